"""DMMA GEMM of F3 vs cuBLAS DGEMM (torch.float64 matmul) at the C5 shape (diagnostics)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
lib = m.load_library()
for (M, N, K) in ((8192, 4096, 4096), (4096, 4096, 4096)):
    ms = C.c_double()
    assert lib.muse_b200_dgemm_time(M, N, K, 10, C.byref(ms)) == 0
    a = torch.full((M, K), 4.7e-4, dtype=torch.float64, device="cuda"); b = torch.full((K, N), 4.7e-4, dtype=torch.float64, device="cuda")
    for _ in range(2): torch.matmul(a, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c = torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    cub = e0.elapsed_time(e1) / 10
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: dmma kernel {ms.value:.3f} ms = {fl/ms.value/1e9:.1f} TFLOP/s | cuBLAS {cub:.3f} ms = {fl/cub/1e9:.1f} TFLOP/s | ratio {cub/ms.value:.2f}")
