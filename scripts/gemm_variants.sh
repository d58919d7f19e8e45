#!/bin/bash
# build the GEMM alone in two variants and time them (diagnostics)
cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
for v in "16 4" "32 3" "8 6"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -DMUSE_GEMM_BK=$1 -DMUSE_GEMM_STAGES=$2 \
     -o /tmp/libgemm_$1.so museinference.jl_b200/csrc/muse_dgemm.cu || exit 1
  python - <<PY
import ctypes as C
lib=C.CDLL("/tmp/libgemm_$1.so")
ms=C.c_double()
lib.muse_b200_dgemm_time.argtypes=[C.c_int32]*4+[C.POINTER(C.c_double)]
for M in (8192, 8320):
    assert lib.muse_b200_dgemm_time(M,4096,4096,10,C.byref(ms))==0
    print("BK=$1 stages=$2 M=%d: %.3f ms %.1f TFLOP/s"%(M, ms.value, 2.0*M*4096*4096/ms.value/1e9))
PY
done
