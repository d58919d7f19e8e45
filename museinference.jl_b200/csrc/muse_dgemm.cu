// muse_dgemm.cu — FP64 GEMM on the tensor cores of sm_100a for the dense correlated-Gaussian family (F3).
//
// What it computes.  The only dense contraction on the path: for a batch of latent vectors stored as rows,
//     Q = S · P            (rows × d) · (d × d),   P = Σ₀⁻¹ symmetric,
// i.e. the P·s products inside ∇z logLike = (x − z) − e^{−θ} P z of F3 (in the reference
// this is whatever AD makes of the user's `logLike`, /root/reference/src/simple.jl:85), and W = ξ · Lᵀ for
// sample_x_z (z = e^{θ/2} L ξ, src/simple.jl:61-65 pattern).
//
// How.  tcgen05/TMEM has no f64 kind, so the FP64 tensor path on Blackwell is the warp-level
// `mma.sync.aligned.m8n8k4.row.col.f64` (SASS: DMMA.8x8x4).  CTA tile 128 × 128 × 32, 8 warps as 2 (M) × 4 (N),
// warp tile 64 × 32 = 8 × 4 MMA tiles (64 accumulator doubles per lane); operands staged in shared memory by a
// 3-stage `cp.async` (LDGSTS) pipeline; rows padded (A: BK + 4, B: 132 doubles) so that every fragment load is
// bank-conflict free per half-warp.  All extents are multiples of the tile (the callers allocate padded, zero
// filled operands), so there is no edge handling.  *Bound: FP64 tensor pipe* — 2·M·N·K flop against
// 24·(M·K + K·N + M·N)… bytes, ≈ 400 flop/B at the C5 shape.
#include <cuda.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "muse_common.cuh"

namespace muse {

namespace {

#ifndef MUSE_GEMM_BK
#define MUSE_GEMM_BK 32       // measured (scripts/gemm_variants.sh): BK 32 × 3 stages 32.1 TFLOP/s, 16 × 4: 31.6, 8 × 6: 30.2
#define MUSE_GEMM_STAGES 3
#endif
constexpr int BM = 128, BN = 128, BK = MUSE_GEMM_BK, STAGES = MUSE_GEMM_STAGES;
constexpr int APAD = BK + 4;      // doubles per A row in shared memory: (BK + 4)·2 ≡ 8 (mod 32) banks
constexpr int BPAD = BN + 4;      // 132 doubles per B row
constexpr int A_STAGE = BM * APAD, B_STAGE = BK * BPAD;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// C[M×N] = A[M×K] · B[K×N]; row-major, leading dimensions lda/ldb/ldc; M % 128 == N % 128 == K % 16 == 0.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int K, int lda,
                  int ldb, int ldc) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;           // 2 × 4 warps
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const double* Ag = A + (size_t)m0 * lda;
    const double* Bg = B + n0;

    auto load_stage = [&](int stage, int kt) {
        double* as = As + stage * A_STAGE;
        double* bs = Bs + stage * B_STAGE;
        const int k0 = kt * BK;
        constexpr int ACH = BK / 2;                    // 16-byte chunks per A row
#pragma unroll
        for (int i = 0; i < BM * ACH / GEMM_THREADS; ++i) {        // A: 128 rows × BK/2 chunks of 2 doubles
            const int c = tid + i * GEMM_THREADS;
            const int r = c / ACH, kc = (c % ACH) * 2;
            cp_async16(as + r * APAD + kc, Ag + (size_t)r * lda + k0 + kc);
        }
#pragma unroll
        for (int i = 0; i < BK * (BN / 2) / GEMM_THREADS; ++i) {   // B: BK rows × 64 chunks
            const int c = tid + i * GEMM_THREADS;
            const int r = c >> 6, nc = (c & 63) * 2;
            cp_async16(bs + r * BPAD + nc, Bg + (size_t)(k0 + r) * ldb + nc);
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = K / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    const int ar = wm * 64 + (lane >> 2), ak = lane & 3;      // A fragment: row lane/4, k lane%4
    const int bk = lane & 3, bn = wn * 32 + (lane >> 2);      // B fragment: k lane%4, col lane/4
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kt + STAGES - 1;
        if (nxt < nk) load_stage(nxt % STAGES, nxt);
        cp_async_commit();
        const double* as = As + (kt % STAGES) * A_STAGE;
        const double* bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = as[(ar + i * 8) * APAD + k4 * 4 + ak];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bs[(k4 * 4 + bk) * BPAD + bn + j * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    // epilogue: lane holds C[row lane/4][cols 2·(lane%4), +1] of every 8 × 8 tile
    const int cr = m0 + wm * 64 + (lane >> 2), cc = n0 + wn * 32 + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(C + (size_t)(cr + i * 8) * ldc + cc + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
}

// ---- TMA-fed form (the default) ------------------------------------------------------------------------------------------------
// C[M×N] = A[M×K] · Btᵀ, Bt[N×K] row-major: BOTH operands K-contiguous, so both go through the same machinery:
//   * 2-D tensor maps (cuTensorMapEncodeTiled, FLOAT64, SWIZZLE_128B), box = 16 k × 128 rows = 128-byte rows; one elected lane of a
//     producer warp issues `cp.async.bulk.tensor.2d … mbarrier::complete_tx::bytes` (SASS: UTMALDG) — 2 boxes of A and 2 of Bt per
//     32-deep k-step — into a 3-stage ring of 64 KB stages guarded by full / empty mbarriers; the 8 consumer warps never touch
//     global memory for operands and never meet at a CTA barrier inside the k loop;
//   * fragment loads straight out of the swizzled boxes, conflict-free: the 4 k's of one m8n8k4 step are {2c, 2c+1, 8+2c, 9+2c}
//     of a box (c = 0…3; summing over k in another order than 0,1,2,… is the only arithmetic difference from the cp.async form), so
//     that the 16 lanes of a half-warp hit the 16 distinct 8-byte slots of a 128-byte line after the XOR with (row mod 8);
//   * a persistent grid (one CTA per SM) walks the output tiles, column tile fastest (P / Lᵀ bands stay in the 126 MB L2).
// P = Σ₀⁻¹ is symmetric, so Bt = P itself; W = ξ·Lᵀ takes Bt = L.
constexpr int TK = 16;                                  // k per box (128 bytes)
constexpr int TMA_STAGES = 3;
constexpr int TMA_BOX_BYTES = BM * TK * 8;              // 16 KB
constexpr int TMA_STAGE_BYTES = 4 * TMA_BOX_BYTES;      // A box 0, A box 1, Bt box 0, Bt box 1
constexpr int TMA_THREADS = 288;                        // 8 consumer warps + 1 producer warp
constexpr int TMA_SMEM = TMA_STAGES * TMA_STAGE_BYTES + 1024;

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double* __restrict__ C, int M, int N, int K, int ldc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * TMA_STAGES];
    const uint32_t smem = (sm_u32(smem_raw) + 1023u) & ~1023u;         // 1024-byte aligned: the swizzle pattern repeats every 8 rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full0 = sm_u32(&bars[0]), empty0 = sm_u32(&bars[TMA_STAGES]);
    if (tid == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) { bar_init(full0 + 8 * s, 1); bar_init(empty0 + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tiles_n = N / BN, ntiles = (M / BM) * tiles_n, nk = K / BK;
    if (warp == 8) {
        // ---------------------------------------------------------------- producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
                for (int kt = 0; kt < nk; ++kt) {
                    bar_wait(empty0 + 8 * stage, phase ^ 1u);
                    const uint32_t full = full0 + 8 * stage, dst = smem + stage * TMA_STAGE_BYTES;
                    bar_expect_tx(full, TMA_STAGE_BYTES);
                    const int k0 = kt * BK;
                    tma_load_2d(dst, &tmA, k0, m0, full);
                    tma_load_2d(dst + TMA_BOX_BYTES, &tmA, k0 + TK, m0, full);
                    tma_load_2d(dst + 2 * TMA_BOX_BYTES, &tmB, k0, n0, full);
                    tma_load_2d(dst + 3 * TMA_BOX_BYTES, &tmB, k0 + TK, n0, full);
                    if (++stage == TMA_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }
    // -------------------------------------------------------------------- consumers: 2 (M) × 4 (N) warps, warp tile 64 × 32
    const int wm = warp >> 2, wn = warp & 3;
    const int r8 = lane >> 2, ak = lane & 3;
    // byte offset of this lane's k inside a 128-byte row for the four steps c = 0…3 of a box: k′ = 8·(ak/2) + 2c + ak%2
    uint32_t koff[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) koff[c] = (uint32_t)((((4 * (ak >> 1) + c) ^ r8) << 4) + ((ak & 1) << 3));
    const uint32_t a_row = (uint32_t)((wm * 64 + r8) * 128), b_row = (uint32_t)((wn * 32 + r8) * 128);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kt = 0; kt < nk; ++kt) {
            bar_wait(full0 + 8 * stage, phase);
            const uint32_t st = smem + stage * TMA_STAGE_BYTES;
#pragma unroll
            for (int box = 0; box < 2; ++box) {
                const uint32_t ab = st + box * TMA_BOX_BYTES + a_row, bb = st + (2 + box) * TMA_BOX_BYTES + b_row;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    double a[8], b[4];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = lds_f64(ab + i * 8 * 128 + koff[c]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) b[j] = lds_f64(bb + j * 8 * 128 + koff[c]);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
                }
            }
            __syncwarp();
            if (lane == 0) bar_arrive(empty0 + 8 * stage);
            if (++stage == TMA_STAGES) { stage = 0; phase ^= 1u; }
        }
        // epilogue: lane holds C[row lane/4][cols 2·(lane%4), +1] of every 8 × 8 tile; the producer is already filling the ring
        // for the next tile
        const int cr = m0 + wm * 64 + r8, cc = n0 + wn * 32 + 2 * ak;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<double2*>(C + (size_t)(cr + i * 8) * ldc + cc + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// rows × K doubles, row stride ld: box = TK k's × 128 rows, 128-byte swizzle
bool make_map(CUtensorMap* m, const double* base, int rows, int K, int ld) {
    EncodeTiledFn f = encode_tiled();
    if (!f) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    return f(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// C = A·Btᵀ: A[M×K] (row stride lda), Bt[N×K] (row stride ldbt), C[M×N]; M % 128 == N % 128 == K % 32 == 0, 16-byte aligned rows.
// MUSE_GEMM=cpasync selects the cp.async form above (A/B; it takes Bt as if it were B — the callers' Bt are symmetric or irrelevant
// to a timing).
cudaError_t launch_dgemm(const double* A, const double* Bt, double* C, int M, int N, int K, int lda, int ldbt, int ldc,
                         cudaStream_t st) {
    if (M % BM || N % BN || K % BK) return cudaErrorInvalidValue;
    static const bool legacy = [] { const char* e = std::getenv("MUSE_GEMM"); return e && std::string(e) == "cpasync"; }();
    CUtensorMap ta, tb;
    if (legacy || !make_map(&ta, A, M, K, lda) || !make_map(&tb, Bt, N, K, ldbt)) {
        // per launch, not cached: the attribute belongs to the current device's context
        cudaError_t e = cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        if (e != cudaSuccess) return e;
        dim3 grid(N / BN, M / BM);
        dgemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(A, Bt, C, K, lda, ldbt, ldc);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (M / BM) * (N / BN);
    dgemm_tma_kernel<<<ntiles < sms ? ntiles : sms, TMA_THREADS, TMA_SMEM, st>>>(ta, tb, C, M, N, K, ldc);
    return cudaGetLastError();
}

}  // namespace muse

// ---- diagnostics entry points (tests, bench): the GEMM on host operands, and timed on device operands -----------
extern "C" int muse_b200_dgemm_host(const double* A, const double* B, double* C, int32_t M, int32_t N, int32_t K) {
    if (!A || !B || !C) return MUSE_EINVAL;
    double *dA = nullptr, *dB = nullptr, *dC = nullptr;
    const size_t sa = (size_t)M * K * 8, sb = (size_t)K * N * 8, sc = (size_t)M * N * 8;
    int rc = MUSE_OK;
    if (cudaMalloc(&dA, sa) != cudaSuccess || cudaMalloc(&dB, sb) != cudaSuccess || cudaMalloc(&dC, sc) != cudaSuccess) rc = MUSE_ENOMEM;
    std::vector<double> bt((size_t)N * K);                      // the kernel takes Bt[N×K]
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) bt[(size_t)n * K + k] = B[(size_t)k * N + n];
    if (rc == MUSE_OK && (cudaMemcpy(dA, A, sa, cudaMemcpyHostToDevice) != cudaSuccess ||
                          cudaMemcpy(dB, bt.data(), sb, cudaMemcpyHostToDevice) != cudaSuccess)) rc = MUSE_ECUDA;
    if (rc == MUSE_OK && muse::launch_dgemm(dA, dB, dC, M, N, K, K, K, N, nullptr) != cudaSuccess) rc = MUSE_EINVAL;
    if (rc == MUSE_OK && cudaMemcpy(C, dC, sc, cudaMemcpyDeviceToHost) != cudaSuccess) rc = MUSE_ECUDA;
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    cudaGetLastError();
    return rc;
}

extern "C" int muse_b200_dgemm_time(int32_t M, int32_t N, int32_t K, int32_t reps, double* ms_per_gemm) {
    if (!ms_per_gemm || reps < 1) return MUSE_EINVAL;
    double *dA = nullptr, *dB = nullptr, *dC = nullptr;
    const size_t sa = (size_t)M * K * 8, sb = (size_t)K * N * 8, sc = (size_t)M * N * 8;
    int rc = MUSE_OK;
    if (cudaMalloc(&dA, sa) != cudaSuccess || cudaMalloc(&dB, sb) != cudaSuccess || cudaMalloc(&dC, sc) != cudaSuccess) rc = MUSE_ENOMEM;
    if (rc == MUSE_OK) {
        cudaMemset(dA, 0x3f, sa);      // 0x3f3f… ≈ 4.7e-4: non-trivial operands
        cudaMemset(dB, 0x3f, sb);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 2 && rc == MUSE_OK; ++i)
            if (muse::launch_dgemm(dA, dB, dC, M, N, K, K, K, N, nullptr) != cudaSuccess) rc = MUSE_EINVAL;
        cudaEventRecord(e0);
        for (int i = 0; i < reps && rc == MUSE_OK; ++i)
            if (muse::launch_dgemm(dA, dB, dC, M, N, K, K, K, N, nullptr) != cudaSuccess) rc = MUSE_EINVAL;
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) rc = MUSE_ECUDA;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_per_gemm = ms / reps;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    cudaGetLastError();
    return rc;
}
