// muse_draws.cu — base normals (ξ, ν) generated on the device.
//
// Replaces the child RNGs of `split_rng` (/root/reference/src/util.jl:85-92) for throughput runs.
// Common-random-number semantics are kept by construction: the normals of sim k are a pure
// function of (seed, global sim index, stream, element), so they are identical at every θ, in
// muse!/get_J!/get_H!, and for any sharding of sims over GPUs.
//
// Generator: Philox4x32-10 (Salmon et al., SC'11) → two 53-bit uniforms → Box–Muller.
//   counter = (pair p, global sim G, stream t, 0), key = (seed lo, seed hi);  t=0: ξ, t=1: ν
//   element 2p = r cos(2πu₂), element 2p+1 = r sin(2πu₂), r = sqrt(-2 ln u₁)
// The master stream's own draw uses G = 0xFFFFFFFF.  oracle/philox.py restates this for tests.
//
// FP64 log/sincospi are ~150 FP64 instructions per pair, so this kernel is compute-bound; it
// runs once per seed (draws are stored and re-read, 16·d B per sim per pass, which is cheaper
// than regenerating them).
#include <cstdlib>

#include "muse_common.cuh"
#include "muse_draw_tables.cuh"
#include "muse_normal_math.cuh"

#ifndef MUSE_DRAWS_IMPL_DEFAULT
#define MUSE_DRAWS_IMPL_DEFAULT 1
#endif

namespace muse {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
    const uint64_t v = ((uint64_t)(hi >> 5) << 26) + (uint64_t)(lo >> 6);
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// Grid: x over pairs of a row, y over (row, stream) — no 64-bit index arithmetic in the loop (the first version
// spent more issue slots on `t / npairs`, `t % npairs` than on the Box–Muller transform: ncu, profiles/README.md).
__global__ void __launch_bounds__(256)
philox_draws_kernel(double* __restrict__ xi, double* __restrict__ nu, int rows, int d, int ld,
                    uint32_t k0, uint32_t k1, int64_t sim_offset, int master_row) {
    const int npairs = (d + 1) >> 1;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= npairs) return;
    for (int rs = blockIdx.y; rs < 2 * rows; rs += gridDim.y) {
        const int stream = rs & 1;
        const int row = rs >> 1;
        const uint32_t G = (row == master_row) ? 0xFFFFFFFFu : (uint32_t)(sim_offset + row);
        uint32_t r[4];
        philox4x32_10((uint32_t)p, G, (uint32_t)stream, 0u, k0, k1, r);
        const double u1 = u53(r[0], r[1]);
        const double u2 = u53(r[2], r[3]);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        double* dst = (stream ? nu : xi) + (size_t)row * ld + 2 * (size_t)p;
        if (2 * p + 1 < d) {
            *reinterpret_cast<double2*>(dst) = make_double2(rad * cs, rad * sn);
        } else {
            dst[0] = rad * cs;
        }
    }
}

// The same generator with the table-driven transform of muse_normal_math.cuh (tables staged in shared memory once per
// CTA): ≈ 55 issue slots for Philox + ≈ 75 for the transform and the store, against ≈ 250 per pair above.
__global__ void __launch_bounds__(256)
philox_draws_tab_kernel(double* __restrict__ xi, double* __restrict__ nu, int rows, int d, int ld,
                        uint32_t k0, uint32_t k1, int64_t sim_offset, int master_row) {
    __shared__ __align__(16) double s_log[91][2];
    __shared__ __align__(16) double s_trig[256][2];
    for (int i = threadIdx.x; i < 91 * 2; i += 256) (&s_log[0][0])[i] = (&kLogTab[0][0])[i];
    for (int i = threadIdx.x; i < 256 * 2; i += 256) (&s_trig[0][0])[i] = (&kTrigTab[0][0])[i];
    __syncthreads();
    const int npairs = (d + 1) >> 1;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= npairs) return;
    for (int rs = blockIdx.y; rs < 2 * rows; rs += gridDim.y) {
        const int stream = rs & 1;
        const int row = rs >> 1;
        const uint32_t G = (row == master_row) ? 0xFFFFFFFFu : (uint32_t)(sim_offset + row);
        uint32_t r[4];
        philox4x32_10((uint32_t)p, G, (uint32_t)stream, 0u, k0, k1, r);
        double n0, n1;
        box_muller_tab(r[0], r[1], r[2], r[3], s_log, s_trig, &n0, &n1);
        double* dst = (stream ? nu : xi) + (size_t)row * ld + 2 * (size_t)p;
        if (2 * p + 1 < d) {
            *reinterpret_cast<double2*>(dst) = make_double2(n0, n1);
        } else {
            dst[0] = n0;
        }
    }
}

// MUSE_DRAWS_IMPL: 1 = table-driven transform (default), 0 = libm log / sincospi (the first implementation, kept for A/B)
static int draws_impl() {
    static const int impl = [] { const char* e = std::getenv("MUSE_DRAWS_IMPL"); return e ? std::atoi(e) : MUSE_DRAWS_IMPL_DEFAULT; }();
    return impl;
}

cudaError_t launch_philox_draws(double* xi, double* nu, int rows, int d, int ld, uint64_t seed,
                                int64_t sim_offset, int master_row, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    const int npairs = (d + 1) >> 1;
    // a thread walks several (row, stream) slots of its pair column: set-up (constants, addressing) is paid once
    const int ny = 2 * rows < 96 ? 2 * rows : 96;
    dim3 grid((unsigned)((npairs + 255) / 256), (unsigned)ny);
    if (draws_impl() == 1)
        philox_draws_tab_kernel<<<grid, 256, 0, st>>>(xi, nu, rows, d, ld, (uint32_t)(seed & 0xFFFFFFFFu),
                                                     (uint32_t)(seed >> 32), sim_offset, master_row);
    else
        philox_draws_kernel<<<grid, 256, 0, st>>>(xi, nu, rows, d, ld, (uint32_t)(seed & 0xFFFFFFFFu),
                                                 (uint32_t)(seed >> 32), sim_offset, master_row);
    return cudaGetLastError();
}

}  // namespace muse
