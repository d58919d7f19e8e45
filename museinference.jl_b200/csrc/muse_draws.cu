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
#define MUSE_DRAWS_IMPL_DEFAULT 2
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

// ---- the same generator and the same arithmetic, instruction count trimmed (the kernel is issue-bound: ncu, profiles/) ------------------
// Against philox_draws_tab_kernel, per pair: the polynomial constants live in registers for the whole row loop instead of being re-read
// from the constant bank, the tables are read with plain ld.shared from 32-bit addresses computed once (a generic pointer makes the
// compiler rebuild the shared window base per access), a thread does both streams of a row back to back (two independent chains in
// flight, one set of addresses), the 53-bit integers are assembled with funnel shifts, and the int → double conversion of the angle
// remainder is an exponent OR plus one exact subtraction.  Every floating-point operation is the one muse_normal_math.cuh spells
// (the host build of that header is the checker), so the output is bit for bit that of philox_draws_tab_kernel.
__device__ __forceinline__ double2 lds_pair(uint32_t addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

struct NMRegs {
    double l0, l1, l2, l3, ln2hi, ln2lo, twopi, s0, s1, s2, c0, c1;
};

__device__ __forceinline__ void box_muller_regs(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t logb, uint32_t trigb,
                                                const NMRegs& K, double* n0, double* n1) {
    // v = ((r1 >> 5) << 26) + (r0 >> 6) as a (hi, lo) word pair
    const uint32_t h1 = r1 >> 5;
    const uint32_t vlo = __funnelshift_r(r0, h1, 6), vhi = h1 >> 6;
    const double y = (double)(((uint64_t)vhi << 32) | vlo) + 0.5;
    const uint32_t yh = (uint32_t)__double2hiint(y), yl = (uint32_t)__double2loint(y);
    const uint32_t mant20 = yh & 0xFFFFFu;
    const bool up = mant20 >= 0x6A09Fu;                                   // m ≥ √2: use m/2 ∈ [√½, 1)
    const int e = (int)(yh >> 20) - (up ? 1075 : 1076);
    const uint32_t mh = mant20 | (up ? 0x3FE00000u : 0x3FF00000u);
    const uint32_t idx = up ? ((mant20 + 8192u) >> 14) - 27u : ((mant20 + 4096u) >> 13) + 37u;
    const double m = __hiloint2double((int)mh, (int)yl);
    const double2 lt = lds_pair(logb + idx * 16u);                        // (1/c, −2 ln c)
    const double r = fma(m, lt.x, -1.0);
    double p = fma(r, K.l0, K.l1);
    p = fma(r, p, K.l2);
    p = fma(r, p, 0.5);
    p = fma(r, p, K.l3);
    p = fma(r, p, 1.0);
    const double P = fma(r * r, p, -2.0 * r);
    const double ed = (double)e;
    const double t = fma(ed, K.ln2hi, lt.y) + fma(ed, K.ln2lo, P);
    const double rad = sqrt(t);

    const uint32_t h3 = r3 >> 5;
    const uint32_t wlo = __funnelshift_r(r2, h3, 6), whi = h3 >> 6;         // v₂ = (whi : wlo), 53 bits
    const uint32_t j = whi >> 13;                                          // v₂ >> 45
    // (v₂ mod 2⁴⁵) − 2⁴⁴ without an integer → double conversion: with the exponent field of 2⁵² the mantissa field IS the integer k
    // (ulp 1), and (2⁵² + k) − (2⁵² + 2⁴⁴) is exact
    const double wk = __hiloint2double((int)((whi & 0x1FFFu) | 0x43300000u), (int)wlo);
    const double bw = ((wk - (0x1p52 + 0x1p44)) + 0.5) * K.twopi;
    const double2 tt = lds_pair(trigb + j * 16u);                          // (cos a_j, sin a_j)
    const double ca = tt.x, sa = tt.y;
    const double b2 = bw * bw;
    double ps = fma(b2, K.s0, K.s1);
    ps = fma(b2, ps, K.s2);
    const double sb = fma(b2 * bw, ps, bw);
    double pc = fma(b2, K.c0, K.c1);
    pc = fma(b2, pc, -0.5);
    const double cm1 = b2 * pc;
    const double cs = ca + fma(ca, cm1, -(sa * sb));
    const double sn = sa + fma(sa, cm1, ca * sb);
    *n0 = rad * cs;
    *n1 = rad * sn;
}

template <bool BOTH>      // BOTH: a thread does the two streams of a row back to back (two chains in flight, 64 registers); else one
__global__ void __launch_bounds__(256)
philox_draws_tab2_kernel(double* __restrict__ xi, double* __restrict__ nu, int rows, int d, int ld,
                         uint32_t k0, uint32_t k1, int64_t sim_offset, int master_row) {
    __shared__ __align__(16) double s_log[91][2];
    __shared__ __align__(16) double s_trig[256][2];
    for (int i = threadIdx.x; i < 91 * 2; i += 256) (&s_log[0][0])[i] = (&kLogTab[0][0])[i];
    for (int i = threadIdx.x; i < 256 * 2; i += 256) (&s_trig[0][0])[i] = (&kTrigTab[0][0])[i];
    __shared__ double s_const[12];
    __shared__ uint32_t s_base[2];
    if (threadIdx.x < 12) s_const[threadIdx.x] = kNM[threadIdx.x];
    if (threadIdx.x == 32) {
        s_base[0] = (uint32_t)__cvta_generic_to_shared(&s_log[0][0]);
        s_base[1] = (uint32_t)__cvta_generic_to_shared(&s_trig[0][0]);
    }
    __syncthreads();
    const int npairs = (d + 1) >> 1;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= npairs) return;
    // ptxas re-reads constant-bank values and rebuilds shared-window addresses inside the loop rather than keep them in registers
    // (≈ 12 issue slots per pair); values that come out of shared memory are not rematerialised
    volatile double* kc = s_const;
    volatile uint32_t* kb = s_base;
    const uint32_t logb = kb[0], trigb = kb[1];
    NMRegs K;
    K.l0 = kc[0]; K.l1 = kc[1]; K.l2 = kc[2]; K.l3 = kc[3];
    K.ln2hi = kc[4]; K.ln2lo = kc[5]; K.twopi = kc[6];
    K.s0 = kc[7]; K.s1 = kc[8]; K.s2 = kc[9]; K.c0 = kc[10]; K.c1 = kc[11];
    const bool whole = 2 * p + 1 < d;
    if (!BOTH) {
        // grid.y walks (row, stream) slots; gridDim.y is even, so a thread stays on one stream
        const uint32_t stream = blockIdx.y & 1u;
        double* dst = (stream ? nu : xi) + (size_t)(blockIdx.y >> 1) * ld + 2 * (size_t)p;
        const size_t step = (size_t)(gridDim.y >> 1) * ld;
        for (int row = blockIdx.y >> 1; row < rows; row += gridDim.y >> 1, dst += step) {
            const uint32_t G = (row == master_row) ? 0xFFFFFFFFu : (uint32_t)(sim_offset + row);
            uint32_t ra[4];
            philox4x32_10((uint32_t)p, G, stream, 0u, k0, k1, ra);
            double a0, a1;
            box_muller_regs(ra[0], ra[1], ra[2], ra[3], logb, trigb, K, &a0, &a1);
            if (whole) *reinterpret_cast<double2*>(dst) = make_double2(a0, a1);
            else dst[0] = a0;
        }
        return;
    }
    const size_t step = (size_t)gridDim.y * ld;
    size_t off = (size_t)blockIdx.y * ld + 2 * (size_t)p;
    for (int row = blockIdx.y; row < rows; row += gridDim.y, off += step) {
        const uint32_t G = (row == master_row) ? 0xFFFFFFFFu : (uint32_t)(sim_offset + row);
        uint32_t ra[4], rb[4];
        philox4x32_10((uint32_t)p, G, 0u, 0u, k0, k1, ra);
        philox4x32_10((uint32_t)p, G, 1u, 0u, k0, k1, rb);
        double a0, a1, b0, b1;
        box_muller_regs(ra[0], ra[1], ra[2], ra[3], logb, trigb, K, &a0, &a1);
        box_muller_regs(rb[0], rb[1], rb[2], rb[3], logb, trigb, K, &b0, &b1);
        if (whole) {
            *reinterpret_cast<double2*>(xi + off) = make_double2(a0, a1);
            *reinterpret_cast<double2*>(nu + off) = make_double2(b0, b1);
        } else {
            xi[off] = a0;
            nu[off] = b0;
        }
    }
}

// MUSE_DRAWS_IMPL: 2 = table-driven transform, trimmed instruction count (default), 1 = the same transform as first written, 0 = libm log / sincospi (the first implementation, kept for A/B)
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
    if (draws_impl() == 2) {
        grid.y = (unsigned)(rows < 48 ? rows : 48);          // a thread does both streams of its rows
        philox_draws_tab2_kernel<true><<<grid, 256, 0, st>>>(xi, nu, rows, d, ld, (uint32_t)(seed & 0xFFFFFFFFu),
                                                            (uint32_t)(seed >> 32), sim_offset, master_row);
    } else if (draws_impl() == 3) {
        grid.y = (unsigned)(rows < 48 ? 2 * rows : 96);
        philox_draws_tab2_kernel<false><<<grid, 256, 0, st>>>(xi, nu, rows, d, ld, (uint32_t)(seed & 0xFFFFFFFFu),
                                                             (uint32_t)(seed >> 32), sim_offset, master_row);
    } else if (draws_impl() == 1)
        philox_draws_tab_kernel<<<grid, 256, 0, st>>>(xi, nu, rows, d, ld, (uint32_t)(seed & 0xFFFFFFFFu),
                                                     (uint32_t)(seed >> 32), sim_offset, master_row);
    else
        philox_draws_kernel<<<grid, 256, 0, st>>>(xi, nu, rows, d, ld, (uint32_t)(seed & 0xFFFFFFFFu),
                                                 (uint32_t)(seed >> 32), sim_offset, master_row);
    return cudaGetLastError();
}

}  // namespace muse
