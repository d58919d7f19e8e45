// muse_iso_stream.cu — single-pass, bulk-async (TMA) streaming MAP + score kernel for the
// isotropic-Gaussian-latent families (F1 funnel, F2 hierarchical Gaussian) at large latent
// dimension, sm_100a.  First pass of every solver launch when d ≥ 4096 (muse_api.cu); units it
// cannot finish are handed to the generic kernel (muse_iso_solver.cu) on the same stream.
//
// What it replaces: the body the reference maps over its worker pool
// (/root/reference/src/muse.jl:170-175, :510-513, :430-432): sample_x_z → ẑ_at_θ (Optim L-BFGS +
// HagerZhang, src/interface.jl:162-166) → ∇θ_logLike (src/simple.jl:92).
//
// Why one pass is enough.  For these families ∇²_z(−logLike) = (1+a)·I, so along the first L-BFGS
// direction s = −∇f(z₀) the line minimum is at c* = 1/(1+a) *whatever the data*.  The reference
// algorithm gets there in three value+gradient evaluations: z₀, the InitialStatic trial z₀+s, and
// the secant point z₀ + c·s with c = φ'(0)/(φ'(0) − φ'(1)) = c* up to round-off.  All three are
// elementwise in (x_j, z₀_j) plus global sums, and only c depends on those sums.  So one sweep
// evaluates the first two honestly and the third *speculatively* at c_spec = 1/(1+a), committing
// ẑ = z₀ + c_spec·s as it goes.  The unit's 15 sums are then run through the scalar optimiser's decisions
// (fast_replay below: the same tests, in the same order, as Controller::solve / hager_zhang in muse_iso_ctl.cuh,
// the code the generic kernel executes): initial convergence, φ(1) and the bracket [0, 1], the secant step c, and —
// iff |c − c_spec| ≤ 1e-11·c_spec — the Wolfe and convergence tests on the speculative sums.  Anything else (another
// bracket, a rejected trial, a second iteration, a non-finite value after the step, a start vector that had to be
// kept for a 0-iteration solve …) hands the unit back: nothing of it is published and its index goes to a
// device-side list that the generic two-sweep kernel, launched right behind, re-solves from the untouched start.
// Parity is unaffected: ẑ differs from the non-speculative result by ≤ 1e-11 relative (tests: rtol 1e-8), iteration
// and evaluation counts are those of the algorithm.
//
// HBM traffic per unit = the fused floor: read ξ, ν [, z₀], write ẑ (nothing for finite-difference
// virtual sims, whose ẑ the reference discards too).  x is never materialised.
//
// Structure.  One CTA per SM, three roles:
//   warp 0, one lane   producer: walks this CTA's work items — (unit, segment) pairs, round-robin
//                      over CTAs — and streams their rows chunk by chunk into a ring of shared-memory
//                      stages with cp.async.bulk + mbarrier complete_tx, running ahead across item
//                      boundaries (bytes in flight per SM = the ring, ~190 KB, independent of registers)
//   warp 1             finisher: sums the consumer warps' partials in index order, publishes the
//                      segment's partial to global memory; the CTA that publishes a unit's last
//                      segment sums the segments in index order and replays the optimiser's
//                      decisions on them (one lane, ≈ 100 FP64 instructions), then writes the outputs
//   warps 2..17        consumers: wait full[stage], fused elementwise work out of shared memory,
//                      128-bit coalesced stores of ẑ, arrive empty[stage]; at the end of an item a
//                      shuffle tree reduces the 15 running sums, which go to the finisher through a
//                      two-slot mailbox — consumers never wait for the scalar code.
// Every reduction is a fixed tree (thread → warp butterfly → warps in order → segments in order),
// so results do not depend on scheduling, grid size or how sims are sharded over GPUs.
#include "muse_iso_ctl.cuh"
#include "muse_outer_dev.cuh"

namespace muse {

namespace {

constexpr int kNC = 512;               // consumer threads
constexpr int kNCW = kNC / 32;         // consumer warps
constexpr int kThreads = kNC + 64;     // + producer warp + finisher warp
constexpr int kChunk = 2048;           // elements per row per stage (16 KB)
constexpr int kMaxStages = 8;
constexpr int kDescRing = 16;
// [host-test:begin red-slots]  (tests/test_fast_replay.py compiles the marked blocks for the host)
constexpr int kNRed = 15;              // running sums per unit
constexpr int kRedPad = 16;
constexpr double kSpecTol = 1e-11;

// reduction slots; kMaxMask marks the max-reductions
enum Red : int {
    rR0 = 0, rS2_0, rS1_0, rGG0,                // sums at z₀
    rR1, rS2_1, rDP1,                           // sums at z₀ + s
    rRT, rS2T, rS1T, rDPT, rGGT,                // sums at z₀ + c_spec·s
    rGM0, rGMT, rXC,                            // maxima: ‖∇f(z₀)‖∞, ‖∇f(z₀ + c_spec·s)‖∞, max|Δz|
};
constexpr int kNSum = 12;                       // slots [0, kNSum) are sums, [kNSum, kNRed) maxima of non-negative values
constexpr unsigned kMaxMask = (1u << rGM0) | (1u << rGMT) | (1u << rXC);
static_assert(rGM0 == kNSum && rXC == kNRed - 1, "reduction slot layout");
// [host-test:end red-slots]

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// [host-test:begin item-desc]
// what the producer tells the other roles about one work item
struct ItemDesc {
    double* zout;        // where ẑ goes (null: discarded)
    double sig, mus;     // sample_x_z constants of the unit (z = mus + sig·ξ)
    int unit;            // launch item
    int seg;
    int chunk0, nch;
    int sim;             // rows 0,1 = ξ,ν (else row 0 = x)
    int zk;              // 0: z₀ ≡ 0, 1: z₀ streamed in row 2, 2: z₀ = simulated latent (truth)
    int start_kind;      // StartKind of the unit
    int zst_accept;      // ZState of the unit once the committed step is accepted (buffer that holds ẑ)
    int* zstate_row;     // the unit's zstate cell (null: none)
    unsigned levmask;    // lazy ẑ: bit l set ⇔ the unit's solve of level l took its step (a 0-iteration solve keeps z)
};
// [host-test:end item-desc]

struct Shared {
    uint64_t full[kMaxStages], empty[kMaxStages];
    uint64_t part_full[2], part_empty[2];
    ItemDesc desc[kDescRing];
    double part[2][kNCW][kRedPad];
};

// [host-test:begin elem3]
struct Acc {
    double v[kNRed];
};

// one element: the three evaluations of the fast path (see the file comment)
template <bool SIM, int ZK, bool LEAN = false>
__device__ __forceinline__ double elem3(double p, double q, double z0in, const IsoEval& ev, double sig, double mus,
                                        Acc& A) {
    double x, z0;
    if (SIM) {
        const double zt = fma(sig, p, mus);     // simulated latent            (src/simple.jl:62)
        x = zt + q;                             // simulated data              (src/simple.jl:63)
        z0 = (ZK == 2) ? zt : z0in;
    } else {
        x = p;
        z0 = z0in;
    }
    // f, ∇f at z₀
    const double r0 = x - z0, w0 = z0 - ev.mu;
    const double g0 = fma(ev.a, w0, -r0);
    A.v[rR0] = fma(r0, r0, A.v[rR0]);
    A.v[rS2_0] = fma(w0, w0, A.v[rS2_0]);
    A.v[rS1_0] += w0;
    A.v[rGG0] = fma(g0, g0, A.v[rGG0]);
    A.v[rGM0] = fmax(A.v[rGM0], fabs(g0));
    // φ(1), φ'(1) along s = −∇f(z₀)   (LEAN: left to fast_replay's closed forms, see SolveLaunch::lean)
    if (!LEAN) {
        const double z1 = z0 - g0;
        const double r1 = x - z1, w1 = z1 - ev.mu;
        const double g1 = fma(ev.a, w1, -r1);
        A.v[rR1] = fma(r1, r1, A.v[rR1]);
        A.v[rS2_1] = fma(w1, w1, A.v[rS2_1]);
        A.v[rDP1] = fma(g1, -g0, A.v[rDP1]);
    }
    // speculated committed trial at c_spec
    const double zt = fma(ev.cspec, -g0, z0);
    const double rt = x - zt, wt = zt - ev.mu;
    const double gt = fma(ev.a, wt, -rt);
    A.v[rRT] = fma(rt, rt, A.v[rRT]);
    A.v[rS2T] = fma(wt, wt, A.v[rS2T]);
    A.v[rS1T] += wt;
    A.v[rDPT] = fma(gt, -g0, A.v[rDPT]);
    A.v[rGGT] = fma(gt, gt, A.v[rGGT]);
    A.v[rGMT] = fmax(A.v[rGMT], fabs(gt));
    if (!LEAN) A.v[rXC] = fmax(A.v[rXC], fabs(zt - z0));
    return zt;
}
// [host-test:end elem3]

// lazy ẑ: the unit's current ẑ from its base normals — the committed points of the earlier passes, replayed with the very
// operations elem3 used to produce them (bit for bit what the chain of launches would have stored and read back).
// The levels' constants sit in shared memory and are read with ld.shared through a pure (non-volatile) asm: the compiler is free to
// keep them in registers across elements or to reload them — a generic load per element is what it emitted for a plain pointer.
__device__ __forceinline__ double lds_f64(uint32_t saddr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
    return v;
}
template <bool SIM>
__device__ __forceinline__ double lazy_level(double p, double q, double z, uint32_t lev) {     // lev: shared address of a LazyLevel
    const double sig = lds_f64(lev), mus = lds_f64(lev + 8), a = lds_f64(lev + 16), mu = lds_f64(lev + 24), cspec = lds_f64(lev + 32);
    const double x = SIM ? fma(sig, p, mus) + q : p;
    const double r0 = x - z, w0 = z - mu;
    const double g0 = fma(a, w0, -r0);
    return fma(cspec, -g0, z);
}
static_assert(sizeof(LazyLevel) == 40 && offsetof(LazyLevel, cspec) == 32, "LazyLevel layout (lazy_level reads it field by field)");

// ---- the funnel (F1) specialised: μ = μ_s = 0 -------------------------------------------------------------------------------
// The same VALUES as elem3<…, LEAN> / lazy_level bit for bit — z − 0 = z, fma(a, 0, −x) = −x, fma(c, x, 0) = c·x, Σ∇f(0)² = Σx² are
// exact identities, the sums are accumulated in the same order, Σ(z − μ) is not needed by the funnel's score — in 12 (cold) /
// 16 (+3 or 5 per lazy level) FP64 instructions per element instead of 20 / 20 (+6): the passes are bound by the FP64 pipe (DESIGN.md
// §3.6).  __dmul_rn / __dadd_rn keep the compiler from contracting what elem3 rounds separately.
template <bool SIM, bool ZERO>   // ZERO: z₀ ≡ 0
__device__ __forceinline__ double elem3_funnel(double p, double q, double z0, const IsoEval& ev, double sig, Acc& A) {
    const double x = SIM ? __dadd_rn(__dmul_rn(sig, p), q) : p;
    double g0;
    if (ZERO) {
        g0 = -x;                                            // ∇f(0) = −x
        A.v[rR0] = fma(x, x, A.v[rR0]);
        A.v[rGG0] = A.v[rR0];                               // Σ∇f(0)² is the same sequence of operations as Σ(x − 0)²
    } else {
        const double r0 = x - z0;
        g0 = fma(ev.a, z0, -r0);
        A.v[rR0] = fma(r0, r0, A.v[rR0]);
        A.v[rS2_0] = fma(z0, z0, A.v[rS2_0]);
        A.v[rGG0] = fma(g0, g0, A.v[rGG0]);
    }
    A.v[rGM0] = fmax(A.v[rGM0], fabs(g0));
    const double zt = ZERO ? __dmul_rn(ev.cspec, x) : fma(ev.cspec, -g0, z0);
    const double rt = x - zt;
    const double gt = fma(ev.a, zt, -rt);
    A.v[rRT] = fma(rt, rt, A.v[rRT]);
    A.v[rS2T] = fma(zt, zt, A.v[rS2T]);
    A.v[rDPT] = fma(gt, -g0, A.v[rDPT]);
    A.v[rGGT] = fma(gt, gt, A.v[rGGT]);
    A.v[rGMT] = fmax(A.v[rGMT], fabs(gt));
    return zt;
}
template <bool SIM, bool ZERO>
__device__ __forceinline__ double lazy_level_funnel(double p, double q, double z, uint32_t lev) {
    const double sig = lds_f64(lev), a = lds_f64(lev + 16), cspec = lds_f64(lev + 32);
    const double x = SIM ? __dadd_rn(__dmul_rn(sig, p), q) : p;
    if (ZERO) return __dmul_rn(cspec, x);
    const double r0 = x - z;
    const double g0 = fma(a, z, -r0);
    return fma(cspec, -g0, z);
}
// what a consumer needs of the launch's lazy description
struct LazyView {
    uint32_t lev0;           // shared address of level 0
    int nlev;
};
template <bool SIM, bool FUNNEL, bool FROM_ZERO>
__device__ __forceinline__ double lazy_z0(double p, double q, double zstart, const LazyView& LV, unsigned mask) {
    // no branches: a level the unit did not step at (a 0-iteration solve — rare) is evaluated and dropped by a select, so that the
    // elements of a thread stay independent instruction streams
    double z = zstart;
    if (LV.nlev >= 1) {                      // the hot case: the second pass of a solve
        const double zn = FUNNEL ? lazy_level_funnel<SIM, FROM_ZERO>(p, q, z, LV.lev0) : lazy_level<SIM>(p, q, z, LV.lev0);
        z = (mask & 1u) ? zn : z;
    }
    for (int l = 1; l < LV.nlev; ++l) {
        const uint32_t lev = LV.lev0 + (uint32_t)l * (uint32_t)sizeof(LazyLevel);
        const double zn = FUNNEL ? lazy_level_funnel<SIM, false>(p, q, z, lev) : lazy_level<SIM>(p, q, z, lev);
        z = ((mask >> l) & 1u) ? zn : z;
    }
    return z;
}
// what a consumer keeps of an item's descriptor
struct ItemRegs {
    double sig, mus;
    double* zout;
    unsigned mask;
};
// ZK: 0 z₀ ≡ 0 · 1 z₀ streamed · 2 z₀ = simulated latent · 3 lazy from zero · 4 lazy from a streamed start row
// LEAN: 0 every trial element by element · 1 lean α = 1 trial · 2 lean + the funnel's specialised element code
template <bool SIM, int ZK, int LEAN>
__device__ __forceinline__ double elem_zk(double p, double q, double z0in, const IsoEval& ev, const ItemRegs& R, const LazyView& LV, Acc& A) {
    if constexpr (LEAN == 2 && ZK != 2) {
        if (ZK == 0) return elem3_funnel<SIM, true>(p, q, 0.0, ev, R.sig, A);
        if (ZK == 1) return elem3_funnel<SIM, false>(p, q, z0in, ev, R.sig, A);
        if (ZK == 3) return elem3_funnel<SIM, false>(p, q, lazy_z0<SIM, true, true>(p, q, 0.0, LV, R.mask), ev, R.sig, A);
        return elem3_funnel<SIM, false>(p, q, lazy_z0<SIM, true, false>(p, q, z0in, LV, R.mask), ev, R.sig, A);
    } else {
        if (ZK >= 3) return elem3<SIM, 1, (LEAN != 0)>(p, q, lazy_z0<SIM, false, false>(p, q, ZK == 4 ? z0in : 0.0, LV, R.mask), ev, R.sig, R.mus, A);
        return elem3<SIM, (ZK >= 3 ? 1 : ZK), (LEAN != 0)>(p, q, z0in, ev, R.sig, R.mus, A);
    }
}
__device__ __forceinline__ LazyView lazy_view(const SolveLaunch& L) {
    LazyView LV;
    LV.lev0 = L.lazy ? smem_u32(&L.lazy->lev[0]) : 0u;
    LV.nlev = L.lazy ? L.lazy->nlev : 0;
    return LV;
}

__device__ __forceinline__ void st2_stream(double* p, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

// one chunk of one item, executed by the consumer threads
template <bool SIM, int ZK, int LEAN>
__device__ __forceinline__ void consume_chunk(const double* buf, int base, int len, int d, const ItemRegs& R,
                                              const IsoEval& ev, int ct, uint64_t pol, const LazyView& LV, Acc& A) {
    constexpr bool ZROW = (ZK == 1 || ZK == 4);
    constexpr int U = kChunk / (2 * kNC);
    const double* ra = buf;
    const double* rb = buf + kChunk;
    const double* rz = buf + 2 * kChunk;
    double* zout = R.zout;
    if (len == kChunk && base + kChunk <= d) {
        // a full chunk (all but a row's last): every load first, then the thread's 2·U elements as independent instruction streams
        double2 p[U], q[U], z[U], zt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q2 = 2 * (ct + u * kNC);
            p[u] = lds2(ra + q2);
            q[u] = SIM ? lds2(rb + q2) : make_double2(0.0, 0.0);
            z[u] = ZROW ? lds2(rz + q2) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            zt[u].x = elem_zk<SIM, ZK, LEAN>(p[u].x, q[u].x, z[u].x, ev, R, LV, A);
            zt[u].y = elem_zk<SIM, ZK, LEAN>(p[u].y, q[u].y, z[u].y, ev, R, LV, A);
        }
        if (zout) {
#pragma unroll
            for (int u = 0; u < U; ++u) st2_stream(zout + base + 2 * (ct + u * kNC), zt[u], pol);
        }
    } else {   // the row's last chunk: it may be short, and elements ≥ d are padding
        for (int q2 = 2 * ct; q2 < len; q2 += 2 * kNC) {
            const int j = base + q2;
            if (j >= d) break;
            const double2 p = lds2(ra + q2);
            const double2 q = SIM ? lds2(rb + q2) : make_double2(0.0, 0.0);
            const double2 z = ZROW ? lds2(rz + q2) : make_double2(0.0, 0.0);
            double2 zt;
            zt.x = elem_zk<SIM, ZK, LEAN>(p.x, q.x, z.x, ev, R, LV, A);
            zt.y = 0.0;
            if (j + 1 < d) zt.y = elem_zk<SIM, ZK, LEAN>(p.y, q.y, z.y, ev, R, LV, A);
            if (zout) {
                if (j + 1 < d) st2_stream(zout + j, zt, pol);
                else zout[j] = zt.x;
            }
        }
    }
}

// Warp reduction of the 15 running sums.  Sums: the xor butterfly (16, 8, 4, 2, 1) done "transposed" — at every level
// a lane keeps half of its slots and ships the other half — 16 shuffles instead of 60, same association tree; slot k
// ends in lane 2k (`sum_k` of lane 2k).  Maxima (non-negative values): two integer REDUX over the high and low words,
// result in every lane.
__device__ __forceinline__ void warp_reduce(const Acc& A, int lane, double& sum_k, double (&mv)[kNRed - kNSum]) {
    double sv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) sv[k] = k < kNSum ? A.v[k] : 0.0;
#pragma unroll
    for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const double send = up ? sv[k] : sv[k + half];
            const double keep = up ? sv[k + half] : sv[k];
            sv[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    sum_k = sv[0] + __shfl_xor_sync(0xffffffffu, sv[0], 1);
#pragma unroll
    for (int k = 0; k < kNRed - kNSum; ++k) {
        const unsigned hi = (unsigned)__double2hiint(A.v[kNSum + k]), lo = (unsigned)__double2loint(A.v[kNSum + k]);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        mv[k] = __hiloint2double((int)mh, (int)ml);
    }
}

// ---- the scalar optimiser on the fast path, replayed on a unit's 15 sums (one thread) ------------------------------
// Follows, decision by decision, what Controller::solve / hager_zhang (muse_iso_ctl.cuh — the code the generic kernel
// runs) do when the first L-BFGS iteration ends the solve:
//   initial_convergence           non-finite f → stop; ‖∇f(z₀)‖∞ ≤ atol → 0 iterations
//   HagerZhang, c = 1             φ(1), φ′(1) finite;  φ′(1) ≥ 0 ⇒ bracket [0, 1]                       (B0)
//   secant²                       c = (a·φ′_b − b·φ′_a)/(φ′_b − φ′_a) with (a, b) = (0, 1); the committed trial at c is the
//                                 speculated one iff |c − c_spec| ≤ kSpecTol·c_spec; it must satisfy the (approximate) Wolfe conditions
//   assess_convergence            x unchanged or ‖∇f‖∞ ≤ atol ⇒ done: 1 iteration, 3 evaluations
// Anything else — another bracket, a rejected secant point, a second iteration, an iteration cap of 0, a non-finite
// value after the step, a 0-iteration solve whose start vector was not materialised — returns false: the unit is
// left untouched and the generic kernel solves it from scratch.
// [host-test:begin fast-replay]
struct FastResult {
    double f, gmax, s1, s2;
    int iters, fg, status;
    bool flip;
};

__device__ __noinline__ bool fast_replay(const SolveLaunch& L, int start_kind, const double (&t)[kNRed], FastResult& r) {
    const IsoEval ev = launch_ev(L);
    constexpr double delta = 0.1, sigma = 0.9, epsilon = 1e-6;
    const double f0 = fma(0.5, fma(ev.a, t[rS2_0], t[rR0]), ev.half_cst);
    const double gg0 = t[rGG0], gmax0 = t[rGM0];
    r.f = f0; r.gmax = gmax0; r.s1 = t[rS1_0]; r.s2 = t[rS2_0];
    r.iters = 0; r.fg = 1; r.flip = false;
    const bool keep_start = (start_kind == kStartTruth || start_kind == kStartSharedKeep);
    if (!fin(f0) || !fin(gg0)) { r.status = MUSE_STATUS_NONFINITE; return !keep_start; }
    if (gmax0 <= L.atol) { r.status = MUSE_STATUS_G_CONVERGED; return !keep_start; }
    if (L.max_iters < 1) return false;
    // hager_zhang(c = 1, φ₀ = f, φ′₀ = −‖∇f‖²)
    const double phi_0 = f0, dphi_0 = -gg0;
    if (dphi_0 >= 0.0 || dphi_0 >= kEpsD * fabs(phi_0)) return false;
    const double phi_lim = phi_0 + epsilon * fabs(phi_0);
    const bool lean = L.lean != 0;
    // lean: ∇f(z₀ + s) = −a·∇f(z₀) for these families, hence φ′(1) = a‖∇f(z₀)‖² and φ(1) = φ(0) + φ′(0) + ½(1 + a)‖∇f(z₀)‖²
    const double phi1 = lean ? phi_0 + dphi_0 + 0.5 * (1.0 + ev.a) * gg0 : fma(0.5, fma(ev.a, t[rS2_1], t[rR1]), ev.half_cst);
    const double dphi1 = lean ? ev.a * gg0 : t[rDP1];
    if (!(fin(phi1) && fin(dphi1))) return false;
    if (!(dphi1 >= 0.0)) return false;                               // B0: bracket (a, b) = (0, 1)
    const double c = (0.0 * dphi1 - 1.0 * dphi_0) / (dphi1 - dphi_0);    // secant(a, b)
    if (!(fabs(c - ev.cspec) <= kSpecTol * ev.cspec)) return false;  // the speculated trial is the one asked for
    const double phic = fma(0.5, fma(ev.a, t[rS2T], t[rRT]), ev.half_cst), dphic = t[rDPT];
    const bool w1 = (delta * dphi_0 >= (phic - phi_0) / c) && (dphic >= sigma * dphi_0);
    const bool w2 = ((2 * delta - 1) * dphi_0 >= dphic) && (dphic >= sigma * dphi_0) && (phic <= phi_lim);
    if (!(w1 || w2)) return false;
    // the step is taken; assess_convergence
    const double gg = t[rGGT], gmax = t[rGMT];
    if (!fin(phic) || !fin(gg)) return false;
    const bool x_conv = !lean && t[rXC] <= 0.0, g_conv = gmax <= L.atol;     // lean: max|Δz| is not tracked — an exactly unchanged x is handed back
    if (!(x_conv || g_conv)) return false;                           // a second iteration would follow
    r.f = phic; r.gmax = gmax; r.s1 = t[rS1T]; r.s2 = t[rS2T];
    r.iters = 1; r.fg = 3; r.flip = true;
    r.status = g_conv ? MUSE_STATUS_G_CONVERGED : MUSE_STATUS_XF_CONVERGED;
    return true;
}
// [host-test:end fast-replay]

// [host-test:begin publish]
__device__ __forceinline__ void publish_unit(const SolveLaunch& L, const ItemDesc& it, const double (&t)[kNRed]) {
    FastResult r;
    if (!fast_replay(L, it.start_kind, t, r)) {
        L.redo_items[atomicAdd(L.redo_count, 1)] = it.unit;          // hand back to the generic kernel
        atomicAdd(L.redo_total, 1ULL);
        return;
    }
    const IsoEval ev = launch_ev(L);
    double* g = L.g_out + (size_t)it.unit * L.ntheta;
    if (L.family == MUSE_FAMILY_FUNNEL) {
        g[0] = 0.5 * ev.a * r.s2 - 0.5 * (double)L.d;                // ∇θ logLike = ½ e^{−θ} Σz² − d/2   (src/simple.jl:66-68)
    } else {
        g[0] = ev.a * r.s1;                                          // (e^{−2ℓ} Σ(z−μ), e^{−2ℓ} Σ(z−μ)² − d)
        g[1] = ev.a * r.s2 - (double)L.d;
    }
    L.iters_out[it.unit] = r.iters;
    L.fg_out[it.unit] = r.fg;
    L.gnorm_out[it.unit] = r.gmax;
    L.f_out[it.unit] = r.f;
    L.status_out[it.unit] = r.status;
    if (it.zstate_row) {
        if (r.flip) *it.zstate_row = it.zst_accept;
        // 0 iterations from zero(z): the unit's ẑ IS zero — say so, or a buffer left by an earlier solve would pass for it
        // (found by the host-side fuzz, tests/test_generic_solver_host.py; the generic kernel always did this)
        else if (it.start_kind == kStartZero) *it.zstate_row = kZZero;
    }
}

// unit → pointers (the Controller's own set-up code, muse_iso_ctl.cuh); no sweeps are ever issued through it here
struct NoIssuer {
    __device__ void operator()(Cmd&, double (&)[7]) {}
};
struct WarpCtx {
    int tid;
};
// [host-test:end publish]

// A consumer's view of the ring: where it stands, and the loop over an item's chunks with ONE variant of the per-element code
// (the variant is chosen per item, not per chunk: the loop body stays small and resident in the instruction cache).
struct RingPos {
    int stage;
    uint32_t phase;
};
template <bool SIM, int ZK, int LEAN>
__device__ __forceinline__ void consume_item(const SolveLaunch& L, Shared& sh, const double* ring, int rows, int stages, int chunk0, int nch,
                                             const ItemRegs& R, const IsoEval& ev, int ct, int lane, uint64_t pol, const LazyView& LV,
                                             RingPos& rp, Acc& A) {
    for (int k = 0; k < nch; ++k) {
        if (k) mbar_wait(&sh.full[rp.stage], rp.phase);         // the first chunk was waited for by the caller (descriptor hand-over)
        const double* buf = ring + (size_t)rp.stage * rows * kChunk;
        const int base = (chunk0 + k) * kChunk;
        consume_chunk<SIM, ZK, LEAN>(buf, base, min(kChunk, L.ld - base), L.d, R, ev, ct, pol, LV, A);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty[rp.stage]);
        if (++rp.stage == stages) { rp.stage = 0; rp.phase ^= 1u; }
    }
}
template <int LEAN>
__device__ __forceinline__ void consume_item_l(const SolveLaunch& L, Shared& sh, const double* ring, int rows, int stages, const ItemDesc& it,
                                               const ItemRegs& R, const IsoEval& ev, int ct, int lane, uint64_t pol, const LazyView& LV,
                                               RingPos& rp, Acc& A) {
#define MUSE_ITEM(SIM_, ZK_) consume_item<SIM_, ZK_, LEAN>(L, sh, ring, rows, stages, it.chunk0, it.nch, R, ev, ct, lane, pol, LV, rp, A)
    if (it.sim) {
        if (it.zk == 0) MUSE_ITEM(true, 0);
        else if (it.zk == 1) MUSE_ITEM(true, 1);
        else if (it.zk == 2) MUSE_ITEM(true, 2);
        else if (it.zk == 3) MUSE_ITEM(true, 3);
        else MUSE_ITEM(true, 4);
    } else {
        if (it.zk == 1) MUSE_ITEM(false, 1);
        else if (it.zk == 3) MUSE_ITEM(false, 3);
        else if (it.zk == 4) MUSE_ITEM(false, 4);
        else MUSE_ITEM(false, 0);
    }
#undef MUSE_ITEM
}
// lean evaluation and lazy ẑ exist in solve_persist_kernel only (LEAN is a parameter of that kernel): the single-pass kernels of the
// chain of launches keep their variants
template <bool PERSIST, int LEAN>
__device__ __forceinline__ void consume_item_any(const SolveLaunch& L, Shared& sh, const double* ring, int rows, int stages, const ItemDesc& it,
                                                 const ItemRegs& R, const IsoEval& ev, int ct, int lane, uint64_t pol, const LazyView& LV,
                                                 RingPos& rp, Acc& A) {
    if constexpr (PERSIST) {
        consume_item_l<LEAN>(L, sh, ring, rows, stages, it, R, ev, ct, lane, pol, LV, rp, A);
    } else {
#define MUSE_ITEM(SIM_, ZK_) consume_item<SIM_, ZK_, false>(L, sh, ring, rows, stages, it.chunk0, it.nch, R, ev, ct, lane, pol, LV, rp, A)
        if (it.sim) {
            if (it.zk == 0) MUSE_ITEM(true, 0);
            else if (it.zk == 1) MUSE_ITEM(true, 1);
            else MUSE_ITEM(true, 2);
        } else {
            if (it.zk == 1) MUSE_ITEM(false, 1);
            else MUSE_ITEM(false, 0);
        }
#undef MUSE_ITEM
    }
}

// lazy ẑ: what changes in a unit's descriptor — no start row unless the user gave one, no ẑ store unless this pass
// materialises it, the level mask out of the unit's state cell
__device__ __forceinline__ void lazy_item(const SolveLaunch& L, const Cmd& c, int* zs, const double* zshared, ItemDesc& it, const double*& rz) {
    const LazyLevels& LZ = *L.lazy;
    it.levmask = LZ.nlev ? (unsigned)ld_state(zs) : 0u;
    it.zk = L.zrows ? 4 : (LZ.nlev ? 3 : 0);             // the launch's first pass from zero(z) is simply a cold item
    rz = L.zrows ? zshared : nullptr;
    it.zout = LZ.store ? c.zA : nullptr;
    it.zst_accept = kZA;
    it.zstate_row = zs;
    // a pass that materialises ẑ cannot serve a 0-iteration solve (the committed point is written as it goes): treated like a
    // start that has to be kept — handed back
    it.start_kind = LZ.store ? (int)kStartSharedKeep : (LZ.nlev ? (int)kStartOwn : c.start_kind);
}
// after publish_unit (same thread): the unit's level mask for the next pass (its iteration count says whether it stepped)
__device__ __forceinline__ void lazy_after_publish(const SolveLaunch& L, const ItemDesc& it) {
    const LazyLevels& LZ = *L.lazy;
    if (LZ.store || !it.zstate_row) return;
    const unsigned stepped = L.iters_out[it.unit] > 0 ? 1u : 0u;
    *it.zstate_row = (int)(it.levmask | (stepped << LZ.nlev));
}

// One streaming pass over the launch's units by the three roles of one CTA (file comment).  The body of iso_stream_kernel, and
// of every phase of solve_persist_kernel (PERSIST: the barriers of the previous phase are invalidated and set up again, the
// proxies are fenced around the phase — ẑ written with ordinary stores by one phase is read by bulk copies in the next —
// and the CTA ends the phase together).
template <bool PERSIST, int LEAN = false>
__device__ __forceinline__ void stream_pass(const SolveLaunch& L, Shared& sh, double* const ring, bool reinit) {
    const int rows = L.zrows ? 3 : 2;
    const int stages = L.stream_stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if (PERSIST && reinit) {
            for (int s = 0; s < kMaxStages; ++s) { mbar_inval(&sh.full[s]); mbar_inval(&sh.empty[s]); }
            for (int s = 0; s < 2; ++s) { mbar_inval(&sh.part_full[s]); mbar_inval(&sh.part_empty[s]); }
        }
        for (int s = 0; s < (PERSIST ? kMaxStages : stages); ++s) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], kNCW);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sh.part_full[s], kNCW);
            mbar_init(&sh.part_empty[s], 1);
        }
        fence_mbar_init();
    }
    if (PERSIST) fence_proxy_async();       // what earlier phases stored (any CTA, ordered by the phase barrier) → this phase's bulk copies
    __syncthreads();

    // diagnostics (muse_b200_debug_timeline): per CTA [start ns, end ns, smid, producer done ns, finisher done ns, consumers done ns]
    long long* const dbg = L.dbg ? L.dbg + (size_t)blockIdx.x * 16 : nullptr;
    auto now_ns = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
    if (dbg && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        dbg[0] = now_ns();
        dbg[2] = smid;
    }

    const int nchunks = (L.ld + kChunk - 1) / kChunk;
    const int nseg = L.nseg;
    const int total = L.nitems * nseg;

    // Work items (unit, segment) are handed out dynamically (one global counter, zeroed by the host before the
    // launch): SMs differ in speed by ±15 % on this workload, a static split leaves the slow ones as a tail.
    // The producer is the only role that talks to the counter; the others learn each item — and the end of work,
    // an item with unit = −1 — from the descriptor ring.
    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const L2Policy pol = make_policies();
            WarpCtx ctx{0};
            NoIssuer none;
            Controller<WarpCtx, NoIssuer> u(ctx, L, none);        // only for setup_unit (pointer logic)
            const double* zshared = u.resolve_zshared();
            const uint64_t zpol = (L.start_kind == kStartShared || L.start_kind == kStartSharedKeep) ? pol.last : pol.first;
            int stage = 0;
            uint32_t phase = 0;
            int w = atomicAdd(L.work_next, 1);
            for (int i = 0;; ++i) {
                ItemDesc& it = sh.desc[i % kDescRing];
                if (w >= total) {      // end of work: an empty stage carrying the sentinel descriptor
                    it.unit = -1;
                    it.nch = 0;
                    mbar_wait(&sh.empty[stage], phase ^ 1u);
                    mbar_arrive(&sh.full[stage]);
                    break;
                }
                const int wnext = atomicAdd(L.work_next, 1);      // fetched early: its latency hides behind this item
                const int unit = w / nseg, seg = w % nseg;
                int* zs = u.setup_unit(unit, zshared, nullptr);
                const Cmd& c = u.cur;
                it.unit = unit;
                it.start_kind = c.start_kind;
                it.zstate_row = zs;
                it.zst_accept = (c.start_kind == kStartTruth || c.start_kind == kStartSharedKeep) ? kZB
                                                                                                  : (c.zalt == c.zA ? kZA : kZB);
                it.seg = seg;
                it.chunk0 = seg * L.seg_chunks;
                it.nch = min(L.seg_chunks, nchunks - it.chunk0);
                it.sim = c.xi != nullptr;
                it.sig = c.smp.sig;
                it.mus = c.smp.mu;
                it.zk = (c.start_kind == kStartTruth) ? 2 : (c.zcur ? 1 : 0);
                it.zout = L.discard_z ? nullptr : c.zalt;
                it.levmask = 0u;
                const double* ra = it.sim ? c.xi : L.xdat;
                const double* rb = it.sim ? c.nu : nullptr;
                const double* rz = it.zk == 1 ? c.zcur : nullptr;
                if (PERSIST && L.lazy) lazy_item(L, c, zs, zshared, it, rz);
                const uint64_t apol = it.sim ? pol.first : pol.last;
                const int nrows = 1 + (rb != nullptr) + (rz != nullptr);
                for (int k = 0; k < it.nch; ++k) {
                    const int base = (it.chunk0 + k) * kChunk;
                    const uint32_t bytes = (uint32_t)min(kChunk, L.ld - base) * 8u;
                    mbar_wait(&sh.empty[stage], phase ^ 1u);
                    mbar_expect_tx(&sh.full[stage], bytes * (uint32_t)nrows);
                    double* dst = ring + (size_t)stage * rows * kChunk;
                    bulk_g2s(dst, ra + base, bytes, &sh.full[stage], apol);
                    if (rb) bulk_g2s(dst + kChunk, rb + base, bytes, &sh.full[stage], pol.first);
                    if (rz) bulk_g2s(dst + 2 * kChunk, rz + base, bytes, &sh.full[stage], zpol);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
                w = wnext;
            }
            if (dbg) dbg[3] = now_ns();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ finisher
        for (int i = 0;; ++i) {
            const int slot = i & 1;
            mbar_wait(&sh.part_full[slot], (uint32_t)(i >> 1) & 1u);
            const ItemDesc& it = sh.desc[i % kDescRing];
            const int unit = it.unit, seg = it.seg;
            if (unit < 0) break;
            double acc = 0.0;
            const bool mx = (kMaxMask >> lane) & 1u;
            if (lane < kNRed) {
                acc = sh.part[slot][0][lane];
                for (int wv = 1; wv < kNCW; ++wv) {
                    const double o = sh.part[slot][wv][lane];
                    acc = mx ? fmax(acc, o) : acc + o;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.part_empty[slot]);
            if (nseg > 1) {
                // publish the segment; the CTA that publishes a unit's last segment sums them in index order
                if (lane < kNRed) __stcg(L.gpart + ((size_t)unit * nseg + seg) * kRedPad + lane, acc);
                __threadfence();
                __syncwarp();
                int last = 0;
                if (lane == 0) last = atomicAdd(L.gcount + unit, 1) == nseg - 1;
                last = __shfl_sync(0xffffffffu, last, 0);
                if (!last) continue;
                __threadfence();
                __syncwarp();
                if (lane < kNRed) {
                    // segments in index order; their partials are fetched eight at a time (independent loads), not one round trip
                    // to the L2 per segment — the fiducial solve of the one-launch form has one segment per chunk (32 at d = 65 536)
                    const double* gp = L.gpart + (size_t)unit * nseg * kRedPad + lane;
                    acc = __ldcg(gp);
                    for (int sg0 = 1; sg0 < nseg; sg0 += 8) {
                        double o[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) o[u] = sg0 + u < nseg ? __ldcg(gp + (size_t)(sg0 + u) * kRedPad) : 0.0;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (sg0 + u < nseg) acc = mx ? fmax(acc, o[u]) : acc + o[u];
                    }
                }
                if (lane == 0) L.gcount[unit] = 0;      // ready for the next launch
            }
            double t[kNRed];
#pragma unroll
            for (int k = 0; k < kNRed; ++k) t[k] = __shfl_sync(0xffffffffu, acc, k);
            if (lane == 0) {
                publish_unit(L, it, t);
                if (PERSIST && L.lazy) lazy_after_publish(L, it);
            }
            __syncwarp();
        }
        if (dbg && lane == 0) dbg[4] = now_ns();
    } else {
        // ------------------------------------------------------------------ consumers
        const int ct = (int)threadIdx.x - 64, cw = warp - 2;
        const IsoEval ev = launch_ev(L);
        const L2Policy pol = make_policies();
        const LazyView LV = PERSIST ? lazy_view(L) : LazyView{0u, 0};
        RingPos rp{0, 0u};
        for (int i = 0;; ++i) {
            Acc A;
#pragma unroll
            for (int k = 0; k < kNRed; ++k) A.v[k] = 0.0;
            mbar_wait(&sh.full[rp.stage], rp.phase);    // the item's first chunk has landed ⇒ its descriptor is visible
            const ItemDesc& it = sh.desc[i % kDescRing];
            const int slot = i & 1;
            if (it.unit < 0) {                          // end of work: wake the finisher (mailbox protocol as for an item)
                if (lane == 0) {
                    mbar_wait(&sh.part_empty[slot], ((uint32_t)(i >> 1) & 1u) ^ 1u);
                    mbar_arrive(&sh.part_full[slot]);
                }
                break;
            }
            const ItemRegs R{it.sig, it.mus, it.zout, it.levmask};
            consume_item_any<PERSIST, LEAN>(L, sh, ring, rows, stages, it, R, ev, ct, lane, pol.first, LV, rp, A);
            // warp reduction, then the warp's partial goes to the finisher's mailbox
            double sum_k, mv[kNRed - kNSum];
            warp_reduce(A, lane, sum_k, mv);
            if (lane == 0) mbar_wait(&sh.part_empty[slot], ((uint32_t)(i >> 1) & 1u) ^ 1u);
            __syncwarp();
            if (!(lane & 1) && (lane >> 1) < kNSum) sh.part[slot][cw][lane >> 1] = sum_k;
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < kNRed - kNSum; ++k) sh.part[slot][cw][kNSum + k] = mv[k];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.part_full[slot]);
        }
        if (dbg && ct == 0) dbg[5] = now_ns();
    }
    if (PERSIST) {
        fence_proxy_async();
        __syncthreads();
    }
    if (dbg) {
        __syncthreads();
        if (threadIdx.x == 0) dbg[1] = now_ns();
    }
}

__global__ void __launch_bounds__(kThreads, 1)
iso_stream_kernel(const __grid_constant__ SolveLaunch L) {
    if (launch_skipped(L)) return;      // device-resident outer loop: this pass is not needed (uniform over the grid)
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ Shared sh;
    stream_pass<false>(L, sh, reinterpret_cast<double*>(dyn), false);
}

// ---- small latent dimension: one warp per unit, same single pass, straight from global memory ---------------------
// For d below a few thousand a unit is a few KB: no ring, no roles — a warp streams its unit's rows with 128-bit loads
// (4 pairs per lane in flight), keeps the same 15 sums, reduces them with the same warp tree and lane 0 replays the
// optimiser's decisions (fast_replay).  Units are dealt round-robin to the grid's warps; a launch of 10⁴ units is one wave.
constexpr int kWarpCta = 256;

// one unit's sweep by one warp
template <bool SIM, int ZK, int LEAN>
__device__ __forceinline__ void warp_unit(const ItemDesc& it, const double* ra, const double* rb, const double* rz, int d,
                                          const IsoEval& ev, int lane, uint64_t pol, const LazyView& LV, Acc& A) {
    const ItemRegs R{it.sig, it.mus, it.zout, it.levmask};
    const int npairs = d >> 1;
    constexpr int U = 4;
    for (int p0 = lane; p0 < npairs; p0 += U * 32) {
        double2 a[U], b[U], z[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int p = p0 + k * 32;
            const bool ok = p < npairs;
            a[k] = ok ? ld2(ra, p) : make_double2(0.0, 0.0);
            b[k] = (SIM && ok) ? ld2(rb, p) : make_double2(0.0, 0.0);
            z[k] = ((ZK == 1 || ZK == 4) && ok) ? ld2(rz, p) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int p = p0 + k * 32;
            if (p < npairs) {
                double2 zt;
                zt.x = elem_zk<SIM, ZK, LEAN>(a[k].x, b[k].x, z[k].x, ev, R, LV, A);
                zt.y = elem_zk<SIM, ZK, LEAN>(a[k].y, b[k].y, z[k].y, ev, R, LV, A);
                if (it.zout) st2_stream(it.zout + 2 * (size_t)p, zt, pol);
            }
        }
    }
    if ((d & 1) && lane == 0) {                               // odd d: the last element
        const int j = d - 1;
        const double zt = elem_zk<SIM, ZK, LEAN>(ra[j], SIM ? rb[j] : 0.0, (ZK == 1 || ZK == 4) ? rz[j] : 0.0, ev, R, LV, A);
        if (it.zout) it.zout[j] = zt;
    }
}

template <int LEAN>
__device__ __forceinline__ void warp_dispatch(const ItemDesc& it, const double* ra, const double* rb, const double* rz, int d,
                                              const IsoEval& ev, int lane, uint64_t pol, const LazyView& LV, Acc& A) {
    if (it.sim) {
        if (it.zk == 0) warp_unit<true, 0, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
        else if (it.zk == 1) warp_unit<true, 1, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
        else if (it.zk == 2) warp_unit<true, 2, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
        else if (it.zk == 3) warp_unit<true, 3, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
        else warp_unit<true, 4, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
    } else {
        if (it.zk == 1) warp_unit<false, 1, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
        else if (it.zk == 3) warp_unit<false, 3, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
        else if (it.zk == 4) warp_unit<false, 4, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
        else warp_unit<false, 0, LEAN>(it, ra, rb, rz, d, ev, lane, pol, LV, A);
    }
}

// the launch's units dealt round-robin to the grid's warps (body of iso_warp_stream_kernel and of every phase of
// solve_persist_kernel's small-d form)
template <bool PERSIST, int LEAN = false>
__device__ __forceinline__ void warp_pass(const SolveLaunch& L) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * kWarpCta + threadIdx.x) >> 5, nw = (gridDim.x * kWarpCta) >> 5;
    const IsoEval ev = launch_ev(L);
    const L2Policy pol = make_policies();
    WarpCtx ctx{lane};
    NoIssuer none;
    Controller<WarpCtx, NoIssuer> u(ctx, L, none);            // only for setup_unit (pointer logic)
    const double* zshared = u.resolve_zshared();
    const LazyView LV = PERSIST ? lazy_view(L) : LazyView{0u, 0};
    for (int unit = gw; unit < L.nitems; unit += nw) {
        int* zs = u.setup_unit(unit, zshared, nullptr);
        const Cmd& c = u.cur;
        // [host-test:begin warp-item]
        ItemDesc it;
        it.unit = unit;
        it.start_kind = c.start_kind;
        it.zstate_row = zs;
        it.zst_accept = (c.start_kind == kStartTruth || c.start_kind == kStartSharedKeep) ? kZB : (c.zalt == c.zA ? kZA : kZB);
        it.sim = c.xi != nullptr;
        it.sig = c.smp.sig;
        it.mus = c.smp.mu;
        it.zk = (c.start_kind == kStartTruth) ? 2 : (c.zcur ? 1 : 0);
        it.zout = L.discard_z ? nullptr : c.zalt;
        const double* ra = it.sim ? c.xi : L.xdat;
        // [host-test:end warp-item]
        it.levmask = 0u;
        const double* rz = c.zcur;
        if (PERSIST && L.lazy) lazy_item(L, c, zs, zshared, it, rz);
        Acc A;
#pragma unroll
        for (int k = 0; k < kNRed; ++k) A.v[k] = 0.0;
        if constexpr (PERSIST) {
            warp_dispatch<LEAN>(it, ra, c.nu, rz, L.d, ev, lane, pol.first, LV, A);
        } else {
            if (it.sim) {
                if (it.zk == 0) warp_unit<true, 0, false>(it, ra, c.nu, rz, L.d, ev, lane, pol.first, LV, A);
                else if (it.zk == 1) warp_unit<true, 1, false>(it, ra, c.nu, rz, L.d, ev, lane, pol.first, LV, A);
                else warp_unit<true, 2, false>(it, ra, c.nu, rz, L.d, ev, lane, pol.first, LV, A);
            } else {
                if (it.zk == 1) warp_unit<false, 1, false>(it, ra, c.nu, rz, L.d, ev, lane, pol.first, LV, A);
                else warp_unit<false, 0, false>(it, ra, c.nu, rz, L.d, ev, lane, pol.first, LV, A);
            }
        }
        double sum_k, mv[kNRed - kNSum];
        warp_reduce(A, lane, sum_k, mv);
        double t[kNRed];
#pragma unroll
        for (int k = 0; k < kNSum; ++k) t[k] = __shfl_sync(0xffffffffu, sum_k, 2 * k);
#pragma unroll
        for (int k = 0; k < kNRed - kNSum; ++k) t[kNSum + k] = mv[k];
        if (lane == 0) {
            publish_unit(L, it, t);
            if (PERSIST && L.lazy) lazy_after_publish(L, it);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kWarpCta, 2)
iso_warp_stream_kernel(const __grid_constant__ SolveLaunch L) {
    if (launch_skipped(L)) return;
    warp_pass<false>(L);
}

// ---- the whole solve in ONE launch ------------------------------------------------------------------------------------
// solve_persist_kernel runs what muse_b200_muse_solve otherwise enqueues as a chain of ≈ 13 kernels and 2 copies (muse_outer.cu):
// the passes of muse!'s first iterations (/root/reference/src/muse.jl:159-236), the θ update between them, and — once the
// loop has ended — get_H!'s fiducial solve and finite-difference sims (:411-433).  A cooperative launch (every CTA resident);
// a phase = one streaming pass by all CTAs (stream_pass / warp_pass above, unchanged per-unit arithmetic); at its end every CTA
// arrives on a counter, CTA 0 waits for all of them, runs the arithmetic the host loop does between two passes
// (theta_step_body / cov_prep_body of muse_outer_dev.cuh — the code and reduction tree of theta_step_kernel, hence the same θ
// bit for bit), publishes the constants of the next phase and releases a flag the other CTAs spin on.  With several GPUs CTA 0
// also performs the exchange step: it stores this rank's score rows straight into every peer's gathered-score buffer over
// NVLink (peer-mapped memory, muse_comm.cu) and raises a flag there; no collective launch, no staging copy.
// A unit that leaves the fast path (hand-back) makes the launch give up: state `abort`, the host re-runs the solve on the chain
// of launches, whose generic kernel handles such units.  Every spin is bounded (kSpinTimeoutNs): a rank that never shows up
// ends in an error, not in a hung GPU.
constexpr long long kSpinTimeoutNs = 4000000000LL;

__device__ __forceinline__ long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct PersistShared {
    SolveLaunch L;                    // the phase in flight
    DynConsts dyn[2];                 // its θ-dependent constants ([1]: FD sims)
    LazyLevels lazy;                  // constants of the launch's earlier passes (lazy ẑ)
    double red[kMaxTheta][32];        // CTA 0: scratch of the reduction tree
    StepCache cache;                  // CTA 0: θ, the last history row's θ, the last variance (muse_outer_dev.cuh)
    OuterParams step;                 // CTA 0: the θ-step's parameters for the pass just executed (shared, not a copy per thread:
    CovParams cov;                    //        the rank table is indexed per sim)
    int bad;
    int done, error, abort, timeout;  // CTA 0's message after the phase, as every CTA has read it
    int l_abort;                      // CTA 0: this phase handed units back (here or on a peer)
};
static_assert(sizeof(DynConsts) % 8 == 0, "DynConsts is copied in 8-byte words");

// ps.L ← the launch description of phase ph: what muse_pass_enqueue / fd_launch (muse_api.cu) fill in on the host.  Two steps:
// every thread copies one word of the common part out of the kernel's parameter space (one thread doing the 650-byte copy took
// ≈ 1 µs per phase), then — after a barrier — thread 0 patches the phase's own fields.
static_assert(sizeof(SolveLaunch) % 8 == 0 && sizeof(SolveLaunch) / 8 <= kWarpCta, "phase_base: one 8-byte word per thread");
__device__ __forceinline__ void phase_base(const PersistParams& P, PersistShared& ps) {
    constexpr int kWords = (int)(sizeof(SolveLaunch) / 8);
    if ((int)threadIdx.x < kWords)
        reinterpret_cast<unsigned long long*>(&ps.L)[threadIdx.x] = reinterpret_cast<const unsigned long long*>(&P.base)[threadIdx.x];
    __syncthreads();
}
__device__ void phase_launch(const PersistParams& P, int ph, PersistShared& ps) {
    SolveLaunch& L = ps.L;
    PersistCtl* const ctl = P.ctl;
    L.redo_count = &ctl->redo[ph];
    L.work_next = &ctl->work[ph];
    L.redo_total = &ctl->redo_sink;
    L.lean = P.lean;
    const OutPtrs* ob;
    if (ph < kOuterSlots) {              // pass ph + 1 of muse!: data + local sims, start zeros / user z₀ on the first, previous ẑ afterwards
        L.nitems = P.step.units_local;
        L.mode = 0;
        L.include_data = 1;
        L.first_sim = 0;
        L.start_kind = ph == 0 ? P.first_kind : (int)kStartOwn;
        L.zshared = P.z0user;
        L.dyn = &ps.dyn[0];
        ob = &P.slot[ph];
        if (P.lazy) {
            // this pass's constants become level ph of the passes that follow; it materialises ẑ only if it is the launch's last
            // possible pass and the loop may go on after it (on the chain of launches, which starts from stored vectors)
            ps.lazy.lev[ph] = LazyLevel{ps.dyn[0].smp[0].sig, ps.dyn[0].smp[0].mu, ps.dyn[0].ev.a, ps.dyn[0].ev.mu, ps.dyn[0].ev.cspec};
            ps.lazy.nlev = ph;
            ps.lazy.store = (ph == P.max_pass - 1 && P.max_pass < P.step.maxsteps) ? 1 : 0;
            L.lazy = &ps.lazy;
        }
    } else if (ph == kPhaseFid) {        // fiducial MAP of the master stream's draw from zero(z)   — src/muse.jl:417-423
        L.nitems = 1;
        L.mode = 2;
        L.start_kind = kStartZero;
        L.zA = P.zfidA;
        L.zB = P.zfidB;
        L.zstate = &ctl->zfid_state;
        L.dyn = &ps.dyn[0];
        // one unit is all this phase has: one chunk per segment spreads it over as many CTAs as it has chunks (its ẑ is elementwise,
        // so the segmentation does not change it; its sums only feed the accept / hand-back decision)
        if (P.fid_seg_chunks > 0) {
            L.seg_chunks = P.fid_seg_chunks;
            L.nseg = ((L.ld + kChunk - 1) / kChunk + L.seg_chunks - 1) / L.seg_chunks;
        }
        ob = &P.fd;
    } else {                             // virtual sims at the 2·nθ sample points, MAP + score at θ̂ from the fiducial start — :426-433
        L.nitems = P.nh_mine * P.step.nt * 2;
        L.mode = 1;
        L.start_kind = kStartShared;
        L.zshared = nullptr;
        L.zshared_state = &ctl->zfid_state;
        L.zsharedA = P.zfidA;
        L.zsharedB = P.zfidB;
        L.xi = P.xi_fd;
        L.nu = P.nu_fd;
        L.zA = L.zB = nullptr;
        L.zstate = nullptr;
        L.discard_z = 1;
        L.dyn = &ps.dyn[1];
        ob = &P.fd;
    }
    L.g_out = ob->g; L.iters_out = ob->iters; L.fg_out = ob->fg;
    L.gnorm_out = ob->gnorm; L.f_out = ob->f; L.status_out = ob->status;
    L.zrows = (L.start_kind == kStartOwn || L.start_kind == kStartShared || L.start_kind == kStartSharedKeep) ? 1 : 0;
    if (L.lazy) L.zrows = P.first_kind == kStartSharedKeep ? 1 : 0;      // only a user start vector is ever streamed
    L.stream_stages = L.zrows ? 4 : 6;
}

template <int STREAM, int LEAN>
__device__ __forceinline__ void run_phase(const SolveLaunch& L, bool reinit) {
    if constexpr (STREAM == 1) {
        extern __shared__ __align__(128) unsigned char dynsm[];
        __shared__ Shared sh;
        stream_pass<true, LEAN>(L, sh, reinterpret_cast<double*>(dynsm), reinit);
    } else {
        warp_pass<true, LEAN>(L);
        __syncthreads();
    }
}

// device → mapped pinned host memory, 16 bytes per thread and step (both 16-byte aligned; a trailing 8 bytes handled apart).
// part / nparts: this caller's slice (a CTA of the grid, or everything for CTA 0 alone).  The data were published by other CTAs
// before the phase barrier: read at the L2.
__device__ void copy_to_host(unsigned char* dst, const unsigned char* src, unsigned long long bytes, int part, int nparts) {
    const unsigned long long words = bytes / 16, per = (words + nparts - 1) / nparts;
    const unsigned long long w0 = (unsigned long long)part * per, w1 = w0 + per < words ? w0 + per : words;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (unsigned long long w = w0 + threadIdx.x; w < w1; w += blockDim.x) d4[w] = __ldcg(s4 + w);
    if (part == nparts - 1 && threadIdx.x == 0 && (bytes & 8ULL))
        *reinterpret_cast<unsigned long long*>(dst + words * 16) = __ldcg(reinterpret_cast<const unsigned long long*>(src + words * 16));
}

// thread 0 of CTA 0: until every CTA has arrived at the end of phase ph
__device__ void wait_arrivals(const PersistParams& P, int ph, PersistShared& ps) {
    const long long t0 = gtime_ns();
    while (ld_acquire_gpu(&P.ctl->arrive[ph]) < (int)gridDim.x)
        if (gtime_ns() - t0 > kSpinTimeoutNs) { ps.timeout = 1; break; }
    if (__ldcg(&P.ctl->redo[ph]) > 0) ps.l_abort = 1;
}

// The exchange step by CTA 0 (all its threads): `n` doubles of this rank (rows src, one status per `per_row` doubles; failed units
// go out as NaN so that every rank takes the same error decision) → offset my_off of dst[q] on EVERY rank q, then flag `ep`
// there; then wait for the flags of all peers.  An abort bit rides on the flag.
__device__ void exchange_rows(const PersistParams& P, PersistShared& ps, const double* src, const int* status, int n, int per_row,
                              double* const* dst, long long my_off, unsigned long long ep) {
    const XchgParams& X = P.x;
    const int tid = threadIdx.x;
    for (int e = tid; e < n; e += blockDim.x) {
        const bool bad = __ldcg(status + e / per_row) == MUSE_STATUS_NONFINITE;
        const double v = bad ? __longlong_as_double(0x7ff8000000000000LL) : __ldcg(src + e);
        for (int q = 0; q < X.nranks; ++q) dst[q][(size_t)my_off + e] = v;
    }
    // the rows of all threads happen-before the barrier, the barrier before the flag threads' release stores: one system-scope
    // release per peer orders everything (a __threadfence_system() in each of the CTA's 18 warps cost more than the exchange)
    __syncthreads();
    if (tid < X.nranks && tid != X.rank) {
        st_release_sys(X.flags[tid] + 2 * X.rank, ep * 2 + (ps.l_abort ? 1ULL : 0ULL));
        const long long t0 = gtime_ns();
        unsigned long long v;
        while (((v = ld_acquire_sys(X.flags[X.rank] + 2 * tid)) >> 1) < ep)
            if (gtime_ns() - t0 > kSpinTimeoutNs) { ps.timeout = 1; break; }
        if ((v >> 1) == ep && (v & 1ULL)) ps.l_abort = 1;
    }
    __syncthreads();
}

// CTA 0 after pass i: wait for the grid, exchange, θ-step (and, if the loop has ended, the covariance stage's constants), release
template <int V>
__device__ void leader_step(const PersistParams& P, PersistShared& ps, int i, unsigned seq) {
    PersistCtl* const ctl = P.ctl;
    volatile OuterState* const st = P.step.st;
    const int tid = threadIdx.x, nt = P.step.nt, ph = i - 1;
    const OutPtrs& ob = P.slot[ph];
    if (tid == 0) {
        ps.l_abort = 0;
        ps.step = P.step;
        ps.step.iter = i;
        ps.step.g_local = ob.g;
        ps.step.status_local = ob.status;
        ps.step.g_all = P.x.nranks > 1 ? P.x.gall[P.x.rank][ph] : ob.g + nt;
        ps.step.dense = 1;
        ps.step.dyn_next = &ctl->dyn[0];
        ps.cov = P.cov;
        ps.cov.dyn_fid = &ctl->dyn[0];
        ps.cov.dyn_fd = &ctl->dyn[1];
        wait_arrivals(P, ph, ps);
        P.stamps[1 + 2 * ph] = gtime_ns();
    }
    __syncthreads();
    if (P.x.nranks > 1 && !ps.timeout) {
        double* dst[kMaxRanks];
        for (int q = 0; q < P.x.nranks; ++q) dst[q] = P.x.gall[q][ph];
        // score rows land in GLOBAL sim order (rank r's rows behind those of ranks < r): the θ-step reads them like one GPU's
        exchange_rows(P, ps, ob.g + nt, ob.status + 1, P.step.counts[P.x.rank] * nt, nt, dst, (long long)P.x.row0 * nt, P.x.epoch0 + i);
        if (tid == 0 && ph < 2) P.stamps[11 + ph] = gtime_ns();          // diagnostics: rows of every peer are here
        if (P.gall_h[ph] && !ps.timeout)
            copy_to_host(reinterpret_cast<unsigned char*>(P.gall_h[ph]), reinterpret_cast<const unsigned char*>(P.x.gall[P.x.rank][ph]),
                         (unsigned long long)P.step.n_total * nt * sizeof(double), 0, 1);
    }
    if (!ps.l_abort && !ps.timeout) {
        theta_step_body<V>(ps.step, ps.red, &ps.bad, &ps.cache);
        __syncthreads();
        if (P.get_cov && ps.cache.done && !ps.cache.error) cov_prep_body<V>(ps.cov, ps.red, &ps.cache);
    }
    __syncthreads();
    if (tid == 0) {
        if (ps.timeout) { st->error = 3; st->done = 1; ps.cache.error = 3; ps.cache.done = 1; }
        if (ps.l_abort) st->abort = 1;
        ctl->msg_done = ps.cache.done;
        ctl->msg_error = ps.cache.error;
        ctl->msg_abort = ps.l_abort;
        P.stamps[2 + 2 * ph] = gtime_ns();
        __threadfence();
        st_release_gpu(&ctl->step_flag, seq);
    }
}

// every CTA: until CTA 0 has released `seq`; then its message and the constants it published are in ps (one round trip to the L2
// for both: every thread fetches one word)
__device__ bool wait_step(const PersistParams& P, PersistShared& ps, unsigned seq) {
    PersistCtl* const ctl = P.ctl;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const long long t0 = gtime_ns();
        while (ld_acquire_gpu(&ctl->step_flag) < seq)
            if (gtime_ns() - t0 > kSpinTimeoutNs + 1000000000LL) { ps.timeout = 1; break; }
    }
    __syncthreads();
    if (ps.timeout) return false;
    constexpr int kWords = 2 * (int)sizeof(DynConsts) / 8;
    if (tid < kWords) reinterpret_cast<double*>(ps.dyn)[tid] = __ldcg(reinterpret_cast<const double*>(ctl->dyn) + tid);
    else if (tid == kWords) ps.done = __ldcg(&ctl->msg_done);
    else if (tid == kWords + 1) ps.error = __ldcg(&ctl->msg_error);
    else if (tid == kWords + 2) ps.abort = __ldcg(&ctl->msg_abort);
    __syncthreads();
    return true;
}
static_assert(2 * sizeof(DynConsts) / 8 + 3 <= kWarpCta, "wait_step: one word per thread");

template <int STREAM, int LEAN>
__global__ void __launch_bounds__(STREAM == 1 ? kThreads : kWarpCta, STREAM == 1 ? 1 : 2)
solve_persist_kernel(const __grid_constant__ PersistParams P) {
    constexpr int V = STREAM == 1 ? 2 : 4;            // lanes of the θ-step's reduction tree per thread (512 / 256 threads carry them)
    __shared__ PersistShared ps;
    PersistCtl* const ctl = P.ctl;
    const int tid = threadIdx.x;
    const bool lead = blockIdx.x == 0;
    if (tid == 0) {
        ps.done = ps.error = ps.abort = ps.timeout = ps.l_abort = 0;
        ps.dyn[0] = P.first;
        if (lead) {
            OuterState* st = P.step.st;
            st->n_iter = 0; st->done = 0; st->error = 0; st->abort = 0;
            for (int c = 0; c < kMaxTheta; ++c) {
                st->theta[c] = P.theta0[c]; st->step[c] = 0.0;
                ps.cache.theta[c] = P.theta0[c]; ps.cache.row_theta[c] = 0.0; ps.cache.var[c] = 0.0;
            }
            ps.cache.n_iter = ps.cache.done = ps.cache.error = 0;
            for (int k = 0; k < 16; ++k) P.stamps[k] = 0;
            P.stamps[0] = gtime_ns();
        }
    }
    unsigned seq = 0;
    bool reinit = false, finished = false;
    for (int i = 1; i <= P.max_pass; ++i) {
        phase_base(P, ps);
        if (tid == 0) phase_launch(P, i - 1, ps);
        __syncthreads();
        run_phase<STREAM, LEAN>(ps.L, reinit);
        reinit = true;
        if (tid == 0) { __threadfence(); atomicAdd(&ctl->arrive[i - 1], 1); }
        ++seq;
        if (lead) leader_step<V>(P, ps, i, seq);
        if (!wait_step(P, ps, seq)) break;
        // this pass's per-unit results → the host mirror, a slice per CTA (posted writes: they overlap what follows)
        if (P.slot_h[i - 1]) copy_to_host(P.slot_h[i - 1], P.slot_d[i - 1], P.slot_bytes, blockIdx.x, gridDim.x);
        if (ps.abort || ps.error) break;
        if (ps.done) { finished = true; break; }
    }
    if (finished && P.get_cov) {
        bool ran_fd = false;
        if (P.nh_mine > 0) {
            phase_base(P, ps);
            if (tid == 0) phase_launch(P, kPhaseFid, ps);
            __syncthreads();
            run_phase<STREAM, LEAN>(ps.L, reinit);
            reinit = true;
            if (tid == 0) { __threadfence(); atomicAdd(&ctl->arrive[kPhaseFid], 1); }
            ++seq;
            if (lead) {
                if (tid == 0) {
                    ps.l_abort = 0;
                    wait_arrivals(P, kPhaseFid, ps);
                    P.stamps[1 + 2 * kPhaseFid] = gtime_ns();
                    ctl->msg_abort = ps.l_abort;
                    if (ps.timeout) ctl->msg_error = 3;
                    __threadfence();
                    st_release_gpu(&ctl->step_flag, seq);
                }
            }
            if (wait_step(P, ps, seq) && !ps.abort && !ps.error) {
                phase_base(P, ps);
                if (tid == 0) phase_launch(P, kPhaseFd, ps);
                __syncthreads();
                run_phase<STREAM, LEAN>(ps.L, reinit);
                if (tid == 0) { __threadfence(); atomicAdd(&ctl->arrive[kPhaseFd], 1); }
                ran_fd = true;
            }
        }
        if (lead) {                       // the end of the covariance stage: hand-backs, and the exchange of the FD scores
            __syncthreads();
            if (tid == 0) {
                if (ran_fd) wait_arrivals(P, kPhaseFd, ps);
                else if (ps.abort) ps.l_abort = 1;
                P.stamps[1 + 2 * kPhaseFd] = gtime_ns();
            }
            __syncthreads();
            if (P.x.nranks > 1 && !ps.timeout) {
                double* dst[kMaxRanks];
                for (int q = 0; q < P.x.nranks; ++q) dst[q] = P.x.fdall[q];
                const int nt = P.step.nt;
                exchange_rows(P, ps, P.fd.g, P.fd.status, P.nh_mine * nt * 2 * nt, nt, dst, (long long)P.x.rank * P.x.need_fd, P.x.epoch0 + kOuterSlots + 1);
            }
            if (tid == 0) {
                volatile OuterState* st = P.step.st;
                if (ps.l_abort) st->abort = 1;
                if (ps.timeout) st->error = 3;
                P.stamps[2 + 2 * kPhaseFd] = gtime_ns();
            }
        }
    }
    // CTA 0: what is left for the host — the FD block (its units were published before the arrivals CTA 0 waited for), the
    // gathered FD rows, and the state (header + the history rows that were written)
    __syncthreads();
    if (lead) {
        if (tid == 0 && ps.timeout) { volatile OuterState* st = P.step.st; st->error = 3; }
        __syncthreads();
        if (P.fd_h && P.fd_bytes) copy_to_host(P.fd_h, P.fd_d, P.fd_bytes, 0, 1);
        if (P.fdall_h && P.x.nranks > 1)
            copy_to_host(reinterpret_cast<unsigned char*>(P.fdall_h), reinterpret_cast<const unsigned char*>(P.x.fdall[P.x.rank]),
                         (unsigned long long)P.x.nranks * P.x.need_fd * sizeof(double), 0, 1);
        if (P.st_h) {
            __threadfence();           // thread 0's own state writes, before everybody reads them back
            __syncthreads();
            const unsigned long long nb = (offsetof(OuterState, row) + (unsigned long long)kOuterSlots * sizeof(OuterRow) + 15ULL) & ~15ULL;
            copy_to_host(reinterpret_cast<unsigned char*>(P.st_h), reinterpret_cast<const unsigned char*>(P.step.st), nb, 0, 1);
        }
    }
    // the last CTA to leave clears the control block for the next launch — and tells the host: every CTA's stores to the host
    // mirrors are ordered (system scope) before its exit count, so the word below is the last thing to land and the host need not wait
    // for the launch to retire
    __syncthreads();
    if (tid == 0) {
        if (P.done_h) __threadfence_system(); else __threadfence();
        if (atomicAdd(&ctl->exit_count, 1) == (int)gridDim.x - 1) {
            __threadfence();
            volatile int* w = reinterpret_cast<volatile int*>(ctl);
            for (int k = 0; k < (int)(offsetof(PersistCtl, dyn) / sizeof(int)); ++k) w[k] = 0;
            __threadfence();
            if (P.done_h) {
                __threadfence_system();
                *reinterpret_cast<volatile unsigned long long*>(P.done_h) = P.done_seq;
            }
        }
    }
}
}  // namespace

// Geometry: chunks, segments, ring depth.  One CTA per SM.
cudaError_t iso_stream_geometry(int d, int ld, int device, Geometry* geo) {
    (void)d;
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    const int nchunks = (ld + kChunk - 1) / kChunk;
    geo->stream = 1;
    geo->stream_grid = sms;
    // A unit is always cut into the same segments (≤ 8 chunks = 128 KB per row each), whatever else is in the
    // launch: the segment is a node of the fixed reduction tree, so a unit's sums — and with them ẑ and g — do not
    // depend on how many units a launch holds or on how sims are sharded over GPUs.
    geo->nseg = (nchunks + 7) / 8;
    geo->seg_chunks = (nchunks + geo->nseg - 1) / geo->nseg;
    geo->nseg = (nchunks + geo->seg_chunks - 1) / geo->seg_chunks;
    geo->smem_bytes = 4 * 3 * kChunk * 8;                     // ring: 4 stages × 3 rows or 6 stages × 2 rows (192 KB)
    e = cudaFuncSetAttribute(iso_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, geo->smem_bytes);
    return e;
}

// pass 1: the streaming kernel (geo.stream == 1: TMA ring, one CTA per SM; == 2: one warp per unit, small d)
cudaError_t launch_iso_stream(SolveLaunch& L, const Geometry& geo, cudaStream_t st) {
    if (geo.stream == 2) {
        L.nseg = 1;
        const int warps_per_cta = kWarpCta / 32;
        int grid = (L.nitems + warps_per_cta - 1) / warps_per_cta;   // one unit per warp (the loop only runs again beyond 2²⁰ CTAs)
        if (grid > (1 << 20)) grid = 1 << 20;
        if (grid < 1) grid = 1;
        iso_warp_stream_kernel<<<grid, kWarpCta, 0, st>>>(L);
        return cudaGetLastError();
    }
    L.seg_chunks = geo.seg_chunks;
    L.nseg = geo.nseg;
    L.zrows = (L.start_kind == kStartOwn || L.start_kind == kStartShared || L.start_kind == kStartSharedKeep) ? 1 : 0;
    L.stream_stages = L.zrows ? 4 : 6;
    const long long total = (long long)L.nitems * L.nseg;
    int grid = geo.stream_grid;
    if (grid > total) grid = (int)total;
    if (grid < 1) grid = 1;
    cudaError_t e = cudaFuncSetAttribute(iso_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.smem_bytes);
    if (e != cudaSuccess) return e;
    iso_stream_kernel<<<grid, kThreads, geo.smem_bytes, st>>>(L);
    return cudaGetLastError();
}

// solve_persist_kernel: a cooperative grid — one CTA per SM with the ring (d ≥ 4096), as many 256-thread CTAs as are resident
// otherwise
cudaError_t iso_persist_geometry(const Geometry& geo, int device, int* grid, int* threads) {
    int sms = 0, coop = 0, per_sm = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (e != cudaSuccess) return e;
    if (!coop || (geo.stream != 1 && geo.stream != 2)) { *grid = 0; *threads = 0; return cudaSuccess; }
    if (geo.stream == 1) {
        e = cudaFuncSetAttribute(solve_persist_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.smem_bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_persist_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.smem_bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_persist_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.smem_bytes);
        if (e != cudaSuccess) return e;
        int a = 0, b = 0, c = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, solve_persist_kernel<1, 0>, kThreads, geo.smem_bytes);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, solve_persist_kernel<1, 1>, kThreads, geo.smem_bytes);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, solve_persist_kernel<1, 2>, kThreads, geo.smem_bytes);
        per_sm = a < b ? (a < c ? a : c) : (b < c ? b : c);
        *threads = kThreads;
    } else {
        int a = 0, b = 0, c = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, solve_persist_kernel<2, 0>, kWarpCta, 0);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, solve_persist_kernel<2, 1>, kWarpCta, 0);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, solve_persist_kernel<2, 2>, kWarpCta, 0);
        per_sm = a < b ? (a < c ? a : c) : (b < c ? b : c);
        *threads = kWarpCta;
    }
    if (e != cudaSuccess) return e;
    *grid = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_iso_persist(const PersistParams& P, const Geometry& geo, int grid, cudaStream_t st) {
    void* args[] = {const_cast<PersistParams*>(&P)};
    const void* const fns[2][3] = {
        {reinterpret_cast<const void*>(&solve_persist_kernel<1, 0>), reinterpret_cast<const void*>(&solve_persist_kernel<1, 1>), reinterpret_cast<const void*>(&solve_persist_kernel<1, 2>)},
        {reinterpret_cast<const void*>(&solve_persist_kernel<2, 0>), reinterpret_cast<const void*>(&solve_persist_kernel<2, 1>), reinterpret_cast<const void*>(&solve_persist_kernel<2, 2>)}};
    const void* fn = fns[geo.stream == 1 ? 0 : 1][P.lean < 0 || P.lean > 2 ? 0 : P.lean];
    if (geo.stream == 1) return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, (size_t)geo.smem_bytes, st);
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kWarpCta), args, 0, st);
}

// warp-per-unit geometry (small d)
cudaError_t iso_warp_stream_geometry(int device, Geometry* geo) {
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    geo->stream = 2;
    geo->stream_grid = sms;
    geo->seg_chunks = 0;
    geo->nseg = 1;
    geo->smem_bytes = 0;
    return cudaSuccess;
}

}  // namespace muse
