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
// ẑ = z₀ + c_spec·s as it goes.  The scalar optimiser (the same Controller code as the generic
// kernel, muse_iso_ctl.cuh) then replays L-BFGS/Hager–Zhang on the reduced sums; when it asks for
// the committed trial at a step c with |c − c_spec| ≤ 1e-11·c_spec the speculative sums answer it.
// Any other request (another trial, a second iteration, a start vector that had to be kept …)
// aborts the unit: nothing of it is published and its index goes to a device-side list that the
// generic two-sweep kernel re-solves from scratch.  Parity is unaffected: ẑ differs from the
// non-speculative result by ≤ 1e-11 relative (tests: rtol 1e-8), iteration and evaluation
// counts are those of the replayed algorithm.
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
//                      segment sums the segments in index order and runs the scalar optimiser
//   warps 2..17        consumers: wait full[stage], fused elementwise work out of shared memory,
//                      128-bit coalesced stores of ẑ, arrive empty[stage]; at the end of an item a
//                      shuffle tree reduces the 15 running sums, which go to the finisher through a
//                      two-slot mailbox — consumers never wait for the scalar code.
// Every reduction is a fixed tree (thread → warp butterfly → warps in order → segments in order),
// so results do not depend on scheduling, grid size or how sims are sharded over GPUs.
#include "muse_iso_ctl.cuh"

namespace muse {

namespace {

constexpr int kNC = 512;               // consumer threads
constexpr int kNCW = kNC / 32;         // consumer warps
constexpr int kThreads = kNC + 64;     // + producer warp + finisher warp
constexpr int kChunk = 2048;           // elements per row per stage (16 KB)
constexpr int kMaxStages = 8;
constexpr int kDescRing = 16;
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
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// what the producer tells the other roles about one work item
struct ItemDesc {
    double* zout;        // where ẑ goes (null: discarded)
    double sig, mus;     // sample_x_z constants of the unit (z = mus + sig·ξ)
    int unit;            // launch item
    int seg;
    int chunk0, nch;
    int sim;             // rows 0,1 = ξ,ν (else row 0 = x)
    int zk;              // 0: z₀ ≡ 0, 1: z₀ streamed in row 2, 2: z₀ = simulated latent (truth)
};

struct Shared {
    uint64_t full[kMaxStages], empty[kMaxStages];
    uint64_t part_full[2], part_empty[2];
    ItemDesc desc[kDescRing];
    double part[2][kNCW][kRedPad];
};

struct Acc {
    double v[kNRed];
};

// one element: the three evaluations of the fast path (see the file comment)
template <bool SIM, int ZK>
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
    // φ(1), φ'(1) along s = −∇f(z₀)
    const double z1 = z0 - g0;
    const double r1 = x - z1, w1 = z1 - ev.mu;
    const double g1 = fma(ev.a, w1, -r1);
    A.v[rR1] = fma(r1, r1, A.v[rR1]);
    A.v[rS2_1] = fma(w1, w1, A.v[rS2_1]);
    A.v[rDP1] = fma(g1, -g0, A.v[rDP1]);
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
    A.v[rXC] = fmax(A.v[rXC], fabs(zt - z0));
    return zt;
}

__device__ __forceinline__ void st2_stream(double* p, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

// one chunk of one item, executed by the consumer threads
template <bool SIM, int ZK>
__device__ __forceinline__ void consume_chunk(const double* buf, int base, int len, int d, const ItemDesc& it,
                                              const IsoEval& ev, int ct, uint64_t pol, Acc& A) {
    const double* ra = buf;
    const double* rb = buf + kChunk;
    const double* rz = buf + 2 * kChunk;
    double* zout = it.zout;
    if (base + len <= d) {
#pragma unroll
        for (int u = 0; u < kChunk / (2 * kNC); ++u) {
            const int q2 = 2 * (ct + u * kNC);
            if (q2 < len) {
                const double2 p = lds2(ra + q2);
                const double2 q = SIM ? lds2(rb + q2) : make_double2(0.0, 0.0);
                const double2 z = (ZK == 1) ? lds2(rz + q2) : make_double2(0.0, 0.0);
                double2 zt;
                zt.x = elem3<SIM, ZK>(p.x, q.x, z.x, ev, it.sig, it.mus, A);
                zt.y = elem3<SIM, ZK>(p.y, q.y, z.y, ev, it.sig, it.mus, A);
                if (zout) st2_stream(zout + base + q2, zt, pol);
            }
        }
    } else {   // the row's last chunk: elements ≥ d are padding
        for (int q2 = 2 * ct; q2 < len; q2 += 2 * kNC) {
            const int j = base + q2;
            if (j >= d) break;
            const double2 p = lds2(ra + q2);
            const double2 q = SIM ? lds2(rb + q2) : make_double2(0.0, 0.0);
            const double2 z = (ZK == 1) ? lds2(rz + q2) : make_double2(0.0, 0.0);
            double2 zt;
            zt.x = elem3<SIM, ZK>(p.x, q.x, z.x, ev, it.sig, it.mus, A);
            zt.y = 0.0;
            if (j + 1 < d) zt.y = elem3<SIM, ZK>(p.y, q.y, z.y, ev, it.sig, it.mus, A);
            if (zout) {
                if (j + 1 < d) st2_stream(zout + j, zt, pol);
                else zout[j] = zt.x;
            }
        }
    }
}

// issuer of the streaming kernel: answers the fast path's two sweeps from the reduced sums
struct SpecIssuer {
    const SolveLaunch& L;
    double t[kNRed];
    bool* abort_flag;
    __device__ SpecIssuer(const SolveLaunch& l) : L(l), abort_flag(nullptr) {}
    __device__ __noinline__ void operator()(Cmd& cur, double (&red)[7]) {
        const IsoEval& ev = L.ev;
        if (cur.op == kOpInit) {
            red[0] = fma(ev.a, t[rS2_0], t[rR0]);       // e₀ = Σ(x−z)² + a Σ(z−μ)²
            red[1] = t[rGG0];
            red[2] = t[rGM0];
            red[3] = t[rS1_0];
            red[4] = t[rS2_0];
            red[5] = fma(ev.a, t[rS2_1], t[rR1]);       // e(z₀+s)
            red[6] = t[rDP1];
            return;
        }
        if (cur.op == kOpTrial && cur.lazy && cur.commit && fabs(cur.c - ev.cspec) <= kSpecTol * ev.cspec) {
            red[0] = fma(ev.a, t[rS2T], t[rRT]);
            red[1] = t[rDPT];
            red[2] = t[rGGT];
            red[3] = t[rGMT];
            red[4] = t[rS1T];
            red[5] = t[rS2T];
            red[6] = t[rXC];
            return;
        }
        *abort_flag = true;
#pragma unroll
        for (int k = 0; k < 7; ++k) red[k] = NAN;
    }
};

struct WarpCtx {
    int tid;
};

__global__ void __launch_bounds__(kThreads, 1)
iso_stream_kernel(const __grid_constant__ SolveLaunch L) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ Shared sh;
    double* const ring = reinterpret_cast<double*>(dyn);
    const int rows = L.zrows ? 3 : 2;
    const int stages = L.stream_stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], kNCW);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sh.part_full[s], kNCW);
            mbar_init(&sh.part_empty[s], 1);
        }
        fence_mbar_init();
    }
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
            SpecIssuer dummy(L);
            Controller<WarpCtx, SpecIssuer> u(ctx, L, dummy);     // only for setup_unit (pointer logic)
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
                u.setup_unit(unit, zshared, nullptr);
                const Cmd& c = u.cur;
                it.unit = unit;
                it.seg = seg;
                it.chunk0 = seg * L.seg_chunks;
                it.nch = min(L.seg_chunks, nchunks - it.chunk0);
                it.sim = c.xi != nullptr;
                it.sig = c.smp.sig;
                it.mus = c.smp.mu;
                it.zk = (c.start_kind == kStartTruth) ? 2 : (c.zcur ? 1 : 0);
                it.zout = L.discard_z ? nullptr : c.zalt;
                const double* ra = it.sim ? c.xi : L.xdat;
                const double* rb = it.sim ? c.nu : nullptr;
                const double* rz = it.zk == 1 ? c.zcur : nullptr;
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
            const int unit = sh.desc[i % kDescRing].unit, seg = sh.desc[i % kDescRing].seg;
            if (unit < 0) break;
            double acc = 0.0;
            if (lane < kNRed) {
                acc = sh.part[slot][0][lane];
                const bool mx = (kMaxMask >> lane) & 1u;
                for (int wv = 1; wv < kNCW; ++wv) {
                    const double o = sh.part[slot][wv][lane];
                    acc = mx ? fmax(acc, o) : acc + o;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.part_empty[slot]);
            if (lane < kNRed) L.gpart[((size_t)unit * nseg + seg) * kRedPad + lane] = acc;   // read by iso_replay_kernel
        }
        if (dbg && lane == 0) dbg[4] = now_ns();
    } else {
        // ------------------------------------------------------------------ consumers
        const int ct = (int)threadIdx.x - 64, cw = warp - 2;
        const IsoEval ev = L.ev;
        const L2Policy pol = make_policies();
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0;; ++i) {
            Acc A;
#pragma unroll
            for (int k = 0; k < kNRed; ++k) A.v[k] = 0.0;
            mbar_wait(&sh.full[stage], phase);          // the item's first chunk has landed ⇒ its descriptor is visible
            const ItemDesc it = sh.desc[i % kDescRing];
            const int slot = i & 1;
            if (it.unit < 0) {                          // end of work: wake the finisher (mailbox protocol as for an item)
                if (lane == 0) {
                    mbar_wait(&sh.part_empty[slot], ((uint32_t)(i >> 1) & 1u) ^ 1u);
                    mbar_arrive(&sh.part_full[slot]);
                }
                break;
            }
            for (int k = 0; k < it.nch; ++k) {
                if (k) mbar_wait(&sh.full[stage], phase);
                const double* buf = ring + (size_t)stage * rows * kChunk;
                const int base = (it.chunk0 + k) * kChunk;
                const int len = min(kChunk, L.ld - base);
                if (it.sim) {
                    if (it.zk == 0) consume_chunk<true, 0>(buf, base, len, L.d, it, ev, ct, pol.first, A);
                    else if (it.zk == 1) consume_chunk<true, 1>(buf, base, len, L.d, it, ev, ct, pol.first, A);
                    else consume_chunk<true, 2>(buf, base, len, L.d, it, ev, ct, pol.first, A);
                } else {
                    if (it.zk == 1) consume_chunk<false, 1>(buf, base, len, L.d, it, ev, ct, pol.first, A);
                    else consume_chunk<false, 0>(buf, base, len, L.d, it, ev, ct, pol.first, A);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            // Warp reduction, then the warp's partial goes to the finisher's mailbox.
            // Sums: the xor butterfly (16, 8, 4, 2, 1) done "transposed" — at every level a lane keeps half of its
            // slots and ships the other half — 16 shuffles instead of 60, same association tree, slot k ends in
            // lane 2k.  Maxima (non-negative): two integer REDUX over the high and low words.
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
            sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
            double mv[kNRed - kNSum];
#pragma unroll
            for (int k = 0; k < kNRed - kNSum; ++k) {
                const unsigned hi = (unsigned)__double2hiint(A.v[kNSum + k]), lo = (unsigned)__double2loint(A.v[kNSum + k]);
                const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
                const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
                mv[k] = __hiloint2double((int)mh, (int)ml);
            }
            if (lane == 0) mbar_wait(&sh.part_empty[slot], ((uint32_t)(i >> 1) & 1u) ^ 1u);
            __syncwarp();
            if (!(lane & 1) && (lane >> 1) < kNSum) sh.part[slot][cw][lane >> 1] = sv[0];
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < kNRed - kNSum; ++k) sh.part[slot][cw][kNSum + k] = mv[k];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.part_full[slot]);
        }
        if (dbg && ct == 0) dbg[5] = now_ns();
    }
    if (dbg) {
        __syncthreads();
        if (threadIdx.x == 0) dbg[1] = now_ns();
    }
}

// Scalar replay: one warp per unit sums the unit's segment partials in index order and runs the
// L-BFGS / Hager–Zhang controller on them (see the file comment).  Accepted units get their outputs
// and their ẑ buffer flip; the others go to the hand-back list of the generic kernel.
constexpr int kReplayWarps = 4;

__global__ void __launch_bounds__(kReplayWarps * 32)
iso_replay_kernel(const __grid_constant__ SolveLaunch L) {
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kReplayWarps + (threadIdx.x >> 5);
    if (unit >= L.nitems) return;
    const int nseg = L.nseg;
    double acc = 0.0;
    if (lane < kNRed) {
        const bool mx = (kMaxMask >> lane) & 1u;
        acc = L.gpart[(size_t)unit * nseg * kRedPad + lane];
        for (int s = 1; s < nseg; ++s) {
            const double o = L.gpart[((size_t)unit * nseg + s) * kRedPad + lane];
            acc = mx ? fmax(acc, o) : acc + o;
        }
    }
    WarpCtx ctx{lane};
    SpecIssuer issuer(L);
    Controller<WarpCtx, SpecIssuer> ctl(ctx, L, issuer);
    issuer.abort_flag = &ctl.abort;
    ctl.spec_mode = true;
    ctl.dxh = ctl.dgh = nullptr;
#pragma unroll
    for (int k = 0; k < kNRed; ++k) issuer.t[k] = __shfl_sync(0xffffffffu, acc, k);
    ctl.cur.sbuf = nullptr;
    ctl.cur.v1 = ctl.cur.v2 = nullptr;
    ctl.cur.w1 = ctl.cur.w2 = nullptr;
    ctl.cur.c = 0.0;
    int* zs = ctl.setup_unit(unit, ctl.resolve_zshared(), nullptr);
    ctl.solve(unit, zs);
    if (ctl.abort && lane == 0) {
        L.redo_items[atomicAdd(L.redo_count, 1)] = unit;
        atomicAdd(L.redo_total, 1ULL);
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

// pass 1 (streaming) + its scalar replay
cudaError_t launch_iso_stream(SolveLaunch& L, const Geometry& geo, cudaStream_t st) {
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
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    SolveLaunch R = L;
    R.dbg = nullptr;      // the per-CTA rows of the streaming kernel own the diagnostics buffer
    iso_replay_kernel<<<(L.nitems + kReplayWarps - 1) / kReplayWarps, kReplayWarps * 32, 0, st>>>(R);
    return cudaGetLastError();
}

}  // namespace muse
