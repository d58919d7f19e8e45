// muse_iso_tma.cu — the large-d variant of the persistent MAP + score solver: sweeps are fed by a
// bulk-async (TMA, `cp.async.bulk`) producer/consumer pipeline through shared memory, and the
// unit's simulated data x can stay resident in the shared memory of a thread-block cluster.
//
// Same algorithm, same controller (muse_iso_ctl.cuh) and same results contract as
// muse_iso_solver.cu — see that file for what is being replaced in the reference
// (/root/reference/src/muse.jl:170-175, src/interface.jl:162-166, src/simple.jl:61-68, 92).
//
// Why a second kernel.  With plain loads the number of bytes a thread keeps in flight is bounded
// by its registers (profiles/r01: 4.6-5.0 TB/s, and the cold TRIAL sweep, which reads one vector,
// is latency-bound).  Here warp 0 of every CTA is the controller *and* the TMA producer: for each
// sweep it streams the CTA's chunks of the input rows into a ring of `stages` shared-memory
// buffers with `cp.async.bulk … mbarrier::complete_tx`, while the NC consumer threads wait on the
// `full` mbarriers, do the fused elementwise work out of shared memory, write results with
// coalesced 128-bit stores and hand the stage back through the `empty` mbarriers.  Bytes in
// flight per SM = stages × chunk bytes, independent of registers.
//
// Resident x.  A unit's x = sample(θ_sim; ξ, ν) is produced by the INIT sweep and consumed again by
// the TRIAL sweep(s).  When the cluster's CTAs can hold their slices of x in shared memory
// (`resident`), x never goes to global memory on the fast path: HBM traffic per unit is then the
// fused floor — read ξ, ν [, z₀], write ẑ.  If a unit needs the history path the
// controller first spills x to the group's slot row (kOpSpillX).
//
// Ownership: chunk k (CH elements) of every row belongs to CTA (k mod CLUSTER) of the cluster;
// inside a chunk consumer thread t owns pairs t, t+NC, ….  The mapping is identical in every
// sweep, so a thread only ever re-reads global data it wrote itself; data a later bulk copy will
// read is published with __threadfence + the sweep's closing barrier + fence.proxy.async.
#include "muse_iso_ctl.cuh"

namespace muse {

namespace {

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

struct Pipe {
    double* ring;       // stages × 3 × ch doubles
    double* resx;       // slice_cap doubles (resident x of the unit in flight)
    uint64_t* full;     // stages
    uint64_t* empty;    // stages
    uint32_t stage, phase;   // this thread's position in the ring
    __device__ __forceinline__ void advance(int stages) {
        if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
    }
};

// what a sweep streams through the ring: up to three rows (slot 0, 1, 2)
struct Streams {
    const double* v[3];
    uint64_t pol[3];
    int n;      // number of non-null rows
};

template <int CLUSTER>
__device__ __forceinline__ Streams streams_of(const SolveLaunch& L, const Cmd& c, const L2Policy& pol) {
    Streams s;
    s.v[0] = s.v[1] = s.v[2] = nullptr;
    s.pol[0] = s.pol[1] = s.pol[2] = pol.first;
    if (c.op == kOpInit) {
        if (c.xi) { s.v[0] = c.xi; s.v[1] = c.nu; }
        else { s.v[0] = c.xsrc; s.pol[0] = pol.last; }
        if (c.zcur && c.start_kind != kStartTruth) { s.v[2] = c.zcur; s.pol[2] = pol.last; }
    } else {   // kOpTrial
        if (!c.resident) { s.v[0] = c.xsrc; s.pol[0] = pol.last; }
        if (c.zcur) s.v[1] = c.zcur;
        if (!c.lazy) { s.v[2] = c.sbuf; s.pol[2] = pol.last; }
    }
    s.n = (s.v[0] != nullptr) + (s.v[1] != nullptr) + (s.v[2] != nullptr);
    return s;
}

// ---- producer side of a streaming sweep (one elected thread) ----------------------------------
template <int CLUSTER>
__device__ __forceinline__ void produce(const SolveLaunch& L, const Streams& s, Pipe& P, int rank) {
    if (s.n == 0) return;
    fence_proxy_async();    // earlier generic-proxy global writes (published by the closing barrier) → async proxy
    const int ch = L.ch;
    const int nchunks = (L.ld + ch - 1) / ch;
    for (int ck = rank; ck < nchunks; ck += CLUSTER) {
        const int base = ck * ch;
        const int len = min(ch, L.ld - base);
        const uint32_t bytes = (uint32_t)len * 8u;
        mbar_wait(&P.empty[P.stage], P.phase ^ 1u);
        mbar_expect_tx(&P.full[P.stage], bytes * (uint32_t)s.n);
        double* dst = P.ring + (size_t)P.stage * 3 * ch;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (s.v[k]) bulk_g2s(dst + (size_t)k * ch, s.v[k] + base, bytes, &P.full[P.stage], s.pol[k]);
        P.advance(L.stages);
    }
}

// ---- consumer side ----------------------------------------------------------------------------
// Generic chunk walk: fn(j, q2, buf, lc) for every owned pair (element index j even, q2 = 2q offset
// in the chunk, buf = this chunk's stage base or null, lc = local chunk index).
template <int NC, int CLUSTER, bool RING, class F>
__device__ __forceinline__ void consume(const SolveLaunch& L, Pipe& P, int rank, int ct, F&& fn) {
    const int ch = L.ch;
    const int nchunks = (L.ld + ch - 1) / ch;
    int lc = 0;
    for (int ck = rank; ck < nchunks; ck += CLUSTER, ++lc) {
        const int base = ck * ch;
        const int len = min(ch, L.ld - base);
        const double* buf = nullptr;
        if (RING) {
            mbar_wait(&P.full[P.stage], P.phase);
            buf = P.ring + (size_t)P.stage * 3 * ch;
        }
        for (int q2 = 2 * ct; q2 < len; q2 += 2 * NC) {
            const int j = base + q2;
            if (j < L.d) fn(j, q2, buf, lc);
        }
        if (RING) {
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&P.empty[P.stage]);
            P.advance(L.stages);
        }
    }
}

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int NC, int CLUSTER, class G>
__device__ __forceinline__ void consumer_init(G& grp, const SolveLaunch& L, const Cmd& c, Pipe& P, int rank, int ct,
                                              double (&red)[7]) {
    const IsoEval ev = L.ev;
    const IsoSample sp = c.smp;
    const L2Policy pol = make_policies();
    const bool sim = (c.xi != nullptr);
    const int sk = c.start_kind;
    const bool zload = (c.zcur != nullptr) && sk != kStartTruth;
    const bool resident = c.resident != 0;
    double *xw = c.xw, *zA = c.zA, *resx = P.resx;
    const int ch = L.ch, d = L.d;
    double e0 = 0, gg0 = 0, gmax0 = 0, s1 = 0, s2 = 0, e1 = 0, dphi1 = 0;
    auto body = [&](double x, double z0) {
        const Elem a = elem(x, z0, ev);
        e0 += a.e;
        gg0 = fma(a.g, a.g, gg0);
        gmax0 = fmax(gmax0, fabs(a.g));
        s1 += a.w;
        s2 = fma(a.w, a.w, s2);
        const double z1 = z0 - a.g;
        const Elem b = elem(x, z1, ev);
        e1 += b.e;
        dphi1 = fma(b.g, -a.g, dphi1);
    };
    const bool dbg = L.dbg && ct == 0 && rank == 0;
    if (dbg) L.dbg[(size_t)c.item * 16 + 6] = clock64();
    consume<NC, CLUSTER, true>(L, P, rank, ct, [&](int j, int q2, const double* buf, int lc) {
        const double2 a = lds2(buf + q2);
        double2 x, z0 = make_double2(0.0, 0.0);
        if (sim) {
            const double2 b = lds2(buf + ch + q2);
            const double zt0 = fma(sp.sig, a.x, sp.mu), zt1 = fma(sp.sig, a.y, sp.mu);
            x = make_double2(zt0 + b.x, zt1 + b.y);
            if (resident) *reinterpret_cast<double2*>(resx + (size_t)lc * ch + q2) = x;
            else st2_hint(xw, j >> 1, x, pol.last);
            if (sk == kStartTruth) {
                z0 = make_double2(zt0, zt1);
                st2_hint(zA, j >> 1, z0, pol.last);
            }
        } else {
            x = a;
        }
        if (zload) z0 = lds2(buf + 2 * ch + q2);
        if (sk == kStartSharedKeep) st2_hint(zA, j >> 1, z0, pol.last);
        body(x.x, z0.x);
        if (j + 1 < d) body(x.y, z0.y);
    });
    if (dbg) L.dbg[(size_t)c.item * 16 + 7] = clock64();
    __threadfence();
    if (dbg) L.dbg[(size_t)c.item * 16 + 8] = clock64();
    red[0] = e0; red[1] = gg0; red[2] = gmax0; red[3] = s1; red[4] = s2; red[5] = e1; red[6] = dphi1;
    grp.template allreduce<7, 0x04u>(red);
    if (dbg) L.dbg[(size_t)c.item * 16 + 9] = clock64();
}

template <int NC, int CLUSTER, class G, bool RING>
__device__ __forceinline__ void consumer_trial(G& grp, const SolveLaunch& L, const Cmd& cm, Pipe& P, int rank, int ct,
                                               double (&red)[7]) {
    const IsoEval ev = L.ev;
    const L2Policy pol = make_policies();
    const double c = cm.c;
    const bool commit = cm.commit != 0, lazy = cm.lazy != 0, resident = cm.resident != 0;
    const bool zload = cm.zcur != nullptr;
    double* zalt = cm.zalt;
    const double* resx = P.resx;
    const int ch = L.ch, d = L.d;
    double e = 0, dphi = 0, gg_ = 0, gmax_ = 0, s1_ = 0, s2_ = 0, xchg = 0;
    auto body = [&](double x, double z, double s) -> double {
        if (lazy) s = -elem(x, z, ev).g;
        const double zt = fma(c, s, z);
        const Elem b = elem(x, zt, ev);
        e += b.e;
        dphi = fma(b.g, s, dphi);
        gg_ = fma(b.g, b.g, gg_);
        gmax_ = fmax(gmax_, fabs(b.g));
        s1_ += b.w;
        s2_ = fma(b.w, b.w, s2_);
        xchg = fmax(xchg, fabs(zt - z));
        return zt;
    };
    const bool dbg = L.dbg && ct == 0 && rank == 0;
    if (dbg) L.dbg[(size_t)cm.item * 16 + 10] = clock64();
    consume<NC, CLUSTER, RING>(L, P, rank, ct, [&](int j, int q2, const double* buf, int lc) {
        const double2 x = resident ? lds2(resx + (size_t)lc * ch + q2) : lds2(buf + q2);
        double2 z = make_double2(0.0, 0.0), s = make_double2(0.0, 0.0);
        if (RING) {
            if (zload) z = lds2(buf + ch + q2);
            if (!lazy) s = lds2(buf + 2 * ch + q2);
        }
        double2 zt;
        zt.x = body(x.x, z.x, s.x);
        zt.y = (j + 1 < d) ? body(x.y, z.y, s.y) : 0.0;
        if (commit) st2_hint(zalt, j >> 1, zt, pol.first);
    });
    if (dbg) L.dbg[(size_t)cm.item * 16 + 11] = clock64();
    if (commit) __threadfence();
    if (dbg) L.dbg[(size_t)cm.item * 16 + 12] = clock64();
    red[0] = e; red[1] = dphi; red[2] = gg_; red[3] = gmax_; red[4] = s1_; red[5] = s2_; red[6] = xchg;
    grp.template allreduce<7, 0x48u>(red);
    if (dbg) L.dbg[(size_t)cm.item * 16 + 13] = clock64();
}

// element iteration of the slow-path ops in the TMA kernel's ownership
template <int NC, int CLUSTER>
struct OwnedIter {
    const SolveLaunch& L;
    int rank, ct;       // ct < 0: producer warp, owns nothing
    template <class F>
    __device__ __forceinline__ void operator()(F&& fn) const {
        if (ct < 0) return;
        const int ch = L.ch;
        const int nchunks = (L.ld + ch - 1) / ch;
        for (int ck = rank; ck < nchunks; ck += CLUSTER) {
            const int base = ck * ch;
            const int len = min(ch, L.ld - base);
            for (int q2 = 2 * ct; q2 < len; q2 += 2 * NC) {
                const int j = base + q2;
                if (j < L.d) fn(j);
                if (j + 1 < L.d) fn(j + 1);
            }
        }
    }
};

// ring needed for this command?
__device__ __forceinline__ bool uses_ring(const Cmd& c) {
    if (c.op == kOpInit) return true;
    if (c.op == kOpTrial) return !c.resident || c.zcur != nullptr || !c.lazy;
    return false;
}

template <int NC, int CLUSTER, class G>
__device__ __forceinline__ void run_consumer(G& grp, const SolveLaunch& L, const Cmd& c, Pipe& P, int rank, int ct,
                                             double (&red)[7]) {
    if (c.op == kOpInit) {
        consumer_init<NC, CLUSTER>(grp, L, c, P, rank, ct, red);
    } else if (c.op == kOpTrial) {
        if (uses_ring(c)) consumer_trial<NC, CLUSTER, G, true>(grp, L, c, P, rank, ct, red);
        else consumer_trial<NC, CLUSTER, G, false>(grp, L, c, P, rank, ct, red);
    } else if (c.op == kOpSpillX) {
        const L2Policy pol = make_policies();
        consume<NC, CLUSTER, false>(L, P, rank, ct, [&](int j, int q2, const double*, int lc) {
            st2_hint(c.xw, j >> 1, lds2(P.resx + (size_t)lc * L.ch + q2), pol.last);
        });
        __threadfence();
        __syncthreads();
    } else {
        sweep_misc(grp, L, c, red, OwnedIter<NC, CLUSTER>{L, rank, ct});
        // (sweep_misc ends with a reduction or a barrier; global writes are fenced below)
    }
}

// issuer: runs in the controller warp.  Lane 0 publishes the command; the warp then acts as the
// TMA producer of the sweep and joins its closing reduction with neutral contributions.
template <int NC, int CLUSTER, class G>
struct TmaIssuer {
    G& grp;
    const SolveLaunch& L;
    Cmd* scmd;
    Pipe P;
    int rank;
    __device__ TmaIssuer(G& g, const SolveLaunch& l, Cmd* s, const Pipe& p, int r) : grp(g), L(l), scmd(s), P(p), rank(r) {}

    __device__ __noinline__ void operator()(Cmd& cur, double (&red)[7]) {
        const int lane = threadIdx.x & 31;
        if (lane == 0) *scmd = cur;
        G::cmd_barrier();
        const int op = cur.op;
        if ((op == kOpInit || op == kOpTrial) && uses_ring(cur)) {
            if (lane == 0) {
                const L2Policy pol = make_policies();
                const Streams s = streams_of<CLUSTER>(L, cur, pol);
                produce<CLUSTER>(L, s, P, rank);
            }
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) red[k] = 0.0;
        if (op == kOpInit) grp.template allreduce<7, 0x04u>(red);
        else if (op == kOpTrial) grp.template allreduce<7, 0x48u>(red);
        else if (op == kOpSpillX) __syncthreads();
        else sweep_misc(grp, L, *scmd, red, OwnedIter<NC, CLUSTER>{L, rank, -1});
    }
};

template <int NC, int CLUSTER>
__global__ void __launch_bounds__(NC + 32, (NC <= 256 ? 2 : 1))
iso_tma_kernel(const __grid_constant__ SolveLaunch L) {
    using G = Group<NC + 32, false, CLUSTER>;
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ typename G::Smem smem;
    __shared__ Cmd scmd;
    G grp(&smem);

    // carve dynamic shared memory: ring | resident x | full[] | empty[]
    Pipe P;
    P.ring = reinterpret_cast<double*>(dyn);
    P.resx = P.ring + (size_t)L.stages * 3 * L.ch;
    P.full = reinterpret_cast<uint64_t*>(P.resx + L.slice_cap);
    P.empty = P.full + L.stages;
    P.stage = 0;
    P.phase = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < L.stages; ++s) {
            mbar_init(&P.full[s], 1);
            mbar_init(&P.empty[s], NC / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int rank = CLUSTER > 1 ? (int)cg::this_cluster().block_rank() : 0;

    if ((threadIdx.x >> 5) == 0) {
        TmaIssuer<NC, CLUSTER, G> issuer(grp, L, &scmd, P, rank);
        Controller<G, TmaIssuer<NC, CLUSTER, G>> ctl(grp, L, issuer);
        ctl.run_items();
        if ((threadIdx.x & 31) == 0) scmd.op = kOpExit;
        G::cmd_barrier();
    } else {
        const int ct = (int)threadIdx.x - 32;
        for (;;) {
            G::cmd_barrier();
            if (scmd.op == kOpExit) break;
            double red[7];
            run_consumer<NC, CLUSTER>(grp, L, scmd, P, rank, ct, red);
        }
    }
    if (CLUSTER > 1) cg::this_cluster().sync();
}

template <int NC, int CLUSTER>
cudaError_t tma_prepare(int smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(iso_tma_kernel<NC, CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    if (CLUSTER > 8) e = cudaFuncSetAttribute(iso_tma_kernel<NC, CLUSTER>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

template <int NC, int CLUSTER>
cudaError_t tma_launch(const SolveLaunch& L, const Geometry& geo, cudaStream_t st) {
    cudaError_t e = tma_prepare<NC, CLUSTER>(geo.smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)geo.grid);
    cfg.blockDim = dim3(NC + 32);
    cfg.dynamicSmemBytes = (size_t)geo.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (CLUSTER > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        nattr = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
    return cudaLaunchKernelEx(&cfg, iso_tma_kernel<NC, CLUSTER>, L);
}

template <int NC, int CLUSTER>
cudaError_t tma_occupancy(int device, int smem_bytes, int* groups, int* grid) {
    cudaError_t e = tma_prepare<NC, CLUSTER>(smem_bytes);
    if (e != cudaSuccess) return e;
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, iso_tma_kernel<NC, CLUSTER>, NC + 32, (size_t)smem_bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    int ctas = sms * per_sm;
    if (CLUSTER > 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(ctas / CLUSTER * CLUSTER));
        cfg.blockDim = dim3(NC + 32);
        cfg.dynamicSmemBytes = (size_t)smem_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        e = cudaOccupancyMaxActiveClusters(&nclusters, iso_tma_kernel<NC, CLUSTER>, &cfg);
        if (e != cudaSuccess) return e;
        if (nclusters < 1) return cudaErrorInvalidConfiguration;
        ctas = nclusters * CLUSTER;
    }
    *grid = ctas;
    *groups = ctas / CLUSTER;
    return cudaSuccess;
}

}  // namespace

#define MUSE_TMA_VARIANTS(X) \
    X(256, 1)                \
    X(512, 1)                \
    X(256, 2)                \
    X(512, 2)                \
    X(256, 4)                \
    X(512, 4)                \
    X(256, 8)                \
    X(512, 8)

// Choose chunk size, pipeline depth and (if it fits) resident-x mode for a latent dimension.
cudaError_t iso_tma_geometry(int d, int ld, int want_group, int want_cluster, int want_resident, int device,
                             Geometry* geo) {
    (void)d;
    int nc = (want_group == 256 || want_group == 512) ? want_group : 512;
    int cluster = want_cluster > 0 ? want_cluster : 1;
    const int ch = 2 * nc;                       // one pair per consumer thread per chunk
    const int nchunks = (ld + ch - 1) / ch;
    const int per_cta = (nchunks + cluster - 1) / cluster;
    const int slice = per_cta * ch;              // resident x elements per CTA
    const int budget = 200 * 1024;               // dynamic shared memory we allow ourselves per CTA
    int resident = 0, stages = 0;
    if (want_resident != 0) {
        const int left = budget - slice * 8;
        const int st = left / (3 * ch * 8);
        if (st >= 2) { resident = 1; stages = st > 8 ? 8 : st; }
        else if (want_resident > 0) return cudaErrorInvalidConfiguration;
    }
    if (!resident) {
        stages = budget / (nc <= 256 ? 2 : 1) / (3 * ch * 8);      // NC=256: leave room for two CTAs per SM
        if (stages > 8) stages = 8;
        if (stages < 2) stages = 2;
    }
    geo->group_threads = nc;
    geo->cta_threads = nc + 32;
    geo->cluster = cluster;
    geo->tma = 1;
    geo->ch = ch;
    geo->stages = stages;
    geo->resident = resident;
    geo->slice_cap = resident ? slice : 0;
    geo->smem_bytes = stages * 3 * ch * 8 + geo->slice_cap * 8 + 2 * stages * 8 + 128;
#define X(N, C) \
    if (nc == N && cluster == C) return tma_occupancy<N, C>(device, geo->smem_bytes, &geo->groups, &geo->grid);
    MUSE_TMA_VARIANTS(X)
#undef X
    return cudaErrorInvalidConfiguration;
}

cudaError_t launch_iso_tma(const SolveLaunch& L, const Geometry& geo, cudaStream_t st) {
    Geometry g = geo;
    int need_ctas = L.nitems * g.cluster;
    if (need_ctas < g.cluster) need_ctas = g.cluster;
    if (g.grid > need_ctas) g.grid = need_ctas;
#define X(N, C) \
    if (g.group_threads == N && g.cluster == C) return tma_launch<N, C>(L, g, st);
    MUSE_TMA_VARIANTS(X)
#undef X
    return cudaErrorInvalidConfiguration;
}

}  // namespace muse
