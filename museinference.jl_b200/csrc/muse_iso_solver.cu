// muse_iso_solver.cu — persistent batched MAP + score solver for the isotropic-Gaussian-latent
// families (F1 Neal's funnel, F2 hierarchical Gaussian), sm_100a.
//
// What it replaces.  One *unit* is one execution of the body the reference maps over its worker
// pool (/root/reference/src/muse.jl:170-175, :510-513, :430-432):
//     x  ← prob.x | sample_x_z(rng_k, θ_sim).x                    (src/simple.jl:61-65)
//     ẑ  ← Optim.optimize(only_fg(z -> .-logLike_and_∇z_logLike(x,z,θ)), z₀, LBFGS(),
//                         Options(g_tol = atol)).minimizer         (src/interface.jl:162-166)
//     g  ← ∇θ_logLike(x, ẑ, θ)                                     (src/simple.jl:92)
// The reference gets ∇z and ∇θ by AD of `logLike` (src/simple.jl:84-85); here they are the
// analytic derivatives, fused into the same passes over memory as the objective value.
//
// Design.  One launch solves all units.  A *group* of threads (a warp, a CTA or a
// thread-block cluster, by latent dimension) owns one unit at a time and runs the complete
// L-BFGS(m) + Hager–Zhang algorithm for it.  The scalar optimiser logic is written in direct
// style (it reads like oracle/lbfgs.py + oracle/hagerzhang.py) and runs only in the *controller*
// warp of each CTA; every objective evaluation is a *sweep* command the controller broadcasts
// through shared memory and that all warps execute cooperatively: one 128-bit vectorised pass
// over the unit's rows with a fixed-tree all-reduce (muse_group.cuh).  No host round trip, no
// inter-kernel state machine; units that need more iterations simply keep their group longer.
// (Keeping the scalar state in one warp matters: when every thread carried it, its spills and
// stack traffic added ~30 % DRAM writes — profiles/r01_*.)
//
// Traffic minimisation.  For these families ∇z is elementwise in (x_j, z_j), so the gradient is
// never stored: it is recomputed wherever it is needed.  Fusions, in reference terms:
//   * INIT sweep   = sample_x_z + initial value_gradient!! + the first line-search trial
//                    φ(1), φ'(1) along −g₀ (InitialStatic α=1 and the first L-BFGS direction
//                    −g are known a priori)               reads ξ, ν [, z₀]; writes x
//   * TRIAL sweep  = φ(c), φ'(c) of Hager–Zhang; trials issued by the secant² stage also
//                    commit z + c·s to the unit's other ẑ buffer together with ‖∇z‖_∞, the score
//                    sums and max|Δz|, so an accepted step costs no further sweep
//                                                          reads x, z [, s]; writes ẑ
// For the registered (quadratic, isotropic-Hessian) families the algorithm takes 1 iteration /
// 3 evaluations, i.e. exactly INIT + one committed TRIAL per unit.  Everything else (two-loop
// recursion over the (dx, dg) history kept in per-slot scratch, bisection, bracket expansion,
// direction resets) is implemented and exercised by tests but is not on the fast path.
#include "muse_iso_ctl.cuh"

namespace muse {

namespace {

// [host-test:begin sweeps]
// ---- the sweeps (executed by every thread of the group) -----------------------------------
// red: [e0, gg0, gmax0, s1, s2, e1, dphi1]
template <class G>
__device__ __forceinline__ void sweep_init(G& grp, const SolveLaunch& L, const Cmd& c, double (&red)[7]) {
    const IsoEval ev = launch_ev(L);
    const IsoSample sp = c.smp;
    const double *xi = c.xi, *nu = c.nu, *xsrc = c.xsrc, *zcur = c.zcur;
    double *xw = c.xw, *zA = c.zA;
    const bool sim = (xi != nullptr);
    const int sk = c.start_kind;
    const int npairs = L.d >> 1;
    const L2Policy pol = make_policies();
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
    // U independent iterations are loaded before any is consumed (memory-level parallelism:
    // U × 32-48 B in flight per thread); the compiler keeps the batch in registers.
    constexpr int U = kBatch;
    for (int p0 = grp.tid; p0 < npairs; p0 += U * G::kSize) {
        double2 a[U], b[U], z0[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * G::kSize;
            const bool ok = p < npairs;
            if (sim) {
                a[u] = ok ? ld2_stream(xi, p, pol.first) : make_double2(0.0, 0.0);
                b[u] = ok ? ld2_stream(nu, p, pol.first) : make_double2(0.0, 0.0);
            } else {
                a[u] = ok ? ld2(xsrc, p) : make_double2(0.0, 0.0);
            }
            z0[u] = (ok && zcur && sk != kStartTruth) ? ld2_hint(zcur, p, pol.last) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * G::kSize;
            if (p < npairs) {
                double2 x;
                if (sim) {
                    const double zt0 = fma(sp.sig, a[u].x, sp.mu), zt1 = fma(sp.sig, a[u].y, sp.mu);
                    x = make_double2(zt0 + b[u].x, zt1 + b[u].y);
                    st2_hint(xw, p, x, pol.last);
                    if (sk == kStartTruth) {
                        z0[u] = make_double2(zt0, zt1);
                        st2_hint(zA, p, z0[u], pol.last);
                    }
                } else {
                    x = a[u];
                }
                if (sk == kStartSharedKeep) st2_hint(zA, p, z0[u], pol.last);
                body(x.x, z0[u].x);
                body(x.y, z0[u].y);
            }
        }
    }
    if ((L.d & 1) && grp.tid == 0) {
        const int j = L.d - 1;
        double x, z0;
        if (sim) {
            const double zt = fma(sp.sig, xi[j], sp.mu);
            x = zt + nu[j];
            xw[j] = x;
            if (sk == kStartTruth) { z0 = zt; zA[j] = z0; }
            else { z0 = zcur ? zcur[j] : 0.0; if (sk == kStartSharedKeep) zA[j] = z0; }
        } else {
            x = xsrc[j];
            z0 = zcur ? zcur[j] : 0.0;
            if (sk == kStartSharedKeep) zA[j] = z0;
        }
        body(x, z0);
    }
    red[0] = e0; red[1] = gg0; red[2] = gmax0; red[3] = s1; red[4] = s2; red[5] = e1; red[6] = dphi1;
    grp.template allreduce<7, 0x04u>(red);
}

// zt = zcur + c·s with s = −∇f(zcur) (lazy) or sbuf.  red: [e, dphi, gg, gmax, s1, s2, xchg]
template <class G, bool LAZY, bool ZNULL>
__device__ __forceinline__ void sweep_trial(G& grp, const SolveLaunch& L, const Cmd& cm, double (&red)[7]) {
    const IsoEval ev = launch_ev(L);
    const double c = cm.c;
    const bool commit = cm.commit != 0;
    const double *xsrc = cm.xsrc, *zcur = cm.zcur, *sb = cm.sbuf;
    double* zalt = cm.zalt;
    const int npairs = L.d >> 1;
    const L2Policy pol = make_policies();
    double e = 0, dphi = 0, gg_ = 0, gmax_ = 0, s1_ = 0, s2_ = 0, xchg = 0;
    auto body = [&](double x, double z, double s) -> double {
        if (LAZY) s = -elem(x, z, ev).g;
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
    constexpr int U = (LAZY && ZNULL) ? 2 * kBatch : kBatch;   // cold start: only x is loaded, keep as many bytes in flight
    for (int p0 = grp.tid; p0 < npairs; p0 += U * G::kSize) {
        double2 x[U], z[U], sv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * G::kSize;
            const bool ok = p < npairs;
            x[u] = ok ? ld2_hint(xsrc, p, pol.last) : make_double2(0.0, 0.0);
            z[u] = (!ZNULL && ok) ? ld2_hint(zcur, p, pol.first) : make_double2(0.0, 0.0);
            sv[u] = (!LAZY && ok) ? ld2(sb, p) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * G::kSize;
            if (p < npairs) {
                double2 zt;
                zt.x = body(x[u].x, z[u].x, sv[u].x);
                zt.y = body(x[u].y, z[u].y, sv[u].y);
                if (commit) st2_hint(zalt, p, zt, pol.first);
            }
        }
    }
    if ((L.d & 1) && grp.tid == 0) {
        const int j = L.d - 1;
        const double zt = body(xsrc[j], (!ZNULL && zcur) ? zcur[j] : 0.0, LAZY ? 0.0 : sb[j]);
        if (commit) zalt[j] = zt;
    }
    red[0] = e; red[1] = dphi; red[2] = gg_; red[3] = gmax_; red[4] = s1_; red[5] = s2_; red[6] = xchg;
    grp.template allreduce<7, 0x48u>(red);
}

// element iteration of the register-loop kernel: pairs strided over the group's threads
template <class G>
struct StridedIter {
    G& grp;
    const SolveLaunch& L;
    template <class F>
    __device__ __forceinline__ void operator()(F&& fn) const {
        const int npairs = L.d >> 1;
        for (int p = grp.tid; p < npairs; p += G::kSize) { fn(2 * p); fn(2 * p + 1); }
        if ((L.d & 1) && grp.tid == 0) fn(L.d - 1);
    }
};

template <class G>
__device__ __noinline__ void run_op(G& grp, const SolveLaunch& L, const Cmd& c, double (&red)[7]) {
    if (c.op == kOpInit) sweep_init(grp, L, c, red);
    else if (c.op == kOpTrial) {
        if (c.lazy) {
            if (c.zcur) sweep_trial<G, true, false>(grp, L, c, red);
            else sweep_trial<G, true, true>(grp, L, c, red);
        } else {
            if (c.zcur) sweep_trial<G, false, false>(grp, L, c, red);
            else sweep_trial<G, false, true>(grp, L, c, red);
        }
    } else sweep_misc(grp, L, c, red, StridedIter<G>{grp, L});
}

// [host-test:end sweeps]

// issuer of the register-loop kernel: broadcast the command, run it with every thread of the group
template <class G>
struct RegIssuer {
    G& grp;
    const SolveLaunch& L;
    Cmd* scmd;
    __device__ RegIssuer(G& g, const SolveLaunch& l, Cmd* s) : grp(g), L(l), scmd(s) {}
    __device__ __noinline__ void operator()(Cmd& cur, double (&red)[7]) {
        if (G::kWarpGroup) {
            run_op(grp, L, cur, red);
        } else {
            if ((threadIdx.x & 31) == 0) *scmd = cur;
            G::cmd_barrier();
            run_op(grp, L, *scmd, red);
        }
    }
};

template <int CTA_THREADS, bool WARP_GROUP, int CLUSTER>
__global__ void __launch_bounds__(CTA_THREADS, (CTA_THREADS >= 512 ? 1 : (CTA_THREADS == 256 ? 2 : 4)))
iso_solver_kernel(const __grid_constant__ SolveLaunch L) {
    if (launch_skipped(L)) return;      // uniform over the grid (and over a cluster): nobody reaches a barrier
    using G = Group<CTA_THREADS, WARP_GROUP, CLUSTER>;
    __shared__ typename G::Smem smem;
    __shared__ Cmd scmd;
    G grp(&smem);

    if (WARP_GROUP || (threadIdx.x >> 5) == 0) {
        // controller warp (every warp, for warp groups)
        RegIssuer<G> issuer(grp, L, &scmd);
        Controller<G, RegIssuer<G>> ctl(grp, L, issuer);
        ctl.run_items();
        if (!WARP_GROUP) {
            if ((threadIdx.x & 31) == 0) scmd.op = kOpExit;
            G::cmd_barrier();
        }
    } else {
        // worker warps: execute the sweeps the controller broadcasts
        for (;;) {
            G::cmd_barrier();
            if (scmd.op == kOpExit) break;
            double red[7];
            run_op(grp, L, scmd, red);
        }
    }
    if (CLUSTER > 1) cg::this_cluster().sync();   // keep DSMEM alive until every CTA is done
}

template <int CTA_THREADS, bool WARP_GROUP, int CLUSTER>
cudaError_t launch_variant(const SolveLaunch& L, const Geometry& geo, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)geo.grid);
    cfg.blockDim = dim3(CTA_THREADS);
    cfg.dynamicSmemBytes = 0;
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
    return cudaLaunchKernelEx(&cfg, iso_solver_kernel<CTA_THREADS, WARP_GROUP, CLUSTER>, L);
}

template <int CTA_THREADS, bool WARP_GROUP, int CLUSTER>
cudaError_t occupancy_variant(int device, int* groups, int* grid) {
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, iso_solver_kernel<CTA_THREADS, WARP_GROUP, CLUSTER>,
                                                      CTA_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int ctas = sms * per_sm;
    if (CLUSTER > 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(ctas / CLUSTER * CLUSTER));
        cfg.blockDim = dim3(CTA_THREADS);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        e = cudaOccupancyMaxActiveClusters(&nclusters, iso_solver_kernel<CTA_THREADS, WARP_GROUP, CLUSTER>, &cfg);
        if (e != cudaSuccess) return e;
        if (nclusters < 1) nclusters = 1;
        ctas = nclusters * CLUSTER;
    }
    *grid = ctas;
    *groups = WARP_GROUP ? ctas * (CTA_THREADS / 32) : ctas / CLUSTER;
    return cudaSuccess;
}

}  // namespace

// Variant table: (group shape) → kernel instantiation.
#define MUSE_VARIANTS(X)      \
    X(128, true, 1)           \
    X(256, false, 1)          \
    X(256, false, 2)          \
    X(256, false, 4)          \
    X(512, false, 1)          \
    X(1024, false, 1)         \
    X(512, false, 2)          \
    X(512, false, 4)          \
    X(1024, false, 2)

cudaError_t iso_solver_geometry(int d, int want_group, int want_cluster, int device, Geometry* geo) {
    int group = want_group, cluster = want_cluster;
    if (group <= 0) group = (d <= 2048) ? 32 : 256;   // measured on C3 (profiles/r01): 256 threads × 2 CTAs/SM
    if (cluster <= 0) cluster = 1;
    if (group == 32) cluster = 1;
    geo->group_threads = group;
    geo->cluster = cluster;
    geo->cta_threads = (group == 32) ? 128 : group;
#define X(T, W, C)                                                                       \
    if (geo->cta_threads == T && (group == 32) == W && cluster == C)                     \
        return occupancy_variant<T, W, C>(device, &geo->groups, &geo->grid);
    MUSE_VARIANTS(X)
#undef X
    return cudaErrorInvalidConfiguration;
}

cudaError_t launch_iso_solver(const SolveLaunch& L, const Geometry& geo, cudaStream_t st) {
    Geometry g = geo;
    // never launch more groups than units
    const int per_cta = (g.group_threads == 32) ? g.cta_threads / 32 : 1;
    int need_ctas = (g.group_threads == 32) ? (L.nitems + per_cta - 1) / per_cta : L.nitems * g.cluster;
    if (need_ctas < g.cluster) need_ctas = g.cluster;
    if (g.grid > need_ctas) g.grid = need_ctas;
#define X(T, W, C)                                                                       \
    if (g.cta_threads == T && (g.group_threads == 32) == W && g.cluster == C)            \
        return launch_variant<T, W, C>(L, g, st);
    MUSE_VARIANTS(X)
#undef X
    return cudaErrorInvalidConfiguration;
}

}  // namespace muse
