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
#include <cmath>

#include "muse_common.cuh"
#include "muse_group.cuh"

namespace muse {

namespace {

constexpr double kEpsD = 2.220446049250313e-16;

__device__ __forceinline__ double next_up(double x) {   // Julia nextfloat(x) for finite x
    if (x == 0.0) return __longlong_as_double(1LL);
    const long long b = __double_as_longlong(x);
    return __longlong_as_double(x > 0.0 ? b + 1 : b - 1);
}
__device__ __forceinline__ double eps_of(double x) {   // Julia eps(x::Float64)
    const double ax = fabs(x);
    return next_up(ax) - ax;
}
__device__ __forceinline__ bool fin(double x) { return isfinite(x); }

// ---- sweep commands ---------------------------------------------------------------------
enum Op : int {
    kOpInit = 0,     // sample + f,g at z₀ + score sums + first trial           → red[7]
    kOpTrial,        // φ(c), φ'(c) [+ commit]                                   → red[7]
    kOpHist,         // dx = c·s → w1, dg = ∇f(v2) − ∇f(v1) → w2                 → red[2] = dx·dg, dg·dg
    kOpGrad,         // sbuf ← ∇f(zcur)
    kOpDot,          // red[0] = v1 · sbuf
    kOpAxpy,         // sbuf ← sbuf + c·v1
    kOpScale,        // sbuf ← c·sbuf
    kOpNegDotG,      // sbuf ← −sbuf; red[0] = ∇f(zcur) · sbuf
    kOpExit,
};

struct Cmd {
    int op;
    int commit;
    int lazy;          // search direction s ≡ −∇f(zcur), not stored
    int start_kind;
    double c;
    IsoSample smp;
    const double* xi;  // null for the data unit
    const double* nu;
    const double* xsrc;   // materialised x (data: xdat; sims: the unit's x row once INIT has run)
    double* xw;           // where INIT materialises x (null for data)
    const double* zcur;   // current iterate (null ⇒ z ≡ 0)
    double* zalt;         // buffer a committed iterate goes to
    double* zA;           // the unit's buffer A (INIT materialises a truth / user start there)
    double* sbuf;         // search direction / two-loop work vector (slot scratch)
    const double* v1;
    const double* v2;
    double* w1;
    double* w2;
};

// elementwise pieces:  g = a (z-μ) - (x-z);   e = (x-z)² + a (z-μ)²
struct Elem {
    double g, e, w;
};
__device__ __forceinline__ Elem elem(double x, double z, const IsoEval& ev) {
    const double r = x - z;
    const double w = z - ev.mu;
    Elem o;
    o.g = fma(ev.a, w, -r);
    o.e = fma(ev.a * w, w, r * r);
    o.w = w;
    return o;
}
__device__ __forceinline__ double2 ld2(const double* p, int i) {
    return p ? *reinterpret_cast<const double2*>(p + 2 * (size_t)i) : make_double2(0.0, 0.0);
}
__device__ __forceinline__ void st2(double* p, int i, double2 v) {
    *reinterpret_cast<double2*>(p + 2 * (size_t)i) = v;
}

// ---- the sweeps (executed by every thread of the group) -----------------------------------
// red: [e0, gg0, gmax0, s1, s2, e1, dphi1]
template <class G>
__device__ __forceinline__ void sweep_init(G& grp, const SolveLaunch& L, const Cmd& c, double (&red)[7]) {
    const IsoEval ev = L.ev;
    const IsoSample sp = c.smp;
    const double *xi = c.xi, *nu = c.nu, *xsrc = c.xsrc, *zcur = c.zcur;
    double *xw = c.xw, *zA = c.zA;
    const bool sim = (xi != nullptr);
    const int sk = c.start_kind;
    const int npairs = L.d >> 1;
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
#pragma unroll 2
    for (int p = grp.tid; p < npairs; p += G::kSize) {
        double2 x, z0;
        if (sim) {
            const double2 a = ld2(xi, p), b = ld2(nu, p);
            const double zt0 = fma(sp.sig, a.x, sp.mu), zt1 = fma(sp.sig, a.y, sp.mu);
            x = make_double2(zt0 + b.x, zt1 + b.y);
            st2(xw, p, x);
            if (sk == kStartTruth) {
                z0 = make_double2(zt0, zt1);
                st2(zA, p, z0);
            } else {
                z0 = ld2(zcur, p);
                if (sk == kStartSharedKeep) st2(zA, p, z0);
            }
        } else {
            x = ld2(xsrc, p);
            z0 = ld2(zcur, p);
            if (sk == kStartSharedKeep) st2(zA, p, z0);
        }
        body(x.x, z0.x);
        body(x.y, z0.y);
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
template <class G, bool LAZY>
__device__ __forceinline__ void sweep_trial(G& grp, const SolveLaunch& L, const Cmd& cm, double (&red)[7]) {
    const IsoEval ev = L.ev;
    const double c = cm.c;
    const bool commit = cm.commit != 0;
    const double *xsrc = cm.xsrc, *zcur = cm.zcur, *sb = cm.sbuf;
    double* zalt = cm.zalt;
    const int npairs = L.d >> 1;
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
#pragma unroll 2
    for (int p = grp.tid; p < npairs; p += G::kSize) {
        const double2 x = ld2(xsrc, p);
        const double2 z = ld2(zcur, p);
        double2 s = make_double2(0.0, 0.0);
        if (!LAZY) s = ld2(sb, p);
        double2 zt;
        zt.x = body(x.x, z.x, s.x);
        zt.y = body(x.y, z.y, s.y);
        if (commit) st2(zalt, p, zt);
    }
    if ((L.d & 1) && grp.tid == 0) {
        const int j = L.d - 1;
        const double zt = body(xsrc[j], zcur ? zcur[j] : 0.0, LAZY ? 0.0 : sb[j]);
        if (commit) zalt[j] = zt;
    }
    red[0] = e; red[1] = dphi; red[2] = gg_; red[3] = gmax_; red[4] = s1_; red[5] = s2_; red[6] = xchg;
    grp.template allreduce<7, 0x48u>(red);
}

template <class G, class F>
__device__ __forceinline__ void for_each_elem(G& grp, const SolveLaunch& L, F&& fn) {
    const int npairs = L.d >> 1;
    for (int p = grp.tid; p < npairs; p += G::kSize) { fn(2 * p); fn(2 * p + 1); }
    if ((L.d & 1) && grp.tid == 0) fn(L.d - 1);
}

// slow-path vector operations (only reached when a unit needs more than one L-BFGS iteration)
template <class G>
__device__ __noinline__ void sweep_misc(G& grp, const SolveLaunch& L, const Cmd& c, double (&red)[7]) {
    const IsoEval ev = L.ev;
    const double* x = c.xsrc;
    double* sb = c.sbuf;
    auto grad = [&](const double* z, int j) { return elem(x[j], z ? z[j] : 0.0, ev).g; };
    switch (c.op) {
        case kOpHist: {
            double a = 0, b = 0;
            const double alpha = c.c;
            const double *zp = c.v1, *zn = c.v2;
            double *dx = c.w1, *dg = c.w2;
            const bool lazy = c.lazy != 0;
            for_each_elem(grp, L, [&](int j) {
                const double gp = grad(zp, j), gn = grad(zn, j);
                const double s = lazy ? -gp : sb[j];
                const double dxj = alpha * s, dgj = gn - gp;
                dx[j] = dxj;
                dg[j] = dgj;
                a = fma(dxj, dgj, a);
                b = fma(dgj, dgj, b);
            });
            double r2[2] = {a, b};
            grp.template allreduce<2, 0u>(r2);
            red[0] = r2[0];
            red[1] = r2[1];
            break;
        }
        case kOpGrad: {
            const double* z = c.zcur;
            for_each_elem(grp, L, [&](int j) { sb[j] = grad(z, j); });
            grp.sync_exec();
            break;
        }
        case kOpDot: {
            double a = 0;
            const double* v = c.v1;
            for_each_elem(grp, L, [&](int j) { a = fma(v[j], sb[j], a); });
            double r1[1] = {a};
            grp.template allreduce<1, 0u>(r1);
            red[0] = r1[0];
            break;
        }
        case kOpAxpy: {
            const double cf = c.c;
            const double* v = c.v1;
            for_each_elem(grp, L, [&](int j) { sb[j] = fma(cf, v[j], sb[j]); });
            grp.sync_exec();
            break;
        }
        case kOpScale: {
            const double cf = c.c;
            for_each_elem(grp, L, [&](int j) { sb[j] *= cf; });
            grp.sync_exec();
            break;
        }
        case kOpNegDotG: {
            double a = 0;
            const double* z = c.zcur;
            for_each_elem(grp, L, [&](int j) {
                const double s = -sb[j];
                sb[j] = s;
                a = fma(grad(z, j), s, a);
            });
            double r1[1] = {a};
            grp.template allreduce<1, 0u>(r1);
            red[0] = r1[0];
            break;
        }
        default: break;
    }
}

template <class G>
__device__ __noinline__ void run_op(G& grp, const SolveLaunch& L, const Cmd& c, double (&red)[7]) {
    if (c.op == kOpInit) sweep_init(grp, L, c, red);
    else if (c.op == kOpTrial) {
        if (c.lazy) sweep_trial<G, true>(grp, L, c, red);
        else sweep_trial<G, false>(grp, L, c, red);
    } else sweep_misc(grp, L, c, red);
}

// ---- the controller: scalar L-BFGS + Hager–Zhang, one warp per CTA ------------------------
struct Red7 {
    double e, dphi, gg, gmax, s1, s2, xchg;
};

template <class G>
struct Controller {
    G& grp;
    const SolveLaunch& L;
    Cmd* scmd;             // shared-memory command slot (CTA / cluster groups)
    Cmd cur;               // the unit's pointers + the command being built
    double* zother;        // the unit's other own buffer (becomes zalt after a flip)
    double* dxh;
    double* dgh;
    // scalar optimiser state
    double f, gg, gmax, s1, s2;
    int fg_evals;
    bool pre_valid;        // first trial prefetched by the INIT sweep
    double pre_phi, pre_dphi;
    double com_alpha;      // last committed trial (NaN ⇒ none)
    Red7 com;
    double last_eval_alpha, last_phi, last_dphi;

    __device__ Controller(G& g, const SolveLaunch& l, Cmd* s) : grp(g), L(l), scmd(s) {}

    // broadcast `cur` and execute it with the whole group
    __device__ __noinline__ void issue(double (&red)[7]) {
        if (G::kWarpGroup) {
            run_op(grp, L, cur, red);
        } else {
            if ((threadIdx.x & 31) == 0) *scmd = cur;
            G::cmd_barrier();
            run_op(grp, L, *scmd, red);
        }
    }

    // φ, φ' at step c (Hager–Zhang's ϕdϕ).  Counts one value+gradient evaluation unless the
    // point equals the last one evaluated (NLSolversBase caching semantics).
    __device__ __noinline__ void phidphi(double c, bool commit, double& phi, double& dphi) {
        if (pre_valid && c == 1.0) {           // prefetched by the INIT sweep
            pre_valid = false;
            phi = pre_phi;
            dphi = pre_dphi;
            fg_evals += 1;
            last_eval_alpha = c;
            last_phi = phi;
            last_dphi = dphi;
            return;
        }
        pre_valid = false;
        if (c == last_eval_alpha && !(commit && com_alpha != c)) {
            phi = last_phi;
            dphi = last_dphi;
            return;
        }
        double red[7];
        cur.op = kOpTrial;
        cur.c = c;
        cur.commit = commit ? 1 : 0;
        issue(red);
        phi = fma(0.5, red[0], L.ev.half_cst);
        dphi = red[1];
        if (c != last_eval_alpha) fg_evals += 1;
        last_eval_alpha = c;
        last_phi = phi;
        last_dphi = dphi;
        if (commit) {
            com_alpha = c;
            com.e = red[0]; com.dphi = red[1]; com.gg = red[2]; com.gmax = red[3];
            com.s1 = red[4]; com.s2 = red[5]; com.xchg = red[6];
        }
    }

    // ------------------------------------------------------------------ Hager–Zhang
    // [EXT LineSearches.jl src/hagerzhang.jl] delta=.1 sigma=.9 alphamax=Inf rho=5 epsilon=1e-6
    // gamma=.66 linesearchmax=50 psi3=.1, mayterminate=false (InitialStatic never sets it).
    // O(1)-state formulation: the upstream routine appends every trial to alphas/values/slopes
    // and addresses them by index; only entries ia, ib, ic and entry 1 are ever read back, and
    // the bracketing scan `for i = ib-1:-1:1` always stops at ib-1 (every point pushed by the
    // expansion branch satisfies value ≤ phi_lim, as does entry 1).  oracle/hagerzhang.py keeps
    // the index-based form; tests compare the two.
    struct Pt {
        double al, phi, dphi;
    };

    __device__ bool wolfe(const Pt& c, double phi_0, double dphi_0, double phi_lim) const {
        constexpr double delta = 0.1, sigma = 0.9;
        const bool w1 = (delta * dphi_0 >= (c.phi - phi_0) / c.al) && (c.dphi >= sigma * dphi_0);
        const bool w2 = ((2 * delta - 1) * dphi_0 >= c.dphi) && (c.dphi >= sigma * dphi_0) && (c.phi <= phi_lim);
        return w1 || w2;
    }

    __device__ __noinline__ void hz_bisect(Pt& a, Pt& b, double phi_lim) {
        while (b.al - a.al > eps_of(b.al)) {
            Pt d;
            d.al = (a.al + b.al) / 2.0;
            phidphi(d.al, false, d.phi, d.dphi);
            if (d.dphi >= 0.0) { b = d; return; }
            if (d.phi <= phi_lim) a = d; else b = d;
        }
    }

    // update!: (a,b) ← best bracket given c
    __device__ __noinline__ void hz_update(Pt& a, Pt& b, const Pt& c, double phi_lim, bool& a_is_c, bool& b_is_c) {
        a_is_c = b_is_c = false;
        if (c.al < a.al || c.al > b.al) return;
        if (c.dphi >= 0.0) { b = c; b_is_c = true; return; }
        if (c.phi <= phi_lim) { a = c; a_is_c = true; return; }
        Pt bb = c;
        hz_bisect(a, bb, phi_lim);
        b = bb;
        b_is_c = (bb.al == c.al);   // bisect! left ib == ic
    }

    // returns 0 ok (alpha, phi_alpha set), 1 LineSearchException (alpha = ex.alpha)
    __device__ __noinline__ int hager_zhang(double c, double phi_0, double dphi_0, double& alpha, double& phi_alpha) {
        constexpr double rho = 5.0, epsilon = 1e-6, gamma = 0.66, psi3 = 0.1;
        constexpr int linesearchmax = 50, iterfinitemax = 53;   // ceil(-log2(eps))
        alpha = 0.0;
        phi_alpha = phi_0;
        if (!(fin(phi_0) && fin(dphi_0))) return 1;
        if (dphi_0 >= kEpsD * fabs(phi_0)) return 1;
        else if (dphi_0 >= 0.0) return 0;
        const double phi_lim = phi_0 + epsilon * fabs(phi_0);
        if (c <= kEpsD) return 0;
        Pt p0{0.0, phi_0, dphi_0};
        Pt pc;
        pc.al = c;
        phidphi(pc.al, false, pc.phi, pc.dphi);
        int iterfinite = 1;
        while (!(fin(pc.phi) && fin(pc.dphi)) && iterfinite < iterfinitemax) {
            iterfinite += 1;
            pc.al *= psi3;
            phidphi(pc.al, false, pc.phi, pc.dphi);
        }
        if (!(fin(pc.phi) && fin(pc.dphi))) return 0;   // alpha = 0

        // bracketing (B0-B3)
        bool isbracketed = false;
        Pt a = p0, b = pc, prev = p0;
        int iter = 1;
        while (!isbracketed && iter < linesearchmax) {
            if (pc.dphi >= 0.0) {
                b = pc;
                a = prev;
                isbracketed = true;
            } else if (pc.phi > phi_lim) {
                a = p0;
                b = pc;
                hz_bisect(a, b, phi_lim);
                isbracketed = true;
            } else {
                const Pt cold = pc;
                Pt nc;
                nc.al = pc.al * rho;
                phidphi(nc.al, false, nc.phi, nc.dphi);
                iterfinite = 1;
                while (!(fin(nc.phi) && fin(nc.dphi)) && nc.al > next_up(cold.al) && iterfinite < iterfinitemax) {
                    iterfinite += 1;
                    nc.al = (cold.al + nc.al) / 2.0;
                    phidphi(nc.al, false, nc.phi, nc.dphi);
                }
                if (!(fin(nc.phi) && fin(nc.dphi))) {
                    alpha = cold.al;
                    phi_alpha = cold.phi;
                    return 0;
                }
                prev = cold;
                pc = nc;
            }
            iter += 1;
        }

        // secant² / bisection
        while (iter < linesearchmax) {
            if (b.al - a.al <= eps_of(b.al)) {
                alpha = a.al;
                phi_alpha = a.phi;
                return 0;
            }
            // ---- secant2!
            Pt A = a, B = b;
            bool iswolfe = false;
            {
                Pt cc;
                cc.al = (a.al * b.dphi - b.al * a.dphi) / (b.dphi - a.dphi);
                phidphi(cc.al, true, cc.phi, cc.dphi);
                if (wolfe(cc, phi_0, dphi_0, phi_lim)) {
                    iswolfe = true;
                    A = B = cc;
                } else {
                    bool a_is_c, b_is_c;
                    hz_update(A, B, cc, phi_lim, a_is_c, b_is_c);
                    double c2 = 0.0;
                    if (b_is_c) c2 = (b.al * B.dphi - B.al * b.dphi) / (B.dphi - b.dphi);
                    else if (a_is_c) c2 = (a.al * A.dphi - A.al * a.dphi) / (A.dphi - a.dphi);
                    if ((a_is_c || b_is_c) && A.al <= c2 && c2 <= B.al) {
                        Pt c2p;
                        c2p.al = c2;
                        phidphi(c2p.al, true, c2p.phi, c2p.dphi);
                        if (wolfe(c2p, phi_0, dphi_0, phi_lim)) {
                            iswolfe = true;
                            A = B = c2p;
                        } else {
                            bool x1, x2;
                            hz_update(A, B, c2p, phi_lim, x1, x2);
                        }
                    }
                }
            }
            if (iswolfe) {
                alpha = A.al;
                phi_alpha = A.phi;
                return 0;
            }
            if (B.al - A.al < gamma * (b.al - a.al)) {
                if (next_up(a.phi) >= b.phi && next_up(A.phi) >= B.phi) {
                    alpha = A.al;
                    phi_alpha = A.phi;
                    return 0;
                }
                a = A;
                b = B;
            } else {
                Pt m;
                m.al = (A.al + B.al) / 2.0;
                phidphi(m.al, false, m.phi, m.dphi);
                bool x1, x2;
                hz_update(A, B, m, phi_lim, x1, x2);
                a = A;
                b = B;
            }
            iter += 1;
        }
        alpha = a.al;    // LineSearchException(alphas[ia])
        return 1;
    }

    // ------------------------------------------------------------------ L-BFGS pieces
    // twoloop!: s ← −H·∇f(zcur) into sbuf; returns ∇f·s   [EXT Optim.jl l_bfgs.jl twoloop!]
    __device__ __noinline__ double twoloop(int pseudo_iter, const double* rho, const double* dxdg_h,
                                           const double* dgdg_h, double* alpha_tl) {
        const int m = L.lbfgs_m;
        const int lower = pseudo_iter - m, upper = pseudo_iter - 1;
        double red[7];
        cur.op = kOpGrad;
        issue(red);
        for (int index = upper; index >= lower; --index) {
            if (index < 1) continue;
            const int i = (index - 1) % m;
            cur.op = kOpDot;
            cur.v1 = dxh + (size_t)i * L.ld;
            issue(red);
            const double al = rho[i] * red[0];
            alpha_tl[i] = al;
            cur.op = kOpAxpy;
            cur.c = -al;
            cur.v1 = dgh + (size_t)i * L.ld;
            issue(red);
        }
        if (pseudo_iter > 1) {     // scaleinvH0
            const int i = (upper - 1) % m;
            cur.op = kOpScale;
            cur.c = dxdg_h[i] / dgdg_h[i];
            issue(red);
        }
        for (int index = lower; index <= upper; ++index) {
            if (index < 1) continue;
            const int i = (index - 1) % m;
            cur.op = kOpDot;
            cur.v1 = dgh + (size_t)i * L.ld;
            issue(red);
            const double beta = rho[i] * red[0];
            cur.op = kOpAxpy;
            cur.c = alpha_tl[i] - beta;
            cur.v1 = dxh + (size_t)i * L.ld;
            issue(red);
        }
        cur.op = kOpNegDotG;
        issue(red);
        return red[0];
    }

    // ------------------------------------------------------------------ one unit
    // `cur` holds the unit's pointers (xi, nu, xsrc, xw, zcur, zalt, zA, sbuf, smp, start_kind).
    __device__ __noinline__ void solve(int item, int* zstate_row) {
        const IsoEval ev = L.ev;
        double red[7];
        com_alpha = NAN;
        last_eval_alpha = NAN;
        last_phi = last_dphi = NAN;

        cur.op = kOpInit;
        cur.lazy = 1;
        cur.commit = 0;
        issue(red);
        if (cur.xw) cur.xsrc = cur.xw;
        f = fma(0.5, red[0], ev.half_cst);
        gg = red[1];
        gmax = red[2];
        s1 = red[3];
        s2 = red[4];
        fg_evals = 1;
        pre_valid = true;
        pre_phi = fma(0.5, red[5], ev.half_cst);
        pre_dphi = red[6];

        // where the start vector now lives
        int zst;    // ZState of the current iterate if it is one of the unit's own buffers, else -1
        if (cur.start_kind == kStartZero) zst = kZZero;
        else if (cur.start_kind == kStartOwn) zst = *zstate_row;
        else if (cur.start_kind == kStartTruth || cur.start_kind == kStartSharedKeep) { zst = kZA; cur.zcur = cur.zA; }
        else zst = -1;

        int status = MUSE_STATUS_G_CONVERGED;
        int iter = 0;
        bool stopped = !fin(f) || !fin(gg);
        bool converged = gmax <= L.atol;
        if (stopped) status = MUSE_STATUS_NONFINITE;

        // L-BFGS bookkeeping (m ≤ 16)
        double rho[16], dxdg_h[16], dgdg_h[16], alpha_tl[16];
        int pseudo_iter = 0;
        int counter_f_tol = 0;

        while (!converged && !stopped && iter < L.max_iters) {
            iter += 1;
            pseudo_iter += 1;
            double dphi_0;
            if (pseudo_iter > 1) {
                cur.lazy = 0;
                dphi_0 = twoloop(pseudo_iter, rho, dxdg_h, dgdg_h, alpha_tl);
                pre_valid = false;
            } else {
                cur.lazy = 1;
                dphi_0 = -gg;
            }
            if (dphi_0 >= 0.0 && pseudo_iter > 1) {      // reset_search_direction!
                pseudo_iter = 1;
                cur.lazy = 1;
                dphi_0 = -gg;
            }
            const double phi_0 = f;
            const double f_prev = f;
            com_alpha = NAN;
            last_eval_alpha = NAN;
            double alpha, phi_alpha;
            const int ls = hager_zhang(1.0, phi_0, dphi_0, alpha, phi_alpha);   // InitialStatic(alpha = 1)
            pre_valid = false;

            const double* zprev = cur.zcur;
            if (alpha == 0.0) {
                com.xchg = 0.0;                           // x unchanged
                if (ls != 0) { status = MUSE_STATUS_LS_FAILED; break; }
            } else {
                if (!(com_alpha == alpha)) {              // accepted point is not the last committed trial
                    const bool need_eval = (ls == 0) && !(last_eval_alpha == alpha);
                    cur.op = kOpTrial;
                    cur.c = alpha;
                    cur.commit = 1;
                    issue(red);
                    com.e = red[0]; com.dphi = red[1]; com.gg = red[2]; com.gmax = red[3];
                    com.s1 = red[4]; com.s2 = red[5]; com.xchg = red[6];
                    com_alpha = alpha;
                    if (need_eval) fg_evals += 1;
                }
                // flip buffers
                double* newcur = cur.zalt;
                cur.zalt = zother;
                zother = newcur;
                cur.zcur = newcur;
                zst = (newcur == cur.zA) ? kZA : kZB;
                if (ls != 0) {      // linesearch exception: x moved, objective not re-evaluated
                    status = MUSE_STATUS_LS_FAILED;
                    s1 = com.s1; s2 = com.s2; gmax = com.gmax;   // report at the point returned
                    break;
                }
                f = fma(0.5, com.e, ev.half_cst);
                gg = com.gg;
                gmax = com.gmax;
                s1 = com.s1;
                s2 = com.s2;
            }
            // assess_convergence  [EXT Optim.jl]
            const bool x_conv = com.xchg <= 0.0;
            const bool f_conv = fabs(f - f_prev) <= 0.0;
            const bool g_conv = gmax <= L.atol;
            counter_f_tol = f_conv ? counter_f_tol + 1 : 0;
            converged = x_conv || g_conv || (counter_f_tol > 1);
            if (converged) status = g_conv ? MUSE_STATUS_G_CONVERGED : MUSE_STATUS_XF_CONVERGED;
            if (!fin(f) || !fin(gg)) { status = MUSE_STATUS_NONFINITE; break; }
            // update_h! (no observable effect once the loop is about to end)
            if (!converged && iter < L.max_iters) {
                if (alpha == 0.0) {
                    pseudo_iter = 0;                     // dx·dg = 0 ⇒ rho = Inf
                } else {
                    const int idx = (pseudo_iter - 1) % L.lbfgs_m;
                    cur.op = kOpHist;
                    cur.c = alpha;
                    cur.v1 = zprev;
                    cur.v2 = cur.zcur;
                    cur.w1 = dxh + (size_t)idx * L.ld;
                    cur.w2 = dgh + (size_t)idx * L.ld;
                    issue(red);
                    const double dxdg = red[0], dgdg = red[1];
                    const double rho_it = 1.0 / dxdg;
                    if (isinf(rho_it)) pseudo_iter = 0;
                    else { rho[idx] = rho_it; dxdg_h[idx] = dxdg; dgdg_h[idx] = dgdg; }
                }
            }
        }
        if (!converged && !stopped && status == MUSE_STATUS_G_CONVERGED && iter >= L.max_iters)
            status = MUSE_STATUS_MAXITER;

        // outputs
        if (grp.tid == 0) {
            if (zstate_row && zst >= 0) *zstate_row = zst;
            double* g = L.g_out + (size_t)item * L.ntheta;
            if (L.family == MUSE_FAMILY_FUNNEL) {
                g[0] = 0.5 * ev.a * s2 - 0.5 * (double)L.d;
            } else {
                g[0] = ev.a * s1;
                g[1] = ev.a * s2 - (double)L.d;
            }
            L.iters_out[item] = iter;
            L.fg_out[item] = fg_evals;
            L.gnorm_out[item] = gmax;
            L.f_out[item] = f;
            L.status_out[item] = status;
        }
    }

    // ------------------------------------------------------------------ unit setup
    __device__ __noinline__ void run_items() {
        const int gi = grp.group_index();
        const int gn = grp.group_count();
        const size_t ld = (size_t)L.ld;
        cur.sbuf = L.sbuf + (size_t)gi * ld;
        dxh = L.dxh + (size_t)gi * L.lbfgs_m * ld;
        dgh = L.dgh + (size_t)gi * L.lbfgs_m * ld;
        cur.v1 = cur.v2 = nullptr;
        cur.w1 = cur.w2 = nullptr;
        cur.c = 0.0;

        const double* zshared = L.zshared;
        if (L.zshared_state) {       // shared start = result of an earlier launch (fiducial ẑ)
            const int st = *L.zshared_state;
            zshared = st == kZA ? L.zsharedA : (st == kZB ? L.zsharedB : nullptr);
        }

        for (int item = gi; item < L.nitems; item += gn) {
            int row, draw, tsel = 0;
            if (L.mode == 0) {
                if (L.include_data && item == 0) { row = 0; draw = -1; }
                else {
                    const int k = L.first_sim + item - (L.include_data ? 1 : 0);
                    row = 1 + k;
                    draw = k;
                }
            } else if (L.mode == 1) {
                // finite-difference virtual sims: item = (k·ntheta + n)·2 + sgn, θ_sim = smp[2n + sgn]
                tsel = item % (2 * L.ntheta);
                row = item;
                draw = item / (2 * L.ntheta);
            } else {
                // the master stream's own draw (fiducial solve of get_H!, src/muse.jl:418)
                row = 0;
                draw = L.master_row;
            }
            cur.smp = L.smp[tsel];
            cur.start_kind = (draw < 0 && L.start_kind == kStartTruth) ? kStartZero : L.start_kind;
            cur.xi = draw >= 0 ? L.xi + (size_t)draw * ld : nullptr;
            cur.nu = draw >= 0 ? L.nu + (size_t)draw * ld : nullptr;
            cur.xw = draw >= 0 ? L.x + (size_t)row * ld : nullptr;
            cur.xsrc = draw >= 0 ? cur.xw : L.xdat;
            double* zA = L.zA + (size_t)row * ld;
            double* zB = L.zB + (size_t)row * ld;
            cur.zA = zA;
            int* zs = L.zstate ? L.zstate + row : nullptr;
            switch (cur.start_kind) {
                case kStartOwn: {
                    const int st = *zs;
                    cur.zcur = st == kZZero ? nullptr : (st == kZA ? zA : zB);
                    cur.zalt = st == kZA ? zB : zA;
                    zother = st == kZA ? zA : zB;
                    break;
                }
                case kStartShared:
                    cur.zcur = zshared; cur.zalt = zA; zother = zB; break;
                case kStartSharedKeep:
                    cur.zcur = zshared; cur.zalt = zB; zother = zA; break;
                case kStartTruth:
                    cur.zcur = nullptr; cur.zalt = zB; zother = zA; break;
                default:   // zeros
                    cur.zcur = nullptr; cur.zalt = zA; zother = zB; break;
            }
            solve(item, zs);
        }
    }
};

template <int CTA_THREADS, bool WARP_GROUP, int CLUSTER>
__global__ void __launch_bounds__(CTA_THREADS, (CTA_THREADS >= 1024 ? 1 : (CTA_THREADS == 512 ? 2 : (CTA_THREADS == 256 ? 3 : 6))))
iso_solver_kernel(const __grid_constant__ SolveLaunch L) {
    using G = Group<CTA_THREADS, WARP_GROUP, CLUSTER>;
    __shared__ typename G::Smem smem;
    __shared__ Cmd scmd;
    G grp(&smem);

    if (WARP_GROUP || (threadIdx.x >> 5) == 0) {
        // controller warp (every warp, for warp groups)
        Controller<G> ctl(grp, L, &scmd);
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
    X(512, false, 1)          \
    X(1024, false, 1)         \
    X(512, false, 2)          \
    X(512, false, 4)          \
    X(1024, false, 2)

cudaError_t iso_solver_geometry(int d, int want_group, int want_cluster, int device, Geometry* geo) {
    int group = want_group, cluster = want_cluster;
    if (group <= 0) {
        if (d <= 2048) group = 32;
        else if (d <= 16384) group = 256;
        else group = 512;
    }
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
