// muse_iso_ctl.cuh — the scalar optimiser ("controller") shared by the solver kernels, and the
// sweep command it issues.  See muse_iso_solver.cu for the overall design.
#pragma once
#include <cmath>

#include "muse_common.cuh"
#include "muse_group.cuh"

namespace muse {
namespace {

// [host-test:begin ctl-a]  (tests/test_generic_solver_host.py compiles the marked blocks for the host)
constexpr double kEpsD = 2.220446049250313e-16;
#ifndef MUSE_BATCH
#define MUSE_BATCH 4
#endif
constexpr int kBatch = MUSE_BATCH;   // independent 16-byte loads per vector per thread kept in flight

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
    int item;          // unit index of the launch (diagnostics)
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

// [host-test:end ctl-a]
// L2 residency control.  The base normals are streamed once per launch
// (evict_first); the unit's materialised x lives in a per-slot scratch row that is rewritten by
// the next unit of the same group, so with evict_last it stays in the 126 MB L2 between the INIT
// and TRIAL sweeps and is never written back to HBM; the final ẑ is a streaming store.
struct L2Policy {
    uint64_t first, last;
};
__device__ __forceinline__ L2Policy make_policies() {
    L2Policy p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.first));
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.last));
    return p;
}
__device__ __forceinline__ double2 ld2_hint(const double* p, int i, uint64_t pol) {
    double2 v;
    asm("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
        : "=d"(v.x), "=d"(v.y)
        : "l"(p + 2 * (size_t)i), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 ld2_stream(const double* p, int i, uint64_t pol) {
    double2 v;
    asm("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
        : "=d"(v.x), "=d"(v.y)
        : "l"(p + 2 * (size_t)i), "l"(pol));
    return v;
}
__device__ __forceinline__ void st2_hint(double* p, int i, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;"
                 :
                 : "l"(p + 2 * (size_t)i), "d"(v.x), "d"(v.y), "l"(pol));
}

// a unit's ẑ-state cell (or the fiducial ẑ's): written by an earlier phase of the SAME launch in solve_persist_kernel, so it
// is read at the L2, never from a stale L1 line
__device__ __forceinline__ int ld_state(const int* p) { return __ldcg(p); }

// [host-test:begin ctl-b]
// slow-path vector operations (only reached when a unit needs more than one L-BFGS iteration)
template <class G, class Iter>
__device__ __noinline__ void sweep_misc(G& grp, const SolveLaunch& L, const Cmd& c, double (&red)[7], Iter iter) {
    const IsoEval ev = launch_ev(L);
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
            iter([&](int j) {
                const double gp = grad(zp, j), gn = grad(zn, j);
                const double s = lazy ? -gp : sb[j];
                const double dxj = alpha * s, dgj = gn - gp;
                dx[j] = dxj;
                dg[j] = dgj;
                a = fma(dxj, dgj, a);
                b = fma(dgj, dgj, b);
            });
            __threadfence();
            double r2[2] = {a, b};
            grp.template allreduce<2, 0u>(r2);
            red[0] = r2[0];
            red[1] = r2[1];
            break;
        }
        case kOpGrad: {
            const double* z = c.zcur;
            iter([&](int j) { sb[j] = grad(z, j); });
            __threadfence();
            grp.sync_exec();
            break;
        }
        case kOpDot: {
            double a = 0;
            const double* v = c.v1;
            iter([&](int j) { a = fma(v[j], sb[j], a); });
            double r1[1] = {a};
            grp.template allreduce<1, 0u>(r1);
            red[0] = r1[0];
            break;
        }
        case kOpAxpy: {
            const double cf = c.c;
            const double* v = c.v1;
            iter([&](int j) { sb[j] = fma(cf, v[j], sb[j]); });
            __threadfence();
            grp.sync_exec();
            break;
        }
        case kOpScale: {
            const double cf = c.c;
            iter([&](int j) { sb[j] *= cf; });
            __threadfence();
            grp.sync_exec();
            break;
        }
        case kOpNegDotG: {
            double a = 0;
            const double* z = c.zcur;
            iter([&](int j) {
                const double s = -sb[j];
                sb[j] = s;
                a = fma(grad(z, j), s, a);
            });
            __threadfence();
            double r1[1] = {a};
            grp.template allreduce<1, 0u>(r1);
            red[0] = r1[0];
            break;
        }
        default: break;
    }
}


// ---- the controller: scalar L-BFGS + Hager–Zhang, one warp per CTA ------------------------
struct Red7 {
    double e, dphi, gg, gmax, s1, s2, xchg;
};

template <class G, class Issuer>
struct Controller {
    G& grp;
    const SolveLaunch& L;
    Issuer& issuer;        // broadcasts `cur` and executes it with the whole group
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
    __device__ Controller(G& g, const SolveLaunch& l, Issuer& is) : grp(g), L(l), issuer(is) {}

    __device__ __forceinline__ void issue(double (&red)[7]) { issuer(cur, red); }

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
        phi = fma(0.5, red[0], launch_ev(L).half_cst);
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
    __device__ __forceinline__ void stamp(int item, int k) {
        if (L.dbg && grp.tid == 0) L.dbg[(size_t)item * 16 + k] = clock64();
    }

    __device__ __noinline__ void solve(int item, int* zstate_row) {
        const IsoEval ev = launch_ev(L);
        double red[7];
        stamp(item, 0);
        com_alpha = NAN;
        last_eval_alpha = NAN;
        last_phi = last_dphi = NAN;

        cur.op = kOpInit;
        cur.lazy = 1;
        cur.commit = 0;
        issue(red);
        stamp(item, 1);
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
            stamp(item, 2);
            const int ls = hager_zhang(1.0, phi_0, dphi_0, alpha, phi_alpha);   // InitialStatic(alpha = 1)
            stamp(item, 3);
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
        stamp(item, 4);
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
        stamp(item, 5);
    }

    // ------------------------------------------------------------------ unit setup
    // Fills `cur` / `zother` with the pointers of launch item `item` (what the reference passes to
    // sample_x_z / ẑ_at_θ for that unit) and returns the unit's zstate cell (or null).
    // xslot: the group's x scratch row (null in the streaming kernel, which never materialises x).
    __device__ __forceinline__ int* setup_unit(int item, const double* zshared, double* xslot) {
        const size_t ld = (size_t)L.ld;
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
        cur.item = item;
        cur.smp = launch_smp(L, tsel);
        cur.start_kind = (draw < 0 && L.start_kind == kStartTruth) ? kStartZero : L.start_kind;
        cur.xi = draw >= 0 ? L.xi + (size_t)draw * ld : nullptr;
        cur.nu = draw >= 0 ? L.nu + (size_t)draw * ld : nullptr;
        cur.xw = draw >= 0 ? xslot : nullptr;
        cur.xsrc = draw >= 0 ? cur.xw : L.xdat;
        double* zA = L.zA ? L.zA + (size_t)row * ld : nullptr;
        double* zB = L.zB ? L.zB + (size_t)row * ld : nullptr;
        cur.zA = zA;
        int* zs = L.zstate ? L.zstate + row : nullptr;
        switch (cur.start_kind) {
            case kStartOwn: {
                const int st = ld_state(zs);
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
        return zs;
    }

    __device__ __forceinline__ const double* resolve_zshared() const {
        if (L.zshared_state) {       // shared start = result of an earlier launch (fiducial ẑ)
            const int st = ld_state(L.zshared_state);
            return st == kZA ? L.zsharedA : (st == kZB ? L.zsharedB : nullptr);
        }
        return L.zshared;
    }

    __device__ __noinline__ void run_items() {
        const int gi = grp.group_index();
        const int gn = grp.group_count();
        const size_t ld = (size_t)L.ld;
        cur.sbuf = L.sbuf + (size_t)gi * ld;
        double* const xslot = L.xslot + (size_t)gi * ld;   // this group's x scratch row (L2-resident)
        dxh = L.dxh + (size_t)gi * L.lbfgs_m * ld;
        dgh = L.dgh + (size_t)gi * L.lbfgs_m * ld;
        cur.v1 = cur.v2 = nullptr;
        cur.w1 = cur.w2 = nullptr;
        cur.c = 0.0;
        const double* zshared = resolve_zshared();
        // items of the launch, or the units the streaming kernel handed back (device-side list)
        const int n = L.item_count ? *L.item_count : L.nitems;
        for (int ii = gi; ii < n; ii += gn) {
            const int item = L.item_list ? L.item_list[ii] : ii;
            int* zs = setup_unit(item, zshared, xslot);
            solve(item, zs);
        }
    }
};
// [host-test:end ctl-b]


}  // namespace
}  // namespace muse
