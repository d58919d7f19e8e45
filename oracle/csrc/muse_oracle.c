/* muse_oracle.c — plain-C restatement of the per-simulation body of MUSE.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): this is the checker and the timed CPU
 * baseline (bench.py cpu_baseline / --impl reference), never part of the product.
 * PARITY UNPINNED: the reference is Julia and cannot run here; this file follows the same
 * sources as oracle/*.py and is cross-checked against them in tests/test_oracle_c.py.
 *
 * What it restates, per unit (unit 0 may be the observed data, src/muse.jl:170):
 *     x ← prob.x | sample_x_z(rng_k, θ_sim).x              /root/reference/src/simple.jl:61-65
 *     ẑ ← Optim.optimize(only_fg(-logLike, -∇z), z₀, LBFGS(), Options(g_tol=atol))
 *                                                           /root/reference/src/interface.jl:162-166
 *     g ← ∇θ_logLike(x, ẑ, θ_eval)                          /root/reference/src/simple.jl:92
 * i.e. the body of the mapped blocks at src/muse.jl:169-176 and :508-514.
 * [EXT] Optim.jl 1.x L-BFGS (m=10, InitialStatic α=1, scaleinvH0) and LineSearches.jl
 * HagerZhang (delta=.1 sigma=.9 rho=5 epsilon=1e-6 gamma=.66 linesearchmax=50 psi3=.1) are
 * restated from their published sources as in oracle/lbfgs.py and oracle/hagerzhang.py.
 * Gradients are analytic (the reference differentiates logLike with ForwardDiff/Zygote,
 * src/simple.jl:84-85, which is strictly slower: ⌈d/12⌉ sweeps per ∇z with ForwardDiff).
 *
 * Units are independent; they are distributed dynamically over POSIX threads (libgomp is not in
 * this image).  The reference's default pool is a serial map, src/util.jl:73-76; its parallel path
 * is Distributed.pmap over processes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define FAM_FUNNEL 1
#define FAM_HIER 2
#define FAM_CORR 3   /* dense correlated Gaussian: -logLike = ½[Σ(x-z)² + a zᵀPz] + half_cst, P = Σ₀⁻¹ */
#define FAM_TWO 4    /* the toy hierarchy of /root/reference/src/turing.jl:63-79 (oracle/families.py TwoLayer): latent u = (z, w), data (x, y),
                        d = 2n; -logLike = ½[bΣz² + Σ(w-z)² + Σ(x-w)² + Σ(y-x)²] + nσ/4, b = e^{-σ/2} (held in `a`) */

typedef struct {
    int family, d;
    double a, mu, half_cst; /* -logLike = ½[Σ(x-z)² + aΣ(z-μ)²] + half_cst */
    const double* P;        /* FAM_CORR: Σ₀⁻¹, d × d row-major */
    const double* L;        /* FAM_CORR: chol(Σ₀), lower, d × d row-major */
} model_t;

static void model_at(model_t* m, int family, int d, const double* th) {
    m->family = family;
    m->d = d;
    if (family == FAM_TWO) {
        m->a = exp(-0.5 * th[0]);
        m->mu = 0.0;
        m->half_cst = 0.25 * (d / 2) * th[0];
    } else if (family == FAM_FUNNEL || family == FAM_CORR) {
        m->a = exp(-th[0]);
        m->mu = 0.0;
        m->half_cst = 0.5 * d * th[0];
    } else {
        m->a = exp(-2.0 * th[1]);
        m->mu = th[0];
        m->half_cst = (double)d * th[1];
    }
}

/* g ← P z (dense symmetric mat-vec), returns zᵀPz */
static double apply_P(const model_t* m, const double* z, double* g) {
    const int d = m->d;
    double zPz = 0.0;
    for (int i = 0; i < d; ++i) {
        const double* row = m->P + (size_t)i * d;
        double acc = 0.0;
        for (int j = 0; j < d; ++j) acc += row[j] * z[j];
        g[i] = acc;
        zPz += z[i] * acc;
    }
    return zPz;
}

static double fg(const model_t* m, const double* x, const double* z, double* g) {
    double e = 0.0;
    const double a = m->a, mu = m->mu;
    if (m->family == FAM_CORR) {
        const double zPz = apply_P(m, z, g);
        double rr = 0.0;
        for (int j = 0; j < m->d; ++j) {
            const double r = x[j] - z[j];
            rr += r * r;
            g[j] = a * g[j] - r;
        }
        return 0.5 * (rr + a * zPz) + m->half_cst;
    }
    if (m->family == FAM_TWO) {                       /* x = (x, y), z = (z, w) stacked */
        const int n = m->d / 2;
        for (int j = 0; j < n; ++j) {
            const double zz = z[j], ww = z[n + j], xx = x[j], yy = x[n + j];
            const double r1 = ww - zz, r2 = xx - ww, r3 = yy - xx;
            e += a * zz * zz + r1 * r1 + r2 * r2 + r3 * r3;
            g[j] = a * zz - r1;
            g[n + j] = r1 - r2;
        }
        return 0.5 * e + m->half_cst;
    }
    for (int j = 0; j < m->d; ++j) {
        const double r = x[j] - z[j], w = z[j] - mu;
        e += r * r + a * w * w;
        g[j] = a * w - r;
    }
    return 0.5 * e + m->half_cst;
}

static void score(const model_t* m, const double* z, double* out, double* scratch) {
    double s1 = 0.0, s2 = 0.0;
    if (m->family == FAM_CORR) {                      /* ∇θ logLike = ½ e^{-θ} zᵀPz − d/2 */
        out[0] = 0.5 * m->a * apply_P(m, z, scratch) - 0.5 * m->d;
        return;
    }
    if (m->family == FAM_TWO) {                       /* ∇σ logLike = ¼ e^{-σ/2} Σz² − n/4 */
        const int n = m->d / 2;
        for (int j = 0; j < n; ++j) s2 += z[j] * z[j];
        out[0] = 0.25 * m->a * s2 - 0.25 * n;
        return;
    }
    for (int j = 0; j < m->d; ++j) {
        const double w = z[j] - m->mu;
        s1 += w;
        s2 += w * w;
    }
    if (m->family == FAM_FUNNEL) out[0] = 0.5 * m->a * s2 - 0.5 * m->d;
    else { out[0] = m->a * s1; out[1] = m->a * s2 - m->d; }
}

/* ---- objective with NLSolversBase-style caching of the last evaluated point ------------- */
typedef struct {
    const model_t* m;
    const double* x;
    double* x_last; /* d */
    double* g;      /* d: gradient at x_last */
    double f;
    int have, f_calls;
} obj_t;

static double value_gradient(obj_t* o, const double* z) {
    const int d = o->m->d;
    if (!o->have || memcmp(z, o->x_last, sizeof(double) * d) != 0) {
        o->f = fg(o->m, o->x, z, o->g);
        memcpy(o->x_last, z, sizeof(double) * d);
        o->have = 1;
        o->f_calls += 1;
    }
    return o->f;
}

/* ---- Hager–Zhang, index-based as upstream ----------------------------------------------- */
#define HZ_CAP 1024
typedef struct {
    obj_t* o;
    const double *x, *s;
    double* xnew;
    double al[HZ_CAP], val[HZ_CAP], slp[HZ_CAP];
    int n;
} ls_t;

static double dot(const double* a, const double* b, int d) {
    double acc = 0.0;
    for (int j = 0; j < d; ++j) acc += a[j] * b[j];
    return acc;
}

static void phidphi(ls_t* L, double a, double* phi, double* dphi) {
    const int d = L->o->m->d;
    for (int j = 0; j < d; ++j) L->xnew[j] = L->x[j] + a * L->s[j];
    *phi = value_gradient(L->o, L->xnew);
    *dphi = dot(L->o->g, L->s, d);
}
static int push(ls_t* L, double a, double v, double s) {
    if (L->n >= HZ_CAP) return -1;
    L->al[L->n] = a; L->val[L->n] = v; L->slp[L->n] = s;
    return L->n++;
}
static double eps_of(double x) { const double ax = fabs(x); return nextafter(ax, INFINITY) - ax; }

static int wolfe(double c, double phi_c, double dphi_c, double phi_0, double dphi_0, double phi_lim) {
    const double delta = 0.1, sigma = 0.9;
    const int w1 = (delta * dphi_0 >= (phi_c - phi_0) / c) && (dphi_c >= sigma * dphi_0);
    const int w2 = ((2 * delta - 1) * dphi_0 >= dphi_c) && (dphi_c >= sigma * dphi_0) && (phi_c <= phi_lim);
    return w1 || w2;
}
static void bisect(ls_t* L, int* ia, int* ib, double phi_lim) {
    double a = L->al[*ia], b = L->al[*ib];
    while (b - a > eps_of(b)) {
        const double dd = (a + b) / 2.0;
        double p, g;
        phidphi(L, dd, &p, &g);
        const int id = push(L, dd, p, g);
        if (g >= 0.0) { *ib = id; return; }
        if (p <= phi_lim) { a = dd; *ia = id; } else { b = dd; *ib = id; }
    }
}
static void update(ls_t* L, int ia, int ib, int ic, double phi_lim, int* oa, int* ob) {
    const double a = L->al[ia], b = L->al[ib], c = L->al[ic];
    *oa = ia; *ob = ib;
    if (c < a || c > b) return;
    if (L->slp[ic] >= 0.0) { *ob = ic; return; }
    if (L->val[ic] <= phi_lim) { *oa = ic; return; }
    *ob = ic;
    bisect(L, oa, ob, phi_lim);
}
static double secant(double a, double b, double da, double db) { return (a * db - b * da) / (db - da); }

static int secant2(ls_t* L, int ia, int ib, double phi_lim, int* iA, int* iB) {
    const double phi_0 = L->val[0], dphi_0 = L->slp[0];
    double c = secant(L->al[ia], L->al[ib], L->slp[ia], L->slp[ib]), p, g;
    phidphi(L, c, &p, &g);
    int ic = push(L, c, p, g);
    if (wolfe(c, p, g, phi_0, dphi_0, phi_lim)) { *iA = *iB = ic; return 1; }
    update(L, ia, ib, ic, phi_lim, iA, iB);
    const double a = L->al[*iA], b = L->al[*iB];
    if (*iB == ic) c = secant(L->al[ib], L->al[*iB], L->slp[ib], L->slp[*iB]);
    else if (*iA == ic) c = secant(L->al[ia], L->al[*iA], L->slp[ia], L->slp[*iA]);
    if ((*iA == ic || *iB == ic) && a <= c && c <= b) {
        phidphi(L, c, &p, &g);
        ic = push(L, c, p, g);
        if (wolfe(c, p, g, phi_0, dphi_0, phi_lim)) { *iA = *iB = ic; return 1; }
        int na, nb;
        update(L, *iA, *iB, ic, phi_lim, &na, &nb);
        *iA = na; *iB = nb;
    }
    return 0;
}

/* returns 0 ok, 1 LineSearchException; *alpha is the step (or ex.alpha) */
static int hager_zhang(ls_t* L, double c, double phi_0, double dphi_0, double* alpha) {
    const double rho = 5.0, epsilon = 1e-6, gamma = 0.66, psi3 = 0.1, EPS = 2.220446049250313e-16;
    const int linesearchmax = 50, iterfinitemax = 53;
    *alpha = 0.0;
    if (!(isfinite(phi_0) && isfinite(dphi_0))) return 1;
    if (dphi_0 >= EPS * fabs(phi_0)) return 1;
    else if (dphi_0 >= 0.0) return 0;
    L->n = 0;
    push(L, 0.0, phi_0, dphi_0);
    const double phi_lim = phi_0 + epsilon * fabs(phi_0);
    if (c <= EPS) return 0;
    double phi_c, dphi_c;
    phidphi(L, c, &phi_c, &dphi_c);
    int iterfinite = 1;
    while (!(isfinite(phi_c) && isfinite(dphi_c)) && iterfinite < iterfinitemax) {
        iterfinite++; c *= psi3; phidphi(L, c, &phi_c, &dphi_c);
    }
    if (!(isfinite(phi_c) && isfinite(dphi_c))) return 0;
    push(L, c, phi_c, dphi_c);
    int isbracketed = 0, ia = 0, ib = 1, iter = 1;
    while (!isbracketed && iter < linesearchmax) {
        if (dphi_c >= 0.0) {
            ib = L->n - 1;
            for (int i = ib - 1; i >= 0; --i) if (L->val[i] <= phi_lim) { ia = i; break; }
            isbracketed = 1;
        } else if (L->val[L->n - 1] > phi_lim) {
            ib = L->n - 1; ia = 0;
            bisect(L, &ia, &ib, phi_lim);
            isbracketed = 1;
        } else {
            const double cold = c, phi_cold = phi_c;
            (void)phi_cold;
            c *= rho;
            phidphi(L, c, &phi_c, &dphi_c);
            iterfinite = 1;
            while (!(isfinite(phi_c) && isfinite(dphi_c)) && c > nextafter(cold, INFINITY) && iterfinite < iterfinitemax) {
                iterfinite++; c = (cold + c) / 2.0; phidphi(L, c, &phi_c, &dphi_c);
            }
            if (!(isfinite(phi_c) && isfinite(dphi_c))) { *alpha = cold; return 0; }
            push(L, c, phi_c, dphi_c);
        }
        iter++;
    }
    while (iter < linesearchmax) {
        const double a = L->al[ia], b = L->al[ib];
        if (b - a <= eps_of(b)) { *alpha = a; return 0; }
        int iA, iB;
        if (secant2(L, ia, ib, phi_lim, &iA, &iB)) { *alpha = L->al[iA]; return 0; }
        const double A = L->al[iA], B = L->al[iB];
        if (B - A < gamma * (b - a)) {
            if (nextafter(L->val[ia], INFINITY) >= L->val[ib] && nextafter(L->val[iA], INFINITY) >= L->val[iB]) {
                *alpha = A; return 0;
            }
            ia = iA; ib = iB;
        } else {
            const double cc = (A + B) / 2.0;
            double p, g;
            phidphi(L, cc, &p, &g);
            const int ic = push(L, cc, p, g);
            update(L, iA, iB, ic, phi_lim, &ia, &ib);
        }
        iter++;
    }
    *alpha = L->al[ia];
    return 1;
}

/* ---- L-BFGS ------------------------------------------------------------------------------ */
typedef struct {
    int d, m;
    double *x, *xprev, *g, *gprev, *s, *dx, *dg, *q, *xnew, *xlast, *gobj;
    double *dxh, *dgh; /* m × d */
    ls_t* ls;
} work_t;

static work_t* work_new(int d, int m) {
    work_t* w = (work_t*)calloc(1, sizeof(work_t));
    w->d = d; w->m = m;
    double** v[] = {&w->x, &w->xprev, &w->g, &w->gprev, &w->s, &w->dx, &w->dg, &w->q, &w->xnew, &w->xlast, &w->gobj};
    for (unsigned i = 0; i < sizeof(v) / sizeof(v[0]); ++i) *v[i] = (double*)malloc(sizeof(double) * d);
    w->dxh = (double*)malloc(sizeof(double) * d * m);
    w->dgh = (double*)malloc(sizeof(double) * d * m);
    w->ls = (ls_t*)malloc(sizeof(ls_t));
    return w;
}
static void work_free(work_t* w) {
    free(w->x); free(w->xprev); free(w->g); free(w->gprev); free(w->s); free(w->dx); free(w->dg); free(w->q);
    free(w->xnew); free(w->xlast); free(w->gobj); free(w->dxh); free(w->dgh); free(w->ls); free(w);
}

static void twoloop(work_t* w, const double* rho, int pseudo, double* alpha) {
    const int d = w->d, m = w->m, lower = pseudo - m, upper = pseudo - 1;
    memcpy(w->q, w->g, sizeof(double) * d);
    for (int index = upper; index >= lower; --index) {
        if (index < 1) continue;
        const int i = (index - 1) % m;
        const double *dxi = w->dxh + (size_t)i * d, *dgi = w->dgh + (size_t)i * d;
        alpha[i] = rho[i] * dot(dxi, w->q, d);
        for (int j = 0; j < d; ++j) w->q[j] -= alpha[i] * dgi[j];
    }
    if (pseudo > 1) {
        const int i = (upper - 1) % m;
        const double *dxi = w->dxh + (size_t)i * d, *dgi = w->dgh + (size_t)i * d;
        const double scaling = dot(dxi, dgi, d) / dot(dgi, dgi, d);
        for (int j = 0; j < d; ++j) w->s[j] = scaling * w->q[j];
    } else {
        memcpy(w->s, w->q, sizeof(double) * d);
    }
    for (int index = lower; index <= upper; ++index) {
        if (index < 1) continue;
        const int i = (index - 1) % m;
        const double *dxi = w->dxh + (size_t)i * d, *dgi = w->dgh + (size_t)i * d;
        const double beta = rho[i] * dot(dgi, w->s, d);
        for (int j = 0; j < d; ++j) w->s[j] += dxi[j] * (alpha[i] - beta);
    }
    for (int j = 0; j < d; ++j) w->s[j] = -w->s[j];
}

static double maxabs(const double* v, int d) {
    double mx = 0.0;
    for (int j = 0; j < d; ++j) { const double a = fabs(v[j]); if (a > mx) mx = a; }
    return mx;
}

/* minimise from w->x (in/out); returns iterations; *f_calls, *gres, *status out */
static int lbfgs(work_t* w, const model_t* mdl, const double* xdata, double g_tol, int max_iters, int* f_calls,
                 double* gres, int* status) {
    const int d = w->d, m = w->m;
    obj_t o = {mdl, xdata, w->xlast, w->gobj, 0.0, 0, 0};
    double rho[64], alpha_tl[64];
    double f_x = value_gradient(&o, w->x);
    memcpy(w->g, o.g, sizeof(double) * d);
    int stopped = !isfinite(f_x);
    for (int j = 0; j < d && !stopped; ++j) if (!isfinite(w->g[j])) stopped = 1;
    int converged = maxabs(w->g, d) <= g_tol;
    int iteration = 0, pseudo = 0, counter_f_tol = 0;
    *status = stopped ? 4 : 0;
    while (!converged && !stopped && iteration < max_iters) {
        iteration++;
        pseudo++;
        twoloop(w, rho, pseudo, alpha_tl);
        memcpy(w->gprev, w->g, sizeof(double) * d);
        double dphi_0 = dot(w->g, w->s, d);
        if (dphi_0 >= 0.0) {
            pseudo = 1;
            for (int j = 0; j < d; ++j) w->s[j] = -w->g[j];
            dphi_0 = dot(w->g, w->s, d);
        }
        const double phi_0 = f_x, f_prev = f_x;
        memcpy(w->xprev, w->x, sizeof(double) * d);
        ls_t* L = w->ls;
        L->o = &o; L->x = w->xprev; L->s = w->s; L->xnew = w->xnew; L->n = 0;
        double alpha;
        const int lsfail = hager_zhang(L, 1.0, phi_0, dphi_0, &alpha);
        for (int j = 0; j < d; ++j) { w->dx[j] = alpha * w->s[j]; w->x[j] = w->x[j] + w->dx[j]; }
        if (lsfail) { *status = 3; break; }
        f_x = value_gradient(&o, w->x);
        memcpy(w->g, o.g, sizeof(double) * d);
        double xch = 0.0;
        for (int j = 0; j < d; ++j) { const double a = fabs(w->x[j] - w->xprev[j]); if (a > xch) xch = a; }
        const int x_conv = xch <= 0.0, f_conv = fabs(f_x - f_prev) <= 0.0, g_conv = maxabs(w->g, d) <= g_tol;
        counter_f_tol = f_conv ? counter_f_tol + 1 : 0;
        converged = x_conv || g_conv || (counter_f_tol > 1);
        if (converged) *status = g_conv ? 0 : 1;
        for (int j = 0; j < d; ++j) w->dg[j] = w->g[j] - w->gprev[j];
        const double dxdg = dot(w->dx, w->dg, d);
        const double rho_it = 1.0 / dxdg;
        if (isinf(rho_it)) pseudo = 0;
        else {
            const int idx = (pseudo - 1) % m;
            memcpy(w->dxh + (size_t)idx * d, w->dx, sizeof(double) * d);
            memcpy(w->dgh + (size_t)idx * d, w->dg, sizeof(double) * d);
            rho[idx] = rho_it;
        }
        if (!isfinite(f_x)) { *status = 4; break; }
    }
    if (!converged && !stopped && *status == 0 && iteration >= max_iters) *status = 2;
    *f_calls = o.f_calls;
    *gres = maxabs(w->g, d);
    return iteration;
}

/* ---- batch entry point -------------------------------------------------------------------
 * start_mode: 0 zeros, 1 z_inout holds the start (previous ẑ), 2 truth (simulated z).
 * z_inout: units × d (may be NULL for modes 0/2 if MAPs are not wanted). */
typedef struct {
    int family, d, nsims, ntheta, units, include_data, start_mode;
    const double *xi, *nu, *xdat;
    double sig, smu, atol;
    model_t mdl;
    double *z_inout, *g_out, *gnorm_out;
    int *iters_out, *fg_out, *status_out;
    atomic_int next;
} job_t;

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    const int d = J->d;
    work_t* w = work_new(d, 10);
    double* x = (double*)malloc(sizeof(double) * d);
    for (;;) {
        const int u = atomic_fetch_add(&J->next, 1);
        if (u >= J->units) break;
        const int is_data = J->include_data && u == 0;
        const int k = u - (J->include_data ? 1 : 0);
        const double* xs;
        if (is_data) {
            xs = J->xdat;
            if (J->start_mode == 1) memcpy(w->x, J->z_inout + (size_t)u * d, sizeof(double) * d);
            else memset(w->x, 0, sizeof(double) * d);
        } else {
            const double *xk = J->xi + (size_t)k * d, *nk = J->nu + (size_t)k * d;
            if (J->family == FAM_TWO) {               /* z = e^{σ/4} ξ_z, w = z + ξ_w, x = w + ν_x, y = x + ν_y */
                const int n = d / 2;
                for (int j = 0; j < n; ++j) {
                    const double zt = J->sig * xk[j], wt = zt + xk[n + j];
                    x[j] = wt + nk[j];
                    x[n + j] = x[j] + nk[n + j];
                    if (J->start_mode == 2) { w->x[j] = zt; w->x[n + j] = wt; }
                }
            } else
            for (int j = 0; j < d; ++j) {
                double base = xk[j];
                if (J->family == FAM_CORR) {          /* z = e^{θ/2} L ξ */
                    const double* row = J->mdl.L + (size_t)j * d;
                    base = 0.0;
                    for (int c = 0; c <= j; ++c) base += row[c] * xk[c];
                }
                const double zt = J->smu + J->sig * base;
                x[j] = zt + nk[j];
                if (J->start_mode == 2) w->x[j] = zt;
            }
            xs = x;
            if (J->start_mode == 1) memcpy(w->x, J->z_inout + (size_t)u * d, sizeof(double) * d);
            else if (J->start_mode == 0) memset(w->x, 0, sizeof(double) * d);
        }
        int fc = 0, st = 0;
        double gres = 0.0;
        const int it = lbfgs(w, &J->mdl, xs, J->atol, 1000, &fc, &gres, &st);
        score(&J->mdl, w->x, J->g_out + (size_t)u * J->ntheta, x);
        if (J->z_inout) memcpy(J->z_inout + (size_t)u * d, w->x, sizeof(double) * d);
        if (J->iters_out) J->iters_out[u] = it;
        if (J->fg_out) J->fg_out[u] = fc;
        if (J->gnorm_out) J->gnorm_out[u] = gres;
        if (J->status_out) J->status_out[u] = st;
    }
    free(x);
    work_free(w);
    return NULL;
}

int muse_oracle_max_threads(void) {
    const long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int muse_oracle_map_score_consts(int family, int d, int nsims, const double* xi, const double* nu, const double* xdat,
                                 const double* theta_sim, const double* theta_eval, double atol, int include_data,
                                 int start_mode, double* z_inout, double* g_out, int* iters_out, int* fg_out,
                                 double* gnorm_out, int* status_out, int nthreads, const double* P, const double* L);

int muse_oracle_map_score(int family, int d, int nsims, const double* xi, const double* nu, const double* xdat,
                          const double* theta_sim, const double* theta_eval, double atol, int include_data,
                          int start_mode, double* z_inout, double* g_out, int* iters_out, int* fg_out,
                          double* gnorm_out, int* status_out, int nthreads) {
    return muse_oracle_map_score_consts(family, d, nsims, xi, nu, xdat, theta_sim, theta_eval, atol, include_data, start_mode,
                                        z_inout, g_out, iters_out, fg_out, gnorm_out, status_out, nthreads, NULL, NULL);
}

/* same, with the constants of the correlated-Gaussian family (P = Σ₀⁻¹, L = chol Σ₀; NULL for the other families) */
int muse_oracle_map_score_consts(int family, int d, int nsims, const double* xi, const double* nu, const double* xdat,
                                 const double* theta_sim, const double* theta_eval, double atol, int include_data,
                                 int start_mode, double* z_inout, double* g_out, int* iters_out, int* fg_out,
                                 double* gnorm_out, int* status_out, int nthreads, const double* P, const double* L) {
    if (family != FAM_FUNNEL && family != FAM_HIER && family != FAM_CORR && family != FAM_TWO) return -5;
    if (family == FAM_TWO && (d & 1)) return -1;
    if (family == FAM_CORR && (!P || !L)) return -1;
    job_t J;
    memset(&J, 0, sizeof(J));
    J.family = family; J.d = d; J.nsims = nsims;
    J.ntheta = family == FAM_HIER ? 2 : 1;
    J.include_data = include_data ? 1 : 0;
    J.units = nsims + J.include_data;
    J.start_mode = start_mode;
    J.xi = xi; J.nu = nu; J.xdat = xdat; J.atol = atol;
    model_at(&J.mdl, family, d, theta_eval);
    J.mdl.P = P;
    J.mdl.L = L;
    J.sig = family == FAM_HIER ? exp(theta_sim[1]) : exp((family == FAM_TWO ? 0.25 : 0.5) * theta_sim[0]);
    J.smu = family == FAM_HIER ? theta_sim[0] : 0.0;
    J.z_inout = z_inout; J.g_out = g_out; J.gnorm_out = gnorm_out;
    J.iters_out = iters_out; J.fg_out = fg_out; J.status_out = status_out;
    atomic_init(&J.next, 0);
    int nt = nthreads > 0 ? nthreads : muse_oracle_max_threads();
    if (nt > J.units) nt = J.units > 0 ? J.units : 1;
    if (nt <= 1) { worker(&J); return 0; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
    for (int i = 0; i < nt; ++i) pthread_create(&th[i], NULL, worker, &J);
    for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
    free(th);
    return 0;
}
