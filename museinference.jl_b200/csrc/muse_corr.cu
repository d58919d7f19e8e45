// muse_corr.cu — MAP + score for the dense correlated-Gaussian family (F3), sm_100a.
//
//   −logLike(x, z | θ) = ½ [ ‖x − z‖² + a·zᵀPz + d·θ ],  a = e^{−θ},  P = Σ₀⁻¹        
//   sample:  z = e^{θ/2}·L ξ,  x = z + ν   (L = chol Σ₀);  ∇z f = (z − x) + a·P z;  ∇θ logLike = ½ a·zᵀPz − d/2
//
// What it replaces: the same mapped body as the other families (/root/reference/src/muse.jl:170-175, 510-513,
// 430-432; ẑ_at_θ = Optim L-BFGS(m) + HagerZhang, src/interface.jl:162-166), for a model whose gradient is a
// dense contraction.
//
// Design.  The Hessian I + a·P is anisotropic, so a solve takes ~7–15 L-BFGS iterations at atol = 1e-2.  All
// units of a pass advance in lock-step, one iteration per round, and the objective is quadratic, so along a search
// direction s
//       φ(α) = f + α·g·s + ½α²·sᵀ(I + aP)s,      φ′(α) = g·s + α·sᵀ(I + aP)s,      ∇f(z + αs) = g + α(s + a·P s):
// ONE product Q = S·P per round — a (units × d)·(d × d) DGEMM on the FP64 tensor cores (muse_dgemm.cu) — serves
// every Hager–Zhang trial of that iteration in closed form and the gradient at the accepted point.  The scalar
// optimiser is the same Controller code as for the other families (muse_iso_ctl.cuh) fed by a closed-form issuer;
// iteration and evaluation counts are those of the reference algorithm (oracle parity: identical counts, ẑ to
// ~1e-15 relative in the NumPy model of this scheme, tests/test_oracle.py::test_f3_lockstep_model).
// Per round: DGEMM (2·units·d² flop, the bound) + one CTA-per-unit kernel doing the line search, the update of
// z, g, the (dx, dg) history and the two-loop recursion for the next direction (≈ 2·m·d doubles read per unit).
//
// F4, the two-layer hierarchy of the Turing adapter's docstring (/root/reference/src/turing.jl:63-79; oracle/families.py TwoLayer):
//   z ~ N(0, e^{σ/2} I_n), w ~ N(z, I_n), x ~ N(w, I_n), y ~ N(x, I_n);  parameter σ, data (x, y), latent u = (z, w), all stacked to d = 2n.
//   −logLike = ½ [ b‖z‖² + ‖w − z‖² + ‖x − w‖² + ‖y − x‖² ] + nσ/4,  b = e^{−σ/2}
//            = ½ ‖X − u‖² + ½ uᵀ(A − I)u + const,   X = (0, x),  A = [[b + 1, −1], [−1, 2]] ⊗ I_n  (two distinct eigenvalues)
// — the same quadratic form with a = 1 and "P" = A − I, a 2 × 2 block per component: the lock-step solver runs unchanged (L-BFGS with
// a live (dx, dg) history: 2 iterations, 6 evaluations per unit), the product per round is an elementwise kernel instead of the DGEMM
// (pair_apply_kernel), and ∇σ logLike = ¼ b‖ẑ_z‖² − n/4.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "muse_handle.cuh"
#include "muse_iso_ctl.cuh"

using namespace muse;

namespace muse {
namespace {

constexpr int kCT = 256;          // threads per unit
constexpr int kMaxM = 16;

struct WarpCtx0 {
    int tid;
};

struct CorrState {
    double f, gmax;
    double cst;          // F4: ½‖y − x‖² of the unit's data (constant in u, part of f)
    double rho[kMaxM], dxdg[kMaxM], dgdg[kMaxM];
    int iter, pseudo, fg, counter_f, status, active, redo, pad;
};

// one batch of units = rows of the arrays below (row stride ld)
struct CorrBatch {
    int rows;            // units
    int mpad;            // rows padded to the GEMM tile
    double *x, *z, *g, *s, *q;      // mpad × ld
    double *dxh, *dgh;              // m × mpad × ld
    CorrState* st;                  // rows
};

struct CorrLaunch {
    int d, ld, m, max_iters;
    int family, n;                       // MUSE_FAMILY_CORRGAUSS | MUSE_FAMILY_TWOLAYER; n = d/2 (F4: components per layer)
    double binv;                         // F4: b = e^{−σ/2} at θ_eval (a = 1)
    double a, half_cst, atol, dhalf;     // dhalf = d/2
    CorrBatch b;
    int row0, nrows;                     // rows [row0, row0 + nrows) take part in this pass
    // init
    int start_kind;                      // kStartZero / kStartOwn / kStartTruth / kStartShared(Keep)
    int data_row;                        // row holding the data unit (−1: none)
    int mode;                            // 0: row r ↔ draw r − 1 + draw_shift; 1: FD virtual sims, row r ↔ draw r / 2, θ_sim = sig[r % 2]
    int draw_shift;
    double sig[2];
    const double *W, *nu, *xdat, *zshared;
    int* active_count;
    // outputs (indexed by item = row − row0)
    double *g_out, *gnorm_out, *f_out;
    int *iters_out, *fg_out, *status_out;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int i = 1; i < kCT / 32; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int i = 1; i < kCT / 32; ++i) t = fmax(t, red[i]);
    return t;
}

// y = P·v for ONE vector (P symmetric, row-major, ld × ld): the data unit's product.  Keeping that single row out of
// the batched GEMM keeps the GEMM at exactly nsims rows (C5: 64 row tiles = 13.8 waves instead of 65 = 14.05 → 15).
__global__ void __launch_bounds__(256) corr_symv_kernel(const double* __restrict__ P, const double* __restrict__ v,
                                                        double* __restrict__ y, int ld) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= ld) return;
    const double* row = P + (size_t)j * ld;
    double acc = 0.0;
    for (int k = 2 * lane; k < ld; k += 64) {
        const double2 p = *reinterpret_cast<const double2*>(row + k);
        const double2 x = *reinterpret_cast<const double2*>(v + k);
        acc = fma(p.x, x.x, acc);
        acc = fma(p.y, x.y, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) y[j] = acc;
}

// F4: Q = V·(A − I), a 2 × 2 block per component — q_z = b·v_z − v_w, q_w = v_w − v_z — for rows [row0, row0 + nrows)
__global__ void __launch_bounds__(256) pair_apply_kernel(const double* __restrict__ V, double* __restrict__ Q, int row0, int n, int ld, double b) {
    const size_t off = (size_t)(row0 + blockIdx.y) * ld;
    const double* v = V + off;
    double* q = Q + off;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
        const double vz = v[j], vw = v[n + j];
        q[j] = fma(b, vz, -vw);
        q[n + j] = vw - vz;
    }
}

// K values at once: slot k is a max if bit k of MAXMASK is set, else a sum (fixed tree: lane butterfly, warps in order)
template <int K, unsigned MAXMASK>
__device__ __forceinline__ void block_reduce(double (&v)[K], double (*red)[4]) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, v[k], off);
            v[k] = ((MAXMASK >> k) & 1u) ? fmax(v[k], o) : v[k] + o;
        }
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[w][k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = red[0][k];
#pragma unroll
        for (int i = 1; i < kCT / 32; ++i) t = ((MAXMASK >> k) & 1u) ? fmax(t, red[i][k]) : t + red[i][k];
        v[k] = t;
    }
}

// ---- pass set-up: x, start vector -------------------------------------------------------------------------
__global__ void __launch_bounds__(kCT) corr_init_kernel(const CorrLaunch L) {
    const int r = L.row0 + blockIdx.x;
    const size_t off = (size_t)r * L.ld;
    double *x = L.b.x + off, *z = L.b.z + off;
    const bool data = (r == L.data_row);
    int draw = 0;
    double sig = L.sig[0];
    if (!data) {
        if (L.mode == 0) draw = r - 1 + L.draw_shift;
        else { draw = r / 2; sig = L.sig[r & 1]; }
    }
    const double* w = L.W + (size_t)draw * L.ld;
    const double* nu = L.nu + (size_t)draw * L.ld;
    const int sk = (data && L.start_kind == kStartTruth) ? kStartZero : L.start_kind;
    if (L.family == MUSE_FAMILY_TWOLAYER) {
        // W rows hold ξ = (ξ_z, ξ_w), ν rows (ν_x, ν_y), xdat = (x, y); the row of the x array is X = (0, x)
        __shared__ double red4[kCT / 32];
        const int n = L.n;
        double c = 0.0;
        for (int j = threadIdx.x; j < L.ld; j += kCT) {
            double xv = 0.0, zt = 0.0;
            if (j < n) {
                if (!data) zt = sig * w[j];                                   // z = e^{σ/4} ξ_z
            } else if (j < L.d) {
                const int i = j - n;
                double xx, yy;
                if (data) { xx = L.xdat[i]; yy = L.xdat[j]; }
                else {
                    zt = sig * w[i] + w[j];                                   // w = z + ξ_w
                    xx = zt + nu[i];                                          // x = w + ν_x
                    yy = xx + nu[j];                                          // y = x + ν_y
                }
                xv = xx;
                const double r3 = yy - xx;
                c = fma(r3, r3, c);
            }
            x[j] = xv;
            if (sk == kStartZero) z[j] = 0.0;
            else if (sk == kStartTruth) z[j] = zt;
            else if (sk == kStartShared || sk == kStartSharedKeep) z[j] = j < L.d ? L.zshared[j] : 0.0;
        }
        c = block_sum(c, red4);
        if (threadIdx.x == 0) L.b.st[r].cst = 0.5 * c;
        return;
    }
    for (int j = threadIdx.x; j < L.ld; j += kCT) {
        const bool in = j < L.d;
        double xv = 0.0, zt = 0.0;
        if (in) {
            if (data) xv = L.xdat[j];
            else { zt = sig * w[j]; xv = zt + nu[j]; }       // z = e^{θ/2} L ξ,  x = z + ν
        }
        x[j] = xv;
        if (sk == kStartZero) z[j] = 0.0;
        else if (sk == kStartTruth) z[j] = zt;
        else if (sk == kStartShared || sk == kStartSharedKeep) z[j] = in ? L.zshared[j] : 0.0;
        // kStartOwn: z keeps the unit's previous ẑ
    }
}

// ---- after Q = Z·P: f, ∇f at the start, convergence at the start, first direction -----------------------------
__global__ void __launch_bounds__(kCT) corr_start_kernel(const CorrLaunch L) {
    __shared__ double red[kCT / 32];
    const int r = L.row0 + blockIdx.x;
    const size_t off = (size_t)r * L.ld;
    const double *x = L.b.x + off, *z = L.b.z + off, *q = L.b.q + off;
    double *g = L.b.g + off, *s = L.b.s + off;
    double rr = 0, zq = 0, gm = 0;
    for (int j = threadIdx.x; j < L.d; j += kCT) {
        const double rj = x[j] - z[j];
        const double gj = fma(L.a, q[j], -rj);
        g[j] = gj;
        s[j] = -gj;
        rr = fma(rj, rj, rr);
        zq = fma(z[j], q[j], zq);
        gm = fmax(gm, fabs(gj));
    }
    rr = block_sum(rr, red);
    zq = block_sum(zq, red);
    gm = block_max(gm, red);
    if (threadIdx.x == 0) {
        CorrState& st = L.b.st[r];
        st.f = fma(0.5, fma(L.a, zq, rr), L.half_cst);
        if (L.family == MUSE_FAMILY_TWOLAYER) st.f += st.cst;
        st.gmax = gm;
        st.iter = 0;
        st.pseudo = 0;
        st.fg = 1;
        st.counter_f = 0;
        st.redo = 0;
        st.status = MUSE_STATUS_G_CONVERGED;
        const bool finite = isfinite(st.f) && isfinite(gm);
        if (!finite) st.status = MUSE_STATUS_NONFINITE;
        st.active = (finite && gm > L.atol && L.max_iters > 0) ? 1 : 0;
        if (st.active) {
            st.iter = 1;            // the first iteration is under way: direction −g, pseudo_iteration 1
            st.pseudo = 1;
            atomicAdd(L.active_count, 1);
        }
    }
}

// closed-form line function for the Controller's Hager–Zhang
struct QuadIssuer {
    double f0, gs, sAs;
    __device__ void operator()(Cmd& cur, double (&red)[7]) {
        const double c = cur.c;
        red[0] = 2.0 * (f0 + c * gs + 0.5 * c * c * sAs);     // phidphi: φ = ½·red[0] + half_cst with half_cst = 0
        red[1] = gs + c * sAs;
        red[2] = red[3] = red[4] = red[5] = red[6] = 0.0;
    }
};

// ---- one lock-step round, after Q = S·P ---------------------------------------------------------------------
__global__ void __launch_bounds__(kCT) corr_iter_kernel(const CorrLaunch L, const SolveLaunch Lq) {
    extern __shared__ double vec[];             // ld doubles: working vector of the two-loop recursion
    __shared__ double red[kCT / 32];
    __shared__ double bc[2];
    __shared__ int flag;
    const int r = L.row0 + blockIdx.x;
    CorrState& st = L.b.st[r];
    if (!st.active) return;
    const size_t off = (size_t)r * L.ld, hstride = (size_t)L.b.mpad * L.ld;
    double *z = L.b.z + off, *g = L.b.g + off, *s = L.b.s + off;
    const double* q = L.b.q + off;
    const int m = L.m;

    __shared__ double red4[kCT / 32][4];
    double r3[3] = {0.0, 0.0, 0.0};           // g·s, s·s, s·Q
    for (int j = threadIdx.x; j < L.d; j += kCT) {
        r3[0] = fma(g[j], s[j], r3[0]);
        r3[1] = fma(s[j], s[j], r3[1]);
        r3[2] = fma(s[j], q[j], r3[2]);
    }
    block_reduce<3, 0u>(r3, red4);
    const double gs = r3[0];
    const double sAs = fma(L.a, r3[2], r3[1]);

    // thread 0: reset test, Hager–Zhang on the closed-form line function
    if (threadIdx.x == 0) {
        flag = 0;
        if (gs >= 0.0 && st.pseudo > 1 && !st.redo) {           // reset_search_direction!: retry this iteration along −g
            st.pseudo = 1;
            st.redo = 1;
            flag = 1;
        } else {
            st.redo = 0;
            WarpCtx0 ctx{0};
            QuadIssuer qi{st.f, gs, sAs};
            Controller<WarpCtx0, QuadIssuer> ctl(ctx, Lq, qi);
            ctl.fg_evals = 0;
            ctl.pre_valid = false;
            ctl.com_alpha = NAN;
            ctl.last_eval_alpha = NAN;
            ctl.last_phi = ctl.last_dphi = NAN;
            ctl.cur.lazy = 0;
            double alpha, phi_alpha;
            const int ls = ctl.hager_zhang(1.0, st.f, gs, alpha, phi_alpha);     // InitialStatic(alpha = 1)
            int fg = ctl.fg_evals;
            if (ls == 0 && alpha != 0.0 && !(ctl.last_eval_alpha == alpha)) fg += 1;   // update_g! at a point not evaluated last
            st.fg += fg;
            bc[0] = alpha;
            if (ls != 0) { st.status = MUSE_STATUS_LS_FAILED; flag = 2; }
        }
    }
    __syncthreads();
    if (flag == 1) {                                   // direction reset: s ← −g, nothing else this round
        for (int j = threadIdx.x; j < L.d; j += kCT) s[j] = -g[j];
        return;
    }
    const double alpha = bc[0];
    const int idx = (st.pseudo - 1) % m;
    double* dx = L.b.dxh + idx * hstride + off;
    double* dg = L.b.dgh + idx * hstride + off;
    double r4[4] = {0.0, 0.0, 0.0, 0.0};      // ‖∇f‖∞, max|Δz|, dx·dg, dg·dg
    for (int j = threadIdx.x; j < L.d; j += kCT) {
        const double dxj = alpha * s[j];
        const double dgj = alpha * fma(L.a, q[j], s[j]);
        const double zn = z[j] + dxj;
        const double gn = g[j] + dgj;
        r4[1] = fmax(r4[1], fabs(zn - z[j]));
        z[j] = zn;
        g[j] = gn;
        dx[j] = dxj;
        dg[j] = dgj;
        r4[0] = fmax(r4[0], fabs(gn));
        r4[2] = fma(dxj, dgj, r4[2]);
        r4[3] = fma(dgj, dgj, r4[3]);
    }
    block_reduce<4, 0x3u>(r4, red4);
    const double gm = r4[0], xc = r4[1], dxdg = r4[2], dgdg = r4[3];
    if (threadIdx.x == 0) {
        const double f_prev = st.f;
        if (flag == 2) {                               // linesearch exception: x moved, optimisation stops
            st.gmax = gm;
            st.active = 0;
        } else {
            st.f = f_prev + alpha * gs + 0.5 * alpha * alpha * sAs;
            st.gmax = gm;
            const bool x_conv = xc <= 0.0, f_conv = fabs(st.f - f_prev) <= 0.0, g_conv = gm <= L.atol;
            st.counter_f = f_conv ? st.counter_f + 1 : 0;
            const bool conv = x_conv || g_conv || st.counter_f > 1;
            if (conv) st.status = g_conv ? MUSE_STATUS_G_CONVERGED : MUSE_STATUS_XF_CONVERGED;
            if (!isfinite(st.f) || !isfinite(gm)) { st.status = MUSE_STATUS_NONFINITE; st.active = 0; }
            else if (conv) st.active = 0;
            else if (st.iter >= L.max_iters) { st.status = MUSE_STATUS_MAXITER; st.active = 0; }
            if (st.active) {                           // update_h!, then the next iteration starts
                const double rho_it = 1.0 / dxdg;
                if (isinf(rho_it)) st.pseudo = 0;
                else { st.rho[idx] = rho_it; st.dxdg[idx] = dxdg; st.dgdg[idx] = dgdg; }
                st.iter += 1;
                st.pseudo += 1;
            }
        }
        if (!st.active) atomicSub(L.active_count, 1);
        flag = st.active;
    }
    __syncthreads();
    if (!flag) return;

    // twoloop!: s ← −H·g for the next round  [EXT Optim.jl l_bfgs.jl]
    const int pseudo = st.pseudo;
    const int lower = pseudo - m, upper = pseudo - 1;
    __shared__ double alpha_tl[kMaxM];
    for (int j = threadIdx.x; j < L.d; j += kCT) vec[j] = g[j];
    __syncthreads();
    for (int index = upper; index >= lower; --index) {
        if (index < 1) continue;
        const int i = (index - 1) % m;
        const double* dxi = L.b.dxh + i * hstride + off;
        const double* dgi = L.b.dgh + i * hstride + off;
        double a = 0;
        for (int j = threadIdx.x; j < L.d; j += kCT) a = fma(dxi[j], vec[j], a);
        a = block_sum(a, red);
        const double al = st.rho[i] * a;
        if (threadIdx.x == 0) alpha_tl[i] = al;
        for (int j = threadIdx.x; j < L.d; j += kCT) vec[j] = fma(-al, dgi[j], vec[j]);
        __syncthreads();
    }
    if (pseudo > 1) {                                  // scaleinvH0
        const int i = (upper - 1) % m;
        const double sc = st.dxdg[i] / st.dgdg[i];
        for (int j = threadIdx.x; j < L.d; j += kCT) vec[j] *= sc;
        __syncthreads();
    }
    for (int index = lower; index <= upper; ++index) {
        if (index < 1) continue;
        const int i = (index - 1) % m;
        const double* dxi = L.b.dxh + i * hstride + off;
        const double* dgi = L.b.dgh + i * hstride + off;
        double b = 0;
        for (int j = threadIdx.x; j < L.d; j += kCT) b = fma(dgi[j], vec[j], b);
        b = block_sum(b, red);
        const double cf = alpha_tl[i] - st.rho[i] * b;
        for (int j = threadIdx.x; j < L.d; j += kCT) vec[j] = fma(cf, dxi[j], vec[j]);
        __syncthreads();
    }
    for (int j = threadIdx.x; j < L.d; j += kCT) s[j] = -vec[j];
}

// ---- score and outputs --------------------------------------------------------------------------
__global__ void __launch_bounds__(kCT) corr_score_kernel(const CorrLaunch L) {
    __shared__ double red[kCT / 32];
    const int r = L.row0 + blockIdx.x;
    const size_t off = (size_t)r * L.ld;
    // a·Pẑ needs no product of its own: the gradient the solver carries is ∇f(ẑ) = (ẑ − x) + a·Pẑ (updated exactly along every
    // accepted step, the objective being quadratic), so a·Pẑ = g − ẑ + x — one DGEMM per pass less than "Q = Ẑ·P, then ẑ·Q"
    const double *z = L.b.z + off, *g = L.b.g + off, *x = L.b.x + off;
    double zq = 0;
    if (L.family == MUSE_FAMILY_TWOLAYER) {
        for (int j = threadIdx.x; j < L.n; j += kCT) zq = fma(z[j], z[j], zq);          // ‖ẑ_z‖²
    } else {
        for (int j = threadIdx.x; j < L.d; j += kCT) zq = fma(z[j], g[j] - z[j] + x[j], zq);
    }
    zq = block_sum(zq, red);
    if (threadIdx.x == 0) {
        const CorrState& st = L.b.st[r];
        const int item = blockIdx.x;
        if (L.family == MUSE_FAMILY_TWOLAYER) L.g_out[item] = 0.25 * L.binv * zq - 0.25 * L.n;      // ∇σ logLike = ¼ e^{−σ/2}‖ẑ_z‖² − n/4
        else L.g_out[item] = 0.5 * zq - L.dhalf;               // ∇θ logLike = ½ e^{−θ} zᵀPz − d/2,  e^{−θ}·Pz = g − z + x
        L.gnorm_out[item] = st.gmax;
        L.f_out[item] = st.f;
        L.iters_out[item] = st.iter;
        L.fg_out[item] = st.fg;
        L.status_out[item] = st.status;
    }
}

// ---- batched conjugate gradients for the implicit-diff branch of get_H! (muse_implicit.cu) --------------------------------------
// Per sim k solve (I + a·P) v = W_k from v = 0, all sims in lock-step: one DGEMM Q = U·P per iteration serves every sim, a CTA per
// sim does the vector updates and the scalars — IterativeSolvers' cg as the reference calls it (src/muse.jl:376-380: maxiter, Pl = I),
// stopped at ‖r‖ ≤ √eps·‖b‖.  (The reference hands cg the Hessian of logLike, −(I + aP), and b = ½σW: the iterates differ by the
// factor −½σ, the iteration count is the same.)
struct CgState {
    double rho, tol;
    int iters, active;
};
struct CgLaunch {
    int d, ld, nrows, maxiter;
    double a;
    const double* W;          // right-hand sides, row k
    double *v, *r, *u, *q;    // solution, residual, direction, Q = U·P
    CgState* st;
    int* active_count;
};
__global__ void __launch_bounds__(kCT) corr_cg_init_kernel(const CgLaunch L) {
    __shared__ double red[kCT / 32];
    const int k = blockIdx.x;
    const size_t off = (size_t)k * L.ld;
    double rr = 0.0;
    for (int j = threadIdx.x; j < L.ld; j += kCT) {
        const double b = j < L.d ? L.W[off + j] : 0.0;
        L.v[off + j] = 0.0;
        L.r[off + j] = b;
        L.u[off + j] = b;                      // first direction: u = r + β·0
        rr = fma(b, b, rr);
    }
    rr = block_sum(rr, red);
    if (threadIdx.x == 0) {
        CgState s;
        s.rho = rr;
        s.tol = 1.4901161193847656e-08 * sqrt(rr);       // √eps(Float64)·‖b‖
        s.iters = 0;
        s.active = sqrt(rr) <= s.tol ? 0 : 1;            // b = 0: cg returns at once
        L.st[k] = s;
        if (s.active) atomicAdd(L.active_count, 1);
    }
}
__global__ void __launch_bounds__(kCT) corr_cg_iter_kernel(const CgLaunch L) {
    __shared__ double red[kCT / 32];
    const int k = blockIdx.x;
    CgState s = L.st[k];
    if (!s.active) return;                     // uniform over the CTA
    const size_t off = (size_t)k * L.ld;
    double uc = 0.0;
    for (int j = threadIdx.x; j < L.d; j += kCT) {
        const double u = L.u[off + j];
        const double c = fma(L.a, L.q[off + j], u);      // c = (I + aP)u
        L.q[off + j] = c;
        uc = fma(u, c, uc);
    }
    uc = block_sum(uc, red);
    const double alpha = s.rho / uc;
    double rr = 0.0;
    for (int j = threadIdx.x; j < L.d; j += kCT) {
        L.v[off + j] = fma(alpha, L.u[off + j], L.v[off + j]);
        const double r = fma(-alpha, L.q[off + j], L.r[off + j]);
        L.r[off + j] = r;
        rr = fma(r, r, rr);
    }
    rr = block_sum(rr, red);
    const int iters = s.iters + 1;
    const bool stop = sqrt(rr) <= s.tol || iters >= L.maxiter;
    if (!stop) {
        const double beta = rr / s.rho;
        for (int j = threadIdx.x; j < L.d; j += kCT) L.u[off + j] = fma(beta, L.u[off + j], L.r[off + j]);
    }
    if (threadIdx.x == 0) {
        s.rho = rr;
        s.iters = iters;
        s.active = stop ? 0 : 1;
        L.st[k] = s;
        if (!stop) atomicAdd(L.active_count, 1);
    }
}
// out[k] = (a·Pẑ)_k · v_k with a·Pẑ = g − ẑ + x of the MAP pass (see corr_score_kernel); rows 1 + k of the main batch
__global__ void __launch_bounds__(kCT) corr_cg_dot_kernel(const double* __restrict__ g, const double* __restrict__ z, const double* __restrict__ x,
                                                          const double* __restrict__ v, int d, int ld, double* __restrict__ out) {
    __shared__ double red[kCT / 32];
    const int k = blockIdx.x;
    const size_t om = (size_t)(1 + k) * ld, ov = (size_t)k * ld;
    double acc = 0.0;
    for (int j = threadIdx.x; j < d; j += kCT) acc = fma(g[om + j] - z[om + j] + x[om + j], v[ov + j], acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[k] = acc;
}

// F4: right-hand sides of the implicit-diff CG, row k = (0, ξ_z of sim k) — and the dot product ẑ_z · v_z with the MAPs in rows 1 + k
__global__ void __launch_bounds__(256) pair_rhs_kernel(const double* __restrict__ W, double* __restrict__ out, int n, int d, int ld) {
    const size_t off = (size_t)blockIdx.y * ld;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < ld; j += gridDim.x * 256) out[off + j] = (j >= n && j < d) ? W[off + j - n] : 0.0;
}
__global__ void __launch_bounds__(kCT) pair_cg_dot_kernel(const double* __restrict__ z, const double* __restrict__ v, int n, int ld,
                                                          double* __restrict__ out) {
    __shared__ double red[kCT / 32];
    const int k = blockIdx.x;
    const size_t om = (size_t)(1 + k) * ld, ov = (size_t)k * ld;
    double acc = 0.0;
    for (int j = threadIdx.x; j < n; j += kCT) acc = fma(z[om + j], v[ov + j], acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[k] = acc;
}

}  // namespace
}  // namespace muse

// =============================================================================== host side
struct muse_corr_ctx {
    int ld = 0;                       // row stride = d rounded up to the GEMM tile (128)
    int draw_rows = 0, draw_pad = 0;  // nsims + 1 (last = master draw), padded to 128
    int h_rows = 0, h_pad = 0;        // separate H shard (multi-GPU)
    double *P = nullptr, *Lt = nullptr;          // ld × ld, zero padded.  The GEMM computes A·Btᵀ (both operands K-contiguous, muse_dgemm.cu):
                                                 // Q = S·P takes Bt = P (symmetric), W = ξ·Lᵀ takes Bt = L — which is what `Lt` holds
    double *W = nullptr, *nu = nullptr, *tmp = nullptr;
    double *W_h = nullptr, *nu_h = nullptr;
    double *xdat = nullptr, *z0user = nullptr;
    double binv = 1.0;                // F4: e^{−σ/2} of the pass under way (pair_apply_kernel)
    bool have_W = false, have_W_h = false;
    CorrBatch main{}, fd{}, fid{};
    int* active = nullptr;
};

namespace {

int round_up_i(int v, int m) { return (v + m - 1) / m * m; }

#define CORR_TRY(h, expr)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
            return e__ == cudaErrorMemoryAllocation ? MUSE_ENOMEM : MUSE_ECUDA;               \
        }                                                                                     \
    } while (0)

void free_batch(CorrBatch& b) {
    cudaFree(b.x); cudaFree(b.z); cudaFree(b.g); cudaFree(b.s); cudaFree(b.q); cudaFree(b.dxh); cudaFree(b.dgh); cudaFree(b.st);
    b = CorrBatch{};
}

int alloc_batch(muse_handle* h, CorrBatch& b, int rows) {
    muse_corr_ctx* c = h->corr;
    if (b.rows >= rows) return MUSE_OK;
    free_batch(b);
    const int mpad = round_up_i(rows, 128) + 128;       // + one tile of slack: GEMM ranges start at any row
    const size_t n = (size_t)mpad * c->ld * sizeof(double), m = (size_t)h->cfg.lbfgs_m;
    CORR_TRY(h, cudaMalloc(&b.x, n));
    CORR_TRY(h, cudaMalloc(&b.z, n));
    CORR_TRY(h, cudaMalloc(&b.g, n));
    CORR_TRY(h, cudaMalloc(&b.s, n));
    CORR_TRY(h, cudaMalloc(&b.q, n));
    CORR_TRY(h, cudaMalloc(&b.dxh, n * m));
    CORR_TRY(h, cudaMalloc(&b.dgh, n * m));
    CORR_TRY(h, cudaMalloc(&b.st, (size_t)mpad * sizeof(CorrState)));
    for (double* p : {b.x, b.z, b.g, b.s, b.q}) CORR_TRY(h, cudaMemsetAsync(p, 0, n, h->stream));
    CORR_TRY(h, cudaMemsetAsync(b.st, 0, (size_t)mpad * sizeof(CorrState), h->stream));
    b.rows = rows;
    b.mpad = mpad;
    return MUSE_OK;
}

// Q[rows] = V[rows]·P for rows [row0, row0 + nrows): a DGEMM over that range rounded up to whole row tiles (the
// rows past the range are scratch: the arrays carry 128 rows of slack), except that a range starting with the data
// unit (row 0 of the main batch) has that one row done by the symmetric matrix-vector kernel.
int gemm_rows(muse_handle* h, const CorrBatch& b, const double* V, int row0, int nrows, bool split_first_row) {
    muse_corr_ctx* c = h->corr;
    if (h->cfg.family == MUSE_FAMILY_TWOLAYER) {
        const int n = h->cfg.d / 2;
        dim3 grid((unsigned)std::min((n + 255) / 256, 64), (unsigned)nrows);
        pair_apply_kernel<<<grid, 256, 0, h->stream>>>(V, b.q, row0, n, c->ld, c->binv);
        CORR_TRY(h, cudaGetLastError());
        h->acc.launches += 1;
        return MUSE_OK;
    }
    if (split_first_row && nrows > 1) {
        corr_symv_kernel<<<(c->ld + 7) / 8, 256, 0, h->stream>>>(c->P, V + (size_t)row0 * c->ld, b.q + (size_t)row0 * c->ld, c->ld);
        CORR_TRY(h, cudaGetLastError());
        h->acc.launches += 1;
        h->acc.solve_flops += 2.0 * (double)c->ld * c->ld;
        row0 += 1;
        nrows -= 1;
    }
    const int m = round_up_i(nrows, 128);
    const size_t off = (size_t)row0 * c->ld;
    CORR_TRY(h, launch_dgemm(V + off, c->P, b.q + off, m, c->ld, c->ld, c->ld, c->ld, c->ld, h->stream));
    h->acc.launches += 1;
    h->acc.solve_flops += 2.0 * m * (double)c->ld * c->ld;
    return MUSE_OK;
}

// the lock-step solve of rows [row0, row0 + nrows) of batch b; outputs go to items 0..nrows−1
int corr_solve(muse_handle* h, CorrBatch& b, CorrLaunch& L) {
    muse_corr_ctx* c = h->corr;
    L.d = h->cfg.d;
    L.ld = c->ld;
    L.family = h->cfg.family;
    L.n = h->cfg.d / 2;
    c->binv = L.binv;
    L.m = h->cfg.lbfgs_m;
    L.max_iters = h->cfg.max_iters;
    L.dhalf = 0.5 * h->cfg.d;
    L.b = b;
    L.active_count = c->active;
    L.g_out = h->g_d; L.gnorm_out = h->gnorm_d; L.f_out = h->f_d;
    L.iters_out = h->iters_d; L.fg_out = h->fg_d; L.status_out = h->status_d;
    SolveLaunch Lq;
    std::memset(&Lq, 0, sizeof(Lq));                  // ev.half_cst = 0: the closed-form issuer returns 2φ
    const int smem = c->ld * (int)sizeof(double);
    if (smem > 48 * 1024) CORR_TRY(h, cudaFuncSetAttribute(corr_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    muse_handle::Rec rec{};
    if (h->prof) {
        CORR_TRY(h, cudaEventCreate(&rec.a));
        CORR_TRY(h, cudaEventCreate(&rec.b));
        CORR_TRY(h, cudaEventRecord(rec.a, h->stream));
    }
    CORR_TRY(h, cudaMemsetAsync(c->active, 0, sizeof(int), h->stream));
    corr_init_kernel<<<L.nrows, kCT, 0, h->stream>>>(L);
    CORR_TRY(h, cudaGetLastError());
    if (L.start_kind == kStartZero) {
        CORR_TRY(h, cudaMemsetAsync(b.q + (size_t)L.row0 * c->ld, 0, (size_t)L.nrows * c->ld * sizeof(double), h->stream));
    } else {
        const int rc = gemm_rows(h, b, b.z, L.row0, L.nrows, L.data_row == L.row0);
        if (rc != MUSE_OK) return rc;
    }
    corr_start_kernel<<<L.nrows, kCT, 0, h->stream>>>(L);
    CORR_TRY(h, cudaGetLastError());
    h->acc.launches += 2;
    const int max_rounds = 2 * h->cfg.max_iters + 8;
    for (int round = 0; round < max_rounds; ++round) {
        int active = 0;
        CORR_TRY(h, cudaMemcpyAsync(&active, c->active, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CORR_TRY(h, cudaStreamSynchronize(h->stream));
        if (active <= 0) break;
        const int rc = gemm_rows(h, b, b.s, L.row0, L.nrows, L.data_row == L.row0);
        if (rc != MUSE_OK) return rc;
        corr_iter_kernel<<<L.nrows, kCT, smem, h->stream>>>(L, Lq);
        CORR_TRY(h, cudaGetLastError());
        h->acc.launches += 1;
    }
    corr_score_kernel<<<L.nrows, kCT, 0, h->stream>>>(L);
    CORR_TRY(h, cudaGetLastError());
    h->acc.launches += 1;
    h->acc.solve_launches += 1;
    if (h->prof) {
        CORR_TRY(h, cudaEventRecord(rec.b, h->stream));
        rec.cls = 0;
        rec.units = L.nrows;
        rec.bytes = 0.0;
        rec.kind = h->pass_kind;
        h->recs.push_back(rec);
    }
    return MUSE_OK;
}

}  // namespace

int muse_corr_create(muse_handle* h) {
    const muse_cfg& cfg = h->cfg;
    const bool pairf = cfg.family == MUSE_FAMILY_TWOLAYER;
    if (!pairf && (!cfg.P || !cfg.L)) { h->err = "corrgauss needs cfg.P = Σ₀⁻¹ and cfg.L = chol(Σ₀) (d × d, row-major)"; return MUSE_EINVAL; }
    if (pairf && (cfg.d & 1)) { h->err = "twolayer: d = 2n (latent (z, w) and data (x, y) stacked) must be even"; return MUSE_EINVAL; }
    muse_corr_ctx* c = new (std::nothrow) muse_corr_ctx();
    if (!c) return MUSE_ENOMEM;
    h->corr = c;
    const int d = cfg.d;
    c->ld = round_up_i(d, 128);
    c->draw_rows = cfg.nsims + 1;
    c->draw_pad = round_up_i(c->draw_rows, 128);
    c->h_rows = cfg.nsims_h;
    c->h_pad = round_up_i(cfg.nsims_h > 0 ? cfg.nsims_h : 1, 128);
    const size_t ld = c->ld, mat = ld * ld * sizeof(double), B = sizeof(double);
    if (!pairf) {
        CORR_TRY(h, cudaMalloc(&c->P, mat));
        CORR_TRY(h, cudaMalloc(&c->Lt, mat));
        CORR_TRY(h, cudaMemsetAsync(c->P, 0, mat, h->stream));
        CORR_TRY(h, cudaMemsetAsync(c->Lt, 0, mat, h->stream));
        CORR_TRY(h, cudaMemcpy2DAsync(c->P, ld * B, cfg.P, (size_t)d * B, (size_t)d * B, d, cudaMemcpyHostToDevice, h->stream));
        CORR_TRY(h, cudaMemcpy2DAsync(c->Lt, ld * B, cfg.L, (size_t)d * B, (size_t)d * B, d, cudaMemcpyHostToDevice, h->stream));
    }
    const size_t dr = (size_t)c->draw_pad * ld * B;
    CORR_TRY(h, cudaMalloc(&c->W, dr));
    CORR_TRY(h, cudaMalloc(&c->nu, dr));
    CORR_TRY(h, cudaMalloc(&c->tmp, (size_t)std::max(c->draw_pad, c->h_pad) * ld * B));
    CORR_TRY(h, cudaMemsetAsync(c->W, 0, dr, h->stream));
    CORR_TRY(h, cudaMemsetAsync(c->nu, 0, dr, h->stream));
    if (cfg.nsims_h > 0) {
        const size_t hr = (size_t)c->h_pad * ld * B;
        CORR_TRY(h, cudaMalloc(&c->W_h, hr));
        CORR_TRY(h, cudaMalloc(&c->nu_h, hr));
        CORR_TRY(h, cudaMemsetAsync(c->W_h, 0, hr, h->stream));
        CORR_TRY(h, cudaMemsetAsync(c->nu_h, 0, hr, h->stream));
    }
    CORR_TRY(h, cudaMalloc(&c->xdat, ld * B));
    CORR_TRY(h, cudaMalloc(&c->z0user, ld * B));
    CORR_TRY(h, cudaMemsetAsync(c->xdat, 0, ld * B, h->stream));
    CORR_TRY(h, cudaMemsetAsync(c->z0user, 0, ld * B, h->stream));
    CORR_TRY(h, cudaMalloc(&c->active, sizeof(int)));
    int rc = alloc_batch(h, c->main, cfg.nsims + 1);
    if (rc != MUSE_OK) return rc;
    rc = alloc_batch(h, c->fid, 1);
    if (rc != MUSE_OK) return rc;
    CORR_TRY(h, cudaStreamSynchronize(h->stream));     // lt goes out of scope
    return MUSE_OK;
}

void muse_corr_destroy(muse_handle* h) {
    muse_corr_ctx* c = h->corr;
    if (!c) return;
    cudaFree(c->P); cudaFree(c->Lt); cudaFree(c->W); cudaFree(c->nu); cudaFree(c->tmp); cudaFree(c->W_h); cudaFree(c->nu_h);
    cudaFree(c->xdat); cudaFree(c->z0user); cudaFree(c->active);
    free_batch(c->main); free_batch(c->fd); free_batch(c->fid);
    delete c;
    h->corr = nullptr;
}

// ξ rows (host or already on the device in c->tmp) → W = ξ·Lᵀ
static int corr_make_W(muse_handle* h, double* Wdst, int pad_rows) {
    muse_corr_ctx* c = h->corr;
    if (h->cfg.family == MUSE_FAMILY_TWOLAYER) {      // no mixing matrix: the rows are the latent normals themselves
        CORR_TRY(h, cudaMemcpyAsync(Wdst, c->tmp, (size_t)pad_rows * c->ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return MUSE_OK;
    }
    CORR_TRY(h, launch_dgemm(c->tmp, c->Lt, Wdst, pad_rows, c->ld, c->ld, c->ld, c->ld, c->ld, h->stream));
    h->acc.launches += 1;
    return MUSE_OK;
}

int muse_corr_set_data(muse_handle* h, const double* x) {
    CORR_TRY(h, cudaMemcpyAsync(h->corr->xdat, x, (size_t)h->cfg.d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CORR_TRY(h, cudaStreamSynchronize(h->stream));
    return MUSE_OK;
}

int muse_corr_set_z0(muse_handle* h, const double* z0) {
    CORR_TRY(h, cudaMemcpyAsync(h->corr->z0user, z0, (size_t)h->cfg.d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CORR_TRY(h, cudaStreamSynchronize(h->stream));
    return MUSE_OK;
}

int muse_corr_set_draws(muse_handle* h, const double* xi, const double* nu, const double* xi_m, const double* nu_m, bool hshard) {
    muse_corr_ctx* c = h->corr;
    const size_t w = (size_t)h->cfg.d * sizeof(double), pitch = (size_t)c->ld * sizeof(double);
    const int n = hshard ? h->cfg.nsims_h : h->cfg.nsims;
    const int pad = hshard ? c->h_pad : c->draw_pad;
    double* nud = hshard ? c->nu_h : c->nu;
    CORR_TRY(h, cudaMemsetAsync(c->tmp, 0, (size_t)pad * pitch, h->stream));
    if (n) {
        CORR_TRY(h, cudaMemcpy2DAsync(c->tmp, pitch, xi, w, w, n, cudaMemcpyHostToDevice, h->stream));
        CORR_TRY(h, cudaMemcpy2DAsync(nud, pitch, nu, w, w, n, cudaMemcpyHostToDevice, h->stream));
    }
    if (!hshard) {
        CORR_TRY(h, cudaMemcpyAsync(c->tmp + (size_t)n * c->ld, xi_m, w, cudaMemcpyHostToDevice, h->stream));
        CORR_TRY(h, cudaMemcpyAsync(nud + (size_t)n * c->ld, nu_m, w, cudaMemcpyHostToDevice, h->stream));
    }
    const int rc = corr_make_W(h, hshard ? c->W_h : c->W, pad);
    if (rc != MUSE_OK) return rc;
    CORR_TRY(h, cudaStreamSynchronize(h->stream));
    (hshard ? c->have_W_h : c->have_W) = true;
    return MUSE_OK;
}

int muse_corr_seed_draws(muse_handle* h, uint64_t seed) {
    muse_corr_ctx* c = h->corr;
    const size_t pitch = (size_t)c->ld * sizeof(double);
    CORR_TRY(h, cudaMemsetAsync(c->tmp, 0, (size_t)c->draw_pad * pitch, h->stream));
    CORR_TRY(h, launch_philox_draws(c->tmp, c->nu, h->cfg.nsims + 1, h->cfg.d, c->ld, seed, h->cfg.sim_offset, h->cfg.nsims, h->stream));
    int rc = corr_make_W(h, c->W, c->draw_pad);
    if (rc != MUSE_OK) return rc;
    c->have_W = true;
    if (h->cfg.nsims_h > 0) {
        CORR_TRY(h, cudaMemsetAsync(c->tmp, 0, (size_t)c->h_pad * pitch, h->stream));
        CORR_TRY(h, launch_philox_draws(c->tmp, c->nu_h, h->cfg.nsims_h, h->cfg.d, c->ld, seed, h->cfg.h_sim_offset, -1, h->stream));
        rc = corr_make_W(h, c->W_h, c->h_pad);
        if (rc != MUSE_OK) return rc;
        c->have_W_h = true;
    }
    h->acc.launches += 1;
    h->acc.draw_launches += 1;
    CORR_TRY(h, cudaStreamSynchronize(h->stream));
    return MUSE_OK;
}

// θ-dependent constants of a pass: evaluation point (a, b, the constant of f) and the scale of the simulated latent
static void corr_eval_consts(const muse_handle* h, double theta_eval, CorrLaunch& L) {
    if (h->cfg.family == MUSE_FAMILY_TWOLAYER) {
        L.a = 1.0;
        L.binv = std::exp(-0.5 * theta_eval);
        L.half_cst = 0.25 * (h->cfg.d / 2) * theta_eval;
    } else {
        L.a = std::exp(-theta_eval);
        L.binv = 1.0;
        L.half_cst = 0.5 * h->cfg.d * theta_eval;
    }
}
static double corr_sim_scale(const muse_handle* h, double theta_sim) {
    return std::exp((h->cfg.family == MUSE_FAMILY_TWOLAYER ? 0.25 : 0.5) * theta_sim);
}

int muse_corr_map_score(muse_handle* h, const double* theta_sim, const double* theta_eval, double atol, int include_data,
                        int warm_start, int first_sim, int count) {
    muse_corr_ctx* c = h->corr;
    if (include_data && first_sim != 0) { h->err = "corrgauss: include_data needs first_sim = 0"; return MUSE_EINVAL; }
    CorrLaunch L{};
    corr_eval_consts(h, theta_eval[0], L);
    L.atol = atol;
    L.sig[0] = L.sig[1] = corr_sim_scale(h, theta_sim[0]);
    L.mode = 0;
    L.data_row = include_data ? 0 : -1;
    L.row0 = include_data ? 0 : 1 + first_sim;
    L.nrows = count + (include_data ? 1 : 0);
    L.W = c->W; L.nu = c->nu; L.xdat = c->xdat;
    switch (warm_start) {
        case MUSE_START_ZEROS: L.start_kind = kStartZero; break;
        case MUSE_START_PREV: L.start_kind = kStartOwn; break;
        case MUSE_START_TRUTH: L.start_kind = kStartTruth; break;
        default: L.start_kind = kStartShared; L.zshared = c->z0user; break;
    }
    return corr_solve(h, c->main, L);
}

// fiducial solve + the 2·n_H virtual sims of get_H! (src/muse.jl:417-442); scores land in items 2k + sgn
int muse_corr_fd_launch(muse_handle* h, const double* theta0, const double* th_pts, int nsims_H, double atol) {
    muse_corr_ctx* c = h->corr;
    const bool hshard = h->cfg.nsims_h > 0;
    CorrLaunch F{};
    corr_eval_consts(h, theta0[0], F);
    F.atol = atol;
    F.sig[0] = F.sig[1] = corr_sim_scale(h, theta0[0]);
    F.mode = 0;
    F.data_row = -1;
    F.row0 = 0;
    F.nrows = 1;
    F.draw_shift = h->cfg.nsims + 1;      // row 0 of the fiducial batch ↔ the master stream's own draw (row nsims)
    F.W = c->W;
    F.nu = c->nu;
    F.start_kind = kStartZero;
    h->pass_kind = MUSE_PASS_FIDUCIAL;
    int rc = corr_solve(h, c->fid, F);
    if (rc != MUSE_OK) return rc;
    rc = alloc_batch(h, c->fd, 2 * nsims_H);
    if (rc != MUSE_OK) return rc;
    CorrLaunch L{};
    L.a = F.a; L.binv = F.binv; L.half_cst = F.half_cst; L.atol = atol;
    L.sig[0] = corr_sim_scale(h, th_pts[0]);      // the "−" and "+" sample points of the single column
    L.sig[1] = corr_sim_scale(h, th_pts[1]);
    L.mode = 1;
    L.data_row = -1;
    L.row0 = 0;
    L.nrows = 2 * nsims_H;
    L.W = hshard ? c->W_h : c->W;
    L.nu = hshard ? c->nu_h : c->nu;
    L.start_kind = kStartShared;
    L.zshared = c->fid.z;
    h->pass_kind = MUSE_PASS_FD;
    return corr_solve(h, c->fd, L);
}

int muse_corr_get_maps(muse_handle* h, int first_unit, int count, double* z_out) {
    muse_corr_ctx* c = h->corr;
    const size_t w = (size_t)h->cfg.d * sizeof(double), pitch = (size_t)c->ld * sizeof(double);
    if (count) CORR_TRY(h, cudaMemcpy2DAsync(z_out, w, c->main.z + (size_t)first_unit * c->ld, pitch, w, count, cudaMemcpyDeviceToHost, h->stream));
    CORR_TRY(h, cudaStreamSynchronize(h->stream));
    return MUSE_OK;
}

bool muse_corr_have_draws(muse_handle* h, bool hshard) { return hshard ? h->corr->have_W_h : h->corr->have_W; }

// implicit-diff branch of get_H! for the dense correlated Gaussian (muse_implicit.cu has the formulas):
//   H_k = ½ a σ (Pẑ_k)·(I + aP)⁻¹ W_k,   ẑ_k the MAP of sim k at θ₀ (atol 1e-1), W_k = L ξ_k
int muse_corr_implicit_h(muse_handle* h, const double* theta0, int nsims_H, int start, int cg_maxiter, double* Hs_out, int32_t* cg_iters_out,
                         int32_t* status_out) {
    muse_corr_ctx* c = h->corr;
    const bool pairf = h->cfg.family == MUSE_FAMILY_TWOLAYER;
    // (1) MAPs of the H sims; the pass leaves ẑ, x and g = ∇f(ẑ), hence a·Pẑ = g − ẑ + x
    h->pass_kind = MUSE_PASS_COLD;
    int rc = muse_corr_map_score(h, theta0, theta0, 1e-1, 0, start, 0, nsims_H);
    if (rc != MUSE_OK) return rc;
    if (status_out) CORR_TRY(h, cudaMemcpyAsync(status_out, h->status_d, (size_t)nsims_H * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    // (2) batched CG on the FD batch's arrays
    rc = alloc_batch(h, c->fd, nsims_H);
    if (rc != MUSE_OK) return rc;
    CgState* st = nullptr;
    double* dots = nullptr;
    CORR_TRY(h, cudaMalloc(&st, (size_t)nsims_H * sizeof(CgState)));
    if (cudaMalloc(&dots, (size_t)nsims_H * sizeof(double)) != cudaSuccess) { cudaFree(st); cudaGetLastError(); h->err = "implicit_h: allocation failed"; return MUSE_ENOMEM; }
    auto done = [&](int code) { cudaFree(st); cudaFree(dots); return code; };
    CgLaunch L{};
    L.d = h->cfg.d; L.ld = c->ld; L.nrows = nsims_H; L.maxiter = cg_maxiter;
    L.a = pairf ? 1.0 : std::exp(-theta0[0]);
    L.W = c->W;
    if (pairf) {
        // F4: ∂σ_sim ∇u logLike = (0, ¼e^{σ/4} ξ_z) (only the w-half sees the data, x = e^{σ/4}ξ_z + ξ_w + ν_x); the scalar goes to the end
        const int n = h->cfg.d / 2;
        dim3 grid((unsigned)std::min((c->ld + 255) / 256, 64), (unsigned)nsims_H);
        pair_rhs_kernel<<<grid, 256, 0, h->stream>>>(c->W, c->tmp, n, h->cfg.d, c->ld);
        h->acc.launches += 1;
        L.W = c->tmp;
        c->binv = std::exp(-0.5 * theta0[0]);
    }
    L.v = c->fd.z; L.r = c->fd.g; L.u = c->fd.s; L.q = c->fd.q;
    L.st = st;
    L.active_count = c->active;
    if (cudaMemsetAsync(c->active, 0, sizeof(int), h->stream) != cudaSuccess) return done(MUSE_ECUDA);
    corr_cg_init_kernel<<<nsims_H, kCT, 0, h->stream>>>(L);
    h->acc.launches += 1;
    for (int it = 0; it <= cg_maxiter; ++it) {
        int active = 0;
        if (cudaMemcpyAsync(&active, c->active, sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "implicit_h: CG poll failed"; return done(MUSE_ECUDA); }
        if (active <= 0) break;
        if (cudaMemsetAsync(c->active, 0, sizeof(int), h->stream) != cudaSuccess) return done(MUSE_ECUDA);
        rc = gemm_rows(h, c->fd, c->fd.s, 0, nsims_H, false);
        if (rc != MUSE_OK) return done(rc);
        corr_cg_iter_kernel<<<nsims_H, kCT, 0, h->stream>>>(L);
        h->acc.launches += 1;
    }
    if (pairf) pair_cg_dot_kernel<<<nsims_H, kCT, 0, h->stream>>>(c->main.z, c->fd.z, h->cfg.d / 2, c->ld, dots);
    else corr_cg_dot_kernel<<<nsims_H, kCT, 0, h->stream>>>(c->main.g, c->main.z, c->main.x, c->fd.z, h->cfg.d, c->ld, dots);
    h->acc.launches += 1;
    std::vector<double> dh((size_t)nsims_H);
    std::vector<CgState> sh((size_t)nsims_H);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(dh.data(), dots, dh.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sh.data(), st, sh.size() * sizeof(CgState), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { h->err = std::string("implicit_h: ") + cudaGetErrorString(e); return done(MUSE_ECUDA); }
    // F3: ½ σ (a·Pẑ)·(I + aP)⁻¹W;   F4: (∂σ ∇u logLike)·A⁻¹·(∂σ_sim ∇u logLike) = ½b ẑ_z · [A⁻¹(0, ¼e^{σ/4} ξ_z)]_z
    const double sig = pairf ? 0.125 * std::exp(-0.5 * theta0[0]) * std::exp(0.25 * theta0[0]) : 0.5 * std::exp(0.5 * theta0[0]);
    for (int k = 0; k < nsims_H; ++k) {
        Hs_out[k] = sig * dh[k];
        if (cg_iters_out) cg_iters_out[k] = sh[k].iters;
    }
    return done(MUSE_OK);
}
