// muse_driver.cu — the outer θ loop of muse! (/root/reference/src/muse.jl:159-236) run inside the library for the
// common configuration (include/muse_b200.h: muse_b200_muse_iterate).  Pure host code: O(N·nθ) arithmetic on the
// gathered scores between two solver passes; every pass is muse_b200_map_score_async + the exchange step +
// muse_b200_fetch.  The general configuration (callable α, regularize, Broyden updates, save_MAPs, resume) stays in
// the host-language driver (museinference.jl_b200/muse.py), which mirrors the reference line by line.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "muse_handle.cuh"

extern "C" int muse_b200_allgather_scores(muse_handle* h, int32_t first_row, const int32_t* counts, double* out_host);
extern "C" int muse_b200_allgather_rows(muse_handle* h, const double* local_host, int32_t ncol, const int32_t* counts, double* out_host);

namespace {

const bool g_debug_timing = std::getenv("MUSE_DEBUG_TIMING") != nullptr;

// Σ f(k), k < n, in double with 8 interleaved partial sums (vectorisable; error growth like pairwise summation).
// The first version accumulated in long double: x87 arithmetic made mean/var of 16 384 scores cost ≈ 0.1 ms per
// pass at 8 ranks — on the critical path between two solver passes of every rank.
template <class F>
inline double sum8(int n, F&& f) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int k = 0;
    for (; k + 8 <= n; k += 8)
        for (int u = 0; u < 8; ++u) acc[u] += f(k + u);
    for (; k < n; ++k) acc[k & 7] += f(k);
    return ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
}

// mean and corrected variance of column c of an n × nt row-major matrix (two-pass, like Statistics.mean / var)
void mean_var(const double* g, int n, int nt, int c, double* mean, double* var) {
    const double m = sum8(n, [&](int k) { return g[(size_t)k * nt + c]; }) / n;
    const double q = sum8(n, [&](int k) { const double d = g[(size_t)k * nt + c] - m; return d * d; });
    *mean = m;
    *var = q / (n - 1);
}

// in-place Gauss–Jordan inverse with partial pivoting of an n × n row-major matrix (n ≤ MUSE_MAX_NTHETA)
bool invert_small(double* a, int n) {
    double inv[MUSE_MAX_NTHETA * MUSE_MAX_NTHETA];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) inv[i * n + j] = i == j ? 1.0 : 0.0;
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(a[r * n + c]) > std::fabs(a[p * n + c])) p = r;
        if (a[p * n + c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < n; ++j) { std::swap(a[p * n + j], a[c * n + j]); std::swap(inv[p * n + j], inv[c * n + j]); }
        const double piv = a[c * n + c];
        for (int j = 0; j < n; ++j) { a[c * n + j] /= piv; inv[c * n + j] /= piv; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) { a[r * n + j] -= f * a[c * n + j]; inv[r * n + j] -= f * inv[c * n + j]; }
        }
    }
    std::memcpy(a, inv, sizeof(double) * n * n);
    return true;
}

}  // namespace

// J from the scores, exchange of the per-sim Jacobians, H and Σ — the host arithmetic of the covariance stage, shared by
// muse_b200_muse_covariance (below) and the device-resident loop (muse_outer.cu).  out->step is left to the caller.
int muse_cov_finish(muse_handle* h, const double* theta, const double* gs, int nsims_total, int nsims_h_total, const int32_t* counts_h,
                    const double* Hs_local, int mine, const double* prior_sigma, muse_cov_out* out, const double* Hs_all) {
    (void)theta;
    const int nt = h->cfg.ntheta;
    const bool multi = h->comm != nullptr && h->comm_nranks > 1;
    // J = var(gs) | cov(SimpleCovariance(corrected = true), gs)                       src/muse.jl:529
    std::vector<double> mean(nt);
    for (int c = 0; c < nt; ++c) mean[c] = sum8(nsims_total, [&](int k) { return gs[(size_t)k * nt + c]; }) / nsims_total;
    for (int a = 0; a < nt; ++a)
        for (int b = a; b < nt; ++b) {
            const double q = sum8(nsims_total, [&](int k) { return (gs[(size_t)k * nt + a] - mean[a]) * (gs[(size_t)k * nt + b] - mean[b]); });
            out->J[a * nt + b] = out->J[b * nt + a] = q / (nsims_total - 1);
        }
    if (Hs_all) {            // already gathered (peer exchange of solve_persist_kernel)
        std::memcpy(out->Hs, Hs_all, (size_t)nsims_h_total * nt * nt * sizeof(double));
    } else if (multi) {
        const int rc = muse_b200_allgather_rows(h, Hs_local, nt * nt, counts_h, out->Hs);
        if (rc != MUSE_OK) return rc;
    } else {
        std::memcpy(out->Hs, Hs_local, (size_t)mine * nt * nt * sizeof(double));
    }
    // H = mean(Hs)                                                                    src/muse.jl:446
    for (int e = 0; e < nt * nt; ++e)
        out->H[e] = sum8(nsims_h_total, [&](int k) { return out->Hs[(size_t)k * nt * nt + e]; }) / nsims_h_total;
    // finalize_result!: Σ⁻¹ = H'·inv(J)·H + H_prior, H_prior = −∇²logPrior(θ); Σ = inv(Σ⁻¹)   src/muse.jl:535-541
    double Jinv[MUSE_MAX_NTHETA * MUSE_MAX_NTHETA], tmp[MUSE_MAX_NTHETA * MUSE_MAX_NTHETA];
    std::memcpy(Jinv, out->J, sizeof(double) * nt * nt);
    if (!invert_small(Jinv, nt)) { h->err = "finalize_result!: J is singular"; return MUSE_ESTATE; }
    for (int i = 0; i < nt; ++i)
        for (int j = 0; j < nt; ++j) {
            double s = 0.0;
            for (int k = 0; k < nt; ++k) s += Jinv[i * nt + k] * out->H[k * nt + j];
            tmp[i * nt + j] = s;
        }
    for (int i = 0; i < nt; ++i)
        for (int j = 0; j < nt; ++j) {
            double s = 0.0;
            for (int k = 0; k < nt; ++k) s += out->H[k * nt + i] * tmp[k * nt + j];
            if (i == j && prior_sigma) s += 1.0 / (prior_sigma[i] * prior_sigma[i]);
            out->Sigma_inv[i * nt + j] = s;
        }
    std::memcpy(out->Sigma, out->Sigma_inv, sizeof(double) * nt * nt);
    if (!invert_small(out->Sigma, nt)) { h->err = "finalize_result!: Σ⁻¹ is singular"; return MUSE_ESTATE; }
    return MUSE_OK;
}

extern "C" int muse_b200_muse_covariance(muse_handle* h, const double* theta, const double* gs, int32_t nsims_total,
                                         int32_t nsims_h_total, const int32_t* counts_h, double atol,
                                         const double* prior_sigma, muse_cov_out* out) {
    if (!h || !theta || !gs || !out || nsims_total < 2 || nsims_h_total < 1) return MUSE_EINVAL;
    const int nt = h->cfg.ntheta;
    const bool multi = h->comm != nullptr && h->comm_nranks > 1;
    if (multi && !counts_h) { h->err = "counts_h (H sims per rank) required with a communicator"; return MUSE_EINVAL; }
    // step = 0.1 ./ std(gs)                                                           src/muse.jl:411-413
    for (int c = 0; c < nt; ++c) {
        double m, v;
        mean_var(gs, nsims_total, nt, c, &m, &v);
        out->step[c] = 0.1 / std::sqrt(v);
    }
    // per-sim finite-difference Jacobians of this rank's H shard                      src/muse.jl:417-442
    const int mine = multi ? counts_h[h->comm_rank] : nsims_h_total;
    std::vector<double> local((size_t)(mine > 0 ? mine : 1) * nt * nt);
    std::vector<int32_t> status((size_t)(mine > 0 ? mine : 1) * nt * 2, 0);
    int rc = muse_b200_fd_jacobian(h, theta, out->step, mine, atol, local.data(), status.data());
    if (rc != MUSE_OK) return rc;
    for (size_t i = 0; i < (size_t)mine * nt * 2; ++i)
        if (status[i] == MUSE_STATUS_NONFINITE) { h->err = "get_H!: MAP solution failed with a non-finite objective"; return MUSE_ESTATE; }
    return muse_cov_finish(h, theta, gs, nsims_total, nsims_h_total, counts_h, local.data(), mine, prior_sigma, out, nullptr);
}

extern "C" int muse_b200_muse_iterate(muse_handle* h, const double* theta0, int32_t nsims_total, const int32_t* counts,
                                      int32_t maxsteps, double theta_rtol, double atol, double alpha, int32_t first_start,
                                      const double* prior_mean, const double* prior_sigma, muse_iterate_out* out) {
    if (!h || !theta0 || !out || maxsteps < 0 || nsims_total < 2) return MUSE_EINVAL;
    if (first_start != MUSE_START_ZEROS && first_start != MUSE_START_USER) { h->err = "first_start must be ZEROS or USER"; return MUSE_EINVAL; }
    const int nt = h->cfg.ntheta, nloc = h->cfg.nsims, units = nloc + 1;
    const bool multi = h->comm != nullptr && h->comm_nranks > 1;
    if (multi && !counts) { h->err = "counts (sims per rank) required with a communicator"; return MUSE_EINVAL; }
    if (!multi && nsims_total != nloc) { h->err = "nsims_total must equal the handle's nsims without a communicator"; return MUSE_EINVAL; }
    if ((prior_mean == nullptr) != (prior_sigma == nullptr)) return MUSE_EINVAL;

    std::vector<double> theta(theta0, theta0 + nt), hpost_prev(nt, 0.0);
    std::vector<double> g_local((size_t)units * nt);
    out->n_iter = 0;
    bool have_prev = false, have_prev2 = false;
    for (int i = 1; i <= maxsteps; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        if (i > 2 && have_prev2) {                                           // src/muse.jl:163-166
            const double* th_a = out->theta_hist + (size_t)(i - 2) * nt;     // history[end].θ
            const double* th_b = out->theta_hist + (size_t)(i - 3) * nt;     // history[end-1].θ
            double q = 0.0;
            for (int c = 0; c < nt; ++c) { const double d = th_a[c] - th_b[c]; q += d * hpost_prev[c] * d; }
            q = -q;
            if (q < 0.0) { h->err = "DomainError: sqrt of a negative number in the θ convergence test (src/muse.jl:165)"; return MUSE_ESTATE; }
            if (std::sqrt(q) < theta_rtol) break;
        }
        const int row = i - 1;
        const auto t_a = std::chrono::steady_clock::now();
        int rc = muse_b200_map_score_async(h, theta.data(), theta.data(), atol, 1, i == 1 ? first_start : MUSE_START_PREV, 0, nloc);
        if (rc != MUSE_OK) return rc;
        double* gs = out->g_sims_hist + (size_t)row * nsims_total * nt;
        if (multi) {
            rc = muse_comm_allgather_scores_enqueue(h, 1, counts);           // the one exchange step (NCCL on the stream)
            if (rc != MUSE_OK) return rc;
        }
        rc = muse_b200_fetch(h, units, g_local.data(), out->iters_hist + (size_t)row * units, out->fg_hist + (size_t)row * units,
                             out->gnorm_hist + (size_t)row * units, out->status_hist + (size_t)row * units);
        if (rc != MUSE_OK) return rc;
        // the error decision must be the same on every rank (a rank that returned alone would leave the others waiting in the
        // next exchange): the data unit is replicated, and the gathered rows of failed sims arrive as NaN (muse_comm.cu)
        bool failed = out->status_hist[(size_t)row * units] == MUSE_STATUS_NONFINITE;
        if (!multi) {
            for (int u = 1; u < units; ++u) failed = failed || out->status_hist[(size_t)row * units + u] == MUSE_STATUS_NONFINITE;
        }
        if (g_debug_timing) {
            const auto t_b = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[muse_iterate rank %d] iter %d: enqueue+wait %.1f us (since loop top %.1f us)\n", h->comm_rank, i,
                         std::chrono::duration<double, std::micro>(t_b - t_a).count(),
                         std::chrono::duration<double, std::micro>(t_b - t0).count());
        }
        if (multi) {
            muse_comm_unpack(h, nt, counts, gs);                              // fetch() has synchronised the stream
            for (size_t e = 0; e < (size_t)nsims_total * nt; ++e) failed = failed || std::isnan(gs[e]);
        } else {
            std::memcpy(gs, g_local.data() + nt, (size_t)nloc * nt * sizeof(double));
        }
        if (failed) { h->err = "muse!: MAP solution failed with a non-finite objective"; return MUSE_ESTATE; }
        double* th_row = out->theta_hist + (size_t)row * nt;
        for (int c = 0; c < nt; ++c) {
            double m, v;
            mean_var(gs, nsims_total, nt, c, &m, &v);
            const double g_dat = g_local[c];
            const double g_like = g_dat - m;                                                  // :183
            const double g_prior = prior_sigma ? -(theta[c] - prior_mean[c]) / (prior_sigma[c] * prior_sigma[c]) : 0.0;   // :184
            const double g_post = g_like + g_prior;                                           // :185
            const double h_inv_like = -1.0 / v;                                               // :188
            const double h_prior = prior_sigma ? -1.0 / (prior_sigma[c] * prior_sigma[c]) : 0.0;   // :207
            const double h_inv_post = 1.0 / (1.0 / h_inv_like + h_prior);                     // :208 (diagonal)
            th_row[c] = theta[c];
            out->g_dat_hist[(size_t)row * nt + c] = g_dat;
            out->g_like_hist[(size_t)row * nt + c] = g_like;
            out->g_prior_hist[(size_t)row * nt + c] = g_prior;
            out->h_inv_like_hist[(size_t)row * nt + c] = h_inv_like;
            out->h_prior_hist[(size_t)row * nt + c] = h_prior;
            out->h_inv_post_hist[(size_t)row * nt + c] = h_inv_post;
            hpost_prev[c] = h_inv_post;
            theta[c] = theta[c] - alpha * (h_inv_post * g_post);                              // :224
        }
        have_prev2 = have_prev;
        have_prev = true;
        out->n_iter = i;
        for (int c = 0; c < nt; ++c) out->theta_final[c] = theta[c];                          // :230
        out->seconds_hist[row] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return MUSE_OK;
}
