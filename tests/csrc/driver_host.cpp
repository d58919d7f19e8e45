// driver_host.cpp — host build of museinference.jl_b200/csrc/muse_driver.cu (TEST INFRASTRUCTURE ONLY).
//
// muse_driver.cu is pure host code: the outer θ loop and the covariance stage around the solver passes.  Compiled here with
// g++ against stubs of the entry points it calls (the passes themselves need a GPU), it lets a CPU test check its arithmetic
// — mean / variance, prior terms, Newton step, convergence test, J, the small Gauss–Jordan inverse, Σ — for every nθ up to
// MUSE_MAX_NTHETA against NumPy, although the registered families only exercise nθ ≤ 2 on the GPU.
#include <cstring>
#include <vector>

#include "../../museinference.jl_b200/csrc/muse_driver.cu"

namespace {
// canned outputs of the passes: per iteration (units × nθ) scores, row 0 = data
const double* g_canned = nullptr;
int g_units = 0, g_nt = 0, g_iter = 0;
const double* g_Hs = nullptr;      // canned per-sim Jacobians (nh × nθ × nθ)
}  // namespace

extern "C" {
int muse_b200_map_score_async(muse_handle*, const double*, const double*, double, int32_t, int32_t, int32_t, int32_t) { return MUSE_OK; }
int muse_b200_fetch(muse_handle*, int32_t units, double* g_out, int32_t* iters, int32_t* fg, double* gnorm, int32_t* status) {
    std::memcpy(g_out, g_canned + (size_t)g_iter * g_units * g_nt, (size_t)units * g_nt * sizeof(double));
    for (int u = 0; u < units; ++u) { iters[u] = 1; fg[u] = 3; gnorm[u] = 0.0; status[u] = 0; }
    ++g_iter;
    return MUSE_OK;
}
int muse_b200_fd_jacobian(muse_handle* h, const double*, const double*, int32_t nsims_H, double, double* Hs_out, int32_t* status_out) {
    const int nt = h->cfg.ntheta;
    std::memcpy(Hs_out, g_Hs, (size_t)nsims_H * nt * nt * sizeof(double));
    if (status_out) std::memset(status_out, 0, (size_t)nsims_H * nt * 2 * sizeof(int32_t));
    return MUSE_OK;
}
int muse_comm_allgather_scores_enqueue(muse_handle*, int, const int32_t*) { return MUSE_OK; }
void muse_comm_unpack(muse_handle*, int, const int32_t*, double*) {}
int muse_b200_allgather_rows(muse_handle*, const double*, int32_t, const int32_t*, double*) { return MUSE_OK; }
int muse_b200_allgather_scores(muse_handle*, int32_t, const int32_t*, double*) { return MUSE_OK; }

// run muse_b200_muse_iterate + muse_b200_muse_covariance on canned scores; returns the iterate call's code
int host_driver_run(int nt, int nsims, int maxsteps, double theta_rtol, double alpha, const double* theta0, const double* prior_mean,
                    const double* prior_sigma, const double* canned /* maxsteps × (nsims+1) × nt */, const double* Hs, int nh,
                    int* n_iter, double* theta_final, double* theta_hist, double* h_inv_post_hist, double* g_like_hist,
                    double* J, double* step, double* H, double* Sigma_inv, double* Sigma) {
    muse_handle h;
    h.cfg.ntheta = nt;
    h.cfg.nsims = nsims;
    g_canned = canned; g_units = nsims + 1; g_nt = nt; g_iter = 0; g_Hs = Hs;
    const int units = nsims + 1;
    std::vector<double> g_dat(maxsteps * nt), g_sims((size_t)maxsteps * nsims * nt), g_prior(maxsteps * nt), h_like(maxsteps * nt), h_prior(maxsteps * nt),
        secs(maxsteps), gnorm((size_t)maxsteps * units);
    std::vector<int32_t> iters((size_t)maxsteps * units), fg((size_t)maxsteps * units), status((size_t)maxsteps * units);
    muse_iterate_out o;
    o.n_iter = 0; o.theta_final = theta_final; o.theta_hist = theta_hist; o.g_dat_hist = g_dat.data(); o.g_sims_hist = g_sims.data();
    o.g_like_hist = g_like_hist; o.g_prior_hist = g_prior.data(); o.h_inv_like_hist = h_like.data(); o.h_prior_hist = h_prior.data();
    o.h_inv_post_hist = h_inv_post_hist; o.seconds_hist = secs.data(); o.iters_hist = iters.data(); o.fg_hist = fg.data();
    o.gnorm_hist = gnorm.data(); o.status_hist = status.data();
    int rc = muse_b200_muse_iterate(&h, theta0, nsims, nullptr, maxsteps, theta_rtol, 1e-2, alpha, MUSE_START_ZEROS, prior_mean, prior_sigma, &o);
    *n_iter = o.n_iter;
    if (rc != MUSE_OK) return rc;
    std::vector<double> Hs_out((size_t)nh * nt * nt);
    muse_cov_out c;
    c.J = J; c.step = step; c.Hs = Hs_out.data(); c.H = H; c.Sigma_inv = Sigma_inv; c.Sigma = Sigma;
    const double* gs_last = g_sims.data() + (size_t)(o.n_iter - 1) * nsims * nt;
    return muse_b200_muse_covariance(&h, theta_final, gs_last, nsims, nh, nullptr, 1e-2, prior_sigma, &c);
}
}
