/* solve_funnel.c — the C ABI of libmuse_b200.so used from plain C, no Python, no torch: a full MUSE solve (θ̂, J, H, Σ) of
 * Neal's funnel (/root/reference/src/simple.jl:58-76; docs/src/index.md:154-176) through muse_b200_muse_solve.
 *
 *   gcc -std=c99 -O2 -I include examples/solve_funnel.c -o solve_funnel -L museinference.jl_b200 -lmuse_b200 -lm \
 *       -Wl,-rpath,$PWD/museinference.jl_b200
 *   ./solve_funnel [d] [nsims]
 *
 * Without a CUDA device the program reports MUSE_ENODEVICE and exits with status 3: the library has no CPU fallback.
 * With one it checks the result against the closed forms of SURVEY.md §8(c): J = H = d·s²/2 with s = 1/(1+e^{−θ}).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "muse_b200.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double uniform01(void) {          /* splitmix64 → (0,1) */
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
}
static double normal(void) { return sqrt(-2.0 * log(uniform01())) * cos(6.283185307179586 * uniform01()); }

int main(int argc, char** argv) {
    const int d = argc > 1 ? atoi(argv[1]) : 4096, nsims = argc > 2 ? atoi(argv[2]) : 512, maxsteps = 50;
    muse_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = MUSE_B200_ABI_VERSION;
    cfg.family = MUSE_FAMILY_FUNNEL;
    cfg.d = d;
    cfg.ntheta = 1;
    cfg.nsims = nsims;
    muse_handle* h = NULL;
    int rc = muse_b200_create(&cfg, &h);
    if (rc != MUSE_OK) {
        fprintf(stderr, "muse_b200_create: %d (%s)\n", rc, muse_b200_last_error(NULL));
        return rc == MUSE_ENODEVICE ? 3 : 1;
    }
    /* observed data at θ_true = 0: z ~ N(0, I), x ~ N(z, I) */
    double* x = malloc(sizeof(double) * d);
    for (int j = 0; j < d; ++j) x[j] = normal() + normal();
    rc = muse_b200_set_data(h, x);
    if (rc == MUSE_OK) rc = muse_b200_seed_draws(h, 20261017ull);          /* split_rng: src/util.jl:85-92 */
    if (rc != MUSE_OK) { fprintf(stderr, "setup: %s\n", muse_b200_last_error(h)); return 1; }

    /* muse(prob, θ₀ = 1; nsims, get_covariance = true) with prior N(0, 3) (src/simple.jl:69-71) */
    const int units = nsims + 1, nh = nsims / 10 > 0 ? nsims / 10 : 1;
    muse_iterate_out it;
    memset(&it, 0, sizeof it);
    double theta_final[1];
    it.theta_final = theta_final;
    it.theta_hist = calloc(maxsteps, sizeof(double));
    it.g_dat_hist = calloc(maxsteps, sizeof(double));
    it.g_sims_hist = calloc((size_t)maxsteps * nsims, sizeof(double));
    it.g_like_hist = calloc(maxsteps, sizeof(double));
    it.g_prior_hist = calloc(maxsteps, sizeof(double));
    it.h_inv_like_hist = calloc(maxsteps, sizeof(double));
    it.h_prior_hist = calloc(maxsteps, sizeof(double));
    it.h_inv_post_hist = calloc(maxsteps, sizeof(double));
    it.seconds_hist = calloc(maxsteps, sizeof(double));
    it.iters_hist = calloc((size_t)maxsteps * units, sizeof(int32_t));
    it.fg_hist = calloc((size_t)maxsteps * units, sizeof(int32_t));
    it.gnorm_hist = calloc((size_t)maxsteps * units, sizeof(double));
    it.status_hist = calloc((size_t)maxsteps * units, sizeof(int32_t));
    double J[1], step[1], H[1], Sinv[1], S[1];
    muse_cov_out cov;
    cov.J = J; cov.step = step; cov.H = H; cov.Sigma_inv = Sinv; cov.Sigma = S;
    cov.Hs = calloc(nh, sizeof(double));
    const double theta0[1] = {1.0}, prior_mean[1] = {0.0}, prior_sigma[1] = {3.0};
    rc = muse_b200_muse_solve(h, theta0, nsims, NULL, maxsteps, 1e-1, 1e-2, 0.7, MUSE_START_ZEROS, prior_mean, prior_sigma,
                              1, nh, NULL, &it, &cov);
    if (rc != MUSE_OK) { fprintf(stderr, "muse_b200_muse_solve: %d (%s)\n", rc, muse_b200_last_error(h)); return 1; }
    const double th = theta_final[0], s = 1.0 / (1.0 + exp(-it.theta_hist[it.n_iter - 1])), sh = 1.0 / (1.0 + exp(-th));
    printf("iterations %d  theta_hat %.6f  sigma %.6f  J %.3f (closed form %.3f)  H %.3f (closed form %.3f)\n", it.n_iter, th,
           sqrt(S[0]), J[0], 0.5 * d * s * s, H[0], 0.5 * d * sh * sh);
    /* J is a sample variance over nsims scores, H a mean over nsims/10 Jacobians: allow for their sampling error */
    const int ok = fabs(J[0] / (0.5 * d * s * s) - 1.0) < 6.0 * sqrt(2.0 / nsims) && fabs(H[0] / (0.5 * d * sh * sh) - 1.0) < 0.2 &&
                   fabs(th) < 6.0 * sqrt(S[0]) + 0.5;
    muse_b200_destroy(h);
    printf(ok ? "OK\n" : "MISMATCH\n");
    return ok ? 0 : 2;
}
