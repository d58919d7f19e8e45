// muse_outer.cu — the outer θ loop of muse! (/root/reference/src/muse.jl:159-236) and its covariance stage (:244-247) with
// the θ update ON THE DEVICE, so that consecutive solver passes follow each other on the stream without a host round trip.
//
// muse_driver.cu runs the same loop with the O(N·nθ) arithmetic on the host: every pass ends in a device→host copy, a stream
// synchronisation and a fresh launch — ≈ 50 µs during which the GPU idles, three times per full solve.  That is 10 % of a C3
// solve (d = 65 536) and most of a C1/C2 solve (d = 512), whose passes take 20–50 µs.  Here one single-CTA kernel per pass
// (theta_step_kernel) does what the host did between two passes — status check, mean and variance of the scores
// (:183, :188), prior terms (:184, :207), H⁻¹_post (:208), the Newton step (:224), the convergence test of the NEXT
// iteration (:163-166) — and writes the θ-dependent constants of the next pass into device memory (DynConsts), which the
// solver kernels read instead of launch parameters.  The host enqueues a chunk of passes (two first — the fewest the
// convergence test needs — then three at a time) speculatively; once the loop has ended the remaining passes see skip = 1
// and return at once.  cov_prep_kernel then derives the finite-difference
// step 0.1 ./ std(gs) (:411-413) and the constants of the fiducial solve and of the 2·nθ sample points θ̂ ± h·eₙ (:417-433), so
// that get_H!'s launches ride the same stream.  One synchronisation per chunk; the typical solve (2 iterations + the
// convergence test + covariance) needs exactly one.
//
// Differences from the host loop: the reductions are parallel trees (deterministic, but not the host's summation order) and
// e^{·} comes from the device's libm, so θ agrees with muse_driver.cu to round-off (≈ 1e-16 relative), not bit for bit.
// The trees are ordered by the GLOBAL sim index, and every rank of a multi-GPU job runs the identical kernel on the
// identical gathered scores: θ stays bit-identical across ranks and for any sharding, as before.  history[i].t is the chunk's wall time divided by its iterations (no per-iteration host clock exists).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "muse_handle.cuh"

using namespace muse;

extern "C" int muse_b200_allgather_rows(muse_handle* h, const double* local_host, int32_t ncol, const int32_t* counts, double* out_host);
int muse_cov_finish(muse_handle* h, const double* theta, const double* gs, int nsims_total, int nsims_h_total, const int32_t* counts_h,
                    const double* Hs_local, int mine, const double* prior_sigma, muse_cov_out* out);

namespace {

constexpr int kStepThreads = 1024;
constexpr int kMaxRanks = 16;

struct OuterParams {
    int nt, family, d;
    int iter;                 // 1-based index of the pass just executed
    int maxsteps;
    int units_local;          // units of the pass on this rank (data + local sims)
    int nranks, n_total;
    int counts[kMaxRanks];    // sims per rank
    long long need;           // doubles per rank slot of g_all
    double alpha, theta_rtol;
    int have_prior;
    double prior_mean[kMaxTheta], prior_sigma[kMaxTheta];
    const double* g_local;    // this pass: local score rows, row 0 = data
    const int* status_local;
    const double* g_all;      // this pass: sim scores of all ranks (rank q's rows at g_all + q·need)
    OuterState* st;
    DynConsts* dyn_next;      // constants of the next pass
};

struct CovParams {
    int nt, family, d;
    int nranks, n_total;
    int counts[kMaxRanks];
    long long need;
    const double* g_all_slot[kOuterSlots];   // gathered sim scores of iteration i live in slot (i − 1) % kOuterSlots
    OuterState* st;
    DynConsts* dyn_fid;
    DynConsts* dyn_fd;
};

// θ → constants, as theta_consts() of muse_api.cu (device libm)
__device__ void consts_of(int family, int d, const double* th_sim, const double* th_eval, IsoSample* smp, IsoEval* ev) {
    const double dd = (double)d;
    if (family == MUSE_FAMILY_FUNNEL) {
        if (smp) { smp->sig = exp(0.5 * th_sim[0]); smp->mu = 0.0; }
        if (ev) { ev->a = exp(-th_eval[0]); ev->mu = 0.0; ev->half_cst = 0.5 * dd * th_eval[0]; ev->cspec = 1.0 / (1.0 + ev->a); }
    } else {
        if (smp) { smp->sig = exp(th_sim[1]); smp->mu = th_sim[0]; }
        if (ev) { ev->a = exp(-2.0 * th_eval[1]); ev->mu = th_eval[0]; ev->half_cst = dd * th_eval[1]; ev->cspec = 1.0 / (1.0 + ev->a); }
    }
}

// Σ_k f_c(k) for every θ-component c over all sims, in an order fixed by the GLOBAL sim index: thread t takes sims t, t + T,
// t + 2T, … (loads issued four at a time — a lone CTA is latency-bound, not bandwidth-bound), then a shuffle tree per warp
// and a second one over the warp results.  The order does not depend on how the sims are sharded over ranks, so θ — like
// every per-sim result — is bit-identical for any number of GPUs.  f(o, c) reads element c of the score row at offset o of
// the gathered layout (rank q's rows start at q·need).  Results land in out[0..nt).
template <class F>
__device__ void block_sums(const int* counts, int nranks, long long need, int nt, int n_total, F&& f, double (*sh)[32], double* out) {
    double acc[kMaxTheta];
#pragma unroll
    for (int c = 0; c < kMaxTheta; ++c) acc[c] = 0.0;
    constexpr int T = kStepThreads;
    auto row_off = [&](int k) -> size_t {                 // global sim k → offset of its row
        int q = 0, base = 0;
        while (q + 1 < nranks && k >= base + counts[q]) { base += counts[q]; ++q; }
        return (size_t)q * (size_t)need + (size_t)(k - base) * nt;
    };
    int k = threadIdx.x;
    for (; k + 3 * T < n_total; k += 4 * T) {
        const size_t o0 = row_off(k), o1 = row_off(k + T), o2 = row_off(k + 2 * T), o3 = row_off(k + 3 * T);
#pragma unroll
        for (int c = 0; c < kMaxTheta; ++c) {
            if (c < nt) {
                const double v0 = f(o0, c), v1 = f(o1, c), v2 = f(o2, c), v3 = f(o3, c);
                acc[c] += v0; acc[c] += v1; acc[c] += v2; acc[c] += v3;
            }
        }
    }
    for (; k < n_total; k += T) {
        const size_t o = row_off(k);
#pragma unroll
        for (int c = 0; c < kMaxTheta; ++c)
            if (c < nt) acc[c] += f(o, c);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < kMaxTheta; ++c) {
        if (c < nt) {
            double v = acc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) sh[c][warp] = v;
        }
    }
    __syncthreads();
    for (int c = 0; c < nt; ++c) {
        double v = (lane < T / 32) ? sh[c][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        out[c] = v;                                      // every warp computes the same value
    }
    __syncthreads();
}

// mean and corrected variance of every component (two passes, like Statistics.mean / var)
__device__ void block_mean_var(const double* g, const int* counts, int nranks, long long need, int nt, int n_total,
                               double (*sh)[32], double* mean, double* var) {
    block_sums(counts, nranks, need, nt, n_total, [&](size_t o, int c) { return g[o + c]; }, sh, mean);
    for (int c = 0; c < nt; ++c) mean[c] /= n_total;
    block_sums(counts, nranks, need, nt, n_total, [&](size_t o, int c) { const double dlt = g[o + c] - mean[c]; return dlt * dlt; }, sh, var);
    for (int c = 0; c < nt; ++c) var[c] /= (n_total - 1);
}

__global__ void __launch_bounds__(kStepThreads) theta_step_kernel(const OuterParams P) {
    __shared__ double sh[kMaxTheta][32];
    __shared__ int bad;
    OuterState* st = P.st;
    if (st->done) {                                            // the pass before this step was skipped: keep skipping
        if (threadIdx.x == 0 && P.dyn_next) P.dyn_next->skip = 1;
        return;
    }
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    // src/interface.jl:170.  The decision must be the same on every rank: with several ranks it is taken from what all of them
    // see — the replicated data unit and the gathered rows, in which a failed sim arrives as NaN (muse_comm.cu / publish_peer)
    if (P.nranks > 1) {
        if (threadIdx.x == 0 && __ldcg(P.status_local) == MUSE_STATUS_NONFINITE) bad = 1;
        for (int q = 0; q < P.nranks; ++q)
            for (int e = threadIdx.x; e < P.counts[q] * P.nt; e += kStepThreads)
                if (isnan(__ldcg(P.g_all + (size_t)q * P.need + e))) bad = 1;
    } else {
        for (int u = threadIdx.x; u < P.units_local; u += kStepThreads)
            if (__ldcg(P.status_local + u) == MUSE_STATUS_NONFINITE) bad = 1;
    }
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0) { st->error = 1; st->done = 1; if (P.dyn_next) P.dyn_next->skip = 1; }
        return;
    }
    const int row = P.iter - 1;
    double mean[kMaxTheta], var[kMaxTheta];
    block_mean_var(P.g_all, P.counts, P.nranks, P.need, P.nt, P.n_total, sh, mean, var);
    if (threadIdx.x != 0) return;
    double th_new[kMaxTheta];
    for (int c = 0; c < P.nt; ++c) {
        const double th = st->theta[c];
        const double g_dat = P.g_local[c];
        const double g_like = g_dat - mean[c];                                                     // :183
        const double g_prior = P.have_prior ? -(th - P.prior_mean[c]) / (P.prior_sigma[c] * P.prior_sigma[c]) : 0.0;   // :184
        const double g_post = g_like + g_prior;                                                    // :185
        const double h_inv_like = -1.0 / var[c];                                                   // :188
        const double h_prior = P.have_prior ? -1.0 / (P.prior_sigma[c] * P.prior_sigma[c]) : 0.0;  // :207
        const double h_inv_post = 1.0 / (1.0 / h_inv_like + h_prior);                              // :208 (diagonal)
        if (row < kOuterMaxIter) {
            OuterRow& R = st->row[row];
            R.theta[c] = th; R.g_dat[c] = g_dat; R.g_like[c] = g_like; R.g_prior[c] = g_prior;
            R.h_inv_like[c] = h_inv_like; R.h_prior[c] = h_prior; R.h_inv_post[c] = h_inv_post;
        }
        th_new[c] = th - P.alpha * (h_inv_post * g_post);                                          // :224
    }
    for (int c = 0; c < P.nt; ++c) st->theta[c] = th_new[c];                                       // :230
    st->n_iter = P.iter;
    int done = 0;
    if (P.iter >= 2) {                                   // the test at the top of iteration iter + 1 > 2   (:163-166)
        double q = 0.0;
        for (int c = 0; c < P.nt; ++c) {
            const double dlt = st->row[row].theta[c] - st->row[row - 1].theta[c];
            q += dlt * st->row[row].h_inv_post[c] * dlt;
        }
        q = -q;
        if (q < 0.0) { st->error = 2; done = 1; }        // DomainError of sqrt in the reference
        else if (sqrt(q) < P.theta_rtol) done = 1;
    }
    if (P.iter >= P.maxsteps) done = 1;
    st->done = done;
    if (P.dyn_next) {
        consts_of(P.family, P.d, th_new, th_new, &P.dyn_next->smp[0], &P.dyn_next->ev);
        P.dyn_next->skip = done;
    }
}

// after the last θ-step of a chunk: if the loop has ended, the constants of get_H!'s launches; otherwise they are skipped
__global__ void __launch_bounds__(kStepThreads) cov_prep_kernel(const CovParams P) {
    __shared__ double sh[kMaxTheta][32];
    OuterState* st = P.st;
    if (!st->done || st->error || st->n_iter < 1) {
        if (threadIdx.x == 0) { P.dyn_fid->skip = 1; P.dyn_fd->skip = 1; }
        return;
    }
    const double* gall = P.g_all_slot[(st->n_iter - 1) % kOuterSlots];
    double step[kMaxTheta], mean[kMaxTheta], var[kMaxTheta];
    block_mean_var(gall, P.counts, P.nranks, P.need, P.nt, P.n_total, sh, mean, var);
    for (int c = 0; c < P.nt; ++c) step[c] = 0.1 / sqrt(var[c]);       // step = 0.1 ./ std(gs)   (:411-413), gs = the last scores (:231)
    if (threadIdx.x != 0) return;
    double th0[kMaxTheta];
    for (int c = 0; c < P.nt; ++c) { th0[c] = st->theta[c]; st->step[c] = step[c]; }
    consts_of(P.family, P.d, th0, th0, &P.dyn_fid->smp[0], &P.dyn_fid->ev);
    P.dyn_fid->skip = 0;
    consts_of(P.family, P.d, th0, th0, nullptr, &P.dyn_fd->ev);
    for (int n = 0; n < P.nt; ++n)
        for (int s = 0; s < 2; ++s) {
            double th[kMaxTheta];
            for (int c = 0; c < P.nt; ++c) th[c] = th0[c];
            const double eps = 0.0 + step[n] * (s ? 1.0 : -1.0);     // x .+ step .* grid   (src/util.jl:15)
            th[n] = th0[n] + eps;
            consts_of(P.family, P.d, th, th0, &P.dyn_fd->smp[2 * n + s], nullptr);
        }
    P.dyn_fd->skip = 0;
}

#define OUTER_TRY(h, expr)                                                                    \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
            return e__ == cudaErrorMemoryAllocation ? MUSE_ENOMEM : MUSE_ECUDA;               \
        }                                                                                     \
    } while (0)

void outer_graph_release(muse_handle* h) {
    if (h->outer_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->outer_exec);
    h->outer_exec = nullptr;
    for (auto& r : h->outer_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    h->outer_recs.clear();
    h->outer_key.clear();
}

size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

// the arena [state | slot 0 | slot 1 | FD block | slot 2] (device + pinned mirror) and the exchange mirrors
int outer_ensure(muse_handle* h, int units, int fd_items, size_t gall_doubles) {
    if (units > h->outer_units_cap || fd_items > h->outer_fd_cap || !h->outer_arena_d) {
        outer_graph_release(h);
        cudaFree(h->outer_arena_d); cudaFreeHost(h->outer_arena_h);
        h->outer_arena_d = h->outer_arena_h = nullptr;
        const int ucap = std::max(units, h->outer_units_cap), fcap = std::max(std::max(fd_items, 1), h->outer_fd_cap);
        const size_t sz_state = up256(sizeof(OuterState)), sz_slot = up256(muse_outblock_bytes(h, ucap)), sz_fd = up256(muse_outblock_bytes(h, fcap));
        const size_t total = sz_state + 3 * sz_slot + sz_fd;
        OUTER_TRY(h, cudaMalloc(&h->outer_arena_d, total));
        OUTER_TRY(h, cudaMallocHost(&h->outer_arena_h, total));
        OUTER_TRY(h, cudaMemsetAsync(h->outer_arena_d, 0, total, h->stream));
        std::memset(h->outer_arena_h, 0, total);
        size_t off = 0;
        h->outer_st_d = reinterpret_cast<OuterState*>(h->outer_arena_d);
        h->outer_st_h = reinterpret_cast<OuterState*>(h->outer_arena_h);
        off += sz_state;
        muse_outblock_carve(h, h->outer_slot[0], h->outer_arena_d + off, h->outer_arena_h + off, ucap); off += sz_slot;
        muse_outblock_carve(h, h->outer_slot[1], h->outer_arena_d + off, h->outer_arena_h + off, ucap); off += sz_slot;
        muse_outblock_carve(h, h->outer_fd, h->outer_arena_d + off, h->outer_arena_h + off, fcap); off += sz_fd;
        h->outer_arena_head = off;
        muse_outblock_carve(h, h->outer_slot[2], h->outer_arena_d + off, h->outer_arena_h + off, ucap); off += sz_slot;
        h->outer_arena_bytes = off;
        h->outer_units_cap = ucap;
        h->outer_fd_cap = fcap;
    }
    if (!h->outer_dyn) {
        OUTER_TRY(h, cudaMalloc(&h->outer_dyn, 4 * sizeof(DynConsts)));
        OUTER_TRY(h, cudaMallocHost(&h->outer_st_stage, sizeof(OuterState)));
    }
    if (gall_doubles > h->outer_gall_doubles) {
        outer_graph_release(h);
        for (int s = 0; s < kOuterSlots; ++s) { cudaFree(h->outer_gall[s]); h->outer_gall[s] = nullptr; }
        cudaFreeHost(h->outer_gall_h);
        h->outer_gall_h = nullptr;
        h->outer_gall_doubles = 0;
        for (int s = 0; s < kOuterSlots; ++s) OUTER_TRY(h, cudaMalloc(&h->outer_gall[s], gall_doubles * sizeof(double)));
        OUTER_TRY(h, cudaMallocHost(&h->outer_gall_h, kOuterSlots * gall_doubles * sizeof(double)));
        h->outer_gall_doubles = gall_doubles;
    }
    return MUSE_OK;
}

}  // namespace

void muse_outer_release(muse_handle* h) {
    outer_graph_release(h);
    cudaFree(h->outer_arena_d);
    cudaFreeHost(h->outer_arena_h);
    cudaFreeHost(h->outer_st_stage);
    cudaFree(h->outer_dyn);
    for (int s = 0; s < kOuterSlots; ++s) { cudaFree(h->outer_gall[s]); h->outer_gall[s] = nullptr; h->outer_slot[s] = OutBlock{}; }
    cudaFreeHost(h->outer_gall_h);
    h->outer_fd = OutBlock{};
    h->outer_arena_d = h->outer_arena_h = nullptr;
    h->outer_st_d = h->outer_st_h = h->outer_st_stage = nullptr;
    h->outer_dyn = nullptr;
    h->outer_gall_h = nullptr;
    h->outer_gall_doubles = 0;
    h->outer_units_cap = h->outer_fd_cap = 0;
}

extern "C" int muse_b200_muse_solve(muse_handle* h, const double* theta0, int32_t nsims_total, const int32_t* counts,
                                    int32_t maxsteps, double theta_rtol, double atol, double alpha, int32_t first_start,
                                    const double* prior_mean, const double* prior_sigma, int32_t get_covariance,
                                    int32_t nsims_h_total, const int32_t* counts_h, muse_iterate_out* out, muse_cov_out* cov) {
    if (!h || !theta0 || !out || maxsteps < 1 || nsims_total < 2) return MUSE_EINVAL;
    if (h->corr) { h->err = "muse_solve: the device-resident loop serves the isotropic families; corrgauss uses muse_iterate"; return MUSE_EUNSUPPORTED; }
    if (maxsteps > kOuterMaxIter) { h->err = "muse_solve: maxsteps exceeds the device history (64 rows); use muse_iterate"; return MUSE_EUNSUPPORTED; }
    if (first_start != MUSE_START_ZEROS && first_start != MUSE_START_USER) { h->err = "first_start must be ZEROS or USER"; return MUSE_EINVAL; }
    if ((prior_mean == nullptr) != (prior_sigma == nullptr)) return MUSE_EINVAL;
    if (get_covariance && (!cov || nsims_h_total < 1)) return MUSE_EINVAL;
    const int nt = h->cfg.ntheta, nloc = h->cfg.nsims, units = nloc + 1;
    const bool multi = h->comm != nullptr && h->comm_nranks > 1;
    if (multi && (!counts || (get_covariance && !counts_h))) { h->err = "counts (and counts_h) required with a communicator"; return MUSE_EINVAL; }
    if (multi && h->comm_nranks > kMaxRanks) { h->err = "muse_solve: more ranks than the θ-step kernel's table"; return MUSE_EUNSUPPORTED; }
    if (!multi && nsims_total != nloc) { h->err = "nsims_total must equal the handle's nsims without a communicator"; return MUSE_EINVAL; }
    if (!h->have_data) { h->err = "observed data not set (muse_b200_set_data)"; return MUSE_ESTATE; }
    if (!h->have_draws) { h->err = "no draws installed (set_draws / seed_draws)"; return MUSE_ESTATE; }
    if (first_start == MUSE_START_USER && !h->have_z0) { h->err = "user z0 not set (muse_b200_set_z0)"; return MUSE_ESTATE; }
    OUTER_TRY(h, cudaSetDevice(h->cfg.device));

    // layout of the gathered scores the θ-step reads
    OuterParams P{};
    P.nt = nt; P.family = h->cfg.family; P.d = h->cfg.d;
    P.maxsteps = maxsteps; P.units_local = units; P.n_total = nsims_total;
    P.alpha = alpha; P.theta_rtol = theta_rtol;
    P.have_prior = prior_sigma ? 1 : 0;
    for (int c = 0; c < nt; ++c) { P.prior_mean[c] = prior_mean ? prior_mean[c] : 0.0; P.prior_sigma[c] = prior_sigma ? prior_sigma[c] : 1.0; }
    int maxc = 1;
    if (multi) {
        P.nranks = h->comm_nranks;
        for (int q = 0; q < P.nranks; ++q) { P.counts[q] = counts[q]; maxc = counts[q] > maxc ? counts[q] : maxc; }
        P.need = (long long)maxc * nt;
    } else {
        P.nranks = 1; P.counts[0] = nloc; P.need = (long long)nloc * nt;
    }
    const size_t gall_doubles = multi ? (size_t)P.need * P.nranks : 0;
    const int nh_mine = get_covariance ? (multi ? counts_h[h->comm_rank] : nsims_h_total) : 0;
    int rc = outer_ensure(h, units, std::max(0, nh_mine) * nt * 2, gall_doubles);
    if (rc != MUSE_OK) return rc;
    if (get_covariance) {
        const bool hshard = h->cfg.nsims_h > 0;
        if (nh_mine < 0 || nh_mine > (hshard ? h->cfg.nsims_h : h->cfg.nsims)) { h->err = "nsims_H outside the handle's H shard"; return MUSE_EINVAL; }
        if (hshard && !h->have_draws_h) { h->err = "no H-shard draws installed (set_draws_h / seed_draws)"; return MUSE_ESTATE; }
    }

    OuterState* sd = h->outer_st_d;
    OuterState* sh_ = h->outer_st_h;
    P.st = sd;
    DynConsts* dyn = h->outer_dyn;

    // Everything one chunk puts on the stream: (first chunk) the initial state and the constants of pass 1 from pinned
    // staging, the passes with their exchange and θ-step, the conditional covariance stage, the copies of the results.
    // The same code runs eagerly or under stream capture.
    auto enqueue_chunk = [&](int first, int last) -> int {
        int rc2;
        // ONE upload initialises the state, zeroes the chains' counters and carries the constants of pass 1; later chunks
        // only need fresh counters
        if (first == 1) OUTER_TRY(h, cudaMemcpyAsync(sd, h->outer_st_stage, offsetof(OuterState, row), cudaMemcpyHostToDevice, h->stream));
        else OUTER_TRY(h, cudaMemsetAsync(sd->ctr, 0, sizeof(sd->ctr), h->stream));
        int chain = 0;
        for (int i = first; i <= last; ++i) {
            const int slot = (i - 1) % kOuterSlots;
            const OutBlock& ob = h->outer_slot[slot];
            // pass i: data + local sims, start zeros / user z₀ on the first, previous ẑ afterwards (:169-176)
            h->rec_tag = i;
            h->ctr_override = &sd->ctr[2 * chain++];
            rc2 = muse_pass_enqueue(h, nullptr, nullptr, atol, 1, i == 1 ? first_start : MUSE_START_PREV, 0, nloc, &ob,
                                    i == 1 ? &sd->dyn_first : &dyn[i & 1]);
            h->ctr_override = nullptr;
            h->rec_tag = 0;
            if (rc2 != MUSE_OK) return rc2;
            P.iter = i;
            P.g_local = ob.g_d;
            P.status_local = ob.status_d;
            if (multi) {                    // the one exchange step, on the stream
                size_t need = 0;
                rc2 = muse_comm_allgather_dev_enqueue(h, ob.g_d + nt, ob.status_d + 1, nt, counts, &need);
                if (rc2 != MUSE_OK) return rc2;
                OUTER_TRY(h, cudaMemcpyAsync(h->outer_gall[slot], h->comm_recv, gall_doubles * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
                P.g_all = h->outer_gall[slot];
            } else {
                P.g_all = ob.g_d + nt;
            }
            P.dyn_next = &dyn[(i + 1) & 1];
            theta_step_kernel<<<1, kStepThreads, 0, h->stream>>>(P);
            OUTER_TRY(h, cudaGetLastError());
            (h->capturing ? h->cap_launches : h->acc.launches) += 1;
        }
        if (get_covariance) {
            CovParams C{};
            C.nt = nt; C.family = P.family; C.d = P.d; C.nranks = P.nranks; C.n_total = nsims_total; C.need = P.need;
            for (int q = 0; q < P.nranks; ++q) C.counts[q] = P.counts[q];
            for (int s = 0; s < kOuterSlots; ++s) C.g_all_slot[s] = multi ? h->outer_gall[s] : h->outer_slot[s].g_d + nt;
            C.st = sd; C.dyn_fid = &dyn[2]; C.dyn_fd = &dyn[3];
            cov_prep_kernel<<<1, kStepThreads, 0, h->stream>>>(C);
            OUTER_TRY(h, cudaGetLastError());
            (h->capturing ? h->cap_launches : h->acc.launches) += 1;
            if (nh_mine > 0) {
                h->rec_tag = -1;
                h->ctr_override = &sd->ctr[2 * chain];          // fiducial: this pair, virtual sims: the next one
                h->zfid_override = &sd->ctr[12];
                rc2 = muse_fd_enqueue(h, nullptr, nullptr, nh_mine, atol, &dyn[2], &dyn[3], &h->outer_fd);
                h->ctr_override = nullptr;
                h->zfid_override = nullptr;
                h->rec_tag = 0;
                if (rc2 != MUSE_OK) return rc2;
            }
        }
        // results of the chunk in ONE copy: the arena's head [state | slot 0 | slot 1 | FD block] holds everything a first chunk
        // produces; later chunks (which also use slot 2) copy the whole arena
        const bool head_only = first == 1 && last <= 2;
        const size_t nbytes = head_only ? (get_covariance && nh_mine > 0 ? h->outer_arena_head
                                                                          : (size_t)(h->outer_slot[1].d - h->outer_arena_d) + h->outer_slot[1].bytes)
                                        : h->outer_arena_bytes;
        OUTER_TRY(h, cudaMemcpyAsync(h->outer_arena_h, h->outer_arena_d, nbytes, cudaMemcpyDeviceToHost, h->stream));
        if (multi)
            for (int i = first; i <= last; ++i)
                OUTER_TRY(h, cudaMemcpyAsync(h->outer_gall_h + (size_t)((i - 1) % kOuterSlots) * gall_doubles, h->outer_gall[(i - 1) % kOuterSlots],
                                             gall_doubles * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        return MUSE_OK;
    };

    // what a captured graph bakes in: every by-value launch parameter and every pointer the chunk touches
    std::vector<unsigned char> key;
    {
        auto put = [&](const void* p, size_t n) { const unsigned char* b = static_cast<const unsigned char*>(p); key.insert(key.end(), b, b + n); };
        const int ints[] = {maxsteps, first_start, get_covariance, nh_mine, nsims_total, nloc, h->prof ? 1 : 0, P.have_prior, h->out_cap, h->h_cap};
        const double dbl[] = {theta_rtol, atol, alpha};
        put(ints, sizeof(ints)); put(dbl, sizeof(dbl)); put(P.prior_mean, sizeof(P.prior_mean)); put(P.prior_sigma, sizeof(P.prior_sigma));
        const void* ptrs[] = {h->stream, h->out_d, h->zHA, h->outer_arena_d, h->outer_arena_h, h->outer_st_stage, dyn, h->gpart, h->dbg, h->xi, h->xi_h,
                              h->comm_send, h->comm_recv, h->comm_host, h->outer_gall[0], h->outer_gall[1], h->outer_gall[2], h->outer_gall_h};
        put(ptrs, sizeof(ptrs));
    }
    // MUSE_OUTER_GRAPH=0: eager enqueue only.  MUSE_OUTER_GRAPH_MULTI=1: also capture the chunk when a communicator is bound
    // (NCCL's all-gather inside the capture) — off by default: validated eagerly at N = 2, the captured form is unmeasured.
    static const bool graphs_on = [] { const char* e = std::getenv("MUSE_OUTER_GRAPH"); return !e || std::atoi(e) != 0; }();
    static const bool graphs_multi = [] { const char* e = std::getenv("MUSE_OUTER_GRAPH_MULTI"); return e && std::atoi(e) != 0; }();
    const bool use_graph = graphs_on && (!multi || graphs_multi);

    out->n_iter = 0;
    int it_done = 0;                        // iterations whose history has been copied out
    bool finished = false;
    while (!finished) {
        const auto t0 = std::chrono::steady_clock::now();
        const int first = it_done + 1;
        // the first chunk holds the two passes the convergence test needs before it can stop the loop (the reference's
        // typical solve: two iterations, then `break` at the top of the third, :163-166); later chunks hold three
        const int last = std::min(maxsteps, it_done + (it_done == 0 ? 2 : kOuterSlots));
        bool via_graph = false;
        if (first == 1) {
            OuterState* stg = h->outer_st_stage;
            std::memset(stg, 0, offsetof(OuterState, row));
            for (int c = 0; c < nt; ++c) stg->theta[c] = theta0[c];
            if (muse_theta_consts(h->cfg, theta0, theta0, &stg->dyn_first.smp[0], &stg->dyn_first.ev) != 0) { h->err = "family"; return MUSE_EUNSUPPORTED; }
            if (use_graph && h->outer_exec && key == h->outer_key) {
                via_graph = true;
            } else if (use_graph && key == h->outer_warm_key) {
                // second solve with these parameters: every buffer exists, capture the chunk the eager path would enqueue
                outer_graph_release(h);
                cudaGraph_t graph = nullptr;
                h->capturing = true;
                h->cap_launches = h->cap_solve_launches = 0;
                cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed);
                rc = e == cudaSuccess ? enqueue_chunk(first, last) : MUSE_ECUDA;
                const cudaError_t e2 = e == cudaSuccess ? cudaStreamEndCapture(h->stream, &graph) : e;
                h->capturing = false;
                cudaGraphExec_t exec = nullptr;
                cudaError_t e3 = cudaSuccess;
                if (rc == MUSE_OK && e2 == cudaSuccess && (e3 = cudaGraphInstantiate(&exec, graph, 0)) == cudaSuccess) {
                    h->outer_exec = exec;
                    h->outer_key = key;
                    via_graph = true;
                } else {
                    cudaGetLastError();
                    outer_graph_release(h);             // fall back to the eager path below
                }
                if (std::getenv("MUSE_DEBUG_TIMING")) {
                    size_t nodes = 0;
                    if (graph) cudaGraphGetNodes(graph, nullptr, &nodes);
                    std::fprintf(stderr, "[muse_solve] graph capture: begin=%s enqueue rc=%d end=%s instantiate=%s nodes=%zu -> %s\n",
                                 cudaGetErrorName(e), rc, cudaGetErrorName(e2), cudaGetErrorName(e3), nodes, via_graph ? "graph" : "eager");
                }
                if (graph) cudaGraphDestroy(graph);
            }
        }
        if (via_graph) {
            OUTER_TRY(h, cudaGraphLaunch((cudaGraphExec_t)h->outer_exec, h->stream));
            h->acc.launches += h->cap_launches;
            h->acc.solve_launches += h->cap_solve_launches;
        } else {
            rc = enqueue_chunk(first, last);
            if (rc != MUSE_OK) return rc;
            if (first == 1) h->outer_warm_key = key;
        }
        OUTER_TRY(h, cudaStreamSynchronize(h->stream));
        const double chunk_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        if (sh_->error == 1) { h->err = "muse!: MAP solution failed with a non-finite objective"; return MUSE_ESTATE; }
        const int n_now = sh_->n_iter;
        {   // statistics: passes the device skipped (speculative launches after the loop ended, a covariance stage that was
            // not due yet) returned at once — they are neither solver passes nor algorithmic bytes
            const bool cov_ran = sh_->done != 0 && sh_->error == 0;
            int skipped = std::max(0, last - std::max(n_now, first - 1));
            if (get_covariance && nh_mine > 0 && !cov_ran) skipped += 2;
            h->acc.solve_launches -= skipped;
            if (via_graph) {                 // the graph's own event pairs: fold the passes that ran into the accumulators now
                for (const muse_handle::Rec& r : h->outer_recs) {
                    if (r.tag > n_now || (r.tag == -1 && !cov_ran)) continue;
                    float ms = 0.f;
                    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
                    h->acc.solve_ms += ms; h->acc.solve_units += r.units; h->acc.solve_bytes += r.bytes;
                    const int kd = r.kind >= 0 && r.kind < MUSE_PASS_KINDS ? r.kind : MUSE_PASS_COLD;
                    h->acc_pass.launches[kd] += 1; h->acc_pass.ms[kd] += ms; h->acc_pass.units[kd] += r.units; h->acc_pass.bytes[kd] += r.bytes;
                }
            }
            for (size_t k = 0; !via_graph && k < h->recs.size();) {
                const muse_handle::Rec& r = h->recs[k];
                if (r.tag > n_now || (r.tag == -1 && !cov_ran)) {
                    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
                    h->recs.erase(h->recs.begin() + k);
                } else {
                    if (h->recs[k].tag != 0) h->recs[k].tag = 0;      // settled
                    ++k;
                }
            }
        }
        for (int i = first; i <= n_now && i <= last; ++i) {
            const int row = i - 1, slot = row % kOuterSlots;
            const OutBlock& ob = h->outer_slot[slot];
            const OuterRow& R = sh_->row[row];
            for (int c = 0; c < nt; ++c) {
                out->theta_hist[(size_t)row * nt + c] = R.theta[c];
                out->g_dat_hist[(size_t)row * nt + c] = R.g_dat[c];
                out->g_like_hist[(size_t)row * nt + c] = R.g_like[c];
                out->g_prior_hist[(size_t)row * nt + c] = R.g_prior[c];
                out->h_inv_like_hist[(size_t)row * nt + c] = R.h_inv_like[c];
                out->h_prior_hist[(size_t)row * nt + c] = R.h_prior[c];
                out->h_inv_post_hist[(size_t)row * nt + c] = R.h_inv_post[c];
            }
            double* gs = out->g_sims_hist + (size_t)row * nsims_total * nt;
            if (multi) {
                const double* src = h->outer_gall_h + (size_t)slot * gall_doubles;
                size_t off = 0;
                for (int q = 0; q < P.nranks; ++q) {
                    std::memcpy(gs + off, src + (size_t)q * P.need, (size_t)counts[q] * nt * sizeof(double));
                    off += (size_t)counts[q] * nt;
                }
            } else {
                std::memcpy(gs, ob.g_h + nt, (size_t)nloc * nt * sizeof(double));
            }
            std::memcpy(out->iters_hist + (size_t)row * units, ob.iters_h, (size_t)units * sizeof(int));
            std::memcpy(out->fg_hist + (size_t)row * units, ob.fg_h, (size_t)units * sizeof(int));
            std::memcpy(out->gnorm_hist + (size_t)row * units, ob.gnorm_h, (size_t)units * sizeof(double));
            std::memcpy(out->status_hist + (size_t)row * units, ob.status_h, (size_t)units * sizeof(int));
            out->seconds_hist[row] = chunk_s / std::max(1, n_now - first + 1);
        }
        out->n_iter = n_now;
        for (int c = 0; c < nt; ++c) out->theta_final[c] = sh_->theta[c];
        if (sh_->error == 2) { h->err = "DomainError: sqrt of a negative number in the θ convergence test (src/muse.jl:165)"; return MUSE_ESTATE; }
        it_done = n_now;
        finished = sh_->done != 0 || it_done >= maxsteps;
        if (!finished && n_now < last) { h->err = "muse_solve: internal error (loop stalled)"; return MUSE_ESTATE; }
    }

    if (get_covariance) {
        // the FD scores of this rank's H shard are in the main output block; combine with the device's step (:411-413)
        std::vector<double> Hs_local((size_t)std::max(1, nh_mine) * nt * nt);
        std::vector<int32_t> status((size_t)std::max(1, nh_mine) * nt * 2, 0);
        for (int c = 0; c < nt; ++c) cov->step[c] = sh_->step[c];
        if (nh_mine > 0) {
            muse_fd_combine_host(h, h->outer_fd.g_h, h->outer_fd.status_h, sh_->step, nh_mine, Hs_local.data(), status.data());
            for (size_t i = 0; i < (size_t)nh_mine * nt * 2; ++i)
                if (status[i] == MUSE_STATUS_NONFINITE) { h->err = "get_H!: MAP solution failed with a non-finite objective"; return MUSE_ESTATE; }
        }
        const double* gs_last = out->g_sims_hist + (size_t)(out->n_iter - 1) * nsims_total * nt;
        return muse_cov_finish(h, out->theta_final, gs_last, nsims_total, nsims_h_total, counts_h, Hs_local.data(), nh_mine, prior_sigma, cov);
    }
    return MUSE_OK;
}
