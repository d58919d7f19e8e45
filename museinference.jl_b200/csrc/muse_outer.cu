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
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "muse_outer_dev.cuh"

using namespace muse;

extern "C" int muse_b200_allgather_rows(muse_handle* h, const double* local_host, int32_t ncol, const int32_t* counts, double* out_host);
int muse_cov_finish(muse_handle* h, const double* theta, const double* gs, int nsims_total, int nsims_h_total, const int32_t* counts_h,
                    const double* Hs_local, int mine, const double* prior_sigma, muse_cov_out* out, const double* Hs_all);

namespace {

constexpr int kStepThreads = kStepLanes;

__global__ void __launch_bounds__(kStepThreads) theta_step_kernel(const OuterParams P) {
    __shared__ double sh[kMaxTheta][32];
    __shared__ int bad;
    if (P.st->done) {                                          // the pass before this step was skipped: keep skipping
        if (threadIdx.x == 0 && P.dyn_next) P.dyn_next->skip = 1;
        return;
    }
    theta_step_body<1>(P, sh, &bad);
}

// after the last θ-step of a chunk: if the loop has ended, the constants of get_H!'s launches; otherwise they are skipped
__global__ void __launch_bounds__(kStepThreads) cov_prep_kernel(const CovParams P) {
    __shared__ double sh[kMaxTheta][32];
    cov_prep_body<1>(P, sh);
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

// the one-launch solve returns as soon as its completion word is in host memory; the event pair around the launch is read later —
// at the next solve or when the profile is asked for — when the launch has long retired
void muse_persist_flush(muse_handle* h) {
    if (!h->persist_pend) return;
    h->persist_pend = false;
    float ms = 0.f;
    cudaError_t e = cudaEventElapsedTime(&ms, h->persist_ev[0], h->persist_ev[1]);
    if (e == cudaErrorNotReady) {
        cudaGetLastError();
        if (cudaEventSynchronize(h->persist_ev[1]) == cudaSuccess) e = cudaEventElapsedTime(&ms, h->persist_ev[0], h->persist_ev[1]);
    }
    if (e != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    h->acc.solve_ms += ms; h->acc.solve_units += h->persist_pend_units; h->acc.solve_bytes += h->persist_pend_bytes;
}

void muse_outer_release(muse_handle* h) {
    outer_graph_release(h);
    h->persist_pend = false;
    cudaFreeHost(h->persist_done_h);
    h->persist_done_h = nullptr;
    cudaFree(h->outer_arena_d);
    cudaFreeHost(h->outer_arena_h);
    cudaFreeHost(h->outer_st_stage);
    cudaFree(h->outer_dyn);
    cudaFree(h->persist_ctl);
    h->persist_ctl = nullptr;
    for (auto& e : h->persist_ev) { if (e) cudaEventDestroy(e); e = nullptr; }
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
    const auto t_entry = std::chrono::steady_clock::now();
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
    if (muse_ensure_outputs(h, std::max(units, std::max(0, nh_mine) * nt * 2)) != 0) return MUSE_ECUDA;   // segment sums / hand-back list of every unit
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

    // copy the history rows of iterations [first, upto] out of the host mirrors (state, per-pass output blocks, gathered scores)
    const double* gall_h = h->outer_gall_h;          // host mirror of the gathered score slots (multi-GPU) and its slot stride
    size_t gall_h_stride = gall_doubles;
    bool gall_dense = false;                         // rows in global sim order (peer exchange) instead of one padded slot per rank
    auto copy_history = [&](int first, int upto, double chunk_s) {
        for (int i = first; i <= upto; ++i) {
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
            if (multi && gall_dense) {     // peer-exchange mirror: blocks [slot 0 | slot 1 | FD | slot 2]
                std::memcpy(gs, gall_h + (size_t)(slot < 2 ? slot : slot + 1) * gall_h_stride, (size_t)nsims_total * nt * sizeof(double));
            } else if (multi) {
                const double* src = gall_h + (size_t)slot * gall_h_stride;
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
            out->seconds_hist[row] = chunk_s / std::max(1, upto - first + 1);
        }
    };

    out->n_iter = 0;
    int it_done = 0;                        // iterations whose history has been copied out
    bool finished = false;
    std::vector<double> Hs_all;             // multi-GPU persistent launch: the FD Jacobians of ALL ranks (no second exchange)

    // ---- the first chunk — normally the whole solve — in ONE cooperative launch (solve_persist_kernel, muse_iso_stream.cu) ----
    // MUSE_PERSIST=0 switches it off (A/B against the chain of launches below).
    // (read at every call, so that one process can hold the two paths against each other)
    const bool persist_on = [] { const char* e = std::getenv("MUSE_PERSIST"); return !e || std::atoi(e) != 0; }();
    const bool p2p_on = [] { const char* e = std::getenv("MUSE_EXCHANGE"); return !e || std::strcmp(e, "nccl") != 0; }();
    // lazy ẑ and lean evaluation (muse_common.cuh: LazyLevels, SolveLaunch::lean): on unless MUSE_LAZY=0 / MUSE_LEAN=0.
    // Measured on C3 (one B200, profiles/README.md): the passes are co-limited by the FP64 pipe, so lazy ẑ alone (16·d instead of
    // 24·d / 32·d bytes per sim) gains ≈ 4 % and lean evaluation alone ≈ 5 %; together 1.48 → 1.30 ms per solve, C4 4.56 → 3.94 ms.
    const bool lazy_on = [] { const char* e = std::getenv("MUSE_LAZY"); return !e || std::atoi(e) != 0; }();
    const bool lean_on = [] { const char* e = std::getenv("MUSE_LEAN"); return !e || std::atoi(e) != 0; }();
    const bool funnel_on = [] { const char* e = std::getenv("MUSE_FUNNEL_SPEC"); return !e || std::atoi(e) != 0; }();
    if (h->persist_grid < 0) {
        int g = 0, t = 0;
        if (iso_persist_geometry(h->geo, h->cfg.device, &g, &t) != cudaSuccess) { cudaGetLastError(); g = 0; }
        h->persist_grid = g;
        h->persist_threads = t;
    }
    const int fd_items = std::max(0, nh_mine) * nt * 2;
    bool p2p_fit = false;
    if (multi && h->p2p_ready && p2p_on && h->p2p_nranks == P.nranks && h->p2p_rank == h->comm_rank) {
        int maxh = 1;
        if (get_covariance) for (int q = 0; q < P.nranks; ++q) maxh = std::max(maxh, (int)counts_h[q]);
        const long long need_fd = (long long)maxh * 2 * nt * nt;
        p2p_fit = (long long)P.nranks * std::max(P.need, need_fd) <= h->p2p_block;
    }
    // what makes units leave the fast path is the tolerance (below round-off) or the data, not the launch mechanics
    std::vector<unsigned char> pkey;
    {
        const double dbl[] = {atol};
        const int ints[] = {first_start, nloc, nsims_total};
        pkey.insert(pkey.end(), reinterpret_cast<const unsigned char*>(dbl), reinterpret_cast<const unsigned char*>(dbl) + sizeof(dbl));
        pkey.insert(pkey.end(), reinterpret_cast<const unsigned char*>(ints), reinterpret_cast<const unsigned char*>(ints) + sizeof(ints));
    }
    if (persist_on && h->persist_grid > 0 && nt <= 2 && !h->dbg && !h->fd_start_user && (!multi || p2p_fit) && pkey != h->persist_off_key) {
        const auto t0 = std::chrono::steady_clock::now();
        if (!h->persist_ctl) {
            OUTER_TRY(h, cudaMalloc(&h->persist_ctl, sizeof(PersistCtl)));
            OUTER_TRY(h, cudaMemsetAsync(h->persist_ctl, 0, sizeof(PersistCtl), h->stream));
        }
        PersistParams Q{};
        muse_fill_common(h, Q.base);
        Q.base.atol = atol;
        Q.base.zA = h->zA; Q.base.zB = h->zB; Q.base.zstate = h->zstate;
        Q.base.dbg = nullptr;
        Q.base.seg_chunks = h->geo.seg_chunks;
        Q.base.nseg = h->geo.stream == 1 ? h->geo.nseg : 1;
        Q.step = P;
        Q.cov.nt = nt; Q.cov.family = P.family; Q.cov.d = P.d; Q.cov.nranks = P.nranks; Q.cov.n_total = nsims_total; Q.cov.need = P.need;
        for (int q = 0; q < P.nranks; ++q) Q.cov.counts[q] = P.counts[q];
        Q.cov.st = sd;
        if (muse_theta_consts(h->cfg, theta0, theta0, &Q.first.smp[0], &Q.first.ev) != 0) { h->err = "family"; return MUSE_EUNSUPPORTED; }
        for (int c = 0; c < nt; ++c) Q.theta0[c] = theta0[c];
        Q.first_kind = first_start == MUSE_START_USER ? kStartSharedKeep : kStartZero;
        Q.max_pass = std::min((int)maxsteps, kOuterSlots);
        Q.get_cov = get_covariance ? 1 : 0;
        Q.lazy = lazy_on ? 1 : 0;
        Q.lean = lean_on ? (h->cfg.family == MUSE_FAMILY_FUNNEL && funnel_on ? 2 : 1) : 0;     // 2: + the funnel's specialised element code
        if (h->geo.stream == 1) {           // room in the segment-sum array for the fiducial unit cut into single chunks?
            const long long nchunks = (h->ld + 2047) / 2048;
            Q.fid_seg_chunks = (long long)h->out_cap * h->geo.nseg >= nchunks ? 1 : 0;
        }
        Q.nh_mine = std::max(0, nh_mine);
        Q.z0user = h->z0user;
        const bool hshard = h->cfg.nsims_h > 0;
        Q.xi_fd = hshard ? h->xi_h : h->xi;
        Q.nu_fd = hshard ? h->nu_h : h->nu;
        Q.zfidA = h->zfidA; Q.zfidB = h->zfidB;
        auto optrs = [](const OutBlock& ob) { return OutPtrs{ob.g_d, ob.gnorm_d, ob.f_d, ob.iters_d, ob.fg_d, ob.status_d}; };
        for (int s2 = 0; s2 < kOuterSlots; ++s2) Q.slot[s2] = optrs(h->outer_slot[s2]);
        Q.fd = optrs(h->outer_fd);
        Q.ctl = static_cast<PersistCtl*>(h->persist_ctl);
        Q.stamps = sd->stamp;
        // results streamed to the pinned mirrors by the kernel itself (MUSE_HOSTWRITE=0: copy nodes behind the kernel instead)
        const bool hostwrite = [] { const char* e = std::getenv("MUSE_HOSTWRITE"); return !e || std::atoi(e) != 0; }();
        if (hostwrite) {
            for (int s2 = 0; s2 < kOuterSlots; ++s2) { Q.slot_d[s2] = h->outer_slot[s2].d; Q.slot_h[s2] = h->outer_slot[s2].hst; }
            Q.slot_bytes = h->outer_slot[0].bytes;
            Q.fd_d = h->outer_fd.d; Q.fd_h = h->outer_fd.hst;
            Q.fd_bytes = nh_mine > 0 && get_covariance ? muse_outblock_bytes(h, fd_items) : 0;
            Q.st_h = sh_;
            // … and the host waits for the kernel's completion word instead of the stream (MUSE_HOSTSPIN=0: cudaStreamSynchronize)
            static const bool spin_on = [] { const char* e = std::getenv("MUSE_HOSTSPIN"); return !e || std::atoi(e) != 0; }();
            if (spin_on) {
                if (!h->persist_done_h) {
                    OUTER_TRY(h, cudaMallocHost(&h->persist_done_h, 64));
                    *h->persist_done_h = 0ULL;
                }
                Q.done_h = h->persist_done_h;
                Q.done_seq = ++h->persist_done_seq;
            }
        }
        Q.x.nranks = 1;
        size_t x_blk = 0;
        int parity = 0;
        if (multi) {
            const unsigned long long seq = ++h->p2p_seq;
            parity = (int)(seq & 1ULL);
            x_blk = (size_t)h->p2p_block;
            Q.x.nranks = P.nranks;
            Q.x.rank = h->comm_rank;
            Q.x.row0 = 0;
            for (int q = 0; q < h->comm_rank; ++q) Q.x.row0 += counts[q];
            Q.x.epoch0 = seq * 8ULL;
            int maxh = 1;
            for (int q = 0; q < P.nranks; ++q) {
                Q.x.counts_h[q] = get_covariance ? counts_h[q] : 0;
                maxh = std::max(maxh, Q.x.counts_h[q]);
            }
            Q.x.need_fd = (long long)maxh * 2 * nt * nt;
            for (int q = 0; q < P.nranks; ++q) {
                unsigned char* base = h->p2p_peer[q];
                Q.x.flags[q] = reinterpret_cast<unsigned long long*>(base);
                double* blocks = reinterpret_cast<double*>(base + 256) + (size_t)parity * (kOuterSlots + 1) * x_blk;
                // block order [slot 0 | slot 1 | FD | slot 2]: the typical solve's results are the first three, copied back in one go
                for (int s2 = 0; s2 < kOuterSlots; ++s2) Q.x.gall[q][s2] = blocks + (size_t)(s2 < 2 ? s2 : s2 + 1) * x_blk;
                Q.x.fdall[q] = blocks + (size_t)2 * x_blk;
            }
            for (int s2 = 0; s2 < kOuterSlots; ++s2) Q.cov.g_all_slot[s2] = Q.x.gall[h->comm_rank][s2];
            if (hostwrite) {
                for (int s2 = 0; s2 < kOuterSlots; ++s2) Q.gall_h[s2] = h->p2p_host + (size_t)(s2 < 2 ? s2 : s2 + 1) * x_blk;
                Q.fdall_h = get_covariance ? h->p2p_host + (size_t)2 * x_blk : nullptr;
            }
        } else {
            for (int s2 = 0; s2 < kOuterSlots; ++s2) Q.cov.g_all_slot[s2] = h->outer_slot[s2].g_d + nt;
        }
        cudaEvent_t ea = nullptr, eb = nullptr;
        if (h->prof) {
            if (!h->persist_ev[0]) {
                OUTER_TRY(h, cudaEventCreate(&h->persist_ev[0]));
                OUTER_TRY(h, cudaEventCreate(&h->persist_ev[1]));
            }
            muse_persist_flush(h);                    // the pair is reused: fold the previous launch's times in first
            ea = h->persist_ev[0]; eb = h->persist_ev[1];
            OUTER_TRY(h, cudaEventRecord(ea, h->stream));
        }
        const cudaError_t le = launch_iso_persist(Q, h->geo, h->persist_grid, h->stream);
        const auto t_launched = std::chrono::steady_clock::now();
        if (le != cudaSuccess) {
            // the cooperative launch was refused (not every CTA can be resident: SMs shared with another context, MPS limits …):
            // nothing has run — this handle uses the chain of launches from now on.  A single rank must not decide that alone.
            cudaGetLastError();
            if (multi) { h->err = std::string("muse_solve: cooperative launch failed on a rank of a multi-GPU solve: ") + cudaGetErrorString(le); return MUSE_ECUDA; }
            h->persist_grid = 0;
            goto chain_of_launches;
        }
        if (h->prof) OUTER_TRY(h, cudaEventRecord(eb, h->stream));
        h->acc.launches += 1;
        if (!hostwrite) {
            // results: [state | slot 0 | slot 1 | FD block] in one copy (the typical solve); slot 2 only if a third pass ran
            OUTER_TRY(h, cudaMemcpyAsync(h->outer_arena_h, h->outer_arena_d, h->outer_arena_head, cudaMemcpyDeviceToHost, h->stream));
            if (multi)
                OUTER_TRY(h, cudaMemcpyAsync(h->p2p_host, Q.x.gall[h->comm_rank][0], (size_t)3 * x_blk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        bool spun = false;
        if (Q.done_h) {
            // the last CTA stores done_seq after every host mirror is complete: results are readable ≈ 15 µs before the stream reports
            // idle.  The stream is queried now and then so that a faulted launch cannot hang the caller.
            volatile unsigned long long* flag = Q.done_h;
            const auto t_spin = std::chrono::steady_clock::now();
            double next_query_s = 2e-3;
            for (unsigned n = 1;; ++n) {
                if (*flag == Q.done_seq) { spun = true; break; }
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
                if ((n & 0xFFu) == 0) {
                    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_spin).count();
                    if (el > next_query_s) {
                        const cudaError_t q = cudaStreamQuery(h->stream);
                        if (q != cudaErrorNotReady) { if (q != cudaSuccess) cudaGetLastError(); spun = (*flag == Q.done_seq); break; }
                        next_query_s = el + 1e-3;
                    }
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        if (!spun) OUTER_TRY(h, cudaStreamSynchronize(h->stream));
        if (sh_->error == 3) {
            cudaMemsetAsync(h->persist_ctl, 0, sizeof(PersistCtl), h->stream);
            h->err = "muse_solve: a rank did not arrive at the exchange step within 4 s (peer-mapped exchange of the persistent launch)";
            return MUSE_ECUDA;
        }
        if (sh_->abort) {
            // a unit left the fast path somewhere (on any rank: the flag travels with the exchange): this configuration goes
            // through the chain of launches, whose generic kernel re-solves such units — from scratch, now and from now on
            h->persist_off_key = pkey;
            if (lazy_on) OUTER_TRY(h, cudaMemsetAsync(h->zstate, 0, (size_t)h->rows * sizeof(int), h->stream));
        } else {
            const int n_now = sh_->n_iter;
            if (n_now > 2 && !hostwrite) {
                OUTER_TRY(h, cudaMemcpyAsync(h->outer_slot[2].hst, h->outer_slot[2].d, h->outer_slot[2].bytes, cudaMemcpyDeviceToHost, h->stream));
                if (multi)
                    OUTER_TRY(h, cudaMemcpyAsync(h->p2p_host + (size_t)3 * x_blk, Q.x.gall[h->comm_rank][2], x_blk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
                OUTER_TRY(h, cudaStreamSynchronize(h->stream));
            }
            const double chunk_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (sh_->error == 1) { h->err = "muse!: MAP solution failed with a non-finite objective"; return MUSE_ESTATE; }
            const bool cov_ran = get_covariance && sh_->done != 0 && sh_->error == 0;
            static const bool dbg_timing = [] { const char* e = std::getenv("MUSE_DEBUG_TIMING"); return e && *e; }();
            if (dbg_timing) {
                const long long* T = sh_->stamp;
                char line[512];
                int off = std::snprintf(line, sizeof line, "[muse_solve persist rank %d] n_iter=%d host: entry→launched %.1f us (set-up+launch %.1f), until synchronised %.1f us | kernel stamps (us since start):",
                                        h->comm_rank, n_now, std::chrono::duration<double>(t_launched - t_entry).count() * 1e6,
                                        std::chrono::duration<double>(t_launched - t0).count() * 1e6, chunk_s * 1e6);
                for (int k = 1; k < 13 && off < (int)sizeof line - 16; ++k)
                    off += std::snprintf(line + off, sizeof line - off, " %.1f", T[k] ? (double)(T[k] - T[0]) * 1e-3 : -1.0);
                std::fprintf(stderr, "%s\n", line);
            }
            {   // statistics: one launch; units and algorithmic bytes of the phases that ran, their times from CTA 0's globaltimer stamps
                const double d8 = 8.0 * h->cfg.d;
                const long long* T = sh_->stamp;
                double units_sum = 0.0, bytes_sum = 0.0;
                auto add = [&](int kind, double u, double b, long long t_a, long long t_b) {
                    units_sum += u; bytes_sum += b;
                    if (!h->prof) return;
                    h->acc_pass.launches[kind] += 1; h->acc_pass.ms[kind] += (double)(t_b - t_a) * 1e-6;
                    h->acc_pass.units[kind] += u; h->acc_pass.bytes[kind] += b;
                };
                for (int i = 1; i <= n_now; ++i) {
                    const bool warm = i > 1 || first_start == MUSE_START_USER;
                    const double z0b = warm ? d8 : 0.0;
                    // algorithmic bytes: per sim read ξ, ν [+ z₀], write ẑ; lazy ẑ: read ξ, ν only (a user start row is shared by
                    // all units: counted once), ẑ written only by a pass that materialises it
                    double bytes = nloc * (3 * d8 + z0b) + (2 * d8 + z0b);
                    if (lazy_on) {
                        const bool stored = i == Q.max_pass && Q.max_pass < maxsteps;
                        bytes = nloc * 2 * d8 + d8 + (first_start == MUSE_START_USER ? d8 : 0.0) + (stored ? units * d8 : 0.0);
                    }
                    add(warm ? MUSE_PASS_WARM : MUSE_PASS_COLD, units, bytes, i == 1 ? T[0] : T[2 + 2 * (i - 2)], T[1 + 2 * (i - 1)]);
                }
                if (cov_ran && nh_mine > 0) {
                    add(MUSE_PASS_FIDUCIAL, 1, 3 * d8, T[2 + 2 * (n_now - 1)], T[1 + 2 * kPhaseFid]);
                    add(MUSE_PASS_FD, fd_items, fd_items * 2 * d8 + d8, T[1 + 2 * kPhaseFid], T[1 + 2 * kPhaseFd]);
                }
                h->acc.solve_launches += 1;
                if (h->prof) {              // the launch's event pair: now if the stream has been synchronised, else when it has retired
                    h->persist_pend = true; h->persist_pend_units = units_sum; h->persist_pend_bytes = bytes_sum;
                    if (!spun) muse_persist_flush(h);
                }
            }
            // lazy ẑ: the state cells hold level masks unless the last pass materialised ẑ — no resident ẑ is left behind
            if (lazy_on && !(n_now == Q.max_pass && Q.max_pass < maxsteps))
                OUTER_TRY(h, cudaMemsetAsync(h->zstate, 0, (size_t)h->rows * sizeof(int), h->stream));
            if (multi) { gall_h = h->p2p_host; gall_h_stride = x_blk; gall_dense = true; }
            copy_history(1, n_now, chunk_s);
            if (multi) { gall_h = h->outer_gall_h; gall_h_stride = gall_doubles; gall_dense = false; }
            out->n_iter = n_now;
            for (int c = 0; c < nt; ++c) out->theta_final[c] = sh_->theta[c];
            if (sh_->error == 2) { h->err = "DomainError: sqrt of a negative number in the θ convergence test (src/muse.jl:165)"; return MUSE_ESTATE; }
            it_done = n_now;
            finished = sh_->done != 0 || it_done >= maxsteps;
            if (finished && cov_ran) {
                if (multi) {        // every rank's FD scores are here: combine them all, in rank order
                    Hs_all.assign((size_t)std::max(1, (int)nsims_h_total) * nt * nt, 0.0);
                    const double* fd_h = h->p2p_host + (size_t)2 * x_blk;
                    size_t off = 0;
                    for (int q = 0; q < P.nranks; ++q) {
                        const double* rows = fd_h + (size_t)q * Q.x.need_fd;
                        for (size_t e = 0; e < (size_t)counts_h[q] * 2 * nt * nt; ++e)
                            if (std::isnan(rows[e])) { h->err = "get_H!: MAP solution failed with a non-finite objective"; return MUSE_ESTATE; }
                        if (counts_h[q] > 0) muse_fd_combine_host(h, rows, nullptr, sh_->step, counts_h[q], Hs_all.data() + off, nullptr);
                        off += (size_t)counts_h[q] * nt * nt;
                    }
                }
            } else if (!finished) {
                // the loop goes on in chunks of launches: they take the constants of the next pass from outer_dyn
                OUTER_TRY(h, cudaMemcpyAsync(&dyn[(n_now + 1) & 1], &static_cast<PersistCtl*>(h->persist_ctl)->dyn[0], sizeof(DynConsts),
                                             cudaMemcpyDeviceToDevice, h->stream));
            } else if (get_covariance) {
                h->err = "muse_solve: internal error (the persistent launch ended the loop without its covariance stage)";
                return MUSE_ESTATE;
            }
        }
    }

chain_of_launches:
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
        copy_history(first, std::min(n_now, last), chunk_s);
        out->n_iter = n_now;
        for (int c = 0; c < nt; ++c) out->theta_final[c] = sh_->theta[c];
        if (sh_->error == 2) { h->err = "DomainError: sqrt of a negative number in the θ convergence test (src/muse.jl:165)"; return MUSE_ESTATE; }
        it_done = n_now;
        finished = sh_->done != 0 || it_done >= maxsteps;
        if (!finished && n_now < last) { h->err = "muse_solve: internal error (loop stalled)"; return MUSE_ESTATE; }
    }

    static const bool dbg_total = [] { const char* e = std::getenv("MUSE_DEBUG_TIMING"); return e && *e; }();
    struct TotalTimer {
        std::chrono::steady_clock::time_point t0; bool on;
        ~TotalTimer() { if (on) std::fprintf(stderr, "[muse_solve] whole call %.1f us\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6); }
    } total_timer{t_entry, dbg_total};
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
        return muse_cov_finish(h, out->theta_final, gs_last, nsims_total, nsims_h_total, counts_h, Hs_local.data(), nh_mine, prior_sigma, cov,
                               Hs_all.empty() ? nullptr : Hs_all.data());
    }
    return MUSE_OK;
}
