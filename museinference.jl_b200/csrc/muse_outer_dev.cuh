// muse_outer_dev.cuh — device code of the outer θ iteration of muse! (/root/reference/src/muse.jl:163-166, 183-224) and of the
// covariance stage's preparation (:411-413), shared by
//   * theta_step_kernel / cov_prep_kernel (muse_outer.cu): one single-CTA launch between two solver passes, and
//   * solve_persist_kernel (muse_iso_stream.cu): the whole solve in ONE cooperative launch, where CTA 0 runs this code between
//     the passes and releases the other CTAs through a flag.
// Both run the SAME reduction tree — 1024 "lanes" (lane t sums sims t, t + 1024, … in that order), a shuffle butterfly per
// group of 32 lanes, a second one over the 32 group results — whatever the CTA size: a CTA of 1024 / V threads gives every
// thread V lanes.  θ is therefore bit-identical between the two drivers, for any sharding and any number of GPUs.
#pragma once
#include "muse_handle.cuh"

namespace muse {

constexpr int kStepLanes = 1024;     // lanes of the reduction tree (= threads of theta_step_kernel)
constexpr int kMaxRanks = 16;

struct OuterParams {
    int nt, family, d;
    int iter;                 // 1-based index of the pass just executed
    int maxsteps;
    int units_local;          // units of the pass on this rank (data + local sims)
    int nranks, n_total;
    int counts[kMaxRanks];    // sims per rank
    long long need;           // doubles per rank slot of g_all
    int dense;                // 1: g_all holds the N sims' rows consecutively in global order (the peer exchange of the one-launch
                              //    solve writes them there); 0: rank q's rows start at q·need (NCCL all-gather slots)
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

// ---- whole solve in one launch (solve_persist_kernel) ----------------------------------------------------------------
constexpr int kPersistPhases = kOuterSlots + 2;       // passes of the chunk, fiducial solve, finite-difference sims
constexpr int kPhaseFid = kOuterSlots, kPhaseFd = kOuterSlots + 1;

// Control block in device memory, all zero between launches: the last CTA to leave a launch clears it again.
struct PersistCtl {
    int work[8];              // per phase: dynamic work counter of the streaming pass
    int redo[8];              // per phase: units handed back (any ⇒ the launch gives up, the host re-runs the solve on the chain of launches)
    int arrive[8];            // per phase: CTAs that have finished their share
    int zfid_state;           // ZState of the fiducial ẑ
    int exit_count;
    unsigned int step_flag;   // released by CTA 0 after the arithmetic that follows phase k (value k + 1)
    int msg_done, msg_error, msg_abort;
    unsigned long long redo_sink;
    DynConsts dyn[2];         // CTA 0 → everyone: constants of the next pass ([0]; after the loop: fiducial solve) and of the FD sims ([1])
};

struct OutPtrs {              // device pointers of one OutBlock
    double *g, *gnorm, *f;
    int *iters, *fg, *status;
};

// cross-GPU exchange through peer-mapped memory (muse_comm.cu: muse_b200_p2p_*): every rank owns one region
//   [ flags: kMaxRanks × 2 u64 | 2 parities × (kOuterSlots score slots + 1 FD slot) ]
// and stores its rows straight into every peer's region, then raises its flag there
struct XchgParams {
    int nranks, rank;
    int row0;                         // global index of this rank's first sim: its score rows go to offset row0·nθ of every gathered slot
    unsigned long long epoch0;        // flags of this solve: epoch0 + phase + 1
    unsigned long long* flags[kMaxRanks];   // peer q's flag array (entry [2·rank] is mine to write there); [rank] = my own region
    double* gall[kMaxRanks][kOuterSlots];   // peer q's gathered-score slot s (this solve's parity)
    double* fdall[kMaxRanks];               // peer q's gathered FD-score block
    long long need_fd;                      // doubles per rank slot of the FD block
    int counts_h[kMaxRanks];
};

struct PersistParams {
    SolveLaunch base;         // everything the passes share (sizes, draws, scratch, atol …); per-phase fields are filled in on the device
    OuterParams step;
    CovParams cov;
    XchgParams x;
    DynConsts first;          // constants of pass 1 at θ₀ (host libm, like the other drivers)
    double theta0[kMaxTheta];
    int first_kind;           // StartKind of pass 1
    int max_pass;             // passes this launch may run (≤ kOuterSlots)
    int get_cov, nh_mine;
    int fid_seg_chunks;       // > 0: chunks per segment of the fiducial solve's one unit (finer than the passes': more CTAs share it)
    int lean;                 // SolveLaunch::lean in every phase: 0 off, 1 lean α = 1 trial, 2 lean + the funnel's specialised element code
    int lazy;                 // 1: ẑ is recomputed from the base normals instead of stored and re-read (muse_common.cuh: LazyLevels)
    const double *z0user, *xi_fd, *nu_fd;
    double *zfidA, *zfidB;
    OutPtrs slot[kOuterSlots], fd;
    // Results go to the host WHILE the solve runs: the pinned mirrors of the output blocks are mapped into the device's address
    // space, and after every pass each CTA stores its slice of that pass's block there (posted PCIe writes that overlap the next
    // phase); CTA 0 adds the FD block, the state and — multi-GPU — the gathered rows at the end.  The host only synchronises:
    // no copy node behind the kernel (≈ 10 µs of latency for C3's 150 KB, 30 µs for C2's 0.7 MB).
    const unsigned char* slot_d[kOuterSlots];     // device blocks (whole OutBlock, capacity layout) and their host mirrors
    unsigned char* slot_h[kOuterSlots];
    const unsigned char* fd_d;
    unsigned char* fd_h;
    unsigned long long slot_bytes, fd_bytes;
    OuterState* st_h;
    unsigned long long* done_h;                   // pinned word: the last CTA to leave stores done_seq there once every host mirror is complete
    unsigned long long done_seq;
    double* gall_h[kOuterSlots];                  // multi-GPU: host mirrors of this rank's gathered-score slots / FD block
    double* fdall_h;
    PersistCtl* ctl;
    long long* stamps;        // globaltimer stamps of CTA 0: [0] start, then per phase [1 + 2k] units done, [2 + 2k] arithmetic done
};

#if defined(__CUDACC__)

// θ → constants, as theta_consts() of muse_api.cu (device libm)
__device__ inline void consts_of(int family, int d, const double* th_sim, const double* th_eval, IsoSample* smp, IsoEval* ev) {
    const double dd = (double)d;
    if (family == MUSE_FAMILY_FUNNEL) {
        if (smp) { smp->sig = exp(0.5 * th_sim[0]); smp->mu = 0.0; }
        if (ev) { ev->a = exp(-th_eval[0]); ev->mu = 0.0; ev->half_cst = 0.5 * dd * th_eval[0]; ev->cspec = 1.0 / (1.0 + ev->a); }
    } else {
        if (smp) { smp->sig = exp(th_sim[1]); smp->mu = th_sim[0]; }
        if (ev) { ev->a = exp(-2.0 * th_eval[1]); ev->mu = th_eval[0]; ev->half_cst = dd * th_eval[1]; ev->cspec = 1.0 / (1.0 + ev->a); }
    }
}

// Σ_k f_c(k) for every θ-component c over all sims, in an order fixed by the GLOBAL sim index (see the file comment).
// Called by ALL threads of the CTA; the first kStepLanes / V of them carry V lanes each.  f(o, c) reads element c of the score
// row at offset o of the gathered layout (rank q's rows start at q·need) — a bare load; map(v, c) is what is summed.  Keeping the two
// apart lets every load of a batch be in flight before the first value is touched (arithmetic inside f serialised the loads: one
// round trip to the L2 per sim and lane, 34 µs per θ-step over 10⁴ sims).  Results land in out[0..nt) of every thread.
template <int V, int NT, class F, class M>
__device__ inline void block_sums(const int* counts, int nranks, long long need, int nt, int n_total, F&& f, M&& map, double (*sh)[32], double* out) {
    constexpr int T = kStepLanes, PT = T / V;
    const int lane = threadIdx.x & 31, pw = threadIdx.x >> 5;
    auto row_off = [&](int k) -> size_t {                 // global sim k → offset of its row
        if (nranks == 1) return (size_t)k * nt;           // (uniform branch: one rank, rows are simply consecutive)
        int q = 0, base = 0;
        while (q + 1 < nranks && k >= base + counts[q]) { base += counts[q]; ++q; }
        return (size_t)q * (size_t)need + (size_t)(k - base) * nt;
    };
    if ((int)threadIdx.x < PT) {
        double acc[V][NT];
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
            for (int c = 0; c < NT; ++c) acc[j][c] = 0.0;
        // every lane walks its sims in increasing order (k = t, t + T, t + 2T, …).  The loads of four steps of ALL the thread's V
        // lanes are issued before the first addition: a lone CTA is bound by the latency of dependent round trips to the L2, not by
        // bandwidth (with the lanes one after the other a θ-step over 10⁴ sims took 29 µs)
        constexpr int S = 4;
        for (int k0 = 0; k0 < n_total; k0 += S * T) {
            double v[S][V][NT];
            bool ok[S][V];
#pragma unroll
            for (int st = 0; st < S; ++st)
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const int k = k0 + st * T + (pw + j * (PT / 32)) * 32 + lane;
                    ok[st][j] = k < n_total;
                    const size_t o = ok[st][j] ? row_off(k) : 0;
#pragma unroll
                    for (int c = 0; c < NT; ++c) v[st][j][c] = (ok[st][j] && c < nt) ? f(o, c) : 0.0;
                }
#pragma unroll
            for (int st = 0; st < S; ++st)
#pragma unroll
                for (int j = 0; j < V; ++j)
#pragma unroll
                    for (int c = 0; c < NT; ++c)
                        if (ok[st][j] && c < nt) acc[j][c] += map(v[st][j][c], c);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                if (c < nt) {
                    double v = acc[j][c];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) sh[c][pw + j * (PT / 32)] = v;
                }
            }
        }
    }
    __syncthreads();
    for (int c = 0; c < nt; ++c) {
        double v = sh[c][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        out[c] = v;                                      // every warp computes the same value
    }
    __syncthreads();
}

// mean and corrected variance of every component (two passes, like Statistics.mean / var)
template <int V, int NT>
__device__ inline void block_mean_var_nt(const double* g, const int* counts, int nranks, long long need, int nt, int n_total,
                                         double (*sh)[32], double* mean, double* var, bool* saw_nan = nullptr) {
    auto ld = [&](size_t o, int c) { return __ldcg(g + o + c); };
    bool nanv = false;       // multi-GPU: a failed sim arrives as a NaN row — noticed while the mean is summed, no scan of its own
    block_sums<V, NT>(counts, nranks, need, nt, n_total, ld, [&](double v, int) { nanv = nanv || isnan(v); return v; }, sh, mean);
    if (saw_nan) *saw_nan = nanv;
    for (int c = 0; c < nt; ++c) mean[c] /= n_total;
    block_sums<V, NT>(counts, nranks, need, nt, n_total, ld, [&](double v, int c) { const double dlt = v - mean[c]; return dlt * dlt; }, sh, var);
    for (int c = 0; c < nt; ++c) var[c] /= (n_total - 1);
}
template <int V>
__device__ inline void block_mean_var(const double* g, const int* counts, int nranks, long long need, int nt, int n_total,
                                      double (*sh)[32], double* mean, double* var, bool* saw_nan = nullptr) {
    if (nt == 1) block_mean_var_nt<V, 1>(g, counts, nranks, need, nt, n_total, sh, mean, var, saw_nan);
    else if (nt == 2 || V > 1) block_mean_var_nt<V, 2>(g, counts, nranks, need, nt, n_total, sh, mean, var, saw_nan);   // V > 1 (one-launch solve): nθ ≤ 2 (host check)
    else block_mean_var_nt<V, kMaxTheta>(g, counts, nranks, need, nt, n_total, sh, mean, var, saw_nan);
}

// solve_persist_kernel: CTA 0 lives through the whole solve, so what one θ-step needs of the previous one (θ, the last history
// row, the variance the covariance stage starts from) stays in its shared memory instead of making dependent round trips to the L2
struct StepCache {
    double theta[kMaxTheta];        // θ after the last update
    double row_theta[kMaxTheta];    // θ at which the last executed iteration evaluated
    double var[kMaxTheta];          // corrected variance of the last scores
    int n_iter, done, error;
};

// What the host loop does between two passes (src/muse.jl:183-224, the test of :163-166 for the NEXT iteration), by all
// threads of one CTA (≥ kStepLanes / V threads).  Thread 0 writes the history row, θ, the flags and the constants of the next
// pass (*dyn_next, which may live in global or shared memory).
template <int V>
__device__ inline void theta_step_body(const OuterParams& P, double (*sh)[32], int* bad, StepCache* cache = nullptr) {
    OuterState* st = P.st;
    if (threadIdx.x == 0) *bad = 0;
    __syncthreads();
    // src/interface.jl:170.  The decision must be the same on every rank: with several ranks it is taken from what all of them
    // see — the replicated data unit and the gathered rows, in which a failed sim arrives as NaN (muse_comm.cu / the exchange
    // of solve_persist_kernel)
    // (loads issued eight at a time: one CTA scanning 10⁴ units is bound by the latency of dependent round trips to the L2)
    bool mine = false;
    const bool fused_nan = P.nranks > 1 && P.dense;          // the NaN rows are noticed by the mean pass below
    if (P.nranks > 1) {
        if (threadIdx.x == 0 && __ldcg(P.status_local) == MUSE_STATUS_NONFINITE) mine = true;
        for (int q = 0; q < P.nranks && !fused_nan; ++q) {
            const double* g = P.g_all + (size_t)q * P.need;
            const int n = P.counts[q] * P.nt;
            for (int e = threadIdx.x; e < n; e += 8 * blockDim.x) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int i = e + u * blockDim.x; v[u] = i < n ? __ldcg(g + i) : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; ++u) mine = mine || isnan(v[u]);
            }
        }
    } else {
        const int n = P.units_local;
        for (int e = threadIdx.x; e < n; e += 8 * blockDim.x) {
            int v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = e + u * blockDim.x; v[u] = i < n ? __ldcg(P.status_local + i) : 0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) mine = mine || v[u] == MUSE_STATUS_NONFINITE;
        }
    }
    double mean[kMaxTheta], var[kMaxTheta];
    if (fused_nan) {
        // dense layout (one-launch solve): rows of all ranks are consecutive in global sim order — mean, variance and the NaN
        // test in the two passes the statistics need anyway.  (A NaN row only poisons sums that are thrown away.)
        bool saw = false;
        block_mean_var<V>(P.g_all, P.counts, 1, P.need, P.nt, P.n_total, sh, mean, var, &saw);
        mine = mine || saw;
    }
    if (mine) *bad = 1;
    __syncthreads();
    if (*bad) {
        if (threadIdx.x == 0) {
            st->error = 1; st->done = 1;
            if (cache) { cache->error = 1; cache->done = 1; }
            if (P.dyn_next) P.dyn_next->skip = 1;
        }
        return;
    }
    const int row = P.iter - 1;
    if (!fused_nan) block_mean_var<V>(P.g_all, P.counts, P.dense ? 1 : P.nranks, P.need, P.nt, P.n_total, sh, mean, var);
    if (threadIdx.x != 0) return;
    double th_new[kMaxTheta], hip[kMaxTheta];
    for (int c = 0; c < P.nt; ++c) {
        const double th = cache ? cache->theta[c] : st->theta[c];
        const double g_dat = __ldcg(P.g_local + c);
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
        hip[c] = h_inv_post;
        th_new[c] = th - P.alpha * (h_inv_post * g_post);                                          // :224
    }
    for (int c = 0; c < P.nt; ++c) st->theta[c] = th_new[c];                                       // :230
    st->n_iter = P.iter;
    int done = 0, err2 = 0;
    if (P.iter >= 2) {                                   // the test at the top of iteration iter + 1 > 2   (:163-166)
        double q = 0.0;
        for (int c = 0; c < P.nt; ++c) {
            // (θ of this row = θ before this update; the row before it comes from the cache when there is one — same values)
            const double th_row = cache ? cache->theta[c] : st->row[row].theta[c];
            const double th_prev = cache ? cache->row_theta[c] : st->row[row - 1].theta[c];
            const double dlt = th_row - th_prev;
            q += dlt * hip[c] * dlt;
        }
        q = -q;
        if (q < 0.0) { st->error = 2; err2 = 1; done = 1; }        // DomainError of sqrt in the reference
        else if (sqrt(q) < P.theta_rtol) done = 1;
    }
    if (P.iter >= P.maxsteps) done = 1;
    st->done = done;
    if (cache) {
        for (int c = 0; c < P.nt; ++c) { cache->row_theta[c] = cache->theta[c]; cache->theta[c] = th_new[c]; cache->var[c] = var[c]; }
        cache->n_iter = P.iter; cache->done = done; cache->error = err2 ? 2 : 0;
    }
    if (P.dyn_next) {
        consts_of(P.family, P.d, th_new, th_new, &P.dyn_next->smp[0], &P.dyn_next->ev);
        P.dyn_next->skip = done;
    }
}

// after the last θ-step: if the loop has ended, step = 0.1 ./ std(gs) (:411-413) and the constants of get_H!'s launches;
// otherwise they are skipped.  All threads of one CTA.
template <int V>
__device__ inline void cov_prep_body(const CovParams& P, double (*sh)[32], const StepCache* cache = nullptr) {
    volatile OuterState* st = P.st;      // written by this CTA's thread 0 a barrier ago when the θ-step ran in the same launch
    const int s_done = cache ? cache->done : st->done, s_error = cache ? cache->error : st->error, s_iter = cache ? cache->n_iter : st->n_iter;
    if (!s_done || s_error || s_iter < 1) {
        if (threadIdx.x == 0) { P.dyn_fid->skip = 1; P.dyn_fd->skip = 1; }
        return;
    }
    double step[kMaxTheta], mean[kMaxTheta], var[kMaxTheta];
    if (cache) {
        // the θ-step of this very launch has just reduced the same scores with the same tree: its variance IS the one below
        for (int c = 0; c < P.nt; ++c) var[c] = cache->var[c];
    } else {
        const double* gall = P.g_all_slot[(s_iter - 1) % kOuterSlots];
        block_mean_var<V>(gall, P.counts, P.nranks, P.need, P.nt, P.n_total, sh, mean, var);
    }
    for (int c = 0; c < P.nt; ++c) step[c] = 0.1 / sqrt(var[c]);       // step = 0.1 ./ std(gs)   (:411-413), gs = the last scores (:231)
    if (threadIdx.x != 0) return;
    double th0[kMaxTheta];
    for (int c = 0; c < P.nt; ++c) { th0[c] = cache ? cache->theta[c] : st->theta[c]; st->step[c] = step[c]; }
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

#endif  // __CUDACC__

cudaError_t iso_persist_geometry(const Geometry& geo, int device, int* grid, int* threads);
cudaError_t launch_iso_persist(const PersistParams& P, const Geometry& geo, int grid, cudaStream_t st);

}  // namespace muse
