// muse_handle.cuh — the state behind a muse_handle (shared by muse_api.cu and muse_comm.cu).
#pragma once
#include <string>
#include <vector>

#include "muse_common.cuh"

using muse::Geometry;

struct muse_corr_ctx;   // correlated-Gaussian family (muse_corr.cu)

// One block of per-unit solver outputs (device + pinned host mirror), laid out as muse_api.cu describes.
struct OutBlock {
    unsigned char *d = nullptr, *hst = nullptr;
    size_t bytes = 0;
    int cap = 0;
    bool owned = false;         // false: carved out of a larger allocation (the device-resident loop's arena)
    double *g_d = nullptr, *gnorm_d = nullptr, *f_d = nullptr;
    int *iters_d = nullptr, *fg_d = nullptr, *status_d = nullptr;
    double *g_h = nullptr, *gnorm_h = nullptr;
    int *iters_h = nullptr, *fg_h = nullptr, *status_h = nullptr;
};

// Device-resident outer loop (muse_outer.cu): state of the θ iteration kept on the device between passes.
constexpr int kOuterSlots = 3;      // passes enqueued per host synchronisation (2 iterations + the convergence test is typical)
constexpr int kOuterMaxIter = 64;   // history rows held on the device
struct OuterRow {                                        // history row of one iteration (what muse_iterate_out holds per row)
    double theta[MUSE_MAX_NTHETA];                       // θ at which the iteration evaluated
    double g_dat[MUSE_MAX_NTHETA], g_like[MUSE_MAX_NTHETA], g_prior[MUSE_MAX_NTHETA];
    double h_inv_like[MUSE_MAX_NTHETA], h_prior[MUSE_MAX_NTHETA], h_inv_post[MUSE_MAX_NTHETA];
};
struct OuterState {
    // header: uploaded from pinned staging at the start of a solve (one copy initialises everything below)
    int n_iter, done, error;
    int abort;                                           // solve_persist_kernel gave up (a unit left the fast path): re-run on the chain of launches
    double theta[MUSE_MAX_NTHETA];                       // θ after the last update (= θ of the next pass)
    double step[MUSE_MAX_NTHETA];                        // 0.1 ./ std(gs) of the covariance stage
    int ctr[16];                                         // per launch chain of a chunk: [2k] hand-back count, [2k+1] streaming work
                                                         // counter (k = 0..4); [12] state of the fiducial ẑ — zero at chunk start
    long long stamp[16];                                 // solve_persist_kernel: globaltimer stamps of CTA 0 (muse_outer_dev.cuh)
    muse::DynConsts dyn_first;                           // constants of pass 1 at θ₀ (host libm, like the other drivers)
    // history
    OuterRow row[kOuterMaxIter];
};

struct muse_handle {
    muse_cfg cfg{};
    int ld = 0;
    int rows = 0;               // 1 + nsims (unit 0 = data)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    Geometry geo{};
    bool have_data = false, have_draws = false, have_z0 = false;
    bool fd_start_user = false;     // get_H!'s fiducial solve starts from the user's z₀ (muse_b200_fd_start)

    // device arrays
    double *xi = nullptr, *nu = nullptr;          // (nsims+1) × ld, last row = master draw
    double *xi_h = nullptr, *nu_h = nullptr;      // nsims_h × ld: draws of the get_H! shard (multi-GPU)
    bool have_draws_h = false;
    double *xdat = nullptr, *z0user = nullptr;    // ld
    double *xslot = nullptr;                             // slots × ld (x of the unit in flight, per group)
    double *zA = nullptr, *zB = nullptr;                 // rows × ld
    int* zstate = nullptr;                        // rows
    double *sbuf = nullptr, *dxh = nullptr, *dgh = nullptr;   // per-slot scratch
    // outputs (device + pinned host mirror), capacity out_cap items
    int out_cap = 0;
    unsigned char *out_d = nullptr, *out_h = nullptr;   // one block each; the typed pointers below point into them
    size_t out_bytes = 0;
    double *g_d = nullptr, *gnorm_d = nullptr, *f_d = nullptr;
    int *iters_d = nullptr, *fg_d = nullptr, *status_d = nullptr;
    double *g_h = nullptr, *gnorm_h = nullptr;
    int *iters_h = nullptr, *fg_h = nullptr, *status_h = nullptr;
    // finite-difference scratch
    int h_cap = 0;
    double *zHA = nullptr, *zHB = nullptr;
    double *zfidA = nullptr, *zfidB = nullptr;
    int* zfid_state = nullptr;

    // streaming kernel: per-(unit, segment) partial sums, arrival counters, hand-back list
    double* gpart = nullptr;
    int *gcount = nullptr, *redo_count = nullptr, *redo_items = nullptr;
    unsigned long long* redo_total = nullptr;

    muse_corr_ctx* corr = nullptr;

    // device-resident outer loop (muse_outer.cu)
    // one device allocation with a pinned host mirror of the same layout — [state | slot 0 | slot 1 | FD block | slot 2] — so
    // that the results of a typical solve come back in ONE copy (every DMA node costs ≈ 8 µs of latency)
    unsigned char *outer_arena_d = nullptr, *outer_arena_h = nullptr;
    size_t outer_arena_bytes = 0, outer_arena_head = 0;  // head: bytes up to the end of the FD block
    int outer_units_cap = 0, outer_fd_cap = 0;
    OutBlock outer_fd;                                   // outputs of get_H!'s launches
    int* ctr_override = nullptr;                         // launch_solver: this chain's counter pair instead of redo_count + memset
    int* zfid_override = nullptr;                        // fd_launch: the fiducial ẑ's state cell instead of zfid_state + memset
    OutBlock outer_slot[kOuterSlots];                    // per-pass outputs of the iterations of one chunk
    double* outer_gall[kOuterSlots] = {nullptr, nullptr, nullptr};   // multi-GPU: gathered score rows per pass (device)
    double* outer_gall_h = nullptr;                      // pinned mirror of the three gathered blocks
    size_t outer_gall_doubles = 0;
    OuterState *outer_st_d = nullptr, *outer_st_h = nullptr;
    muse::DynConsts* outer_dyn = nullptr;                // [0], [1]: passes (alternating); [2]: fiducial; [3]: FD sims
    OuterState* outer_st_stage = nullptr;                // pinned: the state header a solve starts from (uploaded by the first node)
    // CUDA graph of the first chunk of the device-resident loop — the whole of a typical solve: one cudaGraphLaunch replaces
    // ≈ 25 stream operations (launches, memsets, copies, event records) whose CPU-side issue cost, not the GPU, bounded the
    // small configurations.  Captured from the very enqueue code the eager path runs, on the second solve with a given key.
    bool capturing = false;                              // launch_solver: record events as external nodes, count into cap_*
    void* outer_exec = nullptr;                          // cudaGraphExec_t
    std::vector<unsigned char> outer_key, outer_warm_key;   // parameters baked into the graph / of the last eager solve
    int64_t cap_launches = 0, cap_solve_launches = 0;    // kernel launches / solver passes inside the graph

    // the whole solve in one cooperative launch (solve_persist_kernel, muse_iso_stream.cu)
    void* persist_ctl = nullptr;                         // muse::PersistCtl, device, zero between launches
    int persist_grid = -1, persist_threads = 0;          // −1: not queried yet; 0: unavailable
    cudaEvent_t persist_ev[2] = {nullptr, nullptr};      // profiling: the launch's event pair, read once the launch has retired (muse_persist_flush)
    bool persist_pend = false;                           // … an event pair not yet folded into acc
    double persist_pend_units = 0.0, persist_pend_bytes = 0.0;
    unsigned long long* persist_done_h = nullptr;        // pinned completion word of the launch (written by its last CTA)
    unsigned long long persist_done_seq = 0;
    std::vector<unsigned char> persist_off_key;          // parameters with which the launch gave up (hand-backs): straight to the chain

    // exchange through peer-mapped memory (muse_comm.cu: muse_b200_p2p_*): this rank's region and the peers' mappings of theirs
    unsigned char* p2p_region = nullptr;                 // [flags 256 B | 2 parities × (kOuterSlots + 1) blocks of p2p_block doubles]
    unsigned char* p2p_peer[16] = {};                    // region of rank q as mapped here ([rank] = p2p_region)
    double* p2p_host = nullptr;                          // pinned mirror of one parity's blocks
    long long p2p_block = 0;
    int p2p_nranks = 0, p2p_rank = 0;
    bool p2p_ready = false;
    unsigned long long p2p_seq = 0;                      // persistent multi-GPU solves so far (flag epochs, buffer parity)

    // exchange step (muse_comm.cu): NCCL communicator and staging buffers
    void* comm = nullptr;
    int comm_nranks = 0, comm_rank = 0, comm_cap = 0;
    double *comm_send = nullptr, *comm_recv = nullptr, *comm_host = nullptr;

    long long* dbg = nullptr;   // diagnostics timeline (muse_b200_debug_timeline)
    int dbg_cap = 0;

    // profiling
    bool prof = false;
    struct Rec { cudaEvent_t a, b; int cls; double units, bytes; int kind; int tag; };
    std::vector<Rec> recs;
    std::vector<Rec> outer_recs;         // event pairs recorded by nodes of the graph (re-recorded at every graph launch)
    muse_profile acc{};
    int pass_kind = MUSE_PASS_COLD;      // kind of the solver pass being enqueued (set by the entry point)
    int rec_tag = 0;                     // device-resident loop: iteration (> 0) or covariance stage (−1) a launch belongs to,
                                         // so that launches the device skipped can be dropped from the statistics afterwards
    muse_pass_profile acc_pass{};        // per-kind split of acc.solve_* (muse_b200_profile_passes)
};

// internal entry points of muse_api.cu used by the device-resident outer loop (muse_outer.cu)
int  muse_theta_consts(const muse_cfg& c, const double* th_sim, const double* th_eval, muse::IsoSample* smp, muse::IsoEval* ev);
int  muse_outblock_ensure(muse_handle* h, OutBlock& ob, int items);
void muse_outblock_free(OutBlock& ob);
int  muse_pass_enqueue(muse_handle* h, const double* theta_sim, const double* theta_eval, double atol, int include_data,
                       int warm_start, int first_sim, int count, const OutBlock* ob, const muse::DynConsts* dyn);
extern "C" int  muse_fd_enqueue(muse_handle* h, const double* theta0, const double* th_pts, int nsims_H, double atol,
                                const muse::DynConsts* dyn_fid, const muse::DynConsts* dyn_fd, const OutBlock* ob);
extern "C" void muse_fd_combine_host(muse_handle* h, const double* g_h, const int* status_h, const double* step, int nsims_H,
                                     double* Hs_out, int32_t* status_out);
size_t muse_outblock_bytes(const muse_handle* h, int items);
void muse_outblock_carve(const muse_handle* h, OutBlock& ob, unsigned char* dev, unsigned char* host, int items);
void muse_outer_release(muse_handle* h);
void muse_persist_flush(muse_handle* h);      // folds the one-launch solve's pending event pair into the profile (waits for it if need be)
extern "C" int  muse_comm_allgather_dev_enqueue(muse_handle* h, const double* src_dev, const int* status_dev, int ncol, const int32_t* counts, size_t* need_out);

void muse_comm_release(muse_handle* h);
void muse_p2p_release(muse_handle* h);
int  muse_ensure_outputs(muse_handle* h, int items);
void muse_fill_common(muse_handle* h, muse::SolveLaunch& L);
extern "C" void muse_comm_unpack(muse_handle* h, int ncol, const int32_t* counts, double* out_host);
extern "C" int  muse_comm_allgather_scores_enqueue(muse_handle* h, int first_row, const int32_t* counts);

// correlated-Gaussian family (muse_corr.cu)
int  muse_corr_create(muse_handle* h);
void muse_corr_destroy(muse_handle* h);
int  muse_corr_set_data(muse_handle* h, const double* x);
int  muse_corr_set_z0(muse_handle* h, const double* z0);
int  muse_corr_set_draws(muse_handle* h, const double* xi, const double* nu, const double* xi_m, const double* nu_m, bool hshard);
int  muse_corr_seed_draws(muse_handle* h, uint64_t seed);
int  muse_corr_map_score(muse_handle* h, const double* theta_sim, const double* theta_eval, double atol, int include_data,
                         int warm_start, int first_sim, int count);
int  muse_corr_fd_launch(muse_handle* h, const double* theta0, const double* th_pts /* 2 sample points */, int nsims_H, double atol);
int  muse_corr_get_maps(muse_handle* h, int first_unit, int count, double* z_out);
bool muse_corr_have_draws(muse_handle* h, bool hshard);
