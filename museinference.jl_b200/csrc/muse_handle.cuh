// muse_handle.cuh — the state behind a muse_handle (shared by muse_api.cu and muse_comm.cu).
#pragma once
#include <string>
#include <vector>

#include "muse_common.cuh"

using muse::Geometry;

struct muse_corr_ctx;   // correlated-Gaussian family (muse_corr.cu)

struct muse_handle {
    muse_cfg cfg{};
    int ld = 0;
    int rows = 0;               // 1 + nsims (unit 0 = data)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    Geometry geo{};
    bool have_data = false, have_draws = false, have_z0 = false;

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

    // exchange step (muse_comm.cu): NCCL communicator and staging buffers
    void* comm = nullptr;
    int comm_nranks = 0, comm_rank = 0, comm_cap = 0;
    double *comm_send = nullptr, *comm_recv = nullptr, *comm_host = nullptr;

    long long* dbg = nullptr;   // diagnostics timeline (muse_b200_debug_timeline)
    int dbg_cap = 0;

    // profiling
    bool prof = false;
    struct Rec { cudaEvent_t a, b; int cls; double units, bytes; int kind; };
    std::vector<Rec> recs;
    muse_profile acc{};
    int pass_kind = MUSE_PASS_COLD;      // kind of the solver pass being enqueued (set by the entry point)
    muse_pass_profile acc_pass{};        // per-kind split of acc.solve_* (muse_b200_profile_passes)
};

void muse_comm_release(muse_handle* h);
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
