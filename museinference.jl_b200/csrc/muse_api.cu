// muse_api.cu — the C ABI of libmuse_b200.so (include/muse_b200.h).
//
// Host-side glue only: argument checking, device-memory ownership, θ → kernel-constant
// conversion, launch of the persistent solver, result copies.  All arithmetic on the N×d batch
// happens in the CUDA kernels (muse_iso_stream.cu, muse_iso_solver.cu, muse_draws.cu).  There is no CPU fallback:
// every entry point needs a live CUDA context on an sm_100 device.
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "muse_handle.cuh"

using namespace muse;

static thread_local std::string g_create_err;

#define MUSE_FAIL(h, code, msg)        \
    do {                               \
        (h)->err = (msg);              \
        return (code);                 \
    } while (0)

#define CUDA_TRY(h, expr)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
            return e__ == cudaErrorMemoryAllocation ? MUSE_ENOMEM : MUSE_ECUDA;               \
        }                                                                                     \
    } while (0)

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// θ → constants of the isotropic families (host libm, so oracle and kernel see identical scalars)
static int theta_consts(const muse_cfg& c, const double* th_sim, const double* th_eval, IsoSample* smp, IsoEval* ev) {
    const double d = (double)c.d;
    if (c.family == MUSE_FAMILY_FUNNEL) {
        if (smp) { smp->sig = std::exp(0.5 * th_sim[0]); smp->mu = 0.0; }
        if (ev) { ev->a = std::exp(-th_eval[0]); ev->mu = 0.0; ev->half_cst = 0.5 * d * th_eval[0]; ev->cspec = 1.0 / (1.0 + ev->a); }
        return 0;
    }
    if (c.family == MUSE_FAMILY_HIERGAUSS) {
        if (smp) { smp->sig = std::exp(th_sim[1]); smp->mu = th_sim[0]; }
        if (ev) { ev->a = std::exp(-2.0 * th_eval[1]); ev->mu = th_eval[0]; ev->half_cst = d * th_eval[1]; ev->cspec = 1.0 / (1.0 + ev->a); }
        return 0;
    }
    return -1;
}

int muse_theta_consts(const muse_cfg& c, const double* th_sim, const double* th_eval, IsoSample* smp, IsoEval* ev) {
    return theta_consts(c, th_sim, th_eval, smp, ev);
}

// Outputs live in ONE device block with a pinned host mirror of the same layout
//   [ g: cap×nθ f64 | ‖∇z‖∞: cap f64 | f: cap f64 | iterations: cap i32 | f/g evaluations: cap i32 | status: cap i32 ]
// so that a fetch is a single device→host copy (five small copies cost ~8 µs each on the copy engine).
static int ensure_outputs(muse_handle* h, int items) {
    if (items <= h->out_cap) return 0;
    cudaFree(h->out_d); cudaFreeHost(h->out_h);
    cudaFree(h->gpart); cudaFree(h->gcount); cudaFree(h->redo_items);
    h->out_d = h->out_h = nullptr;
    h->gpart = nullptr; h->gcount = h->redo_items = nullptr;
    h->out_cap = 0;
    const size_t n = (size_t)items, nt = (size_t)h->cfg.ntheta;
    if (h->geo.stream) {
        CUDA_TRY(h, cudaMalloc(&h->gpart, n * (size_t)h->geo.nseg * 16 * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&h->redo_items, n * sizeof(int)));
        CUDA_TRY(h, cudaMalloc(&h->gcount, n * sizeof(int)));
        CUDA_TRY(h, cudaMemsetAsync(h->gcount, 0, n * sizeof(int), h->stream));
    }
    const size_t n_i = (n + 1) & ~(size_t)1;                       // keep every section 8-byte aligned
    h->out_bytes = n * (nt + 2) * sizeof(double) + 3 * n_i * sizeof(int);
    CUDA_TRY(h, cudaMalloc(&h->out_d, h->out_bytes));
    CUDA_TRY(h, cudaMallocHost(&h->out_h, h->out_bytes));
    auto carve = [&](unsigned char* base, double*& g, double*& gn, double*& f, int*& it, int*& fg, int*& st) {
        g = reinterpret_cast<double*>(base);
        gn = g + n * nt;
        f = gn + n;
        it = reinterpret_cast<int*>(f + n);
        fg = it + n_i;
        st = fg + n_i;
    };
    carve(h->out_d, h->g_d, h->gnorm_d, h->f_d, h->iters_d, h->fg_d, h->status_d);
    double* f_h_unused;
    carve(h->out_h, h->g_h, h->gnorm_h, f_h_unused, h->iters_h, h->fg_h, h->status_h);
    h->out_cap = items;
    return 0;
}

// output blocks of the same layout outside the handle's main block (device-resident outer loop)
size_t muse_outblock_bytes(const muse_handle* h, int items) {
    const size_t n = (size_t)items, nt = (size_t)h->cfg.ntheta;
    const size_t n_i = (n + 1) & ~(size_t)1;
    return n * (nt + 2) * sizeof(double) + 3 * n_i * sizeof(int);
}

void muse_outblock_carve(const muse_handle* h, OutBlock& ob, unsigned char* dev, unsigned char* host, int items) {
    const size_t n = (size_t)items, nt = (size_t)h->cfg.ntheta;
    const size_t n_i = (n + 1) & ~(size_t)1;
    ob.d = dev; ob.hst = host; ob.bytes = muse_outblock_bytes(h, items); ob.cap = items; ob.owned = false;
    ob.g_d = reinterpret_cast<double*>(ob.d); ob.gnorm_d = ob.g_d + n * nt; ob.f_d = ob.gnorm_d + n;
    ob.iters_d = reinterpret_cast<int*>(ob.f_d + n); ob.fg_d = ob.iters_d + n_i; ob.status_d = ob.fg_d + n_i;
    ob.g_h = reinterpret_cast<double*>(ob.hst); ob.gnorm_h = ob.g_h + n * nt;
    ob.iters_h = reinterpret_cast<int*>(ob.gnorm_h + 2 * n); ob.fg_h = ob.iters_h + n_i; ob.status_h = ob.fg_h + n_i;
}

int muse_outblock_ensure(muse_handle* h, OutBlock& ob, int items) {
    if (items <= ob.cap) return 0;
    muse_outblock_free(ob);
    unsigned char *dev = nullptr, *host = nullptr;
    CUDA_TRY(h, cudaMalloc(&dev, muse_outblock_bytes(h, items)));
    CUDA_TRY(h, cudaMallocHost(&host, muse_outblock_bytes(h, items)));
    muse_outblock_carve(h, ob, dev, host, items);
    ob.owned = true;
    return 0;
}

void muse_outblock_free(OutBlock& ob) {
    if (ob.owned) { cudaFree(ob.d); cudaFreeHost(ob.hst); }
    ob = OutBlock{};
}

static void fill_common(muse_handle* h, SolveLaunch& L) {
    std::memset(&L, 0, sizeof(L));
    L.d = h->cfg.d;
    L.ld = h->ld;
    L.ntheta = h->cfg.ntheta;
    L.family = h->cfg.family;
    L.lbfgs_m = h->cfg.lbfgs_m;
    L.max_iters = h->cfg.max_iters;
    L.xi = h->xi;
    L.nu = h->nu;
    L.xdat = h->xdat;
    L.master_row = h->cfg.nsims;
    L.sbuf = h->sbuf;
    L.xslot = h->xslot;
    L.gpart = h->gpart;
    L.gcount = h->gcount;
    L.redo_count = h->redo_count;
    L.work_next = h->redo_count + 1;
    L.redo_items = h->redo_items;
    L.redo_total = h->redo_total;
    L.dxh = h->dxh;
    L.dgh = h->dgh;
    L.g_out = h->g_d;
    L.iters_out = h->iters_d;
    L.fg_out = h->fg_d;
    L.gnorm_out = h->gnorm_d;
    L.f_out = h->f_d;
    L.status_out = h->status_d;
    L.dbg = h->dbg;
}

int muse_ensure_outputs(muse_handle* h, int items) { return ensure_outputs(h, items); }
void muse_fill_common(muse_handle* h, SolveLaunch& L) { fill_common(h, L); }

static int launch_solver(muse_handle* h, const SolveLaunch& L, double bytes) {
    muse_handle::Rec r{};
    // inside a stream capture (device-resident loop, muse_outer.cu) the event records become external event-record nodes of
    // the graph, and the launch counters go to the graph's own tallies (added at every graph launch)
    const unsigned evflag = h->capturing ? cudaEventRecordExternal : cudaEventRecordDefault;
    int64_t& n_launch = h->capturing ? h->cap_launches : h->acc.launches;
    int64_t& n_solve = h->capturing ? h->cap_solve_launches : h->acc.solve_launches;
    if (h->prof) {
        CUDA_TRY(h, cudaEventCreate(&r.a));
        CUDA_TRY(h, cudaEventCreate(&r.b));
        CUDA_TRY(h, cudaEventRecordWithFlags(r.a, h->stream, evflag));
    }
    if (h->geo.stream) {
        // pass 1: single-pass speculative streaming kernel (its finisher warps replay the scalar optimiser on the
        // sums); pass 2: the generic kernel re-solves the units pass 1 handed back (device-side list; normally empty,
        // then its CTAs exit at once)
        SolveLaunch S = L;
        if (h->geo.stream != 1 || h->dbg_cap < h->geo.stream_grid) S.dbg = nullptr;   // the TMA-ring kernel stamps per CTA
        int* ctr = h->redo_count;
        if (h->ctr_override) {            // device-resident loop: a counter pair of its own per chain, zeroed with the state
            ctr = h->ctr_override;
            S.redo_count = ctr;
            S.work_next = ctr + 1;
        } else {
            CUDA_TRY(h, cudaMemsetAsync(h->redo_count, 0, 2 * sizeof(int), h->stream));
        }
        CUDA_TRY(h, launch_iso_stream(S, h->geo, h->stream));
        SolveLaunch R = L;
        R.dbg = nullptr;                  // the diagnostics buffer belongs to the streaming kernel's per-CTA rows
        R.item_list = h->redo_items;
        R.item_count = ctr;
        R.redo_count = ctr;
        R.work_next = ctr + 1;
        CUDA_TRY(h, launch_iso_solver(R, h->geo, h->stream));
        n_launch += 2;
    } else {
        CUDA_TRY(h, launch_iso_solver(L, h->geo, h->stream));
        n_launch += 1;
    }
    n_solve += 1;
    if (h->prof) {
        CUDA_TRY(h, cudaEventRecordWithFlags(r.b, h->stream, evflag));
        r.cls = 0;
        r.units = L.nitems;
        r.bytes = bytes;
        r.kind = h->pass_kind;
        r.tag = h->rec_tag;
        (h->capturing ? h->outer_recs : h->recs).push_back(r);
    }
    return 0;
}

// One solver pass over (data?) + sims [first_sim, first_sim + count) enqueued on the stream.  ob: where the per-unit outputs
// go (null: the handle's main block); dyn: θ-dependent constants in device memory (null: computed here from theta_*).
int muse_pass_enqueue(muse_handle* h, const double* theta_sim, const double* theta_eval, double atol, int include_data,
                      int warm_start, int first_sim, int count, const OutBlock* ob, const DynConsts* dyn) {
    const int items = count + (include_data ? 1 : 0);
    h->pass_kind = warm_start == MUSE_START_ZEROS ? MUSE_PASS_COLD : (warm_start == MUSE_START_TRUTH ? MUSE_PASS_TRUTH : MUSE_PASS_WARM);
    if (h->corr) return muse_corr_map_score(h, theta_sim, theta_eval, atol, include_data, warm_start, first_sim, count);
    SolveLaunch L;
    fill_common(h, L);
    L.nitems = items;
    L.mode = 0;
    L.include_data = include_data ? 1 : 0;
    L.first_sim = first_sim;
    L.atol = atol;
    if (dyn) L.dyn = dyn;
    else if (theta_consts(h->cfg, theta_sim, theta_eval, &L.smp[0], &L.ev) != 0) MUSE_FAIL(h, MUSE_EUNSUPPORTED, "family");
    switch (warm_start) {
        case MUSE_START_ZEROS: L.start_kind = kStartZero; break;
        case MUSE_START_PREV: L.start_kind = kStartOwn; break;
        case MUSE_START_TRUTH: L.start_kind = kStartTruth; break;
        default: L.start_kind = kStartSharedKeep; L.zshared = h->z0user; break;
    }
    L.zA = h->zA;
    L.zB = h->zB;
    L.zstate = h->zstate;
    if (ob) {
        L.g_out = ob->g_d; L.iters_out = ob->iters_d; L.fg_out = ob->fg_d;
        L.gnorm_out = ob->gnorm_d; L.f_out = ob->f_d; L.status_out = ob->status_d;
    }
    // algorithmic bytes: per sim read ξ, ν (16d) [+ z₀ 8d], write ẑ (8d); data unit reads x (8d)
    const double d8 = 8.0 * h->cfg.d;
    const double z0b = (warm_start == MUSE_START_PREV || warm_start == MUSE_START_USER) ? d8 : 0.0;
    const double bytes = count * (3 * d8 + z0b) + (include_data ? (2 * d8 + z0b) : 0.0);
    return launch_solver(h, L, bytes);
}

extern "C" {

int muse_b200_abi_version(void) { return MUSE_B200_ABI_VERSION; }

const char* muse_b200_last_error(const muse_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int muse_b200_create(const muse_cfg* cfg, muse_handle** out) {
    if (!cfg || !out) { g_create_err = "null argument"; return MUSE_EINVAL; }
    *out = nullptr;
    if (cfg->abi_version != MUSE_B200_ABI_VERSION) { g_create_err = "ABI version mismatch"; return MUSE_EINVAL; }
    if (cfg->family != MUSE_FAMILY_FUNNEL && cfg->family != MUSE_FAMILY_HIERGAUSS && cfg->family != MUSE_FAMILY_CORRGAUSS &&
        cfg->family != MUSE_FAMILY_TWOLAYER) {
        g_create_err = "model outside the registered families (funnel, hiergauss, corrgauss, twolayer); "
                       "Turing/Soss-defined models are not supported by the B200 backend";
        return MUSE_EUNSUPPORTED;
    }
    const int want_nt = cfg->family == MUSE_FAMILY_HIERGAUSS ? 2 : 1;
    if (cfg->ntheta != want_nt) { g_create_err = "ntheta does not match the family"; return MUSE_EINVAL; }
    if (cfg->d < 1 || cfg->nsims < 0 || cfg->nsims_h < 0) { g_create_err = "d must be ≥ 1 and nsims, nsims_h ≥ 0"; return MUSE_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_err = "no CUDA device: libmuse_b200 has no CPU fallback";
        return MUSE_ENODEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "bad device ordinal"; return MUSE_EINVAL; }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) {
        cudaGetLastError();
        g_create_err = "device is not compute capability 10.x (kernels are built for sm_100a only)";
        return MUSE_ENODEVICE;
    }
    muse_handle* h = new (std::nothrow) muse_handle();
    if (!h) { g_create_err = "out of host memory"; return MUSE_ENOMEM; }
    h->cfg = *cfg;
    if (h->cfg.lbfgs_m <= 0) h->cfg.lbfgs_m = 10;
    if (h->cfg.lbfgs_m > 16) h->cfg.lbfgs_m = 16;
    if (h->cfg.max_iters <= 0) h->cfg.max_iters = 1000;
    h->ld = round_up(cfg->d, 32);
    h->rows = cfg->nsims + 1;

    auto fail = [&](int code) {
        g_create_err = h->err;
        muse_b200_destroy(h);
        return code;
    };
#define CREATE_TRY(expr)                                                      \
    do {                                                                      \
        cudaError_t e__ = (expr);                                             \
        if (e__ != cudaSuccess) {                                             \
            h->err = std::string(#expr) + ": " + cudaGetErrorString(e__);     \
            return fail(e__ == cudaErrorMemoryAllocation ? MUSE_ENOMEM : MUSE_ECUDA); \
        }                                                                     \
    } while (0)

    CREATE_TRY(cudaSetDevice(cfg->device));
    if (cfg->stream) {
        h->stream = (cudaStream_t)cfg->stream;
    } else {
        CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    if (cfg->family == MUSE_FAMILY_CORRGAUSS || cfg->family == MUSE_FAMILY_TWOLAYER) {
        // anisotropic quadratic families (dense correlated Gaussian, two-layer hierarchy): the lock-step solver with its own state
        // (muse_corr.cu); the isotropic kernels are not involved
        CREATE_TRY(cudaMalloc(&h->redo_total, sizeof(unsigned long long)));
        CREATE_TRY(cudaMemsetAsync(h->redo_total, 0, sizeof(unsigned long long), h->stream));
        const int rc = muse_corr_create(h);
        if (rc != MUSE_OK) return fail(rc);
        if (ensure_outputs(h, h->rows) != 0) return fail(MUSE_ECUDA);
        CREATE_TRY(cudaStreamSynchronize(h->stream));
        *out = h;
        return MUSE_OK;
    }
    {
        // kernel choice: a single-pass speculative kernel runs first — the TMA-ring streaming kernel
        // for d ≥ 4096, its warp-per-unit form below that — and the generic solver (group shape by d: one warp per
        // unit for small d, one CTA otherwise) only re-solves what the first pass hands back.
        // cfg.kernel: 0 auto, 1 generic only, 2 TMA-ring streaming first (any d), 3 warp-per-unit streaming first (any d).
        if (cfg->kernel < 0 || cfg->kernel > 3) { h->err = "cfg.kernel must be 0..3"; return fail(MUSE_EINVAL); }
        CREATE_TRY(iso_solver_geometry(cfg->d, cfg->group, cfg->cluster, cfg->device, &h->geo));
        h->geo.stream = 0;
        if (cfg->kernel == 2 || (cfg->kernel == 0 && cfg->d >= 4096)) CREATE_TRY(iso_stream_geometry(cfg->d, h->ld, cfg->device, &h->geo));
        else if (cfg->kernel == 3 || cfg->kernel == 0) CREATE_TRY(iso_warp_stream_geometry(cfg->device, &h->geo));
    }

    const size_t ld = (size_t)h->ld, rows = (size_t)h->rows, B = sizeof(double);
    CREATE_TRY(cudaMalloc(&h->xi, rows * ld * B));
    CREATE_TRY(cudaMalloc(&h->nu, rows * ld * B));
    if (cfg->nsims_h > 0) {
        CREATE_TRY(cudaMalloc(&h->xi_h, (size_t)cfg->nsims_h * ld * B));
        CREATE_TRY(cudaMalloc(&h->nu_h, (size_t)cfg->nsims_h * ld * B));
        CREATE_TRY(cudaMemsetAsync(h->xi_h, 0, (size_t)cfg->nsims_h * ld * B, h->stream));
        CREATE_TRY(cudaMemsetAsync(h->nu_h, 0, (size_t)cfg->nsims_h * ld * B, h->stream));
    }
    CREATE_TRY(cudaMalloc(&h->xdat, ld * B));
    CREATE_TRY(cudaMalloc(&h->z0user, ld * B));
    CREATE_TRY(cudaMalloc(&h->zA, rows * ld * B));
    CREATE_TRY(cudaMalloc(&h->zB, rows * ld * B));
    CREATE_TRY(cudaMalloc(&h->zstate, rows * sizeof(int)));
    const size_t slots = (size_t)h->geo.groups, m = (size_t)h->cfg.lbfgs_m;
    CREATE_TRY(cudaMalloc(&h->xslot, slots * ld * B));
    CREATE_TRY(cudaMalloc(&h->sbuf, slots * ld * B));
    CREATE_TRY(cudaMalloc(&h->dxh, slots * m * ld * B));
    CREATE_TRY(cudaMalloc(&h->dgh, slots * m * ld * B));
    CREATE_TRY(cudaMalloc(&h->zfidA, ld * B));
    CREATE_TRY(cudaMalloc(&h->zfidB, ld * B));
    CREATE_TRY(cudaMalloc(&h->zfid_state, sizeof(int)));
    CREATE_TRY(cudaMalloc(&h->redo_count, 2 * sizeof(int)));      // [hand-back count, streaming work counter]
    CREATE_TRY(cudaMemsetAsync(h->redo_count, 0, 2 * sizeof(int), h->stream));
    CREATE_TRY(cudaMalloc(&h->redo_total, sizeof(unsigned long long)));
    CREATE_TRY(cudaMemsetAsync(h->redo_total, 0, sizeof(unsigned long long), h->stream));
    CREATE_TRY(cudaMemsetAsync(h->xi, 0, rows * ld * B, h->stream));
    CREATE_TRY(cudaMemsetAsync(h->nu, 0, rows * ld * B, h->stream));
    CREATE_TRY(cudaMemsetAsync(h->xdat, 0, ld * B, h->stream));
    CREATE_TRY(cudaMemsetAsync(h->z0user, 0, ld * B, h->stream));
    CREATE_TRY(cudaMemsetAsync(h->zstate, 0, rows * sizeof(int), h->stream));
    CREATE_TRY(cudaMemsetAsync(h->zfid_state, 0, sizeof(int), h->stream));
    if (ensure_outputs(h, h->rows) != 0) return fail(MUSE_ECUDA);
    CREATE_TRY(cudaStreamSynchronize(h->stream));
#undef CREATE_TRY
    *out = h;
    return MUSE_OK;
}

int muse_b200_destroy(muse_handle* h) {
    if (!h) return MUSE_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& r : h->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    muse_p2p_release(h);
    muse_comm_release(h);
    muse_outer_release(h);
    muse_corr_destroy(h);
    cudaFree(h->xi); cudaFree(h->nu); cudaFree(h->xi_h); cudaFree(h->nu_h); cudaFree(h->xdat); cudaFree(h->z0user);
    cudaFree(h->xslot); cudaFree(h->zA); cudaFree(h->zB); cudaFree(h->zstate);
    cudaFree(h->sbuf); cudaFree(h->dxh); cudaFree(h->dgh);
    cudaFree(h->out_d); cudaFreeHost(h->out_h);
    cudaFree(h->zHA); cudaFree(h->zHB);
    cudaFree(h->dbg);
    cudaFree(h->gpart); cudaFree(h->gcount); cudaFree(h->redo_count); cudaFree(h->redo_items); cudaFree(h->redo_total);
    cudaFree(h->zfidA); cudaFree(h->zfidB); cudaFree(h->zfid_state);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return MUSE_OK;
}

int muse_b200_set_stream(muse_handle* h, void* s) {
    if (!h) return MUSE_EINVAL;
    if (s && h->stream == (cudaStream_t)s) return MUSE_OK;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    if (s) {
        h->stream = (cudaStream_t)s;
    } else {
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    return MUSE_OK;
}

int muse_b200_set_data(muse_handle* h, const double* x_dat) {
    if (!h || !x_dat) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->corr) {
        const int rc = muse_corr_set_data(h, x_dat);
        if (rc == MUSE_OK) h->have_data = true;
        return rc;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->xdat, x_dat, (size_t)h->cfg.d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_data = true;
    return MUSE_OK;
}

int muse_b200_set_z0(muse_handle* h, const double* z0) {
    if (!h || !z0) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->corr) {
        const int rc = muse_corr_set_z0(h, z0);
        if (rc == MUSE_OK) h->have_z0 = true;
        return rc;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->z0user, z0, (size_t)h->cfg.d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_z0 = true;
    return MUSE_OK;
}

int muse_b200_set_draws(muse_handle* h, const double* xi, const double* nu, const double* xi_m, const double* nu_m) {
    if (!h || !xi_m || !nu_m || (h->cfg.nsims > 0 && (!xi || !nu))) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->corr) {
        const int rc = muse_corr_set_draws(h, xi, nu, xi_m, nu_m, false);
        if (rc == MUSE_OK) h->have_draws = true;
        return rc;
    }
    const size_t w = (size_t)h->cfg.d * sizeof(double), pitch = (size_t)h->ld * sizeof(double);
    const size_t n = (size_t)h->cfg.nsims;
    if (n) {
        CUDA_TRY(h, cudaMemcpy2DAsync(h->xi, pitch, xi, w, w, n, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaMemcpy2DAsync(h->nu, pitch, nu, w, w, n, cudaMemcpyHostToDevice, h->stream));
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->xi + n * h->ld, xi_m, w, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->nu + n * h->ld, nu_m, w, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_draws = true;
    return MUSE_OK;
}

int muse_b200_seed_draws(muse_handle* h, uint64_t seed) {
    if (!h) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->corr) {
        const int rc = muse_corr_seed_draws(h, seed);
        if (rc == MUSE_OK) { h->have_draws = true; h->have_draws_h = h->cfg.nsims_h > 0; }
        return rc;
    }
    muse_handle::Rec r{};
    if (h->prof) {
        CUDA_TRY(h, cudaEventCreate(&r.a));
        CUDA_TRY(h, cudaEventCreate(&r.b));
        CUDA_TRY(h, cudaEventRecord(r.a, h->stream));
    }
    CUDA_TRY(h, launch_philox_draws(h->xi, h->nu, h->cfg.nsims + 1, h->cfg.d, h->ld, seed, h->cfg.sim_offset,
                                    h->cfg.nsims, h->stream));
    h->acc.launches += 1;
    h->acc.draw_launches += 1;
    if (h->cfg.nsims_h > 0) {
        CUDA_TRY(h, launch_philox_draws(h->xi_h, h->nu_h, h->cfg.nsims_h, h->cfg.d, h->ld, seed, h->cfg.h_sim_offset,
                                        -1, h->stream));
        h->acc.launches += 1;
        h->acc.draw_launches += 1;
        h->have_draws_h = true;
    }
    if (h->prof) {
        CUDA_TRY(h, cudaEventRecord(r.b, h->stream));
        r.cls = 1;
        h->recs.push_back(r);
    }
    h->have_draws = true;       // no synchronisation: later launches on the stream are ordered behind the generator
    return MUSE_OK;
}

int muse_b200_set_draws_h(muse_handle* h, const double* xi_h, const double* nu_h) {
    if (!h || !xi_h || !nu_h) return MUSE_EINVAL;
    if (h->cfg.nsims_h <= 0) MUSE_FAIL(h, MUSE_ESTATE, "handle was created without a separate H shard (nsims_h = 0)");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->corr) {
        const int rc = muse_corr_set_draws(h, xi_h, nu_h, nullptr, nullptr, true);
        if (rc == MUSE_OK) h->have_draws_h = true;
        return rc;
    }
    const size_t w = (size_t)h->cfg.d * sizeof(double), pitch = (size_t)h->ld * sizeof(double);
    CUDA_TRY(h, cudaMemcpy2DAsync(h->xi_h, pitch, xi_h, w, w, (size_t)h->cfg.nsims_h, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(h->nu_h, pitch, nu_h, w, w, (size_t)h->cfg.nsims_h, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_draws_h = true;
    return MUSE_OK;
}

int muse_b200_get_draws(muse_handle* h, int32_t first, int32_t count, double* xi_out, double* nu_out) {
    if (!h || first < 0 || count < 0 || first + count > h->cfg.nsims + 1) return MUSE_EINVAL;
    if (!h->have_draws) MUSE_FAIL(h, MUSE_ESTATE, "no draws installed (set_draws / seed_draws)");
    if (h->corr) MUSE_FAIL(h, MUSE_EUNSUPPORTED, "corrgauss keeps L·ξ, not ξ: get_draws is not available");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t w = (size_t)h->cfg.d * sizeof(double), pitch = (size_t)h->ld * sizeof(double);
    if (count && xi_out)
        CUDA_TRY(h, cudaMemcpy2DAsync(xi_out, w, h->xi + (size_t)first * h->ld, pitch, w, count, cudaMemcpyDeviceToHost, h->stream));
    if (count && nu_out)
        CUDA_TRY(h, cudaMemcpy2DAsync(nu_out, w, h->nu + (size_t)first * h->ld, pitch, w, count, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MUSE_OK;
}

int muse_b200_map_score_async(muse_handle* h, const double* theta_sim, const double* theta_eval, double atol,
                              int32_t include_data, int32_t warm_start, int32_t first_sim, int32_t count) {
    if (!h || !theta_sim || !theta_eval) return MUSE_EINVAL;
    if (first_sim < 0 || count < 0 || first_sim + count > h->cfg.nsims) MUSE_FAIL(h, MUSE_EINVAL, "sim range outside the handle's shard");
    if (warm_start < MUSE_START_ZEROS || warm_start > MUSE_START_USER) MUSE_FAIL(h, MUSE_EINVAL, "bad warm_start");
    if (include_data && !h->have_data) MUSE_FAIL(h, MUSE_ESTATE, "observed data not set (muse_b200_set_data)");
    if (count > 0 && !h->have_draws) MUSE_FAIL(h, MUSE_ESTATE, "no draws installed (set_draws / seed_draws)");
    if (warm_start == MUSE_START_USER && !h->have_z0) MUSE_FAIL(h, MUSE_ESTATE, "user z0 not set (muse_b200_set_z0)");
    const int items = count + (include_data ? 1 : 0);
    if (items == 0) return MUSE_OK;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    return muse_pass_enqueue(h, theta_sim, theta_eval, atol, include_data, warm_start, first_sim, count, nullptr, nullptr);
}

int muse_b200_fetch(muse_handle* h, int32_t units, double* g_out, int32_t* iters_out, int32_t* fg_out,
                    double* gnorm_out, int32_t* status_out) {
    if (!h || units < 0 || units > h->out_cap) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)units;
    if (n) CUDA_TRY(h, cudaMemcpyAsync(h->out_h, h->out_d, h->out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (n) {
        if (g_out && g_out != h->g_h) std::memcpy(g_out, h->g_h, n * h->cfg.ntheta * sizeof(double));
        if (iters_out && iters_out != h->iters_h) std::memcpy(iters_out, h->iters_h, n * sizeof(int));
        if (fg_out && fg_out != h->fg_h) std::memcpy(fg_out, h->fg_h, n * sizeof(int));
        if (gnorm_out && gnorm_out != h->gnorm_h) std::memcpy(gnorm_out, h->gnorm_h, n * sizeof(double));
        if (status_out && status_out != h->status_h) std::memcpy(status_out, h->status_h, n * sizeof(int));
    }
    return MUSE_OK;
}

int muse_b200_map_score(muse_handle* h, const double* theta_sim, const double* theta_eval, double atol,
                        int32_t include_data, int32_t warm_start, int32_t first_sim, int32_t count, double* g_out,
                        int32_t* iters_out, int32_t* fg_out, double* gnorm_out, int32_t* status_out) {
    const int rc = muse_b200_map_score_async(h, theta_sim, theta_eval, atol, include_data, warm_start, first_sim, count);
    if (rc != MUSE_OK) return rc;
    return muse_b200_fetch(h, count + (include_data ? 1 : 0), g_out, iters_out, fg_out, gnorm_out, status_out);
}

int muse_b200_device_scores(muse_handle* h, double** g_dev, int32_t* capacity_units) {
    if (!h || !g_dev) return MUSE_EINVAL;
    *g_dev = h->g_d;
    if (capacity_units) *capacity_units = h->out_cap;
    return MUSE_OK;
}

// fetch the ± scores and form central_fdm(3,1): sum(fs .* [-1/2, 0, 1/2]) / step   — src/util.jl:13-19
void muse_fd_combine_host(muse_handle* h, const double* g_h, const int* status_h, const double* step, int nsims_H,
                          double* Hs_out, int32_t* status_out) {
    const int nt = h->cfg.ntheta, items = nsims_H * nt * 2;
    if (status_out) std::memcpy(status_out, status_h, (size_t)items * sizeof(int));
    for (int k = 0; k < nsims_H; ++k)
        for (int n = 0; n < nt; ++n) {
            const double* gm = g_h + ((size_t)(k * nt + n) * 2 + 0) * nt;
            const double* gp = g_h + ((size_t)(k * nt + n) * 2 + 1) * nt;
            for (int i = 0; i < nt; ++i) {
                double acc = gm[i] * -0.5;
                acc = acc + 0.0;
                acc = acc + gp[i] * 0.5;
                Hs_out[((size_t)k * nt + i) * nt + n] = acc / step[n];
            }
        }
}

static int fd_combine(muse_handle* h, const double* step, int nsims_H, double* Hs_out, int32_t* status_out) {
    const int nt = h->cfg.ntheta, items = nsims_H * nt * 2;
    int rc = muse_b200_fetch(h, items, h->g_h, nullptr, nullptr, nullptr, status_out ? h->status_h : nullptr);
    if (rc != 0) return rc;
    muse_fd_combine_host(h, h->g_h, h->status_h, step, nsims_H, Hs_out, status_out);
    return MUSE_OK;
}


// Shared launch sequence of muse_b200_fd_jacobian / muse_b200_fd_scores: the fiducial solve and the 2·nθ virtual sims per
// H sim, sampled at the rows of th_pts (row 2n = the "−" point of column n, row 2n+1 its "+" point), MAP + score at theta0.
static int fd_launch(muse_handle* h, const double* theta0, const double* th_pts, int nsims_H, double atol,
                     const DynConsts* dyn_fid = nullptr, const DynConsts* dyn_fd = nullptr, const OutBlock* ob = nullptr) {
    const bool hshard = h->cfg.nsims_h > 0;
    const int nt = h->cfg.ntheta;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int items = nsims_H * nt * 2;
    const size_t ld = (size_t)h->ld, B = sizeof(double);
    if (h->corr) {
        if (ensure_outputs(h, items) != 0) return MUSE_ECUDA;
        return muse_corr_fd_launch(h, theta0, th_pts, nsims_H, atol);
    }
    if (items > h->h_cap) {
        cudaFree(h->zHA); cudaFree(h->zHB);
        h->zHA = h->zHB = nullptr;
        h->h_cap = 0;
        CUDA_TRY(h, cudaMalloc(&h->zHA, (size_t)items * ld * B));
        CUDA_TRY(h, cudaMalloc(&h->zHB, (size_t)items * ld * B));
        h->h_cap = items;
    }
    if (ensure_outputs(h, items) != 0) return MUSE_ECUDA;
    const double d8 = 8.0 * h->cfg.d;

    // (1) fiducial MAP of the master stream's draw from zero(z)   — src/muse.jl:417-423
    SolveLaunch F;
    fill_common(h, F);
    F.nitems = 1;
    F.mode = 2;
    F.atol = atol;
    // start of the fiducial solve: zero(z) (ẑ_guess_from_truth's default) or the user's z₀ (`z₀` keyword of get_H!, src/muse.jl:309, 419)
    F.start_kind = h->fd_start_user ? kStartSharedKeep : kStartZero;
    F.zshared = h->z0user;
    if (dyn_fid) F.dyn = dyn_fid;
    else if (theta_consts(h->cfg, theta0, theta0, &F.smp[0], &F.ev) != 0) MUSE_FAIL(h, MUSE_EUNSUPPORTED, "family");
    F.zA = h->zfidA;
    F.zB = h->zfidB;
    int* zst = h->zfid_override ? h->zfid_override : h->zfid_state;     // override: already zero (uploaded with the loop's state)
    F.zstate = zst;
    if (!h->zfid_override) CUDA_TRY(h, cudaMemsetAsync(h->zfid_state, 0, sizeof(int), h->stream));
    auto outputs_to = [&](SolveLaunch& X) {
        if (!ob) return;
        X.g_out = ob->g_d; X.iters_out = ob->iters_d; X.fg_out = ob->fg_d;
        X.gnorm_out = ob->gnorm_d; X.f_out = ob->f_d; X.status_out = ob->status_d;
    };
    outputs_to(F);
    h->pass_kind = MUSE_PASS_FIDUCIAL;
    int* const ctr0 = h->ctr_override;
    int rc = launch_solver(h, F, 3 * d8);
    if (rc != 0) return rc;
    if (ctr0) h->ctr_override = ctr0 + 2;          // the virtual sims' chain takes the next counter pair
    // (2) virtual sims at the 2·nθ sample points, MAP + score at θ₀ from the fiducial start — src/muse.jl:426-433
    SolveLaunch L;
    fill_common(h, L);
    L.nitems = items;
    L.mode = 1;
    L.atol = atol;
    L.start_kind = kStartShared;
    L.zshared = nullptr;
    L.zshared_state = zst;               // picked on the device: no host sync between the two launches
    outputs_to(L);
    L.zsharedA = h->zfidA;
    L.zsharedB = h->zfidB;
    if (hshard) { L.xi = h->xi_h; L.nu = h->nu_h; }
    if (dyn_fd) L.dyn = dyn_fd;
    else {
        if (theta_consts(h->cfg, theta0, theta0, nullptr, &L.ev) != 0) MUSE_FAIL(h, MUSE_EUNSUPPORTED, "family");
        for (int p = 0; p < 2 * nt; ++p) theta_consts(h->cfg, th_pts + (size_t)p * nt, theta0, &L.smp[p], nullptr);
    }
    L.zA = h->zHA;
    L.zB = h->zHB;
    L.zstate = nullptr;
    L.discard_z = 1;     // `ẑ, = ẑ_at_θ(...)` is only an intermediate of the score (src/muse.jl:431-432)
    // algorithmic bytes: read ξ, ν per virtual sim; the shared start ẑ_fid is read once (L2)
    h->pass_kind = MUSE_PASS_FD;
    rc = launch_solver(h, L, items * 2 * d8 + d8);
    h->ctr_override = ctr0;
    return rc;
}

int muse_fd_enqueue(muse_handle* h, const double* theta0, const double* th_pts, int nsims_H, double atol,
                    const DynConsts* dyn_fid, const DynConsts* dyn_fd, const OutBlock* ob) {
    return fd_launch(h, theta0, th_pts, nsims_H, atol, dyn_fid, dyn_fd, ob);
}

static int fd_check(muse_handle* h, int nsims_H) {
    const bool hshard = h->cfg.nsims_h > 0;
    if (nsims_H < 0 || nsims_H > (hshard ? h->cfg.nsims_h : h->cfg.nsims)) MUSE_FAIL(h, MUSE_EINVAL, "nsims_H outside the handle's H shard");
    if (!h->have_draws || (hshard && !h->have_draws_h)) MUSE_FAIL(h, MUSE_ESTATE, "no draws installed (set_draws[_h] / seed_draws)");
    return MUSE_OK;
}

int muse_b200_fd_start(muse_handle* h, int32_t start) {
    if (!h || (start != MUSE_START_ZEROS && start != MUSE_START_USER)) return MUSE_EINVAL;
    if (start == MUSE_START_USER && !h->have_z0) MUSE_FAIL(h, MUSE_ESTATE, "user z0 not set (muse_b200_set_z0)");
    if (start == MUSE_START_USER && h->corr) MUSE_FAIL(h, MUSE_EUNSUPPORTED, "corrgauss: the fiducial solve of get_H! starts from zero(z)");
    h->fd_start_user = start == MUSE_START_USER;
    return MUSE_OK;
}

int muse_b200_fd_jacobian(muse_handle* h, const double* theta0, const double* step, int32_t nsims_H, double atol,
                          double* Hs_out, int32_t* status_out) {
    if (!h || !theta0 || !step || !Hs_out) return MUSE_EINVAL;
    int rc = fd_check(h, nsims_H);
    if (rc != MUSE_OK) return rc;
    const int nt = h->cfg.ntheta;
    for (int n = 0; n < nt; ++n)
        if (!(step[n] != 0.0) || !std::isfinite(step[n])) MUSE_FAIL(h, MUSE_EINVAL, "finite-difference step must be finite and non-zero");
    if (nsims_H == 0) return MUSE_OK;
    double pts[2 * kMaxTheta * kMaxTheta];
    for (int n = 0; n < nt; ++n)
        for (int s = 0; s < 2; ++s) {
            double* th = pts + (size_t)(2 * n + s) * nt;
            for (int i = 0; i < nt; ++i) th[i] = theta0[i];
            const double eps = 0.0 + step[n] * (s ? 1.0 : -1.0);     // x .+ step .* grid  [EXT FiniteDifferences]
            th[n] = theta0[n] + eps;                                 // src/util.jl:15
        }
    rc = fd_launch(h, theta0, pts, nsims_H, atol);
    if (rc != MUSE_OK) return rc;
    return fd_combine(h, step, nsims_H, Hs_out, status_out);
}

int muse_b200_fd_scores(muse_handle* h, const double* theta_eval, const double* theta_sims, int32_t nsims_H, double atol,
                        double* g_out, int32_t* status_out) {
    if (!h || !theta_eval || !theta_sims || !g_out) return MUSE_EINVAL;
    int rc = fd_check(h, nsims_H);
    if (rc != MUSE_OK) return rc;
    const int nt = h->cfg.ntheta;
    for (int i = 0; i < 2 * nt * nt; ++i)
        if (!std::isfinite(theta_sims[i])) MUSE_FAIL(h, MUSE_EINVAL, "finite-difference sample points must be finite");
    if (nsims_H == 0) return MUSE_OK;
    rc = fd_launch(h, theta_eval, theta_sims, nsims_H, atol);
    if (rc != MUSE_OK) return rc;
    return muse_b200_fetch(h, nsims_H * nt * 2, g_out, nullptr, nullptr, nullptr, status_out);
}

int muse_b200_get_maps(muse_handle* h, int32_t first_unit, int32_t count, double* z_out) {
    if (!h || !z_out || first_unit < 0 || count < 0 || first_unit + count > h->rows) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->corr) return muse_corr_get_maps(h, first_unit, count, z_out);
    std::vector<int> st((size_t)count);
    if (count) CUDA_TRY(h, cudaMemcpyAsync(st.data(), h->zstate + first_unit, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const size_t w = (size_t)h->cfg.d * sizeof(double);
    for (int i = 0; i < count; ++i) {
        double* dst = z_out + (size_t)i * h->cfg.d;
        const size_t off = (size_t)(first_unit + i) * h->ld;
        if (st[i] == kZA) CUDA_TRY(h, cudaMemcpyAsync(dst, h->zA + off, w, cudaMemcpyDeviceToHost, h->stream));
        else if (st[i] == kZB) CUDA_TRY(h, cudaMemcpyAsync(dst, h->zB + off, w, cudaMemcpyDeviceToHost, h->stream));
        else std::memset(dst, 0, w);
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MUSE_OK;
}

int muse_b200_profile_reset(muse_handle* h, int32_t enable) {
    if (!h) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->persist_pend = false;
    for (auto& r : h->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    h->recs.clear();
    h->acc = muse_profile{};
    h->acc_pass = muse_pass_profile{};
    h->prof = enable != 0;
    return MUSE_OK;
}

// fold the finished event pairs into the accumulators (stream must be idle)
static int profile_drain(muse_handle* h) {
    muse_persist_flush(h);
    for (auto& r : h->recs) {
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.cls == 0) {
            h->acc.solve_ms += ms; h->acc.solve_units += r.units; h->acc.solve_bytes += r.bytes;
            const int k = r.kind >= 0 && r.kind < MUSE_PASS_KINDS ? r.kind : MUSE_PASS_COLD;
            h->acc_pass.launches[k] += 1; h->acc_pass.ms[k] += ms; h->acc_pass.units[k] += r.units; h->acc_pass.bytes[k] += r.bytes;
        }
        else if (r.cls == 3) { h->acc.solve_ms += ms; h->acc.solve_units += r.units; h->acc.solve_bytes += r.bytes; }   // a whole solve in one launch
        else if (r.cls == 1) h->acc.draw_ms += ms;
        else { h->acc.other_ms += ms; }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->recs.clear();
    return MUSE_OK;
}

int muse_b200_profile_get(muse_handle* h, muse_profile* out) {
    if (!h || !out) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const int rc = profile_drain(h);
    if (rc != MUSE_OK) return rc;
    unsigned long long redo = 0;
    CUDA_TRY(h, cudaMemcpy(&redo, h->redo_total, sizeof(redo), cudaMemcpyDeviceToHost));
    h->acc.redo_units = (int64_t)redo;
    *out = h->acc;
    return MUSE_OK;
}

int muse_b200_profile_passes(muse_handle* h, muse_pass_profile* out) {
    if (!h || !out) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const int rc = profile_drain(h);
    if (rc != MUSE_OK) return rc;
    *out = h->acc_pass;
    return MUSE_OK;
}

int muse_b200_debug_timeline(muse_handle* h, int32_t items, int64_t* out) {
    if (!h || items < 0) return MUSE_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (!out) {   // (re)arm: allocate room for `items` units; items = 0 disarms
        cudaFree(h->dbg);
        h->dbg = nullptr;
        h->dbg_cap = 0;
        if (items > 0) {
            CUDA_TRY(h, cudaMalloc(&h->dbg, (size_t)items * 16 * sizeof(long long)));
            CUDA_TRY(h, cudaMemset(h->dbg, 0, (size_t)items * 16 * sizeof(long long)));
            h->dbg_cap = items;
        }
        return MUSE_OK;
    }
    if (items > h->dbg_cap) return MUSE_EINVAL;
    CUDA_TRY(h, cudaMemcpy(out, h->dbg, (size_t)items * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return MUSE_OK;
}

int muse_b200_geometry(muse_handle* h, int32_t* group_threads, int32_t* cluster, int32_t* groups) {
    if (!h) return MUSE_EINVAL;
    if (group_threads) *group_threads = h->geo.group_threads;
    if (cluster) *cluster = h->geo.cluster;
    if (groups) *groups = h->geo.groups;
    return MUSE_OK;
}

}  // extern "C"
