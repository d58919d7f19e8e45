/* muse_b200.h — C ABI of libmuse_b200.so, the B200 (sm_100a) backend for the per-simulation
 * hot path of MUSE.
 *
 * The reference (MuseInference.jl, pure Julia) has no FFI: its extension seam is multiple
 * dispatch on `AbstractMuseProblem` (/root/reference/src/interface.jl:4) and the three mapped
 * blocks `pmap(pool, ...) do ... end` of
 *     muse!   /root/reference/src/muse.jl:169-176
 *     get_H!  /root/reference/src/muse.jl:417-423 and 426-442 (+ pjacobian, src/util.jl:9-26)
 *     get_J!  /root/reference/src/muse.jl:508-525
 * A batched GPU backend replaces exactly those blocks.  Each entry point below cites the
 * reference lines it replaces; INTEGRATION.md shows the Julia `ccall` stubs and the Python
 * ctypes binding (museinference.jl_b200/_capi.py) that binds them.
 *
 * Conventions
 *   - all functions return 0 on success or a negative MUSE_E* code; the message is available
 *     from muse_b200_last_error(); no C++ exception crosses this boundary;
 *   - all numeric data is FP64 (the reference computes in Float64, src/muse.jl:152);
 *   - host matrices are dense row-major, "sim-major": row k = simulation k, d contiguous values;
 *   - the caller owns every host pointer (valid for the duration of the call only); the library
 *     owns all device memory until muse_b200_destroy();
 *   - calls are synchronous unless the name ends in _async; a handle is not thread-safe;
 *   - one handle drives one GPU and owns the contiguous shard [sim_offset, sim_offset+nsims)
 *     of the global simulation index space (one process per GPU; the tiny cross-rank exchange
 *     of scores is done by the host with NCCL);
 *   - there is NO CPU fallback: without a CUDA device create() fails with MUSE_ENODEVICE.
 *
 * Unit numbering inside a handle: unit 0 is the observed data (`rng === nothing`,
 * src/muse.jl:170); unit 1+k is local simulation k (global index sim_offset+k).
 */
#ifndef MUSE_B200_H
#define MUSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUSE_B200_ABI_VERSION 1

/* error codes */
#define MUSE_OK            0
#define MUSE_EINVAL       -1   /* bad argument */
#define MUSE_ENODEVICE    -2   /* no CUDA device / wrong architecture */
#define MUSE_ECUDA        -3   /* CUDA runtime error (see last_error) */
#define MUSE_ENOMEM       -4
#define MUSE_EUNSUPPORTED -5   /* model outside the registered families */
#define MUSE_ESTATE       -6   /* call order violated (e.g. no draws installed) */

/* registered model families */
#define MUSE_FAMILY_FUNNEL     1   /* src/simple.jl:58-76: z~N(0,e^θ I), x~N(z,I); scalar θ            */
#define MUSE_FAMILY_HIERGAUSS  2   /* z~N(μ,e^{2ℓ} I), x~N(z,I); θ=(μ,ℓ)                               */
#define MUSE_FAMILY_CORRGAUSS  3   /* z~N(0,e^θ Σ₀), x~N(z,I); scalar θ; P=Σ₀⁻¹, L=chol Σ₀ supplied    */
#define MUSE_FAMILY_TWOLAYER   4   /* src/turing.jl:63-79 (docstring toy): z~N(0,e^{σ/2}I_n), w~N(z,I), x~N(w,I), y~N(x,I); scalar σ;
                                      d = 2n even: latent (z,w), data (x,y), latent normals (ξ_z,ξ_w), noise normals (ν_x,ν_y) stacked */

/* warm-start modes for the latent MAP (what the reference passes as z₀ to ẑ_at_θ) */
#define MUSE_START_ZEROS   0   /* zero(z)                 src/muse.jl:151, src/interface.jl:184-186 */
#define MUSE_START_PREV    1   /* previous ẑ of the unit  src/muse.jl:169,181 (ẑs carried over)     */
#define MUSE_START_TRUTH   2   /* the simulated z         src/muse.jl:511                           */
#define MUSE_START_USER    3   /* user z₀ (muse_b200_set_z0), the `z₀` keyword                     */

/* per-unit solver status (Optim's convergence flags, src/interface.jl:168-171) */
#define MUSE_STATUS_G_CONVERGED   0   /* ‖∇z‖_∞ ≤ atol                                   */
#define MUSE_STATUS_XF_CONVERGED  1   /* x or f stopped changing exactly (x_tol=f_tol=0) */
#define MUSE_STATUS_MAXITER       2   /* 1000 iterations: Optim.converged == false       */
#define MUSE_STATUS_LS_FAILED     3   /* line search exception: optimisation stopped     */
#define MUSE_STATUS_NONFINITE     4   /* non-finite objective / gradient                 */

#define MUSE_MAX_NTHETA 8

typedef struct muse_handle muse_handle;

typedef struct muse_cfg {
    int32_t abi_version;     /* MUSE_B200_ABI_VERSION */
    int32_t family;          /* MUSE_FAMILY_* */
    int32_t d;               /* latent (= data) dimension */
    int32_t ntheta;          /* 1 (funnel, corrgauss, twolayer) or 2 (hiergauss) */
    int32_t nsims;           /* local simulations owned by this handle (rows of the draw arrays) */
    int32_t device;          /* CUDA device ordinal */
    int64_t sim_offset;      /* global index of local simulation 0 */
    int32_t nsims_h;         /* sims of the get_H! shard held as extra draw rows; 0 → the H shard is the
                                first sims of the local shard (single-GPU default) */
    int32_t kernel;          /* solver kernels: 0 auto (single-pass kernel first: TMA-ring streaming for d ≥ 4096, warp per
                                unit below), 1 generic two-sweep solver only, 2 TMA-ring streaming first, 3 warp-per-unit
                                streaming first; the generic kernel re-solves what the first pass hands back */
    int64_t h_sim_offset;    /* global index of H-shard simulation 0 (used when nsims_h > 0) */
    int32_t lbfgs_m;         /* L-BFGS memory; 0 → 10 (Optim.LBFGS default) */
    int32_t max_iters;       /* 0 → 1000 (Optim.Options default) */
    int32_t group;           /* threads cooperating on one MAP solve: 0 auto, 32 = warp, else CTA size */
    int32_t cluster;         /* CTAs per thread-block cluster cooperating on one solve: 0 auto, 1,2,4,8 */
    const double* P;         /* corrgauss: Σ₀⁻¹, d×d row-major host pointer; else NULL */
    const double* L;         /* corrgauss: chol(Σ₀) lower, d×d row-major host pointer; else NULL */
    void* stream;            /* cudaStream_t to launch on, or NULL → library-owned stream */
} muse_cfg;

/* per-kernel-class device timing, measured with CUDA events on the launch stream */
typedef struct muse_profile {
    int64_t launches;        /* kernels launched since the last reset */
    int64_t solve_launches;  /* solver passes (streaming kernel + its re-solve launch count as one) */
    double  solve_ms;        /* summed device time of those launches (events) */
    double  solve_units;     /* MAP+score units they processed */
    double  solve_bytes;     /* algorithmic bytes they account for */
    int64_t draw_launches;
    double  draw_ms;
    int64_t other_launches;
    double  other_ms;
    int64_t redo_units;      /* units the streaming kernel handed back to the generic kernel (since handle creation) */
    double  solve_flops;     /* FP64 tensor-core flops of the solver launches (corrgauss: 2·rows·d² per product) */
} muse_profile;

/* the same solver figures split by the kind of pass (cold pass, warm pass, …) */
#define MUSE_PASS_COLD      0   /* start = zero(z): muse! iteration 1                      src/muse.jl:151, 169-176 */
#define MUSE_PASS_WARM      1   /* start = previous ẑ or user z₀: muse! iterations ≥ 2     src/muse.jl:169, 181     */
#define MUSE_PASS_TRUTH     2   /* start = simulated z: get_J!                             src/muse.jl:508-514      */
#define MUSE_PASS_FIDUCIAL  3   /* get_H! fiducial solve of the master draw                src/muse.jl:417-423      */
#define MUSE_PASS_FD        4   /* get_H! finite-difference virtual sims                   src/muse.jl:426-433      */
#define MUSE_PASS_KINDS     5
typedef struct muse_pass_profile {
    int64_t launches[MUSE_PASS_KINDS];
    double  ms[MUSE_PASS_KINDS];      /* summed device time (CUDA events on the launch stream) */
    double  units[MUSE_PASS_KINDS];   /* MAP+score units */
    double  bytes[MUSE_PASS_KINDS];   /* algorithmic bytes */
} muse_pass_profile;

int  muse_b200_abi_version(void);

/* lifetime ------------------------------------------------------------------------------- */
int  muse_b200_create(const muse_cfg* cfg, muse_handle** out);
int  muse_b200_destroy(muse_handle* h);
const char* muse_b200_last_error(const muse_handle* h);   /* h may be NULL: last create() error */
int  muse_b200_set_stream(muse_handle* h, void* cuda_stream);

/* inputs --------------------------------------------------------------------------------- */
/* observed data `prob.x` (src/simple.jl:5; used at src/muse.jl:170) */
int  muse_b200_set_data(muse_handle* h, const double* x_dat /* d */);
/* parity mode: base normals of the local sims, ξ (latent) and ν (noise), nsims×d each, plus the
 * master stream's own draw (the sim `sample_x_z(prob, copy(rng), θ)` makes at src/muse.jl:151,418).
 * Replaces `split_rng` (src/util.jl:85-92): sim k sees the same normals at every θ. */
int  muse_b200_set_draws(muse_handle* h, const double* xi, const double* nu,
                         const double* xi_master /* d */, const double* nu_master /* d */);
/* parity mode, multi-GPU only: base normals of the get_H! shard (nsims_h×d each) */
int  muse_b200_set_draws_h(muse_handle* h, const double* xi_h, const double* nu_h);
/* throughput mode: Philox4x32-10 + Box–Muller on the device, keyed by (seed, global sim index,
 * stream, element pair) so results do not depend on how sims are sharded over GPUs. */
int  muse_b200_seed_draws(muse_handle* h, uint64_t seed);
/* read back draws (tests): rows [first, first+count) of ξ and ν; row index nsims = master draw */
int  muse_b200_get_draws(muse_handle* h, int32_t first, int32_t count, double* xi_out, double* nu_out);
/* user start guess z₀ (`z₀` keyword of muse!/get_J!/get_H!, src/muse.jl:117, 309, 488) */
int  muse_b200_set_z0(muse_handle* h, const double* z0 /* d */);

/* the hot path --------------------------------------------------------------------------- */
/* Body of the mapped blocks src/muse.jl:169-176 (include_data=1, warm start ZEROS/PREV) and
 * src/muse.jl:508-514 (include_data=0, warm start TRUTH): for the data unit (optional) and local
 * sims [first_sim, first_sim+count):
 *     x ← prob.x  |  sample_x_z(rng_k, theta_sim).x
 *     ẑ ← argmin_z −logLike(x, z, theta_eval) by L-BFGS(m)+HagerZhang from the chosen start until
 *         ‖∇z‖_∞ ≤ atol                                   (src/interface.jl:162-166)
 *     g ← ∇θ logLike(x, ẑ, theta_eval)                    (src/simple.jl:92)
 * Outputs are indexed by unit in the order [data?, sims...]; any output pointer may be NULL.
 * ẑ stays resident on the device (warm start of the next call; muse_b200_get_maps). */
int  muse_b200_map_score(muse_handle* h, const double* theta_sim, const double* theta_eval,
                         double atol, int32_t include_data, int32_t warm_start,
                         int32_t first_sim, int32_t count,
                         double* g_out      /* units × ntheta */,
                         int32_t* iters_out /* units: L-BFGS iterations      */,
                         int32_t* fg_out    /* units: value+gradient evaluations */,
                         double* gnorm_out  /* units: final ‖∇z‖_∞           */,
                         int32_t* status_out/* units: MUSE_STATUS_*          */);
/* same work enqueued on the stream without a host sync; results stay in device buffers until
 * muse_b200_fetch() copies them out (bench: device-resident timing).  For corrgauss the call returns once the
 * lock-step rounds have finished (the host counts the active units between rounds); results still wait for fetch. */
int  muse_b200_map_score_async(muse_handle* h, const double* theta_sim, const double* theta_eval,
                               double atol, int32_t include_data, int32_t warm_start,
                               int32_t first_sim, int32_t count);
int  muse_b200_fetch(muse_handle* h, int32_t units, double* g_out, int32_t* iters_out,
                     int32_t* fg_out, double* gnorm_out, int32_t* status_out);

/* Device address of the score matrix the last map_score[_async] produced (units × ntheta doubles, row-major, in
 * the order [data?, sims...]); valid on the handle's stream until the next call that resizes the outputs.  Lets a
 * multi-GPU host all-gather the scores with NCCL straight from device memory (the one exchange step of the path,
 * src/muse.jl:177-183) without a host round trip. */
int  muse_b200_device_scores(muse_handle* h, double** g_dev, int32_t* capacity_units);

/* The exchange step across ranks (one process per GPU on one node): an NCCL all-gather of score rows, run on the
 * handle's stream directly from the device score matrix.  Replaces the master-side gather of the reference's pmap
 * (src/muse.jl:169, 177-183).  Rank 0 creates the id, the host distributes its 128 bytes, every rank calls
 * comm_init on its handle.  allgather_scores sends this rank's rows [first_row, first_row + counts[rank]) and
 * returns the concatenation over ranks (Σ counts × ntheta doubles, rank order) in out_host; allgather_rows does the
 * same for rows held on the host (the per-sim Jacobians of get_H!, src/muse.jl:426, 446). */
int  muse_b200_comm_unique_id(uint8_t* id_out /* 128 bytes */);
int  muse_b200_comm_init(muse_handle* h, int32_t nranks, int32_t rank, const uint8_t* id /* 128 bytes */);
int  muse_b200_comm_destroy(muse_handle* h);
int  muse_b200_allgather_scores(muse_handle* h, int32_t first_row, const int32_t* counts /* nranks */, double* out_host);
int  muse_b200_allgather_rows(muse_handle* h, const double* local_host, int32_t ncol, const int32_t* counts, double* out_host);

/* The same exchange step WITHOUT a collective launch: with these buffers in place muse_b200_muse_solve runs the whole solve in
 * one cooperative kernel per rank, whose CTA 0 stores this rank's score rows (and get_H!'s finite-difference scores) straight
 * into every peer's gathered buffer over NVLink (peer-mapped device memory, CUDA IPC) and raises a flag there; the θ update
 * then reads the gathered rows locally.  p2p_alloc creates this rank's region — 2 parities × 4 blocks of block_doubles doubles,
 * block_doubles ≥ nranks × max(sims per rank × ntheta, H sims per rank × 2·ntheta²) — and returns its 64-byte IPC handle; the host
 * distributes the handles (like the NCCL id) and every rank calls p2p_connect with all of them in rank order.  Collective:
 * every rank of the communicator must make the same calls.  Results are bit-identical to the NCCL path and to one GPU. */
int  muse_b200_p2p_alloc(muse_handle* h, int32_t nranks, int32_t rank, int64_t block_doubles, uint8_t* handle_out /* 64 bytes */);
int  muse_b200_p2p_connect(muse_handle* h, const uint8_t* handles /* nranks × 64 bytes */);
int  muse_b200_p2p_info(muse_handle* h, int64_t* block_doubles, int32_t* ready);

/* Finite-difference branch of get_H! (src/muse.jl:417-442 + pjacobian src/util.jl:9-26 with
 * fdm = central_fdm(3,1) and an explicit step): one fiducial MAP of the master-stream draw from
 * zero(z) (the reference computes nsims_H identical copies, :417-423), then for the first
 * nsims_H sims of the H shard (cfg.nsims_h / h_sim_offset; by default the local shard) and each θ-component n: sims at theta0 ± step[n]·e_n (same base normals), MAP and
 * score at theta0 from the fiducial start; column n = (½ g₊ − ½ g₋)/step[n].  The centre
 * evaluation, which central_fdm multiplies by 0, is not executed.
 * Hs_out[k][i][n] = ∂ g_i / ∂ θ_n of sim k (row-major ntheta×ntheta per sim). */
int  muse_b200_fd_jacobian(muse_handle* h, const double* theta0, const double* step,
                           int32_t nsims_H, double atol,
                           double* Hs_out /* nsims_H × ntheta × ntheta */,
                           int32_t* status_out /* nsims_H × ntheta × 2, may be NULL */);

/* Start vector of the fiducial solve of the two entry points around this comment: MUSE_START_ZEROS (default; the reference's
 * ẑ_guess_from_truth, src/interface.jl:184-186) or MUSE_START_USER (the `z₀` keyword of get_H!, src/muse.jl:309, 419; needs
 * muse_b200_set_z0).  Stays in force until changed. */
int  muse_b200_fd_start(muse_handle* h, int32_t start);

/* The same launch sequence with the raw scores returned and arbitrary sample points: row 2n / 2n+1 of theta_sims is the
 * "−" / "+" point at which the sims of Jacobian column n are generated (src/muse.jl:430); MAP and score are taken at
 * theta_eval (:431-432).  g_out[k][2n+s][i] = g_i of sim k at point (n, s).  This is what a host needs when θ lives in a
 * bounded domain (transform_θ / inv_transform_θ, src/interface.jl:14-28): the reference perturbs the UNtransformed θ₀
 * (src/muse.jl:428-432, src/util.jl:15), so the points are not symmetric in the kernels' unconstrained parameters and
 * the host forms sum(fs .* [-1/2, 0, 1/2]) / step and the change of variables itself. */
int  muse_b200_fd_scores(muse_handle* h, const double* theta_eval, const double* theta_sims /* 2·ntheta × ntheta */,
                         int32_t nsims_H, double atol,
                         double* g_out /* nsims_H × 2·ntheta × ntheta */,
                         int32_t* status_out /* nsims_H × 2·ntheta, may be NULL */);

/* The implicit-differentiation branch of get_H! (src/muse.jl:335-405, `implicit_diff = true`) for the first nsims_H sims of the
 * shard: ẑ = MAP at theta0 from zero(z) (start = MUSE_START_ZEROS) or the user's z₀ (MUSE_START_USER) with ∇z_logLike_atol = 1e-1
 * (hard-coded in the reference, :346); H = H1 + H2 with H1 = ∂θ_sim ∇θ′logLike at fixed ẑ (zero for the registered families) and
 * H2 = −(∂θ ∇z logLike)ᵀ · A⁻¹ · (∂θ_sim ∇z logLike), A = ∇²z logLike — the nested-AD derivatives of the reference in closed form,
 * A⁻¹ by conjugate gradients (`implicit_diff_cg_kwargs.maxiter`, Pl = I): exact after one iteration for the isotropic families,
 * a batched CG on the FP64 tensor-core DGEMM for corrgauss.  cg_iters_out: CG iterations per sim and Jacobian column
 * (nsims_H × ntheta; the reference's `implicit_diff_cg_hists`).  Not provided on handles with a separate H shard. */
int  muse_b200_implicit_h(muse_handle* h, const double* theta0, int32_t nsims_H, int32_t start, int32_t cg_maxiter,
                          double* Hs_out /* nsims_H × ntheta × ntheta */, int32_t* cg_iters_out /* may be NULL */,
                          int32_t* status_out /* nsims_H, may be NULL */);

/* The outer loop of muse! (src/muse.jl:159-236) for the common configuration — constant α, regularize = identity,
 * H⁻¹_update = :sims, prior flat or independent Normal(mean, sigma) per component — run inside the library so that
 * no interpreter sits between two passes: per iteration one map_score pass (data + local sims, start zeros / user z₀
 * on the first, previous ẑ afterwards), the exchange step when a communicator exists (counts = sims per rank), then
 *     g_like = g_dat − mean(g_sims)                     :183      g_post = g_like + ∇logPrior(θ)           :184-185
 *     H⁻¹_like = Diagonal(−1 ./ var(g_sims))            :188-189  H⁻¹_post = inv(inv(H⁻¹_like) + ∇²logPrior) :207-208
 *     θ ← θ − α H⁻¹_post g_post                         :224
 * and the convergence test sqrt(−Δθ' H⁻¹_post Δθ) < θ_rtol (:163-166) before iteration i > 2.
 * Histories are row-major with one row per executed iteration; *_hist arrays must hold maxsteps rows.
 * Returns MUSE_ESTATE with a message if a unit ended with a non-finite objective (src/interface.jl:170). */
typedef struct muse_iterate_out {
    int32_t n_iter;          /* iterations executed */
    double* theta_final;     /* ntheta: θ after the last update (result.θ, :230) */
    double* theta_hist;      /* maxsteps × ntheta: θ at which iteration i evaluated */
    double* g_dat_hist;      /* maxsteps × ntheta */
    double* g_sims_hist;     /* maxsteps × nsims_total × ntheta (all ranks' sims, global order) */
    double* g_like_hist;     /* maxsteps × ntheta */
    double* g_prior_hist;    /* maxsteps × ntheta */
    double* h_inv_like_hist; /* maxsteps × ntheta (diagonal) */
    double* h_prior_hist;    /* maxsteps × ntheta (diagonal) */
    double* h_inv_post_hist; /* maxsteps × ntheta (diagonal) */
    double* seconds_hist;    /* maxsteps: wall time of the iteration (history[i].t, :210) */
    int32_t* iters_hist;     /* maxsteps × (nsims_local + 1): per-unit L-BFGS iterations (unit 0 = data) */
    int32_t* fg_hist;        /* maxsteps × (nsims_local + 1) */
    double* gnorm_hist;      /* maxsteps × (nsims_local + 1) */
    int32_t* status_hist;    /* maxsteps × (nsims_local + 1) */
} muse_iterate_out;
int  muse_b200_muse_iterate(muse_handle* h, const double* theta0, int32_t nsims_total, const int32_t* counts,
                            int32_t maxsteps, double theta_rtol, double atol, double alpha, int32_t first_start,
                            const double* prior_mean, const double* prior_sigma /* NULL: flat prior */,
                            muse_iterate_out* out);

/* The covariance stage muse!(…; get_covariance = true) runs after the loop (src/muse.jl:244-247) when no new sims are
 * needed for J: J = var | cov(corrected) of the last scores gs (src/muse.jl:499-502, 529); step = 0.1 ./ std(gs)
 * (:411-413); the finite-difference Jacobians of this rank's H shard (muse_b200_fd_jacobian) and their exchange;
 * H = mean(Hs) (:446); Σ⁻¹ = H'·inv(J)·H − ∇²logPrior(θ), Σ = inv(Σ⁻¹) (finalize_result!, :535-541).
 * counts_h = H sims per rank (NULL without a communicator). */
typedef struct muse_cov_out {
    double* J;          /* ntheta × ntheta */
    double* step;       /* ntheta */
    double* Hs;         /* nsims_h_total × ntheta × ntheta (all ranks, global sim order) */
    double* H;          /* ntheta × ntheta */
    double* Sigma_inv;  /* ntheta × ntheta */
    double* Sigma;      /* ntheta × ntheta */
} muse_cov_out;
int  muse_b200_muse_covariance(muse_handle* h, const double* theta, const double* gs /* nsims_total × ntheta */,
                               int32_t nsims_total, int32_t nsims_h_total, const int32_t* counts_h, double atol,
                               const double* prior_sigma /* NULL: flat prior */, muse_cov_out* out);

/* muse_iterate + muse_covariance in ONE call with the θ update on the DEVICE (csrc/muse_outer.cu): a single-CTA kernel per
 * pass does the arithmetic of src/muse.jl:183-224 and the convergence test of :163-166 and writes the θ-dependent constants
 * of the next pass into device memory, so consecutive passes — and get_H!'s fiducial solve and virtual sims, whose step
 * 0.1 ./ std(gs) (:411-413) is derived on the device too — follow each other on the stream without a host round trip; the host
 * synchronises once per chunk of three passes (the typical solve: once).  Isotropic families only (corrgauss: use
 * muse_iterate), maxsteps ≤ 64.  Same outputs as the two calls it replaces; θ agrees with them to round-off (parallel
 * reductions, device libm), not bit for bit; seconds_hist holds the chunk's wall time divided by its iterations.
 * get_covariance = 0: the loop only (cov may be NULL).
 * By default the first three iterations and the covariance stage run as ONE cooperative kernel launch (csrc/muse_iso_stream.cu:
 * solve_persist_kernel; with muse_b200_p2p_* buffers in place also across GPUs, the exchange inside the kernel); that launch keeps
 * no resident MAPs — like muse! itself, whose ẑs do not outlive the call (src/muse.jl:151): after muse_solve, muse_b200_get_maps
 * and warm_start = PREV see zero(z) unless the loop ran past three iterations.  Use map_score / the host driver's save_MAPs to keep ẑ. */
int  muse_b200_muse_solve(muse_handle* h, const double* theta0, int32_t nsims_total, const int32_t* counts,
                          int32_t maxsteps, double theta_rtol, double atol, double alpha, int32_t first_start,
                          const double* prior_mean, const double* prior_sigma /* NULL: flat prior */,
                          int32_t get_covariance, int32_t nsims_h_total, const int32_t* counts_h,
                          muse_iterate_out* out, muse_cov_out* cov);

/* MAPs of units [first_unit, first_unit+count) (unit 0 = data) — `save_MAPs`, src/muse.jl:139-143,219 */
int  muse_b200_get_maps(muse_handle* h, int32_t first_unit, int32_t count, double* z_out /* count × d */);

/* diagnostics ---------------------------------------------------------------------------- */
/* The FP64 tensor-core GEMM of the correlated-Gaussian family (csrc/muse_dgemm.cu), C = A·B row-major with
 * M % 128 == N % 128 == K % 16 == 0: on host operands (tests) and timed on device-resident operands (bench). */
int  muse_b200_dgemm_host(const double* A, const double* B, double* C, int32_t M, int32_t N, int32_t K);
int  muse_b200_dgemm_time(int32_t M, int32_t N, int32_t K, int32_t reps, double* ms_per_gemm);
int  muse_b200_profile_reset(muse_handle* h, int32_t enable);
int  muse_b200_profile_get(muse_handle* h, muse_profile* out);
int  muse_b200_profile_passes(muse_handle* h, muse_pass_profile* out);
/* diagnostics: 16 int64 stamps per row of the last solver launch.  Generic kernel: one row per unit, SM-clock stamps
 * of its controller.  Streaming kernel: one row per CTA — [start ns, end ns, SM id, producer / finisher / consumers
 * done ns] (globaltimer).  out == NULL arms the facility for up to `items` rows (0 disarms); otherwise copies. */
int  muse_b200_debug_timeline(muse_handle* h, int32_t items, int64_t* out);
/* geometry chosen for the solver: threads per solve group, CTAs per cluster, resident groups */
int  muse_b200_geometry(muse_handle* h, int32_t* group_threads, int32_t* cluster, int32_t* groups);

#ifdef __cplusplus
}
#endif
#endif /* MUSE_B200_H */
