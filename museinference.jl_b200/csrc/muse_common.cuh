// muse_common.cuh — shared declarations of libmuse_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/muse_b200.h"

namespace muse {

constexpr int kMaxTheta = MUSE_MAX_NTHETA;
constexpr int kMaxThetaSim = 2 * MUSE_MAX_NTHETA + 1;   // θ_sim table: θ₀ and θ₀ ± h_n e_n

// ---- isotropic-Gaussian-latent families (F1 funnel, F2 hierarchical Gaussian) ------------
//   -logLike(x,z|θ) = ½ [ Σ (x-z)² + a Σ (z-μ)² ] + half_cst
//   F1: a = e^{-θ},  μ = 0, half_cst = ½ d θ      (/root/reference/src/simple.jl:66-68)
//   F2: a = e^{-2ℓ}, μ = μ, half_cst = d ℓ        
struct IsoEval {
    double a;
    double mu;
    double half_cst;
    double cspec;     // 1/(1+a): exact line minimum along −∇f for these families (speculated step)
};
//   sample: z = mu + sig ξ,  x = z + ν            (src/simple.jl:61-65)
struct IsoSample {
    double sig;
    double mu;
};

// θ-dependent constants of a launch held in DEVICE memory: written by the θ-update kernel of the device-resident outer loop
// (muse_outer.cu) and read by the solver kernels of the next pass, so that no host round trip sits between two passes.
// skip != 0: the pass is not needed (the loop has converged or failed) — every kernel of the launch returns at once.
struct DynConsts {
    IsoEval ev;
    IsoSample smp[kMaxThetaSim];
    int skip;
    int pad;
};

// Lazy ẑ (solve_persist_kernel): when every earlier solve of a unit within the launch took the fast path, its ẑ is an elementwise
// function of the unit's base normals — level l: x = mus + sig·ξ + ν, g = a(z − mu) − (x − z), z ← z + cspec·(−g) — so a pass
// recomputes its start vector in registers instead of reading it, and no pass stores ẑ.
constexpr int kMaxLazy = 4;
struct LazyLevel {
    double sig, mus, a, mu, cspec;
};
struct LazyLevels {
    int nlev;        // earlier passes of this launch
    int store;       // 1: this pass materialises ẑ (the chain of launches may continue from it)
    LazyLevel lev[kMaxLazy];
};

// where the start vector z₀ of a unit comes from
enum StartKind : int {
    kStartZero = 0,        // z₀ ≡ 0, never read from memory
    kStartOwn = 1,         // the unit's current resident ẑ
    kStartTruth = 2,       // the simulated latent (materialised into the unit's buffer A)
    kStartShared = 3,      // one vector shared by all units (user z₀ / fiducial ẑ), read-only
    kStartSharedKeep = 4,  // shared vector, materialised into buffer A so the unit keeps it
};

// which buffer holds a unit's resident ẑ
enum ZState : int { kZZero = 0, kZA = 1, kZB = 2 };

// One launch of the persistent solver.
struct SolveLaunch {
    int d;
    int ld;              // row stride (doubles) of every N×d array
    int ntheta;
    int family;
    int nitems;
    int mode;            // 0: units (data? + sims), 1: finite-difference virtual sims
    int include_data;    // mode 0
    int first_sim;       // mode 0
    int start_kind;      // StartKind
    int lbfgs_m;
    int max_iters;
    double atol;
    IsoEval ev;                      // at θ_eval
    IsoSample smp[kMaxThetaSim];     // mode 0: [0]; mode 1: index 2n+s ↔ θ₀ ∓/± h_n e_n
    const DynConsts* dyn;            // non-null: ev / smp come from device memory instead (launch_ev / launch_smp below)
    // inputs
    const double* xi;    // nsims+1 rows (row nsims = master draw)
    const double* nu;
    const double* xdat;  // d
    const double* zshared;  // d (user z₀ or fiducial ẑ)
    // alternative: the shared start is the resident ẑ of an earlier launch on the same stream
    // (fiducial solve → FD sims, no host sync in between): pick A/B/zero from *zshared_state
    const int* zshared_state;
    const double* zsharedA;
    const double* zsharedB;
    int master_row;      // row of the master draw in xi/nu
    // per-unit resident state (mode 0 rows: unit index; mode 1 rows: item index in H scratch)
    double* xslot;       // materialised x of the unit in flight: one row per resident group (slots × ld)
    double* zA;
    double* zB;
    int* zstate;         // rows
    // per-slot scratch (L-BFGS history, search direction), one slot per resident group
    double* sbuf;        // slots × ld
    double* dxh;         // slots × m × ld
    double* dgh;         // slots × m × ld
    // outputs, indexed by item
    double* g_out;       // nitems × ntheta
    int* iters_out;
    int* fg_out;
    double* gnorm_out;
    double* f_out;
    int* status_out;
    long long* dbg;      // optional timeline: 16 clock64 stamps per item (diagnostics), or null
    // generic kernel as the re-solve pass of the streaming kernel: items come from a device-side list
    const int* item_list;    // null: items 0..nitems-1
    const int* item_count;   // null: nitems
    // streaming kernel (muse_iso_stream.cu)
    int discard_z;       // 1: ẑ is not stored (finite-difference virtual sims; the reference drops it too)
    int seg_chunks;      // chunks per segment
    int nseg;            // segments per unit
    int zrows;           // 1: a z₀ row is streamed (ring stages hold 3 rows), 0: stages hold 2 rows
    int stream_stages;
    double* gpart;       // nitems × nseg × 16 partial sums of the streaming kernel's segments
    int* gcount;         // nitems arrival counters (zero between launches: the last arrival resets its counter)
    int* redo_count;     // number of units handed to the generic kernel by this launch
    int* work_next;      // dynamic work counter of the streaming kernel (zeroed before every launch)
    unsigned long long* redo_total;   // … since handle creation (diagnostics)
    int* redo_items;
    int lean;                // 1 (solve_persist_kernel only): the α = 1 trial is not evaluated element by element — for these families
                             // φ′(1) = a·‖∇f(z₀)‖² and φ(1) = φ(0) − ‖∇f‖² + ½(1 + a)‖∇f‖² exactly, and only their sign / finiteness and the
                             // secant step (≡ c_spec) enter the decisions (muse_iso_stream.cu: fast_replay)
    const LazyLevels* lazy;  // non-null (solve_persist_kernel only): z₀ is recomputed, the zstate cells hold per-unit level masks
};

#if defined(__CUDACC__)
__device__ __forceinline__ IsoEval launch_ev(const SolveLaunch& L) { return L.dyn ? L.dyn->ev : L.ev; }
__device__ __forceinline__ IsoSample launch_smp(const SolveLaunch& L, int t) { return L.dyn ? L.dyn->smp[t] : L.smp[t]; }
__device__ __forceinline__ bool launch_skipped(const SolveLaunch& L) { return L.dyn != nullptr && L.dyn->skip != 0; }
#endif

struct Geometry {
    int group_threads;   // threads per CTA taking part in a solve (32 → one warp per solve)
    int cta_threads;
    int cluster;         // CTAs per cluster (≥1); group = cluster × cta when group_threads > 32
    int groups;          // resident solve groups (= scratch slots)
    int grid;            // CTAs launched
    // streaming kernel (muse_iso_stream.cu)
    int stream;          // first pass of every launch: 0 none (generic kernel only), 1 streaming kernel (TMA ring),
                         // 2 warp-per-unit streaming kernel (small d)
    int stream_grid;     // CTAs (one per SM)
    int seg_chunks;      // chunks per segment
    int nseg;            // segments per unit
    int smem_bytes;      // dynamic shared memory per CTA (the ring)
};

cudaError_t launch_iso_solver(const SolveLaunch& L, const Geometry& geo, cudaStream_t st);
cudaError_t iso_solver_geometry(int d, int want_group, int want_cluster, int device, Geometry* geo);
cudaError_t iso_stream_geometry(int d, int ld, int device, Geometry* geo);
cudaError_t iso_warp_stream_geometry(int device, Geometry* geo);
cudaError_t launch_iso_stream(SolveLaunch& L, const Geometry& geo, cudaStream_t st);
cudaError_t launch_dgemm(const double* A, const double* Bt, double* C, int M, int N, int K, int lda, int ldbt, int ldc,
                         cudaStream_t st);      // C = A·Btᵀ, Bt[N×K] (muse_dgemm.cu)
cudaError_t launch_philox_draws(double* xi, double* nu, int rows, int d, int ld, uint64_t seed,
                                int64_t sim_offset, int master_row, cudaStream_t st);

}  // namespace muse
