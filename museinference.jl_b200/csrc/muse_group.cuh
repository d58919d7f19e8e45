// muse_group.cuh — the set of threads that cooperates on one MAP solve, and its all-reduce.
//
// Three shapes, chosen from the latent dimension d:
//   warp      32 threads, shuffle-only reductions, several solves per CTA      (d ≲ 2 K)
//   CTA       all threads of a CTA, one __syncthreads per reduction            (d ≲ 32 K)
//   cluster   CL CTAs of a thread-block cluster, partials exchanged through distributed
//             shared memory, one barrier.cluster per reduction                 (large d)
// Every reduction is a fixed tree (lane butterfly → warps in index order → CTAs in rank order),
// so a unit's result does not depend on scheduling, on the number of units in the launch or on
// how sims are sharded over GPUs.  All threads receive the result (the scalar L-BFGS /
// Hager–Zhang logic then runs redundantly and uniformly in every thread).
#pragma once
#include <cooperative_groups.h>

namespace muse {
namespace cg = cooperative_groups;

constexpr int kRedMax = 8;   // values per all-reduce

template <int CTA_THREADS, bool WARP_GROUP, int CLUSTER>
struct Group {
    static constexpr int kWarps = CTA_THREADS / 32;
    static constexpr int kGroupsPerCta = WARP_GROUP ? kWarps : 1;
    static constexpr int kSize = WARP_GROUP ? 32 : CTA_THREADS * CLUSTER;
    static constexpr bool kWarpGroup = WARP_GROUP;

    // controller → workers hand-off: named barrier 1 over the whole CTA (the controller warp and
    // the worker warps arrive from different program locations, which bar.sync permits)
    // (`bar.sync` is the .aligned form: every lane of a warp must execute it together.  The controller reaches it
    // right after single-lane code — `if (lane == 0) *scmd = cur;` — so the warp is reconverged first; synccheck
    // flagged the version without __syncwarp as divergent.)
    static __device__ __forceinline__ void cmd_barrier() {
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(CTA_THREADS) : "memory");
    }
    // end of a sweep that has no reduction: every thread is done with the command slot
    __device__ __forceinline__ void sync_exec() {
        if (WARP_GROUP) __syncwarp();
        else __syncthreads();
    }

    struct Smem {
        double wpart[2][kWarps][kRedMax];
        double cpart[2][kRedMax];
    };

    Smem* sm;
    int par;
    int tid;      // thread index inside the group
    int rank;     // CTA rank in cluster

    __device__ __forceinline__ Group(Smem* s) : sm(s), par(0) {
        if (WARP_GROUP) {
            tid = threadIdx.x & 31;
            rank = 0;
        } else if (CLUSTER > 1) {
            rank = (int)cg::this_cluster().block_rank();
            tid = rank * CTA_THREADS + threadIdx.x;
        } else {
            rank = 0;
            tid = threadIdx.x;
        }
    }

    // index of this group among all groups of the grid, and their number
    __device__ __forceinline__ int group_index() const {
        if (WARP_GROUP) return blockIdx.x * kWarps + (threadIdx.x >> 5);
        return blockIdx.x / CLUSTER;
    }
    __device__ __forceinline__ int group_count() const {
        if (WARP_GROUP) return gridDim.x * kWarps;
        return gridDim.x / CLUSTER;
    }

    // v[k] ← reduction over the group; slot k is a max if bit k of MAXMASK is set, else a sum.
    template <int K, unsigned MAXMASK>
    __device__ __forceinline__ void allreduce(double (&v)[K]) {
        static_assert(K <= kRedMax, "too many reduction slots");
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double o = __shfl_xor_sync(0xffffffffu, v[k], off);
                v[k] = ((MAXMASK >> k) & 1u) ? fmax(v[k], o) : v[k] + o;
            }
        }
        if (WARP_GROUP) return;

        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) sm->wpart[par][w][k] = v[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double acc = sm->wpart[par][0][k];
            for (int i = 1; i < kWarps; ++i) {
                const double o = sm->wpart[par][i][k];
                acc = ((MAXMASK >> k) & 1u) ? fmax(acc, o) : acc + o;
            }
            v[k] = acc;
        }
        if (CLUSTER > 1) {
            cg::cluster_group cl = cg::this_cluster();
            if (threadIdx.x < K) sm->cpart[par][threadIdx.x] = v[threadIdx.x];
            cl.sync();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < CLUSTER; ++r) {
                    const double* remote = cl.map_shared_rank(&sm->cpart[par][0], r);
                    const double o = remote[k];
                    acc = (r == 0) ? o : (((MAXMASK >> k) & 1u) ? fmax(acc, o) : acc + o);
                }
                v[k] = acc;
            }
        }
        par ^= 1;
    }

    // make global-memory writes of this group's threads visible to the whole group
    __device__ __forceinline__ void sync_mem() {
        if (WARP_GROUP) {
            __syncwarp();
        } else if (CLUSTER > 1) {
            __threadfence();
            cg::this_cluster().sync();
        } else {
            __syncthreads();
        }
    }
};

}  // namespace muse
