// muse_comm.cu — the one exchange step of the path, natively: an NCCL all-gather of per-sim score rows across
// the ranks of one node (one process per GPU), straight from the solver's device output buffer, on the
// solver's stream.
//
// What it replaces.  The reference gathers every per-sim result on the master through Distributed.pmap
// (/root/reference/src/muse.jl:169, 177-183, 508, 529) and reduces there; here each rank keeps its sims' ẑ on its
// GPU and only the N × nθ score matrix (≤ 64 KB) crosses NVLink, once per pass.  All ranks receive it and run
// the same deterministic host arithmetic.
//
// NCCL is loaded lazily with dlopen("libnccl.so.2") so that single-GPU use of the library has no NCCL
// dependency; inside a PyTorch process this resolves to the NCCL build torch has already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>
#include <vector>

#include "muse_handle.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    if (api.lib || !api.err.empty()) return &api;
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return &api; }
#define MUSE_SYM(field, name)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));         \
    if (!api.field) { api.err = std::string("dlsym(") + name + ") failed"; return &api; }
    MUSE_SYM(GetUniqueId, "ncclGetUniqueId")
    MUSE_SYM(CommInitRank, "ncclCommInitRank")
    MUSE_SYM(CommDestroy, "ncclCommDestroy")
    MUSE_SYM(AllGather, "ncclAllGather")
    MUSE_SYM(GetErrorString, "ncclGetErrorString")
#undef MUSE_SYM
    return &api;
}

// this rank's score rows → the send buffer; the rows of units whose solve ended with a non-finite objective go out as NaN,
// so that EVERY rank sees the failure in the gathered matrix and takes the same error decision (src/interface.jl:170) —
// a rank that failed alone would leave the others waiting in the next exchange
__global__ void pack_rows_kernel(const double* __restrict__ src, const int* __restrict__ status, double* __restrict__ dst, int n, int ncol) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * ncol) return;
    const int row = i / ncol;
    const bool bad = status != nullptr && status[row] == MUSE_STATUS_NONFINITE;
    dst[i] = bad ? __longlong_as_double(0x7ff8000000000000LL) : src[i];
}

}  // namespace

void muse_comm_release(muse_handle* h) {
    if (h->comm) {
        NcclApi* a = nccl_api();
        if (a->CommDestroy) a->CommDestroy((ncclComm_t)h->comm);
        h->comm = nullptr;
    }
    cudaFree(h->comm_send); cudaFree(h->comm_recv); cudaFreeHost(h->comm_host);
    h->comm_send = h->comm_recv = h->comm_host = nullptr;
    h->comm_cap = 0;
}

// ---- exchange through peer-mapped memory -----------------------------------------------------------------------------
// solve_persist_kernel (muse_iso_stream.cu) performs the exchange step itself: CTA 0 of every rank stores its score rows into
// every peer's gathered-score buffer over NVLink and raises a flag there.  Here: the buffers.  Every rank allocates one region
//   [ flags: 16 × 2 u64 | 2 parities × (kOuterSlots + 1) blocks of `block_doubles` doubles ]
// exports it as a CUDA IPC handle (64 bytes, distributed by the host like the NCCL id), and maps the regions of its peers.
// Parity = solve number mod 2: a rank that is already in the next solve never overwrites rows a slower peer has not read yet
// (it cannot get two solves ahead — every pass needs every rank).  Flags only grow (epoch = 8 × solve number + phase).
constexpr size_t kP2pFlagBytes = 256;

void muse_p2p_release(muse_handle* h) {
    for (int q = 0; q < 16; ++q) {
        if (h->p2p_peer[q] && h->p2p_peer[q] != h->p2p_region) cudaIpcCloseMemHandle(h->p2p_peer[q]);
        h->p2p_peer[q] = nullptr;
    }
    cudaFree(h->p2p_region);
    cudaFreeHost(h->p2p_host);
    h->p2p_region = nullptr;
    h->p2p_host = nullptr;
    h->p2p_block = 0;
    h->p2p_ready = false;
    cudaGetLastError();
}

extern "C" {

int muse_b200_p2p_alloc(muse_handle* h, int32_t nranks, int32_t rank, int64_t block_doubles, uint8_t* handle_out /* 64 bytes */) {
    if (!h || !handle_out || nranks < 2 || nranks > 16 || rank < 0 || rank >= nranks || block_doubles < 1) return MUSE_EINVAL;
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) { h->err = "cudaSetDevice"; return MUSE_ECUDA; }
    if (h->stream) cudaStreamSynchronize(h->stream);
    muse_p2p_release(h);
    const size_t blocks = 2 * (size_t)(kOuterSlots + 1);
    const size_t bytes = kP2pFlagBytes + blocks * (size_t)block_doubles * sizeof(double);
    if (cudaMalloc(&h->p2p_region, bytes) != cudaSuccess || cudaMemset(h->p2p_region, 0, bytes) != cudaSuccess ||
        cudaMallocHost(&h->p2p_host, (size_t)(kOuterSlots + 1) * (size_t)block_doubles * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        muse_p2p_release(h);
        h->err = "allocation of the peer exchange region failed";
        return MUSE_ENOMEM;
    }
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t size");
    const cudaError_t e = cudaIpcGetMemHandle(&mh, h->p2p_region);
    if (e != cudaSuccess) {
        h->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
        cudaGetLastError();
        muse_p2p_release(h);
        return MUSE_ECUDA;
    }
    std::memcpy(handle_out, &mh, sizeof(mh));
    h->p2p_block = block_doubles;
    h->p2p_nranks = nranks;
    h->p2p_rank = rank;
    return MUSE_OK;
}

int muse_b200_p2p_connect(muse_handle* h, const uint8_t* handles /* nranks × 64 bytes, rank order */) {
    if (!h || !handles) return MUSE_EINVAL;
    if (!h->p2p_region) { h->err = "p2p_connect before p2p_alloc"; return MUSE_ESTATE; }
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) { h->err = "cudaSetDevice"; return MUSE_ECUDA; }
    for (int q = 0; q < h->p2p_nranks; ++q) {
        if (q == h->p2p_rank) { h->p2p_peer[q] = h->p2p_region; continue; }
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, handles + (size_t)q * 64, sizeof(mh));
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            h->err = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(q) + "): " + cudaGetErrorString(e);
            cudaGetLastError();
            muse_p2p_release(h);
            return MUSE_ECUDA;
        }
        h->p2p_peer[q] = static_cast<unsigned char*>(ptr);
    }
    h->p2p_ready = true;
    return MUSE_OK;
}

int muse_b200_p2p_info(muse_handle* h, int64_t* block_doubles, int32_t* ready) {
    if (!h) return MUSE_EINVAL;
    if (block_doubles) *block_doubles = h->p2p_block;
    if (ready) *ready = h->p2p_ready ? 1 : 0;
    return MUSE_OK;
}

int muse_b200_comm_unique_id(uint8_t* out /* 128 bytes */) {
    if (!out) return MUSE_EINVAL;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return MUSE_EUNSUPPORTED;
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) return MUSE_ECUDA;
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    std::memcpy(out, &id, sizeof(id));
    return MUSE_OK;
}

int muse_b200_comm_init(muse_handle* h, int32_t nranks, int32_t rank, const uint8_t* id_bytes) {
    if (!h || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return MUSE_EINVAL;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) { h->err = a->err; return MUSE_EUNSUPPORTED; }
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) { h->err = "cudaSetDevice"; return MUSE_ECUDA; }
    muse_comm_release(h);
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = a->CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) { h->err = std::string("ncclCommInitRank: ") + a->GetErrorString(r); return MUSE_ECUDA; }
    h->comm = comm;
    h->comm_nranks = nranks;
    h->comm_rank = rank;
    return MUSE_OK;
}

int muse_b200_comm_destroy(muse_handle* h) {
    if (!h) return MUSE_EINVAL;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    muse_comm_release(h);
    return MUSE_OK;
}

// after the stream has been synchronised: copy the gathered rows (rank order) out of the pinned receive area
void muse_comm_unpack(muse_handle* h, int ncol, const int32_t* counts, double* out_host) {
    const int R = h->comm_nranks;
    int maxc = 1;
    for (int r = 0; r < R; ++r) maxc = counts[r] > maxc ? counts[r] : maxc;
    const size_t need = (size_t)maxc * ncol;
    size_t off = 0;
    for (int q = 0; q < R; ++q) {
        std::memcpy(out_host + off, h->comm_host + (size_t)q * need, (size_t)counts[q] * ncol * sizeof(double));
        off += (size_t)counts[q] * ncol;
    }
}

// gather `counts[q]` rows of `ncol` doubles from every rank q; this rank's rows come from device memory (`src_dev`)
// or from the host (`src_host`)
static int allgather_impl(muse_handle* h, const double* src_dev, const double* src_host, int ncol, const int32_t* counts,
                          double* out_host, const int* status_dev = nullptr) {
    if (!h->comm) { h->err = "no communicator (muse_b200_comm_init)"; return MUSE_ESTATE; }
    NcclApi* a = nccl_api();
    const int R = h->comm_nranks;
    int maxc = 1;
    for (int r = 0; r < R; ++r) {
        if (counts[r] < 0) return MUSE_EINVAL;
        if (counts[r] > maxc) maxc = counts[r];
    }
    const int mine = counts[h->comm_rank];
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) { h->err = "cudaSetDevice"; return MUSE_ECUDA; }
    const size_t need = (size_t)maxc * ncol;               // doubles per rank slot
    if (need > (size_t)h->comm_cap) {
        cudaFree(h->comm_send); cudaFree(h->comm_recv); cudaFreeHost(h->comm_host);
        h->comm_send = h->comm_recv = h->comm_host = nullptr;
        h->comm_cap = 0;
        if (cudaMalloc(&h->comm_send, need * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&h->comm_recv, need * R * sizeof(double)) != cudaSuccess ||
            cudaMallocHost(&h->comm_host, need * (R + 1) * sizeof(double)) != cudaSuccess) {
            h->err = "allocation of the exchange buffers failed";
            return MUSE_ENOMEM;
        }
        cudaMemsetAsync(h->comm_send, 0, need * sizeof(double), h->stream);
        h->comm_cap = (int)need;
    }
    const size_t bytes_mine = (size_t)mine * ncol * sizeof(double);
    if (mine && src_dev) {
        const int n = mine * ncol;
        pack_rows_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(src_dev, status_dev, h->comm_send, mine, ncol);
        if (cudaGetLastError() != cudaSuccess) { h->err = "pack_rows_kernel launch failed"; return MUSE_ECUDA; }
        h->acc.launches += 1;
    } else if (mine) {
        double* stage = h->comm_host + need * R;            // pinned staging slot behind the receive area
        std::memcpy(stage, src_host, bytes_mine);
        if (cudaMemcpyAsync(h->comm_send, stage, bytes_mine, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { h->err = "cudaMemcpyAsync"; return MUSE_ECUDA; }
    }
    // every rank computes the same `need` from the same counts, so the slots line up
    const ncclResult_t r = a->AllGather(h->comm_send, h->comm_recv, need, ncclDouble, (ncclComm_t)h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + a->GetErrorString(r); return MUSE_ECUDA; }
    h->acc.launches += 1;
    if (cudaMemcpyAsync(h->comm_host, h->comm_recv, need * R * sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) {
        h->err = "exchange copy failed";
        return MUSE_ECUDA;
    }
    if (!out_host) return MUSE_OK;             // enqueue only: the caller synchronises the stream, then unpacks
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "exchange sync failed"; return MUSE_ECUDA; }
    muse_comm_unpack(h, ncol, counts, out_host);
    return MUSE_OK;
}

// enqueue-only variant for the in-library driver: gather on the stream, no synchronisation
int muse_comm_allgather_scores_enqueue(muse_handle* h, int first_row, const int32_t* counts) {
    if (h->comm && first_row + counts[h->comm_rank] > h->out_cap) { h->err = "score rows outside the device output buffer"; return MUSE_EINVAL; }
    return allgather_impl(h, h->g_d + (size_t)first_row * h->cfg.ntheta, nullptr, h->cfg.ntheta, counts, nullptr, h->status_d + first_row);
}

// the same from an arbitrary device source (the device-resident outer loop gathers from a per-pass output block);
// *need_out = doubles per rank slot of the receive area h->comm_recv
int muse_comm_allgather_dev_enqueue(muse_handle* h, const double* src_dev, const int* status_dev, int ncol, const int32_t* counts, size_t* need_out) {
    int maxc = 1;
    for (int r = 0; r < h->comm_nranks; ++r) maxc = counts[r] > maxc ? counts[r] : maxc;
    if (need_out) *need_out = (size_t)maxc * ncol;
    return allgather_impl(h, src_dev, nullptr, ncol, counts, nullptr, status_dev);
}

int muse_b200_allgather_scores(muse_handle* h, int32_t first_row, const int32_t* counts, double* out_host) {
    if (!h || !counts || !out_host || first_row < 0) return MUSE_EINVAL;
    if (h->comm && first_row + counts[h->comm_rank] > h->out_cap) { h->err = "score rows outside the device output buffer"; return MUSE_EINVAL; }
    return allgather_impl(h, h->g_d + (size_t)first_row * h->cfg.ntheta, nullptr, h->cfg.ntheta, counts, out_host, h->status_d + first_row);
}

int muse_b200_allgather_rows(muse_handle* h, const double* local_host, int32_t ncol, const int32_t* counts, double* out_host) {
    if (!h || !counts || !out_host || ncol < 1) return MUSE_EINVAL;
    if (h->comm && counts[h->comm_rank] > 0 && !local_host) return MUSE_EINVAL;
    return allgather_impl(h, nullptr, local_host, ncol, counts, out_host);
}

}  // extern "C"
