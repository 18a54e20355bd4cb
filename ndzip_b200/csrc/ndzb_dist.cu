// ndzb_dist.cu — the multi-GPU data plane of the hot path behind the C ABI (include/ndzip_b200.h, ndzb_dist_*).
//
// New work: the reference is single-GPU. Hypercubes are independent, so the grid is cut into
// slabs of whole cube rows along the slowest dimension, one slab per rank (one process per GPU, or one thread per GPU
// in a single process). A rank compresses its slab into a SELF-CONTAINED ndzip stream; the only data-path exchange is
//   (1) the cross-rank exclusive scan of the compressed word counts: ONE ncclAllGather of one uint32 per rank,
//       enqueued on a high-priority side stream so that it overlaps whatever the caller runs next on the context's
//       stream (the slab's own decompression does not depend on it), followed by one kernel that turns the local
//       "offset_after" header entries into the global ones (reference src/ndzip/common.hh:342-358), and
//   (2) optionally the final stream gather: ncclSend / ncclRecv of every rank's header slice, cube segment and
//       border segment straight to their final position in the root's buffer, which then holds the stream the
//       reference would have produced for the whole grid, bit for bit.
// NCCL is bound at run time (dlopen of libnccl.so.2 — inside a PyTorch process that is the NCCL torch already
// loaded), so the library has no link-time dependency on it and loads on machines without NCCL.
#include "../../include/ndzip_b200.h"

#include <cuda.h>  // CUdeviceptr, CUresult (the entry point is fetched at run time)
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace {

// ---- the part of nccl.h this file needs (ABI-stable since NCCL 2.0; /usr/include/nccl.h) ---------------------
struct ncclComm;
using ncclComm_t = ncclComm *;
struct ncclUniqueId {
    char internal[128];
};
static_assert(sizeof(ncclUniqueId) == NDZB_UNIQUE_ID_BYTES, "unique id size");
using ncclResult_t = int;  // ncclSuccess == 0
enum : int { kNcclUint8 = 1, kNcclUint32 = 3 };

struct nccl_api {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

thread_local char g_dist_error[256] = "no error";

int fail(const char *what, const char *detail) {
    snprintf(g_dist_error, sizeof g_dist_error, "%s: %s", what, detail);
    return NDZB_ERR_CUDA;
}

const nccl_api &nccl() {
    static nccl_api api = [] {
        nccl_api a;
        void *h = nullptr;
        if (const char *path = getenv("NDZB_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        auto sym = [&](const char *name) { return dlsym(h, name); };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllGather && a.Broadcast && a.Send && a.Recv && a.GroupStart
                && a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

#define DIST_CUDA(call)                                                      \
    do {                                                                     \
        const cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return fail(#call, cudaGetErrorString(e__)); \
    } while (0)
#define DIST_NCCL(call)                                                          \
    do {                                                                         \
        const ncclResult_t r__ = (call);                                         \
        if (r__ != 0) return fail(#call, nccl().GetErrorString(r__));            \
    } while (0)
#define DIST_NDZB(call)                 \
    do {                                \
        const int s__ = (call);         \
        if (s__ != NDZB_OK) return s__; \
    } while (0)

constexpr uint32_t side_for(int dims) { return dims == 1 ? 4096u : dims == 2 ? 64u : 16u; }

class on_device {
  public:
    explicit on_device(int device) {
        if (cudaGetDevice(&_previous) != cudaSuccess) _previous = -1;
        if (_previous != device && cudaSetDevice(device) == cudaSuccess) _switched = true;
    }
    ~on_device() {
        if (_switched && _previous >= 0) cudaSetDevice(_previous);
    }

  private:
    int _previous = -1;
    bool _switched = false;
};

}  // namespace

// The slab partition: [begin, end) along dimension 0 for every rank. Interior boundaries are multiples of the cube
// side, cube rows are dealt out as evenly as possible, the last rank also takes the trailing partial rows. With such
// slabs both the cube order and the ascending-linear-index border order of the slab streams concatenate to the
// global stream's (reference src/ndzip/common.hh:245-306, 414-433).
extern "C" int ndzb_dist_plan(int dtype, int dims, const uint32_t *global_size, int world, int rank, ndzb_dist_layout *out) {
    if ((dtype != NDZB_F32 && dtype != NDZB_F64) || dims < 1 || dims > 3 || !global_size || world < 1 || rank < 0 || rank >= world || !out) {
        return NDZB_ERR_INVALID_ARGUMENT;
    }
    const uint32_t side = side_for(dims);
    const uint32_t cube_rows = global_size[0] / side;
    memset(out, 0, sizeof *out);
    uint32_t begin = 0, cube_index = 0;
    uint64_t border_base = 0;
    for (int r = 0; r < world; ++r) {
        const uint32_t rows = cube_rows / world + (static_cast<uint32_t>(r) < cube_rows % world ? 1u : 0u);
        uint32_t end = begin + rows * side;
        if (r == world - 1) end = global_size[0];
        uint32_t slab[3] = {end - begin, dims > 1 ? global_size[1] : 0u, dims > 2 ? global_size[2] : 0u};
        const uint32_t cubes = ndzb_num_hypercubes(dims, slab);
        const uint64_t border = ndzb_border_element_count(dims, slab);
        if (r == rank) {
            out->slab_begin = begin;
            out->slab_end = end;
            for (int d = 0; d < 3; ++d) out->slab_size[d] = slab[d];
            out->local_cubes = cubes;
            out->cube_index_base = cube_index;
            out->local_header_words = ndzb_header_words(dtype, cubes);
            out->local_border_words = border;
            out->border_base = border_base;
            out->local_bound_words = ndzb_compressed_length_bound(dtype, dims, slab);
        }
        begin = end;
        cube_index += cubes;
        border_base += border;
    }
    out->global_cubes = cube_index;
    out->global_header_words = ndzb_header_words(dtype, cube_index);
    out->global_border_words = border_base;
    out->global_bound_words = ndzb_compressed_length_bound(dtype, dims, global_size);
    if (cube_index != ndzb_num_hypercubes(dims, global_size) || border_base != ndzb_border_element_count(dims, global_size)) {
        return NDZB_ERR_INVALID_ARGUMENT;  // cannot happen for slabs of whole cube rows
    }
    return NDZB_OK;
}

struct ndzb_dist {
    int dtype = 0, dims = 0, rank = 0, world = 1, device = 0;
    uint32_t global_size[3] = {0, 0, 0};
    ndzb_dist_layout layout{};
    std::vector<ndzb_dist_layout> peers;  // every rank's layout (static: geometry only)
    ndzb_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;       // the caller's stream (compress / decompress / gather)
    bool owns_stream = false;            // ndzb_dist_create_local makes one per rank
    cudaStream_t side = nullptr;         // high priority: count exchange + header fix-up
    cudaEvent_t ev_compressed = nullptr, ev_exchanged = nullptr;
    uint32_t *d_length = nullptr;        // this rank's stream length in words
    uint32_t *d_lengths = nullptr;       // all ranks' (all-gathered)
    uint32_t *d_overhead = nullptr;      // all ranks' non-cube words (header + border): static
    uint32_t *d_global_header = nullptr; // this rank's slice of the global header (local_cubes entries)
    uint32_t *h_lengths = nullptr;       // pinned copy of d_lengths for the gather's send / receive sizes
    bool exchange_pending = false;
    // ---- gather over peer memory (NVLink stores into the root's buffer instead of ncclSend / ncclRecv)
    bool local_group = false;            // all ranks live in this process (ndzb_dist_create_local): plain pointers are valid
    struct root_info {                   // what the root tells the other ranks about its destination buffer
        unsigned char handle[64];        // cudaIpcMemHandle_t of the allocation (mode 1)
        unsigned long long offset;       // of the buffer inside that allocation (mode 1) / the pointer itself (mode 2)
        unsigned int mode;               // 0: use NCCL send / recv, 1: CUDA IPC mapping, 2: same process
        unsigned int pad;
    };
    root_info *h_root = nullptr;         // pinned
    root_info *d_root = nullptr;
    uint32_t *d_token = nullptr;         // [1 + world]: completion all-gather behind the peer copies
    struct mapping {
        unsigned char handle[64];
        void *base;
    };
    std::vector<mapping> mappings;       // IPC mappings opened so far (closed in ndzb_dist_destroy)
};

namespace {

size_t word_bytes(int dtype) { return dtype == NDZB_F32 ? 4 : 8; }

// base address of the allocation that contains p (cuMemGetAddressRange), 0 if it cannot be determined
uintptr_t allocation_base(const void *p) {
    using fn_t = CUresult (*)(CUdeviceptr *, size_t *, CUdeviceptr);
    static fn_t fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<fn_t>(f);
    }();
    if (!fn) return 0;
    CUdeviceptr base = 0;
    size_t size = 0;
    if (fn(&base, &size, reinterpret_cast<CUdeviceptr>(p)) != CUDA_SUCCESS) return 0;
    return static_cast<uintptr_t>(base);
}

bool gather_over_nccl_forced() {
    const char *env = getenv("NDZB_GATHER");
    return env && !strcmp(env, "nccl");
}

int dist_init(ndzb_dist *d, int dtype, int dims, const uint32_t *global_size, int rank, int world, void *cuda_stream) {
    d->dtype = dtype;
    d->dims = dims;
    d->rank = rank;
    d->world = world;
    d->stream = static_cast<cudaStream_t>(cuda_stream);
    for (int i = 0; i < dims; ++i) d->global_size[i] = global_size[i];
    DIST_CUDA(cudaGetDevice(&d->device));
    d->peers.resize(world);
    for (int r = 0; r < world; ++r) DIST_NDZB(ndzb_dist_plan(dtype, dims, global_size, world, r, &d->peers[r]));
    d->layout = d->peers[rank];
    if (d->layout.global_bound_words >= (1ull << 32)) return NDZB_ERR_INVALID_ARGUMENT;  // the reference's index_type is uint32
    DIST_NDZB(ndzb_ctx_create(&d->ctx, dtype, dims, d->layout.local_cubes, cuda_stream));
    int least = 0, greatest = 0;
    DIST_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    DIST_CUDA(cudaStreamCreateWithPriority(&d->side, cudaStreamNonBlocking, greatest));
    DIST_CUDA(cudaEventCreateWithFlags(&d->ev_compressed, cudaEventDisableTiming));
    DIST_CUDA(cudaEventCreateWithFlags(&d->ev_exchanged, cudaEventDisableTiming));
    DIST_CUDA(cudaMalloc(&d->d_length, sizeof(uint32_t)));
    DIST_CUDA(cudaMalloc(&d->d_lengths, world * sizeof(uint32_t)));
    DIST_CUDA(cudaMalloc(&d->d_overhead, world * sizeof(uint32_t)));
    DIST_CUDA(cudaMalloc(&d->d_global_header, (d->layout.local_cubes ? d->layout.local_cubes : 1) * sizeof(uint32_t)));
    DIST_CUDA(cudaHostAlloc(&d->h_lengths, world * sizeof(uint32_t), cudaHostAllocDefault));
    DIST_CUDA(cudaHostAlloc(&d->h_root, sizeof(ndzb_dist::root_info), cudaHostAllocDefault));
    DIST_CUDA(cudaMalloc(&d->d_root, sizeof(ndzb_dist::root_info)));
    DIST_CUDA(cudaMalloc(&d->d_token, (1 + world) * sizeof(uint32_t)));
    DIST_CUDA(cudaMemset(d->d_token, 0, (1 + world) * sizeof(uint32_t)));
    std::vector<uint32_t> overhead(world);
    for (int r = 0; r < world; ++r) overhead[r] = d->peers[r].local_header_words + static_cast<uint32_t>(d->peers[r].local_border_words);
    DIST_CUDA(cudaMemcpy(d->d_overhead, overhead.data(), world * sizeof(uint32_t), cudaMemcpyHostToDevice));
    DIST_CUDA(cudaMemset(d->d_lengths, 0, world * sizeof(uint32_t)));
    return NDZB_OK;
}

}  // namespace

extern "C" {

const char *ndzb_dist_last_error(void) { return g_dist_error; }

int ndzb_dist_last_gather_path(const ndzb_dist *d) { return d && d->h_root && d->world > 1 ? static_cast<int>(d->h_root->mode) : 0; }

int ndzb_dist_unique_id(void *id) {
    if (!id) return NDZB_ERR_INVALID_ARGUMENT;
    if (!nccl().ok) return fail("NCCL", "libnccl.so.2 could not be loaded (set NDZB_NCCL_LIB)");
    ncclUniqueId uid;
    DIST_NCCL(nccl().GetUniqueId(&uid));
    memcpy(id, &uid, sizeof uid);
    return NDZB_OK;
}

int ndzb_dist_create(ndzb_dist **out, int dtype, int dims, const uint32_t *global_size, const void *unique_id, int rank, int world,
        void *cuda_stream) {
    if (!out) return NDZB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    ndzb_dist_layout probe;
    DIST_NDZB(ndzb_dist_plan(dtype, dims, global_size, world, rank, &probe));
    if (world > 1 && !unique_id) return NDZB_ERR_INVALID_ARGUMENT;
    if (world > 1 && !nccl().ok) return fail("NCCL", "libnccl.so.2 could not be loaded (set NDZB_NCCL_LIB)");
    ndzb_dist *d = new (std::nothrow) ndzb_dist;
    if (!d) return NDZB_ERR_ALLOC;
    int rc = dist_init(d, dtype, dims, global_size, rank, world, cuda_stream);
    if (rc == NDZB_OK && world > 1) {
        ncclUniqueId uid;
        memcpy(&uid, unique_id, sizeof uid);
        const ncclResult_t r = nccl().CommInitRank(&d->comm, world, uid, rank);
        if (r != 0) rc = fail("ncclCommInitRank", nccl().GetErrorString(r));
    }
    if (rc != NDZB_OK) {
        ndzb_dist_destroy(d);
        return rc;
    }
    *out = d;
    return NDZB_OK;
}

int ndzb_dist_create_local(ndzb_dist **out, int dtype, int dims, const uint32_t *global_size, int world, const int *devices) {
    if (!out || world < 1 || !devices) return NDZB_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < world; ++r) out[r] = nullptr;
    if (world > 1 && !nccl().ok) return fail("NCCL", "libnccl.so.2 could not be loaded (set NDZB_NCCL_LIB)");
    int rc = NDZB_OK;
    for (int r = 0; r < world && rc == NDZB_OK; ++r) {
        const on_device here(devices[r]);
        out[r] = new (std::nothrow) ndzb_dist;
        if (!out[r]) {
            rc = NDZB_ERR_ALLOC;
            break;
        }
        cudaStream_t s = nullptr;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) rc = fail("cudaStreamCreate", "failed");
        if (rc == NDZB_OK) {
            out[r]->local_group = true;
            for (int o = 0; o < world; ++o) {  // peer access for the gather over NVLink (already enabled / unsupported: fine)
                int can = 0;
                if (o != r && cudaDeviceCanAccessPeer(&can, devices[r], devices[o]) == cudaSuccess && can) {
                    if (cudaDeviceEnablePeerAccess(devices[o], 0) != cudaSuccess) cudaGetLastError();
                }
            }
            out[r]->owns_stream = true;
            out[r]->stream = s;  // (dist_init sets it again; kept here so that a failing init still releases it)
            rc = dist_init(out[r], dtype, dims, global_size, r, world, s);
        }
    }
    if (rc == NDZB_OK && world > 1) {
        std::vector<ncclComm_t> comms(world);
        const ncclResult_t r = nccl().CommInitAll(comms.data(), world, devices);
        if (r != 0) rc = fail("ncclCommInitAll", nccl().GetErrorString(r));
        else for (int i = 0; i < world; ++i) out[i]->comm = comms[i];
    }
    if (rc != NDZB_OK) {
        for (int r = 0; r < world; ++r) {
            ndzb_dist_destroy(out[r]);
            out[r] = nullptr;
        }
    }
    return rc;
}

void ndzb_dist_destroy(ndzb_dist *d) {
    if (!d) return;
    const on_device here(d->device);
    if (d->comm && nccl().ok) nccl().CommDestroy(d->comm);
    if (d->ctx) ndzb_ctx_destroy(d->ctx);
    if (d->side) cudaStreamDestroy(d->side);
    if (d->owns_stream && d->stream) cudaStreamDestroy(d->stream);
    if (d->ev_compressed) cudaEventDestroy(d->ev_compressed);
    if (d->ev_exchanged) cudaEventDestroy(d->ev_exchanged);
    if (d->d_length) cudaFree(d->d_length);
    if (d->d_lengths) cudaFree(d->d_lengths);
    if (d->d_overhead) cudaFree(d->d_overhead);
    if (d->d_global_header) cudaFree(d->d_global_header);
    if (d->h_lengths) cudaFreeHost(d->h_lengths);
    for (auto &m : d->mappings) cudaIpcCloseMemHandle(m.base);
    if (d->h_root) cudaFreeHost(d->h_root);
    if (d->d_root) cudaFree(d->d_root);
    if (d->d_token) cudaFree(d->d_token);
    delete d;
}

int ndzb_dist_layout_of(const ndzb_dist *d, int rank, ndzb_dist_layout *out) {
    if (!d || !out || rank < 0 || rank >= d->world) return NDZB_ERR_INVALID_ARGUMENT;
    *out = d->peers[rank];
    return NDZB_OK;
}

void *ndzb_dist_stream(const ndzb_dist *d) { return d ? d->stream : nullptr; }
const uint32_t *ndzb_dist_global_header(const ndzb_dist *d) { return d ? d->d_global_header : nullptr; }
const uint32_t *ndzb_dist_gathered_lengths(const ndzb_dist *d) { return d ? d->d_lengths : nullptr; }

int ndzb_dist_compress(ndzb_dist *d, const void *d_slab, void *d_local_stream, uint32_t *d_local_length) {
    if (!d) return NDZB_ERR_INVALID_ARGUMENT;
    const on_device here(d->device);
    // a previous exchange may still be reading d_length / writing d_global_header on the side stream
    if (d->exchange_pending) DIST_CUDA(cudaStreamWaitEvent(d->stream, d->ev_exchanged, 0));
    DIST_NDZB(ndzb_compress(d->ctx, d_slab, d->dims, d->layout.slab_size, d_local_stream, d->d_length));
    if (d_local_length) DIST_CUDA(cudaMemcpyAsync(d_local_length, d->d_length, sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
    DIST_CUDA(cudaEventRecord(d->ev_compressed, d->stream));
    // ---- the exchange, off the caller's stream
    DIST_CUDA(cudaStreamWaitEvent(d->side, d->ev_compressed, 0));
    if (d->world > 1) {
        DIST_NCCL(nccl().AllGather(d->d_length, d->d_lengths, 1, kNcclUint32, d->comm, d->side));
    } else {
        DIST_CUDA(cudaMemcpyAsync(d->d_lengths, d->d_length, sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->side));
    }
    if (d->layout.local_cubes) {
        // ndzb_fixup_header enqueues on the context's stream; the side stream needs its own launch: same kernel through
        // a context-free entry point
        DIST_NDZB(ndzb_fixup_header_on(d->side, static_cast<const uint32_t *>(d_local_stream), d->d_global_header, d->layout.local_cubes,
                d->d_lengths, d->d_overhead, static_cast<uint32_t>(d->rank)));
    }
    DIST_CUDA(cudaEventRecord(d->ev_exchanged, d->side));
    d->exchange_pending = true;
    return NDZB_OK;
}

int ndzb_dist_wait_exchange(ndzb_dist *d) {
    if (!d) return NDZB_ERR_INVALID_ARGUMENT;
    const on_device here(d->device);
    if (d->exchange_pending) DIST_CUDA(cudaStreamWaitEvent(d->stream, d->ev_exchanged, 0));
    return NDZB_OK;
}

int ndzb_dist_decompress(ndzb_dist *d, const void *d_local_stream, void *d_slab) {
    if (!d) return NDZB_ERR_INVALID_ARGUMENT;
    const on_device here(d->device);
    return ndzb_decompress(d->ctx, d_local_stream, d_slab, d->dims, d->layout.slab_size);
}

int ndzb_dist_gather(ndzb_dist *d, const void *d_local_stream, void *d_global_stream, int root, uint64_t *global_length_words) {
    if (!d || root < 0 || root >= d->world || !d_local_stream) return NDZB_ERR_INVALID_ARGUMENT;
    if (d->rank == root && !d_global_stream) return NDZB_ERR_INVALID_ARGUMENT;
    const on_device here(d->device);
    DIST_NDZB(ndzb_dist_wait_exchange(d));
    // The root tells everybody how to reach its buffer: over peer memory (CUDA IPC mapping, or the plain pointer when all
    // ranks share a process) or, failing that, with NCCL send / recv. One small broadcast on the stream.
    if (d->world > 1) {
        if (d->rank == root) {
            ndzb_dist::root_info info{};
            if (!gather_over_nccl_forced()) {
                if (d->local_group) {
                    info.mode = 2;
                    info.offset = reinterpret_cast<uintptr_t>(d_global_stream);
                } else {
                    const uintptr_t base = allocation_base(d_global_stream);
                    cudaIpcMemHandle_t h;
                    if (base && cudaIpcGetMemHandle(&h, reinterpret_cast<void *>(base)) == cudaSuccess) {
                        static_assert(sizeof h == sizeof info.handle, "cudaIpcMemHandle_t size");
                        memcpy(info.handle, &h, sizeof h);
                        info.offset = reinterpret_cast<uintptr_t>(d_global_stream) - base;
                        info.mode = 1;
                    } else {
                        cudaGetLastError();  // e.g. memory from a pool that has no legacy IPC handle: NCCL it is
                    }
                }
            }
            *d->h_root = info;
            DIST_CUDA(cudaMemcpyAsync(d->d_root, d->h_root, sizeof info, cudaMemcpyHostToDevice, d->stream));
        }
        DIST_NCCL(nccl().Broadcast(d->d_root, d->d_root, sizeof(ndzb_dist::root_info), kNcclUint8, root, d->comm, d->stream));
        if (d->rank != root) DIST_CUDA(cudaMemcpyAsync(d->h_root, d->d_root, sizeof(ndzb_dist::root_info), cudaMemcpyDeviceToHost, d->stream));
    }
    // the send / receive sizes have to be known on the host: one 4-byte-per-rank copy and one synchronisation
    DIST_CUDA(cudaMemcpyAsync(d->h_lengths, d->d_lengths, d->world * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
    DIST_CUDA(cudaStreamSynchronize(d->stream));
    const size_t wb = word_bytes(d->dtype);
    std::vector<uint64_t> cube_words(d->world), cube_base(d->world);
    uint64_t total_cube_words = 0;
    for (int r = 0; r < d->world; ++r) {
        const uint64_t overhead = d->peers[r].local_header_words + d->peers[r].local_border_words;
        if (d->h_lengths[r] < overhead) return fail("ndzb_dist_gather", "a rank reported a stream shorter than its header + border");
        cube_words[r] = d->h_lengths[r] - overhead;
        cube_base[r] = total_cube_words;
        total_cube_words += cube_words[r];
    }
    const uint64_t hdr_g = d->layout.global_header_words;
    const uint64_t total = hdr_g + total_cube_words + d->layout.global_border_words;
    if (total >= (1ull << 32)) return NDZB_ERR_INVALID_ARGUMENT;
    if (global_length_words) *global_length_words = total;

    const char *local = static_cast<const char *>(d_local_stream);
    char *global = static_cast<char *>(d_global_stream);
    auto header_dst = [&](int r) { return global + static_cast<size_t>(d->peers[r].cube_index_base) * sizeof(uint32_t); };
    auto cubes_dst = [&](int r) { return global + (hdr_g + cube_base[r]) * wb; };
    auto border_dst = [&](int r) { return global + (hdr_g + total_cube_words + d->peers[r].border_base) * wb; };
    auto cubes_src = [&](int r) { return local + static_cast<size_t>(d->peers[r].local_header_words) * wb; };
    auto border_src = [&](int r) { return local + (d->peers[r].local_header_words + cube_words[r]) * wb; };

    if (d->rank == root) {
        // own pieces: device-to-device copies; f64 with an odd cube count: the header's padding word (cuda_codec.inl:446-452)
        const ndzb_dist_layout &me = d->layout;
        if (me.local_cubes) DIST_CUDA(cudaMemcpyAsync(header_dst(root), d->d_global_header, me.local_cubes * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
        if (cube_words[root]) DIST_CUDA(cudaMemcpyAsync(cubes_dst(root), cubes_src(root), cube_words[root] * wb, cudaMemcpyDeviceToDevice, d->stream));
        if (me.local_border_words) DIST_CUDA(cudaMemcpyAsync(border_dst(root), border_src(root), me.local_border_words * wb, cudaMemcpyDeviceToDevice, d->stream));
        if (d->dtype == NDZB_F64 && (me.global_cubes & 1u)) {
            DIST_CUDA(cudaMemsetAsync(global + static_cast<size_t>(me.global_cubes) * sizeof(uint32_t), 0, sizeof(uint32_t), d->stream));
        }
    }
    if (d->world > 1 && d->h_root->mode != 0) {
        // ---- peer memory: every rank copies its three pieces straight to their final place in the root's buffer over NVLink
        if (d->rank != root) {
            char *remote = nullptr;
            if (d->h_root->mode == 2) {
                remote = reinterpret_cast<char *>(static_cast<uintptr_t>(d->h_root->offset));
            } else {
                void *base = nullptr;
                for (const auto &m : d->mappings) {
                    if (memcmp(m.handle, d->h_root->handle, sizeof m.handle) == 0) base = m.base;
                }
                if (!base) {
                    cudaIpcMemHandle_t h;
                    memcpy(&h, d->h_root->handle, sizeof h);
                    DIST_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
                    ndzb_dist::mapping m{};
                    memcpy(m.handle, d->h_root->handle, sizeof m.handle);
                    m.base = base;
                    d->mappings.push_back(m);
                }
                remote = static_cast<char *>(base) + d->h_root->offset;
            }
            const ndzb_dist_layout &me = d->layout;
            const int r = d->rank;
            global = remote;  // the *_dst helpers now address the root's buffer
            if (me.local_cubes) DIST_CUDA(cudaMemcpyAsync(header_dst(r), d->d_global_header, me.local_cubes * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
            if (cube_words[r]) DIST_CUDA(cudaMemcpyAsync(cubes_dst(r), cubes_src(r), cube_words[r] * wb, cudaMemcpyDeviceToDevice, d->stream));
            if (me.local_border_words) DIST_CUDA(cudaMemcpyAsync(border_dst(r), border_src(r), me.local_border_words * wb, cudaMemcpyDeviceToDevice, d->stream));
        }
        // completion: a tiny all-gather behind the copies — it finishes on the root's stream only after every rank has
        // reached it on its own stream, i.e. after that rank's copies
        DIST_NCCL(nccl().AllGather(d->d_token, d->d_token + 1, 1, kNcclUint32, d->comm, d->stream));
    } else if (d->world > 1) {
        DIST_NCCL(nccl().GroupStart());
        ncclResult_t r0 = 0;
        if (d->rank == root) {
            for (int r = 0; r < d->world && r0 == 0; ++r) {
                if (r == root) continue;
                const ndzb_dist_layout &p = d->peers[r];
                if (p.local_cubes) r0 = nccl().Recv(header_dst(r), p.local_cubes * sizeof(uint32_t), kNcclUint8, r, d->comm, d->stream);
                if (r0 == 0 && cube_words[r]) r0 = nccl().Recv(cubes_dst(r), cube_words[r] * wb, kNcclUint8, r, d->comm, d->stream);
                if (r0 == 0 && p.local_border_words) r0 = nccl().Recv(border_dst(r), p.local_border_words * wb, kNcclUint8, r, d->comm, d->stream);
            }
        } else {
            const ndzb_dist_layout &me = d->layout;
            if (me.local_cubes) r0 = nccl().Send(d->d_global_header, me.local_cubes * sizeof(uint32_t), kNcclUint8, root, d->comm, d->stream);
            if (r0 == 0 && cube_words[d->rank]) r0 = nccl().Send(cubes_src(d->rank), cube_words[d->rank] * wb, kNcclUint8, root, d->comm, d->stream);
            if (r0 == 0 && me.local_border_words) r0 = nccl().Send(border_src(d->rank), me.local_border_words * wb, kNcclUint8, root, d->comm, d->stream);
        }
        const ncclResult_t r1 = nccl().GroupEnd();
        if (r0 != 0) return fail("ncclSend/ncclRecv", nccl().GetErrorString(r0));
        if (r1 != 0) return fail("ncclGroupEnd", nccl().GetErrorString(r1));
    }
    return NDZB_OK;
}

}  // extern "C"
