// ndzb_capi.cu — the C ABI declared in include/ndzip_b200.h: context, host-side stream arithmetic,
// launch sequencing (what cuda_compressor_impl::compress / cuda_decompressor_impl::decompress /
// cuda_offloader do in the reference, src/ndzip/cuda_codec.inl:554-761).
#include "../../include/ndzip_b200.h"
#include "ndzb_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <cctype>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

using namespace ndzb;

namespace {

thread_local char g_cuda_error[256] = "no error";

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_cuda_error, sizeof g_cuda_error, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return NDZB_ERR_CUDA;
}
int driver_fail(CUresult r, const char *what) {
    snprintf(g_cuda_error, sizeof g_cuda_error, "%s: CUresult %d", what, static_cast<int>(r));
    return NDZB_ERR_CUDA;
}
#define NDZB_CUDA(call)                                          \
    do {                                                         \
        const cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
    } while (0)

constexpr uint32_t side_for(int dims) { return dims == 1 ? 4096u : dims == 2 ? 64u : 16u; }

bool valid_profile(int dtype, int dims) { return (dtype == NDZB_F32 || dtype == NDZB_F64) && dims >= 1 && dims <= 3; }

grid_geom make_geom(int dims, const uint32_t *size) {
    grid_geom g{};
    const uint32_t side = side_for(dims);
    for (int d = 0; d < 3; ++d) {
        g.n[d] = 1;
        g.cubes[d] = 1;
    }
    g.num_cubes = 1;
    for (int d = 0; d < dims; ++d) {
        g.n[3 - dims + d] = size[d];
        g.cubes[3 - dims + d] = size[d] / side;
        g.num_cubes *= size[d] / side;  // reference src/ndzip/common.hh:395-402
    }
    g.div_x = make_fastdiv(g.cubes[2]);
    g.div_y = make_fastdiv(g.cubes[1]);
    return g;
}

border_geom make_border(int dims, const uint32_t *size) {
    const uint32_t side = side_for(dims);
    uint64_t n[3] = {1, 1, 1}, in[3] = {1, 1, 1};
    for (int d = 0; d < dims; ++d) {
        n[3 - dims + d] = size[d];
        in[3 - dims + d] = static_cast<uint64_t>(size[d] / side) * side;
    }
    border_geom b{};
    b.n1 = n[1];
    b.n2 = n[2];
    b.in0 = in[0];
    b.in1 = in[1];
    b.in2 = in[2];
    b.slab_border = n[1] * n[2] - in[1] * in[2];
    b.row_border = n[2] - in[2];
    b.count = n[0] * n[1] * n[2] - in[0] * in[1] * in[2];  // reference src/ndzip/common.hh:308-317
    return b;
}

uint64_t num_elements(int dims, const uint32_t *size) {
    uint64_t n = 1;
    for (int d = 0; d < dims; ++d) n *= size[d];
    return n;
}

uint32_t header_words(int dtype, uint32_t num_cubes) {
    return dtype == NDZB_F32 ? num_cubes : (num_cubes + 1) / 2;  // reference src/ndzip/common.hh:350-352
}

size_t word_bytes(int dtype) { return dtype == NDZB_F32 ? 4 : 8; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), the SM count and the occupancy tables are per device: one
// configuration per device, filled on the first context created there.
constexpr int kMaxDevices = 64;
std::mutex g_config_mutex;
kernel_config g_config[kMaxDevices];
bool g_configured[kMaxDevices] = {};

// Makes the context's device current for the duration of an entry point (a context is bound to the device that was
// current when it was created: its scratch, streams and kernel attributes live there).
class device_guard {
  public:
    explicit device_guard(int device) {
        if (cudaGetDevice(&_previous) != cudaSuccess) _previous = -1;
        if (_previous != device && cudaSetDevice(device) == cudaSuccess) _switched = true;
    }
    ~device_guard() {
        if (_switched && _previous >= 0) cudaSetDevice(_previous);
    }
    device_guard(const device_guard &) = delete;
    device_guard &operator=(const device_guard &) = delete;

  private:
    int _previous = -1;
    bool _switched = false;
};

bool verbose() {  // reference src/ndzip/common.hh:628-631
    const char *env = getenv("NDZIP_VERBOSE");
    return env && *env;
}
// what the reference's offloader prints under NDZIP_VERBOSE (src/ndzip/cuda_codec.inl:699-703, 746-748)
void report_kernel_time(uint64_t ns) {
    if (verbose()) printf("[profile] total kernel time %.3fms\n", static_cast<double>(ns) * 1e-6);
}

}  // namespace

struct ndzb_ctx {
    int dtype = 0;
    int dims = 0;
    int device = 0;                   // the device that was current at ndzb_ctx_create; every entry point runs there
    cudaStream_t stream = nullptr;
    uint32_t desc_capacity = 0;
    uint64_t *d_desc = nullptr;       // look-back descriptors
    unsigned long long *d_blocks[2] = {nullptr, nullptr};  // block-level look-back words (one per 32 cubes, capacity as d_desc); the two
                                                            // arrays alternate between launches, each launch zeroes the other one
    int blocks_cur = 0;
    uint32_t *d_counters = nullptr;   // [0] ticket counter, [1] / [2] total compressed words of the last launch (the two
                                      // alternate, so that a chained launch never reads its base from the word it writes)
    int total_cur = 0;                // which of [1] / [2] the LAST launch wrote
    uint32_t ticket_base = 0;
    uint32_t epoch = 1;
    int forced_path = -1;             // NDZB_LOAD_PATH=tma|vec16|scalar (profiling / tests)
    int forced_store = -1;            // NDZB_STORE_PATH=tma|vec16|scalar: the decoder's output path (profiling / tests)
    bool use_dec_ws = true;           // NDZB_DECOMPRESS_KERNEL=v1: decompress_kernel also where decompress_ws_kernel applies (A/B, tests)
    uint32_t *d_watch = nullptr;      // compress_ws_kernel watchdog record (8 words)
    bool ws_check = false;            // NDZB_WS_CHECK=1: synchronise after every launch and report a raised watchdog as an error
    unsigned long long *d_stats = nullptr;  // NDZB_WS_STATS=1: role/wait cycle counters of the Stats kernel variants, printed per launch
    uint32_t ws_debug = 0;            // NDZB_WS_DEBUG: profiling aids of compress_ws_kernel (produce invalid streams)
    int dec_ctas_cap = 0;             // NDZB_DEC_CTAS=n: at most n resident decompress CTAs per SM (tuning runs)
    bool use_ws = true;               // NDZB_COMPRESS_KERNEL=v1 selects compress_kernel also for TMA-compatible inputs
    int ws_variant = 0;               // NDZB_WS_VARIANT=n: (encoder groups, retire warps) tuning variants
    uint32_t last_launches = 0;
    // host-pointer ("offloader") staging, grown on demand and kept across calls
    void *d_in = nullptr;
    size_t d_in_bytes = 0;
    void *d_out = nullptr;
    size_t d_out_bytes = 0;
    uint32_t *d_length = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // pipelined offload: copy-in / copy-out streams, per-chunk events, pinned per-chunk totals
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_done, ev_k0, ev_k1;
    uint32_t *h_totals = nullptr;  // pinned + mapped: the compress kernels store their running totals here themselves
    uint32_t *d_totals = nullptr;  // device-side address of h_totals
};

namespace {

int ensure_descriptors(ndzb_ctx *ctx, uint32_t cubes) {
    if (cubes <= ctx->desc_capacity) return NDZB_OK;
    // Growing synchronises the device (cudaFree/cudaMalloc); contexts sized by
    // compressor_requirements never get here.
    if (ctx->d_desc) NDZB_CUDA(cudaFree(ctx->d_desc));
    for (auto &b : ctx->d_blocks) {
        if (b) NDZB_CUDA(cudaFree(b));
        b = nullptr;
        const size_t bytes = (static_cast<size_t>(cubes) / 32 + 1) * kDescStride * sizeof(unsigned long long);
        NDZB_CUDA(cudaMalloc(&b, bytes));
        NDZB_CUDA(cudaMemsetAsync(b, 0, bytes, ctx->stream));
    }
    ctx->d_desc = nullptr;
    ctx->desc_capacity = 0;
    NDZB_CUDA(cudaMalloc(&ctx->d_desc, static_cast<size_t>(cubes) * kDescStride * sizeof(uint64_t)));
    NDZB_CUDA(cudaMemsetAsync(ctx->d_desc, 0, static_cast<size_t>(cubes) * kDescStride * sizeof(uint64_t), ctx->stream));
    ctx->desc_capacity = cubes;
    return NDZB_OK;
}

int ensure_buffer(void **buf, size_t *have, size_t want) {
    if (want <= *have) return NDZB_OK;
    if (*buf) NDZB_CUDA(cudaFree(*buf));
    *buf = nullptr;
    *have = 0;
    NDZB_CUDA(cudaMalloc(buf, want));
    *have = want;
    return NDZB_OK;
}

load_path choose_path(const ndzb_ctx *ctx, const void *data, const grid_geom &g) {
    const bool aligned = tma_compatible(ctx->dtype, ctx->dims, data, g);
    if (ctx->forced_path == static_cast<int>(load_path::scalar)) return load_path::scalar;
    if (!aligned) return load_path::scalar;
    if (ctx->forced_path == static_cast<int>(load_path::vec16)) return load_path::vec16;
    return load_path::tma;
}

// After a launch that failed or was abandoned by the watchdog the device-side scan state (ticket counter, totals,
// look-back descriptors, block words) no longer matches the host's book-keeping: start over from a clean slate.
int reset_scan_state(ndzb_ctx *ctx) {
    NDZB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(uint32_t), ctx->stream));
    if (ctx->d_desc) NDZB_CUDA(cudaMemsetAsync(ctx->d_desc, 0, static_cast<size_t>(ctx->desc_capacity) * kDescStride * sizeof(uint64_t), ctx->stream));
    for (auto b : ctx->d_blocks) {
        if (b) NDZB_CUDA(cudaMemsetAsync(b, 0, (static_cast<size_t>(ctx->desc_capacity) / 32 + 1) * kDescStride * sizeof(unsigned long long), ctx->stream));
    }
    NDZB_CUDA(cudaMemsetAsync(ctx->d_watch, 0, 7 * sizeof(uint32_t), ctx->stream));  // keeps [7], the mode
    ctx->ticket_base = 0;
    ctx->epoch = 1;
    ctx->blocks_cur = 0;
    ctx->total_cur = 0;
    return NDZB_OK;
}

// Enqueue the compression of cubes [hc_begin, hc_begin + count) (count > 0).
int enqueue_compress_range(ndzb_ctx *ctx, const void *d_data, const grid_geom &g, uint32_t hc_begin, uint32_t count,
        void *out_cubes, uint32_t *out_offsets, uint32_t *pad_word, uint32_t *length_out, uint32_t length_add,
        bool chained = false, uint32_t *total_host = nullptr) {
    if (int rc = ensure_descriptors(ctx, count)) return rc;
    const load_path path = choose_path(ctx, d_data, g);
    CUtensorMap map{};
    if (path == load_path::tma) {
        const CUresult r = make_input_tensor_map(&map, ctx->dtype, ctx->dims, d_data, g);
        if (r != CUDA_SUCCESS) return driver_fail(r, "cuTensorMapEncodeTiled");
    }
    const bool ws = path == load_path::tma && ctx->use_ws;
    const int per_sm = ws ? 1 : g_config[ctx->device].ctas_per_sm[ctx->dtype][ctx->dims - 1][static_cast<int>(path)];
    const uint64_t resident = static_cast<uint64_t>(per_sm > 0 ? per_sm : 1) * g_config[ctx->device].num_sms;
    const uint32_t grid = static_cast<uint32_t>(count < resident ? count : resident);

    compress_launch a{};
    a.data = d_data;
    a.geom = g;
    a.hc_begin = hc_begin;
    a.count = count;
    a.out_cubes = out_cubes;
    a.out_offsets = out_offsets;
    a.pad_word = pad_word;
    // a chained launch starts where the previous one ended: it reads that total from one scalar and writes its own
    // to the other, so the read never races with the write of the launch's last cube
    a.base_words = chained ? ctx->d_counters + 1 + ctx->total_cur : nullptr;
    const int total_next = ctx->total_cur ^ 1;   // (the host's book-keeping only advances once the launch has succeeded)
    a.total_words = ctx->d_counters + 1 + total_next;
    a.total_host = total_host;
    a.length_out = length_out;
    a.length_add = length_add;
    a.desc = ctx->d_desc;
    a.ticket = ctx->d_counters;
    a.ticket_base = ctx->ticket_base;
    a.epoch = ctx->epoch;
    a.watch = ctx->d_watch;
    a.debug_flags = ctx->ws_debug;
    a.stats = ctx->d_stats;
    if (ws) {
        if (compress_ws_uses_blocks(ctx->dtype, ctx->ws_variant)) {
            // this launch accumulates into one array (all zero: zeroed at allocation or by the previous launch) and zeroes
            // the other one for the next launch, whatever its cube count will be
            a.block_desc = ctx->d_blocks[ctx->blocks_cur];
            a.block_desc_next = ctx->d_blocks[ctx->blocks_cur ^ 1];
            a.block_words_next = ctx->desc_capacity / 32 + 1;
        }
        const cudaError_t e = launch_compress_ws(ctx->dtype, ctx->dims, ctx->ws_variant, a, map, grid, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(e, "compress_ws_kernel launch");
        if (a.block_desc) ctx->blocks_cur ^= 1;
        ctx->total_cur = total_next;
        ctx->ticket_base += count + compress_ws_ticket_overdraw(ctx->dtype, ctx->dims, ctx->ws_variant, grid, count);  // wraps together with the device counter
        if (ctx->d_stats) {
            unsigned long long h[16];
            NDZB_CUDA(cudaStreamSynchronize(ctx->stream));
            NDZB_CUDA(cudaMemcpy(h, ctx->d_stats, sizeof h, cudaMemcpyDeviceToHost));
            NDZB_CUDA(cudaMemset(ctx->d_stats, 0, sizeof h));
            const double enc = h[3] ? double(h[3]) : 1.0, ret = h[11] ? double(h[11]) : 1.0, g = grid;
            fprintf(stderr, "ws stats (cycles per cube) enc: wait_tile %.0f phase1 %.0f phase2 %.0f | loader per cube: wait_slot %.0f wait_encoders %.0f total %.0f | "
                    "retire: wait_cube %.0f look_back %.0f wait_image %.0f copy %.0f polls/cube %.2f extra_windows/cube %.2f\n",
                    h[0] / enc, h[1] / enc, h[2] / enc, h[4] / ret, h[5] / ret, h[6] / ret, h[8] / ret, h[9] / ret, h[14] / ret, h[10] / ret, h[12] / ret, h[13] / ret);
            fprintf(stderr, "ws stats (cycles per cube) slot: freed->load issued %.0f, load issued->encoder starts %.0f\n", h[15] / ret, h[7] / enc);
            (void) g;
        }
        if (ctx->ws_check) {
            std::vector<uint32_t> w(kWatchdogWords, 0u);
            NDZB_CUDA(cudaStreamSynchronize(ctx->stream));
            NDZB_CUDA(cudaMemcpy(w.data(), ctx->d_watch, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            if (w[0] != 0) {
                NDZB_CUDA(cudaMemcpy(w.data(), ctx->d_watch, kWatchdogWords * sizeof(uint32_t), cudaMemcpyDeviceToHost));
                snprintf(g_cuda_error, sizeof g_cuda_error, "compress_ws_kernel watchdog: first code 0x%x, %u look-back + %u other waits abandoned (count %u grid %u variant %d)",
                        w[0], w[1], w[2], count, grid, ctx->ws_variant);
                fprintf(stderr, "%s\n", g_cuda_error);
                if (const char *path = getenv("NDZB_WS_WATCH_DUMP")) {  // all records, one per line, for offline analysis
                    if (FILE *f = fopen(path, "w")) {
                        for (uint32_t i = 0; i < 4000; ++i) {
                            const uint32_t *r = w.data() + 8 + 6 * i;
                            if (i < 8 ? i >= w[1] : i - 8 >= w[2]) continue;
                            fprintf(f, "0x%04x %u %u %u %u %u\n", r[0], r[1], r[2], r[3], r[4], r[5]);
                        }
                        fclose(f);
                    }
                }
                if (int rc = reset_scan_state(ctx)) return rc;  // the abandoned launch left tickets, descriptors and totals half-written
                return NDZB_ERR_CUDA;
            }
        }
    } else {
        const cudaError_t e = launch_compress(ctx->dtype, ctx->dims, path, a, &map, grid, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(e, "compress_kernel launch");
        ctx->total_cur = total_next;
        ctx->ticket_base += count + compress_ticket_overdraw(grid);  // wraps together with the device counter
    }
    ctx->last_launches += 1;
    if (++ctx->epoch >= (1u << 30)) {
        NDZB_CUDA(cudaMemsetAsync(ctx->d_desc, 0, static_cast<size_t>(ctx->desc_capacity) * kDescStride * sizeof(uint64_t), ctx->stream));
        ctx->epoch = 1;
    }
    return NDZB_OK;
}

bool store_vectorisable(const ndzb_ctx *ctx, const void *data, const grid_geom &g) {
    return ctx->forced_path != static_cast<int>(load_path::scalar) && tma_compatible(ctx->dtype, ctx->dims, data, g);
}

int enqueue_decompress_range(ndzb_ctx *ctx, const void *stream_cubes, const uint32_t *offsets, void *d_data,
        const grid_geom &g, uint32_t hc_begin, uint32_t count) {
    // output path: TMA tensor store of the decoded tile where the output is TMA-addressable (16-byte aligned base and
    // pitch), else element-wise stores; NDZB_STORE_PATH=tma|vec16|scalar overrides (profiling / tests)
    int store = store_vectorisable(ctx, d_data, g) ? 2 : 0;
    if (store == 2 && ctx->forced_store >= 0 && ctx->forced_store < 2) store = ctx->forced_store;
    CUtensorMap out_map{};
    if (store == 2) {
        const CUresult r = make_output_tensor_map(&out_map, ctx->dtype, ctx->dims, d_data, g);
        if (r != CUDA_SUCCESS) return driver_fail(r, "cuTensorMapEncodeTiled (output)");
    }
    if (store == 2 && ctx->use_dec_ws && decompress_ws_available(ctx->dtype, ctx->dims)) {
        const uint32_t sms = static_cast<uint32_t>(g_config[ctx->device].num_sms);
        decompress_launch a{};
        a.stream_cubes = stream_cubes;
        a.offsets = offsets;
        a.data = d_data;
        a.geom = g;
        a.hc_begin = hc_begin;
        a.count = count;
        const cudaError_t e = launch_decompress_ws(ctx->dims, a, out_map, count < sms ? count : sms, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(e, "decompress_ws_kernel launch");
        ctx->last_launches += 1;
        return NDZB_OK;
    }
    int per_sm = g_config[ctx->device].dec_ctas_per_sm[ctx->dtype][ctx->dims - 1][store];
    if (ctx->dec_ctas_cap > 0 && per_sm > ctx->dec_ctas_cap) per_sm = ctx->dec_ctas_cap;  // NDZB_DEC_CTAS (tuning)
    const uint64_t resident = static_cast<uint64_t>(per_sm > 0 ? per_sm : 1) * g_config[ctx->device].num_sms;
    const uint32_t grid = static_cast<uint32_t>(count < resident ? count : resident);
    decompress_launch a{};
    a.stream_cubes = stream_cubes;
    a.offsets = offsets;
    a.data = d_data;
    a.geom = g;
    a.hc_begin = hc_begin;
    a.count = count;
    const cudaError_t e = launch_decompress(ctx->dtype, ctx->dims, store, a, store == 2 ? &out_map : nullptr, grid, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "decompress_kernel launch");
    ctx->last_launches += 1;
    return NDZB_OK;
}

// ---- pipelined host<->device path -------------------------------------------------
// The grid is cut into slabs of whole cube rows along dimension 0. H2D of slab c+1, the kernels of slab c
// and D2H of slab c-1 run on three streams, so the call is bounded by max(H2D, D2H) over PCIe instead of
// their sum. Used for border-free extents above a size threshold; everything else takes the simple path.
constexpr int kMaxChunks = 128;
constexpr size_t kChunkBytes = size_t{32} << 20;            // compression: the tail after the last H2D copy wants small chunks
constexpr size_t kDecompressChunkBytes = size_t{64} << 20;  // decompression: the D2H copies (the long pole) want few, large copies
constexpr size_t kPipelineMinBytes = size_t{16} << 20;

size_t pipeline_min_bytes() {
    if (const char *env = getenv("NDZB_PIPELINE_MIN_BYTES")) return strtoull(env, nullptr, 10);
    return kPipelineMinBytes;
}

struct chunk_plan {
    int chunks = 0;
    std::vector<uint32_t> row_begin;  // chunk c = cube rows [row_begin[c], row_begin[c + 1]) along dimension 0
    uint32_t cube_rows = 0;
    uint32_t cubes_per_row = 0;
    uint64_t elems_per_cube_row = 0;
};

// `small_first`: the chunks at the FRONT are single cube rows (decompression: the D2H copies are the long pole and
// cannot start before the first chunk has been uploaded and decoded); otherwise the chunks at the BACK are
// (compression: after the last H2D copy only that chunk's kernel and its D2H copy remain).
chunk_plan plan_chunks(int dims, const uint32_t *size, const grid_geom &g, size_t elem_bytes, bool small_first) {
    chunk_plan p;
    p.cube_rows = size[0] / side_for(dims);
    if (p.cube_rows == 0) return p;
    p.cubes_per_row = g.num_cubes / p.cube_rows;
    p.elems_per_cube_row = static_cast<uint64_t>(side_for(dims));
    for (int d = 1; d < dims; ++d) p.elems_per_cube_row *= size[d];
    const uint64_t row_bytes = p.elems_per_cube_row * elem_bytes;
    uint64_t chunk_bytes = small_first ? kDecompressChunkBytes : kChunkBytes;
    // small arrays: at least ~8 chunks so that the three streams overlap at all (64 MiB in two 32 MiB chunks is half serial),
    // but no chunk below 4 MiB (per-chunk launch + event cost)
    const uint64_t total_bytes = row_bytes * p.cube_rows;
    if (total_bytes / 8 < chunk_bytes) chunk_bytes = total_bytes / 8 > (uint64_t{4} << 20) ? total_bytes / 8 : (uint64_t{4} << 20);
    if (const char *env = getenv("NDZB_CHUNK_BYTES")) chunk_bytes = strtoull(env, nullptr, 10);  // tests / tuning
    uint64_t rows = chunk_bytes / (row_bytes ? row_bytes : 1);
    if (rows == 0) rows = 1;
    uint64_t chunks = (p.cube_rows + rows - 1) / rows;
    if (chunks > kMaxChunks - 4) {
        rows = (p.cube_rows + (kMaxChunks - 4) - 1) / (kMaxChunks - 4);
        chunks = (p.cube_rows + rows - 1) / rows;
    }
    // ramp: one nominal chunk at the small end is cut into (at most 4) pieces of rows/4
    std::vector<uint32_t> lens;
    uint64_t left = p.cube_rows;
    const uint64_t piece = (rows + 3) / 4;
    uint64_t ramp = rows > 1 && chunks > 1 ? rows : 0;
    while (ramp > 0 && left > 0) {
        const uint64_t n = ramp < piece ? ramp : piece;
        lens.push_back(static_cast<uint32_t>(n < left ? n : left));
        left -= lens.back();
        ramp -= n;
    }
    while (left > 0) {
        lens.push_back(static_cast<uint32_t>(rows < left ? rows : left));
        left -= lens.back();
    }
    // built small-first; a remainder chunk (cube_rows not a multiple of `rows`) ends up at the back either way
    if (!small_first) std::reverse(lens.begin(), lens.end());
    p.row_begin.push_back(0);
    for (uint32_t n : lens) p.row_begin.push_back(p.row_begin.back() + n);
    p.chunks = static_cast<int>(lens.size());
    return p;
}

int ensure_pipeline(ndzb_ctx *ctx, int chunks) {
    if (!ctx->s_in) NDZB_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) NDZB_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    if (!ctx->h_totals) {
        NDZB_CUDA(cudaHostAlloc(&ctx->h_totals, kMaxChunks * sizeof(uint32_t), cudaHostAllocMapped));
        NDZB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&ctx->d_totals), ctx->h_totals, 0));
    }
    while (static_cast<int>(ctx->ev_in.size()) < chunks) {
        cudaEvent_t a, b, c, d;
        NDZB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        NDZB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        NDZB_CUDA(cudaEventCreate(&c));
        NDZB_CUDA(cudaEventCreate(&d));
        ctx->ev_in.push_back(a);
        ctx->ev_done.push_back(b);
        ctx->ev_k0.push_back(c);
        ctx->ev_k1.push_back(d);
    }
    return NDZB_OK;
}

int sum_kernel_time(ndzb_ctx *ctx, int chunks, uint64_t *kernel_ns) {
    if (!kernel_ns) return NDZB_OK;
    double total_ms = 0;
    for (int c = 0; c < chunks; ++c) {
        float ms = 0;
        NDZB_CUDA(cudaEventElapsedTime(&ms, ctx->ev_k0[c], ctx->ev_k1[c]));
        total_ms += ms;
    }
    *kernel_ns = static_cast<uint64_t>(total_ms * 1e6);
    return NDZB_OK;
}

// NDZB_PIPE_TRACE=1: timeline of one pipelined call on stderr (ms since the call's first enqueue): per chunk the end
// of its H2D copy, start / end of its kernel, end of its D2H copy, and the host clock at every synchronisation.
struct pipe_trace {
    bool on = false;
    cudaEvent_t t0 = nullptr;
    std::vector<cudaEvent_t> in, out;
    std::vector<double> host_sync;
    std::chrono::steady_clock::time_point h0;
    explicit pipe_trace(int chunks) {
        const char *e = getenv("NDZB_PIPE_TRACE");
        on = e && atoi(e) != 0;
        if (!on) return;
        cudaEventCreate(&t0);
        in.resize(chunks);
        out.resize(chunks, nullptr);
        for (auto &x : in) cudaEventCreate(&x);
        h0 = std::chrono::steady_clock::now();
    }
    double host_ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count(); }
    void begin(cudaStream_t s) { if (on) cudaEventRecord(t0, s); }
    void mark_in(int c, cudaStream_t s) { if (on) cudaEventRecord(in[c], s); }
    void mark_out(int c, cudaStream_t s) {
        if (!on) return;
        if (!out[c]) cudaEventCreate(&out[c]);
        cudaEventRecord(out[c], s);
    }
    void mark_host() { if (on) host_sync.push_back(host_ms()); }
    void report(const char *what, const std::vector<cudaEvent_t> &k0, const std::vector<cudaEvent_t> &k1) {
        if (!on) return;
        const double end = host_ms();
        fprintf(stderr, "pipe trace %s: %zu chunks, host total %.3f ms, host syncs at", what, in.size(), end);
        for (double h : host_sync) fprintf(stderr, " %.3f", h);
        fprintf(stderr, "\n  chunk: h2d_end k_begin k_end d2h_end (ms since first enqueue on the copy-in stream)\n");
        for (size_t c = 0; c < in.size(); ++c) {
            float a = 0, b = 0, d = 0, e = -1;
            cudaEventElapsedTime(&a, t0, in[c]);
            cudaEventElapsedTime(&b, t0, k0[c]);
            cudaEventElapsedTime(&d, t0, k1[c]);
            if (out[c]) cudaEventElapsedTime(&e, t0, out[c]);
            fprintf(stderr, "  %3zu: %8.3f %8.3f %8.3f %8.3f\n", c, a, b, d, e);
        }
        cudaEventDestroy(t0);
        for (auto x : in) cudaEventDestroy(x);
        for (auto x : out) if (x) cudaEventDestroy(x);
    }
};

int pipelined_compress(ndzb_ctx *ctx, const void *h_data, int dims, const uint32_t *size, void *h_stream,
        uint32_t *length_words, uint64_t *kernel_ns, const grid_geom &g, const chunk_plan &plan) {
    const size_t wb = word_bytes(ctx->dtype);
    const uint32_t H = g.num_cubes;
    const uint32_t hdr = header_words(ctx->dtype, H);
    const uint64_t n_elems = num_elements(dims, size);
    if (int rc = ensure_pipeline(ctx, plan.chunks)) return rc;
    if (int rc = ensure_descriptors(ctx, H)) return rc;  // before anything is enqueued: growing synchronises
    char *d_in = static_cast<char *>(ctx->d_in);
    char *d_out = static_cast<char *>(ctx->d_out);
    const char *h_in = static_cast<const char *>(h_data);
    char *h_out = static_cast<char *>(h_stream);
    uint32_t *offsets = static_cast<uint32_t *>(ctx->d_out);
    void *cubes = d_out + static_cast<size_t>(hdr) * wb;
    uint32_t *pad = (ctx->dtype == NDZB_F64 && (H & 1u)) ? offsets + H : nullptr;

    // whatever the caller enqueued on the context's stream comes first
    NDZB_CUDA(cudaEventRecord(ctx->ev_begin, ctx->stream));
    NDZB_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->ev_begin, 0));
    NDZB_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_begin, 0));
    pipe_trace trace(plan.chunks);
    trace.begin(ctx->s_in);

    for (int c = 0; c < plan.chunks; ++c) {
        const uint32_t row0 = plan.row_begin[c], row1 = plan.row_begin[c + 1];
        const uint64_t e0 = row0 * plan.elems_per_cube_row;
        const uint64_t e1 = (c + 1 == plan.chunks) ? n_elems : row1 * plan.elems_per_cube_row;
        NDZB_CUDA(cudaMemcpyAsync(d_in + e0 * wb, h_in + e0 * wb, (e1 - e0) * wb, cudaMemcpyHostToDevice, ctx->s_in));
        trace.mark_in(c, ctx->s_in);
        NDZB_CUDA(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
        NDZB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[c], 0));
        NDZB_CUDA(cudaEventRecord(ctx->ev_k0[c], ctx->stream));
        const uint32_t hb = row0 * plan.cubes_per_row, he = row1 * plan.cubes_per_row;
        if (int rc = enqueue_compress_range(ctx, ctx->d_in, g, hb, he - hb, cubes, offsets + hb, c == 0 ? pad : nullptr,
                    c + 1 == plan.chunks ? ctx->d_length : nullptr, hdr, c != 0, ctx->d_totals + c)) {
            return rc;
        }
        NDZB_CUDA(cudaEventRecord(ctx->ev_k1[c], ctx->stream));
        NDZB_CUDA(cudaEventRecord(ctx->ev_done[c], ctx->stream));
    }
    // drain: the compressed cubes of chunk c go back on the copy-out stream as soon as its kernel has finished. The size
    // of that copy is only known on the device; the kernel stores its running total straight into mapped host memory
    // (total_host), so the compute stream holds nothing but kernels. (A 4-byte D2H copy node per chunk in that stream
    // queued behind the large copies of the copy-out stream on the same copy engine and stalled every later kernel:
    // profiles/README.md, round 2, "offloader timeline".) One host synchronisation per chunk; the host has nothing else to do.
    // Every copy but the last ends on a 4 KiB boundary of the stream (the few words beyond it travel with the next
    // chunk), so that all copies start aligned on both sides.
    uint32_t prev = 0;
    size_t sent = static_cast<size_t>(hdr) * wb;  // bytes of the stream already on their way (the header goes last)
    for (int c = 0; c < plan.chunks; ++c) {
        NDZB_CUDA(cudaEventSynchronize(ctx->ev_done[c]));
        trace.mark_host();
        const uint32_t tot = *static_cast<volatile uint32_t *>(ctx->h_totals + c);
        size_t end = (static_cast<size_t>(hdr) + tot) * wb;
        if (c + 1 < plan.chunks) end = end / 4096 * 4096;
        if (end > sent) {
            NDZB_CUDA(cudaMemcpyAsync(h_out + sent, d_out + sent, end - sent, cudaMemcpyDeviceToHost, ctx->s_out));
            sent = end;
        }
        trace.mark_out(c, ctx->s_out);
        prev = tot;
    }
    NDZB_CUDA(cudaMemcpyAsync(h_out, d_out, static_cast<size_t>(hdr) * wb, cudaMemcpyDeviceToHost, ctx->s_out));
    NDZB_CUDA(cudaStreamSynchronize(ctx->s_out));
    NDZB_CUDA(cudaStreamSynchronize(ctx->stream));
    trace.mark_host();
    trace.report("compress", ctx->ev_k0, ctx->ev_k1);
    *length_words = hdr + prev;
    return sum_kernel_time(ctx, plan.chunks, kernel_ns);
}

int pipelined_decompress(ndzb_ctx *ctx, const void *h_stream, void *h_data, int dims, const uint32_t *size,
        uint64_t *kernel_ns, const grid_geom &g, const chunk_plan &plan) {
    const size_t wb = word_bytes(ctx->dtype);
    const uint32_t H = g.num_cubes;
    const uint32_t hdr = header_words(ctx->dtype, H);
    const uint64_t n_elems = num_elements(dims, size);
    if (int rc = ensure_pipeline(ctx, plan.chunks)) return rc;
    char *d_stream = static_cast<char *>(ctx->d_out);
    char *d_data = static_cast<char *>(ctx->d_in);
    const char *h_in = static_cast<const char *>(h_stream);
    const uint32_t *h_offsets = static_cast<const uint32_t *>(h_stream);
    char *h_out = static_cast<char *>(h_data);

    NDZB_CUDA(cudaEventRecord(ctx->ev_begin, ctx->stream));
    NDZB_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->ev_begin, 0));
    NDZB_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_begin, 0));
    pipe_trace trace(plan.chunks);
    trace.begin(ctx->s_in);
    NDZB_CUDA(cudaMemcpyAsync(d_stream, h_in, static_cast<size_t>(hdr) * wb, cudaMemcpyHostToDevice, ctx->s_in));

    for (int c = 0; c < plan.chunks; ++c) {
        const uint32_t row0 = plan.row_begin[c], row1 = plan.row_begin[c + 1];
        const uint32_t hb = row0 * plan.cubes_per_row, he = row1 * plan.cubes_per_row;
        const size_t w0 = static_cast<size_t>(hdr) + (hb ? h_offsets[hb - 1] : 0u);
        const size_t w1 = static_cast<size_t>(hdr) + h_offsets[he - 1];
        // The chunk's compressed cubes start and end at arbitrary words. The copy is widened to 4 KiB boundaries of the
        // stream (re-sending a few bytes of the neighbouring chunks, which hold the same data on both sides): copies
        // that start in the middle of a cache line move over PCIe at ~41 instead of ~50 GB/s next to a concurrent D2H
        // (profiles/README.md, round 2, "offloader timeline").
        constexpr size_t kAlign = 4096;
        const size_t total_bytes = (static_cast<size_t>(hdr) + h_offsets[H - 1]) * wb;
        size_t b0 = w0 * wb / kAlign * kAlign, b1 = (w1 * wb + kAlign - 1) / kAlign * kAlign;
        if (b1 > total_bytes) b1 = total_bytes;
        if (b1 > b0) NDZB_CUDA(cudaMemcpyAsync(d_stream + b0, h_in + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->s_in));
        trace.mark_in(c, ctx->s_in);
        NDZB_CUDA(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
        NDZB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[c], 0));
        NDZB_CUDA(cudaEventRecord(ctx->ev_k0[c], ctx->stream));
        if (int rc = enqueue_decompress_range(ctx, d_stream + static_cast<size_t>(hdr) * wb, reinterpret_cast<const uint32_t *>(d_stream),
                    ctx->d_in, g, hb, he - hb)) {
            return rc;
        }
        NDZB_CUDA(cudaEventRecord(ctx->ev_k1[c], ctx->stream));
        NDZB_CUDA(cudaEventRecord(ctx->ev_done[c], ctx->stream));
        NDZB_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[c], 0));
        const uint64_t e0 = row0 * plan.elems_per_cube_row;
        const uint64_t e1 = (c + 1 == plan.chunks) ? n_elems : row1 * plan.elems_per_cube_row;
        NDZB_CUDA(cudaMemcpyAsync(h_out + e0 * wb, d_data + e0 * wb, (e1 - e0) * wb, cudaMemcpyDeviceToHost, ctx->s_out));
        trace.mark_out(c, ctx->s_out);
    }
    trace.mark_host();
    NDZB_CUDA(cudaStreamSynchronize(ctx->s_out));
    NDZB_CUDA(cudaStreamSynchronize(ctx->stream));
    trace.mark_host();
    trace.report("decompress", ctx->ev_k0, ctx->ev_k1);
    return sum_kernel_time(ctx, plan.chunks, kernel_ns);
}

int check_call(const ndzb_ctx *ctx, int dims, const uint32_t *size) {
    if (!ctx || !size) return NDZB_ERR_INVALID_ARGUMENT;
    if (dims != ctx->dims) return NDZB_ERR_DIMS_MISMATCH;
    if (num_elements(dims, size) >= (1ull << 32)) return NDZB_ERR_INVALID_ARGUMENT;  // reference index_type is uint32
    return NDZB_OK;
}

}  // namespace

extern "C" {

int ndzb_ctx_create(ndzb_ctx **out_ctx, int dtype, int dims, uint32_t max_hypercubes, void *cuda_stream) {
    if (!out_ctx || !valid_profile(dtype, dims)) return NDZB_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int device = 0;
    {
        const cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice (is a CUDA device present?)");
        if (device < 0 || device >= kMaxDevices) return NDZB_ERR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lock(g_config_mutex);
        if (!g_configured[device]) {
            const cudaError_t c = configure_kernels(g_config[device]);
            if (c != cudaSuccess) return cuda_fail(c, "configure_kernels (is a CUDA device present?)");
            g_configured[device] = true;
        }
    }
    ndzb_ctx *ctx = new (std::nothrow) ndzb_ctx;
    if (!ctx) return NDZB_ERR_ALLOC;
    ctx->dtype = dtype;
    ctx->dims = dims;
    ctx->device = device;
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    if (const char *p = getenv("NDZB_LOAD_PATH")) {
        if (!strcmp(p, "tma")) ctx->forced_path = 0;
        else if (!strcmp(p, "vec16")) ctx->forced_path = 1;
        else if (!strcmp(p, "scalar")) ctx->forced_path = 2;
    }
    if (const char *p = getenv("NDZB_DECOMPRESS_KERNEL")) ctx->use_dec_ws = strcmp(p, "v1") != 0;
    if (const char *p = getenv("NDZB_STORE_PATH")) {
        if (!strcmp(p, "tma")) ctx->forced_store = 2;
        else if (!strcmp(p, "vec16")) ctx->forced_store = 1;
        else if (!strcmp(p, "scalar")) ctx->forced_store = 0;
    }
    if (tuning_build()) {  // -DNDZB_TUNING builds only: A/B kernels and variants
        if (const char *p = getenv("NDZB_COMPRESS_KERNEL")) ctx->use_ws = strcmp(p, "v1") != 0;
        if (const char *p = getenv("NDZB_WS_VARIANT")) ctx->ws_variant = atoi(p);
    }
    if (const char *p = getenv("NDZB_DEC_CTAS")) ctx->dec_ctas_cap = atoi(p);
    auto fail = [&](int rc) {
        ndzb_ctx_destroy(ctx);
        return rc;
    };
    cudaError_t e = cudaMalloc(&ctx->d_counters, 4 * sizeof(uint32_t));
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc counters"));
    e = cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(uint32_t), ctx->stream);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMemsetAsync counters"));
    if (const char *p = getenv("NDZB_WS_CHECK")) ctx->ws_check = atoi(p) != 0;
    if (const char *p = getenv("NDZB_WS_DEBUG")) ctx->ws_debug = tuning_build() ? static_cast<uint32_t>(atoi(p)) : 0u;
    if (const char *p = getenv("NDZB_WS_STATS")) {
        if (tuning_build() && atoi(p) != 0) {
            e = cudaMalloc(&ctx->d_stats, 16 * sizeof(unsigned long long));
            if (e == cudaSuccess) e = cudaMemset(ctx->d_stats, 0, 16 * sizeof(unsigned long long));
            if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc stats"));
        }
    }
    e = cudaMalloc(&ctx->d_watch, kWatchdogWords * sizeof(uint32_t));
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc watchdog"));
    {
        e = cudaMemsetAsync(ctx->d_watch, 0, kWatchdogWords * sizeof(uint32_t), ctx->stream);
        const uint32_t mode = ctx->ws_check ? 1u : 0u;
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_watch + 7, &mode, sizeof mode, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // `mode` is on the stack
        if (e != cudaSuccess) return fail(cuda_fail(e, "watchdog init"));
    }
    if (max_hypercubes) {
        if (int rc = ensure_descriptors(ctx, max_hypercubes)) return fail(rc);
    }
    e = cudaEventCreate(&ctx->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_end);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaEventCreate"));
    *out_ctx = ctx;
    return NDZB_OK;
}

int ndzb_offload_chunk_plan(int dtype, int dims, const uint32_t *size, int decompress, uint32_t *row_begin, uint32_t capacity, uint32_t *chunks) {
    if (!valid_profile(dtype, dims) || !size || !chunks) return NDZB_ERR_INVALID_ARGUMENT;
    const grid_geom g = make_geom(dims, size);
    *chunks = 0;
    // the same conditions as ndzb_offload_compress / ndzb_offload_decompress: no border, above the size threshold
    if (g.num_cubes == 0 || make_border(dims, size).count != 0 || num_elements(dims, size) * word_bytes(dtype) < pipeline_min_bytes()) return NDZB_OK;
    const chunk_plan plan = plan_chunks(dims, size, g, word_bytes(dtype), decompress != 0);
    if (plan.chunks <= 1) return NDZB_OK;
    *chunks = static_cast<uint32_t>(plan.chunks);
    if (row_begin) {
        if (capacity < static_cast<uint32_t>(plan.chunks) + 1) return NDZB_ERR_CAPACITY;
        for (int c = 0; c <= plan.chunks; ++c) row_begin[c] = plan.row_begin[c];
    }
    return NDZB_OK;
}

int ndzb_host_alloc(void **out_ptr, size_t bytes) {
    if (!out_ptr) return NDZB_ERR_INVALID_ARGUMENT;
    *out_ptr = nullptr;
    NDZB_CUDA(cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return NDZB_OK;
}

void ndzb_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

// NUMA node of a CUDA device from sysfs (/sys/bus/pci/devices/<domain:bus:dev.fn>/numa_node); -1 if unknown.
static int device_numa_node(int device) {
    char bus_id[32] = {0};
    if (cudaDeviceGetPCIBusId(bus_id, sizeof bus_id, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bus_id; *c; ++c) *c = static_cast<char>(tolower(*c));
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus_id);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

int ndzb_device_numa_node(int device) { return device_numa_node(device); }

int ndzb_bind_host_to_device(int device) {
    const int node = device_numa_node(device);
    if (node < 0) return -1;
    // CPUs of the node: "0-15,64-79"
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char list[4096] = {0};
    const bool got = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!got) return -1;
    cpu_set_t set;
    CPU_ZERO(&set);
    int cpus = 0;
    char *save = nullptr;
    for (char *tok = strtok_r(list, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
        int a = 0, b = 0;
        const int n = sscanf(tok, "%d-%d", &a, &b);
        if (n < 1) continue;
        if (n == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
            CPU_SET(c, &set);
            ++cpus;
        }
    }
    if (cpus == 0) return -1;
    // The calling thread (and the threads it starts) run on the node's cores; new pages — the pinned buffers allocated
    // from here on — come from the node's memory. MPOL_PREFERRED rather than MPOL_BIND: a full node falls back instead
    // of failing. Containers may refuse either call: report the node only if the affinity took.
    if (sched_setaffinity(0, sizeof set, &set) != 0) return -1;
    unsigned long mask[16] = {0};
    if (node < static_cast<int>(sizeof mask * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);  // best effort
    }
    return node;
}

void ndzb_ctx_destroy(ndzb_ctx *ctx) {
    if (!ctx) return;
    const device_guard on_device(ctx->device);
    if (ctx->d_desc) cudaFree(ctx->d_desc);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    for (auto b : ctx->d_blocks) {
        if (b) cudaFree(b);
    }
    if (ctx->d_watch) cudaFree(ctx->d_watch);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_length) cudaFree(ctx->d_length);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    for (auto *v : {&ctx->ev_in, &ctx->ev_done, &ctx->ev_k0, &ctx->ev_k1}) {
        for (cudaEvent_t e : *v) cudaEventDestroy(e);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    delete ctx;
}

int ndzb_compress(ndzb_ctx *ctx, const void *d_data, int dims, const uint32_t *size, void *d_stream,
        uint32_t *d_length_words) {
    if (int rc = check_call(ctx, dims, size)) return rc;
    const device_guard on_device(ctx->device);
    ctx->last_launches = 0;
    const grid_geom g = make_geom(dims, size);
    const border_geom bg = make_border(dims, size);
    const uint32_t H = g.num_cubes;
    const uint32_t hdr = header_words(ctx->dtype, H);
    if (ndzb_compressed_length_bound(ctx->dtype, dims, size) >= (1ull << 32)) return NDZB_ERR_INVALID_ARGUMENT;
    if ((H || bg.count) && (!d_data || !d_stream)) return NDZB_ERR_INVALID_ARGUMENT;

    if (H > 0) {
        uint32_t *offsets = static_cast<uint32_t *>(d_stream);
        uint32_t *pad = (ctx->dtype == NDZB_F64 && (H & 1u)) ? offsets + H : nullptr;
        void *cubes = static_cast<char *>(d_stream) + static_cast<size_t>(hdr) * word_bytes(ctx->dtype);
        if (int rc = enqueue_compress_range(ctx, d_data, g, 0, H, cubes, offsets, pad, d_length_words,
                    hdr + static_cast<uint32_t>(bg.count))) {
            return rc;
        }
    }
    if (bg.count > 0) {
        const cudaError_t e = launch_pack_border(ctx->dtype, d_data, bg, d_stream, hdr, H ? ctx->d_counters + 1 + ctx->total_cur : nullptr, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(e, "pack_border launch");
        ctx->last_launches += 1;
    }
    if (H == 0 && d_length_words) {
        const cudaError_t e = launch_store_length(d_length_words, hdr + static_cast<uint32_t>(bg.count), nullptr, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(e, "store_length launch");
        ctx->last_launches += 1;
    }
    return NDZB_OK;
}

int ndzb_decompress(ndzb_ctx *ctx, const void *d_stream, void *d_data, int dims, const uint32_t *size) {
    if (int rc = check_call(ctx, dims, size)) return rc;
    const device_guard on_device(ctx->device);
    ctx->last_launches = 0;
    const grid_geom g = make_geom(dims, size);
    const border_geom bg = make_border(dims, size);
    const uint32_t H = g.num_cubes;
    const uint32_t hdr = header_words(ctx->dtype, H);
    if ((H || bg.count) && (!d_data || !d_stream)) return NDZB_ERR_INVALID_ARGUMENT;
    const uint32_t *offsets = static_cast<const uint32_t *>(d_stream);
    if (H > 0) {
        const void *cubes = static_cast<const char *>(d_stream) + static_cast<size_t>(hdr) * word_bytes(ctx->dtype);
        if (int rc = enqueue_decompress_range(ctx, cubes, offsets, d_data, g, 0, H)) return rc;
    }
    if (bg.count > 0) {
        const cudaError_t e = launch_unpack_border(ctx->dtype, d_stream, offsets, H, hdr, bg, d_data, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(e, "unpack_border launch");
        ctx->last_launches += 1;
    }
    return NDZB_OK;
}

int ndzb_compress_cubes(ndzb_ctx *ctx, const void *d_data, int dims, const uint32_t *size, uint32_t hc_begin,
        uint32_t hc_end, void *d_cubes, uint32_t *d_offsets_after, uint32_t *d_local_words) {
    if (int rc = check_call(ctx, dims, size)) return rc;
    const device_guard on_device(ctx->device);
    ctx->last_launches = 0;
    const grid_geom g = make_geom(dims, size);
    if (hc_begin > hc_end || hc_end > g.num_cubes || !d_local_words) return NDZB_ERR_INVALID_ARGUMENT;
    if (hc_begin == hc_end) {
        const cudaError_t e = launch_store_length(d_local_words, 0, nullptr, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(e, "store_length launch");
        ctx->last_launches += 1;
        return NDZB_OK;
    }
    if (!d_data || !d_cubes || !d_offsets_after) return NDZB_ERR_INVALID_ARGUMENT;
    return enqueue_compress_range(ctx, d_data, g, hc_begin, hc_end - hc_begin, d_cubes, d_offsets_after, nullptr, d_local_words, 0);
}

int ndzb_add_offset(ndzb_ctx *ctx, uint32_t *d_offsets, uint32_t count, const uint32_t *d_base_words) {
    if (!ctx || (count && (!d_offsets || !d_base_words))) return NDZB_ERR_INVALID_ARGUMENT;
    const device_guard on_device(ctx->device);
    const cudaError_t e = launch_add_offset(d_offsets, count, d_base_words, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "add_offset launch");
    return NDZB_OK;
}

int ndzb_fixup_header(ndzb_ctx *ctx, const uint32_t *d_local_header, uint32_t *d_global_header, uint32_t count,
        const uint32_t *d_gathered_lengths, const uint32_t *d_overhead_words, uint32_t rank) {
    if (!ctx || (count && (!d_local_header || !d_global_header)) || (rank && (!d_gathered_lengths || !d_overhead_words))) {
        return NDZB_ERR_INVALID_ARGUMENT;
    }
    const device_guard on_device(ctx->device);
    const cudaError_t e = launch_fixup_header(d_local_header, d_global_header, count, d_gathered_lengths, d_overhead_words, rank, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "fixup_header launch");
    return NDZB_OK;
}

int ndzb_fixup_header_on(void *cuda_stream, const uint32_t *d_local_header, uint32_t *d_global_header, uint32_t count,
        const uint32_t *d_gathered_lengths, const uint32_t *d_overhead_words, uint32_t rank) {
    if ((count && (!d_local_header || !d_global_header)) || (rank && (!d_gathered_lengths || !d_overhead_words))) {
        return NDZB_ERR_INVALID_ARGUMENT;
    }
    const cudaError_t e = launch_fixup_header(d_local_header, d_global_header, count, d_gathered_lengths, d_overhead_words, rank,
            static_cast<cudaStream_t>(cuda_stream));
    if (e != cudaSuccess) return cuda_fail(e, "fixup_header launch");
    return NDZB_OK;
}

int ndzb_pack_border(ndzb_ctx *ctx, const void *d_data, int dims, const uint32_t *size, void *d_out) {
    if (int rc = check_call(ctx, dims, size)) return rc;
    const device_guard on_device(ctx->device);
    const border_geom bg = make_border(dims, size);
    if (bg.count == 0) return NDZB_OK;
    if (!d_data || !d_out) return NDZB_ERR_INVALID_ARGUMENT;
    const cudaError_t e = launch_pack_border(ctx->dtype, d_data, bg, d_out, 0, nullptr, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "pack_border launch");
    return NDZB_OK;
}

int ndzb_decompress_cubes(ndzb_ctx *ctx, const void *d_stream, void *d_data, int dims, const uint32_t *size,
        uint32_t hc_begin, uint32_t hc_end) {
    if (int rc = check_call(ctx, dims, size)) return rc;
    const device_guard on_device(ctx->device);
    ctx->last_launches = 0;
    const grid_geom g = make_geom(dims, size);
    if (hc_begin > hc_end || hc_end > g.num_cubes) return NDZB_ERR_INVALID_ARGUMENT;
    if (hc_begin == hc_end) return NDZB_OK;
    if (!d_data || !d_stream) return NDZB_ERR_INVALID_ARGUMENT;
    const uint32_t hdr = header_words(ctx->dtype, g.num_cubes);
    const void *cubes = static_cast<const char *>(d_stream) + static_cast<size_t>(hdr) * word_bytes(ctx->dtype);
    return enqueue_decompress_range(ctx, cubes, static_cast<const uint32_t *>(d_stream), d_data, g, hc_begin, hc_end - hc_begin);
}

int ndzb_offload_compress(ndzb_ctx *ctx, const void *h_data, int dims, const uint32_t *size, void *h_stream,
        uint32_t *length_words, uint64_t *kernel_ns) {
    if (int rc = check_call(ctx, dims, size)) return rc;
    if (!length_words) return NDZB_ERR_INVALID_ARGUMENT;
    const size_t wb = word_bytes(ctx->dtype);
    const size_t in_bytes = num_elements(dims, size) * wb;                                        // 64-bit, cf. cuda_codec.inl:681
    const size_t bound_bytes = ndzb_compressed_length_bound(ctx->dtype, dims, size) * wb;
    if (in_bytes && (!h_data || !h_stream)) return NDZB_ERR_INVALID_ARGUMENT;
    const device_guard on_device(ctx->device);
    if (verbose()) printf("Have %u hypercubes\n", ndzb_num_hypercubes(dims, size));  // cuda_codec.inl:676-678
    if (int rc = ensure_buffer(&ctx->d_in, &ctx->d_in_bytes, in_bytes ? in_bytes : 16)) return rc;
    if (int rc = ensure_buffer(&ctx->d_out, &ctx->d_out_bytes, bound_bytes ? bound_bytes : 16)) return rc;
    if (!ctx->d_length) NDZB_CUDA(cudaMalloc(&ctx->d_length, sizeof(uint32_t)));
    {
        const grid_geom g = make_geom(dims, size);
        if (g.num_cubes > 0 && make_border(dims, size).count == 0 && in_bytes >= pipeline_min_bytes() && !getenv("NDZB_NO_PIPELINE")) {
            const chunk_plan plan = plan_chunks(dims, size, g, wb, false);
            if (plan.chunks > 1) {
                ctx->last_launches = 0;
                uint64_t ns = 0;
                const int rc = pipelined_compress(ctx, h_data, dims, size, h_stream, length_words, (kernel_ns || verbose()) ? &ns : nullptr, g, plan);
                ctx->last_launches = static_cast<uint32_t>(plan.chunks);
                if (kernel_ns) *kernel_ns = ns;
                if (rc == NDZB_OK) report_kernel_time(ns);
                return rc;
            }
        }
    }
    if (in_bytes) NDZB_CUDA(cudaMemcpyAsync(ctx->d_in, h_data, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    NDZB_CUDA(cudaEventRecord(ctx->ev_begin, ctx->stream));
    if (int rc = ndzb_compress(ctx, ctx->d_in, dims, size, ctx->d_out, ctx->d_length)) return rc;
    NDZB_CUDA(cudaEventRecord(ctx->ev_end, ctx->stream));
    uint32_t len = 0;
    NDZB_CUDA(cudaMemcpyAsync(&len, ctx->d_length, sizeof len, cudaMemcpyDeviceToHost, ctx->stream));
    NDZB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (len) NDZB_CUDA(cudaMemcpyAsync(h_stream, ctx->d_out, static_cast<size_t>(len) * wb, cudaMemcpyDeviceToHost, ctx->stream));
    NDZB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (kernel_ns || verbose()) {
        float ms = 0;
        NDZB_CUDA(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
        const uint64_t ns = static_cast<uint64_t>(static_cast<double>(ms) * 1e6);
        if (kernel_ns) *kernel_ns = ns;
        report_kernel_time(ns);
    }
    *length_words = len;
    return NDZB_OK;
}

int ndzb_offload_decompress(ndzb_ctx *ctx, const void *h_stream, uint32_t length_words, void *h_data, int dims,
        const uint32_t *size, uint32_t *consumed_words, uint64_t *kernel_ns) {
    if (int rc = check_call(ctx, dims, size)) return rc;
    const size_t wb = word_bytes(ctx->dtype);
    const size_t out_bytes = num_elements(dims, size) * wb;
    if (out_bytes && (!h_data || !h_stream)) return NDZB_ERR_INVALID_ARGUMENT;
    const device_guard on_device(ctx->device);
    // The header is in host memory: check it before anything is enqueued. A truncated or corrupt stream (the tool hands
    // over whatever is left of a file) must not turn into out-of-bounds host reads, copies past the staging buffer or
    // kernels reading beyond it.
    uint64_t stream_words = 0;
    {
        const grid_geom g = make_geom(dims, size);
        const uint32_t H = g.num_cubes, hdr = header_words(ctx->dtype, H);
        stream_words = static_cast<uint64_t>(hdr) + make_border(dims, size).count;
        if (H) {
            if (length_words < hdr) return NDZB_ERR_CORRUPT_STREAM;
            const uint32_t *offsets = static_cast<const uint32_t *>(h_stream);
            const uint32_t bound = ndzb_compressed_cube_bound(ctx->dtype);
            uint32_t previous = 0;
            for (uint32_t i = 0; i < H; ++i) {
                if (offsets[i] < previous || offsets[i] - previous > bound) return NDZB_ERR_CORRUPT_STREAM;
                previous = offsets[i];
            }
            stream_words += previous;
        }
        if (stream_words > length_words) return NDZB_ERR_CORRUPT_STREAM;
    }
    const size_t in_bytes = static_cast<size_t>(stream_words) * wb;  // what the stream occupies, not all the caller offers
    if (int rc = ensure_buffer(&ctx->d_out, &ctx->d_out_bytes, in_bytes ? in_bytes : 16)) return rc;
    if (int rc = ensure_buffer(&ctx->d_in, &ctx->d_in_bytes, out_bytes ? out_bytes : 16)) return rc;
    {
        const grid_geom g = make_geom(dims, size);
        if (g.num_cubes > 0 && make_border(dims, size).count == 0 && out_bytes >= pipeline_min_bytes() && !getenv("NDZB_NO_PIPELINE")) {
            const chunk_plan plan = plan_chunks(dims, size, g, wb, true);
            if (plan.chunks > 1) {
                uint64_t ns = 0;
                const int rc = pipelined_decompress(ctx, h_stream, h_data, dims, size, (kernel_ns || verbose()) ? &ns : nullptr, g, plan);
                ctx->last_launches = static_cast<uint32_t>(plan.chunks);
                if (kernel_ns) *kernel_ns = ns;
                if (rc == NDZB_OK) report_kernel_time(ns);
                if (rc == NDZB_OK && consumed_words) {
                    const uint32_t H = g.num_cubes;
                    *consumed_words = header_words(ctx->dtype, H) + static_cast<const uint32_t *>(h_stream)[H - 1];
                }
                return rc;
            }
        }
    }
    if (in_bytes) NDZB_CUDA(cudaMemcpyAsync(ctx->d_out, h_stream, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    NDZB_CUDA(cudaEventRecord(ctx->ev_begin, ctx->stream));
    if (int rc = ndzb_decompress(ctx, ctx->d_out, ctx->d_in, dims, size)) return rc;
    NDZB_CUDA(cudaEventRecord(ctx->ev_end, ctx->stream));
    if (out_bytes) NDZB_CUDA(cudaMemcpyAsync(h_data, ctx->d_in, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NDZB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (kernel_ns || verbose()) {
        float ms = 0;
        NDZB_CUDA(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
        const uint64_t ns = static_cast<uint64_t>(static_cast<double>(ms) * 1e6);
        if (kernel_ns) *kernel_ns = ns;
        report_kernel_time(ns);
    }
    if (consumed_words) {
        // header + last offset + border (reference cuda_codec.inl:740-745); the header lives in host memory
        const grid_geom g = make_geom(dims, size);
        const uint32_t H = g.num_cubes;
        const uint32_t last = H ? static_cast<const uint32_t *>(h_stream)[H - 1] : 0u;
        *consumed_words = header_words(ctx->dtype, H) + last + static_cast<uint32_t>(make_border(dims, size).count);
    }
    return NDZB_OK;
}

int ndzb_selftest_lookback(ndzb_ctx *ctx, int mode, const uint32_t *d_lengths, uint32_t count, uint32_t base_words, uint32_t *d_exclusive) {
    if (!ctx || mode < 0 || mode > 2 || (count && (!d_lengths || !d_exclusive))) return NDZB_ERR_INVALID_ARGUMENT;
    if (count == 0) return NDZB_OK;
    const device_guard on_device(ctx->device);
    if (int rc = ensure_descriptors(ctx, count)) return rc;
    const uint32_t resident = static_cast<uint32_t>(g_config[ctx->device].num_sms) * 16u;  // all CTAs resident: waiting on earlier tickets cannot deadlock
    const uint32_t grid = count < resident ? count : resident;
    unsigned long long *blocks = ctx->d_blocks[ctx->blocks_cur];
    if (mode == 0) {
        NDZB_CUDA(cudaMemsetAsync(blocks, 0, (static_cast<size_t>(ctx->desc_capacity) / 32 + 1) * kDescStride * sizeof(unsigned long long), ctx->stream));
    }
    const cudaError_t e = launch_selftest_lookback(mode, d_lengths, count, d_exclusive, ctx->d_desc, blocks, ctx->d_counters, ctx->ticket_base,
            ctx->epoch, base_words, ctx->d_watch, grid, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "selftest_lookback launch");
    if (mode == 0) {  // leave the array as a compress launch expects it: all zero
        NDZB_CUDA(cudaMemsetAsync(blocks, 0, (static_cast<size_t>(ctx->desc_capacity) / 32 + 1) * kDescStride * sizeof(unsigned long long), ctx->stream));
    }
    ctx->ticket_base += count + grid;  // every CTA draws one ticket beyond `count`
    if (++ctx->epoch >= (1u << 30)) {
        NDZB_CUDA(cudaMemsetAsync(ctx->d_desc, 0, static_cast<size_t>(ctx->desc_capacity) * kDescStride * sizeof(uint64_t), ctx->stream));
        ctx->epoch = 1;
    }
    return NDZB_OK;
}

int ndzb_selftest_warp_scan(ndzb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, uint32_t n) {
    if (!ctx || (n && (!d_in || !d_out))) return NDZB_ERR_INVALID_ARGUMENT;
    const device_guard on_device(ctx->device);
    const cudaError_t e = launch_selftest_warp_scan(d_in, d_out, n, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "selftest_warp_scan launch");
    return NDZB_OK;
}

uint32_t ndzb_num_hypercubes(int dims, const uint32_t *size) {
    if (dims < 1 || dims > 3 || !size) return 0;
    return make_geom(dims, size).num_cubes;
}

uint64_t ndzb_border_element_count(int dims, const uint32_t *size) {
    if (dims < 1 || dims > 3 || !size) return 0;
    return make_border(dims, size).count;
}

uint32_t ndzb_header_words(int dtype, uint32_t num_hypercubes) { return header_words(dtype, num_hypercubes); }

uint32_t ndzb_compressed_cube_bound(int dtype) { return dtype == NDZB_F32 ? 4224u : 4160u; }

uint64_t ndzb_compressed_length_bound(int dtype, int dims, const uint32_t *size) {
    if (!valid_profile(dtype, dims) || !size) return 0;
    const uint64_t H = make_geom(dims, size).num_cubes;
    return header_words(dtype, static_cast<uint32_t>(H)) + H * ndzb_compressed_cube_bound(dtype) + make_border(dims, size).count;
}

const char *ndzb_strerror(int status) {
    switch (status) {
        case NDZB_OK: return "ok";
        case NDZB_ERR_INVALID_ARGUMENT: return "invalid argument";
        case NDZB_ERR_DIMS_MISMATCH: return "data dimensionality does not match compressor dimensionality";
        case NDZB_ERR_CAPACITY: return "more hypercubes than the context was created for";
        case NDZB_ERR_CUDA: return "CUDA error";
        case NDZB_ERR_ALLOC: return "out of memory";
        case NDZB_ERR_CORRUPT_STREAM: return "compressed input ends inside a stream or its header is corrupt";
        case NDZB_ERR_IO: return "file input / output failed";
        default: return "unknown ndzb status";
    }
}

const char *ndzb_last_cuda_error(void) { return g_cuda_error; }

const char *ndzb_version(void) { return tuning_build() ? "ndzip_b200 0.2 sm_100a tuning" : "ndzip_b200 0.2 sm_100a"; }

uint32_t ndzb_last_launch_count(const ndzb_ctx *ctx) { return ctx ? ctx->last_launches : 0; }

}  // extern "C"
