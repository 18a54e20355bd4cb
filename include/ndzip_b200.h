/* ndzip_b200 — C ABI of the B200-native ndzip hot path (libndzip_b200.so).
 *
 * This is the drop-in boundary for the reference's CUDA back-end: every entry point replaces one
 * piece of celerity/ndzip @ ff4e6702 (paths relative to the reference tree). The C++ adapter in
 * ndzip_b200/csrc/ndzip_adapter.cu implements ndzip::cuda_compressor<T> / cuda_decompressor<T> /
 * offloader<T> (include/ndzip/cuda.hh:10-41, include/ndzip/offload.hh:8-57) on top of these calls;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ / torch types.
 *  - dtype: NDZB_F32 (float, 32-bit stream words) or NDZB_F64 (double, 64-bit stream words).
 *  - size[]: `dims` extents, slowest dimension first (include/ndzip/ndzip.hh:35-160); every array has
 *    fewer than 2^32 elements (ndzip.hh:19-20).
 *  - lengths are counted in stream words (bits_type<T>, ndzip.hh:186-212), like the reference.
 *  - device entry points are asynchronous on the context's CUDA stream and never synchronise;
 *    host ("offload") entry points are synchronous.
 *  - every function returns NDZB_OK (0) or a negative ndzb_status; nothing throws. There is no CPU
 *    fallback: without a CUDA device every compute call fails with NDZB_ERR_CUDA.
 */
#ifndef NDZIP_B200_H
#define NDZIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDZB_F32 0
#define NDZB_F64 1

typedef enum ndzb_status {
    NDZB_OK = 0,
    NDZB_ERR_INVALID_ARGUMENT = -1, /* bad dtype / dims / null pointer */
    NDZB_ERR_DIMS_MISMATCH = -2,    /* reference: std::runtime_error, src/ndzip/cuda_codec.inl:557-559, 631-633 */
    NDZB_ERR_CAPACITY = -3,         /* more hypercubes than the context was created for (compressor_requirements) */
    NDZB_ERR_CUDA = -4,             /* a CUDA runtime/driver call failed; see ndzb_last_cuda_error() */
    NDZB_ERR_ALLOC = -5,
    NDZB_ERR_CORRUPT_STREAM = -6,   /* host-pointer decompression: the header is not monotonic, a cube exceeds its bound, or
                                       the stream ends before header + cubes + border (nothing was enqueued); also: not a
                                       sharded container / corrupt or truncated segment table */
    NDZB_ERR_IO = -7                /* ndzb_container_* file functions: open / read / write failed (errno is set) */
} ndzb_status;

/* Opaque per-stream context. Owns the scratch the reference's cuda_compressor_impl owns
 * (src/ndzip/cuda_codec.inl:536-552): here only the decoupled look-back descriptors
 * (one 8-byte word per 64 bytes and hypercube) and two counters — no chunk scratch, no scan levels.
 * Not thread-safe per object, like the reference. */
typedef struct ndzb_ctx ndzb_ctx;

/* Replaces cuda_compressor_impl's constructor (src/ndzip/cuda_codec.inl:543-552) and
 * make_cuda_compressor / make_cuda_decompressor (src/ndzip/cuda_factory.cu:4-14).
 * max_hypercubes = compressor_requirements' maximum (src/ndzip/common.cc:8-18); 0 is allowed for a
 * decompress-only context. `cuda_stream` is a cudaStream_t (NULL = default stream). */
int ndzb_ctx_create(ndzb_ctx **out_ctx, int dtype, int dims, uint32_t max_hypercubes, void *cuda_stream);
void ndzb_ctx_destroy(ndzb_ctx *ctx);

/* Replaces cuda_compressor_impl::compress (src/ndzip/cuda_codec.inl:554-603): device pointers,
 * `d_stream` holds ndzb_compressed_length_bound() words, `d_length_words` (nullable) receives the
 * stream length (store_stream_length, cuda_codec.inl:507-511). One kernel launch for the cubes
 * (+ one for the border if the extent has one). */
int ndzb_compress(ndzb_ctx *ctx, const void *d_data, int dims, const uint32_t *size, void *d_stream,
        uint32_t *d_length_words);

/* Replaces cuda_decompressor_impl::decompress (src/ndzip/cuda_codec.inl:628-652). */
int ndzb_decompress(ndzb_ctx *ctx, const void *d_stream, void *d_data, int dims, const uint32_t *size);

/* Replaces cuda_offloader::do_compress / do_decompress (src/ndzip/cuda_codec.inl:669-761): host
 * pointers, synchronous; H2D + kernels + D2H. Device staging buffers are cached in the context and
 * byte sizes are 64-bit (the reference re-allocates per call and overflows at 4 GiB,
 * cuda_codec.inl:679-685). `kernel_ns` (nullable) receives the cudaEvent interval around the kernels
 * only, the reference's kernel_duration (cuda_codec.inl:687-704).
 * ndzb_offload_decompress returns in `consumed_words` the stream words consumed (offload.hh:21-24). It validates the
 * header on the host first (offsets non-decreasing, every cube within ndzb_compressed_cube_bound, header + cubes +
 * border inside `length_words`) and fails with NDZB_ERR_CORRUPT_STREAM before touching the device otherwise; only the
 * words the stream occupies are uploaded. */
int ndzb_offload_compress(ndzb_ctx *ctx, const void *h_data, int dims, const uint32_t *size, void *h_stream,
        uint32_t *length_words, uint64_t *kernel_ns);
int ndzb_offload_decompress(ndzb_ctx *ctx, const void *h_stream, uint32_t length_words, void *h_data, int dims,
        const uint32_t *size, uint32_t *consumed_words, uint64_t *kernel_ns);

/* How the two calls above cut an array into chunks of whole cube rows along dimension 0 (H2D copies, kernels and D2H
 * copies of different chunks overlap): *chunks = number of chunks, 0 if the call is not pipelined (border, or below the
 * size threshold); row_begin (nullable, `capacity` >= *chunks + 1 entries) receives the chunks' first cube rows and, last,
 * the number of cube rows. Pure host arithmetic (tests, tuning). */
int ndzb_offload_chunk_plan(int dtype, int dims, const uint32_t *size, int decompress, uint32_t *row_begin, uint32_t capacity,
        uint32_t *chunks);

/* Page-locked host memory for the buffers handed to the two calls above (the offloader overlaps H2D, kernels
 * and D2H only from pinned memory; pageable buffers work, more slowly). Used by tools/ndzip_compress.cc in place of
 * the reference tool's malloc'ed / mmap'ed chunks (reference src/io/io.cc:23-24, 75-76). */
int ndzb_host_alloc(void **out_ptr, size_t bytes);
void ndzb_host_free(void *ptr);

/* Host placement on multi-socket boxes (new work; the reference has no multi-GPU host side). On an 8-GPU node half of
 * the GPUs hang off the other socket: a rank whose staging buffers sit on the wrong NUMA node moves every host<->device
 * byte across the socket interconnect, which all such ranks share. ndzb_device_numa_node returns the NUMA node of CUDA
 * device `device` (sysfs numa_node of its PCI function; -1 if unknown or not a NUMA box). ndzb_bind_host_to_device
 * restricts the calling thread to that node's cores and makes the node the preferred source of new pages — call it
 * once per rank BEFORE allocating pinned buffers (ndzb_host_alloc, cudaHostAlloc). Returns the node, or -1 if nothing
 * was changed (unknown node, or the container forbids sched_setaffinity). */
int ndzb_device_numa_node(int device);
int ndzb_bind_host_to_device(int device);

/* Multi-GPU sharding (new work; nothing to replace in the reference).
 * A rank compresses the hypercube range [hc_begin, hc_end) of the global array `size` that is fully
 * resident on its device (`d_data` points at the global array's element 0 as seen by this rank, i.e.
 * the shard pointer minus the shard's linear offset; only the shard's elements are touched).
 * Output: `d_cubes` receives the rank's compressed cubes back to back (bound:
 * (hc_end-hc_begin) * ndzb_compressed_cube_bound words), `d_offsets_after[i]` the LOCAL inclusive
 * offsets (words) of cube hc_begin+i, `d_local_words` the rank's total. After the cross-rank
 * exclusive scan of the totals, ndzb_add_offset() turns local into global header entries. */
int ndzb_compress_cubes(ndzb_ctx *ctx, const void *d_data, int dims, const uint32_t *size, uint32_t hc_begin,
        uint32_t hc_end, void *d_cubes, uint32_t *d_offsets_after, uint32_t *d_local_words);
int ndzb_add_offset(ndzb_ctx *ctx, uint32_t *d_offsets, uint32_t count, const uint32_t *d_base_words);
/* The exchange step in one launch: given the all-gathered stream lengths of every rank (d_gathered_lengths,
 * words) and every rank's non-cube words (header + border, d_overhead_words), writes this rank's global
 * header entries: d_global_header[i] = d_local_header[i] + sum over lower ranks of their cube words. */
int ndzb_fixup_header(ndzb_ctx *ctx, const uint32_t *d_local_header, uint32_t *d_global_header, uint32_t count,
        const uint32_t *d_gathered_lengths, const uint32_t *d_overhead_words, uint32_t rank);
/* Copies the border elements of `size` (raw bits, ascending linear index) to d_out; needs the whole
 * array addressable from d_data. Returns nothing on an extent without border. */
int ndzb_pack_border(ndzb_ctx *ctx, const void *d_data, int dims, const uint32_t *size, void *d_out);
/* Decompresses the hypercube range [hc_begin, hc_end) of a complete stream into d_data (global base). */
int ndzb_decompress_cubes(ndzb_ctx *ctx, const void *d_stream, void *d_data, int dims, const uint32_t *size,
        uint32_t hc_begin, uint32_t hc_end);

/* Same kernel as ndzb_fixup_header on an explicit stream, without a context (used by the multi-GPU data plane, whose
 * exchange runs on a side stream). */
int ndzb_fixup_header_on(void *cuda_stream, const uint32_t *d_local_header, uint32_t *d_global_header, uint32_t count,
        const uint32_t *d_gathered_lengths, const uint32_t *d_overhead_words, uint32_t rank);

/* ---- Multi-GPU data plane (new work; BASELINE.json configs[3], configs[4]) ---------------------
 * The grid is cut into slabs of whole cube rows along dimension 0, one per rank (one process per GPU, or one host
 * thread per GPU in a single process). Every rank compresses its slab into a SELF-CONTAINED ndzip stream of the slab
 * (the reference decoder reads it with the slab's extent). The data path has one exchange step — an ncclAllGather of
 * one uint32 per rank (the stream lengths) on a high-priority side stream, then one kernel that rewrites the rank's
 * "offset_after" header entries (reference src/ndzip/common.hh:342-358) for the global stream — and an optional
 * final gather (stores over NVLink peer memory, or ncclSend / ncclRecv, straight into place on the root), after which the root holds the stream the
 * reference produces for the whole grid, bit for bit. NCCL is bound with dlopen("libnccl.so.2") on first use. */
typedef struct ndzb_dist ndzb_dist;
#define NDZB_UNIQUE_ID_BYTES 128

/* Where a rank's pieces live (all counts in stream words of the dtype unless noted). Pure geometry: no GPU needed. */
typedef struct ndzb_dist_layout {
    uint32_t slab_begin, slab_end;   /* [begin, end) along dimension 0 of the global grid */
    uint32_t slab_size[3];           /* extent of the slab (dims entries) */
    uint32_t local_cubes;            /* hypercubes of the slab */
    uint32_t cube_index_base;        /* global index of the slab's first hypercube */
    uint32_t local_header_words;
    uint64_t local_border_words;     /* border elements of the slab = words */
    uint64_t border_base;            /* border words of the lower ranks */
    uint64_t local_bound_words;      /* ndzb_compressed_length_bound of the slab: size of the local stream buffer */
    uint32_t global_cubes;
    uint32_t global_header_words;
    uint64_t global_border_words;
    uint64_t global_bound_words;     /* size of the root's buffer for ndzb_dist_gather */
} ndzb_dist_layout;
int ndzb_dist_plan(int dtype, int dims, const uint32_t *global_size, int world, int rank, ndzb_dist_layout *out);

/* Bootstrap. Multi-process: rank 0 calls ndzb_dist_unique_id and ships the 128 bytes to the other ranks by any means
 * (torch.distributed broadcast, MPI, a file); every rank then calls ndzb_dist_create on its own device (collective:
 * ncclCommInitRank). Single process: ndzb_dist_create_local fills out[0..world) for `devices` (ncclCommInitAll,
 * one non-blocking stream per rank); afterwards each rank must be driven by its own host thread.
 * world == 1 needs no NCCL. The object owns an ndzb_ctx for the slab. */
int ndzb_dist_unique_id(void *id128);
int ndzb_dist_create(ndzb_dist **out, int dtype, int dims, const uint32_t *global_size, const void *id128, int rank, int world,
        void *cuda_stream);
int ndzb_dist_create_local(ndzb_dist **out, int dtype, int dims, const uint32_t *global_size, int world, const int *devices);
void ndzb_dist_destroy(ndzb_dist *d);
int ndzb_dist_layout_of(const ndzb_dist *d, int rank, ndzb_dist_layout *out);
void *ndzb_dist_stream(const ndzb_dist *d); /* the cudaStream_t everything below is enqueued on */

/* Compresses the rank's slab (d_slab = its first element) into d_local_stream (layout.local_bound_words words) on the
 * object's stream and starts the exchange on the side stream. d_local_length (nullable, device) receives the local
 * stream length. Asynchronous; nothing after it on the object's stream waits for the exchange. */
int ndzb_dist_compress(ndzb_dist *d, const void *d_slab, void *d_local_stream, uint32_t *d_local_length);
/* Decompresses the slab from its local stream (no communication). */
int ndzb_dist_decompress(ndzb_dist *d, const void *d_local_stream, void *d_slab);
/* Makes the object's stream wait for the exchange of the last ndzb_dist_compress; afterwards the two arrays below are
 * valid there: the rank's slice of the GLOBAL header (local_cubes uint32 entries) and every rank's stream length. */
int ndzb_dist_wait_exchange(ndzb_dist *d);
const uint32_t *ndzb_dist_global_header(const ndzb_dist *d);
const uint32_t *ndzb_dist_gathered_lengths(const ndzb_dist *d);
/* Final stream gather (collective). d_global_stream (root only, layout.global_bound_words words) receives the whole
 * grid's stream; *global_length_words (nullable, host, every rank) its length. Synchronises the object's stream once
 * (the send / receive sizes must be known on the host), then enqueues the transfers and returns. */
int ndzb_dist_gather(ndzb_dist *d, const void *d_local_stream, void *d_global_stream, int root, uint64_t *global_length_words);
/* How the last gather moved the data: 0 = ncclSend / ncclRecv, 1 = stores over NVLink into a CUDA-IPC mapping of the
 * root's buffer (one process per GPU), 2 = the same with plain pointers (all ranks in one process). The peer-memory
 * paths need no receive side: every rank copies its header slice, cube segment and border segment straight to their final
 * offsets, and one 4-byte all-gather behind the copies tells the root's stream that everything has landed. The root
 * announces the path with a small broadcast; NDZB_GATHER=nccl forces path 0, which is also the fallback when the
 * buffer has no IPC handle (pool allocations). */
int ndzb_dist_last_gather_path(const ndzb_dist *d);
/* NCCL / CUDA error text of the last failing ndzb_dist_* call on this thread. */
const char *ndzb_dist_last_error(void);

/* ---- Sharded stream container (new work; nothing to replace in the reference) ------------------
 * Keeps every rank's SELF-CONTAINED slab stream (what ndzb_dist_compress / ndzb_compress produce for the slab) behind
 * a segment table, so that a multi-GPU pipeline that compresses to storage and decompresses again needs neither the
 * cross-rank offset exchange nor the gather to one root:
 *   u32 magic "NDZS" | u32 version | u32 dtype | u32 dims | u32 size[3] | u32 segments
 *   per segment: u32 slab_begin | u32 slab_end (dimension 0) | u64 stream_words | u64 byte_offset
 *   the segments, each starting at a multiple of 16 bytes.
 * Host code only. All functions validate what they read and return NDZB_ERR_CORRUPT_STREAM on anything that is not a
 * well-formed container (bad magic / version, table beyond the buffer, slabs that do not tile dimension 0, misaligned or
 * overlapping segments). */
typedef struct ndzb_container_segment {
    uint32_t slab_begin, slab_end;   /* [begin, end) along dimension 0 of the global grid */
    uint64_t stream_words;           /* length of the slab's ndzip stream in words of the dtype */
    uint64_t byte_offset;            /* where it starts in the container */
} ndzb_container_segment;
typedef struct ndzb_container_info {
    int32_t dtype, dims;
    uint32_t size[3];                /* extent of the whole grid (dims entries) */
    uint32_t segments;
    uint64_t header_bytes;           /* segment table incl. padding = offset of the first segment */
    uint64_t total_bytes;            /* size of the container */
} ndzb_container_info;

uint64_t ndzb_container_header_bytes(uint32_t segments);
/* Segment table for `segments` ranks that own the slabs of ndzb_dist_plan(…, world = segments, …) and whose slab streams
 * are stream_words[r] words long (e.g. ndzb_dist_gathered_lengths). stream_words and out_segments may both be null to
 * query header_bytes only. */
int ndzb_container_plan(int dtype, int dims, const uint32_t *global_size, uint32_t segments, const uint64_t *stream_words,
        ndzb_container_info *info, ndzb_container_segment *out_segments);
int ndzb_container_encode_header(const ndzb_container_info *info, const ndzb_container_segment *segments, void *out, uint64_t out_bytes);
/* Parses the table at the start of `buf` (any alignment). out_segments may be null (then only *info is filled: call
 * again with info->segments entries); NDZB_ERR_CAPACITY if max_segments is too small. */
int ndzb_container_decode_header(const void *buf, uint64_t bytes, ndzb_container_info *info, ndzb_container_segment *out_segments,
        uint32_t max_segments);
/* The reference's single stream of the whole grid from a container in host memory whose slabs are those of
 * ndzb_dist_plan: header entries rebased by the lower ranks' cube words (reference src/ndzip/common.hh:342-358), cube
 * and border segments concatenated in rank order. out_stream may be null to query *out_words. */
int ndzb_container_to_global_stream(const void *container, uint64_t bytes, void *out_stream, uint64_t capacity_words, uint64_t *out_words);
/* Files. One rank creates the file (table + final size), then every rank writes its own segment with pwrite() — no
 * data passes through another rank; readers fetch the table and the segments they own. */
int ndzb_container_create_file(const char *path, const ndzb_container_info *info, const ndzb_container_segment *segments);
int ndzb_container_write_segment(const char *path, int dtype, const ndzb_container_segment *segment, const void *h_stream);
int ndzb_container_read_header(const char *path, ndzb_container_info *info, ndzb_container_segment *out_segments, uint32_t max_segments);
int ndzb_container_read_segment(const char *path, int dtype, const ndzb_container_segment *segment, void *h_out);
/* Sharded decompression: segment `index` of a container in host memory -> its slab (host pointer, slab extent =
 * [slab_end - slab_begin, size[1], size[2]]), through ndzb_offload_decompress on `ctx` (created for the container's
 * dtype / dims and at least the slab's hypercubes). */
int ndzb_container_decompress_segment(ndzb_ctx *ctx, const void *container, uint64_t bytes, uint32_t index, void *h_slab, uint64_t *kernel_ns);

/* Device-side self tests of the scan primitives, the counterpart of the reference's src/test/cuda_bits_test.cu:37-114
 * (warp scan, hierarchical scan in isolation). ndzb_selftest_lookback runs the decoupled look-back of the compress
 * kernel over `count` items with the given lengths — mode 0: two-level (float profiles), 1 / 2: windows of 32 / 64
 * items (double profiles / tuning) — on the context's descriptors and writes base_words + the exclusive prefix sums;
 * ndzb_selftest_warp_scan writes the per-warp inclusive sums the encoder and decoder warps compute. */
int ndzb_selftest_lookback(ndzb_ctx *ctx, int mode, const uint32_t *d_lengths, uint32_t count, uint32_t base_words, uint32_t *d_exclusive);
int ndzb_selftest_warp_scan(ndzb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, uint32_t n);

/* Host-side stream arithmetic (no GPU needed). */
/* src/ndzip/common.hh:395-412 */
uint32_t ndzb_num_hypercubes(int dims, const uint32_t *size);
/* src/ndzip/common.cc:31-52; 64-bit so that callers can detect overflow of the reference's uint32 */
uint64_t ndzb_compressed_length_bound(int dtype, int dims, const uint32_t *size);
/* src/ndzip/common.hh:308-317 */
uint64_t ndzb_border_element_count(int dims, const uint32_t *size);
/* src/ndzip/common.hh:350-352: words occupied by the offset header */
uint32_t ndzb_header_words(int dtype, uint32_t num_hypercubes);
/* src/ndzip/common.hh:391-392: 4224 (f32) / 4160 (f64) */
uint32_t ndzb_compressed_cube_bound(int dtype);

const char *ndzb_strerror(int status);
/* cudaError_t / CUresult of the last NDZB_ERR_CUDA on this thread, as text. */
const char *ndzb_last_cuda_error(void);
/* Library build identification: "ndzip_b200 <version> sm_100a". */
const char *ndzb_version(void);
/* Number of kernels the last compress / decompress call on this context enqueued (bench.py's gpu_launches). */
uint32_t ndzb_last_launch_count(const ndzb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NDZIP_B200_H */
