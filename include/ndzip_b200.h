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
    NDZB_ERR_ALLOC = -5
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
 * ndzb_offload_decompress returns in `consumed_words` the stream words consumed (offload.hh:21-24). */
int ndzb_offload_compress(ndzb_ctx *ctx, const void *h_data, int dims, const uint32_t *size, void *h_stream,
        uint32_t *length_words, uint64_t *kernel_ns);
int ndzb_offload_decompress(ndzb_ctx *ctx, const void *h_stream, uint32_t length_words, void *h_data, int dims,
        const uint32_t *size, uint32_t *consumed_words, uint64_t *kernel_ns);

/* Page-locked host memory for the buffers handed to the two calls above (the offloader overlaps H2D, kernels
 * and D2H only from pinned memory; pageable buffers work, more slowly). Used by tools/ndzip_compress.cc in place of
 * the reference tool's malloc'ed / mmap'ed chunks (reference src/io/io.cc:23-24, 75-76). */
int ndzb_host_alloc(void **out_ptr, size_t bytes);
void ndzb_host_free(void *ptr);

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
