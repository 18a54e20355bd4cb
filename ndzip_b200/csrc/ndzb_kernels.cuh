// ndzb_kernels.cuh — launch interface between the C ABI (ndzb_capi.cu) and the kernels (ndzb_kernels.cu).
#pragma once

#include "ndzb_cube.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

namespace ndzb {

enum class load_path : int {
    tma = 0,     // cp.async.bulk.tensor into the swizzled tile, double-buffered (needs 16-byte aligned base/strides)
    vec16 = 1,   // 16-byte global loads + manual swizzle (same alignment requirement, no TMA)
    scalar = 2,  // element-wise global loads: any shape / alignment
};

// Look-back descriptors are kDescStride 64-bit words apart. Packed (stride 1), the ~200 descriptors that are live at
// any time sit in a dozen 128-byte lines which every retire warp of every SM polls and every cube writes twice: with
// one descriptor per 32-byte sector the kernel is 13 % faster and more retire warps start to pay, with one per 64
// bytes another 2 % (profiles/README.md). Costs 64 instead of 8 bytes of scratch per hypercube.
#ifndef NDZB_DESC_STRIDE
#define NDZB_DESC_STRIDE 8
#endif
constexpr int kDescStride = NDZB_DESC_STRIDE;

constexpr uint32_t kWatchdogWords = 8 + 6 * 4000;

struct compress_launch {
    const void *data;          // element 0 of the (global) array
    grid_geom geom;
    uint32_t hc_begin;         // first hypercube of the range this launch compresses
    uint32_t count;            // number of hypercubes in the range (> 0)
    void *out_cubes;           // bits_type*: where the range's first compressed cube goes
    uint32_t *out_offsets;     // inclusive word offsets, entry i belongs to cube hc_begin + i
    uint32_t *pad_word;        // nullable: header padding word to be zeroed (f64, odd H)
    const uint32_t *base_words; // nullable: device scalar added to every offset of this launch (chained launches)
    uint32_t *total_words;     // device scalar: base + compressed words of the whole range
    uint32_t *total_host;      // nullable: mapped pinned host word that receives the same total (pipelined offloader: the
                               // host sizes the chunk's D2H copy from it without a copy node in the compute stream)
    uint32_t *length_out;      // nullable: receives length_add + total
    uint32_t length_add;
    uint64_t *desc;            // decoupled look-back descriptors, >= count entries
    unsigned long long *block_desc;  // compress_ws_kernel, two-level look-back: one word per 32 cubes (count << 40 | sum of lengths), all zero when the launch starts
    unsigned long long *block_desc_next;  // the array the NEXT launch will use: this launch zeroes its first block_words_next words (nullable)
    uint32_t block_words_next;
    uint32_t *ticket;          // free-running ticket counter
    uint32_t ticket_base;      // value of *ticket when this launch starts
    uint32_t epoch;            // tag that invalidates descriptors of earlier launches (< 2^30)
    uint32_t debug_flags;      // compress_ws_kernel profiling aids (NDZB_WS_DEBUG): 1 = skip the copy-out, 2 = skip the look-back
    unsigned long long *stats; // compress_ws_kernel, Stats instantiations: 16 counters summed over the grid (nullable)
    uint32_t *watch;           // compress_ws_kernel: kWatchdogWords words, spin-loop watchdog ([0] raised flag, [1] records, [7] 1 = no trap, [8..] records)
};

struct decompress_launch {
    const void *stream_cubes;      // bits_type*: first compressed cube of the stream (after the header)
    const uint32_t *offsets;       // the stream header: inclusive offsets of ALL cubes
    void *data;                    // element 0 of the (global) output array
    grid_geom geom;
    uint32_t hc_begin, count;
};

struct kernel_config {
    int num_sms = 0;
    int ctas_per_sm[2][3][3] = {};  // [dtype][dims-1][load_path] occupancy of the compress kernels
    int dec_ctas_per_sm[2][3][3] = {};  // [dtype][dims-1][store path: scalar / vec16 / tma]
};

// Queries occupancy and opts the kernels into their dynamic shared memory size. Returns cudaError_t.
cudaError_t configure_kernels(kernel_config &cfg);

// Number of tickets a compress launch of `grid` CTAs draws beyond `count` (see ndzb_kernels.cu).
uint32_t compress_ticket_overdraw(uint32_t grid);

// All launchers are asynchronous on `stream` and return the launch error, if any.
cudaError_t launch_compress(int dtype, int dims, load_path path, const compress_launch &args, const CUtensorMap *tmap,
        uint32_t grid, cudaStream_t stream);
// Warp-specialised compress kernel (TMA-compatible inputs only): one CTA per SM, `variant` < compress_ws_variants(dtype).
uint32_t compress_ws_ticket_overdraw(int dtype, int dims, int variant, uint32_t grid, uint32_t count);
int compress_ws_variants(int dtype);
bool tuning_build();  // compiled with -DNDZB_TUNING (variants 1-4, statistics, debug aids, compress_kernel for TMA inputs)
bool compress_ws_uses_blocks(int dtype, int variant);  // two-level look-back: needs block_desc / block_desc_next
cudaError_t launch_compress_ws(int dtype, int dims, int variant, const compress_launch &args, const CUtensorMap &in_map,
        uint32_t grid, cudaStream_t stream);
// store: 0 element-wise, 1 vectorised LSU stores, 2 TMA tensor store of the decoded tile (needs out_map)
cudaError_t launch_decompress(int dtype, int dims, int store, const decompress_launch &args, const CUtensorMap *out_map, uint32_t grid,
        cudaStream_t stream);

// Warp-specialised decoder (one persistent CTA per SM, ring of slots, 7 decode groups + loader): float 2-D / 3-D with a
// TMA-addressable output. grid <= number of SMs.
bool decompress_ws_available(int dtype, int dims);
cudaError_t launch_decompress_ws(int dims, const decompress_launch &args, const CUtensorMap &out_map, uint32_t grid, cudaStream_t stream);

// Border: stream_border[i] = bits(data[border_linear_index(i)]) and the inverse.
// `total_words` (nullable) is a device scalar added to border_base (the compressed words, only known on device).
cudaError_t launch_pack_border(int dtype, const void *data, const border_geom &bg, void *stream_words,
        uint64_t border_base, const uint32_t *total_words, cudaStream_t stream);
cudaError_t launch_unpack_border(int dtype, const void *stream_words, const uint32_t *offsets, uint32_t num_cubes,
        uint64_t header_words, const border_geom &bg, void *data, cudaStream_t stream);

// offsets[i] += *base for i < count; also used to store a constant length.
cudaError_t launch_add_offset(uint32_t *offsets, uint32_t count, const uint32_t *base, cudaStream_t stream);
// global_header[i] = local_header[i] + sum_{r<rank}(gathered_lengths[r] - overhead_words[r])
cudaError_t launch_fixup_header(const uint32_t *local_header, uint32_t *global_header, uint32_t count,
        const uint32_t *gathered_lengths, const uint32_t *overhead_words, uint32_t rank, cudaStream_t stream);
cudaError_t launch_store_length(uint32_t *length_out, uint32_t value, const uint32_t *plus, cudaStream_t stream);

// Self tests of the scan primitives (ndzb_selftest_*): the decoupled look-back over `count` items of the given lengths
// (mode 0 two-level, 1 / 2 windows of 32 / 64), `grid` persistent one-warp CTAs; the per-warp inclusive sum.
cudaError_t launch_selftest_lookback(int mode, const uint32_t *lengths, uint32_t count, uint32_t *exclusive_out, uint64_t *desc,
        unsigned long long *blocks, uint32_t *ticket, uint32_t ticket_base, uint32_t epoch, uint32_t base, uint32_t *watch, uint32_t grid,
        cudaStream_t stream);
cudaError_t launch_selftest_warp_scan(const uint32_t *in, uint32_t *out, uint32_t n, cudaStream_t stream);

// TMA tensor map for the compress input tile of (dtype, dims); returns false if the shape / pointer
// does not meet TMA's alignment rules (caller then uses vec16 or scalar).
bool tma_compatible(int dtype, int dims, const void *data, const grid_geom &g);
// Fills `map`; returns CUDA_SUCCESS or the driver error.
CUresult make_input_tensor_map(CUtensorMap *map, int dtype, int dims, const void *data, const grid_geom &g);
// Tensor map of the decoder's output tile (same alignment rules: tma_compatible).
CUresult make_output_tensor_map(CUtensorMap *map, int dtype, int dims, const void *data, const grid_geom &g);

}  // namespace ndzb
