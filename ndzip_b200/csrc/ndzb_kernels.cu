// ndzb_kernels.cu — the sm_100a kernels of the ndzip hot path and their launchers.
//
// compress_kernel   replaces compress_block + hierarchical_inclusive_scan + compact_all_chunks +
//                   store_stream_length of the reference (src/ndzip/cuda_codec.inl:401-457, 507-511,
//                   src/ndzip/cuda_bits.cuh:266-333) with ONE persistent kernel:
//                     ticket -> TMA tensor load of the cube into a swizzled smem tile (double
//                     buffered, mbarrier) -> per-thread stencil residuals + in-register 32x32 bit
//                     transpose -> decoupled look-back over per-cube lengths (single pass, no
//                     scratch stream, no second pass) -> warp-per-chunk compaction straight into the
//                     final stream position + header entry.
// decompress_kernel replaces decompress_block (cuda_codec.inl:477-492).
// border kernels    replace compact_border / expand_border (cuda_codec.inl:463-474, 495-504).
#include "ndzb_kernels.cuh"
#include "ndzb_ptx.cuh"

#include <type_traits>

namespace ndzb {
namespace {

constexpr uint32_t kFullMask = 0xffffffffu;
// pause between two polls of look-back descriptors that are not published yet
#ifndef NDZB_POLL_NS
#define NDZB_POLL_NS 20
#endif
// -DNDZB_TUNING builds the A/B scaffolding: tuning variants 1-4 of compress_ws_kernel, its Stats instantiation, the
// NDZB_WS_DEBUG profiling aids (which emit INVALID streams) and compress_kernel for TMA inputs. The default build
// carries one compress_ws_kernel per profile and none of the debug branches.
#if defined(NDZB_TUNING)
constexpr bool kTuning = true;
#else
constexpr bool kTuning = false;
#endif
constexpr int kSlots = 3;   // cube tiles per CTA: being copied out / being encoded / being loaded by TMA
constexpr int kWarps = kCubeThreads / 32;

template<typename Bits>
struct smem_plan {
    using tr = codec_traits<Bits>;
    static constexpr int tile_words = tr::cube_words32 > tr::image_words32 ? tr::cube_words32 : tr::image_words32;
    static constexpr int slot_bytes = (tile_words * 4 + 1023) / 1024 * 1024;  // SWIZZLE_128B tiles need 1024-byte alignment
};

template<typename Bits>
struct compress_aux {
    uint64_t mbar[kSlots];
    uint32_t ticket[kSlots];
    uint32_t warp_total[kWarps];
    uint32_t prefix[2];
};

template<typename Bits>
struct decompress_aux {
    uint64_t in_bar[2];  // compressed cube has landed in buffer 0 / 1
    Bits segment_total[4][64];  // 2D: totals of the four 16-row segments of every column
    uint32_t warp_total[kWarps];
    Bits warp_sum[kWarps];
};
constexpr int kDecodeBuffers = 2;  // compressed cube being decoded / next one streaming in (cp.async)

template<typename Bits>
constexpr size_t compress_smem_bytes() {
    return static_cast<size_t>(kSlots) * smem_plan<Bits>::slot_bytes + sizeof(compress_aux<Bits>);
}
template<typename Bits>
struct decode_plan {
    // compressed image (+ up to 3 words of alignment shift) or value tile, whichever is larger
    static constexpr int buffer_bytes = ((codec_traits<Bits>::image_words32 + 8) * 4 + 1023) / 1024 * 1024;
};
template<typename Bits>
constexpr size_t decompress_smem_bytes() {
    return static_cast<size_t>(kDecodeBuffers) * decode_plan<Bits>::buffer_bytes + sizeof(decompress_aux<Bits>);
}

__device__ __forceinline__ uint32_t popc_bits(uint32_t v) { return __popc(v); }
__device__ __forceinline__ uint32_t popc_bits(uint64_t v) { return __popcll(v); }

__device__ __forceinline__ uint32_t warp_inclusive_sum(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v += up;
    }
    return v;
}
template<typename Bits>
__device__ __forceinline__ Bits warp_inclusive_sum_bits(Bits v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Bits up = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v += up;
    }
    return v;
}

// ---- decoupled look-back descriptors -------------------------------------------------------------
// One 64-bit word per cube: [63:34] epoch, [33:32] status, [31:0] value (words). The epoch makes the
// descriptors of earlier launches read as "invalid", so the array never needs to be cleared.
constexpr uint32_t kStatusAggregate = 1;  // value = this cube's compressed length
constexpr uint32_t kStatusPrefix = 2;     // value = inclusive offset after this cube

__device__ __forceinline__ uint64_t pack_desc(uint32_t epoch, uint32_t status, uint32_t value) {
    return (static_cast<uint64_t>((epoch << 2) | status) << 32) | value;
}
__device__ __forceinline__ uint32_t desc_status(uint64_t d, uint32_t epoch) {
    const uint32_t hi = static_cast<uint32_t>(d >> 32);
    return (hi >> 2) == epoch ? (hi & 3u) : 0u;
}

// Exclusive offset of cube t (> 0) = sum of the lengths of cubes 0..t-1. Executed by one full warp.
// Each step inspects a window of 32*kLookBackDepth predecessors (lane l looks at idx-l, idx-32-l, ...):
// at ~200 cubes/us and ~1 us L2 latency the nearest cube whose inclusive prefix is already published is
// typically ~100 cubes back, so a 32-wide window needs several dependent round trips (measured: 2-3).
constexpr int kLookBackDepth = 2;

template<int Depth>
struct look_back_window {
    uint64_t d[Depth];
};
using look_back_sample = look_back_window<kLookBackDepth>;

// Loads the descriptors of the window ending at predecessor `idx` (lane l: idx-l, idx-32-l, ...).
// Positions before cube 0 read as a published prefix equal to the launch's base offset.
template<int Depth = kLookBackDepth, bool Cg = false>
__device__ __forceinline__ look_back_window<Depth> look_back_load(const uint64_t *desc, int64_t idx, uint32_t epoch, int lane, uint32_t base) {
    look_back_window<Depth> s;
#pragma unroll
    for (int k = 0; k < Depth; ++k) {
        const int64_t mine = idx - 32 * k - lane;
        s.d[k] = mine >= 0 ? (Cg ? ptx::ld_cg(desc + mine * kDescStride) : ptx::ld_relaxed_gpu(desc + mine * kDescStride)) : pack_desc(epoch, kStatusPrefix, base);
    }
    return s;
}

// Spin-loop watchdog of compress_ws_kernel. watch[0] != 0 means "give up": the first waiter that has been
// stuck for kWatchdogCycles raises the flag; every spin loop polls it. Every warp that abandons a wait
// appends a record {code, block, thread, x, y, z} at watch[8 + 6 * i] (i < kWatchRecords), so the host sees
// the whole wait-for graph. watch[7] == 0 (production): trap right away instead.
constexpr long long kWatchdogCycles = 20000000000ll;  // ~10 s at 1.9 GHz: far beyond any legitimate wait, also under a profiler
constexpr uint32_t kWatchRecords = 4000;
constexpr uint32_t kWatchWords = 8 + 6 * kWatchRecords;
static_assert(kWatchWords <= kWatchdogWords, "watchdog buffer too small");
struct watchdog {
    uint32_t *watch;
    long long start = 0;
    uint32_t spins = 0;
    __device__ __forceinline__ explicit watchdog(uint32_t *w) : watch(w) {}
    // false = keep spinning, true = abandon the wait
    __device__ __forceinline__ bool expired(uint32_t code, uint32_t x, uint32_t y, uint32_t z) {
        if (watch == nullptr) return false;
        if (spins++ == 0) start = clock64();
        if ((spins & 31u) != 0) return false;
        const bool raised = *reinterpret_cast<volatile uint32_t *>(watch) != 0;
        if (!raised && clock64() - start < kWatchdogCycles) return false;
        if (!raised) {
            atomicCAS(watch, 0u, code);
            if (watch[7] == 0u) __trap();  // production: fail loudly (sticky CUDA error) instead of hanging or returning garbage
        }
        if ((threadIdx.x & 31u) == 0) {
            // look-back waits (there are hundreds, all alike) share the first 8 record slots, the rest is for the others
            const bool lb = code == 0x10Bu;
            const uint32_t n = atomicAdd(watch + (lb ? 1 : 2), 1u);
            const uint32_t i = lb ? n : 8 + n;
            if (lb ? n < 8u : i < kWatchRecords) {
                uint32_t *r = watch + 8 + 6 * i;
                r[0] = code;
                r[1] = blockIdx.x;
                r[2] = threadIdx.x;
                r[3] = x;
                r[4] = y;
                r[5] = z;
            }
        }
        return true;
    }
};

// `first` is a sample of the first window taken earlier (its L2 latency hidden behind other work).
// Returns the exclusive offset; `*aborted` (optional) is set when the watchdog gave up.
template<int Depth = kLookBackDepth, bool Cg = false>
__device__ __forceinline__ uint32_t look_back(const uint64_t *desc, uint32_t t, uint32_t epoch, int lane, look_back_window<Depth> first,
        uint32_t base, uint32_t *watch = nullptr, bool *aborted = nullptr, uint32_t *polls = nullptr) {
    uint32_t exclusive = 0;
    int64_t idx = static_cast<int64_t>(t) - 1;
    look_back_window<Depth> s = first;
    watchdog dog(watch);
    while (true) {
        uint32_t status[Depth];
        while (true) {
            // only predecessors nearer than the nearest published prefix have to be valid
            uint32_t need_wait = 0;
            bool found = false;
#pragma unroll
            for (int k = 0; k < Depth; ++k) {
                status[k] = desc_status(s.d[k], epoch);
                const uint32_t invalid = __ballot_sync(kFullMask, status[k] == 0);
                const uint32_t prefix = __ballot_sync(kFullMask, status[k] == kStatusPrefix);
                if (!found) {
                    const uint32_t nearer = prefix ? ((1u << (__ffs(prefix) - 1)) - 1u) : 0xffffffffu;
                    need_wait |= invalid & nearer;
                    found = prefix != 0;
                }
            }
            if (need_wait == 0) break;
            if (__any_sync(kFullMask, dog.expired(0x10Bu, t, static_cast<uint32_t>(idx), need_wait))) {
                if (aborted) *aborted = true;
                return 0;
            }
            __nanosleep(NDZB_POLL_NS);
            if (polls) *polls += 1u;
            s = look_back_load<Depth, Cg>(desc, idx, epoch, lane, base);
        }
#pragma unroll
        for (int k = 0; k < Depth; ++k) {
            const uint32_t prefix_lanes = __ballot_sync(kFullMask, status[k] == kStatusPrefix);
            const int nearest = prefix_lanes ? __ffs(prefix_lanes) - 1 : 32;
            exclusive += __reduce_add_sync(kFullMask, lane <= nearest ? static_cast<uint32_t>(s.d[k]) : 0u);
            if (prefix_lanes) return exclusive;
        }
        idx -= 32 * Depth;
        if (polls) *polls += 0x10000u;
        s = look_back_load<Depth, Cg>(desc, idx, epoch, lane, base);
    }
}

// ---- two-level look-back ---------------------------------------------------------------------------
// look_back() above walks back 32 * Depth cubes per L2 round trip; at ~170 cubes/us the nearest published offset is
// 2-4 windows away, and every window is a dependent round trip of ~1000 cycles. Here cubes are grouped into blocks
// of 32 consecutive tickets with one 64-bit word per block,
//   [63:40] number of its cubes whose length is known, [39:0] sum of those lengths,
// accumulated with one fire-and-forget atomic per cube (zeroed by the host before the launch). A look-back reads
// the cubes before it in its own block (one per lane) and, per lane, one earlier block: its sum and the inclusive
// offset published by the block's last cube. One round trip covers 1024 cubes.
constexpr uint32_t kBlockShift = 5, kBlockCubes = 1u << kBlockShift;
__device__ __forceinline__ uint64_t block_contribution(uint32_t words) { return (1ull << 40) | words; }

__device__ __forceinline__ uint32_t look_back_blocks(const uint64_t *desc, const uint64_t *blocks, uint32_t t, uint32_t epoch, int lane,
        uint32_t base, uint32_t *watch, bool *aborted, uint32_t *polls) {
    const uint32_t first_of_block = t & ~(kBlockCubes - 1u);
    const bool own_present = static_cast<uint32_t>(lane) < t - first_of_block;  // lane l: cube t-1-l, if in t's block
    const uint64_t *own_ptr = desc + static_cast<size_t>(t - 1u - static_cast<uint32_t>(lane)) * kDescStride;
    uint64_t own = own_present ? ptx::ld_relaxed_gpu(own_ptr) : 0ull;
    // lane l: block (t / 32) - 1 - l; "block -1" is the launch's base offset, published from the start
    int64_t m = static_cast<int64_t>(t >> kBlockShift) - 1 - lane;
    auto block_word = [&](int64_t b) { return blocks + b * kDescStride; };
    auto block_end = [&](int64_t b) { return desc + ((b << kBlockShift) + (kBlockCubes - 1)) * kDescStride; };
    uint64_t blk = m >= 0 ? ptx::ld_relaxed_gpu(block_word(m)) : 0ull;
    uint64_t pfx = m >= 0 ? ptx::ld_relaxed_gpu(block_end(m)) : pack_desc(epoch, kStatusPrefix, base);
    watchdog dog(watch);

    // ---- own block: the cubes between the block's start (or a published offset inside it) and t
    uint32_t sum;
    while (true) {
        const uint32_t st = own_present ? desc_status(own, epoch) : kStatusAggregate;
        const uint32_t prefix = __ballot_sync(kFullMask, own_present && st == kStatusPrefix);
        const uint32_t invalid = __ballot_sync(kFullMask, st == 0);
        const int nearest = prefix ? __ffs(prefix) - 1 : 32;
        const uint32_t need = invalid & (prefix ? (1u << nearest) - 1u : 0xffffffffu);
        if (need == 0) {
            sum = __reduce_add_sync(kFullMask, own_present && lane <= nearest ? static_cast<uint32_t>(own) : 0u);
            if (prefix) return sum;
            break;
        }
        if (__any_sync(kFullMask, dog.expired(0x10Bu, t, need, 1u))) {
            *aborted = true;
            return 0;
        }
        __nanosleep(NDZB_POLL_NS);
        if (polls) *polls += 1u;
        if ((need >> lane) & 1u) own = ptx::ld_relaxed_gpu(own_ptr);
    }
    // ---- earlier blocks, nearest first: whole-block sums up to the nearest block whose end offset is published
    while (true) {
        const bool has_prefix = desc_status(pfx, epoch) == kStatusPrefix;
        const bool complete = m < 0 || static_cast<uint32_t>(blk >> 40) == kBlockCubes;
        const uint32_t prefix = __ballot_sync(kFullMask, has_prefix);
        const uint32_t incomplete = __ballot_sync(kFullMask, !complete && !has_prefix);
        const int nearest = prefix ? __ffs(prefix) - 1 : 32;
        const uint32_t need = incomplete & (prefix ? (1u << nearest) - 1u : 0xffffffffu);
        if (need == 0) {
            const uint32_t mine = lane < nearest ? static_cast<uint32_t>(blk) : lane == nearest ? static_cast<uint32_t>(pfx) : 0u;
            sum += __reduce_add_sync(kFullMask, mine);
            if (prefix) return sum;
            m -= 32;  // (only with more than 1024 unresolved cubes in front of t)
            if (polls) *polls += 0x10000u;
            blk = m >= 0 ? ptx::ld_relaxed_gpu(block_word(m)) : 0ull;
            pfx = m >= 0 ? ptx::ld_relaxed_gpu(block_end(m)) : pack_desc(epoch, kStatusPrefix, base);
            continue;
        }
        if (__any_sync(kFullMask, dog.expired(0x10Bu, t, need, 2u))) {
            *aborted = true;
            return 0;
        }
        __nanosleep(NDZB_POLL_NS);
        if (polls) *polls += 1u;
        if ((need >> lane) & 1u) {
            blk = ptx::ld_relaxed_gpu(block_word(m));
            pfx = ptx::ld_relaxed_gpu(block_end(m));
        }
    }
}

// mbarrier wait with the watchdog; false = abandoned
__device__ __forceinline__ bool mbar_wait_watched(uint64_t *bar, uint32_t parity, uint32_t *watch, uint32_t code, uint32_t x, uint32_t y) {
    if (ptx::mbar_try_wait(bar, parity)) return true;
    watchdog dog(watch);
    while (!ptx::mbar_try_wait(bar, parity)) {
        if (dog.expired(code, x, y, parity)) return false;
    }
    return true;
}

// ---- cube input ------------------------------------------------------------------------------------

template<typename Bits, int Dims>
__device__ __forceinline__ void issue_tma_load(
        uint32_t *slot, uint64_t *bar, const CUtensorMap *map, const grid_geom &g, uint32_t hc) {
    constexpr uint32_t bytes = codec_traits<Bits>::cube_words32 * 4;
    ptx::mbar_arrive_expect_tx(bar, bytes);
    uint32_t ucz, ucy, ucx;
    cube_coords<Dims>(g, hc, ucz, ucy, ucx);
    const int cz = static_cast<int>(ucz), cy = static_cast<int>(ucy), cx = static_cast<int>(ucx);
    if constexpr (sizeof(Bits) == 4 && Dims == 1) {
        ptx::tma_load_2d(slot, map, bar, 0, cx * 128);
    } else if constexpr (sizeof(Bits) == 4 && Dims == 2) {
        ptx::tma_load_3d(slot, map, bar, 0, cx * 2, cy * 64);
    } else if constexpr (sizeof(Bits) == 4 && Dims == 3) {
        // ONE copy for both y-parity regions: the view's slowest dimension is the parity (make_input_tensor_map), so the
        // box [par][z][y/2][x] lands as region 0 = even rows, region 1 = odd rows. Two copies of 128 rows of 64 bytes each
        // reach 4.9 TB/s, this one 6.0 TB/s (scripts/ubench/tma_shapes.cu, profiles/README.md round 2).
        ptx::tma_load_4d(slot, map, bar, cx * 16, cy * 8, cz * 16, 0);
    } else {
        // two regions: float 3D = even-y / odd-y rows (view [z][y/2][y parity][x], SWIZZLE_64B);
        //              double   = first / second 16-value half of every run
        constexpr int region_words = input_layout<Bits, Dims>::region_words;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t *dst = slot + h * region_words;
            if constexpr (Dims == 1) {
                ptx::tma_load_3d(dst, map, bar, 0, h, cx * 128);
            } else if constexpr (Dims == 2) {
                ptx::tma_load_4d(dst, map, bar, 0, h, cx * 2, cy * 64);
            } else {
                ptx::tma_load_4d(dst, map, bar, cx * 16, h, cy * 8, cz * 16);
            }
        }
    }
}

// Cooperative global -> tile copy for shapes TMA cannot take (and as an A/B path for profiling).
template<typename Bits, int Dims, bool Vec16>
__device__ __forceinline__ void load_cube_ldg(uint32_t *tile, const Bits *data, const grid_geom &g, uint32_t hc, int tid) {
    using L = input_layout<Bits, Dims>;
    const uint64_t origin = cube_origin<Dims>(g, hc);
    if constexpr (Vec16) {
        constexpr int elems_per_unit = 16 / sizeof(Bits);
        constexpr int units = kCubeElems / elems_per_unit;
#pragma unroll 8
        for (int q = tid; q < units; q += kCubeThreads) {
            const int e = q * elems_per_unit;
            const uint4 v = ptx::ldg_stream_v4(data + origin + cube_local_offset<Dims>(g, e));
            *reinterpret_cast<uint4 *>(tile + L::elem(e)) = v;
        }
    } else {
#pragma unroll 8
        for (int e = tid; e < kCubeElems; e += kCubeThreads) {
            const Bits v = data[origin + cube_local_offset<Dims>(g, e)];
            const int w = L::elem(e);
            if constexpr (sizeof(Bits) == 4) {
                tile[w] = v;
            } else {
                tile[w] = static_cast<uint32_t>(v);
                tile[w + 1] = static_cast<uint32_t>(v >> 32);
            }
        }
    }
}

// =====================================================================================================
// compress
// =====================================================================================================

template<typename Bits, int Dims, load_path Path>
__global__ void __launch_bounds__(kCubeThreads)
        compress_kernel(const compress_launch a, const __grid_constant__ CUtensorMap tmap) {
    using tr = codec_traits<Bits>;
    constexpr int slot_words = smem_plan<Bits>::slot_bytes / 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t *slots = reinterpret_cast<uint32_t *>(smem_raw);
    auto &aux = *reinterpret_cast<compress_aux<Bits> *>(smem_raw + kSlots * smem_plan<Bits>::slot_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Bits *data = static_cast<const Bits *>(a.data);
    Bits *out_cubes = static_cast<Bits *>(a.out_cubes);
    // offset of this launch's first cube: 0, or the running total left by the previous chained launch
    const uint32_t launch_base = a.base_words ? *a.base_words : 0u;

    // Work is handed out by a free-running ticket counter: ticket order == cube order, which is what
    // makes spinning on predecessors in the look-back deadlock-free (a predecessor's ticket was drawn
    // earlier, hence by a resident CTA).
    //
    // Software pipeline over three tiles (profiles/README.md has the measurements behind each choice):
    //   iteration i:  encode cube t_i and publish its length   |  TMA is loading cube t_{i+1}
    //                 then look back + copy out cube t_{i-1}
    // * A cube's length is always published BEFORE this CTA may wait in a look-back, and every look-back
    //   gets a full iteration of slack. (Waiting first made the time from drawing a ticket to publishing
    //   its length depend on other cubes' look-backs: a convoy, 57 polls per cube.)
    // * Thread 32 owns tickets and TMA: the ticket for iteration i+1 is drawn at the top of iteration i
    //   and its TMA is issued right after barrier B1; warp 0 owns the look-back (sampled at B1, resolved
    //   after phase 2) and hands the offset over at barrier B2.
    // * Measured alternatives that were NOT better over the five BASELINE configs (profiles/README.md):
    //   every warp resolving the look-back itself (r1f), a shared-memory hand-off instead of B2 (r1h),
    //   a fifth "control" warp for tickets/TMA/look-back (r1i), the same with an early non-blocking
    //   look-back attempt (r1j), a 128-wide look-back window. The other warps wait for the look-back
    //   ~20 % of their time in all of them; what is left is L2 latency of the descriptor polls.
    constexpr int kTicketThread = 32;
    if (tid == kTicketThread) {
        if constexpr (Path == load_path::tma) {
            ptx::tma_prefetch_desc(&tmap);
            for (int s = 0; s < kSlots; ++s) ptx::mbar_init(&aux.mbar[s], 1);
            ptx::fence_mbar_init();
        }
        const uint32_t t = atomicAdd(a.ticket, 1u) - a.ticket_base;
        aux.ticket[0] = t;
        if constexpr (Path == load_path::tma) {
            if (t < a.count) issue_tma_load<Bits, Dims>(slots, &aux.mbar[0], &tmap, a.geom, a.hc_begin + t);
        }
    }
    __syncthreads();

    constexpr uint32_t kNone = 0xffffffffu;
    uint32_t prev_t = kNone, prev_words = 0;
    int prev_slot = 0;

    for (uint32_t iter = 0;; ++iter) {
        const int s = iter % kSlots;
        const uint32_t t = aux.ticket[s];  // cube index within the launch's range (CTA-uniform)
        const bool have = t < a.count;
        uint32_t *tile = slots + s * slot_words;
        uint32_t cube_words = 0;
        look_back_sample sample{};
        bool sampled = false;

        if (have) {
            uint32_t next_ticket = 0;
            if (tid == kTicketThread) next_ticket = atomicAdd(a.ticket, 1u) - a.ticket_base;

            if constexpr (Path == load_path::tma) {
                ptx::mbar_wait(&aux.mbar[s], (iter / kSlots) & 1u);
            } else {
                load_cube_ldg<Bits, Dims, Path == load_path::vec16>(tile, data, a.geom, a.hc_begin + t, tid);
                __syncthreads();
            }

            // ---- phase 1: residuals of run `tid`, chunk head, plane count -----------------------------
            Bits r[32];
            residual_run<Bits, Dims>(tile, tid, r);

            Bits head = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) head |= r[j];
            uint32_t count;
            if constexpr (sizeof(Bits) == 4) {
                count = popc_bits(head);
            } else {
                head |= __shfl_xor_sync(kFullMask, head, 1);  // chunk = two adjacent runs
                count = (tid & 1) == 0 ? popc_bits(head) : 0u;
            }
            const uint32_t inclusive = warp_inclusive_sum(count, lane);
            if (lane == 31) aux.warp_total[warp] = inclusive;
            if (tid == kTicketThread) aux.ticket[(iter + 1) % kSlots] = next_ticket;
            __syncthreads();  // B1: reads of the input tile done; warp totals and the next ticket visible

            uint32_t before = 0;
            cube_words = tr::chunks;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const uint32_t wt = aux.warp_total[w];
                cube_words += wt;
                if (w < warp) before += wt;
            }
            // word offset of this thread's chunk body inside the cube (double: both threads of the pair)
            uint32_t body = tr::chunks + before + inclusive - count;
            if constexpr (sizeof(Bits) == 8) body = __shfl_sync(kFullMask, body, lane & ~1);

            // ---- publish the cube length + sample the previous cube's look-back window (warp 0);
            //      prefetch the next cube (thread 32). All are L2 round trips hidden behind phase 2.
            if (warp == 0) {
                if (lane == 0) {
                    ptx::st_relaxed_gpu(a.desc + static_cast<size_t>(t) * kDescStride, pack_desc(a.epoch, kStatusAggregate, cube_words));
                }
                if (prev_t != kNone && prev_t != 0) {
                    sample = look_back_load(a.desc, static_cast<int64_t>(prev_t) - 1, a.epoch, lane, launch_base);
                    sampled = true;
                }
            }
            if constexpr (Path == load_path::tma) {
                if (tid == kTicketThread && next_ticket < a.count) {
                    // tile (iter+1)%kSlots was copied out during the previous iteration, i.e. before B1
                    const int sn = (iter + 1) % kSlots;
                    ptx::fence_proxy_async_smem();
                    issue_tma_load<Bits, Dims>(slots + sn * slot_words, &aux.mbar[sn], &tmap, a.geom, a.hc_begin + next_ticket);
                }
            }

            // ---- phase 2: bit planes, compacted into the cube image (in place over the input tile) ----
            if constexpr (sizeof(Bits) == 4) {
                uint32_t planes[32];
                planes_of_run(r, planes);
                compact_planes(tile, tid, head, body, planes);
            } else {
                uint32_t planes_hi[32], planes_lo[32];
                planes_of_run(r, planes_hi, planes_lo);
                compact_planes(tile, tid >> 1, (tid & 1) == 0, head, body, planes_hi, planes_lo);
            }
        }

        // ---- the previous cube: look back (it has had a whole iteration to become cheap) -------------
        if (prev_t != kNone && warp == 0) {
            if (!sampled && prev_t != 0) sample = look_back_load(a.desc, static_cast<int64_t>(prev_t) - 1, a.epoch, lane, launch_base);
            // cube 0 starts at the launch's base offset (0, or the running total of a chained launch); its
            // descriptor was published as an aggregate and is upgraded to a prefix here like any other
            const uint32_t exclusive = prev_t == 0 ? launch_base : look_back(a.desc, prev_t, a.epoch, lane, sample, launch_base);
            if (lane == 0) {
                const uint32_t after = exclusive + prev_words;
                ptx::st_relaxed_gpu(a.desc + static_cast<size_t>(prev_t) * kDescStride, pack_desc(a.epoch, kStatusPrefix, after));
                aux.prefix[iter & 1] = exclusive;
                a.out_offsets[prev_t] = after;  // "offset_after", reference src/ndzip/common.hh:342-347
                if (prev_t == 0 && a.pad_word) *a.pad_word = 0;  // cuda_codec.inl:446-452
                if (prev_t == a.count - 1) {
                    *a.total_words = after;
                    if (a.total_host) *a.total_host = after;
                    if (a.length_out) *a.length_out = a.length_add + after;  // cuda_codec.inl:507-511
                }
            }
        }
        __syncthreads();  // B2: this iteration's cube image and the previous cube's stream offset are visible

        // ---- coalesced copy of the previous cube's image to its final stream position ---------------
        if (prev_t != kNone) {
            constexpr int w32 = sizeof(Bits) / 4;
            const uint32_t *src = slots + prev_slot * slot_words;
            uint32_t *dst = reinterpret_cast<uint32_t *>(out_cubes + aux.prefix[iter & 1]);
            const int n = static_cast<int>(prev_words) * w32;
#pragma unroll 4
            for (int w = tid; w < n; w += kCubeThreads) dst[w] = src[w];
        }
        if (!have) break;
        prev_t = t;
        prev_words = cube_words;
        prev_slot = s;
    }
}

// =====================================================================================================
// compress, warp-specialised (the default for TMA-compatible inputs)
// =====================================================================================================
//
// One persistent CTA per SM owns ALL of the SM's shared memory as a ring of S cube slots and splits its
// warps by role, so that the warps doing the integer work never wait for anything but their input tile:
//
//   loader  (1 warp, 1 lane)  draws tickets (ticket order == cube order), waits for a free slot and issues
//                             the TMA tensor load of the cube into it                      -> full[s]
//   encoder (G groups of 4 warps; group g takes the CTA's cubes g, g+G, ...)
//                             full[s] -> residuals, heads, plane counts -> named barrier of the group ->
//                             publish the cube length -> transpose + compaction into the cube image, in
//                             place over the input tile                                     -> done[s]
//   retire  (R warps; warp r takes cubes r, r+R, ...)
//                             counted[s] (the cube's length is published; done[s] in the "late" variants) ->
//                             decoupled look-back -> header entry -> done[s] -> coalesced copy of the image
//                             from the slot to its final stream position                   -> empty[s]
//
// Against compress_kernel (3 slots per 4 warps, everything done by the same 4 warps): the pooled ring
// feeds 5 instead of 4 encoder groups per SM from the same shared memory, and the look-back latency (an L2
// round trip of ~1000 cycles or several, while the kernel runs) and the copy-out land on warps that have
// nothing else to do; the look-back overlaps phase 2 of the cube. Defaults (kWsVariants*): float 5 groups + 4
// retire warps, double 3 + 2.
//
// The copy-out is done by the retire warp's own loads and stores because a cube's destination is only
// 4-byte aligned (the stream format packs cubes word by word): cp.async.bulk needs 16-byte alignment, and
// TMA tensor stores turned out to need it too — cp.async.bulk.tensor.{1d,2d}.global.shared::cta with an
// element coordinate that is not a multiple of 16 bytes (or is negative) raises "illegal instruction" on
// sm_100a, while the same box at an aligned coordinate stores correctly (scripts/ubench/tma_store_probe.cu,
// results in profiles/README.md).
//
// Deadlock freedom: tickets are drawn in the order in which the CTA loads and its groups encode its cubes, a
// drawn ticket only ever waits for a slot whose release needs lengths of EARLIER tickets, and encoding never
// waits for anything but its tile; so the smallest ticket whose length is not published yet can always make
// progress, and a look-back only waits for lengths of earlier tickets.

constexpr uint32_t kNoTicket = 0xffffffffu;

template<typename Bits>
struct ws_plan {
    static constexpr int slot_bytes = smem_plan<Bits>::slot_bytes;
    static constexpr int aux_bytes = 2048;
    static constexpr int max_slots = (232448 - aux_bytes) / slot_bytes;  // 227 KiB of dynamic shared memory per CTA
#if defined(NDZB_WS_SLOTS)  // tuning builds: a shorter ring
    static constexpr int slots = NDZB_WS_SLOTS < max_slots ? NDZB_WS_SLOTS : max_slots;
#else
    static constexpr int slots = max_slots;
#endif
};

template<int S, int G>
struct ws_aux {
    uint64_t full[S], done[S], empty[S], resolved[S], counted[S];
    uint32_t ticket[S];
    uint32_t seq[S];    // which of the CTA's cubes the slot holds (guards against mbarrier phase-parity aliasing)
    uint32_t words[S];
    uint32_t offset[S];  // CP > 0: the cube's exclusive offset, handed from its retire warp to its copy warp
    uint32_t issued[S];  // Stats instantiations: clock (low word) at which the slot's TMA load was issued
    uint32_t freed[S];   //                       clock at which the slot was handed back to the loader
    uint32_t warp_total[2][G][4];
    uint32_t next_seq;         // Dyn: the next of the CTA's cubes that no encoder group has taken yet
    uint32_t group_seq[2][G];  // Dyn: the cube each group took (double-buffered)
};

template<typename Bits>
constexpr size_t ws_smem_bytes() {
    return static_cast<size_t>(ws_plan<Bits>::slots) * ws_plan<Bits>::slot_bytes + ws_plan<Bits>::aux_bytes;
}

// one aligned LDS.128 whatever the compiler thinks of the uses of its four words
__device__ __forceinline__ quad ld_quad_shared(const quad *p) {
    quad q;
    // volatile (ordered against the mbarrier waits / arrives around the copy, which are volatile too) but WITHOUT a memory
    // clobber: the global stores of one iteration may sink below the loads of the next
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(ptx::smem_addr(p)));
    return q;
}

// Cube image (shared memory, 1024-byte aligned) -> its place in the stream, by one warp. The destination is only
// 4-byte aligned; the words up to the first 16-byte boundary and the last < 4 words are stored one by one, the
// body as aligned 16-byte stores whose four words are picked from two adjacent 16-byte units of the image
// (0.75 instructions per word instead of 2 for a word-by-word copy). Both units are loaded whole with a forced
// ld.shared.v4 (8 conflict-free shared-memory cycles per 512 bytes): left to itself the compiler fetches exactly the
// words it needs — 4- and 8-byte loads 16 bytes apart, i.e. 4- and 2-way bank conflicts, 8 to 12 cycles
// (-DNDZB_COPY_NARROW_LOADS; 1-1.5 % slower on 512^3 float, profiles/r2_copy_aligned_ab.txt).
// -DNDZB_COPY_SHFL is the variant with ONE load per lane and the next unit's words by shuffle from the lane above
// (5 + head cycles per 512 bytes): 4-13 % SLOWER on every workload (profiles/r2_copy_shfl_ab*.txt) — the copy is a
// latency chain on a single warp, and load -> shuffle -> store in lock step keeps fewer loads in flight.
// The image unit behind the last output quad may lie beyond n: it is still inside the slot.
__device__ __forceinline__ void copy_image_out(const uint32_t *img, uint32_t *dst, uint32_t n, int lane) {
    const uint32_t head = (4u - ((static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst)) >> 2) & 3u)) & 3u;
    if (n < head + 4u) {
        for (uint32_t w = lane; w < n; w += 32) dst[w] = img[w];
        return;
    }
    if (static_cast<uint32_t>(lane) < head) dst[lane] = img[lane];
    const uint32_t nq = (n - head) >> 2;
    const quad *sq = reinterpret_cast<const quad *>(img);
    uint4 *dq = reinterpret_cast<uint4 *>(dst + head);
#if !defined(NDZB_COPY_SHFL)
#if !defined(NDZB_COPY_NARROW_LOADS)  // two whole 16-byte units per lane, conflict-free (forced: ld.shared.v4)
#define COPY_LOAD(p) ld_quad_shared(p)
#else  // plain C++: the compiler fetches only the words it needs, 4- and 8-byte loads 16 bytes apart (bank conflicts)
#define COPY_LOAD(p) (*(p))
#endif
    switch (head) {
        case 0:
#pragma unroll 4
            for (uint32_t j = lane; j < nq; j += 32) {
                const quad a = sq[j];
                dq[j] = uint4{a.x, a.y, a.z, a.w};
            }
            break;
        case 1:
#pragma unroll 4
            for (uint32_t j = lane; j < nq; j += 32) {
                const quad a = COPY_LOAD(sq + j), b = COPY_LOAD(sq + j + 1);
                dq[j] = uint4{a.y, a.z, a.w, b.x};
            }
            break;
        case 2:
#pragma unroll 4
            for (uint32_t j = lane; j < nq; j += 32) {
                const quad a = COPY_LOAD(sq + j), b = COPY_LOAD(sq + j + 1);
                dq[j] = uint4{a.z, a.w, b.x, b.y};
            }
            break;
        default:
#pragma unroll 4
            for (uint32_t j = lane; j < nq; j += 32) {
                const quad a = COPY_LOAD(sq + j), b = COPY_LOAD(sq + j + 1);
                dq[j] = uint4{a.w, b.x, b.y, b.z};
            }
            break;
    }
#undef COPY_LOAD
#else
    // the loops run the same number of times on every lane (the shuffles need the whole warp)
    switch (head) {
        case 0:
#pragma unroll 4
            for (uint32_t j = lane; j < nq; j += 32) {
                const quad a = sq[j];
                dq[j] = uint4{a.x, a.y, a.z, a.w};
            }
            break;
        case 1:
#pragma unroll 4
            for (uint32_t j0 = 0; j0 < nq; j0 += 32) {
                const uint32_t j = j0 + lane;
                const quad a = ld_quad_shared(sq + j);
                uint32_t bx = __shfl_down_sync(kFullMask, a.x, 1);
                if (lane == 31) bx = img[4 * (j + 1)];
                if (j < nq) dq[j] = uint4{a.y, a.z, a.w, bx};
            }
            break;
        case 2:
#pragma unroll 4
            for (uint32_t j0 = 0; j0 < nq; j0 += 32) {
                const uint32_t j = j0 + lane;
                const quad a = ld_quad_shared(sq + j);
                uint32_t bx = __shfl_down_sync(kFullMask, a.x, 1), by = __shfl_down_sync(kFullMask, a.y, 1);
                if (lane == 31) {
                    const uint2 b = *reinterpret_cast<const uint2 *>(img + 4 * (j + 1));
                    bx = b.x;
                    by = b.y;
                }
                if (j < nq) dq[j] = uint4{a.z, a.w, bx, by};
            }
            break;
        default:
#pragma unroll 4
            for (uint32_t j0 = 0; j0 < nq; j0 += 32) {
                const uint32_t j = j0 + lane;
                const quad a = ld_quad_shared(sq + j);
                uint32_t bx = __shfl_down_sync(kFullMask, a.x, 1), by = __shfl_down_sync(kFullMask, a.y, 1),
                         bz = __shfl_down_sync(kFullMask, a.z, 1);
                if (lane == 31) {
                    const quad b = sq[j + 1];
                    bx = b.x;
                    by = b.y;
                    bz = b.z;
                }
                if (j < nq) dq[j] = uint4{a.w, bx, by, bz};
            }
            break;
    }
#endif
    const uint32_t w = head + (nq << 2) + lane;
    if (w < n) dst[w] = img[w];
}

// 3-D residuals with the y-neighbour row taken from the lane below instead of shared memory. residual_run (ndzb_cube.cuh)
// loads five neighbouring half-runs per thread for the stencil: the run one plane back (two rows), the row above in the
// own plane and the row above one plane back. The last two only serve to form "row 2p-1 after the z difference" — which is
// exactly what the thread of run u-1 holds in the second half of ITS registers once it has taken its own z difference.
// So: z difference in registers (own - back), then sixteen shuffles from lane-1, then the y and x differences. Runs with
// p = 0 (lanes 0, 8, 16, 24: the first row pair of a plane) have no row above; every other lane's neighbour is in its own
// warp. Per thread: 16 instead of 24 LDS.128 (the shuffles cost half the shared-memory cycles of the loads they replace),
// 64 instead of 96 rotates, 16 subtractions fewer. Bit-identical to residual_run (tests: every 3-D parity case).
#if !defined(NDZB_NO_SHFL_STENCIL)
template<typename Bits>
__device__ __forceinline__ void residual_run_3d_warp(const uint32_t *tile, int u, Bits *r) {
    residual3_zdiff<Bits>(tile, u, r);  // ndzb_cube.cuh: the two halves are host-callable, tests/host_sim drives them lane by lane
    Bits above[16];  // row 2p-1 of plane z after the z difference: the second row of run u-1
#pragma unroll
    for (int i = 0; i < 16; ++i) above[i] = __shfl_up_sync(kFullMask, r[16 + i], 1);
    residual3_finish<Bits>(u, above, r);
}
#endif

template<typename Bits, int Dims, int G, int R, int LB, int LA, int CP, bool Early, bool Dyn, bool Stats>
__global__ void __launch_bounds__((4 * G + 1 + R + CP) * 32, 1)
        compress_ws_kernel(const compress_launch a, const __grid_constant__ CUtensorMap in_map) {
    using tr = codec_traits<Bits>;
    constexpr int S = ws_plan<Bits>::slots;
    constexpr int slot_words = ws_plan<Bits>::slot_bytes / 4;
    constexpr uint32_t kPoison = (G > R ? G : R) > CP ? (G > R ? G : R) : CP;  // end markers: one for every group, retire and copy warp
    static_assert(S > G && S > R && S > CP && S >= static_cast<int>(kPoison), "ring too small");
    static_assert(sizeof(ws_aux<S, G>) <= ws_plan<Bits>::aux_bytes, "aux area too small");
    static_assert(CP == 0 || Early, "copy warps take over behind an early look-back");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t *slots = reinterpret_cast<uint32_t *>(smem_raw);
    auto &aux = *reinterpret_cast<ws_aux<S, G> *>(smem_raw + S * ws_plan<Bits>::slot_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&aux.full[s], 1);
            ptx::mbar_init(&aux.done[s], kCubeThreads);
            ptx::mbar_init(&aux.empty[s], 1);
            ptx::mbar_init(&aux.resolved[s], 1);
            ptx::mbar_init(&aux.counted[s], 1);
        }
        aux.next_seq = 0;
        ptx::fence_mbar_init();
    }
    __syncthreads();  // the only CTA-wide barrier; the roles below never meet again
    // Stats instantiations only (tuning runs): cycles spent per role and wait, summed over the grid into a.stats
    long long st_a = 0, st_b = 0, st_c = 0, st_d = 0;
    uint32_t st_n = 0, st_polls = 0;
    auto now = [] { return Stats ? clock64() : 0ll; };

    if (warp < 4 * G) {
        // ---------------------------------------------------------------------------------- encoder
        const int g = warp >> 2, u = tid & (kCubeThreads - 1), wg = warp & 3;
        // Static (cube g, g+G, ... per group) or dynamic assignment (Dyn: the group that becomes free takes the CTA's
        // next cube). With static assignment a tile that lands while its group is still busy waits although other
        // groups idle, and every cube behind it in the stream waits for its length.
        static_assert(!Dyn || (G >= R && CP == 0), "dynamic assignment: one end marker per group");
        int s = g;
        uint32_t parity = 0, flip = 0, seq = g, grabs = 0;
        for (;; seq += G) {
            if constexpr (Dyn) {
                if (wg == 0 && lane == 0) aux.group_seq[grabs & 1u][g] = atomicAdd(&aux.next_seq, 1u);
                ptx::named_barrier(1 + g, kCubeThreads);
                seq = aux.group_seq[grabs & 1u][g];
                ++grabs;
                s = static_cast<int>(seq % static_cast<uint32_t>(S));
                parity = (seq / static_cast<uint32_t>(S)) & 1u;
            }
            // A parity wait cannot tell "phase r completed" from "phase r-1 still running": when the TMA load of
            // the cube one ring round back is slower than this group (seen on 1-D inputs, about once in 10^5
            // cubes), the wait falls through while the slot still belongs to that cube. The slot's sequence
            // tag, written by the loader before it arms the barrier, tells the two apart.
            bool alive = true;
            const long long c0 = now();
            do {
                alive = mbar_wait_watched(&aux.full[s], parity, a.watch, 0xF011u, static_cast<uint32_t>(s), (seq << 8) | static_cast<uint32_t>(warp));
            } while (alive && *reinterpret_cast<volatile uint32_t *>(&aux.seq[s]) != seq);
            const long long c1 = now();
            st_a += c1 - c0;
            if (Stats) st_d += static_cast<uint32_t>(static_cast<uint32_t>(c1) - aux.issued[s]);  // TMA issued -> encoder starts
            if (!alive) {
                // watchdog: release siblings that may already stand at the group's barrier, then leave
                asm volatile("bar.arrive %0, %1;" ::"r"(1 + g), "r"(kCubeThreads) : "memory");
                break;
            }
            const uint32_t t = aux.ticket[s];
            if (t >= a.count) {
                // end marker number kNoTicket - t: hand it on to the retire warps; leave with the last one
                // that is addressed to this group
                if (Early && u == 0) ptx::mbar_arrive(&aux.counted[s]);
                ptx::mbar_arrive(&aux.done[s]);
                if (Dyn || (kNoTicket - t) + G >= kPoison) break;
                s += G;
                if (s >= S) {
                    s -= S;
                    parity ^= 1u;
                }
                continue;
            }
            uint32_t *tile = slots + s * slot_words;
            if (kTuning && (a.debug_flags & 4u)) {
                // profiling aid: no encoding at all, every cube "compresses" to its 128 head words (garbage). What is
                // left is the load pipeline: tickets, TMA, slot hand-over, look-back.
                if (u == 0) {
                    ptx::st_relaxed_gpu(a.desc + static_cast<size_t>(t) * kDescStride, pack_desc(a.epoch, kStatusAggregate, tr::chunks));
                    if (LB == 0) {
                        atomicAdd(a.block_desc + static_cast<size_t>(t >> kBlockShift) * kDescStride, static_cast<unsigned long long>(block_contribution(tr::chunks)));
                    }
                    aux.words[s] = tr::chunks;
                    if (Early) ptx::mbar_arrive(&aux.counted[s]);
                }
                ptx::mbar_arrive(&aux.done[s]);
                ++st_n;
                s += G;
                if (s >= S) {
                    s -= S;
                    parity ^= 1u;
                }
                continue;
            }

            // phase 1: residuals of run u, chunk head, plane count
            Bits r[32];
#if !defined(NDZB_NO_SHFL_STENCIL)
            if constexpr (Dims == 3) {
                residual_run_3d_warp<Bits>(tile, u, r);
            } else {
                residual_run<Bits, Dims>(tile, u, r);
            }
#else
            residual_run<Bits, Dims>(tile, u, r);
#endif
            Bits head = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) head |= r[j];
            uint32_t count;
            if constexpr (sizeof(Bits) == 4) {
                count = popc_bits(head);
            } else {
                head |= __shfl_xor_sync(kFullMask, head, 1);  // chunk = two adjacent runs
                count = (u & 1) == 0 ? popc_bits(head) : 0u;
            }
            const uint32_t inclusive = warp_inclusive_sum(count, lane);
            if (lane == 31) aux.warp_total[flip][g][wg] = inclusive;
            ptx::named_barrier(1 + g, kCubeThreads);  // reads of the input tile done; warp totals visible

            uint32_t before = 0, cube_words = tr::chunks;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t wt = aux.warp_total[flip][g][w];
                cube_words += wt;
                if (w < wg) before += wt;
            }
            flip ^= 1u;
            uint32_t body = tr::chunks + before + inclusive - count;
            if constexpr (sizeof(Bits) == 8) body = __shfl_sync(kFullMask, body, lane & ~1);
            if (u == 0) {
                ptx::st_relaxed_gpu(a.desc + static_cast<size_t>(t) * kDescStride, pack_desc(a.epoch, kStatusAggregate, cube_words));
                if (LB == 0) {
                    atomicAdd(a.block_desc + static_cast<size_t>(t >> kBlockShift) * kDescStride, static_cast<unsigned long long>(block_contribution(cube_words)));
                }
                aux.words[s] = cube_words;
                if (Early) ptx::mbar_arrive(&aux.counted[s]);  // the retire warp resolves the cube's offset while phase 2 runs
            }
            const long long c2 = now();
            st_b += c2 - c1;

            // phase 2: bit planes, compacted into the cube image (in place over the input tile)
            if constexpr (sizeof(Bits) == 4) {
                uint32_t planes[32];
                planes_of_run(r, planes);
                compact_planes(tile, u, head, body, planes);
            } else {
                uint32_t planes_hi[32], planes_lo[32];
                planes_of_run(r, planes_hi, planes_lo);
                compact_planes(tile, u >> 1, (u & 1) == 0, head, body, planes_hi, planes_lo);
            }
            ptx::mbar_arrive(&aux.done[s]);
            st_c += now() - c2;
            ++st_n;

            s += G;
            if (s >= S) {
                s -= S;
                parity ^= 1u;
            }
        }
        if (Stats && a.stats && (tid & 127) == 0) {
            atomicAdd(a.stats + 0, static_cast<unsigned long long>(st_a));   // encoder: waiting for the input tile
            atomicAdd(a.stats + 1, static_cast<unsigned long long>(st_b));   // encoder: phase 1 (to the published length)
            atomicAdd(a.stats + 2, static_cast<unsigned long long>(st_c));   // encoder: phase 2
            atomicAdd(a.stats + 3, static_cast<unsigned long long>(st_n));   // cubes
            atomicAdd(a.stats + 7, static_cast<unsigned long long>(st_d));   // TMA issue -> encoder start (load latency + waiting for the group)
        }
    } else if (warp == 4 * G) {
        // ----------------------------------------------------------------------------------- loader
        constexpr bool Pairs = LA < -1;      // look-ahead -LA, drawn as one run of consecutive tickets (pairs: -2)
        constexpr int kLA = Pairs ? -LA : LA;
        if (lane != 0) {
            // The loader warp's other 31 lanes have one job: zero the block words of the NEXT launch (the two arrays
            // alternate), which saves a memset node in front of every launch.
            if (LB == 0 && a.block_desc_next) {
                for (uint32_t i = blockIdx.x * 31u + static_cast<uint32_t>(lane - 1); i < a.block_words_next; i += gridDim.x * 31u) {
                    a.block_desc_next[static_cast<size_t>(i) * kDescStride] = 0ull;
                }
            }
            return;
        }
        ptx::tma_prefetch_desc(&in_map);
        // Tickets are drawn LA cubes ahead and kept in registers, so the ~1 us round trip of
        // the atomic is never waited for (a ticket drawn early still only ever waits for EARLIER tickets:
        // for the slot of the CTA's cube S places back, whose retirement needs lengths of earlier cubes).
        // (LA == 0: the ticket is drawn only when the slot is free, right before the load is issued, and the loader waits
        // for the atomic: no SM ever holds a ticket it cannot load yet.)
        constexpr int kTk = kLA > 0 ? kLA : 1;
        uint32_t tk[kTk];
        bool ended = false;  // LA == 0: an out-of-range ticket has been drawn, no more draws
        if constexpr (kLA > 0) {
            const uint32_t first = atomicAdd(a.ticket, static_cast<uint32_t>(kLA)) - a.ticket_base;
#pragma unroll
            for (int j = 0; j < kLA; ++j) tk[j] = first + j;
        }
        int s = 0;
        uint32_t parity = 1;  // parity of the phase of empty[s] that ends the PREVIOUS round (none in round 0)
        bool first_round = true;
        uint32_t poison = 0, seq = 0;
        const long long l0 = now();
        while (poison < kPoison) {
#pragma unroll
            for (int j = 0; j < kTk; ++j) {
                if (poison == kPoison) break;
                const long long c0 = now();
                if (!first_round && !mbar_wait_watched(&aux.empty[s], parity, a.watch, 0xE017u, static_cast<uint32_t>(s), seq)) return;
                const long long c1 = now();
                st_a += c1 - c0;
                st_b += now() - c1;
                if constexpr (kLA == 0) {
                    if (!ended) tk[0] = atomicAdd(a.ticket, 1u) - a.ticket_base;
                    ended = tk[0] >= a.count;
                }
                const uint32_t t = tk[j];
                if (t < a.count) {
                    aux.ticket[s] = t;
                    aux.seq[s] = seq;
                    if (Stats) {
                        const uint32_t c = static_cast<uint32_t>(clock64());
                        if (!first_round) st_c += static_cast<uint32_t>(c - aux.freed[s]);  // freed -> TMA issued
                        aux.issued[s] = c;
                    }
#if defined(NDZB_LOADER_FENCE)
                    ptx::fence_proxy_async_smem();  // the slot's previous life (generic reads/writes) before the TMA write
#endif
                    issue_tma_load<Bits, Dims>(slots + s * slot_words, &aux.full[s], &in_map, a.geom, a.hc_begin + t);
                    // The next ticket is drawn AFTER the load is on its way: the proxy fence above is a MEMBAR, which
                    // waits for every memory operation this thread has in flight — with the atomic in front of it the
                    // loader stood still for an L2 round trip per cube (slot free -> load issued: 2300 cycles).
                    if constexpr (Pairs) {
                        // tickets are drawn two at a time: consecutive cubes (x neighbours: the two halves of the same 128-byte
                        // lines of a 3-D float grid) are loaded back to back by the same SM
                        if (j == kLA - 1) {
                            const uint32_t first = atomicAdd(a.ticket, static_cast<uint32_t>(kLA)) - a.ticket_base;
#pragma unroll
                            for (int i = 0; i < kLA; ++i) tk[i] = first + i;
                        }
                    } else if constexpr (kLA > 0) {
                        tk[j] = atomicAdd(a.ticket, 1u) - a.ticket_base;
                    }
                } else {
                    if constexpr (Pairs) {  // out of range: nothing more to load, no more draws
#pragma unroll
                        for (int i = 0; i < kLA; ++i) tk[i] = t;
                    }
                    aux.ticket[s] = kNoTicket - poison;  // end marker number `poison`
                    aux.seq[s] = seq;
                    ptx::mbar_arrive(&aux.full[s]);
                    ++poison;
                }
                ++seq;
                if (++s == S) {
                    s = 0;
                    parity ^= 1u;
                    first_round = false;
                }
            }
        }
        if (Stats && a.stats) {
            atomicAdd(a.stats + 4, static_cast<unsigned long long>(st_a));          // loader: waiting for a free slot
            atomicAdd(a.stats + 5, static_cast<unsigned long long>(st_b));          // loader: waiting for the encoders (prefetch limit)
            atomicAdd(a.stats + 6, static_cast<unsigned long long>(now() - l0));    // loader: total
            atomicAdd(a.stats + 15, static_cast<unsigned long long>(st_c));         // slot freed -> its next TMA load issued
        }
    } else if (warp < 4 * G + 1 + R) {
        // ----------------------------------------------------------------------------------- retire
        const int rw = warp - (4 * G + 1);
        const uint32_t launch_base = a.base_words ? *a.base_words : 0u;
        Bits *out_cubes = static_cast<Bits *>(a.out_cubes);
        int s = rw;
        uint32_t parity = 0, seq = rw;
        for (;; seq += R) {
            bool alive = true;
            const long long c0 = now();
            do {  // same aliasing guard as in the encoder
                alive = mbar_wait_watched(Early ? &aux.counted[s] : &aux.done[s], parity, a.watch, 0xD01Eu, static_cast<uint32_t>(s), (seq << 8) | static_cast<uint32_t>(warp));
            } while (alive && *reinterpret_cast<volatile uint32_t *>(&aux.seq[s]) != seq);
            if (!alive) break;
            const long long c1 = now();
            st_a += c1 - c0;
            const uint32_t t = aux.ticket[s];
            if (t >= a.count) {
                if (CP > 0 && lane == 0) ptx::mbar_arrive(&aux.resolved[s]);  // hand the end marker on to the copy warps
                if ((kNoTicket - t) + R >= kPoison) break;  // the last end marker addressed to this warp
                s += R;
                if (s >= S) {
                    s -= S;
                    parity ^= 1u;
                }
                continue;
            }
            const uint32_t words = aux.words[s];
            uint32_t exclusive = launch_base;
            if (kTuning && (a.debug_flags & 2u)) {
                exclusive = t * static_cast<uint32_t>(tr::max_cube_words);  // profiling aid: no look-back, fixed-stride output (NOT the stream format)
            } else if (t != 0) {
                bool aborted = false;
                if constexpr (LB == 0) {  // two-level look-back over blocks of 32 cubes
                    exclusive = look_back_blocks(reinterpret_cast<const uint64_t *>(a.desc), reinterpret_cast<const uint64_t *>(a.block_desc), t, a.epoch,
                            lane, launch_base, a.watch, &aborted, Stats ? &st_polls : nullptr);
                } else if constexpr (LB < 0) {  // windows of 32 * -LB cubes read with weak L1-bypassing loads (ld.global.cg)
                    const look_back_window<-LB> first = look_back_load<-LB, true>(a.desc, static_cast<int64_t>(t) - 1, a.epoch, lane, launch_base);
                    exclusive = look_back<-LB, true>(a.desc, t, a.epoch, lane, first, launch_base, a.watch, &aborted, Stats ? &st_polls : nullptr);
                } else {
                    constexpr int D = LB > 0 ? LB : 1;
                    const look_back_window<D> first = look_back_load<D>(a.desc, static_cast<int64_t>(t) - 1, a.epoch, lane, launch_base);
                    exclusive = look_back<D>(a.desc, t, a.epoch, lane, first, launch_base, a.watch, &aborted, Stats ? &st_polls : nullptr);
                }
                if (aborted) break;
            }
            const long long c2 = now();
            st_b += c2 - c1;
            if (lane == 0) {
                const uint32_t after = exclusive + words;
                ptx::st_relaxed_gpu(a.desc + static_cast<size_t>(t) * kDescStride, pack_desc(a.epoch, kStatusPrefix, after));
                a.out_offsets[t] = after;  // "offset_after", reference src/ndzip/common.hh:342-347
                if (t == 0 && a.pad_word) *a.pad_word = 0;  // cuda_codec.inl:446-452
                if (t == a.count - 1) {
                    *a.total_words = after;
                    if (a.total_host) *a.total_host = after;
                    if (a.length_out) *a.length_out = a.length_add + after;  // cuda_codec.inl:507-511
                }
            }
            if constexpr (CP > 0) {
                // the offset is known and published: the copy-out is somebody else's job, this warp goes straight to its
                // next look-back (a look-back that queues behind a 1650-cycle copy holds a slot for nothing)
                if (lane == 0) {
                    aux.offset[s] = exclusive;
                    ptx::mbar_arrive(&aux.resolved[s]);
                }
                ++st_n;
                s += R;
                if (s >= S) {
                    s -= S;
                    parity ^= 1u;
                }
                continue;
            }
            long long c3 = c2;
            if (Early) {
                // the offset is known; now the image has to be complete (same phase and tag as `counted`)
                if (!mbar_wait_watched(&aux.done[s], parity, a.watch, 0xD02Eu, static_cast<uint32_t>(s), (seq << 8) | static_cast<uint32_t>(warp))) break;
                c3 = now();
                st_d += c3 - c2;
            }
            // image -> stream
            if (!(kTuning && (a.debug_flags & 1u))) {  // (profiling aid: bit 0 skips the copy)
                constexpr uint32_t w32 = sizeof(Bits) / 4;
                copy_image_out(slots + s * slot_words, reinterpret_cast<uint32_t *>(out_cubes + exclusive), words * w32, lane);
            }
#if !defined(NDZB_NO_RETIRE_FENCE)
            ptx::fence_proxy_async_smem();   // these generic reads before the next TMA load into the slot
#endif
            __syncwarp();
            if (Stats && lane == 0) aux.freed[s] = static_cast<uint32_t>(clock64());
            if (lane == 0) ptx::mbar_arrive(&aux.empty[s]);
            st_c += now() - c3;
            ++st_n;

            s += R;
            if (s >= S) {
                s -= S;
                parity ^= 1u;
            }
        }
        if (Stats && a.stats && lane == 0) {
            atomicAdd(a.stats + 8, static_cast<unsigned long long>(st_a));     // retire: waiting for an encoded cube
            atomicAdd(a.stats + 9, static_cast<unsigned long long>(st_b));     // retire: look-back
            atomicAdd(a.stats + 10, static_cast<unsigned long long>(st_c));    // retire: header entry + copy-out
            atomicAdd(a.stats + 11, static_cast<unsigned long long>(st_n));    // cubes
            atomicAdd(a.stats + 12, static_cast<unsigned long long>(st_polls & 0xffffu));   // look-back: reloads because a predecessor's length was missing
            atomicAdd(a.stats + 13, static_cast<unsigned long long>(st_polls >> 16));       // look-back: windows beyond the first
            atomicAdd(a.stats + 14, static_cast<unsigned long long>(st_d));    // retire (early look-back): waiting for the image after the offset is known
        }
    } else {
        // ------------------------------------------------------------------------------------- copy (CP > 0)
        // Copy warp c takes the CTA's cubes c, c + CP, ...: once the cube's retire warp has resolved and published its
        // offset (resolved[s]) and its encoder group has finished the image (done[s]), the image goes to its final stream
        // position and the slot back to the loader. The retire warps never copy: their next look-back starts at once.
        const int cw = warp - (4 * G + 1 + R);
        Bits *out_cubes = static_cast<Bits *>(a.out_cubes);
        int s = cw;
        uint32_t parity = 0, seq = cw;
        constexpr int kStep = CP > 0 ? CP : 1;
        for (;; seq += kStep) {
            bool alive = true;
            do {  // same aliasing guard as in the encoder
                alive = mbar_wait_watched(&aux.resolved[s], parity, a.watch, 0xC09Eu, static_cast<uint32_t>(s), (seq << 8) | static_cast<uint32_t>(warp));
            } while (alive && *reinterpret_cast<volatile uint32_t *>(&aux.seq[s]) != seq);
            if (!alive) break;
            const uint32_t t = aux.ticket[s];
            if (t >= a.count) {
                if ((kNoTicket - t) + kStep >= kPoison) break;  // the last end marker addressed to this warp
            } else {
                if (!mbar_wait_watched(&aux.done[s], parity, a.watch, 0xC0D0u, static_cast<uint32_t>(s), (seq << 8) | static_cast<uint32_t>(warp))) break;
                constexpr uint32_t w32 = sizeof(Bits) / 4;
                copy_image_out(slots + s * slot_words, reinterpret_cast<uint32_t *>(out_cubes + aux.offset[s]), aux.words[s] * w32, lane);
                ptx::fence_proxy_async_smem();   // these generic reads before the next TMA load into the slot
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&aux.empty[s]);
            }
            s += kStep;
            if (s >= S) {
                s -= S;
                parity ^= 1u;
            }
        }
    }
}

// =====================================================================================================
// decompress
// =====================================================================================================

// ---- column strips: two x-adjacent elements per thread (8 bytes float, 16 bytes double), so that the
// 4096-element cube gives all 128 threads a strip in every pass --------------------------------------
template<typename Bits>
struct strip {
    Bits v[2];

    // p = tile + word address of the strip (strip_addr_* in ndzb_cube.cuh)
    static __device__ __forceinline__ strip load(const uint32_t *p) {
        strip s;
        if constexpr (sizeof(Bits) == 4) {
            const uint2 q = *reinterpret_cast<const uint2 *>(p);
            s.v[0] = q.x;
            s.v[1] = q.y;
        } else {
            ld_pair64(p, s.v[0], s.v[1]);
        }
        return s;
    }
    __device__ __forceinline__ void store(uint32_t *p) const {
        if constexpr (sizeof(Bits) == 4) {
            *reinterpret_cast<uint2 *>(p) = uint2{v[0], v[1]};
        } else {
            st_quad(p, quad{static_cast<uint32_t>(v[0]), static_cast<uint32_t>(v[0] >> 32), static_cast<uint32_t>(v[1]),
                            static_cast<uint32_t>(v[1] >> 32)});
        }
    }
    __device__ __forceinline__ strip operator+(const strip &o) const { return strip{{v[0] + o.v[0], v[1] + o.v[1]}}; }
    // the sign rotation undone (reference src/ndzip/common.hh:441-444): what store() then writes is the final value
    __device__ __forceinline__ strip rotated_back() const { return strip{{rotr1(v[0]), rotr1(v[1])}}; }
    // undo the sign rotation (reference src/ndzip/common.hh:441-444) and write the final values at p
    template<bool Vec>
    __device__ __forceinline__ void emit(Bits *p) const {
        const Bits a = rotr1(v[0]), b = rotr1(v[1]);
        if constexpr (!Vec) {
            p[0] = a;
            p[1] = b;
        } else if constexpr (sizeof(Bits) == 4) {
            asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
        } else {
            ptx::stg_stream_v4(p, uint4{static_cast<uint32_t>(a), static_cast<uint32_t>(a >> 32), static_cast<uint32_t>(b),
                                        static_cast<uint32_t>(b >> 32)});
        }
    }
};

// Streams the compressed cube [begin, end) (stream words) into `buf` with cp.async so that image
// word w lands at buf[shift + w], shift = the cube's misalignment to 16 bytes: the body moves in
// 16-byte copies, the ragged head and tail in 4-byte copies (never touches bytes outside the cube).
template<typename Bits>
__device__ __forceinline__ void stream_in_cube(uint32_t *buf, const Bits *stream_cubes, uint32_t begin, uint32_t end, int tid) {
    using tr = codec_traits<Bits>;
    constexpr int w32 = sizeof(Bits) / 4;
    uint32_t len = end - begin;
    if (len > static_cast<uint32_t>(tr::max_cube_words)) len = tr::max_cube_words;  // corrupt header: stay inside the buffer
    const uint32_t *src = reinterpret_cast<const uint32_t *>(stream_cubes + begin);
    const int n = static_cast<int>(len) * w32;
    const int shift = static_cast<int>((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
    const int head = (4 - shift) & 3;            // words before the first 16-byte boundary
    const int body_groups = n > head ? (n - head) >> 2 : 0;
    const int tail_begin = head + (body_groups << 2);
    for (int g = tid; g < body_groups; g += kCubeThreads) {
        ptx::cp_async_16(buf + shift + head + 4 * g, src + head + 4 * g);
    }
    if (tid < head && tid < n) ptx::cp_async_4(buf + shift + tid, src + tid);
    if (tid >= 32 && tid - 32 < n - tail_begin && tail_begin >= head) ptx::cp_async_4(buf + shift + tail_begin + (tid - 32), src + tail_begin + (tid - 32));
}

// How the decoded hypercube reaches global memory (reference store_hypercube, src/ndzip/cuda_codec.inl:58-65):
//   scalar  element-wise stores: any shape / alignment
//   vec16   8- / 16-byte stores straight from the registers of the last pass (16-byte aligned base and pitch)
//   tma     the last pass writes the finished, rotated-back values into the shared-memory tile in a layout a tensor map
//           describes, and ONE thread issues ONE cp.async.bulk.tensor store per cube (UTMASTG): no per-store 64-bit
//           address arithmetic, 16 shared-memory stores with immediate offsets instead of 16 global stores per thread
enum class store_path : int { scalar = 0, vec16 = 1, tma = 2 };

template<typename Bits, int Dims>
__device__ __forceinline__ void issue_tma_store(const uint32_t *tile, const CUtensorMap *map, const grid_geom &g, uint32_t hc) {
    uint32_t ucz, ucy, ucx;
    cube_coords<Dims>(g, hc, ucz, ucy, ucx);
    const int cz = static_cast<int>(ucz), cy = static_cast<int>(ucy), cx = static_cast<int>(ucx);
    if constexpr (Dims == 1) {
        if constexpr (sizeof(Bits) == 4) ptx::tma_store_2d(map, tile, 0, cx * 128);   // [rows of 32][32]
        else ptx::tma_store_3d(map, tile, 0, cx * 128, 0);                             // [half][runs][16]
    } else if constexpr (Dims == 2) {
        if constexpr (sizeof(Bits) == 4) ptx::tma_store_3d(map, tile, 0, cx * 2, cy * 64);      // [y][x / 32][32]
        else ptx::tma_store_4d(map, tile, 0, cx * 2, cy * 64, 0);                               // [half][y][x / 32][16]
    } else {
        if constexpr (sizeof(Bits) == 4) ptx::tma_store_3d(map, tile, cx * 16, cy * 16, cz * 16);   // [z][y][x], SWIZZLE_64B
        else ptx::tma_store_4d(map, tile, cx * 16, cy * 8, cz * 16, 0);                             // [y parity][z][y / 2][x]
    }
    ptx::tma_store_commit();
}

// Decodes ONE hypercube with 128 threads (tid = 0 .. 127): compressed image (shared memory) -> value tile `tile` (which may
// alias the image) -> global memory (vector / element-wise stores) or, for store_path::tma, the finished tile in the tensor
// map's layout, fenced for the async proxy: the caller closes with a barrier and issues the tensor store. `sync` is the
// barrier of the 128 threads (__syncthreads in decompress_kernel, a named barrier in decompress_ws_kernel);
// `after_first_barrier` runs once behind the first of them (decompress_kernel issues its next copy-in there).
template<typename Bits, int Dims, store_path Out, typename Sync, typename Hook>
__device__ __forceinline__ void decode_cube(uint32_t *tile, const uint32_t *image, uint32_t *warp_total, Bits *warp_sum, Bits (*segment_total)[64],
        const decompress_launch &a, uint32_t hc, int tid, Sync sync, Hook after_first_barrier) {
    constexpr bool Vec16 = Out != store_path::scalar;
    constexpr bool Tma = Out == store_path::tma;
    using tr = codec_traits<Bits>;
    const int lane = tid & 31, warp = tid >> 5;
    Bits *data = static_cast<Bits *>(a.data);
    // ---- chunk heads -> where each chunk's planes start ------------------------------------------------
    Bits head;
    uint32_t count;
    if constexpr (sizeof(Bits) == 4) {
        head = image[tid];
        count = popc_bits(head);
    } else {
        const int c = tid >> 1;
        head = (static_cast<uint64_t>(image[2 * c + 1]) << 32) | image[2 * c];
        count = (tid & 1) == 0 ? popc_bits(head) : 0u;
    }
    const uint32_t inclusive = warp_inclusive_sum(count, lane);
    if (lane == 31) warp_total[warp] = inclusive;
    sync();
    after_first_barrier();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        if (w < warp) before += warp_total[w];
    }
    uint32_t body = tr::chunks + before + inclusive - count;
    if constexpr (sizeof(Bits) == 8) body = __shfl_sync(kFullMask, body, lane & ~1);

    // ---- per-thread: planes -> residuals of run `tid`; x-direction prefix sums ------------------
    Bits r[32];
    if constexpr (sizeof(Bits) == 4) {
        run_of_image(image, head, body, r);
    } else {
        run_of_image(image, (tid & 1) == 0, head, body, r);
    }
    if constexpr (Dims == 3) {
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            r[i] += r[i - 1];
            r[16 + i] += r[16 + i - 1];
        }
    } else {
#pragma unroll
        for (int j = 1; j < 32; ++j) r[j] += r[j - 1];
    }
    if constexpr (Dims == 1) {
        // exclusive block scan of the run totals
        const Bits total = r[31];
        const Bits incl = warp_inclusive_sum_bits<Bits>(total, lane);
        if (lane == 31) warp_sum[warp] = incl;
        sync();
        Bits carry = incl - total;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            if (w < warp) carry += warp_sum[w];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] += carry;
    } else if constexpr (Dims == 2) {
        // a row is two adjacent runs: the right half continues from the left half's total
        const Bits left = __shfl_up_sync(kFullMask, r[31], 1);
        if (tid & 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] += left;
        }
    }
    // the value tile aliases the compressed image: everyone must have read before anyone writes
    sync();
    const uint64_t origin = cube_origin<Dims>(a.geom, hc);
    using S = strip<Bits>;

    if constexpr (Dims == 1 && Tma) {
        // ---- rotate back in registers, run -> tile (the layout the tensor map describes), one tensor store ----
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = rotr1(r[j]);
        store_run(tile, tid, r);
        ptx::fence_proxy_async_smem();
    } else if constexpr (Dims == 1) {
        // ---- rotate back and store, coalesced through the tile (reference cuda_codec.inl:58-65) ----
        store_run(tile, tid, r);
        sync();
        constexpr int UE = 16 / sizeof(Bits);
        constexpr int units = kCubeElems / UE;
        // unit q = tid + 128 i lies 4 UE runs below unit tid: same swizzle, 128 UE words further
        const uint32_t *unit0 = tile + tile_elem<Bits>(tid * UE);
#pragma unroll
        for (int i = 0; i < units / kCubeThreads; ++i) {
            const int e = (tid + i * kCubeThreads) * UE;
            const quad v = ld_quad(unit0 + i * (kCubeThreads * UE));
            if constexpr (sizeof(Bits) == 4) {
                const quad o{rotr1(v.x), rotr1(v.y), rotr1(v.z), rotr1(v.w)};
                if constexpr (Vec16) ptx::stg_stream_v4(data + origin + e, uint4{o.x, o.y, o.z, o.w});
                else { data[origin + e] = o.x; data[origin + e + 1] = o.y; data[origin + e + 2] = o.z; data[origin + e + 3] = o.w; }
            } else {
                const uint64_t a0 = rotr1((static_cast<uint64_t>(v.y) << 32) | v.x), a1 = rotr1((static_cast<uint64_t>(v.w) << 32) | v.z);
                if constexpr (Vec16) ptx::stg_stream_v4(data + origin + e, uint4{static_cast<uint32_t>(a0), static_cast<uint32_t>(a0 >> 32), static_cast<uint32_t>(a1), static_cast<uint32_t>(a1 >> 32)});
                else { data[origin + e] = a0; data[origin + e + 1] = a1; }
            }
        }
    } else if constexpr (Dims == 2) {
        // ---- y direction: 32 two-element column strips x four 16-row segments = 128 threads, scanned
        //      in registers; final values go straight to global memory (no tile write-back)
        store_run(tile, tid, r);
        sync();
        const int xq = tid & 31, seg = tid >> 5;
        const strip_addr_y2<Bits> col(seg, xq);
        S q[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) q[k] = S::load(tile + col.at(k));
#pragma unroll
        for (int k = 1; k < 16; ++k) q[k] = q[k] + q[k - 1];
        segment_total[seg][2 * xq] = q[15].v[0];
        segment_total[seg][2 * xq + 1] = q[15].v[1];
        sync();
        S carry{{0, 0}};
        for (int sg = 0; sg < seg; ++sg) carry = carry + S{{segment_total[sg][2 * xq], segment_total[sg][2 * xq + 1]}};
        if constexpr (Tma) {
            // final values back into the tile, in place (every thread rewrites exactly the strips it read)
#pragma unroll
            for (int k = 0; k < 16; ++k) (q[k] + carry).rotated_back().store(tile + col.at(k));
            ptx::fence_proxy_async_smem();
        } else {
            char *dst = reinterpret_cast<char *>(data + origin + static_cast<uint64_t>(seg * 16) * a.geom.n[2] + xq * 2);
            const uint64_t row_bytes = static_cast<uint64_t>(a.geom.n[2]) * sizeof(Bits);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                (q[k] + carry).template emit<Vec16>(reinterpret_cast<Bits *>(dst));
                dst += row_bytes;
            }
        }
    } else {
        // ---- y direction in the tile, z direction fused with rotate + store; 16 x 8 strips per pass -
        store_run3(tile, tid, r);
        sync();
        const int xq = tid & 7, o = tid >> 3;  // o = z in the y pass, y in the z pass
        {
            const strip_addr_y3<Bits> col(o, xq);
            S q[16];
#pragma unroll
            for (int y = 0; y < 16; ++y) q[y] = S::load(tile + col.at(y));
#pragma unroll
            for (int y = 1; y < 16; ++y) {
                q[y] = q[y] + q[y - 1];
                q[y].store(tile + col.at(y));
            }
        }
        sync();
        {
            const strip_addr_z3<Bits> col(o, xq);
            S q[16];
#pragma unroll
            for (int z = 0; z < 16; ++z) q[z] = S::load(tile + col.at(z));
            if constexpr (Tma) {
#pragma unroll
                for (int z = 1; z < 16; ++z) q[z] = q[z] + q[z - 1];
                if constexpr (sizeof(Bits) == 8) {
                    // double: the value tile already has the layout of the tensor map ([y parity][z][y / 2][x], SWIZZLE_128B):
                    // in place, every thread rewrites the strips it read
#pragma unroll
                    for (int z = 0; z < 16; ++z) q[z].rotated_back().store(tile + col.at(z));
                } else {
                    // float: the passes use a layout with the z parity in the swizzle (tile3_unit), which no tensor map can
                    // describe; the finished values are re-laid out as plain [z][y][x] rows of 64 bytes under SWIZZLE_64B
                    // once everybody has read its column
                    sync();
                    uint32_t *out = tile + o * 16 + (((xq >> 1) ^ ((o >> 1) & 3)) << 2) + ((xq & 1) << 1);
#pragma unroll
                    for (int z = 0; z < 16; ++z) q[z].rotated_back().store(out + z * 256);
                }
                ptx::fence_proxy_async_smem();
            } else {
            // one byte pointer advanced by the plane pitch (a 64-bit add per store) instead of an element index
            // that is multiplied out and scaled for every store (IMAD.WIDE + LEA + LEA.HI.X)
            const uint64_t plane_bytes = static_cast<uint64_t>(a.geom.n[1]) * a.geom.n[2] * sizeof(Bits);
            char *dst = reinterpret_cast<char *>(data + origin + static_cast<uint64_t>(o) * a.geom.n[2] + xq * 2);
            q[0].template emit<Vec16>(reinterpret_cast<Bits *>(dst));
#pragma unroll
            for (int z = 1; z < 16; ++z) {
                q[z] = q[z] + q[z - 1];
                dst += plane_bytes;
                q[z].template emit<Vec16>(reinterpret_cast<Bits *>(dst));
            }
            }
        }
    }
}

template<typename Bits, int Dims, store_path Out>
__global__ void __launch_bounds__(kCubeThreads, sizeof(Bits) == 4 ? 6 : 3) decompress_kernel(const decompress_launch a, const __grid_constant__ CUtensorMap out_map) {
    constexpr bool Vec16 = Out != store_path::scalar;
    constexpr bool Tma = Out == store_path::tma;
    using tr = codec_traits<Bits>;
    constexpr int buf_words = decode_plan<Bits>::buffer_bytes / 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t *bufs = reinterpret_cast<uint32_t *>(smem_raw);
    auto &aux = *reinterpret_cast<decompress_aux<Bits> *>(smem_raw + kDecodeBuffers * decode_plan<Bits>::buffer_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Bits *stream_cubes = static_cast<const Bits *>(a.stream_cubes);
    Bits *data = static_cast<Bits *>(a.data);

    // Cubes are independent: static round-robin over the grid. The stream offsets of a cube are read
    // two iterations ahead and its compressed words are streamed in one iteration ahead.
    // Even lanes hold `begin`, odd lanes `end`: a lane-dependent address keeps the value in an ordinary
    // register until it is shuffled out at its use. (With a CTA-uniform address ptxas moves the result
    // into a uniform register right behind the load and the "prefetch" stalls for the full L2 latency:
    // 13.5 % of all samples in profiles/r1_r1d_decompress.txt.)
    auto offsets_of = [&](uint32_t t) -> uint32_t {
        uint32_t v = 0;
        if (t < a.count) {
            const uint32_t hc = a.hc_begin + t;
            const uint32_t odd = lane & 1;
            if (odd || hc) v = __ldg(a.offsets + hc - 1 + odd);  // reference src/ndzip/common.hh:350-358
        }
        return v;
    };
    // Copy-in: ONE bulk copy per cube (cp.async.bulk, completion on an mbarrier), issued by thread 0, instead
    // of ~5 cp.async + address arithmetic per thread (11 % of the kernel's instructions). The copy covers the
    // 16-byte blocks the cube touches, i.e. up to 12 bytes of its neighbours in the stream; the last cube of the
    // launch's range may have nothing behind it, so it is brought in by the per-thread path (stream_in_cube).
    if (tid == 0) {
        ptx::mbar_init(&aux.in_bar[0], 1);
        ptx::mbar_init(&aux.in_bar[1], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    // The bulk copy also rounds its source DOWN to 16 bytes: harmless inside the stream (it re-reads the tail of the previous
    // cube or of the header), but for the stream's very first cube behind a header shorter than that shift it would
    // read in front of the caller's buffer (the reference API only asks for word alignment): that cube takes the
    // per-thread path too.
    const bool first_cube_unsafe = a.hc_begin == 0
            && (reinterpret_cast<uintptr_t>(stream_cubes) & ~static_cast<uintptr_t>(15)) < reinterpret_cast<uintptr_t>(a.offsets);
    auto bulk_ok = [&](uint32_t t) { return t + 1 < a.count && !(first_cube_unsafe && t == 0); };
    auto copy_in = [&](uint32_t *buf, uint64_t *bar, uint32_t t, uint32_t begin, uint32_t end) {
        if (bulk_ok(t)) {
            if (tid == 0) {
                constexpr uint32_t w32 = sizeof(Bits) / 4;
                if constexpr (Tma) ptx::tma_store_wait_read();  // the tensor store of the cube that was decoded in this buffer
                uint32_t len = end - begin;
                if (len > static_cast<uint32_t>(tr::max_cube_words)) len = tr::max_cube_words;  // corrupt header: stay inside the buffer
                const uint32_t *src = reinterpret_cast<const uint32_t *>(stream_cubes + begin);
                const uint32_t shift = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src) >> 2) & 3u;
                const uint32_t bytes = ((shift + len * w32) * 4u + 15u) & ~15u;
                ptx::fence_proxy_async_smem();  // the buffer's previous life (generic accesses, ordered by the barrier before)
                ptx::mbar_arrive_expect_tx(bar, bytes);
                ptx::bulk_load(buf, src - shift, bytes, bar);
            }
        } else {
            if constexpr (Tma) {  // (rare path: the last cube of the range) every thread writes the buffer: all wait
                if (tid == 0) ptx::tma_store_wait_read();
                __syncthreads();
            }
            stream_in_cube<Bits>(buf, stream_cubes, begin, end, tid);
            ptx::cp_async_commit();
            if (tid == 0) ptx::mbar_arrive(bar);  // keeps the barrier's phase in step; the data is waited for with cp.async.wait_group
        }
    };

    uint32_t cur_offsets = offsets_of(blockIdx.x);
    uint32_t next_offsets = offsets_of(blockIdx.x + gridDim.x);
    if (blockIdx.x < a.count) {
        copy_in(bufs, &aux.in_bar[0], blockIdx.x, __shfl_sync(kFullMask, cur_offsets, 0), __shfl_sync(kFullMask, cur_offsets, 1));
    }

    for (uint32_t k = 0, t = blockIdx.x; t < a.count; ++k, t += gridDim.x) {
        const uint32_t hc = a.hc_begin + t;
        uint32_t *tile = bufs + (k & 1) * buf_words;
        // prefetch: next cube's words into the other buffer (free since the end of the previous iteration),
        // the cube after that's offsets into registers. With tensor stores the other buffer is still being read by the
        // store issued a moment ago: the copy-in is issued a little later, behind the first barrier of the iteration.
        const uint32_t next_begin = __shfl_sync(kFullMask, next_offsets, 0), next_end = __shfl_sync(kFullMask, next_offsets, 1);
        if (!Tma && t + gridDim.x < a.count) {
            copy_in(bufs + ((k + 1) & 1) * buf_words, &aux.in_bar[(k + 1) & 1], t + gridDim.x, next_begin, next_end);
        }
        const uint32_t begin = __shfl_sync(kFullMask, cur_offsets, 0);
        cur_offsets = next_offsets;
        next_offsets = offsets_of(t + 2 * gridDim.x);

        ptx::mbar_wait(&aux.in_bar[k & 1], (k >> 1) & 1u);
        if (!bulk_ok(t)) {
            ptx::cp_async_wait<0>();
            __syncthreads();
        }
        const uint32_t *image = tile + ((reinterpret_cast<uintptr_t>(stream_cubes + begin) >> 2) & 3);

        decode_cube<Bits, Dims, Out>(tile, image, aux.warp_total, aux.warp_sum, aux.segment_total, a, hc, tid, [] { __syncthreads(); }, [&] {
            if (Tma && t + gridDim.x < a.count) {
                copy_in(bufs + ((k + 1) & 1) * buf_words, &aux.in_bar[(k + 1) & 1], t + gridDim.x, next_begin, next_end);
            }
        });
        __syncthreads();  // tile is reused by the cube after next; segment totals by the next cube
        if constexpr (Tma) {
            if (tid == 0) issue_tma_store<Bits, Dims>(tile, &out_map, a.geom, hc);  // everybody's writes + proxy fences are behind the barrier
        }
    }
    if constexpr (Tma) {
        if (tid == 0) ptx::tma_store_wait_all();  // shared memory must outlive the stores that read it
    }
}

// =====================================================================================================
// decompress, warp-specialised (float profiles with a TMA-addressable output)
// =====================================================================================================
//
// decompress_kernel keeps two private buffers per 4-warp CTA, so shared memory (6 x 2 x 17 KiB) caps an SM at 24 decoding
// warps although the tensor-store variants need only 56-64 registers. Here ONE persistent CTA per SM pools the same
// memory as a ring of S slots: a loader warp streams compressed cubes in (one cp.async.bulk each) as far ahead as slots
// are free, G decode groups of 4 warps (group g takes the CTA's cubes g, g + G, ...) decode in place and hand the
// finished tile to a TMA tensor store; a slot returns to the loader once that store has read it. 7 groups = 28 decoding
// warps per SM from the same 13 slots.
//   loader          empty[s] -> offsets -> bulk copy (or, for the two kinds of cube a bulk copy must not touch, a
//                   cooperative copy by the loader warp)                                          -> full[s]
//   decode group    full[s] -> decode_cube (named barrier of the group) -> tensor store; at the start of its next cube
//                   the group's first thread waits until the previous store has read its slot     -> empty[s]
constexpr int kDecodeGroups = 7;  // decompress_ws_kernel: 28 decoding warps + the loader warp = 928 threads, <= 64 registers

template<int S>
struct dws_aux {
    uint64_t full[S], empty[S];
    uint32_t seq[S];      // which of the CTA's cubes the slot holds (mbarrier phase-parity aliasing guard, as in compress_ws_kernel)
    uint32_t shift[S];    // the image starts `shift` words into the slot (the stream position's misalignment to 16 bytes)
    uint32_t warp_total[8][4];
    uint32_t warp_sum[8][4];
};

// Cubes are dealt to the CTAs in runs of consecutive cubes. 3-D float: pairs — x neighbours, i.e. the two 64-byte halves of the
// same 128-byte lines, are then written back to back by one SM (cfg2 decompress 0.1737 -> 0.1671 ms; runs of four: 0.1776).
template<int Dims> struct decode_run { static constexpr uint32_t value = Dims == 3 ? 2u : 1u; };

template<int Dims, int G>
__global__ void __launch_bounds__((4 * G + 1) * 32, 1) decompress_ws_kernel(const decompress_launch a, const __grid_constant__ CUtensorMap out_map) {
    using Bits = uint32_t;
    using tr = codec_traits<Bits>;
    constexpr int slot_bytes = decode_plan<Bits>::buffer_bytes;  // 17 KiB: the value tile (16 KiB) + 1 KiB, enough for the longest image
    constexpr int S = (232448 - 2048) / slot_bytes;
    constexpr int slot_words = slot_bytes / 4;
    static_assert(G <= 8 && S > G, "ring too small");
    static_assert(sizeof(dws_aux<S>) <= 2048, "aux area too small");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t *slots = reinterpret_cast<uint32_t *>(smem_raw);
    auto &aux = *reinterpret_cast<dws_aux<S> *>(smem_raw + S * slot_bytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Bits *stream_cubes = static_cast<const Bits *>(a.stream_cubes);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&aux.full[s], 1);
            ptx::mbar_init(&aux.empty[s], 1);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();  // the only CTA-wide barrier
    constexpr uint32_t kDecRun = decode_run<Dims>::value;
    // this CTA's cubes (cubes are independent: static round-robin over runs of kDecRun consecutive cubes):
    // k = 0 .. K-1 -> t = kDecRun * (blockIdx.x + (k / kDecRun) * gridDim.x) + k % kDecRun
    auto cube_of = [&](uint32_t k) { return kDecRun * (blockIdx.x + (k / kDecRun) * gridDim.x) + k % kDecRun; };
    uint32_t K = 0;
    {
        const uint32_t runs = (a.count + kDecRun - 1) / kDecRun;  // the last one may be short
        if (blockIdx.x < runs) {
            const uint32_t mine = (runs - blockIdx.x + gridDim.x - 1) / gridDim.x;
            K = mine * kDecRun;
            const uint32_t last_run = blockIdx.x + (mine - 1) * gridDim.x;
            if (last_run == runs - 1 && a.count % kDecRun) K -= kDecRun - a.count % kDecRun;
        }
    }

    if (warp == 4 * G) {
        // --------------------------------------------------------------------------------------- loader
        const bool first_cube_unsafe = a.hc_begin == 0
                && (reinterpret_cast<uintptr_t>(stream_cubes) & ~static_cast<uintptr_t>(15)) < reinterpret_cast<uintptr_t>(a.offsets);
        // offsets of cube k: even lanes `begin`, odd lanes `end` (a lane-dependent address keeps the value out of the uniform
        // registers until it is used, see decompress_kernel), fetched one cube ahead
        auto offsets_of = [&](uint32_t k) -> uint32_t {
            uint32_t v = 0;
            if (k < K) {
                const uint32_t hc = a.hc_begin + cube_of(k);
                const uint32_t odd = lane & 1;
                if (odd || hc) v = __ldg(a.offsets + hc - 1 + odd);  // reference src/ndzip/common.hh:350-358
            }
            return v;
        };
        uint32_t cur = offsets_of(0);
        int s = 0;
        uint32_t parity = 1;  // parity of the phase of empty[s] that ends the previous round (none in round 0)
        bool first_round = true;
        for (uint32_t k = 0; k < K; ++k) {
            const uint32_t next = offsets_of(k + 1);
            const uint32_t begin = __shfl_sync(kFullMask, cur, 0), end = __shfl_sync(kFullMask, cur, 1);
            cur = next;
            if (!first_round) {
                if (lane == 0) ptx::mbar_wait(&aux.empty[s], parity);
                __syncwarp();
            }
            const uint32_t t = cube_of(k);
            uint32_t len = end - begin;
            if (len > static_cast<uint32_t>(tr::max_cube_words)) len = tr::max_cube_words;  // corrupt header: stay inside the slot
            const uint32_t *src = reinterpret_cast<const uint32_t *>(stream_cubes + begin);
            const uint32_t shift = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src) >> 2) & 3u;
            uint32_t *slot = slots + s * slot_words;
            // the bulk copy moves whole 16-byte blocks: it must not run past the end of the stream (last cube of the range)
            // nor start in front of the caller's buffer (first cube behind a very short header)
            const bool bulk = t + 1 < a.count && !(first_cube_unsafe && t == 0);
            if (bulk) {
                if (lane == 0) {
                    aux.seq[s] = k;
                    aux.shift[s] = shift;
                    const uint32_t bytes = ((shift + len) * 4u + 15u) & ~15u;
                    ptx::mbar_arrive_expect_tx(&aux.full[s], bytes);
                    ptx::bulk_load(slot, src - shift, bytes, &aux.full[s]);
                }
            } else {
                for (uint32_t w = lane; w < len; w += 32) slot[shift + w] = __ldg(src + w);
                __syncwarp();
                if (lane == 0) {
                    aux.seq[s] = k;
                    aux.shift[s] = shift;
                    ptx::mbar_arrive(&aux.full[s]);  // (release: the lanes' stores above are ordered before it by __syncwarp)
                }
            }
            if (++s == S) {
                s = 0;
                parity ^= 1u;
                first_round = false;
            }
        }
    } else {
        // --------------------------------------------------------------------------------------- decode group
        const int g = warp >> 2, u = tid & (kCubeThreads - 1);
        int prev_slot = -1;
        auto release_previous = [&] {
            // the group's first thread issued the previous cube's tensor store: once the store has read the tile, the slot is free
            if (u == 0 && prev_slot >= 0) {
                ptx::tma_store_wait_read();
                ptx::mbar_arrive(&aux.empty[prev_slot]);
                prev_slot = -1;
            }
        };
        for (uint32_t k = g; k < K; k += G) {
            const int s = static_cast<int>(k % static_cast<uint32_t>(S));
            const uint32_t parity = (k / static_cast<uint32_t>(S)) & 1u;
            do {  // parity wait + sequence tag, see compress_ws_kernel
                ptx::mbar_wait(&aux.full[s], parity);
            } while (*reinterpret_cast<volatile uint32_t *>(&aux.seq[s]) != k);
            uint32_t *tile = slots + s * slot_words;
            const uint32_t *image = tile + aux.shift[s];
            const uint32_t hc = a.hc_begin + cube_of(k);
            // 2-D: the four segments' column totals live in the slot's last KiB (the image is dead by the time they are written)
            auto segment_total = reinterpret_cast<Bits(*)[64]>(tile + 4096);
            decode_cube<Bits, Dims, store_path::tma>(tile, image, aux.warp_total[g], aux.warp_sum[g], segment_total, a, hc, u,
                    [&] { ptx::named_barrier(1 + g, kCubeThreads); }, release_previous);
            ptx::named_barrier(1 + g, kCubeThreads);  // everybody's tile writes and proxy fences are behind this barrier
            if (u == 0) {
                issue_tma_store<Bits, Dims>(tile, &out_map, a.geom, hc);
                prev_slot = s;
            }
        }
        release_previous();
        if (u == 0) ptx::tma_store_wait_all();  // shared memory must outlive the stores that read it
    }
}

// =====================================================================================================
// border, small utilities
// =====================================================================================================

template<typename Bits>
__global__ void pack_border_kernel(
        const Bits *data, border_geom bg, Bits *stream_words, uint64_t border_base, const uint32_t *total_words) {
    const uint64_t base = border_base + (total_words ? *total_words : 0u);
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < bg.count;
            i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        stream_words[base + i] = data[border_linear_index(bg, i)];
    }
}

template<typename Bits>
__global__ void unpack_border_kernel(const Bits *stream_words, const uint32_t *offsets, uint32_t num_cubes,
        uint64_t header_words, border_geom bg, Bits *data) {
    const uint64_t base = header_words + (num_cubes ? offsets[num_cubes - 1] : 0u);  // common.hh:365
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < bg.count;
            i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        data[border_linear_index(bg, i)] = stream_words[base + i];
    }
}

__global__ void add_offset_kernel(uint32_t *offsets, uint32_t count, const uint32_t *base) {
    const uint32_t b = *base;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) offsets[i] += b;
}

// Multi-GPU header fix-up in one launch: every rank holds the all-gathered stream lengths (words) of all
// ranks; its cubes start after the compressed cube words of the lower ranks.
__global__ void fixup_header_kernel(const uint32_t *local_header, uint32_t *global_header, uint32_t count,
        const uint32_t *gathered_lengths, const uint32_t *overhead_words, uint32_t rank) {
    uint32_t base = 0;
    for (uint32_t r = 0; r < rank; ++r) base += gathered_lengths[r] - overhead_words[r];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        global_header[i] = local_header[i] + base;
    }
}

__global__ void store_length_kernel(uint32_t *length_out, uint32_t value, const uint32_t *plus) {
    *length_out = value + (plus ? *plus : 0u);
}

// =====================================================================================================
// device-side self tests of the scan primitives (counterpart of the reference's src/test/cuda_bits_test.cu:37-114,
// which tests warp / block / hierarchical scans in isolation): the same look_back / look_back_blocks / warp_inclusive_sum
// code the compress kernel runs, driven by a list of lengths instead of encoded cubes
// =====================================================================================================

// One warp per CTA, persistent; ticket order == item order as in the compress kernel. Item t has length lengths[t]; the
// warp publishes it after a pseudo-random delay (skew between "SMs"), resolves its exclusive offset with the look-back
// under test and publishes the inclusive one. Mode 0: two-level (blocks of 32), 1 / 2: windows of 32 / 64.
template<int Mode>
__global__ void __launch_bounds__(32) selftest_lookback_kernel(const uint32_t *lengths, uint32_t count, uint32_t *exclusive_out, uint64_t *desc,
        unsigned long long *blocks, uint32_t *ticket, uint32_t ticket_base, uint32_t epoch, uint32_t base, uint32_t *watch) {
    const int lane = threadIdx.x;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u) - ticket_base;
        t = __shfl_sync(kFullMask, t, 0);
        if (t >= count) return;
        const uint32_t len = lengths[t];
        const long long until = clock64() + ((t * 2654435761u) >> 20);  // 0 .. 4095 cycles of skew
        while (clock64() < until) {}
        if (lane == 0) {
            ptx::st_relaxed_gpu(desc + static_cast<size_t>(t) * kDescStride, pack_desc(epoch, kStatusAggregate, len));
            if (Mode == 0) atomicAdd(blocks + static_cast<size_t>(t >> kBlockShift) * kDescStride, static_cast<unsigned long long>(block_contribution(len)));
        }
        uint32_t exclusive = base;
        bool aborted = false;
        if (t != 0) {
            if constexpr (Mode == 0) {
                exclusive = look_back_blocks(desc, reinterpret_cast<const uint64_t *>(blocks), t, epoch, lane, base, watch, &aborted, nullptr);
            } else {
                const look_back_window<Mode> first = look_back_load<Mode>(desc, static_cast<int64_t>(t) - 1, epoch, lane, base);
                exclusive = look_back<Mode>(desc, t, epoch, lane, first, base, watch, &aborted, nullptr);
            }
        }
        if (aborted) return;
        if (lane == 0) {
            ptx::st_relaxed_gpu(desc + static_cast<size_t>(t) * kDescStride, pack_desc(epoch, kStatusPrefix, exclusive + len));
            exclusive_out[t] = exclusive;
        }
    }
}

// out[i] = inclusive sum of in[] over the lanes of i's warp up to i (the scan every encoder / decoder warp runs)
__global__ void selftest_warp_scan_kernel(const uint32_t *in, uint32_t *out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    const uint32_t s = warp_inclusive_sum(v, threadIdx.x & 31);
    if (i < n) out[i] = s;
}

// ---- kernel tables ----------------------------------------------------------------------------------

using compress_fn = void (*)(const compress_launch, const CUtensorMap);
using decompress_fn = void (*)(const decompress_launch, const CUtensorMap);

template<typename Bits, int Dims>
compress_fn compress_for(load_path p) {
    switch (p) {
#if defined(NDZB_TUNING)
        case load_path::tma: return compress_kernel<Bits, Dims, load_path::tma>;
#else
        case load_path::tma: return nullptr;  // TMA-compatible inputs always take compress_ws_kernel
#endif
        case load_path::vec16: return compress_kernel<Bits, Dims, load_path::vec16>;
        default: return compress_kernel<Bits, Dims, load_path::scalar>;
    }
}
compress_fn compress_entry(int dtype, int dims, load_path p) {
    if (dtype == 0) {
        return dims == 1 ? compress_for<uint32_t, 1>(p) : dims == 2 ? compress_for<uint32_t, 2>(p) : compress_for<uint32_t, 3>(p);
    }
    return dims == 1 ? compress_for<uint64_t, 1>(p) : dims == 2 ? compress_for<uint64_t, 2>(p) : compress_for<uint64_t, 3>(p);
}
template<typename Bits, int Dims>
decompress_fn decompress_for(int out) {
    switch (out) {
        case 2: return decompress_kernel<Bits, Dims, store_path::tma>;
        case 1: return decompress_kernel<Bits, Dims, store_path::vec16>;
        default: return decompress_kernel<Bits, Dims, store_path::scalar>;
    }
}
decompress_fn decompress_entry(int dtype, int dims, int out) {
    if (dtype == 0) {
        return dims == 1 ? decompress_for<uint32_t, 1>(out) : dims == 2 ? decompress_for<uint32_t, 2>(out) : decompress_for<uint32_t, 3>(out);
    }
    return dims == 1 ? decompress_for<uint64_t, 1>(out) : dims == 2 ? decompress_for<uint64_t, 2>(out) : decompress_for<uint64_t, 3>(out);
}

using compress_ws_fn = void (*)(const compress_launch, const CUtensorMap);

// (encoder groups, retire warps) per CTA; variant 0 is the default, the others exist for tuning runs
// (NDZB_WS_VARIANT) and are documented with their measurements in profiles/README.md
struct ws_variant {
    int groups, retire;
    int look_back_depth;  // windows of 32 * depth cubes per round trip (negative: read with ld.global.cg)
    int ticket_lookahead;  // tickets a loader holds ahead of its loads; -2: two, drawn as a pair of consecutive cubes
    int copiers;  // copy warps: 0 = the retire warps copy their cubes out themselves
    int early;  // look-back started when the cube's length is known (1) / when its image is complete (0) / 1 for 3-D profiles only (2)
    bool dynamic;  // encoder groups take the CTA's next cube when they become free (instead of cube g, g+G, ...)
    bool stats;
};
// Variant 0 is what the library uses; 1-4 are kept for A/B runs and so that the tests cover every code path (late
// look-back with 64-cube windows, dynamic assignment, weak descriptor loads + late ticket binding, prefetch limit +
// statistics). Measured on
// B200 with descriptors one per 64 bytes (profiles/README.md): float 5 groups + 4 retire warps + two-level look-back
// (3-D 0.193 ms, 1-D 0.297 ms per GiB), double 3 + 2 with 32-cube windows (2-D 0.199 ms).
#if defined(NDZB_TUNING) && defined(NDZB_TUNING_SWEEP2)
// second sweep of (groups, retire warps), with paired tickets, after the encoder got lighter (shuffle stencil)
constexpr ws_variant kWsVariants32[] = {{5, 4, 0, -2, 0, 1, false, false}, {5, 5, 0, -2, 0, 1, false, false}, {4, 5, 0, -2, 0, 1, false, false},
        {4, 4, 0, -2, 0, 1, false, false}, {5, 4, 0, 1, 0, 1, false, true}, {5, 3, 0, -2, 0, 1, false, false}, {4, 6, 0, -2, 0, 1, false, false},
        {5, 4, 1, -2, 0, 1, false, false}};
constexpr ws_variant kWsVariants64[] = {{3, 2, 1, 1, 0, 1, false, false}, {3, 3, 1, 1, 0, 1, false, false}, {3, 4, 1, 1, 0, 1, false, false},
        {2, 3, 1, 1, 0, 1, false, false}, {3, 2, 1, 1, 0, 1, false, true}, {3, 2, 0, 1, 0, 1, false, false}, {3, 3, 0, 1, 0, 1, false, false},
        {3, 2, 2, 1, 0, 1, false, false}};
#elif defined(NDZB_TUNING)
constexpr ws_variant kWsVariants32[] = {{5, 4, 0, 1, 0, 1, false, false}, {5, 4, 0, 1, 3, 1, false, false}, {5, 4, 0, 1, 2, 1, false, false},
        {5, 3, 0, 1, 3, 1, false, false}, {5, 4, 0, 1, 0, 1, false, true}, {4, 4, 0, 1, 4, 1, false, false}, {5, 3, 0, 1, 4, 1, false, false},
        {5, 5, 0, 1, 2, 1, false, false}};
constexpr ws_variant kWsVariants64[] = {{3, 2, 1, 1, 0, 1, false, false}, {4, 2, 1, 1, 0, 1, false, false}, {4, 3, 1, 1, 0, 1, false, false},
        {3, 3, 1, 1, 2, 1, false, false}, {3, 2, 1, 1, 0, 1, false, true}, {3, 2, 1, 1, 3, 1, false, false}, {3, 2, 1, 2, 0, 1, false, false},
        {4, 2, 0, 1, 0, 1, false, false}};
#else
#ifndef NDZB_LA
#define NDZB_LA 1
#endif
#ifndef NDZB_LA32
#define NDZB_LA32 -2  // float: pairs of consecutive tickets for 3-D (falls back to 1 for 1-D / 2-D)
#endif
constexpr ws_variant kWsVariants32[] = {{5, 4, 0, NDZB_LA32, 0, 1, false, false}};
constexpr ws_variant kWsVariants64[] = {{3, 2, 1, NDZB_LA, 0, 1, false, false}};
#endif
constexpr int kNumWsVariants32 = sizeof(kWsVariants32) / sizeof(ws_variant);
constexpr int kNumWsVariants64 = sizeof(kWsVariants64) / sizeof(ws_variant);

template<typename Bits, int Dims, int V>
compress_ws_fn compress_ws_variant_fn() {
    if constexpr (sizeof(Bits) == 4) {
        constexpr ws_variant v = kWsVariants32[V];
        // pairs of consecutive tickets pay for 3-D float only (64-byte rows: x neighbours share 128-byte lines), cf. ws_lookahead()
        return compress_ws_kernel<Bits, Dims, v.groups, v.retire, v.look_back_depth, (v.ticket_lookahead < -1 && Dims != 3) ? 1 : v.ticket_lookahead, v.copiers,
                v.early == 1 || (v.early == 2 && Dims == 3), v.dynamic, v.stats>;
    } else {
        constexpr ws_variant v = kWsVariants64[V];
        return compress_ws_kernel<Bits, Dims, v.groups, v.retire, v.look_back_depth, v.ticket_lookahead, v.copiers,
                v.early == 1 || (v.early == 2 && Dims == 3), v.dynamic, v.stats>;
    }
}

template<typename Bits, int Dims>
compress_ws_fn compress_ws_for(int variant) {
#if defined(NDZB_TUNING)
    switch (variant) {
        case 1: return compress_ws_variant_fn<Bits, Dims, 1>();
        case 2: return compress_ws_variant_fn<Bits, Dims, 2>();
        case 3: return compress_ws_variant_fn<Bits, Dims, 3>();
        case 4: return compress_ws_variant_fn<Bits, Dims, 4>();
        case 5: return compress_ws_variant_fn<Bits, Dims, 5>();
        case 6: return compress_ws_variant_fn<Bits, Dims, 6>();
        case 7: return compress_ws_variant_fn<Bits, Dims, 7>();
        default: break;
    }
#endif
    (void) variant;
    return compress_ws_variant_fn<Bits, Dims, 0>();
}
compress_ws_fn compress_ws_entry(int dtype, int dims, int variant) {
    if (dtype == 0) {
        return dims == 1 ? compress_ws_for<uint32_t, 1>(variant) : dims == 2 ? compress_ws_for<uint32_t, 2>(variant) : compress_ws_for<uint32_t, 3>(variant);
    }
    return dims == 1 ? compress_ws_for<uint64_t, 1>(variant) : dims == 2 ? compress_ws_for<uint64_t, 2>(variant) : compress_ws_for<uint64_t, 3>(variant);
}
size_t compress_ws_smem(int dtype) { return dtype == 0 ? ws_smem_bytes<uint32_t>() : ws_smem_bytes<uint64_t>(); }

size_t compress_smem(int dtype) { return dtype == 0 ? compress_smem_bytes<uint32_t>() : compress_smem_bytes<uint64_t>(); }
size_t decompress_smem(int dtype) { return dtype == 0 ? decompress_smem_bytes<uint32_t>() : decompress_smem_bytes<uint64_t>(); }

}  // namespace

// =====================================================================================================
// host side
// =====================================================================================================

uint32_t compress_ticket_overdraw(uint32_t grid) {
    return grid;  // every CTA draws one ticket up front and one more per cube it processes
}
int compress_ws_variants(int dtype);
// effective ticket look-ahead of (dtype, dims, variant): -2 = pairs (see compress_ws_variant_fn)
static int ws_lookahead(int dtype, int dims, int variant) {
    if (variant < 0 || variant >= compress_ws_variants(dtype)) variant = 0;
    const int la = (dtype == 0 ? kWsVariants32[variant] : kWsVariants64[variant]).ticket_lookahead;
    return la < -1 && !(dtype == 0 && dims == 3) ? 1 : la;
}
// Tickets a launch draws beyond `count` (the host mirrors the device's free-running ticket counter).
uint32_t compress_ws_ticket_overdraw(int dtype, int dims, int variant, uint32_t grid, uint32_t count) {
    const int la = ws_lookahead(dtype, dims, variant);
    // pairs: every loader draws one pair up front and one more per fully valid pair it loads: 2 * grid + 2 * floor(count / 2) in all
    if (la < -1) return static_cast<uint32_t>(-la) * grid - count % static_cast<uint32_t>(-la);  // runs of -la tickets: la * (grid + floor(count / la)) in all
    return grid * static_cast<uint32_t>(la > 0 ? la : 1);  // the look-ahead, or the one ticket that ends a late-binding loader
}
int compress_ws_variants(int dtype) { return dtype == 0 ? kNumWsVariants32 : kNumWsVariants64; }
bool tuning_build() { return kTuning; }
bool compress_ws_uses_blocks(int dtype, int variant) {
    if (variant < 0 || variant >= compress_ws_variants(dtype)) variant = 0;
    return (dtype == 0 ? kWsVariants32[variant] : kWsVariants64[variant]).look_back_depth == 0;
}

cudaError_t configure_kernels(kernel_config &cfg) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&cfg.num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    for (int dtype = 0; dtype < 2; ++dtype) {
        for (int dims = 1; dims <= 3; ++dims) {
            for (int p = 0; p < 3; ++p) {
                auto fn = compress_entry(dtype, dims, static_cast<load_path>(p));
                if (!fn) continue;
                err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(compress_smem(dtype)));
                if (err != cudaSuccess) return err;
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                        &cfg.ctas_per_sm[dtype][dims - 1][p], fn, kCubeThreads, compress_smem(dtype));
                if (err != cudaSuccess) return err;
            }
            for (int v = 0; v < compress_ws_variants(dtype); ++v) {
                err = cudaFuncSetAttribute(compress_ws_entry(dtype, dims, v), cudaFuncAttributeMaxDynamicSharedMemorySize,
                        static_cast<int>(compress_ws_smem(dtype)));
                if (err != cudaSuccess) return err;
            }
            if (decompress_ws_available(dtype, dims)) {
                err = dims == 2 ? cudaFuncSetAttribute(decompress_ws_kernel<2, kDecodeGroups>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)
                                : cudaFuncSetAttribute(decompress_ws_kernel<3, kDecodeGroups>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
                if (err != cudaSuccess) return err;
            }
            for (int v = 0; v < 3; ++v) {
                auto fn = decompress_entry(dtype, dims, v);
                err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(decompress_smem(dtype)));
                if (err != cudaSuccess) return err;
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                        &cfg.dec_ctas_per_sm[dtype][dims - 1][v], fn, kCubeThreads, decompress_smem(dtype));
                if (err != cudaSuccess) return err;
            }
        }
    }
    return cudaSuccess;
}

cudaError_t launch_compress(int dtype, int dims, load_path path, const compress_launch &args, const CUtensorMap *tmap,
        uint32_t grid, cudaStream_t stream) {
    static const CUtensorMap dummy{};
    if (!compress_entry(dtype, dims, path)) return cudaErrorInvalidDeviceFunction;
    compress_entry(dtype, dims, path)<<<grid, kCubeThreads, compress_smem(dtype), stream>>>(args, tmap ? *tmap : dummy);
    return cudaGetLastError();
}

cudaError_t launch_compress_ws(int dtype, int dims, int variant, const compress_launch &args, const CUtensorMap &in_map,
        uint32_t grid, cudaStream_t stream) {
    if (variant < 0 || variant >= compress_ws_variants(dtype)) variant = 0;
    const ws_variant v = dtype == 0 ? kWsVariants32[variant] : kWsVariants64[variant];
    const uint32_t threads = static_cast<uint32_t>(4 * v.groups + 1 + v.retire + v.copiers) * 32u;
    compress_ws_entry(dtype, dims, variant)<<<grid, threads, compress_ws_smem(dtype), stream>>>(args, in_map);
    return cudaGetLastError();
}

bool decompress_ws_available(int dtype, int dims) { return dtype == 0 && dims >= 2; }  // (1-D float needs 80 registers: 6 CTAs of decompress_kernel)

cudaError_t launch_decompress_ws(int dims, const decompress_launch &args, const CUtensorMap &out_map, uint32_t grid, cudaStream_t stream) {
    constexpr uint32_t threads = (4 * kDecodeGroups + 1) * 32;
    constexpr size_t smem = 232448;
    if (dims == 2) decompress_ws_kernel<2, kDecodeGroups><<<grid, threads, smem, stream>>>(args, out_map);
    else decompress_ws_kernel<3, kDecodeGroups><<<grid, threads, smem, stream>>>(args, out_map);
    return cudaGetLastError();
}

cudaError_t launch_decompress(int dtype, int dims, int store, const decompress_launch &args, const CUtensorMap *out_map, uint32_t grid,
        cudaStream_t stream) {
    static const CUtensorMap dummy{};
    decompress_entry(dtype, dims, store)<<<grid, kCubeThreads, decompress_smem(dtype), stream>>>(args, out_map ? *out_map : dummy);
    return cudaGetLastError();
}

static uint32_t border_grid(uint64_t count) {
    const uint64_t blocks = (count + 255) / 256;
    return static_cast<uint32_t>(blocks < 148u * 16u ? (blocks ? blocks : 1) : 148u * 16u);
}

cudaError_t launch_pack_border(int dtype, const void *data, const border_geom &bg, void *stream_words,
        uint64_t border_base, const uint32_t *total_words, cudaStream_t stream) {
    if (dtype == 0) {
        pack_border_kernel<uint32_t><<<border_grid(bg.count), 256, 0, stream>>>(
                static_cast<const uint32_t *>(data), bg, static_cast<uint32_t *>(stream_words), border_base, total_words);
    } else {
        pack_border_kernel<uint64_t><<<border_grid(bg.count), 256, 0, stream>>>(
                static_cast<const uint64_t *>(data), bg, static_cast<uint64_t *>(stream_words), border_base, total_words);
    }
    return cudaGetLastError();
}

cudaError_t launch_unpack_border(int dtype, const void *stream_words, const uint32_t *offsets, uint32_t num_cubes,
        uint64_t header_words, const border_geom &bg, void *data, cudaStream_t stream) {
    if (dtype == 0) {
        unpack_border_kernel<uint32_t><<<border_grid(bg.count), 256, 0, stream>>>(
                static_cast<const uint32_t *>(stream_words), offsets, num_cubes, header_words, bg, static_cast<uint32_t *>(data));
    } else {
        unpack_border_kernel<uint64_t><<<border_grid(bg.count), 256, 0, stream>>>(
                static_cast<const uint64_t *>(stream_words), offsets, num_cubes, header_words, bg, static_cast<uint64_t *>(data));
    }
    return cudaGetLastError();
}

cudaError_t launch_add_offset(uint32_t *offsets, uint32_t count, const uint32_t *base, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const uint32_t blocks = (count + 255) / 256;
    add_offset_kernel<<<blocks < 1184u ? blocks : 1184u, 256, 0, stream>>>(offsets, count, base);
    return cudaGetLastError();
}

cudaError_t launch_fixup_header(const uint32_t *local_header, uint32_t *global_header, uint32_t count,
        const uint32_t *gathered_lengths, const uint32_t *overhead_words, uint32_t rank, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const uint32_t blocks = (count + 255) / 256;
    fixup_header_kernel<<<blocks < 592u ? blocks : 592u, 256, 0, stream>>>(local_header, global_header, count, gathered_lengths, overhead_words, rank);
    return cudaGetLastError();
}

cudaError_t launch_selftest_lookback(int mode, const uint32_t *lengths, uint32_t count, uint32_t *exclusive_out, uint64_t *desc,
        unsigned long long *blocks, uint32_t *ticket, uint32_t ticket_base, uint32_t epoch, uint32_t base, uint32_t *watch, uint32_t grid,
        cudaStream_t stream) {
    if (mode == 0) selftest_lookback_kernel<0><<<grid, 32, 0, stream>>>(lengths, count, exclusive_out, desc, blocks, ticket, ticket_base, epoch, base, watch);
    else if (mode == 1) selftest_lookback_kernel<1><<<grid, 32, 0, stream>>>(lengths, count, exclusive_out, desc, blocks, ticket, ticket_base, epoch, base, watch);
    else selftest_lookback_kernel<2><<<grid, 32, 0, stream>>>(lengths, count, exclusive_out, desc, blocks, ticket, ticket_base, epoch, base, watch);
    return cudaGetLastError();
}

cudaError_t launch_selftest_warp_scan(const uint32_t *in, uint32_t *out, uint32_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    selftest_warp_scan_kernel<<<(n + 255) / 256, 256, 0, stream>>>(in, out, n);
    return cudaGetLastError();
}

cudaError_t launch_store_length(uint32_t *length_out, uint32_t value, const uint32_t *plus, cudaStream_t stream) {
    store_length_kernel<<<1, 1, 0, stream>>>(length_out, value, plus);
    return cudaGetLastError();
}

// ---- TMA tensor maps --------------------------------------------------------------------------------

bool tma_compatible(int dtype, int dims, const void *data, const grid_geom &g) {
    if (reinterpret_cast<uintptr_t>(data) % 16 != 0 || g.num_cubes == 0) return false;
    if (dims == 1) return true;                       // rows of 128 bytes at 128-byte pitch
    const uint32_t nx = g.n[2];
    return dtype == 0 ? nx % 4 == 0 : nx % 2 == 0;    // every global stride must be a multiple of 16 bytes
}

namespace {
using encode_tiled_fn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess
                || q != cudaDriverEntryPointSuccess) {
            p = nullptr;
        }
        return reinterpret_cast<encode_tiled_fn>(p);
    }();
    return fn;
}
}  // namespace

CUresult make_input_tensor_map(CUtensorMap *map, int dtype, int dims, const void *data, const grid_geom &g) {
    const auto encode = get_encode_tiled();
    if (!encode) return CUDA_ERROR_NOT_SUPPORTED;
    const uint64_t n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];  // slowest .. fastest (1-padded in front)
    cuuint64_t gdim[5] = {1, 1, 1, 1, 1};
    cuuint64_t gstride[4] = {16, 16, 16, 16};  // bytes, for dimensions 1..rank-1
    cuuint32_t box[5] = {1, 1, 1, 1, 1};
    cuuint32_t estride[5] = {1, 1, 1, 1, 1};
    cuuint32_t rank = 0;
    CUtensorMapDataType type;
    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
    if (dtype == 0) {
        type = CU_TENSOR_MAP_DATA_TYPE_UINT32;
        if (dims == 1) {            // [rows of 32][32]
            rank = 2;
            gdim[0] = 32; gdim[1] = n2 / 32;
            gstride[0] = 128;
            box[0] = 32; box[1] = 128;
        } else if (dims == 2) {     // [y][x / 32][32]
            rank = 3;
            gdim[0] = 32; gdim[1] = n2 / 32; gdim[2] = n1;
            gstride[0] = 128; gstride[1] = n2 * 4;
            box[0] = 32; box[1] = 2; box[2] = 64;
        } else {                    // [y parity][z][y / 2][x] (strides need not be monotonic); 64-byte inner rows -> SWIZZLE_64B
            rank = 4;
            gdim[0] = n2; gdim[1] = n1 / 2; gdim[2] = n0; gdim[3] = 2;
            gstride[0] = n2 * 8; gstride[1] = n1 * n2 * 4; gstride[2] = n2 * 4;
            box[0] = 16; box[1] = 8; box[2] = 16; box[3] = 2;
            swizzle = CU_TENSOR_MAP_SWIZZLE_64B;
        }
    } else {
        type = CU_TENSOR_MAP_DATA_TYPE_UINT64;
        if (dims == 1) {            // [runs][half][16]
            rank = 3;
            gdim[0] = 16; gdim[1] = 2; gdim[2] = n2 / 32;
            gstride[0] = 128; gstride[1] = 256;
            box[0] = 16; box[1] = 1; box[2] = 128;
        } else if (dims == 2) {     // [y][x / 32][half][16]
            rank = 4;
            gdim[0] = 16; gdim[1] = 2; gdim[2] = n2 / 32; gdim[3] = n1;
            gstride[0] = 128; gstride[1] = 256; gstride[2] = n2 * 8;
            box[0] = 16; box[1] = 1; box[2] = 2; box[3] = 64;
        } else {                    // [z][y / 2][y parity][x]
            rank = 4;
            gdim[0] = n2; gdim[1] = 2; gdim[2] = n1 / 2; gdim[3] = n0;
            gstride[0] = n2 * 8; gstride[1] = n2 * 16; gstride[2] = n1 * n2 * 8;
            box[0] = 16; box[1] = 1; box[2] = 8; box[3] = 16;
        }
    }
    return encode(map, type, rank, const_cast<void *>(data), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// Tensor map of the decoder's output tile (the value tile as the last pass leaves it), per profile:
//   float 1D / 2D  the input views ([rows][32] / [y][x / 32][32], SWIZZLE_128B): the decoder's run rows
//   float 3D       plain [z][y][x], box 16 x 16 x 16, 64-byte rows under SWIZZLE_64B (re-laid out by the last pass)
//   double         the two 16 KiB half regions as the slowest view dimension, so that ONE store covers both
CUresult make_output_tensor_map(CUtensorMap *map, int dtype, int dims, const void *data, const grid_geom &g) {
    const auto encode = get_encode_tiled();
    if (!encode) return CUDA_ERROR_NOT_SUPPORTED;
    const uint64_t n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    cuuint64_t gdim[5] = {1, 1, 1, 1, 1};
    cuuint64_t gstride[4] = {16, 16, 16, 16};
    cuuint32_t box[5] = {1, 1, 1, 1, 1};
    cuuint32_t estride[5] = {1, 1, 1, 1, 1};
    cuuint32_t rank = 0;
    CUtensorMapDataType type;
    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
    if (dtype == 0) {
        type = CU_TENSOR_MAP_DATA_TYPE_UINT32;
        if (dims == 1) {
            rank = 2;
            gdim[0] = 32; gdim[1] = n2 / 32;
            gstride[0] = 128;
            box[0] = 32; box[1] = 128;
        } else if (dims == 2) {
            rank = 3;
            gdim[0] = 32; gdim[1] = n2 / 32; gdim[2] = n1;
            gstride[0] = 128; gstride[1] = n2 * 4;
            box[0] = 32; box[1] = 2; box[2] = 64;
        } else {
            rank = 3;
            gdim[0] = n2; gdim[1] = n1; gdim[2] = n0;
            gstride[0] = n2 * 4; gstride[1] = n1 * n2 * 4;
            box[0] = 16; box[1] = 16; box[2] = 16;
            swizzle = CU_TENSOR_MAP_SWIZZLE_64B;
        }
    } else {
        type = CU_TENSOR_MAP_DATA_TYPE_UINT64;
        if (dims == 1) {            // [half][runs][16]
            rank = 3;
            gdim[0] = 16; gdim[1] = n2 / 32; gdim[2] = 2;
            gstride[0] = 256; gstride[1] = 128;
            box[0] = 16; box[1] = 128; box[2] = 2;
        } else if (dims == 2) {     // [half][y][x / 32][16]
            rank = 4;
            gdim[0] = 16; gdim[1] = n2 / 32; gdim[2] = n1; gdim[3] = 2;
            gstride[0] = 256; gstride[1] = n2 * 8; gstride[2] = 128;
            box[0] = 16; box[1] = 2; box[2] = 64; box[3] = 2;
        } else {                    // [y parity][z][y / 2][x]
            rank = 4;
            gdim[0] = n2; gdim[1] = n1 / 2; gdim[2] = n0; gdim[3] = 2;
            gstride[0] = n2 * 16; gstride[1] = n1 * n2 * 8; gstride[2] = n2 * 8;
            box[0] = 16; box[1] = 8; box[2] = 16; box[3] = 2;
        }
    }
    return encode(map, type, rank, const_cast<void *>(data), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace ndzb
