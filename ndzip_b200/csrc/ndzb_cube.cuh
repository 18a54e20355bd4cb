// ndzb_cube.cuh — per-thread building blocks of the hypercube codec.
//
// Everything here is __host__ __device__ and free of warp intrinsics, so the exact code the kernels
// run can also be driven lane by lane from the CPU simulation in tests/host_sim (which checks it
// bit for bit against the oracle without a GPU). The kernels in ndzb_kernels.cu
// add the CTA-level parts: TMA / global loads, scans, the decoupled look-back, barriers.
//
// Work decomposition (differs from the reference's one-warp-per-chunk scheme,
// reference src/ndzip/cuda_codec.inl:204-274): a cube of 4096 elements is handled by 128 threads;
// thread u owns "run" u = the 32 consecutive cube-local elements [32u, 32u+32) and
//   * computes their Lorenzo residuals straight from shared memory (reads its own run plus the
//     1 / 2 / 5 neighbouring half-runs the stencil needs) — no in-place multi-pass transform,
//   * bit-transposes them in registers with a 5-stage butterfly (32x32 bits per thread),
//   * compacts the non-zero bit planes straight from registers into the cube's compressed image in
//     shared memory (in place over the dead input tile), which is then copied out linearly.
// float : run u == chunk u (32 values x 32 bits).
// double: chunk c (64 values x 64 bits) == runs 2c, 2c+1; each thread transposes the high and the
//         low words of its 32 values (two 32x32 transposes) and owns one 32-bit half of every plane.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define NDZB_HD __host__ __device__ __forceinline__
#else
#define NDZB_HD inline
#endif

namespace ndzb {

constexpr int kCubeElems = 4096;       // reference src/ndzip/common.hh:368-381
constexpr int kCubeThreads = 128;      // threads (= runs) per cube
constexpr int kRunElems = 32;          // elements per run

template<int Dims> struct side_of;     // reference src/ndzip/common.hh:371-378
template<> struct side_of<1> { static constexpr int value = 4096; };
template<> struct side_of<2> { static constexpr int value = 64; };
template<> struct side_of<3> { static constexpr int value = 16; };

template<typename Bits> struct codec_traits;
template<> struct codec_traits<uint32_t> {
    static constexpr int bits = 32;
    static constexpr int chunks = 128;               // chunks per cube (cuda_codec.inl:190-194)
    static constexpr int words32_per_elem = 1;
    static constexpr int cube_words32 = 4096;        // input tile, 32-bit words
    static constexpr int max_cube_words = 4224;      // compressed bound in Bits words (common.hh:391-392)
    static constexpr int image_words32 = 4224;       // compressed cube image in shared memory
};
template<> struct codec_traits<uint64_t> {
    static constexpr int bits = 64;
    static constexpr int chunks = 64;
    static constexpr int words32_per_elem = 2;
    static constexpr int cube_words32 = 8192;
    static constexpr int max_cube_words = 4160;
    static constexpr int image_words32 = 2 * 4160;
};

// ------------------------------------------------------------------------------------------------
// bit helpers

NDZB_HD uint32_t rotl1(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(v, v, 1);
#else
    return (v << 1) | (v >> 31);
#endif
}
NDZB_HD uint64_t rotl1(uint64_t v) { return (v << 1) | (v >> 63); }
NDZB_HD uint32_t rotr1(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(v, v, 1);
#else
    return (v >> 1) | (v << 31);
#endif
}
NDZB_HD uint64_t rotr1(uint64_t v) { return (v >> 1) | (v << 63); }

// reference src/ndzip/common.hh:446-449: negative values have their lower bits flipped, v ^ (sign ? 0x7f..f : 0).
// For a negative v that is 0x7f..f - v = v * -1 + 0x7f..f. On the device: a sign test (alu pipe) plus a PREDICATED
// multiply-add (fma pipe) instead of shift + xor (two alu-pipe slots per value) — the alu pipe is what bounds both
// kernels (scripts/ubench/pipes.cu; profiles/README.md round 2). The two constants are read from constant memory
// because ptxas folds literal ones back into an alu-pipe IADD3.
#if defined(__CUDACC__)
static __constant__ uint32_t kFmaPipeConsts[2] = {0xffffffffu, 0x7fffffffu};
#endif
NDZB_HD uint32_t complement_negative(uint32_t v) {
#if defined(__CUDA_ARCH__) && !defined(NDZB_COMPLEMENT_XOR)
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %0, 0;\n\t@p mad.lo.s32 %0, %0, %1, %2;\n\t}"
            : "+r"(v) : "r"(kFmaPipeConsts[0]), "r"(kFmaPipeConsts[1]));
    return v;
#else
    return v ^ (static_cast<uint32_t>(static_cast<int32_t>(v) >> 31) & 0x7fffffffu);
#endif
}
NDZB_HD uint64_t complement_negative(uint64_t v) {
#if defined(__CUDA_ARCH__) && !defined(NDZB_COMPLEMENT_XOR)
    uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
    // low word: ~lo = lo * -1 + -1; high word: hi * -1 + 0x7fffffff (a bitwise xor: no borrow between the halves)
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\t@p mad.lo.s32 %0, %0, %2, %2;\n\t@p mad.lo.s32 %1, %1, %2, %3;\n\t}"
            : "+r"(lo), "+r"(hi) : "r"(kFmaPipeConsts[0]), "r"(kFmaPipeConsts[1]));
    return (static_cast<uint64_t>(hi) << 32) | lo;
#else
    return v ^ (static_cast<uint64_t>(static_cast<int64_t>(v) >> 63) & 0x7fffffffffffffffull);
#endif
}

NDZB_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

NDZB_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    const uint64_t xy = (static_cast<uint64_t>(y) << 32) | x;
    uint32_t r = 0;
    for (int n = 0; n < 4; ++n) r |= static_cast<uint32_t>((xy >> (8 * ((sel >> (4 * n)) & 7))) & 0xff) << (8 * n);
    return r;
#endif
}

// m ? a : b, bit by bit — ONE LOP3 (lut 0xE4). Written as (x & m) | (y & ~m) the compiler sees two different
// constants and emits two LOP3 per output word (96 extra instructions per 32x32 transpose, 3 per element:
// profiles/README.md, round 2).
NDZB_HD uint32_t bitselect(uint32_t m, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
#else
    return b ^ ((a ^ b) & m);
#endif
}

// In-register 32x32 bit-matrix transpose, LSB-indexed: afterwards bit b of a[k] is what bit k of
// a[b] was. Five butterfly stages; the 16- and 8-bit stages are byte permutes (1 PRMT per word),
// the 4/2/1-bit stages are shift + bit-select (2 ops per word). 256 ALU ops per 1024 bits, versus
// 32 ballots + 32 predicate set-ups PER CHUNK for a ballot transpose.
template<int S, uint32_t M>
NDZB_HD void butterfly_stage(uint32_t *a) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (k & S) continue;
        const uint32_t x = a[k], y = a[k + S];
        a[k] = bitselect(M, x, y << S);              // (x & M) | ((y << S) & ~M)
        a[k + S] = bitselect(M, x >> S, y);           // ((x >> S) & M) | (y & ~M)
    }
}

NDZB_HD void transpose32(uint32_t *a) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t x = a[k], y = a[k + 16];
        a[k] = byte_perm(x, y, 0x5410);        // low halves
        a[k + 16] = byte_perm(x, y, 0x7632);   // high halves
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (k & 8) continue;
        const uint32_t x = a[k], y = a[k + 8];
        a[k] = byte_perm(x, y, 0x6240);        // even bytes
        a[k + 8] = byte_perm(x, y, 0x7351);    // odd bytes
    }
    butterfly_stage<4, 0x0f0f0f0fu>(a);
    butterfly_stage<2, 0x33333333u>(a);
    butterfly_stage<1, 0x55555555u>(a);
}

// ------------------------------------------------------------------------------------------------
// shared-memory tile layout
//
// A cube is kept in 128-byte rows whose eight 16-byte units are XOR-swizzled with (row & 7) — the
// pattern CU_TENSOR_MAP_SWIZZLE_128B produces, so a TMA tensor load can deposit the tile directly.
// Row u holds run u (float: the whole run; double: 16 values, the run's other 16 values sit in row
// u of a second 16 KiB region). Thread u reading its run with eight LDS.128 is then conflict-free:
// the eight lanes of a quarter-warp touch eight different units.

// 32-bit word index of 16-byte unit `unit` of row `row`
NDZB_HD int tile_unit(int row, int unit) { return (row << 5) | ((unit ^ (row & 7)) << 2); }

// 32-bit word index (of the low word) of cube-local element e
template<typename Bits> NDZB_HD int tile_elem(int e);
template<> NDZB_HD int tile_elem<uint32_t>(int e) {
    const int run = e >> 5, j = e & 31;
    return tile_unit(run, j >> 2) + (j & 3);
}
template<> NDZB_HD int tile_elem<uint64_t>(int e) {
    const int run = e >> 5, j = e & 31;
    return ((j >> 4) << 12) + tile_unit(run, (j & 15) >> 1) + ((j & 1) << 1);
}

struct alignas(16) quad { uint32_t x, y, z, w; };

NDZB_HD quad ld_quad(const uint32_t *p) { return *reinterpret_cast<const quad *>(p); }

// Two 64-bit values from one 16-byte unit of a SHARED-memory tile. Written in C++ (a quad load whose halves are then
// combined into two 64-bit values) the compiler splits the access into two LDS.64; with threads 128 bytes apart under
// the tile swizzle, or 16 bytes apart in the decoder's column strips, each of those is a 2-way bank conflict — half of
// all shared-memory wavefronts of the double kernels (profiles/README.md, round 2, item 25). One ld.shared.v2.u64 is
// one LDS.128: conflict-free for both patterns.
NDZB_HD void ld_pair64(const uint32_t *p, uint64_t &a, uint64_t &b) {
#if defined(__CUDA_ARCH__) && !defined(NDZB_SPLIT_LDS64)
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))) : "memory");
#else
    const quad q = ld_quad(p);
    a = (static_cast<uint64_t>(q.y) << 32) | q.x;
    b = (static_cast<uint64_t>(q.w) << 32) | q.z;
#endif
}
NDZB_HD void st_quad(uint32_t *p, quad q) { *reinterpret_cast<quad *>(p) = q; }

// Input-tile layout of the compress kernel, per profile. A run is two half-runs of 16 values:
//   float 1D/2D : the run is one 128-byte row, half h = units 4h..4h+3              (SWIZZLE_128B)
//   float 3D    : a half-run is one 16-element x-row = 64 bytes; even-y rows live in region 0, odd-y
//                 rows in region 1 (8 KiB each, 64-byte rows, units XORed with (row>>1)&3 =
//                 CU_TENSOR_MAP_SWIZZLE_64B). TMA cannot lay a 64-byte inner box densely under
//                 SWIZZLE_128B, and with one region per y parity the eight lanes of a quarter-warp
//                 read eight consecutive rows, which the 64-byte swizzle spreads over all banks.
//   double      : half h = row u of region h (16 KiB each, 128-byte rows)            (SWIZZLE_128B)
template<typename Bits, int Dims>
struct input_layout {
    static constexpr bool split64 = sizeof(Bits) == 4 && Dims == 3;
    static constexpr int units_per_half = sizeof(Bits) == 4 ? 4 : 8;
    static constexpr int region_words = sizeof(Bits) == 8 ? 4096 : 2048;  // only used when halves are regions

    // 32-bit word index of 16-byte unit k of half-run `half` of run `run`
    static NDZB_HD int half_unit(int run, int half, int k) {
        if constexpr (sizeof(Bits) == 8) {
            return half * region_words + tile_unit(run, k);
        } else if constexpr (split64) {
            return half * region_words + (run << 4) + ((k ^ ((run >> 1) & 3)) << 2);
        } else {
            return tile_unit(run, half * 4 + k);
        }
    }
    // 32-bit word index (low word) of cube-local element e
    static NDZB_HD int elem(int e) {
        const int run = e >> 5, j = e & 31;
        if constexpr (sizeof(Bits) == 8) {
            return half_unit(run, j >> 4, (j & 15) >> 1) + ((j & 1) << 1);
        } else {
            return half_unit(run, j >> 4, (j & 15) >> 2) + (j & 3);
        }
    }
};

// Half-run `half` (16 values) of run `run`, bit-cast and rotated left by one (the reference fuses
// the same rotate into its load, cuda_codec.inl:51-55).
template<typename L>
NDZB_HD void load_half_rot(const uint32_t *tile, int run, int half, uint32_t *out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const quad q = ld_quad(tile + L::half_unit(run, half, k));
        out[4 * k + 0] = rotl1(q.x);
        out[4 * k + 1] = rotl1(q.y);
        out[4 * k + 2] = rotl1(q.z);
        out[4 * k + 3] = rotl1(q.w);
    }
}
template<typename L>
NDZB_HD void load_half_rot(const uint32_t *tile, int run, int half, uint64_t *out) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint64_t a, b;
        ld_pair64(tile + L::half_unit(run, half, k), a, b);
        out[2 * k + 0] = rotl1(a);
        out[2 * k + 1] = rotl1(b);
    }
}

// Last element (31) of run `run`, rotated.
template<typename L>
NDZB_HD uint32_t load_last_rot(const uint32_t *tile, int run, uint32_t) {
    return rotl1(tile[L::elem(32 * run + 31)]);
}
template<typename L>
NDZB_HD uint64_t load_last_rot(const uint32_t *tile, int run, uint64_t) {
    const int w = L::elem(32 * run + 31);
    return rotl1((static_cast<uint64_t>(tile[w + 1]) << 32) | tile[w]);
}

// a[0..16) -= b[0..16)
template<typename Bits>
NDZB_HD void sub16(Bits *a, const Bits *b) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] -= b[i];
}

// ------------------------------------------------------------------------------------------------
// forward: residuals of run u
//
// The integer Lorenzo transform of the reference (rotate, x[i] -= x[i-1] along every axis in
// Z/2^k, complement negatives; src/ndzip/common.hh:469-501, GPU version cuda_codec.inl:68-126)
// evaluated as a stencil on the read-only tile: the per-axis differences commute, so the residual
// of an element only depends on the 2^dims corner values, with out-of-cube neighbours = 0.
// r[j] is the residual of cube-local element 32u + j, complement already applied.

template<typename Bits, int Dims>
NDZB_HD void residual_run(const uint32_t *tile, int u, Bits *r) {
    using L = input_layout<Bits, Dims>;
    Bits *lo = r, *hi = r + 16;
    load_half_rot<L>(tile, u, 0, lo);
    load_half_rot<L>(tile, u, 1, hi);
    Bits left = 0;  // value preceding r[0] along x after the higher-axis differences

    if constexpr (Dims == 1) {
        // run u = elements [32u, 32u+32) of the 4096-long line
        if (u > 0) left = load_last_rot<L>(tile, u - 1, Bits{});
    } else if constexpr (Dims == 2) {
        // 64 x 64: run u = row y = u>>1, columns [32g, 32g+32) with g = u&1; the row above is run u-2
        const int y = u >> 1, g = u & 1;
        if (g) left = load_last_rot<L>(tile, u - 1, Bits{});
        if (y > 0) {
            Bits up[16];
            load_half_rot<L>(tile, u - 2, 0, up);
            sub16(lo, up);
            load_half_rot<L>(tile, u - 2, 1, up);
            sub16(hi, up);
            if (g) left -= load_last_rot<L>(tile, u - 3, Bits{});
        }
    } else {
        // 16^3: run u = rows y = 2p, 2p+1 of plane z, with z = u>>3, p = u&7.
        // Row y-1 of the first row is the second half of run u-1; plane z-1 is 8 runs back.
        const int z = u >> 3, p = u & 7;
        Bits above[16];  // row 2p-1 of plane z (after the z difference)
#pragma unroll
        for (int i = 0; i < 16; ++i) above[i] = 0;
        if (p > 0) load_half_rot<L>(tile, u - 1, 1, above);
        if (z > 0) {
            Bits back[16];
            load_half_rot<L>(tile, u - 8, 0, back);
            sub16(lo, back);
            load_half_rot<L>(tile, u - 8, 1, back);
            sub16(hi, back);
            if (p > 0) {
                load_half_rot<L>(tile, u - 9, 1, back);
                sub16(above, back);
            }
        }
        // y difference: row 2p+1 -= row 2p, then row 2p -= row 2p-1
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            hi[i] -= lo[i];
            lo[i] -= above[i];
        }
    }

    // x difference, back to front so every step still sees the undifferenced predecessor
    if constexpr (Dims == 3) {
#pragma unroll
        for (int i = 15; i >= 1; --i) {
            hi[i] -= hi[i - 1];
            lo[i] -= lo[i - 1];
        }
    } else {
#pragma unroll
        for (int j = 31; j >= 1; --j) r[j] -= r[j - 1];
        r[0] -= left;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = complement_negative(r[j]);
}

// 3-D residuals in two halves around a warp shuffle (compress_ws_kernel, residual_run_3d_warp): residual_run loads five
// neighbouring half-runs per thread, but two of them (the row above in the own plane and one plane back) only serve to form
// "row 2p-1 after the z difference" — which the thread of run u-1 holds in the second half of ITS registers once it has
// taken its own z difference. First half: own run minus the run one plane back. Then the caller hands every thread the
// second half of run u-1 (__shfl_up_sync by one lane; runs with p = 0 — lanes 0, 8, 16, 24 — have no row above, every other
// lane's neighbour is in its own warp). Second half: y and x differences, complement. Bit-identical to residual_run.
// a -= b where cond != 0, as ONE predicated subtraction on the device (the compiler's `if (c) a -= b` is select + subtract)
NDZB_HD void sub_if(uint32_t &a, uint32_t b, int cond) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q sub.s32 %0, %0, %1;\n\t}" : "+r"(a) : "r"(b), "r"(cond));
#else
    if (cond) a -= b;
#endif
}
NDZB_HD void sub_if(uint64_t &a, uint64_t b, int cond) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q sub.s64 %0, %0, %1;\n\t}" : "+l"(a) : "l"(b), "r"(cond));
#else
    if (cond) a -= b;
#endif
}

template<typename Bits>
NDZB_HD void residual3_zdiff(const uint32_t *tile, int u, Bits *r) {
    using L = input_layout<Bits, 3>;
    Bits *lo = r, *hi = r + 16;
    load_half_rot<L>(tile, u, 0, lo);
    load_half_rot<L>(tile, u, 1, hi);
    const int z = u >> 3;
    if (z > 0) {
        Bits back[16];
        load_half_rot<L>(tile, u - 8, 0, back);
        sub16(lo, back);
        load_half_rot<L>(tile, u - 8, 1, back);
        sub16(hi, back);
    }
}

// `above` = r[16..32) of run u-1 after residual3_zdiff (ignored when p == 0)
template<typename Bits>
NDZB_HD void residual3_finish(int u, const Bits *above, Bits *r) {
    Bits *lo = r, *hi = r + 16;
    const int p = u & 7;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        hi[i] -= lo[i];
        sub_if(lo[i], above[i], p);
    }
#pragma unroll
    for (int i = 15; i >= 1; --i) {
        hi[i] -= hi[i - 1];
        lo[i] -= lo[i - 1];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = complement_negative(r[j]);
}

// ------------------------------------------------------------------------------------------------
// forward: bit planes of a run
//
// Plane i of a chunk (i = 0 is the MSB plane) has value j of the chunk at bit B-1-j
// (reference src/ndzip/cpu_codec.inl:355-363). transpose32 is LSB-indexed, so values and planes are
// fed / read in reversed register order, which is free with static register names.

// float: planes[i] = plane i of chunk u; returns head = OR of the 32 residuals
NDZB_HD uint32_t planes_of_run(const uint32_t *r, uint32_t *planes) {
    uint32_t head = 0;
    uint32_t a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        head |= r[j];
        a[31 - j] = r[j];
    }
    transpose32(a);
#pragma unroll
    for (int i = 0; i < 32; ++i) planes[i] = a[31 - i];
    return head;
}

// double: this thread's 32-bit half of planes 0..31 (from the high words) and of planes 32..63 (from
// the low words) of the chunk its run belongs to. The thread with the chunk's first 32 values owns
// bits 63..32 of every plane word, its partner bits 31..0 (cuda_codec.inl:241-264 has the same split).
// Returns the partial head (OR over this run's 32 values).
NDZB_HD uint64_t planes_of_run(const uint64_t *r, uint32_t *planes_hi, uint32_t *planes_lo) {
    uint64_t head = 0;
    uint32_t a[32], b[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        head |= r[j];
        a[31 - j] = static_cast<uint32_t>(r[j] >> 32);
        b[31 - j] = static_cast<uint32_t>(r[j]);
    }
    transpose32(a);
    transpose32(b);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        planes_hi[i] = a[31 - i];
        planes_lo[i] = b[31 - i];
    }
    return head;
}

// ------------------------------------------------------------------------------------------------
// forward: per-thread compaction of the planes into the cube's compressed image in shared memory
//
// The compressed cube is assembled in shared memory exactly as it appears in the stream
// ([heads][non-zero planes of chunk 0][chunk 1]...), in place over the dead input tile, and then
// copied out linearly (coalesced). Plane i of a chunk is emitted exactly when bit B-1-i of the chunk
// head is set (reference src/ndzip/cpu_codec.inl:514-538). `body` = word offset of the chunk's first
// plane inside the cube (C + exclusive plane count): a per-thread chain of predicated stores instead of a
// warp-per-chunk pass over all 4096 plane slots.

// On the device the chain is written in PTX: per plane ONE predicated STS and ONE predicated pointer bump that is a
// multiply-add with an opaque constant, so it issues on the fma pipe (the alu pipe bounds the kernel; see
// complement_negative). The compiler's own code for the C loop below is test + predicated store + predicated add +
// predicated move: three issue slots per plane, two of them alu. (A closed-form slot for heads whose set bits form
// one run was tried first: real heads are "sign plane + gap + low run" with a few holes on top in ~15 % of the
// chunks, so nearly every warp ran both paths — profiles/README.md round 2.)
#if defined(__CUDA_ARCH__)
// stores the planes whose head bits are set at *ptr (a shared-space byte address), advancing it. The bit tests stay in
// C so that ptxas extracts the predicates seven at a time (R2P); store and pointer bump are PTX.
__device__ __forceinline__ void emit_planes(uint32_t head, uint32_t &ptr, const uint32_t *planes, int stride_words) {
    const uint32_t neg_stride = static_cast<uint32_t>(-4 * stride_words);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if ((head >> (31 - i)) & 1u) {
            asm volatile("st.shared.u32 [%0], %1;\n\tmad.lo.u32 %0, %2, %3, %0;" : "+r"(ptr) : "r"(planes[i]), "r"(kFmaPipeConsts[0]), "r"(neg_stride) : "memory");
        }
    }
}
#endif

// float: image is an array of 32-bit words
NDZB_HD void compact_planes(uint32_t *image, int chunk, uint32_t head, uint32_t body, const uint32_t *planes) {
    image[chunk] = head;
    uint32_t *out = image + body;
#if defined(__CUDA_ARCH__)
    uint32_t ptr = static_cast<uint32_t>(__cvta_generic_to_shared(out));
    emit_planes(head, ptr, planes, 1);
#else
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if ((head >> (31 - i)) & 1u) *out++ = planes[i];
    }
#endif
}

// double: image is an array of 64-bit words seen as 32-bit halves (little endian: word w = halves
// 2w, 2w+1). The thread holding the chunk's first 32 values (`first`) owns the high halves.
NDZB_HD void compact_planes(uint32_t *image, int chunk, bool first, uint64_t head, uint32_t body,
        const uint32_t *planes_hi, const uint32_t *planes_lo) {
    const uint32_t head_hi = static_cast<uint32_t>(head >> 32), head_lo = static_cast<uint32_t>(head);
    image[2 * chunk + (first ? 1 : 0)] = first ? head_hi : head_lo;
    uint32_t *out = image + 2 * body + (first ? 1 : 0);
#if defined(__CUDA_ARCH__)
    uint32_t ptr = static_cast<uint32_t>(__cvta_generic_to_shared(out));
    emit_planes(head_hi, ptr, planes_hi, 2);
    emit_planes(head_lo, ptr, planes_lo, 2);
#else
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if ((head_hi >> (31 - i)) & 1u) {
            *out = planes_hi[i];
            out += 2;
        }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if ((head_lo >> (31 - i)) & 1u) {
            *out = planes_lo[i];
            out += 2;
        }
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// inverse: residuals of run u from the compressed image in shared memory (absent planes = 0; the
// transpose is an involution, reference src/test/codec_generic_test.cc:65-81), complement undone
// (common.hh:505-507)

#if defined(__CUDA_ARCH__)
// a[31 - i] = plane i of the chunk: loaded from *ptr (shared-space byte address, advanced) when the head bit is set,
// 0 otherwise. Same shape as emit_planes: one predicated LDS and one predicated fma-pipe pointer bump per plane.
__device__ __forceinline__ void fetch_planes(uint32_t head, uint32_t &ptr, uint32_t *a, int stride_words) {
    const uint32_t neg_stride = static_cast<uint32_t>(-4 * stride_words);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t v = 0;
        if ((head >> (31 - i)) & 1u) {
            asm volatile("ld.shared.u32 %1, [%0];\n\tmad.lo.u32 %0, %2, %3, %0;" : "+r"(ptr), "=r"(v) : "r"(kFmaPipeConsts[0]), "r"(neg_stride) : "memory");
        }
        a[31 - i] = v;
    }
}
#endif

NDZB_HD void run_of_image(const uint32_t *image, uint32_t head, uint32_t body, uint32_t *r) {
    uint32_t a[32];
    const uint32_t *in = image + body;
#if defined(__CUDA_ARCH__)
    uint32_t ptr = static_cast<uint32_t>(__cvta_generic_to_shared(in));
    fetch_planes(head, ptr, a, 1);
#else
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t v = 0;
        if ((head >> (31 - i)) & 1u) v = *in++;
        a[31 - i] = v;
    }
#endif
    transpose32(a);
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = complement_negative(a[31 - j]);
}

NDZB_HD void run_of_image(const uint32_t *image, bool first, uint64_t head, uint32_t body, uint64_t *r) {
    const uint32_t head_hi = static_cast<uint32_t>(head >> 32), head_lo = static_cast<uint32_t>(head);
    uint32_t a[32], b[32];
    const uint32_t *in = image + 2 * body + (first ? 1 : 0);
#if defined(__CUDA_ARCH__)
    uint32_t ptr = static_cast<uint32_t>(__cvta_generic_to_shared(in));
    fetch_planes(head_hi, ptr, a, 2);
    fetch_planes(head_lo, ptr, b, 2);
#else
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t v = 0;
        if ((head_hi >> (31 - i)) & 1u) {
            v = *in;
            in += 2;
        }
        a[31 - i] = v;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t v = 0;
        if ((head_lo >> (31 - i)) & 1u) {
            v = *in;
            in += 2;
        }
        b[31 - i] = v;
    }
#endif
    transpose32(a);
    transpose32(b);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        r[j] = complement_negative((static_cast<uint64_t>(a[31 - j]) << 32) | b[31 - j]);
    }
}

// Write run u back into the (swizzled) cube tile, e.g. after the in-register part of the inverse
// transform. No rotate: the tile keeps transform-domain values until the final store.
NDZB_HD void store_run(uint32_t *tile, int u, const uint32_t *r) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        st_quad(tile + tile_unit(u, g), quad{r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]});
    }
}
NDZB_HD void store_run(uint32_t *tile, int u, const uint64_t *r) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint64_t v0 = r[16 * h + 2 * k], v1 = r[16 * h + 2 * k + 1];
            st_quad(tile + (h << 12) + tile_unit(u, k),
                    quad{static_cast<uint32_t>(v0), static_cast<uint32_t>(v0 >> 32), static_cast<uint32_t>(v1),
                            static_cast<uint32_t>(v1 >> 32)});
        }
    }
}

template<typename Bits> NDZB_HD Bits tile_load(const uint32_t *tile, int e);
template<> NDZB_HD uint32_t tile_load<uint32_t>(const uint32_t *tile, int e) { return tile[tile_elem<uint32_t>(e)]; }
template<> NDZB_HD uint64_t tile_load<uint64_t>(const uint32_t *tile, int e) {
    const int w = tile_elem<uint64_t>(e);
    return (static_cast<uint64_t>(tile[w + 1]) << 32) | tile[w];
}
template<typename Bits> NDZB_HD void tile_store(uint32_t *tile, int e, Bits v);
template<> NDZB_HD void tile_store<uint32_t>(uint32_t *tile, int e, uint32_t v) { tile[tile_elem<uint32_t>(e)] = v; }
template<> NDZB_HD void tile_store<uint64_t>(uint32_t *tile, int e, uint64_t v) {
    const int w = tile_elem<uint64_t>(e);
    tile[w] = static_cast<uint32_t>(v);
    tile[w + 1] = static_cast<uint32_t>(v >> 32);
}

// ------------------------------------------------------------------------------------------------
// inverse: strip addresses of the column passes
//
// The y / z prefix-sum passes of the decoder walk two-element strips (2 consecutive x) down a column.
// tile_elem() of every strip costs ~5 integer instructions (the swizzle XOR depends on the row), which
// made address arithmetic ~12 % of the decoder. Along a column the XOR term only takes a few values
// that are known at compile time once the loop is unrolled, so the run-time part is folded into a few
// base registers and everything else into the load/store immediate:  address(k) = base[sel(k)] + imm(k).
// tests/host_sim checks every address against tile_elem().

// Value tile of the 3-D float decoder: tile_unit() with the z parity XORed into bit 2 of the unit. In the
// y pass a warp holds 4 z planes x 8 strips; plain tile_unit() puts the four planes on the same 16 banks
// (2 x the minimum number of wavefronts, profiles/r1_r2a_decompress.txt: 42 % of all shared-memory
// wavefronts were conflicts), with the parity term even and odd planes use complementary bank halves.
// (The double tile already spreads 8 strips of 16 bytes over all banks.)
NDZB_HD int tile3_unit(int row, int unit) { return (row << 5) | ((unit ^ (row & 7) ^ ((row >> 1) & 4)) << 2); }

template<typename Bits> NDZB_HD int tile3_elem(int e);
template<> NDZB_HD int tile3_elem<uint32_t>(int e) {
    const int run = e >> 5, j = e & 31;
    return tile3_unit(run, j >> 2) + (j & 3);
}
template<> NDZB_HD int tile3_elem<uint64_t>(int e) { return tile_elem<uint64_t>(e); }

// store_run for the 3-D decoder's value tile
NDZB_HD void store_run3(uint32_t *tile, int u, const uint32_t *r) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        st_quad(tile + tile3_unit(u, g), quad{r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]});
    }
}
NDZB_HD void store_run3(uint32_t *tile, int u, const uint64_t *r) { store_run(tile, u, r); }

// 3-D, y pass: strip (z = o, y = k, x = 2 xq)
template<typename Bits>
struct strip_addr_y3 {
    int base[8];
    NDZB_HD strip_addr_y3(int o, int xq) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if constexpr (sizeof(Bits) == 4) {
                // m = (unit bit 2 known at compile time) * 4 + (XOR constant of unit bits 0-1)
                base[m] = o * 256 + (((xq >> 1) ^ (m & 3)) << 2) + ((((m >> 2) ^ o) & 1) << 4) + ((xq & 1) << 1);
            } else {
                base[m] = o * 256 + ((xq ^ m) << 2);
            }
        }
    }
    NDZB_HD int at(int k) const {
        const int c = k >> 1;
        if constexpr (sizeof(Bits) == 4) return base[(((k & 1) ^ (c >> 2)) << 2) | (c & 3)] + c * 32;
        else return base[c] + (k & 1) * 4096 + c * 32;
    }
};

// 3-D, z pass: strip (z = k, y = o, x = 2 xq) — the swizzle only depends on the parity of z
template<typename Bits>
struct strip_addr_z3 {
    int base[2];
    NDZB_HD strip_addr_z3(int o, int xq) {
        base[0] = tile3_elem<Bits>(o * 16 + xq * 2);
        base[1] = tile3_elem<Bits>(256 + o * 16 + xq * 2) - 256;
    }
    NDZB_HD int at(int k) const { return base[k & 1] + k * 256; }
};

// 2-D, y pass: strip (y = 16 seg + k, x = 2 xq), xq < 32
template<typename Bits>
struct strip_addr_y2 {
    int base[4];
    NDZB_HD strip_addr_y2(int seg, int xq) {
        const int hb = xq >> 4;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            if constexpr (sizeof(Bits) == 4) {
                base[m] = seg * 1024 + hb * 32 + (((((xq & 15) >> 1) ^ hb) ^ (2 * m)) << 2) + ((xq & 1) << 1);
            } else {
                base[m] = ((xq >> 3) & 1) * 4096 + seg * 1024 + hb * 32 + ((((xq & 7) ^ hb) ^ (2 * m)) << 2);
            }
        }
    }
    NDZB_HD int at(int k) const { return base[k & 3] + k * 64; }
};

// ------------------------------------------------------------------------------------------------
// geometry shared by kernels and host code

// Exact unsigned division by a launch-invariant divisor without the ~20-instruction hardware
// sequence (Granlund-Montgomery round-up method): n / d = (t + ((n - t) >> sh1)) >> sh2, t = mulhi(mul, n).
struct fastdiv {
    uint32_t mul, sh1, sh2;
};

NDZB_HD fastdiv make_fastdiv(uint32_t d) {
    fastdiv f{1, 0, 0};
    if (d <= 1) return f;
    uint32_t log2_ceil = 0;
    while ((1ull << log2_ceil) < d) ++log2_ceil;
    f.mul = static_cast<uint32_t>((((1ull << log2_ceil) - d) << 32) / d + 1);
    f.sh1 = 1;
    f.sh2 = log2_ceil - 1;
    return f;
}

NDZB_HD uint32_t fast_divide(uint32_t n, const fastdiv &f) {
#if defined(__CUDA_ARCH__)
    const uint32_t t = __umulhi(f.mul, n);
#else
    const uint32_t t = static_cast<uint32_t>((static_cast<uint64_t>(f.mul) * n) >> 32);
#endif
    return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

struct grid_geom {
    // element extents, slowest first, padded with 1 in front so index 2 is always the fastest
    uint32_t n[3];
    uint32_t cubes[3];       // whole cubes per dimension (padded with 1)
    uint32_t num_cubes;
    fastdiv div_x, div_y;    // division by cubes[2] / cubes[1]
};

// hc -> cube coordinates (cube units). Cubes are numbered row-major over cube coordinates, slowest
// dimension first (reference src/ndzip/common.hh:414-433, 570-579).
template<int Dims>
NDZB_HD void cube_coords(const grid_geom &g, uint32_t hc, uint32_t &cz, uint32_t &cy, uint32_t &cx) {
    cz = cy = 0;
    if constexpr (Dims == 1) {
        cx = hc;
    } else if constexpr (Dims == 2) {
        cy = fast_divide(hc, g.div_x);
        cx = hc - cy * g.cubes[2];
    } else {
        const uint32_t t = fast_divide(hc, g.div_x);
        cx = hc - t * g.cubes[2];
        cz = fast_divide(t, g.div_y);
        cy = t - cz * g.cubes[1];
    }
}

// Linear element offset of the first element of hypercube `hc`.
template<int Dims>
NDZB_HD uint64_t cube_origin(const grid_geom &g, uint32_t hc) {
    constexpr uint32_t side = side_of<Dims>::value;
    uint32_t cz, cy, cx;
    cube_coords<Dims>(g, hc, cz, cy, cx);
    if constexpr (Dims == 1) {
        return static_cast<uint64_t>(cx) * side;
    } else if constexpr (Dims == 2) {
        return (static_cast<uint64_t>(cy) * side) * g.n[2] + static_cast<uint64_t>(cx) * side;
    } else {
        return ((static_cast<uint64_t>(cz) * side) * g.n[1] + static_cast<uint64_t>(cy) * side) * g.n[2]
                + static_cast<uint64_t>(cx) * side;
    }
}

// Linear element offset, relative to the cube origin, of cube-local element e
// (reference src/ndzip/gpu_common.hh:22-32).
template<int Dims>
NDZB_HD uint64_t cube_local_offset(const grid_geom &g, int e) {
    if constexpr (Dims == 1) {
        return static_cast<uint64_t>(e);
    } else if constexpr (Dims == 2) {
        return static_cast<uint64_t>(e >> 6) * g.n[2] + (e & 63);
    } else {
        return (static_cast<uint64_t>(e >> 8) * g.n[1] + ((e >> 4) & 15)) * g.n[2] + (e & 15);
    }
}

// Border = elements outside every whole cube, in ascending linear index
// (reference src/ndzip/common.hh:245-306, GPU counterpart gpu_common.hh:277-344).
// Closed form for the i-th border element's linear index.
struct border_geom {
    uint64_t n1, n2;         // extents of the two fastest dimensions (1 if absent)
    uint64_t in0, in1, in2;  // extents rounded down to whole cubes (slowest..fastest; 1-padded dims: n)
    uint64_t slab_border;    // border elements per slowest-dimension index below in0: n1*n2 - in1*in2
    uint64_t row_border;     // n2 - in2
    uint64_t count;          // total border elements
};

NDZB_HD uint64_t border_linear_index(const border_geom &b, uint64_t i) {
    const uint64_t head = b.in0 * b.slab_border;
    if (i >= head) return b.in0 * b.n1 * b.n2 + (i - head);  // trailing slabs are border entirely
    const uint64_t z = i / b.slab_border;
    const uint64_t i2 = i % b.slab_border;
    const uint64_t strip = b.in1 * b.row_border;  // x-tails of the rows that intersect cubes
    uint64_t y, x;
    if (i2 < strip) {
        y = i2 / b.row_border;
        x = b.in2 + i2 % b.row_border;
    } else {
        const uint64_t i3 = i2 - strip;
        y = b.in1 + i3 / b.n2;
        x = i3 % b.n2;
    }
    return (z * b.n1 + y) * b.n2 + x;
}

}  // namespace ndzb
