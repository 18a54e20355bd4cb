// Micro-benchmark: issue rates of the integer instructions the codec leans on (per SM, B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/pipes scripts/ubench/pipes.cu
// Each kernel runs ITER iterations of 8 independent dependency chains per thread; 148*4 CTAs x 256 threads.
// Purpose: which ops share the alu pipe (LOP3/SHF/PRMT/IADD3), which go to the fma pipe (IMAD*), and whether a mix
// of the two issues at twice the rate of either alone — the encoder/decoder are alu-pipe bound (profiles/README.md).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

__device__ __forceinline__ uint32_t mulhi_u(uint32_t a, uint32_t b) {
    uint32_t t;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
    return t;
}
__device__ __forceinline__ uint32_t mulhi_s(uint32_t a, uint32_t b) {
    uint32_t t;
    asm volatile("mul.hi.s32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
    return t;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t t;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(b), "r"(c));
    return t;
}
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) {  // (a & m) | (b & ~m)
    uint32_t t;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(t) : "r"(a), "r"(b), "r"(m));
    return t;
}

template<int MODE>
__global__ void bench(uint32_t *out, uint32_t seed) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t o = r[(i + 1) & 7];
            if constexpr (MODE == 0) {          // SHF.R + LOP3 (alu + alu)
                r[i] = (r[i] >> 3) ^ o;
            } else if constexpr (MODE == 1) {   // IMAD.HI.U32 + LOP3
                r[i] = mulhi_u(r[i], 0x20000000u) ^ o;
            } else if constexpr (MODE == 2) {   // LOP3 + LOP3
                r[i] = (r[i] ^ 0x55555555u) & (o | 0x0f0f0f0fu);
            } else if constexpr (MODE == 3) {   // IMAD
                r[i] = r[i] * 5u + 7u;
            } else if constexpr (MODE == 4) {   // LOP3 + IMAD
                r[i] = (r[i] ^ 0x55555555u) * 5u + 7u;
            } else if constexpr (MODE == 5) {   // PRMT
                r[i] = __byte_perm(r[i], o, 0x5410);
            } else if constexpr (MODE == 6) {   // shift left (IMAD.SHL?) + LOP3
                r[i] = (r[i] << 4) ^ o;
            } else if constexpr (MODE == 7) {   // IMAD.HI.S32 (sign mask) + LOP3
                r[i] = mulhi_s(r[i], 1u) ^ o;
            } else if constexpr (MODE == 8) {   // rotate left by 1: SHF.L.W
                r[i] = __funnelshift_l(r[i], r[i], 1) ^ o;
            } else if constexpr (MODE == 9) {   // rotate left by 1 on the fma pipe: mul.hi + mad, + LOP3
                r[i] = mad_lo(r[i], 2u, mulhi_u(r[i], 2u)) ^ o;
            } else if constexpr (MODE == 10) {  // subtract as mad (x + o * -1) + LOP3
                r[i] = mad_lo(o, 0xffffffffu, r[i]) ^ 0x13579bdfu;
            } else if constexpr (MODE == 11) {  // plain subtract + LOP3 (what does ptxas pick?)
                r[i] = (r[i] - o) ^ 0x13579bdfu;
            } else if constexpr (MODE == 12) {  // butterfly step as written in C: 2 shifts + 4 logic
                const uint32_t x = r[i], y = o;
                r[i] = ((x & 0x0f0f0f0fu) | ((y << 4) & 0xf0f0f0f0u)) ^ (((x >> 4) & 0x0f0f0f0fu) | (y & 0xf0f0f0f0u));
            } else if constexpr (MODE == 13) {  // butterfly step with bit-select LOP3: 2 shifts + 2 logic (+1 xor)
                const uint32_t x = r[i], y = o;
                r[i] = bitsel(x, y << 4, 0x0f0f0f0fu) ^ bitsel(x >> 4, y, 0x0f0f0f0fu);
            } else if constexpr (MODE == 14) {  // same, right shift as mul.hi
                const uint32_t x = r[i], y = o;
                r[i] = bitsel(x, y << 4, 0x0f0f0f0fu) ^ bitsel(mulhi_u(x, 0x10000000u), y, 0x0f0f0f0fu);
            } else if constexpr (MODE == 15) {  // complement_negative: SHF.R.S32 + LOP3
                const uint32_t v = r[i] + o;
                r[i] = v ^ (static_cast<uint32_t>(static_cast<int32_t>(v) >> 31) & 0x7fffffffu);
            } else if constexpr (MODE == 16) {  // complement_negative with the sign mask from mul.hi.s32
                const uint32_t v = r[i] + o;
                r[i] = v ^ (mulhi_s(v, 1u) & 0x7fffffffu);
            } else if constexpr (MODE == 17) {  // POPC + add
                r[i] = __popc(r[i]) + o;
            } else if constexpr (MODE == 18) {  // IMAD.WIDE-style 64-bit mad (two results)
                const uint64_t w = static_cast<uint64_t>(r[i]) * 0x10000000ull;
                r[i] = static_cast<uint32_t>(w >> 32) ^ static_cast<uint32_t>(w) ^ o;
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template<int MODE>
void run(const char *name, uint32_t *d) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = 148 * 4, block = 256;
    bench<MODE><<<grid, block>>>(d, 1);
    cudaEventRecord(a);
    bench<MODE><<<grid, block>>>(d, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    // cycles per chain step per warp scheduler (4 per SM, 8 warps... 2 warps each), at 1.965 GHz
    const double steps_per_smsp = double(grid) / 148.0 * (block / 32) / 4.0 * ITER * 8.0;
    printf("%-52s %8.3f ms  %6.2f cycles per step per SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / steps_per_smsp);
}

int main() {
    uint32_t *d;
    cudaMalloc(&d, 148 * 4 * 256 * 4);
    run<0>("0  SHF.R + LOP3", d);
    run<1>("1  mul.hi.u32 + LOP3", d);
    run<2>("2  LOP3 + LOP3", d);
    run<3>("3  IMAD", d);
    run<4>("4  LOP3 + IMAD", d);
    run<5>("5  PRMT", d);
    run<6>("6  shl + LOP3", d);
    run<7>("7  mul.hi.s32(v,1) + LOP3", d);
    run<8>("8  SHF.L.W rotate + LOP3", d);
    run<9>("9  rotate as mul.hi+mad + LOP3", d);
    run<10>("10 sub as mad.lo(-1) + LOP3", d);
    run<11>("11 sub + LOP3", d);
    run<12>("12 butterfly step, C form (+xor)", d);
    run<13>("13 butterfly step, bit-select LOP3 (+xor)", d);
    run<14>("14 butterfly step, bit-select, shr as mul.hi (+xor)", d);
    run<15>("15 add + complement (SHF.R.S32 + LOP3)", d);
    run<16>("16 add + complement (mul.hi.s32 + LOP3)", d);
    run<17>("17 POPC + add", d);
    run<18>("18 mul.wide + 2 xor", d);
    return 0;
}
