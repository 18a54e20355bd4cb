// Micro-benchmark: issue rates of the integer instructions the codec leans on (per SM, B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes scripts/ubench/pipes.cu
// Each kernel runs ITER iterations of 8 independent dependency chains per thread; 148*4 CTAs x 256 threads.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template<int MODE>
__global__ void bench(uint32_t *out, uint32_t seed) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if constexpr (MODE == 0) {          // SHF.R (alu pipe)
                r[i] = (r[i] >> 3) ^ r[(i + 1) & 7];  // SHF + LOP3
            } else if constexpr (MODE == 1) {   // IMAD.HI (fma pipe?) + LOP3
                uint32_t t;
                asm volatile("mul.hi.u32 %0, %1, 0x20000000;" : "=r"(t) : "r"(r[i]));
                r[i] = t ^ r[(i + 1) & 7];
            } else if constexpr (MODE == 2) {   // LOP3 only
                r[i] = (r[i] ^ 0x55555555u) & (r[(i + 1) & 7] | 0x0f0f0f0fu);
            } else if constexpr (MODE == 3) {   // IMAD only (x*5+7)
                r[i] = r[i] * 5u + 7u;
            } else if constexpr (MODE == 4) {   // LOP3 + IMAD interleaved
                r[i] = (r[i] ^ 0x55555555u) * 5u + 7u;
            } else if constexpr (MODE == 5) {   // PRMT
                r[i] = __byte_perm(r[i], r[(i + 1) & 7], 0x5410) ;
            } else if constexpr (MODE == 6) {   // IMAD.SHL + LOP3
                r[i] = (r[i] << 4) ^ r[(i + 1) & 7];
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template<int MODE>
void run(const char *name, int ops_per_step, uint32_t *d) {
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
    const double warp_instr = double(grid) * (block / 32) * ITER * 8.0 * ops_per_step;
    // per SM per cycle at 1.965 GHz
    const double per_sm_clk = warp_instr / 148.0 / (ms * 1e-3 * 1.965e9);
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (%d ops per step)\n", name, ms, per_sm_clk, ops_per_step);
}

int main() {
    uint32_t *d;
    cudaMalloc(&d, 148 * 4 * 256 * 4);
    run<0>("SHF.R + LOP3", 2, d);
    run<1>("IMAD.HI + LOP3", 2, d);
    run<2>("LOP3 + LOP3", 2, d);
    run<3>("IMAD", 1, d);
    run<4>("LOP3 + IMAD", 2, d);
    run<5>("PRMT", 1, d);
    run<6>("SHL(IMAD.SHL?) + LOP3", 2, d);
    return 0;
}
