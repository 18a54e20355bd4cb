// Micro-benchmark: throughput of TMA tensor loads / stores for the tile shapes the codec could use on a 512^3 f32
// (or 256x512x512 f64) grid. One CTA per SM owns a ring of tiles; a producer thread issues the copies, a consumer
// thread hands the slot straight back. Nothing is computed: this is the ceiling the load (compress) or store
// (decompress) side of a kernel can reach with that shape.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/tma_shapes scripts/ubench/tma_shapes.cu -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_addr(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load(int rank, void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
    if (rank == 2) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_addr(dst)), "l"(m), "r"(smem_addr(bar)), "r"(c0), "r"(c1) : "memory");
    else if (rank == 3) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_addr(dst)), "l"(m), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_addr(dst)), "l"(m), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store(int rank, const CUtensorMap *m, const void *src, int c0, int c1, int c2, int c3) {
    if (rank == 2) asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_addr(src)), "r"(c0), "r"(c1) : "memory");
    else if (rank == 3) asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct shape {
    int rank;
    int tiles[3];       // tiles along x, y, z (tile index is decomposed x fastest)
    int step[4];        // coordinate step per tile index for each tensor dimension (0 for the parity dimension)
    int map_dim[4];     // which of (tx, ty, tz) feeds tensor dimension d (-1: the per-load index, e.g. y parity)
    int loads;          // copies per tile
    int load_bytes;     // bytes per copy
};

template<bool Store>
__global__ void __launch_bounds__(64, 1) run(const __grid_constant__ CUtensorMap map, shape sh, int depth, int total_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16], empty[16];
    const int tile_bytes = sh.loads * sh.load_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < depth; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto coords = [&](int t, int l, int *c) {
        const int tx = t % sh.tiles[0], ty = (t / sh.tiles[0]) % sh.tiles[1], tz = t / (sh.tiles[0] * sh.tiles[1]);
        const int idx[3] = {tx, ty, tz};
        for (int d = 0; d < 4; ++d) c[d] = sh.map_dim[d] < 0 ? l : idx[sh.map_dim[d]] * sh.step[d];
    };
    if (threadIdx.x == 0) {
        int s = 0, parity = 1, round0 = 1;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            if constexpr (!Store) {
                if (!round0) mbar_wait(&empty[s], parity);
                mbar_expect(&full[s], tile_bytes);
                for (int l = 0; l < sh.loads; ++l) {
                    int c[4];
                    coords(t, l, c);
                    tma_load(sh.rank, smem + s * tile_bytes + l * sh.load_bytes, &map, &full[s], c[0], c[1], c[2], c[3]);
                }
            } else {
                for (int l = 0; l < sh.loads; ++l) {
                    int c[4];
                    coords(t, l, c);
                    tma_store(sh.rank, &map, smem + s * tile_bytes + l * sh.load_bytes, c[0], c[1], c[2], c[3]);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // keep at most depth-1 tiles' stores in flight (their shared-memory source still being read)
                switch (depth - 1) {
                    case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
                    case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
                    case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
                    case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
                    case 6: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
                    default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
                }
            }
            if (++s == depth) { s = 0; parity ^= 1; round0 = 0; }
        }
        if constexpr (Store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (threadIdx.x == 32 && !Store) {
        int s = 0, parity = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            mbar_wait(&full[s], parity);
            mbar_arrive(&empty[s]);
            if (++s == depth) { s = 0; parity ^= 1; }
        }
    }
}

using encode_fn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    auto encode = reinterpret_cast<encode_fn>(fp);
    const size_t bytes = size_t{512} << 20;
    void *d;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 1, bytes));
    CK(cudaFuncSetAttribute(run<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(run<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    struct cfg { const char *name; CUtensorMapDataType type; int rank; cuuint64_t gdim[4]; cuuint64_t gstride[3]; cuuint32_t box[4]; CUtensorMapSwizzle sw; shape sh; CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B; };
    const cuuint64_t N = 512;
    cfg cfgs[] = {
        {"f32 3D cube, view (x, y/2, z, par) box 16x8x16x2 SW64, 1 copy (parity outermost)", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, {N, N / 2, N, 2}, {N * 8, N * N * 4, N * 4}, {16, 8, 16, 2}, CU_TENSOR_MAP_SWIZZLE_64B,
                {4, {32, 32, 32}, {16, 8, 16, 0}, {0, 1, 2, 0}, 1, 16384}},
        {"f32 3D cube, [z][y/2][par][x] box 16x1x8x16 SW64, 2 copies, L2 promotion 128B", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, {N, 2, N / 2, N}, {N * 4, N * 8, N * N * 4}, {16, 1, 8, 16}, CU_TENSOR_MAP_SWIZZLE_64B,
                {4, {32, 32, 32}, {16, 0, 8, 16}, {0, -1, 1, 2}, 2, 8192}, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"f32 3D cube, [z][y/2][par][x] box 16x1x8x16 SW64, 2 copies, L2 promotion none", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, {N, 2, N / 2, N}, {N * 4, N * 8, N * N * 4}, {16, 1, 8, 16}, CU_TENSOR_MAP_SWIZZLE_64B,
                {4, {32, 32, 32}, {16, 0, 8, 16}, {0, -1, 1, 2}, 2, 8192}, CU_TENSOR_MAP_L2_PROMOTION_NONE},
        {"f32 3D cube, view (x, z, y/2, par)->[par][y/2][z][x] box 16x16x8x2 SW64, 1 copy", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, {N, N, N / 2, 2}, {N * N * 4, N * 8, N * 4}, {16, 16, 8, 2}, CU_TENSOR_MAP_SWIZZLE_64B,
                {4, {32, 32, 32}, {16, 16, 8, 0}, {0, 2, 1, 0}, 1, 16384}},
        {"f32 3D cube, [z][y/2][par][x] box 16x1x8x16 SW64, 2 copies (current)", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, {N, 2, N / 2, N}, {N * 4, N * 8, N * N * 4}, {16, 1, 8, 16}, CU_TENSOR_MAP_SWIZZLE_64B,
                {4, {32, 32, 32}, {16, 0, 8, 16}, {0, -1, 1, 2}, 2, 8192}},
        {"f32 3D cube pair, [z][y/2][par][x] box 32x1x8x16 SW128, 2 copies", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, {N, 2, N / 2, N}, {N * 4, N * 8, N * N * 4}, {32, 1, 8, 16}, CU_TENSOR_MAP_SWIZZLE_128B,
                {4, {16, 32, 32}, {32, 0, 8, 16}, {0, -1, 1, 2}, 2, 16384}},
        {"f32 3D cube, [z][y][x] box 16x16x16 SW64, 1 copy", CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, {N, N, N, 1}, {N * 4, N * N * 4, 0}, {16, 16, 16, 1}, CU_TENSOR_MAP_SWIZZLE_64B,
                {3, {32, 32, 32}, {16, 16, 16, 0}, {0, 1, 2, 0}, 1, 16384}},
        {"f32 3D cube pair, [z][y][x] box 32x16x16 SW128, 1 copy", CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, {N, N, N, 1}, {N * 4, N * N * 4, 0}, {32, 16, 16, 1}, CU_TENSOR_MAP_SWIZZLE_128B,
                {3, {16, 32, 32}, {32, 16, 16, 0}, {0, 1, 2, 0}, 1, 32768}},
        {"f32 3D cube, [z][y][x] box 16x16x16 no swizzle, 1 copy", CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, {N, N, N, 1}, {N * 4, N * N * 4, 0}, {16, 16, 16, 1}, CU_TENSOR_MAP_SWIZZLE_NONE,
                {3, {32, 32, 32}, {16, 16, 16, 0}, {0, 1, 2, 0}, 1, 16384}},
        {"f32 3D cube quad, [z][y][x] box 64x16x16 no swizzle, 1 copy", CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, {N, N, N, 1}, {N * 4, N * N * 4, 0}, {64, 16, 16, 1}, CU_TENSOR_MAP_SWIZZLE_NONE,
                {3, {8, 32, 32}, {64, 16, 16, 0}, {0, 1, 2, 0}, 1, 65536}},
        {"f32 1D cube, [rows][32] box 32x128 SW128, 1 copy (contiguous 16 KiB)", CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, {32, N * N * N / 32, 1, 1}, {128, 0, 0}, {32, 128, 1, 1}, CU_TENSOR_MAP_SWIZZLE_128B,
                {2, {32768, 1, 1}, {0, 128, 0, 0}, {1, 0, 1, 1}, 1, 16384}},
        {"f32 2D cube, [y][x/32][32] box 32x2x64 SW128 (8192x16384), 1 copy", CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, {32, 16384 / 32, 8192, 1}, {128, 16384 * 4, 0}, {32, 2, 64, 1}, CU_TENSOR_MAP_SWIZZLE_128B,
                {3, {256, 128, 1}, {0, 2, 64, 0}, {2, 0, 1, 2}, 1, 16384}},
        {"f64 3D cube (256x512x512), [z][y/2][par][x] box 16x1x8x16 SW128, 2 copies (current)", CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, {N, 2, N / 2, N / 2}, {N * 8, N * 16, N * N * 8}, {16, 1, 8, 16}, CU_TENSOR_MAP_SWIZZLE_128B,
                {4, {32, 32, 16}, {16, 0, 8, 16}, {0, -1, 1, 2}, 2, 16384}},
        {"f64 3D cube (256x512x512), [z][y][x] box 16x16x16 SW128, 1 copy", CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, {N, N, N / 2, 1}, {N * 8, N * N * 8, 0}, {16, 16, 16, 1}, CU_TENSOR_MAP_SWIZZLE_128B,
                {3, {32, 32, 16}, {16, 16, 16, 0}, {0, 1, 2, 0}, 1, 32768}},
    };
    for (auto &c : cfgs) {
        CUtensorMap map;
        cuuint32_t es[4] = {1, 1, 1, 1};
        const CUresult r = encode(&map, c.type, c.rank, d, c.gdim, c.gstride, c.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-84s encode failed (%d)\n", c.name, int(r)); continue; }
        const int tile_bytes = c.sh.loads * c.sh.load_bytes;
        const int total = c.sh.tiles[0] * c.sh.tiles[1] * c.sh.tiles[2];
        for (int store = 0; store < 2; ++store) {
            for (int kb : {64, 128, 192}) {
                int depth = kb * 1024 / tile_bytes;
                if (depth > 8) depth = 8;
                if (depth < 2) continue;
                if (store && depth == 5) depth = 4;
                float best = 1e9f;
                for (int rep = 0; rep < 5; ++rep) {
                    cudaEvent_t a, b;
                    cudaEventCreate(&a); cudaEventCreate(&b);
                    cudaEventRecord(a);
                    if (store) run<true><<<148, 64, depth * tile_bytes>>>(map, c.sh, depth, total);
                    else run<false><<<148, 64, depth * tile_bytes>>>(map, c.sh, depth, total);
                    cudaEventRecord(b);
                    CK(cudaEventSynchronize(b));
                    float ms; cudaEventElapsedTime(&ms, a, b);
                    if (rep > 0 && ms < best) best = ms;
                }
                printf("%-84s %s depth %d (%3d KiB in flight): %.4f ms  %6.0f GB/s\n", c.name, store ? "STORE" : "LOAD ", depth, depth * tile_bytes / 1024, best, bytes / (best * 1e-3) / 1e9);
            }
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
