// Build: nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -o tma_store_probe tma_store_probe.cu -lcuda
// Probe: which forms of cp.async.bulk.tensor stores of 32-bit words to a 4-byte-aligned destination are legal on
// sm_100a? One case per process (CUDA errors are sticky).  usage: tma_store_probe <rank> <box> <x> <y> <rows_log2> [lanes]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int rank, int box, int x, int y, int lanes) {
    extern __shared__ __align__(1024) uint32_t tile[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tile[i] = 0xA0000000u + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < lanes) {
        const uint32_t *src = tile + threadIdx.x * box;
        const int cx = x + threadIdx.x * box;
        if (rank == 1) {
            asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.bulk_group [%0, {%2}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&map)),
                    "r"(smem_addr(src)), "r"(cx) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&map)),
                    "r"(smem_addr(src)), "r"(cx), "r"(y) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main(int argc, char **argv) {
    const int rank = atoi(argv[1]), box = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]), rows_log2 = atoi(argv[5]);
    const int lanes = argc > 6 ? atoi(argv[6]) : 1;
    const uint64_t row_words = 1ull << 20;
    const size_t words = 3 * row_words;
    uint32_t *d = nullptr;
    cudaMalloc(&d, words * 4);
    cudaMemset(d, 0, words * 4);
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map{};
    cuuint64_t gdim[2] = {rank == 1 ? words : row_words, 1ull << rows_log2};
    cuuint64_t gstride[1] = {row_words * 4};
    cuuint32_t bx[2] = {static_cast<cuuint32_t>(box), 1}, es[2] = {1, 1};
    CUresult r = reinterpret_cast<encode_fn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, d, gdim, gstride, bx, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("rank=%d box=%d x=%d y=%d rows=2^%d: ENCODE FAILED %d\n", rank, box, x, y, rows_log2, (int) r); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe<<<1, 64, 32768>>>(map, rank, box, x, y, lanes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rank=%d box=%d x=%d y=%d rows=2^%d lanes=%d: %s\n", rank, box, x, y, rows_log2, lanes, cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> h(words);
    cudaMemcpy(h.data(), d, words * 4, cudaMemcpyDeviceToHost);
    // expected: word (y*row_words + x + i) = 0xA0000000 + i for i < lanes*box, clipped to row y when rank == 2
    long bad = 0, written = 0;
    for (size_t w = 0; w < words; ++w) {
        const long long rel = static_cast<long long>(w) - (static_cast<long long>(y) * (rank == 2 ? (long long) row_words : 0) + x);
        bool inside = rel >= 0 && rel < static_cast<long long>(lanes) * box;
        if (rank == 2 && inside) inside = (w >> 20) == static_cast<size_t>(y);
        const uint32_t expect = inside ? 0xA0000000u + static_cast<uint32_t>(rel) : 0u;
        if (h[w] != expect) ++bad;
        if (h[w]) ++written;
    }
    printf("rank=%d box=%d x=%d y=%d rows=2^%d lanes=%d: ok, written=%ld mismatches=%ld\n", rank, box, x, y, rows_log2, lanes, written, bad);
    return bad != 0;
}
