// Multi-GPU data plane through the C ABI alone, in ONE process (no torchrun, no Python): ndzb_dist_create_local
// (ncclCommInitAll), one host thread per GPU, slab compression + count exchange + gather, and the gathered stream
// compared word for word with (a) the stream one GPU produces for the whole grid and (b) the CPU oracle.
// usage: dist_test [world]   (default: all visible GPUs, at least 1)
#include <ndzip_b200.h>

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" uint32_t ndzo_compress(int dtype, int dims, const uint32_t *size, const void *data, void *stream);  // oracle/ndzip_oracle.c

#define CHECK(x)                                                                                    \
    do {                                                                                            \
        const int s_ = (x);                                                                         \
        if (s_ != 0) {                                                                              \
            printf("FAIL %s:%d %s -> %d (%s | %s)\n", __FILE__, __LINE__, #x, s_, ndzb_strerror(s_), ndzb_dist_last_error()); \
            exit(1);                                                                                \
        }                                                                                           \
    } while (0)
#define CUDA(x)                                                                       \
    do {                                                                              \
        const cudaError_t e_ = (x);                                                   \
        if (e_ != cudaSuccess) {                                                      \
            printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

template<typename T>
std::vector<T> make_grid(int dims, const uint32_t *size) {
    uint64_t n = 1;
    for (int d = 0; d < dims; ++d) n *= size[d];
    std::vector<T> v(n);
    for (uint64_t i = 0; i < n; ++i) {
        v[i] = static_cast<T>(std::sin(0.001 * static_cast<double>(i)) + 1e-3 * static_cast<double>((i * 2654435761u) % 1000) / 1000.0);
    }
    return v;
}

template<typename T>
bool run_case(int world, const int *devices, int dims, std::vector<uint32_t> size) {
    const int dtype = sizeof(T) == 4 ? NDZB_F32 : NDZB_F64;
    using word = typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type;
    size.resize(3, 0);
    const std::vector<T> grid = make_grid<T>(dims, size.data());
    uint64_t row_elems = 1;
    for (int d = 1; d < dims; ++d) row_elems *= size[d];

    std::vector<ndzb_dist *> ranks(world);
    CHECK(ndzb_dist_create_local(ranks.data(), dtype, dims, size.data(), world, devices));
    std::vector<word> gathered;
    uint64_t gathered_words = 0;
    std::vector<int> round_trip_ok(world, 0);

    auto body = [&](int r) {
        CUDA(cudaSetDevice(devices[r]));
        ndzb_dist *d = ranks[r];
        ndzb_dist_layout L;
        CHECK(ndzb_dist_layout_of(d, r, &L));
        cudaStream_t stream = static_cast<cudaStream_t>(ndzb_dist_stream(d));
        const uint64_t slab_elems = static_cast<uint64_t>(L.slab_end - L.slab_begin) * row_elems;
        T *d_slab = nullptr, *d_back = nullptr;
        word *d_local = nullptr, *d_global = nullptr;
        CUDA(cudaMalloc(&d_slab, slab_elems * sizeof(T) + 16));
        CUDA(cudaMalloc(&d_back, slab_elems * sizeof(T) + 16));
        CUDA(cudaMalloc(&d_local, L.local_bound_words * sizeof(word) + 16));
        if (r == 0) CUDA(cudaMalloc(&d_global, L.global_bound_words * sizeof(word) + 16));
        CUDA(cudaMemcpyAsync(d_slab, grid.data() + static_cast<uint64_t>(L.slab_begin) * row_elems, slab_elems * sizeof(T), cudaMemcpyHostToDevice, stream));
        CHECK(ndzb_dist_compress(d, d_slab, d_local, nullptr));
        CHECK(ndzb_dist_decompress(d, d_local, d_back));
        uint64_t total = 0;
        CHECK(ndzb_dist_gather(d, d_local, d_global, 0, &total));
        std::vector<T> back(slab_elems);
        CUDA(cudaMemcpyAsync(back.data(), d_back, slab_elems * sizeof(T), cudaMemcpyDeviceToHost, stream));
        if (r == 0) {
            gathered.resize(total);
            gathered_words = total;
            CUDA(cudaMemcpyAsync(gathered.data(), d_global, total * sizeof(word), cudaMemcpyDeviceToHost, stream));
        }
        CUDA(cudaStreamSynchronize(stream));
        round_trip_ok[r] = slab_elems == 0 || memcmp(back.data(), grid.data() + static_cast<uint64_t>(L.slab_begin) * row_elems, slab_elems * sizeof(T)) == 0;
        cudaFree(d_slab);
        cudaFree(d_back);
        cudaFree(d_local);
        cudaFree(d_global);
    };
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r) threads.emplace_back(body, r);
    for (auto &t : threads) t.join();
    for (auto *d : ranks) ndzb_dist_destroy(d);

    // the oracle's stream of the whole grid
    std::vector<word> expect(ndzb_compressed_length_bound(dtype, dims, size.data()) + 1, 0);
    const uint64_t n = ndzo_compress(dtype, dims, size.data(), grid.data(), expect.data());
    bool ok = n == gathered_words && memcmp(expect.data(), gathered.data(), n * sizeof(word)) == 0;
    for (int r = 0; r < world; ++r) ok = ok && round_trip_ok[r];
    printf("%s %dD %u x %u x %u on %d GPU(s): gathered %llu words, oracle %llu words: %s\n", sizeof(T) == 4 ? "f32" : "f64", dims, size[0], size[1],
            size[2], world, static_cast<unsigned long long>(gathered_words), static_cast<unsigned long long>(n), ok ? "ok" : "MISMATCH");
    return ok;
}

int main(int argc, char **argv) {
    int ndev = 0;
    CUDA(cudaGetDeviceCount(&ndev));
    int world = argc > 1 ? atoi(argv[1]) : ndev;
    if (world < 1 || world > ndev) {
        printf("FAIL: %d GPU(s) visible, world %d\n", ndev, world);
        return 1;
    }
    std::vector<int> devices(world);
    for (int r = 0; r < world; ++r) devices[r] = r;
    bool ok = true;
    ok = run_case<float>(world, devices.data(), 3, {100, 70, 50}) && ok;
    ok = run_case<double>(world, devices.data(), 2, {64 * 5 + 9, 200}) && ok;
    ok = run_case<float>(world, devices.data(), 1, {4096 * 11 + 3}) && ok;
    ok = run_case<double>(world, devices.data(), 3, {16 * 3, 16, 16}) && ok;  // 3 cubes: odd count, header padding word
    ok = run_case<float>(world, devices.data(), 3, {10, 40, 40}) && ok;      // no cubes at all: borders only
    printf(ok ? "PASS\n" : "FAIL\n");
    return ok ? 0 : 1;
}
