// Drop-in check of the C++ boundary: this program is written against the REFERENCE's public API
// (ndzip::make_cuda_offloader / make_cuda_compressor / make_cuda_decompressor / compressed_length_bound /
// compressor_requirements) and is compiled twice by tests/test_cpp_adapter.py:
//   1. against this repo's include/ndzip/*.hh,
//   2. against the reference's own include/ndzip/*.hh (where /root/reference exists),
// and linked against libndzip_b200.so both times. Streams are compared with the CPU oracle
// (oracle/libndzip_oracle.so — test infrastructure), mirroring the reference's cross-encoder
// tests (src/test/codec_profile_test.inl:37-140, 952-1082).
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#include <ndzip/cuda.hh>
#include <ndzip/offload.hh>

extern "C" uint32_t ndzo_compress(int dtype, int dims, const uint32_t *size, const void *data, void *stream);

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
            return 1;                                                                \
        }                                                                            \
    } while (0)

#define CUDA_OK(call)                                                                \
    do {                                                                             \
        const cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                                     \
            std::fprintf(stderr, "CUDA %s: %s\n", #call, cudaGetErrorString(e_));    \
            return 1;                                                                \
        }                                                                            \
    } while (0)

template<typename T>
static std::vector<T> make_field(const ndzip::extent &e) {
    const size_t n = ndzip::num_elements(e);
    std::vector<T> v(n);
    for (size_t i = 0; i < n; ++i) {
        const double x = static_cast<double>(i);
        v[i] = static_cast<T>(std::sin(x * 1e-3) + 0.25 * std::cos(x * 7e-2) + 1e-5 * static_cast<double>((i * 2654435761u) % 1000));
    }
    return v;
}

template<typename T>
static int run_profile(ndzip::dim_type dims, ndzip::index_type n) {
    using bits = ndzip::compressed_type<T>;
    ndzip::extent e(dims);
    for (ndzip::dim_type d = 0; d < dims; ++d) e[d] = n + static_cast<ndzip::index_type>(d);
    const auto data = make_field<T>(e);
    const auto bound = ndzip::compressed_length_bound<T>(e);

    // oracle stream
    uint32_t size[3] = {0, 0, 0};
    for (ndzip::dim_type d = 0; d < dims; ++d) size[d] = e[d];
    std::vector<bits> expect(bound + 1, 0);
    const uint32_t expect_len = ndzo_compress(sizeof(T) == 4 ? 0 : 1, dims, size, data.data(), expect.data());

    // host-pointer API (offload.hh)
    auto off = ndzip::make_offloader<T>(ndzip::target::cuda, dims);
    std::vector<bits> stream(bound + 1, bits{0x5a});
    ndzip::kernel_duration dur{};
    const auto len = off->compress(data.data(), e, stream.data(), &dur);
    CHECK(len == expect_len);
    CHECK(std::memcmp(stream.data(), expect.data(), size_t{len} * sizeof(bits)) == 0);
    CHECK(dur.count() > 0);
    std::vector<T> back(data.size());
    const auto consumed = off->decompress(stream.data(), len, back.data(), e);
    CHECK(consumed == len);
    CHECK(std::memcmp(back.data(), data.data(), data.size() * sizeof(T)) == 0);

    // device-pointer API (cuda.hh), explicit stream, requirements covering a larger extent too
    cudaStream_t cs;
    CUDA_OK(cudaStreamCreate(&cs));
    ndzip::compressor_requirements req(e);
    auto comp = ndzip::make_cuda_compressor<T>(req, cs);
    auto dec = ndzip::make_cuda_decompressor<T>(dims, cs);
    T *d_in = nullptr, *d_back = nullptr;
    bits *d_stream = nullptr;
    ndzip::index_type *d_len = nullptr;
    CUDA_OK(cudaMalloc(&d_in, data.size() * sizeof(T) + 16));
    CUDA_OK(cudaMalloc(&d_back, data.size() * sizeof(T) + 16));
    CUDA_OK(cudaMalloc(&d_stream, (size_t{bound} + 1) * sizeof(bits)));
    CUDA_OK(cudaMalloc(&d_len, sizeof(ndzip::index_type)));
    CUDA_OK(cudaMemcpyAsync(d_in, data.data(), data.size() * sizeof(T), cudaMemcpyHostToDevice, cs));
    for (int rep = 0; rep < 2; ++rep) {  // objects are reusable
        comp->compress(d_in, e, d_stream, d_len);
        dec->decompress(d_stream, d_back, e);
    }
    ndzip::index_type dev_len = 0;
    std::vector<bits> dev_stream(bound + 1);
    CUDA_OK(cudaMemcpyAsync(&dev_len, d_len, sizeof dev_len, cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaMemcpyAsync(dev_stream.data(), d_stream, size_t{bound} * sizeof(bits), cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaMemcpyAsync(back.data(), d_back, data.size() * sizeof(T), cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaStreamSynchronize(cs));
    CHECK(dev_len == expect_len);
    CHECK(std::memcmp(dev_stream.data(), expect.data(), size_t{dev_len} * sizeof(bits)) == 0);
    CHECK(std::memcmp(back.data(), data.data(), data.size() * sizeof(T)) == 0);

    // nullptr length pointer is allowed (cuda.hh:18-21)
    comp->compress(d_in, e, d_stream, nullptr);
    CUDA_OK(cudaStreamSynchronize(cs));

    // dimensionality mismatch -> std::runtime_error (cuda_codec.inl:557-559)
    bool threw = false;
    try {
        ndzip::extent wrong(dims == 3 ? 2 : dims + 1);
        for (ndzip::dim_type d = 0; d < wrong.dimensions(); ++d) wrong[d] = 16;
        comp->compress(d_in, wrong, d_stream, d_len);
    } catch (const std::runtime_error &) { threw = true; }
    CHECK(threw);

    cudaFree(d_in);
    cudaFree(d_back);
    cudaFree(d_stream);
    cudaFree(d_len);
    cudaStreamDestroy(cs);
    std::printf("ok %s %dD n=%u: %u words (bound %u)\n", sizeof(T) == 4 ? "float" : "double", dims, n, len, bound);
    return 0;
}

int main() {
    int rc = 0;
    // 4*side-1 style bordered extents, as in the reference tests
    rc |= run_profile<float>(1, 4096 * 3 + 17);
    rc |= run_profile<float>(2, 255);
    rc |= run_profile<float>(3, 63);
    rc |= run_profile<double>(1, 4096 * 2 + 5);
    rc |= run_profile<double>(2, 191);
    rc |= run_profile<double>(3, 47);
    // empty requirements -> std::runtime_error (common.hh:319-322)
    bool threw = false;
    try {
        ndzip::compressor_requirements empty;
        auto c = ndzip::make_cuda_compressor<float>(empty, nullptr);
    } catch (const std::runtime_error &) { threw = true; }
    if (!threw) {
        std::fprintf(stderr, "FAIL: empty requirements did not throw\n");
        rc = 1;
    }
    std::puts(rc == 0 ? "PASS" : "FAILED");
    return rc;
}
