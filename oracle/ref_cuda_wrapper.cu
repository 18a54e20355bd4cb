// TEST / BENCH INFRASTRUCTURE — C-ABI wrapper around the UNMODIFIED reference CUDA codec
// (reference src/ndzip/cuda_codec.inl, cuda_bits.cuh, cuda_factory.cu), recompiled for sm_100a so that
// bench.py can time "the reference's own kernels on the same B200" next to ours, and tests can use it
// as a second, GPU-side oracle. Built by `make -C oracle refcuda` from a scratch copy of the reference
// tree in which only src/ndzip/cuda_workaround.hh is emptied (its `#undef __noinline__` breaks
// libstdc++ 13 under CUDA 12.9). No reference source is copied into this repository.
#include <array>
#include <cassert>
#include <initializer_list>
#include <iterator>
#include <stdexcept>
#include <string>

#include <ndzip/cuda_codec.inl>  // scratch copy: -I<scratch>/src

namespace ndzip::detail::gpu_cuda {
template class cuda_offloader<profile<float, 1>>;
template class cuda_offloader<profile<float, 2>>;
template class cuda_offloader<profile<float, 3>>;
template class cuda_offloader<profile<double, 1>>;
template class cuda_offloader<profile<double, 2>>;
template class cuda_offloader<profile<double, 3>>;
}  // namespace ndzip::detail::gpu_cuda

namespace {
ndzip::extent make_extent(int dims, const uint32_t *size) {
    ndzip::extent e(dims);
    for (int d = 0; d < dims; ++d) e[d] = size[d];
    return e;
}
struct handle {
    int dtype, dims;
    std::unique_ptr<ndzip::cuda_compressor<float>> cf;
    std::unique_ptr<ndzip::cuda_compressor<double>> cd;
    std::unique_ptr<ndzip::cuda_decompressor<float>> df;
    std::unique_ptr<ndzip::cuda_decompressor<double>> dd;
};
}  // namespace

extern "C" {

// device-pointer API of the reference (include/ndzip/cuda.hh:10-41); stream = cudaStream_t
void *ndzrc_create(int dtype, int dims, const uint32_t *size, void *stream) {
    try {
        auto h = std::make_unique<handle>();
        h->dtype = dtype;
        h->dims = dims;
        const ndzip::compressor_requirements req(make_extent(dims, size));
        if (dtype == 0) {
            h->cf = ndzip::make_cuda_compressor<float>(req, static_cast<cudaStream_t>(stream));
            h->df = ndzip::make_cuda_decompressor<float>(dims, static_cast<cudaStream_t>(stream));
        } else {
            h->cd = ndzip::make_cuda_compressor<double>(req, static_cast<cudaStream_t>(stream));
            h->dd = ndzip::make_cuda_decompressor<double>(dims, static_cast<cudaStream_t>(stream));
        }
        return h.release();
    } catch (...) { return nullptr; }
}

void ndzrc_destroy(void *hp) { delete static_cast<handle *>(hp); }

int ndzrc_compress(void *hp, const void *d_in, const uint32_t *size, void *d_stream, uint32_t *d_len) {
    auto *h = static_cast<handle *>(hp);
    try {
        const auto e = make_extent(h->dims, size);
        if (h->dtype == 0) h->cf->compress(static_cast<const float *>(d_in), e, static_cast<uint32_t *>(d_stream), d_len);
        else h->cd->compress(static_cast<const double *>(d_in), e, static_cast<uint64_t *>(d_stream), d_len);
        return 0;
    } catch (...) { return -1; }
}

int ndzrc_decompress(void *hp, const void *d_stream, void *d_out, const uint32_t *size) {
    auto *h = static_cast<handle *>(hp);
    try {
        const auto e = make_extent(h->dims, size);
        if (h->dtype == 0) h->df->decompress(static_cast<const uint32_t *>(d_stream), static_cast<float *>(d_out), e);
        else h->dd->decompress(static_cast<const uint64_t *>(d_stream), static_cast<double *>(d_out), e);
        return 0;
    } catch (...) { return -1; }
}

}  // extern "C"
