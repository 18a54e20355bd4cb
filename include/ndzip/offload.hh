// ndzip/offload.hh — host-pointer API, same declarations as the reference's include/ndzip/offload.hh:8-71
// for the CUDA target. Synchronous; returns stream lengths in words.
#pragma once

#include "ndzip.hh"

#ifndef NDZIP_CUDA_SUPPORT
#define NDZIP_CUDA_SUPPORT 1
#endif

namespace ndzip {

template<typename T>
class offloader {
  public:
    using value_type = T;
    using compressed_type = detail::bits_type<T>;

    virtual ~offloader() = default;

    // returns the stream length in words; `duration` (optional) receives the kernel-only time
    index_type compress(const value_type *data, const extent &data_size, compressed_type *stream,
            kernel_duration *duration = nullptr) {
        return do_compress(data, data_size, stream, duration);
    }

    // returns the number of stream words consumed
    index_type decompress(const compressed_type *stream, index_type length, value_type *data, const extent &data_size,
            kernel_duration *duration = nullptr) {
        return do_decompress(stream, length, data, data_size, duration);
    }

  protected:
    virtual index_type
    do_compress(const value_type *data, const extent &data_size, compressed_type *stream, kernel_duration *duration)
            = 0;

    virtual index_type do_decompress(const compressed_type *stream, index_type length, value_type *data,
            const extent &data_size, kernel_duration *duration)
            = 0;
};

// Only the CUDA target exists in this library (BASELINE.json north_star: no multi-backend dispatch,
// no CPU fallback). `cpu` is kept so that the enumerator values match the reference's
// (offload.hh:41-49 with NDZIP_CUDA_SUPPORT=1, NDZIP_HIPSYCL_SUPPORT=0); asking for it throws.
enum class target {
    cpu,
    cuda,
};

template<typename T>
std::unique_ptr<offloader<T>> make_cuda_offloader(dim_type dimensions);

template<typename T>
std::unique_ptr<offloader<T>> make_offloader(target target, dim_type dimensions, bool enable_profiling = false) {
    (void) enable_profiling;
    switch (target) {
        case target::cuda: return make_cuda_offloader<T>(dimensions);
        default: throw std::runtime_error("ndzip::make_offloader: invalid target");
    }
}

}  // namespace ndzip
