// ndzip/cuda.hh — device-pointer API, same declarations as the reference's include/ndzip/cuda.hh:10-41.
// Implemented by libndzip_b200.so (ndzip_b200/csrc/ndzip_adapter.cu) on top of the C ABI.
#pragma once

#include "ndzip.hh"

#include <cuda_runtime.h>

namespace ndzip {

// Asynchronous on the stream the object was created with; no internal synchronisation; may be
// called repeatedly; not thread-safe per object (scratch is per object), like the reference.
template<typename T>
class cuda_compressor {
  public:
    using value_type = T;
    using compressed_type = detail::bits_type<T>;

    virtual ~cuda_compressor() = default;

    // out_device_stream holds compressed_length_bound<T>(data_size) words;
    // out_device_stream_length may be nullptr.
    virtual void compress(const value_type *in_device_data, const extent &data_size, compressed_type *out_device_stream,
            index_type *out_device_stream_length)
            = 0;
};

template<typename T>
class cuda_decompressor {
  public:
    using value_type = T;
    using compressed_type = detail::bits_type<T>;

    virtual ~cuda_decompressor() = default;

    virtual void
    decompress(const compressed_type *in_device_stream, value_type *out_device_data, const extent &data_size)
            = 0;
};

template<typename T>
std::unique_ptr<cuda_compressor<T>>
make_cuda_compressor(const compressor_requirements &req, cudaStream_t stream = nullptr);

template<typename T>
std::unique_ptr<cuda_decompressor<T>> make_cuda_decompressor(dim_type dims, cudaStream_t stream = nullptr);

}  // namespace ndzip
