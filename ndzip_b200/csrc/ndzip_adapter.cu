// ndzip_adapter.cu — the reference's C++ interface for this path, implemented on the C ABI.
//
//   ndzip::make_cuda_compressor<T>    replaces src/ndzip/cuda_factory.cu:4-9   + cuda_compressor_impl   (cuda_codec.inl:514-603)
//   ndzip::make_cuda_decompressor<T>  replaces src/ndzip/cuda_factory.cu:11-14 + cuda_decompressor_impl (cuda_codec.inl:605-652)
//   ndzip::make_cuda_offloader<T>     replaces src/ndzip/cuda_factory.cu:16-19 + cuda_offloader         (cuda_codec.inl:654-761)
//   ndzip::compressed_length_bound<T>, compressor_requirements       replace src/ndzip/common.cc:8-55
// Errors become std::runtime_error, as in the reference (cuda_bits.cuh:165-169, cuda_codec.inl:557-559).
#include "../../include/ndzip/cuda.hh"
#include "../../include/ndzip/offload.hh"
#include "../../include/ndzip_b200.h"

#include <algorithm>

namespace ndzip {

namespace {

[[noreturn]] void throw_status(int status) {
    std::string msg = ndzb_strerror(status);
    if (status == NDZB_ERR_CUDA) msg += std::string(": ") + ndzb_last_cuda_error();
    throw std::runtime_error{msg};
}
void check(int status) {
    if (status != NDZB_OK) throw_status(status);
}

template<typename T>
constexpr int dtype_of = sizeof(T) == 4 ? NDZB_F32 : NDZB_F64;

struct size_array {
    uint32_t v[3] = {0, 0, 0};
    explicit size_array(const extent &e) {
        for (dim_type d = 0; d < e.dimensions() && d < 3; ++d) v[d] = e[d];
    }
};

class context {
  public:
    context(int dtype, dim_type dims, index_type max_hypercubes, cudaStream_t stream) {
        if (dims < 1 || dims > max_dimensionality) throw std::runtime_error{"Invalid dimensionality"};  // common.hh:642
        check(ndzb_ctx_create(&_ctx, dtype, dims, max_hypercubes, stream));
    }
    context(const context &) = delete;
    context &operator=(const context &) = delete;
    ~context() { ndzb_ctx_destroy(_ctx); }
    ndzb_ctx *get() const { return _ctx; }

  private:
    ndzb_ctx *_ctx = nullptr;
};

template<typename T>
class b200_compressor final : public cuda_compressor<T> {
  public:
    using compressed_type = detail::bits_type<T>;
    b200_compressor(const compressor_requirements &req, cudaStream_t stream)
        : _ctx(dtype_of<T>, detail::get_dimensionality(req), detail::get_num_hypercubes(req), stream) {}

    void compress(const T *in_device_data, const extent &data_size, compressed_type *out_device_stream,
            index_type *out_device_stream_length) override {
        const size_array size(data_size);
        check(ndzb_compress(_ctx.get(), in_device_data, data_size.dimensions(), size.v, out_device_stream,
                out_device_stream_length));
    }

  private:
    context _ctx;
};

template<typename T>
class b200_decompressor final : public cuda_decompressor<T> {
  public:
    using compressed_type = detail::bits_type<T>;
    b200_decompressor(dim_type dims, cudaStream_t stream) : _ctx(dtype_of<T>, dims, 0, stream) {}

    void decompress(const compressed_type *in_device_stream, T *out_device_data, const extent &data_size) override {
        const size_array size(data_size);
        check(ndzb_decompress(_ctx.get(), in_device_stream, out_device_data, data_size.dimensions(), size.v));
    }

  private:
    context _ctx;
};

template<typename T>
class b200_offloader final : public offloader<T> {
  public:
    using compressed_type = detail::bits_type<T>;
    explicit b200_offloader(dim_type dims) : _ctx(dtype_of<T>, dims, 0, nullptr) {}

  protected:
    index_type do_compress(const T *data, const extent &data_size, compressed_type *stream, kernel_duration *duration) override {
        const size_array size(data_size);
        uint32_t length = 0;
        uint64_t ns = 0;
        check(ndzb_offload_compress(_ctx.get(), data, data_size.dimensions(), size.v, stream, &length, &ns));
        if (duration) *duration = kernel_duration{ns};
        return length;
    }

    index_type do_decompress(const compressed_type *stream, index_type length, T *data, const extent &data_size,
            kernel_duration *duration) override {
        const size_array size(data_size);
        uint32_t consumed = 0;
        uint64_t ns = 0;
        check(ndzb_offload_decompress(_ctx.get(), stream, length, data, data_size.dimensions(), size.v, &consumed, &ns));
        if (duration) *duration = kernel_duration{ns};
        return consumed;
    }

  private:
    context _ctx;
};

}  // namespace

// ---- compressor_requirements (reference src/ndzip/common.cc:8-28) ----------------------------------

namespace detail {
dim_type get_dimensionality(const compressor_requirements &req) {
    if (req._dims == -1) throw std::runtime_error{"Cannot construct a compressor with empty requirements"};
    return req._dims;
}
index_type get_num_hypercubes(const compressor_requirements &req) {
    return req._max_num_hypercubes;
}
}  // namespace detail

void compressor_requirements::include(const extent &data_size) {
    if (_dims == -1) {
        _dims = data_size.dimensions();
    } else if (data_size.dimensions() != _dims) {
        throw std::runtime_error{"Cannot add a " + std::to_string(data_size.dimensions()) + "-dimensional extent to "
                + std::to_string(_dims) + "-dimensional compressor_requirements"};
    }
    const size_array size(data_size);
    _max_num_hypercubes = std::max(_max_num_hypercubes, ndzb_num_hypercubes(data_size.dimensions(), size.v));
}

compressor_requirements::compressor_requirements(const extent &single_data_size) {
    include(single_data_size);
}

compressor_requirements::compressor_requirements(std::initializer_list<extent> data_sizes) {
    for (const auto &e : data_sizes) include(e);
}

// ---- compressed_length_bound (reference src/ndzip/common.cc:31-55) ---------------------------------

template<typename T>
index_type compressed_length_bound(const extent &size) {
    const size_array s(size);
    return static_cast<index_type>(ndzb_compressed_length_bound(dtype_of<T>, size.dimensions(), s.v));
}
template index_type compressed_length_bound<float>(const extent &);
template index_type compressed_length_bound<double>(const extent &);

// ---- factories (reference src/ndzip/cuda_factory.cu:4-32) ------------------------------------------

template<typename T>
std::unique_ptr<cuda_compressor<T>> make_cuda_compressor(const compressor_requirements &req, cudaStream_t stream) {
    return std::make_unique<b200_compressor<T>>(req, stream);
}
template<typename T>
std::unique_ptr<cuda_decompressor<T>> make_cuda_decompressor(dim_type dims, cudaStream_t stream) {
    return std::make_unique<b200_decompressor<T>>(dims, stream);
}
template<typename T>
std::unique_ptr<offloader<T>> make_cuda_offloader(dim_type dimensions) {
    return std::make_unique<b200_offloader<T>>(dimensions);
}

template std::unique_ptr<cuda_compressor<float>> make_cuda_compressor<float>(const compressor_requirements &, cudaStream_t);
template std::unique_ptr<cuda_compressor<double>> make_cuda_compressor<double>(const compressor_requirements &, cudaStream_t);
template std::unique_ptr<cuda_decompressor<float>> make_cuda_decompressor<float>(dim_type, cudaStream_t);
template std::unique_ptr<cuda_decompressor<double>> make_cuda_decompressor<double>(dim_type, cudaStream_t);
template std::unique_ptr<offloader<float>> make_cuda_offloader<float>(dim_type);
template std::unique_ptr<offloader<double>> make_cuda_offloader<double>(dim_type);

}  // namespace ndzip
