// ndzip/ndzip.hh — interface header of ndzip_b200, declaring the same public types and signatures as
// the reference's include/ndzip/ndzip.hh (celerity/ndzip @ ff4e6702) so that code written against
// the reference compiles and links against libndzip_b200.so unchanged. Written from the reference's
// interface, not copied: only the CUDA-facing subset is implemented by this library (see
// INTEGRATION.md); the data layout of `extent` and `compressor_requirements` is kept binary
// compatible (int + uint32[3]; int + uint32), which tests/cpp/ checks by building the same test
// program against the reference's own headers.
#pragma once

#include <cassert>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>

#if defined(__CUDACC__)
#define NDZIP_UNIVERSAL __host__ __device__
#else
#define NDZIP_UNIVERSAL
#endif

namespace ndzip {

using dim_type = int;          // reference ndzip.hh:19
using index_type = uint32_t;   // reference ndzip.hh:20: arrays have < 2^32 elements

inline constexpr dim_type max_dimensionality = 3;

class compressor_requirements;

namespace detail {
template<dim_type Dims>
class static_extent;
dim_type get_dimensionality(const compressor_requirements &req);
index_type get_num_hypercubes(const compressor_requirements &req);
}  // namespace detail

// 1-3 component size, slowest dimension first (reference ndzip.hh:35-160).
class extent {
  public:
    using const_iterator = const index_type *;
    using iterator = index_type *;

    constexpr extent() noexcept = default;

    NDZIP_UNIVERSAL constexpr explicit extent(dim_type dims) noexcept : _dims{dims} {}

    NDZIP_UNIVERSAL constexpr extent(std::initializer_list<index_type> components) noexcept
        : _dims{static_cast<dim_type>(components.size())} {
        dim_type d = 0;
        for (auto c : components) {
            if (d < max_dimensionality) _components[d++] = c;
        }
    }

    NDZIP_UNIVERSAL static constexpr extent broadcast(dim_type dims, index_type scalar) {
        extent e(dims);
        for (dim_type d = 0; d < dims; ++d) e._components[d] = scalar;
        return e;
    }

    NDZIP_UNIVERSAL constexpr dim_type dimensions() const { return _dims; }
    NDZIP_UNIVERSAL index_type &operator[](dim_type d) { return _components[d]; }
    NDZIP_UNIVERSAL index_type operator[](dim_type d) const { return _components[d]; }

    NDZIP_UNIVERSAL iterator begin() { return _components; }
    NDZIP_UNIVERSAL iterator end() { return _components + _dims; }
    NDZIP_UNIVERSAL const_iterator begin() const { return _components; }
    NDZIP_UNIVERSAL const_iterator end() const { return _components + _dims; }

#define NDZIP_EXTENT_OP(op)                                                                      \
    NDZIP_UNIVERSAL extent &operator op##=(const extent & other) {                               \
        for (dim_type d = 0; d < _dims; ++d) _components[d] op## = other._components[d];         \
        return *this;                                                                            \
    }                                                                                            \
    NDZIP_UNIVERSAL friend extent operator op(extent left, const extent &right) { return left op## = right; }
    NDZIP_EXTENT_OP(+)
    NDZIP_EXTENT_OP(-)
#undef NDZIP_EXTENT_OP
#define NDZIP_EXTENT_SCALAR_OP(op)                                                               \
    NDZIP_UNIVERSAL extent &operator op##=(index_type s) {                                       \
        for (dim_type d = 0; d < _dims; ++d) _components[d] op## = s;                            \
        return *this;                                                                            \
    }                                                                                            \
    NDZIP_UNIVERSAL friend extent operator op(extent left, index_type s) { return left op## = s; }
    NDZIP_EXTENT_SCALAR_OP(*)
    NDZIP_EXTENT_SCALAR_OP(/)
#undef NDZIP_EXTENT_SCALAR_OP
    NDZIP_UNIVERSAL friend extent operator*(index_type s, extent right) { return right *= s; }

    NDZIP_UNIVERSAL friend bool operator==(const extent &a, const extent &b) {
        bool same = a._dims == b._dims;
        for (dim_type d = 0; same && d < a._dims; ++d) same = a._components[d] == b._components[d];
        return same;
    }
    NDZIP_UNIVERSAL friend bool operator!=(const extent &a, const extent &b) { return !(a == b); }

  private:
    template<dim_type Dims>
    friend class detail::static_extent;

    dim_type _dims = 1;
    index_type _components[max_dimensionality] = {};
};

// reference ndzip.hh:163-170 — computed in index_type, like the reference
template<typename Extent>
NDZIP_UNIVERSAL index_type num_elements(const Extent &size) {
    index_type n = 1;
    for (dim_type d = 0; d < size.dimensions(); ++d) n *= size[d];
    return n;
}

// reference ndzip.hh:172-180 — row-major
template<typename Extent>
NDZIP_UNIVERSAL index_type linear_index(const Extent &space, const Extent &pos) {
    index_type l = pos[0];
    for (dim_type d = 1; d < space.dimensions(); ++d) l = l * space[d] + pos[d];
    return l;
}

namespace detail {
template<size_t Size> struct bits_type_s;
template<> struct bits_type_s<4> { using type = uint32_t; };
template<> struct bits_type_s<8> { using type = uint64_t; };
template<typename T>
using bits_type = typename bits_type_s<sizeof(T)>::type;  // reference ndzip.hh:186-212
}  // namespace detail

template<typename T>
using compressed_type = detail::bits_type<T>;

// Upper bound of the stream length in words (reference ndzip.hh:224-225, src/ndzip/common.cc:31-55).
template<typename T>
index_type compressed_length_bound(const extent &e);

// Host-pointer codec interfaces (reference ndzip.hh:227-253). This library provides no CPU codec:
// make_compressor / make_decompressor are intentionally not defined (the reference does not export
// them either).
template<typename T>
class compressor {
  public:
    using value_type = T;
    using compressed_type = detail::bits_type<T>;
    virtual ~compressor() = default;
    virtual index_type compress(const value_type *data, const extent &data_size, compressed_type *stream) = 0;
};

template<typename T>
class decompressor {
  public:
    using value_type = T;
    using compressed_type = detail::bits_type<T>;
    virtual ~decompressor() = default;
    virtual index_type decompress(const compressed_type *stream, value_type *data, const extent &data_size) = 0;
};

// Maximum hypercube count over the extents a compressor will be used with
// (reference ndzip.hh:255-269, src/ndzip/common.cc:8-28).
class compressor_requirements {
  public:
    compressor_requirements() = default;
    compressor_requirements(const ndzip::extent &single_data_size);  // NOLINT(google-explicit-constructor)
    compressor_requirements(std::initializer_list<extent> data_sizes);

    void include(const extent &data_size);

  private:
    friend dim_type detail::get_dimensionality(const compressor_requirements &);
    friend index_type detail::get_num_hypercubes(const compressor_requirements &);

    dim_type _dims = -1;
    index_type _max_num_hypercubes = 0;
};

using kernel_duration = std::chrono::duration<uint64_t, std::nano>;

}  // namespace ndzip
