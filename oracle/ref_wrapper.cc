// TEST INFRASTRUCTURE — not part of the product path.
//
// C-ABI wrapper around the UNMODIFIED reference CPU codec, compiled from the sources where they
// lie under /root/reference (see oracle/Makefile, target `ref`). Nothing from the reference is
// copied into this repository: this file only #includes the reference headers by path and
// forwards to their public entry points.
//
//   ndzip::make_cpu_offloader<T>(dims, threads)      reference include/ndzip/offload.hh:54-55
//   ndzip::compressed_length_bound<T>(extent)        reference src/ndzip/common.cc:44-55
//   detail::block_transform / inverse_block_transform reference src/ndzip/common.hh:469-535
//   cpu::transpose_bits_trivial / zero_bit_encode    reference src/ndzip/cpu_codec.inl:355-363, 541-559
//   detail::for_each_border_slice                    reference src/ndzip/common.hh:268-282
//
// The resulting oracle/_ref/libndzip_ref.so is used (a) to pin oracle/ndzip_oracle.c,
// (b) as the bit-exactness checker in tests/, and (c) as bench.py's CPU baseline
// (cpu_baseline.kind = "reference").

// The reference's public headers are not self-contained under GCC 13 (SURVEY.md §8b caveats).
#include <array>
#include <cassert>
#include <cstring>
#include <initializer_list>
#include <iterator>
#include <stdexcept>
#include <string>

#include <ndzip/cpu_codec.inl>  // from -I/root/reference/src ; pulls common.hh + ndzip.hh + offload.hh

#include <cstdint>
#include <memory>
#include <vector>

namespace {

using namespace ndzip;

extent make_extent(int dims, const uint32_t *size) {
    extent e(dims);
    for (int d = 0; d < dims; ++d) e[d] = size[d];
    return e;
}

template<typename T>
struct offloader_box {
    std::unique_ptr<offloader<T>> impl;
};

struct handle {
    int dtype;
    int dims;
    offloader_box<float> f32;
    offloader_box<double> f64;
};

template<typename Bits>
void scalar_forward(Bits *x, int dims) {
    const index_type side = dims == 1 ? 4096 : dims == 2 ? 64 : 16;
    detail::block_transform(x, dims, side);
}

template<typename Bits>
void scalar_inverse(Bits *x, int dims) {
    const index_type side = dims == 1 ? 4096 : dims == 2 ? 64 : 16;
    detail::inverse_block_transform(x, dims, side);
}

template<typename T, dim_type D>
void simd_forward(void *x) {
    using P = detail::profile<T, D>;
    detail::cpu::simd_aligned_buffer<typename P::bits_type> buf(4096);
    memcpy(buf.data(), x, 4096 * sizeof(T));
    detail::cpu::block_transform<P>(buf.data());
    memcpy(x, buf.data(), 4096 * sizeof(T));
}

template<typename T, dim_type D>
void simd_inverse(void *x) {
    using P = detail::profile<T, D>;
    detail::cpu::simd_aligned_buffer<typename P::bits_type> buf(4096);
    memcpy(buf.data(), x, 4096 * sizeof(T));
    detail::cpu::inverse_block_transform<P>(buf.data());
    memcpy(x, buf.data(), 4096 * sizeof(T));
}

template<dim_type D>
int border_slices(const uint32_t *size, uint32_t side, uint32_t *out_pairs, int max_pairs) {
    detail::static_extent<D> e;
    for (dim_type d = 0; d < D; ++d) e[d] = size[d];
    int n = 0;
    detail::for_each_border_slice(e, side, [&](index_type offset, index_type count) {
        if (n < max_pairs) {
            out_pairs[2 * n] = offset;
            out_pairs[2 * n + 1] = count;
        }
        ++n;
    });
    return n;
}

}  // namespace

extern "C" {

// threads: 1 = serial_compressor (`-e cpu -T 1`), 0 = all physical cores, n = n OpenMP threads.
void *ndzr_offloader_create(int dtype, int dims, unsigned threads) {
    try {
        auto h = std::make_unique<handle>();
        h->dtype = dtype;
        h->dims = dims;
        if (dtype == 0) {
            h->f32.impl = make_cpu_offloader<float>(dims, threads);
        } else {
            h->f64.impl = make_cpu_offloader<double>(dims, threads);
        }
        return h.release();
    } catch (...) { return nullptr; }
}

void ndzr_offloader_destroy(void *hp) {
    delete static_cast<handle *>(hp);
}

// Returns the stream length in bits_type words (offload.hh:16-19). The caller zero-fills `stream`
// (the CPU encoder never writes the f64 odd-H header padding word).
uint32_t ndzr_offloader_compress(void *hp, const uint32_t *size, const void *data, void *stream) {
    auto *h = static_cast<handle *>(hp);
    const auto e = make_extent(h->dims, size);
    if (h->dtype == 0) {
        return h->f32.impl->compress(static_cast<const float *>(data), e, static_cast<uint32_t *>(stream));
    }
    return h->f64.impl->compress(static_cast<const double *>(data), e, static_cast<uint64_t *>(stream));
}

// Returns the number of stream words consumed (offload.hh:21-24).
uint32_t ndzr_offloader_decompress(void *hp, const uint32_t *size, const void *stream, uint32_t length, void *data) {
    auto *h = static_cast<handle *>(hp);
    const auto e = make_extent(h->dims, size);
    if (h->dtype == 0) {
        return h->f32.impl->decompress(static_cast<const uint32_t *>(stream), length, static_cast<float *>(data), e);
    }
    return h->f64.impl->decompress(static_cast<const uint64_t *>(stream), length, static_cast<double *>(data), e);
}

uint32_t ndzr_compressed_length_bound(int dtype, int dims, const uint32_t *size) {
    const auto e = make_extent(dims, size);
    return dtype == 0 ? compressed_length_bound<float>(e) : compressed_length_bound<double>(e);
}

uint32_t ndzr_num_hypercubes(int dims, const uint32_t *size) {
    return detail::num_hypercubes(make_extent(dims, size));
}

// Scalar, normative transform on one 4096-element cube of bits_type (common.hh:469-535).
void ndzr_block_transform(int dtype, int dims, void *cube) {
    if (dtype == 0) scalar_forward(static_cast<uint32_t *>(cube), dims);
    else scalar_forward(static_cast<uint64_t *>(cube), dims);
}

void ndzr_inverse_block_transform(int dtype, int dims, void *cube) {
    if (dtype == 0) scalar_inverse(static_cast<uint32_t *>(cube), dims);
    else scalar_inverse(static_cast<uint64_t *>(cube), dims);
}

// The transform the CPU encoder actually runs (AVX2 when compiled in; cpu_codec.inl:325-341).
void ndzr_block_transform_simd(int dtype, int dims, void *cube) {
    if (dtype == 0) {
        if (dims == 1) simd_forward<float, 1>(cube);
        else if (dims == 2) simd_forward<float, 2>(cube);
        else simd_forward<float, 3>(cube);
    } else {
        if (dims == 1) simd_forward<double, 1>(cube);
        else if (dims == 2) simd_forward<double, 2>(cube);
        else simd_forward<double, 3>(cube);
    }
}

void ndzr_inverse_block_transform_simd(int dtype, int dims, void *cube) {
    if (dtype == 0) {
        if (dims == 1) simd_inverse<float, 1>(cube);
        else if (dims == 2) simd_inverse<float, 2>(cube);
        else simd_inverse<float, 3>(cube);
    } else {
        if (dims == 1) simd_inverse<double, 1>(cube);
        else if (dims == 2) simd_inverse<double, 2>(cube);
        else simd_inverse<double, 3>(cube);
    }
}

// B x B bit-matrix transpose, scalar definition (cpu_codec.inl:355-363).
void ndzr_transpose_bits(int dtype, const void *in, void *out) {
    if (dtype == 0) {
        detail::cpu::transpose_bits_trivial(static_cast<const uint32_t *>(in), static_cast<uint32_t *>(out));
    } else {
        detail::cpu::transpose_bits_trivial(static_cast<const uint64_t *>(in), static_cast<uint64_t *>(out));
    }
}

// Residual cube (4096 bits_type) -> compressed cube; returns bytes written (cpu_codec.inl:541-559).
uint64_t ndzr_zero_bit_encode(int dtype, const void *cube, void *stream) {
    if (dtype == 0) {
        detail::cpu::simd_aligned_buffer<uint32_t> buf(4096);
        memcpy(buf.data(), cube, 4096 * 4);
        return detail::cpu::zero_bit_encode<uint32_t>(buf.data(), static_cast<std::byte *>(stream), 4096);
    }
    detail::cpu::simd_aligned_buffer<uint64_t> buf(4096);
    memcpy(buf.data(), cube, 4096 * 8);
    return detail::cpu::zero_bit_encode<uint64_t>(buf.data(), static_cast<std::byte *>(stream), 4096);
}

// Compressed cube -> residual cube; returns bytes consumed (cpu_codec.inl:561-578).
uint64_t ndzr_zero_bit_decode(int dtype, const void *stream, void *cube) {
    if (dtype == 0) {
        detail::cpu::simd_aligned_buffer<uint32_t> buf(4096);
        auto n = detail::cpu::zero_bit_decode<uint32_t>(static_cast<const std::byte *>(stream), buf.data(), 4096);
        memcpy(cube, buf.data(), 4096 * 4);
        return n;
    }
    detail::cpu::simd_aligned_buffer<uint64_t> buf(4096);
    auto n = detail::cpu::zero_bit_decode<uint64_t>(static_cast<const std::byte *>(stream), buf.data(), 4096);
    memcpy(cube, buf.data(), 4096 * 8);
    return n;
}

// (offset, count) border slices in emission order (common.hh:245-282). Returns the slice count.
int ndzr_border_slices(int dims, const uint32_t *size, uint32_t side, uint32_t *out_pairs, int max_pairs) {
    switch (dims) {
        case 1: return border_slices<1>(size, side, out_pairs, max_pairs);
        case 2: return border_slices<2>(size, side, out_pairs, max_pairs);
        default: return border_slices<3>(size, side, out_pairs, max_pairs);
    }
}

int ndzr_has_openmp(void) {
#if NDZIP_OPENMP_SUPPORT
    return 1;
#else
    return 0;
#endif
}

unsigned ndzr_physical_concurrency(void) {
#if NDZIP_OPENMP_SUPPORT
    return boost::thread::physical_concurrency();
#else
    return 1;
#endif
}

}  // extern "C"
