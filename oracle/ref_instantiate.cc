// TEST INFRASTRUCTURE — explicit instantiations that replace the reference's CMake
// "SplitConfiguration" generated translation units (reference CMakeLists.txt:105-121,
// src/ndzip/cpu_codec.inl:661-678, 929-943). Compiled against the reference sources in place.
#include <array>
#include <cassert>
#include <initializer_list>
#include <iterator>
#include <stdexcept>
#include <string>

#include <ndzip/cpu_codec.inl>

namespace ndzip::detail::cpu {

#define NDZR_INSTANTIATE(T, D)                                   \
    template class serial_compressor<profile<T, D>>;             \
    template class serial_decompressor<profile<T, D>>;
NDZR_INSTANTIATE(float, 1)
NDZR_INSTANTIATE(float, 2)
NDZR_INSTANTIATE(float, 3)
NDZR_INSTANTIATE(double, 1)
NDZR_INSTANTIATE(double, 2)
NDZR_INSTANTIATE(double, 3)

#if NDZIP_OPENMP_SUPPORT
#define NDZR_INSTANTIATE_MT(T, D)                                \
    template class openmp_compressor<profile<T, D>>;             \
    template class openmp_decompressor<profile<T, D>>;
NDZR_INSTANTIATE_MT(float, 1)
NDZR_INSTANTIATE_MT(float, 2)
NDZR_INSTANTIATE_MT(float, 3)
NDZR_INSTANTIATE_MT(double, 1)
NDZR_INSTANTIATE_MT(double, 2)
NDZR_INSTANTIATE_MT(double, 3)
#endif

}  // namespace ndzip::detail::cpu
