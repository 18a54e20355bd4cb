// Minimal stand-in for boost::container::static_vector, written for this repo so that the
// reference's OpenMP codec (src/ndzip/cpu_codec.inl:717) compiles without Boost installed.
// Only the members that file uses are provided.
#pragma once
#include <array>
#include <cstddef>

namespace boost::container {

template<typename T, std::size_t Capacity>
class static_vector {
  public:
    void push_back(const T &v) { _items[_count++] = v; }
    void clear() { _count = 0; }
    std::size_t size() const { return _count; }
    bool empty() const { return _count == 0; }
    T &back() { return _items[_count - 1]; }
    const T &back() const { return _items[_count - 1]; }
    T &operator[](std::size_t i) { return _items[i]; }
    const T &operator[](std::size_t i) const { return _items[i]; }

  private:
    std::array<T, Capacity> _items{};
    std::size_t _count = 0;
};

}  // namespace boost::container
