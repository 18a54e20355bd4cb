// Minimal stand-in for boost::lockfree::queue<T, capacity<N>> (mutex-protected ring), written
// for this repo so the reference's OpenMP codec (src/ndzip/cpu_codec.inl:734) compiles without
// Boost. Semantics used there: push / pop(T&) -> bool / consume_all(fn).
#pragma once
#include <cstddef>
#include <deque>
#include <mutex>

namespace boost::lockfree {

template<std::size_t N>
struct capacity {
    static constexpr std::size_t value = N;
};

template<typename T, typename Capacity>
class queue {
  public:
    bool push(const T &v) {
        std::lock_guard<std::mutex> lock(_mutex);
        if (_items.size() >= Capacity::value) return false;
        _items.push_back(v);
        return true;
    }

    bool pop(T &out) {
        std::lock_guard<std::mutex> lock(_mutex);
        if (_items.empty()) return false;
        out = _items.front();
        _items.pop_front();
        return true;
    }

    template<typename Fn>
    std::size_t consume_all(Fn &&fn) {
        std::lock_guard<std::mutex> lock(_mutex);
        std::size_t n = _items.size();
        for (auto &v : _items) fn(v);
        _items.clear();
        return n;
    }

  private:
    std::mutex _mutex;
    std::deque<T> _items;
};

}  // namespace boost::lockfree
