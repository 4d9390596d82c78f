// Host-side mirror of the allocator plumbing every reference signature carries
// (reference: include/flucoma/data/FluidMemory.hpp:11-56).  The reference backs these with foonathan/memory so that
// real-time clients never call malloc on the audio thread; the GPU path is offline/batched, so this mirror keeps the
// names (`Allocator`, `rt::vector`, `FluidDefaultAllocator()`) on top of the standard allocator.
#pragma once
#include <memory>
#include <vector>

namespace fluid {

class Allocator
{
public:
  void* allocate(std::size_t bytes) { return ::operator new(bytes); }
  void  deallocate(void* p) noexcept { ::operator delete(p); }
};

inline Allocator& FluidDefaultAllocator()
{
  static Allocator a;
  return a;
}

namespace rt {
template <typename T>
class vector : public std::vector<T>
{
public:
  using std::vector<T>::vector;
  vector(std::size_t n, Allocator&) : std::vector<T>(n) {}
  vector(std::size_t n, const T& v, Allocator&) : std::vector<T>(n, v) {}
  explicit vector(Allocator&) {}
};
} // namespace rt
} // namespace fluid
