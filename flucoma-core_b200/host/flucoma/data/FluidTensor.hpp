// Host-side containers with the public surface of flucoma-core's FluidTensor / FluidTensorView / Slice
// (reference API: include/flucoma/data/FluidTensor.hpp:65-741, FluidTensor_Support.hpp:32-41,259-482; behaviour pinned by
// tests/data/TestFluidTensor.cpp, TestFluidTensorView.cpp and the death tests in tests/data/death_tests/).
//
// Written from scratch for this repo: no foonathan/memory, no Eigen.  A view is {extents, strides, pointer}; an owning
// tensor is a std::vector plus a dense row-major descriptor.  Everything the GPU shims need from a view is
// `data()`, `extent(i)` and `descriptor().strides[i]` -- the C ABI takes dense arrays, so shims pack strided views first.
#pragma once
#include "FluidIndex.hpp"
#include "FluidMemory.hpp"
#include <algorithm>
#include <array>
#include <cassert>
#include <initializer_list>
#include <ostream>
#include <type_traits>
#include <vector>

namespace fluid {

// Slice{start, length, stride}; length < 0 means "to the end" (FluidTensor_Support.hpp:32-41)
struct Slice
{
  Slice() : start(-1), length(-1), stride(1) {}
  explicit Slice(index s) : start(s), length(-1), stride(1) {}
  Slice(index s, index l, index n = 1) : start(s), length(l), stride(n) {}
  index start, length, stride;
};

template <size_t N>
struct FluidTensorSlice
{
  static constexpr size_t order = N;
  index                size{0};
  index                start{0};
  std::array<index, N> extents{};
  std::array<index, N> strides{};

  FluidTensorSlice() = default;

  template <typename... Dims, typename = std::enable_if_t<sizeof...(Dims) == N>>
  FluidTensorSlice(index s, Dims... dims) : start(s), extents{{static_cast<index>(dims)...}}
  {
    init();
  }
  FluidTensorSlice(index s, std::array<index, N> e) : start(s), extents(e) { init(); }
  FluidTensorSlice(index s, std::array<index, N> e, std::array<index, N> st) : start(s), extents(e), strides(st)
  {
    size = 1;
    for (auto x : extents) size *= x;
  }

  void init()
  { // dense row-major
    index acc = 1;
    for (size_t i = N; i-- > 0;)
    {
      strides[i] = acc;
      acc *= extents[i];
    }
    size = acc;
  }

  template <typename... Idx>
  index operator()(Idx... idx) const
  {
    static_assert(sizeof...(Idx) == N, "wrong number of indices");
    std::array<index, N> a{{static_cast<index>(idx)...}};
    index                off = start;
    for (size_t i = 0; i < N; ++i)
    {
      assert(a[i] >= 0 && a[i] < extents[i] && "FluidTensor: index out of range");
      off += a[i] * strides[i];
    }
    return off;
  }

  FluidTensorSlice transpose() const
  {
    FluidTensorSlice r(*this);
    std::reverse(r.extents.begin(), r.extents.end());
    std::reverse(r.strides.begin(), r.strides.end());
    return r;
  }

  bool operator==(const FluidTensorSlice& o) const
  {
    return start == o.start && extents == o.extents && strides == o.strides;
  }
};

template <typename T, size_t N>
class FluidTensor;
template <typename T, size_t N>
class FluidTensorView;

namespace impl {
// nested initializer lists
template <typename T, size_t N>
struct InitList
{
  using type = std::initializer_list<typename InitList<T, N - 1>::type>;
};
template <typename T>
struct InitList<T, 1>
{
  using type = std::initializer_list<T>;
};

template <typename T, typename L>
void flatten(const L& l, std::vector<T>& out, std::vector<index>& ext, size_t depth)
{
  if (ext.size() <= depth) ext.push_back(static_cast<index>(l.size()));
  assert(ext[depth] == static_cast<index>(l.size()) && "FluidTensor: ragged initializer list");
  for (const auto& x : l)
  {
    if constexpr (std::is_convertible_v<std::decay_t<decltype(x)>, T>) out.push_back(static_cast<T>(x));
    else flatten<T>(x, out, ext, depth + 1);
  }
}

// logical row-major iteration over a strided view
template <typename T, size_t N>
class SliceIterator
{
public:
  using value_type = std::remove_const_t<T>;
  using reference = T&;
  using pointer = T*;
  using difference_type = std::ptrdiff_t;
  using iterator_category = std::forward_iterator_tag;

  SliceIterator(const FluidTensorSlice<N>& d, T* base, bool end = false) : mDesc(d), mBase(base)
  {
    mIdx.fill(0);
    mEnd = end || d.size == 0;
  }
  reference operator*() const
  {
    index off = mDesc.start;
    for (size_t i = 0; i < N; ++i) off += mIdx[i] * mDesc.strides[i];
    return mBase[off];
  }
  pointer        operator->() const { return &**this; }
  SliceIterator& operator++()
  {
    for (size_t i = N; i-- > 0;)
    {
      if (++mIdx[i] < mDesc.extents[i]) return *this;
      mIdx[i] = 0;
    }
    mEnd = true;
    return *this;
  }
  SliceIterator operator++(int)
  {
    SliceIterator t(*this);
    ++*this;
    return t;
  }
  bool operator==(const SliceIterator& o) const { return mEnd == o.mEnd && (mEnd || mIdx == o.mIdx); }
  bool operator!=(const SliceIterator& o) const { return !(*this == o); }

private:
  FluidTensorSlice<N>  mDesc;
  T*                   mBase;
  std::array<index, N> mIdx;
  bool                 mEnd;
};

template <typename... Args>
constexpr bool allIndices()
{
  return (std::is_convertible_v<Args, index> && ...);
}
template <typename... Args>
constexpr bool hasSlice()
{
  return ((std::is_same_v<std::decay_t<Args>, Slice> || std::is_convertible_v<Args, index>) &&...) &&
         (std::is_same_v<std::decay_t<Args>, Slice> || ...);
}

inline void resolve(const Slice& s, index extent, index stride, index& start, index& outExtent, index& outStride)
{ // clamping rules of FluidTensor_Support.hpp:447-464
  Slice r(s);
  if (r.start < 0 || r.start >= extent) r.start = 0;
  if (r.length < 0 || r.start + (r.length - 1) * r.stride >= extent) r.length = (extent - r.start + r.stride - 1) / r.stride;
  start += r.start * stride;
  outExtent = r.length;
  outStride = stride * r.stride;
}
inline void resolve(index i, index extent, index stride, index& start, index& outExtent, index& outStride)
{
  assert(i >= 0 && i < extent && "FluidTensor: slice index out of range");
  start += i * stride;
  outExtent = 1;
  outStride = stride;
}
} // namespace impl

template <typename T, size_t N>
using FluidTensorInitializer = typename impl::InitList<T, N>::type;

// ---------------------------------------------------------------------------------------------------------------
// 0-order view: a reference to one element (FluidTensor.hpp:743-787)
template <typename T>
class FluidTensorView<T, 0>
{
public:
  FluidTensorView(const FluidTensorSlice<0>& s, T* p) : mRef(p + s.start) {}
  explicit FluidTensorView(T* p) : mRef(p) {}
  FluidTensorView& operator=(const std::remove_const_t<T>& v)
  {
    *mRef = v;
    return *this;
  }
  T&       operator()() { return *mRef; }
  const T& operator()() const { return *mRef; }
           operator T&() { return *mRef; }
           operator const T&() const { return *mRef; }
  index    size() const { return 1; }

private:
  T* mRef;
};

// ---------------------------------------------------------------------------------------------------------------
template <typename T, size_t N>
class FluidTensorView
{
public:
  static constexpr size_t order = N;
  using type = std::remove_reference_t<T>;
  using pointer = T*;
  using iterator = impl::SliceIterator<T, N>;
  using const_iterator = impl::SliceIterator<const T, N>;

  FluidTensorView() = delete;
  FluidTensorView(const FluidTensorSlice<N>& s, T* p) : mDesc(s), mRef(p) {}

  // (pointer, start, dims...)  FluidTensor.hpp:468-474
  template <typename... Dims, typename = std::enable_if_t<sizeof...(Dims) == N && impl::allIndices<Dims...>()>>
  FluidTensorView(T* p, index start, Dims... dims) : mDesc(start, dims...), mRef(p)
  {}

  // new leading axis of extent 1  FluidTensor.hpp:478-488
  template <size_t M = N, typename = std::enable_if_t<(M > 1)>>
  explicit FluidTensorView(FluidTensorView<T, N - 1> x) : mRef(x.baseData())
  {
    mDesc.start = x.descriptor().start;
    mDesc.extents[0] = 1;
    mDesc.strides[0] = x.descriptor().size;
    for (size_t i = 0; i < N - 1; ++i)
    {
      mDesc.extents[i + 1] = x.descriptor().extents[i];
      mDesc.strides[i + 1] = x.descriptor().strides[i];
    }
    mDesc.size = x.descriptor().size;
  }

  FluidTensorView(FluidTensor<std::remove_const_t<T>, N>&&) = delete; // no views of temporaries (FluidTensor.hpp:495)
  FluidTensorView(const FluidTensorView&) = default;
  FluidTensorView& operator=(const FluidTensorView&) = default; // shallow re-seat; deep copy is <<=

  operator FluidTensorView<const T, N>() const { return {mDesc, mRef}; }

  // deep copies (FluidTensor.hpp:507-578); extents must match
  template <typename U>
  FluidTensorView& operator<<=(const FluidTensorView<U, N> x)
  {
    static_assert(std::is_convertible_v<U, T>, "Can't convert between types");
    assert(sameExtents(x.descriptor()) && "FluidTensorView: <<= needs matching extents");
    auto d = begin();
    for (auto s = x.begin(); s != x.end(); ++s, ++d) *d = static_cast<std::remove_const_t<T>>(*s);
    return *this;
  }
  template <typename U>
  FluidTensorView& operator<<=(const FluidTensor<U, N>& x)
  {
    return *this <<= FluidTensorView<const U, N>(x);
  }

  template <typename... Dims, typename = std::enable_if_t<sizeof...(Dims) == N>>
  void reset(T* p, index start, Dims... dims)
  {
    mRef = p;
    mDesc = FluidTensorSlice<N>(start, dims...);
  }

  template <typename... Args>
  std::enable_if_t<impl::allIndices<Args...>(), T&> operator()(Args... args) const
  {
    return mRef[mDesc(args...)];
  }

  template <typename... Args>
  std::enable_if_t<impl::hasSlice<Args...>(), FluidTensorView<T, N>> operator()(const Args&... args) const
  {
    static_assert(sizeof...(Args) == N, "wrong number of slices");
    FluidTensorSlice<N> d;
    d.start = mDesc.start;
    size_t i = 0;
    ((impl::resolve(args, mDesc.extents[i], mDesc.strides[i], d.start, d.extents[i], d.strides[i]), ++i), ...);
    d.size = 1;
    for (auto e : d.extents) d.size *= e;
    return {d, mRef};
  }

  iterator       begin() { return {mDesc, mRef}; }
  iterator       end() { return {mDesc, mRef, true}; }
  const_iterator begin() const { return {mDesc, mRef}; }
  const_iterator end() const { return {mDesc, mRef, true}; }

  FluidTensorView<T, N - 1> row(index i) const
  {
    assert(i >= 0 && i < mDesc.extents[0] && "FluidTensorView: row out of range");
    return sub(0, i);
  }
  FluidTensorView<T, N - 1> col(index i) const
  {
    assert(i >= 0 && i < mDesc.extents[N - 1] && "FluidTensorView: col out of range");
    return sub(N - 1, i);
  }
  FluidTensorView<T, N - 1> operator[](index i) const { return row(i); }

  index extent(index n) const { return mDesc.extents[asUnsigned(n)]; }
  index rows() const { return mDesc.extents[0]; }
  index cols() const
  {
    if constexpr (N >= 2) return mDesc.extents[1];
    else return 1;
  }
  index size() const { return mDesc.size; }
  void  fill(const std::remove_const_t<T>& v) const
  {
    for (auto it = iterator(mDesc, mRef); it != iterator(mDesc, mRef, true); ++it) *it = v;
  }
  FluidTensorView transpose() const { return {mDesc.transpose(), mRef}; }

  template <typename F>
  FluidTensorView& apply(F f)
  {
    for (auto& x : *this) f(x);
    return *this;
  }
  template <typename M, typename F>
  FluidTensorView& apply(M m, F f)
  {
    assert(m.size() == size());
    auto j = m.begin();
    for (auto i = begin(); i != end(); ++i, ++j) f(*i, *j);
    return *this;
  }

  // pointer to the first element of the view (FluidTensor.hpp:713); nullptr views stay nullptr
  T*                         data() const { return mRef ? mRef + mDesc.start : nullptr; }
  T*                         baseData() const { return mRef; }
  const FluidTensorSlice<N>& descriptor() const { return mDesc; }
  bool                       operator==(const FluidTensorView& o) const { return mRef == o.mRef && mDesc == o.mDesc; }

  friend std::ostream& operator<<(std::ostream& o, const FluidTensorView& t)
  {
    bool first = true;
    for (auto it = t.begin(); it != t.end(); ++it)
    {
      o << (first ? "" : ",") << *it;
      first = false;
    }
    return o;
  }

private:
  template <size_t M>
  bool sameExtents(const FluidTensorSlice<M>& o) const
  {
    if constexpr (M != N) return false;
    else return mDesc.extents == o.extents;
  }
  FluidTensorView<T, N - 1> sub(size_t dim, index i) const
  {
    if constexpr (N == 1) { return FluidTensorView<T, 0>(mRef + mDesc.start + i * mDesc.strides[0]); }
    else
    {
      FluidTensorSlice<N - 1> d;
      d.start = mDesc.start + i * mDesc.strides[dim];
      size_t j = 0;
      d.size = 1;
      for (size_t k = 0; k < N; ++k)
        if (k != dim)
        {
          d.extents[j] = mDesc.extents[k];
          d.strides[j] = mDesc.strides[k];
          d.size *= d.extents[j];
          ++j;
        }
      return {d, mRef};
    }
  }

  FluidTensorSlice<N> mDesc;
  T*                  mRef;
};

// ---------------------------------------------------------------------------------------------------------------
template <typename T, size_t N>
class FluidTensor
{
public:
  static constexpr size_t order = N;
  using type = std::remove_reference_t<T>;
  using Container = rt::vector<std::remove_const_t<T>>;
  using iterator = typename Container::iterator;
  using const_iterator = typename Container::const_iterator;

  explicit FluidTensor(Allocator& = FluidDefaultAllocator()) {}
  FluidTensor(const FluidTensor&) = default;
  FluidTensor(FluidTensor&&) noexcept = default;
  FluidTensor& operator=(const FluidTensor&) = default;
  FluidTensor& operator=(FluidTensor&&) noexcept = default;

  template <typename... Dims,
            typename = std::enable_if_t<sizeof...(Dims) == N && impl::allIndices<Dims...>()>>
  explicit FluidTensor(Dims... dims) : mDesc(0, dims...), mContainer(asUnsigned(mDesc.size))
  {}
  template <typename... Dims,
            typename = std::enable_if_t<sizeof...(Dims) == N && impl::allIndices<Dims...>()>>
  FluidTensor(Allocator&, Dims... dims) : FluidTensor(dims...)
  {}

  // converting deep copies (FluidTensor.hpp:101-124)
  template <typename U>
  explicit FluidTensor(const FluidTensor<U, N>& x) : mDesc(x.descriptor()), mContainer(x.begin(), x.end())
  {}
  template <typename U>
  explicit FluidTensor(FluidTensorView<U, N> x) : mDesc(0, x.descriptor().extents), mContainer(asUnsigned(x.size()))
  {
    FluidTensorView<T, N>(*this) <<= x;
  }

  FluidTensor(FluidTensorInitializer<T, N> init)
  {
    std::vector<std::remove_const_t<T>> flat;
    std::vector<index>                  ext;
    impl::flatten<std::remove_const_t<T>>(init, flat, ext, 0);
    assert(ext.size() == N);
    std::array<index, N> e{};
    std::copy(ext.begin(), ext.end(), e.begin());
    mDesc = FluidTensorSlice<N>(0, e);
    mContainer.assign(flat.begin(), flat.end());
  }

  // 1-D from pointer + stride (FluidTensor.hpp:229-236)
  template <size_t D = N, typename = std::enable_if_t<D == 1>>
  FluidTensor(const T* input, index dim, index stride = 1) : mDesc(0, dim), mContainer(asUnsigned(dim))
  {
    for (index i = 0; i < dim; ++i) mContainer[asUnsigned(i)] = input[i * stride];
  }

  template <typename U, size_t M>
  FluidTensor& operator<<=(const FluidTensorView<U, M> x)
  {
    FluidTensorView<T, N>(*this) <<= x;
    return *this;
  }

  operator FluidTensorView<T, N>() { return {mDesc, data()}; }
  operator FluidTensorView<const T, N>() const { return {mDesc, data()}; }

  FluidTensorView<T, N - 1>       row(index i) { return FluidTensorView<T, N>(*this).row(i); }
  FluidTensorView<const T, N - 1> row(index i) const { return FluidTensorView<const T, N>(*this).row(i); }
  FluidTensorView<T, N - 1>       col(index i) { return FluidTensorView<T, N>(*this).col(i); }
  FluidTensorView<const T, N - 1> col(index i) const { return FluidTensorView<const T, N>(*this).col(i); }
  FluidTensorView<T, N - 1>       operator[](index i) { return row(i); }
  FluidTensorView<const T, N - 1> operator[](index i) const { return row(i); }

  template <typename... Args>
  std::enable_if_t<impl::allIndices<Args...>(), T&> operator()(Args... args)
  {
    return mContainer[asUnsigned(mDesc(args...))];
  }
  template <typename... Args>
  std::enable_if_t<impl::allIndices<Args...>(), const T&> operator()(Args... args) const
  {
    return mContainer[asUnsigned(mDesc(args...))];
  }
  template <typename... Args>
  std::enable_if_t<impl::hasSlice<Args...>(), FluidTensorView<T, N>> operator()(const Args&... args)
  {
    return FluidTensorView<T, N>(*this)(args...);
  }
  template <typename... Args>
  std::enable_if_t<impl::hasSlice<Args...>(), FluidTensorView<const T, N>> operator()(const Args&... args) const
  {
    return FluidTensorView<const T, N>(*this)(args...);
  }

  iterator       begin() { return mContainer.begin(); }
  iterator       end() { return mContainer.end(); }
  const_iterator begin() const { return mContainer.cbegin(); }
  const_iterator end() const { return mContainer.cend(); }

  index extent(index n) const { return mDesc.extents[asUnsigned(n)]; }
  index rows() const { return mDesc.extents[0]; }
  index cols() const
  {
    if constexpr (N >= 2) return mDesc.extents[1];
    else return 1;
  }
  index                      size() const { return asSigned(mContainer.size()); }
  const FluidTensorSlice<N>& descriptor() const { return mDesc; }
  const T*                   data() const { return mContainer.data(); }
  T*                         data() { return mContainer.data(); }

  template <typename... Dims, typename = std::enable_if_t<sizeof...(Dims) == N>>
  void resize(Dims... dims)
  {
    mDesc = FluidTensorSlice<N>(0, dims...);
    mContainer.resize(asUnsigned(mDesc.size));
  }
  void resizeDim(index dim, index amount)
  {
    if (amount == 0) return;
    auto e = mDesc.extents;
    e[asUnsigned(dim)] += amount;
    mDesc = FluidTensorSlice<N>(0, e);
    mContainer.resize(asUnsigned(mDesc.size));
  }
  void fill(T v) { std::fill(mContainer.begin(), mContainer.end(), v); }

  FluidTensorView<T, N>       transpose() { return {mDesc.transpose(), data()}; }
  FluidTensorView<const T, N> transpose() const { return {mDesc.transpose(), data()}; }

  template <typename F>
  FluidTensor& apply(F f)
  {
    for (auto& x : mContainer) f(x);
    return *this;
  }

  bool operator==(const FluidTensor& rhs) const { return mContainer == rhs.mContainer; }
  bool operator!=(const FluidTensor& rhs) const { return !(*this == rhs); }

  friend std::ostream& operator<<(std::ostream& o, const FluidTensor& t)
  {
    return o << FluidTensorView<const T, N>(t);
  }

private:
  FluidTensorSlice<N> mDesc;
  Container           mContainer;
};

} // namespace fluid
