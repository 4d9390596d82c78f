// Loader for libflucoma_b200.so: dlopen + dlsym of the single entry point fb200_get_api (include/flucoma_b200.h).
// The shims in algorithms/public/*.hpp call the device through this table and nothing else.  There is no CPU
// fallback: if the library or a CUDA device is missing, B200Backend::get() throws (offline clients turn that into
// Result::Status::kError).
#pragma once
#include "../../../../../include/flucoma_b200.h"
#include "../../data/FluidTensor.hpp"
#include <dlfcn.h>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace fluid {
namespace b200 {

class B200Backend
{
public:
  static const fb200_api& get()
  {
    static B200Backend inst;
    return *inst.mApi;
  }

private:
  B200Backend()
  {
    const char* env = std::getenv("FLUCOMA_B200_LIB");
    const char* names[] = {env, "libflucoma_b200.so"};
    for (const char* n : names)
    {
      if (!n) continue;
      mHandle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (mHandle) break;
    }
    if (!mHandle) throw std::runtime_error(std::string("flucoma-b200: cannot load libflucoma_b200.so: ") + dlerror());
    using GetApi = const fb200_api* (*) (uint32_t);
    auto get_api = reinterpret_cast<GetApi>(dlsym(mHandle, "fb200_get_api"));
    if (!get_api) throw std::runtime_error("flucoma-b200: fb200_get_api not exported");
    mApi = get_api(FB200_ABI_VERSION);
    if (!mApi) throw std::runtime_error("flucoma-b200: ABI version mismatch");
  }
  void*            mHandle{nullptr};
  const fb200_api* mApi{nullptr};
};

// RAII plan
class Plan
{
public:
  Plan(index win, index fft, index hop, index maxRank = 64, int device = 0)
  {
    fb200_config cfg{};
    cfg.struct_size = sizeof(cfg);
    cfg.device = device;
    cfg.win = static_cast<int32_t>(win);
    cfg.hop = static_cast<int32_t>(hop);
    cfg.fft = static_cast<int32_t>(fft);
    cfg.max_rank = static_cast<int32_t>(maxRank);
    int32_t st = B200Backend::get().plan_create(&cfg, &mPlan);
    if (st != FB200_OK)
      throw std::runtime_error(std::string("flucoma-b200: plan_create failed: ") + B200Backend::get().last_error(nullptr));
  }
  ~Plan()
  {
    if (mPlan) B200Backend::get().plan_destroy(mPlan);
  }
  Plan(const Plan&) = delete;
  Plan& operator=(const Plan&) = delete;
  Plan(Plan&& o) noexcept : mPlan(o.mPlan) { o.mPlan = nullptr; }
  Plan& operator=(Plan&& o) noexcept
  {
    std::swap(mPlan, o.mPlan);
    return *this;
  }
  fb200_plan* get() const { return mPlan; }
  void        check(int32_t st) const
  {
    if (st < 0) throw std::runtime_error(std::string("flucoma-b200: ") + B200Backend::get().last_error(mPlan));
  }

private:
  fb200_plan* mPlan{nullptr};
};

// dense row-major copies of (possibly strided / transposed) views: the C ABI takes dense arrays
template <typename T, size_t N>
std::vector<std::remove_const_t<T>> pack(const FluidTensorView<T, N>& v)
{
  std::vector<std::remove_const_t<T>> out;
  out.reserve(asUnsigned(v.size()));
  for (auto it = v.begin(); it != v.end(); ++it) out.push_back(*it);
  return out;
}
template <typename T, size_t N, typename U>
void unpack(const std::vector<U>& src, FluidTensorView<T, N> v)
{
  auto s = src.begin();
  for (auto it = v.begin(); it != v.end(); ++it, ++s) *it = static_cast<T>(*s);
}

} // namespace b200
} // namespace fluid
