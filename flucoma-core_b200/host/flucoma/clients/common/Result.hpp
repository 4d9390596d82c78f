// Status object returned by clients (reference: include/flucoma/clients/common/Result.hpp:21-81).
#pragma once
#include <sstream>
#include <string>

namespace fluid {
namespace client {

class Result
{
public:
  enum class Status { kOk, kWarning, kError, kCancelled };

  Result() = default;
  template <typename... Args>
  Result(Status s, Args... args) : mStatus(s)
  {
    std::ostringstream os;
    (void) std::initializer_list<int>{(os << args, 0)...};
    mMsg = os.str();
  }
  bool               ok() const noexcept { return mStatus == Status::kOk; }
  Status             status() const noexcept { return mStatus; }
  const std::string& message() const noexcept { return mMsg; }
  void               set(Status s) noexcept { mStatus = s; }
  void               addMessage(const std::string& m) { mMsg += m; }
  void               reset()
  {
    mStatus = Status::kOk;
    mMsg.clear();
  }

private:
  Status      mStatus{Status::kOk};
  std::string mMsg;
};
} // namespace client
} // namespace fluid
