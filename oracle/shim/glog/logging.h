// Test-infrastructure shim (oracle/): a minimal stand-in for glog's LOG()/DLOG()/VLOG()
// streams used by the reference's core sources. Messages at WARNING and above go to
// stderr when VIYA_ORACLE_LOG is set; everything else is swallowed. Not product code.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace google {
enum LogSeverity { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
inline void InitGoogleLogging(const char *) {}
inline void InstallFailureSignalHandler() {}
class ShimLogMessage {
public:
  explicit ShimLogMessage(int sev) : sev_(sev) {}
  ~ShimLogMessage() {
    static const bool on = std::getenv("VIYA_ORACLE_LOG") != nullptr;
    if (on || sev_ >= GLOG_ERROR) std::cerr << ss_.str() << std::endl;
    if (sev_ == GLOG_FATAL) std::abort();
  }
  std::ostream &stream() { return ss_; }
private:
  int sev_;
  std::ostringstream ss_;
};
struct ShimNullStream {
  template <typename T> ShimNullStream &operator<<(const T &) { return *this; }
  ShimNullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
} // namespace google
#define LOG(sev) ::google::ShimLogMessage(::google::GLOG_##sev).stream()
#define PLOG(sev) LOG(sev)
#define LOG_IF(sev, cond) if (cond) LOG(sev)
#define VLOG(n) ::google::ShimNullStream()
#define DLOG(sev) ::google::ShimNullStream()
#define DVLOG(n) ::google::ShimNullStream()
#define CHECK(cond) if (!(cond)) LOG(FATAL) << "Check failed: " #cond " "
