// Test-infrastructure shim (oracle/): boost::replace_all for src/util/statsd.cc:41.
#pragma once
#include <string>
namespace boost {
inline void replace_all(std::string &s, const std::string &from, const std::string &to) {
  if (from.empty()) return;
  size_t pos = 0;
  while ((pos = s.find(from, pos)) != std::string::npos) {
    s.replace(pos, from.size(), to);
    pos += to.size();
  }
}
namespace algorithm { using boost::replace_all; }
} // namespace boost
