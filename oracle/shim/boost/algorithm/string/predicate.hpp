// Test-infrastructure shim (oracle/): boost::algorithm::ends_with for src/input/watcher.cc:85.
#pragma once
#include <string>
namespace boost {
namespace algorithm {
inline bool ends_with(const std::string &s, const std::string &suffix) {
  return s.size() >= suffix.size() &&
         s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}
inline bool starts_with(const std::string &s, const std::string &prefix) {
  return s.size() >= prefix.size() && s.compare(0, prefix.size(), prefix) == 0;
}
} // namespace algorithm
using algorithm::ends_with;
using algorithm::starts_with;
} // namespace boost
