// Test-infrastructure shim (oracle/): boost::algorithm::join for src/codegen/compiler.cc:124.
#pragma once
#include <string>
namespace boost {
namespace algorithm {
template <typename Seq> std::string join(const Seq &parts, const std::string &sep) {
  std::string out;
  bool first = true;
  for (const auto &p : parts) {
    if (!first) out += sep;
    out += p;
    first = false;
  }
  return out;
}
} // namespace algorithm
} // namespace boost
