// Test-infrastructure shim (oracle/): maps the tiny subset of Boost.Filesystem the
// reference's core sources use (src/codegen/compiler.cc:94-136, src/input/watcher.cc:53-56)
// onto std::filesystem. Boost is not installed in this image. Not product code.
#pragma once
#include <filesystem>
#include <fstream>
#include <string>
namespace boost {
namespace filesystem {
using std::filesystem::path;
using std::filesystem::create_directories;
using std::filesystem::exists;
using std::filesystem::rename;
using std::filesystem::canonical;
using std::filesystem::directory_iterator;
using std::filesystem::is_regular_file;
using std::filesystem::remove;
using std::filesystem::remove_all;
using std::filesystem::temp_directory_path;
using std::filesystem::is_directory;
class ofstream : public std::ofstream {
public:
  ofstream() {}
  explicit ofstream(const char *p) : std::ofstream(p) {}
  explicit ofstream(const std::string &p) : std::ofstream(p) {}
  explicit ofstream(const path &p) : std::ofstream(p) {}
};
} // namespace filesystem
} // namespace boost
