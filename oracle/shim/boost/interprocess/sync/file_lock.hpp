// Test-infrastructure shim (oracle/): boost::interprocess::file_lock over flock(2),
// used by src/codegen/compiler.cc:110-111 to serialise JIT cache writers.
#pragma once
#include <fcntl.h>
#include <stdexcept>
#include <sys/file.h>
#include <unistd.h>
namespace boost {
namespace interprocess {
class file_lock {
public:
  explicit file_lock(const char *name) : fd_(::open(name, O_RDWR)) {
    if (fd_ < 0) throw std::runtime_error("file_lock: cannot open lock file");
  }
  file_lock(const file_lock &) = delete;
  ~file_lock() { if (fd_ >= 0) ::close(fd_); }
  void lock() { ::flock(fd_, LOCK_EX); }
  void unlock() { ::flock(fd_, LOCK_UN); }
private:
  int fd_;
};
} // namespace interprocess
} // namespace boost
