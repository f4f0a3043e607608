// Test-infrastructure shim (oracle/): boost::interprocess::scoped_lock (RAII) for compiler.cc:111.
#pragma once
namespace boost {
namespace interprocess {
template <typename M> class scoped_lock {
public:
  explicit scoped_lock(M &m) : m_(m) { m_.lock(); }
  ~scoped_lock() { m_.unlock(); }
  scoped_lock(const scoped_lock &) = delete;
private:
  M &m_;
};
} // namespace interprocess
} // namespace boost
