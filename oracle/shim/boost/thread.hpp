// Oracle build shim: boost::mutex + scoped_lock over std::mutex.
#ifndef COS_SHIM_BOOST_THREAD_HPP_
#define COS_SHIM_BOOST_THREAD_HPP_
#include <mutex>
namespace boost {
class mutex {
 public:
  class scoped_lock {
   public:
    explicit scoped_lock(mutex& m) : g_(m.m_) {}
   private:
    std::lock_guard<std::mutex> g_;
  };
  void lock() { m_.lock(); }
  void unlock() { m_.unlock(); }
 private:
  std::mutex m_;
};
}  // namespace boost
#endif
