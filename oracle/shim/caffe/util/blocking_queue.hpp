// Oracle build shim: caffe::BlockingQueue (mutex + condition variable FIFO).
#ifndef COS_SHIM_CAFFE_BLOCKING_QUEUE_HPP_
#define COS_SHIM_CAFFE_BLOCKING_QUEUE_HPP_
#include <condition_variable>
#include <mutex>
#include <queue>
#include <string>
namespace caffe {
template <typename T>
class BlockingQueue {
 public:
  void push(const T& t) {
    { std::lock_guard<std::mutex> g(m_); q_.push(t); }
    cv_.notify_one();
  }
  T pop(const std::string& = "") {
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [this] { return !q_.empty(); });
    T t = q_.front();
    q_.pop();
    return t;
  }
  size_t size() { std::lock_guard<std::mutex> g(m_); return q_.size(); }
 private:
  std::queue<T> q_;
  std::mutex m_;
  std::condition_variable cv_;
};
}  // namespace caffe
#endif
