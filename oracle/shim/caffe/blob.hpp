// Oracle build shim: a Blob whose data()/diff() can be re-pointed into the
// flat Params buffers (blob.hpp / syncedmem.hpp behaviour used by
// parallel_cpu.cpp:27-57).
#ifndef COS_SHIM_CAFFE_BLOB_HPP_
#define COS_SHIM_CAFFE_BLOB_HPP_
#include <vector>
#include "caffe/common.hpp"
namespace caffe {
class ShimMem {
 public:
  explicit ShimMem(size_t n) : own_(n), ptr_(own_.data()) {}
  const void* cpu_data() const { return ptr_; }
  void* mutable_cpu_data() { return ptr_; }
  void set_cpu_data(void* p) { ptr_ = p; }
  void set_gpu_data(void* p) { ptr_ = p; }
 private:
  std::vector<double> own_;  // big enough for float or double
  void* ptr_;
};
template <typename Dtype>
class Blob {
 public:
  explicit Blob(int count) : count_(count), data_(new ShimMem(count)), diff_(new ShimMem(count)) {}
  int count() const { return count_; }
  const shared_ptr<ShimMem>& data() const { return data_; }
  const shared_ptr<ShimMem>& diff() const { return diff_; }
  Dtype* mutable_cpu_data() { return static_cast<Dtype*>(data_->mutable_cpu_data()); }
 private:
  int count_;
  shared_ptr<ShimMem> data_, diff_;
};
}  // namespace caffe
#endif
