// Oracle build shim: interface-compatible declaration of caffe::Params<Dtype>
// (caffe-public/include/caffe/parallel.hpp:22-45) -- one flat data_ and diff_
// array of size_ elements.  The constructor (total_size rule of
// parallel.cpp:60-74: max(1, sum of blob counts)) is defined in ref_driver.cpp
// because the reference defines it inside libcaffe's parallel.cpp.
#ifndef COS_SHIM_CAFFE_PARALLEL_HPP_
#define COS_SHIM_CAFFE_PARALLEL_HPP_
#include "caffe/common.hpp"
#include "caffe/solver.hpp"
namespace caffe {
template <typename Dtype>
class Params {
 public:
  explicit Params(shared_ptr<Solver<Dtype> > root_solver);
  virtual ~Params() {}
  size_t size() const { return size_; }
  Dtype* data() const { return data_; }
  Dtype* diff() const { return diff_; }
 protected:
  const size_t size_;
  Dtype* data_;
  Dtype* diff_;
  DISABLE_COPY_AND_ASSIGN(Params);
};
}  // namespace caffe
#endif
