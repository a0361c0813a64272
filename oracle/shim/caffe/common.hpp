// Oracle build shim: the slice of caffe/common.hpp the sync path touches.
#ifndef COS_SHIM_CAFFE_COMMON_HPP_
#define COS_SHIM_CAFFE_COMMON_HPP_
#include <glog/logging.h>
#include <stdint.h>
#include <unistd.h>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>
#define DISABLE_COPY_AND_ASSIGN(classname) \
 private:                                  \
  classname(const classname&);             \
  classname& operator=(const classname&)
#define INSTANTIATE_CLASS(classname)     \
  char gInstantiationGuard##classname;   \
  template class classname<float>;       \
  template class classname<double>
#define NO_GPU LOG(FATAL) << "Cannot use GPU in CPU-only Caffe: check mode."
namespace caffe {
using std::map;
using std::shared_ptr;
using std::string;
using std::vector;
class Caffe {
 public:
  // thread-local in the reference (common.cpp:14-19)
  static int& count_() { static thread_local int c = 1; return c; }
  static int solver_count() { return count_(); }
  static void set_solver_count(int v) { count_() = v; }
};
}  // namespace caffe
#endif
