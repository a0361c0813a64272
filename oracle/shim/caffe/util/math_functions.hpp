// Oracle build shim: the math helpers the sync path calls, as scalar loops.
//   caffe_add       = vsAdd, y[i] = a[i] + b[i]   (caffe/util/mkl_alternate.hpp:60-75)
//   caffe_cpu_scale = scopy + sscal, y[i] = alpha * x[i] (math_functions.cpp:362-366)
#ifndef COS_SHIM_CAFFE_MATH_FUNCTIONS_HPP_
#define COS_SHIM_CAFFE_MATH_FUNCTIONS_HPP_
#include <cstring>
namespace caffe {
inline void caffe_memset(const size_t N, const int alpha, void* X) { std::memset(X, alpha, N); }
template <typename Dtype> void caffe_copy(const int N, const Dtype* X, Dtype* Y) {
  if (X != Y) std::memcpy(Y, X, sizeof(Dtype) * N);
}
template <typename Dtype> void caffe_set(const int N, const Dtype alpha, Dtype* Y) {
  for (int i = 0; i < N; ++i) Y[i] = alpha;
}
template <typename Dtype> void caffe_add(const int n, const Dtype* a, const Dtype* b, Dtype* y) {
  for (int i = 0; i < n; ++i) y[i] = a[i] + b[i];
}
template <typename Dtype> void caffe_cpu_scale(const int n, const Dtype alpha, const Dtype* x, Dtype* y) {
  if (x != y) std::memcpy(y, x, sizeof(Dtype) * n);
  for (int i = 0; i < n; ++i) y[i] = alpha * y[i];
}
}  // namespace caffe
#endif
