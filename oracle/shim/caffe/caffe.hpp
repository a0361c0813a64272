// Oracle build shim: umbrella header.
#ifndef COS_SHIM_CAFFE_CAFFE_HPP_
#define COS_SHIM_CAFFE_CAFFE_HPP_
#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/parallel.hpp"
#include "caffe/solver.hpp"
#include "caffe/util/blocking_queue.hpp"
#include "caffe/util/math_functions.hpp"
#endif
