// Oracle build shim: intentionally empty (nothing from this header is used on the sync path).
#ifndef COS_SHIM_CAFFE_INTERNAL_THREAD_HPP_
#define COS_SHIM_CAFFE_INTERNAL_THREAD_HPP_
#endif
