// Oracle build shim: intentionally empty (nothing from this header is used on the sync path).
#ifndef COS_SHIM_CAFFE_PROTO_CAFFE_PB_H_
#define COS_SHIM_CAFFE_PROTO_CAFFE_PB_H_
#endif
