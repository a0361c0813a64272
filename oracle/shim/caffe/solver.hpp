// Oracle build shim: Solver / Net reduced to what P2PSyncCPU and SocketSyncCPU
// need -- learnable_params(), param().device_id(), iter(), add_callback() and
// the Callback interface (caffe/solver.hpp:78-89) plus helpers for the driver
// to fire the two callbacks in Solver::Step order (solver.cpp:214-216,250-252).
#ifndef COS_SHIM_CAFFE_SOLVER_HPP_
#define COS_SHIM_CAFFE_SOLVER_HPP_
#include <vector>
#include "caffe/blob.hpp"
#include "caffe/common.hpp"
namespace caffe {
class SolverParameter {
 public:
  int device_id() const { return 0; }
};
template <typename Dtype>
class Net {
 public:
  const vector<Blob<Dtype>*>& learnable_params() const { return learnable_; }
  void add_param(int count) {
    owned_.push_back(shared_ptr<Blob<Dtype> >(new Blob<Dtype>(count)));
    learnable_.push_back(owned_.back().get());
  }
 private:
  vector<shared_ptr<Blob<Dtype> > > owned_;
  vector<Blob<Dtype>*> learnable_;
};
template <typename Dtype>
class Solver {
 public:
  class Callback {
   protected:
    virtual void on_start() = 0;
    virtual void on_gradients_ready() = 0;
    template <typename T> friend class Solver;
  };
  Solver() : net_(new Net<Dtype>()), iter_(0) {}
  const shared_ptr<Net<Dtype> >& net() const { return net_; }
  const SolverParameter& param() const { return param_; }
  int iter() const { return iter_; }
  void add_callback(Callback* c) { callbacks_.push_back(c); }
  // driver helpers (Solver::Step fires them in this order)
  void fire_on_start() { for (size_t i = 0; i < callbacks_.size(); ++i) callbacks_[i]->on_start(); }
  void fire_on_gradients_ready() { for (size_t i = 0; i < callbacks_.size(); ++i) callbacks_[i]->on_gradients_ready(); }
  void advance() { ++iter_; }
 private:
  shared_ptr<Net<Dtype> > net_;
  SolverParameter param_;
  int iter_;
  vector<Callback*> callbacks_;
};
}  // namespace caffe
#endif
