/*
 * sync_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See sync_oracle.h for scope, parity status and usage rules.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile)
 * -ffp-contract=off matters: every fp32 operation below is individually
 * rounded, as in the reference's scalar loops / netlib-style BLAS.
 */
#include "sync_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* socket_sync_cpu.cpp:46-54 -- size_t arithmetic, multiply first then divide:
 *   start = (peer + 0) * size_ / peers_.size();
 *   until = (peer + 1) * size_ / peers_.size();                              */
void cos_oracle_chunk(uint64_t P, int N, int peer, uint64_t* offs,
                      uint64_t* size) {
  uint64_t start = ((uint64_t)peer + 0) * P / (uint64_t)N;
  uint64_t until = ((uint64_t)peer + 1) * P / (uint64_t)N;
  *offs = start;
  *size = until - start;
}

/* parallel.cpp:60-68 / parallel_cpu.cpp:62-70 */
uint64_t cos_oracle_total_size(const int64_t* counts, int nblobs) {
  uint64_t size = 0;
  for (int i = 0; i < nblobs; ++i) size += (uint64_t)counts[i];
  return size > 0 ? size : 1;
}

/* sgd_solver.cpp:27-63.  Dtype=float.  The reference calls unqualified
 * pow()/exp() on float arguments; with libstdc++ that resolves to the double
 * versions (the float sub-expressions are still evaluated in float first) and
 * the product with base_lr (a float) is rounded to float on assignment to
 * `Dtype rate`.  Whether a given build instead picked powf is un-pinned
 * (<= 1 ulp of rate); product code computes the rate with the same formula on
 * the host, so oracle and product agree exactly. */
float cos_oracle_learning_rate(int policy, float base_lr, float gamma,
                               float power, int stepsize,
                               const int* stepvalues, int nstepvalues,
                               int max_iter, int iter, int* current_step) {
  float rate = 0.f;
  switch (policy) {
    case COS_ORACLE_LR_FIXED: /* :30-31 */
      rate = base_lr;
      break;
    case COS_ORACLE_LR_STEP: /* :32-35 */
      *current_step = iter / stepsize;
      rate = (float)((double)base_lr * pow((double)gamma, (double)*current_step));
      break;
    case COS_ORACLE_LR_EXP: /* :36-37 */
      rate = (float)((double)base_lr * pow((double)gamma, (double)iter));
      break;
    case COS_ORACLE_LR_INV: { /* :38-41 */
      float base = 1.0f + gamma * (float)iter;
      float e = -power;
      rate = (float)((double)base_lr * pow((double)base, (double)e));
      break;
    }
    case COS_ORACLE_LR_MULTISTEP: /* :42-50 */
      if (*current_step < nstepvalues && iter >= stepvalues[*current_step]) {
        (*current_step)++;
      }
      rate = (float)((double)base_lr * pow((double)gamma, (double)*current_step));
      break;
    case COS_ORACLE_LR_POLY: { /* :51-54 */
      float base = 1.0f - ((float)iter / (float)max_iter);
      rate = (float)((double)base_lr * pow((double)base, (double)power));
      break;
    }
    case COS_ORACLE_LR_SIGMOID: { /* :55-58 */
      float x = -gamma * ((float)iter - (float)stepsize);
      double d = (double)1.0f / ((double)1.0f + exp((double)x));
      rate = (float)((double)base_lr * d);
      break;
    }
    default:
      abort(); /* LOG(FATAL) << "Unknown learning rate policy" :60 */
  }
  return rate;
}

/* socket_sync_cpu.cpp:102-105 -> sync(true) :135-163 with the buffer table of
 * CreateMasterBuffers :56-72 (data_send_ = data_ + own_offs_) and
 * CreateWorkerBuffers :74-91 (data_recv_ -> data_ + offs(peer)): every rank
 * ends with data_[shard(p)] == owner p's copy, a bit copy (socket.cpp:399). */
void cos_oracle_all_gather(int N, uint64_t P, float* const* data) {
  for (int owner = 0; owner < N; ++owner) {
    uint64_t offs, size;
    cos_oracle_chunk(P, N, owner, &offs, &size);
    for (int r = 0; r < N; ++r) {
      if (r == owner) continue;
      memcpy(data[r] + offs, data[owner] + offs, size * sizeof(float));
    }
  }
}

/* parallel_cpu.cpp:120-122: caffe_cpu_scale(size_, Dtype(1.0 / solver_count),
 * diff_, diff_) = cblas_scopy + cblas_sscal (math_functions.cpp:362-366): one
 * fp32 multiply per element by the float literal (float)(1.0 / N). */
void cos_oracle_scale(int solver_count, uint64_t P, float* diff) {
  const float inv = (float)(1.0 / (double)solver_count);
  for (uint64_t i = 0; i < P; ++i) diff[i] = inv * diff[i];
}

/* socket_sync_cpu.cpp:108-133.  Sends (:112-119) capture diff_[chunk(p)] of
 * the sender at send time; since every rank has already run the scale and a
 * rank's own shard is the only region it later overwrites (:129 writes
 * diff_ + own_offs_), and nobody sends its own shard, "all sends first, then
 * all ordered adds" is exactly equivalent to the concurrent execution.
 * Add is caffe_add(n, src=recv, dst, dst) -> y[i] = a[i] + b[i] with a = recv
 * (mkl_alternate.hpp:60-75). */
void cos_oracle_reduce_scatter(int N, uint64_t P, float* const* diff) {
  for (int r = 0; r < N; ++r) {
    uint64_t offs, size;
    cos_oracle_chunk(P, N, r, &offs, &size);
    float* dst = diff[r] + offs;
    int peer = r + 1;
    for (int n = 0; n < N - 1; ++n) {
      if (peer == N) peer = 0;
      const float* src = diff[peer] + offs;
      for (uint64_t i = 0; i < size; ++i) dst[i] = src[i] + dst[i];
      peer++;
    }
  }
}

/* ApplyUpdate sgd_solver.cpp:102-116 per learnable blob k (ClipGradients is a
 * no-op for clip_gradients < 0, Normalize a no-op for iter_size == 1):
 *   Regularize :145-172 (CPU):  local_decay = weight_decay * decay_mult_k;
 *       L2: if (local_decay) caffe_axpy(local_decay, w, g)        g = ld*w + g
 *       L1: caffe_cpu_sign(w -> temp); caffe_axpy(local_decay, temp, g)
 *                                                             g = ld*sign(w) + g
 *   ComputeUpdateValue :213-229:    local_rate = rate * lr_mult_k;
 *       caffe_cpu_axpby(local_rate, g, momentum, h)
 *         = cblas_sscal(momentum, h); cblas_saxpy(local_rate, g, h)
 *           (mkl_alternate.hpp:83-88)                          h = lr*g + (m*h)
 *       caffe_copy(h -> g)                                     g = h
 *   Net::Update net.cpp:924 -> Blob::Update blob.cpp:162-179:
 *       caffe_axpy(-1, g, w)                                   w = (-1*g) + w
 * saxpy is taken as the un-fused y[i] = fl(fl(a*x[i]) + y[i]).               */
void cos_oracle_apply_update_ex(uint64_t begin, uint64_t end, float* data,
                                float* diff, float* hist, int nblobs,
                                const int64_t* counts, const float* lr_mult,
                                const float* decay_mult, float rate,
                                float momentum, float weight_decay, int l1) {
  uint64_t blob_begin = 0;
  for (int k = 0; k < nblobs; ++k) {
    uint64_t blob_end = blob_begin + (uint64_t)counts[k];
    uint64_t lo = blob_begin > begin ? blob_begin : begin;
    uint64_t hi = blob_end < end ? blob_end : end;
    const float local_decay = weight_decay * decay_mult[k];
    const float local_rate = rate * lr_mult[k];
    for (uint64_t i = lo; i < hi; ++i) {
      float g = diff[i];
      float w = data[i];
      float h = hist[i];
      if (local_decay != 0.f) {
        /* L2 :155-160 caffe_axpy(local_decay, data, diff); L1 :161-168
         * caffe_cpu_sign(data -> temp_) = (0 < w) - (w < 0)
         * (math_functions.hpp:113-118 caffe_sign), then the same axpy on temp_ */
        float x = l1 ? (float)((0.f < w) - (w < 0.f)) : w;
        float t = local_decay * x;
        g = t + g;
      }
      h = momentum * h;
      {
        float t = local_rate * g;
        h = t + h;
      }
      g = h;
      {
        float t = -1.0f * g;
        w = t + w;
      }
      diff[i] = g;
      hist[i] = h;
      data[i] = w;
    }
    blob_begin = blob_end;
  }
}

void cos_oracle_apply_update(uint64_t begin, uint64_t end, float* data,
                             float* diff, float* hist, int nblobs,
                             const int64_t* counts, const float* lr_mult,
                             const float* decay_mult, float rate,
                             float momentum, float weight_decay) {
  cos_oracle_apply_update_ex(begin, end, data, diff, hist, nblobs, counts,
                             lr_mult, decay_mult, rate, momentum, weight_decay, 0);
}

/* Solver::Step solver.cpp:194-273 minus ClearParamDiffs/ForwardBackward (the
 * caller provides the local gradients in diff[r]). */
void cos_oracle_step_ex(int N, uint64_t P, float* const* data, float* const* diff,
                        float* const* hist, int nblobs, const int64_t* counts,
                        const float* lr_mult, const float* decay_mult, float rate,
                        float momentum, float weight_decay, int l1) {
  if (N > 1) cos_oracle_all_gather(N, P, data);        /* on_start :214-216 */
  if (N > 1) {                                         /* on_gradients_ready */
    for (int r = 0; r < N; ++r) cos_oracle_scale(N, P, diff[r]);
    cos_oracle_reduce_scatter(N, P, diff);
  }
  /* N == 1: LocalCaffeNet in CPU mode installs no sync object at all
   * (CaffeNet.cpp:206-216, syncs_.resize(0)), so no scale is applied. */
  for (int r = 0; r < N; ++r) {                        /* ApplyUpdate :253 */
    cos_oracle_apply_update_ex(0, P, data[r], diff[r], hist[r], nblobs, counts,
                               lr_mult, decay_mult, rate, momentum, weight_decay, l1);
  }
}

void cos_oracle_step(int N, uint64_t P, float* const* data, float* const* diff,
                     float* const* hist, int nblobs, const int64_t* counts,
                     const float* lr_mult, const float* decay_mult, float rate,
                     float momentum, float weight_decay) {
  cos_oracle_step_ex(N, P, data, diff, hist, nblobs, counts, lr_mult, decay_mult,
                     rate, momentum, weight_decay, 0);
}

void cos_oracle_round_bf16(uint64_t n, float* x) {
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, &x[i], 4);
    if ((u & 0x7f800000u) == 0x7f800000u && (u & 0x007fffffu)) {
      u |= 0x00400000u; /* quiet NaN, keep payload top bits */
      u &= 0xffff0000u;
    } else {
      uint32_t lsb = (u >> 16) & 1u;
      u += 0x7fffu + lsb;
      u &= 0xffff0000u;
    }
    memcpy(&x[i], &u, 4);
  }
}

static inline uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

void cos_oracle_fill(uint64_t n, float* out, uint64_t seed, uint64_t stream,
                     float amp) {
  const uint64_t key = mix64(seed * 0x9e3779b97f4a7c15ULL + stream);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t z = mix64(key + i * 0x9e3779b97f4a7c15ULL);
    int32_t v = (int32_t)(z >> 40) - (1 << 23); /* 24 bits, centred */
    out[i] = amp * ((float)v * (1.0f / 8388608.0f));
  }
}

uint64_t cos_oracle_hash(const void* p, uint64_t nbytes) {
  const unsigned char* b = (const unsigned char*)p;
  uint64_t h = 0xcbf29ce484222325ULL;
  for (uint64_t i = 0; i < nbytes; ++i) {
    h ^= b[i];
    h *= 0x100000001b3ULL;
  }
  return h;
}
