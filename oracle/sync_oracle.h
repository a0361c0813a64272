/*
 * sync_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of CaffeOnSpark's inter-executor gradient-sync + SGD
 * update arithmetic (the SocketSyncCPU path).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it;
 * the product library (caffeonspark_b200/csrc) never links or calls it.
 *
 * Parity status: the exchange part (chunk / all-gather / 1-over-N scale /
 * ordered reduce) is PINNED bit-exactly against the reference's own
 * socket_sync_cpu.cpp + parallel_cpu.cpp + socket.cpp compiled verbatim
 * (oracle/_ref, see oracle/Makefile and tests/test_oracle_vs_ref.py).  The
 * SGD update (Regularize / ComputeUpdateValue / Blob::Update) is a
 * restatement of sgd_solver.cpp on top of an *un-pinned third-party BLAS*
 * (ATLAS/OpenBLAS/MKL, not vendored in /root/reference); it is pinned only
 * to the analytic least-squares solver test constants of
 * caffe-public/src/caffe/test/test_gradient_based_solver.cpp (1e-2 relative
 * there).  We fix the un-fused (no FMA) evaluation order written down in
 * caffe/util/mkl_alternate.hpp:83-88.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef COS_SYNC_ORACLE_H_
#define COS_SYNC_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lr_policy ids (sgd_solver.cpp:27-63) */
enum {
  COS_ORACLE_LR_FIXED = 0,
  COS_ORACLE_LR_STEP = 1,
  COS_ORACLE_LR_EXP = 2,
  COS_ORACLE_LR_INV = 3,
  COS_ORACLE_LR_MULTISTEP = 4,
  COS_ORACLE_LR_POLY = 5,
  COS_ORACLE_LR_SIGMOID = 6
};

/* caffe-distri/src/main/cpp/util/socket_sync_cpu.cpp:46-54 */
void cos_oracle_chunk(uint64_t P, int N, int peer, uint64_t* offs,
                      uint64_t* size);

/* caffe-public/src/caffe/parallel.cpp:60-74 (total_size rule: max(1, sum)) */
uint64_t cos_oracle_total_size(const int64_t* counts, int nblobs);

/* caffe-public/src/caffe/solvers/sgd_solver.cpp:27-63.  *current_step is the
 * solver's current_step_ member (read+written by step / multistep). */
float cos_oracle_learning_rate(int policy, float base_lr, float gamma,
                               float power, int stepsize,
                               const int* stepvalues, int nstepvalues,
                               int max_iter, int iter, int* current_step);

/* socket_sync_cpu.cpp:102-105,135-163: all-gather of owned weight shards. */
void cos_oracle_all_gather(int N, uint64_t P, float* const* data);

/* parallel_cpu.cpp:120-122: diff *= Dtype(1.0 / solver_count), whole buffer */
void cos_oracle_scale(int solver_count, uint64_t P, float* diff);

/* socket_sync_cpu.cpp:108-133: on rank r, for p=r+1..r+N-1 (mod N), in that
 * order: diff_r[own] = diff_p[own] + diff_r[own].  All ranks processed
 * "simultaneously": sends snapshot the already-scaled buffers first. */
void cos_oracle_reduce_scatter(int N, uint64_t P, float* const* diff);

/* sgd_solver.cpp:102-116 (ApplyUpdate) = Regularize :145-204 (L2; _ex: L2 or L1),
 * ComputeUpdateValue :213-243, Net::Update -> Blob::Update blob.cpp:162-179,
 * on the element range [begin,end) of the flat buffer.  The reference runs the
 * full range [0,P) on every rank; only the owned shard is meaningful. */
void cos_oracle_apply_update(uint64_t begin, uint64_t end, float* data,
                             float* diff, float* hist, int nblobs,
                             const int64_t* counts, const float* lr_mult,
                             const float* decay_mult, float rate,
                             float momentum, float weight_decay);

/* Same with regularization_type selectable: l1 = 0 -> "L2" (:155-160), 1 -> "L1" (:161-168). */
void cos_oracle_apply_update_ex(uint64_t begin, uint64_t end, float* data,
                                float* diff, float* hist, int nblobs,
                                const int64_t* counts, const float* lr_mult,
                                const float* decay_mult, float rate,
                                float momentum, float weight_decay, int l1);

/* One complete Solver::Step (solver.cpp:194-273) for N simulated ranks, with
 * the local gradients supplied by the caller in diff[r] (step 2 of SURVEY
 * App. A).  Performs on_start (all-gather), on_gradients_ready (scale +
 * reduce-scatter), ApplyUpdate on the FULL buffer as the reference does.
 * After it returns, data[r][own shard of r] and hist[r][own shard of r] hold
 * the reference's values; other shards are the reference's stale values. */
void cos_oracle_step(int N, uint64_t P, float* const* data, float* const* diff,
                     float* const* hist, int nblobs, const int64_t* counts,
                     const float* lr_mult, const float* decay_mult, float rate,
                     float momentum, float weight_decay);

void cos_oracle_step_ex(int N, uint64_t P, float* const* data, float* const* diff,
                        float* const* hist, int nblobs, const int64_t* counts,
                        const float* lr_mult, const float* decay_mult, float rate,
                        float momentum, float weight_decay, int l1);

/* bf16 round-to-nearest-even of an fp32 gradient buffer, in place (config 3:
 * "reference arithmetic applied to bf16-rounded gradient inputs"). */
void cos_oracle_round_bf16(uint64_t n, float* x);

/* Deterministic synthetic tensors shared by oracle driver, tests and bench:
 * value(i) = amp * (int24(mix64(seed, stream, i)) - 2^23) / 2^23, exactly
 * representable in fp32, identical in C / numpy / CUDA. */
void cos_oracle_fill(uint64_t n, float* out, uint64_t seed, uint64_t stream,
                     float amp);

/* FNV-1a 64-bit over raw bytes (checksum of buffers for cheap equality). */
uint64_t cos_oracle_hash(const void* p, uint64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
