// fused_sync_sgd_tma.cu -- TMA (cp.async.bulk + mbarrier) pipelined variant of
// the fused sync kernel.  Placeholder until the pipeline lands: reports
// "not supported" so callers fall back to an explicit error, never to a
// different code path silently.
#include "fused_sync_sgd.hpp"

namespace cosb {
cudaError_t launch_fused_sync_sgd_tma(const SyncParams&, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace cosb
