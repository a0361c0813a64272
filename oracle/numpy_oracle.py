"""Second, independent restatement of SURVEY.md Appendix A in numpy (TEST INFRASTRUCTURE).

oracle/sync_oracle.c is the oracle the GPU tests use; this file restates the same
iteration in vectorised float32 numpy, written separately from the C code, so the
two can be cross-checked on random cases (tests/test_oracle.py).  numpy float32
arithmetic rounds every operation individually, like the un-fused C loops.
"""
import numpy as np

F = np.float32


def shard(P, N, p):  # socket_sync_cpu.cpp:46-54
    return (p * P) // N, ((p + 1) * P) // N


def step(data, grads, hist, counts, lr_mult, decay_mult, rate, momentum, weight_decay, regularization_type="L2"):
    """One Solver::Step on len(data) ranks; returns new (data, hist) lists (inputs untouched)."""
    N, P = len(data), data[0].size
    data = [d.astype(F).copy() for d in data]
    hist = [h.astype(F).copy() for h in hist]
    g = [x.astype(F).copy() for x in grads]
    if N > 1:
        # 1. on_start: every rank gets the owners' weight shards
        full = np.empty(P, F)
        for p in range(N):
            lo, hi = shard(P, N, p)
            full[lo:hi] = data[p][lo:hi]
        data = [full.copy() for _ in range(N)]
        # 3. scale the whole gradient by (float)(1.0 / N), BEFORE the sum
        inv = F(1.0 / N)
        g = [inv * x for x in g]
        # 4. owner r: s = own; then + peers r+1, r+2, ... in that order (recv + s)
        red = [x.copy() for x in g]
        for r in range(N):
            lo, hi = shard(P, N, r)
            s = g[r][lo:hi].copy()
            for j in range(1, N):
                s = g[(r + j) % N][lo:hi] + s
            red[r][lo:hi] = s
        g = red
    # 5. ApplyUpdate on the FULL buffer of every rank
    lr = np.repeat((F(rate) * np.asarray(lr_mult, F)).astype(F), counts)
    ld = np.repeat((F(weight_decay) * np.asarray(decay_mult, F)).astype(F), counts)
    m = F(momentum)
    for r in range(N):
        x = data[r] if regularization_type == "L2" else np.sign(data[r]).astype(F)  # L1: caffe_cpu_sign first
        gr = np.where(ld != 0, ld * x + g[r], g[r]).astype(F)         # Regularize (saxpy)
        h = (m * hist[r]).astype(F)                                   # axpby = scal ...
        h = (lr * gr + h).astype(F)                                   # ... then axpy
        hist[r] = h
        data[r] = (F(-1.0) * h + data[r]).astype(F)                   # Blob::Update
    return data, hist
