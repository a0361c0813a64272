"""ctypes binding of the CPU ORACLE (oracle/sync_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() may import this module.  The product package
(caffeonspark_b200) never does; its compute path is the CUDA library only.

Also wraps oracle/_ref/ref_sync: the reference's own socket_sync_cpu.cpp /
parallel_cpu.cpp / socket.cpp compiled verbatim (see oracle/Makefile).
"""
import ctypes
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ref_sync")

LR_POLICIES = {"fixed": 0, "step": 1, "exp": 2, "inv": 3, "multistep": 4, "poly": 5, "sigmoid": 6}

_lib = None


def build(with_ref=True):
    """Compile liboracle.so (and _ref/ref_sync when /root/reference exists)."""
    target = "all" if with_ref else "_build/liboracle.so"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(with_ref=False)
        L = ctypes.CDLL(LIB_PATH)
        u64, i32, f32 = ctypes.c_uint64, ctypes.c_int, ctypes.c_float
        pf = ctypes.POINTER(ctypes.c_float)
        ppf = ctypes.POINTER(pf)
        pi64 = ctypes.POINTER(ctypes.c_int64)
        L.cos_oracle_chunk.argtypes = [u64, i32, i32, ctypes.POINTER(u64), ctypes.POINTER(u64)]
        L.cos_oracle_total_size.argtypes = [pi64, i32]
        L.cos_oracle_total_size.restype = u64
        L.cos_oracle_learning_rate.argtypes = [i32, f32, f32, f32, i32, ctypes.POINTER(i32), i32, i32, i32,
                                               ctypes.POINTER(i32)]
        L.cos_oracle_learning_rate.restype = f32
        L.cos_oracle_all_gather.argtypes = [i32, u64, ppf]
        L.cos_oracle_scale.argtypes = [i32, u64, pf]
        L.cos_oracle_reduce_scatter.argtypes = [i32, u64, ppf]
        L.cos_oracle_apply_update.argtypes = [u64, u64, pf, pf, pf, i32, pi64, pf, pf, f32, f32, f32]
        L.cos_oracle_step.argtypes = [i32, u64, ppf, ppf, ppf, i32, pi64, pf, pf, f32, f32, f32]
        L.cos_oracle_apply_update_ex.argtypes = [u64, u64, pf, pf, pf, i32, pi64, pf, pf, f32, f32, f32, i32]
        L.cos_oracle_step_ex.argtypes = [i32, u64, ppf, ppf, ppf, i32, pi64, pf, pf, f32, f32, f32, i32]
        L.cos_oracle_round_bf16.argtypes = [u64, pf]
        L.cos_oracle_fill.argtypes = [u64, pf, u64, u64, f32]
        L.cos_oracle_hash.argtypes = [ctypes.c_void_p, u64]
        L.cos_oracle_hash.restype = u64
        _lib = L
    return _lib


def _pf(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ppf(arrs):
    arr_t = ctypes.POINTER(ctypes.c_float) * len(arrs)
    return arr_t(*[_pf(a) for a in arrs])


def chunk(P, N, peer):
    o, s = ctypes.c_uint64(), ctypes.c_uint64()
    lib().cos_oracle_chunk(P, N, peer, ctypes.byref(o), ctypes.byref(s))
    return o.value, s.value


def total_size(counts):
    c = np.asarray(counts, dtype=np.int64)
    return lib().cos_oracle_total_size(c.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(c))


class LrState:
    """current_step_ of SGDSolver (sgd_solver.cpp:33,43-46)."""

    def __init__(self):
        self.current_step = ctypes.c_int(0)


def learning_rate(policy, base_lr, gamma=0.0, power=0.0, stepsize=1, stepvalues=(), max_iter=1, it=0, state=None):
    state = state or LrState()
    sv = (ctypes.c_int * max(1, len(stepvalues)))(*stepvalues)
    return float(np.float32(lib().cos_oracle_learning_rate(
        LR_POLICIES[policy], base_lr, gamma, power, int(stepsize), sv, len(stepvalues), int(max_iter), int(it),
        ctypes.byref(state.current_step))))


def fill(n, seed, stream, amp):
    out = np.empty(n, dtype=np.float32)
    lib().cos_oracle_fill(n, _pf(out), seed, stream, amp)
    return out


def fill_numpy(n, seed, stream, amp):
    """Pure-numpy twin of cos_oracle_fill (cross-checks the C generator)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def mix(z):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
        return z ^ (z >> np.uint64(31))

    with np.errstate(over="ignore"):
        key = mix(np.uint64(seed) * np.uint64(0x9e3779b97f4a7c15) + np.uint64(stream))
        i = np.arange(n, dtype=np.uint64)
        z = mix(key + i * np.uint64(0x9e3779b97f4a7c15))
    v = (z >> np.uint64(40)).astype(np.int64) - (1 << 23)
    return (np.float32(amp) * (v.astype(np.float32) * np.float32(1.0 / 8388608.0))).astype(np.float32)


def round_bf16(x):
    y = np.ascontiguousarray(x, dtype=np.float32).copy()
    lib().cos_oracle_round_bf16(y.size, _pf(y))
    return y


def hash_bytes(a):
    a = np.ascontiguousarray(a)
    return lib().cos_oracle_hash(a.ctypes.data, a.nbytes)


def all_gather(data):
    lib().cos_oracle_all_gather(len(data), data[0].size, _ppf(data))


def scale(solver_count, diff):
    lib().cos_oracle_scale(solver_count, diff.size, _pf(diff))


def reduce_scatter(diff):
    lib().cos_oracle_reduce_scatter(len(diff), diff[0].size, _ppf(diff))


def _layout(counts, lr_mult, decay_mult):
    c = np.ascontiguousarray(counts, dtype=np.int64)
    lm = np.ascontiguousarray(lr_mult, dtype=np.float32)
    dm = np.ascontiguousarray(decay_mult, dtype=np.float32)
    assert len(c) == len(lm) == len(dm)
    return c, lm, dm


def apply_update(begin, end, data, diff, hist, counts, lr_mult, decay_mult, rate, momentum, weight_decay,
                 regularization_type="L2"):
    c, lm, dm = _layout(counts, lr_mult, decay_mult)
    lib().cos_oracle_apply_update_ex(begin, end, _pf(data), _pf(diff), _pf(hist), len(c),
                                     c.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _pf(lm), _pf(dm),
                                     rate, momentum, weight_decay, int(regularization_type == "L1"))


def step(data, diff, hist, counts, lr_mult, decay_mult, rate, momentum, weight_decay, regularization_type="L2"):
    """One Solver::Step on len(data) simulated ranks (in place)."""
    c, lm, dm = _layout(counts, lr_mult, decay_mult)
    N, P = len(data), data[0].size
    lib().cos_oracle_step_ex(N, P, _ppf(data), _ppf(diff), _ppf(hist), len(c),
                             c.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _pf(lm), _pf(dm),
                             rate, momentum, weight_decay, int(regularization_type == "L1"))


class Simulation:
    """N simulated executors driven exactly like oracle/ref_driver.cpp drives
    the reference: same seeds/streams, same Step order.  Produces, per
    iteration, own-shard weights/history per rank, and the final consistent
    weights after a trailing on_start."""

    def __init__(self, N, counts, lr_mult=None, decay_mult=None, lr_policy="fixed", base_lr=0.01, gamma=0.1,
                 power=0.75, stepsize=1, stepvalues=(), max_iter=1000, momentum=0.9, weight_decay=0.0005,
                 seed=1, w_amp=0.05, g_amp=0.01, bf16=False, regularization_type="L2"):
        self.N = N
        self.counts = list(counts)
        self.lr_mult = list(lr_mult) if lr_mult is not None else [1.0] * len(counts)
        self.decay_mult = list(decay_mult) if decay_mult is not None else [1.0] * len(counts)
        self.P = int(sum(counts))
        self.hp = dict(lr_policy=lr_policy, base_lr=base_lr, gamma=gamma, power=power, stepsize=stepsize,
                       stepvalues=tuple(stepvalues), max_iter=max_iter, momentum=momentum,
                       weight_decay=weight_decay)
        self.seed, self.w_amp, self.g_amp, self.bf16 = seed, w_amp, g_amp, bf16
        self.regularization_type = regularization_type
        w0 = fill(self.P, seed, 0, w_amp)
        self.data = [w0.copy() for _ in range(N)]
        self.hist = [np.zeros(self.P, np.float32) for _ in range(N)]
        self.iter = 0
        self.lr_state = LrState()

    def gradient(self, rank, it):
        g = fill(self.P, self.seed, (it + 1) * 4096 + rank, self.g_amp)
        return round_bf16(g) if self.bf16 else g

    def rate(self):
        h = self.hp
        return learning_rate(h["lr_policy"], h["base_lr"], h["gamma"], h["power"], h["stepsize"], h["stepvalues"],
                             h["max_iter"], self.iter, self.lr_state)

    def step(self, grads=None):
        diff = grads if grads is not None else [self.gradient(r, self.iter) for r in range(self.N)]
        diff = [np.ascontiguousarray(g, dtype=np.float32).copy() for g in diff]
        rate = self.rate()
        step(self.data, diff, self.hist, self.counts, self.lr_mult, self.decay_mult, rate,
             self.hp["momentum"], self.hp["weight_decay"], self.regularization_type)
        self.iter += 1
        return rate

    def own(self, rank):
        o, s = chunk(self.P, self.N, rank)
        return self.data[rank][o:o + s].copy(), self.hist[rank][o:o + s].copy()

    def consistent_weights(self):
        """What every rank holds after the next on_start (= concat of owners)."""
        out = np.empty(self.P, np.float32)
        for r in range(self.N):
            o, s = chunk(self.P, self.N, r)
            out[o:o + s] = self.data[r][o:o + s]
        return out

    def consistent_history(self):
        out = np.empty(self.P, np.float32)
        for r in range(self.N):
            o, s = chunk(self.P, self.N, r)
            out[o:o + s] = self.hist[r][o:o + s]
        return out


def ref_available():
    return os.path.exists(REF_BIN)


def _ref_args(N, counts, lr_mult, decay_mult, hp, iters, seed, w_amp, g_amp, bf16):
    lm = lr_mult if lr_mult is not None else [1.0] * len(counts)
    dm = decay_mult if decay_mult is not None else [1.0] * len(counts)
    args = [f"--ranks={N}", "--counts=" + ",".join(str(int(c)) for c in counts),
            "--lr_mult=" + ",".join(repr(float(x)) for x in lm),
            "--decay_mult=" + ",".join(repr(float(x)) for x in dm),
            f"--iters={iters}", f"--seed={seed}", f"--w_amp={w_amp!r}", f"--g_amp={g_amp!r}",
            f"--bf16={int(bool(bf16))}"]
    for k in ("lr_policy", "base_lr", "gamma", "power", "stepsize", "max_iter", "momentum", "weight_decay"):
        if k in hp:
            args.append(f"--{k}={hp[k]}")
    if hp.get("stepvalues"):
        args.append("--stepvalue=" + ",".join(str(int(v)) for v in hp["stepvalues"]))
    return args


def run_ref_dump(N, counts, lr_mult=None, decay_mult=None, iters=3, seed=1, w_amp=0.05, g_amp=0.01, bf16=False,
                 timeout=600, **hp):
    """Run the reference's own code (N processes over loopback TCP) and return
    (own_w[t][rank], own_h[t][rank], final_full_weights[rank])."""
    P = int(sum(counts))
    with tempfile.TemporaryDirectory(prefix="cos_ref_") as d:
        args = [REF_BIN] + _ref_args(N, counts, lr_mult, decay_mult, hp, iters, seed, w_amp, g_amp, bf16)
        subprocess.run(args + ["--dump=1", f"--dir={d}"], check=True, timeout=timeout,
                       stdout=subprocess.DEVNULL)
        own_w = [[None] * N for _ in range(iters)]
        own_h = [[None] * N for _ in range(iters)]
        final = []
        for r in range(N):
            raw = np.fromfile(os.path.join(d, f"rank_{r}.bin"), dtype=np.float32)
            _, s = chunk(max(P, 1), N, r)
            assert raw.size == 2 * s * iters + P, (raw.size, s, iters, P)
            for t in range(iters):
                own_w[t][r] = raw[(2 * t) * s:(2 * t + 1) * s].copy()
                own_h[t][r] = raw[(2 * t + 1) * s:(2 * t + 2) * s].copy()
            final.append(raw[2 * s * iters:].copy())
        return own_w, own_h, final


def run_ref_time(N, counts, lr_mult=None, decay_mult=None, iters=5, seed=1, timeout=900, **hp):
    """Time the reference's socket sync + update; returns the JSON dict
    printed by ref_sync (ms_per_iter_median is max-over-ranks per iteration)."""
    with tempfile.TemporaryDirectory(prefix="cos_ref_") as d:
        args = [REF_BIN] + _ref_args(N, counts, lr_mult, decay_mult, hp, iters, seed, 0.05, 0.01, False)
        out = subprocess.run(args + ["--time=1", f"--dir={d}"], check=True, timeout=timeout,
                             capture_output=True, text=True).stdout
        for line in out.splitlines():
            if line.startswith("{"):
                return json.loads(line)
    raise RuntimeError("ref_sync printed no timing line")
