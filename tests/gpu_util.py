"""Helpers shared by the -m gpu parity tests: build nets through the C ABI,
load the oracle's seeded synthetic tensors, run N in-process ranks."""
import concurrent.futures as cf

import numpy as np
import torch


def to_dev(view, host):
    view.copy_(torch.from_numpy(np.ascontiguousarray(host)).to(view.device))


def to_host(view):
    return view.detach().cpu().numpy().copy()


class Ranks:
    """N executors inside one process (one CaffeNet each).  With devices=None all
    ranks share cuda:0 -- the kernels then must be co-resident, so the grid is
    kept small -- otherwise rank r uses devices[r]."""

    def __init__(self, cos, desc, N, devices=None, grid=4, block=128, timeout_ms=8000, **options):
        self.cos, self.desc, self.N = cos, desc, N
        self.nets = []
        for r in range(N):
            dev = 0 if devices is None else devices[r]
            net = cos.CaffeNet(desc, "", "", 1, N, r, True, cos.CaffeNet.SOCKET if N > 1 else cos.CaffeNet.NONE,
                               dev - 1, 0)
            assert net.deviceID(0) == dev
            net.set_option("barrier_timeout_ms", timeout_ms)
            if devices is None and N > 1:
                net.set_option("grid", grid)
                net.set_option("block", block)
                # launch the first on_start() from ONE thread (below) like every later step: kernels of
                # in-process ranks launched from racing threads were not reliably co-scheduled
                net.set_option("initial_gather", 0)
            for k, v in options.items():
                net.set_option(k, v)
            self.nets.append(net)

    def set_weights(self, per_rank_weights):
        for net, w in zip(self.nets, per_rank_weights):
            to_dev(net.data(), w)
        torch.cuda.synchronize()

    def connect(self):
        if self.N == 1:
            assert self.nets[0].connect(self.nets[0].localAddresses())
            return
        table = [n.localAddresses() for n in self.nets]

        def go(r):
            addrs = [table[p][r] if p != r else "" for p in range(self.N)]
            ok = self.nets[r].connect(addrs)
            return ok, self.nets[r].last_error()

        with cf.ThreadPoolExecutor(self.N) as ex:
            res = list(ex.map(go, range(self.N)))
        assert all(ok for ok, _ in res), res
        if self.nets[0].get_option("initial_gather") == 0:
            for net in self.nets:
                assert net.all_gather_weights(0), net.last_error()
            for net in self.nets:
                assert net.synchronize(), net.last_error()

    def step(self, grads):
        for net, g in zip(self.nets, grads):
            to_dev(net.diff(), g)
        torch.cuda.synchronize()
        for net in self.nets:            # async launches on each net's own stream
            assert net.sync_step(0), net.last_error()
        for net in self.nets:
            assert net.synchronize(), net.last_error()

    def weights(self, r):
        return to_host(self.nets[r].data())

    def history(self, r):
        return to_host(self.nets[r].history())

    def diff(self, r):
        return to_host(self.nets[r].diff())

    def close(self):
        with cf.ThreadPoolExecutor(max(1, self.N)) as ex:
            list(ex.map(lambda n: n.deallocate(), self.nets))
        self.nets = []


def assert_bits_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
        bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
        raise AssertionError(f"{what}: {bad.size} of {a.size} elements differ bitwise; first at {bad[0]}: "
                             f"{a[bad[0]]!r} vs {b[bad[0]]!r}")
