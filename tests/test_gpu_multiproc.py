"""-m gpu: one PROCESS per GPU (the deployment shape: one Spark executor JVM
per GPU), addresses exchanged through torch.distributed the way the Spark
driver would.  Exercises the cross-process VMM file-descriptor path and real
NVLink peer traffic.  Needs >= 2 GPUs; skipped on the single-GPU box (the
in-process tests of test_gpu_multi.py cover the kernels there)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, gpu_count

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, algo, bf16, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import caffeonspark_b200 as C
    from caffeonspark_b200.harness import Cluster
    from oracle import oracle as O
    from gpu_util import to_dev, to_host
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        counts, lm, dm = [500, 20, 25000, 50, 400000, 500, 5000, 10], [1, 2] * 4, [1, 1] * 4
        hp = dict(lr_policy="inv", base_lr=0.01, gamma=0.0001, power=0.75, momentum=0.9, weight_decay=0.0005)
        desc = C.SolverDesc(counts, lm, dm, grad_dtype="bf16" if bf16 else "fp32", **hp)
        sim = O.Simulation(world, counts, lm, dm, seed=77, bf16=bf16, **hp)
        cl = Cluster(desc, rank=rank, world=world, device=rank)
        net = cl.net
        net.set_option("algo", algo)
        net.set_option("barrier_timeout_ms", 15000)
        to_dev(net.data(), sim.data[rank])
        torch.cuda.synchronize()
        cl.start()
        ok = True
        for t in range(3):
            g = O.fill(sim.P, 77, (t + 1) * 4096 + rank, 0.01)
            sim.step()
            to_dev(net.diff(), g)
            torch.cuda.synchronize()
            assert net.sync_step(0), net.last_error()
            assert net.synchronize(), net.last_error()
            ok = ok and np.array_equal(to_host(net.data()).view(np.uint32), sim.consistent_weights().view(np.uint32))
            o, s = net.shard()
            ok = ok and np.array_equal(to_host(net.history())[o:o + s].view(np.uint32),
                                       sim.consistent_history()[o:o + s].view(np.uint32))
        q.put((rank, bool(ok), int(net.get_option("transport")), float(net.last_kernel_ms())))
        assert net.sync()
        net.deallocate()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(gpu_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("algo,bf16", [(1, False), (2, False), (1, True)])
def test_one_process_per_gpu_bit_exact(cos, oracle, algo, bf16):
    import torch.multiprocessing as mp
    world = min(gpu_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, algo, bf16, q)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(ok for _, ok, _, _ in res), res
