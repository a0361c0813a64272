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


def _worker(rank, world, port, algo, bf16, q, shared_gpu=False, kernel=0, nvls=False, opts=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dev = 0 if shared_gpu else rank
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(dev))
    import torch
    import torch.distributed as dist
    import caffeonspark_b200 as C
    from caffeonspark_b200.harness import Cluster
    from oracle import oracle as O
    from gpu_util import to_dev, to_host
    torch.cuda.set_device(dev)
    if shared_gpu:  # NCCL refuses two ranks on one GPU; the address exchange only needs a host backend
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{dev}"))
    try:
        counts, lm, dm = ([500, 20, 25000, 50, 400000, 500, 5000, 10] if not shared_gpu else
                          [500, 20, 2500, 50, 4000, 500, 5000, 10]), [1, 2] * 4, [1, 1] * 4
        hp = dict(lr_policy="inv", base_lr=0.01, gamma=0.0001, power=0.75, momentum=0.9, weight_decay=0.0005)
        opts = dict(opts or {})
        if "_counts" in opts:  # a bigger layout (several grid-stride rounds per CTA)
            counts = opts.pop("_counts")
            lm, dm = [1, 2] * (len(counts) // 2), [1, 0] * (len(counts) // 2)
        desc = C.SolverDesc(counts, lm, dm, grad_dtype="bf16" if bf16 else "fp32", **hp)
        sim = O.Simulation(world, counts, lm, dm, seed=77, bf16=bf16, **hp)
        cl = Cluster(desc, rank=rank, world=world, device=dev)
        net = cl.net
        net.set_option("algo", algo)
        net.set_option("kernel", kernel)
        net.set_option("nvls", int(nvls))
        for k, v in (opts or {}).items():
            net.set_option(k, v)
        net.set_option("barrier_timeout_ms", 60000 if shared_gpu else 15000)
        if shared_gpu:  # contexts time-slice on the one GPU: keep the spinning grids tiny
            net.set_option("grid", 2)
            net.set_option("block", 128)
        to_dev(net.data(), sim.data[rank])
        torch.cuda.synchronize()
        cl.start()
        ok = True
        for t in range(2 if shared_gpu else 3):
            g = O.fill(sim.P, 77, (t + 1) * 4096 + rank, 0.01)
            sim.step()
            to_dev(net.diff(), g)
            torch.cuda.synchronize()
            assert net.sync_step(0), net.last_error()
            assert net.synchronize(), net.last_error()
            o, s = net.shard()
            if nvls and net.get_option("nvls_active") == 1:
                # the switch chooses the summation order: north-star tolerance, not bit equality
                ok = ok and np.allclose(to_host(net.data()), sim.consistent_weights(), rtol=1e-5, atol=1e-8)
                ok = ok and np.allclose(to_host(net.history())[o:o + s], sim.consistent_history()[o:o + s],
                                        rtol=1e-5, atol=1e-9)
            else:
                ok = ok and np.array_equal(to_host(net.data()).view(np.uint32),
                                           sim.consistent_weights().view(np.uint32))
                ok = ok and np.array_equal(to_host(net.history())[o:o + s].view(np.uint32),
                                           sim.consistent_history()[o:o + s].view(np.uint32))
            ok = ok and not bool(net.diff().any())  # ClearParamDiffs folded into the kernel (every variant)
        q.put((rank, bool(ok), int(net.get_option("nvls_active")), float(net.last_kernel_ms())))
        assert net.sync()
        net.deallocate()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(gpu_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("algo,bf16,kernel", [(1, False, 0), (2, False, 0), (1, True, 0), (1, False, 1), (2, False, 1),
                                              (1, True, 1), (1, False, 2), (1, True, 2), (1, False, 4), (1, True, 4)])
def test_one_process_per_gpu_bit_exact(cos, oracle, algo, bf16, kernel):
    import torch.multiprocessing as mp
    world = min(gpu_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, algo, bf16, q, False, kernel)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(ok for _, ok, _, _ in res), res


@pytest.mark.skipif(gpu_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("unroll,p2p", [(4, 0), (8, 0), (1, 0), (2, 1), (4, 1)])
def test_nvls_two_shot_within_tolerance(cos, oracle, unroll, p2p):
    """NVLS kernel (multimem.ld_reduce / multimem.st through the NVSwitch, optionally sharing the work with
    plain P2P vectors): the in-switch summation order is unspecified, so the bar is the north star's 1e-5
    relative.  If the platform does not expose multicast the library keeps the P2P path (then the result must
    be bit-exact) -- reported, not failed.  The P2P share exists for world sizes 2, 4 and 8."""
    import torch.multiprocessing as mp
    world = min(gpu_count(), 8)
    if p2p and world not in (2, 4, 8):
        world = 4 if world > 4 else 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    opts = {"nvls_unroll": unroll, "nvls_p2p": p2p}
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1, False, q, False, 3, True, opts)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(ok for _, ok, _, _ in res), res
    active = {a for _, _, a, _ in res}
    assert len(active) == 1, "ranks disagree on whether NVLS is active"
    print("NVLS active:", active)


@pytest.mark.skipif(gpu_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("kernel,nvls,bf16", [(3, True, False), (2, False, False), (2, False, True), (1, False, False)])
def test_multi_round_layout_one_process_per_gpu(cos, oracle, kernel, nvls, bf16):
    """12 MB layout: every CTA runs several grid-stride rounds, so the NVLS kernel's zeroing warp follows the
    owners' progress counters over many iterations (and the push / TMA kernels loop).  diff_ must end up zero and
    the weights must match the oracle (bit-exact for P2P, 1e-5 for the in-switch sum)."""
    import torch.multiprocessing as mp
    world = min(gpu_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    opts = {"_counts": [3000000, 1000, 7, 13]}
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1, bf16, q, False, kernel, nvls, opts)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(ok for _, ok, _, _ in res), res


@pytest.mark.skipif(gpu_count() < 1, reason="needs a GPU")
@pytest.mark.parametrize("world,algo,bf16,kernel", [(2, 1, False, 0), (3, 2, False, 1), (5, 1, True, 0),
                                                    (8, 1, False, 1), (8, 2, False, 0), (5, 1, False, 2),
                                                    (8, 1, True, 2), (3, 1, False, 4), (8, 1, True, 4)])
def test_processes_sharing_one_gpu_bit_exact(cos, oracle, world, algo, bf16, kernel):
    """Several executor PROCESSES on the single test GPU: the cross-process
    descriptor-passing / VMM import path and the device-side barriers between
    different CUDA contexts (time-sliced, so slow but must be exact)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, algo, bf16, q, True, kernel)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(420) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(ok for _, ok, _, _ in res), res
