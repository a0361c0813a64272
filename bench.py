#!/usr/bin/env python
"""bench.py -- images/sec of synchronous data-parallel SGD with the B200-native
gradient-sync library, next to the reference's CPU socket path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lenet|cifar10_quick|caffenet]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the reference's own CPU sync path (oracle/_ref)
  python bench.py --sweep --gpus N         # config 5: all-reduce message-size sweep

A "step" is one Solver::Step: Net::ForwardBackward (PyTorch/cuDNN harness, NOT
part of the product) followed by the hot path -- ONE launch of the fused
scale + reduce-scatter + SGD/momentum + weight all-gather kernel.
 value  : inputs resident in HBM, device-timed (CUDA events per step, summed;
          max over ranks), L2 flushed between steps outside the timed events.
 e2e    : the reference-facing call CaffeNet.train(0, FloatBlob[]) with pinned
          HOST blobs: H2D of the batch and D2H of the loss inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# Rank 0 prints exactly ONE JSON line on stdout.  Native libraries write to file descriptor 1 behind
# Python's back (NCCL prints its "NCCL version ..." banner there at every debug level >= VERSION), so the
# real stdout is set aside and fd 1 points at stderr for the whole run; emit() writes the line to the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


NVLINK_MEASURED_GBS = 770.0  # B200_PROFILING.md: measured peer copy per direction (900 nominal)
HBM_FALLBACK_GBS = 6650.0    # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


def env_int(name, default):
    return int(os.environ.get(name, default))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cuda_time_steps(torch, dist, world, steps, warmup, step_fn, flush_fn):
    """W warm-up steps, then K steps each bracketed by CUDA events on the
    current stream (L2 flush between steps, outside the events); barrier +
    synchronize on both sides; returns max-over-ranks total ms."""
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush_fn()
        a.record()
        step_fn()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
    return total


def algorithmic_bytes(P, world, mode, zero_diff, bf16):
    """DESIGN.md section 'algorithmic bytes': per launch, per GPU.
    -> (hbm_bytes, nvlink_bytes_per_direction)"""
    z = 4 * P if zero_diff else 0
    if world == 1:
        return 20 * P + z, 0                      # read g,w,h ; write w,h (+ zero g)
    bg = 2 if bf16 else 4
    f = (world - 1) / world
    cast = (4 * P + 2 * P) if bf16 else 0         # phase 0: read fp32, write bf16 wire
    if mode == 2:                                 # one-shot: read all peers' full gradient, update everything
        return cast + bg * P + 20 * P + z, bg * P * (world - 1)
    hbm = cast + bg * P + 4 * P * f + 12 * P / world + 4 * P / world + z  # serve grads, land pushes, own shard
    return hbm, (bg + 4) * P * f                  # pull grads + push weights


def run_ours(args):
    import torch
    import torch.distributed as dist
    import caffeonspark_b200 as C
    from caffeonspark_b200 import harness, nets

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.backends.cudnn.benchmark = not os.environ.get("COS_BENCH_NO_AUTOTUNE")  # off under ncu launch lists
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True

    name = args.workload
    bf16 = args.grad_dtype == "bf16"
    desc = nets.solver_desc(name, grad_dtype=args.grad_dtype)
    batch = nets.NETS[name]["batch"]
    cl = harness.Cluster(desc, rank=rank, world=world, device=local)
    net = cl.net
    if args.algo:
        net.set_option("algo", args.algo)
    net.set_option("kernel", args.kernel)
    net.set_option("nvls", int(args.nvls))
    net.set_option("timing", 1)
    net.set_option("barrier_timeout_ms", 60000)
    prod = harness.make_producer(name, net)
    cl.start()
    P = net.param_count()
    mode = int(net.get_option("resolved_algo"))
    zero = int(net.get_option("zero_diff"))

    # synthetic batch: uniform[0,1) images, random labels; different per rank
    g = torch.Generator().manual_seed(1 + rank)
    c, h, w = nets.NETS[name]["input"]
    x_host = torch.rand((batch, c, h, w), generator=g).pin_memory()
    y_host = torch.randint(0, nets.NETS[name]["classes"], (batch, 1, 1, 1), generator=g).float().pin_memory()
    x_dev, y_dev = x_host.cuda(non_blocking=True), y_host.cuda().view(-1).long()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    working_set = 3 * 4 * P
    need_flush = working_set < (256 << 20)

    def flush():
        if need_flush:
            flush_buf.zero_()

    # Everything below runs on ONE explicit stream.  (torch's default stream has the handle 0, which the
    # C ABI reads as "use the net's own stream": forward/backward and the sync launch would then sit on
    # two unordered streams and the sync kernel would fall outside the timed events.)
    work = torch.cuda.Stream()
    work.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(work)
    assert torch.cuda.current_stream().cuda_stream != 0

    # Net::ForwardBackward, captured in a CUDA graph when possible (launch-bound for the small nets)
    graph = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    prod.forward_backward(x_dev, y_dev)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss_static = prod.forward_backward(x_dev, y_dev)
            net.diff().zero_()
        except Exception as e:  # eager fallback for the producer only (not the product path)
            graph = None
            if rank == 0:
                print(f"[bench] CUDA-graph capture of the gradient producer failed ({e}); eager", file=sys.stderr)
            net.diff().zero_()
    torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            prod.forward_backward(x_dev, y_dev)
        if not net.sync_step(0, torch.cuda.current_stream().cuda_stream):
            raise RuntimeError(net.last_error())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_end = time.perf_counter() + 0.4  # nvidia-smi needs a moment to emit its first rows: same load meanwhile
    spin = torch.tensor([1], device="cuda")
    while True:                        # collective: every rank runs the same number of extra steps
        for _ in range(10):
            step()
        spin[0] = 1 if time.perf_counter() < t_end else 0
        if world > 1:
            dist.all_reduce(spin, op=dist.ReduceOp.MIN)
        if int(spin.item()) == 0:
            break
    torch.cuda.synchronize()
    launches0 = net.launch_count()
    total_ms = cuda_time_steps(torch, dist, world, args.steps, args.warmup, step, flush)
    launches = net.launch_count() - launches0 - args.warmup
    if not net.synchronize():
        raise RuntimeError(net.last_error())

    # the fused kernel alone: library-side CUDA events around the launch, on the launching stream
    kms = []
    for _ in range(max(3, min(args.steps, 30))):
        flush()
        if graph is not None:
            graph.replay()
        else:
            prod.forward_backward(x_dev, y_dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        net.sync_step(0, torch.cuda.current_stream().cuda_stream)
        kms.append(net.last_kernel_ms())
    k_ms = sorted(kms)[len(kms) // 2]
    if world > 1:
        t = torch.tensor([k_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_ms = float(t.item())
    # forward/backward alone (reported so the split is visible)
    fb_evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay() if graph is not None else prod.forward_backward(x_dev, y_dev)
        b.record()
        fb_evs.append((a, b))
    torch.cuda.synchronize()
    fb_ms = sorted(a.elapsed_time(b) for a, b in fb_evs)[5]
    net.diff().zero_()
    torch.cuda.synchronize()

    # end to end through the reference-facing API: train(solver_index, host blobs)
    for _ in range(args.warmup):
        assert net.train(0, [x_host, y_host]), net.last_error()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_launch0 = net.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if not net.train(0, [x_host, y_host]):
            raise RuntimeError(net.last_error())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = net.launch_count() - e2e_launch0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    last_loss = net.last_loss()
    clocks = sampler.stop() if rank == 0 else None

    peaks, peak_kind = measured_peaks()
    hbm_b, nvl_b = algorithmic_bytes(P, world, mode, zero, bf16)
    if world == 1:
        bound, alg, peak = "hbm", hbm_b, float(peaks["hbm_gbs"])
    else:
        bound, alg, peak = "nvlink", nvl_b, NVLINK_MEASURED_GBS
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{name}_n{world}")
    except Exception:
        pass

    out = {
        "metric": "images/sec", "value": world * batch * args.steps / (total_ms * 1e-3), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if not bf16 else "f32 (bf16 gradient wire)", "data": "synthetic",
        "config": {"workload": f"{name} (batch {batch}/device, P={P} fp32 params), synchronous SGD: "
                               f"PyTorch forward/backward + fused sync kernel",
                   "global_batch": world * batch, "parallelism": f"dp{world}",
                   "algo": {0: "local", 1: "two_shot", 2: "one_shot"}[mode], "grad_dtype": args.grad_dtype,
                   "kernel": KERNEL_NAMES[int(net.get_option("resolved_kernel"))],
                   "nvls": bool(net.get_option("nvls_active")),
                   "producer": "cuda_graph" if graph is not None else "eager",
                   "l2": "flushed between steps (256 MiB write outside the timed events)" if need_flush
                         else f"working set {working_set >> 20} MiB > 126 MiB L2"},
        "clocks": clocks,
        "e2e": {"value": world * batch * args.steps / e2e_s, "unit": "images/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4), "d2h_bytes_per_step": 4,
                "timing": "host wall clock around K train() calls (each returns after the loss D2H), max over ranks",
                "last_loss": last_loss},
        "gpu_launches": int(launches), "e2e_gpu_launches": int(e2e_launches),
        "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": {0: "fused_sync_sgd_kernel", 1: "fused_sync_sgd_tma_kernel",
                                2: "fused_sync_sgd_push_kernel", 3: "fused_sync_sgd_nvls_kernel"}[
                         int(net.get_option("resolved_kernel"))], "kernel_ms": k_ms,
                     "algorithmic_bytes": alg, "peak_source": (peak_kind + " MEASURED_PEAKS.json hbm_gbs") if
                     world == 1 else "B200_PROFILING.md measured peer copy per direction"},
        "split_ms": {"forward_backward": fb_ms, "fused_sync_kernel": k_ms},
    }
    if world > 1:
        out["bus_gbs"] = 4 * P * 2 * (world - 1) / world / (k_ms * 1e-3) / 1e9
        # the library route the fused kernel replaces: NCCL all-reduce of the fp32 gradient (the SGD update
        # would still need a separate >= 20P-byte elementwise pass on top of this)
        buf = torch.zeros(P, device="cuda")
        nc = []
        for i in range(13):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            a.record()
            dist.all_reduce(buf)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                nc.append(a.elapsed_time(b))
        t = torch.tensor([sorted(nc)[len(nc) // 2]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["nccl_allreduce_only_ms"] = float(t.item())
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(name, desc, batch, 1, fb_ms)
    if rank == 0 and args.kernels:
        out["kernel_rooflines"] = kernel_rooflines(C, nets, peaks, world, args.kernel)
    net.deallocate()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def kernel_rooflines(C, nets, peaks, world, kernel=-1):
    """Fused-kernel-only timing of the three BASELINE layouts at N=1 (HBM roofline)."""
    import torch
    res = {}
    if world != 1:
        return res
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in ("lenet", "cifar10_quick", "caffenet"):
        desc = nets.solver_desc(name)
        net = C.CaffeNet(desc, "", "", 1, 1, 0, True, 0, torch.cuda.current_device() - 1, 0)
        net.connect([])
        net.set_option("kernel", kernel)
        net.set_option("timing", 1)
        P = net.param_count()
        net.diff().normal_(0, 0.01)
        ms = []
        for i in range(25):
            flush_buf.zero_()
            torch.cuda.synchronize()
            net.sync_step(0)
            net.synchronize()
            if i >= 5:
                ms.append(net.last_kernel_ms())
        k = sorted(ms)[len(ms) // 2]
        alg = 24 * P
        res[name] = {"P": P, "kernel_ms": k, "achieved_gbs": alg / (k * 1e-3) / 1e9,
                     "frac_of_hbm_peak": alg / (k * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), "algorithmic_bytes": alg}
        net.deallocate()
    return res


def cpu_baseline(name, desc, batch, world, fb_ms):
    """The reference's CPU sync path on this box's host cores, bounded sample."""
    from oracle import oracle as O
    P = desc.param_count
    cores = os.cpu_count()
    # ~10-20 s of CPU work: measured rates are ~0.4 GB/s (socket path) / ~2 GB/s (local update)
    if O.ref_available():
        est_ms = max(0.5, 4 * P / 1e6 * (4.0 if world > 1 else 1.0) * world)
        iters = int(max(4, min(2000, 10000 / est_ms)))  # ~10 s of CPU work, bounded
        r = O.run_ref_time(world, desc.counts, desc.lr_mult, desc.decay_mult, iters=iters, **desc.hyper())
        sync_ms, kind = r["ms_per_iter_median"], "reference"
        used = world  # one solver thread per executor process (+ its receiver threads)
    else:
        import numpy as np
        sim = O.Simulation(world, desc.counts, desc.lr_mult, desc.decay_mult, **desc.hyper())
        grads = [sim.gradient(r, 0) for r in range(world)]
        iters = 5
        t0 = time.perf_counter()
        for _ in range(iters):
            sim.step(grads)
        sync_ms, kind, used = (time.perf_counter() - t0) * 1e3 / iters, "port", 1
        del np
    return {"value": world * batch / ((sync_ms + fb_ms) * 1e-3), "unit": "images/s", "cores": used,
            "host_cores": cores, "kind": kind,
            "sample": f"{iters} iterations of the reference's CPU sync+update at P={P}, N={world} "
                      f"(median {sync_ms:.3f} ms/iter, first dropped) + the same GPU forward/backward time "
                      f"({fb_ms:.3f} ms) as this run, so only the sync differs",
            "sync_ms_per_iter": sync_ms}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path
    (oracle/_ref = socket.cpp + socket_sync_cpu.cpp + parallel_cpu.cpp compiled
    verbatim, N loopback processes) + the PyTorch forward/backward time."""
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    world = max(world, args.gpus)
    import torch
    from caffeonspark_b200 import nets
    name = args.workload
    desc = nets.solver_desc(name)
    batch = nets.NETS[name]["batch"]
    fb_ms = 0.0
    if torch.cuda.is_available():
        torch.backends.cudnn.benchmark = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        mod = nets.torch_module(name).cuda()
        for prm in mod.parameters():
            prm.grad = torch.zeros_like(prm)
        c, h, w = nets.NETS[name]["input"]
        x = torch.rand((batch, c, h, w), device="cuda")
        y = torch.randint(0, nets.NETS[name]["classes"], (batch,), device="cuda")
        lossf = torch.nn.CrossEntropyLoss()

        def fb():
            lossf(mod(x), y).backward()

        # same forward/backward execution as our arm (CUDA-graph replay) so that only the sync differs
        run = fb
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fb()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fb()
            run = graph.replay
        except Exception:
            run = fb
        evs = []
        for i in range(args.warmup + min(args.steps, 20)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            if i >= args.warmup:
                evs.append((a, b))
        torch.cuda.synchronize()
        fb_ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    cb = cpu_baseline(name, desc, batch, world, fb_ms)
    step_ms = cb["sync_ms_per_iter"] + fb_ms
    out = {"impl": "reference", "metric": "images/sec", "value": cb["value"], "unit": "images/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{name} (batch {batch}/device, P={desc.param_count} fp32 params), synchronous SGD: "
                                  f"PyTorch forward/backward + reference CPU socket sync",
                      "global_batch": world * batch, "parallelism": f"dp{world}"},
           "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "host_cores", "kind", "sample")},
           "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


SWEEP_VARIANTS = {
    # name: (algo, options set BEFORE connect)
    "ldg": (1, {"kernel": 0, "nvls": 0}),
    "tma": (1, {"kernel": 1, "nvls": 0}),
    "push": (1, {"kernel": 2, "nvls": 0}),
    "push1": (1, {"kernel": 2, "nvls": 0, "push_vecs": 1}),
    "push4": (1, {"kernel": 2, "nvls": 0, "push_vecs": 4}),
    "push8": (1, {"kernel": 2, "nvls": 0, "push_vecs": 8}),
    "ldg1s": (2, {"kernel": 0, "nvls": 0}),
    "tma1s": (2, {"kernel": 1, "nvls": 0}),
    "nvls1": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 1}),
    "nvls2": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 2}),
    "nvls4": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 4}),
    "nvls8": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 8}),
    "nvls1p": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 1, "nvls_p2p": 1}),
    "nvls2p": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 2, "nvls_p2p": 1}),
    "nvls3p": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 3, "nvls_p2p": 1}),
    "nvls4p": (1, {"kernel": 3, "nvls": 1, "nvls_unroll": 4, "nvls_p2p": 1}),
    "auto": (0, {}),
}
KERNEL_NAMES = {0: "ldg_stg_vector", 1: "tma_bulk_pipeline", 2: "push_store", 3: "nvls_multimem"}


def run_sweep(args):
    """Config 5: all-reduce message sweep 64 KiB - 512 MiB (single segment), every kernel variant asked for
    with --variants, next to NCCL's all-reduce of the same message (which does not include the update)."""
    import torch
    import torch.distributed as dist
    import caffeonspark_b200 as C
    from caffeonspark_b200 import harness
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rows = []
    if args.sizes:
        sizes = [int(float(x) * (1 << 20)) // 4 * 4 for x in args.sizes.split(",")]
    else:
        sizes = [64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20, 512 << 20]
    sizes = [b for b in sizes if args.sweep_min_bytes <= b <= args.sweep_max_bytes]
    names = args.variants.split(",") if args.variants else (["ldg", "tma", "push", "auto"] if world > 1 else ["ldg", "tma"])
    if args.nvls == 1 and "nvls4" not in names:
        names.append("nvls4")
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for S in sizes:
        P = S // 4
        nccl_ms = None
        for vname in names:
            algo, opts = SWEEP_VARIANTS[vname]
            if world == 1 and (vname.startswith(("push", "nvls")) or algo == 2):
                continue
            desc = C.SolverDesc([P], lr_policy="fixed", base_lr=0.01, momentum=0.9, weight_decay=0.0005,
                                grad_dtype=args.grad_dtype)
            cl = harness.Cluster(desc, rank=rank, world=world, device=local)
            net = cl.net
            for k, v in opts.items():
                net.set_option(k, v)
            if world > 1 and algo:
                net.set_option("algo", algo)
            net.set_option("zero_diff", int(args.zero_diff))
            net.set_option("timing", 1)
            net.set_option("barrier_timeout_ms", 60000)
            cl.start()
            net.diff().normal_(0, 0.01)
            ms = []
            for i in range(args.warmup + args.steps):
                if S < (64 << 20):
                    flush_buf.zero_()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                net.sync_step(0)
                if not net.synchronize():
                    raise RuntimeError(net.last_error())
                if i >= args.warmup:
                    ms.append(net.last_kernel_ms())
            k = sorted(ms)[len(ms) // 2]
            if world > 1:
                t = torch.tensor([k], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                k = float(t.item())
            # steady state: launches queued back to back on every rank (no host sync in between), so the
            # host-side launch skew between ranks is not part of the number
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            net.set_option("timing", 0)
            st = torch.cuda.Stream()  # explicit stream: handle 0 would mean "the net's own stream"
            with torch.cuda.stream(st):
                if not net.sync_step(0, st.cuda_stream):
                    raise RuntimeError(net.last_error())
                a.record()
                for _ in range(reps):
                    if not net.sync_step(0, st.cuda_stream):
                        raise RuntimeError(net.last_error())
                b.record()
            torch.cuda.synchronize()
            if not net.synchronize():
                raise RuntimeError(net.last_error())
            piped = a.elapsed_time(b) / reps
            if world > 1:
                t = torch.tensor([piped], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                piped = float(t.item())
            row = {"bytes": S, "variant": vname,
                   "algo": {0: "local", 1: "two_shot", 2: "one_shot"}[int(net.get_option("resolved_algo"))],
                   "kernel": KERNEL_NAMES[int(net.get_option("resolved_kernel"))],
                   "nvls_active": bool(net.get_option("nvls_active")), "grad_dtype": args.grad_dtype,
                   "zero_diff": bool(args.zero_diff),
                   "kernel_ms": k, "min_ms": min(ms), "pipelined_ms": piped}
            if args.trace and int(net.get_option("resolved_kernel")) != 1:
                # where the time goes inside one launch (CTA 0, %globaltimer); stamps per kernel:
                #  ldg/nvls: start, after barrier A, after reduce+update+push, after barrier B, after zero
                #  push    : start, after scatter-push, after barrier A, after reduce+update+push, after barrier B
                net.set_option("trace", 1)
                tr = []
                for _ in range(7):
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    net.sync_step(0)
                    net.synchronize()
                    t = [net.get_option(f"trace_{i}") for i in range(5)]
                    tr.append([(t[i + 1] - t[i]) / 1e3 for i in range(4)])
                net.set_option("trace", 0)
                row["trace_us"] = [sorted(c)[len(c) // 2] for c in zip(*tr)]
            if world > 1:
                row["bus_gbs"] = S * 2 * (world - 1) / world / (k * 1e-3) / 1e9
            if world > 1 and nccl_ms is None:
                # NCCL all-reduce of the same message (the library baseline the fused kernel must beat);
                # it does NOT include the SGD update
                buf = torch.zeros(P, device="cuda")
                evs = []
                for i in range(args.warmup + args.steps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    dist.barrier()
                    a.record()
                    dist.all_reduce(buf)
                    b.record()
                    torch.cuda.synchronize()
                    if i >= args.warmup:
                        evs.append(a.elapsed_time(b))
                nccl_ms = sorted(evs)[len(evs) // 2]
                t = torch.tensor([nccl_ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                nccl_ms = float(t.item())
                del buf
            if world > 1:
                row["nccl_allreduce_ms"] = nccl_ms
            else:
                row["hbm_gbs"] = (24 if args.zero_diff else 20) * P / (k * 1e-3) / 1e9
            rows.append(row)
            if rank == 0:
                print("[sweep] %8d KiB %-7s %-18s nvls=%d  %.1f us (min %.1f, piped %.1f)  nccl %s  trace %s" % (
                    S >> 10, vname, row["kernel"], row["nvls_active"], k * 1e3, min(ms) * 1e3, piped * 1e3,
                    "%.1f us" % (nccl_ms * 1e3) if nccl_ms else "-", row.get("trace_us")), file=sys.stderr)
            net.deallocate()
            if world > 1:
                dist.barrier()
    if rank == 0:
        emit({"sweep": rows, "n_gpus": world, "unit": "ms / GB/s", "steps": args.steps})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lenet", choices=["lenet", "cifar10_quick", "caffenet"])
    ap.add_argument("--grad-dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 two-shot, 2 one-shot")
    ap.add_argument("--kernel", type=int, default=-1, help="-1 auto, 0 LDG/STG vector kernel, 1 TMA bulk-copy pipeline")
    ap.add_argument("--nvls", type=int, default=-1, help="NVSwitch multicast reduce/broadcast: -1 auto, 0 off, 1 on")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernels", action="store_true", default=True)
    ap.add_argument("--no-kernels", dest="kernels", action="store_false")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--trace", action="store_true", help="sweep: per-phase timestamps of the LDG kernel")
    ap.add_argument("--variants", default="", help="sweep: comma list of " + ",".join(SWEEP_VARIANTS))
    ap.add_argument("--sizes", default="", help="sweep: comma list of message sizes in MiB (fractions allowed)")
    ap.add_argument("--zero-diff", type=int, default=1, help="sweep: fold ClearParamDiffs into the kernel")
    ap.add_argument("--sweep-min-bytes", type=int, default=0)
    ap.add_argument("--sweep-max-bytes", type=int, default=1 << 40)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    elif args.sweep:
        run_sweep(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
