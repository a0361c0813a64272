"""world_size-2 gloo test (CPU) of the N>1 host path: the address exchange the
Spark driver performs (CaffeOnSpark.scala:113-154) done over torch.distributed,
then PeerAdapter connect / CTRL barrier / handle passing between two real
processes -- everything connect() does short of mapping device memory."""
import os
import socket
import sys

import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import caffeonspark_b200 as C
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ad = C.PeerAdapter(world, rank)
        # phase 1+2: every rank publishes one address per peer ("" for itself), all-gather = broadcast
        mine = [ad.address() if p != rank else "" for p in range(world)]
        table = [None] * world
        dist.all_gather_object(table, mine)
        addrs = [table[p][rank] if p != rank else "" for p in range(world)]
        assert ad.connect(addrs), "connect failed"
        for _ in range(3):
            assert ad.barrier(10000)
        # handle passing: each rank offers a memfd whose content names the rank
        fd = os.memfd_create(f"cos_rank{rank}")
        os.write(fd, f"arena-of-rank-{rank}".encode())
        ad.offer_fd("arena", fd, f"meta{rank}".encode())
        peer = 1 - rank
        got, meta = ad.fetch_fd(peer, "arena", timeout_ms=10000)
        os.lseek(got, 0, os.SEEK_SET)
        content = os.read(got, 64).decode()
        os.close(got)
        assert ad.barrier(10000)
        # shard table agreement across ranks (integer state must be identical)
        shards = [C.chunk(431080, world, r) for r in range(world)]
        all_shards = [None] * world
        dist.all_gather_object(all_shards, shards)
        assert all_shards[0] == all_shards[1]
        q.put((rank, content, meta.rstrip(b"\0").decode()))
        ad.close()
    finally:
        dist.destroy_process_group()


def test_two_process_adapter_over_gloo(cos):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "arena-of-rank-1", "meta1"), (1, "arena-of-rank-0", "meta0")]
