"""CPU model of the index arithmetic the push / NVLS kernels use (sync_device.cuh: shard_range, vec_elem; the
512-byte aligned iteration space; the NVLS zeroing warp's round -> vector mapping; push_recv_stride), checked for the
BASELINE layouts at every world size incl. the ones no in-process GPU test can reach with real NVLS (N = 8).
This restates formulas, it does not execute CUDA: it guards the arithmetic (coverage exactly once, owner and zeroing
warp agree, slots never overflow), the GPU tests guard the code."""
import numpy as np
import pytest

LAYOUT_P = {"lenet": 431080, "cifar10_quick": 145578, "caffenet": 60965224, "ragged": 1000003, "tiny": 37}


def shard_range(P, N, s):  # sync_device.cuh:shard_range
    lo, hi = s * P // N, (s + 1) * P // N
    vlo, vhi = (lo + 3) >> 2, hi >> 2
    if vhi > vlo:
        r = dict(lo=lo, hi=hi, vec_lo=vlo, nvec=vhi - vlo, head_end=vlo << 2, tail_begin=vhi << 2)
    else:
        r = dict(lo=lo, hi=hi, vec_lo=vlo, nvec=0, head_end=hi, tail_begin=hi)
    r["vec_base"] = (lo >> 2) & ~31
    r["off"] = r["vec_lo"] - r["vec_base"]
    return r


def push_recv_stride(P, N):  # fused_sync_sgd_push.cu
    max_shard = (P + N - 1) // N
    return (max_shard + 4 * 32 + 8 + 31) // 32 * 32


@pytest.mark.parametrize("name", sorted(LAYOUT_P))
@pytest.mark.parametrize("N", [2, 3, 4, 5, 6, 7, 8, 16])
def test_aligned_iteration_space_covers_every_shard_exactly(name, N):
    P = LAYOUT_P[name]
    stride_slots = push_recv_stride(P, N)
    covered = 0
    for s in range(N):
        r = shard_range(P, N, s)
        assert 0 <= r["off"] <= 32 and r["vec_base"] % 32 == 0 and (r["vec_base"] << 2) <= r["lo"]
        # thread index a addresses vector vec_base + a; valid for off <= a < off + nvec  (vec_elem)
        a = np.arange(r["off"], r["off"] + r["nvec"], dtype=np.int64)
        vec = r["vec_base"] + a
        assert (vec == np.arange(r["vec_lo"], r["vec_lo"] + r["nvec"])).all()
        # every warp's 32 lanes (a = 32k .. 32k+31) start on a 512-byte boundary of the buffer
        assert ((r["vec_base"] + (a // 32) * 32) * 16) .__mod__(512).sum() == 0
        # the kernels' loop bounds reach the last vector
        max_nvec = (((P + N - 1) // N + 3) >> 2) + 32
        assert r["off"] + r["nvec"] <= max_nvec
        # receive-slot offsets (element i - base, base = vec_base * 4): body, head and tail fit and do not overlap
        base = r["vec_base"] << 2
        body = (vec << 2) - base
        if r["nvec"]:
            assert body.min() >= r["lo"] - base and body.max() + 3 < stride_slots
        edges = list(range(r["lo"], r["head_end"])) + list(range(r["tail_begin"], r["hi"]))
        assert all(0 <= e - base < stride_slots for e in edges)
        assert len(edges) + 4 * r["nvec"] == r["hi"] - r["lo"]
        if r["nvec"]:
            assert all(e - base < body.min() or e - base > body.max() + 3 for e in edges)
        covered += r["hi"] - r["lo"]
    assert covered == P


@pytest.mark.parametrize("N", [2, 4, 6, 8])
@pytest.mark.parametrize("U", [1, 4])
def test_nvls_zeroing_warp_follows_exactly_what_the_owner_consumed(N, U):
    """fused_sync_sgd_nvls.cu: 480 work threads per CTA; owner CTA b, round `it`, unroll slot u consumes the vectors
    a = b*480 + t + (it*U + u)*stride; the zeroing warp of CTA b on every rank zeroes the same a for the rounds the
    owner has published.  Both must enumerate the shard's body exactly once with the same number of rounds."""
    P, W, grid = LAYOUT_P["caffenet"], 480, 148
    stride = grid * W
    for s in (0, 1, N - 1):
        r = shard_range(P, N, s)
        iters = (r["off"] + r["nvec"] + stride * U - 1) // (stride * U)
        assert iters < (1 << 16)  # kIterBits
        seen = np.zeros(r["nvec"], np.int32)
        for b in (0, 1, 73, grid - 1):  # a few CTAs in full, the rest by count below
            owner = []
            for it in range(iters):
                for u in range(U):
                    a = b * W + np.arange(W) + (it * U + u) * stride
                    ok = (a >= r["off"]) & (a - r["off"] < r["nvec"])
                    owner.append(a[ok])
            owner = np.concatenate(owner)
            zero = []
            for it in range(iters):      # zeroing warp: rounds [done, upto), lanes stride 32 over the 480 vectors
                for u in range(U):
                    first = b * W + (it * U + u) * stride
                    for lane in range(32):
                        a = first + np.arange(lane, W, 32)
                        ok = (a >= r["off"]) & (a - r["off"] < r["nvec"])
                        zero.append(a[ok])
            zero = np.concatenate(zero)
            assert np.array_equal(np.sort(owner), np.sort(zero)) and len(np.unique(owner)) == len(owner)
            seen[owner - r["off"]] += 1
        # all CTAs together: each a in [off, off + nvec) belongs to exactly one (CTA, round, slot)
        a_all = np.arange(r["off"], r["off"] + r["nvec"], dtype=np.int64)
        cta = (a_all % stride) // W
        rnd = a_all // stride
        assert cta.max() < grid and rnd.max() < iters * U
        assert seen.max() <= 1
