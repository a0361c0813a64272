"""Generates tests/golden/ref_sync_cases.npz by RUNNING THE REFERENCE'S OWN CODE
(oracle/_ref/ref_sync = socket.cpp + socket_sync_cpu.cpp + parallel_cpu.cpp of
/root/reference compiled verbatim, see oracle/Makefile) as N loopback
processes.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures pin oracle/sync_oracle.c (tests/test_oracle.py) on boxes where
the reference tree is absent.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CASES = [
    # name, N, counts, lr_mult, decay_mult, iters, seed, bf16, hyper
    ("n2_ragged_inv", 2, [50, 7, 33, 5], [1, 2, 1, 2], [1, 1, 1, 0], 4, 11, False,
     dict(lr_policy="inv", base_lr=0.01, gamma=0.0001, power=0.75, momentum=0.9, weight_decay=0.0005)),
    ("n3_odd_fixed", 3, [101, 3, 1, 64], [1, 2, 1, 2], [1, 0, 1, 0], 3, 12, False,
     dict(lr_policy="fixed", base_lr=0.001, momentum=0.9, weight_decay=0.004)),
    ("n4_step", 4, [1023], [1], [1], 5, 13, False,
     dict(lr_policy="step", base_lr=0.01, gamma=0.1, stepsize=2, momentum=0.9, weight_decay=0.0005)),
    ("n8_tiny", 8, [5, 2], [1, 2], [1, 1], 3, 14, False,
     dict(lr_policy="fixed", base_lr=0.05, momentum=0.5, weight_decay=0.0)),
    ("n2_bf16", 2, [130, 9], [1, 2], [1, 1], 3, 15, True,
     dict(lr_policy="fixed", base_lr=0.001, momentum=0.9, weight_decay=0.004)),
    ("n4_lenet_head", 4, [500, 20, 2500, 50], [1, 2, 1, 2], [1, 1, 1, 1], 3, 16, False,
     dict(lr_policy="inv", base_lr=0.01, gamma=0.0001, power=0.75, momentum=0.9, weight_decay=0.0005)),
    ("n5_multistep", 5, [333, 7, 64], [1, 2, 1], [1, 0, 1], 5, 17, False,
     dict(lr_policy="multistep", base_lr=0.02, gamma=0.5, stepvalues=(2, 4), momentum=0.9, weight_decay=0.001)),
    ("n6_poly", 6, [100, 1, 1, 1, 250], [1, 2, 1, 2, 1], [1, 1, 0, 0, 1], 3, 18, False,
     dict(lr_policy="poly", base_lr=0.01, power=2.0, max_iter=10, momentum=0.5, weight_decay=0.0005)),
    ("n7_plain_sgd", 7, [97, 11], [1, 1], [0, 0], 3, 19, False,
     dict(lr_policy="exp", base_lr=0.05, gamma=0.9, momentum=0.0, weight_decay=0.0)),
    ("n3_bf16_sigmoid", 3, [64, 64, 3], [1, 2, 1], [1, 1, 1], 3, 20, True,
     dict(lr_policy="sigmoid", base_lr=0.01, gamma=-0.5, stepsize=2, momentum=0.9, weight_decay=0.004)),
]


def main():
    O.build(with_ref=True)
    assert O.ref_available(), "oracle/_ref/ref_sync missing (needs /root/reference)"
    out, meta = {}, {}
    for name, N, counts, lm, dm, iters, seed, bf16, hp in CASES:
        ow, oh, fin = O.run_ref_dump(N, counts, lm, dm, iters=iters, seed=seed, bf16=bf16, **hp)
        for r in range(1, N):
            assert np.array_equal(fin[0], fin[r]), "ranks disagree after the trailing on_start"
        for t in range(iters):
            for r in range(N):
                out[f"{name}/w/{t}/{r}"] = ow[t][r]
                out[f"{name}/h/{t}/{r}"] = oh[t][r]
        out[f"{name}/final"] = fin[0]
        meta[name] = dict(N=N, counts=counts, lr_mult=lm, decay_mult=dm, iters=iters, seed=seed, bf16=bf16, hyper=hp)
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "ref_sync_cases.npz"), **out)
    with open(os.path.join(here, "ref_sync_cases.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays for", len(meta), "cases")


if __name__ == "__main__":
    main()
