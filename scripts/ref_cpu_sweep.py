"""Reference socket path (oracle/_ref) message sweep on THIS machine's CPU cores
(config 5's CPU column; indicative -- the GPU box's own cores are timed by bench.py)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
rows = []
for N in (2, 4, 8):
    for S in (64 << 10, 1 << 20, 16 << 20, 64 << 20):
        P = S // 4
        iters = max(3, min(50, int(2000 / max(1.0, S / 1e6 * N))))
        r = O.run_ref_time(N, [P], iters=iters, lr_policy="fixed", base_lr=0.01, momentum=0.9, weight_decay=0.0005)
        bus = S * 2 * (N - 1) / N / (r["ms_sync_median"] * 1e-3) / 1e9
        rows.append(dict(bytes=S, ranks=N, ms_per_iter=r["ms_per_iter_median"], ms_sync=r["ms_sync_median"], bus_gbs=bus,
                         cores=r["cores"]))
        print(rows[-1], flush=True)
json.dump({"host": "build container (8 cores)", "rows": rows}, open("profiles/r01_reference_socket_sweep_container.json", "w"), indent=1)
