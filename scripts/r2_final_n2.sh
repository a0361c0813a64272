#!/bin/bash
# Round 2, final 2-GPU validation of the committed code: parity subset + the default bench line at N=2.
N=2
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "push or golden or multi_round or nvls or odd_shapes or bf16_wire or l1" > $OUT/r2f_pytest_n$N.log 2>&1; echo "rc=$?"; tail -3 $OUT/r2f_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench default"; timeout 600 $TR --master-port 29671 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2f_bench_n$N.json 2> $OUT/r2f_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/r2f_bench_n$N.json") if l.startswith("{")][0]
    r=d["roofline"]
    print("value %.0f e2e %.0f ms/step %.3f kernel %s in-step %.1f us (min-rank %.1f, b2b %.1f) frac %.3f nccl %.1f us traffic %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["impl_config"]["kernel"], r["kernel_ms"]*1e3, r["kernel_ms_min_over_ranks"]*1e3, r["kernel_ms_back_to_back"]*1e3, r["frac"], d.get("nccl_allreduce_only_ms",0)*1e3, r["traffic"]))
    print("parity", json.dumps(d.get("parity"))[:700])
    for k,v in d.get("workloads",{}).items(): print(k, "value %.0f e2e %.0f kernel %s in-step %.1f us b2b %.1f nccl %.1f us" % (v["value"], v["e2e"]["value"], v["impl_config"]["kernel"], v["roofline"]["kernel_ms"]*1e3, v["roofline"]["kernel_ms_back_to_back"]*1e3, v.get("nccl_allreduce_only_ms",0)*1e3))
except Exception as e: print("bench unreadable", e)
PY
tail -2 $OUT/r2f_bench_n$N.err | cut -c1-300
