#!/bin/bash
# One gpurun call, N GPUs: everything round 2 needs to decide the open questions of DESIGN.md section 5
# (NVLS vs P2P at N >= 4, small_grid for small messages, per-phase trace, bf16 wire crossover).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash scripts/round2_matrix.sh 8'
# Budget: ~3 min of box time (x N GPUs).
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
export COS_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multiproc parity (incl. NVLS tolerance test)"
timeout 600 python -m pytest tests/test_gpu_multiproc.py -m gpu -q --timeout 300 -k "one_process_per_gpu or nvls" > $OUT/r2_pytest_n$N.log 2>&1; echo "rc=$?"; tail -2 $OUT/r2_pytest_n$N.log
for wl in lenet caffenet; do
  for extra in "" "--nvls"; do
    tag=$(echo "${wl}${extra}" | tr -d ' -')
    timeout 300 $TR --master-port 29601 bench.py --gpus $N --workload $wl --steps 20 --warmup 5 $extra > $OUT/r2_bench_${tag}_n$N.json 2> $OUT/r2_bench_${tag}_n$N.err; echo "bench $wl $extra rc=$?"
  done
done
timeout 300 $TR --master-port 29602 bench.py --gpus $N --workload caffenet --grad-dtype bf16 --steps 10 --warmup 5 > $OUT/r2_bench_caffenet_bf16_n$N.json 2> /dev/null; echo "bf16 rc=$?"
echo "== sweep with NVLS + trace"
timeout 600 $TR --master-port 29603 bench.py --gpus $N --sweep --nvls --trace --steps 8 --warmup 3 > $OUT/r2_sweep_n$N.json 2> $OUT/r2_sweep_n$N.err; echo "sweep rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/r2_bench_*_n$N.json")):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][0]
        print(f, "value %.0f e2e %.0f kernel %s nvls %s %.1f us frac %.3f bus %s nccl %s"%(d["value"],d["e2e"]["value"],d["config"]["kernel"],d["config"]["nvls"],d["roofline"]["kernel_ms"]*1e3,d["roofline"]["frac"],d.get("bus_gbs"),d.get("nccl_allreduce_only_ms")))
    except Exception as e: print(f, "unreadable", e)
try:
    d=[json.loads(l) for l in open("$OUT/r2_sweep_n$N.json") if l.startswith("{")][0]
    for r in d["sweep"]: print(r["bytes"]>>10,"KiB",r["algo"],r["kernel"],"%.1f us"%(r["kernel_ms"]*1e3),"piped %.1f"%(r["pipelined_ms"]*1e3),"bus %.1f"%r.get("bus_gbs",0),"nccl %.1f us"%(r.get("nccl_allreduce_ms",0)*1e3),r.get("trace_us_barrierA_phase1_barrierB_zero"))
except Exception as e: print("sweep unreadable", e)
PY
