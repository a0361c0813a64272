#!/bin/bash
# Round 2, the 8-GPU call: multi-GPU parity (reduced set), kernel-variant matrices at N=8 (small: LL / push / pull /
# NVLS with traces; large: push / TMA / NVLS unrolls / NVLS+P2P share), bf16 wire, the default bench line
# (CaffeNet + extras + parity vs oracle + NVML NVLink traffic), the reference arm and its message sweep.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1200 -- 'bash scripts/r2_n8.sh 8'
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
export COS_VERBOSE=1
nvidia-smi topo -m > $OUT/r2_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multiproc parity (reduced)"; timeout 400 python -m pytest tests/test_gpu_multiproc.py -m gpu -x -q --timeout 200 \
  -k "(one_process_per_gpu and (1-False-2 or 1-True-4 or 1-False-1)) or (nvls and (4-0 or 2-1))" > $OUT/r2_pytest_n$N.log 2>&1; echo "rc=$?"; tail -3 $OUT/r2_pytest_n$N.log
echo "== matrix small"; timeout 400 $TR --master-port 29641 bench.py --gpus $N --sweep --trace --steps 10 --warmup 3 \
  --sizes 0.0625,0.55,1.64,4,16 --variants ldg,tma,push,ll1,ll,ll4,nvls4 > $OUT/r2_matrix_small_n$N.json 2> $OUT/r2_matrix_small_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2_matrix_small_n$N.err
( timeout 500 python bench.py --impl reference --sweep --gpus $N > $OUT/r2_ref_sweep_n$N.json 2> $OUT/r2_ref_sweep_n$N.err; echo "ref sweep rc=$?" ) &
echo "== matrix large"; timeout 500 $TR --master-port 29642 bench.py --gpus $N --sweep --trace --steps 8 --warmup 3 \
  --sizes 64,232.5,512 --variants tma,push,nvls1,nvls4,nvls8,nvls1p,nvls2p,nvls4p > $OUT/r2_matrix_large_n$N.json 2> $OUT/r2_matrix_large_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2_matrix_large_n$N.err
echo "== matrix bf16"; timeout 300 $TR --master-port 29643 bench.py --gpus $N --sweep --trace --steps 8 --warmup 3 --grad-dtype bf16 \
  --sizes 0.55,232.5 --variants ldg,push,ll > $OUT/r2_matrix_bf16_n$N.json 2> $OUT/r2_matrix_bf16_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2_matrix_bf16_n$N.err
wait
grep "ref sweep" $OUT/r2_ref_sweep_n$N.err | tail -8
echo "== bench default (caffenet + extras + parity)"; timeout 800 $TR --master-port 29644 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/r2_bench_n$N.json") if l.startswith("{")][0]
    print("value %.0f e2e %.0f ms/step %.3f kernel %s nvls %s %.1f us frac %.3f bus %.0f nccl %.1f us traffic %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["impl_config"]["kernel"], d["impl_config"]["nvls"], d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d.get("bus_gbs",0), d.get("nccl_allreduce_only_ms",0)*1e3, d["roofline"]["traffic"]))
    print("traffic_source", d["roofline"].get("traffic_source"))
    print("parity", json.dumps(d.get("parity")))
    for k,v in d.get("workloads",{}).items(): print(k, "value %.0f e2e %.0f kernel %s %.1f us frac %.3f nccl %.1f us" % (v["value"], v["e2e"]["value"], v["impl_config"]["kernel"], v["roofline"]["kernel_ms"]*1e3, v["roofline"]["frac"], v.get("nccl_allreduce_only_ms",0)*1e3))
except Exception as e: print("bench unreadable", e)
PY
tail -3 $OUT/r2_bench_n$N.err | cut -c1-300
echo "== reference arm N=$N (caffenet)"; timeout 400 python bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $OUT/r2_bench_ref_n$N.json 2> $OUT/r2_bench_ref_n$N.err; echo "rc=$?"; cut -c1-600 $OUT/r2_bench_ref_n$N.json
