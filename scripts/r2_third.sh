#!/bin/bash
# Round 2, third GPU call (2 GPUs): LL kernel parity + latency vs push, e2e after the warm-up fix.
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest ll/push subset"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "ll or push or selection or odd_shapes or fill or (one_process_per_gpu and 4)" > $OUT/r2c_pytest_n$N.log 2>&1; echo "rc=$?"; tail -4 $OUT/r2c_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== matrix small"; timeout 600 $TR --master-port 29631 bench.py --gpus $N --sweep --trace --steps 10 --warmup 3 \
  --sizes 0.0625,0.55,1.64,4,8 --variants push,ll1,ll,ll4 > $OUT/r2c_matrix_small_n$N.json 2> $OUT/r2c_matrix_small_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2c_matrix_small_n$N.err
echo "== matrix small bf16"; timeout 600 $TR --master-port 29632 bench.py --gpus $N --sweep --trace --steps 10 --warmup 3 --grad-dtype bf16 \
  --sizes 0.55,1.64 --variants push,ll > $OUT/r2c_matrix_bf16_n$N.json 2> $OUT/r2c_matrix_bf16_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2c_matrix_bf16_n$N.err
echo "== bench lenet (+extras off)"; timeout 600 $TR --master-port 29633 bench.py --gpus $N --workload lenet --no-extras --steps 50 --warmup 5 > $OUT/r2c_bench_lenet_n$N.json 2> $OUT/r2c_bench_lenet_n$N.err; echo "rc=$?"
python - <<PY
import json
for f in ("$OUT/r2c_bench_lenet_n$N.json",):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][0]
        print("value %.0f e2e %.0f ms/step %.4f kernel %s %.1f us frac %.3f traffic %s nccl %.1f us parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["impl_config"]["kernel"], d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["roofline"]["traffic"], d.get("nccl_allreduce_only_ms",0)*1e3, d.get("parity",{}).get("bit_exact")))
    except Exception as e: print("bench unreadable", e)
PY
tail -3 $OUT/r2c_bench_lenet_n$N.err | cut -c1-300
