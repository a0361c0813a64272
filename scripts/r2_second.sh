#!/bin/bash
# Round 2, second GPU call (N GPUs, default 2): changed tests only, the reworked push kernel + barrier-internal
# traces, the restructured bench.py (CaffeNet default + extras + parity vs oracle), an ncu attempt at NVLink
# metrics with one process per GPU.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash scripts/r2_second.sh 2'
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export COS_VERBOSE=1
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "train or push or fill or hdf5 or selection or bf16_wire or odd_shapes" > $OUT/r2b_pytest_n$N.log 2>&1; echo "rc=$?"; tail -4 $OUT/r2b_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== matrix small"; timeout 600 $TR --master-port 29621 bench.py --gpus $N --sweep --trace --steps 10 --warmup 3 \
  --sizes 0.0625,0.55,1.64,4,16 --variants ldg,tma,push1,push,push4 > $OUT/r2b_matrix_small_n$N.json 2> $OUT/r2b_matrix_small_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2b_matrix_small_n$N.err
echo "== matrix large"; timeout 600 $TR --master-port 29622 bench.py --gpus $N --sweep --trace --steps 8 --warmup 3 \
  --sizes 64,232.5 --variants tma,push,push4 > $OUT/r2b_matrix_large_n$N.json 2> $OUT/r2b_matrix_large_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2b_matrix_large_n$N.err
echo "== matrix bf16"; timeout 600 $TR --master-port 29623 bench.py --gpus $N --sweep --trace --steps 8 --warmup 3 --grad-dtype bf16 \
  --sizes 0.55,16,232.5 --variants ldg,push,push4 > $OUT/r2b_matrix_bf16_n$N.json 2> $OUT/r2b_matrix_bf16_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2b_matrix_bf16_n$N.err
echo "== bench default (caffenet + extras + parity)"; timeout 900 $TR --master-port 29624 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2b_bench_n$N.json 2> $OUT/r2b_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/r2b_bench_n$N.json") if l.startswith("{")][0]
    print("value %.0f e2e %.0f ms/step %.3f kernel %s %.1f us frac %.3f traffic %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["impl_config"]["kernel"], d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["roofline"]["traffic"]))
    print("traffic_source", d["roofline"].get("traffic_source"))
    print("parity", json.dumps(d.get("parity")))
    for k,v in d.get("workloads",{}).items(): print(k, "value %.0f e2e %.0f kernel %s %.1f us frac %.3f nccl %.1f us" % (v["value"], v["e2e"]["value"], v["impl_config"]["kernel"], v["roofline"]["kernel_ms"]*1e3, v["roofline"]["frac"], v.get("nccl_allreduce_only_ms",0)*1e3))
except Exception as e: print("bench unreadable", e)
PY
tail -5 $OUT/r2b_bench_n$N.err | cut -c1-300
echo "== ncu one process per GPU: NVLink + DRAM bytes of the fused kernel (single pass metrics)"
timeout 300 ncu --target-processes all --metrics nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,nvlrx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:fused_sync --launch-skip 6 --launch-count 4 --csv --log-file $OUT/r2b_ncu_nvlink_n$N.csv \
  $TR --master-port 29625 bench.py --gpus $N --sweep --steps 3 --warmup 2 --sizes 232.5 --variants tma > $OUT/r2b_ncu_nvlink_n$N.out 2> $OUT/r2b_ncu_nvlink_n$N.err; echo "ncu rc=$?"
tail -12 $OUT/r2b_ncu_nvlink_n$N.csv | cut -c1-400
