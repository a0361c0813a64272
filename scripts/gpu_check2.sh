#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
for k in 0 1; do
echo "== bench caffenet kernel=$k"; timeout 600 python bench.py --workload caffenet --steps 10 --no-kernels --no-cpu-baseline --kernel $k > gpurun_out/bench_caffenet_n1_k$k.json 2> gpurun_out/bench_caffenet_n1_k$k.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_caffenet_n1_k$k.json'));print(d['roofline'],d['value'],d['e2e']['value'])"; tail -3 gpurun_out/bench_caffenet_n1_k$k.err
done
echo "== sweep N=1"; timeout 900 python bench.py --sweep --steps 10 --warmup 3 > gpurun_out/sweep_n1.json 2> gpurun_out/sweep_n1.err; echo "rc=$?"; cat gpurun_out/sweep_n1.json; tail -3 gpurun_out/sweep_n1.err
echo "== ncu tma"; COS_BENCH_NO_AUTOTUNE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_sync_sgd_tma -s 4 -c 2 -f -o gpurun_out/prof_caffenet_n1_tma python bench.py --workload caffenet --steps 3 --warmup 3 --no-cpu-baseline --no-kernels --no-graph --kernel 1 > gpurun_out/ncu_full3.log 2>&1; echo "rc=$?"
