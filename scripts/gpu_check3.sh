#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "rc=$?"; cat gpurun_out/bench_ref_n1.json; tail -3 gpurun_out/bench_ref_n1.err
echo "== ncu launch list (default bench command, autotune off, all launches)"
COS_BENCH_NO_AUTOTUNE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_lenet.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_lenet.csv
