#!/bin/bash
# One-GPU check run under gpurun: smoke, GPU parity tests, a short bench, ncu captures.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/nvidia_smi.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== bench caffenet"; timeout 600 python bench.py --workload caffenet --steps 10 --no-kernels > gpurun_out/bench_caffenet_n1.json 2> gpurun_out/bench_caffenet_n1.err; echo "rc=$?"; cat gpurun_out/bench_caffenet_n1.json; tail -5 gpurun_out/bench_caffenet_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "rc=$?"; cat gpurun_out/bench_ref_n1.json
if [ "$1" = "ncu" ]; then
echo "== ncu launch list (default bench command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_lenet.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_caffenet.csv python bench.py --workload caffenet --steps 3 --warmup 3 --no-cpu-baseline --no-kernels --no-graph > gpurun_out/ncu_bench2.log 2>&1; echo "rc=$?"
echo "== ncu full (fused kernel, caffenet + lenet)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_sync_sgd -s 4 -c 2 -f -o gpurun_out/prof_caffenet_n1 python bench.py --workload caffenet --steps 3 --warmup 3 --no-cpu-baseline --no-kernels --no-graph > gpurun_out/ncu_full1.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_sync_sgd -s 4 -c 2 -f -o gpurun_out/prof_lenet_n1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernels --no-graph > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
ls -la gpurun_out
fi
