#!/bin/bash
# N-GPU run under gpurun --gpus N: cross-process parity tests, bench at N, sweep at N.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
echo "== in-process multi tests (x2 for flakiness)"; for i in 1 2; do timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/pytest_multi_$i.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_multi_$i.log; done
echo "== multiproc tests"; timeout 900 python -m pytest tests/test_gpu_multiproc.py -m gpu -q --timeout 600 > gpurun_out/pytest_multiproc_n$N.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_multiproc_n$N.log
for wl in lenet caffenet; do
echo "== bench $wl N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err; echo "rc=$?"; cat gpurun_out/bench_${wl}_n$N.json; grep -v "^W\|^$\|\*\*\*" gpurun_out/bench_${wl}_n$N.err | tail -8
done
echo "== bench caffenet bf16 N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload caffenet --grad-dtype bf16 --steps 10 --warmup 5 > gpurun_out/bench_caffenet_bf16_n$N.json 2> gpurun_out/bench_caffenet_bf16_n$N.err; echo "rc=$?"; cat gpurun_out/bench_caffenet_bf16_n$N.json
echo "== sweep N=$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --sweep --steps 10 --warmup 3 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err; echo "rc=$?"; cat gpurun_out/sweep_n$N.json; grep -v "^W\|^$\|\*\*\*" gpurun_out/sweep_n$N.err | tail -8
echo "== reference arm N=$N (lenet)"; timeout 900 python bench.py --impl reference --gpus $N --steps 20 > gpurun_out/bench_ref_lenet_n$N.json 2> gpurun_out/bench_ref_lenet_n$N.err; echo "rc=$?"; cat gpurun_out/bench_ref_lenet_n$N.json
