#!/bin/bash
# N-GPU run under gpurun --gpus N: parity tests, bench at N (both kernels), sweep at N.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$2" != "notest" ]; then
if [ "$2" = "quick" ]; then
echo "== pytest multiproc"; timeout 1200 python -m pytest tests/test_gpu_multiproc.py -m gpu -q --timeout 600 -k one_process_per_gpu > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu_n$N.log
else
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu_n$N.log
fi
fi
for k in 0 1; do
for wl in lenet caffenet; do
echo "== bench $wl N=$N kernel=$k"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps 20 --warmup 5 --kernel $k > gpurun_out/bench_${wl}_n${N}_k$k.json 2> gpurun_out/bench_${wl}_n${N}_k$k.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/bench_${wl}_n${N}_k$k.json') if l.startswith('{')][0];print(d['value'],d['e2e']['value'],d['roofline'],d.get('bus_gbs'),d.get('nccl_allreduce_only_ms'),d['split_ms'])"; grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_${wl}_n${N}_k$k.err | tail -5
done
done
echo "== bench caffenet bf16 N=$N kernel=1"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload caffenet --grad-dtype bf16 --steps 10 --warmup 5 --kernel 1 > gpurun_out/bench_caffenet_bf16_n${N}_k1.json 2> gpurun_out/bench_caffenet_bf16_n${N}_k1.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/bench_caffenet_bf16_n${N}_k1.json') if l.startswith('{')][0];print(d['value'],d['roofline'],d.get('bus_gbs'))"
echo "== sweep N=$N"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --sweep --steps 10 --warmup 3 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/sweep_n$N.json') if l.startswith('{')][0]
for r in d['sweep']: print(r['bytes']>>10,'KiB',r['algo'],r['kernel'],'%.1f us'%(r['kernel_ms']*1e3),'piped %.1f us'%(r['pipelined_ms']*1e3),'bus %.1f GB/s'%r.get('bus_gbs',0),'nccl %.1f us'%(r.get('nccl_allreduce_ms',0)*1e3))"; grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/sweep_n$N.err | tail -5
echo "== reference arm N=$N (lenet, caffenet)"; for wl in lenet caffenet; do timeout 900 python bench.py --impl reference --gpus $N --workload $wl --steps 20 > gpurun_out/bench_ref_${wl}_n$N.json 2> gpurun_out/bench_ref_${wl}_n$N.err; echo "rc=$?"; cat gpurun_out/bench_ref_${wl}_n$N.json; done
