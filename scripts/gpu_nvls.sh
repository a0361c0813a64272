#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
export COS_VERBOSE=1
echo "== pytest multiproc (real GPUs)"; timeout 1200 python -m pytest tests/test_gpu_multiproc.py -m gpu -q --timeout 600 -k "one_process_per_gpu or nvls" -s > gpurun_out/pytest_nvls_n$N.log 2>&1; echo "rc=$?"; grep -E "NVLS active|nvls|passed|failed|Error" gpurun_out/pytest_nvls_n$N.log | tail -12
for mode in "" "--nvls"; do
echo "== bench caffenet N=$N $mode"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload caffenet --steps 10 --warmup 5 $mode > gpurun_out/bench_caffenet_n${N}${mode}.json 2> gpurun_out/bench_caffenet_n${N}${mode}.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/bench_caffenet_n${N}${mode}.json') if l.startswith('{')][0];print(d['value'],d['e2e']['value'],d['config']['kernel'],d['config']['nvls'],d['roofline']['kernel_ms'],d['roofline']['frac'],d.get('bus_gbs'),d.get('nccl_allreduce_only_ms'))"; grep -E "caffedistri|Error|error" gpurun_out/bench_caffenet_n${N}${mode}.err | tail -4
done
echo "== bench lenet N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_lenet_n${N}.json 2> gpurun_out/bench_lenet_n${N}.err; echo "rc=$?"; cat gpurun_out/bench_lenet_n${N}.json
echo "== sweep N=$N (with nvls)"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --sweep --nvls --steps 10 --warmup 3 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/sweep_n$N.json') if l.startswith('{')][0]
for r in d['sweep']: print(r['bytes']>>10,'KiB',r['algo'],r['kernel'],'%.1f us'%(r['kernel_ms']*1e3),'piped %.1f us'%(r['pipelined_ms']*1e3),'bus %.1f GB/s'%r.get('bus_gbs',0),'nccl %.1f us'%(r.get('nccl_allreduce_ms',0)*1e3))"; grep -E "Error|error" gpurun_out/sweep_n$N.err | tail -5
