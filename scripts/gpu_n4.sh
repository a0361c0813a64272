#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
export COS_VERBOSE=1
for wl in lenet caffenet; do
echo "== bench $wl N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${wl}_n${N}.json 2> gpurun_out/bench_${wl}_n${N}.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/bench_${wl}_n${N}.json') if l.startswith('{')][0];print(d['value'],d['e2e']['value'],d['config']['kernel'],d['config']['nvls'],d['roofline']['kernel_ms'],d['roofline']['frac'],d.get('bus_gbs'),d.get('nccl_allreduce_only_ms'),d['ms_per_step'])"; grep -E "Error|error" gpurun_out/bench_${wl}_n${N}.err | tail -4
done
echo "== bench caffenet N=$N --nvls"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload caffenet --steps 10 --warmup 5 --nvls > gpurun_out/bench_caffenet_n${N}_nvls.json 2> gpurun_out/bench_caffenet_n${N}_nvls.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/bench_caffenet_n${N}_nvls.json') if l.startswith('{')][0];print(d['value'],d['config']['kernel'],d['config']['nvls'],d['roofline']['kernel_ms'],d.get('bus_gbs'),d.get('nccl_allreduce_only_ms'))"; grep -E "caffedistri|Error|error" gpurun_out/bench_caffenet_n${N}_nvls.err | tail -4
echo "== sweep N=$N (>= 1 MiB, with nvls)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --sweep --nvls --sweep-min-bytes 65536 --steps 8 --warmup 3 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/sweep_n$N.json') if l.startswith('{')][0]
for r in d['sweep']: print(r['bytes']>>10,'KiB',r['algo'],r['kernel'],'%.1f us'%(r['kernel_ms']*1e3),'piped %.1f us'%(r['pipelined_ms']*1e3),'bus %.1f GB/s'%r.get('bus_gbs',0),'nccl %.1f us'%(r.get('nccl_allreduce_ms',0)*1e3))"; grep -E "Error|error" gpurun_out/sweep_n$N.err | tail -5
