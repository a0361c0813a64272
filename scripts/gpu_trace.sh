#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
echo "== pytest multiproc"; timeout 900 python -m pytest tests/test_gpu_multiproc.py -m gpu -q --timeout 600 -k "one_process_per_gpu or nvls" > gpurun_out/pytest_mp_n$N.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_mp_n$N.log
echo "== sweep trace N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --sweep --trace --sweep-max-bytes 4194304 --steps 8 --warmup 3 > gpurun_out/sweep_trace_n$N.json 2> gpurun_out/sweep_trace_n$N.err; echo "rc=$?"; python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/sweep_trace_n$N.json') if l.startswith('{')][0]
for r in d['sweep']: print(r['bytes']>>10,'KiB',r['algo'],r['kernel'],'%.1f us'%(r['kernel_ms']*1e3),'piped %.1f us'%(r['pipelined_ms']*1e3),'nccl %.1f us'%(r.get('nccl_allreduce_ms',0)*1e3), r.get('trace_us_barrierA_phase1_barrierB_zero'))"; grep -E "Error|error" gpurun_out/sweep_trace_n$N.err | tail -5
