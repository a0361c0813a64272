#!/bin/bash
# Round 2, first GPU call (N GPUs, default 2): full -m gpu parity suite (multi-GPU tests un-skip), smoke,
# and the kernel-variant matrix (pull LDG / pull TMA / push / NVLS / NVLS+P2P share) with per-phase traces.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash scripts/r2_first.sh 2'
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export COS_VERBOSE=1
nvidia-smi topo -m > $OUT/r2_topo_n$N.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke_n$N.log 2>&1; echo "rc=$?"; tail -2 $OUT/r2_smoke_n$N.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/r2_pytest_n$N.log 2>&1; echo "rc=$?"; tail -8 $OUT/r2_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== matrix small"; timeout 600 $TR --master-port 29611 bench.py --gpus $N --sweep --trace --steps 10 --warmup 3 \
  --sizes 0.0625,0.5,1.64,4,16 --variants ldg,tma,push1,push,push4,push8,nvls4,auto > $OUT/r2_matrix_small_n$N.json 2> $OUT/r2_matrix_small_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2_matrix_small_n$N.err
echo "== matrix large"; timeout 600 $TR --master-port 29612 bench.py --gpus $N --sweep --trace --steps 8 --warmup 3 \
  --sizes 64,232.5 --variants ldg,tma,push,push4,nvls1,nvls4,nvls8,nvls1p,nvls2p,nvls4p,auto > $OUT/r2_matrix_large_n$N.json 2> $OUT/r2_matrix_large_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2_matrix_large_n$N.err
echo "== matrix bf16"; timeout 600 $TR --master-port 29613 bench.py --gpus $N --sweep --trace --steps 8 --warmup 3 --grad-dtype bf16 \
  --sizes 0.55,16,232.5 --variants ldg,tma,push,push4 > $OUT/r2_matrix_bf16_n$N.json 2> $OUT/r2_matrix_bf16_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2_matrix_bf16_n$N.err
grep -h "nvls" $OUT/r2_matrix_large_n$N.err | grep caffedistri | sort | uniq -c | head
