#!/bin/bash
# Round 2, fourth GPU call (2 GPUs): validates the flattened push / LL loops and the NVLS zeroing warp
# (parity incl. a multi-round layout), and measures their effect.
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "multi_round or nvls or (one_process_per_gpu and (2 or 4)) or golden or l1 or odd_shapes" > $OUT/r2d_pytest_n$N.log 2>&1; echo "rc=$?"; tail -4 $OUT/r2d_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== matrix"; timeout 600 $TR --master-port 29651 bench.py --gpus $N --sweep --trace --steps 10 --warmup 3 \
  --sizes 0.0625,1.64,232.5 --variants push,ll,nvls1,nvls4 > $OUT/r2d_matrix_n$N.json 2> $OUT/r2d_matrix_n$N.err; echo "rc=$?"
grep "^\[sweep\]" $OUT/r2d_matrix_n$N.err
