#!/bin/bash
mkdir -p gpurun_out
echo "== f2 multi-device on real GPUs"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -k "multi_device" > gpurun_out/pytest_f2.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_f2.log
echo "== multiproc (real GPUs)"; timeout 600 python -m pytest tests/test_gpu_multiproc.py -m gpu -q --timeout 300 -k "one_process_per_gpu or nvls" > gpurun_out/pytest_mp.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_mp.log
