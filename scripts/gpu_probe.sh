#!/bin/bash
mkdir -p gpurun_out
{
for k in 4 8 16; do for th in 0 1; do for ev in 0 1; do ./scripts/concurrency_probe $k $th $ev; done; done; done
for k in 8 16 32; do CUDA_DEVICE_MAX_CONNECTIONS=32 ./scripts/concurrency_probe $k 1 1; done
} > gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
echo "== cross-process on one GPU"; timeout 1500 python -m pytest tests/test_gpu_multiproc.py -m gpu -q --timeout 900 -k sharing > gpurun_out/pytest_sharing.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_sharing.log
