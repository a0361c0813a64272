#!/bin/bash
mkdir -p gpurun_out
echo "== f2 multi-device on real GPUs"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -k "multi_device" > gpurun_out/pytest_f2.log 2>&1; echo "rc=$?"; grep -E "^E  .*Error|passed|failed" gpurun_out/pytest_f2.log | cut -c1-600 | tail -8
