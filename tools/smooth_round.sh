#!/bin/bash
# GPU session for the smooth-loss step: its two tests, then the remaining training tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k "smooth" > gpurun_out/smooth_pytest.log 2>&1
echo "exit $?" >> gpurun_out/smooth_pytest.log
tail -30 gpurun_out/smooth_pytest.log | cut -c 1-300
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k "not smooth" 2>&1 | tail -3 | cut -c 1-300
