#!/bin/bash
# the one GPU visit of the alpha-discard feature (written after the round's GPU budget was all but spent: ~40 s of box time):
# only the three alpha tests, no torch import, log pulled back through gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/r2_alpha_gpu_tests.log 2>&1
timeout 36 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "alpha" --durations=5 >> gpurun_out/r2_alpha_gpu_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_alpha_gpu_tests.log
tail -15 gpurun_out/r2_alpha_gpu_tests.log
