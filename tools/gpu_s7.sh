#!/bin/bash
# round 2 session 7: adaptive sparse contacts on the GPU (lock-step parity + reference golden), then the whole GPU suite
mkdir -p gpurun_out
S=${1:-s7}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse or mobility" > gpurun_out/pytest_asc_$S.log 2>&1; echo "asc rc=$?"; tail -25 gpurun_out/pytest_asc_$S.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_$S.log
