#!/bin/bash
# Session-11 verification: GPU test-suite against the re-ordered oracle + fp-sensitive cases.
TAG=${1:-r01_s11}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
python tools/gpu_fp_sensitive_cases.py > gpurun_out/fp_sensitive_$TAG.log 2>&1; cat gpurun_out/fp_sensitive_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -2
