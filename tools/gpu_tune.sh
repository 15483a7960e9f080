#!/bin/bash
# One short gpurun call for tuning rounds: GPU parity tests, the (kernel, dtype) throughput matrix, the TMA-variant sweep.
# Usage (repo root on the GPU box):  bash tools/gpu_tune.sh [tag]
TAG=${1:-r01k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== dtype matrix"; timeout 400 python tools/dtype_matrix.py --out $OUT/matrix.md 2>&1 | tee $OUT/matrix.txt
if [ -x tools/sweep_tma ]; then echo "== TMA sweep"; timeout 200 tools/sweep_tma 2>&1 | tee $OUT/sweep_tma.txt; fi
