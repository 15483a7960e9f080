#!/bin/bash
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
