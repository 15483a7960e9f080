#!/bin/bash
# Short confirmation call after a kernel-geometry change: GPU parity tests, smoke, bench line, division rows of the matrix.
TAG=${1:-r01zg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q -x --timeout 200 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 200 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
echo "== (kernel, dtype) matrix: division / remainder / power rows"
timeout 150 python tools/dtype_matrix.py --only "div,rem,pow" --out $OUT/div_matrix.md > $OUT/div_matrix.log 2>&1; tail -2 $OUT/div_matrix.log
