#!/bin/bash
# r02a: first GPU validation of the round-2 kernels (PDL launch, unified mailbox, batched fold + exchange, group API).
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{ nvidia-smi; nproc; free -g; } > $OUT/box.txt 2>&1
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py -m gpu -q -x --timeout 300 2>&1 | tail -15 | tee $OUT/pytest_new.txt
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py --no-supertable 2>&1 | tail -2 | tee $OUT/bench.json | cut -c1-600
