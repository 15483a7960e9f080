#!/bin/bash
# r02d: packed 8/16-bit division — parity (whole 8-bit domain, 16-bit boundaries, all 2^32 u16/i16 pairs), geometry, matrix rows.
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== narrow division parity + knobs + containers"; timeout 900 python -m pytest tests/test_gpu_narrow_division.py tests/test_gpu_geometry_knobs.py tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_containers.py tests/test_cpp_host.py -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_div.txt
echo "== whole-domain 16-bit division sweep"; timeout 400 python tests/sweep_div16.py 2>&1 | tail -4 | tee $OUT/exhaustive_div16.txt
echo "== geometry of the packed path"; timeout 600 python tools/heavy_exp.py int8,uint8,int16,uint16 2>&1 | grep -v "pow\|^$" | tee $OUT/heavy_exp_narrow.txt | tail -70
