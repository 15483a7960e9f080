#!/bin/bash
# r02h: decoupled overlapped exchange + PDL inside batched calls: group / exchange tests, tiers+batch tests, bench N=1 quick.
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py tests/test_gpu_tiers_batch.py tests/test_gpu_containers.py tests/test_gpu_narrow_division.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== bench"; timeout 600 python bench.py --no-cpu > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log > $OUT/bench.json
python - $OUT/bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value",d["value"],"ms",d["ms_per_step"],"roofline",d["roofline"]["frac"],"alone",d["roofline"]["kernel_ms_launch_timed_alone"])
    print("e2e",{k:v for k,v in d["e2e"].items() if k in("value",)}, d["e2e"]["resident_pipeline"]["runs"])
    for k,v in (d.get("configs") or {}).items():
        print(k,{a:(b["GB/s"] if isinstance(b,dict) and "GB/s" in b else b) for a,b in v.items() if isinstance(b,(dict,int))})
    print(d["secondary"]["c1_i64_1000_sum"])
except Exception as e: print("parse failed",e); print(open(sys.argv[1].replace('.json','.log')).read()[-1500:])
PY
