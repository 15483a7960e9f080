#!/bin/bash
# r02b: group / exchange tests after the reserve fix, the restructured bench (strong-scaling line, configs c3/c5, e2e extras).
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{ nvidia-smi; nproc; free -g; ls /sys/devices/system/node; } > $OUT/box.txt 2>&1
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py -m gpu -q --timeout 300 2>&1 | tail -15 | tee $OUT/pytest_new.txt
echo "== bench"; timeout 900 python bench.py > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log > $OUT/bench.json; tail -c 1500 $OUT/bench.log; echo
python - <<'PY' $OUT/bench.json
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value",d["value"],"ms",d["ms_per_step"],"roofline",d["roofline"]["frac"],"alone",d["roofline"]["kernel_ms_launch_timed_alone"])
    print("e2e",{k:v for k,v in d["e2e"].items() if k in("value","apply_f64_add","resident_pipeline")})
    for k,v in (d.get("configs") or {}).items():
        print(k,{a:(b["GB/s"] if isinstance(b,dict) and "GB/s" in b else b) for a,b in v.items() if isinstance(b,(dict,int))})
    print("cpu",d["cpu_baseline"])
except Exception as e: print("parse failed",e)
PY
