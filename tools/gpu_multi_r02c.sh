#!/bin/bash
# Multi-GPU confirmation with tight timeouts: bash tools/gpu_multi_r02c.sh <tag> <N>
TAG=${1:-r02p}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
echo "== bench strong (headline only)"; timeout 150 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --no-secondary --no-other-scaling 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_headline.json | cut -c1-230
echo "== bench strong steps 20"; timeout 150 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-secondary --no-other-scaling 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_steps20.json | cut -c1-230
echo "== multigpu_check"; timeout 300 $TR tests/multigpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -4 | tee $OUT/multigpu_check.txt
echo "== bench (full line)"; timeout 240 $TR bench.py --gpus $N --steps 200 --warmup 10 > $OUT/bench_strong.log 2>&1; grep "^{" $OUT/bench_strong.log | tail -1 > $OUT/bench_n${N}_strong.json; grep -v "^{" $OUT/bench_strong.log | tail -3
python - $OUT/bench_n${N}_strong.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("strong: value",d["value"],"ms",d["ms_per_step"],"roofline",d["roofline"]["frac"],"alone",d["roofline"]["kernel_ms_launch_timed_alone"],"other",d.get("other_scaling"))
    print("e2e",d["e2e"]["value"],d["e2e"]["pcie_h2d_copy_GBps_per_gpu"])
    for k,v in (d.get("configs") or {}).items():
        print(k,{a:(b["GB/s"] if isinstance(b,dict) and "GB/s" in b else b) for a,b in v.items() if isinstance(b,(dict,int))})
except Exception as e: print("parse failed",e)
PY
