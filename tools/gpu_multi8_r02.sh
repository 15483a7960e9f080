#!/bin/bash
# 8-GPU round (charged 8x): bash tools/gpu_multi8_r02.sh <tag> <N>
TAG=${1:-r02n8}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
{ nvidia-smi -L; nvidia-smi topo -m; free -g; nproc; lscpu | grep -i "numa\|model name\|socket"; } > $OUT/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
echo "== bench (strong, fused, overlap)"; timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > $OUT/bench_strong.log 2>&1; grep "^{" $OUT/bench_strong.log | tail -1 > $OUT/bench_n${N}_strong.json; grep -v "^{" $OUT/bench_strong.log | tail -4
python - $OUT/bench_n${N}_strong.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("strong: value",d["value"],"ms",d["ms_per_step"],"roofline",d["roofline"]["frac"],"alone",d["roofline"]["kernel_ms_launch_timed_alone"],"other",d.get("other_scaling"))
    print("e2e",d["e2e"]["value"],d["e2e"]["pcie_h2d_copy_GBps_per_gpu"],d["e2e"].get("numa"))
    for k,v in (d.get("configs") or {}).items():
        print(k,{a:(b["GB/s"] if isinstance(b,dict) and "GB/s" in b else b) for a,b in v.items() if isinstance(b,(dict,int))})
except Exception as e: print("parse failed",e)
PY
echo "== bench strong, driver-like short run (steps 20)"
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-secondary --no-other-scaling 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_steps20.json | cut -c1-220
echo "== bench strong, serialised launches (no overlap)"
timeout 200 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --no-secondary --no-other-scaling --no-overlap 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_nooverlap.json | cut -c1-220
echo "== bench strong, nccl exchange"
timeout 200 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --no-secondary --no-other-scaling --exchange nccl 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_nccl.json | cut -c1-220
echo "== multigpu_check"; timeout 400 $TR tests/multigpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -6 | tee $OUT/multigpu_check.txt
echo "== C++ group (one process, $N GPUs)"; timeout 200 tests/cpp/test_shard_group 2>&1 | tail -8 | tee $OUT/cpp_group.txt
echo "== H2D diag"; timeout 200 $TR tools/h2d_diag.py 2>&1 | grep "^{" | tee $OUT/h2d_diag.json
echo "== reference arm"; timeout 200 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_reference.json | cut -c1-300
