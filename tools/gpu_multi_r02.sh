#!/bin/bash
# Multi-GPU round (charged N x): bash tools/gpu_multi_r02.sh <tag> <N> [quick]
TAG=${1:-r02m}; N=${2:-2}; QUICK=${3:-}
OUT=gpurun_out/$TAG; mkdir -p $OUT
{ nvidia-smi -L; nvidia-smi topo -m; free -g; nproc; ls /sys/devices/system/node; lscpu | grep -i "numa\|model name\|socket"; } > $OUT/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
if [ -z "$QUICK" ]; then
echo "== multigpu_check"; timeout 600 $TR tests/multigpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -12 | tee $OUT/multigpu_check.txt
echo "== C++ group (one process, $N GPUs)"; timeout 300 tests/cpp/test_shard_group 2>&1 | tail -8 | tee $OUT/cpp_group.txt
echo "== pytest group tests on $N GPUs"; timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_multigpu.py -m gpu -q --timeout 300 2>&1 | tail -5 | tee $OUT/pytest_group.txt
fi
echo "== bench (strong, fused, overlap)"; timeout 900 $TR bench.py --gpus $N --steps 200 --warmup 10 > $OUT/bench_strong.log 2>&1; grep "^{" $OUT/bench_strong.log | tail -1 > $OUT/bench_n${N}_strong.json; tail -c 600 $OUT/bench_strong.log | grep -v "^{" | tail -5
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
echo "== bench strong, no overlap (reduce_overlap off) / nccl exchange"
timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --no-secondary --no-other-scaling --exchange nccl 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_nccl.json | cut -c1-200
echo "== H2D diag"; timeout 300 $TR tools/h2d_diag.py 2>&1 | grep "^{" | tee $OUT/h2d_diag.json; timeout 300 $TR tools/h2d_diag.py --bind 2>&1 | grep "^{" | tee $OUT/h2d_diag_bind.json
echo "== reference arm"; timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_reference.json | cut -c1-300
