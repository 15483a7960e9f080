#!/bin/bash
# Multi-GPU round: bash tools/gpu_multi.sh <tag> <N>
TAG=${1:-r01b}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
echo "== multigpu_check"; timeout 600 $TR tests/multigpu_check.py 2>&1 | tail -15 | tee $OUT/multigpu_check.txt
for X in fused nccl; do
echo "== bench weak N=$N $X"; timeout 900 $TR bench.py --gpus $N --steps 200 --warmup 10 --exchange $X 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_weak_$X.json | cut -c1-600
echo "== bench strong N=$N $X"; timeout 900 $TR bench.py --gpus $N --steps 200 --warmup 10 --scaling strong --no-e2e --exchange $X 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_$X.json | cut -c1-600
done
echo "== reference arm under torchrun"; timeout 600 $TR bench.py --impl reference --gpus $N --steps 10 --warmup 2 2>&1 | grep -v "^W\|^\*\*\*" | grep "^{" | tail -1 | tee $OUT/bench_ref_n${N}.json | cut -c1-400
