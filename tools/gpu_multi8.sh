#!/bin/bash
# 8-GPU round, kept short (charged 8x): bash tools/gpu_multi8.sh <tag> <N>
TAG=${1:-r01g}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
{ nvidia-smi -L; nvidia-smi topo -m; free -g; nproc; } > $OUT/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
echo "== multigpu_check"; timeout 600 $TR tests/multigpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -12 | tee $OUT/multigpu_check.txt
echo "== bench weak fused"; timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --exchange fused 2>&1 | tee $OUT/bench_weak_fused.log | grep "^{" | tail -1 | tee $OUT/bench_n${N}_weak_fused.json | cut -c1-300
echo "== bench strong fused"; timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --scaling strong --no-e2e --exchange fused 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_fused.json | cut -c1-300
echo "== bench strong nccl"; timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --scaling strong --no-e2e --exchange nccl 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_strong_nccl.json | cut -c1-300
tail -5 $OUT/bench_weak_fused.log | cut -c1-300
