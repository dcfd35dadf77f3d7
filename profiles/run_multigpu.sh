#!/bin/bash
# Multi-GPU configs of BASELINE.json (run with `gpurun --gpus G -- 'bash profiles/run_multigpu.sh G'`).
set -u
G=${1:-8}
OUT=gpurun_out/r1_mg$G
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -2 > $OUT/pytest_sharded.txt
$TR --master-port 29601 bench.py --gpus $G --steps 20000 --warmup 2000 2>/dev/null | tail -1 > $OUT/bench_cartpole65536.json
for spec in "Pendulum-v1 $((1048576 / G))" "Acrobot-v1 $((1048576 / G))" "FrozenLake-v1 $((4194304 / G))" "Catch-v0 $((4194304 / G))" "HalfCheetah-v4 $((32768 / G))"; do
  set -- $spec
  steps=2000; [ "$1" = "HalfCheetah-v4" ] && steps=300
  $TR --master-port 29602 bench.py --gpus $G --task $1 --num-envs $2 --steps $steps --warmup 100 --no-cpu 2>/dev/null | tail -1 > $OUT/bench_$1_$2.json
done
echo done > $OUT/DONE
