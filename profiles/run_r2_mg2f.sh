#!/bin/bash
# Round 2 multi-GPU diagnosis 3 (gpurun --gpus 2): device timeline of the captured exchange chain.
N=${1:-2}
O=gpurun_out/r2_mg2f; mkdir -p $O
for mode in side inline; do
ENVPOOL_B200_EXCHANGE_CHAIN=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/exchange_trace.py CartPole-v1 65536 >> $O/trace.jsonl 2>$O/err_$mode.txt
done
ENVPOOL_B200_EXCHANGE_DEPTH=8 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/exchange_trace.py CartPole-v1 65536 >> $O/trace.jsonl 2>$O/err_depth8.txt
grep -v NCCL $O/trace.jsonl > $O/summary.txt
