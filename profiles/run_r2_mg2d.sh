#!/bin/bash
# Round 2 multi-GPU diagnosis 2 (gpurun --gpus 2): un-pipelined stage times of an exchanged step.
N=${1:-2}
O=gpurun_out/r2_mg2d; mkdir -p $O
run() { tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 profiles/exchange_stage_times.py $TASK $NENV >> $O/stages.jsonl 2>$O/err_$tag.txt
}
TASK=CartPole-v1 NENV=65536 run cp_fused A=1
TASK=CartPole-v1 NENV=65536 run cp_push ENVPOOL_B200_EXCHANGE=push
TASK=Pendulum-v1 NENV=524288 run pe_fused A=1
TASK=Pendulum-v1 NENV=524288 run pe_push ENVPOOL_B200_EXCHANGE=push
TASK=CartPole-v1 NENV=4096 run cp4096_fused A=1
cat $O/stages.jsonl | tee -a $O/summary.txt
