#!/bin/bash
# Round 2 multi-GPU diagnosis (gpurun --gpus 2): is the exchange chain serialised by hardware-queue
# sharing (a polling wait kernel at the head of a queue that also carries the next step)?
N=${1:-2}
O=gpurun_out/r2_mg2c; mkdir -p $O
run() { tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-configs --no-cpu > $O/bench_$tag.txt 2>$O/bench_$tag.err
  tail -1 $O/bench_$tag.txt | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); wx = d.get('with_exchange') or {}
    print('$tag', 'value', round(d['value']/1e9,3), 'G  us/step', round(d['ms_per_step']*1e3,2), 'replicas', round(d['replicas']['value']/1e9,2), 'nvlink in', round(wx.get('nvlink_gbs_in_per_gpu',0),1), 'nccl', round((d.get('with_allgather_nccl') or {}).get('value',0)/1e9,2))
except Exception as e:
    print('$tag', 'no line', e)" | tee -a $O/summary.txt
}
run conn8 A=1
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_depth8 CUDA_DEVICE_MAX_CONNECTIONS=32 ENVPOOL_B200_EXCHANGE_DEPTH=8
run conn32_inline CUDA_DEVICE_MAX_CONNECTIONS=32 ENVPOOL_B200_EXCHANGE_CHAIN=inline
