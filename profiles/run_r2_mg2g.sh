#!/bin/bash
# Round 2 multi-GPU check 4 (gpurun --gpus 2): three pushes in flight (per-slot flags) -- exchange
# tests, the bench line at N=2, the device timeline.
N=${1:-2}
O=gpurun_out/r2_mg2g; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_exchange.py -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.txt >> $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > $O/bench.txt 2>$O/bench.err
tail -1 $O/bench.txt > $O/bench_line.json
python - <<PY | tee -a $O/summary.txt
import json
try:
    d = json.loads(open("$O/bench_line.json").read())
    wx = d.get('with_exchange') or {}
    print('value', round(d['value']/1e9,3), 'G  us/step', round(d['ms_per_step']*1e3,2), ' replicas', round(d['replicas']['value']/1e9,2), 'G  nvlink in', round(wx.get('nvlink_gbs_in_per_gpu',0),1), 'GB/s  nccl', round((d.get('with_allgather_nccl') or {}).get('value',0)/1e9,2), 'G  e2e', round(d['e2e']['value']/1e6,1), 'M')
    for c in d.get('configs', []):
        print('   ', c.get('task'), c.get('num_envs_per_gpu'), 'value', round(c.get('value',0)/1e9,3), 'G', c.get('value_is'), 'replicas', round(c.get('replicas',{}).get('value',0)/1e9,3), 'G', 'nvlink', round((c.get('with_exchange') or {}).get('nvlink_gbs_in_per_gpu',0),1), c.get('error',''))
except Exception as e:
    print('no line', e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/exchange_trace.py CartPole-v1 65536 2>$O/err_trace.txt | grep -v NCCL > $O/trace.jsonl
