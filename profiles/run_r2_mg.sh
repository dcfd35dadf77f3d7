#!/bin/bash
# Round 2 multi-GPU call: bash profiles/run_r2_mg.sh N  (gpurun --gpus N).  The driver's launch
# line for bench.py, the 2-process sharded test and exchange A/B switches.
N=${1:-2}
O=gpurun_out/r2_mg$N; mkdir -p $O
if [ "$N" = "2" ]; then
  python profiles/step_ab.py --tag default --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
  ENVPOOL_B200_NO_REFILL=1 python profiles/step_ab.py --tag no_refill_UPPER_BOUND_wrong_results --steps 20 --lead 8 --reps 3 >> $O/step_ab.jsonl 2>>$O/step_ab.err
  cat $O/step_ab.jsonl >> $O/summary.txt
  timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_exchange.py -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
fi
run() { # tag, extra env...
  tag=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 $EXTRA > $O/bench_$tag.txt 2>$O/bench_$tag.err
  tail -1 $O/bench_$tag.txt | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    wx = d.get('with_exchange') or {}
    print('$tag', 'value', round(d['value']/1e9,3), 'G  ms/step', round(d['ms_per_step']*1e3,2), 'us  replicas', round(d['replicas']['value']/1e9,2), 'G  nvlink in', round(wx.get('nvlink_gbs_in_per_gpu',0),1), 'GB/s  nccl', round((d.get('with_allgather_nccl') or {}).get('value',0)/1e9,2), 'G  e2e', round(d['e2e']['value']/1e6,1), 'M')
    for c in d.get('configs', []):
        print('   ', c.get('task'), c.get('num_envs_per_gpu'), 'value', round(c.get('value',0)/1e9,3), 'G', c.get('value_is'), 'replicas', round(c.get('replicas',{}).get('value',0)/1e9,3), 'G', 'nvlink', round((c.get('with_exchange') or {}).get('nvlink_gbs_in_per_gpu',0),1), c.get('error',''))
except Exception as e:
    print('$tag', 'no line', e)
" | tee -a $O/summary.txt
}
EXTRA="" run default A=1
EXTRA="--no-configs" run push ENVPOOL_B200_EXCHANGE=push
EXTRA="--no-configs" run depth2 ENVPOOL_B200_EXCHANGE_DEPTH=2
EXTRA="--no-configs" run depth8 ENVPOOL_B200_EXCHANGE_DEPTH=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > $O/bench_ref.txt 2>$O/bench_ref.err
tail -c 600 $O/bench_ref.txt >> $O/summary.txt
