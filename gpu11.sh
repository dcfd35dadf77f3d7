python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for args in "--task HalfCheetah-v4 --num-envs 32768 --steps 300 --warmup 20" "--task HalfCheetah-v4 --num-envs 4096 --steps 300 --warmup 20"; do
python bench.py $args --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'][:40], d['dtype'], 'step: us=%.2f G/s=%.4f frac=%.4f'%(d['ms_per_step']*1e3, d['value']/1e9, d['roofline']['frac']), 'rollout: T=%d us=%.2f G/s=%.4f'%(d['rollout']['T_per_launch'], d['rollout']['ms_per_step']*1e3, d['rollout']['value']/1e9), 'e2e M/s=%.1f'%(d['e2e']['value']/1e6))
"
done
ENVPOOL_B200_HC_KERNEL=warp python bench.py --task HalfCheetah-v4 --num-envs 32768 --steps 100 --warmup 10 --no-cpu 2>&1 | tail -1 | cut -c1-300
