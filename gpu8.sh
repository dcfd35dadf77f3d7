python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for args in "--task CartPole-v1 --num-envs 65536" "--task CartPole-v1 --num-envs 1048576" "--task FrozenLake-v1 --num-envs 1048576" "--task Acrobot-v1 --num-envs 1048576" "--task HalfCheetah-v4 --num-envs 32768 --steps 300 --warmup 20" "--task HalfCheetah-v4 --num-envs 4096 --steps 300 --warmup 20"; do
python bench.py --steps 4000 --warmup 500 $args --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'][:40], d['dtype'], 'step: us=%.2f G/s=%.3f frac=%.3f'%(d['ms_per_step']*1e3, d['value']/1e9, d['roofline']['frac']), 'rollout: T=%d us=%.2f G/s=%.3f frac=%.3f'%(d['rollout']['T_per_launch'], d['rollout']['ms_per_step']*1e3, d['rollout']['value']/1e9, d['rollout']['roofline']['frac']), 'e2e M/s=%.1f'%(d['e2e']['value']/1e6))
"
done
