for args in "--task CartPole-v1 --num-envs 65536 --precision f32" "--task Pendulum-v1 --num-envs 65536" "--task CartPole-v1 --num-envs 1048576" "--task CartPole-v1 --num-envs 1048576 --precision f32" "--task Acrobot-v1 --num-envs 1048576 --precision f32" "--task Pendulum-v1 --num-envs 1048576 --precision f32"; do
python bench.py $args --steps 4000 --warmup 500 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'][:40], d['dtype'], 'step: us=%.2f G/s=%.2f frac=%.3f'%(d['ms_per_step']*1e3, d['value']/1e9, d['roofline']['frac']), 'rollout: T=%d us=%.2f G/s=%.2f frac=%.3f'%(d['rollout']['T_per_launch'], d['rollout']['ms_per_step']*1e3, d['rollout']['value']/1e9, d['rollout']['roofline']['frac']), 'e2e M/s=%.1f'%(d['e2e']['value']/1e6))
"
done
