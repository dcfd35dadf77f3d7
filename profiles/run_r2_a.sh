#!/bin/bash
# Round 2 call A: reset-ahead records -- parity tests, headline bench, A/B switches, ncu.
O=gpurun_out/r2_a; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench20_$rep.txt 2>$O/bench20_$rep.err
done
timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu > $O/bench2000.txt 2>&1
ENVPOOL_B200_REC_SPEC=0 timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu --profile > $O/bench2000_nospec.txt 2>&1
ENVPOOL_B200_REFILL_FORK=0 timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu --profile > $O/bench2000_nofork.txt 2>&1
for t in Pendulum-v1 Acrobot-v1; do
timeout 300 python bench.py --task $t --num-envs 1048576 --steps 500 --warmup 50 --no-cpu --profile > $O/bench_${t}_1m.txt 2>&1
ENVPOOL_B200_REC_SPEC=1 timeout 300 python bench.py --task $t --num-envs 1048576 --steps 500 --warmup 50 --no-cpu --profile > $O/bench_${t}_1m_spec.txt 2>&1
done
timeout 300 python bench.py --task CartPole-v1 --num-envs 1048576 --steps 500 --warmup 50 --no-cpu --profile > $O/bench_cartpole_1m.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 64 --warmup 8 --profile --no-graph > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 2 -o $O/step_cartpole python bench.py --steps 64 --warmup 8 --profile --no-graph > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:refill_kernel -s 20 -c 2 -o $O/refill_cartpole python bench.py --steps 64 --warmup 8 --profile --no-graph > $O/ncu_full2.log 2>&1
grep -h '"value"' $O/bench*.txt | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['workload'][:40], round(d['ms_per_step']*1e3,3), 'us', round(d['roofline']['frac'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'))
" | tee -a $O/summary.txt
