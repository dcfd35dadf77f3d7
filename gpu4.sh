mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 20000 --warmup 2000 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_cartpole_b.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --profile --steps 60 --warmup 10 --no-graph > gpurun_out/ncu_launch.log 2>&1
for t in Pendulum-v1 Acrobot-v1 FrozenLake-v1 Catch-v0; do python bench.py --task $t --num-envs 1048576 --steps 2000 --warmup 200 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_$t.json; done
