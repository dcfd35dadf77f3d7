mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 20000 --warmup 2000 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_cartpole_c.json
python bench.py --task HalfCheetah-v4 --num-envs 32768 --steps 200 --warmup 20 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_hc.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --profile --steps 60 --warmup 10 --no-graph > gpurun_out/ncu_launch.log 2>&1
