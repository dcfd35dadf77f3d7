set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20000 --warmup 2000 2>&1 | tail -3 | tee gpurun_out/bench_cartpole.json
python bench.py --steps 20000 --warmup 2000 --precision f32 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_cartpole_f32.json
python bench.py --impl reference --steps 100 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile --steps 60 --warmup 10 --no-graph > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 3 -o gpurun_out/prof_cartpole_r1 python bench.py --profile --steps 60 --warmup 10 --no-graph > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
