#!/bin/bash
# Round 2 call N (1 GPU): ncu captures of the final builds (hc_pair_kernel v3, headline kernel).
O=gpurun_out/r2_n; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_pair -s 30 -c 1 -o $O/prof_hc_pair4096 \
    python bench.py --task HalfCheetah-v4 --num-envs 4096 --profile --steps 40 --warmup 3 --no-graph > $O/ncu_hc4096.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_pair -s 30 -c 1 -o $O/prof_hc_pair32768 \
    python bench.py --task HalfCheetah-v4 --num-envs 32768 --profile --steps 40 --warmup 3 --no-graph > $O/ncu_hc32768.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -o $O/prof_step_cartpole65536 \
    python bench.py --profile --steps 60 --warmup 10 --no-graph > $O/ncu_full_cartpole.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file $O/launches_cartpole65536.csv \
    python bench.py --profile --steps 60 --warmup 10 --no-graph > $O/ncu_launches.log 2>&1
echo done > $O/summary.txt
