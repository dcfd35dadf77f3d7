#!/bin/bash
# Round-1 measurement suite (run on the B200 box: `gpurun -- 'bash profiles/run_profiles.sh'`).
# Writes raw results to gpurun_out/r1/; profiles/summarize.py turns them into profiles/*.md|csv.
# Numbers printed by a run under ncu are never used as bench values.
set -u
OUT=gpurun_out/r1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.csv
nproc > $OUT/nproc.txt
# 1. headline: BASELINE configs[1], both arms
python bench.py --steps 20000 --warmup 2000 2> $OUT/bench_cartpole65536.err | tail -1 > $OUT/bench_cartpole65536.json
python bench.py --impl reference --steps 60 --warmup 3 2>/dev/null | tail -1 > $OUT/bench_reference_cartpole65536.json
python bench.py --steps 20000 --warmup 2000 --precision f32 --no-cpu 2>/dev/null | tail -1 > $OUT/bench_cartpole65536_f32.json
# 2. other BASELINE configs on one GPU (per-GPU shard sizes and full sizes)
for spec in "Pendulum-v1 1048576" "Acrobot-v1 1048576" "Pendulum-v1 131072" "Acrobot-v1 131072" \
            "FrozenLake-v1 4194304" "Catch-v0 4194304" "FrozenLake-v1 524288" "Catch-v0 524288" \
            "CartPole-v1 1048576" "HalfCheetah-v4 32768" "HalfCheetah-v4 4096"; do
  set -- $spec
  steps=2000; [ "$1" = "HalfCheetah-v4" ] && steps=300
  python bench.py --task $1 --num-envs $2 --steps $steps --warmup 100 --no-cpu 2>/dev/null | tail -1 > $OUT/bench_$1_$2.json
done
python bench.py --task Acrobot-v1 --num-envs 1048576 --precision f32 --steps 2000 --warmup 100 --no-cpu 2>/dev/null | tail -1 > $OUT/bench_Acrobot-v1_1048576_f32.json
# 3. ncu: launch list of the bench command (kernel share of the step), then --set full on the top kernel
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 80 --csv --log-file $OUT/launches_cartpole65536.csv \
    python bench.py --profile --steps 60 --warmup 10 --no-graph > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 3 -o $OUT/prof_step_cartpole65536 \
    python bench.py --profile --steps 60 --warmup 10 --no-graph > $OUT/ncu_full_cartpole.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 2 -o $OUT/prof_step_pendulum1m \
    python bench.py --task Pendulum-v1 --num-envs 1048576 --profile --steps 20 --warmup 5 --no-graph > $OUT/ncu_full_pendulum.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hc_thread -s 2 -c 1 -o $OUT/prof_hc_thread32768 \
    python bench.py --task HalfCheetah-v4 --num-envs 32768 --profile --steps 4 --warmup 2 --no-graph > $OUT/ncu_full_hc.log 2>&1
METRICS='gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|smsp__inst_executed.sum|smsp__issue_active.avg.pct|smsp__thread_inst_executed_per_inst_executed.ratio|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct_of_peak_sustained_elapsed|stalled_long_scoreboard_per_issue|stalled_wait_per_issue|stalled_no_instruction_per_issue|sm__pipe_fp64_cycles_active.avg.pct'
for rep in prof_step_cartpole65536 prof_step_pendulum1m prof_hc_thread32768; do
  ncu -i $OUT/$rep.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_pick.py "$METRICS" > $OUT/$rep.summary.csv
done
echo done > $OUT/DONE
