#!/bin/bash
# Round 2 call K (1 GPU): the round's reference measurement set -- GPU suite, both bench arms in
# the driver's form, ncu launch list + --set full of the headline kernel and of hc_pair_kernel,
# HalfCheetah 144-register build A/B, e2e breakdown.
O=gpurun_out/r2_k; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.csv
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.txt >> $O/summary.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.txt 2>$O/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench20.txt 2>$O/bench20.err; echo "bench rc=$?" >> $O/summary.txt
tail -1 $O/bench20.txt | cut -c1-1800 >> $O/summary.txt
for n in 16384 32768; do
  python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_255reg >> $O/step_ab.jsonl 2>>$O/step_ab.err
  ENVPOOL_B200_HC_PAIR_MINB=7 python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_144reg >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
python profiles/e2e_diag.py > $O/e2e_diag.json 2>$O/e2e_diag.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file $O/launches_cartpole65536.csv \
    python bench.py --profile --steps 60 --warmup 10 --no-graph > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -o $O/prof_step_cartpole65536 \
    python bench.py --profile --steps 60 --warmup 10 --no-graph > $O/ncu_full_cartpole.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_pair -s 30 -c 1 -o $O/prof_hc_pair4096 \
    python bench.py --task HalfCheetah-v4 --num-envs 4096 --profile --steps 40 --warmup 3 --no-graph > $O/ncu_hc4096.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_pair -s 30 -c 1 -o $O/prof_hc_pair32768 \
    python bench.py --task HalfCheetah-v4 --num-envs 32768 --profile --steps 40 --warmup 3 --no-graph > $O/ncu_hc32768.log 2>&1
cat $O/step_ab.jsonl $O/e2e_diag.json >> $O/summary.txt
