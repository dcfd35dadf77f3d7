#!/bin/bash
# Round 2 call J (1 GPU): pair kernel v2 (LDS rows, precomputed impedance constants, coalesced
# model prologue, 128-register build for big batches) -- GPU suite, A/B over batch size and
# occupancy variant, headline step A/B (next-record prefetch), ncu of both builds.
O=gpurun_out/r2_j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.txt >> $O/summary.txt
for n in 4096 8192 16384 32768; do
  python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_auto >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
for n in 16384 32768; do
  ENVPOOL_B200_HC_PAIR_MINB=4 python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_minb4 >> $O/step_ab.jsonl 2>>$O/step_ab.err
  ENVPOOL_B200_HC_PAIR_MINB=8 python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_minb8 >> $O/step_ab.jsonl 2>>$O/step_ab.err
  ENVPOOL_B200_HC_PAIR_MINB=8 ENVPOOL_B200_HC_PAIR_KS=8 python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_minb8_ks8 >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
# longer steady-state numbers (200 steps into the episodes: contacts everywhere)
python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 4096 --steps 100 --lead 100 --reps 2 --tag pair_steady >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_HC_KERNEL=thread python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 4096 --steps 100 --lead 100 --reps 2 --tag thread_steady >> $O/step_ab.jsonl 2>>$O/step_ab.err
python profiles/step_ab.py --tag cartpole --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
python profiles/step_ab.py --tag cartpole_lead128 --lead 128 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_STEP_BLOCK=128 python profiles/step_ab.py --tag cartpole_b128_lead128 --lead 128 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_pair -s 3 -c 1 -o $O/prof_hc_pair4096 \
    python bench.py --task HalfCheetah-v4 --num-envs 4096 --profile --steps 4 --warmup 3 --no-graph > $O/ncu_hc4096.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_pair -s 3 -c 1 -o $O/prof_hc_pair32768 \
    python bench.py --task HalfCheetah-v4 --num-envs 32768 --profile --steps 4 --warmup 3 --no-graph > $O/ncu_hc32768.log 2>&1
cat $O/step_ab.jsonl >> $O/summary.txt
