#!/bin/bash
# Round 2 call I (1 GPU): pair-lane HalfCheetah kernel -- GPU suite, pair vs thread A/B over the
# batch sizes of configs[4] (4096 = one of 8 shards ... 32768 = the whole batch on one GPU),
# rows-in-smem sweep, source-level ncu capture; e2e after the send reorder.
O=gpurun_out/r2_i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.txt >> $O/summary.txt
for n in 4096 8192 16384 32768; do
  python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair >> $O/step_ab.jsonl 2>>$O/step_ab.err
  ENVPOOL_B200_HC_KERNEL=thread python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag thread >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
for ks in 9 14 20 27; do
  ENVPOOL_B200_HC_PAIR_KS=$ks python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 32768 --steps 10 --lead 4 --reps 3 --tag pair_ks$ks >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
ENVPOOL_B200_HC_PAIR_KS=9 python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 4096 --steps 10 --lead 4 --reps 3 --tag pair_ks9 >> $O/step_ab.jsonl 2>>$O/step_ab.err
python profiles/e2e_diag.py > $O/e2e_diag.json 2>$O/e2e_diag.err
for n in 32768 4096; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_pair -s 3 -c 1 -o $O/prof_hc_pair$n \
    python bench.py --task HalfCheetah-v4 --num-envs $n --profile --steps 4 --warmup 3 --no-graph > $O/ncu_hc$n.log 2>&1
done
cat $O/step_ab.jsonl $O/e2e_diag.json >> $O/summary.txt
