#!/bin/bash
# Round 2 call L (1 GPU): line-search shortcut of the pair kernel, env spreading, the GPU suite
# with the alternative-kernel and chain-mode tests.
O=gpurun_out/r2_l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_gpu.txt >> $O/summary.txt
for n in 4096 8192 16384 32768; do
  python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_ls1 >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
for sp in 1 2; do for n in 4096 8192; do
  ENVPOOL_B200_HC_PAIR_SPREAD=$sp python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 3 --tag pair_ls1_spread$sp >> $O/step_ab.jsonl 2>>$O/step_ab.err
done; done
python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 4096 --steps 100 --lead 100 --reps 2 --tag pair_ls1_steady >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_HC_PAIR_SPREAD=1 python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 4096 --steps 100 --lead 100 --reps 2 --tag pair_ls1_spread1_steady >> $O/step_ab.jsonl 2>>$O/step_ab.err
cat $O/step_ab.jsonl >> $O/summary.txt
