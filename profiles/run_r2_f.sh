#!/bin/bash
# Round 2 call F (1 GPU): record rings -- tests, step A/B over (ring, refill period), bench line.
O=gpurun_out/r2_f; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.txt >> $O/summary.txt
python profiles/step_ab.py --tag q8_r4 --steps 20 200 2000 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REFILL_EVERY=2 python profiles/step_ab.py --tag q8_r2 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REFILL_EVERY=6 python profiles/step_ab.py --tag q8_r6 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REC_Q=16 ENVPOOL_B200_REFILL_EVERY=8 python profiles/step_ab.py --tag q16_r8 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REC_Q=16 ENVPOOL_B200_REFILL_EVERY=14 python profiles/step_ab.py --tag q16_r14 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REC_Q=4 ENVPOOL_B200_REFILL_EVERY=1 python profiles/step_ab.py --tag q4_r1 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
for t in Pendulum-v1 Acrobot-v1 CartPole-v1; do
python profiles/step_ab.py --task $t --num-envs 1048576 --steps 200 --tag 1m >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench20.txt 2>$O/bench20.err
cat $O/step_ab.jsonl >> $O/summary.txt
tail -c 1500 $O/bench20.txt
