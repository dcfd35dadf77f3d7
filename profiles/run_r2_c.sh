#!/bin/bash
# Round 2 call C (1 GPU): everything so far -- tests, kernel A/B, the full bench line, e2e diagnosis, ncu.
O=gpurun_out/r2_c; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.txt >> $O/summary.txt
python profiles/step_ab.py --tag default --steps 20 2000 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REC_SPEC=0 python profiles/step_ab.py --tag no_spec --steps 20 2000 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REFILL_FORK=0 python profiles/step_ab.py --tag no_fork --steps 20 2000 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REC_SPEC=0 ENVPOOL_B200_PDL_GRAPH=1 python profiles/step_ab.py --tag no_spec_pdl --steps 20 2000 >> $O/step_ab.jsonl 2>>$O/step_ab.err
for t in Pendulum-v1 CartPole-v1; do
python profiles/step_ab.py --task $t --num-envs 1048576 --steps 200 --tag 1m >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 32768 --steps 10 --lead 4 --reps 2 --tag hc >> $O/step_ab.jsonl 2>>$O/step_ab.err
python profiles/e2e_diag.py > $O/e2e_diag.json 2>$O/e2e_diag.err
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench20.txt 2>$O/bench20.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench20_ref.txt 2>$O/bench20_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 64 --warmup 8 --profile --no-graph --no-configs > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -o $O/step_cartpole python bench.py --steps 64 --warmup 8 --profile --no-graph --no-configs > $O/ncu_full.log 2>&1
timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:hc_thread -s 4 -c 3 --csv --log-file $O/hc_flops.csv python bench.py --task HalfCheetah-v4 --num-envs 32768 --steps 8 --warmup 4 --profile --no-graph --no-configs > $O/ncu_hc.log 2>&1
cat $O/step_ab.jsonl $O/e2e_diag.json | tee -a $O/summary.txt
tail -c 3000 $O/bench20.txt
