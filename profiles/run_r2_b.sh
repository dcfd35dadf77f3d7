#!/bin/bash
# Round 2 call B: exchange v2 + records + prefetch + small-angle sincos -- tests, step A/B, e2e diagnosis.
O=gpurun_out/r2_b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.txt >> $O/summary.txt
python profiles/step_ab.py --tag default >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_PDL_GRAPH=1 python profiles/step_ab.py --tag pdl_graph >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REFILL_FORK=0 python profiles/step_ab.py --tag no_fork >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REC_SPEC=0 python profiles/step_ab.py --tag no_spec >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_STEP_BLOCK=128 python profiles/step_ab.py --tag block128 >> $O/step_ab.jsonl 2>>$O/step_ab.err
python profiles/step_ab.py --tag f32 --precision f32 >> $O/step_ab.jsonl 2>>$O/step_ab.err
for t in Pendulum-v1 Acrobot-v1 CartPole-v1; do
python profiles/step_ab.py --task $t --num-envs 1048576 --steps 200 --tag 1m >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_REC_SPEC=1 python profiles/step_ab.py --task $t --num-envs 1048576 --steps 200 --tag 1m_spec >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
python profiles/step_ab.py --task FrozenLake-v1 --num-envs 4194304 --steps 100 --tag 4m >> $O/step_ab.jsonl 2>>$O/step_ab.err
python profiles/step_ab.py --task Catch-v0 --num-envs 4194304 --steps 100 --tag 4m >> $O/step_ab.jsonl 2>>$O/step_ab.err
python profiles/e2e_diag.py > $O/e2e_diag_unbound.json 2>$O/e2e_diag.err
python profiles/e2e_diag.py bind > $O/e2e_diag_bound.json 2>>$O/e2e_diag.err
numactl -H > $O/numa.txt 2>&1; nvidia-smi topo -m >> $O/numa.txt 2>&1
cat $O/step_ab.jsonl $O/e2e_diag_unbound.json $O/e2e_diag_bound.json | tee -a $O/summary.txt
