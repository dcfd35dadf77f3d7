#!/bin/bash
# Round 2 call H (1 GPU): e2e breakdown, step A/B (PDL edges in the graph, CTA size), HalfCheetah
# source-level ncu capture at 32768 and 4096 envs.
O=gpurun_out/r2_h; mkdir -p $O
python profiles/e2e_diag.py > $O/e2e_diag.json 2>$O/e2e_diag.err
python profiles/e2e_diag.py bind > $O/e2e_diag_bind.json 2>>$O/e2e_diag.err
python profiles/step_ab.py --tag default --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_PDL_GRAPH=1 python profiles/step_ab.py --tag pdl_graph --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_STEP_BLOCK=128 python profiles/step_ab.py --tag block128 --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_STEP_BLOCK=128 ENVPOOL_B200_PDL_GRAPH=1 python profiles/step_ab.py --tag block128_pdl --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
for n in 4096 32768; do
  python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 2 --tag hc >> $O/step_ab.jsonl 2>>$O/step_ab.err
done
for n in 32768 4096; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_thread -s 3 -c 1 -o $O/prof_hc_thread$n \
    python bench.py --task HalfCheetah-v4 --num-envs $n --profile --steps 4 --warmup 3 --no-graph > $O/ncu_hc$n.log 2>&1
done
cat $O/step_ab.jsonl $O/e2e_diag.json $O/e2e_diag_bind.json > $O/summary.txt
