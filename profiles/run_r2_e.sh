#!/bin/bash
# Round 2 call E (1 GPU): refill upper bound, HalfCheetah lane spreading, fp64 peak.
O=gpurun_out/r2_e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_halfcheetah.py tests/test_gpu_records.py tests/test_abi.py -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
python profiles/step_ab.py --tag default --steps 20 200 >> $O/step_ab.jsonl 2>>$O/step_ab.err
ENVPOOL_B200_NO_REFILL=1 python profiles/step_ab.py --tag no_refill_UPPER_BOUND_wrong_results --steps 20 --lead 8 --reps 3 >> $O/step_ab.jsonl 2>>$O/step_ab.err
for n in 4096 8192 16384 32768; do
  for sh in 0 1 2 3; do
    ENVPOOL_B200_HC_LANE_SHIFT=$sh python profiles/step_ab.py --task HalfCheetah-v4 --num-envs $n --steps 10 --lead 4 --reps 2 --tag hc_shift$sh >> $O/step_ab.jsonl 2>>$O/step_ab.err
  done
done
python profiles/step_ab.py --task HalfCheetah-v4 --num-envs 4096 --steps 10 --lead 4 --reps 2 --tag hc_auto >> $O/step_ab.jsonl 2>>$O/step_ab.err
python -c "
import sys; sys.path.insert(0,'.')
from envpool_b200 import _capi
print('fp64 peak GFLOP/s', _capi.fp64_peak_gflops(0))" >> $O/summary.txt 2>&1
cat $O/step_ab.jsonl >> $O/summary.txt
