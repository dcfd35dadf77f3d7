#!/bin/bash
# Short end-of-round check on one B200: GPU test suite, the driver's bench command, smoke(),
# and a fresh ncu launch list of the bench command (gpurun -- 'bash profiles/run_final.sh').
set -u
OUT=gpurun_out/r1
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $OUT/pytest_gpu_final.txt
cat $OUT/pytest_gpu_final.txt
python bench.py 2>$OUT/bench_cartpole65536.err | tail -1 > $OUT/bench_cartpole65536_final.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r1/bench_cartpole65536_final.json"))
print("bench", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"],
      d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_cartpole65536_final.csv \
    python bench.py --profile --steps 300 --warmup 10 --no-graph > $OUT/ncu_launches_final.log 2>&1
tail -2 $OUT/launches_cartpole65536_final.csv | cut -c1-200
