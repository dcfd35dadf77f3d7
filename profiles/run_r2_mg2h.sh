#!/bin/bash
# Round 2 multi-GPU check 5 (gpurun --gpus 2): where the 10 us of a small push go (two more
# stamps) and whether fewer CTAs (= fewer system fences) or a deeper ring change it.
N=${1:-2}
O=gpurun_out/r2_mg2h; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest.txt >> $O/summary.txt
run() { tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 profiles/exchange_trace.py CartPole-v1 65536 2>$O/err_$tag.txt | grep -v NCCL >> $O/trace.jsonl
}
run default A=1
run ctas64 ENVPOOL_B200_PUSH_CTAS=64
run ctas16 ENVPOOL_B200_PUSH_CTAS=16
run depth8 ENVPOOL_B200_EXCHANGE_DEPTH=8
python - <<PY | tee -a $O/summary.txt
import json
for l in open("$O/trace.jsonl"):
    l = l.strip()
    if not l.startswith("{"): continue
    d = json.loads(l)
    if d["rank"] != 0: continue
    r = d["rows_us"]
    import statistics as st
    def med(f): return round(st.median([f(x) for x in r[2:20]]), 1)
    print("ctas", d["push_ctas"], "depth", d["depth"], "us/step", d["us_per_step"],
          "| credit", med(lambda x: x[1]-x[0]), "stores issued (cta0)", med(lambda x: x[6]-x[1]),
          "fence (cta0)", med(lambda x: x[7]-x[6]), "last publish after cta0 fence", med(lambda x: x[2]-x[7]),
          "wait kernel", med(lambda x: x[5]-x[3]))
PY
