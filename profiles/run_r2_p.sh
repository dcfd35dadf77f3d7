#!/bin/bash
# Round 2 call P (1 GPU): smoke() with its kernel launch list.
O=gpurun_out/r2_p; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
tail -2 $O/smoke.txt >> $O/summary.txt
