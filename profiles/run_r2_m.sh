#!/bin/bash
# Round 2 call M (1 GPU): full GPU suite after the exchange changes (per-slot flags, three push
# branches created on first use), then the driver-form bench line.
O=gpurun_out/r2_m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -6 $O/pytest_gpu.txt >> $O/summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench20.txt 2>$O/bench20.err; echo "bench rc=$?" >> $O/summary.txt
tail -1 $O/bench20.txt | cut -c1-900 >> $O/summary.txt
