#!/bin/bash
# Round 2 call O (1 GPU): last check of the round -- full GPU suite, driver-form bench line.
O=gpurun_out/r2_o; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_gpu.txt >> $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench20.txt 2>$O/bench20.err; echo "bench rc=$?" >> $O/summary.txt
tail -1 $O/bench20.txt | cut -c1-400 >> $O/summary.txt
