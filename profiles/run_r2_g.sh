#!/bin/bash
# Round 2 call G (1 GPU): state check after the container restart -- full GPU suite, the driver's
# bench line (own arm + reference arm), ncu launch list.
O=gpurun_out/r2_g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_gpu.txt >> $O/summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench20.txt 2>$O/bench20.err; echo "bench rc=$?" >> $O/summary.txt
tail -1 $O/bench20.txt >> $O/summary.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.txt 2>$O/bench_ref.err
tail -c 1500 $O/bench_ref.txt >> $O/summary.txt
