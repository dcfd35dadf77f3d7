mkdir -p gpurun_out
ENVPOOL_B200_PDL=0 python bench.py --steps 20000 --warmup 2000 --no-cpu 2>&1 | tail -1 | cut -c1-400
ENVPOOL_B200_PDL=1 python bench.py --steps 20000 --warmup 2000 --no-cpu 2>&1 | tail -1 | cut -c1-400
ENVPOOL_B200_PDL=0 python bench.py --steps 20000 --warmup 2000 --no-cpu --no-graph 2>&1 | tail -1 | cut -c1-400
ENVPOOL_B200_PDL=1 python bench.py --steps 20000 --warmup 2000 --no-cpu --no-graph 2>&1 | tail -1 | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:hc_kernel -s 2 -c 1 -o gpurun_out/prof_hc_r1 python bench.py --task HalfCheetah-v4 --num-envs 8192 --profile --steps 4 --warmup 2 --no-graph > gpurun_out/ncu_hc.log 2>&1
tail -2 gpurun_out/ncu_hc.log
