mkdir -p gpurun_out/r1_px8; TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29611 bench.py --gpus 8 --steps 2000 --warmup 100 --no-cpu 2>gpurun_out/r1_px8/cp.err | tail -1 > gpurun_out/r1_px8/bench_cartpole65536.json
timeout 120 $TR --master-port 29612 bench.py --gpus 8 --task Catch-v0 --num-envs 524288 --steps 300 --warmup 20 --no-cpu 2>gpurun_out/r1_px8/catch.err | tail -1 > gpurun_out/r1_px8/bench_Catch-v0_524288.json
timeout 120 $TR --master-port 29613 bench.py --gpus 8 --task HalfCheetah-v4 --num-envs 4096 --steps 200 --warmup 20 --no-cpu 2>gpurun_out/r1_px8/hc.err | tail -1 > gpurun_out/r1_px8/bench_HalfCheetah-v4_4096.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r1_px8/*.json")):
    try:
        d=json.load(open(f)); print(f, d.get("value"), d.get("ms_per_step"))
        for k in ("with_allgather","with_allgather_nccl","with_allgather_peer"):
            a=d.get(k)
            if a: print("  ",k, a.get("ms_per_step"), a.get("value"), a.get("nvlink_gbs_in_per_gpu"), a.get("unavailable"))
    except Exception as e: print(f, "ERR", e)
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r1_px8/cp.err | tail -5
