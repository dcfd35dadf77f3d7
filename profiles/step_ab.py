"""A/B timing of the per-step chain (run on the GPU box): us/step of CartPole-v1 (or --task) at
--num-envs through epb_step_many_timed -- timestamps inside the chain, `lead` untimed steps
in front of the K timed ones.  Environment switches (ENVPOOL_B200_*) are read by the engine
at pool creation, so every variant is a separate process."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="CartPole-v1")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, nargs="+", default=[20, 200, 2000])
    ap.add_argument("--lead", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch

    from envpool_b200 import _capi

    eng, ms, iopt, _ = bench.TASKS[args.task]
    dev = torch.device("cuda", 0)
    pool = _capi.CPool(eng, args.num_envs, seed=0, max_episode_steps=ms, iopt=iopt,
                       precision=args.precision)
    actions = bench.make_action_stream(torch, args.task, args.num_envs, dev, 2 * bench.L2_BYTES)
    flush = torch.empty(2 * bench.L2_BYTES, dtype=torch.uint8, device=dev)
    pool.reset_device()
    pool.step_many_device(actions, 0, 256)
    pool.sync()
    out = {"tag": args.tag, "task": args.task, "num_envs": args.num_envs,
           "bytes_per_env_step": pool.bytes_per_env_step}
    peak, _ = bench.measured_peak_hbm()
    for K in args.steps:
        best = []
        for r in range(args.reps + 1):
            flush.fill_(r)
            torch.cuda.synchronize()
            msv = pool.step_many_timed(actions, 0, args.lead + K, args.lead, args.lead + K)
            if r:
                best.append(msv / K * 1e3)
        best.sort()
        us = best[len(best) // 2]
        out[f"us_per_step_K{K}"] = round(us, 4)
        out[f"frac_K{K}"] = round(pool.bytes_per_env_step * args.num_envs / (us * 1e-6) / 1e9
                                  / peak, 4)
        out[f"min_us_K{K}"] = round(best[0], 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
