"""Turn the raw results of profiles/run_profiles.sh (gpurun_out/r1/) into the tracked
round summary: profiles/r1_summary.md, profiles/r1_bench_lines.jsonl, profiles/r1_*.csv,
and profiles/traffic.json (ncu dram bytes per launch of the headline kernel, read by
bench.py for roofline.traffic)."""
import csv
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r1")
DST = os.path.join(ROOT, "profiles")


def load(path):
    try:
        with open(path) as f:
            txt = f.read().strip()
        return json.loads(txt.splitlines()[-1]) if txt else None
    except Exception:
        return None


def fmt(x, d=2):
    return "-" if x is None else f"{x:,.{d}f}"


def main():
    lines = []
    benches = {}
    for p in sorted(glob.glob(os.path.join(SRC, "bench_*.json"))):
        d = load(p)
        if d:
            benches[os.path.basename(p)[6:-5]] = d
    with open(os.path.join(DST, "r1_bench_lines.jsonl"), "w") as f:
        for name, d in benches.items():
            f.write(json.dumps({"file": name, **d}) + "\n")
    lines.append("# Round 1 measurements (B200, `profiles/run_profiles.sh`)\n")
    gpu = os.path.join(SRC, "gpu.csv")
    if os.path.exists(gpu):
        lines.append("GPU: `" + open(gpu).read().strip().replace("\n", " | ") + "`; host threads: "
                     + open(os.path.join(SRC, "nproc.txt")).read().strip() + "\n")
    lines.append("Roofline peak = `MEASURED_PEAKS.json` hbm_gbs (copy bandwidth measured on this "
                 "pool). `frac` = algorithmic bytes/env-step x env-steps/s / peak.\n")
    lines.append("## bench.py lines (device-resident single-step API; rollout = fused T-step launch; "
                 "e2e = make().step(numpy))\n")
    lines.append("| workload | dtype | us/step | G env-steps/s | B/env-step | HBM frac | rollout G/s "
                 "(frac) | e2e M/s | SM MHz (reasons) |")
    lines.append("|---|---|---|---|---|---|---|---|---|")
    for name, d in benches.items():
        if d.get("impl") == "reference":
            continue
        r, ro, e, c = d["roofline"], d.get("rollout"), d.get("e2e"), d.get("clocks") or {}
        lines.append("| {} | {} | {} | {} | {} | {} | {} | {} | {} |".format(
            d["config"]["workload"].replace(" per GPU x 1 GPU", ""), d["dtype"],
            fmt(d["ms_per_step"] * 1e3), fmt(d["value"] / 1e9, 3), r["bytes_per_env_step"],
            fmt(r["frac"], 3),
            "-" if not ro else f"{fmt(ro['value'] / 1e9, 3)} ({fmt(ro['roofline']['frac'], 3)})",
            "-" if not e else fmt(e["value"] / 1e6, 1),
            f"{c.get('sm_mhz')} ({','.join(c.get('reasons') or []) or 'none'})"))
    ref = benches.get("reference_cartpole65536")
    ours = benches.get("cartpole65536")
    if ref and ours:
        lines.append("\n## Headline vs the reference arm (same box, same workload)\n")
        lines.append(f"* reference CPU thread pool (`bench.py --impl reference`, "
                     f"{ref['cpu_baseline']['cores']} host threads, oracle/_ref): "
                     f"**{fmt(ref['value'] / 1e6, 3)} M env-steps/s**")
        lines.append(f"* ours, e2e (host buffers, H2D+D2H in the timed region): "
                     f"**{fmt(ours['e2e']['value'] / 1e6, 1)} M env-steps/s** "
                     f"= {fmt(ours['e2e']['value'] / ref['value'], 0)}x")
        lines.append(f"* ours, device-resident: **{fmt(ours['value'] / 1e9, 2)} G env-steps/s** "
                     f"= {fmt(ours['value'] / ref['value'], 0)}x; HBM roofline fraction "
                     f"{fmt(ours['roofline']['frac'], 3)}")
        cb = ours.get("cpu_baseline")
        if cb:
            lines.append(f"* cpu_baseline inside the ours-arm line: {fmt(cb['value'] / 1e6, 3)} M/s "
                         f"({cb['kind']}, {cb['cores']} threads; {cb['sample']})")
    # ncu launch list
    lp = os.path.join(SRC, "launches_cartpole65536.csv")
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(DST, "r1_ncu_launches_cartpole65536.csv"))
        rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = {}
        for r in rows[1:]:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        tot = sum(sum(v) for v in agg.values())
        lines.append("\n## ncu launch list of `bench.py --profile --no-graph` (cold-cache, serialised: "
                     "shares, not absolutes)\n")
        lines.append("| kernel | launches | mean ns | share of GPU time |")
        lines.append("|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| `{k[:70]}` | {len(v)} | {fmt(sum(v) / len(v), 0)} | "
                         f"{fmt(100 * sum(v) / tot, 1)} % |")
    for g in (2, 4, 8):
        mg = os.path.join(ROOT, "gpurun_out", f"r1_mg{g}")
        files = sorted(glob.glob(os.path.join(mg, "bench_*.json")))
        if not files:
            continue
        lines.append(f"\n## {g} x B200, one rank per GPU (`profiles/run_multigpu.sh {g}`): env-id "
                     "sharding; `value` has no data-path collective, `with_allgather` adds one "
                     "NCCL all-gather of the packed outputs per step\n")
        lines.append("| workload | G env-steps/s (box) | us/step | + all-gather: G/s | us/step | "
                     "NVLink GB/s in per GPU | e2e M/s (box) |")
        lines.append("|---|---|---|---|---|---|---|")
        with open(os.path.join(DST, f"r1_bench_lines_{g}gpu.jsonl"), "w") as f:
            for pth in files:
                d = load(pth)
                if not d:
                    continue
                f.write(json.dumps({"file": os.path.basename(pth), **d}) + "\n")
                ag = d.get("with_allgather") or {}
                lines.append("| {} | {} | {} | {} | {} | {} | {} |".format(
                    d["config"]["workload"], fmt(d["value"] / 1e9, 3),
                    fmt(d["ms_per_step"] * 1e3),
                    fmt(ag["value"] / 1e9, 3) if ag else "-",
                    fmt(ag["ms_per_step"] * 1e3) if ag else "-",
                    fmt(ag.get("nvlink_gbs_in_per_gpu"), 0) if ag else "-",
                    fmt(d["e2e"]["value"] / 1e6, 1) if d.get("e2e") else "-"))
        st = os.path.join(mg, "pytest_sharded.txt")
        if os.path.exists(st):
            lines.append(f"\n`pytest tests/test_gpu_sharded.py` on this box: "
                         f"`{open(st).read().strip().splitlines()[-1]}`")
    # engine peer exchange vs NCCL (tracked raw lines; runs of profiles/run_exchange8.sh and
    # the 2-GPU equivalents)
    for g in (2, 8):
        xp = os.path.join(DST, f"r1_bench_lines_{g}gpu_exchange.jsonl")
        if not os.path.exists(xp):
            continue
        lines.append(f"\n## {g} x B200: the engine's NVLink peer exchange (`csrc/exchange.cuh`) vs one "
                     f"NCCL all-gather per step (`profiles/r1_bench_lines_{g}gpu_exchange.jsonl`)\n")
        lines.append("| workload | step only us | + engine exchange us (G env-steps/s, NVLink GB/s in "
                     "per GPU) | + NCCL all-gather us (G/s, GB/s) | exchange producer |")
        lines.append("|---|---|---|---|---|")
        for ln in open(xp):
            d = json.loads(ln)
            px, nc = d.get("with_allgather") or {}, d.get("with_allgather_nccl") or {}
            if "value" not in px:
                continue
            prod = "push kernel" if ("HalfCheetah" in d["config"]["workload"]
                                     or "push" in d.get("file", "")) else "fused step-kernel epilogue"
            cell = lambda a: "-" if "value" not in a else "{} ({}, {})".format(
                fmt(a["ms_per_step"] * 1e3, 1), fmt(a["value"] / 1e9, 3),
                fmt(a["nvlink_gbs_in_per_gpu"], 0))
            lines.append("| {} | {} | {} | {} | {} |".format(
                d["config"]["workload"], fmt(d["ms_per_step"] * 1e3, 1), cell(px), cell(nc), prod))
    traffic = {}
    for rep, label, key in (("prof_step_cartpole65536", "step_kernel<CartPole<double>>, N=65536",
                             "CartPole-v1:65536:f64"),
                            ("prof_step_pendulum1m", "step_kernel<Pendulum<double>>, N=1M",
                             "Pendulum-v1:1048576:f64"),
                            ("prof_hc_thread32768", "hc_thread_kernel, N=32768",
                             "HalfCheetah-v4:32768:f64")):
        sp = os.path.join(SRC, rep + ".summary.csv")
        if not os.path.exists(sp):
            continue
        shutil.copy(sp, os.path.join(DST, "r1_ncu_" + rep + ".csv"))
        rows = list(csv.reader(open(sp)))
        lines.append(f"\n## ncu --set full: {label} (`profiles/r1_ncu_{rep}.csv`)\n")
        lines.append("| metric | unit | per launch |")
        lines.append("|---|---|---|")
        rd = wr = None
        for r in rows[1:]:
            lines.append(f"| {r[0]} | {r[1]} | {', '.join(r[2:])} |")
            try:
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[1], 1)
                vals = [float(x.replace(",", "")) * scale for x in r[2:]]
                if r[0] == "dram__bytes_read.sum":
                    rd = sum(vals) / len(vals)
                if r[0] == "dram__bytes_write.sum":
                    wr = sum(vals) / len(vals)
            except ValueError:
                pass
        if rd is not None and wr is not None:
            traffic[key] = rd + wr
    if traffic:
        with open(os.path.join(DST, "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    with open(os.path.join(DST, "r1_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
