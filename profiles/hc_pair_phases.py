"""Phase breakdown of an hc_pair_kernel ncu capture: stall samples / warp-instructions / active
lanes per phase of pair_substep (mujoco_pair.cuh), inlined helpers charged to their call site.

    python profiles/hc_pair_phases.py sass.csv dis_gi.txt "hc_pair_kernelILi4"
(sass.csv: `ncu -i X.ncu-rep --page source --csv`; dis_gi.txt: `nvdisasm -gi -c` of the cubin)"""
import csv
import re
import sys
from collections import defaultdict

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import sass_by_line as S  # noqa: E402


def phase_ranges(src):
    """(first line, name) of every phase, from the `// ----` markers of pair_substep."""
    marks = []
    for i, ln in enumerate(open(src), 1):
        if "__noinline__ void hcp_sincos" in ln:
            marks.append((i, "sincos (out of line)"))
        m = re.match(r"\s*// ---- (.*?) -+\s*$", ln)
        if m and i > 200:
            marks.append((i, m.group(1)[:44]))
        if "warmstart: the better" in ln:
            marks.append((i, "solve: warmstart choice"))
        if re.search(r"for \(int iter = 0; iter <= cm.max_iter", ln):
            marks.append((i, "solve: Ma, row pass (H, force, cost), grad"))
        if "pair_factor(c, H, FH)" in ln:
            marks.append((i, "solve: factor H, apply, Mv, Jv"))
        if "const double gtol = cm.tolerance" in ln:
            marks.append((i, "solve: line search (first pass fused with J*search)"))
    return sorted(marks)


def main():
    sass, dis, kernel = sys.argv[1:4]
    src = sys.argv[4] if len(sys.argv) > 4 else "envpool_b200/csrc/mujoco_pair.cuh"
    marks = phase_ranges(src)
    lm = S.line_map(dis, kernel, "mujoco_pair.cuh")
    rows = list(csv.reader(open(sass)))
    hdr = rows[1]
    ia, isamp, iex, ith = (hdr.index("Address"), hdr.index("# Samples"),
                           hdr.index("Instructions Executed"),
                           hdr.index("Thread Instructions Executed"))
    agg, tot, base = defaultdict(lambda: [0, 0, 0]), [0, 0, 0], None
    for r in rows[2:]:
        if len(r) <= ith or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        f, ln = lm.get(a - base, ("?", 0))
        if f != "mujoco_pair.cuh":
            name = "kernel wrapper (load/store state, outputs)"
        else:
            name = "helpers"
            for first, nm in marks:
                if ln >= first:
                    name = nm
            if ln == next((m[0] for m in marks if m[1].startswith("sincos")), -1) or \
               ln == next((m[0] for m in marks if m[1].startswith("sincos")), -1) + 1:
                name = "sincos (out of line)"
        v = [int(r[isamp] or 0), int(r[iex] or 0), int(r[ith] or 0)]
        for i in range(3):
            agg[name][i] += v[i]
            tot[i] += v[i]
    print(f"{sass}: {tot[0]} samples, {tot[1]} warp-instructions, "
          f"{tot[2] / max(tot[1], 1):.1f} lanes active on average")
    print(f"| phase | stall samples | warp-instructions | active lanes |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {k} | {100 * v[0] / tot[0]:.1f} % | {100 * v[1] / tot[1]:.1f} % | "
              f"{v[2] / max(v[1], 1):.1f} |")


if __name__ == "__main__":
    main()
