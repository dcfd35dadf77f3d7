"""Join an `ncu --page source --csv` SASS dump (stall samples / executed instructions per SASS
address) with `nvdisasm -g` line info of the same kernel, and print per-source-line totals.

    cuobjdump -xelf all envpool_b200/lib/obj/mujoco.o          # -> mujoco.sm_100a.cubin
    nvdisasm -g -c mujoco.sm_100a.cubin > dis.txt
    ncu -i X.ncu-rep --page source --csv > sass.csv
    python profiles/sass_by_line.py sass.csv dis.txt hc_pair_kernel [top] [outer-file]
(use `nvdisasm -gi` together with outer-file, e.g. mujoco_pair.cuh)
"""
import csv
import re
import sys
from collections import defaultdict


def line_map(dis, kernel, outer=None):
    """offset -> (file, line) for the .text section whose name contains `kernel`.  With
    `nvdisasm -gi` every instruction carries its inline chain (innermost first); `outer` picks
    the OUTERMOST frame inside that file (so inlined helpers are charged to their call site)."""
    m, chain, on, last = {}, [], False, ("?", 0)
    for ln in open(dis, errors="ignore"):
        if ln.startswith("//-") and ".text." in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        f = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if f:
            chain.append((f.group(1).split("/")[-1], int(f.group(2))))
            continue
        a = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
        if a:
            if chain:
                cur = chain[0]
                if outer:
                    inside = [c for c in chain if c[0] == outer]
                    cur = inside[-1] if inside else chain[-1]
                last = cur
                chain = []
            m[int(a.group(1), 16)] = last
    return m


def main():
    sass, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    outer = sys.argv[5] if len(sys.argv) > 5 else None
    lm = line_map(dis, kernel, outer)
    rows = list(csv.reader(open(sass)))
    hdr = rows[1]
    ia, isamp, iex, ith = (hdr.index("Address"), hdr.index("# Samples"),
                           hdr.index("Instructions Executed"),
                           hdr.index("Thread Instructions Executed"))
    base = None
    last = ("?", 0)
    samp, ex, th = defaultdict(int), defaultdict(int), defaultdict(int)
    tot_s = tot_e = tot_t = 0
    for r in rows[2:]:
        if len(r) <= ith or not r[ia].startswith("0x"):
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        key = lm.get(addr - base, ("?", 0))
        s, e, t = int(r[isamp] or 0), int(r[iex] or 0), int(r[ith] or 0)
        samp[key] += s; ex[key] += e; th[key] += t
        tot_s += s; tot_e += e; tot_t += t
    print(f"total samples {tot_s}  warp-instructions {tot_e}  thread-instructions {tot_t}  "
          f"avg lanes {tot_t / max(tot_e, 1):.1f}")
    print(f"{'file:line':32s} {'samples':>8s} {'%':>6s} {'warp-inst':>10s} {'%':>6s} {'lanes':>6s}")
    for key in sorted(samp, key=lambda k: -samp[k])[:top]:
        print(f"{key[0] + ':' + str(key[1]):32s} {samp[key]:8d} {100 * samp[key] / tot_s:6.2f} "
              f"{ex[key]:10d} {100 * ex[key] / tot_e:6.2f} {th[key] / max(ex[key], 1):6.1f}")


if __name__ == "__main__":
    main()
