"""stdin: `ncu --page raw --csv`; argv[1]: regex of metric names; stdout: trimmed csv
(metric, unit, one value per captured launch)."""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    sys.exit(0)
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[1])
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
ki = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
if ki is not None:
    w.writerow(["kernel", ""] + [r[ki][:80] for r in rows[2:]])
for i, h in enumerate(hdr):
    if pat.search(h):
        w.writerow([h, units[i]] + [r[i] for r in rows[2:]])
