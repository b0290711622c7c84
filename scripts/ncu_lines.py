"""Aggregate the stall samples of an ncu SASS source page per CUDA source line.

  python scripts/ncu_lines.py <sass.csv> <kernel substring> <nvdisasm -g -c listing> [top]

The CSV comes from `ncu -i rep --page source --csv --print-source sass`, the listing from
`nvdisasm -g -c <cubin>` of the same build (cuobjdump -xelf all lib.so).  Instructions are matched
by their order inside the kernel (the two tools print the same SASS)."""
import csv
import re
import sys
from collections import defaultdict

csv_path, kname, lst_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25

# ---- ncu rows of the first instance of the kernel
rows, hdr, on = [], None, False
for r in csv.reader(open(csv_path)):
    if r and r[0] == "Kernel Name":
        if on and rows:
            break
        on = kname in r[1]
        hdr = None
        continue
    if not on:
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        rows.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

# ---- listing: instruction -> (file, line); a template has one section per instantiation: take the
# one with as many instructions as the captured launch
sections, cur, on = [], ("?", 0), False
for ln in open(lst_path):
    if ln.startswith(".text."):
        on = kname in ln
        if on:
            sections.append([])
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        sections[-1].append(cur)
insts = min(sections, key=lambda sec: abs(len(sec) - len(rows))) if sections else []
if len(insts) != len(rows):
    print(f"warning: {len(rows)} ncu rows vs {len(insts)} listed instructions", file=sys.stderr)

f = lambda r, k: float(r[ci[k]] or 0)
per = defaultdict(lambda: defaultdict(float))
tot = 0.0
for r, where in zip(rows, insts):
    s = f(r, "# Samples")
    tot += s
    per[where]["n"] += s
    for k in stalls:
        per[where][k] += f(r, k)
src = {}
print(f"{kname}: {tot:.0f} samples")
for where, d in sorted(per.items(), key=lambda kv: -kv[1]["n"])[:top]:
    reasons = sorted(((d[k], k[6:]) for k in stalls), reverse=True)[:3]
    fn, line = where
    if fn not in src:
        try:
            src[fn] = open(f"legolas_b200/csrc/{fn}").read().splitlines()
        except OSError:
            src[fn] = []
    text = src[fn][line - 1].strip()[:70] if 0 < line <= len(src[fn]) else ""
    print(f"{100 * d['n'] / tot:5.1f}%  {fn}:{line:<5d} {text:70s} " +
          " ".join(f"{k}={100 * v / max(d['n'], 1):.0f}%" for v, k in reasons if v > 0))
