"""Summarise an ncu source-page CSV: top SASS instructions by stall samples with the dominant stall reason."""
import csv, sys
hdr, data = None, []
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "Address":
        if hdr is not None:
            break            # first kernel instance only
        hdr = r
    elif hdr is not None and len(r) == len(hdr):
        data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
f = lambda r, k: float(r[ci[k]] or 0)
tot = sum(f(r, "# Samples") for r in data)
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("samples", tot, "by reason:", ", ".join(f"{k[6:]}={100*v/tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    dom = max(stalls, key=lambda k: f(r, k))
    print(f"{100*f(r, '# Samples')/tot:5.1f}%  {r[ci['Source']].strip()[:72]:72s} {dom[6:]}")
