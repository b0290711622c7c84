"""Turn ncu output into the tracked summaries under profiles/.

  python scripts/ncu_summarise.py launches <launches.csv[.gz]> <out.md>
      per-kernel launch count / total / average / share from a
      `--metrics gpu__time_duration.sum --csv` launch list
  python scripts/ncu_summarise.py full <raw.csv> <out.md> [traffic.json]
      key metrics per captured launch from `ncu -i rep --page raw --csv` of a `--set full` capture;
      optionally writes {launch-kind: dram bytes per launch} for bench.py's roofline.traffic
"""
import csv
import gzip
import io
import json
import sys
from collections import OrderedDict, defaultdict

KIND_OF = [  # kernel name prefix -> bench.py kernel-class name (api.cu LaunchKind)
    ("krylov_cgs2", "cgs2_step"), ("krylov_pass_kernel<0, 1>", "dots"), ("krylov_pass_kernel<1, 1>", "dots"),
    ("krylov_pass_kernel<1, 0>", "update"), ("krylov_scale", "scale"),
    ("basis_gemm", "gemm"), ("block_matvec", "matvec"), ("bell_matvec", "matvec"), ("slu_fwd_stage", "fwd_stage"),
    ("slu_bwd_stage", "bwd_stage"), ("slu_top_stage", "top_stage"), ("slu_fused_stage", "top_stage"), ("slu_upper", "top_stage"),
    ("slu_merge", "factor"),
    ("slu_build_rows", "factor"), ("slu_top_factor", "factor"), ("assemble", "assemble"),
    ("boundary", "assemble"),
]


def read_csv(path):
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    lines = [ln for ln in raw.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def short(name):
    name = name.replace("(anonymous namespace)", "anon").split("(")[0].replace("void ", "")
    return name.split("::")[-1]


def launches(path, out):
    rows = read_csv(path)
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
        key = (short(r[ci["Kernel Name"]]), r[ci["Grid Size"]], r[ci["Block Size"]])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    with open(out, "w") as fh:
        fh.write(f"ncu launch list: {n} launches, {tot / 1e3:.2f} ms summed kernel time "
                 "(serialised, cold-cache: use the shares, not the absolutes)\n\n")
        fh.write("| kernel | grid | block | launches | total ms | avg us | share |\n|---|---|---|---|---|---|---|\n")
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {key[0]} | {key[1]} | {key[2]} | {a[0]} | {a[1] / 1e3:.3f} | {a[1] / a[0]:.2f} | "
                     f"{100 * a[1] / tot:.1f}% |\n")
    print(open(out).read())


FULL_COLS = OrderedDict([
    ("gpu__time_duration.sum", "dur us"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__shared_mem_per_block_static", "static smem"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
])


def full(path, out, traffic_out=None):
    rows = read_csv(path)
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    to_base = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3,
               "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    per_kind = defaultdict(list)
    with open(out, "w") as fh:
        cols = [c for c in FULL_COLS if c in ci]
        fh.write("| # | kernel | grid x block | " + " | ".join(FULL_COLS[c] for c in cols) + " |\n")
        fh.write("|---|---|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            if len(r) != len(hdr):
                continue
            name = short(r[ci["Kernel Name"]])
            cells = []
            vals = {}
            for c in cols:
                txt = r[ci[c]].replace(",", "")
                try:
                    v = float(txt)
                except ValueError:
                    cells.append(txt)
                    continue
                v *= to_base.get(units[ci[c]], 1.0)
                vals[c] = v
                cells.append(f"{v:.2f}" if v < 1000 else f"{v:.0f}")
            fh.write(f"| {r[ci['ID']]} | {name} | {r[ci['Grid Size']]} x {r[ci['Block Size']]} | "
                     + " | ".join(cells) + " |\n")
            for prefix, kind in KIND_OF:
                if name.startswith(prefix):
                    grid = int(r[ci["Grid Size"]].replace(",", "").split()[0].strip("()"))
                    if kind in ("fwd_stage", "bwd_stage") and grid >= 148:
                        kind += "0"
                    per_kind[kind].append(1e6 * (vals.get("dram__bytes_read.sum", 0.0)
                                                 + vals.get("dram__bytes_write.sum", 0.0)))
                    break
    print(open(out).read())
    if traffic_out:
        traffic = {k: int(sum(v) / len(v)) for k, v in per_kind.items()}
        with open(traffic_out, "w") as fh:
            json.dump(traffic, fh, indent=1, sort_keys=True)
        print(traffic)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
