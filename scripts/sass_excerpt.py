"""profiles/sass_r2.md: per hot kernel of the built library, the SASS instructions that prove the bulk-copy (TMA without
a tensor map: UBLKCP / UBLKPF), mbarrier (SYNCS), programmatic-launch (ACQBULK / griddepcontrol -> no mnemonic of its own,
shown as the PTX it comes from) and FP64 (DFMA) paths, with counts.  Runs here (cuobjdump needs no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "legolas_b200", "liblegolas_b200.so")
KERNELS = ["slu_fwd_stage_kernel", "slu_bwd_stage_kernel", "slu_upper_kernel", "slu_fused_stage_kernel", "krylov_cgs2_kernelILi9E",
           "krylov_pass_kernelILb1ELb1E",
           "slu_merge_kernel", "assemble_kernel", "basis_gemm_rows_kernel", "bell_matvec_real_kernelILi6E"]
PATTERNS = [("UBLKCP", "cp.async.bulk global -> shared (bulk copy engine)"), ("UBLKPF", "cp.async.bulk.prefetch.L2"),
            ("SYNCS", "mbarrier (init / arrive / expect_tx / try_wait)"), ("DFMA", "FP64 FMA"), ("DMMA", "FP64 tensor core"),
            ("SHFL", "warp shuffle"), ("BAR.SYNC", "CTA / named barrier"), ("LDS", "shared-memory load"), ("STS", "shared-memory store"),
            ("LDG", "global load"), ("STG", "global store"), ("LDL", "local-memory load (spill / dynamic index)"),
            ("STL", "local-memory store"), ("ATOM", "global atomic"), ("CCTL", "cache control"), ("ERRBAR", "error barrier"),
            ("DEPBAR", "dependency barrier"), ("ACQBULK", "griddepcontrol.wait"), ("PREEXIT", "griddepcontrol.launch_dependents"),
            ("FENCE.VIEW.ASYNC", "fence.proxy.async (cross-proxy fence: mbarrier init, ring-slot release)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    out = ["# SASS evidence, round 2 (cuobjdump -sass legolas_b200/liblegolas_b200.so, sm_100a)", "",
           "Instruction counts per kernel (static, one template instantiation each) and the first occurrence of the",
           "bulk-copy / mbarrier / programmatic-launch instructions.", ""]
    header = "| kernel | instructions | " + " | ".join(p for p, _ in PATTERNS) + " |"
    out += [header, "|" + "---|" * (len(PATTERNS) + 2)]
    excerpts = []
    for k in KERNELS:
        body = next((f for f in funcs[1:] if k in f.split("\n", 1)[0]), None)
        if body is None:
            continue
        name = body.split("\n", 1)[0].strip()
        lines = [ln for ln in body.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/", ln)]
        ins = [re.sub(r"/\*[0-9a-f]+\*/", "", ln).split(";")[0].strip() for ln in lines]
        counts = collections.OrderedDict((p, sum(1 for i in ins if re.search(r"\b" + re.escape(p), i))) for p, _ in PATTERNS)
        out.append(f"| `{k}` | {len(ins)} | " + " | ".join(str(c) for c in counts.values()) + " |")
        firsts = []
        for p in ("UBLKCP", "UBLKPF", "SYNCS.EXCH", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "ACQBULK", "PREEXIT", "DMMA"):
            hit = next((i for i in ins if p in i), None)
            if hit:
                firsts.append(f"    {hit}")
        # the consumers' release of a ring slot: fence.proxy.async ... plain mbarrier arrival (A1T0), as scheduled
        rel = next((n for n, i in enumerate(ins) if "SYNCS.ARRIVE" in i and "A1T0" in i), None)
        if rel is not None:
            fence = next((n for n in range(rel - 1, max(-1, rel - 40), -1) if "FENCE.VIEW.ASYNC" in ins[n]), None)
            if fence is not None:
                firsts.append("    -- release of a ring slot (common.cuh: mbar_release_slot) --")
                firsts += [f"    {i}" for i in ins[fence:rel + 1] if any(t in i for t in ("FENCE", "WARPSYNC", "SYNCS", "LDS"))]
        if firsts:
            excerpts += [f"`{name}`", "```"] + firsts + ["```", ""]
    out += ["", "Legend: " + "; ".join(f"{p} = {d}" for p, d in PATTERNS), "", "## First occurrences", ""] + excerpts
    path = os.path.join(ROOT, "profiles", "sass_r2.md")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
