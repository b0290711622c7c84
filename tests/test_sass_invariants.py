"""Static checks on the built library's SASS (cuobjdump needs no GPU).

The ring slots of the solve and Gram-Schmidt kernels are read with ordinary shared-memory loads and refilled by bulk
copies (async proxy); every release of a slot must carry a cross-proxy fence (``common.cuh: mbar_release_slot``).  Without
it the library was reproducible with one context per GPU and wrong once in ~1e3 operator applications with three contexts in
flight (profiles/tuning_log_r2.md, "Several contexts in flight") - a bug no single-context test can see, so the
instruction sequence itself is pinned here: every plain mbarrier arrival (SYNCS.ARRIVE...A1T0: the consumers' release;
the producer's arrive.expect_tx has no A1T0) is preceded by FENCE.VIEW.ASYNC, and every bulk copy has a consumer
that waits for it.
"""
import re
import shutil
import subprocess

import pytest

from legolas_b200 import build

RING_KERNELS = ("slu_fwd_stage_kernel", "slu_bwd_stage_kernel", "slu_fused_stage_kernel", "slu_top_stage_kernel",
                "krylov_cgs2_kernel", "krylov_pass_kernel")


@pytest.fixture(scope="module")
def kernels():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        sass = subprocess.run([exe, "-sass", build.build()], stdout=subprocess.PIPE, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError) as exc:   # no CUDA toolkit on this machine
        pytest.skip(f"cuobjdump not usable: {exc}")
    out = {}
    for body in re.split(r"\n\s*Function : ", sass)[1:]:
        name = body.split("\n", 1)[0].strip()
        ins = [re.sub(r"/\*[0-9a-f]+\*/", "", ln).split(";")[0].strip() for ln in body.splitlines()
               if re.search(r"/\*[0-9a-f]{4,}\*/", ln)]
        out[name] = [i for i in ins if i]
    return out


def test_every_ring_release_is_fenced(kernels):
    seen = 0
    for name, ins in kernels.items():
        if not any(k in name for k in RING_KERNELS):
            continue
        releases = [k for k, i in enumerate(ins) if "SYNCS.ARRIVE" in i and "A1T0" in i]
        assert releases, f"{name}: no consumer release found - has the ring protocol changed?"
        for k in releases:
            window = ins[max(0, k - 40):k]
            assert any("FENCE.VIEW.ASYNC" in i for i in window), (name, k, ins[k])
        seen += len(releases)
    assert seen >= 20   # forward / backward / fused / top stage kernels, six fused-step variants, three staged passes


def test_bulk_copies_and_programmatic_launch_are_in_the_hot_kernels(kernels):
    def has(kernel, mnemonic):
        return any(kernel in name and any(mnemonic in i for i in ins) for name, ins in kernels.items())

    for k in ("slu_fwd_stage_kernel", "slu_bwd_stage_kernel", "slu_upper_kernel", "krylov_cgs2_kernel"):
        assert has(k, "UBLKCP"), k            # cp.async.bulk global -> shared
        assert has(k, "SYNCS.PHASECHK"), k    # mbarrier try_wait
    for k in ("slu_fwd_stage_kernel", "slu_bwd_stage_kernel", "slu_upper_kernel"):
        assert has(k, "ACQBULK") and has(k, "PREEXIT"), k   # griddepcontrol.wait / launch_dependents
    assert has("basis_gemm_mma_kernel", "DMMA")             # FP64 tensor-core restart GEMM
