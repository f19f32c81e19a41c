"""The C-ABI library loads and exports every symbol include/playrho_b200.h declares (no compute)."""
import os
import re
import numpy as np
import pytest
import refworld
from playrho_b200 import abi, world as W

HDR = os.path.join(refworld.ROOT, "include", "playrho_b200.h")
built = pytest.mark.skipif(not os.path.exists(W.LIB_PATH), reason="libplayrho_b200.so not built")


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pr_[a-z0-9_]+)\s*\(", src)))


@built
def test_library_exports_every_declared_symbol():
    lib = W.load_library()
    names = declared_functions()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), n
    assert set(W.EXPORTS) <= set(names)
    assert lib.pr_abi_version() == 1


@built
def test_pod_sizes_match_python_mirror():
    lib = W.load_library()
    for which, dt in enumerate([abi.BODY, abi.CHILD, abi.SHAPE, abi.REVOLUTE, abi.MANIFOLD, abi.CONTACT, abi.PROXY,
                                abi.STEP_CONF, abi.STEP_STATS, abi.WORLD_CONF, abi.COUNTS, abi.STEP_PROFILE]):
        assert lib.pr_abi_sizeof(which) == dt.itemsize, (which, lib.pr_abi_sizeof(which), dt.itemsize)


@built
def test_step_conf_defaults_equal_oracle_defaults():
    a = W.default_step_conf()
    if refworld.have_oracle():
        assert refworld.bits_equal(a, refworld.default_step_conf())


def test_no_cpu_fallback_in_package():
    """The package must not import or load anything under oracle/ or tests/_emul."""
    pkg = os.path.join(refworld.ROOT, "playrho_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert "libpr_oracle" not in txt and "ref_harness" not in txt and "orc_" not in txt, f
                if f.endswith(".py"):
                    assert "_emul" not in txt, f


def test_every_kernel_waits_for_its_predecessor():
    """Every launch is a programmatic dependent launch (PR_LAUNCH, pr_compat.h): a kernel body that touched global memory before
    `griddepcontrol.wait` would race with the previous kernel of the stream. Static check: each kernel starts with
    PR_PDL_ENTER() -- or, for the solver kernels that prefetch immutable rows first, contains pr_pdl_wait()."""
    import re
    csrc = os.path.join(refworld.ROOT, "playrho_b200", "csrc")
    missing, seen = [], 0
    for f in sorted(os.listdir(csrc)):
        if not f.endswith((".cuh", ".cu")):
            continue
        s = open(os.path.join(csrc, f)).read()
        for m in re.finditer(r"(?:PR_KERNEL|__global__ void(?: __launch_bounds__\([^)]*\))?|^)\s*(k_\w+)\(", s, flags=re.M):
            brace, semi = s.find("{", m.end()), s.find(";", m.end())
            head = s[m.start():brace]
            if brace < 0 or (0 <= semi < brace) or "<<<" in head or "=" in head:
                continue  # a declaration or a launch, not a definition
            seen += 1
            body = s[brace:brace + 1500]
            if m.group(1) == "k_solve_fused":
                continue  # cooperative launch (cudaLaunchCooperativeKernel): an ordinary stream dependency
            if "PR_PDL_ENTER" not in body[:200] and "pr_pdl_wait" not in body:
                missing.append("%s:%s" % (f, m.group(1)))
    assert seen > 80, seen
    assert not missing, missing
