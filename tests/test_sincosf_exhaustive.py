"""The device path's sincosf (pr_math.cuh glibc_sincosf: every UnitVec::Get(angle) of the step) against this image's libm over ALL
2^32 float inputs (tools/sincos_exhaustive.cpp, host build of the same source, -ffp-contract=off).
glibc selects its sincosf variant by CPU features: the generic form of the restatement (the default build) must equal libm's
generic variant everywhere; the contracted form (-DPR_SINCOSF_FMA) must equal libm's fma variant everywhere (checked when this
CPU has FMA). See profiles/r01h_sincosf_exhaustive.txt."""
import os
import subprocess
import pytest
import refworld

SRC = os.path.join(refworld.ROOT, "tools", "sincos_exhaustive.cpp")


def _build(tmp_path, name, extra):
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-DPR_HOST_EMUL", "-std=c++17", "-pthread", "-w"] + extra + ["-o", exe, SRC])
    return exe


# every float on a box with >= 8 cores (~12 s per run); every 7th on smaller ones so that the CPU suite stays within minutes
STRIDE = 1 if (os.cpu_count() or 1) >= 8 else 7
EXPECT = "4294967296 inputs checked, 0 mismatches" if STRIDE == 1 else " inputs checked, 0 mismatches"


def _cpu_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


def test_generic_form_equals_libm_generic_variant_for_every_float(tmp_path):
    exe = _build(tmp_path, "sincos_generic", [])
    env = dict(os.environ, GLIBC_TUNABLES="glibc.cpu.hwcaps=-FMA,-AVX2")
    r = subprocess.run([exe, str(STRIDE)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and EXPECT in r.stdout, r.stdout


@pytest.mark.skipif(not _cpu_has_fma(), reason="this CPU has no FMA: libm has only its generic sincosf here")
def test_contracted_form_equals_libm_fma_variant_for_every_float(tmp_path):
    exe = _build(tmp_path, "sincos_fma", ["-mfma", "-DPR_SINCOSF_FMA"])
    r = subprocess.run([exe, str(STRIDE)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and EXPECT in r.stdout, r.stdout
