"""The two builds of the library restate the two glibc sincosf variants, and the host-side rule picks the one this host's libm uses.

UnitVec::Get(angle) (Library/source/playrho/d2/UnitVec.cpp:27-30) goes through libm, whose sincosf is an ifunc: the "-fma"
variant on CPUs with usable FMA + AVX2, the generic one otherwise (profiles/r01h_sincosf_exhaustive.txt: 34 of 2^32 inputs
differ by one ulp, none of them inside [-pi, pi]). Known answers below were produced by tools/sincos_exhaustive.cpp's two builds.
"""
import ctypes as C
import ctypes.util
import os
import struct
import numpy as np
import pytest
import refworld
from refworld import RefWorld, DYNAMIC
from playrho_b200 import abi, world as W
from common import need_ref

ANGLE_BITS = 0x418A3ADD  # ~17.2787 rad: cos = 0xb794f76f (generic) / 0xb794f770 (fma); sin = -1 in both
COS_BITS = {"generic": 0xB794F76F, "fma": 0xB794F770}
EMUL = {"generic": os.path.join(refworld.ROOT, "tests", "_emul", "libplayrho_b200_emul.so"),
        "fma": os.path.join(refworld.ROOT, "tests", "_emul", "libplayrho_b200_emul_fma.so")}


def _f(bits):
    return struct.unpack("<f", struct.pack("<I", bits))[0]


def _bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def _libm_variant():
    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    libm.sincosf.restype = None
    s, c = C.c_float(), C.c_float()
    libm.sincosf(_f(ANGLE_BITS), C.byref(s), C.byref(c))
    assert _bits(s.value) == 0xBF800000
    got = _bits(c.value)
    assert got in COS_BITS.values(), hex(got)
    return "fma" if got == COS_BITS["fma"] else "generic"


def test_host_rule_picks_the_variant_this_libm_uses():
    assert W.host_sincosf_variant() == _libm_variant()


def test_library_selection_by_environment(monkeypatch):
    monkeypatch.delenv("PLAYRHO_B200_LIB", raising=False)
    monkeypatch.delenv("PLAYRHO_B200_SINCOSF", raising=False)
    auto = W.LIB_PATH_FMA if _libm_variant() == "fma" else W.LIB_PATH_GENERIC
    assert W._select_lib_path() == auto  # the default: the build that restates what this host's libm (the reference here) computes
    monkeypatch.setenv("PLAYRHO_B200_SINCOSF", "generic")
    assert W._select_lib_path() == W.LIB_PATH_GENERIC
    monkeypatch.setenv("PLAYRHO_B200_SINCOSF", "fma")
    assert W._select_lib_path() == W.LIB_PATH_FMA
    monkeypatch.setenv("PLAYRHO_B200_SINCOSF", "auto")
    assert W._select_lib_path() == (W.LIB_PATH_FMA if _libm_variant() == "fma" else W.LIB_PATH_GENERIC)
    monkeypatch.setenv("PLAYRHO_B200_SINCOSF", "fast")
    with pytest.raises(RuntimeError):
        W._select_lib_path()
    monkeypatch.setenv("PLAYRHO_B200_LIB", "/x/y.so")
    assert W._select_lib_path() == "/x/y.so"


@pytest.mark.parametrize("path,variant", [(W.LIB_PATH_GENERIC, 0), (W.LIB_PATH_FMA, 1)])
def test_product_builds_report_their_variant(path, variant):
    if not os.path.exists(path):
        pytest.skip(os.path.basename(path) + " not built")
    lib = W.load_library(path)
    assert lib.pr_sincosf_variant() == variant
    assert lib.pr_host_sincosf_variant() == (1 if _libm_variant() == "fma" else 0)  # the C ABI's own statement of the host rule
    for name in W.EXPORTS:
        assert hasattr(lib, name), name


def _drift_against_reference(libs, device):
    ref = RefWorld()
    disk = ref.add_disk(0.5, density=1.0)
    # moves, does not turn: the step recomputes the transformation from an unchanged angle
    ref.add_body(DYNAMIC, pos=(1.0, 2.0), angle=_f(ANGLE_BITS), vel=(1.0, 0.0), shapes=[disk], allow_sleep=False)
    sc = ref.export_scene()
    conf = refworld.default_step_conf(do_toi=0)
    ref.step(conf)
    rb = ref.bodies()
    host = _libm_variant()
    assert _bits(rb["pos1_a"][0]) == ANGLE_BITS
    assert _bits(rb["xf_qc"][0]) == COS_BITS[host]
    for form, path in libs.items():
        lib = W.load_library(path)
        assert lib.pr_sincosf_variant() == (1 if form == "fma" else 0)
        gw = W.World(sc, lib=lib, device=device)
        gw.step(conf)
        gb = gw.bodies()
        assert _bits(gb["xf_qc"][0]) == COS_BITS[form], form
        if form == host:
            assert refworld.bits_equal(gb, rb), refworld.diff_fields(gb, rb)
        else:
            d = refworld.diff_fields(gb, rb)
            assert not refworld.bits_equal(gb, rb) and "xf_qc" in str(d), d


@need_ref
def test_emulated_builds_against_the_reference_on_a_differing_angle():
    """A free body drifting at one of the 34 angles: the reference's transformation after a step carries this host's libm
    bits; the emulated kernels of the matching build give the same bits, the other build is one ulp off in the cosine."""
    if not all(os.path.exists(p) for p in EMUL.values()):
        pytest.skip("tests/_emul variants not built")
    _drift_against_reference(EMUL, device=-1)


@pytest.mark.gpu
@need_ref
def test_cuda_builds_against_the_reference_on_a_differing_angle():
    """The same on the GPU with the two product builds: the one that host_sincosf_variant() names equals the reference."""
    _drift_against_reference({"generic": W.LIB_PATH_GENERIC, "fma": W.LIB_PATH_FMA}, device=0)
