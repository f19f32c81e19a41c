"""Shared comparison helpers for the parity tests (TEST INFRASTRUCTURE)."""
import os
import numpy as np
import pytest
import refworld
from playrho_b200 import abi, world as W

# (the emulated-kernel build of the sincosf variant this host's libm -- hence the oracle -- uses, like the product's default)
EMUL_SO = os.path.join(refworld.ROOT, "tests", "_emul",
                       "libplayrho_b200_emul_fma.so" if W.host_sincosf_variant() == "fma" else "libplayrho_b200_emul.so")
need_ref = pytest.mark.skipif(not refworld.have_ref(), reason="oracle/_ref not built (no /root/reference at build time)")
need_oracle = pytest.mark.skipif(not refworld.have_oracle(), reason="oracle/libpr_oracle.so not built")
need_emul = pytest.mark.skipif(not os.path.exists(EMUL_SO), reason="tests/_emul not built")


def emul_lib():
    return W.load_library(EMUL_SO)


def assert_bits(a, b, what):
    assert refworld.bits_equal(a, b), "%s differ: %s" % (what, refworld.diff_fields(a, b))


def run_vs_oracle(scene, steps, lib=None, conf=None, device=-1, check_every=1, toi_mode=0, large_variants=False, world_kwargs=None):
    """Step the device world and the oracle (fed the device's colour-sorted order) side by side and
    require BIT-EXACT bodies, contacts (ids, keys, flags), manifolds (points, feature ids, impulses),
    proxies (fat AABBs) and step statistics after every step."""
    conf = conf if conf is not None else refworld.default_step_conf(do_toi=0)
    gw = W.World(scene, lib=lib, device=device, record_solve_order=True, large_variants=large_variants, **(world_kwargs or {}))
    gw.set_toi_mode(toi_mode)  # 0 = serial-exact: the reference's TOI event order; 1 = PR_TOI_PARALLEL rounds (product default)
    orc = refworld.OracleWorld(scene)
    orc.set_toi_mode(toi_mode)  # (mode 1: the oracle derives the rounds by rule; the two schedules must be the same)
    last = None
    for s in range(steps):
        gs = gw.step(conf)
        ids, cols = gw.solve_order()
        os_ = orc.step(conf, order=ids)
        where = "step %d" % s
        if toi_mode == 1 and conf["do_toi"]:
            gsch, osch = gw.toi_schedule(), orc.toi_schedule()
            assert gsch.shape == osch.shape and np.array_equal(gsch, osch), "%s TOI schedule differs: device %s oracle %s" % (
                where, gsch[:8].tolist(), osch[:8].tolist())
            run_vs_oracle.toi_events = getattr(run_vs_oracle, "toi_events", 0) + len(gsch)
            run_vs_oracle.toi_multi_rounds = getattr(run_vs_oracle, "toi_multi_rounds", 0) + int(
                (np.bincount(gsch[:, 0]) > 1).sum() if len(gsch) else 0)
        if (s % check_every) and s != steps - 1:
            continue
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gw.bodies(), orc.bodies(), where + " bodies")
        gc, gm = gw.contacts()
        oc, om = orc.contacts()
        assert_bits(gc, oc, where + " contacts")
        assert_bits(gm, om, where + " manifolds")
        assert_bits(gw.proxies(), orc.proxies(), where + " proxies")
        assert_bits(gsa, osa, where + " stats")
        last = (gs, len(gc), gw.profile())
    return last
