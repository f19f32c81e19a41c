"""SetShape (AabbTreeWorld.cpp:1284-1356) for values whose geometry and filter stay = pr_set_shape_material: friction, restitution,
density, sensor flag; bodies using the shape are woken, the shape's contacts are flagged for updating. Reference (its own
SetShape) == restatement == emulated kernels == GPU, bit for bit."""
import numpy as np
import pytest
import refworld
import scenes
import scenes_np
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W

# (step, shape id, friction, restitution, density, sensor)
EDITS = {30: (1, 0.02, 0.0, 5.0, 0), 60: (1, 0.9, 0.6, 5.0, 0), 90: (0, 0.2, 0.3, 0.0, 0), 110: (1, 0.9, 0.6, 1.0, 0)}


@need_ref
@need_oracle
@pytest.mark.parametrize("name", ["pyramid5", "disks100", "mixed100"])
def test_restatement_set_shape_equals_reference(name):
    ref = scenes.SCENES[name]()
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    for s in range(140):
        if s in EDITS:
            ref.set_shape_material(*EDITS[s]); orc.set_shape_material(*EDITS[s])
            assert_bits(ref.bodies(), orc.bodies(), "%s bodies right after SetShape at %d" % (name, s))
            rc, _ = ref.contacts(); oc, _ = orc.contacts()
            assert_bits(rc, oc, "%s contacts right after SetShape at %d" % (name, s))
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "%s step %d stats" % (name, s))
        assert_bits(ref.bodies(), orc.bodies(), "%s step %d bodies" % (name, s))
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "%s step %d contacts" % (name, s))
        assert_bits(rm, om, "%s step %d manifolds" % (name, s))


def _device_vs_oracle(sc, lib=None, device=-1):
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    for s in range(140):
        if s in EDITS:
            gw.set_shape_material(*EDITS[s]); orc.set_shape_material(*EDITS[s])
            assert_bits(gw.bodies(), orc.bodies(), "bodies right after SetShape at %d" % s)
            assert_bits(gw.contacts()[0], orc.contacts()[0], "contacts right after SetShape at %d" % s)
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert_bits(gc, oc, "step %d contacts" % s)
        assert_bits(gm, om, "step %d manifolds" % s)
    return gw


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder", [("pyramid6", lambda: scenes_np.pyramid(6)), ("disks64", lambda: scenes_np.disks_box(64, seed=7))])
def test_emulated_set_shape_equals_oracle(name, builder):
    gw = _device_vs_oracle(builder(), lib=emul_lib())
    assert gw.lib.pr_set_shape_material(gw.h, 1 << 20, 0.1, 0.1, 1.0, 0) == -2


@pytest.mark.gpu
@need_oracle
def test_gpu_set_shape_equals_oracle():
    _device_vs_oracle(scenes_np.pyramid(6), device=0)
