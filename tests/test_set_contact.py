"""SetContact / SetManifold (AabbTreeWorld.cpp:2797-2846) = pr_set_contacts / pr_set_manifolds.

  1. the REAL reference (SetContact / SetManifold through its API, oracle/ref_harness.cpp) == the restatement over a whole
     trajectory in which contacts are edited between steps (conveyor tangent speed, friction, disabled contacts, impulses);
  2. the device kernels compiled for the host (tests/_emul) == the restatement fed the colour-sorted order; 3. same on the GPU;
  4. the reference's refusals: unknown id, freed id, changed contactable / sensor / impenetrable / TOI / TOI count; manifold
     type or point count."""
import numpy as np
import pytest
import refworld
import scenes
import scenes_np
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W

ENABLED, TOUCHING, DIRTY = 0x02, 0x01, 0x10


def _edit(cs, ms, step):
    """Deterministic edits of the current contacts: every 3rd touching contact becomes a conveyor, every 5th loses its friction,
    every 7th is disabled (and re-enabled 20 steps later), every 4th touching manifold gets its impulses halved."""
    rows = cs.copy()
    t = (rows["flags"] & TOUCHING) != 0
    k = np.arange(len(rows))
    conv = t & (k % 3 == 0)
    rows["tangent_speed"][conv] = np.float32(0.75)
    rows["friction"][t & (k % 5 == 0)] = np.float32(0.0)
    rows["restitution"][t & (k % 5 == 1)] = np.float32(0.5)
    if step % 40 == 10:
        rows["flags"][k % 7 == 0] &= ~np.uint32(ENABLED)
    if step % 40 == 30:
        rows["flags"] |= np.uint32(ENABLED)
    rows["flags"][k % 11 == 0] |= np.uint32(DIRTY)
    mids = rows["id"][t & (k % 4 == 0)]
    mrows = ms[t & (k % 4 == 0)].copy()
    for f in ("jn0", "jt0", "jn1", "jt1"):
        mrows[f] = mrows[f] * np.float32(0.5)
    return rows, mids, mrows


@need_ref
@need_oracle
@pytest.mark.parametrize("name", ["pyramid5", "mixed100", "shooter"])
def test_restatement_set_contact_equals_reference(name):
    ref = scenes.SCENES[name]()
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    edits = 0
    for s in range(150):
        if s % 10 == 0 and s:
            cs, ms = ref.contacts()
            rows, mids, mrows = _edit(cs, ms, s)
            ref.set_contacts(rows); orc.set_contacts(rows)
            ref.set_manifold_impulses(mids, mrows); orc.set_manifolds(mids, mrows)
            edits += len(rows)
            rc, rm = ref.contacts(); oc, om = orc.contacts()
            assert_bits(rc, oc, "%s contacts right after the edit at %d" % (name, s))
            assert_bits(rm, om, "%s manifolds right after the edit at %d" % (name, s))
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "%s step %d stats" % (name, s))
        assert_bits(ref.bodies(), orc.bodies(), "%s step %d bodies" % (name, s))
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "%s step %d contacts" % (name, s))
        assert_bits(rm, om, "%s step %d manifolds" % (name, s))
    assert edits > 50


def _device_vs_oracle(sc, steps, lib=None, device=-1):
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    edits = 0
    for s in range(steps):
        if s % 10 == 0 and s:
            cs, ms = gw.contacts()
            rows, mids, mrows = _edit(cs, ms, s)
            gw.set_contacts(rows); orc.set_contacts(rows)
            gw.set_manifolds(mids, mrows); orc.set_manifolds(mids, mrows)
            edits += len(rows)
            gc, gm = gw.contacts(); oc, om = orc.contacts()
            assert_bits(gc, oc, "contacts right after the edit at %d" % s)
            assert_bits(gm, om, "manifolds right after the edit at %d" % s)
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert_bits(gc, oc, "step %d contacts" % s)
        assert_bits(gm, om, "step %d manifolds" % s)
    assert edits > 50
    return gw


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder,steps", [
    ("pyramid6", lambda: scenes_np.pyramid(6), 120),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150),
])
def test_emulated_set_contact_equals_oracle(name, builder, steps):
    _device_vs_oracle(builder(), steps, lib=emul_lib())


@need_emul
def test_set_contact_refusals():
    gw = W.World(scenes_np.pyramid(4), lib=emul_lib())
    conf = refworld.default_step_conf()
    for _ in range(30):
        gw.step(conf)
    cs, ms = gw.contacts()
    assert len(cs) > 4
    lib, h = gw.lib, gw.h

    def put(row):
        one = np.array([row], abi.CONTACT)
        return lib.pr_set_contacts(h, 1, abi.ptr(one))

    ok = cs[0].copy()
    assert put(ok) == 0
    bad = cs[0].copy(); bad["id"] = 1 << 20
    assert put(bad) == -2 and b"no such contact" in lib.pr_last_error(h)
    for field, delta in (("body_a", 1), ("child_b", 1), ("shape_a", 1), ("toi_count", 1)):
        bad = cs[0].copy(); bad[field] += delta
        assert put(bad) == -1, field
    for bit in (0x20, 0x80, 0x40, 0x08):  # sensor, impenetrable, destroyed, has-TOI
        bad = cs[0].copy(); bad["flags"] ^= bit
        assert put(bad) == -1, hex(bit)
    assert "cannot change" in lib.pr_last_error(h).decode() or "may not change" in lib.pr_last_error(h).decode()
    # nothing of a refused batch is applied
    two = np.array([cs[1], cs[0]], abi.CONTACT); two[0]["friction"] = 9.0; two[1]["body_a"] += 1
    assert lib.pr_set_contacts(h, 2, abi.ptr(two)) == -1
    assert_bits(gw.contacts()[0], cs, "contacts after a refused batch")
    # manifolds: type / count are immutable
    t = np.nonzero(ms["count"] > 0)[0][0]
    m = np.array([ms[t]], abi.MANIFOLD); ids = np.array([cs[t]["id"]], np.uint32)
    m["jn0"] = 3.0
    assert lib.pr_set_manifolds(h, 1, abi.ptr(ids), abi.ptr(m)) == 0
    assert gw.contacts()[1][t]["jn0"] == np.float32(3.0)
    m["count"] += 1
    assert lib.pr_set_manifolds(h, 1, abi.ptr(ids), abi.ptr(m)) == -1
    m["count"] -= 1; m["type"] += 1
    assert lib.pr_set_manifolds(h, 1, abi.ptr(ids), abi.ptr(m)) == -1


@pytest.mark.gpu
@need_oracle
def test_gpu_set_contact_equals_oracle():
    _device_vs_oracle(scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150, device=0)
