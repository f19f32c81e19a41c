"""Destroy(world, JointID) (AabbTreeWorld.cpp:1215-1227) = pr_destroy_joints, and the re-use of the freed id by the next CreateJoint.
Three pendulums on revolute joints from one static body; the middle joint is destroyed in mid-swing (its box drops onto the
ground), later a new joint -- which gets the freed id back -- pins the box where it came to rest. Reference (its own Destroy /
CreateJoint) == restatement == emulated kernels == GPU, bit for bit."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W


def _build():
    ref = RefWorld()
    ground = ref.add_edge((-20.0, 0.0), (20.0, 0.0))
    ref.add_body(STATIC, shapes=[ground])
    hub = ref.add_body(STATIC, pos=(0.0, 10.0))
    box = ref.add_box(0.5, 0.3, density=2.0)
    ps = []
    for i in range(3):
        p = ref.add_body(DYNAMIC, pos=(-5.0 + 5.0 * i, 6.0), vel=(2.0 + i, 0.0), acc=(0.0, -9.8), shapes=[box])
        ref.add_revolute(hub, p, (-5.0 + 5.0 * i, 9.0), enable_motor=(i == 2), motor_speed=0.5, max_motor_torque=20.0)
        ps.append(p)
    return ref, hub, ps


def _run(lib=None, device=-1, with_device=True):
    ref, hub, ps = _build()
    sc = ref.export_scene()
    orc = refworld.OracleWorld(sc)
    twins = []
    if with_device:
        gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
        gw.set_toi_mode(0)
        orc2 = refworld.OracleWorld(sc)
        twins = [gw, orc2]
    conf = refworld.default_step_conf()
    for s in range(170):
        if s == 40:
            ref.destroy_joint(1); orc.destroy_joints([1])
            for t in twins:
                t.destroy_joints([1])
            assert_bits(ref.joints(), orc.joints(), "joints right after Destroy")
            assert_bits(ref.bodies(), orc.bodies(), "bodies right after Destroy")
            assert ref.joints()[1]["body_a"] == 0xFFFFFFFF
        if s == 120:
            b = ref.bodies()[ps[1]]
            jid = ref.add_revolute(hub, ps[1], (float(b["xf_px"]), float(b["xf_py"]) + 0.25))
            assert jid == 1  # the freed id comes back (joint pool free list)
            row = ref.joints()[1:2]
            orc.create_joints(row)
            for t in twins:
                t.create_joints(row)
            assert_bits(ref.joints(), orc.joints(), "joints right after the new CreateJoint")
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "step %d stats (reference vs restatement)" % s)
        assert_bits(ref.bodies(), orc.bodies(), "step %d bodies (reference vs restatement)" % s)
        assert_bits(ref.joints(), orc.joints(), "step %d joints (reference vs restatement)" % s)
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "step %d contacts (reference vs restatement)" % s)
        assert_bits(rm, om, "step %d manifolds (reference vs restatement)" % s)
        if twins:
            gw, orc2 = twins
            gs = gw.step(conf)
            o2 = orc2.step(conf, order=gw.solve_order()[0])
            gsa, osa = np.array([gs]), np.array([o2])
            assert_bits(gsa, osa, "step %d stats" % s)
            assert_bits(gw.bodies(), orc2.bodies(), "step %d bodies" % s)
            assert_bits(gw.joints(), orc2.joints(), "step %d joints" % s)
            gc, gm = gw.contacts(); oc2, om2 = orc2.contacts()
            assert_bits(gc, oc2, "step %d contacts" % s)
            assert_bits(gm, om2, "step %d manifolds" % s)
    b = ref.bodies()
    assert b[ps[1]]["xf_py"] < 2.0 and b[ps[0]]["xf_py"] > 4.0   # the freed box is down, the others still hang
    return twins[0] if twins else None


@need_ref
@need_oracle
def test_restatement_destroy_joint_equals_reference():
    _run(with_device=False)


@need_ref
@need_oracle
@need_emul
def test_emulated_destroy_joint_equals_oracle():
    gw = _run(lib=emul_lib())
    ids = np.array([7], np.uint32)
    assert gw.lib.pr_destroy_joints(gw.h, 1, abi.ptr(ids)) == -2
    ids = np.array([0], np.uint32)
    assert gw.lib.pr_destroy_joints(gw.h, 1, abi.ptr(ids)) == 0
    assert gw.lib.pr_destroy_joints(gw.h, 1, abi.ptr(ids)) == 0   # destroying a destroyed joint does nothing
    j = gw.joints()
    assert j[0]["body_a"] == 0xFFFFFFFF and gw.lib.pr_set_revolute_joints(gw.h, 1, abi.ptr(ids), abi.ptr(j[0:1].copy())) == -1


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_destroy_joint_equals_oracle():
    _run(device=0)
