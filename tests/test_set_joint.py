"""SetJoint (AabbTreeWorld.cpp:1108-1130) for the revolute joint = pr_set_revolute_joints: the control input of the tumbler worlds
(motor speed / torque / enable, limits). Reference (its own SetJoint) == restatement == emulated kernels == GPU, bit for bit,
over a trajectory in which the motor is reversed, switched off and limited."""
import numpy as np
import pytest
import refworld
import scenes
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W


def _edit(j, step):
    r = j.copy()
    if step == 40:
        r["motor_speed"] = -2.0 * r["motor_speed"]
    elif step == 80:
        r["enable_motor"] = 0
    elif step == 100:
        r["enable_motor"] = 1; r["max_motor_torque"] = np.float32(50.0); r["motor_speed"] = np.float32(1.5)
    elif step == 120:
        r["enable_limit"] = 1; r["lower_angle"] = np.float32(-0.05); r["upper_angle"] = np.float32(0.4)
    return r


@need_ref
@need_oracle
def test_restatement_set_joint_equals_reference():
    ref = scenes.tumbler(40)
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    for s in range(160):
        if s in (40, 80, 100, 120):
            rows = _edit(ref.joints(), s)
            ref.set_joints([0], rows); orc.set_joints([0], rows)
            assert_bits(ref.joints(), orc.joints(), "joints right after SetJoint at %d" % s)
            assert_bits(ref.bodies(), orc.bodies(), "bodies right after SetJoint at %d" % s)
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "step %d stats" % s)
        assert_bits(ref.bodies(), orc.bodies(), "step %d bodies" % s)
        assert_bits(ref.joints(), orc.joints(), "step %d joints" % s)
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "step %d contacts" % s)
        assert_bits(rm, om, "step %d manifolds" % s)


def _device_vs_oracle(lib=None, device=-1):
    sc = scenes.tumbler(40).export_scene()
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    for s in range(160):
        if s in (40, 80, 100, 120):
            rows = _edit(gw.joints(), s)
            gw.set_joints([0], rows); orc.set_joints([0], rows)
            assert_bits(gw.joints(), orc.joints(), "joints right after SetJoint at %d" % s)
            assert_bits(gw.bodies(), orc.bodies(), "bodies right after SetJoint at %d" % s)
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
        assert_bits(gw.joints(), orc.joints(), "step %d joints" % s)
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert_bits(gc, oc, "step %d contacts" % s)
        assert_bits(gm, om, "step %d manifolds" % s)
    return gw


@need_oracle
@need_emul
@need_ref
def test_emulated_set_joint_equals_oracle():
    gw = _device_vs_oracle(lib=emul_lib())
    # refusals: unknown joint, a change of bodies / collideConnected
    import ctypes as C
    j = gw.joints().copy(); ids = np.array([0], np.uint32)
    bad = np.array([5], np.uint32)
    assert gw.lib.pr_set_revolute_joints(gw.h, 1, abi.ptr(bad), abi.ptr(j)) == -2
    j2 = j.copy(); j2["body_b"] = 100000
    assert gw.lib.pr_set_revolute_joints(gw.h, 1, abi.ptr(ids), abi.ptr(j2)) == -2      # no such body
    j3 = j.copy(); j3["collide_connected"] = (10 << 8) | (int(j3["collide_connected"][0]) & 1)
    assert gw.lib.pr_set_revolute_joints(gw.h, 1, abi.ptr(ids), abi.ptr(j3)) == -6      # to a gear joint: its value is wider than a row
    twice = np.array([0, 0], np.uint32); jj = np.concatenate([j, j])
    assert gw.lib.pr_set_revolute_joints(gw.h, 2, abi.ptr(twice), abi.ptr(jj)) == -1    # an id named twice
    assert_bits(gw.joints(), j, "joints after refused edits")


@pytest.mark.gpu
@need_oracle
@need_ref
def test_gpu_set_joint_equals_oracle():
    _device_vs_oracle(device=0)


# ---- SetJoint with a value that names other bodies / another collideConnected (Remove(id) + assign + Add(id, flag), :1108-1130):
# three overlapping boxes on a floor; the joint first ties box 0 to box 1 (no contact between them), is then moved to boxes 0 and 2
# (the pair 0-1 starts colliding, the contact 0-2 is flagged for filtering and destroyed), and finally told to let its bodies collide
def _rehook_scene():
    ref = refworld.RefWorld()
    G = (0.0, -10.0)
    ref.add_body(refworld.STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(10.0, 0.5, center=(0.0, -0.5))])
    box = ref.add_box(0.5, 0.5, density=1.0)
    b0 = ref.add_body(refworld.DYNAMIC, acc=G, pos=(0.0, 0.6), shapes=[box])
    b1 = ref.add_body(refworld.DYNAMIC, acc=G, pos=(0.7, 0.9), shapes=[box])
    b2 = ref.add_body(refworld.DYNAMIC, acc=G, pos=(-0.7, 0.9), shapes=[box])
    ref.add_revolute(b0, b1, (0.35, 0.75))
    return ref, (b0, b1, b2)


def _rehook(j, step, bodies):
    b0, b1, b2 = bodies
    r = j.copy()
    if step == 30:
        r["body_b"] = b2; r["local_anchor_b"] = (0.35, -0.15)
    elif step == 70:
        r["collide_connected"] = 1
    elif step == 110:
        r["body_a"] = b1; r["collide_connected"] = 0
    return r


def _rehook_run(lib=None, device=-1, with_device=True):
    ref, bodies = _rehook_scene()
    sc = ref.export_scene()
    orc = refworld.OracleWorld(sc)
    twins = []
    if with_device:
        gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
        gw.set_toi_mode(0)
        twins = [gw, refworld.OracleWorld(sc)]
    conf = refworld.default_step_conf()
    seen = set()
    for s in range(160):
        if s in (30, 70, 110):
            rows = _rehook(ref.joints(), s, bodies)
            ref.set_joints([0], rows); orc.set_joints([0], rows)
            assert_bits(ref.joints(), orc.joints(), "joints right after SetJoint at %d (reference vs restatement)" % s)
            assert_bits(ref.bodies(), orc.bodies(), "bodies right after SetJoint at %d (reference vs restatement)" % s)
            assert_bits(ref.contacts()[0], orc.contacts()[0], "contacts right after SetJoint at %d (reference vs restatement)" % s)
            for t in twins:
                t.set_joints([0], _rehook(t.joints(), s, bodies))
            if twins:
                assert_bits(twins[0].bodies(), twins[1].bodies(), "bodies right after SetJoint at %d" % s)
                assert_bits(twins[0].contacts()[0], twins[1].contacts()[0], "contacts right after SetJoint at %d" % s)
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "step %d stats (reference vs restatement)" % s)
        assert_bits(ref.bodies(), orc.bodies(), "step %d bodies (reference vs restatement)" % s)
        assert_bits(ref.joints(), orc.joints(), "step %d joints (reference vs restatement)" % s)
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "step %d contacts (reference vs restatement)" % s)
        assert_bits(rm, om, "step %d manifolds (reference vs restatement)" % s)
        seen |= {(int(a), int(b)) for a, b in zip(rc["body_a"], rc["body_b"])}
        if twins:
            gw, orc2 = twins
            gs = gw.step(conf)
            o2 = orc2.step(conf, order=gw.solve_order()[0])
            assert_bits(np.array([gs]), np.array([o2]), "step %d stats" % s)
            assert_bits(gw.bodies(), orc2.bodies(), "step %d bodies" % s)
            assert_bits(gw.joints(), orc2.joints(), "step %d joints" % s)
            gc, gm = gw.contacts(); oc2, om2 = orc2.contacts()
            assert_bits(gc, oc2, "step %d contacts" % s)
            assert_bits(gm, om2, "step %d manifolds" % s)
    b0, b1, b2 = bodies
    assert (b0, b1) in seen or (b1, b0) in seen     # the pair the joint let go of collided afterwards
    assert (b0, b2) in seen or (b2, b0) in seen


@need_ref
@need_oracle
def test_restatement_set_joint_to_other_bodies_equals_reference():
    _rehook_run(with_device=False)


@need_ref
@need_oracle
@need_emul
def test_emulated_set_joint_to_other_bodies_equals_oracle():
    _rehook_run(lib=emul_lib())


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_set_joint_to_other_bodies_equals_oracle():
    _rehook_run(device=0)
