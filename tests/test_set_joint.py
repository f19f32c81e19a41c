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
    j2 = j.copy(); j2["body_b"] = j2["body_b"] + 1
    assert gw.lib.pr_set_revolute_joints(gw.h, 1, abi.ptr(ids), abi.ptr(j2)) == -6
    j3 = j.copy(); j3["collide_connected"] = 1 - j3["collide_connected"]
    assert gw.lib.pr_set_revolute_joints(gw.h, 1, abi.ptr(ids), abi.ptr(j3)) == -6
    assert_bits(gw.joints(), j, "joints after refused edits")


@pytest.mark.gpu
@need_oracle
@need_ref
def test_gpu_set_joint_equals_oracle():
    _device_vs_oracle(device=0)
