"""FrictionJointConf and MotorJointConf on the device (FrictionJointConf.cpp:76-239, MotorJointConf.cpp:77-235):
a top-down world without gravity -- pucks slowed down by friction joints to a static ground body (one saturating its force cap,
one not), a body driven to an offset pose by a motor joint (reaching it) and one whose motor is too weak -- plus contacts between
the pucks. Reference == restatement == emulated kernels == GPU, bit for bit, incl. the joints' accumulated impulses."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import world as W


def _build():
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 0.0))
    box = ref.add_box(0.5, 0.3, density=2.0)
    disk = ref.add_disk(0.4, density=1.0)
    p1 = ref.add_body(DYNAMIC, pos=(-4.0, 0.0), vel=(6.0, 0.5), w=3.0, shapes=[disk])
    p2 = ref.add_body(DYNAMIC, pos=(-4.0, 2.0), vel=(1.0, 0.0), w=0.5, shapes=[box])
    p3 = ref.add_body(DYNAMIC, pos=(2.0, 0.1), vel=(-3.0, 0.0), shapes=[disk])        # runs into p1
    m1 = ref.add_body(DYNAMIC, pos=(0.0, -3.0), shapes=[box])
    m2 = ref.add_body(DYNAMIC, pos=(4.0, -3.0), shapes=[box])
    ref.add_friction(ground, p1, (-4.0, 0.0), 2.0, 0.5)      # saturates: the puck is fast
    ref.add_friction(ground, p2, (-4.0, 2.0), 50.0, 20.0)    # strong: stops the box at once
    ref.add_friction(p1, p3, (-1.0, 0.0), 1.0, 0.2)          # between two dynamic bodies
    ref.add_motor(ground, m1, (1.5, -2.0), 0.8, 40.0, 30.0)            # drives m1 to (1.5, -2) at 0.8 rad
    ref.add_motor(ground, m2, (4.0, 1.0), -0.5, 0.5, 0.05, 0.5)        # too weak to get there soon
    return ref, (p1, p2, p3, m1, m2)


def _run(lib=None, device=-1, with_device=True):
    ref, bodies = _build()
    sc = ref.export_scene()
    assert [int(j["collide_connected"]) >> 8 for j in sc["joints"]] == [3, 3, 3, 4, 4]
    orc = refworld.OracleWorld(sc)
    twins = []
    if with_device:
        gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
        gw.set_toi_mode(0)
        twins = [gw, refworld.OracleWorld(sc)]
    conf = refworld.default_step_conf()
    for s in range(240):
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
            assert_bits(np.array([gs]), np.array([o2]), "step %d stats" % s)
            assert_bits(gw.bodies(), orc2.bodies(), "step %d bodies" % s)
            assert_bits(gw.joints(), orc2.joints(), "step %d joints" % s)
            gc, gm = gw.contacts(); oc2, om2 = orc2.contacts()
            assert_bits(gc, oc2, "step %d contacts" % s)
            assert_bits(gm, om2, "step %d manifolds" % s)
    b = ref.bodies()
    p1, p2, p3, m1, m2 = bodies
    assert abs(float(b[p2]["vel_x"])) < 1e-3                                              # the strong friction joint stopped the box
    assert abs(float(b[m1]["xf_px"]) - 1.5) < 0.05 and abs(float(b[m1]["xf_py"]) + 2.0) < 0.05   # the motor joint got m1 to its offset
    assert abs(float(b[m2]["xf_py"]) - 1.0) > 0.5                                         # the weak one is still on its way


@need_ref
@need_oracle
def test_restatement_friction_motor_joints_equal_reference():
    _run(with_device=False)


@need_ref
@need_oracle
@need_emul
def test_emulated_friction_motor_joints_equal_oracle():
    _run(lib=emul_lib())


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_friction_motor_joints_equal_oracle():
    _run(device=0)
