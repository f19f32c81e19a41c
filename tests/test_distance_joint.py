"""DistanceJointConf and RopeJointConf on the device (the first two of the ten other joint types;
DistanceJointConf.cpp:81-246, RopeJointConf.cpp:58-214):
a rigid rod, a soft spring, a rod between two dynamic bodies and a revolute joint in one world -- the joint table holds both
types, PR_JOINT_TYPE says which row is which. Reference (its own CreateJoint / Destroy) == restatement == emulated kernels ==
GPU, bit for bit, incl. the joints' accumulated impulses."""
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
    disk = ref.add_disk(0.4, density=1.0)
    a = ref.add_body(DYNAMIC, pos=(-6.0, 6.0), vel=(2.0, 0.0), acc=(0.0, -9.8), shapes=[box])
    b = ref.add_body(DYNAMIC, pos=(-2.0, 6.0), vel=(1.0, 0.0), acc=(0.0, -9.8), shapes=[disk])
    c = ref.add_body(DYNAMIC, pos=(2.0, 6.0), vel=(-1.0, 0.0), acc=(0.0, -9.8), shapes=[box])
    e = ref.add_body(DYNAMIC, pos=(3.5, 5.0), vel=(0.0, 0.0), acc=(0.0, -9.8), shapes=[disk])
    f = ref.add_body(DYNAMIC, pos=(6.0, 6.0), vel=(3.0, 0.0), acc=(0.0, -9.8), shapes=[box])
    ref.add_distance(hub, a, (-6.0, 10.0), (-6.0, 6.2))                                   # rigid rod from the hub
    ref.add_distance(hub, b, (-2.0, 10.0), (-2.0, 6.0), frequency=2.0, damping_ratio=0.3)  # soft spring
    ref.add_distance(hub, c, (2.0, 10.0), (2.0, 6.2))
    ref.add_distance(c, e, (2.3, 6.0), (3.5, 5.0), collide_connected=True)                 # rod between two dynamic bodies
    ref.add_revolute(hub, f, (6.0, 9.0))                                                   # and a revolute joint in the same table
    r1 = ref.add_body(DYNAMIC, pos=(-10.0, 8.0), vel=(4.0, 0.0), acc=(0.0, -9.8), shapes=[disk])   # on a rope: slack at first, then taut
    ref.add_rope(hub, r1, (-10.0, 0.0), (0.0, 0.0), 3.0)
    g1 = ref.add_body(DYNAMIC, pos=(10.0, 0.31), acc=(0.0, -9.8), shapes=[box])            # two boxes stacked on the ground, touching
    g2 = ref.add_body(DYNAMIC, pos=(10.0, 0.92), acc=(0.0, -9.8), shapes=[box])
    ref.rope_body = r1
    return ref, hub, (a, b, c, e, f, g1, g2)


def _run(lib=None, device=-1, with_device=True):
    ref, hub, bodies = _build()
    sc = ref.export_scene()
    assert [int(j["collide_connected"]) >> 8 for j in sc["joints"]] == [1, 1, 1, 1, 0, 2]
    orc = refworld.OracleWorld(sc)
    twins = []
    if with_device:
        gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
        gw.set_toi_mode(0)
        orc2 = refworld.OracleWorld(sc)
        twins = [gw, orc2]
    conf = refworld.default_step_conf()
    for s in range(220):
        if s == 60:  # a joint with collideConnected == false between two bodies in contact: Add(JointID, true) flags their contact
            # for filtering (AabbTreeWorld.cpp:1159-1175) and the next Step's DestroyContacts drops it (:2305-2328)
            g1, g2 = bodies[5], bodies[6]
            def pair_contacts(w):
                c, _ = w.contacts()
                return int((((c["body_a"] == g1) & (c["body_b"] == g2)) | ((c["body_a"] == g2) & (c["body_b"] == g1))).sum())
            assert pair_contacts(ref) == 1
            jid = ref.add_distance(g1, g2, (10.0, 0.31), (10.0, 0.92))
            row = ref.joints()[jid:jid + 1]
            orc.create_joints(row)
            for t in twins:
                t.create_joints(row)
            c, _ = ref.contacts()
            sel = ((c["body_a"] == g1) & (c["body_b"] == g2)) | ((c["body_a"] == g2) & (c["body_b"] == g1))
            assert (c["flags"][sel] & abi.CONTACT_FILTER).all()
            filtered_at = s
        if s == 62:
            assert pair_contacts(ref) == 0 and all(pair_contacts(t) == 0 for t in twins)
        if s == 90:  # the spring is cut; later a new rod takes its id
            ref.destroy_joint(1); orc.destroy_joints([1])
            for t in twins:
                t.destroy_joints([1])
        if s == 150:
            bb = ref.bodies()[bodies[1]]
            jid = ref.add_distance(hub, bodies[1], (float(bb["xf_px"]), 10.0), (float(bb["xf_px"]), float(bb["xf_py"])))
            assert jid == 1
            row = ref.joints()[1:2]
            orc.create_joints(row)
            for t in twins:
                t.create_joints(row)
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
    assert b[bodies[0]]["xf_py"] > 5.0          # the rod holds its box up
    j = ref.joints()
    assert abs(float(j[0]["impulse_x"])) > 0.0  # and carries load
    assert float(j[5]["impulse_x"]) < 0.0 and int(j[5]["limit_state"]) in (0, 2)   # the rope went taut
    assert b[ref.rope_body]["xf_py"] > 6.0
    return twins[0] if twins else None


@need_ref
@need_oracle
def test_restatement_distance_joint_equals_reference():
    _run(with_device=False)


@need_ref
@need_oracle
@need_emul
def test_emulated_distance_joint_equals_oracle():
    _run(lib=emul_lib())


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_distance_joint_equals_oracle():
    _run(device=0)
