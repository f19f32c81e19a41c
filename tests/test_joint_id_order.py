"""The order a body's joints are freed in (Destroy(BodyID), AabbTreeWorld.cpp:1071-1085: m_bodyJoints order, NOT id order) decides which ids the
next CreateJoint calls hand out (ObjectPool: last freed first). A body whose joint list is out of id order -- a freed id re-used for a later
joint, a SetJoint that moved a joint to the end of the list -- is destroyed; the joints created afterwards must get the reference's ids.
Reference == restatement == emulated kernels == GPU: joint tables, bodies, and a few steps afterwards."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi, world as W

G = (0.0, -10.0)


def _run(lib=None, device=-1, with_device=True):
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(10.0, 0.5, center=(0.0, -0.5))])
    box = ref.add_box(0.3, 0.3, density=1.0)
    b = [ref.add_body(DYNAMIC, acc=G, pos=(-3.0 + 1.5 * i, 2.0), shapes=[box]) for i in range(5)]
    ref.add_revolute(b[0], b[1], (-2.25, 2.0))        # joint 0
    ref.add_revolute(b[0], b[2], (-1.5, 2.0))         # joint 1
    ref.add_distance(b[0], b[3], (-3.0, 2.0), (1.5, 2.0))   # joint 2
    sc = ref.export_scene()
    worlds = [refworld.OracleWorld(sc)]
    if with_device:
        gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
        gw.set_toi_mode(0)
        worlds.append(gw)
    conf = refworld.default_step_conf()

    def all_step(n):
        for _ in range(n):
            ref.step(conf)
            for w in worlds:
                if isinstance(w, refworld.OracleWorld):
                    w.step(conf)
                else:
                    w.step(conf)

    def check(where):
        for w in worlds:
            assert_bits(ref.joints(), w.joints(), where + " joints")
            la, lb = ref.bodies(), w.bodies()
            dead = (la["flags"] & np.uint32(abi.BODY_DESTROYED)) != 0   # (a destroyed body's row is not the same zeroed value on both sides)
            assert np.array_equal(dead, (lb["flags"] & np.uint32(abi.BODY_DESTROYED)) != 0), where + ": destroyed sets differ"
            assert_bits(la[~dead], lb[~dead], where + " bodies")

    def mirror_new_joint(jid):
        row = ref.joints()[[jid]]
        for w in worlds:
            w.create_joints(row)

    all_step(3)
    # free joint 0, then create a joint of b0 that re-uses id 0: b0's list is now [1, 2, 0]
    ref.destroy_joint(0)
    for w in worlds:
        w.destroy_joints([0])
    j = ref.add_revolute(b[0], b[4], (0.0, 2.0))
    assert j == 0
    mirror_new_joint(0)
    check("after the id re-use")
    # SetJoint on joint 1 (same value): Remove + Add moves it to the end of b0's list: [2, 0, 1]
    rows = ref.joints()[[1]]
    ref.set_joints([1], rows)
    for w in worlds:
        w.set_joints([1], rows)
    all_step(2)
    check("after SetJoint")
    # destroy b0: its joints go in list order 2, 0, 1 -> free stack [2, 0, 1]: the next three joints get ids 1, 0, 2
    ref.destroy_body(b[0])
    for w in worlds:
        w.destroy_bodies([b[0]])
    check("after Destroy(BodyID)")
    got = []
    for (x, y) in ((1, 2), (2, 3), (3, 4)):
        jid = ref.add_revolute(b[x], b[y], (-3.0 + 1.5 * x + 0.75, 2.0))
        got.append(jid)
        mirror_new_joint(jid)
    assert got == [1, 0, 2], got
    check("after the three new joints")
    for s in range(30):
        rs = ref.step(conf)
        for w in worlds:
            ws = w.step(conf) if isinstance(w, refworld.OracleWorld) else w.step(conf)
            if isinstance(w, refworld.OracleWorld):
                assert_bits(np.array([rs]), np.array([ws]), "step %d stats" % s)
        check("step %d" % s)


@need_ref
@need_oracle
def test_restatement_frees_a_bodys_joints_in_list_order_like_the_reference():
    _run(with_device=False)


@need_ref
@need_oracle
@need_emul
def test_emulated_kernels_free_a_bodys_joints_in_list_order_like_the_reference():
    _run(lib=emul_lib())


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_frees_a_bodys_joints_in_list_order_like_the_reference():
    _run(device=0)
