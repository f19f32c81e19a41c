"""Pins the C++ restatement (oracle/restate) against the REAL reference (oracle/_ref, built from
/root/reference by oracle/Makefile) and against the reference's golden HelloWorld output."""
import os
import numpy as np
import pytest
import refworld
import scenes
from common import need_ref, need_oracle, assert_bits
from playrho_b200 import abi

GOLDEN = os.path.join(refworld.ROOT, "tests", "golden", "HelloWorld_ExpectedOutput.txt")


@need_ref
@need_oracle
def test_step_conf_defaults_match_reference():
    rc = abi.one(abi.STEP_CONF)
    refworld._ref_lib().ref_step_conf_default(abi.ptr(rc))
    assert_bits(rc, refworld.default_step_conf(), "StepConf defaults")


@need_ref
@need_oracle
@pytest.mark.parametrize("name,steps", [("hello", 60), ("pyramid5", 120), ("pyramid10", 200), ("disks100", 200), ("mixed100", 200)])
def test_restatement_bit_exact_vs_reference_reg_phase(name, steps):
    """Whole-trajectory bit equality (bodies, contacts, manifolds, proxies, StepStats), doToi off."""
    ref = scenes.SCENES[name]()
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf(do_toi=0)
    for s in range(steps):
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(ref.bodies(), orc.bodies(), "%s step %d bodies" % (name, s))
        rc, rm = ref.contacts()
        oc, om = orc.contacts()
        assert_bits(rc, oc, "%s step %d contacts" % (name, s))
        assert_bits(rm, om, "%s step %d manifolds" % (name, s))
        rp, op = ref.proxies(), orc.proxies()
        assert_bits(rp[np.argsort(rp["tree_id"])], op[np.argsort(op["tree_id"])], "%s step %d proxies" % (name, s))
        assert_bits(np.array([rs]), np.array([os_]), "%s step %d stats" % (name, s))


@need_ref
@need_oracle
@pytest.mark.parametrize("name,steps", [("hello", 60), ("pyramid5", 150), ("pyramid10", 200), ("disks100", 250), ("mixed100", 250), ("shooter", 300)])
def test_restatement_bit_exact_vs_reference_with_toi(name, steps):
    """Default StepConf (doToi on): SolveToi (AabbTreeWorld.cpp:1752-2142), GetToiViaSat, GJK Distance."""
    ref = scenes.SCENES[name]()
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    events = 0
    for s in range(steps):
        rs, os_ = ref.step(conf), orc.step(conf)
        events += int(rs["toi"]["contacts_found"])
        assert_bits(np.array([rs]), np.array([os_]), "%s step %d stats" % (name, s))
        assert_bits(ref.bodies(), orc.bodies(), "%s step %d bodies" % (name, s))
        rc, rm = ref.contacts()
        oc, om = orc.contacts()
        assert_bits(rc, oc, "%s step %d contacts" % (name, s))
        assert_bits(rm, om, "%s step %d manifolds" % (name, s))
        rp, op = ref.proxies(), orc.proxies()
        assert_bits(rp[np.argsort(rp["tree_id"])], op[np.argsort(op["tree_id"])], "%s step %d proxies" % (name, s))
    assert events > 0, "scene never produced a TOI event"


@need_ref
@need_oracle
def test_restatement_tumbler_revolute_joint_bit_exact_vs_reference():
    """Config 5 unit: motorised revolute joint (RevoluteJointConf.cpp:127-435) driving a 4-box enclosure full of
    squares, default StepConf (TOI on): bodies, contacts, manifolds, stats and the joint's solver-written
    fields (impulse, motor impulse, limit state) bit for bit over 200 steps."""
    ref = scenes.tumbler(40)
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    for s in range(200):
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "tumbler step %d stats" % s)
        assert_bits(ref.bodies(), orc.bodies(), "tumbler step %d bodies" % s)
        rc, rm = ref.contacts()
        oc, om = orc.contacts()
        assert_bits(rc, oc, "tumbler step %d contacts" % s)
        assert_bits(rm, om, "tumbler step %d manifolds" % s)
        assert_bits(ref.joints(), orc.joints(), "tumbler step %d joints" % s)
    assert len(rc) > 20


@need_ref
@need_oracle
def test_restatement_reproduces_golden_helloworld():
    """The restatement itself (default StepConf, TOI on) prints HelloWorld/ExpectedOutput.txt."""
    orc = refworld.OracleWorld(scenes.hello_world().export_scene())
    conf = refworld.default_step_conf()
    lines = []
    for _ in range(60):
        orc.step(conf)
        b = orc.bodies()[1]
        lines.append("%5.2f%5.2f%5.2f" % (b["xf_px"], b["xf_py"], np.arctan2(b["xf_qs"], b["xf_qc"])))
    assert lines == open(GOLDEN).read().splitlines()


@need_ref
def test_reference_reproduces_golden_helloworld():
    """The reference built by oracle/Makefile reproduces HelloWorld/ExpectedOutput.txt."""
    ref = scenes.hello_world()
    conf = refworld.default_step_conf()
    lines = []
    for _ in range(60):
        ref.step(conf)
        b = ref.bodies()[1]
        ang = np.arctan2(b["xf_qs"], b["xf_qc"])
        lines.append("%5.2f%5.2f%5.2f" % (b["xf_px"], b["xf_py"], ang))
    golden = open(GOLDEN).read().splitlines()
    assert lines == golden


@need_ref
@need_oracle
def test_collide_shapes_restatement_vs_reference_random():
    """CollideShapes (Manifold.cpp:393) on random child pairs: restatement == reference, bit for bit."""
    rng = np.random.RandomState(3)
    ref = scenes.mixed(60, seed=5)
    sc = ref.export_scene()
    nchild = len(sc["children"])
    n = 20000
    a = rng.randint(0, nchild, n).astype(np.uint32)
    b = rng.randint(0, nchild, n).astype(np.uint32)
    def xfs():
        ang = rng.uniform(-3.2, 3.2, n).astype(np.float32)
        cs = np.zeros((n, 2), np.float32)
        refworld._ref_lib().ref_unitvec_from_angle(n, abi.ptr(ang), abi.ptr(cs))
        return np.concatenate([rng.uniform(-0.6, 0.6, (n, 2)).astype(np.float32), cs], 1).copy()
    xa, xb = xfs(), xfs()
    # move child B near child A's first vertex so that many pairs actually touch
    va = np.stack([sc["verts"]["x"], sc["verts"]["y"]], 1)
    xa[:, :2] -= va[sc["children"]["vertex_offset"][a]] * 0.0
    out_r = np.zeros(n, abi.MANIFOLD)
    out_o = np.zeros(n, abi.MANIFOLD)
    refworld._ref_lib().ref_collide_shapes(n, abi.ptr(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]),
                                           abi.ptr(a), abi.ptr(xa), abi.ptr(b), abi.ptr(xb), 0.005, 10.0, abi.ptr(out_r))
    refworld._orc_lib().orc_collide_shapes(n, abi.ptr(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]),
                                           abi.ptr(a), abi.ptr(xa), abi.ptr(b), abi.ptr(xb), 0.005, 10.0, abi.ptr(out_o))
    assert (out_r["count"] > 0).sum() > 500
    assert_bits(out_r, out_o, "manifolds")


def random_sweeps(rng, n, spread=1.5, move=2.0, spin=3.0):
    """(n, 8) float32 sweeps: pos0 x,y,a ; pos1 x,y,a ; local centre x,y."""
    p0 = rng.uniform(-spread, spread, (n, 2)); a0 = rng.uniform(-3.2, 3.2, (n, 1))
    d = rng.uniform(-move, move, (n, 2)); da = rng.uniform(-spin, spin, (n, 1))
    lc = rng.uniform(-0.1, 0.1, (n, 2))
    return np.concatenate([p0, a0, p0 + d, a0 + da, lc], 1).astype(np.float32).copy()


@need_ref
@need_oracle
def test_toi_restatement_vs_reference_random():
    """GetToiViaSat (TimeOfImpact.cpp:35) incl. GJK Distance / Simplex / SeparationScenario on random
    child pairs and sweeps: time, state and the three iteration statistics bit-for-bit."""
    rng = np.random.RandomState(11)
    sc = scenes.mixed(60, seed=5).export_scene()
    nchild = len(sc["children"])
    n = 30000
    a = rng.randint(0, nchild, n).astype(np.uint32)
    b = rng.randint(0, nchild, n).astype(np.uint32)
    sa, sb = random_sweeps(rng, n), random_sweeps(rng, n)
    # a third of the pairs start far apart and fly towards each other
    sb[: n // 3, 0] += 3.0; sb[: n // 3, 3] -= 3.0
    conf = refworld.default_step_conf()
    outs = []
    for lib, fn in ((refworld._ref_lib(), "ref_toi"), (refworld._orc_lib(), "orc_toi")):
        t = np.zeros(n, np.float32); st = np.zeros(n, np.int32); it = np.zeros((n, 3), np.uint32)
        getattr(lib, fn)(n, abi.ptr(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]), abi.ptr(a), abi.ptr(sa),
                         abi.ptr(b), abi.ptr(sb), abi.ptr(conf), abi.ptr(t), abi.ptr(st), abi.ptr(it))
        outs.append((t, st, it))
    (rt, rs, ri), (ot, os_, oi) = outs
    states = set(int(s) for s in np.unique(rs))
    assert {1, 2, 3} <= states, states  # touching, separated, overlapped all occur
    assert (rs == 1).sum() > 1000
    assert np.array_equal(rs, os_), np.nonzero(rs != os_)[0][:10]
    assert_bits(rt, ot, "toi time")
    assert np.array_equal(ri, oi)
