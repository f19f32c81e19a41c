"""Parity tests proper: the CUDA path (through the C ABI) against the oracle restatement fed the
device's colour-sorted constraint order, and against the real reference's stage functions when
oracle/_ref is present. Bit-exact everywhere (integer/index work AND floats: --fmad=false build)."""
import numpy as np
import pytest
import refworld
import scenes_np
from common import need_oracle, need_ref, run_vs_oracle, assert_bits
from playrho_b200 import abi, world as W

pytestmark = pytest.mark.gpu


@need_oracle
@pytest.mark.parametrize("name,builder,steps,every", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120, 1),
    ("pyramid6", lambda: scenes_np.pyramid(6), 150, 1),
    ("pyramid20", lambda: scenes_np.pyramid(20), 200, 5),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150, 1),
    ("disks2000", lambda: scenes_np.disks_box(2000, seed=12345), 120, 10),
    ("mixed1500", lambda: scenes_np.mixed(1500, seed=12345, tile_segments=64, tile_width=32.0), 100, 10),
])
def test_cuda_step_bit_exact_vs_oracle(name, builder, steps, every):
    gs, ncontacts, prof = run_vs_oracle(builder(), steps, device=0, check_every=every)
    assert ncontacts > 0
    assert prof["kernel_launches"] > 0


@need_oracle
@pytest.mark.parametrize("name,builder,steps,every", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120, 1),
    ("pyramid10", lambda: scenes_np.pyramid(10), 200, 1),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150, 1),
    ("mixed1500", lambda: scenes_np.mixed(1500, seed=12345, tile_segments=64, tile_width=32.0), 100, 5),
])
def test_cuda_step_with_toi_bit_exact_vs_oracle(name, builder, steps, every):
    """Default StepConf (doToi ON), TOI events in the reference's serial order (PR_TOI_SERIAL_EXACT):
    regular + TOI phase, every state array and every StepStats field bit-exact."""
    gs, ncontacts, prof = run_vs_oracle(builder(), steps, device=0, check_every=every, conf=refworld.default_step_conf())
    assert ncontacts > 0


@need_oracle
@pytest.mark.parametrize("name,builder,steps,every", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120, 1),
    ("pyramid10", lambda: scenes_np.pyramid(10), 200, 1),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150, 1),
    ("mixed1500", lambda: scenes_np.mixed(1500, seed=12345, tile_segments=64, tile_width=32.0), 100, 5),
    ("mixed20000", lambda: scenes_np.mixed_fast(20000, seed=7), 80, 10),
])
def test_cuda_parallel_toi_mode_bit_exact_vs_oracle(name, builder, steps, every):
    """PR_TOI_PARALLEL -- the product default and the mode bench.py times -- with the default StepConf: the oracle derives
    the rounds by the documented rule (oracle/restate/orc_toi_phase.inc solve_toi_parallel: every locally-soonest event with
    a disjoint body neighbourhood runs in the same round) and must agree with the device on the event schedule of EVERY
    step (pr_get_toi_schedule), and bit for bit on bodies, contacts, manifolds, proxies and every StepStats field."""
    run_vs_oracle.toi_events = 0
    gs, ncontacts, prof = run_vs_oracle(builder(), steps, device=0, check_every=every, conf=refworld.default_step_conf(), toi_mode=1)
    assert ncontacts > 0
    if name.startswith("mixed"):
        assert run_vs_oracle.toi_events > 0


@need_oracle
@need_ref
def test_cuda_parallel_toi_mode_shooter_bit_exact_vs_oracle():
    """The TOI stress scene (bullets, spinning kinematic bar, sensor) in PR_TOI_PARALLEL: rounds with several events."""
    import scenes
    run_vs_oracle.toi_events = 0; run_vs_oracle.toi_multi_rounds = 0
    run_vs_oracle(scenes.shooter().export_scene(), 300, device=0, conf=refworld.default_step_conf(), toi_mode=1)
    assert run_vs_oracle.toi_events > 300
    assert run_vs_oracle.toi_multi_rounds > 0


@need_oracle
@pytest.mark.parametrize("toi_mode", [0, 1])
def test_cuda_config2_full_size_pyramid40_bit_exact_vs_oracle(toi_mode):
    """BASELINE config 2 at its stated size: the 40-row pyramid (820 boxes on edge ground), 600 steps, default StepConf with
    TOI on, in both TOI modes."""
    gs, ncontacts, prof = run_vs_oracle(scenes_np.pyramid(40), 600, device=0, check_every=20, conf=refworld.default_step_conf(),
                                        toi_mode=toi_mode)
    assert ncontacts > 2000


@need_oracle
def test_cuda_config3_full_size_100k_disks_bit_exact_vs_oracle():
    """BASELINE config 3 at its stated size: 100k disks dropped into the static 4-edge container, default StepConf (TOI on,
    the benchmarked parallel mode), 60 steps from the drop: pair set / contact ids / manifolds / bodies / stats bit-exact."""
    gs, ncontacts, prof = run_vs_oracle(scenes_np.disks_box(100000, seed=12345), 60, device=0, check_every=15,
                                        conf=refworld.default_step_conf(), toi_mode=1)
    assert ncontacts > 100000


@need_oracle
def test_cuda_config4_full_size_1m_mixed_bit_exact_vs_oracle():
    """BASELINE config 4 at its stated size: 1M mixed disks / 4-8-gons dropped on the tiled chain terrain, default StepConf, TOI on
    in the benchmarked parallel mode, the first 64 steps (the lattice lands and starts to pile up: ~1.4M contacts, TOI events
    against the terrain): bodies, contacts, manifolds, proxies, every StepStats field and the TOI schedule bit-exact."""
    sc = scenes_np.mixed_fast(1000000, seed=12345)
    nb = len(sc["bodies"])
    run_vs_oracle.toi_events = 0
    gs, ncontacts, prof = run_vs_oracle(sc, 64, device=0, check_every=16, conf=refworld.default_step_conf(), toi_mode=1,
                                        world_kwargs=dict(body_capacity=nb + 16, proxy_capacity=nb + 8192, contact_capacity=6 * nb))
    assert ncontacts > 1000000
    assert run_vs_oracle.toi_events > 100


@need_oracle
def test_cuda_binned_narrowphase_variant_bit_exact_vs_oracle():
    """The opt-in kernel variants (narrowphase binned by shape-pair class; the persistent cooperative solver kernel with
    grid barriers between colours) on small scenes: same bits as the oracle, incl. joints and the colour-overflow tail."""
    sc = scenes_np.mixed(1500, seed=12345, tile_segments=64, tile_width=32.0)
    gs, ncontacts, prof = run_vs_oracle(sc, 100, device=0, check_every=5, conf=refworld.default_step_conf(), large_variants=True)
    assert ncontacts > 0
    gs, ncontacts, prof = run_vs_oracle(scenes_np.pyramid(10), 120, device=0, check_every=5, large_variants=True)
    assert ncontacts > 0
    conf = scenes_np.tumbler_step_conf(refworld.default_step_conf())
    gs, ncontacts, prof = run_vs_oracle(scenes_np.tumbler(150), 150, device=0, check_every=10, conf=conf, large_variants=True)
    assert ncontacts > 0


@need_oracle
def test_cuda_tma_staged_narrowphase_variant_bit_exact_vs_oracle():
    """k_update_contacts_tma (debug bit 6): the regular phase's narrowphase with the pair's vertex sets brought into shared
    memory by the bulk-copy engine (cp.async.bulk + mbarrier) -- same bits as the oracle on polygon-heavy scenes."""
    sc = scenes_np.mixed(1500, seed=12345, tile_segments=64, tile_width=32.0)
    gs, ncontacts, prof = run_vs_oracle(sc, 100, device=0, check_every=5, conf=refworld.default_step_conf(), toi_mode=1, world_kwargs=dict(debug_flags=64))
    assert ncontacts > 0
    gs, ncontacts, prof = run_vs_oracle(scenes_np.pyramid(10), 120, device=0, check_every=5, world_kwargs=dict(debug_flags=64))
    assert ncontacts > 0


@need_oracle
@need_ref
def test_cuda_shooter_toi_bit_exact_vs_oracle():
    """Fast bodies, bullets, spinning kinematic bar, sensor shape: hundreds of TOI events (SolveToi mini
    islands, GetToiViaSat, GJK TestOverlap for the sensor)."""
    import scenes
    sc = scenes.shooter().export_scene()
    conf = refworld.default_step_conf()
    gw = W.World(sc, device=0, record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    events, sensor_touch = 0, 0
    for s in range(300):
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        events += int(gs["toi"]["contacts_found"])
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert_bits(gc, oc, "step %d contacts" % s)
        assert_bits(gm, om, "step %d manifolds" % s)
        sensor_touch += int(((gc["flags"] & abi.CONTACT_SENSOR) != 0).sum()) if hasattr(abi, "CONTACT_SENSOR") else 0
    assert events > 300


@need_oracle
@need_ref
def test_cuda_tumbler_revolute_joint_bit_exact_vs_oracle():
    """Config 5 unit (Benchmark/BenchmarkMain.cpp:2614-2722): motorised revolute joint (RevoluteJointConf.cpp
    :127-435) on a 4-box enclosure with 120 squares, default StepConf; joint impulses / limit state included."""
    import scenes
    sc = scenes.tumbler(120).export_scene()
    conf = refworld.default_step_conf()
    gw = W.World(sc, device=0, record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    for s in range(300):
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        if s % 5 and s != 299:
            continue
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
        assert_bits(gw.joints(), orc.joints(), "step %d joints" % s)
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert_bits(gc, oc, "step %d contacts" % s)
        assert_bits(gm, om, "step %d manifolds" % s)
    assert abs(float(gw.joints()[0]["angular_motor_impulse"])) > 0


@need_ref
def test_cuda_toi_batch_bit_exact_vs_reference():
    """GetToiViaSat (TimeOfImpact.cpp:35) on 200k random child pairs and sweeps vs the real reference."""
    import scenes
    from test_oracle_pin import random_sweeps
    rng = np.random.RandomState(21)
    sc = scenes.mixed(80, seed=5).export_scene()
    nchild = len(sc["children"])
    n = 200000
    a = rng.randint(0, nchild, n).astype(np.uint32)
    b = rng.randint(0, nchild, n).astype(np.uint32)
    sa, sb = random_sweeps(rng, n), random_sweeps(rng, n)
    sb[: n // 3, 0] += 3.0; sb[: n // 3, 3] -= 3.0
    conf = refworld.default_step_conf()
    rt = np.zeros(n, np.float32); rs = np.zeros(n, np.int32); ri = np.zeros((n, 3), np.uint32)
    refworld._ref_lib().ref_toi(n, abi.ptr(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]), abi.ptr(a), abi.ptr(sa),
                                abi.ptr(b), abi.ptr(sb), abi.ptr(conf), abi.ptr(rt), abi.ptr(rs), abi.ptr(ri))
    gt, gs, gi = W.toi_batch(sc["children"], sc["verts"], sc["normals"], a, sa, b, sb, conf, device=0)
    assert (rs == 1).sum() > 5000
    assert np.array_equal(rs, gs)
    assert_bits(rt, gt, "toi time")
    assert np.array_equal(ri, gi)


@need_oracle
def test_cuda_parallel_toi_mode_drift_report():
    """PR_TOI_PARALLEL (product default) vs PR_TOI_SERIAL_EXACT on the same scene: same number of bodies
    asleep/awake is not required; trajectories may drift (reported, not gated: BASELINE north_star).
    Gated here: finite state, no tunnelling through the terrain, and determinism of the parallel mode."""
    sc = scenes_np.mixed(1500, seed=12345, tile_segments=64, tile_width=32.0)
    conf = W.default_step_conf()
    runs = []
    for mode in (0, 1, 1):
        gw = W.World(sc, device=0)
        gw.set_toi_mode(mode)
        events = rounds = 0
        for _ in range(150):
            st = gw.step(conf)
            p = gw.profile()
            events += int(p["toi_events"]); rounds += int(p["toi_rounds"])
        runs.append((gw.bodies(), events, rounds))
    (b0, e0, r0), (b1, e1, r1), (b2, e2, r2) = runs
    assert_bits(b1, b2, "parallel TOI mode is deterministic")
    dyn = b0["inv_mass"] > 0
    drift = np.hypot(b0["pos1_x"][dyn] - b1["pos1_x"][dyn], b0["pos1_y"][dyn] - b1["pos1_y"][dyn])
    print("TOI drift parallel vs serial-exact after 150 steps: max %.4g m, mean %.4g m; events %d vs %d; rounds %d vs %d"
          % (drift.max(), drift.mean(), e0, e1, r0, r1))
    assert np.isfinite(b1["pos1_x"]).all() and np.isfinite(b1["pos1_y"]).all()
    assert (b1["pos1_y"][dyn] > -0.5).all(), "a body tunnelled through the terrain"
    assert e0 > 0 and e1 > 0 and r1 <= r0


@need_ref
def test_cuda_collide_shapes_bit_exact_vs_reference():
    """CollideShapes (Manifold.cpp:393) on random child pairs vs the real reference's function."""
    import scenes
    rng = np.random.RandomState(3)
    sc = scenes.mixed(80, seed=5).export_scene()
    nchild = len(sc["children"])
    n = 200000
    a = rng.randint(0, nchild, n).astype(np.uint32)
    b = rng.randint(0, nchild, n).astype(np.uint32)

    def xfs():
        ang = rng.uniform(-3.2, 3.2, n).astype(np.float32)
        cs = np.zeros((n, 2), np.float32)
        refworld._ref_lib().ref_unitvec_from_angle(n, abi.ptr(ang), abi.ptr(cs))
        return np.concatenate([rng.uniform(-0.6, 0.6, (n, 2)).astype(np.float32), cs], 1).copy()
    xa, xb = xfs(), xfs()
    out_r = np.zeros(n, abi.MANIFOLD)
    refworld._ref_lib().ref_collide_shapes(n, abi.ptr(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]),
                                           abi.ptr(a), abi.ptr(xa), abi.ptr(b), abi.ptr(xb), 0.005, 10.0, abi.ptr(out_r))
    out_g = W.collide_shapes(sc["children"], sc["verts"], sc["normals"], a, xa, b, xb, device=0)
    assert (out_r["count"] > 0).sum() > 5000
    assert_bits(out_g, out_r, "manifolds")


def test_cuda_sincos_matches_libm_via_body_transforms():
    """The device glibc-sincosf restatement: body transforms written back by Step must equal libm's
    cos/sin of the body angle bit-for-bit (UnitVec::Get, UnitVec.cpp:27-30)."""
    sc = scenes_np.mixed(400, seed=2, tile_segments=32, tile_width=16.0)
    rng = np.random.RandomState(0)
    sc["bodies"]["vel_w"][:] = np.where(sc["bodies"]["inv_mass"] > 0, rng.uniform(-30, 30, len(sc["bodies"])), 0).astype(np.float32)
    gw = W.World(sc, device=0)
    conf = W.default_step_conf(do_toi=0)
    for _ in range(20):
        gw.step(conf)
    b = gw.bodies()
    dyn = b["inv_mass"] > 0
    # glibc's cosf/sinf (NOT numpy's SIMD float32 loops, which round differently)
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = libm.sinf.restype = ctypes.c_float
    libm.cosf.argtypes = libm.sinf.argtypes = [ctypes.c_float]
    ang = b["pos1_a"][dyn]
    assert len(np.unique(ang)) > 100
    cos = np.array([libm.cosf(float(a)) for a in ang], np.float32)
    sin = np.array([libm.sinf(float(a)) for a in ang], np.float32)
    assert_bits(cos, b["xf_qc"][dyn], "cos")
    assert_bits(sin, b["xf_qs"][dyn], "sin")


def test_cuda_large_scene_invariants():
    """Size-independent properties at 100k bodies: finite state, contact keys strictly increasing
    inside every AddContacts batch is not observable, but keys must be unique, A<B by proxy id,
    every contact's fat boxes overlap, and the clone steps identically (determinism)."""
    sc = scenes_np.disks_box(100000, seed=12345)
    gw = W.World(sc, device=0)
    conf = W.default_step_conf(do_toi=0)
    for _ in range(30):
        gw.step(conf)
    other = gw.clone()
    for _ in range(5):
        s1 = gw.step(conf); s2 = other.step(conf)
        assert_bits(np.array([s1]), np.array([s2]), "stats of clone")
    b1, b2 = gw.bodies(), other.bodies()
    assert_bits(b1, b2, "clone bodies")
    for f in ("pos1_x", "pos1_y", "pos1_a", "vel_x", "vel_y", "vel_w"):
        assert np.isfinite(b1[f]).all()
    # DestroyContacts runs in the PRE-phase (AabbTreeWorld.cpp:2186): a deltaTime=0 step runs only that
    # phase (:2205), after which every surviving contact's fat leaf boxes must overlap
    gw.step(W.default_step_conf(do_toi=0, delta_time=0.0))
    cs, ms = gw.contacts()
    keys = (cs["key_min"].astype(np.uint64) << np.uint64(32)) | cs["key_max"].astype(np.uint64)
    assert len(np.unique(keys)) == len(keys)
    assert (cs["key_min"] < cs["key_max"]).all()
    px = gw.proxies()
    by_tree = {int(t): i for i, t in enumerate(px["tree_id"])}
    ia = np.array([by_tree[int(t)] for t in cs["key_min"]]); ib = np.array([by_tree[int(t)] for t in cs["key_max"]])
    ov = (np.minimum(px["max_x"][ia], px["max_x"][ib]) >= np.maximum(px["min_x"][ia], px["min_x"][ib])) & \
         (np.minimum(px["max_y"][ia], px["max_y"][ib]) >= np.maximum(px["min_y"][ia], px["min_y"][ib]))
    assert ov.all()
    assert len(cs) > 100000


def test_cuda_full_size_mixed_world_invariants():
    """BASELINE config 4 at FULL size (1M mixed disks / 4-8-gons over chain terrain, TOI on, parallel event mode): the
    size-independent properties of the step. (1) determinism: a clone steps to identical bits (stats, bodies, contacts,
    manifolds) through the batched TOI rounds; (2) the contact set is exactly the broadphase's: keys unique, A < B, every
    contact's fat boxes overlap, and for a random sample of proxies EVERY other proxy whose fat box overlaps theirs (other
    body, either accelerable) has its contact -- completeness of the LBVH query against brute force; (3) touching <=> a
    manifold with points; (4) nothing re-allocates while stepping a world created with its capacities."""
    n = 1000000
    sc = scenes_np.mixed_fast(n, seed=12345)
    nb = len(sc["bodies"])
    gw = W.World(sc, device=0, body_capacity=nb + 16, proxy_capacity=nb + 8192, contact_capacity=6 * nb)
    conf = W.default_step_conf()
    events = 0
    for s in range(120):
        st = gw.step(conf)
        events += int(st["toi"]["contacts_found"])
        if s >= 60:
            assert int(gw.profile()["grow_events"]) == 0, "step %d re-allocated a device buffer" % s
    assert events > 0
    other = gw.clone()
    for _ in range(3):
        s1 = gw.step(conf); s2 = other.step(conf)
        assert_bits(np.array([s1]), np.array([s2]), "stats of clone")
    b1 = gw.bodies()
    assert_bits(b1, other.bodies(), "clone bodies")
    c1, m1 = gw.contacts(); c2, m2 = other.contacts()
    assert_bits(c1, c2, "clone contacts"); assert_bits(m1, m2, "clone manifolds")
    for f in ("pos1_x", "pos1_y", "pos1_a", "vel_x", "vel_y", "vel_w"):
        assert np.isfinite(b1[f]).all()
    gw.step(W.default_step_conf(delta_time=0.0))  # pre-phase only: DestroyContacts + FindContacts (AabbTreeWorld.cpp:2186-2205)
    cs, ms = gw.contacts()
    assert len(cs) > n
    keys = (cs["key_min"].astype(np.uint64) << np.uint64(32)) | cs["key_max"].astype(np.uint64)
    assert len(np.unique(keys)) == len(keys)
    assert (cs["key_min"] < cs["key_max"]).all()
    px = gw.proxies()
    slot = np.full(int(px["tree_id"].max()) + 1, -1, np.int64)
    slot[px["tree_id"]] = np.arange(len(px))
    ia, ib = slot[cs["key_min"]], slot[cs["key_max"]]
    assert (ia >= 0).all() and (ib >= 0).all()
    ov = (np.minimum(px["max_x"][ia], px["max_x"][ib]) >= np.maximum(px["min_x"][ia], px["min_x"][ib])) & \
         (np.minimum(px["max_y"][ia], px["max_y"][ib]) >= np.maximum(px["min_y"][ia], px["min_y"][ib]))
    assert ov.all()
    touching = (cs["flags"] & 1) != 0
    sensor = (cs["flags"] & 0x20) != 0
    assert np.array_equal(touching[~sensor], ms["count"][~sensor] > 0)
    # completeness against brute force for a sample of proxies of dynamic bodies
    rows = gw.bodies()
    accel = (rows["flags"] & abi.BODY_ACCELERATION) != 0 if hasattr(abi, "BODY_ACCELERATION") else rows["inv_mass"] > 0
    rng = np.random.default_rng(7)
    dyn_px = np.nonzero(accel[px["body"]])[0]
    have = set(keys.tolist())
    checked = 0
    for q in rng.choice(dyn_px, 200, replace=False):
        hit = (np.minimum(px["max_x"], px["max_x"][q]) >= np.maximum(px["min_x"], px["min_x"][q])) & \
              (np.minimum(px["max_y"], px["max_y"][q]) >= np.maximum(px["min_y"], px["min_y"][q])) & (px["body"] != px["body"][q])
        tq = int(px["tree_id"][q])
        for o in np.nonzero(hit)[0]:
            to = int(px["tree_id"][o])
            k = (min(tq, to) << 32) | max(tq, to)
            assert k in have, "overlapping fat boxes of proxies %d and %d have no contact" % (q, o)
            checked += 1
    assert checked > 400


def test_cuda_column_accessors_equal_row_accessors():
    from test_emul_pipeline import _column_accessors_check
    _column_accessors_check(None, 0)
