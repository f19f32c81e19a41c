"""CPU check of the DEVICE pipeline's logic: the same kernels (playrho_b200/csrc) compiled for the
host with -DPR_HOST_EMUL (tests/_emul, a tests-only artefact) are stepped against the oracle fed
the device's colour-sorted constraint order. Everything must be bit-exact. The real CUDA build is
checked the same way in test_gpu_parity.py."""
import numpy as np
import pytest
import refworld
import scenes_np
from common import need_oracle, need_emul, need_ref, emul_lib, run_vs_oracle
from playrho_b200 import world as W


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder,steps", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120),
    ("pyramid6", lambda: scenes_np.pyramid(6), 150),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150),
])
def test_emulated_device_pipeline_bit_exact(name, builder, steps):
    gs, ncontacts, prof = run_vs_oracle(builder(), steps, lib=emul_lib())
    assert ncontacts > 0


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder,steps", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120),
    ("pyramid6", lambda: scenes_np.pyramid(6), 150),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150),
])
def test_emulated_device_pipeline_with_toi_bit_exact(name, builder, steps):
    """Default StepConf (doToi on), serial-exact TOI event order: reg + TOI phases vs the oracle."""
    gs, ncontacts, prof = run_vs_oracle(builder(), steps, lib=emul_lib(), conf=refworld.default_step_conf())
    assert ncontacts > 0


@need_oracle
@need_emul
def test_emulated_binned_narrowphase_variant_bit_exact():
    """The kernel variants large worlds take (narrowphase binned by shape-pair class through a radix pass) forced on a
    small mixed scene: same bits as the oracle, TOI on."""
    sc = scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0)
    gs, ncontacts, prof = run_vs_oracle(sc, 150, lib=emul_lib(), conf=refworld.default_step_conf(), large_variants=True)
    assert ncontacts > 0


@need_oracle
@need_emul
@need_ref
def test_emulated_device_pipeline_shooter_toi_bit_exact():
    """Fast bodies, bullets, a spinning kinematic bar and a sensor: hundreds of TOI events."""
    import scenes
    sc = scenes.shooter().export_scene()
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=emul_lib(), record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    events = 0
    for s in range(200):
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        events += int(gs["toi"]["contacts_found"])
        gsa, osa = np.array([gs]), np.array([os_])
        assert refworld.bits_equal(gsa, osa), "step %d stats: %s" % (s, refworld.diff_fields(gsa, osa))
        assert refworld.bits_equal(gw.bodies(), orc.bodies()), "step %d bodies: %s" % (s, refworld.diff_fields(gw.bodies(), orc.bodies()))
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert refworld.bits_equal(gc, oc), "step %d contacts: %s" % (s, refworld.diff_fields(gc, oc))
        assert refworld.bits_equal(gm, om), "step %d manifolds: %s" % (s, refworld.diff_fields(gm, om))
    assert events > 100


@need_oracle
@need_emul
@need_ref
def test_emulated_tumbler_revolute_joint_bit_exact():
    """Config 5 unit on the device pipeline: motorised revolute joint + enclosure + squares, default StepConf."""
    import scenes
    sc = scenes.tumbler(40).export_scene()
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=emul_lib(), record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    for s in range(150):
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        gsa, osa = np.array([gs]), np.array([os_])
        assert refworld.bits_equal(gsa, osa), "step %d stats: %s" % (s, refworld.diff_fields(gsa, osa))
        assert refworld.bits_equal(gw.bodies(), orc.bodies()), "step %d bodies: %s" % (s, refworld.diff_fields(gw.bodies(), orc.bodies()))
        assert refworld.bits_equal(gw.joints(), orc.joints()), "step %d joints: %s" % (s, refworld.diff_fields(gw.joints(), orc.joints()))
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert refworld.bits_equal(gc, oc) and refworld.bits_equal(gm, om), "step %d contacts" % s


@need_oracle
@need_emul
@need_ref
def test_emulated_tumbler_colour_overflow_tail_bit_exact():
    """512 squares pile on the enclosure's floor: the enclosure body gathers more constraints than colours are worth (hub rule,
    World::colorCap: beyond 32 colours the colouring stops at 24 and the rest goes to the per-island serial tail, sorted by
    island, one thread per island segment) -- and the result stays bit-exact vs the oracle."""
    import scenes
    sc = scenes.tumbler(512).export_scene()
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=emul_lib(), record_solve_order=True)
    gw.lib.pr_set_debug(gw.h, 1 | 16)  # record the solve order + the (opt-in) hub rule
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    max_colors = last_colors = 0
    for s in range(170):
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        last_colors = int(gw.profile()["n_colors"])
        max_colors = max(max_colors, last_colors)
        if s % 10 and s != 169:
            continue
        assert refworld.bits_equal(gw.bodies(), orc.bodies()), "step %d bodies: %s" % (s, refworld.diff_fields(gw.bodies(), orc.bodies()))
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert refworld.bits_equal(gc, oc) and refworld.bits_equal(gm, om), "step %d contacts" % s
    assert max_colors > 32, max_colors      # the world did need more than 32 colours ...
    assert last_colors == 25, last_colors   # ... and ended on 24 colours + the tail


@need_oracle
@need_emul
def test_set_bodies_moves_proxies_and_flags_contacts():
    sc = scenes_np.disks_box(16, seed=3)
    lib = emul_lib()
    conf = refworld.default_step_conf(do_toi=0)
    gw = W.World(sc, lib=lib, record_solve_order=True)
    orc = refworld.OracleWorld(sc)
    for _ in range(30):
        gw.step(conf); orc.step(conf, order=gw.solve_order()[0])
    rows = gw.bodies()
    row = rows[5:6].copy()
    row["xf_px"] += 3.0; row["pos0_x"] += 3.0; row["pos1_x"] += 3.0
    ids = np.array([5], np.uint32)
    gw.set_bodies(ids, row)
    from playrho_b200 import abi
    refworld._orc_lib().orc_set_bodies(orc.h, 1, abi.ptr(ids), abi.ptr(row))
    for s in range(5):
        gs = gw.step(conf); os_ = orc.step(conf, order=gw.solve_order()[0])
        assert refworld.bits_equal(gw.bodies(), orc.bodies())
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert refworld.bits_equal(gc, oc) and refworld.bits_equal(gm, om)
        if s == 0:
            assert gs["pre"]["proxies_moved"] == os_["pre"]["proxies_moved"] == 1


def _column_accessors_check(lib, device):
    """pr_set_body_accelerations / pr_get_body_poses against the whole-row accessors (pr_set_bodies / pr_get_bodies)."""
    sc = scenes_np.disks_box(32, seed=5)
    a, b = W.World(sc, lib=lib, device=device), W.World(sc, lib=lib, device=device)
    conf = refworld.default_step_conf(do_toi=0)
    for _ in range(10):
        a.step(conf); b.step(conf)
    rows = a.bodies()
    ids = np.arange(len(rows), dtype=np.uint32)[1::2]
    acc = np.stack([np.linspace(-3, 3, len(ids)), np.full(len(ids), -12.5), np.linspace(0.5, -0.5, len(ids))], axis=1).astype(np.float32)
    new = rows[ids].copy()
    new["acc_x"], new["acc_y"], new["acc_w"] = acc[:, 0], acc[:, 1], acc[:, 2]
    a.set_bodies(ids, new)          # SetBody with a Body whose only change is the acceleration
    b.set_accelerations(ids, acc)   # the column write
    for _ in range(20):
        sa, sb = a.step(conf), b.step(conf)
        assert np.array([sa]).tobytes() == np.array([sb]).tobytes()
    ra, rb = a.bodies(), b.bodies()
    assert ra.tobytes() == rb.tobytes()
    xf, vel = np.zeros((len(rb), 4), np.float32), np.zeros((len(rb), 3), np.float32)
    b.poses_into(xf, vel)
    assert np.array_equal(xf.view(np.uint32), np.stack([rb["xf_px"], rb["xf_py"], rb["xf_qc"], rb["xf_qs"]], axis=1).view(np.uint32))
    assert np.array_equal(vel.view(np.uint32), np.stack([rb["vel_x"], rb["vel_y"], rb["vel_w"]], axis=1).view(np.uint32))
    xf2 = np.zeros((5, 4), np.float32)
    b.poses_into(xf2, first=3)
    assert np.array_equal(xf2, xf[3:8])
    with pytest.raises(W.PrError):
        b.set_accelerations(np.array([len(rb)], np.uint32), np.zeros((1, 3), np.float32))
    allacc = np.stack([rb["acc_x"], rb["acc_y"], rb["acc_w"]], axis=1).astype(np.float32) * 0.5
    b.set_accelerations(None, allacc)  # ids == NULL: bodies 0..n-1
    assert np.array_equal(np.stack([b.bodies()[f] for f in ("acc_x", "acc_y", "acc_w")], axis=1), allacc)


@need_emul
def test_emulated_column_accessors_equal_row_accessors():
    _column_accessors_check(emul_lib(), -1)


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder,steps", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120),
    ("pyramid6", lambda: scenes_np.pyramid(6), 150),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150),
    ("mixed400", lambda: scenes_np.mixed(400, seed=3, tile_segments=32, tile_width=16.0), 120),
])
def test_emulated_parallel_toi_mode_bit_exact(name, builder, steps):
    """PR_TOI_PARALLEL (the product default and the benchmarked mode): the oracle derives the rounds by the documented rule
    (oracle/restate/orc_toi_phase.inc, solve_toi_parallel); event schedule, every state array and every StepStats field
    must agree bit for bit."""
    run_vs_oracle.toi_events = 0
    gs, ncontacts, prof = run_vs_oracle(builder(), steps, lib=emul_lib(), conf=refworld.default_step_conf(), toi_mode=1)
    assert ncontacts > 0
    if name.startswith("mixed"):
        assert run_vs_oracle.toi_events > 0


@need_oracle
@need_emul
@need_ref
def test_emulated_parallel_toi_mode_shooter_bit_exact():
    """The TOI stress scene (bullets, kinematic bar, sensor) in PR_TOI_PARALLEL: many rounds hold several events."""
    import scenes
    run_vs_oracle.toi_events = 0; run_vs_oracle.toi_multi_rounds = 0
    run_vs_oracle(scenes.shooter().export_scene(), 200, lib=emul_lib(), conf=refworld.default_step_conf(), toi_mode=1)
    assert run_vs_oracle.toi_events > 100
    assert run_vs_oracle.toi_multi_rounds > 0
