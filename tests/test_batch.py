"""Batched independent worlds (BASELINE config 5): n worlds in one device arena.

Oracle: worlds share nothing, so an arena of K worlds must equal -- bit for bit -- ONE reference-semantics world that
holds the same bodies with world w's shapes on their own collision-Filter layer (category = mask = 1 << w): in both,
pairs exist only inside a world. The arena gets NO filters (its broadphase keeps worlds apart by world index), the
oracle gets the filter layers and the device's colour-sorted constraint order."""
import numpy as np
import pytest

import refworld
import scenes_np
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi, world as W


def _arena_vs_oracle(lib, device, n_worlds, n_squares, steps, every, toi_mode=0, units=None, conf=None, min_contacts=None):
    if units is None:
        units = [scenes_np.tumbler(n_squares, seed=12345 + w) for w in range(n_worlds)]
    plain, world = scenes_np.merge_worlds(units)
    layered, _ = scenes_np.merge_worlds(units, layer_filters=True)
    if conf is None:
        conf = scenes_np.tumbler_step_conf(refworld.default_step_conf())
    gw = W.Batch(n_worlds, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(toi_mode)
    gw.create_shapes(plain)
    gw.create_bodies(plain["bodies"], world, plain["shape_offsets"], plain["shape_ids"])
    gw.create_joints(plain["joints"], plain.get("gears"))
    assert np.array_equal(gw.body_worlds(), world)
    orc = refworld.OracleWorld(layered)
    n_contacts = 0
    for s in range(steps):
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        if (s % every) and s != steps - 1:
            continue
        where = "step %d" % s
        assert_bits(gw.bodies(), orc.bodies(), where + " bodies")
        assert_bits(gw.joints(), orc.joints(), where + " joints")
        gc, gm = gw.contacts()
        oc, om = orc.contacts()
        assert_bits(gc, oc, where + " contacts")
        assert_bits(gm, om, where + " manifolds")
        assert_bits(gw.proxies(), orc.proxies(), where + " proxies")
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, where + " stats")
        # no contact ever joins two worlds
        assert np.array_equal(world[gc["body_a"]], world[gc["body_b"]])
        n_contacts = max(n_contacts, len(gc))
    assert n_contacts > (n_worlds * n_squares // 2 if min_contacts is None else min_contacts)
    return gw


def _joint_units():
    """one world per joint type: the scenes of test_more_joints.py"""
    import test_more_joints as tmj
    units = []
    for name in ("weld", "prismatic", "wheel", "pulley", "target", "chains", "gear"):
        ref, _ = tmj.SCENES[name][0]()
        units.append(ref.export_scene())
    return units


@need_ref
@need_oracle
@need_emul
def test_emulated_batch_arena_of_jointed_worlds_equals_filter_layered_oracle():
    """every joint type in worlds of one arena: the arena's joint kernels (and, on the GPU, the block solver's) against the
    restatement with the worlds on filter layers"""
    units = _joint_units()
    _arena_vs_oracle(emul_lib(), -1, n_worlds=len(units), n_squares=0, steps=120, every=10, units=units, conf=refworld.default_step_conf(), min_contacts=4)


@pytest.mark.gpu
@need_ref
@need_oracle
def test_cuda_batch_arena_of_jointed_worlds_equals_filter_layered_oracle():
    units = _joint_units()
    _arena_vs_oracle(None, 0, n_worlds=len(units), n_squares=0, steps=160, every=10, units=units, conf=refworld.default_step_conf(), min_contacts=4)


@need_oracle
@need_emul
def test_emulated_batch_arena_equals_filter_layered_oracle():
    _arena_vs_oracle(emul_lib(), -1, n_worlds=5, n_squares=30, steps=120, every=10)


@need_oracle
@need_emul
def test_emulated_batch_rejects_misuse():
    b = W.Batch(3, lib=emul_lib())
    sc = scenes_np.tumbler(2)
    b.create_shapes(sc)
    first = W.u32(0)
    rc = b.lib.pr_create_bodies(b.h, len(sc["bodies"]), abi.ptr(sc["bodies"]), abi.ptr(sc["shape_offsets"]),
                                abi.ptr(sc["shape_ids"]), W.C.byref(first))
    assert rc == -1  # PR_ERR_INVALID_ARGUMENT: arena bodies need a world index
    with pytest.raises(W.PrError):
        b.create_bodies(sc["bodies"], np.full(len(sc["bodies"]), 3, np.uint32), sc["shape_offsets"], sc["shape_ids"])
    b.create_bodies(sc["bodies"], np.array([0, 0, 1, 2], np.uint32), sc["shape_offsets"], sc["shape_ids"])
    j = sc["joints"].copy()
    j["body_a"], j["body_b"] = 1, 2  # bodies of worlds 0 and 1
    with pytest.raises(W.PrError):
        b.create_joints(j)


@pytest.mark.gpu
@need_oracle
def test_cuda_batch_arena_equals_filter_layered_oracle():
    _arena_vs_oracle(None, 0, n_worlds=12, n_squares=60, steps=200, every=20)


def _big_arena(n_worlds, n_sq, steps, add_steps):
    units = [scenes_np.tumbler(n_sq, seed=99 + w) for w in range(n_worlds)]
    plain, world = scenes_np.merge_worlds(units)
    conf = scenes_np.tumbler_step_conf(W.default_step_conf())
    gw = W.Batch(n_worlds, device=0)
    gw.create_shapes(plain)
    gw.create_bodies(plain["bodies"], world, plain["shape_offsets"], plain["shape_ids"])
    gw.create_joints(plain["joints"])
    row = scenes_np.tumbler_square_row(units[0])
    sq_ids = (units[0]["square_shape"] + len(units[0]["shapes"]) * np.arange(n_worlds)).astype(np.uint32)
    islands = 0
    for s in range(steps):
        if s < add_steps:  # AddSquare() in every world, one per step (BenchmarkMain.cpp:2716)
            gw.create_bodies(np.repeat(row, n_worlds), np.arange(n_worlds, dtype=np.uint32),
                             np.arange(n_worlds + 1, dtype=np.uint32), sq_ids)
        islands = int(gw.step(conf)["reg"]["islands_found"])
    return gw, world, islands


@pytest.mark.gpu
def test_cuda_batch_arena_at_scale_invariants():
    """Size-independent properties at 512 worlds x (2 + 48 + 20) bodies: no contact joins two worlds, no square
    leaves its enclosure (TOI on: no tunnelling, cf. UnitTests/World.cpp:3132), every world still has its bodies, and
    a second arena built the same way ends bit-identical (the arena step is deterministic)."""
    n_worlds, n_sq, steps, add_steps = 512, 48, 90, 20
    gw, world, islands = _big_arena(n_worlds, n_sq, steps, add_steps)
    rows = gw.bodies()
    bw = gw.body_worlds()
    assert len(rows) == n_worlds * (2 + n_sq + add_steps)
    assert np.array_equal(np.bincount(bw, minlength=n_worlds), np.full(n_worlds, 2 + n_sq + add_steps))
    cs, _ = gw.contacts()
    assert np.array_equal(bw[cs["body_a"]], bw[cs["body_b"]])
    assert len(cs) > n_worlds * n_sq
    dyn = rows["inv_mass"] > 0
    assert np.all(np.abs(rows["pos1_x"][dyn]) < 11.0) and np.all(np.abs(rows["pos1_y"][dyn] - 10.0) < 11.0)
    assert np.all(np.isfinite(rows["vel_x"])) and np.all(np.isfinite(rows["pos1_a"]))
    assert islands >= n_worlds  # at least the enclosure's island in every world
    gw2, _, _ = _big_arena(n_worlds, n_sq, steps, add_steps)
    assert gw2.bodies().tobytes() == rows.tobytes()


def _arena_queries_vs_oracle(lib, device, n_worlds=4, n_squares=25, steps=90):
    """Contact events and ray casts on an arena: events of the arena == events of the filter-layered oracle world; a ray cast
    in world w reports exactly the oracle's hits on bodies of world w (shapes of other worlds overlap it in space)."""
    units = [scenes_np.tumbler(n_squares, seed=12345 + w) for w in range(n_worlds)]
    plain, world = scenes_np.merge_worlds(units)
    layered, _ = scenes_np.merge_worlds(units, layer_filters=True)
    conf = scenes_np.tumbler_step_conf(refworld.default_step_conf())
    gw = W.Batch(n_worlds, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    gw.create_shapes(plain)
    gw.create_bodies(plain["bodies"], world, plain["shape_offsets"], plain["shape_ids"])
    gw.create_joints(plain["joints"])
    orc = refworld.OracleWorld(layered)
    mask = abi.EVENT_BEGIN | abi.EVENT_END | abi.EVENT_POST_SOLVE
    gw.set_event_mask(mask, 1 << 16); orc.set_event_mask(mask)
    n_begin = n_post = 0
    for s in range(steps):
        gw.step(conf)
        orc.step(conf, order=gw.solve_order()[0])
        ge = gw.events(abi.EVENT_BEGIN | abi.EVENT_END)
        oe = orc.events(abi.EVENT_BEGIN | abi.EVENT_END)
        oe = oe[np.lexsort((oe["contact"], oe["kind"]))]
        assert_bits(ge, oe, "step %d begin/end events" % s)
        gi, oi = gw.events(abi.EVENT_POST_SOLVE), orc.events(abi.EVENT_POST_SOLVE)
        key = lambda e: e[np.lexsort((np.ascontiguousarray(e["normal"]).view(np.uint32)[:, 0], e["iterations"], e["contact"], e["phase"]))]
        assert_bits(key(gi), key(oi), "step %d post-solve events" % s)
        n_begin += int((oe["kind"] == abi.EVENT_BEGIN).sum()); n_post += len(oi)
    assert n_begin > n_worlds and n_post > 100
    assert_bits(gw.bodies(), orc.bodies(), "bodies")
    px = gw.proxies()
    rng = np.random.default_rng(3)
    lo = np.array([px["min_x"].min(), px["min_y"].min()]) - 0.5
    hi = np.array([px["max_x"].max(), px["max_y"].max()]) + 0.5
    n = 40 * n_worlds
    rays = np.concatenate([rng.uniform(lo, hi, (n, 2)), rng.uniform(lo, hi, (n, 2)), np.ones((n, 1))], axis=1).astype(np.float32)
    rw = (np.arange(n) % n_worlds).astype(np.uint32)
    gh = gw.ray_cast(rays, ray_world=rw)
    oh = orc.ray_cast(rays)
    oh = oh[world[oh["body"]] == rw[oh["ray"]]]          # the oracle world holds every world's bodies: keep the ray's world
    srt = lambda h: h[np.lexsort((h["child"], h["shape"], h["body"], h["ray"]))]
    assert len(oh) > n_worlds
    assert_bits(srt(gh), srt(oh), "arena ray hits")
    assert (world[gh["body"]] == rw[gh["ray"]]).all()
    import ctypes as C
    out = np.zeros(4, abi.RAY_HIT); k = C.c_uint32(0)
    assert gw.lib.pr_ray_cast(gw.h, 1, abi.ptr(rays), None, abi.RAY_ALL, 4, abi.ptr(out), C.byref(k)) == -1  # arena: world index needed


@need_oracle
@need_emul
def test_emulated_batch_arena_events_and_ray_casts():
    _arena_queries_vs_oracle(emul_lib(), -1)


@pytest.mark.gpu
@need_oracle
def test_cuda_batch_arena_events_and_ray_casts():
    _arena_queries_vs_oracle(None, 0, n_worlds=8, n_squares=40, steps=120)
