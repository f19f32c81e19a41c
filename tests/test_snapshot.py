"""Snapshot / restore (the reference's world is a value, AabbTreeWorld.cpp:842-880): a world saved with
pr_snapshot_save in mid-run -- through a FILE -- and loaded with pr_snapshot_load continues bit for bit like the world it was
taken from: bodies, contacts (ids, list order), manifolds, proxies, joints, every StepStats field, TOI on, and (the parity
link) both still equal the oracle fed the device's constraint order."""
import os
import numpy as np
import pytest
import refworld
import scenes_np
from common import need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import world as W


def _roundtrip(lib, device, scene, conf, before, after, tmp_path, with_oracle=True):
    gw = W.World(scene, lib=lib, device=device, record_solve_order=True)
    orc = refworld.OracleWorld(scene) if with_oracle else None
    if orc:
        orc.set_toi_mode(1)
    for _ in range(before):
        gw.step(conf)
        if orc:
            orc.step(conf, order=gw.solve_order()[0])
    path = os.path.join(str(tmp_path), "world.snap")
    gw.snapshot().tofile(path)
    twin = W.World.restore(np.fromfile(path, np.uint8), lib=lib, device=device)
    assert_bits(gw.bodies(), twin.bodies(), "restored bodies")
    assert_bits(gw.contacts()[0], twin.contacts()[0], "restored contacts")
    for s in range(after):
        a, b = gw.step(conf), twin.step(conf)
        where = "step %d after the restore" % s
        assert_bits(np.array([a]), np.array([b]), where + " stats")
        assert_bits(gw.bodies(), twin.bodies(), where + " bodies")
        ca, ma = gw.contacts(); cb, mb = twin.contacts()
        assert_bits(ca, cb, where + " contacts"); assert_bits(ma, mb, where + " manifolds")
        assert_bits(gw.proxies(), twin.proxies(), where + " proxies")
        if len(scene.get("joints", [])):
            assert_bits(gw.joints(), twin.joints(), where + " joints")
        if orc:
            os_ = orc.step(conf, order=twin.solve_order()[0])
            assert_bits(np.array([b]), np.array([os_]), where + " stats vs the oracle")
            assert_bits(twin.bodies(), orc.bodies(), where + " bodies vs the oracle")
    return gw, twin


def _joint_scene(name):
    """a scene of tests/test_more_joints.py (built through the reference harness): gear joints carry an extension entry per joint, chains the
    level lists of joints that share a body -- both must come back from a snapshot"""
    if not refworld.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    import test_more_joints as tmj
    ref, _ = tmj.SCENES[name][0]()
    return ref.export_scene()


SCENES = [
    ("mixed400", lambda: scenes_np.mixed(400, seed=3, tile_segments=32, tile_width=16.0), refworld.default_step_conf, 60, 40),
    ("pyramid10", lambda: scenes_np.pyramid(10), refworld.default_step_conf, 50, 40),
    ("tumbler80", lambda: scenes_np.tumbler(80), lambda: scenes_np.tumbler_step_conf(refworld.default_step_conf()), 60, 40),
    ("gear", lambda: _joint_scene("gear"), refworld.default_step_conf, 40, 40),
    ("chains", lambda: _joint_scene("chains"), refworld.default_step_conf, 30, 30),
    ("wheel", lambda: _joint_scene("wheel"), refworld.default_step_conf, 50, 40),
]


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder,conf,before,after", SCENES)
def test_emulated_snapshot_restore_is_a_bit_exact_continuation(name, builder, conf, before, after, tmp_path):
    _roundtrip(emul_lib(), -1, builder(), conf(), before, after, tmp_path)


@pytest.mark.gpu
@need_oracle
@pytest.mark.parametrize("name,builder,conf,before,after", SCENES)
def test_cuda_snapshot_restore_is_a_bit_exact_continuation(name, builder, conf, before, after, tmp_path):
    _roundtrip(None, 0, builder(), conf(), before, after, tmp_path)


@pytest.mark.gpu
def test_cuda_snapshot_restore_100k(tmp_path):
    """100k disks, saved while the pile settles (contact pool with freed ids, persistent colours, queued proxies)."""
    sc = scenes_np.disks_box(100000, seed=12345)
    _roundtrip(None, 0, sc, W.default_step_conf(), 45, 6, tmp_path, with_oracle=False)


@need_emul
def test_snapshot_rejects_garbage():
    lib = emul_lib()
    with pytest.raises(W.PrError):
        W.World.restore(np.zeros(4096, np.uint8), lib=lib)
    gw = W.World(scenes_np.pyramid(3), lib=lib)
    snap = gw.snapshot()
    with pytest.raises(W.PrError):
        W.World.restore(snap[: len(snap) // 2], lib=lib)
