"""Spatial query Query(tree, AABB, callback) (DynamicTree.cpp:673-710) on the device LBVH.

Two links of parity: (1) the REFERENCE's own Query over its DynamicTree reports exactly the leaves whose fat AABB overlaps
the box (closed intervals) -- checked against brute force over the reference's exported leaves; (2) pr_query_aabb reports
exactly that set over the device's proxies, whose fat AABBs the step parity tests already pin bit for bit to the reference's."""
import numpy as np
import pytest

import refworld
import scenes_np
from common import need_ref, need_emul, emul_lib
from playrho_b200 import world as W


def _brute(px, box):
    hit = (np.minimum(px["max_x"], box[2]) >= np.maximum(px["min_x"], box[0])) & (np.minimum(px["max_y"], box[3]) >= np.maximum(px["min_y"], box[1]))
    return sorted(zip(px["body"][hit].tolist(), px["shape"][hit].tolist(), px["child"][hit].tolist()))


def _boxes(px, rng, n):
    lo_x, hi_x = float(px["min_x"].min()), float(px["max_x"].max())
    lo_y, hi_y = float(px["min_y"].min()), float(px["max_y"].max())
    c = np.stack([rng.uniform(lo_x, hi_x, n), rng.uniform(lo_y, hi_y, n)], axis=1)
    h = rng.uniform(0.01, 0.15, (n, 2)) * np.array([hi_x - lo_x, hi_y - lo_y])
    b = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    # edge cases: a box touching a leaf exactly at its max corner (closed intervals overlap), a box far away, one covering all
    k = int(np.argmax(px["max_x"]))
    b[0] = [px["max_x"][k], px["max_y"][k], px["max_x"][k] + 1.0, px["max_y"][k] + 1.0]
    b[1] = [hi_x + 10.0, hi_y + 10.0, hi_x + 11.0, hi_y + 11.0]
    b[2] = [lo_x - 1.0, lo_y - 1.0, hi_x + 1.0, hi_y + 1.0]
    return b


@need_ref
def test_reference_query_is_closed_interval_overlap_of_leaf_boxes():
    import scenes
    ref = scenes.mixed(120, seed=3)
    conf = refworld.default_step_conf()
    for _ in range(40):
        ref.step(conf)
    px = ref.proxies()
    rng = np.random.default_rng(1)
    for box in _boxes(px, rng, 40):
        b, s, c = ref.query_aabb(box)
        assert sorted(zip(b.tolist(), s.tolist(), c.tolist())) == _brute(px, box)


def _device_query_check(lib, device, sc, steps, n_boxes):
    gw = W.World(sc, lib=lib, device=device)
    conf = W.default_step_conf(lib=lib)
    for _ in range(steps):
        gw.step(conf)
    px = gw.proxies()
    rng = np.random.default_rng(2)
    boxes = _boxes(px, rng, n_boxes)
    qi, b, s, c = gw.query_aabb(boxes)
    assert len(qi) > n_boxes
    for i, box in enumerate(boxes):
        m = qi == i
        assert sorted(zip(b[m].tolist(), s[m].tolist(), c[m].tolist())) == _brute(px, box), "box %d" % i
    qi2, *_ = gw.query_aabb(boxes, cap=8)  # too-short arrays: retried by the binding with the reported size
    assert len(qi2) == len(qi)
    gw.step(conf)  # the step goes on with the same tree afterwards
    assert len(gw.query_aabb(boxes[2:3])[0]) == len(gw.proxies())


@need_emul
def test_emulated_query_aabb_equals_brute_force():
    _device_query_check(emul_lib(), -1, scenes_np.mixed(150, seed=11, tile_segments=16, tile_width=8.0), 60, 32)


@pytest.mark.gpu
def test_cuda_query_aabb_equals_brute_force():
    _device_query_check(None, 0, scenes_np.mixed(20000, seed=5, tile_segments=64, tile_width=32.0), 80, 300)
