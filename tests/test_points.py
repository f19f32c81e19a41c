"""TestPoint(world, body, shape, point) (WorldShape.cpp:178, DistanceProxy.cpp:107-158) and FindClosestBody(world, location)
(WorldBody.cpp:576-592) on the device (pr_test_points / pr_find_closest_bodies) against the REAL
reference's functions: the device world is stepped, the reference world is given the device's body transformations
(Body::SetSweep with the device's position 1: the transformation follows from it like in the device's write-back), and both answer the same queries -- points on a grid over
every body's surroundings incl. the vertex-radius skin, and locations with exact ties."""
import numpy as np
import pytest
import refworld
import scenes
from common import need_ref, need_emul, emul_lib, assert_bits
from playrho_b200 import world as W

u32, vp = refworld.u32, refworld.vp


def _ref_test_points(ref, body, shape, pts):
    ref.lib.ref_test_points.argtypes = [vp, u32, vp, vp, vp, vp]
    out = np.zeros(len(body), np.uint32)
    assert ref.lib.ref_test_points(ref.h, len(body), refworld.abi.ptr(body), refworld.abi.ptr(shape), refworld.abi.ptr(pts), refworld.abi.ptr(out)) == 0
    return out


def _ref_closest(ref, pts):
    ref.lib.ref_find_closest_bodies.argtypes = [vp, u32, vp, vp]
    out = np.zeros(len(pts), np.uint32)
    assert ref.lib.ref_find_closest_bodies(ref.h, len(pts), refworld.abi.ptr(pts), refworld.abi.ptr(out)) == 0
    return out


def _check(lib, device, name, steps):
    ref = scenes.SCENES[name]()
    sc = ref.export_scene()
    gw = W.World(sc, lib=lib, device=device)
    conf = W.default_step_conf(lib)
    for _ in range(steps):
        gw.step(conf)
    rows = gw.bodies()
    ref.lib.ref_set_body_sweep.argtypes = [vp, u32, refworld.f32, refworld.f32, refworld.f32]
    for b in range(len(rows)):  # the reference world takes the device's poses: both answer about the same world
        assert ref.lib.ref_set_body_sweep(ref.h, b, refworld.f32(rows["pos1_x"][b]), refworld.f32(rows["pos1_y"][b]), refworld.f32(rows["pos1_a"][b])) == 0
    rb = ref.bodies()
    for f in ("xf_px", "xf_py", "xf_qc", "xf_qs"):
        assert_bits(rb[f], rows[f], "transformations given to the reference: " + f)
    rng = np.random.default_rng(11)
    # TestPoint: for every (body, shape) a cloud of points around the body incl. points on its skin
    offs, ids = sc["shape_offsets"], sc["shape_ids"]
    body, shape, pts = [], [], []
    for b in range(len(rows)):
        for s in ids[offs[b]:offs[b + 1]]:
            k = 24
            body += [b] * k; shape += [int(s)] * k
            c = np.array([rows["xf_px"][b], rows["xf_py"][b]], np.float64)
            pts.append(c + rng.normal(0.0, 0.45, (k, 2)))
    body = np.array(body, np.uint32); shape = np.array(shape, np.uint32)
    pts = np.concatenate(pts).astype(np.float32)
    got, want = gw.test_points(body, shape, pts), _ref_test_points(ref, body, shape, pts)
    assert want.sum() > 50 and (want == 0).sum() > 50
    assert np.array_equal(got, want), "TestPoint differs at %s" % np.flatnonzero(got != want)[:8]
    # FindClosestBody: random locations, every body's own location (distance 0), midpoints of body pairs (near-ties)
    locs = [rng.uniform([rows["xf_px"].min() - 2, rows["xf_py"].min() - 2], [rows["xf_px"].max() + 2, rows["xf_py"].max() + 2], (400, 2)),
            np.stack([rows["xf_px"], rows["xf_py"]], 1)]
    i, j = rng.integers(0, len(rows), 200), rng.integers(0, len(rows), 200)
    locs.append(0.5 * (np.stack([rows["xf_px"][i], rows["xf_py"][i]], 1).astype(np.float64) + np.stack([rows["xf_px"][j], rows["xf_py"][j]], 1)))
    locs = np.concatenate(locs).astype(np.float32)
    got, want = gw.find_closest_bodies(locs), _ref_closest(ref, locs)
    assert np.array_equal(got, want), "FindClosestBody differs at %s" % np.flatnonzero(got != want)[:8]


@need_ref
@need_emul
@pytest.mark.parametrize("name,steps", [("pyramid10", 40), ("mixed100", 100), ("shooter", 40)])
def test_emulated_test_point_and_find_closest_body_equal_reference(name, steps):
    _check(emul_lib(), -1, name, steps)


@pytest.mark.gpu
@need_ref
@pytest.mark.parametrize("name,steps", [("pyramid10", 40), ("mixed100", 100), ("shooter", 40)])
def test_cuda_test_point_and_find_closest_body_equal_reference(name, steps):
    _check(None, 0, name, steps)


@need_emul
def test_find_closest_body_in_an_empty_world():
    gw = W.World(None, lib=emul_lib())
    assert gw.find_closest_bodies(np.zeros((2, 2), np.float32)).tolist() == [0xFFFFFFFF, 0xFFFFFFFF]
