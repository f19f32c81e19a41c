"""RayCast(world, input, callback) (RayCastOutput.cpp:241-337 over :41-75 / :139-233) on the device LBVH.

Links of parity, all bit for bit (fraction, normal, point):
  1. the REAL reference's RayCast(world, input, cb -> ResetRay) == the restatement's brute force over the proxies (the tree's
     two leaf tests + the shape cast), as sets of (ray, body, shape, child, fraction, normal, point);
  2. the device kernels compiled for the host (tests/_emul) == the restatement on the same stepped world (PR_RAY_ALL), and
     PR_RAY_CLOSEST == the least-fraction hit per ray;
  3. the same on the GPU (-m gpu)."""
import numpy as np
import pytest
import refworld
import scenes
import scenes_np
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W


def _rays(px, rng, n):
    lo = np.array([float(px["min_x"].min()), float(px["min_y"].min())]) - 1.0
    hi = np.array([float(px["max_x"].max()), float(px["max_y"].max())]) + 1.0
    p1 = rng.uniform(lo, hi, (n, 2))
    p2 = rng.uniform(lo, hi, (n, 2))
    mf = rng.uniform(0.2, 1.0, (n, 1))
    rays = np.concatenate([p1, p2, mf], axis=1).astype(np.float32)
    # edge cases: straight down through the pile, horizontal along the ground, a zero-length ray, a ray far away, full fraction
    cx = 0.5 * (lo[0] + hi[0])
    rays[0] = [cx, hi[1], cx, lo[1], 1.0]
    rays[1] = [lo[0], lo[1] + 1.5, hi[0], lo[1] + 1.5, 1.0]
    rays[2] = [cx, hi[1], cx, hi[1], 1.0]
    rays[3] = [hi[0] + 50.0, hi[1] + 50.0, hi[0] + 60.0, hi[1] + 55.0, 1.0]
    rays[4] = [lo[0], hi[1], hi[0], lo[1], 0.5]
    return rays


def _sorted_hits(h):
    return h[np.lexsort((h["child"], h["shape"], h["body"], h["ray"]))]


@need_ref
@need_oracle
@pytest.mark.parametrize("name,steps", [("pyramid10", 60), ("disks100", 80), ("mixed100", 120), ("shooter", 60)])
def test_restatement_ray_cast_equals_reference(name, steps):
    ref = scenes.SCENES[name]()
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    for _ in range(steps):
        ref.step(conf); orc.step(conf)
    assert_bits(ref.bodies(), orc.bodies(), "bodies")
    rays = _rays(ref.proxies(), np.random.default_rng(5), 300)
    rh, oh = _sorted_hits(ref.ray_cast(rays)), _sorted_hits(orc.ray_cast(rays))
    assert len(rh) > 100
    assert_bits(rh, oh, "%s ray hits" % name)
    assert (rh["fraction"] >= 0).all() and (rh["fraction"] <= rays[rh["ray"], 4]).all()
    assert len(rh[rh["ray"] == 2]) == 0 and len(rh[rh["ray"] == 3]) == 0  # zero-length ray, ray far away


def _closest(h, n_rays):
    out = {}
    for r in h:
        k = int(r["ray"])
        if k not in out or r["fraction"] < out[k]["fraction"]:
            out[k] = r
    return out


def _device_ray_check(sc, steps, lib=None, device=-1, n_rays=300):
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    orc = refworld.OracleWorld(sc)
    for _ in range(steps):
        gw.step(conf)
        orc.step(conf, order=gw.solve_order()[0])
    assert_bits(gw.bodies(), orc.bodies(), "bodies")
    assert_bits(gw.proxies(), orc.proxies(), "proxies")
    rays = _rays(gw.proxies(), np.random.default_rng(5), n_rays)
    gh, oh = _sorted_hits(gw.ray_cast(rays)), _sorted_hits(orc.ray_cast(rays))
    assert len(oh) > 100
    assert_bits(gh, oh, "all hits")
    # too short an output array: the count still says how many there were
    import ctypes as C
    few = np.zeros(8, abi.RAY_HIT); k = C.c_uint32(0)
    assert gw.lib.pr_ray_cast(gw.h, len(rays), abi.ptr(rays), None, abi.RAY_ALL, 8, abi.ptr(few), C.byref(k)) == 0 and k.value == len(oh)
    # closest hit per ray
    gc = gw.ray_cast(rays, mode=abi.RAY_CLOSEST)
    assert len(gc) == len(rays) and (gc["ray"] == np.arange(len(rays))).all()
    want = _closest(oh, len(rays))
    unique = 0
    for i in range(len(rays)):
        if i not in want:
            assert gc[i]["body"] == 0xFFFFFFFF, "ray %d should miss" % i
            continue
        w = want[i]
        assert gc[i]["fraction"].tobytes() == w["fraction"].tobytes(), "ray %d fraction" % i
        ties = oh[(oh["ray"] == i) & (oh["fraction"] == w["fraction"])]
        if len(ties) == 1:
            unique += 1
            assert gc[i].tobytes() == w.tobytes(), "ray %d closest hit" % i
    assert unique > 50
    return gw


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder,steps", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 80),
    ("pyramid6", lambda: scenes_np.pyramid(6), 60),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 120),
])
def test_emulated_ray_cast_equals_oracle(name, builder, steps):
    _device_ray_check(builder(), steps, lib=emul_lib())


@need_emul
def test_ray_cast_argument_errors():
    gw = W.World(scenes_np.disks_box(16, seed=7), lib=emul_lib())
    import ctypes as C
    rays = np.array([[0, 0, 1, 1, 1.5]], np.float32)
    out = np.zeros(4, abi.RAY_HIT); k = C.c_uint32(0)
    assert gw.lib.pr_ray_cast(gw.h, 1, abi.ptr(rays), None, abi.RAY_ALL, 4, abi.ptr(out), C.byref(k)) == -1      # maxFraction > 1
    rays[0, 4] = 1.0
    assert gw.lib.pr_ray_cast(gw.h, 1, abi.ptr(rays), None, 7, 4, abi.ptr(out), C.byref(k)) == -1                 # unknown mode
    assert gw.lib.pr_ray_cast(gw.h, 1, abi.ptr(rays), None, abi.RAY_CLOSEST, 0, abi.ptr(out), C.byref(k)) == -1   # cap < n_rays
    assert gw.lib.pr_ray_cast(gw.h, 0, abi.ptr(rays), None, abi.RAY_ALL, 4, abi.ptr(out), C.byref(k)) == 0 and k.value == 0


@pytest.mark.gpu
@need_oracle
@pytest.mark.parametrize("name,builder,steps", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 80),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 120),
])
def test_gpu_ray_cast_equals_oracle(name, builder, steps):
    _device_ray_check(builder(), steps, device=0)


@pytest.mark.gpu
def test_gpu_ray_cast_large_world_closest_is_least_of_all():
    """20k mixed bodies, 4096 rays: PR_RAY_CLOSEST equals the least fraction of PR_RAY_ALL per ray (size-independent property)."""
    sc = scenes_np.mixed_fast(20000, seed=3)
    gw = W.World(sc, device=0)
    conf = W.default_step_conf()
    for _ in range(60):
        gw.step(conf)
    rays = _rays(gw.proxies(), np.random.default_rng(9), 4096)
    allh = gw.ray_cast(rays, cap=1 << 20)
    gc = gw.ray_cast(rays, mode=abi.RAY_CLOSEST)
    assert len(allh) > 4096
    best = np.full(len(rays), np.inf, np.float32)
    np.minimum.at(best, allh["ray"], allh["fraction"])
    hit = gc["body"] != 0xFFFFFFFF
    assert (hit == np.isfinite(best)).all()
    assert (gc["fraction"][hit] == best[hit]).all()
