"""The scene compute-sanitizer runs (tools/sanitize_gpu.sh): small, but it walks every kernel family of the step -- regular phase
with colouring and the overflow tail, TOI rounds (list kernels + the single-block late rounds), joints, the block solver of a
batch arena, ray casts / AABB queries / TestPoint, events, snapshot -- and checks the result against nothing (parity is the
tests' job): the sanitizer's report is the product."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from playrho_b200 import abi, scenes_np, world as W  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
conf = W.default_step_conf()
# 1. mixed polygons / disks on chain terrain, TOI on (parallel rounds), events on
gw = W.World(scenes_np.mixed(300, seed=3, tile_segments=32, tile_width=16.0), device=0)
gw.set_event_mask(abi.EVENT_BEGIN | abi.EVENT_END | abi.EVENT_POST_SOLVE, 1 << 15)
for _ in range(steps):
    gw.step(conf)
px = gw.proxies()
rays = np.array([[px["min_x"].min(), px["max_y"].max(), px["max_x"].max(), px["min_y"].min(), 1.0]] * 8, np.float32)
gw.ray_cast(rays); gw.ray_cast(rays, mode=abi.RAY_CLOSEST)
gw.query_aabb(np.array([[0, 0, 10, 10]], np.float32))
gw.find_closest_bodies(np.zeros((4, 2), np.float32))
gw.events(abi.EVENT_BEGIN | abi.EVENT_END); gw.events(abi.EVENT_POST_SOLVE)
twin = W.World.restore(gw.snapshot(), device=0)
twin.step(conf)
# 2. serial-exact TOI rounds and the TMA-staged narrowphase variant
gw2 = W.World(scenes_np.pyramid(8), device=0, debug_flags=64)
gw2.set_toi_mode(0)
for _ in range(steps):
    gw2.step(conf)
# 3. a batch arena of tumblers: revolute joints, the block solver (one CTA per world)
units = [scenes_np.tumbler(40, seed=1 + w) for w in range(6)]
plain, world = scenes_np.merge_worlds(units)
b = W.Batch(6, device=0)
b.create_shapes(plain); b.create_bodies(plain["bodies"], world, plain["shape_offsets"], plain["shape_ids"]); b.create_joints(plain["joints"])
tconf = scenes_np.tumbler_step_conf(W.default_step_conf())
for _ in range(steps):
    b.step(tconf)
# 4. every joint type (the scenes of tests/test_more_joints.py, built through the reference harness): level-by-level joint launches,
#    the gear joints' extension table, ShiftOrigin over joints
sys.path.insert(0, os.path.join(ROOT, "tests"))
try:
    import test_more_joints as tmj  # noqa: E402
    for name in sorted(tmj.SCENES):
        ref, _ = tmj.SCENES[name][0]()
        jw = W.World(ref.export_scene(), device=0)
        for s in range(steps):
            if s == steps // 2:
                jw.shift_origin(1.5, -0.75)
            jw.step(conf)
        jw.gear_joints()
except Exception as e:  # (the reference harness is not built: the joint scenes are skipped)
    print("joint scenes skipped:", e)
print("sanitize scene done: %d + %d + %d bodies, %d steps each" % (len(gw.bodies()), len(gw2.bodies()), len(b.bodies()), steps))
