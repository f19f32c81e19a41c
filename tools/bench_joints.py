"""Joint-heavy world on the GPU: C three-link pendulum chains of revolute joints hanging from ONE static body (3C joints, 3C + 1 bodies;
the links' filter keeps them from colliding, so the step is islands + joints). Times pr_step with the joints solved level by level
(default) and, for the A/B, by one thread (pr_set_debug bit 11).
usage: python tools/bench_joints.py [chains] [steps]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from playrho_b200 import abi, world as W
from playrho_b200.scenes_np import SceneBuilder


def scene(chains):
    b = SceneBuilder()
    ground = b.body(0, pos=(0.0, 0.0))
    link = b.box(0.3, 0.05, density=5.0)
    side = int(np.ceil(np.sqrt(chains)))
    rows = []
    for c in range(chains):
        x0, y0 = 2.5 * (c % side), 2.0 * (c // side)
        prev, ppos = ground, (0.0, 0.0)
        for k in range(3):
            pos = (x0 + 0.3 + 0.6 * k, y0)
            i = b.body(2, pos=pos, acc=(0.0, -10.0), shapes=[link])
            wx, wy = x0 + 0.6 * k, y0
            rows.append((prev, i, wx - ppos[0], wy - ppos[1], wx - pos[0], wy - pos[1]))
            prev, ppos = i, pos
    sc = b.finish()
    sc["shapes"]["filter_mask"] = 0      # the links collide with nothing
    j = np.zeros(len(rows), abi.REVOLUTE)
    r = np.array(rows, dtype=np.float64)
    j["body_a"], j["body_b"] = r[:, 0].astype(np.uint32), r[:, 1].astype(np.uint32)
    j["local_anchor_a"]["x"], j["local_anchor_a"]["y"] = r[:, 2], r[:, 3]
    j["local_anchor_b"]["x"], j["local_anchor_b"]["y"] = r[:, 4], r[:, 5]
    sc["joints"] = j
    return sc


def main():
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    sc = scene(chains)
    conf = W.default_step_conf()
    for flags, what in ((0, "level by level"), (2048, "one thread")):
        n = 2 if (flags and chains > 2000) else steps   # (one thread over 300 000 joints x 12 phases takes seconds per step)
        w = W.World(sc, device=0, debug_flags=flags)
        for _ in range(2):
            w.step(conf)
        t0 = time.perf_counter()
        for _ in range(n):
            w.step(conf)
        dt = (time.perf_counter() - t0) / n
        y = w.bodies()["xf_py"]
        print("%d chains (%d joints), joints %s: %.3f ms/step (min y %.3f)" % (chains, 3 * chains, what, dt * 1e3, float(y.min())), flush=True)


if __name__ == "__main__":
    main()
