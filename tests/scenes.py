"""Seeded scene recipes (TEST INFRASTRUCTURE) for the BASELINE.json configs, built through the
reference's own shape/mass code via RefWorld (oracle/ref_harness.cpp), then exported as PODs.
Scene shapes follow SURVEY.md §8(d)."""
import numpy as np
from refworld import RefWorld, STATIC, DYNAMIC, KINEMATIC

GRAVITY = (0.0, -9.8)


def hello_world(ids32=False):
    """HelloWorld/HelloWorld.cpp:31-93 : static 50x10 box at (0,-10), r=1 disk at (0,4)."""
    w = RefWorld(ids32)
    box = w.add_box(50.0, 10.0)
    w.add_body(STATIC, pos=(0.0, -10.0), shapes=[box])
    disk = w.add_disk(1.0)
    w.add_body(DYNAMIC, pos=(0.0, 4.0), acc=GRAVITY, shapes=[disk])
    return w


def pyramid(rows=10, ids32=False):
    """Testbed/Tests/Pyramid.cpp recipe: edge ground, `rows` rows of 1 m boxes (config 2 at rows=40)."""
    w = RefWorld(ids32)
    ground = w.add_edge((-60.0, 0.0), (60.0, 0.0))
    w.add_body(STATIC, shapes=[ground])
    box = w.add_box(0.5, 0.5, density=5.0)
    x = np.array([-7.0, 0.75], dtype=np.float32)
    dx = np.array([0.5625, 1.25], dtype=np.float32)
    dy = np.array([1.125, 0.0], dtype=np.float32)
    for i in range(rows):
        y = x.copy()
        for _ in range(i, rows):
            w.add_body(DYNAMIC, pos=(float(y[0]), float(y[1])), acc=GRAVITY, shapes=[box])
            y = (y + dy).astype(np.float32)
        x = (x + dx).astype(np.float32)
    return w


def disks_box(n=200, seed=12345, ids32=False, radius=0.5):
    """Config 3 recipe: 4 edges on one static body forming a W x 2W box; n disks on a jittered lattice."""
    rng = np.random.RandomState(seed)
    w = RefWorld(ids32)
    side = int(np.ceil(np.sqrt(n)))
    W = 1.1 * side
    edges = [w.add_edge((0.0, 0.0), (W, 0.0)), w.add_edge((W, 0.0), (W, 2 * W)),
             w.add_edge((W, 2 * W), (0.0, 2 * W)), w.add_edge((0.0, 2 * W), (0.0, 0.0))]
    w.add_body(STATIC, shapes=edges)
    disk = w.add_disk(radius, density=1.0)
    pitch = 1.05
    k = 0
    for j in range(side):
        for i in range(side):
            if k >= n:
                break
            jx, jy = rng.uniform(-0.02, 0.02, 2)
            w.add_body(DYNAMIC, pos=(0.6 + i * pitch + jx, 0.6 + j * pitch + jy), acc=GRAVITY, shapes=[disk])
            k += 1
    return w


def mixed(n=200, seed=12345, ids32=False, terrain_segments=32):
    """Config 4 recipe (scaled): sawtooth chain terrain + 50% disks / 50% convex k-gons (k in 4..8)."""
    rng = np.random.RandomState(seed)
    w = RefWorld(ids32)
    side = int(np.ceil(np.sqrt(n)))
    width = side * 1.1 + 2.0
    xs = np.linspace(-1.0, width, terrain_segments + 1)
    ys = np.where(np.arange(terrain_segments + 1) % 2 == 0, 0.0, 0.5)
    ys[0] = ys[-1] = 4.0
    chain = w.add_chain(np.stack([xs, ys], 1))
    w.add_body(STATIC, shapes=[chain])
    k = 0
    for j in range(side):
        for i in range(side):
            if k >= n:
                break
            r = rng.uniform(0.2, 0.5)
            pos = (0.5 + i * 1.1 + rng.uniform(-0.02, 0.02), 2.0 + j * 1.1 + rng.uniform(-0.02, 0.02))
            if rng.rand() < 0.5:
                s = w.add_disk(r, density=1.0)
            else:
                kk = rng.randint(4, 9)
                ang = np.sort(rng.uniform(0, 2 * np.pi, kk))
                s = w.add_polygon(np.stack([r * np.cos(ang), r * np.sin(ang)], 1), density=1.0)
            w.add_body(DYNAMIC, pos=pos, angle=rng.uniform(-3, 3), acc=GRAVITY, shapes=[s])
            k += 1
    return w


def tumbler(n_squares=40, seed=12345, ids32=False):
    """Config 5 unit (Benchmark/BenchmarkMain.cpp:2614-2722): 4-box enclosure on a motorised revolute."""
    rng = np.random.RandomState(seed)
    w = RefWorld(ids32)
    g = w.add_body(STATIC)
    walls = [w.add_box(0.5, 10.0, (10.0, 0.0), 0.0, density=5.0), w.add_box(0.5, 10.0, (-10.0, 0.0), 0.0, density=5.0),
             w.add_box(10.0, 0.5, (0.0, 10.0), 0.0, density=5.0), w.add_box(10.0, 0.5, (0.0, -10.0), 0.0, density=5.0)]
    b = w.add_body(DYNAMIC, pos=(0.0, 10.0), allow_sleep=False, acc=GRAVITY, shapes=walls)
    w.add_revolute(g, b, (0.0, 10.0), enable_motor=True, motor_speed=0.05 * np.pi, max_motor_torque=1e5)
    sq = w.add_box(0.125, 0.125, density=1.0)
    for _ in range(n_squares):
        w.add_body(DYNAMIC, pos=(rng.uniform(-4, 4), 10.0 + rng.uniform(-4, 4)), acc=GRAVITY, shapes=[sq])
    return w


def shooter(n=60, seed=4, ids32=False, sensor=True):
    """TOI stress: closed edge box, fast random bodies (some bullets), a spinning kinematic bar and a
    static sensor disk (exercises GetToiViaSat, SolveToi mini-islands, bullet-vs-dynamic TOI, TestOverlap)."""
    rng = np.random.RandomState(seed)
    w = RefWorld(ids32)
    W = 14.0
    edges = [w.add_edge((0.0, 0.0), (W, 0.0)), w.add_edge((W, 0.0), (W, W)), w.add_edge((W, W), (0.0, W)),
             w.add_edge((0.0, W), (0.0, 0.0))]
    w.add_body(STATIC, shapes=edges)
    bar = w.add_box(2.5, 0.2, density=1.0)
    w.add_body(KINEMATIC, pos=(W / 2, W / 2), w=2.5, shapes=[bar])
    if sensor:
        sd = w.add_disk(1.5, sensor=True)
        w.add_body(STATIC, pos=(3.0, 3.0), shapes=[sd])
    side = int(np.ceil(np.sqrt(n)))
    for k in range(n):
        i, j = k % side, k // side
        pos = (1.0 + i * (W - 2.0) / side + rng.uniform(-0.05, 0.05), 1.0 + j * (W - 2.0) / side + rng.uniform(-0.05, 0.05))
        r = rng.uniform(0.15, 0.4)
        if rng.rand() < 0.5:
            sh = w.add_disk(r, density=1.0, restitution=rng.uniform(0, 0.6))
        else:
            kk = rng.randint(3, 8)
            ang = np.sort(rng.uniform(0, 2 * np.pi, kk))
            sh = w.add_polygon(np.stack([r * np.cos(ang), r * np.sin(ang)], 1), density=1.0, restitution=rng.uniform(0, 0.4))
        sp = rng.uniform(5, 60)
        th = rng.uniform(0, 2 * np.pi)
        w.add_body(DYNAMIC, pos=pos, angle=rng.uniform(-3, 3), vel=(sp * np.cos(th), sp * np.sin(th)), w=rng.uniform(-20, 20),
                   acc=GRAVITY, bullet=bool(rng.rand() < 0.3), shapes=[sh])
    return w


SCENES = {
    "shooter": shooter,
    "tumbler": lambda **k: tumbler(40, **k),
    "hello": hello_world,
    "pyramid5": lambda **k: pyramid(5, **k),
    "pyramid10": lambda **k: pyramid(10, **k),
    "disks100": lambda **k: disks_box(100, **k),
    "mixed100": lambda **k: mixed(100, **k),
}
