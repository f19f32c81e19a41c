"""The joint types of SURVEY 8(f) rank 1 added after the distance / rope / friction / motor joints, each as a scene stepped by
the real reference, the restatement, the emulated kernels and the GPU -- bit for bit, incl. the joints' accumulated impulses:
  weld  (WeldJointConf.cpp:35-371): a cantilever of boxes welded to a static wall (rigid), a soft-welded pair, a welded pair with
        fixed rotation (K.ez.z == 0 -> the 2x2 branch), all under gravity and hit by a falling disk.
  target (TargetJointConf.cpp:52-243): one-body "mouse" joints dragging boxes through a pile of disks -- a strong one, one
        whose force cap saturates, one on a body that also hangs from a weld; the targets are moved with SetJoint in mid-run.
  prismatic (PrismaticJointConf.cpp:150-533): sliders on a static rail -- free, with limits (lower hit under gravity, upper hit by a motor),
        equal limits, a motor that saturates, a slider between two dynamic bodies, one with fixed rotation (k22 == 0 -> 1).
  wheel (WheelJointConf.cpp:103-362): a car (chassis + two wheels on sprung wheel joints, one driven by a motor) on a bumpy ground,
        and a rigid (frequency 0) wheel joint.
  pulley (PulleyJointConf.cpp:91-282): two unequal weights over a pulley (ratio 1), a ratio-2 block and tackle, a pulley between a
        body and a static one; ShiftOrigin in mid-run moves the ground anchors (and a target joint's target).
  gear (GearJointConf.cpp:171-468; four bodies): two wheels on revolute joints geared 2:1 (one driven by a motor), a rack and pinion
        (revolute + prismatic), two sliders geared against each other (prismatic + prismatic), disks dropped on them."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import world as W


G = (0.0, -10.0)


def _weld_scene():
    ref = RefWorld()
    wall = ref.add_body(STATIC, pos=(0.0, 0.0))
    box = ref.add_box(0.5, 0.125, density=20.0)
    disk = ref.add_disk(0.3, density=5.0)
    prev = wall
    for i in range(4):                                     # cantilever: rigid welds
        b = ref.add_body(DYNAMIC, acc=G, pos=(0.5 + 1.0 * i, 5.0), shapes=[box])
        ref.add_weld(prev, b, (1.0 * i, 5.0))
        prev = b
    s1 = ref.add_body(DYNAMIC, acc=G, pos=(-3.0, 5.0), shapes=[box])
    s2 = ref.add_body(DYNAMIC, acc=G, pos=(-4.0, 5.0), shapes=[box])
    ref.add_weld(wall, s1, (-2.5, 5.0), frequency=5.0, damping_ratio=0.7)   # soft angular spring
    ref.add_weld(s1, s2, (-3.5, 5.0), frequency=8.0, damping_ratio=0.2)
    f1 = ref.add_body(DYNAMIC, acc=G, pos=(0.0, 9.0), shapes=[box], fixed_rotation=True)
    f2 = ref.add_body(DYNAMIC, acc=G, pos=(1.0, 9.0), shapes=[box], fixed_rotation=True)
    ref.add_weld(f1, f2, (0.5, 9.0))                       # no rotational inertia at all: GetInverse22
    ref.add_body(DYNAMIC, acc=G, pos=(3.2, 8.0), shapes=[disk])   # falls on the cantilever's tip
    ref.add_body(DYNAMIC, acc=G, pos=(-3.9, 8.0), shapes=[disk])  # and on the soft one
    return ref, [5, 5, 5, 5, 5, 5, 5]


def _target_scene():
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(10.0, 0.5, center=(0.0, -0.5))])
    box = ref.add_box(0.4, 0.3, density=4.0)
    disk = ref.add_disk(0.25, density=1.0)
    for i in range(12):
        ref.add_body(DYNAMIC, acc=G, pos=(-2.0 + 0.55 * (i % 6), 0.3 + 0.55 * (i // 6)), shapes=[disk])
    b1 = ref.add_body(DYNAMIC, acc=G, pos=(-5.0, 0.35), shapes=[box])
    b2 = ref.add_body(DYNAMIC, acc=G, pos=(5.0, 0.35), shapes=[box])
    b3 = ref.add_body(DYNAMIC, acc=G, pos=(0.0, 4.0), shapes=[box])
    b4 = ref.add_body(DYNAMIC, acc=G, pos=(0.8, 4.0), shapes=[box])
    ref.add_target(b1, (-5.0, 0.35), (-5.0, 0.35), max_force=2000.0)                       # strong: pulled through the pile
    ref.add_target(b2, (5.0, 0.35), (5.2, 0.5), frequency=3.0, damping_ratio=0.3, max_force=30.0)   # saturates
    ref.add_weld(b3, b4, (0.4, 4.0))
    ref.add_target(b3, (0.0, 4.0), (-0.3, 4.2), frequency=4.0, damping_ratio=1.0, max_force=500.0)
    # a target joint whose conf names a bodyA: it only takes part in the island rules -- a disabled one keeps the joint from being solved
    # (AabbTreeWorld.cpp:1443-1466; UnitTests/World.cpp:3377-3396), an enabled static one does not
    spare = ref.add_body(STATIC, pos=(9.0, 9.0), enabled=False)
    b5 = ref.add_body(DYNAMIC, acc=G, pos=(-8.0, 3.0), shapes=[box])
    b6 = ref.add_body(DYNAMIC, acc=G, pos=(8.0, 3.0), shapes=[box])
    ref.add_target(b5, (-8.0, 6.0), (-8.0, 3.0), max_force=500.0, body_a=spare)
    ref.add_target(b6, (8.0, 6.0), (8.0, 3.0), max_force=500.0, body_a=ground)
    return ref, [6, 6, 5, 6, 6, 6]


def _target_moves(step):
    """(joint id, new target) pairs to apply before the given step"""
    if step % 10 == 0 and step < 160:
        k = step // 10
        return [(0, (-5.0 + 0.6 * k, 0.35 + 0.02 * k)), (1, (5.0 - 0.5 * k, 0.35)), (3, (0.3 * k, 4.0 - 0.1 * k))]
    return []


def _prismatic_scene():
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(20.0, 0.5, center=(0.0, -0.5))])
    box = ref.add_box(0.4, 0.3, density=3.0)
    disk = ref.add_disk(0.25, density=2.0)
    def dyn(pos, **kw):
        return ref.add_body(DYNAMIC, acc=G, pos=pos, shapes=[box], **kw)
    s0 = dyn((-8.0, 3.0)); ref.add_prismatic(ground, s0, (-8.0, 3.0), (0.0, 1.0))                                   # free vertical slider: falls to the ground
    s1 = dyn((-6.0, 3.0)); ref.add_prismatic(ground, s1, (-6.0, 3.0), (0.0, 1.0), limits=(-1.0, 1.0))               # lower limit stops it
    s2 = dyn((-4.0, 3.0)); ref.add_prismatic(ground, s2, (-4.0, 3.0), (0.6, 0.8), limits=(-0.5, 1.5), motor=(2.0, 500.0))   # motor drives it to the upper limit
    s3 = dyn((-2.0, 3.0)); ref.add_prismatic(ground, s3, (-2.0, 3.0), (1.0, 0.0), limits=(0.0, 0.0))                # equal limits
    s4 = dyn((0.0, 3.0)); ref.add_prismatic(ground, s4, (0.0, 3.0), (0.0, 1.0), motor=(5.0, 10.0))                  # motor too weak against gravity: saturates
    s5 = dyn((2.0, 2.0)); s6 = dyn((3.0, 2.6))
    ref.add_prismatic(s5, s6, (2.5, 2.3), (0.8, 0.6), limits=(-0.3, 0.6))                                           # two dynamic bodies
    s7 = dyn((5.0, 3.0), fixed_rotation=True); s8 = dyn((5.0, 4.0), fixed_rotation=True)
    ref.add_prismatic(s7, s8, (5.0, 3.5), (1.0, 0.0), limits=(-0.4, 0.4))                                           # no rotational inertia
    ref.add_body(DYNAMIC, acc=G, pos=(-5.9, 6.0), shapes=[disk])
    ref.add_body(DYNAMIC, acc=G, pos=(2.6, 6.0), shapes=[disk])
    ref.add_body(DYNAMIC, acc=G, pos=(5.3, 7.0), shapes=[disk])
    return ref, [7] * 7


def _wheel_scene():
    ref = RefWorld()
    gs = [ref.add_box(30.0, 0.5, center=(0.0, -0.5))] + [ref.add_box(0.3, 0.08, center=(2.0 + 1.7 * i, 0.08)) for i in range(8)]
    ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=gs)
    chassis = ref.add_box(1.5, 0.25, density=1.0)
    wheel = ref.add_disk(0.4, density=1.0, friction=0.9)
    car = ref.add_body(DYNAMIC, acc=G, pos=(0.0, 1.0), shapes=[chassis])
    w1 = ref.add_body(DYNAMIC, acc=G, pos=(-1.0, 0.45), shapes=[wheel])
    w2 = ref.add_body(DYNAMIC, acc=G, pos=(1.0, 0.45), shapes=[wheel])
    ref.add_wheel(car, w1, (-1.0, 0.45), (0.0, 1.0), frequency=4.0, damping_ratio=0.7, motor=(-12.0, 20.0))     # driven
    ref.add_wheel(car, w2, (1.0, 0.45), (0.0, 1.0), frequency=4.0, damping_ratio=0.7)                           # free-rolling
    cart = ref.add_body(DYNAMIC, acc=G, pos=(-6.0, 1.0), shapes=[chassis])
    w3 = ref.add_body(DYNAMIC, acc=G, pos=(-6.0, 0.45), shapes=[wheel])
    ref.add_wheel(cart, w3, (-6.0, 0.45), (0.6, 0.8), frequency=0.0, motor=(3.0, 1.0))                          # no spring; weak motor saturates
    return ref, [8, 8, 8]


def _pulley_scene():
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(20.0, 0.5, center=(0.0, -0.5))])
    light = ref.add_box(0.4, 0.4, density=1.0)
    heavy = ref.add_box(0.4, 0.4, density=2.5)
    a = ref.add_body(DYNAMIC, acc=G, pos=(-3.0, 3.0), shapes=[light])
    b = ref.add_body(DYNAMIC, acc=G, pos=(-1.0, 3.0), shapes=[heavy])
    ref.add_pulley(a, b, (-3.0, 8.0), (-1.0, 8.0), (-3.0, 3.4), (-1.0, 3.4))                   # the heavy one sinks to the ground, lifting the light one
    c = ref.add_body(DYNAMIC, acc=G, pos=(2.0, 4.0), shapes=[light])
    d = ref.add_body(DYNAMIC, acc=G, pos=(4.0, 4.0), shapes=[light])
    ref.add_pulley(c, d, (2.0, 9.0), (4.5, 9.0), (2.0, 4.4), (4.2, 4.3), ratio=2.0)            # block and tackle, slanted rope
    e = ref.add_body(DYNAMIC, acc=G, pos=(7.0, 3.0), vel=(1.0, 0.0), shapes=[heavy])
    ref.add_pulley(e, ground, (7.0, 8.0), (9.0, 8.0), (7.0, 3.4), (9.0, 0.0))                  # the other end is static: a pendulum of fixed length
    f = ref.add_body(DYNAMIC, acc=G, pos=(-7.0, 2.0), shapes=[light])
    ref.add_target(f, (-7.0, 2.0), (-6.0, 4.0), max_force=200.0)
    return ref, [9, 9, 9, 6]


def _gear_scene():
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(20.0, 0.5, center=(0.0, -0.5))])
    big = ref.add_disk(1.0, density=1.0)
    small = ref.add_disk(0.5, density=1.0)
    bar = ref.add_box(0.25, 1.0, density=2.0)
    pebble = ref.add_disk(0.2, density=3.0)
    w1 = ref.add_body(DYNAMIC, acc=G, pos=(-6.0, 4.0), shapes=[big])
    w2 = ref.add_body(DYNAMIC, acc=G, pos=(-4.5, 4.0), shapes=[small])
    r1 = ref.add_revolute(ground, w1, (-6.0, 4.0), enable_motor=True, motor_speed=1.5, max_motor_torque=50.0)
    r2 = ref.add_revolute(ground, w2, (-4.5, 4.0))
    ref.add_gear(r1, r2, ratio=0.5)                                       # revolute + revolute: w2 turns twice as fast, the other way
    w3 = ref.add_body(DYNAMIC, acc=G, pos=(0.0, 4.0), w=2.0, shapes=[big])
    rack = ref.add_body(DYNAMIC, acc=G, pos=(1.25, 4.0), shapes=[bar])
    r3 = ref.add_revolute(ground, w3, (0.0, 4.0))
    p1 = ref.add_prismatic(ground, rack, (1.25, 4.0), (0.0, 1.0), limits=(-2.0, 2.0))
    ref.add_gear(r3, p1, ratio=-1.0)                                      # rack and pinion: the rack's weight turns the wheel
    s1 = ref.add_body(DYNAMIC, acc=G, pos=(5.0, 3.0), shapes=[bar])
    s2 = ref.add_body(DYNAMIC, acc=G, pos=(7.0, 3.0), shapes=[ref.add_box(0.25, 1.0, density=3.0)])
    p2 = ref.add_prismatic(ground, s1, (5.0, 3.0), (0.0, 1.0), limits=(-1.5, 3.0))
    p3 = ref.add_prismatic(ground, s2, (7.0, 3.0), (0.6, 0.8), limits=(-1.5, 3.0))
    ref.add_gear(p2, p3, ratio=1.0)                                       # a counterweight: the heavier slider sinks, lifting the other
    ref.add_body(DYNAMIC, acc=G, pos=(-6.4, 7.0), shapes=[pebble])
    ref.add_body(DYNAMIC, acc=G, pos=(0.3, 7.5), shapes=[pebble])
    ref.add_body(DYNAMIC, acc=G, pos=(5.0, 7.0), shapes=[pebble])
    return ref, [0, 0, 10, 0, 7, 10, 7, 7, 10]


def _chains_scene():
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 10.0))
    link = ref.add_box(0.3, 0.05, density=5.0)
    types = []
    for c in range(40):
        prev = ground
        x0 = -20.0 + 1.0 * c
        for k in range(3):
            b = ref.add_body(DYNAMIC, acc=G, pos=(x0 + 0.3 + 0.6 * k, 10.0 - 0.01 * c), w=0.1 * (c % 5), shapes=[link])
            ref.add_revolute(prev, b, (x0 + 0.6 * k, 10.0 - 0.01 * c))
            types.append(0)
            prev = b
    return ref, types


def _longchain_scene():
    """one chain of 60 links: more joint levels than the device launches (48) -- one thread walks the joints"""
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 10.0))
    link = ref.add_box(0.2, 0.05, density=5.0)
    prev = ground
    for k in range(60):
        b = ref.add_body(DYNAMIC, acc=G, pos=(0.2 + 0.4 * k, 10.0), shapes=[link])
        ref.add_revolute(prev, b, (0.4 * k, 10.0))
        prev = b
    return ref, [0] * 60


def _fuzz_scene(seed):
    """random joints of every type (but gear, which needs two geared joints) between random bodies of a small pile: limits, motors, soft and
    rigid variants, joints sharing bodies in any order -- whatever the generator comes up with must be the reference's bits"""
    def build():
        rng = np.random.RandomState(1000 + seed)
        ref = RefWorld()
        ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(12.0, 0.5, center=(0.0, -0.5))])
        shapes = [ref.add_box(0.3, 0.2, density=2.0), ref.add_disk(0.25, density=1.5), ref.add_box(0.15, 0.4, density=3.0)]
        bodies = []
        for i in range(10):
            pos = (float(rng.uniform(-5.0, 5.0)), float(rng.uniform(0.4, 4.0)))
            bodies.append(ref.add_body(DYNAMIC, acc=G, pos=pos, angle=float(rng.uniform(-0.5, 0.5)), vel=(float(rng.uniform(-1, 1)), 0.0),
                                       w=float(rng.uniform(-1, 1)), shapes=[shapes[int(rng.randint(3))]], fixed_rotation=bool(rng.rand() < 0.15)))
        types = []
        rows = ref.bodies()
        def p(b):
            return (float(rows[b]["xf_px"]), float(rows[b]["xf_py"]))
        for _ in range(9):
            kind = int(rng.randint(10))
            a = ground if rng.rand() < 0.3 else bodies[int(rng.randint(10))]
            b = bodies[int(rng.randint(10))]
            if a == b:
                continue
            pa, pb = p(a) if a != ground else (p(b)[0], p(b)[1] + 1.0), p(b)
            mid = ((pa[0] + pb[0]) / 2.0, (pa[1] + pb[1]) / 2.0)
            ax = rng.uniform(-1, 1, 2); ax = ax / max(1e-3, float(np.hypot(*ax)))
            axis = (float(ax[0]), float(ax[1]))
            if kind == 0:
                ref.add_revolute(a, b, mid, enable_motor=bool(rng.rand() < 0.5), motor_speed=float(rng.uniform(-2, 2)), max_motor_torque=float(rng.uniform(0, 20))); types.append(0)
            elif kind == 1:
                ref.add_distance(a, b, pa, pb, frequency=float(rng.choice([0.0, 3.0])), damping_ratio=0.4); types.append(1)
            elif kind == 2:
                ref.add_weld(a, b, mid, frequency=float(rng.choice([0.0, 4.0])), damping_ratio=0.5); types.append(5)
            elif kind == 3:
                ref.add_friction(a, b, mid, float(rng.uniform(0.5, 20)), float(rng.uniform(0.1, 5))); types.append(3)
            elif kind == 4:
                ref.add_motor(a, b, (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1))), float(rng.uniform(-1, 1)), float(rng.uniform(1, 40)), float(rng.uniform(0.5, 20))); types.append(4)
            elif kind == 5:
                lim = (float(rng.uniform(-1.0, -0.1)), float(rng.uniform(0.1, 1.0))) if rng.rand() < 0.6 else None
                mot = (float(rng.uniform(-2, 2)), float(rng.uniform(1, 60))) if rng.rand() < 0.5 else None
                ref.add_prismatic(a, b, mid, axis, limits=lim, motor=mot); types.append(7)
            elif kind == 6:
                mot = (float(rng.uniform(-6, 6)), float(rng.uniform(0.5, 20))) if rng.rand() < 0.5 else None
                ref.add_wheel(a, b, pb, axis, frequency=float(rng.choice([0.0, 3.0])), damping_ratio=0.6, motor=mot); types.append(8)
            elif kind == 7:
                ref.add_pulley(a, b, (pa[0], pa[1] + 3.0), (pb[0], pb[1] + 3.0), pa, pb, ratio=float(rng.choice([1.0, 1.5]))); types.append(9)
            elif kind == 8:
                ref.add_target(b, (pb[0] + float(rng.uniform(-1, 1)), pb[1] + float(rng.uniform(0, 1.5))), pb, frequency=float(rng.uniform(2, 6)), max_force=float(rng.uniform(20, 400))); types.append(6)
            else:
                ref.add_rope(a, b, (0.0, 0.0), (0.0, 0.0), float(np.hypot(pb[0] - pa[0], pb[1] - pa[1]) + rng.uniform(0.0, 0.5))); types.append(2)
        return ref, types
    return build


SHIFTS = {"pulley": {60: (1.5, -0.75), 130: (-100.0, 33.3)}}
SCENES = {"chains": (_chains_scene, 120), "longchain": (_longchain_scene, 60), "gear": (_gear_scene, 240), "pulley": (_pulley_scene, 200), "weld": (_weld_scene, 200), "target": (_target_scene, 220), "prismatic": (_prismatic_scene, 240), "wheel": (_wheel_scene, 300)}
MOVES = {"target": _target_moves}
def _fuzz_gear_scene(seed):
    """a fuzz scene plus a gear joint between two of its revolute / prismatic joints (whatever bodies those tie)"""
    def build():
        ref, types = _fuzz_scene(seed)()
        rng = np.random.RandomState(seed)
        cand = [i for i, ty in enumerate(types) if ty in (0, 7)]
        assert len(cand) >= 2
        a, b = rng.choice(cand, 2, replace=False)
        ref.add_gear(int(a), int(b), ratio=float(rng.choice([1.0, -1.0, 2.0, 0.5])))
        return ref, types + [10]
    return build


for _seed in range(6):
    SCENES["fuzz%d" % _seed] = (_fuzz_scene(_seed), 150)
for _seed in (103, 110):
    SCENES["fuzzgear%d" % _seed] = (_fuzz_gear_scene(_seed), 120)


def _run(name, lib=None, device=-1, with_device=True):
    build, steps = SCENES[name]
    ref, types = build()
    sc = ref.export_scene()
    assert [(int(j["collide_connected"]) >> 8) & 0xFF for j in sc["joints"]] == types
    orc = refworld.OracleWorld(sc)
    twins = []
    if with_device:
        gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
        gw.set_toi_mode(0)
        twins = [gw, refworld.OracleWorld(sc)]
    conf = refworld.default_step_conf()
    for s in range(steps):
        for jid, tgt in (MOVES[name](s) if name in MOVES else []):
            rows = orc.joints()[[jid]].copy()
            rows["local_anchor_a"] = tgt          # (pr_target_joint: the target sits where the other rows have local_anchor_a)
            ids = np.array([jid], np.uint32)
            ref.set_target(jid, tgt)
            orc.set_joints(ids, rows)
            for t in twins:
                rows2 = t.joints()[[jid]].copy()
                rows2["local_anchor_a"] = tgt
                t.set_joints(ids, rows2)
        if name in SHIFTS and s in SHIFTS[name]:
            for t in [ref, orc] + twins:
                t.shift_origin(*SHIFTS[name][s])
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "step %d stats (reference vs restatement)" % s)
        assert_bits(ref.bodies(), orc.bodies(), "step %d bodies (reference vs restatement)" % s)
        assert_bits(ref.joints(), orc.joints(), "step %d joints (reference vs restatement)" % s)
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "step %d contacts (reference vs restatement)" % s)
        assert_bits(rm, om, "step %d manifolds (reference vs restatement)" % s)
        if twins:
            gw, orc2 = twins
            gs = gw.step(conf)
            o2 = orc2.step(conf, order=gw.solve_order()[0])
            assert_bits(np.array([gs]), np.array([o2]), "step %d stats" % s)
            assert_bits(gw.bodies(), orc2.bodies(), "step %d bodies" % s)
            assert_bits(gw.joints(), orc2.joints(), "step %d joints" % s)
            gc, gm = gw.contacts(); oc2, om2 = orc2.contacts()
            assert_bits(gc, oc2, "step %d contacts" % s)
            assert_bits(gm, om2, "step %d manifolds" % s)
    return ref


@need_ref
@need_oracle
@pytest.mark.parametrize("name", sorted(SCENES))
def test_restatement_equals_reference(name):
    _run(name, with_device=False)


@need_ref
@need_oracle
@need_emul
@pytest.mark.parametrize("name", sorted(SCENES))
def test_emulated_kernels_equal_oracle(name):
    _run(name, lib=emul_lib())


@pytest.mark.gpu
@need_ref
@need_oracle
@pytest.mark.parametrize("name", sorted(SCENES))
def test_gpu_equals_oracle(name):
    _run(name, device=0)


@need_ref
@need_emul
def test_joint_work_reads_back_what_init_velocity_left():
    """pr_get_joint_work: a fresh joint reads back unsolved; after a Step the wheel scene's driven joint carries its motor mass
    (1 / (invI_A + invI_B)), the point-to-line mass and the unit axes -- the values the reference keeps in the JointConf (GetAngularMass...)"""
    ref, _ = _wheel_scene()
    sc = ref.export_scene()
    gw = W.World(sc, lib=emul_lib(), device=-1)
    wk = gw.joint_work()
    assert (wk["island"] == 0xFFFFFFFF).all()
    gw.step(refworld.default_step_conf())
    wk = gw.joint_work()
    assert (wk["island"] != 0xFFFFFFFF).all()
    b = gw.bodies()
    j = gw.joints()
    ii = b["inv_rot_i"]
    expect = np.float32(1.0) / (ii[j["body_a"][0]] + ii[j["body_b"][0]])
    assert wk["angular_mass"][0] == expect            # the driven wheel: motor enabled
    assert wk["angular_mass"][1] == 0.0               # the free-rolling one: no motor mass
    for k in range(3):
        ay, ax = wk["m"][k][0:2], wk["m"][k][5:7]
        assert abs(float(np.hypot(*ay)) - 1.0) < 1e-6
        assert wk["m"][k][4] > 0.0                    # point-to-line mass
        if k < 2:
            assert abs(float(np.hypot(*ax)) - 1.0) < 1e-6 and abs(float(ay @ ax)) < 1e-6   # sprung: ax set, perpendicular to ay
        else:
            assert ax[0] == 0.0 and ax[1] == 0.0     # frequency 0: UnitVec::GetZero()
