"""Random structural edits in mid-run -- CreateJoint of five types, Destroy(JointID), Destroy(BodyID) with its joints, CreateBody (freed ids come
back), SetJoint that re-hooks a revolute joint to another body -- every fourth step of a small pile: reference == restatement and device == its
twin restatement, bit for bit after every edit and every step, and the joint table's ids and bodies the reference's on every side. Also:
Attach / Detach of a second shape (also on a disabled body), SetEnabled and SetType toggles, ShiftOrigin, SetShape with another filter, SetVelocity.
The sweep found two things, both fixed and now pinned here: a contact read back after a SetBody that moved its body lacked the
needs-updating flag (seed 1), and DestroyContacts freed the overlap half's and the filter half's contacts in one pass instead of two, which
changes the ids later contacts get (seed 89).
`python tests/test_fuzz_mutators.py LO HI` sweeps seeds (330 seeds of the first ten kinds of edit and 200 of all thirteen run clean)."""
import sys
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi, world as W

G = (0.0, -10.0)


def run(seed, steps=90, verbose=False, lib=None, device=-1):
    rng = np.random.RandomState(seed)
    ref = RefWorld()
    ground = ref.add_body(STATIC, pos=(0.0, 0.0), shapes=[ref.add_box(12.0, 0.5, center=(0.0, -0.5))])
    box = ref.add_box(0.3, 0.25, density=2.0)
    disk = ref.add_disk(0.25, density=1.0)
    small = ref.add_disk(0.12, loc=(0.3, 0.0), density=1.0)
    live = [ref.add_body(DYNAMIC, acc=G, pos=(float(rng.uniform(-5, 5)), float(rng.uniform(0.4, 3.5))), shapes=[box if i % 2 else disk] + ([small] if i % 3 == 0 else [])) for i in range(9)]
    has_small = {b: (i % 3 == 0) for i, b in enumerate(live)}
    disabled = set(); static = set()
    shape_of = {b: (box if i % 2 else disk) for i, b in enumerate(live)}
    sc = ref.export_scene()
    orc = refworld.OracleWorld(sc)
    gw = W.World(sc, lib=lib if (lib is not None or device >= 0) else emul_lib(), device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    orc2 = refworld.OracleWorld(sc)
    others = [orc, gw, orc2]
    conf = refworld.default_step_conf()
    joints = []

    def pos(b):
        r = ref.bodies()[b]
        return (float(r["xf_px"]), float(r["xf_py"]))

    def new_joint():
        if len(live) < 2:
            return
        a, b = rng.choice(live, 2, replace=False)
        a, b = int(a), int(b)
        if rng.rand() < 0.25:
            a = ground
        kind = int(rng.randint(5))
        pa, pb = pos(a), pos(b)
        mid = ((pa[0] + pb[0]) / 2, (pa[1] + pb[1]) / 2)
        if kind == 0:
            jid = ref.add_revolute(a, b, mid)
        elif kind == 1:
            jid = ref.add_distance(a, b, pa, pb, frequency=float(rng.choice([0.0, 3.0])), damping_ratio=0.5)
        elif kind == 2:
            jid = ref.add_weld(a, b, mid)
        elif kind == 3:
            jid = ref.add_rope(a, b, (0.0, 0.0), (0.0, 0.0), float(np.hypot(pb[0] - pa[0], pb[1] - pa[1]) + 0.3))
        else:
            jid = ref.add_target(b, (pb[0] + 0.5, pb[1] + 1.0), pb, max_force=100.0)
        row = ref.joints()[[jid]]
        for o in others:
            o.create_joints(row)
        if jid not in joints:
            joints.append(jid)
        if verbose:
            print("  new joint", jid, "kind", kind, a, b)

    def cmp(where):
        for o, name in ((orc, "restatement"),):
            la, lb = ref.bodies(), o.bodies()
            dead = (la["flags"] & np.uint32(abi.BODY_DESTROYED)) != 0
            assert np.array_equal(dead, (lb["flags"] & np.uint32(abi.BODY_DESTROYED)) != 0), where + " destroyed sets (%s)" % name
            assert_bits(la[~dead], lb[~dead], where + " bodies (reference vs %s)" % name)
            assert_bits(ref.joints(), o.joints(), where + " joints (reference vs %s)" % name)
            rc, rm = ref.contacts(); oc, om = o.contacts()
            assert_bits(rc, oc, where + " contacts (reference vs %s)" % name)
            assert_bits(rm, om, where + " manifolds (reference vs %s)" % name)
        la, lb = gw.bodies(), orc2.bodies()
        assert_bits(la, lb, where + " bodies (device vs twin)")
        assert_bits(gw.joints(), orc2.joints(), where + " joints (device vs twin)")
        gc, gm = gw.contacts(); oc, om = orc2.contacts()
        assert_bits(gc, oc, where + " contacts (device vs twin)")
        assert_bits(gm, om, where + " manifolds (device vs twin)")
        # ids must be the reference's on every side
        assert np.array_equal(ref.joints()["body_a"], gw.joints()["body_a"]) and np.array_equal(ref.joints()["body_b"], gw.joints()["body_b"]), where + " joint tables (reference vs device)"

    for s in range(steps):
        if s % 4 == 1:
            op = int(rng.randint(13))
            if verbose:
                print("step", s, "op", op)
            if op in (0, 1):
                new_joint()
            elif op == 2 and joints:
                jid = int(rng.choice(joints)); joints.remove(jid)
                ref.destroy_joint(jid)
                for o in others:
                    o.destroy_joints([jid])
            elif op == 3 and len(live) > 3:
                b = int(rng.choice(live)); live.remove(b); disabled.discard(b); static.discard(b)
                gone = [j for j in joints if int(ref.joints()[j]["body_a"]) == b or int(ref.joints()[j]["body_b"]) == b]
                for j in gone:
                    joints.remove(j)
                ref.destroy_body(b)
                for o in others:
                    o.destroy_bodies([b])
            elif op == 4 and True:
                sh = box if rng.rand() < 0.5 else disk
                bid = ref.add_body(DYNAMIC, acc=G, pos=(float(rng.uniform(-5, 5)), float(rng.uniform(3.0, 5.0))), shapes=[sh])
                row = ref.bodies()[bid:bid + 1]
                for o in others:
                    o.add_body(row, [sh])
                live.append(bid); shape_of[bid] = sh; has_small[bid] = False
            elif op == 6:
                b = int(rng.choice(live))
                if has_small.get(b):
                    ref.detach_shape(b, small)
                    for o in others:
                        o.detach_shapes(b, [small])
                    has_small[b] = False
                else:
                    ref.attach_shape(b, small)
                    for o in others:
                        o.attach_shapes(b, [small])
                    has_small[b] = True
                row = ref.bodies()[b:b + 1]
                for o in others:
                    o.set_bodies([b], row)
            elif op == 7:
                b = int(rng.choice(live))
                flag = b in disabled
                ref.set_body_enabled(b, flag)
                for o in others:
                    o.set_body_enabled(b, flag)
                row = ref.bodies()[b:b + 1]
                for o in others:
                    o.set_bodies([b], row)
                (disabled.discard if flag else disabled.add)(b)
            elif op == 8:
                b = int(rng.choice(live))
                to = 2 if b in static else 0
                ref.set_body_type(b, to)
                row = ref.bodies()[b:b + 1]
                for o in others:
                    o.set_body_type(b, row)
                (static.discard if to == 2 else static.add)(b)
            elif op == 10:
                x, y = float(rng.uniform(-2, 2)), float(rng.uniform(-1, 1))
                ref.shift_origin(x, y)
                for o in others:
                    o.shift_origin(x, y)
            elif op == 11:
                mask = 0xFFFF if rng.rand() < 0.5 else 0xFFFD
                cat = 1 if rng.rand() < 0.5 else 2
                ref.set_shape_filter(disk, cat, mask, 0)
                row, _, _, _ = ref.export_shape(disk)
                for o in others:
                    o.set_shape(disk, row)
            elif op == 12:
                b = int(rng.choice(live))
                ref.lib.ref_set_body_velocity(ref.h, b, float(rng.uniform(-3, 3)), float(rng.uniform(0, 4)), float(rng.uniform(-2, 2)))
                row = ref.bodies()[b:b + 1]
                for o in others:
                    o.set_bodies([b], row)
            elif op == 5 and joints:
                jid = int(rng.choice(joints))
                row = ref.joints()[[jid]].copy()
                if ((int(row["collide_connected"][0]) >> 8) & 0xFF) == 0 and len(live) > 2:
                    cand = [b for b in live if b != int(row["body_a"][0]) and b != int(row["body_b"][0])]
                    row["body_b"] = int(rng.choice(cand))
                    ref.set_joints([jid], row)
                    for o in others:
                        o.set_joints([jid], row)
            cmp("after the edit at step %d" % s)
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "step %d stats (reference vs restatement)" % s)
        gs = gw.step(conf)
        o2 = orc2.step(conf, order=gw.solve_order()[0])
        assert_bits(np.array([gs]), np.array([o2]), "step %d stats (device vs twin)" % s)
        cmp("step %d" % s)


@need_ref
@need_oracle
@need_emul
@pytest.mark.parametrize("seed", [1, 17, 89])
def test_emulated_random_edits_with_joints(seed):
    run(seed)


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_random_edits_with_joints():
    run(17, device=0)


if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(lo, hi):
        try:
            run(seed)
        except AssertionError as e:
            bad += 1
            print("seed", seed, "FAILED:", str(e)[:400], flush=True)
    print("seeds", lo, hi, "failures", bad)
