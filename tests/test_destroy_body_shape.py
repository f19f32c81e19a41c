"""Destroy(BodyID) (AabbTreeWorld.cpp:1062), Destroy(ShapeID) (:1259), Detach / SetEnabled / SetType (SetBody :2723-2795) and
SetShape with new geometry or another filter (:1284-1356) = pr_destroy_bodies / pr_destroy_shapes / pr_detach_shapes /
pr_set_body_enabled / pr_set_body_type / pr_set_shape, with the re-use of freed body, shape and DynamicTree leaf ids.
A pile of boxes and disks (some bodies carry two shapes) is edited in mid-run through the reference's own API; reference ==
restatement over the whole trajectory, bit for bit (bodies, contact ids and keys, manifolds, leaf ids), and the device ==
the restatement fed the device's colour-sorted order (emulated kernels here, the B200 under -m gpu)."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W

NB = 30


def _build():
    ref = RefWorld()
    ground = ref.add_edge((-20.0, 0.0), (20.0, 0.0))
    ref.add_body(STATIC, shapes=[ground])
    ref.box = ref.add_box(0.35, 0.35, density=2.0)
    ref.disk = ref.add_disk(0.4, density=1.0)
    ref.small = ref.add_disk(0.2, loc=(0.3, 0.3), density=3.0)
    ref.wide = ref.add_box(0.6, 0.2, density=1.0)
    bodies = []
    for i in range(NB):
        shapes = [ref.box if i % 2 else ref.disk]
        if i % 5 == 0:
            shapes.append(ref.small)
        if i % 7 == 3:
            shapes = [ref.wide]
        bodies.append(ref.add_body(DYNAMIC, pos=(-3.0 + 0.8 * (i % 8), 0.6 + 0.9 * (i // 8) + 0.01 * i), acc=(0.0, -9.8), shapes=shapes))
    return ref, bodies



def _edits(s, ref, others, mirror_rows):
    """The edit script. ref: the reference world (always edited through its own API: it defines what the new ids and shape values
    are); others: worlds that get the same edit through the C ABI; mirror_rows: copy the reference's body rows over (only when
    the others follow the reference's own trajectory)."""
    def rows_to(ids):
        if mirror_rows:
            for b in ids:
                row = ref.bodies()[b:b + 1]
                for o in others:
                    o.set_bodies([b], row)
    if s == 40:  # three bodies of the pile go, one after the other
        for b in (6, 13, 21):
            ref.destroy_body(b)
        for o in others:
            o.destroy_bodies([6, 13, 21])
    if s == 55:  # two new bodies: they get the freed ids back, last freed first, and their proxies the freed leaf ids
        for k, x in enumerate((-1.0, 1.5)):
            bid = ref.add_body(DYNAMIC, pos=(x, 6.0), acc=(0.0, -9.8), shapes=[ref.box, ref.small])
            assert bid == (21, 13)[k]
            row = ref.bodies()[bid:bid + 1]
            for o in others:
                o.add_body(row, [ref.box, ref.small])
    if s == 70:  # Detach: body 1 (disk + small) loses the small disk, body 16 (box + small) the box
        ref.detach_shape(1, ref.small); ref.detach_shape(16, ref.box)
        for o in others:
            o.detach_shapes(1, [ref.small]); o.detach_shapes(16, [ref.box])
        rows_to([1, 16])
    if s == 85:  # Destroy(ShapeID): every body that has the wide box loses it
        users = [b for b in range(1, NB + 1) if (b - 1) % 7 == 3 and b not in (6, 13, 21)]
        ref.destroy_shape(ref.wide)
        for o in others:
            o.destroy_shapes([ref.wide])
        rows_to(users)
    if s == 95:  # a new shape takes the freed shape id; attached to a body that lost its shape
        sid = ref.add_box(0.3, 0.5, density=1.5)
        assert sid == ref.wide
        row, ch, vt, nm = ref.export_shape(sid)
        for o in others:
            o.add_shape(row, ch, vt, nm)
        ref.attach_shape(4, sid)
        for o in others:
            o.attach_shapes(4, [sid])
        rows_to([4])
    if s == 105:
        ref.set_body_enabled(9, False)
        for o in others:
            o.set_body_enabled(9, False)
        rows_to([9])
    if s == 125:
        ref.set_body_enabled(9, True)
        for o in others:
            o.set_body_enabled(9, True)
        rows_to([9])
    if s == 115:  # a dynamic body of the pile turns static
        ref.set_body_type(2, 0)
        if mirror_rows:
            row = ref.bodies()[2:3]
        else:
            row = others[0].bodies()[2:3].copy()
            row["vel_x"] = 0; row["vel_y"] = 0; row["vel_w"] = 0; row["inv_mass"] = 0; row["inv_rot_i"] = 0
            row["acc_x"] = 0; row["acc_y"] = 0; row["acc_w"] = 0; row["under_active_time"] = 0
            row["pos0_x"] = row["pos1_x"]; row["pos0_y"] = row["pos1_y"]; row["pos0_a"] = row["pos1_a"]; row["alpha0"] = 0  # (a static body's sweep does not move)
            row["flags"] = (row["flags"] & ~np.uint32(abi.BODY_VELOCITY | abi.BODY_ACCELERATION | abi.BODY_AWAKE)) | np.uint32(abi.BODY_IMPENETRABLE)
        for o in others:
            o.set_body_type(2, row)
    if s == 135:  # SetShape with new geometry: the disk moves off-centre for every body that has it
        ref.translate_shape(ref.disk, 0.15, 0.1)
        row, ch, vt, nm = ref.export_shape(ref.disk)
        for o in others:
            o.set_shape(ref.disk, row, ch, vt, nm)
    if s == 150:  # SetShape with another filter: boxes stop colliding with the ground's category... (ground keeps 1)
        ref.set_shape_filter(ref.box, 2, 0xFFFE, 0)
        row, _, _, _ = ref.export_shape(ref.box)
        for o in others:
            o.set_shape(ref.box, row)


def _cmp(a, b, what, ref_side=False):
    la, lb = a.bodies(), b.bodies()
    dead = (la["flags"] & np.uint32(abi.BODY_DESTROYED)) != 0
    assert np.array_equal(dead, (lb["flags"] & np.uint32(abi.BODY_DESTROYED)) != 0), what + ": destroyed sets differ"
    assert_bits(la[~dead], lb[~dead], what + " bodies")
    ac, am = a.contacts(); bc, bm = b.contacts()
    assert_bits(ac, bc, what + " contacts")
    assert_bits(am, bm, what + " manifolds")
    ap, bp = a.proxies(), b.proxies()
    if ref_side:
        ap = ap[np.argsort(ap["tree_id"])]; bp = bp[np.argsort(bp["tree_id"])]
    assert_bits(ap, bp, what + " proxies")


def _run_reference():
    ref, bodies = _build()
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    for s in range(175):
        _edits(s, ref, [orc], True)
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "step %d stats (reference vs restatement)" % s)
        _cmp(ref, orc, "step %d (reference vs restatement)" % s, ref_side=True)
    return ref


def _run_device(lib=None, device=-1):
    ref, bodies = _build()  # (edited, never stepped: it hands out the new ids and shape values)
    sc = ref.export_scene()
    orc = refworld.OracleWorld(sc)
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    conf = refworld.default_step_conf()
    for s in range(175):
        _edits(s, ref, [gw, orc], False)
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        assert_bits(np.array([gs]), np.array([os_]), "step %d stats" % s)
        _cmp(gw, orc, "step %d" % s)
    return gw


@need_ref
@need_oracle
def test_restatement_structural_edits_equal_reference():
    ref = _run_reference()
    n = ref.counts()
    assert n["bodies"] == NB + 1  # the two new bodies re-used freed ids; one id is still free


@need_ref
@need_oracle
@need_emul
def test_emulated_structural_edits_equal_oracle():
    gw = _run_device(lib=emul_lib())
    ids = np.array([6], np.uint32)
    assert gw.lib.pr_destroy_bodies(gw.h, 1, abi.ptr(ids)) == -1   # destroyed already (the reference throws WasDestroyed)
    ids = np.array([1000], np.uint32)
    assert gw.lib.pr_destroy_bodies(gw.h, 1, abi.ptr(ids)) == -2


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_structural_edits_equal_oracle():
    _run_device(device=0)
