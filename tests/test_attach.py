"""Attach(world, body, shape) -- SetBody with added shapes (AabbTreeWorld.cpp:2760-2786) = pr_attach_shapes + pr_set_bodies.
Bodies are created bare and get their shapes attached afterwards (the way most of the reference's own tests and Testbed
scenes build bodies), and a second shape is attached to bodies in mid-run: reference (its own Attach, mass data reset) ==
restatement == emulated kernels == GPU, bit for bit (proxy ids, contacts, trajectories)."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W


def _build_ref():
    """Ground + 24 dynamic bodies created WITHOUT shapes; shapes attached afterwards in a shuffled body order."""
    ref = RefWorld()
    ground = ref.add_edge((-20.0, 0.0), (20.0, 0.0))
    ref.add_body(STATIC, shapes=[ground])
    disk = ref.add_disk(0.4, density=1.0)
    box = ref.add_box(0.35, 0.35, density=2.0)
    small = ref.add_disk(0.2, loc=(0.3, 0.3), density=3.0)
    ref.small0 = ref.add_disk(0.2, loc=(0.3, 0.3), density=0.0)  # massless: attaching it leaves the body's mass data as it is
    bodies = []
    for i in range(24):
        bodies.append(ref.add_body(DYNAMIC, pos=(-6.0 + 0.9 * (i % 12), 1.0 + 1.1 * (i // 12) + 0.01 * i), acc=(0.0, -9.8)))
    return ref, bodies, disk, box, small


def _attach_all(ref, others, bodies, shape_of):
    """Attach in the reference, then mirror: attach the same (body, shape) and upload the reference's new body row."""
    for b in bodies:
        ref.attach_shape(b, shape_of(b))
        row = ref.bodies()[b:b + 1]
        for o in others:
            o.attach_shapes(b, [shape_of(b)])
            o.set_bodies([b], row)


def _run(lib=None, device=-1, with_device=True):
    ref, bodies, disk, box, small = _build_ref()
    sc = ref.export_scene()  # bare bodies: no shape lists yet
    orc = refworld.OracleWorld(sc)
    others = [orc]
    gw = None
    if with_device:
        gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
        gw.set_toi_mode(0)
        others.append(gw)
    order = [bodies[(7 * i) % 24] for i in range(24)]
    _attach_all(ref, others, order, lambda b: disk if b % 2 else box)
    conf = refworld.default_step_conf()
    for s in range(140):
        if s == 50:  # a second shape on every third body, in mid-run (piled up, in contact)
            _attach_all(ref, others, bodies[::3], lambda b: small)
        rs = ref.step(conf)
        os_ = orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "step %d stats (reference vs restatement)" % s)
        assert_bits(ref.bodies(), orc.bodies(), "step %d bodies (reference vs restatement)" % s)
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "step %d contacts (reference vs restatement)" % s)
        assert_bits(rm, om, "step %d manifolds (reference vs restatement)" % s)
        rp, op = ref.proxies(), orc.proxies()
        assert_bits(rp[np.argsort(rp["tree_id"])], op[np.argsort(op["tree_id"])], "step %d proxies (reference vs restatement)" % s)
    assert len(ref.proxies()) == 1 + 24 + 8
    return ref, orc


def _run_device(lib=None, device=-1):
    """Device vs restatement fed the colour-sorted order (its own oracle instance, same edits)."""
    ref, bodies, disk, box, small = _build_ref()
    sc = ref.export_scene()
    orc = refworld.OracleWorld(sc)
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    order = [bodies[(7 * i) % 24] for i in range(24)]
    _attach_all(ref, [orc, gw], order, lambda b: disk if b % 2 else box)
    conf = refworld.default_step_conf()
    for s in range(140):
        if s == 50:  # mid-run (the device's trajectory is its own: no reference rows to mirror): a massless second shape
            for b in bodies[::3]:
                for o in (orc, gw):
                    o.attach_shapes(b, [ref.small0])
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert_bits(gc, oc, "step %d contacts" % s)
        assert_bits(gm, om, "step %d manifolds" % s)
        assert_bits(gw.proxies(), orc.proxies(), "step %d proxies" % s)
    assert len(gw.proxies()) == 1 + 24 + 8


@need_ref
@need_oracle
def test_restatement_attach_equals_reference():
    _run(with_device=False)


@need_ref
@need_oracle
@need_emul
def test_emulated_attach_equals_oracle():
    _run_device(lib=emul_lib())


@need_emul
def test_attach_refusals():
    import scenes_np
    gw = W.World(scenes_np.disks_box(8, seed=7), lib=emul_lib())
    ids = np.array([0], np.uint32)
    assert gw.lib.pr_attach_shapes(gw.h, 1 << 20, 1, abi.ptr(ids)) == -2
    bad = np.array([1 << 20], np.uint32)
    assert gw.lib.pr_attach_shapes(gw.h, 1, 1, abi.ptr(bad)) == -2
    assert gw.lib.pr_attach_shapes(gw.h, 1, 0, abi.ptr(ids)) == 0


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_attach_equals_oracle():
    _run_device(device=0)
