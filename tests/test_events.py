"""Contact events (pr_events: begin / end / pre-solve / post-solve) and ShiftOrigin (pr_shift_origin).

Three links, like the rest of the suite:
  1. the REAL reference with its own listeners installed (oracle/ref_harness.cpp: SetBeginContactListener,
     SetEndContactListener, SetPostSolveContactListener; ShiftOrigin) == the restatement's recorded listener calls, in CALL
     ORDER, bit for bit (native constraint order);
  2. the device kernels compiled for the host (tests/_emul) == the restatement fed the device's colour-sorted order: the same
     multiset of events (the device appends from parallel threads, pr_events sorts by (kind, contact)), impulses bit for bit;
  3. the same on the GPU (-m gpu).
Reference: AabbTreeWorld.cpp:2249-2252, :2703-2714 (begin/end), :196-206, :1671-1674, :2053-2055 (post-solve), :2221-2243
(ShiftOrigin), DynamicTree.cpp:650-658."""
import numpy as np
import pytest
import refworld
import scenes
import scenes_np
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W

ALL = abi.EVENT_BEGIN | abi.EVENT_END | abi.EVENT_PRE_SOLVE | abi.EVENT_POST_SOLVE


def _sorted_contact_events(ev):
    return ev[np.lexsort((ev["contact"], ev["kind"]))]


def _sorted_presolve_events(ev):
    if not len(ev):
        return ev
    raw = np.ascontiguousarray(ev).view(np.uint32).reshape(len(ev), -1).astype(np.uint64)
    return ev[np.lexsort(tuple(raw[:, k] for k in range(raw.shape[1] - 1, -1, -1)))]


def _sorted_impulse_events(ev):
    ev = ev.copy()
    key = np.ascontiguousarray(ev["normal"]).view(np.uint32).astype(np.uint64)
    return ev[np.lexsort((key[:, 1], key[:, 0], ev["iterations"], ev["contact"], ev["phase"]))]


@need_ref
@need_oracle
@pytest.mark.parametrize("name,steps", [("hello", 60), ("pyramid5", 150), ("disks100", 200), ("mixed100", 200), ("shooter", 250)])
def test_restatement_listener_calls_equal_reference(name, steps):
    """Every begin/end/post-solve listener call of the real reference, in call order, default StepConf (TOI on)."""
    ref = scenes.SCENES[name]()
    orc = refworld.OracleWorld(ref.export_scene())
    ref.set_event_mask(ALL); orc.set_event_mask(ALL)
    conf = refworld.default_step_conf()
    seen = {abi.EVENT_BEGIN: 0, abi.EVENT_END: 0, abi.EVENT_POST_SOLVE: 0}
    toi_reports = 0
    for s in range(steps):
        ref.step(conf); orc.step(conf)
        for kind in (abi.EVENT_BEGIN, abi.EVENT_END, abi.EVENT_BEGIN | abi.EVENT_END):
            assert_bits(ref.events(kind), orc.events(kind), "%s step %d events kind %d" % (name, s, kind))
        ri, oi = ref.events(abi.EVENT_POST_SOLVE), orc.events(abi.EVENT_POST_SOLVE)
        toi_reports += int((oi["phase"] == 1).sum())
        oi = oi.copy(); oi["phase"] = 0xFFFFFFFF  # a listener cannot see the phase; the harness stores ~0
        assert_bits(ri, oi, "%s step %d post-solve" % (name, s))
        rp_, op_ = ref.events(abi.EVENT_PRE_SOLVE), orc.events(abi.EVENT_PRE_SOLVE)  # (contact, old manifold), call order
        assert_bits(rp_, op_, "%s step %d pre-solve" % (name, s))
        seen[abi.EVENT_PRE_SOLVE] = seen.get(abi.EVENT_PRE_SOLVE, 0) + len(rp_)
        seen[abi.EVENT_BEGIN] += len(ref.events(abi.EVENT_BEGIN))
        seen[abi.EVENT_END] += len(ref.events(abi.EVENT_END))
        seen[abi.EVENT_POST_SOLVE] += len(ri)
    assert seen[abi.EVENT_BEGIN] > 0 and seen[abi.EVENT_POST_SOLVE] > 0 and seen[abi.EVENT_PRE_SOLVE] > 0
    if name in ("disks100", "mixed100", "shooter"):
        assert seen[abi.EVENT_END] > 0
    if name == "shooter":
        assert toi_reports > 50


@need_ref
@need_oracle
@pytest.mark.parametrize("name", ["pyramid5", "mixed100", "shooter"])
def test_restatement_shift_origin_equals_reference(name):
    """ShiftOrigin between steps (twice, by amounts that round): bodies, proxies, contacts, manifolds stay bit-equal."""
    ref = scenes.SCENES[name]()
    orc = refworld.OracleWorld(ref.export_scene())
    conf = refworld.default_step_conf()
    for s in range(120):
        if s in (40, 80):
            dx, dy = (13.37, -2.25) if s == 40 else (-100.1, 7.3)
            ref.shift_origin(dx, dy); orc.shift_origin(dx, dy)
            assert_bits(ref.bodies(), orc.bodies(), "%s bodies right after the shift at %d" % (name, s))
            rp, op = ref.proxies(), orc.proxies()
            assert_bits(rp[np.argsort(rp["tree_id"])], op[np.argsort(op["tree_id"])], "%s proxies right after the shift" % name)
        rs, os_ = ref.step(conf), orc.step(conf)
        assert_bits(np.array([rs]), np.array([os_]), "%s step %d stats" % (name, s))
        assert_bits(ref.bodies(), orc.bodies(), "%s step %d bodies" % (name, s))
        rc, rm = ref.contacts(); oc, om = orc.contacts()
        assert_bits(rc, oc, "%s step %d contacts" % (name, s))
        assert_bits(rm, om, "%s step %d manifolds" % (name, s))
        rp, op = ref.proxies(), orc.proxies()
        assert_bits(rp[np.argsort(rp["tree_id"])], op[np.argsort(op["tree_id"])], "%s step %d proxies" % (name, s))


def run_events_vs_oracle(sc, steps, lib=None, device=-1, shift_at=()):
    conf = refworld.default_step_conf()
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    gw.set_event_mask(ALL, 1 << 16)
    orc = refworld.OracleWorld(sc)
    orc.set_event_mask(ALL)
    n = {abi.EVENT_BEGIN: 0, abi.EVENT_END: 0, abi.EVENT_POST_SOLVE: 0, "toi": 0}
    for s in range(steps):
        if s in shift_at:
            gw.shift_origin(*shift_at[s]); orc.shift_origin(*shift_at[s])
            assert_bits(gw.bodies(), orc.bodies(), "bodies right after the shift at %d" % s)
            assert_bits(gw.proxies(), orc.proxies(), "proxies right after the shift at %d" % s)
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        gsa, osa = np.array([gs]), np.array([os_])
        assert_bits(gsa, osa, "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
        gc, gm = gw.contacts(); oc, om = orc.contacts()
        assert_bits(gc, oc, "step %d contacts" % s)
        assert_bits(gm, om, "step %d manifolds" % s)
        assert_bits(gw.proxies(), orc.proxies(), "step %d proxies" % s)
        ge = gw.events(abi.EVENT_BEGIN | abi.EVENT_END)
        oe = _sorted_contact_events(orc.events(abi.EVENT_BEGIN | abi.EVENT_END))
        assert_bits(ge, oe, "step %d begin/end events" % s)
        assert_bits(gw.events(abi.EVENT_BEGIN), oe[oe["kind"] == abi.EVENT_BEGIN], "step %d begin events" % s)
        gi = _sorted_impulse_events(gw.events(abi.EVENT_POST_SOLVE))
        oi = _sorted_impulse_events(orc.events(abi.EVENT_POST_SOLVE))
        assert_bits(gi, oi, "step %d post-solve events" % s)
        gp_ = _sorted_presolve_events(gw.events(abi.EVENT_PRE_SOLVE))
        op_ = _sorted_presolve_events(orc.events(abi.EVENT_PRE_SOLVE))
        assert_bits(gp_, op_, "step %d pre-solve events" % s)
        n["pre"] = n.get("pre", 0) + len(op_)
        n[abi.EVENT_BEGIN] += int((oe["kind"] == abi.EVENT_BEGIN).sum()); n[abi.EVENT_END] += int((oe["kind"] == abi.EVENT_END).sum())
        n[abi.EVENT_POST_SOLVE] += len(oi); n["toi"] += int((oi["phase"] == 1).sum())
    return n


@need_oracle
@need_emul
@pytest.mark.parametrize("name,builder,steps", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150),
])
def test_emulated_events_equal_oracle(name, builder, steps):
    n = run_events_vs_oracle(builder(), steps, lib=emul_lib())
    assert n[abi.EVENT_BEGIN] > 0 and n[abi.EVENT_END] > 0 and n[abi.EVENT_POST_SOLVE] > 0 and n["pre"] > 0


@need_oracle
@need_emul
@need_ref
def test_emulated_events_shooter_toi_phase():
    """Bullets and a kinematic bar: the TOI phase reports its own post-solve calls and begin events."""
    n = run_events_vs_oracle(scenes.shooter().export_scene(), 200, lib=emul_lib())
    assert n["toi"] > 50 and n[abi.EVENT_END] > 0


@need_oracle
@need_emul
def test_emulated_shift_origin_equals_oracle():
    sc = scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0)
    run_events_vs_oracle(sc, 120, lib=emul_lib(), shift_at={40: (13.37, -2.25), 80: (-100.1, 7.3)})


@need_emul
def test_event_mask_errors_and_overflow():
    sc = scenes_np.disks_box(64, seed=7)
    gw = W.World(sc, lib=emul_lib())
    assert gw.lib.pr_set_event_mask(gw.h, 64, 16) == -1                     # PR_ERR_INVALID_ARGUMENT
    conf = refworld.default_step_conf()
    gw.step(conf)
    assert len(gw.events(abi.EVENT_POST_SOLVE)) == 0                        # nothing recorded while the mask is 0
    gw.set_event_mask(abi.EVENT_POST_SOLVE, 4)
    import ctypes as C
    rc = 0
    for _ in range(60):
        gw.step(conf)
        out = np.zeros(4, abi.IMPULSE_EVENT); k = C.c_uint32(0)
        rc = gw.lib.pr_events(gw.h, abi.EVENT_POST_SOLVE, abi.ptr(out), 4, C.byref(k))
        if rc != 0:
            break
    assert rc == -4 and k.value > 4                                         # PR_ERR_LENGTH: the step dropped events, count says how many


@pytest.mark.gpu
@need_oracle
@pytest.mark.parametrize("name,builder,steps", [
    ("disks64", lambda: scenes_np.disks_box(64, seed=7), 120),
    ("mixed80", lambda: scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0), 150),
])
def test_gpu_events_equal_oracle(name, builder, steps):
    n = run_events_vs_oracle(builder(), steps, device=0)
    assert n[abi.EVENT_BEGIN] > 0 and n[abi.EVENT_END] > 0 and n[abi.EVENT_POST_SOLVE] > 0


@pytest.mark.gpu
@need_oracle
def test_gpu_shift_origin_equals_oracle():
    sc = scenes_np.mixed(80, seed=11, tile_segments=16, tile_width=8.0)
    run_events_vs_oracle(sc, 120, device=0, shift_at={40: (13.37, -2.25), 80: (-100.1, 7.3)})
