"""pr_set_body_velocities = Body::SetVelocity (Body.cpp:184-194) for many bodies in one call (bulk body
access), and the batch entry points SURVEY 8(b) names (pr_batch_create_masked, pr_batch_stats). The device's rows right after
the call equal the real reference's after its own SetVelocity calls -- awake flag and under-active time included -- and the
world continues like the restatement given the same rows."""
import ctypes as C
import numpy as np
import pytest
import refworld
import scenes
import scenes_np
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import abi
from playrho_b200 import world as W


def _run(lib=None, device=-1):
    ref = scenes.SCENES["pyramid5"]()
    sc = ref.export_scene()
    orc = refworld.OracleWorld(sc)
    gw = W.World(sc, lib=lib, device=device, record_solve_order=True)
    gw.set_toi_mode(0)
    conf = refworld.default_step_conf()
    ref.lib.ref_set_body_velocity.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float]
    n = len(ref.bodies())
    for s in range(260):
        if s in (60, 250):  # 250: the pile has gone to sleep by then
            ids = np.array([0, 1, 2, 5, 7], np.uint32)                  # body 0 is the static ground: a non-zero velocity is ignored
            vel = np.array([[1.0, 0.0, 0.0], [0.0, 2.5, 0.0], [0.0, 0.0, 0.0], [-1.0, 3.0, 0.5], [0.0, 0.0, -2.0]], np.float32)
            g0 = gw.bodies().copy()
            gw.set_body_velocities(vel, ids)
            g1 = gw.bodies()
            # the same through the reference's own Body::SetVelocity, applied to the device's rows (the two worlds' trajectories are
            # their own: the reference solves in its native order)
            r = RefRow(ref, g0)
            for i, b in enumerate(ids):
                r.set_velocity(int(b), vel[i])
            assert_bits(g1, r.rows, "rows after the bulk SetVelocity at step %d" % s)
            orc.set_bodies(ids, g1[ids])
            if s == 250:
                assert (g0["flags"][1] & abi.BODY_AWAKE) == 0 and (g1["flags"][1] & abi.BODY_AWAKE) != 0   # woken
                assert (g1["flags"][2] & abi.BODY_AWAKE) == 0                                                # zero velocity: stays asleep
            assert g1["vel_x"][0] == 0.0                                                                    # the static body ignored it
        gs = gw.step(conf)
        os_ = orc.step(conf, order=gw.solve_order()[0])
        assert_bits(np.array([gs]), np.array([os_]), "step %d stats" % s)
        assert_bits(gw.bodies(), orc.bodies(), "step %d bodies" % s)
    assert n > 10


class RefRow:
    """Body::SetVelocity on rows, by the reference's rule (Body.cpp:184-194)."""

    def __init__(self, ref, rows):
        self.rows = rows.copy()

    def set_velocity(self, b, v):
        r = self.rows
        if np.any(v != 0):
            if not (r["flags"][b] & abi.BODY_VELOCITY):
                return
            r["flags"][b] |= abi.BODY_AWAKE
            r["under_active_time"][b] = 0.0
        r["vel_x"][b], r["vel_y"][b], r["vel_w"][b] = v


@need_ref
@need_oracle
@need_emul
def test_emulated_bulk_set_velocity():
    _run(lib=emul_lib())


@need_emul
def test_batch_masked_and_stats():
    lib = emul_lib()
    conf = abi.one(abi.WORLD_CONF)
    lib.pr_world_conf_default(abi.ptr(conf))
    h = C.c_void_p()
    assert lib.pr_batch_create_masked(abi.ptr(conf), 4, 0, C.byref(h)) == -1          # empty mask
    assert lib.pr_batch_create_masked(abi.ptr(conf), 4, 0b101, C.byref(h)) == 0       # lowest device of the mask
    st = abi.one(abi.STEP_STATS)
    sc = refworld.default_step_conf()
    out = abi.one(abi.STEP_STATS)
    assert lib.pr_batch_step(h, abi.ptr(sc), abi.ptr(out)) == 0
    assert lib.pr_batch_stats(h, abi.ptr(st)) == 0
    assert refworld.bits_equal(st, out)
    lib.pr_batch_destroy(h)


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_bulk_set_velocity():
    _run(device=0)
