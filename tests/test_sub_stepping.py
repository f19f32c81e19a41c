"""Sub-stepping (SetSubStepping / IsStepComplete, AabbTreeWorld.cpp:1779-1824): with it on, a Step solves at most ONE TOI event
and leaves the step incomplete; the following Step calls continue the TOI phase. The device world runs its serial-exact event
order while sub-stepping. Checked against the REAL reference directly: the scene of the reference's own World.IsStepComplete
test (two separate disks landing on an edge at the same instant: one-body islands, so the colour-sorted order is the native
order) -- body rows and the IsStepComplete sequence, bit for bit."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_emul, emul_lib, assert_bits
from playrho_b200 import world as W


def _run(lib=None, device=-1):
    ref = RefWorld()
    disk = ref.add_disk(1.0, density=1.0)
    ref.add_body(DYNAMIC, pos=(-2.0, 2.0), acc=(0.0, -9.8), shapes=[disk])
    ref.add_body(DYNAMIC, pos=(2.0, 2.0), acc=(0.0, -9.8), shapes=[disk])
    edge = ref.add_edge((-10.0, 0.0), (10.0, 0.0))
    ref.add_body(STATIC, shapes=[edge])
    sc = ref.export_scene()
    gw = W.World(sc, lib=lib, device=device)   # default TOI mode (parallel): sub-stepping switches to the serial-exact order
    ref.lib.ref_set_sub_stepping(ref.h, 1)
    gw._chk(gw.lib.pr_set_sub_stepping(gw.h, 1))
    conf = refworld.default_step_conf(delta_time=1.0 / 100.0)
    incomplete = 0
    for s in range(160):
        ref.step(conf); gw.step(conf)
        rc, gc = ref.lib.ref_is_step_complete(ref.h), gw.lib.pr_is_step_complete(gw.h)
        assert rc == gc, "step %d: IsStepComplete %d (reference) vs %d" % (s, rc, gc)
        incomplete += 1 - rc
        assert_bits(ref.bodies(), gw.bodies(), "step %d bodies" % s)
    assert incomplete >= 2   # the two simultaneous landings took two extra Step calls


@need_ref
@need_emul
def test_emulated_sub_stepping_equals_reference():
    _run(lib=emul_lib())


@pytest.mark.gpu
@need_ref
def test_gpu_sub_stepping_equals_reference():
    _run(device=0)
