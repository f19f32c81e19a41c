"""DestroyContacts (AabbTreeWorld.cpp:2290-2331) is TWO passes over the contact list: the contacts that stopped overlapping go first, then the
ones flagged for filtering that may not collide any more. Their ids are freed in that order, so the ids later contacts get tell a one-pass
implementation apart: here contact 0 (earlier in the list) is refused by the filter pass -- a joint that forbids the collision was created
between its bodies -- in the same Step in which contact 1 stops overlapping -- its body was moved away; the next contact created must get id 0
(freed last), not 1. Found by tests/test_fuzz_mutators.py; reference == restatement == emulated kernels == GPU."""
import numpy as np
import pytest
import refworld
from refworld import RefWorld, STATIC, DYNAMIC
from common import need_ref, need_oracle, need_emul, emul_lib, assert_bits
from playrho_b200 import world as W


def _run(lib=None, device=-1, with_device=True):
    ref = RefWorld()
    box = ref.add_box(0.5, 0.5, density=1.0)
    b = [ref.add_body(DYNAMIC, pos=p, shapes=[box], allow_sleep=False) for p in ((0.0, 0.0), (0.8, 0.0), (5.0, 0.0), (5.8, 0.0), (10.0, 0.0))]
    sc = ref.export_scene()
    others = [refworld.OracleWorld(sc)]
    if with_device:
        gw = W.World(sc, lib=lib, device=device)
        others.append(gw)
    conf = refworld.default_step_conf()

    def step(n=1):
        for _ in range(n):
            ref.step(conf)
            for o in others:
                o.step(conf)

    def ids():
        c, _ = ref.contacts()
        out = [(int(i), int(a), int(bb)) for i, a, bb in zip(c["id"], c["body_a"], c["body_b"])]
        for o in others:
            oc, _ = o.contacts()
            assert_bits(c, oc, "contacts")
        return out

    def move(body, x, y):
        ref.lib.ref_set_body_transform(ref.h, body, x, y, 0.0)
        row = ref.bodies()[body:body + 1]
        for o in others:
            o.set_bodies([body], row)

    step(2)
    assert ids() == [(0, b[0], b[1]), (1, b[2], b[3])]
    jid = ref.add_revolute(b[0], b[1], (0.4, 0.0))          # collideConnected false: contact 0 is flagged for filtering
    row = ref.joints()[[jid]]
    for o in others:
        o.create_joints(row)
    move(b[3], 20.0, 0.0)                                   # contact 1 stops overlapping
    step(1)
    assert ids() == []                                      # both gone in the same DestroyContacts
    move(b[4], 5.8, 0.0)                                    # a new pair: b2 - b4
    step(1)
    got = ids()
    assert got == [(0, b[2], b[4])], got                    # id 0 was freed last (second pass): it comes back first


@need_ref
@need_oracle
def test_restatement_frees_filtered_contacts_after_the_overlap_pass():
    _run(with_device=False)


@need_ref
@need_oracle
@need_emul
def test_emulated_kernels_free_filtered_contacts_after_the_overlap_pass():
    _run(lib=emul_lib())


@pytest.mark.gpu
@need_ref
@need_oracle
def test_gpu_frees_filtered_contacts_after_the_overlap_pass():
    _run(device=0)
