"""Two world handles used from two host threads at once (include/playrho_b200.h allows one thread per handle), on one device
and -- when the box has two -- on two devices: every entry point runs under the world's own device (DeviceGuard in
pr_world.cu), the launch serial that lets read_counters skip a round trip is atomic, the stage timer is per world."""
import threading
import numpy as np
import pytest
import refworld
import scenes_np
from common import assert_bits
from playrho_b200 import world as W

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


def _run(scene, device, steps, out, key):
    gw = W.World(scene, device=device)
    conf = W.default_step_conf()
    stats = []
    for _ in range(steps):
        stats.append(gw.step(conf))
    out[key] = (gw.bodies(), gw.contacts(), np.array(stats))


@pytest.mark.parametrize("two_devices", [False, True])
def test_two_worlds_two_threads(two_devices):
    if two_devices and _device_count() < 2:
        pytest.skip("needs two GPUs")
    a = scenes_np.mixed(400, seed=3, tile_segments=32, tile_width=16.0)
    b = scenes_np.disks_box(2000, seed=5)
    serial, threaded = {}, {}
    _run(a, 0, 80, serial, "a")
    _run(b, 1 if two_devices else 0, 80, serial, "b")
    ts = [threading.Thread(target=_run, args=(a, 0, 80, threaded, "a")),
          threading.Thread(target=_run, args=(b, 1 if two_devices else 0, 80, threaded, "b"))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for k in ("a", "b"):
        assert k in threaded, "thread %s died" % k
        assert_bits(serial[k][0], threaded[k][0], k + " bodies")
        assert_bits(serial[k][1][0], threaded[k][1][0], k + " contacts")
        assert_bits(serial[k][1][1], threaded[k][1][1], k + " manifolds")
        assert_bits(serial[k][2], threaded[k][2], k + " stats")
