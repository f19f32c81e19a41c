"""N>1 host path: worlds sharded over ranks, step statistics folded by a collective.

World size 2 over gloo on the CPU box; each rank steps ITS block of tumbler worlds in its own arena through
the C ABI (the host-emulated build of the same kernels -- tests only), then `shard.reduce_step_stats`
all-reduces the StepStats. Checked: (1) the blocks partition the worlds, (2) the collective equals the
host-side fold of the two ranks' records field by field (sums / min / max, like StepStats.cpp:29-38),
(3) a shard is bit-identical to the same worlds stepped alone in one process (no cross-rank state),
(4) body-steps and the max-over-ranks time come back on every rank."""
import os
import socket
import sys

import numpy as np
import pytest

import refworld
import scenes_np
from common import need_emul, EMUL_SO
from playrho_b200 import abi, shard, world as W

N_WORLDS, N_SQUARES, STEPS = 5, 12, 40


def _step_block(first, count):
    lib = W.load_library(EMUL_SO)
    units = [scenes_np.tumbler(N_SQUARES, seed=12345 + first + w) for w in range(count)]
    plain, world = scenes_np.merge_worlds(units)
    conf = scenes_np.tumbler_step_conf(refworld.default_step_conf())
    gw = W.Batch(count, lib=lib)
    gw.create_shapes(plain)
    gw.create_bodies(plain["bodies"], world, plain["shape_offsets"], plain["shape_ids"])
    gw.create_joints(plain["joints"])
    recs = [gw.step(conf).copy() for _ in range(STEPS)]
    return gw.bodies(), recs


def _worker(rank, world_size, port, outdir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(refworld.ROOT, "tests"))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world_size)
    first, count = shard.world_block(N_WORLDS, rank, world_size)
    rows, recs = _step_block(first, count)
    local = shard.fold_step_stats(recs)
    folded, body_steps, secs = shard.reduce_step_stats(local, len(rows) * STEPS, 0.25 * (rank + 1))
    np.save(os.path.join(outdir, "rows%d.npy" % rank), rows.view(np.uint8))
    np.save(os.path.join(outdir, "local%d.npy" % rank), np.array([local]).view(np.uint8))
    np.save(os.path.join(outdir, "folded%d.npy" % rank), np.array([folded]).view(np.uint8))
    np.save(os.path.join(outdir, "scalars%d.npy" % rank), np.array([body_steps, secs, first, count]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_blocks_partition_the_worlds():
    for n, g in [(8192, 1), (8192, 8), (5, 2), (7, 4), (3, 8), (0, 2)]:
        blocks = [shard.world_block(n, r, g) for r in range(g)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
            assert f1 == f0 + c0
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
        for r, (f, c) in enumerate(blocks):
            for w in range(f, f + c):
                assert shard.world_owner(w, n, g) == r
    with pytest.raises(ValueError):
        shard.world_block(4, 2, 2)


@need_emul
def test_two_rank_gloo_shards_and_folds_stats(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    locs, folds, scal, rows = [], [], [], []
    for r in range(2):
        locs.append(np.load(tmp_path / ("local%d.npy" % r)).view(abi.STEP_STATS)[0])
        folds.append(np.load(tmp_path / ("folded%d.npy" % r)).view(abi.STEP_STATS)[0])
        scal.append(np.load(tmp_path / ("scalars%d.npy" % r)))
        rows.append(np.load(tmp_path / ("rows%d.npy" % r)).view(abi.BODY))
    # (2) collective == host fold, identical on both ranks
    want = shard.fold_step_stats(locs)
    for f in folds:
        assert np.array([f]).tobytes() == np.array([want]).tobytes()
    assert want["reg"]["islands_found"] == locs[0]["reg"]["islands_found"] + locs[1]["reg"]["islands_found"]
    assert want["reg"]["min_separation"] == min(locs[0]["reg"]["min_separation"], locs[1]["reg"]["min_separation"])
    assert want["reg"]["max_inc_impulse"] == max(locs[0]["reg"]["max_inc_impulse"], locs[1]["reg"]["max_inc_impulse"])
    assert want["reg"]["islands_found"] >= N_WORLDS * STEPS  # every world has at least its enclosure's island
    # (4) body-steps add, time is the max over ranks
    per = 2 + N_SQUARES
    for sc in scal:
        assert sc[0] == N_WORLDS * per * STEPS and sc[1] == 0.5
    assert [int(sc[2]) for sc in scal] == [0, 3] and [int(sc[3]) for sc in scal] == [3, 2]
    # (3) a shard owes nothing to the other rank: the same block stepped alone is bit-identical
    for r in range(2):
        first, count = shard.world_block(N_WORLDS, r, 2)
        alone, _ = _step_block(first, count)
        assert alone.tobytes() == rows[r].tobytes()
        assert len(alone) == count * per
