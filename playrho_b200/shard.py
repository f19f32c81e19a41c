"""Sharding of batches of independent worlds over the GPUs of one box (BASELINE config 5).

Worlds share nothing (World.hpp:652-656), so the only thing that ever crosses NVLink is the per-step
statistics record: every rank steps its own contiguous block of worlds in its own device arena
(`world.Batch`), and `reduce_step_stats` folds the ranks' `StepStats` the way the reference's
`Update(RegStepStats&, IslandStats)` (StepStats.cpp:29-38) folds islands: counters add, `min_separation`
takes the min, `max_inc_impulse` / iteration maxima take the max. The transport is `torch.distributed`
(NCCL on the GPUs, gloo in the CPU tests); there is no data-path collective.
"""
import numpy as np

from . import abi

# StepStats fields that are maxima / minima over islands (everything else is a counter and adds)
_MAX_FIELDS = {("reg", "max_inc_impulse"), ("toi", "max_inc_impulse"), ("reg", "max_island_bodies"),
               ("toi", "max_dist_iters"), ("toi", "max_toi_iters"), ("toi", "max_root_iters")}
_MIN_FIELDS = {("reg", "min_separation"), ("toi", "min_separation")}


def world_block(n_worlds, rank, world_size):
    """Contiguous block of worlds owned by `rank`: (first world, count). Blocks differ by at most one world."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world_size %d" % (rank, world_size))
    base, extra = divmod(n_worlds, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def world_owner(world, n_worlds, world_size):
    """Inverse of world_block: the rank that owns global world index `world`."""
    base, extra = divmod(n_worlds, world_size)
    edge = extra * (base + 1)
    if world < edge:
        return world // (base + 1)
    return extra + (world - edge) // base


def _stat_fields():
    for phase in abi.STEP_STATS.names:
        for name in abi.STEP_STATS[phase].names:
            yield phase, name


def stats_to_vectors(st):
    """Split one STEP_STATS record into (sum vector, max vector, min vector) of float64 (u32 counters are exact)."""
    s, mx, mn = [], [], []
    for phase, name in _stat_fields():
        v = float(st[phase][name])
        if (phase, name) in _MAX_FIELDS:
            mx.append(v)
        elif (phase, name) in _MIN_FIELDS:
            mn.append(v)
        else:
            s.append(v)
    return np.array(s, np.float64), np.array(mx, np.float64), np.array(mn, np.float64)


def vectors_to_stats(s, mx, mn):
    out = abi.one(abi.STEP_STATS)
    si = mi = ni = 0
    for phase, name in _stat_fields():
        if (phase, name) in _MAX_FIELDS:
            v, mi = mx[mi], mi + 1
        elif (phase, name) in _MIN_FIELDS:
            v, ni = mn[ni], ni + 1
        else:
            v, si = s[si], si + 1
        out[phase][name] = v
    return out[0]


def fold_step_stats(records):
    """Host-side fold of several ranks' (or several steps') STEP_STATS records; the collective must agree with it."""
    parts = [stats_to_vectors(r) for r in records]
    return vectors_to_stats(np.sum([p[0] for p in parts], axis=0), np.max([p[1] for p in parts], axis=0),
                            np.min([p[2] for p in parts], axis=0))


def reduce_step_stats(st, body_steps, seconds, device=None):
    """All-reduce one rank's step statistics over the process group (3 small collectives, ~300 B each).

    Returns (folded STEP_STATS record, total body-steps, max seconds over ranks) on every rank. With no
    initialised process group (single GPU) it returns the inputs unchanged."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return st, float(body_steps), float(seconds)
    s, mx, mn = stats_to_vectors(st)
    dev = device if device is not None else "cpu"
    ts = torch.tensor(np.concatenate([s, [float(body_steps)]]), dtype=torch.float64, device=dev)
    tmax = torch.tensor(np.concatenate([mx, [float(seconds)]]), dtype=torch.float64, device=dev)
    tmin = torch.tensor(mn, dtype=torch.float64, device=dev)
    dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    ts, tmax, tmin = ts.cpu().numpy(), tmax.cpu().numpy(), tmin.cpu().numpy()
    return vectors_to_stats(ts[:-1], tmax[:-1], tmin), float(ts[-1]), float(tmax[-1])
