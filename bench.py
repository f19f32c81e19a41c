#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric for the B200-native PlayRho world.

  python bench.py --gpus N --steps K --warmup W [--bodies B] [--workload mixed|disks|tumblers]
  python bench.py --impl reference ...      # the reference's own CPU Step on the host cores

One "step" = one Step(world, StepConf) of the hot path (AabbTreeWorld.cpp:2144) over the synthetic
scene named in config.workload. Default workload: BASELINE config 4 shape (1M mixed polygons+disks
over tiled chain terrain), one such world per GPU (N>1: independent worlds, weak scaling, no data-path
collective; NCCL only reduces step statistics). --workload tumblers: BASELINE config 5, --worlds
(default 8192) independent 514-body tumbler worlds in one device arena per GPU, the worlds sharded
across the N ranks (strong scaling). Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _checker_path():
    """tests/ holds the reference-side harness (refworld, scenes: oracle/_ref bindings). Only the reference arm and the
    cpu_baseline leg put it on the path; the product arm imports nothing from tests/ or oracle/."""
    t = os.path.join(ROOT, "tests")
    if t not in sys.path:
        sys.path.insert(0, t)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--bodies", type=int, default=1000000)
    ap.add_argument("--workload", default=None, help="mixed (BASELINE config 4: the N=1 default) | disks | tumblers (BASELINE config 5, "
                    "--worlds independent worlds sharded over the GPUs: the default for --gpus > 1, the split north_star names)")
    ap.add_argument("--settle", type=int, default=300, help="untimed steps so that the dropped bodies have piled up and are in contact")
    ap.add_argument("--worlds", type=int, default=8192, help="tumblers: total number of independent worlds (sharded over the ranks)")
    ap.add_argument("--squares", type=int, default=512, help="tumblers: squares added one per step per world before timing")
    ap.add_argument("--ref-worlds", type=int, default=0, help="tumblers reference arm: worlds stepped on the host (default: one per host core)")
    ap.add_argument("--ref-bodies", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-toi", action="store_true", help="StepConf.doToi=false (default: TOI enabled as BASELINE config 4 says)")
    ap.add_argument("--toi-mode", type=int, default=1, help="1 = parallel independent TOI events (product default), 0 = serial-exact")
    ap.add_argument("--debug-flags", type=int, default=0, help="pr_set_debug flags (see include/playrho_b200.h; 32 = tumbler arena through the arena-wide per-colour launches instead of the block solver)")
    ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    ap.add_argument("--strip", type=float, default=None, help="mixed: width (m) of the terrain strip the bodies are dropped over (default: a 4:1 lattice, 2 200 m at 1M bodies)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra measurements of the N=1 mixed line (tumbler arena at N=1, serial-exact TOI mode)")
    args = ap.parse_args()
    if args.workload is None:  # one world runs on one GPU (config 4); what shards over the GPUs of a box is the batch of worlds (config 5)
        args.workload = "mixed" if args.gpus <= 1 else "tumblers"
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region. In-process NVML (nvidia_ml_py) every 20 ms; spawning
    `nvidia-smi` instead (the fallback) takes a driver lock for tens of ms and showed up as +0.7 ms/step on the
    host-interactive TOI phase in about one run out of three."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.how = index, [], False, "nvml"
        try:  # NVML is initialised HERE, long before the timed region: the first nvmlInit of a box stalls the driver for ~0.5 s
            import pynvml
            pynvml.nvmlInit()
            self.nvml_handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml_max = pynvml.nvmlDeviceGetMaxClockInfo(self.nvml_handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(self.nvml_handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.nvml_handle)
        except Exception:
            self.nvml_handle = None

    def _nvml_loop(self):
        import pynvml
        if self.nvml_handle is None:
            raise RuntimeError("no NVML")
        h, mx = self.nvml_handle, self.nvml_max
        bits = [("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap)]
        while not self.stop_flag:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.samples.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
            time.sleep(0.02)

    def _smi_loop(self):
        self.how = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def mark(self):
        """Start of the timed region: forget the samples taken while warming up."""
        self.samples = []

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples), "via": self.how}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE regular-phase k_update_contacts launch on this workload, from the
# `ncu --set full` capture summarised under profiles/ (None until a capture of the current kernel exists)
ROOFLINE_TRAFFIC = {"mixed": int(1.27e9 + 0.2387e9)}  # profiles/r02k_ncu_full_mixed1m.txt, k_update_contacts (590 us row; round 2 final build)


def algorithmic_bytes(prof, nv, np_):
    """SURVEY.md §8(d) B_step from this step's device counters (TOI term omitted while TOI is off)."""
    n_awake, p, pm = int(prof["n_awake"]), int(prof["p_total"]), int(prof["p_moved"])
    c, cd, ct = int(prof["c_total"]), int(prof["c_dirty"]), int(prof["c_touching"])
    vbar = 3.5
    logp = int(np.ceil(np.log2(max(2, p))))
    terms = {
        "a1_a12_bodies": 128 * n_awake,
        "a2_proxies": int((36 + 8 * vbar) * n_awake + 16 * pm),
        "a3_bvh": 184 * p if pm else 0,
        "a4_query": pm * (32 * logp + 16 * 6),
        "a6_destroy": 32 * c,
        "a7_narrow": int((184 + 16 * 2 * vbar) * cd),
        "a8_islands_colour": 24 * (ct + n_awake),
        "a9_setup": 360 * ct,
        "a10_velocity": 188 * ct * nv,
        "a11_position": 112 * ct * np_,
        "a12_impulses": 16 * ct,
    }
    return sum(terms.values()), terms


def build_scene(args, seed):
    from playrho_b200 import scenes_np
    if args.workload == "disks":
        return scenes_np.disks_box(args.bodies, seed=seed)
    return scenes_np.mixed_fast(args.bodies, seed=seed, strip_width=args.strip)


# ---------------------------------------------------------------- reference arm (CPU)
def ref_scene(args, n):
    _checker_path()
    import refworld
    import scenes
    ids32 = refworld.have_ref(True)
    return scenes.mixed(n, ids32=ids32) if args.workload != "disks" else scenes.disks_box(n, ids32=ids32)


def ref_tumbler_world(n_squares, conf):
    """One TumblerPlayRho (Benchmark/BenchmarkMain.cpp:2614-2722) in the real reference: Step(); AddSquare() x n."""
    _checker_path()
    import refworld
    import scenes
    w = scenes.tumbler(0)
    sq = w.add_box(0.125, 0.125, density=1.0)
    for _ in range(n_squares):
        w.step(conf)
        w.add_body(refworld.DYNAMIC, pos=(0.0, 10.0), acc=scenes.GRAVITY, shapes=[sq])
    return w


def time_reference_tumblers(args, steps, warmup):
    """Reference Step on `cores` host threads, one independent tumbler world each (the reference is single-threaded
    per world; ctypes releases the GIL inside ref_step_timed). Returns (body-steps/s, cores, bodies per world)."""
    import refworld
    from playrho_b200 import scenes_np
    cores = args.ref_worlds or max(1, len(os.sched_getaffinity(0)))
    conf = scenes_np.tumbler_step_conf(refworld.default_step_conf(do_toi=0 if args.no_toi else 1))
    worlds = [None] * cores

    def build(i):
        worlds[i] = ref_tumbler_world(args.squares, conf)
        worlds[i].step_timed(conf, warmup)
    th = [threading.Thread(target=build, args=(i,)) for i in range(cores)]
    [t.start() for t in th]
    [t.join() for t in th]
    secs = [0.0] * cores

    def run(i):
        secs[i] = worlds[i].step_timed(conf, steps)[0]
    th = [threading.Thread(target=run, args=(i,)) for i in range(cores)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    nb = int(worlds[0].counts()["bodies"])
    return cores * nb * steps / wall, cores, nb, wall


def run_reference(args, rank, world_size):
    """--impl reference: the reference's own Step (oracle/_ref, built from /root/reference) on the host."""
    if rank != 0:
        return
    _checker_path()
    import refworld
    from playrho_b200 import abi  # noqa: F401
    if not refworld.have_ref(refworld.have_ref(True)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built"}))
        return
    if args.workload == "tumblers":
        value, cores, nb, wall = time_reference_tumblers(args, args.steps, args.warmup)
        ms = 1e3 * wall / args.steps
        workload = "tumblers: %d independent %d-body tumbler worlds, one per host thread (bounded sample of the %d-world workload)" % (cores, nb, args.worlds)
        sample = "%d worlds x %d bodies, %d steps after %d add-square steps, %d threads" % (cores, nb, args.steps, args.squares, cores)
    else:
        n = args.ref_bodies
        ref = ref_scene(args, n)
        conf = refworld.default_step_conf(do_toi=0 if args.no_toi else 1)
        ref.step_timed(conf, args.settle)
        ref.step_timed(conf, args.warmup)
        secs, _ = ref.step_timed(conf, args.steps)
        nb = int(ref.counts()["bodies"])
        value, cores, ms = nb * args.steps / secs, 1, 1e3 * secs / args.steps
        workload = "%s-%d (bounded sample of the %d-body workload; reference Step is O(islands x bodies), SURVEY F4)" % (args.workload, nb, args.bodies)
        sample = "%d bodies, %d steps after %d settle steps, 1 thread (reference is single-threaded)" % (nb, args.steps, args.settle)
    line = {
        "impl": "reference", "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.workload == "tumblers" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "toi": not args.no_toi, "step_conf": "default StepConf (8 vel / 3 pos iters)"},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------- device worlds
def make_mixed_world(args, W, rank, local_rank):
    sc = build_scene(args, seed=12345 + rank)
    nb = len(sc["bodies"])
    gw = W.World(sc, device=local_rank, body_capacity=nb + 16, proxy_capacity=nb + 8192, contact_capacity=6 * nb)
    conf = W.default_step_conf(do_toi=0 if args.no_toi else 1)
    gw.set_toi_mode(args.toi_mode)
    if args.debug_flags:
        gw.lib.pr_set_debug(gw.h, args.debug_flags)
    for _ in range(args.settle):
        gw.step(conf)
    desc = "%s-%d bodies per GPU (BASELINE config 4 shape: 50%% disks / 50%% 4-8-gons over tiled 256-segment chain terrain), " \
           "state after %d settle steps (bodies piled up and in contact)" % (args.workload, nb, args.settle)
    return gw, conf, desc, {"settle_steps": args.settle}


def make_tumbler_arena(args, W, rank, world_size, local_rank):
    """BASELINE config 5: this rank's share of --worlds tumbler worlds in one device arena; Step(); AddSquare() x --squares."""
    from playrho_b200 import scenes_np
    from playrho_b200 import shard
    _, n_worlds = shard.world_block(args.worlds, rank, world_size)  # this rank's contiguous block of worlds
    unit = scenes_np.tumbler(0)
    per = 2 + args.squares
    gw = W.Batch(n_worlds, device=local_rank, body_capacity=n_worlds * per + 16, proxy_capacity=n_worlds * (4 + args.squares) + 64,
                 contact_capacity=12 * n_worlds * args.squares)  # (~9 fat-box pairs per square at 512 squares: no growth while timing)
    gw.set_toi_mode(args.toi_mode)
    if args.debug_flags:
        gw.lib.pr_set_debug(gw.h, args.debug_flags)
    gw.create_shapes(unit)  # the 5 shapes (4 enclosure walls + the square) are shared by every world
    nu = len(unit["bodies"])
    offs = np.concatenate([[0], (unit["shape_offsets"][1:][None, :] + int(unit["shape_offsets"][-1]) * np.arange(n_worlds)[:, None]).reshape(-1)])
    n_ids = int(unit["shape_offsets"][-1])
    gw.create_bodies(np.tile(unit["bodies"], n_worlds), np.repeat(np.arange(n_worlds), nu), offs, np.tile(unit["shape_ids"][:n_ids], n_worlds))
    j = np.tile(unit["joints"], n_worlds)
    j["body_a"] += (nu * np.arange(n_worlds)).astype(np.uint32)
    j["body_b"] += (nu * np.arange(n_worlds)).astype(np.uint32)
    gw.create_joints(j)
    conf = scenes_np.tumbler_step_conf(W.default_step_conf(do_toi=0 if args.no_toi else 1))
    row = scenes_np.tumbler_square_row(unit)
    rows, wi = np.repeat(row, n_worlds), np.arange(n_worlds, dtype=np.uint32)
    so, si = np.arange(n_worlds + 1, dtype=np.uint32), np.full(n_worlds, unit["square_shape"], np.uint32)
    for _ in range(args.squares):
        gw.step(conf)
        gw.create_bodies(rows, wi, so, si)
    desc = "tumblers: %d independent worlds (%d on this GPU, one device arena) x %d bodies (BASELINE config 5: enclosure on a " \
           "motorised revolute joint + %d squares added one per step before timing, Benchmark/BenchmarkMain.cpp:2614-2759)" \
           % (args.worlds, n_worlds, per, args.squares)
    return gw, conf, desc, {"worlds": args.worlds, "worlds_per_gpu": n_worlds, "bodies_per_world": per}


def time_steps(gw, conf, stream, steps, warmup=3):
    """Device time of `steps` steps of an already-built world (CUDA events on the world's stream)."""
    import torch
    for _ in range(warmup):
        gw.step(conf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        gw.step(conf)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world_size)
        return
    import torch
    import torch.distributed as dist
    from playrho_b200 import abi, world as W
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tumblers = args.workload == "tumblers"
    sampler = ClockSampler(local_rank)  # (initialises NVML now; sampling starts with the timed region)
    if tumblers:
        gw, conf, desc, extra = make_tumbler_arena(args, W, rank, world_size, local_rank)
    else:
        gw, conf, desc, extra = make_mixed_world(args, W, rank, local_rank)
    nb = int(gw.counts()["bodies"])
    warmup = max(3, args.warmup)
    # the sampling thread runs through the warm-up steps already: the first NVML queries of a thread on a fresh box were seen to
    # stall the driver for ~1 s (one 1.29 s step in the first process of a box); its samples before the timed region are dropped
    sampler.start()
    stream = torch.cuda.ExternalStream(gw.stream(), device=torch.device("cuda", local_rank))
    # everything the timed loop touches is used once while warming up (torch's lazy CUDA initialisation, event pools, profile reads)
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record(stream)
    for _ in range(warmup):
        gw.step(conf)
        gw.profile()
    w1.record(stream)
    torch.cuda.synchronize()
    w0.elapsed_time(w1)
    time.sleep(0.1)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler.mark()
    # ---------------- device-resident timed region: K steps, state already in HBM; CUDA events on the world's stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.cuda_profiler:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    ev0.record(stream)
    launches, stage_ms, profs, step_recs = 0, np.zeros(abi.N_STAGES), [], []
    stats_sum = np.zeros(4, dtype=np.float64)
    for _ in range(args.steps):
        st = gw.step(conf)
        p = gw.profile()
        launches += int(p["kernel_launches"])
        stage_ms += p["ms_stage"]
        profs.append(p.copy())
        step_recs.append(st.copy())
        stats_sum += [st["reg"]["islands_found"], st["reg"]["contacts_added"], st["reg"]["proxies_moved"], st["reg"]["bodies_slept"]]
    ev1.record(stream)
    barrier()
    wall_secs = time.perf_counter() - t0
    secs = ev0.elapsed_time(ev1) / 1e3
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    if rank == 0:  # the slowest steps with their stage times (stderr): growth / rebuild steps show up here
        for k in sorted(range(len(profs)), key=lambda k: -float(profs[k]["ms_total"]))[:3]:
            print("slow step %d: %.2f ms colour rounds %d, buffer growths %d  %s" % (k, float(profs[k]["ms_total"]), int(profs[k]["color_rounds"]), int(profs[k]["grow_events"]), {n: round(float(profs[k]["ms_stage"][i]), 2) for i, n in enumerate(abi.STAGE_NAMES) if profs[k]["ms_stage"][i] > 0.3}), file=sys.stderr)
    # ---------------- end-to-end through the C ABI with HOST buffers, every step: the loop a user of the world runs --
    # upload an acceleration for EVERY body (SetAcceleration), Step, read EVERY body's transformation back (GetTransformation)
    # -- from / into pinned host memory (pr_set_body_accelerations / pr_get_body_poses).
    rows = gw.bodies()
    ids = np.arange(nb, dtype=np.uint32)
    pin_acc = torch.from_numpy(np.stack([rows["acc_x"], rows["acc_y"], rows["acc_w"]], axis=1).astype(np.float32).copy()).pin_memory()
    pin_ids = torch.from_numpy(ids.copy()).pin_memory()
    pin_xf = torch.empty((nb, 4), dtype=torch.float32).pin_memory()
    acc_np, ids_np, xf_np = pin_acc.numpy(), pin_ids.numpy(), pin_xf.numpy()
    e2e_steps = max(2, min(args.steps, 20))

    e2e_host = np.zeros(3)

    def pose_loop(k):
        for _ in range(k):
            t_a = time.perf_counter()
            gw.set_accelerations(None, acc_np)     # H2D: 12 B acceleration per body (bodies 0..n-1: no id array)
            t_b = time.perf_counter()
            gw.step(conf)
            t_c = time.perf_counter()
            gw.poses_into(xf_np)                   # D2H: 16 B transformation per body
            e2e_host[:] += [t_b - t_a, t_c - t_b, time.perf_counter() - t_c]
    pose_loop(3)  # untimed: the first bulk transfers allocate the library's staging buffers
    e2e_host[:] = 0.0
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record(stream)
    pose_loop(e2e_steps)
    ev3.record(stream)
    barrier()
    e2e_secs = ev2.elapsed_time(ev3) / 1e3
    e2e_h2d, e2e_d2h = int(nb * 12), int(nb * 16)
    # ... and the worst case: every body's WHOLE row (SetBody / GetBody value semantics, 104 B) up and down every step
    pinned_in = torch.from_numpy(rows.view(np.uint8).reshape(-1).copy()).pin_memory()
    pinned_out = torch.empty_like(pinned_in).pin_memory()
    rows_in = pinned_in.numpy().view(abi.BODY)
    rows_out = pinned_out.numpy().view(abi.BODY)
    rows_in[:] = gw.bodies()
    full_steps = max(2, min(args.steps, 8))

    def row_loop(k):
        nonlocal rows_in, rows_out
        for _ in range(k):
            gw.set_bodies(ids, rows_in)           # H2D: every body row, from pinned host memory
            gw.step(conf)
            gw.bodies_into(rows_out)              # D2H: every body row, into pinned host memory
            rows_in, rows_out = rows_out, rows_in # next step uploads what this step produced
    row_loop(2)
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev4.record(stream)
    row_loop(full_steps)
    ev5.record(stream)
    barrier()
    full_secs = ev4.elapsed_time(ev5) / 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # ---------------- reduce over ranks (max time; stats summed over NVLink)
    t = torch.tensor([secs, e2e_secs, wall_secs, full_secs], device="cuda", dtype=torch.float64)
    s = torch.tensor(np.concatenate([stats_sum, [nb, launches]]), device="cuda", dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    secs, e2e_secs, wall_secs, full_secs = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    # the step statistics of all ranks, folded like StepStats.cpp:29-38 (the only NVLink traffic of the sharded path)
    from playrho_b200 import shard
    folded, _, _ = shard.reduce_step_stats(shard.fold_step_stats(step_recs), nb * args.steps, secs, device="cuda")
    total_bodies = float(s[4])
    if rank == 0:
        peak, peak_kind = peaks()
        last = profs[-1]
        nv, npi = int(conf["reg_velocity_iters"][0]), int(conf["reg_position_iters"][0])
        bstep, terms = algorithmic_bytes(last, nv, npi)
        ms_step = 1e3 * secs / args.steps
        stage = {n: float(stage_ms[i]) / args.steps for i, n in enumerate(abi.STAGE_NAMES)}
        # dominant HBM-bound kernel: the regular phase's narrowphase launch (k_update_contacts over every contact of the
        # world; one launch per step). Its CUDA-event time is the "narrow" stage (the launch + two 4 MB memsets).
        dirty = float(np.mean([p["c_dirty"] for p in profs]))
        narrow_bytes = (184 + 16 * 2 * 3.5) * dirty      # SURVEY §8(d) a7 term: (184 + 16 (vA + vB)) B per dirty contact
        narrow_ms = stage["narrow"]
        achieved = (narrow_bytes / 1e9) / (narrow_ms / 1e3) if narrow_ms > 0 else 0.0
        # the solve (warm start, velocity and position iterations) is ONE persistent kernel (k_solve_fused): its time is the
        # "velocity" stage and "position" reads 0; with --debug-flags 4 (one launch per colour) the stages are separate
        fused = stage["position"] == 0.0
        solve_bytes = terms["a10_velocity"] + terms["a11_position"]
        solve_ms = stage["velocity"] + stage["position"]
        others = {
            ("k_solve_fused (warm start + %d velocity + %d position iterations, all colours, one persistent kernel)" % (nv, npi)) if fused
            else ("k_solve_velocity/k_solve_position (one launch per colour and iteration)" if (args.debug_flags & 128) else
                  "k_flow_pass (dataflow solver: one launch per warm-start / velocity / position pass over all coloured rows)"):
                {"algorithmic_bytes": solve_bytes, "ms": solve_ms, "achieved": (solve_bytes / 1e9) / (solve_ms / 1e3) if solve_ms > 0 else 0.0},
            "k_build_constraints (setup)":
                {"algorithmic_bytes": terms["a9_setup"], "ms": stage["setup"],
                 "achieved": (terms["a9_setup"] / 1e9) / (stage["setup"] / 1e3) if stage["setup"] > 0 else 0.0},
        }
        def entry(nbytes, ms):
            return {"algorithmic_bytes": nbytes, "ms": ms, "achieved": (nbytes / 1e9) / (ms / 1e3) if ms > 0 else 0.0}
        # broadphase: proxy AABBs + fat-box policy, LBVH refit/rebuild, overlap query -- per stage and together
        others["k_sync_proxies (a2: proxy AABB + fat-box policy)"] = entry(terms["a2_proxies"], stage["sync"])
        others["k_bvh_* (a3: Morton sort + LBVH build/refit)"] = entry(terms["a3_bvh"], stage["bvh"])
        others["k_bvh_query* (a4: overlap query)"] = entry(terms["a4_query"], stage["query"])
        others["broadphase a2-a4 together"] = entry(terms["a2_proxies"] + terms["a3_bvh"] + terms["a4_query"], stage["sync"] + stage["bvh"] + stage["query"])
        for v in others.values():
            v["frac"] = v["achieved"] / peak
        cfg = {"workload": desc, "toi": not args.no_toi, "toi_mode": "parallel" if args.toi_mode else "serial-exact",
               "toi_events_per_step": float(np.mean([p["toi_events"] for p in profs])), "toi_rounds_per_step": float(np.mean([p["toi_rounds"] for p in profs])),
               "step_conf": "default StepConf (8 vel / 3 pos iters, doToi=%s)%s" % (not args.no_toi, " + TumblerPlayRho's displaceMultiplier 4, maxRotation pi/2" if tumblers else ""),
               "l2": "working set (>1 GB) exceeds the 126 MB L2",
               "sincosf_variant": "fma" if W.load_library().pr_sincosf_variant() else "generic",
               "contacts": int(last["c_total"]), "touching": int(last["c_touching"]), "colors": int(last["n_colors"]),
               "islands": int(last["n_islands"]), "awake_bodies": int(last["n_awake"]),
               "proxies": int(last["p_total"]), "proxies_moved": int(last["p_moved"]), "contacts_dirty": int(last["c_dirty"]),
               "candidate_pairs": int(last["candidate_pairs"])}
        cfg.update(extra)
        line = {
            "metric": "body_steps_per_sec", "value": total_bodies * args.steps / secs, "unit": "body-steps/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "wall_ms_per_step": 1e3 * wall_secs / args.steps,
            "higher_is_better": True, "scaling": "strong" if tumblers else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": total_bodies * e2e_steps / e2e_secs, "unit": "body-steps/s",
                    "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h, "ms_per_step": 1e3 * e2e_secs / e2e_steps, "steps": e2e_steps,
                    "loop": "pr_set_body_accelerations(every body) + pr_step + pr_get_body_poses(every body), pinned host buffers; later steps than the device-timed window (the pile keeps growing: compare host_ms_per_step.step)",
                    "host_ms_per_step": {"set_accelerations": 1e3 * e2e_host[0] / e2e_steps, "step": 1e3 * e2e_host[1] / e2e_steps, "get_poses": 1e3 * e2e_host[2] / e2e_steps}},
            "e2e_full_rows": {"value": total_bodies * full_steps / full_secs, "unit": "body-steps/s",
                              "h2d_bytes_per_step": int(nb * abi.BODY.itemsize + 4 * nb), "d2h_bytes_per_step": int(nb * abi.BODY.itemsize),
                              "ms_per_step": 1e3 * full_secs / full_steps, "steps": full_steps,
                              "loop": "pr_set_bodies(every 104-byte row) + pr_step + pr_get_bodies(every row): the worst case of value-semantic SetBody/GetBody"},
            "gpu_launches": int(float(s[5])),
            "roofline": {"bound": "hbm", "kernel": "k_update_contacts (regular-phase narrowphase: CollideShapes + impulse carry-over for every dirty contact, one launch per step)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                         "traffic": ROOFLINE_TRAFFIC.get(args.workload), "algorithmic_bytes_per_launch": narrow_bytes,
                         "launch_ms": narrow_ms, "algorithmic_bytes_per_step": bstep,
                         "step_achieved_gbs": (bstep / 1e9) / (ms_step / 1e3), "step_frac": (bstep / 1e9) / (ms_step / 1e3) / peak,
                         "other_kernels": others},
            "stage_ms_per_step": stage,
            "clocks": sampler.summary(),
            # per-step device times of rank 0 (pr_get_step_profile: CUDA events around each step's stages)
            "step_ms_percentiles": {k: float(np.percentile([float(p["ms_total"]) for p in profs], q)) for k, q in (("p10", 10), ("p50", 50), ("p90", 90), ("max", 100))},
            "toi_ms_percentiles": {k: float(np.percentile([float(p["ms_stage"][13]) + float(p["ms_stage"][14]) for p in profs], q)) for k, q in (("p10", 10), ("p50", 50), ("p90", 90), ("max", 100))},
            "step_stats_all_ranks": {"islands_found": int(folded["reg"]["islands_found"]), "islands_solved": int(folded["reg"]["islands_solved"]),
                                     "contacts_added": int(folded["reg"]["contacts_added"]) + int(folded["pre"]["contacts_added"]) + int(folded["toi"]["contacts_added"]),
                                     "contacts_destroyed": int(folded["pre"]["contacts_destroyed"]), "bodies_slept": int(folded["reg"]["bodies_slept"]),
                                     "toi_events": int(folded["toi"]["contacts_found"]), "min_separation": float(min(folded["reg"]["min_separation"], folded["toi"]["min_separation"])),
                                     "max_inc_impulse": float(max(folded["reg"]["max_inc_impulse"], folded["toi"]["max_inc_impulse"]))},
        }
        if world_size == 1 and not tumblers and not args.no_extras and not args.no_toi and args.toi_mode == 1:
            # the same world with the reference's own TOI event order (PR_TOI_SERIAL_EXACT: one event per round): on record
            # beside the parallel rounds the headline times (both modes are pinned bit for bit to the oracle by the GPU tests)
            try:
                other = gw.clone()
                other.set_toi_mode(0)
                ostream = torch.cuda.ExternalStream(other.stream(), device=torch.device("cuda", local_rank))
                ms0 = time_steps(other, conf, ostream, 2, warmup=1)
                line["toi_serial_exact"] = {"ms_per_step": ms0, "steps": 2, "rounds_per_step": int(other.profile()["toi_rounds"]),
                                            "note": "same state, TOI events one per round in the reference's order"}
                del other
            except Exception as e:  # noqa: BLE001
                line["toi_serial_exact"] = {"ms_per_step": None, "note": "failed: %s" % e}
        if world_size == 1 and not tumblers and not args.no_extras:
            # BASELINE config 5 at N=1: the base of the strong-scaling runs (bench.py --gpus N > 1 shards these worlds)
            try:
                del gw
                targs = argparse.Namespace(**vars(args))
                targs.workload = "tumblers"
                tw, tconf, tdesc, textra = make_tumbler_arena(targs, W, 0, 1, local_rank)
                tstream = torch.cuda.ExternalStream(tw.stream(), device=torch.device("cuda", local_rank))
                tms = time_steps(tw, tconf, tstream, 10)
                tnb = int(tw.counts()["bodies"])
                line["tumblers_n1"] = {"workload": tdesc, "ms_per_step": tms, "value": tnb / (tms / 1e3), "unit": "body-steps/s", "steps": 10,
                                       "note": "the N=1 point of the strong-scaling series that --gpus 2/4/8 continue (same 8192 worlds, sharded)"}
                del tw
            except Exception as e:  # noqa: BLE001
                line["tumblers_n1"] = {"value": None, "note": "failed: %s" % e}
        if not args.no_cpu_baseline and world_size == 1:
            try:
                _checker_path()
                import refworld
                if refworld.have_ref(refworld.have_ref(True)):
                    if tumblers:
                        v, cores, rnb, wall = time_reference_tumblers(args, 20, 3)
                        line["cpu_baseline"] = {"value": v, "unit": "body-steps/s", "cores": cores, "kind": "reference", "ms_per_step": 1e3 * wall / 20,
                                                "sample": "reference Step, %d tumbler worlds x %d bodies (one per host thread), 20 steps after %d add-square steps" % (cores, rnb, args.squares)}
                    else:
                        ref = ref_scene(args, args.ref_bodies)
                        rconf = refworld.default_step_conf(do_toi=0 if args.no_toi else 1)
                        ref.step_timed(rconf, args.settle)
                        rs, _ = ref.step_timed(rconf, 20)
                        rnb = int(ref.counts()["bodies"])
                        line["cpu_baseline"] = {"value": rnb * 20 / rs, "unit": "body-steps/s", "cores": 1, "kind": "reference",
                                                "ms_per_step": 1e3 * rs / 20,
                                                "sample": "reference Step, %s scene at %d bodies, 20 steps after %d settle steps, 1 thread" % (args.workload, rnb, args.settle)}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "body-steps/s", "cores": 1, "kind": "reference", "sample": "failed: %s" % e}
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
