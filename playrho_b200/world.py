"""ctypes binding of the C ABI in include/playrho_b200.h (libplayrho_b200.so, sm_100a CUDA).

This is plumbing for tests/bench: the product is the shared library + the C++17 host type in
playrho_b200/host/. There is NO CPU fallback: if the library is missing or no CUDA device is
present, constructing a World raises.
"""
import ctypes as C
import os
import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# PLAYRHO_B200_LIB: an alternative build of the SAME CUDA library (e.g. a -DPR_TOI_TIMING profiling build)
LIB_PATH_GENERIC = os.path.join(_HERE, "libplayrho_b200.so")
LIB_PATH_FMA = os.path.join(_HERE, "libplayrho_b200_fma.so")  # same sources, -DPR_SINCOSF_FMA (pr_math.cuh)


def host_sincosf_variant():
    """Which sincosf glibc's loader binds on THIS host: "fma" when the CPU has usable FMA and AVX2 (glibc 2.39
    sysdeps/x86_64/fpu/multiarch/ifunc-fma.h: CPU_FEATURE_USABLE_P(FMA) && CPU_FEATURE_USABLE_P(AVX2)), else "generic".
    The reference's UnitVec::Get(angle) (UnitVec.cpp:27-30) inherits that choice; 34 of the 2^32 angles differ by one ulp."""
    if "-FMA" in os.environ.get("GLIBC_TUNABLES", "").upper() or "-AVX2" in os.environ.get("GLIBC_TUNABLES", "").upper():
        return "generic"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return "fma" if {"fma", "avx2"} <= flags else "generic"
    except OSError:
        pass
    return "generic"


def _select_lib_path():
    """PLAYRHO_B200_LIB = explicit path; else PLAYRHO_B200_SINCOSF = auto (default: host_sincosf_variant(), the variant this
    host's glibc -- hence the reference running here -- uses) | generic | fma."""
    explicit = os.environ.get("PLAYRHO_B200_LIB")
    if explicit:
        return explicit
    want = os.environ.get("PLAYRHO_B200_SINCOSF", "auto").lower()
    if want == "auto":
        want = host_sincosf_variant()
    if want not in ("generic", "fma"):
        raise RuntimeError("PLAYRHO_B200_SINCOSF must be generic, fma or auto (got %r)" % want)
    return LIB_PATH_FMA if want == "fma" else LIB_PATH_GENERIC


LIB_PATH = _select_lib_path()

vp, u32, f32, i32 = C.c_void_p, C.c_uint32, C.c_float, C.c_int

EXPORTS = [
    "pr_abi_version", "pr_sincosf_variant", "pr_host_sincosf_variant", "pr_world_conf_default", "pr_step_conf_default", "pr_create", "pr_destroy",
    "pr_clone", "pr_snapshot_save", "pr_snapshot_load", "pr_clear", "pr_last_error", "pr_create_shapes", "pr_create_bodies",
    "pr_create_revolute_joints", "pr_set_bodies", "pr_get_bodies", "pr_get_revolute_joints",
    "pr_get_counts", "pr_get_contacts", "pr_get_proxies", "pr_load_state", "pr_step",
    "pr_get_solve_order", "pr_get_toi_schedule", "pr_set_debug", "pr_get_step_profile", "pr_abi_sizeof", "pr_shift_origin",
    "pr_get_inv_dt0", "pr_get_stream", "pr_collide_shapes", "pr_set_toi_mode", "pr_set_sub_stepping", "pr_get_sub_stepping",
    "pr_is_step_complete", "pr_toi_batch", "pr_batch_create", "pr_batch_destroy", "pr_batch_arena",
    "pr_batch_world_count", "pr_batch_create_bodies", "pr_batch_step", "pr_batch_get_body_worlds",
    "pr_set_body_accelerations", "pr_get_body_poses", "pr_query_aabb", "pr_set_event_mask", "pr_events", "pr_ray_cast", "pr_test_points", "pr_find_closest_bodies", "pr_set_contacts", "pr_set_manifolds", "pr_set_revolute_joints", "pr_attach_shapes", "pr_set_shape_material", "pr_destroy_joints",
    "pr_destroy_bodies", "pr_detach_shapes", "pr_destroy_shapes", "pr_set_shape", "pr_set_body_enabled", "pr_set_body_type",
    "pr_create_distance_joints", "pr_create_rope_joints", "pr_create_friction_joints", "pr_create_motor_joints", "pr_create_weld_joints", "pr_create_target_joints", "pr_create_prismatic_joints", "pr_create_wheel_joints", "pr_create_pulley_joints", "pr_create_gear_joints", "pr_get_gear_joints", "pr_get_joint_work", "pr_set_body_velocities", "pr_batch_create_masked", "pr_batch_stats",
]


def bind(lib):
    """Attach argtypes/restypes to a loaded library exposing the pr_* C ABI."""
    lib.pr_abi_version.restype = i32
    lib.pr_sincosf_variant.restype = i32
    lib.pr_host_sincosf_variant.restype = i32
    lib.pr_world_conf_default.argtypes = [vp]
    lib.pr_step_conf_default.argtypes = [vp]
    lib.pr_create.argtypes = [vp, C.POINTER(vp)]
    lib.pr_destroy.argtypes = [vp]
    lib.pr_clone.argtypes = [vp, C.POINTER(vp)]
    lib.pr_clear.argtypes = [vp]
    lib.pr_snapshot_save.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.pr_snapshot_load.argtypes = [vp, C.c_uint64, i32, C.POINTER(vp)]
    lib.pr_last_error.restype = C.c_char_p
    lib.pr_last_error.argtypes = [vp]
    lib.pr_create_shapes.argtypes = [vp, u32, vp, vp, u32, vp, vp, u32, C.POINTER(u32)]
    lib.pr_create_bodies.argtypes = [vp, u32, vp, vp, vp, C.POINTER(u32)]
    lib.pr_create_revolute_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_set_bodies.argtypes = [vp, u32, vp, vp]
    lib.pr_get_bodies.argtypes = [vp, u32, u32, vp]
    lib.pr_set_body_accelerations.argtypes = [vp, u32, vp, vp]
    lib.pr_get_body_poses.argtypes = [vp, u32, u32, vp, vp]
    lib.pr_query_aabb.argtypes = [vp, u32, vp, vp, u32, vp, vp, vp, vp, C.POINTER(u32)]
    lib.pr_set_event_mask.argtypes = [vp, u32, u32]
    lib.pr_events.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
    lib.pr_ray_cast.argtypes = [vp, u32, vp, vp, u32, u32, vp, C.POINTER(u32)]
    lib.pr_set_contacts.argtypes = [vp, u32, vp]
    lib.pr_test_points.argtypes = [vp, u32, vp, vp, vp, vp]
    lib.pr_find_closest_bodies.argtypes = [vp, u32, vp, vp]
    lib.pr_set_revolute_joints.argtypes = [vp, u32, vp, vp]
    lib.pr_attach_shapes.argtypes = [vp, u32, u32, vp]
    lib.pr_destroy_joints.argtypes = [vp, u32, vp]
    lib.pr_destroy_bodies.argtypes = [vp, u32, vp]
    lib.pr_detach_shapes.argtypes = [vp, u32, u32, vp]
    lib.pr_destroy_shapes.argtypes = [vp, u32, vp]
    lib.pr_set_shape.argtypes = [vp, u32, vp, vp, vp, vp, u32, u32]
    lib.pr_set_body_enabled.argtypes = [vp, u32, u32]
    lib.pr_set_body_type.argtypes = [vp, u32, vp]
    lib.pr_create_distance_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_rope_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_friction_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_motor_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_weld_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_target_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_prismatic_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_wheel_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_pulley_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_create_gear_joints.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_get_gear_joints.argtypes = [vp, u32, u32, vp]
    lib.pr_get_joint_work.argtypes = [vp, u32, u32, vp]
    lib.pr_set_body_velocities.argtypes = [vp, u32, vp, vp]
    lib.pr_batch_create_masked.argtypes = [vp, u32, C.c_uint64, C.POINTER(vp)]
    lib.pr_batch_stats.argtypes = [vp, vp]
    lib.pr_set_shape_material.argtypes = [vp, u32, f32, f32, f32, u32]
    lib.pr_set_manifolds.argtypes = [vp, u32, vp, vp]
    lib.pr_get_revolute_joints.argtypes = [vp, u32, u32, vp]
    lib.pr_get_counts.argtypes = [vp, vp]
    lib.pr_get_contacts.argtypes = [vp, u32, vp, vp, C.POINTER(u32)]
    lib.pr_get_proxies.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_load_state.argtypes = [vp, u32, vp, u32, vp, u32, vp, vp, f32]
    lib.pr_step.argtypes = [vp, vp, vp]
    lib.pr_get_solve_order.argtypes = [vp, u32, vp, vp, C.POINTER(u32)]
    lib.pr_get_toi_schedule.argtypes = [vp, u32, vp, C.POINTER(u32)]
    lib.pr_set_debug.argtypes = [vp, u32]
    lib.pr_get_step_profile.argtypes = [vp, vp]
    lib.pr_abi_sizeof.argtypes = [i32]
    lib.pr_shift_origin.argtypes = [vp, f32, f32]
    lib.pr_get_inv_dt0.restype = f32
    lib.pr_get_inv_dt0.argtypes = [vp]
    lib.pr_get_stream.restype = vp
    lib.pr_get_stream.argtypes = [vp]
    lib.pr_collide_shapes.argtypes = [i32, u32, vp, u32, vp, vp, u32, vp, vp, vp, vp, f32, f32, vp]
    lib.pr_set_toi_mode.argtypes = [vp, u32]
    lib.pr_set_sub_stepping.argtypes = [vp, u32]
    lib.pr_get_sub_stepping.argtypes = [vp]
    lib.pr_is_step_complete.argtypes = [vp]
    lib.pr_toi_batch.argtypes = [i32, u32, vp, u32, vp, vp, u32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.pr_batch_create.argtypes = [vp, u32, C.POINTER(vp)]
    lib.pr_batch_destroy.argtypes = [vp]
    lib.pr_batch_arena.restype = vp
    lib.pr_batch_arena.argtypes = [vp]
    lib.pr_batch_world_count.restype = u32
    lib.pr_batch_world_count.argtypes = [vp]
    lib.pr_batch_create_bodies.argtypes = [vp, u32, vp, vp, vp, vp, C.POINTER(u32)]
    lib.pr_batch_step.argtypes = [vp, vp, vp]
    lib.pr_batch_get_body_worlds.argtypes = [vp, u32, u32, vp]
    return lib


_lib = None


def load_library(path=None):
    """Load the CUDA library. Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if path is None:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(os.path.basename(LIB_PATH) + " not built: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                                   "There is no CPU fallback.")
            _lib = bind(C.CDLL(LIB_PATH))
        return _lib
    return bind(C.CDLL(path))


class PrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pr_status %d: %s" % (code, msg))
        self.code = code


def default_step_conf(lib=None, **kw):
    lib = lib or load_library()
    c = abi.one(abi.STEP_CONF)
    lib.pr_step_conf_default(abi.ptr(c))
    for k, v in kw.items():
        c[k] = v
    return c


class World:
    """Device-resident world. Mirrors the reference's World usage: create shapes/bodies, Step, read back."""

    def __init__(self, scene=None, lib=None, device=-1, body_capacity=0, contact_capacity=0, proxy_capacity=0,
                 record_solve_order=False, large_variants=False, debug_flags=0):
        self.lib = lib or load_library()
        conf = abi.one(abi.WORLD_CONF)
        self.lib.pr_world_conf_default(abi.ptr(conf))
        conf["device"] = device
        conf["body_capacity"] = body_capacity
        conf["contact_capacity"] = contact_capacity
        conf["proxy_capacity"] = proxy_capacity
        h = vp()
        rc = self.lib.pr_create(abi.ptr(conf), C.byref(h))
        if rc != 0:
            raise PrError(rc, self.lib.pr_last_error(None).decode())
        self.h = h
        # bit 0: record the solve order; bits 1,2: the opt-in kernel variants (binned narrowphase, persistent fused solver)
        if record_solve_order or large_variants or debug_flags:
            self.lib.pr_set_debug(self.h, (1 if record_solve_order else 0) | (6 if large_variants else 0) | debug_flags)
        if scene is not None:
            self.create(scene)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pr_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise PrError(rc, self.lib.pr_last_error(self.h).decode())

    def create(self, sc):
        first = u32(0)
        self._chk(self.lib.pr_create_shapes(self.h, len(sc["shapes"]), abi.ptr(sc["shapes"]), abi.ptr(sc["children"]),
                                            len(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]),
                                            len(sc["verts"]), C.byref(first)))
        self._chk(self.lib.pr_create_bodies(self.h, len(sc["bodies"]), abi.ptr(sc["bodies"]), abi.ptr(sc["shape_offsets"]),
                                            abi.ptr(sc["shape_ids"]), C.byref(first)))
        if "joints" in sc and len(sc["joints"]):
            for a, b, gear in abi.joint_batches(sc["joints"]):
                if gear:
                    g = np.ascontiguousarray(sc["gears"][a:b], dtype=abi.GEAR)
                    self._chk(self.lib.pr_create_gear_joints(self.h, 1, abi.ptr(g), C.byref(first)))
                else:
                    rows = np.ascontiguousarray(sc["joints"][a:b])
                    self._chk(self.lib.pr_create_revolute_joints(self.h, len(rows), abi.ptr(rows), C.byref(first)))

    def step(self, conf):
        st = abi.one(abi.STEP_STATS)
        self._chk(self.lib.pr_step(self.h, abi.ptr(conf), abi.ptr(st)))
        return st[0]

    def counts(self):
        c = abi.one(abi.COUNTS)
        self._chk(self.lib.pr_get_counts(self.h, abi.ptr(c)))
        return c[0]

    def bodies(self, first=0, n=None):
        if n is None:
            n = int(self.counts()["bodies"]) - first
        out = np.zeros(n, abi.BODY)
        self._chk(self.lib.pr_get_bodies(self.h, first, n, abi.ptr(out)))
        return out

    def set_bodies(self, ids, rows):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=abi.BODY)
        self._chk(self.lib.pr_set_bodies(self.h, len(ids), abi.ptr(ids), abi.ptr(rows)))

    def set_accelerations(self, ids, acc3):
        """Raw write of (linear x, linear y, angular) acceleration of bodies `ids` (pr_set_body_accelerations)."""
        acc3 = np.ascontiguousarray(acc3, dtype=np.float32)
        if ids is None:  # bodies [0, len(acc3))
            self._chk(self.lib.pr_set_body_accelerations(self.h, acc3.size // 3, None, abi.ptr(acc3)))
            return
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        assert acc3.size == 3 * len(ids)
        self._chk(self.lib.pr_set_body_accelerations(self.h, len(ids), abi.ptr(ids), abi.ptr(acc3)))

    def poses_into(self, xf4, vel3=None, first=0):
        """Transformations (px, py, cos, sin) -- and velocities if vel3 is given -- of bodies [first, first+len(xf4)) into
        caller-owned float32 arrays (e.g. pinned)."""
        assert xf4.dtype == np.float32 and xf4.size % 4 == 0
        n = xf4.size // 4
        self._chk(self.lib.pr_get_body_poses(self.h, first, n, abi.ptr(xf4), abi.ptr(vel3) if vel3 is not None else None))
        return xf4

    def query_aabb(self, boxes, box_world=None, cap=None):
        """Query(tree, AABB) for many boxes (rows of min x, min y, max x, max y): arrays (box index, body, shape, child) of every
        proxy whose fat AABB overlaps the box (pr_query_aabb). Order unspecified."""
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
        bw = None if box_world is None else np.ascontiguousarray(box_world, dtype=np.uint32)
        cap = cap or max(1024, 64 * len(boxes))
        while True:
            out = [np.zeros(cap, np.uint32) for _ in range(4)]
            k = u32(0)
            self._chk(self.lib.pr_query_aabb(self.h, len(boxes), abi.ptr(boxes), abi.ptr(bw), cap, abi.ptr(out[0]), abi.ptr(out[1]),
                                             abi.ptr(out[2]), abi.ptr(out[3]), C.byref(k)))
            if k.value <= cap:
                return tuple(o[:k.value] for o in out)
            cap = k.value

    def ray_cast(self, rays, mode=abi.RAY_ALL, ray_world=None, cap=None):
        """RayCast(world, input, callback) for many rays (rows of p1 x, p1 y, p2 x, p2 y, maxFraction): abi.RAY_HIT rows --
        every hit (RAY_ALL, order unspecified) or the closest hit per ray (RAY_CLOSEST, row i = ray i, body 0xFFFFFFFF = miss)."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 5)
        rw = None if ray_world is None else np.ascontiguousarray(ray_world, dtype=np.uint32)
        cap = len(rays) if mode == abi.RAY_CLOSEST else (cap or max(1024, 16 * len(rays)))
        while True:
            out = np.zeros(max(1, cap), abi.RAY_HIT)
            k = u32(0)
            self._chk(self.lib.pr_ray_cast(self.h, len(rays), abi.ptr(rays), abi.ptr(rw), mode, cap, abi.ptr(out), C.byref(k)))
            if k.value <= cap:
                return out[:k.value]
            cap = k.value

    def test_points(self, body_ids, shape_ids, points):
        """TestPoint(world, body, shape, point) for many triples (pr_test_points): uint32 array of 0/1."""
        b = np.ascontiguousarray(body_ids, dtype=np.uint32); s_ = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
        out = np.zeros(len(b), np.uint32)
        self._chk(self.lib.pr_test_points(self.h, len(b), abi.ptr(b), abi.ptr(s_), abi.ptr(p), abi.ptr(out)))
        return out

    def find_closest_bodies(self, points):
        """FindClosestBody(world, location) for many locations (pr_find_closest_bodies)."""
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
        out = np.zeros(len(p), np.uint32)
        self._chk(self.lib.pr_find_closest_bodies(self.h, len(p), abi.ptr(p), abi.ptr(out)))
        return out

    def set_shape_material(self, shape, friction, restitution, density, sensor):
        """SetShape with unchanged geometry and filter (pr_set_shape_material)."""
        self._chk(self.lib.pr_set_shape_material(self.h, int(shape), friction, restitution, density, int(sensor)))

    def attach_shapes(self, body, shape_ids):
        """Attach(world, body, shape) for each shape id (pr_attach_shapes); upload the body row (new mass) with set_bodies."""
        ids = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        self._chk(self.lib.pr_attach_shapes(self.h, int(body), len(ids), abi.ptr(ids)))

    def joint_work(self):
        """pr_get_joint_work over the whole joint table: what InitVelocity left for each joint in the last Step (island == 0xFFFFFFFF: not solved)"""
        n = int(self.counts()["joints"])
        out = np.zeros(n, abi.JOINT_WORK)
        if n:
            self._chk(self.lib.pr_get_joint_work(self.h, 0, n, abi.ptr(out)))
        return out

    def gear_joints(self):
        """pr_get_gear_joints over the whole joint table (non-gear rows read back with body_a == 0xFFFFFFFF)"""
        n = int(self.counts()["joints"])
        out = np.zeros(n, abi.GEAR)
        if n:
            self._chk(self.lib.pr_get_gear_joints(self.h, 0, n, abi.ptr(out)))
        return out

    def destroy_joints(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._chk(self.lib.pr_destroy_joints(self.h, len(ids), abi.ptr(ids)))

    def add_body(self, row, shape_ids=()):
        """CreateBody for one body row with the shapes listed; returns its id (a destroyed body's id first, last freed first)."""
        row = np.ascontiguousarray(row, dtype=abi.BODY).reshape(1)
        off = np.array([0, len(shape_ids)], np.uint32)
        ids = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        first = u32(0)
        self._chk(self.lib.pr_create_bodies(self.h, 1, abi.ptr(row), abi.ptr(off), abi.ptr(ids), C.byref(first)))
        return first.value

    def add_shape(self, row, children, vertices, normals):
        """CreateShape for one flattened shape; returns its id (a destroyed shape's id first)."""
        row = np.ascontiguousarray(row, dtype=abi.SHAPE).reshape(1).copy()
        children = np.ascontiguousarray(children, dtype=abi.CHILD)
        vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 2)
        normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 2)
        row["child_offset"] = 0; row["child_count"] = len(children)
        first = u32(0)
        self._chk(self.lib.pr_create_shapes(self.h, 1, abi.ptr(row), abi.ptr(children), len(children), abi.ptr(vertices), abi.ptr(normals), len(vertices), C.byref(first)))
        return first.value

    def destroy_bodies(self, ids):
        """Destroy(world, BodyID) for each id (pr_destroy_bodies)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._chk(self.lib.pr_destroy_bodies(self.h, len(ids), abi.ptr(ids)))

    def detach_shapes(self, body, shape_ids):
        """Detach(world, body, shape) for each shape id (pr_detach_shapes); upload the body row (new mass) with set_bodies."""
        ids = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        self._chk(self.lib.pr_detach_shapes(self.h, int(body), len(ids), abi.ptr(ids)))

    def destroy_shapes(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._chk(self.lib.pr_destroy_shapes(self.h, len(ids), abi.ptr(ids)))

    def set_shape(self, shape, row, children=None, vertices=None, normals=None):
        """SetShape with any new value (pr_set_shape); children/vertices/normals given = new geometry."""
        row = np.ascontiguousarray(row, dtype=abi.SHAPE).reshape(1).copy()
        geom = children is not None
        if geom:
            children = np.ascontiguousarray(children, dtype=abi.CHILD)
            vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 2)
            normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 2)
            row["child_count"] = len(children)
            self._chk(self.lib.pr_set_shape(self.h, int(shape), abi.ptr(row), abi.ptr(children), abi.ptr(vertices), abi.ptr(normals), len(vertices), 1))
        else:
            self._chk(self.lib.pr_set_shape(self.h, int(shape), abi.ptr(row), None, None, None, 0, 0))

    def set_body_velocities(self, vel3, ids=None):
        """Body::SetVelocity for many bodies at once (pr_set_body_velocities); vel3: (n, 3) linear x, y, angular."""
        vel3 = np.ascontiguousarray(vel3, dtype=np.float32).reshape(-1, 3)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._chk(self.lib.pr_set_body_velocities(self.h, len(vel3), abi.ptr(ids) if ids is not None else None, abi.ptr(vel3)))

    def set_body_enabled(self, body, enabled):
        self._chk(self.lib.pr_set_body_enabled(self.h, int(body), 1 if enabled else 0))

    def set_body_type(self, body, row):
        row = np.ascontiguousarray(row, dtype=abi.BODY).reshape(1)
        self._chk(self.lib.pr_set_body_type(self.h, int(body), abi.ptr(row)))

    def create_joints(self, rows):
        rows = np.ascontiguousarray(rows, dtype=abi.REVOLUTE)
        first = u32(0)
        self._chk(self.lib.pr_create_revolute_joints(self.h, len(rows), abi.ptr(rows), C.byref(first)))
        return first.value

    def set_joints(self, ids, rows):
        """SetJoint for revolute joints (pr_set_revolute_joints)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=abi.REVOLUTE)
        self._chk(self.lib.pr_set_revolute_joints(self.h, len(ids), abi.ptr(ids), abi.ptr(rows)))

    def set_contacts(self, rows):
        """SetContact for every row of an abi.CONTACT array (row["id"] names the contact)."""
        rows = np.ascontiguousarray(rows, dtype=abi.CONTACT)
        self._chk(self.lib.pr_set_contacts(self.h, len(rows), abi.ptr(rows)))

    def set_manifolds(self, ids, rows):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=abi.MANIFOLD)
        self._chk(self.lib.pr_set_manifolds(self.h, len(ids), abi.ptr(ids), abi.ptr(rows)))

    def shift_origin(self, x, y):
        """ShiftOrigin(world, newOrigin) (pr_shift_origin)."""
        self._chk(self.lib.pr_shift_origin(self.h, float(x), float(y)))

    def set_event_mask(self, mask, capacity=1 << 16):
        """Record contact events (abi.EVENT_* bits) from the next step on (pr_set_event_mask)."""
        self._chk(self.lib.pr_set_event_mask(self.h, mask, capacity))
        self._ev_cap = capacity

    def events(self, kind):
        """Events of the last step (pr_events): abi.CONTACT_EVENT rows for begin/end, abi.IMPULSE_EVENT for post-solve, abi.PRESOLVE_EVENT for pre-solve."""
        dt = abi.IMPULSE_EVENT if kind == abi.EVENT_POST_SOLVE else abi.PRESOLVE_EVENT if kind == abi.EVENT_PRE_SOLVE else abi.CONTACT_EVENT
        cap = getattr(self, "_ev_cap", 0)
        out = np.zeros(max(1, cap), dt)
        k = u32(0)
        self._chk(self.lib.pr_events(self.h, kind, abi.ptr(out), cap, C.byref(k)))
        return out[:k.value]

    def joints(self):
        n = int(self.counts()["joints"])
        out = np.zeros(max(1, n), abi.REVOLUTE)
        if n:
            self._chk(self.lib.pr_get_revolute_joints(self.h, 0, n, abi.ptr(out)))
        return out[:n]

    def bodies_into(self, out, first=0):
        """Download body rows into a caller-owned (e.g. pinned) abi.BODY array."""
        self._chk(self.lib.pr_get_bodies(self.h, first, len(out), abi.ptr(out)))
        return out

    def contacts(self):
        n = max(1, int(self.counts()["contacts"]))
        cs, ms = np.zeros(n, abi.CONTACT), np.zeros(n, abi.MANIFOLD)
        k = u32(0)
        self._chk(self.lib.pr_get_contacts(self.h, n, abi.ptr(cs), abi.ptr(ms), C.byref(k)))
        return cs[:k.value], ms[:k.value]

    def proxies(self):
        n = max(1, int(self.counts()["proxies"]))
        out = np.zeros(n, abi.PROXY)
        k = u32(0)
        self._chk(self.lib.pr_get_proxies(self.h, n, abi.ptr(out), C.byref(k)))
        return out[:k.value]

    def load_state(self, bodies, proxies, contacts, manifolds, inv_dt0):
        self._chk(self.lib.pr_load_state(self.h, len(bodies), abi.ptr(bodies), len(proxies), abi.ptr(proxies),
                                         len(contacts), abi.ptr(contacts), abi.ptr(manifolds), inv_dt0))

    def solve_order(self):
        k = u32(0)
        self.lib.pr_get_solve_order(self.h, 0, None, None, C.byref(k))
        ids, cols = np.zeros(max(1, k.value), np.uint32), np.zeros(max(1, k.value), np.uint32)
        self.lib.pr_get_solve_order(self.h, len(ids), abi.ptr(ids), abi.ptr(cols), C.byref(k))
        return ids[:k.value], cols[:k.value]

    def toi_schedule(self):
        """(round, contact id) rows of the last step's TOI events (needs record_solve_order)."""
        k = u32(0)
        self.lib.pr_get_toi_schedule(self.h, 0, None, C.byref(k))
        rows = np.zeros((max(1, k.value), 2), np.uint32)
        self._chk(self.lib.pr_get_toi_schedule(self.h, len(rows), abi.ptr(rows), C.byref(k)))
        return rows[:k.value]

    def profile(self):
        p = abi.one(abi.STEP_PROFILE)
        self.lib.pr_get_step_profile(self.h, abi.ptr(p))
        return p[0]

    def inv_dt0(self):
        return self.lib.pr_get_inv_dt0(self.h)

    def stream(self):
        """cudaStream_t (as an int) the world issues its work on; wrap with torch.cuda.ExternalStream for events."""
        return self.lib.pr_get_stream(self.h) or 0

    def set_toi_mode(self, mode):
        """0 = serial-exact (reference event order, bit-exact), 1 = parallel independent events."""
        self._chk(self.lib.pr_set_toi_mode(self.h, mode))

    def set_sub_stepping(self, flag):
        self._chk(self.lib.pr_set_sub_stepping(self.h, 1 if flag else 0))

    def is_step_complete(self):
        return bool(self.lib.pr_is_step_complete(self.h))

    def clone(self):
        h = vp()
        self._chk(self.lib.pr_clone(self.h, C.byref(h)))
        other = World.__new__(World)
        other.lib, other.h = self.lib, h
        return other

    def snapshot(self):
        """The whole world as bytes (pr_snapshot_save): World.restore(bytes) continues bit for bit like this world."""
        n = C.c_uint64(0)
        self._chk(self.lib.pr_snapshot_save(self.h, None, 0, C.byref(n)))
        buf = np.zeros(n.value, np.uint8)
        self._chk(self.lib.pr_snapshot_save(self.h, abi.ptr(buf), buf.nbytes, C.byref(n)))
        return buf[:n.value]

    @classmethod
    def restore(cls, buf, lib=None, device=-1):
        lib = lib or load_library()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        h = vp()
        rc = lib.pr_snapshot_load(abi.ptr(buf), buf.nbytes, device, C.byref(h))
        if rc != 0:
            raise PrError(rc, lib.pr_last_error(None).decode())
        other = World.__new__(World)
        other.lib, other.h = lib, h
        return other


class Batch(World):
    """n_worlds independent worlds in one device arena (pr_batch_*, BASELINE config 5). The arena is an ordinary
    world handle, so every World accessor works on it; body ids are arena-wide."""

    def __init__(self, n_worlds, lib=None, device=-1, body_capacity=0, contact_capacity=0, proxy_capacity=0,
                 record_solve_order=False):
        self.lib = lib or load_library()
        conf = abi.one(abi.WORLD_CONF)
        self.lib.pr_world_conf_default(abi.ptr(conf))
        conf["device"] = device
        conf["body_capacity"] = body_capacity
        conf["contact_capacity"] = contact_capacity
        conf["proxy_capacity"] = proxy_capacity
        b = vp()
        rc = self.lib.pr_batch_create(abi.ptr(conf), n_worlds, C.byref(b))
        if rc != 0:
            raise PrError(rc, self.lib.pr_last_error(None).decode())
        self.b = b
        self.h = vp(self.lib.pr_batch_arena(b))
        self.n_worlds = n_worlds
        if record_solve_order:
            self.lib.pr_set_debug(self.h, 1)

    def close(self):
        if getattr(self, "b", None):
            self.lib.pr_batch_destroy(self.b)
            self.b = None
            self.h = None

    def create_shapes(self, sc):
        first = u32(0)
        self._chk(self.lib.pr_create_shapes(self.h, len(sc["shapes"]), abi.ptr(sc["shapes"]), abi.ptr(sc["children"]),
                                            len(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]),
                                            len(sc["verts"]), C.byref(first)))
        return first.value

    def create_bodies(self, bodies, world_index, shape_offsets, shape_ids):
        first = u32(0)
        bodies = np.ascontiguousarray(bodies, dtype=abi.BODY)
        world_index = np.ascontiguousarray(world_index, dtype=np.uint32)
        shape_offsets = np.ascontiguousarray(shape_offsets, dtype=np.uint32)
        shape_ids = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        assert len(world_index) == len(bodies) and len(shape_offsets) == len(bodies) + 1
        self._chk(self.lib.pr_batch_create_bodies(self.b, len(bodies), abi.ptr(bodies), abi.ptr(world_index),
                                                  abi.ptr(shape_offsets), abi.ptr(shape_ids), C.byref(first)))
        return first.value

    def create_joints(self, joints, gears=None):
        """rows of the joint table in id order; gears: the whole values of the gear joints among them (abi.GEAR, same length), if any"""
        first = u32(0)
        joints = np.ascontiguousarray(joints, dtype=abi.REVOLUTE)
        start = None
        for a, b, gear in abi.joint_batches(joints):
            if gear:
                g = np.ascontiguousarray(gears[a:b], dtype=abi.GEAR)
                self._chk(self.lib.pr_create_gear_joints(self.h, 1, abi.ptr(g), C.byref(first)))
            else:
                rows = np.ascontiguousarray(joints[a:b])
                self._chk(self.lib.pr_create_revolute_joints(self.h, len(rows), abi.ptr(rows), C.byref(first)))
            start = first.value if start is None else start
        return 0 if start is None else start

    def step(self, conf):
        st = abi.one(abi.STEP_STATS)
        self._chk(self.lib.pr_batch_step(self.b, abi.ptr(conf), abi.ptr(st)))
        return st[0]

    def body_worlds(self):
        n = int(self.counts()["bodies"])
        out = np.zeros(max(1, n), np.uint32)
        self._chk(self.lib.pr_batch_get_body_worlds(self.b, 0, n, abi.ptr(out)))
        return out[:n]


def toi_batch(children, verts, normals, a_child, sweep_a, b_child, sweep_b, conf, lib=None, device=-1):
    """GetToiViaSat on arrays through the C ABI (pr_toi_batch)."""
    lib = lib or load_library()
    n = len(a_child)
    t, st, it = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros((n, 3), np.uint32)
    rc = lib.pr_toi_batch(device, n, abi.ptr(children), len(children), abi.ptr(verts), abi.ptr(normals), len(verts),
                          abi.ptr(np.ascontiguousarray(a_child, np.uint32)), abi.ptr(np.ascontiguousarray(sweep_a, np.float32)),
                          abi.ptr(np.ascontiguousarray(b_child, np.uint32)), abi.ptr(np.ascontiguousarray(sweep_b, np.float32)),
                          abi.ptr(conf), abi.ptr(t), abi.ptr(st), abi.ptr(it))
    if rc != 0:
        raise PrError(rc, lib.pr_last_error(None).decode())
    return t, st, it


def collide_shapes(children, verts, normals, a_child, xf_a, b_child, xf_b, linear_slop=0.005,
                   max_circles_ratio=10.0, lib=None, device=-1):
    lib = lib or load_library()
    n = len(a_child)
    out = np.zeros(max(1, n), abi.MANIFOLD)
    a_child = np.ascontiguousarray(a_child, np.uint32)
    b_child = np.ascontiguousarray(b_child, np.uint32)
    xf_a = np.ascontiguousarray(xf_a, np.float32)
    xf_b = np.ascontiguousarray(xf_b, np.float32)
    rc = lib.pr_collide_shapes(device, n, abi.ptr(children), len(children), abi.ptr(verts), abi.ptr(normals),
                               len(verts), abi.ptr(a_child), abi.ptr(xf_a), abi.ptr(b_child), abi.ptr(xf_b),
                               linear_slop, max_circles_ratio, abi.ptr(out))
    if rc != 0:
        raise PrError(rc, lib.pr_last_error(None).decode())
    return out[:n]
