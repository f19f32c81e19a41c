"""ctypes wrappers (TEST INFRASTRUCTURE) over
  - oracle/_ref/libref_harness{,32}.so : the real reference, built by oracle/Makefile
  - oracle/libpr_oracle.so             : the C++ restatement
Both speak the POD layouts of include/playrho_b200.h (playrho_b200/abi.py)."""
import ctypes as C
import os
import numpy as np

from playrho_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
REF32_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness32.so")
ORC_SO = os.path.join(ROOT, "oracle", "libpr_oracle.so")

vp, u32, f32, i32 = C.c_void_p, C.c_uint32, C.c_float, C.c_int


def have_ref(ids32=False):
    return os.path.exists(REF32_SO if ids32 else REF_SO)


def have_oracle():
    return os.path.exists(ORC_SO)


_libs = {}


def _ref_lib(ids32=False):
    key = "ref32" if ids32 else "ref"
    if key not in _libs:
        lib = C.CDLL(REF32_SO if ids32 else REF_SO)
        lib.ref_world_create.restype = vp
        lib.ref_world_create.argtypes = [f32, f32]
        lib.ref_world_destroy.argtypes = [vp]
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_last_error.argtypes = [vp]
        common = [f32, f32, f32, i32, C.c_uint, C.c_uint, i32]
        lib.ref_add_disk.restype = u32
        lib.ref_add_disk.argtypes = [vp, f32, f32, f32] + common
        lib.ref_add_polygon.restype = u32
        lib.ref_add_polygon.argtypes = [vp, vp, u32, f32] + common
        lib.ref_add_box.restype = u32
        lib.ref_add_box.argtypes = [vp, f32, f32, f32, f32, f32, f32] + common
        lib.ref_add_edge.restype = u32
        lib.ref_add_edge.argtypes = [vp, f32, f32, f32, f32, f32] + common
        lib.ref_add_chain.restype = u32
        lib.ref_add_chain.argtypes = [vp, vp, u32, f32] + common
        lib.ref_add_body.restype = u32
        lib.ref_add_body.argtypes = [vp, i32] + [f32] * 11 + [C.c_uint, vp, u32]
        lib.ref_add_revolute.restype = u32
        lib.ref_add_revolute.argtypes = [vp, u32, u32, f32, f32, i32, f32, f32, i32]
        lib.ref_export_revolute_joints.argtypes = [vp, vp, u32]
        lib.ref_set_body_velocity.argtypes = [vp, u32, f32, f32, f32]
        lib.ref_set_body_transform.argtypes = [vp, u32, f32, f32, f32]
        lib.ref_step.argtypes = [vp, vp, vp]
        lib.ref_step_timed.restype = C.c_double
        lib.ref_step_timed.argtypes = [vp, vp, u32, vp]
        lib.ref_get_counts.argtypes = [vp, vp]
        lib.ref_export_shapes.argtypes = [vp, vp, vp, vp, vp]
        lib.ref_export_bodies.argtypes = [vp, vp]
        lib.ref_export_body_shapes.argtypes = [vp, vp, vp, u32]
        lib.ref_export_proxies.argtypes = [vp, vp, u32]
        lib.ref_export_contacts.argtypes = [vp, vp, vp, u32]
        lib.ref_get_inv_dt0.restype = f32
        lib.ref_get_inv_dt0.argtypes = [vp]
        lib.ref_collide_shapes.argtypes = [u32, vp, vp, vp, vp, vp, vp, vp, f32, f32, vp]
        lib.ref_unitvec_from_angle.argtypes = [u32, vp, vp]
        lib.ref_toi.argtypes = [u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.ref_step_conf_default.argtypes = [vp]
        _libs[key] = lib
    return _libs[key]


def _orc_lib():
    if "orc" not in _libs:
        lib = C.CDLL(ORC_SO)
        lib.orc_create.restype = vp
        lib.orc_destroy.argtypes = [vp]
        lib.orc_last_error.restype = C.c_char_p
        lib.orc_last_error.argtypes = [vp]
        lib.orc_create_shapes.argtypes = [vp, u32, vp, vp, u32, vp, vp, u32]
        lib.orc_create_bodies.argtypes = [vp, u32, vp, vp, vp]
        lib.orc_create_revolute_joints.argtypes = [vp, u32, vp]
        lib.orc_set_bodies.argtypes = [vp, u32, vp, vp]
        lib.orc_load_state.argtypes = [vp, u32, vp, u32, vp, u32, vp, vp, f32]
        lib.orc_step.argtypes = [vp, vp, vp, vp, u32]
        lib.orc_get_counts.argtypes = [vp, vp]
        lib.orc_get_bodies.argtypes = [vp, u32, u32, vp]
        lib.orc_get_contacts.argtypes = [vp, u32, vp, vp, vp]
        lib.orc_get_proxies.argtypes = [vp, u32, vp, vp]
        lib.orc_get_revolute_joints.argtypes = [vp, u32, u32, vp]
        lib.orc_get_inv_dt0.restype = f32
        lib.orc_get_inv_dt0.argtypes = [vp]
        lib.orc_collide_shapes.argtypes = [u32, vp, vp, vp, vp, vp, vp, vp, f32, f32, vp]
        lib.orc_step_conf_default.argtypes = [vp]
        lib.orc_toi.argtypes = [u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.orc_sincosf_restated.argtypes = [u32, vp, vp]
        _libs["orc"] = lib
    return _libs["orc"]


def default_step_conf(**kw):
    c = abi.one(abi.STEP_CONF)
    _orc_lib().orc_step_conf_default(abi.ptr(c))
    for k, v in kw.items():
        c[k] = v
    return c


STATIC, KINEMATIC, DYNAMIC = 0, 1, 2


class RefWorld:
    """The real reference world (playrho::d2::World{AabbTreeWorld}) behind oracle/ref_harness.cpp."""

    def __init__(self, ids32=False, min_vr=0.0, max_vr=0.0):
        self.lib = _ref_lib(ids32)
        self.h = self.lib.ref_world_create(min_vr, max_vr)
        assert self.h

    def close(self):
        if self.h:
            self.lib.ref_world_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, r):
        if r == 0xFFFFFFFF or r == -1:
            raise RuntimeError(self.lib.ref_last_error(self.h).decode())
        return r

    @staticmethod
    def _common(friction, restitution, density, sensor, cat, mask, group):
        return (friction, restitution, density, int(sensor), cat, mask, group)

    def add_disk(self, radius, loc=(0.0, 0.0), friction=0.2, restitution=0.0, density=0.0,
                 sensor=False, cat=1, mask=0xFFFF, group=0):
        return self._chk(self.lib.ref_add_disk(self.h, radius, loc[0], loc[1],
                                               *self._common(friction, restitution, density, sensor, cat, mask, group)))

    def add_polygon(self, pts, vertex_radius=0.0, friction=0.2, restitution=0.0, density=0.0,
                    sensor=False, cat=1, mask=0xFFFF, group=0):
        xy = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1)
        return self._chk(self.lib.ref_add_polygon(self.h, abi.ptr(xy), len(xy) // 2, vertex_radius,
                                                  *self._common(friction, restitution, density, sensor, cat, mask, group)))

    def add_box(self, hx, hy, center=(0.0, 0.0), angle=0.0, vertex_radius=0.0, friction=0.2,
                restitution=0.0, density=0.0, sensor=False, cat=1, mask=0xFFFF, group=0):
        return self._chk(self.lib.ref_add_box(self.h, hx, hy, center[0], center[1], angle, vertex_radius,
                                              *self._common(friction, restitution, density, sensor, cat, mask, group)))

    def add_edge(self, p0, p1, vertex_radius=0.0, friction=0.2, restitution=0.0, density=0.0,
                 sensor=False, cat=1, mask=0xFFFF, group=0):
        return self._chk(self.lib.ref_add_edge(self.h, p0[0], p0[1], p1[0], p1[1], vertex_radius,
                                               *self._common(friction, restitution, density, sensor, cat, mask, group)))

    def add_chain(self, pts, vertex_radius=0.0, friction=0.2, restitution=0.0, density=0.0,
                  sensor=False, cat=1, mask=0xFFFF, group=0):
        xy = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1)
        return self._chk(self.lib.ref_add_chain(self.h, abi.ptr(xy), len(xy) // 2, vertex_radius,
                                                *self._common(friction, restitution, density, sensor, cat, mask, group)))

    def add_body(self, btype, pos=(0.0, 0.0), angle=0.0, vel=(0.0, 0.0), w=0.0, acc=(0.0, 0.0),
                 aw=0.0, lin_damp=0.0, ang_damp=0.0, bullet=False, fixed_rotation=False,
                 allow_sleep=True, awake=True, enabled=True, shapes=()):
        bits = (1 if bullet else 0) | (2 if fixed_rotation else 0) | (0 if allow_sleep else 4) | \
               (0 if awake else 8) | (0 if enabled else 16)
        sh = np.asarray(list(shapes), dtype=np.uint32)
        return self._chk(self.lib.ref_add_body(self.h, btype, pos[0], pos[1], angle, vel[0], vel[1], w,
                                               acc[0], acc[1], aw, lin_damp, ang_damp, bits,
                                               abi.ptr(sh) if len(sh) else None, len(sh)))

    def add_revolute(self, body_a, body_b, anchor, enable_motor=False, motor_speed=0.0,
                     max_motor_torque=0.0, collide_connected=False):
        return self._chk(self.lib.ref_add_revolute(self.h, body_a, body_b, anchor[0], anchor[1],
                                                   int(enable_motor), motor_speed, max_motor_torque,
                                                   int(collide_connected)))

    def step(self, conf):
        st = abi.one(abi.STEP_STATS)
        self._chk(self.lib.ref_step(self.h, abi.ptr(conf), abi.ptr(st)))
        return st[0]

    def step_timed(self, conf, n):
        st = abi.one(abi.STEP_STATS)
        secs = self.lib.ref_step_timed(self.h, abi.ptr(conf), n, abi.ptr(st))
        return secs, st[0]

    def set_shape_material(self, shape, friction, restitution, density, sensor):
        self.lib.ref_set_shape_material.argtypes = [vp, u32, f32, f32, f32, u32]
        self._chk(self.lib.ref_set_shape_material(self.h, int(shape), friction, restitution, density, int(sensor)))

    def attach_shape(self, body, shape):
        """Attach(world, body, shape) with the mass data reset (the reference's default)."""
        self.lib.ref_attach_shape.argtypes = [vp, u32, u32]
        self._chk(self.lib.ref_attach_shape(self.h, int(body), int(shape)))

    def add_distance(self, body_a, body_b, anchor_a, anchor_b, frequency=0.0, damping_ratio=0.0, collide_connected=False):
        self.lib.ref_add_distance.argtypes = [vp, u32, u32, f32, f32, f32, f32, f32, f32, i32]
        self.lib.ref_add_distance.restype = u32
        jid = self.lib.ref_add_distance(self.h, int(body_a), int(body_b), anchor_a[0], anchor_a[1], anchor_b[0], anchor_b[1],
                                        frequency, damping_ratio, 1 if collide_connected else 0)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def add_rope(self, body_a, body_b, local_a, local_b, max_length, collide_connected=False):
        self.lib.ref_add_rope.argtypes = [vp, u32, u32, f32, f32, f32, f32, f32, i32]
        self.lib.ref_add_rope.restype = u32
        jid = self.lib.ref_add_rope(self.h, int(body_a), int(body_b), local_a[0], local_a[1], local_b[0], local_b[1], max_length,
                                    1 if collide_connected else 0)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def add_weld(self, body_a, body_b, anchor, frequency=0.0, damping_ratio=0.0):
        self.lib.ref_add_weld.argtypes = [vp, u32, u32, f32, f32, f32, f32]
        self.lib.ref_add_weld.restype = u32
        jid = self.lib.ref_add_weld(self.h, int(body_a), int(body_b), anchor[0], anchor[1], frequency, damping_ratio)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def add_target(self, body_b, target, anchor, frequency=5.0, damping_ratio=0.7, max_force=1000.0, body_a=0xFFFFFFFF):
        self.lib.ref_add_target.argtypes = [vp, u32, u32, f32, f32, f32, f32, f32, f32, f32]
        self.lib.ref_add_target.restype = u32
        jid = self.lib.ref_add_target(self.h, int(body_a), int(body_b), target[0], target[1], anchor[0], anchor[1], frequency, damping_ratio, max_force)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def set_target(self, joint, target):
        self.lib.ref_set_target.argtypes = [vp, u32, f32, f32]
        self._chk(self.lib.ref_set_target(self.h, int(joint), target[0], target[1]))

    def add_prismatic(self, body_a, body_b, anchor, axis, limits=None, motor=None):
        """limits = (lower, upper) or None; motor = (speed, max force) or None"""
        self.lib.ref_add_prismatic.argtypes = [vp, u32, u32, f32, f32, f32, f32, i32, f32, f32, i32, f32, f32]
        self.lib.ref_add_prismatic.restype = u32
        lo, hi = limits if limits else (0.0, 0.0)
        sp, mf = motor if motor else (0.0, 0.0)
        jid = self.lib.ref_add_prismatic(self.h, int(body_a), int(body_b), anchor[0], anchor[1], axis[0], axis[1], 1 if limits else 0, lo, hi,
                                         1 if motor else 0, sp, mf)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def add_wheel(self, body_a, body_b, anchor, axis, frequency=2.0, damping_ratio=0.7, motor=None):
        self.lib.ref_add_wheel.argtypes = [vp, u32, u32, f32, f32, f32, f32, f32, f32, i32, f32, f32]
        self.lib.ref_add_wheel.restype = u32
        sp, mt = motor if motor else (0.0, 0.0)
        jid = self.lib.ref_add_wheel(self.h, int(body_a), int(body_b), anchor[0], anchor[1], axis[0], axis[1], frequency, damping_ratio,
                                     1 if motor else 0, sp, mt)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def add_pulley(self, body_a, body_b, ground_a, ground_b, anchor_a, anchor_b, ratio=1.0):
        self.lib.ref_add_pulley.argtypes = [vp, u32, u32] + [f32] * 9
        self.lib.ref_add_pulley.restype = u32
        jid = self.lib.ref_add_pulley(self.h, int(body_a), int(body_b), ground_a[0], ground_a[1], ground_b[0], ground_b[1],
                                      anchor_a[0], anchor_a[1], anchor_b[0], anchor_b[1], ratio)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def add_gear(self, joint1, joint2, ratio=1.0):
        """GetGearJointConf(world, joint1, joint2, ratio): joint1 / joint2 are revolute or prismatic joints"""
        self.lib.ref_add_gear.argtypes = [vp, u32, u32, f32]
        self.lib.ref_add_gear.restype = u32
        jid = self.lib.ref_add_gear(self.h, int(joint1), int(joint2), ratio)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def gears(self):
        """the whole values of the gear joints by joint id (zeroed with body_a == 0xFFFFFFFF for other joints)"""
        out = np.zeros(max(1, int(self.counts()["joints"])), abi.GEAR)
        self.lib.ref_export_gear_joints.argtypes = [vp, vp, u32]
        k = self.lib.ref_export_gear_joints(self.h, abi.ptr(out), len(out))
        assert k >= 0
        return out[:k]

    def add_friction(self, body_a, body_b, anchor, max_force, max_torque):
        self.lib.ref_add_friction.argtypes = [vp, u32, u32, f32, f32, f32, f32]
        self.lib.ref_add_friction.restype = u32
        jid = self.lib.ref_add_friction(self.h, int(body_a), int(body_b), anchor[0], anchor[1], max_force, max_torque)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def add_motor(self, body_a, body_b, linear_offset, angular_offset, max_force, max_torque, correction_factor=0.3):
        self.lib.ref_add_motor.argtypes = [vp, u32, u32, f32, f32, f32, f32, f32, f32]
        self.lib.ref_add_motor.restype = u32
        jid = self.lib.ref_add_motor(self.h, int(body_a), int(body_b), linear_offset[0], linear_offset[1], angular_offset, max_force, max_torque,
                                     correction_factor)
        assert jid != 0xFFFFFFFF, self.lib.ref_last_error(self.h)
        return jid

    def destroy_joint(self, joint):
        self.lib.ref_destroy_joint.argtypes = [vp, u32]
        self._chk(self.lib.ref_destroy_joint(self.h, int(joint)))

    # ---- structural mutators through the reference's own API
    def destroy_body(self, body):
        self.lib.ref_destroy_body.argtypes = [vp, u32]
        self._chk(self.lib.ref_destroy_body(self.h, int(body)))

    def destroy_shape(self, shape):
        self.lib.ref_destroy_shape.argtypes = [vp, u32]
        self._chk(self.lib.ref_destroy_shape(self.h, int(shape)))

    def detach_shape(self, body, shape):
        self.lib.ref_detach_shape.argtypes = [vp, u32, u32]
        self._chk(self.lib.ref_detach_shape(self.h, int(body), int(shape)))

    def set_body_enabled(self, body, flag):
        self.lib.ref_set_body_enabled.argtypes = [vp, u32, i32]
        self._chk(self.lib.ref_set_body_enabled(self.h, int(body), 1 if flag else 0))

    def set_body_type(self, body, btype):
        self.lib.ref_set_body_type.argtypes = [vp, u32, i32]
        self._chk(self.lib.ref_set_body_type(self.h, int(body), int(btype)))

    def set_shape_filter(self, shape, cat, mask, group):
        self.lib.ref_set_shape_filter.argtypes = [vp, u32, u32, u32, i32]
        self._chk(self.lib.ref_set_shape_filter(self.h, int(shape), int(cat), int(mask), int(group)))

    def translate_shape(self, shape, dx, dy):
        self.lib.ref_translate_shape.argtypes = [vp, u32, f32, f32]
        self._chk(self.lib.ref_translate_shape(self.h, int(shape), dx, dy))

    def export_shape(self, shape):
        """(row, children, vertices, normals) of one shape, as pr_set_shape takes them."""
        self.lib.ref_export_shape.argtypes = [vp, u32, vp, vp, u32, vp, vp, u32]
        row = np.zeros(1, abi.SHAPE); ch = np.zeros(64, abi.CHILD); vt = np.zeros((4096, 2), np.float32); nm = np.zeros((4096, 2), np.float32)
        nv = self._chk(self.lib.ref_export_shape(self.h, int(shape), abi.ptr(row), abi.ptr(ch), 64, abi.ptr(vt), abi.ptr(nm), 4096))
        nc = int(row["child_count"][0])
        return row, ch[:nc].copy(), vt[:nv].copy(), nm[:nv].copy()

    def set_joints(self, ids, rows):
        """SetJoint per revolute joint (motor and limit fields of the row)."""
        self.lib.ref_set_revolute_joint.argtypes = [vp, u32, vp]
        for i, r in zip(ids, np.ascontiguousarray(rows, dtype=abi.REVOLUTE)):
            one = np.array([r], abi.REVOLUTE)
            self._chk(self.lib.ref_set_revolute_joint(self.h, int(i), abi.ptr(one)))

    def set_contacts(self, rows):
        """SetContact per row (friction, restitution, tangent speed, enabled / needs-updating flags); raises with the
        reference's exception text."""
        self.lib.ref_set_contact.argtypes = [vp, vp]
        for r in np.ascontiguousarray(rows, dtype=abi.CONTACT):
            one = np.array([r], abi.CONTACT)
            self._chk(self.lib.ref_set_contact(self.h, abi.ptr(one)))

    def set_manifold_impulses(self, ids, rows):
        self.lib.ref_set_manifold_impulses.argtypes = [vp, u32, vp]
        for i, r in zip(ids, np.ascontiguousarray(rows, dtype=abi.MANIFOLD)):
            one = np.array([r], abi.MANIFOLD)
            self._chk(self.lib.ref_set_manifold_impulses(self.h, int(i), abi.ptr(one)))

    def ray_cast(self, rays, cap=1 << 16):
        """The reference's RayCast(world, input, callback -> ResetRay) per ray: abi.RAY_HIT rows in callback order."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 5)
        self.lib.ref_ray_cast.argtypes = [vp, u32, vp, u32, vp, vp]
        out = np.zeros(cap, abi.RAY_HIT)
        k = C.c_uint32(0)
        self._chk(self.lib.ref_ray_cast(self.h, len(rays), abi.ptr(rays), cap, abi.ptr(out), C.byref(k)))
        assert k.value <= cap
        return out[:k.value]

    def shift_origin(self, x, y):
        self.lib.ref_shift_origin.argtypes = [vp, f32, f32]
        self._chk(self.lib.ref_shift_origin(self.h, x, y))

    def set_event_mask(self, mask):
        """Installs the reference's begin/end/post-solve listeners (recorded in call order)."""
        self.lib.ref_set_event_mask.argtypes = [vp, u32]
        self._chk(self.lib.ref_set_event_mask(self.h, mask))

    def events(self, kind, cap=1 << 18):
        self.lib.ref_events.argtypes = [vp, u32, vp, u32, vp]
        out = np.zeros(cap, abi.IMPULSE_EVENT if kind == abi.EVENT_POST_SOLVE else abi.PRESOLVE_EVENT if kind == abi.EVENT_PRE_SOLVE else abi.CONTACT_EVENT)
        k = C.c_uint32(0)
        self._chk(self.lib.ref_events(self.h, kind, abi.ptr(out), cap, C.byref(k)))
        assert k.value <= cap
        return out[:k.value]

    def counts(self):
        c = abi.one(abi.COUNTS)
        self._chk(self.lib.ref_get_counts(self.h, abi.ptr(c)))
        return c[0]

    def export_scene(self):
        """Flattened shapes + bodies + body->shape CSR, the input of pr_create_* / orc_create_*."""
        n = self.counts()
        shapes = np.zeros(n["shapes"], abi.SHAPE)
        children = np.zeros(n["children"], abi.CHILD)
        verts = np.zeros(max(1, n["vertices"]), abi.VEC2)
        normals = np.zeros(max(1, n["vertices"]), abi.VEC2)
        self._chk(self.lib.ref_export_shapes(self.h, abi.ptr(shapes), abi.ptr(children), abi.ptr(verts), abi.ptr(normals)))
        bodies = self.bodies()
        offs = np.zeros(n["bodies"] + 1, np.uint32)
        cap = max(1, int(n["bodies"]) * 4 + 16)
        ids = np.zeros(cap, np.uint32)
        k = self.lib.ref_export_body_shapes(self.h, abi.ptr(offs), abi.ptr(ids), cap)
        assert 0 <= k <= cap
        return dict(shapes=shapes, children=children, verts=verts[:n["vertices"]], normals=normals[:n["vertices"]],
                    bodies=bodies, shape_offsets=offs, shape_ids=ids[:k], joints=self.joints(), gears=self.gears())

    def joints(self):
        out = np.zeros(64, abi.REVOLUTE)
        k = self.lib.ref_export_revolute_joints(self.h, abi.ptr(out), len(out))
        if k > len(out):  # (the call returns the joint range: a bigger table is read again)
            out = np.zeros(k, abi.REVOLUTE)
            k = self.lib.ref_export_revolute_joints(self.h, abi.ptr(out), len(out))
        return out[:max(0, k)].copy()

    def bodies(self):
        n = self.counts()
        bodies = np.zeros(n["bodies"], abi.BODY)
        self._chk(self.lib.ref_export_bodies(self.h, abi.ptr(bodies)))
        return bodies

    def proxies(self):
        n = self.counts()
        out = np.zeros(max(1, n["proxies"]), abi.PROXY)
        k = self.lib.ref_export_proxies(self.h, abi.ptr(out), len(out))
        return out[:k]

    def query_aabb(self, box):
        """The reference's Query(GetTree(world), AABB, callback): (body, shape, child) of every leaf it reports."""
        cap = max(16, int(self.counts()["proxies"]))
        b, s, c = (np.zeros(cap, np.uint32) for _ in range(3))
        self.lib.ref_query_aabb.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        k = self.lib.ref_query_aabb(self.h, float(box[0]), float(box[1]), float(box[2]), float(box[3]), abi.ptr(b), abi.ptr(s), abi.ptr(c), cap)
        assert 0 <= k <= cap
        return b[:k], s[:k], c[:k]

    def contacts(self):
        n = self.counts()
        cs = np.zeros(max(1, n["contacts"]), abi.CONTACT)
        ms = np.zeros(max(1, n["contacts"]), abi.MANIFOLD)
        k = self.lib.ref_export_contacts(self.h, abi.ptr(cs), abi.ptr(ms), len(cs))
        return cs[:k], ms[:k]

    def inv_dt0(self):
        return self.lib.ref_get_inv_dt0(self.h)


class OracleWorld:
    """The C++ restatement (oracle/restate)."""

    def __init__(self, scene=None):
        self.lib = _orc_lib()
        self.h = self.lib.orc_create()
        if scene is not None:
            self.create(scene)

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def create(self, sc):
        self.lib.orc_create_shapes(self.h, len(sc["shapes"]), abi.ptr(sc["shapes"]), abi.ptr(sc["children"]),
                                   len(sc["children"]), abi.ptr(sc["verts"]), abi.ptr(sc["normals"]), len(sc["verts"]))
        self.lib.orc_create_bodies(self.h, len(sc["bodies"]), abi.ptr(sc["bodies"]), abi.ptr(sc["shape_offsets"]),
                                   abi.ptr(sc["shape_ids"]))
        if "joints" in sc and len(sc["joints"]):
            for a, b, gear in abi.joint_batches(sc["joints"]):
                if gear:
                    g = np.ascontiguousarray(sc["gears"][a:b], dtype=abi.GEAR)
                    self.lib.orc_create_gear_joints.argtypes = [vp, u32, vp]
                    self.lib.orc_create_gear_joints(self.h, 1, abi.ptr(g))
                else:
                    rows = np.ascontiguousarray(sc["joints"][a:b])
                    self.lib.orc_create_revolute_joints(self.h, len(rows), abi.ptr(rows))

    def step(self, conf, order=None):
        st = abi.one(abi.STEP_STATS)
        if order is not None:
            order = np.ascontiguousarray(order, dtype=np.uint32)
        r = self.lib.orc_step(self.h, abi.ptr(conf), abi.ptr(st), abi.ptr(order) if order is not None else None,
                              0 if order is None else len(order))
        if r != 0:
            raise RuntimeError(self.lib.orc_last_error(self.h).decode())
        return st[0]

    def set_toi_mode(self, mode):
        self.lib.orc_set_toi_mode.argtypes = [vp, u32]
        assert self.lib.orc_set_toi_mode(self.h, int(mode)) == 0

    def toi_schedule(self):
        self.lib.orc_get_toi_schedule.argtypes = [vp, u32, vp, C.POINTER(u32)]
        k = u32(0)
        self.lib.orc_get_toi_schedule(self.h, 0, None, C.byref(k))
        rows = np.zeros((max(1, k.value), 2), np.uint32)
        self.lib.orc_get_toi_schedule(self.h, len(rows), abi.ptr(rows), C.byref(k))
        rows = rows[:k.value]
        return rows[np.lexsort((rows[:, 1], rows[:, 0]))] if len(rows) else rows

    def set_shape_material(self, shape, friction, restitution, density, sensor):
        self.lib.orc_set_shape_material.argtypes = [vp, u32, f32, f32, f32, u32]
        self.lib.orc_set_shape_material(self.h, int(shape), friction, restitution, density, int(sensor))

    def attach_shapes(self, body, shape_ids):
        ids = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        self.lib.orc_attach_shapes.argtypes = [vp, u32, u32, vp]
        self.lib.orc_attach_shapes(self.h, int(body), len(ids), abi.ptr(ids))

    def set_bodies(self, ids, rows):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=abi.BODY)
        self.lib.orc_set_bodies(self.h, len(ids), abi.ptr(ids), abi.ptr(rows))

    def destroy_joints(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self.lib.orc_destroy_joints.argtypes = [vp, u32, vp]
        self.lib.orc_destroy_joints(self.h, len(ids), abi.ptr(ids))

    def add_body(self, row, shape_ids=()):
        row = np.ascontiguousarray(row, dtype=abi.BODY).reshape(1)
        off = np.array([0, len(shape_ids)], np.uint32)
        ids = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        self.lib.orc_create_bodies(self.h, 1, abi.ptr(row), abi.ptr(off), abi.ptr(ids))

    def add_shape(self, row, children, vertices, normals):
        row = np.ascontiguousarray(row, dtype=abi.SHAPE).reshape(1).copy()
        children = np.ascontiguousarray(children, dtype=abi.CHILD)
        vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 2)
        normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 2)
        row["child_offset"] = 0; row["child_count"] = len(children)
        self.lib.orc_create_shapes(self.h, 1, abi.ptr(row), abi.ptr(children), len(children), abi.ptr(vertices), abi.ptr(normals), len(vertices))

    def destroy_bodies(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self.lib.orc_destroy_bodies.argtypes = [vp, u32, vp]
        assert self.lib.orc_destroy_bodies(self.h, len(ids), abi.ptr(ids)) == 0

    def detach_shapes(self, body, shape_ids):
        ids = np.ascontiguousarray(shape_ids, dtype=np.uint32)
        self.lib.orc_detach_shapes.argtypes = [vp, u32, u32, vp]
        self.lib.orc_detach_shapes(self.h, int(body), len(ids), abi.ptr(ids))

    def destroy_shapes(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self.lib.orc_destroy_shapes.argtypes = [vp, u32, vp]
        assert self.lib.orc_destroy_shapes(self.h, len(ids), abi.ptr(ids)) == 0

    def set_body_enabled(self, body, flag):
        self.lib.orc_set_body_enabled.argtypes = [vp, u32, u32]
        self.lib.orc_set_body_enabled(self.h, int(body), 1 if flag else 0)

    def set_body_type(self, body, row):
        row = np.ascontiguousarray(row, dtype=abi.BODY).reshape(1)
        self.lib.orc_set_body_type.argtypes = [vp, u32, vp]
        self.lib.orc_set_body_type(self.h, int(body), abi.ptr(row))

    def set_shape(self, shape, row, children=None, vertices=None, normals=None):
        self.lib.orc_set_shape.argtypes = [vp, u32, vp, vp, vp, vp, u32, u32]
        row = np.ascontiguousarray(row, dtype=abi.SHAPE).reshape(1).copy()
        if children is not None:
            children = np.ascontiguousarray(children, dtype=abi.CHILD)
            vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 2)
            normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 2)
            row["child_count"] = len(children)
            self.lib.orc_set_shape(self.h, int(shape), abi.ptr(row), abi.ptr(children), abi.ptr(vertices), abi.ptr(normals), len(vertices), 1)
        else:
            self.lib.orc_set_shape(self.h, int(shape), abi.ptr(row), None, None, None, 0, 0)

    def create_joints(self, rows):
        rows = np.ascontiguousarray(rows, dtype=abi.REVOLUTE)
        self.lib.orc_create_revolute_joints(self.h, len(rows), abi.ptr(rows))

    def set_joints(self, ids, rows):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=abi.REVOLUTE)
        self.lib.orc_set_revolute_joints.argtypes = [vp, u32, vp, vp]
        r = self.lib.orc_set_revolute_joints(self.h, len(ids), abi.ptr(ids), abi.ptr(rows))
        if r != 0:
            raise RuntimeError(self.lib.orc_last_error(self.h).decode())

    def set_contacts(self, rows):
        rows = np.ascontiguousarray(rows, dtype=abi.CONTACT)
        self.lib.orc_set_contacts.argtypes = [vp, u32, vp]
        r = self.lib.orc_set_contacts(self.h, len(rows), abi.ptr(rows))
        if r != 0:
            raise RuntimeError(self.lib.orc_last_error(self.h).decode())

    def set_manifolds(self, ids, rows):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=abi.MANIFOLD)
        self.lib.orc_set_manifolds.argtypes = [vp, u32, vp, vp]
        r = self.lib.orc_set_manifolds(self.h, len(ids), abi.ptr(ids), abi.ptr(rows))
        if r != 0:
            raise RuntimeError(self.lib.orc_last_error(self.h).decode())

    def ray_cast(self, rays, cap=1 << 16):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 5)
        self.lib.orc_ray_cast.argtypes = [vp, u32, vp, u32, vp, vp]
        out = np.zeros(cap, abi.RAY_HIT)
        k = C.c_uint32(0)
        self.lib.orc_ray_cast(self.h, len(rays), abi.ptr(rays), cap, abi.ptr(out), C.byref(k))
        assert k.value <= cap
        return out[:k.value]

    def shift_origin(self, x, y):
        self.lib.orc_shift_origin.argtypes = [vp, f32, f32]
        self.lib.orc_shift_origin(self.h, x, y)

    def set_event_mask(self, mask):
        self.lib.orc_set_event_mask.argtypes = [vp, u32]
        self.lib.orc_set_event_mask(self.h, mask)

    def events(self, kind, cap=1 << 18):
        """Listener calls of the last step in the reference's call order."""
        self.lib.orc_events.argtypes = [vp, u32, vp, u32, vp]
        out = np.zeros(cap, abi.IMPULSE_EVENT if kind == abi.EVENT_POST_SOLVE else abi.PRESOLVE_EVENT if kind == abi.EVENT_PRE_SOLVE else abi.CONTACT_EVENT)
        k = C.c_uint32(0)
        self.lib.orc_events(self.h, kind, abi.ptr(out), cap, C.byref(k))
        assert k.value <= cap
        return out[:k.value]

    def counts(self):
        c = abi.one(abi.COUNTS)
        self.lib.orc_get_counts(self.h, abi.ptr(c))
        return c[0]

    def bodies(self):
        n = self.counts()["bodies"]
        out = np.zeros(n, abi.BODY)
        self.lib.orc_get_bodies(self.h, 0, n, abi.ptr(out))
        return out

    def contacts(self):
        n = max(1, int(self.counts()["contacts"]))
        cs, ms = np.zeros(n, abi.CONTACT), np.zeros(n, abi.MANIFOLD)
        k = C.c_uint32(0)
        self.lib.orc_get_contacts(self.h, n, abi.ptr(cs), abi.ptr(ms), C.byref(k))
        return cs[:k.value], ms[:k.value]

    def proxies(self):
        n = max(1, int(self.counts()["proxies"]))
        out = np.zeros(n, abi.PROXY)
        k = C.c_uint32(0)
        self.lib.orc_get_proxies(self.h, n, abi.ptr(out), C.byref(k))
        return out[:k.value]

    def joints(self):
        n = int(self.counts()["joints"])
        out = np.zeros(max(1, n), abi.REVOLUTE)
        self.lib.orc_get_revolute_joints(self.h, 0, n, abi.ptr(out))
        return out[:n]

    def load_state(self, bodies, proxies, contacts, manifolds, inv_dt0):
        r = self.lib.orc_load_state(self.h, len(bodies), abi.ptr(bodies), len(proxies), abi.ptr(proxies),
                                    len(contacts), abi.ptr(contacts), abi.ptr(manifolds), inv_dt0)
        if r != 0:
            raise RuntimeError(self.lib.orc_last_error(self.h).decode())

    def inv_dt0(self):
        return self.lib.orc_get_inv_dt0(self.h)


def bits_equal(a, b):
    """Bitwise equality of two structured arrays (NaN-safe, -0 != +0)."""
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def diff_fields(a, b, limit=5):
    """Human-readable list of differing (row, field) for structured arrays."""
    out = []
    if a.shape != b.shape:
        return ["shape %s vs %s" % (a.shape, b.shape)]
    if a.dtype.names is None:
        ai = a.view(np.uint32) if a.dtype.itemsize == 4 else a
        bi = b.view(np.uint32) if b.dtype.itemsize == 4 else b
        bad = np.nonzero(ai != bi)[0]
        return ["row %d: %r vs %r" % (r, a[r], b[r]) for r in bad[:limit]] + (["... %d rows differ" % len(bad)] if len(bad) > limit else [])
    for name in a.dtype.names:
        av, bv = a[name], b[name]
        if av.dtype.names:
            out += ["%s.%s" % (name, s) for s in diff_fields(av, bv, limit)]
            continue
        ai = av.view(np.uint32) if av.dtype.itemsize == 4 else av
        bi = bv.view(np.uint32) if bv.dtype.itemsize == 4 else bv
        bad = np.nonzero(ai != bi)[0]
        for r in bad[:limit]:
            out.append("row %d %s: %r vs %r" % (r, name, av[r], bv[r]))
        if len(bad) > limit:
            out.append("... %d rows differ in %s" % (len(bad), name))
    return out
