"""numpy / ctypes mirror of the POD layouts in include/playrho_b200.h.

Only layouts and helpers live here (no compute). Field order and sizes must match the header
exactly; tests/test_abi.py checks the sizes against the compiled library (pr_abi_sizeof).
"""
import ctypes as C
import numpy as np

f4, u4, u2, i2, i4 = np.float32, np.uint32, np.uint16, np.int16, np.int32

VEC2 = np.dtype([("x", f4), ("y", f4)])

BODY = np.dtype([(n, f4) for n in (
    "xf_px", "xf_py", "xf_qc", "xf_qs", "pos0_x", "pos0_y", "pos0_a", "pos1_x", "pos1_y", "pos1_a",
    "lc_x", "lc_y", "alpha0", "vel_x", "vel_y", "vel_w", "acc_x", "acc_y", "acc_w", "inv_mass",
    "inv_rot_i", "lin_damp", "ang_damp", "under_active_time")] + [("flags", u4)])

CHILD = np.dtype([("vertex_offset", u4), ("vertex_count", u4), ("vertex_radius", f4)])

SHAPE = np.dtype([("friction", f4), ("restitution", f4), ("density", f4), ("filter_category", u2),
                  ("filter_mask", u2), ("filter_group", i2), ("is_sensor", u2), ("child_offset", u4),
                  ("child_count", u4)])

REVOLUTE = np.dtype([("body_a", u4), ("body_b", u4), ("collide_connected", u4),
                     ("local_anchor_a", VEC2), ("local_anchor_b", VEC2),
                     ("impulse_x", f4), ("impulse_y", f4), ("impulse_z", f4),
                     ("angular_motor_impulse", f4), ("reference_angle", f4),
                     ("enable_limit", u4), ("enable_motor", u4), ("lower_angle", f4),
                     ("upper_angle", f4), ("motor_speed", f4), ("max_motor_torque", f4),
                     ("limit_state", u4)])

GEAR_SIDE = np.dtype([("type", u4), ("local_anchor_a", VEC2), ("local_anchor_b", VEC2), ("local_axis", VEC2), ("reference_angle", f4)])
GEAR = np.dtype([("body_a", u4), ("body_b", u4), ("collide_connected", u4), ("body_c", u4), ("body_d", u4),
                 ("impulse", f4), ("ratio", f4), ("constant", f4), ("ac", GEAR_SIDE), ("bd", GEAR_SIDE)])
JOINT_GEAR = 10
JOINT_WORK = np.dtype([("r_a", f4, 2), ("r_b", f4, 2), ("m", f4, 9), ("angular_mass", f4), ("island", u4), ("reserved", u4), ("ex", f4, 8)])


def joint_type(row):
    return (int(row["collide_connected"]) >> 8) & 0xFF


def joint_batches(rows):
    """(start, end, is_gear) runs of a joint table in id order: gear joints one by one (their value is wider than a row), the others in runs"""
    i, n = 0, len(rows)
    while i < n:
        if joint_type(rows[i]) == JOINT_GEAR:
            yield i, i + 1, True
            i += 1
        else:
            j = i
            while j < n and joint_type(rows[j]) != JOINT_GEAR:
                j += 1
            yield i, j, False
            i = j


MANIFOLD = np.dtype([(n, f4) for n in ("ln_x", "ln_y", "lp_x", "lp_y", "p0_x", "p0_y", "p1_x", "p1_y",
                                        "jn0", "jt0", "jn1", "jt1")] +
                    [("cf0", u4), ("cf1", u4), ("type", u4), ("count", u4)])

CONTACT = np.dtype([("id", u4), ("key_min", u4), ("key_max", u4), ("body_a", u4), ("shape_a", u4),
                    ("child_a", u4), ("body_b", u4), ("shape_b", u4), ("child_b", u4),
                    ("friction", f4), ("restitution", f4), ("tangent_speed", f4), ("toi", f4),
                    ("toi_count", u4), ("flags", u4)])

PROXY = np.dtype([("tree_id", u4), ("body", u4), ("shape", u4), ("child", u4), ("min_x", f4),
                  ("min_y", f4), ("max_x", f4), ("max_y", f4)])

STEP_CONF = np.dtype([(n, f4) for n in (
    "delta_time", "dt_ratio", "min_still_time_to_sleep", "linear_slop", "angular_slop",
    "reg_resolution_rate", "reg_min_separation", "reg_min_momentum", "toi_resolution_rate",
    "toi_min_separation", "toi_min_momentum", "target_depth", "tolerance", "velocity_threshold",
    "max_translation", "max_rotation", "max_linear_correction", "max_angular_correction",
    "linear_sleep_tolerance", "angular_sleep_tolerance", "displace_multiplier", "aabb_extension",
    "max_circles_ratio")] + [(n, u4) for n in (
    "reg_velocity_iters", "reg_position_iters", "toi_velocity_iters", "toi_position_iters",
    "max_toi_root_iters", "max_toi_iters", "max_distance_iters", "max_sub_steps", "do_warm_start",
    "do_toi", "do_blocksolve")])

_PRE = np.dtype([(n, u4) for n in ("proxies_created", "proxies_moved", "contacts_destroyed",
                                   "contacts_added", "contacts_updated", "contacts_skipped")])
_REG = np.dtype([("min_separation", f4), ("max_inc_impulse", f4)] + [(n, u4) for n in (
    "islands_found", "islands_solved", "contacts_added", "bodies_slept", "proxies_moved",
    "sum_pos_iters", "sum_vel_iters", "contacts_updated", "contacts_skipped", "max_island_bodies")])
_TOI = np.dtype([("min_separation", f4), ("max_inc_impulse", f4)] + [(n, u4) for n in (
    "islands_found", "islands_solved", "contacts_found", "contacts_at_max_sub_steps",
    "contacts_updated_toi", "contacts_updated_touching", "contacts_skipped_touching",
    "contacts_added", "proxies_moved", "sum_pos_iters", "sum_vel_iters", "max_dist_iters",
    "max_toi_iters", "max_root_iters")])
STEP_STATS = np.dtype([("pre", _PRE), ("reg", _REG), ("toi", _TOI)])

WORLD_CONF = np.dtype([("min_vertex_radius", f4), ("max_vertex_radius", f4), ("body_capacity", u4),
                       ("proxy_capacity", u4), ("contact_capacity", u4), ("vertex_capacity", u4),
                       ("device", i4)])

COUNTS = np.dtype([(n, u4) for n in ("bodies", "shapes", "children", "vertices", "proxies", "contacts",
                                     "joints", "contact_id_range", "tree_node_capacity",
                                     "tree_leaf_count")])

N_STAGES = 16
STEP_PROFILE = np.dtype([(n, u4) for n in ("n_awake", "p_total", "p_moved", "c_total", "c_dirty",
                                           "c_touching", "n_colors", "n_islands", "vel_iters",
                                           "pos_iters", "candidate_pairs", "kernel_launches", "toi_rounds", "toi_events", "color_rounds", "grow_events")] +
                        [("ms_total", f4), ("ms_stage", f4, (N_STAGES,))])
STAGE_NAMES = ["pre", "islands", "integrate", "setup", "color", "velocity", "position", "writeback",
               "sync", "narrow", "bvh", "query", "add", "toi", "toi_solve"]

# body / contact flag bits (include/playrho_b200.h)
BODY_DESTROYED, BODY_AWAKE, BODY_AUTOSLEEP, BODY_IMPENETRABLE = 0x1, 0x2, 0x4, 0x8
BODY_FIXED_ROTATION, BODY_ENABLED, BODY_VELOCITY, BODY_ACCELERATION = 0x10, 0x20, 0x80, 0x100
BODY_MASS_DIRTY = 0x200
CONTACT_TOUCHING, CONTACT_ENABLED, CONTACT_FILTER, CONTACT_TOI = 0x1, 0x2, 0x4, 0x8
CONTACT_DIRTY, CONTACT_SENSOR, CONTACT_DESTROYED, CONTACT_IMPENETRABLE = 0x10, 0x20, 0x40, 0x80
MANIFOLD_UNSET, MANIFOLD_CIRCLES, MANIFOLD_FACE_A, MANIFOLD_FACE_B = 0, 1, 2, 3


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def one(dt):
    return np.zeros(1, dtype=dt)


def cf_unpack(cf):
    return (cf & 0xFF, (cf >> 8) & 0xFF, (cf >> 16) & 0xFF, (cf >> 24) & 0xFF)

# pr_contact_event / pr_impulse_event (pr_events)
EVENT_BEGIN, EVENT_END, EVENT_PRE_SOLVE, EVENT_POST_SOLVE = 1, 2, 4, 8
CONTACT_EVENT = np.dtype([(n, u4) for n in ("kind", "contact", "body_a", "shape_a", "child_a", "body_b", "shape_b", "child_b")])
IMPULSE_EVENT = np.dtype([("contact", u4), ("count", u4), ("iterations", u4), ("phase", u4), ("normal", f4, (2,)), ("tangent", f4, (2,))])
PRESOLVE_EVENT = np.dtype([("contact", u4), ("pad", u4), ("old_manifold", MANIFOLD)])

# pr_ray_hit (pr_ray_cast)
RAY_ALL, RAY_CLOSEST = 0, 1
RAY_HIT = np.dtype([("ray", u4), ("body", u4), ("shape", u4), ("child", u4), ("fraction", f4), ("normal_x", f4), ("normal_y", f4),
                    ("point_x", f4), ("point_y", f4)])
