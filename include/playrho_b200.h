/*
 * playrho_b200.h -- C ABI of the B200-native, device-resident PlayRho world ("CudaWorld").
 *
 * This is the drop-in boundary for the per-step collision-and-solve pipeline
 *   Step(AabbTreeWorld&, const StepConf&)          Library/source/playrho/d2/AabbTreeWorld.cpp:2144
 * A C++17 host type (playrho_b200/host/CudaWorld.hpp) implements the 52 ADL free functions that
 * playrho::d2::detail::WorldModel<T> (Library/include/playrho/d2/detail/WorldModel.hpp:36-370)
 * forwards to, and calls ONLY the functions below. Plain pointers and sizes, no C++/torch types,
 * never throws; every function returns PR_OK (0) or a negative pr_status; pr_last_error() gives
 * the message. A handle is single-threaded like the reference world (AabbTreeWorld.hpp:866-875).
 *
 * All reals are IEEE binary32 (playrho::Real = float, Library/include/playrho/Real.hpp:69).
 * Ids are 32-bit here; the 16-bit stock ids of the reference are a host-shim concern.
 */
#ifndef PLAYRHO_B200_H
#define PLAYRHO_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PR_ABI_VERSION 1
#define PR_INVALID_ID 0xFFFFFFFFu
#define PR_INVALID_VERTEX 255u /* playrho::InvalidVertex, Settings.hpp:87-90 */

typedef enum pr_status {
    PR_OK = 0,
    PR_ERR_INVALID_ARGUMENT = -1, /* playrho::InvalidArgument */
    PR_ERR_OUT_OF_RANGE = -2,     /* playrho::OutOfRange      */
    PR_ERR_WRONG_STATE = -3,      /* playrho::WrongState (world locked) */
    PR_ERR_LENGTH = -4,           /* playrho::LengthError (capacity) */
    PR_ERR_CUDA = -5,             /* CUDA runtime failure; message has cudaGetErrorString */
    PR_ERR_UNSUPPORTED = -6       /* e.g. joint type the device solver does not know */
} pr_status;

typedef struct pr_vec2 { float x, y; } pr_vec2;

/* --- body flags: bit-identical to Body::Flag, Library/include/playrho/d2/Body.hpp:440-474 --- */
enum {
    PR_BODY_DESTROYED = 0x0001, PR_BODY_AWAKE = 0x0002, PR_BODY_AUTOSLEEP = 0x0004,
    PR_BODY_IMPENETRABLE = 0x0008, PR_BODY_FIXED_ROTATION = 0x0010, PR_BODY_ENABLED = 0x0020,
    PR_BODY_VELOCITY = 0x0080 /* speedable */, PR_BODY_ACCELERATION = 0x0100 /* accelerable */,
    PR_BODY_MASS_DIRTY = 0x0200
};
/* --- contact flags: bit-identical to Contact::Flag, Library/include/playrho/Contact.hpp:285-310 --- */
enum {
    PR_CONTACT_TOUCHING = 0x01, PR_CONTACT_ENABLED = 0x02, PR_CONTACT_FILTER = 0x04,
    PR_CONTACT_TOI = 0x08, PR_CONTACT_DIRTY = 0x10, PR_CONTACT_SENSOR = 0x20,
    PR_CONTACT_DESTROYED = 0x40, PR_CONTACT_IMPENETRABLE = 0x80
};
/* Manifold::Type, Library/include/playrho/d2/Manifold.hpp:96-126 */
enum { PR_MANIFOLD_UNSET = 0, PR_MANIFOLD_CIRCLES = 1, PR_MANIFOLD_FACE_A = 2, PR_MANIFOLD_FACE_B = 3 };
/* ContactFeature::Type, Library/include/playrho/ContactFeature.hpp:44-48 */
enum { PR_CF_VERTEX = 0, PR_CF_FACE = 1 };

/* State of one Body (Body.hpp private block). One row per body id. */
typedef struct pr_body {
    float xf_px, xf_py, xf_qc, xf_qs; /* m_xf: p and q=(cos,sin)               */
    float pos0_x, pos0_y, pos0_a;      /* m_sweep.pos0 (linear, angular)        */
    float pos1_x, pos1_y, pos1_a;      /* m_sweep.pos1                          */
    float lc_x, lc_y;                  /* m_sweep.localCenter                   */
    float alpha0;                      /* m_sweep.alpha0                        */
    float vel_x, vel_y, vel_w;         /* m_linearVelocity, m_angularVelocity   */
    float acc_x, acc_y, acc_w;         /* m_linearAcceleration, m_angularAcc... */
    float inv_mass, inv_rot_i;         /* m_invMass, m_invRotI                  */
    float lin_damp, ang_damp;          /* m_linearDamping, m_angularDamping     */
    float under_active_time;           /* m_underActiveTime                     */
    uint32_t flags;                    /* PR_BODY_*                             */
} pr_body;

/* One convex child of a shape == DistanceProxy{vertexRadius,count,vertices,normals}
 * (Library/include/playrho/d2/DistanceProxy.hpp:51-151). vertex_offset indexes the vertex /
 * normal arrays passed with it (disk: 1 vertex, no normal needed -- pass anything). */
typedef struct pr_child {
    uint32_t vertex_offset;
    uint32_t vertex_count; /* 1..254 */
    float vertex_radius;
} pr_child;

/* Per-shape data the step needs (ShapeConf.hpp:45-51, Filter.hpp). */
typedef struct pr_shape {
    float friction, restitution, density;
    uint16_t filter_category, filter_mask;
    int16_t filter_group;
    uint16_t is_sensor;
    uint32_t child_offset; /* into the pr_child array passed with it */
    uint32_t child_count;
} pr_shape;

/* RevoluteJointConf (Library/include/playrho/d2/RevoluteJointConf.hpp) -- the one joint type the
 * device solver runs. Solver-written fields are readable back. */
typedef struct pr_revolute_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;
    pr_vec2 local_anchor_a, local_anchor_b;
    float impulse_x, impulse_y, impulse_z;
    float angular_motor_impulse;
    float reference_angle;
    uint32_t enable_limit, enable_motor;
    float lower_angle, upper_angle;
    float motor_speed, max_motor_torque;
    uint32_t limit_state;
} pr_revolute_joint;

/* Joint rows share one table (JointIDs are one id space in the reference): every row has pr_revolute_joint's size and leading
 * fields, and bits 8..15 of collide_connected say which joint type it is (bit 0 is the flag itself). */
enum { PR_JOINT_REVOLUTE = 0, PR_JOINT_DISTANCE = 1, PR_JOINT_ROPE = 2, PR_JOINT_FRICTION = 3, PR_JOINT_MOTOR = 4, PR_JOINT_WELD = 5, PR_JOINT_TARGET = 6, PR_JOINT_PRISMATIC = 7, PR_JOINT_WHEEL = 8, PR_JOINT_PULLEY = 9, PR_JOINT_GEAR = 10 };
/* prismatic and wheel rows keep their enable flags in collide_connected (their enable_limit / enable_motor words hold the local axis) */
#define PR_JOINT_ENABLE_LIMIT 0x10000u
#define PR_JOINT_ENABLE_MOTOR 0x20000u
#define PR_JOINT_TYPE(row) (((row).collide_connected >> 8) & 0xFFu)
/* DistanceJointConf (Library/include/playrho/d2/DistanceJointConf.hpp): rod or spring between two anchor points.
 * Solver: Library/source/playrho/d2/DistanceJointConf.cpp:81-246 (InitVelocity, SolveVelocity, SolvePosition). */
typedef struct pr_distance_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_DISTANCE */
    pr_vec2 local_anchor_a, local_anchor_b;
    float impulse;                       /* accumulated impulse (warm start) */
    float reserved0[3];
    float length;                        /* rest length */
    uint32_t reserved1[2];
    float frequency, damping_ratio;      /* frequency 0: rigid rod (position-corrected); > 0: soft spring */
    float reserved2[2];
    uint32_t reserved3;
} pr_distance_joint;

/* RopeJointConf (Library/include/playrho/d2/RopeJointConf.hpp): the two anchor points are never further apart than max_length.
 * Solver: Library/source/playrho/d2/RopeJointConf.cpp:58-214. */
typedef struct pr_rope_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_ROPE */
    pr_vec2 local_anchor_a, local_anchor_b;
    float impulse;                       /* accumulated impulse (warm start), <= 0 */
    float reserved0[3];
    float max_length;
    uint32_t reserved1[2];
    float reserved2[4];
    uint32_t limit_state;                /* LimitState: 0 inactive, 2 at the upper limit (InitVelocity writes it) */
} pr_rope_joint;

/* FrictionJointConf (FrictionJointConf.hpp; solver FrictionJointConf.cpp:76-239): top-down friction, force and torque capped. */
typedef struct pr_friction_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_FRICTION */
    pr_vec2 local_anchor_a, local_anchor_b;
    float linear_impulse_x, linear_impulse_y, angular_impulse; /* accumulated impulses (warm start) */
    float reserved0[2];
    uint32_t reserved1[2];
    float reserved2[2];
    float max_force, max_torque;
    uint32_t reserved3;
} pr_friction_joint;
/* MotorJointConf (MotorJointConf.hpp; solver MotorJointConf.cpp:77-235): drives body B towards an offset pose relative to A. */
typedef struct pr_motor_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_MOTOR */
    pr_vec2 linear_offset;               /* in body A's frame */
    pr_vec2 reserved0;
    float linear_impulse_x, linear_impulse_y, angular_impulse;
    float reserved1;
    float angular_offset;
    uint32_t reserved2[2];
    float correction_factor, reserved3;
    float max_force, max_torque;
    uint32_t reserved4;
} pr_motor_joint;
/* WeldJointConf (WeldJointConf.hpp; solver WeldJointConf.cpp:35-371): glues two bodies together (rigid, or with a soft angular spring). */
typedef struct pr_weld_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_WELD */
    pr_vec2 local_anchor_a, local_anchor_b;
    float impulse_x, impulse_y, impulse_z; /* accumulated impulse (warm start) */
    float reserved0;
    float reference_angle;               /* angle of body B minus angle of body A */
    uint32_t reserved1[2];
    float frequency, damping_ratio;      /* frequency 0: rigid; > 0: the angular constraint is a soft spring */
    float reserved2[2];
    uint32_t reserved3;
} pr_weld_joint;
/* TargetJointConf (TargetJointConf.hpp; solver TargetJointConf.cpp:52-243): drags a point of ONE body towards a world target
 * (the reference's "mouse joint"). The solver only touches body B. When the reference's bodyA is InvalidBodyID the table row's body_a
 * repeats body_b; a valid bodyA only takes part in the island rules (a joint whose other body is disabled is not solved,
 * AabbTreeWorld.cpp:1443-1466). */
typedef struct pr_target_joint {
    uint32_t body_a, body_b;             /* body_b: the dragged body; body_a: body_b again, or the reference's bodyA */
    uint32_t collide_connected;          /* bits 8..15: PR_JOINT_TARGET */
    pr_vec2 target;                      /* world coordinates */
    pr_vec2 local_anchor_b;
    float impulse_x, impulse_y;          /* accumulated impulse (warm start) */
    float reserved0[3];
    uint32_t reserved1[2];
    float frequency, damping_ratio;
    float max_force, reserved2;
    uint32_t reserved3;
} pr_target_joint;
/* PrismaticJointConf (PrismaticJointConf.hpp; solver PrismaticJointConf.cpp:150-533): body B slides along an axis fixed in body A,
 * no relative rotation; optional translation limits and a linear motor. localYAxisA is the reverse perpendicular of local_x_axis_a. */
typedef struct pr_prismatic_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_PRISMATIC; PR_JOINT_ENABLE_LIMIT / _MOTOR */
    pr_vec2 local_anchor_a, local_anchor_b;
    float impulse_x, impulse_y, impulse_z; /* accumulated impulse: perpendicular, angular, limit (warm start) */
    float motor_impulse;
    float reference_angle;
    pr_vec2 local_x_axis_a;              /* unit vector in body A's frame */
    float lower_translation, upper_translation;
    float motor_speed, max_motor_force;
    uint32_t limit_state;                /* LimitState (InitVelocity writes it) */
} pr_prismatic_joint;
/* WheelJointConf (WheelJointConf.hpp; solver WheelJointConf.cpp:103-362): body B's anchor stays on a line fixed in body A, with a
 * spring along the line and a rotational motor -- a vehicle suspension. */
typedef struct pr_wheel_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_WHEEL; PR_JOINT_ENABLE_MOTOR */
    pr_vec2 local_anchor_a, local_anchor_b;
    float impulse, spring_impulse, angular_impulse; /* accumulated impulses (warm start) */
    float reserved0[2];
    pr_vec2 local_x_axis_a;              /* unit vector in body A's frame */
    float frequency, damping_ratio;
    float motor_speed, max_motor_torque;
    uint32_t reserved1;
} pr_wheel_joint;
/* PulleyJointConf (PulleyJointConf.hpp; solver PulleyJointConf.cpp:91-282): length_a + ratio * length_b stays constant, each length
 * measured from a body's anchor to its ground anchor (world points; pr_shift_origin moves them). */
typedef struct pr_pulley_joint {
    uint32_t body_a, body_b;
    uint32_t collide_connected;          /* bit 0: collideConnected (the reference's default for this type: true); bits 8..15: PR_JOINT_PULLEY */
    pr_vec2 local_anchor_a, local_anchor_b;
    float impulse;                       /* accumulated impulse (warm start) */
    pr_vec2 ground_anchor_a, ground_anchor_b;
    float length_a, length_b;            /* rest lengths */
    float ratio;
    float reserved0[3];
    uint32_t reserved1;
} pr_pulley_joint;
/* GearJointConf (GearJointConf.hpp; solver GearJointConf.cpp:171-468): couples the coordinate of one revolute or prismatic joint (between
 * bodies C and A) to that of another (between D and B): coordinate1 + ratio * coordinate2 = constant. The reference copies what it needs
 * from the two geared joints when the conf is made (GetGearJointConf, :91-169); this struct carries those copies. It is wider than a
 * row of the joint table: the table row (pr_get_revolute_joints) holds body_a, body_b, collide_connected, impulse, ratio and constant,
 * pr_get_gear_joints reads the whole value. */
enum { PR_GEAR_NONE = 0, PR_GEAR_PRISMATIC = 1, PR_GEAR_REVOLUTE = 2 };
typedef struct pr_gear_side {
    uint32_t type;                       /* PR_GEAR_*: the geared joint's type */
    pr_vec2 local_anchor_a;              /* prismatic: the geared joint's localAnchorA (on body C / D) */
    pr_vec2 local_anchor_b;              /* prismatic: its localAnchorB (on body A / B) */
    pr_vec2 local_axis;                  /* prismatic: its localXAxisA */
    float reference_angle;               /* revolute: its referenceAngle */
} pr_gear_side;
typedef struct pr_gear_joint {
    uint32_t body_a, body_b;             /* the geared joints' bodyB's */
    uint32_t collide_connected;          /* bit 0: collideConnected; bits 8..15: PR_JOINT_GEAR */
    uint32_t body_c, body_d;             /* the geared joints' bodyA's */
    float impulse;                       /* accumulated impulse (warm start) */
    float ratio, constant;
    pr_gear_side ac, bd;
} pr_gear_joint;
/* What InitVelocity computes for a joint and the reference keeps in the JointConf value ("solver temporaries": rA, rB, effective masses,
 * axes...; GetJoint returns them, e.g. GetAngularMass(world, joint)). One entry per joint id; which word means what for which type is
 * listed at each type's solver in playrho_b200/csrc/pr_joint.cuh ("work:" comments), e.g. revolute: m = the 3x3 mass matrix
 * (m[3 * i + k] = mass[i][k]), angular_mass; wheel: m[0..1] = ay, m[2] = sAy, m[3] = sBy, m[4] = mass, m[5..6] = ax, m[7] = sAx,
 * m[8] = sBx, ex[0] = springMass, ex[1] = bias, ex[2] = gamma, angular_mass. Valid after a Step that solved the joint's island. */
typedef struct pr_joint_work {
    float r_a[2], r_b[2];
    float m[9];
    float angular_mass;
    uint32_t island;                     /* the island the joint was solved with in the last Step, 0xFFFFFFFF: not solved */
    uint32_t reserved;
    float ex[8];
} pr_joint_work;

/* ContactFeature packed little-endian: typeA | indexA<<8 | typeB<<16 | indexB<<24 */
typedef uint32_t pr_contact_feature;

/* Manifold (Library/include/playrho/d2/Manifold.hpp:133-160,411-418): 64 B, one row per contact id */
typedef struct pr_manifold {
    float ln_x, ln_y;         /* m_localNormal (face types)                      */
    float lp_x, lp_y;         /* m_localPoint                                    */
    float p0_x, p0_y, p1_x, p1_y; /* points[i].localPoint                        */
    float jn0, jt0, jn1, jt1; /* points[i].normalImpulse / tangentImpulse        */
    pr_contact_feature cf0, cf1;
    uint32_t type;            /* PR_MANIFOLD_*                                   */
    uint32_t count;           /* GetPointCount()                                 */
} pr_manifold;

/* Contact (Library/include/playrho/Contact.hpp:312-341) + its ContactKey */
typedef struct pr_contact {
    uint32_t id;                            /* ContactID (pool index)            */
    uint32_t key_min, key_max;              /* ContactKey = proxy (tree leaf) ids */
    uint32_t body_a, shape_a, child_a;      /* Contactable A                     */
    uint32_t body_b, shape_b, child_b;      /* Contactable B                     */
    float friction, restitution, tangent_speed;
    float toi;
    uint32_t toi_count;
    uint32_t flags;                         /* PR_CONTACT_*                      */
} pr_contact;

/* Proxy == DynamicTree leaf: Contactable + fat AABB (DynamicTree.hpp TreeNode) */
typedef struct pr_proxy {
    uint32_t tree_id; /* id the reference's DynamicTree would have given this leaf */
    uint32_t body, shape, child;
    float min_x, min_y, max_x, max_y; /* fat AABB */
} pr_proxy;

/* StepConf (Library/include/playrho/StepConf.hpp:163-391), same field meaning and defaults;
 * pr_step_conf_default() fills the reference's defaults. Angles in radians. */
typedef struct pr_step_conf {
    float delta_time, dt_ratio;
    float min_still_time_to_sleep, linear_slop, angular_slop;
    float reg_resolution_rate, reg_min_separation, reg_min_momentum;
    float toi_resolution_rate, toi_min_separation, toi_min_momentum;
    float target_depth, tolerance, velocity_threshold;
    float max_translation, max_rotation, max_linear_correction, max_angular_correction;
    float linear_sleep_tolerance, angular_sleep_tolerance;
    float displace_multiplier, aabb_extension, max_circles_ratio;
    uint32_t reg_velocity_iters, reg_position_iters, toi_velocity_iters, toi_position_iters;
    uint32_t max_toi_root_iters, max_toi_iters, max_distance_iters, max_sub_steps;
    uint32_t do_warm_start, do_toi, do_blocksolve;
} pr_step_conf;

/* StepStats (Library/include/playrho/StepStats.hpp): pre / reg / toi */
typedef struct pr_step_stats {
    struct { uint32_t proxies_created, proxies_moved, contacts_destroyed, contacts_added,
                      contacts_updated, contacts_skipped; } pre;
    struct { float min_separation, max_inc_impulse;
             uint32_t islands_found, islands_solved, contacts_added, bodies_slept, proxies_moved,
                      sum_pos_iters, sum_vel_iters, contacts_updated, contacts_skipped,
                      max_island_bodies; } reg;
    struct { float min_separation, max_inc_impulse;
             uint32_t islands_found, islands_solved, contacts_found, contacts_at_max_sub_steps,
                      contacts_updated_toi, contacts_updated_touching, contacts_skipped_touching,
                      contacts_added, proxies_moved, sum_pos_iters, sum_vel_iters,
                      max_dist_iters, max_toi_iters, max_root_iters; } toi;
} pr_step_stats;

/* WorldConf (Library/include/playrho/d2/WorldConf.hpp) + device arena sizes (0 = grow on demand) */
typedef struct pr_world_conf {
    float min_vertex_radius, max_vertex_radius; /* GetVertexRadiusInterval */
    uint32_t body_capacity, proxy_capacity, contact_capacity, vertex_capacity;
    int32_t device; /* CUDA device ordinal, -1 = current */
} pr_world_conf;

typedef struct pr_counts {
    uint32_t bodies, shapes, children, vertices, proxies, contacts, joints;
    uint32_t contact_id_range; /* GetContactRange(): size of the contact pool incl. free slots */
    uint32_t tree_node_capacity, tree_leaf_count;
} pr_counts;

/* Per-step device-side accounting used for the roofline figure */
typedef struct pr_step_profile {
    uint32_t n_awake, p_total, p_moved, c_total, c_dirty, c_touching, n_colors, n_islands;
    uint32_t vel_iters, pos_iters, candidate_pairs, kernel_launches;
    uint32_t toi_rounds, toi_events; /* TOI phase: kernel rounds run and TOI events solved (SolveToi loop turns) */
    uint32_t color_rounds, grow_events; /* colouring rounds run; device buffers re-allocated (cudaMalloc/cudaFree) during the step */
    float ms_total;            /* CUDA-event time of the whole step on the world's stream */
    float ms_stage[16];        /* see PR_STAGE_* */
} pr_step_profile;
enum { PR_STAGE_PRE = 0, PR_STAGE_ISLANDS, PR_STAGE_INTEGRATE, PR_STAGE_SETUP, PR_STAGE_COLOR,
       PR_STAGE_VELOCITY, PR_STAGE_POSITION, PR_STAGE_WRITEBACK, PR_STAGE_SYNC, PR_STAGE_NARROW,
       PR_STAGE_BVH, PR_STAGE_QUERY, PR_STAGE_ADD, PR_STAGE_TOI, PR_STAGE_TOI_SOLVE, PR_STAGE_COUNT };

typedef struct pr_world pr_world; /* opaque */

/* ---- lifetime (replaces AabbTreeWorld ctor/copy/Clear: AabbTreeWorld.cpp:780-1017) ---- */
int pr_abi_version(void);
/* UnitVec::Get(Angle) (Library/source/playrho/d2/UnitVec.cpp:27-30: cos(angle), sin(angle) -- libm's sincosf / sinf / cosf, one
 * polynomial) gives bits that depend on the ifunc variant glibc's loader picked for the host CPU: generic, or "-fma" on CPUs with
 * usable FMA + AVX2; 34 of the 2^32 inputs differ by one ulp (profiles/r01h_sincosf_exhaustive.txt). Each build of this library
 * restates ONE variant: 0 = generic (libplayrho_b200.so), 1 = fma (libplayrho_b200_fma.so, -DPR_SINCOSF_FMA). The host loads
 * the build that matches the reference's libm on the same machine ("sincosf variant"). */
int pr_sincosf_variant(void);
/* the variant this host's glibc selects (usable FMA + AVX2 -> 1, else 0): what pr_sincosf_variant() of the loaded build should be */
int pr_host_sincosf_variant(void);
void pr_world_conf_default(pr_world_conf* conf);
void pr_step_conf_default(pr_step_conf* conf);          /* StepConf{} defaults */
int pr_create(const pr_world_conf* conf, pr_world** out);
int pr_destroy(pr_world* w);
int pr_clone(const pr_world* w, pr_world** out);
/* Snapshot / restore (the reference's world is a value: copy construction AabbTreeWorld.cpp:842-880, here to and from host
 * memory, e.g. a file). pr_snapshot_save writes the whole device world -- every array of every pool, the counters, the contact
 * pool's free stack and list order, the persistent colours, joints, queued proxies -- into buf (buf == NULL: only the size
 * is reported in *bytes_out; PR_ERR_LENGTH if cap is too small). pr_snapshot_load creates a world from such a buffer on
 * `device` (-1: the current one): it continues bit for bit like the saved world (like a pr_clone). The buffer is only valid
 * for the library version and sincosf build that wrote it. */
int pr_snapshot_save(const pr_world* w, void* buf, uint64_t cap, uint64_t* bytes_out);
int pr_snapshot_load(const void* buf, uint64_t bytes, int device, pr_world** out);        /* WorldModel::Clone_, WorldModel.hpp:96-99 */
int pr_clear(pr_world* w);                              /* Clear(AabbTreeWorld&), AabbTreeWorld.cpp:949 */
const char* pr_last_error(const pr_world* w);           /* w may be NULL: last create error */

/* ---- population (CreateShape :1234, CreateBody :1019, CreateJoint :1132) ----
 * Bulk, flattened: shapes[i].child_offset/children[j].vertex_offset index the arrays given here.
 * Ids are returned in first_id.. first_id+n-1 (pool append order; creation-only pools). */
int pr_create_shapes(pr_world* w, uint32_t n, const pr_shape* shapes, const pr_child* children,
                     uint32_t n_children, const pr_vec2* vertices, const pr_vec2* normals,
                     uint32_t n_vertices, uint32_t* first_id);
/* shape_offsets has n+1 entries into shape_ids (CSR): the ShapeIDs attached to each body.
 * Queues (body,shape) for proxy creation at the next Step like m_fixturesForProxies (:1019-1060). */
int pr_create_bodies(pr_world* w, uint32_t n, const pr_body* bodies, const uint32_t* shape_offsets,
                     const uint32_t* shape_ids, uint32_t* first_id);
int pr_create_revolute_joints(pr_world* w, uint32_t n, const pr_revolute_joint* joints,
                              uint32_t* first_id);
/* CreateJoint for DistanceJointConf values (the type tag is set by the call). Ids come from the same pool as the revolute
 * joints'; pr_get_revolute_joints / pr_set_revolute_joints / pr_destroy_joints work on any row of the table (read a row's
 * type with PR_JOINT_TYPE and view it as the matching struct). */
int pr_create_distance_joints(pr_world* w, uint32_t n, const pr_distance_joint* joints, uint32_t* first_id);
int pr_create_rope_joints(pr_world* w, uint32_t n, const pr_rope_joint* joints, uint32_t* first_id);
int pr_create_friction_joints(pr_world* w, uint32_t n, const pr_friction_joint* joints, uint32_t* first_id);
int pr_create_motor_joints(pr_world* w, uint32_t n, const pr_motor_joint* joints, uint32_t* first_id);
int pr_create_weld_joints(pr_world* w, uint32_t n, const pr_weld_joint* joints, uint32_t* first_id);
int pr_create_target_joints(pr_world* w, uint32_t n, const pr_target_joint* joints, uint32_t* first_id);
int pr_create_prismatic_joints(pr_world* w, uint32_t n, const pr_prismatic_joint* joints, uint32_t* first_id);
int pr_create_wheel_joints(pr_world* w, uint32_t n, const pr_wheel_joint* joints, uint32_t* first_id);
int pr_create_pulley_joints(pr_world* w, uint32_t n, const pr_pulley_joint* joints, uint32_t* first_id);
int pr_create_gear_joints(pr_world* w, uint32_t n, const pr_gear_joint* joints, uint32_t* first_id);
/* the gear joints among joints [first, first + n): out[i] is filled for a gear joint, zeroed with body_a == 0xFFFFFFFF for any other row */
int pr_get_gear_joints(const pr_world* w, uint32_t first, uint32_t n, pr_gear_joint* out);
int pr_get_joint_work(const pr_world* w, uint32_t first, uint32_t n, pr_joint_work* out);

/* SetShape(AabbTreeWorld&, ShapeID, Shape) (AabbTreeWorld.cpp:1284-1356) for a value whose geometry (children) and filter are
 * what they were: friction, restitution, density and the sensor flag are replaced; every body using the shape is woken
 * (:1297-1301); if the sensor flag, friction or restitution changed, every contact of the shape is flagged for updating and
 * its bodies are woken (:1346-1354). Existing contacts keep their mixed friction / restitution and their sensor flag, like
 * the reference's (they are set when a contact is created). Geometry or filter changes: pr_set_shape. */
int pr_set_shape_material(pr_world* w, uint32_t shape, float friction, float restitution, float density, uint32_t is_sensor);
/* Attach(World&, BodyID, ShapeID) = the SetBody whose new value has MORE shapes than the old one (AabbTreeWorld.cpp:2760-2786,
 * GetOldAndNewShapeIDs :647-680): the added shapes are queued for proxy creation by the next Step (m_fixturesForProxies), in
 * the order given (the reference appends them in ShapeID order). The body row itself (mass data after ResetMassData) is
 * uploaded with pr_set_bodies as for any SetBody. Removing shapes: pr_detach_shapes. */
int pr_attach_shapes(pr_world* w, uint32_t body, uint32_t n, const uint32_t* shape_ids);
/* ---- dirty-row upload / download (SetBody :2723 / GetBody; value semantics) ---- */
int pr_set_bodies(pr_world* w, uint32_t n, const uint32_t* ids, const pr_body* rows);
int pr_get_bodies(const pr_world* w, uint32_t first, uint32_t n, pr_body* rows);
/* Column accessors for the per-step loop "apply accelerations -> Step -> read poses" (bulk body
 * get/set instead of one value-semantic SetBody/GetBody per id): a 104-byte row per body each way is 204 MB per step
 * at 1M bodies, these are 16 + 16 (+12) bytes.
 * pr_set_body_accelerations: RAW write of (linear x, linear y, angular) for bodies ids[0..n) (ids == NULL: bodies 0..n-1): what SetBody does with a Body
 * whose only change is the acceleration. Body::SetAcceleration's wake rule (Body.cpp:203-236: awake + underActiveTime = 0 when
 * the acceleration grows or turns) is the host shim's business -- it runs the reference's own Body::SetAcceleration on its
 * mirror and uploads whole rows (pr_set_bodies) for the bodies whose flags changed.
 * pr_get_body_poses: transformation (px, py, cos, sin) of bodies [first, first+n) -- GetTransformation(world, id), Body.hpp --
 * and, if vel3 != NULL, their velocity (linear x, linear y, angular) -- GetVelocity(world, id). */
int pr_set_body_accelerations(pr_world* w, uint32_t n, const uint32_t* ids, const float* acc3);
/* pr_set_body_velocities: Body::SetVelocity (Body.cpp:184-194) of (linear x, linear y, angular) for bodies ids[0..n) (ids == NULL:
 * bodies 0..n-1): a non-zero velocity is ignored for a body that cannot move (no velocity flag) and otherwise wakes the body and
 * resets its under-active time; a zero velocity is just stored. The RL loop's "set actions" call when actions are velocities. */
int pr_set_body_velocities(pr_world* w, uint32_t n, const uint32_t* ids, const float* vel3);
int pr_get_body_poses(const pr_world* w, uint32_t first, uint32_t n, float* xf4, float* vel3);
/* Query(const DynamicTree&, const AABB&, QueryShapeCallback) (DynamicTree.hpp:696-708, DynamicTree.cpp:673-710; SURVEY 8(f)
 * rank 2) for n_boxes boxes at once, on the device LBVH instead of a materialised DynamicTree: every proxy whose FAT AABB
 * overlaps box i (closed intervals, TestOverlap) is reported as (i, body, shape, child) -- the leaves the reference's callback
 * would see, in no particular order. boxes4 = (min x, min y, max x, max y) per box; box_world = world index per box for a batch
 * arena, NULL for a plain world. *n_out = number of hits; if it exceeds cap only the first cap were written. The tree is the one
 * of the last Step (like GetTree): bodies created or moved since are seen after the next Step. */
int pr_query_aabb(const pr_world* w, uint32_t n_boxes, const float* boxes4, const uint32_t* box_world, uint32_t cap,
                  uint32_t* hit_box, uint32_t* hit_body, uint32_t* hit_shape, uint32_t* hit_child, uint32_t* n_out);
/* RayCast(const World&, const RayCastInput&, ShapeRayCastCB) (RayCastOutput.cpp:301-337 over RayCast(const DynamicTree&, ...)
 * :241-299 and RayCast(const DistanceProxy&, ...) :139-233 / the disk case :41-75) for n_rays rays at once on
 * the device LBVH. rays5 = (p1 x, p1 y, p2 x, p2 y, maxFraction) per ray; ray_world as in pr_query_aabb.
 *   PR_RAY_ALL      every shape child the ray hits: what a callback returning RayCastOpcode::ResetRay is called with
 *                   (hits[0..min(*n_out, cap)), no particular order; *n_out > cap: too short, first cap written);
 *   PR_RAY_CLOSEST  per ray the hit with the least fraction: what a callback returning ClipRay last sees. hits[i] belongs to
 *                   ray i (cap >= n_rays), body == 0xFFFFFFFF for a miss; equal fractions: the lower proxy index wins.
 * fraction, normal and point (= p1 + (p2 - p1) * fraction) are the reference's, bit for bit. The tree is the one of the last
 * Step (like GetTree). */
enum { PR_RAY_ALL = 0, PR_RAY_CLOSEST = 1 };
typedef struct pr_ray_hit {
    uint32_t ray, body, shape, child;
    float fraction, normal_x, normal_y, point_x, point_y;
} pr_ray_hit;
int pr_ray_cast(const pr_world* w, uint32_t n_rays, const float* rays5, const uint32_t* ray_world, uint32_t mode, uint32_t cap,
                pr_ray_hit* hits, uint32_t* n_out);
/* TestPoint(const World&, BodyID, ShapeID, const Length2&) (WorldShape.cpp:178-181; DistanceProxy.cpp:107-158) for n (body, shape,
 * world-space point) triples at once: out[i] = 1 if the point, taken into the body's frame, lies in a child of the shape. */
int pr_test_points(const pr_world* w, uint32_t n, const uint32_t* body_ids, const uint32_t* shape_ids, const float* points_xy, uint32_t* out);
/* FindClosestBody(const World&, const Length2&) (WorldBody.cpp:576-592) for n locations at once: the body whose location is
 * nearest, the lowest id on a tie; 0xFFFFFFFF in a world without bodies. */
int pr_find_closest_bodies(const pr_world* w, uint32_t n, const float* points_xy, uint32_t* out_body_ids);
/* SetJoint(AabbTreeWorld&, JointID, Joint) (AabbTreeWorld.cpp:1108-1130) for any row of the joint table: rows[i] replaces joint ids[i] --
 * every field exactly as the value says, the solver's stored impulses included. The joint's OLD bodies are woken and, when the old value
 * kept them from colliding, their contacts flagged for filtering (Remove(JointID) :1177-1213); the NEW value's bodies are not woken, their
 * contacts flagged when collideConnected is false (Add(JointID, flag) :1159-1175). The new value may name other bodies, another
 * collideConnected or another joint type (not to or from a gear joint: PR_ERR_UNSUPPORTED, its value is wider than a row).
 * PR_ERR_OUT_OF_RANGE for an unknown joint or body id, PR_ERR_INVALID_ARGUMENT for a destroyed joint or an id named twice. */
int pr_set_revolute_joints(pr_world* w, uint32_t n, const uint32_t* ids, const pr_revolute_joint* rows);
/* Destroy(AabbTreeWorld&, JointID) (AabbTreeWorld.cpp:1215-1227): both bodies are woken, the joint stops taking part in islands,
 * solves and the collideConnected filter, its row reads back zeroed with body_a == body_b == 0xFFFFFFFF, and its id is the
 * next one pr_create_revolute_joints hands out (last freed first, like the reference's joint pool). Destroying a destroyed
 * joint does nothing. (Contacts between the two bodies that a collideConnected == false joint had kept from forming appear
 * when the broadphase next looks at their proxies, as in the reference.) */
int pr_destroy_joints(pr_world* w, uint32_t n, const uint32_t* ids);
/* Destroy(AabbTreeWorld&, BodyID) (AabbTreeWorld.cpp:1062-1106), one body after the other: the joints attached to the body are
 * destroyed (pr_destroy_joints; in the order they were added to the body, m_bodyJoints -- the order their ids are freed in), its contacts are destroyed in the body's contact-list order (end-contact events
 * for touching ones when the event mask asks for them; the other body is woken when the manifold had points, :2245-2279), its
 * pending (body, shape) pairs leave the proxy queue, its proxies leave the broadphase in leaf-id order (FindProxies :617) and
 * the body id is the next one pr_create_bodies hands out (last freed first, ObjectPool). Leaf ids are recycled like
 * DynamicTree's free list hands them out (DynamicTree.cpp:517-569), so later contact keys stay the reference's. The row of a
 * destroyed body reads back with flags == PR_BODY_DESTROYED. PR_ERR_OUT_OF_RANGE: unknown id; PR_ERR_INVALID_ARGUMENT: the id
 * is destroyed already (the reference's WasDestroyed). */
int pr_destroy_bodies(pr_world* w, uint32_t n, const uint32_t* ids);
/* Detach(World&, BodyID, ShapeID) = the SetBody whose new value has FEWER shapes (AabbTreeWorld.cpp:2760-2785): one entry of
 * shape_ids per attachment to remove. The proxies of the (body, shape) pairs leave the broadphase in the body's proxy-list
 * order, the contacts of each pair are destroyed, the pairs leave the proxy queue. The body row (new mass data) is uploaded
 * with pr_set_bodies as for any SetBody. */
int pr_detach_shapes(pr_world* w, uint32_t body, uint32_t n, const uint32_t* shape_ids);
/* Destroy(AabbTreeWorld&, ShapeID) (:1259-1277): the shape is detached from every body that has it (body-id order, like
 * pr_detach_shapes) and its id is the next one pr_create_shapes hands out. Destroying a destroyed shape does nothing. */
int pr_destroy_shapes(pr_world* w, uint32_t n, const uint32_t* ids);
/* SetShape(AabbTreeWorld&, ShapeID, Shape) (:1284-1356) for ANY new value. def: the new material / sensor flag / filter and,
 * when geometry_changed != 0, child_count children (vertex offsets relative to `vertices`, n_vertices entries; def->child_offset
 * is ignored). Every body using the shape is woken; with new geometry every ENABLED user loses the shape's proxies and contacts
 * and is queued for new proxies at the next Step; with another filter the shape's contacts are flagged for filtering (tested
 * again by the next Step's DestroyContacts, :2305-2328) and the shape's proxies are queued for contact finding; when the sensor
 * flag, friction, restitution or the geometry changed the shape's contacts are flagged for updating and their bodies woken. */
int pr_set_shape(pr_world* w, uint32_t shape, const pr_shape* def, const pr_child* children, const pr_vec2* vertices,
                 const pr_vec2* normals, uint32_t n_vertices, uint32_t geometry_changed);
/* The SetBody whose value differs in IsEnabled (GetOldAndNewShapeIDs, AabbTreeWorld.cpp:647-680): disabling destroys the body's
 * proxies and contacts (its shapes stay attached), enabling queues every attached shape for proxy creation at the next Step. */
int pr_set_body_enabled(pr_world* w, uint32_t body, uint32_t enabled);
/* The SetBody whose value has another BodyType (:2738-2758): the body's contacts are destroyed, then row replaces the body (flags
 * carry the new type's velocity / acceleration bits, mass data as the caller computed them); a body turned static has its
 * proxies synchronised by the next Step (m_bodiesForSync). */
int pr_set_body_type(pr_world* w, uint32_t body, const pr_body* row);
int pr_get_revolute_joints(const pr_world* w, uint32_t first, uint32_t n, pr_revolute_joint* rows);
int pr_get_counts(const pr_world* w, pr_counts* out);
/* contacts in m_contacts order (creation order), manifolds[i] belongs to contacts[i] */
int pr_get_contacts(const pr_world* w, uint32_t cap, pr_contact* contacts, pr_manifold* manifolds,
                    uint32_t* n_out);
int pr_get_proxies(const pr_world* w, uint32_t cap, pr_proxy* proxies, uint32_t* n_out);
/* SetContact(AabbTreeWorld&, ContactID, Contact) (AabbTreeWorld.cpp:2797-2831) for n contacts, rows[i].id naming the contact:
 * the value replaces the contact -- i.e. its flag bits (enabled, touching, dirty/needs-updating, filter), friction, restitution
 * and tangent speed -- after the reference's checks: PR_ERR_OUT_OF_RANGE for an id out of range, PR_ERR_INVALID_ARGUMENT for a
 * freed id (WasDestroyed) or a change of the is-destroyed flag, either contactable, the impenetrable or sensor flag, the TOI or
 * the TOI count. All rows are checked before any is applied.
 * pr_set_manifolds = SetManifold (:2833-2846): type and point count must not change (points, normal and impulses may). */
int pr_set_contacts(pr_world* w, uint32_t n, const pr_contact* rows);
int pr_set_manifolds(pr_world* w, uint32_t n, const uint32_t* ids, const pr_manifold* rows);
/* whole-state restore for stage-parity tests and clone-from-host: replaces every proxy's fat AABB
 * and the complete contact list (ids, flags, manifolds incl. stored impulses). */
int pr_load_state(pr_world* w, uint32_t n_bodies, const pr_body* bodies, uint32_t n_proxies,
                  const pr_proxy* proxies, uint32_t n_contacts, const pr_contact* contacts,
                  const pr_manifold* manifolds, float inv_dt0);

/* ---- the hot path: Step(AabbTreeWorld&, const StepConf&), AabbTreeWorld.cpp:2144-2219 ---- */
int pr_step(pr_world* w, const pr_step_conf* conf, pr_step_stats* stats_out);
/* the constraint order the last regular-phase solve used (colour-sorted):
 * contact ids in solve order, their colour and island label. For the parity oracle. */
int pr_get_solve_order(const pr_world* w, uint32_t cap, uint32_t* contact_ids, uint32_t* colors,
                       uint32_t* n_out);
/* the TOI event schedule of the last Step (recorded while debug bit 0 is set): (round, contact id) pairs ordered by round then
 * id. In PR_TOI_SERIAL_EXACT every round holds the one event GetSoonestContact picks (AabbTreeWorld.cpp:2870-2890); in
 * PR_TOI_PARALLEL a round holds every locally-soonest event. The parity oracle derives the same rounds by rule and must agree. */
int pr_get_toi_schedule(const pr_world* w, uint32_t cap, uint32_t* round_contact_pairs, uint32_t* n_out);
/* debug flags: bit 0 = record the solve order of every step (costs a device->host copy);
 * bit 1 = use the binned narrowphase, bit 2 = use the persistent fused solver kernel (two experiments that did not beat the
 * default kernels on B200; off by default, kept under test); bit 3 = TOI rounds through the list kernels only (no single-block
 * late rounds: A/B measurements); bit 4 = hub rule of the colouring (colours capped at 24 when a few bodies carry very many
 * constraints, the rest in the per-island serial tail: measured no faster on the tumbler arena); bit 5 = a batch arena solves
 * through the arena-wide per-colour launches instead of the block solver (one CTA per world, pr_block_solve.cuh): A/B runs */
int pr_set_debug(pr_world* w, uint32_t flags);
int pr_get_step_profile(const pr_world* w, pr_step_profile* out);
int pr_abi_sizeof(int which); /* sizeof of the POD structs above, for foreign-language mirrors */
/* ShiftOrigin(AabbTreeWorld&, const Length2&), AabbTreeWorld.cpp:2221-2243: every body's sweep positions 0 and 1 move by
 * -newOrigin (SetSweep recomputes the transformation), every contact of every body is flagged for updating, every tree
 * AABB moves by -newOrigin (DynamicTree::ShiftOrigin, DynamicTree.cpp:650-658). The revolute joint keeps no world-space
 * anchor (RevoluteJointConf.hpp:222). PR_ERR_WRONG_STATE while a sub-stepped Step is incomplete. */
int pr_shift_origin(pr_world* w, float x, float y);

/* ---- contact events: the begin/end/post-solve listeners of a Step, recorded on the device ----
 * The reference invokes std::function listeners synchronously INSIDE Step (begin/end: Update, AabbTreeWorld.cpp:2703-2714,
 * and InternalDestroy :2249-2252 for a touching contact that is destroyed; post-solve: Report :196-206 after each regular
 * island :1671-1674 and each TOI island :2053-2055). A device step cannot call back into the host, so the kernels append
 * event records to device buffers instead and the host reads them after pr_step (pr_events); the CudaWorld shim then invokes
 * the stored listeners, in (kind, contact id) order. Pre-solve (:2716-2720) is recorded the same way -- the contact id and the
 * manifold it had BEFORE the update, for every Update that left a non-sensor contact touching -- and delivered after the step:
 * a listener can observe it, but what it does to the contact (the reference lets it disable the contact for the solve of the
 * same step) takes effect from the next step on.
 * Nothing is recorded (and the kernels pay one uniform branch) until pr_set_event_mask asks for it. */
enum { PR_EVENT_BEGIN = 1, PR_EVENT_END = 2, PR_EVENT_PRE_SOLVE = 4, PR_EVENT_POST_SOLVE = 8 };
typedef struct pr_contact_event { /* begin / end: the contact as the listener would have seen it through GetContact */
    uint32_t kind;                      /* PR_EVENT_BEGIN or PR_EVENT_END */
    uint32_t contact;                   /* ContactID */
    uint32_t body_a, shape_a, child_a;
    uint32_t body_b, shape_b, child_b;
} pr_contact_event;
typedef struct pr_impulse_event { /* post-solve: (ContactID, ContactImpulsesList, iteration_type) */
    uint32_t contact;
    uint32_t count;                     /* entries of the ContactImpulsesList = the velocity constraint's point count */
    uint32_t iterations;                /* regular island: positionIters - 1 if solved else 255 (StepConf::InvalidIteration);
                                           TOI island: positionIters */
    uint32_t phase;                     /* 0 regular phase, 1 TOI phase */
    float normal[2], tangent[2];
} pr_impulse_event;
typedef struct pr_presolve_event { /* pre-solve: (ContactID, const Manifold& oldManifold) */
    uint32_t contact;
    uint32_t pad;
    pr_manifold old_manifold;
} pr_presolve_event;
/* mask: PR_EVENT_* bits to record from the next pr_step on (0 = off); capacity: records kept per step and kind group
 * (begin+end share one buffer, post-solve and pre-solve have their own); further events of a step are counted but dropped. */
int pr_set_event_mask(pr_world* w, uint32_t mask, uint32_t capacity);
/* Events of the LAST pr_step. kind = PR_EVENT_BEGIN | PR_EVENT_END (either or both): buf = pr_contact_event[cap];
 * kind = PR_EVENT_POST_SOLVE: buf = pr_impulse_event[cap]; kind = PR_EVENT_PRE_SOLVE: buf = pr_presolve_event[cap] (a contact updated
 * in the regular phase and again in the TOI phase has a record for each). Sorted by (kind, contact id) resp. (phase, contact id); *n_out =
 * matching events (at most cap are written). If the step produced more events than the capacity given to pr_set_event_mask
 * it dropped the rest: PR_ERR_LENGTH, and *n_out = the number the step produced (begin+end together, or post-solve). */
int pr_events(const pr_world* w, uint32_t kind, void* buf, uint32_t cap, uint32_t* n_out);
/* TOI event scheduling (SolveToi loop, AabbTreeWorld.cpp:1752-1831):
 *   PR_TOI_SERIAL_EXACT  one event per round in the reference's order (global soonest first): bit-exact
 *   PR_TOI_PARALLEL      every locally-soonest event whose bodies and contact neighbours are disjoint
 *                        from all sooner events runs in the same round (default; drift vs the serial
 *                        order is reported by the tests, not gated -- BASELINE.json north_star) */
enum { PR_TOI_SERIAL_EXACT = 0, PR_TOI_PARALLEL = 1 };
int pr_set_toi_mode(pr_world* w, uint32_t mode);
int pr_set_sub_stepping(pr_world* w, uint32_t flag);    /* SetSubStepping, AabbTreeWorld.hpp */
int pr_get_sub_stepping(const pr_world* w);             /* GetSubStepping */
int pr_is_step_complete(const pr_world* w);             /* IsStepComplete */
float pr_get_inv_dt0(const pr_world* w);                /* GetInvDeltaTime */
/* the cudaStream_t every kernel and copy of this world is issued on (for the caller's CUDA events / ordering) */
void* pr_get_stream(const pr_world* w);

/* ---- stage entry point for known-answer tests: CollideShapes (Manifold.cpp:393) over arrays ----
 * pair i collides child a_child[i] under xf_a[i] (px,py,c,s) with child b_child[i] under xf_b[i]. */
int pr_collide_shapes(int device, uint32_t n, const pr_child* children, uint32_t n_children,
                      const pr_vec2* vertices, const pr_vec2* normals, uint32_t n_vertices,
                      const uint32_t* a_child, const float* xf_a, const uint32_t* b_child,
                      const float* xf_b, float linear_slop, float max_circles_ratio,
                      pr_manifold* out);

/* GetToiViaSat (TimeOfImpact.cpp:35) over arrays: sweeps are 8 floats (pos0 x,y,a | pos1 x,y,a | local
 * centre x,y), alpha0 = 0; out_iters has 3 entries per pair: toi_iters, max_dist_iters, max_root_iters;
 * out_state is ToiOutput::State (ToiOutput.hpp:73). */
int pr_toi_batch(int device, uint32_t n, const pr_child* children, uint32_t n_children,
                 const pr_vec2* vertices, const pr_vec2* normals, uint32_t n_vertices,
                 const uint32_t* a_child, const float* sweep_a, const uint32_t* b_child,
                 const float* sweep_b, const pr_step_conf* conf, float* out_time, int32_t* out_state,
                 uint32_t* out_iters);

/* ---- batched independent worlds (BASELINE config 5; no reference API, additive: SURVEY §8(b),(e)) ----
 * n_worlds worlds that share nothing (World.hpp:652-656) live in ONE device arena: bodies carry a world
 * index, the broadphase never pairs proxies of different worlds, and one pr_batch_step is one
 * Step(world, conf) of every world -- each kernel launch is n_worlds wide. World w's bodies behave bit for
 * bit like the same bodies in a reference world whose shapes carry a per-world collision Filter.
 * pr_batch_arena() is an ordinary pr_world handle: create shapes (shared by all worlds) and joints on it,
 * read bodies/contacts/joints back from it; body ids are arena-wide. Bodies are created with
 * pr_batch_create_bodies (pr_create_bodies on the arena is PR_ERR_INVALID_ARGUMENT). Statistics returned by
 * pr_batch_step are sums (min/max for the two float fields) over all worlds. */
typedef struct pr_batch pr_batch;
int pr_batch_create(const pr_world_conf* conf, uint32_t n_worlds, pr_batch** out);
int pr_batch_destroy(pr_batch* b);
pr_world* pr_batch_arena(pr_batch* b);
uint32_t pr_batch_world_count(const pr_batch* b);
int pr_batch_create_bodies(pr_batch* b, uint32_t n, const pr_body* bodies, const uint32_t* world_index,
                           const uint32_t* shape_offsets, const uint32_t* shape_ids, uint32_t* first_id);
int pr_batch_step(pr_batch* b, const pr_step_conf* conf, pr_step_stats* stats_sum_out);
int pr_batch_get_body_worlds(const pr_batch* b, uint32_t first, uint32_t n, uint32_t* world_index);
/* pr_batch_create(n_worlds, device_mask) / pr_batch_stats. One process drives one GPU (the batch's worlds of the
 * other GPUs belong to the other ranks, playrho_b200/shard.py): pr_batch_create_masked creates the arena on the LOWEST device
 * whose bit is set in device_mask (PR_ERR_INVALID_ARGUMENT if none of the mask's devices exists); pr_batch_stats returns the
 * statistics of the last pr_batch_step again (the sums over this arena's worlds that the ranks all-reduce). */
int pr_batch_create_masked(const pr_world_conf* conf, uint32_t n_worlds, uint64_t device_mask, pr_batch** out);
int pr_batch_stats(const pr_batch* b, pr_step_stats* stats_sum_out);

#ifdef __cplusplus
}
#endif
#endif /* PLAYRHO_B200_H */
