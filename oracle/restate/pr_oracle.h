/*
 * pr_oracle.h -- TEST INFRASTRUCTURE ONLY (oracle). Never linked into or called by the product.
 *
 * C API of the plain, serial C++ restatement of the reference's Step pipeline
 * (Library/source/playrho/d2/AabbTreeWorld.cpp:2144 and what it calls). It uses the POD layouts
 * of include/playrho_b200.h so the same arrays can be fed to the oracle, to the real reference
 * (oracle/ref_harness.cpp) and to the CUDA world.
 *
 * PINNING: tests/test_oracle_pin.py checks this restatement bit-for-bit against the real reference
 * (oracle/_ref, built from /root/reference by oracle/Makefile) on HelloWorld (incl. the reference's
 * golden ExpectedOutput.txt), pyramid, disk-box, mixed-polygon and chain scenes over whole
 * trajectories, and against known-answer vectors from UnitTests/CollideShapes.cpp.
 */
#ifndef PR_ORACLE_H
#define PR_ORACLE_H
#include "../../include/playrho_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world orc_world;

orc_world* orc_create(void);
void orc_destroy(orc_world* w);
const char* orc_last_error(const orc_world* w);
int orc_create_shapes(orc_world* w, uint32_t n, const pr_shape* shapes, const pr_child* children,
                      uint32_t n_children, const pr_vec2* vertices, const pr_vec2* normals,
                      uint32_t n_vertices);
int orc_create_bodies(orc_world* w, uint32_t n, const pr_body* bodies,
                      const uint32_t* shape_offsets, const uint32_t* shape_ids);
int orc_create_revolute_joints(orc_world* w, uint32_t n, const pr_revolute_joint* joints);
int orc_create_gear_joints(orc_world* w, uint32_t n, const pr_gear_joint* joints);
int orc_set_bodies(orc_world* w, uint32_t n, const uint32_t* ids, const pr_body* rows);
int orc_set_shape_material(orc_world* w, uint32_t shape, float friction, float restitution, float density, uint32_t is_sensor); /* SetShape :1284 */
int orc_attach_shapes(orc_world* w, uint32_t body, uint32_t n, const uint32_t* shape_ids); /* SetBody with added shapes :2760-2786 */
int orc_destroy_joints(orc_world* w, uint32_t n, const uint32_t* ids); /* Destroy(JointID) :1215 */
int orc_set_revolute_joints(orc_world* w, uint32_t n, const uint32_t* ids, const pr_revolute_joint* rows); /* SetJoint :1108 */
int orc_set_contacts(orc_world* w, uint32_t n, const pr_contact* rows);                 /* SetContact :2797 */
int orc_set_manifolds(orc_world* w, uint32_t n, const uint32_t* ids, const pr_manifold* rows); /* SetManifold :2833 */
int orc_shift_origin(orc_world* w, float x, float y); /* ShiftOrigin, AabbTreeWorld.cpp:2221 */
/* the begin/end/post-solve listener calls of the last orc_step, in the reference's call order */
int orc_set_event_mask(orc_world* w, uint32_t mask);
int orc_events(const orc_world* w, uint32_t kind, void* buf, uint32_t cap, uint32_t* n_out);
int orc_load_state(orc_world* w, uint32_t n_bodies, const pr_body* bodies, uint32_t n_proxies,
                   const pr_proxy* proxies, uint32_t n_contacts, const pr_contact* contacts,
                   const pr_manifold* manifolds, float inv_dt0);
/* order == NULL: the reference's native constraint order (island DFS order).
 * order != NULL: n_order contact ids; within each island, constraints are solved in the relative
 * order they appear in `order` (the GPU's colour-sorted order) -- "the reference solver fed the
 * same colour-sorted constraint order" (BASELINE.json north_star). */
int orc_step(orc_world* w, const pr_step_conf* conf, pr_step_stats* stats, const uint32_t* order,
             uint32_t n_order);
int orc_get_counts(const orc_world* w, pr_counts* out);
int orc_get_bodies(const orc_world* w, uint32_t first, uint32_t n, pr_body* rows);
int orc_get_contacts(const orc_world* w, uint32_t cap, pr_contact* contacts, pr_manifold* manifolds,
                     uint32_t* n_out);
int orc_get_proxies(const orc_world* w, uint32_t cap, pr_proxy* proxies, uint32_t* n_out);
int orc_get_revolute_joints(const orc_world* w, uint32_t first, uint32_t n, pr_revolute_joint* rows);
float orc_get_inv_dt0(const orc_world* w);
/* TOI event schedule: 0 = the reference's (one globally soonest contact per loop turn, AabbTreeWorld.cpp:1752-1831);
 * 1 = the product's PR_TOI_PARALLEL rounds, derived here by the documented rule (orc_toi_phase.inc, solve_toi_parallel).
 * orc_get_toi_schedule: (round, contact id) of every event of the last step, mode 1. */
int orc_set_toi_mode(orc_world* w, uint32_t mode);
int orc_get_toi_schedule(const orc_world* w, uint32_t cap, uint32_t* round_contact_pairs, uint32_t* n_out);

/* RayCast(world, input, callback returning ResetRay): every hit, brute force over the proxies with the tree's two leaf tests
 * (RayCastOutput.cpp:241-337); hits in proxy order per ray */
int orc_ray_cast(const orc_world* w, uint32_t n_rays, const float* rays5, uint32_t cap, pr_ray_hit* hits, uint32_t* n_out);

/* stage functions on flat arrays */
int orc_collide_shapes(uint32_t n, const pr_child* children, const pr_vec2* verts,
                       const pr_vec2* normals, const uint32_t* a_child, const float* xf_a,
                       const uint32_t* b_child, const float* xf_b, float linear_slop,
                       float max_circles_ratio, pr_manifold* out);
void orc_step_conf_default(pr_step_conf* c);
int orc_toi(uint32_t n, const pr_child* children, const pr_vec2* verts, const pr_vec2* normals,
            const uint32_t* a_child, const float* sweep_a, const uint32_t* b_child,
            const float* sweep_b, const pr_step_conf* sc, float* out_time, int32_t* out_state,
            uint32_t* out_iters);
/* the glibc-2.39 sincosf restatement that the CUDA side ports (checked against libm here) */
void orc_sincosf_restated(uint32_t n, const float* x, float* cs);

#ifdef __cplusplus
}
#endif
#endif
