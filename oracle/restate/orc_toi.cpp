// orc_toi.cpp -- TEST INFRASTRUCTURE ONLY (oracle restatement). STUBS: filled in by the TOI milestone.
#include "orc_toi.h"
namespace orc {
float test_overlap(const Proxy&, Xf, const Proxy&, Xf, uint32_t) { return -1.0f; }
ToiOut toi_via_sat(const Proxy&, const Sweep&, const Proxy&, const Sweep&, const pr_step_conf&) { return {1.0f, 0, 0, 0, 0}; }
void revolute_init_velocity(pr_revolute_joint&, RevoluteWork&, float, float, V2, Pos, Vel&, float, float, V2, Pos, Vel&, const JointSolverConf&) {}
bool revolute_solve_velocity(pr_revolute_joint&, RevoluteWork&, float, float, Vel&, float, float, Vel&, const JointSolverConf&) { return true; }
bool revolute_solve_position(const pr_revolute_joint&, float, float, V2, Pos&, float, float, V2, Pos&, const JointSolverConf&) { return true; }
}

#include "pr_oracle.h"
extern "C" {
int orc_toi(uint32_t n, const pr_child* children, const pr_vec2* verts, const pr_vec2* normals,
            const uint32_t* a_child, const float* sweep_a, const uint32_t* b_child, const float* sweep_b,
            const pr_step_conf* sc, float* out_time, int32_t* out_state, uint32_t* out_iters)
{
    using namespace orc;
    for (uint32_t i = 0; i < n; ++i) {
        const pr_child& ca = children[a_child[i]];
        const pr_child& cb = children[b_child[i]];
        const Proxy pa{ca.vertex_radius, ca.vertex_count, verts + ca.vertex_offset, normals + ca.vertex_offset};
        const Proxy pb{cb.vertex_radius, cb.vertex_count, verts + cb.vertex_offset, normals + cb.vertex_offset};
        const float* a = sweep_a + 8 * i;
        const float* b = sweep_b + 8 * i;
        const Sweep sa{{{a[0], a[1]}, a[2]}, {{a[3], a[4]}, a[5]}, {a[6], a[7]}, 0.0f};
        const Sweep sb{{{b[0], b[1]}, b[2]}, {{b[3], b[4]}, b[5]}, {b[6], b[7]}, 0.0f};
        const ToiOut o = toi_via_sat(pa, sa, pb, sb, *sc);
        out_time[i] = o.time; out_state[i] = o.state;
        out_iters[3 * i] = o.toi_iters; out_iters[3 * i + 1] = o.max_dist_iters; out_iters[3 * i + 2] = o.max_root_iters;
    }
    return 0;
}
void orc_sincosf_restated(uint32_t n, const float* x, float* cs)
{
    for (uint32_t i = 0; i < n; ++i) { cs[2 * i] = std::cos(x[i]); cs[2 * i + 1] = std::sin(x[i]); }
}
}
