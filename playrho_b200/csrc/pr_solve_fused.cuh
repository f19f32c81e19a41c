// pr_solve_fused.cuh -- the regular phase's constraint solve as ONE persistent cooperative kernel.
//
// SolveRegIslandViaGS (AabbTreeWorld.cpp:1563-1682) is warm start, 8 velocity iterations, IntegratePositions and 3
// position iterations; with graph colouring every iteration is one launch per colour (14 colours at 1M bodies), i.e.
// ~190 dependent launches per step of ~70k rows each -- each too small to fill the GPU and each paying launch ramp-up
// and tail (measured: 8.4 us per launch for 13 MB of rows, 24% of HBM peak). Here the whole sequence runs in one
// kernel of (SM count x resident blocks) CTAs: a colour is a grid-stride loop, colours are separated by a grid barrier
// (one release-add + acquire-poll per CTA), and before arriving at a barrier every thread prefetches the rows it will
// solve in the NEXT phase into L2 (row data does not depend on the previous colour; only the body velocities do), so
// that after the barrier the dependent chain is L2 row load -> body load -> solve -> store.
// The arithmetic is the very same solve_velocity_row / solve_position_row / joint functions the per-colour kernels
// call (those remain: host-emulated tests, and the fallback when a cooperative launch is not possible).
#ifndef PR_SOLVE_FUSED_CUH
#define PR_SOLVE_FUSED_CUH
#ifndef PR_HOST_EMUL

namespace pr {

struct SolvePlan {
    unsigned nColors;               // coloured ranges (without the overflow tail)
    unsigned start[MAX_COLORS], count[MAX_COLORS];
    unsigned tailStart, tailCount, nTailSegs;
    unsigned nBodies, nJoints, jointsSerial, warmStart;
};

// all CTAs of the (co-resident) grid; *counter grows by gridDim.x per barrier and is zeroed by the host before launch
PR_D void grid_barrier(unsigned* counter, unsigned& phase)
{
    ++phase;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = phase * gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncthreads();
}
PR_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// rows [start, start+count) of the next phase, as this thread will walk them
PR_D void prefetch_velocity_rows(const DV& d, unsigned start, unsigned count)
{
    for (unsigned t = PR_TID; t < count; t += PR_NTHREADS) {
        const unsigned i = start + t;
        prefetch_l2(&d.vc0[i]); prefetch_l2(&d.vc1[i]); prefetch_l2(&d.vc2[i]); prefetch_l2(&d.vc3[i]); prefetch_l2(&d.vc4[i]);
        prefetch_l2(&d.vc5[i]); prefetch_l2(&d.vc6[i]); prefetch_l2(&d.vc7[i]); prefetch_l2(&d.vc8[i]);
    }
}
PR_D void prefetch_position_rows(const DV& d, unsigned start, unsigned count)
{
    for (unsigned t = PR_TID; t < count; t += PR_NTHREADS) {
        const unsigned i = start + t;
        prefetch_l2(&d.vc1[i]); prefetch_l2(&d.vc2[i]); prefetch_l2(&d.pc0[i]); prefetch_l2(&d.pc1[i]); prefetch_l2(&d.pc2[i]); prefetch_l2(&d.pc3[i]);
    }
}
PR_D void fused_joints(const DV& d, const pr_step_conf& conf, const SolvePlan& p, unsigned phase)
{
    if (p.jointsSerial) {
        if (PR_TID == 0u)
            for (unsigned ji = 0; ji < p.nJoints; ++ji) {
                if (phase == 0u) joint_init_velocity(d, ji, conf, jview_global(d));
                else if (phase == 1u) joint_solve_velocity(d, ji, conf, jview_global(d));
                else joint_solve_position(d, ji, conf, jview_global(d));
            }
    } else {
        for (unsigned ji = PR_TID; ji < p.nJoints; ji += PR_NTHREADS) {
            if (phase == 0u) joint_init_velocity(d, ji, conf, jview_global(d));
            else if (phase == 1u) joint_solve_velocity(d, ji, conf, jview_global(d));
            else joint_solve_position(d, ji, conf, jview_global(d));
        }
    }
}
template <class F> PR_D void fused_tail(const DV& d, const SolvePlan& p, F&& fn)
{ // colour-overflow tail: one thread per island segment, rows in list order (see k_tail_segments)
    for (unsigned s = PR_TID; s < p.nTailSegs; s += PR_NTHREADS) {
        unsigned t = d.k_ulist[s];
        const unsigned label = d.vc_island[p.tailStart + t];
        for (; t < p.tailCount && d.vc_island[p.tailStart + t] == label; ++t) fn(p.tailStart + t);
    }
}

// ---- the trailing SMALL colours of an iteration in one launch ---------------------------------------------------
// Greedy colouring leaves a tail of nearly empty colours (1M mixed scene: 127k, 125k, ... 21k, 5.4k, 891, 108, 12, 1
// rows): each still cost a launch per iteration. Colours of at most SMALL_COLOR_ROWS rows are solved by ONE block, a
// block barrier between colours (same SM: the barrier orders their global writes).
constexpr unsigned SMALL_COLOR_ROWS = 1024u, SMALL_COLOR_MAX = 24u, SMALL_BLOCK = 1024u;
struct SmallPlan { unsigned n; unsigned start[SMALL_COLOR_MAX], count[SMALL_COLOR_MAX]; };
// mode 0 warm start, 1 velocity, 2 position
__global__ void __launch_bounds__(SMALL_BLOCK, 1) k_solve_small_colors(DV d, SmallPlan p, unsigned mode, float rate, float slop, float maxCorr)
{
    PR_PDL_ENTER();
    const RS r = rs_global(d);
    const unsigned t = threadIdx.x;
    for (unsigned c = 0; c < p.n; ++c) {
        const unsigned start = p.start[c], count = p.count[c];
        if (mode == 0u) { if (t < count) warm_start_row(r, start + t); }
        else if (mode == 1u) {
            bool did = false; float inc = 0.0f; unsigned label = 0u;
            if (t < count) { inc = solve_velocity_row(d, r, start + t, did); label = d.vc_island[start + t]; }
            island_max(d.isl_inc, label, f2u(inc), did);
        } else {
            bool did = false; float sep = PR_INF; unsigned label = 0u;
            if (t < count) { sep = solve_position_row(d, r, start + t, rate, slop, maxCorr, did); label = d.vc_island[start + t]; }
            island_min(d.isl_sep, label, float_flip(sep), did);
        }
        __threadfence_block();
        __syncthreads();
    }
}

constexpr unsigned FUSED_BLOCK = 512u; // 2 CTAs of 512 per SM (<= 64 registers): every colour of the 1M scene is one row per thread
__global__ void __launch_bounds__(FUSED_BLOCK, 2) k_solve_fused(DV d, pr_step_conf conf, SolvePlan p, unsigned* barrier)
{
    unsigned phase = 0u;
    const RS r = rs_global(d);
    const unsigned lane = threadIdx.x & 31u;
    unsigned seenVel = 0u, seenPos = 0u; // cnt->nActiveVel / nActivePos are zero at step start and only grow in here
    // ---- WarmStartVelocities (:264) per colour, then the joints' InitVelocity (:1597)
    if (p.warmStart) {
        for (unsigned c = 0; c < p.nColors; ++c) {
            for (unsigned t = PR_TID; t < p.count[c]; t += PR_NTHREADS) warm_start_row(r, p.start[c] + t);
            if (c + 1u < p.nColors) prefetch_velocity_rows(d, p.start[c + 1u], p.count[c + 1u]);
            grid_barrier(barrier, phase);
        }
        if (p.tailCount) { fused_tail(d, p, [&](unsigned row) { warm_start_row(r, row); }); grid_barrier(barrier, phase); }
    }
    if (p.nJoints) { fused_joints(d, conf, p, 0u); grid_barrier(barrier, phase); }
    // ---- velocity iterations (:1603-1622)
    for (unsigned it = 0; it < conf.reg_velocity_iters; ++it) {
        if (p.nJoints) { fused_joints(d, conf, p, 1u); grid_barrier(barrier, phase); } // joints first (:1605-1608)
        for (unsigned c = 0; c < p.nColors; ++c) {
            const unsigned start = p.start[c], count = p.count[c];
            for (unsigned t = PR_TID; t - lane < count; t += PR_NTHREADS) { // (whole warps iterate together: island_max is a warp primitive)
                bool did = false; float inc = 0.0f; unsigned label = 0u;
                if (t < count) { inc = solve_velocity_row(d, r, start + t, did); label = d.vc_island[start + t]; }
                island_max(d.isl_inc, label, f2u(inc), did);
            }
            const unsigned nc = (c + 1u < p.nColors) ? c + 1u : 0u;
            prefetch_velocity_rows(d, p.start[nc], p.count[nc]);
            grid_barrier(barrier, phase);
        }
        if (p.tailCount) {
            fused_tail(d, p, [&](unsigned row) {
                bool did = false;
                const float inc = solve_velocity_row(d, r, row, did);
                if (did) atomicMax(&d.isl_inc[d.vc_island[row]], f2u(inc));
            });
            grid_barrier(barrier, phase);
        }
        // per-island early exit bookkeeping (k_isl_vel_check); the grid stops iterating once no island is active
        unsigned still = 0u;
        for (unsigned i = PR_TID; i < p.nBodies; i += PR_NTHREADS) {
            if (d.b_island[i] == i) {
                const unsigned st = d.isl_state[i];
                if ((st & ISL_ACTIVE) && !(st & ISL_VEL_DONE)) {
                    const unsigned incBits = d.isl_inc[i];
                    if (incBits > d.isl_maxInc[i]) d.isl_maxInc[i] = incBits;
                    const bool jointsOkay = d.isl_jflag[i] == 0u;
                    if (jointsOkay && u2f(incBits) <= conf.reg_min_momentum) { d.isl_state[i] = st | ISL_VEL_DONE; d.isl_velIters[i] = it + 1u; }
                    else ++still;
                    d.isl_inc[i] = 0u; d.isl_jflag[i] = 0u;
                }
            }
        }
        block_add(&d.cnt->nActiveVel, still); // cumulative over the iterations (zero at step start): never reset inside the kernel
        grid_barrier(barrier, phase);
        const unsigned totalVel = *reinterpret_cast<volatile unsigned*>(&d.cnt->nActiveVel);
        if (totalVel == seenVel) break; // no island still active: every remaining iteration would be a no-op
        seenVel = totalVel;
    }
    // ---- IntegratePositions (:176)
    for (unsigned i = PR_TID; i < p.nBodies; i += PR_NTHREADS) {
        if (d.b_island[i] != INVALID) {
            const float4 v = d.s_vel[i];
            float4 q = d.s_pos[i];
            const float2 tr = conf.delta_time * v2(v.x, v.y);
            const float rot = conf.delta_time * v.z;
            q.x += tr.x; q.y += tr.y; q.z += rot;
            d.s_pos[i] = q;
        }
    }
    if (p.nColors) prefetch_position_rows(d, p.start[0], p.count[0]);
    grid_barrier(barrier, phase);
    // ---- position iterations (:1628-1644)
    for (unsigned it = 0; it < conf.reg_position_iters; ++it) {
        for (unsigned c = 0; c < p.nColors; ++c) {
            const unsigned start = p.start[c], count = p.count[c];
            for (unsigned t = PR_TID; t - lane < count; t += PR_NTHREADS) {
                bool did = false; float sep = PR_INF; unsigned label = 0u;
                if (t < count) {
                    sep = solve_position_row(d, r, start + t, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction, did);
                    label = d.vc_island[start + t];
                }
                island_min(d.isl_sep, label, float_flip(sep), did);
            }
            const unsigned nc = (c + 1u < p.nColors) ? c + 1u : 0u;
            prefetch_position_rows(d, p.start[nc], p.count[nc]);
            grid_barrier(barrier, phase);
        }
        if (p.tailCount) {
            fused_tail(d, p, [&](unsigned row) {
                bool did = false;
                const float sep = solve_position_row(d, r, row, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction, did);
                if (did) atomicMin(&d.isl_sep[d.vc_island[row]], float_flip(sep));
            });
            grid_barrier(barrier, phase);
        }
        if (p.nJoints) { fused_joints(d, conf, p, 2u); grid_barrier(barrier, phase); } // joints after the contacts (:1634-1637)
        unsigned still = 0u;
        for (unsigned i = PR_TID; i < p.nBodies; i += PR_NTHREADS) {
            if (d.b_island[i] == i) {
                const unsigned st = d.isl_state[i];
                if ((st & ISL_ACTIVE) && !(st & ISL_POS_SOLVED)) {
                    const unsigned sepFlip = d.isl_sep[i];
                    if (sepFlip < d.isl_minSep[i]) d.isl_minSep[i] = sepFlip;
                    const bool jointsOkay = d.isl_jflag[i] == 0u;
                    if (jointsOkay && float_unflip(sepFlip) >= conf.reg_min_separation) { d.isl_state[i] = st | ISL_POS_SOLVED; d.isl_posIters[i] = it + 1u; }
                    else ++still;
                    d.isl_sep[i] = float_flip(PR_INF); d.isl_jflag[i] = 0u;
                }
            }
        }
        block_add(&d.cnt->nActivePos, still);
        grid_barrier(barrier, phase);
        const unsigned totalPos = *reinterpret_cast<volatile unsigned*>(&d.cnt->nActivePos);
        if (totalPos == seenPos) break;
        seenPos = totalPos;
    }
}

} // namespace pr
#endif // !PR_HOST_EMUL
#endif
