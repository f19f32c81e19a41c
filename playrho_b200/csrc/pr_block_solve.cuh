// pr_block_solve.cuh -- the regular phase's constraint solve for an ARENA of independent worlds (BASELINE config 5:
// thousands of 512-body tumblers): ONE CTA PER WORLD, the world's BodyConstraints and island flags in shared memory, and
// the whole of SolveRegIslandViaGS (AabbTreeWorld.cpp:1563-1682) -- warm start, the joints' InitVelocity, every velocity
// iteration, IntegratePositions, every position iteration, the per-island early exits -- inside one launch.
//
// Why: solved through the arena-wide per-colour launches a tumbler world costs ~750 DEPENDENT launches per step whatever
// the number of worlds (its enclosure touches ~100 squares: 65 colours + an overflow tail, times 11 iterations), so 1 024
// worlds per GPU took as long as 4 096 (r01h: velocity + position 5.0 of 7.8 ms) and 8 GPUs scaled 3.75x. Worlds never
// share a body, so a world's rows are a closed Gauss-Seidel problem: a CTA walks the world's colours with block barriers
// instead of launches, the colour-overflow tail (the hub's rows) is one thread's loop with no barrier at all, and the
// bodies never leave shared memory between the first warm start and the last position iteration.
//
// Order = the order of the sorted row list (world, colour; tail rows in list order): exactly what pr_get_solve_order
// reports, so the oracle replays it like any other colour-sorted order. Rows of one colour share no mass-carrying body.
// Rows keep SLOTS (index of a body in its world's shared-memory arrays) instead of body ids, and the slot of the island's
// root instead of its label (k_build_constraints, blockMode).
#ifndef PR_BLOCK_SOLVE_CUH
#define PR_BLOCK_SOLVE_CUH

namespace pr {

#ifndef PR_HOST_EMUL
#define BLK_ID blockIdx.x
#define BLK_TID threadIdx.x
#define BLK_T blockDim.x
#define BLK_SYNC() __syncthreads()
#define BLK_ANY(x) (__syncthreads_or(static_cast<int>(x)) != 0)
#else // the tests' host emulation runs a "thread" after another: one emulated thread plays a whole CTA
#define BLK_ID (pr::emul_tid)
#define BLK_TID 0u
#define BLK_T 1u
#define BLK_SYNC() do { } while (0)
#define BLK_ANY(x) ((x) != 0u)
#endif

constexpr unsigned BLK_THREADS = 64u, BLK_MAX_THREADS = 256u; // (PR_BLK_THREADS picks another block size for experiments)
constexpr unsigned BLK_BYTES_PER_BODY = sizeof(float4) + 2u * sizeof(unsigned);
constexpr unsigned BLK_MAX_BODIES = 8192u; // 24 B per body of shared memory: 192 KB of the SM's 227 KB

// ------------------------------------------------------------------ bodies grouped by world (rebuilt when bodies are created)
PR_KERNEL k_blk_count(DV d, unsigned nBodies, unsigned* count)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) atomicAdd(&count[d.b_world[i]], 1u);
}
// sortedWorld / sortedBody: the bodies in (world, id) order
PR_KERNEL k_blk_fill(DV d, unsigned nBodies, const unsigned* sortedWorld, const unsigned* sortedBody)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) {
        const unsigned body = sortedBody[i];
        d.blk_body[i] = body;
        d.b_slot[body] = i - d.blk_off[sortedWorld[i]];
    }
}
PR_KERNEL k_blk_max(DV d, unsigned nWorlds)
{
    PR_PDL_ENTER();
    const unsigned w = PR_TID;
    block_max(&d.cnt->blkMaxBodies, w < nWorlds ? d.blk_off[w + 1u] - d.blk_off[w] : 0u);
}
// this step's constraints: sort key (world, colour) and the number of rows of every world
PR_KERNEL k_blk_row_keys(DV d, unsigned nBound, unsigned* keys, unsigned* rowCount)
{
    PR_PDL_ENTER();
    const unsigned k = PR_TID;
    unsigned color = 0u;
    if (k < nBound && k < d.cnt->nConstraints) {
        const unsigned world = d.b_world[d.c_ab[d.k_contact[k]].x];
        color = d.k_color[k];
        keys[k] = (world << 7) | color;
        atomicAdd(&rowCount[world], 1u);
        color = (color < COLOR_OVERFLOW) ? color + 1u : 0u;
    }
    block_max(&d.cnt->blkColors, color); // colours in use (profile; the overflow tail is not one)
}

// ------------------------------------------------------------------ the solve
// Shared memory per body of a world: ONE float4 (the BodyConstraint velocities during the velocity iterations, then the
// positions), one reduction word and one state word at the slot of an island's root -- 24 B, so that 18 worlds of 514
// bodies are resident per SM. The solve is a chain of dependent single-lane instruction streams (a hub's rows), and what
// hides it is the number of worlds an SM holds at once (measured with 52 B per body and 128 threads per world: 8 worlds
// per SM, issue slots 26% busy, time linear in the number of worlds from 1 024 up).
struct BlkView {
    float4* buf;          // BodyConstraint velocity (velocity phase) / position (position phase) of slot i
    unsigned *acc, *st;   // at the slot of an island's root: max incremental impulse / min separation of the iteration; ISL_* bits
    unsigned* cstart;     // first row of colour c (MAX_COLORS + 1 entries)
};
enum { BLK_ST_ACTIVE = 1u << 8, BLK_ST_ROOT = 1u << 9 }; // a body of an active island; an island's root (beside the ISL_* bits)
// the island checks of k_isl_vel_check / k_isl_pos_check (AabbTreeWorld.cpp:1603-1644) on the block's copies
PR_D unsigned blk_vel_check(const DV& d, const BlkView& s, unsigned i, unsigned body, unsigned iter, float minMomentum)
{
    const unsigned st = s.st[i];
    if (!(st & ISL_ACTIVE) || (st & ISL_VEL_DONE)) return 0u;
    const unsigned incBits = s.acc[i];
    if (incBits > d.isl_maxInc[body]) d.isl_maxInc[body] = incBits;
    unsigned still = 0u;
    if (d.isl_jflag[body] == 0u && u2f(incBits) <= minMomentum) { s.st[i] = st | ISL_VEL_DONE; d.isl_velIters[body] = iter + 1u; }
    else still = 1u;
    s.acc[i] = 0u; d.isl_jflag[body] = 0u;
    return still;
}
PR_D unsigned blk_pos_check(const DV& d, const BlkView& s, unsigned i, unsigned body, unsigned iter, float minSeparation)
{
    const unsigned st = s.st[i];
    if (!(st & ISL_ACTIVE) || (st & ISL_POS_SOLVED)) return 0u;
    const unsigned sepFlip = s.acc[i];
    if (sepFlip < d.isl_minSep[body]) d.isl_minSep[body] = sepFlip;
    unsigned still = 0u;
    if (d.isl_jflag[body] == 0u && float_unflip(sepFlip) >= minSeparation) { s.st[i] = st | ISL_POS_SOLVED; d.isl_posIters[body] = iter + 1u; }
    else still = 1u;
    s.acc[i] = float_flip(PR_INF); d.isl_jflag[body] = 0u;
    return still;
}
PR_D void blk_prefetch(const void* p)
{
#ifndef PR_HOST_EMUL
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
PR_D void blk_prefetch_velocity(const RS& r, unsigned i)
{
    blk_prefetch(&r.vc0[i]); blk_prefetch(&r.vc1[i]); blk_prefetch(&r.vc2[i]); blk_prefetch(&r.vc3[i]); blk_prefetch(&r.vc4[i]);
    blk_prefetch(&r.vc5[i]); blk_prefetch(&r.vc6[i]); blk_prefetch(&r.vc7[i]); blk_prefetch(&r.vc8[i]); blk_prefetch(&r.vc_island[i]);
}
PR_D void blk_prefetch_position(const RS& r, unsigned i)
{
    blk_prefetch(&r.vc1[i]); blk_prefetch(&r.vc2[i]); blk_prefetch(&r.vc_island[i]);
    blk_prefetch(&r.pc0[i]); blk_prefetch(&r.pc1[i]); blk_prefetch(&r.pc2[i]); blk_prefetch(&r.pc3[i]);
}
// The rows of every colour in turn (a block barrier after each), then the colour-overflow tail in list order.
// load(row) fetches a row's planes from HBM/L2 into registers, solve(row, loaded) runs it on the block's bodies.
// The tail is a serial chain (a hub's rows all share the hub): one thread walking it paid a dependent ~700-cycle row fetch
// per row (measured: 1.6 ms per world and step). Warp 0 fetches 32 tail rows at once, one per lane, and then runs them
// lane after lane -- the fetches leave the chain, the order stays the list order.
#ifdef PR_TOI_TIMING
#define PR_BT(k) do { if (BLK_ID == 0u && BLK_TID == 0u) { const long long t_ = clock64(); bt_acc[k] += static_cast<unsigned long long>(t_ - bt_last); bt_last = t_; } } while (0)
__device__ unsigned long long bt_acc[12];
__device__ long long bt_last;
#else
#define PR_BT(k) do { } while (0)
#endif
template <class P, class L, class S> PR_D void blk_for_each_row(const BlkView& s, P&& pre, L&& load, S&& solve, int tk = 0)
{
    PR_BT(11);
    for (unsigned c = 0; c < COLOR_OVERFLOW; ++c) {
        const unsigned c0 = s.cstart[c], c1 = s.cstart[c + 1u];
        if (c0 == c1) continue; // (uniform: the whole block skips the barrier)
        // the rows this thread will run in the NEXT colour (or the tail) start their way from HBM to L2 now: a row does not
        // depend on the solve, only the bodies do, and a colour step was ~2 500 cycles of row fetch + ~1 500 of solve
        for (unsigned row = c1 + BLK_TID; row < c1 + BLK_T && row < s.cstart[MAX_COLORS]; row += BLK_T) pre(row);
        for (unsigned row = c0 + BLK_TID; row < c1; row += BLK_T) { auto x = load(row); solve(row, x); }
        BLK_SYNC();
    }
    PR_BT(tk);
    const unsigned t0 = s.cstart[COLOR_OVERFLOW], t1 = s.cstart[COLOR_OVERFLOW + 1u];
    if (t0 != t1) {
        if (BLK_TID < PR_WARP) {
            for (unsigned base = t0; base < t1; base += PR_WARP) {
                const unsigned row = base + BLK_TID;
                const bool have = row < t1;
                auto x = load(have ? row : t0);
                const unsigned n = (t1 - base < PR_WARP) ? t1 - base : PR_WARP;
                for (unsigned l = 0; l < n; ++l) {
                    if (BLK_TID == l) solve(row, x);
#ifndef PR_HOST_EMUL
                    __syncwarp();
#endif
                }
            }
        }
        BLK_SYNC();
    }
    PR_BT(tk + 1);
}

#ifndef PR_HOST_EMUL
__global__ void __launch_bounds__(BLK_MAX_THREADS)
#else
PR_KERNEL
#endif
k_block_solve(DV d, pr_step_conf conf, const unsigned* sortedKey, unsigned maxBodies, unsigned warmStart)
{
    PR_PDL_ENTER();
#ifndef PR_HOST_EMUL
    extern __shared__ __align__(16) unsigned char blk_smem[];
    unsigned char* base = blk_smem;
    __shared__ unsigned cstartBuf[MAX_COLORS + 1u];
#else
    std::vector<unsigned char> blk_buf(static_cast<size_t>(maxBodies) * BLK_BYTES_PER_BODY + 16u);
    unsigned char* base = blk_buf.data();
    unsigned cstartBuf[MAX_COLORS + 1u];
#endif
    BlkView s;
    s.buf = reinterpret_cast<float4*>(base);
    s.acc = reinterpret_cast<unsigned*>(s.buf + maxBodies); s.st = s.acc + maxBodies;
    s.cstart = cstartBuf;
    const unsigned world = BLK_ID;
    const unsigned off = d.blk_off[world], nB = d.blk_off[world + 1u] - off;
    const unsigned r0 = d.blk_rowOff[world], r1 = d.blk_rowOff[world + 1u];
    const unsigned j0 = d.blk_jOff[world], j1 = d.blk_jOff[world + 1u];
    // ---- the world's BodyConstraint velocities (k_isl_prepare wrote them) and island flags come in
    for (unsigned i = BLK_TID; i < nB; i += BLK_T) {
        const unsigned body = d.blk_body[off + i];
        s.buf[i] = d.s_vel[body];
        const unsigned isl = d.b_island[body];
        const bool root = (isl == body);
        s.st[i] = (root ? d.isl_state[body] : 0u) | (isl != INVALID ? static_cast<unsigned>(BLK_ST_ACTIVE) : 0u) | (root ? static_cast<unsigned>(BLK_ST_ROOT) : 0u);
        s.acc[i] = 0u;
    }
    for (unsigned c = BLK_TID; c <= MAX_COLORS; c += BLK_T) s.cstart[c] = INVALID;
    BLK_SYNC();
    for (unsigned row = r0 + BLK_TID; row < r1; row += BLK_T) {
        const unsigned c = sortedKey[row] & 127u;
        if (row == r0 || (sortedKey[row - 1u] & 127u) != c) s.cstart[c] = row;
    }
    BLK_SYNC();
    if (BLK_TID == 0u) { // colours without rows start where the next one does
        s.cstart[MAX_COLORS] = r1;
        for (unsigned c = MAX_COLORS; c-- > 0u;) if (s.cstart[c] == INVALID) s.cstart[c] = s.cstart[c + 1u];
    }
    BLK_SYNC();
    RS r = rs_global(d);
    r.s_vel = s.buf; r.s_pos = s.buf; r.isl_state = s.st; // (one buffer: velocity rows only touch s_vel, position rows only s_pos)
    JView jv = jview_global(d);
    jv.s_vel = s.buf; jv.slotV = d.b_slot; jv.isl_state = s.st; jv.slotI = d.b_slot; // (positions: still the world's array)
    // ---- WarmStartVelocities (:264), then the joints' InitVelocity (:1597)
    if (warmStart) blk_for_each_row(s, [&](unsigned row) { blk_prefetch_velocity(r, row); }, [&](unsigned row) { return vc_load(r, row); }, [&](unsigned, const VCRow& vc) { warm_start_loaded(r, vc); }, 0);
    if (BLK_TID == 0u) for (unsigned j = j0; j < j1; ++j) joint_init_velocity(d, d.blk_joint[j], conf, jv);
    BLK_SYNC();
    // ---- velocity iterations (:1603-1622): joints first, then the contacts
    for (unsigned it = 0; it < conf.reg_velocity_iters; ++it) {
        if (j0 != j1) {
            if (BLK_TID == 0u) for (unsigned j = j0; j < j1; ++j) joint_solve_velocity(d, d.blk_joint[j], conf, jv);
            BLK_SYNC();
        }
        blk_for_each_row(s, [&](unsigned row) { blk_prefetch_velocity(r, row); }, [&](unsigned row) { return vc_load(r, row); }, [&](unsigned row, const VCRow& vc) {
            bool did = false;
            const float inc = solve_velocity_loaded(d, r, row, vc, did, true);
            if (did) atomicMax(&s.acc[vc.island], f2u(inc));
        }, 2);
        unsigned still = 0u;
        for (unsigned i = BLK_TID; i < nB; i += BLK_T) if (s.st[i] & BLK_ST_ROOT) still |= blk_vel_check(d, s, i, d.blk_body[off + i], it, conf.reg_min_momentum);
        if (!BLK_ANY(still)) break; // (every island of this world is done: the remaining iterations would skip every row)
    }
    BLK_SYNC();
    // ---- the velocities go back, the positions come in: IntegratePositions (:176-189)
    for (unsigned i = BLK_TID; i < nB; i += BLK_T) {
        const unsigned body = d.blk_body[off + i];
        const float4 v = s.buf[i];
        d.s_vel[body] = v;
        float4 p = d.s_pos[body];
        if (s.st[i] & BLK_ST_ACTIVE) {
            const float2 tr = conf.delta_time * v2(v.x, v.y);
            const float rot = conf.delta_time * v.z;
            p.x += tr.x; p.y += tr.y; p.z += rot;
        }
        s.buf[i] = p;
        s.acc[i] = float_flip(PR_INF);
    }
    jv.s_vel = d.s_vel; jv.slotV = nullptr; jv.s_pos = s.buf; jv.slotP = d.b_slot;
    BLK_SYNC();
    // ---- position iterations (:1628-1644): the contacts, then the joints
    for (unsigned it = 0; it < conf.reg_position_iters; ++it) {
        struct PosRow { VCRow vch; PCRow pc; };
        blk_for_each_row(s, [&](unsigned row) { blk_prefetch_position(r, row); }, [&](unsigned row) { PosRow x; vc_load_head(r, row, x.vch); x.pc = pc_load(r, row); return x; }, [&](unsigned, const PosRow& x) {
            bool did = false;
            const float sep = solve_position_loaded(d, r, x.vch, x.pc, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction, did, true);
            if (did) atomicMin(&s.acc[x.vch.island], float_flip(sep));
        }, 4);
        if (j0 != j1) {
            if (BLK_TID == 0u) for (unsigned j = j0; j < j1; ++j) joint_solve_position(d, d.blk_joint[j], conf, jv);
            BLK_SYNC();
        }
        unsigned still = 0u;
        for (unsigned i = BLK_TID; i < nB; i += BLK_T) if (s.st[i] & BLK_ST_ROOT) still |= blk_pos_check(d, s, i, d.blk_body[off + i], it, conf.reg_min_separation);
        if (!BLK_ANY(still)) break;
    }
    BLK_SYNC();
    PR_BT(11);
#ifdef PR_TOI_TIMING
    if (BLK_ID == 0u && BLK_TID == 0u) { for (int k = 0; k < 12; ++k) { d.cnt->dbgBlk[k] = bt_acc[k]; bt_acc[k] = 0ull; } d.cnt->dbgBlk[8] = r1 - r0; d.cnt->dbgBlk[9] = s.cstart[COLOR_OVERFLOW + 1u] - s.cstart[COLOR_OVERFLOW]; d.cnt->dbgBlk[10] = nB; }
#endif
    // ---- the positions and island flags go back (k_assign_impulses / k_writeback_bodies / k_sleep_and_stats read them)
    for (unsigned i = BLK_TID; i < nB; i += BLK_T) {
        const unsigned body = d.blk_body[off + i];
        d.s_pos[body] = s.buf[i];
        if (s.st[i] & BLK_ST_ROOT) d.isl_state[body] = s.st[i] & 0xFFu;
    }
}

} // namespace pr
#endif
