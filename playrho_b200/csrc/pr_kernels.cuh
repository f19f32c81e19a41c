// pr_kernels.cuh -- the sm_100a kernels of the step pipeline.
// One thread per element unless noted; every kernel is HBM/latency bound (no dense contraction,
// tensor cores unused by design). Kernels avoid early returns so that full-warp primitives
// (ballot/reduce aggregation of atomics) stay converged.
#ifndef PR_KERNELS_CUH
#define PR_KERNELS_CUH
#include "pr_world.h"
#include "pr_math.cuh"
#include "pr_collide.cuh"
#include "pr_toi.cuh"

namespace pr {

// ------------------------------------------------------------------ warp-aggregated helpers
#ifndef PR_HOST_EMUL
// slot index for threads with pred (all 32 lanes of the warp must call this, converged)
PR_D unsigned warp_append(unsigned* counter, bool pred)
{
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    const unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    if (mask) {
        const int leader = __ffs(mask) - 1;
        if (static_cast<int>(lane) == leader) base = atomicAdd(counter, static_cast<unsigned>(__popc(mask)));
        base = __shfl_sync(0xffffffffu, base, leader);
    }
    return base + __popc(mask & ((1u << lane) - 1u));
}
// opportunistic aggregation usable from divergent code
PR_D unsigned active_append(unsigned* counter)
{
    const unsigned active = __activemask();
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(active) - 1;
    unsigned base = 0;
    if (static_cast<int>(lane) == leader) base = atomicAdd(counter, static_cast<unsigned>(__popc(active)));
    base = __shfl_sync(active, base, leader);
    return base + __popc(active & ((1u << lane) - 1u));
}
PR_D void warp_add(unsigned* counter, unsigned v)
{
    const unsigned s = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(counter, s);
}
PR_D void warp_max(unsigned* counter, unsigned v)
{
    const unsigned s = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) atomicMax(counter, s);
}
PR_D void warp_min(unsigned* counter, unsigned v)
{
    const unsigned s = __reduce_min_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) atomicMin(counter, s);
}
// per-island atomics with a uniform-label fast path (giant islands would otherwise serialise on one
// address); `valid` lanes contribute v to slot[label]. All 32 lanes must call.
PR_D void island_max(unsigned* slot, unsigned label, unsigned v, bool valid)
{
    const unsigned l0 = __shfl_sync(0xffffffffu, label, 0);
    const bool uniform = __all_sync(0xffffffffu, valid && label == l0);
    if (uniform) {
        const unsigned s = __reduce_max_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0) atomicMax(&slot[l0], s);
    } else if (valid) atomicMax(&slot[label], v);
}
PR_D void island_min(unsigned* slot, unsigned label, unsigned v, bool valid)
{
    const unsigned l0 = __shfl_sync(0xffffffffu, label, 0);
    const bool uniform = __all_sync(0xffffffffu, valid && label == l0);
    if (uniform) {
        const unsigned s = __reduce_min_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0) atomicMin(&slot[l0], s);
    } else if (valid) atomicMin(&slot[label], v);
}
// Block-aggregated counter updates: one global atomic per BLOCK (and none when it would not change the counter).
// Every warp of a 1M-thread launch hitting the same few counter words serialises in one L2 slice (measured:
// k_sleep_and_stats 105 us, k_isl_prepare 226 us for 1M bodies). All threads of the block must call, converged.
PR_D void block_add(unsigned* counter, unsigned v)
{
    __shared__ unsigned acc;
    if (threadIdx.x == 0) acc = 0u;
    __syncthreads();
    const unsigned s = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&acc, s);
    __syncthreads();
    if (threadIdx.x == 0 && acc) atomicAdd(counter, acc);
    __syncthreads();
}
PR_D void block_max(unsigned* counter, unsigned v)
{
    __shared__ unsigned acc;
    if (threadIdx.x == 0) acc = 0u;
    __syncthreads();
    const unsigned s = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicMax(&acc, s);
    __syncthreads();
    if (threadIdx.x == 0 && acc > *reinterpret_cast<volatile unsigned*>(counter)) atomicMax(counter, acc); // (counter only grows)
    __syncthreads();
}
PR_D void block_min(unsigned* counter, unsigned v)
{
    __shared__ unsigned acc;
    if (threadIdx.x == 0) acc = 0xFFFFFFFFu;
    __syncthreads();
    const unsigned s = __reduce_min_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s != 0xFFFFFFFFu) atomicMin(&acc, s);
    __syncthreads();
    if (threadIdx.x == 0 && acc < *reinterpret_cast<volatile unsigned*>(counter)) atomicMin(counter, acc); // (counter only shrinks)
    __syncthreads();
}
// per-island add with the lanes of a warp that share a label combined first (a 100k-body island would otherwise
// put 100k atomics on one word). Divergent callers allowed.
PR_D void island_add_matched(unsigned* slot, unsigned label, unsigned v)
{
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, label);
    const unsigned lane = threadIdx.x & 31;
    unsigned sum = 0u;
    for (unsigned m = peers; m; m &= m - 1u) sum += __shfl_sync(peers, v, __ffs(m) - 1);
    if (lane == static_cast<unsigned>(__ffs(peers) - 1)) atomicAdd(&slot[label], sum);
}
#else
PR_D void block_add(unsigned* counter, unsigned v) { *counter += v; }
PR_D void block_max(unsigned* counter, unsigned v) { if (v > *counter) *counter = v; }
PR_D void block_min(unsigned* counter, unsigned v) { if (v < *counter) *counter = v; }
PR_D void island_add_matched(unsigned* slot, unsigned label, unsigned v) { slot[label] += v; }
PR_D unsigned warp_append(unsigned* counter, bool pred) { return pred ? atomicAdd(counter, 1u) : 0u; }
PR_D unsigned active_append(unsigned* counter) { return atomicAdd(counter, 1u); }
PR_D void warp_add(unsigned* counter, unsigned v) { *counter += v; }
PR_D void warp_max(unsigned* counter, unsigned v) { if (v > *counter) *counter = v; }
PR_D void warp_min(unsigned* counter, unsigned v) { if (v < *counter) *counter = v; }
PR_D void island_max(unsigned* slot, unsigned label, unsigned v, bool valid) { if (valid && v > slot[label]) slot[label] = v; }
PR_D void island_min(unsigned* slot, unsigned label, unsigned v, bool valid) { if (valid && v < slot[label]) slot[label] = v; }
#endif

// ------------------------------------------------------------------ small accessors
PR_D bool f_speedable(unsigned f) { return (f & PR_BODY_VELOCITY) != 0; }
PR_D bool f_accelerable(unsigned f) { return (f & PR_BODY_ACCELERATION) != 0; }
PR_D bool f_awake(unsigned f) { return (f & PR_BODY_AWAKE) != 0; }
PR_D bool f_enabled(unsigned f) { return (f & PR_BODY_ENABLED) != 0; }
PR_D ProxyView child_view(const DV& d, unsigned cg)
{
    ProxyView v;
    const float4 c = d.ch_pack[cg]; // (radius, vertex offset, vertex count) in ONE 16-byte gather instead of three
    v.radius = c.x; v.count = f2u(c.z);
    const unsigned off = f2u(c.y);
    v.v = d.vtx + off; v.n = d.nrm + off;
    return v;
}
PR_D Pos pos_from4(float4 p) { Pos r; r.l = v2(p.x, p.y); r.a = p.z; return r; }

// ------------------------------------------------------------------ bodies: AoS rows <-> SoA
PR_KERNEL k_put_bodies(DV d, unsigned n, unsigned firstId, const unsigned* ids, const pr_body* rows, unsigned isUpdate)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const unsigned id = ids ? ids[i] : firstId + i;
        const pr_body r = rows[i];
        if (isUpdate) { // SetBody transform-change path, AabbTreeWorld.cpp:2723-2795
            const float4 o = d.b_xf[id];
            if (o.x != r.xf_px || o.y != r.xf_py || o.z != r.xf_qc || o.w != r.xf_qs) { d.b_sync[id] = 1u; d.b_moved[id] = 1u; }
            const float4 om = d.b_mass[id];
            if ((om.x != 0.0f || om.y != 0.0f) != (r.inv_mass != 0.0f || r.inv_rot_i != 0.0f)) d.cnt->recolorFlag = 1u;
        }
        d.b_xf[id] = make_float4(r.xf_px, r.xf_py, r.xf_qc, r.xf_qs);
        d.b_pos0[id] = make_float4(r.pos0_x, r.pos0_y, r.pos0_a, r.alpha0);
        d.b_pos1[id] = make_float4(r.pos1_x, r.pos1_y, r.pos1_a, r.under_active_time);
        d.b_vel[id] = make_float4(r.vel_x, r.vel_y, r.vel_w, 0.0f);
        d.b_acc[id] = make_float4(r.acc_x, r.acc_y, r.acc_w, 0.0f);
        d.b_mass[id] = make_float4(r.inv_mass, r.inv_rot_i, r.lin_damp, r.ang_damp);
        d.b_lc[id] = make_float2(r.lc_x, r.lc_y);
        d.b_flags[id] = r.flags;
    }
}

PR_KERNEL k_put_accelerations(DV d, unsigned n, const unsigned* ids, const float* acc3)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) d.b_acc[ids ? ids[i] : i] = make_float4(acc3[3u * i], acc3[3u * i + 1u], acc3[3u * i + 2u], 0.0f);
}
// Body::SetVelocity, Body.cpp:184-194
PR_KERNEL k_put_velocities(DV d, unsigned n, const unsigned* ids, const float* vel3)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const unsigned b = ids ? ids[i] : i;
        const float x = vel3[3u * i], y = vel3[3u * i + 1u], w = vel3[3u * i + 2u];
        if (x != 0.0f || y != 0.0f || w != 0.0f) {
            const unsigned f = d.b_flags[b];
            if (!f_speedable(f)) return;
            if (!f_awake(f)) d.b_flags[b] = f | PR_BODY_AWAKE;
            reinterpret_cast<float*>(&d.b_pos1[b])[3] = 0.0f; // m_underActiveTime = 0
        }
        d.b_vel[b] = make_float4(x, y, w, 0.0f);
    }
}
PR_KERNEL k_get_poses(DV d, unsigned first, unsigned n, float4* xf, float* vel3)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        xf[i] = d.b_xf[first + i];
        if (vel3) { const float4 v = d.b_vel[first + i]; vel3[3u * i] = v.x; vel3[3u * i + 1u] = v.y; vel3[3u * i + 2u] = v.z; }
    }
}
PR_KERNEL k_get_bodies(DV d, unsigned first, unsigned n, pr_body* rows)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const unsigned id = first + i;
        pr_body r;
        const float4 xf = d.b_xf[id], p0 = d.b_pos0[id], p1 = d.b_pos1[id], v = d.b_vel[id], a = d.b_acc[id], m = d.b_mass[id];
        const float2 lc = d.b_lc[id];
        r.xf_px = xf.x; r.xf_py = xf.y; r.xf_qc = xf.z; r.xf_qs = xf.w;
        r.pos0_x = p0.x; r.pos0_y = p0.y; r.pos0_a = p0.z; r.alpha0 = p0.w;
        r.pos1_x = p1.x; r.pos1_y = p1.y; r.pos1_a = p1.z; r.under_active_time = p1.w;
        r.lc_x = lc.x; r.lc_y = lc.y;
        r.vel_x = v.x; r.vel_y = v.y; r.vel_w = v.z;
        r.acc_x = a.x; r.acc_y = a.y; r.acc_w = a.z;
        r.inv_mass = m.x; r.inv_rot_i = m.y; r.lin_damp = m.z; r.ang_damp = m.w;
        r.flags = d.b_flags[id];
        rows[i] = r;
    }
}

// ------------------------------------------------------------------ a2: proxy AABB + fat-box policy
struct Box { float minx, miny, maxx, maxy; };
PR_D Box box_from4(float4 f) { Box b; b.minx = f.x; b.miny = f.y; b.maxx = f.z; b.maxy = f.w; return b; }
PR_D float4 box_to4(Box b) { return make_float4(b.minx, b.miny, b.maxx, b.maxy); }
// ComputeAABB(proxy, xf0, xf1), AABB.cpp:51-62 (+ Interval::Include, GetFattenedAABB(vertexRadius))
PR_D Box compute_aabb(const ProxyView& p, Xf xf0, Xf xf1)
{
    Box r; r.minx = PR_INF; r.miny = PR_INF; r.maxx = -PR_INF; r.maxy = -PR_INF;
    for (unsigned i = 0; i < p.count; ++i) {
        const float2 v = p.vertex(i);
        const float2 a = transform(v, xf0);
        r.minx = stdmin(a.x, r.minx); r.maxx = stdmax(a.x, r.maxx); r.miny = stdmin(a.y, r.miny); r.maxy = stdmax(a.y, r.maxy);
        const float2 b = transform(v, xf1);
        r.minx = stdmin(b.x, r.minx); r.maxx = stdmax(b.x, r.maxx); r.miny = stdmin(b.y, r.miny); r.maxy = stdmax(b.y, r.maxy);
    }
    r.minx -= p.radius; r.maxx += p.radius; r.miny -= p.radius; r.maxy += p.radius;
    return r;
}
// GetFattenedAABB + GetDisplacedAABB, detail/AABB.hpp:305-320; Interval::Expand adds d to min if d<0 else max
PR_D Box fatten_displace(Box a, float ext, float2 dsp)
{
    a.minx -= ext; a.maxx += ext; a.miny -= ext; a.maxy += ext;
    if (dsp.x < 0.0f) a.minx += dsp.x; else a.maxx += dsp.x;
    if (dsp.y < 0.0f) a.miny += dsp.y; else a.maxy += dsp.y;
    return a;
}
PR_D bool box_contains(Box a, Box b) { return a.minx <= b.minx && a.maxx >= b.maxx && a.miny <= b.miny && a.maxy >= b.maxy; }
// closed-interval overlap: min(maxs) >= max(mins) per axis, Interval.hpp IsIntersecting
PR_D bool box_overlap(Box a, Box b)
{
    return (stdmin(a.maxx, b.maxx) >= stdmax(a.minx, b.minx)) && (stdmin(a.maxy, b.maxy) >= stdmax(a.miny, b.miny));
}

// CreateProxies (AabbTreeWorld.cpp:557-579) for proxies [first, first+n): fat box + queue for FindContacts
PR_KERNEL k_create_proxy_aabbs(DV d, pr_step_conf conf, unsigned first, unsigned n)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const unsigned px = first + i;
        const unsigned body = d.p_body[px];
        const Xf xf1 = xf_from4(d.b_xf[body]);
        const Xf xf0 = get_xf(pos_from4(d.b_pos0[body]), d.b_lc[body]);
        const float2 dsp = conf.displace_multiplier * (xf1.p - xf0.p);
        const Box a = fatten_displace(compute_aabb(child_view(d, d.p_cg[px]), xf0, xf1), conf.aabb_extension, dsp);
        d.p_fat[px] = box_to4(a);
        d.p_moved[px] = 1u;
        d.q_list[d.cnt->nQueue + i] = px; // host bumps nQueue afterwards
    }
}

// Synchronize (AabbTreeWorld.cpp:2564-2585) for every proxy whose body is flagged in b_sync
PR_KERNEL k_sync_proxies(DV d, pr_step_conf conf, unsigned nProxies)
{
    PR_PDL_ENTER();
    const unsigned px = PR_TID;
    bool moved = false;
    if (px < nProxies) {
        const unsigned body = d.p_body[px];
        if (d.b_sync[body]) {
            const Xf xf1 = xf_from4(d.b_xf[body]);
            const Xf xf0 = get_xf(pos_from4(d.b_pos0[body]), d.b_lc[body]);
            const Box a = compute_aabb(child_view(d, d.p_cg[px]), xf0, xf1);
            if (!box_contains(box_from4(d.p_fat[px]), a)) {
                const float2 dsp = conf.displace_multiplier * (xf1.p - xf0.p);
                d.p_fat[px] = box_to4(fatten_displace(a, conf.aabb_extension, dsp));
                moved = (d.p_moved[px] == 0u); // queue once
                d.p_moved[px] = 1u;
                if (!moved) atomicAdd(&d.cnt->proxiesMoved, 1u); // re-moved while still queued: count, do not re-queue
            }
        }
    }
    const unsigned slot = warp_append(&d.cnt->nQueue, moved);
    if (moved) { d.q_list[slot] = px; }
    warp_add(&d.cnt->proxiesMoved, moved ? 1u : 0u);
}

PR_KERNEL k_clear_queue_flags(DV d, unsigned nQueue)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nQueue) d.p_moved[d.q_list[i]] = 0u;
}

// ------------------------------------------------------------------ contact-key hash set
PR_D unsigned long long mix64(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
PR_D void hash_insert(const DV& d, unsigned long long key)
{
    unsigned h = static_cast<unsigned>(mix64(key)) & d.ht_mask;
    for (unsigned probe = 0; probe <= d.ht_mask; ++probe) {
        const unsigned long long old = atomicCAS(&d.ht_keys[h], ~0ull, key);
        if (old == ~0ull || old == key) return;
        h = (h + 1u) & d.ht_mask;
    }
    atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_HASH_FULL));
}
// removal leaves a tombstone (probe chains stay intact); the host rebuilds the table when live + dead > half
#define PR_HASH_TOMB (~0ull - 1ull)
PR_D void hash_remove(const DV& d, unsigned long long key)
{
    unsigned h = static_cast<unsigned>(mix64(key)) & d.ht_mask;
    for (unsigned probe = 0; probe <= d.ht_mask; ++probe) {
        const unsigned long long cur = d.ht_keys[h];
        if (cur == key) { d.ht_keys[h] = PR_HASH_TOMB; return; }
        if (cur == ~0ull) return;
        h = (h + 1u) & d.ht_mask;
    }
}
PR_D bool hash_contains(const DV& d, unsigned long long key)
{
    unsigned h = static_cast<unsigned>(mix64(key)) & d.ht_mask;
    for (unsigned probe = 0; probe <= d.ht_mask; ++probe) {
        const unsigned long long cur = d.ht_keys[h];
        if (cur == key) return true;
        if (cur == ~0ull) return false;
        h = (h + 1u) & d.ht_mask;
    }
    return false;
}
PR_KERNEL k_hash_fill(DV d, unsigned nOrder)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) hash_insert(d, d.c_key[d.c_order[i]]);
}
// SolveReg's maxIslandBodies (AabbTreeWorld.cpp:1510) counts a body that cannot move once for every island that reaches it
// (AddToIsland pushes it, does not propagate through it, and RemoveUnspeedablesFromIslanded frees it for the next island,
// :1395-1403): the first thread to insert (island root, body) into the step's pair set counts it. A full probe window is
// reported (islStatOverflow) and the host recounts from the constraint rows with a larger set (recount_island_statics).
PR_D void island_count_static(const DV& d, unsigned root, unsigned body)
{
    const unsigned long long key = (static_cast<unsigned long long>(root) << 32) | body;
    unsigned h = static_cast<unsigned>(mix64(key)) & d.is_mask;
    for (unsigned probe = 0; probe < 64u; ++probe) {
        const unsigned long long old = atomicCAS(&d.is_keys[h], ~0ull, key);
        if (old == ~0ull) { atomicAdd(&d.isl_nstat[root], 1u); return; }
        if (old == key) return;
        h = (h + 1u) & d.is_mask;
    }
    d.cnt->islStatOverflow = 1u;
}

// ShiftOrigin(AabbTreeWorld&, newOrigin), AabbTreeWorld.cpp:2221-2243: sweeps move by -newOrigin and SetSweep recomputes the
// transformation from position 1 (Body.hpp:697-702); b_moved stands for FlagForUpdating of the body's contacts (the next
// pre-phase UpdateContacts reads it); tree boxes move by -newOrigin (DynamicTree.cpp:650-658, GetMovedAABB = min/max + value).
PR_KERNEL k_shift_origin(DV d, unsigned nBodies, unsigned nProxies, float ox, float oy)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) {
        float4 p0 = d.b_pos0[i], p1 = d.b_pos1[i];
        p0.x -= ox; p0.y -= oy; p1.x -= ox; p1.y -= oy;
        d.b_pos0[i] = p0; d.b_pos1[i] = p1;
        const Xf xf = get_xf(pos_from4(p1), d.b_lc[i]);
        d.b_xf[i] = make_float4(xf.p.x, xf.p.y, xf.q.c, xf.q.s);
        d.b_moved[i] = 1u;
    }
    if (i < nProxies) {
        float4 f = d.p_fat[i];
        const float nx = -ox, ny = -oy;
        f.x += nx; f.y += ny; f.z += nx; f.w += ny;
        d.p_fat[i] = f;
    }
}

// ------------------------------------------------------------------ a6: DestroyContacts
PR_D void body_set_awake(const DV& d, unsigned b)
{ // Body::SetAwake, Body.cpp:163-171
    if (f_speedable(d.b_flags[b])) {
        atomicOr(&d.b_flags[b], static_cast<unsigned>(PR_BODY_AWAKE));
        reinterpret_cast<float*>(&d.b_pos1[b])[3] = 0.0f; // underActiveTime
    }
}
PR_D void manifold_clear(const DV& d, unsigned id)
{
    d.m0[id] = make_float4(0, 0, 0, 0); d.m1[id] = make_float4(0, 0, 0, 0); d.m2[id] = make_float4(0, 0, 0, 0);
    d.m3[id] = make_uint4(PR_CF_DEFAULT, PR_CF_DEFAULT, 0u, 0u);
}
// Contact events (pr_events): what the begin/end listeners (AabbTreeWorld.cpp:2249-2252, :2703-2714) and the post-solve listener
// (Report :196-206) would have been called with, appended to device buffers; the host reads them after the step.
PR_D void emit_contact_event(const DV& d, unsigned kind, unsigned id)
{ // (re-reads the contact's bodies: the callers need not keep them in registers across the collision for this rare path)
    const unsigned slot = atomicAdd(&d.cnt->nEvContact, 1u);
    if (slot < d.ev_cap) {
        const uint4 ab = d.c_ab[id], ids = d.c_ids[id];
        pr_contact_event e;
        e.kind = kind; e.contact = id; e.body_a = ab.x; e.shape_a = ids.x; e.child_a = ids.y; e.body_b = ab.y; e.shape_b = ids.z; e.child_b = ids.w;
        d.ev_contact[slot] = e;
    }
}
PR_D void emit_impulse_event(const DV& d, unsigned id, unsigned count, unsigned iterations, unsigned phase, const float4& imp)
{
    const unsigned slot = atomicAdd(&d.cnt->nEvImpulse, 1u);
    if (slot < d.ev_cap) {
        pr_impulse_event e;
        e.contact = id; e.count = count; e.iterations = iterations; e.phase = phase;
        e.normal[0] = count > 0u ? imp.x : 0.0f; e.tangent[0] = count > 0u ? imp.y : 0.0f;
        e.normal[1] = count > 1u ? imp.z : 0.0f; e.tangent[1] = count > 1u ? imp.w : 0.0f;
        d.ev_impulse[slot] = e;
    }
}
PR_D void emit_presolve_event(const DV& d, unsigned id)
{ // preSolveContact(contactID, oldManifold), AabbTreeWorld.cpp:2716-2720: called BEFORE the new manifold is stored
    const unsigned slot = atomicAdd(&d.cnt->nEvPre, 1u);
    if (slot < d.ev_cap) {
        const float4 a = d.m0[id], b = d.m1[id], c = d.m2[id];
        const uint4 e = d.m3[id];
        pr_presolve_event ev;
        ev.contact = id; ev.pad = 0u;
        ev.old_manifold.ln_x = a.x; ev.old_manifold.ln_y = a.y; ev.old_manifold.lp_x = a.z; ev.old_manifold.lp_y = a.w;
        ev.old_manifold.p0_x = b.x; ev.old_manifold.p0_y = b.y; ev.old_manifold.p1_x = b.z; ev.old_manifold.p1_y = b.w;
        ev.old_manifold.jn0 = c.x; ev.old_manifold.jt0 = c.y; ev.old_manifold.jn1 = c.z; ev.old_manifold.jt1 = c.w;
        ev.old_manifold.cf0 = e.x; ev.old_manifold.cf1 = e.y; ev.old_manifold.type = e.z; ev.old_manifold.count = e.w;
        d.ev_presolve[slot] = ev;
    }
}
// keep[i] = 1 if the two FAT leaf boxes still overlap (AabbTreeWorld.cpp:2290-2304), else InternalDestroy (:2245)
PR_KERNEL k_destroy_test(DV d, unsigned nOrder, unsigned* keep, unsigned hashLive)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const uint2 px = d.c_px[id];
        const bool ov = box_overlap(box_from4(d.p_fat[px.x]), box_from4(d.p_fat[px.y]));
        keep[i] = ov ? 1u : 0u;
        if (!ov) {
            const uint4 ab = d.c_ab[id];
            const unsigned flags = d.c_flags[id];
            if ((d.ev_mask & PR_EVENT_END) && (flags & PR_CONTACT_TOUCHING)) emit_contact_event(d, PR_EVENT_END, id); // :2249-2252
            if (d.m3[id].w > 0u && !(flags & PR_CONTACT_SENSOR)) { body_set_awake(d, ab.x); body_set_awake(d, ab.y); }
            d.c_flags[id] = PR_CONTACT_DESTROYED;
            if (hashLive) hash_remove(d, d.c_key[id]);
            d.c_key[id] = ~0ull;
            manifold_clear(d, id);
        }
    }
}
// stable compaction: kept ids to newOrder[pos[i]]; destroyed ids pushed on the free stack in order (ObjectPool::Free)
PR_KERNEL k_destroy_compact(DV d, unsigned nOrder, const unsigned* keep, const unsigned* pos, unsigned* newOrder, unsigned nFree)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        if (keep[i]) newOrder[pos[i]] = id;
        else d.c_free[nFree + (i - pos[i])] = id;
    }
}

// ------------------------------------------------------------------ a7: narrowphase Update
PR_D void store_manifold(const DV& d, unsigned id, const Mf& m)
{
    d.m0[id] = make_float4(m.ln.x, m.ln.y, m.lp.x, m.lp.y);
    d.m1[id] = make_float4(m.p0.x, m.p0.y, m.p1.x, m.p1.y);
    d.m2[id] = make_float4(m.jn0, m.jt0, m.jn1, m.jt1);
    d.m3[id] = make_uint4(m.cf0, m.cf1, m.type, m.count);
}
// Update(ContactID), AabbTreeWorld.cpp:2587-2721: collide (or sensor TestOverlap), carry impulses over by
// ContactFeature equality else nearest old point, touching flag. Called with the contact's flags word.
// (pvA / pvB: the two children's vertex sets -- the shape pool in HBM, or a block's shared-memory copy of them: k_update_contacts_tma)
PR_D void update_contact_v(const DV& d, const pr_step_conf& conf, unsigned id, unsigned flags, const uint4 ab, const ProxyView& pvA, const ProxyView& pvB);
PR_D void update_contact(const DV& d, const pr_step_conf& conf, unsigned id, unsigned flags)
{
    const uint4 ab = d.c_ab[id];
    update_contact_v(d, conf, id, flags, ab, child_view(d, ab.z), child_view(d, ab.w));
}
PR_D void update_contact_v(const DV& d, const pr_step_conf& conf, unsigned id, unsigned flags, const uint4 ab, const ProxyView& pvA, const ProxyView& pvB)
{
    const bool oldTouching = (flags & PR_CONTACT_TOUCHING) != 0;
    bool newTouching = false;
    const Xf xfA = xf_from4(d.b_xf[ab.x]), xfB = xf_from4(d.b_xf[ab.y]);
    if (flags & PR_CONTACT_SENSOR) { // TestOverlap(...) >= 0, AabbTreeWorld.cpp:2622-2626
        newTouching = gjk_test_overlap(pvA, xfA, pvB, xfB, conf.max_distance_iters) >= 0.0f;
        manifold_clear(d, id);
    } else {
        // the old manifold is fetched BEFORE the collision (it does not depend on it): its round trip overlaps the
        // vertex fetches instead of following them (a touching contact will need the points and impulses too)
        const uint4 o3 = d.m3[id];
        float4 o1 = make_float4(0, 0, 0, 0), o2 = make_float4(0, 0, 0, 0);
        if (oldTouching) { o1 = d.m1[id]; o2 = d.m2[id]; }
        Mf nm = collide_shapes(pvA, xfA, pvB, xfB, conf.linear_slop, conf.max_circles_ratio);
        const unsigned oldCount = o3.w, newCount = nm.count;
        newTouching = newCount > 0u;
        if (newCount > 0u && oldCount > 0u) {
            if (!oldTouching) { o1 = d.m1[id]; o2 = d.m2[id]; } // (old points without the touching flag: not produced by Update, kept for safety)
            const unsigned ocf[2] = {o3.x, o3.y};
            const float ojn[2] = {o2.x, o2.z}, ojt[2] = {o2.y, o2.w};
            const float2 op[2] = {v2(o1.x, o1.y), v2(o1.z, o1.w)};
            const unsigned ncf[2] = {nm.cf0, nm.cf1};
            const float2 np[2] = {nm.p0, nm.p1};
            float njn[2] = {0.0f, 0.0f}, njt[2] = {0.0f, 0.0f};
            bool found[2] = {false, newCount < 2u};
            for (unsigned a = 0; a < newCount; ++a) {
                for (unsigned b = 0; b < oldCount; ++b) {
                    if (ncf[a] == ocf[b]) { found[a] = true; njn[a] = ojn[b]; njt[a] = ojt[b]; break; }
                }
            }
            for (unsigned a = 0; a < newCount; ++a) {
                if (!found[a]) {
                    float least = PR_INF;
                    for (unsigned b = 0; b < oldCount; ++b) {
                        const float sq = magsq(op[b] - np[a]);
                        if (least > sq) { least = sq; njn[a] = ojn[b]; njt[a] = ojt[b]; }
                    }
                }
            }
            nm.jn0 = njn[0]; nm.jt0 = njt[0]; nm.jn1 = njn[1]; nm.jt1 = njt[1];
        }
        // an empty manifold over an empty manifold (3 of 4 contacts of a pile are fat-box overlaps only) is the same
        // 64 bytes again (mf_unset() == manifold_clear()): skip the stores
        if ((d.ev_mask & PR_EVENT_PRE_SOLVE) && newTouching) emit_presolve_event(d, id); // (the old manifold is still in place)
        const bool sameEmpty = newCount == 0u && oldCount == 0u && o3.x == PR_CF_DEFAULT && o3.y == PR_CF_DEFAULT && o3.z == MF_UNSET;
        if (!sameEmpty) store_manifold(d, id, nm);
    }
    const unsigned oldFlags = flags;
    flags &= ~static_cast<unsigned>(PR_CONTACT_DIRTY);
    if (!oldTouching && newTouching) flags |= PR_CONTACT_TOUCHING;
    else if (oldTouching && !newTouching) flags &= ~static_cast<unsigned>(PR_CONTACT_TOUCHING);
    if (flags != oldFlags) d.c_flags[id] = flags;
    if (d.ev_mask && oldTouching != newTouching) { // beginContact / endContact, :2703-2714
        const unsigned kind = newTouching ? PR_EVENT_BEGIN : PR_EVENT_END;
        if (d.ev_mask & kind) emit_contact_event(d, kind, id);
    }
}
// UpdateContacts (AabbTreeWorld.cpp:2333-2421): one thread per contact of order[start, start+n).
// A warp's 32 contacts mix disk-disk, disk-polygon and polygon-polygon pairs, and it used to execute the three paths one
// after another (ncu: 4.9 of 32 lanes active; every warp paid for the O(vA x vB) polygon search). The block now deals its
// 256 contacts out again by class -- a counting sort in shared memory, stable, so neighbouring lanes still hold
// neighbouring contact ids and the global accesses keep their locality (a GLOBAL sort by class lost to sector waste).
#if !defined(PR_HOST_EMUL)
#ifndef PR_NARROW_MINBLOCKS
#define PR_NARROW_MINBLOCKS 4 // latency-bound on its gather chain: 64 registers / 4 CTAs per SM measured 0.77 -> 0.65 ms (5 and 6: 0.68)
#endif
__global__ void __launch_bounds__(PR_BLOCK, PR_NARROW_MINBLOCKS)
#else
PR_KERNEL
#endif
k_update_contacts(DV d, pr_step_conf conf, unsigned start, unsigned n, unsigned useMoved, unsigned countIt)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    bool dirty = false;
    unsigned id = 0u, flags = 0u;
    if (i < n) {
        id = d.c_order[start + i];
        flags = d.c_flags[id];
        dirty = (flags & PR_CONTACT_DIRTY) != 0;
        if (useMoved) { const uint4 ab = d.c_ab[id]; dirty = dirty || (d.b_moved[ab.x] | d.b_moved[ab.y]) != 0u; }
    }
#ifndef PR_HOST_EMUL
    constexpr unsigned NCLS = 4u; // 0 point-point, 1 point-polygon, 2 polygon-polygon / sensor, 3 nothing to do
    __shared__ unsigned s_id[PR_BLOCK], s_flags[PR_BLOCK];
    __shared__ unsigned s_warp[PR_BLOCK / 32][NCLS], s_base[NCLS + 1u];
    unsigned cls = 3u;
    if (dirty) {
        const uint4 ab = d.c_ab[id];
        const unsigned ca = f2u(d.ch_pack[ab.z].z), cb = f2u(d.ch_pack[ab.w].z);
        cls = (flags & PR_CONTACT_SENSOR) ? 2u : (ca == 1u && cb == 1u) ? 0u : (ca == 1u || cb == 1u) ? 1u : 2u;
    }
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    unsigned rank = 0u;
#pragma unroll
    for (unsigned c = 0; c < NCLS; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if (cls == c) rank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0u) s_warp[wid][c] = __popc(m);
    }
    __syncthreads();
    if (threadIdx.x == 0u) { // (32 numbers: one thread turns the per-warp counts into class-major offsets)
        unsigned run = 0u;
        for (unsigned c = 0; c < NCLS; ++c) {
            s_base[c] = run;
            for (unsigned w = 0; w < blockDim.x / 32u; ++w) { const unsigned t = s_warp[w][c]; s_warp[w][c] = run; run += t; }
        }
        s_base[NCLS] = run;
    }
    __syncthreads();
    const unsigned slot = s_warp[wid][cls] + rank;
    s_id[slot] = id; s_flags[slot] = flags;
    __syncthreads();
    if (threadIdx.x < s_base[3]) update_contact(d, conf, s_id[threadIdx.x], s_flags[threadIdx.x]);
#else
    if (dirty) update_contact(d, conf, id, flags);
#endif
    block_add(&d.cnt->contactsUpdated, (dirty && countIt) ? 1u : 0u);
}

#ifndef PR_HOST_EMUL
// ---- TMA staging of the vertex sets (BASELINE north_star (2); opt-in, debug bit 6 -- measured) ----------------
// Same kernel, but before the collision every thread whose pair has a polygon / edge child has the bulk-copy engine
// (cp.async.bulk, the non-tensor TMA path: UBLKCP in SASS) bring that child's vertices and normals from the shape pool into the
// block's shared memory, completion counted on one mbarrier; collide_shapes then reads shared memory. A run starts on an
// 8-byte boundary in the pool and the engine wants 16: the copy starts at the aligned address below it and the view skips the
// first element where needed (the pool has slack at its end).
constexpr unsigned TMA_STAGE_BYTES = 40960u;
PR_D unsigned tma_run_bytes(unsigned voff, unsigned count) { return (((voff & 1u) + count) * 8u + 15u) & ~15u; }
PR_D unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__global__ void __launch_bounds__(PR_BLOCK, 3)
k_update_contacts_tma(DV d, pr_step_conf conf, unsigned start, unsigned n, unsigned useMoved, unsigned countIt)
{
    PR_PDL_ENTER();
    extern __shared__ __align__(128) unsigned char tma_stage[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned i = PR_TID;
    bool dirty = false;
    unsigned id = 0u, flags = 0u;
    if (i < n) {
        id = d.c_order[start + i];
        flags = d.c_flags[id];
        dirty = (flags & PR_CONTACT_DIRTY) != 0;
        if (useMoved) { const uint4 ab = d.c_ab[id]; dirty = dirty || (d.b_moved[ab.x] | d.b_moved[ab.y]) != 0u; }
    }
    constexpr unsigned NCLS = 4u;
    __shared__ unsigned s_id[PR_BLOCK], s_flags[PR_BLOCK];
    __shared__ unsigned s_warp[PR_BLOCK / 32][NCLS], s_base[NCLS + 1u], s_scan[PR_BLOCK / 32];
    unsigned cls = 3u;
    if (dirty) {
        const uint4 ab = d.c_ab[id];
        const unsigned ca = f2u(d.ch_pack[ab.z].z), cb = f2u(d.ch_pack[ab.w].z);
        cls = (flags & PR_CONTACT_SENSOR) ? 2u : (ca == 1u && cb == 1u) ? 0u : (ca == 1u || cb == 1u) ? 1u : 2u;
    }
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    unsigned rank = 0u;
#pragma unroll
    for (unsigned c = 0; c < NCLS; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if (cls == c) rank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0u) s_warp[wid][c] = __popc(m);
    }
    if (threadIdx.x == 0u) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(static_cast<unsigned>(PR_BLOCK)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0u) {
        unsigned run = 0u;
        for (unsigned c = 0; c < NCLS; ++c) {
            s_base[c] = run;
            for (unsigned w = 0; w < blockDim.x / 32u; ++w) { const unsigned t = s_warp[w][c]; s_warp[w][c] = run; run += t; }
        }
        s_base[NCLS] = run;
    }
    __syncthreads();
    const unsigned slot = s_warp[wid][cls] + rank;
    s_id[slot] = id; s_flags[slot] = flags;
    __syncthreads();
    // ---- thread t now owns the t-th dirty contact of the block (by class): what must be staged for it
    const bool work = threadIdx.x < s_base[3];
    unsigned myId = 0u, myFlags = 0u, need = 0u;
    uint4 ab = make_uint4(0u, 0u, 0u, 0u);
    unsigned voff[2] = {0u, 0u}, cnt[2] = {0u, 0u};
    float rad[2] = {0.0f, 0.0f};
    if (work) {
        myId = s_id[threadIdx.x]; myFlags = s_flags[threadIdx.x];
        ab = d.c_ab[myId];
        const unsigned cg[2] = {ab.z, ab.w};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            { const float4 c = d.ch_pack[cg[k]]; voff[k] = f2u(c.y); cnt[k] = f2u(c.z); rad[k] = c.x; }
            if (cnt[k] >= 2u) need += 2u * tma_run_bytes(voff[k], cnt[k]); // vertices + normals
        }
    }
    // block-wide exclusive scan of the staging sizes
    unsigned incl = need;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= static_cast<unsigned>(o)) incl += t; }
    if (lane == 31u) s_scan[wid] = incl;
    __syncthreads();
    unsigned off = incl - need;
    for (unsigned w = 0; w < wid; ++w) off += s_scan[w];
    const bool staged = need != 0u && off + need <= TMA_STAGE_BYTES;
    const unsigned barAddr = smem_u32(&bar);
    ProxyView pv[2];
    if (staged) {
        unsigned long long st;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(st) : "r"(barAddr), "r"(need) : "memory");
        (void)st;
        unsigned cur = off;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            pv[k].radius = rad[k]; pv[k].count = cnt[k];
            if (cnt[k] >= 2u) {
                const unsigned bytes = tma_run_bytes(voff[k], cnt[k]);
                const float2* gv = d.vtx + (voff[k] & ~1u);
                const float2* gn = d.nrm + (voff[k] & ~1u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tma_stage + cur)), "l"(gv), "r"(bytes), "r"(barAddr) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tma_stage + cur + bytes)), "l"(gn), "r"(bytes), "r"(barAddr) : "memory");
                pv[k].v = reinterpret_cast<const float2*>(tma_stage + cur) + (voff[k] & 1u);
                pv[k].n = reinterpret_cast<const float2*>(tma_stage + cur + bytes) + (voff[k] & 1u);
                cur += 2u * bytes;
            } else { pv[k].v = d.vtx + voff[k]; pv[k].n = d.nrm + voff[k]; }
        }
    } else {
        unsigned long long st;
        asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(barAddr) : "memory");
        (void)st;
#pragma unroll
        for (int k = 0; k < 2; ++k) { pv[k].radius = rad[k]; pv[k].count = cnt[k]; pv[k].v = d.vtx + voff[k]; pv[k].n = d.nrm + voff[k]; }
    }
    { // every thread arrived once; the phase completes when the engine has delivered every expected byte
        unsigned done = 0u;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(barAddr) : "memory");
    }
    if (work) update_contact_v(d, conf, myId, myFlags, ab, pv[0], pv[1]);
    block_add(&d.cnt->contactsUpdated, (dirty && countIt) ? 1u : 0u);
}
#endif

// Binned narrowphase (the regular phase's UpdateContacts, where nearly every contact is dirty): a warp of the plain
// kernel mixes disk-disk, disk-polygon and k-gon/m-gon pairs and executes them one after another (measured: 4.9 of
// 32 lanes active). k_narrow_keys classifies every dirty contact by shape-pair class and vertex counts into an 8-bit
// key, ONE stable radix pass groups equal keys (ids stay ascending inside a group, so the gathers stay local), and
// k_update_contacts_binned runs the same update_contact on homogeneous warps. Results are per contact: order-free.
PR_KERNEL k_narrow_keys(DV d, unsigned nOrder, unsigned useMoved, unsigned* keys, unsigned* vals)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    bool dirty = false;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const unsigned flags = d.c_flags[id];
        const uint4 ab = d.c_ab[id];
        dirty = (flags & PR_CONTACT_DIRTY) != 0;
        if (useMoved) dirty = dirty || (d.b_moved[ab.x] | d.b_moved[ab.y]) != 0u;
        unsigned key = 0xFFu; // not dirty: sorted to the end, skipped
        if (dirty) {
            unsigned ca = f2u(d.ch_pack[ab.z].z), cb = f2u(d.ch_pack[ab.w].z);
            ca = ca > 8u ? 8u : ca; cb = cb > 8u ? 8u : cb;
            if (flags & PR_CONTACT_SENSOR) key = 0xC0u;
            else if (ca == 1u && cb == 1u) key = 0u;
            else if (ca == 1u || cb == 1u) key = 0x40u | ((ca > cb ? ca : cb) - 1u);
            else key = 0x80u | ((ca - 1u) << 3) | (cb - 1u);
        }
        keys[i] = key; vals[i] = id;
    }
    block_add(&d.cnt->contactsUpdated, dirty ? 1u : 0u);
}
PR_KERNEL k_update_contacts_binned(DV d, pr_step_conf conf, unsigned n, const unsigned* keys, const unsigned* vals)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n && keys[i] != 0xFFu) { const unsigned id = vals[i]; update_contact(d, conf, id, d.c_flags[id]); }
}

// ------------------------------------------------------------------ a3: LBVH (Morton + Karras build + refit)
PR_KERNEL k_bvh_bounds(DV d, unsigned nProxies)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned cx0 = 0xFFFFFFFFu, cy0 = 0xFFFFFFFFu, cx1 = 0u, cy1 = 0u;
    if (i < nProxies) {
        const float4 f = d.p_fat[i];
        const float cx = (f.x + f.z) * 0.5f, cy = (f.y + f.w) * 0.5f;
        cx0 = cx1 = float_flip(cx); cy0 = cy1 = float_flip(cy);
    }
    warp_min(&d.cnt->boundsMinX, cx0); warp_min(&d.cnt->boundsMinY, cy0);
    warp_max(&d.cnt->boundsMaxX, cx1); warp_max(&d.cnt->boundsMaxY, cy1);
}
PR_D unsigned expand_bits16(unsigned v)
{
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu; v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
    return v;
}
PR_KERNEL k_bvh_morton(DV d, unsigned nProxies)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nProxies) {
        const float x0 = float_unflip(d.cnt->boundsMinX), y0 = float_unflip(d.cnt->boundsMinY);
        const float x1 = float_unflip(d.cnt->boundsMaxX), y1 = float_unflip(d.cnt->boundsMaxY);
        const float4 f = d.p_fat[i];
        const float cx = (f.x + f.z) * 0.5f, cy = (f.y + f.w) * 0.5f;
        const float sx = (x1 > x0) ? 65535.0f / (x1 - x0) : 0.0f, sy = (y1 > y0) ? 65535.0f / (y1 - y0) : 0.0f;
        const unsigned qx = static_cast<unsigned>(fminf(fmaxf((cx - x0) * sx, 0.0f), 65535.0f));
        const unsigned qy = static_cast<unsigned>(fminf(fmaxf((cy - y0) * sy, 0.0f), 65535.0f));
        unsigned code = expand_bits16(qx) | (expand_bits16(qy) << 1);
        // batch arena: the world index is the key's prefix, so every world is its own subtree
        if (d.nWorlds) code = (d.b_world[d.p_body[i]] << (32u - d.worldBits)) | (code >> d.worldBits);
        d.bv_code[i] = code;
        d.bv_idx[i] = i;
    }
}
// common-prefix length of sorted keys i and j (index as tie-break), -1 if j out of range
PR_D int bvh_delta(const unsigned* codes, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned a = codes[i], b = codes[j];
    if (a == b) return 32 + __clz(static_cast<unsigned>(i) ^ static_cast<unsigned>(j));
    return __clz(a ^ b);
}
// Karras 2012: one thread per internal node i in [0, n-1)
PR_KERNEL k_bvh_build(DV d, unsigned nLeaves, const unsigned* codes)
{
    PR_PDL_ENTER();
    const int i = static_cast<int>(PR_TID);
    const int n = static_cast<int>(nLeaves);
    if (i < n - 1) {
        const int dir = (bvh_delta(codes, n, i, i + 1) - bvh_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = bvh_delta(codes, n, i, i - dir);
        int lmax = 2;
        while (bvh_delta(codes, n, i, i + lmax * dir) > dmin) lmax *= 2;
        int l = 0;
        for (int t = lmax / 2; t >= 1; t /= 2) if (bvh_delta(codes, n, i, i + (l + t) * dir) > dmin) l += t;
        const int j = i + l * dir;
        const int dnode = bvh_delta(codes, n, i, j);
        int s = 0;
        for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
            if (bvh_delta(codes, n, i, i + (s + t) * dir) > dnode) s += t;
            if (t == 1) break;
        }
        const int gamma = i + s * dir + (dir < 0 ? -1 : 0);
        const int lo = (i < j) ? i : j, hi = (i < j) ? j : i;
        const unsigned left = (lo == gamma) ? (static_cast<unsigned>(gamma) | LEAF_BIT) : static_cast<unsigned>(gamma);
        const unsigned right = (hi == gamma + 1) ? (static_cast<unsigned>(gamma + 1) | LEAF_BIT) : static_cast<unsigned>(gamma + 1);
        d.bv_left[i] = left; d.bv_right[i] = right;
        if (left & LEAF_BIT) d.bv_lparent[left & ~LEAF_BIT] = i; else d.bv_parent[left] = i;
        if (right & LEAF_BIT) d.bv_lparent[right & ~LEAF_BIT] = i; else d.bv_parent[right] = i;
        d.bv_visit[i] = 0u;
        if (d.nWorlds) d.bv_wr[i] = make_uint2(codes[lo] >> (32u - d.worldBits), codes[hi] >> (32u - d.worldBits));
        if (i == 0) d.bv_parent[0] = INVALID;
    }
}
// Node layout: internal node n keeps the boxes of its TWO CHILDREN (bv_cbox[n].l / .r) next to their links, so a
// traversal step is ONE round of independent loads per level (links + both boxes) instead of link -> box chains;
// a leaf child's box is a copy of its proxy's fat AABB (kept current by k_bvh_refit_moved).
PR_D float4 ld_cg4(const float4* p)
{
#ifndef PR_HOST_EMUL
    return __ldcg(p); // written by another thread of this launch: bypass L1
#else
    return *p;
#endif
}
// bottom-up refit: one thread per leaf carries its box upwards; at every node it stores the box into the slot of
// the child it came from, and the second thread to arrive unions both slots and continues
PR_KERNEL k_bvh_refit(DV d, unsigned nLeaves, const unsigned* sortedIdx)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nLeaves) d.bv_leafOf[sortedIdx[i]] = i;
    if (i < nLeaves && nLeaves > 1u) {
        float4 box = d.p_fat[sortedIdx[i]];
        unsigned child = i | LEAF_BIT;
        unsigned node = d.bv_lparent[i];
        while (node != INVALID) {
            const bool isLeft = d.bv_left[node] == child;
            float4* slots = reinterpret_cast<float4*>(&d.bv_cbox[node]);
            slots[isLeft ? 0 : 1] = box;
            __threadfence();
            const unsigned prev = atomicAdd(&d.bv_visit[node], 1u);
            if (prev == 0u) break;
            const float4 o = ld_cg4(&slots[isLeft ? 1 : 0]);
            box = make_float4(fminf(box.x, o.x), fminf(box.y, o.y), fmaxf(box.z, o.z), fmaxf(box.w, o.w));
            child = node;
            node = d.bv_parent[node];
        }
    }
}

// Incremental refit: when only a few proxies got a new fat box (a TOI round, a quiet regular step) the
// tree topology is kept and every ancestor box of a moved leaf is GROWN with float atomics (boxes never
// shrink between full rebuilds, so the tree stays a valid, slightly looser, bounding hierarchy).
PR_D void atomic_min_f(float* addr, float v)
{
#ifndef PR_HOST_EMUL
    v += 0.0f; // -0 -> +0
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
#else
    if (v < *addr) *addr = v;
#endif
}
PR_D void atomic_max_f(float* addr, float v)
{
#ifndef PR_HOST_EMUL
    v += 0.0f; // -0 -> +0
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
#else
    if (v > *addr) *addr = v;
#endif
}
PR_D void bvh_refit_moved_one(const DV& d, unsigned t, unsigned nLeaves)
{ // queue entry t
    const unsigned q = d.q_list[t];
    const float4 f = d.p_fat[q];
    unsigned child = d.bv_leafOf[q] | LEAF_BIT;
    unsigned node = d.bv_lparent[child & ~LEAF_BIT];
    { // the leaf's own slot: an exact copy of its fat box (only this thread writes it)
        float4* slots = reinterpret_cast<float4*>(&d.bv_cbox[node]);
        slots[(d.bv_left[node] == child) ? 0 : 1] = f;
        child = node; node = d.bv_parent[node];
    }
    while (node != INVALID) {
        float4* slot = reinterpret_cast<float4*>(&d.bv_cbox[node]) + ((d.bv_left[node] == child) ? 0 : 1);
        const float4 cur = *slot;
        if (cur.x <= f.x && cur.y <= f.y && cur.z >= f.z && cur.w >= f.w) break; // already inside: so are all ancestors
        float* b = reinterpret_cast<float*>(slot);
        atomic_min_f(b + 0, f.x); atomic_min_f(b + 1, f.y); atomic_max_f(b + 2, f.z); atomic_max_f(b + 3, f.w);
        child = node; node = d.bv_parent[node];
    }
}
PR_KERNEL k_bvh_refit_moved(DV d, unsigned nQueue, unsigned nLeaves)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    if (t < nQueue && nLeaves > 1u) bvh_refit_moved_one(d, t, nLeaves);
}

// ------------------------------------------------------------------ a4 + a5 (filter part): overlap query
PR_D bool filters_allow(unsigned fa, int ga, unsigned fb, int gb)
{ // ShouldCollide(Filter, Filter), Filter.hpp
    if (ga == gb && ga != 0) return ga > 0;
    const unsigned catA = fa & 0xFFFFu, maskA = fa >> 16, catB = fb & 0xFFFFu, maskB = fb >> 16;
    return (maskA & catB) != 0u && (catA & maskB) != 0u;
}
PR_D void query_emit(const DV& d, unsigned q, unsigned o, unsigned capPairs)
{
    if (o == q) return;
    const unsigned bq = d.p_body[q], bo = d.p_body[o];
    if (bq == bo) return;
    if (d.nWorlds && d.b_world[bq] != d.b_world[bo]) return; // batch arena: worlds share coordinates, never pairs
    const unsigned tq = d.p_tree[q], to = d.p_tree[o];
    if (d.p_moved[o] && to < tq) return; // both queued: the smaller tree id emits (unique, like sort+unique :746-749)
    const unsigned fq = d.b_flags[bq], fo = d.b_flags[bo];
    if (!(f_accelerable(fq) || f_accelerable(fo))) return; // EitherIsAccelerable :474
    const unsigned sq = d.p_shape[q], so = d.p_shape[o];
    if (d.nNcPairs) { // ShouldCollide(joints): a joint with collideConnected == false links the two bodies (:479-490)
        const unsigned long long bk = (static_cast<unsigned long long>(bq < bo ? bq : bo) << 32) | (bq < bo ? bo : bq);
        unsigned lo = 0u, hi = d.nNcPairs;
        while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (d.nc_pairs[mid] < bk) lo = mid + 1u; else hi = mid; }
        if (lo < d.nNcPairs && d.nc_pairs[lo] == bk) return;
    }
    if (!filters_allow(d.sh_filter[sq], d.sh_group[sq], d.sh_filter[so], d.sh_group[so])) return;
    const bool qMin = tq < to;
    const unsigned long long key = (static_cast<unsigned long long>(qMin ? tq : to) << 32) | (qMin ? to : tq);
    if (hash_contains(d, key)) return; // contact already exists (:2485-2489)
    const unsigned slot = active_append(&d.cnt->nPairs);
    if (slot < capPairs) { d.np_key[slot] = key; d.np_px[slot] = qMin ? make_uint2(q, o) : make_uint2(o, q); }
    else atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_PAIR_OVERFLOW));
}
PR_D bool node_has_world(const DV& d, unsigned node, unsigned world)
{ // batch arena: does internal node `node` cover leaves of world `world`? (always true for a plain world)
    if (!d.nWorlds) return true;
    const uint2 wr = d.bv_wr[node];
    return world >= wr.x && world <= wr.y;
}
// FindContacts (AabbTreeWorld.cpp:709-751), dense queue: one thread per leaf in Morton order (a leaf that is not
// queued idles), so that the lanes of a warp walk neighbouring parts of the tree. Used right after a full rebuild,
// when most proxies are queued; also the only variant of the host-emulated test build.
PR_D void bvh_query_serial(const DV& d, unsigned q, unsigned nLeaves, const unsigned* sortedIdx, unsigned capPairs)
{
    const Box qb = box_from4(d.p_fat[q]);
    const unsigned qw = d.nWorlds ? d.b_world[d.p_body[q]] : 0u;
    unsigned stack[64];
    int sp = 0;
    stack[sp++] = 0u;
    while (sp > 0) {
        const unsigned node = stack[--sp];
        const unsigned ch[2] = {d.bv_left[node], d.bv_right[node]};
        const CBox cb = d.bv_cbox[node];
        const float4 bx[2] = {cb.l, cb.r};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (!box_overlap(qb, box_from4(bx[c]))) continue;
            if (ch[c] & LEAF_BIT) query_emit(d, q, sortedIdx[ch[c] & ~LEAF_BIT], capPairs);
            else if (node_has_world(d, ch[c], qw)) {
                if (sp < 64) stack[sp++] = ch[c];
                else atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_STACK));
            }
        }
    }
}
PR_KERNEL k_bvh_query(DV d, unsigned nQueue, unsigned nLeaves, const unsigned* sortedIdx, unsigned capPairs, unsigned mortonOrder)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    unsigned q = INVALID;
    if (mortonOrder) { if (t < nLeaves) { const unsigned c = sortedIdx[t]; if (d.p_moved[c]) q = c; } }
    else if (t < nQueue) q = d.q_list[t];
    if (q != INVALID && nLeaves > 1u) bvh_query_serial(d, q, nLeaves, sortedIdx, capPairs);
}
#ifndef PR_HOST_EMUL
// Sparse queue (a settled step, a TOI round): the walk of ONE query is a chain of dependent loads ~100+ nodes long
// (measured 150 us for a single query), so G lanes share a query: they pop up to G nodes of the query's stack (in
// shared memory) at once, test both children of each, and push the survivors back with a ballot-ranked append. The
// chain shrinks to ~tree depth. The visited leaf set -- hence the pair set -- is the same as the serial walk's.
constexpr unsigned QG = 8u, QSTACK = 96u;
// one query by the QG lanes in gmask (gl = lane index inside the group); st = the group's stack in shared memory
PR_D void bvh_query_group(const DV& d, unsigned q, const unsigned* sortedIdx, unsigned capPairs, unsigned* st, unsigned lane, unsigned gl, unsigned gmask)
{
    const Box qb = box_from4(d.p_fat[q]);
    const unsigned qw = d.nWorlds ? d.b_world[d.p_body[q]] : 0u;
    if (gl == 0u) st[0] = 0u;
    unsigned count = 1u;
    __syncwarp(gmask);
    while (count) {
        const unsigned take = count < QG ? count : QG;
        const unsigned node = (gl < take) ? st[count - 1u - gl] : INVALID;
        __syncwarp(gmask);
        count -= take;
        unsigned push0 = INVALID, push1 = INVALID;
        if (node != INVALID) {
            const unsigned c0 = d.bv_left[node], c1 = d.bv_right[node];
            const CBox cb = d.bv_cbox[node];
            if (box_overlap(qb, box_from4(cb.l))) {
                if (c0 & LEAF_BIT) query_emit(d, q, sortedIdx[c0 & ~LEAF_BIT], capPairs);
                else if (node_has_world(d, c0, qw)) push0 = c0;
            }
            if (box_overlap(qb, box_from4(cb.r))) {
                if (c1 & LEAF_BIT) query_emit(d, q, sortedIdx[c1 & ~LEAF_BIT], capPairs);
                else if (node_has_world(d, c1, qw)) push1 = c1;
            }
        }
        const unsigned b0 = __ballot_sync(gmask, push0 != INVALID) & gmask, b1 = __ballot_sync(gmask, push1 != INVALID) & gmask;
        const unsigned below = (1u << lane) - 1u;
        const unsigned n0 = __popc(b0), total = n0 + __popc(b1);
        if (count + total > QSTACK) { // (uniform within the group) cannot happen for a balanced tree; reported, not silently dropped
            if (gl == 0u) atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_STACK));
            break;
        }
        if (push0 != INVALID) st[count + __popc(b0 & below)] = push0;
        if (push1 != INVALID) st[count + n0 + __popc(b1 & below)] = push1;
        count += total;
        __syncwarp(gmask);
    }
    __syncwarp(gmask);
}
PR_KERNEL k_bvh_query_coop(DV d, unsigned nQueue, unsigned nLeaves, const unsigned* sortedIdx, unsigned capPairs)
{
    PR_PDL_ENTER();
    __shared__ unsigned stk[PR_BLOCK / QG][QSTACK];
    const unsigned lane = threadIdx.x & 31u, gl = lane & (QG - 1u);
    const unsigned gmask = ((1u << QG) - 1u) << (lane & ~(QG - 1u));
    const unsigned qi = PR_TID / QG;
    if (qi >= nQueue || nLeaves <= 1u) return; // (whole groups leave together)
    bvh_query_group(d, d.q_list[qi], sortedIdx, capPairs, stk[threadIdx.x / QG], lane, gl, gmask);
}
#endif

// ------------------------------------------------------------------ spatial query: Query(tree, AABB, callback)
// (DynamicTree.cpp:673-710) every proxy whose FAT box overlaps the query box, closed intervals like
// TestOverlap -- the leaves the reference's tree query visits. One group of QG lanes per box (the broadphase's walk).
PR_D void query_box_emit(const DV& d, unsigned qi, unsigned px, unsigned cap, unsigned* out, unsigned* count)
{
    const unsigned slot = atomicAdd(count, 1u);
    if (slot < cap) { out[4u * slot] = qi; out[4u * slot + 1u] = d.p_body[px]; out[4u * slot + 2u] = d.p_shape[px]; out[4u * slot + 3u] = d.p_child[px]; }
}
PR_KERNEL k_query_boxes(DV d, unsigned nBoxes, const float4* boxes, const unsigned* boxWorld, unsigned nLeaves, const unsigned* sortedIdx,
                        unsigned cap, unsigned* out, unsigned* count)
{
    PR_PDL_ENTER();
#ifndef PR_HOST_EMUL
    __shared__ unsigned stk[PR_BLOCK / QG][QSTACK];
    const unsigned lane = threadIdx.x & 31u, gl = lane & (QG - 1u);
    const unsigned gmask = ((1u << QG) - 1u) << (lane & ~(QG - 1u));
    const unsigned qi = PR_TID / QG;
    if (qi >= nBoxes) return; // (whole groups leave together)
    const Box qb = box_from4(boxes[qi]);
    const unsigned qw = boxWorld ? boxWorld[qi] : 0u;
    if (nLeaves == 1u) { if (gl == 0u && box_overlap(qb, box_from4(d.p_fat[0])) && (!d.nWorlds || d.b_world[d.p_body[0]] == qw)) query_box_emit(d, qi, 0u, cap, out, count); return; }
    unsigned* st = stk[threadIdx.x / QG];
    if (gl == 0u) st[0] = 0u;
    unsigned n = 1u;
    __syncwarp(gmask);
    while (n) {
        const unsigned take = n < QG ? n : QG;
        const unsigned node = (gl < take) ? st[n - 1u - gl] : INVALID;
        __syncwarp(gmask);
        n -= take;
        unsigned push0 = INVALID, push1 = INVALID;
        if (node != INVALID) {
            const unsigned c0 = d.bv_left[node], c1 = d.bv_right[node];
            const CBox cb = d.bv_cbox[node];
            if (box_overlap(qb, box_from4(cb.l))) {
                if (c0 & LEAF_BIT) { const unsigned px = sortedIdx[c0 & ~LEAF_BIT]; if (!d.nWorlds || d.b_world[d.p_body[px]] == qw) query_box_emit(d, qi, px, cap, out, count); }
                else if (node_has_world(d, c0, qw)) push0 = c0;
            }
            if (box_overlap(qb, box_from4(cb.r))) {
                if (c1 & LEAF_BIT) { const unsigned px = sortedIdx[c1 & ~LEAF_BIT]; if (!d.nWorlds || d.b_world[d.p_body[px]] == qw) query_box_emit(d, qi, px, cap, out, count); }
                else if (node_has_world(d, c1, qw)) push1 = c1;
            }
        }
        const unsigned b0 = __ballot_sync(gmask, push0 != INVALID) & gmask, b1 = __ballot_sync(gmask, push1 != INVALID) & gmask;
        const unsigned below = (1u << lane) - 1u;
        const unsigned n0 = __popc(b0), total = n0 + __popc(b1);
        if (n + total > QSTACK) { if (gl == 0u) atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_STACK)); break; }
        if (push0 != INVALID) st[n + __popc(b0 & below)] = push0;
        if (push1 != INVALID) st[n + n0 + __popc(b1 & below)] = push1;
        n += total;
        __syncwarp(gmask);
    }
#else
    const unsigned qi = PR_TID;
    if (qi >= nBoxes) return;
    const Box qb = box_from4(boxes[qi]);
    const unsigned qw = boxWorld ? boxWorld[qi] : 0u;
    if (nLeaves == 1u) { if (box_overlap(qb, box_from4(d.p_fat[0])) && (!d.nWorlds || d.b_world[d.p_body[0]] == qw)) query_box_emit(d, qi, 0u, cap, out, count); return; }
    unsigned stack[64]; int sp = 0; stack[sp++] = 0u;
    while (sp > 0) {
        const unsigned node = stack[--sp];
        const unsigned ch[2] = {d.bv_left[node], d.bv_right[node]};
        const CBox cb = d.bv_cbox[node];
        const float4 bx[2] = {cb.l, cb.r};
        for (int c = 0; c < 2; ++c) {
            if (!box_overlap(qb, box_from4(bx[c]))) continue;
            if (ch[c] & LEAF_BIT) { const unsigned px = sortedIdx[ch[c] & ~LEAF_BIT]; if (!d.nWorlds || d.b_world[d.p_body[px]] == qw) query_box_emit(d, qi, px, cap, out, count); }
            else if (node_has_world(d, ch[c], qw) && sp < 64) stack[sp++] = ch[c];
        }
    }
#endif
}

// ------------------------------------------------------------------ RayCast on the LBVH
struct RayOut { bool hit; Rot normal; float fraction; };
PR_D RayOut ray_miss() { RayOut o; o.hit = false; o.normal = mkrot(0.0f, 0.0f); o.fraction = 0.0f; return o; }
// RayCast(Length radius, const Length2& location, const RayCastInput&), RayCastOutput.cpp:41-75
PR_D RayOut ray_cast_circle(float radius, float2 location, float2 p1, float2 p2, float maxFraction)
{
    const float2 s = p1 - location;
    const float b = magsq(s) - square(radius);
    const float2 raySegment = p2 - p1;
    const float c = dot(s, raySegment);
    const float rr = magsq(raySegment);
    const float sigma = square(c) - rr * b;
    if (sigma < 0.0f || almost_zero(rr)) return ray_miss();
    const float a = -(c + sqrtf(sigma));
    const float fraction = a / rr;
    if (fraction >= 0.0f && fraction <= maxFraction) {
        RayOut o; o.hit = true; o.normal = unit_vector(s + fraction * raySegment, mkrot(0.0f, 0.0f)); o.fraction = fraction; return o;
    }
    return ray_miss();
}
// RayCast(const DistanceProxy&, const RayCastInput&, const Transformation&), RayCastOutput.cpp:139-233
PR_D RayOut ray_cast_proxy(const ProxyView& proxy, float2 p1, float2 p2, float maxFraction, Xf xf)
{
    const unsigned count = proxy.count;
    const float radius = proxy.radius;
    float2 v0 = proxy.vertex(0u);
    if (count == 1u) return ray_cast_circle(radius, transform(v0, xf), p1, p2, maxFraction);
    const float2 tp1 = inv_transform(p1, xf), tp2 = inv_transform(p2, xf);
    const float2 ray0 = tp1, ray = tp2 - tp1;
    float minT = nextafterf(maxFraction, 2.0f);
    Rot normalFound = mkrot(0.0f, 0.0f);
    for (unsigned i = 0; i < count; ++i) {
        const RayOut circle = ray_cast_circle(radius, v0, tp1, tp2, maxFraction);
        if (circle.hit && minT > circle.fraction) { minT = circle.fraction; normalFound = circle.normal; }
        const float2 v1 = proxy.vertex(mod_next(i, count));
        const float2 edge = v1 - v0;
        const float rayCrossEdge = cross(ray, edge);
        if (!almost_zero(rayCrossEdge)) {
            const Rot normal = proxy.normal(i);
            const float2 offset = normal * radius;
            const float2 v0off = v0 + offset;
            const float2 qSubP = v0off - ray0;
            const float recip = 1.0f / rayCrossEdge;
            const float t = cross(qSubP, edge) * recip;
            const float u = cross(qSubP, ray) * recip;
            if (t >= 0.0f && t <= 1.0f && u >= 0.0f && u <= 1.0f) {
                if (minT > t) { minT = t; normalFound = normal; }
            }
        }
        v0 = v1;
    }
    if (minT <= maxFraction) { RayOut o; o.hit = true; o.normal = rot_rotate(normalFound, xf.q); o.fraction = minT; return o; }
    return ray_miss();
}
// the two tests of RayCast(const DynamicTree&, RayCastInput, cb) (RayCastOutput.cpp:255-271): the node's box overlaps the
// segment's box, and the segment's line does not separate from the box (|dot(v, p1 - centre)| <= dot(|v|, extents))
struct RaySeg { float2 p1; Rot v; float2 absV; Box seg; };
PR_D Box ray_seg_box(float2 p1, float2 p2, float maxFraction)
{ // GetAABB(const RayCastInput&), AABB.cpp:107-112
    const float2 fractDelta = maxFraction * (p2 - p1);
    const float2 e = p1 + fractDelta;
    Box b; b.minx = stdmin(p1.x, e.x); b.maxx = stdmax(p1.x, e.x); b.miny = stdmin(p1.y, e.y); b.maxy = stdmax(p1.y, e.y);
    return b;
}
PR_D bool ray_box_test(const RaySeg& r, const Box& box, float slack)
{
    if (!box_overlap(box, r.seg)) return false;
    const float2 center = v2((box.minx + box.maxx) / 2.0f, (box.miny + box.maxy) / 2.0f);
    const float2 extents = v2((box.maxx - box.minx) * 0.5f, (box.maxy - box.miny) * 0.5f);
    const float separation = fabsf(dot(rotv(r.v), r.p1 - center)) - dot(r.absV, extents);
    return !(separation > slack);
}
// One thread per ray. A leaf must pass the reference's two tests exactly (slack 0); an internal node -- whose box is the
// union of its leaves' boxes, so it passes whenever one of them does, up to rounding -- is given a small slack instead.
// mode PR_RAY_ALL: hits appended to out[]; PR_RAY_CLOSEST: out[ray] = least (fraction, proxy), the segment box shrinks with
// the best hit so far (what a ClipRay callback makes the reference do).
PR_D void ray_leaf(const DV& d, unsigned ri, unsigned px, RaySeg& r, float2 p1, float2 p2, float& maxFraction, unsigned qw, unsigned mode,
                   unsigned cap, pr_ray_hit* out, unsigned* count, pr_ray_hit& best, unsigned& bestPx)
{
    if (d.nWorlds && d.b_world[d.p_body[px]] != qw) return;
    if (!ray_box_test(r, box_from4(d.p_fat[px]), 0.0f)) return;
    const unsigned body = d.p_body[px];
    const RayOut o = ray_cast_proxy(child_view(d, d.p_cg[px]), p1, p2, maxFraction, xf_from4(d.b_xf[body]));
    if (!o.hit) return;
    const float2 point = p1 + (p2 - p1) * o.fraction;
    pr_ray_hit h;
    h.ray = ri; h.body = body; h.shape = d.p_shape[px]; h.child = d.p_child[px];
    h.fraction = o.fraction; h.normal_x = o.normal.c; h.normal_y = o.normal.s; h.point_x = point.x; h.point_y = point.y;
    if (mode == PR_RAY_ALL) {
        const unsigned slot = atomicAdd(count, 1u);
        if (slot < cap) out[slot] = h;
    } else if (bestPx == INVALID || o.fraction < best.fraction || (o.fraction == best.fraction && px < bestPx)) {
        best = h; bestPx = px;
        maxFraction = o.fraction; r.seg = ray_seg_box(p1, p2, maxFraction); // ClipRay: :290-294
    }
}
PR_KERNEL k_ray_cast(DV d, unsigned nRays, const float* rays5, const unsigned* rayWorld, unsigned nLeaves, const unsigned* sortedIdx,
                     unsigned mode, unsigned cap, pr_ray_hit* out, unsigned* count)
{
    PR_PDL_ENTER();
    const unsigned ri = PR_TID;
    if (ri >= nRays) return;
    const float2 p1 = v2(rays5[5u * ri], rays5[5u * ri + 1u]), p2 = v2(rays5[5u * ri + 2u], rays5[5u * ri + 3u]);
    float maxFraction = rays5[5u * ri + 4u];
    const unsigned qw = rayWorld ? rayWorld[ri] : 0u;
    RaySeg r;
    r.p1 = p1; r.v = rev_perp(unit_vector(p2 - p1, mkrot(0.0f, 0.0f))); r.absV = v2(fabsf(r.v.c), fabsf(r.v.s));
    r.seg = ray_seg_box(p1, p2, maxFraction);
    pr_ray_hit best; best.ray = ri; best.body = INVALID; best.shape = INVALID; best.child = INVALID;
    best.fraction = 0.0f; best.normal_x = best.normal_y = best.point_x = best.point_y = 0.0f;
    unsigned bestPx = INVALID;
    if (nLeaves == 1u) ray_leaf(d, ri, 0u, r, p1, p2, maxFraction, qw, mode, cap, out, count, best, bestPx);
    else {
        const float slack = 1e-3f;
        unsigned stack[64]; int sp = 0; stack[sp++] = 0u;
        while (sp > 0) {
            const unsigned node = stack[--sp];
            const unsigned ch[2] = {d.bv_left[node], d.bv_right[node]};
            const CBox cb = d.bv_cbox[node];
            const float4 bx[2] = {cb.l, cb.r};
            for (int c = 0; c < 2; ++c) {
                if (ch[c] & LEAF_BIT) ray_leaf(d, ri, sortedIdx[ch[c] & ~LEAF_BIT], r, p1, p2, maxFraction, qw, mode, cap, out, count, best, bestPx);
                else if (ray_box_test(r, box_from4(bx[c]), slack) && node_has_world(d, ch[c], qw)) {
                    if (sp < 64) stack[sp++] = ch[c];
                    else atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_STACK));
                }
            }
        }
    }
    if (mode == PR_RAY_CLOSEST) out[ri] = best;
}

// ------------------------------------------------------------------ TestPoint / FindClosestBody
// TestPoint(const DistanceProxy&, const Length2&), DistanceProxy.cpp:107-158
PR_D bool test_point_proxy(const ProxyView& proxy, float2 point)
{
    const unsigned count = proxy.count;
    const float vr = proxy.radius;
    if (count == 0u) return false;
    if (count == 1u) return magsq(point - proxy.vertex(0u)) <= square(vr);
    float maxDot = -3.402823466e+38f; // -MaxFloat
    unsigned maxIdx = 0u;
    for (unsigned i = 0; i < count; ++i) {
        const float2 delta = point - proxy.vertex(i);
        const float dt = dot(rotv(proxy.normal(i)), delta);
        if (dt > vr) return false;
        if (maxDot < dt) { maxDot = dt; maxIdx = i; }
    }
    const float2 v0 = proxy.vertex(maxIdx), v1 = proxy.vertex(mod_next(maxIdx, count));
    const float2 edge = v1 - v0;
    const float2 delta0 = v0 - point;
    if (dot(edge, delta0) >= 0.0f) return magsq(delta0) <= square(vr);
    const float2 delta1 = point - v1;
    if (dot(edge, delta1) >= 0.0f) return magsq(delta1) <= square(vr);
    return true;
}
// TestPoint(const World&, BodyID, ShapeID, const Length2&), WorldShape.cpp:178-181: the point in the body's frame against every
// child of the shape (Shape.cpp:32-41). q4[i] = (body, first child, child count, -), one thread per query.
PR_KERNEL k_test_points(DV d, unsigned n, const uint4* q4, const float2* points, unsigned* out)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const uint4 q = q4[i];
        const float2 local = inv_transform(points[i], xf_from4(d.b_xf[q.x]));
        unsigned hit = 0u;
        for (unsigned c = 0; c < q.z && !hit; ++c) hit = test_point_proxy(child_view(d, q.y + c), local) ? 1u : 0u;
        out[i] = hit;
    }
}
// FindClosestBody(const World&, const Length2&), WorldBody.cpp:576-592: least squared distance of the body's location, the first
// body in m_bodies order on a tie (strictly-less update). One BLOCK per query: a min over (distance bits, body id) keys --
// squared distances are non-negative, so their bit patterns order like the floats.
PR_KERNEL k_find_closest_bodies(DV d, unsigned nBodies, const float2* points, unsigned* out)
{
    PR_PDL_ENTER();
#ifndef PR_HOST_EMUL
    __shared__ unsigned long long best[PR_BLOCK / 32];
    const float2 p = points[blockIdx.x];
    unsigned long long key = ~0ull;
    for (unsigned b = threadIdx.x; b < nBodies; b += blockDim.x) {
        if (!(d.b_flags[b] & PR_BODY_DESTROYED)) {
            const float4 xf = d.b_xf[b];
            const float lsq = magsq(v2(xf.x, xf.y) - p);
            if (lsq == lsq) { const unsigned long long k = (static_cast<unsigned long long>(f2u(lsq + 0.0f)) << 32) | b; key = k < key ? k : key; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o); key = t < key ? t : key; }
    if ((threadIdx.x & 31u) == 0u) best[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0u) {
        for (unsigned w = 1; w < blockDim.x / 32u; ++w) key = best[w] < key ? best[w] : key;
        out[blockIdx.x] = (key == ~0ull) ? INVALID : static_cast<unsigned>(key & 0xFFFFFFFFull);
    }
#else
    const float2 p = points[PR_TID];
    unsigned long long key = ~0ull;
    for (unsigned b = 0; b < nBodies; ++b) {
        if (d.b_flags[b] & PR_BODY_DESTROYED) continue;
        const float4 xf = d.b_xf[b];
        const float lsq = magsq(v2(xf.x, xf.y) - p);
        if (lsq == lsq) { const unsigned long long k = (static_cast<unsigned long long>(f2u(lsq + 0.0f)) << 32) | b; key = k < key ? k : key; }
    }
    out[PR_TID] = (key == ~0ull) ? INVALID : static_cast<unsigned>(key & 0xFFFFFFFFull);
#endif
}

// AddContacts (AabbTreeWorld.cpp:2423-2547), creation part: sorted pair i gets its ContactID from the LIFO
// free stack else by append (ObjectPool::Allocate), is appended to m_contacts in key order.
PR_D void adj_append(const DV& d, unsigned b, unsigned id, bool eligible);
PR_D bool toi_static_eligible(unsigned flags);
PR_D void toi_mark_recompute(const DV& d, unsigned cid);
PR_D void create_contact_one(const DV& d, unsigned i, const unsigned long long* keys, const unsigned* srcIdx,
                              unsigned nOrder, unsigned nFree, unsigned poolSize, unsigned adjLive)
{
    {
        const unsigned id = (i < nFree) ? d.c_free[nFree - 1u - i] : poolSize + (i - nFree);
        const unsigned long long key = keys[i];
        const uint2 px = d.np_px[srcIdx[i]];
        const unsigned bA = d.p_body[px.x], bB = d.p_body[px.y];
        const unsigned sA = d.p_shape[px.x], sB = d.p_shape[px.y];
        const float4 matA = d.sh_mat[sA], matB = d.sh_mat[sB];
        const unsigned fA = d.b_flags[bA], fB = d.b_flags[bB];
        unsigned flags = PR_CONTACT_ENABLED | PR_CONTACT_DIRTY; // Contact ctor, Contact.hpp:343-349
        if ((fA & PR_BODY_IMPENETRABLE) || (fB & PR_BODY_IMPENETRABLE)) flags |= PR_CONTACT_IMPENETRABLE;
        const bool sensor = (matA.z != 0.0f) || (matB.z != 0.0f);
        if (sensor) flags |= PR_CONTACT_SENSOR;
        d.c_key[id] = key;
        d.c_ab[id] = make_uint4(bA, bB, d.p_cg[px.x], d.p_cg[px.y]);
        d.c_px[id] = px;
        d.c_ids[id] = make_uint4(sA, d.p_child[px.x], sB, d.p_child[px.y]);
        // MixFriction = sqrt(f1*f2), MixRestitution = max, Contact.hpp:57-69
        d.c_mat[id] = make_float4(sqrtf(matA.x * matB.x), (matA.y > matB.y) ? matA.y : matB.y, 0.0f, 1.0f);
        d.c_flags[id] = flags;
        d.c_toiIsl[id] = 0u; d.c_color[id] = COLOR_UNSET; d.c_rstamp[id] = 0u;
        manifold_clear(d, id);
        d.c_order[nOrder + i] = id;
        if (adjLive) { // the TOI phase's tables are live: m_bodyContacts push_back + TOI candidate list
            d.c_pos[id] = nOrder + i;
            const bool eligible = toi_static_eligible(flags);
            adj_append(d, bA, id, eligible); adj_append(d, bB, id, eligible);
            if (eligible) { d.toi_list[atomicAdd(&d.cnt->toiListN, 1u)] = id; toi_mark_recompute(d, id); }
        }
        hash_insert(d, key);
        if (!sensor) { // (a body woken while the TOI tables are live: see toi_wake)
            if (f_speedable(fA) && !(atomicOr(&d.b_flags[bA], static_cast<unsigned>(PR_BODY_AWAKE)) & PR_BODY_AWAKE) && adjLive) d.cnt->toiForceFull = 1u;
            if (f_speedable(fB) && !(atomicOr(&d.b_flags[bB], static_cast<unsigned>(PR_BODY_AWAKE)) & PR_BODY_AWAKE) && adjLive) d.cnt->toiForceFull = 1u;
        }
    }
}
PR_KERNEL k_create_contacts(DV d, unsigned nNew, const unsigned long long* keys, const unsigned* srcIdx,
                            unsigned nOrder, unsigned nFree, unsigned poolSize, unsigned adjLive)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nNew) create_contact_one(d, i, keys, srcIdx, nOrder, nFree, poolSize, adjLive);
}

// ------------------------------------------------------------------ a8: islands (parallel union-find)
PR_D unsigned uf_find(unsigned* parent, unsigned x)
{
    unsigned p = parent[x];
    while (p != x) {
        const unsigned gp = parent[p];
        if (gp != p) parent[x] = gp; // path halving (benign race: always an ancestor)
        x = p; p = gp;
    }
    return x;
}
PR_D void uf_union(unsigned* parent, unsigned a, unsigned b)
{
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { const unsigned t = a; a = b; b = t; } // smaller id becomes the root => root = min member
        const unsigned old = atomicCAS(&parent[b], b, a);
        if (old == b) return;
    }
}
PR_KERNEL k_isl_init(DV d, unsigned nBodies, unsigned velIters, unsigned posIters)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) {
        const unsigned f = d.b_flags[i];
        d.b_island[i] = (f_enabled(f) && f_speedable(f)) ? i : INVALID;
        d.isl_state[i] = 0u; d.isl_velIters[i] = velIters; d.isl_posIters[i] = posIters;
        d.isl_inc[i] = 0u; d.isl_maxInc[i] = 0u; d.isl_jflag[i] = 0u;
        d.isl_sep[i] = float_flip(PR_INF); d.isl_minSep[i] = float_flip(PR_INF); d.isl_minUA[i] = float_flip(PR_INF);
        d.isl_nbodies[i] = 0u; d.isl_nstat[i] = 0u;
        d.b_cmask[i] = 0ull; d.b_claim[i] = ~0ull;
        d.b_sync[i] = 0u;
    }
}
PR_D bool contact_solvable(unsigned flags)
{ // AddContactsToIsland, AabbTreeWorld.cpp:1436-1441
    return (flags & PR_CONTACT_ENABLED) && (flags & PR_CONTACT_TOUCHING) && !(flags & PR_CONTACT_SENSOR);
}
PR_KERNEL k_isl_union(DV d, unsigned nOrder)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        if (contact_solvable(d.c_flags[id])) {
            const uint4 ab = d.c_ab[id];
            if (d.b_island[ab.x] != INVALID && d.b_island[ab.y] != INVALID) uf_union(d.b_island, ab.x, ab.y);
        }
    }
}
PR_KERNEL k_isl_flatten(DV d, unsigned nBodies)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) {
        unsigned r = INVALID;
        if (d.b_island[i] != INVALID) {
            r = uf_find(d.b_island, i);
            const unsigned f = d.b_flags[i];
            // SolveReg seed test :1496 (the flag only ever gets set in this kernel: skip the atomic once it is visible)
            if (f_awake(f) && f_enabled(f) && !(*reinterpret_cast<volatile unsigned*>(&d.isl_state[r]) & ISL_ACTIVE))
                atomicOr(&d.isl_state[r], static_cast<unsigned>(ISL_ACTIVE));
        }
        d.b_root[i] = r; // separate array: parent[] is still being path-halved by other threads
    }
}
} // namespace pr
#include "pr_joint.cuh"
namespace pr {
// GetBodyConstraint's velocity (BodyConstraint.hpp:154-164): Cap(GetVelocity(body, h))
PR_D float4 body_constraint_velocity(const DV& d, unsigned i, float4 v, unsigned f, float h, const pr_step_conf& conf)
{
    if (f_accelerable(f)) { // GetVelocity(body, h), Body.cpp:290-310
        const float4 a = d.b_acc[i], m = d.b_mass[i];
        float2 l = v2(v.x, v.y) + h * v2(a.x, a.y);
        float w = v.z + h * a.z;
        l = vdiv(l, 1 + h * m.z);
        w /= (1 + h * m.w);
        v.x = l.x; v.y = l.y; v.z = w;
    }
    // Cap, Velocity.cpp:27-46
    const float2 tr = h * v2(v.x, v.y);
    const float lsq = magsq(tr);
    if (lsq > square(conf.max_translation)) { const float ratio = conf.max_translation / sqrtf(lsq); v.x *= ratio; v.y *= ratio; }
    const float absRot = fabsf(h * v.z);
    if (absRot > conf.max_rotation) { const float ratio = conf.max_rotation / absRot; v.z *= ratio; }
    return v;
}
// a1: GetBodyConstraint (BodyConstraint.hpp:154) = Cap(GetVelocity(body,h)) for bodies of active islands;
// also SolveRegIslandViaGS's pos0 <- pos1 (:1574) and AddToIsland's SetAwakeFlag (:1400).
PR_KERNEL k_isl_prepare(DV d, unsigned nBodies, pr_step_conf conf)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned isRoot = 0u, isActive = 0u;
    if (i < nBodies) {
        const unsigned r = d.b_root[i];
        const bool active = (r != INVALID) && (d.isl_state[r] & ISL_ACTIVE);
        d.b_island[i] = active ? r : INVALID; // from here on b_island is the label of this step's solve
        const float4 p1 = d.b_pos1[i];
        float4 v = d.b_vel[i];
        const unsigned f = d.b_flags[i];
        if (active) {
            isActive = 1u; isRoot = (r == i) ? 1u : 0u;
            if (!f_awake(f)) d.b_flags[i] = f | PR_BODY_AWAKE;
            const float4 p0 = d.b_pos0[i];
            d.b_pos0[i] = make_float4(p1.x, p1.y, p1.z, p0.w);
            island_add_matched(d.isl_nbodies, r, 1u);
            d.b_sync[i] = 1u;
        }
        if (active || !f_speedable(f)) v = body_constraint_velocity(d, i, v, f, conf.delta_time, conf);
        d.s_vel[i] = make_float4(v.x, v.y, v.z, 0.0f);
        d.s_pos[i] = make_float4(p1.x, p1.y, p1.z, 0.0f);
    }
    block_add(&d.cnt->islandsFound, isRoot);
    block_add(&d.cnt->nAwake, isActive);
}
// which contacts become constraints this step: solvable and in an ACTIVE island. A contact that is not a
// constraint this step loses its persistent colour (it was not in the conflict masks of this step's colouring).
PR_KERNEL k_mark_constraints(DV d, unsigned nOrder, unsigned* mark)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned touching = 0u, statTouch = 0u;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        unsigned m = 0u;
        const unsigned flags = d.c_flags[id];
        if (flags & PR_CONTACT_TOUCHING) touching = 1u;
        if (contact_solvable(flags)) {
            const uint4 ab = d.c_ab[id];
            const unsigned rA = d.b_island[ab.x], rB = d.b_island[ab.y];
            m = (rA != INVALID || rB != INVALID) ? 1u : 0u;
            // the body without a label cannot move (a speedable one would have been united with the other): the island's guest
            if (m && (rA == INVALID || rB == INVALID)) { statTouch = 1u; island_count_static(d, rA != INVALID ? rA : rB, rA != INVALID ? ab.y : ab.x); }
        }
        mark[i] = m;
        if (!m) d.c_color[id] = COLOR_UNSET;
    }
    block_add(&d.cnt->nTouching, touching);
    block_add(&d.cnt->nStaticTouch, statTouch);
}
PR_KERNEL k_gather_constraints(DV d, unsigned nOrder, const unsigned* mark, const unsigned* pos, unsigned recolorAll)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder && mark[i]) {
        const unsigned id = d.c_order[i];
        d.k_contact[pos[i]] = id;
        d.k_color[pos[i]] = recolorAll ? COLOR_UNSET : d.c_color[id];
    }
}

// ------------------------------------------------------------------ deterministic graph colouring
// Colours are PERSISTENT per contact: a constraint that was a constraint last step keeps its colour (a
// subset of a conflict-free set is conflict-free), so each step only colours the NEW constraints:
//   k_color_keep   : kept colours -> per-body 64-bit masks; uncoloured constraints -> work list
//   k_color_claim / k_color_commit : Jones-Plassmann rounds with hashed priorities over the work list;
//                    the winner takes the lowest colour free on both of its mass-carrying bodies.
// Bodies with invMass == invRotI == 0 never change in the solver and are excluded from conflicts.
// The result depends only on contact ids, list order and the world's own history (clones copy it).
PR_D unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
PR_D bool body_dyn(const DV& d, unsigned b) { const float4 m = d.b_mass[b]; return m.x != 0.0f || m.y != 0.0f; }
PR_D unsigned long long color_key(unsigned round, unsigned cid, unsigned k)
{
    return (static_cast<unsigned long long>(63u - (round & 63u)) << 58) | (static_cast<unsigned long long>(hash32(cid)) << 26) | (k & 0x3FFFFFFu);
}
PR_KERNEL k_color_keep(DV d, unsigned nBound)
{
    PR_PDL_ENTER();
    const unsigned k = PR_TID;
    bool todo = false;
    if (k < nBound && k < d.cnt->nConstraints) {
        const unsigned c = d.k_color[k];
        if (c == COLOR_UNSET) todo = true;
        else if (c < 64u) {
            const uint4 ab = d.c_ab[d.k_contact[k]];
            if (body_dyn(d, ab.x)) atomicOr(&d.b_cmask[ab.x], 1ull << c);
            if (body_dyn(d, ab.y)) atomicOr(&d.b_cmask[ab.y], 1ull << c);
        }
    }
    const unsigned slot = warp_append(&d.cnt->nColorWork, todo);
    if (todo) d.k_ulist[slot] = k;
}
PR_D void color_claim_one(const DV& d, unsigned t, unsigned round)
{
    const unsigned k = d.k_ulist[t];
    if (d.k_color[k] == COLOR_UNSET) {
        const unsigned cid = d.k_contact[k];
        const uint4 ab = d.c_ab[cid];
        const unsigned long long key = color_key(round, cid, k);
        if (body_dyn(d, ab.x)) atomicMin(&d.b_claim[ab.x], key);
        if (body_dyn(d, ab.y)) atomicMin(&d.b_claim[ab.y], key);
    }
}
// returns 1 when the constraint is still without a colour after this round
PR_D unsigned color_commit_one(const DV& d, unsigned t, unsigned round, unsigned colorCap)
{
    const unsigned k = d.k_ulist[t];
    if (d.k_color[k] != COLOR_UNSET) return 0u;
    const unsigned cid = d.k_contact[k];
    const uint4 ab = d.c_ab[cid];
    const unsigned long long key = color_key(round, cid, k);
    const bool dA = body_dyn(d, ab.x), dB = body_dyn(d, ab.y);
    const bool win = (!dA || d.b_claim[ab.x] == key) && (!dB || d.b_claim[ab.y] == key);
    if (!win) return 1u;
    const unsigned long long used = (dA ? d.b_cmask[ab.x] : 0ull) | (dB ? d.b_cmask[ab.y] : 0ull);
    const unsigned long long freeBits = ~used;
    unsigned c = COLOR_OVERFLOW;
    if (freeBits && static_cast<unsigned>(__ffsll(static_cast<long long>(freeBits))) - 1u < colorCap) { // (colorCap: see World::colorCap)
        c = static_cast<unsigned>(__ffsll(static_cast<long long>(freeBits))) - 1u;
        if (dA) d.b_cmask[ab.x] |= (1ull << c);
        if (dB) d.b_cmask[ab.y] |= (1ull << c);
    }
    d.k_color[k] = c;
    d.c_color[cid] = c;
    return 0u;
}
PR_KERNEL k_color_claim(DV d, unsigned nWork, unsigned round)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    if (t == 0u) d.cnt->nUncolored = 0u; // (k_color_commit of this round counts what is left)
    if (t < nWork) color_claim_one(d, t, round);
}
PR_KERNEL k_color_commit(DV d, unsigned nWork, unsigned round, unsigned colorCap)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    unsigned remaining = 0u;
    if (t < nWork) remaining = color_commit_one(d, t, round, colorCap);
    block_add(&d.cnt->nUncolored, remaining);
}
// A step of a settled world colours a few thousand NEW constraints in a dozen Jones-Plassmann rounds: as two launches per
// round that was ~25 launches of 7-9 us each, most of them for a few hundred leftover constraints (profiles/r02e). After the first
// COLOR_LAUNCH_ROUNDS rounds ONE block runs the remaining rounds back to back (same keys, same winners, same colours: the rounds
// are what they were, only the barriers are block barriers) over a compacted work list; it stops before the claim keys' round
// field would wrap (the host loop goes on). (All rounds in the block: measured slower -- 0.40 vs 0.33 ms for the colour stage --
// a round over 11 000 constraints is a dozen dependent gathers per thread in one block and one per thread across the chip.)
constexpr unsigned COLOR_BLOCK = 1024u, COLOR_BLOCK_MAX_WORK = 1u << 20, COLOR_BLOCK_MAX_ROUNDS = 60u, COLOR_LAUNCH_ROUNDS = 3u, COLOR_BLOCK_TAIL_MAX = 8192u;
PR_KERNEL k_color_rounds_block(DV d, unsigned colorCap, unsigned* scratch, unsigned firstRound)
{
    PR_PDL_ENTER();
#ifndef PR_HOST_EMUL
    __shared__ unsigned s_n;
    const unsigned tid = threadIdx.x, nth = blockDim.x;
#else
    unsigned s_n = 0u;
    const unsigned tid = 0u, nth = 1u;
#endif
    unsigned n = d.cnt->nColorWork;
    unsigned* cur = d.k_ulist; unsigned* nxt = scratch;
    if (firstRound) { // the first rounds ran as launches over the whole list: what they left, compacted
        if (tid == 0u) s_n = 0u;
        __syncthreads();
        for (unsigned t0 = 0; t0 < n; t0 += nth) {
            const unsigned t = t0 + tid;
            unsigned k = 0u;
            const bool left = t < n && d.k_color[k = cur[t]] == COLOR_UNSET;
            const unsigned slot = warp_append(&s_n, left);
            if (left) nxt[slot] = k;
        }
        __threadfence_block(); __syncthreads();
        n = s_n;
        unsigned* tmp = cur; cur = nxt; nxt = tmp;
        __syncthreads();
        if (n > COLOR_BLOCK_TAIL_MAX) { // too much left for one block (a world's first step): the launches go on where they stopped
            if (tid == 0u) { d.cnt->nUncolored = n; d.cnt->colorRoundsDev = firstRound; }
            return;
        }
    }
    unsigned round = firstRound;
    for (; round < COLOR_BLOCK_MAX_ROUNDS && n; ++round) {
        for (unsigned t = tid; t < n; t += nth) {
            const unsigned k = cur[t];
            const unsigned cid = d.k_contact[k];
            const uint4 ab = d.c_ab[cid];
            const unsigned long long key = color_key(round, cid, k);
            if (body_dyn(d, ab.x)) atomicMin(&d.b_claim[ab.x], key);
            if (body_dyn(d, ab.y)) atomicMin(&d.b_claim[ab.y], key);
        }
        if (tid == 0u) s_n = 0u;
        __threadfence_block(); __syncthreads();
        for (unsigned t0 = 0; t0 < n; t0 += nth) { // (every thread of the block walks every stripe: the appends below are warp-wide)
            const unsigned t = t0 + tid;
            bool left = false; unsigned k = 0u;
            if (t < n) {
                k = cur[t];
                const unsigned cid = d.k_contact[k];
                const uint4 ab = d.c_ab[cid];
                const unsigned long long key = color_key(round, cid, k);
                const bool dA = body_dyn(d, ab.x), dB = body_dyn(d, ab.y);
                const bool win = (!dA || d.b_claim[ab.x] == key) && (!dB || d.b_claim[ab.y] == key);
                if (win) {
                    const unsigned long long used = (dA ? d.b_cmask[ab.x] : 0ull) | (dB ? d.b_cmask[ab.y] : 0ull);
                    const unsigned long long freeBits = ~used;
                    unsigned c = COLOR_OVERFLOW;
                    if (freeBits && static_cast<unsigned>(__ffsll(static_cast<long long>(freeBits))) - 1u < colorCap) {
                        c = static_cast<unsigned>(__ffsll(static_cast<long long>(freeBits))) - 1u;
                        if (dA) d.b_cmask[ab.x] |= (1ull << c);
                        if (dB) d.b_cmask[ab.y] |= (1ull << c);
                    }
                    d.k_color[k] = c;
                    d.c_color[cid] = c;
                } else left = true;
            }
            const unsigned slot = warp_append(&s_n, left);
            if (left) nxt[slot] = k;
        }
        __threadfence_block(); __syncthreads();
        n = s_n;
        unsigned* tmp = cur; cur = nxt; nxt = tmp;
        __syncthreads();
    }
    if (tid == 0u) { d.cnt->nUncolored = n; d.cnt->colorRoundsDev = round; }
    // what is left (only when the round budget ran out) goes back to the head of k_ulist for the host-driven rounds
    if (n && cur != d.k_ulist) for (unsigned t = tid; t < n; t += nth) d.k_ulist[t] = cur[t];
    if (tid == 0u && n) d.cnt->nColorWork = n;
}
PR_KERNEL k_color_reset_claims(DV d, unsigned nBodies)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) d.b_claim[i] = ~0ull;
}
PR_KERNEL k_color_starts(DV d, unsigned nCons, const unsigned* sortedColor)
{
    PR_PDL_ENTER();
    const unsigned k = PR_TID;
    if (k < nCons) {
        const unsigned c = sortedColor[k];
        if (k == 0u || sortedColor[k - 1u] != c) d.cnt->colorStart[c] = k;
    }
}

// ------------------------------------------------------------------ a9: constraint setup
PR_D float2 rel_velocity(float4 vA, float2 relA, float4 vB, float2 relB)
{ // GetContactRelVelocity, d2/Math.cpp:99-102
    const float2 velBrot = rev_perp(relB) * (vB.z / 1.0f);
    const float2 velArot = rev_perp(relA) * (vA.z / 1.0f);
    return (v2(vB.x, vB.y) + velBrot) - (v2(vA.x, vA.y) + velArot);
}
// ---- constraint rows in SoA planes (pr_world.h): plane k holds bytes [16k, 16k+16) of the logical row
//   vc0 nx ny friction tangentSpeed | vc1 bodyA bodyB count|dynA<<8|dynB<<9 contact | vc2 invMA invIA invMB invIB
//   vc3 rA0 rB0 | vc4 rA1 rB1 | vc5 nm0 tm0 nm1 tm1 | vc6 bias0 bias1 k00 k11 | vc7 k01 m00 m11 m01
//   vc8 jn0 jt0 jn1 jt1 (the only plane the velocity iterations write) | vc_island
//   pc0 ln lp | pc1 p0 p1 | pc2 lcA lcB | pc3 totalRadius typeCount
// Row store: the planes of the constraint rows plus the BodyConstraint arrays they index. The regular phase keeps
// them in HBM (rs_global); a TOI event's mini island keeps them in its warp's shared-memory slice (pr_toi_phase.cuh).
struct RS {
    float4 *vc0, *vc1, *vc2, *vc3, *vc4, *vc5, *vc6, *vc7, *vc8; unsigned* vc_island;
    float4 *pc0, *pc1, *pc2, *pc3;
    float4 *s_vel, *s_pos;
    const unsigned* isl_state; // per-island early-exit flags, indexed by the rows' island field (HBM, or a block's shared memory)
};
PR_D RS rs_global(const DV& d)
{
    RS r;
    r.isl_state = d.isl_state;
    r.vc0 = d.vc0; r.vc1 = d.vc1; r.vc2 = d.vc2; r.vc3 = d.vc3; r.vc4 = d.vc4; r.vc5 = d.vc5; r.vc6 = d.vc6; r.vc7 = d.vc7; r.vc8 = d.vc8;
    r.vc_island = d.vc_island; r.pc0 = d.pc0; r.pc1 = d.pc1; r.pc2 = d.pc2; r.pc3 = d.pc3; r.s_vel = d.s_vel; r.s_pos = d.s_pos;
    return r;
}
PR_D void vc_store(const RS& r, unsigned i, const VCRow& v)
{
    r.vc0[i] = make_float4(v.nx, v.ny, v.friction, v.tangentSpeed);
    r.vc1[i] = make_float4(u2f(v.bodyA), u2f(v.bodyB), u2f(v.count | (v.dynA << 8) | (v.dynB << 9)), u2f(v.contact));
    r.vc2[i] = make_float4(v.invMA, v.invIA, v.invMB, v.invIB);
    r.vc3[i] = make_float4(v.rA0x, v.rA0y, v.rB0x, v.rB0y);
    r.vc4[i] = make_float4(v.rA1x, v.rA1y, v.rB1x, v.rB1y);
    r.vc5[i] = make_float4(v.nm0, v.tm0, v.nm1, v.tm1);
    r.vc6[i] = make_float4(v.bias0, v.bias1, v.k00, v.k11);
    r.vc7[i] = make_float4(v.k01, v.m00, v.m11, v.m01);
    r.vc8[i] = make_float4(v.jn0, v.jt0, v.jn1, v.jt1);
    r.vc_island[i] = v.island;
}
PR_D void vc_load_head(const RS& r, unsigned i, VCRow& v)
{
    const float4 a = r.vc1[i], b = r.vc2[i];
    v.bodyA = f2u(a.x); v.bodyB = f2u(a.y); v.contact = f2u(a.w);
    const unsigned cf = f2u(a.z);
    v.count = cf & 0xFFu; v.dynA = (cf >> 8) & 1u; v.dynB = (cf >> 9) & 1u; v.pad = 0u;
    v.invMA = b.x; v.invIA = b.y; v.invMB = b.z; v.invIB = b.w;
    v.island = r.vc_island[i];
}
// everything of the row but the accumulated impulses (vc8): immutable between k_build_constraints and the end of the solve
PR_D VCRow vc_load_static(const RS& r, unsigned i)
{
    VCRow v;
    vc_load_head(r, i, v);
    const float4 a = r.vc0[i], c = r.vc3[i], e = r.vc5[i], f = r.vc6[i];
    v.nx = a.x; v.ny = a.y; v.friction = a.z; v.tangentSpeed = a.w;
    v.rA0x = c.x; v.rA0y = c.y; v.rB0x = c.z; v.rB0y = c.w;
    v.nm0 = e.x; v.tm0 = e.y; v.nm1 = e.z; v.tm1 = e.w;
    v.bias0 = f.x; v.bias1 = f.y; v.k00 = f.z; v.k11 = f.w;
    v.jn0 = v.jt0 = v.jn1 = v.jt1 = 0.0f;
    if (v.count == 2u) { // second point and the block-solver matrices only exist for two-point rows
        const float4 dd = r.vc4[i], g = r.vc7[i];
        v.rA1x = dd.x; v.rA1y = dd.y; v.rB1x = dd.z; v.rB1y = dd.w;
        v.k01 = g.x; v.m00 = g.y; v.m11 = g.z; v.m01 = g.w;
    } else {
        v.rA1x = v.rA1y = v.rB1x = v.rB1y = 0.0f; v.k01 = v.m00 = v.m11 = v.m01 = 0.0f;
    }
    return v;
}
PR_D void vc_load_impulses(const RS& r, unsigned i, VCRow& v)
{
    const float4 h = r.vc8[i];
    v.jn0 = h.x; v.jt0 = h.y; v.jn1 = h.z; v.jt1 = h.w;
}
PR_D VCRow vc_load(const RS& r, unsigned i)
{
    VCRow v = vc_load_static(r, i);
    vc_load_impulses(r, i, v);
    return v;
}
PR_D void pc_store(const RS& r, unsigned i, const PCRow& p)
{
    r.pc0[i] = make_float4(p.lnx, p.lny, p.lpx, p.lpy);
    r.pc1[i] = make_float4(p.p0x, p.p0y, p.p1x, p.p1y);
    r.pc2[i] = make_float4(p.lcAx, p.lcAy, p.lcBx, p.lcBy);
    r.pc3[i] = make_float4(p.totalRadius, u2f(p.typeCount), p.rotC, p.rotS);
}
PR_D PCRow pc_load(const RS& r, unsigned i)
{
    PCRow p;
    const float4 a = r.pc0[i], b = r.pc1[i], c = r.pc2[i], e = r.pc3[i];
    p.lnx = a.x; p.lny = a.y; p.lpx = a.z; p.lpy = a.w;
    p.p0x = b.x; p.p0y = b.y; p.p1x = b.z; p.p1y = b.w;
    p.lcAx = c.x; p.lcAy = c.y; p.lcBx = c.z; p.lcBy = c.w;
    p.totalRadius = e.x; p.typeCount = f2u(e.y); p.rotC = e.z; p.rotS = e.w;
    return p;
}

// GetWorldManifold (WorldManifold.cpp:37-130) + VelocityConstraint ctor (VelocityConstraint.cpp:81-196)
// + PositionConstraint (AabbTreeWorld.cpp:289-312) for sorted constraint row i
// r: where the rows and the BodyConstraint arrays live (HBM for the regular phase, shared memory for a TOI mini island);
// sA/sB: the two bodies' indices into r.s_vel / r.s_pos (the body ids themselves in the regular phase)
// storeA/storeB: what the row keeps as its bodies' indices (the block solver, pr_block_solve.cuh, keeps slots of the block's
// shared-memory arrays; everyone else the indices the row was built from)
PR_D void build_constraint_rows(const DV& d, const RS& r, unsigned cid, unsigned i, const pr_step_conf& conf, float dtRatio, unsigned islandLabel,
                                unsigned sA, unsigned sB, unsigned storeA = INVALID, unsigned storeB = INVALID)
{
    {
        const uint4 ab = d.c_ab[cid];
        const float4 mat = d.c_mat[cid];
        const float4 q0 = d.m0[cid], q1 = d.m1[cid], q2 = d.m2[cid];
        const uint4 q3 = d.m3[cid];
        const float rA = d.ch_pack[ab.z].x, rB = d.ch_pack[ab.w].x;
        const float4 pA = r.s_pos[sA], pB = r.s_pos[sB];
        const float4 vA = r.s_vel[sA], vB = r.s_vel[sB];
        const float4 mA = d.b_mass[ab.x], mB = d.b_mass[ab.y];
        const float2 lcA = d.b_lc[ab.x], lcB = d.b_lc[ab.y];
        const Xf xfA = get_xf(pos_from4(pA), lcA), xfB = get_xf(pos_from4(pB), lcB);
        PCRow pc;
        pc.lnx = q0.x; pc.lny = q0.y; pc.lpx = q0.z; pc.lpy = q0.w;
        pc.p0x = q1.x; pc.p0y = q1.y; pc.p1x = q1.z; pc.p1y = q1.w;
        pc.lcAx = lcA.x; pc.lcAy = lcA.y; pc.lcBx = lcB.x; pc.lcBy = lcB.y;
        pc.totalRadius = rA + rB; pc.typeCount = q3.z | (q3.w << 8); pc.rotC = pc.rotS = 0.0f;
        // a body that cannot move (no velocity flag) keeps its angle through the whole solve: its rotation, which the
        // position iterations would recompute with sincosf for every row and iteration, is stored with the row
        if (!f_speedable(d.b_flags[ab.x]) && mA.x == 0.0f && mA.y == 0.0f) { pc.typeCount |= 1u << 16; pc.rotC = xfA.q.c; pc.rotS = xfA.q.s; }
        else if (!f_speedable(d.b_flags[ab.y]) && mB.x == 0.0f && mB.y == 0.0f) { pc.typeCount |= 1u << 17; pc.rotC = xfB.q.c; pc.rotS = xfB.q.s; }
        pc_store(r, i, pc);
        // world manifold
        Rot normal = mkrot(0.0f, 0.0f);
        float2 wp[2] = {v2(0, 0), v2(0, 0)};
        const float2 lp = v2(q0.z, q0.w);
        const float2 mp[2] = {v2(q1.x, q1.y), v2(q1.z, q1.w)};
        const unsigned type = q3.z, count = q3.w;
        if (type == MF_CIRCLES) {
            const float2 pointA = transform(lp, xfA), pointB = transform(mp[0], xfB);
            normal = unit_vector(pointB - pointA, mkrot(1.0f, 0.0f));
            const float2 cA = pointA + (rA * normal), cB = pointB - (rB * normal);
            wp[0] = vdiv(cA + cB, 2.0f);
        } else if (type == MF_FACE_A) {
            normal = rot_rotate(mkrot(q0.x, q0.y), xfA.q);
            const float2 planePoint = transform(lp, xfA);
            for (unsigned j = 0; j < count; ++j) {
                const float2 clip = transform(mp[j], xfB);
                const float2 cA = clip + (rA - dot(clip - planePoint, normal)) * normal;
                const float2 cB = clip - (rB * normal);
                wp[j] = vdiv(cA + cB, 2.0f);
            }
        } else if (type == MF_FACE_B) {
            const Rot nb = rot_rotate(mkrot(q0.x, q0.y), xfB.q);
            const float2 planePoint = transform(lp, xfB);
            for (unsigned j = 0; j < count; ++j) {
                const float2 clip = transform(mp[j], xfA);
                const float2 cB = clip + (rB - dot(clip - planePoint, nb)) * nb;
                const float2 cA = clip - (rA * nb);
                wp[j] = vdiv(cA + cB, 2.0f);
            }
            normal = rneg(nb);
        }
        VCRow vc;
        vc.nx = normal.c; vc.ny = normal.s; vc.friction = mat.x; vc.tangentSpeed = mat.z;
        vc.bodyA = (storeA != INVALID) ? storeA : sA; vc.bodyB = (storeB != INVALID) ? storeB : sB; vc.count = count; vc.contact = cid;
        vc.invMA = mA.x; vc.invIA = mA.y; vc.invMB = mB.x; vc.invIB = mB.y;
        vc.island = islandLabel;
        vc.dynA = (mA.x != 0.0f || mA.y != 0.0f) ? 1u : 0u; vc.dynB = (mB.x != 0.0f || mB.y != 0.0f) ? 1u : 0u; vc.pad = 0u;
        const float invMass = mA.x + mB.x;
        const Rot tangent = fwd_perp(normal);
        const float restitution = mat.y;
        float2 relA[2] = {v2(0, 0), v2(0, 0)}, relB[2] = {v2(0, 0), v2(0, 0)};
        float jn[2] = {0, 0}, jt[2] = {0, 0}, nm[2] = {0, 0}, tm[2] = {0, 0}, bias[2] = {0, 0};
        const float ojn[2] = {q2.x, q2.z}, ojt[2] = {q2.y, q2.w};
        for (unsigned j = 0; j < count; ++j) {
            relA[j] = wp[j] - v2(pA.x, pA.y);
            relB[j] = wp[j] - v2(pB.x, pB.y);
            jn[j] = ojn[j] * dtRatio; jt[j] = ojt[j] * dtRatio;
            const float2 dv = rel_velocity(vA, relA[j], vB, relB[j]);
            const float vn = dot(dv, normal);
            bias[j] = (vn < -conf.velocity_threshold) ? -restitution * vn : 0.0f;
            {
                const float a = mA.y * square(cross(relA[j], normal)) / 1.0f, b = mB.y * square(cross(relB[j], normal)) / 1.0f;
                const float value = invMass + a + b;
                nm[j] = (value != 0.0f) ? 1.0f / value : 0.0f;
            }
            {
                const float a = mA.y * square(cross(relA[j], tangent)) / 1.0f, b = mB.y * square(cross(relB[j], tangent)) / 1.0f;
                const float value = invMass + a + b;
                tm[j] = (value != 0.0f) ? 1.0f / value : 0.0f;
            }
        }
        vc.k00 = vc.k11 = vc.k01 = 0.0f; vc.m00 = vc.m11 = vc.m01 = 0.0f;
        if (conf.do_blocksolve && count == 2u) { // ComputeK + conditioning, VelocityConstraint.cpp:38-133
            const float rnA0 = cross(relA[0], normal), rnB0 = cross(relB[0], normal);
            const float rnA1 = cross(relA[1], normal), rnB1 = cross(relB[1], normal);
            const float iA0 = (mA.y * square(rnA0)) / 1.0f, iA1 = (mA.y * square(rnA1)) / 1.0f;
            const float iB0 = (mB.y * square(rnB0)) / 1.0f, iB1 = (mB.y * square(rnB1)) / 1.0f;
            const float iA = (mA.y * rnA0 * rnA1) / 1.0f, iB = (mB.y * rnB0 * rnB1) / 1.0f;
            const float k00 = invMass + iA0 + iB0, k11 = invMass + iA1 + iB1, k01 = invMass + iA + iB;
            const float kdiff = k00 * k11 - square(k01);
            if (square(k00) < 1000.0f * kdiff) {
                vc.k00 = k00; vc.k11 = k11; vc.k01 = k01;
                const float cp = k00 * k11 - k01 * k01;
                if (!almost_zero(cp)) { const float inv = 1 / cp; vc.m00 = k11 * inv; vc.m11 = k00 * inv; vc.m01 = -k01 * inv; }
            } else { // RemovePoint
                vc.count = 1u;
                relA[1] = relB[1] = v2(0, 0); jn[1] = jt[1] = nm[1] = tm[1] = bias[1] = 0.0f;
            }
        }
        vc.rA0x = relA[0].x; vc.rA0y = relA[0].y; vc.rB0x = relB[0].x; vc.rB0y = relB[0].y;
        vc.jn0 = jn[0]; vc.jt0 = jt[0]; vc.nm0 = nm[0]; vc.tm0 = tm[0];
        vc.rA1x = relA[1].x; vc.rA1y = relA[1].y; vc.rB1x = relB[1].x; vc.rB1y = relB[1].y;
        vc.jn1 = jn[1]; vc.jt1 = jt[1]; vc.nm1 = nm[1]; vc.tm1 = tm[1];
        vc.bias0 = bias[0]; vc.bias1 = bias[1];
        vc_store(r, i, vc);
    }
}
// blockMode: the rows are for the block solver (pr_block_solve.cuh) -- they keep their bodies' slots and the slot of the
// island's root instead of body ids and the island's label
PR_KERNEL k_build_constraints(DV d, unsigned nCons, pr_step_conf conf, float dtRatio, unsigned blockMode)
{
    // (no early trigger: the solver kernels that follow fetch the rows written here BEFORE their own dependency wait)
    pr_pdl_wait();
    const unsigned i = PR_TID;
    if (i < nCons) {
        const unsigned cid = d.k_contact[d.k_sorted[i]];
        const uint4 ab = d.c_ab[cid];
        const unsigned rAi = d.b_island[ab.x], rBi = d.b_island[ab.y];
        const unsigned label = (rAi != INVALID) ? rAi : rBi;
        if (blockMode) build_constraint_rows(d, rs_global(d), cid, i, conf, dtRatio, d.b_slot[label], ab.x, ab.y, d.b_slot[ab.x], d.b_slot[ab.y]);
        else build_constraint_rows(d, rs_global(d), cid, i, conf, dtRatio, label, ab.x, ab.y);
    }
}

// ------------------------------------------------------------------ a9/a10: warm start + velocity solve
PR_D void apply_pair(float4& vA, float4& vB, float2 P, float2 relA, float2 relB, float imA, float iiA, float imB, float iiB)
{
    const float LA = cross(relA, P) / 1.0f, LB = cross(relB, P) / 1.0f;
    const float2 a = v2(vA.x, vA.y) - imA * P; vA.x = a.x; vA.y = a.y; vA.z -= iiA * LA;
    const float2 b = v2(vB.x, vB.y) + imB * P; vB.x = b.x; vB.y = b.y; vB.z += iiB * LB;
}
// WarmStartVelocities, AabbTreeWorld.cpp:222-273
PR_D void warm_start_loaded(const RS& r, const VCRow& vc);
PR_D void warm_start_row(const RS& r, unsigned i) { warm_start_loaded(r, vc_load(r, i)); }
PR_D void warm_start_core(const VCRow& vc, float4& vA, float4& vB);
PR_D void warm_start_loaded(const RS& r, const VCRow& vc)
{
    float4 vA = r.s_vel[vc.bodyA], vB = r.s_vel[vc.bodyB];
    warm_start_core(vc, vA, vB);
    if (vc.dynA) r.s_vel[vc.bodyA] = vA;
    if (vc.dynB) r.s_vel[vc.bodyB] = vB;
}
PR_D void warm_start_core(const VCRow& vc, float4& vA, float4& vB)
{
    const Rot normal = mkrot(vc.nx, vc.ny), tangent = fwd_perp(normal);
    float2 dAl = v2(0, 0), dBl = v2(0, 0); float dAa = 0.0f, dBa = 0.0f;
    const float2 relA[2] = {v2(vc.rA0x, vc.rA0y), v2(vc.rA1x, vc.rA1y)}, relB[2] = {v2(vc.rB0x, vc.rB0y), v2(vc.rB1x, vc.rB1y)};
    const float jn[2] = {vc.jn0, vc.jn1}, jt[2] = {vc.jt0, vc.jt1};
    for (unsigned j = 0; j < vc.count; ++j) {
        const float2 P = jn[j] * normal + jt[j] * tangent;
        const float LA = cross(relA[j], P) / 1.0f, LB = cross(relB[j], P) / 1.0f;
        dAl = dAl - vc.invMA * P; dAa -= vc.invIA * LA;
        dBl = dBl + vc.invMB * P; dBa += vc.invIB * LB;
    }
    vA = make_float4(vA.x + dAl.x, vA.y + dAl.y, vA.z + dAa, vA.w); // (.w: the dataflow solver's row count, else 0)
    vB = make_float4(vB.x + dBl.x, vB.y + dBl.y, vB.z + dBa, vB.w);
}
PR_KERNEL k_warm_start(DV d, unsigned start, unsigned count)
{ // (launched with PR_LAUNCH_PDL: the row is fetched while the previous colour drains; the bodies after it has)
    const unsigned t = PR_TID;
    VCRow vc;
    if (t < count) vc = vc_load(rs_global(d), start + t); // (the stored impulses are k_build_constraints': no colour writes them)
    pr_pdl_launch_dependents();
    pr_pdl_wait();
    if (t < count) warm_start_loaded(rs_global(d), vc);
}
// ---- colour-overflow tail (constraints that found no free colour < 64, e.g. the 65th+ contact of one body):
// the tail is sorted by island label and solved by ONE THREAD PER ISLAND SEGMENT in list order -- a valid
// Gauss-Seidel order inside each island, islands in parallel.
PR_KERNEL k_tail_keys(DV d, unsigned tailStart, unsigned tailCount, unsigned* keys, unsigned* vals)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    if (t < tailCount) {
        const unsigned k = d.k_sorted[tailStart + t];
        const uint4 ab = d.c_ab[d.k_contact[k]];
        const unsigned ia = d.b_island[ab.x];
        keys[t] = (ia != INVALID) ? ia : d.b_island[ab.y];
        vals[t] = k;
    }
}
PR_KERNEL k_tail_segments(DV d, unsigned tailStart, unsigned tailCount, const unsigned* keys, const unsigned* vals)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    bool head = false;
    if (t < tailCount) {
        d.k_sorted[tailStart + t] = vals[t];
        head = (t == 0u) || (keys[t] != keys[t - 1u]);
    }
    const unsigned slot = warp_append(&d.cnt->nTailSegs, head);
    if (head) d.k_ulist[slot] = t;
}
template <class F> PR_D void tail_for_each_row(const DV& d, unsigned tailStart, unsigned tailCount, unsigned nSegs, F&& fn)
{
    const unsigned s = PR_TID;
    if (s < nSegs) {
        unsigned t = d.k_ulist[s];
        const unsigned label = d.vc_island[tailStart + t];
        for (; t < tailCount && d.vc_island[tailStart + t] == label; ++t) fn(tailStart + t);
    }
}
PR_KERNEL k_warm_start_tail(DV d, unsigned tailStart, unsigned tailCount, unsigned nSegs)
{
    PR_PDL_ENTER();
    tail_for_each_row(d, tailStart, tailCount, nSegs, [&](unsigned row) { warm_start_row(rs_global(d), row); });
}

// GaussSeidel::SolveVelocityConstraint (ContactSolver.cpp:481): tangent (:395) then normal
// (block :245 with cases 1-4 :113-243, else sequential :337); points in order 1 then 0.
PR_D float solve_velocity_loaded(const DV& d, const RS& r, unsigned i, VCRow vc, bool& did, bool checkIsland);
PR_D float solve_velocity_row(const DV& d, const RS& r, unsigned i, bool& did, bool checkIsland = true)
{
    return solve_velocity_loaded(d, r, i, vc_load(r, i), did, checkIsland);
}
// vc: the whole row, impulses included
PR_D float solve_velocity_core(const VCRow& vc, float4& vA, float4& vB, float4& impulses);
PR_D float solve_velocity_loaded(const DV& d, const RS& r, unsigned i, VCRow vc, bool& did, bool checkIsland)
{
    did = false;
    // (the two body loads are issued before the island test, not behind it: one round trip less in the dependent chain)
    float4 vA = r.s_vel[vc.bodyA], vB = r.s_vel[vc.bodyB];
    if (checkIsland && (r.isl_state[vc.island] & ISL_VEL_DONE)) return 0.0f;
    did = true;
    float4 impulses;
    const float maxInc = solve_velocity_core(vc, vA, vB, impulses);
    if (vc.dynA) r.s_vel[vc.bodyA] = vA;
    if (vc.dynB) r.s_vel[vc.bodyB] = vB;
    r.vc8[i] = impulses;
    return maxInc;
}
// the row's arithmetic on registers: the two bodies' velocities in and out, the new accumulated impulses out
PR_D float solve_velocity_core(const VCRow& vc, float4& vA, float4& vB, float4& impulses)
{
    const Rot normal = mkrot(vc.nx, vc.ny), tangent = fwd_perp(normal);
    const float2 relA[2] = {v2(vc.rA0x, vc.rA0y), v2(vc.rA1x, vc.rA1y)}, relB[2] = {v2(vc.rB0x, vc.rB0y), v2(vc.rB1x, vc.rB1y)};
    float jn[2] = {vc.jn0, vc.jn1}, jt[2] = {vc.jt0, vc.jt1};
    const float nm[2] = {vc.nm0, vc.nm1}, tm[2] = {vc.tm0, vc.tm1}, bias[2] = {vc.bias0, vc.bias1};
    float maxInc = 0.0f;
    { // SolveTangentConstraint
        float m = 0.0f;
        for (int j = (vc.count == 2u) ? 1 : 0; j >= 0; --j) {
            const float2 closing = rel_velocity(vA, relA[j], vB, relB[j]);
            const float directionalVel = vc.tangentSpeed - dot(closing, tangent);
            const float lambda = tm[j] * directionalVel;
            const float maxImpulse = vc.friction * jn[j];
            const float oldImpulse = jt[j];
            const float newImpulse = stdclamp(oldImpulse + lambda, -maxImpulse, maxImpulse);
            const float inc = newImpulse - oldImpulse;
            apply_pair(vA, vB, inc * tangent, relA[j], relB[j], vc.invMA, vc.invIA, vc.invMB, vc.invIB);
            m = stdmax(m, fabsf(inc));
            jt[j] = oldImpulse + inc;
        }
        maxInc = stdmax(maxInc, m);
    }
    const bool kZero = (vc.k00 == 0.0f && vc.k11 == 0.0f && vc.k01 == 0.0f);
    if (vc.count == 1u || kZero) { // SeqSolveNormalConstraint
        float m = 0.0f;
        for (int j = (vc.count == 2u) ? 1 : 0; j >= 0; --j) {
            const float2 closing = rel_velocity(vA, relA[j], vB, relB[j]);
            const float directionalVel = dot(closing, normal);
            const float lambda = nm[j] * (bias[j] - directionalVel);
            const float oldImpulse = jn[j];
            const float newImpulse = stdmax(oldImpulse + lambda, 0.0f);
            const float inc = newImpulse - oldImpulse;
            apply_pair(vA, vB, inc * normal, relA[j], relB[j], vc.invMA, vc.invIA, vc.invMB, vc.invIB);
            m = stdmax(m, fabsf(inc));
            jn[j] = oldImpulse + inc;
        }
        maxInc = stdmax(maxInc, m);
    } else { // BlockSolveNormalConstraint
        const float2 dv0 = rel_velocity(vA, relA[0], vB, relB[0]), dv1 = rel_velocity(vA, relA[1], vB, relB[1]);
        const float vn0 = dot(dv0, normal), vn1 = dot(dv1, normal);
        const float b0 = vn0 - bias[0], b1 = vn1 - bias[1];
        float t0 = 0.0f; t0 += vc.k00 * jn[0]; t0 += vc.k01 * jn[1];
        float t1 = 0.0f; t1 += vc.k01 * jn[0]; t1 += vc.k11 * jn[1];
        const float bp0 = b0 - t0, bp1 = b1 - t1;
        float n0 = 0.0f, n1 = 0.0f; bool solved = false;
        {
            float x0 = 0.0f; x0 += vc.m00 * bp0; x0 += vc.m01 * bp1;
            float x1 = 0.0f; x1 += vc.m01 * bp0; x1 += vc.m11 * bp1;
            x0 = -x0; x1 = -x1;
            if (x0 >= 0.0f && x1 >= 0.0f) { n0 = x0; n1 = x1; solved = true; }
        }
        if (!solved) {
            const float x0 = -nm[0] * bp0;
            const float vn2 = vc.k01 * x0 + bp1;
            if (x0 >= 0.0f && vn2 >= 0.0f) { n0 = x0; n1 = 0.0f; solved = true; }
        }
        if (!solved) {
            const float x1 = -nm[1] * bp1;
            const float vn1b = vc.k01 * x1 + bp0;
            if (x1 >= 0.0f && vn1b >= 0.0f) { n0 = 0.0f; n1 = x1; solved = true; }
        }
        if (!solved && bp0 >= 0.0f && bp1 >= 0.0f) { n0 = 0.0f; n1 = 0.0f; solved = true; }
        if (solved) { // BlockSolveUpdate + GetVelocityDelta, ContactSolver.cpp:69-111
            const float i0 = n0 - jn[0], i1 = n1 - jn[1];
            const float2 P0 = i0 * normal, P1 = i1 * normal;
            const float2 P = P0 + P1;
            const float LA = (cross(relA[0], P0) + cross(relA[1], P1)) / 1.0f;
            const float LB = (cross(relB[0], P0) + cross(relB[1], P1)) / 1.0f;
            const float2 dAl = -(vc.invMA * P); const float dAa = -(vc.invIA * LA);
            const float2 dBl = vc.invMB * P; const float dBa = vc.invIB * LB;
            vA.x += dAl.x; vA.y += dAl.y; vA.z += dAa;
            vB.x += dBl.x; vB.y += dBl.y; vB.z += dBa;
            jn[0] = n0; jn[1] = n1;
            maxInc = stdmax(maxInc, stdmax(fabsf(n0), fabsf(n1)));
        }
    }
    impulses = make_float4(jn[0], jt[0], jn[1], jt[1]);
    return maxInc;
}
#if !defined(PR_HOST_EMUL)
#ifndef PR_VEL_MINBLOCKS
#define PR_VEL_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(PR_BLOCK, PR_VEL_MINBLOCKS)
#else
PR_KERNEL
#endif
k_solve_velocity(DV d, unsigned start, unsigned count)
{
    // (launched with PR_LAUNCH_PDL) prologue: the immutable planes of the row, while the previous colour drains
    const unsigned t = PR_TID;
    const RS r = rs_global(d);
    bool did = false; float inc = 0.0f; unsigned label = 0u;
    VCRow vc;
    if (t < count) vc = vc_load_static(r, start + t);
    pr_pdl_launch_dependents();
    pr_pdl_wait(); // from here on: what earlier kernels wrote (impulses, body velocities, island flags)
    if (t < count) { vc_load_impulses(r, start + t, vc); inc = solve_velocity_loaded(d, r, start + t, vc, did, true); label = vc.island; }
    island_max(d.isl_inc, label, f2u(inc), did);
}
PR_KERNEL k_solve_velocity_tail(DV d, unsigned tailStart, unsigned tailCount, unsigned nSegs)
{
    PR_PDL_ENTER();
    tail_for_each_row(d, tailStart, tailCount, nSegs, [&](unsigned row) {
        bool did = false;
        const float inc = solve_velocity_row(d, rs_global(d), row, did);
        if (did) atomicMax(&d.isl_inc[d.vc_island[row]], f2u(inc));
    });
}
// per-island early exit of velocity iterations (AabbTreeWorld.cpp:1603-1622)
PR_D unsigned isl_vel_check_root(const DV& d, unsigned i, unsigned iter, float minMomentum)
{ // i is an island root; returns 1 if the island must keep iterating
    const unsigned st = d.isl_state[i];
    if (!(st & ISL_ACTIVE) || (st & ISL_VEL_DONE)) return 0u;
    const unsigned incBits = d.isl_inc[i];
    if (incBits > d.isl_maxInc[i]) d.isl_maxInc[i] = incBits; // non-negative floats order like uints
    const bool jointsOkay = d.isl_jflag[i] == 0u; // every joint's SolveVelocity left the velocities unchanged
    unsigned still = 0u;
    if (jointsOkay && u2f(incBits) <= minMomentum) { d.isl_state[i] = st | ISL_VEL_DONE; d.isl_velIters[i] = iter + 1u; }
    else still = 1u;
    d.isl_inc[i] = 0u; d.isl_jflag[i] = 0u;
    return still;
}
// iteration 0 scans all bodies for island roots and LISTS the islands still iterating (k_ulist2): a pile is a handful of
// islands among hundreds of thousands of single falling bodies, which are done after the first iteration
PR_KERNEL k_isl_vel_check(DV d, unsigned nBodies, unsigned iter, float minMomentum)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned stillActive = 0u;
    if (i < nBodies && d.b_island[i] == i) stillActive = isl_vel_check_root(d, i, iter, minMomentum);
    const unsigned slot = warp_append(&d.cnt->nActiveVel, stillActive != 0u);
    if (stillActive) d.isl_active[slot] = i;
}
// later iterations: only the listed islands (grid-stride; the list never grows)
PR_KERNEL k_isl_vel_check_list(DV d, unsigned iter, float minMomentum)
{
    PR_PDL_ENTER();
    const unsigned n = d.cnt->nActiveVel;
    for (unsigned k = PR_TID; k < n; k += PR_NTHREADS) isl_vel_check_root(d, d.isl_active[k], iter, minMomentum);
}

// ------------------------------------------------------------------ a11: integrate + NGS position solve
PR_KERNEL k_integrate_positions(DV d, unsigned nBodies, float h)
{
    PR_PDL_ENTER(); // IntegratePositions, AabbTreeWorld.cpp:176-189
    const unsigned i = PR_TID;
    if (i < nBodies && d.b_island[i] != INVALID) {
        const float4 v = d.s_vel[i];
        float4 p = d.s_pos[i];
        const float2 tr = h * v2(v.x, v.y);
        const float rot = h * v.z;
        p.x += tr.x; p.y += tr.y; p.z += rot;
        d.s_pos[i] = p;
    }
}
struct PSM { Rot normal; float2 point; float separation; };
PR_D PSM get_psm(const PCRow& pc, unsigned index, Xf xfA, Xf xfB)
{ // GetPSM, PositionSolverManifold.cpp:35-120
    const float2 lp = v2(pc.lpx, pc.lpy);
    const float2 plp = index == 0u ? v2(pc.p0x, pc.p0y) : v2(pc.p1x, pc.p1y);
    const Rot ln = mkrot(pc.lnx, pc.lny);
    const unsigned type = pc.typeCount & 0xFFu;
    PSM r;
    if (type == MF_CIRCLES) {
        const float2 pointA = transform(lp, xfA), pointB = transform(plp, xfB);
        const float2 delta = pointB - pointA;
        r.normal = unit_vector(delta, mkrot(0.0f, 0.0f));
        r.point = vdiv(pointA + pointB, 2.0f);
        r.separation = dot(delta, r.normal);
    } else if (type == MF_FACE_A) {
        const float2 planePoint = transform(lp, xfA);
        r.normal = rot_rotate(ln, xfA.q);
        r.point = transform(plp, xfB);
        r.separation = dot(r.point - planePoint, r.normal);
    } else {
        const float2 planePoint = transform(lp, xfB);
        const Rot n = rot_rotate(ln, xfB.q);
        r.point = transform(plp, xfA);
        r.separation = dot(r.point - planePoint, n);
        r.normal = rneg(n);
    }
    return r;
}
struct PDelta { float2 al; float aa; float2 bl; float ba; float sep; };
PR_D PDelta psm_solve(const PSM& psm, float2 pA, float2 pB, float imA, float iiA, float imB, float iiB, float totalRadius,
                      float rate, float slop, float maxCorr)
{ // solver_fn in SolvePositionConstraint, ContactSolver.cpp:521-545
    const float separation = psm.separation - totalRadius;
    const float2 rA = psm.point - pA, rB = psm.point - pB;
    const float K = (imA + imB) + iiA * square(cross(rA, psm.normal) / 1.0f) + iiB * square(cross(rB, psm.normal) / 1.0f);
    const float C = -stdclamp(rate * (separation + slop), -maxCorr, 0.0f);
    const float2 P = vdiv(psm.normal * C, K);
    const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
    PDelta r;
    r.al = -(imA * P); r.aa = -(iiA * LA); r.bl = imB * P; r.ba = iiB * LB; r.sep = separation;
    return r;
}
// GaussSeidel::SolvePositionConstraint, ContactSolver.cpp:494-643 (moveA = moveB = true)
PR_D float solve_position_loaded(const DV& d, const RS& r, const VCRow& vch, const PCRow& pc, float rate, float slop, float maxCorr, bool& did, bool checkIsland);
PR_D float solve_position_core(const VCRow& vch, const PCRow& pc, float4& a4, float4& b4, float rate, float slop, float maxCorr);
PR_D float solve_position_row(const DV& d, const RS& r, unsigned i, float rate, float slop, float maxCorr, bool& did, bool checkIsland = true)
{
    VCRow vch; vc_load_head(r, i, vch);
    return solve_position_loaded(d, r, vch, pc_load(r, i), rate, slop, maxCorr, did, checkIsland);
}
// vch: the row's head (bodies, inverse masses, island), pc: its position-constraint planes -- both immutable during the solve
PR_D float solve_position_loaded(const DV& d, const RS& r, const VCRow& vch, const PCRow& pc, float rate, float slop, float maxCorr, bool& did, bool checkIsland)
{
    did = false;
    float4 a4 = r.s_pos[vch.bodyA], b4 = r.s_pos[vch.bodyB]; // (issued before the island test: see solve_velocity_row)
    if (checkIsland && (r.isl_state[vch.island] & ISL_POS_SOLVED)) return PR_INF;
    did = true;
    const float minSep = solve_position_core(vch, pc, a4, b4, rate, slop, maxCorr);
    if (vch.dynA) r.s_pos[vch.bodyA] = a4;
    if (vch.dynB) r.s_pos[vch.bodyB] = b4;
    return minSep;
}
// the row's arithmetic on registers: the two bodies' positions (x, y, angle, 0) in and out
PR_D float solve_position_core(const VCRow& vch, const PCRow& pc, float4& a4, float4& b4, float rate, float slop, float maxCorr)
{
    const float imA = vch.invMA, iiA = vch.invIA, imB = vch.invMB, iiB = vch.invIB;
    Pos posA = pos_from4(a4), posB = pos_from4(b4);
    const float2 lcA = v2(pc.lcAx, pc.lcAy), lcB = v2(pc.lcBx, pc.lcBy);
    const unsigned count = (pc.typeCount >> 8) & 0xFFu;
    const bool fixedA = (pc.typeCount & (1u << 16)) != 0u, fixedB = (pc.typeCount & (1u << 17)) != 0u;
    const Rot fixedRot = mkrot(pc.rotC, pc.rotS);
    // GetTransformation(pos, lc) with the static body's UnitVec::Get(angle) taken from the row (same bits: same angle)
    auto xf_a = [&](const Pos& p) { return fixedA ? get_xf(p.l, fixedRot, lcA) : get_xf(p, lcA); };
    auto xf_b = [&](const Pos& p) { return fixedB ? get_xf(p.l, fixedRot, lcB) : get_xf(p, lcB); };
    float minSep = PR_INF;
    if (count == 1u) {
        const PSM psm0 = get_psm(pc, 0u, xf_a(posA), xf_b(posB));
        const PDelta s = psm_solve(psm0, posA.l, posB.l, imA, iiA, imB, iiB, pc.totalRadius, rate, slop, maxCorr);
        posA.l = posA.l + s.al; posA.a += s.aa; posB.l = posB.l + s.bl; posB.a += s.ba;
        minSep = 0.0f + s.sep;
    } else if (count == 2u) {
        const Xf xfA = xf_a(posA), xfB = xf_b(posB);
        const PSM psm0 = get_psm(pc, 0u, xfA, xfB), psm1 = get_psm(pc, 1u, xfA, xfB);
        if (almost_equal(psm0.separation, psm1.separation)) {
            const PDelta s0 = psm_solve(psm0, posA.l, posB.l, imA, iiA, imB, iiB, pc.totalRadius, rate, slop, maxCorr);
            const PDelta s1 = psm_solve(psm1, posA.l, posB.l, imA, iiA, imB, iiB, pc.totalRadius, rate, slop, maxCorr);
            posA.l = (posA.l + s0.al) + s1.al; posA.a = (posA.a + s0.aa) + s1.aa;
            posB.l = (posB.l + s0.bl) + s1.bl; posB.a = (posB.a + s0.ba) + s1.ba;
            minSep = s0.sep;
        } else if (psm0.separation < psm1.separation) {
            const PDelta s0 = psm_solve(psm0, posA.l, posB.l, imA, iiA, imB, iiB, pc.totalRadius, rate, slop, maxCorr);
            posA.l = posA.l + s0.al; posA.a += s0.aa; posB.l = posB.l + s0.bl; posB.a += s0.ba;
            const PSM p1 = get_psm(pc, 1u, xf_a(posA), xf_b(posB));
            const PDelta s1 = psm_solve(p1, posA.l, posB.l, imA, iiA, imB, iiB, pc.totalRadius, rate, slop, maxCorr);
            posA.l = posA.l + s1.al; posA.a += s1.aa; posB.l = posB.l + s1.bl; posB.a += s1.ba;
            minSep = s0.sep;
        } else if (psm1.separation < psm0.separation) {
            const PDelta s1 = psm_solve(psm1, posA.l, posB.l, imA, iiA, imB, iiB, pc.totalRadius, rate, slop, maxCorr);
            posA.l = posA.l + s1.al; posA.a += s1.aa; posB.l = posB.l + s1.bl; posB.a += s1.ba;
            const PSM p0 = get_psm(pc, 0u, xf_a(posA), xf_b(posB));
            const PDelta s0 = psm_solve(p0, posA.l, posB.l, imA, iiA, imB, iiB, pc.totalRadius, rate, slop, maxCorr);
            posA.l = posA.l + s0.al; posA.a += s0.aa; posB.l = posB.l + s0.bl; posB.a += s0.ba;
            minSep = s1.sep;
        }
    }
    a4 = make_float4(posA.l.x, posA.l.y, posA.a, a4.w); // (.w: the dataflow solver's row count, else 0)
    b4 = make_float4(posB.l.x, posB.l.y, posB.a, b4.w);
    return minSep;
}
PR_KERNEL k_solve_position(DV d, unsigned start, unsigned count, float rate, float slop, float maxCorr)
{
    // (launched with PR_LAUNCH_PDL: see k_solve_velocity)
    const unsigned t = PR_TID;
    const RS r = rs_global(d);
    bool did = false; float sep = PR_INF; unsigned label = 0u;
    VCRow vch; PCRow pc;
    if (t < count) { vc_load_head(r, start + t, vch); pc = pc_load(r, start + t); }
    pr_pdl_launch_dependents();
    pr_pdl_wait();
    if (t < count) { sep = solve_position_loaded(d, r, vch, pc, rate, slop, maxCorr, did, true); label = vch.island; }
    island_min(d.isl_sep, label, float_flip(sep), did);
}
// ------------------------------------------------------------------ a10/a11 as ONE launch per pass: the dataflow solver
// The coloured solve is a launch per colour and pass: 15 colours x (1 + 8 + 3) passes = 180 dependent launches of a few
// microseconds each, although a row only ever depends on the rows of LOWER colour that share one of its two bodies. Here a
// pass (warm start, one velocity iteration, one position iteration) is one launch over all coloured rows in solve order:
//   * every BodyConstraint row carries, in its unused fourth word, the number of rows applied to it so far (s_vel.w during
//     the warm start and the velocity iterations, s_pos.w during the position iterations). A row knows from the colour masks
//     how many rows of lower colour each of its bodies has (vc_flow: rank and row count per body), hence the count each body
//     must show before the row may run -- pass * rows(body) + rank(row, body) -- and it stores count + 1 WITH the new
//     velocity / position in one 16-byte store. Per body the rows therefore run in colour order: exactly the order of the
//     per-colour launches, bit for bit, with the waiting confined to true dependencies -- and without a fence: the 16-byte
//     row is the message and the flag at once (aligned 128-bit accesses to L2 are single transactions).
//   * blocks take their 256 rows by a ticket (atomic counter per pass), so rows START in solve order: a waiting row only
//     waits for rows that are already running or done, whatever order the hardware dispatches blocks in -- no deadlock.
// Bodies that carry no mass are never written by the solver and are not waited for (the colouring ignores them as well).
constexpr unsigned FLOW_SPIN_LIMIT = 1u << 22;
#ifndef PR_HOST_EMUL
PR_D float4 flow_ld(const float4* p)
{ // one 128-bit load from L2 (never L1: other SMs write these rows during the kernel)
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
PR_D void flow_st(float4* p, float4 v) { asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
#else
PR_D float4 flow_ld(const float4* p) { return *p; }
PR_D void flow_st(float4* p, float4 v) { *p = v; }
#endif
// rank | rows << 8 per body (a body's rows all have different colours: the colour mask counts them), body B's << 16
PR_KERNEL k_flow_words(DV d, unsigned nRows, const unsigned* sortedColor)
{
    pr_pdl_wait(); // (no early trigger: the first pass reads vc_flow BEFORE its own dependency wait)
    const unsigned i = PR_TID;
    if (i < nRows) {
        const float4 head = d.vc1[i];
        const unsigned cf = f2u(head.z), color = sortedColor[i];
        const unsigned long long below = (1ull << color) - 1ull;
        const unsigned long long mA = ((cf >> 8) & 1u) ? d.b_cmask[f2u(head.x)] : 0ull, mB = ((cf >> 9) & 1u) ? d.b_cmask[f2u(head.y)] : 0ull;
        d.vc_flow[i] = static_cast<unsigned>(__popcll(mA & below)) | (static_cast<unsigned>(__popcll(mA)) << 8) |
                       (static_cast<unsigned>(__popcll(mB & below)) << 16) | (static_cast<unsigned>(__popcll(mB)) << 24);
    }
}
// mode 0: WarmStartVelocities, 1: one velocity iteration, 2: one position iteration; nRows: the coloured rows (the overflow
// tail, if any, follows in its own launch); pass: how many passes ran on the same array (s_vel / s_pos) before this one
#ifndef PR_FLOW_MINBLOCKS
#define PR_FLOW_MINBLOCKS 2 // (measured on B200, 1M mixed: 2 -> velocity 0.60 + position 0.33 ms, 3 -> 0.62 + 0.36 (80 registers, spills), 4 -> 0.66 + 0.37)
#endif
template <unsigned mode> PR_KERNEL_LB(PR_BLOCK, PR_FLOW_MINBLOCKS) k_flow_pass(DV d, unsigned nRows, unsigned ticket, unsigned pass, float rate, float slop, float maxCorr)
{
    pr_pdl_launch_dependents();
#ifndef PR_HOST_EMUL
    __shared__ unsigned s_chunk;
    if (threadIdx.x == 0u) s_chunk = atomicAdd(&d.cnt->flowTicket[ticket & 31u], 1u);
    __syncthreads();
    const unsigned i = s_chunk * PR_BLOCK + threadIdx.x;
#else
    const unsigned i = PR_TID;
#endif
    const RS r = rs_global(d);
    const bool valid = i < nRows;
    VCRow vc; PCRow pc;
    unsigned tA = 0u, tB = 0u;
    if (valid) { // the immutable part of the row, while the previous pass drains
        if (mode == 2u) { vc_load_head(r, i, vc); pc = pc_load(r, i); }
        else vc = vc_load_static(r, i);
        const unsigned fw = d.vc_flow[i];
        tA = pass * ((fw >> 8) & 0xFFu) + (fw & 0xFFu);
        tB = pass * (fw >> 24) + ((fw >> 16) & 0xFFu);
    }
    pr_pdl_wait();
    if (valid && mode != 2u) vc_load_impulses(r, i, vc);
    float4* const rows = (mode == 2u) ? d.s_pos : d.s_vel;
    bool done = !valid, did = false;
    float result = (mode == 2u) ? PR_INF : 0.0f;
    bool skip = false;
    if (valid) skip = (mode == 1u) ? (d.isl_state[vc.island] & ISL_VEL_DONE) != 0u : (mode == 2u) ? (d.isl_state[vc.island] & ISL_POS_SOLVED) != 0u : false;
    for (unsigned spins = 0;; ++spins) {
        if (!done) {
            float4 a = flow_ld(&rows[vc.bodyA]), b = flow_ld(&rows[vc.bodyB]);
            if ((!vc.dynA || f2u(a.w) == tA) && (!vc.dynB || f2u(b.w) == tB)) {
                if (!skip) {
                    if (mode == 0u) warm_start_core(vc, a, b);
                    else if (mode == 1u) { float4 imp; result = solve_velocity_core(vc, a, b, imp); d.vc8[i] = imp; did = true; }
                    else { result = solve_position_core(vc, pc, a, b, rate, slop, maxCorr); did = true; }
                }
                if (vc.dynA) { a.w = u2f(tA + 1u); flow_st(&rows[vc.bodyA], a); }
                if (vc.dynB) { b.w = u2f(tB + 1u); flow_st(&rows[vc.bodyB], b); }
                done = true;
            }
        }
        if (PR_WARP_ALL(done)) break;
#ifdef PR_HOST_EMUL
        if (spins > 2u && getenv("PR_FLOW_DEBUG")) { const float4 a = rows[vc.bodyA], b = rows[vc.bodyB]; fprintf(stderr, "flow stall: mode %u row %u pass %u bodies %u(dyn %u) %u(dyn %u) want %u %u have %u %u flow %08x color %u\n", mode, i, pass, vc.bodyA, vc.dynA, vc.bodyB, vc.dynB, tA, tB, f2u(a.w), f2u(b.w), d.vc_flow[i], d.k_color[d.k_sorted[i]]); }
#endif
        if (spins > FLOW_SPIN_LIMIT) { atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_FLOW_STALL)); break; } // (never seen: a guard against a hung GPU)
    }
    const unsigned label = valid ? vc.island : 0u;
    if (mode == 1u) island_max(d.isl_inc, label, f2u(result), did);
    else if (mode == 2u) island_min(d.isl_sep, label, float_flip(result), did);
}

PR_KERNEL k_solve_position_tail(DV d, unsigned tailStart, unsigned tailCount, unsigned nSegs, float rate, float slop, float maxCorr)
{
    PR_PDL_ENTER();
    tail_for_each_row(d, tailStart, tailCount, nSegs, [&](unsigned row) {
        bool did = false;
        const float sep = solve_position_row(d, rs_global(d), row, rate, slop, maxCorr, did);
        if (did) atomicMin(&d.isl_sep[d.vc_island[row]], float_flip(sep));
    });
}
PR_D unsigned isl_pos_check_root(const DV& d, unsigned i, unsigned iter, float minSeparation)
{ // AabbTreeWorld.cpp:1628-1644
    const unsigned st = d.isl_state[i];
    if (!(st & ISL_ACTIVE) || (st & ISL_POS_SOLVED)) return 0u;
    const unsigned sepFlip = d.isl_sep[i];
    if (sepFlip < d.isl_minSep[i]) d.isl_minSep[i] = sepFlip;
    const bool jointsOkay = d.isl_jflag[i] == 0u;
    unsigned still = 0u;
    if (jointsOkay && float_unflip(sepFlip) >= minSeparation) { d.isl_state[i] = st | ISL_POS_SOLVED; d.isl_posIters[i] = iter + 1u; }
    else still = 1u;
    d.isl_sep[i] = float_flip(PR_INF); d.isl_jflag[i] = 0u;
    return still;
}
PR_KERNEL k_isl_pos_check(DV d, unsigned nBodies, unsigned iter, float minSeparation)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned stillActive = 0u;
    if (i < nBodies && d.b_island[i] == i) stillActive = isl_pos_check_root(d, i, iter, minSeparation);
    const unsigned slot = warp_append(&d.cnt->nActivePos, stillActive != 0u);
    if (stillActive) d.isl_active[slot] = i; // (the velocity iterations are over: their list is free)
}
PR_KERNEL k_isl_pos_check_list(DV d, unsigned iter, float minSeparation)
{
    PR_PDL_ENTER();
    const unsigned n = d.cnt->nActivePos;
    for (unsigned k = PR_TID; k < n; k += PR_NTHREADS) isl_pos_check_root(d, d.isl_active[k], iter, minSeparation);
}

// ------------------------------------------------------------------ a12: write-back, sleep, stats
PR_KERNEL k_assign_impulses(DV d, unsigned nCons)
{
    PR_PDL_ENTER(); // AssignImpulses, AabbTreeWorld.cpp:208-230
    const unsigned i = PR_TID;
    if (i < nCons) {
        const float4 head = d.vc1[i], imp = d.vc8[i];
        const unsigned cid = f2u(head.w), count = f2u(head.z) & 0xFFu;
        float4 m2 = d.m2[cid];
        if (count > 0u) { m2.x = imp.x; m2.y = imp.y; }
        if (count > 1u) { m2.z = imp.z; m2.w = imp.w; }
        d.m2[cid] = m2;
        if (d.ev_mask & PR_EVENT_POST_SOLVE) { // Report(postSolveContact, ...), AabbTreeWorld.cpp:1671-1674
            const uint4 ab = d.c_ab[cid]; // (the island's label from the bodies: the block solver's rows keep slots, not labels)
            const unsigned rA = d.b_island[ab.x];
            const unsigned r = (rA != INVALID) ? rA : d.b_island[ab.y];
            const unsigned it = (d.isl_state[r] & ISL_POS_SOLVED) ? d.isl_posIters[r] - 1u : 255u; // StepConf::InvalidIteration
            emit_impulse_event(d, cid, count, it, 0u, imp);
        }
    }
}
PR_KERNEL k_writeback_bodies(DV d, unsigned nBodies, pr_step_conf conf)
{
    PR_PDL_ENTER(); // AabbTreeWorld.cpp:1652-1664 + UpdateUnderActiveTimes :395-417
    const unsigned i = PR_TID;
    bool member = false; unsigned label = 0u; unsigned uaFlip = 0xFFFFFFFFu;
    if (i < nBodies) {
        const unsigned r = d.b_island[i];
        if (r != INVALID) {
            const float4 sv = d.s_vel[i], sp = d.s_pos[i];
            d.b_vel[i] = make_float4(sv.x, sv.y, sv.z, 0.0f);
            float4 p1 = d.b_pos1[i];
            if (!(p1.x == sp.x && p1.y == sp.y) || p1.z != sp.z) {
                p1.x = sp.x; p1.y = sp.y; p1.z = sp.z;
                const Xf xf = get_xf(pos_from4(sp), d.b_lc[i]);
                d.b_xf[i] = make_float4(xf.p.x, xf.p.y, xf.q.c, xf.q.s);
                d.b_moved[i] = 1u;
            }
            const unsigned f = d.b_flags[i];
            // IsUnderActive, d2/Math.hpp:283-289
            const float lin = magsq(v2(sv.x, sv.y)), ang = square(sv.z);
            const bool ua = (ang <= square(conf.angular_sleep_tolerance)) && (lin <= square(conf.linear_sleep_tolerance));
            const float t = ((f & PR_BODY_AUTOSLEEP) && ua) ? p1.w + conf.delta_time : 0.0f;
            p1.w = t;
            d.b_pos1[i] = p1;
            member = true; label = r; uaFlip = float_flip(t);
        }
    }
    island_min(d.isl_minUA, label, uaFlip, member);
}
PR_KERNEL k_sleep_and_stats(DV d, unsigned nBodies, float minStillTime)
{
    PR_PDL_ENTER(); // Sleepem :419-430 ; Update(RegStepStats&, IslandStats) StepStats.cpp:29-38
    const unsigned i = PR_TID;
    unsigned slept = 0u, solved = 0u, posIt = 0u, velIt = 0u, nb = 0u, sepFlip = 0xFFFFFFFFu, incBits = 0u;
    if (i < nBodies) {
        const unsigned r = d.b_island[i];
        if (r != INVALID) {
            const unsigned st = d.isl_state[r];
            if ((st & ISL_POS_SOLVED) && float_unflip(d.isl_minUA[r]) >= minStillTime) {
                const unsigned f = d.b_flags[i];
                if (f_awake(f) && (!f_speedable(f) || (f & PR_BODY_AUTOSLEEP))) { // Body::UnsetAwake, Body.cpp:173-182
                    d.b_flags[i] = f & ~static_cast<unsigned>(PR_BODY_AWAKE);
                    reinterpret_cast<float*>(&d.b_pos1[i])[3] = 0.0f;
                    d.b_vel[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    slept = 1u;
                }
            }
            if (r == i) {
                solved = (st & ISL_POS_SOLVED) ? 1u : 0u;
                posIt = d.isl_posIters[i]; velIt = d.isl_velIters[i]; nb = d.isl_nbodies[i] + d.isl_nstat[i];
                sepFlip = d.isl_minSep[i]; incBits = d.isl_maxInc[i];
            }
        }
    }
    block_add(&d.cnt->bodiesSlept, slept); block_add(&d.cnt->islandsSolved, solved);
    block_add(&d.cnt->sumPosIters, posIt); block_add(&d.cnt->sumVelIters, velIt);
    block_max(&d.cnt->maxIslandBodies, nb); block_min(&d.cnt->minSepFlip, sepFlip); block_max(&d.cnt->maxIncBits, incBits);
}

// The rare recount after an overflow of the (island, body) pair set: the step's constraint rows and joint work data still
// name every island's guests (a row's body without a label; labels live until the next step's k_isl_init).
PR_KERNEL k_isl_recount_statics(DV d, unsigned nCons, unsigned nJoints)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nCons) { // (bodies and label through the row's contact: the block solver's rows keep slots)
        const uint4 ab = d.c_ab[f2u(d.vc1[i].w)];
        const unsigned rA = d.b_island[ab.x], rB = d.b_island[ab.y];
        if (rA == INVALID || rB == INVALID) island_count_static(d, rA != INVALID ? rA : rB, rA == INVALID ? ab.x : ab.y);
    }
    if (i < nJoints) {
        const unsigned isl = d.jt_work[i].island;
        if (isl != INVALID) {
            const unsigned a = d.jt[i].body_a, b = d.jt[i].body_b;
            if (d.b_island[a] == INVALID) island_count_static(d, isl, a);
            if (d.b_island[b] == INVALID) island_count_static(d, isl, b);
        }
    }
}
PR_KERNEL k_isl_max_bodies(DV d, unsigned nBodies)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned nb = 0u;
    if (i < nBodies && d.b_island[i] == i) nb = d.isl_nbodies[i] + d.isl_nstat[i];
    block_max(&d.cnt->maxIslandBodies, nb);
}

// ------------------------------------------------------------------ export helpers
PR_KERNEL k_export_contacts(DV d, unsigned nOrder, pr_contact* out, pr_manifold* mout)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const unsigned long long key = d.c_key[id];
        const uint4 ab = d.c_ab[id], ids = d.c_ids[id];
        const float4 mat = d.c_mat[id];
        const unsigned fl = d.c_flags[id];
        pr_contact c;
        c.id = id; c.key_min = static_cast<unsigned>(key >> 32); c.key_max = static_cast<unsigned>(key & 0xFFFFFFFFull);
        c.body_a = ab.x; c.shape_a = ids.x; c.child_a = ids.y; c.body_b = ab.y; c.shape_b = ids.z; c.child_b = ids.w;
        c.friction = mat.x; c.restitution = mat.y; c.tangent_speed = mat.z; c.toi = mat.w;
        c.toi_count = fl >> 8; c.flags = fl & 0xFFu;
        // a SetBody that changed the transformation flags the body's contacts for updating (AabbTreeWorld.cpp:2787-2790); on the device that is
        // the body's moved mark until the next Step's UpdateContacts consumes it: the value read back carries the flag either way
        if (d.b_moved[ab.x] | d.b_moved[ab.y]) c.flags |= PR_CONTACT_DIRTY;
        out[i] = c;
        if (mout) {
            const float4 q0 = d.m0[id], q1 = d.m1[id], q2 = d.m2[id]; const uint4 q3 = d.m3[id];
            pr_manifold m;
            m.ln_x = q0.x; m.ln_y = q0.y; m.lp_x = q0.z; m.lp_y = q0.w; m.p0_x = q1.x; m.p0_y = q1.y; m.p1_x = q1.z; m.p1_y = q1.w;
            m.jn0 = q2.x; m.jt0 = q2.y; m.jn1 = q2.z; m.jt1 = q2.w; m.cf0 = q3.x; m.cf1 = q3.y; m.type = q3.z; m.count = q3.w;
            mout[i] = m;
        }
    }
}
PR_KERNEL k_import_contacts(DV d, unsigned n, const pr_contact* in, const pr_manifold* min, const unsigned* pxOfTree)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const pr_contact c = in[i];
        const pr_manifold m = min[i];
        const unsigned id = c.id;
        const unsigned pa = pxOfTree[c.key_min], pb = pxOfTree[c.key_max];
        d.c_key[id] = (static_cast<unsigned long long>(c.key_min) << 32) | c.key_max;
        d.c_ab[id] = make_uint4(c.body_a, c.body_b, d.p_cg[pa], d.p_cg[pb]);
        d.c_px[id] = make_uint2(pa, pb);
        d.c_ids[id] = make_uint4(c.shape_a, c.child_a, c.shape_b, c.child_b);
        d.c_mat[id] = make_float4(c.friction, c.restitution, c.tangent_speed, c.toi);
        d.c_flags[id] = (c.flags & 0xFFu) | (c.toi_count << 8);
        if (m.count == 0u && m.type == MF_UNSET) manifold_clear(d, id); // canonical empty manifold (update_contact relies on it)
        else {
            d.m0[id] = make_float4(m.ln_x, m.ln_y, m.lp_x, m.lp_y); d.m1[id] = make_float4(m.p0_x, m.p0_y, m.p1_x, m.p1_y);
            d.m2[id] = make_float4(m.jn0, m.jt0, m.jn1, m.jt1); d.m3[id] = make_uint4(m.cf0, m.cf1, m.type, m.count);
        }
        d.c_toiIsl[id] = 0u; d.c_color[id] = COLOR_UNSET; d.c_rstamp[id] = 0u;
        d.c_order[i] = id;
    }
}
// SetShape, AabbTreeWorld.cpp:1297-1301 and :1346-1354: wake the listed bodies; flag the shape's contacts for updating and wake
// their bodies
PR_KERNEL k_wake_bodies(DV d, unsigned n, const unsigned* ids)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) body_set_awake(d, ids[i]);
}
PR_KERNEL k_shape_flag_contacts(DV d, unsigned nOrder, unsigned shape)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const uint4 ids = d.c_ids[id];
        if (ids.x == shape || ids.z == shape) {
            atomicOr(&d.c_flags[id], static_cast<unsigned>(PR_CONTACT_DIRTY));
            const uint4 ab = d.c_ab[id];
            body_set_awake(d, ab.x); body_set_awake(d, ab.y);
        }
    }
}
// SetContact(AabbTreeWorld&, ContactID, Contact), AabbTreeWorld.cpp:2797-2831: phase 0 validates every row against the device's
// contact (what the reference throws for, as a code in *err: the least row's code wins), phase 1 applies "elem = value" -- the
// mutable part: the flag bits, friction, restitution, tangent speed.
enum { SETC_OK = 0, SETC_NO_SUCH = 1, SETC_DESTROYED = 2, SETC_IS_DESTROYED = 3, SETC_CONTACTABLE_A = 4, SETC_CONTACTABLE_B = 5,
       SETC_IMPENETRABLE = 6, SETC_SENSOR = 7, SETC_TOI = 8, SETC_TOI_COUNT = 9, SETM_TYPE = 10, SETM_COUNT = 11 };
PR_KERNEL k_set_contacts(DV d, unsigned n, const pr_contact* rows, unsigned poolSize, unsigned phase, unsigned long long* err)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i >= n) return;
    const pr_contact c = rows[i];
    const unsigned id = c.id;
    if (phase == 0u) {
        unsigned code = SETC_OK;
        if (id >= poolSize) code = SETC_NO_SUCH;
        else {
            const unsigned fl = d.c_flags[id];
            const uint4 ab = d.c_ab[id], ids = d.c_ids[id];
            if (d.c_key[id] == ~0ull) code = SETC_DESTROYED; // freed id (ObjectPool::FindFree): WasDestroyed
            else if ((fl & PR_CONTACT_DESTROYED) != (c.flags & PR_CONTACT_DESTROYED)) code = SETC_IS_DESTROYED;
            else if (ab.x != c.body_a || ids.x != c.shape_a || ids.y != c.child_a) code = SETC_CONTACTABLE_A;
            else if (ab.y != c.body_b || ids.z != c.shape_b || ids.w != c.child_b) code = SETC_CONTACTABLE_B;
            else if ((fl & PR_CONTACT_IMPENETRABLE) != (c.flags & PR_CONTACT_IMPENETRABLE)) code = SETC_IMPENETRABLE;
            else if ((fl & PR_CONTACT_SENSOR) != (c.flags & PR_CONTACT_SENSOR)) code = SETC_SENSOR;
            else if ((fl & PR_CONTACT_TOI) != (c.flags & PR_CONTACT_TOI) || ((fl & PR_CONTACT_TOI) && d.c_mat[id].w != c.toi)) code = SETC_TOI;
            else if ((fl >> 8) != c.toi_count) code = SETC_TOI_COUNT;
        }
        if (code != SETC_OK) atomicMin(err, (static_cast<unsigned long long>(i) << 32) | code);
    } else {
        const float toi = d.c_mat[id].w;
        d.c_mat[id] = make_float4(c.friction, c.restitution, c.tangent_speed, toi);
        d.c_flags[id] = (c.flags & 0xFFu) | (d.c_flags[id] & ~0xFFu);
    }
}
// SetManifold(AabbTreeWorld&, ContactID, const Manifold&), AabbTreeWorld.cpp:2833-2846 (same two phases)
PR_KERNEL k_set_manifolds(DV d, unsigned n, const unsigned* ids, const pr_manifold* rows, unsigned poolSize, unsigned phase, unsigned long long* err)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i >= n) return;
    const unsigned id = ids[i];
    const pr_manifold m = rows[i];
    if (phase == 0u) {
        unsigned code = SETC_OK;
        if (id >= poolSize) code = SETC_NO_SUCH;
        else if (d.c_key[id] == ~0ull) code = SETC_DESTROYED;
        else { const uint4 m3 = d.m3[id]; if (m3.z != m.type) code = SETM_TYPE; else if (m3.w != m.count) code = SETM_COUNT; }
        if (code != SETC_OK) atomicMin(err, (static_cast<unsigned long long>(i) << 32) | code);
    } else {
        d.m0[id] = make_float4(m.ln_x, m.ln_y, m.lp_x, m.lp_y); d.m1[id] = make_float4(m.p0_x, m.p0_y, m.p1_x, m.p1_y);
        d.m2[id] = make_float4(m.jn0, m.jt0, m.jn1, m.jt1); d.m3[id] = make_uint4(m.cf0, m.cf1, m.type, m.count);
    }
}
PR_KERNEL k_export_proxies(DV d, unsigned n, pr_proxy* out)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const float4 f = d.p_fat[i];
        pr_proxy p;
        p.tree_id = d.p_tree[i]; p.body = d.p_body[i]; p.shape = d.p_shape[i]; p.child = d.p_child[i];
        p.min_x = f.x; p.min_y = f.y; p.max_x = f.z; p.max_y = f.w;
        out[i] = p;
    }
}
PR_KERNEL k_fill_u32(unsigned* p, unsigned n, unsigned v) { PR_PDL_ENTER(); const unsigned i = PR_TID; if (i < n) p[i] = v; }
PR_KERNEL k_fill_u64(unsigned long long* p, unsigned n, unsigned long long v) { PR_PDL_ENTER(); const unsigned i = PR_TID; if (i < n) p[i] = v; }
PR_KERNEL k_gather_u32(const unsigned* src, const unsigned* idx, unsigned* dst, unsigned n) { PR_PDL_ENTER(); const unsigned i = PR_TID; if (i < n) dst[i] = src[idx[i]]; }
PR_KERNEL k_solve_order(DV d, unsigned nCons, unsigned* ids, unsigned* colors, const unsigned* sortedColor)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nCons) { ids[i] = d.k_contact[d.k_sorted[i]]; colors[i] = sortedColor[i] & 127u; } // (block solver keys: world << 7 | colour)
}

// standalone CollideShapes over arrays (pr_collide_shapes)
PR_KERNEL k_collide_batch(unsigned n, const pr_child* children, const float2* verts, const float2* normals,
                          const unsigned* aChild, const float4* xfA, const unsigned* bChild, const float4* xfB,
                          float linearSlop, float maxCirclesRatio, pr_manifold* out)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const pr_child ca = children[aChild[i]], cb = children[bChild[i]];
        ProxyView A, B;
        A.radius = ca.vertex_radius; A.count = ca.vertex_count; A.v = verts + ca.vertex_offset; A.n = normals + ca.vertex_offset;
        B.radius = cb.vertex_radius; B.count = cb.vertex_count; B.v = verts + cb.vertex_offset; B.n = normals + cb.vertex_offset;
        const Mf m = collide_shapes(A, xf_from4(xfA[i]), B, xf_from4(xfB[i]), linearSlop, maxCirclesRatio);
        pr_manifold o;
        o.ln_x = m.ln.x; o.ln_y = m.ln.y; o.lp_x = m.lp.x; o.lp_y = m.lp.y; o.p0_x = m.p0.x; o.p0_y = m.p0.y; o.p1_x = m.p1.x; o.p1_y = m.p1.y;
        o.jn0 = o.jt0 = o.jn1 = o.jt1 = 0.0f; o.cf0 = m.cf0; o.cf1 = m.cf1; o.type = m.type; o.count = m.count;
        out[i] = o;
    }
}

} // namespace pr
#include "pr_toi_phase.cuh"
#include "pr_block_solve.cuh"
#endif
