// pr_toi_phase.cuh -- device TOI phase: SolveToi(conf) and callees, AabbTreeWorld.cpp:1684-2142.
//
// The reference loop is serial: compute contact TOIs, take the globally soonest, advance its two
// bodies, build a mini island from their contact lists, solve it, re-synchronise, repeat.
// Here each loop turn is one ROUND of kernels over device-resident state:
//   k_toi_alpha / k_toi_advance_bodies / k_toi_compute   UpdateContactTOIs (:1684) for every contact
//                                                        without a valid TOI, one thread per contact
//   k_toi_claim / k_toi_select                           GetSoonestContact (:2870) -> event list
//   k_toi_solve                                          SolveToi(contact) (:1833), ProcessContactsForTOI
//                                                        (:2060), SolveToiViaGS (:1957): ONE thread per
//                                                        event walks the bodies' contact lists in
//                                                        m_contacts order and solves the mini island
//   k_sync_proxies + FindContacts/AddContacts            (:1797-1817)
// mode 0 (serial-exact): one event per round = the reference's order, results bit-identical.
// mode 1 (parallel): every event that is the soonest among all events touching its bodies or their
//   contact neighbours runs in the same round (their body sets are disjoint, so they commute with
//   each other; an event that an earlier one would have influenced through a chain of later events is
//   the documented source of drift -- BASELINE north_star: "drift reported, not gated").
#ifndef PR_TOI_PHASE_CUH
#define PR_TOI_PHASE_CUH

namespace pr {

// ------------------------------------------------------------------ body -> contact adjacency (m_bodyContacts)
// CSR over the set R of speedable bodies that own a TOI-candidate (impenetrable, non-sensor) contact: only
// those can be advanced, walked (ProcessContactsForTOI) or flagged (FlagForUpdating) by a TOI event; a
// 1M-body pile has a few thousand of them. Every segment has slack so contacts created during the TOI rounds
// are appended in place (k_create_contacts); a body's segment is put into m_contacts order by an insertion
// sort right before an event walks it (adj_sort_body).
PR_D bool toi_static_eligible(unsigned flags)
{ // the flag part of UpdateContactTOIs' filter (:1696) that never changes during a step
    return (flags & PR_CONTACT_ENABLED) && !(flags & PR_CONTACT_SENSOR) && (flags & PR_CONTACT_IMPENETRABLE);
}
// pass 1: m_contacts position, TOI candidate list, and the set R of speedable bodies that own at least one
// candidate contact -- only those bodies can ever be advanced, walked or flagged by a TOI event
PR_KERNEL k_toi_list_build(DV d, unsigned nOrder)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    bool listed = false; unsigned id = 0u;
    if (i < nOrder) {
        id = d.c_order[i];
        d.c_pos[id] = i;
        listed = toi_static_eligible(d.c_flags[id]);
        if (listed) {
            const uint4 ab = d.c_ab[id];
            if (f_speedable(d.b_flags[ab.x])) d.b_inR[ab.x] = 1u;
            if (f_speedable(d.b_flags[ab.y])) d.b_inR[ab.y] = 1u;
        }
    }
    const unsigned slot = warp_append(&d.cnt->toiListN, listed);
    if (listed) d.toi_list[slot] = id;
}
PR_KERNEL k_adj_count(DV d, unsigned nOrder)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const uint4 ab = d.c_ab[d.c_order[i]];
        if (d.b_inR[ab.x]) atomicAdd(&d.adj_cnt[ab.x], 1u);
        if (d.b_inR[ab.y]) atomicAdd(&d.adj_cnt[ab.y], 1u);
    }
}
PR_KERNEL k_adj_cap(DV d, unsigned nBodies)
{
    PR_PDL_ENTER();
    const unsigned b = PR_TID;
    if (b < nBodies) {
        const unsigned c = d.adj_cnt[b];
        d.adj_cnt[b] = d.b_inR[b] ? c + ((c >> 2) > 4u ? (c >> 2) : 4u) : 0u; // capacity (scan input)
        d.adj_n[b] = 0u;
    }
}
PR_KERNEL k_adj_fill(DV d, unsigned nOrder)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const uint4 ab = d.c_ab[id];
        if (d.b_inR[ab.x]) d.adj_list[d.adj_off[ab.x] + atomicAdd(&d.adj_n[ab.x], 1u)] = id;
        if (d.b_inR[ab.y]) d.adj_list[d.adj_off[ab.y] + atomicAdd(&d.adj_n[ab.y], 1u)] = id;
    }
}
PR_D unsigned adj_end(const DV& d, unsigned b)
{
    const unsigned lo = d.adj_off[b], cap = d.adj_off[b + 1u] - lo, n = d.adj_n[b];
    return lo + (n < cap ? n : cap);
}
PR_D void adj_sort_body(const DV& d, unsigned b)
{
    const unsigned lo = d.adj_off[b], hi = adj_end(d, b);
    for (unsigned i = lo + 1u; i < hi; ++i) {
        const unsigned id = d.adj_list[i];
        const unsigned key = d.c_pos[id];
        unsigned j = i;
        while (j > lo && d.c_pos[d.adj_list[j - 1u]] > key) { d.adj_list[j] = d.adj_list[j - 1u]; --j; }
        d.adj_list[j] = id;
    }
}
PR_KERNEL k_adj_sort(DV d, unsigned nBodies)
{
    PR_PDL_ENTER();
    const unsigned b = PR_TID;
    if (b < nBodies && f_speedable(d.b_flags[b])) adj_sort_body(d, b);
}
// AddContacts' m_bodyContacts push_back (:2520-2523) for a contact created while the table is live
PR_D void adj_append(const DV& d, unsigned b, unsigned id, bool eligible)
{
    if (d.b_inR[b]) {
        const unsigned slot = d.adj_off[b] + atomicAdd(&d.adj_n[b], 1u);
        if (slot < d.adj_off[b + 1u]) d.adj_list[slot] = id;
        else atomicOr(&d.cnt->adjOverflow, 1u); // host rebuilds the table before its next use
    } else if (eligible && f_speedable(d.b_flags[b])) atomicOr(&d.cnt->adjOverflow, 1u); // b joins R: rebuild
}

// ------------------------------------------------------------------ UpdateContactTOIs (:1684-1750)
PR_D bool toi_eligible(const DV& d, unsigned flags, uint4 ab, unsigned maxSubSteps, bool& atMax)
{
    atMax = false;
    if (flags & PR_CONTACT_TOI) return false;
    if (!(flags & PR_CONTACT_ENABLED) || (flags & PR_CONTACT_SENSOR) || !(flags & PR_CONTACT_IMPENETRABLE)) return false;
    if ((flags >> 8) >= maxSubSteps) { atMax = true; return false; }
    return f_awake(d.b_flags[ab.x]) || f_awake(d.b_flags[ab.y]);
}
// the TOI phase's statistics and round state, reset on the device (the host used to read the counters, clear these and write them back:
// a stream drain per Step)
PR_KERNEL k_toi_step_counters(DV d)
{
    PR_PDL_ENTER();
    if (PR_TID != 0u) return;
    Counters* c = d.cnt;
    c->toiAtMaxSub = c->toiUpdated = c->toiMaxDistIters = c->toiMaxToiIters = c->toiMaxRootIters = 0u;
    c->toiMinSepFlip = float_flip(PR_INF);
    c->toiMaxIncBits = 0u; c->toiIslandsSolved = c->toiSumPosIters = c->toiSumVelIters = 0u;
    c->toiIslandsFound = c->toiUpdTouching = c->toiSkipTouching = 0u;
    c->proxiesMoved = 0u; c->contactsUpdated = 0u; c->toiDeferred = 0u;
    c->toiRoundId = 0u; c->toiReN[0] = c->toiReN[1] = 0u; c->toiAtMaxLive = 0u; c->toiForceFull = 0u; c->toiHalt = TB_RUN;
}
PR_KERNEL k_toi_phase_init(DV d, unsigned nBodies)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) { d.b_toiAlpha[i] = f2u(d.b_pos0[i].w); d.b_claim[i] = ~0ull; d.b_toiStamp[i] = 0u; d.b_toiIsl[i] = 0u; }
}
// alpha0 = max(alphaA, alphaB) of every contact about to get a TOI (:1717): raise both bodies' targets.
// Invariant outside k_toi_alpha..k_toi_advance: b_toiAlpha[b] == bits(alpha0 of b).
// (thread 0 also opens the round: it resets the round's counters, or -- batched rounds -- halts the rest of the batch when
// the previous round overflowed an adjacency segment or raised an error. A halt that becomes visible to only some threads of
// THIS kernel is harmless: the atomicMax below is idempotent and k_toi_advance, which consumes it, is skipped.)
// a contact that must get its TOI (re)computed in the round after the current one: stamped so that it is listed once
PR_D void toi_mark_recompute(const DV& d, unsigned cid)
{
    const unsigned next = d.cnt->toiRoundId + 1u;
    if (atomicExch(&d.c_rstamp[cid], next) != next) {
        const unsigned slot = atomicAdd(&d.cnt->toiReN[next & 1u], 1u);
        ((next & 1u) ? d.toi_re1 : d.toi_re0)[slot] = cid;
    }
}
PR_D void toi_open_round(const DV& d, unsigned roundId)
{ // one thread: the counters of a round, and the (empty) list of what the NEXT round will have to recompute
    d.cnt->toiRoundId = roundId;
    d.cnt->toiMinKey = ~0ull; d.cnt->toiEvents = 0u; d.cnt->toiScratchTop = 0u; d.cnt->syncN = 0u;
    d.cnt->toiReN[(roundId + 1u) & 1u] = 0u;
}
PR_D void toi_alpha_one(const DV& d, unsigned id, unsigned maxSubSteps)
{
    const uint4 ab = d.c_ab[id];
    bool atMax;
    if (toi_eligible(d, d.c_flags[id], ab, maxSubSteps, atMax)) {
        const float a0 = stdmax(d.b_pos0[ab.x].w, d.b_pos0[ab.y].w);
        atomicMax(&d.b_toiAlpha[ab.x], f2u(a0)); // alpha in [0,1): non-negative floats order like uints
        atomicMax(&d.b_toiAlpha[ab.y], f2u(a0));
    }
}
PR_KERNEL k_toi_alpha(DV d, unsigned maxSubSteps, unsigned round)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (d.cnt->toiHalt) return;
    if (i == 0u) {
        if (d.cnt->adjOverflow) d.cnt->toiHalt = TB_ADJ;
        else if (d.cnt->error) d.cnt->toiHalt = TB_ERROR;
        else { // (round == 0: the device numbers the rounds itself -- list rounds enqueued behind single-block late rounds)
            toi_open_round(d, round ? round : d.cnt->toiRoundId + 1u); d.cnt->toiCandN = 0u; d.cnt->toiAtMaxLive = 0u; d.cnt->toiForceFull = 0u;
        }
    }
    if (i < d.cnt->toiListN) toi_alpha_one(d, d.toi_list[i], maxSubSteps);
}
PR_D void advance0(float4& p0, float4 p1, float alpha)
{ // Advance0(Sweep, alpha), Sweep.cpp:30-34
    const float beta = (alpha - p0.w) / (1.0f - p0.w);
    Pos a; a.l = v2(p0.x, p0.y); a.a = p0.z;
    Pos b; b.l = v2(p1.x, p1.y); b.a = p1.z;
    const Pos r = lerp_pos(a, b, beta);
    p0 = make_float4(r.l.x, r.l.y, r.a, alpha);
}
// Advance0(bA, alpha0); Advance0(bB, alpha0) (:1718-1719): once per body per round (stamp = round id)
PR_D void toi_advance_one(const DV& d, unsigned id, unsigned maxSubSteps, unsigned round)
{
    const uint4 ab = d.c_ab[id];
    bool atMax;
    if (toi_eligible(d, d.c_flags[id], ab, maxSubSteps, atMax)) {
        const unsigned two[2] = {ab.x, ab.y};
        for (int s = 0; s < 2; ++s) {
            const unsigned b = two[s];
            const unsigned t = d.b_toiAlpha[b];
            if (atomicExch(&d.b_toiStamp[b], round) != round) {
                float4 p0 = d.b_pos0[b];
                if (f2u(p0.w) != t) { advance0(p0, d.b_pos1[b], u2f(t)); d.b_pos0[b] = p0; }
            }
        }
    }
}
PR_KERNEL k_toi_advance(DV d, unsigned maxSubSteps, unsigned round)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (d.cnt->toiHalt) return;
    if (i < d.cnt->toiListN) toi_advance_one(d, d.toi_list[i], maxSubSteps, round ? round : d.cnt->toiRoundId);
}
PR_D float normalized_angle(float value)
{ // GetNormalized(Angle), Math.cpp:30-93 (trunc variant)
    const float pi = 3.14159265358979323846264338327950288f;
    const float twoPi = 2.0f * pi;
    const float rTwoPi = 1.0f / twoPi;
    float a = value;
    a = a - truncf(a * rTwoPi) * twoPi;
    if (a >= pi) a -= pi * 2;
    else if (a < -pi) a += pi * 2;
    return a;
}
PR_D SweepD normalized_sweep(float4 p0, float4 p1, float2 lc)
{ // GetNormalized(Sweep), d2/Math.hpp:160-167
    SweepD s;
    s.pos0.l = v2(p0.x, p0.y); s.pos1.l = v2(p1.x, p1.y); s.lc = lc;
    const float pos0a = normalized_angle(p0.z);
    const float dd = p0.z - pos0a;
    s.pos0.a = pos0a;
    s.pos1.a = p1.z - dd;
    return s;
}
PR_D unsigned long long toi_key(float toi, unsigned pos) { return (static_cast<unsigned long long>(f2u(toi)) << 32) | pos; }

struct ToiComputeOut { unsigned atMax, updated, distIt, toiIt, rootIt; unsigned long long key; };
// UpdateContactTOIs' body for one listed contact; key != ~0: the contact is a GetSoonestContact candidate (valid TOI < 1)
PR_D ToiComputeOut toi_compute_one(const DV& d, unsigned id, const pr_step_conf& conf)
{
    ToiComputeOut o; o.atMax = o.updated = o.distIt = o.toiIt = o.rootIt = 0u; o.key = ~0ull;
    unsigned flags = d.c_flags[id];
    const uint4 ab = d.c_ab[id];
    bool atMax;
    float toi = d.c_mat[id].w;
    if (toi_eligible(d, flags, ab, conf.max_sub_steps, atMax)) {
        float4 p0A = d.b_pos0[ab.x], p0B = d.b_pos0[ab.y];
        const float4 p1A = d.b_pos1[ab.x], p1B = d.b_pos1[ab.y];
        const float alpha0 = stdmax(p0A.w, p0B.w);
        if (p0A.w != alpha0) advance0(p0A, p1A, alpha0); // not stored: only reachable in parallel mode
        if (p0B.w != alpha0) advance0(p0B, p1B, alpha0);
        const SweepD sA = normalized_sweep(p0A, p1A, d.b_lc[ab.x]), sB = normalized_sweep(p0B, p1B, d.b_lc[ab.y]);
        const ToiResult out = toi_via_sat(child_view(d, ab.z), sA, child_view(d, ab.w), sB, conf.target_depth, conf.tolerance,
                                          conf.max_toi_iters, conf.max_toi_root_iters, conf.max_distance_iters);
        toi = (out.state == TOI_TOUCHING) ? stdmin(alpha0 + (1.0f - alpha0) * out.time, 1.0f) : 1.0f; // :1740
        reinterpret_cast<float*>(&d.c_mat[id])[3] = toi;
        flags |= PR_CONTACT_TOI;
        d.c_flags[id] = flags;
        o.updated = 1u; o.distIt = out.maxDistIters; o.toiIt = out.toiIters; o.rootIt = out.maxRootIters;
    } else if (atMax) o.atMax = 1u;
    if ((flags & PR_CONTACT_TOI) && toi < 1.0f) o.key = toi_key(toi, d.c_pos[id]);
    return o;
}
PR_KERNEL k_toi_compute(DV d, pr_step_conf conf)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    ToiComputeOut o; o.atMax = o.updated = o.distIt = o.toiIt = o.rootIt = 0u; o.key = ~0ull;
    if (!d.cnt->toiHalt && i < d.cnt->toiListN) {
        const unsigned id = d.toi_list[i];
        o = toi_compute_one(d, id, conf);
        if (o.key != ~0ull) d.toi_cand[atomicAdd(&d.cnt->toiCandN, 1u)] = id; // candidates, compacted for the selection kernels
    }
    warp_add(&d.cnt->toiAtMaxSub, o.atMax); warp_add(&d.cnt->toiAtMaxLive, o.atMax); warp_add(&d.cnt->toiUpdated, o.updated);
    warp_max(&d.cnt->toiMaxDistIters, o.distIt); warp_max(&d.cnt->toiMaxToiIters, o.toiIt); warp_max(&d.cnt->toiMaxRootIters, o.rootIt);
    if (o.key != ~0ull) atomicMin(&d.cnt->toiMinKey, o.key);
}

// ------------------------------------------------------------------ event selection
// the bodies an event on contact (ab) can advance, walk or flag: its own two speedable bodies and, through the accelerable
// ones' contact lists, the bodies ProcessContactsForTOI can advance / read (:2079-2087: same filter). One WARP per
// candidate: the lanes take the adjacency entries (a body of a pile has 6-10 of them, each a chain of four dependent
// loads -- one thread walking them all made these kernels 12 us each however few candidates there were).
template <class F> PR_D void toi_for_each_claim_body(const DV& d, uint4 ab, F&& fn)
{
    const unsigned two[2] = {ab.x, ab.y};
    for (int s = 0; s < 2; ++s) {
        const unsigned b = two[s];
        const unsigned f = d.b_flags[b];
        if (f_speedable(f) && PR_LANE == 0u) fn(b);
        if (f_accelerable(f)) {
            const bool bodyImpenetrable = (f & PR_BODY_IMPENETRABLE) != 0u;
            for (unsigned k = d.adj_off[b] + PR_LANE, ke = adj_end(d, b); k < ke; k += PR_WARP) {
                const unsigned c2 = d.adj_list[k];
                if (d.c_flags[c2] & PR_CONTACT_SENSOR) continue;
                const uint4 o = d.c_ab[c2];
                const unsigned other = (o.x == b) ? o.y : o.x;
                const unsigned of = d.b_flags[other];
                if (f_speedable(of) && (bodyImpenetrable || (of & PR_BODY_IMPENETRABLE))) fn(other);
            }
        }
    }
}
PR_D void toi_claim_cand(const DV& d, unsigned id)
{
    const unsigned long long key = toi_key(d.c_mat[id].w, d.c_pos[id]);
    toi_for_each_claim_body(d, d.c_ab[id], [&](unsigned b) { atomicMin(&d.b_claim[b], key); });
}
PR_D void toi_unclaim_cand(const DV& d, unsigned id) { toi_for_each_claim_body(d, d.c_ab[id], [&](unsigned b) { d.b_claim[b] = ~0ull; }); }
PR_D void toi_select_cand(const DV& d, unsigned id, unsigned parallel)
{
    const unsigned long long key = toi_key(d.c_mat[id].w, d.c_pos[id]);
    bool win;
    if (parallel) {
        bool mine = true;
        toi_for_each_claim_body(d, d.c_ab[id], [&](unsigned b) { if (d.b_claim[b] != key) mine = false; });
        win = PR_WARP_ALL(mine);
    } else win = (key == d.cnt->toiMinKey);
    if (win && PR_LANE == 0u) {
        d.toi_events[atomicAdd(&d.cnt->toiEvents, 1u)] = id;
        if (d.toi_log) { // the schedule, for the parity oracle (pr_get_toi_schedule)
            const unsigned slot = atomicAdd(&d.cnt->toiLogN, 1u);
            if (slot < d.toi_logCap) d.toi_log[slot] = make_uint2(d.cnt->toiRoundId, id);
        }
    }
}
PR_KERNEL k_toi_claim(DV d)
{
    PR_PDL_ENTER();
    if (d.cnt->toiHalt) return;
    const unsigned n = d.cnt->toiCandN;
    for (unsigned c = PR_WARP_ID; c < n; c += PR_NWARPS) toi_claim_cand(d, d.toi_cand[c]);
}
PR_KERNEL k_toi_unclaim(DV d)
{
    PR_PDL_ENTER();
    if (d.cnt->toiHalt) return;
    const unsigned n = d.cnt->toiCandN;
    for (unsigned c = PR_WARP_ID; c < n; c += PR_NWARPS) toi_unclaim_cand(d, d.toi_cand[c]);
}
PR_KERNEL k_toi_select(DV d, unsigned parallel)
{
    PR_PDL_ENTER();
    if (d.cnt->toiHalt) return;
    const unsigned n = d.cnt->toiCandN;
    for (unsigned c = PR_WARP_ID; c < n; c += PR_NWARPS) toi_select_cand(d, d.toi_cand[c], parallel);
}

// ------------------------------------------------------------------ SolveToi(contact) (:1833-1955)
// SetAwake on a body that was asleep makes contacts eligible that no list of "contacts to recompute" holds: the next
// round must walk the full TOI list again (k_tb_late hands back to the list kernels)
PR_D void toi_wake(const DV& d, unsigned b)
{
    if (!(atomicOr(&d.b_flags[b], static_cast<unsigned>(PR_BODY_AWAKE)) & PR_BODY_AWAKE)) d.cnt->toiForceFull = 1u;
}
PR_D void toi_flag_body_contacts(const DV& d, unsigned b)
{ // FlagForUpdating(m_bodyContacts[b]), :459-465
    for (unsigned k = d.adj_off[b], ke = adj_end(d, b); k < ke; ++k) atomicOr(&d.c_flags[d.adj_list[k]], static_cast<unsigned>(PR_CONTACT_DIRTY));
}
struct SweepBk { float4 p0, p1, xf; };
PR_D SweepBk toi_backup(const DV& d, unsigned b) { SweepBk s; s.p0 = d.b_pos0[b]; s.p1 = d.b_pos1[b]; s.xf = d.b_xf[b]; return s; }
PR_D void toi_restore(const DV& d, unsigned b, const SweepBk& s)
{ // Body::SetSweep, Body.hpp:697-702 (m_xf = GetTransform1(sweep): the same bits it had)
    d.b_pos0[b] = s.p0; d.b_toiAlpha[b] = f2u(s.p0.w);
    float4 p1 = d.b_pos1[b]; p1.x = s.p1.x; p1.y = s.p1.y; p1.z = s.p1.z; d.b_pos1[b] = p1;
    d.b_xf[b] = s.xf;
}
PR_D bool pos_differs(float4 a, float4 b) { return !(a.x == b.x && a.y == b.y) || a.z != b.z; }
// Advance(Body&, toi), Body.hpp:1004-1009; returns whether pos0 or pos1 changed (:1866, :2095)
PR_D bool toi_advance_body(const DV& d, unsigned b, float toi)
{
    float4 p0 = d.b_pos0[b];
    float4 p1 = d.b_pos1[b];
    const float4 o0 = p0, o1 = p1;
    advance0(p0, p1, toi);
    p1.x = p0.x; p1.y = p0.y; p1.z = p0.z;
    d.b_pos0[b] = p0; d.b_pos1[b] = p1; d.b_toiAlpha[b] = f2u(toi);
    const Xf xf = get_xf(v2(p1.x, p1.y), rot_from_angle_call(p1.z), d.b_lc[b]);
    d.b_xf[b] = make_float4(xf.p.x, xf.p.y, xf.q.c, xf.q.s);
    return pos_differs(p0, o0) || pos_differs(p1, o1);
}
struct ToiIsland { unsigned base, nB, nC; };
PR_D bool toi_body_islanded(const DV& d, const ToiIsland& isl, unsigned b, bool speedable)
{
    if (speedable) return d.b_toiIsl[b] != 0u;
    for (unsigned k = 0; k < isl.nB; ++k) if (d.toi_islBodies[isl.base + k] == b) return true; // non-speedable: this event's list
    return false;
}
// Update(contactID) as the TOI phase calls it: ONE copy of the narrowphase per kernel (it used to be inlined four times into
// the event solve, 23 000 instructions of which a single lane walked a few hundred)
PR_DN void toi_update_contact(const DV& d, const pr_step_conf& conf, unsigned cid, unsigned cf) { update_contact(d, conf, cid, cf); }
// one entry of ProcessContactsForTOI's loop (:2079-2140), everything in the reference's order: used when the entries of a
// chunk may depend on each other (serial-exact mode, or two contacts of the chunk share their other body)
PR_D void toi_process_entry(const DV& d, const pr_step_conf& conf, unsigned id, bool bodyImpenetrable, unsigned cid, ToiIsland& isl,
                            float toi, bool exactMode, unsigned& updatedN, unsigned& skippedN)
{
    if (d.c_toiIsl[cid]) return;
    unsigned cf = d.c_flags[cid];
    if (cf & PR_CONTACT_SENSOR) return;
    const uint4 ab = d.c_ab[cid];
    const unsigned other = (ab.x == id) ? ab.y : ab.x;
    const unsigned of = d.b_flags[other];
    if (!(bodyImpenetrable || (of & PR_BODY_IMPENETRABLE))) return;
    const bool otherSpeedable = f_speedable(of);
    const bool otherIslanded = toi_body_islanded(d, isl, other, otherSpeedable);
    const SweepBk backup = toi_backup(d, other);
    const bool touchOther = !otherIslanded && (otherSpeedable || exactMode);
    if (touchOther) {
        if (toi_advance_body(d, other, toi)) toi_flag_body_contacts(d, other);
    }
    cf = d.c_flags[cid];
    if (cf & PR_CONTACT_DIRTY) { toi_update_contact(d, conf, cid, cf); ++updatedN; }
    else ++skippedN;
    cf = d.c_flags[cid];
    if (!(cf & PR_CONTACT_ENABLED) || !(cf & PR_CONTACT_TOUCHING)) {
        if (touchOther) toi_restore(d, other, backup);
        return;
    }
    d.toi_islContacts[isl.base + isl.nC++] = cid;
    d.c_toiIsl[cid] = 1u;
    if (!otherIslanded) {
        if (otherSpeedable) { toi_wake(d, other); d.b_toiIsl[other] = 1u; }
        d.toi_islBodies[isl.base + isl.nB++] = other;
    }
}
// ProcessContactsForTOI, :2060-2142 -- the whole WARP: a chunk of 32 entries of the body's contact list at a time. Each lane
// advances its entry's other body, updates the contact and restores the body if the contact turned out not to touch -- the
// entries are independent of each other as long as no two of them share a body that can be moved (then, and in serial-exact
// mode, where unspeedable bodies are advanced too, the chunk runs entry by entry) -- and the island lists are appended in lane
// order (= list order) by ballot ranks. isl and the two counts are warp-uniform on entry and on exit... the counts are
// per-lane partial sums (the caller adds them up).
PR_D void toi_process_contacts(const DV& d, const pr_step_conf& conf, unsigned id, ToiIsland& isl, float toi, bool exactMode,
                               unsigned& updatedN, unsigned& skippedN)
{
    const bool bodyImpenetrable = (d.b_flags[id] & PR_BODY_IMPENETRABLE) != 0u;
    const unsigned lo = d.adj_off[id], hi = adj_end(d, id);
    { // contacts appended since the table was built go to their m_bodyContacts position (rare: check first, all lanes)
        bool sorted = true;
        for (unsigned k = lo + 1u + PR_LANE; k < hi; k += PR_WARP) if (d.c_pos[d.adj_list[k - 1u]] > d.c_pos[d.adj_list[k]]) sorted = false;
        if (!PR_WARP_ALL(sorted)) { if (PR_LANE == 0u) adj_sort_body(d, id); }
        PR_SYNCWARP();
    }
    for (unsigned k0 = lo; k0 < hi; k0 += PR_WARP) {
        const unsigned k = k0 + PR_LANE;
        bool cand = false;
        unsigned cid = 0u, other = 0u, of = 0u;
        if (k < hi) {
            cid = d.adj_list[k];
            if (!d.c_toiIsl[cid] && !(d.c_flags[cid] & PR_CONTACT_SENSOR)) {
                const uint4 ab = d.c_ab[cid];
                other = (ab.x == id) ? ab.y : ab.x;
                of = d.b_flags[other];
                cand = bodyImpenetrable || (of & PR_BODY_IMPENETRABLE);
            }
        }
        const bool otherSpeedable = f_speedable(of);
        bool serial = exactMode;
#ifndef PR_HOST_EMUL
        if (!serial) { // two movable entries of the chunk with the same other body?
            const unsigned m = PR_BALLOT(cand && otherSpeedable);
            bool dup = false;
            if (cand && otherSpeedable) dup = __popc(__match_any_sync(m, other)) > 1;
            serial = PR_WARP_ANY(dup);
        }
#endif
        if (serial) {
            for (unsigned l = 0; l < PR_WARP; ++l) {
                if (PR_LANE == l && k < hi) toi_process_entry(d, conf, id, bodyImpenetrable, d.adj_list[k], isl, toi, exactMode, updatedN, skippedN);
                PR_SYNCWARP();
                isl.nB = PR_SHFL(isl.nB, l); isl.nC = PR_SHFL(isl.nC, l);
            }
            continue;
        }
        bool touching = false, otherIslanded = false;
        if (cand) {
            if (otherSpeedable) otherIslanded = d.b_toiIsl[other] != 0u; // (an unspeedable body is never touched here: looked up below)
            const bool touchOther = otherSpeedable && !otherIslanded;
            SweepBk backup;
            if (touchOther) {
                backup = toi_backup(d, other);
                if (toi_advance_body(d, other, toi)) toi_flag_body_contacts(d, other);
            }
            unsigned cf = d.c_flags[cid];
            if (cf & PR_CONTACT_DIRTY) { toi_update_contact(d, conf, cid, cf); ++updatedN; }
            else ++skippedN;
            cf = d.c_flags[cid];
            touching = (cf & PR_CONTACT_ENABLED) && (cf & PR_CONTACT_TOUCHING);
            if (!touching && touchOther) toi_restore(d, other, backup);
        }
        const unsigned mC = PR_BALLOT(touching);
        if (touching) { d.toi_islContacts[isl.base + isl.nC + __popc(mC & PR_LANEMASK_LT)] = cid; d.c_toiIsl[cid] = 1u; }
        bool addBody = false;
        if (touching) {
            if (otherSpeedable) addBody = !otherIslanded;
            else { addBody = true; for (unsigned q = 0; q < isl.nB; ++q) if (d.toi_islBodies[isl.base + q] == other) addBody = false; }
        }
#ifndef PR_HOST_EMUL
        { // an unspeedable body (the terrain) that several entries of the chunk touch joins the list once, with the first of them
            const unsigned m = PR_BALLOT(addBody && !otherSpeedable);
            if (addBody && !otherSpeedable && PR_LANE != static_cast<unsigned>(__ffs(__match_any_sync(m, other)) - 1)) addBody = false;
        }
#endif
        const unsigned mB = PR_BALLOT(addBody);
        if (addBody) {
            if (otherSpeedable) { toi_wake(d, other); d.b_toiIsl[other] = 1u; }
            d.toi_islBodies[isl.base + isl.nB + __popc(mB & PR_LANEMASK_LT)] = other;
        }
        isl.nC += __popc(mC); isl.nB += __popc(mB);
        PR_SYNCWARP();
    }
}

// A TOI event's mini island (rows + BodyConstraints) lives in the solving warp's slice of shared memory when it
// fits: the NGS/GS iterations are a serial chain of dependent loads, and a chain through HBM/L2 costs ~0.5 us a
// link. Larger islands fall back to the global scratch rows at isl.base.
constexpr unsigned TOI_SM_NC = 16u, TOI_SM_NB = 16u;
struct ToiSmem {
    float4 rows[13][TOI_SM_NC];
    float4 svel[TOI_SM_NB], spos[TOI_SM_NB];
    unsigned island[TOI_SM_NC], bodies[TOI_SM_NB];
    unsigned changed[2], pad[2];
};
// the two serial chains of an event (lane 0 runs them; the rows sit in shared memory)
// (measured on B200, 1M mixed scene: as out-of-line calls the two passes are 10-15% SLOWER -- the row store's pointers then
// live on the stack -- and a per-body cache of UnitVec::Get(angle) loses too: 7 950 vs 6 560 cycles per position pass)
PR_D float toi_position_pass(const DV& d, const RS& r, unsigned row0, unsigned nC, float rate, float slop, float maxCorr)
{
    float minSeparation = PR_INF;
    for (unsigned j = 0; j < nC; ++j) {
        bool did;
        minSeparation = stdmin(minSeparation, solve_position_row(d, r, row0 + j, rate, slop, maxCorr, did, false));
    }
    return minSeparation;
}
PR_D float toi_velocity_pass(const DV& d, const RS& r, unsigned row0, unsigned nC)
{
    float newInc = 0.0f;
    for (unsigned j = 0; j < nC; ++j) { bool did; newInc = stdmax(newInc, solve_velocity_row(d, r, row0 + j, did, false)); }
    return newInc;
}
// SolveToi(contact) for event e -- the whole WARP (see k_toi_solve): the two bodies, the entries of their contact lists, the
// island's bodies and rows are dealt out to the lanes; only the Gauss-Seidel iterations themselves run on lane 0
#ifdef PR_TOI_TIMING
#define PR_TT(k) do { if (e == 0u && PR_LANE == 0u) { const long long t_ = clock64(); atomicAdd(&d.cnt->dbg[k], static_cast<unsigned long long>(t_ - tt_)); tt_ = t_; } } while (0)
#else
#define PR_TT(k) do { } while (0)
#endif
PR_D void toi_solve_event(const DV& d, const pr_step_conf& conf, unsigned e, unsigned exactMode, unsigned scratchCap, ToiSmem* sm)
{
#ifdef PR_TOI_TIMING
    long long tt_ = clock64();
#endif
    const unsigned cid = d.toi_events[e];
    const uint4 ab = d.c_ab[cid];
    const unsigned A = ab.x, B = ab.y;
    const float toi = d.c_mat[cid].w;
    unsigned numUpdated = 0u, numSkipped = 0u; // (per-lane partial sums)
    const unsigned fA = d.b_flags[A], fB = d.b_flags[B];
    // island scratch (body list, contact list, VC/PC rows): sized by the two contact lists. An event that does
    // not fit this round's arena is left untouched (still the soonest of its neighbourhood) for the next round.
    ToiIsland isl;
    const unsigned need = 2u + (f_speedable(fA) ? d.adj_n[A] : 0u) + (f_speedable(fB) ? d.adj_n[B] : 0u);
    isl.base = 0u;
    if (PR_LANE == 0u) isl.base = atomicAdd(&d.cnt->toiScratchTop, need);
    isl.base = PR_SHFL(isl.base, 0);
    if (isl.base + need > scratchCap) {
        if (PR_LANE == 0u) {
            if (isl.base == 0u) atomicOr(&d.cnt->error, static_cast<unsigned>(ERR_TOI_SCRATCH));
            atomicAdd(&d.cnt->toiDeferred, 1u);
        }
        return;
    }
    {
        // lanes 0 and 1 advance A and B (the host emulation's single lane takes them one after the other)
        constexpr unsigned NS = (PR_WARP >= 2u) ? 1u : 2u;
        SweepBk bk[NS]; bool touch[NS];
        for (unsigned s = PR_LANE; s < 2u; s += PR_WARP) {
            const unsigned X = s ? B : A, fX = s ? fB : fA;
            bk[s % NS] = toi_backup(d, X);
            touch[s % NS] = f_speedable(fX) || exactMode;
            sm->changed[s] = (touch[s % NS] && toi_advance_body(d, X, toi)) ? 1u : 0u;
        }
        PR_SYNCWARP();
        for (unsigned s = 0; s < 2u; ++s) { // FlagForUpdating of a body that moved: its list over the lanes
            const unsigned X = s ? B : A;
            if (sm->changed[s]) for (unsigned k = d.adj_off[X] + PR_LANE, ke = adj_end(d, X); k < ke; k += PR_WARP) atomicOr(&d.c_flags[d.adj_list[k]], static_cast<unsigned>(PR_CONTACT_DIRTY));
        }
        PR_SYNCWARP();
        unsigned proceed = 0u;
        if (PR_LANE == 0u) {
            unsigned cf = d.c_flags[cid];
            if (cf & PR_CONTACT_DIRTY) { toi_update_contact(d, conf, cid, cf); ++numUpdated; }
            cf = d.c_flags[cid];
            cf &= ~static_cast<unsigned>(PR_CONTACT_TOI); // SetToi(contact, {})
            cf += 0x100u;                                  // IncrementToiCount
            d.c_flags[cid] = cf;
            if ((cf >> 8) == conf.max_sub_steps) atomicAdd(&d.cnt->toiAtMaxLive, 1u); // from now on UpdateContactTOIs counts it every round
            toi_mark_recompute(d, cid);                    // (its TOI is gone: next round looks at it again)
            reinterpret_cast<float*>(&d.c_mat[cid])[3] = 1.0f;
            proceed = ((cf & PR_CONTACT_ENABLED) && (cf & PR_CONTACT_TOUCHING)) ? 1u : 0u;
            if (!proceed) {
                atomicAdd(&d.cnt->toiUpdTouching, numUpdated);
                atomicAdd(&d.cnt->toiSkipTouching, numUpdated ? 0u : 1u);
            }
        }
        proceed = PR_SHFL(proceed, 0);
        if (!proceed) {
            for (unsigned s = PR_LANE; s < 2u; s += PR_WARP) if (touch[s % NS]) toi_restore(d, s ? B : A, bk[s % NS]);
            return;
        }
    }
    PR_TT(0);
    if (PR_LANE == 0u) {
        if (f_speedable(fA)) toi_wake(d, A);
        if (f_speedable(fB)) toi_wake(d, B);
        d.toi_islBodies[isl.base] = A; d.toi_islBodies[isl.base + 1u] = B; d.toi_islContacts[isl.base] = cid;
        if (f_speedable(fA)) d.b_toiIsl[A] = 1u;
        if (f_speedable(fB)) d.b_toiIsl[B] = 1u;
        d.c_toiIsl[cid] = 1u;
    }
    isl.nB = 2u; isl.nC = 1u;
    PR_SYNCWARP();
    for (unsigned s = 0; s < 2u; ++s) { // (a loop: one copy of the code)
        if (f_accelerable(s ? fB : fA)) toi_process_contacts(d, conf, s ? B : A, isl, toi, exactMode != 0u, numUpdated, numSkipped);
        PR_TT(1u + s);
    }
    numUpdated = PR_WARP_SUM(numUpdated); numSkipped = PR_WARP_SUM(numSkipped);

    // ---- SolveToiViaGS (:1957-2058) with subConf.deltaTime = (1 - toi) * deltaTime (:1952-1953)
    pr_step_conf sub = conf;
    sub.delta_time = (1.0f - toi) * conf.delta_time;
    const bool small = isl.nC <= TOI_SM_NC && isl.nB <= TOI_SM_NB;
    RS r = rs_global(d);
    unsigned row0 = isl.base;
    if (small) {
        r.vc0 = sm->rows[0]; r.vc1 = sm->rows[1]; r.vc2 = sm->rows[2]; r.vc3 = sm->rows[3]; r.vc4 = sm->rows[4];
        r.vc5 = sm->rows[5]; r.vc6 = sm->rows[6]; r.vc7 = sm->rows[7]; r.vc8 = sm->rows[8]; r.vc_island = sm->island;
        r.pc0 = sm->rows[9]; r.pc1 = sm->rows[10]; r.pc2 = sm->rows[11]; r.pc3 = sm->rows[12];
        r.s_vel = sm->svel; r.s_pos = sm->spos;
        row0 = 0u;
        for (unsigned k = PR_LANE; k < isl.nB; k += PR_WARP) { sm->bodies[k] = d.toi_islBodies[isl.base + k]; }
        PR_SYNCWARP();
    }
    // index of island body k (id b) in r.s_vel / r.s_pos, and of a contact's body there
    auto slot_of_k = [&](unsigned k, unsigned b) { return small ? k : b; };
    auto slot_of_body = [&](unsigned b) {
        if (!small) return b;
        unsigned k = 0u;
        while (k + 1u < isl.nB && sm->bodies[k] != b) ++k;
        return k;
    };
    float minSepAll = PR_INF;
    unsigned positionIters = conf.toi_position_iters, velocityIters = conf.toi_velocity_iters;
    unsigned solved = 0u;
    float maxIncAll = 0.0f;
    for (unsigned pass = 0; pass < 2u; ++pass) { // position pass, velocity pass (a loop: one copy of the row builder)
        if (pass == 0u) {
            for (unsigned k = PR_LANE; k < isl.nB; k += PR_WARP) { // GetBodyConstraints(h = 0)
                const unsigned b = d.toi_islBodies[isl.base + k];
                const float4 p1 = d.b_pos1[b];
                float4 v = body_constraint_velocity(d, b, d.b_vel[b], d.b_flags[b], 0.0f, sub);
                v.w = 0.0f;
                r.s_vel[slot_of_k(k, b)] = v;
                r.s_pos[slot_of_k(k, b)] = make_float4(p1.x, p1.y, p1.z, 0.0f);
            }
            PR_SYNCWARP();
        }
        for (unsigned j = PR_LANE; j < isl.nC; j += PR_WARP) { // (one row per lane)
            const unsigned c2 = d.toi_islContacts[isl.base + j];
            const uint4 ab2 = d.c_ab[c2];
            build_constraint_rows(d, r, c2, row0 + j, sub, 0.0f, 0u, slot_of_body(ab2.x), slot_of_body(ab2.y));
        }
        PR_SYNCWARP();
        if (pass == 0u) {
            PR_TT(3);
            if (PR_LANE == 0u) {
                for (unsigned it = 0; it < conf.toi_position_iters; ++it) {
                    const float minSeparation = toi_position_pass(d, r, row0, isl.nC, conf.toi_resolution_rate, conf.linear_slop, conf.max_linear_correction);
                    minSepAll = stdmin(minSepAll, minSeparation);
                    if (minSeparation >= conf.toi_min_separation) { positionIters = it + 1u; solved = 1u; break; }
                }
            }
            PR_SYNCWARP();
            PR_TT(4);
            for (unsigned k = PR_LANE; k < isl.nB; k += PR_WARP) { // "leap of faith": pos0 <- solved position (:2006-2012)
                const unsigned b = d.toi_islBodies[isl.base + k];
                if (f_speedable(d.b_flags[b])) {
                    const float4 sp = r.s_pos[slot_of_k(k, b)];
                    float4 p0 = d.b_pos0[b]; p0.x = sp.x; p0.y = sp.y; p0.z = sp.z; d.b_pos0[b] = p0;
                }
            }
        } else {
            if (PR_LANE == 0u) {
                for (unsigned it = 0; it < conf.toi_velocity_iters; ++it) {
                    const float newInc = toi_velocity_pass(d, r, row0, isl.nC);
                    if (newInc <= conf.toi_min_momentum) { velocityIters = it + 1u; break; }
                    maxIncAll = stdmax(maxIncAll, newInc);
                }
            }
            PR_SYNCWARP();
        }
    }
    PR_TT(5);
    positionIters = PR_SHFL(positionIters, 0);
    if (d.ev_mask & PR_EVENT_POST_SOLVE) { // Report(postSolveContact, island.contacts, velConstraints, positionIters), :2053-2055
        for (unsigned j = PR_LANE; j < isl.nC; j += PR_WARP) {
            const float4 head = r.vc1[row0 + j];
            emit_impulse_event(d, f2u(head.w), f2u(head.z) & 0xFFu, positionIters, 1u, r.vc8[row0 + j]);
        }
    }
    const float h = sub.delta_time;
    for (unsigned k0 = 0; k0 < isl.nB; k0 += PR_WARP) { // IntegratePositions + write-back (:2042-2052): a body per lane ...
        const unsigned k = k0 + PR_LANE;
        bool moved = false; unsigned b = 0u;
        if (k < isl.nB) {
            b = d.toi_islBodies[isl.base + k];
            const unsigned f = d.b_flags[b];
            if (f_speedable(f)) {
                const float4 v = r.s_vel[slot_of_k(k, b)];
                float4 sp = r.s_pos[slot_of_k(k, b)];
                const float2 tr = h * v2(v.x, v.y);
                const float rot = h * v.z;
                sp.x += tr.x; sp.y += tr.y; sp.z += rot;
                d.b_vel[b] = make_float4(v.x, v.y, v.z, 0.0f);
                float4 p1 = d.b_pos1[b];
                if (pos_differs(p1, sp)) {
                    p1.x = sp.x; p1.y = sp.y; p1.z = sp.z;
                    d.b_pos1[b] = p1;
                    const Xf xf = get_xf(v2(sp.x, sp.y), rot_from_angle_call(sp.z), d.b_lc[b]);
                    d.b_xf[b] = make_float4(xf.p.x, xf.p.y, xf.q.c, xf.q.s);
                    moved = true;
                }
            }
        }
        // ... then FlagForUpdating of the moved ones, each list over the lanes
        for (unsigned m = PR_BALLOT(moved); m; m &= m - 1u) {
            const unsigned bm = PR_SHFL(b, static_cast<unsigned>(__ffs(static_cast<int>(m)) - 1));
            for (unsigned q = d.adj_off[bm] + PR_LANE, qe = adj_end(d, bm); q < qe; q += PR_WARP) atomicOr(&d.c_flags[d.adj_list[q]], static_cast<unsigned>(PR_CONTACT_DIRTY));
        }
    }
    PR_SYNCWARP();
    // ---- :1797-1811: reset island flags; accelerable bodies get their proxies synchronised and their contacts' TOIs reset
    // (body after body, each one's contact list over the lanes: every entry is a chain of three returning atomics)
    for (unsigned k = 0; k < isl.nB; ++k) {
        const unsigned b = d.toi_islBodies[isl.base + k];
        const unsigned f = d.b_flags[b];
        if (PR_LANE == 0u) {
            if (f_speedable(f)) d.b_toiIsl[b] = 0u;
            if (f_accelerable(f)) d.sync_list[atomicAdd(&d.cnt->syncN, 1u)] = b;
        }
        if (f_accelerable(f)) {
            for (unsigned q = d.adj_off[b] + PR_LANE, qe = adj_end(d, b); q < qe; q += PR_WARP) {
                const unsigned c2 = d.adj_list[q];
                const unsigned f2 = atomicAnd(&d.c_flags[c2], ~static_cast<unsigned>(PR_CONTACT_TOI));
                if (toi_static_eligible(f2)) toi_mark_recompute(d, c2);
                reinterpret_cast<float*>(&d.c_mat[c2])[3] = 1.0f;
                d.c_toiIsl[c2] = 0u;
            }
        }
    }
    PR_SYNCWARP();
    PR_TT(6);
    if (PR_LANE != 0u) return;
#ifdef PR_TOI_TIMING
    if (e == 0u) { atomicAdd(&d.cnt->dbg[8], 1ull); atomicAdd(&d.cnt->dbg[9], isl.nC); atomicAdd(&d.cnt->dbg[10], isl.nB);
                   atomicAdd(&d.cnt->dbg[11], positionIters); atomicAdd(&d.cnt->dbg[12], velocityIters);
                   atomicAdd(&d.cnt->dbg[13], numUpdated); atomicAdd(&d.cnt->dbg[14], numSkipped); }
#endif
    atomicMin(&d.cnt->toiMinSepFlip, float_flip(minSepAll));
    atomicMax(&d.cnt->toiMaxIncBits, f2u(maxIncAll));
    atomicAdd(&d.cnt->toiIslandsSolved, solved);
    atomicAdd(&d.cnt->toiSumPosIters, positionIters);
    atomicAdd(&d.cnt->toiSumVelIters, velocityIters);
    atomicAdd(&d.cnt->toiIslandsFound, (positionIters > 0u || velocityIters > 0u) ? 1u : 0u);
    atomicAdd(&d.cnt->toiUpdTouching, numUpdated);
    atomicAdd(&d.cnt->toiSkipTouching, numSkipped);
}

PR_KERNEL_LB(PR_BLOCK, 2) k_toi_solve(PR_GRID_CONST DV d, PR_GRID_CONST pr_step_conf conf, unsigned nEvents, unsigned exactMode, unsigned scratchCap)
{
    PR_PDL_ENTER();
    // one WARP per event (toi_solve_event deals the event's bodies, contact-list entries and rows out to the lanes): events take
    // wildly different paths (island sizes, manifold cases, iteration counts), so packing 32 of them into one warp serialises
    // them (measured: 3.5 active lanes)
#ifndef PR_HOST_EMUL
    __shared__ ToiSmem sm[PR_BLOCK / 32];
    ToiSmem* mine = &sm[threadIdx.x >> 5];
#else
    ToiSmem local;
    ToiSmem* mine = &local;
#endif
    const unsigned e = PR_TID >> 5;
#ifdef PR_HOST_EMUL
    if ((PR_TID & 31u) != 0u) return; // (a warp of the emulation is one thread)
#endif
    if (e >= nEvents) return;
    toi_solve_event(d, conf, e, exactMode, scratchCap, mine);
}

// no event left (:1768-1777): ResetBodiesForSolveTOI, ResetContactsForSolveTOI
PR_KERNEL k_toi_finish_bodies(DV d, unsigned nBodies)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nBodies) { reinterpret_cast<float*>(&d.b_pos0[i])[3] = 0.0f; d.b_toiAlpha[i] = 0u; }
}
PR_KERNEL k_toi_finish_contacts(DV d)
{
    PR_PDL_ENTER(); // only contacts on the TOI list can carry a TOI or a TOI count
    const unsigned i = PR_TID;
    if (i < d.cnt->toiListN) {
        const unsigned id = d.toi_list[i];
        d.c_flags[id] = d.c_flags[id] & 0xFFu & ~static_cast<unsigned>(PR_CONTACT_TOI); // SetToi({}) + SetToiCount(0)
        d.c_rstamp[id] = 0u; // (round ids restart with every step)
        reinterpret_cast<float*>(&d.c_mat[id])[3] = 1.0f;
    }
}
// Synchronize (:2564-2585) for the bodies the round's events listed (:1800-1805): grid-stride over sync_list
PR_D void sync_body_proxies_one(const DV& d, const pr_step_conf& conf, unsigned body)
{
    const Xf xf1 = xf_from4(d.b_xf[body]);
    const Xf xf0 = get_xf(pos_from4(d.b_pos0[body]), d.b_lc[body]);
    const float2 dsp = conf.displace_multiplier * (xf1.p - xf0.p);
    for (unsigned k = d.bp_off[body]; k < d.bp_off[body + 1u]; ++k) {
        const unsigned px = d.bp_list[k];
        const Box a = compute_aabb(child_view(d, d.p_cg[px]), xf0, xf1);
        if (!box_contains(box_from4(d.p_fat[px]), a)) {
            d.p_fat[px] = box_to4(fatten_displace(a, conf.aabb_extension, dsp));
            if (d.p_moved[px] == 0u) { d.p_moved[px] = 1u; d.q_list[atomicAdd(&d.cnt->nQueue, 1u)] = px; }
            atomicAdd(&d.cnt->proxiesMoved, 1u);
        }
    }
}
PR_KERNEL k_sync_body_proxies(DV d, pr_step_conf conf)
{
    PR_PDL_ENTER();
    if (d.cnt->toiHalt) return;
    const unsigned n = d.cnt->syncN;
    for (unsigned i = PR_TID; i < n; i += PR_NTHREADS) sync_body_proxies_one(d, conf, d.sync_list[i]);
}

// ------------------------------------------------------------------ batched rounds (k_tb_*)
// A round used to cost four device->host reads (events? queued proxies? candidate pairs? errors?) and the launch
// bubbles behind them -- ~165 us of overhead around ~150 us of event solving, eight times a step. Here the host
// enqueues SEVERAL rounds back to back with fixed launch shapes; every size lives in the device counters, every
// kernel is a grid-stride loop over it, and cnt->toiHalt turns the rest of the batch into no-ops once a round finds
// no event (TB_DONE) or meets something only the host can do (grow a buffer, rebuild the tree or the hash:
// TB_HOST_FIND -- the round is then completed by the host's find_add_contacts -- or rebuild the adjacency: TB_ADJ).
// The contact-pool bookkeeping (m_contacts size, free stack, pool size) is mirrored in the counters meanwhile.
PR_KERNEL_LB(PR_BLOCK, 2) k_tb_solve(PR_GRID_CONST DV d, PR_GRID_CONST pr_step_conf conf, unsigned exactMode, unsigned scratchCap)
{
    PR_PDL_ENTER();
#ifndef PR_HOST_EMUL
    __shared__ ToiSmem sm[PR_BLOCK / 32];
    ToiSmem* mine = &sm[threadIdx.x >> 5];
#else
    ToiSmem local;
    ToiSmem* mine = &local;
#endif
    if (d.cnt->toiHalt) return;
    const unsigned nEvents = d.cnt->toiEvents;
    if (PR_TID == 0u) {
        if (nEvents == 0u) d.cnt->toiHalt = TB_DONE; // :1768 (threads that miss this flip see nEvents == 0 and do nothing either)
        else { d.cnt->tbRounds += 1u; d.cnt->tbFound += nEvents; d.cnt->lateWhy[6] = d.cnt->lateWhy[6] * 100000u + nEvents; d.cnt->lateWhy[7] += 1u; }
    }
#ifdef PR_HOST_EMUL
    if ((PR_TID & 31u) != 0u) return; // (a warp of the emulation is one thread)
#endif
    for (unsigned e = PR_TID >> 5; e < nEvents; e += PR_NTHREADS >> 5) { toi_solve_event(d, conf, e, exactMode, scratchCap, mine); PR_SYNCWARP(); } // one warp per event
}
// FindContacts' tree update for the proxies the round's events moved; decides whether the host must take over
PR_KERNEL k_tb_refit(DV d, unsigned nLeaves)
{
    PR_PDL_ENTER();
    if (d.cnt->toiHalt) return;
    const unsigned nQ = d.cnt->nQueue;
    const bool hostMust = nQ != 0u && 8ull * nQ >= nLeaves; // too many moved proxies for in-place refits: full rebuild
    if (PR_TID == 0u) {
        d.cnt->tbFindSkip = (nQ == 0u) ? 1u : 0u;
        d.cnt->tbAddN = 0u;
        if (nQ) { d.cnt->nPairs = 0u; d.cnt->error &= ~static_cast<unsigned>(ERR_PAIR_OVERFLOW); }
        if (hostMust) d.cnt->toiHalt = TB_HOST_FIND;
    }
    if (hostMust) return;
    for (unsigned t = PR_TID; t < nQ; t += PR_NTHREADS) bvh_refit_moved_one(d, t, nLeaves);
}
PR_KERNEL k_tb_query(DV d, unsigned nLeaves, const unsigned* sortedIdx, unsigned capPairs)
{
    PR_PDL_ENTER();
    if (d.cnt->toiHalt || d.cnt->tbFindSkip) return;
    const unsigned nQ = d.cnt->nQueue;
#ifndef PR_HOST_EMUL
    __shared__ unsigned stk[PR_BLOCK / QG][QSTACK];
    const unsigned lane = threadIdx.x & 31u, gl = lane & (QG - 1u);
    const unsigned gmask = ((1u << QG) - 1u) << (lane & ~(QG - 1u));
    for (unsigned qi = PR_TID / QG; qi < nQ; qi += PR_NTHREADS / QG) // (a group's lanes share qi: they loop together)
        bvh_query_group(d, d.q_list[qi], sortedIdx, capPairs, stk[threadIdx.x / QG], lane, gl, gmask);
#else
    for (unsigned t = PR_TID; t < nQ; t += PR_NTHREADS) bvh_query_serial(d, d.q_list[t], nLeaves, sortedIdx, capPairs);
#endif
}
// The rest of the round in ONE block (a TOI round adds at most a few hundred contacts; SMALL_SORT_MAX bounds it):
//   decide  : can the device add this round's pairs itself? (every thread derives the same answer from counters that no
//             thread changes before the decision is used); if so un-queue the proxies (end of FindContacts)
//   sort    : FindContacts' std::sort of the candidate keys (:746), bitonic network in shared memory
//   create  : AddContacts (:2423-2547) with ids from the mirrored pool bookkeeping
//   update  : AddContacts' Update(contactID) of the new contacts (:2531), then the bookkeeping of the round
// (four launches before; the block-wide barriers order the global writes of one stage before the reads of the next)
constexpr unsigned TB_ADD_BLOCK = 512u;
// returns false when the host must take over (toiHalt = TB_HOST_FIND); kbuf / vbuf: SMALL_SORT_MAX entries of shared memory
PR_D bool tb_add_block(const DV& d, const pr_step_conf& conf, unsigned capPairs, unsigned capContacts, unsigned capHash, unsigned hashTombs,
                       unsigned long long* kbuf, unsigned* vbuf)
{
    const unsigned nP = d.cnt->nPairs, order = d.cnt->devOrder, pool = d.cnt->devPool, nFree = d.cnt->devFree, nQ = d.cnt->nQueue;
    const bool hostMust = (d.cnt->error & ERR_PAIR_OVERFLOW) || nP > capPairs || nP > SMALL_SORT_MAX ||
                          order + nP + 1u > capContacts || pool + nP + 1u > capContacts ||
                          2ull * (static_cast<unsigned long long>(order) + nP + hashTombs) > capHash;
    __syncthreads(); // (every thread has read the counters the decision depends on)
    if (hostMust) { if (PR_TID == 0u) d.cnt->toiHalt = TB_HOST_FIND; return false; }
    for (unsigned t = PR_TID; t < nQ; t += PR_NTHREADS) d.p_moved[d.q_list[t]] = 0u;
    __syncthreads();
    if (PR_TID == 0u) { d.cnt->nQueue = 0u; d.cnt->tbPairs += nP; }
    if (nP == 0u) return true;
#ifndef PR_HOST_EMUL
    unsigned N = 32u;
    while (N < nP) N <<= 1;
    for (unsigned i = threadIdx.x; i < N; i += blockDim.x) { kbuf[i] = (i < nP) ? d.np_key[i] : ~0ull; vbuf[i] = i; }
    __syncthreads();
    for (unsigned size = 2; size <= N; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < N / 2u; t += blockDim.x) {
                const unsigned lo = 2u * t - (t & (stride - 1u));
                const unsigned hi = lo + stride;
                const bool asc = ((lo & size) == 0u);
                const unsigned long long a = kbuf[lo], b = kbuf[hi];
                if ((a > b) == asc) { kbuf[lo] = b; kbuf[hi] = a; const unsigned tv = vbuf[lo]; vbuf[lo] = vbuf[hi]; vbuf[hi] = tv; }
            }
            __syncthreads();
        }
    }
    for (unsigned i = threadIdx.x; i < nP; i += blockDim.x) { d.np_key2[i] = kbuf[i]; d.np_idx2[i] = vbuf[i]; }
    __threadfence_block();
    __syncthreads();
#else
    (void)kbuf; (void)vbuf;
    sort_pairs_small(d.np_key, d.np_key2, d.np_idx2, nP, 0);
#endif
    for (unsigned i = PR_TID; i < nP; i += PR_NTHREADS) create_contact_one(d, i, d.np_key2, d.np_idx2, order, nFree, pool, 1u);
    __threadfence_block();
    __syncthreads();
    for (unsigned i = PR_TID; i < nP; i += PR_NTHREADS) { const unsigned id = d.c_order[order + i]; update_contact(d, conf, id, d.c_flags[id]); }
    if (PR_TID == 0u) {
        const unsigned used = nP < nFree ? nP : nFree;
        d.cnt->devFree = nFree - used; d.cnt->devPool = pool + (nP - used); d.cnt->devOrder = order + nP; d.cnt->tbAdded += nP;
    }
    return true;
}
PR_KERNEL k_tb_add(DV d, pr_step_conf conf, unsigned capPairs, unsigned capContacts, unsigned capHash, unsigned hashTombs)
{
    PR_PDL_ENTER();
    if (d.cnt->toiHalt || d.cnt->tbFindSkip) return;
#ifndef PR_HOST_EMUL
    __shared__ unsigned long long k[SMALL_SORT_MAX];
    __shared__ unsigned v[SMALL_SORT_MAX];
    tb_add_block(d, conf, capPairs, capContacts, capHash, hashTombs, k, v);
#else
    tb_add_block(d, conf, capPairs, capContacts, capHash, hashTombs, nullptr, nullptr);
#endif
}

// ------------------------------------------------------------------ late rounds in ONE block (k_tb_late)
// After the first round of a step the TOI phase is a handful of events per round (at 1M bodies: <= 8, for the bodies the
// pile presses into the terrain, one event per sub-step up to maxSubSteps), yet every round cost eleven launches over the
// full candidate list besides the ~130 us event chain. Here ONE block runs whole rounds back to back, block barriers
// between the stages, over compact lists: the contacts the previous round's events invalidated (toi_mark_recompute) and the
// surviving candidates. It hands back to the list kernels (TB_BIG) when a list outgrows it or a sleeping body was woken,
// and to the host for what only the host can do -- exactly like the batched list rounds.
constexpr unsigned LATE_BLOCK = 512u, LATE_MAX_LIST = 2048u, LATE_MAX_CAND = 192u, LATE_SOLVE_WARPS = 8u;
PR_KERNEL_LB(LATE_BLOCK, 1) k_tb_late(PR_GRID_CONST DV d, PR_GRID_CONST pr_step_conf conf, unsigned exactMode, unsigned scratchCap, unsigned nLeaves, const unsigned* sortedIdx,
                    unsigned capPairs, unsigned capContacts, unsigned capHash, unsigned hashTombs, unsigned maxRounds)
{
    PR_PDL_ENTER();
#ifndef PR_HOST_EMUL
    // one shared-memory arena, re-used stage by stage: event islands | query stacks | sort buffers
    __shared__ __align__(16) unsigned char arena[LATE_SOLVE_WARPS * sizeof(ToiSmem)];
    static_assert(sizeof(arena) >= (LATE_BLOCK / QG) * QSTACK * sizeof(unsigned), "query stacks fit the arena");
    static_assert(sizeof(arena) >= SMALL_SORT_MAX * (sizeof(unsigned long long) + sizeof(unsigned)), "sort buffers fit the arena");
    ToiSmem* islands = reinterpret_cast<ToiSmem*>(arena);
#else
    ToiSmem local;
    ToiSmem* islands = &local;
#endif
#ifdef PR_TOI_TIMING
    long long lt_ = clock64();
#define PR_LT(k) do { if (PR_TID == 0u) { const long long t_ = clock64(); d.cnt->dbgLate[k] += static_cast<unsigned long long>(t_ - lt_); lt_ = t_; } } while (0)
#else
#define PR_LT(k) do { } while (0)
#endif
    for (unsigned r = 0; r < maxRounds; ++r) {
        __syncthreads();
        if (d.cnt->toiHalt) return;
        PR_LT(0);
        const unsigned roundId = d.cnt->toiRoundId + 1u;
        const unsigned nR = d.cnt->toiReN[roundId & 1u], nC = d.cnt->toiCandN;
        const unsigned* R = (roundId & 1u) ? d.toi_re1 : d.toi_re0;
        const unsigned stop = d.cnt->adjOverflow ? TB_ADJ : d.cnt->error ? TB_ERROR
                            : (d.cnt->toiForceFull || nR > LATE_MAX_LIST || nC + nR > LATE_MAX_CAND) ? TB_BIG : TB_RUN;
        __syncthreads();
        if (stop == TB_BIG) { // too much for one block (or a body was woken): the list kernels enqueued behind this launch go on
            if (PR_TID == 0u) { // diagnostics for PR_TOI_DEBUG: why, and how long the lists were
                if (d.cnt->toiForceFull) d.cnt->lateWhy[0] += 1u; else if (nR > LATE_MAX_LIST) d.cnt->lateWhy[1] += 1u; else d.cnt->lateWhy[2] += 1u;
                if (nC > d.cnt->lateWhy[3]) d.cnt->lateWhy[3] = nC;
                if (nR > d.cnt->lateWhy[4]) d.cnt->lateWhy[4] = nR;
            }
            return;
        }
        if (PR_TID == 0u && stop == TB_RUN) d.cnt->lateWhy[5] += 1u; // (diagnostics: rounds run by the block)
        if (stop != TB_RUN) { if (PR_TID == 0u) d.cnt->toiHalt = stop; return; }
        if (PR_TID == 0u) { toi_open_round(d, roundId); d.cnt->toiCand2N = 0u; d.cnt->toiAtMaxSub += d.cnt->toiAtMaxLive; }
        // ---- UpdateContactTOIs over the invalidated contacts
        for (unsigned i = PR_TID; i < nR; i += PR_NTHREADS) toi_alpha_one(d, R[i], conf.max_sub_steps);
        __threadfence_block(); __syncthreads();
        for (unsigned i = PR_TID; i < nR; i += PR_NTHREADS) toi_advance_one(d, R[i], conf.max_sub_steps, roundId);
        // (same stage) candidates that kept their TOI stay candidates; the invalidated ones are re-listed by their computation
        for (unsigned i = PR_TID; i < nC; i += PR_NTHREADS) {
            const unsigned id = d.toi_cand[i];
            const float toi = d.c_mat[id].w;
            if ((d.c_flags[id] & PR_CONTACT_TOI) && toi < 1.0f) {
                d.toi_cand2[atomicAdd(&d.cnt->toiCand2N, 1u)] = id;
                atomicMin(&d.cnt->toiMinKey, toi_key(toi, d.c_pos[id]));
            }
        }
        __threadfence_block(); __syncthreads();
        for (unsigned i = PR_TID; i < nR; i += PR_NTHREADS) {
            const unsigned id = R[i];
            const ToiComputeOut o = toi_compute_one(d, id, conf);
            if (o.updated) atomicAdd(&d.cnt->toiUpdated, 1u);
            atomicMax(&d.cnt->toiMaxDistIters, o.distIt); atomicMax(&d.cnt->toiMaxToiIters, o.toiIt); atomicMax(&d.cnt->toiMaxRootIters, o.rootIt);
            if (o.key != ~0ull) { d.toi_cand2[atomicAdd(&d.cnt->toiCand2N, 1u)] = id; atomicMin(&d.cnt->toiMinKey, o.key); }
        }
        __threadfence_block(); __syncthreads();
        PR_LT(1); // UpdateContactTOIs
        const unsigned nCand = d.cnt->toiCand2N;
        for (unsigned i = PR_TID; i < nCand; i += PR_NTHREADS) d.toi_cand[i] = d.toi_cand2[i];
        __syncthreads();
        if (PR_TID == 0u) d.cnt->toiCandN = nCand;
        __threadfence_block(); __syncthreads();
        // ---- GetSoonestContact
        if (!exactMode) {
            for (unsigned c = PR_WARP_ID; c < nCand; c += PR_NWARPS) toi_claim_cand(d, d.toi_cand[c]);
            __threadfence_block(); __syncthreads();
        }
        for (unsigned c = PR_WARP_ID; c < nCand; c += PR_NWARPS) toi_select_cand(d, d.toi_cand[c], exactMode ? 0u : 1u);
        __threadfence_block(); __syncthreads();
        if (!exactMode) {
            for (unsigned c = PR_WARP_ID; c < nCand; c += PR_NWARPS) toi_unclaim_cand(d, d.toi_cand[c]);
        }
        const unsigned nEvents = d.cnt->toiEvents;
        __threadfence_block(); __syncthreads();
        PR_LT(2); // selection
        if (nEvents == 0u) { if (PR_TID == 0u) d.cnt->toiHalt = TB_DONE; return; } // :1768
        if (PR_TID == 0u) { d.cnt->tbRounds += 1u; d.cnt->tbFound += nEvents; }
#ifdef PR_TOI_TIMING
        if (PR_TID == 0u) { d.cnt->dbgLate[8] += 1ull; d.cnt->dbgLate[9] += nEvents; d.cnt->dbgLate[10] += nR; d.cnt->dbgLate[11] += nCand; }
#endif
        // ---- SolveToi per event: one warp each, islands in the arena
        if (PR_WARP_ID < LATE_SOLVE_WARPS)
            for (unsigned e = PR_WARP_ID; e < nEvents; e += (PR_NWARPS < LATE_SOLVE_WARPS ? PR_NWARPS : LATE_SOLVE_WARPS)) {
                toi_solve_event(d, conf, e, exactMode, scratchCap, &islands[PR_WARP_ID < LATE_SOLVE_WARPS ? PR_WARP_ID : 0u]);
                PR_SYNCWARP();
            }
        __threadfence_block(); __syncthreads();
        PR_LT(3); // event solve
        // ---- Synchronize the moved bodies' proxies, FindContacts, AddContacts
        {
            const unsigned n = d.cnt->syncN;
            for (unsigned i = PR_TID; i < n; i += PR_NTHREADS) sync_body_proxies_one(d, conf, d.sync_list[i]);
        }
        __threadfence_block(); __syncthreads();
        const unsigned nQ = d.cnt->nQueue;
        PR_LT(4); // Synchronize
        if (nQ == 0u) continue;
        if (8ull * nQ >= nLeaves) { if (PR_TID == 0u) d.cnt->toiHalt = TB_HOST_FIND; return; } // full rebuild: the host's
        for (unsigned t = PR_TID; t < nQ; t += PR_NTHREADS) bvh_refit_moved_one(d, t, nLeaves);
        if (PR_TID == 0u) { d.cnt->nPairs = 0u; d.cnt->error &= ~static_cast<unsigned>(ERR_PAIR_OVERFLOW); }
        __threadfence_block(); __syncthreads();
#ifndef PR_HOST_EMUL
        {
            unsigned (*stk)[QSTACK] = reinterpret_cast<unsigned (*)[QSTACK]>(arena);
            const unsigned lane = threadIdx.x & 31u, gl = lane & (QG - 1u);
            const unsigned gmask = ((1u << QG) - 1u) << (lane & ~(QG - 1u));
            for (unsigned qi = threadIdx.x / QG; qi < nQ; qi += blockDim.x / QG)
                bvh_query_group(d, d.q_list[qi], sortedIdx, capPairs, stk[threadIdx.x / QG], lane, gl, gmask);
        }
        __threadfence_block(); __syncthreads();
        PR_LT(5); // refit + query
        if (!tb_add_block(d, conf, capPairs, capContacts, capHash, hashTombs, reinterpret_cast<unsigned long long*>(arena),
                          reinterpret_cast<unsigned*>(arena + SMALL_SORT_MAX * sizeof(unsigned long long)))) return;
        PR_LT(6); // add
#else
        for (unsigned t = PR_TID; t < nQ; t += PR_NTHREADS) bvh_query_serial(d, d.q_list[t], nLeaves, sortedIdx, capPairs);
        if (!tb_add_block(d, conf, capPairs, capContacts, capHash, hashTombs, nullptr, nullptr)) return;
#endif
        __threadfence_block();
    }
}

// standalone GetToiViaSat over arrays (pr_toi_batch): sweeps as 8 floats pos0 x,y,a | pos1 x,y,a | lc x,y
PR_KERNEL k_toi_batch(unsigned n, const pr_child* children, const float2* verts, const float2* normals, const unsigned* aChild,
                      const float* sweepA, const unsigned* bChild, const float* sweepB, pr_step_conf conf, float* outTime,
                      int* outState, unsigned* outIters)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) {
        const pr_child ca = children[aChild[i]], cb = children[bChild[i]];
        ProxyView A, B;
        A.radius = ca.vertex_radius; A.count = ca.vertex_count; A.v = verts + ca.vertex_offset; A.n = normals + ca.vertex_offset;
        B.radius = cb.vertex_radius; B.count = cb.vertex_count; B.v = verts + cb.vertex_offset; B.n = normals + cb.vertex_offset;
        const float* a = sweepA + 8u * i; const float* b = sweepB + 8u * i;
        SweepD sA, sB;
        sA.pos0.l = v2(a[0], a[1]); sA.pos0.a = a[2]; sA.pos1.l = v2(a[3], a[4]); sA.pos1.a = a[5]; sA.lc = v2(a[6], a[7]);
        sB.pos0.l = v2(b[0], b[1]); sB.pos0.a = b[2]; sB.pos1.l = v2(b[3], b[4]); sB.pos1.a = b[5]; sB.lc = v2(b[6], b[7]);
        const ToiResult r = toi_via_sat(A, sA, B, sB, conf.target_depth, conf.tolerance, conf.max_toi_iters, conf.max_toi_root_iters,
                                        conf.max_distance_iters);
        outTime[i] = r.time; outState[i] = r.state;
        outIters[3u * i] = r.toiIters; outIters[3u * i + 1u] = r.maxDistIters; outIters[3u * i + 2u] = r.maxRootIters;
    }
}

} // namespace pr
#endif
