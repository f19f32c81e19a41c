// pr_mutate.cuh -- kernels behind the structural mutators of the boundary: Destroy(BodyID) (AabbTreeWorld.cpp:1062),
// Destroy(ShapeID) (:1259), Detach / the SetBody whose value has fewer shapes, another type or another enabled state (:2723-2795),
// SetShape with new geometry or another filter (:1284-1356). None of this is on the step's hot path: a mutator is a handful of
// passes over the contact list and the proxy table, driven by pr_world.cu.
#ifndef PR_MUTATE_CUH
#define PR_MUTATE_CUH

namespace pr {

// Destroy(ContactID, from) (:2281-2288) for every contact that matches: mode 0 = contacts of body a, mode 1 = contacts of
// (body a, shape b) (IsFor, Contact.hpp), mode 2 = contacts whose proxy slot is flagged in dead[] (their leaf is gone).
// keep[i] = 0 for a match; InternalDestroy's effects (:2245-2279) happen here, the compaction of m_contacts and the pool's free
// stack (list order = the order the reference frees them in: m_bodyContacts[a] is m_contacts restricted to a) in k_destroy_compact.
PR_KERNEL k_destroy_mark(DV d, unsigned nOrder, unsigned* keep, unsigned mode, unsigned a, unsigned b, const unsigned* dead, unsigned hashLive)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const uint4 ab = d.c_ab[id];
        bool match;
        if (mode == 0u) match = ab.x == a || ab.y == a;
        else if (mode == 1u) { const uint4 ids = d.c_ids[id]; match = (ab.x == a && ids.x == b) || (ab.y == a && ids.z == b); }
        else { const uint2 px = d.c_px[id]; match = dead[px.x] != 0u || dead[px.y] != 0u; }
        keep[i] = match ? 0u : 1u;
        if (match) {
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
// the filter half of DestroyContacts (:2305-2328): a contact flagged for filtering is tested again -- EitherIsAccelerable,
// ShouldCollide(joints), ShouldCollide(shapes) -- and destroyed or un-flagged. keep[] comes from the overlap half (k_destroy_test).
PR_KERNEL k_destroy_filter(DV d, unsigned nOrder, unsigned* keep, unsigned hashLive)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned destroyed = 0u;
    if (i < nOrder && keep[i]) {
        const unsigned id = d.c_order[i];
        const unsigned flags = d.c_flags[id];
        if (flags & PR_CONTACT_FILTER) {
            const uint4 ab = d.c_ab[id], ids = d.c_ids[id];
            bool ok = f_accelerable(d.b_flags[ab.x]) || f_accelerable(d.b_flags[ab.y]);
            if (ok && d.nNcPairs) {
                const unsigned long long bk = (static_cast<unsigned long long>(ab.x < ab.y ? ab.x : ab.y) << 32) | (ab.x < ab.y ? ab.y : ab.x);
                unsigned lo = 0u, hi = d.nNcPairs;
                while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (d.nc_pairs[mid] < bk) lo = mid + 1u; else hi = mid; }
                if (lo < d.nNcPairs && d.nc_pairs[lo] == bk) ok = false;
            }
            if (ok) ok = filters_allow(d.sh_filter[ids.x], d.sh_group[ids.x], d.sh_filter[ids.z], d.sh_group[ids.z]);
            if (ok) d.c_flags[id] = flags & ~static_cast<unsigned>(PR_CONTACT_FILTER);
            else {
                keep[i] = 0u; destroyed = 1u;
                if ((d.ev_mask & PR_EVENT_END) && (flags & PR_CONTACT_TOUCHING)) emit_contact_event(d, PR_EVENT_END, id);
                if (d.m3[id].w > 0u && !(flags & PR_CONTACT_SENSOR)) { body_set_awake(d, ab.x); body_set_awake(d, ab.y); }
                d.c_flags[id] = PR_CONTACT_DESTROYED;
                if (hashLive) hash_remove(d, d.c_key[id]);
                d.c_key[id] = ~0ull;
                manifold_clear(d, id);
            }
        }
    }
    block_add(&d.cnt->nDestroyed, destroyed);
}
// SetShape with another filter (:1330-1345): the shape's contacts are flagged for filtering and their bodies woken
PR_KERNEL k_shape_flag_filter(DV d, unsigned nOrder, unsigned shape)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    unsigned any = 0u;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const uint4 ids = d.c_ids[id];
        if (ids.x == shape || ids.z == shape) {
            atomicOr(&d.c_flags[id], static_cast<unsigned>(PR_CONTACT_FILTER));
            const uint4 ab = d.c_ab[id];
            body_set_awake(d, ab.x); body_set_awake(d, ab.y);
            any = 1u;
        }
    }
    block_add(&d.cnt->nDestroyed, any); // (scratch use between steps: the host reads "any contact flagged?" from it)
}
// the proxy table without its destroyed slots: map[old] = new slot, or INVALID (stable; built on the host from its mirror)
PR_KERNEL k_px_compact(DV d, unsigned nOld, const unsigned* map, unsigned* tChild, unsigned* tCg, unsigned* tMoved, float4* tFat)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOld) {
        const unsigned k = map[i];
        if (k != INVALID) { tChild[k] = d.p_child[i]; tCg[k] = d.p_cg[i]; tMoved[k] = d.p_moved[i]; tFat[k] = d.p_fat[i]; }
    }
}
PR_KERNEL k_px_remap_contacts(DV d, unsigned nOrder, const unsigned* map)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const uint2 px = d.c_px[id];
        d.c_px[id] = make_uint2(map[px.x], map[px.y]);
    }
}
// m_proxiesForContacts for every proxy of a shape (ForMatchingProxies, :1341-1343)
PR_KERNEL k_queue_shape_proxies(DV d, unsigned nProxies, unsigned shape)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nProxies && d.p_shape[i] == shape && d.p_moved[i] == 0u) { d.p_moved[i] = 1u; d.q_list[atomicAdd(&d.cnt->nQueue, 1u)] = i; }
}

} // namespace pr
#endif
