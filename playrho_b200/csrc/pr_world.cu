// pr_world.cu -- host orchestration of the device-resident world and the extern "C" layer
// declared in include/playrho_b200.h. Step() mirrors Step(AabbTreeWorld&, const StepConf&)
// (Library/source/playrho/d2/AabbTreeWorld.cpp:2144-2219): pre-phase, SolveReg (:1472), SolveToi.
#include "pr_kernels.cuh"
#include "pr_solve_fused.cuh"
#include "pr_mutate.cuh"
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>
#include <unordered_map>

namespace pr {

// (launch count of the step running on this thread; the serial is process-wide and atomic: worlds stepped from several threads
// can only make each other re-read their counters, never skip a read)
static thread_local unsigned g_launches = 0;
static std::atomic<unsigned long long> g_launchSerial{0}; // never reset: "has any kernel been launched since ...?"
void count_launch() { ++g_launches; g_launchSerial.fetch_add(1ull, std::memory_order_relaxed); }
unsigned launches_since_reset() { return g_launches; }
void reset_launch_count() { g_launches = 0; }
#ifdef PR_HOST_EMUL
thread_local unsigned emul_tid = 0, emul_nthreads = 0;
#endif
static std::string g_createError;

// ------------------------------------------------------------------ memory / stream shims
#ifndef PR_HOST_EMUL
struct CudaError { std::string msg; };
#define CK(call)                                                                                   \
    do {                                                                                           \
        const cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) throw CudaError{std::string(#call) + ": " + cudaGetErrorString(e_)}; \
    } while (0)
// (zeroed, like the emulation build's calloc: growing a group copies its whole old capacity, and no kernel ever reads never-written memory)
// The fill runs on the default stream, which is not ordered with a world's (non-blocking) stream: it is waited for here, or a copy the
// caller enqueues into the new buffer could be overtaken by it.
static void* dev_alloc(size_t bytes) { void* p = nullptr; CK(cudaMalloc(&p, bytes ? bytes : 16)); CK(cudaMemset(p, 0, bytes ? bytes : 16)); CK(cudaStreamSynchronize(0)); return p; }
static void dev_free(void* p) { if (p) cudaFree(p); }
static void h2d(World& w, void* dst, const void* src, size_t n) { if (n) CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, w.stream)); }
static unsigned g_hostSyncs = 0; // (diagnostics, PR_SYNC_DEBUG: stream synchronisations since the last report)
static int g_syncTag = 0; static std::map<int, int> g_syncHist; // (which source lines they came from)
static void count_sync() { ++g_hostSyncs; ++g_syncHist[g_syncTag]; g_syncTag = 0; }
static void d2h(World& w, void* dst, const void* src, size_t n) { if (n) { CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, w.stream)); CK(cudaStreamSynchronize(w.stream)); count_sync(); } }
static void d2d(World& w, void* dst, const void* src, size_t n) { if (n) CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, w.stream)); }
#define CK_COPY_D2H_ASYNC(w, dst, src, n) CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, (w).stream))
static void dset(World& w, void* dst, int v, size_t n) { if (n) CK(cudaMemsetAsync(dst, v, n, w.stream)); }
static void sync(World& w) { CK(cudaStreamSynchronize(w.stream)); count_sync(); }
static void check_launch() { CK(cudaGetLastError()); }
struct StageTimer {
    static constexpr int MAXEV = 1024;
    cudaEvent_t ev[MAXEV]; int stage[MAXEV]; int n = 0; bool on = true;
    StageTimer() { for (auto& e : ev) cudaEventCreate(&e); }
    ~StageTimer() { for (auto& e : ev) cudaEventDestroy(e); }
    void mark(World& w, int s) { if (n < MAXEV - 1) { cudaEventRecord(ev[n], w.stream); stage[n++] = s; } }
    void finish(World& w, pr_step_profile& p)
    {
        mark(w, -1);
        cudaEventSynchronize(ev[n - 1]);
        for (int i = 0; i + 1 < n; ++i) {
            float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            if (stage[i] >= 0 && stage[i] < 16) p.ms_stage[stage[i]] += ms;
        }
        float tot = 0; cudaEventElapsedTime(&tot, ev[0], ev[n - 1]); p.ms_total = tot;
        n = 0;
    }
};
#else
struct CudaError { std::string msg; };
#define CK(call) (call)
static void* dev_alloc(size_t bytes) { return std::calloc(1, bytes ? bytes : 16); }
static void dev_free(void* p) { std::free(p); }
static void h2d(World&, void* dst, const void* src, size_t n) { if (n) std::memcpy(dst, src, n); }
static void d2h(World&, void* dst, const void* src, size_t n) { if (n) std::memcpy(dst, src, n); }
static void d2d(World&, void* dst, const void* src, size_t n) { if (n) std::memmove(dst, src, n); }
#define CK_COPY_D2H_ASYNC(w, dst, src, n) std::memcpy(dst, src, n)
static void dset(World&, void* dst, int v, size_t n) { if (n) std::memset(dst, v, n); }
static void sync(World&) {}
static void check_launch() {}
struct StageTimer { void mark(World&, int) {} void finish(World&, pr_step_profile&) {} };
#endif

// ------------------------------------------------------------------ array table (alloc / grow / clone)
enum Group { G_BODY, G_SHAPE, G_CHILD, G_VERT, G_PROXY, G_CONTACT, G_HASH, G_PAIR, G_CONS, G_ADJ, G_JOINT, G_COUNT };
struct ArrDesc { void** slot; size_t elem; Group g; size_t mult; };
static std::vector<ArrDesc> arrays(World& w)
{
    DV& d = w.d;
#define A(field, g) ArrDesc{reinterpret_cast<void**>(&d.field), sizeof(*d.field), g, 1}
#define A2(field, g, m) ArrDesc{reinterpret_cast<void**>(&d.field), sizeof(*d.field), g, m}
    return {
        A(b_xf, G_BODY), A(b_pos0, G_BODY), A(b_pos1, G_BODY), A(b_vel, G_BODY), A(b_acc, G_BODY), A(b_mass, G_BODY),
        A(b_lc, G_BODY), A(b_flags, G_BODY), A(s_vel, G_BODY), A(s_pos, G_BODY), A(b_island, G_BODY), A(b_root, G_BODY),
        A(b_moved, G_BODY), A(b_sync, G_BODY), A(isl_state, G_BODY), A(isl_velIters, G_BODY), A(isl_posIters, G_BODY),
        A(isl_inc, G_BODY), A(isl_maxInc, G_BODY), A(isl_sep, G_BODY), A(isl_minSep, G_BODY), A(isl_minUA, G_BODY),
        A(isl_nbodies, G_BODY), A(isl_nstat, G_BODY), A(isl_active, G_BODY), A(b_claim, G_BODY), A(b_cmask, G_BODY),
        A(adj_off, G_BODY), A(adj_cnt, G_BODY), A(adj_n, G_BODY), A(b_toiIsl, G_BODY), A(b_toiAlpha, G_BODY), A(b_toiStamp, G_BODY), A(b_inR, G_BODY),
        A(sync_list, G_BODY), A(blk_body, G_BODY), A(b_slot, G_BODY), A(blk_joint, G_JOINT), A(bp_off, G_BODY), A(bp_list, G_PROXY), A(bv_leafOf, G_PROXY), A(b_world, G_BODY), A(bv_wr, G_PROXY),
        A(c_pos, G_CONTACT), A(c_color, G_CONTACT), A(c_toiIsl, G_CONTACT), A(toi_events, G_CONTACT), A(toi_list, G_CONTACT), A(toi_cand, G_CONTACT), A(toi_cand2, G_CONTACT), A(toi_re0, G_CONTACT),
        A(toi_re1, G_CONTACT), A(c_rstamp, G_CONTACT),
        A(adj_list, G_ADJ), A(toi_islBodies, G_ADJ), A(toi_islContacts, G_ADJ),
        A(jt, G_JOINT), A(jt_work, G_JOINT), A(jt_gear, G_JOINT), A(jt_order, G_JOINT), A(nc_pairs, G_JOINT), A(isl_jflag, G_BODY),
        A(sh_mat, G_SHAPE), A(sh_filter, G_SHAPE), A(sh_group, G_SHAPE),
        A(ch_radius, G_CHILD), A(ch_voff, G_CHILD), A(ch_count, G_CHILD), A(ch_pack, G_CHILD), A(vtx, G_VERT), A(nrm, G_VERT),
        A(p_body, G_PROXY), A(p_shape, G_PROXY), A(p_child, G_PROXY), A(p_cg, G_PROXY), A(p_tree, G_PROXY), A(p_moved, G_PROXY),
        A(p_fat, G_PROXY), A(q_list, G_PROXY), A(bv_code, G_PROXY), A(bv_code2, G_PROXY), A(bv_idx, G_PROXY), A(bv_idx2, G_PROXY),
        A(bv_left, G_PROXY), A(bv_right, G_PROXY), A(bv_parent, G_PROXY), A(bv_lparent, G_PROXY), A(bv_visit, G_PROXY), A(bv_cbox, G_PROXY),
        A(c_key, G_CONTACT), A(c_ab, G_CONTACT), A(c_px, G_CONTACT), A(c_ids, G_CONTACT), A(c_mat, G_CONTACT), A(c_flags, G_CONTACT),
        A(m0, G_CONTACT), A(m1, G_CONTACT), A(m2, G_CONTACT), A(m3, G_CONTACT), A(c_order, G_CONTACT), A(c_free, G_CONTACT),
        A(c_tmp, G_CONTACT), A(c_tmp2, G_CONTACT),
        A(ht_keys, G_HASH),
        A(np_key, G_PAIR), A(np_key2, G_PAIR), A(np_px, G_PAIR), A(np_idx, G_PAIR), A(np_idx2, G_PAIR),
        A(vc0, G_CONS), A(vc1, G_CONS), A(vc2, G_CONS), A(vc3, G_CONS), A(vc4, G_CONS), A(vc5, G_CONS), A(vc6, G_CONS), A(vc7, G_CONS),
        A(vc8, G_CONS), A(vc_island, G_CONS), A(vc_flow, G_CONS), A(pc0, G_CONS), A(pc1, G_CONS), A(pc2, G_CONS), A(pc3, G_CONS), A(k_sorted, G_CONS),
        // per-contact scratch of the constraint-list build (sized like m_contacts: written before the constraint count is known)
        A(k_contact, G_CONTACT), A(k_color, G_CONTACT), A(k_tmpKey, G_CONTACT), A(k_tmpKey2, G_CONTACT), A(k_tmpVal, G_CONTACT), A(k_ulist, G_CONTACT),
    };
#undef A
#undef A2
}
static unsigned* cap_of(World& w, Group g)
{
    switch (g) {
    case G_BODY: return &w.capBodies; case G_SHAPE: return &w.capShapes; case G_CHILD: return &w.capChildren;
    case G_VERT: return &w.capVerts; case G_PROXY: return &w.capProxies; case G_CONTACT: return &w.capContacts;
    case G_HASH: return &w.capHash; case G_PAIR: return &w.capPairs; case G_CONS: return &w.capCons; case G_ADJ: return &w.capAdj; case G_JOINT: return &w.capJoints;
    default: return nullptr;
    }
}
static unsigned next_pow2(unsigned v) { unsigned p = 1; while (p < v) p <<= 1; return p; }

static void ensure_scratch(World& w, unsigned nMax)
{
    const size_t need = scan_scratch_elems(std::max(nMax, 256u * rs_blocks(nMax))) + 16;
    if (need > w.scanScratchCap) {
        dev_free(w.scanScratch);
        w.scanScratch = static_cast<unsigned*>(dev_alloc(need * sizeof(unsigned)));
        w.scanScratchCap = need; w.rs.scan = w.scanScratch; w.rs.scanCap = need;
    }
    const size_t hneed = 256u * static_cast<size_t>(rs_blocks(nMax)) + 16;
    if (hneed > w.rs.histCap) {
        dev_free(w.rs.hist);
        w.rs.hist = static_cast<unsigned*>(dev_alloc(hneed * sizeof(unsigned)));
        w.rs.histCap = hneed;
    }
}

// grow group g to hold at least `need` elements, preserving contents
static void ensure(World& w, Group g, unsigned need)
{
    unsigned* cap = cap_of(w, g);
    if (need <= *cap) return;
    unsigned newCap = std::max(need + need / 8u, std::max(1024u, 2u * *cap)); // (growth is expensive: double, with headroom)
    ++w.profile.grow_events;
    if (g == G_HASH) newCap = next_pow2(need);
    for (const ArrDesc& a : arrays(w)) {
        if (a.g != g) continue;
        void* np = dev_alloc(a.elem * newCap);
        if (*a.slot && *cap && g != G_HASH) d2d(w, np, *a.slot, a.elem * *cap);
        sync(w);
        dev_free(*a.slot);
        *a.slot = np;
    }
    if (g == G_HASH) { w.d.ht_mask = newCap - 1u; w.hashDirty = true; }
    *cap = newCap;
    ensure_scratch(w, std::max(std::max(std::max(w.capContacts, w.capProxies), std::max(w.capPairs, w.capCons)), std::max(w.capBodies, w.capAdj)));
}

#ifndef PR_HOST_EMUL
// counters travel through pinned host memory: one D2H + stream sync per read, and a ring of pinned slots for
// the asynchronous H2D writes (a slot is reused only after >= one intervening read, i.e. a stream sync)
static void* stage_buf(World& w, int which, size_t bytes)
{
    if (bytes > w.stageCap[which]) {
        sync(w);
        dev_free(w.stageBuf[which]);
        w.stageBuf[which] = dev_alloc(bytes + bytes / 4);
        w.stageCap[which] = bytes + bytes / 4;
    }
    return w.stageBuf[which];
}
static void read_counters(World& w)
{
    // only kernels change the device counters behind the host's back: if none was launched since the last read (or since the
    // write that followed it) the host copy IS the device's, and the round trip (a stream drain, ~20 us of idle GPU) is skipped
    if (w.cntSerialValid && w.cntSerial == g_launchSerial.load(std::memory_order_relaxed)) return;
    CK(cudaMemcpyAsync(&w.pinnedCnt[0], w.d.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, w.stream));
    CK(cudaStreamSynchronize(w.stream));
    count_sync();
    w.hostCnt = w.pinnedCnt[0];
    w.pinnedNext = 1; // everything queued before this point has executed: all write slots are free again
    w.cntSerial = g_launchSerial.load(std::memory_order_relaxed); w.cntSerialValid = true;
}
static void write_counters(World& w)
{
    if (w.pinnedNext >= World::PINNED_SLOTS) { CK(cudaStreamSynchronize(w.stream)); w.pinnedNext = 1; }
    Counters* slot = &w.pinnedCnt[w.pinnedNext++];
    *slot = w.hostCnt;
    CK(cudaMemcpyAsync(w.d.cnt, slot, sizeof(Counters), cudaMemcpyHostToDevice, w.stream));
}
#else
static void* stage_buf(World& w, int which, size_t bytes)
{
    if (bytes > w.stageCap[which]) { dev_free(w.stageBuf[which]); w.stageBuf[which] = dev_alloc(bytes); w.stageCap[which] = bytes; }
    return w.stageBuf[which];
}
static void read_counters(World& w) { d2h(w, &w.hostCnt, w.d.cnt, sizeof(Counters)); }
static void write_counters(World& w) { h2d(w, w.d.cnt, &w.hostCnt, sizeof(Counters)); }
#endif

#ifndef PR_HOST_EMUL
#define read_counters(w) (g_syncTag = __LINE__, read_counters(w))
#define sync(w) (g_syncTag = __LINE__, sync(w))
#endif
// zero single counters on the device without a round trip: what "read, clear a field, write back" did with a stream drain (~20 us of idle
// GPU per drain). The host copy's fields are cleared too, so a host copy that mirrored the device still does.
static void zero_counter(World& w, unsigned Counters::*field)
{
    dset(w, &(w.d.cnt->*field), 0, sizeof(unsigned));
    w.hostCnt.*field = 0u;
}
static void throw_device_error(World& w)
{
    const unsigned e = w.hostCnt.error;
    if (!e) return;
    std::string m = "device error:";
    if (e & ERR_STACK) m += " bvh traversal stack overflow;";
    if (e & ERR_SENSOR) m += " sensor contacts (GJK TestOverlap) not supported yet;";
    if (e & ERR_HASH_FULL) m += " contact hash full;";
    if (e & ERR_TOI_SCRATCH) m += " one TOI island needs more scratch rows than the arena holds;";
    if (e & ERR_FLOW_STALL) m += " dataflow solver: a row waited for its bodies beyond the spin limit;";
    throw CudaError{m};
}

// ids of destroyed bodies / shapes (World::Pools)
static bool body_dead(const World& w, unsigned id) { return id < w.pools.bodyDead.size() && w.pools.bodyDead[id]; }
static bool shape_dead(const World& w, unsigned id) { return id < w.pools.shapeDead.size() && w.pools.shapeDead[id]; }
static unsigned body_flags(World& w, unsigned id) { unsigned f = 0; d2h(w, &f, w.d.b_flags + id, 4); return f; }

// ------------------------------------------------------------------ pre-phase pieces
static unsigned create_queued_proxies(World& w, const pr_step_conf& conf)
{ // Step: CreateProxies for m_fixturesForProxies, AabbTreeWorld.cpp:2158-2168
    if (w.fixturesForProxies.empty()) return 0;
    std::vector<unsigned> body, shape, child, cg, tree;
    for (const auto& bs : w.fixturesForProxies) {
        const pr_shape& s = w.shapes[bs.second];
        for (unsigned c = 0; c < s.child_count; ++c) {
            const unsigned k = w.nProxies + static_cast<unsigned>(body.size());
            body.push_back(bs.first); shape.push_back(bs.second); child.push_back(c); cg.push_back(s.child_offset + c);
            if (w.bodyProxies.size() <= bs.first) w.bodyProxies.resize(bs.first + 1u);
            w.bodyProxies[bs.first].push_back(k);
            // DynamicTree leaf id: the last destroyed leaf's, else the next fresh node (World::Pools)
            unsigned leaf;
            if (!w.pools.treeFree.empty()) { leaf = w.pools.treeFree.back(); w.pools.treeFree.pop_back(); }
            else { leaf = w.pools.treeNextFresh; w.pools.treeNextFresh += (k == 0 ? 1u : 2u); }
            tree.push_back(leaf);
            w.pools.pxBody.push_back(bs.first); w.pools.pxShape.push_back(bs.second); w.pools.pxTree.push_back(leaf);
            w.maxTreeId = std::max(w.maxTreeId, tree.back());
        }
    }
    w.fixturesForProxies.clear();
    const unsigned n = static_cast<unsigned>(body.size());
    if (!n) return 0;
    ensure(w, G_PROXY, w.nProxies + n);
    const unsigned first = w.nProxies;
    h2d(w, w.d.p_body + first, body.data(), n * 4); h2d(w, w.d.p_shape + first, shape.data(), n * 4);
    h2d(w, w.d.p_child + first, child.data(), n * 4); h2d(w, w.d.p_cg + first, cg.data(), n * 4);
    h2d(w, w.d.p_tree + first, tree.data(), n * 4);
    sync(w); // host vectors go out of scope
    PR_LAUNCH(k_create_proxy_aabbs, n, w.stream, w.d, conf, first, n);
    read_counters(w);
    w.hostCnt.nQueue += n;
    write_counters(w);
    w.nProxies += n;
    w.bpDirty = true; w.bvhValid = false;
    return n;
}

// body -> proxy CSR (m_bodyProxies) for the TOI rounds' per-body Synchronize
static void upload_body_proxies(World& w)
{
    if (!w.bpDirty) return;
    const unsigned N = w.nBodies;
    ensure(w, G_BODY, N + 1);
    std::vector<unsigned> off(N + 1, 0u), list;
    list.reserve(w.nProxies);
    for (unsigned b = 0; b < N; ++b) {
        off[b] = static_cast<unsigned>(list.size());
        if (b < w.bodyProxies.size()) list.insert(list.end(), w.bodyProxies[b].begin(), w.bodyProxies[b].end());
    }
    off[N] = static_cast<unsigned>(list.size());
    h2d(w, w.d.bp_off, off.data(), 4 * (N + 1)); h2d(w, w.d.bp_list, list.data(), 4 * list.size());
    sync(w);
    w.bpDirty = false;
}

// m_contacts without the contacts whose keep flag (c_tmp) is 0: stable; their ids go onto the pool's free stack in list order
static unsigned compact_destroyed_contacts(World& w)
{
    DV& d = w.d;
    scan_exclusive_u32(d.c_tmp, d.k_tmpVal, w.nOrder, &d.cnt->nNew, w.scanScratch, w.stream);
    read_counters(w);
    const unsigned kept = w.hostCnt.nNew;
    const unsigned destroyed = w.nOrder - kept;
    if (destroyed) {
        PR_LAUNCH(k_destroy_compact, w.nOrder, w.stream, d, w.nOrder, d.c_tmp, d.k_tmpVal, d.c_tmp2, w.nFree);
        std::swap(d.c_order, d.c_tmp2);
        w.nFree += destroyed;
        w.nOrder = kept;
        w.hashTombs += destroyed; w.adjValid = false;
        if (2ull * (static_cast<unsigned long long>(w.nOrder) + w.hashTombs) > w.capHash) w.hashDirty = true;
    }
    return destroyed;
}
// Destroy(ContactID, from) for every contact that matches (k_destroy_mark, pr_mutate.cuh)
static unsigned destroy_contacts_where(World& w, unsigned mode, unsigned a, unsigned b, const unsigned* dead = nullptr)
{
    if (w.nOrder == 0) return 0;
    ensure(w, G_CONTACT, w.nOrder + 1);
    PR_LAUNCH(k_destroy_mark, w.nOrder, w.stream, w.d, w.nOrder, w.d.c_tmp, mode, a, b, dead, w.hashDirty ? 0u : 1u);
    return compact_destroyed_contacts(w);
}
static unsigned destroy_contacts(World& w)
{ // DestroyContacts, AabbTreeWorld.cpp:2290-2331: the overlap test and, when some contact is flagged for it, the filter test
    if (w.nOrder == 0) { w.needsFiltering = false; return 0; }
    DV& d = w.d;
    ensure(w, G_CONTACT, w.nOrder + 1);
    PR_LAUNCH(k_destroy_test, w.nOrder, w.stream, d, w.nOrder, d.c_tmp, w.hashDirty ? 0u : 1u);
    unsigned destroyed = compact_destroyed_contacts(w);
    // the filter half is a SECOND pass over what the overlap half left (:2305-2328): the contacts it destroys are freed after all of the
    // first pass's, whatever their places in the list -- the next contacts created get their ids in that order
    if (w.needsFiltering) {
        w.needsFiltering = false;
        if (w.nOrder) {
            PR_LAUNCH(k_destroy_mark, w.nOrder, w.stream, d, w.nOrder, d.c_tmp, 0u, INVALID, 0u, static_cast<const unsigned*>(nullptr), w.hashDirty ? 0u : 1u); // (keep = 1 everywhere: no body has this id)
            PR_LAUNCH(k_destroy_filter, w.nOrder, w.stream, d, w.nOrder, d.c_tmp, w.hashDirty ? 0u : 1u);
            destroyed += compact_destroyed_contacts(w);
        }
    }
    return destroyed;
}

static void rebuild_hash(World& w)
{
    ensure(w, G_HASH, next_pow2(std::max(2048u, 4u * w.capContacts)));
    dset(w, w.d.ht_keys, 0xFF, sizeof(unsigned long long) * w.capHash);
    PR_LAUNCH(k_hash_fill, w.nOrder, w.stream, w.d, w.nOrder);
    w.hashDirty = false; w.hashTombs = 0;
}

static int key_bits(unsigned maxVal) { int b = 1; while (b < 32 && (maxVal >> b)) ++b; return ((b + 7) / 8) * 8; }

// full LBVH rebuild over the current fat boxes (hostCnt mirrors the device counters on entry)
static void rebuild_bvh(World& w)
{
    DV& d = w.d;
    const unsigned P = w.nProxies;
    w.hostCnt.boundsMinX = w.hostCnt.boundsMinY = 0xFFFFFFFFu; w.hostCnt.boundsMaxX = w.hostCnt.boundsMaxY = 0u;
    write_counters(w);
    PR_LAUNCH(k_bvh_bounds, P, w.stream, d, P);
    PR_LAUNCH(k_bvh_morton, P, w.stream, d, P);
    const bool idxInB = radix_sort_pairs<unsigned>(d.bv_code, d.bv_idx, d.bv_code2, d.bv_idx2, P, 0, 32, true, w.rs, w.stream);
    const unsigned* codes = idxInB ? d.bv_code2 : d.bv_code;
    w.bvhSortedIdx = idxInB ? d.bv_idx2 : d.bv_idx;
    PR_LAUNCH(k_bvh_build, P - 1, w.stream, d, P, codes);
    PR_LAUNCH(k_bvh_refit, P, w.stream, d, P, w.bvhSortedIdx);
    w.bvhValid = true; w.bvhLeaves = P; w.bvhRefits = 0;
}

static unsigned bvh_refit_steps()
{
    static const unsigned v = [] { const char* e = getenv("PR_BVH_REFIT_STEPS"); return e ? static_cast<unsigned>(atoi(e)) : 3u; }();
    return v;
}
// FindContacts + AddContacts, AabbTreeWorld.cpp:709-751 and :2423-2547
static unsigned find_add_contacts(World& w, const pr_step_conf& conf, StageTimer& tm, bool toiRound = false, bool queueKnownEmpty = false)
{
    DV& d = w.d;
    if (queueKnownEmpty) return 0; // (the caller read nQueue == 0 and nothing has queued a proxy since: no drain to find that out again)
    read_counters(w);
    throw_device_error(w);
    const unsigned nQueue = w.hostCnt.nQueue;
    if (nQueue == 0) return 0;
    const unsigned P = w.nProxies;
    tm.mark(w, PR_STAGE_BVH);
    // few moved proxies and an intact tree: grow the ancestors' boxes in place instead of rebuilding
    const bool refit = w.bvhValid && w.bvhLeaves == P && P > 1 && 8ull * nQueue < P && w.bvhRefits < 256u;
    // many moved proxies, but bodies move a fraction of their size per step: the Morton order of the last rebuild is still a good
    // tree for a few steps -- recompute every box bottom-up over the SAME topology (k_bvh_refit alone: no bounds pass, no sort,
    // no build) and rebuild every bvh_refit_steps()-th time (PR_BVH_REFIT_STEPS, 0 = always rebuild; debug bit 10 too)
    const bool fullRefit = !refit && w.bvhValid && w.bvhLeaves == P && P > 1 && w.bvhFullRefits < bvh_refit_steps() && !(w.debugFlags & 1024u);
    if (refit) {
        PR_LAUNCH(k_bvh_refit_moved, nQueue, w.stream, d, nQueue, P);
        ++w.bvhRefits;
    } else if (fullRefit) {
        dset(w, d.bv_visit, 0, sizeof(unsigned) * P);
        PR_LAUNCH(k_bvh_refit, P, w.stream, d, P, w.bvhSortedIdx);
        ++w.bvhFullRefits;
    } else if (P > 1) { rebuild_bvh(w); w.bvhFullRefits = 0; }
    if (w.hashDirty) rebuild_hash(w);
    tm.mark(w, PR_STAGE_QUERY);
    unsigned nPairs = 0;
    for (;;) { // (hostCnt mirrors the device counters here: refreshed by the read above / at the loop's end)
        ensure(w, G_PAIR, std::max(1024u, nPairs));
        w.hostCnt.nPairs = 0; w.hostCnt.error &= ~static_cast<unsigned>(ERR_PAIR_OVERFLOW);
        write_counters(w);
        const bool morton = !refit && P > 1 && 4ull * nQueue >= P;
#ifndef PR_HOST_EMUL
        if (!morton) PR_LAUNCH(k_bvh_query_coop, static_cast<unsigned long long>(nQueue) * QG, w.stream, d, nQueue, P, w.bvhSortedIdx, w.capPairs);
        else
#endif
        PR_LAUNCH(k_bvh_query, morton ? P : nQueue, w.stream, d, nQueue, P, w.bvhSortedIdx, w.capPairs, morton ? 1u : 0u);
        read_counters(w);
        nPairs = w.hostCnt.nPairs;
        if (nPairs <= w.capPairs) break;
    }
    throw_device_error(w);
    PR_LAUNCH(k_clear_queue_flags, nQueue, w.stream, d, nQueue);
    w.hostCnt.nQueue = 0; w.hostCnt.error &= ~static_cast<unsigned>(ERR_PAIR_OVERFLOW);
    write_counters(w);
    w.profile.candidate_pairs += nPairs;
    if (nPairs == 0) return 0;
    tm.mark(w, PR_STAGE_ADD);
    // sort candidate keys (min tree id << 32 | max tree id) like FindContacts' std::sort (:746)
    const unsigned long long* keys; const unsigned* src;
    if (nPairs <= SMALL_SORT_MAX) { // one-block bitonic sort: the TOI rounds' and quiet steps' latency path
        sort_pairs_small(d.np_key, d.np_key2, d.np_idx2, nPairs, w.stream);
        keys = d.np_key2; src = d.np_idx2;
    } else { // LSD radix over the max field then the min field
        const unsigned maxTree = std::max(1u, w.maxTreeId);
        const int bits = key_bits(maxTree);
        const bool inB = radix_sort_pairs<unsigned long long>(d.np_key, d.np_idx, d.np_key2, d.np_idx2, nPairs, 0, bits, true, w.rs, w.stream);
        unsigned long long* kA = inB ? d.np_key2 : d.np_key; unsigned* vA = inB ? d.np_idx2 : d.np_idx;
        unsigned long long* kB = inB ? d.np_key : d.np_key2; unsigned* vB = inB ? d.np_idx : d.np_idx2;
        const bool inB2 = radix_sort_pairs<unsigned long long>(kA, vA, kB, vB, nPairs, 32, 32 + bits, false, w.rs, w.stream);
        keys = inB2 ? kB : kA; src = inB2 ? vB : vA;
    }
    const unsigned used = std::min(nPairs, w.nFree);
    const unsigned fresh = nPairs - used;
    ensure(w, G_CONTACT, std::max(w.poolSize + fresh, w.nOrder + nPairs) + 1);
    if (2ull * (static_cast<unsigned long long>(w.nOrder) + nPairs + w.hashTombs) > w.capHash) rebuild_hash(w);
    const bool adjLive = toiRound && w.adjValid;
    if (adjLive) ensure(w, G_ADJ, 2u * (w.nOrder + nPairs) + 2u * w.nBodies + 64u);
    PR_LAUNCH(k_create_contacts, nPairs, w.stream, d, nPairs, keys, src, w.nOrder, w.nFree, w.poolSize, adjLive ? 1u : 0u);
    const unsigned oldOrder = w.nOrder;
    w.nFree -= used; w.poolSize += fresh; w.nOrder += nPairs;
    if (adjLive) w.toiListBound += nPairs; else w.adjValid = false;
    // AddContacts ends with Update(contactID) for each new contact (:2531); not part of UpdateContacts' stats
    PR_LAUNCH(k_update_contacts, nPairs, w.stream, d, conf, oldOrder, nPairs, 0u, 0u);
    if (!toiRound) return nPairs; // (no drain: a device error surfaces at the next read of the counters, the adjacency is not live outside the TOI rounds)
    read_counters(w);
    throw_device_error(w);
    if (w.hostCnt.adjOverflow) { w.adjValid = false; w.hostCnt.adjOverflow = 0; write_counters(w); }
    return nPairs;
}

// ------------------------------------------------------------------ block solver plumbing (pr_block_solve.cuh)
static unsigned block_color_cap()
{ // colours a block walks with barriers; the rest of a world's rows (a hub's) is its serial tail. PR_BLK_COLORS overrides.
    static const unsigned cap = [] { const char* e = getenv("PR_BLK_COLORS"); const unsigned v = e ? static_cast<unsigned>(atoi(e)) : 16u; return std::min(64u, std::max(1u, v)); }();
    return cap;
}
static void ensure_block_tables(World& w)
{
    const unsigned need = w.d.nWorlds + 2u;
    if (need <= w.blkWorldCap) return;
    sync(w);
    dev_free(w.d.blk_off); dev_free(w.d.blk_rowOff); dev_free(w.d.blk_jOff);
    w.d.blk_off = static_cast<unsigned*>(dev_alloc(sizeof(unsigned) * need));
    w.d.blk_rowOff = static_cast<unsigned*>(dev_alloc(sizeof(unsigned) * need));
    w.d.blk_jOff = static_cast<unsigned*>(dev_alloc(sizeof(unsigned) * need));
    w.blkWorldCap = need; w.blocksDirty = true; w.blkJointsDirty = true;
}
// bodies grouped by world: counting sort on the device (world index = key, ids ascending inside a world)
static void rebuild_blocks(World& w)
{
    DV& d = w.d;
    const unsigned N = w.nBodies, W = d.nWorlds;
    ensure_block_tables(w);
    if (w.blocksDirty) {
        dset(w, d.blk_off, 0, sizeof(unsigned) * (W + 1u));
        PR_LAUNCH(k_blk_count, N, w.stream, d, N, d.blk_off);
        scan_exclusive_u32(d.blk_off, d.blk_off, W + 1u, nullptr, w.scanScratch, w.stream);
        if (N) {
            unsigned* buf = static_cast<unsigned*>(stage_buf(w, 0, sizeof(unsigned) * 4ull * N));
            unsigned *kA = buf, *vA = buf + N, *kB = buf + 2ull * N, *vB = buf + 3ull * N;
            d2d(w, kA, d.b_world, sizeof(unsigned) * N);
            const bool inB = radix_sort_pairs<unsigned>(kA, vA, kB, vB, N, 0, static_cast<int>(d.worldBits), true, w.rs, w.stream);
            PR_LAUNCH(k_blk_fill, N, w.stream, d, N, inB ? kB : kA, inB ? vB : vA);
        }
        read_counters(w);
        w.hostCnt.blkMaxBodies = 0u;
        write_counters(w);
        PR_LAUNCH(k_blk_max, W, w.stream, d, W);
        read_counters(w);
        w.blkMaxBodies = w.hostCnt.blkMaxBodies;
        w.blocksDirty = false;
    }
    if (w.blkJointsDirty) { // the joints of every world, ids ascending (few: built on the host)
        std::vector<unsigned> off(W + 1u, 0u), list;
        std::vector<std::pair<unsigned, unsigned>> wj;
        for (unsigned j = 0; j < w.joints.size(); ++j) if (w.joints[j].body_a != INVALID) wj.emplace_back(w.bodyWorld[w.joints[j].body_a], j);
        std::sort(wj.begin(), wj.end());
        for (const auto& e : wj) { ++off[e.first + 1u]; list.push_back(e.second); }
        for (unsigned k = 0; k < W; ++k) off[k + 1u] += off[k];
        ensure(w, G_JOINT, static_cast<unsigned>(list.size()) + 1u);
        h2d(w, d.blk_jOff, off.data(), sizeof(unsigned) * (W + 1u));
        h2d(w, d.blk_joint, list.data(), sizeof(unsigned) * list.size());
        sync(w);
        w.blkJointsDirty = false;
    }
}
static unsigned block_threads()
{
    static const unsigned t = [] { const char* e = getenv("PR_BLK_THREADS"); const unsigned v = e ? static_cast<unsigned>(atoi(e)) : BLK_THREADS; return std::min(BLK_MAX_THREADS, std::max(32u, v / 32u * 32u)); }();
    return t;
}
// The block solver wins while the worlds of one GPU are few enough to be latency-bound (B200, 514-body tumblers: 512 worlds 1.8
// vs 3.6 ms, 1 024 worlds 2.3 vs 4.9 ms for the whole solve); from ~2 500 worlds up the arena-wide per-colour launches, which
// run every world's hub rows side by side in full warps, have the higher throughput (8 192 worlds: 9.2 vs 15.7 ms).
// PR_BLK_MAX_WORLDS moves the crossover; debug bit 5 switches the block solver off.
static unsigned block_max_worlds()
{
    static const unsigned v = [] { const char* e = getenv("PR_BLK_MAX_WORLDS"); return e ? static_cast<unsigned>(atoi(e)) : 2560u; }();
    return v;
}
static bool use_block_solver(const World& w) { return w.d.nWorlds != 0u && w.d.nWorlds <= block_max_worlds() && !(w.debugFlags & (4u | 32u)); }

// ------------------------------------------------------------------ regular phase (SolveReg :1472-1561)
// the step's (island, non-speedable body) pair set (island_count_static): sized from what the last step inserted, emptied
static void prepare_island_statics(World& w, unsigned want)
{
    const unsigned cap = next_pow2(std::max(1024u, want));
    if (cap > w.isCap) {
        sync(w);
        dev_free(w.d.is_keys);
        w.d.is_keys = static_cast<unsigned long long*>(dev_alloc(sizeof(unsigned long long) * cap));
        w.isCap = cap;
    }
    w.d.is_mask = cap - 1u;
    dset(w, w.d.is_keys, 0xFF, sizeof(unsigned long long) * cap);
}
// islStatOverflow: the set was too small for this step (the pile just landed): count again from the rows with a larger one
static void recount_island_statics(World& w)
{
    for (unsigned want = 8u * (w.hostCnt.nStaticTouch + 256u);; want *= 4u) {
        prepare_island_statics(w, want);
        dset(w, w.d.isl_nstat, 0, sizeof(unsigned) * w.nBodies);
        w.hostCnt.islStatOverflow = 0u; w.hostCnt.maxIslandBodies = 0u;
        write_counters(w);
        PR_LAUNCH(k_isl_recount_statics, std::max(w.nCons, w.nJoints), w.stream, w.d, w.nCons, w.nJoints);
        PR_LAUNCH(k_isl_max_bodies, w.nBodies, w.stream, w.d, w.nBodies);
        read_counters(w);
        if (!w.hostCnt.islStatOverflow) return;
    }
}
static void solve_reg(World& w, const pr_step_conf& conf, pr_step_stats& stats, StageTimer& tm)
{
    DV& d = w.d;
    const unsigned N = w.nBodies;
    tm.mark(w, PR_STAGE_ISLANDS);
    bool blockMode = use_block_solver(w);
    if (blockMode) {
        if (w.blocksDirty || w.blkJointsDirty || w.blkWorldCap < d.nWorlds + 2u) rebuild_blocks(w);
        if (w.blkMaxBodies > BLK_MAX_BODIES) blockMode = false; // a world too large for one CTA's shared memory: arena-wide launches
    }
    // colours: a block walks few of them with barriers and runs the rest of its world's rows (a hub's) as its serial tail
    if (blockMode) { if (w.colorCap != block_color_cap()) { w.colorCap = block_color_cap(); w.recolorAll = true; } }
    else if (!(w.debugFlags & 16u) && w.colorCap != 64u) { w.colorCap = 64u; w.recolorAll = true; }
    prepare_island_statics(w, 4u * (w.isHint + 256u));
    PR_LAUNCH(k_isl_init, N, w.stream, d, N, conf.reg_velocity_iters, conf.reg_position_iters);
    PR_LAUNCH(k_isl_union, w.nOrder, w.stream, d, w.nOrder);
    PR_LAUNCH(k_isl_union_joints, w.nJoints, w.stream, d, w.nJoints);
    PR_LAUNCH(k_isl_flatten, N, w.stream, d, N);
    tm.mark(w, PR_STAGE_INTEGRATE);
    PR_LAUNCH(k_isl_prepare, N, w.stream, d, N, conf);
    tm.mark(w, PR_STAGE_COLOR);
    unsigned nCons = 0;
    if (w.nOrder) {
        ensure(w, G_CONTACT, w.nOrder + 1);
        PR_LAUNCH(k_mark_constraints, w.nOrder, w.stream, d, w.nOrder, d.c_tmp);
        scan_exclusive_u32(d.c_tmp, d.k_tmpVal, w.nOrder, &d.cnt->nConstraints, w.scanScratch, w.stream);
        PR_LAUNCH(k_gather_constraints, w.nOrder, w.stream, d, w.nOrder, d.c_tmp, d.k_tmpVal, w.recolorAll ? 1u : 0u);
        PR_LAUNCH(k_color_keep, w.nOrder, w.stream, d, w.nOrder); // nColorWork was zeroed by reset_counters at step start
        read_counters(w);
        nCons = w.hostCnt.nConstraints;
        ensure(w, G_CONS, nCons + 1); // the rows (13 planes of 16 B) are sized by the constraints, not by the contacts
    }
    w.nCons = nCons; w.nColors = 0;
    std::vector<std::pair<unsigned, unsigned>> ranges; // (start, count) per colour, overflow tail last
    unsigned tailStart = 0, tailCount = 0, nTailSegs = 0;
    if (nCons) {
        const unsigned nWork = w.hostCnt.nColorWork;
        w.recolorAll = false;
        unsigned nWorkLeft = nWork, firstRound = 0u;
        // the rounds the one-block tail could not finish (rare): launches over the list, the count looked at every 4th round
        auto finish_rounds = [&]() {
            const unsigned nW = nWorkLeft;
            for (unsigned round = firstRound;; ++round) {
                if (round && (round % 64u) == 0u) PR_LAUNCH(k_color_reset_claims, N, w.stream, d, N);
                PR_LAUNCH(k_color_claim, nW, w.stream, d, nW, round);
                PR_LAUNCH(k_color_commit, nW, w.stream, d, nW, round, w.colorCap);
                ++w.profile.color_rounds;
                if ((round & 3u) == 3u || nW < 64u) { // look at the remaining count every 4th round only
                    read_counters(w);
                    if (w.hostCnt.nUncolored == 0) break;
                }
                if (round > 100000u) throw CudaError{"colouring did not converge"};
            }
        };
        // what the one-block tail reported: 0 left; or it handed back at once (too much left: the launches go on over the whole list), or
        // its round budget ran out (the rest, compacted to the head of the list, goes on after a reset of the claims)
        auto after_block = [&]() {
            w.profile.color_rounds += w.hostCnt.colorRoundsDev;
            const bool handedBack = w.hostCnt.colorRoundsDev == COLOR_LAUNCH_ROUNDS;
            nWorkLeft = !w.hostCnt.nUncolored ? 0u : handedBack ? nWork : w.hostCnt.nUncolored;
            firstRound = handedBack ? COLOR_LAUNCH_ROUNDS : 64u;
        };
        bool blockPending = false; // the one-block tail is enqueued and its verdict not read yet
        if (nWork && nWork <= COLOR_BLOCK_MAX_WORK && !(w.debugFlags & 256u)) { // the first rounds as launches, the tail in one block (debug bit 8: launches only)
            for (unsigned round = 0; round < COLOR_LAUNCH_ROUNDS; ++round) {
                PR_LAUNCH(k_color_claim, nWork, w.stream, d, nWork, round);
                PR_LAUNCH(k_color_commit, nWork, w.stream, d, nWork, round, w.colorCap);
            }
            ensure(w, G_PAIR, nWork); // (np_idx: the block's second work list; the candidate-pair buffers are idle here)
#ifndef PR_HOST_EMUL
            PR_LAUNCH_GRID(k_color_rounds_block, 1, COLOR_BLOCK, w.stream, d, w.colorCap, d.np_idx, COLOR_LAUNCH_ROUNDS);
#else
            PR_LAUNCH(k_color_rounds_block, 1, w.stream, d, w.colorCap, d.np_idx, COLOR_LAUNCH_ROUNDS);
#endif
            // the block almost always finishes: outside the block solver the colour sort and the colour starts are enqueued behind it
            // and ONE read fetches both its verdict and the starts (no drain here); if something was left the sort is done again
            if (!blockMode) { blockPending = true; nWorkLeft = 0u; }
            else { read_counters(w); after_block(); }
        }
        if (nWorkLeft) finish_rounds();
        const unsigned* sortedColor = nullptr;
        if (blockMode) { // rows sorted by (world, colour); every world's first row: histogram + scan (pr_block_solve.cuh)
            dset(w, d.blk_rowOff, 0, sizeof(unsigned) * (d.nWorlds + 1u));
            PR_LAUNCH(k_blk_row_keys, nCons, w.stream, d, nCons, d.k_tmpKey, d.blk_rowOff);
            scan_exclusive_u32(d.blk_rowOff, d.blk_rowOff, d.nWorlds + 1u, nullptr, w.scanScratch, w.stream);
            const bool inB = radix_sort_pairs<unsigned>(d.k_tmpKey, d.k_sorted, d.k_tmpKey2, d.k_tmpVal, nCons, 0, 7 + static_cast<int>(d.worldBits), true, w.rs, w.stream);
            sortedColor = inB ? d.k_tmpKey2 : d.k_tmpKey;
            if (inB) d2d(w, d.k_sorted, d.k_tmpVal, nCons * sizeof(unsigned));
        } else {
        // stable sort by colour (one 8-bit radix pass): k_sorted = permutation of the base constraint list
        auto sort_and_starts = [&]() {
            d2d(w, d.k_tmpKey, d.k_color, nCons * sizeof(unsigned));
            const bool inB = radix_sort_pairs<unsigned>(d.k_tmpKey, d.k_sorted, d.k_tmpKey2, d.k_tmpVal, nCons, 0, 8, true, w.rs, w.stream);
            sortedColor = inB ? d.k_tmpKey2 : d.k_tmpKey;
            if (inB) d2d(w, d.k_sorted, d.k_tmpVal, nCons * sizeof(unsigned));
            dset(w, d.cnt->colorStart, 0xFF, sizeof(unsigned) * (MAX_COLORS + 1u)); // (INVALID everywhere, on the device: the host copy may be stale here)
            PR_LAUNCH(k_color_starts, nCons, w.stream, d, nCons, sortedColor);
            read_counters(w);
        };
        sort_and_starts();
        if (blockPending) {
            after_block();
            if (nWorkLeft) { finish_rounds(); sort_and_starts(); }
        }
        unsigned prevStart = nCons;
        for (int c = static_cast<int>(MAX_COLORS) - 1; c >= 0; --c) {
            const unsigned s = w.hostCnt.colorStart[c];
            if (s == INVALID) continue;
            if (static_cast<unsigned>(c) == COLOR_OVERFLOW) { tailStart = s; tailCount = prevStart - s; }
            else ranges.emplace_back(s, prevStart - s);
            prevStart = s;
        }
        std::reverse(ranges.begin(), ranges.end());
        if (tailCount) { // group the overflow tail by island: one thread per island segment solves it in list order
            // (k_tmpKey / k_tmpKey2 hold the sorted colours only up to here; c_tmp / c_tmp2 are free scratch)
            ensure(w, G_PAIR, tailCount);
            PR_LAUNCH(k_tail_keys, tailCount, w.stream, d, tailStart, tailCount, d.c_tmp, d.c_tmp2);
            const bool tb = radix_sort_pairs<unsigned>(d.c_tmp, d.c_tmp2, d.np_idx, d.np_idx2, tailCount, 0, key_bits(std::max(1u, N)), false, w.rs, w.stream);
            w.hostCnt.nTailSegs = 0; write_counters(w);
            PR_LAUNCH(k_tail_segments, tailCount, w.stream, d, tailStart, tailCount, tb ? d.np_idx : d.c_tmp, tb ? d.np_idx2 : d.c_tmp2);
            read_counters(w);
            nTailSegs = w.hostCnt.nTailSegs;
        }
        } // (!blockMode)
        w.blkSortedKey = sortedColor;
        if (w.debugFlags & 1u) { // solve order for the parity oracle (pr_get_solve_order)
            w.solveOrderIds.resize(nCons); w.solveOrderColors.resize(nCons);
            // c_tmp (marks, consumed) and c_tmp2 (stale ping-pong copy of c_order) are free scratch here
            PR_LAUNCH(k_solve_order, nCons, w.stream, d, nCons, d.c_tmp, d.c_tmp2, sortedColor);
            d2h(w, w.solveOrderIds.data(), d.c_tmp, nCons * sizeof(unsigned));
            d2h(w, w.solveOrderColors.data(), d.c_tmp2, nCons * sizeof(unsigned));
        }
        w.nColors = static_cast<unsigned>(ranges.size()) + (tailCount ? 1u : 0u);
        if (blockMode) w.nColors = 0u; // (filled from the device's count at the end of the step)
        else if (nWork == nCons) w.colorsAtFull = w.nColors; // (this step coloured everything from scratch)
        else if (w.nColors > 24u && w.nColors > w.colorsAtFull + 8u) w.recolorAll = true; // persistent colours fragmented
        if (blockMode) { }
        else if ((w.debugFlags & 16u) && w.colorCap == 64u && ranges.size() > 32u) { // hub rule (World::colorCap, opt-in): cap the colours from the next step on?
            unsigned long long beyond = tailCount;
            for (size_t c = COLOR_CAP_HUBS; c < ranges.size(); ++c) beyond += ranges[c].second;
            const unsigned islands = std::max(1u, w.hostCnt.islandsFound);
            if (4ull * beyond <= nCons && beyond / islands <= 256ull) { w.colorCap = COLOR_CAP_HUBS; w.recolorAll = true; }
        } else if (w.colorCap != 64u && nTailSegs && tailCount / nTailSegs > 512u) { // tails too long for one thread each: back
            w.colorCap = 64u; w.recolorAll = true;
        }
        tm.mark(w, PR_STAGE_SETUP);
        PR_LAUNCH(k_build_constraints, nCons, w.stream, d, nCons, conf, conf.do_warm_start ? conf.dt_ratio : 0.0f, blockMode ? 1u : 0u);
    }
    if (blockMode && !nCons) dset(w, d.blk_rowOff, 0, sizeof(unsigned) * (d.nWorlds + 1u)); // (no world has a row)
    const unsigned jSerial = w.jointsSerial ? 1u : 0u, jThreads = w.jointsSerial ? (w.nJoints ? 1u : 0u) : w.nJoints;
    // joints: one thread per joint when no two share a movable body, else level by level (refresh_joint_tables), else one thread for all
    auto launch_joints = [&](unsigned phase) {
        if (w.jointLevelOff.size() > 2u && !(w.debugFlags & 2048u)) { // (debug bit 11: one thread for all joints, the A/B of tools/bench_joints.py)
            for (size_t l = 0; l + 1u < w.jointLevelOff.size(); ++l) {
                const unsigned off = w.jointLevelOff[l], cnt = w.jointLevelOff[l + 1u] - off;
                if (cnt) PR_LAUNCH(k_joints_level, cnt, w.stream, d, conf, off, cnt, phase);
            }
        } else PR_LAUNCH(k_joints, jThreads, w.stream, d, conf, w.nJoints, phase, jSerial);
    };
    bool fused = false;
#ifndef PR_HOST_EMUL
    // The whole solve as ONE persistent cooperative kernel (pr_solve_fused.cuh): opt-in (debug bit 2). Measured on B200 it
    // only ties the per-colour launches on the 1M mixed scene (1.62-1.67 vs 1.61 ms for warm start + 8 + 3 iterations: a
    // grid barrier costs what a back-to-back launch costs, and a resident grid has fewer threads to hide the row -> body
    // load chain) and loses on the 8192-tumbler arena (65 colours + overflow tail: 16.5 vs 10.8 ms).
    if ((w.debugFlags & 4u) && ranges.size() <= MAX_COLORS) {
        if (!w.fusedProbed) {
            w.fusedProbed = true;
            int coop = 0, sms = 0, perSm = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, w.device);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w.device);
            if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_solve_fused, FUSED_BLOCK, 0) == cudaSuccess && perSm > 0) {
                w.fusedGrid = static_cast<unsigned>(sms * perSm);
                w.gridBar = static_cast<unsigned*>(dev_alloc(256));
            }
            cudaGetLastError();
        }
        if (w.fusedGrid) {
            SolvePlan p{};
            p.nColors = static_cast<unsigned>(ranges.size());
            for (unsigned c = 0; c < p.nColors; ++c) { p.start[c] = ranges[c].first; p.count[c] = ranges[c].second; }
            p.tailStart = tailStart; p.tailCount = tailCount; p.nTailSegs = nTailSegs;
            p.nBodies = N; p.nJoints = w.nJoints; p.jointsSerial = jSerial; p.warmStart = (nCons && conf.do_warm_start) ? 1u : 0u;
            tm.mark(w, PR_STAGE_VELOCITY);
            CK(cudaMemsetAsync(w.gridBar, 0, sizeof(unsigned), w.stream));
            pr_step_conf confCopy = conf;
            void* args[] = {&d, &confCopy, &p, &w.gridBar};
            const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_solve_fused), dim3(w.fusedGrid), dim3(FUSED_BLOCK), args, 0, w.stream);
            if (e == cudaSuccess) { fused = true; count_launch(); }
            else { cudaGetLastError(); w.fusedGrid = 0; } // fall back to one launch per colour from now on
        }
    }
#endif
    if (blockMode) { // one CTA per world runs the whole solve (pr_block_solve.cuh)
        tm.mark(w, PR_STAGE_VELOCITY);
        const unsigned maxB = std::max(1u, w.blkMaxBodies);
#ifndef PR_HOST_EMUL
        const unsigned smem = maxB * BLK_BYTES_PER_BODY;
        if (smem > w.blkSmemSet) { CK(cudaFuncSetAttribute(k_block_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); w.blkSmemSet = smem; }
        PR_LAUNCH_GRID_SMEM(k_block_solve, d.nWorlds, block_threads(), smem, w.stream, d, conf, w.blkSortedKey, maxB, (nCons && conf.do_warm_start) ? 1u : 0u);
#else
        PR_LAUNCH(k_block_solve, d.nWorlds, w.stream, d, conf, w.blkSortedKey, maxB, (nCons && conf.do_warm_start) ? 1u : 0u);
#endif
        tm.mark(w, PR_STAGE_POSITION);
    } else if (!fused && !(w.debugFlags & 128u) && 1u + conf.reg_velocity_iters + conf.reg_position_iters <= 32u) {
        // dataflow solver (k_flow_pass, pr_kernels.cuh): ONE launch per pass over all coloured rows; debug bit 7 = the per-colour launches below
        const unsigned nFlow = tailCount ? tailStart : nCons; // (the overflow tail sorts last)
        unsigned ticket = 0u, pass = 0u; // ticket: launches so far; pass: passes so far over the same array (s_vel, then s_pos)
        if (nFlow) PR_LAUNCH(k_flow_words, nFlow, w.stream, d, nFlow, w.blkSortedKey);
        if (nCons && conf.do_warm_start) {
            if (nFlow) { PR_LAUNCH(k_flow_pass<0u>, nFlow, w.stream, d, nFlow, ticket, pass, 0.0f, 0.0f, 0.0f); ++ticket; ++pass; } // (no ++ inside the macro's arguments)
            if (tailCount) PR_LAUNCH(k_warm_start_tail, nTailSegs, w.stream, d, tailStart, tailCount, nTailSegs);
        }
        launch_joints(0u); // InitVelocity after the contacts' warm start (:1597)
        tm.mark(w, PR_STAGE_VELOCITY);
        for (unsigned it = 0; it < conf.reg_velocity_iters; ++it) {
            launch_joints(1u); // joints first (:1605-1608)
            if (nFlow) { PR_LAUNCH(k_flow_pass<1u>, nFlow, w.stream, d, nFlow, ticket, pass, 0.0f, 0.0f, 0.0f); ++ticket; ++pass; }
            if (tailCount) PR_LAUNCH(k_solve_velocity_tail, nTailSegs, w.stream, d, tailStart, tailCount, nTailSegs);
            if (it == 0u) PR_LAUNCH(k_isl_vel_check, N, w.stream, d, N, it, conf.reg_min_momentum);
            else PR_LAUNCH(k_isl_vel_check_list, std::min(N, 64u * 256u), w.stream, d, it, conf.reg_min_momentum);
        }
        tm.mark(w, PR_STAGE_POSITION);
        PR_LAUNCH(k_integrate_positions, N, w.stream, d, N, conf.delta_time);
        pass = 0u;
        for (unsigned it = 0; it < conf.reg_position_iters; ++it) {
            if (nFlow) { PR_LAUNCH(k_flow_pass<2u>, nFlow, w.stream, d, nFlow, ticket, pass, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction); ++ticket; ++pass; }
            if (tailCount)
                PR_LAUNCH(k_solve_position_tail, nTailSegs, w.stream, d, tailStart, tailCount, nTailSegs, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction);
            launch_joints(2u); // joints after the contacts (:1634-1637)
            if (it == 0u) PR_LAUNCH(k_isl_pos_check, N, w.stream, d, N, it, conf.reg_min_separation);
            else PR_LAUNCH(k_isl_pos_check_list, std::min(N, 64u * 256u), w.stream, d, it, conf.reg_min_separation);
        }
    } else if (!fused) {
        // the trailing colours with at most SMALL_COLOR_ROWS rows go through ONE single-block launch per iteration
        size_t nBig = ranges.size();
#ifndef PR_HOST_EMUL
        SmallPlan sp{};
        while (nBig > 0 && ranges[nBig - 1].second <= SMALL_COLOR_ROWS && sp.n < SMALL_COLOR_MAX) { --nBig; ++sp.n; }
        if (sp.n < 2u) { nBig = ranges.size(); sp.n = 0u; } // (one small colour alone is just a launch)
        for (unsigned c = 0; c < sp.n; ++c) { sp.start[c] = ranges[nBig + c].first; sp.count[c] = ranges[nBig + c].second; }
#define PR_SMALL_COLORS(mode) do { if (sp.n) PR_LAUNCH_GRID(k_solve_small_colors, 1, SMALL_BLOCK, w.stream, d, sp, mode, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction); } while (0)
#else
#define PR_SMALL_COLORS(mode) do { } while (0)
#endif
        if (nCons && conf.do_warm_start) {
            for (size_t c = 0; c < nBig; ++c) PR_LAUNCH_PDL(k_warm_start, ranges[c].second, w.stream, d, ranges[c].first, ranges[c].second);
            PR_SMALL_COLORS(0u);
            if (tailCount) PR_LAUNCH(k_warm_start_tail, nTailSegs, w.stream, d, tailStart, tailCount, nTailSegs);
        }
        launch_joints(0u); // InitVelocity after the contacts' warm start (:1597)
        tm.mark(w, PR_STAGE_VELOCITY);
        for (unsigned it = 0; it < conf.reg_velocity_iters; ++it) {
            launch_joints(1u); // joints first (:1605-1608)
            for (size_t c = 0; c < nBig; ++c) PR_LAUNCH_PDL(k_solve_velocity, ranges[c].second, w.stream, d, ranges[c].first, ranges[c].second);
            PR_SMALL_COLORS(1u);
            if (tailCount) PR_LAUNCH(k_solve_velocity_tail, nTailSegs, w.stream, d, tailStart, tailCount, nTailSegs);
            if (it == 0u) PR_LAUNCH(k_isl_vel_check, N, w.stream, d, N, it, conf.reg_min_momentum);
            else PR_LAUNCH(k_isl_vel_check_list, std::min(N, 64u * 256u), w.stream, d, it, conf.reg_min_momentum);
        }
        tm.mark(w, PR_STAGE_POSITION);
        PR_LAUNCH(k_integrate_positions, N, w.stream, d, N, conf.delta_time);
        for (unsigned it = 0; it < conf.reg_position_iters; ++it) {
            for (size_t c = 0; c < nBig; ++c)
                PR_LAUNCH_PDL(k_solve_position, ranges[c].second, w.stream, d, ranges[c].first, ranges[c].second, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction);
            PR_SMALL_COLORS(2u);
            if (tailCount)
                PR_LAUNCH(k_solve_position_tail, nTailSegs, w.stream, d, tailStart, tailCount, nTailSegs, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction);
            launch_joints(2u); // joints after the contacts (:1634-1637)
            if (it == 0u) PR_LAUNCH(k_isl_pos_check, N, w.stream, d, N, it, conf.reg_min_separation);
            else PR_LAUNCH(k_isl_pos_check_list, std::min(N, 64u * 256u), w.stream, d, it, conf.reg_min_separation);
        }
#undef PR_SMALL_COLORS
    }
    tm.mark(w, PR_STAGE_WRITEBACK);
    if (nCons) PR_LAUNCH(k_assign_impulses, nCons, w.stream, d, nCons);
    PR_LAUNCH(k_writeback_bodies, N, w.stream, d, N, conf);
    PR_LAUNCH(k_sleep_and_stats, N, w.stream, d, N, conf.min_still_time_to_sleep);
    tm.mark(w, PR_STAGE_SYNC);
    zero_counter(w, &Counters::proxiesMoved); zero_counter(w, &Counters::contactsUpdated); // (no read: the solve's counters are read after the narrowphase)
    PR_LAUNCH(k_sync_proxies, w.nProxies, w.stream, d, conf, w.nProxies);
    tm.mark(w, PR_STAGE_NARROW);
    const unsigned nOrderAtUpdate = w.nOrder;
    // Binned variant: opt-in only (debug bit 1). Measured at 4.4M contacts it LOSES to the plain kernel (1.02 vs 0.90 ms):
    // homogeneous warps, but a warp's 32 contact ids are no longer consecutive, so every 16-byte row of the seven contact
    // planes drags in a 32-byte sector of its own and the kernel turns bandwidth-bound. Kept for the record and the tests.
    if ((w.debugFlags & 2u) && w.nOrder) {
        ensure(w, G_CONTACT, w.nOrder + 1); // (k_tmpKey/k_tmpKey2/k_tmpVal and c_tmp are free scratch at this point of the step)
        PR_LAUNCH(k_narrow_keys, w.nOrder, w.stream, d, w.nOrder, 1u, d.k_tmpKey, d.k_tmpVal);
        const bool inB = radix_sort_pairs<unsigned>(d.k_tmpKey, d.k_tmpVal, d.k_tmpKey2, d.c_tmp, w.nOrder, 0, 8, false, w.rs, w.stream);
        PR_LAUNCH(k_update_contacts_binned, w.nOrder, w.stream, d, conf, w.nOrder, inB ? d.k_tmpKey2 : d.k_tmpKey, inB ? d.c_tmp : d.k_tmpVal);
    }
#ifndef PR_HOST_EMUL
    else if ((w.debugFlags & 64u) && w.nOrder) { // TMA-staged vertex sets (k_update_contacts_tma): opt-in, measured slower
        if (!w.tmaSmemSet) { CK(cudaFuncSetAttribute(k_update_contacts_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(TMA_STAGE_BYTES))); w.tmaSmemSet = true; }
        PR_LAUNCH_GRID_SMEM(k_update_contacts_tma, (w.nOrder + PR_BLOCK - 1u) / PR_BLOCK, PR_BLOCK, TMA_STAGE_BYTES, w.stream, d, conf, 0u, w.nOrder, 1u, 1u);
    }
#endif
    else
        PR_LAUNCH(k_update_contacts, w.nOrder, w.stream, d, conf, 0u, w.nOrder, 1u, 1u);
    dset(w, d.b_moved, 0, sizeof(unsigned) * N);
    dset(w, d.b_sync, 0, sizeof(unsigned) * N);
    read_counters(w);
    if (w.hostCnt.islStatOverflow) recount_island_statics(w);
    w.isHint = w.hostCnt.nStaticTouch;
    const Counters c = w.hostCnt;
    stats.reg.proxies_moved = c.proxiesMoved;
    stats.reg.contacts_updated = c.contactsUpdated;
    stats.reg.contacts_skipped = nOrderAtUpdate - c.contactsUpdated;
    stats.reg.islands_found = c.islandsFound;
    stats.reg.islands_solved = c.islandsSolved;
    stats.reg.bodies_slept = c.bodiesSlept;
    stats.reg.sum_pos_iters = c.sumPosIters;
    stats.reg.sum_vel_iters = c.sumVelIters;
    stats.reg.max_island_bodies = c.maxIslandBodies;
    stats.reg.min_separation = float_unflip(c.minSepFlip);
    stats.reg.max_inc_impulse = u2f(c.maxIncBits);
    w.profile.n_awake = c.nAwake; w.profile.c_touching = c.nTouching; w.profile.c_dirty = c.contactsUpdated;
    if (blockMode) w.nColors = c.blkColors;
#ifdef PR_TOI_TIMING
    if (blockMode) std::fprintf(stderr, "blk_timing world 0: rows %llu tail %llu bodies %llu | cyc: warm colours %llu tail %llu | vel colours %llu tail %llu | pos colours %llu tail %llu | other %llu\n",
                                c.dbgBlk[8], c.dbgBlk[9], c.dbgBlk[10], c.dbgBlk[0], c.dbgBlk[1], c.dbgBlk[2], c.dbgBlk[3], c.dbgBlk[4], c.dbgBlk[5], c.dbgBlk[11]);
#endif
    w.profile.p_moved = c.proxiesMoved; w.profile.n_islands = c.islandsFound; w.profile.n_colors = w.nColors;
    w.profile.vel_iters = conf.reg_velocity_iters; w.profile.pos_iters = conf.reg_position_iters;
    stats.reg.contacts_added = find_add_contacts(w, conf, tm);
}

// ------------------------------------------------------------------ TOI phase (SolveToi :1752-1831)
// m_bodyContacts as CSR: count, scan, fill, then order each speedable body's segment by m_contacts position
static void build_adjacency(World& w)
{ // leaves hostCnt mirroring the device counters
    DV& d = w.d;
    const unsigned N = w.nBodies;
    ensure(w, G_BODY, N + 1);
    ensure(w, G_ADJ, 2u * w.nOrder + 2u * w.nOrder / 4u + 8u * N + 64u);
    dset(w, d.adj_cnt, 0, sizeof(unsigned) * N);
    dset(w, d.b_inR, 0, sizeof(unsigned) * N);
    zero_counter(w, &Counters::toiListN); zero_counter(w, &Counters::adjOverflow);
    PR_LAUNCH(k_toi_list_build, w.nOrder, w.stream, d, w.nOrder);
    if (w.toiListBound == 0u) { // (the last list was empty, or there was none: look before the three adjacency passes; a world that had TOI candidates
        read_counters(w);       //  last time has them now, almost surely -- no drain for the look)
        if (w.hostCnt.toiListN == 0u) { // no contact can ever get a TOI (e.g. an arena of tumblers: no static or bullet body touches anything):
            w.toiListBound = 0; w.adjValid = true; // the set R is empty, so the adjacency (three more passes over every contact) is too
            return;
        }
    }
    PR_LAUNCH(k_adj_count, w.nOrder, w.stream, d, w.nOrder);
    PR_LAUNCH(k_adj_cap, N, w.stream, d, N);
    scan_exclusive_u32(d.adj_cnt, d.adj_off, N, d.adj_off + N, w.scanScratch, w.stream);
    PR_LAUNCH(k_adj_fill, w.nOrder, w.stream, d, w.nOrder);
    read_counters(w);
    w.toiListBound = w.hostCnt.toiListN;
    w.adjValid = true;
}

static void solve_toi(World& w, const pr_step_conf& conf, pr_step_stats& stats, StageTimer& tm)
{
    DV& d = w.d;
    const unsigned N = w.nBodies;
    tm.mark(w, PR_STAGE_TOI);
    auto& st = stats.toi;
    PR_LAUNCH(k_toi_step_counters, 1, w.stream, d); // (no read: the host copy below is only kept alike; the next read refreshes it)
    w.hostCnt.toiAtMaxSub = w.hostCnt.toiUpdated = w.hostCnt.toiMaxDistIters = w.hostCnt.toiMaxToiIters = w.hostCnt.toiMaxRootIters = 0;
    w.hostCnt.toiMinSepFlip = float_flip(std::numeric_limits<float>::infinity());
    w.hostCnt.toiMaxIncBits = 0; w.hostCnt.toiIslandsSolved = w.hostCnt.toiSumPosIters = w.hostCnt.toiSumVelIters = 0;
    w.hostCnt.toiIslandsFound = w.hostCnt.toiUpdTouching = w.hostCnt.toiSkipTouching = 0;
    w.hostCnt.proxiesMoved = 0; w.hostCnt.contactsUpdated = 0; w.hostCnt.toiDeferred = 0;
    w.hostCnt.toiRoundId = 0; w.hostCnt.toiReN[0] = w.hostCnt.toiReN[1] = 0; w.hostCnt.toiAtMaxLive = 0; w.hostCnt.toiForceFull = 0; w.hostCnt.toiHalt = TB_RUN;
    const bool exact = (w.toiMode == 0u) || w.subStepping; // (sub-stepping = one TOI event per Step call, in the reference's order: :1779-1824)
    unsigned rounds = 0, round = 0, toiBatches = 0, listRoundsUsed = 0;
    if (!w.toiPossible) { // no contact of this world can be a TOI candidate: SolveToi finds nothing (:1768) and its final UpdateContacts skips all
        w.stepComplete = true;
        if (w.toiAlphaDirty) { PR_LAUNCH(k_toi_finish_bodies, N, w.stream, d, N); w.toiAlphaDirty = false; }
        st.contacts_skipped_touching = w.nOrder;
        return;
    }
    upload_body_proxies(w);
    if (w.debugFlags & 1u) { // record the event schedule (pr_get_toi_schedule)
        const unsigned want = std::max(1u << 16, 4u * w.capContacts / 8u);
        if (want > d.toi_logCap) { sync(w); dev_free(d.toi_log); d.toi_log = static_cast<uint2*>(dev_alloc(sizeof(uint2) * want)); d.toi_logCap = want; }
    } else if (d.toi_log) { sync(w); dev_free(d.toi_log); d.toi_log = nullptr; d.toi_logCap = 0u; }
    w.toiScheduleN = 0; w.toiSchedule.clear();
    PR_LAUNCH(k_toi_phase_init, N, w.stream, d, N);
    for (;;) {
        if (w.nOrder == 0) { w.stepComplete = true; break; } // nothing can have a TOI
        if (!w.adjValid) build_adjacency(w);
        // island scratch (body / contact lists, VC/PC rows of islands too large for shared memory) of ONE round's events: an
        // event that does not fit is deferred to the next round (k_toi_solve), so this is a working size, not a worst case
        ensure(w, G_CONS, std::min(2u * w.nOrder, std::max(1u << 16, 16u * w.toiListBound)) + 64u);
        const unsigned P = w.nProxies;
        if (w.toiListBound == 0u) { // no TOI-candidate contact in the world: nothing to compute, no event (:1768-1777)
            w.stepComplete = true;
            if (rounds || w.toiAlphaDirty) { PR_LAUNCH(k_toi_finish_bodies, N, w.stream, d, N); w.toiAlphaDirty = false; }
            break;
        }
        // ---- batched rounds: needs an intact tree to refit in place; otherwise one host-driven round (below)
        const bool canBatch = w.bvhValid && w.bvhLeaves == P && P > 1; // (the refit budget is checked by the next regular step)
        if (canBatch) {
            if (w.hashDirty) rebuild_hash(w);
            // rounds to enqueue: what the last step needed (+1: the round that finds nothing), a few at a time after that
            const unsigned K = w.subStepping ? 1u : std::min(32u, rounds ? 4u : std::max(2u, w.toiRoundsHint + 1u));
            // late: groups of [one round over the full TOI list][up to K rounds in one block over compact lists]; the block
            // hands back (too many candidates yet, a woken body) simply by returning, and the next group's list round goes on
            const bool late = !(w.debugFlags & 8u) && !w.subStepping;
            const unsigned groups = late ? std::max(1u, std::min(4u, w.toiGroupsHint)) : 1u;
            const unsigned listRounds = late ? 1u : K;
            const unsigned batchRounds = groups * (listRounds + (late ? K : 0u)); // every round may add up to SMALL_SORT_MAX contacts
            ensure(w, G_PAIR, SMALL_SORT_MAX);
            ensure(w, G_CONTACT, std::max(w.poolSize, w.nOrder) + batchRounds * SMALL_SORT_MAX + 2u);
            if (2ull * (static_cast<unsigned long long>(w.nOrder) + batchRounds * SMALL_SORT_MAX + w.hashTombs) > w.capHash) rebuild_hash(w);
            const unsigned scratchCap = std::min(w.capCons, w.capAdj);
            read_counters(w);
            throw_device_error(w);
            w.hostCnt.toiHalt = TB_RUN; w.hostCnt.tbRounds = w.hostCnt.tbFound = w.hostCnt.tbAdded = w.hostCnt.tbPairs = 0u;
            w.hostCnt.devOrder = w.nOrder; w.hostCnt.devFree = w.nFree; w.hostCnt.devPool = w.poolSize;
            w.hostCnt.tbFindSkip = 1u; w.hostCnt.tbAddN = w.hostCnt.tbAddOld = 0u;
            write_counters(w);
            const unsigned listBound = w.toiListBound + batchRounds * SMALL_SORT_MAX; // contacts created by earlier rounds of the batch join the list
            const unsigned solveWarps = std::min(592u * 8u, std::max(8u, listBound)); // (8 warps per 256-thread block)
            // one round over the full TOI list (list kernels), then -- unless switched off (debug bit 3) or sub-stepping -- up to
            // K more rounds in ONE block over compact lists (k_tb_late); without it, K list rounds
            for (unsigned g = 0; g < groups; ++g) {
                for (unsigned k = 0; k < listRounds; ++k) {
                    const unsigned rid = late ? 0u : ++round; // (late: the device numbers the rounds)
                    PR_LAUNCH(k_toi_alpha, listBound, w.stream, d, conf.max_sub_steps, rid);
                    PR_LAUNCH(k_toi_advance, listBound, w.stream, d, conf.max_sub_steps, rid);
                    PR_LAUNCH(k_toi_compute, listBound, w.stream, d, conf);
                    if (!exact) PR_LAUNCH(k_toi_claim, listBound, w.stream, d);
                    PR_LAUNCH(k_toi_select, listBound, w.stream, d, exact ? 0u : 1u);
                    if (!exact) PR_LAUNCH(k_toi_unclaim, listBound, w.stream, d);
                    PR_LAUNCH(k_tb_solve, 32u * solveWarps, w.stream, d, conf, exact ? 1u : 0u, scratchCap);
#ifndef PR_HOST_EMUL
                    PR_LAUNCH_GRID(k_sync_body_proxies, 148, PR_BLOCK, w.stream, d, conf);
                    PR_LAUNCH_GRID(k_tb_refit, 32, PR_BLOCK, w.stream, d, P);
                    PR_LAUNCH_GRID(k_tb_query, 148, PR_BLOCK, w.stream, d, P, w.bvhSortedIdx, w.capPairs);
                    PR_LAUNCH_GRID(k_tb_add, 1, TB_ADD_BLOCK, w.stream, d, conf, w.capPairs, w.capContacts, w.capHash, w.hashTombs);
#else
                    PR_LAUNCH(k_sync_body_proxies, 1, w.stream, d, conf);
                    PR_LAUNCH(k_tb_refit, 2, w.stream, d, P);
                    PR_LAUNCH(k_tb_query, 2, w.stream, d, P, w.bvhSortedIdx, w.capPairs);
                    PR_LAUNCH(k_tb_add, 1, w.stream, d, conf, w.capPairs, w.capContacts, w.capHash, w.hashTombs);
#endif
                }
                if (late) {
#ifndef PR_HOST_EMUL
                    PR_LAUNCH_GRID(k_tb_late, 1, LATE_BLOCK, w.stream, d, conf, exact ? 1u : 0u, scratchCap, P, w.bvhSortedIdx, w.capPairs, w.capContacts,
                                   w.capHash, w.hashTombs, K);
#else
                    PR_LAUNCH(k_tb_late, 1, w.stream, d, conf, exact ? 1u : 0u, scratchCap, P, w.bvhSortedIdx, w.capPairs, w.capContacts, w.capHash,
                              w.hashTombs, K);
#endif
                }
            }
            ++toiBatches;
            read_counters(w);
            const Counters c = w.hostCnt;
            if (getenv("PR_TOI_DEBUG"))
                std::fprintf(stderr, "toi batch: groups %u K %u rounds %u found %u halt %u | late: woken %u bigR %u bigC %u maxC %u maxR %u lateRounds %u candN %u listRoundEvents(base 1e5) %u listRounds %u\n", groups, K,
                             c.tbRounds, c.tbFound, c.toiHalt, c.lateWhy[0], c.lateWhy[1], c.lateWhy[2], c.lateWhy[3], c.lateWhy[4], c.lateWhy[5], c.toiCandN, c.lateWhy[6], c.lateWhy[7]);
            listRoundsUsed = c.lateWhy[7]; // (list rounds of this step that had events)
            round = c.toiRoundId; // (the device opened this many rounds so far this step)
            w.nOrder = c.devOrder; w.nFree = c.devFree; w.poolSize = c.devPool;
            w.toiListBound += c.tbAdded; w.bvhRefits += c.tbRounds;
            rounds += c.tbRounds;
            st.contacts_found += c.tbFound; st.contacts_added += c.tbAdded;
            w.profile.candidate_pairs += c.tbPairs;
            const unsigned halt = (c.toiHalt == TB_BIG) ? static_cast<unsigned>(TB_RUN) : c.toiHalt; // (TB_BIG: the next batch starts with a list round anyway)
            w.hostCnt.toiHalt = TB_RUN;
            if (halt == TB_ERROR) { write_counters(w); throw_device_error(w); throw CudaError{"TOI rounds halted on a device error"}; }
            if (halt == TB_ADJ || c.adjOverflow) { w.adjValid = false; w.hostCnt.adjOverflow = 0; }
            write_counters(w);
            if (halt == TB_HOST_FIND) { // the round's events are solved; its FindContacts/AddContacts needs the host (growth / rebuild)
                tm.mark(w, PR_STAGE_TOI_SOLVE);
                st.contacts_added += find_add_contacts(w, conf, tm, true);
                tm.mark(w, PR_STAGE_TOI);
            }
            if (halt == TB_DONE) { // :1768-1777
                w.stepComplete = true;
                if (rounds || w.toiAlphaDirty) { PR_LAUNCH(k_toi_finish_bodies, N, w.stream, d, N); w.toiAlphaDirty = false; }
                if (w.toiListBound) PR_LAUNCH(k_toi_finish_contacts, w.toiListBound, w.stream, d);
                break;
            }
            if (w.subStepping && c.tbRounds) { w.stepComplete = false; w.toiAlphaDirty = true; break; }
            continue;
        }
        // ---- one host-driven round (tree or hash about to be rebuilt)
        const unsigned scratchCap = std::min(w.capCons, w.capAdj);
        const unsigned nList = w.toiListBound; // >= device toiListN (kernels re-check against it)
        ++round;
        if (nList) {
            PR_LAUNCH(k_toi_alpha, nList, w.stream, d, conf.max_sub_steps, round);
            PR_LAUNCH(k_toi_advance, nList, w.stream, d, conf.max_sub_steps, round);
            PR_LAUNCH(k_toi_compute, nList, w.stream, d, conf);
            if (!exact) PR_LAUNCH(k_toi_claim, nList, w.stream, d);
            PR_LAUNCH(k_toi_select, nList, w.stream, d, exact ? 0u : 1u);
            if (!exact) PR_LAUNCH(k_toi_unclaim, nList, w.stream, d);
        }
        read_counters(w);
        if (w.hostCnt.toiHalt == TB_ADJ) { w.adjValid = false; w.hostCnt.adjOverflow = 0; w.hostCnt.toiHalt = TB_RUN; write_counters(w); --round; continue; }
        if (w.hostCnt.toiHalt == TB_ERROR) throw_device_error(w);
        const unsigned nEvents = nList ? w.hostCnt.toiEvents : 0u;
        if (nEvents == 0) { // :1768-1777
            w.stepComplete = true;
            if (rounds || w.toiAlphaDirty) { PR_LAUNCH(k_toi_finish_bodies, N, w.stream, d, N); w.toiAlphaDirty = false; }
            if (nList) PR_LAUNCH(k_toi_finish_contacts, nList, w.stream, d);
            break;
        }
        ++rounds;
        tm.mark(w, PR_STAGE_TOI_SOLVE);
        st.contacts_found += nEvents; // (events deferred for lack of scratch are subtracted below)
        PR_LAUNCH(k_toi_solve, 32u * nEvents, w.stream, d, conf, nEvents, exact ? 1u : 0u, scratchCap);
#ifndef PR_HOST_EMUL
        PR_LAUNCH_GRID(k_sync_body_proxies, 148, PR_BLOCK, w.stream, d, conf);
#else
        PR_LAUNCH(k_sync_body_proxies, 1, w.stream, d, conf);
#endif
        st.contacts_added += find_add_contacts(w, conf, tm, true);
        tm.mark(w, PR_STAGE_TOI);
        if (w.subStepping) { w.stepComplete = false; w.toiAlphaDirty = true; break; }
    }
    w.toiRoundsHint = rounds;
    // groups to enqueue next step = list rounds this step needed (a group beyond that is ~12 launches that find the halt flag set)
    if (toiBatches) w.toiGroupsHint = std::min(4u, std::max(1u, listRoundsUsed));
    // final UpdateContacts (:1825-1827)
    zero_counter(w, &Counters::contactsUpdated); // (proxiesMoved is read below: the final UpdateContacts does not touch it)
    const unsigned nOrderAtUpdate = w.nOrder;
    // (no event this step: nothing was flagged since the regular phase's UpdateContacts -- every contact is skipped)
    if (rounds) PR_LAUNCH(k_update_contacts, w.nOrder, w.stream, d, conf, 0u, w.nOrder, 0u, 1u);
    read_counters(w);
    const Counters c = w.hostCnt;
    st.contacts_at_max_sub_steps = c.toiAtMaxSub; st.contacts_updated_toi = c.toiUpdated;
    st.max_dist_iters = c.toiMaxDistIters; st.max_toi_iters = c.toiMaxToiIters; st.max_root_iters = c.toiMaxRootIters;
    st.min_separation = float_unflip(c.toiMinSepFlip); st.max_inc_impulse = u2f(c.toiMaxIncBits);
    st.islands_found = c.toiIslandsFound; st.islands_solved = c.toiIslandsSolved;
    st.sum_pos_iters = c.toiSumPosIters; st.sum_vel_iters = c.toiSumVelIters;
    st.contacts_updated_touching = c.toiUpdTouching + c.contactsUpdated;
    st.contacts_skipped_touching = c.toiSkipTouching + (nOrderAtUpdate - c.contactsUpdated);
    st.proxies_moved = c.proxiesMoved;
    st.contacts_found -= c.toiDeferred;
    if (d.toi_log) {
        w.toiScheduleN = c.toiLogN;
        w.toiSchedule.resize(std::min(c.toiLogN, d.toi_logCap));
        d2h(w, w.toiSchedule.data(), d.toi_log, sizeof(uint2) * w.toiSchedule.size());
        std::sort(w.toiSchedule.begin(), w.toiSchedule.end(), [](const uint2& a, const uint2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    }
    w.profile.toi_rounds = rounds; w.profile.toi_events = st.contacts_found;
#ifdef PR_TOI_TIMING
    std::fprintf(stderr, "late_timing rounds %llu events %llu nR %llu nCand %llu | cyc: entry %llu tois %llu select %llu solve %llu sync %llu refitquery %llu add %llu\n",
                 c.dbgLate[8], c.dbgLate[9], c.dbgLate[10], c.dbgLate[11], c.dbgLate[0], c.dbgLate[1], c.dbgLate[2], c.dbgLate[3], c.dbgLate[4], c.dbgLate[5], c.dbgLate[6]);
    std::fprintf(stderr, "toi_timing rounds=%u ev0=%llu | cyc: head %llu procA %llu procB %llu build %llu pos %llu vel %llu tailwb %llu | nC %llu nB %llu posIt %llu velIt %llu upd %llu skip %llu\n",
                 rounds, c.dbg[8], c.dbg[0], c.dbg[1], c.dbg[2], c.dbg[3], c.dbg[4], c.dbg[5], c.dbg[6], c.dbg[9], c.dbg[10], c.dbg[11], c.dbg[12], c.dbg[13], c.dbg[14]);
#endif
}

static void reset_counters(World& w)
{
    std::memset(&w.hostCnt, 0, sizeof(Counters));
    w.hostCnt.minSepFlip = float_flip(std::numeric_limits<float>::infinity());
    write_counters(w);
}

static void step_impl(World& w, const pr_step_conf& conf, pr_step_stats& stats)
{
    DV& d = w.d;
    if (!w.timer) w.timer = new StageTimer(); // (the world's own: its CUDA events belong to the world's device)
    StageTimer& tm = *static_cast<StageTimer*>(w.timer);
    std::memset(&stats, 0, sizeof(stats));
    stats.reg.min_separation = std::numeric_limits<float>::infinity();
    stats.toi.min_separation = std::numeric_limits<float>::infinity();
    std::memset(&w.profile, 0, sizeof(w.profile));
    reset_launch_count();
    tm.mark(w, PR_STAGE_PRE);
    // keep the queue count across steps (pr_set_bodies / creation between steps feed it)
    read_counters(w);
    const unsigned keepQueue = w.hostCnt.nQueue, keepToiList = w.hostCnt.toiListN; // state that outlives a step
    if (w.hostCnt.recolorFlag) w.recolorAll = true; // a SetBody changed which bodies carry mass: conflict sets changed
    reset_counters(w);
    w.hostCnt.nQueue = keepQueue; w.hostCnt.toiListN = keepToiList; write_counters(w);
    stats.pre.proxies_created = create_queued_proxies(w, conf);
    if (w.anyBodySet) PR_LAUNCH(k_sync_proxies, w.nProxies, w.stream, d, conf, w.nProxies);
    const bool queueSeen = w.nOrder != 0u; // (destroy_contacts reads the counters then, after the last kernel of the pre-phase that queues a proxy)
    stats.pre.contacts_destroyed = destroy_contacts(w);
    const unsigned queueN = w.hostCnt.nQueue;
    const unsigned nOrderAtUpdate = w.nOrder;
    // UpdateContacts of the pre-phase (:2190): nothing to look at when the last Step ended with every contact updated and no
    // mutator ran since (the usual step-after-step loop): every contact is skipped, as the reference finds out one by one
    const bool preClean = w.contactsClean && !w.anyBodySet;
    const bool hadBodySet = w.anyBodySet;
    if (!preClean) PR_LAUNCH(k_update_contacts, w.nOrder, w.stream, d, conf, 0u, w.nOrder, w.anyBodySet ? 1u : 0u, 1u);
    if (w.anyBodySet) { dset(w, d.b_moved, 0, sizeof(unsigned) * w.nBodies); dset(w, d.b_sync, 0, sizeof(unsigned) * w.nBodies); }
    // the usual step-after-step loop: no proxy was synchronised and no contact updated in this pre-phase (both counters are what the
    // reset above left: 0) and the proxy queue is what destroy_contacts' read saw -- nothing to read back
    const bool quietPre = preClean && !hadBodySet && queueSeen && queueN == 0u;
    w.anyBodySet = false;
    w.contactsClean = false;
    if (!quietPre) read_counters(w);
    stats.pre.proxies_moved = quietPre ? 0u : w.hostCnt.proxiesMoved;
    stats.pre.contacts_updated = quietPre ? 0u : w.hostCnt.contactsUpdated;
    stats.pre.contacts_skipped = nOrderAtUpdate - stats.pre.contacts_updated;
    stats.pre.contacts_added = find_add_contacts(w, conf, tm, false, quietPre);
    if (conf.delta_time != 0.0f) {
        w.inv_dt0 = 1.0f / conf.delta_time;
        if (w.stepComplete) solve_reg(w, conf, stats, tm);
        if (conf.do_toi) solve_toi(w, conf, stats, tm);
    }
    // every contact that was flagged or whose body moved has been updated by now (regular phase :1547, the TOI phase's final
    // UpdateContacts :1825, AddContacts' own Update) unless a sub-stepped TOI phase is still open
    w.contactsClean = w.stepComplete;
    check_launch();
    w.profile.p_total = w.nProxies; w.profile.c_total = w.nOrder;
    tm.finish(w, w.profile);
    sync(w);
    w.profile.kernel_launches = launches_since_reset();
}

} // namespace pr

// =============================================================================================
using namespace pr;
struct pr_world { World w; };

// every entry point runs with the world's device current and puts the caller's device back (two worlds on two GPUs in one
// process, or a caller -- torch -- that changed the current device since pr_create)
#ifndef PR_HOST_EMUL
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess || prev == dev) prev = -1; else cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#else
struct DeviceGuard { explicit DeviceGuard(int) {} };
#endif
#define API_TRY(wptr) try { DeviceGuard pr_guard_((wptr)->w.device);
#define API_CATCH(wptr)                                                                            \
    } catch (const CudaError& e) { (wptr)->w.error = e.msg; return PR_ERR_CUDA; }                  \
    catch (const std::exception& e) { (wptr)->w.error = e.what(); return PR_ERR_CUDA; }

extern "C" {

int pr_abi_version(void) { return PR_ABI_VERSION; }
// which glibc sincosf ifunc variant pr::glibc_sincosf restates in this build (pr_math.cuh): 0 = generic, 1 = x86-64 "-fma"
int pr_sincosf_variant(void)
{
#ifdef PR_SINCOSF_FMA
    return 1;
#else
    return 0;
#endif
}

// which variant this host's glibc selects for sincosf (sysdeps/x86_64/fpu/multiarch/ifunc-fma.h: the "-fma" one on CPUs with
// usable FMA and AVX2, else the generic one) -- i.e. the bits the reference computes when it runs on this host
int pr_host_sincosf_variant(void)
{
#if defined(__x86_64__) && !defined(__CUDA_ARCH__)
    __builtin_cpu_init();
    return (__builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2")) ? 1 : 0;
#else
    return 0;
#endif
}

void pr_world_conf_default(pr_world_conf* c)
{
    c->min_vertex_radius = 0.01f; c->max_vertex_radius = 255.0f; // WorldConf::DefaultVertexRadius
    c->body_capacity = c->proxy_capacity = c->contact_capacity = c->vertex_capacity = 0;
    c->device = -1;
}

void pr_step_conf_default(pr_step_conf* c)
{ // StepConf{} defaults: Settings.hpp:92-256, StepConf.hpp:163-391 (checked against the reference in tests)
    const float pi = 3.14159265358979323846264338327950288f;
    const float slop = 0.005f;
    c->delta_time = 1.0f / 60; c->dt_ratio = 1.0f;
    c->min_still_time_to_sleep = 1.0f / 2; c->linear_slop = slop; c->angular_slop = (pi * 2 * 1.0f) / 180.0f;
    c->reg_resolution_rate = 2.0f / 10; c->reg_min_separation = -slop * 3; c->reg_min_momentum = 0.0f / 100;
    c->toi_resolution_rate = 75.0f / 100; c->toi_min_separation = -slop * 1.5f; c->toi_min_momentum = 0.0f / 100;
    c->target_depth = slop * 3; c->tolerance = slop / 4; c->velocity_threshold = 1.0f;
    c->max_translation = 2.0f; c->max_rotation = 179.0f * (pi / 180.0f);
    c->max_linear_correction = 0.2f; c->max_angular_correction = (8.0f / 180.0f) * pi * 1.0f;
    c->linear_sleep_tolerance = 0.01f; c->angular_sleep_tolerance = 2.0f * (pi / 180.0f);
    c->displace_multiplier = 2.0f; c->aabb_extension = slop * 20.0f; c->max_circles_ratio = 10.0f;
    c->reg_velocity_iters = 8; c->reg_position_iters = 3; c->toi_velocity_iters = 8; c->toi_position_iters = 20;
    c->max_toi_root_iters = 30; c->max_toi_iters = 20; c->max_distance_iters = 20; c->max_sub_steps = 8;
    c->do_warm_start = 1; c->do_toi = 1; c->do_blocksolve = 1;
}

int pr_create(const pr_world_conf* conf, pr_world** out)
{
    *out = nullptr;
    pr_world* pw = nullptr;
    try {
        pw = new pr_world();
        World& w = pw->w;
        if (conf) w.conf = *conf; else pr_world_conf_default(&w.conf);
#ifndef PR_HOST_EMUL
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_createError = "no CUDA device: the B200 path has no CPU fallback"; delete pw; return PR_ERR_CUDA; }
        if (w.conf.device >= 0) CK(cudaSetDevice(w.conf.device));
        CK(cudaGetDevice(&w.device));
        CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
#endif
        w.d.cnt = static_cast<Counters*>(dev_alloc(sizeof(Counters)));
#ifndef PR_HOST_EMUL
        CK(cudaHostAlloc(reinterpret_cast<void**>(&w.pinnedCnt), sizeof(Counters) * World::PINNED_SLOTS, cudaHostAllocDefault));
#endif
        ensure(w, G_BODY, std::max(64u, w.conf.body_capacity));
        ensure(w, G_SHAPE, 64); ensure(w, G_CHILD, 64); ensure(w, G_VERT, std::max(256u, w.conf.vertex_capacity));
        ensure(w, G_PROXY, std::max(64u, w.conf.proxy_capacity));
        ensure(w, G_CONTACT, std::max(256u, w.conf.contact_capacity));
        // The derived buffers start in proportion to the configured contact capacity, so that a world sized by its caller
        // never re-allocates while stepping: a growth is cudaMalloc + copy + cudaFree of up to GBs, measured at 10-400 ms
        // on a shared B200 host (three such steps in a 50-step timed window cost 10% of the mean step time).
        const unsigned cc = w.conf.contact_capacity, bc = w.conf.body_capacity;
        ensure(w, G_PAIR, std::max(1024u, cc / 8u));
        ensure(w, G_CONS, std::max(256u, cc / 3u));
        ensure(w, G_ADJ, std::max(256u, cc ? 2u * cc + cc / 2u + 8u * bc + 64u : 0u));
        ensure(w, G_JOINT, 16);
        ensure(w, G_HASH, next_pow2(2u * w.capContacts));
        reset_counters(w);
        sync(w);
        *out = pw;
        return PR_OK;
    } catch (const CudaError& e) { g_createError = e.msg; delete pw; return PR_ERR_CUDA; }
}

int pr_destroy(pr_world* pw)
{
    if (!pw) return PR_OK;
    World& w = pw->w;
    DeviceGuard guard(w.device);
    delete static_cast<StageTimer*>(w.timer);
    for (const ArrDesc& a : arrays(w)) dev_free(*a.slot);
    dev_free(w.d.ev_contact); dev_free(w.d.ev_impulse); dev_free(w.d.ev_presolve);
    dev_free(w.d.is_keys); dev_free(w.d.blk_off); dev_free(w.d.blk_rowOff); dev_free(w.d.blk_jOff); dev_free(w.d.toi_log); dev_free(w.d.cnt); dev_free(w.scanScratch); dev_free(w.rs.hist); dev_free(w.stageBuf[0]); dev_free(w.stageBuf[1]);
#ifndef PR_HOST_EMUL
    if (w.stream) cudaStreamDestroy(w.stream);
    if (w.pinnedCnt) cudaFreeHost(w.pinnedCnt);
#endif
    delete pw;
    return PR_OK;
}

const char* pr_last_error(const pr_world* pw) { return pw ? pw->w.error.c_str() : g_createError.c_str(); }

int pr_clone(const pr_world* src, pr_world** out)
{
    pr_world* pw = nullptr;
    pr_world_conf conf = src->w.conf;
    conf.device = src->w.device; // (the clone lives where its source lives, whatever device is current)
    const int rc = pr_create(&conf, &pw);
    if (rc != PR_OK) return rc;
    API_TRY(pw)
    World& w = pw->w;
    World& s = const_cast<World&>(src->w);
    sync(s);
    ensure(w, G_BODY, s.capBodies); ensure(w, G_SHAPE, s.capShapes); ensure(w, G_CHILD, s.capChildren); ensure(w, G_VERT, s.capVerts);
    ensure(w, G_PROXY, s.capProxies); ensure(w, G_CONTACT, s.capContacts); ensure(w, G_HASH, s.capHash);
    ensure(w, G_PAIR, s.capPairs); ensure(w, G_CONS, s.capCons); ensure(w, G_ADJ, s.capAdj); ensure(w, G_JOINT, s.capJoints);
    auto da = arrays(w); auto sa = arrays(s);
    for (size_t i = 0; i < da.size(); ++i) d2d(w, *da[i].slot, *sa[i].slot, da[i].elem * *cap_of(s, sa[i].g));
    d2d(w, w.d.cnt, s.d.cnt, sizeof(Counters));
    w.shapes = s.shapes; w.children = s.children; w.bodyShapes = s.bodyShapes; w.joints = s.joints; w.freeJoints = s.freeJoints;
    w.fixturesForProxies = s.fixturesForProxies; w.bodyProxies = s.bodyProxies; w.bodyJointList = s.bodyJointList; w.pools = s.pools; w.needsFiltering = s.needsFiltering;
    w.bpDirty = true; w.bvhValid = false; w.toiAlphaDirty = true; w.recolorAll = s.recolorAll; w.toiPossible = s.toiPossible; w.colorCap = s.colorCap;
    w.nJoints = s.nJoints; w.jointsSerial = s.jointsSerial; w.jointLevelOff = s.jointLevelOff; w.d.nNcPairs = s.d.nNcPairs;
    w.d.nWorlds = s.d.nWorlds; w.d.worldBits = s.d.worldBits; w.bodyWorld = s.bodyWorld; w.blocksDirty = w.blkJointsDirty = true;
    w.nBodies = s.nBodies; w.nShapes = s.nShapes; w.nChildren = s.nChildren; w.nVerts = s.nVerts; w.nProxies = s.nProxies;
    w.maxTreeId = s.maxTreeId; w.debugFlags = s.debugFlags;
    w.nOrder = s.nOrder; w.nFree = s.nFree; w.poolSize = s.poolSize; w.anyBodySet = s.anyBodySet; w.hashDirty = true;
    w.stepComplete = s.stepComplete; w.inv_dt0 = s.inv_dt0; w.subStepping = s.subStepping; w.toiMode = s.toiMode;
    w.adjValid = false;
    sync(w);
    *out = pw;
    return PR_OK;
    API_CATCH(pw)
}

// ------------------------------------------------------------------ snapshot / restore
// What pr_clone copies device to device -- every array of every group, the counters, the host-side tables and scalars -- goes
// through one self-describing byte stream instead: a world restored from it continues exactly like a clone does (= like the
// source: tests/test_snapshot.py), which a pr_load_state of bodies/proxies/contacts cannot (the contact pool's free stack,
// the persistent colours and the m_contacts order are history). Layout: header | scalars | host tables | counters | arrays
// (element size, used count, bytes; groups that the next step rebuilds anyway -- hash, candidate pairs, constraint rows,
// adjacency -- are left out and flagged dirty on load, like pr_clone flags them).
extern "C++" {
namespace {
constexpr unsigned long long SNAP_MAGIC = 0x30325042524e5350ull; // "PSNRBP20"
struct SnapHeader {
    unsigned long long magic; unsigned version, abi, variant, nArrays;
    unsigned szBody, szContact, szCounters, pad;
};
struct SnapScalars {
    pr_world_conf conf;
    unsigned nBodies, nShapes, nChildren, nVerts, nProxies, nOrder, nFree, poolSize, maxTreeId, debugFlags, nJoints, nWorlds, worldBits, nNcPairs;
    unsigned colorsAtFull, colorCap, toiMode, toiRoundsHint, toiGroupsHint, isHint, evMask, evCap;
    unsigned anyBodySet, stepComplete, recolorAll, subStepping, jointsSerial, toiPossible;
    float inv_dt0;
};
struct SnapWriter {
    unsigned char* p; unsigned long long cap, n = 0; bool ok = true;
    void put(const void* src, size_t bytes) { if (p && n + bytes <= cap) std::memcpy(p + n, src, bytes); else if (p) ok = false; n += bytes; }
    template <class T> void pod(const T& v) { put(&v, sizeof(T)); }
    template <class T> void vec(const std::vector<T>& v) { const unsigned long long k = v.size(); pod(k); if (k) put(v.data(), sizeof(T) * k); }
};
struct SnapReader {
    const unsigned char* p; unsigned long long cap, n = 0; bool ok = true;
    void get(void* dst, size_t bytes) { if (n + bytes <= cap) std::memcpy(dst, p + n, bytes); else ok = false; n += bytes; }
    template <class T> void pod(T& v) { get(&v, sizeof(T)); }
    template <class T> void vec(std::vector<T>& v) { unsigned long long k = 0; pod(k); if (!ok || k > (cap - n) / sizeof(T) + 1ull) { ok = false; return; } v.resize(static_cast<size_t>(k)); if (k) get(v.data(), sizeof(T) * k); }
};
// elements of group g that carry state across steps (0: the group is rebuilt by the next step)
unsigned snap_used(const World& w, Group g)
{
    switch (g) {
    case G_BODY: return w.nBodies + 1u; case G_SHAPE: return w.nShapes; case G_CHILD: return w.nChildren; case G_VERT: return w.nVerts;
    case G_PROXY: return w.nProxies + 1u; case G_CONTACT: return std::max(w.poolSize, w.nOrder) + 1u; case G_JOINT: return static_cast<unsigned>(w.joints.size()) + 1u;
    default: return 0u; // G_HASH, G_PAIR, G_CONS, G_ADJ
    }
}
void snap_scalars(const World& w, SnapScalars& s)
{
    std::memset(&s, 0, sizeof(s));
    s.conf = w.conf;
    s.nBodies = w.nBodies; s.nShapes = w.nShapes; s.nChildren = w.nChildren; s.nVerts = w.nVerts; s.nProxies = w.nProxies;
    s.nOrder = w.nOrder; s.nFree = w.nFree; s.poolSize = w.poolSize; s.maxTreeId = w.maxTreeId; s.debugFlags = w.debugFlags; s.nJoints = w.nJoints;
    s.nWorlds = w.d.nWorlds; s.worldBits = w.d.worldBits; s.nNcPairs = w.d.nNcPairs;
    s.colorsAtFull = w.colorsAtFull; s.colorCap = w.colorCap; s.toiMode = w.toiMode; s.toiRoundsHint = w.toiRoundsHint; s.toiGroupsHint = w.toiGroupsHint;
    s.isHint = w.isHint; s.evMask = w.d.ev_mask; s.evCap = w.d.ev_cap;
    s.anyBodySet = w.anyBodySet; s.stepComplete = w.stepComplete; s.recolorAll = w.recolorAll; s.subStepping = w.subStepping;
    s.jointsSerial = w.jointsSerial; s.toiPossible = w.toiPossible; s.inv_dt0 = w.inv_dt0;
}
template <class T> void snap_nested(SnapWriter& o, const std::vector<std::vector<T>>& v)
{
    const unsigned long long k = v.size(); o.pod(k);
    for (const auto& e : v) o.vec(e);
}
template <class T> void snap_nested(SnapReader& in, std::vector<std::vector<T>>& v)
{
    unsigned long long k = 0; in.pod(k);
    if (!in.ok || k > in.cap) { in.ok = false; return; }
    v.resize(static_cast<size_t>(k));
    for (auto& e : v) { in.vec(e); if (!in.ok) return; }
}
} // namespace
} // extern "C++"

int pr_snapshot_save(const pr_world* cpw, void* buf, uint64_t cap, uint64_t* bytes_out)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    sync(w);
    SnapWriter o{static_cast<unsigned char*>(buf), cap};
    auto arrs = arrays(w);
    SnapHeader h{SNAP_MAGIC, 3u, static_cast<unsigned>(PR_ABI_VERSION), static_cast<unsigned>(pr_sincosf_variant()), static_cast<unsigned>(arrs.size()),
                 static_cast<unsigned>(sizeof(pr_body)), static_cast<unsigned>(sizeof(pr_contact)), static_cast<unsigned>(sizeof(Counters)), 0u};
    o.pod(h);
    SnapScalars sc; snap_scalars(w, sc); o.pod(sc);
    o.vec(w.shapes); o.vec(w.children); snap_nested(o, w.bodyShapes); o.vec(w.joints); o.vec(w.freeJoints);
    { std::vector<unsigned> flat; for (const auto& f : w.fixturesForProxies) { flat.push_back(f.first); flat.push_back(f.second); } o.vec(flat); }
    snap_nested(o, w.bodyProxies); snap_nested(o, w.bodyJointList); o.vec(w.bodyWorld);
    o.vec(w.pools.bodyFree); o.vec(w.pools.shapeFree); o.vec(w.pools.treeFree); o.vec(w.pools.bodyDead); o.vec(w.pools.shapeDead);
    o.vec(w.pools.pxBody); o.vec(w.pools.pxShape); o.vec(w.pools.pxTree);
    { const unsigned extra[2] = {w.pools.treeNextFresh, w.needsFiltering ? 1u : 0u}; o.pod(extra); }
    read_counters(w);
    o.pod(w.hostCnt);
    std::vector<unsigned char> tmp;
    for (const ArrDesc& a : arrs) {
        const unsigned long long used = snap_used(w, a.g), elem = a.elem;
        o.pod(elem); o.pod(used);
        const size_t bytes = static_cast<size_t>(elem * used);
        if (bytes) {
            if (o.p && o.n + bytes <= o.cap) d2h(w, o.p + o.n, *a.slot, bytes);
            else if (o.p) o.ok = false;
            o.n += bytes;
        }
    }
    if (bytes_out) *bytes_out = o.n;
    if (buf && !o.ok) { w.error = "pr_snapshot_save: buffer too small"; return PR_ERR_LENGTH; }
    return PR_OK;
    API_CATCH(pw)
}

int pr_snapshot_load(const void* buf, uint64_t bytes, int device, pr_world** out)
{
    *out = nullptr;
    SnapReader in{static_cast<const unsigned char*>(buf), bytes};
    SnapHeader h{}; in.pod(h);
    SnapScalars sc{}; in.pod(sc);
    if (!in.ok || h.magic != SNAP_MAGIC || h.version != 3u || h.abi != static_cast<unsigned>(PR_ABI_VERSION) || h.szBody != sizeof(pr_body) ||
        h.szContact != sizeof(pr_contact) || h.szCounters != sizeof(Counters)) { g_createError = "pr_snapshot_load: not a snapshot of this library version"; return PR_ERR_INVALID_ARGUMENT; }
    if (h.variant != static_cast<unsigned>(pr_sincosf_variant())) { g_createError = "pr_snapshot_load: the snapshot was taken by the other sincosf build of the library"; return PR_ERR_INVALID_ARGUMENT; }
    pr_world* pw = nullptr;
    pr_world_conf conf = sc.conf;
    conf.device = device;
    const int rc = pr_create(&conf, &pw);
    if (rc != PR_OK) return rc;
    try {
        DeviceGuard guard(pw->w.device);
        World& w = pw->w;
        in.vec(w.shapes); in.vec(w.children); snap_nested(in, w.bodyShapes); in.vec(w.joints); in.vec(w.freeJoints);
        { std::vector<unsigned> flat; in.vec(flat); for (size_t i = 0; i + 1 < flat.size(); i += 2) w.fixturesForProxies.emplace_back(flat[i], flat[i + 1]); }
        snap_nested(in, w.bodyProxies); snap_nested(in, w.bodyJointList); in.vec(w.bodyWorld);
        in.vec(w.pools.bodyFree); in.vec(w.pools.shapeFree); in.vec(w.pools.treeFree); in.vec(w.pools.bodyDead); in.vec(w.pools.shapeDead);
        in.vec(w.pools.pxBody); in.vec(w.pools.pxShape); in.vec(w.pools.pxTree);
        { unsigned extra[2] = {0u, 0u}; in.pod(extra); w.pools.treeNextFresh = extra[0]; w.needsFiltering = extra[1] != 0u; }
        in.pod(w.hostCnt);
        if (!in.ok) throw CudaError{"pr_snapshot_load: truncated snapshot"};
        w.nBodies = sc.nBodies; w.nShapes = sc.nShapes; w.nChildren = sc.nChildren; w.nVerts = sc.nVerts; w.nProxies = sc.nProxies;
        w.nOrder = sc.nOrder; w.nFree = sc.nFree; w.poolSize = sc.poolSize; w.maxTreeId = sc.maxTreeId; w.debugFlags = sc.debugFlags; w.nJoints = sc.nJoints;
        w.d.nWorlds = sc.nWorlds; w.d.worldBits = sc.worldBits; w.d.nNcPairs = sc.nNcPairs;
        w.colorsAtFull = sc.colorsAtFull; w.colorCap = sc.colorCap; w.toiMode = sc.toiMode; w.toiRoundsHint = sc.toiRoundsHint; w.toiGroupsHint = sc.toiGroupsHint;
        w.isHint = sc.isHint;
        w.anyBodySet = sc.anyBodySet != 0; w.stepComplete = sc.stepComplete != 0; w.recolorAll = sc.recolorAll != 0; w.subStepping = sc.subStepping != 0;
        w.jointsSerial = sc.jointsSerial != 0; w.toiPossible = sc.toiPossible != 0; w.inv_dt0 = sc.inv_dt0;
        w.jointLevelOff.clear(); w.jointTablesDirty = true; // (the level lists are rebuilt by the next Step)
        ensure(w, G_BODY, w.nBodies + 2u); ensure(w, G_SHAPE, w.nShapes + 1u); ensure(w, G_CHILD, w.nChildren + 1u); ensure(w, G_VERT, w.nVerts + 1u);
        ensure(w, G_PROXY, w.nProxies + 2u); ensure(w, G_CONTACT, std::max(w.poolSize, w.nOrder) + 2u); ensure(w, G_JOINT, static_cast<unsigned>(w.joints.size()) + 2u);
        auto arrs = arrays(w);
        if (arrs.size() != h.nArrays) throw CudaError{"pr_snapshot_load: array table mismatch"};
        for (const ArrDesc& a : arrs) {
            unsigned long long elem = 0, used = 0; in.pod(elem); in.pod(used);
            if (!in.ok || elem != a.elem || used > *cap_of(w, a.g)) throw CudaError{"pr_snapshot_load: array does not fit this build"};
            const size_t nbytes = static_cast<size_t>(elem * used);
            if (nbytes) {
                if (in.n + nbytes > in.cap) throw CudaError{"pr_snapshot_load: truncated snapshot"};
                h2d(w, *a.slot, in.p + in.n, nbytes);
                in.n += nbytes;
            }
        }
        write_counters(w);
        if (sc.evMask) { const int erc = pr_set_event_mask(pw, sc.evMask, std::max(1u, sc.evCap)); if (erc != PR_OK) throw CudaError{"pr_snapshot_load: event buffers"}; }
        w.bpDirty = true; w.bvhValid = false; w.toiAlphaDirty = true; w.hashDirty = true; w.adjValid = false; w.blocksDirty = w.blkJointsDirty = true;
        sync(w);
    } catch (const CudaError& e) { g_createError = e.msg; pr_destroy(pw); return PR_ERR_INVALID_ARGUMENT; }
    catch (const std::exception& e) { g_createError = e.what(); pr_destroy(pw); return PR_ERR_INVALID_ARGUMENT; }
    *out = pw;
    return PR_OK;
}

int pr_clear(pr_world* pw)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    w.shapes.clear(); w.children.clear(); w.bodyShapes.clear(); w.joints.clear(); w.fixturesForProxies.clear();
    w.bodyProxies.clear(); w.bodyJointList.clear(); w.bpDirty = true; w.bvhValid = false; w.hashTombs = 0; w.toiAlphaDirty = true;
    w.nJoints = 0; w.jointsSerial = false; w.jointLevelOff.clear(); w.d.nNcPairs = 0; w.bodyWorld.clear(); w.freeJoints.clear();
    w.pools = World::Pools{}; w.needsFiltering = false;
    w.nBodies = w.nShapes = w.nChildren = w.nVerts = w.nProxies = w.nOrder = w.nFree = w.poolSize = 0; w.maxTreeId = 0;
    w.anyBodySet = false; w.hashDirty = true; w.adjValid = false; w.stepComplete = true; w.toiPossible = false;
    w.inv_dt0 = 0.0f; // m_inv_dt0 = 0_Hz (AabbTreeWorld.cpp:982)
    w.recolorAll = true; w.colorsAtFull = 0; w.colorCap = 64u; w.nCons = w.nColors = 0; w.toiRoundsHint = 1; w.toiGroupsHint = 1; w.isHint = 0;
    if (w.d.ev_mask) { w.hostCnt.nEvContact = w.hostCnt.nEvImpulse = 0u; }
    reset_counters(w);
    return PR_OK;
    API_CATCH(pw)
}

int pr_create_shapes(pr_world* pw, uint32_t n, const pr_shape* shapes, const pr_child* children, uint32_t n_children,
                     const pr_vec2* vertices, const pr_vec2* normals, uint32_t n_vertices, uint32_t* first_id)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    // CreateShape validation (AabbTreeWorld.cpp:1234-1255): every child's vertex radius within the world interval
    for (uint32_t i = 0; i < n_children; ++i) {
        if (children[i].vertex_count == 0 || children[i].vertex_count > 254) { w.error = "child vertex count out of range"; return PR_ERR_INVALID_ARGUMENT; }
        if (!(children[i].vertex_radius >= w.conf.min_vertex_radius && children[i].vertex_radius <= w.conf.max_vertex_radius)) {
            w.error = "vertex radius outside world interval"; return PR_ERR_INVALID_ARGUMENT;
        }
        if (children[i].vertex_offset > n_vertices || children[i].vertex_count > n_vertices - children[i].vertex_offset) { w.error = "child vertex range"; return PR_ERR_OUT_OF_RANGE; }
    }
    for (uint32_t i = 0; i < n; ++i) // (every row is checked before any is applied)
        if (shapes[i].child_offset > n_children || shapes[i].child_count > n_children - shapes[i].child_offset) { w.error = "shape child range"; return PR_ERR_OUT_OF_RANGE; }
    ensure(w, G_SHAPE, w.nShapes + n); ensure(w, G_CHILD, w.nChildren + n_children); ensure(w, G_VERT, w.nVerts + n_vertices + 4u); // (+4: the TMA-staged narrowphase copies whole 16-byte lines)
    // ids of destroyed shapes are handed out again first, last freed first (ObjectPool::Allocate)
    const bool reuse = !w.pools.shapeFree.empty();
    if (reuse && n != 1u) { w.error = "CreateShape: create shapes one at a time while destroyed shape ids are free"; return PR_ERR_UNSUPPORTED; }
    const unsigned firstShape = reuse ? w.pools.shapeFree.back() : w.nShapes;
    if (reuse) { w.pools.shapeFree.pop_back(); w.pools.shapeDead[firstShape] = 0; }
    std::vector<float4> mat(n); std::vector<unsigned> filt(n); std::vector<int> grp(n);
    for (uint32_t i = 0; i < n; ++i) {
        pr_shape s = shapes[i];
        s.child_offset += w.nChildren;
        if (reuse) w.shapes[firstShape] = s; else w.shapes.push_back(s);
        mat[i] = make_float4(s.friction, s.restitution, s.is_sensor ? 1.0f : 0.0f, s.density);
        filt[i] = static_cast<unsigned>(s.filter_category) | (static_cast<unsigned>(s.filter_mask) << 16);
        grp[i] = s.filter_group;
    }
    std::vector<float> rad(n_children); std::vector<unsigned> voff(n_children), cnt(n_children);
    for (uint32_t i = 0; i < n_children; ++i) {
        pr_child c = children[i];
        c.vertex_offset += w.nVerts;
        w.children.push_back(c);
        rad[i] = c.vertex_radius; voff[i] = c.vertex_offset; cnt[i] = c.vertex_count;
    }
    h2d(w, w.d.sh_mat + firstShape, mat.data(), n * sizeof(float4)); h2d(w, w.d.sh_filter + firstShape, filt.data(), n * 4);
    h2d(w, w.d.sh_group + firstShape, grp.data(), n * 4);
    h2d(w, w.d.ch_radius + w.nChildren, rad.data(), n_children * 4); h2d(w, w.d.ch_voff + w.nChildren, voff.data(), n_children * 4);
    h2d(w, w.d.ch_count + w.nChildren, cnt.data(), n_children * 4);
    std::vector<float4> pack(n_children);
    for (uint32_t i = 0; i < n_children; ++i) pack[i] = make_float4(rad[i], u2f(voff[i]), u2f(cnt[i]), 0.0f);
    h2d(w, w.d.ch_pack + w.nChildren, pack.data(), n_children * sizeof(float4));
    h2d(w, w.d.vtx + w.nVerts, vertices, n_vertices * sizeof(pr_vec2)); h2d(w, w.d.nrm + w.nVerts, normals, n_vertices * sizeof(pr_vec2));
    sync(w);
    if (first_id) *first_id = firstShape;
    if (!reuse) w.nShapes += n;
    w.nChildren += n_children; w.nVerts += n_vertices;
    return PR_OK;
    API_CATCH(pw)
}

static int create_bodies(pr_world* pw, uint32_t n, const pr_body* bodies, const uint32_t* shape_offsets, const uint32_t* shape_ids,
                         const uint32_t* world_index, uint32_t* first_id)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false;
    if (w.d.nWorlds) { // batch arena: every body names the world it lives in
        if (!world_index) { w.error = "batch arena: bodies are created with pr_batch_create_bodies (a world index per body)"; return PR_ERR_INVALID_ARGUMENT; }
        for (uint32_t i = 0; i < n; ++i) if (world_index[i] >= w.d.nWorlds) { w.error = "no such world"; return PR_ERR_OUT_OF_RANGE; }
    }
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t k = shape_offsets[i]; k < shape_offsets[i + 1]; ++k)
            if (shape_ids[k] >= w.nShapes) { w.error = "no such shape"; return PR_ERR_OUT_OF_RANGE; }
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t k = shape_offsets[i]; k < shape_offsets[i + 1]; ++k)
            if (shape_dead(w, shape_ids[k])) { w.error = "shape identifier is destroyed"; return PR_ERR_INVALID_ARGUMENT; }
    if (!w.pools.bodyFree.empty()) { // ids of destroyed bodies are handed out again first, last freed first (ObjectPool::Allocate)
        if (n != 1u) { w.error = "CreateBody: create bodies one at a time while destroyed body ids are free"; return PR_ERR_UNSUPPORTED; }
        const unsigned id = w.pools.bodyFree.back();
        w.pools.bodyFree.pop_back(); w.pools.bodyDead[id] = 0;
        const int rc = pr_set_bodies(pw, 1u, &id, bodies);
        if (rc != PR_OK) return rc;
        const unsigned z = 0u;
        h2d(w, w.d.b_moved + id, &z, 4); h2d(w, w.d.b_sync + id, &z, 4);
        if (w.d.nWorlds) { h2d(w, w.d.b_world + id, world_index, 4); w.bodyWorld[id] = world_index[0]; w.blocksDirty = true; }
        sync(w);
        std::vector<unsigned> sh(shape_ids + shape_offsets[0], shape_ids + shape_offsets[1]);
        if (bodies[0].flags & PR_BODY_ENABLED) for (unsigned s : sh) w.fixturesForProxies.emplace_back(id, s);
        if ((bodies[0].flags & PR_BODY_IMPENETRABLE) && !sh.empty()) w.toiPossible = true;
        w.bodyShapes[id] = std::move(sh);
        if (first_id) *first_id = id;
        w.adjValid = false; w.bpDirty = true; w.toiAlphaDirty = true; w.recolorAll = true;
        return PR_OK;
    }
    ensure(w, G_BODY, w.nBodies + n);
    pr_body* staging = static_cast<pr_body*>(dev_alloc(sizeof(pr_body) * n));
    h2d(w, staging, bodies, sizeof(pr_body) * n);
    PR_LAUNCH(k_put_bodies, n, w.stream, w.d, n, w.nBodies, static_cast<const unsigned*>(nullptr), staging, 0u);
    dset(w, w.d.b_moved + w.nBodies, 0, 4 * n); dset(w, w.d.b_sync + w.nBodies, 0, 4 * n);
    if (w.d.nWorlds) { h2d(w, w.d.b_world + w.nBodies, world_index, 4 * n); w.bodyWorld.insert(w.bodyWorld.end(), world_index, world_index + n); w.blocksDirty = true; }
    sync(w);
    dev_free(staging);
    for (uint32_t i = 0; i < n; ++i) {
        std::vector<unsigned> sh(shape_ids + shape_offsets[i], shape_ids + shape_offsets[i + 1]);
        if (bodies[i].flags & PR_BODY_ENABLED) for (unsigned s : sh) w.fixturesForProxies.emplace_back(w.nBodies + i, s); // CreateBody :1046-1052
        if ((bodies[i].flags & PR_BODY_IMPENETRABLE) && !sh.empty()) w.toiPossible = true;
        w.bodyShapes.push_back(std::move(sh));
    }
    if (first_id) *first_id = w.nBodies;
    w.nBodies += n;
    w.adjValid = false; w.bpDirty = true; w.toiAlphaDirty = true;
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_shape_material(pr_world* pw, uint32_t shape, float friction, float restitution, float density, uint32_t is_sensor)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (!w.stepComplete) { w.error = "SetShape: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (shape >= w.nShapes) { w.error = "no such shape"; return PR_ERR_OUT_OF_RANGE; }
    pr_shape& cur = w.shapes[shape];
    const bool touch = (cur.is_sensor != 0) != (is_sensor != 0) || cur.friction != friction || cur.restitution != restitution;
    std::vector<unsigned> users;
    for (unsigned b = 0; b < w.bodyShapes.size(); ++b)
        if (std::find(w.bodyShapes[b].begin(), w.bodyShapes[b].end(), shape) != w.bodyShapes[b].end()) users.push_back(b);
    if (!users.empty()) {
        unsigned* sids = static_cast<unsigned*>(stage_buf(w, 1, 4u * users.size()));
        h2d(w, sids, users.data(), 4u * users.size());
        PR_LAUNCH(k_wake_bodies, static_cast<unsigned>(users.size()), w.stream, w.d, static_cast<unsigned>(users.size()), sids);
    }
    if (touch && w.nOrder) PR_LAUNCH(k_shape_flag_contacts, w.nOrder, w.stream, w.d, w.nOrder, shape);
    cur.friction = friction; cur.restitution = restitution; cur.density = density; cur.is_sensor = is_sensor ? 1 : 0;
    const float4 mat = make_float4(friction, restitution, is_sensor ? 1.0f : 0.0f, density);
    h2d(w, w.d.sh_mat + shape, &mat, sizeof(float4));
    sync(w);
    return PR_OK;
    API_CATCH(pw)
}

// Joints that share a MOVABLE body cannot be solved by concurrent threads, and the device's joint order is id order. Level of a joint
// = 1 + the highest level among the earlier joints it shares a movable body with: joints of one level touch disjoint movable bodies
// (a body that cannot move is only ever written back unchanged), and launching the levels in turn keeps every conflicting pair in id
// order -- the same bits as one thread walking all joints in id order, which is what runs when the chains are longer than
// JOINT_MAX_LEVELS. Also: the body pairs joined by a collideConnected == false joint, sorted, for the broadphase's ShouldCollide.
// Destroyed joints (bodies == INVALID) are skipped.
// m_bodyJoints on the host: Add(JointID) appends to both bodies' lists (:1162-1169), Remove(JointID) erases (:1192-1212)
static void joint_links_add(World& w, unsigned jid, const pr_revolute_joint& j)
{
    auto& l = w.bodyJointList;
    const unsigned hi = std::max(j.body_a, j.body_b);
    if (l.size() <= hi) l.resize(hi + 1u);
    l[j.body_a].push_back(jid);
    if (j.body_b != j.body_a) l[j.body_b].push_back(jid); // (a target joint's row repeats its one body)
}
static void joint_links_remove(World& w, unsigned jid, const pr_revolute_joint& j)
{
    for (unsigned b : {j.body_a, j.body_b}) {
        if (b >= w.bodyJointList.size()) continue;
        auto& lst = w.bodyJointList[b];
        const auto it = std::find(lst.begin(), lst.end(), jid);
        if (it != lst.end()) lst.erase(it);
    }
}
constexpr unsigned JOINT_MAX_LEVELS = 48u;
static void refresh_joint_tables(World& w)
{
    w.blkJointsDirty = true; // (block solver: the joints of every world are listed again before the next solve)
    std::vector<unsigned long long> nc;
    const unsigned nJ = static_cast<unsigned>(w.joints.size());
    std::vector<unsigned> flags; std::vector<GearExt> gext;
    bool anyGear = false;
    for (const auto& j : w.joints) if (j.body_a != INVALID && PR_JOINT_TYPE(j) == PR_JOINT_GEAR) anyGear = true;
    if (nJ) { flags.resize(w.nBodies); d2h(w, flags.data(), w.d.b_flags, sizeof(unsigned) * w.nBodies); }
    if (anyGear) { gext.resize(nJ); d2h(w, gext.data(), w.d.jt_gear, sizeof(GearExt) * nJ); }
    std::unordered_map<unsigned, unsigned> last; // movable body -> level of the last joint on it
    std::vector<unsigned> level(nJ, 0u);
    unsigned nLevels = 0u;
    for (unsigned i = 0; i < nJ; ++i) {
        const auto& j = w.joints[i];
        if (j.body_a == INVALID) continue;
        unsigned bodies[4] = {j.body_a, j.body_b, INVALID, INVALID};
        if (PR_JOINT_TYPE(j) == PR_JOINT_GEAR) { bodies[2] = gext[i].body_c; bodies[3] = gext[i].body_d; }
        unsigned lvl = 0u;
        for (unsigned b : bodies) if (b != INVALID && b < w.nBodies && (flags[b] & PR_BODY_VELOCITY)) { const auto it = last.find(b); if (it != last.end()) lvl = std::max(lvl, it->second); }
        ++lvl;
        for (unsigned b : bodies) if (b != INVALID && b < w.nBodies && (flags[b] & PR_BODY_VELOCITY)) last[b] = lvl;
        level[i] = lvl; nLevels = std::max(nLevels, lvl);
        if (!(j.collide_connected & 1u) && j.body_a != j.body_b) nc.push_back((static_cast<unsigned long long>(std::min(j.body_a, j.body_b)) << 32) | std::max(j.body_a, j.body_b));
    }
    w.jointsSerial = nLevels > 1u;
    w.jointLevelOff.clear();
    if (nLevels > 1u && nLevels <= JOINT_MAX_LEVELS) { // the joint ids sorted by (level, id) + where each level starts
        std::vector<unsigned> order; order.reserve(nJ);
        w.jointLevelOff.assign(nLevels + 1u, 0u);
        for (unsigned i = 0; i < nJ; ++i) if (level[i]) ++w.jointLevelOff[level[i]];
        for (unsigned l = 1; l <= nLevels; ++l) w.jointLevelOff[l] += w.jointLevelOff[l - 1u];
        order.resize(w.jointLevelOff[nLevels]);
        std::vector<unsigned> fill(w.jointLevelOff.begin(), w.jointLevelOff.end() - 1);
        for (unsigned i = 0; i < nJ; ++i) if (level[i]) order[fill[level[i] - 1u]++] = i;
        ensure(w, G_JOINT, nJ + 1u);
        h2d(w, w.d.jt_order, order.data(), sizeof(unsigned) * order.size());
        sync(w);
    }
    std::sort(nc.begin(), nc.end());
    nc.erase(std::unique(nc.begin(), nc.end()), nc.end());
    ensure(w, G_JOINT, static_cast<unsigned>(std::max<size_t>(nc.size(), w.joints.size()) + 1));
    h2d(w, w.d.nc_pairs, nc.data(), sizeof(unsigned long long) * nc.size());
    w.d.nNcPairs = static_cast<unsigned>(nc.size());
}

int pr_attach_shapes(pr_world* pw, uint32_t body, uint32_t n, const uint32_t* shape_ids)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (!w.stepComplete) { w.error = "SetBody: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (body >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    for (uint32_t i = 0; i < n; ++i) if (shape_ids[i] >= w.nShapes) { w.error = "no such shape"; return PR_ERR_OUT_OF_RANGE; }
    if (!n) return PR_OK;
    unsigned flags = 0;
    d2h(w, &flags, w.d.b_flags + body, 4);
    if (!(flags & PR_BODY_ENABLED)) { // a disabled body has no proxies: the shapes only join its list (GetOldAndNewShapeIDs :677-680); enabling it queues them all
        for (uint32_t i = 0; i < n; ++i) w.bodyShapes[body].push_back(shape_ids[i]);
        return PR_OK;
    }
    for (uint32_t i = 0; i < n; ++i) { w.fixturesForProxies.emplace_back(body, shape_ids[i]); w.bodyShapes[body].push_back(shape_ids[i]); } // Append(m_fixturesForProxies, ...) :2786
    if (flags & PR_BODY_IMPENETRABLE) w.toiPossible = true;
    w.adjValid = false; w.bpDirty = true; w.toiAlphaDirty = true;
    return PR_OK;
    API_CATCH(pw)
}

int pr_create_bodies(pr_world* pw, uint32_t n, const pr_body* bodies, const uint32_t* shape_offsets, const uint32_t* shape_ids, uint32_t* first_id)
{
    return create_bodies(pw, n, bodies, shape_offsets, shape_ids, nullptr, first_id);
}

} // extern "C" (helpers of the structural mutators)
// DestroyLeaf for the proxy slots listed, in the order given (= the order the reference destroys the leaves in: their ids go
// onto the leaf free list in that order); the proxy table, the queue, the contacts' slot pairs and m_bodyProxies are renumbered.
// Contacts on these leaves were destroyed by the caller (the reference destroys them around the DestroyLeaf calls too).
static void destroy_proxy_slots(World& w, const std::vector<unsigned>& slots)
{
    if (slots.empty()) return;
    DV& d = w.d;
    const unsigned P = w.nProxies;
    std::vector<unsigned> map(P, 0u);
    for (unsigned s : slots) { map[s] = INVALID; w.pools.treeFree.push_back(w.pools.pxTree[s]); }
    unsigned k = 0;
    for (unsigned i = 0; i < P; ++i) if (map[i] != INVALID) map[i] = k++;
    // device: the columns only the device holds
    unsigned char* tmp = static_cast<unsigned char*>(stage_buf(w, 0, static_cast<size_t>(P) * 28u + 64u));
    unsigned* dmap = static_cast<unsigned*>(stage_buf(w, 1, static_cast<size_t>(P) * 4u + 16u));
    unsigned* tChild = reinterpret_cast<unsigned*>(tmp + static_cast<size_t>(P) * 16u); unsigned* tCg = tChild + P; unsigned* tMoved = tCg + P;
    float4* tFat = reinterpret_cast<float4*>(tmp);
    h2d(w, dmap, map.data(), 4u * P);
    PR_LAUNCH(k_px_compact, P, w.stream, d, P, dmap, tChild, tCg, tMoved, tFat);
    d2d(w, d.p_child, tChild, 4u * k); d2d(w, d.p_cg, tCg, 4u * k); d2d(w, d.p_moved, tMoved, 4u * k); d2d(w, d.p_fat, tFat, sizeof(float4) * k);
    if (w.nOrder) PR_LAUNCH(k_px_remap_contacts, w.nOrder, w.stream, d, w.nOrder, dmap);
    // the queue (m_proxiesForContacts): EraseFirst of the destroyed ones, the others renumbered
    read_counters(w);
    if (w.hostCnt.nQueue) {
        std::vector<unsigned> q(w.hostCnt.nQueue);
        d2h(w, q.data(), d.q_list, 4u * q.size());
        unsigned m = 0;
        for (unsigned e : q) if (map[e] != INVALID) q[m++] = map[e];
        h2d(w, d.q_list, q.data(), 4u * m);
        w.hostCnt.nQueue = m; write_counters(w);
    }
    { // host mirrors
        auto& pl = w.pools;
        for (unsigned i = 0; i < P; ++i) if (map[i] != INVALID) { pl.pxBody[map[i]] = pl.pxBody[i]; pl.pxShape[map[i]] = pl.pxShape[i]; pl.pxTree[map[i]] = pl.pxTree[i]; }
        pl.pxBody.resize(k); pl.pxShape.resize(k); pl.pxTree.resize(k);
        h2d(w, d.p_body, pl.pxBody.data(), 4u * k); h2d(w, d.p_shape, pl.pxShape.data(), 4u * k); h2d(w, d.p_tree, pl.pxTree.data(), 4u * k);
    }
    for (auto& list : w.bodyProxies) {
        unsigned m = 0;
        for (unsigned e : list) if (map[e] != INVALID) list[m++] = map[e];
        list.resize(m);
    }
    sync(w);
    w.nProxies = k; w.bvhValid = false; w.bpDirty = true; w.toiAlphaDirty = true;
}
// the SetBody whose value has lost shapes (AabbTreeWorld.cpp:2760-2785; Detach, Destroy(ShapeID), a body being disabled): the
// proxies of the (body, shape) pairs leave the tree in m_bodyProxies order, then per shape the contacts of that pair are
// destroyed and the pair leaves m_fixturesForProxies. shapeIds: ascending (the reference walks a std::map).
static void body_remove_shapes(World& w, unsigned body, const std::vector<unsigned>& shapeIds, bool keepInList)
{
    std::vector<unsigned> slots;
    if (body < w.bodyProxies.size())
        for (unsigned s : w.bodyProxies[body])
            if (std::find(shapeIds.begin(), shapeIds.end(), w.pools.pxShape[s]) != shapeIds.end()) slots.push_back(s);
    for (unsigned sid : shapeIds) {
        destroy_contacts_where(w, 1u, body, sid);
        auto& f = w.fixturesForProxies;
        f.erase(std::remove(f.begin(), f.end(), std::make_pair(body, sid)), f.end());
    }
    destroy_proxy_slots(w, slots);
    if (!keepInList) {
        auto& list = w.bodyShapes[body];
        for (unsigned sid : shapeIds) { const auto it = std::find(list.begin(), list.end(), sid); if (it != list.end()) list.erase(it); }
    }
}
extern "C" {

int pr_destroy_bodies(pr_world* pw, uint32_t n, const uint32_t* ids)
{ // Destroy(AabbTreeWorld&, BodyID), :1062-1106
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false;
    if (!w.stepComplete) { w.error = "Destroy: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    for (uint32_t i = 0; i < n; ++i) {
        if (ids[i] >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
        if (body_dead(w, ids[i])) { w.error = "body identifier is destroyed"; return PR_ERR_INVALID_ARGUMENT; }
        for (uint32_t k = 0; k < i; ++k) if (ids[k] == ids[i]) { w.error = "body listed twice"; return PR_ERR_INVALID_ARGUMENT; }
    }
    for (uint32_t i = 0; i < n; ++i) {
        const unsigned id = ids[i];
        // the attached joints (:1072-1085), in the order of the body's joint list (the order their ids are freed in)
        const std::vector<unsigned> js = id < w.bodyJointList.size() ? w.bodyJointList[id] : std::vector<unsigned>{};
        if (!js.empty()) { const int rc = pr_destroy_joints(pw, static_cast<uint32_t>(js.size()), js.data()); if (rc != PR_OK) return rc; }
        destroy_contacts_where(w, 0u, id, 0u); // the attached contacts (:1087-1091), in m_bodyContacts order
        { auto& f = w.fixturesForProxies; f.erase(std::remove_if(f.begin(), f.end(), [id](const std::pair<unsigned, unsigned>& e) { return e.first == id; }), f.end()); }
        std::vector<unsigned> slots = id < w.bodyProxies.size() ? w.bodyProxies[id] : std::vector<unsigned>{};
        std::sort(slots.begin(), slots.end(), [&w](unsigned a, unsigned b) { return w.pools.pxTree[a] < w.pools.pxTree[b]; }); // FindProxies: by leaf id (:617-630)
        destroy_proxy_slots(w, slots);
        // Remove(BodyID) (:1046-1060): the slot reads back destroyed and is the next id a CreateBody hands out
        pr_body zero{};
        zero.xf_qc = 1.0f; zero.flags = PR_BODY_DESTROYED;
        const int rc = pr_set_bodies(pw, 1u, &id, &zero);
        if (rc != PR_OK) return rc;
        const unsigned z = 0u;
        h2d(w, w.d.b_sync + id, &z, 4); h2d(w, w.d.b_moved + id, &z, 4);
        sync(w);
        w.bodyShapes[id].clear();
        if (w.pools.bodyDead.size() <= id) w.pools.bodyDead.resize(id + 1u, 0);
        w.pools.bodyDead[id] = 1; w.pools.bodyFree.push_back(id);
        w.recolorAll = true; w.blocksDirty = true;
    }
    return PR_OK;
    API_CATCH(pw)
}

int pr_detach_shapes(pr_world* pw, uint32_t body, uint32_t n, const uint32_t* shape_ids)
{ // the SetBody whose value has fewer shapes (:2760-2785); shape_ids: one entry per attachment to remove
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false;
    if (!w.stepComplete) { w.error = "SetBody: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (body >= w.nBodies || body_dead(w, body)) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    std::vector<unsigned> ids(shape_ids, shape_ids + n), have = w.bodyShapes[body];
    std::sort(ids.begin(), ids.end());
    for (unsigned sid : ids) { const auto it = std::find(have.begin(), have.end(), sid); if (it == have.end()) { w.error = "shape is not attached to the body"; return PR_ERR_INVALID_ARGUMENT; } have.erase(it); }
    if (!n) return PR_OK;
    if (body_flags(w, body) & PR_BODY_ENABLED) body_remove_shapes(w, body, ids, false);
    else w.bodyShapes[body] = have; // (a disabled body has no proxies and no contacts)
    return PR_OK;
    API_CATCH(pw)
}

int pr_destroy_shapes(pr_world* pw, uint32_t n, const uint32_t* ids)
{ // Destroy(AabbTreeWorld&, ShapeID), :1259-1277: detached from every body (in body order), then the id is free
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false;
    if (!w.stepComplete) { w.error = "Destroy: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    for (uint32_t i = 0; i < n; ++i) if (ids[i] >= w.nShapes) { w.error = "no such shape"; return PR_ERR_OUT_OF_RANGE; }
    for (uint32_t i = 0; i < n; ++i) {
        const unsigned sid = ids[i];
        if (shape_dead(w, sid)) continue; // (Free of a freed id: nothing)
        for (unsigned b = 0; b < w.bodyShapes.size(); ++b) {
            const unsigned cnt = static_cast<unsigned>(std::count(w.bodyShapes[b].begin(), w.bodyShapes[b].end(), sid));
            if (!cnt) continue;
            const std::vector<unsigned> rm(cnt, sid);
            const int rc = pr_detach_shapes(pw, b, cnt, rm.data());
            if (rc != PR_OK) return rc;
        }
        if (w.pools.shapeDead.size() <= sid) w.pools.shapeDead.resize(sid + 1u, 0);
        w.pools.shapeDead[sid] = 1; w.pools.shapeFree.push_back(sid);
        w.shapes[sid].child_count = 0;
    }
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_body_enabled(pr_world* pw, uint32_t body, uint32_t enabled)
{ // the SetBody whose value differs in IsEnabled (GetOldAndNewShapeIDs :671-676): all proxies and contacts go / all shapes are queued
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false;
    if (!w.stepComplete) { w.error = "SetBody: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (body >= w.nBodies || body_dead(w, body)) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    unsigned f = body_flags(w, body);
    const bool was = (f & PR_BODY_ENABLED) != 0u, now = enabled != 0u;
    if (was == now) return PR_OK;
    if (!now) {
        std::vector<unsigned> ids = w.bodyShapes[body];
        std::sort(ids.begin(), ids.end());
        body_remove_shapes(w, body, ids, true);
        f &= ~static_cast<unsigned>(PR_BODY_ENABLED);
    } else {
        for (unsigned sid : w.bodyShapes[body]) w.fixturesForProxies.emplace_back(body, sid);
        f |= PR_BODY_ENABLED;
        if ((f & PR_BODY_IMPENETRABLE) && !w.bodyShapes[body].empty()) w.toiPossible = true;
    }
    h2d(w, w.d.b_flags + body, &f, 4);
    sync(w);
    w.adjValid = false; w.bpDirty = true; w.toiAlphaDirty = true; w.recolorAll = true;
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_body_type(pr_world* pw, uint32_t body, const pr_body* row)
{ // the SetBody whose value has another BodyType (:2738-2758): the body's contacts are destroyed; a body turned static is
  // synchronised by the next Step (m_bodiesForSync); then "elem = value"
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false;
    if (!w.stepComplete) { w.error = "SetBody: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (body >= w.nBodies || body_dead(w, body)) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    destroy_contacts_where(w, 0u, body, 0u);
    const int rc = pr_set_bodies(pw, 1u, &body, row);
    if (rc != PR_OK) return rc;
    if (!(row->flags & (PR_BODY_VELOCITY | PR_BODY_ACCELERATION))) { const unsigned one = 1u; h2d(w, w.d.b_sync + body, &one, 4); sync(w); }
    w.recolorAll = true; w.adjValid = false;
    if (w.nJoints) w.jointTablesDirty = true; // (the joint levels depend on which bodies can move)
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_shape(pr_world* pw, uint32_t shape, const pr_shape* def, const pr_child* children, const pr_vec2* vertices, const pr_vec2* normals,
                 uint32_t n_vertices, uint32_t geometry_changed)
{ // SetShape(AabbTreeWorld&, ShapeID, Shape), :1284-1356, for any new value
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false;
    if (!w.stepComplete) { w.error = "SetShape: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (shape >= w.nShapes) { w.error = "no such shape"; return PR_ERR_OUT_OF_RANGE; }
    if (shape_dead(w, shape)) { w.error = "shape identifier is destroyed"; return PR_ERR_INVALID_ARGUMENT; }
    pr_shape cur = w.shapes[shape];
    if (geometry_changed) {
        for (uint32_t i = 0; i < def->child_count; ++i) {
            const pr_child& c = children[i];
            if (c.vertex_count == 0 || c.vertex_count > 254) { w.error = "child vertex count out of range"; return PR_ERR_INVALID_ARGUMENT; }
            if (!(c.vertex_radius >= w.conf.min_vertex_radius && c.vertex_radius <= w.conf.max_vertex_radius)) { w.error = "vertex radius outside world interval"; return PR_ERR_INVALID_ARGUMENT; }
            if (c.vertex_offset > n_vertices || c.vertex_count > n_vertices - c.vertex_offset) { w.error = "child vertex range"; return PR_ERR_OUT_OF_RANGE; }
        }
    }
    std::vector<unsigned> users;
    for (unsigned b = 0; b < w.bodyShapes.size(); ++b)
        if (std::find(w.bodyShapes[b].begin(), w.bodyShapes[b].end(), shape) != w.bodyShapes[b].end()) users.push_back(b);
    if (!users.empty()) { // SetAwake(b) for every body that has the shape (:1301)
        unsigned* sids = static_cast<unsigned*>(stage_buf(w, 1, 4u * users.size()));
        h2d(w, sids, users.data(), 4u * users.size());
        PR_LAUNCH(k_wake_bodies, static_cast<unsigned>(users.size()), w.stream, w.d, static_cast<unsigned>(users.size()), sids);
        sync(w);
    }
    if (geometry_changed) {
        for (unsigned b : users) { // :1302-1328: the enabled users lose the shape's proxies and contacts and are queued for new proxies
            if (!(body_flags(w, b) & PR_BODY_ENABLED)) continue;
            body_remove_shapes(w, b, std::vector<unsigned>{shape}, true);
            w.fixturesForProxies.emplace_back(b, shape);
        }
        // the new children and vertices go to the end of the pools (the old ones are dead storage)
        ensure(w, G_CHILD, w.nChildren + def->child_count); ensure(w, G_VERT, w.nVerts + n_vertices + 4u);
        std::vector<float> rad(def->child_count); std::vector<unsigned> voff(def->child_count), cnt(def->child_count);
        for (uint32_t i = 0; i < def->child_count; ++i) {
            pr_child c = children[i];
            c.vertex_offset += w.nVerts;
            w.children.push_back(c);
            rad[i] = c.vertex_radius; voff[i] = c.vertex_offset; cnt[i] = c.vertex_count;
        }
        h2d(w, w.d.ch_radius + w.nChildren, rad.data(), def->child_count * 4u); h2d(w, w.d.ch_voff + w.nChildren, voff.data(), def->child_count * 4u);
        h2d(w, w.d.ch_count + w.nChildren, cnt.data(), def->child_count * 4u);
        std::vector<float4> pack(def->child_count);
        for (uint32_t i = 0; i < def->child_count; ++i) pack[i] = make_float4(rad[i], u2f(voff[i]), u2f(cnt[i]), 0.0f);
        h2d(w, w.d.ch_pack + w.nChildren, pack.data(), def->child_count * sizeof(float4));
        h2d(w, w.d.vtx + w.nVerts, vertices, n_vertices * sizeof(pr_vec2)); h2d(w, w.d.nrm + w.nVerts, normals, n_vertices * sizeof(pr_vec2));
        cur.child_offset = w.nChildren; cur.child_count = def->child_count;
        w.nChildren += def->child_count; w.nVerts += n_vertices;
    }
    const bool filterChanged = cur.filter_category != def->filter_category || cur.filter_mask != def->filter_mask || cur.filter_group != def->filter_group;
    if (filterChanged) { // :1330-1345 (the shape's NEW filter is what the next DestroyContacts tests with)
        if (w.nOrder) {
            read_counters(w); w.hostCnt.nDestroyed = 0u; write_counters(w);
            PR_LAUNCH(k_shape_flag_filter, w.nOrder, w.stream, w.d, w.nOrder, shape);
            read_counters(w);
            if (w.hostCnt.nDestroyed) w.needsFiltering = true;
            w.hostCnt.nDestroyed = 0u; write_counters(w);
        }
        if (w.nProxies) { ensure(w, G_PROXY, w.nProxies + 1u); PR_LAUNCH(k_queue_shape_proxies, w.nProxies, w.stream, w.d, w.nProxies, shape); }
    }
    const bool touch = (cur.is_sensor != 0) != (def->is_sensor != 0) || cur.friction != def->friction || cur.restitution != def->restitution || geometry_changed;
    if (touch && w.nOrder) PR_LAUNCH(k_shape_flag_contacts, w.nOrder, w.stream, w.d, w.nOrder, shape); // :1346-1354
    cur.friction = def->friction; cur.restitution = def->restitution; cur.density = def->density; cur.is_sensor = def->is_sensor ? 1 : 0;
    cur.filter_category = def->filter_category; cur.filter_mask = def->filter_mask; cur.filter_group = def->filter_group;
    w.shapes[shape] = cur;
    const float4 mat = make_float4(cur.friction, cur.restitution, cur.is_sensor ? 1.0f : 0.0f, cur.density);
    const unsigned filt = static_cast<unsigned>(cur.filter_category) | (static_cast<unsigned>(cur.filter_mask) << 16);
    const int grp = cur.filter_group;
    h2d(w, w.d.sh_mat + shape, &mat, sizeof(float4)); h2d(w, w.d.sh_filter + shape, &filt, 4); h2d(w, w.d.sh_group + shape, &grp, 4);
    sync(w);
    w.adjValid = false;
    return PR_OK;
    API_CATCH(pw)
}

} // extern "C"
// FlagForFiltering (:436-451) for the body pairs of joints with collideConnected == false that are being added or removed
static void flag_joint_pairs_for_filtering(World& w, const std::vector<unsigned>& pairs)
{
    if (pairs.empty() || !w.nOrder) return;
    unsigned* dp = static_cast<unsigned*>(stage_buf(w, 1, 4u * pairs.size()));
    h2d(w, dp, pairs.data(), 4u * pairs.size());
    PR_LAUNCH(k_pair_flag_filter, w.nOrder, w.stream, w.d, w.nOrder, static_cast<unsigned>(pairs.size() / 2u), dp);
    sync(w);
    w.needsFiltering = true; // (the filter pass only looks at flagged contacts: harmless when none was)
}
extern "C" {
int pr_create_revolute_joints(pr_world* pw, uint32_t n, const pr_revolute_joint* joints, uint32_t* first_id)
{ // CreateJoint (:1132-1157) + Add(JointID, !collideConnected) (:1159-1175) for rows of the joint table; creating a joint does not wake its bodies
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (n == 0) return PR_OK;
    for (uint32_t i = 0; i < n; ++i)
        if (joints[i].body_a >= w.nBodies || joints[i].body_b >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    if (w.d.nWorlds)
        for (uint32_t i = 0; i < n; ++i)
            if (w.bodyWorld[joints[i].body_a] != w.bodyWorld[joints[i].body_b]) { w.error = "joint between bodies of different worlds"; return PR_ERR_INVALID_ARGUMENT; }
    if (!w.freeJoints.empty()) { // ids of destroyed joints are handed out again first, last freed first (ObjectPool::Allocate)
        if (n != 1u) { w.error = "CreateJoint: create joints one at a time while destroyed joint ids are free"; return PR_ERR_UNSUPPORTED; }
        const unsigned id = w.freeJoints.back();
        w.freeJoints.pop_back();
        w.joints[id] = joints[0];
        joint_links_add(w, id, joints[0]);
        h2d(w, w.d.jt + id, joints, sizeof(pr_revolute_joint));
        { JointWork fresh{}; fresh.island = INVALID; h2d(w, w.d.jt_work + id, &fresh, sizeof(fresh)); sync(w); } // (not solved yet: pr_get_joint_work)
        w.jointTablesDirty = true;
        if (!(joints[0].collide_connected & 1u)) flag_joint_pairs_for_filtering(w, {joints[0].body_a, joints[0].body_b});
        sync(w);
        if (first_id) *first_id = id;
        return PR_OK;
    }
    w.joints.insert(w.joints.end(), joints, joints + n);
    for (uint32_t i = 0; i < n; ++i) joint_links_add(w, w.nJoints + i, joints[i]);
    ensure(w, G_JOINT, w.nJoints + n + 1);
    h2d(w, w.d.jt + w.nJoints, joints, sizeof(pr_revolute_joint) * n);
    { JointWork one{}; one.island = INVALID; std::vector<JointWork> fresh(n, one); h2d(w, w.d.jt_work + w.nJoints, fresh.data(), sizeof(JointWork) * n); sync(w); } // (not solved yet: pr_get_joint_work)
    w.jointTablesDirty = true;
    { std::vector<unsigned> pairs; for (uint32_t i = 0; i < n; ++i) if (!(joints[i].collide_connected & 1u)) { pairs.push_back(joints[i].body_a); pairs.push_back(joints[i].body_b); }
      flag_joint_pairs_for_filtering(w, pairs); }
    sync(w);
    if (first_id) *first_id = w.nJoints;
    w.nJoints += n;
    return PR_OK;
    API_CATCH(pw)
}

int pr_create_distance_joints(pr_world* pw, uint32_t n, const pr_distance_joint* joints, uint32_t* first_id)
{ // CreateJoint for DistanceJointConf values: rows of the one joint table, tagged
    static_assert(sizeof(pr_distance_joint) == sizeof(pr_revolute_joint), "joint rows share one table");
    std::vector<pr_revolute_joint> rows(n);
    for (uint32_t i = 0; i < n; ++i) {
        std::memcpy(&rows[i], &joints[i], sizeof(pr_revolute_joint));
        rows[i].collide_connected = (joints[i].collide_connected & 1u) | (static_cast<uint32_t>(PR_JOINT_DISTANCE) << 8);
        rows[i].impulse_y = rows[i].impulse_z = rows[i].angular_motor_impulse = 0.0f;
        rows[i].enable_limit = rows[i].enable_motor = 0u; rows[i].motor_speed = rows[i].max_motor_torque = 0.0f; rows[i].limit_state = 0u;
    }
    return pr_create_revolute_joints(pw, n, rows.data(), first_id);
}

int pr_create_rope_joints(pr_world* pw, uint32_t n, const pr_rope_joint* joints, uint32_t* first_id)
{
    static_assert(sizeof(pr_rope_joint) == sizeof(pr_revolute_joint), "joint rows share one table");
    std::vector<pr_revolute_joint> rows(n);
    for (uint32_t i = 0; i < n; ++i) {
        std::memset(&rows[i], 0, sizeof(pr_revolute_joint));
        rows[i].body_a = joints[i].body_a; rows[i].body_b = joints[i].body_b;
        rows[i].collide_connected = (joints[i].collide_connected & 1u) | (static_cast<uint32_t>(PR_JOINT_ROPE) << 8);
        rows[i].local_anchor_a = joints[i].local_anchor_a; rows[i].local_anchor_b = joints[i].local_anchor_b;
        rows[i].impulse_x = joints[i].impulse; rows[i].reference_angle = joints[i].max_length; rows[i].limit_state = joints[i].limit_state;
    }
    return pr_create_revolute_joints(pw, n, rows.data(), first_id);
}

} // extern "C"
template <class Row> static int create_tagged_joints(pr_world* pw, uint32_t n, const Row* joints, uint32_t type, uint32_t* first_id)
{ // rows of the one joint table: same size and leading fields as pr_revolute_joint, tagged with their type
    static_assert(sizeof(Row) == sizeof(pr_revolute_joint), "joint rows share one table");
    std::vector<pr_revolute_joint> rows(n);
    for (uint32_t i = 0; i < n; ++i) {
        std::memcpy(&rows[i], &joints[i], sizeof(pr_revolute_joint));
        rows[i].collide_connected = (joints[i].collide_connected & (1u | PR_JOINT_ENABLE_LIMIT | PR_JOINT_ENABLE_MOTOR)) | (type << 8);
    }
    return pr_create_revolute_joints(pw, n, rows.data(), first_id);
}
extern "C" {
int pr_create_friction_joints(pr_world* pw, uint32_t n, const pr_friction_joint* joints, uint32_t* first_id) { return create_tagged_joints(pw, n, joints, PR_JOINT_FRICTION, first_id); }
int pr_create_motor_joints(pr_world* pw, uint32_t n, const pr_motor_joint* joints, uint32_t* first_id) { return create_tagged_joints(pw, n, joints, PR_JOINT_MOTOR, first_id); }
int pr_create_weld_joints(pr_world* pw, uint32_t n, const pr_weld_joint* joints, uint32_t* first_id) { return create_tagged_joints(pw, n, joints, PR_JOINT_WELD, first_id); }
int pr_create_prismatic_joints(pr_world* pw, uint32_t n, const pr_prismatic_joint* joints, uint32_t* first_id) { return create_tagged_joints(pw, n, joints, PR_JOINT_PRISMATIC, first_id); }
int pr_create_wheel_joints(pr_world* pw, uint32_t n, const pr_wheel_joint* joints, uint32_t* first_id) { return create_tagged_joints(pw, n, joints, PR_JOINT_WHEEL, first_id); }
int pr_create_pulley_joints(pr_world* pw, uint32_t n, const pr_pulley_joint* joints, uint32_t* first_id) { return create_tagged_joints(pw, n, joints, PR_JOINT_PULLEY, first_id); }
static GearExt gear_ext_of(const pr_gear_joint& g)
{
    GearExt e{};
    e.body_c = g.body_c; e.body_d = g.body_d; e.type_ac = g.ac.type; e.type_bd = g.bd.type;
    const pr_gear_side* sides[2] = {&g.ac, &g.bd};
    float* dst[2] = {e.ac, e.bd};
    for (int k = 0; k < 2; ++k) {
        if (sides[k]->type == PR_GEAR_PRISMATIC) {
            dst[k][0] = sides[k]->local_anchor_a.x; dst[k][1] = sides[k]->local_anchor_a.y; dst[k][2] = sides[k]->local_anchor_b.x; dst[k][3] = sides[k]->local_anchor_b.y;
            dst[k][4] = sides[k]->local_axis.x; dst[k][5] = sides[k]->local_axis.y;
        } else if (sides[k]->type == PR_GEAR_REVOLUTE) dst[k][0] = sides[k]->reference_angle;
    }
    return e;
}
int pr_create_gear_joints(pr_world* pw, uint32_t n, const pr_gear_joint* joints, uint32_t* first_id)
{ // CreateJoint of GearJointConf values: a row of the joint table + the GearExt entry of the same id
    API_TRY(pw)
    World& w = pw->w;
    std::vector<pr_revolute_joint> rows(n);
    for (uint32_t i = 0; i < n; ++i) {
        const pr_gear_joint& g = joints[i];
        if (g.body_c >= w.nBodies || g.body_d >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
        if (g.ac.type > PR_GEAR_REVOLUTE || g.bd.type > PR_GEAR_REVOLUTE) { w.error = "gear joint: unknown geared joint type"; return PR_ERR_INVALID_ARGUMENT; }
        if (w.d.nWorlds && g.body_a < w.nBodies && (w.bodyWorld[g.body_c] != w.bodyWorld[g.body_a] || w.bodyWorld[g.body_d] != w.bodyWorld[g.body_a])) {
            w.error = "joint between bodies of different worlds"; return PR_ERR_INVALID_ARGUMENT;
        }
        pr_revolute_joint r{};
        r.body_a = g.body_a; r.body_b = g.body_b;
        r.collide_connected = (g.collide_connected & 1u) | (static_cast<uint32_t>(PR_JOINT_GEAR) << 8);
        r.impulse_x = g.impulse; r.lower_angle = g.ratio; r.upper_angle = g.constant;
        rows[i] = r;
    }
    uint32_t first = 0;
    const int rc = pr_create_revolute_joints(pw, n, rows.data(), &first);
    if (rc != PR_OK) return rc;
    std::vector<GearExt> ext(n);
    for (uint32_t i = 0; i < n; ++i) ext[i] = gear_ext_of(joints[i]);
    if (n) { h2d(w, w.d.jt_gear + first, ext.data(), sizeof(GearExt) * n); sync(w); }
    if (first_id) *first_id = first;
    return PR_OK;
    API_CATCH(pw)
}
int pr_get_gear_joints(const pr_world* cpw, uint32_t first, uint32_t n, pr_gear_joint* out)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    if (first > w.nJoints || n > w.nJoints - first) { w.error = "no such joint"; return PR_ERR_OUT_OF_RANGE; }
    if (n == 0) return PR_OK;
    std::vector<pr_revolute_joint> rows(n); std::vector<GearExt> ext(n);
    d2h(w, rows.data(), w.d.jt + first, sizeof(pr_revolute_joint) * n);
    d2h(w, ext.data(), w.d.jt_gear + first, sizeof(GearExt) * n);
    sync(w);
    for (uint32_t i = 0; i < n; ++i) {
        pr_gear_joint g{};
        const pr_revolute_joint& r = rows[i];
        if (r.body_a == INVALID || PR_JOINT_TYPE(r) != PR_JOINT_GEAR) { g.body_a = g.body_b = g.body_c = g.body_d = INVALID; out[i] = g; continue; }
        g.body_a = r.body_a; g.body_b = r.body_b; g.collide_connected = r.collide_connected;
        g.body_c = ext[i].body_c; g.body_d = ext[i].body_d; g.impulse = r.impulse_x; g.ratio = r.lower_angle; g.constant = r.upper_angle;
        pr_gear_side* sides[2] = {&g.ac, &g.bd};
        const float* src[2] = {ext[i].ac, ext[i].bd};
        const unsigned types[2] = {ext[i].type_ac, ext[i].type_bd};
        for (int k = 0; k < 2; ++k) {
            sides[k]->type = types[k];
            if (types[k] == PR_GEAR_PRISMATIC) {
                sides[k]->local_anchor_a = pr_vec2{src[k][0], src[k][1]}; sides[k]->local_anchor_b = pr_vec2{src[k][2], src[k][3]}; sides[k]->local_axis = pr_vec2{src[k][4], src[k][5]};
            } else if (types[k] == PR_GEAR_REVOLUTE) sides[k]->reference_angle = src[k][0];
        }
        out[i] = g;
    }
    return PR_OK;
    API_CATCH(pw)
}
int pr_get_joint_work(const pr_world* cpw, uint32_t first, uint32_t n, pr_joint_work* out)
{
    static_assert(sizeof(pr_joint_work) == sizeof(JointWork), "pr_joint_work mirrors JointWork");
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    if (first > w.nJoints || n > w.nJoints - first) { w.error = "no such joint"; return PR_ERR_OUT_OF_RANGE; }
    if (n) d2h(w, out, w.d.jt_work + first, sizeof(JointWork) * n);
    return PR_OK;
    API_CATCH(pw)
}
int pr_create_target_joints(pr_world* pw, uint32_t n, const pr_target_joint* joints, uint32_t* first_id) { return create_tagged_joints(pw, n, joints, PR_JOINT_TARGET, first_id); }

int pr_destroy_joints(pr_world* pw, uint32_t n, const uint32_t* ids)
{ // Destroy(AabbTreeWorld&, JointID), :1215-1227: Remove(id) wakes both bodies (:1192-1205), the id goes to the free list
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (!w.stepComplete) { w.error = "Destroy: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    for (uint32_t i = 0; i < n; ++i) if (ids[i] >= w.nJoints) { w.error = "no such joint"; return PR_ERR_OUT_OF_RANGE; }
    std::vector<unsigned> wake, pairs;
    for (uint32_t i = 0; i < n; ++i) {
        pr_revolute_joint& j = w.joints[ids[i]];
        if (j.body_a == INVALID) continue; // already destroyed: not in m_joints, nothing happens (:1222)
        if (!(j.collide_connected & 1u)) { pairs.push_back(j.body_a); pairs.push_back(j.body_b); } // Remove(JointID), :1184-1190
        wake.push_back(j.body_a); wake.push_back(j.body_b);
        joint_links_remove(w, ids[i], j);
        std::memset(&j, 0, sizeof(j));
        j.body_a = j.body_b = INVALID;
        h2d(w, w.d.jt + ids[i], &j, sizeof(j));
        w.freeJoints.push_back(ids[i]);
    }
    if (!wake.empty()) {
        unsigned* sids = static_cast<unsigned*>(stage_buf(w, 1, 4u * wake.size()));
        h2d(w, sids, wake.data(), 4u * wake.size());
        PR_LAUNCH(k_wake_bodies, static_cast<unsigned>(wake.size()), w.stream, w.d, static_cast<unsigned>(wake.size()), sids);
    }
    flag_joint_pairs_for_filtering(w, pairs);
    w.jointTablesDirty = true;
    sync(w);
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_revolute_joints(pr_world* pw, uint32_t n, const uint32_t* ids, const pr_revolute_joint* rows)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (!w.stepComplete) { w.error = "SetJoint: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    bool structural = false; // bodies, collideConnected or type of some joint change: the joint tables (pairs that must not collide, levels) are rebuilt
    for (uint32_t i = 0; i < n; ++i) {
        if (ids[i] >= w.nJoints) { w.error = "no such joint"; return PR_ERR_OUT_OF_RANGE; }
        for (uint32_t k = 0; k < i; ++k) if (ids[k] == ids[i]) { w.error = "SetJoint: a joint id is named twice"; return PR_ERR_INVALID_ARGUMENT; }
        const pr_revolute_joint& cur = w.joints[ids[i]];
        if (cur.body_a == INVALID) { w.error = "joint identifier is destroyed"; return PR_ERR_INVALID_ARGUMENT; }
        if (rows[i].body_a >= w.nBodies || rows[i].body_b >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
        if (body_dead(w, rows[i].body_a) || body_dead(w, rows[i].body_b)) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
        const bool typeChange = PR_JOINT_TYPE(cur) != PR_JOINT_TYPE(rows[i]);
        if (typeChange && (PR_JOINT_TYPE(cur) == PR_JOINT_GEAR || PR_JOINT_TYPE(rows[i]) == PR_JOINT_GEAR)) {
            w.error = "SetJoint: a gear joint's value is wider than a table row (pr_create_gear_joints)"; return PR_ERR_UNSUPPORTED;
        }
        if (w.d.nWorlds && w.bodyWorld[rows[i].body_a] != w.bodyWorld[rows[i].body_b]) { w.error = "joint between bodies of different worlds"; return PR_ERR_INVALID_ARGUMENT; }
        if (typeChange || cur.body_a != rows[i].body_a || cur.body_b != rows[i].body_b || (cur.collide_connected & 1u) != (rows[i].collide_connected & 1u)) structural = true;
    }
    if (!n) return PR_OK;
    // Remove(id) flags the OLD pair's contacts and wakes the OLD bodies, Add(id, !collideConnected) flags the NEW pair's (:1127-1129)
    std::vector<unsigned> pairs;
    for (uint32_t i = 0; i < n; ++i) {
        const pr_revolute_joint& cur = w.joints[ids[i]];
        if (!(cur.collide_connected & 1u)) { pairs.push_back(cur.body_a); pairs.push_back(cur.body_b); }
        if (!(rows[i].collide_connected & 1u)) { pairs.push_back(rows[i].body_a); pairs.push_back(rows[i].body_b); }
    }
    pr_revolute_joint* srows = static_cast<pr_revolute_joint*>(stage_buf(w, 0, sizeof(pr_revolute_joint) * n));
    unsigned* sids = static_cast<unsigned*>(stage_buf(w, 1, 4u * n));
    h2d(w, srows, rows, sizeof(pr_revolute_joint) * n); h2d(w, sids, ids, 4u * n);
    PR_LAUNCH(k_set_joints, n, w.stream, w.d, n, sids, srows);
    sync(w);
    flag_joint_pairs_for_filtering(w, pairs);
    if (structural) w.jointTablesDirty = true;
    for (uint32_t i = 0; i < n; ++i) { joint_links_remove(w, ids[i], w.joints[ids[i]]); w.joints[ids[i]] = rows[i]; joint_links_add(w, ids[i], rows[i]); } // (Remove + Add: to the end of the lists)
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_bodies(pr_world* pw, uint32_t n, const uint32_t* ids, const pr_body* rows)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    uint32_t maxId = 0; // (branch-free so that it vectorises: a bulk SetBody of 1M rows spent ~1 ms in an early-exit loop here)
    for (uint32_t i = 0; i < n; ++i) maxId = ids[i] > maxId ? ids[i] : maxId;
    if (n && maxId >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    if (!n) return PR_OK;
    pr_body* staging = static_cast<pr_body*>(stage_buf(w, 0, sizeof(pr_body) * n));
    unsigned* dids = static_cast<unsigned*>(stage_buf(w, 1, 4 * n));
    h2d(w, staging, rows, sizeof(pr_body) * n); h2d(w, dids, ids, 4 * n);
    PR_LAUNCH(k_put_bodies, n, w.stream, w.d, n, 0u, dids, staging, 1u);
    sync(w); // the caller's rows / ids may be reused once we return
    w.anyBodySet = true; w.toiAlphaDirty = true;
    if (!w.toiPossible)
        for (uint32_t i = 0; i < n; ++i)
            if ((rows[i].flags & PR_BODY_IMPENETRABLE) && ids[i] < w.bodyShapes.size() && !w.bodyShapes[ids[i]].empty()) { w.toiPossible = true; break; }
    return PR_OK;
    API_CATCH(pw)
}

int pr_get_bodies(const pr_world* cpw, uint32_t first, uint32_t n, pr_body* rows)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    if (first > w.nBodies || n > w.nBodies - first) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    if (!n) return PR_OK;
    pr_body* staging = static_cast<pr_body*>(stage_buf(w, 0, sizeof(pr_body) * n));
    PR_LAUNCH(k_get_bodies, n, w.stream, w.d, first, n, staging);
    d2h(w, rows, staging, sizeof(pr_body) * n);
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_body_accelerations(pr_world* pw, uint32_t n, const uint32_t* ids, const float* acc3)
{
    API_TRY(pw)
    World& w = pw->w;
    uint32_t maxId = n ? n - 1u : 0u; // ids == NULL: bodies [0, n)
    if (ids) { maxId = 0; for (uint32_t i = 0; i < n; ++i) maxId = ids[i] > maxId ? ids[i] : maxId; }
    if (n && maxId >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    if (!n) return PR_OK;
    float* sacc = static_cast<float*>(stage_buf(w, 0, sizeof(float) * 3u * n));
    unsigned* dids = ids ? static_cast<unsigned*>(stage_buf(w, 1, 4u * n)) : nullptr;
    h2d(w, sacc, acc3, sizeof(float) * 3u * n);
    if (ids) h2d(w, dids, ids, 4u * n);
    PR_LAUNCH(k_put_accelerations, n, w.stream, w.d, n, dids, sacc);
    sync(w); // the caller's buffers may be reused once we return
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_body_velocities(pr_world* pw, uint32_t n, const uint32_t* ids, const float* vel3)
{
    API_TRY(pw)
    World& w = pw->w;
    uint32_t maxId = n ? n - 1u : 0u; // ids == NULL: bodies [0, n)
    if (ids) { maxId = 0; for (uint32_t i = 0; i < n; ++i) maxId = ids[i] > maxId ? ids[i] : maxId; }
    if (n && maxId >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    if (!n) return PR_OK;
    float* svel = static_cast<float*>(stage_buf(w, 0, sizeof(float) * 3u * n));
    unsigned* dids = ids ? static_cast<unsigned*>(stage_buf(w, 1, 4u * n)) : nullptr;
    h2d(w, svel, vel3, sizeof(float) * 3u * n);
    if (ids) h2d(w, dids, ids, 4u * n);
    PR_LAUNCH(k_put_velocities, n, w.stream, w.d, n, dids, svel);
    sync(w); // the caller's buffers may be reused once we return
    w.anyBodySet = true;
    return PR_OK;
    API_CATCH(pw)
}

int pr_get_body_poses(const pr_world* cpw, uint32_t first, uint32_t n, float* xf4, float* vel3)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    if (first > w.nBodies || n > w.nBodies - first) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
    if (!n) return PR_OK;
    float4* sxf = static_cast<float4*>(stage_buf(w, 0, sizeof(float4) * n));
    float* svel = vel3 ? static_cast<float*>(stage_buf(w, 1, sizeof(float) * 3u * n)) : nullptr;
    PR_LAUNCH(k_get_poses, n, w.stream, w.d, first, n, sxf, svel);
    if (vel3) CK_COPY_D2H_ASYNC(w, vel3, svel, sizeof(float) * 3u * n);
    d2h(w, xf4, sxf, sizeof(float4) * n);
    return PR_OK;
    API_CATCH(pw)
}

int pr_query_aabb(const pr_world* cpw, uint32_t n_boxes, const float* boxes4, const uint32_t* box_world, uint32_t cap, uint32_t* hit_box,
                  uint32_t* hit_body, uint32_t* hit_shape, uint32_t* hit_child, uint32_t* n_out)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    *n_out = 0;
    if (w.d.nWorlds && !box_world) { w.error = "pr_query_aabb on a batch arena needs the world index of every box"; return PR_ERR_INVALID_ARGUMENT; }
    const unsigned P = w.nProxies;
    if (!n_boxes || !P) return PR_OK;
    if (P > 1 && (!w.bvhValid || w.bvhLeaves != P)) { read_counters(w); rebuild_bvh(w); }
    float4* sbox = static_cast<float4*>(stage_buf(w, 0, sizeof(float4) * n_boxes + 4u * n_boxes));
    unsigned* sworld = box_world ? reinterpret_cast<unsigned*>(sbox + n_boxes) : nullptr;
    unsigned* sout = static_cast<unsigned*>(stage_buf(w, 1, 16u * static_cast<size_t>(cap) + 16u));
    unsigned* scount = sout + 4u * static_cast<size_t>(cap);
    h2d(w, sbox, boxes4, sizeof(float4) * n_boxes);
    if (box_world) h2d(w, sworld, box_world, 4u * n_boxes);
    dset(w, scount, 0, 4);
#ifndef PR_HOST_EMUL
    PR_LAUNCH(k_query_boxes, static_cast<unsigned long long>(n_boxes) * QG, w.stream, w.d, n_boxes, sbox, sworld, P, w.bvhSortedIdx, cap, sout, scount);
#else
    PR_LAUNCH(k_query_boxes, n_boxes, w.stream, w.d, n_boxes, sbox, sworld, P, w.bvhSortedIdx, cap, sout, scount);
#endif
    unsigned total = 0;
    d2h(w, &total, scount, 4);
    const unsigned got = std::min(total, cap);
    if (got) {
        std::vector<unsigned> tmp(4u * static_cast<size_t>(got));
        d2h(w, tmp.data(), sout, 16u * static_cast<size_t>(got));
        for (unsigned i = 0; i < got; ++i) { hit_box[i] = tmp[4u * i]; hit_body[i] = tmp[4u * i + 1u]; hit_shape[i] = tmp[4u * i + 2u]; hit_child[i] = tmp[4u * i + 3u]; }
    }
    *n_out = total; // > cap: the caller's arrays were too short, the first `cap` hits were written
    return PR_OK;
    API_CATCH(pw)
}

int pr_test_points(const pr_world* cpw, uint32_t n, const uint32_t* body_ids, const uint32_t* shape_ids, const float* points_xy, uint32_t* out)
{ // TestPoint(const World&, BodyID, ShapeID, const Length2&), WorldShape.cpp:178-181, for many (body, shape, point) at once
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    if (!n) return PR_OK;
    std::vector<uint4> q(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (body_ids[i] >= w.nBodies) { w.error = "no such body"; return PR_ERR_OUT_OF_RANGE; }
        if (shape_ids[i] >= w.nShapes) { w.error = "no such shape"; return PR_ERR_OUT_OF_RANGE; }
        const pr_shape& sh = w.shapes[shape_ids[i]];
        q[i] = make_uint4(body_ids[i], sh.child_offset, sh.child_count, 0u);
    }
    unsigned char* in = static_cast<unsigned char*>(stage_buf(w, 0, (sizeof(uint4) + sizeof(float2)) * static_cast<size_t>(n)));
    unsigned* sout = static_cast<unsigned*>(stage_buf(w, 1, sizeof(unsigned) * static_cast<size_t>(n)));
    uint4* sq = reinterpret_cast<uint4*>(in);
    float2* sp = reinterpret_cast<float2*>(in + sizeof(uint4) * static_cast<size_t>(n));
    h2d(w, sq, q.data(), sizeof(uint4) * n); h2d(w, sp, points_xy, sizeof(float2) * n);
    PR_LAUNCH(k_test_points, n, w.stream, w.d, n, sq, sp, sout);
    d2h(w, out, sout, sizeof(unsigned) * n);
    return PR_OK;
    API_CATCH(pw)
}

int pr_find_closest_bodies(const pr_world* cpw, uint32_t n, const float* points_xy, uint32_t* out_body_ids)
{ // FindClosestBody(const World&, const Length2&), WorldBody.cpp:576-592, for many locations at once (0xFFFFFFFF: no body)
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    if (!n) return PR_OK;
    float2* sp = static_cast<float2*>(stage_buf(w, 0, sizeof(float2) * static_cast<size_t>(n)));
    unsigned* sout = static_cast<unsigned*>(stage_buf(w, 1, sizeof(unsigned) * static_cast<size_t>(n)));
    h2d(w, sp, points_xy, sizeof(float2) * n);
#ifndef PR_HOST_EMUL
    PR_LAUNCH_GRID(k_find_closest_bodies, n, PR_BLOCK, w.stream, w.d, w.nBodies, sp, sout);
#else
    PR_LAUNCH(k_find_closest_bodies, n, w.stream, w.d, w.nBodies, sp, sout);
#endif
    d2h(w, out_body_ids, sout, sizeof(unsigned) * n);
    return PR_OK;
    API_CATCH(pw)
}

static int set_contact_error(World& w, unsigned long long e)
{
    const unsigned code = static_cast<unsigned>(e & 0xFFFFFFFFull);
    static const char* msgs[] = {"", "no such contact", "contact identifier is destroyed", "cannot change is-destroyed value", "cannot change contactable A",
                                 "cannot change contactable B", "change body A or B being impenetrable to change impenetrable state",
                                 "change shape A or B being a sensor to change sensor state", "user may not change the TOI",
                                 "user may not change the TOI count", "cannot change manifold type", "cannot change manifold point count"};
    w.error = std::string(msgs[code < 12u ? code : 0u]) + " (row " + std::to_string(static_cast<unsigned>(e >> 32)) + ")";
    return code == SETC_NO_SUCH ? PR_ERR_OUT_OF_RANGE : PR_ERR_INVALID_ARGUMENT;
}

int pr_set_contacts(pr_world* pw, uint32_t n, const pr_contact* rows)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (!w.stepComplete) { w.error = "SetContact: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (!n) return PR_OK;
    pr_contact* srows = static_cast<pr_contact*>(stage_buf(w, 0, sizeof(pr_contact) * n));
    unsigned long long* serr = static_cast<unsigned long long*>(stage_buf(w, 1, 8));
    h2d(w, srows, rows, sizeof(pr_contact) * n);
    dset(w, serr, 0xFF, 8);
    PR_LAUNCH(k_set_contacts, n, w.stream, w.d, n, srows, w.poolSize, 0u, serr);
    unsigned long long e = 0;
    d2h(w, &e, serr, 8);
    if (e != ~0ull) return set_contact_error(w, e);
    PR_LAUNCH(k_set_contacts, n, w.stream, w.d, n, srows, w.poolSize, 1u, serr);
    sync(w);
    w.adjValid = false; // the enabled flag decides TOI candidacy (toi_static_eligible): rebuild the TOI list
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_manifolds(pr_world* pw, uint32_t n, const uint32_t* ids, const pr_manifold* rows)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (!w.stepComplete) { w.error = "SetManifold: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    if (!n) return PR_OK;
    pr_manifold* srows = static_cast<pr_manifold*>(stage_buf(w, 0, sizeof(pr_manifold) * n + 4u * n));
    unsigned* sids = reinterpret_cast<unsigned*>(srows + n);
    unsigned long long* serr = static_cast<unsigned long long*>(stage_buf(w, 1, 8));
    h2d(w, srows, rows, sizeof(pr_manifold) * n); h2d(w, sids, ids, 4u * n);
    dset(w, serr, 0xFF, 8);
    PR_LAUNCH(k_set_manifolds, n, w.stream, w.d, n, sids, srows, w.poolSize, 0u, serr);
    unsigned long long e = 0;
    d2h(w, &e, serr, 8);
    if (e != ~0ull) return set_contact_error(w, e);
    PR_LAUNCH(k_set_manifolds, n, w.stream, w.d, n, sids, srows, w.poolSize, 1u, serr);
    sync(w);
    return PR_OK;
    API_CATCH(pw)
}

int pr_ray_cast(const pr_world* cpw, uint32_t n_rays, const float* rays5, const uint32_t* ray_world, uint32_t mode, uint32_t cap,
                pr_ray_hit* hits, uint32_t* n_out)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    *n_out = 0;
    if (mode != PR_RAY_ALL && mode != PR_RAY_CLOSEST) { w.error = "pr_ray_cast: unknown mode"; return PR_ERR_INVALID_ARGUMENT; }
    if (w.d.nWorlds && !ray_world) { w.error = "pr_ray_cast on a batch arena needs the world index of every ray"; return PR_ERR_INVALID_ARGUMENT; }
    if (mode == PR_RAY_CLOSEST && cap < n_rays) { w.error = "pr_ray_cast: PR_RAY_CLOSEST writes one hit per ray (cap < n_rays)"; return PR_ERR_INVALID_ARGUMENT; }
    for (uint32_t i = 0; i < n_rays; ++i) // RayCastInput::maxFraction is a UnitInterval<Real>
        if (!(rays5[5u * i + 4u] >= 0.0f && rays5[5u * i + 4u] <= 1.0f)) { w.error = "pr_ray_cast: maxFraction outside [0, 1]"; return PR_ERR_INVALID_ARGUMENT; }
    const unsigned P = w.nProxies;
    if (!n_rays) return PR_OK;
    if (!P) {
        if (mode == PR_RAY_CLOSEST) { for (uint32_t i = 0; i < n_rays; ++i) { std::memset(&hits[i], 0, sizeof(pr_ray_hit)); hits[i].ray = i; hits[i].body = hits[i].shape = hits[i].child = INVALID; } *n_out = n_rays; }
        return PR_OK;
    }
    if (P > 1 && (!w.bvhValid || w.bvhLeaves != P)) { read_counters(w); rebuild_bvh(w); }
    float* srays = static_cast<float*>(stage_buf(w, 0, 20u * static_cast<size_t>(n_rays) + 4u * n_rays));
    unsigned* sworld = ray_world ? reinterpret_cast<unsigned*>(srays + 5u * static_cast<size_t>(n_rays)) : nullptr;
    const unsigned outCap = mode == PR_RAY_CLOSEST ? n_rays : cap;
    pr_ray_hit* sout = static_cast<pr_ray_hit*>(stage_buf(w, 1, sizeof(pr_ray_hit) * static_cast<size_t>(outCap) + 16u));
    unsigned* scount = reinterpret_cast<unsigned*>(sout + outCap);
    h2d(w, srays, rays5, 20u * static_cast<size_t>(n_rays));
    if (ray_world) h2d(w, sworld, ray_world, 4u * n_rays);
    dset(w, scount, 0, 4);
    PR_LAUNCH(k_ray_cast, n_rays, w.stream, w.d, n_rays, srays, sworld, P, w.bvhSortedIdx, mode, outCap, sout, scount);
    unsigned total = n_rays;
    if (mode == PR_RAY_ALL) d2h(w, &total, scount, 4);
    const unsigned got = std::min(total, outCap);
    if (got) d2h(w, hits, sout, sizeof(pr_ray_hit) * static_cast<size_t>(got));
    *n_out = total; // PR_RAY_ALL and > cap: the caller's array was too short, the first `cap` hits were written
    read_counters(w);
    throw_device_error(w);
    return PR_OK;
    API_CATCH(pw)
}

int pr_get_revolute_joints(const pr_world* cpw, uint32_t first, uint32_t n, pr_revolute_joint* rows)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    if (first > w.nJoints || n > w.nJoints - first) { w.error = "no such joint"; return PR_ERR_OUT_OF_RANGE; }
    d2h(w, rows, w.d.jt + first, sizeof(pr_revolute_joint) * n);
    return PR_OK;
    API_CATCH(pw)
}

int pr_get_counts(const pr_world* pw, pr_counts* out)
{
    const World& w = pw->w;
    std::memset(out, 0, sizeof(*out));
    out->bodies = w.nBodies; out->shapes = w.nShapes; out->children = w.nChildren; out->vertices = w.nVerts;
    out->proxies = w.nProxies; out->contacts = w.nOrder; out->joints = w.nJoints; out->contact_id_range = w.poolSize;
    out->tree_leaf_count = w.nProxies;
    // DynamicTree capacity: DefaultTreeCapacity 4096 grown to the next power of two of the node count (DynamicTree.cpp:436-454)
    const unsigned nodes = w.nProxies ? 2u * w.nProxies - 1u : 0u;
    out->tree_node_capacity = std::max(4096u, next_pow2(nodes));
    return PR_OK;
}

int pr_get_contacts(const pr_world* cpw, uint32_t cap, pr_contact* contacts, pr_manifold* manifolds, uint32_t* n_out)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    *n_out = w.nOrder;
    const unsigned n = std::min(cap, w.nOrder);
    if (!n) return PR_OK;
    pr_contact* sc = static_cast<pr_contact*>(dev_alloc(sizeof(pr_contact) * n));
    pr_manifold* sm = manifolds ? static_cast<pr_manifold*>(dev_alloc(sizeof(pr_manifold) * n)) : nullptr;
    PR_LAUNCH(k_export_contacts, n, w.stream, w.d, n, sc, sm);
    d2h(w, contacts, sc, sizeof(pr_contact) * n);
    if (manifolds) d2h(w, manifolds, sm, sizeof(pr_manifold) * n);
    dev_free(sc); dev_free(sm);
    return PR_OK;
    API_CATCH(pw)
}

int pr_get_proxies(const pr_world* cpw, uint32_t cap, pr_proxy* proxies, uint32_t* n_out)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    *n_out = w.nProxies;
    const unsigned n = std::min(cap, w.nProxies);
    if (!n) return PR_OK;
    pr_proxy* sp = static_cast<pr_proxy*>(dev_alloc(sizeof(pr_proxy) * n));
    PR_LAUNCH(k_export_proxies, n, w.stream, w.d, n, sp);
    d2h(w, proxies, sp, sizeof(pr_proxy) * n);
    dev_free(sp);
    return PR_OK;
    API_CATCH(pw)
}

int pr_load_state(pr_world* pw, uint32_t n_bodies, const pr_body* bodies, uint32_t n_proxies, const pr_proxy* proxies,
                  uint32_t n_contacts, const pr_contact* contacts, const pr_manifold* manifolds, float inv_dt0)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (n_bodies != w.nBodies) { w.error = "body count mismatch"; return PR_ERR_INVALID_ARGUMENT; }
    std::vector<unsigned> ids(n_bodies);
    for (unsigned i = 0; i < n_bodies; ++i) ids[i] = i;
    const int rc = pr_set_bodies(pw, n_bodies, ids.data(), bodies);
    if (rc != PR_OK) return rc;
    dset(w, w.d.b_moved, 0, 4 * w.capBodies); dset(w, w.d.b_sync, 0, 4 * w.capBodies);
    w.anyBodySet = false;
    // proxies
    ensure(w, G_PROXY, n_proxies + 1);
    std::vector<unsigned> body(n_proxies), shape(n_proxies), child(n_proxies), cg(n_proxies), tree(n_proxies);
    std::vector<float4> fat(n_proxies);
    unsigned maxTree = 0;
    for (unsigned i = 0; i < n_proxies; ++i) {
        const pr_proxy& p = proxies[i];
        if (p.shape >= w.nShapes || p.body >= w.nBodies) { w.error = "proxy refers to unknown body/shape"; return PR_ERR_OUT_OF_RANGE; }
        if (p.child >= w.shapes[p.shape].child_count) { w.error = "proxy refers to a child its shape does not have"; return PR_ERR_OUT_OF_RANGE; }
        body[i] = p.body; shape[i] = p.shape; child[i] = p.child; cg[i] = w.shapes[p.shape].child_offset + p.child; tree[i] = p.tree_id;
        fat[i] = make_float4(p.min_x, p.min_y, p.max_x, p.max_y);
        maxTree = std::max(maxTree, p.tree_id);
    }
    h2d(w, w.d.p_body, body.data(), 4 * n_proxies); h2d(w, w.d.p_shape, shape.data(), 4 * n_proxies);
    h2d(w, w.d.p_child, child.data(), 4 * n_proxies); h2d(w, w.d.p_cg, cg.data(), 4 * n_proxies);
    h2d(w, w.d.p_tree, tree.data(), 4 * n_proxies); h2d(w, w.d.p_fat, fat.data(), sizeof(float4) * n_proxies);
    dset(w, w.d.p_moved, 0, 4 * w.capProxies);
    w.nProxies = n_proxies; w.maxTreeId = maxTree;
    w.pools.pxBody = body; w.pools.pxShape = shape; w.pools.pxTree = tree;
    { // the leaf allocator as a tree with these leaves and no destroyed node between them would stand (2k - 1 sequence)
        std::vector<char> used(static_cast<size_t>(maxTree) + 2u, 0);
        for (unsigned t : tree) used[t] = 1;
        w.pools.treeFree.clear();
        w.pools.treeNextFresh = n_proxies ? maxTree + 2u : 0u;
        if (n_proxies) for (unsigned t = maxTree; t-- > 0u;) if ((t & 1u) && !used[t]) w.pools.treeFree.push_back(t); // (odd ids below the maximum that no proxy holds: freed leaves)
        if (n_proxies && !used[0] ) w.pools.treeFree.push_back(0u);
    }
    w.bodyProxies.assign(w.nBodies, {});
    for (unsigned i = 0; i < n_proxies; ++i) w.bodyProxies[body[i]].push_back(i);
    w.bpDirty = true; w.bvhValid = false; w.toiAlphaDirty = true;
    w.fixturesForProxies.clear();
    // contacts
    unsigned maxId = 0;
    std::vector<char> treeKnown(static_cast<size_t>(maxTree) + 1u, 0);
    for (unsigned i = 0; i < n_proxies; ++i) treeKnown[tree[i]] = 1;
    for (unsigned i = 0; i < n_contacts; ++i) { // (validated before anything is uploaded: the import kernel indexes pxOfTree with the keys)
        const pr_contact& c = contacts[i];
        if (c.key_min > maxTree || c.key_max > maxTree || !treeKnown[c.key_min] || !treeKnown[c.key_max]) { w.error = "contact key names no proxy"; return PR_ERR_OUT_OF_RANGE; }
        if (c.id == 0xFFFFFFFFu) { w.error = "invalid contact id"; return PR_ERR_OUT_OF_RANGE; }
    }
    for (unsigned i = 0; i < n_contacts; ++i) maxId = std::max(maxId, contacts[i].id + 1);
    ensure(w, G_CONTACT, std::max(maxId, n_contacts) + 1);
    std::vector<unsigned> pxOfTree(maxTree + 1, 0u);
    for (unsigned i = 0; i < n_proxies; ++i) pxOfTree[tree[i]] = i;
    std::vector<char> usedId(maxId, 0);
    for (unsigned i = 0; i < n_contacts; ++i) usedId[contacts[i].id] = 1;
    std::vector<unsigned> freeIds;
    for (unsigned id = maxId; id-- > 0;) if (!usedId[id]) freeIds.push_back(id);
    if (n_contacts) {
        pr_contact* sc = static_cast<pr_contact*>(dev_alloc(sizeof(pr_contact) * n_contacts));
        pr_manifold* sm = static_cast<pr_manifold*>(dev_alloc(sizeof(pr_manifold) * n_contacts));
        unsigned* dpx = static_cast<unsigned*>(dev_alloc(4 * pxOfTree.size()));
        h2d(w, sc, contacts, sizeof(pr_contact) * n_contacts); h2d(w, sm, manifolds, sizeof(pr_manifold) * n_contacts);
        h2d(w, dpx, pxOfTree.data(), 4 * pxOfTree.size());
        PR_LAUNCH(k_import_contacts, n_contacts, w.stream, w.d, n_contacts, sc, sm, dpx);
        sync(w);
        dev_free(sc); dev_free(sm); dev_free(dpx);
    }
    h2d(w, w.d.c_free, freeIds.data(), 4 * freeIds.size());
    sync(w);
    w.nOrder = n_contacts; w.nFree = static_cast<unsigned>(freeIds.size()); w.poolSize = maxId;
    w.hashDirty = true; w.adjValid = false; w.stepComplete = true; w.inv_dt0 = inv_dt0; w.recolorAll = true;
    reset_counters(w);
    return PR_OK;
    API_CATCH(pw)
}

int pr_step(pr_world* pw, const pr_step_conf* conf, pr_step_stats* stats_out)
{
    API_TRY(pw)
    pr_step_stats st;
    if (pw->w.jointTablesDirty) { refresh_joint_tables(pw->w); pw->w.jointTablesDirty = false; } // (built when a Step needs them, not per CreateJoint / Destroy)
    step_impl(pw->w, *conf, st);
#ifndef PR_HOST_EMUL
    if (getenv("PR_SYNC_DEBUG")) {
        std::fprintf(stderr, "step: %u host syncs:", g_hostSyncs);
        for (const auto& kv : g_syncHist) std::fprintf(stderr, " L%d x%d", kv.first, kv.second);
        std::fprintf(stderr, "\n");
        g_hostSyncs = 0; g_syncHist.clear();
    }
#endif
    if (stats_out) *stats_out = st;
    return PR_OK;
    API_CATCH(pw)
}

int pr_get_solve_order(const pr_world* pw, uint32_t cap, uint32_t* contact_ids, uint32_t* colors, uint32_t* n_out)
{
    const World& w = pw->w;
    *n_out = w.nCons;
    const unsigned n = std::min<unsigned>(cap, static_cast<unsigned>(std::min<size_t>(w.nCons, w.solveOrderIds.size())));
    for (unsigned i = 0; i < n; ++i) { contact_ids[i] = w.solveOrderIds[i]; if (colors) colors[i] = w.solveOrderColors[i]; }
    return PR_OK;
}

int pr_get_toi_schedule(const pr_world* pw, uint32_t cap, uint32_t* round_contact_pairs, uint32_t* n_out)
{
    const World& w = pw->w;
    *n_out = w.toiScheduleN;
    const unsigned n = std::min<unsigned>(cap, static_cast<unsigned>(w.toiSchedule.size()));
    for (unsigned i = 0; i < n; ++i) { round_contact_pairs[2u * i] = w.toiSchedule[i].x; round_contact_pairs[2u * i + 1u] = w.toiSchedule[i].y; }
    return w.toiScheduleN <= cap && w.toiSchedule.size() == w.toiScheduleN ? PR_OK : PR_ERR_LENGTH;
}

int pr_set_debug(pr_world* pw, uint32_t flags) { pw->w.debugFlags = flags; return PR_OK; }

int pr_get_step_profile(const pr_world* pw, pr_step_profile* out) { *out = pw->w.profile; return PR_OK; }

int pr_shift_origin(pr_world* pw, float x, float y)
{
    API_TRY(pw)
    World& w = pw->w;
    w.contactsClean = false; // (a mutator: the next Step's pre-phase UpdateContacts must look at every contact)
    if (!w.stepComplete) { w.error = "ShiftOrigin: world is locked (sub-stepped Step incomplete)"; return PR_ERR_WRONG_STATE; }
    const unsigned n = std::max(w.nBodies, w.nProxies);
    if (n) PR_LAUNCH(k_shift_origin, n, w.stream, w.d, w.nBodies, w.nProxies, x, y);
    if (w.nJoints) PR_LAUNCH(k_shift_joints, w.nJoints, w.stream, w.d, w.nJoints, x, y);
    w.anyBodySet = true;  // the pre-phase UpdateContacts of the next Step looks at b_moved (FlagForUpdating, :2235)
    w.bvhValid = false;   // every leaf box moved: rebuild the LBVH
    w.toiAlphaDirty = true;
    return PR_OK;
    API_CATCH(pw)
}

int pr_set_event_mask(pr_world* pw, uint32_t mask, uint32_t capacity)
{
    API_TRY(pw)
    World& w = pw->w;
    if (mask & ~static_cast<uint32_t>(PR_EVENT_BEGIN | PR_EVENT_END | PR_EVENT_PRE_SOLVE | PR_EVENT_POST_SOLVE)) { w.error = "unknown event kind"; return PR_ERR_INVALID_ARGUMENT; }
    sync(w);
    if (mask && capacity > w.d.ev_cap) {
        dev_free(w.d.ev_contact); dev_free(w.d.ev_impulse); dev_free(w.d.ev_presolve);
        w.d.ev_presolve = static_cast<pr_presolve_event*>(dev_alloc(sizeof(pr_presolve_event) * capacity));
        w.d.ev_contact = static_cast<pr_contact_event*>(dev_alloc(sizeof(pr_contact_event) * capacity));
        w.d.ev_impulse = static_cast<pr_impulse_event*>(dev_alloc(sizeof(pr_impulse_event) * capacity));
        w.d.ev_cap = capacity;
    }
    if (mask && !w.d.ev_cap) { w.error = "event capacity is zero"; return PR_ERR_INVALID_ARGUMENT; }
    w.d.ev_mask = mask;
    return PR_OK;
    API_CATCH(pw)
}

int pr_events(const pr_world* cpw, uint32_t kind, void* buf, uint32_t cap, uint32_t* n_out)
{
    pr_world* pw = const_cast<pr_world*>(cpw);
    API_TRY(pw)
    World& w = pw->w;
    const bool contactKinds = (kind & (PR_EVENT_BEGIN | PR_EVENT_END)) != 0 && !(kind & ~static_cast<uint32_t>(PR_EVENT_BEGIN | PR_EVENT_END));
    if (!contactKinds && kind != PR_EVENT_POST_SOLVE && kind != PR_EVENT_PRE_SOLVE) { w.error = "pr_events: kind must be begin|end, pre-solve or post-solve"; return PR_ERR_INVALID_ARGUMENT; }
    read_counters(w);
    const unsigned total = contactKinds ? w.hostCnt.nEvContact : kind == PR_EVENT_PRE_SOLVE ? w.hostCnt.nEvPre : w.hostCnt.nEvImpulse;
    const unsigned have = std::min(total, w.d.ev_cap);
    if (kind == PR_EVENT_PRE_SOLVE) {
        std::vector<pr_presolve_event> ev(have);
        if (have) d2h(w, ev.data(), w.d.ev_presolve, sizeof(pr_presolve_event) * have);
        std::sort(ev.begin(), ev.end(), [](const pr_presolve_event& a, const pr_presolve_event& b) {
            return a.contact != b.contact ? a.contact < b.contact : std::memcmp(&a.old_manifold, &b.old_manifold, sizeof(pr_manifold)) < 0; });
        if (n_out) *n_out = static_cast<uint32_t>(ev.size());
        if (buf && !ev.empty()) std::memcpy(buf, ev.data(), sizeof(pr_presolve_event) * std::min<size_t>(ev.size(), cap));
    } else if (contactKinds) {
        std::vector<pr_contact_event> ev(have);
        if (have) d2h(w, ev.data(), w.d.ev_contact, sizeof(pr_contact_event) * have);
        ev.erase(std::remove_if(ev.begin(), ev.end(), [&](const pr_contact_event& e) { return !(e.kind & kind); }), ev.end());
        std::sort(ev.begin(), ev.end(), [](const pr_contact_event& a, const pr_contact_event& b) { return a.kind != b.kind ? a.kind < b.kind : a.contact < b.contact; });
        if (n_out) *n_out = static_cast<uint32_t>(ev.size());
        if (buf && !ev.empty()) std::memcpy(buf, ev.data(), sizeof(pr_contact_event) * std::min<size_t>(ev.size(), cap));
    } else {
        std::vector<pr_impulse_event> ev(have);
        if (have) d2h(w, ev.data(), w.d.ev_impulse, sizeof(pr_impulse_event) * have);
        std::sort(ev.begin(), ev.end(), [](const pr_impulse_event& a, const pr_impulse_event& b) {
            if (a.phase != b.phase) return a.phase < b.phase;
            if (a.contact != b.contact) return a.contact < b.contact;
            if (a.iterations != b.iterations) return a.iterations < b.iterations;
            return std::memcmp(a.normal, b.normal, 16) < 0; });
        if (n_out) *n_out = static_cast<uint32_t>(ev.size());
        if (buf && !ev.empty()) std::memcpy(buf, ev.data(), sizeof(pr_impulse_event) * std::min<size_t>(ev.size(), cap));
    }
    if (total > w.d.ev_cap) { // the step dropped events: say how many there were
        if (n_out) *n_out = total;
        w.error = "events of the last step exceeded the capacity given to pr_set_event_mask"; return PR_ERR_LENGTH;
    }
    return PR_OK;
    API_CATCH(pw)
}

float pr_get_inv_dt0(const pr_world* pw) { return pw->w.inv_dt0; }
void* pr_get_stream(const pr_world* pw)
{
#ifndef PR_HOST_EMUL
    return reinterpret_cast<void*>(pw->w.stream);
#else
    return nullptr;
#endif
}

int pr_set_toi_mode(pr_world* pw, uint32_t mode)
{
    if (mode > PR_TOI_PARALLEL) { pw->w.error = "unknown TOI mode"; return PR_ERR_INVALID_ARGUMENT; }
    pw->w.toiMode = mode;
    return PR_OK;
}
int pr_set_sub_stepping(pr_world* pw, uint32_t flag) { pw->w.subStepping = flag != 0; return PR_OK; }
int pr_get_sub_stepping(const pr_world* pw) { return pw->w.subStepping ? 1 : 0; }
int pr_is_step_complete(const pr_world* pw) { return pw->w.stepComplete ? 1 : 0; }

int pr_toi_batch(int device, uint32_t n, const pr_child* children, uint32_t n_children, const pr_vec2* vertices,
                 const pr_vec2* normals, uint32_t n_vertices, const uint32_t* a_child, const float* sweep_a,
                 const uint32_t* b_child, const float* sweep_b, const pr_step_conf* conf, float* out_time,
                 int32_t* out_state, uint32_t* out_iters)
{
    try {
        World w; // only for the stream shim
#ifndef PR_HOST_EMUL
        if (device >= 0) CK(cudaSetDevice(device));
        w.stream = 0;
#else
        (void)device;
#endif
        if (!n) return PR_OK;
        pr_child* dch = static_cast<pr_child*>(dev_alloc(sizeof(pr_child) * n_children));
        float2* dv = static_cast<float2*>(dev_alloc(sizeof(float2) * n_vertices));
        float2* dn = static_cast<float2*>(dev_alloc(sizeof(float2) * n_vertices));
        unsigned* da = static_cast<unsigned*>(dev_alloc(4 * n)); unsigned* db = static_cast<unsigned*>(dev_alloc(4 * n));
        float* dsa = static_cast<float*>(dev_alloc(32 * n)); float* dsb = static_cast<float*>(dev_alloc(32 * n));
        float* dt = static_cast<float*>(dev_alloc(4 * n)); int* dst = static_cast<int*>(dev_alloc(4 * n));
        unsigned* dit = static_cast<unsigned*>(dev_alloc(12 * n));
        h2d(w, dch, children, sizeof(pr_child) * n_children); h2d(w, dv, vertices, 8 * n_vertices); h2d(w, dn, normals, 8 * n_vertices);
        h2d(w, da, a_child, 4 * n); h2d(w, db, b_child, 4 * n); h2d(w, dsa, sweep_a, 32 * n); h2d(w, dsb, sweep_b, 32 * n);
        PR_LAUNCH(k_toi_batch, n, w.stream, n, dch, dv, dn, da, dsa, db, dsb, *conf, dt, dst, dit);
        d2h(w, out_time, dt, 4 * n); d2h(w, out_state, dst, 4 * n); d2h(w, out_iters, dit, 12 * n);
        dev_free(dch); dev_free(dv); dev_free(dn); dev_free(da); dev_free(db); dev_free(dsa); dev_free(dsb);
        dev_free(dt); dev_free(dst); dev_free(dit);
        return PR_OK;
    } catch (const CudaError& e) { g_createError = e.msg; return PR_ERR_CUDA; }
}

int pr_collide_shapes(int device, uint32_t n, const pr_child* children, uint32_t n_children, const pr_vec2* vertices,
                      const pr_vec2* normals, uint32_t n_vertices, const uint32_t* a_child, const float* xf_a,
                      const uint32_t* b_child, const float* xf_b, float linear_slop, float max_circles_ratio, pr_manifold* out)
{
    try {
        World w; // only for the stream shim
#ifndef PR_HOST_EMUL
        if (device >= 0) CK(cudaSetDevice(device));
        w.stream = 0;
#else
        (void)device;
#endif
        if (!n) return PR_OK;
        pr_child* dch = static_cast<pr_child*>(dev_alloc(sizeof(pr_child) * n_children));
        float2* dv = static_cast<float2*>(dev_alloc(sizeof(float2) * n_vertices));
        float2* dn = static_cast<float2*>(dev_alloc(sizeof(float2) * n_vertices));
        unsigned* da = static_cast<unsigned*>(dev_alloc(4 * n)); unsigned* db = static_cast<unsigned*>(dev_alloc(4 * n));
        float4* dxa = static_cast<float4*>(dev_alloc(16 * n)); float4* dxb = static_cast<float4*>(dev_alloc(16 * n));
        pr_manifold* dout = static_cast<pr_manifold*>(dev_alloc(sizeof(pr_manifold) * n));
        h2d(w, dch, children, sizeof(pr_child) * n_children); h2d(w, dv, vertices, 8 * n_vertices); h2d(w, dn, normals, 8 * n_vertices);
        h2d(w, da, a_child, 4 * n); h2d(w, db, b_child, 4 * n); h2d(w, dxa, xf_a, 16 * n); h2d(w, dxb, xf_b, 16 * n);
        PR_LAUNCH(k_collide_batch, n, w.stream, n, dch, dv, dn, da, dxa, db, dxb, linear_slop, max_circles_ratio, dout);
        d2h(w, out, dout, sizeof(pr_manifold) * n);
        dev_free(dch); dev_free(dv); dev_free(dn); dev_free(da); dev_free(db); dev_free(dxa); dev_free(dxb); dev_free(dout);
        return PR_OK;
    } catch (const CudaError& e) { g_createError = e.msg; return PR_ERR_CUDA; }
}

// sizes of the POD structs, for the python mirror (tests/test_abi.py)
int pr_abi_sizeof(int which)
{
    switch (which) {
    case 0: return sizeof(pr_body); case 1: return sizeof(pr_child); case 2: return sizeof(pr_shape);
    case 3: return sizeof(pr_revolute_joint); case 4: return sizeof(pr_manifold); case 5: return sizeof(pr_contact);
    case 6: return sizeof(pr_proxy); case 7: return sizeof(pr_step_conf); case 8: return sizeof(pr_step_stats);
    case 9: return sizeof(pr_world_conf); case 10: return sizeof(pr_counts); case 11: return sizeof(pr_step_profile);
    default: return -1;
    }
}

// Batched independent worlds (BASELINE config 5): all n_worlds worlds live in ONE device arena. Bodies carry their world
// index; the index is the prefix of the LBVH sort key, so every world is its own subtree and a query never leaves
// the queried proxy's world (k_bvh_query / query_emit). No contact ever joins two worlds, hence islands, colours, the
// solver's per-island early exits, sleeping and TOI events are per world by construction, and one pr_step of the arena
// is one Step of every world: each kernel launch is n_worlds wide instead of one launch train per world.
struct pr_batch { pr_world* arena; uint32_t nWorlds; pr_step_stats lastStats{}; };
int pr_batch_create(const pr_world_conf* conf, uint32_t n_worlds, pr_batch** out)
{
    *out = nullptr;
    if (n_worlds == 0 || n_worlds > (1u << 20)) { g_createError = "pr_batch_create: n_worlds must be in 1..2^20"; return PR_ERR_INVALID_ARGUMENT; }
    pr_world* a = nullptr;
    const int rc = pr_create(conf, &a);
    if (rc != PR_OK) return rc;
    unsigned bits = 1; while ((1u << bits) < n_worlds) ++bits;
    a->w.d.nWorlds = n_worlds; a->w.d.worldBits = bits;
    pr_batch* b = new pr_batch();
    b->arena = a; b->nWorlds = n_worlds;
    *out = b;
    return PR_OK;
}
int pr_batch_create_masked(const pr_world_conf* conf, uint32_t n_worlds, uint64_t device_mask, pr_batch** out)
{
    *out = nullptr;
    int count = 1;
#ifndef PR_HOST_EMUL
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); count = 0; }
#endif
    for (int dev = 0; dev < count && dev < 64; ++dev)
        if (device_mask & (1ull << dev)) { pr_world_conf c = *conf; c.device = dev; return pr_batch_create(&c, n_worlds, out); }
    g_createError = "pr_batch_create_masked: no device of the mask exists";
    return PR_ERR_INVALID_ARGUMENT;
}
int pr_batch_stats(const pr_batch* b, pr_step_stats* out) { if (!b || !out) return PR_ERR_INVALID_ARGUMENT; *out = b->lastStats; return PR_OK; }
int pr_batch_destroy(pr_batch* b) { if (b) { pr_destroy(b->arena); delete b; } return PR_OK; }
pr_world* pr_batch_arena(pr_batch* b) { return b ? b->arena : nullptr; }
uint32_t pr_batch_world_count(const pr_batch* b) { return b ? b->nWorlds : 0u; }
int pr_batch_create_bodies(pr_batch* b, uint32_t n, const pr_body* bodies, const uint32_t* world_index,
                           const uint32_t* shape_offsets, const uint32_t* shape_ids, uint32_t* first_id)
{
    return create_bodies(b->arena, n, bodies, shape_offsets, shape_ids, world_index, first_id);
}
int pr_batch_step(pr_batch* b, const pr_step_conf* conf, pr_step_stats* sum)
{
    const int rc = pr_step(b->arena, conf, &b->lastStats);
    if (sum) *sum = b->lastStats;
    return rc;
}
int pr_batch_get_body_worlds(const pr_batch* b, uint32_t first, uint32_t n, uint32_t* world_index)
{
    const World& w = b->arena->w;
    if (first > w.nBodies || n > w.nBodies - first) return PR_ERR_OUT_OF_RANGE;
    std::memcpy(world_index, w.bodyWorld.data() + first, sizeof(uint32_t) * n);
    return PR_OK;
}

} // extern "C"
