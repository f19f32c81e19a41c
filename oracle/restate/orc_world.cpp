// orc_world.cpp -- TEST INFRASTRUCTURE ONLY (oracle restatement; see pr_oracle.h).
// Serial restatement of Step(AabbTreeWorld&, const StepConf&) and its callees:
//   Library/source/playrho/d2/AabbTreeWorld.cpp (line numbers cited per function),
//   ContactSolver.cpp, VelocityConstraint.cpp, WorldManifold.cpp, PositionSolverManifold.cpp,
//   Body.cpp, Velocity.cpp, AABB.cpp, Interval.hpp.
// The broadphase is a plain O(queued x proxies) scan: the candidate pair SET of FindContacts
// (AabbTreeWorld.cpp:709-751) does not depend on the tree's shape, only on the fat AABBs.
// Proxy (tree leaf) ids follow the creation-only allocation sequence of DynamicTree
// (DynamicTree.cpp:456-537): leaf 0 -> id 0, leaf k -> id 2k-1.
#include "pr_oracle.h"
#include "orc_math.h"
#include "orc_toi.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace orc;

namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();

struct AABB { float minx, maxx, miny, maxy; };

struct OBody {
    pr_body b;
    std::vector<uint32_t> shapes;
    std::vector<uint32_t> proxies;  // dense proxy indices
    std::vector<uint32_t> contacts; // contact ids, append order (m_bodyContacts)
    std::vector<uint32_t> joints;   // joint ids (m_bodyJoints)
};
struct OProxy { pr_proxy p; };
struct OContact { pr_contact c; pr_manifold m; uint32_t pxA, pxB; };

// BodyConstraint, BodyConstraint.hpp
struct BC { float invM, invI; V2 lc; Pos pos; Vel vel; };
// VelocityConstraint::Point / VelocityConstraint, VelocityConstraint.hpp
struct VCP { V2 relA, relB; float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias; };
struct VC {
    uint32_t bodyA, bodyB;
    Rot normal;
    float friction, restitution, tangentSpeed;
    float K[3];          // k00, k11, k01
    float normalMass[3]; // m00, m11, m01
    VCP pt[2];
    uint32_t count;
};
struct PC { pr_manifold m; uint32_t bodyA, bodyB; float totalRadius; };
struct IslandStats {
    float minSeparation = kInf, maxIncImpulse = 0;
    uint32_t bodiesSlept = 0, positionIters = 0, velocityIters = 0, contactsUpdated = 0, contactsSkipped = 0;
    bool solved = false;
};

} // namespace

struct orc_world {
    std::vector<pr_shape> shapes;
    std::vector<pr_child> children;
    std::vector<pr_vec2> verts, normals;
    std::vector<OBody> bodies;
    std::vector<OProxy> proxies;
    std::vector<OContact> contacts;      // pool indexed by contact id
    std::vector<uint32_t> freeContacts;  // ObjectPool free list (LIFO), ObjectPool.hpp
    std::vector<uint32_t> order;         // m_contacts (ids, creation order)
    std::vector<pr_revolute_joint> joints;
    std::vector<uint32_t> freeJoints;            // joint pool free list (LIFO)
    std::vector<orc::RevoluteWork> jointWork;
    std::vector<pr_gear_joint> gears;   // by joint id: the whole value of a gear joint (GearJointConf), zeroed for other types
    std::vector<std::pair<uint32_t, uint32_t>> fixturesForProxies; // (body, shape)
    // ids of destroyed bodies / shapes (ObjectPool free lists, LIFO) and the DynamicTree's free list as the LEAVES see it:
    // DestroyLeaf frees the leaf's parent node and then the leaf (DynamicTree.cpp:539-569), CreateLeaf takes a leaf and then a
    // parent (:517-537), so the k-th new leaf gets the k-th last destroyed leaf id; with the list empty a leaf takes the
    // next fresh node and its parent the one after it (the first leaf of an empty tree has no parent)
    std::vector<uint32_t> bodyFree, shapeFree, treeFree;
    std::vector<uint32_t> bodyOrder; // m_bodies: live body ids in creation order (a re-used id goes to the END: SolveReg seeds its islands in this order)
    uint32_t treeNextFresh = 0;
    std::vector<char> bodyDead, shapeDead;
    bool needsFiltering = false; // e_needsContactFiltering
    std::vector<uint32_t> proxiesForContacts;
    std::vector<uint32_t> bodiesForSync;
    uint32_t evMask = 0;                          // listeners installed (PR_EVENT_* bits)
    std::vector<pr_contact_event> evContact;      // begin/end listener calls of the last step, call order
    std::vector<pr_impulse_event> evImpulse;      // post-solve listener calls of the last step, call order
    std::vector<pr_presolve_event> evPre;         // pre-solve listener calls of the last step, call order
    std::vector<char> islandedBodies, islandedContacts, islandedJoints;
    bool stepComplete = true;
    bool subStepping = false;
    float inv_dt0 = 0;
    std::string error;
    uint32_t toiMode = 0;                                        // 0: the reference's event order; 1: PR_TOI_PARALLEL rounds (orc_toi_phase.inc)
    std::vector<BC> bcScratch;
    std::vector<uint32_t> lastToiIslandBodies;
    std::vector<std::pair<uint32_t, uint32_t>> toiSchedule;      // (round, contact id) of the last step's TOI events, toiMode 1

    Proxy child(uint32_t shape, uint32_t childIdx) const
    {
        const pr_child& c = children[shapes[shape].child_offset + childIdx];
        return {c.vertex_radius, c.vertex_count, verts.data() + c.vertex_offset, normals.data() + c.vertex_offset};
    }
    float radius(uint32_t shape, uint32_t childIdx) const
    {
        return children[shapes[shape].child_offset + childIdx].vertex_radius;
    }
};

namespace {

inline Xf body_xf(const pr_body& b) { return {{b.xf_px, b.xf_py}, {b.xf_qc, b.xf_qs}}; }
inline Pos body_pos0(const pr_body& b) { return {{b.pos0_x, b.pos0_y}, b.pos0_a}; }
inline Pos body_pos1(const pr_body& b) { return {{b.pos1_x, b.pos1_y}, b.pos1_a}; }
inline V2 body_lc(const pr_body& b) { return {b.lc_x, b.lc_y}; }
inline bool speedable(const pr_body& b) { return (b.flags & PR_BODY_VELOCITY) != 0; }
inline bool accelerable(const pr_body& b) { return (b.flags & PR_BODY_ACCELERATION) != 0; }
inline bool awake(const pr_body& b) { return (b.flags & PR_BODY_AWAKE) != 0; }
inline bool enabled(const pr_body& b) { return (b.flags & PR_BODY_ENABLED) != 0; }
inline bool impenetrable(const pr_body& b) { return (b.flags & PR_BODY_IMPENETRABLE) != 0; }
// Body::SetPosition1, Body.hpp:710-715
inline void set_pos1(pr_body& b, Pos p)
{
    b.pos1_x = p.l.x; b.pos1_y = p.l.y; b.pos1_a = p.a;
    const Xf xf = get_xf(p, body_lc(b));
    b.xf_px = xf.p.x; b.xf_py = xf.p.y; b.xf_qc = xf.q.c; b.xf_qs = xf.q.s;
}
inline void set_pos0(pr_body& b, Pos p) { b.pos0_x = p.l.x; b.pos0_y = p.l.y; b.pos0_a = p.a; }
// Body::SetAwake, Body.cpp:163-171
inline void set_awake(pr_body& b)
{
    if (speedable(b)) { b.flags |= PR_BODY_AWAKE; b.under_active_time = 0; }
}
// Body::UnsetAwake, Body.cpp:173-182; Unawaken(body) returns whether it was awake and now is not
inline bool unawaken(pr_body& b)
{
    if (!awake(b)) return false;
    if (!speedable(b) || (b.flags & PR_BODY_AUTOSLEEP)) {
        b.flags &= ~uint32_t(PR_BODY_AWAKE);
        b.under_active_time = 0; b.vel_x = b.vel_y = b.vel_w = 0;
        return true;
    }
    return false;
}

// ComputeAABB(proxy, xf0, xf1) AABB.cpp:51-62 + Interval::Include + GetFattenedAABB(radius)
AABB compute_aabb(const Proxy& p, Xf xf0, Xf xf1)
{
    AABB r{kInf, -kInf, kInf, -kInf};
    auto include = [&](V2 v) {
        r.minx = std::min(v.x, r.minx); r.maxx = std::max(v.x, r.maxx);
        r.miny = std::min(v.y, r.miny); r.maxy = std::max(v.y, r.maxy);
    };
    for (uint32_t i = 0; i < p.count; ++i) { include(transform(p.vertex(i), xf0)); include(transform(p.vertex(i), xf1)); }
    return {r.minx - p.radius, r.maxx + p.radius, r.miny - p.radius, r.maxy + p.radius};
}
// GetFattenedAABB / GetDisplacedAABB, detail/AABB.hpp:305-320; Interval::Expand/ExpandEqually
AABB fatten_displace(AABB a, float ext, V2 d)
{
    a = {a.minx - ext, a.maxx + ext, a.miny - ext, a.maxy + ext};
    if (d.x < 0.0f) a.minx += d.x; else a.maxx += d.x;
    if (d.y < 0.0f) a.miny += d.y; else a.maxy += d.y;
    return a;
}
inline AABB fat_of(const pr_proxy& p) { return {p.min_x, p.max_x, p.min_y, p.max_y}; }
inline void set_fat(pr_proxy& p, AABB a) { p.min_x = a.minx; p.max_x = a.maxx; p.min_y = a.miny; p.max_y = a.maxy; }
// Contains: IsEntirelyEnclosing per axis, Interval.hpp
inline bool contains(AABB a, AABB b) { return a.minx <= b.minx && a.maxx >= b.maxx && a.miny <= b.miny && a.maxy >= b.maxy; }
// TestOverlap: IsIntersecting per axis (closed), Interval.hpp
inline bool overlap(AABB a, AABB b)
{
    return (std::min(a.maxx, b.maxx) >= std::max(a.minx, b.minx)) && (std::min(a.maxy, b.maxy) >= std::max(a.miny, b.miny));
}

// Report -> GetContactImpulses(vc), ContactImpulsesList.cpp:28-37
void record_impulse_event(orc_world& w, uint32_t id, const VC& vc, uint32_t iterations, uint32_t phase)
{
    pr_impulse_event e;
    std::memset(&e, 0, sizeof(e));
    e.contact = id; e.count = vc.count; e.iterations = iterations; e.phase = phase;
    for (uint32_t j = 0; j < vc.count; ++j) { e.normal[j] = vc.pt[j].normalImpulse; e.tangent[j] = vc.pt[j].tangentImpulse; }
    w.evImpulse.push_back(e);
}
// ---------------- narrowphase Update, AabbTreeWorld.cpp:2587-2721 ----------------
void record_contact_event(orc_world& w, uint32_t kind, uint32_t id)
{
    if (!(w.evMask & kind)) return;
    const pr_contact& c = w.contacts[id].c;
    w.evContact.push_back(pr_contact_event{kind, id, c.body_a, c.shape_a, c.child_a, c.body_b, c.shape_b, c.child_b});
}
void update_contact(orc_world& w, uint32_t id, const pr_step_conf& conf)
{
    OContact& oc = w.contacts[id];
    pr_contact& c = oc.c;
    const pr_manifold oldM = oc.m;
    const bool oldTouching = (c.flags & PR_CONTACT_TOUCHING) != 0;
    bool newTouching = false;
    const Xf xfA = body_xf(w.bodies[c.body_a].b), xfB = body_xf(w.bodies[c.body_b].b);
    const Proxy childA = w.child(c.shape_a, c.child_a), childB = w.child(c.shape_b, c.child_b);
    if (c.flags & PR_CONTACT_SENSOR) {
        // TestOverlap(childA, xfA, childB, xfB, conf.distance) >= 0, Distance.cpp:205-230
        newTouching = orc::test_overlap(childA, xfA, childB, xfB, conf.max_distance_iters) >= 0.0f;
        pr_manifold m; std::memset(&m, 0, sizeof(m)); m.cf0 = m.cf1 = CF_DEFAULT;
        oc.m = m;
    } else {
        pr_manifold nm = collide_shapes(childA, xfA, childB, xfB, conf.linear_slop, conf.max_circles_ratio);
        const uint32_t oldCount = oldM.count, newCount = nm.count;
        newTouching = newCount > 0;
        const pr_contact_feature ocf[2] = {oldM.cf0, oldM.cf1};
        const float ojn[2] = {oldM.jn0, oldM.jn1}, ojt[2] = {oldM.jt0, oldM.jt1};
        const V2 op[2] = {{oldM.p0_x, oldM.p0_y}, {oldM.p1_x, oldM.p1_y}};
        pr_contact_feature ncf[2] = {nm.cf0, nm.cf1};
        float* njn[2] = {&nm.jn0, &nm.jn1};
        float* njt[2] = {&nm.jt0, &nm.jt1};
        const V2 np[2] = {{nm.p0_x, nm.p0_y}, {nm.p1_x, nm.p1_y}};
        bool found[2] = {false, newCount < 2};
        for (uint32_t i = 0; i < newCount; ++i) {
            for (uint32_t j = 0; j < oldCount; ++j) {
                if (ncf[i] == ocf[j]) { found[i] = true; *njn[i] = ojn[j]; *njt[i] = ojt[j]; break; }
            }
        }
        for (uint32_t i = 0; i < newCount; ++i) {
            if (!found[i]) {
                float least = kInf;
                for (uint32_t j = 0; j < oldCount; ++j) {
                    const float sq = magsq(op[j] - np[i]);
                    if (least > sq) { least = sq; *njn[i] = ojn[j]; *njt[i] = ojt[j]; }
                }
            }
        }
        oc.m = nm;
    }
    c.flags &= ~uint32_t(PR_CONTACT_DIRTY);
    if (!oldTouching && newTouching) c.flags |= PR_CONTACT_TOUCHING;
    else if (oldTouching && !newTouching) c.flags &= ~uint32_t(PR_CONTACT_TOUCHING);
    if (!oldTouching && newTouching) record_contact_event(w, PR_EVENT_BEGIN, id);    // :2703-2708
    else if (oldTouching && !newTouching) record_contact_event(w, PR_EVENT_END, id); // :2709-2714
    if (!(c.flags & PR_CONTACT_SENSOR) && newTouching && (w.evMask & PR_EVENT_PRE_SOLVE)) { // :2716-2720
        pr_presolve_event e; std::memset(&e, 0, sizeof(e));
        e.contact = id; e.old_manifold = oldM;
        w.evPre.push_back(e);
    }
}

// UpdateContacts, AabbTreeWorld.cpp:2333-2421
void update_contacts(orc_world& w, const pr_step_conf& conf, uint32_t& updated, uint32_t& skipped)
{
    for (uint32_t id : w.order) {
        if (w.contacts[id].c.flags & PR_CONTACT_DIRTY) { update_contact(w, id, conf); ++updated; }
        else ++skipped;
    }
}

void flag_body_contacts_dirty(orc_world& w, uint32_t body)
{
    for (uint32_t id : w.bodies[body].contacts) w.contacts[id].c.flags |= PR_CONTACT_DIRTY;
}

// InternalDestroy, AabbTreeWorld.cpp:2245-2279
void internal_destroy(orc_world& w, uint32_t id)
{
    OContact& oc = w.contacts[id];
    if (oc.c.flags & PR_CONTACT_TOUCHING) record_contact_event(w, PR_EVENT_END, id); // :2249-2252
    for (uint32_t b : {oc.c.body_a, oc.c.body_b}) {
        auto& lst = w.bodies[b].contacts;
        auto it = std::find(lst.begin(), lst.end(), id);
        if (it != lst.end()) lst.erase(it);
    }
    if (oc.m.count > 0 && !(oc.c.flags & PR_CONTACT_SENSOR)) {
        set_awake(w.bodies[oc.c.body_a].b);
        set_awake(w.bodies[oc.c.body_b].b);
    }
    w.freeContacts.push_back(id);
    std::memset(&oc.c, 0, sizeof(oc.c));
    oc.c.id = id; oc.c.flags = PR_CONTACT_DESTROYED;
    std::memset(&oc.m, 0, sizeof(oc.m)); oc.m.cf0 = oc.m.cf1 = CF_DEFAULT;
}

bool joints_allow_collision(const orc_world& w, uint32_t a, uint32_t b)
{
    // ShouldCollide(jointBuffer, bodyJoints, lhs, rhs), AabbTreeWorld.cpp:479-490
    for (uint32_t j : w.bodies[a].joints) {
        const auto& jt = w.joints[j];
        const uint32_t other = (jt.body_a == a) ? jt.body_b : jt.body_a;
        if (other == b && !(jt.collide_connected & 1u)) return false;
    }
    return true;
}
// ShouldCollide(Filter, Filter), Filter.hpp
bool filters_allow(const pr_shape& a, const pr_shape& b)
{
    if (a.filter_group == b.filter_group && a.filter_group != 0) return a.filter_group > 0;
    return (a.filter_mask & b.filter_category) != 0 && (a.filter_category & b.filter_mask) != 0;
}

// DestroyContacts, AabbTreeWorld.cpp:2290-2331 (overlap part; filter flag part: flagged contacts)
uint32_t destroy_contacts(orc_world& w)
{
    uint32_t destroyed = 0;
    std::vector<uint32_t> keep;
    keep.reserve(w.order.size());
    for (uint32_t id : w.order) {
        const OContact& oc = w.contacts[id];
        if (!overlap(fat_of(w.proxies[oc.pxA].p), fat_of(w.proxies[oc.pxB].p))) { internal_destroy(w, id); ++destroyed; }
        else keep.push_back(id);
    }
    w.order.swap(keep);
    if (w.needsFiltering) { // :2305-2328
        keep.clear();
        for (uint32_t id : w.order) {
            OContact& oc = w.contacts[id];
            if (oc.c.flags & PR_CONTACT_FILTER) {
                const pr_body& bA = w.bodies[oc.c.body_a].b; const pr_body& bB = w.bodies[oc.c.body_b].b;
                if (!(accelerable(bA) || accelerable(bB)) || !joints_allow_collision(w, oc.c.body_a, oc.c.body_b) ||
                    !filters_allow(w.shapes[oc.c.shape_a], w.shapes[oc.c.shape_b])) { internal_destroy(w, id); ++destroyed; continue; }
                oc.c.flags &= ~static_cast<uint32_t>(PR_CONTACT_FILTER);
            }
            keep.push_back(id);
        }
        w.order.swap(keep);
        w.needsFiltering = false;
    }
    return destroyed;
}
// Destroy(ContactID, from), :2281-2288
void destroy_contact(orc_world& w, uint32_t id)
{
    auto it = std::find(w.order.begin(), w.order.end(), id);
    if (it != w.order.end()) w.order.erase(it);
    internal_destroy(w, id);
}
// DestroyLeaf for the proxy slots listed, in that order; the dense proxy table and everything that indexes it are renumbered
void destroy_proxy_slots(orc_world& w, const std::vector<uint32_t>& slots)
{
    if (slots.empty()) return;
    const uint32_t P = static_cast<uint32_t>(w.proxies.size());
    std::vector<uint32_t> map(P, 0u);
    for (uint32_t s : slots) { map[s] = 0xFFFFFFFFu; w.treeFree.push_back(w.proxies[s].p.tree_id); }
    uint32_t k = 0;
    for (uint32_t i = 0; i < P; ++i) if (map[i] != 0xFFFFFFFFu) { w.proxies[k] = w.proxies[i]; map[i] = k++; }
    w.proxies.resize(k);
    for (uint32_t id : w.order) { OContact& oc = w.contacts[id]; oc.pxA = map[oc.pxA]; oc.pxB = map[oc.pxB]; }
    auto remap = [&](std::vector<uint32_t>& v) { uint32_t m = 0; for (uint32_t e : v) if (map[e] != 0xFFFFFFFFu) v[m++] = map[e]; v.resize(m); };
    remap(w.proxiesForContacts);
    for (OBody& ob : w.bodies) remap(ob.proxies);
}
// the SetBody whose value lost shapes (:2760-2785): proxies in m_bodyProxies order, then per shape its contacts and queue entries
void body_remove_shapes(orc_world& w, uint32_t body, const std::vector<uint32_t>& shapeIds, bool keepInList)
{
    OBody& ob = w.bodies[body];
    std::vector<uint32_t> slots;
    for (uint32_t s : ob.proxies) if (std::find(shapeIds.begin(), shapeIds.end(), w.proxies[s].p.shape) != shapeIds.end()) slots.push_back(s);
    for (uint32_t sid : shapeIds) {
        const std::vector<uint32_t> list = ob.contacts;
        for (uint32_t id : list) {
            const pr_contact& c = w.contacts[id].c;
            if ((c.body_a == body && c.shape_a == sid) || (c.body_b == body && c.shape_b == sid)) destroy_contact(w, id);
        }
        auto& f = w.fixturesForProxies;
        f.erase(std::remove(f.begin(), f.end(), std::make_pair(body, sid)), f.end());
    }
    destroy_proxy_slots(w, slots);
    if (!keepInList) for (uint32_t sid : shapeIds) { auto it = std::find(ob.shapes.begin(), ob.shapes.end(), sid); if (it != ob.shapes.end()) ob.shapes.erase(it); }
}

// FindContacts + AddContacts, AabbTreeWorld.cpp:709-751, 2423-2547
uint32_t find_and_add_contacts(orc_world& w, const pr_step_conf& conf)
{
    struct Key { uint32_t mn, mx, pxMin, pxMax; };
    std::vector<Key> keys;
    std::vector<uint32_t> queued;
    queued.swap(w.proxiesForContacts);
    const uint32_t P = static_cast<uint32_t>(w.proxies.size());
    auto test_pair = [&](uint32_t q, uint32_t o) {
        if (o == q) return;
        const pr_proxy& pq = w.proxies[q].p;
        const pr_proxy& po = w.proxies[o].p;
        if (po.body == pq.body) return;
        if (!overlap(fat_of(pq), fat_of(po))) return;
        if (pq.tree_id < po.tree_id) keys.push_back({pq.tree_id, po.tree_id, q, o});
        else keys.push_back({po.tree_id, pq.tree_id, o, q});
    };
    if (static_cast<uint64_t>(queued.size()) * P <= 4000000ull) {
        for (uint32_t q : queued) for (uint32_t o = 0; o < P; ++o) test_pair(q, o);
    } else {
        // Same pair set as the all-pairs loop, found through a uniform grid (the oracle only has to finish in seconds at
        // BASELINE sizes; the keys are sorted and made unique below either way). Closed intervals: two boxes that overlap share
        // a cell because the cell index is a monotone function of the coordinate. Boxes spanning many cells go to a side list.
        std::vector<float> ext(P);
        float x0 = kInf, y0 = kInf;
        for (uint32_t i = 0; i < P; ++i) {
            const AABB a = fat_of(w.proxies[i].p);
            ext[i] = std::max(a.maxx - a.minx, a.maxy - a.miny);
            x0 = std::min(x0, a.minx); y0 = std::min(y0, a.miny);
        }
        std::vector<float> tmp = ext;
        std::nth_element(tmp.begin(), tmp.begin() + P / 2, tmp.end());
        const float cell = std::max(tmp[P / 2] * 2.0f, 1e-3f);
        auto ci = [&](float v, float o) { return static_cast<int64_t>(std::floor((v - o) / cell)); };
        std::vector<uint32_t> large;
        std::vector<std::pair<uint64_t, uint32_t>> cells;
        auto span = [&](uint32_t i, int64_t& cx0, int64_t& cx1, int64_t& cy0, int64_t& cy1) {
            const AABB a = fat_of(w.proxies[i].p);
            cx0 = ci(a.minx, x0); cx1 = ci(a.maxx, x0); cy0 = ci(a.miny, y0); cy1 = ci(a.maxy, y0);
            return (cx1 - cx0 + 1) * (cy1 - cy0 + 1) <= 16;
        };
        auto ckey = [](int64_t cx, int64_t cy) { return (static_cast<uint64_t>(cx) << 32) ^ (static_cast<uint64_t>(cy) & 0xFFFFFFFFull); };
        std::vector<char> isLarge(P, 0);
        for (uint32_t i = 0; i < P; ++i) {
            int64_t cx0, cx1, cy0, cy1;
            if (!span(i, cx0, cx1, cy0, cy1)) { large.push_back(i); isLarge[i] = 1; continue; }
            for (int64_t cx = cx0; cx <= cx1; ++cx) for (int64_t cy = cy0; cy <= cy1; ++cy) cells.emplace_back(ckey(cx, cy), i);
        }
        std::sort(cells.begin(), cells.end());
        for (uint32_t q : queued) {
            if (isLarge[q]) { for (uint32_t o = 0; o < P; ++o) test_pair(q, o); continue; }
            int64_t cx0, cx1, cy0, cy1;
            span(q, cx0, cx1, cy0, cy1);
            for (int64_t cx = cx0; cx <= cx1; ++cx) for (int64_t cy = cy0; cy <= cy1; ++cy) {
                const uint64_t k = ckey(cx, cy);
                for (auto it = std::lower_bound(cells.begin(), cells.end(), std::make_pair(k, 0u)); it != cells.end() && it->first == k; ++it) test_pair(q, it->second);
            }
            for (uint32_t o : large) test_pair(q, o);
        }
    }
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.mn != b.mn ? a.mn < b.mn : a.mx < b.mx; });
    keys.erase(std::unique(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.mn == b.mn && a.mx == b.mx; }), keys.end());
    uint32_t added = 0;
    for (const Key& k : keys) {
        const pr_proxy& pa = w.proxies[k.pxMin].p;
        const pr_proxy& pb = w.proxies[k.pxMax].p;
        OBody& bodyA = w.bodies[pa.body];
        OBody& bodyB = w.bodies[pb.body];
        const pr_shape& shapeA = w.shapes[pa.shape];
        const pr_shape& shapeB = w.shapes[pb.shape];
        if (!(accelerable(bodyA.b) || accelerable(bodyB.b)) || !joints_allow_collision(w, pa.body, pb.body) ||
            !filters_allow(shapeA, shapeB)) continue;
        const auto& search = (bodyA.contacts.size() < bodyB.contacts.size()) ? bodyA.contacts : bodyB.contacts;
        bool exists = false;
        for (uint32_t id : search) {
            const pr_contact& c = w.contacts[id].c;
            if (c.key_min == k.mn && c.key_max == k.mx) { exists = true; break; }
        }
        if (exists) continue;
        uint32_t id;
        if (!w.freeContacts.empty()) { id = w.freeContacts.back(); w.freeContacts.pop_back(); }
        else { id = static_cast<uint32_t>(w.contacts.size()); w.contacts.emplace_back(); }
        OContact& oc = w.contacts[id];
        std::memset(&oc, 0, sizeof(oc));
        oc.pxA = k.pxMin; oc.pxB = k.pxMax;
        pr_contact& c = oc.c;
        c.id = id; c.key_min = k.mn; c.key_max = k.mx;
        c.body_a = pa.body; c.shape_a = pa.shape; c.child_a = pa.child;
        c.body_b = pb.body; c.shape_b = pb.shape; c.child_b = pb.child;
        c.flags = PR_CONTACT_ENABLED | PR_CONTACT_DIRTY; // Contact ctor, Contact.hpp:343-349
        c.toi = 1.0f;
        if (impenetrable(bodyA.b) || impenetrable(bodyB.b)) c.flags |= PR_CONTACT_IMPENETRABLE;
        if (shapeA.is_sensor || shapeB.is_sensor) c.flags |= PR_CONTACT_SENSOR;
        c.friction = std::sqrt(shapeA.friction * shapeB.friction);              // MixFriction, Contact.hpp:57
        c.restitution = (shapeA.restitution > shapeB.restitution) ? shapeA.restitution : shapeB.restitution;
        oc.m.cf0 = oc.m.cf1 = CF_DEFAULT;
        w.order.push_back(id);
        bodyA.contacts.push_back(id);
        bodyB.contacts.push_back(id);
        if (!(c.flags & PR_CONTACT_SENSOR)) {
            if (speedable(bodyA.b)) bodyA.b.flags |= PR_BODY_AWAKE;
            if (speedable(bodyB.b)) bodyB.b.flags |= PR_BODY_AWAKE;
        }
        update_contact(w, id, conf);
        ++added;
    }
    return added;
}

// CreateProxies, AabbTreeWorld.cpp:557-579
uint32_t create_proxies(orc_world& w, uint32_t bodyId, uint32_t shapeId, const pr_step_conf& conf)
{
    OBody& body = w.bodies[bodyId];
    const Xf xf0 = get_xf(body_pos0(body.b), body_lc(body.b));
    const Xf xf1 = body_xf(body.b);
    const V2 displacement = conf.displace_multiplier * (xf1.p - xf0.p);
    const uint32_t n = w.shapes[shapeId].child_count;
    for (uint32_t ci = 0; ci < n; ++ci) {
        const AABB a = fatten_displace(compute_aabb(w.child(shapeId, ci), xf0, xf1), conf.aabb_extension, displacement);
        OProxy op;
        const uint32_t k = static_cast<uint32_t>(w.proxies.size());
        if (!w.treeFree.empty()) { op.p.tree_id = w.treeFree.back(); w.treeFree.pop_back(); }
        else { op.p.tree_id = w.treeNextFresh; w.treeNextFresh += (k == 0) ? 1u : 2u; }
        op.p.body = bodyId; op.p.shape = shapeId; op.p.child = ci;
        set_fat(op.p, a);
        w.proxies.push_back(op);
        body.proxies.push_back(k);
        w.proxiesForContacts.push_back(k);
    }
    return n;
}

// Synchronize, AabbTreeWorld.cpp:2564-2585
uint32_t synchronize(orc_world& w, uint32_t bodyId, Xf xf0, Xf xf1, const pr_step_conf& conf)
{
    uint32_t updated = 0;
    const V2 displacement = conf.displace_multiplier * (xf1.p - xf0.p);
    for (uint32_t px : w.bodies[bodyId].proxies) {
        pr_proxy& p = w.proxies[px].p;
        const AABB a = compute_aabb(w.child(p.shape, p.child), xf0, xf1);
        if (!contains(fat_of(p), a)) {
            set_fat(p, fatten_displace(a, conf.aabb_extension, displacement));
            w.proxiesForContacts.push_back(px);
            ++updated;
        }
    }
    return updated;
}

// ---------------- solver pieces ----------------

// GetVelocity(body,h) Body.cpp:290-310 ; Cap Velocity.cpp:27-46 ; GetBodyConstraint BodyConstraint.hpp:154
BC get_body_constraint(const pr_body& b, float h, const pr_step_conf& conf)
{
    Vel v{{b.vel_x, b.vel_y}, b.vel_w};
    if (accelerable(b)) {
        v.l = v.l + h * V2{b.acc_x, b.acc_y};
        v.a += h * b.acc_w;
        v.l = v.l / (1 + h * b.lin_damp);
        v.a /= (1 + h * b.ang_damp);
    }
    const V2 translation = h * v.l;
    const float lsq = magsq(translation);
    if (lsq > square(conf.max_translation)) {
        const float ratio = conf.max_translation / std::sqrt(lsq);
        v.l = v.l * ratio;
    }
    const float absRot = std::fabs(h * v.a);
    if (absRot > conf.max_rotation) {
        const float ratio = conf.max_rotation / absRot;
        v.a *= ratio;
    }
    return {b.inv_mass, b.inv_rot_i, body_lc(b), body_pos1(b), v};
}

// GetContactRelVelocity, d2/Math.cpp:99-102
inline V2 rel_velocity(Vel vA, V2 relA, Vel vB, V2 relB)
{
    const V2 velBrot = rev_perp(relB) * (vB.a / 1.0f);
    const V2 velArot = rev_perp(relA) * (vA.a / 1.0f);
    return (vB.l + velBrot) - (vA.l + velArot);
}

struct WMPoint { V2 p; float jn, jt, sep; };
struct WM { Rot normal; uint32_t count; WMPoint pt[2]; };

// GetWorldManifold, WorldManifold.cpp:37-130
WM world_manifold(const pr_manifold& m, Xf xfA, float rA, Xf xfB, float rB)
{
    WM w{};
    const V2 lp{m.lp_x, m.lp_y};
    const V2 mp[2] = {{m.p0_x, m.p0_y}, {m.p1_x, m.p1_y}};
    const float jn[2] = {m.jn0, m.jn1}, jt[2] = {m.jt0, m.jt1};
    if (m.type == PR_MANIFOLD_CIRCLES) {
        const V2 pointA = transform(lp, xfA);
        const V2 pointB = transform(mp[0], xfB);
        const Rot normal = unit_vector(pointB - pointA, Rot{1.0f, 0.0f});
        const V2 cA = pointA + (rA * normal);
        const V2 cB = pointB - (rB * normal);
        w.normal = normal; w.count = 1;
        w.pt[0] = {(cA + cB) / 2.0f, jn[0], jt[0], dot(cB - cA, normal)};
    } else if (m.type == PR_MANIFOLD_FACE_A) {
        const Rot normal = rot_rotate(Rot{m.ln_x, m.ln_y}, xfA.q);
        const V2 planePoint = transform(lp, xfA);
        w.normal = normal; w.count = m.count;
        for (uint32_t i = 0; i < m.count; ++i) {
            const V2 clipPoint = transform(mp[i], xfB);
            const V2 cA = clipPoint + (rA - dot(clipPoint - planePoint, normal)) * normal;
            const V2 cB = clipPoint - (rB * normal);
            w.pt[i] = {(cA + cB) / 2.0f, jn[i], jt[i], dot(cB - cA, normal)};
        }
    } else if (m.type == PR_MANIFOLD_FACE_B) {
        const Rot normal = rot_rotate(Rot{m.ln_x, m.ln_y}, xfB.q);
        const V2 planePoint = transform(lp, xfB);
        w.normal = -normal; w.count = m.count;
        for (uint32_t i = 0; i < m.count; ++i) {
            const V2 clipPoint = transform(mp[i], xfA);
            const V2 cB = clipPoint + (rB - dot(clipPoint - planePoint, normal)) * normal;
            const V2 cA = clipPoint - (rA * normal);
            w.pt[i] = {(cA + cB) / 2.0f, jn[i], jt[i], dot(cA - cB, normal)};
        }
    }
    return w;
}

// VelocityConstraint ctor + GetPoint + ComputeK, VelocityConstraint.cpp:38-196
VC make_vc(float friction, float restitution, float tangentSpeed, const WM& wm, uint32_t bA, uint32_t bB,
           const std::vector<BC>& bodies, float dtRatio, float velocityThreshold, bool blockSolve)
{
    VC vc{};
    vc.bodyA = bA; vc.bodyB = bB; vc.normal = wm.normal;
    vc.friction = friction; vc.restitution = restitution; vc.tangentSpeed = tangentSpeed;
    const BC& A = bodies[bA];
    const BC& B = bodies[bB];
    const float invMass = A.invM + B.invM;
    const Rot tangent = fwd_perp(vc.normal);
    for (uint32_t j = 0; j < wm.count; ++j) {
        VCP& p = vc.pt[j];
        const V2 relA = wm.pt[j].p - A.pos.l;
        const V2 relB = wm.pt[j].p - B.pos.l;
        p.normalImpulse = wm.pt[j].jn * dtRatio;
        p.tangentImpulse = wm.pt[j].jt * dtRatio;
        const V2 dv = rel_velocity(A.vel, relA, B.vel, relB);
        const float vn = dot(dv, vc.normal);
        p.velocityBias = (vn < -velocityThreshold) ? -restitution * vn : 0.0f;
        p.relA = relA; p.relB = relB;
        {
            const float a = A.invI * square(cross(relA, vc.normal)) / 1.0f;
            const float b = B.invI * square(cross(relB, vc.normal)) / 1.0f;
            const float value = invMass + a + b;
            p.normalMass = (value != 0.0f) ? 1.0f / value : 0.0f;
        }
        {
            const float a = A.invI * square(cross(relA, tangent)) / 1.0f;
            const float b = B.invI * square(cross(relB, tangent)) / 1.0f;
            const float value = invMass + a + b;
            p.tangentMass = (value != 0.0f) ? 1.0f / value : 0.0f;
        }
        ++vc.count;
    }
    if (blockSolve && vc.count == 2) {
        const float rnA0 = cross(vc.pt[0].relA, vc.normal), rnB0 = cross(vc.pt[0].relB, vc.normal);
        const float rnA1 = cross(vc.pt[1].relA, vc.normal), rnB1 = cross(vc.pt[1].relB, vc.normal);
        const float iA0 = (A.invI * square(rnA0)) / 1.0f, iA1 = (A.invI * square(rnA1)) / 1.0f;
        const float iB0 = (B.invI * square(rnB0)) / 1.0f, iB1 = (B.invI * square(rnB1)) / 1.0f;
        const float iA = (A.invI * rnA0 * rnA1) / 1.0f, iB = (B.invI * rnB0 * rnB1) / 1.0f;
        const float k00 = invMass + iA0 + iB0, k11 = invMass + iA1 + iB1, k01 = invMass + iA + iB;
        const float k00sq = square(k00), k00k11 = k00 * k11, k01sq = square(k01);
        const float kdiff = k00k11 - k01sq;
        if (k00sq < 1000.0f * kdiff) {
            vc.K[0] = k00; vc.K[1] = k11; vc.K[2] = k01;
            // Invert(Matrix22), Math.hpp:430-442 on {{k00,k01},{k01,k11}}
            const float cp = k00 * k11 - k01 * k01; // Cross(col0, col1)
            if (!almost_zero(cp)) {
                const float inv = 1 / cp;
                // normalMass = {{k11*inv, -k01*inv},{-k01*inv, k00*inv}}; stored (m00, m11, m01)
                vc.normalMass[0] = k11 * inv; vc.normalMass[1] = k00 * inv; vc.normalMass[2] = -k01 * inv;
            }
        } else {
            vc.pt[1] = VCP{};
            vc.count = 1;
        }
    }
    return vc;
}

inline void apply_pair(BC& A, BC& B, V2 P, V2 relA, V2 relB)
{
    // newVelA -= Velocity{invMassA * P, invRotInertiaA * LA}; newVelB += ...
    const float LA = cross(relA, P) / 1.0f, LB = cross(relB, P) / 1.0f;
    A.vel.l = A.vel.l - A.invM * P; A.vel.a -= A.invI * LA;
    B.vel.l = B.vel.l + B.invM * P; B.vel.a += B.invI * LB;
}

// WarmStartVelocities, AabbTreeWorld.cpp:222-273
void warm_start(const VC& vc, std::vector<BC>& bodies)
{
    BC& A = bodies[vc.bodyA];
    BC& B = bodies[vc.bodyB];
    Vel dA{{0, 0}, 0}, dB{{0, 0}, 0};
    const Rot tangent = fwd_perp(vc.normal);
    for (uint32_t j = 0; j < vc.count; ++j) {
        const VCP& p = vc.pt[j];
        const V2 P = p.normalImpulse * vc.normal + p.tangentImpulse * tangent;
        const float LA = cross(p.relA, P) / 1.0f, LB = cross(p.relB, P) / 1.0f;
        dA.l = dA.l - A.invM * P; dA.a -= A.invI * LA;
        dB.l = dB.l + B.invM * P; dB.a += B.invI * LB;
    }
    A.vel.l = A.vel.l + dA.l; A.vel.a += dA.a;
    B.vel.l = B.vel.l + dB.l; B.vel.a += dB.a;
}

// SolveTangentConstraint, ContactSolver.cpp:395-455
float solve_tangent(VC& vc, std::vector<BC>& bodies)
{
    float maxInc = 0;
    const Rot direction = fwd_perp(vc.normal);
    BC& A = bodies[vc.bodyA];
    BC& B = bodies[vc.bodyB];
    BC nA = A, nB = B;
    auto proc = [&](uint32_t i) {
        VCP& p = vc.pt[i];
        const V2 closing = rel_velocity(nA.vel, p.relA, nB.vel, p.relB);
        const float directionalVel = vc.tangentSpeed - dot(closing, direction);
        const float lambda = p.tangentMass * directionalVel;
        const float maxImpulse = vc.friction * p.normalImpulse;
        const float oldImpulse = p.tangentImpulse;
        const float newImpulse = std::clamp(oldImpulse + lambda, -maxImpulse, maxImpulse);
        const float inc = newImpulse - oldImpulse;
        const V2 P = inc * direction;
        apply_pair(nA, nB, P, p.relA, p.relB);
        maxInc = std::max(maxInc, std::fabs(inc));
        p.tangentImpulse = oldImpulse + inc;
    };
    if (vc.count == 2) proc(1);
    proc(0);
    A.vel = nA.vel; B.vel = nB.vel;
    return maxInc;
}

// SeqSolveNormalConstraint, ContactSolver.cpp:337-393
float solve_normal_seq(VC& vc, std::vector<BC>& bodies)
{
    float maxInc = 0;
    const Rot direction = vc.normal;
    BC& A = bodies[vc.bodyA];
    BC& B = bodies[vc.bodyB];
    BC nA = A, nB = B;
    auto proc = [&](uint32_t i) {
        VCP& p = vc.pt[i];
        const V2 closing = rel_velocity(nA.vel, p.relA, nB.vel, p.relB);
        const float directionalVel = dot(closing, direction);
        const float lambda = p.normalMass * (p.velocityBias - directionalVel);
        const float oldImpulse = p.normalImpulse;
        const float newImpulse = std::max(oldImpulse + lambda, 0.0f);
        const float inc = newImpulse - oldImpulse;
        const V2 P = inc * direction;
        apply_pair(nA, nB, P, p.relA, p.relB);
        maxInc = std::max(maxInc, std::fabs(inc));
        p.normalImpulse = oldImpulse + inc;
    };
    if (vc.count == 2) proc(1);
    proc(0);
    A.vel = nA.vel; B.vel = nB.vel;
    return maxInc;
}

// BlockSolveUpdate + GetVelocityDelta, ContactSolver.cpp:69-111
float block_update(VC& vc, float n0, float n1, std::vector<BC>& bodies)
{
    BC& A = bodies[vc.bodyA];
    BC& B = bodies[vc.bodyB];
    const float i0 = n0 - vc.pt[0].normalImpulse, i1 = n1 - vc.pt[1].normalImpulse;
    const V2 P0 = i0 * vc.normal, P1 = i1 * vc.normal;
    const V2 P = P0 + P1;
    const float LA = (cross(vc.pt[0].relA, P0) + cross(vc.pt[1].relA, P1)) / 1.0f;
    const float LB = (cross(vc.pt[0].relB, P0) + cross(vc.pt[1].relB, P1)) / 1.0f;
    const V2 dAl = -(A.invM * P); const float dAa = -(A.invI * LA);
    const V2 dBl = B.invM * P;    const float dBa = B.invI * LB;
    A.vel.l = A.vel.l + dAl; A.vel.a += dAa;
    B.vel.l = B.vel.l + dBl; B.vel.a += dBa;
    vc.pt[0].normalImpulse = n0; vc.pt[1].normalImpulse = n1;
    return std::max(std::fabs(n0), std::fabs(n1));
}

// BlockSolveNormalConstraint, ContactSolver.cpp:113-335
float solve_normal_block(VC& vc, std::vector<BC>& bodies)
{
    const BC& A = bodies[vc.bodyA];
    const BC& B = bodies[vc.bodyB];
    const V2 dv0 = rel_velocity(A.vel, vc.pt[0].relA, B.vel, vc.pt[0].relB);
    const V2 dv1 = rel_velocity(A.vel, vc.pt[1].relA, B.vel, vc.pt[1].relB);
    const float vn0 = dot(dv0, vc.normal), vn1 = dot(dv1, vc.normal);
    const float b0 = vn0 - vc.pt[0].velocityBias, b1 = vn1 - vc.pt[1].velocityBias;
    const float k00 = vc.K[0], k11 = vc.K[1], k01 = vc.K[2];
    const float a0 = vc.pt[0].normalImpulse, a1 = vc.pt[1].normalImpulse;
    // b - Transform(a, K): K rows {k00,k01},{k01,k11}; element = 0 + m[b][0]*v0 + m[b][1]*v1
    float t0 = 0.0f; t0 += k00 * a0; t0 += k01 * a1;
    float t1 = 0.0f; t1 += k01 * a0; t1 += k11 * a1;
    const float bp0 = b0 - t0, bp1 = b1 - t1;
    const float m00 = vc.normalMass[0], m11 = vc.normalMass[1], m01 = vc.normalMass[2];
    { // case 1
        float x0 = 0.0f; x0 += m00 * bp0; x0 += m01 * bp1;
        float x1 = 0.0f; x1 += m01 * bp0; x1 += m11 * bp1;
        x0 = -x0; x1 = -x1;
        if (x0 >= 0.0f && x1 >= 0.0f) return block_update(vc, x0, x1, bodies);
    }
    { // case 2
        const float x0 = -vc.pt[0].normalMass * bp0;
        const float vn2 = k01 * x0 + bp1;
        if (x0 >= 0.0f && vn2 >= 0.0f) return block_update(vc, x0, 0.0f, bodies);
    }
    { // case 3
        const float x1 = -vc.pt[1].normalMass * bp1;
        const float vn1b = k01 * x1 + bp0;
        if (x1 >= 0.0f && vn1b >= 0.0f) return block_update(vc, 0.0f, x1, bodies);
    }
    if (bp0 >= 0.0f && bp1 >= 0.0f) return block_update(vc, 0.0f, 0.0f, bodies);
    return 0.0f;
}

// GaussSeidel::SolveVelocityConstraint, ContactSolver.cpp:481-490
float solve_velocity_constraint(VC& vc, std::vector<BC>& bodies)
{
    float maxInc = 0;
    maxInc = std::max(maxInc, solve_tangent(vc, bodies));
    const bool kZero = (vc.K[0] == 0.0f && vc.K[1] == 0.0f && vc.K[2] == 0.0f);
    if (vc.count == 1 || kZero) maxInc = std::max(maxInc, solve_normal_seq(vc, bodies));
    else maxInc = std::max(maxInc, solve_normal_block(vc, bodies));
    return maxInc;
}

struct PSM { Rot normal; V2 point; float separation; };
// GetPSM, PositionSolverManifold.cpp:35-120
PSM get_psm(const pr_manifold& m, uint32_t index, Xf xfA, Xf xfB)
{
    const V2 lp{m.lp_x, m.lp_y};
    const V2 plp = index == 0 ? V2{m.p0_x, m.p0_y} : V2{m.p1_x, m.p1_y};
    const Rot ln{m.ln_x, m.ln_y};
    if (m.type == PR_MANIFOLD_CIRCLES) {
        const V2 pointA = transform(lp, xfA), pointB = transform(plp, xfB);
        const V2 delta = pointB - pointA;
        const Rot normal = unit_vector(delta, Rot{0.0f, 0.0f});
        const V2 mid = (pointA + pointB) / 2.0f;
        return {normal, mid, dot(delta, normal)};
    }
    if (m.type == PR_MANIFOLD_FACE_A) {
        const V2 planePoint = transform(lp, xfA);
        const Rot normal = rot_rotate(ln, xfA.q);
        const V2 clipPoint = transform(plp, xfB);
        return {normal, clipPoint, dot(clipPoint - planePoint, normal)};
    }
    const V2 planePoint = transform(lp, xfB);
    const Rot normal = rot_rotate(ln, xfB.q);
    const V2 clipPoint = transform(plp, xfA);
    return {-normal, clipPoint, dot(clipPoint - planePoint, normal)};
}

struct PosSol { Pos a, b; float minSep; };
// GaussSeidel::SolvePositionConstraint, ContactSolver.cpp:494-643 (moveA = moveB = true)
PosSol solve_position_constraint(const PC& pc, const std::vector<BC>& bodies, float resolutionRate,
                                 float linearSlop, float maxLinearCorrection)
{
    const BC& A = bodies[pc.bodyA];
    const BC& B = bodies[pc.bodyB];
    const float invMassA = A.invM, invIA = A.invI, invMassB = B.invM, invIB = B.invI;
    const float invMassTotal = invMassA + invMassB;
    const float totalRadius = pc.totalRadius;
    struct Delta { Pos a, b; float sep; };
    auto solver = [&](const PSM& psm, V2 pA, V2 pB) -> Delta {
        const float separation = psm.separation - totalRadius;
        const V2 rA = psm.point - pA, rB = psm.point - pB;
        // GetEffectiveInvMass: invRotI * Square(Cross(p, q) / Radian), d2/Math.hpp:292-296
        const float K = invMassTotal + invIA * square(cross(rA, psm.normal) / 1.0f) + invIB * square(cross(rB, psm.normal) / 1.0f);
        const float C = -std::clamp(resolutionRate * (separation + linearSlop), -maxLinearCorrection, 0.0f);
        const V2 P = (psm.normal * C) / K;
        const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
        return {Pos{-(invMassA * P), -(invIA * LA)}, Pos{invMassB * P, invIB * LB}, separation};
    };
    auto add = [](Pos p, Pos d) { return Pos{p.l + d.l, p.a + d.a}; };
    Pos posA = A.pos, posB = B.pos;
    const V2 lcA = A.lc, lcB = B.lc;
    if (pc.m.count == 1) {
        const PSM psm0 = get_psm(pc.m, 0, get_xf(posA, lcA), get_xf(posB, lcB));
        const Delta s = solver(psm0, posA.l, posB.l);
        return {add(posA, s.a), add(posB, s.b), 0.0f + s.sep};
    }
    if (pc.m.count == 2) {
        const Xf xfA = get_xf(posA, lcA), xfB = get_xf(posB, lcB);
        const PSM psm0 = get_psm(pc.m, 0, xfA, xfB), psm1 = get_psm(pc.m, 1, xfA, xfB);
        if (almost_equal(psm0.separation, psm1.separation)) {
            const Delta s0 = solver(psm0, posA.l, posB.l), s1 = solver(psm1, posA.l, posB.l);
            return {add(add(posA, s0.a), s1.a), add(add(posB, s0.b), s1.b), s0.sep};
        }
        if (psm0.separation < psm1.separation) {
            const Delta s0 = solver(psm0, posA.l, posB.l);
            posA = add(posA, s0.a); posB = add(posB, s0.b);
            const PSM p1 = get_psm(pc.m, 1, get_xf(posA, lcA), get_xf(posB, lcB));
            const Delta s1 = solver(p1, posA.l, posB.l);
            posA = add(posA, s1.a); posB = add(posB, s1.b);
            return {posA, posB, s0.sep};
        }
        if (psm1.separation < psm0.separation) {
            const Delta s1 = solver(psm1, posA.l, posB.l);
            posA = add(posA, s1.a); posB = add(posB, s1.b);
            const PSM p0 = get_psm(pc.m, 0, get_xf(posA, lcA), get_xf(posB, lcB));
            const Delta s0 = solver(p0, posA.l, posB.l);
            posA = add(posA, s0.a); posB = add(posB, s0.b);
            return {posA, posB, s1.sep};
        }
    }
    return {posA, posB, kInf};
}

struct Island { std::vector<uint32_t> bodies, contacts, joints; };

// IsUnderActive, d2/Math.hpp:283-289
inline bool under_active(const pr_body& b, float linTol, float angTol)
{
    const float lin = magsq(V2{b.vel_x, b.vel_y});
    const float ang = square(b.vel_w);
    return (ang <= square(angTol)) && (lin <= square(linTol));
}

void build_constraints(orc_world& w, const Island& island, std::vector<BC>& bcs, std::vector<VC>& vcs,
                       std::vector<PC>& pcs, float h, const pr_step_conf& conf, float dtRatio)
{
    for (uint32_t id : island.bodies) bcs[id] = get_body_constraint(w.bodies[id].b, h, conf);
    for (uint32_t cid : island.contacts) {
        const OContact& oc = w.contacts[cid];
        const float rA = w.radius(oc.c.shape_a, oc.c.child_a), rB = w.radius(oc.c.shape_b, oc.c.child_b);
        pcs.push_back({oc.m, oc.c.body_a, oc.c.body_b, rA + rB});
    }
    for (uint32_t cid : island.contacts) {
        const OContact& oc = w.contacts[cid];
        const float rA = w.radius(oc.c.shape_a, oc.c.child_a), rB = w.radius(oc.c.shape_b, oc.c.child_b);
        const BC& A = bcs[oc.c.body_a];
        const BC& B = bcs[oc.c.body_b];
        const Xf xfA = get_xf(A.pos, A.lc), xfB = get_xf(B.pos, B.lc);
        const WM wm = world_manifold(oc.m, xfA, rA, xfB, rB);
        vcs.push_back(make_vc(oc.c.friction, oc.c.restitution, oc.c.tangent_speed, wm, oc.c.body_a, oc.c.body_b,
                              bcs, dtRatio, conf.velocity_threshold, conf.do_blocksolve != 0));
    }
}

// AssignImpulses, AabbTreeWorld.cpp:208-230
void assign_impulses(pr_manifold& m, const VC& vc)
{
    if (vc.count > 0) { m.jn0 = vc.pt[0].normalImpulse; m.jt0 = vc.pt[0].tangentImpulse; }
    if (vc.count > 1) { m.jn1 = vc.pt[1].normalImpulse; m.jt1 = vc.pt[1].tangentImpulse; }
}

// SolveRegIslandViaGS, AabbTreeWorld.cpp:1563-1682
IslandStats solve_reg_island(orc_world& w, const pr_step_conf& conf, const Island& island)
{
    IslandStats results;
    results.positionIters = conf.reg_position_iters;
    const float h = conf.delta_time;
    for (uint32_t id : island.bodies) { pr_body& b = w.bodies[id].b; set_pos0(b, body_pos1(b)); }
    // (BodyConstraints indexed by body id: a world-owned scratch; every body an island's constraints or joints name is set by
    // build_constraints first -- one vector per island made a 100k-island step quadratic)
    std::vector<BC>& bcs = w.bcScratch;
    if (bcs.size() < w.bodies.size()) bcs.resize(w.bodies.size());
    std::vector<VC> vcs; std::vector<PC> pcs;
    build_constraints(w, island, bcs, vcs, pcs, h, conf, conf.do_warm_start ? conf.dt_ratio : 0.0f);
    if (conf.do_warm_start) for (const VC& vc : vcs) warm_start(vc, bcs);
    orc::JointSolverConf jc{conf.delta_time, conf.dt_ratio, conf.do_warm_start != 0, conf.linear_slop, conf.angular_slop,
                            conf.max_linear_correction, conf.max_angular_correction, conf.reg_resolution_rate};
    auto gb = [&](uint32_t id) { return orc::GearBody{bcs[id].invM, bcs[id].invI, bcs[id].lc, &bcs[id].pos, &bcs[id].vel}; };
    for (uint32_t j : island.joints) {
        auto& jt = w.joints[j];
        if (PR_JOINT_TYPE(jt) == PR_JOINT_GEAR) { const auto& g = w.gears[j]; orc::gear_init_velocity(jt, g, w.jointWork[j], gb(jt.body_a), gb(jt.body_b), gb(g.body_c), gb(g.body_d), jc); continue; }
        orc::joint_init_velocity(jt, w.jointWork[j], bcs[jt.body_a].invM, bcs[jt.body_a].invI, bcs[jt.body_a].lc, bcs[jt.body_a].pos, bcs[jt.body_a].vel,
                                    bcs[jt.body_b].invM, bcs[jt.body_b].invI, bcs[jt.body_b].lc, bcs[jt.body_b].pos, bcs[jt.body_b].vel, jc);
    }
    results.velocityIters = conf.reg_velocity_iters;
    for (uint32_t i = 0; i < conf.reg_velocity_iters; ++i) {
        bool jointsOkay = true;
        for (uint32_t j : island.joints) {
            auto& jt = w.joints[j];
            if (PR_JOINT_TYPE(jt) == PR_JOINT_GEAR) { const auto& g = w.gears[j]; jointsOkay &= orc::gear_solve_velocity(jt, w.jointWork[j], gb(jt.body_a), gb(jt.body_b), gb(g.body_c), gb(g.body_d)); continue; }
            jointsOkay &= orc::joint_solve_velocity(jt, w.jointWork[j], bcs[jt.body_a].invM, bcs[jt.body_a].invI, bcs[jt.body_a].vel,
                                                       bcs[jt.body_b].invM, bcs[jt.body_b].invI, bcs[jt.body_b].vel, jc);
        }
        float newInc = 0;
        for (VC& vc : vcs) newInc = std::max(newInc, solve_velocity_constraint(vc, bcs));
        results.maxIncImpulse = std::max(results.maxIncImpulse, newInc);
        if (jointsOkay && newInc <= conf.reg_min_momentum) { results.velocityIters = i + 1; break; }
    }
    // IntegratePositions, AabbTreeWorld.cpp:176-189
    for (uint32_t id : island.bodies) {
        BC& bc = bcs[id];
        const V2 translation = h * bc.vel.l;
        const float rotation = h * bc.vel.a;
        bc.pos = Pos{bc.pos.l + translation, bc.pos.a + rotation};
    }
    for (uint32_t i = 0; i < conf.reg_position_iters; ++i) {
        float minSeparation = kInf;
        for (const PC& pc : pcs) {
            const PosSol r = solve_position_constraint(pc, bcs, conf.reg_resolution_rate, conf.linear_slop, conf.max_linear_correction);
            bcs[pc.bodyA].pos = r.a; bcs[pc.bodyB].pos = r.b;
            minSeparation = std::min(minSeparation, r.minSep);
        }
        results.minSeparation = std::min(results.minSeparation, minSeparation);
        const bool contactsOkay = minSeparation >= conf.reg_min_separation;
        bool jointsOkay = true;
        for (uint32_t j : island.joints) {
            auto& jt = w.joints[j];
            if (PR_JOINT_TYPE(jt) == PR_JOINT_GEAR) { const auto& g = w.gears[j]; jointsOkay &= orc::gear_solve_position(jt, g, gb(jt.body_a), gb(jt.body_b), gb(g.body_c), gb(g.body_d), jc); continue; }
            jointsOkay &= orc::joint_solve_position(jt, w.jointWork[j], bcs[jt.body_a].invM, bcs[jt.body_a].invI, bcs[jt.body_a].lc, bcs[jt.body_a].pos,
                                                       bcs[jt.body_b].invM, bcs[jt.body_b].invI, bcs[jt.body_b].lc, bcs[jt.body_b].pos, jc);
        }
        if (contactsOkay && jointsOkay) { results.positionIters = i + 1; results.solved = true; break; }
    }
    for (size_t i = 0; i < vcs.size(); ++i) assign_impulses(w.contacts[island.contacts[i]].m, vcs[i]);
    if (w.evMask & PR_EVENT_POST_SOLVE) // Report(...), :1671-1674
        for (size_t i = 0; i < vcs.size(); ++i) record_impulse_event(w, island.contacts[i], vcs[i], results.solved ? results.positionIters - 1u : 255u, 0u);
    for (uint32_t id : island.bodies) {
        const BC& bc = bcs[id];
        pr_body& b = w.bodies[id].b;
        b.vel_x = bc.vel.l.x; b.vel_y = bc.vel.l.y; b.vel_w = bc.vel.a;
        const Pos p1 = body_pos1(b);
        if (!(p1.l == bc.pos.l) || p1.a != bc.pos.a) { set_pos1(b, bc.pos); flag_body_contacts_dirty(w, id); }
    }
    // UpdateUnderActiveTimes, AabbTreeWorld.cpp:395-417 ; Sleepem :419-430
    float minUnderActive = kInf;
    for (uint32_t id : island.bodies) {
        pr_body& b = w.bodies[id].b;
        if (speedable(b)) {
            const bool ua = under_active(b, conf.linear_sleep_tolerance, conf.angular_sleep_tolerance);
            const bool sleepable = (b.flags & PR_BODY_AUTOSLEEP) != 0;
            const float t = (sleepable && ua) ? b.under_active_time + conf.delta_time : 0.0f;
            b.under_active_time = t;
            minUnderActive = std::min(minUnderActive, t);
        }
    }
    if (minUnderActive >= conf.min_still_time_to_sleep && results.solved) {
        for (uint32_t id : island.bodies) if (unawaken(w.bodies[id].b)) ++results.bodiesSlept;
    }
    return results;
}

// AddToIsland (DFS), AabbTreeWorld.cpp:1358-1470
void add_to_island(orc_world& w, Island& island, uint32_t seed)
{
    std::vector<uint32_t> stack{seed};
    w.islandedBodies[seed] = 1;
    while (!stack.empty()) {
        const uint32_t bodyId = stack.back();
        stack.pop_back();
        OBody& body = w.bodies[bodyId];
        island.bodies.push_back(bodyId);
        if (!speedable(body.b)) continue;
        body.b.flags |= PR_BODY_AWAKE;
        for (uint32_t cid : body.contacts) {
            if (w.islandedContacts[cid]) continue;
            const pr_contact& c = w.contacts[cid].c;
            if ((c.flags & PR_CONTACT_ENABLED) && (c.flags & PR_CONTACT_TOUCHING) && !(c.flags & PR_CONTACT_SENSOR)) {
                const uint32_t other = (c.body_a == bodyId) ? c.body_b : c.body_a;
                island.contacts.push_back(cid);
                w.islandedContacts[cid] = 1;
                if (!w.islandedBodies[other]) { w.islandedBodies[other] = 1; stack.push_back(other); }
            }
        }
        for (uint32_t jid : body.joints) {
            if (w.islandedJoints[jid]) continue;
            const auto& jt = w.joints[jid];
            const uint32_t other = (jt.body_a == bodyId) ? jt.body_b : jt.body_a;
            if (enabled(w.bodies[other].b)) {
                w.islandedJoints[jid] = 1;
                island.joints.push_back(jid);
                if (!w.islandedBodies[other]) { w.islandedBodies[other] = 1; stack.push_back(other); }
            }
        }
    }
}

// SolveReg, AabbTreeWorld.cpp:1472-1561
void solve_reg(orc_world& w, const pr_step_conf& conf, pr_step_stats& stats,
               const std::unordered_map<uint32_t, uint32_t>* rank)
{
    w.islandedBodies.assign(w.bodies.size(), 0);
    w.islandedContacts.assign(w.contacts.size(), 0);
    w.islandedJoints.assign(w.joints.size(), 0);
    for (uint32_t bodyId : w.bodyOrder) { // (m_bodies order, :1493)
        if (w.islandedBodies[bodyId]) continue;
        const pr_body& b = w.bodies[bodyId].b;
        if (!(awake(b) && enabled(b))) continue;
        ++stats.reg.islands_found;
        Island island;
        add_to_island(w, island, bodyId);
        if (rank) {
            std::stable_sort(island.contacts.begin(), island.contacts.end(), [&](uint32_t a, uint32_t b2) {
                const auto ia = rank->find(a), ib = rank->find(b2);
                const uint32_t ra = ia == rank->end() ? 0xFFFFFFFFu : ia->second;
                const uint32_t rb = ib == rank->end() ? 0xFFFFFFFFu : ib->second;
                return ra < rb;
            });
            // the device walks joints in id order (one thread when joints share a body, pr_joint.cuh); the reference in the island's
            // discovery order (AddJointsToIsland :1443-1466). As the device's twin the restatement takes the device's order, like for contacts.
            std::sort(island.joints.begin(), island.joints.end());
        }
        stats.reg.max_island_bodies = std::max(stats.reg.max_island_bodies, static_cast<uint32_t>(island.bodies.size()));
        for (uint32_t id : island.bodies) if (!speedable(w.bodies[id].b)) w.islandedBodies[id] = 0;
        const IslandStats r = solve_reg_island(w, conf, island);
        // Update(RegStepStats&, IslandStats), StepStats.cpp:29-38
        stats.reg.max_inc_impulse = std::max(stats.reg.max_inc_impulse, r.maxIncImpulse);
        stats.reg.min_separation = std::min(stats.reg.min_separation, r.minSeparation);
        if (r.solved) ++stats.reg.islands_solved;
        stats.reg.sum_pos_iters += r.positionIters;
        stats.reg.sum_vel_iters += r.velocityIters;
        stats.reg.bodies_slept += r.bodiesSlept;
    }
    for (uint32_t bodyId : w.bodyOrder) {
        if (!w.islandedBodies[bodyId]) continue;
        const pr_body& b = w.bodies[bodyId].b;
        if (speedable(b)) {
            stats.reg.proxies_moved += synchronize(w, bodyId, get_xf(body_pos0(b), body_lc(b)), body_xf(b), conf);
        }
    }
    w.islandedBodies.assign(w.bodies.size(), 0);
    w.islandedContacts.assign(w.contacts.size(), 0);
    w.islandedJoints.assign(w.joints.size(), 0);
    update_contacts(w, conf, stats.reg.contacts_updated, stats.reg.contacts_skipped);
    stats.reg.contacts_added = find_and_add_contacts(w, conf);
}

#include "orc_toi_phase.inc"

} // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

orc_world* orc_create(void) { return new orc_world(); }
void orc_destroy(orc_world* w) { delete w; }
const char* orc_last_error(const orc_world* w) { return w->error.c_str(); }

void orc_step_conf_default(pr_step_conf* c)
{
    // StepConf defaults, Settings.hpp:92-256 / StepConf.hpp:163-391 (pinned by test vs ref_step_conf_default)
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

int orc_create_shapes(orc_world* w, uint32_t n, const pr_shape* shapes, const pr_child* children,
                      uint32_t n_children, const pr_vec2* vertices, const pr_vec2* normals, uint32_t n_vertices)
{
    const uint32_t cbase = static_cast<uint32_t>(w->children.size());
    const uint32_t vbase = static_cast<uint32_t>(w->verts.size());
    for (uint32_t i = 0; i < n; ++i) {
        pr_shape s = shapes[i]; s.child_offset += cbase;
        if (!w->shapeFree.empty()) { const uint32_t id = w->shapeFree.back(); w->shapeFree.pop_back(); w->shapeDead[id] = 0; w->shapes[id] = s; } // ObjectPool::Allocate
        else w->shapes.push_back(s);
    }
    for (uint32_t i = 0; i < n_children; ++i) { pr_child c = children[i]; c.vertex_offset += vbase; w->children.push_back(c); }
    w->verts.insert(w->verts.end(), vertices, vertices + n_vertices);
    w->normals.insert(w->normals.end(), normals, normals + n_vertices);
    return 0;
}

int orc_create_bodies(orc_world* w, uint32_t n, const pr_body* bodies, const uint32_t* shape_offsets,
                      const uint32_t* shape_ids)
{
    for (uint32_t i = 0; i < n; ++i) {
        OBody ob;
        ob.b = bodies[i];
        const bool reuse = !w->bodyFree.empty(); // ObjectPool::Allocate: the last freed id first
        const uint32_t id = reuse ? w->bodyFree.back() : static_cast<uint32_t>(w->bodies.size());
        if (reuse) { w->bodyFree.pop_back(); w->bodyDead[id] = 0; }
        for (uint32_t k = shape_offsets[i]; k < shape_offsets[i + 1]; ++k) {
            ob.shapes.push_back(shape_ids[k]);
            if (enabled(ob.b)) w->fixturesForProxies.emplace_back(id, shape_ids[k]); // CreateBody :1046-1052
        }
        if (reuse) w->bodies[id] = ob; else w->bodies.push_back(ob);
        w->bodyOrder.push_back(id);
    }
    return 0;
}

// FlagForFiltering, AabbTreeWorld.cpp:436-451
static void flag_pair_for_filtering(orc_world& w, uint32_t a, uint32_t b)
{
    for (uint32_t id : w.bodies[b].contacts) {
        pr_contact& c = w.contacts[id].c;
        const uint32_t other = (c.body_a == b) ? c.body_b : c.body_a;
        if (other == a) { c.flags |= PR_CONTACT_FILTER; w.needsFiltering = true; }
    }
}
int orc_create_revolute_joints(orc_world* w, uint32_t n, const pr_revolute_joint* joints)
{
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t id;
        if (!w->freeJoints.empty()) { id = w->freeJoints.back(); w->freeJoints.pop_back(); w->joints[id] = joints[i]; }
        else { id = static_cast<uint32_t>(w->joints.size()); w->joints.push_back(joints[i]); w->jointWork.emplace_back(); }
        w->bodies[joints[i].body_a].joints.push_back(id);
        if (joints[i].body_b != joints[i].body_a) w->bodies[joints[i].body_b].joints.push_back(id); // (target joint: one body, :1164-1169)
        // Add(JointID, !collideConnected) :1159-1175: creating a joint does not wake its bodies; the pair's contacts are flagged for filtering
        if (!(joints[i].collide_connected & 1u)) flag_pair_for_filtering(*w, joints[i].body_a, joints[i].body_b);
    }
    return 0;
}

// SetShape with unchanged geometry and filter, AabbTreeWorld.cpp:1284-1356
int orc_set_shape_material(orc_world* w, uint32_t shape, float friction, float restitution, float density, uint32_t is_sensor)
{
    pr_shape& cur = w->shapes[shape];
    for (OBody& ob : w->bodies) // :1297-1301 (m_bodyBuffer order)
        if (std::find(ob.shapes.begin(), ob.shapes.end(), shape) != ob.shapes.end()) set_awake(ob.b);
    if ((cur.is_sensor != 0) != (is_sensor != 0) || cur.friction != friction || cur.restitution != restitution) { // :1346-1354
        for (uint32_t id : w->order) {
            pr_contact& c = w->contacts[id].c;
            if (c.shape_a == shape || c.shape_b == shape) { c.flags |= PR_CONTACT_DIRTY; set_awake(w->bodies[c.body_a].b); set_awake(w->bodies[c.body_b].b); }
        }
    }
    cur.friction = friction; cur.restitution = restitution; cur.density = density; cur.is_sensor = is_sensor ? 1 : 0;
    return 0;
}

// SetBody with added shapes, AabbTreeWorld.cpp:2760-2786: Append(m_fixturesForProxies, id, added)
int orc_attach_shapes(orc_world* w, uint32_t body, uint32_t n, const uint32_t* shape_ids)
{
    const bool on = enabled(w->bodies[body].b); // (a disabled body has no proxies: GetOldAndNewShapeIDs :677-680)
    for (uint32_t i = 0; i < n; ++i) { w->bodies[body].shapes.push_back(shape_ids[i]); if (on) w->fixturesForProxies.emplace_back(body, shape_ids[i]); }
    return 0;
}

int orc_destroy_joints(orc_world* w, uint32_t n, const uint32_t* ids);
// Destroy(AabbTreeWorld&, BodyID), :1062-1106
int orc_destroy_bodies(orc_world* w, uint32_t n, const uint32_t* ids)
{
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t id = ids[i];
        if (id >= w->bodies.size()) { w->error = "no such body"; return -2; }
        if (id < w->bodyDead.size() && w->bodyDead[id]) { w->error = "destroyed"; return -3; }
        OBody& ob = w->bodies[id];
        { const std::vector<uint32_t> js = ob.joints; for (uint32_t j : js) orc_destroy_joints(w, 1u, &j); } // the attached joints in the body's list order (:1072-1085)
        { const std::vector<uint32_t> list = ob.contacts; for (uint32_t c : list) destroy_contact(*w, c); }
        { auto& f = w->fixturesForProxies; f.erase(std::remove_if(f.begin(), f.end(), [id](const std::pair<uint32_t, uint32_t>& e) { return e.first == id; }), f.end()); }
        std::vector<uint32_t> slots = ob.proxies; // FindProxies: ascending leaf id (:617-630)
        std::sort(slots.begin(), slots.end(), [w](uint32_t a, uint32_t b) { return w->proxies[a].p.tree_id < w->proxies[b].p.tree_id; });
        destroy_proxy_slots(*w, slots);
        { auto& bs = w->bodiesForSync; bs.erase(std::remove(bs.begin(), bs.end(), id), bs.end()); } // Remove(BodyID) :1046-1060
        w->bodyOrder.erase(std::remove(w->bodyOrder.begin(), w->bodyOrder.end(), id), w->bodyOrder.end());
        OBody gone; std::memset(&gone.b, 0, sizeof(gone.b)); gone.b.xf_qc = 1.0f; gone.b.flags = PR_BODY_DESTROYED;
        w->bodies[id] = gone;
        if (w->bodyDead.size() <= id) w->bodyDead.resize(id + 1u, 0);
        w->bodyDead[id] = 1; w->bodyFree.push_back(id);
    }
    return 0;
}
// Detach = the SetBody whose value has fewer shapes (:2760-2785)
int orc_detach_shapes(orc_world* w, uint32_t body, uint32_t n, const uint32_t* shape_ids)
{
    std::vector<uint32_t> ids(shape_ids, shape_ids + n);
    std::sort(ids.begin(), ids.end()); // (GetOldAndNewShapeIDs walks a std::map)
    if (enabled(w->bodies[body].b)) body_remove_shapes(*w, body, ids, false);
    else for (uint32_t sid : ids) { auto& l = w->bodies[body].shapes; auto it = std::find(l.begin(), l.end(), sid); if (it != l.end()) l.erase(it); }
    return 0;
}
// Destroy(AabbTreeWorld&, ShapeID), :1259-1277
int orc_destroy_shapes(orc_world* w, uint32_t n, const uint32_t* ids)
{
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t sid = ids[i];
        if (sid >= w->shapes.size()) { w->error = "no such shape"; return -2; }
        if (sid < w->shapeDead.size() && w->shapeDead[sid]) continue;
        for (uint32_t b = 0; b < w->bodies.size(); ++b) {
            const uint32_t cnt = static_cast<uint32_t>(std::count(w->bodies[b].shapes.begin(), w->bodies[b].shapes.end(), sid));
            if (!cnt) continue;
            const std::vector<uint32_t> rm(cnt, sid);
            orc_detach_shapes(w, b, cnt, rm.data());
        }
        if (w->shapeDead.size() <= sid) w->shapeDead.resize(sid + 1u, 0);
        w->shapeDead[sid] = 1; w->shapeFree.push_back(sid);
        w->shapes[sid].child_count = 0;
    }
    return 0;
}
// the SetBody whose value differs in IsEnabled (GetOldAndNewShapeIDs :671-676)
int orc_set_body_enabled(orc_world* w, uint32_t body, uint32_t en)
{
    OBody& ob = w->bodies[body];
    if (enabled(ob.b) == (en != 0)) return 0;
    if (!en) {
        std::vector<uint32_t> ids = ob.shapes; std::sort(ids.begin(), ids.end());
        body_remove_shapes(*w, body, ids, true);
        ob.b.flags &= ~static_cast<uint32_t>(PR_BODY_ENABLED);
    } else {
        for (uint32_t sid : ob.shapes) w->fixturesForProxies.emplace_back(body, sid);
        ob.b.flags |= PR_BODY_ENABLED;
    }
    return 0;
}
// the SetBody whose value has another BodyType (:2738-2758)
int orc_set_body_type(orc_world* w, uint32_t body, const pr_body* row)
{
    OBody& ob = w->bodies[body];
    { const std::vector<uint32_t> list = ob.contacts; for (uint32_t c : list) destroy_contact(*w, c); }
    const pr_body old = ob.b;
    ob.b = *row;
    bool sync = !(row->flags & (PR_BODY_VELOCITY | PR_BODY_ACCELERATION));
    if (old.xf_px != ob.b.xf_px || old.xf_py != ob.b.xf_py || old.xf_qc != ob.b.xf_qc || old.xf_qs != ob.b.xf_qs) { flag_body_contacts_dirty(*w, body); sync = true; }
    if (sync) w->bodiesForSync.push_back(body);
    return 0;
}
// SetShape(AabbTreeWorld&, ShapeID, Shape), :1284-1356, any new value
int orc_set_shape(orc_world* w, uint32_t shape, const pr_shape* def, const pr_child* children, const pr_vec2* vertices, const pr_vec2* normals,
                  uint32_t n_vertices, uint32_t geometry_changed)
{
    pr_shape cur = w->shapes[shape];
    for (uint32_t b = 0; b < w->bodies.size(); ++b) {
        OBody& ob = w->bodies[b];
        if (std::find(ob.shapes.begin(), ob.shapes.end(), shape) == ob.shapes.end()) continue;
        set_awake(ob.b);
        if (geometry_changed && enabled(ob.b)) {
            body_remove_shapes(*w, b, std::vector<uint32_t>{shape}, true);
            w->fixturesForProxies.emplace_back(b, shape);
        }
    }
    if (geometry_changed) {
        const uint32_t cbase = static_cast<uint32_t>(w->children.size()), vbase = static_cast<uint32_t>(w->verts.size());
        for (uint32_t i = 0; i < def->child_count; ++i) { pr_child c = children[i]; c.vertex_offset += vbase; w->children.push_back(c); }
        w->verts.insert(w->verts.end(), vertices, vertices + n_vertices);
        w->normals.insert(w->normals.end(), normals, normals + n_vertices);
        cur.child_offset = cbase; cur.child_count = def->child_count;
    }
    if (cur.filter_category != def->filter_category || cur.filter_mask != def->filter_mask || cur.filter_group != def->filter_group) { // :1330-1345
        for (uint32_t id : w->order) {
            pr_contact& c = w->contacts[id].c;
            if (c.shape_a == shape || c.shape_b == shape) { c.flags |= PR_CONTACT_FILTER; set_awake(w->bodies[c.body_a].b); set_awake(w->bodies[c.body_b].b); w->needsFiltering = true; }
        }
        for (uint32_t k = 0; k < w->proxies.size(); ++k) if (w->proxies[k].p.shape == shape) w->proxiesForContacts.push_back(k);
    }
    if ((cur.is_sensor != 0) != (def->is_sensor != 0) || cur.friction != def->friction || cur.restitution != def->restitution || geometry_changed) { // :1346-1354
        for (uint32_t id : w->order) {
            pr_contact& c = w->contacts[id].c;
            if (c.shape_a == shape || c.shape_b == shape) { c.flags |= PR_CONTACT_DIRTY; set_awake(w->bodies[c.body_a].b); set_awake(w->bodies[c.body_b].b); }
        }
    }
    cur.friction = def->friction; cur.restitution = def->restitution; cur.density = def->density; cur.is_sensor = def->is_sensor ? 1 : 0;
    cur.filter_category = def->filter_category; cur.filter_mask = def->filter_mask; cur.filter_group = def->filter_group;
    w->shapes[shape] = cur;
    return 0;
}

// Destroy(AabbTreeWorld&, JointID), :1215-1227 + Remove(JointID) :1177-1213
int orc_destroy_joints(orc_world* w, uint32_t n, const uint32_t* ids)
{
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t id = ids[i];
        if (id >= w->joints.size()) { w->error = "no such joint"; return -2; }
        pr_revolute_joint& j = w->joints[id];
        if (j.body_a == 0xFFFFFFFFu) continue;
        if (!(j.collide_connected & 1u)) flag_pair_for_filtering(*w, j.body_a, j.body_b); // Remove(JointID), :1184-1190
        for (uint32_t b : {j.body_a, j.body_b}) {
            set_awake(w->bodies[b].b);
            auto& lst = w->bodies[b].joints;
            auto it = std::find(lst.begin(), lst.end(), id);
            if (it != lst.end()) lst.erase(it);
        }
        std::memset(&j, 0, sizeof(j));
        j.body_a = j.body_b = 0xFFFFFFFFu;
        w->freeJoints.push_back(id);
    }
    return 0;
}

// SetJoint, AabbTreeWorld.cpp:1108-1130 (same bodies, same collideConnected): Remove(id) wakes both bodies and takes the joint
// out of their m_bodyJoints lists, Add(id) appends it again
int orc_set_revolute_joints(orc_world* w, uint32_t n, const uint32_t* ids, const pr_revolute_joint* rows)
{
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t id = ids[i];
        if (id >= w->joints.size()) { w->error = "no such joint"; return -2; }
        const pr_revolute_joint old = w->joints[id];
        if (old.body_a == 0xFFFFFFFFu) { w->error = "joint identifier is destroyed"; return -1; }
        if (rows[i].body_a >= w->bodies.size() || rows[i].body_b >= w->bodies.size()) { w->error = "no such body"; return -2; }
        if (PR_JOINT_TYPE(old) != PR_JOINT_TYPE(rows[i]) && (PR_JOINT_TYPE(old) == PR_JOINT_GEAR || PR_JOINT_TYPE(rows[i]) == PR_JOINT_GEAR)) { w->error = "unsupported"; return -6; }
        // Remove(id), :1177-1213: the old pair's contacts flagged, the old bodies woken and the joint taken out of their lists
        if (!(old.collide_connected & 1u)) flag_pair_for_filtering(*w, old.body_a, old.body_b);
        for (uint32_t b : {old.body_a, old.body_b}) {
            set_awake(w->bodies[b].b);
            auto& lst = w->bodies[b].joints;
            auto it = std::find(lst.begin(), lst.end(), id);
            if (it != lst.end()) lst.erase(it);
        }
        w->joints[id] = rows[i];
        w->bodies[rows[i].body_a].joints.push_back(id);
        if (rows[i].body_b != rows[i].body_a) w->bodies[rows[i].body_b].joints.push_back(id);
        if (!(rows[i].collide_connected & 1u)) flag_pair_for_filtering(*w, rows[i].body_a, rows[i].body_b); // Add(id, true)
    }
    return 0;
}

int orc_set_bodies(orc_world* w, uint32_t n, const uint32_t* ids, const pr_body* rows)
{
    // SetBody (transform change path), AabbTreeWorld.cpp:2723-2795
    for (uint32_t i = 0; i < n; ++i) {
        OBody& ob = w->bodies[ids[i]];
        const pr_body old = ob.b;
        ob.b = rows[i];
        if (old.xf_px != ob.b.xf_px || old.xf_py != ob.b.xf_py || old.xf_qc != ob.b.xf_qc || old.xf_qs != ob.b.xf_qs) {
            flag_body_contacts_dirty(*w, ids[i]);
            w->bodiesForSync.push_back(ids[i]);
        }
    }
    return 0;
}

// ShiftOrigin(AabbTreeWorld&, newOrigin), AabbTreeWorld.cpp:2221-2243
int orc_shift_origin(orc_world* w, float x, float y)
{
    for (uint32_t id = 0; id < w->bodies.size(); ++id) {
        pr_body& b = w->bodies[id].b;
        b.pos0_x -= x; b.pos0_y -= y;
        Pos p1 = body_pos1(b);
        p1.l.x -= x; p1.l.y -= y;
        set_pos1(b, p1); // SetSweep: m_xf = GetTransform1(sweep), Body.hpp:697-702
        flag_body_contacts_dirty(*w, id);
    }
    for (pr_revolute_joint& j : w->joints) { // ShiftOrigin(Joint&, newOrigin), AabbTreeWorld.cpp:2237-2240: target joint's target, pulley joint's ground anchors
        if (j.body_a == 0xFFFFFFFFu && j.body_b == 0xFFFFFFFFu) continue;
        if (PR_JOINT_TYPE(j) == PR_JOINT_TARGET) { j.local_anchor_a.x -= x; j.local_anchor_a.y -= y; }
        else if (PR_JOINT_TYPE(j) == PR_JOINT_PULLEY) { j.impulse_y -= x; j.impulse_z -= y; j.angular_motor_impulse -= x; j.reference_angle -= y; }
    }
    const float nx = -x, ny = -y; // DynamicTree::ShiftOrigin: GetMovedAABB(aabb, -newOrigin), DynamicTree.cpp:650-658
    for (OProxy& op : w->proxies) { op.p.min_x += nx; op.p.max_x += nx; op.p.min_y += ny; op.p.max_y += ny; }
    return 0;
}

// ---------------- RayCast, RayCastOutput.cpp ----------------
namespace {
struct RayOut { bool hit; Rot normal; float fraction; };
// RayCast(Length radius, const Length2& location, const RayCastInput&), :41-75
RayOut ray_cast_circle(float radius, V2 location, V2 p1, V2 p2, float maxFraction)
{
    const V2 s = p1 - location;
    const float b = magsq(s) - square(radius);
    const V2 raySegment = p2 - p1;
    const float c = dot(s, raySegment);
    const float rr = magsq(raySegment);
    const float sigma = square(c) - rr * b;
    if (sigma < 0.0f || almost_zero(rr)) return {false, {0, 0}, 0};
    const float a = -(c + std::sqrt(sigma));
    const float fraction = a / rr;
    if (fraction >= 0.0f && fraction <= maxFraction) return {true, unit_vector(s + fraction * raySegment, Rot{0, 0}), fraction};
    return {false, {0, 0}, 0};
}
// RayCast(const DistanceProxy&, const RayCastInput&, const Transformation&), :139-233
RayOut ray_cast_proxy(const Proxy& proxy, V2 p1, V2 p2, float maxFraction, Xf xf)
{
    const uint32_t count = proxy.count;
    const float radius = proxy.radius;
    V2 v0 = proxy.vertex(0);
    if (count == 1) return ray_cast_circle(radius, transform(v0, xf), p1, p2, maxFraction);
    const V2 tp1 = inv_transform(p1, xf), tp2 = inv_transform(p2, xf);
    const V2 ray0 = tp1, ray = tp2 - tp1;
    float minT = std::nextafter(maxFraction, 2.0f);
    Rot normalFound{0, 0};
    for (uint32_t i = 0; i < count; ++i) {
        const RayOut circle = ray_cast_circle(radius, v0, tp1, tp2, maxFraction);
        if (circle.hit && minT > circle.fraction) { minT = circle.fraction; normalFound = circle.normal; }
        const V2 v1 = proxy.vertex(mod_next(i, count));
        const V2 edge = v1 - v0;
        const float rayCrossEdge = cross(ray, edge);
        if (!almost_zero(rayCrossEdge)) {
            const Rot normal = proxy.normal(i);
            const V2 offset = normal * radius;
            const V2 v0off = v0 + offset;
            const V2 qSubP = v0off - ray0;
            const float recip = 1.0f / rayCrossEdge;
            const float t = cross(qSubP, edge) * recip;
            const float u = cross(qSubP, ray) * recip;
            if (t >= 0.0f && t <= 1.0f && u >= 0.0f && u <= 1.0f) {
                if (minT > t) { minT = t; normalFound = normal; }
            }
        }
        v0 = v1;
    }
    if (minT <= maxFraction) return {true, rot_rotate(normalFound, xf.q), minT};
    return {false, {0, 0}, 0};
}
} // namespace

int orc_ray_cast(const orc_world* w, uint32_t n_rays, const float* rays5, uint32_t cap, pr_ray_hit* hits, uint32_t* n_out)
{
    uint32_t n = 0;
    for (uint32_t r = 0; r < n_rays; ++r) {
        const V2 p1{rays5[5 * r], rays5[5 * r + 1]}, p2{rays5[5 * r + 2], rays5[5 * r + 3]};
        const float maxFraction = rays5[5 * r + 4];
        // RayCast(const DynamicTree&, RayCastInput, cb), :241-299: the two tests a leaf must pass
        const Rot v = rev_perp(unit_vector(p2 - p1, Rot{0, 0}));
        const V2 absV{std::fabs(v.c), std::fabs(v.s)};
        const V2 fractDelta = maxFraction * (p2 - p1); // GetAABB(RayCastInput), AABB.cpp:107-112
        const V2 e = p1 + fractDelta;
        const AABB seg{std::min(p1.x, e.x), std::max(p1.x, e.x), std::min(p1.y, e.y), std::max(p1.y, e.y)};
        for (const OProxy& op : w->proxies) {
            const AABB box = fat_of(op.p);
            if (!overlap(box, seg)) continue;
            const V2 center{(box.minx + box.maxx) / 2, (box.miny + box.maxy) / 2};
            const V2 extents{(box.maxx - box.minx) * 0.5f, (box.maxy - box.miny) * 0.5f};
            const float separation = std::fabs(dot(rotv(v), p1 - center)) - dot(absV, extents);
            if (separation > 0.0f) continue;
            const RayOut o = ray_cast_proxy(w->child(op.p.shape, op.p.child), p1, p2, maxFraction, body_xf(w->bodies[op.p.body].b));
            if (!o.hit) continue;
            if (n < cap) {
                const V2 point = p1 + (p2 - p1) * o.fraction;
                hits[n] = pr_ray_hit{r, op.p.body, op.p.shape, op.p.child, o.fraction, o.normal.c, o.normal.s, point.x, point.y};
            }
            ++n;
        }
    }
    *n_out = n;
    return 0;
}

// SetContact / SetManifold, AabbTreeWorld.cpp:2797-2846 (-2: no such contact, -1: what the reference throws InvalidArgument for)
int orc_set_contacts(orc_world* w, uint32_t n, const pr_contact* rows)
{
    for (uint32_t i = 0; i < n; ++i) {
        const pr_contact& v = rows[i];
        if (v.id >= w->contacts.size()) { w->error = "no such contact"; return -2; }
        const pr_contact& e = w->contacts[v.id].c;
        if (std::find(w->freeContacts.begin(), w->freeContacts.end(), v.id) != w->freeContacts.end()) { w->error = "contact identifier is destroyed"; return -1; }
        const bool same = (e.flags & PR_CONTACT_DESTROYED) == (v.flags & PR_CONTACT_DESTROYED) && e.body_a == v.body_a && e.shape_a == v.shape_a &&
                          e.child_a == v.child_a && e.body_b == v.body_b && e.shape_b == v.shape_b && e.child_b == v.child_b &&
                          (e.flags & PR_CONTACT_IMPENETRABLE) == (v.flags & PR_CONTACT_IMPENETRABLE) && (e.flags & PR_CONTACT_SENSOR) == (v.flags & PR_CONTACT_SENSOR) &&
                          (e.flags & PR_CONTACT_TOI) == (v.flags & PR_CONTACT_TOI) && (!(e.flags & PR_CONTACT_TOI) || e.toi == v.toi) && e.toi_count == v.toi_count;
        if (!same) { w->error = "immutable contact field changed"; return -1; }
    }
    for (uint32_t i = 0; i < n; ++i) {
        pr_contact& e = w->contacts[rows[i].id].c;
        e.friction = rows[i].friction; e.restitution = rows[i].restitution; e.tangent_speed = rows[i].tangent_speed; e.flags = rows[i].flags & 0xFFu;
    }
    return 0;
}
int orc_set_manifolds(orc_world* w, uint32_t n, const uint32_t* ids, const pr_manifold* rows)
{
    for (uint32_t i = 0; i < n; ++i) {
        if (ids[i] >= w->contacts.size()) { w->error = "no such contact"; return -2; }
        const pr_manifold& m = w->contacts[ids[i]].m;
        if (std::find(w->freeContacts.begin(), w->freeContacts.end(), ids[i]) != w->freeContacts.end()) { w->error = "contact identifier is destroyed"; return -1; }
        if (m.type != rows[i].type || m.count != rows[i].count) { w->error = "cannot change manifold type / point count"; return -1; }
    }
    for (uint32_t i = 0; i < n; ++i) w->contacts[ids[i]].m = rows[i];
    return 0;
}

int orc_set_event_mask(orc_world* w, uint32_t mask) { w->evMask = mask; return 0; }
int orc_events(const orc_world* w, uint32_t kind, void* buf, uint32_t cap, uint32_t* n_out)
{ // call order (the order the reference's listeners are invoked in)
    if (kind == PR_EVENT_PRE_SOLVE) {
        *n_out = static_cast<uint32_t>(w->evPre.size());
        std::memcpy(buf, w->evPre.data(), sizeof(pr_presolve_event) * std::min<size_t>(cap, w->evPre.size()));
    } else if (kind == PR_EVENT_POST_SOLVE) {
        *n_out = static_cast<uint32_t>(w->evImpulse.size());
        std::memcpy(buf, w->evImpulse.data(), sizeof(pr_impulse_event) * std::min<size_t>(cap, w->evImpulse.size()));
    } else {
        std::vector<pr_contact_event> ev;
        for (const auto& e : w->evContact) if (e.kind & kind) ev.push_back(e);
        *n_out = static_cast<uint32_t>(ev.size());
        std::memcpy(buf, ev.data(), sizeof(pr_contact_event) * std::min<size_t>(cap, ev.size()));
    }
    return 0;
}

int orc_load_state(orc_world* w, uint32_t n_bodies, const pr_body* bodies, uint32_t n_proxies,
                   const pr_proxy* proxies, uint32_t n_contacts, const pr_contact* contacts,
                   const pr_manifold* manifolds, float inv_dt0)
{
    if (n_bodies != w->bodies.size()) { w->error = "body count mismatch"; return -1; }
    for (uint32_t i = 0; i < n_bodies; ++i) { w->bodies[i].b = bodies[i]; w->bodies[i].contacts.clear(); w->bodies[i].proxies.clear(); }
    w->proxies.clear();
    std::unordered_map<uint32_t, uint32_t> byTree;
    for (uint32_t i = 0; i < n_proxies; ++i) {
        OProxy op; op.p = proxies[i];
        byTree[op.p.tree_id] = i;
        w->bodies[op.p.body].proxies.push_back(i);
        w->proxies.push_back(op);
    }
    w->contacts.clear(); w->freeContacts.clear(); w->order.clear();
    uint32_t maxId = 0;
    for (uint32_t i = 0; i < n_contacts; ++i) maxId = std::max(maxId, contacts[i].id + 1);
    w->contacts.resize(maxId);
    std::vector<char> used(maxId, 0);
    for (uint32_t i = 0; i < n_contacts; ++i) {
        OContact& oc = w->contacts[contacts[i].id];
        oc.c = contacts[i]; oc.m = manifolds[i];
        oc.pxA = byTree.at(oc.c.key_min); oc.pxB = byTree.at(oc.c.key_max);
        used[contacts[i].id] = 1;
        w->order.push_back(contacts[i].id);
        w->bodies[oc.c.body_a].contacts.push_back(contacts[i].id);
        w->bodies[oc.c.body_b].contacts.push_back(contacts[i].id);
    }
    // free ids: pushed in ascending order is NOT the reference's history; tests that load state
    // must not rely on id reuse order beyond "smallest id popped last".
    for (uint32_t id = maxId; id-- > 0;) if (!used[id]) { w->freeContacts.push_back(id); w->contacts[id].c.flags = PR_CONTACT_DESTROYED; w->contacts[id].c.id = id; }
    w->fixturesForProxies.clear(); w->proxiesForContacts.clear(); w->bodiesForSync.clear();
    w->inv_dt0 = inv_dt0;
    w->stepComplete = true;
    return 0;
}

// Step, AabbTreeWorld.cpp:2144-2219
int orc_step(orc_world* w, const pr_step_conf* confp, pr_step_stats* stats_out, const uint32_t* order, uint32_t n_order)
{
    const pr_step_conf& conf = *confp;
    pr_step_stats stats;
    std::memset(&stats, 0, sizeof(stats));
    stats.reg.min_separation = kInf;
    stats.toi.min_separation = kInf;
    w->evContact.clear(); w->evImpulse.clear(); w->evPre.clear();
    for (const auto& bs : w->fixturesForProxies) stats.pre.proxies_created += create_proxies(*w, bs.first, bs.second, conf);
    w->fixturesForProxies.clear();
    for (uint32_t bodyId : w->bodiesForSync) {
        const pr_body& b = w->bodies[bodyId].b;
        stats.pre.proxies_moved += synchronize(*w, bodyId, get_xf(body_pos0(b), body_lc(b)), body_xf(b), conf);
    }
    w->bodiesForSync.clear();
    stats.pre.contacts_destroyed = destroy_contacts(*w);
    update_contacts(*w, conf, stats.pre.contacts_updated, stats.pre.contacts_skipped);
    stats.pre.contacts_added = find_and_add_contacts(*w, conf);
    if (conf.delta_time != 0.0f) {
        w->inv_dt0 = 1.0f / conf.delta_time;
        std::unordered_map<uint32_t, uint32_t> rank;
        if (order) for (uint32_t i = 0; i < n_order; ++i) rank[order[i]] = i;
        if (w->stepComplete) solve_reg(*w, conf, stats, order ? &rank : nullptr);
        if (conf.do_toi) solve_toi(*w, conf, stats);
    }
    if (stats_out) *stats_out = stats;
    return 0;
}

int orc_get_counts(const orc_world* w, pr_counts* out)
{
    std::memset(out, 0, sizeof(*out));
    out->bodies = static_cast<uint32_t>(w->bodies.size());
    out->shapes = static_cast<uint32_t>(w->shapes.size());
    out->children = static_cast<uint32_t>(w->children.size());
    out->vertices = static_cast<uint32_t>(w->verts.size());
    out->proxies = static_cast<uint32_t>(w->proxies.size());
    out->contacts = static_cast<uint32_t>(w->order.size());
    out->joints = static_cast<uint32_t>(w->joints.size());
    out->contact_id_range = static_cast<uint32_t>(w->contacts.size());
    out->tree_leaf_count = out->proxies;
    return 0;
}

int orc_get_bodies(const orc_world* w, uint32_t first, uint32_t n, pr_body* rows)
{
    for (uint32_t i = 0; i < n; ++i) rows[i] = w->bodies[first + i].b;
    return 0;
}

int orc_get_contacts(const orc_world* w, uint32_t cap, pr_contact* contacts, pr_manifold* manifolds, uint32_t* n_out)
{
    uint32_t n = 0;
    for (uint32_t id : w->order) {
        if (n < cap) { contacts[n] = w->contacts[id].c; if (manifolds) manifolds[n] = w->contacts[id].m; }
        ++n;
    }
    *n_out = n;
    return 0;
}

int orc_get_proxies(const orc_world* w, uint32_t cap, pr_proxy* proxies, uint32_t* n_out)
{
    uint32_t n = 0;
    for (const auto& p : w->proxies) { if (n < cap) proxies[n] = p.p; ++n; }
    *n_out = n;
    return 0;
}

// CreateJoint of GearJointConf values: the table row (bodies A and B, impulse, ratio, constant) + the whole value by joint id
int orc_create_gear_joints(orc_world* w, uint32_t n, const pr_gear_joint* joints)
{
    for (uint32_t i = 0; i < n; ++i) {
        pr_revolute_joint r{};
        r.body_a = joints[i].body_a; r.body_b = joints[i].body_b;
        r.collide_connected = (joints[i].collide_connected & 1u) | (static_cast<uint32_t>(PR_JOINT_GEAR) << 8);
        r.impulse_x = joints[i].impulse; r.lower_angle = joints[i].ratio; r.upper_angle = joints[i].constant;
        const uint32_t id = w->freeJoints.empty() ? static_cast<uint32_t>(w->joints.size()) : w->freeJoints.back();
        orc_create_revolute_joints(w, 1, &r);
        if (w->gears.size() < w->joints.size()) w->gears.resize(w->joints.size());
        w->gears[id] = joints[i];
    }
    return 0;
}
int orc_get_revolute_joints(const orc_world* w, uint32_t first, uint32_t n, pr_revolute_joint* rows)
{
    for (uint32_t i = 0; i < n; ++i) rows[i] = w->joints[first + i];
    return 0;
}

float orc_get_inv_dt0(const orc_world* w) { return w->inv_dt0; }
int orc_set_toi_mode(orc_world* w, uint32_t mode) { if (mode > 1u) return 1; w->toiMode = mode; return 0; }
int orc_get_toi_schedule(const orc_world* w, uint32_t cap, uint32_t* round_contact_pairs, uint32_t* n_out)
{
    const uint32_t n = static_cast<uint32_t>(w->toiSchedule.size());
    *n_out = n;
    for (uint32_t i = 0; i < n && i < cap; ++i) { round_contact_pairs[2u * i] = w->toiSchedule[i].first; round_contact_pairs[2u * i + 1u] = w->toiSchedule[i].second; }
    return n <= cap ? 0 : 1;
}

int orc_collide_shapes(uint32_t n, const pr_child* children, const pr_vec2* verts, const pr_vec2* normals,
                       const uint32_t* a_child, const float* xf_a, const uint32_t* b_child, const float* xf_b,
                       float linear_slop, float max_circles_ratio, pr_manifold* out)
{
    for (uint32_t i = 0; i < n; ++i) {
        const pr_child& ca = children[a_child[i]];
        const pr_child& cb = children[b_child[i]];
        const Proxy pa{ca.vertex_radius, ca.vertex_count, verts + ca.vertex_offset, normals + ca.vertex_offset};
        const Proxy pb{cb.vertex_radius, cb.vertex_count, verts + cb.vertex_offset, normals + cb.vertex_offset};
        const float* a = xf_a + 4 * i;
        const float* b = xf_b + 4 * i;
        out[i] = collide_shapes(pa, Xf{{a[0], a[1]}, {a[2], a[3]}}, pb, Xf{{b[0], b[1]}, {b[2], b[3]}}, linear_slop, max_circles_ratio);
    }
    return 0;
}

} // extern "C"
