// CudaWorld.cpp -- host shim of the B200-native world (see CudaWorld.hpp). Calls ONLY the C ABI of
// include/playrho_b200.h; mirrors the argument checks and exception types of AabbTreeWorld
// (Library/source/playrho/d2/AabbTreeWorld.cpp, lines cited per function).
#include "CudaWorld.hpp"
#include "../../include/playrho_b200.h"

#include <playrho/InvalidArgument.hpp>
#include <playrho/LengthError.hpp>
#include <playrho/OutOfRange.hpp>
#include <playrho/WrongState.hpp>
#include <playrho/d2/DistanceProxy.hpp>
#include <playrho/d2/RevoluteJointConf.hpp>
#include <playrho/d2/DistanceJointConf.hpp>
#include <playrho/d2/RopeJointConf.hpp>
#include <playrho/d2/FrictionJointConf.hpp>
#include <playrho/d2/MotorJointConf.hpp>
#include <playrho/d2/WeldJointConf.hpp>
#include <playrho/d2/TargetJointConf.hpp>
#include <playrho/d2/PrismaticJointConf.hpp>
#include <playrho/d2/WheelJointConf.hpp>
#include <playrho/d2/PulleyJointConf.hpp>
#include <playrho/d2/GearJointConf.hpp>
#include <variant>
#include <playrho/to_underlying.hpp>

#include <algorithm>
#include <cstring>
#include <map>
#include <limits>
#include <stdexcept>
#include <string>

namespace playrho::b200 {

namespace {
constexpr auto worldIsLockedMsg = "world is locked";
constexpr auto noSuchBodyMsg = "no such body";
constexpr auto noSuchShapeMsg = "no such shape";
constexpr auto noSuchJointMsg = "no such joint";
constexpr auto noSuchContactMsg = "no such contact";
constexpr auto unsupportedMsg = "operation not supported by the B200 device world yet";

template <class ID>
ID MakeId(std::uint32_t v) { return static_cast<ID>(static_cast<typename ID::underlying_type>(v)); }

template <class Container, class ID>
auto& AtId(Container& c, ID id, const char* msg)
{
    const auto i = static_cast<std::size_t>(to_underlying(id));
    if (i >= c.size()) throw OutOfRange<ID>{id, msg}; // At(...), AabbTreeWorld.cpp:163-174
    return c[i];
}

pr_contact_feature Pack(const ContactFeature& cf)
{
    return static_cast<std::uint32_t>(cf.typeA) | (static_cast<std::uint32_t>(cf.indexA) << 8) |
           (static_cast<std::uint32_t>(cf.typeB) << 16) | (static_cast<std::uint32_t>(cf.indexB) << 24);
}
ContactFeature Unpack(pr_contact_feature v)
{
    return ContactFeature{static_cast<ContactFeature::Type>(v & 0xFFu), static_cast<ContactFeature::Index>((v >> 8) & 0xFFu),
                          static_cast<ContactFeature::Type>((v >> 16) & 0xFFu), static_cast<ContactFeature::Index>((v >> 24) & 0xFFu)};
}

// Body value -> device row (Body.hpp private block; flags bit-identical to Body::Flag)
pr_body ToRow(const Body& b)
{
    pr_body o{};
    const auto xf = GetTransformation(b);
    const auto& sw = GetSweep(b);
    const auto v = GetVelocity(b);
    o.xf_px = GetX(xf.p); o.xf_py = GetY(xf.p); o.xf_qc = xf.q.GetX(); o.xf_qs = xf.q.GetY();
    o.pos0_x = GetX(sw.pos0.linear); o.pos0_y = GetY(sw.pos0.linear); o.pos0_a = sw.pos0.angular;
    o.pos1_x = GetX(sw.pos1.linear); o.pos1_y = GetY(sw.pos1.linear); o.pos1_a = sw.pos1.angular;
    o.lc_x = GetX(sw.localCenter); o.lc_y = GetY(sw.localCenter);
    o.alpha0 = sw.alpha0;
    o.vel_x = GetX(v.linear); o.vel_y = GetY(v.linear); o.vel_w = v.angular;
    o.acc_x = GetX(GetLinearAcceleration(b)); o.acc_y = GetY(GetLinearAcceleration(b)); o.acc_w = GetAngularAcceleration(b);
    o.inv_mass = GetInvMass(b); o.inv_rot_i = GetInvRotInertia(b);
    o.lin_damp = GetLinearDamping(b); o.ang_damp = GetAngularDamping(b);
    o.under_active_time = GetUnderActiveTime(b);
    std::uint32_t f = 0;
    if (IsDestroyed(b)) f |= PR_BODY_DESTROYED;
    if (IsAwake(b)) f |= PR_BODY_AWAKE;
    if (IsSleepingAllowed(b)) f |= PR_BODY_AUTOSLEEP;
    if (IsImpenetrable(b)) f |= PR_BODY_IMPENETRABLE;
    if (IsFixedRotation(b)) f |= PR_BODY_FIXED_ROTATION;
    if (IsEnabled(b)) f |= PR_BODY_ENABLED;
    if (IsSpeedable(b)) f |= PR_BODY_VELOCITY;
    if (IsAccelerable(b)) f |= PR_BODY_ACCELERATION;
    if (IsMassDataDirty(b)) f |= PR_BODY_MASS_DIRTY;
    o.flags = f;
    return o;
}

// device row -> Body value: only what a Step can change (sweep, transform, velocity, awake, timer)
void FromRow(const pr_body& r, Body& b)
{
    auto sweep = Sweep{Position{Length2{r.pos0_x, r.pos0_y}, Angle{r.pos0_a}}, Position{Length2{r.pos1_x, r.pos1_y}, Angle{r.pos1_a}},
                       Length2{r.lc_x, r.lc_y}, ZeroToUnderOneFF<Real>(r.alpha0)};
    b.SetSweep(sweep); // (recomputes the transformation from pos1 with the same libm sincos the device restates)
    b.JustSetVelocity(Velocity{LinearVelocity2{r.vel_x, r.vel_y}, AngularVelocity{r.vel_w}});
    if (r.flags & PR_BODY_AWAKE) b.SetAwakeFlag(); else b.UnsetAwakeFlag();
    b.SetUnderActiveTime(Time{r.under_active_time});
}

pr_step_conf ToConf(const StepConf& s)
{
    pr_step_conf c{};
    c.delta_time = s.deltaTime; c.dt_ratio = s.dtRatio;
    c.min_still_time_to_sleep = s.minStillTimeToSleep; c.linear_slop = s.linearSlop; c.angular_slop = s.angularSlop;
    c.reg_resolution_rate = s.regResolutionRate; c.reg_min_separation = s.regMinSeparation; c.reg_min_momentum = s.regMinMomentum;
    c.toi_resolution_rate = s.toiResolutionRate; c.toi_min_separation = s.toiMinSeparation; c.toi_min_momentum = s.toiMinMomentum;
    c.target_depth = s.targetDepth; c.tolerance = s.tolerance; c.velocity_threshold = s.velocityThreshold;
    c.max_translation = s.maxTranslation; c.max_rotation = s.maxRotation;
    c.max_linear_correction = s.maxLinearCorrection; c.max_angular_correction = s.maxAngularCorrection;
    c.linear_sleep_tolerance = s.linearSleepTolerance; c.angular_sleep_tolerance = s.angularSleepTolerance;
    c.displace_multiplier = s.displaceMultiplier; c.aabb_extension = s.aabbExtension; c.max_circles_ratio = s.maxCirclesRatio;
    c.reg_velocity_iters = s.regVelocityIters; c.reg_position_iters = s.regPositionIters;
    c.toi_velocity_iters = s.toiVelocityIters; c.toi_position_iters = s.toiPositionIters;
    c.max_toi_root_iters = s.maxToiRootIters; c.max_toi_iters = s.maxToiIters; c.max_distance_iters = s.maxDistanceIters;
    c.max_sub_steps = s.maxSubSteps;
    c.do_warm_start = s.doWarmStart; c.do_toi = s.doToi; c.do_blocksolve = s.doBlocksolve;
    return c;
}

template <class T, class U> T Narrow(U v) { return static_cast<T>(v); }

StepStats ToStats(const pr_step_stats& o)
{
    StepStats s;
    using PC = decltype(s.pre.proxiesCreated);
    s.pre.proxiesCreated = Narrow<PC>(o.pre.proxies_created); s.pre.proxiesMoved = Narrow<PC>(o.pre.proxies_moved);
    s.pre.contactsDestroyed = Narrow<PC>(o.pre.contacts_destroyed); s.pre.contactsAdded = Narrow<PC>(o.pre.contacts_added);
    s.pre.contactsUpdated = Narrow<PC>(o.pre.contacts_updated); s.pre.contactsSkipped = Narrow<PC>(o.pre.contacts_skipped);
    s.reg.minSeparation = o.reg.min_separation; s.reg.maxIncImpulse = o.reg.max_inc_impulse;
    using RC = decltype(s.reg.islandsFound);
    s.reg.islandsFound = Narrow<RC>(o.reg.islands_found); s.reg.islandsSolved = Narrow<RC>(o.reg.islands_solved);
    s.reg.contactsAdded = Narrow<RC>(o.reg.contacts_added); s.reg.bodiesSlept = Narrow<RC>(o.reg.bodies_slept);
    s.reg.proxiesMoved = Narrow<RC>(o.reg.proxies_moved); s.reg.sumPosIters = Narrow<RC>(o.reg.sum_pos_iters);
    s.reg.sumVelIters = Narrow<RC>(o.reg.sum_vel_iters); s.reg.contactsUpdated = Narrow<RC>(o.reg.contacts_updated);
    s.reg.contactsSkipped = Narrow<RC>(o.reg.contacts_skipped);
    s.reg.maxIslandBodies = Narrow<decltype(s.reg.maxIslandBodies)>(o.reg.max_island_bodies);
    s.toi.minSeparation = o.toi.min_separation; s.toi.maxIncImpulse = o.toi.max_inc_impulse;
    using TC = decltype(s.toi.islandsFound);
    s.toi.islandsFound = Narrow<TC>(o.toi.islands_found); s.toi.islandsSolved = Narrow<TC>(o.toi.islands_solved);
    s.toi.contactsFound = Narrow<TC>(o.toi.contacts_found); s.toi.contactsAtMaxSubSteps = Narrow<TC>(o.toi.contacts_at_max_sub_steps);
    s.toi.contactsUpdatedToi = Narrow<TC>(o.toi.contacts_updated_toi);
    s.toi.contactsUpdatedTouching = Narrow<TC>(o.toi.contacts_updated_touching);
    s.toi.contactsSkippedTouching = Narrow<TC>(o.toi.contacts_skipped_touching);
    s.toi.contactsAdded = Narrow<TC>(o.toi.contacts_added); s.toi.proxiesMoved = Narrow<TC>(o.toi.proxies_moved);
    s.toi.sumPosIters = Narrow<TC>(o.toi.sum_pos_iters); s.toi.sumVelIters = Narrow<TC>(o.toi.sum_vel_iters);
    s.toi.maxDistIters = Narrow<decltype(s.toi.maxDistIters)>(o.toi.max_dist_iters);
    s.toi.maxToiIters = Narrow<decltype(s.toi.maxToiIters)>(o.toi.max_toi_iters);
    s.toi.maxRootIters = Narrow<decltype(s.toi.maxRootIters)>(o.toi.max_root_iters);
    return s;
}

Manifold ToManifold(const pr_manifold& m)
{
    using P = Manifold::Point;
    const auto pt = [&](int k) {
        return k == 0 ? P{Length2{m.p0_x, m.p0_y}, Unpack(m.cf0), Momentum{m.jn0}, Momentum{m.jt0}}
                      : P{Length2{m.p1_x, m.p1_y}, Unpack(m.cf1), Momentum{m.jn1}, Momentum{m.jt1}};
    };
    const auto ln = [&]() { // stored local normal is already unit length: rebuild the UnitVec from its raw bits
        UnitVec u; const float raw[2] = {m.ln_x, m.ln_y};
        static_assert(sizeof(UnitVec) == sizeof(raw), "UnitVec layout");
        std::memcpy(static_cast<void*>(&u), raw, sizeof(raw));
        return u;
    };
    const auto lp = Length2{m.lp_x, m.lp_y};
    switch (m.type) {
    case PR_MANIFOLD_CIRCLES: {
        const auto cf = Unpack(m.cf0);
        auto r = Manifold::GetForCircles(lp, cf.indexA, Length2{m.p0_x, m.p0_y}, cf.indexB);
        r.SetImpulses(0, Momentum{m.jn0}, Momentum{m.jt0});
        return r;
    }
    case PR_MANIFOLD_FACE_A:
        if (m.count == 0) return Manifold::GetForFaceA(ln(), lp);
        if (m.count == 1) return Manifold::GetForFaceA(ln(), lp, pt(0));
        return Manifold::GetForFaceA(ln(), lp, pt(0), pt(1));
    case PR_MANIFOLD_FACE_B:
        if (m.count == 0) return Manifold::GetForFaceB(ln(), lp);
        if (m.count == 1) return Manifold::GetForFaceB(ln(), lp, pt(0));
        return Manifold::GetForFaceB(ln(), lp, pt(0), pt(1));
    default: break;
    }
    return Manifold{};
}

Contact ToContact(const pr_contact& c)
{
    auto r = Contact{Contactable{MakeId<BodyID>(c.body_a), MakeId<ShapeID>(c.shape_a), static_cast<ChildCounter>(c.child_a)},
                     Contactable{MakeId<BodyID>(c.body_b), MakeId<ShapeID>(c.shape_b), static_cast<ChildCounter>(c.child_b)}};
    r.SetFriction(c.friction); r.SetRestitution(c.restitution); r.SetTangentSpeed(c.tangent_speed);
    if (c.flags & PR_CONTACT_TOUCHING) r.SetTouching(); else r.UnsetTouching();
    if (c.flags & PR_CONTACT_ENABLED) r.SetEnabled(); else r.UnsetEnabled();
    if (c.flags & PR_CONTACT_FILTER) r.FlagForFiltering(); else r.UnflagForFiltering();
    if (c.flags & PR_CONTACT_DIRTY) r.FlagForUpdating(); else r.UnflagForUpdating();
    if (c.flags & PR_CONTACT_SENSOR) r.SetSensor(); else r.UnsetSensor();
    if (c.flags & PR_CONTACT_IMPENETRABLE) r.SetImpenetrable(); else r.UnsetImpenetrable();
    if (c.flags & PR_CONTACT_TOI) r.SetToi({UnitIntervalFF<Real>(c.toi)}); else r.SetToi({});
    r.SetToiCount(static_cast<Contact::substep_type>(c.toi_count));
    return r;
}

} // namespace

// ------------------------------------------------------------------------------------------------
struct CudaWorld::Impl {
    pr_world* h = nullptr;
    WorldConf conf;
    unsigned toiMode = PR_TOI_PARALLEL;
    bool locked = false;
    // host mirrors of entity values
    mutable std::vector<Body> bodies;
    std::vector<char> bodyDirty;   // host value newer than the device row
    std::vector<char> bodyLive;    // 0: destroyed (its id is handed out again by the device, last freed first)
    std::vector<BodyID> bodyOrder; // m_bodies: live ids in creation order
    std::vector<char> shapeLive;
    std::vector<BodyID> bodiesForSync; // bodies whose transformation was set since the last Step (m_bodiesForSync)
    bool anyBodyDirty = false;
    mutable bool bodiesStale = false; // device rows newer than the host values (after a Step)
    std::vector<Shape> shapes;
    std::vector<Joint> joints;         // an empty Joint = destroyed (its id is in freeJoints and is handed out again, last freed first)
    std::vector<std::uint32_t> freeJoints;
    std::vector<JointID> jointOrder;   // m_joints: creation order, erased on destroy (GetJoints(world), :1132-1157, :1215-1227)
    std::vector<std::vector<std::pair<BodyID, JointID>>> bodyJoints; // m_bodyJoints: (other body, joint) in Add order (:1159-1175)
    // lazily downloaded contact state
    mutable bool contactsStale = true;
    mutable std::vector<KeyedContactID> contactIds;            // m_contacts order
    mutable std::vector<Contact> contactBuf;                   // indexed by ContactID
    mutable std::vector<Manifold> manifoldBuf;
    mutable std::vector<char> contactLive;
    mutable std::vector<std::vector<std::tuple<ContactKey, ContactID>>> bodyContacts;
    // lazily materialised DynamicTree (GetTree): leaves created in tree-id order replay the id sequence
    mutable bool treeStale = true;
    mutable DynamicTree tree;
    // listeners: begin/end/post-solve are delivered after the device step from pr_events (deliverEvents)
    ShapeFunction shapeDestruction; BodyShapeFunction detach; JointFunction jointDestruction;
    ContactFunction beginContact, endContact; ContactManifoldFunction preSolve; ContactImpulsesFunction postSolve;
    std::uint32_t eventCap = 1u << 16;
    std::uint64_t droppedEvents = 0; // listener calls lost to a full device event buffer (it is grown after such a step)

    void updateEventMask() noexcept
    {
        const std::uint32_t mask = (beginContact ? PR_EVENT_BEGIN : 0u) | (endContact ? PR_EVENT_END : 0u) | (preSolve ? PR_EVENT_PRE_SOLVE : 0u) |
                                   (postSolve ? PR_EVENT_POST_SOLVE : 0u);
        if (h) pr_set_event_mask(h, mask, eventCap);
    }
    // The listeners of the step that just ran, in (kind, contact id) order: all begin calls, all end calls, pre-solve, then post-solve.
    // (The reference interleaves them inside Step; a listener here sees the world AFTER the step -- a pre-solve listener gets the
    // contact id and the manifold the contact had before its update, like in the reference, but cannot change this step's solve.)
    void deliverEvents()
    {
        if (!beginContact && !endContact && !postSolve && !preSolve) return;
        for (;;) { // a step that overflowed the buffers: grow them for the next step, deliver what was kept
            std::vector<pr_contact_event> ce(eventCap);
            std::vector<pr_impulse_event> ie(postSolve ? eventCap : 0u);
            std::vector<pr_presolve_event> pe(preSolve ? eventCap : 0u);
            std::uint32_t nc = 0, ni = 0, np = 0;
            int rc = PR_OK;
            if (beginContact || endContact) rc = pr_events(h, PR_EVENT_BEGIN | PR_EVENT_END, ce.data(), eventCap, &nc);
            int rc2 = PR_OK;
            if (postSolve) rc2 = pr_events(h, PR_EVENT_POST_SOLVE, ie.data(), eventCap, &ni);
            int rc3 = PR_OK;
            if (preSolve) rc3 = pr_events(h, PR_EVENT_PRE_SOLVE, pe.data(), eventCap, &np);
            const bool overflow = rc == PR_ERR_LENGTH || rc2 == PR_ERR_LENGTH || rc3 == PR_ERR_LENGTH;
            if (!overflow) { check(rc, "pr_events"); check(rc2, "pr_events"); check(rc3, "pr_events"); }
            const std::uint32_t keptC = std::min(nc, eventCap), keptI = std::min(ni, eventCap), keptP = std::min(np, eventCap);
            // like inside the reference's Step, the world is locked while listeners run: mutators throw WrongState (:2156)
            struct Lock { bool& f; explicit Lock(bool& b) : f(b) { f = true; } ~Lock() { f = false; } } lock{locked};
            for (std::uint32_t i = 0; i < keptC; ++i) {
                const auto id = MakeId<ContactID>(ce[i].contact);
                if (ce[i].kind == PR_EVENT_BEGIN) { if (beginContact) beginContact(id); }
                else if (endContact) endContact(id);
            }
            for (std::uint32_t i = 0; i < keptP; ++i) preSolve(MakeId<ContactID>(pe[i].contact), ToManifold(pe[i].old_manifold));
            for (std::uint32_t i = 0; i < keptI; ++i) {
                ContactImpulsesList imp;
                for (std::uint32_t j = 0; j < ie[i].count && j < 2u; ++j) imp.AddEntry(Momentum{ie[i].normal[j] * NewtonSecond}, Momentum{ie[i].tangent[j] * NewtonSecond});
                postSolve(MakeId<ContactID>(ie[i].contact), imp, static_cast<unsigned>(ie[i].iterations));
            }
            if (overflow) { // the step is applied and its kept events are delivered: count what was dropped, hold more from now on
                droppedEvents += (nc > keptC ? nc - keptC : 0u) + (ni > keptI ? ni - keptI : 0u) + (np > keptP ? np - keptP : 0u);
                eventCap = std::max(std::max(nc, ni), np) * 2u;
                updateEventMask();
            }
            return;
        }
    }

    void check(int rc, const char* what) const
    {
        if (rc == PR_OK) return;
        const std::string msg = std::string(what) + ": " + (h ? pr_last_error(h) : pr_last_error(nullptr));
        switch (rc) {
        case PR_ERR_INVALID_ARGUMENT: throw InvalidArgument(msg.c_str());
        case PR_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
        case PR_ERR_WRONG_STATE: throw WrongState(msg.c_str());
        case PR_ERR_LENGTH: throw LengthError(msg.c_str());
        case PR_ERR_UNSUPPORTED: throw InvalidArgument(msg.c_str());
        default: throw std::runtime_error(msg);
        }
    }
    void create(const WorldConf& c)
    {
        conf = c;
        pr_world_conf wc;
        pr_world_conf_default(&wc);
        wc.min_vertex_radius = static_cast<float>(Real(c.vertexRadius.GetMin()));
        wc.max_vertex_radius = static_cast<float>(Real(c.vertexRadius.GetMax()));
        wc.contact_capacity = c.contactCapacity;
        wc.proxy_capacity = c.proxyCapacity;
        check(pr_create(&wc, &h), "pr_create");
        pr_set_toi_mode(h, toiMode);
    }
    void refreshBodies() const
    { // device rows -> host values, once after each Step
        if (!bodiesStale) return;
        std::vector<pr_body> rows(bodies.size());
        if (!rows.empty()) check(pr_get_bodies(h, 0, static_cast<std::uint32_t>(rows.size()), rows.data()), "pr_get_bodies");
        for (std::size_t i = 0; i < rows.size(); ++i) if (!bodyDirty[i] && bodyLive[i]) FromRow(rows[i], bodies[i]);
        bodiesStale = false;
    }
    void uploadDirtyBodies()
    {
        if (!anyBodyDirty) return;
        std::vector<std::uint32_t> ids; std::vector<pr_body> rows;
        for (std::size_t i = 0; i < bodies.size(); ++i) if (bodyDirty[i] && bodyLive[i]) { ids.push_back(static_cast<std::uint32_t>(i)); rows.push_back(ToRow(bodies[i])); bodyDirty[i] = 0; }
        if (!ids.empty()) check(pr_set_bodies(h, static_cast<std::uint32_t>(ids.size()), ids.data(), rows.data()), "pr_set_bodies");
        anyBodyDirty = false;
    }
    void refreshContacts() const
    {
        if (!contactsStale) return;
        pr_counts n{};
        pr_get_counts(h, &n);
        std::vector<pr_contact> cs(std::max<std::uint32_t>(1u, n.contacts));
        std::vector<pr_manifold> ms(cs.size());
        std::uint32_t k = 0;
        check(pr_get_contacts(h, static_cast<std::uint32_t>(cs.size()), cs.data(), ms.data(), &k), "pr_get_contacts");
        contactIds.clear();
        contactBuf.assign(n.contact_id_range, Contact{});
        manifoldBuf.assign(n.contact_id_range, Manifold{});
        contactLive.assign(n.contact_id_range, 0);
        bodyContacts.assign(bodies.size(), {});
        for (std::uint32_t i = 0; i < k; ++i) {
            const auto key = ContactKey{static_cast<ContactCounter>(cs[i].key_min), static_cast<ContactCounter>(cs[i].key_max)};
            const auto id = MakeId<ContactID>(cs[i].id);
            contactIds.emplace_back(key, id);
            contactBuf[cs[i].id] = ToContact(cs[i]);
            manifoldBuf[cs[i].id] = ToManifold(ms[i]);
            contactLive[cs[i].id] = 1;
            bodyContacts[cs[i].body_a].emplace_back(key, id); // m_bodyContacts: append order == m_contacts order
            bodyContacts[cs[i].body_b].emplace_back(key, id);
        }
        for (std::size_t i = 0; i < contactBuf.size(); ++i) if (!contactLive[i]) contactBuf[i].SetDestroyed(); // (m_contactBuffer.Free(id).SetDestroyed(), :2277)
        contactsStale = false;
    }
    void mutated() { contactsStale = true; treeStale = true; }
};

CudaWorld::CudaWorld(const WorldConf& conf) : m{new Impl} {
    try { m->create(conf); } catch (...) { delete m; m = nullptr; throw; }
}
CudaWorld::CudaWorld(const CudaWorld& o) : m{new Impl}
{
    try {
        o.m->refreshBodies();
        const_cast<CudaWorld&>(o).m->uploadDirtyBodies();
        m->conf = o.m->conf; m->toiMode = o.m->toiMode;
        m->check(pr_clone(o.m->h, &m->h), "pr_clone");
        m->bodies = o.m->bodies; m->bodyDirty = o.m->bodyDirty; m->anyBodyDirty = o.m->anyBodyDirty; m->bodiesForSync = o.m->bodiesForSync;
        m->bodyLive = o.m->bodyLive; m->bodyOrder = o.m->bodyOrder; m->shapeLive = o.m->shapeLive;
        m->bodiesStale = false; m->shapes = o.m->shapes; m->joints = o.m->joints; m->freeJoints = o.m->freeJoints; m->jointOrder = o.m->jointOrder; m->bodyJoints = o.m->bodyJoints;
        m->shapeDestruction = o.m->shapeDestruction; m->detach = o.m->detach; m->jointDestruction = o.m->jointDestruction;
        m->beginContact = o.m->beginContact; m->endContact = o.m->endContact; m->preSolve = o.m->preSolve; m->postSolve = o.m->postSolve;
        m->eventCap = o.m->eventCap; m->updateEventMask();
    } catch (...) { if (m->h) pr_destroy(m->h); delete m; m = nullptr; throw; }
}
CudaWorld::CudaWorld(CudaWorld&& o) noexcept : m{o.m} { o.m = nullptr; }
CudaWorld& CudaWorld::operator=(const CudaWorld& o) { if (this != &o) { CudaWorld tmp{o}; std::swap(m, tmp.m); } return *this; }
CudaWorld& CudaWorld::operator=(CudaWorld&& o) noexcept { if (this != &o) { std::swap(m, o.m); } return *this; }
CudaWorld::~CudaWorld() noexcept { if (m) { if (m->h) pr_destroy(m->h); delete m; } }
std::uint64_t CudaWorld::GetDroppedEventCount() const noexcept { return m ? m->droppedEvents : 0u; }
void CudaWorld::SetToiMode(ToiMode mode) { m->toiMode = static_cast<unsigned>(mode); m->check(pr_set_toi_mode(m->h, m->toiMode), "pr_set_toi_mode"); }

bool operator==(const CudaWorld& a, const CudaWorld& b)
{ // like AabbTreeWorld's operator== (:927-947): essential state + same touching contacts
    if (a.m == b.m) return true;
    if (!a.m || !b.m) return false;
    a.m->refreshBodies(); b.m->refreshBodies();
    if (a.m->bodies.size() != b.m->bodies.size() || a.m->shapes.size() != b.m->shapes.size()) return false;
    for (std::size_t i = 0; i < a.m->bodies.size(); ++i) if (!(a.m->bodies[i] == b.m->bodies[i])) return false;
    for (std::size_t i = 0; i < a.m->shapes.size(); ++i) if (!(a.m->shapes[i] == b.m->shapes[i])) return false;
    a.m->refreshContacts(); b.m->refreshContacts();
    if (a.m->contactIds.size() != b.m->contactIds.size()) return false;
    for (std::size_t i = 0; i < a.m->contactIds.size(); ++i) {
        const auto ia = to_underlying(std::get<ContactID>(a.m->contactIds[i])), ib = to_underlying(std::get<ContactID>(b.m->contactIds[i]));
        if (a.m->contactBuf[ia].IsTouching() != b.m->contactBuf[ib].IsTouching()) return false;
    }
    return true;
}
bool operator!=(const CudaWorld& a, const CudaWorld& b) { return !(a == b); }

// ---- listeners
void SetShapeDestructionListener(CudaWorld& w, ShapeFunction l) noexcept { w.m->shapeDestruction = std::move(l); }
void SetDetachListener(CudaWorld& w, BodyShapeFunction l) noexcept { w.m->detach = std::move(l); }
void SetJointDestructionListener(CudaWorld& w, JointFunction l) noexcept { w.m->jointDestruction = std::move(l); }
void SetBeginContactListener(CudaWorld& w, ContactFunction l) noexcept { w.m->beginContact = std::move(l); w.m->updateEventMask(); }
void SetEndContactListener(CudaWorld& w, ContactFunction l) noexcept { w.m->endContact = std::move(l); w.m->updateEventMask(); }
// pre-solve: recorded on the device like the others and delivered after the step (contact id + the manifold before the update);
// what the listener does to the contact takes effect from the next step on (include/playrho_b200.h, "contact events")
void SetPreSolveContactListener(CudaWorld& w, ContactManifoldFunction l) noexcept { w.m->preSolve = std::move(l); w.m->updateEventMask(); }
void SetPostSolveContactListener(CudaWorld& w, ContactImpulsesFunction l) noexcept { w.m->postSolve = std::move(l); w.m->updateEventMask(); }

// ---- miscellaneous
std::optional<pmr::StatsResource::Stats> GetResourceStats(const CudaWorld&) noexcept { return {}; }

void Clear(CudaWorld& w) noexcept
{ // Clear(AabbTreeWorld&), :949-1017 (destruction listeners are told about joints and shapes first)
    if (!w.m || !w.m->h) return;
    if (w.m->jointDestruction) // every live joint (:955-965); an exception of one call does not stop the others
        for (const auto jid : std::vector<JointID>(w.m->jointOrder)) { // (m_joints order; a copy: a listener may not mutate, but stay safe)
            try { w.m->jointDestruction(jid); } catch (...) {}
        }
    if (w.m->shapeDestruction) // every non-empty shape slot (:966-980)
        for (std::size_t i = 0; i < w.m->shapes.size(); ++i) {
            if (w.m->shapes[i] == Shape{}) continue;
            try { w.m->shapeDestruction(MakeId<ShapeID>(static_cast<std::uint32_t>(i))); } catch (...) {}
        }
    pr_clear(w.m->h);
    w.m->bodies.clear(); w.m->bodyDirty.clear(); w.m->bodyLive.clear(); w.m->bodyOrder.clear(); w.m->shapeLive.clear(); w.m->anyBodyDirty = false; w.m->bodiesStale = false; w.m->bodiesForSync.clear();
    w.m->shapes.clear(); w.m->joints.clear(); w.m->freeJoints.clear(); w.m->jointOrder.clear(); w.m->bodyJoints.clear();
    w.m->mutated();
}

StepStats Step(CudaWorld& w, const StepConf& conf)
{ // Step(AabbTreeWorld&, const StepConf&), :2144-2219
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    w.m->refreshBodies();       // (host values the user may still read-modify-write must be current)
    w.m->uploadDirtyBodies();
    const pr_step_conf c = ToConf(conf);
    pr_step_stats st{};
    w.m->locked = true;
    const int rc = pr_step(w.m->h, &c, &st);
    w.m->locked = false;
    w.m->check(rc, "pr_step");
    w.m->bodiesStale = true;
    w.m->bodiesForSync.clear();
    w.m->mutated();
    w.m->deliverEvents();
    return ToStats(st);
}
bool IsStepComplete(const CudaWorld& w) noexcept { return pr_is_step_complete(w.m->h) != 0; }
bool GetSubStepping(const CudaWorld& w) noexcept { return pr_get_sub_stepping(w.m->h) != 0; }
void SetSubStepping(CudaWorld& w, bool flag) noexcept { pr_set_sub_stepping(w.m->h, flag ? 1u : 0u); }
bool IsLocked(const CudaWorld& w) noexcept { return w.m->locked; }
void ShiftOrigin(CudaWorld& w, const Length2& newOrigin)
{ // ShiftOrigin(AabbTreeWorld&, const Length2&), :2221-2243
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    w.m->refreshBodies();
    w.m->uploadDirtyBodies();
    w.m->check(pr_shift_origin(w.m->h, static_cast<float>(Real(GetX(newOrigin) / Meter)), static_cast<float>(Real(GetY(newOrigin) / Meter))), "pr_shift_origin");
    for (auto& j : w.m->joints) if (j.has_value()) ::playrho::d2::ShiftOrigin(j, newOrigin); // (:2237-2240: target joints' targets, pulley joints' ground anchors)
    w.m->bodiesStale = true;
    w.m->mutated();
}
Interval<Positive<Length>> GetVertexRadiusInterval(const CudaWorld& w) noexcept { return w.m->conf.vertexRadius; }
Frequency GetInvDeltaTime(const CudaWorld& w) noexcept { return Frequency{pr_get_inv_dt0(w.m->h)}; }

const DynamicTree& GetTree(const CudaWorld& w) noexcept
{ // materialised on demand from the device's fat AABBs; leaf ids replay DynamicTree's allocation
    try {
        if (w.m->treeStale) {
            pr_counts n{};
            pr_get_counts(w.m->h, &n);
            std::vector<pr_proxy> px(std::max<std::uint32_t>(1u, n.proxies));
            std::uint32_t k = 0;
            pr_get_proxies(w.m->h, static_cast<std::uint32_t>(px.size()), px.data(), &k);
            px.resize(k);
            std::sort(px.begin(), px.end(), [](const pr_proxy& a, const pr_proxy& b) { return a.tree_id < b.tree_id; });
            w.m->tree = DynamicTree{w.m->conf.treeCapacity};
            for (const auto& p : px) {
                const auto aabb = AABB{Length2{p.min_x, p.min_y}, Length2{p.max_x, p.max_y}};
                w.m->tree.CreateLeaf(aabb, Contactable{MakeId<BodyID>(p.body), MakeId<ShapeID>(p.shape), static_cast<ChildCounter>(p.child)});
            }
            w.m->treeStale = false;
        }
    } catch (...) {}
    return w.m->tree;
}

// ---- bodies
BodyCounter GetBodyRange(const CudaWorld& w) noexcept { return static_cast<BodyCounter>(w.m->bodies.size()); }
std::vector<BodyID> GetBodies(const CudaWorld& w)
{
    return w.m->bodyOrder; // m_bodies (creation order; a re-used id sits where its new body was created)
}
std::vector<BodyID> GetBodiesForProxies(const CudaWorld& w) { return w.m->bodiesForSync; } // m_bodiesForSync (:2792, cleared by Step :2181)

BodyID CreateBody(CudaWorld& w, Body body)
{ // CreateBody(AabbTreeWorld&, Body), :1019-1060
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    if (w.m->bodyOrder.size() >= MaxBodies) throw LengthError("CreateBody: operation would exceed MaxBodies");
    for (const auto& sid : body.GetShapes()) (void)AtId(w.m->shapes, sid, noSuchShapeMsg); // :1030 validates shape ids
    const pr_body row = ToRow(body);
    std::vector<std::uint32_t> sh;
    for (const auto& sid : body.GetShapes()) sh.push_back(static_cast<std::uint32_t>(to_underlying(sid)));
    const std::uint32_t offs[2] = {0u, static_cast<std::uint32_t>(sh.size())};
    const std::uint32_t dummy = 0;
    std::uint32_t first = 0;
    w.m->check(pr_create_bodies(w.m->h, 1, &row, offs, sh.empty() ? &dummy : sh.data(), &first), "pr_create_bodies");
    if (first < w.m->bodies.size()) { w.m->bodies[first] = std::move(body); w.m->bodyDirty[first] = 0; w.m->bodyLive[first] = 1; } // a destroyed body's id came back
    else { w.m->bodies.push_back(std::move(body)); w.m->bodyDirty.push_back(0); w.m->bodyLive.push_back(1); }
    w.m->bodyOrder.push_back(MakeId<BodyID>(first));
    w.m->mutated();
    return MakeId<BodyID>(first);
}
Body GetBody(const CudaWorld& w, BodyID id)
{
    (void)AtId(w.m->bodies, id, noSuchBodyMsg);
    w.m->refreshBodies();
    return w.m->bodies[to_underlying(id)];
}
void SetBody(CudaWorld& w, BodyID id, Body value)
{ // SetBody(AabbTreeWorld&, BodyID, Body), :2723-2795
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    auto& elem = AtId(w.m->bodies, id, noSuchBodyMsg);
    w.m->refreshBodies();
    if (IsDestroyed(elem)) throw WasDestroyed{id, "ID is destroyed"};
    if (IsDestroyed(value)) throw InvalidArgument("cannot change is-destroyed value");
    for (const auto& sid : value.GetShapes()) (void)AtId(w.m->shapes, sid, noSuchShapeMsg);
    const auto bid = static_cast<std::uint32_t>(to_underlying(id));
    bool rowUploaded = false;
    if (GetType(elem) != GetType(value)) { // :2738-2758: the body's contacts are destroyed; a body turned static is synchronised by the next Step
        w.m->uploadDirtyBodies();
        const pr_body row = ToRow(value);
        w.m->check(pr_set_body_type(w.m->h, bid, &row), "pr_set_body_type");
        if (GetType(value) == BodyType::Static) w.m->bodiesForSync.push_back(id);
        rowUploaded = true;
        w.m->bodiesStale = true; // (the other bodies of destroyed contacts were woken)
    }
    if (IsEnabled(elem) != IsEnabled(value) || elem.GetShapes() != value.GetShapes()) { // GetOldAndNewShapeIDs, :647-680
        w.m->uploadDirtyBodies(); // (the device reads the body's flags)
        std::vector<std::uint32_t> removed, added;
        if (IsEnabled(elem) && IsEnabled(value)) { // multiset difference, in ShapeID order
            std::map<ShapeID, int> oldmap, newmap;
            for (const auto& i : elem.GetShapes()) { ++oldmap[i]; --newmap[i]; }
            for (const auto& i : value.GetShapes()) { --oldmap[i]; ++newmap[i]; }
            for (const auto& e : oldmap) for (int i = 0; i < e.second; ++i) removed.push_back(static_cast<std::uint32_t>(to_underlying(e.first)));
            for (const auto& e : newmap) for (int i = 0; i < e.second; ++i) added.push_back(static_cast<std::uint32_t>(to_underlying(e.first)));
            if (!removed.empty()) w.m->check(pr_detach_shapes(w.m->h, bid, static_cast<std::uint32_t>(removed.size()), removed.data()), "pr_detach_shapes");
            if (!added.empty()) w.m->check(pr_attach_shapes(w.m->h, bid, static_cast<std::uint32_t>(added.size()), added.data()), "pr_attach_shapes");
        } else { // enabled <-> disabled: the device's shape list follows the value's first (no proxies involved either way), then the switch
            if (IsEnabled(elem)) w.m->check(pr_set_body_enabled(w.m->h, bid, 0u), "pr_set_body_enabled"); // all proxies and contacts go (:674-676)
            if (elem.GetShapes() != value.GetShapes()) {
                std::map<ShapeID, int> oldmap;
                for (const auto& i : elem.GetShapes()) ++oldmap[i];
                for (const auto& i : value.GetShapes()) --oldmap[i];
                for (const auto& e : oldmap) { for (int i = 0; i < e.second; ++i) removed.push_back(static_cast<std::uint32_t>(to_underlying(e.first)));
                                               for (int i = 0; i > e.second; --i) added.push_back(static_cast<std::uint32_t>(to_underlying(e.first))); }
                if (!removed.empty()) w.m->check(pr_detach_shapes(w.m->h, bid, static_cast<std::uint32_t>(removed.size()), removed.data()), "pr_detach_shapes");
                if (!added.empty()) { // (pr_attach_shapes wants an enabled body: attach after the switch below when the value is enabled)
                    if (!IsEnabled(value)) throw InvalidArgument(unsupportedMsg);
                }
            }
            if (IsEnabled(value)) {
                w.m->check(pr_set_body_enabled(w.m->h, bid, 1u), "pr_set_body_enabled"); // every attached shape is queued (:671-673)
                if (!added.empty()) w.m->check(pr_attach_shapes(w.m->h, bid, static_cast<std::uint32_t>(added.size()), added.data()), "pr_attach_shapes");
            }
        }
        w.m->bodiesStale = true;
    }
    (void)rowUploaded;
    if (GetTransformation(elem) != GetTransformation(value)) w.m->bodiesForSync.push_back(id); // :2787-2793
    elem = std::move(value);
    w.m->bodyDirty[to_underlying(id)] = 1;
    w.m->anyBodyDirty = true;
    w.m->mutated();
}
void Destroy(CudaWorld& w, BodyID id)
{ // Destroy(AabbTreeWorld&, BodyID), :1062-1106
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    const auto& body = AtId(w.m->bodies, id, noSuchBodyMsg);
    const auto bid = static_cast<std::uint32_t>(to_underlying(id));
    if (!w.m->bodyLive[bid]) throw WasDestroyed{id, "ID is destroyed"};
    w.m->refreshBodies();
    w.m->uploadDirtyBodies();
    // the attached joints (:1072-1085), in the order of the body's joint list: the listener hears of each, then the joint is destroyed
    // (its id is freed in that order: the next CreateJoint gets the last of them)
    while (bid < w.m->bodyJoints.size() && !w.m->bodyJoints[bid].empty()) {
        const auto jointID = std::get<JointID>(w.m->bodyJoints[bid].front());
        if (w.m->jointDestruction) w.m->jointDestruction(jointID);
        const auto before = w.m->bodyJoints[bid].size();
        Destroy(w, jointID);
        if (w.m->bodyJoints[bid].size() == before) w.m->bodyJoints[bid].erase(w.m->bodyJoints[bid].begin()); // (a stale entry: never loop on it)
    }
    if (w.m->endContact) { // InternalDestroy (:2249-2252): end-contact for the touching contacts that go with the body
        w.m->refreshContacts();
        const auto list = w.m->bodyContacts[bid];
        for (const auto& kc : list) if (IsTouching(w.m->contactBuf[to_underlying(std::get<ContactID>(kc))])) w.m->endContact(std::get<ContactID>(kc));
    }
    const auto shapes = body.GetShapes();
    w.m->check(pr_destroy_bodies(w.m->h, 1u, &bid), "pr_destroy_bodies");
    if (w.m->detach) for (const auto& sid : shapes) w.m->detach(std::make_pair(id, sid)); // :1100-1104
    // Remove(BodyID), :1046-1060
    w.m->bodiesForSync.erase(std::remove(w.m->bodiesForSync.begin(), w.m->bodiesForSync.end(), id), w.m->bodiesForSync.end());
    w.m->bodyOrder.erase(std::remove(w.m->bodyOrder.begin(), w.m->bodyOrder.end(), id), w.m->bodyOrder.end());
    w.m->bodies[bid] = Body{};
    w.m->bodies[bid].SetDestroyed();
    w.m->bodyDirty[bid] = 0; w.m->bodyLive[bid] = 0;
    w.m->bodiesStale = true; // (bodies of the destroyed contacts and joints were woken)
    w.m->mutated();
}
std::vector<std::pair<BodyID, JointID>> GetJoints(const CudaWorld& w, BodyID id)
{ // m_bodyJoints: (other body, joint) pairs in the order the joints were added to the body
    (void)AtId(w.m->bodies, id, noSuchBodyMsg);
    const auto i = static_cast<std::size_t>(to_underlying(id));
    return i < w.m->bodyJoints.size() ? w.m->bodyJoints[i] : std::vector<std::pair<BodyID, JointID>>{};
}
std::vector<std::tuple<ContactKey, ContactID>> GetContacts(const CudaWorld& w, BodyID id)
{
    (void)AtId(w.m->bodies, id, noSuchBodyMsg);
    w.m->refreshContacts();
    return w.m->bodyContacts[to_underlying(id)];
}
std::vector<ShapeID> GetShapes(const CudaWorld& w, BodyID id) { return AtId(w.m->bodies, id, noSuchBodyMsg).GetShapes(); }

// ---- joints: the device solver runs RevoluteJointConf; other joint types are rejected
JointCounter GetJointRange(const CudaWorld& w) noexcept { return static_cast<JointCounter>(w.m->joints.size()); }
// m_joints / m_bodyJoints bookkeeping on the host: Add(JointID) appends (:1162-1169), Remove(JointID) erases (:1192-1212)
static void AddJointLinks(CudaWorld& w, JointID id, const Joint& joint)
{
    const auto a = GetBodyA(joint), b = GetBodyB(joint);
    auto& bj = w.m->bodyJoints;
    const auto need = [&](BodyID x) { const auto i = static_cast<std::size_t>(to_underlying(x)); if (bj.size() <= i) bj.resize(i + 1u); return i; };
    if (a != InvalidBodyID) bj[need(a)].emplace_back(b, id);
    if (b != InvalidBodyID) bj[need(b)].emplace_back(a, id);
}
static void RemoveJointLinks(CudaWorld& w, JointID id, const Joint& joint)
{
    auto& bj = w.m->bodyJoints;
    for (const auto x : {GetBodyA(joint), GetBodyB(joint)}) {
        if (x == InvalidBodyID) continue;
        const auto i = static_cast<std::size_t>(to_underlying(x));
        if (i >= bj.size()) continue;
        auto& lst = bj[i];
        const auto it = std::find_if(lst.begin(), lst.end(), [&](const auto& e) { return std::get<JointID>(e) == id; });
        if (it != lst.end()) lst.erase(it);
    }
}
std::vector<JointID> GetJoints(const CudaWorld& w) { return w.m->jointOrder; }
static pr_revolute_joint ToJointRow(const RevoluteJointConf& jc)
{
    pr_revolute_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = jc.collideConnected ? 1u : 0u;
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse_x = GetX(jc.impulse); o.impulse_y = GetY(jc.impulse); o.impulse_z = GetZ(jc.impulse);
    o.angular_motor_impulse = jc.angularMotorImpulse; o.reference_angle = jc.referenceAngle;
    o.enable_limit = jc.enableLimit ? 1u : 0u; o.enable_motor = jc.enableMotor ? 1u : 0u;
    o.lower_angle = jc.lowerAngle; o.upper_angle = jc.upperAngle;
    o.motor_speed = jc.motorSpeed; o.max_motor_torque = jc.maxMotorTorque;
    o.limit_state = static_cast<std::uint32_t>(jc.limitState);
    return o;
}
static pr_revolute_joint ToJointRow(const DistanceJointConf& jc)
{ // the same table row viewed as pr_distance_joint (include/playrho_b200.h)
    pr_distance_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_DISTANCE) << 8);
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse = jc.impulse; o.length = jc.length; o.frequency = jc.frequency; o.damping_ratio = jc.dampingRatio;
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const RopeJointConf& jc)
{ // the same table row viewed as pr_rope_joint
    pr_rope_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_ROPE) << 8);
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse = jc.impulse; o.max_length = jc.maxLength; o.limit_state = static_cast<std::uint32_t>(jc.limitState);
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const FrictionJointConf& jc)
{ // the same table row viewed as pr_friction_joint
    pr_friction_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_FRICTION) << 8);
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.linear_impulse_x = GetX(jc.linearImpulse); o.linear_impulse_y = GetY(jc.linearImpulse); o.angular_impulse = jc.angularImpulse;
    o.max_force = jc.maxForce; o.max_torque = jc.maxTorque;
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const MotorJointConf& jc)
{ // the same table row viewed as pr_motor_joint
    pr_motor_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_MOTOR) << 8);
    o.linear_offset = pr_vec2{GetX(jc.linearOffset), GetY(jc.linearOffset)};
    o.linear_impulse_x = GetX(jc.linearImpulse); o.linear_impulse_y = GetY(jc.linearImpulse); o.angular_impulse = jc.angularImpulse;
    o.angular_offset = jc.angularOffset; o.correction_factor = jc.correctionFactor;
    o.max_force = jc.maxForce; o.max_torque = jc.maxTorque;
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const WeldJointConf& jc)
{ // the same table row viewed as pr_weld_joint
    pr_weld_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_WELD) << 8);
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse_x = GetX(jc.impulse); o.impulse_y = GetY(jc.impulse); o.impulse_z = GetZ(jc.impulse);
    o.reference_angle = jc.referenceAngle; o.frequency = jc.frequency; o.damping_ratio = jc.dampingRatio;
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const TargetJointConf& jc)
{ // the same table row viewed as pr_target_joint: one body, the row's body_a repeats body_b
    pr_target_joint o{};
    o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.body_a = (jc.bodyA != InvalidBodyID) ? static_cast<std::uint32_t>(to_underlying(jc.bodyA)) : o.body_b;
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_TARGET) << 8);
    o.target = pr_vec2{GetX(jc.target), GetY(jc.target)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse_x = GetX(jc.impulse); o.impulse_y = GetY(jc.impulse);
    o.frequency = static_cast<float>(Real(Frequency(jc.frequency) / Hertz)); o.damping_ratio = static_cast<float>(Real(jc.dampingRatio));
    o.max_force = static_cast<float>(Real(Force(jc.maxForce) / Newton));
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const PrismaticJointConf& jc)
{ // the same table row viewed as pr_prismatic_joint
    pr_prismatic_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_PRISMATIC) << 8) |
                          (jc.enableLimit ? PR_JOINT_ENABLE_LIMIT : 0u) | (jc.enableMotor ? PR_JOINT_ENABLE_MOTOR : 0u);
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse_x = GetX(jc.impulse); o.impulse_y = GetY(jc.impulse); o.impulse_z = GetZ(jc.impulse);
    o.motor_impulse = jc.motorImpulse; o.reference_angle = jc.referenceAngle;
    o.local_x_axis_a = pr_vec2{jc.localXAxisA.GetX(), jc.localXAxisA.GetY()};
    o.lower_translation = jc.lowerTranslation; o.upper_translation = jc.upperTranslation;
    o.motor_speed = jc.motorSpeed; o.max_motor_force = jc.maxMotorForce;
    o.limit_state = static_cast<std::uint32_t>(jc.limitState);
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const WheelJointConf& jc)
{ // the same table row viewed as pr_wheel_joint
    pr_wheel_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_WHEEL) << 8) | (jc.enableMotor ? PR_JOINT_ENABLE_MOTOR : 0u);
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse = jc.impulse; o.spring_impulse = jc.springImpulse; o.angular_impulse = jc.angularImpulse;
    o.local_x_axis_a = pr_vec2{jc.localXAxisA.GetX(), jc.localXAxisA.GetY()};
    o.frequency = jc.frequency; o.damping_ratio = jc.dampingRatio;
    o.motor_speed = jc.motorSpeed; o.max_motor_torque = jc.maxMotorTorque;
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static pr_revolute_joint ToJointRow(const PulleyJointConf& jc)
{ // the same table row viewed as pr_pulley_joint
    pr_pulley_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_PULLEY) << 8);
    o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
    o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
    o.impulse = jc.impulse;
    o.ground_anchor_a = pr_vec2{GetX(jc.groundAnchorA), GetY(jc.groundAnchorA)};
    o.ground_anchor_b = pr_vec2{GetX(jc.groundAnchorB), GetY(jc.groundAnchorB)};
    o.length_a = jc.lengthA; o.length_b = jc.lengthB; o.ratio = jc.ratio;
    pr_revolute_joint r{};
    static_assert(sizeof(r) == sizeof(o), "joint rows share one table");
    std::memcpy(&r, &o, sizeof(r));
    return r;
}
static void ToGearSide(const GearJointConf::TypeData& td, pr_gear_side& o)
{
    if (std::holds_alternative<GearJointConf::PrismaticData>(td)) {
        const auto& d = std::get<GearJointConf::PrismaticData>(td);
        o.type = PR_GEAR_PRISMATIC;
        o.local_anchor_a = pr_vec2{GetX(d.localAnchorA), GetY(d.localAnchorA)};
        o.local_anchor_b = pr_vec2{GetX(d.localAnchorB), GetY(d.localAnchorB)};
        o.local_axis = pr_vec2{d.localAxis.GetX(), d.localAxis.GetY()};
    } else if (std::holds_alternative<GearJointConf::RevoluteData>(td)) {
        o.type = PR_GEAR_REVOLUTE;
        o.reference_angle = std::get<GearJointConf::RevoluteData>(td).referenceAngle;
    } else o.type = PR_GEAR_NONE;
}
static pr_gear_joint ToGearJoint(const GearJointConf& jc)
{ // a gear joint's value is wider than a row of the joint table: pr_create_gear_joints
    pr_gear_joint o{};
    o.body_a = static_cast<std::uint32_t>(to_underlying(jc.bodyA)); o.body_b = static_cast<std::uint32_t>(to_underlying(jc.bodyB));
    o.body_c = static_cast<std::uint32_t>(to_underlying(jc.bodyC)); o.body_d = static_cast<std::uint32_t>(to_underlying(jc.bodyD));
    o.collide_connected = (jc.collideConnected ? 1u : 0u) | (static_cast<std::uint32_t>(PR_JOINT_GEAR) << 8);
    o.impulse = jc.impulse; o.ratio = jc.ratio; o.constant = jc.constant;
    ToGearSide(jc.typeDataAC, o.ac); ToGearSide(jc.typeDataBD, o.bd);
    return o;
}
// the joint types the device solves (SURVEY a15 + the first of 8(f) rank 1); the others are rejected
static bool ToAnyJointRow(const Joint& def, pr_revolute_joint& row, BodyID& a, BodyID& b)
{
    if (GetType(def) == GetTypeID<RevoluteJointConf>()) { const auto jc = TypeCast<RevoluteJointConf>(def); row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true; }
    if (GetType(def) == GetTypeID<DistanceJointConf>()) { const auto jc = TypeCast<DistanceJointConf>(def); row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true; }
    if (GetType(def) == GetTypeID<FrictionJointConf>()) { const auto jc = TypeCast<FrictionJointConf>(def); row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true; }
    if (GetType(def) == GetTypeID<MotorJointConf>()) { const auto jc = TypeCast<MotorJointConf>(def); row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true; }
    if (GetType(def) == GetTypeID<RopeJointConf>()) { const auto jc = TypeCast<RopeJointConf>(def); row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true; }
    if (GetType(def) == GetTypeID<WeldJointConf>()) { const auto jc = TypeCast<WeldJointConf>(def); row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true; }
    if (GetType(def) == GetTypeID<PrismaticJointConf>()) { // (the device derives localYAxisA from localXAxisA, like the reference's constructor)
        const auto jc = TypeCast<PrismaticJointConf>(def);
        if (jc.localYAxisA != GetRevPerpendicular(jc.localXAxisA)) return false;
        row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true;
    }
    if (GetType(def) == GetTypeID<WheelJointConf>()) {
        const auto jc = TypeCast<WheelJointConf>(def);
        if (jc.localYAxisA != GetRevPerpendicular(jc.localXAxisA)) return false;
        row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true;
    }
    if (GetType(def) == GetTypeID<PulleyJointConf>()) { const auto jc = TypeCast<PulleyJointConf>(def); row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true; }
    if (GetType(def) == GetTypeID<TargetJointConf>()) { // (bodyA is ignored by the target joint's solver)
        const auto jc = TypeCast<TargetJointConf>(def);
        if (jc.bodyB == InvalidBodyID) return false;
        row = ToJointRow(jc); a = jc.bodyA; b = jc.bodyB; return true;
    }
    return false;
}
JointID CreateJoint(CudaWorld& w, Joint def)
{ // CreateJoint(AabbTreeWorld&, Joint), :1132-1157
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    if (w.m->joints.size() >= MaxJoints) throw LengthError("CreateJoint: operation would exceed MaxJoints");
    if (!def.has_value()) throw WasDestroyed{def, "cannot be empty"};
    if (GetType(def) == GetTypeID<GearJointConf>()) { // four bodies: its own entry point
        const auto gc = TypeCast<GearJointConf>(def);
        for (const auto b : {gc.bodyA, gc.bodyB, gc.bodyC, gc.bodyD}) {
            if (b == InvalidBodyID) throw InvalidArgument(unsupportedMsg); // (the reference's gear solver does nothing then; not built here)
            (void)AtId(w.m->bodies, b, noSuchBodyMsg);
        }
        const auto g = ToGearJoint(gc);
        w.m->refreshBodies();
        w.m->uploadDirtyBodies();
        std::uint32_t first = 0;
        w.m->check(pr_create_gear_joints(w.m->h, 1, &g, &first), "pr_create_gear_joints");
        if (first < w.m->joints.size()) {
            w.m->joints[first] = std::move(def);
            w.m->freeJoints.erase(std::remove(w.m->freeJoints.begin(), w.m->freeJoints.end(), first), w.m->freeJoints.end());
        } else w.m->joints.push_back(std::move(def));
        w.m->jointOrder.push_back(MakeId<JointID>(first)); AddJointLinks(w, MakeId<JointID>(first), w.m->joints[first]);
        w.m->bodiesStale = true; w.m->contactsStale = true;
        return MakeId<JointID>(first);
    }
    pr_revolute_joint o{}; BodyID bA{}, bB{};
    if (!ToAnyJointRow(def, o, bA, bB)) throw InvalidArgument(unsupportedMsg);
    if (bA != InvalidBodyID) (void)AtId(w.m->bodies, bA, noSuchBodyMsg); // (:1144-1149)
    if (bB != InvalidBodyID) (void)AtId(w.m->bodies, bB, noSuchBodyMsg);
    w.m->refreshBodies();       // Add(JointID) wakes both bodies on the device: keep the host values in step
    w.m->uploadDirtyBodies();
    std::uint32_t first = 0;
    w.m->check(pr_create_revolute_joints(w.m->h, 1, &o, &first), "pr_create_revolute_joints");
    if (first < w.m->joints.size()) { // a destroyed joint's id came back
        w.m->joints[first] = std::move(def);
        w.m->freeJoints.erase(std::remove(w.m->freeJoints.begin(), w.m->freeJoints.end(), first), w.m->freeJoints.end());
    } else w.m->joints.push_back(std::move(def));
    w.m->jointOrder.push_back(MakeId<JointID>(first)); AddJointLinks(w, MakeId<JointID>(first), w.m->joints[first]);
    w.m->bodiesStale = true; w.m->contactsStale = true; // (Add(JointID, flagForFiltering) flags the pair's contacts, :1170-1174)
    return MakeId<JointID>(first);
}
// the solver temporaries InitVelocity leaves in the JointConf value (pr_get_joint_work; layout per type: pr_joint.cuh)
static UnitVec MakeUnit(float c, float s)
{
    if (c == 0.0f && s == 0.0f) return UnitVec::GetZero();
    try { return UnitVec{Vec2{c, s}}; } catch (...) { return std::get<0>(UnitVec::Get(c, s)); }
}
static Mat33 MakeMat33(const float* m) { return Mat33{Vec3{m[0], m[1], m[2]}, Vec3{m[3], m[4], m[5]}, Vec3{m[6], m[7], m[8]}}; }
static Mass22 MakeMass22(const float* m) { return Mass22{Vector2<Mass>{m[0] * Kilogram, m[1] * Kilogram}, Vector2<Mass>{m[2] * Kilogram, m[3] * Kilogram}}; }
Joint GetJoint(const CudaWorld& w, JointID id)
{ // the solver-written fields (impulses, limit state) come back from the device
    auto& elem = AtId(w.m->joints, id, noSuchJointMsg);
    if (!elem.has_value()) return Joint{}; // GetJoint = At(m_jointBuffer, id) (:1103): a freed slot reads back empty
    pr_revolute_joint r{};
    w.m->check(pr_get_revolute_joints(w.m->h, static_cast<std::uint32_t>(to_underlying(id)), 1, &r), "pr_get_revolute_joints");
    pr_joint_work k{};
    w.m->check(pr_get_joint_work(w.m->h, static_cast<std::uint32_t>(to_underlying(id)), 1, &k), "pr_get_joint_work");
    const bool solved = k.island != 0xFFFFFFFFu; // (never solved: the conf's temporaries stay what the caller gave)
    const auto rA = Length2{k.r_a[0] * Meter, k.r_a[1] * Meter}, rB = Length2{k.r_b[0] * Meter, k.r_b[1] * Meter};
    if (GetType(elem) == GetTypeID<DistanceJointConf>()) {
        auto dc = TypeCast<DistanceJointConf>(elem);
        dc.impulse = Momentum{r.impulse_x * NewtonSecond};
        if (solved) { dc.u = MakeUnit(k.m[0], k.m[1]); dc.rA = rA; dc.rB = rB; dc.invGamma = InvMass{k.m[2] / Kilogram}; dc.bias = LinearVelocity{k.m[3] * MeterPerSecond}; dc.mass = Mass{k.m[4] * Kilogram}; }
        return Joint{dc};
    }
    if (GetType(elem) == GetTypeID<FrictionJointConf>()) {
        auto fc = TypeCast<FrictionJointConf>(elem);
        fc.linearImpulse = Momentum2{r.impulse_x * NewtonSecond, r.impulse_y * NewtonSecond};
        fc.angularImpulse = AngularMomentum{r.impulse_z * NewtonMeterSecond};
        if (solved) { fc.rA = rA; fc.rB = rB; fc.linearMass = MakeMass22(k.m); fc.angularMass = RotInertia{k.angular_mass * SquareMeter * Kilogram / SquareRadian}; }
        return Joint{fc};
    }
    if (GetType(elem) == GetTypeID<MotorJointConf>()) {
        auto mc = TypeCast<MotorJointConf>(elem);
        mc.linearImpulse = Momentum2{r.impulse_x * NewtonSecond, r.impulse_y * NewtonSecond};
        mc.angularImpulse = AngularMomentum{r.impulse_z * NewtonMeterSecond};
        if (solved) {
            mc.rA = rA; mc.rB = rB; mc.linearMass = MakeMass22(k.m); mc.angularMass = RotInertia{k.angular_mass * SquareMeter * Kilogram / SquareRadian};
            mc.linearError = Length2{k.m[4] * Meter, k.m[5] * Meter}; mc.angularError = Angle{k.m[6] * Radian};
        }
        return Joint{mc};
    }
    if (GetType(elem) == GetTypeID<WeldJointConf>()) {
        auto wc = TypeCast<WeldJointConf>(elem);
        wc.impulse = Vec3{r.impulse_x, r.impulse_y, r.impulse_z};
        if (solved) { wc.rA = rA; wc.rB = rB; wc.mass = MakeMat33(k.m); wc.gamma = InvRotInertia{k.angular_mass * SquareRadian / (SquareMeter * Kilogram)}; wc.bias = AngularVelocity{k.ex[0] * RadianPerSecond}; }
        return Joint{wc};
    }
    if (GetType(elem) == GetTypeID<TargetJointConf>()) {
        auto tc = TypeCast<TargetJointConf>(elem);
        tc.impulse = Momentum2{r.impulse_x * NewtonSecond, r.impulse_y * NewtonSecond};
        if (solved) { tc.rB = rB; tc.mass = MakeMass22(k.m); tc.gamma = InvMass{k.angular_mass / Kilogram}; tc.C = LinearVelocity2{k.ex[0] * MeterPerSecond, k.ex[1] * MeterPerSecond}; }
        return Joint{tc};
    }
    if (GetType(elem) == GetTypeID<PrismaticJointConf>()) {
        auto pc = TypeCast<PrismaticJointConf>(elem);
        pr_prismatic_joint p{}; std::memcpy(&p, &r, sizeof(p));
        pc.impulse = Vec3{p.impulse_x, p.impulse_y, p.impulse_z};
        pc.motorImpulse = Momentum{p.motor_impulse * NewtonSecond};
        pc.limitState = static_cast<LimitState>(p.limit_state);
        if (solved) {
            pc.axis = MakeUnit(k.ex[0], k.ex[1]); pc.perp = MakeUnit(k.ex[2], k.ex[3]);
            pc.s1 = Length{k.ex[4] * Meter}; pc.s2 = Length{k.ex[5] * Meter}; pc.a1 = Length{k.ex[6] * Meter}; pc.a2 = Length{k.ex[7] * Meter};
            pc.K = MakeMat33(k.m); pc.motorMass = Mass{k.angular_mass * Kilogram};
        }
        return Joint{pc};
    }
    if (GetType(elem) == GetTypeID<WheelJointConf>()) {
        auto wc = TypeCast<WheelJointConf>(elem);
        pr_wheel_joint p{}; std::memcpy(&p, &r, sizeof(p));
        wc.impulse = Momentum{p.impulse * NewtonSecond};
        wc.springImpulse = Momentum{p.spring_impulse * NewtonSecond};
        wc.angularImpulse = AngularMomentum{p.angular_impulse * NewtonMeterSecond};
        if (solved) {
            wc.ay = MakeUnit(k.m[0], k.m[1]); wc.sAy = Length{k.m[2] * Meter}; wc.sBy = Length{k.m[3] * Meter}; wc.mass = Mass{k.m[4] * Kilogram};
            wc.ax = MakeUnit(k.m[5], k.m[6]); wc.sAx = Length{k.m[7] * Meter}; wc.sBx = Length{k.m[8] * Meter};
            wc.springMass = Mass{k.ex[0] * Kilogram}; wc.bias = LinearVelocity{k.ex[1] * MeterPerSecond}; wc.gamma = InvMass{k.ex[2] / Kilogram};
            wc.angularMass = RotInertia{k.angular_mass * SquareMeter * Kilogram / SquareRadian};
        }
        return Joint{wc};
    }
    if (GetType(elem) == GetTypeID<PulleyJointConf>()) {
        auto pc = TypeCast<PulleyJointConf>(elem);
        pc.impulse = Momentum{r.impulse_x * NewtonSecond};
        if (solved) { pc.uA = MakeUnit(k.m[0], k.m[1]); pc.uB = MakeUnit(k.m[2], k.m[3]); pc.rA = rA; pc.rB = rB; pc.mass = Mass{k.m[4] * Kilogram}; }
        return Joint{pc};
    }
    if (GetType(elem) == GetTypeID<GearJointConf>()) {
        auto gc = TypeCast<GearJointConf>(elem);
        gc.impulse = Momentum{r.impulse_x * NewtonSecond};
        if (solved) {
            gc.JvAC = Vec2{k.m[0], k.m[1]}; gc.JvBD = Vec2{k.m[2], k.m[3]};
            gc.JwA = Length{k.m[4] * Meter}; gc.JwB = Length{k.m[5] * Meter}; gc.JwC = Length{k.m[6] * Meter}; gc.JwD = Length{k.m[7] * Meter}; gc.mass = k.m[8];
        }
        return Joint{gc};
    }
    if (GetType(elem) == GetTypeID<RopeJointConf>()) {
        auto rc = TypeCast<RopeJointConf>(elem);
        rc.impulse = Momentum{r.impulse_x * NewtonSecond};
        rc.limitState = static_cast<LimitState>(r.limit_state);
        if (solved) { rc.u = MakeUnit(k.m[0], k.m[1]); rc.rA = rA; rc.rB = rB; rc.mass = Mass{k.m[4] * Kilogram}; rc.length = Length{k.m[5] * Meter}; }
        return Joint{rc};
    }
    auto jc = TypeCast<RevoluteJointConf>(elem);
    jc.impulse = Vec3{r.impulse_x, r.impulse_y, r.impulse_z};
    jc.angularMotorImpulse = r.angular_motor_impulse;
    jc.limitState = static_cast<LimitState>(r.limit_state);
    if (solved) { jc.rA = rA; jc.rB = rB; jc.mass = MakeMat33(k.m); jc.angularMass = RotInertia{k.angular_mass * SquareMeter * Kilogram / SquareRadian}; }
    return Joint{jc};
}
void SetJoint(CudaWorld& w, JointID id, Joint def)
{ // SetJoint(AabbTreeWorld&, JointID, Joint), :1108-1130 (any row of the joint table; the value may name other bodies or another collideConnected: pr_set_revolute_joints)
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    if (!AtId(w.m->joints, id, noSuchJointMsg).has_value()) throw WasDestroyed{id, "ID is destroyed"};
    if (const auto b = GetBodyA(def); b != InvalidBodyID) (void)AtId(w.m->bodies, b, noSuchBodyMsg);
    if (const auto b = GetBodyB(def); b != InvalidBodyID) (void)AtId(w.m->bodies, b, noSuchBodyMsg);
    if (!def.has_value()) throw WasDestroyed{def, "cannot be empty"};
    pr_revolute_joint o{}; BodyID bA{}, bB{};
    if (GetType(def) == GetTypeID<GearJointConf>()) { // the row part (impulse, ratio, constant) can change; the geared bodies and joint data cannot
        const auto& old = w.m->joints[to_underlying(id)];
        if (GetType(old) != GetTypeID<GearJointConf>()) throw InvalidArgument(unsupportedMsg);
        const auto gn = ToGearJoint(TypeCast<GearJointConf>(def)), go = ToGearJoint(TypeCast<GearJointConf>(old));
        if (gn.body_c != go.body_c || gn.body_d != go.body_d || std::memcmp(&gn.ac, &go.ac, sizeof(gn.ac)) != 0 || std::memcmp(&gn.bd, &go.bd, sizeof(gn.bd)) != 0)
            throw InvalidArgument(unsupportedMsg);
        o.body_a = gn.body_a; o.body_b = gn.body_b; o.collide_connected = gn.collide_connected;
        o.impulse_x = gn.impulse; o.lower_angle = gn.ratio; o.upper_angle = gn.constant;
    } else if (!ToAnyJointRow(def, o, bA, bB)) throw InvalidArgument(unsupportedMsg);
    w.m->refreshBodies();       // Remove(JointID) wakes both bodies on the device: keep the host values in step
    w.m->uploadDirtyBodies();
    const std::uint32_t jid = static_cast<std::uint32_t>(to_underlying(id));
    w.m->check(pr_set_revolute_joints(w.m->h, 1u, &jid, &o), "pr_set_revolute_joints");
    RemoveJointLinks(w, id, w.m->joints[to_underlying(id)]); // Remove(id) ... Add(id): the joint moves to the end of its bodies' lists
    w.m->joints[to_underlying(id)] = std::move(def);
    AddJointLinks(w, id, w.m->joints[to_underlying(id)]);
    w.m->bodiesStale = true; w.m->contactsStale = true;
}
void Destroy(CudaWorld& w, JointID id)
{ // Destroy(AabbTreeWorld&, JointID), :1215-1227
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    auto& elem = AtId(w.m->joints, id, noSuchJointMsg);
    if (!elem.has_value()) return; // not in m_joints any more: nothing happens
    w.m->refreshBodies();       // Remove(JointID) wakes both bodies on the device: keep the host values in step
    w.m->uploadDirtyBodies();
    const std::uint32_t jid = static_cast<std::uint32_t>(to_underlying(id));
    w.m->check(pr_destroy_joints(w.m->h, 1u, &jid), "pr_destroy_joints");
    RemoveJointLinks(w, id, elem);
    w.m->jointOrder.erase(std::remove(w.m->jointOrder.begin(), w.m->jointOrder.end(), id), w.m->jointOrder.end());
    elem = Joint{};
    w.m->freeJoints.push_back(jid);
    w.m->bodiesStale = true; w.m->contactsStale = true; // (Remove(JointID) flags the pair's contacts, :1184-1190)
}

// ---- shapes
ShapeCounter GetShapeRange(const CudaWorld& w) noexcept { return static_cast<ShapeCounter>(w.m->shapes.size()); }
ShapeID CreateShape(CudaWorld& w, Shape def)
{ // CreateShape(AabbTreeWorld&, Shape), :1234-1260; every child flattens to a DistanceProxy
    if (!def.has_value()) throw WasDestroyed{def, "cannot be empty"};
    const auto vmin = w.m->conf.vertexRadius.GetMin();
    const auto vmax = w.m->conf.vertexRadius.GetMax();
    const auto nChildren = GetChildCount(def);
    pr_shape s{};
    s.friction = static_cast<float>(Real(GetFriction(def))); s.restitution = GetRestitution(def);
    s.density = static_cast<float>(Real(GetDensity(def)));
    const auto f = GetFilter(def);
    s.filter_category = f.categoryBits; s.filter_mask = f.maskBits; s.filter_group = f.groupIndex;
    s.is_sensor = IsSensor(def) ? 1 : 0;
    s.child_offset = 0; s.child_count = nChildren;
    std::vector<pr_child> children; std::vector<pr_vec2> verts, normals;
    for (ChildCounter c = 0; c < nChildren; ++c) {
        const auto dp = GetChild(def, c);
        const auto vr = dp.GetVertexRadius();
        if (!(vr >= vmin)) throw InvalidArgument("CreateShape: vertex radius < min");
        if (!(vr <= vmax)) throw InvalidArgument("CreateShape: vertex radius > max");
        pr_child ch{};
        ch.vertex_offset = static_cast<std::uint32_t>(verts.size()); ch.vertex_count = dp.GetVertexCount(); ch.vertex_radius = vr;
        for (VertexCounter v = 0; v < dp.GetVertexCount(); ++v) {
            verts.push_back(pr_vec2{GetX(dp.GetVertex(v)), GetY(dp.GetVertex(v))});
            normals.push_back(dp.GetVertexCount() > 1 ? pr_vec2{dp.GetNormal(v).GetX(), dp.GetNormal(v).GetY()} : pr_vec2{0, 0});
        }
        children.push_back(ch);
    }
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    if (w.m->shapes.size() >= MaxShapes) throw LengthError("CreateShape: operation would exceed MaxShapes");
    std::uint32_t first = 0;
    const pr_child noChild{}; const pr_vec2 noVert{};
    w.m->check(pr_create_shapes(w.m->h, 1, &s, children.empty() ? &noChild : children.data(), static_cast<std::uint32_t>(children.size()),
                                verts.empty() ? &noVert : verts.data(), normals.empty() ? &noVert : normals.data(),
                                static_cast<std::uint32_t>(verts.size()), &first), "pr_create_shapes");
    if (first < w.m->shapes.size()) { w.m->shapes[first] = std::move(def); w.m->shapeLive[first] = 1; } // a destroyed shape's id came back
    else { w.m->shapes.push_back(std::move(def)); w.m->shapeLive.push_back(1); }
    return MakeId<ShapeID>(first);
}
Shape GetShape(const CudaWorld& w, ShapeID id) { return AtId(w.m->shapes, id, noSuchShapeMsg); }
void SetShape(CudaWorld& w, ShapeID id, Shape def)
{
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    auto& shape = AtId(w.m->shapes, id, noSuchShapeMsg);
    if (!def.has_value()) throw WasDestroyed{def, "cannot be empty"};
    // SetShape(AabbTreeWorld&, ShapeID, Shape), :1284-1356
    if (!shape.has_value()) throw WasDestroyed{id, "ID is destroyed"};
    bool geometryChanged = GetChildCount(shape) != GetChildCount(def); // IsGeomChanged, :778-793
    for (ChildCounter c = 0; !geometryChanged && c < GetChildCount(def); ++c) geometryChanged = GetChild(shape, c) != GetChild(def, c);
    w.m->refreshBodies();       // bodies are woken on the device: keep the host values in step
    w.m->uploadDirtyBodies();
    if (!geometryChanged && GetFilter(shape) == GetFilter(def)) {
        w.m->check(pr_set_shape_material(w.m->h, static_cast<std::uint32_t>(to_underlying(id)), static_cast<float>(Real(GetFriction(def))),
                                         static_cast<float>(GetRestitution(def)), static_cast<float>(Real(GetDensity(def) / KilogramPerSquareMeter)),
                                         IsSensor(def) ? 1u : 0u), "pr_set_shape_material");
    } else {
        pr_shape row{};
        row.friction = static_cast<float>(Real(GetFriction(def))); row.restitution = GetRestitution(def);
        row.density = static_cast<float>(Real(GetDensity(def) / KilogramPerSquareMeter));
        const auto f = GetFilter(def);
        row.filter_category = f.categoryBits; row.filter_mask = f.maskBits; row.filter_group = f.groupIndex;
        row.is_sensor = IsSensor(def) ? 1 : 0;
        row.child_offset = 0; row.child_count = GetChildCount(def);
        std::vector<pr_child> children; std::vector<pr_vec2> verts, normals;
        if (geometryChanged) {
            const auto vmin = w.m->conf.vertexRadius.GetMin();
            const auto vmax = w.m->conf.vertexRadius.GetMax();
            for (ChildCounter c = 0; c < GetChildCount(def); ++c) {
                const auto dp = GetChild(def, c);
                const auto vr = dp.GetVertexRadius();
                if (!(vr >= vmin) || !(vr <= vmax)) throw InvalidArgument("SetShape: vertex radius outside the world's interval");
                pr_child ch{};
                ch.vertex_offset = static_cast<std::uint32_t>(verts.size()); ch.vertex_count = dp.GetVertexCount(); ch.vertex_radius = vr;
                for (VertexCounter v = 0; v < dp.GetVertexCount(); ++v) {
                    verts.push_back(pr_vec2{GetX(dp.GetVertex(v)), GetY(dp.GetVertex(v))});
                    normals.push_back(dp.GetVertexCount() > 1 ? pr_vec2{dp.GetNormal(v).GetX(), dp.GetNormal(v).GetY()} : pr_vec2{0, 0});
                }
                children.push_back(ch);
            }
        }
        const pr_child noChild{}; const pr_vec2 noVert{};
        w.m->check(pr_set_shape(w.m->h, static_cast<std::uint32_t>(to_underlying(id)), &row, children.empty() ? &noChild : children.data(),
                                verts.empty() ? &noVert : verts.data(), normals.empty() ? &noVert : normals.data(),
                                static_cast<std::uint32_t>(verts.size()), geometryChanged ? 1u : 0u), "pr_set_shape");
    }
    shape = std::move(def);
    w.m->bodiesStale = true;
    w.m->mutated();
}
void Destroy(CudaWorld& w, ShapeID id)
{ // Destroy(AabbTreeWorld&, ShapeID), :1259-1277: detached from every body that has it, then the id is free
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    (void)AtId(w.m->shapes, id, noSuchShapeMsg);
    w.m->refreshBodies();
    w.m->uploadDirtyBodies();
    const auto sid = static_cast<std::uint32_t>(to_underlying(id));
    w.m->check(pr_destroy_shapes(w.m->h, 1u, &sid), "pr_destroy_shapes");
    for (std::size_t b = 0; b < w.m->bodies.size(); ++b) { // the values follow: body.Detach(id) until it has none left (:1266-1275)
        if (!w.m->bodyLive[b]) continue;
        while (w.m->bodies[b].Detach(id)) {}
    }
    w.m->shapes[sid] = Shape{};
    w.m->shapeLive[sid] = 0;
    w.m->bodiesStale = true;
    w.m->mutated();
}

// ---- contacts
ContactCounter GetContactRange(const CudaWorld& w) noexcept
{
    pr_counts n{};
    pr_get_counts(w.m->h, &n);
    return static_cast<ContactCounter>(n.contact_id_range);
}
std::vector<KeyedContactID> GetContacts(const CudaWorld& w) { w.m->refreshContacts(); return w.m->contactIds; }
Contact GetContact(const CudaWorld& w, ContactID id)
{
    w.m->refreshContacts();
    return AtId(w.m->contactBuf, id, noSuchContactMsg);
}
void SetContact(CudaWorld& w, ContactID id, Contact value)
{ // SetContact(AabbTreeWorld&, ContactID, Contact), :2797-2831 -- same checks, same exception types, then "elem = value"
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    (void)GetBody(w, GetBodyA(value));
    (void)GetBody(w, GetBodyB(value));
    { const auto shapeA = GetShape(w, GetShapeA(value)); (void)GetChild(shapeA, GetChildIndexA(value)); }
    { const auto shapeB = GetShape(w, GetShapeB(value)); (void)GetChild(shapeB, GetChildIndexB(value)); }
    w.m->refreshContacts();
    const auto& elem = AtId(w.m->contactBuf, id, noSuchContactMsg);
    if (!w.m->contactLive[to_underlying(id)]) throw WasDestroyed{id, "ID is destroyed"};
    if (elem.IsDestroyed() != value.IsDestroyed()) throw InvalidArgument("cannot change is-destroyed value");
    if (elem.GetContactableA() != value.GetContactableA()) throw InvalidArgument("cannot change contactable A");
    if (elem.GetContactableB() != value.GetContactableB()) throw InvalidArgument("cannot change contactable B");
    if (IsImpenetrable(elem) != IsImpenetrable(value)) throw InvalidArgument("change body A or B being impenetrable to change impenetrable state");
    if (IsSensor(elem) != IsSensor(value)) throw InvalidArgument("change shape A or B being a sensor to change sensor state");
    if (GetToi(elem) != GetToi(value)) throw InvalidArgument("user may not change the TOI");
    if (GetToiCount(elem) != GetToiCount(value)) throw InvalidArgument("user may not change the TOI count");
    pr_contact row{};
    row.id = static_cast<std::uint32_t>(to_underlying(id));
    row.body_a = static_cast<std::uint32_t>(to_underlying(GetBodyA(value))); row.shape_a = static_cast<std::uint32_t>(to_underlying(GetShapeA(value)));
    row.child_a = static_cast<std::uint32_t>(GetChildIndexA(value));
    row.body_b = static_cast<std::uint32_t>(to_underlying(GetBodyB(value))); row.shape_b = static_cast<std::uint32_t>(to_underlying(GetShapeB(value)));
    row.child_b = static_cast<std::uint32_t>(GetChildIndexB(value));
    row.friction = static_cast<float>(GetFriction(value)); row.restitution = static_cast<float>(GetRestitution(value));
    row.tangent_speed = static_cast<float>(Real(GetTangentSpeed(value) / MeterPerSecond));
    row.toi = HasValidToi(value) ? static_cast<float>(*GetToi(value)) : 1.0f;
    row.toi_count = GetToiCount(value);
    row.flags = (IsTouching(value) ? PR_CONTACT_TOUCHING : 0u) | (IsEnabled(value) ? PR_CONTACT_ENABLED : 0u) | (NeedsFiltering(value) ? PR_CONTACT_FILTER : 0u) |
                (HasValidToi(value) ? PR_CONTACT_TOI : 0u) | (NeedsUpdating(value) ? PR_CONTACT_DIRTY : 0u) | (IsSensor(value) ? PR_CONTACT_SENSOR : 0u) |
                (IsDestroyed(value) ? PR_CONTACT_DESTROYED : 0u) | (IsImpenetrable(value) ? PR_CONTACT_IMPENETRABLE : 0u);
    w.m->check(pr_set_contacts(w.m->h, 1u, &row), "pr_set_contacts");
    w.m->contactBuf[to_underlying(id)] = value;
}
Manifold GetManifold(const CudaWorld& w, ContactID id)
{
    w.m->refreshContacts();
    return AtId(w.m->manifoldBuf, id, "no such manifold");
}
void SetManifold(CudaWorld& w, ContactID id, const Manifold& value)
{ // SetManifold(AabbTreeWorld&, ContactID, const Manifold&), :2833-2846
    if (w.m->locked) throw WrongState(worldIsLockedMsg);
    w.m->refreshContacts();
    const auto& manifold = AtId(w.m->manifoldBuf, id, "no such manifold");
    if (!w.m->contactLive[to_underlying(id)]) throw WasDestroyed{id, "ID is destroyed"};
    if (manifold.GetType() != value.GetType()) throw InvalidArgument("cannot change manifold type");
    if (manifold.GetPointCount() != value.GetPointCount()) throw InvalidArgument("cannot change manifold point count");
    pr_manifold row{};
    const auto ln = value.GetLocalNormal();
    const auto lp = value.GetLocalPoint();
    row.ln_x = static_cast<float>(ln.GetX()); row.ln_y = static_cast<float>(ln.GetY());
    row.lp_x = static_cast<float>(Real(GetX(lp) / Meter)); row.lp_y = static_cast<float>(Real(GetY(lp) / Meter));
    row.cf0 = row.cf1 = Pack(ContactFeature{ContactFeature::e_vertex, InvalidVertex, ContactFeature::e_vertex, InvalidVertex});
    row.type = value.GetType() == Manifold::e_circles ? PR_MANIFOLD_CIRCLES : value.GetType() == Manifold::e_faceA ? PR_MANIFOLD_FACE_A
             : value.GetType() == Manifold::e_faceB ? PR_MANIFOLD_FACE_B : PR_MANIFOLD_UNSET;
    row.count = value.GetPointCount();
    if (row.count > 0) { const auto p = value.GetPoint(0); row.p0_x = static_cast<float>(Real(GetX(p.localPoint) / Meter)); row.p0_y = static_cast<float>(Real(GetY(p.localPoint) / Meter));
                         row.cf0 = Pack(p.contactFeature); row.jn0 = static_cast<float>(Real(p.normalImpulse / NewtonSecond)); row.jt0 = static_cast<float>(Real(p.tangentImpulse / NewtonSecond)); }
    if (row.count > 1) { const auto p = value.GetPoint(1); row.p1_x = static_cast<float>(Real(GetX(p.localPoint) / Meter)); row.p1_y = static_cast<float>(Real(GetY(p.localPoint) / Meter));
                         row.cf1 = Pack(p.contactFeature); row.jn1 = static_cast<float>(Real(p.normalImpulse / NewtonSecond)); row.jt1 = static_cast<float>(Real(p.tangentImpulse / NewtonSecond)); }
    const std::uint32_t cid = static_cast<std::uint32_t>(to_underlying(id));
    w.m->check(pr_set_manifolds(w.m->h, 1u, &cid, &row), "pr_set_manifolds");
    w.m->manifoldBuf[to_underlying(id)] = value;
}

} // namespace playrho::b200
