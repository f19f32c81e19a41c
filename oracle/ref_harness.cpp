// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY. Not product code; never linked into
// libplayrho_b200.so. A C-ABI shim of OURS over the UNMODIFIED reference library (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libPlayRho{,32}.a) so that python tests and
// bench.py's cpu_baseline leg can (a) build scenes with the reference's own shape/mass code,
// (b) run the reference's Step, (c) dump its state in the POD layout of include/playrho_b200.h,
// (d) call the reference's pure stage functions (CollideShapes, ...) on flat arrays.
// Only the reference's PUBLIC API is used (playrho/d2/World*.hpp, Manifold.hpp, ...).
#include <playrho/d2/BasicAPI.hpp>
#include <playrho/d2/AabbTreeWorld.hpp>
#include <playrho/d2/DynamicTree.hpp>
#include <playrho/d2/Manifold.hpp>
#include <playrho/d2/DistanceProxy.hpp>
#include <playrho/d2/WorldManifold.hpp>
#include <playrho/d2/RevoluteJointConf.hpp>
#include <playrho/d2/TimeOfImpact.hpp>
#include <playrho/d2/RayCastOutput.hpp>
#include <playrho/RayCastOpcode.hpp>
#include <playrho/ToiConf.hpp>
#include <playrho/ToiOutput.hpp>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <variant>
#include <vector>

#include "../include/playrho_b200.h"

using namespace playrho;
using namespace playrho::d2;

namespace {

struct RefWorld {
    World world;
    std::string error;
    std::vector<pr_contact_event> evContact; // begin/end listener calls of the last ref_step, call order
    std::vector<pr_impulse_event> evImpulse; // post-solve listener calls of the last ref_step, call order
    std::vector<pr_presolve_event> evPre;    // pre-solve listener calls of the last ref_step, call order
    bool inToiPhase = false;
    explicit RefWorld(const WorldConf& c) : world{c} {}
};

StepConf ToStepConf(const pr_step_conf& c)
{
    StepConf s;
    s.deltaTime = c.delta_time;
    s.dtRatio = c.dt_ratio;
    s.minStillTimeToSleep = c.min_still_time_to_sleep;
    s.linearSlop = c.linear_slop;
    s.angularSlop = c.angular_slop;
    s.regResolutionRate = c.reg_resolution_rate;
    s.regMinSeparation = c.reg_min_separation;
    s.regMinMomentum = c.reg_min_momentum;
    s.toiResolutionRate = c.toi_resolution_rate;
    s.toiMinSeparation = c.toi_min_separation;
    s.toiMinMomentum = c.toi_min_momentum;
    s.targetDepth = c.target_depth;
    s.tolerance = c.tolerance;
    s.velocityThreshold = c.velocity_threshold;
    s.maxTranslation = c.max_translation;
    s.maxRotation = c.max_rotation;
    s.maxLinearCorrection = c.max_linear_correction;
    s.maxAngularCorrection = c.max_angular_correction;
    s.linearSleepTolerance = c.linear_sleep_tolerance;
    s.angularSleepTolerance = c.angular_sleep_tolerance;
    s.displaceMultiplier = c.displace_multiplier;
    s.aabbExtension = c.aabb_extension;
    s.maxCirclesRatio = c.max_circles_ratio;
    s.regVelocityIters = static_cast<StepConf::iteration_type>(c.reg_velocity_iters);
    s.regPositionIters = static_cast<StepConf::iteration_type>(c.reg_position_iters);
    s.toiVelocityIters = static_cast<StepConf::iteration_type>(c.toi_velocity_iters);
    s.toiPositionIters = static_cast<StepConf::iteration_type>(c.toi_position_iters);
    s.maxToiRootIters = static_cast<StepConf::iteration_type>(c.max_toi_root_iters);
    s.maxToiIters = static_cast<StepConf::iteration_type>(c.max_toi_iters);
    s.maxDistanceIters = static_cast<StepConf::iteration_type>(c.max_distance_iters);
    s.maxSubSteps = static_cast<StepConf::iteration_type>(c.max_sub_steps);
    s.doWarmStart = c.do_warm_start != 0;
    s.doToi = c.do_toi != 0;
    s.doBlocksolve = c.do_blocksolve != 0;
    return s;
}

void FromStepConf(const StepConf& s, pr_step_conf& c)
{
    c.delta_time = s.deltaTime;
    c.dt_ratio = s.dtRatio;
    c.min_still_time_to_sleep = s.minStillTimeToSleep;
    c.linear_slop = s.linearSlop;
    c.angular_slop = s.angularSlop;
    c.reg_resolution_rate = s.regResolutionRate;
    c.reg_min_separation = s.regMinSeparation;
    c.reg_min_momentum = s.regMinMomentum;
    c.toi_resolution_rate = s.toiResolutionRate;
    c.toi_min_separation = s.toiMinSeparation;
    c.toi_min_momentum = s.toiMinMomentum;
    c.target_depth = s.targetDepth;
    c.tolerance = s.tolerance;
    c.velocity_threshold = s.velocityThreshold;
    c.max_translation = s.maxTranslation;
    c.max_rotation = s.maxRotation;
    c.max_linear_correction = s.maxLinearCorrection;
    c.max_angular_correction = s.maxAngularCorrection;
    c.linear_sleep_tolerance = s.linearSleepTolerance;
    c.angular_sleep_tolerance = s.angularSleepTolerance;
    c.displace_multiplier = s.displaceMultiplier;
    c.aabb_extension = s.aabbExtension;
    c.max_circles_ratio = s.maxCirclesRatio;
    c.reg_velocity_iters = s.regVelocityIters;
    c.reg_position_iters = s.regPositionIters;
    c.toi_velocity_iters = s.toiVelocityIters;
    c.toi_position_iters = s.toiPositionIters;
    c.max_toi_root_iters = s.maxToiRootIters;
    c.max_toi_iters = s.maxToiIters;
    c.max_distance_iters = s.maxDistanceIters;
    c.max_sub_steps = s.maxSubSteps;
    c.do_warm_start = s.doWarmStart;
    c.do_toi = s.doToi;
    c.do_blocksolve = s.doBlocksolve;
}

void FromStats(const StepStats& s, pr_step_stats& o)
{
    std::memset(&o, 0, sizeof(o));
    o.pre.proxies_created = s.pre.proxiesCreated;
    o.pre.proxies_moved = s.pre.proxiesMoved;
    o.pre.contacts_destroyed = s.pre.contactsDestroyed;
    o.pre.contacts_added = s.pre.contactsAdded;
    o.pre.contacts_updated = s.pre.contactsUpdated;
    o.pre.contacts_skipped = s.pre.contactsSkipped;
    o.reg.min_separation = s.reg.minSeparation;
    o.reg.max_inc_impulse = s.reg.maxIncImpulse;
    o.reg.islands_found = s.reg.islandsFound;
    o.reg.islands_solved = s.reg.islandsSolved;
    o.reg.contacts_added = s.reg.contactsAdded;
    o.reg.bodies_slept = s.reg.bodiesSlept;
    o.reg.proxies_moved = s.reg.proxiesMoved;
    o.reg.sum_pos_iters = s.reg.sumPosIters;
    o.reg.sum_vel_iters = s.reg.sumVelIters;
    o.reg.contacts_updated = s.reg.contactsUpdated;
    o.reg.contacts_skipped = s.reg.contactsSkipped;
    o.reg.max_island_bodies = s.reg.maxIslandBodies;
    o.toi.min_separation = s.toi.minSeparation;
    o.toi.max_inc_impulse = s.toi.maxIncImpulse;
    o.toi.islands_found = s.toi.islandsFound;
    o.toi.islands_solved = s.toi.islandsSolved;
    o.toi.contacts_found = s.toi.contactsFound;
    o.toi.contacts_at_max_sub_steps = s.toi.contactsAtMaxSubSteps;
    o.toi.contacts_updated_toi = s.toi.contactsUpdatedToi;
    o.toi.contacts_updated_touching = s.toi.contactsUpdatedTouching;
    o.toi.contacts_skipped_touching = s.toi.contactsSkippedTouching;
    o.toi.contacts_added = s.toi.contactsAdded;
    o.toi.proxies_moved = s.toi.proxiesMoved;
    o.toi.sum_pos_iters = s.toi.sumPosIters;
    o.toi.sum_vel_iters = s.toi.sumVelIters;
    o.toi.max_dist_iters = s.toi.maxDistIters;
    o.toi.max_toi_iters = s.toi.maxToiIters;
    o.toi.max_root_iters = s.toi.maxRootIters;
}

template <class Conf>
Conf& ApplyCommon(Conf& conf, float friction, float restitution, float density, int sensor,
                  unsigned cat, unsigned mask, int group)
{
    conf.UseFriction(friction).UseRestitution(restitution).UseDensity(density * 1_kgpm2);
    conf.UseIsSensor(sensor != 0);
    Filter f;
    f.categoryBits = static_cast<Filter::bits_type>(cat);
    f.maskBits = static_cast<Filter::bits_type>(mask);
    f.groupIndex = static_cast<Filter::index_type>(group);
    conf.UseFilter(f);
    return conf;
}

pr_contact_feature Pack(const ContactFeature& cf)
{
    return static_cast<uint32_t>(cf.typeA) | (static_cast<uint32_t>(cf.indexA) << 8) |
           (static_cast<uint32_t>(cf.typeB) << 16) | (static_cast<uint32_t>(cf.indexB) << 24);
}

void FromManifold(const Manifold& m, pr_manifold& o)
{
    std::memset(&o, 0, sizeof(o));
    o.type = static_cast<uint32_t>(m.GetType());
    o.count = m.GetPointCount();
    o.cf0 = o.cf1 = Pack(ContactFeature{});
    if (o.type == Manifold::e_unset) {
        return;
    }
    if (o.type != Manifold::e_circles) {
        const auto ln = m.GetLocalNormal();
        o.ln_x = ln.GetX();
        o.ln_y = ln.GetY();
    }
    const auto lp = m.GetLocalPoint();
    o.lp_x = GetX(lp);
    o.lp_y = GetY(lp);
    if (o.count > 0) {
        const auto& p = m.GetPoint(0);
        o.p0_x = GetX(p.localPoint);
        o.p0_y = GetY(p.localPoint);
        o.jn0 = p.normalImpulse;
        o.jt0 = p.tangentImpulse;
        o.cf0 = Pack(p.contactFeature);
    }
    if (o.count > 1) {
        const auto& p = m.GetPoint(1);
        o.p1_x = GetX(p.localPoint);
        o.p1_y = GetY(p.localPoint);
        o.jn1 = p.normalImpulse;
        o.jt1 = p.tangentImpulse;
        o.cf1 = Pack(p.contactFeature);
    }
}

void FromBody(const Body& b, pr_body& o)
{
    const auto xf = GetTransformation(b);
    const auto& sw = GetSweep(b);
    const auto v = GetVelocity(b);
    o.xf_px = GetX(xf.p);
    o.xf_py = GetY(xf.p);
    o.xf_qc = xf.q.GetX();
    o.xf_qs = xf.q.GetY();
    o.pos0_x = GetX(sw.pos0.linear);
    o.pos0_y = GetY(sw.pos0.linear);
    o.pos0_a = sw.pos0.angular;
    o.pos1_x = GetX(sw.pos1.linear);
    o.pos1_y = GetY(sw.pos1.linear);
    o.pos1_a = sw.pos1.angular;
    o.lc_x = GetX(sw.localCenter);
    o.lc_y = GetY(sw.localCenter);
    o.alpha0 = sw.alpha0;
    o.vel_x = GetX(v.linear);
    o.vel_y = GetY(v.linear);
    o.vel_w = v.angular;
    o.acc_x = GetX(GetLinearAcceleration(b));
    o.acc_y = GetY(GetLinearAcceleration(b));
    o.acc_w = GetAngularAcceleration(b);
    o.inv_mass = GetInvMass(b);
    o.inv_rot_i = GetInvRotInertia(b);
    o.lin_damp = GetLinearDamping(b);
    o.ang_damp = GetAngularDamping(b);
    o.under_active_time = GetUnderActiveTime(b);
    uint32_t f = 0;
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
}

template <class ID>
ID MakeId(uint32_t v)
{
    return static_cast<ID>(static_cast<typename ID::underlying_type>(v));
}

} // namespace

extern "C" {

int ref_ids_are_32bit(void) { return sizeof(BodyID) == 4; }

void ref_step_conf_default(pr_step_conf* out) { FromStepConf(StepConf{}, *out); }

void* ref_world_create(float minVertexRadius, float maxVertexRadius)
{
    try {
        auto conf = WorldConf{};
        if (minVertexRadius > 0 && maxVertexRadius > 0) {
            conf.vertexRadius = Interval<Positive<Length>>{Positive<Length>{minVertexRadius},
                                                           Positive<Length>{maxVertexRadius}};
        }
        return new RefWorld(conf);
    } catch (...) {
        return nullptr;
    }
}

void ref_world_destroy(void* h) { delete static_cast<RefWorld*>(h); }

const char* ref_last_error(void* h) { return static_cast<RefWorld*>(h)->error.c_str(); }

#define REF_TRY(h) auto* rw = static_cast<RefWorld*>(h); try {
#define REF_CATCH(bad) } catch (const std::exception& e) { rw->error = e.what(); return bad; }

uint32_t ref_add_disk(void* h, float radius, float lx, float ly, float friction, float restitution,
                      float density, int sensor, unsigned cat, unsigned mask, int group)
{
    REF_TRY(h)
    auto conf = DiskShapeConf{};
    conf.UseRadius(radius).UseLocation(Length2{lx, ly});
    ApplyCommon(conf, friction, restitution, density, sensor, cat, mask, group);
    return to_underlying(CreateShape(rw->world, conf));
    REF_CATCH(PR_INVALID_ID)
}

// Polygon through PolygonShapeConf::Set (convex hull, like any user of the reference would)
uint32_t ref_add_polygon(void* h, const float* xy, uint32_t n, float vertexRadius, float friction,
                         float restitution, float density, int sensor, unsigned cat, unsigned mask,
                         int group)
{
    REF_TRY(h)
    std::vector<Length2> pts;
    for (uint32_t i = 0; i < n; ++i) pts.push_back(Length2{xy[2 * i], xy[2 * i + 1]});
    auto conf = PolygonShapeConf{};
    if (vertexRadius > 0) conf.UseVertexRadius(vertexRadius);
    conf.Set(Span<const Length2>(pts.data(), pts.size()));
    ApplyCommon(conf, friction, restitution, density, sensor, cat, mask, group);
    return to_underlying(CreateShape(rw->world, conf));
    REF_CATCH(PR_INVALID_ID)
}

uint32_t ref_add_box(void* h, float hx, float hy, float cx, float cy, float angle,
                     float vertexRadius, float friction, float restitution, float density,
                     int sensor, unsigned cat, unsigned mask, int group)
{
    REF_TRY(h)
    auto conf = PolygonShapeConf{};
    if (vertexRadius > 0) conf.UseVertexRadius(vertexRadius);
    if (cx == 0 && cy == 0 && angle == 0) conf.SetAsBox(hx, hy);
    else conf.SetAsBox(hx, hy, Length2{cx, cy}, angle);
    ApplyCommon(conf, friction, restitution, density, sensor, cat, mask, group);
    return to_underlying(CreateShape(rw->world, conf));
    REF_CATCH(PR_INVALID_ID)
}

uint32_t ref_add_edge(void* h, float x0, float y0, float x1, float y1, float vertexRadius,
                      float friction, float restitution, float density, int sensor, unsigned cat,
                      unsigned mask, int group)
{
    REF_TRY(h)
    auto conf = EdgeShapeConf{};
    if (vertexRadius > 0) conf.UseVertexRadius(vertexRadius);
    conf.Set(Length2{x0, y0}, Length2{x1, y1});
    ApplyCommon(conf, friction, restitution, density, sensor, cat, mask, group);
    return to_underlying(CreateShape(rw->world, conf));
    REF_CATCH(PR_INVALID_ID)
}

uint32_t ref_add_chain(void* h, const float* xy, uint32_t n, float vertexRadius, float friction,
                       float restitution, float density, int sensor, unsigned cat, unsigned mask,
                       int group)
{
    REF_TRY(h)
    std::vector<Length2> pts;
    for (uint32_t i = 0; i < n; ++i) pts.push_back(Length2{xy[2 * i], xy[2 * i + 1]});
    auto conf = ChainShapeConf{};
    if (vertexRadius > 0) conf.UseVertexRadius(vertexRadius);
    conf.Set(pts);
    ApplyCommon(conf, friction, restitution, density, sensor, cat, mask, group);
    return to_underlying(CreateShape(rw->world, conf));
    REF_CATCH(PR_INVALID_ID)
}

// type: 0 static, 1 kinematic, 2 dynamic (BodyType.hpp). bits: 1 bullet, 2 fixedRotation,
// 4 disallow sleep, 8 start asleep, 16 disabled
uint32_t ref_add_body(void* h, int type, float x, float y, float angle, float vx, float vy, float w,
                      float ax, float ay, float aw, float linDamp, float angDamp, unsigned bits,
                      const uint32_t* shapes, uint32_t nshapes)
{
    REF_TRY(h)
    auto bc = BodyConf{};
    bc.Use(type == 0 ? BodyType::Static : type == 1 ? BodyType::Kinematic : BodyType::Dynamic);
    bc.UseLocation(Length2{x, y}).UseAngle(angle);
    bc.UseLinearVelocity(LinearVelocity2{vx, vy}).UseAngularVelocity(w);
    bc.UseLinearAcceleration(LinearAcceleration2{ax, ay}).UseAngularAcceleration(aw);
    bc.UseLinearDamping(linDamp).UseAngularDamping(angDamp);
    bc.UseBullet((bits & 1u) != 0).UseFixedRotation((bits & 2u) != 0);
    bc.UseAllowSleep((bits & 4u) == 0).UseAwake((bits & 8u) == 0).UseEnabled((bits & 16u) == 0);
    for (uint32_t i = 0; i < nshapes; ++i) bc.Use(MakeId<ShapeID>(shapes[i]));
    return to_underlying(CreateBody(rw->world, bc));
    REF_CATCH(PR_INVALID_ID)
}

uint32_t ref_add_revolute(void* h, uint32_t bodyA, uint32_t bodyB, float ax, float ay,
                          int enableMotor, float motorSpeed, float maxMotorTorque,
                          int collideConnected)
{
    REF_TRY(h)
    auto jc = GetRevoluteJointConf(rw->world, MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB),
                                   Length2{ax, ay});
    jc.enableMotor = enableMotor != 0;
    jc.motorSpeed = motorSpeed;
    jc.maxMotorTorque = maxMotorTorque;
    jc.collideConnected = collideConnected != 0;
    return to_underlying(CreateJoint(rw->world, jc));
    REF_CATCH(PR_INVALID_ID)
}

// every joint as a row of the joint table (revolute and distance joints: PR_JOINT_TYPE says which)
// CreateJoint(world, GetDistanceJointConf(world, bodyA, bodyB, anchorA, anchorB)) with frequency / damping ratio put in
uint32_t ref_add_distance(void* h, uint32_t bodyA, uint32_t bodyB, float ax, float ay, float bx, float by, float frequency, float dampingRatio,
                          int collideConnected)
{
    REF_TRY(h)
    auto jc = GetDistanceJointConf(rw->world, MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB), Length2{ax * 1_m, ay * 1_m}, Length2{bx * 1_m, by * 1_m});
    jc.frequency = frequency * 1_Hz; jc.dampingRatio = dampingRatio; jc.collideConnected = collideConnected != 0;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
// CreateJoint(world, RopeJointConf{bodyA, bodyB}) with local anchors and the maximum length put in
uint32_t ref_add_rope(void* h, uint32_t bodyA, uint32_t bodyB, float lax, float lay, float lbx, float lby, float maxLength, int collideConnected)
{
    REF_TRY(h)
    auto jc = RopeJointConf{MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB)};
    jc.localAnchorA = Length2{lax * 1_m, lay * 1_m}; jc.localAnchorB = Length2{lbx * 1_m, lby * 1_m};
    jc.maxLength = maxLength * 1_m; jc.collideConnected = collideConnected != 0;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
uint32_t ref_add_friction(void* h, uint32_t bodyA, uint32_t bodyB, float ax, float ay, float maxForce, float maxTorque)
{
    REF_TRY(h)
    auto jc = GetFrictionJointConf(rw->world, MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB), Length2{ax * 1_m, ay * 1_m});
    jc.maxForce = maxForce * 1_N; jc.maxTorque = maxTorque * 1_Nm;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
uint32_t ref_add_motor(void* h, uint32_t bodyA, uint32_t bodyB, float ox, float oy, float angularOffset, float maxForce, float maxTorque, float correctionFactor)
{
    REF_TRY(h)
    auto jc = MotorJointConf{MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB), Length2{ox * 1_m, oy * 1_m}, angularOffset * 1_rad};
    jc.maxForce = maxForce * 1_N; jc.maxTorque = maxTorque * 1_Nm; jc.correctionFactor = correctionFactor;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
uint32_t ref_add_weld(void* h, uint32_t bodyA, uint32_t bodyB, float ax, float ay, float frequency, float dampingRatio)
{
    REF_TRY(h)
    auto jc = GetWeldJointConf(rw->world, MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB), Length2{ax * 1_m, ay * 1_m});
    jc.frequency = frequency * 1_Hz; jc.dampingRatio = dampingRatio;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
uint32_t ref_add_target(void* h, uint32_t bodyA, uint32_t bodyB, float tx, float ty, float ax, float ay, float frequency, float dampingRatio, float maxForce)
{ // bodyA: 0xFFFFFFFF = none (the usual target joint), else the conf's bodyA (it only takes part in the island rules)
    REF_TRY(h)
    auto jc = TargetJointConf{};
    if (bodyA != 0xFFFFFFFFu) jc.bodyA = MakeId<BodyID>(bodyA);
    jc.bodyB = MakeId<BodyID>(bodyB);
    jc.target = Length2{tx * 1_m, ty * 1_m};
    jc.localAnchorB = GetLocalPoint(rw->world, jc.bodyB, Length2{ax * 1_m, ay * 1_m});
    jc.frequency = frequency * 1_Hz; jc.dampingRatio = dampingRatio; jc.maxForce = maxForce * 1_N;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
int ref_set_target(void* h, uint32_t joint, float tx, float ty)
{
    REF_TRY(h)
    const auto id = MakeId<JointID>(joint);
    auto jc = TypeCast<TargetJointConf>(GetJoint(rw->world, id));
    jc.target = Length2{tx * 1_m, ty * 1_m};
    SetJoint(rw->world, id, Joint{jc});
    return 0;
    REF_CATCH(-1)
}
uint32_t ref_add_prismatic(void* h, uint32_t bodyA, uint32_t bodyB, float ax, float ay, float dirx, float diry, int enableLimit, float lower, float upper,
                           int enableMotor, float motorSpeed, float maxMotorForce)
{
    REF_TRY(h)
    auto jc = GetPrismaticJointConf(rw->world, MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB), Length2{ax * 1_m, ay * 1_m}, UnitVec::Get(dirx, diry).first);
    jc.enableLimit = enableLimit != 0; jc.lowerTranslation = lower * 1_m; jc.upperTranslation = upper * 1_m;
    jc.enableMotor = enableMotor != 0; jc.motorSpeed = motorSpeed * 1_rad / 1_s; jc.maxMotorForce = maxMotorForce * 1_N;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
uint32_t ref_add_wheel(void* h, uint32_t bodyA, uint32_t bodyB, float ax, float ay, float dirx, float diry, float frequency, float dampingRatio,
                       int enableMotor, float motorSpeed, float maxMotorTorque)
{
    REF_TRY(h)
    auto jc = GetWheelJointConf(rw->world, MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB), Length2{ax * 1_m, ay * 1_m}, UnitVec::Get(dirx, diry).first);
    jc.frequency = frequency * 1_Hz; jc.dampingRatio = dampingRatio;
    jc.enableMotor = enableMotor != 0; jc.motorSpeed = motorSpeed * 1_rad / 1_s; jc.maxMotorTorque = maxMotorTorque * 1_Nm;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
uint32_t ref_add_pulley(void* h, uint32_t bodyA, uint32_t bodyB, float gax, float gay, float gbx, float gby, float ax, float ay, float bx, float by, float ratio)
{
    REF_TRY(h)
    auto jc = GetPulleyJointConf(rw->world, MakeId<BodyID>(bodyA), MakeId<BodyID>(bodyB), Length2{gax * 1_m, gay * 1_m}, Length2{gbx * 1_m, gby * 1_m},
                                 Length2{ax * 1_m, ay * 1_m}, Length2{bx * 1_m, by * 1_m});
    jc.ratio = ratio;
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
uint32_t ref_add_gear(void* h, uint32_t joint1, uint32_t joint2, float ratio)
{
    REF_TRY(h)
    const auto jc = GetGearJointConf(rw->world, MakeId<JointID>(joint1), MakeId<JointID>(joint2), ratio);
    return static_cast<uint32_t>(to_underlying(CreateJoint(rw->world, Joint{jc})));
    REF_CATCH(0xFFFFFFFFu)
}
static void export_gear_side(const GearJointConf::TypeData& td, pr_gear_side& o)
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
int ref_export_gear_joints(void* h, pr_gear_joint* out, uint32_t cap)
{
    REF_TRY(h)
    const auto n = static_cast<uint32_t>(GetJointRange(rw->world));
    const auto live = GetJoints(rw->world);
    for (uint32_t i = 0; i < n && i < cap; ++i) {
        auto& o = out[i];
        std::memset(&o, 0, sizeof(o));
        o.body_a = o.body_b = o.body_c = o.body_d = 0xFFFFFFFFu;
        if (std::find(live.begin(), live.end(), MakeId<JointID>(i)) == live.end()) continue;
        const auto joint = GetJoint(rw->world, MakeId<JointID>(i));
        if (GetType(joint) != GetTypeID<GearJointConf>()) continue;
        const auto gc = TypeCast<GearJointConf>(joint);
        o.body_a = to_underlying(gc.bodyA); o.body_b = to_underlying(gc.bodyB); o.body_c = to_underlying(gc.bodyC); o.body_d = to_underlying(gc.bodyD);
        o.collide_connected = (gc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_GEAR) << 8);
        o.impulse = gc.impulse; o.ratio = gc.ratio; o.constant = gc.constant;
        export_gear_side(gc.typeDataAC, o.ac); export_gear_side(gc.typeDataBD, o.bd);
    }
    return static_cast<int>(n);
    REF_CATCH(-1)
}
int ref_export_revolute_joints(void* h, pr_revolute_joint* out, uint32_t cap)
{
    REF_TRY(h)
    const auto n = static_cast<uint32_t>(GetJointRange(rw->world));
    const auto live = GetJoints(rw->world);
    for (uint32_t i = 0; i < n && i < cap; ++i) {
        auto& o = out[i];
        std::memset(&o, 0, sizeof(o));
        if (std::find(live.begin(), live.end(), MakeId<JointID>(i)) == live.end()) { o.body_a = o.body_b = 0xFFFFFFFFu; continue; } // destroyed
        const auto joint = GetJoint(rw->world, MakeId<JointID>(i));
        if (GetType(joint) == GetTypeID<DistanceJointConf>()) { // the same row, viewed as pr_distance_joint
            const auto dc = TypeCast<DistanceJointConf>(joint);
            o.body_a = to_underlying(dc.bodyA); o.body_b = to_underlying(dc.bodyB);
            o.collide_connected = (dc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_DISTANCE) << 8);
            o.local_anchor_a = pr_vec2{GetX(dc.localAnchorA), GetY(dc.localAnchorA)};
            o.local_anchor_b = pr_vec2{GetX(dc.localAnchorB), GetY(dc.localAnchorB)};
            o.impulse_x = dc.impulse; o.reference_angle = dc.length; o.lower_angle = dc.frequency; o.upper_angle = dc.dampingRatio;
            continue;
        }
        if (GetType(joint) == GetTypeID<FrictionJointConf>()) { // the same row, viewed as pr_friction_joint
            const auto fc = TypeCast<FrictionJointConf>(joint);
            o.body_a = to_underlying(fc.bodyA); o.body_b = to_underlying(fc.bodyB);
            o.collide_connected = (fc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_FRICTION) << 8);
            o.local_anchor_a = pr_vec2{GetX(fc.localAnchorA), GetY(fc.localAnchorA)};
            o.local_anchor_b = pr_vec2{GetX(fc.localAnchorB), GetY(fc.localAnchorB)};
            o.impulse_x = GetX(fc.linearImpulse); o.impulse_y = GetY(fc.linearImpulse); o.impulse_z = fc.angularImpulse;
            o.motor_speed = fc.maxForce; o.max_motor_torque = fc.maxTorque;
            continue;
        }
        if (GetType(joint) == GetTypeID<MotorJointConf>()) { // the same row, viewed as pr_motor_joint
            const auto mc = TypeCast<MotorJointConf>(joint);
            o.body_a = to_underlying(mc.bodyA); o.body_b = to_underlying(mc.bodyB);
            o.collide_connected = (mc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_MOTOR) << 8);
            o.local_anchor_a = pr_vec2{GetX(mc.linearOffset), GetY(mc.linearOffset)};
            o.impulse_x = GetX(mc.linearImpulse); o.impulse_y = GetY(mc.linearImpulse); o.impulse_z = mc.angularImpulse;
            o.reference_angle = mc.angularOffset; o.lower_angle = mc.correctionFactor;
            o.motor_speed = mc.maxForce; o.max_motor_torque = mc.maxTorque;
            continue;
        }
        if (GetType(joint) == GetTypeID<WeldJointConf>()) { // the same row, viewed as pr_weld_joint
            const auto wc = TypeCast<WeldJointConf>(joint);
            o.body_a = to_underlying(wc.bodyA); o.body_b = to_underlying(wc.bodyB);
            o.collide_connected = (wc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_WELD) << 8);
            o.local_anchor_a = pr_vec2{GetX(wc.localAnchorA), GetY(wc.localAnchorA)};
            o.local_anchor_b = pr_vec2{GetX(wc.localAnchorB), GetY(wc.localAnchorB)};
            o.impulse_x = GetX(wc.impulse); o.impulse_y = GetY(wc.impulse); o.impulse_z = GetZ(wc.impulse);
            o.reference_angle = wc.referenceAngle; o.lower_angle = wc.frequency; o.upper_angle = wc.dampingRatio;
            continue;
        }
        if (GetType(joint) == GetTypeID<TargetJointConf>()) { // the same row, viewed as pr_target_joint (body_a repeats body_b)
            const auto tc = TypeCast<TargetJointConf>(joint);
            o.body_b = to_underlying(tc.bodyB); o.body_a = (tc.bodyA != InvalidBodyID) ? to_underlying(tc.bodyA) : o.body_b;
            o.collide_connected = (tc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_TARGET) << 8);
            o.local_anchor_a = pr_vec2{GetX(tc.target), GetY(tc.target)};
            o.local_anchor_b = pr_vec2{GetX(tc.localAnchorB), GetY(tc.localAnchorB)};
            o.impulse_x = GetX(tc.impulse); o.impulse_y = GetY(tc.impulse);
            o.lower_angle = tc.frequency; o.upper_angle = tc.dampingRatio; o.motor_speed = tc.maxForce;
            continue;
        }
        if (GetType(joint) == GetTypeID<PrismaticJointConf>()) { // the same row, viewed as pr_prismatic_joint
            const auto pc = TypeCast<PrismaticJointConf>(joint);
            pr_prismatic_joint p{};
            p.body_a = to_underlying(pc.bodyA); p.body_b = to_underlying(pc.bodyB);
            p.collide_connected = (pc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_PRISMATIC) << 8) |
                                  (pc.enableLimit ? PR_JOINT_ENABLE_LIMIT : 0u) | (pc.enableMotor ? PR_JOINT_ENABLE_MOTOR : 0u);
            p.local_anchor_a = pr_vec2{GetX(pc.localAnchorA), GetY(pc.localAnchorA)};
            p.local_anchor_b = pr_vec2{GetX(pc.localAnchorB), GetY(pc.localAnchorB)};
            p.impulse_x = GetX(pc.impulse); p.impulse_y = GetY(pc.impulse); p.impulse_z = GetZ(pc.impulse);
            p.motor_impulse = pc.motorImpulse; p.reference_angle = pc.referenceAngle;
            p.local_x_axis_a = pr_vec2{pc.localXAxisA.GetX(), pc.localXAxisA.GetY()};
            p.lower_translation = pc.lowerTranslation; p.upper_translation = pc.upperTranslation;
            p.motor_speed = pc.motorSpeed; p.max_motor_force = pc.maxMotorForce;
            p.limit_state = static_cast<uint32_t>(pc.limitState);
            std::memcpy(&o, &p, sizeof(o));
            continue;
        }
        if (GetType(joint) == GetTypeID<WheelJointConf>()) { // the same row, viewed as pr_wheel_joint
            const auto wc = TypeCast<WheelJointConf>(joint);
            pr_wheel_joint p{};
            p.body_a = to_underlying(wc.bodyA); p.body_b = to_underlying(wc.bodyB);
            p.collide_connected = (wc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_WHEEL) << 8) | (wc.enableMotor ? PR_JOINT_ENABLE_MOTOR : 0u);
            p.local_anchor_a = pr_vec2{GetX(wc.localAnchorA), GetY(wc.localAnchorA)};
            p.local_anchor_b = pr_vec2{GetX(wc.localAnchorB), GetY(wc.localAnchorB)};
            p.impulse = wc.impulse; p.spring_impulse = wc.springImpulse; p.angular_impulse = wc.angularImpulse;
            p.local_x_axis_a = pr_vec2{wc.localXAxisA.GetX(), wc.localXAxisA.GetY()};
            p.frequency = wc.frequency; p.damping_ratio = wc.dampingRatio;
            p.motor_speed = wc.motorSpeed; p.max_motor_torque = wc.maxMotorTorque;
            std::memcpy(&o, &p, sizeof(o));
            continue;
        }
        if (GetType(joint) == GetTypeID<PulleyJointConf>()) { // the same row, viewed as pr_pulley_joint
            const auto pc = TypeCast<PulleyJointConf>(joint);
            pr_pulley_joint p{};
            p.body_a = to_underlying(pc.bodyA); p.body_b = to_underlying(pc.bodyB);
            p.collide_connected = (pc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_PULLEY) << 8);
            p.local_anchor_a = pr_vec2{GetX(pc.localAnchorA), GetY(pc.localAnchorA)};
            p.local_anchor_b = pr_vec2{GetX(pc.localAnchorB), GetY(pc.localAnchorB)};
            p.impulse = pc.impulse;
            p.ground_anchor_a = pr_vec2{GetX(pc.groundAnchorA), GetY(pc.groundAnchorA)};
            p.ground_anchor_b = pr_vec2{GetX(pc.groundAnchorB), GetY(pc.groundAnchorB)};
            p.length_a = pc.lengthA; p.length_b = pc.lengthB; p.ratio = pc.ratio;
            std::memcpy(&o, &p, sizeof(o));
            continue;
        }
        if (GetType(joint) == GetTypeID<GearJointConf>()) { // the table row of a gear joint (the whole value: ref_export_gear_joints)
            const auto gc = TypeCast<GearJointConf>(joint);
            o.body_a = to_underlying(gc.bodyA); o.body_b = to_underlying(gc.bodyB);
            o.collide_connected = (gc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_GEAR) << 8);
            o.impulse_x = gc.impulse; o.lower_angle = gc.ratio; o.upper_angle = gc.constant;
            continue;
        }
        if (GetType(joint) == GetTypeID<RopeJointConf>()) { // the same row, viewed as pr_rope_joint
            const auto rc = TypeCast<RopeJointConf>(joint);
            o.body_a = to_underlying(rc.bodyA); o.body_b = to_underlying(rc.bodyB);
            o.collide_connected = (rc.collideConnected ? 1u : 0u) | (static_cast<uint32_t>(PR_JOINT_ROPE) << 8);
            o.local_anchor_a = pr_vec2{GetX(rc.localAnchorA), GetY(rc.localAnchorA)};
            o.local_anchor_b = pr_vec2{GetX(rc.localAnchorB), GetY(rc.localAnchorB)};
            o.impulse_x = rc.impulse; o.reference_angle = rc.maxLength; o.limit_state = static_cast<uint32_t>(rc.limitState);
            continue;
        }
        const auto jc = TypeCast<RevoluteJointConf>(joint);
        o.body_a = to_underlying(jc.bodyA); o.body_b = to_underlying(jc.bodyB);
        o.collide_connected = jc.collideConnected ? 1u : 0u;
        o.local_anchor_a = pr_vec2{GetX(jc.localAnchorA), GetY(jc.localAnchorA)};
        o.local_anchor_b = pr_vec2{GetX(jc.localAnchorB), GetY(jc.localAnchorB)};
        o.impulse_x = GetX(jc.impulse); o.impulse_y = GetY(jc.impulse); o.impulse_z = GetZ(jc.impulse);
        o.angular_motor_impulse = jc.angularMotorImpulse;
        o.reference_angle = jc.referenceAngle;
        o.enable_limit = jc.enableLimit ? 1u : 0u; o.enable_motor = jc.enableMotor ? 1u : 0u;
        o.lower_angle = jc.lowerAngle; o.upper_angle = jc.upperAngle;
        o.motor_speed = jc.motorSpeed; o.max_motor_torque = jc.maxMotorTorque;
        o.limit_state = static_cast<uint32_t>(jc.limitState);
    }
    return static_cast<int>(n);
    REF_CATCH(-1)
}

// SetShape through the reference's API: the current shape with friction / restitution / density / sensor put in
int ref_set_shape_material(void* h, uint32_t shape, float friction, float restitution, float density, uint32_t is_sensor)
{
    REF_TRY(h)
    const auto id = MakeId<ShapeID>(shape);
    auto s = GetShape(rw->world, id);
    SetFriction(s, friction);
    SetRestitution(s, restitution);
    SetDensity(s, density * 1_kgpm2);
    SetSensor(s, is_sensor != 0);
    SetShape(rw->world, id, s);
    return 0;
    REF_CATCH(-1)
}

int ref_set_sub_stepping(void* h, int flag) { REF_TRY(h) SetSubStepping(rw->world, flag != 0); return 0; REF_CATCH(-1) }
int ref_is_step_complete(void* h) { return IsStepComplete(static_cast<RefWorld*>(h)->world) ? 1 : 0; }

// Attach(World&, BodyID, ShapeID, resetMassData = true), WorldBody.cpp
int ref_attach_shape(void* h, uint32_t body, uint32_t shape)
{
    REF_TRY(h)
    Attach(rw->world, MakeId<BodyID>(body), MakeId<ShapeID>(shape));
    return 0;
    REF_CATCH(-1)
}

// the structural mutators, through the reference's own API (World.hpp / WorldBody.hpp / WorldShape.hpp)
int ref_destroy_body(void* h, uint32_t body) { REF_TRY(h) Destroy(rw->world, MakeId<BodyID>(body)); return 0; REF_CATCH(-1) }
int ref_destroy_shape(void* h, uint32_t shape) { REF_TRY(h) Destroy(rw->world, MakeId<ShapeID>(shape)); return 0; REF_CATCH(-1) }
int ref_detach_shape(void* h, uint32_t body, uint32_t shape) { REF_TRY(h) Detach(rw->world, MakeId<BodyID>(body), MakeId<ShapeID>(shape)); return 0; REF_CATCH(-1) }
int ref_set_body_enabled(void* h, uint32_t body, int flag) { REF_TRY(h) SetEnabled(rw->world, MakeId<BodyID>(body), flag != 0); return 0; REF_CATCH(-1) }
int ref_set_body_type(void* h, uint32_t body, int type)
{
    REF_TRY(h)
    SetType(rw->world, MakeId<BodyID>(body), type == 0 ? BodyType::Static : type == 1 ? BodyType::Kinematic : BodyType::Dynamic);
    return 0;
    REF_CATCH(-1)
}
int ref_set_shape_filter(void* h, uint32_t shape, uint32_t category, uint32_t mask, int group)
{
    REF_TRY(h)
    Filter f; f.categoryBits = static_cast<Filter::bits_type>(category); f.maskBits = static_cast<Filter::bits_type>(mask); f.groupIndex = static_cast<Filter::index_type>(group);
    SetFilterData(rw->world, MakeId<ShapeID>(shape), f);
    return 0;
    REF_CATCH(-1)
}
int ref_translate_shape(void* h, uint32_t shape, float dx, float dy) { REF_TRY(h) Translate(rw->world, MakeId<ShapeID>(shape), Length2{dx * 1_m, dy * 1_m}); return 0; REF_CATCH(-1) }
// one shape flattened like ref_export_shapes does all of them (child_offset 0, vertex offsets relative to verts); returns the vertex count
int ref_export_shape(void* h, uint32_t sid, pr_shape* o, pr_child* children, uint32_t capChildren, pr_vec2* verts, pr_vec2* normals, uint32_t capVerts)
{
    REF_TRY(h)
    const auto& shape = GetShape(rw->world, MakeId<ShapeID>(sid));
    o->friction = GetFriction(shape); o->restitution = GetRestitution(shape); o->density = GetDensity(shape);
    const auto f = GetFilter(shape);
    o->filter_category = f.categoryBits; o->filter_mask = f.maskBits; o->filter_group = f.groupIndex;
    o->is_sensor = IsSensor(shape) ? 1 : 0;
    o->child_offset = 0; o->child_count = GetChildCount(shape);
    uint32_t vi = 0;
    for (ChildCounter c = 0; c < o->child_count; ++c) {
        const auto dp = GetChild(shape, c);
        if (c < capChildren) { children[c].vertex_offset = vi; children[c].vertex_count = dp.GetVertexCount(); children[c].vertex_radius = dp.GetVertexRadius(); }
        for (VertexCounter v = 0; v < dp.GetVertexCount(); ++v) {
            if (vi < capVerts) {
                verts[vi] = pr_vec2{GetX(dp.GetVertex(v)), GetY(dp.GetVertex(v))};
                normals[vi] = dp.GetVertexCount() > 1 ? pr_vec2{dp.GetNormal(v).GetX(), dp.GetNormal(v).GetY()} : pr_vec2{0, 0};
            }
            ++vi;
        }
    }
    return static_cast<int>(vi);
    REF_CATCH(-1)
}

int ref_destroy_joint(void* h, uint32_t joint)
{
    REF_TRY(h)
    Destroy(rw->world, MakeId<JointID>(joint));
    return 0;
    REF_CATCH(-1)
}

// SetJoint through the reference's API: the current RevoluteJointConf with the row's motor / limit fields put in
int ref_set_revolute_joint(void* h, uint32_t joint, const pr_revolute_joint* row)
{
    REF_TRY(h)
    const auto id = MakeId<JointID>(joint);
    auto jc = TypeCast<RevoluteJointConf>(GetJoint(rw->world, id));
    // the whole value as the row says: bodies, collideConnected and anchors too (SetJoint = Remove + assign + Add, :1108-1130)
    jc.bodyA = MakeId<BodyID>(row->body_a); jc.bodyB = MakeId<BodyID>(row->body_b); jc.collideConnected = (row->collide_connected & 1u) != 0;
    jc.localAnchorA = Length2{row->local_anchor_a.x * 1_m, row->local_anchor_a.y * 1_m};
    jc.localAnchorB = Length2{row->local_anchor_b.x * 1_m, row->local_anchor_b.y * 1_m};
    jc.referenceAngle = row->reference_angle * 1_rad;
    jc.enableMotor = row->enable_motor != 0; jc.motorSpeed = row->motor_speed * 1_rad / 1_s; jc.maxMotorTorque = row->max_motor_torque * 1_Nm;
    jc.enableLimit = row->enable_limit != 0; jc.lowerAngle = row->lower_angle * 1_rad; jc.upperAngle = row->upper_angle * 1_rad;
    SetJoint(rw->world, id, Joint{jc});
    return 0;
    REF_CATCH(-1)
}

int ref_set_body_velocity(void* h, uint32_t body, float vx, float vy, float w)
{
    REF_TRY(h)
    SetVelocity(rw->world, MakeId<BodyID>(body), Velocity{LinearVelocity2{vx, vy}, w});
    return 0;
    REF_CATCH(-1)
}

int ref_set_body_transform(void* h, uint32_t body, float x, float y, float angle)
{
    REF_TRY(h)
    SetTransform(rw->world, MakeId<BodyID>(body), Length2{x, y}, angle);
    return 0;
    REF_CATCH(-1)
}

// Body::SetSweep(Sweep{Position{(x, y), angle}, localCenter}) through SetBody: pos0 = pos1 = the given position, the
// transformation recomputed from it (Body.hpp:697-702) -- how tests give the reference world another world's poses
int ref_set_body_sweep(void* h, uint32_t body, float x, float y, float angle)
{
    REF_TRY(h)
    auto b = GetBody(rw->world, MakeId<BodyID>(body));
    SetSweep(b, Sweep{Position{Length2{x * 1_m, y * 1_m}, angle * 1_rad}, GetLocalCenter(b)});
    SetBody(rw->world, MakeId<BodyID>(body), b);
    return 0;
    REF_CATCH(-1)
}

// Installs the reference's own listeners (World.hpp SetBeginContactListener / SetEndContactListener /
// SetPostSolveContactListener); each call is recorded with what the listener can see through GetContact at that moment.
int ref_set_event_mask(void* h, uint32_t mask)
{
    REF_TRY(h)
    auto record = [rw](uint32_t kind, ContactID id) {
        const auto c = GetContact(rw->world, id);
        rw->evContact.push_back(pr_contact_event{kind, static_cast<uint32_t>(to_underlying(id)),
            static_cast<uint32_t>(to_underlying(GetBodyA(c))), static_cast<uint32_t>(to_underlying(GetShapeA(c))), static_cast<uint32_t>(GetChildIndexA(c)),
            static_cast<uint32_t>(to_underlying(GetBodyB(c))), static_cast<uint32_t>(to_underlying(GetShapeB(c))), static_cast<uint32_t>(GetChildIndexB(c))});
    };
    if (mask & PR_EVENT_BEGIN) SetBeginContactListener(rw->world, [record](ContactID id) { record(PR_EVENT_BEGIN, id); });
    else SetBeginContactListener(rw->world, {});
    if (mask & PR_EVENT_END) SetEndContactListener(rw->world, [record](ContactID id) { record(PR_EVENT_END, id); });
    else SetEndContactListener(rw->world, {});
    if (mask & PR_EVENT_POST_SOLVE)
        SetPostSolveContactListener(rw->world, [rw](ContactID id, const ContactImpulsesList& imp, unsigned iters) {
            pr_impulse_event e;
            std::memset(&e, 0, sizeof(e));
            e.contact = static_cast<uint32_t>(to_underlying(id)); e.count = imp.GetCount(); e.iterations = iters; e.phase = 0xFFFFFFFFu; // (phase: not visible to a listener)
            for (unsigned j = 0; j < imp.GetCount() && j < 2; ++j) { e.normal[j] = imp.GetEntryNormal(j); e.tangent[j] = imp.GetEntryTanget(j); }
            rw->evImpulse.push_back(e);
        });
    else SetPostSolveContactListener(rw->world, {});
    if (mask & PR_EVENT_PRE_SOLVE)
        SetPreSolveContactListener(rw->world, [rw](ContactID id, const Manifold& oldManifold) {
            pr_presolve_event e;
            std::memset(&e, 0, sizeof(e));
            e.contact = static_cast<uint32_t>(to_underlying(id));
            FromManifold(oldManifold, e.old_manifold);
            rw->evPre.push_back(e);
        });
    else SetPreSolveContactListener(rw->world, {});
    return 0;
    REF_CATCH(-1)
}

int ref_events(void* h, uint32_t kind, void* buf, uint32_t cap, uint32_t* n_out)
{
    REF_TRY(h)
    if (kind == PR_EVENT_PRE_SOLVE) {
        *n_out = static_cast<uint32_t>(rw->evPre.size());
        std::memcpy(buf, rw->evPre.data(), sizeof(pr_presolve_event) * std::min<size_t>(cap, rw->evPre.size()));
    } else if (kind == PR_EVENT_POST_SOLVE) {
        *n_out = static_cast<uint32_t>(rw->evImpulse.size());
        std::memcpy(buf, rw->evImpulse.data(), sizeof(pr_impulse_event) * std::min<size_t>(cap, rw->evImpulse.size()));
    } else {
        std::vector<pr_contact_event> ev;
        for (const auto& e : rw->evContact) if (e.kind & kind) ev.push_back(e);
        *n_out = static_cast<uint32_t>(ev.size());
        std::memcpy(buf, ev.data(), sizeof(pr_contact_event) * std::min<size_t>(cap, ev.size()));
    }
    return 0;
    REF_CATCH(-1)
}

// RayCast(const World&, const RayCastInput&, ShapeRayCastCB) with a callback that records every call and returns ResetRay
// (RayCastOutput.cpp:301-337); hits in callback order
int ref_ray_cast(void* h, uint32_t n_rays, const float* rays5, uint32_t cap, pr_ray_hit* hits, uint32_t* n_out)
{
    REF_TRY(h)
    uint32_t n = 0;
    for (uint32_t r = 0; r < n_rays; ++r) {
        const RayCastInput input{Length2{rays5[5 * r] * 1_m, rays5[5 * r + 1] * 1_m}, Length2{rays5[5 * r + 2] * 1_m, rays5[5 * r + 3] * 1_m}, rays5[5 * r + 4]};
        RayCast(rw->world, input, [&](BodyID b, ShapeID s, ChildCounter c, const Length2& point, const UnitVec& normal) {
            if (n < cap) {
                // the fraction is not handed to the callback: recompute the shape cast the way RayCast(world) did
                const auto shape = GetShape(rw->world, s);
                const auto out = RayCast(GetChild(shape, c), input, GetTransformation(rw->world, b));
                hits[n] = pr_ray_hit{r, static_cast<uint32_t>(to_underlying(b)), static_cast<uint32_t>(to_underlying(s)), static_cast<uint32_t>(c),
                                     out.has_value() ? static_cast<float>(out->fraction) : -1.0f, static_cast<float>(normal.GetX()), static_cast<float>(normal.GetY()),
                                     static_cast<float>(Real(GetX(point) / Meter)), static_cast<float>(Real(GetY(point) / Meter))};
            }
            ++n;
            return RayCastOpcode::ResetRay;
        });
    }
    *n_out = n;
    return 0;
    REF_CATCH(-1)
}

// TestPoint(const World&, BodyID, ShapeID, const Length2&), WorldShape.cpp:178; FindClosestBody(const World&, const Length2&), WorldBody.cpp:576
int ref_test_points(void* h, uint32_t n, const uint32_t* body, const uint32_t* shape, const float* xy, uint32_t* out)
{
    REF_TRY(h)
    for (uint32_t i = 0; i < n; ++i)
        out[i] = TestPoint(rw->world, MakeId<BodyID>(body[i]), MakeId<ShapeID>(shape[i]), Length2{xy[2 * i] * 1_m, xy[2 * i + 1] * 1_m}) ? 1u : 0u;
    return 0;
    REF_CATCH(-1)
}
int ref_find_closest_bodies(void* h, uint32_t n, const float* xy, uint32_t* out)
{
    REF_TRY(h)
    for (uint32_t i = 0; i < n; ++i) {
        const auto id = FindClosestBody(rw->world, Length2{xy[2 * i] * 1_m, xy[2 * i + 1] * 1_m});
        out[i] = (id == InvalidBodyID) ? 0xFFFFFFFFu : static_cast<uint32_t>(to_underlying(id));
    }
    return 0;
    REF_CATCH(-1)
}

int ref_shift_origin(void* h, float x, float y)
{
    REF_TRY(h)
    ShiftOrigin(rw->world, Length2{x, y});
    return 0;
    REF_CATCH(-1)
}

int ref_step(void* h, const pr_step_conf* conf, pr_step_stats* stats)
{
    REF_TRY(h)
    rw->evContact.clear(); rw->evImpulse.clear(); rw->evPre.clear();
    const auto s = Step(rw->world, ToStepConf(*conf));
    if (stats) FromStats(s, *stats);
    return 0;
    REF_CATCH(-1)
}

// steps n times; returns wall seconds spent inside Step() (cpu_baseline)
double ref_step_timed(void* h, const pr_step_conf* conf, uint32_t n, pr_step_stats* last)
{
    auto* rw = static_cast<RefWorld*>(h);
    const auto sc = ToStepConf(*conf);
    StepStats s;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < n; ++i) s = Step(rw->world, sc);
    const auto t1 = std::chrono::steady_clock::now();
    if (last) FromStats(s, *last);
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_get_counts(void* h, pr_counts* out)
{
    REF_TRY(h)
    std::memset(out, 0, sizeof(*out));
    out->bodies = static_cast<uint32_t>(GetBodyRange(rw->world));
    out->shapes = static_cast<uint32_t>(GetShapeRange(rw->world));
    out->joints = static_cast<uint32_t>(GetJointRange(rw->world));
    out->contacts = static_cast<uint32_t>(GetContacts(rw->world).size());
    out->contact_id_range = static_cast<uint32_t>(GetContactRange(rw->world));
    const auto& tree = GetTree(rw->world);
    out->proxies = tree.GetLeafCount();
    out->tree_leaf_count = tree.GetLeafCount();
    out->tree_node_capacity = tree.GetNodeCapacity();
    uint32_t nchild = 0, nvert = 0;
    for (uint32_t s = 0; s < out->shapes; ++s) {
        const auto& shape = GetShape(rw->world, MakeId<ShapeID>(s));
        const auto cc = GetChildCount(shape);
        nchild += cc;
        for (ChildCounter c = 0; c < cc; ++c) nvert += GetChild(shape, c).GetVertexCount();
    }
    out->children = nchild;
    out->vertices = nvert;
    return 0;
    REF_CATCH(-1)
}

// Flatten every shape: SURVEY F7 (each child is a DistanceProxy{radius,count,vertices,normals})
int ref_export_shapes(void* h, pr_shape* shapes, pr_child* children, pr_vec2* verts, pr_vec2* normals)
{
    REF_TRY(h)
    const auto ns = static_cast<uint32_t>(GetShapeRange(rw->world));
    uint32_t ci = 0, vi = 0;
    for (uint32_t s = 0; s < ns; ++s) {
        const auto& shape = GetShape(rw->world, MakeId<ShapeID>(s));
        auto& o = shapes[s];
        o.friction = GetFriction(shape);
        o.restitution = GetRestitution(shape);
        o.density = GetDensity(shape);
        const auto f = GetFilter(shape);
        o.filter_category = f.categoryBits;
        o.filter_mask = f.maskBits;
        o.filter_group = f.groupIndex;
        o.is_sensor = IsSensor(shape) ? 1 : 0;
        o.child_offset = ci;
        o.child_count = GetChildCount(shape);
        for (ChildCounter c = 0; c < o.child_count; ++c) {
            const auto dp = GetChild(shape, c);
            children[ci].vertex_offset = vi;
            children[ci].vertex_count = dp.GetVertexCount();
            children[ci].vertex_radius = dp.GetVertexRadius();
            for (VertexCounter v = 0; v < dp.GetVertexCount(); ++v) {
                verts[vi] = pr_vec2{GetX(dp.GetVertex(v)), GetY(dp.GetVertex(v))};
                if (dp.GetVertexCount() > 1) {
                    normals[vi] = pr_vec2{dp.GetNormal(v).GetX(), dp.GetNormal(v).GetY()};
                } else {
                    normals[vi] = pr_vec2{0, 0};
                }
                ++vi;
            }
            ++ci;
        }
    }
    return 0;
    REF_CATCH(-1)
}

int ref_export_bodies(void* h, pr_body* bodies)
{
    REF_TRY(h)
    const auto nb = static_cast<uint32_t>(GetBodyRange(rw->world));
    for (uint32_t b = 0; b < nb; ++b) FromBody(GetBody(rw->world, MakeId<BodyID>(b)), bodies[b]);
    return 0;
    REF_CATCH(-1)
}

// CSR of shape ids per body; offsets has nb+1 entries
int ref_export_body_shapes(void* h, uint32_t* offsets, uint32_t* shape_ids, uint32_t cap)
{
    REF_TRY(h)
    const auto nb = static_cast<uint32_t>(GetBodyRange(rw->world));
    uint32_t k = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        offsets[b] = k;
        for (const auto& s : GetShapes(rw->world, MakeId<BodyID>(b))) {
            if (k < cap) shape_ids[k] = to_underlying(s);
            ++k;
        }
    }
    offsets[nb] = k;
    return static_cast<int>(k);
    REF_CATCH(-1)
}

int ref_export_proxies(void* h, pr_proxy* out, uint32_t cap)
{
    REF_TRY(h)
    const auto& tree = GetTree(rw->world);
    uint32_t n = 0;
    for (DynamicTree::Size i = 0; i < tree.GetNodeCapacity(); ++i) {
        if (DynamicTree::IsLeaf(tree.GetHeight(i))) {
            if (n < cap) {
                const auto leaf = tree.GetLeafData(i);
                const auto aabb = tree.GetAABB(i);
                out[n].tree_id = i;
                out[n].body = to_underlying(leaf.bodyId);
                out[n].shape = to_underlying(leaf.shapeId);
                out[n].child = leaf.childId;
                out[n].min_x = aabb.ranges[0].GetMin();
                out[n].max_x = aabb.ranges[0].GetMax();
                out[n].min_y = aabb.ranges[1].GetMin();
                out[n].max_y = aabb.ranges[1].GetMax();
            }
            ++n;
        }
    }
    return static_cast<int>(n);
    REF_CATCH(-1)
}

// Query(const DynamicTree&, const AABB&, QueryShapeCallback) (DynamicTree.hpp:708): the reference's own spatial query over its tree
int ref_query_aabb(void* h, float min_x, float min_y, float max_x, float max_y, uint32_t* body, uint32_t* shape, uint32_t* child, uint32_t cap)
{
    REF_TRY(h)
    const AABB aabb{LengthInterval{min_x * 1_m, max_x * 1_m}, LengthInterval{min_y * 1_m, max_y * 1_m}};
    uint32_t n = 0;
    Query(GetTree(rw->world), aabb, [&](BodyID b, ShapeID s, ChildCounter c) {
        if (n < cap) { body[n] = to_underlying(b); shape[n] = to_underlying(s); child[n] = c; }
        ++n;
        return true;
    });
    return static_cast<int>(n);
    REF_CATCH(-1)
}

int ref_export_contacts(void* h, pr_contact* out, pr_manifold* manifolds, uint32_t cap)
{
    REF_TRY(h)
    const auto contacts = GetContacts(rw->world);
    uint32_t n = 0;
    for (const auto& kc : contacts) {
        if (n < cap) {
            const auto key = std::get<ContactKey>(kc);
            const auto id = std::get<ContactID>(kc);
            const auto& c = GetContact(rw->world, id);
            auto& o = out[n];
            o.id = to_underlying(id);
            o.key_min = key.GetMin();
            o.key_max = key.GetMax();
            o.body_a = to_underlying(GetBodyA(c));
            o.shape_a = to_underlying(GetShapeA(c));
            o.child_a = GetChildIndexA(c);
            o.body_b = to_underlying(GetBodyB(c));
            o.shape_b = to_underlying(GetShapeB(c));
            o.child_b = GetChildIndexB(c);
            o.friction = GetFriction(c);
            o.restitution = GetRestitution(c);
            o.tangent_speed = GetTangentSpeed(c);
            o.toi = HasValidToi(c) ? static_cast<float>(*GetToi(c)) : 1.0f;
            o.toi_count = GetToiCount(c);
            uint32_t f = 0;
            if (IsTouching(c)) f |= PR_CONTACT_TOUCHING;
            if (IsEnabled(c)) f |= PR_CONTACT_ENABLED;
            if (NeedsFiltering(c)) f |= PR_CONTACT_FILTER;
            if (HasValidToi(c)) f |= PR_CONTACT_TOI;
            if (NeedsUpdating(c)) f |= PR_CONTACT_DIRTY;
            if (IsSensor(c)) f |= PR_CONTACT_SENSOR;
            if (IsDestroyed(c)) f |= PR_CONTACT_DESTROYED;
            if (IsImpenetrable(c)) f |= PR_CONTACT_IMPENETRABLE;
            o.flags = f;
            if (manifolds) FromManifold(GetManifold(rw->world, id), manifolds[n]);
        }
        ++n;
    }
    return static_cast<int>(n);
    REF_CATCH(-1)
}

// SetContact through the reference's API: the current value with the row's mutable fields put in (friction, restitution,
// tangent speed, enabled / needs-updating flags). Returns 0, or -1 with the reference's exception text in ref_last_error.
int ref_set_contact(void* h, const pr_contact* row)
{
    REF_TRY(h)
    const auto id = MakeId<ContactID>(row->id);
    auto c = GetContact(rw->world, id);
    SetFriction(c, row->friction);
    SetRestitution(c, row->restitution);
    SetTangentSpeed(c, row->tangent_speed * 1_mps);
    if (row->flags & PR_CONTACT_ENABLED) SetEnabled(c); else UnsetEnabled(c);
    if (row->flags & PR_CONTACT_DIRTY) c.FlagForUpdating(); else c.UnflagForUpdating();
    SetContact(rw->world, id, c);
    return 0;
    REF_CATCH(-1)
}
// SetManifold: the current manifold with the row's impulses (the one thing the reference lets a user change)
int ref_set_manifold_impulses(void* h, uint32_t contact, const pr_manifold* row)
{
    REF_TRY(h)
    const auto id = MakeId<ContactID>(contact);
    auto m = GetManifold(rw->world, id);
    if (m.GetPointCount() > 0) m.SetImpulses(0, row->jn0 * NewtonSecond, row->jt0 * NewtonSecond);
    if (m.GetPointCount() > 1) m.SetImpulses(1, row->jn1 * NewtonSecond, row->jt1 * NewtonSecond);
    SetManifold(rw->world, id, m);
    return 0;
    REF_CATCH(-1)
}

float ref_get_inv_dt0(void* h) { return GetInvDeltaTime(static_cast<RefWorld*>(h)->world); }

// ---- stage functions of the reference on flat arrays ----

// CollideShapes (Manifold.cpp:393) for n pairs of children
int ref_collide_shapes(uint32_t n, const pr_child* children, const pr_vec2* verts,
                       const pr_vec2* normals, const uint32_t* a_child, const float* xf_a,
                       const uint32_t* b_child, const float* xf_b, float linearSlop,
                       float maxCirclesRatio, pr_manifold* out)
{
    static_assert(sizeof(Length2) == sizeof(pr_vec2), "layout");
    static_assert(sizeof(UnitVec) == sizeof(pr_vec2), "layout");
    auto conf = Manifold::Conf{};
    conf.linearSlop = linearSlop;
    conf.maxCirclesRatio = maxCirclesRatio;
    for (uint32_t i = 0; i < n; ++i) {
        const auto& ca = children[a_child[i]];
        const auto& cb = children[b_child[i]];
        const auto pa = DistanceProxy{ca.vertex_radius, static_cast<VertexCounter>(ca.vertex_count),
            reinterpret_cast<const Length2*>(verts + ca.vertex_offset),
            reinterpret_cast<const UnitVec*>(normals + ca.vertex_offset)};
        const auto pb = DistanceProxy{cb.vertex_radius, static_cast<VertexCounter>(cb.vertex_count),
            reinterpret_cast<const Length2*>(verts + cb.vertex_offset),
            reinterpret_cast<const UnitVec*>(normals + cb.vertex_offset)};
        const float* a = xf_a + 4 * i;
        const float* b = xf_b + 4 * i;
        const auto qa = UnitVec::Get(a[2], a[3]).first; // exact when (c,s) already unit: see below
        const auto qb = UnitVec::Get(b[2], b[3]).first;
        (void)qa; (void)qb;
        // Build the transformation from raw (c,s) bits without renormalising:
        Transformation xa, xb;
        xa.p = Length2{a[0], a[1]};
        xb.p = Length2{b[0], b[1]};
        std::memcpy(static_cast<void*>(&xa.q), a + 2, sizeof(float) * 2);
        std::memcpy(static_cast<void*>(&xb.q), b + 2, sizeof(float) * 2);
        FromManifold(CollideShapes(pa, xa, pb, xb, conf), out[i]);
    }
    return 0;
}

// UnitVec::Get(angle) == {cos, sin} as the reference computes it (UnitVec.cpp:27-30)
void ref_unitvec_from_angle(uint32_t n, const float* angles, float* cs)
{
    for (uint32_t i = 0; i < n; ++i) {
        const auto u = UnitVec::Get(Angle{angles[i]});
        cs[2 * i] = u.GetX();
        cs[2 * i + 1] = u.GetY();
    }
}

// GetToiViaSat (TimeOfImpact.cpp:35) on flat arrays: sweeps as 8 floats
// (pos0 x,y,a, pos1 x,y,a, lc x,y); alpha0 = 0. out: time, state, toi_iters, max_dist, max_root
int ref_toi(uint32_t n, const pr_child* children, const pr_vec2* verts, const pr_vec2* normals,
            const uint32_t* a_child, const float* sweep_a, const uint32_t* b_child,
            const float* sweep_b, const pr_step_conf* sc, float* out_time, int32_t* out_state,
            uint32_t* out_iters)
{
    const auto conf = GetToiConf(ToStepConf(*sc));
    for (uint32_t i = 0; i < n; ++i) {
        const auto& ca = children[a_child[i]];
        const auto& cb = children[b_child[i]];
        const auto pa = DistanceProxy{ca.vertex_radius, static_cast<VertexCounter>(ca.vertex_count),
            reinterpret_cast<const Length2*>(verts + ca.vertex_offset),
            reinterpret_cast<const UnitVec*>(normals + ca.vertex_offset)};
        const auto pb = DistanceProxy{cb.vertex_radius, static_cast<VertexCounter>(cb.vertex_count),
            reinterpret_cast<const Length2*>(verts + cb.vertex_offset),
            reinterpret_cast<const UnitVec*>(normals + cb.vertex_offset)};
        const float* a = sweep_a + 8 * i;
        const float* b = sweep_b + 8 * i;
        const auto sa = Sweep{Position{Length2{a[0], a[1]}, a[2]}, Position{Length2{a[3], a[4]}, a[5]},
                              Length2{a[6], a[7]}};
        const auto sb = Sweep{Position{Length2{b[0], b[1]}, b[2]}, Position{Length2{b[3], b[4]}, b[5]},
                              Length2{b[6], b[7]}};
        const auto o = GetToiViaSat(pa, sa, pb, sb, conf);
        out_time[i] = o.time;
        out_state[i] = static_cast<int32_t>(o.state);
        out_iters[3 * i + 0] = o.stats.toi_iters;
        out_iters[3 * i + 1] = o.stats.max_dist_iters;
        out_iters[3 * i + 2] = o.stats.max_root_iters;
    }
    return 0;
}

} // extern "C"
