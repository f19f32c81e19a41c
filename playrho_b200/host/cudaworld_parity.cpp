// cudaworld_parity.cpp -- drives World{AabbTreeWorld} and World{CudaWorld} side by side through the
// reference's own public API (CreateShape / CreateBody / Step / GetBody / GetContacts / GetManifold) and
// reports the differences. Used by tests/test_cudaworld_host.py on the GPU box (TEST DRIVER, links oracle/_ref).
#include "CudaWorld.hpp"
#include <playrho/d2/AabbTreeWorld.hpp>
#include <playrho/d2/BasicAPI.hpp>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>

using namespace playrho;
using namespace playrho::d2;

static void BuildPyramid(World& world, int rows)
{ // Testbed/Tests/Pyramid.cpp recipe
    const auto ground = CreateShape(world, EdgeShapeConf{Length2{-60_m, 0_m}, Length2{60_m, 0_m}});
    CreateBody(world, BodyConf{}.Use(ground));
    const auto box = CreateShape(world, PolygonShapeConf{}.UseDensity(5_kgpm2).SetAsBox(0.5_m, 0.5_m));
    auto x = Length2{-7_m, 0.75_m};
    const auto dx = Length2{0.5625_m, 1.25_m}, dy = Length2{1.125_m, 0_m};
    for (int i = 0; i < rows; ++i) {
        auto y = x;
        for (int j = i; j < rows; ++j) {
            CreateBody(world, BodyConf{}.Use(BodyType::Dynamic).UseLocation(y).UseLinearAcceleration(EarthlyGravity).Use(box));
            y += dy;
        }
        x += dx;
    }
}

static void BuildDisks(World& world, int n)
{ // closed edge box + n disks (some fast, some bullets) : exercises TOI
    const Real W = 12;
    const ShapeID edges[4] = {CreateShape(world, EdgeShapeConf{Length2{0_m, 0_m}, Length2{W * 1_m, 0_m}}),
                              CreateShape(world, EdgeShapeConf{Length2{W * 1_m, 0_m}, Length2{W * 1_m, W * 1_m}}),
                              CreateShape(world, EdgeShapeConf{Length2{W * 1_m, W * 1_m}, Length2{0_m, W * 1_m}}),
                              CreateShape(world, EdgeShapeConf{Length2{0_m, W * 1_m}, Length2{0_m, 0_m}})};
    auto gconf = BodyConf{};
    for (auto e : edges) gconf.Use(e);
    CreateBody(world, gconf);
    const auto disk = CreateShape(world, DiskShapeConf{}.UseRadius(0.3_m).UseDensity(1_kgpm2).UseRestitution(Real(0.3)));
    unsigned seed = 12345;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return Real((seed >> 8) & 0xFFFF) / Real(65535); };
    for (int k = 0; k < n; ++k) {
        const auto px = Real(1) + Real(k % 10), py = Real(1) + Real(k / 10);
        const auto sp = Real(5) + Real(40) * rnd(), th = Real(6.28318) * rnd();
        CreateBody(world, BodyConf{}.Use(BodyType::Dynamic).UseLocation(Length2{px * 1_m, py * 1_m})
                              .UseLinearVelocity(LinearVelocity2{sp * std::cos(th) * 1_mps, sp * std::sin(th) * 1_mps})
                              .UseLinearAcceleration(EarthlyGravity).UseBullet(k % 3 == 0).Use(disk));
    }
}

int main(int argc, char** argv)
{
    const char* scene = argc > 1 ? argv[1] : "pyramid";
    const int steps = argc > 2 ? std::atoi(argv[2]) : 120;
    World ref{AabbTreeWorld{}};
    auto dev = b200::CudaWorld{};
    dev.SetToiMode(b200::CudaWorld::ToiMode::SerialExact);
    World cud{dev};
    for (World* w : {&ref, &cud}) {
        if (!std::strcmp(scene, "pyramid")) BuildPyramid(*w, 8); else BuildDisks(*w, 40);
    }
    // the reference's listener API on both worlds: begin/end/post-solve calls are counted; a begin listener also reads the
    // contact through the world (GetContact must work from inside a listener)
    unsigned long nBegin[2] = {0, 0}, nEnd[2] = {0, 0}, nPost[2] = {0, 0}, badListenerRead = 0;
    double postImpulse[2] = {0, 0};
    int wi = 0;
    for (World* w : {&ref, &cud}) {
        const int k = wi++;
        SetBeginContactListener(*w, [&, w, k](ContactID id) { ++nBegin[k]; if (!IsTouching(GetContact(*w, id)) && k == 0) ++badListenerRead; });
        SetEndContactListener(*w, [&, k](ContactID) { ++nEnd[k]; });
        SetPostSolveContactListener(*w, [&, k](ContactID, const ContactImpulsesList& imp, unsigned) {
            ++nPost[k];
            for (unsigned j = 0; j < imp.GetCount(); ++j) postImpulse[k] += static_cast<double>(Real(imp.GetEntryNormal(j) / NewtonSecond));
        });
    }
    unsigned long setCalls = 0, setRefused = 0;
    const bool shift = argc > 3 && !std::strcmp(argv[3], "shift");
    const auto conf = StepConf{};
    double maxPos = 0, maxVel = 0;
    unsigned long touchMismatch = 0, contactCountMismatch = 0, toiEvents = 0, toiEventsRef = 0;
    for (int s = 0; s < steps; ++s) {
        if (shift && s == steps / 3) { // SetContact / SetManifold through the World API on both worlds, and the refusals
            for (World* w : {&ref, &cud}) {
                for (const auto& kc : GetContacts(*w)) {
                    const auto id = std::get<ContactID>(kc);
                    auto c = GetContact(*w, id);
                    SetFriction(c, GetFriction(c)); // (value unchanged: this test reports drift; the effect of edits is gated bit for bit in tests/test_set_contact.py)
                    SetContact(*w, id, c);
                    SetManifold(*w, id, GetManifold(*w, id));
                    ++setCalls;
                    auto bad = c;
                    bad.SetToiCount(static_cast<Contact::substep_type>(GetToiCount(c) + 1));
                    try { SetContact(*w, id, bad); } catch (const InvalidArgument&) { ++setRefused; }
                }
                try { SetContact(*w, ContactID(60000), Contact{}); } catch (const std::out_of_range&) { ++setRefused; } catch (const InvalidArgument&) { ++setRefused; }
            }
        }
        if (shift && s == steps / 2) { ShiftOrigin(ref, Length2{13.37_m, -2.25_m}); ShiftOrigin(cud, Length2{13.37_m, -2.25_m}); }
        const auto rs = Step(ref, conf);
        const auto cs = Step(cud, conf);
        toiEvents += cs.toi.contactsFound; toiEventsRef += rs.toi.contactsFound;
        const auto rc = GetContacts(ref), cc = GetContacts(cud);
        if (rc.size() != cc.size()) ++contactCountMismatch;
        std::set<std::pair<unsigned, unsigned>> rt, ct;
        for (const auto& kc : rc) if (IsTouching(GetContact(ref, std::get<ContactID>(kc)))) rt.insert({std::get<ContactKey>(kc).GetMin(), std::get<ContactKey>(kc).GetMax()});
        for (const auto& kc : cc) if (IsTouching(GetContact(cud, std::get<ContactID>(kc)))) ct.insert({std::get<ContactKey>(kc).GetMin(), std::get<ContactKey>(kc).GetMax()});
        if (rt != ct) ++touchMismatch;
    }
    const auto nb = GetBodyRange(ref);
    for (BodyCounter i = 0; i < nb; ++i) {
        const auto id = BodyID(i);
        const auto a = GetBody(ref, id), b = GetBody(cud, id);
        const auto dp = GetLocation(a) - GetLocation(b);
        maxPos = std::max(maxPos, static_cast<double>(std::hypot(Real(GetX(dp)), Real(GetY(dp)))));
        const auto dv = GetVelocity(a).linear - GetVelocity(b).linear;
        maxVel = std::max(maxVel, static_cast<double>(std::hypot(Real(GetX(dv)), Real(GetY(dv)))));
    }
    // manifold accessors work unchanged on the new type
    unsigned long points = 0;
    for (const auto& kc : GetContacts(cud)) points += GetManifold(cud, std::get<ContactID>(kc)).GetPointCount();
    std::printf("scene=%s steps=%d bodies=%u contacts_ref=%zu contacts_dev=%zu manifold_points_dev=%lu toi_events_ref=%lu toi_events_dev=%lu "
                "steps_with_contact_count_mismatch=%lu steps_with_touching_set_mismatch=%lu max_pos_diff=%.6g max_vel_diff=%.6g type_is_cudaworld=%d "
                "begin_ref=%lu begin_dev=%lu end_ref=%lu end_dev=%lu post_ref=%lu post_dev=%lu post_impulse_ref=%.6g post_impulse_dev=%.6g shifted=%d set_contact_calls=%lu set_contact_refusals=%lu\n",
                scene, steps, static_cast<unsigned>(nb), GetContacts(ref).size(), GetContacts(cud).size(), points, toiEventsRef, toiEvents,
                contactCountMismatch, touchMismatch, maxPos, maxVel, TypeCast<b200::CudaWorld>(&cud) != nullptr ? 1 : 0,
                nBegin[0], nBegin[1], nEnd[0], nEnd[1], nPost[0], nPost[1], postImpulse[0], postImpulse[1], shift ? 1 : 0, setCalls, setRefused);
    (void)badListenerRead;
    return 0;
}
