// orc_toi.h -- TEST INFRASTRUCTURE ONLY (oracle restatement): GJK distance, TOI and revolute joint.
#ifndef ORC_TOI_H
#define ORC_TOI_H
#include "orc_math.h"
namespace orc {

// TestOverlap(proxyA, xfA, proxyB, xfB, conf), Distance.cpp:205-230: returns squared-area-like
// signed value; >= 0 means overlapping.
float test_overlap(const Proxy& a, Xf xfA, const Proxy& b, Xf xfB, uint32_t maxIters);

struct ToiOut { float time; int state; uint32_t toi_iters, max_dist_iters, max_root_iters; };
struct Sweep { Pos pos0, pos1; V2 lc; float alpha0; };
// GetToiViaSat, TimeOfImpact.cpp:35-251
ToiOut toi_via_sat(const Proxy& a, const Sweep& sA, const Proxy& b, const Sweep& sB, const pr_step_conf& conf);

struct JointSolverConf { float dt, dtRatio; bool warmStart; float linearSlop, angularSlop, maxLinearCorrection, maxAngularCorrection; float resolutionRate = 0.2f; };
struct RevoluteWork { V2 rA{0, 0}, rB{0, 0}; float mass[9] = {0}; float angularMass = 0; float ex[8] = {0}; };
void revolute_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                            float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c);
bool revolute_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& velA,
                             float imB, float iiB, Vel& velB, const JointSolverConf& c);
bool revolute_solve_position(const pr_revolute_joint& j, float imA, float iiA, V2 lcA, Pos& posA,
                             float imB, float iiB, V2 lcB, Pos& posB, const JointSolverConf& c);
// the joint table's rows carry their type (PR_JOINT_TYPE): these dispatch to the revolute or the distance joint
// (DistanceJointConf.cpp:81-246; work: rA, rB, mass[0..1] = u, mass[2] = invGamma, mass[3] = bias, mass[4] = mass)
void joint_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                         float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c);
bool joint_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& velA,
                          float imB, float iiB, Vel& velB, const JointSolverConf& c);
bool joint_solve_position(const pr_revolute_joint& j, const RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos& posA,
                          float imB, float iiB, V2 lcB, Pos& posB, const JointSolverConf& c);
// GearJointConf.cpp:171-468: four bodies (A, B = the row's; C, D and the copies of the geared joints' data in the pr_gear_joint value)
struct GearBody { float invM, invI; V2 lc; Pos* pos; Vel* vel; };
void gear_init_velocity(pr_revolute_joint& j, const pr_gear_joint& g, RevoluteWork& wk, GearBody A, GearBody B, GearBody C, GearBody D, const JointSolverConf& c);
bool gear_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, GearBody A, GearBody B, GearBody C, GearBody D);
bool gear_solve_position(const pr_revolute_joint& j, const pr_gear_joint& g, GearBody A, GearBody B, GearBody C, GearBody D, const JointSolverConf& c);
} // namespace orc
#endif
