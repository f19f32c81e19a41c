// pr_joint.cuh -- joints on the device: the revolute joint and the distance joint (below):
// InitVelocity / SolveVelocity / SolvePosition
// of Library/source/playrho/d2/RevoluteJointConf.cpp:127-435 (GetMat33 :35-62, Solve33 / Solve22 Math.hpp:447-474),
// one thread per joint. Joints run before the contact colours in every velocity iteration and after them in
// every position iteration, exactly where SolveRegIslandViaGS calls them (AabbTreeWorld.cpp:1597-1644).
// When two joints share a speedable body the host launches these kernels with ONE thread that walks the joints
// in id order (a valid Gauss-Seidel order); otherwise one thread per joint.
#ifndef PR_JOINT_CUH
#define PR_JOINT_CUH

namespace pr {

struct V3 { float x, y, z; };
PR_D float dot3(V3 a, V3 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; }
PR_D V3 cross3(V3 a, V3 b) { V3 r; r.x = a.y * b.z - a.z * b.y; r.y = a.z * b.x - a.x * b.z; r.z = a.x * b.y - a.y * b.x; return r; }
PR_D V3 mcol(const float* m, int c) { V3 r; r.x = m[3 * c]; r.y = m[3 * c + 1]; r.z = m[3 * c + 2]; return r; }
PR_D V3 solve33(const float* m, V3 b)
{
    const float dp = dot3(mcol(m, 0), cross3(mcol(m, 1), mcol(m, 2)));
    const float det = (dp != 0) ? (1 / dp) : dp;
    V3 r;
    r.x = det * dot3(b, cross3(mcol(m, 1), mcol(m, 2)));
    r.y = det * dot3(mcol(m, 0), cross3(b, mcol(m, 2)));
    r.z = det * dot3(mcol(m, 0), cross3(mcol(m, 1), b));
    return r;
}
PR_D float2 solve22(const float* m, float2 b)
{
    const float matXX = m[0], matXY = m[3], matYX = m[1], matYY = m[4];
    const float cp = matXX * matYY - matXY * matYX;
    const float det = (cp != 0) ? (1 / cp) : cp;
    return v2(det * (matYY * b.x - matXY * b.y), det * (matXX * b.y - matYX * b.x));
}
enum { LS_INACTIVE = 0, LS_AT_LOWER = 1, LS_AT_UPPER = 2, LS_EQUAL = 3 };

// the island a joint is solved with: the active island of either of its bodies (AddToIsland :1443-1466);
// INVALID when no active island reached it or the other body is disabled
PR_D unsigned joint_island(const DV& d, const pr_revolute_joint& j)
{
    if (j.body_a == INVALID) return INVALID; // destroyed joint
    const unsigned ia = d.b_island[j.body_a], ib = d.b_island[j.body_b];
    if (ia != INVALID) return f_enabled(d.b_flags[j.body_b]) ? ia : INVALID;
    if (ib != INVALID) return f_enabled(d.b_flags[j.body_a]) ? ib : INVALID;
    return INVALID;
}

PR_D unsigned joint_type(const pr_revolute_joint& j) { return (j.collide_connected >> 8) & 0xFFu; }
// UnitVec::Get(x, y) with the magnitude (UnitVec.hpp:176-232, default fallback = the zero vector)
PR_D Rot unit_vector_mag(float2 v, float& mag)
{
    const float x = v.x, y = v.y;
    const float ms = x * x + y * y;
    if (is_normal(ms)) {
        mag = sqrtf(ms);
        const float inv = 1.0f / mag;
        return mkrot(x * inv, y * inv);
    }
    mag = hypotf(x, y);
    if (!is_normal(mag)) { mag = 0.0f; return mkrot(0.0f, 0.0f); }
    return mkrot(x / mag, y / mag);
}

// Where a joint's BodyConstraints and its island's flags live: the world's HBM arrays indexed by body id / island label
// (jview_global), or a block's shared-memory arrays indexed by slot (pr_block_solve.cuh; slot[] maps a body id to its slot).
// slotV / slotP / slotI: body id -> index into s_vel / s_pos / isl_state (nullptr: the id itself). isl_jflag is always the
// world's array, indexed by island label.
struct JView { float4 *s_vel, *s_pos; const unsigned* isl_state; unsigned* isl_jflag; const unsigned *slotV, *slotP, *slotI; };
PR_D JView jview_global(const DV& d)
{
    JView v; v.s_vel = d.s_vel; v.s_pos = d.s_pos; v.isl_state = d.isl_state; v.isl_jflag = d.isl_jflag; v.slotV = v.slotP = v.slotI = nullptr;
    return v;
}
PR_D unsigned jv_idx(const unsigned* slot, unsigned id) { return slot ? slot[id] : id; }

// ---- DistanceJointConf.cpp:81-246. Row overlay (pr_distance_joint): impulse = impulse_x, length = reference_angle, frequency =
// lower_angle, dampingRatio = upper_angle; work: rA, rB, m[0..1] = u, m[2] = invGamma, m[3] = bias, m[4] = mass
PR_D void distance_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const Rot qA = rot_from_angle(pA.z), qB = rot_from_angle(pB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    const float2 delta = (v2(pB.x, pB.y) + rB) - (v2(pA.x, pA.y) + rA);
    float length;
    const Rot u = unit_vector_mag(delta, length);
    const float crAu = cross(rA, u) / 1.0f, crBu = cross(rB, u) / 1.0f;
    const float invRotMassA = iiA * square(crAu), invRotMassB = iiB * square(crBu);
    float invMass = imA + invRotMassA + imB + invRotMassB;
    float mass = (invMass != 0.0f) ? 1.0f / invMass : 0.0f;
    float invGamma = 0.0f, bias = 0.0f;
    const float frequency = j.lower_angle, dampingRatio = j.upper_angle, restLength = j.reference_angle;
    if (frequency > 0.0f) {
        const float C = length - restLength;
        const float omega = 2.0f * 3.14159265358979323846264338327950288f * frequency;
        const float dd = 2.0f * mass * dampingRatio * omega;
        const float k = mass * square(omega);
        const float h = conf.delta_time;
        const float gamma = h * (dd + h * k);
        invGamma = (gamma != 0.0f) ? 1.0f / gamma : 0.0f;
        bias = C * h * k * invGamma;
        invMass += invGamma;
        mass = (invMass != 0.0f) ? 1.0f / invMass : 0.0f;
    }
    wk.m[0] = u.c; wk.m[1] = u.s; wk.m[2] = invGamma; wk.m[3] = bias; wk.m[4] = mass;
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio;
        const float2 P = j.impulse_x * u;
        const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else j.impulse_x = 0.0f;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void distance_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const float2 rA = v2(wk.rAx, wk.rAy), rB = v2(wk.rBx, wk.rBy);
    const Rot u = mkrot(wk.m[0], wk.m[1]);
    const float2 vpA = v2(velA.x, velA.y) + rev_perp(rA) * (velA.z / 1.0f);
    const float2 vpB = v2(velB.x, velB.y) + rev_perp(rB) * (velB.z / 1.0f);
    const float Cdot = dot(u, vpB - vpA);
    const float impulse = -wk.m[4] * (Cdot + wk.m[3] + wk.m[2] * j.impulse_x);
    j.impulse_x += impulse;
    const float2 P = impulse * u;
    const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
    const float2 la = v2(velA.x, velA.y) - mA.x * P; velA.x = la.x; velA.y = la.y; velA.z -= mA.y * LA;
    const float2 lb = v2(velB.x, velB.y) + mB.x * P; velB.x = lb.x; velB.y = lb.y; velB.z += mB.y * LB;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    if (!(impulse == 0.0f)) v.isl_jflag[wk.island] = 1u; // (:190: solved when the impulse is zero)
}
PR_D void distance_solve_position(const DV& d, const pr_revolute_joint& j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    if (j.lower_angle > 0.0f) return; // no position correction for soft distance constraints (:203-206)
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    float4 posA = v.s_pos[ia], posB = v.s_pos[ib];
    const Rot qA = rot_from_angle(posA.z), qB = rot_from_angle(posB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    const float2 relLoc = (v2(posB.x, posB.y) + rB) - (v2(posA.x, posA.y) + rA);
    float length;
    const Rot u = unit_vector_mag(relLoc, length);
    const float deltaLength = length - j.reference_angle;
    const float C = stdclamp(deltaLength, -conf.max_linear_correction, conf.max_linear_correction);
    const float impulse = -wk.m[4] * C;
    const float2 P = impulse * u;
    const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
    const float2 la = v2(posA.x, posA.y) - mA.x * P; posA.x = la.x; posA.y = la.y; posA.z -= mA.y * LA;
    const float2 lb = v2(posB.x, posB.y) + mB.x * P; posB.x = lb.x; posB.y = lb.y; posB.z += mB.y * LB;
    v.s_pos[ia] = posA; v.s_pos[ib] = posB;
    if (!(fabsf(C) < conf.linear_slop)) v.isl_jflag[wk.island] = 1u;
}

// ---- RopeJointConf.cpp:58-214. Row overlay (pr_rope_joint): impulse = impulse_x, maxLength = reference_angle, limitState = limit_state;
// work: rA, rB, m[0..1] = u, m[4] = mass, m[5] = length
PR_D void rope_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const Rot qA = rot_from_angle(pA.z), qB = rot_from_angle(pB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    const float2 posDelta = (v2(pB.x, pB.y) + rB) - (v2(pA.x, pA.y) + rA);
    float length;
    const Rot uv = unit_vector_mag(posDelta, length);
    wk.m[5] = length;
    const float C = length - j.reference_angle;
    j.limit_state = (C > 0.0f) ? LS_AT_UPPER : LS_INACTIVE;
    if (!(length > conf.linear_slop)) { // (:96-104: the bodies' velocities are left as they are)
        wk.m[0] = 0.0f; wk.m[1] = 0.0f; wk.m[4] = 0.0f;
        j.impulse_x = 0.0f;
        d.jt[ji] = j; d.jt_work[ji] = wk;
        return;
    }
    const Rot u = uv;
    wk.m[0] = u.c; wk.m[1] = u.s;
    const float crA = cross(rA, u) / 1.0f, crB = cross(rB, u) / 1.0f;
    const float invRotMassA = iiA * square(crA), invRotMassB = iiB * square(crB);
    const float invMass = imA + imB + invRotMassA + invRotMassB;
    wk.m[4] = (invMass != 0.0f) ? 1.0f / invMass : 0.0f;
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio;
        const float2 P = j.impulse_x * u;
        const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else j.impulse_x = 0.0f;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void rope_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const float2 rA = v2(wk.rAx, wk.rAy), rB = v2(wk.rBx, wk.rBy);
    const Rot u = mkrot(wk.m[0], wk.m[1]);
    const float2 vpA = v2(velA.x, velA.y) + rev_perp(rA) * (velA.z / 1.0f);
    const float2 vpB = v2(velB.x, velB.y) + rev_perp(rB) * (velB.z / 1.0f);
    const float C = wk.m[5] - j.reference_angle;
    const float inv_h = (conf.delta_time != 0.0f) ? 1.0f / conf.delta_time : 0.0f;
    const float Cdot = dot(u, vpB - vpA) + ((C < 0.0f) ? inv_h * C : 0.0f);
    float localImpulse = -wk.m[4] * Cdot;
    const float oldImpulse = j.impulse_x;
    j.impulse_x = stdmin(0.0f, j.impulse_x + localImpulse);
    localImpulse = j.impulse_x - oldImpulse;
    const float2 P = localImpulse * u;
    const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
    const float2 la = v2(velA.x, velA.y) - mA.x * P; velA.x = la.x; velA.y = la.y; velA.z -= mA.y * LA;
    const float2 lb = v2(velB.x, velB.y) + mB.x * P; velB.x = lb.x; velB.y = lb.y; velB.z += mB.y * LB;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    if (!(localImpulse == 0.0f)) v.isl_jflag[wk.island] = 1u;
}
PR_D void rope_solve_position(const DV& d, const pr_revolute_joint& j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    float4 posA = v.s_pos[ia], posB = v.s_pos[ib];
    const Rot qA = rot_from_angle(posA.z), qB = rot_from_angle(posB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    const float2 posDelta = (v2(posB.x, posB.y) + rB) - (v2(posA.x, posA.y) + rA);
    float length;
    const Rot uv = unit_vector_mag(posDelta, length);
    const float C = stdclamp(length - j.reference_angle, 0.0f, conf.max_linear_correction);
    const float localImpulse = -wk.m[4] * C;
    const float2 P = localImpulse * uv;
    const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
    const float2 la = v2(posA.x, posA.y) - mA.x * P; posA.x = la.x; posA.y = la.y; posA.z -= mA.y * LA;
    const float2 lb = v2(posB.x, posB.y) + mB.x * P; posB.x = lb.x; posB.y = lb.y; posB.z += mB.y * LB;
    v.s_pos[ia] = posA; v.s_pos[ib] = posB;
    if (!((length - j.reference_angle) < conf.linear_slop)) v.isl_jflag[wk.island] = 1u;
}

// ---- FrictionJointConf.cpp:76-239 and MotorJointConf.cpp:77-235 (one code path: the motor joint is the friction joint with a
// pose error fed back). Row overlay: linearImpulse = impulse_x/y, angularImpulse = impulse_z, maxForce = motor_speed, maxTorque =
// max_motor_torque; motor: linearOffset = local_anchor_a, angularOffset = reference_angle, correctionFactor = lower_angle.
// work: rA, rB, m[0..3] = linearMass (rows), angularMass; motor: m[4..5] = linearError, m[6] = angularError
PR_D float2 mat22_transform(const float* m, float2 v)
{ // Transform(v, M) = M * v, Vector.hpp:458-472 (each element accumulated from zero)
    float x = 0.0f; x += m[0] * v.x; x += m[1] * v.y;
    float y = 0.0f; y += m[2] * v.x; y += m[3] * v.y;
    return v2(x, y);
}
PR_D void fm_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v, bool motor)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const Rot qA = rot_from_angle(pA.z), qB = rot_from_angle(pB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = motor ? rotate(-d.b_lc[j.body_b], qB) : rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    float exx, exy, eyy;
    if (motor) {
        exx = imA + imB + iiA * square(rA.y) / 1.0f + iiB * square(rB.y) / 1.0f;
        exy = -iiA * rA.x * rA.y / 1.0f + -iiB * rB.x * rB.y / 1.0f;
        eyy = imA + imB + iiA * square(rA.x) / 1.0f + iiB * square(rB.x) / 1.0f;
    } else {
        exx = imA + iiA * square(rA.y) / 1.0f + imB + iiB * square(rB.y) / 1.0f;
        exy = -iiA * rA.x * rA.y / 1.0f + -iiB * rB.x * rB.y / 1.0f;
        eyy = imA + iiA * square(rA.x) / 1.0f + imB + iiB * square(rB.x) / 1.0f;
    }
    { // Invert(Matrix22), Math.hpp:430-442: columns (exx, exy), (exy, eyy)
        const float cp = exx * eyy - exy * exy;
        if (!almost_zero(cp)) {
            const float inverse = 1 / cp;
            wk.m[0] = eyy * inverse; wk.m[1] = -exy * inverse; wk.m[2] = -exy * inverse; wk.m[3] = exx * inverse;
        } else wk.m[0] = wk.m[1] = wk.m[2] = wk.m[3] = 0.0f;
    }
    const float invRotInertia = iiA + iiB;
    wk.angularMass = (invRotInertia > 0.0f) ? 1.0f / invRotInertia : 0.0f;
    if (motor) {
        const float2 le = (v2(pB.x, pB.y) + rB) - (v2(pA.x, pA.y) + rA);
        wk.m[4] = le.x; wk.m[5] = le.y;
        wk.m[6] = (pB.z - pA.z) - j.reference_angle;
    }
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio; j.impulse_y *= conf.dt_ratio; j.impulse_z *= conf.dt_ratio;
        const float2 P = v2(j.impulse_x, j.impulse_y);
        const float crossAP = cross(rA, P) / 1.0f, crossBP = cross(rB, P) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * (crossAP + j.impulse_z);
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * (crossBP + j.impulse_z);
    } else { j.impulse_x = j.impulse_y = 0.0f; j.impulse_z = 0.0f; }
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void fm_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const pr_step_conf& conf, const JView& v, bool motor)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const float2 rA = v2(wk.rAx, wk.rAy), rB = v2(wk.rBx, wk.rBy);
    const float h = conf.delta_time;
    const float inv_h = (h != 0.0f) ? 1.0f / h : 0.0f;
    const float maxForce = j.motor_speed, maxTorque = j.max_motor_torque, correction = j.lower_angle;
    bool solved = true;
    { // angular
        const float Cdot = motor ? (velB.z - velA.z) + inv_h * correction * wk.m[6] : (velB.z - velA.z);
        const float angularImpulse = -wk.angularMass * Cdot;
        const float oldAngularImpulse = j.impulse_z;
        const float maxAngularImpulse = h * maxTorque;
        j.impulse_z = stdclamp(oldAngularImpulse + angularImpulse, -maxAngularImpulse, maxAngularImpulse);
        const float inc = j.impulse_z - oldAngularImpulse;
        if (inc != 0.0f) solved = false;
        if (!motor || inc != 0.0f) { velA.z -= iiA * inc; velB.z += iiB * inc; }
    }
    { // linear
        const float2 vb = v2(velB.x, velB.y) + (rev_perp(rB) * (velB.z / 1.0f));
        const float2 va = v2(velA.x, velA.y) + (rev_perp(rA) * (velA.z / 1.0f));
        const float2 Cdot = motor ? (vb - va) + (inv_h * correction) * v2(wk.m[4], wk.m[5]) : vb - va;
        const float2 impulse = -mat22_transform(wk.m, Cdot);
        const float2 oldImpulse = v2(j.impulse_x, j.impulse_y);
        float2 acc = oldImpulse + impulse;
        const float maxImpulse = h * maxForce;
        if (magsq(acc) > square(maxImpulse)) { const Rot u = unit_vector(acc, mkrot(0.0f, 0.0f)); acc = u * maxImpulse; }
        j.impulse_x = acc.x; j.impulse_y = acc.y;
        const float2 inc = acc - oldImpulse;
        const float angImpulseA = cross(rA, inc) / 1.0f, angImpulseB = cross(rB, inc) / 1.0f;
        const bool any = motor ? (inc.x != 0.0f || inc.y != 0.0f || angImpulseA != 0.0f || angImpulseB != 0.0f) : (inc.x != 0.0f || inc.y != 0.0f);
        if (any) solved = false;
        if (!motor || any) {
            const float2 la = v2(velA.x, velA.y) - imA * inc; velA.x = la.x; velA.y = la.y; velA.z -= iiA * angImpulseA;
            const float2 lb = v2(velB.x, velB.y) + imB * inc; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * angImpulseB;
        }
    }
    d.jt[ji] = j;
    if (motor && solved) return; // (:222-224: the velocities are what they were)
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    if (!solved) v.isl_jflag[wk.island] = 1u;
}

// ---- WeldJointConf.cpp:35-371. Row overlay (pr_weld_joint): impulse = impulse_x/y/z, referenceAngle = reference_angle, frequency =
// lower_angle, dampingRatio = upper_angle; work: rA, rB, m[3 * i + k] = mass[i][k], angularMass = gamma, ex[0] = bias
PR_D void weld_k(float* m, float imA, float2 rA, float iiA, float imB, float2 rB, float iiB)
{ // GetMat33, :35-62
    const float exx = imA + square(rA.y) * iiA / 1.0f + imB + square(rB.y) * iiB / 1.0f;
    const float eyx = -rA.y * rA.x * iiA / 1.0f + -rB.y * rB.x * iiB / 1.0f;
    const float ezx = -rA.y * iiA * 1.0f / 1.0f + -rB.y * iiB * 1.0f / 1.0f;
    const float eyy = imA + square(rA.x) * iiA / 1.0f + imB + square(rB.x) * iiB / 1.0f;
    const float ezy = rA.x * iiA * 1.0f / 1.0f + rB.x * iiB * 1.0f / 1.0f;
    const float ezz = (iiA + iiB) * 1.0f / 1.0f;
    m[0] = exx; m[3] = eyx; m[6] = ezx;
    m[1] = m[3]; m[4] = eyy; m[7] = ezy;
    m[2] = m[6]; m[5] = m[7]; m[8] = ezz;
}
PR_D void mat33_inverse22(const float* k, float* r)
{ // GetInverse22, Math.hpp:478-489
    const float a = k[0], b = k[3], c = k[1], dd = k[4];
    float det = (a * dd) - (b * c);
    if (det != 0.0f) det = 1.0f / det;
    r[0] = det * dd; r[1] = -det * c; r[2] = 0.0f;
    r[3] = -det * b; r[4] = det * a; r[5] = 0.0f;
    r[6] = 0.0f; r[7] = 0.0f; r[8] = 0.0f;
}
PR_D void mat33_sym_inverse33(const float* k, float* r)
{ // GetSymInverse33, Math.hpp:493-509
    const float invDet = dot3(mcol(k, 0), cross3(mcol(k, 1), mcol(k, 2)));
    const float det = (invDet != 0.0f) ? (1.0f / invDet) : invDet;
    const float a11 = k[0], a12 = k[3], a13 = k[6], a22 = k[4], a23 = k[7], a33 = k[8];
    const float ex_y = det * (a13 * a23 - a12 * a33);
    const float ey_z = det * (a13 * a12 - a11 * a23);
    const float ex_z = det * (a12 * a23 - a13 * a22);
    r[0] = det * (a22 * a33 - a23 * a23); r[1] = ex_y; r[2] = ex_z;
    r[3] = ex_y; r[4] = det * (a11 * a33 - a13 * a13); r[5] = ey_z;
    r[6] = ex_z; r[7] = ey_z; r[8] = det * (a11 * a22 - a12 * a12);
}
PR_D float2 mat33_transform2(const float* A, float2 v)
{ // Transform(const Vec2&, const Mat33&), Math.hpp:522-526
    return v2(A[0] * v.x + A[3] * v.y, A[1] * v.x + A[4] * v.y);
}
PR_D V3 mat33_transform3(const float* m, V3 v)
{ // Transform(Vec3, Mat33) = m * v, Vector.hpp:458-472 (each element accumulated from zero)
    V3 r;
    r.x = 0.0f; r.x += m[0] * v.x; r.x += m[1] * v.y; r.x += m[2] * v.z;
    r.y = 0.0f; r.y += m[3] * v.x; r.y += m[4] * v.y; r.y += m[5] * v.z;
    r.z = 0.0f; r.z += m[6] * v.x; r.z += m[7] * v.y; r.z += m[8] * v.z;
    return r;
}
PR_D void weld_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const Rot qA = rot_from_angle(pA.z), qB = rot_from_angle(pB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    float K[9];
    weld_k(K, imA, rA, iiA, imB, rB, iiB);
    const float frequency = j.lower_angle, dampingRatio = j.upper_angle;
    if (frequency > 0.0f) {
        mat33_inverse22(K, wk.m);
        float invRotInertia = iiA + iiB;
        const float rotInertia = (invRotInertia > 0.0f) ? 1.0f / invRotInertia : 0.0f;
        const float C = pB.z - pA.z - j.reference_angle;
        const float omega = 2.0f * 3.14159265358979323846264338327950288f * frequency;
        const float dd = 2.0f * rotInertia * dampingRatio * omega;
        const float k = rotInertia * omega * omega;
        const float h = conf.delta_time;
        const float invGamma = h * (dd + h * k);
        wk.angularMass = (invGamma != 0.0f) ? 1.0f / invGamma : 0.0f;
        wk.ex[0] = C * h * k * wk.angularMass;
        invRotInertia += wk.angularMass;
        wk.m[8] = (invRotInertia != 0.0f) ? 1.0f / invRotInertia : 0.0f;
    } else if (K[8] == 0.0f) {
        mat33_inverse22(K, wk.m); wk.angularMass = 0.0f; wk.ex[0] = 0.0f;
    } else {
        mat33_sym_inverse33(K, wk.m); wk.angularMass = 0.0f; wk.ex[0] = 0.0f;
    }
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio; j.impulse_y *= conf.dt_ratio; j.impulse_z *= conf.dt_ratio;
        const float2 P = v2(j.impulse_x * 1.0f, j.impulse_y * 1.0f);
        const float L = j.impulse_z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = L + cross(rA, P) / 1.0f;
        const float LB = L + cross(rB, P) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else j.impulse_x = j.impulse_y = j.impulse_z = 0.0f;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void weld_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 oldA = v.s_vel[ia], oldB = v.s_vel[ib];
    float4 velA = oldA, velB = oldB;
    const float2 rA = v2(wk.rAx, wk.rAy), rB = v2(wk.rBx, wk.rBy);
    if (j.lower_angle > 0.0f) {
        const float Cdot2 = velB.z - velA.z;
        const float gamma = wk.angularMass * j.impulse_z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float impulse2 = -wk.m[8] * (Cdot2 + wk.ex[0] + gamma);
        j.impulse_z += impulse2;
        velA.z -= iiA * impulse2 * 1.0f * 1.0f / (1.0f * 1.0f);
        velB.z += iiB * impulse2 * 1.0f * 1.0f / (1.0f * 1.0f);
        const float2 vb = v2(velB.x, velB.y) + rev_perp(rB) * (velB.z / 1.0f);
        const float2 va = v2(velA.x, velA.y) + rev_perp(rA) * (velA.z / 1.0f);
        const float2 Cdot1 = vb - va;
        const float2 impulse1 = -mat33_transform2(wk.m, v2(Cdot1.x / 1.0f, Cdot1.y / 1.0f));
        j.impulse_x += impulse1.x; j.impulse_y += impulse1.y;
        const float2 P = v2(impulse1.x * 1.0f, impulse1.y * 1.0f);
        const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else {
        const float2 vb = v2(velB.x, velB.y) + rev_perp(rB) * (velB.z / 1.0f);
        const float2 va = v2(velA.x, velA.y) + rev_perp(rA) * (velA.z / 1.0f);
        const float2 Cdot1 = vb - va;
        V3 Cdot; Cdot.x = Cdot1.x / 1.0f; Cdot.y = Cdot1.y / 1.0f; Cdot.z = (velB.z - velA.z) / 1.0f;
        const V3 t = mat33_transform3(wk.m, Cdot);
        V3 impulse; impulse.x = -t.x; impulse.y = -t.y; impulse.z = -t.z;
        j.impulse_x += impulse.x; j.impulse_y += impulse.y; j.impulse_z += impulse.z;
        const float2 P = v2(impulse.x * 1.0f, impulse.y * 1.0f);
        const float L = impulse.z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = L + cross(rA, P) / 1.0f;
        const float LB = L + cross(rB, P) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    }
    d.jt[ji] = j;
    const bool changed = velA.x != oldA.x || velA.y != oldA.y || velA.z != oldA.z || velB.x != oldB.x || velB.y != oldB.y || velB.z != oldB.z;
    if (changed) { v.s_vel[ia] = velA; v.s_vel[ib] = velB; v.isl_jflag[wk.island] = 1u; }
}
PR_D void weld_solve_position(const DV& d, const pr_revolute_joint& j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    float4 posA = v.s_pos[ia], posB = v.s_pos[ib];
    const Rot qA = rot_from_angle(posA.z), qB = rot_from_angle(posB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    float positionError, angularError;
    float K[9];
    weld_k(K, imA, rA, iiA, imB, rB, iiB);
    const float2 C1 = (v2(posB.x, posB.y) + rB) - (v2(posA.x, posA.y) + rA);
    positionError = sqrtf(magsq(C1));
    if (j.lower_angle > 0.0f) {
        angularError = 0.0f;
        const float2 P = -solve22(K, C1) * 1.0f;
        const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
        const float2 la = v2(posA.x, posA.y) - imA * P; posA.x = la.x; posA.y = la.y; posA.z -= iiA * LA;
        const float2 lb = v2(posB.x, posB.y) + imB * P; posB.x = lb.x; posB.y = lb.y; posB.z += iiB * LB;
    } else {
        const float C2 = posB.z - posA.z - j.reference_angle;
        angularError = fabsf(C2);
        V3 impulse;
        if (K[8] > 0.0f) {
            V3 C; C.x = C1.x; C.y = C1.y; C.z = C2;
            const V3 s = solve33(K, C);
            impulse.x = -s.x; impulse.y = -s.y; impulse.z = -s.z;
        } else {
            const float2 impulse2 = -solve22(K, C1);
            impulse.x = impulse2.x; impulse.y = impulse2.y; impulse.z = 0.0f;
        }
        const float2 P = v2(impulse.x * 1.0f, impulse.y * 1.0f) * 1.0f;
        const float L = impulse.z * 1.0f * 1.0f / 1.0f;
        const float LA = L + cross(rA, P) / 1.0f;
        const float LB = L + cross(rB, P) / 1.0f;
        const float2 la = v2(posA.x, posA.y) - imA * P; posA.x = la.x; posA.y = la.y; posA.z -= iiA * LA;
        const float2 lb = v2(posB.x, posB.y) + imB * P; posB.x = lb.x; posB.y = lb.y; posB.z += iiB * LB;
    }
    v.s_pos[ia] = posA; v.s_pos[ib] = posB;
    if (!((positionError <= conf.linear_slop) && (angularError <= conf.angular_slop))) v.isl_jflag[wk.island] = 1u;
}

// ---- TargetJointConf.cpp:52-243. One body: the row has body_a == body_b (the reference's bodyA is InvalidBodyID). Row overlay
// (pr_target_joint): target = local_anchor_a, localAnchorB = local_anchor_b, impulse = impulse_x/y, frequency = lower_angle,
// dampingRatio = upper_angle, maxForce = motor_speed; work: rB, m[0..3] = mass, angularMass = gamma, ex[0..1] = C
PR_D void target_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ib = jv_idx(v.slotV, j.body_b);
    const float4 mB = d.b_mass[j.body_b];
    const float imB = mB.x, iiB = mB.y;
    const float4 pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velB = v.s_vel[ib];
    const Rot qB = rot_from_angle(pB.z);
    const float mass = (imB != 0.0f) ? (1.0f / imB) : 0.0f;
    const float omega = 2.0f * 3.14159265358979323846264338327950288f * j.lower_angle;
    const float dd = 2.0f * mass * j.upper_angle * omega;
    const float k = mass * square(omega);
    const float h = conf.delta_time;
    const float tmp = dd + h * k;
    const float invGamma = h * tmp;
    const float gamma = (invGamma != 0.0f) ? 1.0f / invGamma : 0.0f;
    wk.angularMass = gamma;
    const float beta = h * k * gamma;
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = 0.0f; wk.rAy = 0.0f; wk.rBx = rB.x; wk.rBy = rB.y;
    { // GetEffectiveMassMatrix, :52-75
        const float exx = imB + (iiB * square(rB.y) / 1.0f) + gamma;
        const float exy = -iiB * rB.x * rB.y / 1.0f;
        const float eyy = imB + (iiB * square(rB.x) / 1.0f) + gamma;
        const float cp = exx * eyy - exy * exy; // Invert(Matrix22), Math.hpp:430-442
        if (!almost_zero(cp)) {
            const float inverse = 1 / cp;
            wk.m[0] = eyy * inverse; wk.m[1] = -exy * inverse; wk.m[2] = -exy * inverse; wk.m[3] = exx * inverse;
        } else wk.m[0] = wk.m[1] = wk.m[2] = wk.m[3] = 0.0f;
    }
    const float2 C = ((v2(pB.x, pB.y) + rB) - v2(j.local_anchor_a.x, j.local_anchor_a.y)) * beta;
    wk.ex[0] = C.x; wk.ex[1] = C.y;
    velB.z *= 0.98f; // "cheat with some damping" (:125-127)
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio; j.impulse_y *= conf.dt_ratio;
        const float2 P = v2(j.impulse_x, j.impulse_y);
        const float crossBP = cross(rB, P) / 1.0f;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * crossBP;
    } else j.impulse_x = j.impulse_y = 0.0f;
    v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void target_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ib = jv_idx(v.slotV, j.body_b);
    const float4 mB = d.b_mass[j.body_b];
    float4 velB = v.s_vel[ib];
    const float2 rB = v2(wk.rBx, wk.rBy);
    const float2 Cdot = v2(velB.x, velB.y) + (rev_perp(rB) * (velB.z / 1.0f));
    const float2 ev = Cdot + (v2(wk.ex[0], wk.ex[1]) + (wk.angularMass * v2(j.impulse_x, j.impulse_y)));
    const float2 oldImpulse = v2(j.impulse_x, j.impulse_y);
    const float2 addImpulse = mat22_transform(wk.m, -ev);
    float2 impulse = oldImpulse + addImpulse;
    const float maxImpulse = conf.delta_time * j.motor_speed;
    if (magsq(impulse) > square(maxImpulse)) { const Rot u = unit_vector(impulse, mkrot(0.0f, 0.0f)); impulse = u * maxImpulse; }
    j.impulse_x = impulse.x; j.impulse_y = impulse.y;
    const float2 inc = impulse - oldImpulse;
    const float angImpulseB = cross(rB, inc) / 1.0f;
    const float2 lb = v2(velB.x, velB.y) + mB.x * inc; velB.x = lb.x; velB.y = lb.y; velB.z += mB.y * angImpulseB;
    v.s_vel[ib] = velB;
    d.jt[ji] = j;
    if (!(inc.x == 0.0f && inc.y == 0.0f)) v.isl_jflag[wk.island] = 1u;
}

// ---- PrismaticJointConf.cpp:150-533. Row overlay (pr_prismatic_joint): impulse = impulse_x/y/z, motorImpulse = angular_motor_impulse,
// localXAxisA = the bits of enable_limit / enable_motor (localYAxisA is its reverse perpendicular), enableLimit / enableMotor = bits
// 16 / 17 of collide_connected, lower / upperTranslation = lower_angle / upper_angle, motorSpeed, maxMotorForce = max_motor_torque;
// work: m[3 * i + k] = K[i][k], angularMass = motorMass, ex[0..1] = axis, ex[2..3] = perp, ex[4] = s1, ex[5] = s2, ex[6] = a1, ex[7] = a2
PR_D bool j_enable_limit(const pr_revolute_joint& j) { return (j.collide_connected & PR_JOINT_ENABLE_LIMIT) != 0u; }
PR_D bool j_enable_motor(const pr_revolute_joint& j) { return (j.collide_connected & PR_JOINT_ENABLE_MOTOR) != 0u; }
PR_D Rot j_local_x_axis(const pr_revolute_joint& j) { return mkrot(u2f(j.enable_limit), u2f(j.enable_motor)); }
PR_D void prismatic_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const Rot qA = rot_from_angle(pA.z), qB = rot_from_angle(pB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    const float2 dd = (v2(pB.x, pB.y) - v2(pA.x, pA.y)) + rB - rA;
    const Rot lx = j_local_x_axis(j);
    const Rot axis = rot_rotate(lx, qA);
    const float a1 = cross(dd + rA, axis), a2 = cross(rB, axis);
    const float invRotMassA = iiA * square(a1) / 1.0f, invRotMassB = iiB * square(a2) / 1.0f;
    const float totalInvMass = imA + imB + invRotMassA + invRotMassB;
    wk.angularMass = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
    const Rot perp = rot_rotate(rev_perp(lx), qA);
    const float s1 = cross(dd + rA, perp), s2 = cross(rB, perp);
    {
        const float invRotMassA2 = iiA * square(s1) / 1.0f, invRotMassB2 = iiB * square(s2) / 1.0f;
        const float k11 = imA + imB + invRotMassA2 + invRotMassB2;
        const float k12 = (iiA * s1 + iiB * s2) * 1.0f * 1.0f / 1.0f;
        const float k13 = (iiA * s1 * a1 + iiB * s2 * a2) / 1.0f;
        const float totalInvRotInertia = iiA + iiB;
        const float k22 = (totalInvRotInertia == 0.0f) ? 1.0f : totalInvRotInertia;
        const float k23 = (iiA * a1 + iiB * a2) * 1.0f * 1.0f / 1.0f;
        const float k33 = totalInvMass;
        float* m = wk.m;
        m[0] = k11; m[1] = k12; m[2] = k13;
        m[3] = k12; m[4] = k22; m[5] = k23;
        m[6] = k13; m[7] = k23; m[8] = k33;
    }
    wk.ex[0] = axis.c; wk.ex[1] = axis.s; wk.ex[2] = perp.c; wk.ex[3] = perp.s; wk.ex[4] = s1; wk.ex[5] = s2; wk.ex[6] = a1; wk.ex[7] = a2;
    if (j_enable_limit(j)) {
        const float jointTranslation = dot(axis, dd);
        if (fabsf(j.upper_angle - j.lower_angle) < (conf.linear_slop * 2.0f)) j.limit_state = LS_EQUAL;
        else if (jointTranslation <= j.lower_angle) {
            if (j.limit_state != LS_AT_LOWER) { j.limit_state = LS_AT_LOWER; j.impulse_z = 0; }
        } else if (jointTranslation >= j.upper_angle) {
            if (j.limit_state != LS_AT_UPPER) { j.limit_state = LS_AT_UPPER; j.impulse_z = 0; }
        } else { j.limit_state = LS_INACTIVE; j.impulse_z = 0; }
    } else { j.limit_state = LS_INACTIVE; j.impulse_z = 0; }
    if (!j_enable_motor(j)) j.angular_motor_impulse = 0.0f;
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio; j.impulse_y *= conf.dt_ratio; j.impulse_z *= conf.dt_ratio;
        j.angular_motor_impulse *= conf.dt_ratio;
        const float2 ulImpulseX = j.impulse_x * perp;
        const float2 Px = v2(ulImpulseX.x * 1.0f, ulImpulseX.y * 1.0f);
        const float Pxs1 = j.impulse_x * s1 * 1.0f / 1.0f, Pxs2 = j.impulse_x * s2 * 1.0f / 1.0f;
        const float PzLength = j.angular_motor_impulse + j.impulse_z * 1.0f;
        const float2 Pz = PzLength * axis;
        const float2 P = Px + Pz;
        const float L = j.impulse_y * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = L + (Pxs1 * 1.0f + PzLength * a1) / 1.0f;
        const float LB = L + (Pxs2 * 1.0f + PzLength * a2) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else { j.impulse_x = j.impulse_y = j.impulse_z = 0.0f; j.angular_motor_impulse = 0.0f; }
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void prismatic_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 oldA = v.s_vel[ia], oldB = v.s_vel[ib];
    float4 velA = oldA, velB = oldB;
    const Rot axis = mkrot(wk.ex[0], wk.ex[1]), perp = mkrot(wk.ex[2], wk.ex[3]);
    const float s1 = wk.ex[4], s2 = wk.ex[5], a1 = wk.ex[6], a2 = wk.ex[7];
    if (j_enable_motor(j) && j.limit_state != LS_EQUAL) {
        const float vDot = dot(axis, v2(velB.x, velB.y) - v2(velA.x, velA.y));
        const float Cdot = vDot + (a2 * velB.z - a1 * velA.z) / 1.0f;
        float impulse = wk.angularMass * (j.motor_speed * 1.0f / 1.0f - Cdot);
        const float oldImpulse = j.angular_motor_impulse;
        const float maxImpulse = conf.delta_time * j.max_motor_torque;
        j.angular_motor_impulse = stdclamp(j.angular_motor_impulse + impulse, -maxImpulse, maxImpulse);
        impulse = j.angular_motor_impulse - oldImpulse;
        const float2 P = impulse * axis;
        const float LA = impulse * a1 / 1.0f, LB = impulse * a2 / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    }
    const float2 velDelta = v2(velB.x, velB.y) - v2(velA.x, velA.y);
    const float sRotSpeed = (s2 * velB.z - s1 * velA.z) / 1.0f;
    const float2 Cdot1 = v2(dot(perp, velDelta) + sRotSpeed, velB.z - velA.z);
    if (j_enable_limit(j) && (j.limit_state != LS_INACTIVE)) {
        const float deltaDot = dot(axis, velDelta);
        const float aRotSpeed = (a2 * velB.z - a1 * velA.z) / 1.0f;
        const float Cdot2 = deltaDot + aRotSpeed;
        V3 nC; nC.x = -Cdot1.x; nC.y = -Cdot1.y; nC.z = -Cdot2;
        V3 f1; f1.x = j.impulse_x; f1.y = j.impulse_y; f1.z = j.impulse_z;
        const V3 sol = solve33(wk.m, nC);
        V3 imp; imp.x = f1.x + sol.x; imp.y = f1.y + sol.y; imp.z = f1.z + sol.z;
        if (j.limit_state == LS_AT_LOWER) imp.z = stdmax(imp.z, 0.0f);
        else if (j.limit_state == LS_AT_UPPER) imp.z = stdmin(imp.z, 0.0f);
        const float2 b = -Cdot1 - (imp.z - f1.z) * v2(wk.m[6], wk.m[7]);
        const float2 f2r = solve22(wk.m, b) + v2(f1.x, f1.y);
        imp.x = f2r.x; imp.y = f2r.y;
        j.impulse_x = imp.x; j.impulse_y = imp.y; j.impulse_z = imp.z;
        V3 df; df.x = imp.x - f1.x; df.y = imp.y - f1.y; df.z = imp.z - f1.z;
        const float2 ulP = df.x * perp + df.z * axis;
        const float2 P = v2(ulP.x * 1.0f, ulP.y * 1.0f);
        const float LA = (df.x * s1 + df.y * 1.0f + df.z * a1) * 1.0f / 1.0f;
        const float LB = (df.x * s2 + df.y * 1.0f + df.z * a2) * 1.0f / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else {
        const float2 df = solve22(wk.m, -Cdot1);
        j.impulse_x += df.x; j.impulse_y += df.y;
        const float2 ulP = df.x * perp;
        const float2 P = v2(ulP.x * 1.0f, ulP.y * 1.0f);
        const float LA = (df.x * s1 + df.y * 1.0f) * 1.0f / 1.0f;
        const float LB = (df.x * s2 + df.y * 1.0f) * 1.0f / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    }
    d.jt[ji] = j;
    const bool changed = velA.x != oldA.x || velA.y != oldA.y || velA.z != oldA.z || velB.x != oldB.x || velB.y != oldB.y || velB.z != oldB.z;
    if (changed) { v.s_vel[ia] = velA; v.s_vel[ib] = velB; v.isl_jflag[wk.island] = 1u; }
}
PR_D void prismatic_solve_position(const DV& d, const pr_revolute_joint& j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    float4 posA = v.s_pos[ia], posB = v.s_pos[ib];
    const Rot qA = rot_from_angle(posA.z), qB = rot_from_angle(posB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    const float2 dd = (v2(posB.x, posB.y) + rB) - (v2(posA.x, posA.y) + rA);
    const Rot lx = j_local_x_axis(j);
    const Rot axis = rot_rotate(lx, qA);
    const float a1 = cross(dd + rA, axis), a2 = cross(rB, axis);
    const Rot perp = rot_rotate(rev_perp(lx), qA);
    const float s1 = cross(dd + rA, perp), s2 = cross(rB, perp);
    const float2 C1 = v2(dot(perp, dd) / 1.0f, (posB.z - posA.z - j.reference_angle) / 1.0f);
    float linearError = fabsf(C1.x) * 1.0f;
    const float angularError = fabsf(C1.y) * 1.0f;
    bool active = false;
    float C2 = 0.0f;
    if (j_enable_limit(j)) {
        const float translation = dot(axis, dd);
        if (fabsf(j.upper_angle - j.lower_angle) < (2.0f * conf.linear_slop)) {
            C2 = stdclamp(translation, -conf.max_linear_correction, conf.max_linear_correction);
            linearError = stdmax(linearError, fabsf(translation));
            active = true;
        } else if (translation <= j.lower_angle) {
            C2 = stdclamp(translation - j.lower_angle + conf.linear_slop, -conf.max_linear_correction, 0.0f);
            linearError = stdmax(linearError, j.lower_angle - translation);
            active = true;
        } else if (translation >= j.upper_angle) {
            C2 = stdclamp(translation - j.upper_angle - conf.linear_slop, 0.0f, conf.max_linear_correction);
            linearError = stdmax(linearError, translation - j.upper_angle);
            active = true;
        }
    }
    V3 impulse;
    if (active) {
        const float k11 = imA + iiA * square(s1) / 1.0f + imB + iiB * square(s2) / 1.0f;
        const float k12 = iiA * s1 * 1.0f / 1.0f + iiB * s2 * 1.0f / 1.0f;
        const float k13 = iiA * s1 * a1 / 1.0f + iiB * s2 * a2 / 1.0f;
        float k22 = iiA + iiB;
        if (k22 == 0.0f) k22 = 1.0f * 1.0f / (1.0f * 1.0f);
        const float k23 = iiA * a1 * 1.0f / 1.0f + iiB * a2 * 1.0f / 1.0f;
        const float k33 = imA + iiA * square(a1) / 1.0f + imB + iiB * square(a2) / 1.0f;
        float K[9];
        K[0] = k11; K[1] = k12; K[2] = k13; K[3] = k12; K[4] = k22; K[5] = k23; K[6] = k13; K[7] = k23; K[8] = k33;
        V3 nC; nC.x = -C1.x; nC.y = -C1.y; nC.z = -C2;
        impulse = solve33(K, nC);
    } else {
        const float k11 = imA + iiA * s1 * s1 / 1.0f + imB + iiB * s2 * s2 / 1.0f;
        const float k12 = iiA * s1 * 1.0f / 1.0f + iiB * s2 * 1.0f / 1.0f;
        float k22 = iiA + iiB;
        if (k22 == 0.0f) k22 = 1.0f;
        const float cp = cross(v2(k11, k12), v2(k12, k22)); // Solve(Matrix22, Vector2), Math.hpp:414-426
        const float2 b = -C1;
        float2 impulse1 = v2(0.0f, 0.0f);
        if (!almost_zero(cp)) {
            const float inverse = 1 / cp;
            impulse1 = v2((k22 * b.x - k12 * b.y) * inverse, (k11 * b.y - k12 * b.x) * inverse);
        }
        impulse.x = impulse1.x; impulse.y = impulse1.y; impulse.z = 0.0f;
    }
    const float2 P = (impulse.x * perp + impulse.z * axis) * 1.0f * 1.0f;
    const float LA = (impulse.x * s1 + impulse.y * 1.0f + impulse.z * a1) * 1.0f * 1.0f / 1.0f;
    const float LB = (impulse.x * s2 + impulse.y * 1.0f + impulse.z * a2) * 1.0f * 1.0f / 1.0f;
    const float2 la = v2(posA.x, posA.y) - imA * P; posA.x = la.x; posA.y = la.y; posA.z -= iiA * LA;
    const float2 lb = v2(posB.x, posB.y) + imB * P; posB.x = lb.x; posB.y = lb.y; posB.z += iiB * LB;
    v.s_pos[ia] = posA; v.s_pos[ib] = posB;
    if (!((linearError <= conf.linear_slop) && (angularError <= conf.angular_slop))) v.isl_jflag[wk.island] = 1u;
}

// ---- WheelJointConf.cpp:103-362. Row overlay (pr_wheel_joint): impulse = impulse_x, springImpulse = impulse_y, angularImpulse = impulse_z,
// localXAxisA = the bits of enable_limit / enable_motor, enableMotor = bit 17 of collide_connected, frequency = lower_angle, dampingRatio =
// upper_angle, motorSpeed, maxMotorTorque; work: m[0..1] = ay, m[2] = sAy, m[3] = sBy, m[4] = mass, m[5..6] = ax, m[7] = sAx, m[8] = sBx,
// ex[0] = springMass, ex[1] = bias, ex[2] = gamma, angularMass
PR_D void wheel_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const Rot qA = rot_from_angle(pA.z), qB = rot_from_angle(pB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    const float2 dd = (v2(pB.x, pB.y) + rB) - (v2(pA.x, pA.y) + rA);
    const Rot lx = j_local_x_axis(j);
    const Rot ay = rot_rotate(rev_perp(lx), qA);
    const float sAy = cross(dd + rA, ay), sBy = cross(rB, ay);
    {
        const float invRotMassA = iiA * square(sAy) / 1.0f, invRotMassB = iiB * square(sBy) / 1.0f;
        const float invMass = imA + imB + invRotMassA + invRotMassB;
        wk.m[4] = (invMass > 0.0f) ? 1.0f / invMass : 0.0f;
    }
    wk.m[0] = ay.c; wk.m[1] = ay.s; wk.m[2] = sAy; wk.m[3] = sBy;
    float springMass = 0.0f, bias = 0.0f, gamma = 0.0f;
    Rot ax = mkrot(0.0f, 0.0f);
    float sAx = 0.0f, sBx = 0.0f;
    if (j.lower_angle > 0.0f) {
        ax = rot_rotate(lx, qA);
        sAx = cross(dd + rA, ax); sBx = cross(rB, ax);
        const float invRotMassA = iiA * square(sAx) / 1.0f, invRotMassB = iiB * square(sBx) / 1.0f;
        const float invMass = imA + imB + invRotMassA + invRotMassB;
        if (invMass > 0.0f) {
            springMass = 1.0f / invMass;
            const float C = dot(dd, ax);
            const float omega = 2.0f * 3.14159265358979323846264338327950288f * j.lower_angle;
            const float dc = 2.0f * springMass * j.upper_angle * omega;
            const float k = springMass * omega * omega;
            const float h = conf.delta_time;
            const float invGamma = h * (dc + h * k);
            gamma = (invGamma > 0.0f) ? 1.0f / invGamma : 0.0f;
            bias = C * h * k * gamma;
            const float totalInvMass = invMass + gamma;
            springMass = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
        }
    } else j.impulse_y = 0.0f;
    wk.m[5] = ax.c; wk.m[6] = ax.s; wk.m[7] = sAx; wk.m[8] = sBx;
    wk.ex[0] = springMass; wk.ex[1] = bias; wk.ex[2] = gamma;
    if (j_enable_motor(j)) {
        const float invRotInertia = iiA + iiB;
        wk.angularMass = (invRotInertia > 0.0f) ? 1.0f / invRotInertia : 0.0f;
    } else { wk.angularMass = 0.0f; j.impulse_z = 0.0f; }
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio; j.impulse_y *= conf.dt_ratio; j.impulse_z *= conf.dt_ratio;
        const float2 P = j.impulse_x * ay + j.impulse_y * ax;
        const float LA = (j.impulse_x * sAy + j.impulse_y * sAx) / 1.0f + j.impulse_z;
        const float LB = (j.impulse_x * sBy + j.impulse_y * sBx) / 1.0f + j.impulse_z;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else j.impulse_x = j.impulse_y = j.impulse_z = 0.0f;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void wheel_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 oldA = v.s_vel[ia], oldB = v.s_vel[ib];
    float4 velA = oldA, velB = oldB;
    const Rot ay = mkrot(wk.m[0], wk.m[1]), ax = mkrot(wk.m[5], wk.m[6]);
    const float sAy = wk.m[2], sBy = wk.m[3], sAx = wk.m[7], sBx = wk.m[8];
    { // spring
        const float dt = dot(ax, v2(velB.x, velB.y) - v2(velA.x, velA.y));
        const float Cdot = dt + sBx * velB.z / 1.0f - sAx * velA.z / 1.0f;
        const float impulse = -wk.ex[0] * (Cdot + wk.ex[1] + wk.ex[2] * j.impulse_y);
        j.impulse_y += impulse;
        const float2 P = impulse * ax;
        const float LA = impulse * sAx / 1.0f, LB = impulse * sBx / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    }
    { // rotational motor
        const float Cdot = (velB.z - velA.z - j.motor_speed);
        float impulse = -wk.angularMass * Cdot;
        const float oldImpulse = j.impulse_z;
        const float maxImpulse = conf.delta_time * j.max_motor_torque;
        j.impulse_z = stdclamp(j.impulse_z + impulse, -maxImpulse, maxImpulse);
        impulse = j.impulse_z - oldImpulse;
        velA.z -= iiA * impulse;
        velB.z += iiB * impulse;
    }
    { // point to line
        const float dt = dot(ay, v2(velB.x, velB.y) - v2(velA.x, velA.y));
        const float Cdot = dt + sBy * velB.z / 1.0f - sAy * velA.z / 1.0f;
        const float impulse = -wk.m[4] * Cdot;
        j.impulse_x += impulse;
        const float2 P = impulse * ay;
        const float LA = impulse * sAy / 1.0f, LB = impulse * sBy / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    }
    d.jt[ji] = j;
    const bool changed = velA.x != oldA.x || velA.y != oldA.y || velA.z != oldA.z || velB.x != oldB.x || velB.y != oldB.y || velB.z != oldB.z;
    if (changed) { v.s_vel[ia] = velA; v.s_vel[ib] = velB; v.isl_jflag[wk.island] = 1u; }
}
PR_D void wheel_solve_position(const DV& d, const pr_revolute_joint& j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    float4 posA = v.s_pos[ia], posB = v.s_pos[ib];
    const Rot qA = rot_from_angle(posA.z), qB = rot_from_angle(posB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    const float2 dd = (v2(posB.x, posB.y) - v2(posA.x, posA.y)) + (rB - rA);
    const Rot ay = rot_rotate(rev_perp(j_local_x_axis(j)), qA);
    const float sAy = cross(dd + rA, ay), sBy = cross(rB, ay);
    const float C = dot(dd, ay);
    // (:339-340: the inverse rotational masses come from the sAy / sBy InitVelocity stored, not from the fresh ones)
    const float invRotMassA = iiA * square(wk.m[2]) / 1.0f, invRotMassB = iiB * square(wk.m[3]) / 1.0f;
    const float k = imA + imB + invRotMassA + invRotMassB;
    const float impulse = (k != 0.0f) ? -(C / k) : 0.0f * 1.0f * 1.0f;
    const float2 P = impulse * ay;
    const float LA = impulse * sAy / 1.0f, LB = impulse * sBy / 1.0f;
    const float2 la = v2(posA.x, posA.y) - imA * P; posA.x = la.x; posA.y = la.y; posA.z -= iiA * LA;
    const float2 lb = v2(posB.x, posB.y) + imB * P; posB.x = lb.x; posB.y = lb.y; posB.z += iiB * LB;
    v.s_pos[ia] = posA; v.s_pos[ib] = posB;
    if (!(fabsf(C) <= conf.linear_slop)) v.isl_jflag[wk.island] = 1u;
}

// ---- PulleyJointConf.cpp:91-282. Row overlay (pr_pulley_joint): impulse = impulse_x, groundAnchorA = impulse_y / impulse_z, groundAnchorB =
// angular_motor_impulse / reference_angle, lengthA / lengthB = the bits of enable_limit / enable_motor, ratio = lower_angle;
// work: rA, rB, m[0..1] = uA, m[2..3] = uB, m[4] = mass
PR_D void pulley_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const Rot qA = rot_from_angle(pA.z), qB = rot_from_angle(pB.z);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    const float ratio = j.lower_angle;
    const float2 pulleyAxisA = v2(pA.x, pA.y) + rA - v2(j.impulse_y, j.impulse_z);
    const float2 pulleyAxisB = v2(pB.x, pB.y) + rB - v2(j.angular_motor_impulse, j.reference_angle);
    const Rot uA = unit_vector(pulleyAxisA, mkrot(0.0f, 0.0f)), uB = unit_vector(pulleyAxisB, mkrot(0.0f, 0.0f));
    const float ruA = cross(rA, uA), ruB = cross(rB, uB);
    const float totInvMassA = imA + (iiA * square(ruA)) / 1.0f;
    const float totInvMassB = imB + (iiB * square(ruB)) / 1.0f;
    const float totalInvMass = totInvMassA + ratio * ratio * totInvMassB;
    wk.m[0] = uA.c; wk.m[1] = uA.s; wk.m[2] = uB.c; wk.m[3] = uB.s;
    wk.m[4] = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio;
        const float2 PA = -(j.impulse_x) * uA;
        const float2 PB = (-ratio * j.impulse_x) * uB;
        const float2 la = v2(velA.x, velA.y) + imA * PA; velA.x = la.x; velA.y = la.y; velA.z += iiA * cross(rA, PA) / 1.0f;
        const float2 lb = v2(velB.x, velB.y) + imB * PB; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * cross(rB, PB) / 1.0f;
    } else j.impulse_x = 0.0f;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void pulley_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const JView& v)
{
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const float2 rA = v2(wk.rAx, wk.rAy), rB = v2(wk.rBx, wk.rBy);
    const Rot uA = mkrot(wk.m[0], wk.m[1]), uB = mkrot(wk.m[2], wk.m[3]);
    const float ratio = j.lower_angle;
    const float2 vpA = v2(velA.x, velA.y) + rev_perp(rA) * (velA.z / 1.0f);
    const float2 vpB = v2(velB.x, velB.y) + rev_perp(rB) * (velB.z / 1.0f);
    const float Cdot = -dot(uA, vpA) - ratio * dot(uB, vpB);
    const float impulse = -wk.m[4] * Cdot;
    j.impulse_x += impulse;
    const float2 PA = -impulse * uA;
    const float2 PB = -ratio * impulse * uB;
    const float2 la = v2(velA.x, velA.y) + mA.x * PA; velA.x = la.x; velA.y = la.y; velA.z += mA.y * cross(rA, PA) / 1.0f;
    const float2 lb = v2(velB.x, velB.y) + mB.x * PB; velB.x = lb.x; velB.y = lb.y; velB.z += mB.y * cross(rB, PB) / 1.0f;
    v.s_vel[ia] = velA; v.s_vel[ib] = velB;
    d.jt[ji] = j;
    if (!(impulse == 0.0f)) v.isl_jflag[wk.island] = 1u;
}
PR_D void pulley_solve_position(const DV& d, const pr_revolute_joint& j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    float4 posA = v.s_pos[ia], posB = v.s_pos[ib];
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], rot_from_angle(posA.z));
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], rot_from_angle(posB.z));
    const float ratio = j.lower_angle;
    float lengthA, lengthB;
    const float2 pA = v2(posA.x, posA.y) + rA - v2(j.impulse_y, j.impulse_z);
    const Rot uA = unit_vector_mag(pA, lengthA);
    const float2 pB = v2(posB.x, posB.y) + rB - v2(j.angular_motor_impulse, j.reference_angle);
    const Rot uB = unit_vector_mag(pB, lengthB);
    const float ruA = cross(rA, uA), ruB = cross(rB, uB);
    const float totalInvMassA = imA + iiA * square(ruA) / 1.0f;
    const float totalInvMassB = imB + iiB * square(ruB) / 1.0f;
    const float totalInvMass = totalInvMassA + square(ratio) * totalInvMassB;
    const float mass = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
    const float srcLengthRatio = u2f(j.enable_limit) + ratio * u2f(j.enable_motor);
    const float dstLengthRatio = lengthA + ratio * lengthB;
    const float C = srcLengthRatio - dstLengthRatio;
    const float linearError = fabsf(C);
    const float impulse = -mass * C;
    const float2 PA = -impulse * uA;
    const float2 PB = -ratio * impulse * uB;
    const float2 la = v2(posA.x, posA.y) + imA * PA; posA.x = la.x; posA.y = la.y; posA.z += iiA * cross(rA, PA) / 1.0f;
    const float2 lb = v2(posB.x, posB.y) + imB * PB; posB.x = lb.x; posB.y = lb.y; posB.z += iiB * cross(rB, PB) / 1.0f;
    v.s_pos[ia] = posA; v.s_pos[ib] = posB;
    if (!(linearError < conf.linear_slop)) v.isl_jflag[wk.island] = 1u;
}

// ---- GearJointConf.cpp:171-468: four bodies -- A and B (the row's bodies) geared through their joints to C and D (GearExt). Row overlay:
// impulse = impulse_x, ratio = lower_angle, constant = upper_angle; work: m[0..1] = JvAC, m[2..3] = JvBD, m[4] = JwA, m[5] = JwB, m[6] = JwC,
// m[7] = JwD, m[8] = mass. (A body C or D the island did not reach reads a default BodyConstraint in the reference, :282-285; here it reads the body's own.)
PR_D void gear_init_velocity(const DV& d, unsigned ji, pr_revolute_joint j, JointWork wk, const pr_step_conf& conf, const JView& v)
{
    const GearExt g = d.jt_gear[ji];
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b), ic = jv_idx(v.slotV, g.body_c), id = jv_idx(v.slotV, g.body_d);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b], mC = d.b_mass[g.body_c], mD = d.b_mass[g.body_d];
    const float ratio = j.lower_angle;
    float2 JvAC, JvBD; float JwA, JwB, JwC, JwD;
    float invMass = 0.0f;
    if (g.type_ac == 1u) {
        const Rot qA = rot_from_angle(v.s_pos[jv_idx(v.slotP, j.body_a)].z), qC = rot_from_angle(v.s_pos[jv_idx(v.slotP, g.body_c)].z);
        const Rot u = rot_rotate(mkrot(g.ac[4], g.ac[5]), qC);
        const float2 rC = rotate(v2(g.ac[0], g.ac[1]) - d.b_lc[g.body_c], qC);
        const float2 rA = rotate(v2(g.ac[2], g.ac[3]) - d.b_lc[j.body_a], qA);
        JvAC = 1.0f * u;
        JwC = cross(rC, u); JwA = cross(rA, u);
        const float invRotMassC = mC.y * square(JwC) / 1.0f, invRotMassA = mA.y * square(JwA) / 1.0f;
        const float invLinMass = mC.x + mA.x + invRotMassC + invRotMassA;
        invMass += invLinMass;
    } else {
        JvAC = v2(0.0f, 0.0f); JwA = 1.0f; JwC = 1.0f;
        invMass += mA.y + mC.y;
    }
    if (g.type_bd == 1u) {
        const Rot qB = rot_from_angle(v.s_pos[jv_idx(v.slotP, j.body_b)].z), qD = rot_from_angle(v.s_pos[jv_idx(v.slotP, g.body_d)].z);
        const Rot u = rot_rotate(mkrot(g.bd[4], g.bd[5]), qD);
        const float2 rD = rotate(v2(g.bd[0], g.bd[1]) - d.b_lc[g.body_d], qD);
        const float2 rB = rotate(v2(g.bd[2], g.bd[3]) - d.b_lc[j.body_b], qB);
        JvBD = ratio * u;
        JwD = ratio * cross(rD, u); JwB = ratio * cross(rB, u);
        const float invRotMassD = mD.y * square(JwD) / 1.0f, invRotMassB = mB.y * square(JwB) / 1.0f;
        const float invLinMass = square(ratio) * (mD.x + mB.x) + invRotMassD + invRotMassB;
        invMass += invLinMass;
    } else {
        JvBD = v2(0.0f, 0.0f); JwB = ratio * 1.0f; JwD = ratio * 1.0f;
        invMass += square(ratio) * (mB.y + mD.y);
    }
    wk.m[0] = JvAC.x; wk.m[1] = JvAC.y; wk.m[2] = JvBD.x; wk.m[3] = JvBD.y; wk.m[4] = JwA; wk.m[5] = JwB; wk.m[6] = JwC; wk.m[7] = JwD;
    wk.m[8] = (invMass > 0.0f) ? 1.0f / invMass : 0.0f;
    if (conf.do_warm_start) { // (:244-253: each body's velocity is read, changed and written in turn -- the four need not be different bodies)
        const float imp = j.impulse_x;
        float4 vel = v.s_vel[ia];
        { const float2 l = v2(vel.x, vel.y) + (mA.x * imp) * JvAC; vel.x = l.x; vel.y = l.y; vel.z += mA.y * imp * JwA / 1.0f; }
        float4 velB = v.s_vel[ib], velC = v.s_vel[ic], velD = v.s_vel[id];
        { const float2 l = v2(velB.x, velB.y) + (mB.x * imp) * JvBD; velB.x = l.x; velB.y = l.y; velB.z += mB.y * imp * JwB / 1.0f; }
        { const float2 l = v2(velC.x, velC.y) - (mC.x * imp) * JvAC; velC.x = l.x; velC.y = l.y; velC.z -= mC.y * imp * JwC / 1.0f; }
        { const float2 l = v2(velD.x, velD.y) - (mD.x * imp) * JvBD; velD.x = l.x; velD.y = l.y; velD.z -= mD.y * imp * JwD / 1.0f; }
        v.s_vel[ia] = vel; v.s_vel[ib] = velB; v.s_vel[ic] = velC; v.s_vel[id] = velD;
    } else j.impulse_x = 0.0f;
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}
PR_D void gear_solve_velocity(const DV& d, unsigned ji, pr_revolute_joint j, const JointWork& wk, const JView& v)
{
    const GearExt g = d.jt_gear[ji];
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b), ic = jv_idx(v.slotV, g.body_c), id = jv_idx(v.slotV, g.body_d);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b], mC = d.b_mass[g.body_c], mD = d.b_mass[g.body_d];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib], velC = v.s_vel[ic], velD = v.s_vel[id];
    const float2 JvAC = v2(wk.m[0], wk.m[1]), JvBD = v2(wk.m[2], wk.m[3]);
    const float JwA = wk.m[4], JwB = wk.m[5], JwC = wk.m[6], JwD = wk.m[7];
    const float acDot = dot(JvAC, v2(velA.x, velA.y) - v2(velC.x, velC.y));
    const float bdDot = dot(JvBD, v2(velB.x, velB.y) - v2(velD.x, velD.y));
    const float Cdot = acDot + bdDot + (JwA * velA.z - JwC * velC.z) / 1.0f + (JwB * velB.z - JwD * velD.z) / 1.0f;
    const float impulse = -wk.m[8] * 1.0f * Cdot;
    j.impulse_x += impulse;
    { const float2 l = v2(velA.x, velA.y) + (mA.x * impulse) * JvAC; velA.x = l.x; velA.y = l.y; velA.z += mA.y * impulse * JwA / 1.0f; }
    { const float2 l = v2(velB.x, velB.y) + (mB.x * impulse) * JvBD; velB.x = l.x; velB.y = l.y; velB.z += mB.y * impulse * JwB / 1.0f; }
    { const float2 l = v2(velC.x, velC.y) - (mC.x * impulse) * JvAC; velC.x = l.x; velC.y = l.y; velC.z -= mC.y * impulse * JwC / 1.0f; }
    { const float2 l = v2(velD.x, velD.y) - (mD.x * impulse) * JvBD; velD.x = l.x; velD.y = l.y; velD.z -= mD.y * impulse * JwD / 1.0f; }
    v.s_vel[ia] = velA; v.s_vel[ib] = velB; v.s_vel[ic] = velC; v.s_vel[id] = velD;
    d.jt[ji] = j;
    if (!(impulse == 0.0f)) v.isl_jflag[wk.island] = 1u;
}
PR_D float4 gear_cap(float2 lin, float ang, const pr_step_conf& conf)
{ // Cap(Position, ConstraintSolverConf), d2/Math.cpp:63-71
    const float lsquared = magsq(lin);
    if (lsquared > square(conf.max_linear_correction)) { const float sc = conf.max_linear_correction / sqrtf(lsquared); lin.x *= sc; lin.y *= sc; }
    ang = stdclamp(ang, -conf.max_angular_correction, +conf.max_angular_correction);
    return make_float4(lin.x, lin.y, ang, 0.0f);
}
PR_D void gear_solve_position(const DV& d, unsigned ji, const pr_revolute_joint& j, const JointWork& wk, const pr_step_conf& conf, const JView& v)
{
    const GearExt g = d.jt_gear[ji];
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b), ic = jv_idx(v.slotP, g.body_c), id = jv_idx(v.slotP, g.body_d);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b], mC = d.b_mass[g.body_c], mD = d.b_mass[g.body_d];
    const float4 posA = v.s_pos[ia], posB = v.s_pos[ib], posC = v.s_pos[ic], posD = v.s_pos[id];
    const float ratio = j.lower_angle, constant = j.upper_angle;
    if (((f2u(constant) >> 23) & 0xFFu) == 0xFFu) return; // (!isfinite(constant), :300-302)
    float2 JvAC = v2(0.0f, 0.0f), JvBD = v2(0.0f, 0.0f);
    float JwA = 0.0f, JwB = 0.0f, JwC = 0.0f, JwD = 0.0f, coordinateA = 0.0f, coordinateB = 0.0f, invMass = 0.0f;
    if (g.type_ac == 1u) {
        const Rot qA = rot_from_angle(posA.z), qC = rot_from_angle(posC.z);
        const Rot lax = mkrot(g.ac[4], g.ac[5]);
        const Rot u = rot_rotate(lax, qC);
        const float2 rC = rotate(v2(g.ac[0], g.ac[1]) - d.b_lc[g.body_c], qC);
        const float2 rA = rotate(v2(g.ac[2], g.ac[3]) - d.b_lc[j.body_a], qA);
        JvAC = u * 1.0f;
        JwC = cross(rC, u); JwA = cross(rA, u);
        const float invLinMass = mC.x + mA.x;
        const float invRotMassC = mC.y * square(JwC * 1.0f / 1.0f), invRotMassA = mA.y * square(JwA * 1.0f / 1.0f);
        invMass += invLinMass + invRotMassC + invRotMassA;
        const float2 pC = v2(g.ac[0], g.ac[1]) - d.b_lc[g.body_c];
        const float2 pA = inv_rotate(rA + (v2(posA.x, posA.y) - v2(posC.x, posC.y)), qC);
        coordinateA = dot(pA - pC, lax) / 1.0f;
    } else if (g.type_ac == 2u) {
        JwA = 1.0f; JwC = 1.0f;
        invMass += mA.y + mC.y;
        coordinateA = ((posA.z - posC.z) - g.ac[0]) / 1.0f;
    }
    if (g.type_bd == 1u) {
        const Rot qB = rot_from_angle(posB.z), qD = rot_from_angle(posD.z);
        const Rot lax = mkrot(g.bd[4], g.bd[5]);
        const Rot u = rot_rotate(lax, qD);
        const float2 rD = rotate(v2(g.bd[0], g.bd[1]) - d.b_lc[g.body_d], qD);
        const float2 rB = rotate(v2(g.bd[2], g.bd[3]) - d.b_lc[j.body_b], qB);
        JvBD = ratio * u;
        JwD = ratio * cross(rD, u); JwB = ratio * cross(rB, u);
        const float invLinMass = square(ratio) * (mD.x + mB.x);
        const float invRotMassD = mD.y * square(JwD * 1.0f / 1.0f), invRotMassB = mB.y * square(JwB * 1.0f / 1.0f);
        invMass += invLinMass + invRotMassD + invRotMassB;
        const float2 pD = v2(g.bd[0], g.bd[1]) - d.b_lc[g.body_d];
        const float2 pB = inv_rotate(rB + (v2(posB.x, posB.y) - v2(posD.x, posD.y)), qD);
        coordinateB = dot(pB - pD, lax) / 1.0f;
    } else if (g.type_bd == 2u) {
        JwB = ratio; JwD = ratio;
        invMass += square(ratio) * (mB.y + mD.y);
        coordinateB = ((posB.z - posD.z) - g.bd[0]) / 1.0f;
    }
    const float C = ((coordinateA + ratio * coordinateB) - constant) * conf.reg_resolution_rate;
    const float impulse = ((invMass > 0.0f) ? -C / invMass : 0.0f) * 1.0f * 1.0f;
    const float4 deltaA = gear_cap(mA.x * impulse * JvAC, mA.y * impulse * JwA * 1.0f / 1.0f, conf);
    const float4 deltaC = gear_cap(mC.x * impulse * JvAC, mC.y * impulse * JwC * 1.0f / 1.0f, conf);
    const float4 deltaB = gear_cap(mB.x * impulse * JvBD, mB.y * impulse * JwB * 1.0f / 1.0f, conf);
    const float4 deltaD = gear_cap(mD.x * impulse * JvBD, mD.y * impulse * JwD * 1.0f / 1.0f, conf);
    // (:453-456: SetPosition(posX +/- deltaX) from the positions read at the top, in the order A, B, C, D)
    { float4 p = v.s_pos[ia]; p.x = posA.x + deltaA.x; p.y = posA.y + deltaA.y; p.z = posA.z + deltaA.z; v.s_pos[ia] = p; }
    { float4 p = v.s_pos[ib]; p.x = posB.x + deltaB.x; p.y = posB.y + deltaB.y; p.z = posB.z + deltaB.z; v.s_pos[ib] = p; }
    { float4 p = v.s_pos[ic]; p.x = posC.x - deltaC.x; p.y = posC.y - deltaC.y; p.z = posC.z - deltaC.z; v.s_pos[ic] = p; }
    { float4 p = v.s_pos[id]; p.x = posD.x - deltaD.x; p.y = posD.y - deltaD.y; p.z = posD.z - deltaD.z; v.s_pos[id] = p; }
    float linearError = 0.0f, angularError = 0.0f;
    linearError = stdmax(linearError, magsq(v2(deltaA.x, deltaA.y)));
    linearError = stdmax(linearError, magsq(v2(deltaB.x, deltaB.y)));
    linearError = stdmax(linearError, magsq(v2(deltaC.x, deltaC.y)));
    linearError = stdmax(linearError, magsq(v2(deltaD.x, deltaD.y)));
    angularError = stdmax(angularError, fabsf(deltaA.z));
    angularError = stdmax(angularError, fabsf(deltaB.z));
    angularError = stdmax(angularError, fabsf(deltaC.z));
    angularError = stdmax(angularError, fabsf(deltaD.z));
    if (!((linearError <= square(conf.linear_slop)) && (angularError <= conf.angular_slop))) v.isl_jflag[wk.island] = 1u;
}

PR_D void joint_init_velocity(const DV& d, unsigned ji, const pr_step_conf& conf, const JView& v)
{
    pr_revolute_joint j = d.jt[ji];
    JointWork wk;
    wk.island = joint_island(d, j);
    if (wk.island == INVALID) { d.jt_work[ji] = wk; return; }
    if (joint_type(j) == PR_JOINT_DISTANCE) {
        if (d.b_island[j.body_a] == INVALID) island_count_static(d, wk.island, j.body_a);
        if (d.b_island[j.body_b] == INVALID) island_count_static(d, wk.island, j.body_b);
        distance_init_velocity(d, ji, j, wk, conf, v);
        return;
    }
    if (joint_type(j) == PR_JOINT_ROPE) {
        if (d.b_island[j.body_a] == INVALID) island_count_static(d, wk.island, j.body_a);
        if (d.b_island[j.body_b] == INVALID) island_count_static(d, wk.island, j.body_b);
        rope_init_velocity(d, ji, j, wk, conf, v);
        return;
    }
    if (joint_type(j) == PR_JOINT_FRICTION || joint_type(j) == PR_JOINT_MOTOR) {
        if (d.b_island[j.body_a] == INVALID) island_count_static(d, wk.island, j.body_a);
        if (d.b_island[j.body_b] == INVALID) island_count_static(d, wk.island, j.body_b);
        fm_init_velocity(d, ji, j, wk, conf, v, joint_type(j) == PR_JOINT_MOTOR);
        return;
    }
    if (joint_type(j) == PR_JOINT_TARGET) {
        if (j.body_a != j.body_b && d.b_island[j.body_a] == INVALID) island_count_static(d, wk.island, j.body_a);
        if (d.b_island[j.body_b] == INVALID) island_count_static(d, wk.island, j.body_b);
        target_init_velocity(d, ji, j, wk, conf, v);
        return;
    }
    if (joint_type(j) == PR_JOINT_GEAR) { gear_init_velocity(d, ji, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_WELD || joint_type(j) == PR_JOINT_PRISMATIC || joint_type(j) == PR_JOINT_WHEEL || joint_type(j) == PR_JOINT_PULLEY) {
        if (d.b_island[j.body_a] == INVALID) island_count_static(d, wk.island, j.body_a);
        if (d.b_island[j.body_b] == INVALID) island_count_static(d, wk.island, j.body_b);
        if (joint_type(j) == PR_JOINT_WELD) weld_init_velocity(d, ji, j, wk, conf, v);
        else if (joint_type(j) == PR_JOINT_PRISMATIC) prismatic_init_velocity(d, ji, j, wk, conf, v);
        else if (joint_type(j) == PR_JOINT_PULLEY) pulley_init_velocity(d, ji, j, wk, conf, v);
        else wheel_init_velocity(d, ji, j, wk, conf, v);
        return;
    }
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    // AddJointsToIsland (:1443-1466) puts the other body on the island's stack whatever its type: a body that cannot move is
    // the island's guest (maxIslandBodies, see island_count_static)
    if (d.b_island[j.body_a] == INVALID) island_count_static(d, wk.island, j.body_a);
    if (d.b_island[j.body_b] == INVALID) island_count_static(d, wk.island, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 pA = v.s_pos[jv_idx(v.slotP, j.body_a)], pB = v.s_pos[jv_idx(v.slotP, j.body_b)];
    float4 velA = v.s_vel[ia], velB = v.s_vel[ib];
    const float aA = pA.z, aB = pB.z;
    const Rot qA = rot_from_angle(aA), qB = rot_from_angle(aB);
    const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
    const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
    wk.rAx = rA.x; wk.rAy = rA.y; wk.rBx = rB.x; wk.rBy = rB.y;
    const float totInvI = iiA + iiB;
    const bool fixedRotation = (totInvI == 0.0f);
    {
        const float exx = imA + (square(rA.y) * iiA / 1.0f) + imB + (square(rB.y) * iiB / 1.0f);
        const float eyx = (-rA.y * rA.x * iiA / 1.0f) + (-rB.y * rB.x * iiB / 1.0f);
        const float ezx = (-rA.y * iiA * 1.0f / 1.0f) + (-rB.y * iiB * 1.0f / 1.0f);
        const float eyy = imA + (square(rA.x) * iiA / 1.0f) + imB + (square(rB.x) * iiB / 1.0f);
        const float ezy = (rA.x * iiA * 1.0f / 1.0f) + (rB.x * iiB * 1.0f / 1.0f);
        float* m = wk.m;
        m[0] = exx; m[3] = eyx; m[6] = ezx;
        m[1] = m[3]; m[4] = eyy; m[7] = ezy;
        m[2] = m[6]; m[5] = m[7]; m[8] = totInvI;
    }
    wk.angularMass = (totInvI > 0.0f) ? (1.0f / totInvI) : 0.0f;
    if (!j.enable_motor || fixedRotation) j.angular_motor_impulse = 0.0f;
    if (j.enable_limit && !fixedRotation) {
        const float jointAngle = aB - aA - j.reference_angle;
        if (fabsf(j.upper_angle - j.lower_angle) < (conf.angular_slop * 2)) j.limit_state = LS_EQUAL;
        else if (jointAngle <= j.lower_angle) {
            if (j.limit_state != LS_AT_LOWER) { j.limit_state = LS_AT_LOWER; j.impulse_z = 0; }
        } else if (jointAngle >= j.upper_angle) {
            if (j.limit_state != LS_AT_UPPER) { j.limit_state = LS_AT_UPPER; j.impulse_z = 0; }
        } else { j.limit_state = LS_INACTIVE; j.impulse_z = 0; }
    } else j.limit_state = LS_INACTIVE;
    if (conf.do_warm_start) {
        j.impulse_x *= conf.dt_ratio; j.impulse_y *= conf.dt_ratio; j.impulse_z *= conf.dt_ratio;
        j.angular_motor_impulse *= conf.dt_ratio;
        const float2 P = v2(j.impulse_x * 1.0f, j.impulse_y * 1.0f);
        const float L = j.angular_motor_impulse + (j.impulse_z * 1.0f * 1.0f / (1.0f * 1.0f));
        const float LA = cross(rA, P) / 1.0f + L;
        const float LB = cross(rB, P) / 1.0f + L;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else {
        j.impulse_x = j.impulse_y = j.impulse_z = 0.0f;
        j.angular_motor_impulse = 0.0f;
    }
    v.s_vel[ia] = velA; v.s_vel[ib] = velB; // bodyConstraint.SetVelocity (:224-225)
    d.jt[ji] = j;
    d.jt_work[ji] = wk;
}

PR_D void joint_solve_velocity(const DV& d, unsigned ji, const pr_step_conf& conf, const JView& v)
{
    const JointWork wk = d.jt_work[ji];
    if (wk.island == INVALID) return;
    if (v.isl_state[jv_idx(v.slotI, wk.island)] & ISL_VEL_DONE) return;
    pr_revolute_joint j = d.jt[ji];
    if (joint_type(j) == PR_JOINT_DISTANCE) { distance_solve_velocity(d, ji, j, wk, v); return; }
    if (joint_type(j) == PR_JOINT_ROPE) { rope_solve_velocity(d, ji, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_FRICTION || joint_type(j) == PR_JOINT_MOTOR) { fm_solve_velocity(d, ji, j, wk, conf, v, joint_type(j) == PR_JOINT_MOTOR); return; }
    if (joint_type(j) == PR_JOINT_WELD) { weld_solve_velocity(d, ji, j, wk, v); return; }
    if (joint_type(j) == PR_JOINT_TARGET) { target_solve_velocity(d, ji, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_PRISMATIC) { prismatic_solve_velocity(d, ji, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_WHEEL) { wheel_solve_velocity(d, ji, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_PULLEY) { pulley_solve_velocity(d, ji, j, wk, v); return; }
    if (joint_type(j) == PR_JOINT_GEAR) { gear_solve_velocity(d, ji, j, wk, v); return; }
    const unsigned ia = jv_idx(v.slotV, j.body_a), ib = jv_idx(v.slotV, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    const float4 oldA = v.s_vel[ia], oldB = v.s_vel[ib];
    float4 velA = oldA, velB = oldB;
    const float2 rA = v2(wk.rAx, wk.rAy), rB = v2(wk.rBx, wk.rBy);
    const bool fixedRotation = (iiA + iiB == 0.0f);
    if (j.enable_motor && (j.limit_state != LS_EQUAL) && !fixedRotation) {
        const float impulse = -wk.angularMass * (velB.z - velA.z - j.motor_speed);
        const float oldImpulse = j.angular_motor_impulse;
        const float maxImpulse = conf.delta_time * j.max_motor_torque;
        j.angular_motor_impulse = stdclamp(j.angular_motor_impulse + impulse, -maxImpulse, maxImpulse);
        const float incImpulse = j.angular_motor_impulse - oldImpulse;
        velA.z -= iiA * incImpulse;
        velB.z += iiB * incImpulse;
    }
    const float2 vb = v2(velB.x, velB.y) + rev_perp(rB) * (velB.z / 1.0f);
    const float2 va = v2(velA.x, velA.y) + rev_perp(rA) * (velA.z / 1.0f);
    const float2 vDelta = vb - va;
    if (j.enable_limit && (j.limit_state != LS_INACTIVE) && !fixedRotation) {
        V3 Cdot; Cdot.x = vDelta.x / 1.0f; Cdot.y = vDelta.y / 1.0f; Cdot.z = (velB.z - velA.z) / 1.0f;
        const V3 s = solve33(wk.m, Cdot);
        V3 impulse; impulse.x = -s.x; impulse.y = -s.y; impulse.z = -s.z;
        V3 acc; acc.x = j.impulse_x; acc.y = j.impulse_y; acc.z = j.impulse_z;
        const float newZ = j.impulse_z + impulse.z;
        const bool reduce = (j.limit_state == LS_AT_LOWER && newZ < 0) || (j.limit_state == LS_AT_UPPER && newZ > 0);
        if (reduce) { // UpdateImpulse lambda (:277-285)
            const float2 zcol = v2(wk.m[6], wk.m[7]);
            const float2 rhs = -v2(vDelta.x / 1.0f, vDelta.y / 1.0f) + j.impulse_z * zcol;
            const float2 reduced = solve22(wk.m, rhs);
            impulse.x = reduced.x; impulse.y = reduced.y; impulse.z = -j.impulse_z;
            acc.x = j.impulse_x + reduced.x; acc.y = j.impulse_y + reduced.y; acc.z = 0.0f;
        } else { acc.x += impulse.x; acc.y += impulse.y; acc.z += impulse.z; }
        j.impulse_x = acc.x; j.impulse_y = acc.y; j.impulse_z = acc.z;
        const float2 P = v2(impulse.x * 1.0f, impulse.y * 1.0f);
        const float L = impulse.z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = cross(rA, P) / 1.0f + L;
        const float LB = cross(rB, P) / 1.0f + L;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    } else {
        const float2 impulse = solve22(wk.m, -v2(vDelta.x / 1.0f, vDelta.y / 1.0f));
        j.impulse_x += impulse.x; j.impulse_y += impulse.y;
        const float2 P = v2(impulse.x * 1.0f, impulse.y * 1.0f);
        const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
        const float2 la = v2(velA.x, velA.y) - imA * P; velA.x = la.x; velA.y = la.y; velA.z -= iiA * LA;
        const float2 lb = v2(velB.x, velB.y) + imB * P; velB.x = lb.x; velB.y = lb.y; velB.z += iiB * LB;
    }
    d.jt[ji] = j;
    const bool changed = velA.x != oldA.x || velA.y != oldA.y || velA.z != oldA.z || velB.x != oldB.x || velB.y != oldB.y || velB.z != oldB.z;
    if (changed) {
        v.s_vel[ia] = velA; v.s_vel[ib] = velB;
        v.isl_jflag[wk.island] = 1u; // jointsOkay = false for this island this iteration (:1604-1608)
    }
}

PR_D void joint_solve_position(const DV& d, unsigned ji, const pr_step_conf& conf, const JView& v)
{
    const JointWork wk = d.jt_work[ji];
    if (wk.island == INVALID) return;
    if (v.isl_state[jv_idx(v.slotI, wk.island)] & ISL_POS_SOLVED) return;
    const pr_revolute_joint j = d.jt[ji];
    if (joint_type(j) == PR_JOINT_DISTANCE) { distance_solve_position(d, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_ROPE) { rope_solve_position(d, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_FRICTION || joint_type(j) == PR_JOINT_MOTOR || joint_type(j) == PR_JOINT_TARGET) return; // SolvePosition: nothing to do, always solved
    if (joint_type(j) == PR_JOINT_WELD) { weld_solve_position(d, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_PRISMATIC) { prismatic_solve_position(d, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_WHEEL) { wheel_solve_position(d, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_PULLEY) { pulley_solve_position(d, j, wk, conf, v); return; }
    if (joint_type(j) == PR_JOINT_GEAR) { gear_solve_position(d, ji, j, wk, conf, v); return; }
    const unsigned ia = jv_idx(v.slotP, j.body_a), ib = jv_idx(v.slotP, j.body_b);
    const float4 mA = d.b_mass[j.body_a], mB = d.b_mass[j.body_b];
    const float imA = mA.x, iiA = mA.y, imB = mB.x, iiB = mB.y;
    float4 posA = v.s_pos[ia], posB = v.s_pos[ib];
    const bool fixedRotation = ((iiA + iiB) == 0.0f);
    float angularError = 0.0f;
    if (j.enable_limit && (j.limit_state != LS_INACTIVE) && !fixedRotation) {
        const float angle = posB.z - posA.z - j.reference_angle;
        float limitImpulse = 0.0f;
        if (j.limit_state == LS_AT_LOWER) {
            float C = angle - j.lower_angle;
            angularError = -C;
            C = stdclamp(C + conf.angular_slop, -conf.max_angular_correction, 0.0f);
            limitImpulse = -wk.angularMass * C;
        } else if (j.limit_state == LS_AT_UPPER) {
            float C = angle - j.upper_angle;
            angularError = C;
            C = stdclamp(C - conf.angular_slop, 0.0f, conf.max_angular_correction);
            limitImpulse = -wk.angularMass * C;
        } else {
            const float C = stdclamp(angle - j.lower_angle, -conf.max_angular_correction, conf.max_angular_correction);
            limitImpulse = -wk.angularMass * C;
            angularError = fabsf(C);
        }
        posA.z -= iiA * limitImpulse;
        posB.z += iiB * limitImpulse;
    }
    float positionError = 0.0f;
    {
        const Rot qA = rot_from_angle(posA.z), qB = rot_from_angle(posB.z);
        const float2 rA = rotate(v2(j.local_anchor_a.x, j.local_anchor_a.y) - d.b_lc[j.body_a], qA);
        const float2 rB = rotate(v2(j.local_anchor_b.x, j.local_anchor_b.y) - d.b_lc[j.body_b], qB);
        const float2 C = (v2(posB.x, posB.y) + rB) - (v2(posA.x, posA.y) + rA);
        positionError = magsq(C);
        const float exx = imA + (iiA * square(rA.y) / 1.0f) + imB + (iiB * square(rB.y) / 1.0f);
        const float exy = (-iiA * rA.x * rA.y / 1.0f) + (-iiB * rB.x * rB.y / 1.0f);
        const float eyy = imA + (iiA * square(rA.x) / 1.0f) + imB + (iiB * square(rB.x) / 1.0f);
        const float cp = exx * eyy - exy * exy; // Solve(Matrix22, Vector2), Math.hpp:414-426
        float2 sol = v2(0.0f, 0.0f);
        if (!almost_zero(cp)) {
            const float inverse = 1 / cp;
            sol = v2((eyy * C.x - exy * C.y) * inverse, (exx * C.y - exy * C.x) * inverse);
        }
        const float2 P = -sol;
        const float2 la = v2(posA.x, posA.y) - imA * P; posA.x = la.x; posA.y = la.y; posA.z -= iiA * cross(rA, P) / 1.0f;
        const float2 lb = v2(posB.x, posB.y) + imB * P; posB.x = lb.x; posB.y = lb.y; posB.z += iiB * cross(rB, P) / 1.0f;
    }
    v.s_pos[ia] = posA; v.s_pos[ib] = posB;
    const bool ok = (positionError <= square(conf.linear_slop)) && (angularError <= conf.angular_slop);
    if (!ok) v.isl_jflag[wk.island] = 1u;
}

// phase: 0 = InitVelocity, 1 = SolveVelocity, 2 = SolvePosition. serial != 0: one thread walks all joints.
PR_KERNEL k_joints(DV d, pr_step_conf conf, unsigned nJoints, unsigned phase, unsigned serial)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    const unsigned lo = serial ? 0u : t, hi = serial ? (t == 0u ? nJoints : 0u) : (t < nJoints ? t + 1u : 0u);
    const JView v = jview_global(d);
    for (unsigned ji = lo; ji < hi; ++ji) {
        if (phase == 0u) joint_init_velocity(d, ji, conf, v);
        else if (phase == 1u) joint_solve_velocity(d, ji, conf, v);
        else joint_solve_position(d, ji, conf, v);
    }
}
// one level of joints that share no movable body (pr_world.cu: refresh_joint_tables): one thread per joint of d.jt_order[start, start + count)
PR_KERNEL k_joints_level(DV d, pr_step_conf conf, unsigned start, unsigned count, unsigned phase)
{
    PR_PDL_ENTER();
    const unsigned t = PR_TID;
    if (t >= count) return;
    const unsigned ji = d.jt_order[start + t];
    const JView v = jview_global(d);
    if (phase == 0u) joint_init_velocity(d, ji, conf, v);
    else if (phase == 1u) joint_solve_velocity(d, ji, conf, v);
    else joint_solve_position(d, ji, conf, v);
}
// islands: a joint links its two bodies (AddToIsland :1443-1466) when both can be island members
PR_KERNEL k_isl_union_joints(DV d, unsigned nJoints)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nJoints) {
        const unsigned a = d.jt[i].body_a, b = d.jt[i].body_b;
        if (a != INVALID && d.b_island[a] != INVALID && d.b_island[b] != INVALID) uf_union(d.b_island, a, b);
    }
}
// Add(JointID) wakes both bodies (AabbTreeWorld.cpp:1159-1180)
PR_KERNEL k_set_joints(DV d, unsigned n, const unsigned* ids, const pr_revolute_joint* rows)
{ // SetJoint: Remove(id) wakes both bodies (:1192-1205), joint = def (:1128)
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) { // (the OLD row's bodies are the ones Remove(id) wakes: the same bodies unless the new value names others)
        const pr_revolute_joint old = d.jt[ids[i]];
        body_set_awake(d, old.body_a); body_set_awake(d, old.body_b);
        d.jt[ids[i]] = rows[i];
    }
}
// Add(JointID, flagForFiltering) / Remove(JointID) of a joint with collideConnected == false (AabbTreeWorld.cpp:1159-1191, FlagForFiltering
// :436-451): the contacts between its two bodies are flagged for the next Step's filter test (pr_mutate.cuh: k_destroy_filter)
PR_KERNEL k_pair_flag_filter(DV d, unsigned nOrder, unsigned n, const unsigned* pairs)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nOrder) {
        const unsigned id = d.c_order[i];
        const uint4 ab = d.c_ab[id];
        for (unsigned k = 0; k < n; ++k) {
            const unsigned a = pairs[2u * k], b = pairs[2u * k + 1u];
            if ((ab.x == a && ab.y == b) || (ab.x == b && ab.y == a)) { atomicOr(&d.c_flags[id], static_cast<unsigned>(PR_CONTACT_FILTER)); break; }
        }
    }
}
// ShiftOrigin(Joint&, newOrigin) (AabbTreeWorld.cpp:2237-2240): the joints that hold world points move them -- the target joint's target
// (TargetJointConf.hpp:200-204) and the pulley joint's ground anchors (PulleyJointConf.cpp:284-289)
PR_KERNEL k_shift_joints(DV d, unsigned nJoints, float x, float y)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < nJoints) {
        pr_revolute_joint j = d.jt[i];
        if (j.body_a == INVALID && j.body_b == INVALID) return;
        if (joint_type(j) == PR_JOINT_TARGET) { j.local_anchor_a.x -= x; j.local_anchor_a.y -= y; d.jt[i] = j; }
        else if (joint_type(j) == PR_JOINT_PULLEY) { j.impulse_y -= x; j.impulse_z -= y; j.angular_motor_impulse -= x; j.reference_angle -= y; d.jt[i] = j; }
    }
}
PR_KERNEL k_joint_wake(DV d, unsigned first, unsigned n)
{
    PR_PDL_ENTER();
    const unsigned i = PR_TID;
    if (i < n) { body_set_awake(d, d.jt[first + i].body_a); body_set_awake(d, d.jt[first + i].body_b); }
}

} // namespace pr
#endif
