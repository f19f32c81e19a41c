// orc_toi.cpp -- TEST INFRASTRUCTURE ONLY (oracle restatement): GJK Distance (Distance.cpp, Simplex.cpp),
// SeparationScenario.cpp, GetToiViaSat (TimeOfImpact.cpp) and the revolute joint (RevoluteJointConf.cpp).
#include <algorithm>
#include <cstring>
#include "orc_toi.h"
namespace orc {
namespace {

constexpr uint32_t kInvalidVertex = 255u; // InvalidVertex, Settings.hpp:87-90
constexpr float kInfF = std::numeric_limits<float>::infinity();

// SimplexEdge, SimplexEdge.hpp: world points of one vertex of A and one of B, with their indices
struct SEdge { V2 wA; uint32_t iA; V2 wB; uint32_t iB; };
inline V2 delta(const SEdge& e) { return e.wB - e.wA; } // GetPointDelta, SimplexEdge.hpp:84-87
struct SEdges { SEdge e[3]; uint32_t n = 0; };
struct Simplex { SEdges edges; float coef[3] = {0, 0, 0}; };
struct IndexPair { uint32_t a = kInvalidVertex, b = kInvalidVertex; };
inline bool operator==(IndexPair x, IndexPair y) { return x.a == y.a && x.b == y.b; }
struct IndexPair3 { IndexPair p[3]; };
inline uint32_t num_valid(const IndexPair3& ip)
{ // GetNumValidIndices, IndexPair.hpp:66-72
    const IndexPair inv;
    return (ip.p[0] == inv ? 0u : 1u) + (ip.p[1] == inv ? 0u : 1u) + (ip.p[2] == inv ? 0u : 1u);
}
struct Cache { float metric = 0.0f; IndexPair3 indices; };

// GetSimplexEdge, Distance.cpp:37-44
inline SEdge simplex_edge(const Proxy& a, Xf xfA, uint32_t iA, const Proxy& b, Xf xfB, uint32_t iB)
{
    return {transform(a.vertex(iA), xfA), iA, transform(b.vertex(iB), xfB), iB};
}
// GetIndexPairs, Simplex.cpp:27-42
inline IndexPair3 index_pairs(const SEdges& s)
{
    IndexPair3 r;
    for (uint32_t i = 0; i < s.n; ++i) r.p[i] = {s.e[i].iA, s.e[i].iB};
    return r;
}
// Simplex::CalcMetric, Simplex.cpp:218-238
inline float calc_metric(const SEdges& s)
{
    if (s.n == 2) return std::sqrt(magsq(delta(s.e[1]) - delta(s.e[0])));
    if (s.n == 3) return cross(delta(s.e[1]) - delta(s.e[0]), delta(s.e[2]) - delta(s.e[0]));
    return 0.0f;
}
// CalcSearchDirection, Simplex.cpp:44-62
inline V2 search_direction(const SEdges& s)
{
    if (s.n == 1) return -delta(s.e[0]);
    if (s.n == 2) {
        const V2 e12 = delta(s.e[1]) - delta(s.e[0]);
        const V2 e0 = delta(s.e[0]);
        const float sgn = cross(e12, -e0);
        return (sgn > 0.0f) ? V2{-e12.y, e12.x} : V2{e12.y, -e12.x};
    }
    return {0.0f, 0.0f};
}
// Simplex::Get(s0, s1), Simplex.cpp:69-116
inline Simplex simplex2(const SEdge& s0, const SEdge& s1)
{
    Simplex r;
    const V2 w1 = delta(s0), w2 = delta(s1);
    const V2 e12 = w2 - w1;
    const float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) { r.edges.e[0] = s0; r.edges.n = 1; r.coef[0] = 1.0f; return r; }
    const float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) { r.edges.e[0] = s1; r.edges.n = 1; r.coef[0] = 1.0f; return r; }
    const float inv_sum = 1.0f / (d12_1 + d12_2);
    r.edges.e[0] = s0; r.edges.e[1] = s1; r.edges.n = 2; r.coef[0] = d12_1 * inv_sum; r.coef[1] = d12_2 * inv_sum;
    return r;
}
// Simplex::Get(s0, s1, s2), Simplex.cpp:118-199
inline Simplex simplex3(const SEdge& s0, const SEdge& s1, const SEdge& s2)
{
    Simplex r;
    auto one = [&](const SEdge& s) { r.edges.e[0] = s; r.edges.n = 1; r.coef[0] = 1.0f; return r; };
    auto two = [&](const SEdge& a, const SEdge& b, float ca, float cb) {
        r.edges.e[0] = a; r.edges.e[1] = b; r.edges.n = 2; r.coef[0] = ca; r.coef[1] = cb; return r;
    };
    const V2 w1 = delta(s0), w2 = delta(s1), w3 = delta(s2);
    const V2 e12 = w2 - w1;
    const float d12_1 = dot(w2, e12), d12_2 = -dot(w1, e12);
    const V2 e13 = w3 - w1;
    const float d13_1 = dot(w3, e13), d13_2 = -dot(w1, e13);
    const V2 e23 = w3 - w2;
    const float d23_1 = dot(w3, e23), d23_2 = -dot(w2, e23);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) return one(s0);
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) return one(s1);
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) return one(s2);
    const float n123 = cross(e12, e13);
    const float d123_3 = n123 * cross(w1, w2);
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
        const float inv_sum = 1.0f / (d12_1 + d12_2);
        return two(s0, s1, d12_1 * inv_sum, d12_2 * inv_sum);
    }
    const float d123_2 = n123 * cross(w3, w1);
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
        const float inv_sum = 1.0f / (d13_1 + d13_2);
        return two(s0, s2, d13_1 * inv_sum, d13_2 * inv_sum);
    }
    const float d123_1 = n123 * cross(w2, w3);
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
        const float inv_sum = 1.0f / (d23_1 + d23_2);
        return two(s2, s1, d23_2 * inv_sum, d23_1 * inv_sum); // note the swapped order, Simplex.cpp:193
    }
    const float inv_sum = 1.0f / (d123_1 + d123_2 + d123_3);
    r.edges.e[0] = s0; r.edges.e[1] = s1; r.edges.e[2] = s2; r.edges.n = 3;
    r.coef[0] = d123_1 * inv_sum; r.coef[1] = d123_2 * inv_sum; r.coef[2] = d123_3 * inv_sum;
    return r;
}
inline Simplex simplex_get(const SEdges& s)
{ // Simplex::Get(edges), Simplex.cpp:201-216
    if (s.n == 1) { Simplex r; r.edges = s; r.coef[0] = 1.0f; return r; }
    if (s.n == 2) return simplex2(s.e[0], s.e[1]);
    if (s.n == 3) return simplex3(s.e[0], s.e[1], s.e[2]);
    return Simplex{};
}
// GetSupportIndex, DistanceProxy.hpp:232-249 (first strict maximum of Dot(vertex, dir))
inline uint32_t support_index(const Proxy& p, V2 dir)
{
    uint32_t index = kInvalidVertex;
    float maxValue = -kInfF;
    for (uint32_t i = 0; i < p.count; ++i) {
        const float value = dot(p.vertex(i), dir);
        if (maxValue < value) { maxValue = value; index = i; }
    }
    return index;
}
enum DistState { D_Unknown, D_UnfitSearchDir, D_HitMaxIters, D_MaxPoints, D_DuplicateIndexPair };
struct DistOut { Simplex simplex; uint32_t iterations; int state; };

// Distance, Distance.cpp:100-203
DistOut distance(const Proxy& a, Xf xfA, const Proxy& b, Xf xfB, const Cache& cache, uint32_t maxIters)
{
    IndexPair3 saved = cache.indices;
    SEdges edges;
    const uint32_t count = num_valid(saved);
    for (uint32_t i = 0; i < count; ++i) edges.e[i] = simplex_edge(a, xfA, saved.p[i].a, b, xfB, saved.p[i].b);
    edges.n = count;
    if (edges.n > 1) {
        const float metric1 = cache.metric;
        const float metric2 = calc_metric(edges);
        if ((metric2 < (metric1 / 2)) || (metric2 > (metric1 * 2)) || (metric2 < 0) || almost_zero(metric2)) edges.n = 0;
    }
    if (edges.n == 0) {
        edges.e[0] = simplex_edge(a, xfA, 0, b, xfB, 0); edges.n = 1;
        saved = IndexPair3{}; saved.p[0] = {0, 0};
    }
    Simplex simplex;
    int state = D_HitMaxIters;
    uint32_t iter = 0;
    while (iter < maxIters) {
        ++iter;
        simplex = simplex_get(edges);
        edges = simplex.edges;
        if (edges.n == 3) { state = D_MaxPoints; break; }
        const V2 d = search_direction(edges);
        if (almost_zero(magsq(d))) { state = D_UnfitSearchDir; break; }
        const uint32_t indexA = support_index(a, inv_rotate(-d, xfA.q));
        const uint32_t indexB = support_index(b, inv_rotate(d, xfB.q));
        const IndexPair key{indexA, indexB};
        if (saved.p[0] == key || saved.p[1] == key || saved.p[2] == key) { state = D_DuplicateIndexPair; break; }
        edges.e[edges.n++] = simplex_edge(a, xfA, indexA, b, xfB, indexB);
        saved = index_pairs(edges);
    }
    return {simplex, iter, state};
}
// GetWitnessPoints, Distance.cpp:72-94, and |pB - pA|^2
inline float witness_dist_sq(const Simplex& s)
{
    V2 pA{0, 0}, pB{0, 0};
    for (uint32_t i = 0; i < s.edges.n; ++i) {
        pA = pA + s.edges.e[i].wA * s.coef[i];
        pB = pB + s.edges.e[i].wB * s.coef[i];
    }
    return magsq(pB - pA);
}

// SeparationScenario.cpp: type 0 = points, 1 = faceA, 2 = faceB
struct Scenario { int type; Rot axis; V2 localPoint; };
// GetSeparationScenario, SeparationScenario.cpp:138-190
Scenario separation_scenario(const IndexPair3& idx, const Proxy& a, Xf xfA, const Proxy& b, Xf xfB)
{
    const Rot zero{0.0f, 0.0f};
    const uint32_t n = num_valid(idx);
    if (n >= 2 && idx.p[0].a == idx.p[1].a) { // FaceB: two points on B, one on A
        const V2 b1 = b.vertex(idx.p[0].b), b2 = b.vertex(idx.p[1].b);
        const V2 e = b2 - b1;
        const Rot axis = unit_vector(V2{e.y, -e.x}, zero);
        const Rot normal = rot_rotate(axis, xfB.q);
        const V2 localPoint = (b1 + b2) / 2.0f;
        const V2 pointB = transform(localPoint, xfB);
        const V2 pointA = transform(a.vertex(idx.p[0].a), xfA);
        const V2 dp = pointA - pointB;
        return {2, (dot(dp, normal) < 0.0f) ? -axis : axis, localPoint};
    }
    if (n >= 2) { // FaceA
        const V2 a1 = a.vertex(idx.p[0].a), a2 = a.vertex(idx.p[1].a);
        const V2 e = a2 - a1;
        const Rot axis = unit_vector(V2{e.y, -e.x}, zero);
        const Rot normal = rot_rotate(axis, xfA.q);
        const V2 localPoint = (a1 + a2) / 2.0f;
        const V2 pointA = transform(localPoint, xfA);
        const V2 pointB = transform(b.vertex(idx.p[0].b), xfB);
        const V2 dp = pointB - pointA;
        return {1, (dot(dp, normal) < 0.0f) ? -axis : axis, localPoint};
    }
    const V2 pointA = transform(a.vertex(idx.p[0].a), xfA);
    const V2 pointB = transform(b.vertex(idx.p[0].b), xfB);
    return {0, unit_vector(pointB - pointA, zero), {0, 0}};
}
struct MinSep { float distance; IndexPair ip; };
// FindMinSeparation, SeparationScenario.cpp:29-74
MinSep find_min_separation(const Scenario& s, const Proxy& a, Xf xfA, const Proxy& b, Xf xfB)
{
    if (s.type == 0) {
        const V2 ax = rotv(s.axis);
        const uint32_t iA = support_index(a, inv_rotate(ax, xfA.q));
        const uint32_t iB = support_index(b, inv_rotate(-ax, xfB.q));
        const V2 pA = transform(a.vertex(iA), xfA), pB = transform(b.vertex(iB), xfB);
        return {dot(pB - pA, s.axis), {iA, iB}};
    }
    if (s.type == 1) {
        const Rot normal = rot_rotate(s.axis, xfA.q);
        const V2 pA = transform(s.localPoint, xfA);
        const uint32_t iB = support_index(b, inv_rotate(-rotv(normal), xfB.q));
        const V2 pB = transform(b.vertex(iB), xfB);
        return {dot(pB - pA, normal), {kInvalidVertex, iB}};
    }
    const Rot normal = rot_rotate(s.axis, xfB.q);
    const uint32_t iA = support_index(a, inv_rotate(-rotv(normal), xfA.q));
    const V2 pA = transform(a.vertex(iA), xfA);
    const V2 pB = transform(s.localPoint, xfB);
    return {dot(pA - pB, normal), {iA, kInvalidVertex}};
}
// Evaluate, SeparationScenario.cpp:76-114
float evaluate(const Scenario& s, const Proxy& a, Xf xfA, const Proxy& b, Xf xfB, IndexPair ip)
{
    if (s.type == 0) {
        const V2 pA = transform(a.vertex(ip.a), xfA), pB = transform(b.vertex(ip.b), xfB);
        return dot(pB - pA, s.axis);
    }
    if (s.type == 1) {
        const Rot normal = rot_rotate(s.axis, xfA.q);
        const V2 pA = transform(s.localPoint, xfA), pB = transform(b.vertex(ip.b), xfB);
        return dot(pB - pA, normal);
    }
    const Rot normal = rot_rotate(s.axis, xfB.q);
    const V2 pB = transform(s.localPoint, xfB), pA = transform(a.vertex(ip.a), xfA);
    return dot(pA - pB, normal);
}
// GetTransformation(sweep, beta), d2/Math.hpp:251-254
inline Xf sweep_xf(const Sweep& s, float beta) { return get_xf(lerp_pos(s.pos0, s.pos1, beta), s.lc); }

} // namespace

// TestOverlap, Distance.cpp:205-218 (default-constructed simplex cache)
float test_overlap(const Proxy& a, Xf xfA, const Proxy& b, Xf xfB, uint32_t maxIters)
{
    const DistOut d = distance(a, xfA, b, xfB, Cache{}, maxIters);
    const float distSq = witness_dist_sq(d.simplex);
    const float totalRadiusSq = square(a.radius + b.radius);
    return totalRadiusSq - distSq;
}

// GetToiViaSat, TimeOfImpact.cpp:35-251. States follow ToiOutput::State (ToiOutput.hpp:73-140).
ToiOut toi_via_sat(const Proxy& a, const Sweep& sA, const Proxy& b, const Sweep& sB, const pr_step_conf& conf)
{
    enum { e_unknown, e_touching, e_separated, e_overlapped, e_maxRootIters, e_nextAfter, e_maxToiIters,
           e_belowMinTarget, e_maxDistIters, e_targetDepthExceedsTotalRadius, e_minTargetSquaredOverflow,
           e_maxTargetSquaredOverflow, e_notFinite };
    ToiOut out{0.0f, e_unknown, 0, 0, 0};
    const float timeMax = 1.0f; // GetToiConf, ToiConf.cpp:26-35
    const float totalRadius = a.radius + b.radius;
    if (conf.target_depth > totalRadius) { out.state = e_targetDepthExceedsTotalRadius; return out; }
    const float target = totalRadius - conf.target_depth;
    const float maxTarget = std::max(target + conf.tolerance, 0.0f);
    const float minTarget = std::max(target - conf.tolerance, 0.0f);
    const float minTargetSquared = square(minTarget);
    if (!std::isfinite(minTargetSquared) && std::isfinite(minTarget)) { out.state = e_minTargetSquaredOverflow; return out; }
    const float maxTargetSquared = square(maxTarget);
    if (!std::isfinite(maxTargetSquared) && std::isfinite(maxTarget)) { out.state = e_maxTargetSquaredOverflow; return out; }
    float timeLo = 0.0f;
    Xf loA = sweep_xf(sA, timeLo), loB = sweep_xf(sB, timeLo);
    Cache cache;
    auto done = [&](float t, int st) { out.time = t; out.state = st; return out; };
    while (out.toi_iters < conf.max_toi_iters) {
        const DistOut dinfo = distance(a, loA, b, loB, cache, conf.max_distance_iters);
        ++out.toi_iters;
        out.max_dist_iters = std::max(out.max_dist_iters, dinfo.iterations);
        if (dinfo.state == D_HitMaxIters) return done(timeLo, e_maxDistIters);
        cache.metric = calc_metric(dinfo.simplex.edges);
        cache.indices = index_pairs(dinfo.simplex.edges);
        const float distSquared = witness_dist_sq(dinfo.simplex);
        if (distSquared < minTargetSquared) return done(timeLo, e_overlapped);
        if (distSquared <= maxTargetSquared) return done(timeLo, e_touching);
        const Scenario fcn = separation_scenario(cache.indices, a, loA, b, loB);
        float timeHi = timeMax;
        Xf hiA = sweep_xf(sA, timeHi), hiB = sweep_xf(sB, timeHi);
        for (uint32_t pbIter = 0; pbIter < 254u; ++pbIter) { // MaxShapeVertices, Settings.hpp
            const MinSep hiMin = find_min_separation(fcn, a, hiA, b, hiB);
            if (hiMin.distance > maxTarget) return done(timeHi, e_separated);
            if (hiMin.distance >= minTarget) {
                if (timeHi == timeLo) return done(timeHi, e_separated);
                timeLo = timeHi; loA = hiA; loB = hiB;
                break;
            }
            const float loEval = evaluate(fcn, a, loA, b, loB, hiMin.ip);
            if (loEval <= maxTarget) return done(timeLo, (loEval < minTarget) ? e_belowMinTarget : e_touching);
            float a1 = timeLo, a2 = timeHi, s1 = loEval, s2 = hiMin.distance;
            uint32_t roots = 0;
            float t = timeLo;
            for (;;) {
                if (roots == conf.max_toi_root_iters) {
                    out.max_root_iters = std::max(out.max_root_iters, roots);
                    return done(t, e_maxRootIters);
                }
                if (std::nextafter(a1, a2) >= a2) {
                    out.max_root_iters = std::max(out.max_root_iters, roots);
                    return done(t, e_nextAfter);
                }
                // Secant / Bisect, Math.hpp:92-108
                t = (roots % 2u) ? (a1 + (target - s1) * (a2 - a1) / (s2 - s1)) : ((a1 + a2) / 2);
                ++roots;
                const Xf tA = sweep_xf(sA, t), tB = sweep_xf(sB, t);
                const float s = evaluate(fcn, a, tA, b, tB, hiMin.ip);
                if (std::fabs(s - target) <= conf.tolerance) { timeHi = t; hiA = tA; hiB = tB; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
            }
            out.max_root_iters = std::max(out.max_root_iters, roots);
        }
    }
    return done(timeLo, e_maxToiIters);
}

namespace {
// Mat33 is three column vectors ex, ey, ez (GetX(GetY(m)) == ey.x); stored here as m[3*col + row]
inline float m33(const float* m, int col, int row) { return m[3 * col + row]; }
struct V3 { float x, y, z; };
inline float dot3(V3 a, V3 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; } // Math.hpp:320
inline V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; } // Math.hpp:397
inline V3 col(const float* m, int c) { return {m[3 * c], m[3 * c + 1], m[3 * c + 2]}; }
// Solve33, Math.hpp:447-456
inline V3 solve33(const float* m, V3 b)
{
    const float dp = dot3(col(m, 0), cross3(col(m, 1), col(m, 2)));
    const float det = (dp != 0) ? (1 / dp) : dp;
    return {det * dot3(b, cross3(col(m, 1), col(m, 2))), det * dot3(col(m, 0), cross3(b, col(m, 2))),
            det * dot3(col(m, 0), cross3(col(m, 1), b))};
}
// Solve22, Math.hpp:462-474
inline V2 solve22(const float* m, V2 b)
{
    const float matXX = m33(m, 0, 0), matXY = m33(m, 1, 0), matYX = m33(m, 0, 1), matYY = m33(m, 1, 1);
    const float cp = matXX * matYY - matXY * matYX;
    const float det = (cp != 0) ? (1 / cp) : cp;
    return {det * (matYY * b.x - matXY * b.y), det * (matXX * b.y - matYX * b.x)};
}
enum { LS_INACTIVE = 0, LS_AT_LOWER = 1, LS_AT_UPPER = 2, LS_EQUAL = 3 }; // LimitState.hpp:32-45
} // namespace

// InitVelocity(RevoluteJointConf&, ...), RevoluteJointConf.cpp:127-226
void revolute_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                            float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const float aA = posA.a, aB = posB.a;
    const Rot qA = rot_from_angle(aA), qB = rot_from_angle(aB);
    wk.rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    wk.rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const float totInvI = iiA + iiB;
    const bool fixedRotation = (totInvI == 0.0f);
    { // GetMat33, RevoluteJointConf.cpp:35-62
        const V2 rA = wk.rA, rB = wk.rB;
        const float exx = imA + (square(rA.y) * iiA / 1.0f) + imB + (square(rB.y) * iiB / 1.0f);
        const float eyx = (-rA.y * rA.x * iiA / 1.0f) + (-rB.y * rB.x * iiB / 1.0f);
        const float ezx = (-rA.y * iiA * 1.0f / 1.0f) + (-rB.y * iiB * 1.0f / 1.0f);
        const float eyy = imA + (square(rA.x) * iiA / 1.0f) + imB + (square(rB.x) * iiB / 1.0f);
        const float ezy = (rA.x * iiA * 1.0f / 1.0f) + (rB.x * iiB * 1.0f / 1.0f);
        float* m = wk.mass;
        m[0] = exx; m[3] = eyx; m[6] = ezx;   // GetX(GetX), GetX(GetY), GetX(GetZ)
        m[1] = m[3]; m[4] = eyy; m[7] = ezy;  // GetY(GetX) = GetX(GetY), GetY(GetY), GetY(GetZ)
        m[2] = m[6]; m[5] = m[7]; m[8] = totInvI;
    }
    wk.angularMass = (totInvI > 0.0f) ? (1.0f / totInvI) : 0.0f;
    if (!j.enable_motor || fixedRotation) j.angular_motor_impulse = 0.0f;
    if (j.enable_limit && !fixedRotation) {
        const float jointAngle = aB - aA - j.reference_angle;
        if (std::fabs(j.upper_angle - j.lower_angle) < (c.angularSlop * 2)) j.limit_state = LS_EQUAL;
        else if (jointAngle <= j.lower_angle) {
            if (j.limit_state != LS_AT_LOWER) { j.limit_state = LS_AT_LOWER; j.impulse_z = 0; }
        } else if (jointAngle >= j.upper_angle) {
            if (j.limit_state != LS_AT_UPPER) { j.limit_state = LS_AT_UPPER; j.impulse_z = 0; }
        } else { j.limit_state = LS_INACTIVE; j.impulse_z = 0; }
    } else j.limit_state = LS_INACTIVE;
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio; j.impulse_y *= c.dtRatio; j.impulse_z *= c.dtRatio;
        j.angular_motor_impulse *= c.dtRatio;
        const V2 P{j.impulse_x * 1.0f, j.impulse_y * 1.0f};
        const float L = j.angular_motor_impulse + (j.impulse_z * 1.0f * 1.0f / (1.0f * 1.0f));
        const float LA = cross(wk.rA, P) / 1.0f + L;
        const float LB = cross(wk.rB, P) / 1.0f + L;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else {
        j.impulse_x = j.impulse_y = j.impulse_z = 0.0f;
        j.angular_motor_impulse = 0.0f;
    }
}

// SolveVelocity(RevoluteJointConf&, ...), RevoluteJointConf.cpp:228-331; returns "no change"
bool revolute_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& vA, float imB, float iiB,
                             Vel& vB, const JointSolverConf& c)
{
    const Vel oldA = vA, oldB = vB;
    Vel velA = vA, velB = vB;
    const bool fixedRotation = (iiA + iiB == 0.0f);
    if (j.enable_motor && (j.limit_state != LS_EQUAL) && !fixedRotation) {
        const float impulse = -wk.angularMass * (velB.a - velA.a - j.motor_speed);
        const float oldImpulse = j.angular_motor_impulse;
        const float maxImpulse = c.dt * j.max_motor_torque;
        j.angular_motor_impulse = std::clamp(j.angular_motor_impulse + impulse, -maxImpulse, maxImpulse);
        const float incImpulse = j.angular_motor_impulse - oldImpulse;
        velA.a -= iiA * incImpulse;
        velB.a += iiB * incImpulse;
    }
    const V2 vb = velB.l + rev_perp(wk.rB) * (velB.a / 1.0f);
    const V2 va = velA.l + rev_perp(wk.rA) * (velA.a / 1.0f);
    const V2 vDelta = vb - va;
    if (j.enable_limit && (j.limit_state != LS_INACTIVE) && !fixedRotation) {
        const V3 Cdot{vDelta.x / 1.0f, vDelta.y / 1.0f, (velB.a - velA.a) / 1.0f};
        const V3 s = solve33(wk.mass, Cdot);
        V3 impulse{-s.x, -s.y, -s.z};
        auto updateImpulse = [&](V3& imp) -> V3 {
            const V2 zcol{m33(wk.mass, 2, 0), m33(wk.mass, 2, 1)};
            const V2 rhs = -V2{vDelta.x / 1.0f, vDelta.y / 1.0f} + j.impulse_z * zcol;
            const V2 reduced = solve22(wk.mass, rhs);
            imp = {reduced.x, reduced.y, -j.impulse_z};
            return {j.impulse_x + reduced.x, j.impulse_y + reduced.y, 0.0f};
        };
        V3 acc{j.impulse_x, j.impulse_y, j.impulse_z};
        if (j.limit_state == LS_AT_LOWER) {
            const float newZ = j.impulse_z + impulse.z;
            acc = (newZ < 0) ? updateImpulse(impulse) : V3{acc.x + impulse.x, acc.y + impulse.y, acc.z + impulse.z};
        } else if (j.limit_state == LS_AT_UPPER) {
            const float newZ = j.impulse_z + impulse.z;
            acc = (newZ > 0) ? updateImpulse(impulse) : V3{acc.x + impulse.x, acc.y + impulse.y, acc.z + impulse.z};
        } else acc = {acc.x + impulse.x, acc.y + impulse.y, acc.z + impulse.z};
        j.impulse_x = acc.x; j.impulse_y = acc.y; j.impulse_z = acc.z;
        const V2 P{impulse.x * 1.0f, impulse.y * 1.0f};
        const float L = impulse.z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = cross(wk.rA, P) / 1.0f + L;
        const float LB = cross(wk.rB, P) / 1.0f + L;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else {
        const V2 impulse = solve22(wk.mass, -V2{vDelta.x / 1.0f, vDelta.y / 1.0f});
        j.impulse_x += impulse.x; j.impulse_y += impulse.y;
        const V2 P{impulse.x * 1.0f, impulse.y * 1.0f};
        const float LA = cross(wk.rA, P) / 1.0f, LB = cross(wk.rB, P) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    }
    const bool changed = (velA.l != oldA.l) || (velA.a != oldA.a) || (velB.l != oldB.l) || (velB.a != oldB.a);
    if (changed) { vA = velA; vB = velB; return false; }
    return true;
}

// SolvePosition(const RevoluteJointConf&, ...), RevoluteJointConf.cpp:333-435
bool revolute_solve_position(const pr_revolute_joint& j, float imA, float iiA, V2 lcA, Pos& pA, float imB, float iiB, V2 lcB,
                             Pos& pB, const JointSolverConf& c)
{
    Pos posA = pA, posB = pB;
    const bool fixedRotation = ((iiA + iiB) == 0.0f);
    float angularError = 0.0f;
    if (j.enable_limit && (j.limit_state != LS_INACTIVE) && !fixedRotation) {
        const float totInvI = iiA + iiB;
        const float angularMass = (totInvI > 0.0f) ? (1.0f / totInvI) : 0.0f; // object.angularMass as set by InitVelocity
        const float angle = posB.a - posA.a - j.reference_angle;
        float limitImpulse = 0.0f;
        if (j.limit_state == LS_AT_LOWER) {
            float C = angle - j.lower_angle;
            angularError = -C;
            C = std::clamp(C + c.angularSlop, -c.maxAngularCorrection, 0.0f);
            limitImpulse = -angularMass * C;
        } else if (j.limit_state == LS_AT_UPPER) {
            float C = angle - j.upper_angle;
            angularError = C;
            C = std::clamp(C - c.angularSlop, 0.0f, c.maxAngularCorrection);
            limitImpulse = -angularMass * C;
        } else {
            const float C = std::clamp(angle - j.lower_angle, -c.maxAngularCorrection, c.maxAngularCorrection);
            limitImpulse = -angularMass * C;
            angularError = std::fabs(C);
        }
        posA.a -= iiA * limitImpulse;
        posB.a += iiB * limitImpulse;
    }
    float positionError = 0.0f;
    {
        const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
        const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
        const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
        const V2 C = (posB.l + rB) - (posA.l + rA);
        positionError = magsq(C);
        const float exx = imA + (iiA * square(rA.y) / 1.0f) + imB + (iiB * square(rB.y) / 1.0f);
        const float exy = (-iiA * rA.x * rA.y / 1.0f) + (-iiB * rB.x * rB.y / 1.0f);
        const float eyy = imA + (iiA * square(rA.x) / 1.0f) + imB + (iiB * square(rB.x) / 1.0f);
        // Solve(Matrix22 K, C), Math.hpp:414-426, with K = {{exx, exy}, {exy, eyy}}
        const float cp = exx * eyy - exy * exy; // Cross(get<0>(mat), get<1>(mat))
        V2 sol{0.0f, 0.0f};
        if (!almost_zero(cp)) {
            const float inverse = 1 / cp;
            sol = {(eyy * C.x - exy * C.y) * inverse, (exx * C.y - exy * C.x) * inverse};
        }
        const V2 P = -sol;
        posA.l = posA.l - imA * P; posA.a -= iiA * cross(rA, P) / 1.0f;
        posB.l = posB.l + imB * P; posB.a += iiB * cross(rB, P) / 1.0f;
    }
    pA = posA; pB = posB;
    return (positionError <= square(c.linearSlop)) && (angularError <= c.angularSlop);
}

// ---------------- DistanceJointConf.cpp:81-246 ----------------
namespace {
// UnitVec::Get(x, y) with the magnitude, UnitVec.hpp:176-232 (default fallback: the zero vector)
Rot unit_vector_mag(V2 v, float& mag)
{
    const float x = v.x, y = v.y;
    const float ms = x * x + y * y;
    if (std::isnormal(ms)) { mag = std::sqrt(ms); const float inv = 1.0f / mag; return Rot{x * inv, y * inv}; }
    mag = std::hypot(x, y);
    if (!std::isnormal(mag)) { mag = 0.0f; return Rot{0.0f, 0.0f}; }
    return Rot{x / mag, y / mag};
}
void distance_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                            float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    wk.rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    wk.rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const V2 delta = (posB.l + wk.rB) - (posA.l + wk.rA);
    float length;
    const Rot u = unit_vector_mag(delta, length);
    const V2 uv{u.c, u.s};
    const float crAu = cross(wk.rA, uv) / 1.0f, crBu = cross(wk.rB, uv) / 1.0f;
    const float invRotMassA = iiA * square(crAu), invRotMassB = iiB * square(crBu);
    float invMass = imA + invRotMassA + imB + invRotMassB;
    float mass = (invMass != 0.0f) ? 1.0f / invMass : 0.0f;
    float invGamma = 0.0f, bias = 0.0f;
    const float frequency = j.lower_angle, dampingRatio = j.upper_angle, restLength = j.reference_angle;
    if (frequency > 0.0f) {
        const float C = length - restLength;
        const float omega = 2.0f * 3.14159265358979323846264338327950288f * frequency;
        const float d = 2.0f * mass * dampingRatio * omega;
        const float k = mass * square(omega);
        const float h = c.dt;
        const float gamma = h * (d + h * k);
        invGamma = (gamma != 0.0f) ? 1.0f / gamma : 0.0f;
        bias = C * h * k * invGamma;
        invMass += invGamma;
        mass = (invMass != 0.0f) ? 1.0f / invMass : 0.0f;
    }
    wk.mass[0] = u.c; wk.mass[1] = u.s; wk.mass[2] = invGamma; wk.mass[3] = bias; wk.mass[4] = mass;
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio;
        const V2 P = j.impulse_x * uv;
        const float LA = cross(wk.rA, P) / 1.0f, LB = cross(wk.rB, P) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else j.impulse_x = 0.0f;
}
bool distance_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& velA, float imB, float iiB, Vel& velB)
{
    const V2 uv{wk.mass[0], wk.mass[1]};
    const V2 vpA = velA.l + rev_perp(wk.rA) * (velA.a / 1.0f);
    const V2 vpB = velB.l + rev_perp(wk.rB) * (velB.a / 1.0f);
    const float Cdot = dot(uv, vpB - vpA);
    const float impulse = -wk.mass[4] * (Cdot + wk.mass[3] + wk.mass[2] * j.impulse_x);
    j.impulse_x += impulse;
    const V2 P = impulse * uv;
    const float LA = cross(wk.rA, P) / 1.0f, LB = cross(wk.rB, P) / 1.0f;
    velA.l = velA.l - imA * P; velA.a -= iiA * LA;
    velB.l = velB.l + imB * P; velB.a += iiB * LB;
    return impulse == 0.0f;
}
bool distance_solve_position(const pr_revolute_joint& j, const RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos& posA,
                             float imB, float iiB, V2 lcB, Pos& posB, const JointSolverConf& c)
{
    if (j.lower_angle > 0.0f) return true; // soft constraint: no position correction
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const V2 relLoc = (posB.l + rB) - (posA.l + rA);
    float length;
    const Rot u = unit_vector_mag(relLoc, length);
    const V2 uv{u.c, u.s};
    const float deltaLength = length - j.reference_angle;
    const float C = std::clamp(deltaLength, -c.maxLinearCorrection, c.maxLinearCorrection);
    const float impulse = -wk.mass[4] * C;
    const V2 P = impulse * uv;
    const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
    posA.l = posA.l - imA * P; posA.a -= iiA * LA;
    posB.l = posB.l + imB * P; posB.a += iiB * LB;
    return std::fabs(C) < c.linearSlop;
}
// ---------------- RopeJointConf.cpp:58-214 (mass[0..1] = u, mass[4] = mass, mass[5] = length) ----------------
void rope_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                        float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    wk.rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    wk.rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const V2 posDelta = (posB.l + wk.rB) - (posA.l + wk.rA);
    float length;
    const Rot u = unit_vector_mag(posDelta, length);
    wk.mass[5] = length;
    const float C = length - j.reference_angle;
    j.limit_state = (C > 0.0f) ? LS_AT_UPPER : LS_INACTIVE;
    if (!(length > c.linearSlop)) { wk.mass[0] = wk.mass[1] = 0.0f; wk.mass[4] = 0.0f; j.impulse_x = 0.0f; return; }
    wk.mass[0] = u.c; wk.mass[1] = u.s;
    const V2 uv{u.c, u.s};
    const float crA = cross(wk.rA, uv) / 1.0f, crB = cross(wk.rB, uv) / 1.0f;
    const float invRotMassA = iiA * square(crA), invRotMassB = iiB * square(crB);
    const float invMass = imA + imB + invRotMassA + invRotMassB;
    wk.mass[4] = (invMass != 0.0f) ? 1.0f / invMass : 0.0f;
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio;
        const V2 P = j.impulse_x * uv;
        const float LA = cross(wk.rA, P) / 1.0f, LB = cross(wk.rB, P) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else j.impulse_x = 0.0f;
}
bool rope_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& velA, float imB, float iiB, Vel& velB, const JointSolverConf& c)
{
    const V2 uv{wk.mass[0], wk.mass[1]};
    const V2 vpA = velA.l + rev_perp(wk.rA) * (velA.a / 1.0f);
    const V2 vpB = velB.l + rev_perp(wk.rB) * (velB.a / 1.0f);
    const float C = wk.mass[5] - j.reference_angle;
    const float inv_h = (c.dt != 0.0f) ? 1.0f / c.dt : 0.0f;
    const float Cdot = dot(uv, vpB - vpA) + ((C < 0.0f) ? inv_h * C : 0.0f);
    float localImpulse = -wk.mass[4] * Cdot;
    const float oldImpulse = j.impulse_x;
    j.impulse_x = std::min(0.0f, j.impulse_x + localImpulse);
    localImpulse = j.impulse_x - oldImpulse;
    const V2 P = localImpulse * uv;
    const float LA = cross(wk.rA, P) / 1.0f, LB = cross(wk.rB, P) / 1.0f;
    velA.l = velA.l - imA * P; velA.a -= iiA * LA;
    velB.l = velB.l + imB * P; velB.a += iiB * LB;
    return localImpulse == 0.0f;
}
bool rope_solve_position(const pr_revolute_joint& j, const RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos& posA,
                         float imB, float iiB, V2 lcB, Pos& posB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const V2 posDelta = (posB.l + rB) - (posA.l + rA);
    float length;
    const Rot u = unit_vector_mag(posDelta, length);
    const V2 uv{u.c, u.s};
    const float C = std::clamp(length - j.reference_angle, 0.0f, c.maxLinearCorrection);
    const float localImpulse = -wk.mass[4] * C;
    const V2 P = localImpulse * uv;
    const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
    posA.l = posA.l - imA * P; posA.a -= iiA * LA;
    posB.l = posB.l + imB * P; posB.a += iiB * LB;
    return (length - j.reference_angle) < c.linearSlop;
}
// ---------------- FrictionJointConf.cpp:76-239, MotorJointConf.cpp:77-235 ----------------
// (mass[0..3] = linearMass rows, angularMass; motor: mass[4..5] = linearError, mass[6] = angularError)
V2 mat22_transform(const float* m, V2 v)
{ // Transform(v, M) = M * v, Vector.hpp:458-472
    float x = 0.0f; x += m[0] * v.x; x += m[1] * v.y;
    float y = 0.0f; y += m[2] * v.x; y += m[3] * v.y;
    return {x, y};
}
void fm_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                      float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c, bool motor)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    wk.rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    wk.rB = motor ? rotate(-lcB, qB) : rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const V2 rA = wk.rA, rB = wk.rB;
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
    const float cp = exx * eyy - exy * exy; // Invert(Matrix22), Math.hpp:430-442
    if (!almost_zero(cp)) {
        const float inverse = 1 / cp;
        wk.mass[0] = eyy * inverse; wk.mass[1] = -exy * inverse; wk.mass[2] = -exy * inverse; wk.mass[3] = exx * inverse;
    } else wk.mass[0] = wk.mass[1] = wk.mass[2] = wk.mass[3] = 0.0f;
    const float invRotInertia = iiA + iiB;
    wk.angularMass = (invRotInertia > 0.0f) ? 1.0f / invRotInertia : 0.0f;
    if (motor) {
        const V2 le = (posB.l + rB) - (posA.l + rA);
        wk.mass[4] = le.x; wk.mass[5] = le.y;
        wk.mass[6] = (posB.a - posA.a) - j.reference_angle;
    }
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio; j.impulse_y *= c.dtRatio; j.impulse_z *= c.dtRatio;
        const V2 P{j.impulse_x, j.impulse_y};
        const float crossAP = cross(rA, P) / 1.0f, crossBP = cross(rB, P) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * (crossAP + j.impulse_z);
        velB.l = velB.l + imB * P; velB.a += iiB * (crossBP + j.impulse_z);
    } else { j.impulse_x = j.impulse_y = 0.0f; j.impulse_z = 0.0f; }
}
bool fm_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& vA, float imB, float iiB, Vel& vB,
                       const JointSolverConf& c, bool motor)
{
    Vel velA = vA, velB = vB;
    const V2 rA = wk.rA, rB = wk.rB;
    const float h = c.dt;
    const float inv_h = (h != 0.0f) ? 1.0f / h : 0.0f;
    const float maxForce = j.motor_speed, maxTorque = j.max_motor_torque, correction = j.lower_angle;
    bool solved = true;
    {
        const float Cdot = motor ? (velB.a - velA.a) + inv_h * correction * wk.mass[6] : (velB.a - velA.a);
        const float angularImpulse = -wk.angularMass * Cdot;
        const float oldAngularImpulse = j.impulse_z;
        const float maxAngularImpulse = h * maxTorque;
        j.impulse_z = std::clamp(oldAngularImpulse + angularImpulse, -maxAngularImpulse, maxAngularImpulse);
        const float inc = j.impulse_z - oldAngularImpulse;
        if (inc != 0.0f) solved = false;
        if (!motor || inc != 0.0f) { velA.a -= iiA * inc; velB.a += iiB * inc; }
    }
    {
        const V2 vb = velB.l + (rev_perp(rB) * (velB.a / 1.0f));
        const V2 va = velA.l + (rev_perp(rA) * (velA.a / 1.0f));
        const V2 Cdot = motor ? (vb - va) + (inv_h * correction) * V2{wk.mass[4], wk.mass[5]} : vb - va;
        const V2 t = mat22_transform(wk.mass, Cdot);
        const V2 impulse{-t.x, -t.y};
        const V2 oldImpulse{j.impulse_x, j.impulse_y};
        V2 acc = oldImpulse + impulse;
        const float maxImpulse = h * maxForce;
        if (magsq(acc) > square(maxImpulse)) { const Rot u = unit_vector(acc, Rot{0.0f, 0.0f}); acc = V2{u.c * maxImpulse, u.s * maxImpulse}; }
        j.impulse_x = acc.x; j.impulse_y = acc.y;
        const V2 inc = acc - oldImpulse;
        const float angImpulseA = cross(rA, inc) / 1.0f, angImpulseB = cross(rB, inc) / 1.0f;
        const bool any = motor ? (inc.x != 0.0f || inc.y != 0.0f || angImpulseA != 0.0f || angImpulseB != 0.0f) : (inc.x != 0.0f || inc.y != 0.0f);
        if (any) solved = false;
        if (!motor || any) {
            velA.l = velA.l - imA * inc; velA.a -= iiA * angImpulseA;
            velB.l = velB.l + imB * inc; velB.a += iiB * angImpulseB;
        }
    }
    if (motor && solved) return true;
    vA = velA; vB = velB;
    return solved;
}
// ---------------- WeldJointConf.cpp:35-371 (mass[3 * i + k] = mass[i][k], angularMass = gamma, ex[0] = bias) ----------------
void weld_k(float* m, float imA, V2 rA, float iiA, float imB, V2 rB, float iiB)
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
void inverse22_of33(const float* k, float* r)
{ // GetInverse22, Math.hpp:478-489
    const float a = k[0], b = k[3], c = k[1], d = k[4];
    float det = (a * d) - (b * c);
    if (det != 0.0f) det = 1.0f / det;
    const float out[9] = {det * d, -det * c, 0.0f, -det * b, det * a, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int i = 0; i < 9; ++i) r[i] = out[i];
}
void sym_inverse33(const float* k, float* r)
{ // GetSymInverse33, Math.hpp:493-509
    const float invDet = dot3(col(k, 0), cross3(col(k, 1), col(k, 2)));
    const float det = (invDet != 0.0f) ? (1.0f / invDet) : invDet;
    const float a11 = k[0], a12 = k[3], a13 = k[6], a22 = k[4], a23 = k[7], a33 = k[8];
    const float ex_y = det * (a13 * a23 - a12 * a33);
    const float ey_z = det * (a13 * a12 - a11 * a23);
    const float ex_z = det * (a12 * a23 - a13 * a22);
    const float out[9] = {det * (a22 * a33 - a23 * a23), ex_y, ex_z, ex_y, det * (a11 * a33 - a13 * a13), ey_z, ex_z, ey_z, det * (a11 * a22 - a12 * a12)};
    for (int i = 0; i < 9; ++i) r[i] = out[i];
}
V2 transform2_by33(const float* A, V2 v) { return {A[0] * v.x + A[3] * v.y, A[1] * v.x + A[4] * v.y}; } // Math.hpp:522-526
V3 transform3_by33(const float* m, V3 v)
{ // m * v, Vector.hpp:458-472
    float e[3];
    const float vv[3] = {v.x, v.y, v.z};
    for (int b = 0; b < 3; ++b) { float el = 0.0f; for (int a = 0; a < 3; ++a) el += m[3 * b + a] * vv[a]; e[b] = el; }
    return {e[0], e[1], e[2]};
}
void weld_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                        float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    wk.rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    wk.rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    float K[9];
    weld_k(K, imA, wk.rA, iiA, imB, wk.rB, iiB);
    const float frequency = j.lower_angle, dampingRatio = j.upper_angle;
    if (frequency > 0.0f) {
        inverse22_of33(K, wk.mass);
        float invRotInertia = iiA + iiB;
        const float rotInertia = (invRotInertia > 0.0f) ? 1.0f / invRotInertia : 0.0f;
        const float C = posB.a - posA.a - j.reference_angle;
        const float omega = 2.0f * 3.14159265358979323846264338327950288f * frequency;
        const float d = 2.0f * rotInertia * dampingRatio * omega;
        const float k = rotInertia * omega * omega;
        const float h = c.dt;
        const float invGamma = h * (d + h * k);
        wk.angularMass = (invGamma != 0.0f) ? 1.0f / invGamma : 0.0f;
        wk.ex[0] = C * h * k * wk.angularMass;
        invRotInertia += wk.angularMass;
        wk.mass[8] = (invRotInertia != 0.0f) ? 1.0f / invRotInertia : 0.0f;
    } else if (K[8] == 0.0f) { inverse22_of33(K, wk.mass); wk.angularMass = 0.0f; wk.ex[0] = 0.0f; }
    else { sym_inverse33(K, wk.mass); wk.angularMass = 0.0f; wk.ex[0] = 0.0f; }
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio; j.impulse_y *= c.dtRatio; j.impulse_z *= c.dtRatio;
        const V2 P{j.impulse_x * 1.0f, j.impulse_y * 1.0f};
        const float L = j.impulse_z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = L + cross(wk.rA, P) / 1.0f, LB = L + cross(wk.rB, P) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else j.impulse_x = j.impulse_y = j.impulse_z = 0.0f;
}
bool weld_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& vA, float imB, float iiB, Vel& vB)
{
    const Vel oldA = vA, oldB = vB;
    Vel velA = vA, velB = vB;
    if (j.lower_angle > 0.0f) {
        const float Cdot2 = velB.a - velA.a;
        const float gamma = wk.angularMass * j.impulse_z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float impulse2 = -wk.mass[8] * (Cdot2 + wk.ex[0] + gamma);
        j.impulse_z += impulse2;
        velA.a -= iiA * impulse2 * 1.0f * 1.0f / (1.0f * 1.0f);
        velB.a += iiB * impulse2 * 1.0f * 1.0f / (1.0f * 1.0f);
        const V2 vb = velB.l + rev_perp(wk.rB) * (velB.a / 1.0f);
        const V2 va = velA.l + rev_perp(wk.rA) * (velA.a / 1.0f);
        const V2 Cdot1 = vb - va;
        const V2 t = transform2_by33(wk.mass, V2{Cdot1.x / 1.0f, Cdot1.y / 1.0f});
        const V2 impulse1{-t.x, -t.y};
        j.impulse_x += impulse1.x; j.impulse_y += impulse1.y;
        const V2 P{impulse1.x * 1.0f, impulse1.y * 1.0f};
        const float LA = cross(wk.rA, P) / 1.0f, LB = cross(wk.rB, P) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else {
        const V2 vb = velB.l + rev_perp(wk.rB) * (velB.a / 1.0f);
        const V2 va = velA.l + rev_perp(wk.rA) * (velA.a / 1.0f);
        const V2 Cdot1 = vb - va;
        const V3 Cdot{Cdot1.x / 1.0f, Cdot1.y / 1.0f, (velB.a - velA.a) / 1.0f};
        const V3 t = transform3_by33(wk.mass, Cdot);
        const V3 impulse{-t.x, -t.y, -t.z};
        j.impulse_x += impulse.x; j.impulse_y += impulse.y; j.impulse_z += impulse.z;
        const V2 P{impulse.x * 1.0f, impulse.y * 1.0f};
        const float L = impulse.z * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = L + cross(wk.rA, P) / 1.0f, LB = L + cross(wk.rB, P) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    }
    const bool changed = (velA.l != oldA.l) || (velA.a != oldA.a) || (velB.l != oldB.l) || (velB.a != oldB.a);
    if (changed) { vA = velA; vB = velB; return false; }
    return true;
}
bool weld_solve_position(const pr_revolute_joint& j, float imA, float iiA, V2 lcA, Pos& posA, float imB, float iiB, V2 lcB, Pos& posB,
                         const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    float K[9];
    weld_k(K, imA, rA, iiA, imB, rB, iiB);
    const V2 C1 = (posB.l + rB) - (posA.l + rA);
    const float positionError = std::sqrt(magsq(C1));
    float angularError = 0.0f;
    if (j.lower_angle > 0.0f) {
        const V2 s = solve22(K, C1);
        const V2 P = V2{-s.x, -s.y} * 1.0f;
        const float LA = cross(rA, P) / 1.0f, LB = cross(rB, P) / 1.0f;
        posA.l = posA.l - imA * P; posA.a -= iiA * LA;
        posB.l = posB.l + imB * P; posB.a += iiB * LB;
    } else {
        const float C2 = posB.a - posA.a - j.reference_angle;
        angularError = std::fabs(C2);
        V3 impulse;
        if (K[8] > 0.0f) { const V3 s = solve33(K, V3{C1.x, C1.y, C2}); impulse = V3{-s.x, -s.y, -s.z}; }
        else { const V2 s = solve22(K, C1); impulse = V3{-s.x, -s.y, 0.0f}; }
        const V2 P = V2{impulse.x * 1.0f, impulse.y * 1.0f} * 1.0f;
        const float L = impulse.z * 1.0f * 1.0f / 1.0f;
        const float LA = L + cross(rA, P) / 1.0f, LB = L + cross(rB, P) / 1.0f;
        posA.l = posA.l - imA * P; posA.a -= iiA * LA;
        posB.l = posB.l + imB * P; posB.a += iiB * LB;
    }
    return (positionError <= c.linearSlop) && (angularError <= c.angularSlop);
}
// ---------------- TargetJointConf.cpp:52-243 (one body: the row's body_a repeats body_b; mass[0..3] = mass, angularMass = gamma, ex[0..1] = C) ----
void target_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const Rot qB = rot_from_angle(posB.a);
    const float mass = (imB != 0.0f) ? (1.0f / imB) : 0.0f;
    const float omega = 2.0f * 3.14159265358979323846264338327950288f * j.lower_angle;
    const float d = 2.0f * mass * j.upper_angle * omega;
    const float k = mass * square(omega);
    const float h = c.dt;
    const float tmp = d + h * k;
    const float invGamma = h * tmp;
    wk.angularMass = (invGamma != 0.0f) ? 1.0f / invGamma : 0.0f;
    const float beta = h * k * wk.angularMass;
    wk.rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const float exx = imB + (iiB * square(wk.rB.y) / 1.0f) + wk.angularMass;
    const float exy = -iiB * wk.rB.x * wk.rB.y / 1.0f;
    const float eyy = imB + (iiB * square(wk.rB.x) / 1.0f) + wk.angularMass;
    const float cp = exx * eyy - exy * exy;
    if (!almost_zero(cp)) {
        const float inverse = 1 / cp;
        wk.mass[0] = eyy * inverse; wk.mass[1] = -exy * inverse; wk.mass[2] = -exy * inverse; wk.mass[3] = exx * inverse;
    } else wk.mass[0] = wk.mass[1] = wk.mass[2] = wk.mass[3] = 0.0f;
    const V2 C = ((posB.l + wk.rB) - V2{j.local_anchor_a.x, j.local_anchor_a.y}) * beta;
    wk.ex[0] = C.x; wk.ex[1] = C.y;
    velB.a *= 0.98f;
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio; j.impulse_y *= c.dtRatio;
        const V2 P{j.impulse_x, j.impulse_y};
        const float crossBP = cross(wk.rB, P) / 1.0f;
        velB.l = velB.l + imB * P; velB.a += iiB * crossBP;
    } else j.impulse_x = j.impulse_y = 0.0f;
}
bool target_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imB, float iiB, Vel& velB, const JointSolverConf& c)
{
    const V2 Cdot = velB.l + (rev_perp(wk.rB) * (velB.a / 1.0f));
    const V2 ev = Cdot + (V2{wk.ex[0], wk.ex[1]} + (wk.angularMass * V2{j.impulse_x, j.impulse_y}));
    const V2 oldImpulse{j.impulse_x, j.impulse_y};
    const V2 addImpulse = mat22_transform(wk.mass, V2{-ev.x, -ev.y});
    V2 impulse = oldImpulse + addImpulse;
    const float maxImpulse = c.dt * j.motor_speed;
    if (magsq(impulse) > square(maxImpulse)) { const Rot u = unit_vector(impulse, Rot{0.0f, 0.0f}); impulse = V2{u.c * maxImpulse, u.s * maxImpulse}; }
    j.impulse_x = impulse.x; j.impulse_y = impulse.y;
    const V2 inc = impulse - oldImpulse;
    const float angImpulseB = cross(wk.rB, inc) / 1.0f;
    velB.l = velB.l + imB * inc; velB.a += iiB * angImpulseB;
    return inc.x == 0.0f && inc.y == 0.0f;
}
// ---------------- PrismaticJointConf.cpp:150-533 ----------------
// (mass[3 * i + k] = K[i][k], angularMass = motorMass, ex[0..1] = axis, ex[2..3] = perp, ex[4] = s1, ex[5] = s2, ex[6] = a1, ex[7] = a2)
inline float bits_to_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline Rot local_x_axis(const pr_revolute_joint& j) { return Rot{bits_to_float(j.enable_limit), bits_to_float(j.enable_motor)}; }
inline bool has_limit(const pr_revolute_joint& j) { return (j.collide_connected & PR_JOINT_ENABLE_LIMIT) != 0u; }
inline bool has_motor(const pr_revolute_joint& j) { return (j.collide_connected & PR_JOINT_ENABLE_MOTOR) != 0u; }
void prismatic_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                             float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    wk.rA = rA; wk.rB = rB;
    const V2 d = (posB.l - posA.l) + rB - rA;
    const Rot axis = rot_rotate(local_x_axis(j), qA);
    const float a1 = cross(d + rA, axis), a2 = cross(rB, axis);
    const float invRotMassA = iiA * square(a1) / 1.0f, invRotMassB = iiB * square(a2) / 1.0f;
    const float totalInvMass = imA + imB + invRotMassA + invRotMassB;
    wk.angularMass = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
    const Rot perp = rot_rotate(rev_perp(local_x_axis(j)), qA);
    const float s1 = cross(d + rA, perp), s2 = cross(rB, perp);
    const float invRotMassA2 = iiA * square(s1) / 1.0f, invRotMassB2 = iiB * square(s2) / 1.0f;
    const float k11 = imA + imB + invRotMassA2 + invRotMassB2;
    const float k12 = (iiA * s1 + iiB * s2) * 1.0f * 1.0f / 1.0f;
    const float k13 = (iiA * s1 * a1 + iiB * s2 * a2) / 1.0f;
    const float totalInvRotInertia = iiA + iiB;
    const float k22 = (totalInvRotInertia == 0.0f) ? 1.0f : totalInvRotInertia;
    const float k23 = (iiA * a1 + iiB * a2) * 1.0f * 1.0f / 1.0f;
    const float k33 = totalInvMass;
    const float K[9] = {k11, k12, k13, k12, k22, k23, k13, k23, k33};
    for (int i = 0; i < 9; ++i) wk.mass[i] = K[i];
    wk.ex[0] = axis.c; wk.ex[1] = axis.s; wk.ex[2] = perp.c; wk.ex[3] = perp.s; wk.ex[4] = s1; wk.ex[5] = s2; wk.ex[6] = a1; wk.ex[7] = a2;
    if (has_limit(j)) {
        const float jointTranslation = dot(axis, d);
        if (std::fabs(j.upper_angle - j.lower_angle) < (c.linearSlop * 2.0f)) j.limit_state = LS_EQUAL;
        else if (jointTranslation <= j.lower_angle) {
            if (j.limit_state != LS_AT_LOWER) { j.limit_state = LS_AT_LOWER; j.impulse_z = 0; }
        } else if (jointTranslation >= j.upper_angle) {
            if (j.limit_state != LS_AT_UPPER) { j.limit_state = LS_AT_UPPER; j.impulse_z = 0; }
        } else { j.limit_state = LS_INACTIVE; j.impulse_z = 0; }
    } else { j.limit_state = LS_INACTIVE; j.impulse_z = 0; }
    if (!has_motor(j)) j.angular_motor_impulse = 0.0f;
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio; j.impulse_y *= c.dtRatio; j.impulse_z *= c.dtRatio;
        j.angular_motor_impulse *= c.dtRatio;
        const V2 ul = j.impulse_x * perp;
        const V2 Px{ul.x * 1.0f, ul.y * 1.0f};
        const float Pxs1 = j.impulse_x * s1 * 1.0f / 1.0f, Pxs2 = j.impulse_x * s2 * 1.0f / 1.0f;
        const float PzLength = j.angular_motor_impulse + j.impulse_z * 1.0f;
        const V2 Pz = PzLength * axis;
        const V2 P = Px + Pz;
        const float L = j.impulse_y * 1.0f * 1.0f / (1.0f * 1.0f);
        const float LA = L + (Pxs1 * 1.0f + PzLength * a1) / 1.0f;
        const float LB = L + (Pxs2 * 1.0f + PzLength * a2) / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else { j.impulse_x = j.impulse_y = j.impulse_z = 0.0f; j.angular_motor_impulse = 0.0f; }
}
bool prismatic_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& vA, float imB, float iiB, Vel& vB, const JointSolverConf& c)
{
    const Vel oldA = vA, oldB = vB;
    Vel velA = vA, velB = vB;
    const Rot axis{wk.ex[0], wk.ex[1]}, perp{wk.ex[2], wk.ex[3]};
    const float s1 = wk.ex[4], s2 = wk.ex[5], a1 = wk.ex[6], a2 = wk.ex[7];
    if (has_motor(j) && j.limit_state != LS_EQUAL) {
        const float vDot = dot(axis, velB.l - velA.l);
        const float Cdot = vDot + (a2 * velB.a - a1 * velA.a) / 1.0f;
        float impulse = wk.angularMass * (j.motor_speed * 1.0f / 1.0f - Cdot);
        const float oldImpulse = j.angular_motor_impulse;
        const float maxImpulse = c.dt * j.max_motor_torque;
        j.angular_motor_impulse = std::clamp(j.angular_motor_impulse + impulse, -maxImpulse, maxImpulse);
        impulse = j.angular_motor_impulse - oldImpulse;
        const V2 P = impulse * axis;
        const float LA = impulse * a1 / 1.0f, LB = impulse * a2 / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    }
    const V2 velDelta = velB.l - velA.l;
    const float sRotSpeed = (s2 * velB.a - s1 * velA.a) / 1.0f;
    const V2 Cdot1{dot(perp, velDelta) + sRotSpeed, velB.a - velA.a};
    if (has_limit(j) && (j.limit_state != LS_INACTIVE)) {
        const float deltaDot = dot(axis, velDelta);
        const float aRotSpeed = (a2 * velB.a - a1 * velA.a) / 1.0f;
        const float Cdot2 = deltaDot + aRotSpeed;
        const V3 f1{j.impulse_x, j.impulse_y, j.impulse_z};
        const V3 sol = solve33(wk.mass, V3{-Cdot1.x, -Cdot1.y, -Cdot2});
        V3 imp{f1.x + sol.x, f1.y + sol.y, f1.z + sol.z};
        if (j.limit_state == LS_AT_LOWER) imp.z = std::max(imp.z, 0.0f);
        else if (j.limit_state == LS_AT_UPPER) imp.z = std::min(imp.z, 0.0f);
        const V2 b = -Cdot1 - (imp.z - f1.z) * V2{wk.mass[6], wk.mass[7]};
        const V2 f2r = solve22(wk.mass, b) + V2{f1.x, f1.y};
        imp.x = f2r.x; imp.y = f2r.y;
        j.impulse_x = imp.x; j.impulse_y = imp.y; j.impulse_z = imp.z;
        const V3 df{imp.x - f1.x, imp.y - f1.y, imp.z - f1.z};
        const V2 ulP = df.x * perp + df.z * axis;
        const V2 P{ulP.x * 1.0f, ulP.y * 1.0f};
        const float LA = (df.x * s1 + df.y * 1.0f + df.z * a1) * 1.0f / 1.0f;
        const float LB = (df.x * s2 + df.y * 1.0f + df.z * a2) * 1.0f / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else {
        const V2 df = solve22(wk.mass, -Cdot1);
        j.impulse_x += df.x; j.impulse_y += df.y;
        const V2 ulP = df.x * perp;
        const V2 P{ulP.x * 1.0f, ulP.y * 1.0f};
        const float LA = (df.x * s1 + df.y * 1.0f) * 1.0f / 1.0f;
        const float LB = (df.x * s2 + df.y * 1.0f) * 1.0f / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    }
    const bool changed = (velA.l != oldA.l) || (velA.a != oldA.a) || (velB.l != oldB.l) || (velB.a != oldB.a);
    if (changed) { vA = velA; vB = velB; return false; }
    return true;
}
bool prismatic_solve_position(const pr_revolute_joint& j, float imA, float iiA, V2 lcA, Pos& posA, float imB, float iiB, V2 lcB, Pos& posB,
                              const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const V2 d = (posB.l + rB) - (posA.l + rA);
    const Rot axis = rot_rotate(local_x_axis(j), qA);
    const float a1 = cross(d + rA, axis), a2 = cross(rB, axis);
    const Rot perp = rot_rotate(rev_perp(local_x_axis(j)), qA);
    const float s1 = cross(d + rA, perp), s2 = cross(rB, perp);
    const V2 C1{dot(perp, d) / 1.0f, (posB.a - posA.a - j.reference_angle) / 1.0f};
    float linearError = std::fabs(C1.x) * 1.0f;
    const float angularError = std::fabs(C1.y) * 1.0f;
    bool active = false;
    float C2 = 0.0f;
    if (has_limit(j)) {
        const float translation = dot(axis, d);
        if (std::fabs(j.upper_angle - j.lower_angle) < (2.0f * c.linearSlop)) {
            C2 = std::clamp(translation, -c.maxLinearCorrection, c.maxLinearCorrection);
            linearError = std::max(linearError, std::fabs(translation));
            active = true;
        } else if (translation <= j.lower_angle) {
            C2 = std::clamp(translation - j.lower_angle + c.linearSlop, -c.maxLinearCorrection, 0.0f);
            linearError = std::max(linearError, j.lower_angle - translation);
            active = true;
        } else if (translation >= j.upper_angle) {
            C2 = std::clamp(translation - j.upper_angle - c.linearSlop, 0.0f, c.maxLinearCorrection);
            linearError = std::max(linearError, translation - j.upper_angle);
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
        const float K[9] = {k11, k12, k13, k12, k22, k23, k13, k23, k33};
        impulse = solve33(K, V3{-C1.x, -C1.y, -C2});
    } else {
        const float k11 = imA + iiA * s1 * s1 / 1.0f + imB + iiB * s2 * s2 / 1.0f;
        const float k12 = iiA * s1 * 1.0f / 1.0f + iiB * s2 * 1.0f / 1.0f;
        float k22 = iiA + iiB;
        if (k22 == 0.0f) k22 = 1.0f;
        const float cp = cross(V2{k11, k12}, V2{k12, k22}); // Solve(Matrix22, Vector2), Math.hpp:414-426
        const V2 b = -C1;
        V2 impulse1{0.0f, 0.0f};
        if (!almost_zero(cp)) {
            const float inverse = 1 / cp;
            impulse1 = V2{(k22 * b.x - k12 * b.y) * inverse, (k11 * b.y - k12 * b.x) * inverse};
        }
        impulse = V3{impulse1.x, impulse1.y, 0.0f};
    }
    const V2 P = (impulse.x * perp + impulse.z * axis) * 1.0f * 1.0f;
    const float LA = (impulse.x * s1 + impulse.y * 1.0f + impulse.z * a1) * 1.0f * 1.0f / 1.0f;
    const float LB = (impulse.x * s2 + impulse.y * 1.0f + impulse.z * a2) * 1.0f * 1.0f / 1.0f;
    posA.l = posA.l - imA * P; posA.a -= iiA * LA;
    posB.l = posB.l + imB * P; posB.a += iiB * LB;
    return (linearError <= c.linearSlop) && (angularError <= c.angularSlop);
}
// ---------------- WheelJointConf.cpp:103-362 ----------------
// (mass[0..1] = ay, mass[2] = sAy, mass[3] = sBy, mass[4] = mass, mass[5..6] = ax, mass[7] = sAx, mass[8] = sBx, ex[0] = springMass, ex[1] = bias,
//  ex[2] = gamma, angularMass)
void wheel_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                         float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    wk.rA = rA; wk.rB = rB;
    const V2 dd = (posB.l + rB) - (posA.l + rA);
    const Rot ay = rot_rotate(rev_perp(local_x_axis(j)), qA);
    const float sAy = cross(dd + rA, ay), sBy = cross(rB, ay);
    {
        const float invRotMassA = iiA * square(sAy) / 1.0f, invRotMassB = iiB * square(sBy) / 1.0f;
        const float invMass = imA + imB + invRotMassA + invRotMassB;
        wk.mass[4] = (invMass > 0.0f) ? 1.0f / invMass : 0.0f;
    }
    wk.mass[0] = ay.c; wk.mass[1] = ay.s; wk.mass[2] = sAy; wk.mass[3] = sBy;
    float springMass = 0.0f, bias = 0.0f, gamma = 0.0f;
    Rot ax{0.0f, 0.0f};
    float sAx = 0.0f, sBx = 0.0f;
    if (j.lower_angle > 0.0f) {
        ax = rot_rotate(local_x_axis(j), qA);
        sAx = cross(dd + rA, ax); sBx = cross(rB, ax);
        const float invRotMassA = iiA * square(sAx) / 1.0f, invRotMassB = iiB * square(sBx) / 1.0f;
        const float invMass = imA + imB + invRotMassA + invRotMassB;
        if (invMass > 0.0f) {
            springMass = 1.0f / invMass;
            const float C = dot(dd, ax);
            const float omega = 2.0f * 3.14159265358979323846264338327950288f * j.lower_angle;
            const float d = 2.0f * springMass * j.upper_angle * omega;
            const float k = springMass * omega * omega;
            const float h = c.dt;
            const float invGamma = h * (d + h * k);
            gamma = (invGamma > 0.0f) ? 1.0f / invGamma : 0.0f;
            bias = C * h * k * gamma;
            const float totalInvMass = invMass + gamma;
            springMass = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
        }
    } else j.impulse_y = 0.0f;
    wk.mass[5] = ax.c; wk.mass[6] = ax.s; wk.mass[7] = sAx; wk.mass[8] = sBx;
    wk.ex[0] = springMass; wk.ex[1] = bias; wk.ex[2] = gamma;
    if (has_motor(j)) {
        const float invRotInertia = iiA + iiB;
        wk.angularMass = (invRotInertia > 0.0f) ? 1.0f / invRotInertia : 0.0f;
    } else { wk.angularMass = 0.0f; j.impulse_z = 0.0f; }
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio; j.impulse_y *= c.dtRatio; j.impulse_z *= c.dtRatio;
        const V2 P = j.impulse_x * ay + j.impulse_y * ax;
        const float LA = (j.impulse_x * sAy + j.impulse_y * sAx) / 1.0f + j.impulse_z;
        const float LB = (j.impulse_x * sBy + j.impulse_y * sBx) / 1.0f + j.impulse_z;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    } else j.impulse_x = j.impulse_y = j.impulse_z = 0.0f;
}
bool wheel_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& vA, float imB, float iiB, Vel& vB, const JointSolverConf& c)
{
    const Vel oldA = vA, oldB = vB;
    Vel velA = vA, velB = vB;
    const Rot ay{wk.mass[0], wk.mass[1]}, ax{wk.mass[5], wk.mass[6]};
    const float sAy = wk.mass[2], sBy = wk.mass[3], sAx = wk.mass[7], sBx = wk.mass[8];
    {
        const float dt = dot(ax, velB.l - velA.l);
        const float Cdot = dt + sBx * velB.a / 1.0f - sAx * velA.a / 1.0f;
        const float impulse = -wk.ex[0] * (Cdot + wk.ex[1] + wk.ex[2] * j.impulse_y);
        j.impulse_y += impulse;
        const V2 P = impulse * ax;
        const float LA = impulse * sAx / 1.0f, LB = impulse * sBx / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    }
    {
        const float Cdot = (velB.a - velA.a - j.motor_speed);
        float impulse = -wk.angularMass * Cdot;
        const float oldImpulse = j.impulse_z;
        const float maxImpulse = c.dt * j.max_motor_torque;
        j.impulse_z = std::clamp(j.impulse_z + impulse, -maxImpulse, maxImpulse);
        impulse = j.impulse_z - oldImpulse;
        velA.a -= iiA * impulse;
        velB.a += iiB * impulse;
    }
    {
        const float dt = dot(ay, velB.l - velA.l);
        const float Cdot = dt + sBy * velB.a / 1.0f - sAy * velA.a / 1.0f;
        const float impulse = -wk.mass[4] * Cdot;
        j.impulse_x += impulse;
        const V2 P = impulse * ay;
        const float LA = impulse * sAy / 1.0f, LB = impulse * sBy / 1.0f;
        velA.l = velA.l - imA * P; velA.a -= iiA * LA;
        velB.l = velB.l + imB * P; velB.a += iiB * LB;
    }
    const bool changed = (velA.l != oldA.l) || (velA.a != oldA.a) || (velB.l != oldB.l) || (velB.a != oldB.a);
    if (changed) { vA = velA; vB = velB; return false; }
    return true;
}
bool wheel_solve_position(const pr_revolute_joint& j, const RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos& posA, float imB, float iiB, V2 lcB,
                          Pos& posB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const V2 d = (posB.l - posA.l) + (rB - rA);
    const Rot ay = rot_rotate(rev_perp(local_x_axis(j)), qA);
    const float sAy = cross(d + rA, ay), sBy = cross(rB, ay);
    const float C = dot(d, ay);
    const float invRotMassA = iiA * square(wk.mass[2]) / 1.0f, invRotMassB = iiB * square(wk.mass[3]) / 1.0f; // (:339-340: the stored sAy / sBy)
    const float k = imA + imB + invRotMassA + invRotMassB;
    const float impulse = (k != 0.0f) ? -(C / k) : 0.0f * 1.0f * 1.0f;
    const V2 P = impulse * ay;
    const float LA = impulse * sAy / 1.0f, LB = impulse * sBy / 1.0f;
    posA.l = posA.l - imA * P; posA.a -= iiA * LA;
    posB.l = posB.l + imB * P; posB.a += iiB * LB;
    return std::fabs(C) <= c.linearSlop;
}
// ---------------- PulleyJointConf.cpp:91-282 (mass[0..1] = uA, mass[2..3] = uB, mass[4] = mass) ----------------
// row: impulse = impulse_x, groundAnchorA = (impulse_y, impulse_z), groundAnchorB = (angular_motor_impulse, reference_angle),
// lengthA / lengthB = the bits of enable_limit / enable_motor, ratio = lower_angle
void pulley_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                          float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    const Rot qA = rot_from_angle(posA.a), qB = rot_from_angle(posB.a);
    wk.rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, qA);
    wk.rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, qB);
    const float ratio = j.lower_angle;
    const V2 pulleyAxisA = posA.l + wk.rA - V2{j.impulse_y, j.impulse_z};
    const V2 pulleyAxisB = posB.l + wk.rB - V2{j.angular_motor_impulse, j.reference_angle};
    const Rot uA = unit_vector(pulleyAxisA, Rot{0.0f, 0.0f}), uB = unit_vector(pulleyAxisB, Rot{0.0f, 0.0f});
    const float ruA = cross(wk.rA, uA), ruB = cross(wk.rB, uB);
    const float totInvMassA = imA + (iiA * square(ruA)) / 1.0f;
    const float totInvMassB = imB + (iiB * square(ruB)) / 1.0f;
    const float totalInvMass = totInvMassA + ratio * ratio * totInvMassB;
    wk.mass[0] = uA.c; wk.mass[1] = uA.s; wk.mass[2] = uB.c; wk.mass[3] = uB.s;
    wk.mass[4] = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
    if (c.warmStart) {
        j.impulse_x *= c.dtRatio;
        const V2 PA = -(j.impulse_x) * uA;
        const V2 PB = (-ratio * j.impulse_x) * uB;
        velA.l = velA.l + imA * PA; velA.a += iiA * cross(wk.rA, PA) / 1.0f;
        velB.l = velB.l + imB * PB; velB.a += iiB * cross(wk.rB, PB) / 1.0f;
    } else j.impulse_x = 0.0f;
}
bool pulley_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& velA, float imB, float iiB, Vel& velB)
{
    const Rot uA{wk.mass[0], wk.mass[1]}, uB{wk.mass[2], wk.mass[3]};
    const float ratio = j.lower_angle;
    const V2 vpA = velA.l + rev_perp(wk.rA) * (velA.a / 1.0f);
    const V2 vpB = velB.l + rev_perp(wk.rB) * (velB.a / 1.0f);
    const float Cdot = -dot(uA, vpA) - ratio * dot(uB, vpB);
    const float impulse = -wk.mass[4] * Cdot;
    j.impulse_x += impulse;
    const V2 PA = -impulse * uA;
    const V2 PB = -ratio * impulse * uB;
    velA.l = velA.l + imA * PA; velA.a += iiA * cross(wk.rA, PA) / 1.0f;
    velB.l = velB.l + imB * PB; velB.a += iiB * cross(wk.rB, PB) / 1.0f;
    return impulse == 0.0f;
}
bool pulley_solve_position(const pr_revolute_joint& j, float imA, float iiA, V2 lcA, Pos& posA, float imB, float iiB, V2 lcB, Pos& posB,
                           const JointSolverConf& c)
{
    const V2 rA = rotate(V2{j.local_anchor_a.x, j.local_anchor_a.y} - lcA, rot_from_angle(posA.a));
    const V2 rB = rotate(V2{j.local_anchor_b.x, j.local_anchor_b.y} - lcB, rot_from_angle(posB.a));
    const float ratio = j.lower_angle;
    float lengthA, lengthB;
    const V2 pA = posA.l + rA - V2{j.impulse_y, j.impulse_z};
    const Rot uA = unit_vector_mag(pA, lengthA);
    const V2 pB = posB.l + rB - V2{j.angular_motor_impulse, j.reference_angle};
    const Rot uB = unit_vector_mag(pB, lengthB);
    const float ruA = cross(rA, uA), ruB = cross(rB, uB);
    const float totalInvMassA = imA + iiA * square(ruA) / 1.0f;
    const float totalInvMassB = imB + iiB * square(ruB) / 1.0f;
    const float totalInvMass = totalInvMassA + square(ratio) * totalInvMassB;
    const float mass = (totalInvMass > 0.0f) ? 1.0f / totalInvMass : 0.0f;
    const float srcLengthRatio = bits_to_float(j.enable_limit) + ratio * bits_to_float(j.enable_motor);
    const float dstLengthRatio = lengthA + ratio * lengthB;
    const float C = srcLengthRatio - dstLengthRatio;
    const float linearError = std::fabs(C);
    const float impulse = -mass * C;
    const V2 PA = -impulse * uA;
    const V2 PB = -ratio * impulse * uB;
    posA.l = posA.l + imA * PA; posA.a += iiA * cross(rA, PA) / 1.0f;
    posB.l = posB.l + imB * PB; posB.a += iiB * cross(rB, PB) / 1.0f;
    return linearError < c.linearSlop;
}
} // namespace
// ---------------- GearJointConf.cpp:171-468 ----------------
// (mass[0..1] = JvAC, mass[2..3] = JvBD, mass[4] = JwA, mass[5] = JwB, mass[6] = JwC, mass[7] = JwD, mass[8] = mass)
void gear_init_velocity(pr_revolute_joint& j, const pr_gear_joint& g, RevoluteWork& wk, GearBody A, GearBody B, GearBody C, GearBody D, const JointSolverConf& c)
{
    Vel velA = *A.vel, velB = *B.vel, velC = *C.vel, velD = *D.vel;
    const float ratio = j.lower_angle;
    V2 JvAC, JvBD; float JwA, JwB, JwC, JwD;
    float invMass = 0.0f;
    if (g.ac.type == PR_GEAR_PRISMATIC) {
        const Rot qA = rot_from_angle(A.pos->a), qC = rot_from_angle(C.pos->a);
        const Rot u = rot_rotate(Rot{g.ac.local_axis.x, g.ac.local_axis.y}, qC);
        const V2 rC = rotate(V2{g.ac.local_anchor_a.x, g.ac.local_anchor_a.y} - C.lc, qC);
        const V2 rA = rotate(V2{g.ac.local_anchor_b.x, g.ac.local_anchor_b.y} - A.lc, qA);
        JvAC = 1.0f * u;
        JwC = cross(rC, u); JwA = cross(rA, u);
        const float invRotMassC = C.invI * square(JwC) / 1.0f, invRotMassA = A.invI * square(JwA) / 1.0f;
        invMass += C.invM + A.invM + invRotMassC + invRotMassA;
    } else {
        JvAC = V2{0.0f, 0.0f}; JwA = 1.0f; JwC = 1.0f;
        invMass += A.invI + C.invI;
    }
    if (g.bd.type == PR_GEAR_PRISMATIC) {
        const Rot qB = rot_from_angle(B.pos->a), qD = rot_from_angle(D.pos->a);
        const Rot u = rot_rotate(Rot{g.bd.local_axis.x, g.bd.local_axis.y}, qD);
        const V2 rD = rotate(V2{g.bd.local_anchor_a.x, g.bd.local_anchor_a.y} - D.lc, qD);
        const V2 rB = rotate(V2{g.bd.local_anchor_b.x, g.bd.local_anchor_b.y} - B.lc, qB);
        JvBD = ratio * u;
        JwD = ratio * cross(rD, u); JwB = ratio * cross(rB, u);
        const float invRotMassD = D.invI * square(JwD) / 1.0f, invRotMassB = B.invI * square(JwB) / 1.0f;
        invMass += square(ratio) * (D.invM + B.invM) + invRotMassD + invRotMassB;
    } else {
        JvBD = V2{0.0f, 0.0f}; JwB = ratio * 1.0f; JwD = ratio * 1.0f;
        invMass += square(ratio) * (B.invI + D.invI);
    }
    wk.mass[0] = JvAC.x; wk.mass[1] = JvAC.y; wk.mass[2] = JvBD.x; wk.mass[3] = JvBD.y; wk.mass[4] = JwA; wk.mass[5] = JwB; wk.mass[6] = JwC; wk.mass[7] = JwD;
    wk.mass[8] = (invMass > 0.0f) ? 1.0f / invMass : 0.0f;
    if (c.warmStart) {
        const float imp = j.impulse_x;
        velA.l = velA.l + (A.invM * imp) * JvAC; velA.a += A.invI * imp * JwA / 1.0f;
        velB.l = velB.l + (B.invM * imp) * JvBD; velB.a += B.invI * imp * JwB / 1.0f;
        velC.l = velC.l - (C.invM * imp) * JvAC; velC.a -= C.invI * imp * JwC / 1.0f;
        velD.l = velD.l - (D.invM * imp) * JvBD; velD.a -= D.invI * imp * JwD / 1.0f;
    } else j.impulse_x = 0.0f;
    *A.vel = velA; *B.vel = velB; *C.vel = velC; *D.vel = velD;
}
bool gear_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, GearBody A, GearBody B, GearBody C, GearBody D)
{
    Vel velA = *A.vel, velB = *B.vel, velC = *C.vel, velD = *D.vel;
    const V2 JvAC{wk.mass[0], wk.mass[1]}, JvBD{wk.mass[2], wk.mass[3]};
    const float JwA = wk.mass[4], JwB = wk.mass[5], JwC = wk.mass[6], JwD = wk.mass[7];
    const float acDot = dot(JvAC, velA.l - velC.l);
    const float bdDot = dot(JvBD, velB.l - velD.l);
    const float Cdot = acDot + bdDot + (JwA * velA.a - JwC * velC.a) / 1.0f + (JwB * velB.a - JwD * velD.a) / 1.0f;
    const float impulse = -wk.mass[8] * 1.0f * Cdot;
    j.impulse_x += impulse;
    velA.l = velA.l + (A.invM * impulse) * JvAC; velA.a += A.invI * impulse * JwA / 1.0f;
    velB.l = velB.l + (B.invM * impulse) * JvBD; velB.a += B.invI * impulse * JwB / 1.0f;
    velC.l = velC.l - (C.invM * impulse) * JvAC; velC.a -= C.invI * impulse * JwC / 1.0f;
    velD.l = velD.l - (D.invM * impulse) * JvBD; velD.a -= D.invI * impulse * JwD / 1.0f;
    *A.vel = velA; *B.vel = velB; *C.vel = velC; *D.vel = velD;
    return impulse == 0.0f;
}
namespace {
Pos cap_position(V2 lin, float ang, const JointSolverConf& c)
{ // Cap(Position, ConstraintSolverConf), d2/Math.cpp:63-71
    const float lsquared = magsq(lin);
    if (lsquared > square(c.maxLinearCorrection)) { const float sc = c.maxLinearCorrection / std::sqrt(lsquared); lin.x *= sc; lin.y *= sc; }
    return Pos{lin, std::clamp(ang, -c.maxAngularCorrection, +c.maxAngularCorrection)};
}
}
bool gear_solve_position(const pr_revolute_joint& j, const pr_gear_joint& g, GearBody A, GearBody B, GearBody C, GearBody D, const JointSolverConf& c)
{
    const Pos posA = *A.pos, posB = *B.pos, posC = *C.pos, posD = *D.pos;
    const float ratio = j.lower_angle, constant = j.upper_angle;
    if (!std::isfinite(constant)) return true;
    V2 JvAC{0.0f, 0.0f}, JvBD{0.0f, 0.0f};
    float JwA = 0.0f, JwB = 0.0f, JwC = 0.0f, JwD = 0.0f, coordinateA = 0.0f, coordinateB = 0.0f, invMass = 0.0f;
    if (g.ac.type == PR_GEAR_PRISMATIC) {
        const Rot qA = rot_from_angle(posA.a), qC = rot_from_angle(posC.a);
        const Rot lax{g.ac.local_axis.x, g.ac.local_axis.y};
        const Rot u = rot_rotate(lax, qC);
        const V2 rC = rotate(V2{g.ac.local_anchor_a.x, g.ac.local_anchor_a.y} - C.lc, qC);
        const V2 rA = rotate(V2{g.ac.local_anchor_b.x, g.ac.local_anchor_b.y} - A.lc, qA);
        JvAC = u * 1.0f;
        JwC = cross(rC, u); JwA = cross(rA, u);
        const float invLinMass = C.invM + A.invM;
        const float invRotMassC = C.invI * square(JwC * 1.0f / 1.0f), invRotMassA = A.invI * square(JwA * 1.0f / 1.0f);
        invMass += invLinMass + invRotMassC + invRotMassA;
        const V2 pC = V2{g.ac.local_anchor_a.x, g.ac.local_anchor_a.y} - C.lc;
        const V2 pA = inv_rotate(rA + (posA.l - posC.l), qC);
        coordinateA = dot(pA - pC, lax) / 1.0f;
    } else if (g.ac.type == PR_GEAR_REVOLUTE) {
        JwA = 1.0f; JwC = 1.0f;
        invMass += A.invI + C.invI;
        coordinateA = ((posA.a - posC.a) - g.ac.reference_angle) / 1.0f;
    }
    if (g.bd.type == PR_GEAR_PRISMATIC) {
        const Rot qB = rot_from_angle(posB.a), qD = rot_from_angle(posD.a);
        const Rot lax{g.bd.local_axis.x, g.bd.local_axis.y};
        const Rot u = rot_rotate(lax, qD);
        const V2 rD = rotate(V2{g.bd.local_anchor_a.x, g.bd.local_anchor_a.y} - D.lc, qD);
        const V2 rB = rotate(V2{g.bd.local_anchor_b.x, g.bd.local_anchor_b.y} - B.lc, qB);
        JvBD = ratio * u;
        JwD = ratio * cross(rD, u); JwB = ratio * cross(rB, u);
        const float invLinMass = square(ratio) * (D.invM + B.invM);
        const float invRotMassD = D.invI * square(JwD * 1.0f / 1.0f), invRotMassB = B.invI * square(JwB * 1.0f / 1.0f);
        invMass += invLinMass + invRotMassD + invRotMassB;
        const V2 pD = V2{g.bd.local_anchor_a.x, g.bd.local_anchor_a.y} - D.lc;
        const V2 pB = inv_rotate(rB + (posB.l - posD.l), qD);
        coordinateB = dot(pB - pD, lax) / 1.0f;
    } else if (g.bd.type == PR_GEAR_REVOLUTE) {
        JwB = ratio; JwD = ratio;
        invMass += square(ratio) * (B.invI + D.invI);
        coordinateB = ((posB.a - posD.a) - g.bd.reference_angle) / 1.0f;
    }
    const float Cc = ((coordinateA + ratio * coordinateB) - constant) * c.resolutionRate;
    const float impulse = ((invMass > 0.0f) ? -Cc / invMass : 0.0f) * 1.0f * 1.0f;
    const Pos deltaA = cap_position(A.invM * impulse * JvAC, A.invI * impulse * JwA * 1.0f / 1.0f, c);
    const Pos deltaC = cap_position(C.invM * impulse * JvAC, C.invI * impulse * JwC * 1.0f / 1.0f, c);
    const Pos deltaB = cap_position(B.invM * impulse * JvBD, B.invI * impulse * JwB * 1.0f / 1.0f, c);
    const Pos deltaD = cap_position(D.invM * impulse * JvBD, D.invI * impulse * JwD * 1.0f / 1.0f, c);
    *A.pos = Pos{posA.l + deltaA.l, posA.a + deltaA.a};
    *B.pos = Pos{posB.l + deltaB.l, posB.a + deltaB.a};
    *C.pos = Pos{posC.l - deltaC.l, posC.a - deltaC.a};
    *D.pos = Pos{posD.l - deltaD.l, posD.a - deltaD.a};
    float linearError = 0.0f, angularError = 0.0f;
    linearError = std::max(linearError, magsq(deltaA.l));
    linearError = std::max(linearError, magsq(deltaB.l));
    linearError = std::max(linearError, magsq(deltaC.l));
    linearError = std::max(linearError, magsq(deltaD.l));
    angularError = std::max(angularError, std::fabs(deltaA.a));
    angularError = std::max(angularError, std::fabs(deltaB.a));
    angularError = std::max(angularError, std::fabs(deltaC.a));
    angularError = std::max(angularError, std::fabs(deltaD.a));
    return (linearError <= square(c.linearSlop)) && (angularError <= c.angularSlop);
}
void joint_init_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos posA, Vel& velA,
                         float imB, float iiB, V2 lcB, Pos posB, Vel& velB, const JointSolverConf& c)
{
    if (PR_JOINT_TYPE(j) == PR_JOINT_DISTANCE) distance_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c);
    else if (PR_JOINT_TYPE(j) == PR_JOINT_ROPE) rope_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c);
    else if (PR_JOINT_TYPE(j) == PR_JOINT_FRICTION || PR_JOINT_TYPE(j) == PR_JOINT_MOTOR)
        fm_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c, PR_JOINT_TYPE(j) == PR_JOINT_MOTOR);
    else if (PR_JOINT_TYPE(j) == PR_JOINT_TARGET) target_init_velocity(j, wk, imB, iiB, lcB, posB, velB, c);
    else if (PR_JOINT_TYPE(j) == PR_JOINT_PRISMATIC) prismatic_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c);
    else if (PR_JOINT_TYPE(j) == PR_JOINT_WHEEL) wheel_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c);
    else if (PR_JOINT_TYPE(j) == PR_JOINT_PULLEY) pulley_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c);
    else if (PR_JOINT_TYPE(j) == PR_JOINT_WELD) weld_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c);
    else revolute_init_velocity(j, wk, imA, iiA, lcA, posA, velA, imB, iiB, lcB, posB, velB, c);
}
bool joint_solve_velocity(pr_revolute_joint& j, RevoluteWork& wk, float imA, float iiA, Vel& velA, float imB, float iiB, Vel& velB, const JointSolverConf& c)
{
    if (PR_JOINT_TYPE(j) == PR_JOINT_DISTANCE) return distance_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB);
    if (PR_JOINT_TYPE(j) == PR_JOINT_ROPE) return rope_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_FRICTION || PR_JOINT_TYPE(j) == PR_JOINT_MOTOR)
        return fm_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB, c, PR_JOINT_TYPE(j) == PR_JOINT_MOTOR);
    if (PR_JOINT_TYPE(j) == PR_JOINT_TARGET) return target_solve_velocity(j, wk, imB, iiB, velB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_PRISMATIC) return prismatic_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_WHEEL) return wheel_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_PULLEY) return pulley_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB);
    if (PR_JOINT_TYPE(j) == PR_JOINT_WELD) return weld_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB);
    return revolute_solve_velocity(j, wk, imA, iiA, velA, imB, iiB, velB, c);
}
bool joint_solve_position(const pr_revolute_joint& j, const RevoluteWork& wk, float imA, float iiA, V2 lcA, Pos& posA,
                          float imB, float iiB, V2 lcB, Pos& posB, const JointSolverConf& c)
{
    if (PR_JOINT_TYPE(j) == PR_JOINT_DISTANCE) return distance_solve_position(j, wk, imA, iiA, lcA, posA, imB, iiB, lcB, posB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_ROPE) return rope_solve_position(j, wk, imA, iiA, lcA, posA, imB, iiB, lcB, posB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_FRICTION || PR_JOINT_TYPE(j) == PR_JOINT_MOTOR || PR_JOINT_TYPE(j) == PR_JOINT_TARGET) return true;
    if (PR_JOINT_TYPE(j) == PR_JOINT_PRISMATIC) return prismatic_solve_position(j, imA, iiA, lcA, posA, imB, iiB, lcB, posB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_WHEEL) return wheel_solve_position(j, wk, imA, iiA, lcA, posA, imB, iiB, lcB, posB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_PULLEY) return pulley_solve_position(j, imA, iiA, lcA, posA, imB, iiB, lcB, posB, c);
    if (PR_JOINT_TYPE(j) == PR_JOINT_WELD) return weld_solve_position(j, imA, iiA, lcA, posA, imB, iiB, lcB, posB, c);
    return revolute_solve_position(j, imA, iiA, lcA, posA, imB, iiB, lcB, posB, c);
}
}

#include "pr_oracle.h"
extern "C" {
int orc_toi(uint32_t n, const pr_child* children, const pr_vec2* verts, const pr_vec2* normals,
            const uint32_t* a_child, const float* sweep_a, const uint32_t* b_child, const float* sweep_b,
            const pr_step_conf* sc, float* out_time, int32_t* out_state, uint32_t* out_iters)
{
    using namespace orc;
    for (uint32_t i = 0; i < n; ++i) {
        const pr_child& ca = children[a_child[i]];
        const pr_child& cb = children[b_child[i]];
        const Proxy pa{ca.vertex_radius, ca.vertex_count, verts + ca.vertex_offset, normals + ca.vertex_offset};
        const Proxy pb{cb.vertex_radius, cb.vertex_count, verts + cb.vertex_offset, normals + cb.vertex_offset};
        const float* a = sweep_a + 8 * i;
        const float* b = sweep_b + 8 * i;
        const Sweep sa{{{a[0], a[1]}, a[2]}, {{a[3], a[4]}, a[5]}, {a[6], a[7]}, 0.0f};
        const Sweep sb{{{b[0], b[1]}, b[2]}, {{b[3], b[4]}, b[5]}, {b[6], b[7]}, 0.0f};
        const ToiOut o = toi_via_sat(pa, sa, pb, sb, *sc);
        out_time[i] = o.time; out_state[i] = o.state;
        out_iters[3 * i] = o.toi_iters; out_iters[3 * i + 1] = o.max_dist_iters; out_iters[3 * i + 2] = o.max_root_iters;
    }
    return 0;
}
void orc_sincosf_restated(uint32_t n, const float* x, float* cs)
{
    for (uint32_t i = 0; i < n; ++i) { cs[2 * i] = std::cos(x[i]); cs[2 * i + 1] = std::sin(x[i]); }
}
}
