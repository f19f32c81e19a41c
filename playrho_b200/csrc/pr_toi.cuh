// pr_toi.cuh -- device GJK distance and conservative-advancement time of impact.
// Re-expression, in registers and in the reference's operation order, of
//   Distance / TestOverlap        Library/source/playrho/d2/Distance.cpp:100-218
//   Simplex::Get / CalcMetric     Library/source/playrho/d2/Simplex.cpp:44-238
//   GetSeparationScenario/FindMinSeparation/Evaluate   SeparationScenario.cpp:29-217
//   GetToiViaSat                  Library/source/playrho/d2/TimeOfImpact.cpp:35-251
// One thread runs one whole query: the loops are short, data dependent and bounded by
// maxToiIters (20) x maxToiRootIters (30) x maxDistanceIters (20); vertices come from the pooled
// float2 arrays through the read-only path.
#ifndef PR_TOI_CUH
#define PR_TOI_CUH
#include "pr_collide.cuh"

namespace pr {

struct SEdge { float2 wA, wB; unsigned iA, iB; };     // SimplexEdge.hpp
PR_D float2 se_delta(const SEdge& e) { return e.wB - e.wA; } // GetPointDelta, SimplexEdge.hpp:84-87
struct GjkSimplex { SEdge e[3]; float coef[3]; unsigned n; };
// IndexPair3 packed: pair k in bits [16k, 16k+16) as (a | b << 8); 0xFFFF = InvalidIndexPair
struct GjkCache { float metric; unsigned long long idx; };
#define PR_IP_INVALID 0xFFFFull
#define PR_IP3_INVALID 0xFFFFFFFFFFFFull
PR_D unsigned ip_get(unsigned long long idx, unsigned k) { return static_cast<unsigned>((idx >> (16u * k)) & 0xFFFFull); }
PR_D unsigned ip_count(unsigned long long idx)
{ // GetNumValidIndices, IndexPair.hpp:66-72
    return (ip_get(idx, 0) != 0xFFFFu ? 1u : 0u) + (ip_get(idx, 1) != 0xFFFFu ? 1u : 0u) + (ip_get(idx, 2) != 0xFFFFu ? 1u : 0u);
}
PR_D unsigned long long ip_pack(const GjkSimplex& s)
{ // GetIndexPairs, Simplex.cpp:27-42
    unsigned long long r = PR_IP3_INVALID;
    for (unsigned k = 0; k < s.n; ++k) {
        const unsigned long long p = static_cast<unsigned long long>(s.e[k].iA | (s.e[k].iB << 8));
        r = (r & ~(0xFFFFull << (16u * k))) | (p << (16u * k));
    }
    return r;
}
PR_D SEdge se_make(const ProxyView& a, Xf xfA, unsigned iA, const ProxyView& b, Xf xfB, unsigned iB)
{ // GetSimplexEdge, Distance.cpp:37-44
    SEdge e; e.wA = transform(a.vertex(iA), xfA); e.iA = iA; e.wB = transform(b.vertex(iB), xfB); e.iB = iB; return e;
}
PR_D float gjk_metric(const GjkSimplex& s)
{ // Simplex::CalcMetric, Simplex.cpp:218-238
    if (s.n == 2u) return sqrtf(magsq(se_delta(s.e[1]) - se_delta(s.e[0])));
    if (s.n == 3u) return cross(se_delta(s.e[1]) - se_delta(s.e[0]), se_delta(s.e[2]) - se_delta(s.e[0]));
    return 0.0f;
}
PR_D float2 gjk_search_dir(const GjkSimplex& s)
{ // CalcSearchDirection, Simplex.cpp:44-62
    if (s.n == 1u) return -se_delta(s.e[0]);
    if (s.n == 2u) {
        const float2 e12 = se_delta(s.e[1]) - se_delta(s.e[0]);
        const float2 e0 = se_delta(s.e[0]);
        const float sgn = cross(e12, -e0);
        return (sgn > 0.0f) ? v2(-e12.y, e12.x) : v2(e12.y, -e12.x);
    }
    return v2(0.0f, 0.0f);
}
PR_D void gjk_set1(GjkSimplex& r, const SEdge& s) { r.e[0] = s; r.n = 1u; r.coef[0] = 1.0f; }
PR_D void gjk_set2(GjkSimplex& r, const SEdge& a, const SEdge& b, float ca, float cb)
{
    r.e[0] = a; r.e[1] = b; r.n = 2u; r.coef[0] = ca; r.coef[1] = cb;
}
// Simplex::Get(edges): Simplex.cpp:64-216 (1, 2 and 3 edge cases)
PR_D void gjk_solve(GjkSimplex& s)
{
    if (s.n == 2u) {
        const SEdge s0 = s.e[0], s1 = s.e[1];
        const float2 w1 = se_delta(s0), w2 = se_delta(s1);
        const float2 e12 = w2 - w1;
        const float d12_2 = -dot(w1, e12);
        if (d12_2 <= 0.0f) { gjk_set1(s, s0); return; }
        const float d12_1 = dot(w2, e12);
        if (d12_1 <= 0.0f) { gjk_set1(s, s1); return; }
        const float inv_sum = 1.0f / (d12_1 + d12_2);
        s.coef[0] = d12_1 * inv_sum; s.coef[1] = d12_2 * inv_sum;
        return;
    }
    if (s.n == 3u) {
        const SEdge s0 = s.e[0], s1 = s.e[1], s2 = s.e[2];
        const float2 w1 = se_delta(s0), w2 = se_delta(s1), w3 = se_delta(s2);
        const float2 e12 = w2 - w1;
        const float d12_1 = dot(w2, e12), d12_2 = -dot(w1, e12);
        const float2 e13 = w3 - w1;
        const float d13_1 = dot(w3, e13), d13_2 = -dot(w1, e13);
        const float2 e23 = w3 - w2;
        const float d23_1 = dot(w3, e23), d23_2 = -dot(w2, e23);
        if (d12_2 <= 0.0f && d13_2 <= 0.0f) { gjk_set1(s, s0); return; }
        if (d12_1 <= 0.0f && d23_2 <= 0.0f) { gjk_set1(s, s1); return; }
        if (d13_1 <= 0.0f && d23_1 <= 0.0f) { gjk_set1(s, s2); return; }
        const float n123 = cross(e12, e13);
        const float d123_3 = n123 * cross(w1, w2);
        if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
            const float inv_sum = 1.0f / (d12_1 + d12_2);
            gjk_set2(s, s0, s1, d12_1 * inv_sum, d12_2 * inv_sum); return;
        }
        const float d123_2 = n123 * cross(w3, w1);
        if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
            const float inv_sum = 1.0f / (d13_1 + d13_2);
            gjk_set2(s, s0, s2, d13_1 * inv_sum, d13_2 * inv_sum); return;
        }
        const float d123_1 = n123 * cross(w2, w3);
        if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
            const float inv_sum = 1.0f / (d23_1 + d23_2);
            gjk_set2(s, s2, s1, d23_2 * inv_sum, d23_1 * inv_sum); return; // swapped order: Simplex.cpp:193
        }
        const float inv_sum = 1.0f / (d123_1 + d123_2 + d123_3);
        s.coef[0] = d123_1 * inv_sum; s.coef[1] = d123_2 * inv_sum; s.coef[2] = d123_3 * inv_sum;
        return;
    }
    s.coef[0] = 1.0f;
}
PR_D unsigned support_index(const ProxyView& p, float2 dir)
{ // GetSupportIndex, DistanceProxy.hpp:232-249: first strict maximum
    unsigned index = PR_INVALID_VTX;
    float maxValue = -PR_INF;
    for (unsigned i = 0; i < p.count; ++i) {
        const float value = dot(p.vertex(i), dir);
        if (maxValue < value) { maxValue = value; index = i; }
    }
    return index;
}
enum { GJK_UNKNOWN = 0, GJK_UNFIT_SEARCH_DIR, GJK_HIT_MAX_ITERS, GJK_MAX_POINTS, GJK_DUPLICATE_INDEX_PAIR };

// Distance, Distance.cpp:100-203. Returns the state; `s` is the final simplex, `iters` the count.
PR_D int gjk_distance(const ProxyView& a, Xf xfA, const ProxyView& b, Xf xfB, const GjkCache& cache, unsigned maxIters,
                      GjkSimplex& s, unsigned& iters)
{
    unsigned long long saved = cache.idx;
    const unsigned count = ip_count(saved);
    for (unsigned k = 0; k < count; ++k) {
        const unsigned p = ip_get(saved, k);
        s.e[k] = se_make(a, xfA, p & 0xFFu, b, xfB, p >> 8);
    }
    s.n = count;
    if (s.n > 1u) {
        const float metric1 = cache.metric;
        const float metric2 = gjk_metric(s);
        if ((metric2 < (metric1 / 2)) || (metric2 > (metric1 * 2)) || (metric2 < 0) || almost_zero(metric2)) s.n = 0u;
    }
    if (s.n == 0u) {
        s.e[0] = se_make(a, xfA, 0u, b, xfB, 0u); s.n = 1u;
        saved = (PR_IP3_INVALID & ~0xFFFFull); // {(0,0), invalid, invalid}
    }
    int state = GJK_HIT_MAX_ITERS;
    unsigned iter = 0;
    while (iter < maxIters) {
        ++iter;
        gjk_solve(s);
        if (s.n == 3u) { state = GJK_MAX_POINTS; break; }
        const float2 dir = gjk_search_dir(s);
        if (almost_zero(magsq(dir))) { state = GJK_UNFIT_SEARCH_DIR; break; }
        const unsigned indexA = support_index(a, inv_rotate(-dir, xfA.q));
        const unsigned indexB = support_index(b, inv_rotate(dir, xfB.q));
        const unsigned key = indexA | (indexB << 8);
        if (ip_get(saved, 0) == key || ip_get(saved, 1) == key || ip_get(saved, 2) == key) { state = GJK_DUPLICATE_INDEX_PAIR; break; }
        s.e[s.n] = se_make(a, xfA, indexA, b, xfB, indexB);
        s.coef[s.n] = 0.0f;
        ++s.n;
        saved = ip_pack(s);
    }
    iters = iter;
    return state;
}
PR_D float gjk_witness_dist_sq(const GjkSimplex& s)
{ // GetWitnessPoints, Distance.cpp:72-94; |pointB - pointA|^2
    float2 pA = v2(0.0f, 0.0f), pB = v2(0.0f, 0.0f);
    for (unsigned i = 0; i < s.n; ++i) {
        pA = pA + s.e[i].wA * s.coef[i];
        pB = pB + s.e[i].wB * s.coef[i];
    }
    return magsq(pB - pA);
}
// TestOverlap, Distance.cpp:205-218 (sensor contacts): >= 0 means overlapping
PR_D float gjk_test_overlap(const ProxyView& a, Xf xfA, const ProxyView& b, Xf xfB, unsigned maxIters)
{
    GjkCache cache; cache.metric = 0.0f; cache.idx = PR_IP3_INVALID;
    GjkSimplex s; unsigned iters;
    gjk_distance(a, xfA, b, xfB, cache, maxIters, s, iters);
    return square(a.radius + b.radius) - gjk_witness_dist_sq(s);
}

// ---- SeparationScenario.cpp: type 0 = points, 1 = face A, 2 = face B
struct SepScenario { int type; Rot axis; float2 lp; };
PR_D SepScenario sep_scenario(unsigned long long idx, const ProxyView& a, Xf xfA, const ProxyView& b, Xf xfB)
{ // GetSeparationScenario, SeparationScenario.cpp:138-190
    SepScenario r;
    const Rot zero = mkrot(0.0f, 0.0f);
    const unsigned n = ip_count(idx);
    const unsigned p0 = ip_get(idx, 0), p1 = ip_get(idx, 1);
    if (n >= 2u && (p0 & 0xFFu) == (p1 & 0xFFu)) {
        const float2 b1 = b.vertex(p0 >> 8), b2 = b.vertex(p1 >> 8);
        const float2 e = b2 - b1;
        const Rot axis = unit_vector(v2(e.y, -e.x), zero);
        const Rot normal = rot_rotate(axis, xfB.q);
        r.lp = vdiv(b1 + b2, 2.0f);
        const float2 pointB = transform(r.lp, xfB);
        const float2 pointA = transform(a.vertex(p0 & 0xFFu), xfA);
        r.type = 2; r.axis = (dot(pointA - pointB, normal) < 0.0f) ? rneg(axis) : axis;
        return r;
    }
    if (n >= 2u) {
        const float2 a1 = a.vertex(p0 & 0xFFu), a2 = a.vertex(p1 & 0xFFu);
        const float2 e = a2 - a1;
        const Rot axis = unit_vector(v2(e.y, -e.x), zero);
        const Rot normal = rot_rotate(axis, xfA.q);
        r.lp = vdiv(a1 + a2, 2.0f);
        const float2 pointA = transform(r.lp, xfA);
        const float2 pointB = transform(b.vertex(p0 >> 8), xfB);
        r.type = 1; r.axis = (dot(pointB - pointA, normal) < 0.0f) ? rneg(axis) : axis;
        return r;
    }
    const float2 pointA = transform(a.vertex(p0 & 0xFFu), xfA);
    const float2 pointB = transform(b.vertex(p0 >> 8), xfB);
    r.type = 0; r.axis = unit_vector(pointB - pointA, zero); r.lp = v2(0.0f, 0.0f);
    return r;
}
PR_D float sep_find_min(const SepScenario& s, const ProxyView& a, Xf xfA, const ProxyView& b, Xf xfB, unsigned& iA, unsigned& iB)
{ // FindMinSeparation, SeparationScenario.cpp:29-74
    if (s.type == 0) {
        const float2 ax = rotv(s.axis);
        iA = support_index(a, inv_rotate(ax, xfA.q));
        iB = support_index(b, inv_rotate(-ax, xfB.q));
        return dot(transform(b.vertex(iB), xfB) - transform(a.vertex(iA), xfA), s.axis);
    }
    if (s.type == 1) {
        const Rot normal = rot_rotate(s.axis, xfA.q);
        const float2 pA = transform(s.lp, xfA);
        iA = PR_INVALID_VTX;
        iB = support_index(b, inv_rotate(-rotv(normal), xfB.q));
        return dot(transform(b.vertex(iB), xfB) - pA, normal);
    }
    const Rot normal = rot_rotate(s.axis, xfB.q);
    iA = support_index(a, inv_rotate(-rotv(normal), xfA.q));
    iB = PR_INVALID_VTX;
    const float2 pA = transform(a.vertex(iA), xfA);
    return dot(pA - transform(s.lp, xfB), normal);
}
PR_D float sep_evaluate(const SepScenario& s, const ProxyView& a, Xf xfA, const ProxyView& b, Xf xfB, unsigned iA, unsigned iB)
{ // Evaluate, SeparationScenario.cpp:76-114
    if (s.type == 0) return dot(transform(b.vertex(iB), xfB) - transform(a.vertex(iA), xfA), s.axis);
    if (s.type == 1) {
        const Rot normal = rot_rotate(s.axis, xfA.q);
        const float2 pA = transform(s.lp, xfA);
        return dot(transform(b.vertex(iB), xfB) - pA, normal);
    }
    const Rot normal = rot_rotate(s.axis, xfB.q);
    const float2 pB = transform(s.lp, xfB);
    return dot(transform(a.vertex(iA), xfA) - pB, normal);
}

struct SweepD { Pos pos0, pos1; float2 lc; };
PR_D Pos lerp_pos(Pos p0, Pos p1, float beta)
{ // GetPosition(pos0, pos1, beta), d2/Math.cpp:30-60
    Pos r; r.l = p0.l + (p1.l - p0.l) * beta; r.a = p0.a + (p1.a - p0.a) * beta; return r;
}
PR_D Xf sweep_xf(const SweepD& s, float beta) { return get_xf(lerp_pos(s.pos0, s.pos1, beta), s.lc); }

// ToiOutput::State, ToiOutput.hpp:73-140
enum { TOI_UNKNOWN = 0, TOI_TOUCHING, TOI_SEPARATED, TOI_OVERLAPPED, TOI_MAX_ROOT_ITERS, TOI_NEXT_AFTER, TOI_MAX_TOI_ITERS,
       TOI_BELOW_MIN_TARGET, TOI_MAX_DIST_ITERS, TOI_TARGET_DEPTH_EXCEEDS_TOTAL_RADIUS, TOI_MIN_TARGET_SQ_OVERFLOW,
       TOI_MAX_TARGET_SQ_OVERFLOW, TOI_NOT_FINITE };
struct ToiResult { float time; int state; unsigned toiIters, maxDistIters, maxRootIters; };
PR_D bool pr_isfinite(float v) { return ((f2u(v) >> 23) & 0xFFu) != 0xFFu; }

// GetToiViaSat, TimeOfImpact.cpp:35-251 with GetToiConf(StepConf) (ToiConf.cpp:26-35: timeMax = 1)
PR_D ToiResult toi_via_sat(const ProxyView& a, const SweepD& sA, const ProxyView& b, const SweepD& sB, float targetDepth,
                           float tolerance, unsigned maxToiIters, unsigned maxRootIters, unsigned maxDistIters)
{
    ToiResult out; out.time = 0.0f; out.state = TOI_UNKNOWN; out.toiIters = out.maxDistIters = out.maxRootIters = 0u;
    const float totalRadius = a.radius + b.radius;
    if (targetDepth > totalRadius) { out.state = TOI_TARGET_DEPTH_EXCEEDS_TOTAL_RADIUS; return out; }
    const float target = totalRadius - targetDepth;
    const float maxTarget = stdmax(target + tolerance, 0.0f);
    const float minTarget = stdmax(target - tolerance, 0.0f);
    const float minTargetSquared = square(minTarget);
    if (!pr_isfinite(minTargetSquared) && pr_isfinite(minTarget)) { out.state = TOI_MIN_TARGET_SQ_OVERFLOW; return out; }
    const float maxTargetSquared = square(maxTarget);
    if (!pr_isfinite(maxTargetSquared) && pr_isfinite(maxTarget)) { out.state = TOI_MAX_TARGET_SQ_OVERFLOW; return out; }
    float timeLo = 0.0f;
    Xf loA = sweep_xf(sA, timeLo), loB = sweep_xf(sB, timeLo);
    GjkCache cache; cache.metric = 0.0f; cache.idx = PR_IP3_INVALID;
    while (out.toiIters < maxToiIters) {
        GjkSimplex sx; unsigned dIters;
        const int dstate = gjk_distance(a, loA, b, loB, cache, maxDistIters, sx, dIters);
        ++out.toiIters;
        out.maxDistIters = (out.maxDistIters < dIters) ? dIters : out.maxDistIters;
        if (dstate == GJK_HIT_MAX_ITERS) { out.time = timeLo; out.state = TOI_MAX_DIST_ITERS; return out; }
        cache.metric = gjk_metric(sx);
        cache.idx = ip_pack(sx);
        const float distSquared = gjk_witness_dist_sq(sx);
        if (distSquared < minTargetSquared) { out.time = timeLo; out.state = TOI_OVERLAPPED; return out; }
        if (distSquared <= maxTargetSquared) { out.time = timeLo; out.state = TOI_TOUCHING; return out; }
        const SepScenario fcn = sep_scenario(cache.idx, a, loA, b, loB);
        float timeHi = 1.0f;
        Xf hiA = sweep_xf(sA, timeHi), hiB = sweep_xf(sB, timeHi);
        for (unsigned pbIter = 0; pbIter < 254u; ++pbIter) { // MaxShapeVertices
            unsigned iA, iB;
            const float hiMin = sep_find_min(fcn, a, hiA, b, hiB, iA, iB);
            if (hiMin > maxTarget) { out.time = timeHi; out.state = TOI_SEPARATED; return out; }
            if (hiMin >= minTarget) {
                if (timeHi == timeLo) { out.time = timeHi; out.state = TOI_SEPARATED; return out; }
                timeLo = timeHi; loA = hiA; loB = hiB;
                break;
            }
            const float loEval = sep_evaluate(fcn, a, loA, b, loB, iA, iB);
            if (loEval <= maxTarget) { out.time = timeLo; out.state = (loEval < minTarget) ? TOI_BELOW_MIN_TARGET : TOI_TOUCHING; return out; }
            float a1 = timeLo, a2 = timeHi, s1 = loEval, s2 = hiMin;
            unsigned roots = 0;
            float t = timeLo;
            for (;;) {
                if (roots == maxRootIters) {
                    out.maxRootIters = (out.maxRootIters < roots) ? roots : out.maxRootIters;
                    out.time = t; out.state = TOI_MAX_ROOT_ITERS; return out;
                }
                if (nextafterf(a1, a2) >= a2) {
                    out.maxRootIters = (out.maxRootIters < roots) ? roots : out.maxRootIters;
                    out.time = t; out.state = TOI_NEXT_AFTER; return out;
                }
                t = (roots & 1u) ? (a1 + (target - s1) * (a2 - a1) / (s2 - s1)) : ((a1 + a2) / 2); // Secant / Bisect, Math.hpp:92-108
                ++roots;
                const Xf tA = sweep_xf(sA, t), tB = sweep_xf(sB, t);
                const float s = sep_evaluate(fcn, a, tA, b, tB, iA, iB);
                if (fabsf(s - target) <= tolerance) { timeHi = t; hiA = tA; hiB = tB; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
            }
            out.maxRootIters = (out.maxRootIters < roots) ? roots : out.maxRootIters;
        }
    }
    out.time = timeLo; out.state = TOI_MAX_TOI_ITERS;
    return out;
}

} // namespace pr
#endif
