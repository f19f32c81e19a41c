// orc_collide.cpp -- TEST INFRASTRUCTURE ONLY (oracle restatement of the narrowphase).
// Restates CollideShapes and what it calls:
//   Library/source/playrho/d2/Manifold.cpp:61-438, ShapeSeparation.cpp:39-128.
#include "orc_math.h"
#include <cstring>

namespace orc {

namespace {

constexpr uint32_t kInvalidVertex = 255u;
constexpr uint32_t kFace = PR_CF_FACE, kVertex = PR_CF_VERTEX;

pr_manifold manifold_unset()
{
    pr_manifold m;
    std::memset(&m, 0, sizeof(m));
    m.cf0 = m.cf1 = CF_DEFAULT;
    return m;
}

// Manifold::GetForCircles, Manifold.hpp:168-171
pr_manifold for_circles(V2 vA, uint32_t iA, V2 vB, uint32_t iB)
{
    pr_manifold m = manifold_unset();
    m.type = PR_MANIFOLD_CIRCLES;
    m.lp_x = vA.x; m.lp_y = vA.y;
    m.p0_x = vB.x; m.p0_y = vB.y;
    m.cf0 = cf_pack(kVertex, iA, kVertex, iB);
    m.count = 1;
    return m;
}

// Manifold::GetForFaceA/B with 0 or 1 points, Manifold.hpp:178-228
pr_manifold for_face(uint32_t type, Rot ln, V2 lp)
{
    pr_manifold m = manifold_unset();
    m.type = type;
    m.ln_x = ln.c; m.ln_y = ln.s;
    m.lp_x = lp.x; m.lp_y = lp.y;
    return m;
}

// Manifold::AddPoint, Manifold.hpp:461-478
void add_point(pr_manifold& m, V2 lp, pr_contact_feature cf)
{
    if (m.count == 0) { m.p0_x = lp.x; m.p0_y = lp.y; m.cf0 = cf; }
    else { m.p1_x = lp.x; m.p1_y = lp.y; m.cf1 = cf; }
    ++m.count;
}

pr_manifold for_face(uint32_t type, Rot ln, V2 lp, V2 pt, pr_contact_feature cf)
{
    pr_manifold m = for_face(type, ln, lp);
    add_point(m, pt, cf);
    return m;
}

struct SepInfo { float distance; uint32_t first; uint32_t second0, second1; };
struct LenIdx { float distance; uint32_t first, second; };

// GetMinSeparationInfo, ShapeSeparation.cpp:39-63
LenIdx min_separation_info(V2 origin, Rot direction, const V2* verts, uint32_t n)
{
    float minSep = std::numeric_limits<float>::infinity();
    uint32_t first = kInvalidVertex, second = kInvalidVertex;
    for (uint32_t i = 0; i < n; ++i) {
        const float s = dot(direction, verts[i] - origin);
        if (minSep > s) { minSep = s; first = i; second = kInvalidVertex; }
        else if (minSep == s) { second = i; }
    }
    return {minSep, first, second};
}

// GetMaxSeparation4x4, ShapeSeparation.cpp:66-103
SepInfo max_separation_4x4(const Proxy& p1, Xf xf1, const Proxy& p2, Xf xf2)
{
    float separation = -std::numeric_limits<float>::infinity();
    uint32_t firstIndex = kInvalidVertex, s0 = kInvalidVertex, s1 = kInvalidVertex;
    const Xf xf = mulT(xf1, xf2);
    const V2 p2v[4] = {transform(p2.vertex(0), xf), transform(p2.vertex(1), xf),
                       transform(p2.vertex(2), xf), transform(p2.vertex(3), xf)};
    for (uint32_t i = 0; i < 4; ++i) {
        const LenIdx ap = min_separation_info(p1.vertex(i), p1.normal(i), p2v, 4);
        if (separation < ap.distance) {
            separation = ap.distance; s0 = ap.first; s1 = ap.second; firstIndex = i;
        }
    }
    return {separation, firstIndex, s0, s1};
}

// GetMaxSeparation, ShapeSeparation.cpp:106-128
SepInfo max_separation(const Proxy& p1, Xf xf1, const Proxy& p2, Xf xf2)
{
    float separation = -std::numeric_limits<float>::infinity();
    uint32_t firstIndex = kInvalidVertex, s0 = kInvalidVertex, s1 = kInvalidVertex;
    const Xf xf = mulT(xf2, xf1);
    V2 verts2[254];
    for (uint32_t j = 0; j < p2.count; ++j) verts2[j] = p2.vertex(j);
    for (uint32_t i = 0; i < p1.count; ++i) {
        const V2 origin = transform(p1.vertex(i), xf);
        const Rot normal = rot_rotate(p1.normal(i), xf.q);
        const LenIdx ap = min_separation_info(origin, normal, verts2, p2.count);
        if (separation < ap.distance) {
            separation = ap.distance; s0 = ap.first; s1 = ap.second; firstIndex = i;
        }
    }
    return {separation, firstIndex, s0, s1};
}

struct ClipVertex { V2 v; pr_contact_feature cf; };
struct ClipList { ClipVertex e[2]; uint32_t n = 0; void push(ClipVertex c) { e[n++] = c; } };

// ClipSegmentToLine, Manifold.cpp:61-101
ClipList clip_segment_to_line(const ClipList& vIn, Rot normal, float offset, uint32_t indexA)
{
    ClipList vOut;
    if (vIn.n == 2) {
        const float d0 = dot(normal, vIn.e[0].v) - offset;
        const float d1 = dot(normal, vIn.e[1].v) - offset;
        if (d0 <= 0.0f || almost_zero(d0)) vOut.push(vIn.e[0]);
        if (d1 <= 0.0f || almost_zero(d1)) vOut.push(vIn.e[1]);
        if (vOut.n < 2 && std::signbit(d0) != std::signbit(d1)) {
            const float interp = d0 / (d0 - d1);
            const V2 vertex = vIn.e[0].v + (vIn.e[1].v - vIn.e[0].v) * interp;
            const uint32_t idxB = (vIn.e[0].cf >> 24) & 0xFFu;
            vOut.push({vertex, cf_pack(kVertex, indexA, kFace, idxB)});
        }
    }
    return vOut;
}

struct IdxPair { uint32_t first, second; };

// GetMostAntiParallelEdge, Manifold.cpp:122-141
IdxPair most_anti_parallel_edge(Rot shape0_rel_n0, Xf xf0, const Proxy& shape1, Xf xf1,
                                uint32_t firstIdx, uint32_t secondIdx)
{
    if (secondIdx == kInvalidVertex) {
        const Rot normal = rot_inv_rotate(rot_rotate(shape0_rel_n0, xf0.q), xf1.q);
        const uint32_t count = shape1.count;
        const uint32_t prevIdx = mod_prev(firstIdx, count);
        const float prevDot = dot(normal, shape1.normal(prevIdx));
        const float currDot = dot(normal, shape1.normal(firstIdx));
        return (prevDot < currDot) ? IdxPair{prevIdx, firstIdx}
                                   : IdxPair{firstIdx, mod_next(firstIdx, count)};
    }
    return ((secondIdx > firstIdx) && ((firstIdx + 1) == secondIdx)) ? IdxPair{firstIdx, secondIdx}
                                                                    : IdxPair{secondIdx, firstIdx};
}

// GetClipPoints, Manifold.cpp:143-157
ClipList get_clip_points(V2 s0v0, V2 s0v1, IdxPair s0e, Rot s0dir, V2 s1v0, V2 s1v1, IdxPair s1e)
{
    ClipList ie;
    ie.push({s1v0, cf_pack(kFace, s0e.first, kVertex, s1e.first)});
    ie.push({s1v1, cf_pack(kFace, s0e.first, kVertex, s1e.second)});
    const float dp0 = -dot(s0dir, s0v0);
    const float dp1 = +dot(s0dir, s0v1);
    const ClipList points = clip_segment_to_line(ie, -s0dir, dp0, s0e.first);
    return clip_segment_to_line(points, s0dir, dp1, s0e.second);
}

// GetManifold(flipped, shape0, xf0, idx0, shape1, xf1, indices1, conf), Manifold.cpp:168-291
pr_manifold manifold_poly_poly(bool flipped, const Proxy& shape0, Xf xf0, uint32_t idx0,
                               const Proxy& shape1, Xf xf1, uint32_t ind1a, uint32_t ind1b,
                               float maxCirclesRatio)
{
    const float r0 = shape0.radius, r1 = shape1.radius;
    const float totalRadius = r0 + r1;
    const uint32_t idx0Next = mod_next(idx0, shape0.count);
    const V2 s0rv0 = shape0.vertex(idx0), s0rv1 = shape0.vertex(idx0Next);
    const V2 s0av0 = transform(s0rv0, xf0), s0av1 = transform(s0rv1, xf0);
    const float lenEdge0 = magsq(s0rv1 - s0rv0);
    const Rot s0n0 = shape0.normal(idx0);
    const Rot s0dir = rev_perp(s0n0);
    const IdxPair s1e = most_anti_parallel_edge(s0n0, xf0, shape1, xf1, ind1a, ind1b);
    const V2 s1rv0 = shape1.vertex(s1e.first);
    const V2 s1av0 = transform(s1rv0, xf1);
    const V2 s1rv1 = shape1.vertex(s1e.second);
    const V2 s1av1 = transform(s1rv1, xf1);
    const uint32_t faceType = flipped ? PR_MANIFOLD_FACE_B : PR_MANIFOLD_FACE_A;
    {
        const Rot absDir = rot_rotate(s0dir, xf0.q);
        const ClipList cps = get_clip_points(s0av0, s0av1, {idx0, idx0Next}, absDir, s1av0, s1av1, s1e);
        if (cps.n == 2) {
            const Rot absNormal = fwd_perp(absDir);
            const V2 relMid = (s0rv0 + s0rv1) / 2.0f;
            const float absOffset = dot(absNormal, s0av0);
            const Rot normal = fwd_perp(s0dir);
            pr_manifold m = for_face(faceType, normal, relMid);
            for (uint32_t k = 0; k < 2; ++k) {
                const ClipVertex& cp = cps.e[k];
                if ((dot(absNormal, cp.v) - absOffset) <= totalRadius) {
                    add_point(m, inv_transform(cp.v, xf1), flipped ? cf_flip(cp.cf) : cp.cf);
                }
            }
            if (m.count > 0) return m;
        }
    }
    const float totalRadiusSq = square(totalRadius);
    const bool mustFace = lenEdge0 > square(maxCirclesRatio * r0);
    const Rot fn = fwd_perp(s0dir);
    struct Try { V2 a0; V2 a1; V2 r0; uint32_t i0; V2 r1; uint32_t i1; };
    const Try tries[4] = {
        {s0av0, s1av0, s0rv0, idx0, s1rv0, s1e.first},
        {s0av1, s1av1, s0rv1, idx0Next, s1rv1, s1e.second},
        {s0av0, s1av1, s0rv0, idx0, s1rv1, s1e.second},
        {s0av1, s1av0, s0rv1, idx0Next, s1rv0, s1e.first},
    };
    for (const Try& t : tries) {
        if (magsq(t.a0 - t.a1) <= totalRadiusSq) {
            if (mustFace) {
                return !flipped
                    ? for_face(PR_MANIFOLD_FACE_A, fn, t.r0, t.r1, cf_pack(kFace, t.i0, kVertex, t.i1))
                    : for_face(PR_MANIFOLD_FACE_B, fn, t.r0, t.r1, cf_pack(kVertex, t.i1, kFace, t.i0));
            }
            return !flipped ? for_circles(t.r0, t.i0, t.r1, t.i1) : for_circles(t.r1, t.i1, t.r0, t.i0);
        }
    }
    return manifold_unset();
}

// GetManifold(flipped, totalRadius, shape, sxf, point, xfm), Manifold.cpp:293-372
pr_manifold manifold_poly_point(bool flipped, float totalRadius, const Proxy& shape, Xf sxf, V2 point,
                                Xf xfm)
{
    const uint32_t vertexCount = shape.count;
    const V2 cLocal = inv_transform(transform(point, xfm), sxf);
    uint32_t indexOfMax = 0;
    float maxSeparation = -FLT_MAX;
    for (uint32_t i = 0; i < vertexCount; ++i) {
        const float s = dot(shape.normal(i), cLocal - shape.vertex(i));
        if (s > totalRadius) return manifold_unset();
        if (maxSeparation < s) { maxSeparation = s; indexOfMax = i; }
    }
    const uint32_t indexOfMax2 = mod_next(indexOfMax, vertexCount);
    const V2 v1 = shape.vertex(indexOfMax), v2 = shape.vertex(indexOfMax2);
    const uint32_t faceType = flipped ? PR_MANIFOLD_FACE_B : PR_MANIFOLD_FACE_A;
    const pr_contact_feature faceCf = !flipped ? cf_pack(kFace, indexOfMax, kVertex, 0)
                                               : cf_pack(kVertex, 0, kFace, indexOfMax);
    if (maxSeparation < 0.0f) {
        const V2 faceCenter = (v1 + v2) / 2.0f;
        return for_face(faceType, shape.normal(indexOfMax), faceCenter, point, faceCf);
    }
    const V2 cLocalV1 = cLocal - v1;
    if (dot(cLocalV1, v2 - v1) <= 0.0f) {
        if (magsq(cLocalV1) > square(totalRadius)) return manifold_unset();
        return !flipped ? for_circles(v1, indexOfMax, point, 0) : for_circles(point, 0, v1, indexOfMax);
    }
    const V2 cLocalV2 = cLocal - v2;
    if (dot(cLocalV2, v1 - v2) <= 0.0f) {
        if (magsq(cLocalV2) > square(totalRadius)) return manifold_unset();
        return !flipped ? for_circles(v2, indexOfMax2, point, 0) : for_circles(point, 0, v2, indexOfMax2);
    }
    const V2 faceCenter = (v1 + v2) / 2.0f;
    if (dot(cLocal - faceCenter, shape.normal(indexOfMax)) > totalRadius) return manifold_unset();
    return for_face(faceType, shape.normal(indexOfMax), faceCenter, point, faceCf);
}

} // namespace

// CollideShapes, Manifold.cpp:393-438
pr_manifold collide_shapes(const Proxy& A, Xf xfA, const Proxy& B, Xf xfB, float linearSlop,
                           float maxCirclesRatio)
{
    const float totalRadius = A.radius + B.radius;
    const uint32_t countA = A.count, countB = B.count;
    if (countA == 1 && countB == 1) {
        // GetManifold(locationA, xfA, locationB, xfB, totalRadius), Manifold.cpp:374-386
        const V2 pA = transform(A.vertex(0), xfA), pB = transform(B.vertex(0), xfB);
        const float lenSq = magsq(pB - pA);
        const float totSq = square(totalRadius);
        if (lenSq > totSq) return manifold_unset();
        return for_circles(A.vertex(0), 0, B.vertex(0), 0);
    }
    if (countA == 1) return manifold_poly_point(true, totalRadius, B, xfB, A.vertex(0), xfA);
    if (countB == 1) return manifold_poly_point(false, totalRadius, A, xfA, B.vertex(0), xfB);
    const bool do4x4 = (countA == 4) && (countB == 4);
    const SepInfo sepA = do4x4 ? max_separation_4x4(A, xfA, B, xfB) : max_separation(A, xfA, B, xfB);
    if (sepA.distance > totalRadius) return manifold_unset();
    const SepInfo sepB = do4x4 ? max_separation_4x4(B, xfB, A, xfA) : max_separation(B, xfB, A, xfA);
    if (sepB.distance > totalRadius) return manifold_unset();
    const float k_tol = linearSlop / 10;
    return (sepB.distance > (sepA.distance + k_tol))
        ? manifold_poly_poly(true, B, xfB, sepB.first, A, xfA, sepB.second0, sepB.second1, maxCirclesRatio)
        : manifold_poly_poly(false, A, xfA, sepA.first, B, xfB, sepA.second0, sepA.second1, maxCirclesRatio);
}

} // namespace orc
