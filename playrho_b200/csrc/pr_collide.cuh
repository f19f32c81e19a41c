// pr_collide.cuh -- device narrowphase: CollideShapes over two DistanceProxy children.
// Re-expression of Library/source/playrho/d2/Manifold.cpp:61-438 and ShapeSeparation.cpp:39-128
// (SAT max separation both ways, reference/incident edge clipping, rounded-corner fallback,
// circle cases). Result is the 64-byte device manifold row (4 x float4, see pr_world.h).
#ifndef PR_COLLIDE_CUH
#define PR_COLLIDE_CUH
#include "pr_math.cuh"

namespace pr {

struct ProxyView {
    float radius;
    unsigned count;
    const float2* v;
    const float2* n;
    PR_D float2 vertex(unsigned i) const { return v[i]; }
    PR_D Rot normal(unsigned i) const { const float2 t = n[i]; return mkrot(t.x, t.y); }
};
// (Measured and rejected on B200, 4.4M contacts: staging both polygons in shared memory before the separation
// search -- narrowphase 0.87 -> 1.18 ms; staging them in registers with unrolled 8x8 loops -- 0.90 -> 1.07 ms;
// binning contacts by shape-pair class through a radix pass -- 0.90 -> 1.02 ms. What did pay: stopping the search
// at the first separating edge (0.87 -> 0.80) and not re-storing an empty manifold over an empty one (-> 0.75).)

// manifold in registers; stored as 4 x float4: (ln, lp) (p0, p1) (jn0 jt0 jn1 jt1) (cf0 cf1 type count)
struct Mf {
    float2 ln, lp, p0, p1;
    float jn0, jt0, jn1, jt1;
    unsigned cf0, cf1, type, count;
};

enum { MF_UNSET = 0, MF_CIRCLES = 1, MF_FACE_A = 2, MF_FACE_B = 3 };
enum { CF_VERTEX = 0, CF_FACE = 1 };
#define PR_INVALID_VTX 255u

PR_D Mf mf_unset()
{
    Mf m;
    m.ln = v2(0, 0); m.lp = v2(0, 0); m.p0 = v2(0, 0); m.p1 = v2(0, 0);
    m.jn0 = m.jt0 = m.jn1 = m.jt1 = 0.0f;
    m.cf0 = m.cf1 = PR_CF_DEFAULT; m.type = MF_UNSET; m.count = 0;
    return m;
}
PR_D Mf mf_circles(float2 vA, unsigned iA, float2 vB, unsigned iB)
{
    Mf m = mf_unset();
    m.type = MF_CIRCLES; m.lp = vA; m.p0 = vB; m.cf0 = cf_pack(CF_VERTEX, iA, CF_VERTEX, iB); m.count = 1;
    return m;
}
PR_D Mf mf_face(unsigned type, Rot ln, float2 lp)
{
    Mf m = mf_unset();
    m.type = type; m.ln = rotv(ln); m.lp = lp;
    return m;
}
PR_D void mf_add(Mf& m, float2 lp, unsigned cf)
{
    if (m.count == 0) { m.p0 = lp; m.cf0 = cf; } else { m.p1 = lp; m.cf1 = cf; }
    ++m.count;
}
PR_D Mf mf_face1(unsigned type, Rot ln, float2 lp, float2 pt, unsigned cf)
{
    Mf m = mf_face(type, ln, lp);
    mf_add(m, pt, cf);
    return m;
}

struct SepInfo { float distance; unsigned first, second0, second1; };

// GetMaxSeparation4x4, ShapeSeparation.cpp:66-103 (shape-2 vertices moved into shape-1's frame)
// stopAbove: CollideShapes only asks whether the max separation exceeds totalRadius (Manifold.cpp:419-427) before it
// looks at the indices, so the search may stop at the first edge whose separation does: the caller's outcome
// (an unset manifold) is the same as after the full search.
PR_D SepInfo max_separation_4x4(const ProxyView& p1, Xf xf1, const ProxyView& p2, Xf xf2, float stopAbove)
{
    SepInfo r; r.distance = -PR_INF; r.first = r.second0 = r.second1 = PR_INVALID_VTX;
    const Xf xf = mulT(xf1, xf2);
    float2 p2v[4];
#pragma unroll
    for (unsigned j = 0; j < 4; ++j) p2v[j] = transform(p2.vertex(j), xf);
#pragma unroll
    for (unsigned i = 0; i < 4; ++i) {
        const float2 origin = p1.vertex(i);
        const Rot normal = p1.normal(i);
        float minSep = PR_INF; unsigned first = PR_INVALID_VTX, second = PR_INVALID_VTX;
#pragma unroll
        for (unsigned j = 0; j < 4; ++j) {
            const float s = dot(normal, p2v[j] - origin);
            if (minSep > s) { minSep = s; first = j; second = PR_INVALID_VTX; }
            else if (minSep == s) { second = j; }
        }
        if (r.distance < minSep) { r.distance = minSep; r.second0 = first; r.second1 = second; r.first = i; }
        if (minSep > stopAbove) return r;
    }
    return r;
}

// GetMaxSeparation, ShapeSeparation.cpp:106-128 (shape-1 vertex+normal moved into shape-2's frame)
PR_D SepInfo max_separation(const ProxyView& p1, Xf xf1, const ProxyView& p2, Xf xf2, float stopAbove)
{
    // (staging both polygons in registers with fully unrolled, predicated 8x8 loops was measured SLOWER on B200:
    // narrowphase 0.90 -> 1.07 ms at 4.4M contacts -- every pair pays for 64 predicated iterations and 92 registers)
    SepInfo r; r.distance = -PR_INF; r.first = r.second0 = r.second1 = PR_INVALID_VTX;
    const Xf xf = mulT(xf2, xf1);
    const unsigned c1 = p1.count, c2 = p2.count;
    for (unsigned i = 0; i < c1; ++i) {
        const float2 origin = transform(p1.vertex(i), xf);
        const Rot normal = rot_rotate(p1.normal(i), xf.q);
        float minSep = PR_INF; unsigned first = PR_INVALID_VTX, second = PR_INVALID_VTX;
        for (unsigned j = 0; j < c2; ++j) {
            const float s = dot(normal, p2.vertex(j) - origin);
            if (minSep > s) { minSep = s; first = j; second = PR_INVALID_VTX; }
            else if (minSep == s) { second = j; }
        }
        if (r.distance < minSep) { r.distance = minSep; r.second0 = first; r.second1 = second; r.first = i; }
        if (minSep > stopAbove) return r;
    }
    return r;
}

struct ClipV { float2 v; unsigned cf; };

// ClipSegmentToLine, Manifold.cpp:61-101. in/out: two-slot lists with counts.
PR_D unsigned clip_segment(const ClipV* vin, unsigned nin, Rot normal, float offset, unsigned indexA, ClipV* vout)
{
    unsigned n = 0;
    if (nin == 2) {
        const float d0 = dot(normal, vin[0].v) - offset;
        const float d1 = dot(normal, vin[1].v) - offset;
        if (d0 <= 0.0f || almost_zero(d0)) vout[n++] = vin[0];
        if (d1 <= 0.0f || almost_zero(d1)) vout[n++] = vin[1];
        const bool sb0 = (f2u(d0) >> 31) != 0, sb1 = (f2u(d1) >> 31) != 0;
        if (n < 2 && sb0 != sb1) {
            const float interp = d0 / (d0 - d1);
            ClipV c;
            c.v = vin[0].v + (vin[1].v - vin[0].v) * interp;
            c.cf = cf_pack(CF_VERTEX, indexA, CF_FACE, (vin[0].cf >> 24) & 0xFFu);
            vout[n++] = c;
        }
    }
    return n;
}

// GetManifold(flipped, shape0, xf0, idx0, shape1, xf1, indices1, conf), Manifold.cpp:168-291
PR_D Mf manifold_poly_poly(bool flipped, const ProxyView& shape0, Xf xf0, unsigned idx0,
                           const ProxyView& shape1, Xf xf1, unsigned ind1a, unsigned ind1b, float maxCirclesRatio)
{
    const float r0 = shape0.radius, r1 = shape1.radius;
    const float totalRadius = r0 + r1;
    const unsigned idx0Next = mod_next(idx0, shape0.count);
    const float2 s0rv0 = shape0.vertex(idx0), s0rv1 = shape0.vertex(idx0Next);
    const float2 s0av0 = transform(s0rv0, xf0), s0av1 = transform(s0rv1, xf0);
    const float lenEdge0 = magsq(s0rv1 - s0rv0);
    const Rot s0n0 = shape0.normal(idx0);
    const Rot s0dir = rev_perp(s0n0);
    // GetMostAntiParallelEdge, Manifold.cpp:122-141
    unsigned e1first, e1second;
    if (ind1b == PR_INVALID_VTX) {
        const Rot normal = rot_inv_rotate(rot_rotate(s0n0, xf0.q), xf1.q);
        const unsigned count = shape1.count;
        const unsigned prevIdx = mod_prev(ind1a, count);
        const float prevDot = dot(normal, shape1.normal(prevIdx));
        const float currDot = dot(normal, shape1.normal(ind1a));
        if (prevDot < currDot) { e1first = prevIdx; e1second = ind1a; }
        else { e1first = ind1a; e1second = mod_next(ind1a, count); }
    } else if ((ind1b > ind1a) && ((ind1a + 1) == ind1b)) { e1first = ind1a; e1second = ind1b; }
    else { e1first = ind1b; e1second = ind1a; }
    const float2 s1rv0 = shape1.vertex(e1first);
    const float2 s1av0 = transform(s1rv0, xf1);
    const float2 s1rv1 = shape1.vertex(e1second);
    const float2 s1av1 = transform(s1rv1, xf1);
    const unsigned faceType = flipped ? MF_FACE_B : MF_FACE_A;
    {
        const Rot absDir = rot_rotate(s0dir, xf0.q);
        // GetClipPoints, Manifold.cpp:143-157
        ClipV ie[2], mid[2], cps[2];
        ie[0].v = s1av0; ie[0].cf = cf_pack(CF_FACE, idx0, CF_VERTEX, e1first);
        ie[1].v = s1av1; ie[1].cf = cf_pack(CF_FACE, idx0, CF_VERTEX, e1second);
        const float dp0 = -dot(absDir, s0av0);
        const float dp1 = +dot(absDir, s0av1);
        const unsigned nmid = clip_segment(ie, 2, rneg(absDir), dp0, idx0, mid);
        const unsigned ncp = clip_segment(mid, nmid, absDir, dp1, idx0Next, cps);
        if (ncp == 2) {
            const Rot absNormal = fwd_perp(absDir);
            const float2 relMid = vdiv(s0rv0 + s0rv1, 2.0f);
            const float absOffset = dot(absNormal, s0av0);
            const Rot normal = fwd_perp(s0dir);
            Mf m = mf_face(faceType, normal, relMid);
#pragma unroll
            for (unsigned k = 0; k < 2; ++k) {
                if ((dot(absNormal, cps[k].v) - absOffset) <= totalRadius) {
                    mf_add(m, inv_transform(cps[k].v, xf1), flipped ? cf_flip(cps[k].cf) : cps[k].cf);
                }
            }
            if (m.count > 0) return m;
        }
    }
    const float totalRadiusSq = square(totalRadius);
    const bool mustFace = lenEdge0 > square(maxCirclesRatio * r0);
    const Rot fn = fwd_perp(s0dir);
#pragma unroll
    for (unsigned t = 0; t < 4; ++t) {
        // fixed order (v0,v0') (v1,v1') (v0,v1') (v1,v0'), Manifold.cpp:236-289
        const bool use1 = (t == 1 || t == 3);
        const bool use1p = (t == 1 || t == 2);
        const float2 a0 = use1 ? s0av1 : s0av0;
        const float2 a1 = use1p ? s1av1 : s1av0;
        if (magsq(a0 - a1) <= totalRadiusSq) {
            const float2 rr0 = use1 ? s0rv1 : s0rv0;
            const unsigned i0 = use1 ? idx0Next : idx0;
            const float2 rr1 = use1p ? s1rv1 : s1rv0;
            const unsigned i1 = use1p ? e1second : e1first;
            if (mustFace) {
                return !flipped ? mf_face1(MF_FACE_A, fn, rr0, rr1, cf_pack(CF_FACE, i0, CF_VERTEX, i1))
                                : mf_face1(MF_FACE_B, fn, rr0, rr1, cf_pack(CF_VERTEX, i1, CF_FACE, i0));
            }
            return !flipped ? mf_circles(rr0, i0, rr1, i1) : mf_circles(rr1, i1, rr0, i0);
        }
    }
    return mf_unset();
}

// GetManifold(flipped, totalRadius, shape, sxf, point, xfm), Manifold.cpp:293-372
PR_D Mf manifold_poly_point(bool flipped, float totalRadius, const ProxyView& shape, Xf sxf, float2 point, Xf xfm)
{
    const unsigned vertexCount = shape.count;
    const float2 cLocal = inv_transform(transform(point, xfm), sxf);
    unsigned indexOfMax = 0;
    float maxSeparation = -PR_FLT_MAX;
    for (unsigned i = 0; i < vertexCount; ++i) {
        const float s = dot(shape.normal(i), cLocal - shape.vertex(i));
        if (s > totalRadius) return mf_unset();
        if (maxSeparation < s) { maxSeparation = s; indexOfMax = i; }
    }
    const unsigned indexOfMax2 = mod_next(indexOfMax, vertexCount);
    const float2 v1 = shape.vertex(indexOfMax), vv2 = shape.vertex(indexOfMax2);
    const unsigned faceType = flipped ? MF_FACE_B : MF_FACE_A;
    const unsigned faceCf = !flipped ? cf_pack(CF_FACE, indexOfMax, CF_VERTEX, 0) : cf_pack(CF_VERTEX, 0, CF_FACE, indexOfMax);
    if (maxSeparation < 0.0f) {
        return mf_face1(faceType, shape.normal(indexOfMax), vdiv(v1 + vv2, 2.0f), point, faceCf);
    }
    const float2 cLocalV1 = cLocal - v1;
    if (dot(cLocalV1, vv2 - v1) <= 0.0f) {
        if (magsq(cLocalV1) > square(totalRadius)) return mf_unset();
        return !flipped ? mf_circles(v1, indexOfMax, point, 0) : mf_circles(point, 0, v1, indexOfMax);
    }
    const float2 cLocalV2 = cLocal - vv2;
    if (dot(cLocalV2, v1 - vv2) <= 0.0f) {
        if (magsq(cLocalV2) > square(totalRadius)) return mf_unset();
        return !flipped ? mf_circles(vv2, indexOfMax2, point, 0) : mf_circles(point, 0, vv2, indexOfMax2);
    }
    const float2 faceCenter = vdiv(v1 + vv2, 2.0f);
    if (dot(cLocal - faceCenter, shape.normal(indexOfMax)) > totalRadius) return mf_unset();
    return mf_face1(faceType, shape.normal(indexOfMax), faceCenter, point, faceCf);
}

// CollideShapes, Manifold.cpp:393-438
PR_D Mf collide_shapes(const ProxyView& A, Xf xfA, const ProxyView& B, Xf xfB, float linearSlop, float maxCirclesRatio)
{
    const float totalRadius = A.radius + B.radius;
    const unsigned countA = A.count, countB = B.count;
    if (countA == 1 && countB == 1) {
        const float2 pA = transform(A.vertex(0), xfA), pB = transform(B.vertex(0), xfB);
        const float lenSq = magsq(pB - pA);
        const float totSq = square(totalRadius);
        if (lenSq > totSq) return mf_unset();
        return mf_circles(A.vertex(0), 0, B.vertex(0), 0);
    }
    if (countA == 1) return manifold_poly_point(true, totalRadius, B, xfB, A.vertex(0), xfA);
    if (countB == 1) return manifold_poly_point(false, totalRadius, A, xfA, B.vertex(0), xfB);
    const bool do4x4 = (countA == 4) && (countB == 4);
    const SepInfo sepA = do4x4 ? max_separation_4x4(A, xfA, B, xfB, totalRadius) : max_separation(A, xfA, B, xfB, totalRadius);
    if (sepA.distance > totalRadius) return mf_unset();
    const SepInfo sepB = do4x4 ? max_separation_4x4(B, xfB, A, xfA, totalRadius) : max_separation(B, xfB, A, xfA, totalRadius);
    if (sepB.distance > totalRadius) return mf_unset();
    const float k_tol = linearSlop / 10;
    return (sepB.distance > (sepA.distance + k_tol))
        ? manifold_poly_poly(true, B, xfB, sepB.first, A, xfA, sepB.second0, sepB.second1, maxCirclesRatio)
        : manifold_poly_poly(false, A, xfA, sepA.first, B, xfB, sepA.second0, sepA.second1, maxCirclesRatio);
}

} // namespace pr
#endif
