// orc_math.h -- TEST INFRASTRUCTURE ONLY (oracle restatement). Plain scalar float math restating
// the reference's value-type arithmetic in the SAME operation order: no FMA
// (compile with -ffp-contract=off), vector/scalar division is multiply-by-reciprocal
// (Library/include/playrho/Vector.hpp:367-373,511-521), Dot accumulates from 0 left-to-right
// (Library/include/playrho/Math.hpp:320-333).
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <limits>
#include "../../include/playrho_b200.h"

namespace orc {

struct V2 { float x, y; };
struct Rot { float c, s; };          // UnitVec: x = cos, y = sin
struct Xf { V2 p; Rot q; };          // Transformation
struct Pos { V2 l; float a; };       // Position
struct Vel { V2 l; float a; };       // Velocity

inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator-(V2 a) { return {-a.x, -a.y}; }
inline V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }
inline V2 operator*(float s, V2 a) { return {s * a.x, s * a.y}; }
// Vector / scalar: Vector.hpp:511-521 (reciprocal then multiply)
inline V2 operator/(V2 a, float s) { const float inv = 1.0f / s; return {a.x * inv, a.y * inv}; }
inline bool operator==(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(V2 a, V2 b) { return !(a == b); }
inline V2 rotv(Rot r) { return {r.c, r.s}; }
inline Rot operator-(Rot r) { return {-r.c, -r.s}; }
inline V2 operator*(float s, Rot r) { return {s * r.c, s * r.s}; }
inline V2 operator*(Rot r, float s) { return {r.c * s, r.s * s}; }

// Math.hpp:320-333
inline float dot(V2 a, V2 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; return r; }
inline float dot(Rot a, V2 b) { return dot(rotv(a), b); }
inline float dot(V2 a, Rot b) { return dot(a, rotv(b)); }
inline float dot(Rot a, Rot b) { return dot(rotv(a), rotv(b)); }
// Math.hpp:365-386
inline float cross(V2 a, V2 b) { const float m = a.x * b.y; const float s = a.y * b.x; return m - s; }
inline float cross(V2 a, Rot b) { return cross(a, rotv(b)); }
// Math.hpp:274-283
inline float magsq(V2 a) { float r = 0.0f; r += a.x * a.x; r += a.y * a.y; return r; }
inline float square(float v) { return v * v; }

// d2/Math.hpp:113-133
inline V2 rotate(V2 v, Rot q) { return {(q.c * v.x) - (q.s * v.y), (q.s * v.x) + (q.c * v.y)}; }
inline V2 inv_rotate(V2 v, Rot q) { return {(q.c * v.x) + (q.s * v.y), (q.c * v.y) - (q.s * v.x)}; }
// UnitVec::Rotate, UnitVec.hpp:353-358
inline Rot rot_rotate(Rot v, Rot amount)
{
    return {v.c * amount.c - v.s * amount.s, v.s * amount.c + v.c * amount.s};
}
inline Rot flip_y(Rot r) { return {r.c, -r.s}; }
inline Rot rot_inv_rotate(Rot v, Rot angle) { return rot_rotate(v, flip_y(angle)); } // UnitVec.hpp:461
inline Rot rev_perp(Rot r) { return {-r.s, r.c}; }  // UnitVec.hpp:364
inline Rot fwd_perp(Rot r) { return {r.s, -r.c}; }  // UnitVec.hpp:374
inline V2 rev_perp(V2 v) { return {-v.y, v.x}; }    // Vector2.hpp:98
// d2/Math.hpp:194-211
inline V2 transform(V2 v, Xf xf) { return rotate(v, xf.q) + xf.p; }
inline V2 inv_transform(V2 v, Xf xf) { return inv_rotate(v - xf.p, xf.q); }
// d2/Math.hpp:225-229
inline Xf mulT(Xf A, Xf B) { const V2 dp = B.p - A.p; return {inv_rotate(dp, A.q), rot_rotate(B.q, flip_y(A.q))}; }

// UnitVec::Get(angle), UnitVec.cpp:27-30 -- libm cos/sin (glibc; SURVEY §8c third-party)
inline Rot rot_from_angle(float a) { return {std::cos(a), std::sin(a)}; }
// GetTransformation(pos, localCenter), d2/Math.hpp:232-245
inline Xf get_xf(V2 ctr, Rot rot, V2 lc) { return {ctr - rotate(lc, rot), rot}; }
inline Xf get_xf(Pos pos, V2 lc) { return get_xf(pos.l, rot_from_angle(pos.a), lc); }

// UnitVec::Get(x, y, fallback), UnitVec.hpp:169-232. ok=false => fallback used.
inline Rot unit_vector(V2 v, Rot fallback)
{
    const float x = v.x, y = v.y;
    const float ms = x * x + y * y;
    if (std::isnormal(ms)) {
        const float mag = std::sqrt(ms);
        const float inv = 1.0f / mag;
        return {x * inv, y * inv};
    }
    const float mag = std::hypot(x, y);
    if (!std::isnormal(mag)) return fallback;
    return {x / mag, y / mag};
}

// Math.hpp:177-194
inline bool almost_zero(float v) { return std::fabs(v) < FLT_MIN; }
inline bool almost_equal(float a, float b, int ulp = 4)
{
    for (; (a != b) && (ulp > 0); --ulp) a = std::nextafter(a, b);
    return a == b;
}

// GetPosition(pos0,pos1,beta), d2/Math.cpp:30-60: pos0 + (pos1 - pos0) * beta
inline Pos lerp_pos(Pos p0, Pos p1, float beta)
{
    return {p0.l + (p1.l - p0.l) * beta, p0.a + (p1.a - p0.a) * beta};
}

// child accessors over the flat pools
struct Proxy {
    float radius;
    uint32_t count;
    const pr_vec2* v;
    const pr_vec2* n;
    V2 vertex(uint32_t i) const { return {v[i].x, v[i].y}; }
    Rot normal(uint32_t i) const { return {n[i].x, n[i].y}; }
};
inline uint32_t mod_next(uint32_t i, uint32_t count) { return (i + 1u < count) ? i + 1u : 0u; } // Math.hpp GetModuloNext
inline uint32_t mod_prev(uint32_t i, uint32_t count) { return i ? i - 1u : count - 1u; }

inline pr_contact_feature cf_pack(uint32_t tA, uint32_t iA, uint32_t tB, uint32_t iB)
{
    return tA | (iA << 8) | (tB << 16) | (iB << 24);
}
inline pr_contact_feature cf_flip(pr_contact_feature cf)
{
    return ((cf >> 16) & 0xFFFFu) | ((cf & 0xFFFFu) << 16);
}
static const pr_contact_feature CF_DEFAULT = (0u) | (255u << 8) | (0u << 16) | (255u << 24);

pr_manifold collide_shapes(const Proxy& A, Xf xfA, const Proxy& B, Xf xfB, float linearSlop,
                           float maxCirclesRatio);

} // namespace orc
#endif
