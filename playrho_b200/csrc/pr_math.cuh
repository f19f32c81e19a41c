// pr_math.cuh -- device-side scalar float math for the PlayRho step, written so that every result
// is bit-identical to the reference's x86-64 build: compiled with --fmad=false,
// IEEE div/sqrt (nvcc defaults), the same association order as the reference's value types:
//   Dot: 0 + a0*b0 + a1*b1                 Library/include/playrho/Math.hpp:320-333
//   Vector / scalar = Vector * (1/scalar)  Library/include/playrho/Vector.hpp:367-373,511-521
//   Rotate / InverseRotate / MulT          Library/include/playrho/d2/Math.hpp:113-133,225-229
#ifndef PR_MATH_CUH
#define PR_MATH_CUH
#include "pr_compat.h"

namespace pr {

struct Rot { float c, s; };              // UnitVec (x = cos, y = sin)
struct Xf { float2 p; Rot q; };          // Transformation
struct Pos { float2 l; float a; };       // Position
struct Vel { float2 l; float a; };       // Velocity

#define PR_INF (u2f(0x7f800000u))
#define PR_FLT_MIN 1.17549435e-38f
#define PR_FLT_MAX 3.402823466e+38f

PR_HD float2 v2(float x, float y) { return make_float2(x, y); }
PR_HD float2 operator+(float2 a, float2 b) { return v2(a.x + b.x, a.y + b.y); }
PR_HD float2 operator-(float2 a, float2 b) { return v2(a.x - b.x, a.y - b.y); }
PR_HD float2 operator-(float2 a) { return v2(-a.x, -a.y); }
PR_HD float2 operator*(float2 a, float s) { return v2(a.x * s, a.y * s); }
PR_HD float2 operator*(float s, float2 a) { return v2(s * a.x, s * a.y); }
PR_HD float2 vdiv(float2 a, float s) { const float inv = 1.0f / s; return v2(a.x * inv, a.y * inv); }
PR_HD bool veq(float2 a, float2 b) { return a.x == b.x && a.y == b.y; }
PR_HD float2 rotv(Rot r) { return v2(r.c, r.s); }
PR_HD Rot rneg(Rot r) { Rot o; o.c = -r.c; o.s = -r.s; return o; }
PR_HD Rot mkrot(float c, float s) { Rot o; o.c = c; o.s = s; return o; }
PR_HD float2 operator*(float s, Rot r) { return v2(s * r.c, s * r.s); }
PR_HD float2 operator*(Rot r, float s) { return v2(r.c * s, r.s * s); }

PR_HD float dot(float2 a, float2 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; return r; }
PR_HD float dot(Rot a, float2 b) { return dot(rotv(a), b); }
PR_HD float dot(float2 a, Rot b) { return dot(a, rotv(b)); }
PR_HD float dot(Rot a, Rot b) { return dot(rotv(a), rotv(b)); }
PR_HD float cross(float2 a, float2 b) { const float m = a.x * b.y; const float s = a.y * b.x; return m - s; }
PR_HD float cross(float2 a, Rot b) { return cross(a, rotv(b)); }
PR_HD float magsq(float2 a) { float r = 0.0f; r += a.x * a.x; r += a.y * a.y; return r; }
PR_HD float square(float v) { return v * v; }
// std::min / std::max / std::clamp semantics (not fminf/fmaxf: NaN and signed-zero behaviour differ)
PR_HD float stdmin(float a, float b) { return (b < a) ? b : a; }
PR_HD float stdmax(float a, float b) { return (a < b) ? b : a; }
PR_HD float stdclamp(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

PR_HD float2 rotate(float2 v, Rot q) { return v2((q.c * v.x) - (q.s * v.y), (q.s * v.x) + (q.c * v.y)); }
PR_HD float2 inv_rotate(float2 v, Rot q) { return v2((q.c * v.x) + (q.s * v.y), (q.c * v.y) - (q.s * v.x)); }
PR_HD Rot rot_rotate(Rot v, Rot amount) { return mkrot(v.c * amount.c - v.s * amount.s, v.s * amount.c + v.c * amount.s); }
PR_HD Rot flip_y(Rot r) { return mkrot(r.c, -r.s); }
PR_HD Rot rot_inv_rotate(Rot v, Rot angle) { return rot_rotate(v, flip_y(angle)); }
PR_HD Rot rev_perp(Rot r) { return mkrot(-r.s, r.c); }
PR_HD Rot fwd_perp(Rot r) { return mkrot(r.s, -r.c); }
PR_HD float2 rev_perp(float2 v) { return v2(-v.y, v.x); }
PR_HD Xf mkxf(float2 p, Rot q) { Xf x; x.p = p; x.q = q; return x; }
PR_HD float2 transform(float2 v, Xf xf) { return rotate(v, xf.q) + xf.p; }
PR_HD float2 inv_transform(float2 v, Xf xf) { return inv_rotate(v - xf.p, xf.q); }
PR_HD Xf mulT(Xf A, Xf B) { const float2 dp = B.p - A.p; return mkxf(inv_rotate(dp, A.q), rot_rotate(B.q, flip_y(A.q))); }
PR_HD Xf xf_from4(float4 v) { return mkxf(v2(v.x, v.y), mkrot(v.z, v.w)); }

PR_HD bool is_normal(float v)
{
    const unsigned e = (f2u(v) >> 23) & 0xFFu;
    return e != 0u && e != 0xFFu;
}
PR_HD bool almost_zero(float v) { return fabsf(v) < PR_FLT_MIN; }
// AlmostEqual(a, b, ulp=4), Library/include/playrho/Math.hpp:187-194
PR_HD bool almost_equal(float a, float b)
{
    for (int ulp = 4; (a != b) && (ulp > 0); --ulp) a = nextafterf(a, b);
    return a == b;
}

// ---- sincosf, bit-for-bit glibc 2.39 (sysdeps/ieee754/flt-32/s_sincosf.c, s_sincosf.h,
// s_sincosf_data.c; ARM optimized-routines algorithm by Szabolcs Nagy). Third-party code not under
// /root/reference: restated from its published algorithm; the same restatement is
// checked exhaustively (all 2^32 inputs) against the image's libm in tests/ (oracle/restate).
// glibc ships sincosf (and sinf / cosf, same polynomial) in ifunc variants: the generic one evaluates every a*b+c of the
// polynomial and of the range reduction with two roundings; the x86-64 "-fma" variant, which the dynamic loader selects on CPUs
// with usable FMA + AVX2, contracts them into fused multiply-adds. The two differ, by one ulp of the float result, for 34 of the
// 2^32 inputs (tools/sincos_exhaustive.cpp checks all of them against the libm of the machine it runs on; results in
// profiles/r01h_sincosf_exhaustive.txt). So the reference's own bits depend on the CPU it runs on. This restatement is the
// GENERIC form by default (0 mismatches against a libm that selected the generic variant; --fmad=false keeps nvcc from
// contracting it); -DPR_SINCOSF_FMA builds the contracted form with explicit fma() (0 mismatches against a libm that selected
// the fma variant). Run-time selection by the host's CPU features, like the loader's, is a next step.
#ifdef PR_SINCOSF_FMA
#define PR_MAD(a, b, c) fma((a), (b), (c))
#else
#define PR_MAD(a, b, c) ((c) + (a) * (b))
#endif
struct SinCosTab { double c0, c1, c2, c3, c4, s1, s2, s3; };
PR_HD void sincosf_poly(double x, double x2, bool alt, int n, float* sinp, float* cosp)
{
    const double sg = alt ? -1.0 : 1.0;
    const double C0 = sg * 0x1p0, C1 = sg * -0x1.ffffffd0c621cp-2, C2 = sg * 0x1.55553e1068f19p-5,
                 C3 = sg * -0x1.6c087e89a359dp-10, C4 = sg * 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const double x4 = x2 * x2;
    const double x3 = x2 * x;
    const double c2 = PR_MAD(x2, C4, C3);
    const double s1 = PR_MAD(x2, S3, S2);
    const double c1 = PR_MAD(x2, C1, C0);
    const double x5 = x3 * x2;
    const double x6 = x4 * x2;
    const double s = PR_MAD(x3, S1, x);
    const double c = PR_MAD(x4, C2, c1);
    const float rs = static_cast<float>(PR_MAD(x5, s1, s));
    const float rc = static_cast<float>(PR_MAD(x6, c2, c));
    if (n & 1) { *cosp = rs; *sinp = rc; } else { *sinp = rs; *cosp = rc; }
}
PR_HD unsigned abstop12(float x) { return (f2u(x) >> 20) & 0x7ffu; }
PR_HD void glibc_sincosf(float y, float* sinp, float* cosp)
{
    double x = static_cast<double>(y);
    if (abstop12(y) < 0x3f4u) { // |y| < pi/4   (abstop12(0x1.921FB6p-1f) == 0x3f4)
        const double x2 = x * x;
        if (abstop12(y) < 0x398u) { // |y| < 2^-12
            *sinp = y; *cosp = 1.0f; return;
        }
        sincosf_poly(x, x2, false, 0, sinp, cosp);
    } else if (abstop12(y) < 0x42fu) { // |y| < 120
        const double r = x * 0x1.45F306DC9C883p+23;
        const int n = (static_cast<int>(r) + 0x800000) >> 24;
        x = PR_MAD(-static_cast<double>(n), 0x1.921FB54442D18p0, x); // x - n * hpi
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0; // sign[] = {1,-1,-1,1}
        sincosf_poly(x * s, x * x, (n & 2) != 0, n, sinp, cosp);
    } else if (abstop12(y) < 0x7f8u) { // finite
        const unsigned inv_pio4[24] = {0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44,
            0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5,
            0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599,
            0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
        unsigned xi = f2u(y);
        const int sign = static_cast<int>(xi >> 31);
        const unsigned* arr = &inv_pio4[(xi >> 26) & 15];
        const int shift = (xi >> 23) & 7;
        xi = (xi & 0xffffffu) | 0x800000u;
        xi <<= shift;
        unsigned long long res0 = static_cast<unsigned long long>(xi * arr[0]); // 32-bit product, as in glibc
        const unsigned long long res1 = static_cast<unsigned long long>(xi) * arr[4];
        const unsigned long long res2 = static_cast<unsigned long long>(xi) * arr[8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        const unsigned long long nn = (res0 + (1ULL << 61)) >> 62;
        res0 -= nn << 62;
        const int n = static_cast<int>(nn);
        x = static_cast<double>(static_cast<long long>(res0)) * 0x1.921FB54442D18p-62;
        const int q = (n + sign) & 3;
        const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
        sincosf_poly(x * s, x * x, ((n + sign) & 2) != 0, n, sinp, cosp);
    } else {
        *sinp = *cosp = y - y;
    }
}
// UnitVec::Get(angle), Library/source/playrho/d2/UnitVec.cpp:27-30
PR_HD Rot rot_from_angle(float a) { Rot r; glibc_sincosf(a, &r.s, &r.c); return r; }
// GetTransformation(pos, localCenter), d2/Math.hpp:232-245
// the same function as ONE out-of-line copy: for code that is a single lane's serial chain (the TOI event solve), where
// twenty inlined copies of the double-precision polynomial only inflate the kernel
#ifndef PR_HOST_EMUL
__device__ __noinline__ Rot rot_from_angle_call(float a) { return rot_from_angle(a); }
#else
inline Rot rot_from_angle_call(float a) { return rot_from_angle(a); }
#endif
PR_HD Xf get_xf(float2 ctr, Rot rot, float2 lc) { return mkxf(ctr - rotate(lc, rot), rot); }
PR_HD Xf get_xf(Pos pos, float2 lc) { return get_xf(pos.l, rot_from_angle(pos.a), lc); }

// UnitVec::Get(x, y, fallback), Library/include/playrho/d2/UnitVec.hpp:169-232
PR_HD Rot unit_vector(float2 v, Rot fallback)
{
    const float x = v.x, y = v.y;
    const float ms = x * x + y * y;
    if (is_normal(ms)) {
        const float mag = sqrtf(ms);
        const float inv = 1.0f / mag;
        return mkrot(x * inv, y * inv);
    }
    const float mag = hypotf(x, y);
    if (!is_normal(mag)) return fallback;
    return mkrot(x / mag, y / mag);
}

PR_HD unsigned mod_next(unsigned i, unsigned count) { return (i + 1u < count) ? i + 1u : 0u; }
PR_HD unsigned mod_prev(unsigned i, unsigned count) { return i ? i - 1u : count - 1u; }

PR_HD unsigned cf_pack(unsigned tA, unsigned iA, unsigned tB, unsigned iB) { return tA | (iA << 8) | (tB << 16) | (iB << 24); }
PR_HD unsigned cf_flip(unsigned cf) { return ((cf >> 16) & 0xFFFFu) | ((cf & 0xFFFFu) << 16); }
#define PR_CF_DEFAULT 0xFF00FF00u

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats of either sign
PR_HD unsigned float_flip(float f)
{
    const unsigned u = f2u(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
PR_HD float float_unflip(unsigned u)
{
    return u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

} // namespace pr
#endif
