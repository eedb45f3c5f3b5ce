// rt_math.cuh -- scalar float math with the reference's exact operation order.
//
// Parity with Ref (reference internal/CoreRef.cpp, ShadeRef.cpp) needs every rounding to happen where the reference's
// SSE2 build puts it: the TU is compiled with -msse2 -mno-avx (reference CMakeLists.txt:41), i.e. no fma contraction,
// IEEE div/sqrt; this TU is compiled with -fmad=false and default (IEEE) division and square root.  Vector helpers
// reproduce the association order of the reference's 4-wide `fvec4` reductions (internal/simd/simd_sse.h):
//   dot/length : (x*x' + y*y') + (w*w' + z*z')            simd_sse.h:120-142, 252-260
//   hsum       : ((x + y) + z) + w                        simd_sse.h:144-152 (no SSE4.1 in the Ref TU)
//   min/max    : _mm_min_ps / _mm_max_ps operand order    simd_sse.h:180-188
// v3 is an fvec4 whose 4th lane is known to be 0, so w*w' + z*z' == z*z'.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_types.h"

#define RT_DEV __device__ __forceinline__
// RT_FN marks the heavy leaf functions (BSDF nodes, light sampling, normalize, ...).  History, all on hall-250k:
//   everything inlined, 128-thread blocks:   k_shade = 460 KB of SASS, I-cache hit 65 %, no_instruction 4-5 stalls/issue
//   RT_FN = noinline (real calls):           177 KB, -25 % time, but every context struct lives in local memory
//                                            (2 GB of DRAM writes per launch)
//   inlined again + block-synchronised warps (RT_SHADE_SYNC, rt_kernels.cuh): the warps of a 512-thread block walk
//   the straight-line code together and share instruction-cache lines -> 118 -> 54 ms per 16 samples.
// Inlining never changes results here (-fmad=false, no contraction across calls).  -DRT_FN="__device__ __noinline__"
// restores the call-based build for A/B measurements.
#ifndef RT_FN
#define RT_FN __device__ __forceinline__
#endif

namespace rt {

struct v2 {
    float x, y;
};
struct v3 {
    float x, y, z;
};

struct c4 { // rgb + a 4th lane (pdf for BSDF results, alpha for pixel colours)
    float x, y, z, w;
};
struct v4 {
    float x, y, z, w;
};

RT_DEV v3 mk3(float x, float y, float z) { return v3{x, y, z}; }
RT_DEV v3 mk3(const float *p) { return v3{p[0], p[1], p[2]}; }
RT_DEV v2 mk2(float x, float y) { return v2{x, y}; }

RT_DEV v3 operator+(v3 a, v3 b) { return v3{a.x + b.x, a.y + b.y, a.z + b.z}; }
RT_DEV v3 operator-(v3 a, v3 b) { return v3{a.x - b.x, a.y - b.y, a.z - b.z}; }
RT_DEV v3 operator*(v3 a, v3 b) { return v3{a.x * b.x, a.y * b.y, a.z * b.z}; }
RT_DEV v3 operator/(v3 a, v3 b) { return v3{a.x / b.x, a.y / b.y, a.z / b.z}; }
RT_DEV v3 operator*(v3 a, float s) { return v3{a.x * s, a.y * s, a.z * s}; }
RT_DEV v3 operator*(float s, v3 a) { return v3{s * a.x, s * a.y, s * a.z}; }
RT_DEV v3 operator/(v3 a, float s) { return v3{a.x / s, a.y / s, a.z / s}; }
RT_DEV v3 operator-(v3 a) { return v3{-a.x, -a.y, -a.z}; }
RT_DEV v3 &operator+=(v3 &a, v3 b) {
    a = a + b;
    return a;
}
RT_DEV v3 &operator*=(v3 &a, v3 b) {
    a = a * b;
    return a;
}
RT_DEV v3 &operator*=(v3 &a, float s) {
    a = a * s;
    return a;
}
RT_DEV v3 &operator/=(v3 &a, float s) {
    a = a / s;
    return a;
}

RT_DEV v2 operator+(v2 a, v2 b) { return v2{a.x + b.x, a.y + b.y}; }
RT_DEV v2 operator-(v2 a, v2 b) { return v2{a.x - b.x, a.y - b.y}; }
RT_DEV v2 operator*(v2 a, v2 b) { return v2{a.x * b.x, a.y * b.y}; }
RT_DEV v2 operator*(v2 a, float s) { return v2{a.x * s, a.y * s}; }
RT_DEV v2 operator*(float s, v2 a) { return v2{s * a.x, s * a.y}; }

// simd_sse.h:252-260 with the w lane zero.
RT_DEV float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z); }
RT_DEV float length(v3 a) { return sqrtf(dot(a, a)); }
RT_DEV float length2(v3 a) { return dot(a, a); }
RT_FN v3 normalize(v3 a) { return a / length(a); }
RT_FN v3 normalize_len(v3 a, float &len) {
    len = length(a);
    return a / len;
}
// generic fvec<2>: accumulates left to right from 0 (simd.h:304-314,475-479)
RT_DEV float dot(v2 a, v2 b) { return (0.0f + a.x * b.x) + a.y * b.y; }
RT_DEV float length(v2 a) { return sqrtf(dot(a, a)); }

// CoreRef.h:281-285
RT_DEV v3 cross(v3 a, v3 b) { return v3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// _mm_min_ps(a,b) = a < b ? a : b ; _mm_max_ps(a,b) = a > b ? a : b  (second operand on NaN / equal)
RT_DEV float sse_min(float a, float b) { return a < b ? a : b; }
RT_DEV float sse_max(float a, float b) { return a > b ? a : b; }
// std::min / std::max as used by the generic fvec<2> (simd.h:388-398)
RT_DEV float std_min(float a, float b) { return (b < a) ? b : a; }
RT_DEV float std_max(float a, float b) { return (a < b) ? b : a; }

RT_DEV float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); } // Core.h:537-539
RT_DEV float saturatef(float v) { return clampf(v, 0.0f, 1.0f); }
RT_DEV float sqr(float x) { return x * x; }
RT_DEV float mixf(float a, float b, float k) { return (1.0f - k) * a + k * b; } // ShadeRef.cpp:30
RT_DEV v3 mix3(v3 a, v3 b, float k) { return (1.0f - k) * a + k * b; }          // simd.h:609-611
RT_DEV float fractf(float v) { return v - floorf(v); }                          // CoreRef.h:158

RT_DEV float safe_sqrt(float v) { return sqrtf(fmaxf(v, 0.0f)); }
RT_DEV float safe_div(float a, float b) { return b != 0.0f ? (a / b) : kFltMax; }
RT_DEV float safe_div_pos(float a, float b) { return a / fmaxf(b, kFltEps); }
RT_DEV float safe_div_neg(float a, float b) { return a / fminf(b, -kFltEps); }
RT_FN v3 safe_normalize(v3 a) {
    const float l = length(a);
    return l > 0.0f ? (a / l) : a;
}
// CoreRef.h:193-197
RT_FN v3 safe_invert(v3 d) {
    v3 r;
    r.x = 1.0f / ((fabsf(d.x) > kFltEps) ? d.x : copysignf(kFltEps, d.x));
    r.y = 1.0f / ((fabsf(d.y) > kFltEps) ? d.y : copysignf(kFltEps, d.y));
    r.z = 1.0f / ((fabsf(d.z) > kFltEps) ? d.z : copysignf(kFltEps, d.z));
    return r;
}

RT_DEV float lum(v3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; } // CoreRef.h:398-404
RT_DEV float power_heuristic(float a, float b) {                                         // CoreRef.h:424-427
    const float t = a * a;
    return t / (b * b + t);
}

RT_DEV float fast_log2(float val) { // CoreRef.h:406-417
    int x = __float_as_int(val);
    float log_2 = float(((x >> 23) & 255) - 128);
    x &= ~(255 << 23);
    x += 127 << 23;
    const float m = __int_as_float(x);
    log_2 += ((-0.34484843f) * m + 2.02466578f) * m - 0.67487759f;
    return log_2;
}

// "A Fast and Robust Method for Avoiding Self-Intersection" as in CoreRef.h:447-462 (ivec4(float) truncates).
RT_FN v3 offset_ray(v3 p, v3 n) {
    const float Origin = 1.0f / 32.0f;
    const float FloatScale = 1.0f / 65536.0f;
    const float IntScale = 128.0f;
    const int ox = __float2int_rz(IntScale * n.x), oy = __float2int_rz(IntScale * n.y),
              oz = __float2int_rz(IntScale * n.z);
    const float ix = __int_as_float(__float_as_int(p.x) + ((p.x < 0.0f) ? -ox : ox));
    const float iy = __int_as_float(__float_as_int(p.y) + ((p.y < 0.0f) ? -oy : oy));
    const float iz = __int_as_float(__float_as_int(p.z) + ((p.z < 0.0f) ? -oz : oz));
    return v3{fabsf(p.x) < Origin ? (p.x + FloatScale * n.x) : ix, fabsf(p.y) < Origin ? (p.y + FloatScale * n.y) : iy,
              fabsf(p.z) < Origin ? (p.z + FloatScale * n.z) : iz};
}

// ---- trigonometry: the reference's own polynomials (CoreRef.cpp:1131-1270), not CUDA's --------------------------
// portable_cos/sin evaluate three range-shifted copies of one polynomial and select with a 0/1 mask through a dot
// product; exactly one mask lane is +-1 and the others are 0, so evaluating only the selected lane is bit-identical.
RT_DEV float trig_poly(float arg) {
    arg = arg * arg;
    float res = -25.0407296503853054f * arg + 60.1524123580209817f;
    res = res * arg - 85.4539888046442542f;
    res = res * arg + 64.9393549651994562f;
    res = res * arg - 19.7392086060579359f;
    res = res * arg + 0.9999999998415476f;
    return res;
}
RT_DEV float trig_select(float a) { // a in [0,1): fraction of a full turn
    if (a < 0.25f) {
        return trig_poly(a);
    } else if (a >= 0.75f) {
        return trig_poly(a - 1.0f);
    }
    return -trig_poly(a - 0.5f);
}
RT_FN float portable_cos(float a) { return trig_select(fractf(fabsf(a) * 0.15915494309189535f)); }
RT_FN float portable_sin(float a) {
    return trig_select(fractf(fabsf(a - 1.5707963267948966f) * 0.15915494309189535f));
}
// returns {sin, cos} like Ref::portable_sincos
RT_FN v2 portable_sincos(float a) { return v2{portable_sin(a), portable_cos(a)}; }

RT_DEV float asin_tail(float x) {
    return (kPi / 2) - ((x + 2.71745038f) * x + 14.0375338f) * (0.00440413551f * ((x - 8.31223679f) * x + 25.3978882f)) *
                           sqrtf(1 - x);
}
RT_FN float portable_asinf(float x) {
    if (fabsf(x) > 0.57f) {
        const float ret = asin_tail(fabsf(x));
        return (x < 0.0f) ? -ret : ret;
    } else {
        const float x2 = x * x;
        return x + (0.0517513789f * ((x2 + 1.83372748f) * x2 + 1.56678128f)) * x *
                       (x2 * ((x2 - 1.48268414f) * x2 + 2.05554748f));
    }
}
RT_FN float portable_acosf(float x) {
    if (x < -0.62f) {
        return kPi - (((x - 2.71850395f) * x + 14.7303705f)) * (0.00393401226f * ((x + 8.60734272f) * x + 27.0927486f)) *
                         sqrtf(1 + x);
    } else if (x <= 0.62f) {
        const float x2 = x * x;
        return (kPi / 2) - x -
               (0.0700945929f * x * ((x2 + 1.57144082f) * x2 + 1.25210774f)) *
                   (x2 * ((x2 - 1.53757966f) * x2 + 1.89929986f));
    } else {
        return (((x + 2.71850395f) * x + 14.7303705f)) * (0.00393401226f * ((x - 8.60734272f) * x + 27.0927486f)) *
               sqrtf(1 - x);
    }
}

// acosf as glibc 2.39 computes it for binary32 (sysdeps/ieee754/flt-32/e_acosf.c, the fdlibm algorithm in float
// arithmetic).  The reference calls libm acosf in slerp() (CoreRef.cpp:1110-1126) and for spot lights; CUDA's acosf
// is a different <=1-ulp approximation, and one ulp in a sampled direction is enough to flip a later discrete
// decision.  tests/test_libm.py checks this restatement against the host libm over the whole [-1,1] domain.
RT_FN float libm_acosf(float x) {
    const float one = 1.0000000000e+00f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f,
                pio2_lo = 7.5497894159e-08f, pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f,
                pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f,
                qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    const int hx = __float_as_int(x);
    const int ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) {
        if (hx > 0) {
            return 0.0f;
        }
        return pi + 2.0f * pio2_lo;
    } else if (ix > 0x3f800000) {
        return (x - x) / (x - x);
    }
    if (ix < 0x3f000000) { // |x| < 0.5
        if (ix <= 0x23000000) {
            return pio2_hi + pio2_lo;
        }
        const float z = x * x;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (hx < 0) { // x < -0.5
        const float z = (one + x) * 0.5f;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float s = sqrtf(z);
        const float r = p / q;
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    } else { // x > 0.5
        const float z = (one - x) * 0.5f;
        const float s = sqrtf(z);
        const float df = __int_as_float(__float_as_int(s) & 0xfffff000);
        const float c = (z - df * df) / (s + df);
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = p / q;
        const float w = r * s + c;
        return 2.0f * (df + w);
    }
}

// sinf / cosf / atan2f of the host libm (glibc 2.39), restated.  The environment-map code of the reference calls them
// (SampleLatlong_RGBE CoreRef.cpp:2995-3039, CanonicalToDir / DirToCanonical Core.cpp:110-143) and the results index
// texels / quad-tree cells, so one ulp matters.  sinf/cosf: glibc's sysdeps/ieee754/flt-32/s_sincosf.h (the Arm
// optimized-routines algorithm: double-precision polynomial after a fast reduction by pi/2; arguments here are angles in
// [-7, 7], far below the 120.0f where the big-argument reduction starts).  atan2f/atanf: the fdlibm float versions
// (e_atan2f.c, s_atanf.c).  tests/test_libm.py compares all of them bit for bit with the host libm.
RT_FN float libm_sincosf_poly(double x, double x2, bool second_table, int n) {
    // __sincosf_table[0] / [1]: the second table negates the cosine coefficients
    const double sg = second_table ? -1.0 : 1.0;
    const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5,
                 c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double t1 = s2 + x2 * s3;
        const double x7 = x3 * x2;
        const double s = x + x3 * s1;
        return float(s + x7 * t1);
    } else {
        const double x4 = x2 * x2;
        const double t2 = c3 + x2 * c4;
        const double t1 = c0 + x2 * c1;
        const double x6 = x4 * x2;
        const double c = t1 + x4 * c2;
        return float(c + x6 * t2);
    }
}

RT_FN float libm_sinf(float y) {
    double x = double(y);
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ffu;
    if (top < 0x3f4u) { // |y| < pi/4
        if (top < 0x398u) { // |y| < 2^-12
            return y;
        }
        return libm_sincosf_poly(x, x * x, false, 0);
    }
    const double r = x * 0x1.45F306DC9C883p+23;
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = x - double(n) * 0x1.921FB54442D18p0;
    const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return libm_sincosf_poly(x * sign, x * x, (n & 2) != 0, n);
}

RT_FN float libm_cosf(float y) {
    double x = double(y);
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ffu;
    if (top < 0x3f4u) {
        if (top < 0x398u) {
            return 1.0f;
        }
        return libm_sincosf_poly(x, x * x, false, 1);
    }
    const double r = x * 0x1.45F306DC9C883p+23;
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = x - double(n) * 0x1.921FB54442D18p0;
    const int m = n + 1;
    const double sign = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
    return libm_sincosf_poly(x * sign, x * x, (m & 2) != 0, n ^ 1);
}

RT_FN float libm_atanf(float x) {
    const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
    const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const int hx = __float_as_int(x);
    const int ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c800000) { // |x| >= 2^26
        if (ix > 0x7f800000) {
            return x + x;
        }
        return (hx > 0) ? (hi3 + lo3) : (-hi3 - lo3);
    }
    if (ix < 0x3ee00000) { // |x| < 0.4375
        if (ix < 0x31000000) { // |x| < 2^-29
            return x;
        }
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) { // |x| < 1.1875
            if (ix < 0x3f300000) { // 7/16 <= |x| < 11/16
                id = 0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                id = 1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000) { // |x| < 2.4375
                id = 2;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else {
                id = 3;
                x = -1.0f / x;
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) {
        return x - x * (s1 + s2);
    }
    const float hi = (id == 0) ? hi0 : ((id == 1) ? hi1 : ((id == 2) ? hi2 : hi3));
    const float lo = (id == 0) ? lo0 : ((id == 1) ? lo1 : ((id == 2) ? lo2 : lo3));
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return (hx < 0) ? -r : r;
}

RT_FN float libm_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const int hx = __float_as_int(x), hy = __float_as_int(y);
    const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) {
        return x + y;
    }
    if (hx == 0x3f800000) {
        return libm_atanf(y);
    }
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        return (m < 2) ? y : ((m == 2) ? (pi + tiny) : (-pi - tiny));
    }
    if (ix == 0) {
        return (hy < 0) ? (-pi_o_2 - tiny) : (pi_o_2 + tiny);
    }
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            return (m == 0) ? (pi_o_4 + tiny) : ((m == 1) ? (-pi_o_4 - tiny) : ((m == 2) ? (3.0f * pi_o_4 + tiny) : (-3.0f * pi_o_4 - tiny)));
        }
        return (m == 0) ? 0.0f : ((m == 1) ? -0.0f : ((m == 2) ? (pi + tiny) : (-pi - tiny)));
    }
    if (iy == 0x7f800000) {
        return (hy < 0) ? (-pi_o_2 - tiny) : (pi_o_2 + tiny);
    }
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) {
        z = pi_o_2 + 0.5f * pi_lo;
    } else if (hx < 0 && k < -60) {
        z = 0.0f;
    } else {
        z = libm_atanf(fabsf(y / x));
    }
    return (m == 0) ? z : ((m == 1) ? -z : ((m == 2) ? (pi - (z - pi_lo)) : ((z - pi_lo) - pi)));
}

// expf of the host libm (glibc 2.39 sysdeps/ieee754/flt-32/e_expf.c, the Arm optimized-routines algorithm: 2^(k/32) table
// times a cubic in double precision).  The NLM denoiser's weights are expf(-distance) (DenoiseRef.cpp:55,76).  The table is
// T[i] = bits(2^(i/32)) - (i << 47), regenerated from its definition; tests/test_libm.py compares with the host expf.
RT_FN float libm_expf(float x) {
    static const uint64_t T[32] = {0x3ff0000000000000ULL,0x3fefd9b0d3158574ULL,0x3fefb5586cf9890fULL,0x3fef9301d0125b51ULL,0x3fef72b83c7d517bULL,0x3fef54873168b9aaULL,0x3fef387a6e756238ULL,0x3fef1e9df51fdee1ULL,0x3fef06fe0a31b715ULL,0x3feef1a7373aa9cbULL,0x3feedea64c123422ULL,0x3feece086061892dULL,0x3feebfdad5362a27ULL,0x3feeb42b569d4f82ULL,0x3feeab07dd485429ULL,0x3feea47eb03a5585ULL,0x3feea09e667f3bcdULL,0x3fee9f75e8ec5f74ULL,0x3feea11473eb0187ULL,0x3feea589994cce13ULL,0x3feeace5422aa0dbULL,0x3feeb737b0cdc5e5ULL,0x3feec49182a3f090ULL,0x3feed503b23e255dULL,0x3feee89f995ad3adULL,0x3feeff76f2fb5e47ULL,0x3fef199bdd85529cULL,0x3fef3720dcef9069ULL,0x3fef5818dcfba487ULL,0x3fef7c97337b9b5fULL,0x3fefa4afa2a490daULL,0x3fefd0765b6e4540ULL};
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
    const double xd = double(x);
    const uint32_t abstop = (__float_as_uint(x) >> 20) & 0x7ffu;
    if (abstop >= 0x42bu) { // |x| >= 88 or NaN
        if (__float_as_uint(x) == 0xff800000u) {
            return 0.0f;
        }
        if (abstop >= 0x7f8u) {
            return x + x;
        }
        if (x > 0x1.62e42ep6f) {
            return __uint_as_float(0x7f800000u);
        }
        if (x < -0x1.9fe368p6f) {
            return 0.0f;
        }
    }
    double z = InvLn2N * xd;
    double kd = z + SHIFT;
    const uint64_t ki = uint64_t(__double_as_longlong(kd));
    kd -= SHIFT;
    const double r = z - kd;
    uint64_t t = T[ki % 32];
    t += ki << (52 - 5);
    const double s = __longlong_as_double((long long)t);
    z = C0 * r + C1;
    const double r2 = r * r;
    double y = C2 * r + 1;
    y = z * r2 + y;
    y = y * s;
    return float(y);
}

// logf and powf of the host libm (glibc 2.39 sysdeps/ieee754/flt-32/e_logf.c, e_powf.c: the Arm optimized-routines
// algorithms -- a 16-entry {1/c, log(c)} table indexed by the top mantissa bits, a short polynomial in double precision,
// for powf followed by the exp2 kernel of libm_expf).  The reference calls logf in D_GTR1 (clearcoat lobe,
// ShadeRef.cpp) and powf in the display transform (TonemapRef.h:20-47: the sRGB OETF and 1/gamma), so with these the
// tonemapped plane and the clearcoat pdf are the host's floats, not CUDA's <= 2 ulp versions.  Table values are the
// published constants of those files as found in this image's libm.so.6; tests/test_libm.py compares tens of millions of
// arguments with the host functions.
struct LibmLogTab {
    double invc, logc;
};
RT_FN float libm_logf(float x) {
    static const LibmLogTab T[16] = {
        {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2},
        {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
        {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2},
        {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
        {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3},
        {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
        {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4},
        {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
        {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
        {0x1.0000000000000p+0, 0x0.0p+0},
        {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},
        {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
        {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},
        {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
        {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
        {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
    const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) {
        return 0.0f;
    }
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        // x < 0x1p-126 or inf or nan
        if (ix * 2 == 0) {
            return -__uint_as_float(0x7f800000u); // log(+-0) = -inf
        }
        if (ix == 0x7f800000u) {
            return x; // log(inf) = inf
        }
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) {
            return __uint_as_float(0x7fc00000u) ; // log(negative) / log(nan)
        }
        // subnormal: normalise
        ix = __float_as_uint(x * 0x1p23f);
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = int((tmp >> (23 - 4)) % 16u);
    const int k = int(tmp) >> 23; // arithmetic shift
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = T[i].invc, logc = T[i].logc;
    const double z = double(__uint_as_float(iz));
    const double r = z * invc - 1;
    const double y0 = logc + double(k) * Ln2;
    const double r2 = r * r;
    double y = A1 * r + A2;
    y = A0 * r2 + y;
    y = y * r2 + (y0 + r);
    return float(y);
}

// x > 0 finite or x == 0, y finite: the calls the display transform makes (x = a linear pixel value, y = 1/2.4 or
// 1/gamma).  Negative x (a non-integer power is NaN), infinities and NaNs take the device powf: they never reach a
// comparable pixel.
RT_FN float libm_powf(float x, float y) {
    static const LibmLogTab T[16] = {
        {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2},
        {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
        {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2},
        {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
        {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2},
        {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
        {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3},
        {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
        {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5},
        {0x1.0000000000000p+0, 0x0.0p+0},
        {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},
        {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
        {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},
        {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
        {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},
        {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
    const double P0 = 0x1.27616c9496e0bp-2, P1 = -0x1.71969a075c67ap-2, P2 = 0x1.ec70a6ca7baddp-2, P3 = -0x1.7154748bef6c8p-1,
                 P4 = 0x1.71547652ab82bp+0;
    static const uint64_t E[32] = {0x3ff0000000000000ULL,0x3fefd9b0d3158574ULL,0x3fefb5586cf9890fULL,0x3fef9301d0125b51ULL,0x3fef72b83c7d517bULL,0x3fef54873168b9aaULL,0x3fef387a6e756238ULL,0x3fef1e9df51fdee1ULL,0x3fef06fe0a31b715ULL,0x3feef1a7373aa9cbULL,0x3feedea64c123422ULL,0x3feece086061892dULL,0x3feebfdad5362a27ULL,0x3feeb42b569d4f82ULL,0x3feeab07dd485429ULL,0x3feea47eb03a5585ULL,0x3feea09e667f3bcdULL,0x3fee9f75e8ec5f74ULL,0x3feea11473eb0187ULL,0x3feea589994cce13ULL,0x3feeace5422aa0dbULL,0x3feeb737b0cdc5e5ULL,0x3feec49182a3f090ULL,0x3feed503b23e255dULL,0x3feee89f995ad3adULL,0x3feeff76f2fb5e47ULL,0x3fef199bdd85529cULL,0x3fef3720dcef9069ULL,0x3fef5818dcfba487ULL,0x3fef7c97337b9b5fULL,0x3fefa4afa2a490daULL,0x3fefd0765b6e4540ULL};
    const double SHIFT = 0x1.8p+52 / 32, C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
    uint32_t ix = __float_as_uint(x);
    const uint32_t iy = __float_as_uint(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || (iy * 2 - 1 >= 2u * 0x7f800000u - 1)) {
        if (iy * 2 == 0) {
            return 1.0f; // x^0
        }
        if (ix == 0x3f800000u) {
            return 1.0f;
        }
        if (ix * 2 == 0 && iy * 2 < 2u * 0x7f800000u) {
            return (iy & 0x80000000u) ? __uint_as_float(0x7f800000u) : 0.0f; // (+-0)^y, y non-integer sign ignored: |result|
        }
        if ((ix & 0x80000000u) || ix >= 0x7f800000u || iy * 2 >= 2u * 0x7f800000u) {
            return powf(x, y); // negative base / inf / nan: off the compared path
        }
        // subnormal x: normalise
        ix = __float_as_uint(x * 0x1p23f);
        ix &= 0x7fffffffu;
        ix -= 23u << 23;
    }
    // log2_inline
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = int((tmp >> (23 - 4)) % 16u);
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = int(tmp) >> 23;
    const double invc = T[i].invc, logc = T[i].logc;
    const double z = double(__uint_as_float(iz));
    const double r = z * invc - 1;
    const double y0 = logc + double(k);
    const double r2 = r * r;
    double yy = P0 * r + P1;
    const double p = P2 * r + P3;
    const double r4 = r2 * r2;
    double q = P4 * r + y0;
    q = p * r2 + q;
    yy = yy * r4 + q;
    const double ylogx = double(y) * yy;
    if (((uint64_t(__double_as_longlong(ylogx)) >> 47) & 0xffffu) >= (uint64_t(__double_as_longlong(126.0)) >> 47)) {
        // |y * log2(x)| >= 126
        if (ylogx > 0x1.fffffffd1d571p+6) {
            return __uint_as_float(0x7f800000u);
        }
        if (ylogx <= -150.0) {
            return 0.0f;
        }
    }
    // exp2_inline
    double kd = ylogx + SHIFT;
    const uint64_t ki = uint64_t(__double_as_longlong(kd));
    kd -= SHIFT;
    const double rr = ylogx - kd;
    uint64_t t = E[ki % 32];
    t += ki << (52 - 5);
    const double s = __longlong_as_double((long long)t);
    const double zz = C0 * rr + C1;
    const double rr2 = rr * rr;
    double o = C2 * rr + 1;
    o = zz * rr2 + o;
    o = o * s;
    return float(o);
}

// exp2f(float(e) - 128.0f) of rgbe_to_rgb (CoreRef.h:234-237): an exact power of two for every byte e
RT_DEV float rgbe_scale(uint32_t e) {
    const int n = int(e) - 128;
    return (n >= -126) ? __int_as_float((n + 127) << 23) : __int_as_float(1 << (149 + n));
}

// CoreRef.cpp:771-802
RT_DEV float approx_atan2(float y, float x) {
    float t0, t1, t3, t4;
    t3 = fabsf(x);
    t1 = fabsf(y);
    t0 = fmaxf(t3, t1);
    t1 = fminf(t3, t1);
    t3 = 1.0f / t0;
    t3 = t1 * t3;
    t4 = t3 * t3;
    t0 = -0.013480470f;
    t0 = t0 * t4 + 0.057477314f;
    t0 = t0 * t4 - 0.121239071f;
    t0 = t0 * t4 + 0.195635925f;
    t0 = t0 * t4 - 0.332994597f;
    t0 = t0 * t4 + 0.999995630f;
    t3 = t0 * t3;
    t3 = (fabsf(y) > fabsf(x)) ? 1.570796327f - t3 : t3;
    t3 = (x < 0) ? 3.141592654f - t3 : t3;
    t3 = (y < 0) ? -t3 : t3;
    return t3;
}

// ---- matrices (column-major 4x4), CoreRef.cpp:2789-2816 -----------------------------------------------------------
RT_DEV v3 transform_point(v3 p, const float *m) {
    return v3{m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
              m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
}
RT_DEV v3 transform_direction(v3 p, const float *m) {
    return v3{m[0] * p.x + m[4] * p.y + m[8] * p.z, m[1] * p.x + m[5] * p.y + m[9] * p.z,
              m[2] * p.x + m[6] * p.y + m[10] * p.z};
}
RT_DEV v3 transform_normal(v3 n, const float *inv) {
    return v3{inv[0] * n.x + inv[1] * n.y + inv[2] * n.z, inv[4] * n.x + inv[5] * n.y + inv[6] * n.z,
              inv[8] * n.x + inv[9] * n.y + inv[10] * n.z};
}

RT_DEV v3 world_from_tangent(v3 T, v3 B, v3 N, v3 V) { return V.x * T + V.y * B + V.z * N; } // CoreRef.h:296-298
RT_DEV v3 tangent_from_world(v3 T, v3 B, v3 N, v3 V) { return v3{dot(V, T), dot(V, B), dot(V, N)}; }

// CoreRef.cpp:675-688
RT_FN void create_tbn(v3 N, v3 &T, v3 &B) {
    v3 U;
    if (fabsf(N.y) < 0.999f) {
        U = v3{0.0f, 1.0f, 0.0f};
    } else {
        U = v3{1.0f, 0.0f, 0.0f};
    }
    T = normalize(cross(U, N));
    B = cross(N, T);
}

// ---- ray depth packing, CoreRef.h:253-280 --------------------------------------------------------------------------
RT_DEV int diff_depth(uint32_t d) { return int(d & 0x7f); }
RT_DEV int spec_depth(uint32_t d) { return int(d >> 7) & 0x7f; }
RT_DEV int refr_depth(uint32_t d) { return int(d >> 14) & 0x7f; }
RT_DEV int transp_depth(uint32_t d) { return int(d >> 21) & 0x7f; }
RT_DEV int total_depth(uint32_t d) { return diff_depth(d) + spec_depth(d) + refr_depth(d) + transp_depth(d); }
RT_DEV int ray_type(uint32_t d) { return int(d >> 28) & 0xf; }
RT_DEV bool is_indirect(uint32_t d) { return (d & 0x001fffff) != 0; }
RT_DEV uint32_t pack_depth(int diff, int spec, int refr, int transp) {
    return uint32_t(diff) | (uint32_t(spec) << 7) | (uint32_t(refr) << 14) | (uint32_t(transp) << 21);
}

// ---- sampler: Owen-scrambled lookup into the PMJ02 table, CoreRef.cpp:1068-1101, 1418-1427 -------------------------
RT_DEV uint32_t hash_u32(uint32_t x) { // CoreRef.h:133-141
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
RT_DEV uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
RT_DEV uint32_t laine_karras(uint32_t x, uint32_t seed) {
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return x;
}
RT_DEV uint32_t owen_scramble(uint32_t x, uint32_t seed) { return __brev(laine_karras(__brev(x), seed)); }
RT_DEV float scramble_unorm(uint32_t seed, uint32_t val) { return float(owen_scramble(val, seed) >> 8) / 16777216.0f; }
RT_FN v2 rand2d(uint32_t dim, uint32_t seed, int sample, const uint32_t *__restrict__ seq) {
    const uint32_t sd = owen_scramble(dim, seed) & (kRandDims - 1);
    const uint32_t si = owen_scramble(uint32_t(sample), hash_combine(seed, dim)) & (kRandSamples - 1);
    const uint2 s = __ldg(reinterpret_cast<const uint2 *>(seq + sd * 2 * kRandSamples + 2 * si));
    return v2{scramble_unorm(hash_combine(seed, 2 * dim + 0), s.x), scramble_unorm(hash_combine(seed, 2 * dim + 1), s.y)};
}

} // namespace rt
