// ORACLE (test infrastructure, not product code).
// CPU restatement of the scalar arithmetic on the LT-removert hot path of gisbi-kim/lt-mapper.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything under oracle/.  The reference ships no tests or golden vectors.
// The first-party arithmetic here (cart2sph, rad2deg, pixel index, image size) is pinned bit-for-bit
// against the reference's own utility.cpp compiled behind stand-in headers (oracle/ref_shim,
// tests/test_ref_pin.py), against this container's libm (atan2f) and against hand-computed known
// answers (tests/test_oracle_kat.py).  PARITY UNPINNED for the PCL transform restated at the end.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Build with -ffp-contract=off and no -march flag: the reference is a plain x86-64 Release build
// (ltremovert/CMakeLists.txt:4-5), i.e. SSE2 scalar IEEE float/double with no FMA contraction.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace ltr_oracle {

// ---------------------------------------------------------------------------------------------
// atan2f: the reference calls std::atan2(float,float) (ltremovert/src/utility.cpp:46-47), i.e.
// glibc's atan2f.  glibc <= 2.39 implements it with Sun's fdlibm algorithm (e_atan2f.c /
// s_atanf.c, public domain).  Restated here in plain IEEE float ops so that the same algorithm
// can be written for the device; tests assert bit-equality with the container's libm.
// ---------------------------------------------------------------------------------------------
static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static inline float ref_atanf(float x) {
    static const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    static const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    static const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                                 9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                                 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25 (glibc 2.39's s_atanf.c; measured on this container's libm: atanf(2^25 - 2) takes the
                             // polynomial path, atanf(2^25) returns atanhi[3] + atanlo[3]; FreeBSD's msun uses 2^26 here)
        if (ix > 0x7f800000) return x + x;  // NaN
        return (hx > 0) ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {          // |x| < 0.4375
        if (ix < 0x31000000) return x;  // |x| < 2^-29
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3f980000) {      // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }   // 7/16 <= |x| < 11/16
            else                 { id = 1; x = (x - 1.0f) / (x + 1.0f); }          // 11/16 <= |x| < 19/16
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }   // |x| < 2.4375
            else                 { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (hx < 0) ? -r : r;
}

static inline float ref_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;  // NaN
    if (hx == 0x3f800000) return ref_atanf(y);              // x == 1.0
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);      // 2*sign(x) + sign(y)
    if (iy == 0) {
        switch (m) {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return pi + tiny;
                default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = ref_atanf(std::fabs(y / x));
    switch (m) {
        case 0: return z;
        case 1: return u2f(f2u(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// ---------------------------------------------------------------------------------------------
// cart2sph (ltremovert/src/utility.cpp:38-51) with SphericalPoint (utility.h:96-101)
// ---------------------------------------------------------------------------------------------
struct Sph { float az, el, r; };

template <bool kUseLibm>
static inline Sph cart2sph_t(float x, float y, float z) {
    Sph s;
    const float rho2 = x * x + y * y;
    if (kUseLibm) {
        s.az = atan2f(y, x);
        s.el = atan2f(z, sqrtf(rho2));
    } else {
        s.az = ref_atan2f(y, x);
        s.el = ref_atan2f(z, sqrtf(rho2));
    }
    s.r = sqrtf(x * x + y * y + z * z);
    return s;
}
static inline Sph cart2sph(float x, float y, float z) { return cart2sph_t<false>(x, y, z); }

// rad2deg (utility.cpp:53-56): float -> double mul, double div by M_PI, rounded back to float.
static inline float rad2deg(float radians) { return (float)((double)radians * 180.0 / M_PI); }

// resetRimgSize (utility.cpp:222-236): int = std::round(float * float)
static inline void resetRimgSize(float vfov, float hfov, float alpha, int* rows, int* cols) {
    *rows = (int)std::round(vfov * alpha);
    *cols = (int)std::round(hfov * alpha);
}

// Pixel index (utility.cpp:118-123 == Removerter.cpp:133-138); all f32 except rad2deg.
static inline void pixelIndex(const Sph& s, float vfov, float hfov, int rows, int cols, int* row, int* col) {
    const float lower = 0.0f;
    const float r = std::round(rows * (1 - (rad2deg(s.el) + (vfov / float(2.0))) / (vfov - float(0.0))));
    const float c = std::round(cols * ((rad2deg(s.az) + (hfov / float(2.0))) / (hfov - float(0.0))));
    *row = int(std::min(std::max(r, lower), float(rows - 1)));
    *col = int(std::min(std::max(c, lower), float(cols - 1)));
}

// ---------------------------------------------------------------------------------------------
// pcl::transformPointCloud(in, out, Eigen::Matrix4d) on PointXYZI (call sites utility.cpp:70-71,
// 164-165, 184-185, 198-199).  PCL is NOT under /root/reference -> semantics restated from the
// published PCL sources (UNPINNED): double math, result rounded to float.
//   order 0: ((m00*x + m01*y) + m02*z) + m03   PCL <= 1.9 generic code path
//   order 1: ((m03 + m00*x) + m01*y) + m02*z   PCL >= 1.10 SSE2 pcl::detail::Transformer<double>::se3
// m is row-major 4x4.
// ---------------------------------------------------------------------------------------------
static inline void transformPoint(const double* m, int order, float x, float y, float z, float* ox, float* oy, float* oz) {
    const double px = x, py = y, pz = z;
    if (order == 0) {
        *ox = (float)(m[0] * px + m[1] * py + m[2] * pz + m[3]);
        *oy = (float)(m[4] * px + m[5] * py + m[6] * pz + m[7]);
        *oz = (float)(m[8] * px + m[9] * py + m[10] * pz + m[11]);
    } else {
        *ox = (float)(((m[3] + px * m[0]) + py * m[1]) + pz * m[2]);
        *oy = (float)(((m[7] + px * m[4]) + py * m[5]) + pz * m[6]);
        *oz = (float)(((m[11] + px * m[8]) + py * m[9]) + pz * m[10]);
    }
}

}  // namespace ltr_oracle
