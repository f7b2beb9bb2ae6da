// Device restatement of the scalar arithmetic of the LT-removert hot path, bit-exact with the reference's
// x86-64 Release build (no FMA contraction, IEEE float/double, glibc atan2f):
//   cart2sph            ltremovert/src/utility.cpp:38-51
//   rad2deg             ltremovert/src/utility.cpp:53-56
//   pixel index         ltremovert/src/utility.cpp:118-123 == ltremovert/src/Removerter.cpp:133-138
//   transformPointCloud PCL semantics at ltremovert/src/utility.cpp:70-71, 164-165, 198-199
// Every operation is spelled with round-to-nearest intrinsics so the compiler cannot contract a
// multiply-add into an FMA regardless of -fmad; glibc's atan2f (<= 2.39) is Sun's fdlibm algorithm
// (public domain e_atan2f.c / s_atanf.c), written out here in IEEE float ops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltr {

#define LTR_DEV __device__ __forceinline__

LTR_DEV float fm(float a, float b) { return __fmul_rn(a, b); }
LTR_DEV float fa(float a, float b) { return __fadd_rn(a, b); }
LTR_DEV float fs(float a, float b) { return __fsub_rn(a, b); }
LTR_DEV float fd(float a, float b) { return __fdiv_rn(a, b); }
LTR_DEV double dm(double a, double b) { return __dmul_rn(a, b); }
LTR_DEV double da(double a, double b) { return __dadd_rn(a, b); }

LTR_DEV float ref_atanf(float x) {
    const int32_t hx = __float_as_int(x);
    const int32_t ix = hx & 0x7fffffff;
    float hi, lo;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25 (glibc 2.39 s_atanf.c; the oracle's copy is checked against the container's libm around this value)
        if (ix > 0x7f800000) return fa(x, x);
        const float r = fa(1.5707962513e+00f, 7.5497894159e-08f);
        return (hx > 0) ? r : -r;
    }
    if (ix < 0x3ee00000) {  // |x| < 0.4375
        if (ix < 0x31000000) return x;
        id = -1; hi = 0.0f; lo = 0.0f;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; hi = 4.6364760399e-01f; lo = 5.0121582440e-09f; x = fd(fs(fm(2.0f, x), 1.0f), fa(2.0f, x)); }
            else                 { id = 1; hi = 7.8539812565e-01f; lo = 3.7748947079e-08f; x = fd(fs(x, 1.0f), fa(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; hi = 9.8279368877e-01f; lo = 3.4473217170e-08f; x = fd(fs(x, 1.5f), fa(1.0f, fm(1.5f, x))); }
            else                 { id = 3; hi = 1.5707962513e+00f; lo = 7.5497894159e-08f; x = fd(-1.0f, x); }
        }
    }
    const float z = fm(x, x);
    const float w = fm(z, z);
    const float s1 = fm(z, fa(3.3333334327e-01f, fm(w, fa(1.4285714924e-01f, fm(w, fa(9.0908870101e-02f,
                     fm(w, fa(6.6610731184e-02f, fm(w, fa(4.9768779427e-02f, fm(w, 1.6285819933e-02f)))))))))));
    const float s2 = fm(w, fa(-2.0000000298e-01f, fm(w, fa(-1.1111110449e-01f, fm(w, fa(-7.6918758452e-02f,
                     fm(w, fa(-5.8335702866e-02f, fm(w, -3.6531571299e-02f)))))))));
    if (id < 0) return fs(x, fm(x, fa(s1, s2)));
    const float r = fs(hi, fs(fs(fm(x, fa(s1, s2)), lo), x));
    return (hx < 0) ? -r : r;
}

LTR_DEV float ref_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = __float_as_int(x), hy = __float_as_int(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return fa(x, y);
    if (hx == 0x3f800000) return ref_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        if (m < 2) return y;
        return (m == 2) ? fa(pi, tiny) : fs(-pi, tiny);
    }
    if (ix == 0) return (hy < 0) ? fs(-pi_o_2, tiny) : fa(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return fa(pi_o_4, tiny);
                case 1: return fs(-pi_o_4, tiny);
                case 2: return fa(fm(3.0f, pi_o_4), tiny);
                default: return fs(fm(-3.0f, pi_o_4), tiny);
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return fa(pi, tiny);
                default: return fs(-pi, tiny);
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? fs(-pi_o_2, tiny) : fa(pi_o_2, tiny);
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = fa(pi_o_2, fm(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = ref_atanf(fabsf(fd(y, x)));
    switch (m) {
        case 0: return z;
        case 1: return __int_as_float(__float_as_int(z) ^ 0x80000000);
        case 2: return fs(pi, fs(z, pi_lo));
        default: return fs(fs(z, pi_lo), pi);
    }
}

struct Sph { float az, el, r; };

// cart2sph (utility.cpp:38-51)
LTR_DEV Sph cart2sph(float x, float y, float z) {
    Sph s;
    const float xx = fm(x, x), yy = fm(y, y);
    const float rho2 = fa(xx, yy);
    s.az = ref_atan2f(y, x);
    s.el = ref_atan2f(z, __fsqrt_rn(rho2));
    s.r = __fsqrt_rn(fa(rho2, fm(z, z)));
    return s;
}

// rad2deg (utility.cpp:53-56): (float)((double)rad * 180.0 / M_PI)
LTR_DEV float rad2deg(float r) {
    return __double2float_rn(__ddiv_rn(dm((double)r, 180.0), 3.14159265358979323846));
}

// std::round(float): halfway cases away from zero, exact
LTR_DEV float round_half_away(float v) {
    const float t = truncf(v);
    const float f = fs(v, t);  // exact
    if (f >= 0.5f) return fa(t, 1.0f);
    if (f <= -0.5f) return fs(t, 1.0f);
    return t;
}

struct ImgShape { int rows, cols; float vfov, hfov; };

// pixel index (utility.cpp:118-123)
LTR_DEV void pixel_index(const Sph& s, const ImgShape& g, int* row, int* col) {
    const float rr = round_half_away(fm((float)g.rows, fs(1.0f, fd(fa(rad2deg(s.el), fd(g.vfov, 2.0f)), fs(g.vfov, 0.0f)))));
    const float cc = round_half_away(fm((float)g.cols, fd(fa(rad2deg(s.az), fd(g.hfov, 2.0f)), fs(g.hfov, 0.0f))));
    *row = (int)fminf(fmaxf(rr, 0.0f), (float)(g.rows - 1));
    *col = (int)fminf(fmaxf(cc, 0.0f), (float)(g.cols - 1));
}

// pcl::transformPointCloud, one point, 3x4 row-major double matrix m (see include/ltr_b200.h transform_order)
LTR_DEV void transform_point(const double* __restrict__ m, int order, float x, float y, float z, float* ox, float* oy, float* oz) {
    const double px = (double)x, py = (double)y, pz = (double)z;
    if (order == 0) {
        *ox = __double2float_rn(da(da(da(dm(m[0], px), dm(m[1], py)), dm(m[2], pz)), m[3]));
        *oy = __double2float_rn(da(da(da(dm(m[4], px), dm(m[5], py)), dm(m[6], pz)), m[7]));
        *oz = __double2float_rn(da(da(da(dm(m[8], px), dm(m[9], py)), dm(m[10], pz)), m[11]));
    } else {
        *ox = __double2float_rn(da(da(da(m[3], dm(px, m[0])), dm(py, m[1])), dm(pz, m[2])));
        *oy = __double2float_rn(da(da(da(m[7], dm(px, m[4])), dm(py, m[5])), dm(pz, m[6])));
        *oz = __double2float_rn(da(da(da(m[11], dm(px, m[8])), dm(py, m[9])), dm(pz, m[10])));
    }
}

// FLANN L2_Simple<float> squared distance: ((dx*dx) + dy*dy) + dz*dz with dx = query - point
LTR_DEV float l2_simple(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = fs(qx, px), dy = fs(qy, py), dz = fs(qz, pz);
    return fa(fa(fm(dx, dx), fm(dy, dy)), fm(dz, dz));
}

}  // namespace ltr
