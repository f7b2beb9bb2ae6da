// Exactness-preserving fast rejection path of the map projection (used by ltr_remove_pass / ltr_parse_projected).
//
// The reference arithmetic per (map point, keyframe) -- two double-precision transforms with float rounding,
// fdlibm atan2f twice, IEEE divisions and square roots, double rad2deg (ltremovert/src/utility.cpp:38-56, 64-72,
// 118-123) -- costs ~500 non-FMA instructions.  Almost every (point, keyframe) pair has NO effect on the result:
//   * scan-minus-map variants (HD / revert / PD): only points with scan_range - range > thres can win the per-pixel
//     atomic; static structure (range ~ scan range) and occluded points never do;
//   * true-min variants (ND, visible-point extraction): only points at most as near as the pixel's current minimum.
// So each pair is first evaluated with ~90 FMA-based float instructions whose error against the reference value is
// bounded (margins below).  The pair is DROPPED only if, for every value inside the error bound, the reference
// arithmetic would also have produced "no effect"; otherwise it is queued (warp-aggregated, shared memory) and the
// queue is drained 32 entries at a time through the bit-exact path of ref_math.cuh.  Results are therefore identical
// to the all-exact kernel; tests/test_gpu_fastpath.py checks the margins against measured deviations and
// the flags / visible points against the oracle with the fast path on and off.
#pragma once
#include "ref_math.cuh"

namespace ltr {

// Per-keyframe single-precision constants: q ~= A * (p - c_hi) - A * c_lo, c = c_hi + c_lo = sensor origin in the map frame.
// Subtracting the origin first makes the rounding error proportional to the RANGE, not to the map coordinates.
struct KfFast { float A[9]; float chi[3]; float tlo[3]; float ok; };  // 16 floats: q ~= A (p - c_hi) - t_lo; ok == 0 -> exact path only

struct FastCfg {
    float col_scale, col_off;   // V_col ~= az * col_scale + col_off   (= C * ((deg(az) + H/2) / H))
    float row_scale, row_off;   // V_row ~= row_off - el * row_scale   (= R * (1 - (deg(el) + V/2) / V))
    float m_col_a, m_col_b;     // column margin [px] = m_col_a + m_col_b * (r / rho)
    float m_row;                // row margin [px]
    float m_r_rel, m_r_abs;     // range margin [m] = m_r_abs + m_r_rel * r
    int empty_scan_shortcut;    // 1: no map point can be >= 9000 m from a keyframe -> pixels without a scan return never flag
};

// Deviation budget (validated by ltr_debug_fast_project over >1e8 samples, see test): the fast pre-round pixel
// coordinate differs from the reference's own pre-round float by
//   transform: |dq| <= ~6e-7 * r per component  -> azimuth 6e-7 * r/rho rad, elevation ~1e-6 rad
//   atan (degree-8 minimax in a^2, approximate reciprocal): <= 3e-7 rad;  quadrant fix-ups <= 3e-7 rad
//   reference's own float chain (rad2deg rounding, (x + H/2)/H, scaling): <= 4e-7 * C px
// Margins are set >= 3x the measured maxima (measured: deviation / margin <= 0.26 / 0.23 / 0.21 for column / row / range).
__host__ inline FastCfg make_fast_cfg(int rows, int cols, float vfov, float hfov, int empty_scan_shortcut) {
    FastCfg f;
    const double kPi = 3.14159265358979323846;
    const double ppr_c = (double)cols * 180.0 / (kPi * (double)hfov);   // pixels per radian
    const double ppr_r = (double)rows * 180.0 / (kPi * (double)vfov);
    f.col_scale = (float)ppr_c; f.col_off = 0.5f * (float)cols;
    f.row_scale = (float)ppr_r; f.row_off = 0.5f * (float)rows;
    f.m_col_a = (float)(ppr_c * 1.2e-6 + (double)cols * 4.0e-7);
    f.m_col_b = (float)(ppr_c * 1.2e-6);
    f.m_row = (float)(ppr_r * 2.4e-6 + (double)rows * 4.0e-7);
    f.m_r_rel = 2.0e-6f; f.m_r_abs = 5.0e-6f;
    f.empty_scan_shortcut = empty_scan_shortcut;
    return f;
}

// single-instruction approximations (MUFU) without the denormal fix-up sequences of rsqrtf()/__fdividef()
LTR_DEV float mufu_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
LTR_DEV float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// atan(a) for a in [0, 1]: a * P(a^2), degree-8 least-squares/Chebyshev fit, |err| <= 1.1e-7 in f32 Horner form
LTR_DEV float fast_atan01(float a) {
    const float z = __fmul_rn(a, a);
    float p = 0.0028340641874819994f;
    p = __fmaf_rn(p, z, -0.016005029901862144f);
    p = __fmaf_rn(p, z, 0.042587608098983765f);
    p = __fmaf_rn(p, z, -0.07495445758104324f);
    p = __fmaf_rn(p, z, 0.10636754333972931f);
    p = __fmaf_rn(p, z, -0.14202570915222168f);
    p = __fmaf_rn(p, z, 0.19992484152317047f);
    p = __fmaf_rn(p, z, -0.3333306610584259f);
    p = __fmaf_rn(p, z, 1.0f);
    return __fmul_rn(p, a);
}

struct FastProj { float az, el, r, rho_inv_r; };  // rho_inv_r = r / rho

// approximate cart2sph of a point in the sensor frame
LTR_DEV FastProj fast_sph(float qx, float qy, float qz) {
    const float rho2 = __fmaf_rn(qy, qy, __fmul_rn(qx, qx));
    const float r2 = __fmaf_rn(qz, qz, rho2);
    const float irho = mufu_rsq(rho2), ir = mufu_rsq(r2);
    const float rho = __fmul_rn(rho2, irho);
    FastProj o;
    o.r = __fmul_rn(r2, ir);
    o.rho_inv_r = __fmul_rn(o.r, irho);
    // azimuth = atan2(qy, qx); 0/0 -> NaN (NaN always routes to the exact path)
    {
        const float ax = fabsf(qx), ay = fabsf(qy);
        float t = fast_atan01(__fmul_rn(fminf(ax, ay), mufu_rcp(fmaxf(ax, ay))));
        if (ay > ax) t = __fsub_rn(1.57079632679f, t);
        if (qx < 0.0f) t = __fsub_rn(3.14159265359f, t);
        o.az = copysignf(t, qy);
    }
    // elevation = atan2(qz, rho), rho >= 0
    {
        const float az_ = fabsf(qz);
        float t = fast_atan01(__fmul_rn(fminf(rho, az_), mufu_rcp(fmaxf(rho, az_))));
        if (az_ > rho) t = __fsub_rn(1.57079632679f, t);
        o.el = copysignf(t, qz);
    }
    return o;
}

// kf: A[0..8], c_hi[9..11], t_lo[12..14] = A * c_lo, ok[15]:  q ~= A * (p - c_hi) - t_lo
LTR_DEV FastProj fast_project(const float* __restrict__ kf, float x, float y, float z) {
    const float dx = __fsub_rn(x, kf[9]), dy = __fsub_rn(y, kf[10]), dz = __fsub_rn(z, kf[11]);
    const float qx = __fmaf_rn(kf[2], dz, __fmaf_rn(kf[1], dy, __fmaf_rn(kf[0], dx, -kf[12])));
    const float qy = __fmaf_rn(kf[5], dz, __fmaf_rn(kf[4], dy, __fmaf_rn(kf[3], dx, -kf[13])));
    const float qz = __fmaf_rn(kf[8], dz, __fmaf_rn(kf[7], dy, __fmaf_rn(kf[6], dx, -kf[14])));
    return fast_sph(qx, qy, qz);
}

// Rounds a pre-round pixel coordinate when that is certain: returns true and the clamped index iff every value within
// +-margin of v rounds (half away from zero) to the same integer.  NaN -> false.  |v| < 2^22 (pixel coordinates).
LTR_DEV bool certain_round(float v, float margin, int hi, int* idx) {
    const float kMagic = 12582912.0f;                         // 1.5 * 2^23: v + kMagic has the nearest integer in its low mantissa bits
    const float y = __fadd_rn(v, kMagic);
    const float n = __fsub_rn(y, kMagic);                     // nearest integer (ties to even; ties are inside the margin anyway)
    const float d = fabsf(__fsub_rn(v, n));                   // distance to the nearest integer, in [0, 0.5]
    *idx = min(max(__float_as_int(y) - 0x4B400000, 0), hi);
    return d < __fsub_rn(0.5f, margin);                       // i.e. farther than `margin` from the rounding boundary n +- 0.5
}

constexpr uint32_t kNoPointBitsF = 0x461C4000u;  // 10000.0f

// queue entry: [63:32] map index | [31:14] pixel (row * cols + col, < 2^18) | [13:0] keyframe within the batch
LTR_DEV uint64_t q_pack(uint32_t i, uint32_t px, uint32_t k) { return ((uint64_t)i << 32) | (px << 14) | k; }

// Kind A: pixel index known to be right, only the exact range is needed (double transform + sqrt): ~60 instructions.
template <bool kCandidatesOnly>
LTR_DEV void exact_range_pair(const PtrView& map, uint64_t e, const double* __restrict__ poses, int kf0, const double* __restrict__ ext,
                              int ext_identity, int order, uint32_t npx, const uint32_t* __restrict__ scan_rimg, float thres,
                              uint64_t* __restrict__ win, int shortcut, unsigned* n_atomics) {
    const uint32_t i = (uint32_t)(e >> 32), pxl = ((uint32_t)e >> 14) & 0x3ffffu, k = (uint32_t)e & 0x3fffu;
    float lx, ly, lz;
    transform_point(poses + (size_t)(kf0 + k) * 24, order, map.x[i], map.y[i], map.z[i], &lx, &ly, &lz);
    if (!ext_identity) transform_point(ext, order, lx, ly, lz, &lx, &ly, &lz);
    const float r = __fsqrt_rn(fa(fa(fm(lx, lx), fm(ly, ly)), fm(lz, lz)));   // == cart2sph().r
    const size_t px = (size_t)k * npx + pxl;
    const uint64_t packed = ((uint64_t)__float_as_uint(r) << 32) | i;
    if (kCandidatesOnly) {
        const uint32_t sb = scan_rimg[px];
        if (shortcut && sb == kNoPointBitsF && r < 9000.0f) return;
        if (fs(__uint_as_float(sb), r) > thres && packed < win[px]) { atomicMin((unsigned long long*)&win[px], (unsigned long long)packed); ++*n_atomics; }
    } else {
        if (packed < win[px]) { atomicMin((unsigned long long*)&win[px], (unsigned long long)packed); ++*n_atomics; }
    }
}

// Kind B: nothing is known (pixel near a rounding boundary, degenerate geometry, keyframe without fast constants):
// the full reference arithmetic, identical to map_project_kernel's body.
template <bool kCandidatesOnly>
__device__ __noinline__ void exact_full_pair(const PtrView& map, uint64_t e, const double* __restrict__ poses, int kf0,
                                             const double* __restrict__ ext, int ext_identity, int order, const ImgShape& g,
                                             const uint32_t* __restrict__ scan_rimg, float thres, uint64_t* __restrict__ win, int shortcut,
                                             unsigned* n_atomics) {
    const uint32_t i = (uint32_t)(e >> 32), k = (uint32_t)e & 0x3fffu;
    float lx, ly, lz;
    transform_point(poses + (size_t)(kf0 + k) * 24, order, map.x[i], map.y[i], map.z[i], &lx, &ly, &lz);
    if (!ext_identity) transform_point(ext, order, lx, ly, lz, &lx, &ly, &lz);
    const Sph s = cart2sph(lx, ly, lz);
    int r, c;
    pixel_index(s, g, &r, &c);
    const size_t px = (size_t)k * g.rows * g.cols + (size_t)r * g.cols + c;
    const uint64_t packed = ((uint64_t)__float_as_uint(s.r) << 32) | i;
    if (kCandidatesOnly) {
        const uint32_t sb = scan_rimg[px];
        if (shortcut && sb == kNoPointBitsF && s.r < 9000.0f) return;  // empty scan pixel: diff > 200 for every map point there
        if (fs(__uint_as_float(sb), s.r) > thres && packed < win[px]) { atomicMin((unsigned long long*)&win[px], (unsigned long long)packed); ++*n_atomics; }
    } else {
        if (packed < win[px]) { atomicMin((unsigned long long*)&win[px], (unsigned long long)packed); ++*n_atomics; }
    }
}

}  // namespace ltr
#include "project_cull.cuh"
namespace ltr {

constexpr int kFastThreads = 256;
constexpr int kFastPts = 4;      // map points per thread (registers), strided by the block size for coalescing
constexpr int kQueueCap = 160;   // per warp and kind: < 32 left over + 4 * 32 pushed per keyframe step

// `approx_min` (u32 per pixel, float bits, initialised to +inf): running minimum of the APPROXIMATE ranges of the pairs
// that were queued for pixel px.  A pair whose approximate range exceeds it by more than twice the range margin is
// provably farther than an already-queued pair and cannot win the pixel, so it is dropped without exact arithmetic.
//
// Hot loop budget (ncu, r01): the kernel is issue-bound, so the common "no effect" path is kept branch-free: one
// gather from the scan image (HD/PD) or from approx_min (ND / visible points), compare, done.  Rare pairs reserve a
// queue slot with a shared-memory atomic (no ballots on the common path).
template <bool kCandidatesOnly>
__global__ void __launch_bounds__(kFastThreads, 4) map_project_fast_kernel(PtrView map, const float* __restrict__ kf_fast, const double* __restrict__ poses,
                                                                        int kf0, int nb, const double* __restrict__ ext, int ext_identity, int order,
                                                                        ImgShape g, FastCfg fc, const uint32_t* __restrict__ scan_rimg, float thres,
                                                                        uint64_t* __restrict__ win, uint32_t* __restrict__ approx_min,
                                                                        CullArgs ca, unsigned long long* __restrict__ counters,
                                                                        unsigned int* __restrict__ work_counter) {
    extern __shared__ float s_dyn[];
    float* s_kf = s_dyn;                                                                    // nb * 16
    const int warp = threadIdx.x >> 5;
    uint64_t* s_qa = (uint64_t*)(s_dyn + ((nb * 16 + 3) & ~3)) + warp * (2 * kQueueCap);
    uint64_t* s_qb = s_qa + kQueueCap;
    __shared__ int s_cnt[kFastThreads / 32][2];
    for (int t = threadIdx.x; t < nb * 16; t += blockDim.x) s_kf[t] = kf_fast[(size_t)kf0 * 16 + t];
    if ((threadIdx.x & 31) < 2) s_cnt[warp][threadIdx.x & 31] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const uint32_t npx = (uint32_t)(g.rows * g.cols);
    const int hi_c = g.cols - 1, hi_r = g.rows - 1;
    // folded constants
    const float kMagic = 12582912.0f;
    const float lim_c = 0.5f - fc.m_col_a, lim_r = 0.5f - fc.m_row;
    const float thr_lo = thres - fc.m_r_abs, rel_m1 = fc.m_r_rel - 1.0f;
    const float mr2_rel = 2.0f * fc.m_r_rel, mr2_abs = 2.0f * fc.m_r_abs;
    const uint32_t empty_bits = fc.empty_scan_shortcut ? 0xff800000u : kNoPointBitsF;
    unsigned n_a = 0, n_b = 0, n_atomics = 0, n_culled = 0;
    // Persistent warps: each warp claims 128-point tiles from a global counter until the map is exhausted, so warps whose
    // tiles are culled in many keyframes do not leave issue slots idle while their CTA-mates finish.
    const long long total_tiles = (map.n + (32 * kFastPts) - 1) / (32 * kFastPts);
    for (;;) {
    unsigned tile_u = 0;
    if (lane == 0) tile_u = atomicAdd(work_counter, 1u);
    tile_u = __shfl_sync(0xffffffffu, tile_u, 0);
    const long long tile = (long long)tile_u;
    if (tile >= total_tiles) break;
    // lane l holds points tile * 128 + 32 j + l, j = 0..3 (coalesced per j)
    const int64_t base = (int64_t)tile * (32 * kFastPts) + lane;
    float px_[kFastPts], py_[kFastPts], pz_[kFastPts];
    uint32_t pi_[kFastPts];
#pragma unroll
    for (int j = 0; j < kFastPts; ++j) {
        // the last tile re-processes point n-1 in its padding lanes: harmless (atomicMin of an identical key)
        const int64_t i = min(base + (int64_t)j * 32, map.n - 1);
        pi_[j] = (uint32_t)i;
        px_[j] = map.x[i]; py_[j] = map.y[i]; pz_[j] = map.z[i];
    }
    // Tile culling (project_cull.cuh): lane l tests this 128-point tile against keyframe l of the launch; the ballot is the
    // set of keyframes in which no point of the tile can be a candidate -- those iterations of the k loop are skipped.
    unsigned cull_mask = 0u;
    if (kCandidatesOnly && ca.enabled) {
        const float4 t = ca.tiles[tile];
        const bool c = ((int)lane < nb) && tile_culled(s_kf + 16 * lane, t, fc, g, ca, (int)lane, thres);
        cull_mask = __ballot_sync(0xffffffffu, c);
        if (lane == 0) n_culled += __popc(cull_mask);
    }
    for (int k = 0; k < nb; ++k) {
        if ((cull_mask >> k) & 1u) continue;   // warp-uniform
        float kf[16];
        {
            const float4* s4 = reinterpret_cast<const float4*>(s_kf + 16 * k);
            const float4 a = s4[0], b = s4[1], c = s4[2], d = s4[3];
            kf[0] = a.x; kf[1] = a.y; kf[2] = a.z; kf[3] = a.w; kf[4] = b.x; kf[5] = b.y; kf[6] = b.z; kf[7] = b.w;
            kf[8] = c.x; kf[9] = c.y; kf[10] = c.z; kf[11] = c.w; kf[12] = d.x; kf[13] = d.y; kf[14] = d.z; kf[15] = d.w;
        }
        const bool kf_ok = kf[15] != 0.0f;  // warp-uniform
        const uint32_t kbase = (uint32_t)k * npx;   // 32-bit pixel offsets: nb * npx < 2^31 (checked by the host)
#pragma unroll
        for (int j = 0; j < kFastPts; ++j) {
            const FastProj f = fast_project(kf, px_[j], py_[j], pz_[j]);
            const float vcol = __fmaf_rn(f.az, fc.col_scale, fc.col_off);
            const float vrow = __fmaf_rn(-f.el, fc.row_scale, fc.row_off);
            const float yc = __fadd_rn(vcol, kMagic), yr = __fadd_rn(vrow, kMagic);
            const float dc = fabsf(__fsub_rn(vcol, __fsub_rn(yc, kMagic)));   // distance to the nearest integer
            const float dr = fabsf(__fsub_rn(vrow, __fsub_rn(yr, kMagic)));
            const int c = min(max(__float_as_int(yc) - 0x4B400000, 0), hi_c);
            const int r = min(max(__float_as_int(yr) - 0x4B400000, 0), hi_r);
            // certain <=> both coordinates are farther than their margin from a rounding boundary (false for NaN)
            const bool certain = kf_ok & (__fmaf_rn(fc.m_col_b, f.rho_inv_r, dc) < lim_c) & (dr < lim_r);
            const uint32_t pxl = (uint32_t)(r * g.cols + c);
            const uint32_t idx = kbase + pxl;
            if (certain) {
                bool maybe = true;
                if (kCandidatesOnly) {
                    uint32_t sb = scan_rimg[idx];
                    if (sb == kNoPointBitsF) sb = empty_bits;   // -inf when pixels without a scan return can never flag
                    // candidate for SOME range within the margin: scan - range > thres   (sr - r + m_rel r >= thres - m_abs)
                    maybe = !(__fmaf_rn(f.r, rel_m1, __uint_as_float(sb)) < thr_lo);
                }
                if (maybe) {
                    const float cur = __uint_as_float(approx_min[idx]);
                    if (!(f.r > __fadd_rn(cur, __fmaf_rn(mr2_rel, f.r, mr2_abs)))) {   // not provably farther than an already-queued pair
                        if (f.r < cur) atomicMin(&approx_min[idx], __float_as_uint(f.r));
                        s_qa[atomicAdd(&s_cnt[warp][0], 1)] = q_pack(pi_[j], pxl, (uint32_t)k);
                    }
                }
            } else {
                s_qb[atomicAdd(&s_cnt[warp][1], 1)] = q_pack(pi_[j], 0u, (uint32_t)k);
            }
        }
        __syncwarp();
        int qa = *(volatile int*)&s_cnt[warp][0];
        while (qa >= 32) {
            qa -= 32;
            const uint64_t e = s_qa[qa + lane];
            exact_range_pair<kCandidatesOnly>(map, e, poses, kf0, ext, ext_identity, order, npx, scan_rimg, thres, win, fc.empty_scan_shortcut, &n_atomics);
            ++n_a;
        }
        int qb = *(volatile int*)&s_cnt[warp][1];
        while (qb >= 32) {
            qb -= 32;
            const uint64_t e = s_qb[qb + lane];
            exact_full_pair<kCandidatesOnly>(map, e, poses, kf0, ext, ext_identity, order, g, scan_rimg, thres, win, fc.empty_scan_shortcut, &n_atomics);
            ++n_b;
        }
        __syncwarp();
        if (lane == 0) { s_cnt[warp][0] = qa; s_cnt[warp][1] = qb; }
        __syncwarp();
    }
    }  // persistent tile loop
    const int qa = *(volatile int*)&s_cnt[warp][0], qb = *(volatile int*)&s_cnt[warp][1];
    if ((int)lane < qa) {
        exact_range_pair<kCandidatesOnly>(map, s_qa[lane], poses, kf0, ext, ext_identity, order, npx, scan_rimg, thres, win, fc.empty_scan_shortcut, &n_atomics);
        ++n_a;
    }
    if ((int)lane < qb) {
        exact_full_pair<kCandidatesOnly>(map, s_qb[lane], poses, kf0, ext, ext_identity, order, g, scan_rimg, thres, win, fc.empty_scan_shortcut, &n_atomics);
        ++n_b;
    }
    // statistics: [0] pairs through the exact-range path, [1] atomics on the winner image, [2] pairs through the full exact path, [3] culled pairs
    for (int o = 16; o > 0; o >>= 1) {
        n_a += __shfl_down_sync(0xffffffffu, n_a, o); n_b += __shfl_down_sync(0xffffffffu, n_b, o); n_atomics += __shfl_down_sync(0xffffffffu, n_atomics, o);
    }
    if (lane == 0 && counters) {
        atomicAdd(&counters[0], (unsigned long long)n_a); atomicAdd(&counters[1], (unsigned long long)n_atomics); atomicAdd(&counters[2], (unsigned long long)n_b);
        if (n_culled) atomicAdd(&counters[3], (unsigned long long)n_culled * (32ull * kFastPts));   // pairs skipped by tile culling
    }
}

}  // namespace ltr
