// libltr_b200.so -- range-image projection kernels: remove / revert / ND / PD pass and visible-point extraction.
//
// Reference functions replaced (paths relative to the lt-mapper repository):
//   Removerter::scan2RangeImg                                   ltremovert/src/Removerter.cpp:109-156
//   transformGlobalMapToLocal + map2RangeImg                    ltremovert/src/utility.cpp:64-72, 92-142
//   range diff + calcDescrepancyAndParseDynamicPointIdx         ltremovert/src/Removerter.cpp:572/459/516, 381-413
//   calcDescrepancyAndParseDynamicPointIdxForEachScan{,ND,PD}   ltremovert/src/Removerter.cpp:542-593, 485-540, 429-482
//   parseProjectedPoints / Session::parseScansViaProjection     ltremovert/src/utility.cpp:74-89, ltremovert/src/Session.cpp:348-360
//
// Determinism: the reference's per-pixel min is a racy OpenMP loop; its sequential meaning (min range, lowest
// index among equal ranges) is obtained here with one 64-bit atomicMin on (float_bits(range) << 32 | index).
#include "ltr_internal.cuh"
#include "ref_math.cuh"
#include "project_fast.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace ltr {

constexpr uint32_t kNoPointBits = 0x461C4000u;                       // 10000.0f == kFlagNoPOINT (utility.h:93)
constexpr uint64_t kWinEmpty = ((uint64_t)kNoPointBits << 32);        // (range 10000, index 0) == utility.cpp:103-104
constexpr uint64_t kWinNone = ~0ull;                                  // "no candidate" for the scan-minus-map variants
constexpr float kValidDiffUpperBound = 200.0f;                        // utility.h:94

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, uint32_t v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_u64_kernel(uint64_t* __restrict__ p, uint64_t v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__device__ __forceinline__ int find_kf_rel(const int64_t* __restrict__ off, int nb, int64_t i) {
    int lo = 0, hi = nb;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

// scan2RangeImg for the keyframes [kf0, kf0+nb): per-pixel min range via 32-bit atomicMin on the float bits (ranges >= 0).
__global__ void __launch_bounds__(256) scan_rimg_kernel(PtrView scans, const int64_t* __restrict__ off, int kf0, int nb, ImgShape g,
                                                        uint32_t* __restrict__ rimg) {
    const int64_t begin = off[kf0], end = off[kf0 + nb];
    const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const int k = find_kf_rel(off + kf0, nb, i);
    const Sph s = cart2sph(scans.x[i], scans.y[i], scans.z[i]);
    int r, c;
    pixel_index(s, g, &r, &c);
    atomicMin(&rimg[(size_t)k * g.rows * g.cols + (size_t)r * g.cols + c], __float_as_uint(s.r));
}

// The same image through the fast pixel evaluation of project_fast.cuh.  Scan points are already in the sensor frame, so only the
// angle polynomials are approximate; the column / row are accepted when every value within the margins (the ones validated for the map
// projection, which additionally cover a transform error that does not exist here) rounds to the same pixel, otherwise -- pixel
// boundary, point on the z axis -- the reference arithmetic decides.  The range written is always the exact one.
template <bool kElDirect>
__global__ void __launch_bounds__(256) scan_rimg_fast_kernel(PtrView scans, const int64_t* __restrict__ off, int kf0, int nb, ImgShape g, FastCfg fc,
                                                             uint32_t* __restrict__ rimg) {
    const int64_t begin = off[kf0], end = off[kf0 + nb];
    const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const int k = find_kf_rel(off + kf0, nb, i);
    const float x = scans.x[i], y = scans.y[i], z = scans.z[i];
    const FastProj f = fast_sph<kElDirect>(x, y, z);
    int r, c;
    const bool okc = certain_round(__fmaf_rn(f.az, fc.col_scale, fc.col_off), __fmaf_rn(fc.m_col_b, f.rho_inv_r, fc.m_col_a), g.cols - 1, &c);
    const bool okr = certain_round(__fmaf_rn(f.el, fc.neg_row_scale, fc.row_off), fc.m_row, g.rows - 1, &r);
    float range;
    if (okc & okr) {
        range = __fsqrt_rn(fa(fa(fm(x, x), fm(y, y)), fm(z, z)));   // cart2sph's r (utility.cpp:48)
    } else {
        const Sph s = cart2sph(x, y, z);
        pixel_index(s, g, &r, &c);
        range = s.r;
    }
    atomicMin(&rimg[(size_t)k * g.rows * g.cols + (size_t)r * g.cols + c], __float_as_uint(range));
}

// test image of the scan-minus-map fast path (project_fast.cuh): the scan image with "no return" pixels set to -inf, which makes
// "scan - range > thres" false for every range -- valid when no map point can be ~10000 m away (empty_scan_shortcut_ok)
__global__ void __launch_bounds__(256) scan_test_image_kernel(const uint32_t* __restrict__ rimg, uint32_t* __restrict__ test_img, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = rimg[i];
        test_img[i] = (b == kNoPointBits) ? 0xff800000u : b;
    }
}

// Exact projection of every map point into every keyframe of the batch.
//   kCandidatesOnly = true  (HD / revert / PD, diff = scan - map): by monotonicity of f32 subtraction the set
//       {points with scan - range > thres} is a prefix in range order of the pixel's points, so the pixel winner among
//       them equals the global pixel winner whenever any exists; only those points touch the atomic.
//   kCandidatesOnly = false (ND, visible-point extraction): true per-pixel minimum with a read-before-atomic filter.
template <bool kCandidatesOnly>
__global__ void __launch_bounds__(256) map_project_kernel(PtrView map, const double* __restrict__ poses, int kf0, int nb,
                                                          const double* __restrict__ ext, int ext_identity, int order, ImgShape g,
                                                          const uint32_t* __restrict__ scan_rimg, float thres, uint64_t* __restrict__ win) {
    extern __shared__ double s_pose[];  // nb * 12: inverse poses of the batch
    for (int t = threadIdx.x; t < nb * 12; t += blockDim.x) s_pose[t] = poses[(size_t)(kf0 + t / 12) * 24 + (t % 12)];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= map.n) return;
    const float x = map.x[i], y = map.y[i], z = map.z[i];
    const size_t npx = (size_t)g.rows * g.cols;
    for (int k = 0; k < nb; ++k) {
        float lx, ly, lz;
        transform_point(s_pose + 12 * k, order, x, y, z, &lx, &ly, &lz);        // utility.cpp:70
        if (!ext_identity) transform_point(ext, order, lx, ly, lz, &lx, &ly, &lz);  // utility.cpp:71
        const Sph s = cart2sph(lx, ly, lz);
        int r, c;
        pixel_index(s, g, &r, &c);
        const size_t px = (size_t)k * npx + (size_t)r * g.cols + c;
        const uint64_t packed = ((uint64_t)__float_as_uint(s.r) << 32) | (uint32_t)i;
        if (kCandidatesOnly) {
            const float sr = __uint_as_float(scan_rimg[px]);
            if (fs(sr, s.r) > thres) atomicMin((unsigned long long*)&win[px], (unsigned long long)packed);
        } else {
            if (packed < win[px]) atomicMin((unsigned long long*)&win[px], (unsigned long long)packed);
        }
    }
}

// calcDescrepancyAndParseDynamicPointIdx (Removerter.cpp:381-413) over the batch images; also re-arms `win`.
template <bool kCandidatesOnly>
__global__ void __launch_bounds__(256) resolve_kernel(const uint32_t* __restrict__ scan_rimg, uint64_t* __restrict__ win, int64_t total_px,
                                                      float thres, uint8_t* __restrict__ flags) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total_px) return;
    const uint64_t w = win[p];
    if (kCandidatesOnly) {
        if (w == kWinNone) return;
        win[p] = kWinNone;
        const float mr = __uint_as_float((uint32_t)(w >> 32));
        const float diff = fs(__uint_as_float(scan_rimg[p]), mr);  // scan - map (Removerter.cpp:572, 459)
        if (diff < kValidDiffUpperBound && diff > thres) flags[(uint32_t)w] = 1;
    } else {
        if (w != kWinEmpty) win[p] = kWinEmpty;
        const float mr = __uint_as_float((uint32_t)(w >> 32));
        const float diff = fs(mr, __uint_as_float(scan_rimg[p]));  // map - scan (Removerter.cpp:516)
        if (diff < kValidDiffUpperBound && diff > thres) flags[(uint32_t)w] = 1;
    }
}

// parseProjectedPoints (utility.cpp:80-87): row-major scan of the index image, skipping index 0.  Each keyframe's image is cut
// into chunks of kParseChunk pixels, one block per (chunk, keyframe): a count kernel, then a scatter kernel that starts at the sum of
// the preceding chunks' counts and compacts its chunk in pixel order (and resets the winner image for the next batch).
constexpr int kParseChunk = 8192;

__global__ void __launch_bounds__(1024) parse_count_kernel(const uint64_t* __restrict__ win, int npx, int nchunk, unsigned int* __restrict__ chunk_cnt) {
    __shared__ int s_warp[32];
    const int k = blockIdx.y, c = blockIdx.x;
    const uint64_t* w = win + (size_t)k * npx;
    const int p1 = min((c + 1) * kParseChunk, npx);
    int n = 0;
    for (int p = c * kParseChunk + threadIdx.x; p < p1; p += blockDim.x) n += ((uint32_t)w[p] != 0u) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x < 32) {
        n = s_warp[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
        if (threadIdx.x == 0) chunk_cnt[(size_t)k * nchunk + c] = (unsigned int)n;
    }
}

__global__ void __launch_bounds__(1024) parse_compact_kernel(uint64_t* __restrict__ win, int npx, int nchunk, const unsigned int* __restrict__ chunk_cnt,
                                                             uint32_t* __restrict__ list, unsigned int* __restrict__ count) {
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const int k = blockIdx.y, c = blockIdx.x;
    uint64_t* w = win + (size_t)k * npx;
    uint32_t* out = list + (size_t)k * npx;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {   // where this chunk's output starts: the counts of the chunks before it
        int b = 0;
        for (int j = (int)lane; j < c; j += 32) b += (int)chunk_cnt[(size_t)k * nchunk + j];
        for (int o = 16; o > 0; o >>= 1) b += __shfl_down_sync(0xffffffffu, b, o);
        if (lane == 0) s_base = b;
    }
    __syncthreads();
    const int p1 = min((c + 1) * kParseChunk, npx);
    for (int p0 = c * kParseChunk; p0 < p1; p0 += blockDim.x) {
        const int p = p0 + threadIdx.x;
        uint32_t idx = 0;
        if (p < p1) { const uint64_t v = w[p]; idx = (uint32_t)v; if (v != kWinEmpty) w[p] = kWinEmpty; }
        const bool keep = idx != 0;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        const int wrank = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        if (warp == 0) {
            const int v = s_warp[lane];
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
            s_warp[lane] = incl - v;
            if (lane == 31) s_warp[32] = incl;
        }
        __syncthreads();
        if (keep) out[s_base + s_warp[warp] + wrank] = idx;
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[32];
        __syncthreads();
    }
    if (c == nchunk - 1 && threadIdx.x == 0) count[k] = (unsigned int)s_base;
}

// emits map_local[ptidx] for every listed index: exact two-step transform of the winning map point
__global__ void __launch_bounds__(256) parse_emit_kernel(PtrView map, const double* __restrict__ poses, int kf_begin, const int64_t* __restrict__ out_off,
                                                         int K, const uint32_t* __restrict__ list, int npx, const double* __restrict__ ext,
                                                         int ext_identity, int order, DevCloud out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= out.n) return;
    const int k = find_kf_rel(out_off, K, j);
    const uint32_t idx = list[(size_t)k * npx + (size_t)(j - out_off[k])];
    float x = map.x[idx], y = map.y[idx], z = map.z[idx];
    transform_point(poses + (size_t)(kf_begin + k) * 24, order, x, y, z, &x, &y, &z);
    if (!ext_identity) transform_point(ext, order, x, y, z, &x, &y, &z);
    out.x()[j] = x; out.y()[j] = y; out.z()[j] = z; out.i()[j] = map.i[idx];
}

__global__ void debug_pixel_kernel(const float* __restrict__ xyz, int64_t n, ImgShape g, int* __restrict__ row, int* __restrict__ col,
                                   float* __restrict__ range, float* __restrict__ az, float* __restrict__ el) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Sph s = cart2sph(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    int r, c;
    pixel_index(s, g, &r, &c);
    row[i] = r; col[i] = c; range[i] = s.r; az[i] = s.az; el[i] = s.el;
}

// records an event pair around the dominant kernel; collected by prof_collect after the pass's final synchronisation
static inline void prof_begin(ltr_ctx* ctx) { if (ctx->ev_used + 2 <= (int)ctx->ev_pool.size()) cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream); }
static inline void prof_end(ltr_ctx* ctx) { if (ctx->ev_used + 2 <= (int)ctx->ev_pool.size()) { cudaEventRecord(ctx->ev_pool[ctx->ev_used + 1], ctx->stream); ctx->ev_used += 2; } }
static inline void prof_collect(ltr_ctx* ctx, int slot, double bytes_per_launch_unit, double proj_per_launch_unit, const std::vector<int>& units) {
    // units[i] = keyframes in launch i (only the launches that got an event pair are accumulated)
    for (int i = 0; i * 2 + 1 < ctx->ev_used && i < (int)units.size(); ++i) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[2 * i], ctx->ev_pool[2 * i + 1]) != cudaSuccess) continue;
        ctx->prof[slot + 0] += (double)ms * 1000.0;
        ctx->prof[slot + 1] += 1.0;
        ctx->prof[slot + 2] += bytes_per_launch_unit * units[i];
        ctx->prof[slot + 3] += proj_per_launch_unit * units[i];
    }
    ctx->ev_used = 0;
}

// 1 iff no map point can be >= 8000 m away from any keyframe origin in [k0, k1): then a pixel without a scan return
// (scan range 10000) has scan - map > 200 for every map point and can never flag (see project_fast.cuh).
static int empty_scan_shortcut_ok(ltr_ctx* ctx, const DevCloud& map, const DevPoses& poses, int k0, int k1, int* ok) {
    *ok = 0;
    float mn[3], mx[3];
    LTR_TRY(minmax_xyz(ctx, map, mn, mx));
    double worst = 0.0;
    for (int k = k0; k < k1; ++k) {
        const float* f = &poses.h_fast[(size_t)k * 16];
        if (f[15] == 0.0f) return LTR_OK;
        double d2 = 0.0;
        for (int d = 0; d < 3; ++d) {
            const double c = (double)f[9 + d];  // c_hi (the f32 part of the sensor origin is plenty for a 7 km bound)
            const double e = std::max(std::fabs(c - (double)mn[d]), std::fabs(c - (double)mx[d]));
            d2 += e * e;
        }
        worst = std::max(worst, d2);
    }
    *ok = (worst < 7000.0 * 7000.0) ? 1 : 0;
    return LTR_OK;
}

template <bool kElDirect>
__global__ void debug_fast_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ kf, const double* __restrict__ pose,
                                  const double* __restrict__ ext, int ext_identity, int order, ImgShape g, FastCfg fc,
                                  float* __restrict__ out /* n x 8: vcol_f vrow_f r_f rho_inv_r | vcol_e vrow_e r_e unused */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const FastProj f = fast_project<kElDirect>(kf, x, y, z);
    out[8 * i + 0] = __fmaf_rn(f.az, fc.col_scale, fc.col_off);
    out[8 * i + 1] = __fmaf_rn(f.el, fc.neg_row_scale, fc.row_off);
    out[8 * i + 2] = f.r;
    out[8 * i + 3] = f.rho_inv_r;
    float lx, ly, lz;
    transform_point(pose, order, x, y, z, &lx, &ly, &lz);
    if (!ext_identity) transform_point(ext, order, lx, ly, lz, &lx, &ly, &lz);
    const Sph s = cart2sph(lx, ly, lz);
    // the reference's pre-round floats (utility.cpp:122-123)
    out[8 * i + 4] = fm((float)g.cols, fd(fa(rad2deg(s.az), fd(g.hfov, 2.0f)), fs(g.hfov, 0.0f)));
    out[8 * i + 5] = fm((float)g.rows, fs(1.0f, fd(fa(rad2deg(s.el), fd(g.vfov, 2.0f)), fs(g.vfov, 0.0f))));
    out[8 * i + 6] = s.r;
    out[8 * i + 7] = 0.0f;
}

// Exhaustive check of the two arctangent polynomials of the fast path: every float a with bits in [0, last_bits] (i.e. all of [0, 1]
// or [0, 0.5]) is evaluated and compared with atan((double)a); the maximum absolute error is kept (positive doubles order like their bits).
__global__ void __launch_bounds__(256) atan_sweep_kernel(int which, uint32_t last_bits, unsigned long long* __restrict__ max_err_bits, uint32_t* __restrict__ arg_bits) {
    double worst = 0.0;
    uint32_t worst_arg = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= last_bits; b += (uint64_t)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((uint32_t)b);
        const float v = which == 0 ? fast_atan01(a) : fast_atan_half(a);
        const double e = fabs((double)v - atan((double)a));
        if (e > worst) { worst = e; worst_arg = (uint32_t)b; }
    }
    const unsigned long long wb = (unsigned long long)__double_as_longlong(worst);
    const unsigned long long old = atomicMax(max_err_bits, wb);
    if (wb > old) *arg_bits = worst_arg;   // racy between equal-ish maxima; the argument is informational only
}

// persistent grid: 4 resident CTAs per SM (64 registers, ~22 KB shared memory each), never more CTAs than tiles need
static inline unsigned fast_grid(const ltr_ctx* ctx, int64_t n) {
    const int64_t tiles = (n + 32 * kFastPts - 1) / (32 * kFastPts);
    const int64_t ctas = (tiles + kFastThreads / 32 - 1) / (kFastThreads / 32);
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)ctx->sm_count * kFastCtasPerSm));
}

static inline size_t fast_smem_bytes() { return (size_t)(kFastThreads / 32) * 2 * kQueueCap * sizeof(uint64_t); }

static inline unsigned grid_for(int64_t n, int threads, int max_blocks) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, max_blocks));
}

static inline KfBatch make_kf_batch(const DevPoses& poses, int k0, int nb) {
    KfBatch kb;
    std::memset(&kb, 0, sizeof(kb));
    std::memcpy(kb.kf, &poses.h_fast[(size_t)k0 * 16], (size_t)nb * 16 * sizeof(float));
    return kb;
}

// one launch of the fast projection over the keyframes [k0, k0 + nb) of `poses`
template <bool kCand, bool kDeferred = false>
static void launch_fast(ltr_ctx* ctx, const DevCloud& map, const DevPoses& poses, int k0, int nb, const ImgShape& g, const FastCfg& fc,
                        const uint32_t* rimg, const uint32_t* test_img, float thres, uint64_t* win, uint32_t* amin, const CullArgs& ca,
                        DeferredArgs da = DeferredArgs{nullptr, nullptr, nullptr, nullptr, 0}) {
    const unsigned fb = fast_grid(ctx, map.n);
    unsigned int* work = (unsigned int*)(ctx->d_counters + 4);
    cudaMemsetAsync(work, 0, sizeof(unsigned int), ctx->stream);
    const KfBatch kb = make_kf_batch(poses, k0, nb);
    if (kDeferred) {
        cudaMemsetAsync(da.best, 0xff, (size_t)nb * g.rows * g.cols * sizeof(unsigned long long), ctx->stream);   // "no pair yet"
        cudaMemsetAsync(da.count, 0, sizeof(unsigned int), ctx->stream);
    }
    if (fc.el_direct) map_project_fast_kernel<kCand, true, kDeferred><<<fb, kFastThreads, fast_smem_bytes(), ctx->stream>>>(view(map), kb, poses.d, k0, nb, ctx->d_ext,
        ctx->ext_identity ? 1 : 0, ctx->cfg.transform_order, g, fc, rimg, test_img, thres, win, amin, ca, ctx->d_counters, work, da);
    else map_project_fast_kernel<kCand, false, kDeferred><<<fb, kFastThreads, fast_smem_bytes(), ctx->stream>>>(view(map), kb, poses.d, k0, nb, ctx->d_ext,
        ctx->ext_identity ? 1 : 0, ctx->cfg.transform_order, g, fc, rimg, test_img, thres, win, amin, ca, ctx->d_counters, work, da);
    if (kDeferred) {
        const int64_t total = (int64_t)nb * g.rows * g.cols;
        ctx->launches++;
        deferred_resolve_kernel<<<grid_for(total, 256, ctx->sm_count * 16), 256, 0, ctx->stream>>>(view(map), poses.d, k0, nb, ctx->d_ext, ctx->ext_identity ? 1 : 0,
            ctx->cfg.transform_order, (uint32_t)(g.rows * g.cols), da, win);
    }
}

// scratch of the deferred true-minimum mode for launches of up to B keyframes: best image + overflow list + its counter
constexpr unsigned kDeferredListCap = 8u << 20;   // entries (64 MB); a typical launch appends a few 10^4
static int deferred_alloc(ltr_ctx* ctx, int B, int64_t npx, void** block, DeferredArgs* da) {
    const size_t best_bytes = (size_t)B * npx * sizeof(unsigned long long);
    LTR_TRY(dev_alloc(ctx, block, best_bytes + (size_t)kDeferredListCap * sizeof(unsigned long long) + 256));
    da->best = (unsigned long long*)*block;
    da->list = da->best + (size_t)B * npx;
    da->count = (unsigned int*)(da->list + kDeferredListCap);
    da->overflow = da->count + 1;
    da->capacity = kDeferredListCap;
    LTR_CUDA(ctx, cudaMemsetAsync(da->count, 0, 2 * sizeof(unsigned int), ctx->stream));
    return LTR_OK;
}


}  // namespace ltr

using namespace ltr;

extern "C" {

static int remove_pass_impl(ltr_ctx* ctx, ltr_cloud map_h, ltr_scanset scans_h, ltr_poses poses_h, int32_t kf_begin, int32_t kf_end,
                            int32_t mode, float res_alpha, float diff_thres, int32_t accumulate, int64_t* n_dynamic, bool allow_deferred, bool* overflowed);

int ltr_remove_pass(ltr_ctx* ctx, ltr_cloud map_h, ltr_scanset scans_h, ltr_poses poses_h, int32_t kf_begin, int32_t kf_end,
                    int32_t mode, float res_alpha, float diff_thres, int32_t accumulate, int64_t* n_dynamic) {
    ApiTrace tr__(ctx, "ltr_remove_pass");
    if (!ctx) return LTR_ERR_INVALID;
    bool overflowed = false;
    // ND (true per-pixel minimum) runs in the deferred mode of project_fast.cuh unless the flags accumulate into earlier ones (a run whose
    // overflow list filled up would have to be undone); a full list -- never seen, the list holds 8 M entries -- repeats the pass in the immediate mode
    LTR_TRY(remove_pass_impl(ctx, map_h, scans_h, poses_h, kf_begin, kf_end, mode, res_alpha, diff_thres, accumulate, n_dynamic, !accumulate, &overflowed));
    if (overflowed) LTR_TRY(remove_pass_impl(ctx, map_h, scans_h, poses_h, kf_begin, kf_end, mode, res_alpha, diff_thres, accumulate, n_dynamic, false, &overflowed));
    return LTR_OK;
}

static int remove_pass_impl(ltr_ctx* ctx, ltr_cloud map_h, ltr_scanset scans_h, ltr_poses poses_h, int32_t kf_begin, int32_t kf_end,
                            int32_t mode, float res_alpha, float diff_thres, int32_t accumulate, int64_t* n_dynamic, bool allow_deferred, bool* overflowed) {
    *overflowed = false;
    DevCloud* map;
    DevScanSet* scans;
    DevPoses* poses;
    LTR_TRY(cloud_get(ctx, map_h, &map));
    LTR_TRY(scanset_get(ctx, scans_h, &scans));
    LTR_TRY(poses_get(ctx, poses_h, &poses));
    if (mode != LTR_MODE_HD && mode != LTR_MODE_ND && mode != LTR_MODE_PD) return fail(ctx, LTR_ERR_INVALID, "unknown pass mode %d", mode);
    if (scans->K != poses->K) return fail(ctx, LTR_ERR_INVALID, "pose count %d != keyframe count %d (Session.cpp:117)", poses->K, scans->K);
    if (kf_begin < 0 || kf_end > scans->K || kf_begin > kf_end) return fail(ctx, LTR_ERR_INVALID, "keyframe range [%d,%d) outside [0,%d)", kf_begin, kf_end, scans->K);
    if (map->n >= ((int64_t)1 << 32)) return fail(ctx, LTR_ERR_UNSUPPORTED, "map larger than 2^32 points");
    int32_t rows, cols;
    ltr_reset_rimg_size(ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, res_alpha, &rows, &cols);
    if (rows < 1 || cols < 1) return fail(ctx, LTR_ERR_INVALID, "range image %dx%d is empty (res_alpha %g)", rows, cols, res_alpha);
    LTR_TRY(cloud_ensure_flags(ctx, map));
    if (!accumulate && map->n > 0) LTR_CUDA(ctx, cudaMemsetAsync(map->flags, 0, (size_t)map->n, ctx->stream));
    const ImgShape g{rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg};
    const int64_t npx = (int64_t)rows * cols;
    const bool cand = (mode != LTR_MODE_ND);
    const bool use_fast = ctx->cfg.fast_path && npx <= (1 << 18);   // queue-entry field widths, 32-bit pixel offsets (project_fast.cuh)
    // fast path: at most kFastMaxBatch keyframes per launch (their constants travel as a kernel parameter); exact path: 12 doubles of shared memory per keyframe
    const int B = std::max(1, std::min(std::min(ctx->cfg.keyframe_batch, use_fast ? kFastMaxBatch : 400), kf_end - kf_begin));
    int shortcut = 0;
    if (use_fast && cand && map->n > 0 && kf_end > kf_begin) LTR_TRY(empty_scan_shortcut_ok(ctx, *map, *poses, kf_begin, kf_end, &shortcut));
    const FastCfg fc = make_fast_cfg(rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, shortcut);
    if (use_fast) LTR_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    const bool use_deferred = use_fast && !cand && allow_deferred;
    void* p_def = nullptr;
    ScratchGuard g_def(ctx, &p_def);
    DeferredArgs da{nullptr, nullptr, nullptr, nullptr, 0};
    if (use_deferred && map->n > 0 && kf_end > kf_begin) LTR_TRY(deferred_alloc(ctx, B, npx, &p_def, &da));
    // tile culling (project_cull.cuh): scan-minus-map variants only, one keyframe per lane -> launches of at most 32 keyframes
    CullArgs ca;
    std::memset(&ca, 0, sizeof(ca));
    const bool use_cull = use_fast && cand && ctx->cfg.fast_path >= 2 && B <= 32 && map->n >= kTilePts;
    void *p_tiles = nullptr, *p_pyr = nullptr;
    ScratchGuard g_tiles(ctx, &p_tiles), g_pyr(ctx, &p_pyr);
    if (use_cull) {
        ca.enabled = 1;
        ca.ntiles = (map->n + kTilePts - 1) / kTilePts;
        unsigned off = 0;
        for (int L = 1; L <= kPyrLevels; ++L) {
            const int RL = (rows + (1 << L) - 1) >> L, CL = (cols + (1 << L) - 1) >> L;
            ca.lvl_off[L] = off; ca.lvl_cols[L] = CL;
            off += (unsigned)(RL * CL);
        }
        ca.pyr_stride = off;
        LTR_TRY(dev_alloc(ctx, &p_tiles, (size_t)ca.ntiles * sizeof(float4)));
        LTR_TRY(dev_alloc(ctx, &p_pyr, (size_t)B * ca.pyr_stride * sizeof(float)));
        ca.tiles = (const float4*)p_tiles; ca.pyr = (const float*)p_pyr;
        tile_sphere_kernel<<<(unsigned)((ca.ntiles * 32 + 255) / 256), 256, 0, ctx->stream>>>(view(*map), (float4*)p_tiles, ca.ntiles);
        LTR_LAUNCH_CHECK(ctx);
    }
    LTR_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    std::vector<int> launch_units;
    ctx->ev_used = 0;
    if (map->n > 0 && kf_end > kf_begin) {
        void *p_rimg = nullptr, *p_win = nullptr;
        ScratchGuard g_rimg(ctx, &p_rimg), g_win(ctx, &p_win);
        LTR_TRY(dev_alloc(ctx, &p_rimg, (size_t)B * npx * sizeof(uint32_t) * (use_fast ? 3 : 1)));
        LTR_TRY(dev_alloc(ctx, &p_win, (size_t)B * npx * sizeof(uint64_t)));
        uint32_t* rimg = (uint32_t*)p_rimg;
        uint32_t* amin = rimg + (size_t)B * npx;   // approximate running minimum (fast path only)
        uint32_t* timg = amin + (size_t)B * npx;   // test image of the scan-minus-map variants when pixels without a return can never flag
        uint64_t* win = (uint64_t*)p_win;
        const int fill_blocks = ctx->sm_count * 8;
        fill_u64_kernel<<<grid_for(B * npx, 256, fill_blocks), 256, 0, ctx->stream>>>(win, cand ? kWinNone : kWinEmpty, B * npx);
        LTR_LAUNCH_CHECK(ctx);
        for (int k0 = kf_begin; k0 < kf_end; k0 += B) {
            const int nb = std::min(B, kf_end - k0);
            fill_u32_kernel<<<grid_for(nb * npx, 256, fill_blocks), 256, 0, ctx->stream>>>(rimg, kNoPointBits, nb * npx);
            LTR_LAUNCH_CHECK(ctx);
            if (use_fast) {
                fill_u32_kernel<<<grid_for(nb * npx, 256, fill_blocks), 256, 0, ctx->stream>>>(amin, 0x7f800000u, nb * npx);
                LTR_LAUNCH_CHECK(ctx);
            }

            const int64_t npts = scans->h_off[k0 + nb] - scans->h_off[k0];
            if (npts > 0) {
                if (use_fast && fc.el_direct) scan_rimg_fast_kernel<true><<<(unsigned)((npts + 255) / 256), 256, 0, ctx->stream>>>(view(scans->pts), scans->d_off, k0, nb, g, fc, rimg);
                else if (use_fast) scan_rimg_fast_kernel<false><<<(unsigned)((npts + 255) / 256), 256, 0, ctx->stream>>>(view(scans->pts), scans->d_off, k0, nb, g, fc, rimg);
                else scan_rimg_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, ctx->stream>>>(view(scans->pts), scans->d_off, k0, nb, g, rimg);
                LTR_LAUNCH_CHECK(ctx);
            }
            if (use_cull) {
                const float empty_value = shortcut ? -__builtin_huge_valf() : __builtin_huge_valf();   // empty scan pixel: never flags / unknown
                scan_pyramid_kernel<<<(unsigned)nb, 1024, 0, ctx->stream>>>(rimg, nb, rows, cols, empty_value, ca, (float*)p_pyr);
                LTR_LAUNCH_CHECK(ctx);
            }
            const uint32_t* test_img = rimg;
            if (use_fast && cand && shortcut) {
                scan_test_image_kernel<<<grid_for(nb * npx, 256, fill_blocks), 256, 0, ctx->stream>>>(rimg, timg, nb * npx);
                LTR_LAUNCH_CHECK(ctx);
                test_img = timg;
            }
            prof_begin(ctx);
            launch_units.push_back(nb);
            if (use_fast) {
                if (cand) launch_fast<true>(ctx, *map, *poses, k0, nb, g, fc, rimg, test_img, diff_thres, win, amin, ca);
                else if (use_deferred) launch_fast<false, true>(ctx, *map, *poses, k0, nb, g, fc, rimg, amin, diff_thres, win, amin, ca, da);
                else launch_fast<false>(ctx, *map, *poses, k0, nb, g, fc, rimg, amin, diff_thres, win, amin, ca);
            } else {
                const unsigned mb = (unsigned)((map->n + 255) / 256);
                const size_t smem = (size_t)nb * 12 * sizeof(double);
                if (cand) map_project_kernel<true><<<mb, 256, smem, ctx->stream>>>(view(*map), poses->d, k0, nb, ctx->d_ext, ctx->ext_identity ? 1 : 0,
                                                                                 ctx->cfg.transform_order, g, rimg, diff_thres, win);
                else map_project_kernel<false><<<mb, 256, smem, ctx->stream>>>(view(*map), poses->d, k0, nb, ctx->d_ext, ctx->ext_identity ? 1 : 0,
                                                                              ctx->cfg.transform_order, g, rimg, diff_thres, win);
            }
            prof_end(ctx);
            LTR_LAUNCH_CHECK(ctx);
            const unsigned rb = (unsigned)((nb * npx + 255) / 256);
            if (cand) resolve_kernel<true><<<rb, 256, 0, ctx->stream>>>(rimg, win, nb * npx, diff_thres, map->flags);
            else resolve_kernel<false><<<rb, 256, 0, ctx->stream>>>(rimg, win, nb * npx, diff_thres, map->flags);
            LTR_LAUNCH_CHECK(ctx);
        }
    }
    g_tiles.release();
    g_pyr.release();
    LTR_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    if (n_dynamic) LTR_TRY(count_flags(ctx, map->flags, map->n, n_dynamic));
    unsigned int h_over = 0;
    if (da.overflow) LTR_CUDA(ctx, cudaMemcpyAsync(&h_over, da.overflow, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    LTR_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    if (da.overflow) { LTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); *overflowed = h_over != 0; }
    g_def.release();
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    prof_collect(ctx, 0, 12.0 * (double)map->n + (double)map->n / 8.0, (double)map->n, launch_units);
    ctx->stats[0] = (double)map->n * (kf_end - kf_begin);
    ctx->stats[1] = 0; ctx->stats[2] = ctx->stats[0]; ctx->stats[3] = 0;
    ctx->stats_counters_pending = use_fast;   // device counters are fetched lazily by ltr_last_pass_stats
    ctx->stats[4] = (double)ms * 1000.0;
    return LTR_OK;
}

static int parse_projected_impl(ltr_ctx* ctx, ltr_cloud map_h, ltr_poses poses_h, int32_t kf_begin, int32_t kf_end, float res_alpha, ltr_scanset* out,
                                bool allow_deferred, bool* overflowed);

int ltr_parse_projected(ltr_ctx* ctx, ltr_cloud map_h, ltr_poses poses_h, int32_t kf_begin, int32_t kf_end, float res_alpha, ltr_scanset* out) {
    ApiTrace tr__(ctx, "ltr_parse_projected");
    if (!ctx || !out) return fail(ctx, LTR_ERR_INVALID, "null argument");
    bool overflowed = false;
    LTR_TRY(parse_projected_impl(ctx, map_h, poses_h, kf_begin, kf_end, res_alpha, out, true, &overflowed));
    if (overflowed) {   // the deferred mode's candidate list filled up (8 M entries; never seen): repeat with the immediate mode
        LTR_TRY(ltr_scanset_free(ctx, *out));
        LTR_TRY(parse_projected_impl(ctx, map_h, poses_h, kf_begin, kf_end, res_alpha, out, false, &overflowed));
    }
    return LTR_OK;
}

static int parse_projected_impl(ltr_ctx* ctx, ltr_cloud map_h, ltr_poses poses_h, int32_t kf_begin, int32_t kf_end, float res_alpha, ltr_scanset* out,
                                bool allow_deferred, bool* overflowed) {
    *overflowed = false;
    DevCloud* map;
    DevPoses* poses;
    LTR_TRY(cloud_get(ctx, map_h, &map));
    LTR_TRY(poses_get(ctx, poses_h, &poses));
    if (kf_begin < 0 || kf_end > poses->K || kf_begin > kf_end) return fail(ctx, LTR_ERR_INVALID, "keyframe range [%d,%d) outside [0,%d)", kf_begin, kf_end, poses->K);
    if (map->n >= ((int64_t)1 << 32)) return fail(ctx, LTR_ERR_UNSUPPORTED, "map larger than 2^32 points");
    int32_t rows, cols;
    ltr_reset_rimg_size(ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, res_alpha, &rows, &cols);
    if (rows < 1 || cols < 1) return fail(ctx, LTR_ERR_INVALID, "range image %dx%d is empty", rows, cols);
    const ImgShape g{rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg};
    const int64_t npx = (int64_t)rows * cols;
    const int K = kf_end - kf_begin;
    const DevCloud mapc = *map;
    const DevPoses posc = *poses;
    std::vector<int64_t> off((size_t)K + 1, 0);
    const FastCfg fc = make_fast_cfg(rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, 0);
    const bool use_fast = ctx->cfg.fast_path && npx <= (1 << 18);
    if (use_fast) LTR_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    LTR_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    void *p_list = nullptr, *p_cnt = nullptr, *p_def = nullptr;
    ScratchGuard g_list(ctx, &p_list), g_cnt(ctx, &p_cnt), g_def(ctx, &p_def);
    DeferredArgs da{nullptr, nullptr, nullptr, nullptr, 0};
    const bool use_deferred = use_fast && allow_deferred;
    std::vector<int> launch_units;
    ctx->ev_used = 0;
    if (K > 0 && mapc.n > 0) {
        const int B = std::max(1, std::min(std::min(ctx->cfg.keyframe_batch, use_fast ? kFastMaxBatch : 400), K));
        if (use_deferred) LTR_TRY(deferred_alloc(ctx, B, npx, &p_def, &da));
        void *p_win = nullptr, *p_amin = nullptr;
        ScratchGuard g_win(ctx, &p_win), g_amin(ctx, &p_amin);
        LTR_TRY(dev_alloc(ctx, &p_win, (size_t)B * npx * sizeof(uint64_t)));
        if (use_fast) LTR_TRY(dev_alloc(ctx, &p_amin, (size_t)B * npx * sizeof(uint32_t)));
        uint32_t* amin = (uint32_t*)p_amin;
        LTR_TRY(dev_alloc(ctx, &p_list, (size_t)K * npx * sizeof(uint32_t)));
        const int nchunk = (int)((npx + kParseChunk - 1) / kParseChunk);
        LTR_TRY(dev_alloc(ctx, &p_cnt, ((size_t)K + (size_t)B * nchunk) * sizeof(unsigned int)));
        unsigned int* chunk_cnt = (unsigned int*)p_cnt + K;
        uint64_t* win = (uint64_t*)p_win;
        fill_u64_kernel<<<grid_for(B * npx, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(win, kWinEmpty, B * npx);
        LTR_LAUNCH_CHECK(ctx);
        for (int k0 = 0; k0 < K; k0 += B) {
            const int nb = std::min(B, K - k0);
            prof_begin(ctx);
            launch_units.push_back(nb);
            if (use_fast) {
                fill_u32_kernel<<<grid_for(nb * npx, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(amin, 0x7f800000u, nb * npx);
                LTR_LAUNCH_CHECK(ctx);
                CullArgs no_cull;
                std::memset(&no_cull, 0, sizeof(no_cull));
                if (use_deferred) launch_fast<false, true>(ctx, mapc, posc, kf_begin + k0, nb, g, fc, nullptr, amin, 0.0f, win, amin, no_cull, da);
                else launch_fast<false>(ctx, mapc, posc, kf_begin + k0, nb, g, fc, nullptr, amin, 0.0f, win, amin, no_cull);
            } else {
                const unsigned mb = (unsigned)((mapc.n + 255) / 256);
                map_project_kernel<false><<<mb, 256, (size_t)nb * 12 * sizeof(double), ctx->stream>>>(view(mapc), posc.d, kf_begin + k0, nb, ctx->d_ext,
                    ctx->ext_identity ? 1 : 0, ctx->cfg.transform_order, g, nullptr, 0.0f, win);
            }
            prof_end(ctx);
            LTR_LAUNCH_CHECK(ctx);
            parse_count_kernel<<<dim3((unsigned)nchunk, (unsigned)nb), 1024, 0, ctx->stream>>>(win, (int)npx, nchunk, chunk_cnt);
            LTR_LAUNCH_CHECK(ctx);
            parse_compact_kernel<<<dim3((unsigned)nchunk, (unsigned)nb), 1024, 0, ctx->stream>>>(win, (int)npx, nchunk, chunk_cnt, (uint32_t*)p_list + (size_t)k0 * npx,
                                                                                              (unsigned int*)p_cnt + k0);
            LTR_LAUNCH_CHECK(ctx);
        }
        g_win.release();
        g_amin.release();
        std::vector<unsigned int> cnt((size_t)K);
        unsigned int h_over = 0;
        LTR_CUDA(ctx, cudaMemcpyAsync(cnt.data(), p_cnt, (size_t)K * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        if (da.overflow) LTR_CUDA(ctx, cudaMemcpyAsync(&h_over, da.overflow, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        LTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        g_def.release();
        if (h_over) *overflowed = true;
        for (int k = 0; k < K; ++k) off[k + 1] = off[k] + cnt[k];
    }
    LTR_TRY(scanset_new(ctx, off, out));
    DevScanSet& os = ctx->scansets[*out];
    if (os.pts.n > 0) {
        parse_emit_kernel<<<(unsigned)((os.pts.n + 255) / 256), 256, 0, ctx->stream>>>(view(mapc), posc.d, kf_begin, os.d_off, K, (const uint32_t*)p_list,
            (int)npx, ctx->d_ext, ctx->ext_identity ? 1 : 0, ctx->cfg.transform_order, os.pts);
        LTR_LAUNCH_CHECK(ctx);
    }
    g_list.release();
    g_cnt.release();
    LTR_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    LTR_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    prof_collect(ctx, 4, 12.0 * (double)mapc.n, (double)mapc.n, launch_units);
    ctx->stats[0] = (double)mapc.n * K;
    ctx->stats[1] = 0; ctx->stats[2] = ctx->stats[0]; ctx->stats[3] = 0;
    ctx->stats_counters_pending = use_fast;
    ctx->stats[4] = (double)ms * 1000.0;
    return LTR_OK;
}

int ltr_debug_pixel_index(ltr_ctx* ctx, const float* xyz, int64_t n, int32_t rows, int32_t cols, int32_t* row, int32_t* col, float* range,
                          float* az, float* el) {
    ApiTrace tr__(ctx, "ltr_debug_pixel_index");
    if (!ctx || !xyz || n < 0) return fail(ctx, LTR_ERR_INVALID, "bad argument");
    if (n == 0) return LTR_OK;
    void* p = nullptr;
    ScratchGuard g_p(ctx, &p);
    const size_t bytes = (size_t)n * (3 + 5) * 4;
    LTR_TRY(dev_alloc(ctx, &p, bytes));
    float* d_xyz = (float*)p;
    int* d_row = (int*)(d_xyz + 3 * n);
    int* d_col = d_row + n;
    float* d_rng = (float*)(d_col + n);
    float* d_az = d_rng + n;
    float* d_el = d_az + n;
    LTR_CUDA(ctx, cudaMemcpyAsync(d_xyz, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    const ImgShape g{rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg};
    debug_pixel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_xyz, n, g, d_row, d_col, d_rng, d_az, d_el);
    LTR_LAUNCH_CHECK(ctx);
    if (row) LTR_CUDA(ctx, cudaMemcpyAsync(row, d_row, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (col) LTR_CUDA(ctx, cudaMemcpyAsync(col, d_col, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (range) LTR_CUDA(ctx, cudaMemcpyAsync(range, d_rng, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (az) LTR_CUDA(ctx, cudaMemcpyAsync(az, d_az, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (el) LTR_CUDA(ctx, cudaMemcpyAsync(el, d_el, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    LTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LTR_OK;
}

int ltr_debug_margins(ltr_ctx* ctx, float res_alpha, float* m) {
    if (!ctx || !m) return fail(ctx, LTR_ERR_INVALID, "bad argument");
    int32_t rows, cols;
    ltr_reset_rimg_size(ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, res_alpha, &rows, &cols);
    const FastCfg fc = make_fast_cfg(rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, 0);
    m[0] = fc.m_col_a; m[1] = fc.m_col_b; m[2] = fc.m_row; m[3] = fc.m_r_rel; m[4] = fc.m_r_abs; m[5] = (float)fc.el_direct;
    m[6] = (float)((fc.m_col_a + fc.m_col_b) / kMarginSafety); m[7] = (float)(fc.m_row / kMarginSafety);
    return LTR_OK;
}

int ltr_debug_atan_sweep(ltr_ctx* ctx, int32_t which, double* max_abs_err, float* arg_at_max) {
    ApiTrace tr__(ctx, "ltr_debug_atan_sweep");
    if (!ctx || !max_abs_err || (which != 0 && which != 1)) return fail(ctx, LTR_ERR_INVALID, "bad argument");
    void* p = nullptr;
    ScratchGuard g_p(ctx, &p);
    LTR_TRY(dev_alloc(ctx, &p, 16));
    LTR_CUDA(ctx, cudaMemsetAsync(p, 0, 16, ctx->stream));
    const uint32_t last = which == 0 ? 0x3F800000u : 0x3F000000u;   // 1.0f / 0.5f
    atan_sweep_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(which, last, (unsigned long long*)p, (uint32_t*)((char*)p + 8));
    LTR_LAUNCH_CHECK(ctx);
    unsigned long long h[2] = {0, 0};
    LTR_CUDA(ctx, cudaMemcpyAsync(h, p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    LTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(max_abs_err, &h[0], 8);
    if (arg_at_max) { const uint32_t b = (uint32_t)h[1]; std::memcpy(arg_at_max, &b, 4); }
    return LTR_OK;
}

int ltr_debug_scan_rimg(ltr_ctx* ctx, ltr_scanset scans_h, int32_t kf, float res_alpha, float* out) {
    ApiTrace tr__(ctx, "ltr_debug_scan_rimg");
    if (!ctx || !out) return fail(ctx, LTR_ERR_INVALID, "null argument");
    DevScanSet* scans;
    LTR_TRY(scanset_get(ctx, scans_h, &scans));
    if (kf < 0 || kf >= scans->K) return fail(ctx, LTR_ERR_INVALID, "keyframe %d outside [0,%d)", kf, scans->K);
    int32_t rows, cols;
    ltr_reset_rimg_size(ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, res_alpha, &rows, &cols);
    if (rows < 1 || cols < 1) return fail(ctx, LTR_ERR_INVALID, "range image %dx%d is empty", rows, cols);
    const ImgShape g{rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg};
    const int64_t npx = (int64_t)rows * cols;
    void* p = nullptr;
    ScratchGuard g_p(ctx, &p);
    LTR_TRY(dev_alloc(ctx, &p, (size_t)npx * sizeof(uint32_t)));
    fill_u32_kernel<<<grid_for(npx, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>((uint32_t*)p, kNoPointBits, npx);
    LTR_LAUNCH_CHECK(ctx);
    const int64_t npts = scans->h_off[kf + 1] - scans->h_off[kf];
    if (npts > 0) {
        const FastCfg fcd = make_fast_cfg(rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, 0);
        if (ctx->cfg.fast_path && npx <= (1 << 18) && fcd.el_direct)
            scan_rimg_fast_kernel<true><<<(unsigned)((npts + 255) / 256), 256, 0, ctx->stream>>>(view(scans->pts), scans->d_off, kf, 1, g, fcd, (uint32_t*)p);
        else if (ctx->cfg.fast_path && npx <= (1 << 18))
            scan_rimg_fast_kernel<false><<<(unsigned)((npts + 255) / 256), 256, 0, ctx->stream>>>(view(scans->pts), scans->d_off, kf, 1, g, fcd, (uint32_t*)p);
        else scan_rimg_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, ctx->stream>>>(view(scans->pts), scans->d_off, kf, 1, g, (uint32_t*)p);
        LTR_LAUNCH_CHECK(ctx);
    }
    LTR_CUDA(ctx, cudaMemcpyAsync(out, p, (size_t)npx * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LTR_OK;
}

int ltr_debug_fast_project(ltr_ctx* ctx, const float* xyz, int64_t n, const double* inv_pose16, float res_alpha, float* out8, float* margins4) {
    ApiTrace tr__(ctx, "ltr_debug_fast_project");
    if (!ctx || !xyz || !inv_pose16 || !out8 || n < 0) return fail(ctx, LTR_ERR_INVALID, "bad argument");
    int32_t rows, cols;
    ltr_reset_rimg_size(ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, res_alpha, &rows, &cols);
    const ImgShape g{rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg};
    const FastCfg fc = make_fast_cfg(rows, cols, ctx->cfg.vfov_deg, ctx->cfg.hfov_deg, 0);
    if (margins4) { margins4[0] = fc.m_col_a; margins4[1] = fc.m_col_b; margins4[2] = fc.m_row; margins4[3] = fc.m_r_rel; }
    if (n == 0) return LTR_OK;
    ltr_poses ph;
    double id[16];
    for (int i = 0; i < 16; ++i) id[i] = (i % 5 == 0) ? 1.0 : 0.0;
    LTR_TRY(ltr_poses_upload(ctx, id, inv_pose16, 1, &ph));
    const DevPoses pp = ctx->poses[ph];
    void* p = nullptr;
    ScratchGuard g_p(ctx, &p);
    LTR_TRY(dev_alloc(ctx, &p, (size_t)n * (3 + 8) * sizeof(float)));
    float* d_xyz = (float*)p;
    float* d_out = d_xyz + 3 * n;
    LTR_CUDA(ctx, cudaMemcpyAsync(d_xyz, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    if (fc.el_direct) debug_fast_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_xyz, n, pp.d_fast, pp.d, ctx->d_ext, ctx->ext_identity ? 1 : 0,
                                                                        ctx->cfg.transform_order, g, fc, d_out);
    else debug_fast_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_xyz, n, pp.d_fast, pp.d, ctx->d_ext, ctx->ext_identity ? 1 : 0,
                                                                        ctx->cfg.transform_order, g, fc, d_out);
    LTR_LAUNCH_CHECK(ctx);
    LTR_CUDA(ctx, cudaMemcpyAsync(out8, d_out, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g_p.release();
    return ltr_poses_free(ctx, ph);
}

}  // extern "C"
