// libltr_b200.so -- exact k-nearest-neighbour labelling on a dense uniform grid.
//
// Reference functions replaced:
//   Session::extractLowDynPointsViaKnnDiff / extractHighDynPointsViaKnnDiff   ltremovert/src/Session.cpp:393-427, 487-504
//   Session::partition{Low,High}DynamicPointsOfScanByKnn                      ltremovert/src/Session.cpp:537-607, 610-642
//   Session::removeWeakNDMapPointsHavingStrongNDInNear                        ltremovert/src/Session.cpp:452-484
// The reference queries pcl::KdTreeFLANN (FLANN KDTreeSingleIndex, L2_Simple<float>, exact) and tests
//   |(float)(sum_double of the k smallest SQUARED distances) / float(k)| < thr.
// Exactness of the grid search: every squared distance is >= 0, so a point can only be labelled "near" if all of its
// k nearest squared distances are < k*thr.  With cells of side >= sqrt(k*thr)*(1+slack) every such neighbour lies in the
// 27 cells around the query's cell; if one of the true k nearest lies outside them, its squared distance alone already
// exceeds k*thr and both searches label the point "far".  Squared distances are evaluated exactly as FLANN's L2_Simple:
// ((dx*dx) + dy*dy) + dz*dz in f32, dx = query - point, no FMA.
#include "ltr_internal.cuh"
#include "ref_math.cuh"
#include <algorithm>
#include <cmath>

namespace ltr {

constexpr int kMaxK = 16;

struct Grid {
    double origin[3];
    double inv_cell;
    int dim[3];
};

__device__ __forceinline__ void cell_of(const Grid& g, float x, float y, float z, int* cx, int* cy, int* cz) {
    *cx = (int)floor(((double)x - g.origin[0]) * g.inv_cell);
    *cy = (int)floor(((double)y - g.origin[1]) * g.inv_cell);
    *cz = (int)floor(((double)z - g.origin[2]) * g.inv_cell);
}

__global__ void __launch_bounds__(256) grid_count_kernel(PtrView t, Grid g, uint32_t* __restrict__ cell_cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.n) return;
    int cx, cy, cz;
    cell_of(g, t.x[i], t.y[i], t.z[i], &cx, &cy, &cz);
    cx = min(max(cx, 0), g.dim[0] - 1); cy = min(max(cy, 0), g.dim[1] - 1); cz = min(max(cz, 0), g.dim[2] - 1);
    atomicAdd(&cell_cnt[((size_t)cz * g.dim[1] + cy) * g.dim[0] + cx], 1u);
}

__global__ void __launch_bounds__(256) grid_fill_kernel(PtrView t, Grid g, const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cursor,
                                                        float4* __restrict__ sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.n) return;
    const float x = t.x[i], y = t.y[i], z = t.z[i];
    int cx, cy, cz;
    cell_of(g, x, y, z, &cx, &cy, &cz);
    cx = min(max(cx, 0), g.dim[0] - 1); cy = min(max(cy, 0), g.dim[1] - 1); cz = min(max(cz, 0), g.dim[2] - 1);
    const size_t c = ((size_t)cz * g.dim[1] + cy) * g.dim[0] + cx;
    const uint32_t slot = cell_start[c] + atomicAdd(&cursor[c], 1u);
    sorted[slot] = make_float4(x, y, z, 0.0f);
}

// k smallest squared distances of (qx,qy,qz) among the 27 neighbouring cells -> decision "far" (1) or "near" (0).
// KT >= k is the compile-time capacity of the register-resident ascending top list.
template <int KT>
__device__ __forceinline__ uint8_t knn_label(const Grid& g, const uint32_t* __restrict__ cell_start, const float4* __restrict__ sorted,
                                             float qx, float qy, float qz, int k, float thr) {
    const float kInf = __int_as_float(0x7f800000);
    float best[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) best[j] = kInf;
    int cx, cy, cz;
    cell_of(g, qx, qy, qz, &cx, &cy, &cz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
    if (x0 <= x1) {
        // rows of 3 x-adjacent cells, the query's own row first: most points have their k neighbours right there and
        // leave through the early exit below
        // (dy, dz) + 1 packed 2 bits per step: dy = {0,-1,1,0,0,-1,1,-1,1}, dz = {0,0,0,-1,1,-1,-1,1,1}
        const unsigned pack_dy = 0x22161U, pack_dz = 0x28215U;
#pragma unroll 1
        for (int o = 0; o < 9; ++o) {
            const int z = cz + (int)((pack_dz >> (2 * o)) & 3u) - 1, y = cy + (int)((pack_dy >> (2 * o)) & 3u) - 1;
            if (z < 0 || z >= g.dim[2] || y < 0 || y >= g.dim[1]) continue;
            const size_t row = ((size_t)z * g.dim[1] + y) * g.dim[0];
            const uint32_t s = cell_start[row + x0], e = cell_start[row + x1 + 1];
            for (uint32_t p = s; p < e; ++p) {
                const float4 t = sorted[p];
                const float d = l2_simple(qx, qy, qz, t.x, t.y, t.z);
                if (d < best[KT - 1]) {
                    best[KT - 1] = d;
#pragma unroll
                    for (int j = KT - 1; j > 0; --j)
                        if (best[j] < best[j - 1]) { const float t2 = best[j]; best[j] = best[j - 1]; best[j - 1] = t2; }
                    // Early exit: the final k smallest distances are element-wise <= the current ones, and double sum, float
                    // rounding and division are monotone, so "current mean < thr" already decides "near".
                    if (KT <= 4 && best[k - 1 < KT ? k - 1 : KT - 1] != kInf) {
                        double sum = 0.0;
#pragma unroll
                        for (int j = 0; j < KT; ++j) if (j < k) sum = da(sum, (double)best[j]);
                        if (fabsf(fd(__double2float_rn(sum), (float)k)) < thr) return 0;
                    }
                }
            }
        }
    }
    // Fewer than k neighbours inside the 27 cells: a true k-th neighbour lies farther than the cell size, i.e. beyond
    // sqrt(k*thr) -> "far" (see file header).  The API rejects targets with fewer than k points up front.
    // accumulate(..., 0.0) in double (Session.cpp:593), float(sum) / float(k) (Session.cpp:594).
    double sum = 0.0;
    bool full = true;
#pragma unroll
    for (int j = 0; j < KT; ++j)
        if (j < k) { if (best[j] == kInf) full = false; else sum = da(sum, (double)best[j]); }
    if (!full) return 1;
    const float avg = fd(__double2float_rn(sum), (float)k);
    return (fabsf(avg) < thr) ? 0 : 1;
}

// Target with fewer than k points (1 <= n < k <= kMaxK): PCL's nearestKSearch returns all n of them (ascending), the reference still
// divides their sum by k (Session.cpp:470-471, 592-594).  Brute force over the n points held in shared memory.
__device__ __forceinline__ uint8_t knn_label_small(const float4* __restrict__ s_t, int n, float qx, float qy, float qz, int k, float thr) {
    float d[kMaxK];
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) d[j] = __int_as_float(0x7f800000);
    for (int p = 0; p < n; ++p) {
        float v = l2_simple(qx, qy, qz, s_t[p].x, s_t[p].y, s_t[p].z);
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) if (v < d[j]) { const float t = d[j]; d[j] = v; v = t; }   // ascending insertion
    }
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) if (j < n) sum = da(sum, (double)d[j]);
    return (fabsf(fd(__double2float_rn(sum), (float)k)) < thr) ? 0 : 1;
}

__device__ __forceinline__ uint8_t knn_label_dispatch(const Grid& g, const uint32_t* __restrict__ cell_start, const float4* __restrict__ sorted,
                                                      float qx, float qy, float qz, int k, float thr) {
    switch (k) {
        case 1: return knn_label<1>(g, cell_start, sorted, qx, qy, qz, k, thr);
        case 2: return knn_label<2>(g, cell_start, sorted, qx, qy, qz, k, thr);
        case 3: return knn_label<3>(g, cell_start, sorted, qx, qy, qz, k, thr);
        case 4: return knn_label<4>(g, cell_start, sorted, qx, qy, qz, k, thr);
        default: return knn_label<kMaxK>(g, cell_start, sorted, qx, qy, qz, k, thr);
    }
}

// scans: local -> (base2lidar as written at Session.cpp:545) -> pose => global query; label
__global__ void __launch_bounds__(128) knn_scan_label_kernel(PtrView scans, const int64_t* __restrict__ off, int K, const double* __restrict__ poses,
                                                             int pose_offset, const double* __restrict__ ext, int ext_identity, int order, Grid g,
                                                             const uint32_t* __restrict__ cell_start, const float4* __restrict__ sorted, int k,
                                                             float thr, uint8_t* __restrict__ label, float* __restrict__ gx, float* __restrict__ gy,
                                                             float* __restrict__ gz, PtrView small_target) {
    __shared__ float4 s_t[kMaxK];
    if (small_target.n > 0 && threadIdx.x < small_target.n) s_t[threadIdx.x] = make_float4(small_target.x[threadIdx.x], small_target.y[threadIdx.x], small_target.z[threadIdx.x], 0.0f);
    if (small_target.n > 0) __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= scans.n) return;
    int lo = 0, hi = K;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
    float x = scans.x[i], y = scans.y[i], z = scans.z[i];
    if (!ext_identity) transform_point(ext, order, x, y, z, &x, &y, &z);                              // utility.cpp:164 with base2lidar (as written)
    transform_point(poses + (size_t)(pose_offset + lo) * 24 + 12, order, x, y, z, &x, &y, &z);        // utility.cpp:165
    gx[i] = x; gy[i] = y; gz[i] = z;
    label[i] = small_target.n > 0 ? knn_label_small(s_t, (int)small_target.n, x, y, z, k, thr) : knn_label_dispatch(g, cell_start, sorted, x, y, z, k, thr);
}

// global2local of the partitioned points (Session.cpp:603-604)
__global__ void __launch_bounds__(256) knn_relocalise_kernel(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                                                             int64_t n, const int64_t* __restrict__ off, int K, const double* __restrict__ poses,
                                                             int pose_offset, const double* __restrict__ ext, int ext_identity, int order,
                                                             float* __restrict__ ox, float* __restrict__ oy, float* __restrict__ oz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = K;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
    float x = gx[i], y = gy[i], z = gz[i];
    transform_point(poses + (size_t)(pose_offset + lo) * 24, order, x, y, z, &x, &y, &z);   // utility.cpp:198
    if (!ext_identity) transform_point(ext, order, x, y, z, &x, &y, &z);                    // utility.cpp:199
    ox[i] = x; oy[i] = y; oz[i] = z;
}

__global__ void __launch_bounds__(128) knn_cloud_label_kernel(PtrView q, Grid g, const uint32_t* __restrict__ cell_start,
                                                              const float4* __restrict__ sorted, int k, float thr, uint8_t* __restrict__ label,
                                                              PtrView small_target) {
    __shared__ float4 s_t[kMaxK];
    if (small_target.n > 0 && threadIdx.x < small_target.n) s_t[threadIdx.x] = make_float4(small_target.x[threadIdx.x], small_target.y[threadIdx.x], small_target.z[threadIdx.x], 0.0f);
    if (small_target.n > 0) __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.n) return;
    label[i] = small_target.n > 0 ? knn_label_small(s_t, (int)small_target.n, q.x[i], q.y[i], q.z[i], k, thr)
                                  : knn_label_dispatch(g, cell_start, sorted, q.x[i], q.y[i], q.z[i], k, thr);
}

struct GridBuf {
    Grid g;
    uint32_t* cell_start = nullptr;  // ncell + 1
    float4* sorted = nullptr;
    void* block = nullptr;
};

static int build_grid(ltr_ctx* ctx, const DevCloud& target, int k, float thr, GridBuf* gb) {
    float mn[3], mx[3];
    LTR_TRY(minmax_xyz(ctx, target, mn, mx));
    const double reach = std::sqrt((double)k * (double)thr) * 1.001 + 1e-6;
    double cell = reach;
    // bound the dense grid to 2^29 cells by growing the cell (exactness only needs cell >= reach)
    for (;;) {
        double cells = 1.0;
        for (int d = 0; d < 3; ++d) cells *= std::floor(((double)mx[d] - (double)mn[d]) / cell) + 3.0;
        if (cells <= 536870912.0) break;
        cell *= 1.25;
    }
    Grid g;
    g.inv_cell = 1.0 / cell;
    size_t ncell = 1;
    for (int d = 0; d < 3; ++d) {
        g.origin[d] = (double)mn[d] - cell;  // one guard cell below the minimum
        g.dim[d] = (int)std::floor(((double)mx[d] - g.origin[d]) * g.inv_cell) + 2;
        ncell *= (size_t)g.dim[d];
    }
    void* p = nullptr;
    ScratchGuard g_p(ctx, &p);   // handed to the caller (gb->block) only when the grid is complete
    const size_t cs_bytes = (ncell + 1) * sizeof(uint32_t);
    const size_t cs_pad = (cs_bytes + 255) / 256 * 256;
    LTR_TRY(dev_alloc(ctx, &p, 2 * cs_pad + (size_t)target.n * sizeof(float4)));
    uint32_t* cnt = (uint32_t*)p;
    uint32_t* start = (uint32_t*)((char*)p + cs_pad);
    float4* sorted = (float4*)((char*)p + 2 * cs_pad);
    LTR_CUDA(ctx, cudaMemsetAsync(cnt, 0, cs_pad, ctx->stream));
    const unsigned nb = (unsigned)((target.n + 255) / 256);
    grid_count_kernel<<<nb, 256, 0, ctx->stream>>>(view(target), g, cnt);
    LTR_LAUNCH_CHECK(ctx);
    LTR_TRY(exclusive_scan_u32(ctx, cnt, start, (int64_t)ncell + 1));
    LTR_CUDA(ctx, cudaMemsetAsync(cnt, 0, cs_pad, ctx->stream));
    grid_fill_kernel<<<nb, 256, 0, ctx->stream>>>(view(target), g, start, cnt, sorted);
    LTR_LAUNCH_CHECK(ctx);
    gb->g = g; gb->cell_start = start; gb->sorted = sorted; gb->block = p;
    p = nullptr;   // ownership moves to gb
    return LTR_OK;
}

}  // namespace ltr

using namespace ltr;

extern "C" {

int ltr_knn_diff(ltr_ctx* ctx, ltr_scanset scans_h, ltr_poses poses_h, int32_t pose_offset, ltr_cloud target_h, int32_t k, float thr,
                 ltr_scanset* out_coexist, ltr_scanset* out_diff) {
    ApiTrace tr__(ctx, "ltr_knn_diff");
    if (!ctx) return LTR_ERR_INVALID;
    DevScanSet* sp;
    DevPoses* pp;
    DevCloud* tp;
    LTR_TRY(scanset_get(ctx, scans_h, &sp));
    LTR_TRY(poses_get(ctx, poses_h, &pp));
    LTR_TRY(cloud_get(ctx, target_h, &tp));
    if (k < 1 || k > kMaxK) return fail(ctx, LTR_ERR_UNSUPPORTED, "k = %d outside [1, %d]", k, kMaxK);
    if (!(thr > 0.0f)) return fail(ctx, LTR_ERR_INVALID, "threshold must be positive");
    if (pose_offset < 0 || pose_offset + sp->K > pp->K) return fail(ctx, LTR_ERR_INVALID, "pose range [%d,%d) outside [0,%d)", pose_offset, pose_offset + sp->K, pp->K);
    if (tp->n < 1) return fail(ctx, LTR_ERR_UNSUPPORTED, "empty target map (PCL's nearestKSearch asserts on an empty tree)");
    const bool small = tp->n < k;   // fewer than k target points: all of them are returned and the sum is still divided by k (Session.cpp:592-594)
    const DevScanSet scans = *sp;
    const DevPoses poses = *pp;
    const DevCloud target = *tp;
    const int64_t n = scans.pts.n;
    GridBuf gb;
    ScratchGuard g_grid(ctx, &gb.block);
    void* p = nullptr;
    ScratchGuard g_p(ctx, &p);
    uint8_t* label = nullptr;
    float *gx = nullptr, *gy = nullptr, *gz = nullptr;
    if (n > 0) {
        PtrView small_view{nullptr, nullptr, nullptr, nullptr, 0};
        if (small) small_view = view(target);
        else LTR_TRY(build_grid(ctx, target, k, thr, &gb));
        LTR_TRY(dev_alloc(ctx, &p, (size_t)n * (3 * sizeof(float) + 1) + 64));
        gx = (float*)p; gy = gx + n; gz = gy + n; label = (uint8_t*)(gz + n);
        knn_scan_label_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(view(scans.pts), scans.d_off, scans.K, poses.d, pose_offset, ctx->d_ext,
            ctx->ext_identity ? 1 : 0, ctx->cfg.transform_order, gb.g, gb.cell_start, gb.sorted, k, thr, label, gx, gy, gz, small_view);
        LTR_LAUNCH_CHECK(ctx);
        g_grid.release();
        // re-localise in place (every point; partitioning afterwards keeps the arithmetic identical to the reference,
        // which transforms the two partitions separately with the same per-point operations)
        // NOTE: intensity is carried unchanged.
        DevCloud tmp = scans.pts;  // reuse layout: write relocalised xyz into a scratch cloud with the same stride
        ltr_cloud scratch;
        LTR_TRY(cloud_new(ctx, n, &scratch));
        struct CloudGuard { ltr_ctx* c; ltr_cloud h; bool armed; ~CloudGuard() { if (armed) ltr_cloud_free(c, h); } } g_scratch{ctx, scratch, true};   // error paths below
        DevCloud sc = ctx->clouds[scratch];
        knn_relocalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(gx, gy, gz, n, scans.d_off, scans.K, poses.d, pose_offset, ctx->d_ext,
            ctx->ext_identity ? 1 : 0, ctx->cfg.transform_order, sc.x(), sc.y(), sc.z());
        LTR_LAUNCH_CHECK(ctx);
        LTR_CUDA(ctx, cudaMemcpyAsync(sc.i(), tmp.i(), (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        g_scratch.armed = false;
        DevScanSet view_set = scans;
        view_set.pts = sc;
        const int rc = split_scanset_by_flag(ctx, view_set, label, out_coexist, out_diff);
        ltr_cloud_free(ctx, scratch);
        return rc;
    }
    // empty input: empty outputs with K keyframes
    std::vector<int64_t> off((size_t)scans.K + 1, 0);
    if (out_coexist) LTR_TRY(scanset_new(ctx, off, out_coexist));
    if (out_diff) LTR_TRY(scanset_new(ctx, off, out_diff));
    return LTR_OK;
}

int ltr_knn_split_cloud(ltr_ctx* ctx, ltr_cloud query_h, ltr_cloud target_h, int32_t k, float thr, ltr_cloud* out_near, ltr_cloud* out_far) {
    ApiTrace tr__(ctx, "ltr_knn_split_cloud");
    if (!ctx || !out_near || !out_far) return fail(ctx, LTR_ERR_INVALID, "null argument");
    DevCloud *qp, *tp;
    LTR_TRY(cloud_get(ctx, query_h, &qp));
    LTR_TRY(cloud_get(ctx, target_h, &tp));
    if (k < 1 || k > kMaxK) return fail(ctx, LTR_ERR_UNSUPPORTED, "k = %d outside [1, %d]", k, kMaxK);
    if (!(thr > 0.0f)) return fail(ctx, LTR_ERR_INVALID, "threshold must be positive");
    if (tp->n < 1) return fail(ctx, LTR_ERR_UNSUPPORTED, "empty target (PCL's nearestKSearch asserts on an empty tree)");
    const bool small = tp->n < k;
    const DevCloud q = *qp, target = *tp;
    LTR_TRY(cloud_new(ctx, q.n, out_near));
    LTR_TRY(cloud_new(ctx, q.n, out_far));
    if (q.n == 0) return LTR_OK;
    GridBuf gb;
    ScratchGuard g_grid(ctx, &gb.block);
    PtrView small_view{nullptr, nullptr, nullptr, nullptr, 0};
    if (small) small_view = view(target);
    else LTR_TRY(build_grid(ctx, target, k, thr, &gb));
    void* p = nullptr;
    ScratchGuard g_p(ctx, &p);
    LTR_TRY(dev_alloc(ctx, &p, (size_t)q.n));
    knn_cloud_label_kernel<<<(unsigned)((q.n + 127) / 128), 128, 0, ctx->stream>>>(view(q), gb.g, gb.cell_start, gb.sorted, k, thr, (uint8_t*)p, small_view);
    LTR_LAUNCH_CHECK(ctx);
    g_grid.release();
    DevCloud o0 = ctx->clouds[*out_near], o1 = ctx->clouds[*out_far];
    int64_t nf = 0;
    LTR_TRY(stable_partition_by_flag(ctx, q, (const uint8_t*)p, &nf, &o0, &o1));
    ctx->clouds[*out_near].n = o0.n;
    ctx->clouds[*out_far].n = o1.n;
    return LTR_OK;
}

}  // extern "C"
