// ORACLE (test infrastructure, not product code) -- see oracle.h for the scope statement.
// First-party logic pinned against the compiled reference sources (oracle/ref_shim, tests/test_ref_pin.py);
// PARITY UNPINNED for the third-party semantics restated here (no reference tests exist, see comments).
#include "oracle.h"
#ifdef _OPENMP
#include <parallel/algorithm>
#endif
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <set>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace ltr_oracle {

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Params::Params() {
    for (int i = 0; i < 16; ++i) lidar2base.m[i] = base2lidar.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

// General 4x4 inverse by cofactors, standing in for Eigen::Matrix4d::inverse() (Session.cpp:110,
// RosParamServer.cpp:30).  Eigen is not under /root/reference (UNPINNED): its op order may differ in
// the last double bits; the C-ABI takes inverse poses as INPUT so oracle and GPU share the same doubles.
Mat4 inverse4x4(const Mat4& A) {
    const double* m = A.m;
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    Mat4 R;
    const double idet = 1.0 / det;
    for (int i = 0; i < 16; ++i) R.m[i] = inv[i] * idet;
    return R;
}

// ------------------------------------------------------------------------------------------------
// Range images.  Sequential semantics: strict "<" in point order => per pixel the
// minimum range wins and among equal ranges the LOWEST point index.  The reference's OMP loop is
// racy (Removerter.cpp:142-148, utility.cpp:127-133); `faithful` mode reproduces that structure for
// timing only and its images are never used for parity.
// ------------------------------------------------------------------------------------------------
void scan2RangeImg(const Cloud& scan, const Params& p, int rows, int cols, std::vector<float>& rimg) {
    rimg.assign((size_t)rows * cols, kFlagNoPOINT);
    const int n = (int)scan.size();
    if (p.faithful) {
#pragma omp parallel for num_threads(p.omp_cores)
        for (int i = 0; i < n; ++i) {
            const Pt& q = scan[i];
            const Sph s = cart2sph(q.x, q.y, q.z);
            int r, c;
            pixelIndex(s, p.vfov, p.hfov, rows, cols, &r, &c);
            if (s.r < rimg[(size_t)r * cols + c]) rimg[(size_t)r * cols + c] = s.r;
        }
        return;
    }
    for (int i = 0; i < n; ++i) {
        const Pt& q = scan[i];
        const Sph s = cart2sph(q.x, q.y, q.z);
        int r, c;
        pixelIndex(s, p.vfov, p.hfov, rows, cols, &r, &c);
        if (s.r < rimg[(size_t)r * cols + c]) rimg[(size_t)r * cols + c] = s.r;
    }
}

void map2RangeImg(const Cloud& scan, const Params& p, int rows, int cols, std::vector<float>& rimg, std::vector<int>& ptidx) {
    rimg.assign((size_t)rows * cols, kFlagNoPOINT);
    ptidx.assign((size_t)rows * cols, 0);  // utility.cpp:104: index image initialised to 0
    const int n = (int)scan.size();
    if (p.faithful) {
#pragma omp parallel for num_threads(p.omp_cores < 16 ? p.omp_cores : 16)  // utility.cpp:109 hard-codes 16
        for (int i = 0; i < n; ++i) {
            const Pt& q = scan[i];
            const Sph s = cart2sph(q.x, q.y, q.z);
            int r, c;
            pixelIndex(s, p.vfov, p.hfov, rows, cols, &r, &c);
            const size_t px = (size_t)r * cols + c;
            if (s.r < rimg[px]) { rimg[px] = s.r; ptidx[px] = i; }
        }
        return;
    }
    for (int i = 0; i < n; ++i) {
        const Pt& q = scan[i];
        const Sph s = cart2sph(q.x, q.y, q.z);
        int r, c;
        pixelIndex(s, p.vfov, p.hfov, rows, cols, &r, &c);
        const size_t px = (size_t)r * cols + c;
        if (s.r < rimg[px]) { rimg[px] = s.r; ptidx[px] = i; }
    }
}

void transformPointCloud(const Cloud& in, Cloud& out, const Mat4& T, int order) {
    if (&in != &out) out.resize(in.size());
    const size_t n = in.size();
    for (size_t i = 0; i < n; ++i) {
        Pt q = in[i];
        transformPoint(T.m, order, q.x, q.y, q.z, &q.x, &q.y, &q.z);
        out[i] = q;
    }
}

void transformGlobalMapToLocal(const Cloud& map_global, const Mat4& inv_pose, const Mat4& base2lidar, int order, Cloud& map_local) {
    transformPointCloud(map_global, map_local, inv_pose, order);   // utility.cpp:70
    transformPointCloud(map_local, map_local, base2lidar, order);  // utility.cpp:71
}

Cloud parseProjectedPoints(const Cloud& map_local, const Params& p, int rows, int cols) {
    std::vector<float> rimg;
    std::vector<int> ptidx;
    map2RangeImg(map_local, p, rows, cols, rimg, ptidx);
    Cloud out;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const int id = ptidx[(size_t)r * cols + c];
            if (id == 0) continue;  // utility.cpp:82: 0 means "no point" (so map point 0 is never emitted)
            out.push_back(map_local[id]);
        }
    return out;
}

Cloud local2global(const Cloud& scan_local, const Mat4& pose, const Mat4& lidar2base, int order) {
    Cloud g;
    transformPointCloud(scan_local, g, lidar2base, order);  // utility.cpp:164
    transformPointCloud(g, g, pose, order);                 // utility.cpp:165
    return g;
}

Cloud global2local(const Cloud& scan_global, const Mat4& inv_pose, const Mat4& base2lidar, int order) {
    Cloud l;
    transformPointCloud(scan_global, l, inv_pose, order);  // utility.cpp:198
    transformPointCloud(l, l, base2lidar, order);          // utility.cpp:199
    return l;
}

Cloud mergeScansWithinGlobalCoordUtil(const std::vector<Cloud>& scans, const std::vector<Mat4>& poses, const Mat4& lidar2base, int order) {
    Cloud merged;
    for (size_t k = 0; k < scans.size(); ++k) {
        Cloud g = local2global(scans[k], poses[k], lidar2base, order);  // utility.cpp:184-185
        merged.insert(merged.end(), g.begin(), g.end());               // utility.cpp:188
    }
    return merged;
}

// ------------------------------------------------------------------------------------------------
// octreeDownsampling (utility.cpp:204-219) = pcl::octree::OctreePointCloudVoxelCentroid<PointXYZI>:
// setInputCloud; defineBoundingBox(); addPointsFromInputCloud(); getVoxelCentroids().
// PCL is not under /root/reference: restated from the published PCL 1.10 sources (UNPINNED,
// SURVEY.md §A.5):
//  * bounding box = getMinMax3D of the cloud, max padded by 512*FLT_EPSILON (float add), then
//    expanded to a cube of side 2^depth*resolution, centred by moving min/max outwards by half
//    the slack when the slack exceeds FLT_EPSILON (OctreePointCloud::getKeyBitSize);
//  * key = (unsigned)(((double)p - min) / resolution) per axis (genOctreeKeyforPoint);
//  * leaf container sums x,y,z,intensity in f32 in insertion order, centroid = sum / (float)count;
//  * getVoxelCentroids walks depth-first, children in index order (x_bit<<2)|(y_bit<<1)|z_bit from
//    the MSB  =>  ascending Morton order with x most significant.
// Returns -1 (unsupported) if a point falls outside the defined box (PCL would grow the tree).
// ------------------------------------------------------------------------------------------------
struct OctreeBox { double min[3], res; int depth; };

static OctreeBox octreeDefineBox(const Cloud& src, float leaf) {
    OctreeBox b;
    b.res = (double)leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const Pt& q : src) {
        mn[0] = std::min(mn[0], q.x); mn[1] = std::min(mn[1], q.y); mn[2] = std::min(mn[2], q.z);
        mx[0] = std::max(mx[0], q.x); mx[1] = std::max(mx[1], q.y); mx[2] = std::max(mx[2], q.z);
    }
    const float pad = FLT_EPSILON * 512.0f;
    double lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] = mn[d];
        hi[d] = (float)(mx[d] + pad);  // float add, then widened
        const double a = std::min(lo[d], hi[d]);
        const double c = std::max(a, hi[d]);
        lo[d] = a; hi[d] = c;
    }
    const float eps = FLT_EPSILON;
    unsigned mk[3];
    for (int d = 0; d < 3; ++d) mk[d] = (unsigned)std::ceil((hi[d] - lo[d] - eps) / b.res);
    const unsigned max_voxels = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
    const double lg = std::log((double)max_voxels) / std::log(2.0);
    b.depth = (int)std::max(std::min(32u, (unsigned)std::ceil(lg - eps)), 0u);
    const double side = (double)(1u << b.depth) * b.res;
    for (int d = 0; d < 3; ++d) {
        const double over = (side - (hi[d] - lo[d])) / 2.0;
        if (over > eps) { lo[d] -= over; hi[d] += over; }
        b.min[d] = lo[d];
    }
    return b;
}

static inline uint64_t mortonXYZ(unsigned kx, unsigned ky, unsigned kz, int depth) {
    uint64_t code = 0;
    for (int bit = depth - 1; bit >= 0; --bit)
        code = (code << 3) | (uint64_t)((((kx >> bit) & 1u) << 2) | (((ky >> bit) & 1u) << 1) | ((kz >> bit) & 1u));
    return code;
}

int octreeDownsampling(const Cloud& src_in, Cloud& dst, float leaf) {
    const Cloud src = src_in;  // src and dst may alias (utility.cpp call sites pass the same cloud)
    dst.clear();
    if (src.empty()) return 0;
    const OctreeBox b = octreeDefineBox(src, leaf);
    if (b.depth > 21) return -1;  // 3*depth must fit 63 bits
    const size_t n = src.size();
    std::vector<std::pair<uint64_t, uint32_t>> keyed(n);
    const unsigned lim = (b.depth >= 32) ? 0xffffffffu : ((1u << b.depth) - 1u);
    for (size_t i = 0; i < n; ++i) {
        const Pt& q = src[i];
        const double fx = ((double)q.x - b.min[0]) / b.res;
        const double fy = ((double)q.y - b.min[1]) / b.res;
        const double fz = ((double)q.z - b.min[2]) / b.res;
        if (fx < 0 || fy < 0 || fz < 0) return -1;
        const unsigned kx = (unsigned)fx, ky = (unsigned)fy, kz = (unsigned)fz;
        if (kx > lim || ky > lim || kz > lim) return -1;
        keyed[i] = std::make_pair(mortonXYZ(kx, ky, kz, b.depth), (uint32_t)i);
    }
    // (code, original index) pairs are all distinct, so any correct sort gives the same order: stable w.r.t. insertion order
#ifdef _OPENMP
    if (n > (size_t)1 << 20) __gnu_parallel::sort(keyed.begin(), keyed.end());
    else
#endif
    std::sort(keyed.begin(), keyed.end());
    size_t i = 0;
    while (i < n) {
        size_t j = i;
        float sx = 0.0f, sy = 0.0f, sz = 0.0f, si = 0.0f;
        while (j < n && keyed[j].first == keyed[i].first) {
            const Pt& q = src[keyed[j].second];
            sx += q.x; sy += q.y; sz += q.z; si += q.i;
            ++j;
        }
        const float cnt = (float)(j - i);
        Pt c;
        c.x = sx / cnt; c.y = sy / cnt; c.z = sz / cnt; c.i = si / cnt;
        dst.push_back(c);
        i = j;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// One pass over all source keyframes (Removerter.cpp:542-593 HD/revert, :485-540 ND, :429-482 PD).
// ------------------------------------------------------------------------------------------------
static void dynIdxOfOneScan(const Cloud& map_global, const Cloud& scan, const Mat4& inv_pose, const Params& p, int mode,
                            int rows, int cols, float thres, std::vector<float>& scan_rimg, std::vector<float>& map_rimg,
                            std::vector<int>& ptidx, Cloud* map_local_buf, std::vector<int>& out) {
    scan2RangeImg(scan, p, rows, cols, scan_rimg);
    if (p.faithful) {
        transformGlobalMapToLocal(map_global, inv_pose, p.base2lidar, p.transform_order, *map_local_buf);
        map2RangeImg(*map_local_buf, p, rows, cols, map_rimg, ptidx);
    } else {
        // fused: identical arithmetic per point, no materialised map_local
        map_rimg.assign((size_t)rows * cols, kFlagNoPOINT);
        ptidx.assign((size_t)rows * cols, 0);
        const int n = (int)map_global.size();
        for (int i = 0; i < n; ++i) {
            float x, y, z;
            transformPoint(inv_pose.m, p.transform_order, map_global[i].x, map_global[i].y, map_global[i].z, &x, &y, &z);
            transformPoint(p.base2lidar.m, p.transform_order, x, y, z, &x, &y, &z);
            const Sph s = cart2sph(x, y, z);
            int r, c;
            pixelIndex(s, p.vfov, p.hfov, rows, cols, &r, &c);
            const size_t px = (size_t)r * cols + c;
            if (s.r < map_rimg[px]) { map_rimg[px] = s.r; ptidx[px] = i; }
        }
    }
    // calcDescrepancyAndParseDynamicPointIdx (Removerter.cpp:381-413)
    const size_t npx = (size_t)rows * cols;
    for (size_t px = 0; px < npx; ++px) {
        const float diff = (mode == MODE_ND) ? (map_rimg[px] - scan_rimg[px])   // :516 reversed diff
                                             : (scan_rimg[px] - map_rimg[px]);  // :572, :459
        if (diff < kValidDiffUpperBound && diff > thres) out.push_back(ptidx[px]);
    }
}

std::vector<int> calcDescrepancyAndParseDynamicPointIdxForEachScan(
    const Cloud& map_global, const std::vector<Cloud>& scans, const std::vector<Mat4>& inv_poses,
    const Params& p, int mode, int rows, int cols, float diff_thres) {
    const int K = (int)scans.size();
    std::vector<int> all;
    if (p.faithful) {
        std::vector<float> srimg, mrimg;
        std::vector<int> ptidx;
        Cloud map_local;
        for (int k = 0; k < K; ++k)
            dynIdxOfOneScan(map_global, scans[k], inv_poses[k], p, mode, rows, cols, diff_thres, srimg, mrimg, ptidx, &map_local, all);
        std::set<int> s(all.begin(), all.end());   // Removerter.cpp:589
        return std::vector<int>(s.begin(), s.end());
    }
    const int T = std::max(1, p.threads);
    std::vector<std::vector<int>> per(T);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        std::vector<float> srimg, mrimg;
        std::vector<int> ptidx;
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < K; ++k)
            dynIdxOfOneScan(map_global, scans[k], inv_poses[k], p, mode, rows, cols, diff_thres, srimg, mrimg, ptidx, nullptr, per[t]);
    }
    for (auto& v : per) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    return all;
}

// getStaticIdxFromDynamicIdx (Removerter.cpp:675-687) builds linspace<int>(0, N, N) (utility.h:158-167):
// step = N / (N-1) in int arithmetic == 1 only for N >= 3; N == 1 divides by zero and N == 2 yields
// {0, 2} (out of range) in the reference.  Those sizes are rejected here (-1).
int partitionByIdx(const Cloud& map, const std::vector<int>& dyn_idx, Cloud& stat, Cloud& dyn) {
    const size_t n = map.size();
    stat.clear(); dyn.clear();
    if (n == 0) return 0;
    if (n < 3) return -1;
    std::vector<uint8_t> flag(n, 0);
    for (int id : dyn_idx) { flag[(size_t)id] = 1; dyn.push_back(map[(size_t)id]); }  // ExtractIndices keeps index order (sorted)
    for (size_t i = 0; i < n; ++i) if (!flag[i]) stat.push_back(map[i]);
    return 0;
}

static int partitionByIdxFaithful(const Cloud& map, const std::vector<int>& dyn_idx, Cloud& stat, Cloud& dyn) {
    // reference structure: std::set of ALL indices, erase dynamic ones (Removerter.cpp:677-684)
    const int n = (int)map.size();
    stat.clear(); dyn.clear();
    if (n == 0) return 0;
    if (n < 3) return -1;
    std::set<int> all;
    for (int i = 0; i < n; ++i) all.insert(all.end(), i);
    for (int id : dyn_idx) { all.erase(id); dyn.push_back(map[(size_t)id]); }
    for (int id : all) stat.push_back(map[(size_t)id]);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// kNN partition of one scan (Session.cpp:537-607 / 610-642)
// ------------------------------------------------------------------------------------------------
void partitionScanByKnn(const Cloud& scan_local, const Mat4& pose, const Mat4& inv_pose, const KdTree& target,
                        const Params& p, int k, float thr, Cloud& coexist_local, Cloud& diff_local, std::vector<uint8_t>* labels) {
    // NOTE: the reference passes kSE3MatExtrinsicPoseBasetoLiDAR as the lidar->base argument (Session.cpp:545, 618);
    // reproduced as written.
    const Cloud g = local2global(scan_local, pose, p.base2lidar, p.transform_order);
    Cloud co, di;
    if (labels) labels->assign(g.size(), 0);
    std::vector<float> d2((size_t)std::max(k, 1));
    for (size_t i = 0; i < g.size(); ++i) {
        const float q[3] = {g[i].x, g[i].y, g[i].z};
        const int kk = target.knn(q, k, d2.data());
        double sum = 0.0;  // accumulate(..., 0.0): double accumulator (Session.cpp:593)
        for (int j = 0; j < kk; ++j) sum += d2[j];
        const float sumf = (float)sum;
        const float avg = sumf / float(k);  // divides by k even if fewer were found (Session.cpp:594)
        if (std::abs(avg) < thr) co.push_back(g[i]);
        else { di.push_back(g[i]); if (labels) (*labels)[i] = 1; }
    }
    coexist_local = global2local(co, inv_pose, p.base2lidar, p.transform_order);  // Session.cpp:603
    diff_local = global2local(di, inv_pose, p.base2lidar, p.transform_order);     // Session.cpp:604
}

// ------------------------------------------------------------------------------------------------
// Removerter
// ------------------------------------------------------------------------------------------------
Removerter::Removerter() {
    central_sess_.sess_type_ = "Central";
    query_sess_.sess_type_ = "Query";
    hd_schedule.push_back({OP_REMOVE, 2.5f});
}

void Removerter::precleaningKeyframes(float radius) {
    for (Session* s : {&central_sess_, &query_sess_})
        for (Cloud& scan : s->keyframe_scans_) {
            Cloud cleaned;
            for (const Pt& q : scan) {
                const Sph sp = cart2sph(q.x, q.y, q.z);
                if ((sp.r < radius) & (q.z < 0.5) & (-0.5 < q.z)) continue;  // Session.cpp:522-526
                cleaned.push_back(q);
            }
            scan.swap(cleaned);
        }
}

void Removerter::makeGlobalMap() {
    for (Session* s : {&central_sess_, &query_sess_}) {
        s->map_global_orig_ = mergeScansWithinGlobalCoordUtil(s->keyframe_scans_, s->keyframe_poses_, P.lidar2base, P.transform_order);  // Session.cpp:186-202
        octreeDownsampling(s->map_global_orig_, s->map_global_curr_, P.downsample_voxel);  // Removerter.cpp:225
        saved["OriginalNoisy" + s->sess_type_ + "MapGlobal"] = s->map_global_curr_;         // :231
    }
}

static void partitionCurrentMapGeneric(Removerter& R, const Cloud& map, const std::vector<Cloud>& scans, const std::vector<Mat4>& inv_poses,
                                       int mode, float res, float thres, const char* what, Cloud& stat, Cloud& dyn) {
    int rows, cols;
    resetRimgSize(R.P.vfov, R.P.hfov, res, &rows, &cols);  // Removerter.cpp:807
    const std::vector<int> dyn_idx = calcDescrepancyAndParseDynamicPointIdxForEachScan(map, scans, inv_poses, R.P, mode, rows, cols, thres);
    if (R.P.faithful) partitionByIdxFaithful(map, dyn_idx, stat, dyn);
    else partitionByIdx(map, dyn_idx, stat, dyn);
    PassLog l;
    l.what = what; l.n_map = (long)map.size(); l.n_dynamic = (long)dyn_idx.size(); l.n_static_after = l.n_dynamic_after = -1;
    R.log.push_back(l);
}

void Removerter::removeOnce(Session& t, const Session& s, float res) {
    Cloud st, dy;
    partitionCurrentMapGeneric(*this, t.map_global_curr_, s.keyframe_scans_, s.keyframe_inverse_poses_, MODE_HD, res, 0.1f, "removeOnce", st, dy);
    t.map_global_curr_static_ = st;                                             // :894-895
    octreeDownsampling(t.map_global_curr_static_, t.map_global_curr_static_, 0.05f);  // :896
    t.map_global_curr_ = t.map_global_curr_static_;                             // :899-900
    t.map_global_curr_dynamic_.insert(t.map_global_curr_dynamic_.end(), dy.begin(), dy.end());  // :902
    octreeDownsampling(t.map_global_curr_dynamic_, t.map_global_curr_dynamic_, 0.05f);          // :903
    log.back().n_static_after = (long)t.map_global_curr_static_.size();
    log.back().n_dynamic_after = (long)t.map_global_curr_dynamic_.size();
}

void Removerter::revertOnce(Session& t, const Session& s, float res) {
    Cloud st, dy;
    partitionCurrentMapGeneric(*this, t.map_global_curr_, s.keyframe_scans_, s.keyframe_inverse_poses_, MODE_HD, res, 0.1f, "revertOnce", st, dy);
    t.map_global_curr_dynamic_ = dy;                                            // :919-920
    octreeDownsampling(t.map_global_curr_dynamic_, t.map_global_curr_dynamic_, 0.05f);  // :921
    t.map_global_curr_ = t.map_global_curr_dynamic_;                            // :924-925
    t.map_global_curr_static_.insert(t.map_global_curr_static_.end(), st.begin(), st.end());  // :927
    octreeDownsampling(t.map_global_curr_static_, t.map_global_curr_static_, 0.05f);          // :928
    log.back().n_static_after = (long)t.map_global_curr_static_.size();
    log.back().n_dynamic_after = (long)t.map_global_curr_dynamic_.size();
}

void Removerter::iremoveOnceForND(Session& t, const Session& s, float res) {
    Cloud st, dy;
    partitionCurrentMapGeneric(*this, t.map_global_nd_, s.keyframe_scans_static_projected_, s.keyframe_inverse_poses_, MODE_ND, res, 0.1f, "iremoveOnceForND", st, dy);
    t.map_global_nd_strong_ = st;                                               // :843-844
    octreeDownsampling(t.map_global_nd_strong_, t.map_global_nd_strong_, 0.05f);
    t.map_global_nd_ = t.map_global_nd_strong_;                                 // :848-849
    t.map_global_nd_weak_.insert(t.map_global_nd_weak_.end(), dy.begin(), dy.end());  // :851
    octreeDownsampling(t.map_global_nd_weak_, t.map_global_nd_weak_, 0.05f);
    log.back().n_static_after = (long)t.map_global_nd_strong_.size();
    log.back().n_dynamic_after = (long)t.map_global_nd_weak_.size();
}

void Removerter::removeOnceForPD(Session& t, const Session& s, float res) {
    Cloud st, dy;
    partitionCurrentMapGeneric(*this, t.map_global_pd_, s.keyframe_scans_static_projected_, s.keyframe_inverse_poses_, MODE_PD, res, 0.1f, "removeOnceForPD", st, dy);
    t.map_global_pd_strong_ = st;                                               // :868-869
    octreeDownsampling(t.map_global_pd_strong_, t.map_global_pd_strong_, 0.05f);
    t.map_global_pd_ = t.map_global_pd_strong_;                                 // :873-874
    t.map_global_pd_weak_.insert(t.map_global_pd_weak_.end(), dy.begin(), dy.end());  // :876
    octreeDownsampling(t.map_global_pd_weak_, t.map_global_pd_weak_, 0.05f);
    log.back().n_static_after = (long)t.map_global_pd_strong_.size();
    log.back().n_dynamic_after = (long)t.map_global_pd_weak_.size();
}

void Removerter::runSchedule(Session& s) {
    for (const ScheduleOp& op : hd_schedule) {
        if (op.op == OP_REMOVE) {
            removeOnce(s, s, op.res);
        } else {
            s.map_global_curr_ = s.map_global_curr_dynamic_;  // resetCurrrentMapAsDynamic (:714-732)
            revertOnce(s, s, op.res);
            s.map_global_curr_ = s.map_global_curr_static_;   // resetCurrrentMapAsStatic (:734-737)
        }
    }
}

void Removerter::extractHighDynPointsViaKnnDiff(Session& s, const Cloud& target) {
    KdTree tree;
    tree.build(&target[0].x, 4, (int)target.size());
    const int K = (int)s.keyframe_scans_.size();
    s.keyframe_scans_dynamic_.assign(K, Cloud());
    const int T = P.faithful ? P.omp_cores : std::max(1, P.threads);
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        Cloud co, di;
        partitionScanByKnn(s.keyframe_scans_[k], s.keyframe_poses_[k], s.keyframe_inverse_poses_[k], tree, P, P.num_knn, P.knn_thr, co, di, nullptr);
        s.keyframe_scans_dynamic_[k] = di;
    }
}

void Removerter::removeHighDynamicPoints() {
    double t0 = now_s();
    runSchedule(central_sess_);  // shipped: removeOnce(central, central, 2.5) (:1584)
    runSchedule(query_sess_);    // shipped: removeOnce(query, query, 2.5) (:1587)
    timing["hd_remove"] += now_s() - t0;
    if (do_high_dyn_knn) {
        t0 = now_s();
        extractHighDynPointsViaKnnDiff(central_sess_, central_sess_.map_global_curr_static_);  // :1591
        extractHighDynPointsViaKnnDiff(query_sess_, query_sess_.map_global_curr_static_);      // :1592
        Cloud c = mergeScansWithinGlobalCoordUtil(central_sess_.keyframe_scans_dynamic_, central_sess_.keyframe_poses_, P.lidar2base, P.transform_order);
        Cloud q = mergeScansWithinGlobalCoordUtil(query_sess_.keyframe_scans_dynamic_, query_sess_.keyframe_poses_, P.lidar2base, P.transform_order);
        octreeDownsampling(c, c, 0.05f);
        octreeDownsampling(q, q, 0.05f);
        saved["central_sess_high_dyn"] = c;  // :1600
        saved["query_sess_high_dyn"] = q;    // :1601
        timing["hd_knn"] += now_s() - t0;
    }
}

void Removerter::parseScansViaProjection(const Session& s, const Cloud& map, std::vector<Cloud>& out) {
    const int K = (int)s.keyframe_scans_.size();  // Session.cpp:353 loops over keyframe_scans_.size()
    out.assign(K, Cloud());
    int rows, cols;
    resetRimgSize(P.vfov, P.hfov, 3.0f, &rows, &cols);  // kReprojectionAlpha (Session.h:13)
    if (P.faithful) {
        Cloud map_local;
        for (int k = 0; k < K; ++k) {
            transformGlobalMapToLocal(map, s.keyframe_inverse_poses_[k], P.base2lidar, P.transform_order, map_local);
            out[k] = parseProjectedPoints(map_local, P, rows, cols);
        }
        return;
    }
    const int T = std::max(1, P.threads);
#pragma omp parallel num_threads(T)
    {
        Cloud map_local;
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < K; ++k) {
            transformGlobalMapToLocal(map, s.keyframe_inverse_poses_[k], P.base2lidar, P.transform_order, map_local);
            out[k] = parseProjectedPoints(map_local, P, rows, cols);
        }
    }
}

void Removerter::parseStaticScansViaProjection() {
    const double t0 = now_s();
    parseScansViaProjection(central_sess_, central_sess_.map_global_curr_, central_sess_.keyframe_scans_static_projected_);  // Session.cpp:305-308
    parseScansViaProjection(query_sess_, query_sess_.map_global_curr_, query_sess_.keyframe_scans_static_projected_);
    timing["parse_static"] += now_s() - t0;
}

void Removerter::extractLowDynPointsViaKnnDiff(Session& s, const Cloud& target) {
    if (P.faithful) {  // dead weight the reference executes: 0.4 m downsample for the disabled ICP (Session.cpp:395-402)
        Cloud down;
        octreeDownsampling(target, down, 0.4f);
    }
    KdTree tree;
    tree.build(&target[0].x, 4, (int)target.size());  // Session.cpp:404
    const int K = (int)s.keyframe_scans_static_projected_.size();
    s.scans_knn_coexist_.assign(K, Cloud());
    s.scans_knn_diff_.assign(K, Cloud());
    const int T = P.faithful ? P.omp_cores : std::max(1, P.threads);
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (int k = 0; k < K; ++k)
        partitionScanByKnn(s.keyframe_scans_static_projected_[k], s.keyframe_poses_[k], s.keyframe_inverse_poses_[k], tree, P,
                           P.num_knn, P.knn_thr, s.scans_knn_coexist_[k], s.scans_knn_diff_[k], nullptr);
}

void Removerter::removeWeakNDMapPointsHavingStrongNDInNear(Session& s) {
    if (s.map_global_nd_strong_.empty()) return;  // Session.cpp:454-455
    KdTree tree;
    tree.build(&s.map_global_nd_strong_[0].x, 4, (int)s.map_global_nd_strong_.size());
    Cloud added, new_weak;
    const int k = 2;          // Session.cpp:468
    const float thr = 1.0f;   // Session.cpp:469
    float d2[2];
    for (const Pt& q : s.map_global_nd_weak_) {
        const float qq[3] = {q.x, q.y, q.z};
        const int kk = tree.knn(qq, k, d2);
        double sum = 0.0;
        for (int j = 0; j < kk; ++j) sum += d2[j];
        const float avg = (float)sum / float(k);
        if (std::abs(avg) < thr) added.push_back(q); else new_weak.push_back(q);
    }
    s.map_global_nd_strong_.insert(s.map_global_nd_strong_.end(), added.begin(), added.end());  // :482
    s.map_global_nd_weak_ = new_weak;                                                            // :483
}

void Removerter::detectLowDynamicPoints() {
    Session& C = central_sess_;
    Session& Q = query_sess_;
    double t0 = now_s();
    extractLowDynPointsViaKnnDiff(C, Q.map_global_curr_static_);  // :1416
    extractLowDynPointsViaKnnDiff(Q, C.map_global_curr_static_);  // :1418
    timing["ld_knn"] += now_s() - t0;

    t0 = now_s();
    // strong ND
    C.map_global_nd_ = mergeScansWithinGlobalCoordUtil(C.scans_knn_diff_, C.keyframe_poses_, P.lidar2base, P.transform_order);  // Session.cpp:430-435
    octreeDownsampling(C.map_global_nd_, C.map_global_nd_, 0.05f);
    for (int i = 0; i < 3; ++i) iremoveOnceForND(C, Q, 2.5f);  // filterStrongND (:1403-1411)
    removeWeakNDMapPointsHavingStrongNDInNear(C);               // :1424
    // strong PD
    Q.map_global_pd_ = mergeScansWithinGlobalCoordUtil(Q.scans_knn_diff_, Q.keyframe_poses_, P.lidar2base, P.transform_order);  // Session.cpp:437-445
    octreeDownsampling(Q.map_global_pd_, Q.map_global_pd_, 0.05f);
    Q.map_global_pd_orig_ = Q.map_global_pd_;
    for (int i = 0; i < 3; ++i) removeOnceForPD(Q, C, 2.5f);   // filterStrongPD (:1395-1401)
    // revertStrongPDMapPointsHavingWeakPDInNear: empty TODO (Session.cpp:447-450)
    C.map_global_pd_ = Q.map_global_pd_;                        // :1434
    C.map_global_pd_orig_ = Q.map_global_pd_orig_;              // :1435
    C.map_global_pd_strong_ = Q.map_global_pd_strong_;          // :1436
    timing["ld_filter"] += now_s() - t0;

    // always-on viz block (:1442-1480), including its in-place re-downsampling side effects
    t0 = now_s();
    Cloud uq = mergeScansWithinGlobalCoordUtil(Q.scans_knn_coexist_, Q.keyframe_poses_, P.lidar2base, P.transform_order);
    octreeDownsampling(uq, uq, 0.05f); saved["union_map_queryside"] = uq;
    Cloud uc = mergeScansWithinGlobalCoordUtil(C.scans_knn_coexist_, C.keyframe_poses_, P.lidar2base, P.transform_order);
    octreeDownsampling(uc, uc, 0.05f); saved["union_map_centralside"] = uc;
    Cloud pd = mergeScansWithinGlobalCoordUtil(Q.scans_knn_diff_, Q.keyframe_poses_, P.lidar2base, P.transform_order);
    octreeDownsampling(pd, pd, 0.05f); saved["pd_map"] = pd;
    Cloud nd = mergeScansWithinGlobalCoordUtil(C.scans_knn_diff_, C.keyframe_poses_, P.lidar2base, P.transform_order);
    octreeDownsampling(nd, nd, 0.05f); saved["nd_map"] = nd;
    if (!C.map_global_nd_strong_.empty()) {
        octreeDownsampling(C.map_global_nd_strong_, C.map_global_nd_strong_, 0.05f);
        saved["strong_nd_map"] = C.map_global_nd_strong_;
    }
    octreeDownsampling(C.map_global_nd_weak_, C.map_global_nd_weak_, 0.05f); saved["weak_nd_map"] = C.map_global_nd_weak_;
    octreeDownsampling(Q.map_global_pd_strong_, Q.map_global_pd_strong_, 0.05f); saved["strong_pd_map"] = Q.map_global_pd_strong_;
    octreeDownsampling(Q.map_global_pd_weak_, Q.map_global_pd_weak_, 0.05f); saved["weak_pd_map"] = Q.map_global_pd_weak_;
    timing["ld_merge_viz"] += now_s() - t0;
}

void Removerter::updateCurrentMap() {
    Session& C = central_sess_;
    Session& Q = query_sess_;
    Cloud uq = mergeScansWithinGlobalCoordUtil(Q.scans_knn_coexist_, Q.keyframe_poses_, P.lidar2base, P.transform_order);
    octreeDownsampling(uq, uq, 0.05f);
    Cloud uc = mergeScansWithinGlobalCoordUtil(C.scans_knn_coexist_, C.keyframe_poses_, P.lidar2base, P.transform_order);
    octreeDownsampling(uc, uc, 0.05f);
    Cloud upd = uq;                                                           // :1495
    upd.insert(upd.end(), uc.begin(), uc.end());                              // :1496
    upd.insert(upd.end(), C.map_global_nd_weak_.begin(), C.map_global_nd_weak_.end());  // :1500
    Cloud strong = upd;                                                       // :1505
    strong.insert(strong.end(), C.map_global_pd_strong_.begin(), C.map_global_pd_strong_.end());  // :1506
    octreeDownsampling(strong, strong, 0.05f);                                // :1507
    upd.insert(upd.end(), C.map_global_pd_orig_.begin(), C.map_global_pd_orig_.end());  // :1511
    octreeDownsampling(upd, upd, 0.05f);                                      // :1512
    C.map_global_updated_ = upd; saved["updated_map"] = upd;                  // :1516-1517
    C.map_global_updated_strong_ = strong; saved["updated_map_strong"] = strong;  // :1519-1520
}

void Removerter::parseUpdatedStaticScansViaProjection() {
    Session& C = central_sess_;
    parseScansViaProjection(C, C.map_global_updated_, C.keyframe_scans_updated_);
    parseScansViaProjection(C, C.map_global_updated_strong_, C.keyframe_scans_updated_strong_);
}

void Removerter::parseLDScansViaProjection() {
    Session& C = central_sess_;
    parseScansViaProjection(C, C.map_global_pd_orig_, C.keyframe_scans_pd_);
    parseScansViaProjection(C, C.map_global_pd_strong_, C.keyframe_scans_strong_pd_);
    parseScansViaProjection(C, C.map_global_nd_weak_, C.keyframe_scans_weak_nd_);
    parseScansViaProjection(C, C.map_global_nd_strong_, C.keyframe_scans_strong_nd_);
}

void Removerter::updateScansScanwise() {
    Session& C = central_sess_;
    for (size_t k = 0; k < C.keyframe_scans_updated_.size(); ++k) {
        Cloud f = C.keyframe_scans_updated_[k];
        f.insert(f.end(), C.keyframe_scans_weak_nd_[k].begin(), C.keyframe_scans_weak_nd_[k].end());
        f.insert(f.end(), C.keyframe_scans_pd_[k].begin(), C.keyframe_scans_pd_[k].end());
        octreeDownsampling(f, f, 0.05f);
        C.keyframe_scans_updated_[k] = f;
    }
}

void Removerter::runStep0() {
    double t0 = now_s();
    precleaningKeyframes(2.5f);  // run() :1660
    makeGlobalMap();             // run() :1662
    timing["step0"] += now_s() - t0;
}

void Removerter::runStep12() {
    removeHighDynamicPoints();        // :1665
    parseStaticScansViaProjection();  // :1666
    detectLowDynamicPoints();         // :1669
}

void Removerter::runStep3() {
    const double t0 = now_s();
    updateCurrentMap();
    parseUpdatedStaticScansViaProjection();
    parseLDScansViaProjection();
    updateScansScanwise();
    timing["step3"] += now_s() - t0;
}

}  // namespace ltr_oracle
