// ORACLE (test infrastructure, not product code): C entry points so that tests/ and bench.py's CPU
// legs can drive the oracle through ctypes.  See oracle.h for scope and for what is / is not pinned.
#include "oracle.h"
#include <cstring>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace ltr_oracle;

static Cloud toCloud(const float* xyzi, int64_t n) {
    Cloud c((size_t)n);
    if (n > 0) std::memcpy(c.data(), xyzi, (size_t)n * sizeof(Pt));
    return c;
}
static Mat4 toMat(const double* m) { Mat4 r; std::memcpy(r.m, m, sizeof(r.m)); return r; }

static Params makeParams(float vfov, float hfov, const double* lidar2base, int order, int threads, int faithful) {
    Params p;
    p.vfov = vfov; p.hfov = hfov; p.transform_order = order; p.threads = threads; p.faithful = faithful;
    if (lidar2base) { p.lidar2base = toMat(lidar2base); p.base2lidar = inverse4x4(p.lidar2base); }
    return p;
}

static Cloud* findCloud(Removerter* R, const char* name, int sess, int kf) {
    Session& S = sess == 0 ? R->central_sess_ : R->query_sess_;
    const std::string n(name);
    if (n.rfind("saved:", 0) == 0) {
        auto it = R->saved.find(n.substr(6));
        return it == R->saved.end() ? nullptr : &it->second;
    }
#define MAPC(x) if (n == #x) return &S.x;
    MAPC(map_global_orig_) MAPC(map_global_curr_) MAPC(map_global_curr_static_) MAPC(map_global_curr_dynamic_)
    MAPC(map_global_updated_) MAPC(map_global_updated_strong_) MAPC(map_global_nd_) MAPC(map_global_nd_strong_)
    MAPC(map_global_nd_weak_) MAPC(map_global_pd_) MAPC(map_global_pd_orig_) MAPC(map_global_pd_strong_) MAPC(map_global_pd_weak_)
#undef MAPC
#define VECC(x) if (n == #x) return (kf >= 0 && kf < (int)S.x.size()) ? &S.x[(size_t)kf] : nullptr;
    VECC(keyframe_scans_) VECC(keyframe_scans_static_projected_) VECC(keyframe_scans_dynamic_) VECC(scans_knn_coexist_)
    VECC(scans_knn_diff_) VECC(keyframe_scans_updated_) VECC(keyframe_scans_updated_strong_) VECC(keyframe_scans_pd_)
    VECC(keyframe_scans_strong_pd_) VECC(keyframe_scans_strong_nd_) VECC(keyframe_scans_weak_nd_)
#undef VECC
    return nullptr;
}

extern "C" {

int ltro_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---------------- scalar / unit level ----------------
void ltro_atan2f(const float* y, const float* x, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = ref_atan2f(y[i], x[i]);
}
void ltro_libm_atan2f(const float* y, const float* x, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = atan2f(y[i], x[i]);
}
// counts bit mismatches between ref_atan2f and the container's libm on the cart2sph call pattern
int64_t ltro_atan2f_selfcheck(uint64_t seed, int64_t n) {
    int64_t bad = 0;
    uint64_t s = seed * 0x9e3779b97f4a7c15ull + 1;
    auto next = [&s]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int64_t i = 0; i < n; ++i) {
        const float x = (float)((double)(next() >> 11) / 9007199254740992.0 * 200.0 - 100.0);
        const float y = (float)((double)(next() >> 11) / 9007199254740992.0 * 200.0 - 100.0);
        const float z = (float)((double)(next() >> 11) / 9007199254740992.0 * 20.0 - 10.0);
        const float rho = sqrtf(x * x + y * y);
        if (f2u(atan2f(y, x)) != f2u(ref_atan2f(y, x))) ++bad;
        if (f2u(atan2f(z, rho)) != f2u(ref_atan2f(z, rho))) ++bad;
    }
    return bad;
}
// the same over log-uniform magnitudes (|x|, |y| in 2^-40 .. 2^40, random signs): ratios far beyond what scan geometry produces,
// where the branch thresholds of atanf / atan2f live (|y/x| >= 2^25, < 2^-29, exponent difference > 60 ...)
int64_t ltro_atan2f_selfcheck_wide(uint64_t seed, int64_t n) {
    int64_t bad = 0;
    uint64_t s = seed * 0x9e3779b97f4a7c15ull + 7;
    auto next = [&s]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    auto draw = [&]() {
        const uint64_t r = next();
        const uint32_t e = 127u - 40u + (uint32_t)((r >> 40) % 81u);
        const uint32_t bits = ((uint32_t)(r & 1u) << 31) | (e << 23) | (uint32_t)((r >> 8) & 0x7fffffu);
        return u2f(bits);
    };
    for (int64_t i = 0; i < n; ++i) {
        const float x = draw(), y = draw();
        if (f2u(atan2f(y, x)) != f2u(ref_atan2f(y, x))) ++bad;
    }
    return bad;
}
void ltro_reset_rimg_size(float vfov, float hfov, float alpha, int* rows, int* cols) { resetRimgSize(vfov, hfov, alpha, rows, cols); }
void ltro_pixel_index(const float* xyz, int64_t n, int stride, float vfov, float hfov, int rows, int cols, int* row, int* col, float* range) {
    for (int64_t i = 0; i < n; ++i) {
        const Sph s = cart2sph(xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2]);
        pixelIndex(s, vfov, hfov, rows, cols, &row[i], &col[i]);
        if (range) range[i] = s.r;
    }
}
void ltro_transform(const float* xyzi, int64_t n, const double* T, int order, float* out) {
    Cloud c = toCloud(xyzi, n), o;
    transformPointCloud(c, o, toMat(T), order);
    if (n > 0) std::memcpy(out, o.data(), (size_t)n * sizeof(Pt));
}
void ltro_inverse4x4(const double* in, double* out) { Mat4 r = inverse4x4(toMat(in)); std::memcpy(out, r.m, sizeof(r.m)); }

void ltro_scan2rimg(const float* xyzi, int64_t n, float vfov, float hfov, int rows, int cols, float* rimg) {
    Params p = makeParams(vfov, hfov, nullptr, 0, 1, 0);
    std::vector<float> r;
    scan2RangeImg(toCloud(xyzi, n), p, rows, cols, r);
    std::memcpy(rimg, r.data(), r.size() * 4);
}
void ltro_map2rimg(const float* xyzi, int64_t n, float vfov, float hfov, int rows, int cols, float* rimg, int* ptidx) {
    Params p = makeParams(vfov, hfov, nullptr, 0, 1, 0);
    std::vector<float> r;
    std::vector<int> id;
    map2RangeImg(toCloud(xyzi, n), p, rows, cols, r, id);
    std::memcpy(rimg, r.data(), r.size() * 4);
    std::memcpy(ptidx, id.data(), id.size() * 4);
}

// One remove/revert/ND/PD pass: flags[N] (1 = dynamic). Returns the number of dynamic points.
int64_t ltro_remove_pass(const float* map_xyzi, int64_t N, const float* scans_xyzi, const int64_t* offsets, const double* inv_poses, int K,
                         float vfov, float hfov, const double* lidar2base, int order, int mode, float alpha, float thres, int threads,
                         uint8_t* flags) {
    Params p = makeParams(vfov, hfov, lidar2base, order, threads, 0);
    int rows, cols;
    resetRimgSize(vfov, hfov, alpha, &rows, &cols);
    Cloud map = toCloud(map_xyzi, N);
    std::vector<Cloud> scans((size_t)K);
    std::vector<Mat4> inv((size_t)K);
    for (int k = 0; k < K; ++k) { scans[k] = toCloud(scans_xyzi + 4 * offsets[k], offsets[k + 1] - offsets[k]); inv[k] = toMat(inv_poses + 16 * (size_t)k); }
    const std::vector<int> idx = calcDescrepancyAndParseDynamicPointIdxForEachScan(map, scans, inv, p, mode, rows, cols, thres);
    std::memset(flags, 0, (size_t)N);
    for (int id : idx) flags[id] = 1;
    return (int64_t)idx.size();
}

// Visible points of one keyframe (Session.cpp:353-357): out_xyzi gets the emitted local points, out_idx the map indices.
int64_t ltro_parse_projected(const float* map_xyzi, int64_t N, const double* inv_pose, float vfov, float hfov, const double* lidar2base,
                             int order, float alpha, float* out_xyzi, int* out_idx, int64_t cap) {
    Params p = makeParams(vfov, hfov, lidar2base, order, 1, 0);
    int rows, cols;
    resetRimgSize(vfov, hfov, alpha, &rows, &cols);
    Cloud map = toCloud(map_xyzi, N), local;
    transformGlobalMapToLocal(map, toMat(inv_pose), p.base2lidar, order, local);
    std::vector<float> rimg;
    std::vector<int> ptidx;
    map2RangeImg(local, p, rows, cols, rimg, ptidx);
    int64_t m = 0;
    for (size_t px = 0; px < ptidx.size(); ++px) {
        if (ptidx[px] == 0) continue;
        if (m < cap) {
            if (out_xyzi) std::memcpy(out_xyzi + 4 * m, &local[(size_t)ptidx[px]], sizeof(Pt));
            if (out_idx) out_idx[m] = ptidx[px];
        }
        ++m;
    }
    return m;
}

int64_t ltro_voxel(const float* xyzi, int64_t n, float leaf, float* out, int64_t cap) {
    Cloud c = toCloud(xyzi, n), o;
    if (octreeDownsampling(c, o, leaf) != 0) return -1;
    const int64_t m = (int64_t)o.size();
    if (m <= cap && m > 0) std::memcpy(out, o.data(), (size_t)m * sizeof(Pt));
    return m;
}

// k smallest squared distances (ascending) for each query; out is n*k floats (inf-padded)
void ltro_knn_dists(const float* q_xyzi, int64_t n, const float* t_xyzi, int64_t T, int k, float* out, int brute) {
    if (brute) {
        for (int64_t i = 0; i < n; ++i) {
            std::vector<float> d((size_t)T);
            for (int64_t j = 0; j < T; ++j) d[(size_t)j] = KdTree::dist2(q_xyzi + 4 * i, t_xyzi + 4 * j);
            std::sort(d.begin(), d.end());
            for (int j = 0; j < k; ++j) out[i * k + j] = j < T ? d[(size_t)j] : INFINITY;
        }
        return;
    }
    KdTree tree;
    tree.build(t_xyzi, 4, (int)T);
    for (int64_t i = 0; i < n; ++i) {
        for (int j = 0; j < k; ++j) out[i * k + j] = INFINITY;
        tree.knn(q_xyzi + 4 * i, k, out + i * k);
    }
}

// labels[i] = 1 if point i of the (local) scan is "diff" (Session.cpp:588-600); also returns the two
// re-localised partitions if the output pointers are non-null (capacity n each).
int64_t ltro_knn_partition(const float* scan_xyzi, int64_t n, const double* pose, const double* inv_pose, const float* t_xyzi, int64_t T,
                           const double* lidar2base, int order, int k, float thr, uint8_t* labels, float* coexist_out, float* diff_out) {
    Params p = makeParams(50, 360, lidar2base, order, 1, 0);
    KdTree tree;
    tree.build(t_xyzi, 4, (int)T);
    Cloud co, di;
    std::vector<uint8_t> lab;
    partitionScanByKnn(toCloud(scan_xyzi, n), toMat(pose), toMat(inv_pose), tree, p, k, thr, co, di, &lab);
    if (labels && n > 0) std::memcpy(labels, lab.data(), (size_t)n);
    if (coexist_out && !co.empty()) std::memcpy(coexist_out, co.data(), co.size() * sizeof(Pt));
    if (diff_out && !di.empty()) std::memcpy(diff_out, di.data(), di.size() * sizeof(Pt));
    return (int64_t)di.size();
}

// ---------------- pipeline level ----------------
void* ltro_create() { return new Removerter(); }
void ltro_destroy(void* h) { delete (Removerter*)h; }

void ltro_set_params(void* h, float vfov, float hfov, const double* lidar2base, int order, int num_knn, float knn_thr, float voxel,
                     int threads, int faithful, int omp_cores, int do_high_dyn_knn) {
    Removerter* R = (Removerter*)h;
    R->P = makeParams(vfov, hfov, lidar2base, order, threads, faithful);
    R->P.num_knn = num_knn; R->P.knn_thr = knn_thr; R->P.downsample_voxel = voxel; R->P.omp_cores = omp_cores;
    R->do_high_dyn_knn = do_high_dyn_knn != 0;
}
void ltro_set_schedule(void* h, const int* ops, const float* res, int n) {
    Removerter* R = (Removerter*)h;
    R->hd_schedule.clear();
    for (int i = 0; i < n; ++i) R->hd_schedule.push_back({ops[i], res[i]});
}
void ltro_load_session(void* h, int sess, const float* xyzi, const int64_t* offsets, const double* poses, const double* inv_poses, int K) {
    Removerter* R = (Removerter*)h;
    Session& S = sess == 0 ? R->central_sess_ : R->query_sess_;
    S.keyframe_scans_.assign((size_t)K, Cloud());
    S.keyframe_poses_.resize((size_t)K);
    S.keyframe_inverse_poses_.resize((size_t)K);
    for (int k = 0; k < K; ++k) {
        S.keyframe_scans_[k] = toCloud(xyzi + 4 * offsets[k], offsets[k + 1] - offsets[k]);
        S.keyframe_poses_[k] = toMat(poses + 16 * (size_t)k);
        S.keyframe_inverse_poses_[k] = inv_poses ? toMat(inv_poses + 16 * (size_t)k) : inverse4x4(S.keyframe_poses_[k]);
    }
}
void ltro_set_map(void* h, int sess, const char* name, const float* xyzi, int64_t n) {
    Cloud* c = findCloud((Removerter*)h, name, sess, -1);
    if (c) *c = toCloud(xyzi, n);
}
int ltro_run(void* h, int mask) {
    Removerter* R = (Removerter*)h;
    if (mask & 1) R->runStep0();
    if (mask & 2) R->runStep12();
    if (mask & 4) R->runStep3();
    return 0;
}
// finer-grained stage entry points (mirror run()'s call graph) for stage-by-stage parity tests
int ltro_stage(void* h, const char* stage) {
    Removerter* R = (Removerter*)h;
    const std::string s(stage);
    if (s == "precleaningKeyframes") R->precleaningKeyframes(2.5f);
    else if (s == "makeGlobalMap") R->makeGlobalMap();
    else if (s == "removeHighDynamicPoints") R->removeHighDynamicPoints();
    else if (s == "parseStaticScansViaProjection") R->parseStaticScansViaProjection();
    else if (s == "detectLowDynamicPoints") R->detectLowDynamicPoints();
    else if (s == "updateCurrentMap") R->updateCurrentMap();
    else if (s == "parseUpdatedStaticScansViaProjection") R->parseUpdatedStaticScansViaProjection();
    else if (s == "parseLDScansViaProjection") R->parseLDScansViaProjection();
    else if (s == "updateScansScanwise") R->updateScansScanwise();
    else return -1;
    return 0;
}
int64_t ltro_cloud_size(void* h, const char* name, int sess, int kf) {
    Cloud* c = findCloud((Removerter*)h, name, sess, kf);
    return c ? (int64_t)c->size() : -1;
}
int64_t ltro_cloud_copy(void* h, const char* name, int sess, int kf, float* out, int64_t cap) {
    Cloud* c = findCloud((Removerter*)h, name, sess, kf);
    if (!c) return -1;
    const int64_t n = (int64_t)c->size();
    if (n <= cap && n > 0) std::memcpy(out, c->data(), (size_t)n * sizeof(Pt));
    return n;
}
int ltro_num_keyframes(void* h, int sess) {
    Removerter* R = (Removerter*)h;
    return (int)(sess == 0 ? R->central_sess_ : R->query_sess_).keyframe_scans_.size();
}
int ltro_log_count(void* h) { return (int)((Removerter*)h)->log.size(); }
int ltro_log_get(void* h, int i, char* what, int cap, int64_t* vals /*4*/) {
    Removerter* R = (Removerter*)h;
    if (i < 0 || i >= (int)R->log.size()) return -1;
    const PassLog& l = R->log[(size_t)i];
    std::snprintf(what, (size_t)cap, "%s", l.what.c_str());
    vals[0] = l.n_map; vals[1] = l.n_dynamic; vals[2] = l.n_static_after; vals[3] = l.n_dynamic_after;
    return 0;
}
double ltro_timing(void* h, const char* key) {
    Removerter* R = (Removerter*)h;
    auto it = R->timing.find(key);
    return it == R->timing.end() ? 0.0 : it->second;
}

}  // extern "C"
