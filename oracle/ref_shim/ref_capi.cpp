// ORACLE / REFERENCE SHIM (test infrastructure, not product code).
// Shim state, PCD file I/O for the pcl::io stand-ins, and a flat C view of the reference's own classes
// (ltremovert::Removerter / Session, compiled unmodified from /root/reference) for the ctypes tests.
// Built by oracle/Makefile into oracle/_ref/libltremovert_ref.so (git-ignored; travels to the GPU box).
#include <algorithm>
#include <array>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <ctime>
#include <deque>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <mutex>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "ltr_shim_core.h"

// the sessions are private members of Removerter (Removerter.h:12-16); the tests need to read them.  Access specifiers
// do not change the object layout, so this translation unit stays ABI-compatible with the unmodified reference objects.
#define private public
#include "removert/Removerter.h"
#undef private

namespace ltr_shim {
std::map<std::string, ParamValue>& params() { static std::map<std::string, ParamValue> p; return p; }
int& transform_order() { static int o = 0; return o; }
int& verbose() { static int v = 0; return v; }
std::vector<SavedCloud>& saved() { static std::vector<SavedCloud> s; return s; }
int& write_files() { static int w = 1; return w; }
void fatal(const char* what) { std::fprintf(stderr, "[ref_shim] %s\n", what); std::abort(); }
}  // namespace ltr_shim

namespace pcl { namespace io {

// PCD v0.7 reader for the fields the pipeline uses (x y z intensity as 4-byte floats; extra fields are skipped),
// DATA ascii | binary.  binary_compressed is not modelled.
int load_pcd_xyzi(const std::string& path, std::vector<float>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) return -1;
    std::vector<std::string> fields; std::vector<int> sizes, counts; std::vector<char> types;
    long npts = -1; std::string data, line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line); std::string key; ss >> key;
        if (key == "FIELDS") { std::string s; while (ss >> s) fields.push_back(s); }
        else if (key == "SIZE") { int s; while (ss >> s) sizes.push_back(s); }
        else if (key == "TYPE") { char c; while (ss >> c) types.push_back(c); }
        else if (key == "COUNT") { int c; while (ss >> c) counts.push_back(c); }
        else if (key == "POINTS") ss >> npts;
        else if (key == "DATA") { ss >> data; break; }
    }
    if (npts < 0 || fields.empty() || sizes.size() != fields.size()) return -1;
    if (counts.empty()) counts.assign(fields.size(), 1);
    int off[4] = {-1, -1, -1, -1}, col[4] = {-1, -1, -1, -1}, stride = 0, ncol = 0;
    const char* want[4] = {"x", "y", "z", "intensity"};
    for (std::size_t i = 0; i < fields.size(); ++i) {
        for (int k = 0; k < 4; ++k) if (fields[i] == want[k]) { if (sizes[i] != 4 || (!types.empty() && types[i] != 'F')) return -1; off[k] = stride; col[k] = ncol; }
        stride += sizes[i] * counts[i]; ncol += counts[i];
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0) return -1;
    out.assign((std::size_t)npts * 4, 0.0f);
    if (data == "binary") {
        std::vector<char> buf((std::size_t)npts * stride);
        f.read(buf.data(), (std::streamsize)buf.size());
        if ((std::size_t)f.gcount() != buf.size()) return -1;
        for (long i = 0; i < npts; ++i) for (int k = 0; k < 4; ++k) if (off[k] >= 0) std::memcpy(&out[(std::size_t)i * 4 + k], &buf[(std::size_t)i * stride + off[k]], 4);
    } else if (data == "ascii") {
        for (long i = 0; i < npts; ++i) {
            if (!std::getline(f, line)) return -1;
            std::istringstream ss(line); std::string tok;
            for (int c = 0; ss >> tok; ++c) for (int k = 0; k < 4; ++k) if (col[k] == c) out[(std::size_t)i * 4 + k] = std::strtof(tok.c_str(), nullptr);
        }
    } else return -1;
    return 0;
}

int save_pcd_xyzi(const std::string& path, const float* xyzi, std::size_t n) {
    ltr_shim::SavedCloud s; s.path = path; s.xyzi.assign(xyzi, xyzi + n * 4);
    ltr_shim::saved().push_back(std::move(s));
    if (!ltr_shim::write_files()) return 0;
    std::ofstream f(path, std::ios::binary);
    if (!f.good()) return -1;
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
      << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    if (n) f.write((const char*)xyzi, (std::streamsize)(n * 16));
    return f.good() ? 0 : -1;
}
}}  // namespace pcl::io

using ltremovert::Removerter;
using ltremovert::Session;
typedef pcl::PointCloud<PointType> PC;

static Session& sess_of(Removerter* R, int s) { return s == 0 ? R->central_sess_ : R->query_sess_; }

static PC::Ptr* cloud_member(Session& S, const std::string& n) {
#define M(x) if (n == #x) return &S.x;
    M(map_global_orig_) M(map_global_curr_) M(map_local_curr_) M(map_global_curr_static_) M(map_global_curr_dynamic_)
    M(map_global_updated_) M(map_global_updated_strong_) M(map_global_nd_) M(map_global_nd_strong_) M(map_global_nd_weak_)
    M(map_global_pd_) M(map_global_pd_orig_) M(map_global_pd_strong_) M(map_global_pd_weak_) M(target_map_down_for_knn_)
#undef M
    return nullptr;
}
static std::vector<PC::Ptr>* scans_member(Session& S, const std::string& n) {
#define M(x) if (n == #x) return &S.x;
    M(keyframe_scans_) M(keyframe_scans_static_) M(keyframe_scans_static_projected_) M(keyframe_scans_dynamic_)
    M(scans_knn_coexist_) M(scans_knn_diff_) M(keyframe_scans_updated_) M(keyframe_scans_updated_strong_)
    M(keyframe_scans_pd_) M(keyframe_scans_strong_pd_) M(keyframe_scans_strong_nd_) M(keyframe_scans_weak_nd_)
#undef M
    return nullptr;
}
static PC::Ptr make_cloud(const float* xyzi, int64_t n) {
    PC::Ptr c(new PC());
    c->points.resize((size_t)n);
    if (n) std::memcpy(c->points.data(), xyzi, (size_t)n * 16);
    c->width = (uint32_t)n; c->height = 1;
    return c;
}
static int64_t copy_out(const PC& c, float* out, int64_t cap) {
    const int64_t n = (int64_t)c.points.size();
    if (out && cap >= n && n) std::memcpy(out, c.points.data(), (size_t)n * 16);
    return n;
}

extern "C" {

// ---- parameters (what the launch file's yaml would put on the ROS parameter server) ----
void ref_params_clear() { ltr_shim::params().clear(); }
void ref_param_num(const char* name, const double* v, int n) { ltr_shim::ParamValue p; p.nums.assign(v, v + n); ltr_shim::params()[name] = p; }
void ref_param_str(const char* name, const char* s) { ltr_shim::ParamValue p; p.str = s; ltr_shim::params()[name] = p; }
void ref_set_transform_order(int o) { ltr_shim::transform_order() = o; }
void ref_set_verbose(int v) { ltr_shim::verbose() = v; }
void ref_set_write_files(int w) { ltr_shim::write_files() = w; }

// ---- free functions of utility.cpp ----
void ref_cart2sph(const float* xyz, int64_t n, float* az_el_r) {
    for (int64_t i = 0; i < n; ++i) {
        PointType p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        const SphericalPoint s = cart2sph(p);
        az_el_r[3 * i] = s.az; az_el_r[3 * i + 1] = s.el; az_el_r[3 * i + 2] = s.r;
    }
}
void ref_rad2deg(const float* in, int64_t n, float* out) { for (int64_t i = 0; i < n; ++i) out[i] = rad2deg(in[i]); }
void ref_reset_rimg_size(float vfov, float hfov, float alpha, int* rows, int* cols) {
    const std::pair<int, int> s = resetRimgSize({vfov, hfov}, alpha);
    *rows = s.first; *cols = s.second;
}
void ref_map2rimg(const float* xyzi, int64_t n, float vfov, float hfov, int rows, int cols, float* rimg, int* ptidx) {
    auto [r, p] = map2RangeImg(make_cloud(xyzi, n), {vfov, hfov}, {rows, cols});
    std::memcpy(rimg, r.buf->data(), (size_t)rows * cols * 4);
    std::memcpy(ptidx, p.buf->data(), (size_t)rows * cols * 4);
}
int64_t ref_parse_projected(const float* xyzi, int64_t n, float vfov, float hfov, int rows, int cols, float* out, int64_t cap) {
    return copy_out(*parseProjectedPoints(make_cloud(xyzi, n), {vfov, hfov}, {rows, cols}), out, cap);
}
void ref_transform_global_to_local(const float* xyzi, int64_t n, const double* inv_pose, const double* base2lidar, float* out) {
    Eigen::Matrix4d a, b; std::memcpy(a.m, inv_pose, sizeof(a.m)); std::memcpy(b.m, base2lidar, sizeof(b.m));
    PC::Ptr loc(new PC());
    transformGlobalMapToLocal(make_cloud(xyzi, n), a, b, loc);
    copy_out(*loc, out, n);
}
int64_t ref_octree_downsampling(const float* xyzi, int64_t n, float leaf, float* out, int64_t cap) {
    PC::Ptr dst(new PC());
    octreeDownsampling(make_cloud(xyzi, n), dst, leaf);
    return copy_out(*dst, out, cap);
}
int ref_linspace_int(int a, int b, int n, int* out) { const std::vector<int> v = linspace<int>(a, b, (size_t)n); std::copy(v.begin(), v.end(), out); return (int)v.size(); }
void ref_inverse4x4(const double* in, double* out) { Eigen::Matrix4d a; std::memcpy(a.m, in, sizeof(a.m)); const Eigen::Matrix4d r = a.inverse(); std::memcpy(out, r.m, sizeof(r.m)); }

// ---- Removerter ----
void* ref_create() { return new Removerter(); }
void ref_destroy(void* h) { delete (Removerter*)h; }
void ref_run(void* h) { ((Removerter*)h)->run(); }

// member functions of Removerter that run() (Removerter.cpp:1654-1677) and selfRemovert (:1378-1393) are made of
int ref_stage(void* h, const char* stage) {
    Removerter* R = (Removerter*)h;
    const std::string s = stage;
    if (s == "loadSessionInfo") R->loadSessionInfo();
    else if (s == "parseKeyframes") R->parseKeyframes();
    else if (s == "loadKeyframes") R->loadKeyframes();
    else if (s == "precleaningKeyframes") R->precleaningKeyframes(2.5);
    else if (s == "makeGlobalMap") R->makeGlobalMap();
    else if (s == "removeHighDynamicPoints") R->removeHighDynamicPoints();
    else if (s == "parseStaticScansViaProjection") R->parseStaticScansViaProjection();
    else if (s == "detectLowDynamicPoints") R->detectLowDynamicPoints();
    else if (s == "updateCurrentMap") R->updateCurrentMap();
    else if (s == "parseUpdatedStaticScansViaProjection") R->parseUpdatedStaticScansViaProjection();
    else if (s == "parseLDScansViaProjection") R->parseLDScansViaProjection();
    else if (s == "updateScansScanwise") R->updateScansScanwise();
    else if (s == "saveAllTypeOfScans") R->saveAllTypeOfScans();
    else return -1;
    return 0;
}
// op 0 removeOnce, 1 revertOnce, 2 resetCurrrentMapAsDynamic, 3 resetCurrrentMapAsStatic, 4 iremoveOnceForND, 5 removeOnceForPD
int ref_pass(void* h, int op, int target, int source, float res) {
    Removerter* R = (Removerter*)h;
    Session& t = sess_of(R, target); Session& s = sess_of(R, source);
    switch (op) {
        case 0: R->removeOnce(t, s, res); break;
        case 1: R->revertOnce(t, s, res); break;
        case 2: R->resetCurrrentMapAsDynamic(t); break;
        case 3: R->resetCurrrentMapAsStatic(t); break;
        case 4: R->iremoveOnceForND(t, s, res); break;
        case 5: R->removeOnceForPD(t, s, res); break;
        default: return -1;
    }
    return 0;
}
// Step 1 with an explicit remove / revert schedule (op 0 = removeOnce(res); op 1 = resetCurrrentMapAsDynamic, revertOnce(res),
// resetCurrrentMapAsStatic -- the building blocks of selfRemovert, Removerter.cpp:1378-1393) applied to each session, followed by the
// rest of removeHighDynamicPoints (Removerter.cpp:1591-1604).  Every call is one of the reference's own member functions.
void ref_high_dyn_with_schedule(void* h, const int* ops, const float* res, int n) {
    Removerter* R = (Removerter*)h;
    for (int s = 0; s < 2; ++s) {
        Session& S = sess_of(R, s);
        for (int i = 0; i < n; ++i) {
            if (ops[i] == 0) R->removeOnce(S, S, res[i]);
            else { R->resetCurrrentMapAsDynamic(S); R->revertOnce(S, S, res[i]); R->resetCurrrentMapAsStatic(S); }
        }
    }
    Session& C = R->central_sess_; Session& Q = R->query_sess_;
    C.extractHighDynPointsViaKnnDiff(C.map_global_curr_static_);
    Q.extractHighDynPointsViaKnnDiff(Q.map_global_curr_static_);
    auto mc = mergeScansWithinGlobalCoordUtil(C.keyframe_scans_dynamic_, C.keyframe_poses_, C.kSE3MatExtrinsicLiDARtoPoseBase);
    auto mq = mergeScansWithinGlobalCoordUtil(Q.keyframe_scans_dynamic_, Q.keyframe_poses_, Q.kSE3MatExtrinsicLiDARtoPoseBase);
    octreeDownsampling(mc, mc, 0.05);
    octreeDownsampling(mq, mq, 0.05);
    pcl::io::savePCDFileBinary(R->save_pcd_directory_ + "central_sess_high_dyn.pcd", *mc);
    pcl::io::savePCDFileBinary(R->save_pcd_directory_ + "query_sess_high_dyn.pcd", *mq);
}
void ref_self_removert(void* h, int sess, int repeat) { Removerter* R = (Removerter*)h; R->selfRemovert(sess_of(R, sess), repeat); }

int64_t ref_scan2rimg(void* h, const float* xyzi, int64_t n, int rows, int cols, float* rimg) {
    Removerter* R = (Removerter*)h;
    const cv::Mat m = R->scan2RangeImg(make_cloud(xyzi, n), R->kFOV, {rows, cols});
    std::memcpy(rimg, m.buf->data(), (size_t)rows * cols * 4);
    return (int64_t)rows * cols;
}
// calcDescrepancyAndParseDynamicPointIdxForEachScan{,ForND,ForPD}: mode 0 / 1 / 2 -> sorted unique indices
int64_t ref_dynamic_idx(void* h, int mode, int target, int source, int rows, int cols, int* out, int64_t cap) {
    Removerter* R = (Removerter*)h;
    Session& t = sess_of(R, target); Session& s = sess_of(R, source);
    std::vector<int> v;
    if (mode == 0) v = R->calcDescrepancyAndParseDynamicPointIdxForEachScan(t, s, {rows, cols});
    else if (mode == 1) v = R->calcDescrepancyAndParseDynamicPointIdxForEachScanForND(t, s, {rows, cols});
    else v = R->calcDescrepancyAndParseDynamicPointIdxForEachScanForPD(t, s, {rows, cols});
    if (out && cap >= (int64_t)v.size()) std::copy(v.begin(), v.end(), out);
    return (int64_t)v.size();
}
int64_t ref_static_idx(void* h, const int* dyn, int64_t ndyn, int num_all, int* out, int64_t cap) {
    Removerter* R = (Removerter*)h;
    const std::vector<int> d(dyn, dyn + ndyn);
    const std::vector<int> v = R->getStaticIdxFromDynamicIdx(d, num_all);
    if (out && cap >= (int64_t)v.size()) std::copy(v.begin(), v.end(), out);
    return (int64_t)v.size();
}

// ---- state access ----
int ref_num_keyframes(void* h, int sess) { return (int)sess_of((Removerter*)h, sess).keyframe_poses_.size(); }
int ref_num_scans(void* h, int sess) { return (int)sess_of((Removerter*)h, sess).scan_poses_.size(); }
int ref_keyframe_name(void* h, int sess, int k, char* out, int cap) {
    const std::string& s = sess_of((Removerter*)h, sess).keyframe_names_.at((size_t)k);
    std::snprintf(out, (size_t)cap, "%s", s.c_str());
    return (int)s.size();
}
void ref_keyframe_pose(void* h, int sess, int k, double* pose, double* inv) {
    Session& S = sess_of((Removerter*)h, sess);
    std::memcpy(pose, S.keyframe_poses_.at((size_t)k).m, 128);
    std::memcpy(inv, S.keyframe_inverse_poses_.at((size_t)k).m, 128);
}
void ref_extrinsics(void* h, double* lidar2base, double* base2lidar) {
    Removerter* R = (Removerter*)h;
    std::memcpy(lidar2base, R->kSE3MatExtrinsicLiDARtoPoseBase.m, 128);
    std::memcpy(base2lidar, R->kSE3MatExtrinsicPoseBasetoLiDAR.m, 128);
}
int64_t ref_cloud(void* h, int sess, const char* name, float* out, int64_t cap) {
    PC::Ptr* p = cloud_member(sess_of((Removerter*)h, sess), name);
    if (!p || !*p) return -1;
    return copy_out(**p, out, cap);
}
int ref_set_cloud(void* h, int sess, const char* name, const float* xyzi, int64_t n) {
    PC::Ptr* p = cloud_member(sess_of((Removerter*)h, sess), name);
    if (!p) return -1;
    **p = *make_cloud(xyzi, n);
    return 0;
}
int ref_scans_count(void* h, int sess, const char* name) {
    std::vector<PC::Ptr>* v = scans_member(sess_of((Removerter*)h, sess), name);
    return v ? (int)v->size() : -1;
}
int64_t ref_scan(void* h, int sess, const char* name, int k, float* out, int64_t cap) {
    std::vector<PC::Ptr>* v = scans_member(sess_of((Removerter*)h, sess), name);
    if (!v || k < 0 || k >= (int)v->size()) return -1;
    return copy_out(*v->at((size_t)k), out, cap);
}
int ref_set_scans(void* h, int sess, const char* name, const float* xyzi, const int64_t* offsets, int K) {
    std::vector<PC::Ptr>* v = scans_member(sess_of((Removerter*)h, sess), name);
    if (!v) return -1;
    v->clear();
    for (int k = 0; k < K; ++k) v->push_back(make_cloud(xyzi + 4 * offsets[k], offsets[k + 1] - offsets[k]));
    return 0;
}
// load keyframes from memory instead of files: poses + already down-sampled scans (what loadKeyframes leaves behind)
void ref_load_session_mem(void* h, int sess, const float* xyzi, const int64_t* offsets, const double* poses, int K) {
    Session& S = sess_of((Removerter*)h, sess);
    S.sess_type_ = sess == 0 ? "Central" : "Query";
    S.keyframe_poses_.clear(); S.keyframe_inverse_poses_.clear(); S.keyframe_scans_.clear(); S.keyframe_names_.clear();
    for (int k = 0; k < K; ++k) {
        Eigen::Matrix4d P; std::memcpy(P.m, poses + 16 * k, 128);
        S.keyframe_poses_.push_back(P);
        S.keyframe_inverse_poses_.push_back(P.inverse());
        S.keyframe_scans_.push_back(make_cloud(xyzi + 4 * offsets[k], offsets[k + 1] - offsets[k]));
        char nm[32]; std::snprintf(nm, sizeof(nm), "%06d.pcd", k);
        S.keyframe_names_.push_back(nm);
    }
}

// Session::extractLowDynPointsViaKnnDiff (Session.cpp:393-427; reads keyframe_scans_static_projected_) / extractHighDynPointsViaKnnDiff
// (:487-504; reads keyframe_scans_) of session `sess` against a target map given by value; k / thr are the per-object copies of
// removert/num_nn_points_within and dist_nn_points_within (every Session is its own RosParamServer).
int ref_extract_knn_diff(void* h, int sess, int low, const float* target_xyzi, int64_t n, int k, float thr) {
    Session& S = sess_of((Removerter*)h, sess);
    S.kNumKnnPointsToCompare = k;
    S.kScanKnnAndMapKnnAvgDiffThreshold = thr;
    PC::Ptr t = make_cloud(target_xyzi, n);
    if (low) S.extractLowDynPointsViaKnnDiff(t); else S.extractHighDynPointsViaKnnDiff(t);
    return 0;
}

// ---- timing of the reference's own per-keyframe loops (bench.py --impl reference; see its docstring) ----
// Each call runs ONE loop of the reference over the keyframes currently loaded in the session(s), on whatever maps the caller put
// into the members, and returns the seconds spent inside (steady clock around the reference's own member function).
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// mode 0 / 1 / 2 = calcDescrepancyAndParseDynamicPointIdxForEachScan / ...ForND / ...ForPD (Removerter.cpp:542-593, 485-540, 429-482)
double ref_time_dynamic_idx(void* h, int mode, int target, int source, int rows, int cols, int64_t* n_dynamic) {
    Removerter* R = (Removerter*)h;
    Session& t = sess_of(R, target); Session& s = sess_of(R, source);
    const double t0 = now_s();
    std::vector<int> v;
    if (mode == 0) v = R->calcDescrepancyAndParseDynamicPointIdxForEachScan(t, s, {rows, cols});
    else if (mode == 1) v = R->calcDescrepancyAndParseDynamicPointIdxForEachScanForND(t, s, {rows, cols});
    else v = R->calcDescrepancyAndParseDynamicPointIdxForEachScanForPD(t, s, {rows, cols});
    const double dt = now_s() - t0;
    if (n_dynamic) *n_dynamic = (int64_t)v.size();
    return dt;
}
// Session::parseStaticScansViaProjection (Session.cpp:305-308 -> 348-360) of one session
double ref_time_parse_static(void* h, int sess) {
    Session& S = sess_of((Removerter*)h, sess);
    const double t0 = now_s();
    S.parseStaticScansViaProjection();
    return now_s() - t0;
}
// kd-tree construction of Session::extract{Low,High}DynPointsViaKnnDiff (Session.cpp:403 / :489), a per-call fixed cost
double ref_knn_set_target(void* h, int sess, const float* xyzi, int64_t n, int k, float thr) {
    Session& S = sess_of((Removerter*)h, sess);
    S.kNumKnnPointsToCompare = k;
    S.kScanKnnAndMapKnnAvgDiffThreshold = thr;
    const double t0 = now_s();
    S.kdtree_target_map_global_->setInputCloud(make_cloud(xyzi, n));
    return now_s() - t0;
}
// the keyframe loop of the same two functions (Session.cpp:408-414 / :491-497, OpenMP over keyframes as in the reference)
double ref_time_knn_queries(void* h, int sess, int low, int omp_cores, int64_t* n_diff) {
    Session& S = sess_of((Removerter*)h, sess);
    const int K = low ? (int)S.keyframe_scans_static_projected_.size() : (int)S.keyframe_scans_.size();
    std::vector<PC::Ptr> co((size_t)K), di((size_t)K);
    const double t0 = now_s();
#pragma omp parallel for num_threads(omp_cores)
    for (int k = 0; k < K; ++k) {
        auto pr = low ? S.partitionLowDynamicPointsOfScanByKnn(k) : S.partitionHighDynamicPointsOfScanByKnn(k);
        co[(size_t)k] = pr.first; di[(size_t)k] = pr.second;
    }
    const double dt = now_s() - t0;
    int64_t nd = 0;
    for (int k = 0; k < K; ++k) nd += (int64_t)di[(size_t)k]->points.size();
    if (n_diff) *n_diff = nd;
    if (low) { S.scans_knn_coexist_ = co; S.scans_knn_diff_ = di; } else S.keyframe_scans_dynamic_ = di;
    return dt;
}
// mergeScansWithinGlobalCoordUtil (utility.cpp:170-192) over one per-keyframe member of the session
double ref_time_merge(void* h, int sess, const char* name, int64_t* n_out) {
    Session& S = sess_of((Removerter*)h, sess);
    std::vector<PC::Ptr>* v = scans_member(S, name);
    if (!v) return -1.0;
    const double t0 = now_s();
    auto m = mergeScansWithinGlobalCoordUtil(*v, S.keyframe_poses_, S.kSE3MatExtrinsicLiDARtoPoseBase);
    const double dt = now_s() - t0;
    if (n_out) *n_out = (int64_t)m->points.size();
    return dt;
}

// ---- what the run wrote through pcl::io::savePCDFileBinary ----
void ref_saved_clear() { ltr_shim::saved().clear(); }
int ref_saved_count() { return (int)ltr_shim::saved().size(); }
int64_t ref_saved_get(int i, char* path, int cap, float* out, int64_t out_cap) {
    const ltr_shim::SavedCloud& s = ltr_shim::saved().at((size_t)i);
    if (path) std::snprintf(path, (size_t)cap, "%s", s.path.c_str());
    const int64_t n = (int64_t)s.xyzi.size() / 4;
    if (out && out_cap >= n && n) std::memcpy(out, s.xyzi.data(), (size_t)n * 16);
    return n;
}

}  // extern "C"
