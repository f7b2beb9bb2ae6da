// Host orchestration of the LT-removert / LT-map hot path on top of the device C-ABI (include/ltr_b200.h).
// Statement order follows ltremovert/src/Removerter.cpp and ltremovert/src/Session.cpp (cited per function);
// see removerter.h.  No arithmetic on points happens here.
#include "removerter.h"
#include "io.h"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <cstdio>

namespace ltremovert_b200 {

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define CK(expr)                                                          \
    do {                                                                  \
        const int rc__ = (expr);                                          \
        if (rc__ != LTR_OK) {                                             \
            if (err.empty()) err = ctx ? ltr_last_error(ctx) : "no context"; \
            return rc__;                                                  \
        }                                                                 \
    } while (0)

struct StageTimer {
    Removerter& R; std::string key; double t0;
    StageTimer(Removerter& r, const char* k) : R(r), key(k) { if (R.ctx) ltr_synchronize(R.ctx); t0 = now_s(); }
    ~StageTimer() { if (R.ctx) ltr_synchronize(R.ctx); R.timing[key] += now_s() - t0; }
};

std::map<std::string, ltr_cloud*> Session::cloud_names() {
    return {{"map_global_orig_", &map_global_orig_}, {"map_global_curr_", &map_global_curr_},
            {"map_global_curr_static_", &map_global_curr_static_}, {"map_global_curr_dynamic_", &map_global_curr_dynamic_},
            {"map_global_updated_", &map_global_updated_}, {"map_global_updated_strong_", &map_global_updated_strong_},
            {"map_global_nd_", &map_global_nd_}, {"map_global_nd_strong_", &map_global_nd_strong_}, {"map_global_nd_weak_", &map_global_nd_weak_},
            {"map_global_pd_", &map_global_pd_}, {"map_global_pd_orig_", &map_global_pd_orig_},
            {"map_global_pd_strong_", &map_global_pd_strong_}, {"map_global_pd_weak_", &map_global_pd_weak_}};
}
std::map<std::string, ltr_scanset*> Session::scanset_names() {
    return {{"keyframe_scans_", &keyframe_scans_}, {"keyframe_scans_static_projected_", &keyframe_scans_static_projected_},
            {"keyframe_scans_dynamic_", &keyframe_scans_dynamic_}, {"scans_knn_coexist_", &scans_knn_coexist_},
            {"scans_knn_diff_", &scans_knn_diff_}, {"keyframe_scans_updated_", &keyframe_scans_updated_},
            {"keyframe_scans_updated_strong_", &keyframe_scans_updated_strong_}, {"keyframe_scans_pd_", &keyframe_scans_pd_},
            {"keyframe_scans_strong_pd_", &keyframe_scans_strong_pd_}, {"keyframe_scans_strong_nd_", &keyframe_scans_strong_nd_},
            {"keyframe_scans_weak_nd_", &keyframe_scans_weak_nd_}};
}

Removerter::Removerter(const ltrh_params& p) : P(p) {
    central_sess_.sess_type_ = "Central"; central_sess_.id = 0;
    query_sess_.sess_type_ = "Query"; query_sess_.id = 1;
    comm.rank = 0; comm.world = 1;
}

Removerter::~Removerter() {
    ltr_pinned_free(pin_in_);
    ltr_pinned_free(pin_out_);
    if (ctx) ltr_destroy(ctx);
}

int Removerter::fail(int code, const std::string& msg) { err = msg; return code; }

// general 4x4 inverse by cofactors (io.cpp `inverse`, the one implementation in this library); stands in for Eigen's
// kSE3MatExtrinsicLiDARtoPoseBase.inverse() (RosParamServer.cpp:30) and Session.cpp:110
static void invert4x4(const double* m, double* out) {
    Mat4 a;
    std::copy(m, m + 16, a.begin());
    const Mat4 r = inverse(a);
    std::copy(r.begin(), r.end(), out);
}

void invert4x4_public(const double* m, double* out) { invert4x4(m, out); }

int Removerter::init() {
    ltr_config cfg;
    ltr_config_default(&cfg);
    cfg.device = P.device;
    cfg.vfov_deg = P.sequence_vfov;
    cfg.hfov_deg = P.sequence_hfov;
    std::memcpy(cfg.lidar2base, P.ExtrinsicLiDARtoPoseBase, sizeof(cfg.lidar2base));
    bool ident = true;
    for (int i = 0; i < 16; ++i) if (cfg.lidar2base[i] != ((i % 5 == 0) ? 1.0 : 0.0)) ident = false;
    if (ident) std::memcpy(cfg.base2lidar, cfg.lidar2base, sizeof(cfg.base2lidar));
    else invert4x4(cfg.lidar2base, cfg.base2lidar);
    cfg.transform_order = P.transform_order;
    cfg.keyframe_batch = P.keyframe_batch;
    cfg.fast_path = P.fast_path;
    const int rc = ltr_create(&ctx, &cfg);
    if (rc != LTR_OK) err = ltr_last_error(nullptr);
    if (const char* e = std::getenv("LTR_DIST_VOXEL_MIN")) dist_voxel_min = std::atoll(e);   // tests force the distributed path on small clouds
    return rc;
}

int Removerter::set(ltr_cloud* slot, ltr_cloud v) {
    if (*slot >= 0 && *slot != v) CK(ltr_cloud_free(ctx, *slot));
    *slot = v;
    return LTR_OK;
}
int Removerter::set(ltr_scanset* slot, ltr_scanset v, bool) {
    if (*slot >= 0 && *slot != v) CK(ltr_scanset_free(ctx, *slot));
    *slot = v;
    return LTR_OK;
}
int Removerter::assign(ltr_cloud* dst, ltr_cloud src) {
    ltr_cloud c;
    if (src < 0) { CK(ltr_cloud_alloc(ctx, 0, &c)); }
    else CK(ltr_cloud_copy(ctx, src, &c));
    return set(dst, c);
}
int Removerter::append(ltr_cloud* dst, ltr_cloud src) {
    if (*dst < 0) return assign(dst, src);
    if (src < 0) return LTR_OK;
    ltr_cloud c;
    CK(ltr_cloud_concat(ctx, *dst, src, &c));
    return set(dst, c);
}
int Removerter::save(const std::string& name, ltr_cloud c) {
    ltr_cloud copy;
    CK(ltr_cloud_copy(ctx, c, &copy));
    auto it = saved.find(name);
    if (it != saved.end()) { CK(ltr_cloud_free(ctx, it->second)); it->second = copy; }
    else saved[name] = copy;
    return LTR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Parallel decomposition.  One process per GPU.
//   * Keyframe sharding: the ranks that own a session hold contiguous keyframe blocks of it (scans + poses) in rank order; every
//     per-keyframe loop of the reference runs on the owner.  Maps are replicated on those ranks; after a pass the per-rank flag
//     arrays are OR-ed (reduce_flags) and every rank applies the same partition / voxelisation, so replicas stay bit-identical.
//     Clouds merged over all keyframes are concatenated in rank == keyframe order (gather_clouds).
//   * Session split (world even, split_sessions): ranks [0, world/2) own the central session, the others the query session.
//     Step 1 of the two sessions (Removerter.cpp:1584 / 1587) and the two halves of Step 2 are independent, so the two groups
//     run them concurrently and the replicated per-pass work (partition, voxelisation) is done once per group, not once per
//     rank.  A map crosses groups exactly where the reference reads the other session's member: rank r hands it to rank
//     r + world/2 (exchange_with_partner), one peer-to-peer transfer per rank pair.
// Transport: the native NCCL entry points of libltr_b200.so (ltr_nccl_*), on the context's stream, no host code in the loop.
// The legacy ltr_comm hooks remain for host-memory tests of the exchange logic (no session split).
// ---------------------------------------------------------------------------------------------------------------
int Removerter::comm_init_nccl(const uint8_t* id128, int rank_, int world_, int split_sessions) {
    if (world_ < 1 || rank_ < 0 || rank_ >= world_) return fail(LTR_ERR_INVALID, "bad rank / world");
    rank = rank_; world = world_;
    split = split_sessions && world >= 2 && world % 2 == 0;
    CK(ltr_nccl_init(ctx, id128, rank, world, &nccl_world));
    if (split) {
        int g;
        CK(ltr_nccl_split(ctx, nccl_world, rank < world / 2 ? 0 : 1, rank, &g));
        nccl_group[0] = nccl_group[1] = g;    // a rank only ever uses the group of the session it owns
        group_world = world / 2;
    } else {
        nccl_group[0] = nccl_group[1] = nccl_world;
        group_world = world;
    }
    return LTR_OK;
}

int Removerter::reduce_flags(ltr_cloud map, const Session& over) {
    if (nccl_world >= 0) {
        if (group_world <= 1) return LTR_OK;
        return ltr_nccl_allreduce_flags(ctx, nccl_group[over.id], map) == LTR_OK ? LTR_OK : fail(LTR_ERR_CUDA, ltr_last_error(ctx));
    }
    if (!has_comm || comm.world <= 1) return LTR_OK;
    uint8_t* p; int64_t n;
    CK(ltr_flags_device_ptr(ctx, map, &p, &n));
    CK(ltr_synchronize(ctx));
    if (n > 0 && comm.allreduce_max_u8(comm.user, p, n) != 0) return fail(LTR_ERR_CUDA, "allreduce_max_u8 hook failed");
    return LTR_OK;
}

int Removerter::gather_clouds(const Session& over, int n, ltr_cloud** clouds) {
    if (nccl_world >= 0) {
        if (group_world <= 1) return LTR_OK;
        ltr_cloud local[16] = {0}, out[16] = {0};
        if (n > 16) return fail(LTR_ERR_INVALID, "too many clouds in one gather");
        for (int i = 0; i < n; ++i) local[i] = *clouds[i];
        CK(ltr_nccl_allgather_clouds(ctx, nccl_group[over.id], n, local, out));
        for (int i = 0; i < n; ++i) CK(set(clouds[i], out[i]));
        return LTR_OK;
    }
    if (!has_comm || comm.world <= 1) return LTR_OK;
    for (int i = 0; i < n; ++i) {
        ltr_cloud* cloud = clouds[i];
        float *sx, *sy, *sz, *si; int64_t m;
        CK(ltr_cloud_device_ptrs(ctx, *cloud, &sx, &sy, &sz, &si, &m));
        std::vector<int64_t> counts((size_t)comm.world), displs((size_t)comm.world);
        if (comm.allgather_i64(comm.user, m, counts.data()) != 0) return fail(LTR_ERR_CUDA, "allgather_i64 hook failed");
        int64_t total = 0;
        for (int r = 0; r < comm.world; ++r) { displs[r] = total; total += counts[r]; }
        ltr_cloud g;
        CK(ltr_cloud_alloc(ctx, total, &g));
        float *dx, *dy, *dz, *di; int64_t mm;
        CK(ltr_cloud_device_ptrs(ctx, g, &dx, &dy, &dz, &di, &mm));
        CK(ltr_synchronize(ctx));
        const float* src[4] = {sx, sy, sz, si};
        float* dst[4] = {dx, dy, dz, di};
        for (int c = 0; c < 4; ++c)
            if (comm.allgatherv_f32(comm.user, src[c], m, dst[c], counts.data(), displs.data()) != 0) return fail(LTR_ERR_CUDA, "allgatherv_f32 hook failed");
        CK(set(cloud, g));
    }
    return LTR_OK;
}

int Removerter::exchange_with_partner(int n_send, const ltr_cloud* send, int n_recv, ltr_cloud* recv) {
    if (!split) return fail(LTR_ERR_INVALID, "exchange_with_partner outside session-split mode");
    return ltr_nccl_exchange_clouds(ctx, nccl_world, partner(), n_send, send, n_recv, recv) == LTR_OK ? LTR_OK : fail(LTR_ERR_CUDA, ltr_last_error(ctx));
}

int Removerter::load_session(int sess, const float* xyzi, const int64_t* offsets, const double* poses, const double* inv_poses, int K) {
    Session& S = sess == 0 ? central_sess_ : query_sess_;
    ltr_scanset ss; ltr_poses ps;
    CK(ltr_scanset_upload(ctx, xyzi, offsets, K, &ss));
    CK(ltr_poses_upload(ctx, poses, inv_poses, K, &ps));
    CK(set(&S.keyframe_scans_, ss, true));
    if (S.keyframe_poses_ >= 0) CK(ltr_poses_free(ctx, S.keyframe_poses_));
    S.keyframe_poses_ = ps;
    S.num_keyframes_ = K;
    return LTR_OK;
}

int Removerter::octreeDownsampling(ltr_cloud* cloud, float leaf) {
    ltr_cloud v;
    CK(ltr_voxel_centroid(ctx, *cloud, leaf, &v));
    return set(cloud, v);
}

int Removerter::octreeDownsamplingAppended(ltr_cloud* cloud, float leaf) {
    int64_t n = 0;
    CK(ltr_cloud_size(ctx, *cloud, &n));
    if (!(nccl_world >= 0 && group_world > 1) || n < dist_voxel_min) return octreeDownsampling(cloud, leaf);
    const int g = nccl_group[0];            // the rank's own group (both entries are the same communicator)
    int32_t r = 0, w = 1;
    CK(ltr_nccl_info(ctx, g, &r, &w));
    ltr_cloud part, v;
    CK(ltr_cloud_slice(ctx, *cloud, n * r / w, n * (r + 1) / w, &part));
    const int rc = ltr_nccl_voxel_centroid_merged(ctx, g, part, leaf, &v);
    ltr_cloud_free(ctx, part);
    if (rc != LTR_OK) return fail(rc, ltr_last_error(ctx));
    return set(cloud, v);
}

int Removerter::mergeScansWithinGlobalCoordUtil(Session& s, ltr_scanset scans, ltr_cloud* out) {
    ltr_cloud m;
    CK(ltr_merge_scans_global(ctx, scans, s.keyframe_poses_, &m));
    CK(gather_cloud(s, &m));  // keyframe order == rank order (contiguous keyframe blocks)
    *out = m;
    return LTR_OK;
}

int Removerter::mergeAndDownsample(Session& s, ltr_scanset scans, float leaf, ltr_cloud* out) {
    if (nccl_world >= 0 && group_world > 1) {
        ltr_cloud m, v;
        CK(ltr_merge_scans_global(ctx, scans, s.keyframe_poses_, &m));      // this rank's keyframes only
        const int rc = ltr_nccl_voxel_centroid_merged(ctx, nccl_group[s.id], m, leaf, &v);
        ltr_cloud_free(ctx, m);
        if (rc != LTR_OK) return fail(rc, ltr_last_error(ctx));
        *out = v;
        return LTR_OK;
    }
    ltr_cloud m;
    CK(mergeScansWithinGlobalCoordUtil(s, scans, &m));
    CK(octreeDownsampling(&m, leaf));
    *out = m;
    return LTR_OK;
}

int Removerter::precleaningKeyframes(float radius) {
    for (Session* s : {&central_sess_, &query_sess_}) {
        if (!owns(*s)) continue;
        ltr_scanset c;
        CK(ltr_preclean(ctx, s->keyframe_scans_, radius, &c));
        CK(set(&s->keyframe_scans_, c, true));
    }
    return LTR_OK;
}

int Removerter::makeGlobalMap(Session& s) {
    ltr_cloud orig, curr;
    if (nccl_world >= 0 && group_world > 1) {
        // several ranks: the voxelised map comes from the distributed voxeliser (each rank sorts its key range of the ~10^8 raw points);
        // map_global_orig_ keeps its meaning (the merged raw cloud of ALL keyframes) through a plain gather afterwards
        CK(ltr_merge_scans_global(ctx, s.keyframe_scans_, s.keyframe_poses_, &orig));       // _sess.mergeScansWithinGlobalCoord() (:218), this rank's keyframes
        if (ltr_nccl_voxel_centroid_merged(ctx, nccl_group[s.id], orig, P.downsample_voxel_size, &curr) != LTR_OK) return fail(LTR_ERR_CUDA, ltr_last_error(ctx));   // :225
        CK(gather_cloud(s, &orig));
        CK(set(&s.map_global_orig_, orig));
    } else {
        CK(mergeScansWithinGlobalCoordUtil(s, s.keyframe_scans_, &orig));       // _sess.mergeScansWithinGlobalCoord() (:218)
        CK(set(&s.map_global_orig_, orig));
        CK(ltr_voxel_centroid(ctx, s.map_global_orig_, P.downsample_voxel_size, &curr));  // :225
    }
    CK(set(&s.map_global_curr_, curr));
    CK(save("OriginalNoisy" + s.sess_type_ + "MapGlobal", s.map_global_curr_));       // :231
    return LTR_OK;
}
int Removerter::makeGlobalMap() {
    if (owns(central_sess_)) CK(makeGlobalMap(central_sess_));
    if (owns(query_sess_)) CK(makeGlobalMap(query_sess_));
    return LTR_OK;
}

// partitionCurrentMap / ForND / ForPD (:801-828, 771-799, 740-768)
int Removerter::partitionCurrentMapGeneric(ltr_cloud map, Session& source, ltr_scanset scans, int mode, float res, const char* what,
                                           ltr_cloud* stat, ltr_cloud* dyn) {
    int64_t n_map = 0;
    CK(ltr_cloud_size(ctx, map, &n_map));
    // n_dynamic is taken from the partition below (after the cross-rank flag union), so no count is requested here
    CK(ltr_remove_pass(ctx, map, scans, source.keyframe_poses_, 0, source.num_keyframes_, mode, res, 0.1f, 0, nullptr));
    CK(reduce_flags(map, source));
    CK(ltr_apply_partition(ctx, map, stat, dyn));
    int64_t n_dyn = 0;
    CK(ltr_cloud_size(ctx, *dyn, &n_dyn));
    log.push_back(PassLog{what, n_map, n_dyn, -1, -1});
    return LTR_OK;
}

int Removerter::removeOnce(Session& t, Session& s, float res) {
    ltr_cloud st, dy;
    CK(partitionCurrentMapGeneric(t.map_global_curr_, s, s.keyframe_scans_, LTR_MODE_HD, res, "removeOnce", &st, &dy));
    CK(set(&t.map_global_curr_static_, st));                       // :894-895
    CK(octreeDownsampling(&t.map_global_curr_static_, 0.05f));     // :896
    CK(assign(&t.map_global_curr_, t.map_global_curr_static_));    // :899-900
    CK(append(&t.map_global_curr_dynamic_, dy));                   // :902
    CK(ltr_cloud_free(ctx, dy));
    CK(octreeDownsamplingAppended(&t.map_global_curr_dynamic_, 0.05f));    // :903
    CK(ltr_cloud_size(ctx, t.map_global_curr_static_, &log.back().n_static_after));
    CK(ltr_cloud_size(ctx, t.map_global_curr_dynamic_, &log.back().n_dynamic_after));
    return LTR_OK;
}

int Removerter::revertOnce(Session& t, Session& s, float res) {
    ltr_cloud st, dy;
    CK(partitionCurrentMapGeneric(t.map_global_curr_, s, s.keyframe_scans_, LTR_MODE_HD, res, "revertOnce", &st, &dy));
    CK(set(&t.map_global_curr_dynamic_, dy));                      // :919-920
    CK(octreeDownsampling(&t.map_global_curr_dynamic_, 0.05f));    // :921
    CK(assign(&t.map_global_curr_, t.map_global_curr_dynamic_));   // :924-925
    CK(append(&t.map_global_curr_static_, st));                    // :927
    CK(ltr_cloud_free(ctx, st));
    CK(octreeDownsamplingAppended(&t.map_global_curr_static_, 0.05f));     // :928
    CK(ltr_cloud_size(ctx, t.map_global_curr_static_, &log.back().n_static_after));
    CK(ltr_cloud_size(ctx, t.map_global_curr_dynamic_, &log.back().n_dynamic_after));
    return LTR_OK;
}

int Removerter::iremoveOnceForND(Session& t, Session& s, float res) {
    ltr_cloud st, dy;
    CK(partitionCurrentMapGeneric(t.map_global_nd_, s, s.keyframe_scans_static_projected_, LTR_MODE_ND, res, "iremoveOnceForND", &st, &dy));
    CK(set(&t.map_global_nd_strong_, st));                         // :843-844
    CK(octreeDownsampling(&t.map_global_nd_strong_, 0.05f));       // :845
    CK(assign(&t.map_global_nd_, t.map_global_nd_strong_));        // :848-849
    CK(append(&t.map_global_nd_weak_, dy));                        // :851
    CK(ltr_cloud_free(ctx, dy));
    CK(octreeDownsampling(&t.map_global_nd_weak_, 0.05f));         // :852
    CK(ltr_cloud_size(ctx, t.map_global_nd_strong_, &log.back().n_static_after));
    CK(ltr_cloud_size(ctx, t.map_global_nd_weak_, &log.back().n_dynamic_after));
    return LTR_OK;
}

int Removerter::removeOnceForPD(Session& t, Session& s, float res) {
    ltr_cloud st, dy;
    CK(partitionCurrentMapGeneric(t.map_global_pd_, s, s.keyframe_scans_static_projected_, LTR_MODE_PD, res, "removeOnceForPD", &st, &dy));
    CK(set(&t.map_global_pd_strong_, st));                         // :868-869
    CK(octreeDownsampling(&t.map_global_pd_strong_, 0.05f));       // :870
    CK(assign(&t.map_global_pd_, t.map_global_pd_strong_));        // :873-874
    CK(append(&t.map_global_pd_weak_, dy));                        // :876
    CK(ltr_cloud_free(ctx, dy));
    CK(octreeDownsampling(&t.map_global_pd_weak_, 0.05f));         // :878
    CK(ltr_cloud_size(ctx, t.map_global_pd_strong_, &log.back().n_static_after));
    CK(ltr_cloud_size(ctx, t.map_global_pd_weak_, &log.back().n_dynamic_after));
    return LTR_OK;
}

int Removerter::resetCurrrentMapAsDynamic(Session& s, bool as_dynamic) {
    return assign(&s.map_global_curr_, as_dynamic ? s.map_global_curr_dynamic_ : s.map_global_curr_static_);  // :714-732
}

int Removerter::selfRemovert(Session& s) {
    for (int i = 0; i < P.n_schedule; ++i) {
        if (P.schedule_op[i] == LTRH_OP_REMOVE) {
            CK(removeOnce(s, s, P.schedule_res[i]));
        } else {
            CK(resetCurrrentMapAsDynamic(s, true));
            CK(revertOnce(s, s, P.schedule_res[i]));
            CK(resetCurrrentMapAsDynamic(s, false));
        }
    }
    return LTR_OK;
}

int Removerter::extractHighDynPointsViaKnnDiff(Session& s, ltr_cloud target_map) {
    ltr_scanset di;
    CK(ltr_knn_diff(ctx, s.keyframe_scans_, s.keyframe_poses_, 0, target_map, P.num_nn_points_within, P.dist_nn_points_within, nullptr, &di));
    return set(&s.keyframe_scans_dynamic_, di, true);
}

int Removerter::removeHighDynamicPoints() {
    Session& C = central_sess_;
    Session& Q = query_sess_;
    {
        StageTimer t(*this, "hd_remove");
        if (owns(C)) CK(selfRemovert(C));  // shipped schedule: removeOnce(central, central, 2.5) (:1584)
        if (owns(Q)) CK(selfRemovert(Q));  // shipped schedule: removeOnce(query, query, 2.5)     (:1587)
    }
    if (P.extract_high_dyn_knn) {
        StageTimer t(*this, "hd_knn");
        if (owns(C)) CK(extractHighDynPointsViaKnnDiff(C, C.map_global_curr_static_));  // :1591
        if (owns(Q)) CK(extractHighDynPointsViaKnnDiff(Q, Q.map_global_curr_static_));  // :1592
        ltr_cloud c = -1, q = -1;
        if (owns(C)) CK(mergeAndDownsample(C, C.keyframe_scans_dynamic_, 0.05f, &c));  // :1594, :1597
        if (owns(Q)) CK(mergeAndDownsample(Q, Q.keyframe_scans_dynamic_, 0.05f, &q));  // :1595, :1598
        if (owns(C)) { CK(save("central_sess_high_dyn", c)); CK(ltr_cloud_free(ctx, c)); }  // :1600
        if (owns(Q)) { CK(save("query_sess_high_dyn", q)); CK(ltr_cloud_free(ctx, q)); }    // :1601
    }
    if (split) {
        // Step 2 reads the OTHER session's static map (Removerter.cpp:1416, 1418): hand mine to my partner, take theirs
        StageTimer t(*this, "exchange");
        Session& mine = owns(C) ? C : Q;
        Session& other = owns(C) ? Q : C;
        ltr_cloud got = -1;
        CK(exchange_with_partner(1, &mine.map_global_curr_static_, 1, &got));
        CK(set(&other.map_global_curr_static_, got));
    }
    return LTR_OK;
}

int Removerter::parseScansViaProjection(Session& s, ltr_cloud map, ltr_scanset* vec_to_store) {
    ltr_scanset out;
    CK(ltr_parse_projected(ctx, map, s.keyframe_poses_, 0, s.num_keyframes_, Session::kReprojectionAlpha, &out));
    return set(vec_to_store, out, true);
}

int Removerter::parseStaticScansViaProjection() {
    StageTimer t(*this, "parse_static");
    if (owns(central_sess_)) CK(parseScansViaProjection(central_sess_, central_sess_.map_global_curr_, &central_sess_.keyframe_scans_static_projected_));  // Session.cpp:305-308
    if (owns(query_sess_)) CK(parseScansViaProjection(query_sess_, query_sess_.map_global_curr_, &query_sess_.keyframe_scans_static_projected_));
    return LTR_OK;
}

int Removerter::extractLowDynPointsViaKnnDiff(Session& s, ltr_cloud target_map) {
    // Session.cpp:395-402 (0.4 m downsample + ICP target) feed an ICP refinement that is hard-disabled (Session.cpp:551): skipped.
    ltr_scanset co, di;
    CK(ltr_knn_diff(ctx, s.keyframe_scans_static_projected_, s.keyframe_poses_, 0, target_map, P.num_nn_points_within, P.dist_nn_points_within, &co, &di));
    CK(set(&s.scans_knn_coexist_, co, true));
    CK(set(&s.scans_knn_diff_, di, true));
    return LTR_OK;
}

int Removerter::constructGlobalNDMap(Session& s) {
    ltr_cloud m;
    CK(mergeAndDownsample(s, s.scans_knn_diff_, 0.05f, &m));   // Session.cpp:432-434
    return set(&s.map_global_nd_, m);
}

int Removerter::constructGlobalPDMap(Session& s) {
    ltr_cloud m;
    CK(mergeAndDownsample(s, s.scans_knn_diff_, 0.05f, &m));   // Session.cpp:439-441
    CK(set(&s.map_global_pd_, m));
    return assign(&s.map_global_pd_orig_, s.map_global_pd_);  // Session.cpp:444
}

int Removerter::removeWeakNDMapPointsHavingStrongNDInNear(Session& s) {
    int64_t n_strong = 0;
    CK(ltr_cloud_size(ctx, s.map_global_nd_strong_, &n_strong));
    if (n_strong == 0) return LTR_OK;  // Session.cpp:454-455
    ltr_cloud near_, far_;
    // hard-coded k = 2, threshold 1.0 (Session.cpp:468-469)
    CK(ltr_knn_split_cloud(ctx, s.map_global_nd_weak_, s.map_global_nd_strong_, 2, 1.0f, &near_, &far_));
    CK(append(&s.map_global_nd_strong_, near_));  // Session.cpp:482 (+=)
    CK(ltr_cloud_free(ctx, near_));
    CK(set(&s.map_global_nd_weak_, far_));        // Session.cpp:483 (=)
    return LTR_OK;
}

int Removerter::filterStrongND(Session& t, Session& s) {
    for (int i = 0; i < 3; ++i) CK(iremoveOnceForND(t, s, 2.5f));  // :1407-1410
    return LTR_OK;
}
int Removerter::filterStrongPD(Session& t, Session& s) {
    for (int i = 0; i < 3; ++i) CK(removeOnceForPD(t, s, 2.5f));   // :1397-1400
    return LTR_OK;
}

// In session-split mode a member lives on the ranks that run the loops producing it:
//   owners of the query keyframes:   filterStrongND (it projects into QUERY keyframes, :1403-1411) -> central map_global_nd_{,strong_,weak_},
//                                    union_map_queryside, pd_map, strong_nd_map, weak_nd_map
//   owners of the central keyframes: filterStrongPD (it projects into CENTRAL keyframes, :1395-1401) -> query map_global_pd_{,strong_,weak_} and the
//                                    central copies (:1434-1436), union_map_centralside, nd_map, strong_pd_map, weak_pd_map
int Removerter::detectLowDynamicPoints() {
    Session& C = central_sess_;
    Session& Q = query_sess_;
    {
        StageTimer t(*this, "ld_knn");
        if (owns(C)) CK(extractLowDynPointsViaKnnDiff(C, Q.map_global_curr_static_));  // :1416
        if (owns(Q)) CK(extractLowDynPointsViaKnnDiff(Q, C.map_global_curr_static_));  // :1418
    }
    {
        StageTimer t(*this, "ld_filter");
        if (owns(C)) CK(constructGlobalNDMap(C));           // :1421
        if (split) {
            // :1427 is independent of :1421-1424, so the query side builds its PD map now and the two maps cross over
            if (owns(Q)) CK(constructGlobalPDMap(Q));
            ltr_cloud got = -1;
            // the sender drops its copy: from here on the member lives only where its filter passes run
            if (owns(C)) {
                CK(exchange_with_partner(1, &C.map_global_nd_, 1, &got));
                CK(set(&C.map_global_nd_, -1));
                CK(set(&Q.map_global_pd_, got)); CK(assign(&Q.map_global_pd_orig_, Q.map_global_pd_));   // Session.cpp:444
            } else {
                CK(exchange_with_partner(1, &Q.map_global_pd_, 1, &got));
                CK(set(&Q.map_global_pd_, -1)); CK(set(&Q.map_global_pd_orig_, -1));
                CK(set(&C.map_global_nd_, got));
            }
        }
        if (owns(Q)) {
            if (C.map_global_nd_weak_ < 0) CK(assign(&C.map_global_nd_weak_, -1));
            CK(filterStrongND(C, Q));                           // :1423
            CK(removeWeakNDMapPointsHavingStrongNDInNear(C));   // :1424
        }
        if (!split) CK(constructGlobalPDMap(Q));                // :1427
        if (owns(C)) {
            if (Q.map_global_pd_weak_ < 0) CK(assign(&Q.map_global_pd_weak_, -1));
            CK(filterStrongPD(Q, C));                           // :1429
            // revertStrongPDMapPointsHavingWeakPDInNear: empty TODO in the reference (Session.cpp:447-450)
            CK(assign(&C.map_global_pd_, Q.map_global_pd_));                  // :1434
            CK(assign(&C.map_global_pd_orig_, Q.map_global_pd_orig_));        // :1435
            CK(assign(&C.map_global_pd_strong_, Q.map_global_pd_strong_));    // :1436
        }
    }
    {
        // always-on "save merged maps for visual debug" block (:1442-1480) including its in-place re-downsampling.
        // With several ranks the four merged clouds are voxelised without being gathered (mergeAndDownsample).
        StageTimer t(*this, "ld_merge_viz");
        ltr_cloud qco = -1, cco = -1, qdi = -1, cdi = -1;
        if (owns(Q)) CK(mergeAndDownsample(Q, Q.scans_knn_coexist_, 0.05f, &qco));   // :1443-1444
        if (owns(C)) CK(mergeAndDownsample(C, C.scans_knn_coexist_, 0.05f, &cco));   // :1448-1449
        if (owns(Q)) CK(mergeAndDownsample(Q, Q.scans_knn_diff_, 0.05f, &qdi));      // :1453-1454
        if (owns(C)) CK(mergeAndDownsample(C, C.scans_knn_diff_, 0.05f, &cdi));      // :1458-1459
        if (owns(Q)) { CK(save("union_map_queryside", qco)); CK(ltr_cloud_free(ctx, qco)); }     // :1446
        if (owns(C)) { CK(save("union_map_centralside", cco)); CK(ltr_cloud_free(ctx, cco)); }   // :1451
        if (owns(Q)) { CK(save("pd_map", qdi)); CK(ltr_cloud_free(ctx, qdi)); }                  // :1456
        if (owns(C)) { CK(save("nd_map", cdi)); CK(ltr_cloud_free(ctx, cdi)); }                  // :1461
        if (owns(Q)) {   // the ND members live where filterStrongND ran
            int64_t n = 0;
            CK(ltr_cloud_size(ctx, C.map_global_nd_strong_, &n));
            if (n != 0) { CK(octreeDownsampling(&C.map_global_nd_strong_, 0.05f)); CK(save("strong_nd_map", C.map_global_nd_strong_)); }   // :1463-1468
            CK(octreeDownsampling(&C.map_global_nd_weak_, 0.05f)); CK(save("weak_nd_map", C.map_global_nd_weak_));                          // :1470-1472
        }
        if (owns(C)) {   // the PD members live where filterStrongPD ran
            CK(octreeDownsampling(&Q.map_global_pd_strong_, 0.05f)); CK(save("strong_pd_map", Q.map_global_pd_strong_));   // :1474-1476
            CK(octreeDownsampling(&Q.map_global_pd_weak_, 0.05f)); CK(save("weak_pd_map", Q.map_global_pd_weak_));         // :1478-1480
        }
    }
    return LTR_OK;
}

int Removerter::updateCurrentMap() {
    Session& C = central_sess_;
    Session& Q = query_sess_;
    ltr_cloud uq = -1, uc = -1, upd = -1, strong = -1;
    if (owns(Q)) CK(mergeAndDownsample(Q, Q.scans_knn_coexist_, 0.05f, &uq));  // :1489-1490
    if (split) {
        // Step 3 runs on the owners of the central keyframes; the query side hands over what it holds of the central session
        if (owns(Q)) {
            const ltr_cloud send[3] = {uq, C.map_global_nd_weak_, C.map_global_nd_strong_};
            CK(exchange_with_partner(3, send, 0, nullptr));
            CK(ltr_cloud_free(ctx, uq));
            return LTR_OK;
        }
        ltr_cloud got[3] = {-1, -1, -1};
        CK(exchange_with_partner(0, nullptr, 3, got));
        uq = got[0];
        CK(set(&C.map_global_nd_weak_, got[1]));
        CK(set(&C.map_global_nd_strong_, got[2]));
    }
    CK(mergeAndDownsample(C, C.scans_knn_coexist_, 0.05f, &uc));  // :1492-1493
    CK(assign(&upd, uq));                       // :1495
    CK(append(&upd, uc));                       // :1496
    CK(append(&upd, C.map_global_nd_weak_));    // :1500
    CK(assign(&strong, upd));                   // :1505
    CK(append(&strong, C.map_global_pd_strong_));  // :1506
    CK(octreeDownsampling(&strong, 0.05f));     // :1507
    CK(append(&upd, C.map_global_pd_orig_));    // :1511
    CK(octreeDownsampling(&upd, 0.05f));        // :1512
    CK(set(&C.map_global_updated_, upd)); CK(save("updated_map", upd));                      // :1516-1517
    CK(set(&C.map_global_updated_strong_, strong)); CK(save("updated_map_strong", strong));  // :1519-1520
    CK(ltr_cloud_free(ctx, uq));
    CK(ltr_cloud_free(ctx, uc));
    return LTR_OK;
}

int Removerter::parseUpdatedStaticScansViaProjection() {
    Session& C = central_sess_;
    if (!owns(C)) return LTR_OK;
    CK(parseScansViaProjection(C, C.map_global_updated_, &C.keyframe_scans_updated_));
    CK(parseScansViaProjection(C, C.map_global_updated_strong_, &C.keyframe_scans_updated_strong_));
    return LTR_OK;
}

int Removerter::parseLDScansViaProjection() {
    Session& C = central_sess_;
    if (!owns(C)) return LTR_OK;
    CK(parseScansViaProjection(C, C.map_global_pd_orig_, &C.keyframe_scans_pd_));
    CK(parseScansViaProjection(C, C.map_global_pd_strong_, &C.keyframe_scans_strong_pd_));
    CK(parseScansViaProjection(C, C.map_global_nd_weak_, &C.keyframe_scans_weak_nd_));
    CK(parseScansViaProjection(C, C.map_global_nd_strong_, &C.keyframe_scans_strong_nd_));
    return LTR_OK;
}

int Removerter::updateScansScanwise() {
    Session& C = central_sess_;
    if (!owns(C)) return LTR_OK;
    ltr_scanset a, b, v;
    CK(ltr_scanset_concat_per_keyframe(ctx, C.keyframe_scans_updated_, C.keyframe_scans_weak_nd_, &a));  // Session.cpp:367-370
    CK(ltr_scanset_concat_per_keyframe(ctx, a, C.keyframe_scans_pd_, &b));                               // Session.cpp:371
    CK(ltr_voxel_centroid_per_keyframe(ctx, b, 0.05f, &v));                                              // Session.cpp:374
    CK(ltr_scanset_free(ctx, a));
    CK(ltr_scanset_free(ctx, b));
    return set(&C.keyframe_scans_updated_, v, true);                                                    // Session.cpp:377
}

int Removerter::run_step0() {
    StageTimer t(*this, "step0");
    CK(precleaningKeyframes(2.5f));  // run() :1660
    CK(makeGlobalMap());             // run() :1662
    return LTR_OK;
}
int Removerter::run_step12() {
    StageTimer t(*this, "step12");
    CK(removeHighDynamicPoints());        // :1665
    CK(parseStaticScansViaProjection());  // :1666
    CK(detectLowDynamicPoints());         // :1669
    return LTR_OK;
}
int Removerter::run_step3() {
    StageTimer t(*this, "step3");
    CK(updateCurrentMap());                       // :1672
    CK(parseUpdatedStaticScansViaProjection());   // :1673
    CK(parseLDScansViaProjection());              // :1674
    CK(updateScansScanwise());                    // :1675
    return LTR_OK;
}

int Removerter::reset_to_step0() {
    for (Session* s : {&central_sess_, &query_sess_}) {
        auto it = saved.find("OriginalNoisy" + s->sess_type_ + "MapGlobal");
        if (!owns(*s)) {   // session-split: everything held of the other session was derived after Step 0
            for (auto& kv : s->cloud_names()) if (*kv.second >= 0) { CK(ltr_cloud_free(ctx, *kv.second)); *kv.second = -1; }
            continue;
        }
        if (it == saved.end()) return fail(LTR_ERR_INVALID, "reset_to_step0: Step 0 has not run");
        for (auto& kv : s->cloud_names()) {
            if (kv.first == "map_global_orig_" || kv.first == "map_global_curr_") continue;
            if (*kv.second >= 0) { CK(ltr_cloud_free(ctx, *kv.second)); *kv.second = -1; }
        }
        for (auto& kv : s->scanset_names()) {
            if (kv.first == "keyframe_scans_") continue;
            if (*kv.second >= 0) { CK(ltr_scanset_free(ctx, *kv.second)); *kv.second = -1; }
        }
        CK(assign(&s->map_global_curr_, it->second));
    }
    for (auto it = saved.begin(); it != saved.end();) {
        if (it->first.rfind("OriginalNoisy", 0) == 0) { ++it; continue; }
        CK(ltr_cloud_free(ctx, it->second));
        it = saved.erase(it);
    }
    log.clear();
    timing.clear();
    return LTR_OK;
}

// LT-map cascade.  The reference chains sessions through its file protocol: a run's scans_updated/ directory and the
// central pose file are handed to the next run as its central session, where Session::loadKeyframes (Session.cpp:272-303) reads
// each scan back and passes it through pcl::VoxelGrid(downsample_voxel_size).  This does the same in memory: the updated scans
// (Removerter.cpp:1606-1650, what saveUpdatedScans writes) are brought to the host, voxel-gridded with the same load-time
// restatement the file driver uses (io.cpp), and become keyframe_scans_ of the central session; its poses stay.  Everything
// derived from the finished pair, and the query session, is released; the caller loads the next query and runs Steps 0-3 again.
int Removerter::cascade_promote_updated() {
    Session& C = central_sess_;
    if (owns(C) && C.keyframe_scans_updated_ < 0) return fail(LTR_ERR_INVALID, "cascade: Step 3 has not run (no keyframe_scans_updated_)");
    int32_t K = 0; int64_t total = 0;
    if (owns(C)) CK(ltr_scanset_info(ctx, C.keyframe_scans_updated_, &K, &total));
    // page-locked staging in both directions (a pageable 1 GB round trip costs more than Steps 0-3 of a session)
    // (kept across promotions: page-locking a gigabyte is itself slow)
    auto grow = [&](void** p, size_t* cap, size_t bytes) {
        if (*cap >= bytes) return true;
        ltr_pinned_free(*p); *p = nullptr; *cap = 0;
        if (ltr_pinned_alloc(2 * bytes, p) != LTR_OK) return false;   // the live scans grow from session to session: leave room
        *cap = 2 * bytes;
        return true;
    };
    if (!grow(&pin_in_, &pin_in_cap_, (size_t)std::max<int64_t>(total, 1) * 16)) return fail(LTR_ERR_NOMEM, "cascade: pinned staging allocation failed");
    float* xyzi = (float*)pin_in_;
    std::vector<int64_t> off((size_t)K + 1, 0);
    if (owns(C)) CK(ltr_scanset_download(ctx, C.keyframe_scans_updated_, xyzi, total, off.data()));
    // host-side load-time VoxelGrid, keyframes are independent (the reference's loadKeyframes loop is serial; the per-scan
    // arithmetic and its std::sort are unchanged, only different scans run on different threads).  Scans that take PCL's overflow
    // exit -- the common case for 0.05 m leaves on outdoor scans -- go from one staging buffer to the other with a single memcpy.
    VoxelGridBatch vg;
    vg.plan(xyzi, off.data(), K, P.downsample_voxel_size);
    const std::vector<int64_t>& out_off = vg.out_off;
    const float* out = xyzi;            // every scan passed through unchanged: the downloaded buffer IS the next session's input
    if (std::find(vg.unchanged.begin(), vg.unchanged.end(), (uint8_t)0) != vg.unchanged.end()) {
        if (!grow(&pin_out_, &pin_out_cap_, (size_t)std::max<int64_t>(out_off[(size_t)K], 1) * 16)) return fail(LTR_ERR_NOMEM, "cascade: pinned staging allocation failed");
        vg.emit(xyzi, off.data(), (float*)pin_out_);
        out = (const float*)pin_out_;
    }
    for (Session* s : {&central_sess_, &query_sess_}) {
        for (auto& kv : s->cloud_names()) if (*kv.second >= 0) { CK(ltr_cloud_free(ctx, *kv.second)); *kv.second = -1; }
        for (auto& kv : s->scanset_names()) if (*kv.second >= 0) { CK(ltr_scanset_free(ctx, *kv.second)); *kv.second = -1; }
    }
    for (auto& kv : saved) CK(ltr_cloud_free(ctx, kv.second));
    saved.clear();
    if (query_sess_.keyframe_poses_ >= 0) { CK(ltr_poses_free(ctx, query_sess_.keyframe_poses_)); query_sess_.keyframe_poses_ = -1; }
    query_sess_.num_keyframes_ = 0;
    if (owns(C)) {
        ltr_scanset ss;
        CK(ltr_scanset_upload(ctx, out, out_off.data(), K, &ss));
        C.keyframe_scans_ = ss;
    }
    log.clear();
    timing.clear();
    return LTR_OK;
}

}  // namespace ltremovert_b200

// ---------------------------------------------------------------------------------------------------------------
// flat C view (include/ltr_removert.h)
// ---------------------------------------------------------------------------------------------------------------
using ltremovert_b200::Removerter;
using ltremovert_b200::Session;

struct ltrh_removerter { Removerter* R; };
static thread_local std::string g_err;

extern "C" {

void ltrh_params_default(ltrh_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->device = 0;
    p->sequence_vfov = 50.0f;   // RosParamServer.cpp:15
    p->sequence_hfov = 360.0f;  // RosParamServer.cpp:16
    for (int i = 0; i < 16; ++i) p->ExtrinsicLiDARtoPoseBase[i] = (i % 5 == 0) ? 1.0 : 0.0;  // params_ltmapper.yaml:28-31
    p->num_nn_points_within = 2;        // params_ltmapper.yaml:65 (C++ default 3, RosParamServer.cpp:24)
    p->dist_nn_points_within = 0.01f;   // params_ltmapper.yaml:66 (C++ default 0.1, RosParamServer.cpp:25)
    p->downsample_voxel_size = 0.05f;   // RosParamServer.cpp:33
    p->n_schedule = 1;                  // shipped run(): removeOnce(2.5) (Removerter.cpp:1584, 1587)
    p->schedule_op[0] = LTRH_OP_REMOVE;
    p->schedule_res[0] = 2.5f;
    p->extract_high_dyn_knn = 1;
    p->transform_order = 0;
    p->keyframe_batch = 0;
    p->fast_path = 2;
}

int ltrh_create(ltrh_removerter** out, const ltrh_params* p) {
    if (!out || !p) { g_err = "null argument"; return LTR_ERR_INVALID; }
    if (p->n_schedule < 0 || p->n_schedule > LTRH_MAX_SCHEDULE) { g_err = "schedule too long"; return LTR_ERR_INVALID; }
    Removerter* R = new Removerter(*p);
    const int rc = R->init();
    if (rc != LTR_OK) { g_err = R->err; delete R; return rc; }
    *out = new ltrh_removerter{R};
    return LTR_OK;
}
void ltrh_destroy(ltrh_removerter* r) { if (r) { delete r->R; delete r; } }
const char* ltrh_last_error(const ltrh_removerter* r) { return r ? r->R->err.c_str() : g_err.c_str(); }
int ltrh_set_comm(ltrh_removerter* r, const ltr_comm* comm) {
    if (!r) return LTR_ERR_INVALID;
    if (!comm) { r->R->has_comm = false; return LTR_OK; }
    if (comm->world > 1 && (!comm->allreduce_max_u8 || !comm->allgather_i64 || !comm->allgatherv_f32)) { r->R->err = "comm hooks missing"; return LTR_ERR_INVALID; }
    r->R->comm = *comm;
    r->R->has_comm = true;
    return LTR_OK;
}
ltr_ctx* ltrh_context(ltrh_removerter* r) { return r ? r->R->ctx : nullptr; }
int ltrh_comm_init_nccl(ltrh_removerter* r, const uint8_t* id128, int32_t rank, int32_t world, int32_t split_sessions) {
    if (!r || !id128) return LTR_ERR_INVALID;
    r->R->err.clear();
    return r->R->comm_init_nccl(id128, rank, world, split_sessions);
}
// Eigen::Matrix4d::inverse() stand-in (general 4x4 inverse by cofactors) for the inverse keyframe poses of Session.cpp:110
void ltrh_invert_poses(const double* poses16, int32_t K, double* out16) {
    for (int k = 0; k < K; ++k) ltremovert_b200::invert4x4_public(poses16 + 16 * (size_t)k, out16 + 16 * (size_t)k);
}
int ltrh_owns_session(ltrh_removerter* r, int32_t sess) { return r && r->R->owns(sess == 0 ? r->R->central_sess_ : r->R->query_sess_) ? 1 : 0; }

int ltrh_load_session(ltrh_removerter* r, int32_t sess, const float* xyzi, const int64_t* offsets, const double* poses, const double* inv_poses, int32_t K) {
    if (!r || (sess != 0 && sess != 1)) return LTR_ERR_INVALID;
    r->R->err.clear();
    return r->R->load_session(sess, xyzi, offsets, poses, inv_poses, K);
}
int ltrh_run_step0(ltrh_removerter* r) { r->R->err.clear(); return r->R->run_step0(); }
int ltrh_run_step12(ltrh_removerter* r) { r->R->err.clear(); return r->R->run_step12(); }
int ltrh_run_step3(ltrh_removerter* r) { r->R->err.clear(); return r->R->run_step3(); }
int ltrh_reset_to_step0(ltrh_removerter* r) { r->R->err.clear(); return r->R->reset_to_step0(); }
int ltrh_cascade_promote_updated(ltrh_removerter* r) { if (!r) return LTR_ERR_INVALID; r->R->err.clear(); return r->R->cascade_promote_updated(); }

int ltrh_stage(ltrh_removerter* r, const char* name) {
    Removerter& R = *r->R;
    R.err.clear();
    const std::string s(name);
    if (s == "precleaningKeyframes") return R.precleaningKeyframes(2.5f);
    if (s == "makeGlobalMap") return R.makeGlobalMap();
    if (s == "removeHighDynamicPoints") return R.removeHighDynamicPoints();
    if (s == "parseStaticScansViaProjection") return R.parseStaticScansViaProjection();
    if (s == "detectLowDynamicPoints") return R.detectLowDynamicPoints();
    if (s == "updateCurrentMap") return R.updateCurrentMap();
    if (s == "parseUpdatedStaticScansViaProjection") return R.parseUpdatedStaticScansViaProjection();
    if (s == "parseLDScansViaProjection") return R.parseLDScansViaProjection();
    if (s == "updateScansScanwise") return R.updateScansScanwise();
    R.err = "unknown stage " + s;
    return LTR_ERR_INVALID;
}

int ltrh_cloud(ltrh_removerter* r, const char* name, int32_t sess, ltr_cloud* out) {
    Removerter& R = *r->R;
    const std::string n(name);
    if (n.rfind("saved:", 0) == 0) {
        auto it = R.saved.find(n.substr(6));
        if (it == R.saved.end()) { R.err = "no saved cloud " + n; return LTR_ERR_INVALID; }
        *out = it->second;
        return LTR_OK;
    }
    Session& S = sess == 0 ? R.central_sess_ : R.query_sess_;
    auto m = S.cloud_names();
    auto it = m.find(n);
    if (it == m.end() || *it->second < 0) { R.err = "no cloud " + n; return LTR_ERR_INVALID; }
    *out = *it->second;
    return LTR_OK;
}
int ltrh_scanset(ltrh_removerter* r, const char* name, int32_t sess, ltr_scanset* out) {
    Removerter& R = *r->R;
    Session& S = sess == 0 ? R.central_sess_ : R.query_sess_;
    auto m = S.scanset_names();
    auto it = m.find(name);
    if (it == m.end() || *it->second < 0) { R.err = std::string("no scanset ") + name; return LTR_ERR_INVALID; }
    *out = *it->second;
    return LTR_OK;
}
double ltrh_timing(ltrh_removerter* r, const char* key) {
    auto it = r->R->timing.find(key);
    return it == r->R->timing.end() ? 0.0 : it->second;
}
int32_t ltrh_log_count(ltrh_removerter* r) { return (int32_t)r->R->log.size(); }
int ltrh_log_get(ltrh_removerter* r, int32_t i, char* what, int32_t cap, int64_t* vals) {
    if (i < 0 || i >= (int32_t)r->R->log.size()) return LTR_ERR_INVALID;
    const auto& l = r->R->log[(size_t)i];
    std::snprintf(what, (size_t)cap, "%s", l.what.c_str());
    vals[0] = l.n_map; vals[1] = l.n_dynamic; vals[2] = l.n_static_after; vals[3] = l.n_dynamic_after;
    return LTR_OK;
}

}  // extern "C"
