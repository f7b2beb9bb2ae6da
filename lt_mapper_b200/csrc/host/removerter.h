// Host-side mirror of ltremovert::Removerter / ltremovert::Session (reference: ltremovert/include/removert/
// Removerter.h:9-204, Session.h:9-136).  Same member and method names, same statement order as
// ltremovert/src/Removerter.cpp / Session.cpp; every cloud is a device handle of the C-ABI in include/ltr_b200.h
// and every loop over keyframes or points is one C-ABI call.  Plain C++17, no CUDA, no ROS/PCL: a ROS build would
// convert pcl::PointCloud<PointXYZI> to/from the float[n][4] arrays at load/save time only.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include "../../../include/ltr_removert.h"

namespace ltremovert_b200 {

void invert4x4_public(const double* m, double* out);   // general 4x4 inverse by cofactors (removerter.cpp)

struct PassLog { std::string what; int64_t n_map, n_dynamic, n_static_after, n_dynamic_after; };

class Removerter;

class Session {
public:
    static constexpr float kReprojectionAlpha = 3.0f;  // Session.h:13
    std::string sess_type_;
    int id = 0;                                        // 0 = central, 1 = query
    int num_keyframes_ = 0;                            // this rank's keyframes
    ltr_poses keyframe_poses_ = -1;                    // keyframe_poses_ + keyframe_inverse_poses_ (Session.h:36-37)
    // per-keyframe clouds (Session.h:40-56)
    ltr_scanset keyframe_scans_ = -1, keyframe_scans_static_projected_ = -1, keyframe_scans_dynamic_ = -1;
    ltr_scanset scans_knn_coexist_ = -1, scans_knn_diff_ = -1;
    ltr_scanset keyframe_scans_updated_ = -1, keyframe_scans_updated_strong_ = -1, keyframe_scans_pd_ = -1,
                keyframe_scans_strong_pd_ = -1, keyframe_scans_strong_nd_ = -1, keyframe_scans_weak_nd_ = -1;
    // maps (Session.h:67-87)
    ltr_cloud map_global_orig_ = -1, map_global_curr_ = -1, map_global_curr_static_ = -1, map_global_curr_dynamic_ = -1;
    ltr_cloud map_global_updated_ = -1, map_global_updated_strong_ = -1;
    ltr_cloud map_global_nd_ = -1, map_global_nd_strong_ = -1, map_global_nd_weak_ = -1;
    ltr_cloud map_global_pd_ = -1, map_global_pd_orig_ = -1, map_global_pd_strong_ = -1, map_global_pd_weak_ = -1;

    std::map<std::string, ltr_cloud*> cloud_names();
    std::map<std::string, ltr_scanset*> scanset_names();
};

class Removerter {
public:
    explicit Removerter(const ltrh_params& p);
    ~Removerter();
    int init();  // creates the device context; returns ltr_status

    // ---- reference call graph (Removerter.cpp) ----
    int precleaningKeyframes(float radius);                                    // :102-106 -> Session.cpp:506-533
    int makeGlobalMap();                                                       // :248-252
    int makeGlobalMap(Session& s);                                             // :213-245
    int removeOnce(Session& target, Session& source, float res);               // :882-905
    int revertOnce(Session& target, Session& source, float res);               // :908-931
    int iremoveOnceForND(Session& target, Session& source, float res);         // :831-854
    int removeOnceForPD(Session& target, Session& source, float res);          // :856-880
    int resetCurrrentMapAsDynamic(Session& s, bool as_dynamic);                // :714-737
    int selfRemovert(Session& s);                                              // :1378-1393 generalised to the configured schedule
    int removeHighDynamicPoints();                                             // :1580-1604
    int parseStaticScansViaProjection();                                       // :1534-1538
    int detectLowDynamicPoints();                                              // :1413-1481
    int filterStrongND(Session& target, Session& source);                      // :1403-1411
    int filterStrongPD(Session& target, Session& source);                      // :1395-1401
    int updateCurrentMap();                                                    // :1483-1524
    int parseUpdatedStaticScansViaProjection();                                // :1551-1561
    int parseLDScansViaProjection();                                           // :1564-1577
    int updateScansScanwise();                                                 // :1540-1548 -> Session.cpp:362-380
    int run_step0();
    int run_step12();
    int run_step3();
    int reset_to_step0();  // benchmark plumbing, see include/ltr_removert.h
    int cascade_promote_updated();  // LT-map cascade: keyframe_scans_updated_ -> the next run's central keyframe_scans_ (see .cpp)

    // ---- Session methods (Session.cpp) ----
    int parseScansViaProjection(Session& s, ltr_cloud map, ltr_scanset* vec_to_store);     // Session.cpp:348-360
    int extractLowDynPointsViaKnnDiff(Session& s, ltr_cloud target_map);                   // Session.cpp:393-427
    int extractHighDynPointsViaKnnDiff(Session& s, ltr_cloud target_map);                  // Session.cpp:487-504
    int constructGlobalNDMap(Session& s);                                                  // Session.cpp:430-435
    int constructGlobalPDMap(Session& s);                                                  // Session.cpp:437-445
    int removeWeakNDMapPointsHavingStrongNDInNear(Session& s);                             // Session.cpp:452-484
    int mergeScansWithinGlobalCoordUtil(Session& s, ltr_scanset scans, ltr_cloud* out);    // utility.cpp:170-192 (+ rank gather)
    int octreeDownsampling(ltr_cloud* cloud, float leaf);                                  // utility.cpp:204-219, in place
    // the same for a freshly appended cloud ("*a += *b", never in octree order): with several ranks every rank contributes its slice of
    // the (replicated) cloud to the distributed voxeliser, so each sorts 1/G of it instead of all of it
    int octreeDownsamplingAppended(ltr_cloud* cloud, float leaf);
    // mergeScansWithinGlobalCoordUtil followed by octreeDownsampling of the result (the only way the path consumes a merged cloud):
    // with several ranks the raw merged cloud is never gathered (ltr_nccl_voxel_centroid_merged)
    int mergeAndDownsample(Session& s, ltr_scanset scans, float leaf, ltr_cloud* out);

    // ---- plumbing ----
    int load_session(int sess, const float* xyzi, const int64_t* offsets, const double* poses, const double* inv_poses, int K);
    int set(ltr_cloud* slot, ltr_cloud v);         // frees what the slot held, then stores v
    int set(ltr_scanset* slot, ltr_scanset v, bool);
    int assign(ltr_cloud* dst, ltr_cloud src);     // "*dst = *src" (deep copy)
    int append(ltr_cloud* dst, ltr_cloud src);     // "*dst += *src"
    int save(const std::string& name, ltr_cloud c);  // stands for pcl::io::savePCDFileBinary: keeps a device copy by name
    // ---- multi-GPU (see "Parallel decomposition" in removerter.cpp) ----
    int comm_init_nccl(const uint8_t* id128, int rank, int world, int split_sessions);   // native NCCL transport (ltr_nccl_*)
    bool owns(const Session& s) const { return !split || (rank < world / 2) == (s.id == 0); }
    bool multi() const { return world > 1; }
    int partner() const { return (rank + world / 2) % world; }
    int reduce_flags(ltr_cloud map, const Session& over);               // OR of the dynamic flags over the ranks that own `over`'s keyframes
    int gather_clouds(const Session& over, int n, ltr_cloud** clouds);  // each *clouds[i] <- rank-ordered concatenation over those ranks
    int gather_cloud(const Session& over, ltr_cloud* cloud) { ltr_cloud* one[1] = {cloud}; return gather_clouds(over, 1, one); }
    int exchange_with_partner(int n_send, const ltr_cloud* send, int n_recv, ltr_cloud* recv);   // split mode only
    int fail(int code, const std::string& msg);

    ltrh_params P;
    ltr_ctx* ctx = nullptr;
    ltr_comm comm{};          // legacy caller-supplied hooks (host-memory tests); superseded by the native transport when nccl_world >= 0
    bool has_comm = false;
    int rank = 0, world = 1;
    bool split = false;       // session-split: ranks [0, world/2) own the central keyframes, the others the query keyframes
    int nccl_world = -1;      // ltr_nccl communicator handles: everyone, ...
    int nccl_group[2] = {-1, -1};   // ... and the ranks that own session s
    int group_world = 1;
    // appended (replicated) clouds of at least this many points go through the distributed voxeliser, every rank contributing a slice.
    // Off by default: on 4 NVLinked ranks the two size exchanges + the exchange + the gather of a ~10 M point cloud cost more than sorting it on
    // every rank (hd_remove 115.7 -> 125.5 ms at 8 GPUs, profiles/r02_scaling.md); LTR_DIST_VOXEL_MIN turns it on (the multi-rank tests do).
    int64_t dist_voxel_min = INT64_MAX;
    void *pin_in_ = nullptr, *pin_out_ = nullptr;     // page-locked staging of cascade_promote_updated, grown on demand
    size_t pin_in_cap_ = 0, pin_out_cap_ = 0;
    Session central_sess_, query_sess_;
    std::map<std::string, ltr_cloud> saved;
    std::map<std::string, double> timing;
    std::vector<PassLog> log;
    std::string err;

private:
    int partitionCurrentMapGeneric(ltr_cloud map, Session& source, ltr_scanset scans, int mode, float res, const char* what,
                                   ltr_cloud* stat, ltr_cloud* dyn);
};

}  // namespace ltremovert_b200
