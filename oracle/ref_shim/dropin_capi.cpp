// ORACLE / REFERENCE SHIM (test infrastructure, not product code).
// INTEGRATION.md "B" made executable: the reference's own Removerter object (compiled from /root/reference behind the stand-in
// headers) keeps its loaders and writers, and everything between them -- Removerter::run()'s Steps 0-3 -- is handed to
// libltr_removert.so / libltr_b200.so through include/ltr_pcl_adapter.hpp.  tests/test_gpu_driver.py compares the files this writes
// with the files the unmodified reference run writes.  Built by `make -C oracle ref` into oracle/_ref/libltremovert_dropin.so.
#include <algorithm>
#include <array>
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
#define private public
#include "removert/Removerter.h"
#undef private
#include "ltr_pcl_adapter.hpp"

using ltremovert::Removerter;

#define DROPIN_TRY(x) do { const int rc__ = (x); if (rc__ != LTR_OK) { std::snprintf(err, (size_t)cap, "%s -> %d: %s", #x, rc__, G ? ltrh_last_error(G) : ltrh_last_error(nullptr)); if (G) ltrh_destroy(G); return rc__; } } while (0)

extern "C" int ref_dropin_run(void* h, int transform_order, char* err, int cap) {
    Removerter* R = (Removerter*)h;
    ltrh_removerter* G = nullptr;
    // ---- the reference's own loaders (run() :1656-1659) ----
    R->loadSessionInfo();
    R->parseKeyframes();
    R->loadKeyframes();
    // ---- the B200 orchestrator configured from the node's parameters (RosParamServer.cpp:4-63) ----
    ltrh_params p;
    ltrh_params_default(&p);
    p.sequence_vfov = R->kVFOV; p.sequence_hfov = R->kHFOV;
    p.num_nn_points_within = R->kNumKnnPointsToCompare; p.dist_nn_points_within = R->kScanKnnAndMapKnnAvgDiffThreshold;
    p.downsample_voxel_size = R->kDownsampleVoxelSize;
    ltr_pcl::to_row_major(R->kSE3MatExtrinsicLiDARtoPoseBase, p.ExtrinsicLiDARtoPoseBase);
    p.transform_order = transform_order;
    DROPIN_TRY(ltrh_create(&G, &p));
    auto& C = R->central_sess_;
    auto& Q = R->query_sess_;
    DROPIN_TRY(ltr_pcl::load_session(G, 0, C.keyframe_scans_, C.keyframe_poses_, C.keyframe_inverse_poses_));
    DROPIN_TRY(ltr_pcl::load_session(G, 1, Q.keyframe_scans_, Q.keyframe_poses_, Q.keyframe_inverse_poses_));
    // ---- run() :1660-1675 on the device ----
    DROPIN_TRY(ltrh_run_step0(G));
    DROPIN_TRY(ltrh_run_step12(G));
    DROPIN_TRY(ltrh_run_step3(G));
    // ---- results back into the node's own members; its own writers put them on disk ----
    pcl::PointCloud<PointType> tmp;
    if (R->kFlagSaveMapPointcloud) {
        for (const char* s : {"Central", "Query"}) {
            const std::string n = std::string("OriginalNoisy") + s + "MapGlobal";
            DROPIN_TRY(ltr_pcl::fetch_cloud(G, ("saved:" + n).c_str(), 0, tmp));
            pcl::io::savePCDFileBinary(R->save_pcd_directory_ + n + ".pcd", tmp);               // Removerter.cpp:231-232
        }
    }
    for (const char* n : {"central_sess_high_dyn", "query_sess_high_dyn", "union_map_queryside", "union_map_centralside", "pd_map", "nd_map",
                          "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map", "updated_map", "updated_map_strong"}) {
        ltr_cloud hc;
        if (ltrh_cloud(G, (std::string("saved:") + n).c_str(), 0, &hc) != LTR_OK) continue;   // strong_nd_map is only written when non-empty (:1464)
        DROPIN_TRY(ltr_pcl::fetch_cloud(G, (std::string("saved:") + n).c_str(), 0, tmp));
        pcl::io::savePCDFileBinary(R->save_pcd_directory_ + n + ".pcd", tmp);                   // :1446-1477, 1517-1520, 1600-1601
    }
    DROPIN_TRY(ltr_pcl::fetch_scans(G, "keyframe_scans_updated_", 0, C.keyframe_scans_updated_));
    DROPIN_TRY(ltr_pcl::fetch_scans(G, "keyframe_scans_updated_strong_", 0, C.keyframe_scans_updated_strong_));
    DROPIN_TRY(ltr_pcl::fetch_scans(G, "keyframe_scans_pd_", 0, C.keyframe_scans_pd_));
    DROPIN_TRY(ltr_pcl::fetch_scans(G, "keyframe_scans_strong_pd_", 0, C.keyframe_scans_strong_pd_));
    DROPIN_TRY(ltr_pcl::fetch_scans(G, "keyframe_scans_strong_nd_", 0, C.keyframe_scans_strong_nd_));
    R->saveAllTypeOfScans();                                                                    // the reference's own writer (:1606-1650)
    ltrh_destroy(G);
    return LTR_OK;
}
