/*
 * ltr_removert.h -- host-side orchestrator (libltr_removert.so) that mirrors the reference's
 * ltremovert::Removerter / ltremovert::Session call graph (ltremovert/src/Removerter.cpp:1653-1678,
 * ltremovert/include/removert/Session.h:39-136) on top of the device C-ABI in ltr_b200.h.
 *
 * The C++ classes live in lt_mapper_b200/csrc/host/removerter.{h,cpp}; this header is the flat C view of
 * them used by the Python test/bench harness and by any non-C++ caller.  Sessions are handed over in
 * memory in the form the reference holds them after Session::loadKeyframes (Session.cpp:266-302):
 * per keyframe a cloud of x, y, z, intensity floats plus a pose and its inverse.
 *
 * Multi-GPU: one process per GPU, each owning a contiguous block of keyframes of both sessions; maps are
 * replicated.  The two exchange points of the path (per-pass dynamic-flag union, rank-ordered gather of
 * merged clouds) are expressed as the ltr_comm hooks below so that the caller chooses the transport
 * (torch.distributed/NCCL in bench.py; any NCCL communicator natively).
 */
#ifndef LTR_REMOVERT_H_
#define LTR_REMOVERT_H_

#include "ltr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ltrh_removerter ltrh_removerter;

#define LTRH_MAX_SCHEDULE 32
enum { LTRH_OP_REMOVE = 0, LTRH_OP_REVERT = 1 };

/* Same keys as the reference's RosParamServer (ltremovert/src/RosParamServer.cpp:4-63). */
typedef struct ltrh_params {
    int32_t device;
    float sequence_vfov, sequence_hfov;          /* removert/sequence_vfov, sequence_hfov */
    double ExtrinsicLiDARtoPoseBase[16];         /* removert/ExtrinsicLiDARtoPoseBase, row-major */
    int32_t num_nn_points_within;                /* kNumKnnPointsToCompare */
    float dist_nn_points_within;                 /* kScanKnnAndMapKnnAvgDiffThreshold (mean SQUARED distance) */
    float downsample_voxel_size;                 /* kDownsampleVoxelSize */
    /* HD-removal schedule: the shipped run() does one removeOnce(2.5) per session (Removerter.cpp:1584,1587);
     * selfRemovert (Removerter.cpp:1378-1393) is {REMOVE r, REVERT 0.95 r, REMOVE r} per resolution r. */
    int32_t n_schedule;
    int32_t schedule_op[LTRH_MAX_SCHEDULE];
    float schedule_res[LTRH_MAX_SCHEDULE];
    int32_t extract_high_dyn_knn;                /* extractHighDynPointsViaKnnDiff in removeHighDynamicPoints (viz output) */
    int32_t transform_order;                     /* see ltr_config */
    int32_t keyframe_batch, fast_path;           /* see ltr_config */
} ltrh_params;

/* Collective hooks; NULL pointers (or world == 1) mean single process.  Device pointers belong to the
 * context's GPU; the orchestrator synchronises its stream before calling a hook and expects the hook to
 * return after the data is in place (or ordered on the default stream and synchronised). */
typedef struct ltr_comm {
    int32_t rank, world;
    void* user;
    /* in-place max-reduction of n bytes across ranks (logical OR of the per-rank dynamic flags) */
    int (*allreduce_max_u8)(void* user, uint8_t* dev, int64_t n);
    /* all-gather of one int64 per rank into out[world] (host memory) */
    int (*allgather_i64)(void* user, int64_t local, int64_t* out);
    /* rank-ordered variable all-gather of float arrays: dst[displs[r] .. displs[r]+counts[r]) = rank r's src[0..counts[r]) */
    int (*allgatherv_f32)(void* user, const float* dev_src, int64_t n_local, float* dev_dst, const int64_t* counts, const int64_t* displs);
} ltr_comm;

void ltrh_params_default(ltrh_params* p);
int ltrh_create(ltrh_removerter** out, const ltrh_params* p);
void ltrh_destroy(ltrh_removerter* r);
const char* ltrh_last_error(const ltrh_removerter* r);
int ltrh_set_comm(ltrh_removerter* r, const ltr_comm* comm);
ltr_ctx* ltrh_context(ltrh_removerter* r);
/* Native multi-GPU transport (ltr_nccl_* of ltr_b200.h; NCCL on the context's stream, no host code in the loop).  id128 = the 128 bytes
 * of ltr_nccl_unique_id() taken on one rank and shared with the others by the launcher.  split_sessions != 0 with an even world:
 * ranks [0, world/2) own the keyframes of the central session, the others those of the query session (each a contiguous block,
 * in rank order); Step 1 of the two sessions and the two halves of Step 2 then run concurrently and a map crosses over exactly
 * where the reference reads the other session's member.  Otherwise every rank owns a block of BOTH sessions.
 * ltrh_load_session on a rank that does not own the session is called with K = 0. */
int ltrh_comm_init_nccl(ltrh_removerter* r, const uint8_t* id128, int32_t rank, int32_t world, int32_t split_sessions);
int ltrh_owns_session(ltrh_removerter* r, int32_t sess);   /* 1 iff this rank holds keyframes of `sess` */

/* Inverse keyframe poses the way the reference obtains them (Session.cpp:110, Eigen::Matrix4d::inverse(): general 4x4 inverse by cofactors).
 * ltrh_load_session takes inverse poses as an INPUT so that every implementation works from identical doubles; callers that do not have the
 * reference's inverses at hand should use this one rather than an LU-based inverse, whose last bits differ. */
void ltrh_invert_poses(const double* poses16, int32_t K, double* out16);

/* sess: 0 = central, 1 = query.  Keyframes are this rank's block (all keyframes when world == 1). */
int ltrh_load_session(ltrh_removerter* r, int32_t sess, const float* xyzi, const int64_t* offsets, const double* poses,
                      const double* inv_poses, int32_t K);

/* Removerter::run() split at its step comments (Removerter.cpp:1655-1676) */
int ltrh_run_step0(ltrh_removerter* r);   /* precleaningKeyframes(2.5) + makeGlobalMap            (:1660-1662) */
int ltrh_run_step12(ltrh_removerter* r);  /* removeHighDynamicPoints + parseStaticScansViaProjection + detectLowDynamicPoints (:1665-1669) */
int ltrh_run_step3(ltrh_removerter* r);   /* updateCurrentMap .. updateScansScanwise               (:1672-1675) */
/* Benchmark/test plumbing (not in the reference): restores the state right after Step 0 (map_global_curr_ = the
 * voxelised original map, everything derived freed) so that Step 1+2 can be timed repeatedly; clears log and timings. */
int ltrh_reset_to_step0(ltrh_removerter* r);
/* LT-map cascade (multi-session chaining).  The reference has no call for this: it is done by pointing the next run's
 * central_sess_scan_dir at the previous run's scans_updated/ (written by saveUpdatedScans, Removerter.cpp:1620-1623) with the same
 * central pose file; Session::loadKeyframes (Session.cpp:272-303) then voxel-grids each scan at downsample_voxel_size.  This
 * entry does exactly that in memory after ltrh_run_step3: central keyframe_scans_ <- VoxelGrid(keyframe_scans_updated_), central poses
 * kept, query session and every derived cloud released, log cleared.  Then ltrh_load_session(r, 1, next query) and run the steps again. */
int ltrh_cascade_promote_updated(ltrh_removerter* r);
/* a single member function of Removerter by name, e.g. "removeHighDynamicPoints" */
int ltrh_stage(ltrh_removerter* r, const char* name);

/* Named device clouds: Session members ("map_global_curr_static_", ...) or what the reference saves as PCD
 * ("saved:pd_map", ...).  Scan sets: per-keyframe members ("keyframe_scans_static_projected_", ...). */
int ltrh_cloud(ltrh_removerter* r, const char* name, int32_t sess, ltr_cloud* out);
int ltrh_scanset(ltrh_removerter* r, const char* name, int32_t sess, ltr_scanset* out);
double ltrh_timing(ltrh_removerter* r, const char* key);  /* seconds spent in a stage (host clock around synchronised calls) */
int32_t ltrh_log_count(ltrh_removerter* r);
/* pass log (the counters the reference prints at Removerter.cpp:811-822, 897, 904): vals = n_map, n_dynamic, n_static_after, n_dynamic_after */
int ltrh_log_get(ltrh_removerter* r, int32_t i, char* what, int32_t cap, int64_t* vals);

/* ---- file-level helpers of the ltremovert node surface (lt_mapper_b200/csrc/host/io.h), used by apps/ltremovert_b200.cpp ---- */
const char* ltrh_io_last_error(void);
/* binary/ascii PCD with float32 x y z [intensity] (pcl::io::loadPCDFile at Session.cpp:275); returns the point count or -1 */
int64_t ltrh_io_read_pcd(const char* path, float* xyzi, int64_t capacity);
/* pcl::io::savePCDFileBinary<PointXYZI>; octree_layout: WIDTH 1 / HEIGHT n as set by octreeDownsampling (utility.cpp:217-218) */
int ltrh_io_write_pcd(const char* path, const float* xyzi, int64_t n, int32_t octree_layout);
/* pose file: one line of 12 or 16 numbers per scan (Session.cpp:102-114); returns the pose count or -1 */
int32_t ltrh_io_read_poses(const char* path, double* poses16, int32_t capacity);
/* Session::parseKeyframes(range, gap) (Session.cpp:138-174, including its skip-two quirk) */
int32_t ltrh_io_parse_keyframes(int32_t num_scans, int32_t start_idx, int32_t end_idx, int32_t gap, int32_t* out, int32_t capacity);
/* Session::parseKeyframesInROI (Session.cpp:230-263, 10 m) */
int32_t ltrh_io_parse_keyframes_in_roi(const double* scan_poses16, int32_t n, const double* roi_poses16, int32_t m, int32_t gap, int32_t* out, int32_t capacity);
/* pcl::VoxelGrid at load (Session.cpp:284-289) incl. the int32 overflow fallback that returns the input unchanged */
int64_t ltrh_io_voxel_grid(const float* xyzi, int64_t n, float leaf, float* out, int64_t capacity, int32_t* overflowed);
/* the same filter over the K scans of a session held back to back (scan k = points [off[k], off[k+1])), scans on different OpenMP
 * threads (Session::loadKeyframes' loop, Session.cpp:272-303).  Returns the total number of output points and fills out_off[K + 1];
 * `out` is written only if capacity (points) suffices. */
int64_t ltrh_io_voxel_grid_scans(const float* xyzi, const int64_t* off, int32_t K, float leaf, float* out, int64_t capacity, int64_t* out_off);
/* one key of a roslaunch-style yaml ("removert/key"): scalar text and/or numeric list */
int ltrh_io_yaml_get(const char* path, const char* key, char* value, int32_t capacity, double* list, int32_t list_capacity, int32_t* list_n);

#ifdef __cplusplus
}
#endif
#endif /* LTR_REMOVERT_H_ */
