// ltr_pcl_adapter.hpp -- header-only glue between the reference's host types (pcl::PointCloud<pcl::PointXYZI>, Eigen::Matrix4d,
// ltremovert/include/removert/utility.h:90-91) and the C-ABI of libltr_b200.so / libltr_removert.so, for a ROS build of the
// ltremovert node that links the B200 library.  It is NOT used by anything in this repository's product or
// bench path (PCL and Eigen do not exist here); tests/test_abi.py compiles it against the stand-in headers of oracle/ref_shim to keep
// it honest, and it only uses members that exist in real PCL >= 1.8 / Eigen 3.3 as well.
//
// Include <pcl/point_cloud.h>, <pcl/point_types.h> and <Eigen/Dense> before this header.
#pragma once
#include <cstdint>
#include <vector>

#include "ltr_b200.h"
#include "ltr_removert.h"

namespace ltr_pcl {

typedef pcl::PointXYZI PointType;   // utility.h:90
typedef pcl::PointCloud<PointType> Cloud;

// pcl::PointXYZI is 32 B on the host (xyz + pad | intensity + pad); the ABI takes float[n][4] = x y z intensity
inline void pack(const Cloud& c, std::vector<float>& xyzi) {
    xyzi.resize(4 * c.points.size());
    for (std::size_t i = 0; i < c.points.size(); ++i) {
        xyzi[4 * i] = c.points[i].x; xyzi[4 * i + 1] = c.points[i].y; xyzi[4 * i + 2] = c.points[i].z; xyzi[4 * i + 3] = c.points[i].intensity;
    }
}
// octree_layout: width = 1, height = n as octreeDownsampling leaves it (utility.cpp:217-218); otherwise width = n, height = 1
inline void unpack(const float* xyzi, std::int64_t n, Cloud& c, bool octree_layout = false) {
    c.points.resize((std::size_t)n);
    for (std::int64_t i = 0; i < n; ++i) {
        PointType p;
        p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3];
        c.points[(std::size_t)i] = p;
    }
    c.width = octree_layout ? 1u : (std::uint32_t)n;
    c.height = octree_layout ? (std::uint32_t)n : 1u;
}
// keyframe scans of a session (Session::keyframe_scans_, Session.h:39) -> one packed buffer + K+1 offsets
template <class CloudPtrVector>
inline void pack_scans(const CloudPtrVector& scans, std::vector<float>& xyzi, std::vector<std::int64_t>& offsets) {
    offsets.assign(1, 0);
    xyzi.clear();
    std::vector<float> one;
    for (const auto& s : scans) {
        pack(*s, one);
        xyzi.insert(xyzi.end(), one.begin(), one.end());
        offsets.push_back(offsets.back() + (std::int64_t)s->points.size());
    }
}
// Eigen::Matrix4d -> the ABI's row-major double[16] (the layout of the pose files, Session.cpp:109)
inline void to_row_major(const Eigen::Matrix4d& m, double* out16) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) out16[4 * r + c] = m(r, c);
}
template <class PoseVector>   // std::vector<Eigen::Matrix4d> (aligned allocator or not)
inline void pack_poses(const PoseVector& poses, std::vector<double>& out) {
    out.resize(16 * poses.size());
    for (std::size_t k = 0; k < poses.size(); ++k) to_row_major(poses[k], &out[16 * k]);
}

// ---- device round trips through libltr_b200.so ----
inline int to_device(ltr_ctx* ctx, const Cloud& c, ltr_cloud* out) {
    std::vector<float> buf;
    pack(c, buf);
    return ltr_cloud_upload(ctx, buf.data(), (std::int64_t)c.points.size(), out);
}
inline int to_host(ltr_ctx* ctx, ltr_cloud h, Cloud& c, bool octree_layout = false) {
    std::int64_t n = 0;
    int rc = ltr_cloud_size(ctx, h, &n);
    if (rc != LTR_OK) return rc;
    std::vector<float> buf(4 * (std::size_t)(n > 0 ? n : 1));
    rc = ltr_cloud_download(ctx, h, buf.data(), n, &n);
    if (rc == LTR_OK) unpack(buf.data(), n, c, octree_layout);
    return rc;
}

// ---- wholesale replacement: what Removerter::loadKeyframes leaves in a Session goes to ltrh_load_session ----
// scans: Session::keyframe_scans_, poses / inverse poses: Session::keyframe_poses_ / keyframe_inverse_poses_ (Session.h:33-40),
// passed through unchanged so that the node's own Eigen inverse is the one used on the device.
template <class CloudPtrVector, class PoseVector>
inline int load_session(ltrh_removerter* r, int sess, const CloudPtrVector& scans, const PoseVector& poses, const PoseVector& inverse_poses) {
    std::vector<float> xyzi;
    std::vector<std::int64_t> off;
    std::vector<double> P, IP;
    pack_scans(scans, xyzi, off);
    pack_poses(poses, P);
    pack_poses(inverse_poses, IP);
    if (xyzi.empty()) xyzi.resize(4);
    return ltrh_load_session(r, sess, xyzi.data(), off.data(), P.data(), IP.data(), (std::int32_t)scans.size());
}
// a named device cloud of the orchestrator ("saved:nd_map", "map_global_curr_static_", ...) back into a PCL cloud
inline int fetch_cloud(ltrh_removerter* r, const char* name, int sess, Cloud& out, bool octree_layout = true) {
    ltr_cloud h;
    const int rc = ltrh_cloud(r, name, sess, &h);
    return rc != LTR_OK ? rc : to_host(ltrh_context(r), h, out, octree_layout);
}

// a named per-keyframe scan set ("keyframe_scans_updated_", ...) back into the vector of clouds the Session holds (Session.h:39-60)
template <class CloudPtrVector>
inline int fetch_scans(ltrh_removerter* r, const char* name, int sess, CloudPtrVector& out) {
    ltr_scanset h;
    int rc = ltrh_scanset(r, name, sess, &h);
    if (rc != LTR_OK) return rc;
    std::int32_t K = 0;
    std::int64_t total = 0;
    rc = ltr_scanset_info(ltrh_context(r), h, &K, &total);
    if (rc != LTR_OK) return rc;
    std::vector<float> buf(4 * (std::size_t)(total > 0 ? total : 1));
    std::vector<std::int64_t> off((std::size_t)K + 1, 0);
    rc = ltr_scanset_download(ltrh_context(r), h, buf.data(), total, off.data());
    if (rc != LTR_OK) return rc;
    out.clear();
    for (std::int32_t k = 0; k < K; ++k) {
        typename CloudPtrVector::value_type c(new Cloud());
        unpack(buf.data() + 4 * off[(std::size_t)k], off[(std::size_t)k + 1] - off[(std::size_t)k], *c);
        out.push_back(c);
    }
    return LTR_OK;
}

}  // namespace ltr_pcl
