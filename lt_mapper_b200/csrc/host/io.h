// File-level surface of the ltremovert node: yaml parameters, pose files, binary PCD
// scans, keyframe selection and the output tree -- plain C++17, no ROS/PCL.  Mirrors:
//   RosParamServer::RosParamServer      ltremovert/src/RosParamServer.cpp:4-63   (yaml keys, defaults)
//   Session::loadSessionInfo            ltremovert/src/Session.cpp:80-118        (sorted directory listing, pose lines)
//   splitPoseLine                       ltremovert/src/utility.cpp:28-36
//   Session::parseKeyframes / InROI     ltremovert/src/Session.cpp:138-174, 230-263
//   Session::loadKeyframes              ltremovert/src/Session.cpp:266-302       (loadPCDFile + VoxelGrid)
//   Removerter::Removerter / save*      ltremovert/src/Removerter.cpp:17-73, 231, 318-338, 1446-1477, 1517-1520, 1600-1650
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace ltremovert_b200 {

struct PointXYZI { float x, y, z, intensity; };
typedef std::vector<PointXYZI> HostCloud;
typedef std::array<double, 16> Mat4;  // row-major

// ---- yaml (the subset roslaunch's rosparam load needs for params_ltmapper.yaml) ----
struct YamlParams {
    std::map<std::string, std::string> scalars;               // "removert/key" -> raw text
    std::map<std::string, std::vector<double>> lists;         // "removert/key" -> numbers
    bool load(const std::string& path, std::string* err);
    std::string str(const std::string& key, const std::string& def) const;
    double num(const std::string& key, double def) const;
    bool boolean(const std::string& key, bool def) const;
    std::vector<double> list(const std::string& key) const;
};

// ---- poses ----
// One pose per line, 12 (3x4) or 16 numbers separated by single spaces (utility.cpp:28-36, Session.cpp:102-114).
// Empty tokens (consecutive spaces) make std::stod throw in the reference; here they are skipped.
bool read_pose_file(const std::string& path, std::vector<Mat4>* poses, std::string* err);
Mat4 inverse(const Mat4& m);  // general 4x4 inverse, stands in for Eigen::Matrix4d::inverse() (Session.cpp:110)

// ---- PCD ----
// Reads x, y, z (+ intensity if present) from an ascii or binary PCD with 4-byte float fields (what SC-LIO-SAM /
// pcl::io::savePCDFileBinary write); other fields are skipped.  DATA ascii | binary | binary_compressed (LZF, field-major payload).
bool read_pcd(const std::string& path, HostCloud* out, std::string* err);
// pcl::io::savePCDFileBinary<PointXYZI>: FIELDS x y z intensity, 16 B/point.  width/height as the reference sets them
// (octreeDownsampling sets width = 1, height = n, utility.cpp:217-218; everything else width = n, height = 1).
bool write_pcd_binary(const std::string& path, const HostCloud& c, bool octree_layout, std::string* err);

// pcl::VoxelGrid<PointXYZI> with one leaf size (Session.cpp:284-289), PCL 1.10 semantics restated (UNPINNED):
// int64 overflow check -> returns the input unchanged (with the reference's warning) when
// (dx*dy*dz) > INT32_MAX, which is the common case for 0.05 m leaves on outdoor scans; otherwise voxel index sort
// (std::sort on idx, as PCL) and float centroid accumulation in sorted order.
HostCloud voxel_grid(const HostCloud& in, float leaf, bool* overflowed);
// The same filter over the K scans of a session held back to back (interleaved x y z intensity, scan k = points
// [off[k], off[k+1])): what Session::loadKeyframes does scan by scan (Session.cpp:272-303).  Two phases so that the caller can
// size (page-locked) destination memory in between; scans run on different OpenMP threads, and a scan that takes PCL's
// overflow exit is passed through with one memcpy instead of being rebuilt point by point.
struct VoxelGridBatch {
    std::vector<HostCloud> grids;       // filtered scans (empty for pass-through scans)
    std::vector<uint8_t> unchanged;     // 1: overflow exit (or empty scan), output = input
    std::vector<int64_t> out_off;       // K + 1 point offsets of the result
    void plan(const float* xyzi, const int64_t* off, int K, float leaf);
    void emit(const float* xyzi, const int64_t* off, float* out) const;   // out: out_off[K] points
};

// ---- session bookkeeping ----
struct SessionFiles {
    std::vector<std::string> scan_names, scan_paths;   // sorted (Session.cpp:87-92)
    std::vector<Mat4> scan_poses, scan_inverse_poses;
    std::vector<int> keyframe_idx;                     // indices into the above
};
bool list_session(const std::string& scan_dir, const std::string& pose_path, SessionFiles* s, std::string* err);
// Session::parseKeyframes(range, gap) including the reference's quirk: an out-of-range index also skips the next one
// (the extra curr_idx++ at Session.cpp:150), and `remainder(num_valid_parsed, gap) != 0` gap test.
std::vector<int> parse_keyframes(int num_scans, int start_idx, int end_idx, int gap);
// Session::parseKeyframesInROI: scans within 10 m (xyz) of any ROI pose (Session.cpp:230-263)
std::vector<int> parse_keyframes_in_roi(const std::vector<Mat4>& scan_poses, const std::vector<Mat4>& roi_poses, int gap);

}  // namespace ltremovert_b200
