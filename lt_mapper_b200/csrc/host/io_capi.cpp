// Flat C view of the file-level helpers (io.h) for the Python test harness; declared in include/ltr_removert.h.
#include "io.h"
#include "../../../include/ltr_removert.h"
#include <cstring>

using namespace ltremovert_b200;
static thread_local std::string g_io_err;

extern "C" {

const char* ltrh_io_last_error(void) { return g_io_err.c_str(); }

int64_t ltrh_io_read_pcd(const char* path, float* xyzi, int64_t capacity) {
    HostCloud c;
    if (!read_pcd(path, &c, &g_io_err)) return -1;
    if ((int64_t)c.size() <= capacity && !c.empty()) std::memcpy(xyzi, c.data(), c.size() * sizeof(PointXYZI));
    return (int64_t)c.size();
}
int ltrh_io_write_pcd(const char* path, const float* xyzi, int64_t n, int32_t octree_layout) {
    HostCloud c((size_t)n);
    if (n) std::memcpy(c.data(), xyzi, (size_t)n * sizeof(PointXYZI));
    return write_pcd_binary(path, c, octree_layout != 0, &g_io_err) ? 0 : -1;
}
int32_t ltrh_io_read_poses(const char* path, double* poses16, int32_t capacity) {
    std::vector<Mat4> p;
    if (!read_pose_file(path, &p, &g_io_err)) return -1;
    for (size_t i = 0; i < p.size() && (int32_t)i < capacity; ++i) std::memcpy(poses16 + 16 * i, p[i].data(), 16 * sizeof(double));
    return (int32_t)p.size();
}
int32_t ltrh_io_parse_keyframes(int32_t num_scans, int32_t start_idx, int32_t end_idx, int32_t gap, int32_t* out, int32_t capacity) {
    const std::vector<int> v = parse_keyframes(num_scans, start_idx, end_idx, gap);
    for (size_t i = 0; i < v.size() && (int32_t)i < capacity; ++i) out[i] = v[i];
    return (int32_t)v.size();
}
int32_t ltrh_io_parse_keyframes_in_roi(const double* scan_poses16, int32_t n, const double* roi_poses16, int32_t m, int32_t gap, int32_t* out, int32_t capacity) {
    std::vector<Mat4> a((size_t)n), b((size_t)m);
    for (int i = 0; i < n; ++i) std::memcpy(a[i].data(), scan_poses16 + 16 * (size_t)i, 16 * sizeof(double));
    for (int i = 0; i < m; ++i) std::memcpy(b[i].data(), roi_poses16 + 16 * (size_t)i, 16 * sizeof(double));
    const std::vector<int> v = parse_keyframes_in_roi(a, b, gap);
    for (size_t i = 0; i < v.size() && (int32_t)i < capacity; ++i) out[i] = v[i];
    return (int32_t)v.size();
}
int64_t ltrh_io_voxel_grid(const float* xyzi, int64_t n, float leaf, float* out, int64_t capacity, int32_t* overflowed) {
    HostCloud c((size_t)n);
    if (n) std::memcpy(c.data(), xyzi, (size_t)n * sizeof(PointXYZI));
    bool ov = false;
    const HostCloud o = voxel_grid(c, leaf, &ov);
    if (overflowed) *overflowed = ov ? 1 : 0;
    if ((int64_t)o.size() <= capacity && !o.empty()) std::memcpy(out, o.data(), o.size() * sizeof(PointXYZI));
    return (int64_t)o.size();
}
// all K scans of a session at once (VoxelGridBatch): returns the total number of output points and fills out_off[K + 1];
// `out` is written only when it is large enough (call with capacity 0 first to size it)
int64_t ltrh_io_voxel_grid_scans(const float* xyzi, const int64_t* off, int32_t K, float leaf, float* out, int64_t capacity, int64_t* out_off) {
    if (K < 0 || !off || (K > 0 && !xyzi && off[K] > 0)) return -1;
    VoxelGridBatch b;
    b.plan(xyzi, off, K, leaf);
    if (out_off) for (int k = 0; k <= K; ++k) out_off[k] = b.out_off[(size_t)k];
    if (out && b.out_off[(size_t)K] <= capacity) b.emit(xyzi, off, out);
    return b.out_off[(size_t)K];
}
int ltrh_io_yaml_get(const char* path, const char* key, char* value, int32_t capacity, double* list, int32_t list_capacity, int32_t* list_n) {
    YamlParams y;
    if (!y.load(path, &g_io_err)) return -1;
    if (value && capacity > 0) { std::snprintf(value, (size_t)capacity, "%s", y.str(key, "").c_str()); }
    const std::vector<double> l = y.list(key);
    if (list_n) *list_n = (int32_t)l.size();
    for (size_t i = 0; i < l.size() && (int32_t)i < list_capacity; ++i) list[i] = l[i];
    return 0;
}

}  // extern "C"
