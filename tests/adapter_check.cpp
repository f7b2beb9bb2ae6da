// Compiled, linked and run by tests/test_abi.py::test_pcl_adapter_builds_against_stand_in_headers: include/ltr_pcl_adapter.hpp against the
// stand-in PCL / Eigen headers of oracle/ref_shim (the real ones do not exist here).  Host-side conversions only; no device call is made.
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Dense>
#include "ltr_pcl_adapter.hpp"
#include <cstdio>
int main() {
    ltr_pcl::Cloud c;
    for (int i = 0; i < 5; ++i) { ltr_pcl::PointType p; p.x = i; p.y = 2 * i; p.z = -i; p.intensity = 0.5f * i; c.push_back(p); }
    std::vector<float> buf;
    ltr_pcl::pack(c, buf);
    ltr_pcl::Cloud d;
    ltr_pcl::unpack(buf.data(), 5, d, true);
    bool ok = d.points.size() == 5 && d.width == 1 && d.height == 5;
    for (int i = 0; i < 5; ++i) ok = ok && d.points[i].x == c.points[i].x && d.points[i].intensity == c.points[i].intensity;
    std::vector<ltr_pcl::Cloud::Ptr> scans;
    scans.push_back(ltr_pcl::Cloud::Ptr(new ltr_pcl::Cloud(c)));
    scans.push_back(ltr_pcl::Cloud::Ptr(new ltr_pcl::Cloud()));
    std::vector<float> xyzi; std::vector<std::int64_t> off;
    ltr_pcl::pack_scans(scans, xyzi, off);
    ok = ok && off.size() == 3 && off[1] == 5 && off[2] == 5 && xyzi.size() == 20;
    Eigen::Matrix4d m = Eigen::Matrix4d::Identity(); m(0, 3) = 7.0; m(2, 1) = -3.0;
    double rm[16]; ltr_pcl::to_row_major(m, rm);
    ok = ok && rm[3] == 7.0 && rm[9] == -3.0 && rm[15] == 1.0;
    std::printf(ok ? "adapter ok\n" : "adapter BROKEN\n");
    return ok ? 0 : 1;
}
