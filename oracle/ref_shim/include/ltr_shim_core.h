// ORACLE / REFERENCE SHIM (test infrastructure, not product code).
//
// Minimal stand-ins for the third-party headers the ltremovert translation units include (ROS, PCL, FLANN via PCL,
// Eigen, OpenCV, cv_bridge, image_transport, tf, Boost).  None of those libraries is installed in this image and there
// is no network, so the reference cannot be linked against the real ones.  With these stand-ins the reference's OWN
// sources (ltremovert/src/{utility,RosParamServer,Session,Removerter}.cpp, compiled unmodified from where they lie under
// /root/reference by oracle/Makefile -> oracle/_ref/) build and run here, which pins every line of first-party logic the
// oracle restates: pixel indexing, range-image min selection, discrepancy test, index set arithmetic, the linspace
// quirk, the kNN-threshold test, keyframe parsing, the run() order.
//
// What is NOT pinned by this: the behaviour of the third-party algorithms themselves.  The stand-ins implement
//   pcl::transformPointCloud, OctreePointCloudVoxelCentroid, KdTreeFLANN::nearestKSearch, Eigen::Matrix4d::inverse
// by calling the oracle's restatements (oracle/ref_math.h, oracle.cpp, kdtree.h), and pcl::VoxelGrid / pcl::io PCD
// from a restatement written here.  Their headers say which published behaviour they follow.
//
// Types only carry the members the reference touches.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <type_traits>
#include <unistd.h>
#include <utility>
#include <vector>

#include "../../oracle.h"   // restated third-party algorithms (ltr_oracle::transformPoint, octreeDownsampling, KdTree, inverse4x4)

// ------------------------------------------------------------------------------------------------ shim state
namespace ltr_shim {
struct ParamValue { std::vector<double> nums; std::string str; };
std::map<std::string, ParamValue>& params();                 // what rosparam / the launch file's yaml would hold
int& transform_order();                                      // which PCL summation order to stand in for (ref_math.h)
int& verbose();                                              // 1: ROS_INFO_STREAM goes to stdout
struct SavedCloud { std::string path; std::vector<float> xyzi; };
std::vector<SavedCloud>& saved();                            // every pcl::io::savePCDFileBinary call, in order
int& write_files();                                          // 1: savePCDFileBinary also writes the file
void fatal(const char* what);                                // prints and aborts: a stand-in reached a path it does not model
}  // namespace ltr_shim

// ------------------------------------------------------------------------------------------------ boost
namespace boost {
using std::make_shared;
using std::shared_ptr;
}  // namespace boost

// ------------------------------------------------------------------------------------------------ Eigen
namespace Eigen {
enum { ColMajor = 0, RowMajor = 1, Dynamic = -1 };

template <class S>
struct Mat4T {
    S m[16];   // row-major
    Mat4T() { for (int i = 0; i < 16; ++i) m[i] = S(0); }
    S& operator()(int r, int c) { return m[r * 4 + c]; }
    const S& operator()(int r, int c) const { return m[r * 4 + c]; }
    static Mat4T Identity() { Mat4T r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = S(1); return r; }
    // Eigen::Matrix4d::inverse(): stands in through the oracle's cofactor restatement (oracle.cpp inverse4x4)
    Mat4T inverse() const {
        static_assert(std::is_same<S, double>::value, "only Matrix4d::inverse is used by the reference");
        ltr_oracle::Mat4 a; std::memcpy(a.m, m, sizeof(m));
        const ltr_oracle::Mat4 r = ltr_oracle::inverse4x4(a);
        Mat4T o; std::memcpy(o.m, r.m, sizeof(m)); return o;
    }
};
typedef Mat4T<double> Matrix4d;
typedef Mat4T<float> Matrix4f;

template <class S, int R, int C, int Opt = 0> struct Matrix {};
template <class M> struct Map;
template <> struct Map<const Matrix<double, -1, -1, RowMajor>> {
    const double* p; int r, c;
    Map(const double* p_, int r_, int c_) : p(p_), r(r_), c(c_) {}
    operator Matrix4d() const {
        if (r != 4 || c != 4) ltr_shim::fatal("Eigen::Map: only 4x4 row-major maps are modelled");
        Matrix4d o; std::memcpy(o.m, p, sizeof(o.m)); return o;
    }
};
}  // namespace Eigen

// ------------------------------------------------------------------------------------------------ ROS
namespace ros {
struct Time { double t = 0; static Time now() { return Time(); } double toSec() const { return t; } };
struct TransportHints { TransportHints& tcpNoDelay(bool = true) { return *this; } };
struct Subscriber {};
struct Publisher {
    template <class M> void publish(const M&) const {}
    int getNumSubscribers() const { return 0; }
};
struct NodeHandle {
    template <class T> bool param(const std::string& name, T& var, const T& def) const {
        auto it = ltr_shim::params().find(name);
        if (it == ltr_shim::params().end()) { var = def; return false; }
        const ltr_shim::ParamValue& v = it->second;
        if constexpr (std::is_same<T, std::string>::value) var = v.str;
        else if constexpr (std::is_same<T, std::vector<float>>::value) var.assign(v.nums.begin(), v.nums.end());
        else if constexpr (std::is_same<T, std::vector<double>>::value) var = v.nums;
        else if constexpr (std::is_same<T, bool>::value) var = !v.nums.empty() && v.nums[0] != 0.0;
        else var = v.nums.empty() ? def : T(v.nums[0]);
        return true;
    }
    template <class M, class... A> Subscriber subscribe(A&&...) { return Subscriber(); }
    template <class M> Publisher advertise(const std::string&, int) { return Publisher(); }
};
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
}  // namespace ros

#define ROS_INFO(...) do { if (ltr_shim::verbose()) { std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)
#define ROS_INFO_STREAM(x) do { if (ltr_shim::verbose()) { std::cout << x << std::endl; } } while (0)
#define ROS_WARN_STREAM(x) ROS_INFO_STREAM(x)
#define ROS_ERROR_STREAM(x) ROS_INFO_STREAM(x)

namespace std_msgs { struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; }; }
namespace sensor_msgs {
struct PointCloud2 { std_msgs::Header header; };
typedef boost::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
struct Image { std_msgs::Header header; int height = 0, width = 0; };
typedef boost::shared_ptr<Image> ImagePtr;
}  // namespace sensor_msgs
namespace image_transport {
struct Publisher { void publish(const sensor_msgs::ImagePtr&) const {} };
struct ImageTransport {
    explicit ImageTransport(const ros::NodeHandle&) {}
    Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace image_transport

// ------------------------------------------------------------------------------------------------ OpenCV
#define CV_8UC1 0
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_8UC3 16
namespace cv {
struct Scalar {
    double v;
    Scalar(double x = 0) : v(x) {}
    static Scalar all(double x) { return Scalar(x); }
};
enum { COLORMAP_JET = 2 };
struct Mat {
    int rows = 0, cols = 0, type_ = CV_32FC1;
    std::shared_ptr<std::vector<unsigned char>> buf;   // shared on copy like cv::Mat's ref-counted header
    static int elem(int t) { return t == CV_8UC1 ? 1 : t == CV_8UC3 ? 3 : 4; }
    Mat() {}
    Mat(int r, int c, int t, const Scalar& s = Scalar()) : rows(r), cols(c), type_(t), buf(new std::vector<unsigned char>((size_t)r * c * elem(t))) {
        for (int i = 0; i < r * c; ++i) set(i, s.v);
    }
    int type() const { return type_; }
    double get(int i) const {
        if (type_ == CV_32FC1) return ((const float*)buf->data())[i];
        if (type_ == CV_32SC1) return ((const int*)buf->data())[i];
        return buf->data()[(size_t)i * elem(type_)];
    }
    void set(int i, double v) {
        if (type_ == CV_32FC1) ((float*)buf->data())[i] = (float)v;
        else if (type_ == CV_32SC1) ((int*)buf->data())[i] = (int)std::nearbyint(v);
        else { const double c = std::min(255.0, std::max(0.0, std::nearbyint(v))); for (int k = 0; k < elem(type_); ++k) buf->data()[(size_t)i * elem(type_) + k] = (unsigned char)c; }
    }
    template <class T> T& at(int r, int c) { return ((T*)buf->data())[(size_t)r * cols + c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)buf->data())[(size_t)r * cols + c]; }
    Mat clone() const { Mat o = *this; if (buf) o.buf.reset(new std::vector<unsigned char>(*buf)); return o; }
    void convertTo(Mat& dst, int t) const { Mat o(rows, cols, t); for (int i = 0; i < rows * cols; ++i) o.set(i, get(i)); dst = o; }
};
// float images: elementwise IEEE f32 subtraction (cv::subtract on CV_32F); other types only feed the rviz colour maps
inline Mat operator-(const Mat& a, const Mat& b) {
    if (a.rows != b.rows || a.cols != b.cols || a.type_ != CV_32FC1 || b.type_ != CV_32FC1) ltr_shim::fatal("cv::Mat - cv::Mat: only equal-size CV_32FC1 is modelled");
    Mat o(a.rows, a.cols, CV_32FC1);
    const float* pa = (const float*)a.buf->data(); const float* pb = (const float*)b.buf->data(); float* po = (float*)o.buf->data();
    for (int i = 0; i < a.rows * a.cols; ++i) po[i] = pa[i] - pb[i];
    return o;
}
inline Mat operator-(const Mat& a, double s) { Mat o(a.rows, a.cols, a.type_); for (int i = 0; i < a.rows * a.cols; ++i) o.set(i, a.get(i) - s); return o; }
inline Mat operator*(double s, const Mat& a) { Mat o(a.rows, a.cols, a.type_); for (int i = 0; i < a.rows * a.cols; ++i) o.set(i, a.get(i) * s); return o; }
inline Mat operator/(const Mat& a, double s) { Mat o(a.rows, a.cols, a.type_); for (int i = 0; i < a.rows * a.cols; ++i) o.set(i, a.get(i) / s); return o; }
inline void applyColorMap(const Mat& src, Mat& dst, int) { Mat o(src.rows, src.cols, CV_8UC3); for (int i = 0; i < src.rows * src.cols; ++i) o.set(i, src.get(i)); dst = o; }
}  // namespace cv
namespace cv_bridge {
struct CvImage {
    cv::Mat image;
    CvImage(const std_msgs::Header&, const std::string&, const cv::Mat& img) : image(img) {}
    sensor_msgs::ImagePtr toImageMsg() const { sensor_msgs::ImagePtr p(new sensor_msgs::Image()); p->height = image.rows; p->width = image.cols; return p; }
};
}  // namespace cv_bridge

// ------------------------------------------------------------------------------------------------ PCL
namespace pcl {
struct PointXYZI {
    float x, y, z, intensity;
    PointXYZI() : x(0), y(0), z(0), intensity(0) {}
};
static_assert(sizeof(PointXYZI) == sizeof(ltr_oracle::Pt), "PointXYZI stand-in must alias ltr_oracle::Pt");

template <class P>
struct PointCloud {
    typedef boost::shared_ptr<PointCloud<P>> Ptr;
    typedef boost::shared_ptr<const PointCloud<P>> ConstPtr;
    std::vector<P> points;
    std::uint32_t width = 0, height = 0;
    bool is_dense = true;
    void push_back(const P& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
    void clear() { points.clear(); width = height = 0; }
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    P& operator[](std::size_t i) { return points[i]; }
    const P& operator[](std::size_t i) const { return points[i]; }
    PointCloud& operator+=(const PointCloud& o) {
        const std::vector<P> tmp = o.points;   // (self-append safe)
        points.insert(points.end(), tmp.begin(), tmp.end());
        width = (std::uint32_t)points.size(); height = 1;
        return *this;
    }
};

namespace console { enum VERBOSITY_LEVEL { L_ALWAYS, L_ERROR, L_WARN, L_INFO, L_DEBUG, L_VERBOSE }; inline void setVerbosityLevel(VERBOSITY_LEVEL) {} }

// pcl::transformPointCloud(cloud_in, cloud_out, Eigen::Matrix4d): xyz in double, rounded to float, intensity copied,
// in-place allowed.  Stand-in = ltr_oracle::transformPoint (ref_math.h), summation order = ltr_shim::transform_order().
template <class P>
void transformPointCloud(const PointCloud<P>& in, PointCloud<P>& out, const Eigen::Matrix4d& T) {
    if (&in != &out) { out.points.resize(in.points.size()); out.width = in.width; out.height = in.height; out.is_dense = in.is_dense; }
    const int order = ltr_shim::transform_order();
    for (std::size_t i = 0; i < in.points.size(); ++i) {
        const P p = in.points[i];
        P q = p;
        ltr_oracle::transformPoint(T.m, order, p.x, p.y, p.z, &q.x, &q.y, &q.z);
        out.points[i] = q;
    }
}
template <class P>
void transformPointCloud(const PointCloud<P>&, PointCloud<P>&, const Eigen::Matrix4f&) { ltr_shim::fatal("transformPointCloud(Matrix4f) (ICP refinement branch, disabled in the reference) is not modelled"); }

// pcl::KdTreeFLANN<PointXYZI>: exact kNN on xyz, squared L2 distances ascending, k clamped to the indexed count.
// Stand-in = ltr_oracle::KdTree (kdtree.h).  Neighbour indices are not modelled (the reference never reads them).
template <class P>
struct KdTreeFLANN {
    typedef boost::shared_ptr<KdTreeFLANN<P>> Ptr;
    ltr_oracle::KdTree tree;
    void setInputCloud(const typename PointCloud<P>::ConstPtr& cloud) { tree.build(cloud->points.empty() ? nullptr : &cloud->points[0].x, (int)(sizeof(P) / sizeof(float)), (int)cloud->points.size()); }
    int nearestKSearch(const P& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
        if (tree.n == 0) ltr_shim::fatal("KdTreeFLANN::nearestKSearch on an empty tree (PCL asserts here)");
        const int kk = std::min(k, tree.n);
        idx.assign(kk, -1);
        d2.resize(kk);
        const float qq[3] = {q.x, q.y, q.z};
        tree.knn(qq, kk, d2.data());
        return kk;
    }
};

// pcl::octree::OctreePointCloudVoxelCentroid<PointXYZI>.  Stand-in = ltr_oracle::octreeDownsampling (oracle.cpp).
namespace octree {
template <class P>
struct OctreePointCloudVoxelCentroid {
    typedef std::vector<P> AlignedPointTVector;
    double res;
    typename PointCloud<P>::ConstPtr input;
    explicit OctreePointCloudVoxelCentroid(double r) : res(r) {}
    void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { input = c; }
    void defineBoundingBox() {}
    void addPointsFromInputCloud() {}
    std::size_t getVoxelCentroids(AlignedPointTVector& out) const {
        ltr_oracle::Cloud src(input->points.size()), dst;
        if (!src.empty()) std::memcpy(src.data(), input->points.data(), src.size() * sizeof(P));
        if (ltr_oracle::octreeDownsampling(src, dst, (float)res) < 0) ltr_shim::fatal("octree deeper than the restatement supports");
        out.resize(dst.size());
        if (!dst.empty()) std::memcpy(out.data(), dst.data(), dst.size() * sizeof(P));
        return out.size();
    }
};
}  // namespace octree

// pcl::VoxelGrid<PointXYZI> (pcl/filters/impl/voxel_grid.hpp applyFilter, all fields, no min-points limit): leaf index from
// floor(coord * inverse_leaf) - min_b, std::sort on the leaf index, per-leaf centroid of x, y, z, intensity in float.
template <class P>
struct VoxelGrid {
    float leaf[3] = {0, 0, 0};
    typename PointCloud<P>::ConstPtr input;
    void setLeafSize(float lx, float ly, float lz) { leaf[0] = lx; leaf[1] = ly; leaf[2] = lz; }
    void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { input = c; }
    void filter(PointCloud<P>& out) const {
        const std::vector<P>& in = input->points;
        out.clear();
        if (in.empty()) return;
        const float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};
        float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
        for (const P& p : in) {
            mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
            mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
        }
        std::int64_t d[3];
        for (int a = 0; a < 3; ++a) d[a] = (std::int64_t)((mx[a] - mn[a]) * inv[a]) + 1;
        if (d[0] * d[1] * d[2] > (std::int64_t)std::numeric_limits<std::int32_t>::max()) { out = *input; return; }   // "Leaf size is too small": output = input
        int min_b[3], div_b[3];
        for (int a = 0; a < 3; ++a) { min_b[a] = (int)std::floor(mn[a] * inv[a]); div_b[a] = (int)std::floor(mx[a] * inv[a]) - min_b[a] + 1; }
        struct Entry { unsigned idx, pt; bool operator<(const Entry& o) const { return idx < o.idx; } };
        std::vector<Entry> e;
        e.reserve(in.size());
        for (std::size_t i = 0; i < in.size(); ++i) {
            const int i0 = (int)(std::floor(in[i].x * inv[0]) - (float)min_b[0]);
            const int i1 = (int)(std::floor(in[i].y * inv[1]) - (float)min_b[1]);
            const int i2 = (int)(std::floor(in[i].z * inv[2]) - (float)min_b[2]);
            e.push_back(Entry{(unsigned)(i0 + i1 * div_b[0] + i2 * div_b[0] * div_b[1]), (unsigned)i});
        }
        std::sort(e.begin(), e.end(), std::less<Entry>());
        for (std::size_t i = 0; i < e.size();) {
            std::size_t j = i;
            float s[4] = {0, 0, 0, 0};
            for (; j < e.size() && e[j].idx == e[i].idx; ++j) { const P& p = in[e[j].pt]; s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.intensity; }
            const float n = (float)(j - i);
            P c; c.x = s[0] / n; c.y = s[1] / n; c.z = s[2] / n; c.intensity = s[3] / n;
            out.push_back(c);
            i = j;
        }
    }
};

template <class P>
struct ExtractIndices {
    typename PointCloud<P>::ConstPtr input;
    boost::shared_ptr<std::vector<int>> indices;
    bool negative = false;
    void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { input = c; }
    void setIndices(const boost::shared_ptr<std::vector<int>>& i) { indices = i; }
    void setNegative(bool n) { negative = n; }
    void filter(PointCloud<P>& out) const {
        if (negative) ltr_shim::fatal("ExtractIndices::setNegative(true) is not modelled (unused by the reference)");
        std::vector<P> r;
        r.reserve(indices->size());
        for (int i : *indices) r.push_back(input->points.at((std::size_t)i));   // PCL copies the indexed points in index order
        out.points.swap(r);
        out.width = (std::uint32_t)out.points.size(); out.height = 1;
    }
};

template <class S, class T>
struct IterativeClosestPoint {   // only reachable with useICPrefinement = true (hard-coded false, Session.cpp:551)
    void setMaxCorrespondenceDistance(double) {}
    void setMaximumIterations(int) {}
    void setTransformationEpsilon(double) {}
    void setEuclideanFitnessEpsilon(double) {}
    void setRANSACIterations(int) {}
    void setInputTarget(const typename PointCloud<T>::ConstPtr&) {}
    void setInputSource(const typename PointCloud<S>::ConstPtr&) {}
    void align(PointCloud<S>&) { ltr_shim::fatal("ICP refinement is not modelled (disabled in the reference)"); }
    Eigen::Matrix4f getFinalTransformation() const { return Eigen::Matrix4f::Identity(); }
    double getFitnessScore() const { return 0.0; }
};

template <class P> void toROSMsg(const PointCloud<P>&, sensor_msgs::PointCloud2&) {}

namespace io {
int load_pcd_xyzi(const std::string& path, std::vector<float>& xyzi);      // ref_shim.cpp
int save_pcd_xyzi(const std::string& path, const float* xyzi, std::size_t n);
template <class P> int loadPCDFile(const std::string& path, PointCloud<P>& cloud) {
    std::vector<float> v;
    if (load_pcd_xyzi(path, v) != 0) return -1;
    cloud.points.resize(v.size() / 4);
    if (!v.empty()) std::memcpy(cloud.points.data(), v.data(), v.size() * sizeof(float));
    cloud.width = (std::uint32_t)cloud.points.size(); cloud.height = 1;
    return 0;
}
template <class P> int savePCDFileBinary(const std::string& path, const PointCloud<P>& cloud) {
    return save_pcd_xyzi(path, cloud.points.empty() ? nullptr : &cloud.points[0].x, cloud.points.size());
}
}  // namespace io
}  // namespace pcl
