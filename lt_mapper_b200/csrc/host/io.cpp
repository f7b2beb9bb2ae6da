// File-level surface of the ltremovert node -- see io.h for the reference lines mirrored.
#include "io.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <limits>
#include <sstream>

namespace fs = std::filesystem;

namespace ltremovert_b200 {

// ------------------------------------------------------------------------------------------------ yaml
static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
static std::string strip_comment(const std::string& s) {
    bool in_q = false;
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        if (in_q) { if (s[i] == q) in_q = false; }
        else if (s[i] == '"' || s[i] == '\'') { in_q = true; q = s[i]; }
        else if (s[i] == '#' && (i == 0 || std::isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
    }
    return s;
}
static std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
    return s;
}

bool YamlParams::load(const std::string& path, std::string* err) {
    std::ifstream f(path);
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::string line, ns, pending_key, pending_list;
    bool in_list = false;
    auto finish_list = [&](const std::string& key, const std::string& body) {
        std::vector<double> v;
        std::string tok;
        std::stringstream ss(body);
        while (std::getline(ss, tok, ',')) { tok = trim(tok); if (!tok.empty()) v.push_back(std::stod(tok)); }
        lists[key] = v;
    };
    while (std::getline(f, line)) {
        line = strip_comment(line);
        if (in_list) {
            const size_t e = line.find(']');
            pending_list += (e == std::string::npos) ? line + " " : line.substr(0, e);
            if (e != std::string::npos) { finish_list(pending_key, pending_list); in_list = false; }
            continue;
        }
        if (trim(line).empty()) continue;
        const size_t indent = line.find_first_not_of(' ');
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = trim(line.substr(0, colon));
        const std::string val = trim(line.substr(colon + 1));
        if (indent == 0 && val.empty()) { ns = key + "/"; continue; }
        const std::string full = (indent == 0 ? std::string() : ns) + key;
        if (!val.empty() && val[0] == '[') {
            const size_t e = val.find(']');
            if (e != std::string::npos) finish_list(full, val.substr(1, e - 1));
            else { in_list = true; pending_key = full; pending_list = val.substr(1) + " "; }
        } else {
            scalars[full] = unquote(val);
        }
    }
    return true;
}
std::string YamlParams::str(const std::string& key, const std::string& def) const { auto it = scalars.find(key); return it == scalars.end() ? def : it->second; }
double YamlParams::num(const std::string& key, double def) const { auto it = scalars.find(key); return it == scalars.end() ? def : std::stod(it->second); }
bool YamlParams::boolean(const std::string& key, bool def) const {
    auto it = scalars.find(key);
    if (it == scalars.end()) return def;
    std::string v = it->second;
    std::transform(v.begin(), v.end(), v.begin(), ::tolower);
    return v == "true" || v == "1" || v == "yes" || v == "on";
}
std::vector<double> YamlParams::list(const std::string& key) const { auto it = lists.find(key); return it == lists.end() ? std::vector<double>() : it->second; }

// ------------------------------------------------------------------------------------------------ poses
bool read_pose_file(const std::string& path, std::vector<Mat4>* poses, std::string* err) {
    std::ifstream f(path);
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::string line;
    poses->clear();
    while (std::getline(f, line)) {
        std::vector<double> v;
        std::stringstream ss(line);
        std::string tok;
        while (std::getline(ss, tok, ' ')) { if (trim(tok).empty()) continue; v.push_back(std::stod(tok)); }  // utility.cpp:28-36
        if (v.empty()) continue;
        if (v.size() == 12) { v.push_back(0.0); v.push_back(0.0); v.push_back(0.0); v.push_back(1.0); }    // Session.cpp:106-108
        if (v.size() != 16) { if (err) *err = "pose line with " + std::to_string(v.size()) + " numbers in " + path; return false; }
        Mat4 m;
        std::copy(v.begin(), v.end(), m.begin());
        poses->push_back(m);
    }
    return true;
}

Mat4 inverse(const Mat4& A) {
    const double* m = A.data();
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    const double idet = 1.0 / det;
    Mat4 R;
    for (int i = 0; i < 16; ++i) R[i] = inv[i] * idet;
    return R;
}

// ------------------------------------------------------------------------------------------------ PCD
bool read_pcd(const std::string& path, HostCloud* out, std::string* err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::vector<std::string> fields;
    std::vector<int> sizes, counts;
    std::vector<char> types;
    long width = 0, height = 1, points = -1;
    std::string data, line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::stringstream ss(line);
        std::string key;
        ss >> key;
        if (key == "FIELDS") { std::string t; while (ss >> t) fields.push_back(t); }
        else if (key == "SIZE") { int t; while (ss >> t) sizes.push_back(t); }
        else if (key == "TYPE") { char t; while (ss >> t) types.push_back(t); }
        else if (key == "COUNT") { int t; while (ss >> t) counts.push_back(t); }
        else if (key == "WIDTH") ss >> width;
        else if (key == "HEIGHT") ss >> height;
        else if (key == "POINTS") ss >> points;
        else if (key == "DATA") { ss >> data; break; }
    }
    if (fields.empty() || sizes.size() != fields.size()) { if (err) *err = "bad PCD header in " + path; return false; }
    if (counts.empty()) counts.assign(fields.size(), 1);
    if (types.size() != fields.size()) types.assign(fields.size(), 'F');
    if (points < 0) points = width * height;
    int off[4] = {-1, -1, -1, -1}, col[4] = {-1, -1, -1, -1};
    int stride = 0, ncol = 0;
    for (size_t i = 0; i < fields.size(); ++i) {
        const char* names[4] = {"x", "y", "z", "intensity"};
        for (int j = 0; j < 4; ++j)
            if (fields[i] == names[j]) {
                if (sizes[i] != 4 || types[i] != 'F') { if (err) *err = "field " + fields[i] + " is not float32 in " + path; return false; }
                off[j] = stride; col[j] = ncol;
            }
        stride += sizes[i] * counts[i];
        ncol += counts[i];
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0) { if (err) *err = "PCD without x y z in " + path; return false; }
    out->assign((size_t)points, PointXYZI{0, 0, 0, 0});
    if (data == "binary") {
        std::vector<char> buf((size_t)points * stride);
        f.read(buf.data(), (std::streamsize)buf.size());
        if ((size_t)f.gcount() != buf.size()) { if (err) *err = "truncated PCD " + path; return false; }
        for (long i = 0; i < points; ++i) {
            const char* p = buf.data() + (size_t)i * stride;
            PointXYZI q{0, 0, 0, 0};
            std::memcpy(&q.x, p + off[0], 4); std::memcpy(&q.y, p + off[1], 4); std::memcpy(&q.z, p + off[2], 4);
            if (off[3] >= 0) std::memcpy(&q.intensity, p + off[3], 4);
            (*out)[(size_t)i] = q;
        }
    } else if (data == "ascii") {
        for (long i = 0; i < points; ++i) {
            if (!std::getline(f, line)) { if (err) *err = "truncated PCD " + path; return false; }
            std::stringstream ss(line);
            std::vector<double> v;
            double t;
            while (ss >> t) v.push_back(t);
            PointXYZI q{0, 0, 0, 0};
            if ((int)v.size() <= std::max(col[0], std::max(col[1], col[2]))) { if (err) *err = "short PCD row in " + path; return false; }
            q.x = (float)v[col[0]]; q.y = (float)v[col[1]]; q.z = (float)v[col[2]];
            if (col[3] >= 0 && (int)v.size() > col[3]) q.intensity = (float)v[col[3]];
            (*out)[(size_t)i] = q;
        }
    } else if (data == "binary_compressed") {
        // pcl::PCDWriter::writeBinaryCompressed: uint32 compressed size, uint32 uncompressed size, LZF stream; the uncompressed payload is
        // stored field by field (all x, then all y, ...), each field block = points * size * count bytes
        uint32_t csize = 0, usize = 0;
        f.read(reinterpret_cast<char*>(&csize), 4); f.read(reinterpret_cast<char*>(&usize), 4);
        if (!f || usize != (uint64_t)points * (uint64_t)stride) { if (err) *err = "bad binary_compressed sizes in " + path; return false; }
        std::vector<unsigned char> cbuf(csize), ubuf(usize);
        f.read(reinterpret_cast<char*>(cbuf.data()), (std::streamsize)csize);
        if ((size_t)f.gcount() != cbuf.size()) { if (err) *err = "truncated PCD " + path; return false; }
        {   // LZF decompression (liblzf format: control byte < 32 = literal run of ctrl + 1 bytes, otherwise a back reference)
            size_t ip = 0, op = 0;
            while (ip < cbuf.size()) {
                unsigned ctrl = cbuf[ip++];
                if (ctrl < 32) {
                    ++ctrl;
                    if (ip + ctrl > cbuf.size() || op + ctrl > ubuf.size()) { if (err) *err = "corrupt LZF stream in " + path; return false; }
                    std::memcpy(&ubuf[op], &cbuf[ip], ctrl);
                    ip += ctrl; op += ctrl;
                } else {
                    size_t len = ctrl >> 5;
                    if (ip >= cbuf.size()) { if (err) *err = "corrupt LZF stream in " + path; return false; }
                    if (len == 7) { len += cbuf[ip++]; if (ip >= cbuf.size()) { if (err) *err = "corrupt LZF stream in " + path; return false; } }
                    const size_t dist = ((size_t)(ctrl & 0x1f) << 8) + cbuf[ip++] + 1;
                    len += 2;
                    if (dist > op || op + len > ubuf.size()) { if (err) *err = "corrupt LZF stream in " + path; return false; }
                    for (size_t t = 0; t < len; ++t, ++op) ubuf[op] = ubuf[op - dist];   // may overlap: byte by byte
                }
            }
            if (op != ubuf.size()) { if (err) *err = "short LZF stream in " + path; return false; }
        }
        // field blocks in header order; block start of field i = points * (bytes of the fields before it)
        for (int j = 0; j < 4; ++j) {
            if (off[j] < 0) continue;
            const unsigned char* blk = ubuf.data() + (size_t)points * (size_t)off[j];
            for (long i = 0; i < points; ++i) {
                float v;
                std::memcpy(&v, blk + (size_t)i * 4, 4);
                PointXYZI& q = (*out)[(size_t)i];
                (j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.intensity) = v;
            }
        }
    } else {
        if (err) *err = "unsupported PCD DATA '" + data + "' in " + path;
        return false;
    }
    return true;
}

bool write_pcd_binary(const std::string& path, const HostCloud& c, bool octree_layout, std::string* err) {
    std::ofstream f(path, std::ios::binary);
    if (!f) { if (err) *err = "cannot write " + path; return false; }
    const size_t n = c.size();
    std::ostringstream h;
    h << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
      << "WIDTH " << (octree_layout ? 1 : n) << "\nHEIGHT " << (octree_layout ? n : 1) << "\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    const std::string hs = h.str();
    f.write(hs.data(), (std::streamsize)hs.size());
    if (n) f.write(reinterpret_cast<const char*>(c.data()), (std::streamsize)(n * sizeof(PointXYZI)));
    return (bool)f;
}

// ------------------------------------------------------------------------------------------------ pcl::VoxelGrid
// Core of the filter on a span of points.  Returns false when PCL takes its overflow exit ("Leaf size is too small for the input
// dataset": output = *input_); `out` is then left untouched so that callers can pass the input through without a copy.
static bool voxel_grid_span(const PointXYZI* in, size_t n, float leaf, HostCloud* out) {
    const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Ones / leaf_size_
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[1], -mn[2]};
    for (size_t i = 0; i < n; ++i) {
        const PointXYZI& p = in[i];
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) return false;
    int min_b[3], max_b[3], div_b[3];
    for (int d = 0; d < 3; ++d) { min_b[d] = (int)std::floor(mn[d] * inv); max_b[d] = (int)std::floor(mx[d] * inv); div_b[d] = max_b[d] - min_b[d] + 1; }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    struct cloud_point_index_idx {
        unsigned int idx, cloud_point_index;
        bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }
    };
    std::vector<cloud_point_index_idx> iv;
    iv.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        const int ijk0 = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
        const int ijk1 = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
        const int ijk2 = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
        iv.push_back({(unsigned int)(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]), (unsigned int)i});
    }
    std::sort(iv.begin(), iv.end(), std::less<cloud_point_index_idx>());
    out->clear();
    size_t i = 0;
    while (i < iv.size()) {
        size_t j = i;
        float sx = 0, sy = 0, sz = 0, si = 0;
        while (j < iv.size() && iv[j].idx == iv[i].idx) {
            const auto& p = in[iv[j].cloud_point_index];
            sx += p.x; sy += p.y; sz += p.z; si += p.intensity;   // CentroidPoint accumulators (float)
            ++j;
        }
        const float cnt = (float)(j - i);
        out->push_back({sx / cnt, sy / cnt, sz / cnt, si / cnt});
        i = j;
    }
    return true;
}

HostCloud voxel_grid(const HostCloud& in, float leaf, bool* overflowed) {
    if (overflowed) *overflowed = false;
    if (in.empty()) return in;
    HostCloud out;
    if (!voxel_grid_span(in.data(), in.size(), leaf, &out)) {
        if (overflowed) *overflowed = true;
        return in;                                                      // output = *input_
    }
    return out;
}

void VoxelGridBatch::plan(const float* xyzi, const int64_t* off, int K, float leaf) {
    static_assert(sizeof(PointXYZI) == 4 * sizeof(float), "PointXYZI must be four packed floats");
    const PointXYZI* pts = reinterpret_cast<const PointXYZI*>(xyzi);
    grids.assign((size_t)K, HostCloud());
    unchanged.assign((size_t)K, 0);
    out_off.assign((size_t)K + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        const size_t n = (size_t)(off[k + 1] - off[k]);
        if (n == 0 || !voxel_grid_span(pts + off[k], n, leaf, &grids[(size_t)k])) unchanged[(size_t)k] = 1;
    }
    for (int k = 0; k < K; ++k)
        out_off[(size_t)k + 1] = out_off[(size_t)k] + (unchanged[(size_t)k] ? off[k + 1] - off[k] : (int64_t)grids[(size_t)k].size());
}

void VoxelGridBatch::emit(const float* xyzi, const int64_t* off, float* out) const {
    const int K = (int)grids.size();
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        float* o = out + (size_t)out_off[(size_t)k] * 4;
        const size_t n = (size_t)(out_off[(size_t)k + 1] - out_off[(size_t)k]);
        if (n == 0) continue;
        if (unchanged[(size_t)k]) std::memcpy(o, xyzi + (size_t)off[k] * 4, n * sizeof(PointXYZI));
        else std::memcpy(o, grids[(size_t)k].data(), n * sizeof(PointXYZI));
    }
}

// ------------------------------------------------------------------------------------------------ sessions
bool list_session(const std::string& scan_dir, const std::string& pose_path, SessionFiles* s, std::string* err) {
    std::error_code ec;
    if (!fs::is_directory(scan_dir, ec)) { if (err) *err = "not a directory: " + scan_dir; return false; }
    s->scan_names.clear(); s->scan_paths.clear();
    for (auto& e : fs::directory_iterator(scan_dir)) { s->scan_names.push_back(e.path().filename().string()); s->scan_paths.push_back(e.path().string()); }
    std::sort(s->scan_names.begin(), s->scan_names.end());   // Session.cpp:91
    std::sort(s->scan_paths.begin(), s->scan_paths.end());   // Session.cpp:92
    if (!read_pose_file(pose_path, &s->scan_poses, err)) return false;
    s->scan_inverse_poses.clear();
    for (const auto& p : s->scan_poses) s->scan_inverse_poses.push_back(inverse(p));
    if (s->scan_paths.size() != s->scan_poses.size()) {      // the reference only asserts this (Session.cpp:117, compiled out in Release)
        if (err) *err = "scan count " + std::to_string(s->scan_paths.size()) + " != pose count " + std::to_string(s->scan_poses.size());
        return false;
    }
    return true;
}

std::vector<int> parse_keyframes(int num_scans, int start_idx, int end_idx, int gap) {
    std::vector<int> out;
    int num_valid_parsed = 0;
    for (int curr_idx = 0; curr_idx < num_scans; curr_idx++) {
        if (curr_idx > end_idx || curr_idx < start_idx) { curr_idx++; continue; }                  // Session.cpp:149-152 (skips two)
        if (std::remainder((double)num_valid_parsed, (double)gap) != 0) { num_valid_parsed++; continue; }   // :154-157
        out.push_back(curr_idx);
        num_valid_parsed++;
    }
    return out;
}

std::vector<int> parse_keyframes_in_roi(const std::vector<Mat4>& scan_poses, const std::vector<Mat4>& roi_poses, int gap) {
    std::vector<int> out;
    const double inplace_thres = 10.0;  // Session.cpp:234
    int num_valid_parsed = 0;
    for (int i = 0; i < (int)scan_poses.size(); ++i) {
        double nn = 10000000000.0;      // Session.cpp:218
        for (const auto& r : roi_poses) {
            const double dx = scan_poses[i][3] - r[3], dy = scan_poses[i][7] - r[7], dz = scan_poses[i][11] - r[11];
            const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (d < nn) nn = d;
        }
        if (nn > inplace_thres) continue;
        if (std::remainder((double)num_valid_parsed, (double)gap) != 0) { num_valid_parsed++; continue; }
        out.push_back(i);
        num_valid_parsed++;
    }
    return out;
}

}  // namespace ltremovert_b200
