// ltremovert_b200 -- standalone file-in/file-out driver with the surface of the `removert_removert` ROS node
// (ltremovert/launch/run_ltmapper.launch:6-13, ltremovert/src/removert_main.cpp:3-13): reads the same yaml keys
// (ltremovert/config/params_ltmapper.yaml, namespace "removert"), the same scan directories (binary PCD, fields
// x y z intensity) and pose files (12 numbers per line), runs Removerter::run() (ltremovert/src/Removerter.cpp:1653-1678)
// on the GPU through libltr_removert/libltr_b200, and writes the same output tree (Removerter.cpp:26-50, 231, 1446-1477,
// 1517-1520, 1600-1601, 1607-1650).  No ROS, no PCL; rviz publishing is dropped (side effect only).
//
//   ltremovert_b200 --config params.yaml [--selfremovert] [--pcl110] [--no-split]
//     --selfremovert   run selfRemovert over remove_resolution_list (Removerter.cpp:1378-1393) instead of the shipped
//                      single removeOnce(2.5) (Removerter.cpp:1584, 1587)
//     --pcl110         PCL >= 1.10 transformPointCloud summation order (ltr_config.transform_order = 1)
//
// Multi-GPU: start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK in the environment (torchrun, mpirun -x, a shell loop);
// every rank reads the yaml and the pose files, but loads only the scans of ITS keyframe block (even world: ranks < world/2 take the
// central session, the others the query session; --no-split: a block of both).  The NCCL id travels through a file
// ($LTR_NCCL_ID_FILE, default /tmp/ltr_nccl_id_<parent pid>_<MASTER_PORT>).  Maps are written by the first rank of the group that holds
// them, per-keyframe scans by the rank that owns the keyframe; the output tree is the single-process one.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>
#include "../lt_mapper_b200/csrc/host/io.h"
#include "../lt_mapper_b200/csrc/host/removerter.h"

namespace fs = std::filesystem;
using namespace ltremovert_b200;

static void fsmkdir(const std::string& p) { if (!fs::is_directory(p) || !fs::exists(p)) fs::create_directories(p); }  // Removerter.cpp:6-10

struct LoadedSession {
    SessionFiles files;
    std::vector<std::string> keyframe_names;
    std::vector<float> xyzi;
    std::vector<int64_t> offsets;
    std::vector<double> poses, inv_poses;
};

// the keyframes [k0, k1) of the session's keyframe list (a rank's block; everything in a single-process run)
static bool load_keyframes(LoadedSession& s, float voxel, size_t k0, size_t k1, std::string* err) {  // Session::loadKeyframes (Session.cpp:266-302)
    s.offsets.assign(1, 0);
    int overflow_warnings = 0;
    for (size_t kk = k0; kk < k1; ++kk) {
        const int idx = s.files.keyframe_idx[kk];
        HostCloud pts;
        if (!read_pcd(s.files.scan_paths[(size_t)idx], &pts, err)) return false;
        bool ov = false;
        const HostCloud down = voxel_grid(pts, voxel, &ov);
        overflow_warnings += ov ? 1 : 0;
        for (const auto& p : down) { s.xyzi.push_back(p.x); s.xyzi.push_back(p.y); s.xyzi.push_back(p.z); s.xyzi.push_back(p.intensity); }
        s.offsets.push_back(s.offsets.back() + (int64_t)down.size());
        s.keyframe_names.push_back(s.files.scan_names[(size_t)idx]);
        const Mat4& P = s.files.scan_poses[(size_t)idx];
        const Mat4& IP = s.files.scan_inverse_poses[(size_t)idx];
        s.poses.insert(s.poses.end(), P.begin(), P.end());
        s.inv_poses.insert(s.inv_poses.end(), IP.begin(), IP.end());
    }
    if (overflow_warnings)
        std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] Leaf size is too small for the input dataset (%d scans returned unchanged)\n", overflow_warnings);
    return true;
}

static bool save_cloud(Removerter& R, ltr_cloud h, const std::string& path, bool octree_layout) {
    int64_t n = 0;
    if (ltr_cloud_size(R.ctx, h, &n) != LTR_OK) return false;
    std::vector<float> buf((size_t)n * 4 + 4);
    if (ltr_cloud_download(R.ctx, h, buf.data(), n, &n) != LTR_OK) return false;
    HostCloud c((size_t)n);
    if (n) std::memcpy(c.data(), buf.data(), (size_t)n * 16);
    std::string err;
    if (!write_pcd_binary(path, c, octree_layout, &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return false; }
    return true;
}

static bool save_scans(Removerter& R, ltr_scanset s, const std::vector<std::string>& names, const std::string& dir, bool octree_layout) {  // saveScans (:1632-1650)
    int32_t K = 0; int64_t total = 0;
    if (s < 0 || ltr_scanset_info(R.ctx, s, &K, &total) != LTR_OK) return false;
    std::vector<float> buf((size_t)total * 4 + 4);
    std::vector<int64_t> off((size_t)K + 1);
    if (ltr_scanset_download(R.ctx, s, buf.data(), total, off.data()) != LTR_OK) return false;
    for (int k = 0; k < K; ++k) {
        HostCloud c((size_t)(off[k + 1] - off[k]));
        if (!c.empty()) std::memcpy(c.data(), buf.data() + 4 * off[k], c.size() * 16);
        std::string err;
        if (!write_pcd_binary(dir + "/" + names[(size_t)k], c, octree_layout, &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return false; }
    }
    return true;
}

static int env_int(const char* name, int dflt) { const char* v = std::getenv(name); return v && *v ? std::atoi(v) : dflt; }

// the 128 bytes of ncclGetUniqueId from rank 0 to everybody else, through a file that appears atomically (write + rename)
static bool share_nccl_id(int rank, unsigned char id[128]) {
    const char* e = std::getenv("LTR_NCCL_ID_FILE");
    const char* port = std::getenv("MASTER_PORT");
    const std::string path = e && *e ? e : "/tmp/ltr_nccl_id_" + std::to_string((long)getppid()) + "_" + (port ? port : "0");
    if (rank == 0) {
        if (ltr_nccl_unique_id(id) != LTR_OK) return false;
        { std::ofstream f(path + ".tmp", std::ios::binary); f.write((const char*)id, 128); if (!f) return false; }
        return std::rename((path + ".tmp").c_str(), path.c_str()) == 0;
    }
    for (int tries = 0; tries < 600; ++tries) {   // up to 60 s
        std::ifstream f(path, std::ios::binary);
        if (f && f.read((char*)id, 128) && f.gcount() == 128) return true;
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    return false;
}

int main(int argc, char** argv) {
    std::string cfg_path;
    bool selfremovert = false, pcl110 = false, no_split = false;
    const char* usage = "usage: %s --config params.yaml [--selfremovert] [--pcl110] [--no-split]\n";
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--config") && i + 1 < argc) cfg_path = argv[++i];
        else if (!std::strcmp(argv[i], "--selfremovert")) selfremovert = true;
        else if (!std::strcmp(argv[i], "--pcl110")) pcl110 = true;
        else if (!std::strcmp(argv[i], "--no-split")) no_split = true;
        else { std::fprintf(stderr, usage, argv[0]); return 2; }
    }
    if (cfg_path.empty()) { std::fprintf(stderr, usage, argv[0]); return 2; }
    const int rank = env_int("RANK", 0), world = std::max(1, env_int("WORLD_SIZE", 1));
    const bool split = world >= 2 && world % 2 == 0 && !no_split;
    YamlParams y;
    std::string err;
    if (!y.load(cfg_path, &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }

    // RosParamServer.cpp:4-63 (same keys, same C++ defaults)
    ltrh_params P;
    ltrh_params_default(&P);
    P.device = env_int("LOCAL_RANK", 0);
    P.sequence_vfov = (float)y.num("removert/sequence_vfov", 50.0);
    P.sequence_hfov = (float)y.num("removert/sequence_hfov", 360.0);
    P.num_nn_points_within = (int)y.num("removert/num_nn_points_within", 3);
    P.dist_nn_points_within = (float)y.num("removert/dist_nn_points_within", 0.1);
    P.downsample_voxel_size = (float)y.num("removert/downsample_voxel_size", 0.05);
    const std::vector<double> ext = y.list("removert/ExtrinsicLiDARtoPoseBase");
    if (ext.size() == 16) for (int i = 0; i < 16; ++i) P.ExtrinsicLiDARtoPoseBase[i] = ext[(size_t)i];
    P.transform_order = pcl110 ? 1 : 0;
    const std::vector<double> remove_res = y.list("removert/remove_resolution_list");
    if (selfremovert) {
        P.n_schedule = 0;
        for (double r : remove_res) {   // selfRemovert (Removerter.cpp:1380-1389)
            if (P.n_schedule + 3 > LTRH_MAX_SCHEDULE) break;
            P.schedule_op[P.n_schedule] = LTRH_OP_REMOVE; P.schedule_res[P.n_schedule++] = (float)r;
            P.schedule_op[P.n_schedule] = LTRH_OP_REVERT; P.schedule_res[P.n_schedule++] = (float)(0.95 * (float)r);
            P.schedule_op[P.n_schedule] = LTRH_OP_REMOVE; P.schedule_res[P.n_schedule++] = (float)r;
        }
    }
    std::string save_dir = y.str("removert/save_pcd_directory", "/");
    if (save_dir.empty() || save_dir.back() != '/') save_dir += "/";      // Removerter.cpp:26-27
    const bool save_map_pcd = y.boolean("removert/saveMapPCD", false);
    const int start_idx = (int)y.num("removert/start_idx", 1), end_idx = (int)y.num("removert/end_idx", 100);
    const int keyframe_gap = (int)y.num("removert/keyframe_gap", 10);

    // output tree (Removerter.cpp:28-50)
    fsmkdir(save_dir);
    const std::string d_updated = save_dir + "scans_updated", d_updated_strong = save_dir + "scans_updated_strong", d_pd = save_dir + "scans_pd",
                      d_pd_strong = save_dir + "scans_pd_strong", d_nd_strong = save_dir + "scans_nd_strong";
    for (const std::string& d : {d_updated, d_updated_strong, d_pd, d_pd_strong, d_nd_strong, save_dir + "map_static", save_dir + "map_dynamic"}) fsmkdir(d);

    // Step 0 host side: loadSessionInfo, parseKeyframes, loadKeyframes (run() :1656-1659)
    LoadedSession C, Q;
    if (!list_session(y.str("removert/central_sess_scan_dir", ""), y.str("removert/central_sess_pose_path", ""), &C.files, &err) ||
        !list_session(y.str("removert/query_sess_scan_dir", ""), y.str("removert/query_sess_pose_path", ""), &Q.files, &err)) {
        std::fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    if (rank == 0) std::printf(" Total : %zu / %zu scans in the directories.\n", C.files.scan_paths.size(), Q.files.scan_paths.size());
    C.files.keyframe_idx = parse_keyframes((int)C.files.scan_paths.size(), start_idx, end_idx, keyframe_gap);            // Removerter.cpp:92
    std::vector<Mat4> roi;
    for (int i : C.files.keyframe_idx) roi.push_back(C.files.scan_poses[(size_t)i]);
    Q.files.keyframe_idx = parse_keyframes_in_roi(Q.files.scan_poses, roi, keyframe_gap);                               // Removerter.cpp:93
    if (rank == 0)
        std::printf(" Total %zu central keyframes from the index range [%d, %d], %zu query keyframes in the map's ROI\n",
                    C.files.keyframe_idx.size(), start_idx, end_idx, Q.files.keyframe_idx.size());
    // this rank's contiguous keyframe block of each session (keyframe order == rank order); everything when world == 1
    auto block = [&](size_t K, bool owns, size_t* k0, size_t* k1) {
        const int g = split ? world / 2 : world, r = split ? rank % (world / 2) : rank;
        *k0 = owns ? K * (size_t)r / (size_t)g : 0;
        *k1 = owns ? K * (size_t)(r + 1) / (size_t)g : 0;
    };
    const bool owns_c = !split || rank < world / 2, owns_q = !split || rank >= world / 2;
    size_t c0, c1, q0, q1;
    block(C.files.keyframe_idx.size(), owns_c, &c0, &c1);
    block(Q.files.keyframe_idx.size(), owns_q, &q0, &q1);
    if (!load_keyframes(C, P.downsample_voxel_size, c0, c1, &err) || !load_keyframes(Q, P.downsample_voxel_size, q0, q1, &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }

    Removerter R(P);
    if (R.init() != LTR_OK) { std::fprintf(stderr, "GPU context: %s\n", R.err.c_str()); return 1; }
    auto ck = [&](int rc, const char* what) { if (rc != LTR_OK) { std::fprintf(stderr, "[rank %d] %s failed (%d): %s\n", rank, what, rc, R.err.c_str()); std::exit(1); } };
    if (world > 1) {
        unsigned char id[128];
        if (!share_nccl_id(rank, id)) { std::fprintf(stderr, "[rank %d] could not exchange the NCCL id\n", rank); return 1; }
        ck(R.comm_init_nccl(id, rank, world, split ? 1 : 0), "NCCL init");
    }
    ck(R.load_session(0, C.xyzi.data(), C.offsets.data(), C.poses.data(), C.inv_poses.data(), (int)C.keyframe_names.size()), "load central");
    ck(R.load_session(1, Q.xyzi.data(), Q.offsets.data(), Q.poses.data(), Q.inv_poses.data(), (int)Q.keyframe_names.size()), "load query");
    ck(R.run_step0(), "Step 0");
    ck(R.run_step12(), "Step 1-2");
    ck(R.run_step3(), "Step 3");
    if (rank == 0) for (const auto& l : R.log)
        std::printf(" %-18s map %lld  dynamic %lld  -> static %lld  dynamic %lld\n", l.what.c_str(), (long long)l.n_map, (long long)l.n_dynamic,
                    (long long)l.n_static_after, (long long)l.n_dynamic_after);

    // saved maps (all come out of octreeDownsampling: WIDTH 1 / HEIGHT n)
    bool ok = true;
    // a saved map is replicated on the ranks of the group that produced it: its first rank writes it
    const bool group_leader = split ? (rank == 0 || rank == world / 2) : rank == 0;
    for (const auto& kv : R.saved) {
        if (!group_leader) break;
        if (kv.first.rfind("OriginalNoisy", 0) == 0 && !save_map_pcd) continue;   // kFlagSaveMapPointcloud (Removerter.cpp:228)
        ok &= save_cloud(R, kv.second, save_dir + kv.first + ".pcd", true);
    }
    // saveAllTypeOfScans (Removerter.cpp:1607-1630); only keyframe_scans_updated_ went through octreeDownsampling (Session.cpp:374)
    Session& S = R.central_sess_;
    if (owns_c) {   // per-keyframe scans: the owner of the keyframe writes them, under the input scan's name
        ok &= save_scans(R, S.keyframe_scans_updated_, C.keyframe_names, d_updated, true);
        ok &= save_scans(R, S.keyframe_scans_updated_strong_, C.keyframe_names, d_updated_strong, false);
        ok &= save_scans(R, S.keyframe_scans_pd_, C.keyframe_names, d_pd, false);
        ok &= save_scans(R, S.keyframe_scans_strong_pd_, C.keyframe_names, d_pd_strong, false);
        ok &= save_scans(R, S.keyframe_scans_strong_nd_, C.keyframe_names, d_nd_strong, false);
    }
    if (world > 1 && R.nccl_world >= 0) ltr_nccl_barrier(R.ctx, R.nccl_world);   // nobody leaves before everybody's files are written
    if (rank == 0) std::printf(" outputs written to %s (%s)\n", save_dir.c_str(), ok ? "ok" : "WITH ERRORS");
    return ok ? 0 : 1;
}
