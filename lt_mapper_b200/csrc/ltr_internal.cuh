// Internal context / device-container definitions shared by the translation units of libltr_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler (nsys, ncu --nvtx) is attached
#include <stdint.h>
#include <string>
#include <vector>
#include <cstdio>
#include <map>
#include <chrono>
#include "../../include/ltr_b200.h"

namespace ltr {

struct DevCloud {
    float* base = nullptr;      // one allocation: x | y | z | i, each `cap` floats (cap multiple of 64 -> 256 B aligned components)
    int64_t n = 0, cap = 0;
    uint8_t* flags = nullptr;   // per-point dynamic flags of the last remove pass (lazily allocated, cap bytes)
    bool used = false;
    bool borrowed = false;      // a view into another cloud's storage (ltr_cloud_slice): releasing it frees nothing
    __host__ __device__ float* x() const { return base; }
    __host__ __device__ float* y() const { return base + cap; }
    __host__ __device__ float* z() const { return base + 2 * cap; }
    __host__ __device__ float* i() const { return base + 3 * cap; }
};

struct DevScanSet {
    DevCloud pts;                  // all keyframes back to back
    std::vector<int64_t> h_off;    // K+1 host offsets
    int64_t* d_off = nullptr;      // K+1 device offsets
    int K = 0;
    bool used = false;
};

// 24 doubles per keyframe: inverse pose rows 0..2 (12) then pose rows 0..2 (12), row-major 3x4
struct DevPoses {
    double* d = nullptr;
    std::vector<double> h;
    float* d_fast = nullptr;       // K x 16 floats (KfFast, project_fast.cuh)
    std::vector<float> h_fast;
    int K = 0;
    bool used = false;
};

struct PtrView { const float *x, *y, *z, *i; int64_t n; };

// one NCCL communicator of a context (nccl_comm.cu)
struct NcclComm {
    void* comm = nullptr;       // ncclComm_t
    int rank = 0, world = 1;
    bool used = false;
    int64_t* h_cnt = nullptr;   // pinned scratch for size exchanges
    int64_t* d_cnt = nullptr;   // device scratch for size exchanges
};

}  // namespace ltr

struct ltr_ctx {
    ltr_config cfg;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_timer0 = nullptr, ev_timer1 = nullptr;
    std::string err;
    std::vector<ltr::DevCloud> clouds;
    std::vector<ltr::DevScanSet> scansets;
    std::vector<ltr::DevPoses> poses;
    std::vector<ltr::NcclComm> nccl;           // communicators created by ltr_nccl_init / ltr_nccl_split
    int64_t launches = 0;
    bool rs_attr_set = false;    // radix-sort scatter kernel opted in to > 48 KB shared memory on this device
    int64_t vox_shortcuts = 0;   // voxelisations answered by the already-one-point-per-voxel shortcut (util.cu)
    bool ext_identity = true;     // base2lidar/lidar2base exactly identity -> second transform step is exact and skipped
    double* d_ext = nullptr;      // 24 doubles: base2lidar rows 0..2, lidar2base rows 0..2
    double stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool stats_counters_pending = false;       // d_counters of the last pass not yet folded into stats[]
    unsigned long long* d_counters = nullptr;  // 4 device counters for pass statistics
    // per-kernel profile of the dominant kernel (map projection): CUDA-event time, launches, algorithmic bytes
    std::vector<cudaEvent_t> ev_pool;          // pairs
    int ev_used = 0;
    double prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // caching device allocator (single stream => a freed block can be handed out again immediately, stream-ordered)
    std::multimap<size_t, void*> free_blocks;  // size -> block
    std::map<void*, size_t> live_blocks;       // block -> size
    size_t cached_bytes = 0, live_bytes = 0, peak_live_bytes = 0;
    size_t cache_limit_bytes = (size_t)48 << 30;
    long n_cuda_malloc = 0, n_cache_hits = 0, n_purges = 0;
    bool trace = false;                        // LTR_TRACE=1: per-entry-point synchronised host timings, dumped at ltr_destroy
    std::map<std::string, std::pair<double, long>> trace_acc; // [0] remove-pass map kernel us, [1] launches, [2] algorithmic bytes, [3] point-projections
                                               // [4] parse map kernel us, [5] launches, [6] algorithmic bytes, [7] point-projections
};

namespace ltr {

extern thread_local std::string g_create_err;

int fail(ltr_ctx* ctx, int code, const char* fmt, ...);

#define LTR_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) return ::ltr::fail(ctx, LTR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                                                   cudaGetErrorString(e__), __FILE__, __LINE__);     \
    } while (0)

#define LTR_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != LTR_OK) return rc__; \
    } while (0)

#define LTR_LAUNCH_CHECK(ctx)                                                                         \
    do {                                                                                              \
        (ctx)->launches++;                                                                            \
        cudaError_t e__ = cudaGetLastError();                                                         \
        if (e__ != cudaSuccess) return ::ltr::fail(ctx, LTR_ERR_CUDA, "kernel launch failed: %s (%s:%d)", \
                                                   cudaGetErrorString(e__), __FILE__, __LINE__);      \
    } while (0)

// Entry-point guard: opens an NVTX range, makes the context's GPU the calling thread's current device for the duration of the call (the caller may have
// switched devices -- torch.cuda.set_device, a second context on another GPU -- and cudaMalloc / kernel launches follow the CURRENT
// device) and restores the previous one afterwards; with LTR_TRACE=1 it also times the call on the host between two stream
// synchronisations.
struct ApiTrace {
    ltr_ctx* c; const char* name; double t0; int prev_dev = -1;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    ApiTrace(ltr_ctx* ctx, const char* n) : c(ctx), name(n), t0(0) {
        nvtxRangePushA(n);            // every C-ABI entry point is an NVTX range named after itself
        if (!c) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != c->device) { prev_dev = cur; cudaSetDevice(c->device); }
        if (c->trace) { cudaStreamSynchronize(c->stream); t0 = now(); }
    }
    ~ApiTrace() {
        nvtxRangePop();
        if (!c) return;
        if (c->trace) { cudaStreamSynchronize(c->stream); auto& a = c->trace_acc[name]; a.first += now() - t0; a.second += 1; }
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
    }
};

// allocation helpers (stream-ordered)
int dev_alloc(ltr_ctx* ctx, void** p, size_t bytes);
void dev_free(ltr_ctx* ctx, void* p);
// Returns a scratch block to the context's cache on every exit path (LTR_TRY / LTR_CUDA return early on errors).
// Declare it right after the pointer it watches; release() frees now (so the block can be reused by the next allocation).
struct ScratchGuard {
    ltr_ctx* ctx;
    void** pp;
    ScratchGuard(ltr_ctx* c, void** p) : ctx(c), pp(p) {}
    ~ScratchGuard() { release(); }
    void release() { if (*pp) { dev_free(ctx, *pp); *pp = nullptr; } }
    ScratchGuard(const ScratchGuard&) = delete;
    ScratchGuard& operator=(const ScratchGuard&) = delete;
};

int cloud_new(ltr_ctx* ctx, int64_t n, ltr_cloud* out);          // uninitialised cloud of n points
int cloud_get(ltr_ctx* ctx, ltr_cloud h, DevCloud** c);
int scanset_new(ltr_ctx* ctx, const std::vector<int64_t>& off, ltr_scanset* out);  // allocates points + uploads offsets
int scanset_get(ltr_ctx* ctx, ltr_scanset h, DevScanSet** s);
int poses_get(ltr_ctx* ctx, ltr_poses h, DevPoses** p);
int cloud_ensure_flags(ltr_ctx* ctx, DevCloud* c);
void cloud_release(ltr_ctx* ctx, DevCloud* c);

inline int64_t round_cap(int64_t n) { return ((n > 0 ? n : 1) + 63) / 64 * 64; }
inline PtrView view(const DevCloud& c) { return PtrView{c.x(), c.y(), c.z(), c.i(), c.n}; }

// implemented in util.cu
int stable_partition_by_flag(ltr_ctx* ctx, const DevCloud& in, const uint8_t* flags, int64_t* n_flagged,
                             DevCloud* out_unflagged, DevCloud* out_flagged);   // outputs must be pre-allocated with cap >= in.n
int count_flags(ltr_ctx* ctx, const uint8_t* flags, int64_t n, int64_t* count);
int exclusive_scan_u32(ltr_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n);  // out[n] elements; returns via out
int minmax_xyz(ltr_ctx* ctx, const DevCloud& c, float mn[3], float mx[3]);
// implemented in nccl_comm.cu (internal helpers of the distributed voxeliser in util.cu)
int nccl_comm_info(ltr_ctx* ctx, int comm, int* rank, int* world);
int nccl_allreduce_u32(ltr_ctx* ctx, int comm, uint32_t* dev, size_t count, int op /* 0 sum, 1 min, 2 max */);
// `send` holds this rank's points grouped by destination rank (segment d has scount[d] points, segments back to back); *recv_out
// receives, in source-rank order, what every rank sent to this one
int nccl_alltoallv_cloud(ltr_ctx* ctx, int comm, const DevCloud& send, const std::vector<int64_t>& scount, ltr_cloud* recv_out);

int segment_counts(ltr_ctx* ctx, const uint8_t* flags, const DevScanSet& s, std::vector<int64_t>* counts);
int split_scanset_by_flag(ltr_ctx* ctx, const DevScanSet& in, const uint8_t* flags, ltr_scanset* out_unflagged, ltr_scanset* out_flagged);

}  // namespace ltr
