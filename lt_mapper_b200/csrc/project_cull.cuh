// Tile-level occlusion / range culling for the scan-minus-map variants (HD / revert / PD) of the map projection.
//
// Map clouds come out of the voxeliser in octree (Morton) order, so 128 consecutive points form a compact patch
// (median radius 0.6 m on the BASELINE maps).  For a tile with bounding sphere (c, rho) and a keyframe with origin o:
// every point p of the tile has range >= |c - o| - rho and lies inside the cone of half-angle asin(rho / |c - o|) around
// c - o, hence inside a pixel rectangle around the pixel of c.  If
//        max over that rectangle of the scan range image  <=  (|c - o| - rho) + thres - margin
// then scan(px_p) - range_p <= thres for every p in the tile: no point of the tile is a candidate, so all
// 128 (point, keyframe) pairs are skipped.  This removes the map regions that are occluded in, or beyond the reach of, a
// keyframe's scan (~30 % of all pairs on the BASELINE scene) and never changes a result.  The rectangle maximum comes
// from a max-pyramid of the scan image (levels 1..5, at most 3 x 3 lookups).  The 32 tests of one tile against the 32
// keyframes of a launch are evaluated by the 32 lanes of the warp in parallel (one keyframe per lane) and exchanged with
// one ballot, so the test costs ~1 % of the work it saves.
//
// Not applicable to the true-min variants (ND, visible points): there a far, otherwise invisible point still wins its pixel.
#pragma once
// included from the middle of project_fast.cuh (needs FastCfg / fast_project, is needed by map_project_fast_kernel)

namespace ltr {

constexpr int kPyrLevels = 5;  // levels 1..5: blocks of 2, 4, 8, 16, 32 pixels

struct CullArgs {
    const float4* tiles;        // per 128-point tile: sphere centre xyz, radius (inflated)
    long long ntiles;
    const float* pyr;           // [keyframes in launch][pyr_stride]: levels 1..5 of max(scan range), empty pixel -> empty_value
    unsigned pyr_stride;
    unsigned lvl_off[kPyrLevels + 1];
    int lvl_cols[kPyrLevels + 1];
    int enabled;
};

constexpr int kTilePts = 128;   // points per tile == the points one warp of map_project_fast_kernel owns (32 lanes x 4)

// bounding sphere of each 128-point tile: one warp per tile, 4 points per lane
__global__ void __launch_bounds__(256) tile_sphere_kernel(PtrView map, float4* __restrict__ tiles, long long ntiles) {
    const long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    const unsigned lane = threadIdx.x & 31;
    float px[4], py[4], pz[4];
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t i = min((int64_t)tile * kTilePts + j * 32 + lane, map.n - 1);   // padding repeats the last point (inside the last tile)
        px[j] = map.x[i]; py[j] = map.y[i]; pz[j] = map.z[i];
        lo[0] = fminf(lo[0], px[j]); lo[1] = fminf(lo[1], py[j]); lo[2] = fminf(lo[2], pz[j]);
        hi[0] = fmaxf(hi[0], px[j]); hi[1] = fmaxf(hi[1], py[j]); hi[2] = fmaxf(hi[2], pz[j]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
    float r2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float dx = px[j] - cx, dy = py[j] - cy, dz = pz[j] - cz; r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz); }
    for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    if (lane == 0) tiles[tile] = make_float4(cx, cy, cz, sqrtf(r2) * 1.00001f + 1.0e-4f);   // inflated: f32 rounding of centre / distance
}

// max-pyramid of the scan range images of one launch; level 0 is the scan image itself (float bits, 10000.0f = empty).
// One CTA per keyframe builds level 1 from the image and every further level from the previous one (4 reads per output).
__global__ void __launch_bounds__(1024) scan_pyramid_kernel(const uint32_t* __restrict__ scan_rimg, int nb, int rows, int cols, float empty_value,
                                                            CullArgs ca, float* __restrict__ pyr) {
    const int k = blockIdx.x;
    if (k >= nb) return;
    const uint32_t* img = scan_rimg + (size_t)k * rows * cols;
    float* P = pyr + (size_t)k * ca.pyr_stride;
    const float ninf = -__int_as_float(0x7f800000);
    int pr = rows, pc = cols;                       // dimensions of the previous level
    for (int L = 1; L <= kPyrLevels; ++L) {
        const int RL = (pr + 1) >> 1, CL = (pc + 1) >> 1;
        float* out = P + ca.lvl_off[L];
        const float* prev = (L > 1) ? P + ca.lvl_off[L - 1] : nullptr;
        for (int e = threadIdx.x; e < RL * CL; e += blockDim.x) {
            const int br = e / CL, bc = e - br * CL;
            float m = ninf;
#pragma unroll
            for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                for (int dc = 0; dc < 2; ++dc) {
                    const int r = 2 * br + dr, c = 2 * bc + dc;
                    if (r < pr && c < pc) {
                        float v;
                        if (L == 1) { const uint32_t b = img[(size_t)r * cols + c]; v = (b == 0x461C4000u) ? empty_value : __uint_as_float(b); }
                        else v = prev[r * pc + c];
                        m = fmaxf(m, v);
                    }
                }
            out[e] = m;
        }
        __syncthreads();
        pr = RL; pc = CL;
    }
}

// true iff no point of the tile can be a candidate in keyframe k (kf = its fast constants).  Every comparison is written so
// that NaN / degenerate geometry yields "false" (keep the tile).
LTR_DEV bool tile_culled(const float* __restrict__ kf, float4 tile, const FastCfg& fc, const ImgShape& g, const CullArgs& ca, int k, float thres) {
    if (kf[15] == 0.0f) return false;
    const FastProj f = fast_project(kf, tile.x, tile.y, tile.z);   // f.r = |c - o|, f.rho_inv_r = 1 / cos(elevation of c)
    const float d = f.r;
    const float x = __fmul_rn(__fmul_rn(tile.w, mufu_rcp(d)), 1.00001f);     // sin(theta) = rho / d
    const float y = __fmul_rn(x, f.rho_inv_r);                                 // sin(theta) / cos(el_c) bounds sin(azimuth half-width)
    if (!(x <= 0.5f) || !(y <= 0.5f)) return false;                           // asin(t) <= 1.05 t on [0, 0.5]; the cone must not contain the pole
    const float vcol = __fmaf_rn(f.az, fc.col_scale, fc.col_off);
    const float vrow = __fmaf_rn(-f.el, fc.row_scale, fc.row_off);
    const float dc = __fmaf_rn(__fmul_rn(1.05f, y), fc.col_scale, 1.5f);       // half-widths in pixels (+1.5 px: rounding, approximation)
    const float dr = __fmaf_rn(__fmul_rn(1.05f, x), fc.row_scale, 1.5f);
    const float c0f = floorf(__fsub_rn(vcol, dc)), c1f = ceilf(__fadd_rn(vcol, dc));
    if (!(c0f >= 0.0f) || !(c1f <= (float)(g.cols - 1))) return false;         // azimuth seam (or NaN): keep
    if (!(vrow == vrow) || !(dr == dr)) return false;
    const int c0 = (int)c0f, c1 = (int)c1f;
    const int r0 = min(max((int)fminf(fmaxf(floorf(__fsub_rn(vrow, dr)), -1.0f), 1.0e6f), 0), g.rows - 1);  // rows clamp exactly like the pixel index does
    const int r1 = min(max((int)fminf(fmaxf(ceilf(__fadd_rn(vrow, dr)), -1.0f), 1.0e6f), 0), g.rows - 1);
    const int e = max(c1 - c0, r1 - r0) + 1;
    int L = 32 - __clz(e - 1) - 1;                                             // blocks of 2^L >= e / 2: at most 3 x 3 blocks
    L = max(L, 1);
    if (L > kPyrLevels) return false;
    const float* P = ca.pyr + (size_t)k * ca.pyr_stride + ca.lvl_off[L];
    const int CL = ca.lvl_cols[L];
    float smax = -__int_as_float(0x7f800000);
    for (int br = r0 >> L; br <= (r1 >> L); ++br)
        for (int bc = c0 >> L; bc <= (c1 >> L); ++bc) smax = fmaxf(smax, P[br * CL + bc]);
    const float rmin = __fsub_rn(__fsub_rn(__fmul_rn(d, 0.999996f), tile.w), 1.0e-4f);
    return smax <= __fsub_rn(__fadd_rn(rmin, thres), 1.0e-3f);                 // scan - range <= thres - 1 mm for every point of the tile
}

}  // namespace ltr
