// Deterministic synthetic two-session LiDAR generator: "ParkingLot-shaped" scene,
// 64-beam x 1800-step spinning sensor, session 2 = shifted copy of session 1 with removed cars (ND),
// new cars (PD) and moving boxes (HD).  Data generator only: neither product path nor oracle.
// Output formats mirror what the reference consumes after loading (Session.cpp:102-114, 266-302):
// per keyframe a cloud of (x, y, z, intensity) float points in the LiDAR frame and a 4x4 row-major
// double pose (LiDAR -> world; extrinsic identity as in params_ltmapper.yaml:28-31).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Box { double cx, cy, cz, hx, hy, hz, cyaw, syaw; };

inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
inline uint64_t key4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    return mix64(mix64(mix64(mix64(mix64(seed) ^ a) ^ b) ^ c) ^ d);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct SynthCfg {
    uint64_t seed;
    int session;       // 0 = central, 1 = query
    int K;
    int beams, az_steps;
    double max_range, noise_sigma;
    double spacing;    // metres between keyframes
    int n_cars, n_poles, n_movers;
};

const double kPi = 3.14159265358979323846;

Box makeBox(double cx, double cy, double cz, double lx, double ly, double lz, double yaw) {
    Box b; b.cx = cx; b.cy = cy; b.cz = cz; b.hx = lx / 2; b.hy = ly / 2; b.hz = lz / 2; b.cyaw = std::cos(yaw); b.syaw = std::sin(yaw);
    return b;
}

// static scene of one session
void buildStatic(const SynthCfg& c, std::vector<Box>& boxes) {
    // facades around a 150 x 100 m lot, 30 m tall so that almost every beam returns
    boxes.push_back(makeBox(0, 54, 15, 170, 8, 30, 0));
    boxes.push_back(makeBox(0, -54, 15, 170, 8, 30, 0));
    boxes.push_back(makeBox(79, 0, 15, 8, 116, 30, 0));
    boxes.push_back(makeBox(-79, 0, 15, 8, 116, 30, 0));
    // poles / trees on a jittered grid
    for (int i = 0; i < c.n_poles; ++i) {
        const uint64_t h = key4(c.seed, 1001, (uint64_t)i, 0, 0);
        const double x = -70 + 140 * u01(h), y = -47 + 94 * u01(mix64(h));
        const double ht = 5 + 4 * u01(mix64(h + 7));
        boxes.push_back(makeBox(x, y, ht / 2, 0.3, 0.3, ht, 0));
        if (i % 3 == 0) boxes.push_back(makeBox(x, y, ht + 1.0, 2.5, 2.5, 2.0, 0.3 * i));  // canopy
    }
    // parked cars: 8 rows between the lanes, slots every 3 m; which slots are occupied differs by session
    int made = 0;
    for (int row = 0; row < 8; ++row)
        for (int slot = 0; slot < 39; ++slot) {
            const uint64_t h = key4(c.seed, 2002, (uint64_t)row, (uint64_t)slot, 0);
            const double occ = u01(h);
            if (occ > (double)c.n_cars / 312.0) continue;
            const double life = u01(mix64(h + 1));
            // 15 % only in session 0 (-> ND), 15 % only in session 1 (-> PD), 70 % in both
            const bool only0 = life < 0.15, only1 = life >= 0.85;
            if (c.session == 0 && only1) continue;
            if (c.session == 1 && only0) continue;
            const double y = -35 + 10 * row, x = -57 + 3 * slot;
            const double yaw = kPi / 2 + (u01(mix64(h + 2)) - 0.5) * 0.1;
            boxes.push_back(makeBox(x, y, 0.75, 4.5, 1.8, 1.5, yaw));
            ++made;
        }
    (void)made;
}

// serpentine path: 9 lanes along x at y = -40 + 10 j, joined by 10 m connectors; arc-length parameterised
void pathAt(double s, double* x, double* y, double* yaw) {
    const double lane = 120.0, conn = 10.0, period = 9 * lane + 8 * conn;
    double t = std::fmod(s, 2 * period);
    bool back = false;
    if (t > period) { t = 2 * period - t; back = true; }
    int j = 0;
    double rem = t;
    for (; j < 9; ++j) {
        if (rem <= lane || j == 8) break;
        rem -= lane;
        if (rem <= conn) {  // on connector j -> j+1
            const double xe = (j % 2 == 0) ? 60.0 : -60.0;
            *x = xe; *y = -40 + 10 * j + rem; *yaw = back ? -kPi / 2 : kPi / 2;
            return;
        }
        rem -= conn;
    }
    rem = std::min(rem, lane);
    const bool fwd = (j % 2 == 0);
    *x = fwd ? (-60.0 + rem) : (60.0 - rem);
    *y = -40 + 10 * j;
    *yaw = (fwd != back) ? 0.0 : kPi;
}

void poseOf(const SynthCfg& c, int k, double* T /*16 row-major*/) {
    double x, y, yaw;
    const double s0 = (c.session == 0) ? 0.0 : 0.3;
    pathAt(s0 + c.spacing * k, &x, &y, &yaw);
    const uint64_t h = key4(c.seed, 3003, (uint64_t)c.session, (uint64_t)k, 0);
    double roll = (u01(h) - 0.5) * 0.017, pitch = (u01(mix64(h)) - 0.5) * 0.017;
    double z = 1.9 + (u01(mix64(h + 3)) - 0.5) * 0.04;
    if (c.session == 1) { y += 0.5; yaw += 2.0 * kPi / 180.0; }
    yaw += (u01(mix64(h + 5)) - 0.5) * 0.02;
    const double cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch), cy = std::cos(yaw), sy = std::sin(yaw);
    // R = Rz(yaw) Ry(pitch) Rx(roll)
    T[0] = cy * cp; T[1] = cy * sp * sr - sy * cr; T[2] = cy * sp * cr + sy * sr; T[3] = x;
    T[4] = sy * cp; T[5] = sy * sp * sr + cy * cr; T[6] = sy * sp * cr - cy * sr; T[7] = y;
    T[8] = -sp;     T[9] = cp * sr;                T[10] = cp * cr;               T[11] = z;
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

void moversAt(const SynthCfg& c, int k, std::vector<Box>& boxes) {
    for (int i = 0; i < c.n_movers; ++i) {
        const uint64_t h = key4(c.seed, 4004, (uint64_t)c.session, (uint64_t)i, 0);
        const int lane = (int)(u01(h) * 9) % 9;
        const double side = (u01(mix64(h)) < 0.5) ? -1.6 : 1.6;
        const double x0 = -60 + 120 * u01(mix64(h + 1));
        const double v = 0.4 + 0.8 * u01(mix64(h + 2));
        double x = x0 + v * k;
        x = -60 + std::fmod(x + 60, 120.0);
        boxes.push_back(makeBox(x, -40 + 10 * lane + side, 0.85, 1.2, 0.7, 1.7, 0));
    }
}

inline bool rayBox(const Box& b, const double* o, const double* d, double tmax, double* thit) {
    const double ox = o[0] - b.cx, oy = o[1] - b.cy, oz = o[2] - b.cz;
    const double lx = b.cyaw * ox + b.syaw * oy, ly = -b.syaw * ox + b.cyaw * oy;
    const double dx = b.cyaw * d[0] + b.syaw * d[1], dy = -b.syaw * d[0] + b.cyaw * d[1], dz = d[2];
    double t0 = 0.0, t1 = tmax;
    const double lo[3] = {lx, ly, oz}, dd[3] = {dx, dy, dz}, hh[3] = {b.hx, b.hy, b.hz};
    for (int a = 0; a < 3; ++a) {
        if (std::fabs(dd[a]) < 1e-12) { if (std::fabs(lo[a]) > hh[a]) return false; continue; }
        const double inv = 1.0 / dd[a];
        double ta = (-hh[a] - lo[a]) * inv, tb = (hh[a] - lo[a]) * inv;
        if (ta > tb) std::swap(ta, tb);
        t0 = std::max(t0, ta); t1 = std::min(t1, tb);
        if (t0 > t1) return false;
    }
    if (t0 <= 1e-6) return false;  // origin inside the box
    *thit = t0;
    return true;
}

}  // namespace

extern "C" {

// poses_out: K*16 doubles for keyframes k0 .. k0+K-1.
void ltr_synth_poses(uint64_t seed, int session, int k0, int K, double spacing, double* poses_out) {
    SynthCfg c{}; c.seed = seed; c.session = session; c.K = K; c.spacing = spacing;
    for (int k = 0; k < K; ++k) poseOf(c, k0 + k, poses_out + 16 * (size_t)k);
}

// Generates K scans. xyzi_out must hold K*beams*az_steps*4 floats (upper bound); offsets_out K+1 int64.
// Returns the total number of points written.
int64_t ltr_synth_session(uint64_t seed, int session, int k0, int K, int beams, int az_steps, double max_range, double noise_sigma,
                          double spacing, int n_cars, int n_poles, int n_movers, int threads,
                          float* xyzi_out, int64_t* offsets_out, double* poses_out) {
    SynthCfg c{}; c.seed = seed; c.session = session; c.K = K; c.beams = beams; c.az_steps = az_steps; c.max_range = max_range;
    c.noise_sigma = noise_sigma; c.spacing = spacing; c.n_cars = n_cars; c.n_poles = n_poles; c.n_movers = n_movers;
    std::vector<Box> stat;
    buildStatic(c, stat);
    const size_t per = (size_t)beams * az_steps;
    std::vector<int64_t> counts(K, 0);
    std::vector<float> tmp((size_t)K * per * 4);  // scratch, compacted afterwards
    const int nb = 720;                            // world-azimuth bins for box culling
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
    for (int kk = 0; kk < K; ++kk) {
        const int k = k0 + kk;  // absolute keyframe index along the trajectory
        double T[16];
        poseOf(c, k, T);
        std::memcpy(poses_out + 16 * (size_t)kk, T, sizeof(T));
        std::vector<Box> boxes = stat;
        moversAt(c, k, boxes);
        const double o[3] = {T[3], T[7], T[11]};
        // per-bin candidate lists from footprint azimuth extents
        std::vector<std::vector<int>> bins(nb);
        for (int bi = 0; bi < (int)boxes.size(); ++bi) {
            const Box& b = boxes[bi];
            const double ox = o[0] - b.cx, oy = o[1] - b.cy;
            const double lx = b.cyaw * ox + b.syaw * oy, ly = -b.syaw * ox + b.cyaw * oy;
            if (std::fabs(lx) <= b.hx + 0.05 && std::fabs(ly) <= b.hy + 0.05) { for (int q = 0; q < nb; ++q) bins[q].push_back(bi); continue; }
            double amin = 1e9, amax = -1e9, aref = std::atan2(b.cy - o[1], b.cx - o[0]);
            for (int sx = -1; sx <= 1; sx += 2) for (int sy = -1; sy <= 1; sy += 2) {
                const double wx = b.cx + b.cyaw * sx * b.hx - b.syaw * sy * b.hy, wy = b.cy + b.syaw * sx * b.hx + b.cyaw * sy * b.hy;
                double a = std::atan2(wy - o[1], wx - o[0]) - aref;
                while (a > kPi) a -= 2 * kPi;
                while (a < -kPi) a += 2 * kPi;
                amin = std::min(amin, a); amax = std::max(amax, a);
            }
            const int q0 = (int)std::floor((aref + amin + kPi) / (2 * kPi) * nb) - 1, q1 = (int)std::floor((aref + amax + kPi) / (2 * kPi) * nb) + 1;
            for (int q = q0; q <= q1; ++q) bins[((q % nb) + nb) % nb].push_back(bi);
        }
        float* out = tmp.data() + (size_t)kk * per * 4;
        int64_t cnt = 0;
        for (int a = 0; a < az_steps; ++a) {
            const double az = -kPi + 2 * kPi * (a + 0.5) / az_steps;
            for (int bm = 0; bm < beams; ++bm) {
                const double el = (-22.5 + 45.0 * bm / (beams - 1)) * kPi / 180.0;
                const double dl[3] = {std::cos(el) * std::cos(az), std::cos(el) * std::sin(az), std::sin(el)};
                const double d[3] = {T[0] * dl[0] + T[1] * dl[1] + T[2] * dl[2], T[4] * dl[0] + T[5] * dl[1] + T[6] * dl[2],
                                     T[8] * dl[0] + T[9] * dl[1] + T[10] * dl[2]};
                double t = max_range;
                bool hit = false;
                if (d[2] < -1e-9) { const double tg = -o[2] / d[2]; if (tg < t) { t = tg; hit = true; } }
                double wa = std::atan2(d[1], d[0]);
                int q = (int)std::floor((wa + kPi) / (2 * kPi) * nb);
                q = std::min(std::max(q, 0), nb - 1);
                for (int bi : bins[q]) { double th; if (rayBox(boxes[bi], o, d, t, &th) && th < t) { t = th; hit = true; } }
                if (!hit) continue;
                const uint64_t h = key4(seed, (uint64_t)session, (uint64_t)k, (uint64_t)bm, (uint64_t)a);
                const double u1 = std::max(u01(h), 1e-12), u2 = u01(mix64(h));
                const double g = std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * kPi * u2);
                const double r = t + noise_sigma * g;
                if (r <= 0.3 || r >= max_range) continue;
                out[4 * cnt + 0] = (float)(r * dl[0]);
                out[4 * cnt + 1] = (float)(r * dl[1]);
                out[4 * cnt + 2] = (float)(r * dl[2]);
                out[4 * cnt + 3] = (float)(mix64(h + 9) & 0xff);
                ++cnt;
            }
        }
        counts[kk] = cnt;
    }
    offsets_out[0] = 0;
    for (int k = 0; k < K; ++k) offsets_out[k + 1] = offsets_out[k] + counts[k];
    for (int k = 0; k < K; ++k)
        std::memmove(xyzi_out + 4 * (size_t)offsets_out[k], tmp.data() + (size_t)k * per * 4, (size_t)counts[k] * 16);
    return offsets_out[K];
}

}  // extern "C"
