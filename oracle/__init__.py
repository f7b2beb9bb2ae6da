"""ORACLE (test infrastructure, not product code).

ctypes front-end of the dependency-free CPU restatement of the LT-removert hot path (oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; lt_mapper_b200 never does.  First-party logic is pinned against the reference's own sources compiled
behind stand-in third-party headers (oracle/ref.py, tests/test_ref_pin.py); PARITY UNPINNED for the restated PCL / FLANN /
Eigen semantics (the reference ships no tests).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_HD, MODE_ND, MODE_PD = 0, 1, 2
OP_REMOVE, OP_REVERT = 0, 1
IDENTITY = np.eye(4, dtype=np.float64)


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        i64, f32, f64, vp, ci, cp = ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
        L.ltro_atan2f_selfcheck.restype = i64
        L.ltro_atan2f_selfcheck.argtypes = [ctypes.c_uint64, i64]
        L.ltro_atan2f_selfcheck_wide.restype = i64
        L.ltro_atan2f_selfcheck_wide.argtypes = [ctypes.c_uint64, i64]
        L.ltro_atan2f.argtypes = [vp, vp, vp, i64]
        L.ltro_libm_atan2f.argtypes = [vp, vp, vp, i64]
        L.ltro_reset_rimg_size.argtypes = [f32, f32, f32, vp, vp]
        L.ltro_pixel_index.argtypes = [vp, i64, ci, f32, f32, ci, ci, vp, vp, vp]
        L.ltro_transform.argtypes = [vp, i64, vp, ci, vp]
        L.ltro_inverse4x4.argtypes = [vp, vp]
        L.ltro_scan2rimg.argtypes = [vp, i64, f32, f32, ci, ci, vp]
        L.ltro_map2rimg.argtypes = [vp, i64, f32, f32, ci, ci, vp, vp]
        L.ltro_remove_pass.restype = i64
        L.ltro_remove_pass.argtypes = [vp, i64, vp, vp, vp, ci, f32, f32, vp, ci, ci, f32, f32, ci, vp]
        L.ltro_parse_projected.restype = i64
        L.ltro_parse_projected.argtypes = [vp, i64, vp, f32, f32, vp, ci, f32, vp, vp, i64]
        L.ltro_voxel.restype = i64
        L.ltro_voxel.argtypes = [vp, i64, f32, vp, i64]
        L.ltro_knn_dists.argtypes = [vp, i64, vp, i64, ci, vp, ci]
        L.ltro_knn_partition.restype = i64
        L.ltro_knn_partition.argtypes = [vp, i64, vp, vp, vp, i64, vp, ci, ci, f32, vp, vp, vp]
        L.ltro_create.restype = vp
        L.ltro_destroy.argtypes = [vp]
        L.ltro_set_params.argtypes = [vp, f32, f32, vp, ci, ci, f32, f32, ci, ci, ci, ci]
        L.ltro_set_schedule.argtypes = [vp, vp, vp, ci]
        L.ltro_load_session.argtypes = [vp, ci, vp, vp, vp, vp, ci]
        L.ltro_set_map.argtypes = [vp, ci, cp, vp, i64]
        L.ltro_run.argtypes = [vp, ci]
        L.ltro_stage.argtypes = [vp, cp]
        L.ltro_cloud_size.restype = i64
        L.ltro_cloud_size.argtypes = [vp, cp, ci, ci]
        L.ltro_cloud_copy.restype = i64
        L.ltro_cloud_copy.argtypes = [vp, cp, ci, ci, vp, i64]
        L.ltro_num_keyframes.argtypes = [vp, ci]
        L.ltro_log_get.argtypes = [vp, ci, cp, ci, vp]
        L.ltro_log_count.argtypes = [vp]
        L.ltro_timing.restype = f64
        L.ltro_timing.argtypes = [vp, cp]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return _lib().ltro_max_threads()


def atan2f_selfcheck_wide(seed, n):
    """bit mismatches between ref_atan2f and the container's libm over log-uniform magnitudes 2^-40 .. 2^40 (branch thresholds)."""
    return _lib().ltro_atan2f_selfcheck_wide(seed, n)


def atan2f_selfcheck(seed, n):
    """Number of bit mismatches between ref_atan2f and this machine's libm atan2f."""
    return _lib().ltro_atan2f_selfcheck(seed, n)


def atan2f(y, x, libm=False):
    y, x = _f32(y), _f32(x)
    out = np.empty_like(y)
    (_lib().ltro_libm_atan2f if libm else _lib().ltro_atan2f)(y.ctypes.data, x.ctypes.data, out.ctypes.data, y.size)
    return out


def reset_rimg_size(alpha, vfov=50.0, hfov=360.0):
    r, c = ctypes.c_int(), ctypes.c_int()
    _lib().ltro_reset_rimg_size(vfov, hfov, alpha, ctypes.byref(r), ctypes.byref(c))
    return r.value, c.value


def pixel_index(xyz, rows, cols, vfov=50.0, hfov=360.0):
    xyz = _f32(xyz)
    n, stride = xyz.shape
    row = np.empty(n, np.int32); col = np.empty(n, np.int32); rng = np.empty(n, np.float32)
    _lib().ltro_pixel_index(xyz.ctypes.data, n, stride, vfov, hfov, rows, cols, row.ctypes.data, col.ctypes.data, rng.ctypes.data)
    return row, col, rng


def transform(xyzi, T, order=0):
    xyzi = _f32(xyzi); T = _f64(T)
    out = np.empty_like(xyzi)
    _lib().ltro_transform(xyzi.ctypes.data, len(xyzi), T.ctypes.data, order, out.ctypes.data)
    return out


def inverse4x4(T):
    T = _f64(T); out = np.empty((4, 4), np.float64)
    _lib().ltro_inverse4x4(T.ctypes.data, out.ctypes.data)
    return out


def inverse_poses(poses):
    return np.stack([inverse4x4(p) for p in poses]) if len(poses) else np.zeros((0, 4, 4))


def scan2rimg(xyzi, rows, cols, vfov=50.0, hfov=360.0):
    xyzi = _f32(xyzi); rimg = np.empty((rows, cols), np.float32)
    _lib().ltro_scan2rimg(xyzi.ctypes.data, len(xyzi), vfov, hfov, rows, cols, rimg.ctypes.data)
    return rimg


def map2rimg(xyzi, rows, cols, vfov=50.0, hfov=360.0):
    xyzi = _f32(xyzi); rimg = np.empty((rows, cols), np.float32); idx = np.empty((rows, cols), np.int32)
    _lib().ltro_map2rimg(xyzi.ctypes.data, len(xyzi), vfov, hfov, rows, cols, rimg.ctypes.data, idx.ctypes.data)
    return rimg, idx


def remove_pass(map_xyzi, scans_xyzi, offsets, inv_poses, mode, alpha, thres=0.1, vfov=50.0, hfov=360.0,
                lidar2base=IDENTITY, order=0, threads=None):
    """flags (N,) uint8, 1 = dynamic.  Restates Removerter.cpp:542-593 / 485-540 / 429-482."""
    m = _f32(map_xyzi); s = _f32(scans_xyzi); o = np.ascontiguousarray(offsets, np.int64); ip = _f64(inv_poses)
    l2b = _f64(lidar2base)
    flags = np.zeros(len(m), np.uint8)
    n = _lib().ltro_remove_pass(m.ctypes.data, len(m), s.ctypes.data, o.ctypes.data, ip.ctypes.data, len(o) - 1, vfov, hfov,
                                l2b.ctypes.data, order, mode, alpha, thres, threads or max_threads(), flags.ctypes.data)
    assert n == int(flags.sum())
    return flags


def parse_projected(map_xyzi, inv_pose, alpha=3.0, vfov=50.0, hfov=360.0, lidar2base=IDENTITY, order=0):
    """(points (V,4) in the keyframe's LiDAR frame, map indices (V,)).  Restates Session.cpp:353-357."""
    m = _f32(map_xyzi); ip = _f64(inv_pose); l2b = _f64(lidar2base)
    rows, cols = reset_rimg_size(alpha, vfov, hfov)
    cap = rows * cols
    out = np.empty((cap, 4), np.float32); idx = np.empty(cap, np.int32)
    n = _lib().ltro_parse_projected(m.ctypes.data, len(m), ip.ctypes.data, vfov, hfov, l2b.ctypes.data, order, alpha,
                                    out.ctypes.data, idx.ctypes.data, cap)
    return out[:n].copy(), idx[:n].copy()


def voxel(xyzi, leaf):
    """octreeDownsampling (utility.cpp:204-219)."""
    x = _f32(xyzi)
    out = np.empty((max(len(x), 1), 4), np.float32)
    n = _lib().ltro_voxel(x.ctypes.data, len(x), leaf, out.ctypes.data, len(out))
    if n < 0:
        raise ValueError("oracle voxel: unsupported input")
    return out[:n].copy()


def knn_dists(q_xyzi, t_xyzi, k, brute=False):
    q = _f32(q_xyzi); t = _f32(t_xyzi)
    out = np.empty((len(q), k), np.float32)
    _lib().ltro_knn_dists(q.ctypes.data, len(q), t.ctypes.data, len(t), k, out.ctypes.data, int(brute))
    return out


def knn_partition(scan_xyzi, pose, inv_pose, target_xyzi, k, thr, lidar2base=IDENTITY, order=0):
    """(labels (n,) uint8 1 = diff, coexist_local, diff_local).  Restates Session.cpp:537-607."""
    s = _f32(scan_xyzi); t = _f32(target_xyzi); P = _f64(pose); IP = _f64(inv_pose); l2b = _f64(lidar2base)
    n = len(s)
    labels = np.zeros(n, np.uint8); co = np.empty((max(n, 1), 4), np.float32); di = np.empty((max(n, 1), 4), np.float32)
    nd = _lib().ltro_knn_partition(s.ctypes.data, n, P.ctypes.data, IP.ctypes.data, t.ctypes.data, len(t), l2b.ctypes.data, order,
                                   k, thr, labels.ctypes.data, co.ctypes.data, di.ctypes.data)
    return labels, co[:n - nd].copy(), di[:nd].copy()


class Removerter:
    """Pipeline-level oracle mirroring ltremovert::Removerter (Removerter.cpp:1653-1678)."""

    def __init__(self, vfov=50.0, hfov=360.0, lidar2base=IDENTITY, order=0, num_knn=2, knn_thr=0.01, voxel=0.05,
                 threads=None, faithful=False, omp_cores=16, do_high_dyn_knn=True, schedule=None):
        self._h = _lib().ltro_create()
        l2b = _f64(lidar2base)
        _lib().ltro_set_params(self._h, vfov, hfov, l2b.ctypes.data, order, num_knn, knn_thr, voxel,
                               threads or max_threads(), int(faithful), omp_cores, int(do_high_dyn_knn))
        if schedule is not None:
            ops = np.array([s[0] for s in schedule], np.int32); res = np.array([s[1] for s in schedule], np.float32)
            _lib().ltro_set_schedule(self._h, ops.ctypes.data, res.ctypes.data, len(ops))

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().ltro_destroy(self._h)
            self._h = None

    def load_session(self, sess, xyzi, offsets, poses, inv_poses=None):
        x = _f32(xyzi); o = np.ascontiguousarray(offsets, np.int64); p = _f64(poses)
        ip = _f64(inv_poses) if inv_poses is not None else None
        _lib().ltro_load_session(self._h, sess, x.ctypes.data, o.ctypes.data, p.ctypes.data,
                                 ip.ctypes.data if ip is not None else None, len(o) - 1)

    def set_map(self, sess, name, xyzi):
        x = _f32(xyzi)
        _lib().ltro_set_map(self._h, sess, name.encode(), x.ctypes.data, len(x))

    def run(self, step0=True, step12=True, step3=False):
        _lib().ltro_run(self._h, (1 if step0 else 0) | (2 if step12 else 0) | (4 if step3 else 0))

    def stage(self, name):
        if _lib().ltro_stage(self._h, name.encode()) != 0:
            raise KeyError(name)

    def num_keyframes(self, sess):
        return _lib().ltro_num_keyframes(self._h, sess)

    def cloud(self, name, sess=0, kf=-1):
        n = _lib().ltro_cloud_size(self._h, name.encode(), sess, kf)
        if n < 0:
            raise KeyError(name)
        out = np.empty((n, 4), np.float32)
        _lib().ltro_cloud_copy(self._h, name.encode(), sess, kf, out.ctypes.data, n)
        return out

    def clouds(self, name, sess=0):
        return [self.cloud(name, sess, k) for k in range(self.num_keyframes(sess))]

    def log(self):
        out = []
        for i in range(_lib().ltro_log_count(self._h)):
            buf = ctypes.create_string_buffer(64); vals = np.zeros(4, np.int64)
            _lib().ltro_log_get(self._h, i, buf, 64, vals.ctypes.data)
            out.append((buf.value.decode(), *[int(v) for v in vals]))
        return out

    def timing(self, key):
        return _lib().ltro_timing(self._h, key.encode())
