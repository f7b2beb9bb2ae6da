"""ORACLE / REFERENCE SHIM (test infrastructure, not product code).

ctypes view of oracle/_ref/libltremovert_ref.so: the reference's own ltremovert translation units, compiled unmodified
from /root/reference against the stand-in third-party headers under oracle/ref_shim/include (see ltr_shim_core.h for what
that does and does not pin).  Used by tests/ to pin the oracle and to generate the golden fixtures under tests/golden/.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libltremovert_ref.so")
LIB_OMP_PATH = os.path.join(HERE, "_ref", "libltremovert_ref_omp.so")
LIB_DROPIN_PATH = os.path.join(HERE, "_ref", "libltremovert_dropin.so")   # reference loaders / writers around libltr_removert.so

_lib_cache = {}


def _path(omp):
    return LIB_DROPIN_PATH if omp == "dropin" else LIB_OMP_PATH if omp else LIB_PATH


def available(omp=False):
    """omp: False = deterministic parity build, True = OpenMP timing build, "dropin" = the drop-in demonstration library."""
    return os.path.exists(_path(omp))


def _lib(omp=False):
    if omp not in _lib_cache:
        L = ctypes.CDLL(_path(omp))
        if omp == "dropin":
            L.ref_dropin_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        i64, f32, vp, cp, ci = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int
        L.ref_create.restype = vp
        L.ref_destroy.argtypes = [vp]
        L.ref_run.argtypes = [vp]
        L.ref_stage.argtypes = [vp, cp]
        L.ref_pass.argtypes = [vp, ci, ci, ci, f32]
        L.ref_self_removert.argtypes = [vp, ci, ci]
        L.ref_high_dyn_with_schedule.argtypes = [vp, vp, vp, ci]
        L.ref_param_num.argtypes = [cp, vp, ci]
        L.ref_param_str.argtypes = [cp, cp]
        L.ref_cart2sph.argtypes = [vp, i64, vp]
        L.ref_rad2deg.argtypes = [vp, i64, vp]
        L.ref_reset_rimg_size.argtypes = [f32, f32, f32, vp, vp]
        L.ref_map2rimg.argtypes = [vp, i64, f32, f32, ci, ci, vp, vp]
        L.ref_parse_projected.argtypes = [vp, i64, f32, f32, ci, ci, vp, i64]; L.ref_parse_projected.restype = i64
        L.ref_transform_global_to_local.argtypes = [vp, i64, vp, vp, vp]
        L.ref_octree_downsampling.argtypes = [vp, i64, f32, vp, i64]; L.ref_octree_downsampling.restype = i64
        L.ref_linspace_int.argtypes = [ci, ci, ci, vp]
        L.ref_inverse4x4.argtypes = [vp, vp]
        L.ref_scan2rimg.argtypes = [vp, vp, i64, ci, ci, vp]; L.ref_scan2rimg.restype = i64
        L.ref_dynamic_idx.argtypes = [vp, ci, ci, ci, ci, ci, vp, i64]; L.ref_dynamic_idx.restype = i64
        L.ref_static_idx.argtypes = [vp, vp, i64, ci, vp, i64]; L.ref_static_idx.restype = i64
        L.ref_num_keyframes.argtypes = [vp, ci]
        L.ref_num_scans.argtypes = [vp, ci]
        L.ref_keyframe_name.argtypes = [vp, ci, ci, cp, ci]
        L.ref_keyframe_pose.argtypes = [vp, ci, ci, vp, vp]
        L.ref_extrinsics.argtypes = [vp, vp, vp]
        L.ref_cloud.argtypes = [vp, ci, cp, vp, i64]; L.ref_cloud.restype = i64
        L.ref_set_cloud.argtypes = [vp, ci, cp, vp, i64]
        L.ref_scans_count.argtypes = [vp, ci, cp]
        L.ref_scan.argtypes = [vp, ci, cp, ci, vp, i64]; L.ref_scan.restype = i64
        L.ref_set_scans.argtypes = [vp, ci, cp, vp, vp, ci]
        L.ref_extract_knn_diff.argtypes = [vp, ci, ci, vp, i64, ci, f32]
        L.ref_load_session_mem.argtypes = [vp, ci, vp, vp, vp, ci]
        L.ref_saved_get.argtypes = [ci, cp, ci, vp, i64]; L.ref_saved_get.restype = i64
        f64 = ctypes.c_double
        L.ref_time_dynamic_idx.argtypes = [vp, ci, ci, ci, ci, ci, vp]; L.ref_time_dynamic_idx.restype = f64
        L.ref_time_parse_static.argtypes = [vp, ci]; L.ref_time_parse_static.restype = f64
        L.ref_knn_set_target.argtypes = [vp, ci, vp, i64, ci, f32]; L.ref_knn_set_target.restype = f64
        L.ref_time_knn_queries.argtypes = [vp, ci, ci, ci, vp]; L.ref_time_knn_queries.restype = f64
        L.ref_time_merge.argtypes = [vp, ci, cp, vp]; L.ref_time_merge.restype = f64
        _lib_cache[omp] = L
    return _lib_cache[omp]


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def set_params(params, omp=False):
    """params: dict name (without the 'removert/' prefix) -> bool | int | float | str | list of numbers."""
    L = _lib(omp)
    L.ref_params_clear()
    for k, v in params.items():
        name = ("removert/" + k).encode()
        if isinstance(v, str):
            L.ref_param_str(name, v.encode())
        else:
            a = np.atleast_1d(np.asarray(v, np.float64)).ravel()
            L.ref_param_num(name, a.ctypes.data, len(a))


# ---- free functions of utility.cpp -------------------------------------------------------------------------------
def cart2sph(xyz):
    x = _f32(xyz)[:, :3].copy(); out = np.empty((len(x), 3), np.float32)
    _lib().ref_cart2sph(x.ctypes.data, len(x), out.ctypes.data)
    return out


def rad2deg(a):
    a = _f32(a); out = np.empty_like(a)
    _lib().ref_rad2deg(a.ctypes.data, a.size, out.ctypes.data)
    return out


def reset_rimg_size(alpha, vfov=50.0, hfov=360.0):
    r = ctypes.c_int(); c = ctypes.c_int()
    _lib().ref_reset_rimg_size(vfov, hfov, alpha, ctypes.byref(r), ctypes.byref(c))
    return r.value, c.value


def map2rimg(xyzi, rows, cols, vfov=50.0, hfov=360.0):
    x = _f32(xyzi); rimg = np.empty((rows, cols), np.float32); idx = np.empty((rows, cols), np.int32)
    _lib().ref_map2rimg(x.ctypes.data, len(x), vfov, hfov, rows, cols, rimg.ctypes.data, idx.ctypes.data)
    return rimg, idx


def parse_projected(xyzi, rows, cols, vfov=50.0, hfov=360.0):
    x = _f32(xyzi); out = np.empty((rows * cols, 4), np.float32)
    n = _lib().ref_parse_projected(x.ctypes.data, len(x), vfov, hfov, rows, cols, out.ctypes.data, len(out))
    return out[:n].copy()


def transform_global_to_local(xyzi, inv_pose, base2lidar):
    x = _f32(xyzi); a = np.ascontiguousarray(inv_pose, np.float64); b = np.ascontiguousarray(base2lidar, np.float64)
    out = np.empty_like(x)
    _lib().ref_transform_global_to_local(x.ctypes.data, len(x), a.ctypes.data, b.ctypes.data, out.ctypes.data)
    return out


def linspace_int(a, b, n):
    out = np.empty(n, np.int32)
    _lib().ref_linspace_int(a, b, n, out.ctypes.data)
    return out


def inverse4x4(T):
    T = np.ascontiguousarray(T, np.float64); out = np.empty((4, 4), np.float64)
    _lib().ref_inverse4x4(T.ctypes.data, out.ctypes.data)
    return out


# ---- the node ----------------------------------------------------------------------------------------------------
class Removerter:
    """ltremovert::Removerter of the reference (Removerter.h), constructed from the current parameter set."""

    def __init__(self, params, transform_order=0, omp=False, verbose=False, write_files=True):
        self._omp = omp
        L = _lib(omp)
        set_params(params, omp)
        L.ref_set_transform_order(transform_order)
        L.ref_set_verbose(int(verbose))
        L.ref_set_write_files(int(write_files))
        L.ref_saved_clear()
        self._h = L.ref_create()

    def close(self):
        if getattr(self, "_h", None):
            _lib(self._omp).ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self):
        _lib(self._omp).ref_run(self._h)

    def dropin_run(self, transform_order=0):
        """omp="dropin" only: the node's own loaders and writers with Steps 0-3 on the B200 library (oracle/ref_shim/dropin_capi.cpp)."""
        err = ctypes.create_string_buffer(1024)
        rc = _lib(self._omp).ref_dropin_run(self._h, transform_order, err, 1024)
        if rc != 0:
            raise RuntimeError(f"drop-in run failed ({rc}): {err.value.decode()}")

    def stage(self, name):
        if _lib(self._omp).ref_stage(self._h, name.encode()) != 0:
            raise KeyError(name)

    OPS = {"removeOnce": 0, "revertOnce": 1, "resetAsDynamic": 2, "resetAsStatic": 3, "iremoveOnceForND": 4, "removeOnceForPD": 5}

    def op(self, name, target=0, source=None, res=0.0):
        _lib(self._omp).ref_pass(self._h, self.OPS[name], target, target if source is None else source, res)

    def high_dyn_with_schedule(self, schedule):
        """Step 1 with an explicit [(op, res)] schedule (op 0 remove, 1 revert) + the rest of removeHighDynamicPoints."""
        ops = np.array([s[0] for s in schedule], np.int32); res = np.array([s[1] for s in schedule], np.float32)
        _lib(self._omp).ref_high_dyn_with_schedule(self._h, ops.ctypes.data, res.ctypes.data, len(ops))

    def self_removert(self, sess, repeat):
        _lib(self._omp).ref_self_removert(self._h, sess, repeat)

    def load_session_mem(self, sess, xyzi, offsets, poses):
        x = _f32(xyzi); o = np.ascontiguousarray(offsets, np.int64); p = np.ascontiguousarray(poses, np.float64)
        _lib(self._omp).ref_load_session_mem(self._h, sess, x.ctypes.data, o.ctypes.data, p.ctypes.data, len(o) - 1)

    def scan2rimg(self, xyzi, rows, cols):
        x = _f32(xyzi); out = np.empty((rows, cols), np.float32)
        _lib(self._omp).ref_scan2rimg(self._h, x.ctypes.data, len(x), rows, cols, out.ctypes.data)
        return out

    def dynamic_idx(self, mode, target, source, rows, cols, cap):
        out = np.empty(cap, np.int32)
        n = _lib(self._omp).ref_dynamic_idx(self._h, mode, target, source, rows, cols, out.ctypes.data, cap)
        return out[:n].copy()

    def static_idx(self, dyn, num_all):
        d = np.ascontiguousarray(dyn, np.int32); out = np.empty(num_all + 2, np.int32)
        n = _lib(self._omp).ref_static_idx(self._h, d.ctypes.data, len(d), num_all, out.ctypes.data, len(out))
        return out[:n].copy()

    def num_keyframes(self, sess):
        return _lib(self._omp).ref_num_keyframes(self._h, sess)

    def num_scans(self, sess):
        return _lib(self._omp).ref_num_scans(self._h, sess)

    def keyframe_names(self, sess):
        out = []
        for k in range(self.num_keyframes(sess)):
            buf = ctypes.create_string_buffer(256)
            _lib(self._omp).ref_keyframe_name(self._h, sess, k, buf, 256)
            out.append(buf.value.decode())
        return out

    def keyframe_poses(self, sess):
        K = self.num_keyframes(sess)
        P = np.empty((K, 4, 4), np.float64); IP = np.empty((K, 4, 4), np.float64)
        for k in range(K):
            _lib(self._omp).ref_keyframe_pose(self._h, sess, k, P[k].ctypes.data, IP[k].ctypes.data)
        return P, IP

    def extrinsics(self):
        a = np.empty((4, 4), np.float64); b = np.empty((4, 4), np.float64)
        _lib(self._omp).ref_extrinsics(self._h, a.ctypes.data, b.ctypes.data)
        return a, b

    def cloud(self, name, sess=0):
        L = _lib(self._omp)
        n = L.ref_cloud(self._h, sess, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty((n, 4), np.float32)
        L.ref_cloud(self._h, sess, name.encode(), out.ctypes.data, n)
        return out

    def set_cloud(self, name, xyzi, sess=0):
        x = _f32(xyzi)
        if _lib(self._omp).ref_set_cloud(self._h, sess, name.encode(), x.ctypes.data, len(x)) != 0:
            raise KeyError(name)

    def set_scans(self, name, xyzi, offsets, sess=0):
        x = _f32(xyzi); o = np.ascontiguousarray(offsets, np.int64)
        if _lib(self._omp).ref_set_scans(self._h, sess, name.encode(), x.ctypes.data, o.ctypes.data, len(o) - 1) != 0:
            raise KeyError(name)

    def extract_knn_diff(self, sess, target_xyzi, k, thr, low=True):
        """Session::extractLowDynPointsViaKnnDiff (low) / extractHighDynPointsViaKnnDiff of `sess` against target_xyzi."""
        t = _f32(target_xyzi)
        _lib(self._omp).ref_extract_knn_diff(self._h, sess, int(low), t.ctypes.data, len(t), k, thr)

    # ---- timed per-keyframe loops (bench.py --impl reference) ----
    def time_dynamic_idx(self, mode, target, source, rows, cols):
        n = ctypes.c_int64()
        return _lib(self._omp).ref_time_dynamic_idx(self._h, mode, target, source, rows, cols, ctypes.byref(n)), n.value

    def time_parse_static(self, sess):
        return _lib(self._omp).ref_time_parse_static(self._h, sess)

    def knn_set_target(self, sess, target_xyzi, k, thr):
        t = _f32(target_xyzi)
        return _lib(self._omp).ref_knn_set_target(self._h, sess, t.ctypes.data, len(t), k, thr)

    def time_knn_queries(self, sess, low, omp_cores):
        n = ctypes.c_int64()
        return _lib(self._omp).ref_time_knn_queries(self._h, sess, int(low), omp_cores, ctypes.byref(n)), n.value

    def time_merge(self, sess, name):
        n = ctypes.c_int64()
        return _lib(self._omp).ref_time_merge(self._h, sess, name.encode(), ctypes.byref(n)), n.value

    def scans(self, name, sess=0):
        L = _lib(self._omp)
        cnt = L.ref_scans_count(self._h, sess, name.encode())
        if cnt < 0:
            raise KeyError(name)
        out = []
        for k in range(cnt):
            n = L.ref_scan(self._h, sess, name.encode(), k, None, 0)
            a = np.empty((n, 4), np.float32)
            L.ref_scan(self._h, sess, name.encode(), k, a.ctypes.data, n)
            out.append(a)
        return out

    def saved(self):
        """[(path, (n,4) float32)] of every pcl::io::savePCDFileBinary call so far, in call order."""
        L = _lib(self._omp)
        out = []
        for i in range(L.ref_saved_count()):
            buf = ctypes.create_string_buffer(1024)
            n = L.ref_saved_get(i, buf, 1024, None, 0)
            a = np.empty((n, 4), np.float32)
            L.ref_saved_get(i, buf, 1024, a.ctypes.data, n)
            out.append((buf.value.decode(), a))
        return out
