"""Python view of the host orchestrator (include/ltr_removert.h, libltr_removert.so).

`Removerter` mirrors ltremovert::Removerter (ltremovert/src/Removerter.cpp:1653-1678): load two sessions,
run Step 0 / Step 1+2 / Step 3, read back named clouds.  All work happens in the C++ orchestrator and the
CUDA library underneath; this module only moves numpy arrays across the C-ABI and, for multi-GPU runs,
provides the collective hooks on top of torch.distributed (NCCL).
"""
import ctypes
import os

import numpy as np

from . import binding

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libltr_removert.so")
_LIB = None

OP_REMOVE, OP_REVERT = 0, 1
MAX_SCHEDULE = 32


class Params(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("sequence_vfov", ctypes.c_float), ("sequence_hfov", ctypes.c_float),
                ("ExtrinsicLiDARtoPoseBase", ctypes.c_double * 16), ("num_nn_points_within", ctypes.c_int32),
                ("dist_nn_points_within", ctypes.c_float), ("downsample_voxel_size", ctypes.c_float),
                ("n_schedule", ctypes.c_int32), ("schedule_op", ctypes.c_int32 * MAX_SCHEDULE),
                ("schedule_res", ctypes.c_float * MAX_SCHEDULE), ("extract_high_dyn_knn", ctypes.c_int32),
                ("transform_order", ctypes.c_int32), ("keyframe_batch", ctypes.c_int32), ("fast_path", ctypes.c_int32)]


_ALLREDUCE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)
_ALLGATHER_I64 = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64))
_ALLGATHERV = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                               ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64))


class Comm(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("user", ctypes.c_void_p),
                ("allreduce_max_u8", _ALLREDUCE), ("allgather_i64", _ALLGATHER_I64), ("allgatherv_f32", _ALLGATHERV)]


HOST_EXPORTS = ["ltrh_params_default", "ltrh_create", "ltrh_destroy", "ltrh_last_error", "ltrh_set_comm", "ltrh_context", "ltrh_comm_init_nccl", "ltrh_owns_session", "ltrh_invert_poses",
                "ltrh_load_session", "ltrh_run_step0", "ltrh_run_step12", "ltrh_run_step3", "ltrh_reset_to_step0", "ltrh_cascade_promote_updated", "ltrh_stage", "ltrh_cloud",
                "ltrh_scanset", "ltrh_timing", "ltrh_log_count", "ltrh_log_get", "ltrh_io_last_error", "ltrh_io_read_pcd", "ltrh_io_write_pcd",
                "ltrh_io_read_poses", "ltrh_io_parse_keyframes", "ltrh_io_parse_keyframes_in_roi", "ltrh_io_voxel_grid", "ltrh_io_voxel_grid_scans", "ltrh_io_yaml_get"]


def host_lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    binding.lib()  # the device library must be loadable first (fails loudly if it has not been built)
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
    L = ctypes.CDLL(HOST_LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    P = ctypes.POINTER
    L.ltrh_params_default.argtypes = [P(Params)]
    L.ltrh_params_default.restype = None
    L.ltrh_create.argtypes = [P(vp), P(Params)]
    L.ltrh_destroy.argtypes = [vp]
    L.ltrh_destroy.restype = None
    L.ltrh_last_error.argtypes = [vp]
    L.ltrh_last_error.restype = ctypes.c_char_p
    L.ltrh_set_comm.argtypes = [vp, P(Comm)]
    L.ltrh_context.argtypes = [vp]
    L.ltrh_context.restype = vp
    L.ltrh_comm_init_nccl.argtypes = [vp, vp, i32, i32, i32]
    L.ltrh_owns_session.argtypes = [vp, i32]
    L.ltrh_invert_poses.argtypes = [vp, i32, vp]
    L.ltrh_invert_poses.restype = None
    L.ltrh_load_session.argtypes = [vp, i32, vp, vp, vp, vp, i32]
    for f in ("ltrh_run_step0", "ltrh_run_step12", "ltrh_run_step3", "ltrh_reset_to_step0", "ltrh_cascade_promote_updated"):
        getattr(L, f).argtypes = [vp]
    L.ltrh_stage.argtypes = [vp, ctypes.c_char_p]
    L.ltrh_cloud.argtypes = [vp, ctypes.c_char_p, i32, P(i32)]
    L.ltrh_scanset.argtypes = [vp, ctypes.c_char_p, i32, P(i32)]
    L.ltrh_timing.argtypes = [vp, ctypes.c_char_p]
    L.ltrh_timing.restype = ctypes.c_double
    L.ltrh_log_count.argtypes = [vp]
    L.ltrh_log_get.argtypes = [vp, i32, ctypes.c_char_p, i32, vp]
    L.ltrh_io_last_error.restype = ctypes.c_char_p
    L.ltrh_io_read_pcd.argtypes = [ctypes.c_char_p, vp, i64]
    L.ltrh_io_read_pcd.restype = i64
    L.ltrh_io_write_pcd.argtypes = [ctypes.c_char_p, vp, i64, i32]
    L.ltrh_io_read_poses.argtypes = [ctypes.c_char_p, vp, i32]
    L.ltrh_io_parse_keyframes.argtypes = [i32, i32, i32, i32, vp, i32]
    L.ltrh_io_parse_keyframes_in_roi.argtypes = [vp, i32, vp, i32, i32, vp, i32]
    L.ltrh_io_voxel_grid.argtypes = [vp, i64, ctypes.c_float, vp, i64, P(i32)]
    L.ltrh_io_voxel_grid.restype = i64
    L.ltrh_io_voxel_grid_scans.argtypes = [vp, vp, i32, ctypes.c_float, vp, i64, vp]
    L.ltrh_io_voxel_grid_scans.restype = i64
    L.ltrh_io_yaml_get.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, i32, vp, i32, P(i32)]
    _LIB = L
    return L


class _DevArray:
    """Minimal __cuda_array_interface__ wrapper so torch can alias raw device memory without a copy."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class TorchDistComm:
    """ltr_comm hooks on top of torch.distributed (NCCL for CUDA tensors; gloo works through host staging in tests)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.trace = os.environ.get("LTR_COMM_TRACE") == "1"
        self.stats = {"allreduce": [0, 0.0, 0], "allgather_i64": [0, 0.0, 0], "allgatherv": [0, 0.0, 0]}   # calls, seconds, bytes
        self._cb = (_ALLREDUCE(self._timed("allreduce", self._allreduce)), _ALLGATHER_I64(self._timed("allgather_i64", self._allgather_i64)),
                    _ALLGATHERV(self._timed("allgatherv", self._allgatherv)))
        self.struct = Comm(self.rank, self.world, None, *self._cb)

    def _timed(self, name, fn):
        import time

        def wrapped(*a):
            t0 = time.perf_counter()
            rc = fn(*a)
            st = self.stats[name]
            st[0] += 1
            st[1] += time.perf_counter() - t0
            st[2] += int(a[2]) * (1 if name == "allreduce" else 4) if name != "allgather_i64" else 8
            return rc
        return wrapped

    def _tensor(self, ptr, n, typestr, dtype):
        if self.device.type == "cuda":
            return self.torch.as_tensor(_DevArray(ptr, n, typestr), device=self.device)
        # host memory (gloo tests of the hook logic): alias the buffer through numpy
        ct = ctypes.c_uint8 if typestr == "|u1" else ctypes.c_float
        arr = np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(ptr), ctypes.POINTER(ct)), shape=(int(n),))
        return self.torch.from_numpy(arr)

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize()

    def _allreduce(self, user, dev, n):
        try:
            t = self._tensor(dev, n, "|u1", self.torch.uint8)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001 -- must not propagate through the C frame
            print("allreduce hook failed:", e)
            return 1

    def _allgather_i64(self, user, local, out):
        try:
            t = self.torch.tensor([local], dtype=self.torch.int64, device=self.device)
            g = [self.torch.zeros_like(t) for _ in range(self.world)]
            self.dist.all_gather(g, t, group=self.group)
            for r in range(self.world):
                out[r] = int(g[r].item())
            return 0
        except Exception as e:  # noqa: BLE001
            print("allgather_i64 hook failed:", e)
            return 1

    def _allgatherv(self, user, src, n_local, dst, counts, displs):
        try:
            torch = self.torch
            cnt = [int(counts[r]) for r in range(self.world)]
            total = sum(cnt)
            if total == 0:
                return 0
            mx = max(cnt)
            send = torch.zeros(mx, dtype=torch.float32, device=self.device)
            if n_local > 0:
                send[:n_local] = self._tensor(src, n_local, "<f4", torch.float32)
            recv = torch.empty(mx * self.world, dtype=torch.float32, device=self.device)
            if self.backend == "gloo":
                self.dist.all_gather(list(recv.view(self.world, mx).unbind(0)), send, group=self.group)
            else:
                self.dist.all_gather_into_tensor(recv, send, group=self.group)
            out = self._tensor(dst, total, "<f4", torch.float32)
            for r in range(self.world):
                if cnt[r]:
                    d = int(displs[r])
                    out[d:d + cnt[r]] = recv[r * mx:r * mx + cnt[r]]
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001
            print("allgatherv hook failed:", e)
            return 1


def inverse_poses(poses):
    """Inverse keyframe poses as the reference computes them (general 4x4 cofactor inverse, Session.cpp:110); np.linalg.inv differs in the
    last bits, which is enough to move a few emitted coordinates by one ulp."""
    p = np.ascontiguousarray(poses, np.float64).reshape(-1, 4, 4)
    out = np.empty_like(p)
    host_lib().ltrh_invert_poses(p.ctypes.data, len(p), out.ctypes.data)
    return out


def nccl_unique_id():
    """128 bytes of ncclGetUniqueId (call on ONE rank, hand the bytes to the others through the launcher's own channel)."""
    buf = ctypes.create_string_buffer(128)
    rc = binding.lib().ltr_nccl_unique_id(buf)
    if rc != 0:
        raise binding.LtrError(rc, binding.lib().ltr_last_error(None).decode())
    return bytes(buf.raw)


def selfremovert_schedule(resolutions):
    """Removerter::selfRemovert (Removerter.cpp:1378-1393): remove(r), revert(0.95 r), remove(r) per resolution.  The revert
    resolution is `0.95 * _res_alpha` with a float argument, i.e. a DOUBLE product narrowed once when passed on as float (:1385)."""
    s = []
    for r in resolutions:
        s += [(OP_REMOVE, r), (OP_REVERT, float(np.float32(0.95 * float(np.float32(r))))), (OP_REMOVE, r)]
    return s


class Removerter:
    def __init__(self, device=0, vfov=50.0, hfov=360.0, lidar2base=None, num_knn=2, knn_thr=0.01, voxel=0.05,
                 schedule=((OP_REMOVE, 2.5),), extract_high_dyn_knn=True, transform_order=0, keyframe_batch=0,
                 fast_path=True, cull=True, comm=None):
        L = host_lib()
        p = Params()
        L.ltrh_params_default(ctypes.byref(p))
        p.device = device
        p.sequence_vfov, p.sequence_hfov = vfov, hfov
        if lidar2base is not None:
            l2b = np.ascontiguousarray(lidar2base, np.float64).reshape(16)
            for i in range(16):
                p.ExtrinsicLiDARtoPoseBase[i] = l2b[i]
        p.num_nn_points_within, p.dist_nn_points_within, p.downsample_voxel_size = num_knn, knn_thr, voxel
        schedule = list(schedule)
        p.n_schedule = len(schedule)
        for i, (op, res) in enumerate(schedule):
            p.schedule_op[i], p.schedule_res[i] = op, res
        p.extract_high_dyn_knn = int(extract_high_dyn_knn)
        p.transform_order, p.keyframe_batch, p.fast_path = transform_order, keyframe_batch, ((2 if cull else 1) if fast_path else 0)
        self._h = ctypes.c_void_p()
        rc = L.ltrh_create(ctypes.byref(self._h), ctypes.byref(p))
        if rc != 0:
            msg = L.ltrh_last_error(None).decode()
            self._h = None
            raise binding.LtrError(rc, msg)
        self._comm = comm
        if comm is not None:
            self._ck(L.ltrh_set_comm(self._h, ctypes.byref(comm.struct)))
        # a binding.Context view that shares the orchestrator's ltr_ctx (not owned)
        self.ctx = binding.Context.__new__(binding.Context)
        self.ctx._h = ctypes.c_void_p(L.ltrh_context(self._h))
        self.K = [0, 0]
        self.rank, self.world, self.split = 0, 1, False

    def init_nccl(self, id128, rank, world, split_sessions=True):
        """Native multi-GPU transport (ltrh_comm_init_nccl): every rank passes the same 128-byte id (nccl_unique_id() of one rank)."""
        self._ck(host_lib().ltrh_comm_init_nccl(self._h, id128, rank, world, 1 if split_sessions else 0))
        self.rank, self.world = rank, world
        self.split = bool(split_sessions) and world >= 2 and world % 2 == 0

    def owns(self, sess):
        return bool(host_lib().ltrh_owns_session(self._h, sess))

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._h = None
            try:
                host_lib().ltrh_destroy(self._h)
            except TypeError:  # interpreter shutdown
                pass
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise binding.LtrError(rc, host_lib().ltrh_last_error(self._h).decode())

    def load_session(self, sess, xyzi, offsets, poses, inv_poses):
        x = np.ascontiguousarray(xyzi, np.float32)
        o = np.ascontiguousarray(offsets, np.int64)
        p = np.ascontiguousarray(poses, np.float64)
        ip = np.ascontiguousarray(inv_poses, np.float64)
        self._ck(host_lib().ltrh_load_session(self._h, sess, x.ctypes.data, o.ctypes.data, p.ctypes.data, ip.ctypes.data, len(o) - 1))
        self.K[sess] = len(o) - 1

    def run_step0(self):
        self._ck(host_lib().ltrh_run_step0(self._h))

    def run_step12(self):
        self._ck(host_lib().ltrh_run_step12(self._h))

    def run_step3(self):
        self._ck(host_lib().ltrh_run_step3(self._h))

    def reset_to_step0(self):
        self._ck(host_lib().ltrh_reset_to_step0(self._h))

    def cascade_promote_updated(self):
        """LT-map cascade: the updated scans become the central session of the next run (ltrh_cascade_promote_updated)."""
        self._ck(host_lib().ltrh_cascade_promote_updated(self._h))

    def stage(self, name):
        self._ck(host_lib().ltrh_stage(self._h, name.encode()))

    def cloud_handle(self, name, sess=0):
        h = ctypes.c_int32()
        self._ck(host_lib().ltrh_cloud(self._h, name.encode(), sess, ctypes.byref(h)))
        return h.value

    def scanset_handle(self, name, sess=0):
        h = ctypes.c_int32()
        self._ck(host_lib().ltrh_scanset(self._h, name.encode(), sess, ctypes.byref(h)))
        return h.value

    def cloud(self, name, sess=0, out=None):
        return self.ctx.cloud_download(self.cloud_handle(name, sess), out=out)

    def cloud_size(self, name, sess=0):
        return self.ctx.cloud_size(self.cloud_handle(name, sess))

    def scanset(self, name, sess=0):
        return self.ctx.scanset_download(self.scanset_handle(name, sess))

    def timing(self, key):
        return host_lib().ltrh_timing(self._h, key.encode())

    def log(self):
        out = []
        for i in range(host_lib().ltrh_log_count(self._h)):
            buf = ctypes.create_string_buffer(64)
            vals = np.zeros(4, np.int64)
            host_lib().ltrh_log_get(self._h, i, buf, 64, vals.ctypes.data)
            out.append((buf.value.decode(), *[int(v) for v in vals]))
        return out


# ---- file-level helpers (include/ltr_removert.h ltrh_io_*): thin ctypes views used by tests and tools ----
def read_pcd(path):
    L = host_lib()
    n = L.ltrh_io_read_pcd(path.encode(), None, 0)
    if n < 0:
        raise IOError(L.ltrh_io_last_error().decode())
    out = np.empty((n, 4), np.float32)
    L.ltrh_io_read_pcd(path.encode(), out.ctypes.data, n)
    return out


def write_pcd(path, xyzi, octree_layout=False):
    x = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    if host_lib().ltrh_io_write_pcd(path.encode(), x.ctypes.data, len(x), int(octree_layout)) != 0:
        raise IOError(host_lib().ltrh_io_last_error().decode())


def read_poses(path):
    L = host_lib()
    n = L.ltrh_io_read_poses(path.encode(), None, 0)
    if n < 0:
        raise IOError(L.ltrh_io_last_error().decode())
    out = np.empty((n, 4, 4), np.float64)
    L.ltrh_io_read_poses(path.encode(), out.ctypes.data, n)
    return out


def parse_keyframes(num_scans, start_idx, end_idx, gap):
    out = np.empty(max(num_scans, 1), np.int32)
    n = host_lib().ltrh_io_parse_keyframes(num_scans, start_idx, end_idx, gap, out.ctypes.data, len(out))
    return out[:n].copy()


def parse_keyframes_in_roi(scan_poses, roi_poses, gap):
    a = np.ascontiguousarray(scan_poses, np.float64).reshape(-1, 16)
    b = np.ascontiguousarray(roi_poses, np.float64).reshape(-1, 16)
    out = np.empty(max(len(a), 1), np.int32)
    n = host_lib().ltrh_io_parse_keyframes_in_roi(a.ctypes.data, len(a), b.ctypes.data, len(b), gap, out.ctypes.data, len(out))
    return out[:n].copy()


def voxel_grid(xyzi, leaf):
    x = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    out = np.empty((max(len(x), 1), 4), np.float32)
    ov = ctypes.c_int32()
    n = host_lib().ltrh_io_voxel_grid(x.ctypes.data, len(x), leaf, out.ctypes.data, len(out), ctypes.byref(ov))
    return out[:n].copy(), bool(ov.value)


def voxel_grid_scans(xyzi, offsets, leaf):
    """The load-time filter over all scans of a session at once (scan k = rows [offsets[k], offsets[k+1])).  Returns (points, offsets)."""
    x = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    off = np.ascontiguousarray(offsets, np.int64)
    K = len(off) - 1
    out_off = np.zeros(K + 1, np.int64)
    out = np.empty((max(len(x), 1), 4), np.float32)     # the filter never grows a scan
    n = host_lib().ltrh_io_voxel_grid_scans(x.ctypes.data, off.ctypes.data, K, leaf, out.ctypes.data, len(out), out_off.ctypes.data)
    if n < 0:
        raise ValueError("ltrh_io_voxel_grid_scans: bad argument")
    return out[:n].copy(), out_off


def yaml_get(path, key):
    buf = ctypes.create_string_buffer(4096)
    lst = np.zeros(64, np.float64)
    n = ctypes.c_int32()
    if host_lib().ltrh_io_yaml_get(path.encode(), key.encode(), buf, 4096, lst.ctypes.data, 64, ctypes.byref(n)) != 0:
        raise IOError(host_lib().ltrh_io_last_error().decode())
    return buf.value.decode(), lst[:n.value].copy()
