"""Deterministic synthetic two-session LiDAR data.  Data generator only.

`make_session` returns scans in the form the reference holds them after loading
(ltremovert/src/Session.cpp:266-302): per keyframe an (n, 4) float32 array x, y, z, intensity in the
LiDAR frame, and a 4x4 float64 pose (row-major, LiDAR -> world, ltremovert/src/Session.cpp:102-114).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
SEED = 20220523


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libltr_synth.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = ctypes.CDLL(path)
        _LIB.ltr_synth_session.restype = ctypes.c_int64
        _LIB.ltr_synth_session.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


class SessionData:
    """K keyframe scans (concatenated) + poses of one session."""

    def __init__(self, xyzi, offsets, poses):
        self.xyzi = xyzi            # (M, 4) float32
        self.offsets = offsets      # (K + 1,) int64
        self.poses = poses          # (K, 4, 4) float64

    @property
    def K(self):
        return len(self.offsets) - 1

    def scan(self, k):
        return self.xyzi[self.offsets[k]:self.offsets[k + 1]]

    def subset(self, k0, k1):
        o = self.offsets[k0:k1 + 1]
        return SessionData(self.xyzi[o[0]:o[-1]].copy(), (o - o[0]).copy(), self.poses[k0:k1].copy())


def make_session(session, K, seed=SEED, beams=64, az_steps=1800, max_range=100.0, noise=0.02, spacing=1.0,
                 n_cars=200, n_poles=40, n_movers=10, threads=None, k0=0):
    """Keyframes k0 .. k0+K-1 of `session` (0 = central, 1 = query)."""
    lib = _lib()
    threads = threads or os.cpu_count() or 1
    cap = K * beams * az_steps
    xyzi = np.empty((cap, 4), dtype=np.float32)
    offsets = np.zeros(K + 1, dtype=np.int64)
    poses = np.zeros((K, 4, 4), dtype=np.float64)
    n = lib.ltr_synth_session(seed, session, k0, K, beams, az_steps, max_range, noise, spacing, n_cars, n_poles, n_movers,
                              threads, xyzi.ctypes.data, offsets.ctypes.data, poses.ctypes.data)
    return SessionData(xyzi[:n].copy() if n < cap else xyzi, offsets, poses)


def make_pair(K, **kw):
    """(central, query) sessions of K keyframes each."""
    return make_session(0, K, **kw), make_session(1, K, **kw)
