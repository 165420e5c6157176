"""ctypes binding of oracle/liboracle.so (oracle_cpu.cpp) — TEST INFRASTRUCTURE ONLY.

The double-precision C++ restatement of the reference's CPU pseudo-likelihood
path.  Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs; never by anything under lite_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# keep in sync with lite_b200.capi._SoA / include/lite_b200.h (lite_tess_soa)
_FIELDS = [
    ("n_vertices", C.c_int32), ("vx", C.c_void_p), ("vy", C.c_void_p),
    ("n_halfedges", C.c_int32),
    ("he_src", C.c_void_p), ("he_twin", C.c_void_p), ("he_next", C.c_void_p),
    ("he_face", C.c_void_p), ("he_seg", C.c_void_p),
    ("he_next_hf", C.c_void_p), ("he_prev_hf", C.c_void_p),
    ("he_dir", C.c_void_p), ("he_length", C.c_void_p),
    ("n_faces", C.c_int32), ("face_inside", C.c_void_p),
    ("face_he_offset", C.c_void_p), ("face_he_count", C.c_void_p),
    ("n_face_he", C.c_int32), ("face_he_list", C.c_void_p),
    ("n_segments", C.c_int32), ("seg_he_start", C.c_void_p), ("seg_he_end", C.c_void_p),
    ("seg_n_edges", C.c_void_p), ("seg_on_boundary", C.c_void_p),
    ("n_non_blocking", C.c_int32), ("n_blocking", C.c_int32),
    ("non_blocking", C.c_void_p), ("blocking", C.c_void_p),
    ("int_length", C.c_double),
]
_DT = dict(vx=np.float64, vy=np.float64, he_src=np.int32, he_twin=np.int32, he_next=np.int32,
           he_face=np.int32, he_seg=np.int32, he_next_hf=np.int32, he_prev_hf=np.int32,
           he_dir=np.uint8, he_length=np.float64, face_inside=np.uint8,
           face_he_offset=np.int32, face_he_count=np.int32, face_he_list=np.int32,
           seg_he_start=np.int32, seg_he_end=np.int32, seg_n_edges=np.int32,
           seg_on_boundary=np.uint8, non_blocking=np.int32, blocking=np.int32)
_COUNT = dict(n_vertices="vx", n_halfedges="he_src", n_faces="face_inside", n_face_he="face_he_list",
              n_segments="seg_he_start", n_non_blocking="non_blocking", n_blocking="blocking")

FEATURE_IDS = {
    "is_point_inside_window": 0, "edge_length": 1, "seg_number": 2, "is_segment_internal": 3,
    "minus_is_segment_internal": 4, "segment_size_2": 5, "face_number": 6, "face_area_2": 7,
    "face_perimeter": 8, "face_shape": 9, "face_sum_of_angles": 10, "min_angle": 11,
    "edge_length_internal": 12,
}


class _SoA(C.Structure):
    _fields_ = _FIELDS


def lib():
    global _LIB
    if _LIB is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            subprocess.check_call(["make", "-C", _HERE])
        L = C.CDLL(p)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_split_stats.restype = C.c_int64
        L.oracle_merge_stats.restype = C.c_int64
        L.oracle_flip_stats.restype = C.c_int64
        L.oracle_sample_splits.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int64,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_split_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_merge_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_flip_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_pl.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int]
        L.oracle_pl_pass.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64,
                                     C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def hardware_threads():
    return int(lib().oracle_hardware_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Tess:
    """Holds the numpy arrays of a flattened tessellation alive and exposes the
    ``lite_tess_soa`` struct the C functions take."""

    def __init__(self, arrays):
        self.a = {k: np.ascontiguousarray(arrays[k], dtype=dt) for k, dt in _DT.items()}
        s = _SoA()
        for cnt, arr in _COUNT.items():
            setattr(s, cnt, len(self.a[arr]))
        for k in _DT:
            setattr(s, k, self.a[k].ctypes.data)
        s.int_length = float(arrays["int_length"])
        self.int_length = s.int_length
        self.s = s
        self.n_nb = len(self.a["non_blocking"])
        self.n_b = len(self.a["blocking"])

    def ref(self):
        return C.addressof(self.s)


def _ids(features):
    return np.array([FEATURE_IDS[f] if isinstance(f, str) else int(f) for f in features], dtype=np.int32)


def sample_splits(t, n_batch, seed, ctr0=0, j0=0, n=None, nthreads=1):
    n = n_batch if n is None else n
    he = np.zeros(n, dtype=np.int32)
    x = np.zeros(n)
    ang = np.zeros(n)
    if lib().oracle_sample_splits(t.ref(), n_batch, seed, ctr0, j0, n, _p(he), _p(x), _p(ang), nthreads):
        raise RuntimeError(lib().oracle_last_error().decode())
    return he, x, ang


def lines_clip(t, n_batch, seed, ctr0=0, j0=0, n=None, nthreads=1):
    """Sample + clip only (config C4); returns (checksum[2], lines without exit)."""
    n = n_batch if n is None else n
    cs = np.zeros(2, dtype=np.uint64)
    L = lib()
    L.oracle_lines_clip.restype = C.c_int64
    L.oracle_lines_clip.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_int]
    bad = L.oracle_lines_clip(t.ref(), n_batch, seed, ctr0, j0, n, _p(cs), nthreads)
    if bad < 0:
        raise RuntimeError(L.oracle_last_error().decode())
    return cs, int(bad)


def split_stats(t, features, he, x, angle, nthreads=1):
    fid = _ids(features)
    he = np.ascontiguousarray(he, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    angle = np.ascontiguousarray(angle, dtype=np.float64)
    n, d = len(he), len(fid)
    e2 = np.zeros(n, dtype=np.int32)
    pts = np.zeros((n, 4))
    st = np.zeros((n, d))
    bad = lib().oracle_split_stats(t.ref(), d, _p(fid), n, _p(he), _p(x), _p(angle), _p(e2), _p(pts),
                                   _p(st), nthreads)
    if bad < 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return dict(split_e2=e2, split_p1=pts[:, :2].copy(), split_p2=pts[:, 2:].copy(), split_stat=st, bad=bad)


def merge_stats(t, features, nthreads=1):
    fid = _ids(features)
    d = len(fid)
    he = np.zeros(t.n_nb, dtype=np.int32)
    st = np.zeros((t.n_nb, d))
    if lib().oracle_merge_stats(t.ref(), d, _p(fid), _p(he), _p(st), nthreads) < 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return dict(merge_he=he, merge_stat=st)


def flip_stats(t, features, nthreads=1):
    fid = _ids(features)
    d, nf = len(fid), 2 * t.n_b
    e1 = np.zeros(nf, dtype=np.int32)
    e2 = np.zeros(nf, dtype=np.int32)
    right = np.zeros(nf, dtype=np.int32)
    p2 = np.zeros((nf, 2))
    st = np.zeros((nf, d))
    bad = lib().oracle_flip_stats(t.ref(), d, _p(fid), _p(e1), _p(e2), _p(right), _p(p2), _p(st), nthreads)
    if bad < 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return dict(flip_e1=e1, flip_e2=e2, flip_right=right, flip_p2=p2, flip_stat=st, bad=bad)


def pl(theta, split_stat, flip_stat, sum_m, sum_f, int_length, n_global=None, nthreads=1):
    th = np.ascontiguousarray(theta, dtype=np.float64)
    d = len(th)
    ss = np.ascontiguousarray(split_stat, dtype=np.float64).reshape(-1, d)
    fs = np.ascontiguousarray(flip_stat, dtype=np.float64).reshape(-1, d)
    sm = np.ascontiguousarray(sum_m, dtype=np.float64)
    sf = np.ascontiguousarray(sum_f, dtype=np.float64)
    v = C.c_double()
    g = np.zeros(d)
    H = np.zeros((d, d))
    ng = len(ss) if n_global is None else n_global
    if lib().oracle_pl(d, _p(th), len(ss), _p(ss), len(fs), _p(fs), _p(sm), _p(sf), float(int_length), ng,
                       C.byref(v), _p(g), _p(H), nthreads):
        raise RuntimeError(lib().oracle_last_error().decode())
    return v.value, g, H


def pl_pass(t, features, theta, n_splits, seed, ctr0=0, nthreads=1):
    """One whole CPU pass: SetEnergy + AddSplits(n_splits) + value/gradient/Hessian."""
    fid = _ids(features)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    d = len(fid)
    v = C.c_double()
    g = np.zeros(d)
    H = np.zeros((d, d))
    if lib().oracle_pl_pass(t.ref(), d, _p(fid), _p(th), n_splits, seed, ctr0, C.byref(v), _p(g), _p(H), nthreads):
        raise RuntimeError(lib().oracle_last_error().decode())
    return v.value, g, H
