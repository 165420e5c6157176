"""ctypes binding of include/lite_b200.h.  No fallback: if ``libliteb200.so`` is
absent or a call fails, ``LiteError`` is raised."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FEATURE_IDS = {
    "is_point_inside_window": 0,
    "edge_length": 1,
    "seg_number": 2,
    "is_segment_internal": 3,
    "minus_is_segment_internal": 4,
    "segment_size_2": 5,
    "face_number": 6,
    "face_area_2": 7,
    "face_perimeter": 8,
    "face_shape": 9,
    "face_sum_of_angles": 10,
    "min_angle": 11,
    "edge_length_internal": 12,  # extension, not in the reference
}

# every symbol declared in include/lite_b200.h
SYMBOLS = [
    "lite_last_error", "lite_nccl_unique_id", "lite_ctx_create", "lite_ctx_destroy",
    "lite_ctx_sync", "lite_ctx_set_option", "lite_tess_upload", "lite_energy_set",
    "lite_candidates_merges_flips", "lite_candidates_add_splits_given",
    "lite_candidates_add_splits_philox", "lite_candidates_clear_splits",
    "lite_candidates_count", "lite_pl_eval", "lite_pl_sums", "lite_lines_sample_clip",
    "lite_debug_fetch", "lite_last_kernel_ms", "lite_kernel_launch_count",
    "lite_measure_fp64_peak", "lite_last_reduce_ms", "lite_timer_mark", "lite_timer_elapsed",
    "lite_smf_create", "lite_smf_destroy", "lite_smf_set_model", "lite_smf_run", "lite_smf_stats",
    "lite_smf_trace", "lite_smf_export", "lite_smf_count", "lite_smf_last_ms",
    "lite_pl_combine", "lite_shard_range", "lite_pl_failures", "lite_comm_info", "lite_smf_intersections", "lite_candidates_reserve", "lite_timers_accumulated",
]

# columns of lite_smf_run / lite_smf_stats (applications/checkACS.cpp:136-148 + energy)
SMF_COLS = ["l", "n", "n_bl", "n_nbl", "m", "prop_S", "acc_S", "prop_M", "acc_M", "prop_F", "acc_F", "energy"]

FETCH = dict(SPLIT_E1=0, SPLIT_E2=1, SPLIT_P1=2, SPLIT_P2=3, SPLIT_X=4, SPLIT_ANGLE=5,
             SPLIT_STAT=6, MERGE_HE=7, MERGE_STAT=8, FLIP_E1=9, FLIP_E2=10, FLIP_P2=11,
             FLIP_RIGHT=12, FLIP_STAT=13, SPLIT_U=14, STATUS=15)


class LiteError(RuntimeError):
    pass


class _SoA(C.Structure):
    _fields_ = [
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


class TessSoA:
    """Host-side flattened tessellation (dict of numpy arrays + int_length) in the
    layout of ``lite_tess_soa``."""

    def __init__(self, arrays):
        self.a = {}
        for k, dt in _DT.items():
            self.a[k] = np.ascontiguousarray(arrays[k], dtype=dt)
        self.int_length = float(arrays["int_length"])

    @classmethod
    def load(cls, path):
        z = np.load(path)
        return cls({k: z[k] for k in z.files})

    def save(self, path):
        np.savez_compressed(path, int_length=np.float64(self.int_length), **self.a)

    def n_cells(self):
        return int(self.a["face_inside"].sum())

    def as_struct(self):
        """The lite_tess_soa view of the arrays (built once: the arrays are not to be modified
        afterwards)."""
        if getattr(self, "_struct", None) is not None:
            return self._struct
        s = _SoA()
        a = self.a
        s.n_vertices = len(a["vx"])
        s.n_halfedges = len(a["he_src"])
        s.n_faces = len(a["face_inside"])
        s.n_face_he = len(a["face_he_list"])
        s.n_segments = len(a["seg_he_start"])
        s.n_non_blocking = len(a["non_blocking"])
        s.n_blocking = len(a["blocking"])
        s.int_length = self.int_length
        for k in _DT:
            setattr(s, k, a[k].ctypes.data_as(C.c_void_p))
        self._struct = s
        return s


def lib_path():
    # LITE_B200_LIB selects another build of the same library (kernel tuning experiments)
    return os.environ.get("LITE_B200_LIB") or os.path.join(_HERE, "libliteb200.so")


_LIB = None


def load_library():
    """dlopen the in-tree CUDA library; raises LiteError when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise LiteError("%s not found: build it with `make` (python -c 'import __graft_entry__ as g; g.build()'). "
                        "There is no CPU fallback." % p)
    L = C.CDLL(p)
    L.lite_last_error.restype = C.c_char_p
    L.lite_kernel_launch_count.restype = C.c_int64
    L.lite_kernel_launch_count.argtypes = [C.c_void_p]
    L.lite_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.lite_ctx_destroy.argtypes = [C.c_void_p]
    L.lite_ctx_sync.argtypes = [C.c_void_p]
    L.lite_ctx_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    L.lite_tess_upload.argtypes = [C.c_void_p, C.POINTER(_SoA)]
    L.lite_energy_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.lite_candidates_merges_flips.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.lite_candidates_add_splits_given.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_int64]
    L.lite_candidates_add_splits_philox.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64]
    L.lite_candidates_clear_splits.argtypes = [C.c_void_p]
    L.lite_candidates_reserve.argtypes = [C.c_void_p, C.c_int64]
    L.lite_candidates_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.lite_pl_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lite_pl_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.lite_lines_sample_clip.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    L.lite_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    L.lite_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.lite_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.lite_nccl_unique_id.argtypes = [C.c_void_p]
    L.lite_timers_accumulated.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.lite_last_reduce_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.lite_timer_mark.argtypes = [C.c_void_p, C.c_int]
    L.lite_timer_elapsed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.lite_smf_create.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_uint64]
    L.lite_smf_destroy.argtypes = [C.c_void_p]
    L.lite_smf_set_model.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
    L.lite_smf_run.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
    L.lite_smf_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.lite_smf_trace.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]
    L.lite_smf_export.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
    L.lite_smf_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64)]
    L.lite_smf_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.lite_smf_intersections.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_int64)]
    L.lite_pl_combine.argtypes = [C.c_int, C.c_int64, C.c_double] + [C.c_void_p] * 8
    L.lite_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.lite_pl_failures.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.lite_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    _LIB = L
    return L


def _ck(L, rc):
    if rc != 0:
        raise LiteError("lite_b200 error %d: %s" % (rc, L.lite_last_error().decode()))


def nccl_unique_id():
    L = load_library()
    buf = (C.c_char * 128)()
    _ck(L, L.lite_nccl_unique_id(buf))
    return bytes(buf)


def shard_range(n, rank, world):
    """Contiguous shard [r*n/G, (r+1)*n/G) of rank r: lite_shard_range, the partition
    lite_candidates_add_splits_philox uses.  Returns (first, count)."""
    L = load_library()
    a, b = C.c_int64(), C.c_int64()
    _ck(L, L.lite_shard_range(int(n), int(rank), int(world), C.byref(a), C.byref(b)))
    return a.value, b.value


def pl_combine(theta, n_splits_global, int_length, S, F, M, Fl):
    """lite_pl_combine: log-PL, gradient, Hessian from reduced sums (host arithmetic of
    lite_pl_eval; S, F packed [S0, S1_k, S2_kl (k<=l)])."""
    L = load_library()
    th = np.ascontiguousarray(theta, dtype=np.float64)
    d = len(th)
    arr = [np.ascontiguousarray(a, dtype=np.float64) for a in (S, F, M, Fl)]
    nt = 1 + d + d * (d + 1) // 2
    assert len(arr[0]) >= nt and len(arr[1]) >= nt and len(arr[2]) >= d and len(arr[3]) >= d
    lpl = C.c_double()
    g = np.zeros(d)
    H = np.zeros((d, d))
    _ck(L, L.lite_pl_combine(d, int(n_splits_global), float(int_length), th.ctypes.data_as(C.c_void_p),
                             *[a.ctypes.data_as(C.c_void_p) for a in arr], C.cast(C.byref(lpl), C.c_void_p),
                             g.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p)))
    return lpl.value, g, H


class Context:
    """One GPU context (``lite_ctx``).  Method names follow the C ABI."""

    def __init__(self, device=0, rank=0, world_size=1, nccl_id=None):
        self.L = load_library()
        self.h = C.c_void_p()
        idbuf = None
        if nccl_id is not None:
            idbuf = C.create_string_buffer(bytes(nccl_id), 128)
        _ck(self.L, self.L.lite_ctx_create(C.byref(self.h), device, rank, world_size, idbuf))
        self.d = 0
        self.rank, self.world = rank, world_size
        self.device = device

    def close(self):
        if self.h:
            self.L.lite_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _ck(self.L, self.L.lite_ctx_sync(self.h))

    def record_splits(self, on=True):
        _ck(self.L, self.L.lite_ctx_set_option(self.h, 1, 1 if on else 0))

    def validate_on_host(self, on=True):
        """LITE_OPT_VALIDATE_HOST: run the export assertions inside tess_upload (default: on the
        device behind the copy, reported by the next synchronising call)."""
        _ck(self.L, self.L.lite_ctx_set_option(self.h, 3, 1 if on else 0))

    def sampler(self, iid):
        """LITE_OPT_SAMPLER: i.i.d. sorted positions (the reference's split_sample law) or the
        default stratified sampler."""
        _ck(self.L, self.L.lite_ctx_set_option(self.h, 2, 1 if iid else 0))

    def tess_upload(self, soa: TessSoA):
        s = soa.as_struct()
        _ck(self.L, self.L.lite_tess_upload(self.h, C.byref(s)))
        self.n_nb = len(soa.a["non_blocking"])
        self.n_b = len(soa.a["blocking"])

    def energy_set(self, features):
        ids = np.array([FEATURE_IDS[f] if isinstance(f, str) else int(f) for f in features], dtype=np.int32)
        _ck(self.L, self.L.lite_energy_set(self.h, len(ids), ids.ctypes.data_as(C.c_void_p)))
        self.d = len(ids)

    def merges_flips(self):
        nm, nf = C.c_int32(), C.c_int32()
        _ck(self.L, self.L.lite_candidates_merges_flips(self.h, C.byref(nm), C.byref(nf)))
        return nm.value, nf.value

    def add_splits_given(self, he, x, angle, n_global_total=None):
        he = np.ascontiguousarray(he, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        angle = np.ascontiguousarray(angle, dtype=np.float64)
        n = len(he)
        if n_global_total is None:
            n_global_total = n
        _ck(self.L, self.L.lite_candidates_add_splits_given(
            self.h, n, he.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
            angle.ctypes.data_as(C.c_void_p), n_global_total))

    def add_splits_philox(self, n_global, seed, offset=0):
        _ck(self.L, self.L.lite_candidates_add_splits_philox(self.h, n_global, seed, offset))

    def reserve_splits(self, n_more_local):
        _ck(self.L, self.L.lite_candidates_reserve(self.h, int(n_more_local)))

    def clear_splits(self):
        _ck(self.L, self.L.lite_candidates_clear_splits(self.h))

    def count(self):
        a, b = C.c_int64(), C.c_int64()
        _ck(self.L, self.L.lite_candidates_count(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def pl_eval(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        d = self.d
        assert len(th) == d
        lpl = C.c_double()
        g = np.zeros(d)
        H = np.zeros((d, d))
        _ck(self.L, self.L.lite_pl_eval(self.h, th.ctypes.data_as(C.c_void_p), C.byref(lpl),
                                        g.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p)))
        return lpl.value, g, H

    def pl_failures(self):
        """(split, flip) candidates behind the last pl_eval whose ray found no exit edge."""
        a, b = C.c_int64(), C.c_int64()
        _ck(self.L, self.L.lite_pl_failures(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def comm_info(self):
        """(mode, exchange_ms): 0 single rank, 1 ncclAllReduce, 2 peer memory in k_reduce."""
        m, t = C.c_int32(), C.c_float()
        _ck(self.L, self.L.lite_comm_info(self.h, C.byref(m), C.byref(t)))
        return m.value, t.value

    def pl_sums(self):
        m = np.zeros(self.d)
        f = np.zeros(self.d)
        _ck(self.L, self.L.lite_pl_sums(self.h, m.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p)))
        return m, f

    def lines_sample_clip(self, n_global, seed, offset=0, mode=0):
        cs = np.zeros(2, dtype=np.uint64)
        _ck(self.L, self.L.lite_lines_sample_clip(self.h, n_global, seed, offset, mode,
                                                  cs.ctypes.data_as(C.c_void_p)))
        return cs

    def fetch(self, which, first=0, count=None):
        w = FETCH[which] if isinstance(which, str) else which
        name = which if isinstance(which, str) else {v: k for k, v in FETCH.items()}[which]
        if name == "STATUS":
            out = np.zeros(4, dtype=np.int64)
            _ck(self.L, self.L.lite_debug_fetch(self.h, w, 0, 4, out.ctypes.data_as(C.c_void_p)))
            return out
        if count is None:
            if name.startswith("SPLIT"):
                count = self.count()[0] - first
            elif name.startswith("MERGE"):
                count = self.n_nb - first
            else:
                count = 2 * self.n_b - first
        if name.endswith("_STAT"):
            out = np.zeros((count, self.d))
        elif name in ("SPLIT_P1", "SPLIT_P2", "FLIP_P2"):
            out = np.zeros((count, 2))
        elif name in ("SPLIT_X", "SPLIT_ANGLE", "SPLIT_U"):
            out = np.zeros(count)
        else:
            out = np.zeros(count, dtype=np.int32)
        if count:
            _ck(self.L, self.L.lite_debug_fetch(self.h, w, first, count, out.ctypes.data_as(C.c_void_p)))
        return out

    def last_kernel_ms(self):
        a, b, m = C.c_float(), C.c_float(), C.c_float()
        _ck(self.L, self.L.lite_last_kernel_ms(self.h, C.byref(a), C.byref(b), C.byref(m)))
        return a.value, b.value, m.value

    def timers_accumulated(self, reset=False):
        """(k_split ms, k_reduce ms, merges+flips ms, exchange ms, evaluations) summed since the last reset."""
        out = np.zeros(5)
        _ck(self.L, self.L.lite_timers_accumulated(self.h, out.ctypes.data_as(C.c_void_p), 1 if reset else 0))
        return out

    def last_reduce_ms(self):
        t = C.c_float()
        _ck(self.L, self.L.lite_last_reduce_ms(self.h, C.byref(t)))
        return t.value

    def timer_mark(self, slot):
        _ck(self.L, self.L.lite_timer_mark(self.h, slot))

    def timer_elapsed(self, a, b):
        t = C.c_float()
        _ck(self.L, self.L.lite_timer_elapsed(self.h, a, b, C.byref(t)))
        return t.value

    def launch_count(self):
        return int(self.L.lite_kernel_launch_count(self.h))

    def measure_fp64_peak(self):
        t = C.c_double()
        _ck(self.L, self.L.lite_measure_fp64_peak(self.h, C.byref(t)))
        return t.value

    # ---- ensemble of SMF chains (BASELINE.json config 5) ----
    def smf_create(self, n_chains_global, max_segments, rect=(0.0, 0.0, 1.0, 1.0), seed=5):
        _ck(self.L, self.L.lite_smf_create(self.h, int(n_chains_global), int(max_segments),
                                           *map(float, rect), int(seed)))

    def smf_destroy(self):
        _ck(self.L, self.L.lite_smf_destroy(self.h))

    def smf_set_model(self, features, theta, p_split=0.33, p_merge=0.33):
        ids = np.array([FEATURE_IDS[f] if isinstance(f, str) else int(f) for f in features], dtype=np.int32)
        th = np.ascontiguousarray(theta, dtype=np.float64)
        assert len(th) == len(ids)
        _ck(self.L, self.L.lite_smf_set_model(self.h, len(ids), ids.ctypes.data_as(C.c_void_p),
                                              th.ctypes.data_as(C.c_void_p), float(p_split), float(p_merge)))

    def smf_count(self):
        a, b, f, sz = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _ck(self.L, self.L.lite_smf_count(self.h, C.byref(a), C.byref(b), C.byref(f), C.byref(sz)))
        return a.value, b.value, f.value, sz.value

    def smf_run(self, n_samples, steps_per_sample, rows=True):
        """Returns an array [n_samples, 12, n_local] (columns SMF_COLS) or None."""
        out = None
        if rows:
            out = np.zeros((int(n_samples), len(SMF_COLS), self.smf_count()[0]))
        _ck(self.L, self.L.lite_smf_run(self.h, int(n_samples), int(steps_per_sample),
                                        out.ctypes.data_as(C.c_void_p) if rows else None))
        return out

    def smf_stats(self, validate=False):
        out = np.zeros((len(SMF_COLS), self.smf_count()[0]))
        _ck(self.L, self.L.lite_smf_stats(self.h, out.ctypes.data_as(C.c_void_p), 1 if validate else 0))
        return out

    def smf_trace(self, chain_local, n_steps):
        """Advance one chain by n_steps proposals; returns the raw trace (23 doubles per step)."""
        out = np.zeros((int(n_steps), 23))
        done = C.c_int64()
        _ck(self.L, self.L.lite_smf_trace(self.h, int(chain_local), int(n_steps), out.ctypes.data_as(C.c_void_p),
                                          C.byref(done)))
        return out[:done.value]

    def smf_export(self, chain_local):
        """One chain's tessellation as a host model (lite_b200.HostTess)."""
        from .host import HostTess
        h = C.c_void_p()
        _ck(self.L, self.L.lite_smf_export(self.h, int(chain_local), C.byref(h)))
        return HostTess(handle=h)

    def smf_intersections(self, probes):
        """checkACS's intersection output for the current state of every local chain: arrays
        (chain, probe, xy) sorted by (chain, probe, x, y).  probes: [n, 4] segments x0 y0 x1 y1."""
        pr = np.ascontiguousarray(probes, dtype=np.float64).reshape(-1, 4)
        n = C.c_int64()
        _ck(self.L, self.L.lite_smf_intersections(self.h, len(pr), pr.ctypes.data_as(C.c_void_p), 0, None, None, None, C.byref(n)))
        cap = max(1, n.value)
        ch, kk, xy = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32), np.zeros((cap, 2))
        _ck(self.L, self.L.lite_smf_intersections(self.h, len(pr), pr.ctypes.data_as(C.c_void_p), cap, ch.ctypes.data_as(C.c_void_p),
                                                  kk.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p), C.byref(n)))
        m = n.value
        order = np.lexsort((xy[:m, 1], xy[:m, 0], kk[:m], ch[:m]))
        return ch[:m][order], kk[:m][order], xy[:m][order]

    def smf_last_ms(self):
        t = C.c_float()
        _ck(self.L, self.L.lite_smf_last_ms(self.h, C.byref(t)))
        return t.value
