"""ctypes binding of include/lite_b200_host.h (host model: build, CRTT input
generator, moves, export).  Thin wrapper; the implementation is C++."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import LiteError, TessSoA, _SoA, _DT, _ck, load_library

HOST_SYMBOLS = [
    "lite_host_create_rect", "lite_host_create_from_text", "lite_host_destroy",
    "lite_host_crtt_run", "lite_host_update_split", "lite_host_update_merge",
    "lite_host_update_flip", "lite_host_counts", "lite_host_check", "lite_host_export",
    "lite_host_write_text", "lite_host_clip_segment", "lite_host_insert_segment",
    "lite_host_is_t_tessellation", "lite_host_edges", "lite_host_remove_ivertices",
    "lite_host_remove_lvertices", "lite_host_remove_xvertices", "lite_host_vertex_census",
    "lite_host_non_t_vertices",
]

_COUNT_OF = dict(vx="n_vertices", vy="n_vertices", he_src="n_halfedges", he_twin="n_halfedges",
                 he_next="n_halfedges", he_face="n_halfedges", he_seg="n_halfedges",
                 he_next_hf="n_halfedges", he_prev_hf="n_halfedges", he_dir="n_halfedges",
                 he_length="n_halfedges", face_inside="n_faces", face_he_offset="n_faces",
                 face_he_count="n_faces", face_he_list="n_face_he", seg_he_start="n_segments",
                 seg_he_end="n_segments", seg_n_edges="n_segments", seg_on_boundary="n_segments",
                 non_blocking="n_non_blocking", blocking="n_blocking")


def _lib():
    L = load_library()
    if not getattr(L, "_host_ready", False):
        L.lite_host_create_rect.argtypes = [C.POINTER(C.c_void_p), C.c_double, C.c_double, C.c_double, C.c_double]
        L.lite_host_create_from_text.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
        L.lite_host_destroy.argtypes = [C.c_void_p]
        L.lite_host_crtt_run.argtypes = [C.c_void_p, C.c_double, C.c_uint64, C.c_uint64, C.c_void_p]
        L.lite_host_update_split.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double]
        L.lite_host_update_merge.argtypes = [C.c_void_p, C.c_int32]
        L.lite_host_update_flip.argtypes = [C.c_void_p, C.c_int32, C.c_int]
        L.lite_host_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.lite_host_check.argtypes = [C.c_void_p]
        L.lite_host_export.argtypes = [C.c_void_p, C.POINTER(_SoA)]
        L.lite_host_write_text.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.lite_host_write_text.restype = C.c_int64
        L.lite_host_clip_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lite_host_insert_segment.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        L.lite_host_is_t_tessellation.argtypes = [C.c_void_p]
        L.lite_host_edges.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lite_host_edges.restype = C.c_int64
        L.lite_host_remove_ivertices.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.lite_host_remove_lvertices.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
        L.lite_host_remove_xvertices.argtypes = [C.c_void_p, C.c_double]
        L.lite_host_vertex_census.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 4
        L.lite_host_non_t_vertices.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
        L._host_ready = True
    return L


class HostTess:
    """A dynamic T-tessellation on the host (C++ ``lite::HostTess``)."""

    def __init__(self, rect=(0.0, 0.0, 1.0, 1.0), text=None, handle=None):
        self.L = _lib()
        self.h = C.c_void_p()
        if handle is not None:
            self.h = handle  # takes ownership (lite_smf_export)
        elif text is not None:
            _ck(self.L, self.L.lite_host_create_from_text(C.byref(self.h), text.encode()))
        else:
            _ck(self.L, self.L.lite_host_create_rect(C.byref(self.h), *map(float, rect)))

    def close(self):
        if self.h:
            self.L.lite_host_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def crtt_run(self, tau, n_steps, seed):
        c = np.zeros(6, dtype=np.int64)
        _ck(self.L, self.L.lite_host_crtt_run(self.h, float(tau), int(n_steps), int(seed), c.ctypes.data_as(C.c_void_p)))
        return c

    def update_split(self, he, x, angle):
        _ck(self.L, self.L.lite_host_update_split(self.h, int(he), float(x), float(angle)))

    def update_merge(self, nb_index):
        _ck(self.L, self.L.lite_host_update_merge(self.h, int(nb_index)))

    def update_flip(self, b_index, side):
        _ck(self.L, self.L.lite_host_update_flip(self.h, int(b_index), int(side)))

    def counts(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _ck(self.L, self.L.lite_host_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def check(self):
        _ck(self.L, self.L.lite_host_check(self.h))

    def export(self) -> TessSoA:
        s = _SoA()
        _ck(self.L, self.L.lite_host_export(self.h, C.byref(s)))
        arrays = {}
        for k, dt in _DT.items():
            n = getattr(s, _COUNT_OF[k])
            ptr = getattr(s, k)
            if n == 0 or not ptr:
                arrays[k] = np.zeros(0, dtype=dt)
            else:
                buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
                arrays[k] = np.frombuffer(buf, dtype=dt, count=n).copy()
        arrays["int_length"] = s.int_length
        return TessSoA(arrays)

    def clip_segment(self, seg):
        """clip_segment_by_polygon (lib/ttessel.cpp:3762-3786) by this tessellation's window."""
        a = np.asarray(seg, dtype=np.float64).reshape(4).copy()
        out = np.zeros(4)
        _ck(self.L, self.L.lite_host_clip_segment(self.h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    def insert_segment(self, seg):
        """LineTes::insert_segment (lib/ttessel.cpp:904-980); returns the new segment's id."""
        a = np.asarray(seg, dtype=np.float64).reshape(4).copy()
        sid = C.c_int32(-1)
        _ck(self.L, self.L.lite_host_insert_segment(self.h, a.ctypes.data_as(C.c_void_p), C.byref(sid)))
        return sid.value

    def is_t_tessellation(self):
        return self.L.lite_host_is_t_tessellation(self.h) == 1

    def edges(self):
        """Every halfedge: (end points [n,4], segment id, next_hf, prev_hf), T-tessellation or not."""
        n = self.L.lite_host_edges(self.h, 0, None, None, None, None)
        xy = np.zeros((n, 4)); seg = np.zeros(n, np.int32); nx = np.zeros(n, np.int32); pv = np.zeros(n, np.int32)
        self.L.lite_host_edges(self.h, n, xy.ctypes.data_as(C.c_void_p), seg.ctypes.data_as(C.c_void_p),
                               nx.ctypes.data_as(C.c_void_p), pv.ctypes.data_as(C.c_void_p))
        return xy, seg, nx, pv

    def remove_ivertices(self, imax=0):
        """LineTes::remove_ivertices (lib/ttessel.cpp:982-1086); returns (removed, added) length."""
        a, b = C.c_double(), C.c_double()
        _ck(self.L, self.L.lite_host_remove_ivertices(self.h, int(imax), C.byref(a), C.byref(b)))
        return a.value, b.value

    def remove_lvertices(self, imax=0):
        """LineTes::remove_lvertices (lib/ttessel.cpp:1088-1179); returns the added length."""
        a = C.c_double()
        _ck(self.L, self.L.lite_host_remove_lvertices(self.h, int(imax), C.byref(a)))
        return a.value

    def remove_xvertices(self, step):
        """LineTes::remove_xvertices (lib/ttessel.cpp:1181-1296)."""
        _ck(self.L, self.L.lite_host_remove_xvertices(self.h, float(step)))

    def vertex_census(self):
        """(free ends, corners inside the window, crossings, crossings remove_xvertices would break)"""
        v = [C.c_int32() for _ in range(4)]
        _ck(self.L, self.L.lite_host_vertex_census(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def non_t_vertices(self):
        """(count, position of the first) of the vertices that are not T (nor plain boundary) vertices."""
        n = C.c_int32()
        xy = np.zeros(2)
        _ck(self.L, self.L.lite_host_non_t_vertices(self.h, C.byref(n), xy.ctypes.data_as(C.c_void_p)))
        return n.value, (float(xy[0]), float(xy[1]))

    def write_text(self):
        n = self.L.lite_host_write_text(self.h, None, 0)
        buf = C.create_string_buffer(n)
        self.L.lite_host_write_text(self.h, buf, n)
        return buf.value.decode()


def generate_crtt(n_cells_target, seed=5, rect=(0.0, 0.0, 1.0, 1.0), sweeps=60):
    """CRTT tessellation with about ``n_cells_target`` cells in the unit square:
    tau is chosen from the calibration cells ~= 14.5*tau^2 (measured with the
    exact oracle chain) and the chain is run for ``sweeps`` proposals per cell."""
    tau = (n_cells_target / 14.5) ** 0.5
    t = HostTess(rect)
    t.crtt_run(tau, int(sweeps * n_cells_target), seed)
    return t, tau
