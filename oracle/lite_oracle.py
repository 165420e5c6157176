"""Exact-arithmetic CPU restatement of LiTe's pseudo-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lite_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline leg
use it, and only as the checker.

The reference (kien-kieu/lite) computes every predicate and construction with
``CGAL::Lazy_exact_nt<Gmpq>`` (reference ``include/ttessel.h:53-56``).  CGAL is
not available here, so this file restates the algorithm with Python
``fractions.Fraction`` (exact where the reference is exact, IEEE doubles through
``math`` exactly where the reference drops to ``double``).  Every function
cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).

Third-party arithmetic that is *not* in the reference tree and is restated
here from its published behaviour: CGAL ``Arrangement_2`` (DCEL surgery),
``Polygon_2::area`` / ``bounded_side_2`` / ``collinear`` / ``intersection``,
``Linear_algebraCd::inverse``; version unpinned by the reference
(``CMakeLists.txt:12``).  Halfedge container order, the ``outer_ccb()``
representative of a face and ``CGAL::Random`` streams are CGAL internals that
no reference test observes: *parity unpinned* for those (they are carried
explicitly through the SoA export instead).

Pinned by: the reference's split/flip constructor known answers
(``tests/check_lite.cpp:671-873``), its face-prediction property
(``tests/check_lite.cpp:339-397``, ``:875-1049``: the scripted flip, the square and the
non-convex holed domain, i.e. all four cases of ``hpolygon_insert_edge`` /
``hpolygon_remove_edge``), the 11-cell fixture
(``wrap/R/RLiTe/inst/tessellation_data.txt``) and the closed-form CRTT
pseudo-likelihood (``demos/pseudo_lik_inference.cpp:80-81``); see
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

CGAL_PI = 3.14159265358979323846  # same literal as CGAL's CGAL_PI

Fr = Fraction


# ----------------------------------------------------------------------------
# exact 2-D helpers (stand-ins for the CGAL kernel predicates the path uses)
# ----------------------------------------------------------------------------
def _pt(x, y):
    return (Fr(x), Fr(y))


def _cross(ax, ay, bx, by):
    return ax * by - ay * bx


def orientation(p, q, r):
    """CGAL::orientation(p,q,r): sign of (q-p) x (r-p)."""
    c = _cross(q[0] - p[0], q[1] - p[1], r[0] - p[0], r[1] - p[1])
    return (c > 0) - (c < 0)


def collinear(p, q, r):
    return orientation(p, q, r) == 0


def sqdist(p, q):
    dx = q[0] - p[0]
    dy = q[1] - p[1]
    return dx * dx + dy * dy


def poly_area(poly):
    """CGAL Polygon_2::area() (signed, exact)."""
    n = len(poly)
    s = Fr(0)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        s += x0 * y1 - x1 * y0
    return s / 2


def _on_segment(p, a, b):
    """Segment_2::has_on(p) (closed segment)."""
    if not collinear(a, b, p):
        return False
    return (min(a[0], b[0]) <= p[0] <= max(a[0], b[0])
            and min(a[1], b[1]) <= p[1] <= max(a[1], b[1]))


ON_UNBOUNDED_SIDE, ON_BOUNDARY, ON_BOUNDED_SIDE = -1, 0, 1


def bounded_side_2(poly, p):
    """CGAL::bounded_side_2 for a simple polygon (exact crossing number)."""
    n = len(poly)
    inside = False
    for i in range(n):
        a = poly[i]
        b = poly[(i + 1) % n]
        if _on_segment(p, a, b):
            return ON_BOUNDARY
        if (a[1] > p[1]) != (b[1] > p[1]):
            # x coordinate of the edge at height p.y
            xi = a[0] + (p[1] - a[1]) * (b[0] - a[0]) / (b[1] - a[1])
            if xi > p[0]:
                inside = not inside
    return ON_BOUNDED_SIDE if inside else ON_UNBOUNDED_SIDE


def _dir_rank(nx, ny, dx, dy):
    """Rank of direction d by its counter-clockwise angle from direction n,
    in (0, 2pi]; returns (half, ) key usable with _ccw_before."""
    c = _cross(nx, ny, dx, dy)
    d = nx * dx + ny * dy
    if c > 0:
        return 1  # (0, pi)
    if c == 0 and d < 0:
        return 2  # pi
    if c < 0:
        return 3  # (pi, 2pi)
    return 4  # same direction: 2pi


def _ccw_before(n, a, b):
    """True if direction a is met strictly before b when turning CCW from n."""
    ra = _dir_rank(n[0], n[1], a[0], a[1])
    rb = _dir_rank(n[0], n[1], b[0], b[1])
    if ra != rb:
        return ra < rb
    if ra in (2, 4):
        return False
    return _cross(a[0], a[1], b[0], b[1]) > 0


# ----------------------------------------------------------------------------
# DCEL (stand-in for CGAL::Arrangement_2 with LineTes_halfedge payload,
# reference include/ttessel.h:125-130, :402-451)
# ----------------------------------------------------------------------------
class Vertex:
    __slots__ = ("p", "inc")

    def __init__(self, p):
        self.p = p
        self.inc = None  # one incoming halfedge (target == self)

    def incoming(self):
        """Halfedges whose target is this vertex, clockwise, starting at inc
        (CGAL Halfedge_around_vertex_circulator order)."""
        res = []
        h = self.inc
        if h is None:
            return res
        while True:
            res.append(h)
            h = h.next.twin
            if h is self.inc:
                break
        return res

    def degree(self):
        return len(self.incoming())


class Halfedge:
    __slots__ = ("src", "twin", "next", "prev", "face",
                 "length", "next_hf", "prev_hf", "dir", "seg")

    def __init__(self, src):
        self.src = src
        self.twin = None
        self.next = None
        self.prev = None
        self.face = None
        self.length = 0.0
        self.next_hf = None
        self.prev_hf = None
        self.dir = True
        self.seg = None

    @property
    def tgt(self):
        return self.twin.src

    def set_length(self):
        # reference lib/ttessel.cpp:1459-1462
        self.length = math.sqrt(float(sqdist(self.src.p, self.tgt.p)))


class Face:
    __slots__ = ("outer", "holes", "data")

    def __init__(self):
        self.outer = None  # representative halfedge of the outer ccb
        self.holes = []    # representative halfedges of inner ccbs
        self.data = False  # True = inside the window (ttessel.h:125-127)


def ccb(h):
    """All halfedges of the connected boundary component through h."""
    res = [h]
    e = h.next
    while e is not h:
        res.append(e)
        e = e.next
    return res


def ccb_points(h):
    return [e.tgt.p for e in ccb(h)]


class Seg:
    """reference include/ttessel.h:182-212, lib/ttessel.cpp:788-873."""
    __slots__ = ("e",)

    def __init__(self, e=None):
        self.e = e

    def set_halfedge_handle(self, hh):
        # lib/ttessel.cpp:841-851
        if hh is None:
            self.e = None
        elif hh.dir:
            self.e = hh
        else:
            self.e = hh.twin

    def get_halfedge_handle(self):
        return self.e

    def halfedges_start(self):
        e = self.e
        while e.prev_hf is not None:
            e = e.prev_hf
        return e

    def halfedges_end(self):
        e = self.e
        while e.next_hf is not None:
            e = e.next_hf
        return e

    def number_of_edges(self):
        e = self.halfedges_start()
        i = 1
        while e.next_hf is not None:
            e = e.next_hf
            i += 1
        return i

    def number_of_edges_is_greater_than(self, n):
        e = self.halfedges_start()
        i = 1
        while e.next_hf is not None:
            i += 1
            e = e.next_hf
            if i > n:
                return True
        return False

    def pointSource(self):
        return self.halfedges_start().src.p

    def pointTarget(self):
        return self.halfedges_end().tgt.p

    def list_of_points(self):
        # lib/ttessel.cpp:864-873
        e = self.halfedges_start()
        lp = [e.src.p]
        while e is not None:
            lp.append(e.tgt.p)
            e = e.next_hf
        return lp


def set_junction(e1, e2):
    """lib/ttessel.cpp:3909-3930."""
    if e1 is not None:
        e1.next_hf = e2
        e1.twin.prev_hf = e2.twin if e2 is not None else None
    if e2 is not None:
        e2.prev_hf = e1
        e2.twin.next_hf = e1.twin if e1 is not None else None


def _obtuse(p, q, r):
    """CGAL::angle(p,q,r)==OBTUSE: (p-q).(r-q) < 0."""
    return ((p[0] - q[0]) * (r[0] - q[0]) + (p[1] - q[1]) * (r[1] - q[1])) < 0


def compute_prev_hf(e):
    """lib/ttessel.cpp:3853-3874."""
    for hf in e.src.incoming():
        if hf is not e.twin:
            if (collinear(e.tgt.p, e.src.p, hf.src.p)
                    and _obtuse(e.tgt.p, e.src.p, hf.src.p)):
                return hf
    return None


def compute_next_hf(e):
    """lib/ttessel.cpp:3882-3902."""
    for hf in e.tgt.incoming():
        if hf is not e:
            if (collinear(e.src.p, e.tgt.p, hf.src.p)
                    and _obtuse(e.src.p, e.tgt.p, hf.src.p)):
                return hf.twin
    return None


class LineTes:
    """Line tessellation: DCEL + segments (reference include/ttessel.h:171-377).

    Container order of ``halfedges`` mirrors CGAL's in-place list: new
    halfedges are appended (as a pair h, twin(h)); removed ones are unlinked.
    """

    def __init__(self):
        self.vertices = []
        self.halfedges = []
        self.faces = []
        self.unbounded = Face()
        self.faces.append(self.unbounded)
        self.window = []       # list of (outer, [holes]) polygons
        self.all_segments = []

    # ---- raw DCEL surgery (CGAL Arrangement_2 equivalents) ----------------
    def _new_vertex(self, p):
        v = Vertex(p)
        self.vertices.append(v)
        return v

    def _new_edge(self, v1, v2):
        h = Halfedge(v1)
        t = Halfedge(v2)
        h.twin = t
        t.twin = h
        self.halfedges.append(h)
        self.halfedges.append(t)
        return h

    def _prev_around(self, v, d):
        """Incoming halfedge at v after which a new outgoing edge with
        direction d must be linked (so that h_in.next becomes the new edge)."""
        best = None
        bestdir = None
        for hin in v.incoming():
            o = hin.twin  # outgoing
            od = (o.tgt.p[0] - v.p[0], o.tgt.p[1] - v.p[1])
            if best is None or _ccw_before(d, od, bestdir):
                best = hin
                bestdir = od
        return best

    def insert_polygon(self, pts, containing_face):
        """Insert a closed polygon disjoint from everything else inside
        ``containing_face``.  Returns the list of halfedges following the
        polygon orientation (one per polygon edge)."""
        n = len(pts)
        vs = [self._new_vertex(p) for p in pts]
        hs = [self._new_edge(vs[i], vs[(i + 1) % n]) for i in range(n)]
        for i in range(n):
            h = hs[i]
            h.next = hs[(i + 1) % n]
            h.prev = hs[(i - 1) % n]
            t = h.twin
            t.next = hs[(i - 1) % n].twin
            t.prev = hs[(i + 1) % n].twin
            vs[(i + 1) % n].inc = h
        nf = Face()
        self.faces.append(nf)
        area = poly_area(pts)
        if area > 0:
            inner, outerc = hs, [h.twin for h in hs]
        else:
            inner, outerc = [h.twin for h in hs], hs
        for h in inner:
            h.face = nf
        nf.outer = inner[0]
        for h in outerc:
            h.face = containing_face
        containing_face.holes.append(outerc[0])
        return hs

    def arr_split_edge(self, e, p):
        """Arrangement_2::split_edge: e keeps its source, its target becomes a
        new vertex at p; returns e (first half).  e.next is the second half."""
        v = self._new_vertex(p)
        t = e.twin
        old_tgt = t.src
        e2 = self._new_edge(v, old_tgt)
        t2 = e2.twin  # old_tgt -> v
        # links on e's side: e -> e2 -> old e.next
        e2.next = e.next
        e.next.prev = e2
        e.next = e2
        e2.prev = e
        # links on twin side: old t.prev -> t2 -> t
        t2.prev = t.prev
        t.prev.next = t2
        t2.next = t
        t.prev = t2
        t.src = v
        e2.face = e.face
        t2.face = t.face
        v.inc = e
        if old_tgt.inc is e:
            old_tgt.inc = e2
        return e

    def arr_merge_edge(self, e1, e2):
        """Arrangement_2::merge_edge(e1,e2): e1: a->v, e2: v->b with v of degree
        2; afterwards e1: a->b."""
        v = e1.tgt
        t1 = e1.twin
        t2 = e2.twin
        b = e2.tgt
        e1.next = e2.next
        e2.next.prev = e1
        t1.prev = t2.prev
        t2.prev.next = t1
        t1.src = b
        if b.inc is e2:
            b.inc = e1
        for f, old, new in ((e2.face, e2, e1), (t2.face, t2, t1)):
            if f.outer is old:
                f.outer = new
            for i, hrep in enumerate(f.holes):
                if hrep is old:
                    f.holes[i] = new
        self.halfedges.remove(e2)
        self.halfedges.remove(t2)
        self.vertices.remove(v)
        return e1

    def arr_insert_at_vertices(self, v1, v2):
        """Arrangement_2::insert_at_vertices: new edge v1->v2 inside one face.
        Returns the halfedge directed v1->v2."""
        d12 = (v2.p[0] - v1.p[0], v2.p[1] - v1.p[1])
        d21 = (-d12[0], -d12[1])
        prev1 = self._prev_around(v1, d12)
        prev2 = self._prev_around(v2, d21)
        f = prev1.face
        if prev2.face is not f:
            raise AssertionError("insert_at_vertices: vertices not on one face")
        same_ccb = any(e is prev2 for e in ccb(prev1))
        # which ccb(s) of f are involved
        def rep_of(h):
            cyc = set(map(id, ccb(h)))
            if f.outer is not None and id(f.outer) in cyc:
                return "outer", f.outer
            for hr in f.holes:
                if id(hr) in cyc:
                    return "hole", hr
            raise AssertionError("ccb representative not found")
        kind1, rep1 = rep_of(prev1)
        kind2, rep2 = (kind1, rep1) if same_ccb else rep_of(prev2)
        h = self._new_edge(v1, v2)
        t = h.twin
        a1 = prev1.next
        a2 = prev2.next
        prev1.next = h
        h.prev = prev1
        h.next = a2
        a2.prev = h
        prev2.next = t
        t.prev = prev2
        t.next = a1
        a1.prev = t
        h.face = f
        t.face = f
        if same_ccb:
            nf = Face()
            nf.data = f.data
            self.faces.append(nf)
            if kind1 == "outer":
                # cycle(h) becomes the new face, cycle(t) keeps f
                newc, oldrep = ccb(h), t
                f.outer = oldrep
            else:
                # a hole boundary is cut: the CCW cycle is the new face
                ch = ccb(h)
                if poly_area([e.tgt.p for e in ch]) > 0:
                    newc, oldrep = ch, t
                else:
                    newc, oldrep = ccb(t), h
                f.holes[[i for i, hr in enumerate(f.holes) if hr is rep1][0]] = oldrep
            for e in newc:
                e.face = nf
            nf.outer = newc[0]
            # dispatch the remaining holes of f
            poly = [e.tgt.p for e in newc]
            keep = []
            newset = set(map(id, newc))
            for hr in f.holes:
                if id(hr) in newset:
                    continue
                if hr is oldrep:
                    keep.append(hr)
                    continue
                if bounded_side_2(poly, hr.tgt.p) == ON_BOUNDED_SIDE:
                    nf.holes.append(hr)
                    for e in ccb(hr):
                        e.face = nf
                else:
                    keep.append(hr)
            f.holes = keep
        else:
            # two boundary components are joined: one fewer hole
            if kind1 == "outer":
                f.holes.remove(rep2)
            elif kind2 == "outer":
                f.holes.remove(rep1)
            else:
                f.holes.remove(rep2)
        return h

    def arr_remove_edge(self, h):
        """Arrangement_2::remove_edge(e,false,false): end vertices are kept.
        Returns the face that contains the region after removal."""
        t = h.twin
        f1, f2 = h.face, t.face
        p1, n1, p2, n2 = h.prev, h.next, t.prev, t.next
        if n1 is t or n2 is h:
            raise NotImplementedError("antenna removal is not on the path")

        def kind(f, e):
            cyc = set(map(id, ccb(e)))
            if f.outer is not None and id(f.outer) in cyc:
                return "outer", f.outer
            for hr in f.holes:
                if id(hr) in cyc:
                    return "hole", hr
            raise AssertionError
        k1, r1 = kind(f1, h)
        if f1 is f2:
            k2, r2 = k1, r1
        else:
            k2, r2 = kind(f2, t)
        # fix vertex incident pointers
        if h.tgt.inc is h:
            h.tgt.inc = p2
        if t.tgt.inc is t:
            t.tgt.inc = p1
        p1.next = n2
        n2.prev = p1
        p2.next = n1
        n1.prev = p2
        self.halfedges.remove(h)
        self.halfedges.remove(t)
        if f1 is not f2:
            if k1 == "outer" and k2 == "outer":
                keep, gone = f1, f2
                keep.outer = p1
            elif k1 == "outer":  # f1 lies in a hole of f2
                keep, gone = f2, f1
                keep.holes[[i for i, hr in enumerate(keep.holes) if hr is r2][0]] = p1
            else:
                keep, gone = f1, f2
                keep.holes[[i for i, hr in enumerate(keep.holes) if hr is r1][0]] = p1
            for hr in gone.holes:
                keep.holes.append(hr)
                for e in ccb(hr):
                    e.face = keep
            for e in ccb(p1):
                e.face = keep
            self.faces.remove(gone)
            return keep
        # same face on both sides: the boundary component splits in two
        c1 = ccb(p1)
        c2 = ccb(p2)
        f = f1
        if k1 == "outer":
            a1 = poly_area([e.tgt.p for e in c1])
            if a1 > 0:
                f.outer = p1
                f.holes.append(p2)
            else:
                f.outer = p2
                f.holes.append(p1)
        else:
            f.holes[[i for i, hr in enumerate(f.holes) if hr is r1][0]] = p1
            f.holes.append(p2)
        return f

    # ---- LineTes layer -----------------------------------------------------
    def insert_window(self, window):
        """lib/ttessel.cpp:60-120.  ``window`` = list of (outer, [holes])."""
        self.window = [(simplify(list(o)), [simplify(list(h)) for h in hs])
                       for o, hs in window]
        for outer, holes in self.window:
            bnds = boundaries((outer, holes))
            inner_face = None
            for bi, poly in enumerate(bnds):
                cont = self.unbounded if bi == 0 else inner_face
                hs = self.insert_polygon(poly, cont)
                for e in hs:
                    e.set_length()
                    e.twin.length = e.length
                    set_junction(e, None)
                    set_junction(None, e)
                    e.dir = True
                    e.twin.dir = False
                    s = Seg()
                    s.set_halfedge_handle(e)
                    e.seg = s
                    e.twin.seg = s
                    self.all_segments.append(s)
                e = hs[-1]
                if poly_area(poly) < 0:
                    e.twin.face.data = False
                else:
                    e.face.data = True
                    inner_face = e.face

    def is_on_boundary_he(self, e):
        """lib/ttessel.cpp:574-588."""
        if e.face.data and e.twin.face.data:
            return 0
        return 1 if e.face.data else 2

    def is_on_boundary_seg(self, s):
        return self.is_on_boundary_he(s.get_halfedge_handle()) > 0

    def split_edge(self, e, p):
        """lib/ttessel.cpp:606-655."""
        e_next_hf = e.next_hf
        s = e.seg
        if s.get_halfedge_handle() is e:
            s.set_halfedge_handle(None)
        e_dir = e.dir
        e1 = self.arr_split_edge(e, p)
        e2 = e1.next
        set_junction(e1, e2)
        set_junction(e2, e_next_hf)
        e1.dir = e_dir
        e1.twin.dir = not e1.dir
        e2.dir = e1.dir
        e2.twin.dir = not e2.dir
        e1.set_length()
        e2.set_length()
        e1.twin.length = e1.length
        e2.twin.length = e2.length
        for x in (e1, e1.twin, e2, e2.twin):
            x.seg = s
        s.set_halfedge_handle(e1)
        return e1

    def merge_edge(self, e1, e2):
        """lib/ttessel.cpp:658-698."""
        hf_avant = e1.prev_hf
        hf_apres = e2.next_hf
        new_length = e1.length + e2.length
        s = e1.seg
        d = e1.dir
        hf = self.arr_merge_edge(e1, e2)
        set_junction(hf_avant, hf)
        set_junction(hf, hf_apres)
        hf.length = new_length
        hf.twin.length = new_length
        hf.dir = d
        hf.twin.dir = not d
        hf.seg = s
        hf.twin.seg = s
        s.set_halfedge_handle(hf)
        return hf

    def add_edge(self, v1, v2):
        """lib/ttessel.cpp:700-740."""
        e = self.arr_insert_at_vertices(v1, v2)
        e.set_length()
        e.twin.length = e.length
        set_junction(e, compute_next_hf(e))
        set_junction(compute_prev_hf(e), e)
        if e.prev_hf is not None or e.next_hf is not None:
            # LineTes_halfedge::set_dir(), lib/ttessel.cpp:1480-1488
            e.dir = e.next_hf.dir if e.next_hf is not None else e.prev_hf.dir
        else:
            e.dir = True
        e.twin.dir = not e.dir
        if e.next_hf is None and e.prev_hf is None:
            s = Seg()
            s.set_halfedge_handle(e)
            self.all_segments.append(s)
        elif e.next_hf is not None:
            s = e.next_hf.seg
        else:
            s = e.prev_hf.seg
        e.seg = s
        e.twin.seg = s
        return e

    def remove_edge(self, e):
        """lib/ttessel.cpp:743-780."""
        if self.is_on_boundary_he(e) != 0:
            return None
        prev_hf, next_hf = e.prev_hf, e.next_hf
        set_junction(prev_hf, None)
        set_junction(None, next_hf)
        s = e.seg
        if next_hf is None and prev_hf is None:
            self.all_segments.remove(s)
        elif prev_hf is not None:
            s.set_halfedge_handle(prev_hf)
        else:
            s.set_halfedge_handle(next_hf)
        return self.arr_remove_edge(e)

    def split_from_edge(self, e1, p1, e2, p2):
        """lib/ttessel.cpp:131-152."""
        if self.is_on_boundary_he(e1) == 2:
            return None
        e1_new = self.split_edge(e1, p1)
        e2_new = self.split_edge(e2, p2)
        new_e = self.add_edge(e1_new.tgt, e2_new.tgt)
        new_e.face.data = True
        new_e.twin.face.data = True
        return new_e

    def split_from_vertex(self, v, e, p):
        """lib/ttessel.cpp:161-189 (degree-1 merge branch not on the path)."""
        if self.is_on_boundary_he(e) == 2:
            return None
        if v.degree() == 1:
            raise NotImplementedError("split_from_vertex from an I-vertex")
        e2 = self.split_edge(e, p)
        new_e = self.add_edge(v, e2.tgt)
        new_e.face.data = True
        new_e.twin.face.data = True
        return new_e

    def suppress_edge(self, e):
        """lib/ttessel.cpp:201-236."""
        if self.is_on_boundary_he(e) > 0 or (e.prev_hf is not None
                                             and e.next_hf is not None):
            return None
        v1, v2 = e.src, e.tgt
        f = self.remove_edge(e)
        if f is None:
            return None
        for v in (v1, v2):
            inc = v.incoming()
            if len(inc) == 2 and v.inc.next_hf is not None:
                self.merge_edge(v.inc, v.inc.next_hf)
        f.data = True
        return f

    def number_of_window_edges(self):
        return sum(len(b) for w in self.window for b in boundaries(w))

    def internal_faces(self):
        return [f for f in self.faces if f.data]


# ----------------------------------------------------------------------------
# polygon helpers (reference lib/ttessel.cpp:4314-4336, :4453-4506, :4600-5064)
# ----------------------------------------------------------------------------
def boundaries(hp):
    """lib/ttessel.cpp:4453-4469.  hp = (outer, holes).  CGAL's
    reverse_orientation keeps the first vertex in place."""
    outer, holes = hp
    b = []
    ob = list(outer)
    if poly_area(ob) < 0:
        ob = [ob[0]] + ob[:0:-1]
    b.append(ob)
    for h in holes:
        p = list(h)
        if poly_area(p) > 0:
            p = [p[0]] + p[:0:-1]
        b.append(p)
    return b


def simplify(p):
    """lib/ttessel.cpp:4474-4492: drop vertices collinear with their
    neighbours; order preserved, scan starts at vertex 0."""
    n = len(p)
    if n == 0:
        raise ValueError("simplify(Polygon), input polygon is empty")
    sp = []
    for i in range(n):
        if not collinear(p[(i - 1) % n], p[i], p[(i + 1) % n]):
            sp.append(p[i])
    return sp


def simplify_h(hp):
    """lib/ttessel.cpp:4495-4506."""
    b = boundaries(hp)
    return (simplify(b[0]), [simplify(h) for h in b[1:]])


def face2poly(f, simplify_flag=True):
    """lib/ttessel.cpp:4314-4336."""
    ccbs = [f.outer] + list(f.holes)
    polys = []
    for c in ccbs:
        poly = []
        for e in ccb(c):
            if (not simplify_flag) or e.next_hf is None or (e.next_hf is not e.next):
                poly.append(e.tgt.p)
        polys.append(poly)
    return (polys[0], polys[1:])


class PECirc:
    """Polygon edge circulator: (polygon list, index of the edge's source)."""
    __slots__ = ("poly", "i")

    def __init__(self, poly, i):
        self.poly = poly
        self.i = i % len(poly)

    def source(self):
        return self.poly[self.i]

    def target(self):
        return self.poly[(self.i + 1) % len(self.poly)]

    def inc(self):
        return PECirc(self.poly, self.i + 1)

    def same(self, other):
        return self.poly is other.poly and self.i == other.i


def find_edge_in_polygons(polys, seg):
    """lib/ttessel.cpp:5046-5064.  Returns (index, circulator)."""
    for i, poly in enumerate(polys):
        n = len(poly)
        for k in range(n):
            if poly[k] == seg[0] and poly[(k + 1) % n] == seg[1]:
                return i, PECirc(poly, k)
    return len(polys), None


def polygon_insert_edge(e1, e2, p1, p2):
    """lib/ttessel.cpp:4600-4631."""
    res = [p1, p2]
    e = e2
    if e.target() == p2:
        e = e.inc()
    while True:
        res.append(e.target())
        e = e.inc()
        if e.same(e2) or e.same(e1):
            break
    if e.same(e1):
        if e1.source() == p1:
            res.pop()
        return res
    res.append(p2)
    res.append(p1)
    e = e1
    if e1.target() == p1:
        e = e.inc()
    while True:
        res.append(e.target())
        e = e.inc()
        if e.same(e1):
            break
    if e1.source() == p1:
        res.pop()
    return res


def filter_holes(holes, poly):
    """lib/ttessel.cpp:4996-5009: which holes have their first vertex strictly inside poly."""
    return [bounded_side_2(poly, h[0]) > 0 for h in holes]


def hpolygon_insert_edge(hpoly, i1, i2, e1, e2, p1, p2):
    """lib/ttessel.cpp:4659-4795.  hpoly = [outer, hole, ...] as given by boundaries();
    returns a list of (outer, holes): two faces when the inserted segment joins two points
    of the outer boundary (cas 1) or of one hole (cas 2), one modified face when it
    joins the outer boundary to a hole (cas 3) or two holes (cas 4)."""
    if len(hpoly) == 1:
        cas = 1
    elif i1 == i2:
        cas = 1 if i1 == 0 else 2
    elif i1 == 0 or i2 == 0:
        cas = 3
    else:
        cas = 4
    if cas == 1:
        outer1 = polygon_insert_edge(e1, e2, p1, p2)
        outer2 = polygon_insert_edge(e2, e1, p2, p1)
        holes1, holes2 = [], []
        if len(hpoly) > 1:
            inside = filter_holes(hpoly[1:], outer1)
            for ii in range(1, len(hpoly)):
                (holes1 if inside[ii - 1] else holes2).append(hpoly[ii])
        return [(outer1, holes1), (outer2, holes2)]
    if cas == 2:
        # mother face (the old face minus the daughter) and daughter face (along the hole)
        outer1 = hpoly[0]
        c21 = polygon_insert_edge(e2, e1, p2, p1)
        if bounded_side_2(c21, e2.target()) >= 0:   # != ON_UNBOUNDED_SIDE
            outer2 = polygon_insert_edge(e1, e2, p1, p2)
            enlarged_hole = c21
        else:
            outer2 = c21
            enlarged_hole = polygon_insert_edge(e1, e2, p1, p2)
        inside = filter_holes(hpoly[1:], outer2)
        holes1, holes2 = [], []
        for ii in range(1, len(hpoly)):
            if ii == i1:
                holes1.append(enlarged_hole)
            elif inside[ii - 1]:
                holes2.append(hpoly[ii])
            else:
                holes1.append(hpoly[ii])
        return [(outer1, holes1), (outer2, holes2)]
    if cas == 3:
        # the outer boundary folds into the face and takes the hole in
        outer1 = polygon_insert_edge(e1, e2, p1, p2)
        j = i1 if i1 > 0 else i2
        return [(outer1, [hpoly[ii] for ii in range(1, len(hpoly)) if ii != j])]
    # cas 4: two holes become one
    holes1 = [hpoly[ii] for ii in range(1, len(hpoly)) if ii != i1 and ii != i2]
    holes1.append(polygon_insert_edge(e1, e2, p1, p2))
    return [(hpoly[0], holes1)]


def polygon_remove_edge(e1, e2):
    """lib/ttessel.cpp:4945-4966."""
    res = []
    circ = e1.inc()
    while not circ.same(e1) and not circ.same(e2):
        res.append(circ.target())
        circ = circ.inc()
    if circ.same(e2):
        return res
    circ = e2.inc()
    while not circ.same(e2):
        res.append(circ.target())
        circ = circ.inc()
    return res


def hpolygon_remove_edge(hpoly, hpoly_twin, i, i_twin, e, e_twin, single):
    """lib/ttessel.cpp:4816-4925.  Returns (outer, holes)."""
    if (not single) and len(hpoly) == 1 and len(hpoly_twin) == 1:
        cas = 1
    elif single and i == i_twin:
        cas = 3 if i == 0 else 4
    elif i > 0:
        cas = 2
    else:
        cas = 1 if i_twin == 0 else 2
    if cas == 1:
        # the edge separates two faces: they merge, holes of both are kept
        return (polygon_remove_edge(e, e_twin), list(hpoly[1:]) + list(hpoly_twin[1:]))
    if cas == 2:
        # the edge joins two vertices of one hole of the mother face: the daughter face
        # on its other side dissolves into the mother
        mother, daughter = (hpoly, hpoly_twin) if i > 0 else (hpoly_twin, hpoly)
        j = i if i > 0 else i_twin
        holes = list(daughter[1:])
        for ii in range(1, len(mother)):
            if ii != j:
                holes.append(mother[ii])
            else:
                holes.append(polygon_remove_edge(e, e_twin) if i > 0 else polygon_remove_edge(e_twin, e))
        return (mother[0], holes)
    if cas == 3:
        # an isthmus from the outer boundary to what becomes a hole again
        outer_new = polygon_remove_edge(e, e_twin)
        inner_new = polygon_remove_edge(e_twin, e)
        if poly_area(outer_new) <= 0:
            outer_new, inner_new = inner_new, outer_new
        return (outer_new, list(hpoly[1:]) + [inner_new])
    # cas 4: an isthmus between two holes: the merged hole divides again
    holes = list(hpoly[1:])
    del holes[i - 1]
    holes.append(polygon_remove_edge(e, e_twin))
    holes.append(polygon_remove_edge(e_twin, e))
    return (hpoly[0], holes)


# ----------------------------------------------------------------------------
# ray casting (reference lib/ttessel.cpp:4515-4576)
# ----------------------------------------------------------------------------
def _ray_segment_point(src, d, a, b):
    """CGAL::intersection(Ray_2, Segment_2) when the result is a single point;
    None otherwise (empty, or a collinear overlap which is not a Point)."""
    ex, ey = b[0] - a[0], b[1] - a[1]
    den = _cross(d[0], d[1], ex, ey)
    wx, wy = a[0] - src[0], a[1] - src[1]
    if den == 0:
        # parallel.  A single-point result is only possible at the ray source.
        if _cross(wx, wy, d[0], d[1]) != 0:
            return None
        ta = wx * d[0] + wy * d[1]
        tb = (b[0] - src[0]) * d[0] + (b[1] - src[1]) * d[1]
        if max(ta, tb) == 0:
            return src
        return None
    t = _cross(wx, wy, ex, ey) / den
    s = _cross(wx, wy, d[0], d[1]) / den
    if t < 0 or s < 0 or s > 1:
        return None
    return (src[0] + t * d[0], src[1] + t * d[1])


def ray_exit_face(src, d, f):
    """lib/ttessel.cpp:4515-4576.  Returns (point, halfedge)."""
    all_edges = ccb(f.outer)
    for ho in f.holes:
        all_edges.extend(ccb(ho))
    lp = 1.7976931348623157e308  # std::numeric_limits<double>::max()
    res = None
    e = None
    found = False
    for ce in all_edges:
        ip = _ray_segment_point(src, d, ce.src.p, ce.tgt.p)
        if ip is None:
            continue
        ln = math.sqrt(float(sqdist(src, ip)))
        if ln == 0:
            continue
        if found and ce is e.twin:
            if orientation(ce.src.p, ce.tgt.p, src) > 0:
                e = ce
            continue
        if ln < lp:
            res = ip
            found = True
            e = ce
            lp = ln
    if not found:
        raise RuntimeError("ray_exit_face: intersection of ray with face boundaries not found")
    return res, e


# ----------------------------------------------------------------------------
# TTessel and its modifications (reference lib/ttessel.cpp:1506-2713)
# ----------------------------------------------------------------------------
class ModList:
    """include/ttessel.h:489-498."""

    def __init__(self):
        self.del_vertices = []
        self.add_vertices = []
        self.del_edges = []
        self.add_edges = []
        self.del_faces = []
        self.add_faces = []
        self.del_segs = []
        self.add_segs = []


class Split:
    """lib/ttessel.cpp:2090-2232."""

    def __init__(self, e, x=0.5, angle=CGAL_PI / 2):
        self.e1 = e
        self._x, self._angle = x, angle
        sx, sy = e.src.p
        tx, ty = e.tgt.p
        e1_angle = math.atan2(float(ty - sy), float(tx - sx))
        l_angle = e1_angle + angle
        a = math.sin(l_angle)
        b = -math.cos(l_angle)
        u = a * float(sx) + b * float(sy)
        v = a * float(tx) + b * float(ty)
        w = u + x * (v - u)
        A, B, C = Fr(a), Fr(b), Fr(-w)
        # p1 = Segment(e1) ∩ Line(A,B,C)   (:2110-2117)
        ss = A * sx + B * sy + C
        st = A * tx + B * ty + C
        if ss == st:
            raise RuntimeError("Split constructor: l does not hit e1")
        t = ss / (ss - st)
        if t < 0 or t > 1:
            raise RuntimeError("Split constructor: l does not hit e1")
        self.p1 = (sx + t * (tx - sx), sy + t * (ty - sy))
        # l_dir (:2121-2126); Line_2(a,b,c) has direction (b,-a)
        if st < 0:
            d = (B, -A)
        else:
            d = (-B, A)
        self.p2, self.e2 = ray_exit_face(self.p1, d, e.face)
        self.line = (A, B, C)

    def get_e1(self):
        return self.e1

    def get_e2(self):
        return self.e2

    def get_p1(self):
        return self.p1

    def get_p2(self):
        return self.p2

    def modified_elements(self):
        m = ModList()
        e1, e2, p1, p2 = self.e1, self.e2, self.p1, self.p2
        m.add_vertices += [p1, p2]
        m.del_edges.append((e1.src.p, e1.tgt.p))
        m.del_edges.append((e2.src.p, e2.tgt.p))
        m.add_edges += [(e1.src.p, p1), (p1, e1.tgt.p),
                        (e2.src.p, p2), (p2, e2.tgt.p), (p1, p2)]
        m.del_faces.append(face2poly(e1.face))
        fp = face2poly(e1.face, False)
        fb = boundaries(fp)
        i1, pe1 = find_edge_in_polygons(fb, (e1.src.p, e1.tgt.p))
        i2, pe2 = find_edge_in_polygons(fb, (e2.src.p, e2.tgt.p))
        for hp in hpolygon_insert_edge(fb, i1, i2, pe1, pe2, p1, p2):
            m.add_faces.append(simplify_h(hp))
        for (e, p) in ((e1, p1), (e2, p2)):
            seg = e.seg.list_of_points()
            m.del_segs.append(list(seg))
            key = e.tgt.p if e.dir else e.src.p
            seg.insert(seg.index(key), p)
            m.add_segs.append(seg)
        m.add_segs.append([p1, p2])
        return m


class Merge:
    """lib/ttessel.cpp:2299-2404, include/ttessel.h:657-661."""

    def __init__(self, e):
        self.e = e

    def get_e(self):
        return self.e

    def e1(self):
        return self.e.twin.next

    def e2(self):
        return self.e.next

    def p1(self):
        return self.e.src.p

    def p2(self):
        return self.e.tgt.p

    def modified_elements(self):
        m = ModList()
        e = self.e
        e1, e2, p1, p2 = self.e1(), self.e2(), self.p1(), self.p2()
        m.del_vertices += [p1, p2]
        m.del_edges.append((e.src.p, e.tgt.p))
        m.del_edges.append((p2, e2.tgt.p))
        m.del_edges.append((e2.prev_hf.src.p, p2))
        m.del_edges.append((p1, e1.tgt.p))
        m.del_edges.append((e1.prev_hf.src.p, p1))
        m.add_edges.append((e2.prev_hf.src.p, e2.tgt.p))
        m.add_edges.append((e1.prev_hf.src.p, e1.tgt.p))
        f = face2poly(e.face, False)
        single = e.face is e.twin.face
        fb = boundaries(f)
        i, pe = find_edge_in_polygons(fb, (e.src.p, e.tgt.p))
        m.del_faces.append(simplify_h(f))
        if single:
            # both sides of e are the same face (an isthmus), :2343-2352
            it, pet = find_edge_in_polygons(fb, (e.tgt.p, e.src.p))
            hp = hpolygon_remove_edge(fb, fb, i, it, pe, pet, True)
            m.add_faces.append(simplify_h(hp))
            return self._segments(m, e1, e2, p1, p2)
        ft = face2poly(e.twin.face, False)
        ftb = boundaries(ft)
        it, pet = find_edge_in_polygons(ftb, (e.tgt.p, e.src.p))
        m.del_faces.append(simplify_h(ft))
        hp = hpolygon_remove_edge(fb, ftb, i, it, pe, pet, single)
        m.add_faces.append(simplify_h(hp))
        return self._segments(m, e1, e2, p1, p2)

    @staticmethod
    def _segments(m, e1, e2, p1, p2):
        """:2366-2403: the two segments e abutted on lose a vertex, e's segment goes."""
        seg = e1.seg.list_of_points()
        m.del_segs.append(list(seg))
        seg.remove(p1)
        m.add_segs.append(seg)
        seg = e2.seg.list_of_points()
        m.del_segs.append(list(seg))
        seg.remove(p2)
        m.add_segs.append(seg)
        m.del_segs.append([p1, p2])
        return m


class Flip:
    """lib/ttessel.cpp:2432-2713."""

    def __init__(self, he):
        self.e1 = he
        s1 = he.seg
        e3 = None
        for c in he.src.incoming():
            if c.seg is not s1:
                e3 = c
                break
        self.e3 = e3
        src = e3.tgt.p
        d = (e3.tgt.p[0] - e3.src.p[0], e3.tgt.p[1] - e3.src.p[1])
        if orientation(he.src.p, he.tgt.p, e3.src.p) > 0:
            f = he.twin.face
            self.right = True
        else:
            f = he.face
            self.right = False
        self.p2, self.e2 = ray_exit_face(src, d, f)

    def get_e1(self):
        return self.e1

    def get_e2(self):
        return self.e2

    def get_p2(self):
        return self.p2

    def to_right(self):
        return self.right

    def modified_elements(self):
        m = ModList()
        e1, e2, p2 = self.e1, self.e2, self.p2
        m.del_vertices.append(e1.tgt.p)
        m.add_vertices.append(p2)
        s_merge = (e1.src.p, e1.tgt.p)
        s1_merge = (e1.next.src.p, e1.next.tgt.p)
        s2_merge = (e1.next.prev_hf.tgt.p, e1.next.prev_hf.src.p)
        p_merge = s1_merge[0]
        s_split = (e1.src.p, p2)
        s1_split = (p2, e2.tgt.p)
        s2_split = (p2, e2.src.p)
        m.del_edges += [s_merge, s1_merge, s2_merge]
        generic = s1_split[1] != p_merge and s2_split[1] != p_merge
        if generic:
            m.del_edges.append((s1_split[1], s2_split[1]))
        m.add_edges.append(s_split)
        if generic:
            m.add_edges += [s1_split, s2_split, (s1_merge[1], s2_merge[1])]
        if s1_split[1] == p_merge:
            m.add_edges += [s2_split, (s1_split[0], s1_merge[1])]
        if s2_split[1] == p_merge:
            m.add_edges += [s1_split, (s2_split[0], s2_merge[1])]
        # faces (:2515-2655)
        f2 = face2poly(e2.face, False)
        f2b = boundaries(f2)
        pt1 = e1.src.p
        i2, pe2 = find_edge_in_polygons(f2b, (e2.src.p, e2.tgt.p))
        if self.right:
            seg_buf = (e1.twin.src.p, pt1)
        else:
            seg_buf = (pt1, e1.tgt.p)
        i1, pe_p1 = find_edge_in_polygons(f2b, seg_buf)
        ins = hpolygon_insert_edge(f2b, i1, i2, pe_p1, pe2, pt1, p2)
        match = [False] * len(ins)
        del_face_f1 = del_face_f1_twin = False
        found = False
        seg_buf = (pt1, e1.tgt.p)
        poly = None
        pe1 = None
        if len(ins) == 2:
            poly = boundaries(ins[1])
            i1, pe1 = find_edge_in_polygons(poly, seg_buf)
            if i1 < len(poly):
                found = True
                match[1] = True
        if not found:
            poly = boundaries(ins[0])
            i1, pe1 = find_edge_in_polygons(poly, seg_buf)
            if i1 < len(poly):
                found = True
                match[0] = True
        if not found:
            poly = boundaries(face2poly(e1.face, False))
            i1, pe1 = find_edge_in_polygons(poly, seg_buf)
            if i1 < len(poly):
                found = True
                del_face_f1 = True
            else:
                raise RuntimeError("Flip::modified_elements, unexpected case: e1 not found in any face")
        seg_buf = (seg_buf[1], seg_buf[0])
        poly_twin = None
        poly_twin_is_poly = False
        found = False
        if match[0]:
            i1t, pe1t = find_edge_in_polygons(poly, seg_buf)
            if i1t < len(poly):
                found = True
                poly_twin_is_poly = True
        else:
            poly_twin = boundaries(ins[0])
            i1t, pe1t = find_edge_in_polygons(poly_twin, seg_buf)
            if i1t < len(poly_twin):
                found = True
                match[0] = True
        if not found and len(ins) == 2:
            if match[1]:
                i1t, pe1t = find_edge_in_polygons(poly, seg_buf)
                if i1t < len(poly):
                    found = True
                    poly_twin_is_poly = True
            else:
                poly_twin = boundaries(ins[1])
                i1t, pe1t = find_edge_in_polygons(poly_twin, seg_buf)
                if i1t < len(poly_twin):
                    found = True
                    match[1] = True
        if not found:
            poly_twin = boundaries(face2poly(e1.twin.face, False))
            i1t, pe1t = find_edge_in_polygons(poly_twin, seg_buf)
            if i1t < len(poly_twin):
                found = True
                del_face_f1_twin = True
                if del_face_f1:
                    raise RuntimeError("Flip::modified_elements, unexpected case")
            else:
                raise RuntimeError("Flip::modified_elements, twin of e1 not found anywhere")
        if poly_twin_is_poly:
            add_face_remove = hpolygon_remove_edge(poly, poly, i1, i1t, pe1, pe1t, True)
        else:
            add_face_remove = hpolygon_remove_edge(poly, poly_twin, i1, i1t, pe1, pe1t, False)
        m.del_faces.append(simplify_h(f2))
        if del_face_f1:
            m.del_faces.append(simplify_h((poly[0], poly[1:])))
        if del_face_f1_twin:
            m.del_faces.append(simplify_h((poly_twin[0], poly_twin[1:])))
        for k in range(len(ins)):
            if not match[k]:
                m.add_faces.append(simplify_h(ins[k]))
        m.add_faces.append(simplify_h(add_face_remove))
        # segments (:2659-2712)
        seg = e1.seg.list_of_points()
        m.del_segs.append(list(seg))
        if e1.dir:
            seg.pop()
        else:
            seg.pop(0)
        m.add_segs.append(seg)
        e3 = self.e3
        seg = e3.seg.list_of_points()
        m.del_segs.append(list(seg))
        if e3.dir:
            seg.append(p2)
        else:
            seg.insert(0, p2)
        m.add_segs.append(seg)
        seg = e2.seg.list_of_points()
        m.del_segs.append(list(seg))
        key = e2.tgt.p if e2.dir else e2.src.p
        seg.insert(seg.index(key), p2)
        if e1.next.seg is e2.seg:
            seg.remove(e1.tgt.p)
            m.add_segs.append(seg)
            return m
        m.add_segs.append(seg)
        seg = e1.next.seg.list_of_points()
        m.del_segs.append(list(seg))
        seg.remove(e1.tgt.p)
        m.add_segs.append(seg)
        return m


class Rng:
    """Stand-in for the global ``CGAL::Random *rnd`` (include/ttessel.h:139).
    CGAL::Random is un-vendored and unpinned, so the stream itself is not
    reproduced: numpy PCG64, same call pattern (Appendix B of SURVEY.md)."""

    def __init__(self, seed):
        self.g = np.random.Generator(np.random.PCG64(seed))

    def get_double(self, lo, hi):
        return lo + (hi - lo) * float(self.g.random())

    def get_int(self, lo, hi):
        return int(self.g.integers(lo, hi))


class TTessel(LineTes):
    """lib/ttessel.cpp:1506-2008."""

    def __init__(self, rng=None):
        super().__init__()
        self.blocking_segments = []
        self.non_blocking_segments = []
        self.int_length = 0.0
        self.rnd = rng
        self.arg_order = "gxx"  # g++ evaluates call arguments right to left

    def insert_window(self, window):
        super().insert_window(window)
        self.int_length = 0.0
        for w in window:
            for b in boundaries(w):
                n = len(b)
                for i in range(n):
                    self.int_length += math.sqrt(float(sqdist(b[i], b[(i + 1) % n])))

    def insert_window_rect(self, x0, y0, x1, y1):
        # Iso_rectangle_2 vertices 0..3 are counter-clockwise from (xmin,ymin)
        pts = [_pt(x0, y0), _pt(x1, y0), _pt(x1, y1), _pt(x0, y1)]
        self.insert_window([(pts, [])])

    def get_total_internal_length(self):
        return self.int_length

    def number_of_blocking_segments(self):
        return len(self.blocking_segments)

    def number_of_non_blocking_segments(self):
        return len(self.non_blocking_segments)

    # -- updates --
    def update_split(self, spl):
        """lib/ttessel.cpp:1580-1597."""
        e = self.split_from_edge(spl.e1, spl.p1, spl.e2, spl.p2)
        self.int_length += 2 * e.length
        self.non_blocking_segments.append(e.seg)
        for s in (e.next.seg, e.twin.next.seg):
            if (not self.is_on_boundary_seg(s)) and not s.number_of_edges_is_greater_than(2):
                self.non_blocking_segments.remove(s)
                self.blocking_segments.append(s)
        return e

    def update_merge(self, me):
        """lib/ttessel.cpp:1604-1638."""
        e = me.e
        s = e.seg
        self.non_blocking_segments.remove(s)
        for sx in (e.next.seg, e.twin.next.seg):
            if (not self.is_on_boundary_seg(sx)) and not sx.number_of_edges_is_greater_than(2):
                self.non_blocking_segments.append(sx)
                self.blocking_segments.remove(sx)
        suppressed_length = e.length
        f = self.suppress_edge(e)
        self.int_length -= 2 * suppressed_length
        return f

    def update_flip(self, flip):
        """lib/ttessel.cpp:1644-1701."""
        e1 = flip.e1
        s1 = e1.seg
        s1_i = e1.next.seg
        e2c = None
        for c in e1.src.incoming():
            if c.seg is not s1:
                e2c = c
                break
        s2 = e2c.seg
        new_e = self.split_from_vertex(e1.src, flip.e2, flip.p2)
        self.int_length += 2 * new_e.length
        self.int_length -= 2 * e1.length
        s2_i = new_e.next.seg
        self.suppress_edge(e1)
        if not s2.number_of_edges_is_greater_than(2):
            self.non_blocking_segments.remove(s2)
            self.blocking_segments.append(s2)
        if not s1.number_of_edges_is_greater_than(1):
            self.blocking_segments.remove(s1)
            self.non_blocking_segments.append(s1)
        if s2_i is not s1_i:
            if (not self.is_on_boundary_seg(s2_i)) and not s2_i.number_of_edges_is_greater_than(2):
                self.non_blocking_segments.remove(s2_i)
                self.blocking_segments.append(s2_i)
            if (not self.is_on_boundary_seg(s1_i)) and not s1_i.number_of_edges_is_greater_than(1):
                self.blocking_segments.remove(s1_i)
                self.non_blocking_segments.append(s1_i)
        return new_e

    def update(self, modif):
        if isinstance(modif, Split):
            return self.update_split(modif)
        if isinstance(modif, Merge):
            return self.update_merge(modif)
        return self.update_flip(modif)

    # -- proposals / enumeration --
    def _draw_x_angle(self):
        # Split(e, rnd->get_double(0,1), acos(1-2*rnd->get_double(0,1))):
        # argument evaluation order unspecified (lib/ttessel.cpp:1757, :1792)
        if self.arg_order == "gxx":
            u = self.rnd.get_double(0, 1)
            x = self.rnd.get_double(0, 1)
        else:
            x = self.rnd.get_double(0, 1)
            u = self.rnd.get_double(0, 1)
        return x, math.acos(1 - 2 * u)

    def alt_length_weighted_random_halfedge(self):
        """lib/ttessel.cpp:1948-1966."""
        select = self.rnd.get_double(0, self.int_length)
        cumlen = 0.0
        for e in self.halfedges:
            if not e.face.data:
                continue
            cumlen += e.length
            if cumlen > select:
                return e
        return None

    def alt_length_weighted_random_halfedges(self, n):
        """lib/ttessel.cpp:1968-1991."""
        til = self.int_length
        select = [self.rnd.get_double(0, til) for _ in range(n)]
        select.sort(reverse=True)
        res = []
        cumlen = 0.0
        for e in self.halfedges:
            if not e.face.data:
                continue
            cumlen += e.length
            while select and cumlen > select[-1]:
                res.append(e)
                select.pop()
            if not select:
                break
        return res

    def propose_split(self):
        e = self.alt_length_weighted_random_halfedge()
        x, ang = self._draw_x_angle()
        return Split(e, x, ang)

    def propose_merge(self):
        n = self.rnd.get_int(0, len(self.non_blocking_segments))
        return Merge(self.non_blocking_segments[n].get_halfedge_handle())

    def propose_flip(self):
        n = self.rnd.get_int(0, len(self.blocking_segments))
        s = self.blocking_segments[n]
        side = self.rnd.get_int(0, 2)
        e = s.halfedges_start().twin if side == 0 else s.halfedges_end()
        return Flip(e)

    def split_sample(self, n):
        """lib/ttessel.cpp:1783-1795."""
        es = self.alt_length_weighted_random_halfedges(n)
        res = []
        for e in es:
            x, ang = self._draw_x_angle()
            res.append(Split(e, x, ang))
        return res

    def all_merges(self):
        return [Merge(s.get_halfedge_handle()) for s in self.non_blocking_segments]

    def all_flips(self):
        """lib/ttessel.cpp:1834-1846."""
        n = len(self.blocking_segments)
        flips = [None] * (2 * n)
        for i, s in enumerate(self.blocking_segments):
            flips[i] = Flip(s.halfedges_start().twin)
            flips[i + n] = Flip(s.halfedges_end())
        return flips

    # -- validity (lib/ttessel.cpp:1721-1753 and :1307-1329 in spirit) --
    def is_valid(self):
        for s in self.non_blocking_segments:
            if s.number_of_edges() != 1:
                return False
        for s in self.blocking_segments:
            if s.number_of_edges() < 2:
                return False
        for h in self.halfedges:
            if h.next.prev is not h or h.twin.twin is not h or h.next.src is not h.tgt:
                return False
        return True


# ----------------------------------------------------------------------------
# feature functions (reference lib/ttessel.cpp:4124-4434), as written
# ----------------------------------------------------------------------------
def is_inside(pt, hp):
    """lib/ttessel.cpp:3583-3597."""
    outer, holes = hp
    if bounded_side_2(outer, pt) != ON_BOUNDED_SIDE:
        return False
    for h in holes:
        if bounded_side_2(h, pt) != ON_UNBOUNDED_SIDE:
            return False
    return True


def clip_segment_by_polygons(seg, window):
    """clip_segment_by_polygon(Segment, HPolygons), lib/ttessel.cpp:3696-3786, in exact arithmetic:
    the longest piece of ``seg`` = (p, q) inside the union of the *open* polygons with holes
    (boundaries cut the segment and are not part of the result).  The reference builds Nef
    polyhedra; the intersection of a segment with an open polygonal set is what is computed here
    directly: cut the segment at every boundary edge it meets, keep the pieces whose midpoint
    is strictly inside.  Returns (a, b) along the direction of seg; ValueError if there is none
    (std::domain_error there)."""
    p, q = (Fr(seg[0][0]), Fr(seg[0][1])), (Fr(seg[1][0]), Fr(seg[1][1]))
    dx, dy = q[0] - p[0], q[1] - p[1]
    cuts = {Fr(0), Fr(1)}
    for hp in window:
        for b in boundaries(hp):
            n = len(b)
            for i in range(n):
                a, c = b[i], b[(i + 1) % n]
                ex, ey = c[0] - a[0], c[1] - a[1]
                den = dx * ey - dy * ex
                if den != 0:
                    ts = ((a[0] - p[0]) * ey - (a[1] - p[1]) * ex) / den
                    te = ((a[0] - p[0]) * dy - (a[1] - p[1]) * dx) / den
                    if 0 <= ts <= 1 and 0 <= te <= 1:
                        cuts.add(ts)
                elif (a[0] - p[0]) * dy - (a[1] - p[1]) * dx == 0:   # on the segment's line
                    for v in (a, c):
                        t = ((v[0] - p[0]) * dx + (v[1] - p[1]) * dy) / (dx * dx + dy * dy)
                        if 0 < t < 1:
                            cuts.add(t)
    cuts = sorted(cuts)
    best = None
    for t0, t1 in zip(cuts[:-1], cuts[1:]):
        tm = (t0 + t1) / 2
        m = (p[0] + tm * dx, p[1] + tm * dy)
        if any(is_inside(m, hp) for hp in window) and (best is None or t1 - t0 > best[1] - best[0]):
            best = (t0, t1)
    if best is None:
        raise ValueError("clip_segment_by_polygon: the segment does not hit the polygons")
    return ((p[0] + best[0] * dx, p[1] + best[0] * dy), (p[0] + best[1] * dx, p[1] + best[1] * dy))


def arrangement_census(window, segs):
    """What CGAL::insert leaves after LineTes::insert_segment of every segment of ``segs`` (already
    clipped; lib/ttessel.cpp:904-910), counted exactly: (vertices, edges, proper crossings, ends on
    another segment or on the window boundary, free ends).  General position is assumed and
    asserted: no three segments through one point, no overlapping collinear segments, no segment
    through a window vertex."""
    sides = []
    for hp in window:
        for b in boundaries(hp):
            sides += [(b[i], b[(i + 1) % len(b)]) for i in range(len(b))]
    L = sides + [((Fr(s[0][0]), Fr(s[0][1])), (Fr(s[1][0]), Fr(s[1][1]))) for s in segs]
    nw = len(sides)
    on = [set() for _ in L]            # points on each segment besides its two ends
    crossings = 0
    for i in range(len(L)):
        for j in range(max(i + 1, nw), len(L)):
            (a, b), (c, d) = L[i], L[j]
            r = (b[0] - a[0], b[1] - a[1]); w = (d[0] - c[0], d[1] - c[1])
            den = _cross(r[0], r[1], w[0], w[1])
            assert den != 0 or _cross(r[0], r[1], c[0] - a[0], c[1] - a[1]) != 0, "collinear segments"
            if den == 0:
                continue
            t = _cross(c[0] - a[0], c[1] - a[1], w[0], w[1]) / den
            u = _cross(c[0] - a[0], c[1] - a[1], r[0], r[1]) / den
            if 0 < t < 1 and 0 < u < 1:
                crossings += 1
                x = (a[0] + t * r[0], a[1] + t * r[1])
                on[i].add(x); on[j].add(x)
            elif 0 < t < 1 and u in (0, 1):     # an end of j on the interior of i
                on[i].add(c if u == 0 else d)
            elif 0 < u < 1 and t in (0, 1):
                assert i >= nw, "a segment through a window vertex"
                on[j].add(a if t == 0 else b)
    pts = set()
    for k, (a, b) in enumerate(L):
        pts |= {a, b} | on[k]
    edges = sum(1 + len(o) for o in on)
    ends = [e for (a, b) in L[nw:] for e in (a, b)]
    attached = sum(1 for e in ends if any(e in on[k] for k in range(len(L))))
    multiplicity = {}
    for k in range(len(L)):
        for x in on[k] | set(L[k]):
            multiplicity[x] = multiplicity.get(x, 0) + 1
    assert all(m <= 2 for m in multiplicity.values()), "three segments through one point"
    return len(pts), edges, crossings, attached, len(ends) - attached


def is_point_inside_window(pt, t):
    return 1.0 if any(is_inside(pt, hp) for hp in t.window) else 0.0


def is_segment_internal(s, t):
    """lib/ttessel.cpp:4172-4196."""
    p1, p2 = s[0], s[-1]
    if is_point_inside_window(p1, t) > .5 or is_point_inside_window(p2, t) > .5:
        return 1.0
    for outer, holes in t.window:
        for b in [outer] + list(holes):
            n = len(b)
            for i in range(n):
                a, c = b[i], b[(i + 1) % n]
                if _on_segment(p1, a, c) and _on_segment(p2, a, c):
                    return 0.0
    return 1.0


def minus_is_segment_internal(s, t):
    return -is_segment_internal(s, t)


def edge_length(seg, t):
    """lib/ttessel.cpp:4141-4149 (returns the length only for *boundary* edges)."""
    if is_segment_internal([seg[0], seg[1]], t) > 0:
        return 0.0
    return math.sqrt(float(sqdist(seg[0], seg[1])))


def edge_length_internal(seg, t):
    """EXTENSION (not in the reference): length of *internal* edges, the evident
    intent of edge_length's doc comment (lib/ttessel.cpp:4137-4139)."""
    if is_segment_internal([seg[0], seg[1]], t) > 0:
        return math.sqrt(float(sqdist(seg[0], seg[1])))
    return 0.0


def seg_number(s, t):
    return 1.0


def face_number(f, t):
    return 1.0


def face_area_2(f, t):
    """lib/ttessel.cpp:4215-4223."""
    area = 0.0
    for p in boundaries(f):
        a = float(poly_area(p))
        area += a * a
    return area


def face_perimeter(f, t):
    """lib/ttessel.cpp:4229-4239."""
    l = 0.0
    for p in boundaries(f):
        n = len(p)
        for j in range(n):
            l += math.sqrt(float(sqdist(p[j], p[(j + 1) % n])))
    return l


def face_shape(f, t):
    """lib/ttessel.cpp:4247-4249."""
    return face_perimeter(f, t) / (4 * math.pow(face_area_2(f, t), 0.25))


def angle_between_vectors(v1, v2):
    """lib/ttessel.cpp:4255-4268."""
    n1 = v1[0] * v1[0] + v1[1] * v1[1]
    n2 = v2[0] * v2[0] + v2[1] * v2[1]
    p = v1[0] * v2[0] + v1[1] * v2[1]
    if p * p == n1 * n2:
        return 0.0 if p > 0 else CGAL_PI
    return math.acos(float(p) / math.sqrt(float(n1 * n2)))


def _vec(a, b):
    return (b[0] - a[0], b[1] - a[1])


def face_sum_of_angles(f, t):
    """lib/ttessel.cpp:4285-4305."""
    ob = f[0]
    n = len(ob)
    tot = 0.0
    for i in range(n):
        j = (i - 1) % n
        v1 = _vec(ob[i], ob[j])
        v2 = _vec(ob[i], ob[(i + 1) % n])
        a = CGAL_PI / 2 - angle_between_vectors(v1, v2)
        if a >= 0:
            tot += a
    return tot


def min_angle(f, t):
    """lib/ttessel.cpp:4348-4367 (seeded with the turning angle at corner 0)."""
    ob = f[0]
    n = len(ob)
    m = angle_between_vectors(_vec(ob[n - 1], ob[0]), _vec(ob[0], ob[1]))
    for i in range(n):
        j = (i - 1) % n
        a = angle_between_vectors(_vec(ob[i], ob[j]), _vec(ob[i], ob[(i + 1) % n]))
        if a < m:
            m = a
    return m


def segment_size_2(s, t):
    """lib/ttessel.cpp:4424-4434 (non-zero only for *boundary* segments)."""
    if is_segment_internal([s[0], s[-1]], t) > 0:
        return 0.0
    return float((len(s) - 1) * (len(s) - 1))


# Feature ids shared with include/lite_b200.h (LITE_FEAT_*) and category codes.
CAT_VERTICES, CAT_EDGES, CAT_FACES, CAT_SEGS = 0, 1, 2, 3
FEATURES = {
    "is_point_inside_window": (0, CAT_VERTICES, is_point_inside_window),
    "edge_length": (1, CAT_EDGES, edge_length),
    "seg_number": (2, CAT_SEGS, seg_number),
    "is_segment_internal": (3, CAT_SEGS, is_segment_internal),
    "minus_is_segment_internal": (4, CAT_SEGS, minus_is_segment_internal),
    "segment_size_2": (5, CAT_SEGS, segment_size_2),
    "face_number": (6, CAT_FACES, face_number),
    "face_area_2": (7, CAT_FACES, face_area_2),
    "face_perimeter": (8, CAT_FACES, face_perimeter),
    "face_shape": (9, CAT_FACES, face_shape),
    "face_sum_of_angles": (10, CAT_FACES, face_sum_of_angles),
    "min_angle": (11, CAT_FACES, min_angle),
    "edge_length_internal": (12, CAT_EDGES, edge_length_internal),
}
FEATURE_BY_ID = {v[0]: (k, v[1], v[2]) for k, v in FEATURES.items()}


# ----------------------------------------------------------------------------
# Energy (reference lib/ttessel.cpp:2856-3104)
# ----------------------------------------------------------------------------
class Energy:
    """theta/feature registry.  The flattened order of a CatVector is
    vertices, edges, faces, segs (lib/ttessel.cpp:2750-2761)."""

    def __init__(self):
        self.ttes = None
        self.theta = {c: [] for c in range(4)}
        self.features = {c: [] for c in range(4)}
        self.value = 0.0

    def add(self, name, theta):
        fid, cat, fn = FEATURES[name]
        self.features[cat].append(fn)
        self.theta[cat].append(float(theta))

    def feature_names(self):
        inv = {v[2]: k for k, v in FEATURES.items()}
        return [inv[fn] for c in range(4) for fn in self.features[c]]

    def feature_ids(self):
        return [FEATURES[n][0] for n in self.feature_names()]

    def set_ttessel(self, t):
        self.ttes = t

    def get_ttessel(self):
        return self.ttes

    def get_theta(self):
        return [x for c in range(4) for x in self.theta[c]]

    def set_theta(self, flat):
        k = 0
        for c in range(4):
            for i in range(len(self.theta[c])):
                self.theta[c][i] = float(flat[k])
                k += 1

    def dim(self):
        return sum(len(self.theta[c]) for c in range(4))

    def statistic_variation(self, modif):
        """lib/ttessel.cpp:2914-2980."""
        ml = modif.modified_elements()
        t = self.ttes
        out = []
        lists = ((ml.del_vertices, ml.add_vertices), (ml.del_edges, ml.add_edges),
                 (ml.del_faces, ml.add_faces), (ml.del_segs, ml.add_segs))
        for c in range(4):
            dels, adds = lists[c]
            for i in range(len(self.theta[c])):
                fn = self.features[c][i]
                loc = 0.0
                for el in dels:
                    loc -= fn(el, t)
                for el in adds:
                    loc += fn(el, t)
                out.append(loc)
        return out

    def variation(self, modif):
        """lib/ttessel.cpp:2994-3012."""
        sv = self.statistic_variation(modif)
        th = self.get_theta()
        var = 0.0
        for s, q in zip(sv, th):
            var += s * q
        return var


# ----------------------------------------------------------------------------
# SMFChain (reference lib/ttessel.cpp:3121-3311) — used to generate inputs
# ----------------------------------------------------------------------------
class SMFChain:
    def __init__(self, energy, p_split, p_merge):
        self.engy = energy
        self.p_split = p_split
        self.p_merge = p_merge
        self.p_split_merge = p_split + p_merge

    def hasting_split(self, s):
        t = self.engy.ttes
        delta_nb = 1
        for sg in (s.e1.seg, s.e2.seg):
            if sg.number_of_edges() == 1 and not t.is_on_boundary_seg(sg):
                delta_nb -= 1
        r = self.p_merge / self.p_split
        r *= (1 / CGAL_PI) * t.int_length / (len(t.non_blocking_segments) + delta_nb)
        ev = self.engy.variation(s)
        return r * math.exp(-ev), ev

    def hasting_merge(self, m):
        t = self.engy.ttes
        r = self.p_split / self.p_merge
        r *= CGAL_PI * len(t.non_blocking_segments) / (t.int_length - 2 * m.e.length)
        ev = self.engy.variation(m)
        return r * math.exp(-ev), ev

    def hasting_flip(self, f):
        t = self.engy.ttes
        db = 0
        s_short = f.e1.seg
        if s_short.number_of_edges() == 2:
            db -= 1
        if f.e3.seg.number_of_edges() == 1:
            db += 1
        i_short = f.e1.next.seg
        i_ext = f.e2.seg
        if i_short is not i_ext:
            if i_short.number_of_edges() == 2 and not t.is_on_boundary_seg(i_short):
                db -= 1
            if i_ext.number_of_edges() == 1 and not t.is_on_boundary_seg(i_ext):
                db += 1
        sb = len(t.blocking_segments)
        r = float(sb) if sb == -db else 1.0 * sb / (sb + db)
        ev = self.engy.variation(f)
        return r * math.exp(-ev), ev

    def step(self, n):
        t = self.engy.ttes
        rnd = t.rnd
        counts = [0] * 6
        for _ in range(n):
            r0 = rnd.get_double(0, 1)
            x = rnd.get_double(0, 1)
            if r0 <= self.p_split:
                counts[0] += 1
                s = t.propose_split()
                r, ev = self.hasting_split(s)
                if r > 1 or x < r:
                    counts[1] += 1
                    t.update(s)
                    self.engy.value += ev
            elif r0 <= self.p_split_merge:
                counts[2] += 1
                if len(t.non_blocking_segments) > 0:
                    m = t.propose_merge()
                    r, ev = self.hasting_merge(m)
                    if r > 1 or x < r:
                        counts[3] += 1
                        t.update(m)
                        self.engy.value += ev
            else:
                counts[4] += 1
                if len(t.blocking_segments) > 0:
                    f = t.propose_flip()
                    r, ev = self.hasting_flip(f)
                    if r > 1 or x < r:
                        counts[5] += 1
                        t.update(f)
                        self.engy.value += ev
        return counts


# ----------------------------------------------------------------------------
# PseudoLikDiscrete / PLInferenceNOIS (reference lib/ttessel.cpp:3319-3540)
# ----------------------------------------------------------------------------
class PseudoLikDiscrete:
    def __init__(self, energy=None):
        self.stat_splits = []
        self.stat_flips = []
        if energy is not None:
            self.SetEnergy(energy)

    def SetEnergy(self, e):
        """lib/ttessel.cpp:3327-3348."""
        self.engy = e
        tes = e.get_ttessel()
        d = e.dim()
        self.sum_stat_merges = [0.0] * d
        for m in tes.all_merges():
            sv = e.statistic_variation(m)
            self.sum_stat_merges = [a + (-1.0 * b) for a, b in zip(self.sum_stat_merges, sv)]
        self.sum_stat_flips = [0.0] * d
        self.stat_flips = []
        for f in tes.all_flips():
            p = e.statistic_variation(f)
            self.stat_flips.append([-1.0 * x for x in p])
            self.sum_stat_flips = [a + (-1.0 * b) for a, b in zip(self.sum_stat_flips, p)]

    def AddSplits(self, n):
        tes = self.engy.get_ttessel()
        for s in tes.split_sample(n):
            self.AddGivenSplit(s)

    def AddGivenSplit(self, s):
        p = self.engy.statistic_variation(s)
        self.stat_splits.append([-1.0 * x for x in p])

    def ClearSplits(self):
        self.stat_splits = []

    def _c(self):
        tes = self.engy.get_ttessel()
        return tes.get_total_internal_length() / (CGAL_PI * len(self.stat_splits))

    def GetValue(self, theta):
        """lib/ttessel.cpp:3387-3404."""
        def dot(a, b):  # sum(theta*(*p)), lib/ttessel.cpp:2801-2817
            return _seqsum([x * y for x, y in zip(a, b)])
        lpl = -dot(theta, self.sum_stat_merges)
        buf = 0.0
        for p in self.stat_splits:
            buf += math.exp(dot(theta, p))
        lpl -= self._c() * buf
        lpl -= dot(theta, self.sum_stat_flips)
        buf = 0.0
        for p in self.stat_flips:
            buf += math.exp(dot(theta, p))
        lpl -= buf
        return lpl

    def GetGradient(self, theta):
        """lib/ttessel.cpp:3412-3429."""
        d = len(theta)
        g = [-1.0 * x for x in self.sum_stat_merges]
        buf = [0.0] * d
        for p in self.stat_splits:
            w = math.exp(_seqsum([x * y for x, y in zip(theta, p)]))
            buf = [b + w * x for b, x in zip(buf, p)]
        c = self._c()
        g = [a - c * b for a, b in zip(g, buf)]
        g = [a - b for a, b in zip(g, self.sum_stat_flips)]
        buf = [0.0] * d
        for p in self.stat_flips:
            w = math.exp(_seqsum([x * y for x, y in zip(theta, p)]))
            buf = [b + w * x for b, x in zip(buf, p)]
        return [a - b for a, b in zip(g, buf)]

    def GetHessian(self, theta):
        """lib/ttessel.cpp:3433-3443."""
        d = len(theta)
        H = [[0.0] * d for _ in range(d)]
        for p in self.stat_splits:
            w = math.exp(_seqsum([x * y for x, y in zip(theta, p)]))
            for i in range(d):
                for j in range(d):
                    H[i][j] -= w * (p[i] * p[j])
        c = self._c()
        H = [[c * x for x in row] for row in H]
        for p in self.stat_flips:
            w = math.exp(_seqsum([x * y for x, y in zip(theta, p)]))
            for i in range(d):
                for j in range(d):
                    H[i][j] -= w * (p[i] * p[j])
        return H


def _seqsum(xs):
    s = 0.0
    for x in xs:
        s += x
    return s


class PLInferenceNOIS(PseudoLikDiscrete):
    """lib/ttessel.cpp:3454-3540.  ``fill_bug`` replicates the reference's
    ``fill()`` (lib/ttessel.cpp:2745-2746: the segs loop never advances j)."""

    def __init__(self, eng, store=False, lam=1.0, fill_bug=True):
        super().__init__(eng)
        self.store_path = store
        self.stepsize = lam
        self.fill_bug = fill_bug
        self.estimates = []
        self.values = []
        self.AddSplits(len(eng.get_ttessel().all_merges()))
        if store:
            self.estimates.append(eng.get_theta())
            self.values.append(self.GetValue(eng.get_theta()))

    def _fill(self, shift_like, x):
        e = self.engy
        out = []
        j = 0
        for c in range(3):
            for _ in range(len(e.theta[c])):
                out.append(x[j])
                j += 1
        for _ in range(len(e.theta[3])):
            out.append(x[j])
            if not self.fill_bug:
                j += 1
        return out

    def Step(self, n=1):
        e = self.engy
        for _ in range(n):
            th = e.get_theta()
            g = np.array(self.GetGradient(th))
            H = np.array(self.GetHessian(th))
            eps = -np.linalg.inv(H) @ g
            shift = self._fill(th, list(eps))
            e.set_theta([a + self.stepsize * b for a, b in zip(th, shift)])
            if self.store_path:
                self.estimates.append(e.get_theta())
            self.AddSplits(len(e.get_ttessel().all_merges()))
            if self.store_path:
                self.values.append(self.GetValue(e.get_theta()))

    def GetEstimate(self):
        return self.engy.get_theta()

    def Run(self, tol=0.05, nmax=100):
        if nmax == 0:
            return 0
        i = 0
        while True:
            prev = self.GetValue(self.GetEstimate())
            self.Step(1)
            cur = self.GetValue(self.GetEstimate())
            i += 1
            if not (abs(cur - prev) >= tol * (abs(cur) + tol) and i < nmax):
                break
        return i


# ----------------------------------------------------------------------------
# text format (reference lib/ttessel.cpp:1331-1441) and bulk construction
# ----------------------------------------------------------------------------
def parse_text(text):
    """Returns (window, segments) with exact rational coordinates."""
    window = []
    cur_outer = None
    cur_holes = []
    segs = []
    for line in text.splitlines():
        toks = line.split()
        if not toks:
            continue
        c = [Fr(tk) for tk in toks]
        if len(c) % 2:
            raise ValueError("Uneven number of coordinates on the current line")
        if len(c) > 4:
            poly = [(c[i], c[i + 1]) for i in range(0, len(c), 2)]
            a = poly_area(poly)
            if a > 0:
                if cur_outer is not None:
                    window.append((cur_outer, cur_holes))
                cur_outer, cur_holes = poly, []
            elif a < 0:
                cur_holes.append(poly)
            else:
                raise ValueError("Invalid orientation of a domain ccb")
        else:
            segs.append(((c[0], c[1]), (c[2], c[3])))
    if cur_outer is not None:
        window.append((cur_outer, cur_holes))
    return window, segs


def write_text(t):
    """LineTes::write, lib/ttessel.cpp:1348-1370."""
    def q(x):
        x = Fr(x)
        return "%d/%d" % (x.numerator, x.denominator)
    lines = []
    for w in t.window:
        for b in boundaries(w):
            lines.append(" ".join(q(p[0]) + " " + q(p[1]) for p in b))
    for s in t.all_segments:
        if not t.is_on_boundary_seg(s):
            a, b = s.pointSource(), s.pointTarget()
            lines.append(" ".join(q(v) for v in (a[0], a[1], b[0], b[1])))
    return "\n".join(lines) + "\n"


def ttessel_from_segments(window, segs, rng=None):
    """Build a TTessel from a window and the internal maximal segments of a
    T-tessellation: what LineTes::read + TTessel(LineTes&) produce
    (lib/ttessel.cpp:1377-1441, :1513-1540), up to CGAL's internal ordering.
    The planar graph is assembled in one go (vertices = all segment ends,
    edges = consecutive points along each segment), so cyclic T-dependencies
    (pinwheels) are fine."""
    import functools
    t = TTessel(rng)
    t.window = [(simplify(list(o)), [simplify(list(h)) for h in hs]) for o, hs in window]
    chains = []  # (a, b, on_window)
    for w in t.window:
        for b in boundaries(w):
            n = len(b)
            for i in range(n):
                chains.append((b[i], b[(i + 1) % n], True))
    for (a, b) in segs:
        chains.append((a, b, False))
    allpts = set()
    for (a, b, _) in chains:
        allpts.add(a)
        allpts.add(b)
    vmap = {}

    def vert(p):
        v = vmap.get(p)
        if v is None:
            v = t._new_vertex(p)
            vmap[p] = v
        return v
    seg_chains = []
    for (a, b, onw) in chains:
        pts = [p for p in allpts if p != a and p != b and _on_segment(p, a, b)]
        pts.sort(key=lambda p: sqdist(a, p))
        pts = [a] + pts + [b]
        hs = []
        for k in range(len(pts) - 1):
            h = t._new_edge(vert(pts[k]), vert(pts[k + 1]))
            hs.append(h)
        seg_chains.append((hs, onw))
    # angular order around each vertex
    out = {}
    for h in t.halfedges:
        out.setdefault(id(h.src), []).append(h)
    ref = (Fr(1), Fr(0))

    def cmpdir(h1, h2):
        d1 = (h1.tgt.p[0] - h1.src.p[0], h1.tgt.p[1] - h1.src.p[1])
        d2 = (h2.tgt.p[0] - h2.src.p[0], h2.tgt.p[1] - h2.src.p[1])
        if _ccw_before(ref, d1, d2):
            return -1
        if _ccw_before(ref, d2, d1):
            return 1
        return 0
    for v in t.vertices:
        os_ = sorted(out[id(v)], key=functools.cmp_to_key(cmpdir))
        k = len(os_)
        for i in range(k):
            o_i = os_[i]
            o_n = os_[(i + 1) % k]
            o_n.twin.next = o_i
            o_i.prev = o_n.twin
        v.inc = os_[0].twin
    # faces
    seen = set()
    cw_cycles = []
    ccw_faces = []
    for h in t.halfedges:
        if id(h) in seen:
            continue
        cyc = ccb(h)
        for e in cyc:
            seen.add(id(e))
        poly = [e.tgt.p for e in cyc]
        if poly_area(poly) > 0:
            f = Face()
            f.outer = h
            t.faces.append(f)
            for e in cyc:
                e.face = f
            ccw_faces.append((f, poly))
        else:
            cw_cycles.append((h, cyc))
    hole_sets = [frozenset(hh) for (_, hs_) in t.window for hh in hs_]
    for f, poly in ccw_faces:
        f.data = frozenset(simplify(poly)) not in hole_sets
    for h, cyc in cw_cycles:
        best = None
        for f, poly in ccw_faces:
            if bounded_side_2(poly, h.tgt.p) == ON_BOUNDED_SIDE:
                a = poly_area(poly)
                if best is None or a < best[1]:
                    best = (f, a)
        cont = t.unbounded if best is None else best[0]
        cont.holes.append(h)
        for e in cyc:
            e.face = cont
    # LiTe payload
    for hs, onw in seg_chains:
        s = Seg()
        for k, h in enumerate(hs):
            h.set_length()
            h.twin.length = h.length
            h.dir = True
            h.twin.dir = False
            h.seg = s
            h.twin.seg = s
        for k in range(len(hs)):
            set_junction(hs[k], hs[k + 1] if k + 1 < len(hs) else None)
        set_junction(None, hs[0])
        s.set_halfedge_handle(hs[0])
        t.all_segments.append(s)
    # TTessel bookkeeping: insert_window (:1563-1574) then TTessel(LineTes&) (:1523-1536)
    t.int_length = 0.0
    for w in window:
        for b in boundaries(w):
            n = len(b)
            for i in range(n):
                t.int_length += math.sqrt(float(sqdist(b[i], b[(i + 1) % n])))
    for s in t.all_segments:
        if t.is_on_boundary_seg(s):
            continue
        seglen = math.sqrt(float(sqdist(s.pointSource(), s.pointTarget())))
        t.int_length += 2 * seglen
        if s.number_of_edges() == 1:
            t.non_blocking_segments.append(s)
        else:
            t.blocking_segments.append(s)
    return t


# ----------------------------------------------------------------------------
# SoA export (the schema of include/lite_b200.h :: lite_tess_soa)
# ----------------------------------------------------------------------------
def export_soa(t):
    """Flatten a TTessel into the structure-of-arrays the C ABI uploads.
    Halfedges are numbered in container order (a2 in SURVEY.md §8 depends on
    it); per-face halfedge lists start at outer_ccb()."""
    vid = {id(v): i for i, v in enumerate(t.vertices)}
    hid = {id(h): i for i, h in enumerate(t.halfedges)}
    fid = {id(f): i for i, f in enumerate(t.faces)}
    sid = {id(s): i for i, s in enumerate(t.all_segments)}
    nH = len(t.halfedges)
    g = lambda h: -1 if h is None else hid[id(h)]
    soa = {}
    soa["vx"] = np.array([float(v.p[0]) for v in t.vertices], dtype=np.float64)
    soa["vy"] = np.array([float(v.p[1]) for v in t.vertices], dtype=np.float64)
    soa["he_src"] = np.array([vid[id(h.src)] for h in t.halfedges], dtype=np.int32)
    soa["he_twin"] = np.array([hid[id(h.twin)] for h in t.halfedges], dtype=np.int32)
    soa["he_next"] = np.array([hid[id(h.next)] for h in t.halfedges], dtype=np.int32)
    soa["he_face"] = np.array([fid[id(h.face)] for h in t.halfedges], dtype=np.int32)
    soa["he_seg"] = np.array([sid[id(h.seg)] for h in t.halfedges], dtype=np.int32)
    soa["he_next_hf"] = np.array([g(h.next_hf) for h in t.halfedges], dtype=np.int32)
    soa["he_prev_hf"] = np.array([g(h.prev_hf) for h in t.halfedges], dtype=np.int32)
    soa["he_dir"] = np.array([1 if h.dir else 0 for h in t.halfedges], dtype=np.uint8)
    soa["he_length"] = np.array([h.length for h in t.halfedges], dtype=np.float64)
    inside, off, cnt, lst = [], [], [], []
    for f in t.faces:
        inside.append(1 if f.data else 0)
        off.append(len(lst))
        if f.data:
            # only outer_ccb() is listed (the layout of lite_tess_soa): the inner boundaries of
            # a face with holes are the other halfedges whose face it is
            c = ccb(f.outer)
            lst.extend(hid[id(h)] for h in c)
            cnt.append(len(c))
        else:
            cnt.append(0)
    soa["face_inside"] = np.array(inside, dtype=np.uint8)
    soa["face_he_offset"] = np.array(off, dtype=np.int32)
    soa["face_he_count"] = np.array(cnt, dtype=np.int32)
    soa["face_he_list"] = np.array(lst, dtype=np.int32)
    soa["seg_he_start"] = np.array([hid[id(s.halfedges_start())] for s in t.all_segments], dtype=np.int32)
    soa["seg_he_end"] = np.array([hid[id(s.halfedges_end())] for s in t.all_segments], dtype=np.int32)
    soa["seg_n_edges"] = np.array([s.number_of_edges() for s in t.all_segments], dtype=np.int32)
    soa["seg_on_boundary"] = np.array([1 if t.is_on_boundary_seg(s) else 0 for s in t.all_segments], dtype=np.uint8)
    soa["non_blocking"] = np.array([sid[id(s)] for s in t.non_blocking_segments], dtype=np.int32)
    soa["blocking"] = np.array([sid[id(s)] for s in t.blocking_segments], dtype=np.int32)
    soa["int_length"] = float(t.int_length)
    assert nH == len(soa["he_src"])
    return soa


# ----------------------------------------------------------------------------
# helpers for the parity tests
# ----------------------------------------------------------------------------
ALL12 = ["is_point_inside_window", "edge_length", "face_number", "face_area_2",
         "face_perimeter", "face_shape", "face_sum_of_angles", "min_angle",
         "seg_number", "is_segment_internal", "minus_is_segment_internal", "segment_size_2"]
ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]


def acs_theta(tau=12.0, fi1=0.05, fi2=1.0):
    """demos/sim_acs.cpp:64-76, flattened in CatVector order edges, faces, faces, segs."""
    return [(tau - 1) / CGAL_PI, tau ** 4 * fi1, fi2, -math.log(tau)]


def make_energy(names, thetas=None, t=None):
    e = Energy()
    for i, n in enumerate(names):
        e.add(n, 0.0 if thetas is None else thetas[i])
    if t is not None:
        e.set_ttessel(t)
    return e


def draw_split_candidates(t, n):
    """The (halfedge, x, angle) list TTessel::split_sample(n) would build
    (lib/ttessel.cpp:1783-1795), as indices into the SoA export order."""
    hid = {id(h): i for i, h in enumerate(t.halfedges)}
    es = t.alt_length_weighted_random_halfedges(n)
    he, xs, angs = [], [], []
    for e in es:
        x, a = t._draw_x_angle()
        he.append(hid[id(e)])
        xs.append(x)
        angs.append(a)
    return (np.array(he, dtype=np.int32), np.array(xs, dtype=np.float64),
            np.array(angs, dtype=np.float64))


def eval_candidates(t, energy, he, xs, angs):
    """Exact evaluation of a fed split list, all merges and all flips.
    Returns a dict of numpy arrays (integer outputs + stored statistics, i.e.
    MINUS the statistic variation as PseudoLikDiscrete keeps them)."""
    hid = {id(h): i for i, h in enumerate(t.halfedges)}
    d = energy.dim()
    out = {}
    n = len(he)
    e2 = np.zeros(n, dtype=np.int32)
    p1 = np.zeros((n, 2))
    p2 = np.zeros((n, 2))
    st = np.zeros((n, d))
    for i in range(n):
        s = Split(t.halfedges[int(he[i])], float(xs[i]), float(angs[i]))
        e2[i] = hid[id(s.e2)]
        p1[i] = [float(s.p1[0]), float(s.p1[1])]
        p2[i] = [float(s.p2[0]), float(s.p2[1])]
        st[i] = [-v for v in energy.statistic_variation(s)]
    out.update(split_e2=e2, split_p1=p1, split_p2=p2, split_stat=st)
    ms = t.all_merges()
    out["merge_he"] = np.array([hid[id(m.e)] for m in ms], dtype=np.int32)
    out["merge_stat"] = np.array([[-v for v in energy.statistic_variation(m)] for m in ms]).reshape(len(ms), d)
    fs = t.all_flips()
    out["flip_e1"] = np.array([hid[id(f.e1)] for f in fs], dtype=np.int32)
    out["flip_e2"] = np.array([hid[id(f.e2)] for f in fs], dtype=np.int32)
    out["flip_right"] = np.array([1 if f.right else 0 for f in fs], dtype=np.int32)
    out["flip_p2"] = np.array([[float(f.p2[0]), float(f.p2[1])] for f in fs]).reshape(len(fs), 2)
    out["flip_stat"] = np.array([[-v for v in energy.statistic_variation(f)] for f in fs]).reshape(len(fs), d)
    return out


def pl_from_stats(theta, split_stat, merge_stat, flip_stat, int_length, n_splits_global=None):
    """lib/ttessel.cpp:3387-3443 on stored statistic arrays (sequential sums)."""
    pl = PseudoLikDiscrete()

    class _T:
        def get_total_internal_length(self_inner):
            return int_length

    class _E:
        def get_ttessel(self_inner):
            return _T()
    pl.engy = _E()
    d = len(theta)
    pl.stat_splits = [list(map(float, r)) for r in split_stat]
    pl.stat_flips = [list(map(float, r)) for r in flip_stat]
    pl.sum_stat_merges = [0.0] * d
    for r in merge_stat:
        pl.sum_stat_merges = [a + float(b) for a, b in zip(pl.sum_stat_merges, r)]
    pl.sum_stat_flips = [0.0] * d
    for r in flip_stat:
        pl.sum_stat_flips = [a + float(b) for a, b in zip(pl.sum_stat_flips, r)]
    if n_splits_global is not None and n_splits_global != len(pl.stat_splits):
        k = len(pl.stat_splits) / float(n_splits_global)
        base_c = pl._c
        pl._c = lambda: base_c() * k
    th = [float(x) for x in theta]
    return pl.GetValue(th), np.array(pl.GetGradient(th)), np.array(pl.GetHessian(th))
