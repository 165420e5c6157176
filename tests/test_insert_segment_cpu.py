"""LineTes::insert_segment and clip_segment_by_polygon on the host model
(reference lib/ttessel.cpp:904-980, :3696-3786), against the reference's own known answers
(tests/check_lite.cpp:544-624) and against the batch reader."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
sys.path.insert(0, HERE)

import lite_b200
import lite_oracle as lo
from lite_b200.capi import LiteError
from test_oracle_golden import holed_polygons


def window_text(window):
    t = lo.TTessel(lo.Rng(0))
    t.insert_window(window)
    return lo.write_text(t)


def find_halfedge(h, a, b, tol=1e-8):
    xy, seg, nx, pv = h.edges()
    for k in range(len(xy)):
        if np.allclose(xy[k], [a[0], a[1], b[0], b[1]], rtol=tol, atol=0):
            return k, seg[k], nx[k], pv[k]
    return None


def test_clip_segment_by_holed_polygon_known_answers():
    """check_lite.cpp:544-560: one polygon with holes; the piece ends on the outer boundary and on
    a hole; a segment that misses the polygon is an error (std::domain_error there)."""
    h = lite_b200.HostTess(text=window_text(holed_polygons()[:1]))
    c = h.clip_segment([-1, 8, 3, 12])
    assert np.array_equal(c, [0, 9, 2, 11])
    with pytest.raises(LiteError, match="does not hit"):
        h.clip_segment([-1, -1, -2, -2])


def test_clip_segment_by_holed_polygons_known_answers():
    """check_lite.cpp:562-579 (the test's 19/2, 9/2, 39/2 are integer divisions: 9, 4, 19): the
    segment runs along a hole's side, through the first polygon (a piece of length 2) and into
    the second (length 5, up to a hole): the longest piece wins."""
    h = lite_b200.HostTess(text=window_text(holed_polygons()))
    assert np.array_equal(h.clip_segment([9, 4, 19, 4]), [14, 4, 19, 4])
    assert np.array_equal(h.clip_segment([19, 4, 9, 4]), [19, 4, 14, 4])   # direction kept
    with pytest.raises(LiteError, match="does not hit"):
        h.clip_segment([-1, -1, -2, -2])
    assert np.array_equal(h.clip_segment([9, 4, 13, 4]), [10, 4, 12, 4])   # hits one polygon only


def test_insert_segment_known_answers():
    """check_lite.cpp:581-624: clipped at both ends (outer boundary, hole) and at one end; the
    halfedge along the clipped segment has no neighbour along its segment."""
    h = lite_b200.HostTess(text=window_text(holed_polygons()))
    n0 = len(h.edges()[0])
    h.insert_segment([3, 12, 3, 16])
    e = find_halfedge(h, (3, 13), (3, 15))
    assert e is not None and e[2] == -1 and e[3] == -1
    assert h.is_t_tessellation()          # both ends on the window's boundaries
    h.check()
    assert len(h.edges()[0]) == n0 + 2 * 3   # two boundary sides split, one new edge
    h.insert_segment([1, 2, 5, 2])
    e = find_halfedge(h, (1, 2), (4, 2))
    assert e is not None and e[2] == -1 and e[3] == -1
    assert not h.is_t_tessellation()      # (1, 2) is a free end
    with pytest.raises(LiteError, match="T-tessellation"):
        h.check()
    with pytest.raises(LiteError, match="T-tessellation"):
        h.export()


def fixture_text():
    return open(os.path.join(HERE, "golden", "tessellation_data.txt")).read()


def soa_edges(s):
    out = {}
    for k in range(len(s["he_src"])):
        p, q = s["he_src"][k], s["he_src"][s["he_twin"][k]]
        out[(float(s["vx"][p]), float(s["vy"][p]), float(s["vx"][q]), float(s["vy"][q]))] = (
            bool(s["face_inside"][s["he_face"][k]]), int(s["seg_n_edges"][s["he_seg"][k]]), bool(s["seg_on_boundary"][s["he_seg"][k]]))
    return out


@pytest.mark.parametrize("order_seed", [0, 1, 2])
def test_segment_by_segment_equals_the_batch_reader(order_seed):
    """LineTes::read inserts the file's segments one at a time (lib/ttessel.cpp:1417-1437); a
    T-tessellation holds cycles of segments each ending on the next, so whatever the order some
    prefix is not a T-tessellation.  Any order ends on the tessellation the batch reader builds."""
    text = fixture_text()
    lines = [l for l in text.splitlines() if l.strip()]
    win = [l for l in lines if len(l.split()) > 4]
    segs = [l for l in lines if len(l.split()) == 4]
    ref = lite_b200.HostTess(text=text)
    rng = np.random.default_rng(order_seed)
    h = lite_b200.HostTess(text="\n".join(win) + "\n")
    was_not_t = False
    for k in rng.permutation(len(segs)):
        c = [float(lo.Fr(v)) for v in segs[k].split()]
        h.insert_segment(c)
        was_not_t = was_not_t or not h.is_t_tessellation()
    assert was_not_t
    assert h.is_t_tessellation()
    h.check()
    assert h.counts() == ref.counts()
    assert soa_edges(h.export().a) == soa_edges(ref.export().a)
    assert h.export().int_length == pytest.approx(ref.export().int_length, rel=1e-14)


def test_crtt_tessellation_rebuilt_by_insertion():
    """300 cells from the CRTT chain, its segments inserted in a random order (over-long by a
    tenth on each side where that stays inside the cell structure is not assumed: the ends are
    exact), then the three moves work on the result."""
    t, _ = lite_b200.generate_crtt(300, seed=11)
    text = t.write_text()
    lines = [l for l in text.splitlines() if l.strip()]
    h = lite_b200.HostTess((0, 0, 1, 1))
    segs = [[float(lo.Fr(v)) for v in l.split()] for l in lines if len(l.split()) == 4]
    for k in np.random.default_rng(3).permutation(len(segs)):
        h.insert_segment(segs[k])
    assert h.is_t_tessellation()
    h.check()
    assert h.counts() == t.counts()
    assert soa_edges(h.export().a) == soa_edges(t.export().a)
    h.update_split(0, 0.4, 1.1)
    h.update_merge(0)
    h.check()


def test_general_arrangements():
    """What the reference's arrangement accepts and remove_[ilx]vertices would clean: a crossing
    (X-vertex), a corner (L-vertex), a free-standing segment, a segment over-running the window."""
    h = lite_b200.HostTess((0, 0, 4, 4))
    a = h.insert_segment([-1, 2, 5, 2])          # clipped to the window: (0,2)-(4,2)
    assert h.is_t_tessellation() and h.counts() == (2, 1, 0)
    b = h.insert_segment([2, 1, 2, 3])           # crosses it, both ends free
    assert not h.is_t_tessellation()
    xy, seg, nx, pv = h.edges()
    assert len(xy) == 2 * (6 + 2 + 2)            # 4 sides (two of them split), 2 + 2 edges
    k = find_halfedge(h, (2, 1), (2, 2))
    assert k is not None and k[1] == b and k[3] == -1 and np.allclose(xy[k[2]], [2, 2, 2, 3])
    k = find_halfedge(h, (0, 2), (2, 2))
    assert k is not None and k[1] == a and np.allclose(xy[k[2]], [2, 2, 4, 2])
    # the free ends closed by two more segments: T again, five cells
    h.insert_segment([2, 3, 2, 10])              # from the free end (L-vertex, collinear) to the top side
    assert not h.is_t_tessellation()
    text = h.write_text()
    assert len([l for l in text.splitlines() if len(l.split()) == 4]) == 3
    # a free-standing segment lies in one face as an inner contour
    g = lite_b200.HostTess((0, 0, 4, 4))
    g.insert_segment([1, 1, 3, 3])
    assert not g.is_t_tessellation()
    xy, seg, nx, pv = g.edges()
    assert len(xy) == 2 * 5
    g.insert_segment([1, 3, 3, 1])               # crosses it
    assert len(g.edges()[0]) == 2 * 8
    with pytest.raises(LiteError, match="T-tessellation"):
        g.update_split(0, 0.5, 1.0)
    with pytest.raises(LiteError, match="does not hit"):
        g.insert_segment([5, 5, 6, 6])


def test_remove_ivertices_known_answers():
    """check_lite.cpp:626-669: a segment with both ends free, close to the boundaries, is
    prolonged at both ends (outer boundary below, hole above); a short free end past a
    crossing is cut back."""
    h = lite_b200.HostTess(text=window_text(holed_polygons()))
    h.insert_segment([5, 3, 5, 9])
    assert h.vertex_census()[0] == 2
    removed, added = h.remove_ivertices()
    assert (removed, added) == (0.0, pytest.approx(1.0, rel=1e-12))
    e = find_halfedge(h, (5, 2.5), (5, 9.5))
    assert e is not None and e[2] == -1 and e[3] == -1
    assert h.is_t_tessellation()
    h.check()
    # the test's Split((0,15)->(0,0), 0.5, pi/2) is the segment (0,7.5)-(10.25,7.5)
    g = lite_b200.HostTess(text=window_text(holed_polygons()))
    g.insert_segment([0, 7.5, 10.25, 7.5])
    assert g.is_t_tessellation()
    g.insert_segment([8.5, 15, 8.5, 7])
    assert find_halfedge(g, (8.5, 7), (8.5, 7.5)) is not None
    assert g.vertex_census()[:3] == (1, 0, 1)
    removed, added = g.remove_ivertices()
    assert (removed, added) == (0.5, 0.0)
    assert find_halfedge(g, (8.5, 7.5), (8.5, 7)) is None
    assert g.is_t_tessellation() and g.vertex_census() == (0, 0, 0, 0)
    g.check()


def test_remove_xvertices_known_answers():
    """check_lite.cpp:1051-1074: two crossings inside the window; none left afterwards."""
    h = lite_b200.HostTess((0, 0, 4, 4))
    for s in ([2, 0, 2, 4], [0, 2, 4, 2], [0, 1, 2, 1], [1, 0, 1, 2]):
        h.insert_segment(s)
    assert h.vertex_census()[3] == 2
    h.remove_xvertices(0.1)
    assert h.vertex_census()[3] == 0
    # breaking the segment (2,0)-(2,4) moved the half on which (0,1)-(2,1) ended: that one now
    # crosses it with a free end next to the crossing ("this procedure can generate new
    # I-vertices", lib/ttessel.cpp:1193), which remove_ivertices takes away
    assert h.vertex_census()[:3] == (1, 0, 1)
    h.remove_ivertices()
    assert h.vertex_census() == (0, 0, 0, 0) and h.is_t_tessellation()
    h.check()
    assert sum(h.counts()[1:]) == 6   # four segments, two of them broken in two


def test_remove_xvertices_recursion_known_answers():
    """check_lite.cpp:1076-1127: one crossing to remove (the others have a free end next to them);
    none left after remove_xvertices(2)."""
    h = lite_b200.HostTess((0, 0, 8, 8))
    segs = [[4, 0, 4, 8], [0, 4, 8, 4], [0, 2, 2, 2], [2, 2, 2, 0], [1, 3.9, 1, 1], [1, 1, 3.9, 1],
            [0, 6, 2, 6], [2, 6, 2, 8], [1, 4.1, 1, 7], [1, 7, 3.9, 7], [4.1, 7, 7, 7], [7, 7, 7, 4.1],
            [6, 8, 6, 6], [6, 6, 8, 6], [4.1, 1, 7, 1], [7, 1, 7, 3.9], [6, 0, 6, 2], [6, 2, 8, 2]]
    for s in segs:
        h.insert_segment(s)
    assert h.vertex_census()[3] == 1
    h.remove_xvertices(2)
    assert h.vertex_census()[3] == 0
    # the free ends at (1, 3.9) and (1, 4.1) are then prolonged onto the same point of the segment
    # between them: a vertex of degree 4 that is neither T nor X, which no pass removes (in the
    # reference either); the census is clean and the count of non-T vertices says what is left
    for _ in range(3):
        h.remove_ivertices()
        h.remove_lvertices()
        h.remove_xvertices(2)
    assert h.vertex_census() == (0, 0, 0, 0) and not h.is_t_tessellation()
    n, first = h.non_t_vertices()
    assert n == 4 and first in [(1.0, 4.0), (4.0, 1.0), (4.0, 7.0), (7.0, 4.0)]
    # the arrangement still holds together: every halfedge of a segment chain is linked both ways
    xy, seg, nx, pv = h.edges()
    for k in range(len(xy)):
        if nx[k] >= 0:
            assert pv[nx[k]] == k and seg[nx[k]] == seg[k] and np.array_equal(xy[nx[k]][:2], xy[k][2:])


def test_cleaning_a_random_line_pattern_gives_a_t_tessellation():
    """The use the three passes are made for (docfiles of the reference, RLiTe's cleaning
    pipeline): random segments in a window -> crossings broken, free ends and corners removed ->
    a T-tessellation the device path accepts."""
    rng = np.random.default_rng(8)
    h = lite_b200.HostTess((0, 0, 1, 1))
    for _ in range(25):
        c = rng.uniform(0.05, 0.95, 2)
        a = rng.uniform(0, np.pi)
        l = rng.uniform(0.1, 0.5)
        d = 0.5 * l * np.array([np.cos(a), np.sin(a)])
        h.insert_segment([*(c - d), *(c + d)])
    free, corners, cross, todo = h.vertex_census()
    assert free > 20 and cross > 5 and not h.is_t_tessellation()
    for _ in range(20):
        h.remove_xvertices(0.01)
        h.remove_ivertices()
        h.remove_lvertices()
        if h.is_t_tessellation():
            break
    assert h.vertex_census() == (0, 0, 0, 0)
    assert h.is_t_tessellation()
    h.check()
    s = h.export()
    assert h.counts()[0] == int(np.sum(s.a["face_inside"])) > 5


@pytest.mark.parametrize("seed", [0, 3, 7, 15, 39, 40, 77])
def test_cleaning_random_patterns_in_plain_and_holed_windows(seed):
    """Random segments in the unit square or (every fourth seed) in the reference's two holed
    polygons, then the three passes in turn until nothing is left to remove: a T-tessellation that
    passes the export assertions.  Seed 39 is a pattern whose shortest edge is far below the
    step (the moved end must not fall back onto the crossing it leaves)."""
    rng = np.random.default_rng(seed)
    holed = seed % 4 == 3
    h = lite_b200.HostTess(text=window_text(holed_polygons())) if holed else lite_b200.HostTess((0, 0, 1, 1))
    sc = 20.0 if holed else 1.0
    for _ in range(int(rng.integers(5, 40))):
        c = rng.uniform(0.05, 0.95, 2) * sc
        a = rng.uniform(0, np.pi)
        d = 0.5 * rng.uniform(0.1, 0.6) * sc * np.array([np.cos(a), np.sin(a)])
        try:
            h.insert_segment([*(c - d), *(c + d)])
        except LiteError as e:     # wholly inside a hole or outside both polygons
            assert "does not hit" in str(e)
    for _ in range(30):
        h.remove_xvertices(0.01 * sc)
        h.remove_ivertices()
        h.remove_lvertices()
        if h.is_t_tessellation():
            break
    assert h.vertex_census() == (0, 0, 0, 0) and h.is_t_tessellation()
    h.check()
    s = h.export()
    assert int(np.sum(s.a["face_inside"])) == h.counts()[0]
    assert lite_b200.HostTess(text=h.write_text()).counts() == h.counts()   # the text round-trips


def test_clip_against_the_exact_oracle():
    """lite::clip_segment (doubles, tolerance 1e-9 of the window) against the oracle's exact
    clip on 400 random segments over the reference's two holed polygons: same verdict (hit or
    not), same piece to 1e-12."""
    win = holed_polygons()
    h = lite_b200.HostTess(text=window_text(win))
    rng = np.random.default_rng(21)
    hits = misses = 0
    for _ in range(400):
        a = rng.uniform(-2, 23, 2)
        if rng.uniform() < 0.5:
            b = a + rng.uniform(-3, 3, 2)       # short: often wholly inside a hole or outside
        else:
            b = rng.uniform(-2, 23, 2)
        seg = [float(a[0]), float(a[1]), float(b[0]), float(b[1])]
        try:
            want = lo.clip_segment_by_polygons(((seg[0], seg[1]), (seg[2], seg[3])), win)
        except ValueError:
            with pytest.raises(LiteError, match="does not hit"):
                h.clip_segment(seg)
            misses += 1
            continue
        got = h.clip_segment(seg)
        assert np.allclose(got, [float(want[0][0]), float(want[0][1]), float(want[1][0]), float(want[1][1])], rtol=0, atol=1e-12 * 23)
        hits += 1
    assert hits > 200 and misses > 20


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_arrangement_against_the_exact_census(seed):
    """The lenient builder (crossings, free ends, ends on other segments) against exact counting:
    numbers of vertices, edges, crossings and free ends of the arrangement of random segments in
    the unit square, some of them over-running the window."""
    rng = np.random.default_rng(seed)
    square = [([lo._pt(0, 0), lo._pt(1, 0), lo._pt(1, 1), lo._pt(0, 1)], [])]
    h = lite_b200.HostTess((0, 0, 1, 1))
    exact = []
    for _ in range(int(rng.integers(10, 30))):
        c = rng.uniform(0.0, 1.0, 2)
        a = rng.uniform(0, np.pi)
        d = 0.5 * rng.uniform(0.2, 0.9) * np.array([np.cos(a), np.sin(a)])
        seg = [float(c[0] - d[0]), float(c[1] - d[1]), float(c[0] + d[0]), float(c[1] + d[1])]
        try:
            exact.append(lo.clip_segment_by_polygons(((seg[0], seg[1]), (seg[2], seg[3])), square))
        except ValueError:
            continue
        h.insert_segment(seg)
    n_v, n_e, n_x, n_attached, n_free = lo.arrangement_census(square, exact)
    xy, seg, nx, pv = h.edges()
    assert len(xy) == 2 * n_e
    assert len({(round(x, 9), round(y, 9)) for x, y, _, _ in xy}) == n_v
    free, corners, crossings, _ = h.vertex_census()
    assert (free, corners, crossings) == (n_free, 0, n_x)
