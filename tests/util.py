"""Shared helpers for the tests (golden loading, tolerances)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

# floating-point parity bar (BASELINE.json north_star): relative 1e-9 on the
# sums; per-candidate statistics get the same relative bar plus an absolute
# floor because the as-written edge_length delta is pure rounding noise
# (SURVEY.md §7 "Cancellation").
RTOL = 1e-9
ATOL_STAT = 1e-12


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    soa = {k[4:]: z[k] for k in z.files if k.startswith("soa_")}
    rest = {k: z[k] for k in z.files if not k.startswith("soa_")}
    return soa, rest


def inside_cum(soa):
    """cum / cum_he exactly as lite_tess_upload builds them: inside halfedges in
    slot order (faces in index order, each in ccb order from outer_ccb())."""
    lens = soa["he_length"]
    ids, cum = [], []
    c = 0.0
    for f in range(len(soa["face_inside"])):
        if not soa["face_inside"][f]:
            continue
        off, cnt = int(soa["face_he_offset"][f]), int(soa["face_he_count"][f])
        for h in soa["face_he_list"][off:off + cnt]:
            c += float(lens[h])
            cum.append(c)
            ids.append(int(h))
    return np.array(cum), np.array(ids, dtype=np.int32)


def exact_split_face_stats(a, h, x, angle):
    """-(phi(after) - phi(before)) of the face features for Split(h, x, angle), computed as
    the reference computes it: the line (a, b, -w) built in double precision
    (lib/ttessel.cpp:2094-2105) and then taken as EXACT, so that p1, p2 and the two child
    polygons are exact; 60-digit arithmetic stands in for the rationals.  Used to
    adjudicate the rare candidates on which the double-precision CPU port (which rounds p1
    and p2) is itself ill conditioned: a chord of length c turns the 1e-16 rounding of its
    end points into an angle error of 1e-16 / c."""
    import math
    import mpmath as mp
    mp.mp.dps = 60
    f = a["he_face"][h]
    off, cnt = int(a["face_he_offset"][f]), int(a["face_he_count"][f])
    hs = [int(k) for k in a["face_he_list"][off:off + cnt]]
    n = len(hs)
    P = [(mp.mpf(float(a["vx"][a["he_src"][k]])), mp.mpf(float(a["vy"][a["he_src"][k]]))) for k in hs]
    corner = []
    for q in range(n):
        hp = hs[q - 1]
        corner.append(bool(a["he_next_hf"][hp] < 0 or a["he_next_hf"][hp] != a["he_next"][hp]))
    k1 = hs.index(int(h))
    sx, sy = float(P[k1][0]), float(P[k1][1])
    tx, ty = float(P[(k1 + 1) % n][0]), float(P[(k1 + 1) % n][1])
    line_angle = math.atan2(ty - sy, tx - sx) + float(angle)
    da, db = math.sin(line_angle), -math.cos(line_angle)
    u, v = da * sx + db * sy, da * tx + db * ty
    w = u + float(x) * (v - u)
    A, B, W = mp.mpf(da), mp.mpf(db), mp.mpf(w)

    def resid(p):
        return A * p[0] + B * p[1] - W

    def inter(pa, pb):
        sa, sb = resid(pa), resid(pb)
        t = sa / (sa - sb)
        return (pa[0] + t * (pb[0] - pa[0]), pa[1] + t * (pb[1] - pa[1]))

    p1 = inter(P[k1], P[(k1 + 1) % n])
    k2 = None
    for k in range(n):   # convex face: the other halfedge the line crosses
        if k != k1 and resid(P[k]) * resid(P[(k + 1) % n]) < 0:
            k2 = k
    assert k2 is not None
    p2 = inter(P[k2], P[(k2 + 1) % n])

    def ring(first, last):
        out, k = [], first
        while True:
            out.append(k)
            if k == last:
                return out
            k = (k + 1) % n

    def feats(poly, cf):
        m = len(poly)
        area2 = sum(poly[i][0] * poly[(i + 1) % m][1] - poly[(i + 1) % m][0] * poly[i][1] for i in range(m))
        perim = sum(mp.sqrt((poly[(i + 1) % m][0] - poly[i][0]) ** 2 + (poly[(i + 1) % m][1] - poly[i][1]) ** 2)
                    for i in range(m))
        sumang, angles = mp.mpf(0), []
        for i in range(m):
            if not cf[i]:
                continue
            ax, ay = poly[i - 1][0] - poly[i][0], poly[i - 1][1] - poly[i][1]
            bx, by = poly[(i + 1) % m][0] - poly[i][0], poly[(i + 1) % m][1] - poly[i][1]
            ang = mp.acos((ax * bx + ay * by) / mp.sqrt((ax * ax + ay * ay) * (bx * bx + by * by)))
            angles.append(ang)
            if mp.pi / 2 - ang >= 0:
                sumang += mp.pi / 2 - ang
        area = area2 / 2
        minang = min([mp.pi - angles[0]] + angles) if angles else mp.mpf(0)   # lib/ttessel.cpp:4348-4367
        return dict(face_number=mp.mpf(1), face_area_2=area * area, face_perimeter=perim,
                    face_shape=perim / (4 * mp.sqrt(abs(area))), face_sum_of_angles=sumang, min_angle=minang)

    r1, r2 = ring((k2 + 1) % n, k1), ring((k1 + 1) % n, k2)
    c1 = feats([p1, p2] + [P[k] for k in r1], [True, True] + [corner[k] for k in r1])
    c2 = feats([p2, p1] + [P[k] for k in r2], [True, True] + [corner[k] for k in r2])
    # face2poly order: the parent polygon starts at the target of the outer_ccb halfedge
    order = ring(1 % n, 0)
    par = feats([P[k] for k in order], [corner[k] for k in order])
    return {k: -float(c1[k] + c2[k] - par[k]) for k in par}
