// gen_face.h — candidates on faces that are not simple polygons: faces with holes (inner
// boundaries) and faces whose boundary runs along both sides of an edge (an isthmus from the
// outer boundary to a hole, or between two holes).  Windows with holes produce them
// (reference tests/check_lite.cpp:129-229); a rectangular window never does.
//
// The fast kernels (kernels.cuh) assume one simple boundary cycle per face and closed forms
// that follow from it.  This file is the general restatement, polygon by polygon, of
//   ray_exit_face                       lib/ttessel.cpp:4515-4576 (all cycles, twin tie-break)
//   face2poly / boundaries / simplify   :4314-4336, :4453-4506
//   polygon_insert_edge                 :4600-4631
//   hpolygon_insert_edge (cas 1-4)      :4659-4795
//   polygon_remove_edge                 :4945-4966
//   hpolygon_remove_edge (cas 1-4)      :4816-4925
//   the face part of Split/Merge/Flip::modified_elements  :2135-2232, :2302-2404, :2515-2655
//   face features                       :4206-4367
// for the few candidates that touch such a face (k_gen_splits / k_gen_mf in capi.cu run it, one
// candidate per thread, polygons in a per-thread scratch block).  The same source compiles for the
// host: tests/cpp/gen_host.cpp replays it against the exact oracle without a GPU.
//
// Exact-arithmetic notes.  The reference decides "is this polygon vertex collinear with its
// neighbours" (simplify) and "which polygon edge is this" (find_edge_in_polygons) on exact
// rationals.  Here both are decided topologically: a polygon vertex carries the id of the DCEL
// vertex it is (new points get negative ids), a polygon edge lies on the segment of the DCEL
// halfedge joining its end vertices (or on the inserted chord), and a vertex is flat iff the two
// polygon edges around it lie on the same segment.  Coordinates are doubles.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/lite_b200.h"

#if defined(__CUDACC__)
#define GEN_HD __host__ __device__
#else
#define GEN_HD
#endif

// test harness hook: which case of hpolygon_insert_edge / _remove_edge a call took
#ifndef GEN_CASE
#define GEN_CASE(kind, cas)
#endif

namespace gen {

constexpr double kPi = 3.14159265358979323846;
constexpr int kMaxHoles = 24;     // boundaries per face handled (outer + holes)
constexpr int kNew1 = -1;         // vertex id of a split's p1 (when it is not an existing vertex)
constexpr int kNew2 = -2;         // vertex id of p2
constexpr int kChord = -2;        // "segment id" of the inserted edge

struct Prog {  // layout of kernels.cuh::Prog
  int d;
  int fid[LITE_MAX_D];
  uint32_t mask;
};

// raw DCEL (the uploaded lite_tess_soa) + the boundary cycles of the faces
struct Tess {
  int n_he, n_faces;
  const double *vx, *vy;
  const int *he_src, *he_twin, *he_next, *he_face, *he_seg, *he_next_hf, *he_prev_hf;
  const uint8_t *he_dir, *face_inside, *seg_bnd;
  const int *seg_n_edges;
  const int *f_outer;                    // first halfedge of outer_ccb(), -1 for outside faces
  const int *f_hole_off, *f_hole_cnt;    // inner boundaries of a face: hole_he[off .. off+cnt)
  const int *hole_he;                    // one halfedge of each inner boundary
};

struct Pt {
  double x, y;
  int vid;  // DCEL vertex id, kNew1 / kNew2 for the points a move creates
  int he;   // DCEL halfedge the polygon edge LEAVING this vertex lies on, kChord for the inserted edge
};
struct Poly {
  Pt *p;
  int n;
};
struct HPoly {        // outer boundary + holes (pointers into the scratch block)
  Poly b[kMaxHoles];
  int nb;
};
struct Scratch {
  Pt *base;
  int cap, used;
  int overflow;
  GEN_HD Pt *take(int n) {
    if (used + n > cap) { overflow = 1; used = 0; }
    Pt *r = base + used;
    used += n;
    return r;
  }
};

// ---------------------------------------------------------------- DCEL helpers
GEN_HD inline int tgt(const Tess &T, int h) { return T.he_src[T.he_twin[h]]; }
GEN_HD inline int prev_he(const Tess &T, int e) {
  // the halfedge whose next is e: rotate around src(e)
  int o = e;
  for (int guard = 0; guard < 64; guard++) {
    const int i = T.he_twin[o];
    if (T.he_next[i] == e) return i;
    o = T.he_next[i];
  }
  return -1;
}
GEN_HD inline bool on_window_boundary(const Tess &T, int h) { return !T.face_inside[T.he_face[T.he_twin[h]]]; }
GEN_HD inline bool face_is_general(const Tess &T, int f) {
  if (T.f_hole_cnt[f] > 0) return true;
  const int h0 = T.f_outer[f];
  if (h0 < 0) return false;
  int h = h0;
  do {
    if (T.he_face[T.he_twin[h]] == f) return true;
    h = T.he_next[h];
  } while (h != h0);
  return false;
}
// ---------------------------------------------------------------- polygons
GEN_HD inline double poly_area(const Poly &P) {
  if (P.n < 3) return 0.0;
  double acc = 0;
  const double ox = P.p[0].x, oy = P.p[0].y;
  for (int i = 1; i + 1 < P.n; i++)
    acc += (P.p[i].x - ox) * (P.p[i + 1].y - oy) - (P.p[i].y - oy) * (P.p[i + 1].x - ox);
  return 0.5 * acc;
}
GEN_HD inline double poly_perimeter(const Poly &P) {
  double l = 0;
  for (int i = 0; i < P.n; i++) {
    const Pt &a = P.p[i], &b = P.p[(i + 1 == P.n) ? 0 : i + 1];
    l += sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y));
  }
  return l;
}
// CGAL::bounded_side_2 for a point that is not on the boundary: +1 inside, -1 outside
GEN_HD inline int bounded_side(const Poly &P, double x, double y) {
  bool in = false;
  for (int i = 0, j = P.n - 1; i < P.n; j = i++) {
    const Pt &a = P.p[i], &b = P.p[j];
    if ((a.y > y) != (b.y > y)) {
      const double xc = a.x + (y - a.y) * (b.x - a.x) / (b.y - a.y);
      if (x < xc) in = !in;
    }
  }
  return in ? 1 : -1;
}

// one boundary cycle as face2poly(f, false) lists it (:4321-4330): the targets of its
// halfedges, starting with the target of h0.  The polygon edge of halfedge number k of
// the cycle runs from vertex k-1 to vertex k.
GEN_HD inline Poly cycle_poly(const Tess &T, Scratch &S, int h0) {
  int n = 0, h = h0;
  do { n++; h = T.he_next[h]; } while (h != h0 && n < (1 << 20));
  Poly P;
  P.p = S.take(n); P.n = n;
  h = h0;
  for (int k = 0; k < n; k++) {
    const int v = tgt(T, h);
    h = T.he_next[h];
    P.p[k].x = T.vx[v]; P.p[k].y = T.vy[v]; P.p[k].vid = v; P.p[k].he = h;
  }
  return P;
}
// boundaries(face2poly(f, false)): outer cycle then holes
GEN_HD inline void face_boundaries(const Tess &T, Scratch &S, int f, HPoly &H) {
  H.nb = 0;
  H.b[H.nb++] = cycle_poly(T, S, T.f_outer[f]);
  for (int k = 0; k < T.f_hole_cnt[f] && H.nb < kMaxHoles; k++)
    H.b[H.nb++] = cycle_poly(T, S, T.hole_he[T.f_hole_off[f] + k]);
  if (T.f_hole_cnt[f] + 1 > kMaxHoles) S.overflow = 1;
}

struct PE {  // polygon edge circulator: the edge from P->p[i] to P->p[i+1]
  const Poly *P;
  int i;
};
GEN_HD inline PE pe_inc(PE e) { e.i = (e.i + 1 == e.P->n) ? 0 : e.i + 1; return e; }
GEN_HD inline bool pe_same(const PE &a, const PE &b) { return a.P->p == b.P->p && a.i == b.i; }
GEN_HD inline const Pt &pe_src(const PE &e) { return e.P->p[e.i]; }
GEN_HD inline const Pt &pe_tgt(const PE &e) { return e.P->p[(e.i + 1 == e.P->n) ? 0 : e.i + 1]; }

// find_edge_in_polygons (:5046-5064) by vertex ids; returns the boundary index, H.nb if absent
GEN_HD inline int find_edge(const HPoly &H, int va, int vb, PE &out) {
  for (int i = 0; i < H.nb; i++) {
    const Poly &P = H.b[i];
    for (int k = 0; k < P.n; k++)
      if (P.p[k].vid == va && P.p[(k + 1 == P.n) ? 0 : k + 1].vid == vb) {
        out.P = &H.b[i]; out.i = k;
        return i;
      }
  }
  return H.nb;
}

// polygon_insert_edge (:4600-4631).  p1 lies on e1, p2 on e2.  Every vertex copied from a source
// polygon keeps the halfedge of its outgoing edge, which stays right because its successor
// is copied with it; the new points get theirs here (the chord, or the rest of e1 / e2).
GEN_HD inline Poly polygon_insert_edge(Scratch &S, PE e1, PE e2, const Pt &p1, const Pt &p2) {
  const int room = e1.P->n + e2.P->n + 6;
  Poly R;
  R.p = S.take(room); R.n = 0;
  PE e = e2;
  const bool p2_is_tgt = pe_tgt(e2).vid == p2.vid, p1_is_tgt = pe_tgt(e1).vid == p1.vid;
  R.p[R.n] = p1; R.p[R.n++].he = kChord;
  R.p[R.n] = p2; R.p[R.n++].he = p2_is_tgt ? pe_tgt(e2).he : pe_src(e2).he;
  if (p2_is_tgt) e = pe_inc(e);
  for (;;) {
    R.p[R.n++] = pe_tgt(e);
    e = pe_inc(e);
    if (pe_same(e, e2) || pe_same(e, e1) || R.n >= room - 1) break;
  }
  if (pe_same(e, e1)) {
    if (pe_src(e1).vid == p1.vid) R.n--;
    return R;
  }
  R.p[R.n] = p2; R.p[R.n++].he = kChord;
  R.p[R.n] = p1; R.p[R.n++].he = p1_is_tgt ? pe_tgt(e1).he : pe_src(e1).he;
  e = e1;
  if (p1_is_tgt) e = pe_inc(e);
  for (;;) {
    R.p[R.n++] = pe_tgt(e);
    e = pe_inc(e);
    if (pe_same(e, e1) || R.n >= room) break;
  }
  if (pe_src(e1).vid == p1.vid) R.n--;
  return R;
}

// polygon_remove_edge (:4945-4966).  e2 is the opposite of e1.  At the seams the copied vertex
// is followed by a vertex of the other side: its outgoing edge is the one after the removed edge.
GEN_HD inline Poly polygon_remove_edge(Scratch &S, PE e1, PE e2) {
  const int room = e1.P->n + e2.P->n + 2;
  Poly R;
  R.p = S.take(room); R.n = 0;
  const int after_e1 = pe_tgt(e1).he, after_e2 = pe_tgt(e2).he;  // halfedges of e1+1 and e2+1
  PE c = pe_inc(e1);
  while (!pe_same(c, e1) && !pe_same(c, e2) && R.n < room) { R.p[R.n++] = pe_tgt(c); c = pe_inc(c); }
  if (pe_same(c, e2)) {           // one polygon visiting the edge twice: the part in front of e1
    if (R.n) R.p[R.n - 1].he = after_e1;   // src(e2) = src(e1+1) closes onto tgt(e1+1)
    return R;
  }
  if (R.n) R.p[R.n - 1].he = after_e2;     // src(e1) = src(e2+1) goes on along e2's polygon
  c = pe_inc(e2);
  while (!pe_same(c, e2) && R.n < room) { R.p[R.n++] = pe_tgt(c); c = pe_inc(c); }
  if (R.n) R.p[R.n - 1].he = after_e1;     // src(e2) = src(e1+1) closes onto tgt(e1+1)
  return R;
}

// hpolygon_insert_edge (:4659-4795): fills res[0..1], returns how many faces come out
GEN_HD inline int hpolygon_insert_edge(Scratch &S, const HPoly &hp, int i1, int i2, PE e1, PE e2, const Pt &p1,
                                       const Pt &p2, HPoly *res) {
  int cas;
  if (hp.nb == 1) cas = 1;
  else if (i1 == i2) cas = (i1 == 0) ? 1 : 2;
  else if (i1 == 0 || i2 == 0) cas = 3;
  else cas = 4;
  GEN_CASE(0, cas);
  if (cas == 1) {
    res[0].nb = res[1].nb = 0;
    res[0].b[res[0].nb++] = polygon_insert_edge(S, e1, e2, p1, p2);
    res[1].b[res[1].nb++] = polygon_insert_edge(S, e2, e1, p2, p1);
    for (int ii = 1; ii < hp.nb; ii++) {  // filter_holes (:4996-5009): first vertex strictly inside outer1
      const bool in1 = bounded_side(res[0].b[0], hp.b[ii].p[0].x, hp.b[ii].p[0].y) > 0;
      HPoly &dst = in1 ? res[0] : res[1];
      dst.b[dst.nb++] = hp.b[ii];
    }
    return 2;
  }
  if (cas == 2) {
    // mother face (the old face minus the daughter) and daughter face (along the hole)
    Poly c21 = polygon_insert_edge(S, e2, e1, p2, p1), outer2, enlarged;
    const Pt &t2 = pe_tgt(e2);
    // bounded_side_2(c21, target of e2) != ON_UNBOUNDED_SIDE (:4737).  The target can lie ON c21: a
    // hole made of two holes joined by an isthmus visits the isthmus' end vertices twice.  A DCEL
    // vertex is on the boundary of c21 iff it is one of its vertices.
    bool not_outside = false;
    for (int k = 0; k < c21.n && !not_outside; k++) not_outside = c21.p[k].vid == t2.vid;
    if (!not_outside) not_outside = bounded_side(c21, t2.x, t2.y) > 0;
    if (not_outside) { outer2 = polygon_insert_edge(S, e1, e2, p1, p2); enlarged = c21; }
    else { outer2 = c21; enlarged = polygon_insert_edge(S, e1, e2, p1, p2); }
    res[0].nb = res[1].nb = 0;
    res[0].b[res[0].nb++] = hp.b[0];
    res[1].b[res[1].nb++] = outer2;
    for (int ii = 1; ii < hp.nb; ii++) {
      if (ii == i1) res[0].b[res[0].nb++] = enlarged;
      else if (bounded_side(outer2, hp.b[ii].p[0].x, hp.b[ii].p[0].y) > 0) res[1].b[res[1].nb++] = hp.b[ii];
      else res[0].b[res[0].nb++] = hp.b[ii];
    }
    return 2;
  }
  if (cas == 3) {
    // the outer boundary folds into the face and takes the hole in
    res[0].nb = 0;
    res[0].b[res[0].nb++] = polygon_insert_edge(S, e1, e2, p1, p2);
    const int j = i1 > 0 ? i1 : i2;
    for (int ii = 1; ii < hp.nb; ii++)
      if (ii != j) res[0].b[res[0].nb++] = hp.b[ii];
    return 1;
  }
  // cas 4: two holes become one
  res[0].nb = 0;
  res[0].b[res[0].nb++] = hp.b[0];
  for (int ii = 1; ii < hp.nb; ii++)
    if (ii != i1 && ii != i2) res[0].b[res[0].nb++] = hp.b[ii];
  res[0].b[res[0].nb++] = polygon_insert_edge(S, e1, e2, p1, p2);
  return 1;
}

// hpolygon_remove_edge (:4816-4925)
GEN_HD inline void hpolygon_remove_edge(Scratch &S, const HPoly &hp, const HPoly &hpt, int i, int it, PE e, PE et,
                                        bool single, HPoly &out) {
  int cas;
  if (!single && hp.nb == 1 && hpt.nb == 1) cas = 1;
  else if (single && i == it) cas = (i == 0) ? 3 : 4;
  else if (i > 0) cas = 2;
  else cas = (it == 0) ? 1 : 2;
  GEN_CASE(1, cas);
  out.nb = 0;
  if (cas == 1) {
    out.b[out.nb++] = polygon_remove_edge(S, e, et);
    for (int k = 1; k < hp.nb && out.nb < kMaxHoles; k++) out.b[out.nb++] = hp.b[k];
    for (int k = 1; k < hpt.nb && out.nb < kMaxHoles; k++) out.b[out.nb++] = hpt.b[k];
    return;
  }
  if (cas == 2) {
    // the edge joins two vertices of one hole of the mother face: the daughter on its other side dissolves
    const HPoly &mother = (i > 0) ? hp : hpt, &daughter = (i > 0) ? hpt : hp;
    const int j = (i > 0) ? i : it;
    out.b[out.nb++] = mother.b[0];
    for (int k = 1; k < daughter.nb && out.nb < kMaxHoles; k++) out.b[out.nb++] = daughter.b[k];
    for (int k = 1; k < mother.nb && out.nb < kMaxHoles; k++) {
      if (k != j) out.b[out.nb++] = mother.b[k];
      else out.b[out.nb++] = (i > 0) ? polygon_remove_edge(S, e, et) : polygon_remove_edge(S, et, e);
    }
    return;
  }
  if (cas == 3) {
    // an isthmus from the outer boundary to what becomes a hole again
    Poly o = polygon_remove_edge(S, e, et), in = polygon_remove_edge(S, et, e);
    if (poly_area(o) <= 0) { Poly t = o; o = in; in = t; }
    out.b[out.nb++] = o;
    for (int k = 1; k < hp.nb && out.nb < kMaxHoles - 1; k++) out.b[out.nb++] = hp.b[k];
    out.b[out.nb++] = in;
    return;
  }
  // cas 4: an isthmus between two holes: the merged hole divides again
  out.b[out.nb++] = hp.b[0];
  for (int k = 1; k < hp.nb && out.nb < kMaxHoles - 2; k++)
    if (k != i) out.b[out.nb++] = hp.b[k];
  out.b[out.nb++] = polygon_remove_edge(S, e, et);
  out.b[out.nb++] = polygon_remove_edge(S, et, e);
}

// ---------------------------------------------------------------- face features of a holed polygon
struct FaceFeat {
  double area2, perim, sumang, minang;  // face_area_2 (:4215), face_perimeter (:4229), :4285, :4348
};
GEN_HD inline int seg_of(const Tess &T, int he) { return he < 0 ? he : T.he_seg[he]; }
GEN_HD inline double angle_between(double ax, double ay, double bx, double by) {  // :4255-4268
  const double n1 = ax * ax + ay * ay, n2 = bx * bx + by * by, p = ax * bx + ay * by;
  double c = p / sqrt(n1 * n2);
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  return acos(c);
}
// Features of simplify_h(hp) (:4495-4506): every boundary contributes its squared area and its
// length; the angles are those of the simplified OUTER boundary, in the order simplify leaves
// it (scan from vertex 0 of the polygon as built, after boundaries() has turned a clockwise
// outer boundary round keeping its first vertex).
GEN_HD inline FaceFeat hp_features(const Tess &T, Scratch &S, const HPoly &hp, bool need_ang) {
  FaceFeat F;
  F.area2 = 0; F.perim = 0; F.sumang = 0; F.minang = 0;
  double a0 = 0;
  for (int k = 0; k < hp.nb; k++) {
    const double a = poly_area(hp.b[k]);
    if (k == 0) a0 = a;
    F.area2 += a * a;
    F.perim += poly_perimeter(hp.b[k]);
  }
  if (!need_ang) return F;
  Poly ob = hp.b[0];
  const int n = ob.n;
  if (n < 3) return F;
  if (a0 < 0) {  // reverse_orientation keeps vertex 0 (:4458-4460); edge i of the result is old edge n-1-i reversed
    Poly r;
    r.p = S.take(n); r.n = n;
    for (int i = 0; i < n; i++) {
      const int src = (i == 0) ? 0 : n - i;
      r.p[i] = ob.p[src];
      r.p[i].he = ob.p[n - 1 - i].he;
    }
    ob = r;
  }
  // corners: the two polygon edges around the vertex lie on different segments
  int first = -1;
  for (int i = 0; i < n && first < 0; i++)
    if (seg_of(T, ob.p[(i == 0) ? n - 1 : i - 1].he) != seg_of(T, ob.p[i].he)) first = i;
  if (first < 0) return F;
  int prev = first;  // the corner before `first` (going backwards from it)
  for (int k = 1; k <= n; k++) {
    const int i = (first - k + n) % n;
    if (seg_of(T, ob.p[(i == 0) ? n - 1 : i - 1].he) != seg_of(T, ob.p[i].he)) { prev = i; break; }
  }
  double minang = 1e300, seed = -1;
  int cur = first, guard = 0;
  do {
    int nxt = cur;
    for (int k = 1; k <= n; k++) {
      const int i = (cur + k) % n;
      if (seg_of(T, ob.p[(i == 0) ? n - 1 : i - 1].he) != seg_of(T, ob.p[i].he)) { nxt = i; break; }
    }
    const Pt &c = ob.p[cur], &a = ob.p[prev], &b = ob.p[nxt];
    const double ang = angle_between(a.x - c.x, a.y - c.y, b.x - c.x, b.y - c.y);
    if (seed < 0) seed = angle_between(c.x - a.x, c.y - a.y, b.x - c.x, b.y - c.y);  // turning angle at corner 0 (:4354)
    const double q = kPi / 2 - ang;
    if (q >= 0) F.sumang += q;
    if (ang < minang) minang = ang;
    prev = cur; cur = nxt;
  } while (cur != first && ++guard < n + 2);
  F.minang = seed < minang ? seed : minang;
  return F;
}
GEN_HD inline void add_face_feats(const Prog &G, double *acc, double sign, const FaceFeat &F) {
  for (int i = 0; i < G.d; i++) {
    switch (G.fid[i]) {
      case LITE_FEAT_FACE_NUMBER: acc[i] += sign; break;
      case LITE_FEAT_FACE_AREA_2: acc[i] += sign * F.area2; break;
      case LITE_FEAT_FACE_PERIMETER: acc[i] += sign * F.perim; break;
      case LITE_FEAT_FACE_SHAPE: acc[i] += sign * (F.perim / (4.0 * sqrt(sqrt(F.area2)))); break;  // :4247-4249
      case LITE_FEAT_FACE_SUM_OF_ANGLES: acc[i] += sign * F.sumang; break;
      case LITE_FEAT_MIN_ANGLE: acc[i] += sign * F.minang; break;
      default: break;
    }
  }
}
GEN_HD inline bool needs_angles(const Prog &G) {
  return G.mask & ((1u << LITE_FEAT_FACE_SUM_OF_ANGLES) | (1u << LITE_FEAT_MIN_ANGLE));
}
GEN_HD inline bool needs_faces(const Prog &G) {
  return G.mask & ((1u << LITE_FEAT_FACE_NUMBER) | (1u << LITE_FEAT_FACE_AREA_2) | (1u << LITE_FEAT_FACE_PERIMETER) |
                   (1u << LITE_FEAT_FACE_SHAPE) | (1u << LITE_FEAT_FACE_SUM_OF_ANGLES) | (1u << LITE_FEAT_MIN_ANGLE));
}
// features of an unmodified face (a deleted face of a move): simplify_h(face2poly(f))
GEN_HD inline FaceFeat face_features(const Tess &T, Scratch &S, int f, bool need_ang) {
  const int mark = S.used;
  HPoly H;
  face_boundaries(T, S, f, H);
  const FaceFeat F = hp_features(T, S, H, need_ang);
  S.used = mark;
  return F;
}

// ---------------------------------------------------------------- ray_exit_face (:4515-4576)
// Ray from (px,py) along (dx,dy) inside face f; line a*X + b*Y = w through the source.
// skip_he / skip_twin: the halfedge the source lies on (distance 0 in exact arithmetic, :4552);
// skip_v: a DCEL vertex the source coincides with (every edge incident to it is at distance 0).
GEN_HD inline int ray_exit_general(const Tess &T, int f, double a, double b, double w, double px, double py,
                                   double dx, double dy, int skip_he, int skip_v, double &hx, double &hy) {
  double best = 1.7976931348623157e308;
  int e = -1;
  const int ncyc = 1 + T.f_hole_cnt[f];
  for (int c = 0; c < ncyc; c++) {
    const int h0 = (c == 0) ? T.f_outer[f] : T.hole_he[T.f_hole_off[f] + c - 1];
    int h = h0;
    do {
      const int vs = T.he_src[h], vt = tgt(T, h);
      const bool skip = (h == skip_he) || (skip_he >= 0 && h == T.he_twin[skip_he]) || vs == skip_v || vt == skip_v;
      if (!skip) {
        const double sx = T.vx[vs], sy = T.vy[vs], tx = T.vx[vt], ty = T.vy[vt];
        const double sk = fma(a, sx, fma(b, sy, -w)), sn = fma(a, tx, fma(b, ty, -w));
        const bool cr = (sk <= 0.0 && sn >= 0.0) || (sk >= 0.0 && sn <= 0.0);
        if (cr && sk != sn) {
          const double sg = sk / (sk - sn);
          const double X = fma(sg, tx - sx, sx), Y = fma(sg, ty - sy, sy);
          const double rx = X - px, ry = Y - py;
          if (rx * dx + ry * dy >= 0.0) {
            const double len = sqrt(rx * rx + ry * ry);
            if (len != 0.0) {
              if (e >= 0 && h == T.he_twin[e]) {
                // both an halfedge and its twin bound the face (:4554-4562): keep the one that has
                // the ray source on its left
                if ((tx - sx) * (py - sy) - (ty - sy) * (px - sx) > 0.0) e = h;
              } else if (len < best) {
                best = len; e = h; hx = X; hy = Y;
              }
            }
          }
        }
      }
      h = T.he_next[h];
    } while (h != h0);
  }
  return e;
}

// ---------------------------------------------------------------- the three moves
struct SplitRes {
  int e2;
  double p1x, p1y, p2x, p2y;
};
GEN_HD inline double edge_cut(double ax, double ay, double mx, double my, double bx, double by) {
  return sqrt((mx - ax) * (mx - ax) + (my - ay) * (my - ay)) + sqrt((bx - mx) * (bx - mx) + (by - my) * (by - my)) -
         sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
}
GEN_HD inline double dist(double ax, double ay, double bx, double by) {
  return sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
}

// Split(e1, x, angle): constructor (:2090-2132) + statistic variation of modified_elements (:2135-2232).
// stat[k * stride] receives MINUS the variation of feature k, as PseudoLikDiscrete stores it.
GEN_HD inline void split_general(const Tess &T, const Prog &G, Scratch &S, int h1, double x, double angle,
                                 double *stat, size_t stride, SplitRes &o) {
  S.used = 0;
  const int vs = T.he_src[h1], vt = tgt(T, h1);
  const double sx = T.vx[vs], sy = T.vy[vs], tx = T.vx[vt], ty = T.vy[vt];
  const double l = atan2(ty - sy, tx - sx) + angle;
  const double a = sin(l), b = -cos(l);
  const double u = a * sx + b * sy, v = a * tx + b * ty;
  const double w = u + x * (v - u);
  const double sP = fma(a, sx, fma(b, sy, -w)), sQ = fma(a, tx, fma(b, ty, -w));
  const double t1 = sP / (sP - sQ);
  // x == 0: w == u and the reference's exact intersection is the source vertex itself
  const double p1x = (x == 0.0) ? sx : fma(t1, tx - sx, sx), p1y = (x == 0.0) ? sy : fma(t1, ty - sy, sy);
  double dx, dy;
  if (sQ < 0.0) { dx = b; dy = -a; } else { dx = -b; dy = a; }
  const int f = T.he_face[h1];
  double p2x = 0, p2y = 0;
  const int h2 = (sP != sQ) ? ray_exit_general(T, f, a, b, w, p1x, p1y, dx, dy, h1, (x == 0.0) ? vs : -1, p2x, p2y) : -1;
  o.e2 = h2; o.p1x = p1x; o.p1y = p1y; o.p2x = p2x; o.p2y = p2y;
  if (h2 < 0) {
    for (int k = 0; k < G.d; k++) stat[(size_t)k * stride] = 0.0;
    return;
  }
  const bool bnd1 = on_window_boundary(T, h1), bnd2 = on_window_boundary(T, h2);
  const int v2s = T.he_src[h2], v2t = tgt(T, h2);
  const double d01 = dist(p1x, p1y, p2x, p2y);
  double acc[LITE_MAX_D];
  for (int k = 0; k < G.d; k++) acc[k] = 0.0;
  if (needs_faces(G)) {
    const bool ang = needs_angles(G);
    HPoly fb, res[2];
    face_boundaries(T, S, f, fb);
    add_face_feats(G, acc, -1.0, hp_features(T, S, fb, ang));
    PE pe1{nullptr, 0}, pe2{nullptr, 0};
    const int i1 = find_edge(fb, vs, vt, pe1), i2 = find_edge(fb, v2s, v2t, pe2);
    Pt p1, p2;
    p1.x = p1x; p1.y = p1y; p1.vid = (x == 0.0) ? vs : kNew1; p1.he = kChord;
    p2.x = p2x; p2.y = p2y; p2.vid = kNew2; p2.he = kChord;
    if (i1 < fb.nb && i2 < fb.nb) {
      const int nr = hpolygon_insert_edge(S, fb, i1, i2, pe1, pe2, p1, p2, res);
      for (int r = 0; r < nr; r++) add_face_feats(G, acc, +1.0, hp_features(T, S, res[r], ang));
    } else S.overflow = 1;
  }
  for (int k = 0; k < G.d; k++) {
    double dv = acc[k];
    switch (G.fid[k]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW: {
        bool in1 = !bnd1;
        if (x == 0.0) in1 = !bnd1 && !on_window_boundary(T, prev_he(T, h1));
        dv = (in1 ? 1.0 : 0.0) + (bnd2 ? 0.0 : 1.0);
      } break;
      case LITE_FEAT_EDGE_LENGTH:
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: {
        const bool wantb = G.fid[k] == LITE_FEAT_EDGE_LENGTH;
        dv = wantb ? 0.0 : d01;
        if (bnd1 == wantb) dv += edge_cut(sx, sy, p1x, p1y, tx, ty);
        if (bnd2 == wantb) dv += edge_cut(T.vx[v2s], T.vy[v2s], p2x, p2y, T.vx[v2t], T.vy[v2t]);
      } break;
      case LITE_FEAT_SEG_NUMBER: dv = 1.0; break;
      case LITE_FEAT_IS_SEGMENT_INTERNAL: dv = 1.0; break;
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: dv = -1.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2:
        dv = 0.0;
        if (bnd1) dv += 2.0 * T.seg_n_edges[T.he_seg[h1]] + 1.0;
        if (bnd2) dv += 2.0 * T.seg_n_edges[T.he_seg[h2]] + 1.0;
        break;
      default: break;
    }
    stat[(size_t)k * stride] = -dv;
  }
}

// Merge(e) (:2299-2404), e = the single halfedge (dir == true) of a non-blocking segment
GEN_HD inline void merge_general(const Tess &T, const Prog &G, Scratch &S, int e, double *stat, size_t stride) {
  S.used = 0;
  const int t = T.he_twin[e];
  const int e2 = T.he_next[e], e1 = T.he_next[t];  // Merge::e2(), e1() (include/ttessel.h:657-661)
  const int v1 = T.he_src[e], v2 = T.he_src[t];    // p1, p2
  const double p1x = T.vx[v1], p1y = T.vy[v1], p2x = T.vx[v2], p2y = T.vy[v2];
  const bool bnd_e2 = on_window_boundary(T, e2), bnd_e1 = on_window_boundary(T, e1);
  const int x2 = tgt(T, e2), y2 = T.he_src[T.he_prev_hf[e2]], x1 = tgt(T, e1), y1 = T.he_src[T.he_prev_hf[e1]];
  const double elen = dist(p1x, p1y, p2x, p2y);
  double acc[LITE_MAX_D];
  for (int k = 0; k < G.d; k++) acc[k] = 0.0;
  if (needs_faces(G)) {
    const bool ang = needs_angles(G);
    const int f = T.he_face[e], ft = T.he_face[t];
    HPoly fb, ftb, out;
    face_boundaries(T, S, f, fb);
    add_face_feats(G, acc, -1.0, hp_features(T, S, fb, ang));
    PE pe{nullptr, 0}, pet{nullptr, 0};
    const int i = find_edge(fb, v1, v2, pe);
    if (f == ft) {  // both sides of e are the same face (:2343-2352)
      const int it = find_edge(fb, v2, v1, pet);
      if (i < fb.nb && it < fb.nb) {
        hpolygon_remove_edge(S, fb, fb, i, it, pe, pet, true, out);
        add_face_feats(G, acc, +1.0, hp_features(T, S, out, ang));
      } else S.overflow = 1;
    } else {
      face_boundaries(T, S, ft, ftb);
      add_face_feats(G, acc, -1.0, hp_features(T, S, ftb, ang));
      const int it = find_edge(ftb, v2, v1, pet);
      if (i < fb.nb && it < ftb.nb) {
        hpolygon_remove_edge(S, fb, ftb, i, it, pe, pet, false, out);
        add_face_feats(G, acc, +1.0, hp_features(T, S, out, ang));
      } else S.overflow = 1;
    }
  }
  for (int k = 0; k < G.d; k++) {
    double dv = acc[k];
    switch (G.fid[k]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW: dv = -((bnd_e1 ? 0.0 : 1.0) + (bnd_e2 ? 0.0 : 1.0)); break;
      case LITE_FEAT_EDGE_LENGTH:
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: {
        const bool wantb = G.fid[k] == LITE_FEAT_EDGE_LENGTH;
        dv = wantb ? 0.0 : -elen;
        if (bnd_e2 == wantb)
          dv += dist(T.vx[y2], T.vy[y2], T.vx[x2], T.vy[x2]) - dist(p2x, p2y, T.vx[x2], T.vy[x2]) - dist(T.vx[y2], T.vy[y2], p2x, p2y);
        if (bnd_e1 == wantb)
          dv += dist(T.vx[y1], T.vy[y1], T.vx[x1], T.vy[x1]) - dist(p1x, p1y, T.vx[x1], T.vy[x1]) - dist(T.vx[y1], T.vy[y1], p1x, p1y);
      } break;
      case LITE_FEAT_SEG_NUMBER: dv = -1.0; break;
      case LITE_FEAT_IS_SEGMENT_INTERNAL: dv = -1.0; break;
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: dv = 1.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2:
        dv = 0.0;
        if (bnd_e1) dv -= 2.0 * T.seg_n_edges[T.he_seg[e1]] - 1.0;
        if (bnd_e2) dv -= 2.0 * T.seg_n_edges[T.he_seg[e2]] - 1.0;
        break;
      default: break;
    }
    stat[(size_t)k * stride] = -dv;
  }
}

struct FlipRes {
  int e2, right;
  double p2x, p2y;
};
// Flip(e1) (:2432-2454) and the statistic variation of its modified_elements (:2459-2713)
GEN_HD inline void flip_general(const Tess &T, const Prog &G, Scratch &S, int e1, double *stat, size_t stride, FlipRes &o) {
  S.used = 0;
  const int t1 = T.he_twin[e1];
  const int va = T.he_src[e1], vq = T.he_src[t1];
  const double ax = T.vx[va], ay = T.vy[va], qx = T.vx[vq], qy = T.vy[vq];
  // e3: the halfedge arriving at a that is not on e1's segment (:2435-2438)
  const int pe1h = prev_he(T, e1);
  const bool right = (pe1h != T.he_prev_hf[e1]);  // prev(e1) is e3: it arrives from the left of e1
  int s3;                                         // source vertex of e3
  if (right) s3 = T.he_src[pe1h];
  else s3 = tgt(T, T.he_next[t1]);                // e3 = twin(next(twin e1))
  double dx = ax - T.vx[s3], dy = ay - T.vy[s3];
  { const double nrm = sqrt(dx * dx + dy * dy); dx /= nrm; dy /= nrm; }
  const double la = dy, lb = -dx, lw = la * ax + lb * ay;
  const int fR = right ? T.he_face[t1] : T.he_face[e1];
  double p2x = 0, p2y = 0;
  const int e2 = ray_exit_general(T, fR, la, lb, lw, ax, ay, dx, dy, -1, va, p2x, p2y);
  o.e2 = e2; o.right = right ? 1 : 0; o.p2x = p2x; o.p2y = p2y;
  if (e2 < 0) {
    for (int k = 0; k < G.d; k++) stat[(size_t)k * stride] = 0.0;
    return;
  }
  const int n1 = T.he_next[e1];                     // leaves q along the blocking segment
  const bool bnd_n1 = on_window_boundary(T, n1), bnd_e2 = on_window_boundary(T, e2);
  const int vX = tgt(T, n1), vY = T.he_src[T.he_prev_hf[n1]];
  const int v2s = T.he_src[e2], v2t = tgt(T, e2);
  const bool tgt_is_q = v2t == vq, src_is_q = v2s == vq;
  const double l_e1 = dist(ax, ay, qx, qy), l_new = dist(ax, ay, p2x, p2y);
  double acc[LITE_MAX_D];
  for (int k = 0; k < G.d; k++) acc[k] = 0.0;
  if (needs_faces(G)) {
    const bool ang = needs_angles(G);
    // faces (:2515-2655): insert (a,p2) into the face of e2, then remove e1 from whatever it bounds now
    HPoly f2b, ins[2], other, other_t, rem;
    face_boundaries(T, S, fR, f2b);
    PE pe2{nullptr, 0}, pep1{nullptr, 0};
    const int i2 = find_edge(f2b, v2s, v2t, pe2);
    const int i1 = right ? find_edge(f2b, vq, va, pep1) : find_edge(f2b, va, vq, pep1);
    Pt pt1, p2;
    pt1.x = ax; pt1.y = ay; pt1.vid = va; pt1.he = kChord;
    p2.x = p2x; p2.y = p2y; p2.vid = kNew2; p2.he = kChord;
    if (i1 >= f2b.nb || i2 >= f2b.nb) { S.overflow = 1; }
    else {
      const int nins = hpolygon_insert_edge(S, f2b, i1, i2, pep1, pe2, pt1, p2, ins);
      bool match[2] = {false, false};
      bool del_f1 = false, del_f1t = false;
      // e1 = (a, q) among the inserted faces, else in its own (untouched) face
      const HPoly *poly = nullptr, *poly_twin = nullptr;
      PE pe1{nullptr, 0}, pe1t{nullptr, 0};
      int j1 = 0, j1t = 0;
      bool found = false;
      if (nins == 2) {
        j1 = find_edge(ins[1], va, vq, pe1);
        if (j1 < ins[1].nb) { found = true; match[1] = true; poly = &ins[1]; }
      }
      if (!found) {
        j1 = find_edge(ins[0], va, vq, pe1);
        if (j1 < ins[0].nb) { found = true; match[0] = true; poly = &ins[0]; }
      }
      if (!found) {
        face_boundaries(T, S, T.he_face[e1], other);
        j1 = find_edge(other, va, vq, pe1);
        if (j1 < other.nb) { found = true; del_f1 = true; poly = &other; }
      }
      bool ok = found;
      // its twin (q, a)
      bool twin_in_poly = false;
      found = false;
      if (ok) {
        if (match[0]) {
          j1t = find_edge(*poly, vq, va, pe1t);
          if (j1t < poly->nb) { found = true; twin_in_poly = true; }
        } else {
          j1t = find_edge(ins[0], vq, va, pe1t);
          if (j1t < ins[0].nb) { found = true; match[0] = true; poly_twin = &ins[0]; }
        }
        if (!found && nins == 2) {
          if (match[1]) {
            j1t = find_edge(*poly, vq, va, pe1t);
            if (j1t < poly->nb) { found = true; twin_in_poly = true; }
          } else {
            j1t = find_edge(ins[1], vq, va, pe1t);
            if (j1t < ins[1].nb) { found = true; match[1] = true; poly_twin = &ins[1]; }
          }
        }
        if (!found) {
          face_boundaries(T, S, T.he_face[t1], other_t);
          j1t = find_edge(other_t, vq, va, pe1t);
          if (j1t < other_t.nb && !del_f1) { found = true; del_f1t = true; poly_twin = &other_t; }
        }
        ok = found;
      }
      if (!ok) S.overflow = 1;
      else {
        if (twin_in_poly) hpolygon_remove_edge(S, *poly, *poly, j1, j1t, pe1, pe1t, true, rem);
        else hpolygon_remove_edge(S, *poly, *poly_twin, j1, j1t, pe1, pe1t, false, rem);
        add_face_feats(G, acc, -1.0, hp_features(T, S, f2b, ang));
        if (del_f1) add_face_feats(G, acc, -1.0, hp_features(T, S, *poly, ang));
        if (del_f1t) add_face_feats(G, acc, -1.0, hp_features(T, S, *poly_twin, ang));
        for (int k = 0; k < nins; k++)
          if (!match[k]) add_face_feats(G, acc, +1.0, hp_features(T, S, ins[k], ang));
        add_face_feats(G, acc, +1.0, hp_features(T, S, rem, ang));
      }
    }
  }
  for (int k = 0; k < G.d; k++) {
    double dv = acc[k];
    switch (G.fid[k]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW: dv = (bnd_e2 ? 0.0 : 1.0) - (bnd_n1 ? 0.0 : 1.0); break;
      case LITE_FEAT_EDGE_LENGTH:
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: {
        const bool wantb = G.fid[k] == LITE_FEAT_EDGE_LENGTH;
        const double Xx = T.vx[vX], Xy = T.vy[vX], Yx = T.vx[vY], Yy = T.vy[vY];
        const double Sx = T.vx[v2s], Sy = T.vy[v2s], Ex = T.vx[v2t], Ey = T.vy[v2t];
        dv = wantb ? 0.0 : l_new - l_e1;
        if (bnd_n1 == wantb) dv -= dist(qx, qy, Xx, Xy) + dist(qx, qy, Yx, Yy);
        if (!tgt_is_q && !src_is_q) {
          if (bnd_e2 == wantb) dv += dist(p2x, p2y, Ex, Ey) + dist(p2x, p2y, Sx, Sy) - dist(Ex, Ey, Sx, Sy);
          if (bnd_n1 == wantb) dv += dist(Xx, Xy, Yx, Yy);
        } else if (tgt_is_q) {
          if (bnd_e2 == wantb) dv += dist(p2x, p2y, Sx, Sy) + dist(p2x, p2y, Xx, Xy);
        } else {
          if (bnd_e2 == wantb) dv += dist(p2x, p2y, Ex, Ey) + dist(p2x, p2y, Yx, Yy);
        }
      } break;
      case LITE_FEAT_SEG_NUMBER:
      case LITE_FEAT_IS_SEGMENT_INTERNAL:
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: dv = 0.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2: {
        dv = 0.0;
        const int sg2 = T.he_seg[e2], sgn = T.he_seg[n1];
        if (sg2 != sgn) {
          if (bnd_e2) dv += 2.0 * T.seg_n_edges[sg2] + 1.0;
          if (bnd_n1) dv -= 2.0 * T.seg_n_edges[sgn] - 1.0;
        }
      } break;
      default: break;
    }
    stat[(size_t)k * stride] = -dv;
  }
}


}  // namespace gen
#include <utility>
#include <vector>
namespace gen {
// Boundary cycles of the inside faces from the uploaded arrays (host, once per upload): the
// caller lists outer_ccb() of every inside face; any other halfedge bounding an inside face
// lies on an inner boundary (a hole) of it.  Holes of a face come out in halfedge-id order.
struct Cycles {
  std::vector<int> f_outer, f_hole_off, f_hole_cnt, hole_he;
  int n_general = 0;  // faces with holes or with an edge they bound on both sides
};
inline void derive_cycles(int n_he, int n_faces, const int *he_next, const int *he_face, const int *he_twin,
                          const uint8_t *face_inside, const int *f_off, const int *f_cnt, const int *f_list,
                          Cycles &C) {
  C.f_outer.assign(n_faces, -1);
  C.f_hole_off.assign(n_faces, 0);
  C.f_hole_cnt.assign(n_faces, 0);
  C.hole_he.clear();
  C.n_general = 0;
  std::vector<uint8_t> seen(n_he, 0), twice(n_faces, 0);
  for (int f = 0; f < n_faces; f++) {
    if (!face_inside[f] || f_cnt[f] <= 0) continue;
    C.f_outer[f] = f_list[f_off[f]];
    for (int k = 0; k < f_cnt[f]; k++) {
      const int h = f_list[f_off[f] + k];
      seen[h] = 1;
      if (he_face[he_twin[h]] == f) twice[f] = 1;
    }
  }
  std::vector<std::pair<int, int>> holes;  // (face, halfedge)
  for (int h = 0; h < n_he; h++) {
    const int f = he_face[h];
    if (seen[h] || !face_inside[f]) continue;
    holes.push_back({f, h});
    for (int g = h, guard = 0; !seen[g] && guard < n_he; g = he_next[g], guard++) seen[g] = 1;
  }
  for (auto &fh : holes) C.f_hole_cnt[fh.first]++;
  int run = 0;
  for (int f = 0; f < n_faces; f++) { C.f_hole_off[f] = run; run += C.f_hole_cnt[f]; }
  C.hole_he.assign(run > 0 ? run : 1, -1);
  std::vector<int> fill(n_faces, 0);
  for (auto &fh : holes) C.hole_he[C.f_hole_off[fh.first] + fill[fh.first]++] = fh.second;
  for (int f = 0; f < n_faces; f++)
    if (face_inside[f] && (C.f_hole_cnt[f] > 0 || twice[f])) C.n_general++;
}

}  // namespace gen
