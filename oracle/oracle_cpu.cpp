// oracle/oracle_cpu.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Double-precision C++ restatement of the reference's CPU pseudo-likelihood
// path, written the way the reference computes it: per candidate the parent
// and child polygons and the segment point lists are REBUILT
// (modified_elements), every registered feature function is then called on
// every removed/added element (Energy::statistic_variation), and value,
// gradient and Hessian are three separate passes over the stored statistics.
// It is used
//   * by tests/ as a second, fast checker (itself pinned against the exact
//     Python oracle's golden vectors, tests/test_oracle_cpu.py), and
//   * by bench.py as the timed CPU baseline (`cpu_baseline.kind = "port"`,
//     and the `--impl reference` arm), threaded over candidates.
// The reference itself (CGAL Lazy_exact_nt<Gmpq>) cannot be built here; this
// port uses doubles where the reference is exact, so it is FASTER than the
// real reference would be.  Exact predicates of the reference are replaced by
// forward-error-bounded ones (see `collinear`, `on_segment`).
//
// Reference functions restated (paths relative to /root/reference):
//   Split ctor                lib/ttessel.cpp:2090-2132
//   ray_exit_face             :4515-4576
//   Flip ctor                 :2432-2454
//   all_merges / all_flips    :1822-1846
//   Split/Merge/Flip::modified_elements  :2135-2232, :2302-2404, :2459-2713
//   face2poly / boundaries / simplify    :4314-4336, :4453-4506
//   polygon_insert_edge / polygon_remove_edge / find_edge_in_polygons
//                             :4600-4631, :4945-4966, :5046-5064
//   Seg::list_of_points       :864-873
//   Energy::statistic_variation  :2914-2980
//   feature functions         :4124-4434
//   PseudoLikDiscrete::GetValue/GetGradient/GetHessian  :3387-3443
// Scope: hole-free faces ("cas 1" of hpolygon_insert_edge/_remove_edge), as
// the device path.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../include/lite_b200.h"

namespace {

const double ORA_PI = 3.14159265358979323846;  // CGAL_PI

struct Pt {
  double x, y;
  int id;  // >= 0: vertex of the tessellation; < 0: constructed point
};
inline bool same(const Pt &a, const Pt &b) {
  if (a.id >= 0 && b.id >= 0) return a.id == b.id;
  return a.x == b.x && a.y == b.y;
}
typedef std::vector<Pt> Poly;
struct Edge {
  Pt a, b;
};
struct ModList {  // include/ttessel.h:489-498
  std::vector<Pt> del_vertices, add_vertices;
  std::vector<Edge> del_edges, add_edges;
  std::vector<Poly> del_faces, add_faces;
  std::vector<Poly> del_segs, add_segs;
};

inline double cross(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }
inline double dist(const Pt &a, const Pt &b) {
  double x = b.x - a.x, y = b.y - a.y;
  return sqrt(x * x + y * y);
}

struct View {
  const lite_tess_soa *s;
  Poly window;  // counter-clockwise
  double tol;   // absolute coordinate tolerance (64 ulp of the window size)
  Pt V(int v) const { return Pt{s->vx[v], s->vy[v], v}; }
  int twin(int h) const { return s->he_twin[h]; }
  int next(int h) const { return s->he_next[h]; }
  Pt src(int h) const { return V(s->he_src[h]); }
  Pt tgt(int h) const { return V(s->he_src[s->he_twin[h]]); }
};

// CGAL::collinear(p,q,r) with a forward error bound instead of exact arithmetic
inline bool collinear(const View &T, const Pt &p, const Pt &q, const Pt &r) {
  double c = cross(p.x - q.x, p.y - q.y, r.x - q.x, r.y - q.y);
  return fabs(c) <= T.tol * (dist(p, q) + dist(q, r));
}
inline bool on_segment(const View &T, const Pt &p, const Pt &a, const Pt &b) {
  double ex = b.x - a.x, ey = b.y - a.y, l = sqrt(ex * ex + ey * ey);
  if (l == 0) return false;
  if (fabs(cross(ex, ey, p.x - a.x, p.y - a.y)) > T.tol * l) return false;
  double t = ((p.x - a.x) * ex + (p.y - a.y) * ey) / l;
  return t >= -T.tol && t <= l + T.tol;
}
double poly_area(const Poly &p) {  // Polygon_2::area (signed)
  double a = 0;
  size_t n = p.size();
  for (size_t i = 0; i < n; i++) {
    const Pt &u = p[i], &v = p[(i + 1) % n];
    a += cross(u.x - p[0].x, u.y - p[0].y, v.x - p[0].x, v.y - p[0].y);
  }
  return 0.5 * a;
}

// ---------------------------------------------------------------------------
// feature functions, as written (lib/ttessel.cpp:4124-4434).  Arguments by
// value, like the reference's function-pointer signatures
// (include/ttessel.h:933-936).
// ---------------------------------------------------------------------------
double is_point_inside_window(Pt p, const View *T) {  // :4131-4135, :3583-3597
  Poly w = T->window;  // the reference copies the window on every call (:4133)
  size_t n = w.size();
  for (size_t i = 0; i < n; i++)
    if (on_segment(*T, p, w[i], w[(i + 1) % n])) return 0.0;  // ON_BOUNDARY
  bool in = false;
  for (size_t i = 0, j = n - 1; i < n; j = i++) {
    if ((w[i].y > p.y) != (w[j].y > p.y) &&
        p.x < (w[j].x - w[i].x) * (p.y - w[i].y) / (w[j].y - w[i].y) + w[i].x)
      in = !in;
  }
  return in ? 1.0 : 0.0;
}
double is_segment_internal(Poly s, const View *T) {  // :4172-4196
  Pt p1 = s.front(), p2 = s.back();
  if (is_point_inside_window(p1, T) > .5 || is_point_inside_window(p2, T) > .5) return 1.0;
  size_t n = T->window.size();  // window corners only: each polygon edge is one side
  for (size_t i = 0; i < n; i++) {
    const Pt &a = T->window[i], &b = T->window[(i + 1) % n];
    if (on_segment(*T, p1, a, b) && on_segment(*T, p2, a, b)) return 0.0;
  }
  return 1.0;
}
double minus_is_segment_internal(Poly s, const View *T) { return -is_segment_internal(s, T); }
double edge_length(Edge e, const View *T) {  // :4141-4149 (boundary edges only, as written)
  Poly s{e.a, e.b};
  if (is_segment_internal(s, T) > 0) return 0.0;
  return dist(e.a, e.b);
}
double edge_length_internal(Edge e, const View *T) {  // EXTENSION (not in the reference)
  Poly s{e.a, e.b};
  if (is_segment_internal(s, T) > 0) return dist(e.a, e.b);
  return 0.0;
}
double seg_number(Poly, const View *) { return 1.0; }   // :4157
double face_number(Poly, const View *) { return 1.0; }  // :4206
double face_area_2(Poly f, const View *) {              // :4215-4223
  double a = poly_area(f);
  return a * a;
}
double face_perimeter(Poly f, const View *) {  // :4229-4239
  double l = 0;
  size_t n = f.size();
  for (size_t j = 0; j < n; j++) l += dist(f[j], f[(j + 1) % n]);
  return l;
}
double face_shape(Poly f, const View *T) {  // :4247-4249
  return face_perimeter(f, T) / (4 * pow(face_area_2(f, T), 0.25));
}
double angle_between_vectors(double ax, double ay, double bx, double by) {  // :4255-4268
  double n1 = ax * ax + ay * ay, n2 = bx * bx + by * by, p = ax * bx + ay * by;
  double c = p / sqrt(n1 * n2);
  if (c >= 1.0) return 0.0;
  if (c <= -1.0) return ORA_PI;
  return acos(c);
}
double face_sum_of_angles(Poly ob, const View *) {  // :4285-4305
  size_t n = ob.size();
  double tot = 0;
  for (size_t i = 0; i < n; i++) {
    size_t j = (i + n - 1) % n, k = (i + 1) % n;
    double a = ORA_PI / 2 - angle_between_vectors(ob[j].x - ob[i].x, ob[j].y - ob[i].y,
                                                  ob[k].x - ob[i].x, ob[k].y - ob[i].y);
    if (a >= 0) tot += a;
  }
  return tot;
}
double min_angle(Poly ob, const View *) {  // :4348-4367 (seeded with the turning angle at corner 0)
  size_t n = ob.size();
  double m = angle_between_vectors(ob[0].x - ob[n - 1].x, ob[0].y - ob[n - 1].y,
                                   ob[1].x - ob[0].x, ob[1].y - ob[0].y);
  for (size_t i = 0; i < n; i++) {
    size_t j = (i + n - 1) % n, k = (i + 1) % n;
    double a = angle_between_vectors(ob[j].x - ob[i].x, ob[j].y - ob[i].y, ob[k].x - ob[i].x,
                                     ob[k].y - ob[i].y);
    if (a < m) m = a;
  }
  return m;
}
double segment_size_2(Poly s, const View *T) {  // :4424-4434 (boundary segments only, as written)
  Poly ends{s.front(), s.back()};
  if (is_segment_internal(ends, T) > 0) return 0.0;
  double k = (double)s.size() - 1;
  return k * k;
}

// Energy: feature registry in flattened CatVector order (lib/ttessel.cpp:2750-2761)
struct Energy {
  int d = 0;
  int fid[LITE_MAX_D];
  int cat[LITE_MAX_D];
};
int category_of(int f) {
  switch (f) {
    case LITE_FEAT_IS_POINT_INSIDE_WINDOW: return 0;
    case LITE_FEAT_EDGE_LENGTH: case LITE_FEAT_EDGE_LENGTH_INTERNAL: return 1;
    case LITE_FEAT_FACE_NUMBER: case LITE_FEAT_FACE_AREA_2: case LITE_FEAT_FACE_PERIMETER:
    case LITE_FEAT_FACE_SHAPE: case LITE_FEAT_FACE_SUM_OF_ANGLES: case LITE_FEAT_MIN_ANGLE: return 2;
    case LITE_FEAT_SEG_NUMBER: case LITE_FEAT_IS_SEGMENT_INTERNAL:
    case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: case LITE_FEAT_SEGMENT_SIZE_2: return 3;
    default: return -1;
  }
}
double call_vertex(int f, const Pt &p, const View *T) { (void)f; return is_point_inside_window(p, T); }
double call_edge(int f, const Edge &e, const View *T) {
  return f == LITE_FEAT_EDGE_LENGTH ? edge_length(e, T) : edge_length_internal(e, T);
}
double call_face(int f, const Poly &p, const View *T) {
  switch (f) {
    case LITE_FEAT_FACE_NUMBER: return face_number(p, T);
    case LITE_FEAT_FACE_AREA_2: return face_area_2(p, T);
    case LITE_FEAT_FACE_PERIMETER: return face_perimeter(p, T);
    case LITE_FEAT_FACE_SHAPE: return face_shape(p, T);
    case LITE_FEAT_FACE_SUM_OF_ANGLES: return face_sum_of_angles(p, T);
    default: return min_angle(p, T);
  }
}
double call_seg(int f, const Poly &s, const View *T) {
  switch (f) {
    case LITE_FEAT_SEG_NUMBER: return seg_number(s, T);
    case LITE_FEAT_IS_SEGMENT_INTERNAL: return is_segment_internal(s, T);
    case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: return minus_is_segment_internal(s, T);
    default: return segment_size_2(s, T);
  }
}
// Energy::statistic_variation (lib/ttessel.cpp:2914-2980); out = MINUS the
// variation, the form PseudoLikDiscrete stores (:3337, :3345, :3363)
void statistic_variation_neg(const Energy &E, const ModList &m, const View *T, double *out) {
  for (int i = 0; i < E.d; i++) {
    double loc = 0;
    int f = E.fid[i];
    switch (E.cat[i]) {
      case 0:
        for (const Pt &p : m.del_vertices) loc -= call_vertex(f, p, T);
        for (const Pt &p : m.add_vertices) loc += call_vertex(f, p, T);
        break;
      case 1:
        for (const Edge &e : m.del_edges) loc -= call_edge(f, e, T);
        for (const Edge &e : m.add_edges) loc += call_edge(f, e, T);
        break;
      case 2:
        for (const Poly &p : m.del_faces) loc -= call_face(f, p, T);
        for (const Poly &p : m.add_faces) loc += call_face(f, p, T);
        break;
      default:
        for (const Poly &p : m.del_segs) loc -= call_seg(f, p, T);
        for (const Poly &p : m.add_segs) loc += call_seg(f, p, T);
        break;
    }
    out[i] = -loc;
  }
}

// ---------------------------------------------------------------------------
// polygon helpers
// ---------------------------------------------------------------------------
Poly face2poly(const View &T, int f, bool simplify) {  // :4314-4336 (outer ccb only)
  const lite_tess_soa *s = T.s;
  Poly p;
  int off = s->face_he_offset[f], n = s->face_he_count[f];
  for (int k = 0; k < n; k++) {
    int e = s->face_he_list[off + k];
    if (!simplify || s->he_next_hf[e] < 0 || s->he_next_hf[e] != s->he_next[e]) p.push_back(T.tgt(e));
  }
  return p;
}
Poly boundaries(Poly ob) {  // :4453-4469, outer boundary; reverse_orientation keeps vertex 0
  if (poly_area(ob) < 0) std::reverse(ob.begin() + 1, ob.end());
  return ob;
}
Poly simplify(const View &T, const Poly &p) {  // :4474-4492
  Poly sp;
  size_t n = p.size();
  for (size_t i = 0; i < n; i++)
    if (!collinear(T, p[(i + n - 1) % n], p[i], p[(i + 1) % n])) sp.push_back(p[i]);
  return sp;
}
struct PECirc {  // polygon edge circulator: edge k goes poly[k] -> poly[k+1]
  const Poly *poly;
  int i;
  const Pt &source() const { return (*poly)[i]; }
  const Pt &target() const { return (*poly)[(i + 1) % (int)poly->size()]; }
  void inc() { i = (i + 1) % (int)poly->size(); }
  bool operator==(const PECirc &o) const { return poly == o.poly && i == o.i; }
  bool operator!=(const PECirc &o) const { return !(*this == o); }
};
bool find_edge_in_polygon(const Poly &poly, const Pt &a, const Pt &b, PECirc &out) {  // :5046-5064
  int n = (int)poly.size();
  for (int k = 0; k < n; k++)
    if (same(poly[k], a) && same(poly[(k + 1) % n], b)) {
      out.poly = &poly;
      out.i = k;
      return true;
    }
  return false;
}
Poly polygon_insert_edge(PECirc e1, PECirc e2, const Pt &p1, const Pt &p2) {  // :4600-4631
  Poly res;
  res.push_back(p1);
  res.push_back(p2);
  PECirc e = e2;
  if (same(e.target(), p2)) e.inc();
  do {
    res.push_back(e.target());
    e.inc();
  } while (e != e2 && e != e1);
  if (e == e1) {
    if (same(e1.source(), p1)) res.pop_back();
    return res;
  }
  res.push_back(p2);
  res.push_back(p1);
  e = e1;
  if (same(e1.target(), p1)) e.inc();
  do {
    res.push_back(e.target());
    e.inc();
  } while (e != e1);
  if (same(e1.source(), p1)) res.pop_back();
  return res;
}
Poly polygon_remove_edge(PECirc e1, PECirc e2) {  // :4945-4966
  Poly res;
  PECirc circ = e1;
  circ.inc();
  while (circ != e1 && circ != e2) {
    res.push_back(circ.target());
    circ.inc();
  }
  if (circ == e2) return res;
  circ = e2;
  circ.inc();
  while (circ != e2) {
    res.push_back(circ.target());
    circ.inc();
  }
  return res;
}
Poly list_of_points(const View &T, int seg) {  // Seg::list_of_points, :864-873
  Poly lp;
  int e = T.s->seg_he_start[seg];
  lp.push_back(T.src(e));
  while (e >= 0) {
    lp.push_back(T.tgt(e));
    e = T.s->he_next_hf[e];
  }
  return lp;
}
void insert_before(Poly &seg, const Pt &key, const Pt &p) {
  for (size_t i = 0; i < seg.size(); i++)
    if (same(seg[i], key)) {
      seg.insert(seg.begin() + i, p);
      return;
    }
  seg.push_back(p);  // std::find == end(): insert at end
}
void remove_point(Poly &seg, const Pt &key) {
  for (size_t i = 0; i < seg.size(); i++)
    if (same(seg[i], key)) {
      seg.erase(seg.begin() + i);
      return;
    }
}

// ---------------------------------------------------------------------------
// ray_exit_face (:4515-4576) for a hole-free face.  The line through the ray
// is a*X + b*Y = w; a vertex is classified once by the sign of its residual,
// so consecutive halfedges agree at their shared vertex.  The reference's
// `len==0` skip (:4552) becomes: skip halfedges that contain the ray source
// (the halfedge `on_he` it was constructed on, and those incident to the
// vertex `at_vertex` when the source is a vertex).
// ---------------------------------------------------------------------------
bool ray_exit_face(const View &T, int f, double a, double b, double w, const Pt &src, double dx,
                   double dy, int on_he, int at_vertex, Pt &hit, int &e_out) {
  const lite_tess_soa *s = T.s;
  int off = s->face_he_offset[f], n = s->face_he_count[f];
  double lp = 1.7976931348623157e308;
  bool found = false;
  for (int k = 0; k < n; k++) {
    int ce = s->face_he_list[off + k];
    if (ce == on_he) continue;
    int vs = s->he_src[ce], vt = s->he_src[s->he_twin[ce]];
    if (at_vertex >= 0 && (vs == at_vertex || vt == at_vertex)) continue;
    double sx = s->vx[vs], sy = s->vy[vs], tx = s->vx[vt], ty = s->vy[vt];
    double sk = a * sx + b * sy - w, sn = a * tx + b * ty - w;
    bool cr = (sk <= 0 && sn >= 0) || (sk >= 0 && sn <= 0);
    if (!cr || sk == sn) continue;
    double g = sk / (sk - sn);
    double X = sx + g * (tx - sx), Y = sy + g * (ty - sy);
    double rx = X - src.x, ry = Y - src.y;
    if (rx * dx + ry * dy < 0) continue;  // behind the ray source
    double len = sqrt(rx * rx + ry * ry);
    if (len == 0) continue;
    if (len < lp) {
      lp = len;
      hit = Pt{X, Y, -2};
      e_out = ce;
      found = true;
    }
  }
  return found;
}

struct SplitC {  // TTessel::Split, :2090-2132
  int e1 = -1, e2 = -1;
  Pt p1{0, 0, -1}, p2{0, 0, -2};
  bool ok = false;
};
SplitC make_split(const View &T, int e, double x, double angle) {
  SplitC r;
  r.e1 = e;
  Pt S = T.src(e), Q = T.tgt(e);
  double e1_angle = atan2(Q.y - S.y, Q.x - S.x);
  double l_angle = e1_angle + angle;
  double a = sin(l_angle), b = -cos(l_angle);
  double u = a * S.x + b * S.y, v = a * Q.x + b * Q.y;
  double w = u + x * (v - u);
  double ss = a * S.x + b * S.y - w, st = a * Q.x + b * Q.y - w;
  if (ss == st) return r;
  double t = ss / (ss - st);
  if (t < 0 || t > 1) return r;
  r.p1 = Pt{S.x + t * (Q.x - S.x), S.y + t * (Q.y - S.y), -1};
  double dx, dy;  // Line_2(a,b,c) has direction (b,-a) (:2121-2126)
  if (st < 0) { dx = b; dy = -a; } else { dx = -b; dy = a; }
  int at_vertex = -1;
  if (x == 0.0) { r.p1 = S; at_vertex = S.id; }
  r.ok = ray_exit_face(T, T.s->he_face[e], a, b, w, r.p1, dx, dy, e, at_vertex, r.p2, r.e2);
  return r;
}

ModList split_modified_elements(const View &T, const SplitC &sp) {  // :2135-2232
  ModList m;
  int e1 = sp.e1, e2 = sp.e2;
  const Pt &p1 = sp.p1, &p2 = sp.p2;
  m.add_vertices.push_back(p1);
  m.add_vertices.push_back(p2);
  m.del_edges.push_back(Edge{T.src(e1), T.tgt(e1)});
  m.del_edges.push_back(Edge{T.src(e2), T.tgt(e2)});
  m.add_edges.push_back(Edge{T.src(e1), p1});
  m.add_edges.push_back(Edge{p1, T.tgt(e1)});
  m.add_edges.push_back(Edge{T.src(e2), p2});
  m.add_edges.push_back(Edge{p2, T.tgt(e2)});
  m.add_edges.push_back(Edge{p1, p2});
  int f = T.s->he_face[e1];
  m.del_faces.push_back(face2poly(T, f, true));
  Poly fb = boundaries(face2poly(T, f, false));
  PECirc pe1{nullptr, 0}, pe2{nullptr, 0};
  find_edge_in_polygon(fb, T.src(e1), T.tgt(e1), pe1);
  find_edge_in_polygon(fb, T.src(e2), T.tgt(e2), pe2);
  // hpolygon_insert_edge cas 1 (:4700-4705)
  m.add_faces.push_back(simplify(T, boundaries(polygon_insert_edge(pe1, pe2, p1, p2))));
  m.add_faces.push_back(simplify(T, boundaries(polygon_insert_edge(pe2, pe1, p2, p1))));
  const int es[2] = {e1, e2};
  const Pt ps[2] = {p1, p2};
  for (int k = 0; k < 2; k++) {
    Poly seg = list_of_points(T, T.s->he_seg[es[k]]);
    m.del_segs.push_back(seg);
    insert_before(seg, T.s->he_dir[es[k]] ? T.tgt(es[k]) : T.src(es[k]), ps[k]);
    m.add_segs.push_back(seg);
  }
  m.add_segs.push_back(Poly{p1, p2});
  return m;
}

ModList merge_modified_elements(const View &T, int e) {  // :2302-2404
  ModList m;
  const lite_tess_soa *s = T.s;
  int te = T.twin(e);
  int e1 = T.next(te), e2 = T.next(e);  // Merge::e1(), e2() (include/ttessel.h:657-661)
  Pt p1 = T.src(e), p2 = T.tgt(e);
  m.del_vertices.push_back(p1);
  m.del_vertices.push_back(p2);
  m.del_edges.push_back(Edge{p1, p2});
  m.del_edges.push_back(Edge{p2, T.tgt(e2)});
  m.del_edges.push_back(Edge{T.src(s->he_prev_hf[e2]), p2});
  m.del_edges.push_back(Edge{p1, T.tgt(e1)});
  m.del_edges.push_back(Edge{T.src(s->he_prev_hf[e1]), p1});
  m.add_edges.push_back(Edge{T.src(s->he_prev_hf[e2]), T.tgt(e2)});
  m.add_edges.push_back(Edge{T.src(s->he_prev_hf[e1]), T.tgt(e1)});
  Poly fb = boundaries(face2poly(T, s->he_face[e], false));
  PECirc pe{nullptr, 0}, pet{nullptr, 0};
  find_edge_in_polygon(fb, p1, p2, pe);
  m.del_faces.push_back(simplify(T, fb));
  Poly ftb = boundaries(face2poly(T, s->he_face[te], false));
  find_edge_in_polygon(ftb, p2, p1, pet);
  m.del_faces.push_back(simplify(T, ftb));
  m.add_faces.push_back(simplify(T, boundaries(polygon_remove_edge(pe, pet))));  // cas 1 (:4859-4865)
  Poly seg = list_of_points(T, s->he_seg[e1]);
  m.del_segs.push_back(seg);
  remove_point(seg, p1);
  m.add_segs.push_back(seg);
  seg = list_of_points(T, s->he_seg[e2]);
  m.del_segs.push_back(seg);
  remove_point(seg, p2);
  m.add_segs.push_back(seg);
  m.del_segs.push_back(Poly{p1, p2});
  return m;
}

struct FlipC {  // TTessel::Flip, :2432-2454
  int e1 = -1, e2 = -1, e3 = -1;
  Pt p2{0, 0, -2};
  bool right = false, ok = false;
};
FlipC make_flip(const View &T, int e1) {
  FlipC r;
  r.e1 = e1;
  const lite_tess_soa *s = T.s;
  int s1 = s->he_seg[e1];
  // incoming halfedges at src(e1): twin(e1), then twin(next(c)) around the vertex
  int c = T.twin(e1), c0 = c;
  do {
    if (s->he_seg[c] != s1) { r.e3 = c; break; }
    c = T.twin(T.next(c));
  } while (c != c0);
  if (r.e3 < 0) return r;
  Pt a = T.src(e1), q = T.tgt(e1), s3 = T.src(r.e3);
  double dx = a.x - s3.x, dy = a.y - s3.y;
  // CGAL::orientation(e1.source, e1.target, e3.source) (:2442)
  double o = cross(q.x - a.x, q.y - a.y, s3.x - a.x, s3.y - a.y);
  int f;
  if (o > 0) { f = s->he_face[T.twin(e1)]; r.right = true; } else { f = s->he_face[e1]; r.right = false; }
  double la = dy, lb = -dx, lw = la * a.x + lb * a.y;
  r.ok = ray_exit_face(T, f, la, lb, lw, a, dx, dy, -1, a.id, r.p2, r.e2);
  return r;
}

bool flip_modified_elements(const View &T, const FlipC &fl, ModList &m) {  // :2459-2713
  const lite_tess_soa *s = T.s;
  int e1 = fl.e1, e2 = fl.e2;
  const Pt &p2 = fl.p2;
  Pt a = T.src(e1), q = T.tgt(e1);
  m.del_vertices.push_back(q);
  m.add_vertices.push_back(p2);
  int n1 = T.next(e1), ph = s->he_prev_hf[n1];
  Edge s_merge{a, q}, s1_merge{T.src(n1), T.tgt(n1)}, s2_merge{T.tgt(ph), T.src(ph)};
  Pt p_merge = s1_merge.a;
  Edge s_split{a, p2}, s1_split{p2, T.tgt(e2)}, s2_split{p2, T.src(e2)};
  m.del_edges.push_back(s_merge);
  m.del_edges.push_back(s1_merge);
  m.del_edges.push_back(s2_merge);
  bool generic = !same(s1_split.b, p_merge) && !same(s2_split.b, p_merge);
  if (generic) m.del_edges.push_back(Edge{s1_split.b, s2_split.b});
  m.add_edges.push_back(s_split);
  if (generic) {
    m.add_edges.push_back(s1_split);
    m.add_edges.push_back(s2_split);
    m.add_edges.push_back(Edge{s1_merge.b, s2_merge.b});
  }
  if (same(s1_split.b, p_merge)) {
    m.add_edges.push_back(s2_split);
    m.add_edges.push_back(Edge{s1_split.a, s1_merge.b});
  }
  if (same(s2_split.b, p_merge)) {
    m.add_edges.push_back(s1_split);
    m.add_edges.push_back(Edge{s2_split.a, s2_merge.b});
  }
  // faces (:2515-2655): insert (a,p2) into e2's face, then remove e1
  Poly f2b = boundaries(face2poly(T, s->he_face[e2], false));
  PECirc pe2{nullptr, 0}, pe_p1{nullptr, 0};
  if (!find_edge_in_polygon(f2b, T.src(e2), T.tgt(e2), pe2)) return false;
  bool okp = fl.right ? find_edge_in_polygon(f2b, T.src(T.twin(e1)), a, pe_p1)
                      : find_edge_in_polygon(f2b, a, q, pe_p1);
  if (!okp) return false;
  Poly ins[2];
  ins[0] = boundaries(polygon_insert_edge(pe_p1, pe2, a, p2));
  ins[1] = boundaries(polygon_insert_edge(pe2, pe_p1, p2, a));
  bool match[2] = {false, false};
  bool del_f1 = false, del_f1_twin = false;
  const Poly *poly = nullptr, *poly_twin = nullptr;
  Poly face_e1, face_e1t;
  PECirc pe1{nullptr, 0}, pe1t{nullptr, 0};
  if (find_edge_in_polygon(ins[1], a, q, pe1)) { poly = &ins[1]; match[1] = true; }
  else if (find_edge_in_polygon(ins[0], a, q, pe1)) { poly = &ins[0]; match[0] = true; }
  else {
    face_e1 = boundaries(face2poly(T, s->he_face[e1], false));
    if (!find_edge_in_polygon(face_e1, a, q, pe1)) return false;
    poly = &face_e1;
    del_f1 = true;
  }
  bool found = false, twin_is_poly = false;
  if (match[0]) {
    if (find_edge_in_polygon(*poly, q, a, pe1t)) { found = true; twin_is_poly = true; }
  } else if (find_edge_in_polygon(ins[0], q, a, pe1t)) {
    poly_twin = &ins[0]; found = true; match[0] = true;
  }
  if (!found) {
    if (match[1]) {
      if (find_edge_in_polygon(*poly, q, a, pe1t)) { found = true; twin_is_poly = true; }
    } else if (find_edge_in_polygon(ins[1], q, a, pe1t)) {
      poly_twin = &ins[1]; found = true; match[1] = true;
    }
  }
  if (!found) {
    face_e1t = boundaries(face2poly(T, s->he_face[T.twin(e1)], false));
    if (!find_edge_in_polygon(face_e1t, q, a, pe1t) || del_f1) return false;
    poly_twin = &face_e1t;
    del_f1_twin = true;
  }
  if (twin_is_poly) return false;  // edge inside one face: holed result, outside the scope
  Poly removed = polygon_remove_edge(pe1, pe1t);
  m.del_faces.push_back(simplify(T, f2b));
  if (del_f1) m.del_faces.push_back(simplify(T, *poly));
  if (del_f1_twin) m.del_faces.push_back(simplify(T, *poly_twin));
  for (int k = 0; k < 2; k++)
    if (!match[k]) m.add_faces.push_back(simplify(T, ins[k]));
  m.add_faces.push_back(simplify(T, boundaries(removed)));
  // segments (:2659-2712)
  Poly seg = list_of_points(T, s->he_seg[e1]);
  m.del_segs.push_back(seg);
  if (s->he_dir[e1]) seg.pop_back(); else seg.erase(seg.begin());
  m.add_segs.push_back(seg);
  int e3 = fl.e3;
  seg = list_of_points(T, s->he_seg[e3]);
  m.del_segs.push_back(seg);
  if (s->he_dir[e3]) seg.push_back(p2); else seg.insert(seg.begin(), p2);
  m.add_segs.push_back(seg);
  seg = list_of_points(T, s->he_seg[e2]);
  m.del_segs.push_back(seg);
  insert_before(seg, s->he_dir[e2] ? T.tgt(e2) : T.src(e2), p2);
  if (s->he_seg[n1] == s->he_seg[e2]) {
    remove_point(seg, q);
    m.add_segs.push_back(seg);
    return true;
  }
  m.add_segs.push_back(seg);
  seg = list_of_points(T, s->he_seg[n1]);
  m.del_segs.push_back(seg);
  remove_point(seg, q);
  m.add_segs.push_back(seg);
  return true;
}

bool build_view(const lite_tess_soa *s, View &T, std::string &err) {
  T.s = s;
  T.window.clear();
  // the window boundary = the ccb of the outside face, reversed to counter-clockwise
  int start = -1;
  for (int h = 0; h < s->n_halfedges; h++)
    if (!s->face_inside[s->he_face[h]]) { start = h; break; }
  if (start < 0) { err = "no outside face"; return false; }
  int h = start, guard = 0;
  do {
    T.window.push_back(T.src(h));
    h = s->he_next[h];
    if (++guard > s->n_halfedges) { err = "outside face boundary does not close"; return false; }
  } while (h != start);
  std::reverse(T.window.begin(), T.window.end());
  double lo = 1e300, hi = -1e300;
  for (const Pt &p : T.window) {
    lo = std::min(lo, std::min(p.x, p.y));
    hi = std::max(hi, std::max(p.x, p.y));
  }
  T.tol = 64 * 2.220446049250313e-16 * std::max(hi - lo, std::max(fabs(hi), fabs(lo)));
  T.window = simplify(T, T.window);  // the domain polygon: drop the T-vertices lying on its sides
  return true;
}

template <typename F>
void parallel_for(int64_t n, int nthreads, F fn) {
  if (nthreads <= 1 || n < 2) { fn(0, n, 0); return; }
  std::vector<std::thread> th;
  int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    int64_t a = t * chunk, b = std::min(n, a + chunk);
    if (a >= b) break;
    th.emplace_back([=] { fn(a, b, t); });
  }
  for (auto &x : th) x.join();
}

bool set_energy(int d, const int32_t *fid, Energy &E) {
  if (d < 1 || d > LITE_MAX_D) return false;
  E.d = d;
  for (int i = 0; i < d; i++) {
    E.fid[i] = fid[i];
    E.cat[i] = category_of(fid[i]);
    if (E.cat[i] < 0) return false;
  }
  return true;
}

// Philox4x32-10 and the stratified length-weighted sampler, identical to the
// device sampler (lite_b200/csrc/kernels.cuh) so both arms see the same list.
inline void philox4x32_10(uint64_t ctr, uint32_t stream, uint64_t seed, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
inline double u53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

thread_local std::string g_err;
}  // namespace

extern "C" {

const char *oracle_last_error(void) { return g_err.c_str(); }

int oracle_hardware_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

/* TTessel::split_sample(n) with the Philox stream of the device path:
 * candidates j0 .. j0+n-1 of a batch of n_batch. */
int oracle_sample_splits(const lite_tess_soa *s, int64_t n_batch, uint64_t seed, uint64_t ctr0,
                         int64_t j0, int64_t n, int32_t *he, double *x, double *angle, int nthreads) {
  std::vector<double> cum;
  std::vector<int> cum_he;
  double cl = 0;
  // slot order: faces in index order, each in ccb order from outer_ccb()
  for (int f = 0; f < s->n_faces; f++) {
    if (!s->face_inside[f]) continue;
    for (int k = 0; k < s->face_he_count[f]; k++) {
      int h = s->face_he_list[s->face_he_offset[f] + k];
      cl += s->he_length[h];  // cumlen += e->get_length(), lib/ttessel.cpp:1980
      cum.push_back(cl);
      cum_he.push_back(h);
    }
  }
  if (cum.empty()) { g_err = "no inside halfedge"; return -1; }
  parallel_for(n, nthreads, [&](int64_t a, int64_t b, int) {
    for (int64_t i = a; i < b; i++) {
      uint64_t j = (uint64_t)(j0 + i);
      uint32_t r[4];
      philox4x32_10(ctr0 + j, 0u, seed, r);
      x[i] = u53(r[0], r[1]);
      uint32_t spare = ((r[0] & 31u) << 17) | ((r[1] & 63u) << 11) | ((r[2] & 31u) << 6) | (r[3] & 63u);
      double U = u53(r[2], r[3]), jit = (double)spare * (1.0 / 4194304.0);
      double sel = ((double)j + jit) * (cl / (double)n_batch);
      size_t k = std::upper_bound(cum.begin(), cum.end(), sel) - cum.begin();  // first cum > sel
      if (k >= cum.size()) k = cum.size() - 1;
      he[i] = cum_he[k];
      angle[i] = acos(1.0 - 2.0 * U);  // lib/ttessel.cpp:1792
    }
  });
  return 0;
}

/* Config "random_lines sampling and face clipping" (loop 3a alone): split_sample
 * (lib/ttessel.cpp:1783-1795) + the Split constructor (:2090-2132) for candidates
 * j0 .. j0+n-1 of a batch, nothing stored; checksum2 as lite_lines_sample_clip:
 * [0] xor of hash(e1<<32|e2), [1] sum of e2.  Returns the number of lines without exit. */
int64_t oracle_lines_clip(const lite_tess_soa *s, int64_t n_batch, uint64_t seed, uint64_t ctr0, int64_t j0,
                          int64_t n, uint64_t *checksum2, int nthreads) {
  View T;
  std::string err;
  if (!build_view(s, T, err)) { g_err = err; return -1; }
  std::vector<double> cum;
  std::vector<int> cum_he;
  double cl = 0;
  for (int f = 0; f < s->n_faces; f++) {
    if (!s->face_inside[f]) continue;
    for (int k = 0; k < s->face_he_count[f]; k++) {
      int h = s->face_he_list[s->face_he_offset[f] + k];
      cl += s->he_length[h];
      cum.push_back(cl);
      cum_he.push_back(h);
    }
  }
  if (cum.empty()) { g_err = "no inside halfedge"; return -1; }
  std::atomic<int64_t> bad(0);
  std::atomic<uint64_t> cx(0), cs(0);
  parallel_for(n, nthreads, [&](int64_t a, int64_t b, int) {
    uint64_t lx = 0, ls = 0;
    int64_t lbad = 0;
    for (int64_t i = a; i < b; i++) {
      uint64_t j = (uint64_t)(j0 + i);
      uint32_t r[4];
      philox4x32_10(ctr0 + j, 0u, seed, r);
      double x = u53(r[0], r[1]);
      uint32_t spare = ((r[0] & 31u) << 17) | ((r[1] & 63u) << 11) | ((r[2] & 31u) << 6) | (r[3] & 63u);
      double U = u53(r[2], r[3]), jit = (double)spare * (1.0 / 4194304.0);
      double sel = ((double)j + jit) * (cl / (double)n_batch);
      size_t k = std::upper_bound(cum.begin(), cum.end(), sel) - cum.begin();
      if (k >= cum.size()) k = cum.size() - 1;
      int he = cum_he[k];
      SplitC sp = make_split(T, he, x, acos(1.0 - 2.0 * U));
      int e2 = sp.ok ? sp.e2 : -1;
      if (!sp.ok) lbad++;
      uint64_t key = ((uint64_t)(uint32_t)he << 32) | (uint32_t)e2;
      key *= 0x9E3779B97F4A7C15ull; key ^= key >> 29;
      lx ^= key; ls += (uint64_t)(uint32_t)e2;
    }
    cx ^= lx; cs += ls; bad += lbad;
  });
  if (checksum2) { checksum2[0] = cx.load(); checksum2[1] = cs.load(); }
  return bad.load();
}

/* PseudoLikDiscrete::AddGivenSplit over a list (lib/ttessel.cpp:3360-3364):
 * stat[i*d+k] = -statistic_variation(Split(he[i],x[i],angle[i]))[k].
 * e2_out / pts_out (p1x,p1y,p2x,p2y per candidate) may be NULL.
 * Returns the number of candidates whose ray found no exit (their rows are 0). */
int64_t oracle_split_stats(const lite_tess_soa *s, int d, const int32_t *fid, int64_t n,
                           const int32_t *he, const double *x, const double *angle,
                           int32_t *e2_out, double *pts_out, double *stat, int nthreads) {
  View T;
  Energy E;
  std::string err;
  if (!build_view(s, T, err) || !set_energy(d, fid, E)) { g_err = err.empty() ? "bad energy" : err; return -1; }
  std::atomic<int64_t> bad(0);
  parallel_for(n, nthreads, [&](int64_t a, int64_t b, int) {
    for (int64_t i = a; i < b; i++) {
      SplitC sp = make_split(T, he[i], x[i], angle[i]);
      if (e2_out) e2_out[i] = sp.ok ? sp.e2 : -1;
      if (pts_out) {
        pts_out[4 * i] = sp.p1.x; pts_out[4 * i + 1] = sp.p1.y;
        pts_out[4 * i + 2] = sp.p2.x; pts_out[4 * i + 3] = sp.p2.y;
      }
      if (!sp.ok) {
        bad++;
        for (int k = 0; k < d; k++) stat[i * d + k] = 0;
        continue;
      }
      ModList m = split_modified_elements(T, sp);
      statistic_variation_neg(E, m, &T, stat + i * d);
    }
  });
  return bad.load();
}

/* all_merges (lib/ttessel.cpp:1822-1829) + statistic variations (:3335-3338) */
int64_t oracle_merge_stats(const lite_tess_soa *s, int d, const int32_t *fid, int32_t *he_out,
                           double *stat, int nthreads) {
  View T;
  Energy E;
  std::string err;
  if (!build_view(s, T, err) || !set_energy(d, fid, E)) { g_err = err.empty() ? "bad energy" : err; return -1; }
  parallel_for(s->n_non_blocking, nthreads, [&](int64_t a, int64_t b, int) {
    for (int64_t i = a; i < b; i++) {
      int e = s->seg_he_start[s->non_blocking[i]];  // Seg::get_halfedge_handle(): dir == true
      if (he_out) he_out[i] = e;
      ModList m = merge_modified_elements(T, e);
      statistic_variation_neg(E, m, &T, stat + i * d);
    }
  });
  return 0;
}

/* all_flips (lib/ttessel.cpp:1834-1846): flips[i] = Flip(twin(start)), flips[i+n] = Flip(end) */
int64_t oracle_flip_stats(const lite_tess_soa *s, int d, const int32_t *fid, int32_t *e1_out,
                          int32_t *e2_out, int32_t *right_out, double *p2_out, double *stat,
                          int nthreads) {
  View T;
  Energy E;
  std::string err;
  if (!build_view(s, T, err) || !set_energy(d, fid, E)) { g_err = err.empty() ? "bad energy" : err; return -1; }
  const int nb = s->n_blocking;
  std::atomic<int64_t> bad(0);
  parallel_for(2 * (int64_t)nb, nthreads, [&](int64_t a, int64_t b, int) {
    for (int64_t i = a; i < b; i++) {
      int sg = s->blocking[i < nb ? i : i - nb];
      int e1 = i < nb ? s->he_twin[s->seg_he_start[sg]] : s->seg_he_end[sg];
      FlipC f = make_flip(T, e1);
      if (e1_out) e1_out[i] = e1;
      if (e2_out) e2_out[i] = f.ok ? f.e2 : -1;
      if (right_out) right_out[i] = f.right ? 1 : 0;
      if (p2_out) { p2_out[2 * i] = f.p2.x; p2_out[2 * i + 1] = f.p2.y; }
      ModList m;
      if (!f.ok || !flip_modified_elements(T, f, m)) {
        bad++;
        for (int k = 0; k < d; k++) stat[i * d + k] = 0;
        continue;
      }
      statistic_variation_neg(E, m, &T, stat + i * d);
    }
  });
  return bad.load();
}

/* PseudoLikDiscrete::GetValue / GetGradient / GetHessian (lib/ttessel.cpp:3387-3443):
 * three separate passes, one exp per candidate per pass, as the reference.
 * split_stat: n_s x d, flip_stat: n_f x d (row major, negated variations);
 * sum_m / sum_f: the sums kept by SetEnergy.  n_global = |S| of the normaliser. */
int oracle_pl(int d, const double *theta, int64_t n_s, const double *split_stat, int64_t n_f,
              const double *flip_stat, const double *sum_m, const double *sum_f, double int_length,
              int64_t n_global, double *lpl, double *grad, double *hess, int nthreads) {
  if (n_global <= 0) { g_err = "no dummy splits"; return -1; }
  const double c = int_length / (ORA_PI * (double)n_global);
  int nt = nthreads < 1 ? 1 : nthreads;
  auto dot = [&](const double *p) {
    double v = 0;
    for (int k = 0; k < d; k++) v += theta[k] * p[k];
    return v;
  };
  // GetValue
  if (lpl) {
    std::vector<double> part(nt, 0.0), partf(nt, 0.0);
    parallel_for(n_s, nt, [&](int64_t a, int64_t b, int t) {
      double v = 0;
      for (int64_t i = a; i < b; i++) v += exp(dot(split_stat + i * d));
      part[t] = v;
    });
    parallel_for(n_f, nt, [&](int64_t a, int64_t b, int t) {
      double v = 0;
      for (int64_t i = a; i < b; i++) v += exp(dot(flip_stat + i * d));
      partf[t] = v;
    });
    double S = 0, F = 0;
    for (double v : part) S += v;
    for (double v : partf) F += v;
    *lpl = -dot(sum_m) - c * S - dot(sum_f) - F;
  }
  // GetGradient
  if (grad) {
    std::vector<double> part((size_t)nt * d, 0.0), partf((size_t)nt * d, 0.0);
    parallel_for(n_s, nt, [&](int64_t a, int64_t b, int t) {
      for (int64_t i = a; i < b; i++) {
        const double *p = split_stat + i * d;
        double e = exp(dot(p));
        for (int k = 0; k < d; k++) part[(size_t)t * d + k] += e * p[k];
      }
    });
    parallel_for(n_f, nt, [&](int64_t a, int64_t b, int t) {
      for (int64_t i = a; i < b; i++) {
        const double *p = flip_stat + i * d;
        double e = exp(dot(p));
        for (int k = 0; k < d; k++) partf[(size_t)t * d + k] += e * p[k];
      }
    });
    for (int k = 0; k < d; k++) {
      double S = 0, F = 0;
      for (int t = 0; t < nt; t++) { S += part[(size_t)t * d + k]; F += partf[(size_t)t * d + k]; }
      grad[k] = -sum_m[k] - c * S - sum_f[k] - F;
    }
  }
  // GetHessian
  if (hess) {
    std::vector<double> part((size_t)nt * d * d, 0.0), partf((size_t)nt * d * d, 0.0);
    parallel_for(n_s, nt, [&](int64_t a, int64_t b, int t) {
      double *o = part.data() + (size_t)t * d * d;
      for (int64_t i = a; i < b; i++) {
        const double *p = split_stat + i * d;
        double e = exp(dot(p));
        for (int k = 0; k < d; k++)
          for (int l = 0; l < d; l++) o[k * d + l] += e * p[k] * p[l];
      }
    });
    parallel_for(n_f, nt, [&](int64_t a, int64_t b, int t) {
      double *o = partf.data() + (size_t)t * d * d;
      for (int64_t i = a; i < b; i++) {
        const double *p = flip_stat + i * d;
        double e = exp(dot(p));
        for (int k = 0; k < d; k++)
          for (int l = 0; l < d; l++) o[k * d + l] += e * p[k] * p[l];
      }
    });
    for (int k = 0; k < d * d; k++) {
      double S = 0, F = 0;
      for (int t = 0; t < nt; t++) { S += part[(size_t)t * d * d + k]; F += partf[(size_t)t * d * d + k]; }
      hess[k] = -c * S - F;
    }
  }
  return 0;
}

/* One whole pass of the hot path on the CPU, for timing: SetEnergy (all
 * merges + all flips), AddSplits(n) with the Philox sampler, then value,
 * gradient and Hessian at theta.  Returns 0, fills lpl/grad/hess. */
int oracle_pl_pass(const lite_tess_soa *s, int d, const int32_t *fid, const double *theta,
                   int64_t n_splits, uint64_t seed, uint64_t ctr0, double *lpl, double *grad,
                   double *hess, int nthreads) {
  const int64_t nm = s->n_non_blocking, nf = 2 * (int64_t)s->n_blocking;
  std::vector<double> ms((size_t)std::max<int64_t>(nm, 1) * d), fs((size_t)std::max<int64_t>(nf, 1) * d);
  std::vector<double> ss((size_t)std::max<int64_t>(n_splits, 1) * d);
  std::vector<int32_t> he(n_splits);
  std::vector<double> x(n_splits), ang(n_splits);
  if (oracle_merge_stats(s, d, fid, nullptr, ms.data(), nthreads) < 0) return -1;
  if (oracle_flip_stats(s, d, fid, nullptr, nullptr, nullptr, nullptr, fs.data(), nthreads) < 0) return -1;
  if (oracle_sample_splits(s, n_splits, seed, ctr0, 0, n_splits, he.data(), x.data(), ang.data(), nthreads)) return -1;
  if (oracle_split_stats(s, d, fid, n_splits, he.data(), x.data(), ang.data(), nullptr, nullptr, ss.data(), nthreads) < 0)
    return -1;
  std::vector<double> sm(d, 0.0), sf(d, 0.0);
  for (int64_t i = 0; i < nm; i++)
    for (int k = 0; k < d; k++) sm[k] += ms[i * d + k];
  for (int64_t i = 0; i < nf; i++)
    for (int k = 0; k < d; k++) sf[k] += fs[i * d + k];
  return oracle_pl(d, theta, n_splits, ss.data(), nf, fs.data(), sm.data(), sf.data(), s->int_length,
                   n_splits, lpl, grad, hess, nthreads);
}

}  // extern "C"
