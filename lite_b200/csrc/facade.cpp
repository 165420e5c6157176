// facade.cpp — implementation of include/ttessel.h: the CGAL-free, source-compatible
// C++ API of LiTe's pseudo-likelihood path over the host model (host_tess.h) and
// the C ABI (include/lite_b200.h).  Every statistic variation and every
// value / gradient / Hessian is computed on the GPU; nothing here is a CPU
// fallback of the hot path.  Reference citations are relative to /root/reference.
#include "../../include/ttessel.h"

#include <iostream>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <sstream>

#include "host_tess.h"

CGAL::Random *rnd = 0;  // include/ttessel.h:139, lib/ttessel.cpp:21

using lite::HostTess;

static void ck(int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + lite_last_error());
}
static CGAL::Random &the_rnd() {
  if (!rnd) throw std::logic_error("the global generator `rnd` is not set (rnd = new CGAL::Random(seed))");
  return *rnd;
}

// ---------------------------------------------------------------------------
// kernel objects
// ---------------------------------------------------------------------------
std::ostream &operator<<(std::ostream &os, const Point2 &p) { return os << p.x() << " " << p.y(); }

double Polygon::area() const {
  double a = 0;
  Size n = v_.size();
  for (Size i = 0; i < n; i++) {
    const Point2 &u = v_[i], &w = v_[(i + 1) % n];
    a += (u.x() - v_[0].x()) * (w.y() - v_[0].y()) - (u.y() - v_[0].y()) * (w.x() - v_[0].x());
  }
  return 0.5 * a;
}
void Polygon::reverse_orientation() {
  if (v_.size() > 1) std::reverse(v_.begin() + 1, v_.end());  // CGAL keeps the first vertex
}

// ---------------------------------------------------------------------------
// LineTes
// ---------------------------------------------------------------------------
LineTes::LineTes() : impl_(new HostTess()), version_(0) {}
LineTes::~LineTes() { delete impl_; }
LineTes::LineTes(const LineTes &o) : impl_(new HostTess(*o.impl_)), window_(o.window_), version_(o.version_ + 1) {}
LineTes &LineTes::operator=(const LineTes &o) {
  if (this != &o) { *impl_ = *o.impl_; window_ = o.window_; version_++; }
  return *this;
}

Point2 LineTes::Vertex_handle::point() const { return Point2(t_->impl_->vx[id_], t_->impl_->vy[id_]); }
Size LineTes::Vertex_handle::degree() const {
  const HostTess &h = *t_->impl_;
  int h0 = h.v_inc[id_], e = h0;
  Size n = 0;
  do { n++; e = h.h_next[e] ^ 1; } while (e != h0);
  return n;
}
std::vector<LineTes::Halfedge_handle> LineTes::Vertex_handle::incident_halfedges() const {
  const HostTess &h = *t_->impl_;
  std::vector<Halfedge_handle> r;
  int h0 = h.v_inc[id_], e = h0;
  do { r.push_back(Halfedge_handle(t_, e)); e = h.h_next[e] ^ 1; } while (e != h0);
  return r;
}
std::vector<LineTes::Halfedge_handle> LineTes::Face_handle::holes() const {
  // inner boundaries: halfedges that carry the face's id and are not on its outer boundary
  const HostTess &h = *t_->impl_;
  std::vector<Halfedge_handle> r;
  if (!h.has_holes) return r;
  std::vector<char> seen(h.h_src.size(), 0);
  int e = h.f_he[id_];
  do { seen[e] = 1; e = h.h_next[e]; } while (e != h.f_he[id_]);
  for (size_t k = 0; k < h.h_src.size(); k++) {
    if (!h.h_alive[k] || seen[k] || h.h_face[k] != id_) continue;
    r.push_back(Halfedge_handle(t_, (int)k));
    int q = (int)k;
    do { seen[q] = 1; q = h.h_next[q]; } while (q != (int)k);
  }
  return r;
}
bool LineTes::Face_handle::data() const { return t_->impl_->f_inside[id_] != 0; }
LineTes::Halfedge_handle LineTes::Face_handle::outer_ccb() const { return Halfedge_handle(t_, t_->impl_->f_he[id_]); }
LineTes::Vertex_handle LineTes::Halfedge_handle::source() const { return Vertex_handle(t_, t_->impl_->h_src[id_]); }
LineTes::Vertex_handle LineTes::Halfedge_handle::target() const { return Vertex_handle(t_, t_->impl_->h_src[id_ ^ 1]); }
LineTes::Halfedge_handle LineTes::Halfedge_handle::twin() const { return Halfedge_handle(t_, id_ ^ 1); }
LineTes::Halfedge_handle LineTes::Halfedge_handle::next() const { return Halfedge_handle(t_, t_->impl_->h_next[id_]); }
LineTes::Halfedge_handle LineTes::Halfedge_handle::prev() const { return Halfedge_handle(t_, t_->impl_->h_prev[id_]); }
LineTes::Face_handle LineTes::Halfedge_handle::face() const { return Face_handle(t_, t_->impl_->h_face[id_]); }
double LineTes::Halfedge_handle::get_length() const { return t_->impl_->h_len[id_]; }
bool LineTes::Halfedge_handle::get_dir() const { return t_->impl_->h_dir[id_] != 0; }
LineTes::Halfedge_handle LineTes::Halfedge_handle::get_next_hf() const {
  int n = t_->impl_->h_next_hf[id_];
  return n < 0 ? Halfedge_handle() : Halfedge_handle(t_, n);
}
LineTes::Halfedge_handle LineTes::Halfedge_handle::get_prev_hf() const {
  int n = t_->impl_->h_prev_hf[id_];
  return n < 0 ? Halfedge_handle() : Halfedge_handle(t_, n);
}
LineTes::Seg LineTes::Halfedge_handle::segment() const { return Seg(t_, t_->impl_->h_seg[id_]); }

LineTes::Halfedge_handle LineTes::Seg::get_halfedge_handle() const { return Halfedge_handle(t_, t_->impl_->s_he[id_]); }
LineTes::Halfedge_handle LineTes::Seg::halfedges_start() const { return Halfedge_handle(t_, t_->impl_->seg_start(id_)); }
LineTes::Halfedge_handle LineTes::Seg::halfedges_end() const { return Halfedge_handle(t_, t_->impl_->seg_end(id_)); }
Size LineTes::Seg::number_of_edges() const { return (Size)t_->impl_->s_nedges[id_]; }
Point2 LineTes::Seg::pointSource() const { return halfedges_start().source().point(); }
Point2 LineTes::Seg::pointTarget() const { return halfedges_end().target().point(); }
std::vector<Point2> LineTes::Seg::list_of_points() const {  // lib/ttessel.cpp:864-873
  const HostTess &h = *t_->impl_;
  std::vector<Point2> lp;
  int e = h.seg_start(id_);
  lp.push_back(Point2(h.vx[h.h_src[e]], h.vy[h.h_src[e]]));
  while (e >= 0) {
    int v = h.h_src[e ^ 1];
    lp.push_back(Point2(h.vx[v], h.vy[v]));
    e = h.h_next_hf[e];
  }
  return lp;
}

void LineTes::insert_window(Rectangle r) {
  Polygon p;
  p.push_back(Point2(r.xmin(), r.ymin())); p.push_back(Point2(r.xmax(), r.ymin()));
  p.push_back(Point2(r.xmax(), r.ymax())); p.push_back(Point2(r.xmin(), r.ymax()));
  insert_window(p);
}
void LineTes::insert_window(Polygon &p) {
  if (p.size() < 3) throw std::domain_error("insert_window: a window needs at least three vertices");
  Polygon q(p);
  if (q.area() < 0) q.reverse_orientation();
  std::vector<double> x, y;
  for (Size i = 0; i < q.size(); i++) { x.push_back(q[i].x()); y.push_back(q[i].y()); }
  impl_->insert_window_polygon(x, y);
  window_.clear();
  window_.push_back(HPolygon(q));
  touch();
}
static void boundaries_of(const HPolygons &w, std::vector<std::vector<double>> &bx, std::vector<std::vector<double>> &by) {
  // outer boundaries counter-clockwise, holes clockwise (the window on the left of every side)
  for (size_t i = 0; i < w.size(); i++) {
    std::vector<Polygon> b;
    b.push_back(w[i].outer_boundary());
    if (b[0].area() < 0) b[0].reverse_orientation();
    for (HPolygon::Hole_const_iterator h = w[i].holes_begin(); h != w[i].holes_end(); ++h) {
      b.push_back(*h);
      if (b.back().area() > 0) b.back().reverse_orientation();
    }
    for (size_t k = 0; k < b.size(); k++) {
      bx.push_back(std::vector<double>()); by.push_back(std::vector<double>());
      for (Size j = 0; j < b[k].size(); j++) { bx.back().push_back(b[k][j].x()); by.back().push_back(b[k][j].y()); }
    }
  }
}
void LineTes::insert_window(HPolygons &w) {
  if (w.size() == 1 && w[0].number_of_holes() == 0) { Polygon p(w[0].outer_boundary()); insert_window(p); return; }
  std::vector<std::vector<double>> bx, by;
  boundaries_of(w, bx, by);
  std::string err;
  if (!lite::build_from_boundaries(*impl_, bx, by, std::vector<std::vector<double>>(), err))
    throw std::domain_error("insert_window: " + err);
  window_ = w;
  touch();
}
LineTes::Seg LineTes::insert_segment(Segment sg) {
  const double c[4] = {sg.source().x(), sg.source().y(), sg.target().x(), sg.target().y()};
  int id = -1;
  std::string err;
  if (!lite::insert_segment(*impl_, c, &id, err)) throw std::domain_error("insert_segment: " + err);
  touch();
  return Seg(this, id);
}
bool LineTes::is_a_T_tessellation(bool verbose, std::ostream &out) {
  if (impl_->arrangement_only && verbose) {
    double x = 0, y = 0;
    const int n = lite::count_non_t_vertices(*impl_, &x, &y);
    out << "not a T-tessellation: " << impl_->not_t_reason;
    if (n > 0) out << "; " << n << " vertices are not T-vertices, the first at " << Point2(x, y);
    out << std::endl;
  }
  return !impl_->arrangement_only;
}
void LineTes::remove_ivertices(Size imax, bool verbose) {
  std::string err;
  lite::CleanStats st;
  if (!lite::remove_ivertices(*impl_, imax, &st, err)) throw std::domain_error("remove_ivertices: " + err);
  if (verbose)
    std::cout << st.n_removed << " I-vertices removed" << std::endl
              << "removed length " << st.removed_length << std::endl
              << "added length " << st.added_length << std::endl;
  touch();
}
void LineTes::remove_lvertices(Size imax, bool verbose) {
  std::string err;
  lite::CleanStats st;
  if (!lite::remove_lvertices(*impl_, imax, &st, err)) throw std::domain_error("remove_lvertices: " + err);
  if (verbose)
    std::cout << st.n_removed << " L-vertices removed" << std::endl << "added length " << st.added_length << std::endl;
  touch();
}
void LineTes::remove_xvertices(double step) {
  std::string err;
  if (!lite::remove_xvertices(*impl_, step, 0, err)) throw std::domain_error("remove_xvertices: " + err);
  touch();
}
Size LineTes::number_of_xvertices(bool exclude_boundaries, bool exclude_ineighbours) {
  return (Size)lite::count_xvertices(*impl_, exclude_boundaries, exclude_ineighbours);
}
double LineTes::min_edge_length() {
  double m = HUGE_VAL;
  for (size_t h = 0; h < impl_->h_src.size(); h += 2) if (impl_->h_alive[h]) m = std::fmin(m, impl_->h_len[h]);
  return m;
}
Segment clip_segment_by_polygon(Segment sg, HPolygons w) {
  std::vector<std::vector<double>> bx, by;
  boundaries_of(w, bx, by);
  const double c[4] = {sg.source().x(), sg.source().y(), sg.target().x(), sg.target().y()};
  double o[4];
  std::string err;
  if (!lite::clip_segment(bx, by, c, o, err))
    throw std::domain_error("function clip_segment_by_polygon failed to return a segment: " + err);
  return Segment(Point2(o[0], o[1]), Point2(o[2], o[3]));
}
Segment clip_segment_by_polygon(Segment sg, HPolygon w) { return clip_segment_by_polygon(sg, HPolygons(1, w)); }
Segment clip_segment_by_polygon(Segment sg, Polygon w) { return clip_segment_by_polygon(sg, HPolygons(1, HPolygon(w))); }

Size LineTes::number_of_internal_segments() {
  Size n = 0;
  for (size_t i = 0; i < impl_->s_alive.size(); i++) n += impl_->s_alive[i] && !impl_->s_bnd[i];
  return n;
}
Size LineTes::number_of_window_edges() {
  Size n = 0;
  for (size_t i = 0; i < window_.size(); i++) {
    Polygons b = boundaries(window_[i]);
    for (size_t k = 0; k < b.size(); k++) n += b[k].size();
  }
  return n;
}
double LineTes::get_window_perimeter() {
  double per = 0;
  for (size_t i = 0; i < window_.size(); i++) {
    Polygons b = boundaries(window_[i]);
    for (size_t k = 0; k < b.size(); k++)
      for (Size j = 0; j < b[k].size(); j++) per += sqrt(Segment(b[k][j], b[k][(j + 1) % b[k].size()]).squared_length());
  }
  return per;
}
LineTes::Seg_list_iterator LineTes::segments_begin() {
  seg_cache_.clear();
  for (size_t i = 0; i < impl_->s_alive.size(); i++) if (impl_->s_alive[i]) seg_cache_.push_back(Seg(this, (int)i));
  return seg_cache_.begin();
}
LineTes::Seg_list_iterator LineTes::segments_end() { return seg_cache_.end(); }
bool LineTes::is_on_boundary(Seg s) const { return impl_->s_bnd[s.id()] != 0; }
std::vector<LineTes::Vertex_handle> LineTes::vertices() const {
  std::vector<Vertex_handle> r;
  for (size_t i = 0; i < impl_->v_alive.size(); i++) if (impl_->v_alive[i]) r.push_back(Vertex_handle(this, (int)i));
  return r;
}
std::vector<LineTes::Face_handle> LineTes::faces() const {
  std::vector<Face_handle> r;
  for (size_t i = 0; i < impl_->f_alive.size(); i++) if (impl_->f_alive[i]) r.push_back(Face_handle(this, (int)i));
  return r;
}
Size LineTes::number_of_vertices() const {
  Size n = 0;
  for (size_t i = 0; i < impl_->v_alive.size(); i++) n += impl_->v_alive[i];
  return n;
}
Size LineTes::number_of_halfedges() const { return (Size)impl_->n_alive_halfedges(); }
Size LineTes::number_of_faces() const {
  Size n = 0;
  for (size_t i = 0; i < impl_->f_alive.size(); i++) n += impl_->f_alive[i];
  return n;
}
Size LineTes::number_of_segments() const {
  Size n = 0;
  for (size_t i = 0; i < impl_->s_alive.size(); i++) n += impl_->s_alive[i];
  return n;
}
std::vector<LineTes::Halfedge_handle> LineTes::halfedges() const {
  std::vector<Halfedge_handle> r;
  for (size_t i = 0; i < impl_->h_alive.size(); i++)
    if (impl_->h_alive[i]) r.push_back(Halfedge_handle(this, (int)i));
  return r;
}
bool LineTes::is_on_boundary(Halfedge_handle e) const { return impl_->s_bnd[impl_->h_seg[e.id()]] != 0; }
bool LineTes::is_valid(bool verbose) {
  if (impl_->arrangement_only) {   // not a T-tessellation: the links of the arrangement only (:284-300)
    const HostTess &h = *impl_;
    for (size_t e = 0; e < h.h_src.size(); e++) {
      if (!h.h_alive[e]) continue;
      const int n = h.h_next[e], nf = h.h_next_hf[e];
      const bool ok = n >= 0 && h.h_prev[n] == (int)e && h.h_src[n] == h.h_src[e ^ 1] && h.h_face[n] == h.h_face[e] &&
                      (nf < 0 || (h.h_prev_hf[nf] == (int)e && h.h_seg[nf] == h.h_seg[e] && h.h_src[nf] == h.h_src[e ^ 1]));
      if (!ok) {
        if (verbose) std::clog << "LineTes::is_valid: broken links at halfedge " << e << std::endl;
        return false;
      }
    }
    return true;
  }
  std::string e = impl_->check();
  if (!e.empty() && verbose) std::clog << "LineTes::is_valid: " << e << std::endl;
  return e.empty();
}
void LineTes::write(std::ostream &os) const { os << lite::write_text(*impl_); }
void LineTes::read(std::istream &is) {
  std::stringstream ss;
  ss << is.rdbuf();
  std::string err;
  if (!lite::read_text(*impl_, ss.str(), err, /*accept_any=*/true)) throw std::domain_error("LineTes::read: " + err);
  // the window as the file gave it: a counter-clockwise boundary opens a polygon, the clockwise
  // ones that follow are its holes (lib/ttessel.cpp:1395-1415)
  window_.clear();
  Polygon outer;
  Polygons holes;
  for (size_t b = 0; b < impl_->win_bx.size(); b++) {
    Polygon q;
    for (size_t i = 0; i < impl_->win_bx[b].size(); i++) q.push_back(Point2(impl_->win_bx[b][i], impl_->win_by[b][i]));
    if (q.area() > 0) {
      if (!outer.is_empty()) window_.push_back(HPolygon(outer, holes.begin(), holes.end()));
      outer = q; holes.clear();
    } else {
      holes.push_back(q);
    }
  }
  if (!outer.is_empty()) window_.push_back(HPolygon(outer, holes.begin(), holes.end()));
  touch();
}
void LineTes::print_all_segments() const {
  for (size_t s = 0; s < impl_->s_alive.size(); s++) {
    if (!impl_->s_alive[s]) continue;
    Seg sg(this, (int)s);
    std::cout << sg.pointSource() << " " << sg.pointTarget() << std::endl;
  }
}

// ---------------------------------------------------------------------------
// TTessel
// ---------------------------------------------------------------------------
TTessel::TTessel() {}
TTessel::TTessel(LineTes &lt) : LineTes(lt) {  // lib/ttessel.cpp:1506-1540
  if (!lt.is_a_T_tessellation()) throw std::domain_error("TTessel(LinTes&) input is not a T-tesselation");
}
void TTessel::insert_window(Rectangle r) { LineTes::insert_window(r); }
void TTessel::insert_window(Polygon &p) { LineTes::insert_window(p); }
void TTessel::insert_window(HPolygons &w) { LineTes::insert_window(w); }
void TTessel::clear(bool remove_window) {
  if (remove_window || window_.empty()) { *impl_ = HostTess(); window_.clear(); touch(); return; }
  HPolygons w(window_);
  LineTes::insert_window(w);
}
double TTessel::get_total_internal_length() { return impl_->int_length; }
Size TTessel::number_of_blocking_segments() { return impl_->b_list.size(); }
Size TTessel::number_of_non_blocking_segments() { return impl_->nb_list.size(); }
void TTessel::refresh_sublists() {
  blocking_cache_.clear(); non_blocking_cache_.clear();
  for (size_t i = 0; i < impl_->b_list.size(); i++) blocking_cache_.push_back(Seg(this, impl_->b_list[i]));
  for (size_t i = 0; i < impl_->nb_list.size(); i++) non_blocking_cache_.push_back(Seg(this, impl_->nb_list[i]));
}
TTessel::Seg_sublist_iterator TTessel::blocking_segments_begin() { refresh_sublists(); return blocking_cache_.begin(); }
TTessel::Seg_sublist_iterator TTessel::blocking_segments_end() { return blocking_cache_.end(); }
TTessel::Seg_sublist_iterator TTessel::non_blocking_segments_begin() { refresh_sublists(); return non_blocking_cache_.begin(); }
TTessel::Seg_sublist_iterator TTessel::non_blocking_segments_end() { return non_blocking_cache_.end(); }
bool TTessel::is_valid(bool verbose) { return LineTes::is_valid(verbose); }
void TTessel::print_blocking_segments() {
  for (size_t i = 0; i < impl_->b_list.size(); i++) {
    Seg s(this, impl_->b_list[i]);
    std::cout << s.pointSource() << " " << s.pointTarget() << std::endl;
  }
}
void TTessel::print_non_blocking_segments() {
  for (size_t i = 0; i < impl_->nb_list.size(); i++) {
    Seg s(this, impl_->nb_list[i]);
    std::cout << s.pointSource() << " " << s.pointTarget() << std::endl;
  }
}
void TTessel::simulate_crtt(double tau, unsigned long n_steps, std::uint64_t seed) {
  lite::crtt_chain(*impl_, tau, n_steps, seed);
  touch();
}

static Polygon face_polygon(const HostTess &h, int f, bool simplify) {  // face2poly, lib/ttessel.cpp:4314-4336
  Polygon p;
  int h0 = h.f_he[f], e = h0;
  do {
    if (!simplify || h.h_next_hf[e] < 0 || h.h_next_hf[e] != h.h_next[e]) {
      int v = h.h_src[e ^ 1];
      p.push_back(Point2(h.vx[v], h.vy[v]));
    }
    e = h.h_next[e];
  } while (e != h0);
  return p;
}
// ---- random halfedges (lib/ttessel.cpp:1932-2010), from the global `rnd` ----
TTessel::Halfedge_handle TTessel::random_halfedge() {
  std::vector<Halfedge_handle> hs = halfedges();
  for (;;) {
    Halfedge_handle e = hs[(size_t)the_rnd().get_int(0, (int)hs.size())];
    if (e->face()->data()) return e;
  }
}
TTessel::Halfedge_handle TTessel::length_weighted_random_halfedge() {  // rejection, :1932-1946
  const double til = total_internal_length();
  for (;;) {
    Halfedge_handle e = random_halfedge();
    if (the_rnd().get_double(0, til) < e->get_length()) return e;
  }
}
TTessel::Halfedge_handle TTessel::alt_length_weighted_random_halfedge() {  // :1948-1966
  const double select = the_rnd().get_double(0, total_internal_length());
  double cumlen = 0;
  const HostTess &h = *impl_;
  for (size_t e = 0; e < h.h_src.size(); e++) {
    if (!h.h_alive[e] || !h.f_inside[h.h_face[e]]) continue;
    cumlen += h.h_len[e];
    if (cumlen > select) return Halfedge_handle(this, (int)e);
  }
  return Halfedge_handle();
}
std::vector<TTessel::Halfedge_handle> TTessel::alt_length_weighted_random_halfedge(unsigned int n) {  // :1968-1991
  const double til = total_internal_length();
  std::vector<double> select(n);
  for (unsigned int i = 0; i != n; i++) select[i] = the_rnd().get_double(0, til);
  std::sort(select.begin(), select.end(), std::greater<double>());
  std::vector<Halfedge_handle> res;
  double cumlen = 0;
  const HostTess &h = *impl_;
  for (size_t e = 0; e < h.h_src.size() && !select.empty(); e++) {
    if (!h.h_alive[e] || !h.f_inside[h.h_face[e]]) continue;
    cumlen += h.h_len[e];
    while (!select.empty() && cumlen > select.back()) { res.push_back(Halfedge_handle(this, (int)e)); select.pop_back(); }
  }
  return res;
}
void TTessel::printRCALI(std::ostream &out) {  // :2012-2060
  const HostTess &h = *impl_;
  out << impl_->n_cells() << std::endl;
  int counter = 0;
  for (size_t f = 0; f < h.f_alive.size(); f++) {
    if (!h.f_alive[f] || !h.f_inside[f]) continue;
    counter++;
    // the corners after the first halfedge's target, which is always printed (as written)
    std::vector<int> vs;
    int e0 = h.f_he[f], e = e0, corners = 0;
    do {
      const bool corner = h.h_next_hf[e] < 0 || h.h_next_hf[e] != h.h_next[e];
      if (corner) corners++;
      if (e == e0 || corner) vs.push_back(h.h_src[e ^ 1]);
      e = h.h_next[e];
    } while (e != e0);
    out << counter << " " << counter << " " << corners << std::endl;
    for (size_t k = 0; k < vs.size(); k++) out << (k ? " " : "") << h.vx[vs[k]];
    out << std::endl;
    for (size_t k = 0; k < vs.size(); k++) out << (k ? " " : "") << h.vy[vs[k]];
    out << std::endl;
  }
}
HPolygons TTessel::all_faces() {
  HPolygons r;
  for (size_t f = 0; f < impl_->f_alive.size(); f++)
    if (impl_->f_alive[f] && impl_->f_inside[f]) r.push_back(HPolygon(face_polygon(*impl_, (int)f, true)));
  return r;
}

ModList TTessel::Modification::modified_elements() { return ModList(); }

// ---- Split (lib/ttessel.cpp:2090-2132) ----
TTessel::Split::Split() : x_(0), angle_(0) {}
TTessel::Split::Split(Halfedge_handle e, double x, double angle) : e1(e), x_(x), angle_(angle) {
  const LineTes *lt = e.tess();
  if (!lt || e.is_null()) return;
  lite::SplitMove m = lt->host()->make_split(e.id(), x, angle);
  if (!m.valid) {  // reference: message + null handles (:2112-2117, :4571-4574)
    std::cerr << "Split constructor: l does not hit e1 or ray_exit_face found no intersection" << std::endl;
    e2 = Halfedge_handle();
    return;
  }
  e2 = Halfedge_handle(lt, m.e2);
  p1 = Point2(m.p1x, m.p1y);
  p2 = Point2(m.p2x, m.p2y);
}
bool TTessel::Split::is_valid() { return !e1.is_null() && !e2.is_null() && e1 != e2; }

static const LineTes *tess_of(const LineTes::Halfedge_handle &e) { return e.tess(); }
static Point2 pt(const HostTess &h, int v) { return Point2(h.vx[v], h.vy[v]); }
static bool flat_at_source(const HostTess &h, int e) {  // the source of e is not a corner of face(e)
  int p = h.h_prev[e];
  return h.h_next_hf[p] >= 0 && h.h_next_hf[p] == e;
}
// corners of the face met when walking its halfedges from next(from) to `to` inclusive
// (the source vertex of each halfedge)
static void append_chain(const HostTess &h, int from, int to, std::vector<Point2> &out) {
  for (int e = h.h_next[from];; e = h.h_next[e]) {
    if (!flat_at_source(h, e)) out.push_back(pt(h, h.h_src[e]));
    if (e == to) break;
  }
}

ModList TTessel::Split::modified_elements() {  // lib/ttessel.cpp:2135-2232, hole-free faces
  ModList m;
  if (!is_valid()) return m;
  const HostTess &h = *tess_of(e1)->host();
  int a = e1.id(), b = e2.id();
  Point2 s1 = pt(h, h.h_src[a]), t1 = pt(h, h.h_src[a ^ 1]), s2 = pt(h, h.h_src[b]), t2 = pt(h, h.h_src[b ^ 1]);
  m.add_vertices.push_back(p1); m.add_vertices.push_back(p2);
  m.del_edges.push_back(Segment(s1, t1)); m.del_edges.push_back(Segment(s2, t2));
  m.add_edges.push_back(Segment(s1, p1)); m.add_edges.push_back(Segment(p1, t1));
  m.add_edges.push_back(Segment(s2, p2)); m.add_edges.push_back(Segment(p2, t2));
  m.add_edges.push_back(Segment(p1, p2));
  m.del_faces.push_back(HPolygon(face_polygon(h, h.h_face[a], true)));
  // C1 = [p1, p2, tgt(e2) .. src(e1)],  C2 = [p2, p1, tgt(e1) .. src(e2)]  (:4603-4616), simplified
  std::vector<Point2> c1, c2;
  c1.push_back(p1); c1.push_back(p2);
  append_chain(h, b, a, c1);
  c2.push_back(p2); c2.push_back(p1);
  append_chain(h, a, b, c2);
  m.add_faces.push_back(HPolygon(Polygon(c1.begin(), c1.end())));
  m.add_faces.push_back(HPolygon(Polygon(c2.begin(), c2.end())));
  const int es[2] = {a, b};
  const Point2 ps[2] = {p1, p2};
  for (int k = 0; k < 2; k++) {
    std::vector<Point2> seg = Seg(tess_of(e1), h.h_seg[es[k]]).list_of_points();
    m.del_segs.push_back(seg);
    Point2 key = h.h_dir[es[k]] ? pt(h, h.h_src[es[k] ^ 1]) : pt(h, h.h_src[es[k]]);
    std::vector<Point2>::iterator it = std::find(seg.begin(), seg.end(), key);
    seg.insert(it, ps[k]);
    m.add_segs.push_back(seg);
  }
  std::vector<Point2> ns;
  ns.push_back(p1); ns.push_back(p2);
  m.add_segs.push_back(ns);
  return m;
}

// ---- Merge (:2299-2404) ----
TTessel::Merge::Merge() {}
TTessel::Merge::Merge(Halfedge_handle he) : e(he) {}
bool TTessel::Merge::is_valid() {
  if (e.is_null()) return false;
  const HostTess &h = *tess_of(e)->host();
  int s = h.h_seg[e.id()];
  return h.s_nedges[s] == 1 && !h.s_bnd[s];
}
ModList TTessel::Merge::modified_elements() {
  ModList m;
  if (!is_valid()) return m;
  const HostTess &h = *tess_of(e)->host();
  int a = e.id(), ta = a ^ 1;
  int e1 = h.h_next[ta], e2 = h.h_next[a];
  Point2 p1 = pt(h, h.h_src[a]), p2 = pt(h, h.h_src[ta]);
  Point2 X2 = pt(h, h.h_src[e2 ^ 1]), Y2 = pt(h, h.h_src[h.h_prev_hf[e2]]);
  Point2 X1 = pt(h, h.h_src[e1 ^ 1]), Y1 = pt(h, h.h_src[h.h_prev_hf[e1]]);
  m.del_vertices.push_back(p1); m.del_vertices.push_back(p2);
  m.del_edges.push_back(Segment(p1, p2));
  m.del_edges.push_back(Segment(p2, X2)); m.del_edges.push_back(Segment(Y2, p2));
  m.del_edges.push_back(Segment(p1, X1)); m.del_edges.push_back(Segment(Y1, p1));
  m.add_edges.push_back(Segment(Y2, X2)); m.add_edges.push_back(Segment(Y1, X1));
  m.del_faces.push_back(HPolygon(face_polygon(h, h.h_face[a], true)));
  m.del_faces.push_back(HPolygon(face_polygon(h, h.h_face[ta], true)));
  // merged polygon (:4945-4966): face(e) after p2 round to p1, face(twin e) after p1 round to p2;
  // p1 and p2 become flat
  std::vector<Point2> mp;
  for (int k = h.h_next[e2]; k != a; k = h.h_next[k])
    if (!flat_at_source(h, k)) mp.push_back(pt(h, h.h_src[k]));
  for (int k = h.h_next[e1]; k != ta; k = h.h_next[k])
    if (!flat_at_source(h, k)) mp.push_back(pt(h, h.h_src[k]));
  m.add_faces.push_back(HPolygon(Polygon(mp.begin(), mp.end())));
  const LineTes *lt = tess_of(e);
  std::vector<Point2> seg = Seg(lt, h.h_seg[e1]).list_of_points();
  m.del_segs.push_back(seg);
  seg.erase(std::find(seg.begin(), seg.end(), p1));
  m.add_segs.push_back(seg);
  seg = Seg(lt, h.h_seg[e2]).list_of_points();
  m.del_segs.push_back(seg);
  seg.erase(std::find(seg.begin(), seg.end(), p2));
  m.add_segs.push_back(seg);
  std::vector<Point2> gone;
  gone.push_back(p1); gone.push_back(p2);
  m.del_segs.push_back(gone);
  return m;
}

// ---- Flip (:2432-2713) ----
TTessel::Flip::Flip() : right(false) {}
TTessel::Flip::Flip(Halfedge_handle he) : e1(he), right(false) {
  const LineTes *lt = tess_of(he);
  lite::FlipMove f = lt->host()->make_flip(he.id());
  right = f.right;
  if (!f.valid) {
    std::cerr << "ray_exit_face: intersection of ray with face boundaries not found" << std::endl;
    return;
  }
  e2 = Halfedge_handle(lt, f.e2);
  p2 = Point2(f.p2x, f.p2y);
}
ModList TTessel::Flip::modified_elements() {
  ModList m;
  if (e1.is_null() || e2.is_null()) return m;
  const LineTes *lt = tess_of(e1);
  const HostTess &h = *lt->host();
  int a1 = e1.id(), b = e2.id();
  int n1 = h.h_next[a1], ph = h.h_prev_hf[n1];
  Point2 a = pt(h, h.h_src[a1]), q = pt(h, h.h_src[a1 ^ 1]);
  Point2 X = pt(h, h.h_src[n1 ^ 1]), Y = pt(h, h.h_src[ph]);
  Point2 S2 = pt(h, h.h_src[b]), E2 = pt(h, h.h_src[b ^ 1]);
  m.del_vertices.push_back(q);
  m.add_vertices.push_back(p2);
  m.del_edges.push_back(Segment(a, q)); m.del_edges.push_back(Segment(q, X)); m.del_edges.push_back(Segment(q, Y));
  bool tgt_is_q = h.h_src[b ^ 1] == h.h_src[a1 ^ 1], src_is_q = h.h_src[b] == h.h_src[a1 ^ 1];
  m.add_edges.push_back(Segment(a, p2));
  if (!tgt_is_q && !src_is_q) {  // :2484-2512
    m.del_edges.push_back(Segment(E2, S2));
    m.add_edges.push_back(Segment(p2, E2)); m.add_edges.push_back(Segment(p2, S2)); m.add_edges.push_back(Segment(X, Y));
  } else if (tgt_is_q) {
    m.add_edges.push_back(Segment(p2, S2)); m.add_edges.push_back(Segment(p2, X));
  } else {
    m.add_edges.push_back(Segment(p2, E2)); m.add_edges.push_back(Segment(p2, Y));
  }
  // faces (:2515-2655), generic case: R = face the ray enters, L = the other face of e1;
  // -{R, L} +{R2, R1 u L}
  int fR = right ? h.h_face[a1 ^ 1] : h.h_face[a1], fL = right ? h.h_face[a1] : h.h_face[a1 ^ 1];
  m.del_faces.push_back(HPolygon(face_polygon(h, fR, true)));
  m.del_faces.push_back(HPolygon(face_polygon(h, fL, true)));
  std::vector<Point2> r2, rl;
  if (right) {
    // R = face(twin e1).  R2 = [p2, a, ... vertices of R after a round to src(e2)]
    int t1 = a1 ^ 1;
    r2.push_back(p2); r2.push_back(a);
    for (int k = h.h_next[h.h_next[t1]]; ; k = h.h_next[k]) {
      if (!flat_at_source(h, k)) r2.push_back(pt(h, h.h_src[k]));
      if (k == b) break;
    }
    // R1 u L: L after q round to a (flat), p2, R from tgt(e2) round to q (flat)
    for (int k = h.h_next[n1]; k != a1; k = h.h_next[k])
      if (!flat_at_source(h, k)) rl.push_back(pt(h, h.h_src[k]));
    rl.push_back(p2);
    for (int k = h.h_next[b]; k != t1; k = h.h_next[k])
      if (!flat_at_source(h, k)) rl.push_back(pt(h, h.h_src[k]));
  } else {
    // R = face(e1).  R2 = [a, p2, vertices of R from tgt(e2) round to before a]
    r2.push_back(a); r2.push_back(p2);
    for (int k = h.h_next[b]; k != a1; k = h.h_next[k])
      if (!flat_at_source(h, k)) r2.push_back(pt(h, h.h_src[k]));
    int t1 = a1 ^ 1;
    if (b != n1)  // e2 leaving q: R1 is the triangle (a, q, p2) and adds no vertex
      for (int k = h.h_next[n1]; ; k = h.h_next[k]) {
        if (!flat_at_source(h, k)) rl.push_back(pt(h, h.h_src[k]));
        if (k == b) break;
      }
    rl.push_back(p2);
    for (int k = h.h_next[h.h_next[t1]]; k != t1; k = h.h_next[k])
      if (!flat_at_source(h, k)) rl.push_back(pt(h, h.h_src[k]));
  }
  m.add_faces.push_back(HPolygon(Polygon(r2.begin(), r2.end())));
  m.add_faces.push_back(HPolygon(Polygon(rl.begin(), rl.end())));
  // segments (:2659-2712)
  lite::FlipMove f = h.make_flip(a1);
  std::vector<Point2> seg = Seg(lt, h.h_seg[a1]).list_of_points();
  m.del_segs.push_back(seg);
  if (h.h_dir[a1]) seg.pop_back(); else seg.erase(seg.begin());
  m.add_segs.push_back(seg);
  seg = Seg(lt, h.h_seg[f.e3]).list_of_points();
  m.del_segs.push_back(seg);
  if (h.h_dir[f.e3]) seg.push_back(p2); else seg.insert(seg.begin(), p2);
  m.add_segs.push_back(seg);
  seg = Seg(lt, h.h_seg[b]).list_of_points();
  m.del_segs.push_back(seg);
  Point2 key = h.h_dir[b] ? E2 : S2;
  seg.insert(std::find(seg.begin(), seg.end(), key), p2);
  if (h.h_seg[n1] == h.h_seg[b]) {
    seg.erase(std::find(seg.begin(), seg.end(), q));
    m.add_segs.push_back(seg);
    return m;
  }
  m.add_segs.push_back(seg);
  seg = Seg(lt, h.h_seg[n1]).list_of_points();
  m.del_segs.push_back(seg);
  seg.erase(std::find(seg.begin(), seg.end(), q));
  m.add_segs.push_back(seg);
  return m;
}

// ---- moves ----
static void require_t(const HostTess &h, const char *what, bool moving) {
  if (h.arrangement_only)
    throw std::domain_error(std::string(what) + ": the inserted segments do not form a T-tessellation (" + h.not_t_reason + ")");
  if (moving && h.has_holes)
    throw std::domain_error(std::string(what) + ": the host model does not move tessellations of windows with holes");
}
TTessel::Halfedge_handle TTessel::update(Split s) {  // lib/ttessel.cpp:1580-1597
  require_t(*impl_, "TTessel::update(Split)", true);
  if (!s.is_valid()) throw std::domain_error("TTessel::update(Split): invalid split");
  lite::SplitMove m;
  m.e1 = s.get_e1().id(); m.e2 = s.get_e2().id();
  m.p1x = s.get_p1().x(); m.p1y = s.get_p1().y(); m.p2x = s.get_p2().x(); m.p2y = s.get_p2().y();
  m.valid = true;
  int e = impl_->update_split(m);
  touch();
  return Halfedge_handle(this, e);
}
TTessel::Face_handle TTessel::update(Merge m) {  // :1604-1638
  require_t(*impl_, "TTessel::update(Merge)", true);
  if (!m.is_valid()) throw std::domain_error("TTessel::update(Merge): not a non-blocking segment");
  int e = m.get_e().id();
  const HostTess &h = *impl_;
  // halfedges around the two faces; merge_edge keeps one id of each merged pair, so
  // whichever of these is still alive afterwards bounds the merged face
  int cand[6] = {h.h_next[h.h_next[e]], h.h_next[h.h_next[e ^ 1]], h.h_next[e], h.h_prev[e],
                 h.h_next[e ^ 1], h.h_prev[e ^ 1]};
  impl_->update_merge(e);
  touch();
  for (int k = 0; k < 6; k++)
    if (cand[k] != e && cand[k] != (e ^ 1) && impl_->h_alive[cand[k]]) return Face_handle(this, impl_->h_face[cand[k]]);
  return Face_handle();
}
TTessel::Halfedge_handle TTessel::update(Flip f) {  // :1644-1701
  require_t(*impl_, "TTessel::update(Flip)", true);
  if (f.get_e1().is_null() || f.get_e2().is_null()) throw std::domain_error("TTessel::update(Flip): invalid flip");
  lite::FlipMove m = impl_->make_flip(f.get_e1().id());
  if (!m.valid) throw std::domain_error("TTessel::update(Flip): invalid flip");
  int e = impl_->update_flip(m);
  touch();
  return Halfedge_handle(this, e);
}
TTessel::Split TTessel::propose_split() {  // :1755-1758, :1948-1966
  CGAL::Random &r = the_rnd();
  int e = impl_->sample_halfedge(r.get_double(0, 1));
  double U = r.get_double(0, 1), x = r.get_double(0, 1);  // g++ evaluates the arguments right to left
  return Split(Halfedge_handle(this, e), x, acos(1 - 2 * U));
}
TTessel::Merge TTessel::propose_merge() {  // :1760-1764
  int i = the_rnd().get_int(0, (int)impl_->nb_list.size());
  return Merge(Halfedge_handle(this, impl_->s_he[impl_->nb_list[i]]));
}
TTessel::Flip TTessel::propose_flip() {  // :1766-1777
  CGAL::Random &r = the_rnd();
  int i = r.get_int(0, (int)impl_->b_list.size());
  int side = r.get_int(0, 2);
  return Flip(Halfedge_handle(this, impl_->flip_halfedge(impl_->b_list[i], side)));
}
TTessel::Split_list TTessel::split_sample(Size n) {  // :1783-1795, :1968-1991
  CGAL::Random &r = the_rnd();
  std::vector<double> sel(n);
  for (Size i = 0; i < n; i++) sel[i] = r.get_double(0, 1);
  std::sort(sel.begin(), sel.end());  // emitted in halfedge container order, like the reference's scan
  Split_list out;
  out.reserve(n);
  for (Size i = 0; i < n; i++) {
    int e = impl_->sample_halfedge(sel[i]);
    double U = r.get_double(0, 1), x = r.get_double(0, 1);
    out.push_back(Split(Halfedge_handle(this, e), x, acos(1 - 2 * U)));
  }
  return out;
}
TTessel::Split_list TTessel::poisson_splits(double lambda) {  // :1804-1818 (Knuth's product method)
  CGAL::Random &r = the_rnd();
  double L = exp(-lambda), p = 1.0;
  Size k = 0;
  do { k++; p *= r.get_double(0, 1); } while (p > L);
  return split_sample(k - 1);
}
TTessel::Merge_list TTessel::all_merges() {  // :1822-1829
  Merge_list out;
  for (size_t i = 0; i < impl_->nb_list.size(); i++)
    out.push_back(Merge(Halfedge_handle(this, impl_->s_he[impl_->nb_list[i]])));
  return out;
}
TTessel::Flip_list TTessel::all_flips() {  // :1834-1846: flips[i] = twin(start), flips[i+n] = end
  size_t n = impl_->b_list.size();
  Flip_list out(2 * n);
  for (size_t i = 0; i < n; i++) {
    out[i] = Flip(Halfedge_handle(this, impl_->flip_halfedge(impl_->b_list[i], 0)));
    out[i + n] = Flip(Halfedge_handle(this, impl_->flip_halfedge(impl_->b_list[i], 1)));
  }
  return out;
}

// ---------------------------------------------------------------------------
// CatVector helpers (lib/ttessel.cpp:2737-2846)
// ---------------------------------------------------------------------------
void fill(CatVector &v, FVector &x) {
  Size j = 0;
  for (Size i = 0; i < v.vertices.size(); i++, j++) v.vertices[i] = x[j];
  for (Size i = 0; i < v.edges.size(); i++, j++) v.edges[i] = x[j];
  for (Size i = 0; i < v.faces.size(); i++, j++) v.faces[i] = x[j];
  for (Size i = 0; i < v.segs.size(); i++, j++) v.segs[i] = x[j];
}
std::vector<double> asVectorOfDoubles(const CatVector &v) {
  std::vector<double> r;
  r.insert(r.end(), v.vertices.begin(), v.vertices.end());
  r.insert(r.end(), v.edges.begin(), v.edges.end());
  r.insert(r.end(), v.faces.begin(), v.faces.end());
  r.insert(r.end(), v.segs.begin(), v.segs.end());
  return r;
}
FVector asFVector(const CatVector &v) { return asVectorOfDoubles(v); }
FMatrix asFMatrix(const CatMatrix &m) {
  FMatrix r;
  for (Size i = 0; i < m.vertices.size(); i++) r.push_back(asVectorOfDoubles(m.vertices[i]));
  for (Size i = 0; i < m.edges.size(); i++) r.push_back(asVectorOfDoubles(m.edges[i]));
  for (Size i = 0; i < m.faces.size(); i++) r.push_back(asVectorOfDoubles(m.faces[i]));
  for (Size i = 0; i < m.segs.size(); i++) r.push_back(asVectorOfDoubles(m.segs[i]));
  return r;
}
std::ostream &operator<<(std::ostream &os, const CatVector &v) {
  std::vector<double> x = asVectorOfDoubles(v);
  for (Size i = 0; i < x.size(); i++) os << (i ? " " : "") << x[i];
  return os;
}
double sum(const CatVector &v) {
  double s = 0;
  std::vector<double> x = asVectorOfDoubles(v);
  for (Size i = 0; i < x.size(); i++) s += x[i];
  return s;
}
CatMatrix outer(const CatVector &a, const CatVector &b) {
  CatMatrix m;
  for (Size i = 0; i < a.vertices.size(); i++) m.vertices.push_back(a.vertices[i] * b);
  for (Size i = 0; i < a.edges.size(); i++) m.edges.push_back(a.edges[i] * b);
  for (Size i = 0; i < a.faces.size(); i++) m.faces.push_back(a.faces[i] * b);
  for (Size i = 0; i < a.segs.size(); i++) m.segs.push_back(a.segs[i] * b);
  return m;
}
static void shape_like(const CatVector &like, const double *flat, CatVector *out) {
  *out = like;
  Size j = 0;
  for (Size i = 0; i < out->vertices.size(); i++) out->vertices[i] = flat[j++];
  for (Size i = 0; i < out->edges.size(); i++) out->edges[i] = flat[j++];
  for (Size i = 0; i < out->faces.size(); i++) out->faces[i] = flat[j++];
  for (Size i = 0; i < out->segs.size(); i++) out->segs[i] = flat[j++];
}

// ---------------------------------------------------------------------------
// feature functions on the host, in doubles (lib/ttessel.cpp:4124-4434, as written).
// They exist so that a caller may evaluate them and so that Energy can recognise
// them by address; the pseudo-likelihood path evaluates their device twins.
// ---------------------------------------------------------------------------
static double window_tol(const Polygon &w) {
  double m = 1.0;
  for (Size i = 0; i < w.size(); i++) m = std::max(m, std::max(fabs(w[i].x()), fabs(w[i].y())));
  return 64 * 2.220446049250313e-16 * m;
}
static bool on_seg(const Point2 &p, const Point2 &a, const Point2 &b, double tol) {
  double ex = b.x() - a.x(), ey = b.y() - a.y(), l = sqrt(ex * ex + ey * ey);
  if (l == 0) return false;
  if (fabs(ex * (p.y() - a.y()) - ey * (p.x() - a.x())) > tol * l) return false;
  double t = ((p.x() - a.x()) * ex + (p.y() - a.y()) * ey) / l;
  return t >= -tol && t <= l + tol;
}
// ---- helpers of the reference header ----
bool are_aligned(Point2 p, Point2 q, Point2 r, bool verbose) {  // :3556-3577
  const double cr = (q.x() - p.x()) * (r.y() - p.y()) - (q.y() - p.y()) * (r.x() - p.x());
  if (cr != 0.0) { if (verbose) std::clog << "are_aligned function: points are not collinear" << std::endl; return false; }
  if (p == q || q == r) { if (verbose) std::clog << "are_aligned function: the intermediate point is equal to an end point" << std::endl; return false; }
  if ((p.x() - q.x()) * (r.x() - q.x()) + (p.y() - q.y()) * (r.y() - q.y()) > 0) {
    if (verbose) std::clog << "are_aligned function: the intermediate point is not between the end points" << std::endl;
    return false;
  }
  return true;
}
// +1 strictly inside, 0 on the boundary, -1 outside (CGAL::bounded_side_2)
static int bounded_side(const Polygon &poly, const Point2 &pt) {
  bool in = false;
  const Size n = poly.size();
  for (Size i = 0; i < n; i++) {
    const Point2 &a = poly[i], &b = poly[(i + 1) % n];
    const double cr = (b.x() - a.x()) * (pt.y() - a.y()) - (b.y() - a.y()) * (pt.x() - a.x());
    if (cr == 0.0 && std::min(a.x(), b.x()) <= pt.x() && pt.x() <= std::max(a.x(), b.x()) &&
        std::min(a.y(), b.y()) <= pt.y() && pt.y() <= std::max(a.y(), b.y()))
      return 0;
    if ((a.y() > pt.y()) != (b.y() > pt.y())) {
      const double xc = a.x() + (pt.y() - a.y()) * (b.x() - a.x()) / (b.y() - a.y());
      if (pt.x() < xc) in = !in;
    }
  }
  return in ? 1 : -1;
}
bool is_inside(Point2 pt, HPolygon &poly) {  // :3583-3598
  if (bounded_side(poly.outer_boundary(), pt) != 1) return false;
  for (HPolygon::Hole_const_iterator h = poly.holes_begin(); h != poly.holes_end(); ++h)
    if (bounded_side(*h, pt) != -1) return false;
  return true;
}
bool is_inside(Point2 pt, HPolygons &polys) {
  for (Size i = 0; i != polys.size(); i++) if (is_inside(pt, polys[i])) return true;
  return false;
}
Segment clip_segment_by_convex_polygon(Segment sg, Polygon w) { return clip_segment_by_polygon(sg, w); }
LineTes::Halfedge_handle find_halfedge(LineTes &t, Point2 p1, Point2 p2) {  // :3955-3969
  std::vector<LineTes::Halfedge_handle> hs = t.halfedges();
  for (size_t i = 0; i < hs.size(); i++)
    if (hs[i]->source()->point() == p1 && hs[i]->target()->point() == p2) return hs[i];
  return NULL_HALFEDGE_HANDLE;
}
bool is_a_T_vertex(LineTes::Vertex_handle v, bool verbose) {  // :4002-4046
  if (v->degree() != 3) {
    if (verbose) std::clog << "is_a_T_vertex function: vertex " << v->point() << " not a T-vertex because its degree not equal to 3" << std::endl;
    return false;
  }
  std::vector<LineTes::Halfedge_handle> in = v->incident_halfedges();
  int n = 0;
  for (size_t i = 0; i < in.size(); i++) n += in[i]->get_next_hf() == NULL_HALFEDGE_HANDLE;
  if (n != 1) {
    if (verbose) std::clog << "is_a_T_vertex function: vertex " << v->point() << " not a T-vertex because "
                           << (n == 3 ? "it is a Y-vertex" : "of an unexpected reason") << std::endl;
    return false;
  }
  return true;
}
bool is_an_X_vertex(LineTes::Vertex_handle v) {  // :4058-4067
  if (v->degree() != 4) return false;
  std::vector<LineTes::Halfedge_handle> e = v->incident_halfedges();
  return e[0]->get_next_hf() == e[2]->twin() && e[1]->get_next_hf() == e[3]->twin();
}
bool has_an_I_vertex_neighbour(LineTes::Vertex_handle v) {  // :4073-4081
  std::vector<LineTes::Halfedge_handle> e = v->incident_halfedges();
  for (size_t i = 0; i < e.size(); i++) if (e[i]->source()->degree() == 1) return true;
  return false;
}
bool is_an_irreducible_X_vertex(LineTes::Vertex_handle v) { return is_an_X_vertex(v) && !has_an_I_vertex_neighbour(v); }
unsigned long int number_of_internal_vertices(TTessel &t) {  // :4105-4115: vertices minus the edges of the window sides
  unsigned long int compt = 0;
  LineTes::Seg_list_iterator s = t.segments_begin();
  for (Size i = 0; i < t.number_of_window_edges() && s != t.segments_end(); i++, ++s) compt += s->number_of_edges();
  return (unsigned long int)t.number_of_vertices() - compt;
}
static Polygon cycle_polygon(const HostTess &h, int h0, bool simplify) {
  Polygon p;
  int e = h0;
  do {
    if (!simplify || h.h_next_hf[e] < 0 || h.h_next_hf[e] != h.h_next[e]) p.push_back(pt(h, h.h_src[e ^ 1]));
    e = h.h_next[e];
  } while (e != h0);
  return p;
}
HPolygon face2poly(TTessel::Face_handle f, bool simplify) {  // :4314-4336
  const HostTess &h = *f.tess()->host();
  Polygon outer = cycle_polygon(h, h.f_he[f.id()], simplify);
  Polygons holes;
  std::vector<LineTes::Halfedge_handle> hs = f.holes();
  for (size_t i = 0; i < hs.size(); i++) holes.push_back(cycle_polygon(h, hs[i].id(), simplify));
  return HPolygon(outer, holes.begin(), holes.end());
}
static double sum_over_faces(TTessel *t, double (*feature)(HPolygon, TTessel *)) {
  double s = 0;
  std::vector<LineTes::Face_handle> fs = t->faces();
  for (size_t i = 0; i < fs.size(); i++) if (fs[i].data()) s += feature(face2poly(fs[i]), t);
  return s;
}
double sum_of_faces_squared_areas(TTessel *t) { return sum_over_faces(t, face_area_2); }
double sum_of_min_angles(TTessel *t) { return sum_over_faces(t, min_angle); }
double sum_of_angles_obt(TTessel *t) { return sum_over_faces(t, face_sum_of_angles); }
double sum_of_segment_squared_sizes(TTessel *t) {  // :4440-4445
  double s4 = 0;
  for (LineTes::Seg_list_iterator si = t->segments_begin(); si != t->segments_end(); ++si) s4 += segment_size_2(si->list_of_points(), t);
  return s4;
}
Polygons boundaries(HPolygon hp) {  // :4453-4469
  Polygons b;
  Polygon ob(hp.outer_boundary());
  if (ob.area() < 0) ob.reverse_orientation();
  b.push_back(ob);
  for (HPolygon::Hole_const_iterator h = hp.holes_begin(); h != hp.holes_end(); ++h) {
    Polygon q(*h);
    if (q.area() > 0) q.reverse_orientation();
    b.push_back(q);
  }
  return b;
}
Polygon simplify(Polygon p) {  // :4474-4492
  if (p.is_empty()) throw std::domain_error("simplify(Polygon), input polygon is empty");
  Polygon sp;
  const Size n = p.size();
  for (Size i = 0; i < n; i++) {
    const Point2 &a = p[(i + n - 1) % n], &b = p[i], &c = p[(i + 1) % n];
    if ((b.x() - a.x()) * (c.y() - a.y()) - (b.y() - a.y()) * (c.x() - a.x()) != 0.0) sp.push_back(b);
  }
  return sp;
}
HPolygon simplify(HPolygon p) {  // :4495-4506
  Polygons b = boundaries(p);
  for (size_t i = 0; i < b.size(); i++) b[i] = simplify(b[i]);
  return HPolygon(b[0], b.begin() + 1, b.end());
}
bool has_holes(LineTes::Face_handle &f) { return !f.holes().empty(); }
bool has_holes(HPolygon &p) { return p.number_of_holes() > 0; }

double is_point_inside_window(Point2 p, LineTes *t) {  // :4131-4135
  HPolygons ws = t->get_window();
  for (Size k = 0; k < ws.size(); k++) {
    const Polygon &w = ws[k].outer_boundary();
    double tol = window_tol(w);
    Size n = w.size();
    bool on = false;
    for (Size i = 0; i < n && !on; i++) on = on_seg(p, w[i], w[(i + 1) % n], tol);
    if (on) continue;
    bool in = false;
    for (Size i = 0, j = n - 1; i < n; j = i++)
      if ((w[i].y() > p.y()) != (w[j].y() > p.y()) &&
          p.x() < (w[j].x() - w[i].x()) * (p.y() - w[i].y()) / (w[j].y() - w[i].y()) + w[i].x())
        in = !in;
    if (in) return 1.0;
  }
  return 0.0;
}
double is_point_inside_window(Point2 p, TTessel *t) { return is_point_inside_window(p, static_cast<LineTes *>(t)); }
double is_segment_internal(std::vector<Point2> s, TTessel *t) {  // :4172-4196
  Point2 p1 = s.front(), p2 = s.back();
  if (is_point_inside_window(p1, t) > .5 || is_point_inside_window(p2, t) > .5) return 1.0;
  HPolygons ws = t->get_window();
  for (Size k = 0; k < ws.size(); k++) {
    const Polygon &w = ws[k].outer_boundary();
    double tol = window_tol(w);
    for (Size i = 0; i < w.size(); i++)
      if (on_seg(p1, w[i], w[(i + 1) % w.size()], tol) && on_seg(p2, w[i], w[(i + 1) % w.size()], tol)) return 0.0;
  }
  return 1.0;
}
double minus_is_segment_internal(std::vector<Point2> s, TTessel *t) { return -is_segment_internal(s, t); }
double edge_length(Segment e, TTessel *t) {  // :4141-4149: boundary edges only, as written
  std::vector<Point2> s;
  s.push_back(e.source()); s.push_back(e.target());
  if (is_segment_internal(s, t) > 0) return 0.0;
  return sqrt(e.squared_length());
}
double edge_length_internal(Segment e, TTessel *t) {
  std::vector<Point2> s;
  s.push_back(e.source()); s.push_back(e.target());
  if (is_segment_internal(s, t) > 0) return sqrt(e.squared_length());
  return 0.0;
}
double seg_number(std::vector<Point2>, TTessel *) { return 1.0; }
double face_number(HPolygon, TTessel *) { return 1.0; }
double face_area_2(HPolygon f, TTessel *) {  // :4215-4223
  double a = f.outer_boundary().area(), r = a * a;
  for (HPolygon::Hole_const_iterator h = f.holes_begin(); h != f.holes_end(); h++) r += h->area() * h->area();
  return r;
}
static double perimeter_of(const Polygon &p) {
  double l = 0;
  for (Size j = 0; j < p.size(); j++) l += sqrt(Segment(p[j], p[(j + 1) % p.size()]).squared_length());
  return l;
}
double face_perimeter(HPolygon f, TTessel *) {  // :4229-4239
  double l = perimeter_of(f.outer_boundary());
  for (HPolygon::Hole_const_iterator h = f.holes_begin(); h != f.holes_end(); h++) l += perimeter_of(*h);
  return l;
}
double face_shape(HPolygon f, TTessel *t) { return face_perimeter(f, t) / (4 * pow(face_area_2(f, t), 0.25)); }
double angle_between_vectors(double ax, double ay, double bx, double by) {  // :4255-4268
  double n1 = ax * ax + ay * ay, n2 = bx * bx + by * by, p = ax * bx + ay * by;
  double c = p / sqrt(n1 * n2);
  if (c >= 1.0) return 0.0;
  if (c <= -1.0) return CGAL_PI;
  return acos(c);
}
double face_sum_of_angles(HPolygon f, TTessel *) {  // :4285-4305
  const Polygon &ob = f.outer_boundary();
  Size n = ob.size();
  double tot = 0;
  for (Size i = 0; i < n; i++) {
    Size j = (i + n - 1) % n, k = (i + 1) % n;
    double a = CGAL_PI / 2 - angle_between_vectors(ob[j].x() - ob[i].x(), ob[j].y() - ob[i].y(),
                                                   ob[k].x() - ob[i].x(), ob[k].y() - ob[i].y());
    if (a >= 0) tot += a;
  }
  return tot;
}
double min_angle(HPolygon f, TTessel *) {  // :4348-4367
  const Polygon &ob = f.outer_boundary();
  Size n = ob.size();
  double m = angle_between_vectors(ob[0].x() - ob[n - 1].x(), ob[0].y() - ob[n - 1].y(), ob[1].x() - ob[0].x(),
                                   ob[1].y() - ob[0].y());
  for (Size i = 0; i < n; i++) {
    Size j = (i + n - 1) % n, k = (i + 1) % n;
    double a = angle_between_vectors(ob[j].x() - ob[i].x(), ob[j].y() - ob[i].y(), ob[k].x() - ob[i].x(),
                                     ob[k].y() - ob[i].y());
    if (a < m) m = a;
  }
  return m;
}
double segment_size_2(std::vector<Point2> s, TTessel *t) {  // :4424-4434
  std::vector<Point2> ends;
  ends.push_back(s.front()); ends.push_back(s.back());
  if (is_segment_internal(ends, t) > 0) return 0.0;
  double k = (double)s.size() - 1;
  return k * k;
}

// ---------------------------------------------------------------------------
// Energy (lib/ttessel.cpp:2856-3104)
// ---------------------------------------------------------------------------
namespace lite {
// GPU copy of (tessellation, feature program) used by Energy::statistic_variation
struct DeviceMirror {
  lite_ctx *ctx;
  const TTessel *tes;
  unsigned long version;
  std::vector<int32_t> prog;
  std::vector<int> he_export;            // internal halfedge id -> id in the uploaded SoA
  std::map<int, int> merge_row, flip_row;  // internal halfedge id -> row of the stat blocks
  std::vector<double> merge_stat, flip_stat;
  DeviceMirror() : ctx(0), tes(0), version(~0ul) {}
  ~DeviceMirror() { if (ctx) lite_ctx_destroy(ctx); }
};
}  // namespace lite

static std::vector<int> export_map(const HostTess &h) {
  std::vector<int> m(h.h_alive.size(), -1);
  int k = 0;
  for (size_t i = 0; i < h.h_alive.size(); i++)
    if (h.h_alive[i]) m[i] = k++;
  return m;
}
static std::vector<int> inverse_map(const std::vector<int> &m) {
  std::vector<int> inv;
  for (size_t i = 0; i < m.size(); i++)
    if (m[i] >= 0) { if ((size_t)m[i] >= inv.size()) inv.resize(m[i] + 1); inv[m[i]] = (int)i; }
  return inv;
}

Energy::Energy() : ttes(0), value(0), mirror(0) {}
Energy::~Energy() { delete mirror; }
Energy::Energy(const Energy &o) : ttes(o.ttes), theta(o.theta), features(o.features), value(o.value), mirror(0) {}
Energy &Energy::operator=(const Energy &o) {
  if (this != &o) {
    ttes = o.ttes; theta = o.theta; features = o.features; value = o.value;
    delete mirror; mirror = 0;
  }
  return *this;
}
void Energy::add_features_vertices(double (*f)(Point2, TTessel *)) { features.vertices.push_back(f); }
void Energy::add_features_edges(double (*f)(Segment, TTessel *)) { features.edges.push_back(f); }
void Energy::add_features_faces(double (*f)(HPolygon, TTessel *)) { features.faces.push_back(f); }
void Energy::add_features_segs(double (*f)(std::vector<Point2>, TTessel *)) { features.segs.push_back(f); }
void Energy::add_theta_vertices(double v) { theta.vertices.push_back(v); }
void Energy::add_theta_edges(double v) { theta.edges.push_back(v); }
void Energy::add_theta_faces(double v) { theta.faces.push_back(v); }
void Energy::add_theta_segs(double v) { theta.segs.push_back(v); }
void Energy::del_theta_vertices() { theta.vertices.clear(); }
void Energy::del_theta_edges() { theta.edges.clear(); }
void Energy::del_theta_faces() { theta.faces.clear(); }
void Energy::del_theta_segs() { theta.segs.clear(); }
void Energy::del_features_vertices() { features.vertices.clear(); }
void Energy::del_features_edges() { features.edges.clear(); }
void Energy::del_features_faces() { features.faces.clear(); }
void Energy::del_features_segs() { features.segs.clear(); }

std::vector<int32_t> Energy::device_program() const {
  if (theta.vertices.size() != features.vertices.size() || theta.edges.size() != features.edges.size() ||
      theta.faces.size() != features.faces.size() || theta.segs.size() != features.segs.size())
    throw std::logic_error("Energy: each category needs as many theta components as feature functions");
  std::vector<int32_t> p;
  double (*ipw)(Point2, TTessel *) = is_point_inside_window;
  for (Size i = 0; i < features.vertices.size(); i++) {
    if (features.vertices[i] == ipw) p.push_back(LITE_FEAT_IS_POINT_INSIDE_WINDOW);
    else throw std::invalid_argument("Energy: vertex feature is not a built-in function (no CPU fallback)");
  }
  for (Size i = 0; i < features.edges.size(); i++) {
    if (features.edges[i] == edge_length) p.push_back(LITE_FEAT_EDGE_LENGTH);
    else if (features.edges[i] == edge_length_internal) p.push_back(LITE_FEAT_EDGE_LENGTH_INTERNAL);
    else throw std::invalid_argument("Energy: edge feature is not a built-in function (no CPU fallback)");
  }
  for (Size i = 0; i < features.faces.size(); i++) {
    double (*f)(HPolygon, TTessel *) = features.faces[i];
    if (f == face_number) p.push_back(LITE_FEAT_FACE_NUMBER);
    else if (f == face_area_2) p.push_back(LITE_FEAT_FACE_AREA_2);
    else if (f == face_perimeter) p.push_back(LITE_FEAT_FACE_PERIMETER);
    else if (f == face_shape) p.push_back(LITE_FEAT_FACE_SHAPE);
    else if (f == face_sum_of_angles) p.push_back(LITE_FEAT_FACE_SUM_OF_ANGLES);
    else if (f == min_angle) p.push_back(LITE_FEAT_MIN_ANGLE);
    else throw std::invalid_argument("Energy: face feature is not a built-in function (no CPU fallback)");
  }
  for (Size i = 0; i < features.segs.size(); i++) {
    double (*f)(std::vector<Point2>, TTessel *) = features.segs[i];
    if (f == seg_number) p.push_back(LITE_FEAT_SEG_NUMBER);
    else if (f == is_segment_internal) p.push_back(LITE_FEAT_IS_SEGMENT_INTERNAL);
    else if (f == minus_is_segment_internal) p.push_back(LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL);
    else if (f == segment_size_2) p.push_back(LITE_FEAT_SEGMENT_SIZE_2);
    else throw std::invalid_argument("Energy: segment feature is not a built-in function (no CPU fallback)");
  }
  if (p.empty()) throw std::logic_error("Energy: no feature registered");
  return p;
}

void Energy::set_ttessel(TTessel *t) {  // :2862-2912: the current energy is computed from scratch
  ttes = t;
  value = 0;
  const HostTess &h = *t->host();
  for (Size i = 0; i != theta.vertices.size(); i++)
    for (size_t v = 0; v < h.v_alive.size(); v++)
      if (h.v_alive[v]) value += theta.vertices[i] * features.vertices[i](Point2(h.vx[v], h.vy[v]), t);
  for (Size i = 0; i != theta.edges.size(); i++)
    for (size_t e = 0; e < h.h_alive.size(); e += 2)
      if (h.h_alive[e])
        value += theta.edges[i] * features.edges[i](Segment(pt(h, h.h_src[e]), pt(h, h.h_src[e ^ 1])), t);
  for (Size i = 0; i != theta.faces.size(); i++)
    for (size_t f = 0; f < h.f_alive.size(); f++)
      if (h.f_alive[f] && h.f_inside[f])
        value += theta.faces[i] * features.faces[i](HPolygon(face_polygon(h, (int)f, true)), t);
  for (Size i = 0; i != theta.segs.size(); i++)
    for (size_t s = 0; s < h.s_alive.size(); s++)
      if (h.s_alive[s] && !h.s_bnd[s])
        value += theta.segs[i] * features.segs[i](LineTes::Seg(t, (int)s).list_of_points(), t);
}

static void sync_mirror(lite::DeviceMirror *&M, Energy &E, TTessel *t) {
  if (!t) throw std::logic_error("Energy: no T-tessellation set (set_ttessel)");
  std::vector<int32_t> prog = E.device_program();
  if (!M) M = new lite::DeviceMirror();
  if (!M->ctx) ck(lite_ctx_create(&M->ctx, 0, 0, 1, 0), "lite_ctx_create");
  if (M->tes == t && M->version == t->version() && M->prog == prog) return;
  HostTess &h = *t->host();
  require_t(h, "Energy", false);
  const lite_tess_soa *soa = h.export_soa();
  ck(lite_candidates_clear_splits(M->ctx), "lite_candidates_clear_splits");
  ck(lite_tess_upload(M->ctx, soa), "lite_tess_upload");
  ck(lite_energy_set(M->ctx, (int)prog.size(), prog.data()), "lite_energy_set");
  int32_t nm = 0, nf = 0;
  ck(lite_candidates_merges_flips(M->ctx, &nm, &nf), "lite_candidates_merges_flips");
  const int d = (int)prog.size();
  M->he_export = export_map(h);
  std::vector<int> inv = inverse_map(M->he_export);
  std::vector<int32_t> mh(nm ? nm : 1), fh(nf ? nf : 1);
  M->merge_stat.assign((size_t)(nm ? nm : 1) * d, 0.0);
  M->flip_stat.assign((size_t)(nf ? nf : 1) * d, 0.0);
  M->merge_row.clear(); M->flip_row.clear();
  if (nm) {
    ck(lite_debug_fetch(M->ctx, LITE_FETCH_MERGE_HE, 0, nm, mh.data()), "fetch merges");
    ck(lite_debug_fetch(M->ctx, LITE_FETCH_MERGE_STAT, 0, nm, M->merge_stat.data()), "fetch merges");
    for (int i = 0; i < nm; i++) M->merge_row[inv[mh[i]]] = i;
  }
  if (nf) {
    ck(lite_debug_fetch(M->ctx, LITE_FETCH_FLIP_E1, 0, nf, fh.data()), "fetch flips");
    ck(lite_debug_fetch(M->ctx, LITE_FETCH_FLIP_STAT, 0, nf, M->flip_stat.data()), "fetch flips");
    for (int i = 0; i < nf; i++) M->flip_row[inv[fh[i]]] = i;
  }
  M->tes = t; M->version = t->version(); M->prog = prog;
}

CatVector Energy::statistic_variation(TTessel::Modification &modif) {  // :2914-2980, on the device
  sync_mirror(mirror, *this, ttes);
  const int d = (int)mirror->prog.size();
  std::vector<double> st(d, 0.0);
  if (TTessel::Split *s = dynamic_cast<TTessel::Split *>(&modif)) {
    if (!s->is_valid()) throw std::domain_error("statistic_variation: invalid split");
    int32_t he = mirror->he_export[s->get_e1().id()];
    double x = s->get_x(), ang = s->get_angle();
    ck(lite_candidates_clear_splits(mirror->ctx), "clear");
    ck(lite_candidates_add_splits_given(mirror->ctx, 1, &he, &x, &ang, 1), "lite_candidates_add_splits_given");
    ck(lite_debug_fetch(mirror->ctx, LITE_FETCH_SPLIT_STAT, 0, 1, st.data()), "fetch split");
  } else if (TTessel::Merge *m = dynamic_cast<TTessel::Merge *>(&modif)) {
    std::map<int, int>::iterator it = mirror->merge_row.find(m->get_e().id());
    if (it == mirror->merge_row.end()) throw std::domain_error("statistic_variation: not a merge of this tessellation");
    for (int k = 0; k < d; k++) st[k] = mirror->merge_stat[(size_t)it->second * d + k];
  } else if (TTessel::Flip *f = dynamic_cast<TTessel::Flip *>(&modif)) {
    std::map<int, int>::iterator it = mirror->flip_row.find(f->get_e1().id());
    if (it == mirror->flip_row.end()) throw std::domain_error("statistic_variation: not a flip of this tessellation");
    for (int k = 0; k < d; k++) st[k] = mirror->flip_stat[(size_t)it->second * d + k];
  } else {
    throw std::invalid_argument("statistic_variation: unknown modification type");
  }
  for (int k = 0; k < d; k++) st[k] = -st[k];  // the device keeps MINUS the variation, as PseudoLikDiscrete does
  CatVector out;
  shape_like(theta, st.data(), &out);
  return out;
}
double Energy::variation(TTessel::Modification &modif) {  // :2994-3012
  return sum(theta * statistic_variation(modif));
}

// ---------------------------------------------------------------------------
// SMFChain (lib/ttessel.cpp:3121-3311)
// ---------------------------------------------------------------------------
SMFChain::SMFChain(Energy *e, double ps, double pm) : engy(e) { set_smf_prob(ps, pm); }
void SMFChain::set_smf_prob(double ps, double pm) {
  p_split = ps; p_merge = pm; p_flip = 1 - ps - pm; p_split_merge = ps + pm;
}
SMFChain::ModType SMFChain::propose_modif_type() {  // :3143-3151
  double r = the_rnd().get_double(0, 1);
  if (r <= p_split) return SPLIT;
  if (r <= p_split_merge) return MERGE;
  return FLIP;
}
double SMFChain::Hasting_ratio(TTessel::Split &s, double *ev) {  // :3158-3176
  TTessel *t = engy->get_ttessel();
  const HostTess &h = *t->host();
  int delta_nb = 1;
  int sg1 = h.h_seg[s.get_e1().id()], sg2 = h.h_seg[s.get_e2().id()];
  if (h.s_nedges[sg1] == 1 && !h.s_bnd[sg1]) delta_nb--;
  if (h.s_nedges[sg2] == 1 && !h.s_bnd[sg2]) delta_nb--;
  double r = p_merge / p_split;
  r *= (1 / CGAL_PI) * t->get_total_internal_length() / ((double)t->number_of_non_blocking_segments() + delta_nb);
  *ev = engy->variation(s);
  return r * exp(-*ev);
}
double SMFChain::Hasting_ratio(TTessel::Merge &m, double *ev) {  // :3181-3192
  TTessel *t = engy->get_ttessel();
  double r = p_split / p_merge;
  r *= CGAL_PI * (double)t->number_of_non_blocking_segments() /
       (t->get_total_internal_length() - 2 * m.get_e().get_length());
  *ev = engy->variation(m);
  return r * exp(-*ev);
}
double SMFChain::Hasting_ratio(TTessel::Flip &f, double *ev) {  // :3197-3252
  TTessel *t = engy->get_ttessel();
  const HostTess &h = *t->host();
  lite::FlipMove fm = h.make_flip(f.get_e1().id());
  int db = 0;
  int s_short = h.h_seg[fm.e1];
  if (h.s_nedges[s_short] == 2) db--;
  if (h.s_nedges[h.h_seg[fm.e3]] == 1) db++;
  int i_short = h.h_seg[h.h_next[fm.e1]], i_ext = h.h_seg[fm.e2];
  if (i_short != i_ext) {
    if (h.s_nedges[i_short] == 2 && !h.s_bnd[i_short]) db--;
    if (h.s_nedges[i_ext] == 1 && !h.s_bnd[i_ext]) db++;
  }
  double sb = (double)t->number_of_blocking_segments();
  double r = (sb == -db) ? sb : sb / (sb + db);
  *ev = engy->variation(f);
  return r * exp(-*ev);
}
ModifCounts SMFChain::step(unsigned long int n) {  // :3257-3311
  ModifCounts c = {0, 0, 0, 0, 0, 0};
  TTessel *t = engy->get_ttessel();
  for (unsigned long i = 0; i != n; i++) {
    ModType ty = propose_modif_type();
    double x = the_rnd().get_double(0, 1), ev = 0, r;
    if (ty == SPLIT) {
      c.proposed_S++;
      TTessel::Split s = t->propose_split();
      if (!s.is_valid()) continue;
      r = Hasting_ratio(s, &ev);
      if (r > 1 || x < r) { c.accepted_S++; t->update(s); engy->add_value(ev); }
    } else if (ty == MERGE) {
      c.proposed_M++;
      if (t->number_of_non_blocking_segments() == 0) continue;
      TTessel::Merge m = t->propose_merge();
      r = Hasting_ratio(m, &ev);
      if (r > 1 || x < r) { c.accepted_M++; t->update(m); engy->add_value(ev); }
    } else {
      c.proposed_F++;
      if (t->number_of_blocking_segments() == 0) continue;
      TTessel::Flip f = t->propose_flip();
      if (f.get_e2().is_null()) continue;
      r = Hasting_ratio(f, &ev);
      if (r > 1 || x < r) { c.accepted_F++; t->update(f); engy->add_value(ev); }
    }
  }
  return c;
}

// ---------------------------------------------------------------------------
// PseudoLikDiscrete / PLInferenceNOIS (lib/ttessel.cpp:3319-3540)
// ---------------------------------------------------------------------------
PseudoLikDiscrete::PseudoLikDiscrete()
    : engy(0), ctx_(0), device_(0), rank_(0), world_(1), seed_(0x5EED), offset_(0), n_splits_(0) {}
PseudoLikDiscrete::PseudoLikDiscrete(Energy *e)
    : engy(0), ctx_(0), device_(0), rank_(0), world_(1), seed_(0x5EED), offset_(0), n_splits_(0) {
  SetEnergy(e);
}
PseudoLikDiscrete::~PseudoLikDiscrete() { if (ctx_) lite_ctx_destroy(ctx_); }
void PseudoLikDiscrete::SetDevice(int device, int rank, int world_size, const void *id) {
  if (ctx_) { lite_ctx_destroy(ctx_); ctx_ = 0; }
  device_ = device; rank_ = rank; world_ = world_size;
  nccl_id_.clear();
  if (id) nccl_id_.assign((const char *)id, (const char *)id + 128);
  if (engy) SetEnergy(engy);
}
void PseudoLikDiscrete::ensure_ctx() {
  if (!ctx_) {
    ck(lite_ctx_create(&ctx_, device_, rank_, world_, nccl_id_.empty() ? 0 : nccl_id_.data()), "lite_ctx_create");
    ck(lite_ctx_set_option(ctx_, LITE_OPT_SAMPLER, iid_ ? LITE_SAMPLER_IID : LITE_SAMPLER_STRATIFIED), "lite_ctx_set_option");
  }
}
void PseudoLikDiscrete::SetIndependentSampling(bool on) {
  iid_ = on;
  if (ctx_) ck(lite_ctx_set_option(ctx_, LITE_OPT_SAMPLER, on ? LITE_SAMPLER_IID : LITE_SAMPLER_STRATIFIED), "lite_ctx_set_option");
}
void PseudoLikDiscrete::SetEnergy(Energy *e) {  // :3327-3348: all merges, all flips, their sums
  engy = e;
  TTessel *t = e->get_ttessel();
  if (!t) throw std::logic_error("PseudoLikDiscrete::SetEnergy: the energy has no T-tessellation");
  std::vector<int32_t> prog = e->device_program();
  ensure_ctx();
  HostTess &h = *t->host();
  require_t(h, "PseudoLikDiscrete::SetEnergy", false);
  ck(lite_candidates_clear_splits(ctx_), "lite_candidates_clear_splits");
  ck(lite_tess_upload(ctx_, h.export_soa()), "lite_tess_upload");
  ck(lite_energy_set(ctx_, (int)prog.size(), prog.data()), "lite_energy_set");
  int32_t nm = 0, nf = 0;
  ck(lite_candidates_merges_flips(ctx_, &nm, &nf), "lite_candidates_merges_flips");
  he_export_ = export_map(h);
  n_splits_ = 0;
}
void PseudoLikDiscrete::AddSplits(Size n) {  // :3356-3365; split_sample on the device (Philox)
  if (!ctx_) throw std::logic_error("PseudoLikDiscrete::AddSplits: SetEnergy first");
  ck(lite_candidates_add_splits_philox(ctx_, (int64_t)n, seed_, offset_), "lite_candidates_add_splits_philox");
  offset_ += n;
  n_splits_ += (std::int64_t)n;
}
void PseudoLikDiscrete::AddGivenSplit(TTessel::Split s) {  // :3371-3374
  if (!ctx_) throw std::logic_error("PseudoLikDiscrete::AddGivenSplit: SetEnergy first");
  if (!s.is_valid()) throw std::domain_error("AddGivenSplit: invalid split");
  int32_t he = he_export_.at(s.get_e1().id());
  double x = s.get_x(), ang = s.get_angle();
  // every rank holds the same given split: rank 0 contributes it, the count is global
  int64_t mine = (rank_ == 0) ? 1 : 0;
  ck(lite_candidates_add_splits_given(ctx_, mine, &he, &x, &ang, 1), "lite_candidates_add_splits_given");
  n_splits_ += 1;
}
void PseudoLikDiscrete::ClearSplits() {  // :3378-3380
  if (ctx_) ck(lite_candidates_clear_splits(ctx_), "lite_candidates_clear_splits");
  n_splits_ = 0;
}
void PseudoLikDiscrete::shape(const CatVector &like, const double *flat, CatVector *out) const { shape_like(like, flat, out); }
void PseudoLikDiscrete::GetValueGradientHessian(CatVector theta, double *v, CatVector *g, CatMatrix *H) {
  if (!ctx_) throw std::logic_error("PseudoLikDiscrete: SetEnergy first");
  std::vector<double> th = asVectorOfDoubles(theta);
  const int d = (int)th.size();
  std::vector<double> gg(d), hh((size_t)d * d);
  double lpl = 0;
  ck(lite_pl_eval(ctx_, th.data(), &lpl, g ? gg.data() : 0, H ? hh.data() : 0), "lite_pl_eval");
  {  // as the reference, which reports on std::cerr and goes on (lib/ttessel.cpp:4571-4574)
    int64_t fs = 0, ff = 0;
    if (lite_pl_failures(ctx_, &fs, &ff) == 0 && (fs || ff))
      std::cerr << "ray_exit_face: intersection of ray with face boundaries not found for " << fs << " split(s) and "
                << ff << " flip(s): their statistics count as zero" << std::endl;
  }
  if (v) *v = lpl;
  if (g) shape_like(theta, gg.data(), g);
  if (H) {
    CatMatrix m;
    std::vector<CatVector> rows(d);
    for (int i = 0; i < d; i++) shape_like(theta, hh.data() + (size_t)i * d, &rows[i]);
    Size j = 0;
    for (Size i = 0; i < theta.vertices.size(); i++) m.vertices.push_back(rows[j++]);
    for (Size i = 0; i < theta.edges.size(); i++) m.edges.push_back(rows[j++]);
    for (Size i = 0; i < theta.faces.size(); i++) m.faces.push_back(rows[j++]);
    for (Size i = 0; i < theta.segs.size(); i++) m.segs.push_back(rows[j++]);
    *H = m;
  }
}
double PseudoLikDiscrete::GetValue(CatVector theta) {  // :3387-3404
  double v = 0;
  GetValueGradientHessian(theta, &v, 0, 0);
  return v;
}
CatVector PseudoLikDiscrete::GetGradient(CatVector theta) {  // :3412-3429
  CatVector g;
  GetValueGradientHessian(theta, 0, &g, 0);
  return g;
}
CatMatrix PseudoLikDiscrete::GetHessian(CatVector theta) {  // :3433-3443
  CatMatrix H;
  GetValueGradientHessian(theta, 0, 0, &H);
  return H;
}

// d x d inverse by Gauss-Jordan with partial pivoting (stands for CGAL
// Linear_algebraCd<double>::inverse, lib/ttessel.cpp:3480)
static bool invert(FMatrix &a) {
  const Size n = a.size();
  FMatrix inv(n, std::vector<double>(n, 0.0));
  for (Size i = 0; i < n; i++) inv[i][i] = 1.0;
  for (Size c = 0; c < n; c++) {
    Size p = c;
    for (Size r = c + 1; r < n; r++)
      if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    if (a[p][c] == 0.0) return false;
    std::swap(a[p], a[c]); std::swap(inv[p], inv[c]);
    double piv = a[c][c];
    for (Size k = 0; k < n; k++) { a[c][k] /= piv; inv[c][k] /= piv; }
    for (Size r = 0; r < n; r++) {
      if (r == c) continue;
      double f = a[r][c];
      if (f == 0.0) continue;
      for (Size k = 0; k < n; k++) { a[r][k] -= f * a[c][k]; inv[r][k] -= f * inv[c][k]; }
    }
  }
  a = inv;
  return true;
}

PLInferenceNOIS::PLInferenceNOIS(Energy *eng, bool store, double lambda)
    : PseudoLikDiscrete(eng), stepsize(lambda), previous_lpl(0), store_path(store) {  // :3454-3463
  AddSplits(GetEnergy()->get_ttessel()->number_of_non_blocking_segments());
  if (store_path) {
    estimates.push_back(eng->get_theta());
    values.push_back(GetValue(eng->get_theta()));
  }
}
void PLInferenceNOIS::Step(unsigned int n) {  // :3470-3505
  Energy *e = GetEnergy();
  for (unsigned int i = 0; i != n; i++) {
    CatVector cur = e->get_theta(), grad;
    CatMatrix H;
    GetValueGradientHessian(cur, 0, &grad, &H);  // one fused device pass instead of two host loops
    FMatrix FH = asFMatrix(H);
    FVector FG = asFVector(grad);
    if (!invert(FH)) throw std::domain_error("PLInferenceNOIS::Step: singular Hessian");
    FVector eps(FG.size(), 0.0);
    for (Size r = 0; r < FG.size(); r++)
      for (Size c = 0; c < FG.size(); c++) eps[r] -= FH[r][c] * FG[c];
    CatVector shift(cur);
    fill(shift, eps);
    e->set_theta(cur + stepsize * shift);
    previous_theta = cur;
    if (store_path) estimates.push_back(e->get_theta());
    AddSplits(e->get_ttessel()->number_of_non_blocking_segments());
    if (store_path) values.push_back(GetValue(e->get_theta()));
  }
}
CatVector PLInferenceNOIS::GetEstimate() { return GetEnergy()->get_theta(); }
unsigned int PLInferenceNOIS::Run(double tol, unsigned int nmax) {  // :3526-3540
  if (nmax == 0) return 0;
  if (context())  // every iteration appends n_nb dummy splits (:3500): make room once
    lite_candidates_reserve(context(), (int64_t)nmax * (int64_t)GetEnergy()->get_ttessel()->number_of_non_blocking_segments());
  unsigned int i = 0;
  double current_lpl;
  do {
    previous_lpl = GetValue(GetEstimate());
    Step(1);
    current_lpl = GetValue(GetEstimate());
    i++;
  } while (fabs(current_lpl - previous_lpl) >= tol * (fabs(current_lpl) + tol) && i < nmax);
  return i;
}
