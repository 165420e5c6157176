// host_tess.cpp — see host_tess.h.
#include "host_tess.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <set>
#include <unordered_map>
#include <sstream>

namespace lite {

static const double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------
// Fenwick tree over halfedge ids (length-weighted sampling in O(log n))
// ---------------------------------------------------------------------------
void Fenwick::resize(size_t n) {
  if (n <= w.size()) return;
  size_t cap = w.size() ? w.size() : 64;
  while (cap < n) cap *= 2;
  std::vector<double> old = w;
  w.assign(cap, 0.0);
  t.assign(cap + 1, 0.0);
  for (size_t i = 0; i < old.size(); i++) w[i] = old[i];
  for (size_t i = 0; i < cap; i++) {
    t[i + 1] += w[i];
    size_t j = (i + 1) + ((i + 1) & (~(i + 1) + 1));
    if (j <= cap) t[j] += t[i + 1];
  }
}
void Fenwick::set(int i, double v) {
  double d = v - w[i];
  if (d == 0.0) return;
  w[i] = v;
  for (size_t j = (size_t)i + 1; j < t.size(); j += j & (~j + 1)) t[j] += d;
}
double Fenwick::total() const {
  double s = 0;
  for (size_t j = t.size() - 1; j > 0; j -= j & (~j + 1)) s += t[j];
  return s;
}
int Fenwick::find(double target) const {
  size_t pos = 0, n = t.size() - 1;
  size_t step = 1;
  while (step * 2 <= n) step *= 2;
  for (; step; step >>= 1) {
    size_t nx = pos + step;
    if (nx <= n && t[nx] <= target) {
      pos = nx;
      target -= t[nx];
    }
  }
  if (pos >= n) pos = n - 1;
  return (int)pos;
}

// ---------------------------------------------------------------------------
// element management
// ---------------------------------------------------------------------------
HostTess::HostTess() {}

int HostTess::new_vertex(double x, double y) {
  int v;
  if (!free_v.empty()) { v = free_v.back(); free_v.pop_back(); }
  else { v = (int)vx.size(); vx.push_back(0); vy.push_back(0); v_inc.push_back(-1); v_alive.push_back(0); }
  vx[v] = x; vy[v] = y; v_inc[v] = -1; v_alive[v] = 1;
  return v;
}
void HostTess::del_vertex(int v) { v_alive[v] = 0; free_v.push_back(v); }

int HostTess::new_edge_pair(int v1, int v2) {
  int h;
  if (!free_hp.empty()) { h = free_hp.back(); free_hp.pop_back(); }
  else {
    h = (int)h_src.size();
    for (int k = 0; k < 2; k++) {
      h_src.push_back(-1); h_next.push_back(-1); h_prev.push_back(-1); h_face.push_back(-1);
      h_seg.push_back(-1); h_next_hf.push_back(-1); h_prev_hf.push_back(-1);
      h_dir.push_back(0); h_alive.push_back(0); h_len.push_back(0);
    }
    fen.resize(h_src.size());
  }
  for (int k = 0; k < 2; k++) {
    int e = h + k;
    h_next[e] = h_prev[e] = h_face[e] = h_seg[e] = h_next_hf[e] = h_prev_hf[e] = -1;
    h_alive[e] = 1; h_len[e] = 0; h_dir[e] = (k == 0);
  }
  h_src[h] = v1; h_src[h + 1] = v2;
  return h;
}
void HostTess::del_edge_pair(int h) {
  h &= ~1;
  for (int k = 0; k < 2; k++) { h_alive[h + k] = 0; fen.set(h + k, 0.0); }
  free_hp.push_back(h);
}
int HostTess::new_face(bool inside) {
  int f;
  if (!free_f.empty()) { f = free_f.back(); free_f.pop_back(); }
  else { f = (int)f_he.size(); f_he.push_back(-1); f_inside.push_back(0); f_alive.push_back(0); }
  f_he[f] = -1; f_inside[f] = inside; f_alive[f] = 1;
  return f;
}
void HostTess::del_face(int f) { f_alive[f] = 0; free_f.push_back(f); }
int HostTess::new_seg(int he, bool bnd) {
  int s;
  if (!free_s.empty()) { s = free_s.back(); free_s.pop_back(); }
  else {
    s = (int)s_he.size();
    s_he.push_back(-1); s_nedges.push_back(0); s_pos.push_back(-1); s_bnd.push_back(0);
    s_alive.push_back(0); s_kind.push_back(0);
  }
  s_he[s] = he; s_nedges[s] = 1; s_pos[s] = -1; s_bnd[s] = bnd; s_alive[s] = 1; s_kind[s] = 0;
  return s;
}
void HostTess::del_seg(int s) { s_alive[s] = 0; free_s.push_back(s); }

void HostTess::set_len(int h) {
  double dx = vx[tgt(h)] - vx[h_src[h]], dy = vy[tgt(h)] - vy[h_src[h]];
  double l = sqrt(dx * dx + dy * dy);  // LineTes_halfedge::set_length, lib/ttessel.cpp:1459
  h_len[h] = l; h_len[h ^ 1] = l;
}
void HostTess::refresh_weight(int h) {
  for (int k = 0; k < 2; k++) {
    int e = (h & ~1) + k;
    bool in = h_alive[e] && h_face[e] >= 0 && f_inside[h_face[e]];
    fen.set(e, in ? h_len[e] : 0.0);
  }
}
void HostTess::set_junction(int e1, int e2) {  // lib/ttessel.cpp:3909-3930
  if (e1 >= 0) { h_next_hf[e1] = e2; h_prev_hf[e1 ^ 1] = (e2 >= 0) ? (e2 ^ 1) : -1; }
  if (e2 >= 0) { h_prev_hf[e2] = e1; h_next_hf[e2 ^ 1] = (e1 >= 0) ? (e1 ^ 1) : -1; }
}
void HostTess::seg_set_handle(int s, int h) {  // Seg::set_halfedge_handle, :841-851
  s_he[s] = h_dir[h] ? h : (h ^ 1);
}
int HostTess::seg_start(int s) const {
  int e = s_he[s];
  while (h_prev_hf[e] >= 0) e = h_prev_hf[e];
  return e;
}
int HostTess::seg_end(int s) const {
  int e = s_he[s];
  while (h_next_hf[e] >= 0) e = h_next_hf[e];
  return e;
}
int HostTess::flip_halfedge(int seg, int side) const {  // TTessel::propose_flip, :1766-1777
  return side == 0 ? (seg_start(seg) ^ 1) : seg_end(seg);
}
int HostTess::degree(int v) const {
  int h0 = v_inc[v], h = h0, n = 0;
  do { n++; h = h_next[h] ^ 1; } while (h != h0);
  return n;
}
void HostTess::list_add(int s, int kind) {
  std::vector<int> &L = (kind == 1) ? nb_list : b_list;
  s_kind[s] = kind; s_pos[s] = (int)L.size(); L.push_back(s);
}
void HostTess::list_remove(int s) {
  int kind = s_kind[s];
  if (!kind) return;
  std::vector<int> &L = (kind == 1) ? nb_list : b_list;
  int p = s_pos[s];
  if (ordered_lists) {
    L.erase(L.begin() + p);
    for (size_t i = p; i < L.size(); i++) s_pos[L[i]] = (int)i;
  } else {
    int last = L.back();
    L[p] = last; s_pos[last] = p; L.pop_back();
  }
  s_kind[s] = 0; s_pos[s] = -1;
}
int HostTess::n_cells() const {
  int n = 0;
  for (size_t f = 0; f < f_he.size(); f++) n += (f_alive[f] && f_inside[f]);
  return n;
}
int HostTess::n_alive_halfedges() const {
  int n = 0;
  for (size_t h = 0; h < h_src.size(); h++) n += h_alive[h];
  return n;
}

// ---------------------------------------------------------------------------
// window (LineTes::insert_window, lib/ttessel.cpp:60-120; hole-free polygon)
// ---------------------------------------------------------------------------
void HostTess::insert_window_polygon(const std::vector<double> &x, const std::vector<double> &y) {
  *this = HostTess();
  win_x = x; win_y = y;
  int n = (int)x.size();
  std::vector<int> vs(n), hs(n);
  for (int i = 0; i < n; i++) vs[i] = new_vertex(x[i], y[i]);
  for (int i = 0; i < n; i++) hs[i] = new_edge_pair(vs[i], vs[(i + 1) % n]);
  int fo = new_face(false), fi = new_face(true);
  for (int i = 0; i < n; i++) {
    int h = hs[i], hn = hs[(i + 1) % n], hp = hs[(i + n - 1) % n];
    h_next[h] = hn; h_prev[h] = hp;
    h_next[h ^ 1] = hp ^ 1; h_prev[h ^ 1] = hn ^ 1;
    h_face[h] = fi; h_face[h ^ 1] = fo;
    v_inc[vs[(i + 1) % n]] = h;
    h_dir[h] = 1; h_dir[h ^ 1] = 0;
    set_len(h);
    int s = new_seg(h, true);
    h_seg[h] = h_seg[h ^ 1] = s;
  }
  f_he[fi] = hs[0]; f_he[fo] = hs[0] ^ 1;
  int_length = 0;
  for (int i = 0; i < n; i++) { int_length += h_len[hs[i]]; refresh_weight(hs[i]); }
}
void HostTess::insert_window_rect(double x0, double y0, double x1, double y1) {
  insert_window_polygon({x0, x1, x1, x0}, {y0, y0, y1, y1});
}

// ---------------------------------------------------------------------------
// DCEL surgery
// ---------------------------------------------------------------------------
int HostTess::split_edge(int e, double px, double py) {  // lib/ttessel.cpp:606-655
  int e_next_hf = h_next_hf[e];
  int s = h_seg[e];
  bool e_dir = h_dir[e];
  int t = e ^ 1;
  int b = h_src[t];
  int v = new_vertex(px, py);
  int e2 = new_edge_pair(v, b), t2 = e2 ^ 1;
  h_next[e2] = h_next[e]; h_prev[h_next[e]] = e2; h_next[e] = e2; h_prev[e2] = e;
  h_prev[t2] = h_prev[t]; h_next[h_prev[t]] = t2; h_next[t2] = t; h_prev[t] = t2;
  h_src[t] = v;
  h_face[e2] = h_face[e]; h_face[t2] = h_face[t];
  v_inc[v] = e;
  if (v_inc[b] == e) v_inc[b] = e2;
  set_junction(e, e2);
  set_junction(e2, e_next_hf);
  h_dir[e] = e_dir; h_dir[t] = !e_dir; h_dir[e2] = e_dir; h_dir[t2] = !e_dir;
  set_len(e); set_len(e2);
  h_seg[e2] = h_seg[t2] = s;
  s_nedges[s]++;
  seg_set_handle(s, e);
  refresh_weight(e); refresh_weight(e2);
  return e;
}

int HostTess::merge_edge(int e1, int e2) {  // lib/ttessel.cpp:658-698
  int hf_avant = h_prev_hf[e1], hf_apres = h_next_hf[e2];
  double new_len = h_len[e1] + h_len[e2];
  int s = h_seg[e1];
  int v = tgt(e1), t1 = e1 ^ 1, t2 = e2 ^ 1, b = tgt(e2);
  h_next[e1] = h_next[e2]; h_prev[h_next[e2]] = e1;
  h_prev[t1] = h_prev[t2]; h_next[h_prev[t2]] = t1;
  h_src[t1] = b;
  if (v_inc[b] == e2) v_inc[b] = e1;
  if (f_he[h_face[e2]] == e2) f_he[h_face[e2]] = e1;
  if (f_he[h_face[t2]] == t2) f_he[h_face[t2]] = t1;
  del_edge_pair(e2);
  del_vertex(v);
  set_junction(hf_avant, e1);
  set_junction(e1, hf_apres);
  h_len[e1] = h_len[t1] = new_len;
  s_nedges[s]--;
  seg_set_handle(s, e1);
  refresh_weight(e1);
  return e1;
}

// new edge from tgt(prev1) to tgt(prev2) inside their common face; ext_prev is
// the halfedge whose straight continuation the new edge is (-1: new segment)
int HostTess::add_edge(int prev1, int prev2, int ext_prev) {  // lib/ttessel.cpp:700-740
  int v1 = tgt(prev1), v2 = tgt(prev2);
  int f = h_face[prev1];
  int a1 = h_next[prev1], a2 = h_next[prev2];
  int h = new_edge_pair(v1, v2), t = h ^ 1;
  h_next[prev1] = h; h_prev[h] = prev1; h_next[h] = a2; h_prev[a2] = h;
  h_next[prev2] = t; h_prev[t] = prev2; h_next[t] = a1; h_prev[a1] = t;
  int nf = new_face(true);
  f_he[f] = t; h_face[t] = f;
  f_he[nf] = h;
  int e = h;
  do { h_face[e] = nf; e = h_next[e]; } while (e != h);
  set_len(h);
  if (ext_prev >= 0) {
    set_junction(ext_prev, h);
    h_dir[h] = h_dir[ext_prev]; h_dir[t] = !h_dir[h];
    int s = h_seg[ext_prev];
    h_seg[h] = h_seg[t] = s;
    s_nedges[s]++;
  } else {
    h_dir[h] = 1; h_dir[t] = 0;
    int s = new_seg(h, false);
    h_seg[h] = h_seg[t] = s;
  }
  // refresh weights of the halfedges that changed face
  e = h;
  do { refresh_weight(e); e = h_next[e]; } while (e != h);
  refresh_weight(t);
  return h;
}

int HostTess::remove_edge(int e) {  // lib/ttessel.cpp:743-780
  int t = e ^ 1;
  int f1 = h_face[e], f2 = h_face[t];
  int p1 = h_prev[e], n1 = h_next[e], p2 = h_prev[t], n2 = h_next[t];
  int phf = h_prev_hf[e], nhf = h_next_hf[e];
  set_junction(phf, -1);
  set_junction(-1, nhf);
  int s = h_seg[e];
  if (phf < 0 && nhf < 0) {
    list_remove(s);
    del_seg(s);
  } else {
    s_nedges[s]--;
    seg_set_handle(s, phf >= 0 ? phf : nhf);
  }
  if (v_inc[tgt(e)] == e) v_inc[tgt(e)] = p2;
  if (v_inc[h_src[e]] == t) v_inc[h_src[e]] = p1;
  h_next[p1] = n2; h_prev[n2] = p1; h_next[p2] = n1; h_prev[n1] = p2;
  del_edge_pair(e);
  f_he[f1] = p1;
  int c = p1;
  do { h_face[c] = f1; c = h_next[c]; } while (c != p1);
  if (f2 != f1) del_face(f2);
  return f1;
}

void HostTess::suppress_edge(int e) {  // lib/ttessel.cpp:201-236
  int v1 = h_src[e], v2 = tgt(e);
  remove_edge(e);
  for (int v : {v1, v2}) {
    if (degree(v) == 2 && h_next_hf[v_inc[v]] >= 0) merge_edge(v_inc[v], h_next_hf[v_inc[v]]);
  }
}

// ---------------------------------------------------------------------------
// moves (double restatement of the reference constructors)
// ---------------------------------------------------------------------------
int HostTess::ray_exit(int start_he, double a, double b, double w, double px, double py,
                       double dx, double dy, int skip0, int skip1, double &hx,
                       double &hy) const {  // ray_exit_face, lib/ttessel.cpp:4515-4576
  double best = 1.7976931348623157e308;
  int res = -1;
  int h = start_he;
  do {
    if (h != skip0 && h != skip1) {
      int vs = h_src[h], vt = tgt(h);
      double sk = fma(a, vx[vs], fma(b, vy[vs], -w)), sn = fma(a, vx[vt], fma(b, vy[vt], -w));
      bool cr = (sk <= 0.0 && sn >= 0.0) || (sk >= 0.0 && sn <= 0.0);
      if (cr && sk != sn) {
        double sg = sk / (sk - sn);
        double X = fma(sg, vx[vt] - vx[vs], vx[vs]), Y = fma(sg, vy[vt] - vy[vs], vy[vs]);
        double rx = X - px, ry = Y - py;
        if (rx * dx + ry * dy >= 0.0) {
          double len = sqrt(rx * rx + ry * ry);
          if (len != 0.0 && len < best) { best = len; res = h; hx = X; hy = Y; }
        }
      }
    }
    h = h_next[h];
  } while (h != start_he);
  return res;
}

SplitMove HostTess::make_split(int e, double x, double angle) const {  // lib/ttessel.cpp:2090-2132
  SplitMove m;
  m.e1 = e;
  int vs = h_src[e], vt = tgt(e);
  double sx = vx[vs], sy = vy[vs], tx = vx[vt], ty = vy[vt];
  double e1_angle = atan2(ty - sy, tx - sx);
  double l = e1_angle + angle;
  double a = sin(l), b = -cos(l);
  double u = a * sx + b * sy, v = a * tx + b * ty;
  double w = u + x * (v - u);
  double sP = fma(a, sx, fma(b, sy, -w)), sQ = fma(a, tx, fma(b, ty, -w));
  if (!(sP != sQ)) return m;
  double t1 = sP / (sP - sQ);
  m.p1x = fma(t1, tx - sx, sx); m.p1y = fma(t1, ty - sy, sy);
  double dx, dy;
  if (sQ < 0.0) { dx = b; dy = -a; } else { dx = -b; dy = a; }
  int skip1 = (x == 0.0) ? h_prev[e] : -1;
  int e2 = ray_exit(f_he[h_face[e]], a, b, w, m.p1x, m.p1y, dx, dy, e, skip1, m.p2x, m.p2y);
  if (e2 < 0) return m;
  m.e2 = e2;
  m.valid = true;
  return m;
}

FlipMove HostTess::make_flip(int e1) const {  // lib/ttessel.cpp:2432-2454
  FlipMove m;
  m.e1 = e1;
  int hp = h_prev[e1];
  bool right = (hp != h_prev_hf[e1]);
  int e3 = right ? hp : (h_next[e1 ^ 1] ^ 1);
  m.e3 = e3; m.right = right;
  int a = h_src[e1];
  double ax = vx[a], ay = vy[a];
  double dx = ax - vx[h_src[e3]], dy = ay - vy[h_src[e3]];
  double nrm = sqrt(dx * dx + dy * dy);
  dx /= nrm; dy /= nrm;
  double la = dy, lb = -dx, lw = la * ax + lb * ay;
  int f = right ? h_face[e1 ^ 1] : h_face[e1];
  int skip0, skip1;
  if (right) { skip0 = e1 ^ 1; skip1 = h_next[e1 ^ 1]; }
  else { skip0 = e1; skip1 = h_prev[e1]; }
  int e2 = ray_exit(f_he[f], la, lb, lw, ax, ay, dx, dy, skip0, skip1, m.p2x, m.p2y);
  if (e2 < 0) return m;
  m.e2 = e2;
  m.valid = true;
  return m;
}

int HostTess::update_split(const SplitMove &sp) {  // TTessel::update(Split), :1580-1597
  int e1n = split_edge(sp.e1, sp.p1x, sp.p1y);
  int e2n = split_edge(sp.e2, sp.p2x, sp.p2y);
  int e = add_edge(e1n, e2n, -1);
  int_length += 2 * h_len[e];
  list_add(h_seg[e], 1);
  int ends[2] = {h_seg[h_next[e]], h_seg[h_next[e ^ 1]]};
  for (int s : ends) {
    if (!s_bnd[s] && s_nedges[s] <= 2 && s_kind[s] == 1) {
      list_remove(s);
      list_add(s, 2);
    }
  }
  return e;
}

void HostTess::update_merge(int e) {  // TTessel::update(Merge), :1604-1638
  int s = h_seg[e];
  list_remove(s);
  int ends[2] = {h_seg[h_next[e]], h_seg[h_next[e ^ 1]]};
  for (int sx : ends) {
    if (!s_bnd[sx] && s_nedges[sx] <= 2 && s_kind[sx] == 2) {
      list_remove(sx);
      list_add(sx, 1);
    }
  }
  double suppressed = h_len[e];
  suppress_edge(e);
  int_length -= 2 * suppressed;
}

int HostTess::update_flip(const FlipMove &fl) {  // TTessel::update(Flip), :1644-1701
  int e1 = fl.e1;
  int s1 = h_seg[e1];
  int s1_i = h_seg[h_next[e1]];
  int e3 = fl.e3;
  int s2 = h_seg[e3];
  // split_from_vertex (lib/ttessel.cpp:161-189): the halfedge arriving at a in the face of e2
  int prev_a = fl.right ? (e1 ^ 1) : h_prev[e1];
  int e2n = split_edge(fl.e2, fl.p2x, fl.p2y);
  int new_e = add_edge(prev_a, e2n, e3);
  int_length += 2 * h_len[new_e];
  int_length -= 2 * h_len[e1];
  int s2_i = h_seg[h_next[new_e]];
  suppress_edge(e1);
  if (s_nedges[s2] <= 2 && s_kind[s2] == 1) { list_remove(s2); list_add(s2, 2); }
  if (s_nedges[s1] <= 1 && s_kind[s1] == 2) { list_remove(s1); list_add(s1, 1); }
  if (s2_i != s1_i) {
    if (!s_bnd[s2_i] && s_nedges[s2_i] <= 2 && s_kind[s2_i] == 1) { list_remove(s2_i); list_add(s2_i, 2); }
    if (!s_bnd[s1_i] && s_nedges[s1_i] <= 1 && s_kind[s1_i] == 2) { list_remove(s1_i); list_add(s1_i, 1); }
  }
  return new_e;
}

void HostTess::finish_import() {
  free_v.clear(); free_hp.clear(); free_f.clear(); free_s.clear();
  for (int v = (int)vx.size() - 1; v >= 0; v--) if (!v_alive[v]) free_v.push_back(v);
  for (int h = (int)h_src.size() - 2; h >= 0; h -= 2) if (!h_alive[h]) free_hp.push_back(h);
  for (int f = (int)f_he.size() - 1; f >= 0; f--) if (!f_alive[f]) free_f.push_back(f);
  for (int s = (int)s_he.size() - 1; s >= 0; s--) if (!s_alive[s]) free_s.push_back(s);
  fen = Fenwick();
  fen.resize(h_src.size());
  for (size_t h = 0; h < h_src.size(); h += 2) if (h_alive[h]) refresh_weight((int)h);
}

int HostTess::sample_halfedge(double u01) const {
  double tot = fen.total();
  return fen.find(u01 * tot);
}

// ---------------------------------------------------------------------------
// export
// ---------------------------------------------------------------------------
const lite_tess_soa *HostTess::export_soa() {
  Export &E = ex;
  size_t nV0 = vx.size(), nH0 = h_src.size(), nF0 = f_he.size(), nS0 = s_he.size();
  std::vector<int> vmap(nV0, -1), hmap(nH0, -1), fmap(nF0, -1), smap(nS0, -1);
  int nV = 0, nH = 0, nF = 0, nS = 0;
  for (size_t i = 0; i < nV0; i++) if (v_alive[i]) vmap[i] = nV++;
  for (size_t i = 0; i < nH0; i++) if (h_alive[i]) hmap[i] = nH++;
  for (size_t i = 0; i < nF0; i++) if (f_alive[i]) fmap[i] = nF++;
  for (size_t i = 0; i < nS0; i++) if (s_alive[i]) smap[i] = nS++;
  E.vx.resize(nV); E.vy.resize(nV);
  for (size_t i = 0; i < nV0; i++) if (v_alive[i]) { E.vx[vmap[i]] = vx[i]; E.vy[vmap[i]] = vy[i]; }
  E.he_src.resize(nH); E.he_twin.resize(nH); E.he_next.resize(nH); E.he_face.resize(nH);
  E.he_seg.resize(nH); E.he_next_hf.resize(nH); E.he_prev_hf.resize(nH); E.he_dir.resize(nH);
  E.he_length.resize(nH);
  for (size_t i = 0; i < nH0; i++) {
    if (!h_alive[i]) continue;
    int k = hmap[i];
    E.he_src[k] = vmap[h_src[i]]; E.he_twin[k] = hmap[i ^ 1]; E.he_next[k] = hmap[h_next[i]];
    E.he_face[k] = fmap[h_face[i]]; E.he_seg[k] = smap[h_seg[i]];
    E.he_next_hf[k] = h_next_hf[i] >= 0 ? hmap[h_next_hf[i]] : -1;
    E.he_prev_hf[k] = h_prev_hf[i] >= 0 ? hmap[h_prev_hf[i]] : -1;
    E.he_dir[k] = h_dir[i]; E.he_length[k] = h_len[i];
  }
  E.face_inside.resize(nF); E.f_off.resize(nF); E.f_cnt.resize(nF); E.f_list.clear();
  for (size_t i = 0; i < nF0; i++) {
    if (!f_alive[i]) continue;
    int k = fmap[i];
    E.face_inside[k] = f_inside[i];
    E.f_off[k] = (int)E.f_list.size();
    int cnt = 0;
    if (f_inside[i]) {
      int h0 = f_he[i], h = h0;
      do { E.f_list.push_back(hmap[h]); cnt++; h = h_next[h]; } while (h != h0);
    }
    E.f_cnt[k] = cnt;
  }
  E.seg_start.resize(nS); E.seg_end.resize(nS); E.seg_n.resize(nS); E.seg_bnd.resize(nS);
  for (size_t i = 0; i < nS0; i++) {
    if (!s_alive[i]) continue;
    int k = smap[i];
    E.seg_start[k] = hmap[seg_start((int)i)]; E.seg_end[k] = hmap[seg_end((int)i)];
    E.seg_n[k] = s_nedges[i]; E.seg_bnd[k] = s_bnd[i];
  }
  E.nb.resize(nb_list.size()); E.b.resize(b_list.size());
  for (size_t i = 0; i < nb_list.size(); i++) E.nb[i] = smap[nb_list[i]];
  for (size_t i = 0; i < b_list.size(); i++) E.b[i] = smap[b_list[i]];
  lite_tess_soa &S = E.soa;
  S.n_vertices = nV; S.vx = E.vx.data(); S.vy = E.vy.data();
  S.n_halfedges = nH; S.he_src = E.he_src.data(); S.he_twin = E.he_twin.data();
  S.he_next = E.he_next.data(); S.he_face = E.he_face.data(); S.he_seg = E.he_seg.data();
  S.he_next_hf = E.he_next_hf.data(); S.he_prev_hf = E.he_prev_hf.data();
  S.he_dir = E.he_dir.data(); S.he_length = E.he_length.data();
  S.n_faces = nF; S.face_inside = E.face_inside.data(); S.face_he_offset = E.f_off.data();
  S.face_he_count = E.f_cnt.data(); S.n_face_he = (int)E.f_list.size(); S.face_he_list = E.f_list.data();
  S.n_segments = nS; S.seg_he_start = E.seg_start.data(); S.seg_he_end = E.seg_end.data();
  S.seg_n_edges = E.seg_n.data(); S.seg_on_boundary = E.seg_bnd.data();
  S.n_non_blocking = (int)E.nb.size(); S.n_blocking = (int)E.b.size();
  S.non_blocking = E.nb.data(); S.blocking = E.b.data();
  // int_length as the reference keeps it (incrementally); the sum of inside
  // halfedge lengths agrees to rounding
  S.int_length = int_length;
  return &S;
}

std::string HostTess::check() const {  // in the spirit of TTessel::is_valid, :1721-1753
  char buf[256];
  for (size_t h = 0; h < h_src.size(); h++) {
    if (!h_alive[h]) continue;
    if (h_prev[h_next[h]] != (int)h || h_src[h_next[h]] != tgt((int)h)) {
      snprintf(buf, sizeof buf, "halfedge %zu: broken next/prev", h);
      return buf;
    }
    if (h_face[h_next[h]] != h_face[h]) { snprintf(buf, sizeof buf, "halfedge %zu: face differs along ccb", h); return buf; }
    int nh = h_next_hf[h];
    if (nh >= 0 && (h_prev_hf[nh] != (int)h || h_src[nh] != tgt((int)h) || h_seg[nh] != h_seg[h])) {
      snprintf(buf, sizeof buf, "halfedge %zu: broken segment link", h);
      return buf;
    }
  }
  for (int s : nb_list) if (s_nedges[s] != 1 || s_bnd[s]) return "non-blocking segment with != 1 edge";
  for (int s : b_list) if (s_nedges[s] < 2 || s_bnd[s]) return "blocking segment with < 2 edges";
  for (size_t s = 0; s < s_he.size(); s++) {
    if (!s_alive[s]) continue;
    int n = 1, e = seg_start((int)s);
    while (h_next_hf[e] >= 0) { e = h_next_hf[e]; n++; }
    if (n != s_nedges[s]) { snprintf(buf, sizeof buf, "segment %zu: edge count %d != %d", s, s_nedges[s], n); return buf; }
    if (!s_bnd[s] && s_kind[s] != (n == 1 ? 1 : 2)) { snprintf(buf, sizeof buf, "segment %zu: wrong list", s); return buf; }
  }
  return "";
}

// ---------------------------------------------------------------------------
// CRTT chain (SMFChain::step, lib/ttessel.cpp:3257-3311; Hastings ratios :3158-3252)
// ---------------------------------------------------------------------------
ChainCounts crtt_chain(HostTess &t, double tau, uint64_t n_steps, uint64_t seed, double p_split,
                       double p_merge) {
  ChainCounts cc;
  std::mt19937_64 g(seed);
  auto U = [&]() { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); };
  const double p_sm = p_split + p_merge;
  bool save = t.ordered_lists;
  t.ordered_lists = false;  // the chain draws uniformly from the lists: order is irrelevant
  for (uint64_t it = 0; it < n_steps; it++) {
    double r0 = U(), x = U();
    if (r0 <= p_split) {
      cc.c[0]++;
      int e = t.sample_halfedge(U());
      double xx = U(), ang = acos(1 - 2 * U());
      SplitMove s = t.make_split(e, xx, ang);
      if (!s.valid) continue;
      int delta_nb = 1;
      if (t.s_nedges[t.h_seg[s.e1]] == 1 && !t.s_bnd[t.h_seg[s.e1]]) delta_nb--;
      if (t.s_nedges[t.h_seg[s.e2]] == 1 && !t.s_bnd[t.h_seg[s.e2]]) delta_nb--;
      double r = p_merge / p_split;
      r *= (1 / kPi) * t.int_length / ((double)t.nb_list.size() + delta_nb);
      r *= tau;  // exp(-dE), dE = log(tau)*(-1)
      if (r > 1 || x < r) { cc.c[1]++; t.update_split(s); }
    } else if (r0 <= p_sm) {
      cc.c[2]++;
      if (t.nb_list.empty()) continue;
      int s = t.nb_list[(size_t)(U() * t.nb_list.size())];
      int e = t.s_he[s];
      double r = p_split / p_merge;
      r *= kPi * (double)t.nb_list.size() / (t.int_length - 2 * t.h_len[e]);
      r /= tau;
      if (r > 1 || x < r) { cc.c[3]++; t.update_merge(e); }
    } else {
      cc.c[4]++;
      if (t.b_list.empty()) continue;
      int s = t.b_list[(size_t)(U() * t.b_list.size())];
      int side = U() < 0.5 ? 0 : 1;
      FlipMove f = t.make_flip(t.flip_halfedge(s, side));
      if (!f.valid) continue;
      int db = 0;
      int s_short = t.h_seg[f.e1];
      if (t.s_nedges[s_short] == 2) db--;
      if (t.s_nedges[t.h_seg[f.e3]] == 1) db++;
      int i_short = t.h_seg[t.h_next[f.e1]], i_ext = t.h_seg[f.e2];
      if (i_short != i_ext) {
        if (t.s_nedges[i_short] == 2 && !t.s_bnd[i_short]) db--;
        if (t.s_nedges[i_ext] == 1 && !t.s_bnd[i_ext]) db++;
      }
      double sb = (double)t.b_list.size();
      double r = (sb == -db) ? sb : sb / (sb + db);
      if (r > 1 || x < r) { cc.c[5]++; t.update_flip(f); }
    }
  }
  t.ordered_lists = save;
  return cc;
}

// ---------------------------------------------------------------------------
// text format (LineTes::write/read, lib/ttessel.cpp:1331-1441), doubles
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// simultaneous construction from maximal segments
// ---------------------------------------------------------------------------
bool build_from_segments(HostTess &t, const std::vector<double> &wx, const std::vector<double> &wy,
                         const std::vector<std::vector<double>> &segs, std::string &err) {
  return build_from_boundaries(t, {wx}, {wy}, segs, err);
}

bool build_from_boundaries(HostTess &t, const std::vector<std::vector<double>> &bx,
                           const std::vector<std::vector<double>> &by,
                           const std::vector<std::vector<double>> &segs, std::string &err) {
  return build_arrangement(t, bx, by, segs, false, err);
}

bool build_arrangement(HostTess &t, const std::vector<std::vector<double>> &bx,
                       const std::vector<std::vector<double>> &by,
                       const std::vector<std::vector<double>> &segs, bool lenient, std::string &err) {
  if (bx.empty() || bx.size() != by.size()) { err = "no window"; return false; }
  // all boundary vertices in one list; boundary side i runs wfrom[i] -> wto[i] (vertex ids)
  std::vector<double> wx, wy;
  std::vector<int> wto;
  for (size_t b = 0; b < bx.size(); b++) {
    const int n = (int)bx[b].size(), base = (int)wx.size();
    if (n < 3 || by[b].size() != bx[b].size()) { err = "a window boundary needs at least three vertices"; return false; }
    for (int i = 0; i < n; i++) { wx.push_back(bx[b][i]); wy.push_back(by[b][i]); wto.push_back(base + (i + 1) % n); }
  }
  const int nw = (int)wx.size();
  struct S { double x0, y0, x1, y1; std::vector<std::pair<double, int>> on; };  // (parameter, vertex)
  std::vector<S> L;
  for (int i = 0; i < nw; i++) L.push_back({wx[i], wy[i], wx[wto[i]], wy[wto[i]], {}});
  for (const auto &c : segs) {
    if (c.size() != 4) { err = "a segment needs four coordinates"; return false; }
    L.push_back({c[0], c[1], c[2], c[3], {}});
  }
  const int nL = (int)L.size();
  double xmin = wx[0], xmax = wx[0], ymin = wy[0], ymax = wy[0];
  for (int i = 0; i < nw; i++) { xmin = fmin(xmin, wx[i]); xmax = fmax(xmax, wx[i]); ymin = fmin(ymin, wy[i]); ymax = fmax(ymax, wy[i]); }
  const double scale = fmax(xmax - xmin, ymax - ymin);
  const double eps = 1e-9 * scale;
  // uniform grid over the window: every segment is listed in the cells its box meets
  const int G = std::max(1, std::min(2048, (int)sqrt((double)nL)));
  const double cw = (xmax - xmin) / G + 1e-300, ch = (ymax - ymin) / G + 1e-300;
  auto cx = [&](double x) { return std::max(0, std::min(G - 1, (int)((x - xmin) / cw))); };
  auto cy = [&](double y) { return std::max(0, std::min(G - 1, (int)((y - ymin) / ch))); };
  std::vector<std::vector<int>> grid((size_t)G * G);
  for (int i = 0; i < nL; i++) {
    // walk the cells along the segment (conservative: the cells of its bounding box row by row,
    // restricted to the columns the segment spans in that row)
    const S &s = L[i];
    const int r0 = cy(fmin(s.y0, s.y1) - eps), r1 = cy(fmax(s.y0, s.y1) + eps);
    for (int r = r0; r <= r1; r++) {
      double xa = fmin(s.x0, s.x1), xb = fmax(s.x0, s.x1);
      if (s.y1 != s.y0) {
        const double ylo = ymin + r * ch - eps, yhi = ymin + (r + 1) * ch + eps;
        const double ta = (ylo - s.y0) / (s.y1 - s.y0), tb = (yhi - s.y0) / (s.y1 - s.y0);
        const double t0 = fmax(0.0, fmin(ta, tb)), t1 = fmin(1.0, fmax(ta, tb));
        const double x0 = s.x0 + t0 * (s.x1 - s.x0), x1 = s.x0 + t1 * (s.x1 - s.x0);
        xa = fmin(x0, x1); xb = fmax(x0, x1);
      }
      for (int q = cx(xa - eps); q <= cx(xb + eps); q++) grid[(size_t)r * G + q].push_back(i);
    }
  }
  // no two segments may cross in their interiors (X-vertex): distances of the ends of one
  // to the line of the other have opposite signs, both ways (a general arrangement keeps the
  // pairs: each becomes a vertex on both segments)
  std::set<std::pair<int, int>> crossings;
  {
    auto side = [&](const S &a, double px, double py) {
      const double ex = a.x1 - a.x0, ey = a.y1 - a.y0;
      const double d = (ex * (py - a.y0) - ey * (px - a.x0)) / sqrt(ex * ex + ey * ey);
      return d > eps ? 1 : (d < -eps ? -1 : 0);
    };
    for (const auto &cell : grid)
      for (size_t p = 0; p < cell.size(); p++)
        for (size_t q = p + 1; q < cell.size(); q++) {
          const S &a = L[cell[p]], &b = L[cell[q]];
          if (side(a, b.x0, b.y0) * side(a, b.x1, b.y1) < 0 && side(b, a.x0, a.y0) * side(b, a.x1, a.y1) < 0) {
            if (!lenient) {
              err = "two segments cross (X-vertex): not a T-tessellation";
              return false;
            }
            crossings.insert({std::min(cell[p], cell[q]), std::max(cell[p], cell[q])});
          }
        }
  }
  // vertices: window corners first, then the two ends of every internal segment
  t = HostTess();
  t.win_x = bx[0]; t.win_y = by[0];
  t.win_bx = bx; t.win_by = by;
  t.has_holes = bx.size() > 1;
  std::vector<double> vx, vy;
  for (int i = 0; i < nw; i++) { vx.push_back(wx[i]); vy.push_back(wy[i]); }
  for (int i = 0; i < nw; i++) { L[i].on.push_back({0.0, i}); L[i].on.push_back({1.0, wto[i]}); }
  // general arrangement: ends that coincide (with a window corner or with each other: L-vertices)
  // share one vertex
  std::vector<int> end_vertex(2 * (size_t)nL, -1);
  // vertices by cell of a fine grid (cells of 4 eps: a point within eps of another lies in its
  // cell or in one of the eight around it)
  std::unordered_map<uint64_t, std::vector<int>> by_cell;
  auto cell_of = [&](double x, double y, int dx, int dy) {
    const uint64_t cxq = (uint64_t)(int64_t)floor((x - xmin) / (4 * eps)) + 8 + dx, cyq = (uint64_t)(int64_t)floor((y - ymin) / (4 * eps)) + 8 + dy;
    return (cxq << 32) ^ cyq;
  };
  auto shared_vertex = [&](double px, double py) {   // the vertex within eps of the point, made if there is none
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++) {
        auto it = by_cell.find(cell_of(px, py, dx, dy));
        if (it == by_cell.end()) continue;
        for (int k : it->second)
          if (fabs(vx[k] - px) <= eps && fabs(vy[k] - py) <= eps) return k;
      }
    const int v = (int)vx.size();
    vx.push_back(px); vy.push_back(py);
    by_cell[cell_of(px, py, 0, 0)].push_back(v);
    return v;
  };
  if (lenient) {
    for (int k = 0; k < nw; k++) by_cell[cell_of(vx[k], vy[k], 0, 0)].push_back(k);
    for (int i = nw; i < nL; i++)
      for (int e = 0; e < 2; e++)
        end_vertex[2 * (size_t)i + e] = shared_vertex(e ? L[i].x1 : L[i].x0, e ? L[i].y1 : L[i].y0);
  }
  for (int i = nw; i < nL; i++) {
    for (int e = 0; e < 2; e++) {
      const double px = e ? L[i].x1 : L[i].x0, py = e ? L[i].y1 : L[i].y0;
      const int v = lenient ? end_vertex[2 * (size_t)i + e] : (int)vx.size();
      if (!lenient) { vx.push_back(px); vy.push_back(py); }
      L[i].on.push_back({(double)e, v});
      // the segment this end point abuts on
      int host = -1; double best = eps, host_t = 0;
      for (int j : grid[(size_t)cy(py) * G + cx(px)]) {
        if (j == i) continue;
        const S &h = L[j];
        const double ex = h.x1 - h.x0, ey = h.y1 - h.y0, l2 = ex * ex + ey * ey;
        const double tt = ((px - h.x0) * ex + (py - h.y0) * ey) / l2;
        if (tt < 0 || tt > 1) continue;
        const double qx = h.x0 + tt * ex - px, qy = h.y0 + tt * ey - py;
        const double d = sqrt(qx * qx + qy * qy);
        if (lenient) {
          // every segment whose interior holds the point gets the vertex (once); a segment that
          // merely ends there shares it already
          const double hl = sqrt(l2);
          if (d > eps || tt * hl <= eps || (1 - tt) * hl <= eps) continue;
          bool there = false;
          for (const auto &o : L[j].on) there = there || o.second == v;
          if (!there) L[j].on.push_back({tt, v});
          continue;
        }
        if (d <= best) { best = d; host = j; host_t = tt; }
      }
      if (host < 0) {
        if (lenient) continue;  // a free end (I-vertex), or an end shared with other ends only
        err = "a segment end lies on no other segment and not on the window boundary"; return false;
      }
      const double hl = sqrt((L[host].x1 - L[host].x0) * (L[host].x1 - L[host].x0) + (L[host].y1 - L[host].y0) * (L[host].y1 - L[host].y0));
      if (host_t * hl <= eps || (1 - host_t) * hl <= eps) { err = "two segments meet at their end points (L- or X-vertex): not a T-tessellation"; return false; }
      L[host].on.push_back({host_t, v});
    }
  }
  for (const auto &c : crossings) {   // general arrangement only
    const S &a = L[c.first], &b = L[c.second];
    const double ax = a.x1 - a.x0, ay = a.y1 - a.y0, bxx = b.x1 - b.x0, byy = b.y1 - b.y0;
    const double den = ax * byy - ay * bxx;
    const double ta = ((b.x0 - a.x0) * byy - (b.y0 - a.y0) * bxx) / den;
    const double tb = ((b.x0 - a.x0) * ay - (b.y0 - a.y0) * ax) / den;
    // a segment end may sit on the crossing already: one vertex
    const double cxp = a.x0 + ta * ax, cyp = a.y0 + ta * ay;
    const int v = shared_vertex(cxp, cyp);
    for (int side = 0; side < 2; side++) {
      auto &on = L[side ? c.second : c.first].on;
      bool there = false;
      for (const auto &o : on) there = there || o.second == v;
      if (!there) on.push_back({side ? tb : ta, v});
    }
  }
  // edges along every segment
  struct E { int v0, v1, seg, prev_hf, next_hf; bool bnd; };
  std::vector<E> edges;
  for (int i = 0; i < nL; i++) {
    auto &on = L[i].on;
    std::sort(on.begin(), on.end());
    for (size_t k = 0; k + 1 < on.size(); k++) {
      if (on[k + 1].first - on[k].first <= 0) { err = "two vertices coincide on a segment"; return false; }
      E e; e.v0 = on[k].second; e.v1 = on[k + 1].second; e.seg = i; e.bnd = i < nw;
      e.prev_hf = k ? (int)edges.size() - 1 : -1; e.next_hf = (k + 2 < on.size()) ? (int)edges.size() + 1 : -1;
      edges.push_back(e);
    }
  }
  const int nE = (int)edges.size(), nV = (int)vx.size();
  // DCEL arrays: halfedge 2e runs v0 -> v1 (dir true), 2e+1 back
  t.vx = vx; t.vy = vy; t.v_inc.assign(nV, -1); t.v_alive.assign(nV, 1);
  t.h_src.assign(2 * nE, -1); t.h_next.assign(2 * nE, -1); t.h_prev.assign(2 * nE, -1); t.h_face.assign(2 * nE, -1);
  t.h_seg.assign(2 * nE, -1); t.h_next_hf.assign(2 * nE, -1); t.h_prev_hf.assign(2 * nE, -1);
  t.h_dir.assign(2 * nE, 0); t.h_alive.assign(2 * nE, 1); t.h_len.assign(2 * nE, 0.0);
  std::vector<std::vector<int>> out(nV);
  for (int e = 0; e < nE; e++) {
    const E &ed = edges[e];
    t.h_src[2 * e] = ed.v0; t.h_src[2 * e + 1] = ed.v1;
    t.h_seg[2 * e] = t.h_seg[2 * e + 1] = ed.seg;
    t.h_dir[2 * e] = 1; t.h_dir[2 * e + 1] = 0;
    t.h_next_hf[2 * e] = ed.next_hf >= 0 ? 2 * ed.next_hf : -1;
    t.h_prev_hf[2 * e] = ed.prev_hf >= 0 ? 2 * ed.prev_hf : -1;
    t.h_next_hf[2 * e + 1] = ed.prev_hf >= 0 ? 2 * ed.prev_hf + 1 : -1;
    t.h_prev_hf[2 * e + 1] = ed.next_hf >= 0 ? 2 * ed.next_hf + 1 : -1;
    const double dx = vx[ed.v1] - vx[ed.v0], dy = vy[ed.v1] - vy[ed.v0];
    t.h_len[2 * e] = t.h_len[2 * e + 1] = sqrt(dx * dx + dy * dy);
    out[ed.v0].push_back(2 * e); out[ed.v1].push_back(2 * e + 1);
    t.v_inc[ed.v1] = 2 * e; t.v_inc[ed.v0] = 2 * e + 1;
  }
  // next(): at the head of h, turn to the outgoing halfedge that follows twin(h) clockwise
  for (int v = 0; v < nV; v++) {
    auto &o = out[v];
    if (o.empty() || (o.size() < 2 && !lenient)) { err = "dangling vertex"; return false; }
    std::sort(o.begin(), o.end(), [&](int a, int b) {
      const int ta = t.h_src[a ^ 1], tb = t.h_src[b ^ 1];
      return atan2(vy[ta] - vy[v], vx[ta] - vx[v]) < atan2(vy[tb] - vy[v], vx[tb] - vx[v]);
    });
    const int k = (int)o.size();
    for (int i = 0; i < k; i++) {
      const int in = o[i] ^ 1;              // arrives at v
      const int nx = o[(i + k - 1) % k];    // the next outgoing halfedge clockwise
      t.h_next[in] = nx; t.h_prev[nx] = in;
    }
  }
  // Faces.  The window is on the LEFT of every boundary side as given, so halfedge 2e of a
  // boundary edge (it runs along the side) bounds an inside face and its twin an outside one
  // (the unbounded face or the inside of a hole).  A counter-clockwise cycle is the outer boundary
  // of a face: inside unless it holds the twin of a boundary halfedge.  A clockwise cycle is an
  // inner boundary: of the unbounded face if it holds such twins (one per window polygon),
  // otherwise of the inside face that surrounds it (a hole that no segment touches).
  std::vector<int> cw_cycles;   // one halfedge of each clockwise cycle of inside-facing halfedges
  int n_outer = 0;
  for (int h0 = 0; h0 < 2 * nE; h0++) {
    if (t.h_face[h0] >= 0 || t.h_face[h0] == -2) continue;
    double a2 = 0;
    bool outside_side = false;
    int h = h0, guard = 0;
    do {
      const int s = t.h_src[h], d = t.h_src[h ^ 1];
      a2 += vx[s] * vy[d] - vx[d] * vy[s];
      if (edges[h >> 1].bnd && (h & 1)) outside_side = true;
      h = t.h_next[h];
      if (h < 0 || ++guard > 2 * nE) { err = "broken boundary cycle"; return false; }
    } while (h != h0);
    // (a general arrangement also has the contours of free-standing groups of segments: trees have
    // no area, up to rounding)
    if ((a2 < 0 || (lenient && a2 <= 1e-12 * scale * scale)) && !outside_side) {   // inner boundary of an inside face: assigned below
      cw_cycles.push_back(h0);
      h = h0;
      do { t.h_face[h] = -2; h = t.h_next[h]; } while (h != h0);
      continue;
    }
    const bool inside = a2 > 0 && !outside_side;
    if (a2 < 0) n_outer++;
    const int f = (int)t.f_he.size();
    t.f_he.push_back(h0); t.f_inside.push_back(inside); t.f_alive.push_back(1);
    h = h0;
    do { t.h_face[h] = f; h = t.h_next[h]; } while (h != h0);
  }
  if (n_outer < 1 || n_outer > (int)bx.size()) { err = "the segments do not form a tessellation of the window"; return false; }
  for (int h0 : cw_cycles) {
    // the smallest inside face whose outer boundary surrounds a vertex of the cycle
    const double px = vx[t.h_src[h0]], py = vy[t.h_src[h0]];
    int best = -1; double best_a = 0;
    for (size_t f = 0; f < t.f_he.size(); f++) {
      if (!t.f_inside[f]) continue;
      bool in = false; double a2 = 0;
      int h = t.f_he[f];
      const int hs = h;
      do {
        const int s = t.h_src[h], d = t.h_src[h ^ 1];
        a2 += vx[s] * vy[d] - vx[d] * vy[s];
        if ((vy[s] > py) != (vy[d] > py)) {
          const double xc = vx[s] + (py - vy[s]) * (vx[d] - vx[s]) / (vy[d] - vy[s]);
          if (px < xc) in = !in;
        }
        h = t.h_next[h];
      } while (h != hs);
      if (in && (best < 0 || a2 < best_a)) { best = (int)f; best_a = a2; }
    }
    if (best < 0) { err = "a boundary cycle lies in no face of the window"; return false; }
    int h = h0;
    do { t.h_face[h] = best; h = t.h_next[h]; } while (h != h0);
    t.has_holes = true;
  }
  // segments and the blocking / non-blocking lists, in input order
  t.s_he.assign(nL, -1); t.s_nedges.assign(nL, 0); t.s_pos.assign(nL, -1); t.s_bnd.assign(nL, 0);
  t.s_alive.assign(nL, 1); t.s_kind.assign(nL, 0);
  t.nb_list.clear(); t.b_list.clear();
  for (int e = 0; e < nE; e++) { const int s = edges[e].seg; if (t.s_he[s] < 0) t.s_he[s] = 2 * e; t.s_nedges[s]++; }
  t.int_length = 0;
  for (int h = 0; h < 2 * nE; h++) if (t.f_inside[t.h_face[h]]) t.int_length += t.h_len[h];
  for (int s = 0; s < nL; s++) {
    t.s_bnd[s] = s < nw;
    if (s < nw) continue;
    if (t.s_nedges[s] == 1) { t.s_kind[s] = 1; t.s_pos[s] = (int)t.nb_list.size(); t.nb_list.push_back(s); }
    else { t.s_kind[s] = 2; t.s_pos[s] = (int)t.b_list.size(); t.b_list.push_back(s); }
  }
  // a window side must bound an outside face on its right and an inside one on its left
  for (int i = 0; i < nw; i++)
    if (t.f_inside[t.h_face[t.s_he[i] ^ 1]] || !t.f_inside[t.h_face[t.s_he[i]]]) { err = "window side with the window on the wrong side (outer boundaries counter-clockwise, holes clockwise)"; return false; }
  t.finish_import();
  if (lenient) {   // links only: the T-tessellation invariants of check() do not apply
    t.arrangement_only = true;
    for (size_t h = 0; h < t.h_src.size(); h++)
      if (t.h_face[h] < 0 || t.h_next[h] < 0 || t.h_prev[t.h_next[h]] != (int)h) { err = "inconsistent arrangement"; return false; }
    return true;
  }
  std::string bad = t.check();
  if (!bad.empty()) { err = "inconsistent tessellation: " + bad; return false; }
  return true;
}

// ---------------------------------------------------------------------------
// clip_segment_by_polygon (lib/ttessel.cpp:3696-3786): the longest piece of a segment inside
// the open window (its boundaries cut the segment and are not part of the result)
// ---------------------------------------------------------------------------
static bool point_in_open_window(const std::vector<std::vector<double>> &bx, const std::vector<std::vector<double>> &by,
                                 double px, double py, double eps) {
  bool in = false;
  for (size_t b = 0; b < bx.size(); b++) {
    const size_t n = bx[b].size();
    for (size_t i = 0; i < n; i++) {
      const size_t j = (i + 1) % n;
      const double x0 = bx[b][i], y0 = by[b][i], x1 = bx[b][j], y1 = by[b][j];
      const double ex = x1 - x0, ey = y1 - y0, l2 = ex * ex + ey * ey;
      double tt = ((px - x0) * ex + (py - y0) * ey) / l2;
      tt = fmax(0.0, fmin(1.0, tt));
      const double qx = x0 + tt * ex - px, qy = y0 + tt * ey - py;
      if (qx * qx + qy * qy <= eps * eps) return false;   // on a boundary: outside the open set
      if ((y0 > py) != (y1 > py)) {
        const double xc = x0 + (py - y0) * ex / ey;
        if (px < xc) in = !in;
      }
    }
  }
  return in;   // odd number of boundaries crossed: inside an outer boundary and in none of its holes
}

bool clip_segment(const std::vector<std::vector<double>> &bx, const std::vector<std::vector<double>> &by,
                  const double seg[4], double out[4], std::string &err) {
  if (bx.empty() || bx.size() != by.size()) { err = "no window"; return false; }
  const double sx = seg[2] - seg[0], sy = seg[3] - seg[1], sl = sqrt(sx * sx + sy * sy);
  if (!(sl > 0)) { err = "degenerate segment"; return false; }
  double scale = sl;
  for (size_t b = 0; b < bx.size(); b++)
    for (size_t i = 0; i < bx[b].size(); i++) scale = fmax(scale, fmax(fabs(bx[b][i]), fabs(by[b][i])));
  const double eps = 1e-9 * scale;
  std::vector<double> cuts = {0.0, 1.0};
  for (size_t b = 0; b < bx.size(); b++) {
    const size_t n = bx[b].size();
    for (size_t i = 0; i < n; i++) {
      const size_t j = (i + 1) % n;
      const double x0 = bx[b][i], y0 = by[b][i], ex = bx[b][j] - x0, ey = by[b][j] - y0;
      const double el = sqrt(ex * ex + ey * ey);
      const double den = sx * ey - sy * ex;
      if (fabs(den) > 1e-14 * sl * el) {
        const double ts = ((x0 - seg[0]) * ey - (y0 - seg[1]) * ex) / den;   // along the segment
        const double te = ((x0 - seg[0]) * sy - (y0 - seg[1]) * sx) / den;   // along the boundary edge
        if (ts * sl >= -eps && (1 - ts) * sl >= -eps && te * el >= -eps && (1 - te) * el >= -eps)
          cuts.push_back(fmax(0.0, fmin(1.0, ts)));
      } else {
        // parallel: a boundary edge lying on the segment's line cuts it at the edge's two ends
        const double d = ((x0 - seg[0]) * sy - (y0 - seg[1]) * sx) / sl;
        if (fabs(d) > eps) continue;
        const double ta = ((x0 - seg[0]) * sx + (y0 - seg[1]) * sy) / (sl * sl);
        const double tb = ((x0 + ex - seg[0]) * sx + (y0 + ey - seg[1]) * sy) / (sl * sl);
        if (ta > 0 && ta < 1) cuts.push_back(ta);
        if (tb > 0 && tb < 1) cuts.push_back(tb);
      }
    }
  }
  std::sort(cuts.begin(), cuts.end());
  double best = -1, b0 = 0, b1 = 0;
  for (size_t k = 0; k + 1 < cuts.size(); k++) {
    const double ta = cuts[k], tb = cuts[k + 1];
    if ((tb - ta) * sl <= eps) continue;
    const double tm = 0.5 * (ta + tb);
    if (!point_in_open_window(bx, by, seg[0] + tm * sx, seg[1] + tm * sy, eps)) continue;
    if (tb - ta > best) { best = tb - ta; b0 = ta; b1 = tb; }
  }
  if (best < 0) { err = "the segment does not hit the window"; return false; }
  out[0] = b0 == 0.0 ? seg[0] : seg[0] + b0 * sx; out[1] = b0 == 0.0 ? seg[1] : seg[1] + b0 * sy;
  out[2] = b1 == 1.0 ? seg[2] : seg[0] + b1 * sx; out[3] = b1 == 1.0 ? seg[3] : seg[1] + b1 * sy;
  return true;
}

// ---------------------------------------------------------------------------
// LineTes::insert_segment (lib/ttessel.cpp:904-980): the segment, clipped by the window, joins
// the maximal segments already there and the arrangement is built again from all of them --
// as a T-tessellation if they form one, else as a general arrangement (free ends, crossings,
// corners) that can be inspected, written and inserted into further, nothing else.
// ---------------------------------------------------------------------------
bool insert_segment(HostTess &t, const double seg[4], int *seg_id, std::string &err) {
  std::vector<std::vector<double>> bx = t.win_bx, by = t.win_by;
  if (bx.empty() && !t.win_x.empty()) { bx = {t.win_x}; by = {t.win_y}; }
  if (bx.empty()) { err = "no window"; return false; }
  double c[4];
  if (!clip_segment(bx, by, seg, c, err)) return false;
  std::vector<std::vector<double>> segs;
  for (size_t s = 0; s < t.s_he.size(); s++) {
    if (!t.s_alive[s] || t.s_bnd[s]) continue;
    const int a = t.seg_start((int)s), b = t.seg_end((int)s);
    segs.push_back({t.vx[t.h_src[a]], t.vy[t.h_src[a]], t.vx[t.tgt(b)], t.vy[t.tgt(b)]});
  }
  segs.push_back({c[0], c[1], c[2], c[3]});
  size_t nw = 0;
  for (const auto &b : bx) nw += b.size();
  HostTess nt;
  std::string strict_err;
  if (!build_arrangement(nt, bx, by, segs, false, strict_err)) {
    if (!build_arrangement(nt, bx, by, segs, true, err)) return false;
    nt.not_t_reason = strict_err;
  }
  t = nt;
  if (seg_id) *seg_id = (int)(nw + segs.size() - 1);
  return true;
}

// ---------------------------------------------------------------------------
// LineTes::remove_ivertices / remove_lvertices / remove_xvertices (lib/ttessel.cpp:982-1296):
// the passes that turn an arrangement of segments into a T-tessellation.  They work on the
// list of maximal segments -- move one end, drop a segment, break one in two -- and build the
// arrangement again after every change (same device as insert_segment).
// ---------------------------------------------------------------------------
namespace {
struct Soup {
  std::vector<std::vector<double>> bx, by, segs;
  int nw = 0;
};
static bool soup_of(const HostTess &t, Soup &S, std::string &err) {
  S.bx = t.win_bx; S.by = t.win_by;
  if (S.bx.empty() && !t.win_x.empty()) { S.bx = {t.win_x}; S.by = {t.win_y}; }
  if (S.bx.empty()) { err = "no window"; return false; }
  S.nw = 0;
  for (const auto &b : S.bx) S.nw += (int)b.size();
  S.segs.clear();
  for (size_t s = 0; s < t.s_he.size(); s++) {
    if (!t.s_alive[s] || t.s_bnd[s]) continue;
    const int a = t.seg_start((int)s), b = t.seg_end((int)s);
    S.segs.push_back({t.vx[t.h_src[a]], t.vy[t.h_src[a]], t.vx[t.tgt(b)], t.vy[t.tgt(b)]});
  }
  return true;
}
// after this the segment ids are nw + (index in S.segs)
static bool rebuild(HostTess &t, const Soup &S, std::string &err) {
  HostTess nt;
  std::string strict_err;
  if (!build_arrangement(nt, S.bx, S.by, S.segs, false, strict_err)) {
    if (!build_arrangement(nt, S.bx, S.by, S.segs, true, err)) return false;
    nt.not_t_reason = strict_err;
  }
  t = nt;
  return true;
}
struct Around {   // halfedges arriving at each vertex
  std::vector<std::vector<int>> in;
  explicit Around(const HostTess &t) : in(t.vx.size()) {
    for (size_t h = 0; h < t.h_src.size(); h++) if (t.h_alive[h]) in[t.tgt((int)h)].push_back((int)h);
  }
  int degree(int v) const { return (int)in[v].size(); }
};
// the first edge met when the edge e is prolonged beyond its target (ray_exit_face on the face
// ahead, lib/ttessel.cpp:1017-1022, precompute_lengthening :3800-3830): distance and point
static double prolong(const HostTess &t, int e, double &px, double &py) {
  const int v = t.tgt(e), u = t.h_src[e];
  const double ox = t.vx[v], oy = t.vy[v];
  double dx = ox - t.vx[u], dy = oy - t.vy[u];
  const double dl = sqrt(dx * dx + dy * dy);
  dx /= dl; dy /= dl;
  double best = HUGE_VAL;
  for (size_t h = 0; h < t.h_src.size(); h += 2) {
    if (!t.h_alive[h] || (int)h == (e & ~1)) continue;
    const int a = t.h_src[h], b = t.h_src[h ^ 1];
    if (a == v || b == v) continue;   // edges that leave the vertex itself
    const double ex = t.vx[b] - t.vx[a], ey = t.vy[b] - t.vy[a];
    const double den = dx * ey - dy * ex;
    if (fabs(den) <= 1e-14 * sqrt(ex * ex + ey * ey)) continue;
    const double tr = ((t.vx[a] - ox) * ey - (t.vy[a] - oy) * ex) / den;
    const double te = ((t.vx[a] - ox) * dy - (t.vy[a] - oy) * dx) / den;
    if (tr > 0 && te >= -1e-12 && te <= 1 + 1e-12 && tr < best) {
      best = tr;
      const double tc = fmax(0.0, fmin(1.0, te));
      px = t.vx[a] + tc * ex; py = t.vy[a] + tc * ey;   // on the edge that is hit
    }
  }
  return best;
}
// which end of its maximal segment the target of e is (0: the segment's source, 1: its target)
static int seg_end_at(const HostTess &t, int e) {
  const int s = t.h_seg[e];
  return t.h_src[t.seg_start(s)] == t.tgt(e) ? 0 : 1;
}
static bool on_window_boundary(const HostTess &t, const Around &A, int v) {
  for (int h : A.in[v]) if (t.s_bnd[t.h_seg[h]]) return true;
  return false;
}
static bool is_x_vertex(const HostTess &t, const Around &A, int v) {   // :4058-4067
  if (A.degree(v) != 4) return false;
  int through = 0;
  for (int h : A.in[v]) if (t.h_next_hf[h] >= 0 && t.h_src[t.h_next_hf[h]] == v) through++;
  int sg[4], n = 0;
  for (int h : A.in[v]) sg[n++] = t.h_seg[h];
  std::sort(sg, sg + 4);
  return through == 4 && sg[0] == sg[1] && sg[2] == sg[3] && sg[1] != sg[2];
}
static bool selected_x_vertex(const HostTess &t, const Around &A, int v, bool exclude_boundaries, bool exclude_ineighbours) {
  if (!is_x_vertex(t, A, v)) return false;
  if (exclude_boundaries && on_window_boundary(t, A, v)) return false;
  if (exclude_ineighbours)
    for (int h : A.in[v]) if (A.degree(t.h_src[h]) == 1) return false;   // :4073-4093
  return true;
}
}  // namespace

int count_xvertices(const HostTess &t, bool exclude_boundaries, bool exclude_ineighbours) {
  Around A(t);
  int n = 0;
  for (size_t v = 0; v < t.vx.size(); v++)
    if (t.v_alive[v] && selected_x_vertex(t, A, (int)v, exclude_boundaries, exclude_ineighbours)) n++;
  return n;
}

int count_vertices_of_degree(const HostTess &t, int degree, bool internal_only) {
  Around A(t);
  int n = 0;
  for (size_t v = 0; v < t.vx.size(); v++)
    if (t.v_alive[v] && A.degree((int)v) == degree && !(internal_only && on_window_boundary(t, A, (int)v))) n++;
  return n;
}

// LineTes::is_a_T_tessellation, vertex by vertex (lib/ttessel.cpp:1307-1329, is_a_T_vertex
// :4002-4046): a vertex off the window boundary must have degree 3 with exactly one incident
// segment ending there; one on the boundary may also have degree 2
int count_non_t_vertices(const HostTess &t, double *first_x, double *first_y) {
  Around A(t);
  int n = 0;
  for (size_t v = 0; v < t.vx.size(); v++) {
    if (!t.v_alive[v]) continue;
    int terminal = 0;
    for (int h : A.in[v]) terminal += t.h_next_hf[h] < 0;
    const bool is_t = A.degree((int)v) == 3 && terminal == 1;
    const bool ok = is_t || (on_window_boundary(t, A, (int)v) && A.degree((int)v) == 2);
    if (!ok && n++ == 0) { if (first_x) *first_x = t.vx[v]; if (first_y) *first_y = t.vy[v]; }
  }
  return n;
}

bool remove_ivertices(HostTess &t, size_t imax, CleanStats *st, std::string &err) {
  if (imax == 0) imax = (size_t)-1;
  Soup S;
  if (!soup_of(t, S, err) || !rebuild(t, S, err)) return false;
  CleanStats loc;
  const size_t guard = 8 * (t.h_src.size() + 16);   // every removal takes a free end away for good
  for (size_t counter = 1;; counter++) {
    if (counter > guard) { err = "remove_ivertices does not terminate (a prolonged end did not land on its edge)"; return false; }
    Around A(t);
    double best = HUGE_VAL, bx = 0, by = 0;
    int be = -1;
    bool shorten = false;
    for (size_t v = 0; v < t.vx.size(); v++) {
      if (!t.v_alive[v] || A.degree((int)v) != 1) continue;
      const int e = A.in[v][0];
      double px = 0, py = 0;
      const double lm = t.h_len[e], lp = prolong(t, e, px, py);
      const double changed = lm <= lp ? lm : lp;   // shortening preferred on a tie (:1026-1032)
      if (changed < best) { best = changed; be = e; shorten = lm <= lp; bx = px; by = py; }
    }
    if (be < 0) break;
    const int k = t.h_seg[be] - S.nw, end = seg_end_at(t, be);
    if (shorten) {
      if (t.s_nedges[t.h_seg[be]] == 1) S.segs.erase(S.segs.begin() + k);
      else { S.segs[k][2 * end] = t.vx[t.h_src[be]]; S.segs[k][2 * end + 1] = t.vy[t.h_src[be]]; }
      loc.removed_length += best;
    } else {
      S.segs[k][2 * end] = bx; S.segs[k][2 * end + 1] = by;
      loc.added_length += best;
    }
    loc.n_removed++;
    if (!rebuild(t, S, err)) return false;
    if (counter > imax) break;   // as written (:1068-1078): imax + 1 removals
  }
  if (st) *st = loc;
  return true;
}

bool remove_lvertices(HostTess &t, size_t imax, CleanStats *st, std::string &err) {
  if (imax == 0) imax = (size_t)-1;
  Soup S;
  if (!soup_of(t, S, err) || !rebuild(t, S, err)) return false;
  CleanStats loc;
  const size_t guard = 8 * (t.h_src.size() + 16);
  for (size_t counter = 1;; counter++) {
    if (counter > guard) { err = "remove_lvertices does not terminate (a prolonged end did not land on its edge)"; return false; }
    Around A(t);
    double best = HUGE_VAL, bx = 0, by = 0;
    int be = -1;
    for (size_t v = 0; v < t.vx.size(); v++) {
      if (!t.v_alive[v] || A.degree((int)v) != 2 || on_window_boundary(t, A, (int)v)) continue;
      const int e1 = A.in[v][0], e2 = A.in[v][1];
      // two aligned edges are one line broken at a vertex, not a corner
      const double ux = t.vx[v] - t.vx[t.h_src[e1]], uy = t.vy[v] - t.vy[t.h_src[e1]];
      const double wx = t.vx[v] - t.vx[t.h_src[e2]], wy = t.vy[v] - t.vy[t.h_src[e2]];
      if (fabs(ux * wy - uy * wx) <= 1e-12 * t.h_len[e1] * t.h_len[e2]) continue;
      double p1x = 0, p1y = 0, p2x = 0, p2y = 0;
      const double l1 = prolong(t, e1, p1x, p1y), l2 = prolong(t, e2, p2x, p2y);
      const bool first = l1 < l2;   // compare_ivertices (:1135)
      const double l = first ? l1 : l2;
      if (l < best) { best = l; be = first ? e1 : e2; bx = first ? p1x : p2x; by = first ? p1y : p2y; }
    }
    if (be < 0 || best == HUGE_VAL) break;
    const int k = t.h_seg[be] - S.nw, end = seg_end_at(t, be);
    S.segs[k][2 * end] = bx; S.segs[k][2 * end + 1] = by;
    loc.added_length += best;
    loc.n_removed++;
    if (!rebuild(t, S, err)) return false;
    if (counter > imax) break;
  }
  if (st) *st = loc;
  return true;
}

bool remove_xvertices(HostTess &t, double step, CleanStats *st, std::string &err) {
  if (!(step > 0)) { err = "the step must be positive"; return false; }
  Soup S;
  if (!soup_of(t, S, err) || !rebuild(t, S, err)) return false;
  CleanStats loc;
  for (int guard = 0;; guard++) {
    if (guard > 100000) { err = "remove_xvertices does not terminate"; return false; }
    Around A(t);
    int v = -1;
    for (size_t u = 0; u < t.vx.size() && v < 0; u++)
      if (t.v_alive[u] && selected_x_vertex(t, A, (int)u, true, true)) v = (int)u;
    if (v < 0) break;
    // the segment that is broken: the one with the smaller id; curr arrives at v along its direction
    int curr = -1;
    for (int h : A.in[v])
      if (t.h_dir[h] && (curr < 0 || t.h_seg[h] < t.h_seg[curr])) curr = h;
    const int s = t.h_seg[curr], k = s - S.nw;
    const int other = t.h_next[curr] ^ 1;   // the next halfedge arriving at v, clockwise: on the other segment
    double mel = HUGE_VAL;                   // min_edge_length (:547-557)
    for (size_t h = 0; h < t.h_src.size(); h += 2) if (t.h_alive[h]) mel = fmin(mel, t.h_len[h]);
    double eps = fmin(step, 0.9 * mel);
    // in doubles two ends closer than the builder's tolerance are one vertex: where the shortest
    // edge is that small the end is moved by half of the edge it moves along instead
    double xmin = HUGE_VAL, xmax = -HUGE_VAL, ymin = HUGE_VAL, ymax = -HUGE_VAL;
    for (size_t b = 0; b < S.bx.size(); b++)
      for (size_t i = 0; i < S.bx[b].size(); i++) {
        xmin = fmin(xmin, S.bx[b][i]); xmax = fmax(xmax, S.bx[b][i]); ymin = fmin(ymin, S.by[b][i]); ymax = fmax(ymax, S.by[b][i]);
      }
    const double tol = 1e-7 * fmax(xmax - xmin, ymax - ymin);
    if (eps < tol) eps = fmin(step, 0.5 * t.h_len[other]);
    if (eps < tol) { err = "remove_xvertices: a crossing lies closer to its neighbours than the coordinates resolve"; return false; }
    const int q = t.h_src[other];
    const double ex = t.vx[v] + eps * (t.vx[q] - t.vx[v]) / t.h_len[other];
    const double ey = t.vy[v] + eps * (t.vy[q] - t.vy[v]) / t.h_len[other];
    // the first half, from the segment's source, now ends on the other segment a little aside
    // of the crossing; the second half starts at the crossing (:1233-1252)
    const std::vector<double> half = {S.segs[k][0], S.segs[k][1], ex, ey};
    S.segs[k][0] = t.vx[v]; S.segs[k][1] = t.vy[v];
    S.segs.push_back(half);
    loc.n_removed++;
    if (!rebuild(t, S, err)) return false;
  }
  if (st) *st = loc;
  return true;
}

static bool parse_ratio(const std::string &tok, double &out) {
  size_t sl = tok.find('/');
  std::string num = tok.substr(0, sl), den = sl == std::string::npos ? "1" : tok.substr(sl + 1);
  auto to_ld = [](const std::string &s, long double &v) {
    if (s.empty()) return false;
    size_t i = 0;
    bool neg = false;
    if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; }
    if (i >= s.size()) return false;
    v = 0;
    for (; i < s.size(); i++) {
      if (s[i] < '0' || s[i] > '9') return false;
      v = v * 10.0L + (long double)(s[i] - '0');
    }
    if (neg) v = -v;
    return true;
  };
  long double n, d;
  if (!to_ld(num, n) || !to_ld(den, d) || d == 0) return false;
  out = (double)(n / d);
  return true;
}

bool read_text(HostTess &t, const std::string &text, std::string &err, bool accept_any) {
  std::istringstream is(text);
  std::string line;
  std::vector<std::vector<double>> segs, bx, by;
  while (std::getline(is, line)) {
    std::istringstream ls(line);
    std::string tok;
    std::vector<double> c;
    while (ls >> tok) {
      double v;
      if (!parse_ratio(tok, v)) { err = "bad coordinate '" + tok + "'"; return false; }
      c.push_back(v);
    }
    if (c.empty()) continue;
    if (c.size() % 2) { err = "Uneven number of coordinates on the current line"; return false; }
    if (c.size() > 4) {
      // a boundary of the window (lib/ttessel.cpp:1395-1415): counter-clockwise = the outer boundary
      // of a new polygon, clockwise = a hole of the current one
      std::vector<double> x, y;
      for (size_t i = 0; i < c.size(); i += 2) { x.push_back(c[i]); y.push_back(c[i + 1]); }
      double a = 0;
      for (size_t i = 0; i < x.size(); i++) { size_t j = (i + 1) % x.size(); a += x[i] * y[j] - x[j] * y[i]; }
      if (a == 0) { err = "Invalid orientation of a domain ccb"; return false; }
      if (a < 0 && bx.empty()) { err = "Invalid orientation of a domain ccb: a hole (clockwise) before any outer boundary"; return false; }
      bx.push_back(x); by.push_back(y);
    } else {
      if (bx.empty()) { err = "segment before window"; return false; }
      segs.push_back(c);
    }
  }
  if (build_from_boundaries(t, bx, by, segs, err)) return true;
  if (!accept_any) return false;
  // LineTes::read inserts whatever the file holds, one clipped segment at a time
  // (lib/ttessel.cpp:1417-1437): the arrangement of the segments that hit the window
  const std::string strict_err = err;
  std::vector<std::vector<double>> clipped;
  for (const auto &c : segs) {
    double o[4];
    std::string e2;
    if (clip_segment(bx, by, c.data(), o, e2)) clipped.push_back({o[0], o[1], o[2], o[3]});
  }
  err.clear();
  if (!build_arrangement(t, bx, by, clipped, true, err)) return false;
  t.not_t_reason = strict_err;
  return true;
}

// a double is m*2^e exactly: print it as the rational num/den the format wants
static std::string exact_ratio(double v) {
  if (v == 0.0) return "0/1";
  int e;
  double m = frexp(v, &e);                // v = m * 2^e, 0.5 <= |m| < 1
  long long M = (long long)ldexp(m, 53);  // 53-bit integer mantissa
  e -= 53;
  while (M % 2 == 0 && e < 0) { M /= 2; e++; }
  char buf[400];
  if (e >= 0) snprintf(buf, sizeof buf, "%.0Lf/1", (long double)M * ldexpl(1.0L, e));
  else snprintf(buf, sizeof buf, "%lld/%.0Lf", M, ldexpl(1.0L, -e));
  return buf;
}

std::string write_text(const HostTess &t) {
  std::ostringstream os;
  if (t.win_bx.empty()) {
    for (size_t i = 0; i < t.win_x.size(); i++)
      os << (i ? " " : "") << exact_ratio(t.win_x[i]) << " " << exact_ratio(t.win_y[i]);
    os << "\n";
  }
  for (size_t b = 0; b < t.win_bx.size(); b++) {
    for (size_t i = 0; i < t.win_bx[b].size(); i++)
      os << (i ? " " : "") << exact_ratio(t.win_bx[b][i]) << " " << exact_ratio(t.win_by[b][i]);
    os << "\n";
  }
  for (size_t s = 0; s < t.s_he.size(); s++) {
    if (!t.s_alive[s] || t.s_bnd[s]) continue;
    int a = t.seg_start((int)s), b = t.seg_end((int)s);
    os << exact_ratio(t.vx[t.h_src[a]]) << " " << exact_ratio(t.vy[t.h_src[a]]) << " "
       << exact_ratio(t.vx[t.tgt(b)]) << " " << exact_ratio(t.vy[t.tgt(b)]) << "\n";
  }
  return os.str();
}

}  // namespace lite
