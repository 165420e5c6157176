// host_capi.cpp — implementation of include/lite_b200_host.h over lite::HostTess.
#include <string.h>

#include <string>

#include "../../include/lite_b200_host.h"
#include "host_tess.h"
#include "smf_chain.h"

extern "C" int lite_set_error_(int code, const char *msg);  // defined in capi.cu

struct lite_host_tess {
  lite::HostTess t;
  std::vector<int> he_ids;  // export id -> internal id of the last export
};

static const char *NOT_T = "the segments inserted so far do not form a T-tessellation (lite_host_is_t_tessellation): the arrangement can be inspected, written and inserted into, nothing else";

static void remember_ids(lite_host_tess *h) {
  h->he_ids.clear();
  for (size_t i = 0; i < h->t.h_src.size(); i++)
    if (h->t.h_alive[i]) h->he_ids.push_back((int)i);
}

extern "C" int lite_host_create_rect(lite_host_tess **out, double x0, double y0, double x1, double y1) {
  if (!out || !(x1 > x0) || !(y1 > y0)) return lite_set_error_(LITE_EINVAL, "bad rectangle");
  lite_host_tess *h = new lite_host_tess();
  h->t.insert_window_rect(x0, y0, x1, y1);
  remember_ids(h);
  *out = h;
  return 0;
}

extern "C" int lite_host_create_from_text(lite_host_tess **out, const char *text) {
  if (!out || !text) return lite_set_error_(LITE_EINVAL, "null argument");
  lite_host_tess *h = new lite_host_tess();
  std::string err;
  if (!lite::read_text(h->t, text, err)) {
    delete h;
    return lite_set_error_(LITE_EINVAL, err.c_str());
  }
  remember_ids(h);
  *out = h;
  return 0;
}

extern "C" int lite_host_destroy(lite_host_tess *t) {
  delete t;
  return 0;
}

extern "C" int lite_host_crtt_run(lite_host_tess *h, double tau, uint64_t n_steps, uint64_t seed,
                                  int64_t counts6[6]) {
  if (h && h->t.arrangement_only) return lite_set_error_(LITE_ESTATE, NOT_T);
  if (h && h->t.has_holes) return lite_set_error_(LITE_ESTATE, "the host model reads, checks, exports and writes tessellations of windows with holes but does not move them");
  if (!h || !(tau > 0)) return lite_set_error_(LITE_EINVAL, "bad argument");
  lite::ChainCounts c = lite::crtt_chain(h->t, tau, n_steps, seed);
  if (counts6) memcpy(counts6, c.c, sizeof c.c);
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_update_split(lite_host_tess *h, int32_t he, double x, double angle) {
  if (h && h->t.arrangement_only) return lite_set_error_(LITE_ESTATE, NOT_T);
  if (h && h->t.has_holes) return lite_set_error_(LITE_ESTATE, "the host model reads, checks, exports and writes tessellations of windows with holes but does not move them");
  if (!h || he < 0 || he >= (int)h->he_ids.size()) return lite_set_error_(LITE_EINVAL, "bad halfedge id");
  int e = h->he_ids[he];
  if (!h->t.f_inside[h->t.h_face[e]]) return lite_set_error_(LITE_EINVAL, "halfedge bounds the outer face");
  lite::SplitMove m = h->t.make_split(e, x, angle);
  if (!m.valid) return lite_set_error_(LITE_EINVAL, "Split constructor: l does not hit e1 / no exit found");
  h->t.update_split(m);
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_update_merge(lite_host_tess *h, int32_t i) {
  if (h && h->t.arrangement_only) return lite_set_error_(LITE_ESTATE, NOT_T);
  if (h && h->t.has_holes) return lite_set_error_(LITE_ESTATE, "the host model reads, checks, exports and writes tessellations of windows with holes but does not move them");
  if (!h || i < 0 || i >= (int)h->t.nb_list.size()) return lite_set_error_(LITE_EINVAL, "bad non-blocking index");
  h->t.update_merge(h->t.s_he[h->t.nb_list[i]]);
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_update_flip(lite_host_tess *h, int32_t i, int side) {
  if (h && h->t.arrangement_only) return lite_set_error_(LITE_ESTATE, NOT_T);
  if (h && h->t.has_holes) return lite_set_error_(LITE_ESTATE, "the host model reads, checks, exports and writes tessellations of windows with holes but does not move them");
  if (!h || i < 0 || i >= (int)h->t.b_list.size()) return lite_set_error_(LITE_EINVAL, "bad blocking index");
  lite::FlipMove f = h->t.make_flip(h->t.flip_halfedge(h->t.b_list[i], side));
  if (!f.valid) return lite_set_error_(LITE_EINVAL, "ray_exit_face: intersection of ray with face boundaries not found");
  h->t.update_flip(f);
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_counts(lite_host_tess *h, int32_t *n_cells, int32_t *n_nb, int32_t *n_b) {
  if (!h) return lite_set_error_(LITE_EINVAL, "null argument");
  if (n_cells) *n_cells = h->t.n_cells();
  if (n_nb) *n_nb = (int)h->t.nb_list.size();
  if (n_b) *n_b = (int)h->t.b_list.size();
  return 0;
}

extern "C" int lite_host_check(lite_host_tess *h) {
  if (!h) return lite_set_error_(LITE_EINVAL, "null argument");
  if (h->t.arrangement_only) return lite_set_error_(LITE_EINVAL, ("not a T-tessellation: " + h->t.not_t_reason).c_str());
  std::string e = h->t.check();
  if (!e.empty()) return lite_set_error_(LITE_EINVAL, e.c_str());
  return 0;
}

extern "C" int lite_host_export(lite_host_tess *h, lite_tess_soa *out) {
  if (!h || !out) return lite_set_error_(LITE_EINVAL, "null argument");
  if (h->t.arrangement_only) return lite_set_error_(LITE_ESTATE, NOT_T);
  *out = *h->t.export_soa();
  remember_ids(h);
  return 0;
}

extern "C" int64_t lite_host_write_text(lite_host_tess *h, char *buf, int64_t cap) {
  if (!h) return -1;
  std::string s = lite::write_text(h->t);
  if (buf && cap > 0) {
    int64_t n = (int64_t)s.size() < cap - 1 ? (int64_t)s.size() : cap - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size() + 1;
}

extern "C" int lite_host_clip_segment(lite_host_tess *h, const double seg[4], double out[4]) {
  if (!h || !seg || !out) return lite_set_error_(LITE_EINVAL, "null argument");
  std::vector<std::vector<double>> bx = h->t.win_bx, by = h->t.win_by;
  if (bx.empty() && !h->t.win_x.empty()) { bx = {h->t.win_x}; by = {h->t.win_y}; }
  std::string err;
  if (!lite::clip_segment(bx, by, seg, out, err)) return lite_set_error_(LITE_EINVAL, ("clip_segment_by_polygon: " + err).c_str());
  return 0;
}

extern "C" int lite_host_insert_segment(lite_host_tess *h, const double seg[4], int32_t *seg_id) {
  if (!h || !seg) return lite_set_error_(LITE_EINVAL, "null argument");
  std::string err;
  int id = -1;
  if (!lite::insert_segment(h->t, seg, &id, err)) return lite_set_error_(LITE_EINVAL, ("insert_segment: " + err).c_str());
  if (seg_id) *seg_id = id;
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_is_t_tessellation(lite_host_tess *h) {
  if (!h) return lite_set_error_(LITE_EINVAL, "null argument");
  return h->t.arrangement_only ? 0 : 1;
}

extern "C" int64_t lite_host_edges(lite_host_tess *h, int64_t cap, double *xy, int32_t *seg, int32_t *next_hf,
                                   int32_t *prev_hf) {
  if (!h) return -1;
  const lite::HostTess &t = h->t;
  std::vector<int> id(t.h_src.size(), -1);
  int64_t n = 0;
  for (size_t e = 0; e < t.h_src.size(); e++) if (t.h_alive[e]) id[e] = (int)n++;
  if (cap < n) return n;
  for (size_t e = 0; e < t.h_src.size(); e++) {
    if (!t.h_alive[e]) continue;
    const int k = id[e], a = t.h_src[e], b = t.h_src[e ^ 1];
    if (xy) { xy[4 * k] = t.vx[a]; xy[4 * k + 1] = t.vy[a]; xy[4 * k + 2] = t.vx[b]; xy[4 * k + 3] = t.vy[b]; }
    if (seg) seg[k] = t.h_seg[e];
    if (next_hf) next_hf[k] = t.h_next_hf[e] >= 0 ? id[t.h_next_hf[e]] : -1;
    if (prev_hf) prev_hf[k] = t.h_prev_hf[e] >= 0 ? id[t.h_prev_hf[e]] : -1;
  }
  return n;
}

extern "C" int lite_host_remove_ivertices(lite_host_tess *h, int64_t imax, double *removed_length, double *added_length) {
  if (!h || imax < 0) return lite_set_error_(LITE_EINVAL, "bad argument");
  std::string err;
  lite::CleanStats st;
  if (!lite::remove_ivertices(h->t, (size_t)imax, &st, err)) return lite_set_error_(LITE_EINVAL, ("remove_ivertices: " + err).c_str());
  if (removed_length) *removed_length = st.removed_length;
  if (added_length) *added_length = st.added_length;
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_remove_lvertices(lite_host_tess *h, int64_t imax, double *added_length) {
  if (!h || imax < 0) return lite_set_error_(LITE_EINVAL, "bad argument");
  std::string err;
  lite::CleanStats st;
  if (!lite::remove_lvertices(h->t, (size_t)imax, &st, err)) return lite_set_error_(LITE_EINVAL, ("remove_lvertices: " + err).c_str());
  if (added_length) *added_length = st.added_length;
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_remove_xvertices(lite_host_tess *h, double step) {
  if (!h) return lite_set_error_(LITE_EINVAL, "null argument");
  std::string err;
  if (!lite::remove_xvertices(h->t, step, nullptr, err)) return lite_set_error_(LITE_EINVAL, ("remove_xvertices: " + err).c_str());
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_vertex_census(lite_host_tess *h, int32_t *n_free_ends, int32_t *n_corners, int32_t *n_crossings,
                                       int32_t *n_crossings_to_remove) {
  if (!h) return lite_set_error_(LITE_EINVAL, "null argument");
  if (n_free_ends) *n_free_ends = lite::count_vertices_of_degree(h->t, 1, false);
  if (n_corners) *n_corners = lite::count_vertices_of_degree(h->t, 2, true);
  if (n_crossings) *n_crossings = lite::count_xvertices(h->t, false, false);
  if (n_crossings_to_remove) *n_crossings_to_remove = lite::count_xvertices(h->t, true, true);
  return 0;
}

extern "C" int lite_host_non_t_vertices(lite_host_tess *h, int32_t *n, double first_xy[2]) {
  if (!h || !n) return lite_set_error_(LITE_EINVAL, "null argument");
  double x = 0, y = 0;
  *n = lite::count_non_t_vertices(h->t, &x, &y);
  if (first_xy) { first_xy[0] = x; first_xy[1] = y; }
  return 0;
}

// One SMF chain brought back from the GPU (lite_smf_export, capi.cu) -> host model.
// Ids are kept; elements on the chain's free lists stay dead.
extern "C" int lite_smf_blob_to_host_(char *blob, int cap, lite_host_tess **out) {
  smf::Chain c;
  smf::Header H;
  c.H = &H;
  smf::bind(c, blob, cap);
  smf::load(c);
  if (c.H->cap != cap) return lite_set_error_(LITE_EINVAL, "chain block does not match the ensemble capacity");
  lite_host_tess *h = new lite_host_tess();
  lite::HostTess &t = h->t;
  const int nH = 2 * c.H->n_p, nF = c.H->n_f, nS = c.H->n_s;
  t.h_src.assign(nH, -1); t.h_next.assign(nH, -1); t.h_prev.assign(nH, -1); t.h_face.assign(nH, -1);
  t.h_seg.assign(nH, -1); t.h_next_hf.assign(nH, -1); t.h_prev_hf.assign(nH, -1);
  t.h_dir.assign(nH, 0); t.h_alive.assign(nH, 0); t.h_len.assign(nH, 0.0);
  t.f_he.assign(nF, -1); t.f_inside.assign(nF, 0); t.f_alive.assign(nF, 0);
  t.s_he.assign(nS, -1); t.s_nedges.assign(nS, 0); t.s_pos.assign(nS, -1); t.s_bnd.assign(nS, 0);
  t.s_alive.assign(nS, 0); t.s_kind.assign(nS, 0);
  for (int e = 0; e < nH; e++) {
    if (!smf::pair_alive(c, e >> 1)) continue;
    t.h_alive[e] = 1;
    t.h_next[e] = smf::h_next(c, e); t.h_prev[e] = smf::h_prev(c, e);
    t.h_face[e] = smf::h_face(c, e); t.h_seg[e] = smf::h_seg(c, e);
    t.h_next_hf[e] = smf::h_nhf(c, e); t.h_prev_hf[e] = smf::h_phf(c, e);
    t.h_dir[e] = smf::h_dir(c, e); t.h_len[e] = smf::h_len(c, e);
    const int f = t.h_face[e], s = t.h_seg[e];
    t.f_alive[f] = 1; t.f_inside[f] = (f != 0); t.f_he[f] = c.f_he[f];
    t.s_alive[s] = 1;
  }
  // The chain stores a vertex as the source point of each halfedge leaving it: give the
  // halfedges around one vertex (h -> twin(prev(h))) one id.  The window corners come
  // first (ids 0..3, halfedges 0, 2, 4, 6 leave them), as in insert_window.
  t.vx.clear(); t.vy.clear(); t.v_inc.clear(); t.v_alive.clear();
  for (int pass = 0; pass < 2; pass++) {
    for (int e = 0; e < nH; e += (pass == 0 ? 2 : 1)) {
      if (pass == 0 && e >= 8) break;
      if (!t.h_alive[e] || t.h_src[e] >= 0) continue;
      const int v = (int)t.vx.size();
      t.vx.push_back(c.he[e].x); t.vy.push_back(c.he[e].y); t.v_inc.push_back(e ^ 1); t.v_alive.push_back(1);
      int k = e;
      do { t.h_src[k] = v; k = t.h_prev[k] ^ 1; } while (k != e && t.h_alive[k] && t.h_src[k] < 0);
    }
  }
  for (int s = 0; s < nS; s++) {
    if (!t.s_alive[s]) continue;
    t.s_he[s] = c.seg[s].he; t.s_nedges[s] = c.seg[s].nedges; t.s_bnd[s] = smf::seg_bnd(s);
    t.s_kind[s] = (uint8_t)c.seg[s].kind;
    t.s_pos[s] = (c.seg[s].kind == 1 || c.seg[s].kind == 2) ? (int)c.seg[s].pos : -1;
  }
  t.nb_list.assign(c.nb_list, c.nb_list + c.H->n_nb);
  t.b_list.assign(c.b_list, c.b_list + c.H->n_b);
  t.int_length = c.H->int_length;
  // the window: the four corner vertices are ids 0..3 and never die
  t.win_x = {t.vx[0], t.vx[1], t.vx[2], t.vx[3]};
  t.win_y = {t.vy[0], t.vy[1], t.vy[2], t.vy[3]};
  t.finish_import();
  std::string err = t.check();
  if (!err.empty()) {
    delete h;
    return lite_set_error_(LITE_EINVAL, ("exported chain is inconsistent: " + err).c_str());
  }
  remember_ids(h);
  *out = h;
  return 0;
}
