// host_capi.cpp — implementation of include/lite_b200_host.h over lite::HostTess.
#include <string.h>

#include <string>

#include "../../include/lite_b200_host.h"
#include "host_tess.h"

extern "C" int lite_set_error_(int code, const char *msg);  // defined in capi.cu

struct lite_host_tess {
  lite::HostTess t;
  std::vector<int> he_ids;  // export id -> internal id of the last export
};

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
  if (!h || !(tau > 0)) return lite_set_error_(LITE_EINVAL, "bad argument");
  lite::ChainCounts c = lite::crtt_chain(h->t, tau, n_steps, seed);
  if (counts6) memcpy(counts6, c.c, sizeof c.c);
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_update_split(lite_host_tess *h, int32_t he, double x, double angle) {
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
  if (!h || i < 0 || i >= (int)h->t.nb_list.size()) return lite_set_error_(LITE_EINVAL, "bad non-blocking index");
  h->t.update_merge(h->t.s_he[h->t.nb_list[i]]);
  remember_ids(h);
  return 0;
}

extern "C" int lite_host_update_flip(lite_host_tess *h, int32_t i, int side) {
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
  std::string e = h->t.check();
  if (!e.empty()) return lite_set_error_(LITE_EINVAL, e.c_str());
  return 0;
}

extern "C" int lite_host_export(lite_host_tess *h, lite_tess_soa *out) {
  if (!h || !out) return lite_set_error_(LITE_EINVAL, "null argument");
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
