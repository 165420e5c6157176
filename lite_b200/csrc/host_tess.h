// host_tess.h — CGAL-free dynamic T-tessellation on the host (double DCEL).
//
// This is the host-side state behind the TTessel facade (include/ttessel.h)
// and the input generator for the benchmarks.  It mirrors the reference's
// LineTes/TTessel bookkeeping for hole-free windows (paths relative to
// /root/reference):
//   LineTes::split_edge / merge_edge / add_edge / remove_edge   lib/ttessel.cpp:606-780
//   LineTes::split_from_edge / split_from_vertex / suppress_edge  :131-236
//   TTessel::update(Split/Merge/Flip)                             :1580-1701
//   TTessel::Split / Flip constructors (double arithmetic)        :2090-2132, :2432-2454
// Halfedges are allocated in pairs (twin(h) == h^1); ids of removed elements
// are recycled.  The "container order" seen by the SoA export is id order.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/lite_b200.h"

namespace lite {

struct SplitMove {
  int e1 = -1, e2 = -1;
  double p1x = 0, p1y = 0, p2x = 0, p2y = 0;
  bool valid = false;
};
struct FlipMove {
  int e1 = -1, e2 = -1, e3 = -1;
  double p2x = 0, p2y = 0;
  bool right = false, valid = false;
};

struct Fenwick {
  std::vector<double> t;
  std::vector<double> w;
  void resize(size_t n);
  void set(int i, double v);
  double total() const;
  int find(double target) const;  // smallest i with prefix(i) > target
};

class HostTess {
 public:
  // vertices
  std::vector<double> vx, vy;
  std::vector<int> v_inc;  // one incoming halfedge
  std::vector<uint8_t> v_alive;
  // halfedges
  std::vector<int> h_src, h_next, h_prev, h_face, h_seg, h_next_hf, h_prev_hf;
  std::vector<uint8_t> h_dir, h_alive;
  std::vector<double> h_len;
  // faces
  std::vector<int> f_he;
  std::vector<uint8_t> f_inside, f_alive;
  // segments
  std::vector<int> s_he, s_nedges, s_pos;
  std::vector<uint8_t> s_bnd, s_alive, s_kind;  // kind 0 none, 1 non-blocking, 2 blocking
  std::vector<int> nb_list, b_list;
  double int_length = 0;
  bool ordered_lists = true;  // erase keeps order (reference Seg_sublist::suppress)
  std::vector<double> win_x, win_y;   // the (first) outer boundary of the window, counter-clockwise
  // every boundary of the window: outer boundaries counter-clockwise, holes clockwise, in file order
  std::vector<std::vector<double>> win_bx, win_by;
  bool has_holes = false;  // window with holes or several polygons: read / check / export / write only (no moves)
  // built by insert_segment from segments that do not form a T-tessellation (free ends, crossings,
  // corners): links, lengths and segments are there to be inspected and written; no moves, no export
  bool arrangement_only = false;
  std::string not_t_reason;

  HostTess();
  void insert_window_polygon(const std::vector<double> &x, const std::vector<double> &y);
  void insert_window_rect(double x0, double y0, double x1, double y1);

  inline int twin(int h) const { return h ^ 1; }
  inline int tgt(int h) const { return h_src[h ^ 1]; }
  int n_cells() const;
  int n_alive_halfedges() const;

  // moves
  SplitMove make_split(int e, double x, double angle) const;
  FlipMove make_flip(int e1) const;
  int update_split(const SplitMove &s);  // returns the new halfedge p1->p2
  void update_merge(int e);
  int update_flip(const FlipMove &f);
  int flip_halfedge(int seg, int side) const;  // side 0: twin(start), 1: end
  int seg_start(int s) const;
  int seg_end(int s) const;

  // length-weighted halfedge among those bounding inside faces
  int sample_halfedge(double u01) const;  // u01 in [0,1)
  double weight_total() const { return fen.total(); }

  // after the public arrays have been filled from outside (an SMF chain brought back
  // from the GPU): rebuild the free lists and the sampling tree from the alive flags
  void finish_import();

  // export; the returned struct points into buffers owned by this object
  const lite_tess_soa *export_soa();
  std::string check() const;  // empty if consistent

 private:
  std::vector<int> free_v, free_hp, free_f, free_s;
  Fenwick fen;
  int new_vertex(double x, double y);
  int new_edge_pair(int v1, int v2);
  int new_face(bool inside);
  int new_seg(int he, bool bnd);
  void del_vertex(int v);
  void del_edge_pair(int h);
  void del_face(int f);
  void del_seg(int s);
  void set_len(int h);
  void refresh_weight(int h);
  void set_junction(int e1, int e2);
  void seg_set_handle(int s, int h);
  int split_edge(int e, double px, double py);
  int merge_edge(int e1, int e2);
  int add_edge(int prev1, int prev2, int ext_prev);
  int remove_edge(int e);
  void suppress_edge(int e);
  int degree(int v) const;
  void list_add(int s, int kind);
  void list_remove(int s);
  int ray_exit(int start_he, double a, double b, double w, double px, double py, double dx,
               double dy, int skip0, int skip1, double &hx, double &hy) const;
  // export buffers
  struct Export {
    std::vector<double> vx, vy, he_length;
    std::vector<int32_t> he_src, he_twin, he_next, he_face, he_seg, he_next_hf, he_prev_hf;
    std::vector<uint8_t> he_dir, face_inside, seg_bnd;
    std::vector<int32_t> f_off, f_cnt, f_list, seg_start, seg_end, seg_n, nb, b;
    lite_tess_soa soa;
  } ex;
};

// SMFChain restricted to the CRTT energy theta*minus_is_segment_internal with
// theta = log(tau) (reference demos/pseudo_lik_inference.cpp:56-69,
// lib/ttessel.cpp:3143-3311): the input generator of the benchmarks.
struct ChainCounts {
  int64_t c[6] = {0, 0, 0, 0, 0, 0};
};
ChainCounts crtt_chain(HostTess &t, double tau, uint64_t n_steps, uint64_t seed,
                       double p_split = 0.33, double p_merge = 0.33);

// Build a T-tessellation from its maximal segments, all at once (the reference reads a
// file by LineTes::insert_segment on each line, lib/ttessel.cpp:904-980, which handles any
// order; successive splits cannot, because T-tessellations contain cycles of segments each
// ending on the next).  window: a hole-free polygon, counter-clockwise; segs: x0 y0 x1 y1
// per internal segment, every end point on the interior of another segment or of a window
// side.  Anything else (crossings, L/X/I-vertices, ends in the void) is an error.
bool build_from_segments(HostTess &t, const std::vector<double> &wx, const std::vector<double> &wy,
                         const std::vector<std::vector<double>> &segs, std::string &err);
// The same for a window given by all its boundaries (outer boundaries counter-clockwise, holes
// clockwise: the window is on the left of every boundary edge, as LineTes::read requires,
// lib/ttessel.cpp:1395-1415).  Faces that surround a hole no segment touches keep the hole's
// boundary as an inner boundary: its halfedges carry the face's id and are not listed on its outer
// boundary, which is how lite_tess_soa describes a face with holes.  Such a tessellation can be
// checked, exported (hence evaluated on the GPU) and written; the moves of the host model do not
// handle inner boundaries and refuse it.
bool build_from_boundaries(HostTess &t, const std::vector<std::vector<double>> &bx,
                           const std::vector<std::vector<double>> &by,
                           const std::vector<std::vector<double>> &segs, std::string &err);

// The same builder, strict (a T-tessellation or an error) or lenient (any arrangement of the
// segments: sets HostTess::arrangement_only).
bool build_arrangement(HostTess &t, const std::vector<std::vector<double>> &bx,
                       const std::vector<std::vector<double>> &by,
                       const std::vector<std::vector<double>> &segs, bool lenient, std::string &err);
// clip_segment_by_polygon for a union of polygons with holes (lib/ttessel.cpp:3762-3786): the
// longest piece of seg (x0 y0 x1 y1) inside the open window; false if it does not hit it.
bool clip_segment(const std::vector<std::vector<double>> &bx, const std::vector<std::vector<double>> &by,
                  const double seg[4], double out[4], std::string &err);
// LineTes::insert_segment (lib/ttessel.cpp:904-980); seg_id receives the id of the new segment.
// Every handle into t is invalidated (the arrangement is rebuilt).
bool insert_segment(HostTess &t, const double seg[4], int *seg_id, std::string &err);

// LineTes::remove_ivertices / remove_lvertices / remove_xvertices (lib/ttessel.cpp:982-1296) on the
// arrangement a sequence of insert_segment calls left: free ends are cut back or prolonged to the
// next edge (whichever changes the smaller length, smallest change first), corners are prolonged,
// crossings are broken into two T-vertices `step` apart.  imax as in the reference (0: all).
struct CleanStats {
  int n_removed = 0;
  double removed_length = 0, added_length = 0;
};
bool remove_ivertices(HostTess &t, size_t imax, CleanStats *st, std::string &err);
bool remove_lvertices(HostTess &t, size_t imax, CleanStats *st, std::string &err);
bool remove_xvertices(HostTess &t, double step, CleanStats *st, std::string &err);
// LineTes::number_of_xvertices (:272-278); vertices of a given degree (1: free ends)
int count_xvertices(const HostTess &t, bool exclude_boundaries, bool exclude_ineighbours);
int count_vertices_of_degree(const HostTess &t, int degree, bool internal_only);
// vertices that keep the arrangement from being a T-tessellation (:1307-1329): free ends, corners
// and crossings, and also what no cleaning pass removes (three ends on one point, two ends on the
// same point of a third segment); the first one's position if asked for
int count_non_t_vertices(const HostTess &t, double *first_x, double *first_y);

// LineTes::read / write text format (lib/ttessel.cpp:1331-1441), double precision
// accept_any: a file whose segments do not form a T-tessellation is read as the arrangement of
// its segments, each clipped by the window (what LineTes::read does with any file); otherwise it
// is an error
bool read_text(HostTess &t, const std::string &text, std::string &err, bool accept_any = false);
std::string write_text(const HostTess &t);

}  // namespace lite
