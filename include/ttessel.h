/* ttessel.h — source-compatible C++ facade of LiTe's pseudo-likelihood path.
 *
 * Drop-in for the subset of the reference header (kien-kieu/lite
 * include/ttessel.h) that the pseudo-likelihood hot path and its callers use:
 *   TTessel with Split / Merge / Flip, update(), propose_*, split_sample,
 *   all_merges, all_flips                      reference :500-795
 *   CatItems / CatVector / CatMatrix            :806-926
 *   Features, Energy                            :932-1020
 *   SMFChain, ModifCounts                       :1027-1103
 *   PseudoLikDiscrete, PLInferenceNOIS          :1201-1313
 *   the built-in feature functions              :1384-1405
 * Same class names, member names, argument meaning and error behaviour; no
 * CGAL.  Geometry is held in doubles by a host model (lite_b200/csrc/host_tess.h);
 * every pseudo-likelihood computation (statistic variations of the candidate
 * moves, value / gradient / Hessian) runs on the GPU through the C ABI of
 * include/lite_b200.h.  There is no CPU fallback: Energy only accepts the built-in
 * feature functions, which it recognises by address and maps to device feature
 * ids (an unknown function pointer is a std::invalid_argument when the energy is
 * first evaluated).
 *
 * Deliberate differences from the reference, all documented in DESIGN.md:
 *   - handles are (tessellation, id) pairs, not CGAL iterators; `e->source()->point()`
 *     style access works, CGAL-specific members do not exist;
 *   - the host model's moves need a single hole-free polygon as window; windows with holes or
 *     of several polygons are read, built by insert_segment, checked, written and evaluated on
 *     the device;
 *   - insert_segment and remove_[ilx]vertices rebuild the arrangement (handles obtained before
 *     are invalidated); an arrangement that is not a T-tessellation can be inspected, written,
 *     inserted into and cleaned, nothing else;
 *   - `rnd` is a CGAL::Random look-alike over std::mt19937_64 (CGAL's stream is
 *     un-vendored and unpinned by the reference);
 *   - AddSplits() samples on the device with a counter-based Philox stream.
 */
#ifndef LITE_B200_TTESSEL_H
#define LITE_B200_TTESSEL_H

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "lite_b200.h"

#ifndef CGAL_PI
#define CGAL_PI 3.14159265358979323846
#endif

typedef std::size_t Size;

namespace CGAL {
/* Stand-in for CGAL::Random (reference include/ttessel.h:139): same calls. */
class Random {
 public:
  Random() : g_(5489u), seed_(0) {}
  explicit Random(unsigned int seed) : g_(seed), seed_(seed) {}
  double get_double(double lower = 0.0, double upper = 1.0) {
    return lower + (upper - lower) * ((double)(g_() >> 11) * (1.0 / 9007199254740992.0));
  }
  int get_int(int lower, int upper) {  // uniform on [lower, upper)
    return lower + (int)(get_double() * (double)(upper - lower));
  }
  unsigned int get_seed() const { return seed_; }
  std::uint64_t get_bits64() { return g_(); }
 private:
  std::mt19937_64 g_;
  unsigned int seed_;
};
}  // namespace CGAL

/* The global generator of the reference (include/ttessel.h:139): the caller sets
 * it before any stochastic call (docfiles/start.md:100-102). */
extern CGAL::Random *rnd;

/* ---- kernel objects (doubles) ------------------------------------------- */
class Point2 {
 public:
  Point2() : x_(0), y_(0) {}
  Point2(double x, double y) : x_(x), y_(y) {}
  double x() const { return x_; }
  double y() const { return y_; }
  bool operator==(const Point2 &o) const { return x_ == o.x_ && y_ == o.y_; }
  bool operator!=(const Point2 &o) const { return !(*this == o); }
 private:
  double x_, y_;
};
std::ostream &operator<<(std::ostream &, const Point2 &);

class Segment {
 public:
  Segment() {}
  Segment(const Point2 &s, const Point2 &t) : s_(s), t_(t) {}
  const Point2 &source() const { return s_; }
  const Point2 &target() const { return t_; }
  double squared_length() const {
    double dx = t_.x() - s_.x(), dy = t_.y() - s_.y();
    return dx * dx + dy * dy;
  }
 private:
  Point2 s_, t_;
};

class Rectangle {  /* CGAL::Iso_rectangle_2 look-alike */
 public:
  Rectangle() {}
  Rectangle(const Point2 &p, const Point2 &q) : lo_(p), hi_(q) {}
  double xmin() const { return lo_.x(); }
  double ymin() const { return lo_.y(); }
  double xmax() const { return hi_.x(); }
  double ymax() const { return hi_.y(); }
 private:
  Point2 lo_, hi_;
};

class Polygon {  /* CGAL::Polygon_2 look-alike: a vertex list */
 public:
  typedef std::vector<Point2>::const_iterator Vertex_const_iterator;
  Polygon() {}
  template <class It>
  Polygon(It b, It e) : v_(b, e) {}
  void push_back(const Point2 &p) { v_.push_back(p); }
  Size size() const { return v_.size(); }
  bool is_empty() const { return v_.empty(); }
  const Point2 &operator[](Size i) const { return v_[i]; }
  Vertex_const_iterator vertices_begin() const { return v_.begin(); }
  Vertex_const_iterator vertices_end() const { return v_.end(); }
  double area() const;  /* signed */
  void reverse_orientation();
  const std::vector<Point2> &container() const { return v_; }
 private:
  std::vector<Point2> v_;
};
typedef std::vector<Polygon> Polygons;

class HPolygon {  /* CGAL::Polygon_with_holes_2 look-alike */
 public:
  typedef Polygons::const_iterator Hole_const_iterator;
  HPolygon() {}
  explicit HPolygon(const Polygon &outer) : outer_(outer) {}
  template <class It>
  HPolygon(const Polygon &outer, It hb, It he) : outer_(outer), holes_(hb, he) {}
  const Polygon &outer_boundary() const { return outer_; }
  Hole_const_iterator holes_begin() const { return holes_.begin(); }
  Hole_const_iterator holes_end() const { return holes_.end(); }
  Size number_of_holes() const { return holes_.size(); }
 private:
  Polygon outer_;
  Polygons holes_;
};
typedef std::vector<HPolygon> HPolygons;

namespace lite {
class HostTess;
struct DeviceMirror;
}

/* ---- LineTes: the part of the reference class the path's callers touch --- */
class LineTes {
 public:
  class Seg;
  class Halfedge_handle;
  class Vertex_handle {
   public:
    Vertex_handle() : t_(0), id_(-1) {}
    Vertex_handle(const LineTes *t, int id) : t_(t), id_(id) {}
    const Vertex_handle *operator->() const { return this; }
    Point2 point() const;
    Size degree() const;
    std::vector<Halfedge_handle> incident_halfedges() const;  /* arriving at the vertex, clockwise */
    const LineTes *tess() const { return t_; }
    int id() const { return id_; }
    bool operator==(const Vertex_handle &o) const { return id_ == o.id_ && t_ == o.t_; }
    bool operator!=(const Vertex_handle &o) const { return !(*this == o); }
   private:
    const LineTes *t_;
    int id_;
  };
  class Face_handle {
   public:
    Face_handle() : t_(0), id_(-1) {}
    Face_handle(const LineTes *t, int id) : t_(t), id_(id) {}
    const Face_handle *operator->() const { return this; }
    bool data() const;          /* true for faces inside the window (:125-127) */
    bool is_unbounded() const { return !data(); }
    Halfedge_handle outer_ccb() const;
    std::vector<Halfedge_handle> holes() const;  /* one halfedge of each inner boundary */
    const LineTes *tess() const { return t_; }
    int id() const { return id_; }
    bool operator==(const Face_handle &o) const { return id_ == o.id_ && t_ == o.t_; }
    bool operator!=(const Face_handle &o) const { return !(*this == o); }
   private:
    const LineTes *t_;
    int id_;
  };
  class Halfedge_handle {  /* LineTes_halfedge payload + DCEL links (:445-450) */
   public:
    Halfedge_handle() : t_(0), id_(-1) {}
    Halfedge_handle(const LineTes *t, int id) : t_(t), id_(id) {}
    const Halfedge_handle *operator->() const { return this; }
    Vertex_handle source() const;
    Vertex_handle target() const;
    Halfedge_handle twin() const;
    Halfedge_handle next() const;
    Halfedge_handle prev() const;
    Face_handle face() const;
    double get_length() const;
    bool get_dir() const;
    Halfedge_handle get_next_hf() const;
    Halfedge_handle get_prev_hf() const;
    Seg segment() const;
    int id() const { return id_; }
    const LineTes *tess() const { return t_; }
    bool is_null() const { return id_ < 0; }
    bool operator==(const Halfedge_handle &o) const { return id_ == o.id_ && (id_ < 0 || t_ == o.t_); }
    bool operator!=(const Halfedge_handle &o) const { return !(*this == o); }
   private:
    const LineTes *t_;
    int id_;
  };
  class Seg {  /* LineTes::Seg (:182-212) */
   public:
    Seg() : t_(0), id_(-1) {}
    Seg(const LineTes *t, int id) : t_(t), id_(id) {}
    const Seg *operator->() const { return this; }
    Halfedge_handle get_halfedge_handle() const;
    Halfedge_handle halfedges_start() const;
    Halfedge_handle halfedges_end() const;
    Size number_of_edges() const;
    Point2 pointSource() const;
    Point2 pointTarget() const;
    std::vector<Point2> list_of_points() const;
    int id() const { return id_; }
    bool operator==(const Seg &o) const { return id_ == o.id_ && t_ == o.t_; }
   private:
    const LineTes *t_;
    int id_;
  };
  typedef std::vector<Seg> Seg_sublist;
  typedef Seg_sublist::const_iterator Seg_sublist_iterator;

  LineTes();
  virtual ~LineTes();
  LineTes(const LineTes &);
  LineTes &operator=(const LineTes &);

  virtual void insert_window(Rectangle);
  virtual void insert_window(Polygon &);
  virtual void insert_window(HPolygons &);  /* :283: read / check / evaluate; no host moves */
  /* :904-980.  The segment is clipped by the window and joins the maximal segments; every
   * handle obtained before is invalidated (the arrangement is rebuilt).  While the segments do
   * not form a T-tessellation the object can be inspected, written and inserted into. */
  Seg insert_segment(Segment);
  bool is_a_T_tessellation(bool verbose = false, std::ostream &out = std::cout); /* :1307-1329 */
  /* :982-1296, the passes that make a T-tessellation of an arrangement of segments */
  void remove_ivertices(Size imax = 0, bool verbose = false);
  void remove_lvertices(Size imax = 0, bool verbose = false);
  void remove_xvertices(double step);
  Size number_of_xvertices(bool exclude_boundaries = false, bool exclude_ineighbours = false); /* :272-278 */
  double min_edge_length();                                                                    /* :547-557 */
  Size number_of_internal_segments();   /* :256-262 */
  Size number_of_window_edges();        /* :521-530 */
  double get_window_perimeter();        /* :533-544 */
  typedef std::vector<Seg> Seg_list;    /* every segment, window sides first (container order) */
  typedef Seg_list::const_iterator Seg_list_iterator;
  Seg_list_iterator segments_begin();
  Seg_list_iterator segments_end();
  bool is_on_boundary(Seg) const;
  Size number_of_vertices() const;
  Size number_of_halfedges() const;
  Size number_of_edges() const { return number_of_halfedges() / 2; }
  Size number_of_faces() const;   /* including the unbounded one */
  Size number_of_segments() const; /* window sides included, as the reference */
  std::vector<Halfedge_handle> halfedges() const; /* container order */
  HPolygons get_window() const { return window_; }
  bool is_on_boundary(Halfedge_handle) const;
  bool is_valid(bool verbose = false);
  void write(std::ostream &) const;   /* lib/ttessel.cpp:1348-1370 */
  void read(std::istream &);          /* :1377-1441 */
  void print_all_segments() const;
  unsigned long version() const { return version_; }
  lite::HostTess *host() const { return impl_; }
  std::vector<Vertex_handle> vertices() const;
  std::vector<Face_handle> faces() const;   /* the unbounded face included */

 protected:
  lite::HostTess *impl_;
  HPolygons window_;
  unsigned long version_;
  Seg_list seg_cache_;
  void touch() { version_++; }
};

#define NULL_HALFEDGE_HANDLE LineTes::Halfedge_handle()

/* lib/ttessel.cpp:3696-3786: the longest piece of the segment inside the open polygon(s);
 * std::domain_error when the segment does not hit them */
Segment clip_segment_by_polygon(Segment, Polygon);
Segment clip_segment_by_polygon(Segment, HPolygon);
Segment clip_segment_by_polygon(Segment, HPolygons);

/* include/ttessel.h:489-498 */
struct ModList {
  std::vector<Point2> del_vertices, add_vertices;
  std::vector<Segment> del_edges, add_edges;
  HPolygons del_faces, add_faces;
  std::vector<std::vector<Point2> > del_segs, add_segs;
};

/* ---- TTessel (:500-795) --------------------------------------------------- */
class TTessel : public LineTes {
 public:
  class Modification {
   public:
    virtual ~Modification() {}
    virtual ModList modified_elements();
  };
  class Split : public Modification {
   public:
    Split();
    Split(Halfedge_handle, double = .5, double = CGAL_PI / 2);
    inline Halfedge_handle get_e1() { return e1; }
    inline Point2 get_p1() { return p1; }
    inline Halfedge_handle get_e2() { return e2; }
    inline Point2 get_p2() { return p2; }
    virtual ModList modified_elements();
    bool is_valid();
    double get_x() const { return x_; }        /* the (x, angle) the split was built from */
    double get_angle() const { return angle_; }
   private:
    Halfedge_handle e1, e2;
    Point2 p1, p2;
    double x_, angle_;
  };
  class Merge : public Modification {
   public:
    Merge();
    Merge(Halfedge_handle);
    inline Halfedge_handle get_e() { return e; }
    virtual ModList modified_elements();
    bool is_valid();
   private:
    Halfedge_handle e;
  };
  class Flip : public Modification {
   public:
    Flip();
    Flip(Halfedge_handle);
    inline Halfedge_handle get_e1() { return e1; }
    inline Halfedge_handle get_e2() { return e2; }
    inline Point2 get_p2() { return p2; }
    inline bool to_right() { return right; }
    virtual ModList modified_elements();
   private:
    Halfedge_handle e1, e2;
    Point2 p2;
    bool right;
  };
  typedef std::vector<Split> Split_list;
  typedef std::vector<Merge> Merge_list;
  typedef std::vector<Flip> Flip_list;
  typedef Split_list::iterator Split_list_iterator;
  typedef Merge_list::iterator Merge_list_iterator;
  typedef Flip_list::iterator Flip_list_iterator;

  TTessel();
  TTessel(LineTes &);   /* :1506-1540: std::domain_error unless the line tessellation is a T-tessellation */
  virtual void insert_window(Rectangle);
  virtual void insert_window(Polygon &);
  virtual void insert_window(HPolygons &);
  Halfedge_handle update(Split);
  Face_handle update(Merge);
  Halfedge_handle update(Flip);
  void clear(bool remove_window = true);
  Split propose_split();
  Merge propose_merge();
  Flip propose_flip();
  Split_list split_sample(Size);
  Split_list poisson_splits(double);
  Merge_list all_merges();
  Flip_list all_flips();
  Seg_sublist_iterator blocking_segments_begin();
  Seg_sublist_iterator blocking_segments_end();
  Seg_sublist_iterator non_blocking_segments_begin();
  Seg_sublist_iterator non_blocking_segments_end();
  HPolygons all_faces();
  Size number_of_blocking_segments();
  Size number_of_non_blocking_segments();
  double get_total_internal_length();
  inline double total_internal_length() { return get_total_internal_length(); }   /* :788 */
  /* :1932-2010, draws from the global `rnd` */
  Halfedge_handle random_halfedge();
  Halfedge_handle length_weighted_random_halfedge();
  Halfedge_handle alt_length_weighted_random_halfedge();
  std::vector<Halfedge_handle> alt_length_weighted_random_halfedge(unsigned int);
  void printRCALI(std::ostream &);   /* :2012-2060, the format califlopp reads */
  bool is_valid(bool verbose = false);
  void print_blocking_segments();
  void print_non_blocking_segments();
  /* EXTENSION: the fast CRTT generator used for the benchmark inputs
   * (SMFChain restricted to theta*minus_is_segment_internal, O(log n) proposals) */
  void simulate_crtt(double tau, unsigned long n_steps, std::uint64_t seed);

 private:
  Seg_sublist blocking_cache_, non_blocking_cache_;
  void refresh_sublists();
};

/* ---- CatItems / CatVector / CatMatrix (:806-926) ---------------------------- */
template <typename T>
struct CatItems {
  std::vector<T> vertices, edges, faces, segs;
  const CatItems<T> &operator+=(const CatItems<T> &incr) { *this = *this + incr; return *this; }
  const CatItems<T> &operator-=(const CatItems<T> &decr) { *this = *this - decr; return *this; }
  bool IsEmpty() { return vertices.empty() && edges.empty() && faces.empty() && segs.empty(); }
};
template <typename T>
const CatItems<T> operator+(const CatItems<T> &x1, const CatItems<T> &x2) {
  CatItems<T> r(x1);
  for (Size i = 0; i != x2.vertices.size(); i++) r.vertices[i] = r.vertices[i] + x2.vertices[i];
  for (Size i = 0; i != x2.edges.size(); i++) r.edges[i] = r.edges[i] + x2.edges[i];
  for (Size i = 0; i != x2.faces.size(); i++) r.faces[i] = r.faces[i] + x2.faces[i];
  for (Size i = 0; i != x2.segs.size(); i++) r.segs[i] = r.segs[i] + x2.segs[i];
  return r;
}
template <typename T>
const CatItems<T> operator*(const double &s, const CatItems<T> &x) {
  CatItems<T> r(x);
  for (Size i = 0; i != x.vertices.size(); i++) r.vertices[i] = s * r.vertices[i];
  for (Size i = 0; i != x.edges.size(); i++) r.edges[i] = s * r.edges[i];
  for (Size i = 0; i != x.faces.size(); i++) r.faces[i] = s * r.faces[i];
  for (Size i = 0; i != x.segs.size(); i++) r.segs[i] = s * r.segs[i];
  return r;
}
template <typename T>
const CatItems<T> operator-(const CatItems<T> &x1, const CatItems<T> &x2) {
  return x1 + (-1.0 * x2);
}
template <typename T>
const CatItems<T> operator*(const CatItems<T> &x1, const CatItems<T> &x2) {  /* element-wise */
  CatItems<T> r(x1);
  for (Size i = 0; i != x2.vertices.size(); i++) r.vertices[i] = r.vertices[i] * x2.vertices[i];
  for (Size i = 0; i != x2.edges.size(); i++) r.edges[i] = r.edges[i] * x2.edges[i];
  for (Size i = 0; i != x2.faces.size(); i++) r.faces[i] = r.faces[i] * x2.faces[i];
  for (Size i = 0; i != x2.segs.size(); i++) r.segs[i] = r.segs[i] * x2.segs[i];
  return r;
}
typedef CatItems<double> CatVector;
typedef CatItems<CatVector> CatMatrix;
typedef std::vector<double> FVector;                 /* flat vector, CatVector order */
typedef std::vector<std::vector<double> > FMatrix;   /* flat matrix, row major */
/* lib/ttessel.cpp:2737-2846.  `fill` is the corrected mapping; the reference's
 * fill() never advances its index in the segs loop (:2745-2746), which only
 * matters with two or more segment features. */
void fill(CatVector &, FVector &);
std::vector<double> asVectorOfDoubles(const CatVector &);
FVector asFVector(const CatVector &);
FMatrix asFMatrix(const CatMatrix &);
std::ostream &operator<<(std::ostream &, const CatVector &);
double sum(const CatVector &);
CatMatrix outer(const CatVector &, const CatVector &);

/* ---- Features / Energy (:932-1020) ------------------------------------------ */
struct Features {
  std::vector<double (*)(Point2, TTessel *)> vertices;
  std::vector<double (*)(Segment, TTessel *)> edges;
  std::vector<double (*)(HPolygon, TTessel *)> faces;
  std::vector<double (*)(std::vector<Point2>, TTessel *)> segs;
};

class Energy {
 public:
  Energy();
  ~Energy();
  Energy(const Energy &);
  Energy &operator=(const Energy &);
  void add_features_vertices(double (*)(Point2, TTessel *));
  void add_features_edges(double (*)(Segment, TTessel *));
  void add_features_faces(double (*)(HPolygon, TTessel *));
  void add_features_segs(double (*)(std::vector<Point2>, TTessel *));
  void add_theta_vertices(double);
  void add_theta_edges(double);
  void add_theta_faces(double);
  void add_theta_segs(double);
  inline void set_theta(CatVector par) { theta = par; }
  void set_ttessel(TTessel *);
  inline TTessel *get_ttessel() { return ttes; }
  inline double get_value() { return value; }
  inline CatVector get_theta() { return theta; }
  inline void add_value(double delta) { value += delta; }
  void del_theta_vertices();
  void del_theta_edges();
  void del_theta_faces();
  void del_theta_segs();
  void del_features_vertices();
  void del_features_edges();
  void del_features_faces();
  void del_features_segs();
  double variation(TTessel::Modification &);
  CatVector statistic_variation(TTessel::Modification &);
  /* device feature ids in flattened CatVector order; throws std::invalid_argument
   * on a feature function that is not one of the built-in ones, and
   * std::logic_error if #theta != #features in a category (UB in the reference,
   * lib/ttessel.cpp:2925-2929) */
  std::vector<int32_t> device_program() const;
  const Features &get_features() const { return features; }

 private:
  TTessel *ttes;
  CatVector theta;
  Features features;
  double value;
  lite::DeviceMirror *mirror;  /* GPU copy of the tessellation for statistic_variation */
};

/* ---- SMFChain (:1027-1103) --------------------------------------------------- */
struct ModifCounts {
  int proposed_S, accepted_S, proposed_M, accepted_M, proposed_F, accepted_F;
};
class SMFChain {
 public:
  enum ModType { SPLIT = 1, MERGE, FLIP };
  SMFChain() : engy(0), p_split(0), p_merge(0), p_flip(0), p_split_merge(0) {}
  SMFChain(Energy *, double, double);
  inline void set_energy(Energy *e) { engy = e; }
  void set_smf_prob(double, double);
  inline Energy *get_energy() { return engy; }
  ModType propose_modif_type();
  double Hasting_ratio(TTessel::Split &, double *);
  double Hasting_ratio(TTessel::Merge &, double *);
  double Hasting_ratio(TTessel::Flip &, double *);
  ModifCounts step(unsigned long int = 1);
 private:
  Energy *engy;
  double p_split, p_merge, p_flip, p_split_merge;
};

/* ---- PseudoLikDiscrete / PLInferenceNOIS (:1201-1313) -------------------------- */
class PseudoLikDiscrete {
 public:
  PseudoLikDiscrete();
  PseudoLikDiscrete(Energy *);
  virtual ~PseudoLikDiscrete();
  void SetEnergy(Energy *);
  void AddSplits(Size);
  void AddGivenSplit(TTessel::Split);
  void ClearSplits();
  inline Energy *GetEnergy() { return engy; }
  double GetValue(CatVector);
  CatVector GetGradient(CatVector);
  CatMatrix GetHessian(CatVector);
  /* EXTENSIONS: one fused device pass for all three; the device context for
   * multi-GPU use (rank/world/NCCL id); the Philox seed of AddSplits */
  void GetValueGradientHessian(CatVector, double *, CatVector *, CatMatrix *);
  void SetDevice(int device, int rank = 0, int world_size = 1, const void *nccl_unique_id = 0);
  void SetSeed(std::uint64_t seed) { seed_ = seed; }
  /* AddSplits draws its positions as TTessel::split_sample does (n independent uniforms, sorted,
   * lib/ttessel.cpp:1968-1991) instead of the default stratified sampler, which keeps each split's
   * law but not the independence between splits (lower variance): see lite_b200.h, LITE_OPT_SAMPLER */
  void SetIndependentSampling(bool on);
  Size NumberOfSplits() const { return (Size)n_splits_; }
  lite_ctx *context() { return ctx_; }

 private:
  PseudoLikDiscrete(const PseudoLikDiscrete &);
  PseudoLikDiscrete &operator=(const PseudoLikDiscrete &);
  Energy *engy;
  lite_ctx *ctx_;
  int device_, rank_, world_;
  std::vector<char> nccl_id_;
  std::uint64_t seed_, offset_;
  std::int64_t n_splits_;
  bool iid_ = false;
  std::vector<int> he_export_;  /* internal halfedge id -> id in the uploaded SoA */
  void ensure_ctx();
  void shape(const CatVector &like, const double *flat, CatVector *out) const;
};

class PLInferenceNOIS : public PseudoLikDiscrete {
 public:
  PLInferenceNOIS(Energy *, bool = false, double = 1.0);
  inline void SetStepSize(double lambda) { stepsize = lambda; }
  inline double GetStepSize() { return stepsize; }
  CatVector GetEstimate();
  inline std::vector<CatVector> GetEstimates() { return estimates; }
  inline std::vector<double> GetValues() { return values; }
  void Step(unsigned int = 1);
  unsigned int Run(double = 0.05, unsigned int = 100);
 private:
  double stepsize;
  CatVector previous_theta;
  double previous_lpl;
  bool store_path;
  std::vector<CatVector> estimates;
  std::vector<double> values;
};

/* ---- built-in feature functions (:1384-1405, lib/ttessel.cpp:4124-4434) --------
 * Real functions (a caller may evaluate them on the host, in doubles); Energy
 * recognises them by address. */
double is_point_inside_window(Point2, LineTes *);
double is_point_inside_window(Point2, TTessel *);
double edge_length(Segment, TTessel *);
double seg_number(std::vector<Point2>, TTessel *);
double is_segment_internal(std::vector<Point2>, TTessel *);
double minus_is_segment_internal(std::vector<Point2>, TTessel *);
double face_number(HPolygon, TTessel *);
double face_area_2(HPolygon, TTessel *);
double face_perimeter(HPolygon, TTessel *);
double face_shape(HPolygon, TTessel *);
double face_sum_of_angles(HPolygon, TTessel *);
double min_angle(HPolygon, TTessel *);
double segment_size_2(std::vector<Point2>, TTessel *);
double angle_between_vectors(double v1x, double v1y, double v2x, double v2y);

/* ---- helpers of the reference header (:1319-1370) --------------------------- */
bool are_aligned(Point2 p, Point2 q, Point2 r, bool verbose = false);          /* :3556-3577 */
bool is_inside(Point2, HPolygon &);                                              /* :3583-3598: strictly inside */
bool is_inside(Point2, HPolygons &);
Segment clip_segment_by_convex_polygon(Segment, Polygon);                        /* :3620-3641 */
LineTes::Halfedge_handle find_halfedge(LineTes &, Point2, Point2);               /* :3955-3969: exact end points */
bool is_a_T_vertex(LineTes::Vertex_handle, bool verbose = false);                /* :4002-4046 */
bool is_an_X_vertex(LineTes::Vertex_handle);                                     /* :4058-4067 */
bool has_an_I_vertex_neighbour(LineTes::Vertex_handle);                          /* :4073-4081 */
bool is_an_irreducible_X_vertex(LineTes::Vertex_handle);                         /* :4087-4091 */
unsigned long int number_of_internal_vertices(TTessel &);                        /* :4105-4115 */
HPolygon face2poly(TTessel::Face_handle, bool simplify = true);                  /* :4314-4336 */
double sum_of_faces_squared_areas(TTessel *);                                    /* :4372-4382 */
double sum_of_min_angles(TTessel *);                                             /* :4387-4397 */
double sum_of_angles_obt(TTessel *);                                             /* :4402-4412 */
double sum_of_segment_squared_sizes(TTessel *);                                  /* :4440-4445 */
Polygons boundaries(HPolygon);                                                   /* :4453-4469 */
Polygon simplify(Polygon);                                                       /* :4474-4492 */
HPolygon simplify(HPolygon);                                                     /* :4495-4506 */
bool has_holes(LineTes::Face_handle &);                                          /* :4972-4974 */
bool has_holes(HPolygon &);                                                      /* :4980-4982 */
/* EXTENSION (not in the reference): length of internal edges, the evident intent
 * of edge_length's doc comment */
double edge_length_internal(Segment, TTessel *);

#endif /* LITE_B200_TTESSEL_H */
