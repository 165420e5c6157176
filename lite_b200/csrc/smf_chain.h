// smf_chain.h — one SMF (split / merge / flip) Metropolis-Hastings chain on a
// fixed-capacity T-tessellation, written once for the device (one thread per
// chain, state in HBM) and for the host-side unit test that checks the same
// source against the exact oracle without a GPU (tests/cpp/smf_host.cpp).  The
// library itself only runs it on the GPU (smf_ensemble.cuh).
//
// Reference behaviour restated (paths relative to /root/reference):
//   SMFChain::step / Hasting_ratio x3         lib/ttessel.cpp:3143-3311
//   TTessel::propose_split/merge/flip         :1755-1777
//   TTessel::update(Split/Merge/Flip)         :1580-1701
//   LineTes::split_edge/merge_edge/add_edge/remove_edge, suppress_edge  :606-780, :201-236
//   Split / Flip constructors, ray_exit_face  :2090-2132, :2432-2454, :4515-4576
//   Split/Merge/Flip::modified_elements       :2135-2232, :2302-2404, :2459-2713
//   Energy::statistic_variation / variation   :2914-3012, features :4124-4434
//
// Layout of one chain (a contiguous block of chain_bytes(cap) bytes, cap = capacity
// in segments, the four window sides included).  Indices are 16 bit and a halfedge
// record carries its own source point: 32 bytes = one memory sector, an edge (both
// twins, hence both end points and all links) = one aligned 64-byte line.  There is
// no vertex array: walking a face costs one dependent miss per halfedge instead of
// two (record, then vertex), and the chain is bound by exactly those misses
// (profiles/r1_k_smf_run.txt).  A vertex of degree 3 is stored three times; vertices
// never move, so the copies are bit-identical.
//   Header | he HE[6cap] | w double[3cap] | bsum double[3cap/32]
//          | seg SegRec[cap] | f_he u16[cap+1] | nb_list u16[cap] | b_list u16[cap]
// Halfedges are allocated in pairs (twin(h) = h^1); face 0 is the outside face;
// segments 0..3 are the window sides.  w[p] is the sampling weight of edge pair p
// (length x number of its halfedges that bound an inside face) and bsum the sums of
// 32 consecutive weights: length-weighted halfedge sampling reads ~n/32 + 32
// doubles instead of the reference's scan of every halfedge (:1956-1964).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/lite_b200.h"
#include "philox.h"

// The heavy steps of a proposal stay out of line on the device: inlined into one
// kernel body they need 255 registers and still spill; as calls the kernel fits 64-128
// registers, which is what decides how many chains a SM keeps in flight (the kernel is
// bound by the latency of dependent loads, not by issue slots: profiles/r1_k_smf_run.txt).
#ifdef __CUDACC__
#define SMF_FN __host__ __device__ __noinline__
#else
#define SMF_FN inline
#endif

namespace smf {

typedef uint16_t idx;
static const int NIL = 0xFFFF;
static const int WBLOCK = 32;
static const double kPi = 3.14159265358979323846;
static const double kHalfPi = kPi / 2;

struct alignas(32) HE {
  double x, y;  // source point
  idx next, prev, face, segd /* seg | dir << 15; 0xFFFF = dead pair */, nhf, phf;
  uint32_t pad;
};
struct alignas(16) SegRec {
  idx he, nedges, pos, kind;  // kind: 0 none (window side), 1 non-blocking, 2 blocking, NIL dead
  idx start, end;             // first and last halfedge along the segment's direction (dir == true):
  idx pad[2];                 // Seg::halfedges_start()/end() without walking the segment
};

enum Status { ST_OK = 0, ST_CAPACITY = 1, ST_BROKEN = 2 };

struct alignas(16) Header {
  int32_t cap;
  int32_t n_p, n_f, n_s;              // high-water marks of the id spaces (edge pairs, faces, segments)
  int32_t free_p, free_f, free_s;     // heads of the intrusive free lists, -1 = empty
  int32_t n_nb, n_b;                       // sizes of the non-blocking / blocking lists
  int32_t n_seg;                           // live segments, window sides included
  int32_t status;
  int32_t n_dirty;                         // weight blocks whose sum is stale (empty between proposals)
  int32_t pad0[2];
  double int_length;                       // TTessel::int_length (include/ttessel.h:786)
  double perimeter;                        // of the window
  double energy;                           // Energy::value, updated on acceptance (:3284)
  uint64_t step;                           // proposals made so far = Philox counter
  int64_t counts[6];                       // ModifCounts since the last reset
  int16_t dirty[8];                        // their ids
};

struct Energy {
  int d;
  int fid[LITE_MAX_D];
  uint32_t mask;
  double theta[LITE_MAX_D];
};

struct Pt {
  double x, y;
};

struct Chain {
  Header *H;   // working copy (shared memory in the kernel); load()/store() move it from/to the block
  Header *Hm;  // its home at the start of the chain's block
  HE *he;
  double *w, *bsum;
  SegRec *seg;
  idx *f_he, *nb_list, *b_list;
};

LITE_HD inline size_t align_up(size_t b, size_t a) { return (b + a - 1) / a * a; }
LITE_HD inline int n_blocks_cap(int cap) { return (3 * cap + WBLOCK - 1) / WBLOCK; }
LITE_HD inline size_t chain_bytes(int cap) {
  size_t b = align_up(sizeof(Header), 128);
  b += (size_t)6 * cap * sizeof(HE);
  b += (size_t)n_blocks_cap(cap) * WBLOCK * sizeof(double);
  b += align_up((size_t)n_blocks_cap(cap) * sizeof(double), 16);
  b += (size_t)cap * sizeof(SegRec);
  b += align_up((size_t)(cap + 1) * sizeof(idx), 16);
  b += align_up((size_t)cap * sizeof(idx), 16) * 2;
  return align_up(b, 256);
}
LITE_HD inline void bind(Chain &c, char *base, int cap) {
  char *p = base;
  c.Hm = (Header *)p; p += align_up(sizeof(Header), 128);
  c.he = (HE *)p; p += (size_t)6 * cap * sizeof(HE);
  c.w = (double *)p; p += (size_t)n_blocks_cap(cap) * WBLOCK * sizeof(double);
  c.bsum = (double *)p; p += align_up((size_t)n_blocks_cap(cap) * sizeof(double), 16);
  c.seg = (SegRec *)p; p += (size_t)cap * sizeof(SegRec);
  c.f_he = (idx *)p; p += align_up((size_t)(cap + 1) * sizeof(idx), 16);
  c.nb_list = (idx *)p; p += align_up((size_t)cap * sizeof(idx), 16);
  c.b_list = (idx *)p;
}
LITE_HD inline void load(Chain &c) { *c.H = *c.Hm; }
LITE_HD inline void store(Chain &c) { *c.Hm = *c.H; }

// ---------------------------------------------------------------------------
// field access
// ---------------------------------------------------------------------------
LITE_HD inline int h_next(const Chain &c, int h) { return c.he[h].next; }
LITE_HD inline int h_prev(const Chain &c, int h) { return c.he[h].prev; }
LITE_HD inline int h_face(const Chain &c, int h) { return c.he[h].face; }
LITE_HD inline int h_seg(const Chain &c, int h) { return c.he[h].segd & 0x7FFF; }
LITE_HD inline bool h_dir(const Chain &c, int h) { return (c.he[h].segd >> 15) != 0; }
LITE_HD inline int h_nhf(const Chain &c, int h) { int v = c.he[h].nhf; return v == NIL ? -1 : v; }
LITE_HD inline int h_phf(const Chain &c, int h) { int v = c.he[h].phf; return v == NIL ? -1 : v; }
LITE_HD inline void set_segdir(Chain &c, int h, int s, bool dir) { c.he[h].segd = (idx)(s | (dir ? 0x8000 : 0)); }
LITE_HD inline void set_dir(Chain &c, int h, bool dir) { c.he[h].segd = (idx)((c.he[h].segd & 0x7FFF) | (dir ? 0x8000 : 0)); }
LITE_HD inline Pt P_src(const Chain &c, int h) { Pt p; p.x = c.he[h].x; p.y = c.he[h].y; return p; }
LITE_HD inline Pt P_tgt(const Chain &c, int h) { return P_src(c, h ^ 1); }
LITE_HD inline void set_src(Chain &c, int h, Pt p) { c.he[h].x = p.x; c.he[h].y = p.y; }
LITE_HD inline bool same_pt(Pt a, Pt b) { return a.x == b.x && a.y == b.y; }
LITE_HD inline bool pair_alive(const Chain &c, int p) { return c.he[2 * p].segd != 0xFFFF; }
LITE_HD inline bool seg_bnd(int s) { return s < 4; }
LITE_HD inline bool h_bnd(const Chain &c, int h) { return h_seg(c, h) < 4; }
LITE_HD inline double h_len(const Chain &c, int h) {
  double w = c.w[h >> 1];
  return h_bnd(c, h) ? w : 0.5 * w;
}
// the source of h is a polygon corner of face(h): the previous halfedge does not
// continue straight into h (face2poly + simplify, lib/ttessel.cpp:4314-4336)
LITE_HD inline bool corner_at_src(const Chain &c, int h) { return c.he[c.he[h].prev].nhf != (idx)h; }

LITE_HD inline double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
LITE_HD inline double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
LITE_HD inline double dist(Pt a, Pt b) {
  double x = b.x - a.x, y = b.y - a.y;
  return sqrt(x * x + y * y);
}
LITE_HD inline double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

// ---------------------------------------------------------------------------
// sampling weights
// ---------------------------------------------------------------------------
// A move changes up to five weights in as many calls.  Recomputing a block sum right away
// costs one memory round trip per call, one after the other; instead the block is marked,
// its two cache lines are requested (non-blocking), and all marked blocks are summed once
// at the end of the proposal, when their lines arrive together.  In this latency-bound
// kernel the number of DEPENDENT round trips is what counts.
LITE_HD inline void prefetch_line(const void *p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
LITE_HD inline void sum_block(Chain &c, int b) {
  const double *w = c.w + b * WBLOCK;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int k = 0; k < WBLOCK; k += 4) { s0 += w[k]; s1 += w[k + 1]; s2 += w[k + 2]; s3 += w[k + 3]; }
  c.bsum[b] = (s0 + s1) + (s2 + s3);
}
LITE_HD inline void flush_dirty(Chain &c) {
  for (int i = 0; i < c.H->n_dirty; i++) sum_block(c, c.H->dirty[i]);
  c.H->n_dirty = 0;
}
LITE_HD inline void set_pair_weight(Chain &c, int p, double wnew) {
  c.w[p] = wnew;
  const int b = p / WBLOCK;
  for (int i = 0; i < c.H->n_dirty; i++) if (c.H->dirty[i] == b) return;
  if (c.H->n_dirty == 8) flush_dirty(c);
  c.H->dirty[c.H->n_dirty++] = (int16_t)b;
  prefetch_line(c.w + b * WBLOCK);
  prefetch_line(c.w + b * WBLOCK + WBLOCK / 2);
}
LITE_HD inline void set_weight(Chain &c, int h, double len) {
  set_pair_weight(c, h >> 1, h_bnd(c, h) ? len : 2.0 * len);
}
// LineTes_halfedge::set_length, lib/ttessel.cpp:1459
LITE_HD inline void set_len(Chain &c, int h) { set_weight(c, h, dist(P_src(c, h), P_tgt(c, h))); }

// ---------------------------------------------------------------------------
// element management (ids of removed elements are recycled, last freed first)
// ---------------------------------------------------------------------------
LITE_HD inline int new_pair(Chain &c, Pt v1, Pt v2) {
  int p;
  if (c.H->free_p >= 0) { p = c.H->free_p; int nx = c.he[2 * p].next; c.H->free_p = (nx == NIL) ? -1 : nx; }
  else {
    if (c.H->n_p >= 3 * c.H->cap) { c.H->status = ST_CAPACITY; return -1; }
    p = c.H->n_p++;
  }
  HE r;
  r.next = r.prev = r.face = r.nhf = r.phf = (idx)NIL; r.pad = 0; r.segd = 0x7FFF;
  r.x = v1.x; r.y = v1.y; c.he[2 * p] = r;
  r.x = v2.x; r.y = v2.y; c.he[2 * p + 1] = r;
  return 2 * p;
}
LITE_HD inline void del_pair(Chain &c, int h) {
  const int p = h >> 1;
  set_pair_weight(c, p, 0.0);
  c.he[2 * p].segd = 0xFFFF; c.he[2 * p + 1].segd = 0xFFFF;
  c.he[2 * p].next = (idx)(c.H->free_p < 0 ? NIL : c.H->free_p);
  c.H->free_p = p;
}
LITE_HD inline int new_face(Chain &c) {
  int f;
  if (c.H->free_f >= 0) { f = c.H->free_f; int nx = c.f_he[f]; c.H->free_f = (nx == NIL) ? -1 : nx; }
  else {
    if (c.H->n_f >= c.H->cap + 1) { c.H->status = ST_CAPACITY; return -1; }
    f = c.H->n_f++;
  }
  c.f_he[f] = (idx)NIL;
  return f;
}
LITE_HD inline void del_face(Chain &c, int f) {
  c.f_he[f] = (idx)(c.H->free_f < 0 ? NIL : c.H->free_f);
  c.H->free_f = f;
}
LITE_HD inline int new_seg(Chain &c, int he) {
  int s;
  if (c.H->free_s >= 0) { s = c.H->free_s; int nx = c.seg[s].he; c.H->free_s = (nx == NIL) ? -1 : nx; }
  else {
    if (c.H->n_s >= c.H->cap) { c.H->status = ST_CAPACITY; return -1; }
    s = c.H->n_s++;
  }
  SegRec r; r.he = (idx)he; r.nedges = 1; r.pos = (idx)NIL; r.kind = 0;
  r.start = r.end = (idx)he; r.pad[0] = r.pad[1] = 0;  // callers pass the dir == true halfedge
  c.seg[s] = r;
  c.H->n_seg++;
  return s;
}
LITE_HD inline void del_seg(Chain &c, int s) {
  c.seg[s].kind = (idx)NIL;
  c.seg[s].he = (idx)(c.H->free_s < 0 ? NIL : c.H->free_s);
  c.H->free_s = s;
  c.H->n_seg--;
}

// blocking / non-blocking lists (include/ttessel.h:784-785).  The chain draws
// uniformly from them, so removal may reorder (swap with the last element).
LITE_HD inline void list_add(Chain &c, int s, int kind) {
  c.seg[s].kind = (idx)kind;
  if (kind == 1) { c.seg[s].pos = (idx)c.H->n_nb; c.nb_list[c.H->n_nb++] = (idx)s; }
  else { c.seg[s].pos = (idx)c.H->n_b; c.b_list[c.H->n_b++] = (idx)s; }
}
LITE_HD inline void list_remove(Chain &c, int s) {
  const int kind = c.seg[s].kind;
  if (kind != 1 && kind != 2) return;
  idx *L = (kind == 1) ? c.nb_list : c.b_list;
  int &n = (kind == 1) ? c.H->n_nb : c.H->n_b;
  const int p = c.seg[s].pos;
  const int last = L[n - 1];
  L[p] = (idx)last; c.seg[last].pos = (idx)p;
  n--;
  c.seg[s].kind = 0; c.seg[s].pos = (idx)NIL;
}

// LineTes::set_junction, lib/ttessel.cpp:3909-3930
LITE_HD inline void set_junction(Chain &c, int e1, int e2) {
  if (e1 >= 0) { c.he[e1].nhf = (idx)(e2 >= 0 ? e2 : NIL); c.he[e1 ^ 1].phf = (idx)(e2 >= 0 ? (e2 ^ 1) : NIL); }
  if (e2 >= 0) { c.he[e2].phf = (idx)(e1 >= 0 ? e1 : NIL); c.he[e2 ^ 1].nhf = (idx)(e1 >= 0 ? (e1 ^ 1) : NIL); }
}
// Seg::set_halfedge_handle, lib/ttessel.cpp:841-851: keep the halfedge with dir == true
LITE_HD inline void seg_set_handle(Chain &c, int s, int h) { c.seg[s].he = (idx)(h_dir(c, h) ? h : (h ^ 1)); }
LITE_HD inline int seg_start(const Chain &c, int s) { return c.seg[s].start; }
LITE_HD inline int seg_end(const Chain &c, int s) { return c.seg[s].end; }
// the same by walking the along-segment links (consistency check only)
LITE_HD inline int seg_start_walk(const Chain &c, int s) {
  int e = c.seg[s].he;
  for (int p; (p = h_phf(c, e)) >= 0;) e = p;
  return e;
}
LITE_HD inline int seg_end_walk(const Chain &c, int s) {
  int e = c.seg[s].he;
  for (int n; (n = h_nhf(c, e)) >= 0;) e = n;
  return e;
}
// the halfedge of the pair of h that runs along its segment's direction
LITE_HD inline int dir_true(const Chain &c, int h) { return h_dir(c, h) ? h : (h ^ 1); }

// ---------------------------------------------------------------------------
// window (LineTes::insert_window, lib/ttessel.cpp:60-120; rectangle)
// ---------------------------------------------------------------------------
LITE_HD inline void init_window(Chain &c, int cap, double x0, double y0, double x1, double y1) {
  Header &H = *c.H;
  H.cap = cap;
  H.n_p = H.n_f = H.n_s = 0;
  H.free_p = H.free_f = H.free_s = -1;
  H.n_nb = H.n_b = 0; H.n_seg = 0; H.status = ST_OK; H.n_dirty = 0; H.pad0[0] = H.pad0[1] = 0;
  for (int k = 0; k < 8; k++) H.dirty[k] = 0;
  H.int_length = 0; H.energy = 0; H.step = 0;
  for (int k = 0; k < 6; k++) H.counts[k] = 0;
  const int nb = n_blocks_cap(cap);
  for (int i = 0; i < nb * WBLOCK; i++) c.w[i] = 0.0;
  for (int i = 0; i < nb; i++) c.bsum[i] = 0.0;
  const double xs[4] = {x0, x1, x1, x0}, ys[4] = {y0, y0, y1, y1};
  Pt vs[4];
  int hs[4];
  for (int i = 0; i < 4; i++) { vs[i].x = xs[i]; vs[i].y = ys[i]; }
  for (int i = 0; i < 4; i++) hs[i] = new_pair(c, vs[i], vs[(i + 1) & 3]);  // pairs 0..3: the window sides
  const int fo = new_face(c), fi = new_face(c);  // 0 = outside, 1 = the window
  for (int i = 0; i < 4; i++) {
    const int h = hs[i], hn = hs[(i + 1) & 3], hp = hs[(i + 3) & 3];
    c.he[h].next = (idx)hn; c.he[h].prev = (idx)hp;
    c.he[h ^ 1].next = (idx)(hp ^ 1); c.he[h ^ 1].prev = (idx)(hn ^ 1);
    c.he[h].face = (idx)fi; c.he[h ^ 1].face = (idx)fo;
    const int s = new_seg(c, h);  // ids 0..3: the window sides
    set_segdir(c, h, s, true); set_segdir(c, h ^ 1, s, false);
    set_len(c, h);
  }
  c.f_he[fi] = (idx)hs[0]; c.f_he[fo] = (idx)(hs[0] ^ 1);
  flush_dirty(c);
  H.perimeter = 0;
  for (int i = 0; i < 4; i++) H.perimeter += h_len(c, hs[i]);
  H.int_length = H.perimeter;
}

// ---------------------------------------------------------------------------
// DCEL surgery
// ---------------------------------------------------------------------------
// LineTes::split_edge, lib/ttessel.cpp:606-655: e keeps its source, the new pair runs
// from the new vertex to the old target
SMF_FN int split_edge(Chain &c, int e, double px, double py) {
  const int e_nhf = h_nhf(c, e);
  const int s = h_seg(c, e);
  const bool e_dir = h_dir(c, e);
  const int t = e ^ 1;
  const Pt b = P_src(c, t);
  Pt v; v.x = px; v.y = py;
  const int e2 = new_pair(c, v, b), t2 = e2 ^ 1;
  const int en = h_next(c, e), tp = h_prev(c, t);
  c.he[e2].next = (idx)en; c.he[en].prev = (idx)e2; c.he[e].next = (idx)e2; c.he[e2].prev = (idx)e;
  c.he[t2].prev = (idx)tp; c.he[tp].next = (idx)t2; c.he[t2].next = (idx)t; c.he[t].prev = (idx)t2;
  set_src(c, t, v);
  c.he[e2].face = c.he[e].face; c.he[t2].face = c.he[t].face;
  set_junction(c, e, e2);
  set_junction(c, e2, e_nhf);
  set_segdir(c, e2, s, e_dir); set_segdir(c, t2, s, !e_dir);
  set_len(c, e); set_len(c, e2);
  c.seg[s].nedges++;
  seg_set_handle(c, s, e);
  // the new pair follows e: after it along the segment if e runs with the segment, before it otherwise
  if (e_dir) { if (c.seg[s].end == (idx)e) c.seg[s].end = (idx)e2; }
  else { if (c.seg[s].start == (idx)t) c.seg[s].start = (idx)t2; }
  return e;
}

// LineTes::merge_edge, lib/ttessel.cpp:658-698: e2 = next_hf(e1) disappears with the
// vertex between them; the length is the sum of the two lengths, as in the reference
LITE_HD inline int merge_edge(Chain &c, int e1, int e2) {
  const int hf_before = h_phf(c, e1), hf_after = h_nhf(c, e2);
  const double new_len = h_len(c, e1) + h_len(c, e2);
  const int s = h_seg(c, e1);
  const int t1 = e1 ^ 1, t2 = e2 ^ 1;
  const Pt b = P_tgt(c, e2);
  const int e2n = h_next(c, e2), t2p = h_prev(c, t2);
  c.he[e1].next = (idx)e2n; c.he[e2n].prev = (idx)e1;
  c.he[t1].prev = (idx)t2p; c.he[t2p].next = (idx)t1;
  set_src(c, t1, b);
  const int fe = h_face(c, e2), ft = h_face(c, t2);
  const idx rep_e = c.f_he[fe], rep_t = c.f_he[ft];   // both requested before either is tested
  if (rep_e == (idx)e2) c.f_he[fe] = (idx)e1;
  if (rep_t == (idx)t2) c.f_he[ft] = (idx)t1;
  del_pair(c, e2);
  set_junction(c, hf_before, e1);
  set_junction(c, e1, hf_after);
  set_weight(c, e1, new_len);
  c.seg[s].nedges--;
  seg_set_handle(c, s, e1);
  if (h_dir(c, e1)) { if (c.seg[s].end == (idx)e2) c.seg[s].end = (idx)e1; }
  else { if (c.seg[s].start == (idx)t2) c.seg[s].start = (idx)t1; }
  return e1;
}

// The halfedges first, next(first), ..., last take face id f.  Walked from both ends at once:
// the two pointer chases (next from the front, prev from the back) are independent, so their
// misses overlap and the dependent depth is half the length of the run.  Same stores as a walk
// from one end.
LITE_HD inline void relabel_cycle(Chain &c, int first, int last, int f) {
  int a = first, b = last;
  for (;;) {
    c.he[a].face = (idx)f;
    if (a == b) break;
    c.he[b].face = (idx)f;
    const int an = h_next(c, a), bp = h_prev(c, b);   // both requested before either is used
    if (an == b) break;
    a = an; b = bp;
  }
}

// LineTes::add_edge, lib/ttessel.cpp:700-740: new edge from tgt(prev1) to tgt(prev2)
// inside their common face; ext_prev is the halfedge whose straight continuation the
// new edge is (-1: it starts a new segment)
SMF_FN int add_edge(Chain &c, int prev1, int prev2, int ext_prev) {
  const Pt v1 = P_tgt(c, prev1), v2 = P_tgt(c, prev2);
  const int f = h_face(c, prev1);
  const int a1 = h_next(c, prev1), a2 = h_next(c, prev2);
  const int h = new_pair(c, v1, v2), t = h ^ 1;
  c.he[prev1].next = (idx)h; c.he[h].prev = (idx)prev1; c.he[h].next = (idx)a2; c.he[a2].prev = (idx)h;
  c.he[prev2].next = (idx)t; c.he[t].prev = (idx)prev2; c.he[t].next = (idx)a1; c.he[a1].prev = (idx)t;
  const int nf = new_face(c);
  c.f_he[f] = (idx)t; c.he[t].face = (idx)f;
  c.f_he[nf] = (idx)h;
  relabel_cycle(c, h, h_prev(c, h), nf);
  if (ext_prev >= 0) {
    set_junction(c, ext_prev, h);
    const int s = h_seg(c, ext_prev);
    const bool d = h_dir(c, ext_prev);
    set_segdir(c, h, s, d); set_segdir(c, t, s, !d);
    c.seg[s].nedges++;
    if (d) { if (c.seg[s].end == (idx)ext_prev) c.seg[s].end = (idx)h; }
    else { if (c.seg[s].start == (idx)(ext_prev ^ 1)) c.seg[s].start = (idx)t; }
  } else {
    const int s = new_seg(c, h);
    set_segdir(c, h, s, true); set_segdir(c, t, s, false);
  }
  set_len(c, h);
  return h;
}

// LineTes::remove_edge, lib/ttessel.cpp:743-780
LITE_HD inline void remove_edge(Chain &c, int e) {
  const int t = e ^ 1;
  const int f1 = h_face(c, e), f2 = h_face(c, t);
  const int p1 = h_prev(c, e), n1 = h_next(c, e), p2 = h_prev(c, t), n2 = h_next(c, t);
  const int phf = h_phf(c, e), nhf = h_nhf(c, e);
  set_junction(c, phf, -1);
  set_junction(c, -1, nhf);
  const int s = h_seg(c, e);
  if (phf < 0 && nhf < 0) {
    list_remove(c, s);
    del_seg(c, s);
  } else {
    c.seg[s].nedges--;
    seg_set_handle(c, s, phf >= 0 ? phf : nhf);
    {
      // e is the first or the last edge of its segment: the neighbour takes over that end
      const int d = dir_true(c, e);
      const int dn = h_dir(c, e) ? nhf : (phf >= 0 ? (phf ^ 1) : -1);   // after d along the segment
      const int dp = h_dir(c, e) ? phf : (nhf >= 0 ? (nhf ^ 1) : -1);   // before d
      if (c.seg[s].start == (idx)d && dn >= 0) c.seg[s].start = (idx)dn;
      if (c.seg[s].end == (idx)d && dp >= 0) c.seg[s].end = (idx)dp;
    }
  }
  c.he[p1].next = (idx)n2; c.he[n2].prev = (idx)p1; c.he[p2].next = (idx)n1; c.he[n1].prev = (idx)p2;
  del_pair(c, e);
  c.f_he[f1] = (idx)p1;
  if (f2 != f1) {
    // only the halfedges that bounded f2 change face: n2 .. p2 on the merged ccb
    relabel_cycle(c, n2, p2, f1);
    del_face(c, f2);
  }
}

// LineTes::suppress_edge, lib/ttessel.cpp:201-236: remove the edge, then dissolve its
// end vertices where they are left with two collinear edges
SMF_FN void suppress_edge(Chain &c, int e) {
  const int p1 = h_prev(c, e), p2 = h_prev(c, e ^ 1);  // arrive at src(e), tgt(e)
  remove_edge(c, e);
  prefetch_line(&c.he[h_next(c, p2)]);   // the second end's neighbour travels while the first end is dissolved
  const int ps[2] = {p1, p2};
  for (int k = 0; k < 2; k++) {
    const int p = ps[k];
    const int o = h_next(c, p);                       // leaves the vertex
    const bool degree2 = h_next(c, o ^ 1) == (p ^ 1);  // the only other edge at the vertex
    const int nh = h_nhf(c, p);
    if (degree2 && nh >= 0) merge_edge(c, p, nh);
  }
}

// ---------------------------------------------------------------------------
// moves
// ---------------------------------------------------------------------------
struct SplitMove {
  int e1, e2;
  double x, angle;
  Pt p1, p2;
  bool valid;
};
struct FlipMove {
  int e1, e2, e3;
  Pt p2;
  bool right, valid;
};

// ray_exit_face, lib/ttessel.cpp:4515-4576 on the ccb of a hole-free face: nearest
// forward hit of the line a*X+b*Y=w from p along (dx,dy); strict minimum, first wins
SMF_FN int ray_exit(const Chain &c, int start_he, double a, double b, double w, Pt p, double dx,
                            double dy, int skip0, int skip1, Pt &hit) {
  // The window of an ensemble is a rectangle, so every face is convex: the line meets its
  // boundary once besides the ray's own edge and the first forward hit is the nearest one
  // (:4563).  The ccb is walked from both ends at once -- `next` from start_he, `prev` from the
  // halfedge before it -- so that the two pointer chases overlap and the crossed edge is met
  // after half as many dependent loads.  The front applies the reference's rule as it stands
  // (first edge in ccb order whose end values s = a*X+b*Y-w straddle or touch zero, forward of
  // p, not at p).  The back only takes an edge whose end values are non-zero and of opposite
  // signs: on a convex face no edge before it can qualify then (the values change sign twice
  // around the face, once on the ray's own edge), so it is the edge the front would have found.
  // An edge with a zero end value met from the back is left to the front.
  int hf = start_he, hb = h_prev(c, start_he);
  Pt Sf = P_src(c, hf);
  double skf = fma(a, Sf.x, fma(b, Sf.y, -w));
  Pt Tb = Sf;          // the target of the last halfedge is the source of the first
  double snb = skf;
  for (;;) {
    const Pt T = P_tgt(c, hf);
    const Pt S = P_src(c, hb);              // both lines requested before either is used
    const int fn = h_next(c, hf), bp = h_prev(c, hb);
    const double sn = fma(a, T.x, fma(b, T.y, -w));
    if (hf != skip0 && hf != skip1) {
      const bool cr = (skf <= 0.0 && sn >= 0.0) || (skf >= 0.0 && sn <= 0.0);
      if (cr && skf != sn) {
        const double sg = skf / (skf - sn);
        const double X = fma(sg, T.x - Sf.x, Sf.x), Y = fma(sg, T.y - Sf.y, Sf.y);
        const double rx = X - p.x, ry = Y - p.y;
        if (rx * dx + ry * dy >= 0.0 && sqrt(rx * rx + ry * ry) != 0.0) { hit.x = X; hit.y = Y; return hf; }
      }
    }
    if (hf == hb) return -1;
    const double sk = fma(a, S.x, fma(b, S.y, -w));
    if (hb != skip0 && hb != skip1 && ((sk < 0.0 && snb > 0.0) || (sk > 0.0 && snb < 0.0))) {
      const double sg = sk / (sk - snb);
      const double X = fma(sg, Tb.x - S.x, S.x), Y = fma(sg, Tb.y - S.y, S.y);
      const double rx = X - p.x, ry = Y - p.y;
      if (rx * dx + ry * dy >= 0.0 && sqrt(rx * rx + ry * ry) != 0.0) { hit.x = X; hit.y = Y; return hb; }
    }
    if (fn == hb) return -1;
    Sf = T; skf = sn; hf = fn;
    Tb = S; snb = sk; hb = bp;
  }
}

// What ray_exit needs: the supporting line, the ray, where to start the walk and the two
// halfedges whose hit is the ray's own source.  Split and flip proposals build one and
// share a single ray_exit call site in smf_step (lanes of a warp working on either kind of
// move run the face walk together instead of one kind after the other).
struct Ray {
  double a, b, w, dx, dy;
  Pt p;
  int start, skip0, skip1;
  bool ok;
};
// Split::Split(e,x,angle), lib/ttessel.cpp:2090-2132, up to the ray
// The line's normal (sin l, -cos l), l = atan2(e1) + angle (:2094-2100), comes from the
// addition formulas with e1's unit direction and (cos, sin) of the proposal angle: for a
// drawn angle cos = 1 - 2U and sin = 2 sqrt(U(1-U)), so no atan2 / sin / cos / acos is
// evaluated on the way (the angle itself is only needed by the trace).
LITE_HD inline void prepare_split(const Chain &c, int e, double x, double angle, double cos_angle, double sin_angle,
                                  SplitMove &m, Ray &r) {
  m.e1 = e; m.e2 = -1; m.x = x; m.angle = angle; m.valid = false;
  m.p1.x = m.p1.y = m.p2.x = m.p2.y = 0;
  r.ok = false;
  const Pt P = P_src(c, e), Q = P_tgt(c, e);
  const double ex = Q.x - P.x, ey = Q.y - P.y;
  const double inv = 1.0 / sqrt(ex * ex + ey * ey);
  const double ux = ex * inv, uy = ey * inv;
  const double a = uy * cos_angle + ux * sin_angle, b = uy * sin_angle - ux * cos_angle;
  const double u = dadd(dmul(a, P.x), dmul(b, P.y)), v = dadd(dmul(a, Q.x), dmul(b, Q.y));
  const double w = dadd(u, dmul(x, dadd(v, -u)));
  const double sP = fma(a, P.x, fma(b, P.y, -w)), sQ = fma(a, Q.x, fma(b, Q.y, -w));
  if (!(sP != sQ)) return;
  const double t1 = sP / (sP - sQ);
  m.p1.x = fma(t1, Q.x - P.x, P.x); m.p1.y = fma(t1, Q.y - P.y, P.y);
  r.a = a; r.b = b; r.w = w; r.p = m.p1;
  if (sQ < 0.0) { r.dx = b; r.dy = -a; } else { r.dx = -b; r.dy = a; }
  r.start = e; r.skip0 = e; r.skip1 = (x == 0.0) ? h_prev(c, e) : -1;
  r.ok = true;
}
// Flip::Flip(e1), lib/ttessel.cpp:2432-2454, up to the ray: a = src(e1) is where another
// segment (the one of e3) ends on e1's segment; the flip removes e1 and prolongs e3 beyond a
LITE_HD inline void prepare_flip(const Chain &c, int e1, FlipMove &m, Ray &r) {
  m.e1 = e1; m.e2 = -1; m.valid = false; m.p2.x = m.p2.y = 0;
  const int hp = h_prev(c, e1);
  const bool right = (hp != h_phf(c, e1));
  const int e3 = right ? hp : (h_next(c, e1 ^ 1) ^ 1);
  m.e3 = e3; m.right = right;
  const Pt A = P_src(c, e1), S3 = P_src(c, e3);
  double dx = A.x - S3.x, dy = A.y - S3.y;
  const double nrm = sqrt(dx * dx + dy * dy);
  dx /= nrm; dy /= nrm;
  r.a = dy; r.b = -dx; r.w = r.a * A.x + r.b * A.y; r.p = A; r.dx = dx; r.dy = dy;
  if (right) { r.skip0 = e1 ^ 1; r.skip1 = h_next(c, e1 ^ 1); r.start = r.skip0; }
  else { r.skip0 = e1; r.skip1 = h_prev(c, e1); r.start = e1; }
  r.ok = true;
}

// TTessel::length_weighted_random_halfedge (lib/ttessel.cpp:1948-1966) by two-level
// sums: u01 in [0,1) -> a halfedge bounding an inside face, with probability
// proportional to its length
SMF_FN int sample_halfedge(const Chain &c, double u01) {
  // First pass: the total, with every load independent of the others, so the misses
  // of the whole array overlap; the early-exit scan below then runs out of L1.  (Taking
  // int_length as the total and scanning once exposes one miss after the other: measured
  // 8% slower, gpurun_out/smf_ab2.log.)
  const int nb = (c.H->n_p + WBLOCK - 1) / WBLOCK;
  double total = 0;
  for (int b = 0; b < nb; b++) total += c.bsum[b];
  const double target = u01 * total;
  double cum = 0;
  int b = 0, last_b = 0;
  for (; b < nb; b++) {
    const double s = c.bsum[b];
    if (s > 0) last_b = b;
    if (cum + s > target) break;
    cum += s;
  }
  if (b >= nb) { b = last_b; cum = target; }  // rounding at the very end of the range
  int p = -1, last_p = -1;
  double r = target - cum;
  const double *w = c.w + b * WBLOCK;
  double acc = 0;
  for (int k = 0; k < WBLOCK; k++) {
    const double wk = w[k];
    if (wk > 0) last_p = k;
    if (acc + wk > r) { p = k; r -= acc; break; }
    acc += wk;
  }
  if (p < 0) { p = last_p; r = 0; }
  p += b * WBLOCK;
  if (p < b * WBLOCK) {
    // the block is empty and its sum is rounding residue: take the first live pair
    for (p = 0; p < c.H->n_p && !(c.w[p] > 0); p++) {}
    r = 0;
  }
  const int h = 2 * p;
  if (h_bnd(c, h)) return (h_face(c, h) != 0) ? h : (h ^ 1);
  return (r >= 0.5 * c.w[p]) ? (h ^ 1) : h;
}

// ---------------------------------------------------------------------------
// face features from a streamed polygon: vertices in order with their corner
// flag (non-corners are the T-vertices simplify() removes, :4474-4506)
// ---------------------------------------------------------------------------
LITE_HD inline double angle_between(double ax, double ay, double bx, double by) {  // :4255-4268
  const double n1 = ax * ax + ay * ay, n2 = bx * bx + by * by;
  double cs = (ax * bx + ay * by) / sqrt(n1 * n2);
  cs = fmin(1.0, fmax(-1.0, cs));
  return acos(cs);
}
struct FaceFeat {
  double area, perim, sumang, minang;
};
struct Poly {
  Pt p0, p1, a, b;  // first two vertices; vertex i-2, vertex i-1
  bool c0, c1, bc;
  int n;
  bool need_ang;
  double area2, perim, sumang, minang, seed;
};
LITE_HD inline void poly_begin(Poly &P, bool need_ang) {
  P.n = 0; P.need_ang = need_ang; P.area2 = 0; P.perim = 0; P.sumang = 0;
  P.minang = 1.7976931348623157e308; P.seed = -1.0;
  P.c0 = P.c1 = P.bc = false;
}
LITE_HD inline void poly_corner(Poly &P, Pt prev, Pt cur, Pt next, bool first_vertex) {
  const double a = angle_between(prev.x - cur.x, prev.y - cur.y, next.x - cur.x, next.y - cur.y);
  // min_angle is seeded with the turning angle at the first corner (:4354)
  if (first_vertex || P.seed < 0) P.seed = kPi - a;
  const double q = kHalfPi - a;
  if (q >= 0) P.sumang += q;
  P.minang = fmin(P.minang, a);
}
LITE_HD inline void poly_push(Poly &P, Pt v, bool corner) {
  if (P.n == 0) { P.p0 = v; P.c0 = corner; }
  else {
    if (P.n == 1) { P.p1 = v; P.c1 = corner; }
    else {
      if (P.need_ang && P.bc) poly_corner(P, P.a, P.b, v, false);
      P.area2 += cross2(P.b.x - P.p0.x, P.b.y - P.p0.y, v.x - P.p0.x, v.y - P.p0.y);
    }
    P.perim += dist(P.b, v);
    P.a = P.b;
  }
  P.b = v; P.bc = corner;
  P.n++;
}
LITE_HD inline FaceFeat poly_end(Poly &P) {
  FaceFeat F;
  P.perim += dist(P.b, P.p0);
  if (P.need_ang) {
    if (P.bc) poly_corner(P, P.a, P.b, P.p0, false);
    if (P.c0) poly_corner(P, P.b, P.p0, P.p1, true);
  }
  F.area = 0.5 * P.area2; F.perim = P.perim; F.sumang = P.sumang;
  F.minang = (P.seed >= 0) ? fmin(P.seed, P.minang) : 0.0;
  return F;
}
// push src(h) for h = from, next(from), ... up to `to` (to_inclusive) or just before it
LITE_HD inline void poly_walk(const Chain &c, Poly &P, int from, int to, bool to_inclusive) {
  int h = from;
  if (!to_inclusive && h == to) return;
  for (;;) {
    poly_push(P, P_src(c, h), corner_at_src(c, h));
    if (to_inclusive && h == to) break;
    h = h_next(c, h);
    if (!to_inclusive && h == to) break;
  }
}
// an unmodified face in face2poly order: the first vertex is the target of outer_ccb()
SMF_FN FaceFeat face_feat(const Chain &c, int f, bool need_ang) {
  Poly P; poly_begin(P, need_ang);
  const int h0 = h_next(c, c.f_he[f]);
  int h = h0;
  do { poly_push(P, P_src(c, h), corner_at_src(c, h)); h = h_next(c, h); } while (h != h0);
  return poly_end(P);
}
LITE_HD inline double shape_of(double perim, double area) {  // face_shape, :4247-4249
  return perim / (4.0 * sqrt(fabs(area)));
}

#define SMF_BIT(f) (1u << (f))
static const uint32_t kMaskFaceAng = SMF_BIT(LITE_FEAT_FACE_SUM_OF_ANGLES) | SMF_BIT(LITE_FEAT_MIN_ANGLE);
static const uint32_t kMaskFaceWalk = SMF_BIT(LITE_FEAT_FACE_AREA_2) | SMF_BIT(LITE_FEAT_FACE_SHAPE) | kMaskFaceAng;

LITE_HD inline void add_face_feats(const Energy &G, double *acc, double sign, const FaceFeat &F) {
  for (int i = 0; i < G.d; i++) {
    switch (G.fid[i]) {
      case LITE_FEAT_FACE_AREA_2: acc[i] += sign * (F.area * F.area); break;
      case LITE_FEAT_FACE_SHAPE: acc[i] += sign * shape_of(F.perim, F.area); break;
      case LITE_FEAT_FACE_SUM_OF_ANGLES: acc[i] += sign * F.sumang; break;
      case LITE_FEAT_MIN_ANGLE: acc[i] += sign * F.minang; break;
      default: break;
    }
  }
}
// |a-m| + |m-b| - |a-b|: what a cut edge contributes to an edge-length sum
LITE_HD inline double edge_cut_delta(Pt a, Pt m, Pt b) { return dist(a, m) + dist(m, b) - dist(a, b); }

// ---------------------------------------------------------------------------
// statistic variations  phi(after) - phi(before), one value per registered feature
// ---------------------------------------------------------------------------
// Split::modified_elements, lib/ttessel.cpp:2135-2232
SMF_FN void split_variation(const Chain &c, const Energy &G, const SplitMove &m, double *dv) {
  const int e1 = m.e1, e2 = m.e2;
  const bool bnd1 = h_bnd(c, e1), bnd2 = h_bnd(c, e2);
  const Pt P = P_src(c, e1), Q = P_tgt(c, e1), S2 = P_src(c, e2), E2 = P_tgt(c, e2);
  const double d01 = dist(m.p1, m.p2);
  double acc[LITE_MAX_D];
  for (int i = 0; i < G.d; i++) acc[i] = 0.0;
  if (G.mask & kMaskFaceWalk) {
    const bool ang = (G.mask & kMaskFaceAng) != 0;
    // children (polygon_insert_edge, :4600-4631):
    //   C1 = [p1, p2, tgt(e2) .. src(e1)],  C2 = [p2, p1, tgt(e1) .. src(e2)]
    Poly Pl;
    poly_begin(Pl, ang);
    poly_push(Pl, m.p1, true); poly_push(Pl, m.p2, true);
    poly_walk(c, Pl, h_next(c, e2), e1, true);
    const FaceFeat C1 = poly_end(Pl);
    poly_begin(Pl, ang);
    poly_push(Pl, m.p2, true); poly_push(Pl, m.p1, true);
    poly_walk(c, Pl, h_next(c, e1), e2, true);
    const FaceFeat C2 = poly_end(Pl);
    const FaceFeat F0 = face_feat(c, h_face(c, e1), ang);
    for (int i = 0; i < G.d; i++) {
      switch (G.fid[i]) {
        case LITE_FEAT_FACE_AREA_2: acc[i] = -2.0 * C1.area * C2.area; break;  // a1^2+a2^2-(a1+a2)^2
        case LITE_FEAT_FACE_SHAPE: acc[i] = shape_of(C1.perim, C1.area) + shape_of(C2.perim, C2.area) - shape_of(F0.perim, F0.area); break;
        case LITE_FEAT_FACE_SUM_OF_ANGLES: acc[i] = C1.sumang + C2.sumang - F0.sumang; break;
        case LITE_FEAT_MIN_ANGLE: acc[i] = C1.minang + C2.minang - F0.minang; break;
        default: break;
      }
    }
  }
  for (int i = 0; i < G.d; i++) {
    double v = 0.0;
    switch (G.fid[i]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW: {
        bool in1 = !bnd1;
        if (m.x == 0.0) in1 = !bnd1 && !h_bnd(c, h_prev(c, e1));
        v = (in1 ? 1.0 : 0.0) + (bnd2 ? 0.0 : 1.0);
      } break;
      case LITE_FEAT_EDGE_LENGTH:
        // as written only window-boundary edges count (:4141-4149): rounding noise
        if (bnd1) v += edge_cut_delta(P, m.p1, Q);
        if (bnd2) v += edge_cut_delta(S2, m.p2, E2);
        break;
      case LITE_FEAT_EDGE_LENGTH_INTERNAL:
        if (!bnd1) v += edge_cut_delta(P, m.p1, Q);
        if (!bnd2) v += edge_cut_delta(S2, m.p2, E2);
        v += d01;
        break;
      case LITE_FEAT_SEG_NUMBER: v = 1.0; break;
      case LITE_FEAT_IS_SEGMENT_INTERNAL: v = 1.0; break;
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: v = -1.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2:
        if (bnd1) v += 2.0 * c.seg[h_seg(c, e1)].nedges + 1.0;
        if (bnd2) v += 2.0 * c.seg[h_seg(c, e2)].nedges + 1.0;
        break;
      case LITE_FEAT_FACE_NUMBER: v = 1.0; break;
      case LITE_FEAT_FACE_PERIMETER: v = 2.0 * d01; break;
      default: v = acc[i]; break;
    }
    dv[i] = v;
  }
}

// Merge::modified_elements, lib/ttessel.cpp:2302-2404; e = the only edge of a
// non-blocking segment
SMF_FN void merge_variation(const Chain &c, const Energy &G, int e, double *dv) {
  const int t = e ^ 1;
  const int e2h = h_next(c, e), e1h = h_next(c, t);  // Merge::e2(), Merge::e1()
  const bool bnd_e2 = h_bnd(c, e2h), bnd_e1 = h_bnd(c, e1h);
  const Pt p1 = P_src(c, e), p2 = P_src(c, t);
  const Pt X2 = P_tgt(c, e2h), X1 = P_tgt(c, e1h);
  const Pt Y2 = P_src(c, h_prev(c, t)), Y1 = P_src(c, h_prev(c, e));
  const double elen = dist(p1, p2);
  double acc[LITE_MAX_D];
  for (int i = 0; i < G.d; i++) acc[i] = 0.0;
  if (G.mask & kMaskFaceWalk) {
    const bool ang = (G.mask & kMaskFaceAng) != 0;
    add_face_feats(G, acc, -1.0, face_feat(c, h_face(c, e), ang));
    add_face_feats(G, acc, -1.0, face_feat(c, h_face(c, t), ang));
    // polygon_remove_edge (:4945-4966): face(e) after p2 up to p1 (flat), then
    // face(twin e) after p1 up to p2 (flat)
    Poly Pl; poly_begin(Pl, ang);
    poly_walk(c, Pl, h_next(c, e2h), e, false);
    poly_push(Pl, p1, false);
    poly_walk(c, Pl, h_next(c, e1h), t, false);
    poly_push(Pl, p2, false);
    add_face_feats(G, acc, +1.0, poly_end(Pl));
  }
  for (int i = 0; i < G.d; i++) {
    double v = 0.0;
    switch (G.fid[i]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW: v = -((bnd_e1 ? 0.0 : 1.0) + (bnd_e2 ? 0.0 : 1.0)); break;
      case LITE_FEAT_EDGE_LENGTH:
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: {
        const bool wantb = (G.fid[i] == LITE_FEAT_EDGE_LENGTH);
        if (!wantb) v -= elen;
        if (bnd_e2 == wantb) v += dist(Y2, X2) - dist(p2, X2) - dist(Y2, p2);
        if (bnd_e1 == wantb) v += dist(Y1, X1) - dist(p1, X1) - dist(Y1, p1);
      } break;
      case LITE_FEAT_SEG_NUMBER: v = -1.0; break;
      case LITE_FEAT_IS_SEGMENT_INTERNAL: v = -1.0; break;
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: v = 1.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2:
        if (bnd_e1) v -= 2.0 * c.seg[h_seg(c, e1h)].nedges - 1.0;
        if (bnd_e2) v -= 2.0 * c.seg[h_seg(c, e2h)].nedges - 1.0;
        break;
      case LITE_FEAT_FACE_NUMBER: v = -1.0; break;
      case LITE_FEAT_FACE_PERIMETER: v = -2.0 * elen; break;
      default: v = acc[i]; break;
    }
    dv[i] = v;
  }
}

// Flip::modified_elements, lib/ttessel.cpp:2459-2713
SMF_FN void flip_variation(const Chain &c, const Energy &G, const FlipMove &m, double *dv) {
  const int e1 = m.e1, t1 = e1 ^ 1, e2 = m.e2;
  const bool right = m.right;
  const Pt a = P_src(c, e1), q = P_src(c, t1), p2 = m.p2;
  const int n1 = h_next(c, e1);  // leaves q along the incident segment
  const bool bnd_n1 = h_bnd(c, n1), bnd_e2 = h_bnd(c, e2);
  const Pt X = P_tgt(c, n1), Y = P_src(c, h_prev(c, t1));
  const Pt S2 = P_src(c, e2), E2 = P_tgt(c, e2);
  const bool tgt_is_q = same_pt(E2, q), src_is_q = same_pt(S2, q);  // copies of a vertex are bit-identical
  const double l_e1 = dist(a, q), l_new = dist(a, p2);
  double acc[LITE_MAX_D];
  for (int i = 0; i < G.d; i++) acc[i] = 0.0;
  if (G.mask & kMaskFaceWalk) {
    const bool ang = (G.mask & kMaskFaceAng) != 0;
    const int fL = h_face(c, e1), fT = h_face(c, t1);
    Poly Pl;
    if (right) {
      // the ray enters face(twin e1); deleted faces: the face of e2 first (:2644-2651)
      add_face_feats(G, acc, -1.0, face_feat(c, fT, ang));
      add_face_feats(G, acc, -1.0, face_feat(c, fL, ang));
      // [p2, a, after a in face(twin e1) .. src(e2)]
      poly_begin(Pl, ang);
      poly_push(Pl, p2, true); poly_push(Pl, a, true);
      poly_walk(c, Pl, h_next(c, h_next(c, t1)), e2, true);
      add_face_feats(G, acc, +1.0, poly_end(Pl));
      // face(e1) after q .. before a, a (flat), p2, tgt(e2) .. before q, q (flat)
      poly_begin(Pl, ang);
      poly_walk(c, Pl, h_next(c, n1), e1, false);
      poly_push(Pl, a, false); poly_push(Pl, p2, true);
      poly_walk(c, Pl, h_next(c, e2), t1, false);
      poly_push(Pl, q, false);
      add_face_feats(G, acc, +1.0, poly_end(Pl));
    } else {
      add_face_feats(G, acc, -1.0, face_feat(c, fL, ang));
      add_face_feats(G, acc, -1.0, face_feat(c, fT, ang));
      // [a, p2, tgt(e2) .. before a]
      poly_begin(Pl, ang);
      poly_push(Pl, a, true); poly_push(Pl, p2, true);
      poly_walk(c, Pl, h_next(c, e2), e1, false);
      add_face_feats(G, acc, +1.0, poly_end(Pl));
      // face(e1) after q .. src(e2), p2, a (flat), face(twin e1) after a .. before q, q (flat)
      poly_begin(Pl, ang);
      if (e2 != n1) poly_walk(c, Pl, h_next(c, n1), e2, true);  // e2 == n1: the ray ends on the edge leaving q
      poly_push(Pl, p2, true); poly_push(Pl, a, false);
      poly_walk(c, Pl, h_next(c, h_next(c, t1)), t1, false);
      poly_push(Pl, q, false);
      add_face_feats(G, acc, +1.0, poly_end(Pl));
    }
  }
  for (int i = 0; i < G.d; i++) {
    double v = 0.0;
    switch (G.fid[i]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW: v = (bnd_e2 ? 0.0 : 1.0) - (bnd_n1 ? 0.0 : 1.0); break;
      case LITE_FEAT_EDGE_LENGTH:
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: {
        const bool wantb = (G.fid[i] == LITE_FEAT_EDGE_LENGTH);
        if (!wantb) v += l_new - l_e1;
        if (bnd_n1 == wantb) v -= dist(q, X) + dist(q, Y);  // :2476-2486
        if (!tgt_is_q && !src_is_q) {
          if (bnd_e2 == wantb) v += dist(p2, E2) + dist(p2, S2) - dist(E2, S2);
          if (bnd_n1 == wantb) v += dist(X, Y);
        } else if (tgt_is_q) {
          if (bnd_e2 == wantb) v += dist(p2, S2) + dist(p2, X);
        } else {
          if (bnd_e2 == wantb) v += dist(p2, E2) + dist(p2, Y);
        }
      } break;
      case LITE_FEAT_SEG_NUMBER:
      case LITE_FEAT_IS_SEGMENT_INTERNAL:
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: v = 0.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2: {
        const int sg2 = h_seg(c, e2), sgn = h_seg(c, n1);
        if (sg2 != sgn) {  // else one segment, cut and rejoined at once (:2696-2702)
          if (bnd_e2) v += 2.0 * c.seg[sg2].nedges + 1.0;
          if (bnd_n1) v -= 2.0 * c.seg[sgn].nedges - 1.0;
        }
      } break;
      case LITE_FEAT_FACE_NUMBER: v = 0.0; break;
      case LITE_FEAT_FACE_PERIMETER: v = 2.0 * (l_new - l_e1); break;
      default: v = acc[i]; break;
    }
    dv[i] = v;
  }
}

LITE_HD inline double dot_theta(const Energy &G, const double *dv) {  // Energy::variation, :2994-3012
  double e = 0;
  for (int i = 0; i < G.d; i++) e += dv[i] * G.theta[i];
  return e;
}

// ---------------------------------------------------------------------------
// one proposal  (SMFChain::step body, lib/ttessel.cpp:3266-3307)
// ---------------------------------------------------------------------------
// What a step did, for the parity tests (device -> host trace, replayed by the oracle).
struct TraceRec {
  double type;          // 0 split, 1 merge, 2 flip
  double status;        // 0 no candidate (empty list), 1 evaluated + rejected, 2 accepted, -1 invalid geometry
  double r, e_var, x_accept;
  double e1s[2], e1t[2];  // endpoints of e1 (split, flip) or of the merged edge
  double x, angle;        // split parameters
  double p1[2], p2[2];
  double e2s[2], e2t[2];
  double n_seg, n_nb, n_b, int_length;  // after the step
};

struct StepRng {
  uint64_t seed, chain;  // Philox key and the chain's global id
};
// draw k (0,1,2) of proposal t of a chain: 4 words
LITE_HD inline void draw4(const StepRng &R, uint64_t t, uint32_t k, uint32_t out[4]) {
  // counter = (t, chain-low | k): chains and proposals never collide
  uint32_t c0 = (uint32_t)t, c1 = (uint32_t)(t >> 32), c2 = (uint32_t)R.chain, c3 = (uint32_t)(R.chain >> 32) ^ (k << 28);
  uint32_t k0 = (uint32_t)R.seed, k1 = (uint32_t)(R.seed >> 32);
  for (int r = 0; r < 10; r++) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct ChainParams {
  double p_split, p_merge;
  // bit 0: request the lines around e1 as soon as it is known, bit 1: the heads of the free
  // lists and the tails of the segment lists at the start of a proposal, bit 2: the lines around
  // e2 when the ray has found it.  Requests only (prefetch_line): the trajectory is unchanged.
  int prefetch = 0;
};

// The lines an accepted move reads and writes around halfedge e: its four neighbours on the two
// faces, its neighbours along the segment, the segment record, its weight.  Their indices sit in
// e's own 64-byte line, so the requests leave together, one round trip before the DCEL
// primitives (or, for e1, the ray walk) would ask for them one after the other.
LITE_HD inline void prefetch_around(const Chain &c, int e) {
#ifdef __CUDA_ARCH__
  const HE &a = c.he[e], &b = c.he[e ^ 1];
  prefetch_line(&c.he[a.next]); prefetch_line(&c.he[a.prev]);
  prefetch_line(&c.he[b.next]); prefetch_line(&c.he[b.prev]);
  if (a.nhf != NIL) prefetch_line(&c.he[a.nhf]);
  if (a.phf != NIL) prefetch_line(&c.he[a.phf]);
  prefetch_line(&c.seg[a.segd & 0x7FFF]);
  prefetch_line(&c.w[e >> 1]);
#else
  (void)c; (void)e;
#endif
}

// Returns false if the chain cannot go on (capacity exhausted / broken state).
//
// The three kinds of move are laid out in common phases rather than three branches:
// candidate -> ray (splits, flips) -> Hastings ratio -> DCEL primitives -> bookkeeping, and
// every out-of-line step (ray_exit, split_edge, add_edge, suppress_edge) has ONE call site.
// 32 chains share a warp; with one site per primitive the lanes that need it run it
// together whatever their kind of move.
LITE_HD inline bool smf_step(Chain &c, const Energy &G, const ChainParams &cp, const StepRng &R, TraceRec *tr) {
  if (c.H->status != ST_OK) return false;
  const uint64_t t = c.H->step;
  uint32_t r4[4];
  draw4(R, t, 0, r4);
  const double r0 = u53(r4[0], r4[1]);  // SMFChain::propose_modif_type, :3143-3151
  const double x = u53(r4[2], r4[3]);   // acceptance uniform, :3268
  const int type = (r0 <= cp.p_split) ? 0 : ((r0 <= cp.p_split + cp.p_merge) ? 1 : 2);
  c.H->counts[2 * type]++;
  if (tr) {
    tr->type = type; tr->status = 0; tr->r = 0; tr->e_var = 0; tr->x_accept = x;
    tr->e1s[0] = tr->e1s[1] = tr->e1t[0] = tr->e1t[1] = 0; tr->x = tr->angle = 0;
    tr->p1[0] = tr->p1[1] = tr->p2[0] = tr->p2[1] = 0;
    tr->e2s[0] = tr->e2s[1] = tr->e2t[0] = tr->e2t[1] = 0;
  }
  if (cp.prefetch & 2) {
    prefetch_line(&c.he[2 * (c.H->free_p >= 0 ? c.H->free_p : c.H->n_p)]);
    prefetch_line(&c.f_he[c.H->free_f >= 0 ? c.H->free_f : c.H->n_f]);
    prefetch_line(&c.seg[c.H->free_s >= 0 ? c.H->free_s : c.H->n_s]);
    prefetch_line(&c.nb_list[c.H->n_nb]);
    prefetch_line(&c.b_list[c.H->n_b]);
  }
  // ---- the candidate (TTessel::propose_split / _merge / _flip, :1755-1777)
  bool have = false;
  int e1 = -1;
  SplitMove sm; FlipMove fm; Ray ray;
  sm.valid = false; sm.e2 = -1; fm.valid = false; fm.e2 = -1; fm.e3 = -1; fm.right = false; ray.ok = false;
  if (type == 0) {
    if (c.H->n_seg + 1 > c.H->cap) { c.H->status = ST_CAPACITY; return false; }
    draw4(R, t, 1, r4);
    e1 = sample_halfedge(c, u53(r4[0], r4[1]));
    if (cp.prefetch & 1) prefetch_around(c, e1);
    const double xs = u53(r4[2], r4[3]);
    draw4(R, t, 2, r4);
    // angle = acos(1 - 2U), TTessel::propose_split :1757
    const double U = u53(r4[0], r4[1]);
    const double ca = 1.0 - 2.0 * U, sa = 2.0 * sqrt(U * (1.0 - U));
    const double ang = tr ? acos(ca) : 0.0;
    prepare_split(c, e1, xs, ang, ca, sa, sm, ray);
    have = true;
    if (tr) { tr->x = xs; tr->angle = ang; }
  } else if (type == 1) {
    if (c.H->n_nb > 0) {
      draw4(R, t, 1, r4);
      int n = (int)(u53(r4[0], r4[1]) * c.H->n_nb);
      if (n >= c.H->n_nb) n = c.H->n_nb - 1;
      e1 = c.seg[c.nb_list[n]].he;
      if (cp.prefetch & 1) prefetch_around(c, e1);
      have = true;
    }
  } else {
    if (c.H->n_b > 0) {
      draw4(R, t, 1, r4);
      int n = (int)(u53(r4[0], r4[1]) * c.H->n_b);
      if (n >= c.H->n_b) n = c.H->n_b - 1;
      const int sg = c.b_list[n];
      const int side = (u53(r4[2], r4[3]) < 0.5) ? 0 : 1;
      e1 = side == 0 ? (seg_start(c, sg) ^ 1) : seg_end(c, sg);
      if (cp.prefetch & 1) prefetch_around(c, e1);
      prepare_flip(c, e1, fm, ray);
      have = true;
    }
  }
  if (tr && have) {
    const Pt A = P_src(c, e1), B = P_tgt(c, e1);
    tr->e1s[0] = A.x; tr->e1s[1] = A.y; tr->e1t[0] = B.x; tr->e1t[1] = B.y;
    tr->status = (type == 1) ? 0 : -1;
  }
  // ---- the ray of a split or a flip: one call site
  int e2 = -1;
  Pt p2; p2.x = p2.y = 0;
  if (ray.ok) e2 = ray_exit(c, ray.start, ray.a, ray.b, ray.w, ray.p, ray.dx, ray.dy, ray.skip0, ray.skip1, p2);
  if ((cp.prefetch & 4) && e2 >= 0) prefetch_around(c, e2);
  bool valid = have && (type == 1 || e2 >= 0);
  if (type == 0) { sm.e2 = e2; sm.p2 = p2; sm.valid = valid; }
  if (type == 2) { fm.e2 = e2; fm.p2 = p2; fm.valid = valid; }
  // ---- Hastings ratio (SMFChain::Hasting_ratio x3, :3158-3252) and acceptance
  bool acc = false;
  double ev = 0;
  if (valid) {
    double dv[LITE_MAX_D];
    double r;
    if (type == 0) {
      int delta_nb = 1;
      const int sa = h_seg(c, e1), sb = h_seg(c, e2);
      const int n_a = c.seg[sa].nedges, n_b2 = c.seg[sb].nedges;
      if (n_a == 1 && !seg_bnd(sa)) delta_nb--;
      if (n_b2 == 1 && !seg_bnd(sb)) delta_nb--;
      r = cp.p_merge / cp.p_split;
      r *= (1 / kPi) * c.H->int_length / (c.H->n_nb + delta_nb);
      split_variation(c, G, sm, dv);
    } else if (type == 1) {
      r = cp.p_split / cp.p_merge;
      r *= kPi * c.H->n_nb / (c.H->int_length - 2 * h_len(c, e1));
      merge_variation(c, G, e1, dv);
    } else {
      int db = 0;
      // the four segment records are independent: asked for together, then tested
      const int i_short = h_seg(c, h_next(c, e1)), i_ext = h_seg(c, e2);
      const int n_e1 = c.seg[h_seg(c, e1)].nedges, n_e3 = c.seg[h_seg(c, fm.e3)].nedges;
      const int n_short = c.seg[i_short].nedges, n_ext = c.seg[i_ext].nedges;
      if (n_e1 == 2) db--;
      if (n_e3 == 1) db++;
      if (i_short != i_ext) {
        if (n_short == 2 && !seg_bnd(i_short)) db--;
        if (n_ext == 1 && !seg_bnd(i_ext)) db++;
      }
      const double sb = (double)c.H->n_b;
      r = (c.H->n_b == -db) ? sb : sb / (sb + db);
      flip_variation(c, G, fm, dv);
    }
    ev = dot_theta(G, dv);
    r *= exp(-ev);
    acc = (r > 1 || x < r);
    if (tr) {
      if (type != 1) {
        const Pt A = P_src(c, e2), B = P_tgt(c, e2);
        tr->e2s[0] = A.x; tr->e2s[1] = A.y; tr->e2t[0] = B.x; tr->e2t[1] = B.y;
        tr->p2[0] = p2.x; tr->p2[1] = p2.y;
        if (type == 0) { tr->p1[0] = sm.p1.x; tr->p1[1] = sm.p1.y; }
      }
      tr->r = r; tr->e_var = ev; tr->status = acc ? 2 : 1;
    }
  }
  // ---- TTessel::update (:1580-1701) as a sequence of shared primitives
  if (acc) {
    c.H->counts[2 * type + 1]++;
    // what the bookkeeping needs from the tessellation BEFORE it changes
    int s1 = -1, s1_i = -1, s2 = -1, prev_a = -1;
    double gone = 0;
    if (type == 2) {
      s1 = h_seg(c, e1); s1_i = h_seg(c, h_next(c, e1)); s2 = h_seg(c, fm.e3);
      prev_a = fm.right ? (e1 ^ 1) : h_prev(c, e1);  // split_from_vertex, :161-189
    }
    if (type == 1) {  // TTessel::update(Merge), :1604-1638
      const int ends[2] = {h_seg(c, h_next(c, e1)), h_seg(c, h_next(c, e1 ^ 1))};
      prefetch_line(&c.seg[ends[0]]); prefetch_line(&c.seg[ends[1]]);
      list_remove(c, h_seg(c, e1));
      for (int k = 0; k < 2; k++) {
        const int sx = ends[k];
        if (!seg_bnd(sx) && c.seg[sx].nedges <= 2 && c.seg[sx].kind == 2) { list_remove(c, sx); list_add(c, sx, 1); }
      }
      gone = h_len(c, e1);
    }
    // split_edge: at p1 on e1 (splits), then at p2 on e2 (splits and flips)
    int en[2] = {-1, -1};
    for (int k = (type == 0 ? 0 : 1); k < 2 && type != 1; k++) {
      const int he = (k == 0) ? e1 : e2;
      const Pt q = (k == 0) ? sm.p1 : p2;
      en[k] = split_edge(c, he, q.x, q.y);
    }
    int ne = -1;
    if (type != 1) ne = add_edge(c, type == 0 ? en[0] : prev_a, en[1], type == 0 ? -1 : fm.e3);
    if (type == 0) {  // TTessel::update(Split), :1580-1597
      c.H->int_length += 2 * h_len(c, ne);
      const int ends[2] = {h_seg(c, h_next(c, ne)), h_seg(c, h_next(c, ne ^ 1))};
      prefetch_line(&c.seg[ends[0]]); prefetch_line(&c.seg[ends[1]]);
      list_add(c, h_seg(c, ne), 1);
      for (int k = 0; k < 2; k++) {
        const int s = ends[k];
        if (!seg_bnd(s) && c.seg[s].nedges <= 2 && c.seg[s].kind == 1) { list_remove(c, s); list_add(c, s, 2); }
      }
    }
    int s2_i = -1;
    if (type == 2) {  // TTessel::update(Flip), :1644-1701
      c.H->int_length += 2 * h_len(c, ne);
      c.H->int_length -= 2 * h_len(c, e1);
      s2_i = h_seg(c, h_next(c, ne));
    }
    if (type != 0) suppress_edge(c, e1);
    if (type == 1) c.H->int_length -= 2 * gone;
    if (type == 2) {
      if (c.seg[s2].nedges <= 2 && c.seg[s2].kind == 1) { list_remove(c, s2); list_add(c, s2, 2); }
      if (c.seg[s1].nedges <= 1 && c.seg[s1].kind == 2) { list_remove(c, s1); list_add(c, s1, 1); }
      if (s2_i != s1_i) {
        if (!seg_bnd(s2_i) && c.seg[s2_i].nedges <= 2 && c.seg[s2_i].kind == 1) { list_remove(c, s2_i); list_add(c, s2_i, 2); }
        if (!seg_bnd(s1_i) && c.seg[s1_i].nedges <= 1 && c.seg[s1_i].kind == 2) { list_remove(c, s1_i); list_add(c, s1_i, 1); }
      }
    }
    c.H->energy += ev;
    flush_dirty(c);
  }
  c.H->step = t + 1;
  if (tr) {
    tr->n_seg = c.H->n_seg - 4; tr->n_nb = c.H->n_nb; tr->n_b = c.H->n_b; tr->int_length = c.H->int_length;
  }
  return c.H->status == ST_OK;
}

// the columns applications/checkACS.cpp:136-148 writes after each batch of steps
static const int N_STAT_COLS = 12;
LITE_HD inline void stat_row(const Chain &c, double *row, size_t stride) {
  int bnd_edges = 0;
  for (int s = 0; s < 4; s++) bnd_edges += c.seg[s].nedges;
  const int n_int = c.H->n_seg - 4;
  row[0 * stride] = (c.H->int_length - c.H->perimeter) / 2;  // l(T)
  row[1 * stride] = n_int;                                 // n(T)
  row[2 * stride] = c.H->n_b;                               // n_bl(T)
  row[3 * stride] = c.H->n_nb;                              // n_nbl(T)
  row[4 * stride] = (2 * n_int + 4) - bnd_edges;           // m(T): number_of_internal_vertices, :4105-4115
  for (int k = 0; k < 6; k++) row[(5 + k) * stride] = (double)c.H->counts[k];
  row[11 * stride] = c.H->energy;
}

// consistency check in the spirit of TTessel::is_valid (:1721-1753); 0 = valid
LITE_HD inline int check(const Chain &c) {
  int live_pairs = 0;
  for (int p = 0; p < c.H->n_p; p++) {
    if (!pair_alive(c, p)) { if (c.w[p] != 0.0) return 1; continue; }
    live_pairs++;
    for (int k = 0; k < 2; k++) {
      const int h = 2 * p + k;
      const int n = h_next(c, h);
      if (h_prev(c, n) != h || !same_pt(P_src(c, n), P_tgt(c, h))) return 2;
      if (h_face(c, n) != h_face(c, h)) return 3;
      const int nh = h_nhf(c, h);
      if (nh >= 0 && (h_phf(c, nh) != h || !same_pt(P_src(c, nh), P_tgt(c, h)) || h_seg(c, nh) != h_seg(c, h))) return 4;
      if (h_seg(c, h) != h_seg(c, h ^ 1) || h_dir(c, h) == h_dir(c, h ^ 1)) return 5;
    }
    const double l = dist(P_src(c, 2 * p), P_tgt(c, 2 * p));
    if (fabs(h_len(c, 2 * p) - l) > 1e-9) return 6;
  }
  const int n_int = c.H->n_seg - 4;
  if (live_pairs != 3 * n_int + 4) return 7;
  if (c.H->n_nb + c.H->n_b != n_int) return 8;
  for (int i = 0; i < c.H->n_nb; i++) { const int s = c.nb_list[i]; if (c.seg[s].nedges != 1 || c.seg[s].kind != 1 || c.seg[s].pos != i) return 9; }
  for (int i = 0; i < c.H->n_b; i++) { const int s = c.b_list[i]; if (c.seg[s].nedges < 2 || c.seg[s].kind != 2 || c.seg[s].pos != i) return 10; }
  for (int s = 0; s < c.H->n_s; s++) {
    if (c.seg[s].kind == NIL) continue;
    if (seg_start(c, s) != seg_start_walk(c, s) || seg_end(c, s) != seg_end_walk(c, s)) return 17;
    int n = 1, e = seg_start(c, s);
    if (h_seg(c, e) != s) return 11;
    for (int nx; (nx = h_nhf(c, e)) >= 0;) { e = nx; n++; }
    if (n != c.seg[s].nedges) return 12;
  }
  double tot = 0;
  for (int b = 0; b < (c.H->n_p + WBLOCK - 1) / WBLOCK; b++) tot += c.bsum[b];
  if (fabs(tot - c.H->int_length) > 1e-9 * (1 + tot)) return 13;
  for (int b = 0; b < (c.H->n_p + WBLOCK - 1) / WBLOCK; b++) {
    double s = 0;
    for (int k = 0; k < WBLOCK; k++) s += c.w[b * WBLOCK + k];
    if (fabs(s - c.bsum[b]) > 1e-11) return 14;
  }
  return 0;
}

}  // namespace smf
