// kernels.cuh — sm_100a device code for the LiTe pseudo-likelihood hot path.
//
// Layout in HBM: the tessellation is re-indexed by *slot*: one
// slot per halfedge bounding an inside face, faces stored contiguously in ccb
// order starting at outer_ccb().  A split candidate reads ~6 consecutive slots
// (double2 points, angles, lengths, corner angles) of one face; candidates are
// emitted in halfedge order, so a warp's 32 candidates share one or two faces
// and those loads are L1 broadcast hits.  Per-candidate statistics are stored
// structure-of-arrays (d columns of N doubles) for coalesced writes (K2) and
// reads (K3).
//
// Reference behaviour restated here (paths relative to /root/reference):
//   Split ctor + ray_exit_face      lib/ttessel.cpp:2090-2132, :4515-4576
//   Split/Merge/Flip modified_elements  :2135-2232, :2302-2404, :2459-2713
//   Energy::statistic_variation     :2914-2980
//   feature functions               :4124-4434
//   PseudoLikDiscrete::Get*         :3387-3443
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/lite_b200.h"

#define LITE_PI 3.14159265358979323846
#define LITE_HALF_PI (LITE_PI / 2)

// flag bits of a slot
#define SF_CORNER 1u  // the source vertex of the slot is a polygon corner of its face
#define SF_BND 2u     // the halfedge lies on the window boundary
#define SF_DIR 4u     // LineTes_halfedge::get_dir()
#define SF_GENERAL 8u // the face has holes or bounds an edge on both sides: gen_face.h evaluates it

struct TessDev {
  int n_slots, n_faces, n_he, n_seg;
  const int *err;      // != 0: the uploaded tessellation failed the device checks (k_validate); kernels then do nothing
  // per slot
  const double2 *pt;   // source point
  const double *ang;   // atan2(dy,dx) of the halfedge (lib/ttessel.cpp:2095)
  const double2 *udir; // its unit direction from the end points
  const double *len;   // LineTes_halfedge length
  const double *ca;    // interior angle at the source vertex (corners only)
  const uint32_t *flg; // SF_*
  const int *slot_he, *slot_seg, *slot_face;
  const int2 *sfo;     // (offset, count) of the slot's face: saves the slot->face->offset chain
  const double *pre;   // prefix shoelace sum: sum_{i<k} cross(V_i-O, V_{i+1}-O), O = first vertex of the face
  const double *plen;  // prefix sum of the halfedge lengths inside the face
  // per halfedge
  const int *he_slot, *he_twin, *he_src, *he_next_hf, *he_prev_hf;
  // per face
  const int *f_off, *f_cnt;
  const uint8_t *f_gen;  // 1 = not a simple polygon (holes, doubly-bounded edges); null if the tessellation has none
  const double *f_area, *f_perim, *f_sumang, *f_minang;
  const double *f_tot2;  // whole-face shoelace sum (twice the area) by the same prefix summation
  const double *stot2;   // the same, copied to every slot of the face
  // per segment
  const int *seg_n_edges, *seg_he_start, *seg_he_end;
  const uint8_t *seg_bnd;
  // length-weighted sampling: cumulative halfedge lengths in SLOT order (slot m
  // ends at cum[m]); guide[b] = first m with cum[m] > b*total_len/n_guide
  const double *cum;
  const int *guide;
  int n_cum, n_guide;
  double total_len, guide_scale;  // guide_scale = n_guide/total_len
};

struct Prog {
  int d;
  int fid[LITE_MAX_D];
  uint32_t mask;  // bit i set if feature id i is used
};

struct Theta {
  double v[LITE_MAX_D];
};

#include "philox.h"

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 ldpt(const double2 *p, int i) { return __ldg(p + i); }
__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by) {
  return ax * by - ay * bx;
}
__device__ __forceinline__ double plen(double2 a, double2 b) {
  double x = b.x - a.x, y = b.y - a.y;
  return sqrt(x * x + y * y);
}
__device__ __forceinline__ double wrap_pi(double d) {
  // to (-pi, pi]
  return d - (2 * LITE_PI) * rint(d * (1.0 / (2 * LITE_PI)));
}
// acos(c) for the cosine c of an angle in [0, pi] whose sine s = sqrt(1 - c^2) >= 0 is already
// known (a sampled proposal: c = 1 - 2U, s = 2 sqrt(U (1 - U))).  The smaller of |c| and s is at
// most sin(pi/4), so one odd polynomial asin(z) = z + z^3 Q(z^2) on z^2 <= 1/2 serves every case
// without the square root and the two divergent branches of the library acos (15.8 % of the
// kernel's instructions, profiles/r1_k_split.txt); |error| <= 1 ulp of the result.  Q: degree-18
// interpolant at the Chebyshev nodes of [0, 0.5], fitted in 60-digit arithmetic.
__constant__ double c_asin_q[19] = {   // Q, ascending powers of z^2 (constant bank: the DFMAs read them as operands)
    0.16666666666666669,   0.07499999999998412,  0.044642857146655654, 0.030381944084992746, 0.022372177006627192,
    0.017352221536902322,  0.013975730224099243, 0.01139924543091176,  0.011310930442084787, -0.003294232668448929,
    0.07377365207108956,   -0.28099565631446904, 0.9522680642152734,   -2.353535874499935,   4.384722306509683,
    -5.883246068449167,    5.436394922152539,    -3.0948600063047733,  0.8387686144806944};
__device__ __forceinline__ double acos_cs(double c, double s) {
  const double ac = fabs(c);
  const bool small_c = ac <= s;
  const double z = small_c ? ac : s;
  const double w = z * z;
#ifdef LITE_ACOS_HORNER
  double q = c_asin_q[18];
#pragma unroll
  for (int k = 17; k >= 0; k--) q = fma(q, w, c_asin_q[k]);
#else
  // even and odd coefficients as two Horner chains in w^2: the same 19 multiply-adds (+2), half
  // the dependent latency (the kernel waits on such chains: issue slots are 70 % busy)
  const double w2 = w * w;
  double qe = c_asin_q[18], qo = c_asin_q[17];
#pragma unroll
  for (int k = 16; k >= 0; k -= 2) qe = fma(qe, w2, c_asin_q[k]);
#pragma unroll
  for (int k = 15; k >= 1; k -= 2) qo = fma(qo, w2, c_asin_q[k]);
  const double q = fma(qo, w, qe);
#endif
  const double r = fma(z * w, q, z);  // asin(z)
  // |c| <= s: acos(c) = pi/2 - asin(c);  otherwise the angle is r (c > 0) or pi - r (c < 0)
  const double base = small_c ? LITE_HALF_PI : (c > 0.0 ? 0.0 : LITE_PI);
  const double sgn = (small_c == (c > 0.0)) ? -1.0 : 1.0;
  return fma(sgn, r, base);
}

// n / d for operands in the normal range, without the library's special-case branch: 20-bit
// reciprocal seed, two Newton steps, one correction of the quotient (error <= 1 ulp)
__device__ __forceinline__ double div_fast(double n, double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  const double q = n * r;
  return fma(fma(-d, q, n), r, q);
}

// angle_between_vectors, lib/ttessel.cpp:4255-4268
__device__ __forceinline__ double angle_between(double ax, double ay, double bx, double by) {
  double n1 = ax * ax + ay * ay, n2 = bx * bx + by * by;
  double p = ax * bx + ay * by;
  double c = p / sqrt(n1 * n2);
  c = fmin(1.0, fmax(-1.0, c));
  return acos(c);
}

// ---------------------------------------------------------------------------
// virtual polygons: a few explicit points and cyclic slot ranges of faces
// ---------------------------------------------------------------------------
#define VP_MAX 6
struct VPoly {
  int np, total;
  int kind[VP_MAX];  // 0 point, 1 range
  double x[VP_MAX], y[VP_MAX];
  int cf[VP_MAX];  // corner flag of an explicit point
  int off[VP_MAX], n[VP_MAX], s0[VP_MAX], cnt[VP_MAX];
};
__device__ inline void vp_init(VPoly &P) { P.np = 0; P.total = 0; }
__device__ inline void vp_point(VPoly &P, double x, double y, int corner) {
  int i = P.np++;
  P.kind[i] = 0; P.x[i] = x; P.y[i] = y; P.cf[i] = corner; P.cnt[i] = 1;
  P.total += 1;
}
__device__ inline void vp_range(VPoly &P, int off, int n, int s0, int cnt) {
  if (cnt <= 0) return;
  int i = P.np++;
  P.kind[i] = 1; P.off[i] = off; P.n[i] = n; P.s0[i] = ((s0 % n) + n) % n; P.cnt[i] = cnt;
  P.total += cnt;
}
__device__ inline void vp_get(const TessDev &T, const VPoly &P, int i, double &x, double &y,
                              int &corner) {
  int k = 0;
  while (i >= P.cnt[k]) { i -= P.cnt[k]; k++; }
  if (P.kind[k] == 0) {
    x = P.x[k]; y = P.y[k]; corner = P.cf[k];
  } else {
    int s = P.s0[k] + i;
    if (s >= P.n[k]) s -= P.n[k];
    double2 v = ldpt(T.pt, P.off[k] + s);
    x = v.x; y = v.y;
    corner = (int)(T.flg[P.off[k] + s] & SF_CORNER);
  }
}

struct FaceFeat {
  double area, perim, sumang, minang;
};
// face_area_2 / face_perimeter / face_sum_of_angles / min_angle inputs of one
// (simplified) polygon, lib/ttessel.cpp:4215-4239, :4285-4305, :4348-4367.
// `need_ang`: compute the acos-based quantities.
__device__ inline FaceFeat vp_features(const TessDev &T, const VPoly &P, bool need_ang) {
  FaceFeat r;
  int n = P.total;
  double x0, y0; int c0;
  vp_get(T, P, 0, x0, y0, c0);
  double px, py; int pc;   // previous vertex (i-1)
  vp_get(T, P, n - 1, px, py, pc);
  double cx = x0, cy = y0; int cc = c0;  // current vertex i
  double area2 = 0, perim = 0, sumang = 0, minang = CUDART_INF, seed = -1.0;
  for (int i = 0; i < n; i++) {
    double nx, ny; int nc;
    if (i + 1 < n) vp_get(T, P, i + 1, nx, ny, nc);
    else { nx = x0; ny = y0; nc = c0; }
    area2 += cross2(cx - x0, cy - y0, nx - x0, ny - y0);
    double ex = nx - cx, ey = ny - cy;
    perim += sqrt(ex * ex + ey * ey);
    if (need_ang && cc) {
      double a = angle_between(px - cx, py - cy, ex, ey);
      if (seed < 0) seed = LITE_PI - a;  // turning angle at the first corner (:4354)
      double q = LITE_HALF_PI - a;
      if (q >= 0) sumang += q;
      minang = fmin(minang, a);
    }
    px = cx; py = cy; cx = nx; cy = ny; cc = nc;
  }
  r.area = 0.5 * area2;
  r.perim = perim;
  r.sumang = sumang;
  r.minang = (seed >= 0) ? fmin(seed, minang) : 0.0;
  return r;
}
__device__ __forceinline__ double shape_of(double perim, double area) {
  // face_shape: perimeter/(4*pow(area^2,0.25)), lib/ttessel.cpp:4247-4249
  return perim / (4.0 * sqrt(fabs(area)));
}

// Accumulate  sign * f(face polygon)  for every face feature of the program.
__device__ inline void add_face_feats(const Prog &G, double *acc, double sign, const FaceFeat &F) {
  for (int i = 0; i < G.d; i++) {
    switch (G.fid[i]) {
      case LITE_FEAT_FACE_NUMBER: acc[i] += sign; break;
      case LITE_FEAT_FACE_AREA_2: acc[i] += sign * (F.area * F.area); break;
      case LITE_FEAT_FACE_PERIMETER: acc[i] += sign * F.perim; break;
      case LITE_FEAT_FACE_SHAPE: acc[i] += sign * shape_of(F.perim, F.area); break;
      case LITE_FEAT_FACE_SUM_OF_ANGLES: acc[i] += sign * F.sumang; break;
      case LITE_FEAT_MIN_ANGLE: acc[i] += sign * F.minang; break;
      default: break;
    }
  }
}
// features of an unmodified face of the tessellation: precomputed by k_prep_faces in
// face2poly order, so a deleted parent face costs four loads instead of a polygon walk
__device__ __forceinline__ FaceFeat parent_feat(const TessDev &T, int f) {
  FaceFeat r;
  r.area = __ldg(T.f_area + f); r.perim = __ldg(T.f_perim + f);
  r.sumang = __ldg(T.f_sumang + f); r.minang = __ldg(T.f_minang + f);
  return r;
}
#define MASK_FACE_ANG ((1u << LITE_FEAT_FACE_SUM_OF_ANGLES) | (1u << LITE_FEAT_MIN_ANGLE))
#define MASK_FACE_ANY                                                                     \
  ((1u << LITE_FEAT_FACE_NUMBER) | (1u << LITE_FEAT_FACE_AREA_2) |                        \
   (1u << LITE_FEAT_FACE_PERIMETER) | (1u << LITE_FEAT_FACE_SHAPE) | MASK_FACE_ANG)
#define MASK_FACE_AREA ((1u << LITE_FEAT_FACE_AREA_2) | (1u << LITE_FEAT_FACE_SHAPE))
#define MASK_NEED_D01                                                                     \
  ((1u << LITE_FEAT_FACE_PERIMETER) | (1u << LITE_FEAT_FACE_SHAPE) | (1u << LITE_FEAT_EDGE_LENGTH_INTERNAL))
#define MASK_FACE_GEOM                                                                    \
  ((1u << LITE_FEAT_FACE_AREA_2) | (1u << LITE_FEAT_FACE_SHAPE) | (1u << LITE_FEAT_MIN_ANGLE))

// ---------------------------------------------------------------------------
// K0: per-face preparation (slot arrays + parent-face features)
// ---------------------------------------------------------------------------
struct PrepArgs {
  // raw SoA on device
  const double *vx, *vy;
  const int *he_src, *he_twin, *he_next, *he_face, *he_seg, *he_next_hf;
  const uint8_t *he_dir, *face_inside;
  const uint8_t *f_gen;  // null when every face is a simple polygon
  const double *he_length;
  const int *face_he_offset, *face_he_count, *face_he_list;
  int n_faces;
  // outputs
  double2 *pt; double2 *udir; double *ang, *len, *ca, *el; uint32_t *flg;
  int *slot_he, *slot_seg, *slot_face, *he_slot;
  double *f_area, *f_perim, *f_sumang, *f_minang;
  int *err;  // consistency errors
};

// ---------------------------------------------------------------------------
// Export assertions on the device (in the spirit of TTessel::is_valid, lib/ttessel.cpp:1721-1753).
// lite_tess_upload used to run them on four host threads in front of the copy (0.2-0.27 ms of
// a 1.2 ms end-to-end step); here they run at L2 speed behind it.  Every index that a later
// kernel dereferences is range-checked before anything follows it; the first violation leaves
// code*2^24 + element in *err, every kernel that reads the tessellation returns at once when
// *err != 0, and the host reports it at its next synchronisation (or right away with
// LITE_OPT_VALIDATE_SYNC).
// ---------------------------------------------------------------------------
struct ValidateArgs {
  int nV, nH, nF, nS, nFH, nNB, nB;
  const int *he_src, *he_twin, *he_next, *he_face, *he_seg, *he_next_hf, *he_prev_hf;
  const uint8_t *face_inside, *seg_bnd, *f_gen;
  const int *f_off, *f_cnt, *f_list, *seg_he_start, *seg_he_end, *seg_n_edges, *nb_list, *b_list;
  int *err;
  int *he_slot;  // set to -1 here (k_prep_slots fills the listed ones behind this kernel)
};
enum { LITE_BAD_HALFEDGE = 1, LITE_BAD_FACE = 2, LITE_BAD_FACE_LIST = 3, LITE_BAD_SEGMENT = 4, LITE_BAD_NON_BLOCKING = 5,
       LITE_BAD_BLOCKING = 6 };
__device__ __forceinline__ void flag_bad(int *err, int code, int elem) {
  atomicCAS(err, 0, (code << 24) | (elem & 0xffffff));
}
__global__ void k_validate(ValidateArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.nH) {  // links in range, twin of twin
    A.he_slot[i] = -1;
    const int f = A.he_face[i], tw = A.he_twin[i], v = A.he_src[i], sg = A.he_seg[i], nx = A.he_next[i];
    const int nhf = A.he_next_hf[i], phf = A.he_prev_hf[i];
    bool ok = (unsigned)f < (unsigned)A.nF && (unsigned)tw < (unsigned)A.nH && (unsigned)v < (unsigned)A.nV &&
              (unsigned)sg < (unsigned)A.nS && (unsigned)nx < (unsigned)A.nH && nhf >= -1 && nhf < A.nH && phf >= -1 &&
              phf < A.nH;
    if (ok) ok = A.he_twin[tw] == i;
    if (!ok) flag_bad(A.err, LITE_BAD_HALFEDGE, i);
  }
  if (i < A.nF) {  // list range; an inside face lists at least three halfedges
    const int off = A.f_off[i], cnt = A.f_cnt[i];
    if (cnt < 0 || off < 0 || off > A.nFH || cnt > A.nFH - off || (A.face_inside[i] && cnt < 3)) flag_bad(A.err, LITE_BAD_FACE, i);
  }
  if (i < A.nS) {
    const int a = A.seg_he_start[i], b = A.seg_he_end[i];
    if ((unsigned)a >= (unsigned)A.nH || (unsigned)b >= (unsigned)A.nH || A.seg_n_edges[i] < 1) flag_bad(A.err, LITE_BAD_SEGMENT, i);
  }
  if (i < A.nNB) {
    const int sg = A.nb_list[i];
    if ((unsigned)sg >= (unsigned)A.nS || A.seg_n_edges[sg] != 1 || A.seg_bnd[sg]) flag_bad(A.err, LITE_BAD_NON_BLOCKING, i);
  }
  if (i < A.nB) {
    const int sg = A.b_list[i];
    if ((unsigned)sg >= (unsigned)A.nS || A.seg_n_edges[sg] < 2 || A.seg_bnd[sg]) flag_bad(A.err, LITE_BAD_BLOCKING, i);
  }
  if (i < A.nFH) {  // a list entry is a halfedge (that it bounds the face and chains to the next entry: k_prep_all)
    if ((unsigned)A.f_list[i] >= (unsigned)A.nH) flag_bad(A.err, LITE_BAD_FACE_LIST, i);
  }
}

// Everything per slot with one thread per slot: the loads of a slot chain four deep (list ->
// halfedge -> vertex / twin -> face), and one thread per FACE walking its ~6 slots one after
// the other took 50 us at 10^4 cells.  The face of slot m is the last one whose list starts at
// or before m (lite_tess_upload has checked that the lists tile the slots in face order).
__global__ void k_prep_slots(PrepArgs A, int n_slots) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_slots || *A.err) return;
  int f;
  {
    int lo = 0, hi = A.n_faces - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(A.face_he_offset + mid) <= m) lo = mid; else hi = mid - 1;
    }
    f = lo;
  }
  if (!A.face_inside[f] || m >= A.face_he_offset[f] + A.face_he_count[f]) { A.slot_face[m] = -1; return; }  // a slot of an outside face
  A.slot_face[m] = f;
  const int n = A.face_he_count[f], off = A.face_he_offset[f], k = m - off;
  int h = A.face_he_list[off + k];
  int hp, hn;
  const bool general = A.f_gen && A.f_gen[f];
  if (general) {
    // the slots of such a face list several boundary cycles one after the other: neighbours
    // come from the DCEL (previous halfedge: rotate around the source vertex)
    hn = A.he_next[h];
    hp = h;
    for (int o = h, guard = 0; guard < 64; guard++) {
      const int i = A.he_twin[o];
      if (A.he_next[i] == h) { hp = i; break; }
      o = A.he_next[i];
    }
    if (A.he_face[h] != f) flag_bad(A.err, LITE_BAD_FACE_LIST, m);
  } else {
    hp = A.face_he_list[off + (k == 0 ? n - 1 : k - 1)];
    hn = A.face_he_list[off + (k + 1 == n ? 0 : k + 1)];
    if (A.he_next[h] != hn || A.he_face[h] != f) flag_bad(A.err, LITE_BAD_FACE_LIST, m);
  }
  int vs = A.he_src[h], vt = A.he_src[hn];
  double sx = A.vx[vs], sy = A.vy[vs], tx = A.vx[vt], ty = A.vy[vt];
  A.pt[m] = make_double2(sx, sy);
  A.ang[m] = atan2(ty - sy, tx - sx);
  {
    const double ex = tx - sx, ey = ty - sy;
    const double inv = 1.0 / sqrt(ex * ex + ey * ey);
    A.udir[m] = make_double2(ex * inv, ey * inv);
  }
  A.len[m] = A.he_length[h];
  uint32_t fl = 0;
  int nhf = A.he_next_hf[hp];
  if (nhf < 0 || nhf != A.he_next[hp]) fl |= SF_CORNER;  // face2poly, lib/ttessel.cpp:4325-4327
  // the two expensive terms of the face features, one slot per thread (k_prep_faces only adds
  // them up in the reference's order): the edge length from its end points and the interior
  // angle at the source, angle_between at a corner and pi elsewhere
  {
    const double ex = tx - sx, ey = ty - sy;
    A.el[m] = sqrt(ex * ex + ey * ey);
    double a = LITE_PI;
    if (fl & SF_CORNER) {
      const int vp = A.he_src[hp];
      a = angle_between(A.vx[vp] - sx, A.vy[vp] - sy, ex, ey);
    }
    A.ca[m] = a;
  }
  if (!A.face_inside[A.he_face[A.he_twin[h]]]) fl |= SF_BND;
  if (A.he_dir[h]) fl |= SF_DIR;
  if (general) fl |= SF_GENERAL;
  A.flg[m] = fl;
  A.slot_he[m] = h;
  A.slot_seg[m] = A.he_seg[h];
  A.he_slot[h] = m;
}

__global__ void k_prep_faces(TessDev T, const double *ca, const double *el, double *f_area, double *f_perim,
                             double *f_sumang, double *f_minang, double *pre, double *plen,
                             double *f_tot2, int2 *sfo, double *stot2) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= T.n_faces || *T.err) return;
  int n = T.f_cnt[f];
  if (n == 0) return;
  int off = T.f_off[f];
  {  // prefix sums used by the O(1) child-area / child-perimeter formulas of split_eval
    double2 O = ldpt(T.pt, off), Vk = O;
    double acc = 0, lacc = 0;
    for (int k = 0; k < n; k++) {
      double2 Vn = ldpt(T.pt, off + (k + 1 == n ? 0 : k + 1));
      pre[off + k] = acc;
      plen[off + k] = lacc;
      sfo[off + k] = make_int2(off, n);
      acc += cross2(Vk.x - O.x, Vk.y - O.y, Vn.x - O.x, Vn.y - O.y);
      lacc += T.len[off + k];
      Vk = Vn;
    }
    f_tot2[f] = acc;
    for (int k = 0; k < n; k++) stot2[off + k] = acc;
  }
  // Corner angles at the slot sources and the features of the unmodified face, in one walk in
  // face2poly order (first vertex = target of the outer_ccb halfedge = source of slot 1), with
  // the arithmetic of vp_features (lib/ttessel.cpp:4215-4239, :4285-4305, :4348-4367): shoelace
  // from the first vertex, edge lengths from the end points, angle_between at the corners, the
  // minimum seeded with the turning angle at the first corner.
  {
    const double2 V0 = ldpt(T.pt, off + (1 == n ? 0 : 1));
    double2 C = V0;   // current vertex
    double area2 = 0, perim = 0, sumang = 0, minang = CUDART_INF, seed = -1.0;
    int kc = (1 == n) ? 0 : 1, kn = (kc + 1 == n) ? 0 : kc + 1;   // slot of the current vertex, of the next
    for (int i = 0; i < n; i++) {
      const int sc = off + kc;                            // slot whose source is the current vertex
      const double2 Nx = (i + 1 < n) ? ldpt(T.pt, off + kn) : V0;
      area2 += cross2(C.x - V0.x, C.y - V0.y, Nx.x - V0.x, Nx.y - V0.y);
      perim += __ldg(el + sc);
      if (T.flg[sc] & SF_CORNER) {
        const double a = __ldg(ca + sc);
        if (seed < 0) seed = LITE_PI - a;
        const double q = LITE_HALF_PI - a;
        if (q >= 0) sumang += q;
        minang = fmin(minang, a);
      }
      C = Nx;
      kc = kn; kn = (kn + 1 == n) ? 0 : kn + 1;
    }
    f_area[f] = 0.5 * area2; f_perim[f] = perim; f_sumang[f] = sumang;
    f_minang[f] = (seed >= 0) ? fmin(seed, minang) : 0.0;
  }
}

// guide table of the sampler: guide[b] = first slot m with cum[m] > b * step, clamped
__global__ void k_build_guide(const double *cum, int n_cum, double step, int n_guide, int *guide) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_guide) return;  // reads cum only (host-built), no guard needed
  const double lo = (double)b * step;
  int l = 0, h = n_cum - 1;
  while (l < h) {
    int mid = (l + h) >> 1;
    if (__ldg(cum + mid) > lo) h = mid; else l = mid + 1;
  }
  guide[b] = l;
}

// ---------------------------------------------------------------------------
// ray exit search in a hole-free face (ray_exit_face, lib/ttessel.cpp:4515-4576)
//   line a*X + b*Y = w through the ray source p, direction (dx,dy);
//   skip0/skip1: slots whose hit is the ray source itself (distance 0).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int ray_exit(const TessDev &T, int off, int n, double a, double b,
                                        double w, double2 p, double dx, double dy, int skip0,
                                        int skip1, double2 &hit) {
  double2 V0 = ldpt(T.pt, off);
  double s0 = fma(a, V0.x, fma(b, V0.y, -w));
  double2 Vk = V0; double sk = s0;
  double best = 1.7976931348623157e308;
  int k2 = -1;
  for (int k = 0; k < n; k++) {
    double2 Vn; double sn;
    if (k + 1 < n) { Vn = ldpt(T.pt, off + k + 1); sn = fma(a, Vn.x, fma(b, Vn.y, -w)); }
    else { Vn = V0; sn = s0; }
    if (k != skip0 && k != skip1) {
      bool cr = (sk <= 0.0 && sn >= 0.0) || (sk >= 0.0 && sn <= 0.0);
      if (cr && sk != sn) {
        double sg = sk / (sk - sn);
        double X = fma(sg, Vn.x - Vk.x, Vk.x), Y = fma(sg, Vn.y - Vk.y, Vk.y);
        double rx = X - p.x, ry = Y - p.y;
        if (rx * dx + ry * dy >= 0.0) {
          double len = sqrt(rx * rx + ry * ry);
          if (len != 0.0 && len < best) { best = len; k2 = k; hit = make_double2(X, Y); }
        }
      }
    }
    Vk = Vn; sk = sn;
  }
  return k2;
}

// ---------------------------------------------------------------------------
// K1+K2: one split candidate  (Split ctor + modified_elements + statistic_variation)
//
// Cost model.  Candidates reach the kernel sorted by slot (the sampler is
// stratified along the cumulative halfedge length), and a warp owns a contiguous
// run of them, so the lanes of a warp sit on the same halfedge of the same face
// and each lane meets a new slot only every few iterations: the slot's data
// (SlotData) stay in registers in between and the sampler needs no memory access.
// After the crossed-edge search the work is O(min(child sizes)): the child with
// fewer parent vertices is walked from p1, the other is the parent minus it.
// The feature program is a template policy: ProgDyn reads it from the kernel
// argument (any registered energy), ProgACS is the sim_acs energy folded at
// compile time; both run the same code.
// ---------------------------------------------------------------------------
struct SplitOut {
  int e2;       // halfedge id, -1 if no exit was found
  double2 p1, p2;
};

struct ProgDyn {
  static constexpr bool kStatic = false;
  __device__ static __forceinline__ uint32_t mask(const Prog &G) { return G.mask; }
  __device__ static __forceinline__ int d(const Prog &G) { return G.d; }
  __device__ static __forceinline__ int fid(const Prog &G, int i) { return G.fid[i]; }
};
// demos/sim_acs.cpp:64-76 in CatVector order: edges edge_length, faces face_area_2,
// faces face_sum_of_angles, segs is_segment_internal
struct ProgACS {
  static constexpr bool kStatic = true;
  __device__ static __forceinline__ constexpr uint32_t mask(const Prog &) {
    return (1u << LITE_FEAT_EDGE_LENGTH) | (1u << LITE_FEAT_FACE_AREA_2) |
           (1u << LITE_FEAT_FACE_SUM_OF_ANGLES) | (1u << LITE_FEAT_IS_SEGMENT_INTERNAL);
  }
  __device__ static __forceinline__ constexpr int d(const Prog &) { return 4; }
  __device__ static __forceinline__ constexpr int fid(const Prog &, int i) {
    return i == 0 ? LITE_FEAT_EDGE_LENGTH
                  : i == 1 ? LITE_FEAT_FACE_AREA_2
                           : i == 2 ? LITE_FEAT_FACE_SUM_OF_ANGLES : LITE_FEAT_IS_SEGMENT_INTERNAL;
  }
};

// sum of a prefix array over slots i, i+1, .., j-1 (cyclic, empty when i == j)
__device__ __forceinline__ double chain(const double *pre, double tot, int off, int i, int j) {
  double a = __ldg(pre + off + j) - __ldg(pre + off + i);
  return (j >= i) ? a : a + tot;
}

// what a candidate needs from the halfedge it starts on
struct SlotData {
  int off, n, k1;     // face slot range, position of the halfedge in it
  double2 P, Q;       // source and target of the halfedge
  double ang1;        // atan2 of its direction
  double ux, uy;      // its unit direction (cos, sin of ang1) from the end points
  double half_tot;    // area of its face (shoelace, prefix summation)
  uint32_t fl1;       // SF_* of the slot
};
__device__ __forceinline__ SlotData load_slot(const TessDev &T, int s1) {
  SlotData S;
  const int2 fo = __ldg(T.sfo + s1);
  S.off = fo.x; S.n = fo.y; S.k1 = s1 - fo.x;
  S.P = ldpt(T.pt, s1);
  S.Q = ldpt(T.pt, fo.x + ((S.k1 + 1 == fo.y) ? 0 : S.k1 + 1));
  S.ang1 = __ldg(T.ang + s1);
  {
    const double2 u = __ldg(T.udir + s1);
    S.ux = u.x; S.uy = u.y;
  }
  S.half_tot = 0.5 * __ldg(T.stot2 + s1);
  S.fl1 = __ldg(T.flg + s1);
  return S;
}

// out of line on purpose: only taken for the few candidates that start or end on a
// window-boundary edge (edge_length as written) or when the feature is registered
__device__ __noinline__ double edge_cut_delta(double ax, double ay, double mx, double my, double bx,
                                              double by) {
  // |a-m| + |m-b| - |a-b| for a point m on the edge (a,b)
  return sqrt((mx - ax) * (mx - ax) + (my - ay) * (my - ay)) +
         sqrt((bx - mx) * (bx - mx) + (by - my) * (by - my)) -
         sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
}
// min over the parent's corner angles at slots ks .. ke (cyclic, inclusive)
__device__ __noinline__ double corner_min(const uint32_t *flg, const double *ca, int off, int n, int ks,
                                          int ke) {
  double m = CUDART_INF;
  for (int k = ks;; k = (k + 1 == n) ? 0 : k + 1) {
    if (__ldg(flg + off + k) & SF_CORNER) m = fmin(m, __ldg(ca + off + k));
    if (k == ke) break;
  }
  return m;
}

// The proposal angle arrives as (cos, sin) and, when a registered feature or the record
// needs the angle itself, as `angle`.  The line's normal is a = sin(l), b = -cos(l) with
// l = ang1 + angle (lib/ttessel.cpp:2094-2100), formed by the addition formulas from the
// halfedge's unit direction: for a sampled candidate cos(angle) = 1 - 2U is the draw
// itself, so neither acos nor sincos is on the path unless an angle feature asks for it
// (they were 24 % of the kernel's instructions, profiles/r1_k_split.txt).
#ifndef LITE_SCAN_UNROLL
#define LITE_SCAN_UNROLL 4
#endif
#define LITE_PRAGMA_(x) _Pragma(#x)
#define LITE_UNROLL(n) LITE_PRAGMA_(unroll n)
template <bool WITH_STAT, class PG>
__device__ __forceinline__ void split_eval(const TessDev &T, const Prog &G, const SlotData &S, double x,
                                           double angle, double cos_angle, double sin_angle,
                                           double *__restrict__ stat_col0, size_t cap, SplitOut &o) {
  const int off = S.off, n = S.n, k1 = S.k1;
  const int k1n = (k1 + 1 == n) ? 0 : k1 + 1;
  const double2 P = S.P, Q = S.Q;
  // lib/ttessel.cpp:2094-2105, evaluated as the reference does (no contraction)
  const double l = S.ang1 + angle;  // only read by the angle features
  const double a = S.uy * cos_angle + S.ux * sin_angle;   // sin(l)
  const double b = S.uy * sin_angle - S.ux * cos_angle;   // -cos(l)
  const double u = __dadd_rn(__dmul_rn(a, P.x), __dmul_rn(b, P.y));
  const double v = __dadd_rn(__dmul_rn(a, Q.x), __dmul_rn(b, Q.y));
  const double w = __dadd_rn(u, __dmul_rn(x, __dadd_rn(v, -u)));
  // p1 = e1 ∩ line(a,b,-w)   (:2110)
  const double sP = fma(a, P.x, fma(b, P.y, -w)), sQ = fma(a, Q.x, fma(b, Q.y, -w));
  const double t1 = div_fast(sP, sP - sQ);
  const double2 p1 = make_double2(fma(t1, Q.x - P.x, P.x), fma(t1, Q.y - P.y, P.y));
  // ray direction (:2121-2127): Line_2(a,b,c) has direction (b,-a)
  double dx, dy;
  if (sQ < 0.0) { dx = b; dy = -a; } else { dx = -b; dy = a; }
  const int skip1 = (x == 0.0) ? (k1 == 0 ? n - 1 : k1 - 1) : -1;

  // crossed-edge search (ray_exit_face, :4515-4576).  A vertex is classified once
  // by the sign of its residual s = a*X + b*Y - w, so the scan is watertight; a
  // halfedge crosses the line when its end residuals have opposite signs or one is
  // zero (s_k * s_{k+1} <= 0, not both zero).  In a convex face exactly one
  // halfedge other than e1 does, unless the line passes through a vertex.
  int cnt = 0, k2 = -1;
#ifndef LITE_SCAN_LOOP
  if (n <= 32) {
    // sign bit of every vertex residual, shifted in from the right (vertex k ends at bit
    // n-1-k: one funnel shift per vertex); `zacc` stays non-zero while no residual is exactly
    // zero (then a halfedge crosses iff the signs of its ends differ)
    uint32_t neg = 0, zacc = 0xffffffffu;
LITE_UNROLL(LITE_SCAN_UNROLL)
    for (int k = 0; k < n; k++) {
      const double2 V = ldpt(T.pt, off + k);
      const double sv = fma(a, V.x, fma(b, V.y, -w));
      const uint32_t hi = (uint32_t)__double2hiint(sv), lo = (uint32_t)__double2loint(sv);
      neg = __funnelshift_l(hi, neg, 1);
      zacc = min(zacc, (hi & 0x7fffffffu) | lo);
    }
    if (zacc != 0u) {
      const uint32_t full = (n == 32) ? 0xffffffffu : ((1u << n) - 1u);
      const uint32_t nxt = ((neg << 1) | (neg >> (n - 1))) & full;   // sign of vertex k+1 at the bit of vertex k
      uint32_t cross = (neg ^ nxt) & ~(1u << (n - 1 - k1));          // bit n-1-k: halfedge k crosses
      if (skip1 >= 0) cross &= ~(1u << (n - 1 - skip1));
      cnt = __popc(cross);
      k2 = cross ? n - 32 + __clz(cross) : -1;                       // the first crossing halfedge in slot order
    } else {
      cnt = -1;  // a vertex lies exactly on the line: the general loop below decides
    }
  } else {
    cnt = -1;
  }
  if (cnt < 0)
#endif
  {
    cnt = 0; k2 = -1;
    const double2 V0 = ldpt(T.pt, off);
    const double s0 = fma(a, V0.x, fma(b, V0.y, -w));
    double sk = s0;
#pragma unroll 4
    for (int k = 1; k < n; k++) {
      const double2 Vn = ldpt(T.pt, off + k);
      const double sn = fma(a, Vn.x, fma(b, Vn.y, -w));
      if (sk * sn <= 0.0 && sk != sn && (k - 1) != k1 && (k - 1) != skip1) { if (cnt == 0) k2 = k - 1; cnt++; }
      sk = sn;
    }
    if (sk * s0 <= 0.0 && sk != s0 && (n - 1) != k1 && (n - 1) != skip1) { if (cnt == 0) k2 = n - 1; cnt++; }
  }
  double2 p2 = make_double2(0.0, 0.0);
  if (cnt == 1) {
    // the usual case (a convex face, the line through no vertex): one candidate halfedge, so the
    // nearest-hit comparison needs no distance: forward and not at the source
    const int kn = (k2 + 1 == n) ? 0 : k2 + 1;
    const double2 Vk = ldpt(T.pt, off + k2), Vn = ldpt(T.pt, off + kn);
    const double sk = fma(a, Vk.x, fma(b, Vk.y, -w)), sn = fma(a, Vn.x, fma(b, Vn.y, -w));
    const double sg = div_fast(sk, sk - sn);
    const double X = fma(sg, Vn.x - Vk.x, Vk.x), Y = fma(sg, Vn.y - Vk.y, Vk.y);
    const double rx = X - p1.x, ry = Y - p1.y;
    if (rx * dx + ry * dy >= 0.0 && (rx != 0.0 || ry != 0.0)) p2 = make_double2(X, Y);
    else k2 = -1;
  } else if (cnt > 1) {
    // nearest forward hit among the crossing halfedges, first wins ties (:4563)
    double best = 1.7976931348623157e308;
    int kb = -1;
    for (int k = k2, left = cnt;;) {
      const int kn = (k + 1 == n) ? 0 : k + 1;
      const double2 Vk = ldpt(T.pt, off + k), Vn = ldpt(T.pt, off + kn);
      const double sk = fma(a, Vk.x, fma(b, Vk.y, -w)), sn = fma(a, Vn.x, fma(b, Vn.y, -w));
      const double sg = sk / (sk - sn);
      const double X = fma(sg, Vn.x - Vk.x, Vk.x), Y = fma(sg, Vn.y - Vk.y, Vk.y);
      const double rx = X - p1.x, ry = Y - p1.y;
      if (rx * dx + ry * dy >= 0.0) {
        const double len = sqrt(rx * rx + ry * ry);
        if (len != 0.0 && len < best) { best = len; kb = k; p2 = make_double2(X, Y); }
      }
      if (--left == 0) break;
      for (;;) {  // next crossing halfedge after k
        k++;
        if (k == k1 || k == skip1) continue;
        const int kn2 = (k + 1 == n) ? 0 : k + 1;
        const double2 A_ = ldpt(T.pt, off + k), B_ = ldpt(T.pt, off + kn2);
        const double sa = fma(a, A_.x, fma(b, A_.y, -w)), sb = fma(a, B_.x, fma(b, B_.y, -w));
        if (sa * sb <= 0.0 && sa != sb) break;
      }
    }
    k2 = kb;
  }
  o.p1 = p1; o.p2 = p2;
  const int D = PG::d(G);
  const uint32_t M = PG::mask(G);
  if (k2 < 0 || !(sP != sQ)) {
    o.e2 = -1;
    if (WITH_STAT) for (int i = 0; i < D; i++) stat_col0[(size_t)i * cap] = 0.0;
    return;
  }
  o.e2 = __ldg(T.slot_he + off + k2);
  if (!WITH_STAT) return;

  const int s1 = off + k1;
  const int k2n = (k2 + 1 == n) ? 0 : k2 + 1;
  const uint32_t fl2 = __ldg(T.flg + off + k2);
  const bool bnd1 = S.fl1 & SF_BND, bnd2 = fl2 & SF_BND;
  const double d01x = p2.x - p1.x, d01y = p2.y - p1.y;

  // children (polygon_insert_edge, :4600-4631):
  //   C1 = [p1, p2, V(k2+1) .. V(k1)],  C2 = [p2, p1, V(k1+1) .. V(k2)]
  // The child with fewer parent vertices is walked from p1 (accurate however small
  // it is: a corner cut off by the split has relative area ~x^2); the other one is
  // the parent minus it, and its perimeter comes from the prefix sums.
  double A1 = 0, A2 = 0, d01 = 0, L1 = 0, L2 = 0;
  if (M & MASK_NEED_D01) d01 = sqrt(d01x * d01x + d01y * d01y);
  if (M & MASK_FACE_AREA) {
    const bool need_len = M & (1u << LITE_FEAT_FACE_SHAPE);
    int dk = k1 - k2n; if (dk < 0) dk += n;          // C1 holds dk+1 parent vertices, C2 the other n-dk-1
    const bool walk1 = 2 * (dk + 1) <= n;
    const double half_tot = S.half_tot;
    const double ltot = need_len ? __ldg(T.plen + off + n - 1) + __ldg(T.len + off + n - 1) : 0.0;
    double Aw, Lw = 0;
    {
      const int ks = walk1 ? k2n : k1n, ke = walk1 ? k1 : k2;
      double2 Vk = ldpt(T.pt, off + ks);
      double rx = Vk.x - p1.x, ry = Vk.y - p1.y;
      double acc = walk1 ? cross2(d01x, d01y, rx, ry) : 0.0;
      if (need_len) Lw = d01 + (walk1 ? plen(p2, Vk) : sqrt(rx * rx + ry * ry));
      for (int k = ks; k != ke;) {
        const int kn = (k + 1 == n) ? 0 : k + 1;
        const double2 Vn = ldpt(T.pt, off + kn);
        const double qx = Vn.x - p1.x, qy = Vn.y - p1.y;
        acc += cross2(rx, ry, qx, qy);
        if (need_len) Lw += __ldg(T.len + off + k);
        rx = qx; ry = qy; k = kn;
      }
      if (!walk1) acc += cross2(rx, ry, d01x, d01y);
      if (need_len) Lw += walk1 ? sqrt(rx * rx + ry * ry) : sqrt((rx - d01x) * (rx - d01x) + (ry - d01y) * (ry - d01y));
      Aw = 0.5 * acc;
    }
    double Ao = half_tot - Aw;  // the two children tile the parent
    if (Ao < 1e-6 * half_tot) {
      // the many-vertex child is the sliver (micro-edges): walk it too
      const int ks = walk1 ? k1n : k2n, ke = walk1 ? k2 : k1;
      double2 Vk = ldpt(T.pt, off + ks);
      double rx = Vk.x - p1.x, ry = Vk.y - p1.y;
      double acc = walk1 ? 0.0 : cross2(d01x, d01y, rx, ry);
      for (int k = ks; k != ke;) {
        const int kn = (k + 1 == n) ? 0 : k + 1;
        const double2 Vn = ldpt(T.pt, off + kn);
        const double qx = Vn.x - p1.x, qy = Vn.y - p1.y;
        acc += cross2(rx, ry, qx, qy);
        rx = qx; ry = qy; k = kn;
      }
      if (walk1) acc += cross2(rx, ry, d01x, d01y);
      Ao = 0.5 * acc;
    }
    double Lo = 0;
    if (need_len) {
      const double2 S2 = ldpt(T.pt, off + k2), E2 = ldpt(T.pt, off + k2n);
      Lo = walk1 ? d01 + plen(p1, Q) + chain(T.plen, ltot, off, k1n, k2) + plen(S2, p2)
                 : d01 + plen(p2, E2) + chain(T.plen, ltot, off, k2n, k1) + plen(P, p1);
    }
    A1 = walk1 ? Aw : Ao; A2 = walk1 ? Ao : Aw;
    L1 = walk1 ? Lw : Lo; L2 = walk1 ? Lo : Lw;
  }
  // new corner angles: alpha at p1 (C2 side), beta at p2 (C1 side)
  const double alpha = angle;
  double beta = 0;
  if (M & MASK_FACE_ANG) beta = fabs(wrap_pi(__ldg(T.ang + off + k2) - l - LITE_PI));

  // Every feature value is computed under a branch on the (uniform) program mask,
  // so that a program pays only for the features it registers; the final loop just
  // routes the values to their columns.
  double v_edge_b = 0, v_edge_i = 0, v_inside = 0, v_segsize = 0, v_shape = 0, v_minang = 0;
  if (M & (1u << LITE_FEAT_IS_POINT_INSIDE_WINDOW)) {
    // add_vertices = {p1,p2} (:2143-2144); a point on an edge is strictly inside the
    // window iff the edge is internal
    bool in1 = !bnd1;
    if (x == 0.0) {  // p1 == source vertex of e1
      int kp = (k1 == 0) ? n - 1 : k1 - 1;
      in1 = !bnd1 && !(T.flg[off + kp] & SF_BND);
    }
    v_inside = (in1 ? 1.0 : 0.0) + (bnd2 ? 0.0 : 1.0);
  }
  if (M & ((1u << LITE_FEAT_EDGE_LENGTH) | (1u << LITE_FEAT_EDGE_LENGTH_INTERNAL))) {
    // -{e1,e2} +{(s1,p1),(p1,t1),(s2,p2),(p2,t2),(p1,p2)} (:2148-2164): the two halves of
    // a cut edge minus the edge, as the reference sums them (rounding noise, not 0).
    // edge_length as written only sees window-boundary edges: a rare branch.
    const bool wi = M & (1u << LITE_FEAT_EDGE_LENGTH_INTERNAL);
    double c1 = 0, c2 = 0;
    if (bnd1 || wi) c1 = edge_cut_delta(P.x, P.y, p1.x, p1.y, Q.x, Q.y);
    if (bnd2 || wi) {
      const double2 S2 = ldpt(T.pt, off + k2), E2 = ldpt(T.pt, off + k2n);
      c2 = edge_cut_delta(S2.x, S2.y, p2.x, p2.y, E2.x, E2.y);
    }
    v_edge_b = (bnd1 ? c1 : 0.0) + (bnd2 ? c2 : 0.0);
    v_edge_i = (bnd1 ? 0.0 : c1) + (bnd2 ? 0.0 : c2) + d01;
  }
  if (M & (1u << LITE_FEAT_SEGMENT_SIZE_2)) {
    // boundary segments gain one edge: (k+1)^2-k^2 (:2202-2224, :4424-4434)
    if (bnd1) v_segsize += 2.0 * T.seg_n_edges[T.slot_seg[s1]] + 1.0;
    if (bnd2) v_segsize += 2.0 * T.seg_n_edges[T.slot_seg[off + k2]] + 1.0;
  }
  if (M & (1u << LITE_FEAT_FACE_SHAPE)) {
    const int f = T.slot_face[s1];
    v_shape = shape_of(L1, A1) + shape_of(L2, A2) - shape_of(T.f_perim[f], T.f_area[f]);
  }
  if (M & (1u << LITE_FEAT_MIN_ANGLE)) {
    const int f = T.slot_face[s1];
    const double m2 = corner_min(T.flg, T.ca, off, n, k1n, k2);  // parent corners inherited by C2
    const double m1 = corner_min(T.flg, T.ca, off, n, k2n, k1);  // ... by C1
    // C1 = [p1,p2,...]: seed pi-(pi-alpha); corners pi-alpha (p1), beta (p2)
    double ma1 = fmin(fmin(alpha, LITE_PI - alpha), fmin(beta, m1));
    // C2 = [p2,p1,...]: seed pi-(pi-beta); corners pi-beta (p2), alpha (p1)
    double ma2 = fmin(fmin(beta, LITE_PI - beta), fmin(alpha, m2));
    v_minang = ma1 + ma2 - T.f_minang[f];
  }
#pragma unroll
  for (int i = 0; i < (PG::kStatic ? PG::d(G) : LITE_MAX_D); i++) {
    if (!PG::kStatic && i >= D) break;
    double dv = 0.0;  // phi(after) - phi(before)
    switch (PG::fid(G, i)) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW: dv = v_inside; break;
      case LITE_FEAT_EDGE_LENGTH: dv = v_edge_b; break;
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: dv = v_edge_i; break;
      case LITE_FEAT_SEG_NUMBER: dv = 1.0; break;
      case LITE_FEAT_IS_SEGMENT_INTERNAL: dv = 1.0; break;
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: dv = -1.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2: dv = v_segsize; break;
      case LITE_FEAT_FACE_NUMBER: dv = 1.0; break;
      case LITE_FEAT_FACE_AREA_2: dv = -2.0 * A1 * A2; break;  // a1^2+a2^2-(a1+a2)^2
      case LITE_FEAT_FACE_PERIMETER: dv = 2.0 * d01; break;
      case LITE_FEAT_FACE_SHAPE: dv = v_shape; break;
      case LITE_FEAT_FACE_SUM_OF_ANGLES:
        // untouched corners cancel; new corners alpha, pi-alpha at p1 and beta, pi-beta at p2
        dv = fabs(LITE_HALF_PI - alpha) + fabs(LITE_HALF_PI - beta);
        break;
      case LITE_FEAT_MIN_ANGLE: dv = v_minang; break;
      default: break;
    }
    stat_col0[(size_t)i * cap] = -dv;
  }
}

struct SplitArgs {
  long n_local;      // candidates handled by this launch
  long j0;           // index of the first one inside the batch
  double sel_step;   // total_len / batch size: width of one stratum of the sampler
  uint64_t seed, ctr0;  // philox: counter of batch element j is ctr0 + j
  const int *in_he; const double *in_x, *in_angle;  // given path
  double *stat; size_t cap; long base;              // column k at stat[k*cap + base + i]
  // optional record
  int *r_e1, *r_e2; double2 *r_p1, *r_p2; double *r_x, *r_angle, *r_u;
  unsigned long long *status;  // [0] no-exit count
  unsigned long long *checksum;  // [0] xor hash, [1] sum of e2
  unsigned long long *ticket;  // [0] run scheduler of this launch, [1] warps that have left; both zero between launches
  // i.i.d. sampler (LITE_SAMPLER_IID): positions along the cumulative halfedge length drawn
  // independently and sorted (split_sample, lib/ttessel.cpp:1968-1975), with the batch index of
  // each draw; null = the stratified sampler
  const double *sel_sorted; const long long *sel_id;
};

// Stratified length-weighted halfedge of candidate j (the device form of
// alt_length_weighted_random_halfedge(n), lib/ttessel.cpp:1968-1991): slot m with
// cum[m-1] <= sel < cum[m] (last slot if sel is beyond), found from a guide table
// in O(1) expected loads.
__device__ __forceinline__ int sample_slot(const TessDev &T, double sel) {
  int b = (int)(sel * T.guide_scale);
  b = min(max(b, 0), T.n_guide - 1);
  int m = __ldg(T.guide + b);
  while (m > 0 && __ldg(T.cum + m - 1) > sel) m--;
  while (m < T.n_cum - 1 && __ldg(T.cum + m) <= sel) m++;
  return m;
}

// i.i.d. positions of a batch shard: sel_j = U_j * total_len with U_j the 53-bit uniform of
// Philox stream 1 at counter ctr0 + j (the candidate's x and angle come from stream 0)
#define LITE_IID_STREAM 1u
__global__ void k_iid_draw(long n_local, long j0, uint64_t seed, uint64_t ctr0, double total_len, double *sel,
                           long long *id) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  uint32_t r[4];
  philox4x32_10(ctr0 + (uint64_t)(j0 + i), LITE_IID_STREAM, seed, r);
  sel[i] = u53(r[0], r[1]) * total_len;
  id[i] = j0 + i;
}

#ifndef LITE_SPLIT_TB
#define LITE_SPLIT_TB 128
#endif
#ifndef LITE_SPLIT_MINB
#define LITE_SPLIT_MINB 6
#endif
// MODE bit0: philox sampling (else given list); bit1: record; bit2: statistics; bit3: checksum
template <int MODE, class PG>
__global__ void __launch_bounds__(LITE_SPLIT_TB, LITE_SPLIT_MINB) k_split(TessDev T, Prog G, SplitArgs A) {
  constexpr bool PHILOX = MODE & 1, RECORD = MODE & 2, STAT = MODE & 4, CHECK = MODE & 8;
  if (*T.err) return;  // the tessellation failed k_validate
  // Guided self-scheduling: a warp takes contiguous runs of candidates (32 per
  // iteration) from a ticket counter; run lengths halve from n/(2*warps) down to 128,
  // so the slot cache pays off on the long early runs and the short late ones even
  // out the finish whatever the residency of the blocks.
  const int lane = threadIdx.x & 31;
  const long n_warps = ((long)gridDim.x * blockDim.x) >> 5;
  long run0 = (A.n_local + 2 * n_warps - 1) / (2 * n_warps);
  run0 = max((run0 + 31) & ~31L, 128L);
  unsigned long long cx = 0, cs = 0;
  unsigned int bad = 0;
  int m = -1;          // current slot of this lane (philox path)
  double cum_hi = 0;   // cum[m]
  SlotData S;
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(A.ticket, 1ULL);
    t = __shfl_sync(0xffffffffu, t, 0);
    long beg = 0, len = run0;
    while (t >= (unsigned long long)n_warps && len > 128) {  // skip whole levels of the schedule
      beg += n_warps * len; t -= n_warps; len = max(((len >> 1) + 31) & ~31L, 128L);
    }
    beg += (long)t * len;
    if (beg >= A.n_local) {
      // every warp leaves here exactly once; the last one re-arms the scheduler for the next launch
      if (lane == 0 && atomicAdd(A.ticket + 1, 1ULL) == (unsigned long long)(n_warps - 1)) { A.ticket[0] = 0ULL; A.ticket[1] = 0ULL; }
      break;
    }
    const long end = min(beg + len, A.n_local);
    m = -1;
  // (a contiguous piece of the run per lane instead of this interleaving keeps a lane on its slot
  // longer but scatters the statistic stores: measured 2x slower at 10^4 and at 10^5 cells)
  for (long i = beg + lane; i < end; i += 32) {
    double x, angle, ca, sa, U = 0.0;
    if (PHILOX) {
      const bool iid = A.sel_sorted != nullptr;
      const long j = iid ? (long)A.sel_id[i] : A.j0 + i;
      uint32_t r[4];
      philox4x32_10(A.ctr0 + (uint64_t)j, 0u, A.seed, r);
      x = u53(r[0], r[1]);
      U = u53(r[2], r[3]);
      // the 22 bits u53 leaves unused place the candidate inside its stratum
      const uint32_t spare = ((r[0] & 31u) << 17) | ((r[1] & 63u) << 11) | ((r[2] & 31u) << 6) | (r[3] & 63u);
      const double sel = iid ? A.sel_sorted[i] : ((double)j + (double)spare * (1.0 / 4194304.0)) * A.sel_step;
      // sel grows along the lane's run (stratified by construction, i.i.d. after the sort), so
      // the slot only moves forward
      int m_new = m;
      if (m < 0) m_new = sample_slot(T, sel);
      else while (m_new < T.n_cum - 1 && sel >= cum_hi) cum_hi = __ldg(T.cum + ++m_new);
      if (m_new != m) { m = m_new; cum_hi = __ldg(T.cum + m); S = load_slot(T, m); }
      // angle = acos(1 - 2U) (lib/ttessel.cpp:1792): cos(angle) is the draw, sin(angle) >= 0
      ca = 1.0 - 2.0 * U;
      sa = 2.0 * sqrt(U * (1.0 - U));
      const bool need_angle = RECORD || (STAT && (PG::mask(G) & MASK_FACE_ANG));
#ifdef LITE_LIBM_ACOS
      angle = need_angle ? acos(ca) : 0.0;
#else
      angle = need_angle ? acos_cs(ca, sa) : 0.0;
#endif
    } else {
      S = load_slot(T, __ldg(T.he_slot + A.in_he[i])); x = A.in_x[i]; angle = A.in_angle[i];
      sincos(angle, &sa, &ca);
    }
    if (S.fl1 & SF_GENERAL) continue;  // k_gen_splits evaluates it (gen_face.h)
    SplitOut o;
    split_eval<STAT, PG>(T, G, S, x, angle, ca, sa, A.stat + A.base + i, A.cap, o);
    if (o.e2 < 0) bad++;
    if (RECORD || CHECK) {
      const int h = __ldg(T.slot_he + S.off + S.k1);
      if (RECORD) {
        A.r_e1[A.base + i] = h; A.r_e2[A.base + i] = o.e2;
        A.r_p1[A.base + i] = o.p1; A.r_p2[A.base + i] = o.p2;
        A.r_x[A.base + i] = x; A.r_angle[A.base + i] = angle; A.r_u[A.base + i] = U;
      }
      if (CHECK) {
        unsigned long long key = ((unsigned long long)(unsigned)h << 32) | (unsigned)o.e2;
        key *= 0x9E3779B97F4A7C15ull; key ^= key >> 29;
        cx ^= key; cs += (unsigned long long)(unsigned)o.e2;
      }
    }
  }
  }
  if (CHECK) {
    for (int off = 16; off; off >>= 1) {
      cx ^= __shfl_xor_sync(0xffffffffu, cx, off);
      cs += __shfl_xor_sync(0xffffffffu, cs, off);
    }
    if (lane == 0) { atomicXor(A.checksum, cx); atomicAdd(A.checksum + 1, cs); }
  }
  if (bad) atomicAdd(A.status, (unsigned long long)bad);
}

// ---------------------------------------------------------------------------
// K4: merges  (Merge::modified_elements, lib/ttessel.cpp:2302-2404)
// ---------------------------------------------------------------------------

__device__ __forceinline__ void merge_one(const TessDev &T, const Prog &G, const int *nb_seg, int i,
                                          double *stat, size_t cap, int *out_he) {
  int sg = nb_seg[i];
  int e = T.seg_he_start[sg];  // single edge, dir == true: Seg::get_halfedge_handle()
  int t = T.he_twin[e];
  out_he[i] = e;
  int se = T.he_slot[e], st_ = T.he_slot[t];
  if ((se | st_) < 0) {  // not an internal segment: both sides must bound inside faces
    for (int k = 0; k < G.d; k++) stat[(size_t)k * cap + i] = 0.0;
    return;
  }
  int f1 = T.slot_face[se], f2 = T.slot_face[st_];
  if (T.f_gen && (T.f_gen[f1] | T.f_gen[f2])) return;  // k_gen_mf has written this row (gen_face.h)
  int o1 = T.f_off[f1], n1 = T.f_cnt[f1], o2 = T.f_off[f2], n2 = T.f_cnt[f2];
  int ke = se - o1, kt = st_ - o2;
  int ke1 = (ke + 1) % n1, kt1 = (kt + 1) % n2;  // e2() = next(e), e1() = next(twin e)
  double2 p1 = ldpt(T.pt, se), p2 = ldpt(T.pt, st_);
  uint32_t fe2 = T.flg[o1 + ke1], fe1 = T.flg[o2 + kt1];
  bool bnd_e2 = fe2 & SF_BND, bnd_e1 = fe1 & SF_BND;
  // neighbours along the blocking lines: X = tgt(next), Y = src(prev_hf(next))
  double2 X2 = ldpt(T.pt, o1 + (ke1 + 1) % n1), X1 = ldpt(T.pt, o2 + (kt1 + 1) % n2);
  // Y2 is the vertex before p2 in face(twin e), Y1 the vertex before p1 in face(e)
  double2 Y2 = ldpt(T.pt, o2 + (kt + n2 - 1) % n2), Y1 = ldpt(T.pt, o1 + (ke + n1 - 1) % n1);
  double elen = plen(p1, p2);

  double acc[LITE_MAX_D];
  for (int k = 0; k < G.d; k++) acc[k] = 0.0;
  if ((G.mask & MASK_FACE_ANY) && !(G.mask & (1u << LITE_FEAT_MIN_ANGLE))) {
    // Closed forms from the parents' precomputed features.  The merged face is
    // face(e) + face(twin e) with p1, p2 flat (:4945-4966): areas add, the perimeter
    // loses e twice, every other corner keeps its angle, so only the four corners at
    // p1 and p2 leave the sum of angles.
    const FaceFeat F1 = parent_feat(T, f1), F2 = parent_feat(T, f2);
    double gone = 0;
    if (G.mask & (1u << LITE_FEAT_FACE_SUM_OF_ANGLES)) {
      const int cs[4] = {o1 + ke, o1 + ke1, o2 + kt, o2 + kt1};
#pragma unroll
      for (int c = 0; c < 4; c++)
        if (T.flg[cs[c]] & SF_CORNER) gone += fmax(0.0, LITE_HALF_PI - T.ca[cs[c]]);
    }
    for (int k = 0; k < G.d; k++) {
      switch (G.fid[k]) {
        case LITE_FEAT_FACE_NUMBER: acc[k] = -1.0; break;
        case LITE_FEAT_FACE_AREA_2: acc[k] = 2.0 * F1.area * F2.area; break;  // (a1+a2)^2-a1^2-a2^2
        case LITE_FEAT_FACE_PERIMETER: acc[k] = -2.0 * elen; break;
        case LITE_FEAT_FACE_SHAPE:
          acc[k] = shape_of(F1.perim + F2.perim - 2.0 * elen, F1.area + F2.area) - shape_of(F1.perim, F1.area) -
                   shape_of(F2.perim, F2.area);
          break;
        case LITE_FEAT_FACE_SUM_OF_ANGLES: acc[k] = -gone; break;
        default: break;
      }
    }
  } else if (G.mask & MASK_FACE_ANY) {
    bool ang = G.mask & MASK_FACE_ANG;
    VPoly P;
    FaceFeat F;
    // del: simplify(face(e)), simplify(face(twin e)) in face2poly order (:2333-2362)
    add_face_feats(G, acc, -1.0, parent_feat(T, f1));
    add_face_feats(G, acc, -1.0, parent_feat(T, f2));
    // add: polygon_remove_edge (:4945-4966): face(e) after p2 .. p1, face(twin) after p1 .. p2
    vp_init(P);
    vp_range(P, o1, n1, ke + 2, n1 - 2);
    vp_point(P, p1.x, p1.y, 0);
    vp_range(P, o2, n2, kt + 2, n2 - 2);
    vp_point(P, p2.x, p2.y, 0);
    F = vp_features(T, P, ang); add_face_feats(G, acc, +1.0, F);
  }
  for (int k = 0; k < G.d; k++) {
    double dv = 0.0;
    switch (G.fid[k]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW:
        dv = -((bnd_e1 ? 0.0 : 1.0) + (bnd_e2 ? 0.0 : 1.0));
        break;
      case LITE_FEAT_EDGE_LENGTH:
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: {
        bool wantb = (G.fid[k] == LITE_FEAT_EDGE_LENGTH);
        if (!wantb) dv -= elen;
        if (bnd_e2 == wantb) dv += plen(Y2, X2) - plen(p2, X2) - plen(Y2, p2);
        if (bnd_e1 == wantb) dv += plen(Y1, X1) - plen(p1, X1) - plen(Y1, p1);
      } break;
      case LITE_FEAT_SEG_NUMBER: dv = -1.0; break;
      case LITE_FEAT_IS_SEGMENT_INTERNAL: dv = -1.0; break;
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: dv = 1.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2: {
        if (bnd_e1) dv -= 2.0 * T.seg_n_edges[T.slot_seg[o2 + kt1]] - 1.0;
        if (bnd_e2) dv -= 2.0 * T.seg_n_edges[T.slot_seg[o1 + ke1]] - 1.0;
      } break;
      default: dv = acc[k]; break;
    }
    stat[(size_t)k * cap + i] = -dv;
  }
}

// ---------------------------------------------------------------------------
// K5: flips  (Flip ctor :2432-2454, Flip::modified_elements :2459-2713)
// flip i < n_b : e1 = twin(halfedges_start(seg)),  i >= n_b: e1 = halfedges_end(seg)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void flip_one(const TessDev &T, const Prog &G, const int *b_seg, int nb, int i,
                                         double *stat, size_t cap, int *out_e1, int *out_e2,
                                         double2 *out_p2, int *out_right, unsigned long long *status) {
  int sg = b_seg[i < nb ? i : i - nb];
  int e1 = (i < nb) ? T.he_twin[T.seg_he_start[sg]] : T.seg_he_end[sg];
  int t1 = T.he_twin[e1];
  int s1 = T.he_slot[e1], st1 = T.he_slot[t1];
  if ((s1 | st1) < 0) {  // not an internal segment
    out_e1[i] = e1; out_e2[i] = -1; out_right[i] = 0; out_p2[i] = make_double2(0, 0);
    atomicAdd(status, 1ull);
    for (int k = 0; k < G.d; k++) stat[(size_t)k * cap + i] = 0.0;
    return;
  }
  int fL = T.slot_face[s1], fT = T.slot_face[st1];  // face(e1), face(twin e1)
  if (T.f_gen && (T.f_gen[fL] | T.f_gen[fT])) return;  // k_gen_mf has written this row and its outputs
  int oL = T.f_off[fL], nL = T.f_cnt[fL], oT = T.f_off[fT], nT = T.f_cnt[fT];
  int k1 = s1 - oL, kt = st1 - oT;
  // e3: incoming halfedge at a = src(e1) that is not on e1's segment (:2435-2438)
  int kp = (k1 + nL - 1) % nL;  // prev(e1) in face(e1), arrives at a
  int hp = T.slot_he[oL + kp];
  bool right = (hp != T.he_prev_hf[e1]);  // prev(e1) is e3: it comes from the left of e1
  double2 a = ldpt(T.pt, s1), q = ldpt(T.pt, st1);
  int oR, nR, ka;  // face the ray enters; ka = slot index (in R) whose source is a
  double ang3;
  if (right) {
    ang3 = T.ang[oL + kp];
    oR = oT; nR = nT; ka = (kt + 1) % nT;
  } else {
    // e3 = twin(next(twin e1)): reverse of the slot leaving a in face(twin e1)
    ang3 = T.ang[oT + (kt + 1) % nT] + LITE_PI;
    oR = oL; nR = nL; ka = k1;
  }
  double dx, dy;
  sincos(ang3, &dy, &dx);
  if (right) {
    double2 s3 = ldpt(T.pt, oL + kp);  // exact direction of e3 from its endpoints
    dx = a.x - s3.x; dy = a.y - s3.y;
  } else {
    double2 s3 = ldpt(T.pt, oT + (kt + 2) % nT);
    dx = a.x - s3.x; dy = a.y - s3.y;
  }
  { double nrm = sqrt(dx * dx + dy * dy); dx /= nrm; dy /= nrm; }
  double la = dy, lb = -dx, lw = la * a.x + lb * a.y;  // line through a along d
  double2 p2 = make_double2(0, 0);
  int k2 = ray_exit(T, oR, nR, la, lb, lw, a, dx, dy, ka, (ka + nR - 1) % nR, p2);
  out_e1[i] = e1; out_right[i] = right ? 1 : 0; out_p2[i] = p2;
  if (k2 < 0) {
    out_e2[i] = -1;
    atomicAdd(status, 1ull);
    for (int k = 0; k < G.d; k++) stat[(size_t)k * cap + i] = 0.0;
    return;
  }
  int e2 = T.slot_he[oR + k2];
  out_e2[i] = e2;
  int k2n = (k2 + 1) % nR;
  // next(e1): leaves q along the blocking line; X = its target, Y = the vertex
  // before q on the other side (source of prev_hf(next(e1)))
  int kn1 = (k1 + 1) % nL;
  bool bnd_n1 = T.flg[oL + kn1] & SF_BND, bnd_e2 = T.flg[oR + k2] & SF_BND;
  double2 X = ldpt(T.pt, oL + (kn1 + 1) % nL), Y = ldpt(T.pt, oT + (kt + nT - 1) % nT);
  double2 S2 = ldpt(T.pt, oR + k2), E2 = ldpt(T.pt, oR + k2n);
  int vq = T.he_src[t1];
  bool tgt_is_q = T.he_src[T.he_twin[e2]] == vq, src_is_q = T.he_src[e2] == vq;

  double acc[LITE_MAX_D];
  for (int k = 0; k < G.d; k++) acc[k] = 0.0;
  const double l_e1 = plen(a, q), l_new = plen(a, p2);
  if ((G.mask & MASK_FACE_ANY) && !(G.mask & (1u << LITE_FEAT_MIN_ANGLE))) {
    // Closed forms.  R = the face the ray enters, L = the other face
    // of e1.  The chord (a,p2) cuts R into R1 (the part bounded by e1) and R2; the
    // flip gives -{R, L} +{R2, R1 u L} (:2515-2655).  Areas and perimeters follow
    // from those of R, L (precomputed) and of ONE of the two parts, walked from a
    // (the one with fewer vertices).  In the sum of angles only three vertices move:
    // a keeps its corner (it passes from L to R2), q loses its two corners delta and
    // pi-delta, p2 gains beta and pi-beta.
    const int fR = right ? fT : fL, fO = right ? fL : fT;
    const FaceFeat FR = parent_feat(T, fR), FO = parent_feat(T, fO);
    const bool need_len = G.mask & (1u << LITE_FEAT_FACE_SHAPE);
    double r1 = 0, pr1 = 0;  // area and perimeter of R1
    if (G.mask & MASK_FACE_AREA) {
      // left part  [a, p2, V(k2n) .. V(ka-1)]  and right part  [a, V(ka+1) .. V(k2), p2]
      int cl = ka - 1 - k2n; cl = ((cl % nR) + nR) % nR + 1;
      const bool walk_left = 2 * cl <= nR - 1;
      const double px = p2.x - a.x, py = p2.y - a.y;
      double Aw, Lw = 0;
      {
        const int ks = walk_left ? k2n : (ka + 1) % nR, ke_ = walk_left ? (ka + nR - 1) % nR : k2;
        double2 Vk = ldpt(T.pt, oR + ks);
        double rx = Vk.x - a.x, ry = Vk.y - a.y;
        double acc2 = walk_left ? cross2(px, py, rx, ry) : 0.0;
        if (need_len) Lw = l_new + (walk_left ? plen(p2, Vk) : T.len[oR + ka]);
        for (int k = ks; k != ke_;) {
          const int kn = (k + 1 == nR) ? 0 : k + 1;
          const double2 Vn = ldpt(T.pt, oR + kn);
          const double qx = Vn.x - a.x, qy = Vn.y - a.y;
          acc2 += cross2(rx, ry, qx, qy);
          if (need_len) Lw += T.len[oR + k];
          rx = qx; ry = qy; k = kn;
        }
        if (!walk_left) acc2 += cross2(rx, ry, px, py);
        if (need_len) Lw += walk_left ? T.len[oR + ke_] : sqrt((rx - px) * (rx - px) + (ry - py) * (ry - py));
        Aw = 0.5 * acc2;
      }
      double Ao = FR.area - Aw;
      if (Ao < 1e-6 * FR.area) {  // the unwalked part is a sliver: walk it as well
        const int ks = walk_left ? (ka + 1) % nR : k2n, ke_ = walk_left ? k2 : (ka + nR - 1) % nR;
        double2 Vk = ldpt(T.pt, oR + ks);
        double rx = Vk.x - a.x, ry = Vk.y - a.y;
        double acc2 = walk_left ? 0.0 : cross2(px, py, rx, ry);
        for (int k = ks; k != ke_;) {
          const int kn = (k + 1 == nR) ? 0 : k + 1;
          const double2 Vn = ldpt(T.pt, oR + kn);
          const double qx = Vn.x - a.x, qy = Vn.y - a.y;
          acc2 += cross2(rx, ry, qx, qy);
          rx = qx; ry = qy; k = kn;
        }
        if (walk_left) acc2 += cross2(rx, ry, px, py);
        Ao = 0.5 * acc2;
      }
      const double Lo = FR.perim + 2.0 * l_new - Lw;  // the two parts share the chord
      // R1 is the left part when the ray went to the right of e1 (it holds twin(e1): q -> a)
      const bool r1_is_walked = (walk_left == right);
      r1 = r1_is_walked ? Aw : Ao;
      pr1 = r1_is_walked ? Lw : Lo;
    }
    double d_sumang = 0;
    if (G.mask & (1u << LITE_FEAT_FACE_SUM_OF_ANGLES)) {
      const int sq = right ? oL + (k1 + 1) % nL : oT + kt;  // slot of L whose source is q
      const double delta = (T.flg[sq] & SF_CORNER) ? T.ca[sq] : LITE_PI;
      const double beta = fabs(wrap_pi(T.ang[oR + k2] - ang3 - LITE_PI));
      d_sumang = fabs(LITE_HALF_PI - beta) - fabs(LITE_HALF_PI - delta);
    }
    for (int k = 0; k < G.d; k++) {
      switch (G.fid[k]) {
        case LITE_FEAT_FACE_NUMBER: acc[k] = 0.0; break;
        // a(R2)^2 + a(R1 u L)^2 - a(R)^2 - a(L)^2 with a(R2) = a(R) - r1, a(R1 u L) = a(L) + r1
        case LITE_FEAT_FACE_AREA_2: acc[k] = 2.0 * r1 * (FO.area - FR.area + r1); break;
        case LITE_FEAT_FACE_PERIMETER: acc[k] = 2.0 * (l_new - l_e1); break;
        case LITE_FEAT_FACE_SHAPE:
          acc[k] = shape_of(FR.perim + 2.0 * l_new - pr1, FR.area - r1) + shape_of(FO.perim + pr1 - 2.0 * l_e1, FO.area + r1) -
                   shape_of(FR.perim, FR.area) - shape_of(FO.perim, FO.area);
          break;
        case LITE_FEAT_FACE_SUM_OF_ANGLES: acc[k] = d_sumang; break;
        default: break;
      }
    }
  } else if (G.mask & MASK_FACE_ANY) {
    bool ang = G.mask & MASK_FACE_ANG;
    VPoly P; FaceFeat F;
    // deleted faces: the face of e2 first, then the other face bounded by e1 (:2644-2651)
    if (right) {
      add_face_feats(G, acc, -1.0, parent_feat(T, fT));
      add_face_feats(G, acc, -1.0, parent_feat(T, fL));
      // unmatched inserted face: outer2 = [p2, a, V(kt+2) .. V(k2)]
      vp_init(P);
      vp_point(P, p2.x, p2.y, 1); vp_point(P, a.x, a.y, 1);
      vp_range(P, oT, nT, kt + 2, ((k2 - (kt + 1)) % nT + nT) % nT);
      F = vp_features(T, P, ang); add_face_feats(G, acc, +1.0, F);
      // merged face: L after q .. a(flat), p2, R V(k2+1) .. V(kt-1), q(flat)
      vp_init(P);
      vp_range(P, oL, nL, k1 + 2, nL - 2);
      vp_point(P, a.x, a.y, 0); vp_point(P, p2.x, p2.y, 1);
      vp_range(P, oT, nT, k2 + 1, ((kt - (k2 + 1)) % nT + nT) % nT);
      vp_point(P, q.x, q.y, 0);
      F = vp_features(T, P, ang); add_face_feats(G, acc, +1.0, F);
    } else {
      add_face_feats(G, acc, -1.0, parent_feat(T, fL));
      add_face_feats(G, acc, -1.0, parent_feat(T, fT));
      // unmatched inserted face: outer1 = [a, p2, V(k2+1) .. V(k1-1)]
      vp_init(P);
      vp_point(P, a.x, a.y, 1); vp_point(P, p2.x, p2.y, 1);
      vp_range(P, oL, nL, k2 + 1, ((k1 - (k2 + 1)) % nL + nL) % nL);
      F = vp_features(T, P, ang); add_face_feats(G, acc, +1.0, F);
      // merged face: R V(k1+2) .. V(k2), p2, a(flat), L' after a .. before q, q(flat)
      vp_init(P);
      vp_range(P, oL, nL, k1 + 2, ((k2 - (k1 + 1)) % nL + nL) % nL);
      vp_point(P, p2.x, p2.y, 1); vp_point(P, a.x, a.y, 0);
      vp_range(P, oT, nT, kt + 2, nT - 2);
      vp_point(P, q.x, q.y, 0);
      F = vp_features(T, P, ang); add_face_feats(G, acc, +1.0, F);
    }
  }
  for (int k = 0; k < G.d; k++) {
    double dv = 0.0;
    switch (G.fid[k]) {
      case LITE_FEAT_IS_POINT_INSIDE_WINDOW:
        dv = (bnd_e2 ? 0.0 : 1.0) - (bnd_n1 ? 0.0 : 1.0);
        break;
      case LITE_FEAT_EDGE_LENGTH:
      case LITE_FEAT_EDGE_LENGTH_INTERNAL: {
        bool wantb = (G.fid[k] == LITE_FEAT_EDGE_LENGTH);
        if (!wantb) dv += l_new - l_e1;
        // -(q,X) -(q,Y) (:2476-2486)
        if (bnd_n1 == wantb) dv -= plen(q, X) + plen(q, Y);
        if (!tgt_is_q && !src_is_q) {
          if (bnd_e2 == wantb) dv += plen(p2, E2) + plen(p2, S2) - plen(E2, S2);
          if (bnd_n1 == wantb) dv += plen(X, Y);
        } else if (tgt_is_q) {
          if (bnd_e2 == wantb) dv += plen(p2, S2) + plen(p2, X);
        } else {
          if (bnd_e2 == wantb) dv += plen(p2, E2) + plen(p2, Y);
        }
      } break;
      case LITE_FEAT_SEG_NUMBER:
      case LITE_FEAT_IS_SEGMENT_INTERNAL:
      case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: dv = 0.0; break;
      case LITE_FEAT_SEGMENT_SIZE_2: {
        int sg2 = T.slot_seg[oR + k2], sgn = T.slot_seg[oL + kn1];
        if (sg2 != sgn) {  // else: one segment, split and merged at once (:2696-2702)
          if (bnd_e2) dv += 2.0 * T.seg_n_edges[sg2] + 1.0;
          if (bnd_n1) dv -= 2.0 * T.seg_n_edges[sgn] - 1.0;
        }
      } break;
      default: dv = acc[k]; break;
    }
    stat[(size_t)k * cap + i] = -dv;
  }
}

// all merges and all flips in one launch: blocks [0, gm) take the merges, the rest the flips
struct MFArgs {
  const int *nb_seg; int nm; double *mstat; int *m_he; int gm;
  const int *b_seg; int nb; double *fstat; int *f_e1, *f_e2, *f_right; double2 *f_p2;
  unsigned long long *status;
  double *part;  // [gridDim.x][LITE_MAX_D]: per-block column sums of the statistics just written
};
// Each block also leaves the column sums of its 128 statistics rows (fixed tree): the host
// adds the blocks' partials in block order to get sum_stat_merges / sum_stat_flips
// (lib/ttessel.cpp:3337, :3345).  A separate 1024-thread summing kernel found no room
// beside k_split, which fills every SM, and finished only after it (scripts/step_variants.py).
__global__ void __launch_bounds__(128) k_merges_flips(TessDev T, Prog G, MFArgs A) {
  __shared__ double sh[128];
  if (*T.err) return;  // the tessellation failed k_validate (block-uniform)
  const bool merges = (int)blockIdx.x < A.gm;
  const int i = (merges ? blockIdx.x : blockIdx.x - A.gm) * blockDim.x + threadIdx.x;
  const int n = merges ? A.nm : 2 * A.nb;
  const double *stat = merges ? A.mstat : A.fstat;
  if (i < n) {
    if (merges) merge_one(T, G, A.nb_seg, i, A.mstat, (size_t)A.nm, A.m_he);
    else flip_one(T, G, A.b_seg, A.nb, i, A.fstat, (size_t)(2 * A.nb), A.f_e1, A.f_e2, A.f_p2, A.f_right, A.status);
  }
  for (int k = 0; k < G.d; k++) {
    sh[threadIdx.x] = (i < n) ? stat[(size_t)k * n + i] : 0.0;
    __syncthreads();
    for (int w = 64; w; w >>= 1) {
      if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) A.part[(size_t)blockIdx.x * LITE_MAX_D + k] = sh[0];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K3: theta-reduce.  One pass over the stored statistics gives
//   S0 = sum e_i, S1_k = sum e_i p_ik, S2_kl = sum e_i p_ik p_il (k<=l),
//   e_i = exp(theta . p_i)      (lib/ttessel.cpp:3392-3394, :3417-3419, :3435-3436)
// warp shuffle -> block tree -> fixed-order finish by the last block.
// ---------------------------------------------------------------------------
#ifndef LITE_REDUCE_LD
#define LITE_REDUCE_LD 0
#endif
__device__ __forceinline__ double2 ld_stat2(const double2 *p) {
#if LITE_REDUCE_LD == 0
  return __ldcs(p);   // streaming: evict first
#elif LITE_REDUCE_LD == 1
  return __ldg(p);    // normal caching: the tail of a pass can stay in L2 for the next one
#else
  return __ldlu(p);
#endif
}
template <int D, int NT>
__device__ __forceinline__ void reduce_one(const Theta &th, const double (&p)[D], double (&acc)[NT]) {
  double dot = 0.0;
#pragma unroll
  for (int k = 0; k < D; k++) dot += th.v[k] * p[k];
  double e = exp(dot);
  acc[0] += e;
  int t = 1 + D;
#pragma unroll
  for (int k = 0; k < D; k++) {
    double ek = e * p[k];
    acc[1 + k] += ek;
#pragma unroll
    for (int l = k; l < D; l++) acc[t++] += ek * p[l];
  }
}

// Each thread streams two candidates per 16-byte load and keeps two such loads
// per column in flight (64*D bytes per thread), which is what it takes to cover
// the HBM latency-bandwidth product with 2 x 256 threads per SM.  `rev` walks the
// block from its end: right after k_split, and on every other evaluation of a
// Newton loop, the tail of the statistic block is still in the 126 MB L2.
// One launch reduces two statistic blocks: blocks [0, A.grid) the dummy splits of this rank,
// the next B.grid blocks the flips (a few thousand rows: as a launch of its own it only added
// its latency to the critical path).  Each segment has its own partials, counter and output.
struct RedSeg {
  const double *stat; size_t cap; long n; int rev;
  double *partials; unsigned int *counter; double *out;
  int grid;
};

// Cross-GPU sum of the split sums, fused into the last block of k_reduce (the
// one exchange step of the path, 1+d+d(d+1)/2 doubles).  Every rank owns a *mailbox* in its
// HBM and maps its peers' mailboxes over NVLink (CUDA IPC, capi.cu: peer_setup).  Slot
// (parity, source rank) of a mailbox holds LITE_BOX_SLOT doubles: the NT sums, and in the
// last word the sequence number of the evaluation they belong to.  The finishing block
// stores its sums into its slot of EVERY mailbox (its own included), fences, publishes the
// sequence number, waits until all world slots of its own mailbox carry that number and
// adds them in rank order — so every rank gets the bitwise identical total, with no kernel
// launch and no host rendezvous on the critical path (an ncclAllReduce of 120 bytes there
// cost a launch plus its small-message latency on every evaluation).  Two parities: rank A
// can only reach evaluation s+2 after B has published s+1, i.e. after B has read s.
#define LITE_MAX_WORLD 16
#define LITE_BOX_SLOT 128
#define LITE_BOX_DOUBLES (2 * LITE_MAX_WORLD * LITE_BOX_SLOT)
struct PeerArgs {
  int world, rank;
  unsigned long long seq;          // this evaluation (>= 1, the same on every rank)
  double *box[LITE_MAX_WORLD];     // mailbox of rank r as this GPU addresses it; box[rank] is local
  unsigned long long *info;        // [0] set to seq when a wait timed out, [1] ns spent waiting for the peers
};
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ size_t box_slot(int parity, int src_rank) {
  return ((size_t)parity * LITE_MAX_WORLD + src_rank) * LITE_BOX_SLOT;
}
#ifndef LITE_PEER_TIMEOUT_NS
#define LITE_PEER_TIMEOUT_NS 4000000000ull
#endif

template <int D>
__global__ void __launch_bounds__(256) k_reduce(RedSeg A, RedSeg B, Theta th, PeerArgs P) {
  const bool second = (int)blockIdx.x >= A.grid;
  const double *__restrict__ stat = second ? B.stat : A.stat;
  const size_t cap = second ? B.cap : A.cap;
  const long n = second ? B.n : A.n;
  const int rev = second ? B.rev : A.rev;
  double *partials = second ? B.partials : A.partials;
  unsigned int *counter = second ? B.counter : A.counter;
  double *out = second ? B.out : A.out;
  const unsigned int nblk = (unsigned int)(second ? B.grid : A.grid);
  const unsigned int bid = blockIdx.x - (second ? (unsigned int)A.grid : 0u);
  constexpr int NT = 1 + D + D * (D + 1) / 2;
  double acc[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) acc[t] = 0.0;
  const long stride = (long)nblk * blockDim.x;
  const long gtid = (long)bid * blockDim.x + threadIdx.x;
  const long npair = n >> 1;
  long i = gtid;
  // two loads per column in flight only while the accumulators leave room for them
  for (; D <= 6 && i + stride < npair; i += 2 * stride) {
    const long ia = rev ? npair - 1 - i : i, ib = rev ? npair - 1 - (i + stride) : i + stride;
    double2 va[D], vb[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
      va[k] = ld_stat2((const double2 *)(stat + (size_t)k * cap) + ia);
      vb[k] = ld_stat2((const double2 *)(stat + (size_t)k * cap) + ib);
    }
    double p[D];
#pragma unroll
    for (int k = 0; k < D; k++) p[k] = va[k].x;
    reduce_one<D, NT>(th, p, acc);
#pragma unroll
    for (int k = 0; k < D; k++) p[k] = va[k].y;
    reduce_one<D, NT>(th, p, acc);
#pragma unroll
    for (int k = 0; k < D; k++) p[k] = vb[k].x;
    reduce_one<D, NT>(th, p, acc);
#pragma unroll
    for (int k = 0; k < D; k++) p[k] = vb[k].y;
    reduce_one<D, NT>(th, p, acc);
  }
  for (; i < npair; i += stride) {
    const long ia = rev ? npair - 1 - i : i;
    double2 va[D];
#pragma unroll
    for (int k = 0; k < D; k++) va[k] = ld_stat2((const double2 *)(stat + (size_t)k * cap) + ia);
    double p[D];
#pragma unroll
    for (int k = 0; k < D; k++) p[k] = va[k].x;
    reduce_one<D, NT>(th, p, acc);
#pragma unroll
    for (int k = 0; k < D; k++) p[k] = va[k].y;
    reduce_one<D, NT>(th, p, acc);
  }
  if ((n & 1) && gtid == 0) {  // odd tail
    double p[D];
#pragma unroll
    for (int k = 0; k < D; k++) p[k] = stat[(size_t)k * cap + n - 1];
    reduce_one<D, NT>(th, p, acc);
  }
  __shared__ double sh[8][NT];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < NT; t++) {
    double v = acc[t];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[wid][t] = v;
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x < NT) {
    double v = 0;
    int nw = blockDim.x >> 5;
    for (int w2 = 0; w2 < nw; w2++) v += sh[w2][threadIdx.x];
    partials[(size_t)bid * NT + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(counter, 1u);
    last = (prev == nblk - 1);
  }
  __syncthreads();
  if (last) {
    // fixed-order finish: warp w sums output t = w, w+8, ..; its lanes stride over the
    // blocks' partials and a shuffle tree combines them (same order every run)
    __threadfence();
    __shared__ double fin[NT];
    for (int t = wid; t < NT; t += (blockDim.x >> 5)) {
      double v = 0;
      for (unsigned int b = lane; b < nblk; b += 32) v += __ldcg(partials + (size_t)b * NT + t);
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) fin[t] = v;
    }
    if (threadIdx.x == 0) *counter = 0;
    __syncthreads();
    if (second || P.world <= 1) {  // flips are replicated on every rank: nothing to exchange
      if (threadIdx.x < NT) out[threadIdx.x] = fin[threadIdx.x];
      return;
    }
    const int par = (int)(P.seq & 1ull);
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) t0 = gtimer_ns();
    for (int idx = threadIdx.x; idx < P.world * NT; idx += blockDim.x) {
      const int r = idx / NT, t = idx - r * NT;
      *(volatile double *)(P.box[r] + box_slot(par, P.rank) + t) = fin[t];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < P.world) {
      const int r = threadIdx.x;
      st_release_sys((unsigned long long *)(P.box[r] + box_slot(par, P.rank) + (LITE_BOX_SLOT - 1)), P.seq);
      const unsigned long long *mine =
          (const unsigned long long *)(P.box[P.rank] + box_slot(par, r) + (LITE_BOX_SLOT - 1));
      const unsigned long long w0 = gtimer_ns();
      while (ld_acquire_sys(mine) < P.seq) {
        if (gtimer_ns() - w0 > LITE_PEER_TIMEOUT_NS) { atomicExch(P.info, P.seq); break; }
      }
    }
    __syncthreads();
    if (threadIdx.x < NT) {
      double v = 0;
      for (int r = 0; r < P.world; r++) v += __ldcg(P.box[P.rank] + box_slot(par, r) + threadIdx.x);
      out[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) P.info[1] = gtimer_ns() - t0;
  }
}

// ---------------------------------------------------------------------------
// FP64 peak: dependent DFMA chains, 8 per thread
// ---------------------------------------------------------------------------
__global__ void k_dfma_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;
}
