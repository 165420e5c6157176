// gen_launch.cuh — kernels for the candidates that touch a face which is not a simple polygon
// (gen_face.h).  Launched only when the uploaded tessellation has such faces (a window with
// holes); a rectangular window never gets here, so the fast kernels pay one flag test.
//
// k_gen_splits re-derives every candidate's halfedge (same Philox draw and sampler as k_split,
// or the fed list) and evaluates the ones k_split skipped (slot flag SF_GENERAL); k_gen_mf
// does the merges and flips with a general face on either side BEFORE k_merges_flips runs, which
// skips them but still sums their rows.  One candidate per thread, polygons in a per-thread
// block of `scratch`; the work is rare and irregular, nothing here is tuned.
#pragma once
#include "gen_face.h"
#include "kernels.cuh"

#define LITE_GEN_TB 32
#define LITE_GEN_PTS 4096  // polygon vertices a thread can hold at once (96 KB per thread)

struct GenSplitArgs {
  SplitArgs A;
  int philox, record, stat, check;
  gen::Pt *scratch;
  unsigned long long *overflow;  // candidates whose polygons did not fit the scratch block
};

__global__ void __launch_bounds__(LITE_GEN_TB) k_gen_splits(TessDev T, gen::Tess GT, Prog G, GenSplitArgs X) {
  const SplitArgs &A = X.A;
  if (*T.err) return;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  gen::Scratch S;
  S.base = X.scratch + (size_t)tid * LITE_GEN_PTS; S.cap = LITE_GEN_PTS; S.used = 0; S.overflow = 0;
  gen::Prog GP;
  GP.d = X.stat ? G.d : 0; GP.mask = X.stat ? G.mask : 0;
  for (int k = 0; k < LITE_MAX_D; k++) GP.fid[k] = G.fid[k];
  unsigned long long cx = 0, cs = 0;
  unsigned int bad = 0, over = 0;
  double dummy[1];
  for (long i = tid; i < A.n_local; i += stride) {
    int m;
    double x, angle, U = 0.0;
    if (X.philox) {
      const bool iid = A.sel_sorted != nullptr;
      const long j = iid ? (long)A.sel_id[i] : A.j0 + i;
      uint32_t r[4];
      philox4x32_10(A.ctr0 + (uint64_t)j, 0u, A.seed, r);
      const uint32_t spare = ((r[0] & 31u) << 17) | ((r[1] & 63u) << 11) | ((r[2] & 31u) << 6) | (r[3] & 63u);
      const double sel = iid ? A.sel_sorted[i] : ((double)j + (double)spare * (1.0 / 4194304.0)) * A.sel_step;
      m = sample_slot(T, sel);
      if (!(__ldg(T.flg + m) & SF_GENERAL)) continue;
      x = u53(r[0], r[1]);
      U = u53(r[2], r[3]);
      angle = acos(1.0 - 2.0 * U);
    } else {
      m = __ldg(T.he_slot + A.in_he[i]);
      if (!(__ldg(T.flg + m) & SF_GENERAL)) continue;
      x = A.in_x[i]; angle = A.in_angle[i];
    }
    const int h1 = __ldg(T.slot_he + m);
    gen::SplitRes o;
    S.overflow = 0;
    gen::split_general(GT, GP, S, h1, x, angle, X.stat ? A.stat + A.base + i : dummy, A.cap, o);
    if (S.overflow) { over++; for (int k = 0; k < GP.d; k++) A.stat[(size_t)k * A.cap + A.base + i] = 0.0; }
    if (o.e2 < 0) bad++;
    if (X.record) {
      A.r_e1[A.base + i] = h1; A.r_e2[A.base + i] = o.e2;
      A.r_p1[A.base + i] = make_double2(o.p1x, o.p1y); A.r_p2[A.base + i] = make_double2(o.p2x, o.p2y);
      A.r_x[A.base + i] = x; A.r_angle[A.base + i] = angle; A.r_u[A.base + i] = U;
    }
    if (X.check) {
      unsigned long long key = ((unsigned long long)(unsigned)h1 << 32) | (unsigned)o.e2;
      key *= 0x9E3779B97F4A7C15ull; key ^= key >> 29;
      cx ^= key; cs += (unsigned long long)(unsigned)o.e2;
    }
  }
  if (X.check && (cx | cs)) { atomicXor(A.checksum, cx); atomicAdd(A.checksum + 1, cs); }
  if (bad) atomicAdd(A.status, (unsigned long long)bad);
  if (over) atomicAdd(X.overflow, (unsigned long long)over);
}

__global__ void __launch_bounds__(LITE_GEN_TB) k_gen_mf(TessDev T, gen::Tess GT, Prog G, MFArgs A, gen::Pt *scratch,
                                                        unsigned long long *overflow) {
  if (*T.err) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  gen::Scratch S;
  S.base = scratch + (size_t)tid * LITE_GEN_PTS; S.cap = LITE_GEN_PTS; S.used = 0; S.overflow = 0;
  gen::Prog GP;
  GP.d = G.d; GP.mask = G.mask;
  for (int k = 0; k < LITE_MAX_D; k++) GP.fid[k] = G.fid[k];
  const int nm = A.nm, nf = 2 * A.nb;
  unsigned int over = 0;
  for (int c = tid; c < nm + nf; c += stride) {
    S.overflow = 0;
    if (c < nm) {
      const int e = T.seg_he_start[A.nb_seg[c]];
      const int f1 = GT.he_face[e], f2 = GT.he_face[GT.he_twin[e]];
      if (!(T.f_gen[f1] | T.f_gen[f2])) continue;
      A.m_he[c] = e;
      gen::merge_general(GT, GP, S, e, A.mstat + c, (size_t)nm);
      if (S.overflow) { over++; for (int k = 0; k < G.d; k++) A.mstat[(size_t)k * nm + c] = 0.0; }
    } else {
      const int i = c - nm;
      const int sg = A.b_seg[i < A.nb ? i : i - A.nb];
      const int e1 = (i < A.nb) ? T.he_twin[T.seg_he_start[sg]] : T.seg_he_end[sg];
      const int f1 = GT.he_face[e1], f2 = GT.he_face[GT.he_twin[e1]];
      if (!(T.f_gen[f1] | T.f_gen[f2])) continue;
      gen::FlipRes o;
      gen::flip_general(GT, GP, S, e1, A.fstat + i, (size_t)nf, o);
      A.f_e1[i] = e1; A.f_e2[i] = o.e2; A.f_right[i] = o.right; A.f_p2[i] = make_double2(o.p2x, o.p2y);
      if (o.e2 < 0) atomicAdd(A.status, 1ull);
      if (S.overflow) { over++; for (int k = 0; k < G.d; k++) A.fstat[(size_t)k * nf + i] = 0.0; }
    }
  }
  if (over) atomicAdd(overflow, (unsigned long long)over);
}
