// gen_host.cpp — TEST HARNESS, not part of the library.  Compiles the general-face source the
// GPU kernels run (lite_b200/csrc/gen_face.h) for the host, so that its logic can be checked
// against the exact oracle on a machine without a GPU (tests/test_gen_face_cpu.py).  The
// product path is the CUDA kernels only.
//
// usage: gen_host <in.bin> <out.bin>
// in.bin : int32 header [nV nH nF nFH nS nNB nB d nC], then the lite_tess_soa arrays in the
//          order of the struct, then d feature ids (int32), then nC candidates: he (int32[nC]),
//          x (f64[nC]), angle (f64[nC]).
// out.bin: split e2 (int32[nC]), split p1, p2 (f64[2nC] each), split stat (f64[nC*d]),
//          merge stat (f64[nNB*d]), flip e2, right (int32[2nB] each), flip p2 (f64[4nB]),
//          flip stat (f64[2nB*d]), then int32 [n_general_faces, overflow, insert cas 1-4, remove cas 1-4].
#include <stdio.h>
#include <stdlib.h>

#include <vector>

static int g_cases[2][5];
#define GEN_CASE(kind, cas) (g_cases[kind][cas]++)
#include "../../lite_b200/csrc/gen_face.h"

template <typename T>
static std::vector<T> rd(FILE *f, size_t n) {
  std::vector<T> v(n ? n : 1);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}
template <typename T>
static void wr(FILE *f, const std::vector<T> &v, size_t n) {
  if (n) fwrite(v.data(), sizeof(T), n, f);
}

int main(int argc, char **argv) {
  if (argc != 3) { fprintf(stderr, "usage: gen_host in.bin out.bin\n"); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  auto hd = rd<int32_t>(f, 9);
  const int nV = hd[0], nH = hd[1], nF = hd[2], nFH = hd[3], nS = hd[4], nNB = hd[5], nB = hd[6], d = hd[7], nC = hd[8];
  auto vx = rd<double>(f, nV), vy = rd<double>(f, nV);
  auto he_src = rd<int32_t>(f, nH), he_twin = rd<int32_t>(f, nH), he_next = rd<int32_t>(f, nH), he_face = rd<int32_t>(f, nH),
       he_seg = rd<int32_t>(f, nH), he_next_hf = rd<int32_t>(f, nH), he_prev_hf = rd<int32_t>(f, nH);
  auto he_dir = rd<uint8_t>(f, nH);
  auto he_length = rd<double>(f, nH);
  auto face_inside = rd<uint8_t>(f, nF);
  auto f_off = rd<int32_t>(f, nF), f_cnt = rd<int32_t>(f, nF), f_list = rd<int32_t>(f, nFH);
  auto seg_start = rd<int32_t>(f, nS), seg_end = rd<int32_t>(f, nS), seg_n = rd<int32_t>(f, nS);
  auto seg_bnd = rd<uint8_t>(f, nS);
  auto nb = rd<int32_t>(f, nNB), bl = rd<int32_t>(f, nB);
  auto fid = rd<int32_t>(f, d);
  auto c_he = rd<int32_t>(f, nC);
  auto c_x = rd<double>(f, nC), c_ang = rd<double>(f, nC);
  fclose(f);

  gen::Cycles C;
  gen::derive_cycles(nH, nF, he_next.data(), he_face.data(), he_twin.data(), face_inside.data(), f_off.data(),
                     f_cnt.data(), f_list.data(), C);
  gen::Tess T{};
  T.n_he = nH; T.n_faces = nF; T.vx = vx.data(); T.vy = vy.data();
  T.he_src = he_src.data(); T.he_twin = he_twin.data(); T.he_next = he_next.data(); T.he_face = he_face.data();
  T.he_seg = he_seg.data(); T.he_next_hf = he_next_hf.data(); T.he_prev_hf = he_prev_hf.data();
  T.he_dir = he_dir.data(); T.face_inside = face_inside.data(); T.seg_bnd = seg_bnd.data(); T.seg_n_edges = seg_n.data();
  T.f_outer = C.f_outer.data(); T.f_hole_off = C.f_hole_off.data(); T.f_hole_cnt = C.f_hole_cnt.data();
  T.hole_he = C.hole_he.data();
  gen::Prog G{};
  G.d = d;
  for (int k = 0; k < d; k++) { G.fid[k] = fid[k]; G.mask |= 1u << fid[k]; }
  std::vector<gen::Pt> buf(1 << 16);
  gen::Scratch S{buf.data(), (int)buf.size(), 0, 0};
  int overflow = 0;

  std::vector<int32_t> s_e2(nC ? nC : 1);
  std::vector<double> s_p1(2 * (size_t)nC + 1), s_p2(2 * (size_t)nC + 1), s_stat((size_t)nC * d + 1);
  for (int i = 0; i < nC; i++) {
    gen::SplitRes r;
    gen::split_general(T, G, S, c_he[i], c_x[i], c_ang[i], &s_stat[(size_t)i * d], 1, r);
    s_e2[i] = r.e2; s_p1[2 * i] = r.p1x; s_p1[2 * i + 1] = r.p1y; s_p2[2 * i] = r.p2x; s_p2[2 * i + 1] = r.p2y;
    overflow |= S.overflow;
  }
  std::vector<double> m_stat((size_t)nNB * d + 1);
  for (int i = 0; i < nNB; i++) {
    gen::merge_general(T, G, S, seg_start[nb[i]], &m_stat[(size_t)i * d], 1);
    overflow |= S.overflow;
  }
  const int nFl = 2 * nB;
  std::vector<int32_t> f_e2(nFl ? nFl : 1), f_right(nFl ? nFl : 1);
  std::vector<double> f_p2(2 * (size_t)nFl + 1), f_stat((size_t)nFl * d + 1);
  for (int i = 0; i < nFl; i++) {
    const int sg = bl[i < nB ? i : i - nB];
    const int e1 = (i < nB) ? he_twin[seg_start[sg]] : seg_end[sg];   // all_flips, lib/ttessel.cpp:1834-1846
    gen::FlipRes r;
    gen::flip_general(T, G, S, e1, &f_stat[(size_t)i * d], 1, r);
    f_e2[i] = r.e2; f_right[i] = r.right; f_p2[2 * i] = r.p2x; f_p2[2 * i + 1] = r.p2y;
    overflow |= S.overflow;
  }
  f = fopen(argv[2], "wb");
  if (!f) { perror(argv[2]); return 2; }
  wr(f, s_e2, nC); wr(f, s_p1, 2 * (size_t)nC); wr(f, s_p2, 2 * (size_t)nC); wr(f, s_stat, (size_t)nC * d);
  wr(f, m_stat, (size_t)nNB * d);
  wr(f, f_e2, nFl); wr(f, f_right, nFl); wr(f, f_p2, 2 * (size_t)nFl); wr(f, f_stat, (size_t)nFl * d);
  int32_t tail[10] = {C.n_general, overflow, g_cases[0][1], g_cases[0][2], g_cases[0][3], g_cases[0][4],
                      g_cases[1][1], g_cases[1][2], g_cases[1][3], g_cases[1][4]};
  fwrite(tail, sizeof(int32_t), 10, f);
  fclose(f);
  return 0;
}
