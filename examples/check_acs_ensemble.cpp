// examples/check_acs_ensemble.cpp — applications/checkACS.cpp for MANY seeds at once.
// The reference program simulates one SMF chain under the ACS model (segs
// is_segment_internal theta = -log tau, edges edge_length theta = (tau-1)/pi, :113-117) and
// writes one row of statistics per batch of `nipl` proposals (:136-148); a Monte Carlo
// envelope test runs it once per seed.  Here the chains are one ensemble on the GPU
// (lite_smf_*, include/lite_b200.h) and the same columns come back for every chain.
// Options as the reference: -t tau -s seed -w width -i n_i -j nipl, plus -n chains.
//
//   g++ -std=c++17 -Iinclude examples/check_acs_ensemble.cpp -Llite_b200 -lliteb200
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lite_b200.h"
#include "lite_b200_host.h"

#define CK(x)                                                                  \
  do {                                                                         \
    if ((x) != 0) { fprintf(stderr, "%s: %s\n", #x, lite_last_error()); return 1; } \
  } while (0)

int main(int argc, char **argv) {
  int opt, seed = 5, width = 1, n_i = 100, nipl = 1, chains = 1024, prec = 16;
  double tau = 6;
  while ((opt = getopt(argc, argv, "t:s:w:i:j:n:")) != EOF) {
    switch (opt) {
      case 't': tau = atof(optarg); break;
      case 's': seed = atoi(optarg); break;
      case 'w': width = atoi(optarg); break;
      case 'i': n_i = atoi(optarg); break;
      case 'j': nipl = atoi(optarg); break;
      case 'n': chains = atoi(optarg); break;
      default: printf("Invalid argument\n"); break;
    }
  }
  lite_ctx *ctx;
  CK(lite_ctx_create(&ctx, 0, 0, 1, NULL));
  // capacity: the stationary size grows like tau^2 * width^2 (about 15 tau^2 cells in the unit square)
  const int cap = (int)std::fmin(10896.0, 40.0 * tau * tau * width * width + 64);
  CK(lite_smf_create(ctx, chains, cap, 0, 0, width, width, (uint64_t)seed));
  const int32_t ids[2] = {LITE_FEAT_IS_SEGMENT_INTERNAL, LITE_FEAT_EDGE_LENGTH};
  const double th[2] = {-log(tau), (tau - 1) / 3.14159265358979323846};
  CK(lite_smf_set_model(ctx, 2, ids, th, 0.33, 0.33));
  std::vector<double> rows((size_t)n_i * LITE_SMF_COLS * chains);
  CK(lite_smf_run(ctx, n_i, nipl, rows.data()));
  float ms = 0;
  lite_smf_last_ms(ctx, &ms);
  fprintf(stderr, "%d chains x %d batches x %d proposals in %.1f ms (%.3g chain-steps/s)\n", chains, n_i, nipl, ms,
          (double)chains * n_i * nipl / (ms * 1e-3));
  // the reference's stat file, with the chain id in front (applications/checkACS.cpp:121-148)
  printf("chain i l(T) n(T) n_bl(T) n_nbl(T) m(T) prop_S acc_S prop_M acc_M prop_F acc_F\n");
  for (int c = 0; c < chains; c++)
    for (int i = 0; i < n_i; i++) {
      const double *r = rows.data() + (size_t)i * LITE_SMF_COLS * chains + c;
      printf("%d %d %.*g", c, i, prec, r[0]);
      for (int k = 1; k < 11; k++) printf(" %.0f", r[(size_t)k * chains]);
      printf("\n");
    }
  // the reference's intersection file for the final state (:100-103, :151-172): the four window
  // edges and the two mid-lines against every internal edge of every chain
  {
    const double w = width;
    const double probes[6 * 4] = {0, 0, w, 0,  w, 0, w, w,  w, w, 0, w,  0, w, 0, 0,  0, w / 2, w, w / 2,  w / 2, 0, w / 2, w};
    int64_t n_found = 0;
    CK(lite_smf_intersections(ctx, 6, probes, 0, NULL, NULL, NULL, &n_found));
    std::vector<int32_t> ch((size_t)n_found + 1), seg((size_t)n_found + 1);
    std::vector<double> xy(2 * (size_t)n_found + 2);
    CK(lite_smf_intersections(ctx, 6, probes, n_found, ch.data(), seg.data(), xy.data(), &n_found));
    FILE *fi = fopen("inter_final.txt", "w");
    if (fi) {
      fprintf(fi, "chain seg x y\n");
      for (int64_t k = 0; k < n_found; k++) fprintf(fi, "%d %d %.*g %.*g\n", ch[k], seg[k], prec, xy[2 * k], prec, xy[2 * k + 1]);
      fclose(fi);
    }
    fprintf(stderr, "%lld intersections of internal edges with the 6 probe segments\n", (long long)n_found);
  }
  // the final tessellation of the first chain in the reference's text format (:175)
  lite_host_tess *t;
  CK(lite_smf_export(ctx, 0, &t));
  std::vector<char> text((size_t)lite_host_write_text(t, NULL, 0) + 1);
  lite_host_write_text(t, text.data(), (int64_t)text.size());
  FILE *f = fopen("tessel_chain0.txt", "w");
  if (f) { fputs(text.data(), f); fclose(f); }
  lite_host_destroy(t);
  lite_ctx_destroy(ctx);
  return 0;
}
