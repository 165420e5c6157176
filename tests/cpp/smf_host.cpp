// smf_host.cpp — TEST HARNESS, not part of the library.  Compiles the chain source
// the GPU kernel runs (lite_b200/csrc/smf_chain.h) for the host, so that its logic
// can be replayed against the exact oracle on a machine without a GPU
// (tests/test_smf_chain_cpu.py).  The product path is the CUDA kernel only.
//
// usage: smf_host <cap> <n_steps> <seed> <chain_id> <p_split> <p_merge> <trace.bin> <d> (<feature_id> <theta>) x d
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../lite_b200/csrc/smf_chain.h"

int main(int argc, char **argv) {
  if (argc < 9) { fprintf(stderr, "bad arguments\n"); return 2; }
  const int cap = atoi(argv[1]);
  const long n_steps = atol(argv[2]);
  smf::StepRng R;
  R.seed = strtoull(argv[3], nullptr, 0);
  R.chain = strtoull(argv[4], nullptr, 0);
  smf::ChainParams cp;
  cp.p_split = atof(argv[5]); cp.p_merge = atof(argv[6]);
  const char *out = argv[7];
  smf::Energy G;
  G.d = atoi(argv[8]); G.mask = 0;
  if (argc != 9 + 2 * G.d) { fprintf(stderr, "bad feature list\n"); return 2; }
  for (int i = 0; i < G.d; i++) {
    G.fid[i] = atoi(argv[9 + 2 * i]);
    G.theta[i] = atof(argv[10 + 2 * i]);
    G.mask |= 1u << G.fid[i];
  }
  std::vector<char> blob(smf::chain_bytes(cap) + 256);
  char *base = (char *)(((uintptr_t)blob.data() + 255) & ~(uintptr_t)255);
  smf::Chain c;
  smf::Header H;
  c.H = &H;
  smf::bind(c, base, cap);
  smf::init_window(c, cap, 0.0, 0.0, 1.0, 1.0);
  std::vector<smf::TraceRec> tr((size_t)n_steps);
  long done = 0;
  for (; done < n_steps; done++) {
    if (!smf::smf_step(c, G, cp, R, &tr[done])) break;
    const int bad = smf::check(c);
    if (bad) { fprintf(stderr, "inconsistent chain state (code %d) after step %ld\n", bad, done); return 1; }
  }
  smf::store(c);
  FILE *f = fopen(out, "wb");
  if (!f) { perror(out); return 2; }
  fwrite(tr.data(), sizeof(smf::TraceRec), (size_t)done, f);
  fclose(f);
  double row[smf::N_STAT_COLS];
  smf::stat_row(c, row, 1);
  printf("steps %ld status %d", done, c.H->status);
  for (int k = 0; k < smf::N_STAT_COLS; k++) printf(" %.17g", row[k]);
  printf("\n");
  return 0;
}
