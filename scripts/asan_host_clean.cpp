// Random line patterns through insert_segment and the three cleaning passes under ASan + UBSan:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -o /tmp/asan_host scripts/asan_host_clean.cpp lite_b200/csrc/host_tess.cpp && /tmp/asan_host
#include <cstdio>
#include <random>
#include <cmath>
#include "../lite_b200/csrc/host_tess.h"
int main() {
  int bad = 0;
  for (int seed = 0; seed < 120; seed++) {
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<double> U(0, 1);
    lite::HostTess t;
    t.insert_window_rect(0, 0, 1, 1);
    std::string err;
    int n = 5 + (int)(U(g) * 35);
    for (int i = 0; i < n; i++) {
      double cx = 0.05 + 0.9 * U(g), cy = 0.05 + 0.9 * U(g), a = M_PI * U(g), l = 0.1 + 0.6 * U(g);
      double s[4] = {cx - 0.5 * l * cos(a), cy - 0.5 * l * sin(a), cx + 0.5 * l * cos(a), cy + 0.5 * l * sin(a)};
      int id;
      if (!lite::insert_segment(t, s, &id, err)) { printf("seed %d insert: %s\n", seed, err.c_str()); }
    }
    for (int it = 0; it < 30 && t.arrangement_only; it++) {
      if (!lite::remove_xvertices(t, 0.01, nullptr, err)) { printf("seed %d X: %s\n", seed, err.c_str()); break; }
      if (!lite::remove_ivertices(t, 0, nullptr, err)) { printf("seed %d I: %s\n", seed, err.c_str()); break; }
      if (!lite::remove_lvertices(t, 0, nullptr, err)) { printf("seed %d L: %s\n", seed, err.c_str()); break; }
    }
    if (t.arrangement_only) { bad++; continue; }
    std::string c = t.check();
    if (!c.empty()) { printf("seed %d check: %s\n", seed, c.c_str()); bad++; }
    t.export_soa();
    std::string txt = lite::write_text(t);
    lite::HostTess u;
    if (!lite::read_text(u, txt, err)) { printf("seed %d reread: %s\n", seed, err.c_str()); bad++; }
  }
  printf("not cleaned or failed: %d of 120\n", bad);
  return 0;
}
