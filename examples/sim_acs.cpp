// examples/sim_acs.cpp — the reference's second demo (demos/sim_acs.cpp) as a client of this
// repository's include/ttessel.h: one SMF chain under the ACS model with its two optional
// face terms, the reference's per-loop trace and end-of-run feature summary.
//   energy = -log(tau) * #internal segments + (tau-1)/pi * edge_length
//            + tau^4 * f * sum of squared face areas + g * sum of complementary angles
// (demos/sim_acs.cpp:64-76).  Options as there: -t tau -f f -g g -s seed -w width -i loops
// -j proposals per loop.  Every energy variation of the chain is evaluated on the GPU
// (Energy::variation -> lite_candidates_* / lite_pl_sums); at the end the energy the chain has
// accumulated is compared with the energy of the final tessellation recomputed from scratch by
// the host feature functions, and the exit code says whether they agree.
//
//   g++ -std=c++17 -Iinclude examples/sim_acs.cpp -Llite_b200 -lliteb200
#include <unistd.h>

#include <cmath>
#include <cstdlib>
#include <iostream>

#include "ttessel.h"

namespace {
struct Options {
  int seed = 5, width = 1, loops = 10, per_loop = 250;
  double tau = 12, f = 0, g = 0;
};
bool parse(int argc, char **argv, Options &o) {
  int c;
  bool ok = true;
  while ((c = getopt(argc, argv, "t:f:g:s:w:i:j:")) != EOF) {
    if (c == 't') o.tau = atof(optarg);
    else if (c == 'f') o.f = atof(optarg);
    else if (c == 'g') o.g = atof(optarg);
    else if (c == 's') o.seed = atoi(optarg);
    else if (c == 'w') o.width = atoi(optarg);
    else if (c == 'i') o.loops = atoi(optarg);
    else if (c == 'j') o.per_loop = atoi(optarg);
    else { std::cout << "Invalid argument" << std::endl; ok = false; }
  }
  return ok;
}
void acs_model(Energy &e, const Options &o) {
  e.add_theta_segs(-log(o.tau));
  e.add_features_segs(is_segment_internal);
  e.add_theta_edges((o.tau - 1) / CGAL_PI);
  e.add_features_edges(edge_length);
  e.add_features_faces(face_area_2);
  e.add_theta_faces(pow(o.tau, 4) * o.f);
  e.add_features_faces(face_sum_of_angles);
  e.add_theta_faces(o.g);
}
}  // namespace

int main(int argc, char **argv) {
  Options o;
  parse(argc, argv, o);
  rnd = new CGAL::Random(o.seed);
  TTessel tes;
  tes.insert_window(Rectangle(Point2(0, 0), Point2(o.width, o.width)));
  Energy model;
  SMFChain chain(&model, 0.33, 0.33);
  acs_model(model, o);
  model.set_ttessel(&tes);

  std::cerr << "step energy nb_of_int_segments int_length" << std::endl;
  ModifCounts total = {};
  for (int i = 0; i != o.loops; i++) {
    ModifCounts c = chain.step(o.per_loop);
    total.proposed_S += c.proposed_S; total.accepted_S += c.accepted_S;
    total.proposed_M += c.proposed_M; total.accepted_M += c.accepted_M;
    total.proposed_F += c.proposed_F; total.accepted_F += c.accepted_F;
    std::cerr << i << " " << model.get_value() << " " << tes.number_of_segments() - 4 << " "
              << (tes.get_total_internal_length() - 4 * o.width) / 2 << std::endl;
  }
  const double inner = tes.get_total_internal_length() - 4 * o.width;
  std::cerr << std::endl << "Tessellation features:" << std::endl << std::endl;
  std::cerr << "Number of segments: " << tes.number_of_segments() - 4 << std::endl;
  std::cerr << "Total internal length: " << inner << std::endl;
  std::cerr << "Sum of squared area of faces: " << sum_of_faces_squared_areas(&tes) << std::endl;
  std::cerr << "Angle statistic: " << sum_of_min_angles(&tes) << std::endl << std::endl;
  std::cerr << "accepted/proposed: splits " << total.accepted_S << "/" << total.proposed_S << ", merges "
            << total.accepted_M << "/" << total.proposed_M << ", flips " << total.accepted_F << "/"
            << total.proposed_F << std::endl;
  std::cout << "Coordinates of segments: " << std::endl;
  tes.print_all_segments();

  // the energy the chain accumulated, variation by variation on the device, against the energy of
  // the final tessellation computed from scratch by the host feature functions (Energy::set_ttessel)
  Energy scratch;
  acs_model(scratch, o);
  scratch.set_ttessel(&tes);
  const double tracked = model.get_value(), from_scratch = scratch.get_value();
  std::cerr << "energy accumulated by the chain " << tracked << ", recomputed " << from_scratch << std::endl;
  delete rnd;
  const bool ok = tes.is_valid() && std::fabs(tracked - from_scratch) <= 1e-8 * (1.0 + std::fabs(from_scratch));
  return ok ? 0 : 1;
}
