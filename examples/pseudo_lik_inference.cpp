// examples/pseudo_lik_inference.cpp — the reference's demo (demos/pseudo_lik_inference.cpp)
// against this repository's include/ttessel.h: simulate a CRTT tessellation with
// SMFChain, print the closed-form pseudo-likelihood estimate (:80-81), then fit the model.
// The reference demo fits with `PseudoLikStochApprox`, a class that does not exist in its
// library; the fit here uses PLInferenceNOIS (lib/ttessel.cpp:3454-3540), the
// estimator the library does provide.  Same options as the reference demo: -t tau -s seed
// -w width -b burn-in.
//
//   g++ -std=c++17 -Iinclude examples/pseudo_lik_inference.cpp -Llite_b200 -lliteb200
#include <unistd.h>

#include <cmath>
#include <cstdlib>
#include <iostream>

#include "ttessel.h"

int main(int argc, char **argv) {
  int opt, seed = 5, width = 1, n_burnin = 5000;
  double tau = 2;
  while ((opt = getopt(argc, argv, "t:s:w:b:")) != EOF) {
    switch (opt) {
      case 't': tau = atof(optarg); break;
      case 's': seed = atoi(optarg); break;
      case 'w': width = atoi(optarg); break;
      case 'b': n_burnin = atoi(optarg); break;
      default: std::cout << "Invalid argument" << std::endl; break;
    }
  }
  rnd = new CGAL::Random(seed);

  TTessel obsTes;
  Energy trueMod, fitMod;
  obsTes.insert_window(Rectangle(Point2(0, 0), Point2(width, width)));
  SMFChain smf = SMFChain(&trueMod, 0.33, 0.33);
  trueMod.add_theta_segs(log(tau));
  trueMod.add_features_segs(minus_is_segment_internal);
  trueMod.set_ttessel(&obsTes);
  std::cout << "True parameter value " << trueMod.get_theta().segs[0] << std::endl;
  smf.step(n_burnin);
  std::cout << std::endl << "Some features of the \"observed\" tessellation:" << std::endl;
  std::cout << "\tnumber of internal segments " << obsTes.number_of_segments() - 4 << std::endl;
  std::cout << "\tnumber of internal non-blocking segments " << obsTes.number_of_non_blocking_segments() << std::endl;
  std::cout << "\tnumber of internal blocking segments " << obsTes.number_of_blocking_segments() << std::endl;
  std::cout << "\ttotal perimeter " << obsTes.get_total_internal_length() << std::endl;
  const double exact = log(obsTes.number_of_non_blocking_segments() * CGAL_PI / obsTes.get_total_internal_length());
  std::cout << "Exact pseudo-likelihood estimate " << exact << std::endl << std::endl;

  fitMod.add_theta_segs(0.0);
  fitMod.add_features_segs(minus_is_segment_internal);
  fitMod.set_ttessel(&obsTes);
  PLInferenceNOIS pl(&fitMod, /*store=*/true, /*lambda=*/1.0);
  const unsigned int n_iter = pl.Run(1e-6, 100);
  const double est = pl.GetEstimate().segs[0];
  std::cout << "PLInferenceNOIS: " << n_iter << " Newton iterations, " << pl.NumberOfSplits()
            << " dummy splits, parameter estimate: " << est << std::endl;
  delete rnd;
  // the CRTT optimum does not depend on the dummy sample: the fit must land on it
  return std::fabs(est - exact) < 1e-4 ? 0 : 1;
}
