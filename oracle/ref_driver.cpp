// ORACLE — TEST INFRASTRUCTURE ONLY.
// C-ABI driver around the UNMODIFIED reference class HmcSampler
// (tmg/src/HmcSampler.{h,cpp}, compiled from the sources extracted from
// /root/reference/dependecies/tmg_0.3.tar.gz into oracle/_ref/src by the Makefile).
// It does what tmg/src/rtmg.cpp:14-66 does, minus Rcpp (absent from this image):
// construct, addLinearConstraint per row of F, setInitialValue, n x sampleNext.
#include <cstdint>
#include "HmcSampler.h"

extern "C" int ref_rtmg(int n, int seed, int d, const double* initial, int numlin,
                        const double* F /*c x d col-major*/, const double* g,
                        double* out /*n x d col-major, row = sample*/) {
  HmcSampler hmc1(d, seed);                       // rtmg.cpp:23
  for (int i = 0; i < numlin; ++i) {              // rtmg.cpp:35-37
    VectorXd f(d);
    for (int j = 0; j < d; ++j) f(j) = F[(size_t)i + (size_t)j * numlin];
    hmc1.addLinearConstraint(f, g[i]);
  }
  VectorXd init(d);
  for (int j = 0; j < d; ++j) init(j) = initial[j];
  hmc1.setInitialValue(init);                     // rtmg.cpp:55
  for (int i = 0; i < n; ++i) {                   // rtmg.cpp:59-61
    MatrixXd s = hmc1.sampleNext();
    for (int j = 0; j < d; ++j) out[(size_t)i + (size_t)j * n] = s(j);
  }
  return 0;
}
