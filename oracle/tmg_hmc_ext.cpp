// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// The chain of oracle/tmg_hmc.cpp (tmg/src/HmcSampler.cpp:43-133,152-189,244-265) once more, with
// the same control flow but in EXTENDED precision (x87 long double, 64-bit mantissa), injected
// draws only.  Purpose: when the GPU path and the double-precision oracle differ by 1e-9..1e-7 at
// the ill-conditioned headline shapes (cond(Sigma_eta) ~ 1e7), this tells which side carries the
// rounding: both are compared with the extended chain (tests/test_gpu_hmc.py).  Thresholds (min_t)
// keep their double values so that the discrete events are the reference's.
// Also here: a batch form of the scalar hit time for the randomised test of the device function.
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

typedef long double R;
const R kPiL = 3.141592653589793238462643383279502884L;
const R kMinT = (R)0.00001;                       // HmcSampler.cpp:21, the double value

struct SamplerExt {
  int dim = 0;
  std::vector<std::vector<R>> f;
  std::vector<R> g;
  std::vector<R> last;
  const double* inj = nullptr;
  int64_t n_inj = 0, pos = 0;
  bool exhausted = false;
  int64_t bounces = 0, vel_restarts = 0, chk_restarts = 0, hit_searches = 0, max_restarts = 0;
  bool gave_up = false;

  R next() {
    if (pos >= n_inj) { exhausted = true; return 0; }
    return (R)inj[pos++];
  }
  static R dot(const std::vector<R>& x, const std::vector<R>& y) {
    R s = 0;
    for (size_t i = 0; i < x.size(); ++i) s += x[i] * y[i];
    return s;
  }
  void next_linear_hit(const std::vector<R>& a, const std::vector<R>& b, R& hit_time, int& cn) {   // :152-189
    hit_time = 0;
    ++hit_searches;
    for (size_t i = 0; i != f.size(); ++i) {
      const R fa = dot(f[i], a), fb = dot(f[i], b), gi = g[i];
      const R u = std::sqrt(fa * fa + fb * fb);
      if (u > gi && u > -gi) {
        const R phi = std::atan2(-fa, fb);
        R t1 = std::acos(-gi / u) - phi;
        if (t1 < 0) t1 += 2 * kPiL;
        if (std::fabs(t1) < kMinT) t1 = 0;
        else if (std::fabs(t1 - 2 * kPiL) < kMinT) t1 = 0;
        R t2 = -t1 - 2 * phi;
        if (t2 < 0) t2 += 2 * kPiL;
        if (t2 < 0) t2 += 2 * kPiL;
        if (std::fabs(t2) < kMinT) t2 = 0;
        else if (std::fabs(t2 - 2 * kPiL) < kMinT) t2 = 0;
        R t = t1;
        if (t1 == 0) t = t2;
        else if (t2 == 0) t = t1;
        else t = (t1 < t2 ? t1 : t2);
        if (t > kMinT && (hit_time == 0 || t < hit_time)) { hit_time = t; cn = (int)i; }
      }
    }
  }
  R verify(const std::vector<R>& b) {             // :244-265
    R r = 0;
    for (size_t i = 0; i != f.size(); ++i) {
      const R check = dot(f[i], b) + g[i];
      if (i == 0 || check < r) r = check;
    }
    return r;
  }
  void sample_next() {                            // :43-133
    const R T = kPiL / 2;
    std::vector<R> b = last, a(dim), new_b(dim), hit_vel(dim), bb(dim);
    int64_t restarts = 0;
    while (2) {
      if (max_restarts > 0 && restarts > max_restarts) { gave_up = true; return; }
      R velsign = 0;
      for (int i = 0; i < dim; ++i) a[i] = next();
      if (exhausted) return;
      R tt = T;
      R t1; int cn1 = 0;
      while (1) {
        t1 = 0;
        if (!f.empty()) next_linear_hit(a, b, t1, cn1);
        const R t = t1;
        if (t == 0 || tt < t) break;
        ++bounces;
        tt = tt - t;
        const R st = std::sin(t), ct = std::cos(t);
        for (int i = 0; i < dim; ++i) {
          new_b[i] = st * a[i] + ct * b[i];
          hit_vel[i] = ct * a[i] - st * b[i];
        }
        b = new_b;
        const std::vector<R>& fr = f[cn1];
        const R alpha = dot(fr, hit_vel) / dot(fr, fr);
        for (int i = 0; i < dim; ++i) a[i] = hit_vel[i] - 2 * alpha * fr[i];
        velsign = dot(a, fr);
        if (velsign < 0) break;
      }
      if (velsign < 0) { ++vel_restarts; ++restarts; continue; }
      const R st = std::sin(tt), ct = std::cos(tt);
      for (int i = 0; i < dim; ++i) bb[i] = st * a[i] + ct * b[i];
      if (verify(bb) >= 0) { last = bb; return; }
      ++chk_restarts; ++restarts;
    }
  }
};

}  // namespace

extern "C" {

// Same contract as oracle_rtmg (canonical frame, F c x d column-major, out n x d column-major with
// row = sample, rounded to double on output), injected draws mandatory, extended precision inside.
// unwhiten (nullable, d x d column-major Ri) and shift (nullable, d): when given, out holds
// Ri z + shift evaluated in extended precision, i.e. the eta-frame sample tmg:R/tmg.R:106 returns.
// stats[0..4] = bounces, vel_restarts, chk_restarts, draws consumed, hit searches.
int oracle_rtmg_ext(int n, int d, const double* initial, int numlin, const double* F, const double* g,
                    const double* injected, int64_t n_injected, int64_t max_restarts,
                    const double* unwhiten, const double* shift, double* out, int64_t* stats) {
  SamplerExt s;
  s.dim = d;
  s.inj = injected; s.n_inj = n_injected;
  s.max_restarts = max_restarts;
  s.f.assign(numlin, std::vector<R>(d));
  s.g.resize(numlin);
  for (int i = 0; i < numlin; ++i) {
    for (int j = 0; j < d; ++j) s.f[i][j] = (R)F[(size_t)i + (size_t)j * numlin];
    s.g[i] = (R)g[i];
  }
  s.last.resize(d);
  for (int j = 0; j < d; ++j) s.last[j] = (R)initial[j];
  int rc = 0;
  for (int i = 0; i < n; ++i) {
    s.sample_next();
    if (s.exhausted) { rc = 1; break; }
    if (s.gave_up) { rc = 2; break; }
    for (int j = 0; j < d; ++j) {
      R v = s.last[j];
      if (unwhiten) {
        v = shift ? (R)shift[j] : 0;
        for (int k = 0; k < d; ++k) v += (R)unwhiten[(size_t)j + (size_t)k * d] * s.last[k];
      }
      out[(size_t)i + (size_t)j * n] = (double)v;
    }
  }
  if (stats) { stats[0] = s.bounces; stats[1] = s.vel_restarts; stats[2] = s.chk_restarts; stats[3] = s.pos; stats[4] = s.hit_searches; }
  return rc;
}

double oracle_hit_time(double fa, double fb, double g);     // tmg_hmc.cpp

// batch form of oracle_hit_time for the randomised test of the device function
void oracle_hit_time_batch(int64_t n, const double* fa, const double* fb, const double* g, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = oracle_hit_time(fa[i], fb[i], g[i]);
}

}  // extern "C"
