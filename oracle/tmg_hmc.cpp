// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; never linked
// into liblineqgpr_b200.so.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this.
//
// CPU restatement (C++17, std::vector, no Eigen/Rcpp) of the exact bouncing HMC
// of tmg 0.3 for LINEAR constraints only (lineqGPR always passes q=NULL):
//   tmg/src/HmcSampler.cpp:21      min_t
//   tmg/src/HmcSampler.cpp:23-28   ctor / seeding of knuth_b
//   tmg/src/HmcSampler.cpp:43-133  sampleNext
//   tmg/src/HmcSampler.cpp:152-189 _getNextLinearHitTime
//   tmg/src/HmcSampler.cpp:244-265 _verifyConstraints
//   tmg/src/rtmg.cpp:14-66         driver loop
// (tmg = /root/reference/dependecies/tmg_0.3.tar.gz)
//
// Parity pinning: `make -C oracle ref` compiles the UNMODIFIED reference
// HmcSampler.cpp against oracle/eigen_shim (a minimal Eigen-API stand-in,
// Eigen itself is not in this image) into oracle/_ref/libtmg_ref.so;
// tests/test_oracle.py checks this restatement bit-for-bit against it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

namespace {

constexpr double kMinT = 0.00001;  // HmcSampler.cpp:21

struct Draws {
  // mode 0: libstdc++ knuth_b + normal_distribution, exactly HmcSampler.h:50-52
  // mode 1: injected stream consumed in order (d values per (re)start)
  std::knuth_b eng;
  std::normal_distribution<> nd;
  const double* inj = nullptr;
  int64_t n_inj = 0, pos = 0;
  bool exhausted = false;
  double next() {
    if (inj) {
      if (pos >= n_inj) { exhausted = true; return 0.0; }
      return inj[pos++];
    }
    ++pos;
    return nd(eng);
  }
};

struct LinCon {           // HmcSampler.h:21-24 (array of structs, one heap vector per row)
  std::vector<double> f;
  double g;
};

struct Sampler {
  int dim;
  std::vector<LinCon> lin;
  std::vector<double> last;
  Draws rng;
  bool faithful_copy;     // keep the per-constraint struct copy of HmcSampler.cpp:156,256
  // counters
  int64_t bounces = 0, vel_restarts = 0, chk_restarts = 0, hit_searches = 0;
  // optional per-bounce trace
  double* trace_t = nullptr; int32_t* trace_cn = nullptr; int64_t trace_cap = 0, n_trace = 0;
  int64_t max_restarts = 0;  // 0 = unbounded like the reference
  bool gave_up = false;

  static double dot(const std::vector<double>& f, const std::vector<double>& v) {
    double s = 0.0;
    const size_t n = f.size();
    for (size_t i = 0; i < n; ++i) s += f[i] * v[i];
    return s;
  }

  // HmcSampler.cpp:152-189
  void next_linear_hit(const std::vector<double>& a, const std::vector<double>& b,
                       double& hit_time, int& cn) {
    hit_time = 0;
    ++hit_searches;
    for (size_t i = 0; i != lin.size(); ++i) {
      double fa, fb, g;
      if (faithful_copy) {
        LinCon lc = lin[i];                       // :156 copies the row
        fa = dot(lc.f, a); fb = dot(lc.f, b); g = lc.g;
      } else {
        const LinCon& lc = lin[i];
        fa = dot(lc.f, a); fb = dot(lc.f, b); g = lc.g;
      }
      double u = std::sqrt(fa * fa + fb * fb);    // :159
      if (u > g && u > -g) {                      // :160
        double phi = std::atan2(-fa, fb);         // :161
        double t1 = std::acos(-g / u) - phi;      // :162
        if (t1 < 0) t1 += 2 * M_PI;               // :165
        if (std::abs(t1) < kMinT) t1 = 0;         // :166
        else if (std::abs(t1 - 2 * M_PI) < kMinT) t1 = 0;   // :167
        double t2 = -t1 - 2 * phi;                // :170 (uses wrapped/zeroed t1)
        if (t2 < 0) t2 += 2 * M_PI;               // :171
        if (t2 < 0) t2 += 2 * M_PI;               // :172
        if (std::abs(t2) < kMinT) t2 = 0;         // :174
        else if (std::abs(t2 - 2 * M_PI) < kMinT) t2 = 0;   // :175
        double t = t1;                            // :178-181
        if (t1 == 0) t = t2;
        else if (t2 == 0) t = t1;
        else t = (t1 < t2 ? t1 : t2);
        if (t > kMinT && (hit_time == 0 || t < hit_time)) {  // :183
          hit_time = t;
          cn = (int)i;
        }
      }
    }
  }

  // HmcSampler.cpp:244-265 (linear part)
  double verify(const std::vector<double>& b) {
    double r = 0;
    for (size_t i = 0; i != lin.size(); ++i) {
      double check;
      if (faithful_copy) { LinCon lc = lin[i]; check = dot(lc.f, b) + lc.g; }
      else               { const LinCon& lc = lin[i]; check = dot(lc.f, b) + lc.g; }
      if (i == 0 || check < r) r = check;
    }
    return r;
  }

  // HmcSampler.cpp:43-133, returnTrace=false, no quadratic constraints
  void sample_next() {
    const double T = M_PI / 2;                    // :47
    std::vector<double> b = last;                 // :48 (outside the retry loop)
    std::vector<double> a(dim), new_b(dim), hit_vel(dim), bb(dim);
    int64_t restarts = 0;
    while (2) {                                   // :51
      if (max_restarts > 0 && restarts > max_restarts) { gave_up = true; return; }
      double velsign = 0;                         // :53
      for (int i = 0; i < dim; ++i) a[i] = rng.next();   // :54-55
      if (rng.exhausted) return;
      double tt = T;                              // :57
      double t1; int cn1 = 0;
      while (1) {                                 // :64
        t1 = 0;
        if (!lin.empty()) next_linear_hit(a, b, t1, cn1);   // :66-68
        double t = t1;                            // :75 (t2 == 0: no quadratics)
        if (t == 0 || tt < t) break;              // :82
        if (trace_cap > n_trace) { trace_t[n_trace] = t; trace_cn[n_trace] = cn1; }
        ++n_trace;
        ++bounces;
        tt = tt - t;                              // :90
        const double st = std::sin(t), ct = std::cos(t);
        for (int i = 0; i < dim; ++i) {           // :91-93
          new_b[i] = st * a[i] + ct * b[i];
          hit_vel[i] = ct * a[i] - st * b[i];
        }
        b = new_b;
        const std::vector<double>& f = lin[cn1].f;           // :106
        double f2 = dot(f, f);                    // :107
        double alpha = dot(f, hit_vel) / f2;      // :108
        const double two_alpha = 2 * alpha;
        for (int i = 0; i < dim; ++i) a[i] = hit_vel[i] - two_alpha * f[i];   // :109
        velsign = dot(a, f);                      // :110
        if (velsign < 0) break;                   // :112
      }
      if (velsign < 0) { ++vel_restarts; ++restarts; continue; }   // :117
      const double st = std::sin(tt), ct = std::cos(tt);
      for (int i = 0; i < dim; ++i) bb[i] = st * a[i] + ct * b[i];   // :119
      double check = verify(bb);                  // :121
      if (check >= 0) { last = bb; return; }      // :122-128
      ++chk_restarts; ++restarts;                 // :130 fallthrough to while(2)
    }
  }
};

}  // namespace

extern "C" {

// Mirror of the .Call("rtmg", n, seed, initial, numlin, F, g, 0, NULL) contract
// (tmg/R/tmg.R:104, tmg/src/rtmg.cpp:14-66).  F is c x d COLUMN-major (R layout),
// out is n x d COLUMN-major with row = sample, all in the canonical (whitened) frame.
// injected: nullable; when non-null the Gaussian draws are read from it in order.
// stats[0..5] = bounces, vel_restarts, chk_restarts, draws_consumed, hit_searches, n_trace
// returns 0 ok, 1 injected stream exhausted, 2 restart cap hit.
int oracle_rtmg(int n, int seed, int d, const double* initial, int numlin,
                const double* F, const double* g,
                const double* injected, int64_t n_injected,
                int faithful_copy, int64_t max_restarts,
                double* trace_t, int32_t* trace_cn, int64_t trace_cap,
                double* out, int64_t* stats) {
  Sampler s;
  s.dim = d;
  s.rng.eng.seed(seed);                           // HmcSampler.cpp:26
  s.rng.inj = injected; s.rng.n_inj = n_injected;
  s.faithful_copy = faithful_copy != 0;
  s.max_restarts = max_restarts;
  s.trace_t = trace_t; s.trace_cn = trace_cn; s.trace_cap = trace_cap;
  s.lin.resize(numlin);
  for (int i = 0; i < numlin; ++i) {              // rtmg.cpp:35-37
    s.lin[i].f.resize(d);
    for (int j = 0; j < d; ++j) s.lin[i].f[j] = F[(size_t)i + (size_t)j * numlin];
    s.lin[i].g = g[i];
  }
  s.last.assign(initial, initial + d);            // rtmg.cpp:55
  int rc = 0;
  for (int i = 0; i < n; ++i) {                   // rtmg.cpp:59-61
    s.sample_next();
    if (s.rng.exhausted) { rc = 1; break; }
    if (s.gave_up) { rc = 2; break; }
    for (int j = 0; j < d; ++j) out[(size_t)i + (size_t)j * n] = s.last[j];
  }
  if (stats) {
    stats[0] = s.bounces; stats[1] = s.vel_restarts; stats[2] = s.chk_restarts;
    stats[3] = s.rng.pos; stats[4] = s.hit_searches; stats[5] = s.n_trace;
  }
  return rc;
}

// The stream tmg would draw for a given seed: knuth_b + std::normal_distribution
// (HmcSampler.h:50-52, .cpp:26,55).  Used to feed the SAME draws to the GPU path.
void oracle_normal_stream(int seed, int64_t n, double* out) {
  std::knuth_b eng; eng.seed(seed);
  std::normal_distribution<> nd;
  for (int64_t i = 0; i < n; ++i) out[i] = nd(eng);
}

// One hit-time evaluation (HmcSampler.cpp:157-181 for a single constraint given
// fa, fb, g); returns the candidate t (0 = none).  For unit-testing the device
// hit-time function against the same scalar inputs.
double oracle_hit_time(double fa, double fb, double g) {
  double u = std::sqrt(fa * fa + fb * fb);
  if (!(u > g && u > -g)) return 0.0;
  double phi = std::atan2(-fa, fb);
  double t1 = std::acos(-g / u) - phi;
  if (t1 < 0) t1 += 2 * M_PI;
  if (std::abs(t1) < kMinT) t1 = 0;
  else if (std::abs(t1 - 2 * M_PI) < kMinT) t1 = 0;
  double t2 = -t1 - 2 * phi;
  if (t2 < 0) t2 += 2 * M_PI;
  if (t2 < 0) t2 += 2 * M_PI;
  if (std::abs(t2) < kMinT) t2 = 0;
  else if (std::abs(t2 - 2 * M_PI) < kMinT) t2 = 0;
  double t = t1;
  if (t1 == 0) t = t2;
  else if (t2 == 0) t = t1;
  else t = (t1 < t2 ? t1 : t2);
  return t > kMinT ? t : 0.0;
}

}  // extern "C"
