// sk_plan_host.cpp -- host-side planning for the B200 type-3 NUFFT and the quadrature rules.
//
//  * exp-of-semicircle tap polynomials and the deconvolution factor (2/w)/phihat(xi), fitted once
//    per context in __float128 so that the double coefficients are correctly rounded;
//  * the canonical Gauss-Legendre / Gauss-Jacobi(0,p) rules the reference takes from
//    FastGaussQuadrature (src/quadrature.jl:36-42), by Newton on the three-term recurrence in
//    long double.  Callers that already own a rule (the Julia host) pass it through sk_rule_set
//    instead and this generator is not used.
#include "sk_plan.h"

#include <quadmath.h>

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

typedef __float128 q128;
typedef long double ld;

static const q128 Q_PI = 3.14159265358979323846264338327950288419716939937510Q;

static q128 es_phi_q(q128 z, q128 beta) {
  q128 t = 1 - z * z;
  if (t <= 0) return 0;
  return expq(beta * (sqrtq(t) - 1));
}

// Chebyshev interpolation of f at n first-kind nodes on [-1,1] -> monomial coefficients (ascending)
template <class F>
static void cheb_fit_monomial(int n, F f, std::vector<q128> &mono) {
  std::vector<q128> fx(n), c(n);
  for (int k = 0; k < n; ++k) fx[k] = f(cosq((2 * k + 1) * Q_PI / (2 * n)));
  for (int j = 0; j < n; ++j) {
    q128 acc = 0;
    for (int k = 0; k < n; ++k) acc += fx[k] * cosq(j * (2 * k + 1) * Q_PI / (2 * n));
    c[j] = acc * 2 / n;
  }
  c[0] /= 2;
  // T_0 = 1, T_1 = s, T_{j+1} = 2 s T_j - T_{j-1}
  std::vector<q128> tm1(n, 0), t0(n, 0), t1(n, 0);
  mono.assign(n, 0);
  tm1[0] = 1;
  mono[0] += c[0];
  if (n > 1) {
    t0[1] = 1;
    mono[1] += c[1];
  }
  for (int j = 2; j < n; ++j) {
    for (int q = 0; q < n; ++q) t1[q] = (q > 0 ? 2 * t0[q - 1] : 0) - tm1[q];
    for (int q = 0; q < n; ++q) mono[q] += c[j] * t1[q];
    tm1 = t0;
    t0 = t1;
  }
}

template <class F>
static void cheb_fit_series(int n, F f, std::vector<q128> &c) {
  std::vector<q128> fx(n);
  c.assign(n, 0);
  for (int k = 0; k < n; ++k) fx[k] = f(cosq((2 * k + 1) * Q_PI / (2 * n)));
  for (int j = 0; j < n; ++j) {
    q128 acc = 0;
    for (int k = 0; k < n; ++k) acc += fx[k] * cosq(j * (2 * k + 1) * Q_PI / (2 * n));
    c[j] = acc * 2 / n;
  }
  c[0] /= 2;
}

static int sk_plan_make_es_fit(int w, SkEsPlan *out) {
  if (w < 4 || w > SK_WMAX || (w & 1)) return -1;
  std::memset(out, 0, sizeof(*out));
  out->w = w;
  out->beta = 2.30 * w;  // sigma = 2 (Barnett, Magland, af Klinteberg 2019)
  const q128 beta = out->beta;
  // ---- tap polynomials --------------------------------------------------------------------
  for (int i = 0; i < w / 2; ++i) {
    std::vector<q128> mono;
    cheb_fit_monomial(SK_NC, [&](q128 s) {
      q128 z = (s / 2 + i - (q128)w / 2 + 0.5Q) * 2 / w;
      return es_phi_q(z, beta);
    }, mono);
    for (int q = 0; q < SK_NC / 2; ++q) {
      out->E[i][q] = (double)mono[2 * q];
      out->O[i][q] = (double)mono[2 * q + 1];
    }
  }
  // ---- deconvolution factor -----------------------------------------------------------------
  // phihat(xi) = 2 int_0^1 phi(z) cos(xi z) dz by Gauss-Legendre (long-double nodes, q128 sums)
  const int ng = 128;
  std::vector<double> gx(ng), gw(ng);
  // nodes good to 1e-16 are enough: the integrand is entire in z on the node set
  // (quadrature error, not node error, is what matters), but use the long-double generator anyway.
  std::vector<ld> gxl(ng), gwl(ng);
  {
    std::vector<double> xd(ng), wd(ng);
    if (sk_plan_gauss_rule(ng, 0.0, xd.data(), wd.data()) != 0) return -2;
    for (int i = 0; i < ng; ++i) { gxl[i] = xd[i]; gwl[i] = wd[i]; }
  }
  out->ximax = (double)(Q_PI * w / 4) * 1.002;      // sigma >= 1.996 on both the mode and the target side
  const q128 ximax = out->ximax;
  auto phihat = [&](q128 xi) {
    q128 acc = 0;
    for (int i = 0; i < ng; ++i) {
      q128 z = ((q128)1 + (q128)gxl[i]) / 2;
      acc += (q128)gwl[i] / 2 * es_phi_q(z, beta) * cosq(xi * z);
    }
    return 2 * acc;
  };
  std::vector<q128> qc;
  cheb_fit_series(SK_NQMAX, [&](q128 tau) {
    q128 xi = ximax * sqrtq((tau + 1) / 2);
    return ((q128)2 / w) / phihat(xi);
  }, qc);
  out->nq = SK_NQMAX;
  // drop trailing terms below 1e-18 relative
  while (out->nq > 4 && fabsq(qc[out->nq - 1]) < 1e-18Q * fabsq(qc[0])) out->nq--;
  for (int j = 0; j < SK_NQMAX; ++j) out->qc[j] = j < out->nq ? (double)qc[j] : 0.0;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Gauss-Jacobi rule for the weight (1+x)^p on [-1,1] (alpha = 0, beta = p); p == 0 is Legendre.
// Evaluates P_n^{(0,p)} and P_{n-1}^{(0,p)} by the recurrence and polishes asymptotic guesses
// of the zeros with Newton; Christoffel weights w_k = 2^{p+1} / ((1-x_k^2) P_n'(x_k)^2).  (The
// algebraically equivalent form in P_{n-1} alone loses ~6 digits in the few weights next to x = +-1.)
// ---------------------------------------------------------------------------------------------
namespace {
struct JacobiEval {
  int n;
  ld p;
  std::vector<ld> A, B, C;  // P_{k+1} = (A_k x + B_k) P_k - C_k P_{k-1}
  JacobiEval(int n_, ld p_) : n(n_), p(p_), A(n_), B(n_), C(n_) {
    for (int k = 1; k < n; ++k) {
      ld kk = k, s = 2 * kk + p;
      ld den = 2 * (kk + 1) * (kk + p + 1) * s;
      A[k] = (s + 1) * (s + 2) * s / den;
      B[k] = -(s + 1) * p * p / den;
      C[k] = 2 * kk * (kk + p) * (s + 2) / den;
    }
  }
  void eval(ld x, ld &pn, ld &pnm1) const {
    ld p0 = 1, p1 = ((p + 2) * x - p) / 2;
    if (n == 0) { pn = 1; pnm1 = 0; return; }
    for (int k = 1; k < n; ++k) {
      ld p2 = (A[k] * x + B[k]) * p1 - C[k] * p0;
      p0 = p1;
      p1 = p2;
    }
    pn = p1;
    pnm1 = p0;
  }
};
}  // namespace

// recurrence coefficients of P_k^{(0,p)} as (hi, lo) double pairs for the device generator (sk_rules.cuh);
// out arrays hold 2*n doubles each, entry k at [2k], [2k+1]; k = 0 is unused
int sk_plan_jacobi_coeffs(int n, double p_, double *A, double *B, double *C) {
  if (n < 1 || !(p_ > -1.0)) return -1;
  const ld p = p_;
  auto split = [](ld v, double *dst) {
    const double hi = (double)v;
    dst[0] = hi;
    dst[1] = (double)(v - (ld)hi);
  };
  A[0] = A[1] = B[0] = B[1] = C[0] = C[1] = 0.0;
  for (int k = 1; k < n; ++k) {
    const ld kk = k, s = 2 * kk + p;
    const ld den = 2 * (kk + 1) * (kk + p + 1) * s;
    split((s + 1) * (s + 2) * s / den, A + 2 * k);
    split(-(s + 1) * p * p / den, B + 2 * k);
    split(2 * kk * (kk + p) * (s + 2) / den, C + 2 * k);
  }
  return 0;
}

int sk_plan_gauss_rule(int n, double p_, double *no, double *wt) {
  if (n < 1 || !(p_ > -1.0)) return -1;
  const ld p = p_;
  const ld PI_L = 3.14159265358979323846264338327950288L;
  JacobiEval J(n, p);
  const ld nn = n, s = 2 * nn + p;
  const ld wfac = powl(2.0L, p + 1);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    int k = n - i;  // counted from x = +1
    ld x = cosl((2 * (ld)k - 0.5L) * PI_L / (2 * nn + p + 1));
    for (int it = 0; it < 60; ++it) {
      ld pn, pm;
      J.eval(x, pn, pm);
      // (2n+p)(1-x^2) P_n' = n(-p - (2n+p) x) P_n + 2 n (n+p) P_{n-1}
      ld dp = (nn * (-p - s * x) * pn + 2 * nn * (nn + p) * pm) / (s * (1 - x * x));
      ld dx = pn / dp;
      x -= dx;
      if (fabsl(dx) <= 4.4e-19L * (1 + fabsl(x))) break;
    }
    ld pn, pm;
    J.eval(x, pn, pm);
    const ld dp = (nn * (-p - s * x) * pn + 2 * nn * (nn + p) * pm) / (s * (1 - x * x));
    no[i] = (double)x;
    wt[i] = (double)(wfac / ((1 - x * x) * dp * dp));
  }
  for (int i = 1; i < n; ++i)
    if (!(no[i] > no[i - 1])) return -3;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// J_nu(z), nu = 0 .. numax, 0 <= z <= 2 nint, as piecewise polynomials: interval i = [2i, 2i+2],
// nc monomial coefficients in t = z - (2i+1).  Values by Miller's backward recurrence in __float128
// (normalised with J_0 + 2 sum_k J_2k = 1), fitted at Chebyshev nodes: the double coefficients are
// correctly rounded and the table is good to ~1e-16 absolute (sk_hankel.h uses it for the local
// Chebyshev expansions of the nonuniform Hankel transform).
// ---------------------------------------------------------------------------------------------
static void bessel_miller_q(int numax, q128 z, q128 *J) {
  if (z == 0) {
    for (int n = 0; n <= numax; ++n) J[n] = (n == 0) ? 1 : 0;
    return;
  }
  const int N = 2 * ((int)(double)z / 2) + 120;      // even, well above z
  q128 jp1 = 0, j = 1e-300Q, sum = 0;
  std::vector<q128> keep(numax + 1, 0);
  for (int n = N; n >= 1; --n) {
    const q128 jm1 = (2 * (q128)n / z) * j - jp1;     // J_{n-1} = (2n/z) J_n - J_{n+1}
    jp1 = j;
    j = jm1;
    // j is now J_{n-1} (unnormalised)
    if (n - 1 <= numax) keep[n - 1] = j;
    if ((n - 1) % 2 == 0) sum += (n - 1 == 0) ? j : 2 * j;
  }
  for (int n = 0; n <= numax; ++n) J[n] = keep[n] / sum;
}

static int sk_plan_bessel_table_fit(int numax, int nint, int nc, double *tab) {
  if (numax < 0 || nint < 1 || nc < 2 || !tab) return -1;
  for (int nu = 0; nu <= numax; ++nu)
    for (int i = 0; i < nint; ++i) {
      std::vector<q128> mono;
      cheb_fit_monomial(nc, [&](q128 t) {
        std::vector<q128> J(numax + 1);
        bessel_miller_q(numax, (q128)(2 * i + 1) + t, J.data());
        return J[nu];
      }, mono);
      for (int q = 0; q < nc; ++q) tab[((size_t)nu * nint + i) * nc + q] = (double)mono[q];
    }
  return 0;
}

// The fits above run in __float128 (~20 ms for a plan, more for the Bessel table): done once per process and width,
// not once per context -- a device group opens one context per GPU.
int sk_plan_make_es(int w, SkEsPlan *out) {
  static std::mutex mu;
  static std::map<int, SkEsPlan> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(w);
  if (it == cache.end()) {
    SkEsPlan P;
    const int rc = sk_plan_make_es_fit(w, &P);
    if (rc != 0) return rc;
    it = cache.emplace(w, P).first;
  }
  *out = it->second;
  return 0;
}

int sk_plan_bessel_table(int numax, int nint, int nc, double *tab) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, std::vector<double>> cache;
  if (numax < 0 || nint < 1 || nc < 2 || !tab) return -1;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(numax, nint, nc);
  auto it = cache.find(key);
  if (it == cache.end()) {
    std::vector<double> t((size_t)(numax + 1) * nint * nc);
    const int rc = sk_plan_bessel_table_fit(numax, nint, nc, t.data());
    if (rc != 0) return rc;
    it = cache.emplace(key, std::move(t)).first;
  }
  std::memcpy(tab, it->second.data(), sizeof(double) * it->second.size());
  return 0;
}
