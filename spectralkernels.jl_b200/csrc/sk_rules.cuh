// sk_rules.cuh -- canonical Gauss-Legendre / Gauss-Jacobi(0,p) rules on the device (QuadRule,
// src/quadrature.jl:27-47; the reference takes them from FastGaussQuadrature.jl).
//
// One thread per node: Newton on P_n^{(0,p)} evaluated with the three-term recurrence in double-double
// arithmetic (~106 bits, more than the 64-bit long-double host generator it replaces), Christoffel weights
// w = 2^(p+1) / ((1 - x^2) P_n'(x)^2).  n = 8192 takes a few milliseconds instead of ~0.3 s on the host,
// which matters when the singularity exponent is a fitted parameter and every optimiser step needs a new
// Jacobi rule.  The recurrence coefficients are prepared on the host in long double and passed as
// (hi, lo) pairs.
#pragma once
#include "sk_math.h"

struct sk_dd { double hi, lo; };

SK_HD sk_dd dd_make(double a) { sk_dd r; r.hi = a; r.lo = 0.0; return r; }
SK_HD sk_dd dd_fast_renorm(double s, double e) {   // |s| >= |e|
  sk_dd r;
  r.hi = sk_add(s, e);
  r.lo = sk_add(e, -sk_add(r.hi, -s));
  return r;
}
SK_HD sk_dd dd_add(sk_dd a, sk_dd b) {
  // two-sum of the high parts, then fold in the low parts
  const double s = sk_add(a.hi, b.hi);
  const double bb = sk_add(s, -a.hi);
  double e = sk_add(sk_add(a.hi, -sk_add(s, -bb)), sk_add(b.hi, -bb));
  e = sk_add(e, sk_add(a.lo, b.lo));
  return dd_fast_renorm(s, e);
}
SK_HD sk_dd dd_neg(sk_dd a) { a.hi = -a.hi; a.lo = -a.lo; return a; }
SK_HD sk_dd dd_mul(sk_dd a, sk_dd b) {
  const double p = sk_mul(a.hi, b.hi);
  double e = sk_fma(a.hi, b.hi, -p);
  e = sk_fma(a.hi, b.lo, e);
  e = sk_fma(a.lo, b.hi, e);
  return dd_fast_renorm(p, e);
}
SK_HD sk_dd dd_div(sk_dd a, sk_dd b) {
  const double q1 = a.hi / b.hi;
  sk_dd r = dd_add(a, dd_neg(dd_mul(b, dd_make(q1))));
  const double q2 = r.hi / b.hi;
  r = dd_add(r, dd_neg(dd_mul(b, dd_make(q2))));
  const double q3 = r.hi / b.hi;
  return dd_add(dd_fast_renorm(q1, q2), dd_make(q3));
}

struct SkRuleJob {          // one rule: n nodes, exponent p
  int n;
  double p;
  const sk_dd *A, *B, *C;   // recurrence coefficients, k = 1 .. n-1:  P_{k+1} = (A_k x + B_k) P_k - C_k P_{k-1}
  double *no, *wt;
};

SK_HD void sk_jacobi_eval_dd(const SkRuleJob &J, sk_dd x, sk_dd *pn, sk_dd *pm) {
  sk_dd p0 = dd_make(1.0);
  // P_1 = ((p + 2) x - p) / 2
  sk_dd p1 = dd_mul(dd_add(dd_mul(dd_make(J.p + 2.0), x), dd_make(-J.p)), dd_make(0.5));
  for (int k = 1; k < J.n; ++k) {
    const sk_dd t = dd_mul(dd_add(dd_mul(J.A[k], x), J.B[k]), p1);
    const sk_dd p2 = dd_add(t, dd_neg(dd_mul(J.C[k], p0)));
    p0 = p1;
    p1 = p2;
  }
  *pn = p1;
  *pm = p0;
}

// P_n'(x) from (2n+p)(1-x^2) P_n' = n(-p - (2n+p) x) P_n + 2 n (n+p) P_{n-1}
SK_HD sk_dd sk_jacobi_deriv_dd(const SkRuleJob &J, sk_dd x, sk_dd pn, sk_dd pm) {
  const double nn = (double)J.n;
  const sk_dd s = dd_add(dd_make(2.0 * nn), dd_make(J.p));
  const sk_dd t1 = dd_mul(dd_mul(dd_make(nn), dd_add(dd_make(-J.p), dd_neg(dd_mul(s, x)))), pn);
  const sk_dd t2 = dd_mul(dd_mul(dd_make(2.0 * nn), dd_add(dd_make(nn), dd_make(J.p))), pm);
  const sk_dd omx2 = dd_add(dd_make(1.0), dd_neg(dd_mul(x, x)));
  return dd_div(dd_add(t1, t2), dd_mul(s, omx2));
}

// P_n and P_{n-1} in plain double (for the first Newton steps only)
SK_HD void sk_jacobi_eval_d(const SkRuleJob &J, double x, double *pn, double *pm) {
  double p0 = 1.0, p1 = ((J.p + 2.0) * x - J.p) * 0.5;
  for (int k = 1; k < J.n; ++k) {
    const double p2 = (J.A[k].hi * x + J.B[k].hi) * p1 - J.C[k].hi * p0;
    p0 = p1;
    p1 = p2;
  }
  *pn = p1;
  *pm = p0;
}

SK_HD void sk_gauss_node(const SkRuleJob &J, int i) {
  if (J.n == 1 && J.p == 0.0) { J.no[0] = 0.0; J.wt[0] = 2.0; return; }
  const int k = J.n - i;                                   // counted from x = +1 (ascending output)
  const double th = (2.0 * k - 0.5) * 3.141592653589793 / (2.0 * J.n + J.p + 1.0);
  // Newton in plain double down to rounding level (a double-double evaluation of the recurrence costs ~8x more) ...
  double xd = cos(th);
  const double nn = (double)J.n, sp = 2.0 * nn + J.p;
  for (int it = 0; it < 6; ++it) {
    double pn, pm;
    sk_jacobi_eval_d(J, xd, &pn, &pm);
    const double dp = (nn * (-J.p - sp * xd) * pn + 2.0 * nn * (nn + J.p) * pm) / (sp * (1.0 - xd * xd));
    const double dx = pn / dp;
    xd -= dx;
    if (fabs(dx) <= 1e-13 * (1.0 + fabs(xd))) break;
  }
  // ... then in double-double: each step squares the error, so one or two steps reach ~1e-30
  sk_dd x = dd_make(xd);
  for (int it = 0; it < 8; ++it) {
    sk_dd pn, pm;
    sk_jacobi_eval_dd(J, x, &pn, &pm);
    const sk_dd dp = sk_jacobi_deriv_dd(J, x, pn, pm);
    const sk_dd dx = dd_div(pn, dp);
    x = dd_add(x, dd_neg(dx));
    if (fabs(dx.hi) <= 1e-13 * (1.0 + fabs(x.hi))) break;    // the next correction would be ~dx^2 < 1e-26
  }
  sk_dd pn, pm;
  sk_jacobi_eval_dd(J, x, &pn, &pm);
  const sk_dd dp = sk_jacobi_deriv_dd(J, x, pn, pm);
  const sk_dd omx2 = dd_add(dd_make(1.0), dd_neg(dd_mul(x, x)));
  const sk_dd w = dd_div(dd_make(exp2(J.p + 1.0)), dd_mul(omx2, dd_mul(dp, dp)));
  J.no[i] = x.hi;              // x.hi is the correctly rounded double of hi + lo (|lo| <= ulp/2)
  J.wt[i] = w.hi;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(64) k_gauss_rules(const SkRuleJob *__restrict__ jobs, int njobs) {
  const SkRuleJob J = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < J.n) sk_gauss_node(J, i);
}
#endif
