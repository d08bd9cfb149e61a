/*
 * sk_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * CPU restatement, in plain C, of the arithmetic on the K(r) hot path of
 * pbeckman/SpectralKernels.jl.  Nothing under oracle/ is part of the shipped
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * CPU baseline, never as the thing measured as "ours".
 *
 * What is restated here (reference file:line):
 *   sko_direct_cis     src/quadrature.jl:113-128  direct Fourier summation,
 *                      int[j] = sum_k buf[k]*cispi(2*no[k]*x[j]); this branch is
 *                      the reference's own definition of what
 *                      finufft1d3 (src/utils.jl:10) approximates.
 *   sko_direct_bessel  src/quadrature.jl:145-160  direct Bessel summation (dim>=2).
 *   sko_gauss_rule     src/quadrature.jl:36-42    gausslegendre(n) /
 *                      gaussjacobi(n, 0.0, p).  The reference takes these from
 *                      FastGaussQuadrature.jl (third party, unpinned in
 *                      Project.toml:10, source absent from /root/reference);
 *                      here the published definition (zeros of P_n^{(0,p)},
 *                      Christoffel weights) is evaluated by Newton iteration on
 *                      the three-term recurrence in 80-bit long double.
 *   sko_nufft1d3       src/utils.jl:10  f_j = sum_k c_k exp(+i 2 pi x_j w_k) by a
 *                      from-scratch CPU type-3 NUFFT (exp-of-semicircle kernel,
 *                      sigma=2) -- the same algorithm class as FINUFFT (third
 *                      party, unpinned, absent).  Used as the CPU baseline
 *                      ("CPU restatement, not FINUFFT") and validated against
 *                      sko_direct_cis in tests/.
 *
 * Parity pinning: see oracle/README.md.  Built by oracle/Makefile with
 * -ffp-contract=off so that a*b+c is two roundings, as in Julia.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

void sko_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int sko_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* direct Fourier summation, src/quadrature.jl:113-128                        */
/* cispi(2*no*x): the product 2*no[k]*x[j] is rounded once, then reduced      */
/* exactly mod 2 (as Julia's cispi does) before sin/cos.                      */
/* ------------------------------------------------------------------------- */
static inline void cispi_d(double t, double *c, double *s) {
  /* t in "half turns": cos(pi t), sin(pi t) with exact argument reduction */
  double r = t - 2.0 * rint(t * 0.5); /* exact: r in [-1,1] */
  sincos(M_PI * r, s, c);
  /* tidy exact cases */
  if (r == 0.5) { *c = 0.0; *s = 1.0; }
  else if (r == -0.5) { *c = 0.0; *s = -1.0; }
  else if (r == 1.0 || r == -1.0) { *c = -1.0; *s = 0.0; }
}

void sko_direct_cis(int64_t M, const double *no, const double *buf_re,
                    const double *buf_im, int64_t N, const double *xs,
                    double *out_re, double *out_im) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) {
    double xj = xs[j];
    double ar = 0.0, ai = 0.0;
    for (int64_t k = 0; k < M; ++k) {
      double c, s;
      cispi_d(2.0 * no[k] * xj, &c, &s);
      double br = buf_re[k], bi = buf_im ? buf_im[k] : 0.0;
      ar += br * c - bi * s;
      ai += br * s + bi * c;
    }
    out_re[j] = ar;
    out_im[j] = ai;
  }
}

/* direct Bessel summation, src/quadrature.jl:145-160 (integer order nu) */
void sko_direct_bessel(int nu, int64_t M, const double *no, const double *buf,
                       int64_t N, const double *xs, double *out) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) {
    double xj = xs[j];
    double acc = 0.0;
    for (int64_t k = 0; k < M; ++k) acc += buf[k] * jn(nu, 2.0 * M_PI * no[k] * xj);
    out[j] = acc;
  }
}

/* ------------------------------------------------------------------------- */
/* Gauss-Jacobi rule with weight (1-x)^0 (1+x)^p on [-1,1], ascending nodes.  */
/* p == 0 gives Gauss-Legendre.  (src/quadrature.jl:36-42)                    */
/* ------------------------------------------------------------------------- */
typedef long double ld;

/* P_n^{(0,b)}(x) and P_{n-1}^{(0,b)}(x) by the three-term recurrence */
static void jacobi_pn(int n, ld b, ld x, ld *pn, ld *pnm1) {
  ld p0 = 1.0L;
  ld p1 = 0.5L * ((b + 2.0L) * x - b); /* (a-b)/2 + (a+b+2)x/2 with a=0 */
  if (n == 0) { *pn = p0; *pnm1 = 0.0L; return; }
  for (int k = 1; k < n; ++k) {
    ld kk = (ld)k;
    ld s = 2.0L * kk + b; /* 2k+a+b */
    ld c1 = 2.0L * (kk + 1.0L) * (kk + b + 1.0L) * s;
    ld c2 = (s + 1.0L) * ((s + 2.0L) * s * x - b * b);
    ld c3 = 2.0L * kk * (kk + b) * (s + 2.0L);
    ld p2 = (c2 * p1 - c3 * p0) / c1;
    p0 = p1;
    p1 = p2;
  }
  *pn = p1;
  *pnm1 = p0;
}

static inline ld jacobi_dpn(int n, ld b, ld x, ld pn, ld pnm1) {
  /* (2n+a+b)(1-x^2) P_n' = n[a-b-(2n+a+b)x] P_n + 2(n+a)(n+b) P_{n-1}, a=0 */
  ld nn = (ld)n, s = 2.0L * nn + b;
  return (nn * (-b - s * x) * pn + 2.0L * nn * (nn + b) * pnm1) / (s * (1.0L - x * x));
}

/* returns 0 on success, nonzero if the nodes failed to come out strictly ascending */
int sko_gauss_rule(int n, double p, double *xout, double *wout) {
  ld b = (ld)p;
  ld piL = 3.14159265358979323846264338327950288L;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    /* i-th node ascending; k counts from the x=+1 end */
    int k = n - i;
    /* interior asymptotic guess: theta_k ~ (2k - 1/2 + a) pi / (2n + a + b + 1), a = 0 */
    ld th = (2.0L * k - 0.5L) * piL / (2.0L * n + b + 1.0L);
    ld x = cosl(th);
    for (int it = 0; it < 100; ++it) {
      ld pn, pnm1;
      jacobi_pn(n, b, x, &pn, &pnm1);
      ld dp = jacobi_dpn(n, b, x, pn, pnm1);
      ld dx = pn / dp;
      x -= dx;
      if (fabsl(dx) <= 4.0L * 1.0842e-19L * (1.0L + fabsl(x))) break;
    }
    ld pn, pnm1;
    jacobi_pn(n, b, x, &pn, &pnm1);
    ld dp = jacobi_dpn(n, b, x, pn, pnm1);
    ld w = powl(2.0L, b + 1.0L) / ((1.0L - x * x) * dp * dp);
    xout[i] = (double)x;
    wout[i] = (double)w;
  }
  for (int i = 1; i < n; ++i)
    if (!(xout[i] > xout[i - 1])) bad = 1;
  return bad;
}

/* ------------------------------------------------------------------------- */
/* CPU type-3 NUFFT (baseline + mid-size checker).                            */
/*   f_j = sum_k c_k exp(+i 2 pi x_j w_k),  j = 0..N-1                        */
/* Exp-of-semicircle kernel phi(z) = exp(beta (sqrt(1-z^2) - 1)), width nsp   */
/* grid cells, upsampling sigma = 2 in both the spread and the inner type-2.  */
/* ------------------------------------------------------------------------- */
typedef double complex dc;

static int64_t next235even(int64_t n) {
  if (n <= 2) return 2;
  if (n % 2) ++n;
  for (;; n += 2) {
    int64_t m = n;
    while (m % 2 == 0) m /= 2;
    while (m % 3 == 0) m /= 3;
    while (m % 5 == 0) m /= 5;
    if (m == 1) return n;
  }
}

/* in-place iterative-free recursive mixed radix (2,3,5) DIT FFT, sign = +1 */
static void fft_rec(int64_t n, int64_t stride, const dc *in, dc *out, const dc *tw, int64_t ntw) {
  if (n == 1) { out[0] = in[0]; return; }
  int r = (n % 2 == 0) ? 2 : (n % 3 == 0) ? 3 : 5;
  int64_t m = n / r;
  for (int q = 0; q < r; ++q) fft_rec(m, stride * r, in + q * stride, out + q * m, tw, ntw);
  int64_t tstep = ntw / n;
  dc tmp[5];
  for (int64_t k = 0; k < m; ++k) {
    for (int q = 0; q < r; ++q) tmp[q] = out[q * m + k] * tw[(q * k * tstep) % ntw];
    for (int s = 0; s < r; ++s) {
      dc acc = 0;
      for (int q = 0; q < r; ++q) acc += tmp[q] * tw[((int64_t)q * s * m * tstep) % ntw];
      out[s * m + k] = acc;
    }
  }
}

static void fft_forward_plus(int64_t n, const dc *in, dc *out) {
  dc *tw = (dc *)malloc(sizeof(dc) * n);
  for (int64_t k = 0; k < n; ++k) {
    double a = 2.0 * M_PI * (double)k / (double)n;
    tw[k] = cos(a) + I * sin(a);
  }
  fft_rec(n, 1, in, out, tw, n);
  free(tw);
}

static inline double es_kernel(double z, double beta) {
  double t = 1.0 - z * z;
  return t > 0.0 ? exp(beta * (sqrt(t) - 1.0)) : 0.0;
}

/* phihat(xi) = int_{-1}^{1} phi(z) cos(xi z) dz by Gauss-Legendre (nq nodes on [0,1]) */
typedef struct { int nq; double *z, *w; double beta; } phihat_t;
static void phihat_init(phihat_t *ph, double beta, int nsp) {
  int n = 2 * (int)(2 + 1.5 * nsp); /* enough for |xi| <= pi*nsp/4 at beta = 2.3 nsp */
  double *x = (double *)malloc(sizeof(double) * n), *w = (double *)malloc(sizeof(double) * n);
  sko_gauss_rule(n, 0.0, x, w);
  ph->nq = n / 2;
  ph->z = (double *)malloc(sizeof(double) * ph->nq);
  ph->w = (double *)malloc(sizeof(double) * ph->nq);
  for (int i = 0; i < ph->nq; ++i) {
    ph->z[i] = x[n / 2 + i];
    ph->w[i] = 2.0 * w[n / 2 + i] * es_kernel(ph->z[i], beta);
  }
  ph->beta = beta;
  free(x); free(w);
}
static inline double phihat_eval(const phihat_t *ph, double xi) {
  double acc = 0.0;
  for (int i = 0; i < ph->nq; ++i) acc += ph->w[i] * cos(xi * ph->z[i]);
  return acc;
}
static void phihat_free(phihat_t *ph) { free(ph->z); free(ph->w); }

/* s: interleaved complex strengths (re,im); out: interleaved complex */
int sko_nufft1d3(int64_t M, const double *w, const double *s, int64_t N, const double *x,
                 double *out, double eps) {
  if (M <= 0 || N <= 0) return 0;
  int nsp = (int)ceil(-log10(eps / 10.0));
  if (nsp < 2) nsp = 2;
  if (nsp > 16) nsp = 16;
  const double sigma = 2.0;
  const double beta = 2.30 * nsp;
  double wmin = w[0], wmax = w[0], xmin = x[0], xmax = x[0];
  for (int64_t k = 1; k < M; ++k) { if (w[k] < wmin) wmin = w[k]; if (w[k] > wmax) wmax = w[k]; }
  for (int64_t j = 1; j < N; ++j) { if (x[j] < xmin) xmin = x[j]; if (x[j] > xmax) xmax = x[j]; }
  double wc = 0.5 * (wmin + wmax), X = 0.5 * (wmax - wmin);
  double D = 0.5 * (xmin + xmax), S = 0.5 * (xmax - xmin);
  /* guard degenerate widths so that the space-bandwidth product is >= O(1) */
  if (X == 0.0 && S == 0.0) { X = 1.0; S = 1.0; }
  else if (X == 0.0) X = 1.0 / (4.0 * S);
  else if (S * X < 0.0625) S = 0.0625 / X;
  double hu = 1.0 / (2.0 * sigma * S);          /* spread-grid spacing in w */
  int64_t nf = (int64_t)ceil(2.0 * X / hu) + nsp + 2;
  if (nf % 2) ++nf;
  if (nf < 2 * nsp) nf = 2 * nsp;
  int64_t nf2 = next235even((int64_t)ceil(sigma * (double)nf));
  dc *b = (dc *)calloc((size_t)nf, sizeof(dc));
  dc *d = (dc *)calloc((size_t)nf2, sizeof(dc));
  dc *g = (dc *)malloc(sizeof(dc) * (size_t)nf2);
  if (!b || !d || !g) { free(b); free(d); free(g); return -1; }
  phihat_t ph;
  phihat_init(&ph, beta, nsp);
  const double half = 0.5 * nsp;
  /* step A: pre-phase and spread; sources are split over the threads, each with a private grid */
  {
    int nth = sko_num_threads();
    if (nth > 32) nth = 32;
    if (M < 4096) nth = 1;
    dc *priv = nth > 1 ? (dc *)calloc((size_t)nf * (size_t)nth, sizeof(dc)) : NULL;
    if (nth > 1 && !priv) nth = 1;
#pragma omp parallel num_threads(nth)
    {
#ifdef _OPENMP
      int tid = omp_get_thread_num();
#else
      int tid = 0;
#endif
      dc *bb = nth > 1 ? priv + (size_t)tid * (size_t)nf : b;
#pragma omp for schedule(static)
      for (int64_t k = 0; k < M; ++k) {
        double u = w[k] - wc;
        double ang = 2.0 * M_PI * (u * D);
        dc c = (s[2 * k] + I * s[2 * k + 1]) * (cos(ang) + I * sin(ang));
        double pos = u / hu + 0.5 * (double)nf;
        int64_t l0 = (int64_t)ceil(pos - half);
        for (int i = 0; i < nsp; ++i) {
          int64_t l = l0 + i;
          if (l < 0 || l >= nf) continue;
          bb[l] += c * es_kernel(((double)l - pos) / half, beta);
        }
      }
    }
    if (nth > 1) {
#pragma omp parallel for schedule(static)
      for (int64_t l = 0; l < nf; ++l) {
        dc acc = 0;
        for (int t = 0; t < nth; ++t) acc += priv[(size_t)t * (size_t)nf + l];
        b[l] = acc;
      }
      free(priv);
    }
  }
  /* step B: deconvolve the modes n = l - nf/2, modulate by (-1)^n (output shift nf2/2), zero-pad */
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < nf; ++l) {
    int64_t n = l - nf / 2;
    double q = (2.0 / nsp) / phihat_eval(&ph, M_PI * nsp * (double)n / (double)nf2);
    if (n & 1) q = -q;
    int64_t idx = n >= 0 ? n : n + nf2;
    d[idx] = b[l] * q;
  }
  fft_forward_plus(nf2, d, g);
  /* step C: interpolate at the targets, deconvolve the spread kernel, post-phase */
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) {
    double v = x[j] - D;
    double y = hu * v * (double)nf2 + 0.5 * (double)nf2;
    int64_t l0 = (int64_t)ceil(y - half);
    dc acc = 0;
    for (int i = 0; i < nsp; ++i) {
      int64_t l = l0 + i;
      l = ((l % nf2) + nf2) % nf2;
      acc += g[l] * es_kernel(((double)(l0 + i) - y) / half, beta);
    }
    double q = (2.0 / nsp) / phihat_eval(&ph, M_PI * nsp * hu * v);
    double t = wc * x[j];
    t -= rint(t);
    double ang = 2.0 * M_PI * t;
    dc f = acc * q * (cos(ang) + I * sin(ang));
    out[2 * j] = creal(f);
    out[2 * j + 1] = cimag(f);
  }
  phihat_free(&ph);
  free(b); free(d); free(g);
  return 0;
}
