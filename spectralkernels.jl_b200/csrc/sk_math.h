// sk_math.h -- per-element arithmetic of the B200 K(r) path, written once as host/device inline
// functions.  The CUDA kernels in sk_kernels.cuh are thin launch wrappers around these; the
// host-emulation harness in tests/emul/ compiles the same functions with g++ so the index
// arithmetic can be checked on a machine without a GPU.
//
// Type-3 NUFFT geometry (one per sub-interval, shared by the m- and 2m-node rules):
//   f_j = sum_k c_k exp(2 pi i w_k r_j)        (src/utils.jl:10 with s_j = 2 pi r_j)
//   w = wc + u, r = D + v:   w r = wc r + u D + u v
//   sources are spread at pos = u * inv_hu on a grid of nf cells (modes n = -nf/2 .. nf/2-1),
//   the modes are deconvolved and zero-padded to nf2 >= 2 nf, one FFT of size nf2 gives samples of
//   B(t) = sum_n b_n exp(2 pi i n t) at t = (l - nf2/2)/nf2, and the targets interpolate at
//   y = v * kappa (kappa = nf2 / inv_hu) with the same exp-of-semicircle kernel.
// Positions (pos, y) and the phases (wc r, u D) are carried as unevaluated sums of two doubles so
// that the O(eps * space-bandwidth product) position error of a plain-double NUFFT does not
// appear: the result agrees with the direct sum to ~1e-14 * sum|c_k| instead of ~1e-11.
#pragma once
#include "sk_plan.h"

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SK_HD __host__ __device__ __forceinline__
#else
#define SK_HD inline
#endif

struct alignas(16) sk_cplx { double x, y; };

#define SK_KMAX 128          // largest k (sub-panels per NUFFT) supported by the device generator
#define SK_NPARAM_MAX 8

struct SkGeom {
  double wc;          // centre of the source interval
  double D;           // centre of the target interval (0 when the targets start near the origin)
  double inv_hu;      // spread-grid cells per unit frequency
  double kap_hi;      // fine-grid cells per unit distance, kappa = nf2 / inv_hu, as hi + lo
  double kap_lo;
  double t_cell;      // (pi w / nf2) / ximax : normalised deconvolution argument per fine-grid cell / mode index
  long long nf;       // spread-grid size (even)
  long long nf2;      // FFT size (2^k or 3*2^(k-1))
};

struct SkPanelSpec {     // updatequadbufs! (src/quadrature.jl:49-95) for one sub-interval
  int m, k;
  int origin_jacobi;     // p != 0 && a == 0: Gauss-Jacobi on the first sub-panel, f(no) only
  int weight_in_f;       // else-branch of src/quadrature.jl:240-247: integrand is w^p [log w] f(w)
  int logw;
  int family, deriv, nparam;
  int variant;           // integrand: 0 f;  1 f + w log(w) f'(w);  2 w log(w) f(w)   (the two integrands of the
  int _pad;              //   integration by parts of the log-weighted origin sub-interval, src/quadrature.jl:192, :198)
  double p;
  double jac_scale;      // bmad2[0]^(p+1), src/quadrature.jl:69,73
  double params[SK_NPARAM_MAX];
  double bmad2[SK_KMAX];
  double bpad2[SK_KMAX];
};

// ---- small portability layer ---------------------------------------------------------------------
SK_HD double sk_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}
SK_HD double sk_mul(double a, double b) {  // a*b rounded once, never contracted into an FMA
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;   // host builds use -ffp-contract=off
#endif
}
SK_HD double sk_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
SK_HD void sk_sincospi(double x, double *s, double *c) {
#if defined(__CUDA_ARCH__)
  sincospi(x, s, c);
#else
  const double pi = 3.14159265358979323846;
  double r = x - 2.0 * rint(0.5 * x);
  *s = sin(pi * r);
  *c = cos(pi * r);
#endif
}

// a + b = s + e exactly
SK_HD void sk_two_sum(double a, double b, double *s, double *e) {
  double ss = sk_add(a, b);
  double bb = sk_add(ss, -a);
  *e = sk_add(sk_add(a, -sk_add(ss, -bb)), sk_add(b, -bb));
  *s = ss;
}

// fractional part of a*b (in cycles), |result| <= 1/2 + tiny, with the product formed exactly
SK_HD double sk_frac_prod(double a, double b, double extra) {
  double p = sk_mul(a, b);
  double e = sk_fma(a, b, -p);
  return (p - rint(p)) + (e + extra);
}

// ---- exp-of-semicircle kernel --------------------------------------------------------------------
SK_HD double sk_es_direct(double z, double beta) {
  double t = 1.0 - z * z;
  return t > 0.0 ? exp(beta * (sqrt(t) - 1.0)) : 0.0;
}

// all W taps at offset s = 2x, x in [-1/2, 1/2): Horner on the even / odd parts in s^2
template <int W>
SK_HD void sk_es_taps(const SkEsPlan &P, double s, double *taps) {
  const double s2 = s * s;
#pragma unroll
  for (int i = 0; i < W / 2; ++i) {
    double e = P.E[i][SK_NC / 2 - 1];
    double o = P.O[i][SK_NC / 2 - 1];
#pragma unroll
    for (int q = SK_NC / 2 - 2; q >= 0; --q) {
      e = sk_fma(e, s2, P.E[i][q]);
      o = sk_fma(o, s2, P.O[i][q]);
    }
    taps[i] = sk_fma(s, o, e);
    taps[W - 1 - i] = sk_fma(-s, o, e);
  }
}

// (2/w)/phihat(xi) at t = xi/ximax in [0,1], by Clenshaw in tau = 2 t^2 - 1
SK_HD double sk_deconv(const SkEsPlan &P, double t) {
  const double tau2 = sk_fma(4.0 * t, t, -2.0);  // 2*tau
  double b1 = 0.0, b2 = 0.0;
  for (int j = P.nq - 1; j >= 1; --j) {
    double b0 = sk_fma(tau2, b1, P.qc[j] - b2);
    b2 = b1;
    b1 = b0;
  }
  return sk_fma(0.5 * tau2, b1, P.qc[0] - b2);
}

// ---- spectral-density families (device-evaluated integrands) --------------------------------------
// family ids as in include/spectralkernels_b200.h.  deriv = 0: S; j>=1: dS/dparams[j-1].
SK_HD double sk_sdf_eval(int family, int deriv, const double *q, double w) {
  if (family == 1) {  // Matern: phi (rho^2 + w^2)^(-nu - d/2), scripts/matern_pair.jl:17
    const double phi = q[0], rho = q[1], nu = q[2], d = q[3];
    const double base = sk_add(sk_mul(rho, rho), sk_mul(w, w));
    const double ex = -nu - d / 2;
    if (deriv == 0) return phi * pow(base, ex);
    if (deriv == 1) return pow(base, ex);
    if (deriv == 2) return phi * ex * pow(base, ex - 1.0) * 2.0 * rho;
    if (deriv == 3) return -phi * pow(base, ex) * log(base);
    return 0.0;
  }
  if (family == 2) {  // exponential: phi exp(-alpha |w|), test/derivatives/jacobian.jl:5
    const double phi = q[0], al = q[1];
    const double e = exp(-al * fabs(w));
    if (deriv == 0) return phi * e;
    if (deriv == 1) return e;
    if (deriv == 2) return -phi * fabs(w) * e;
    return 0.0;
  }
  return 0.0;
}

// dS/dw of the density itself (the `df` keyword of AdaptiveKernelConfig, needed by logw = true: src/quadrature.jl:192)
SK_HD double sk_sdf_dw(int family, const double *q, double w) {
  if (family == 1) {
    const double phi = q[0], rho = q[1], nu = q[2], d = q[3];
    const double base = sk_add(sk_mul(rho, rho), sk_mul(w, w));
    const double ex = -nu - d / 2;
    return phi * ex * pow(base, ex - 1.0) * 2.0 * w;
  }
  if (family == 2) {
    const double phi = q[0], al = q[1];
    return -al * (w > 0 ? 1.0 : (w < 0 ? -1.0 : 0.0)) * phi * exp(-al * fabs(w));
  }
  return 0.0;
}
// the integrand the panel spec asks for (SkPanelSpec::variant)
SK_HD double sk_integrand(const SkPanelSpec &S, double w) {
  const double f = sk_sdf_eval(S.family, S.deriv, S.params, w);
  if (S.variant == 0) return f;
  const double wl = w * log(w);
  if (S.variant == 1) return f + wl * sk_sdf_dw(S.family, S.params, w);
  return wl * f;
}

// node and (real) strength of source `idx` of rule `rule` (0: m nodes per sub-panel, 1: 2m),
// following updatequadbufs! operation by operation (src/quadrature.jl:61-92)
SK_HD void sk_gen_source(const SkPanelSpec &S, int rule, long long idx, const double *leg_no, const double *leg_wt,
                         const double *jac_no, const double *jac_wt, double *no_out, double *buf_out) {
  const int mm = rule == 0 ? S.m : 2 * S.m;
  const int i = (int)(idx / mm);  // sub-panel
  const int j = (int)(idx - (long long)i * mm);
  const double bmad2 = S.bmad2[i], bpad2 = S.bpad2[i];
  double no, buf;
  if (S.origin_jacobi && i == 0) {
    no = sk_add(sk_mul(bmad2, jac_no[j]), bpad2);                                  // :68,72
    buf = sk_mul(sk_mul(jac_wt[j], S.jac_scale), sk_integrand(S, no));  // :69,73
  } else {
    no = sk_add(sk_mul(bmad2, leg_no[j]), bpad2);                                  // :85,89
    const double f = sk_integrand(S, no);
    const double wb = sk_mul(leg_wt[j], bmad2);
    if (S.weight_in_f) {
      // integrand w -> w^p * (logw ? log(w) : 1) * f(w), quadrature rule power 0 (src/quadrature.jl:242, :86)
      double g = S.p == 0.0 ? 1.0 : pow(no, S.p);
      g = sk_mul(g, S.logw ? log(no) : 1.0);
      g = sk_mul(g, f);
      buf = sk_mul(sk_mul(wb, 1.0), g);
    } else {
      // Legendre sub-panels of an origin interval: wt*bmad2 * no^p * f(no) (src/quadrature.jl:86,90)
      const double np_ = S.p == 0.0 ? 1.0 : pow(no, S.p);
      buf = sk_mul(sk_mul(wb, np_), f);
    }
  }
  *no_out = no;
  *buf_out = buf;
}

// ---- source side: position on the spread grid and pre-phase ---------------------------------------
SK_HD void sk_source_prep(const SkGeom &G, double no, double c_re, double c_im, double *pos_hi, double *pos_lo,
                          double *o_re, double *o_im) {
  double u_hi, u_lo;
  sk_two_sum(no, -G.wc, &u_hi, &u_lo);
  const double ph = sk_mul(u_hi, G.inv_hu);
  const double pl = sk_fma(u_hi, G.inv_hu, -ph) + u_lo * G.inv_hu;
  *pos_hi = ph;
  *pos_lo = pl;
  if (G.D != 0.0) {
    const double fr = sk_frac_prod(u_hi, G.D, u_lo * G.D);
    double sn, cs;
    sk_sincospi(2.0 * fr, &sn, &cs);
    *o_re = c_re * cs - c_im * sn;
    *o_im = c_re * sn + c_im * cs;
  } else {
    *o_re = c_re;
    *o_im = c_im;
  }
}

// ---- spread (gather) + mode deconvolution + zero-pad: the value of FFT-input element j -------------
// pos_hi must be ascending.  cs: complex (pre-phased) strengths.
SK_HD void sk_spread_mode(const SkEsPlan &P, const SkGeom &G, long long j, const double *pos_hi, const double *pos_lo,
                          const sk_cplx *cs, long long M, double *o_re, double *o_im) {
  const long long n = (j < G.nf2 / 2) ? j : j - G.nf2;  // signed mode index
  if (n < -(G.nf / 2) || n >= G.nf / 2) {
    *o_re = 0.0;
    *o_im = 0.0;
    return;
  }
  const double ctr = (double)n;
  const double half = 0.5 * P.w;
  const double lo_edge = ctr - half;
  // first source with pos_hi > ctr - w/2 - guard (guard covers |pos_lo|)
  long long a = 0, b = M;
  while (a < b) {
    long long mid = (a + b) >> 1;
    if (pos_hi[mid] < lo_edge - 1e-6) a = mid + 1; else b = mid;
  }
  double ar = 0.0, ai = 0.0;
  const double hi_edge = ctr + half + 1e-6;
  const double inv_half = 1.0 / half;
  for (long long k = a; k < M && pos_hi[k] <= hi_edge; ++k) {
    const double z = ((ctr - pos_hi[k]) - pos_lo[k]) * inv_half;
    const double wgt = sk_es_direct(z, P.beta);
    const sk_cplx c = cs[k];
    ar = sk_fma(wgt, c.x, ar);
    ai = sk_fma(wgt, c.y, ai);
  }
  double q = sk_deconv(P, G.t_cell * fabs(ctr));
  if (n & 1) q = -q;  // shifts the FFT output by nf2/2 so the used window is contiguous
  *o_re = ar * q;
  *o_im = ai * q;
}

// ---- target side ----------------------------------------------------------------------------------
struct SkTargetCoord {
  long long l0;   // first fine-grid index of the w-wide window (already shifted by nf2/2, clamped)
                  // (the cell centre s = 0 sits at y = l0 - nf2/2 + w/2 - 1/2)
  double s;       // 2x, x in [-1/2, 1/2) the offset inside the window
  double yabs;    // |y| for the deconvolution argument
};

template <int W>
SK_HD SkTargetCoord sk_target_coord(const SkGeom &G, double r) {
  double v_hi = r, v_lo = 0.0;
  if (G.D != 0.0) sk_two_sum(r, -G.D, &v_hi, &v_lo);
  const double y_hi = sk_mul(v_hi, G.kap_hi);
  const double y_lo = sk_fma(v_hi, G.kap_hi, -y_hi) + sk_fma(v_hi, G.kap_lo, v_lo * G.kap_hi);
  const double c0 = ceil(y_hi - 0.5 * W);
  const double x0 = (c0 - y_hi) - y_lo;          // in [-W/2, -W/2 + 1)
  SkTargetCoord t;
  t.s = 2.0 * (x0 + (0.5 * W - 0.5));
  // window start in 32-bit arithmetic (nf2 < 2^31); one unsigned min clamps both ends (a negative start
  // wraps to a huge unsigned).  Targets inside the range the geometry was built for never hit the clamp.
  const unsigned int hi_ok = (unsigned int)(G.nf2 - W);
  unsigned int l0u = (unsigned int)((int)c0 + (int)(G.nf2 / 2));
  l0u = l0u < hi_ok ? l0u : hi_ok;
  t.l0 = (long long)l0u;
  t.yabs = fabs(y_hi);
  return t;
}

// post-phase exp(2 pi i wc r) with the product wc*r formed exactly
SK_HD void sk_post_phase(const SkGeom &G, double r, double *sn, double *cs) {
  const double fr = sk_frac_prod(G.wc, r, 0.0);
  sk_sincospi(2.0 * fr, sn, cs);
}

// Interpolate NR interleaved grids (layout grid[l*NR + rule]) at distance r and apply
// the target-side deconvolution and the post-phase: out = f_rule(r) as (re, im).
template <int W, int NR>
SK_HD void sk_interp_point(const SkEsPlan &P, const SkGeom &G, double r, const sk_cplx *grid, double *out_re,
                           double *out_im) {
  const SkTargetCoord t = sk_target_coord<W>(G, r);
  double taps[W];
  sk_es_taps<W>(P, t.s, taps);
  double ar[NR], ai[NR];
#pragma unroll
  for (int q = 0; q < NR; ++q) ar[q] = ai[q] = 0.0;
  const sk_cplx *g = grid + (size_t)t.l0 * NR;
#pragma unroll
  for (int i = 0; i < W; ++i) {
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      const sk_cplx gv = g[i * NR + q];
      ar[q] = sk_fma(taps[i], gv.x, ar[q]);
      ai[q] = sk_fma(taps[i], gv.y, ai[q]);
    }
  }
  const double qf = sk_deconv(P, G.t_cell * t.yabs);
  double sn, cs;
  sk_post_phase(G, r, &sn, &cs);
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    out_re[q] = qf * (ar[q] * cs - ai[q] * sn);
    out_im[q] = qf * (ar[q] * sn + ai[q] * cs);
  }
}

// ---- lean sincos of 2 pi f for |f| <= 1/2 + tiny -----------------------------------------------------
// f = k/64 + g with k = rint(64 f): a 65-entry table (cos, sin)(2 pi k / 64), k = -32..32, and short Taylor
// series in theta = 2 pi g, |theta| <= pi/64 (sin through theta^7, cos through theta^8: truncation < 1e-16),
// combined by one complex rotation.  No branches or selects -- about 20 FP64 operations, against ~60
// instructions for the generic sincospi.  tab: sk_cplx[65] (shared or constant memory).
SK_HD void sk_sincos2pi_table_fill(sk_cplx *tab, int k /* 0..64 */) {
  double s, c;
  sk_sincospi((double)(k - 32) / 32.0, &s, &c);      // 2 pi (k-32)/64 = pi (k-32)/32
  tab[k].x = c;
  tab[k].y = s;
}
SK_HD void sk_sincos2pi(const sk_cplx *tab, double f, double *sn, double *cs) {
  const double kf = rint(64.0 * f);
  const double th = 6.283185307179586 * (f - kf * 0.015625);          // exact subtraction
  const double t2 = th * th;
  double ps = sk_fma(t2, -1.0 / 5040.0, 1.0 / 120.0);
  ps = sk_fma(ps, t2, -1.0 / 6.0);
  ps = sk_fma(ps * t2, th, th);                                        // sin(theta)
  double pc = sk_fma(t2, 1.0 / 40320.0, -1.0 / 720.0);
  pc = sk_fma(pc, t2, 1.0 / 24.0);
  pc = sk_fma(pc, t2, -0.5);
  pc = sk_fma(pc, t2, 1.0);                                            // cos(theta)
  const sk_cplx e = tab[(int)kf + 32];
  *cs = sk_fma(e.x, pc, -e.y * ps);
  *sn = sk_fma(e.y, pc, e.x * ps);
}

// ---- cell polynomials ---------------------------------------------------------------------------------
// All targets whose w-wide window starts at the same fine-grid index ("cell") see the same w grid
// values, so  sum_i tap_i(s) g_i = sum_q C_q s^q  with coefficients that depend on the cell only:
//   C_{2q'}   = sum_{i<w/2} E_i[q'] (g_i + g_{w-1-i}),   C_{2q'+1} = sum_{i<w/2} O_i[q'] (g_i - g_{w-1-i}).
// With many targets per cell (N >> nf2: 1e7 targets on ~6.6e4 used cells) the per-target work drops
// from w tap polynomials + w complex MACs to one degree-(SK_NC-1) Horner per component.
// g: window of W points, `stride` doubles between consecutive points, component `comp` selected by offset.
template <int W>
SK_HD double sk_cell_coef(const double *E, const double *O, const double *g, int stride, int q) {
  const int qh = q >> 1;
  const bool odd = q & 1;
  const double *T = odd ? O : E;   // flattened [W/2][SK_NC/2]
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < W / 2; ++i) {
    const double a = g[i * stride], b = g[(W - 1 - i) * stride];
    acc = sk_fma(T[i * (SK_NC / 2) + qh], odd ? (a - b) : (a + b), acc);
  }
  return acc;
}

// Horner in s on NCOMP interleaved components; coef layout [SK_NC][NCOMP]
template <int NCOMP>
SK_HD void sk_cell_horner(const double *coef, double s, double *out) {
#pragma unroll
  for (int c = 0; c < NCOMP; ++c) out[c] = coef[(SK_NC - 1) * NCOMP + c];
#pragma unroll
  for (int q = SK_NC - 2; q >= 0; --q) {
#pragma unroll
    for (int c = 0; c < NCOMP; ++c) out[c] = sk_fma(out[c], s, coef[q * NCOMP + c]);
  }
}

// The target-side deconvolution factor q(y) = (2/w)/phihat(pi w y / nf2) changes by ~1e-4 relative across
// one cell, so inside a cell it is a cubic in s to ~1e-17 (interpolated at the 4 Chebyshev nodes).  Folding
// that cubic into the cell polynomial (a truncated polynomial product: the dropped degree-16..18 terms are
// ~1e-4 * 1e-16) makes the deconvolution free per target.  ymid = y at s = 0 (the cell centre), y = ymid - s/2.
// q[0..3]: values at the Chebyshev nodes s = cos(pi/8), cos(3pi/8), -cos(3pi/8), -cos(pi/8)  ->  monomial
// coefficients a[0..3] of the interpolating cubic in s
SK_HD void sk_cheb4_to_monomial(const double *q, double *a) {
  const double n0 = 0.9238795325112867, n1 = 0.3826834323650898;   // cos(pi/8), cos(3 pi/8)
  const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  // Chebyshev coefficients of the interpolant: c_j = (2/4) sum_k q_k T_j(s_k), c_0 halved
  const double t2a = 2.0 * n0 * n0 - 1.0, t2b = 2.0 * n1 * n1 - 1.0;               // T2 at +-n0, +-n1
  const double t3a = (4.0 * n0 * n0 - 3.0) * n0, t3b = (4.0 * n1 * n1 - 3.0) * n1; // T3 at n0, n1 (odd)
  const double c0 = 0.25 * ((q0 + q3) + (q1 + q2));
  const double c1 = 0.5 * (n0 * (q0 - q3) + n1 * (q1 - q2));
  const double c2 = 0.5 * (t2a * (q0 + q3) + t2b * (q1 + q2));
  const double c3 = 0.5 * (t3a * (q0 - q3) + t3b * (q1 - q2));
  a[0] = c0 - c2;          // T0 = 1, T1 = s, T2 = 2 s^2 - 1, T3 = 4 s^3 - 3 s
  a[1] = c1 - 3.0 * c3;
  a[2] = 2.0 * c2;
  a[3] = 4.0 * c3;
}
SK_HD void sk_cell_deconv_cubic(const SkEsPlan &P, const SkGeom &G, double ymid, double *a) {
  const double n0 = 0.9238795325112867, n1 = 0.3826834323650898;
  double q[4];
  q[0] = sk_deconv(P, G.t_cell * fabs(ymid - 0.5 * n0));
  q[1] = sk_deconv(P, G.t_cell * fabs(ymid - 0.5 * n1));
  q[2] = sk_deconv(P, G.t_cell * fabs(ymid + 0.5 * n1));
  q[3] = sk_deconv(P, G.t_cell * fabs(ymid + 0.5 * n0));
  sk_cheb4_to_monomial(q, a);
}

// in-place product of one coefficient column (stride doubles apart) with the cubic a, truncated at SK_NC
SK_HD void sk_cell_fold(double *col, int stride, const double *a) {
#pragma unroll
  for (int q = SK_NC - 1; q >= 0; --q) {
    double v = a[0] * col[q * stride];
    if (q >= 1) v = sk_fma(a[1], col[(q - 1) * stride], v);
    if (q >= 2) v = sk_fma(a[2], col[(q - 2) * stride], v);
    if (q >= 3) v = sk_fma(a[3], col[(q - 3) * stride], v);
    col[q * stride] = v;
  }
}

// ---- convergence predicate, src/adaptive.jl:222-233 -------------------------------------------------
// trunc_a, trunc_num are computed on the host (they do not depend on the target).
SK_HD double sk_trunc_err(double trunc_a, double trunc_num, double xpow, double x, int criteria_panel) {
  if (criteria_panel) return 0.0;  // src/adaptive.jl:186
  const double xp = (xpow == 1.0) ? x : pow(x, xpow);
  const double t2 = trunc_num / sk_mul(6.283185307179586, xp);   // 2pi*x^((dim+1)/2)
  // Julia's min(a, b) propagates a NaN operand (C's fmin drops it): a NaN bound then fails `trunc_err < tol`
  // in sk_converged, i.e. the target stays active, exactly as in src/adaptive.jl:185-197
  if (trunc_a != trunc_a || t2 != t2) return trunc_a + t2;
  return fmin(trunc_a, t2);
}
SK_HD bool sk_converged(double trunc_err, double panel_k, double tau, int criteria) {
  // (criteria == :panel || trunc_err < tol) && (criteria == :tails || abs(panel_k) < tol)
  const bool c1 = (criteria == 0) || (trunc_err < tau);
  const bool c2 = (criteria == 1) || (fabs(panel_k) < tau);
  return c1 && c2;
}
