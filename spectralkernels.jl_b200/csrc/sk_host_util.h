// sk_host_util.h -- host-side geometry / panel bookkeeping shared by the C-ABI translation unit and
// the host-emulation test harness.
#pragma once
#include "sk_hankel.h"
#include "sk_math.h"

#include <cmath>
#include <cstring>

// Geometry of one type-3 transform: sources in [w_lo, w_hi], targets in [r_lo, r_hi].
// sigma = 2 for both the spread and the inner type-2 step.
// Returns 0, or -1 if the grid would be unreasonably large.
static inline int sk_make_geom(const SkEsPlan &P, double w_lo, double w_hi, double r_lo, double r_hi, SkGeom *G,
                               bool force_center = false, bool pow2_only = false) {
  const double sigma = 2.0;
  const double PI = 3.14159265358979323846;
  double X = 0.5 * (w_hi - w_lo);
  G->wc = 0.5 * (w_lo + w_hi);
  double S;
  // centre the targets only when they form a narrow band away from the origin; otherwise keep
  // D = 0 (no pre-phase, r - D exact) -- the adaptive driver's targets always start near 0.
  if ((r_lo > 0.5 * r_hi || force_center) && r_lo > 0.0) {
    G->D = 0.5 * (r_lo + r_hi);
    S = 0.5 * (r_hi - r_lo);
  } else if (r_hi < 0.0 && r_hi < 0.5 * r_lo) {
    G->D = 0.5 * (r_lo + r_hi);
    S = 0.5 * (r_hi - r_lo);
  } else {
    G->D = 0.0;
    S = std::fmax(std::fabs(r_lo), std::fabs(r_hi));
  }
  // keep the space-bandwidth product >= O(1) so that the grids never degenerate
  if (!(X > 0.0) && !(S > 0.0)) { X = 1.0; S = 1.0; }
  else if (!(X > 0.0)) X = 1.0 / (16.0 * S);
  if (S * X < 0.0625) S = 0.0625 / X;
  G->inv_hu = 2.0 * sigma * S;
  const double cells = 2.0 * X * G->inv_hu;
  if (!(cells < 2.0e8)) return -1;
  long long nf = (long long)std::ceil(cells) + P.w + 2;
  if (nf & 1) ++nf;
  if (nf < 2 * P.w) nf = 2 * P.w;
  G->nf = nf;
  // FFT size: oversampling >= 1.999 (the deconvolution fit covers up to 2/1.996; the kernel's aliasing error is flat
  // around sigma = 2), rounded up to the next 2^k or 3*2^(k-1).  cuFFT is ~3x faster on such sizes than on sizes
  // with large 3^k 5^l factors, and -- as important for a fitting loop, where every hyperparameter vector moves the
  // panel ends a little -- the set of sizes is small, so cuFFT plans (tens of ms to create) are reused instead of
  // being rebuilt for every new geometry.  The adaptive driver's default geometry (nf = 131 090) lands on 2^18.
  const long long need = (long long)std::ceil(1.999 * (double)nf);
  long long pow2 = 2;
  while (pow2 < need) pow2 <<= 1;
  // (pow2_only: the octave groups of the Hankel transform come in many sizes that move with every hyperparameter
  //  vector of a fit; powers of two keep them to one cuFFT kernel family and ~10 plans -- the first use of a
  //  3*2^k size made cuFFT load another module, 0.6-1 s, in the middle of a fitting loop)
  G->nf2 = (!pow2_only && pow2 >= 8 && 3 * (pow2 / 4) >= need) ? 3 * (pow2 / 4) : pow2;
  const double n2 = (double)G->nf2;
  G->kap_hi = n2 / G->inv_hu;
  G->kap_lo = -std::fma(G->kap_hi, G->inv_hu, -n2) / G->inv_hu;
  G->t_cell = (PI * P.w / n2) / P.ximax;
  return 0;
}

// range(a, b, length=k+1) (src/quadrature.jl:56): Julia's StepRangeLen keeps the step in twice
// precision, i.e. element i is a + i (b-a)/k rounded once; long double reproduces that.
static inline void sk_fill_subpanels(double a, double b, int k, double *bmad2, double *bpad2) {
  const long double al = a, bl = b;
  double prev = a;
  for (int i = 1; i <= k; ++i) {
    double e = (i == k) ? b : (double)(al + (long double)i * ((bl - al) / (long double)k));
    bmad2[i - 1] = (e - prev) / 2;   // src/quadrature.jl:63,83
    bpad2[i - 1] = (e + prev) / 2;
    prev = e;
  }
}

// Plan of one nonuniform Hankel transform (sk_hankel.h): sources in [a, b], active targets in [r_lo, r_hi].
// Fills H and groups[0 .. H->ngroups) and returns the total number of sk_cplx grid entries, or -1 when the
// dyadic scheme does not apply (order not tabulated, too many levels, a grid too large): the caller then uses
// the direct Bessel summation.
static inline long long sk_hk_make_plan(const SkEsPlan &P, int nu, double a, double b, double r_lo, double r_hi,
                                        SkHankelPlan *H, SkHankelGroup *groups) {
  const double PI = 3.14159265358979323846;
  if (nu < 0 || nu > SK_HK_NUMAX || !(r_hi > 0.0) || !(b > a) || !(a >= 0.0) || !(r_lo > 0.0) || !(r_lo <= r_hi)) return -1;
  std::memset(H, 0, sizeof(*H));
  H->nu = nu;
  H->r_hi = r_hi;
  H->wT = SK_HK_ZL / (2.0 * PI * r_hi);
  const double phi = nu * PI / 2 + PI / 4;
  H->cphi = std::cos(phi);
  H->sphi = std::sin(phi);
  for (int n = 0; n < SK_HK_K; ++n) H->ratio[n] = (4.0 * nu * nu - (2.0 * n + 1.0) * (2.0 * n + 1.0)) / (8.0 * (n + 1.0));
  {
    // |a_n| / z^n <= 1e-17  <=>  z >= (|a_n| 1e17)^(1/n); made non-increasing from the top so that the terms a
    // target keeps are always a leading block 0 .. k-1
    double an = 1.0;
    H->zthr[0] = 1e300;
    for (int n = 1; n < SK_HK_K; ++n) {
      an *= H->ratio[n - 1];
      H->zthr[n] = std::pow(std::fabs(an) * 1e17, 1.0 / n);
    }
    for (int n = SK_HK_K - 2; n >= 1; --n) H->zthr[n] = std::fmax(H->zthr[n], H->zthr[n + 1]);
  }
  H->q_lo = sk_hk_level(H->wT, a);
  H->q_hi = sk_hk_level(H->wT, b);
  if (H->q_hi >= SK_HK_NLEV - 1) return -1;
  H->t_full = H->q_lo >= 2 ? H->q_lo - 2 : -1;
  H->t_last = H->q_hi - 2;                                     // may be negative: no asymptotic part at all
  const int t_need = sk_hk_octave(r_hi, r_lo);
  long long total = 0;
  int ng = 0;
  auto add = [&](const SkGeom *share, double w_lo, double rl, double rh, double w_ref, int q_cut, bool center) -> bool {
    if (ng >= SK_HK_NGRP) return false;
    SkHankelGroup &g = groups[ng];
    std::memset(&g, 0, sizeof(g));
    if (share) g.G = *share;
    else if (sk_make_geom(P, w_lo, b, rl, rh, &g.G, center, true) != 0) return false;
    g.w_ref = w_ref;
    g.q_cut = q_cut;
    g.q_from = q_cut;
    g.q_to = SK_HK_NLEV;
    g.shared = 0;
    g.grid_off = total;
    total += g.G.nf2 * (long long)(2 * SK_HK_K);
    ++ng;
    return true;
  };
  if (H->t_full >= 0) {
    const int tm = H->t_full < t_need ? H->t_full : t_need;
    const double rl = std::fmax(r_lo, std::ldexp(r_hi, -(tm + 1)));
    if (!add(nullptr, a, rl, r_hi, std::ldexp(H->wT, H->q_lo - 1), 0, false)) return -1;
  }
  const int t_first = H->t_full + 1, t_end = H->t_last < t_need ? H->t_last : t_need;      // octaves with own groups
  // the small octaves share one geometry and one w_ref (see SkHankelGroup): t_share .. t_end
  const int t_share = t_first > SK_HK_T_SHARE ? t_first : SK_HK_T_SHARE;
  const bool sharing = t_end - t_share >= 1;
  SkGeom Gs;
  double w_ref_s = 0.0;
  if (sharing) {
    w_ref_s = std::ldexp(H->wT, t_share + 1);
    const double rl = std::fmax(r_lo, std::ldexp(r_hi, -(t_end + 1)));
    if (sk_make_geom(P, w_ref_s, b, rl, std::ldexp(r_hi, -t_share), &Gs, false, true) != 0) return -1;
  }
  for (int t = t_first; t <= t_end; ++t) {
    if (sharing && t >= t_share) {
      if (!add(&Gs, 0.0, 0.0, 0.0, w_ref_s, t + 2, false)) return -1;
      SkHankelGroup &g = groups[ng - 1];
      g.q_to = (t == t_end) ? SK_HK_NLEV : t + 3;          // the deepest octave takes everything above its cut
      g.shared = (t == t_end) ? 1 : 2;
    } else {
      const double w_ref = std::ldexp(H->wT, t + 1);             // lower boundary of level t+2
      if (!add(nullptr, w_ref, std::ldexp(r_hi, -(t + 1)), std::ldexp(r_hi, -t), w_ref, t + 2, true)) return -1;
    }
  }
  H->ngroups = ng;
  if (total > (1LL << 31)) return -1;
  return total;
}
