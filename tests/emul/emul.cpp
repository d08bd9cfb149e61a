// emul.cpp -- TEST HARNESS ONLY.  Compiles the product's per-element arithmetic
// (spectralkernels.jl_b200/csrc/sk_math.h, sk_host_util.h, sk_plan_host.cpp) with g++ and runs it in
// plain loops, so that the index arithmetic of the CUDA kernels can be checked against the oracle on
// a machine without a GPU.  It is not a product path: nothing in spectralkernels.jl_b200/ loads it.
#include "sk_host_util.h"
#include "sk_rules.cuh"
#include "sk_k8.h"

#include <algorithm>
#include <cstring>
#include <vector>

extern "C" {

int emul_es_plan(int w, SkEsPlan *out) { return sk_plan_make_es(w, out); }

int emul_geom(const SkEsPlan *P, double w_lo, double w_hi, double r_lo, double r_hi, SkGeom *G) {
  return sk_make_geom(*P, w_lo, w_hi, r_lo, r_hi, G);
}

// spread + deconvolve + zero-pad: fills fft_in[nf2] (interleaved complex) for one rule
void emul_spread(const SkEsPlan *P, const SkGeom *G, long long M, const double *no, const double *c /*interleaved*/,
                 double *fft_in) {
  std::vector<double> ph(M), pl(M);
  std::vector<sk_cplx> cs(M);
  for (long long k = 0; k < M; ++k) sk_source_prep(*G, no[k], c[2 * k], c[2 * k + 1], &ph[k], &pl[k], &cs[k].x, &cs[k].y);
#pragma omp parallel for schedule(static)
  for (long long j = 0; j < G->nf2; ++j)
    sk_spread_mode(*P, *G, j, ph.data(), pl.data(), cs.data(), M, &fft_in[2 * j], &fft_in[2 * j + 1]);
}

// interpolate one grid (interleaved complex, nf2 entries) at N targets
void emul_interp(const SkEsPlan *P, const SkGeom *G, long long N, const double *r, const double *grid, double *out) {
  std::vector<sk_cplx> g(G->nf2);
  std::memcpy(g.data(), grid, sizeof(sk_cplx) * G->nf2);
#pragma omp parallel for schedule(static)
  for (long long j = 0; j < N; ++j) sk_interp_point<16, 1>(*P, *G, r[j], g.data(), &out[2 * j], &out[2 * j + 1]);
}

// the block algorithm of k_interp_cells (csrc/sk_kernels.cuh) in plain loops: two interleaved grids
// (layout [nf2][2]), sorted targets, blocks of `tpb` targets, cell polynomials when a block spans at
// most cmax cells, per-target taps otherwise.  out: [N][2] complex (rule 0, rule 1).  Returns the
// number of blocks that took the cell-polynomial path.
long long emul_interp_cells(const SkEsPlan *P, const SkGeom *G, long long N, const double *r, const double *grid2,
                            int tpb, int cmax, double *out) {
  const int W = 16;
  std::vector<sk_cplx> g((size_t)G->nf2 * 2);
  std::memcpy(g.data(), grid2, sizeof(sk_cplx) * G->nf2 * 2);
  std::vector<double> E(8 * 8), O(8 * 8);
  for (int i = 0; i < 8; ++i) for (int q = 0; q < 8; ++q) { E[i * 8 + q] = P->E[i][q]; O[i * 8 + q] = P->O[i][q]; }
  long long npoly = 0;
  sk_cplx tab[65];
  for (int k = 0; k < 65; ++k) sk_sincos2pi_table_fill(tab, k);
  for (long long j0 = 0; j0 < N; j0 += tpb) {
    const int cnt = (int)std::min<long long>(tpb, N - j0);
    const long long lf = sk_target_coord<W>(*G, r[j0]).l0, ll = sk_target_coord<W>(*G, r[j0 + cnt - 1]).l0;
    const long long ncell = ll - lf + 1;
    if (ncell >= 1 && ncell <= cmax) {
      ++npoly;
      const double *win = reinterpret_cast<const double *>(g.data() + (size_t)lf * 2);   // [ncell+W-1][4]
      std::vector<double> coef((size_t)ncell * SK_NC * 4);
      for (long long t = 0; t < ncell * SK_NC * 4; ++t) {
        const int comp = t & 3, q = (t >> 2) & (SK_NC - 1);
        const long long cell = t / (SK_NC * 4);
        coef[t] = sk_cell_coef<W>(E.data(), O.data(), win + cell * 4 + comp, 4, q);
      }
      for (long long t = 0; t < ncell * 4; ++t) {
        const int comp = t & 3;
        const long long cell = t >> 2;
        double a[4];
        sk_cell_deconv_cubic(*P, *G, (double)(lf + cell - G->nf2 / 2) + (0.5 * W - 0.5), a);
        sk_cell_fold(coef.data() + (size_t)cell * SK_NC * 4 + comp, 4, a);
      }
      for (int t = 0; t < cnt; ++t) {
        const double rr = r[j0 + t];
        const SkTargetCoord tc = sk_target_coord<W>(*G, rr);
        long long cell = tc.l0 - lf;
        double a[4];
        sk_cell_horner<4>(coef.data() + (size_t)cell * SK_NC * 4, tc.s, a);
        double sn, cs;
        sk_sincos2pi(tab, sk_frac_prod(G->wc, rr, 0.0), &sn, &cs);
        double *o = out + (j0 + t) * 4;
        o[0] = a[0] * cs - a[1] * sn; o[1] = a[0] * sn + a[1] * cs;
        o[2] = a[2] * cs - a[3] * sn; o[3] = a[2] * sn + a[3] * cs;
      }
    } else {
      // sparse block: the same cell arithmetic, one cell at a time (the warp path of the kernel)
      for (int t = 0; t < cnt; ++t) {
        const double rr = r[j0 + t];
        const SkTargetCoord tc = sk_target_coord<W>(*G, rr);
        const double *win = reinterpret_cast<const double *>(g.data() + (size_t)tc.l0 * 2);
        double coef[SK_NC * 4];
        for (int i = 0; i < SK_NC * 4; ++i) coef[i] = sk_cell_coef<W>(E.data(), O.data(), win + (i & 3), 4, i >> 2);
        double a4[4];
        sk_cell_deconv_cubic(*P, *G, (double)(tc.l0 - G->nf2 / 2) + (0.5 * W - 0.5), a4);
        for (int comp = 0; comp < 4; ++comp) sk_cell_fold(coef + comp, 4, a4);
        double a[4];
        sk_cell_horner<4>(coef, tc.s, a);
        double sn, cs;
        sk_sincos2pi(tab, sk_frac_prod(G->wc, rr, 0.0), &sn, &cs);
        double *o = out + (j0 + t) * 4;
        o[0] = a[0] * cs - a[1] * sn; o[1] = a[0] * sn + a[1] * cs;
        o[2] = a[2] * cs - a[3] * sn; o[3] = a[2] * sn + a[3] * cs;
      }
    }
  }
  return npoly;
}

// per-target taps on two interleaved grids (the k_interp_session body)
void emul_interp2(const SkEsPlan *P, const SkGeom *G, long long N, const double *r, const double *grid2, double *out) {
  std::vector<sk_cplx> g((size_t)G->nf2 * 2);
  std::memcpy(g.data(), grid2, sizeof(sk_cplx) * G->nf2 * 2);
#pragma omp parallel for schedule(static)
  for (long long j = 0; j < N; ++j) {
    double fre[2], fim[2];
    sk_interp_point<16, 2>(*P, *G, r[j], g.data(), fre, fim);
    double *o = out + j * 4;
    o[0] = fre[0]; o[1] = fim[0]; o[2] = fre[1]; o[3] = fim[1];
  }
}

// updatequadbufs! on the "device" generator
void emul_gen_sources(int m, int k, double a, double b, double p, int origin_jacobi, int weight_in_f, int logw,
                      int family, int deriv, const double *params, int nparam, const double *leg_no1,
                      const double *leg_wt1, const double *leg_no2, const double *leg_wt2, const double *jac_no1,
                      const double *jac_wt1, const double *jac_no2, const double *jac_wt2, double *no1, double *buf1,
                      double *no2, double *buf2) {
  SkPanelSpec S;
  std::memset(&S, 0, sizeof(S));
  S.m = m; S.k = k; S.origin_jacobi = origin_jacobi; S.weight_in_f = weight_in_f; S.logw = logw;
  S.family = family; S.deriv = deriv; S.nparam = nparam; S.p = p;
  for (int i = 0; i < nparam; ++i) S.params[i] = params[i];
  sk_fill_subpanels(a, b, k, S.bmad2, S.bpad2);
  S.jac_scale = std::pow(S.bmad2[0], p + 1);
  for (long long i = 0; i < (long long)m * k; ++i) sk_gen_source(S, 0, i, leg_no1, leg_wt1, jac_no1, jac_wt1, &no1[i], &buf1[i]);
  for (long long i = 0; i < 2LL * m * k; ++i) sk_gen_source(S, 1, i, leg_no2, leg_wt2, jac_no2, jac_wt2, &no2[i], &buf2[i]);
}

int emul_gauss_rule(int n, double p, double *no, double *wt) { return sk_plan_gauss_rule(n, p, no, wt); }

// the device rule generator (sk_rules.cuh: double-double Newton) in a plain loop
int emul_gauss_rule_dd(int n, double p, double *no, double *wt) {
  std::vector<double> A(2 * (size_t)n), B(2 * (size_t)n), C(2 * (size_t)n);
  if (sk_plan_jacobi_coeffs(n, p, A.data(), B.data(), C.data()) != 0) return -1;
  SkRuleJob J;
  J.n = n; J.p = p;
  J.A = (const sk_dd *)A.data(); J.B = (const sk_dd *)B.data(); J.C = (const sk_dd *)C.data();
  J.no = no; J.wt = wt;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < n; ++i) sk_gauss_node(J, i);
  return 0;
}

// max error of the lean sincos against libm over a sweep of |f| <= 1/2
double emul_sincos_err(int n) {
  sk_cplx tab[65];
  for (int k = 0; k < 65; ++k) sk_sincos2pi_table_fill(tab, k);
  double worst = 0.0;
  for (int i = 0; i <= n; ++i) {
    const double f = -0.5 + (double)i / n;
    double s, c;
    sk_sincos2pi(tab, f, &s, &c);
    const long double th = 6.283185307179586476925286766559L * (long double)f;
    worst = std::fmax(worst, std::fabs((double)(sinl(th) - s)));
    worst = std::fmax(worst, std::fabs((double)(cosl(th) - c)));
  }
  return worst;
}

double emul_trunc_err(double ta, double tn, double xpow, double x, int panel) { return sk_trunc_err(ta, tn, xpow, x, panel); }
int emul_converged(double te, double pk, double tau, int crit) { return sk_converged(te, pk, tau, crit) ? 1 : 0; }

// ---- nonuniform Hankel transform (sk_hankel.h), the arithmetic of sk_hankel.cuh in plain loops ----------
int emul_bessel_table(double *tab) { return sk_plan_bessel_table(SK_HK_NUMAX, SK_HK_TAB_INT, SK_HK_TAB_NC, tab); }
double emul_bessel_eval(const double *tab, int nu, double z) { return sk_bessel_tab(tab, nu, z); }
int emul_hk_sizes(int *plan_bytes, int *group_bytes, int *K, int *nlev, int *nch, int *ngrp) {
  *plan_bytes = (int)sizeof(SkHankelPlan); *group_bytes = (int)sizeof(SkHankelGroup);
  *K = SK_HK_K; *nlev = SK_HK_NLEV; *nch = SK_HK_NCH; *ngrp = SK_HK_NGRP;
  return 0;
}
long long emul_hk_plan(const SkEsPlan *P, int nu, double a, double b, double r_lo, double r_hi, SkHankelPlan *H,
                       SkHankelGroup *groups) {
  return sk_hk_make_plan(*P, nu, a, b, r_lo, r_hi, H, groups);
}
void emul_hk_plan_info(const SkHankelPlan *H, int *out /*q_lo q_hi t_full t_last ngroups*/) {
  out[0] = H->q_lo; out[1] = H->q_hi; out[2] = H->t_full; out[3] = H->t_last; out[4] = H->ngroups;
}
void emul_hk_group_info(const SkHankelGroup *g, int gi, long long *nf2, long long *off, double *D) {
  *nf2 = g[gi].G.nf2; *off = g[gi].grid_off; *D = g[gi].G.D;
}
// fields of group gi: out = {q_cut, q_from, q_to, shared, nf, nf2}; geo = {w_ref, wc, D, inv_hu}
void emul_hk_group_fields(const SkHankelGroup *g, int gi, long long *out, double *geo) {
  out[0] = g[gi].q_cut; out[1] = g[gi].q_from; out[2] = g[gi].q_to; out[3] = g[gi].shared;
  out[4] = g[gi].G.nf; out[5] = g[gi].G.nf2;
  geo[0] = g[gi].w_ref; geo[1] = g[gi].G.wc; geo[2] = g[gi].G.D; geo[3] = g[gi].G.inv_hu;
}
int emul_hk_level(double wT, double w) { return sk_hk_level(wT, w); }
int emul_hk_octave(double r_hi, double r) { return sk_hk_octave(r_hi, r); }
// Chebyshev coefficients of every level's partial sum for one rule: cheb[NLEV][NCH][2 rules], column `rule`
void emul_hk_fit(const SkHankelPlan *H, const double *tab, int rule, long long M, const double *no, const double *buf, double *cheb) {
  std::vector<long long> start(SK_HK_NLEV + 1, M);
  for (long long k = M - 1; k >= 0; --k) {                 // lev_start[q] = first source of level >= q
    const int l = sk_hk_level(H->wT, no[k]);
    for (int q = 0; q <= l; ++q) start[q] = k;
  }
  for (int i = 0; i < SK_HK_NLEV * SK_HK_NCH; ++i) cheb[2 * i + rule] = 0.0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int q = H->q_lo; q <= H->q_hi; ++q) {
    const double R = sk_hk_level_radius(H->r_hi, q);
    double vals[SK_HK_NCH];
    for (int i = 0; i < SK_HK_NCH; ++i) {
      const double rho = 0.5 * R * (sk_hk_cheb_node(i) + 1.0);
      double acc = 0.0;
      for (long long k = start[q]; k < start[q + 1]; ++k) acc += sk_hk_fit_term(tab, H->nu, no[k], buf[k], rho);
      vals[i] = acc;
    }
    for (int m = 0; m < SK_HK_NCH; ++m) cheb[(q * SK_HK_NCH + m) * 2 + rule] = sk_hk_cheb_coef(vals, m);
  }
}
// per-octave piecewise expansions from the per-level Chebyshev coefficients: loc[NLEV][NSUB][NLOC][2]
void emul_hk_local_poly(const SkHankelPlan *H, const double *cheb, double *loc) {
  std::memset(loc, 0, sizeof(double) * SK_HK_NLEV * SK_HK_NSUB * SK_HK_NLOC * 2);
  for (int tt = 0; tt <= H->q_hi; ++tt)
    for (int s = 0; s < SK_HK_NSUB; ++s) {
      double vals[SK_HK_NLOC * 2];
      for (int i = 0; i < SK_HK_NLOC; ++i)
        sk_hk_local(*H, cheb, sk_hk_local_node(*H, tt, s, i), tt < H->q_hi ? tt : 4096, &vals[2 * i]);
      for (int m = 0; m < SK_HK_NLOC; ++m)
        for (int rule = 0; rule < 2; ++rule)
          loc[(((size_t)tt * SK_HK_NSUB + s) * SK_HK_NLOC + m) * 2 + rule] = sk_hk_local_coef(vals + rule, m);
    }
}
// FFT input of group gi for one rule: fft_in[nf2][K][2] (interleaved complex), entries of the other rule untouched
void emul_hk_spread(const SkEsPlan *P, const SkHankelPlan *H, const SkHankelGroup *groups, int gi, int rule, long long M,
                    const double *no, const double *buf, double *fft_in) {
  const SkHankelGroup &g = groups[gi];
  std::vector<double> ph(M), pl(M), lam(M);
  std::vector<sk_cplx> cs(M);
  for (long long k = 0; k < M; ++k) sk_hk_source_prep(g, H->wT, g.q_cut, SK_HK_NLEV, no[k], buf[k], &ph[k], &pl[k], &cs[k], &lam[k]);
#pragma omp parallel for schedule(static)
  for (long long j = 0; j < g.G.nf2; ++j) {
    sk_cplx o[SK_HK_K];
    sk_hk_spread_mode(*P, *H, g.G, j, ph.data(), pl.data(), cs.data(), lam.data(), M, o);
    for (int n = 0; n < SK_HK_K; ++n) {
      fft_in[((j * SK_HK_K + n) * 2 + rule) * 2] = o[n].x;
      fft_in[((j * SK_HK_K + n) * 2 + rule) * 2 + 1] = o[n].y;
    }
  }
}
// all targets: out[N][2]
void emul_hk_eval(const SkEsPlan *P, const SkHankelPlan *H, const SkHankelGroup *groups, const double *grid,
                  const double *loc, long long N, const double *r, double *out) {
#pragma omp parallel for schedule(static)
  for (long long j = 0; j < N; ++j)
    sk_hk_point<16>(*P, *H, groups, reinterpret_cast<const sk_cplx *>(grid), loc, r[j], &out[2 * j]);
}

// the cell-polynomial variant (k_hankel_cells): every target builds the polynomial of its cell with the same
// per-item formulas the warp kernel distributes over its lanes
void emul_hk_eval_cells(const SkEsPlan *P, const SkHankelPlan *H, const SkHankelGroup *groups, const double *grid,
                        const double *loc, long long N, const double *r, double *out, long long *ncell_path) {
  std::vector<double> E(8 * 8), O(8 * 8);
  for (int i = 0; i < 8; ++i) for (int q = 0; q < 8; ++q) { E[i * 8 + q] = P->E[i][q]; O[i * 8 + q] = P->O[i][q]; }
  long long cnt = 0;
  sk_cplx tab[65];
  for (int k = 0; k < 65; ++k) sk_sincos2pi_table_fill(tab, k);
#pragma omp parallel for schedule(static) reduction(+ : cnt)
  for (long long j = 0; j < N; ++j) {
    const int t = sk_hk_octave(H->r_hi, r[j]);
    double lo[2], asy[2] = {0.0, 0.0};
    sk_hk_local2(*H, loc, r[j], t, lo);
    const int gi = sk_hk_group_of_octave(*H, t);
    if (gi >= 0 && gi < H->ngroups) {
      const SkHankelGroup &g = groups[gi];
      const sk_cplx *gg = reinterpret_cast<const sk_cplx *>(grid) + g.grid_off;
      const SkTargetCoord tc = sk_target_coord<16>(g.G, r[j]);
      if (sk_hk_cell_ok<16>(g, tc.l0)) {
        double coef[SK_NC * 4];
        sk_hk_cell_build<16>(*P, *H, g, gg, tc.l0, E.data(), O.data(), coef);
        sk_hk_cell_eval(coef, tab, g.G, r[j], tc.s, asy);
        ++cnt;
      } else {
        sk_hk_interp_point<16>(*P, *H, g, gg, r[j], asy);
      }
    }
    out[2 * j] = asy[0] + lo[0];
    out[2 * j + 1] = asy[1] + lo[1];
  }
  *ncell_path = cnt;
}

// K8 (csrc/sk_k8.cuh) in plain loops with the index arithmetic of csrc/sk_k8.h: stats, sampled coarse histogram,
// plan, scatter into fine bins, per-bin counting sort + group ranks + heads, sequential prefix instead of the
// look-back.  Returns 0 (ok), 1 (a fine bin overflowed: the product then takes the general sort), 2 (invalid input),
// 3 / 4 (a bug: fine bin out of range / fine-bin map not monotone).
// diag[0] = number of fine bins, diag[1] = largest fill, diag[2] = largest sub-bin group, diag[3] = presorted.
int emul_k8(const double *xs, long long n, unsigned int samp, double *uxs, unsigned int *inv, long long *n_unique,
            long long *diag) {
  auto keyof = [](double x, bool *bad) -> unsigned long long {
    if (!(x >= 0.0) || std::isinf(x)) { *bad = true; x = 0.0; }
    if (x == 0.0) x = 0.0;
    unsigned long long k;
    std::memcpy(&k, &x, 8);
    return k;
  };
  bool bad = false;
  unsigned long long kmin = ~0ull, kmax = 0ull, nzero = 0ull, ndesc = 0ull;
  for (long long j = 0; j < n; ++j) {
    const unsigned long long k = keyof(xs[j], &bad);
    if (j > 0 && !(xs[j] > xs[j - 1])) ++ndesc;
    if (k == 0ull) ++nzero; else { kmin = std::min(kmin, k); kmax = std::max(kmax, k); }
  }
  if (bad) return 2;
  diag[0] = diag[1] = diag[2] = 0;
  const bool unsorted = ndesc != 0ull;
  if (samp == 0u) samp = sk_k8_samp((unsigned long long)n, ndesc);     // 0: the product's choice
  diag[3] = unsorted ? 0 : 1;
  if (!unsorted) {
    for (long long j = 0; j < n; ++j) { uxs[j] = xs[j] == 0.0 ? 0.0 : xs[j]; inv[j] = (unsigned int)j; }
    *n_unique = n;
    return 0;
  }
  const unsigned long long npos = (unsigned long long)n - nzero;
  std::vector<unsigned int> chist(SK_K8_NC, 0u), ccum(SK_K8_NC), ccnt(SK_K8_NC);
  unsigned long long mul = 0;
  unsigned int nfine = 0;
  if (npos) {
    mul = sk_k8_mul(kmin, kmax);
    unsigned long long sampled = 0;
    for (long long s = 0; s < (n + 31) / 32; ++s) {
      if (!sk_k8_sampled((unsigned long long)s, samp)) continue;
      for (long long j = s * 32; j < std::min<long long>(n, s * 32 + 32); ++j) {
        bool b2 = false;
        const unsigned long long k = keyof(xs[j], &b2);
        if (k) {
          unsigned int cb;
          unsigned long long frac;
          sk_k8_coarse(k, kmin, mul, &cb, &frac);
          if (cb >= (unsigned int)SK_K8_NC) return 3;
          ++chist[cb];
          ++sampled;
        }
      }
    }
    const bool single = npos <= (unsigned long long)SK_K8_CAP;
    auto scaled = [&](unsigned long long cum) { return sampled ? cum * npos / sampled : 0ull; };
    unsigned long long run = 0;
    for (int c = 0; c < SK_K8_NC; ++c) {
      const unsigned long long lo = scaled(run), hi = scaled(run + chist[c]);
      run += chist[c];
      ccum[c] = single ? 0u : (unsigned int)lo;
      ccnt[c] = single ? 0u : (unsigned int)(hi - lo);
    }
    nfine = single ? 1u : (unsigned int)((scaled(sampled) >> SK_K8_TARGET_LOG) + 1ull);
  }
  diag[0] = nfine;
  if ((unsigned long long)nfine > (unsigned long long)(n >> SK_K8_TARGET_LOG) + 2ull) return 3;   // the host's bound
  std::vector<std::vector<std::pair<unsigned long long, unsigned int>>> bins(nfine);
  for (long long j = 0; j < n; ++j) {
    bool b2 = false;
    const unsigned long long k = keyof(xs[j], &b2);
    if (k == 0ull) { inv[j] = 0u; continue; }
    const unsigned int f = sk_k8_fine_bin(k, kmin, mul, ccum.data(), ccnt.data());
    if (f >= nfine) return 3;                                        // would be out of bounds on the device
    bins[f].push_back({k, (unsigned int)j});
  }
  int rc = 0;
  unsigned long long uoff = nzero ? 1ull : 0ull;
  if (nzero) uxs[0] = 0.0;
  unsigned long long last_hi = 0ull;
  for (unsigned int f = 0; f < nfine; ++f) {
    auto &B = bins[f];
    const int cnt = (int)B.size();
    diag[1] = std::max<long long>(diag[1], cnt);
    if (cnt > SK_K8_CAP) { rc = 1; continue; }
    if (cnt == 0) continue;
    unsigned long long lo = ~0ull, hi = 0ull;
    for (auto &e : B) { lo = std::min(lo, e.first); hi = std::max(hi, e.first); }
    if (lo < last_hi) return 4;                                      // the fine-bin map is not monotone
    last_hi = hi;
    const double scale = sk_k8_ssb_scale(lo, hi);
    std::vector<int> off(SK_K8_NSSB + 1, 0), ssb(cnt), rk(cnt);
    for (int t = 0; t < cnt; ++t) { ssb[t] = sk_k8_ssb(B[t].first, lo, scale); rk[t] = off[ssb[t]]++; }
    int run = 0;
    for (int i = 0; i <= SK_K8_NSSB; ++i) { const int c = i < SK_K8_NSSB ? off[i] : 0; off[i] = run; run += c; }
    std::vector<unsigned long long> placed(cnt), fin_key(cnt);
    std::vector<unsigned char> head(cnt);
    std::vector<int> fin(cnt);
    for (int t = 0; t < cnt; ++t) placed[off[ssb[t]] + rk[t]] = B[t].first;
    for (int t = 0; t < cnt; ++t) {
      const int o = off[ssb[t]], ge = off[ssb[t] + 1], me = o + rk[t];
      diag[2] = std::max<long long>(diag[2], ge - o);
      int less = 0, eqb = 0;
      for (int p = o; p < ge; ++p) { less += placed[p] < B[t].first; eqb += (placed[p] == B[t].first) && (p < me); }
      fin[t] = o + less + eqb;
      fin_key[fin[t]] = B[t].first;
      head[fin[t]] = eqb == 0;
    }
    std::vector<int> luid(cnt);
    int h = 0;
    for (int p = 0; p < cnt; ++p) {
      h += head[p];
      luid[p] = h;
      if (head[p]) std::memcpy(&uxs[uoff + h - 1], &fin_key[p], 8);
    }
    for (int t = 0; t < cnt; ++t) inv[B[t].second] = (unsigned int)(uoff + luid[fin[t]] - 1);
    uoff += h;
  }
  *n_unique = (long long)uoff;
  return rc;
}

}  // extern "C"
