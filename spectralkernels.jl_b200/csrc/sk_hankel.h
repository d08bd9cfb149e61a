// sk_hankel.h -- per-element arithmetic of the O(N) nonuniform Hankel transform
//
//     g_j = sum_k c_k J_nu(2 pi w_k r_j)                      (src/quadrature.jl:137-161, `nufht`)
//
// which the reference takes from FastHankelTransform.jl for dim >= 2.  Written from scratch on top of
// the type-3 NUFFT of sk_math.h; host/device inline functions, so the CUDA kernels of sk_hankel.cuh are
// thin wrappers and tests/emul/ can run the same arithmetic with g++.
//
// Scheme.  With ZL = 32 the Hankel expansion
//     J_nu(z) = sqrt(2/(pi z)) Re[ e^{i(z - nu pi/2 - pi/4)} sum_{n<K} a_n(nu) (i/z)^n ]   (DLMF 10.17.3)
// with K = 12 terms is exact to ~1e-15 for z >= ZL.  The (frequency, distance) plane is cut dyadically:
//     frequency levels   L_0 = [0, wT),  L_q = [wT 2^(q-1), wT 2^q)          wT = ZL / (2 pi r_hi)
//     distance octaves   O_t = (r_hi 2^-(t+1), r_hi 2^-t]                     t = 0, 1, ...
// For a target in octave t every source of level >= t+2 has z >= ZL: that suffix of the (ascending) source
// list is summed with K type-3 NUFFTs (one per term: strengths c_k a_n (w_ref/w_k)^(n+1/2), the target
// applies (2 pi w_ref r)^(-n) by Horner in i/z_ref), all K terms and both quadrature rules in one batched
// grid.  Levels q <= t+1 have z < 2 ZL on the whole of [0, R_q], R_q = r_hi 2^(1-q): there the level's
// partial sum is a band-limited function of r and is replaced by its Chebyshev interpolant of SK_HK_NCH
// terms on [0, R_q], built from direct sums at the Chebyshev nodes (J_nu from a piecewise-polynomial table on
// [0, 64] fitted in quad precision).  Nothing is evaluated per (source, target) pair: the work per target is
// K * 4 * w FMAs (interpolation) plus one short Clenshaw recurrence: the local levels of an octave are summed
// once per sub-interval into a piecewise expansion of the octave (SK_HK_NSUB pieces x SK_HK_NLOC terms).
// Octaves t with t+2 <= level(a) see the whole sub-interval [a, b] asymptotically and share one transform; the
// small octaves (t >= SK_HK_T_SHARE) share one geometry and are spread incrementally.  In the default
// interpolation kernel the K terms are folded into one polynomial per fine-grid cell (see "cell polynomials
// across the K terms" below), so a target costs four Horner chains.
#pragma once
#include "sk_math.h"

#define SK_HK_ZL 32.0
#define SK_HK_K 12           // terms of the Hankel expansion
#define SK_HK_NCH 72         // Chebyshev terms per level (z < 2 ZL = 64 needs ~64/2 + 26)
#define SK_HK_NLEV 48        // dyadic frequency levels
#define SK_HK_NSUB 16        // pieces per octave of the per-octave local expansion
#define SK_HK_NLOC 16        // Chebyshev terms per piece (z changes by 2 across a piece: truncation < 1e-18)
#define SK_HK_NGRP 48        // transforms per sub-interval
#define SK_HK_NUMAX 3        // tabulated Bessel orders 0..3 (dim <= 6, or dim <= 4 with derivatives)
#define SK_HK_TAB_INT 32     // table intervals [2i, 2i+2]
#define SK_HK_TAB_NC 16      // monomial coefficients per interval, in t = z - (2i+1)
#define SK_HK_TAB_SIZE ((SK_HK_NUMAX + 1) * SK_HK_TAB_INT * SK_HK_TAB_NC)

struct SkHankelGroup {       // one batched transform: a suffix of the sources against a band of targets
  SkGeom G;
  double w_ref;              // strengths carry (w_ref / w)^(n + 1/2), targets (2 pi w_ref r)^(-n - 1/2)
  int q_cut;                 // sources of level >= q_cut belong to the group
  // Incremental spreading.  The groups of the small octaves (t >= SK_HK_T_SHARE) share one geometry and one w_ref;
  // they are processed from the deepest octave up and each adds only the levels [q_from, q_to) to a running mode
  // buffer (modes of octave t = modes of octave t+1 + level t+2), which is then transformed into the group's own
  // grids: one full pass over the sources for the whole set instead of one per octave.  A standalone group has
  // [q_from, q_to) = [q_cut, SK_HK_NLEV) and shared = 0.
  int q_from, q_to;
  int shared;                // 0: standalone; 1: first (deepest) group of the shared set; 2: later group of the set
  long long grid_off;        // offset (in sk_cplx) of the group's grids; layout [nf2][K][2 rules]
};
#define SK_HK_T_SHARE 3

struct SkHankelPlan {
  int nu;
  int q_lo, q_hi;            // levels that can hold sources of [a, b]
  int t_full;                // octaves t <= t_full share group 0 (all of [a, b] is asymptotic); -1: none
  int t_last;                // octaves t > t_last have no asymptotic part
  int ngroups;
  double wT;                 // ZL / (2 pi r_hi)
  double r_hi;               // largest active distance (global over the ranks of a sharded run)
  double cphi, sphi;         // cos, sin of nu pi/2 + pi/4
  double ratio[SK_HK_K];     // a_{n+1}(nu) / a_n(nu) = (4 nu^2 - (2n+1)^2) / (8 (n+1))
  double zthr[SK_HK_K];      // term n contributes less than 1e-17 once z >= zthr[n] (non-increasing in n >= 1)
};

// ---- J_nu(z), 0 <= z <= 64, from the table ----------------------------------------------------------
SK_HD double sk_bessel_tab(const double *tab, int nu, double z) {
  int i = (int)(0.5 * z);
  i = i < 0 ? 0 : (i > SK_HK_TAB_INT - 1 ? SK_HK_TAB_INT - 1 : i);
  const double t = z - (double)(2 * i + 1);
  const double *c = tab + ((size_t)nu * SK_HK_TAB_INT + i) * SK_HK_TAB_NC;
  double v = c[SK_HK_TAB_NC - 1];
#pragma unroll
  for (int q = SK_HK_TAB_NC - 2; q >= 0; --q) v = sk_fma(v, t, c[q]);
  return v;
}

// ---- dyadic cuts (exact comparisons against power-of-two multiples: the same answer on host and device) --
// level of a frequency: the number of boundaries wT 2^(q-1), q >= 1, that are <= w
SK_HD int sk_hk_level(double wT, double w) {
  if (!(w >= wT)) return 0;
  const int e = ilogb(w) - ilogb(wT);                 // w / wT in (2^(e-1), 2^(e+1))
  const int fl = (w >= ldexp(wT, e)) ? e : e - 1;     // floor(log2(w / wT))
  const int q = fl + 1;
  return q > SK_HK_NLEV - 1 ? SK_HK_NLEV - 1 : q;
}
// octave of a distance: t >= 0 with r in (r_hi 2^-(t+1), r_hi 2^-t]; r >= r_hi gives 0
SK_HD int sk_hk_octave(double r_hi, double r) {
  if (!(r < r_hi)) return 0;
  if (!(r > 0.0)) return 4096;
  const int e = ilogb(r_hi) - ilogb(r);               // r_hi / r in (2^(e-1), 2^(e+1))
  return (r <= ldexp(r_hi, -e)) ? e : e - 1;          // floor(log2(r_hi / r))
}
// number of leading terms of the expansion a target needs: every source of its group has z >= z_ref
SK_HD int sk_hk_nterms(const SkHankelPlan &H, double zref) {
  int k = 1;
#pragma unroll
  for (int n = 1; n < SK_HK_K; ++n) k = (zref < H.zthr[n]) ? n + 1 : k;
  return k;
}
SK_HD int sk_hk_group_of_octave(const SkHankelPlan &H, int t) {
  if (t > H.t_last) return -1;
  if (t <= H.t_full) return 0;
  return (t - H.t_full - 1) + (H.t_full >= 0 ? 1 : 0);
}
// right end of the interval on which level q is expanded locally
SK_HD double sk_hk_level_radius(double r_hi, int q) { return q >= 1 ? ldexp(r_hi, 1 - q) : r_hi; }

// ---- local part: Chebyshev interpolants of the levels' partial sums ------------------------------------
SK_HD double sk_hk_cheb_node(int i) {   // first-kind nodes, descending
  double s, c;
  sk_sincospi(((double)i + 0.5) / (double)SK_HK_NCH, &s, &c);
  return c;
}
// one term of the direct sum at a node: c_k J_nu(2 pi w_k rho)
SK_HD double sk_hk_fit_term(const double *tab, int nu, double no, double buf, double rho) {
  return sk_mul(buf, sk_bessel_tab(tab, nu, sk_mul(sk_mul(6.283185307179586, no), rho)));
}
// coefficient m of the interpolant from the NCH node values
SK_HD double sk_hk_cheb_coef(const double *vals, int m) {
  double acc = 0.0;
  for (int i = 0; i < SK_HK_NCH; ++i) {
    const int k = (m * (2 * i + 1)) % (4 * SK_HK_NCH);          // cos(pi k / (2 NCH)), exact reduction
    double s, c;
    sk_sincospi((double)k / (double)(2 * SK_HK_NCH), &s, &c);
    acc = sk_fma(vals[i], c, acc);
  }
  return acc * (m == 0 ? 1.0 : 2.0) / (double)SK_HK_NCH;
}
// value of level q's interpolant at distance r (both rules); cheb layout [SK_HK_NLEV][SK_HK_NCH][2 rules]
SK_HD void sk_hk_local_level(const SkHankelPlan &H, const double *cheb, double r, int q, double *out) {
  const double R = sk_hk_level_radius(H.r_hi, q);
  const double x2 = 2.0 * sk_fma(r, 2.0 / R, -1.0);             // 2x, x in [-1, 1]
  const double *c = cheb + (size_t)q * (SK_HK_NCH * 2);
  double b1 = 0.0, b2 = 0.0, d1 = 0.0, d2 = 0.0;
  for (int j = SK_HK_NCH - 1; j >= 1; --j) {
    const double b0 = sk_fma(x2, b1, c[2 * j] - b2);
    const double d0 = sk_fma(x2, d1, c[2 * j + 1] - d2);
    b2 = b1; b1 = b0;
    d2 = d1; d1 = d0;
  }
  out[0] = sk_fma(0.5 * x2, b1, c[0] - b2);
  out[1] = sk_fma(0.5 * x2, d1, c[1] - d2);
}
// sum over the local levels of a target in octave t, in level order
SK_HD void sk_hk_local(const SkHankelPlan &H, const double *cheb, double r, int t, double *out) {
  out[0] = out[1] = 0.0;
  const int qe = (t + 1 < H.q_hi) ? t + 1 : H.q_hi;
  for (int q = H.q_lo; q <= qe; ++q) {
    double v[2];
    sk_hk_local_level(H, cheb, r, q, v);
    out[0] += v[0];
    out[1] += v[1];
  }
}

// Per-octave local expansion.  All targets of octave t need the same sum over the levels q <= t+1, a function of r
// with z < 2 ZL on the octave: it is tabulated once per sub-interval as SK_HK_NSUB pieces of SK_HK_NLOC Chebyshev
// terms (table index tt = min(t, q_hi); tt = q_hi is the catch-all piece set on [0, r_hi 2^-q_hi], where every
// level is local and z < ZL).  Per target this replaces ~2-3 recurrences of 72 terms by one of 16.
SK_HD void sk_hk_local_piece(const SkHankelPlan &H, double r, int t, int *tt, int *sp, double *u) {
  const int ti = t < H.q_hi ? t : H.q_hi;
  double v;
  if (ti < H.q_hi) v = sk_fma(r, 32.0 / ldexp(H.r_hi, -ti), -16.0);     // r in (R/2, R]  ->  (0, 16]
  else v = r * (16.0 / ldexp(H.r_hi, -H.q_hi));                          // r in (0, Rc]   ->  (0, 16]
  int s = (int)v;
  s = s < 0 ? 0 : (s > SK_HK_NSUB - 1 ? SK_HK_NSUB - 1 : s);
  *tt = ti;
  *sp = s;
  *u = 2.0 * (v - (double)s) - 1.0;
}
// distance of Chebyshev node i of piece s of table octave tt (the inverse of the map above)
SK_HD double sk_hk_local_node(const SkHankelPlan &H, int tt, int s, int i) {
  double sn, cs;
  sk_sincospi(((double)i + 0.5) / (double)SK_HK_NLOC, &sn, &cs);
  const double v = (double)s + 0.5 * (cs + 1.0);
  if (tt < H.q_hi) return (v + 16.0) * (ldexp(H.r_hi, -tt) / 32.0);
  return v * (ldexp(H.r_hi, -H.q_hi) / 16.0);
}
SK_HD double sk_hk_local_coef(const double *vals /*[NLOC], stride 2*/, int m) {
  double acc = 0.0;
  for (int i = 0; i < SK_HK_NLOC; ++i) {
    const int k = (m * (2 * i + 1)) % (4 * SK_HK_NLOC);
    double s, c;
    sk_sincospi((double)k / (double)(2 * SK_HK_NLOC), &s, &c);
    acc = sk_fma(vals[2 * i], c, acc);
  }
  return acc * (m == 0 ? 1.0 : 2.0) / (double)SK_HK_NLOC;
}
// loc layout [table octave][SK_HK_NSUB][SK_HK_NLOC][2 rules]
SK_HD void sk_hk_local2(const SkHankelPlan &H, const double *loc, double r, int t, double *out) {
  int tt, s;
  double u;
  sk_hk_local_piece(H, r, t, &tt, &s, &u);
  const double *c = loc + ((size_t)tt * SK_HK_NSUB + s) * (SK_HK_NLOC * 2);
  const double x2 = 2.0 * u;
  double b1 = 0.0, b2 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
  for (int j = SK_HK_NLOC - 1; j >= 1; --j) {
    const double b0 = sk_fma(x2, b1, c[2 * j] - b2);
    const double d0 = sk_fma(x2, d1, c[2 * j + 1] - d2);
    b2 = b1; b1 = b0;
    d2 = d1; d1 = d0;
  }
  out[0] = sk_fma(u, b1, c[0] - b2);
  out[1] = sk_fma(u, d1, c[1] - d2);
}

// ---- asymptotic part ---------------------------------------------------------------------------------------
// position on the group's spread grid, term-0 strength c_k (w_ref/w_k)^(1/2) (pre-phased), and the ratio
// lam = w_ref / w_k that advances the strength from one term to the next
// (levels [q_from, q_to): the emulation spreads every group in one go and passes [g.q_cut, SK_HK_NLEV))
SK_HD void sk_hk_source_prep(const SkHankelGroup &g, double wT, int q_from, int q_to, double no, double buf, double *pos_hi,
                             double *pos_lo, sk_cplx *cs, double *lam) {
  const int lev = sk_hk_level(wT, no);
  const bool in = lev >= q_from && lev < q_to && no > 0.0;
  const double l = in ? g.w_ref / no : 0.0;
  sk_source_prep(g.G, no, in ? sk_mul(buf, sqrt(l)) : 0.0, 0.0, pos_hi, pos_lo, &cs->x, &cs->y);
  *lam = l;
}

// spread + mode deconvolution + zero-pad for all K terms of one rule: FFT-input element j (emulation
// twin of k_spread_hankel; exp-of-semicircle evaluated directly instead of through the tap polynomials)
SK_HD void sk_hk_spread_mode(const SkEsPlan &P, const SkHankelPlan &H, const SkGeom &G, long long j, const double *pos_hi,
                             const double *pos_lo, const sk_cplx *cs, const double *lam, long long M, sk_cplx *out /*[K]*/) {
  for (int n = 0; n < SK_HK_K; ++n) out[n].x = out[n].y = 0.0;
  const long long nm = (j < G.nf2 / 2) ? j : j - G.nf2;
  if (nm < -(G.nf / 2) || nm >= G.nf / 2) return;
  const double ctr = (double)nm, half = 0.5 * P.w;
  long long a = 0, b = M;
  while (a < b) {
    const long long mid = (a + b) >> 1;
    if (pos_hi[mid] < ctr - half - 1e-6) a = mid + 1; else b = mid;
  }
  for (long long k = a; k < M && pos_hi[k] <= ctr + half + 1e-6; ++k) {
    const double z = ((ctr - pos_hi[k]) - pos_lo[k]) / half;
    const double wgt = sk_es_direct(z, P.beta);
    double cx = wgt * cs[k].x, cy = wgt * cs[k].y;
    for (int n = 0; n < SK_HK_K; ++n) {
      out[n].x += cx;
      out[n].y += cy;
      const double f = lam[k] * H.ratio[n];
      cx *= f;
      cy *= f;
    }
  }
  double q = sk_deconv(P, G.t_cell * fabs(ctr));
  if (nm & 1) q = -q;
  for (int n = 0; n < SK_HK_K; ++n) { out[n].x *= q; out[n].y *= q; }
}

// interpolate the K x 2 grids of a group at distance r and sum the expansion: out[rule]
template <int W>
SK_HD void sk_hk_interp_point(const SkEsPlan &P, const SkHankelPlan &H, const SkHankelGroup &g, const sk_cplx *grid,
                              double r, double *out) {
  const SkTargetCoord t = sk_target_coord<W>(g.G, r);
  double taps[W];
  sk_es_taps<W>(P, t.s, taps);
  const double zref = sk_mul(sk_mul(6.283185307179586, g.w_ref), r);
  const double iz = 1.0 / zref;
  double c0r = 0.0, c0i = 0.0, c1r = 0.0, c1i = 0.0;
  const sk_cplx *gp = grid + (size_t)t.l0 * (SK_HK_K * 2);
#pragma unroll 1
  for (int n = sk_hk_nterms(H, zref) - 1; n >= 0; --n) {
    double a0r = 0.0, a0i = 0.0, a1r = 0.0, a1i = 0.0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const sk_cplx v0 = gp[(i * SK_HK_K + n) * 2], v1 = gp[(i * SK_HK_K + n) * 2 + 1];
      a0r = sk_fma(taps[i], v0.x, a0r);
      a0i = sk_fma(taps[i], v0.y, a0i);
      a1r = sk_fma(taps[i], v1.x, a1r);
      a1i = sk_fma(taps[i], v1.y, a1i);
    }
    // C <- raw_n + (i / z_ref) C
    const double n0r = sk_fma(-c0i, iz, a0r), n0i = sk_fma(c0r, iz, a0i);
    const double n1r = sk_fma(-c1i, iz, a1r), n1i = sk_fma(c1r, iz, a1i);
    c0r = n0r; c0i = n0i; c1r = n1r; c1i = n1i;
  }
  const double qf = sk_deconv(P, g.G.t_cell * t.yabs);
  double sn, cs;
  sk_post_phase(g.G, r, &sn, &cs);
  // e^{i (2 pi wc r - phi)}
  const double er = sk_fma(cs, H.cphi, sn * H.sphi), ei = sk_fma(sn, H.cphi, -cs * H.sphi);
  const double amp = qf * sqrt(0.6366197723675814 * iz);             // sqrt(2 / (pi z_ref))
  out[0] = amp * sk_fma(c0r, er, -c0i * ei);
  out[1] = amp * sk_fma(c1r, er, -c1i * ei);
}

// the whole transform at one target (both rules)
template <int W>
SK_HD void sk_hk_point(const SkEsPlan &P, const SkHankelPlan &H, const SkHankelGroup *groups, const sk_cplx *grid,
                       const double *loctab, double r, double *out) {
  const int t = sk_hk_octave(H.r_hi, r);
  double loc[2];
  sk_hk_local2(H, loctab, r, t, loc);
  const int gi = sk_hk_group_of_octave(H, t);
  double asy[2] = {0.0, 0.0};
  if (gi >= 0 && gi < H.ngroups) sk_hk_interp_point<W>(P, H, groups[gi], grid + groups[gi].grid_off, r, asy);
  out[0] = asy[0] + loc[0];
  out[1] = asy[1] + loc[1];
}

// ---- cell polynomials across the K terms --------------------------------------------------------------------
// All targets whose w-wide window starts at the same fine-grid index ("cell") see the same 16 x K x 2 grid values,
// and inside a cell r = r_mid - s h (h = half a cell) moves 1/z by a relative eps = h / r_mid only.  With
//   (1/z)^(n + 1/2) = z_mid^-(n + 1/2) (1 - eps s)^-(n + 1/2) = z_mid^-(n + 1/2) sum_j beta_{n,j} (eps s)^j
// truncated at j <= 3 (eps <= 2^-12: the dropped term is < 1e-17 of the sum), the whole expansion
//   sqrt(2/(pi z)) e^{-i phi} sum_n a_n (i/z)^n sum_i tap_i(s) g_{i,n}
// becomes ONE complex polynomial of degree 15 in s per rule and cell (layout [SK_NC][4], exactly the cell
// polynomial of K4, sk_math.h), and a target costs four Horner chains instead of K x 4 x w FMAs.  The 16 x K x 2
// grid values of the cell are first combined over n for each j (weights w_{n,j}), then turned into polynomial
// coefficients (sk_cell_coef), shifted-added over j and folded with the deconvolution cubic (sk_cell_fold).
// Every output item is computed by a fixed formula from the cell's data only, so results do not depend on which
// thread, warp or GPU evaluates a target.
#define SK_HK_NJ 4
#define SK_HK_CELL_MIN 2048.0     // cells between r = 0 and the cell centre: eps = 1 / (2 * that) <= 2^-12

struct SkHkCell {
  double ymid;      // y at s = 0 (cell centre), y = ymid - s/2
  double eps;       // relative change of r per unit s
  double iz_mid;    // 1 / (2 pi w_ref r_mid)
  int nt;           // leading terms kept (for the smallest z of the cell)
  int ok;           // eps small enough for the truncated binomial series
};

// only the admission test of sk_hk_cell_setup (what a target needs to know about its cell)
template <int W>
SK_HD bool sk_hk_cell_ok(const SkHankelGroup &g, long long l0) {
  const double ymid = (double)(l0 - g.G.nf2 / 2) + (0.5 * W - 0.5);
  return sk_fma(g.G.D, g.G.kap_hi, ymid) >= SK_HK_CELL_MIN;
}

template <int W>
SK_HD SkHkCell sk_hk_cell_setup(const SkHankelPlan &H, const SkHankelGroup &g, long long l0) {
  SkHkCell c;
  c.ymid = (double)(l0 - g.G.nf2 / 2) + (0.5 * W - 0.5);
  const double cfo = sk_fma(g.G.D, g.G.kap_hi, c.ymid);           // cells from the origin r = 0 to the cell centre
  c.ok = cfo >= SK_HK_CELL_MIN;
  c.eps = 0.5 / cfo;
  const double r_mid = cfo / g.G.kap_hi;
  const double z_mid = sk_mul(sk_mul(6.283185307179586, g.w_ref), r_mid);
  c.iz_mid = 1.0 / z_mid;
  c.nt = sk_hk_nterms(H, z_mid * (1.0 - c.eps));                  // s = +1 is the smallest r of the cell
  return c;
}

// weight of grid term n in the j-th power of (eps s): amp_mid e^{-i phi} i^n z_mid^-n beta_{n,j} eps^j
SK_HD void sk_hk_cell_weight(const SkHankelPlan &H, const SkHkCell &c, int n, int j, double *wr, double *wi) {
  double m = sqrt(0.6366197723675814 * c.iz_mid);                 // sqrt(2 / (pi z_mid))
  for (int k = 0; k < n; ++k) m *= c.iz_mid;
  const double e = (double)n + 0.5;
  double beta = 1.0;
  for (int k = 0; k < j; ++k) beta = beta * (e + (double)k) / (double)(k + 1) * c.eps;
  m *= beta;
  // (cphi - i sphi) i^n
  double rr = H.cphi, ri = -H.sphi;
  for (int k = 0; k < (n & 3); ++k) { const double t = rr; rr = -ri; ri = t; }
  *wr = m * rr;
  *wi = m * ri;
}

// combination over n of the grid values at window point i of one rule, for the SK_HK_NJ powers:
// out[j] (complex) = sum_{n < nt} w[n][j] * g[i][n][rule];   w: [SK_HK_K][SK_HK_NJ] complex
SK_HD void sk_hk_cell_combine(const sk_cplx *w, int nt, const sk_cplx *gpoint /* &grid[((l0+i) K + 0) 2 + rule] */, sk_cplx *out) {
  for (int j = 0; j < SK_HK_NJ; ++j) out[j].x = out[j].y = 0.0;
  for (int n = 0; n < nt; ++n) {
    const sk_cplx gv = gpoint[n * 2];
#pragma unroll
    for (int j = 0; j < SK_HK_NJ; ++j) {
      const sk_cplx ww = w[n * SK_HK_NJ + j];
      out[j].x = sk_fma(ww.x, gv.x, sk_fma(-ww.y, gv.y, out[j].x));
      out[j].y = sk_fma(ww.x, gv.y, sk_fma(ww.y, gv.x, out[j].y));
    }
  }
}

// coefficient q, component comp of the cell polynomial before the deconvolution fold; cj: [SK_HK_NJ][SK_NC][4]
SK_HD double sk_hk_cell_shift_add(const double *cj, int q, int comp) {
  double v = cj[q * 4 + comp];
  if (q >= 1) v += cj[(1 * SK_NC + (q - 1)) * 4 + comp];
  if (q >= 2) v += cj[(2 * SK_NC + (q - 2)) * 4 + comp];
  if (q >= 3) v += cj[(3 * SK_NC + (q - 3)) * 4 + comp];
  return v;
}

// the whole cell polynomial in one place (plain loops; the warp kernel distributes exactly these items)
template <int W>
SK_HD void sk_hk_cell_build(const SkEsPlan &P, const SkHankelPlan &H, const SkHankelGroup &g, const sk_cplx *grid, long long l0,
                            const double *E, const double *O, double *coef /*[SK_NC][4]*/) {
  const SkHkCell c = sk_hk_cell_setup<W>(H, g, l0);
  sk_cplx w[SK_HK_K * SK_HK_NJ];
  for (int n = 0; n < c.nt; ++n)
    for (int j = 0; j < SK_HK_NJ; ++j) sk_hk_cell_weight(H, c, n, j, &w[n * SK_HK_NJ + j].x, &w[n * SK_HK_NJ + j].y);
  double gc[SK_HK_NJ][W][4];
  for (int i = 0; i < W; ++i)
    for (int rule = 0; rule < 2; ++rule) {
      sk_cplx o[SK_HK_NJ];
      sk_hk_cell_combine(w, c.nt, grid + ((size_t)(l0 + i) * SK_HK_K) * 2 + rule, o);
      for (int j = 0; j < SK_HK_NJ; ++j) { gc[j][i][rule * 2] = o[j].x; gc[j][i][rule * 2 + 1] = o[j].y; }
    }
  double cj[SK_HK_NJ * SK_NC * 4];
  for (int j = 0; j < SK_HK_NJ; ++j)
    for (int q = 0; q < SK_NC; ++q)
      for (int comp = 0; comp < 4; ++comp) cj[(j * SK_NC + q) * 4 + comp] = sk_cell_coef<W>(E, O, &gc[j][0][comp], 4, q);
  for (int q = 0; q < SK_NC; ++q)
    for (int comp = 0; comp < 4; ++comp) coef[q * 4 + comp] = sk_hk_cell_shift_add(cj, q, comp);
  double a[4];
  sk_cell_deconv_cubic(P, g.G, c.ymid, a);
  for (int comp = 0; comp < 4; ++comp) sk_cell_fold(coef + comp, 4, a);
}

// one target through its cell polynomial (asymptotic part only): Horner, post-phase exp(2 pi i wc r) with the
// product formed exactly and the lean table sincos of K4 (tab: sk_sincos2pi_table_fill), real part
SK_HD void sk_hk_cell_eval(const double *coef, const sk_cplx *tab, const SkGeom &G, double r, double s, double *out) {
  double a[4];
  sk_cell_horner<4>(coef, s, a);
  double sn, cs;
  sk_sincos2pi(tab, sk_frac_prod(G.wc, r, 0.0), &sn, &cs);
  out[0] = sk_fma(a[0], cs, -sk_mul(a[1], sn));
  out[1] = sk_fma(a[2], cs, -sk_mul(a[3], sn));
}
