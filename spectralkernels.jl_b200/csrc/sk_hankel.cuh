// sk_hankel.cuh -- sm_100a kernels of the O(N) nonuniform Hankel transform (dim >= 2 branch of the K(r)
// path: src/quadrature.jl:137-161, where the reference calls FastHankelTransform.jl's `nufht`).
// The arithmetic is in sk_hankel.h; these are the launch wrappers plus block-level cooperation.
//
//   K9a k_hankel_levels   first source of every dyadic frequency level (both rules)
//   K9b k_hankel_fit      direct sums of each level at the Chebyshev nodes of its local interval
//       k_hankel_cheb     node values -> Chebyshev coefficients
//   K9c k_hankel_prep     per group: grid positions, term-0 strengths, term ratios
//       k_spread_hankel   deterministic gather spread of all K terms of a rule in one pass + mode
//                         deconvolution + zero pad (taps evaluated once per source, not once per term)
//       cuFFT Z2Z         batch 2K interleaved (sk_api.cu)
//   K9d k_hankel_interp   per target: octave -> group; w taps once, K x 2 grids, Horner in i/z_ref, local
//                         Chebyshev levels, *c, / x^(dim/2-1), stage (I2, |I2-I1|), block max
#pragma once
#include "sk_hankel.h"
#include "sk_kernels.cuh"

__global__ void k_hankel_levels(double wT, const double *__restrict__ no1, long long M1, const double *__restrict__ no2,
                                long long M2, long long *__restrict__ lev_start /*[2][SK_HK_NLEV + 1]*/) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M1 + M2) return;
  const int rule = t < M1 ? 0 : 1;
  const long long k = rule ? t - M1 : t, M = rule ? M2 : M1;
  const double *no = rule ? no2 : no1;
  long long *ls = lev_start + rule * (SK_HK_NLEV + 1);
  const int l = sk_hk_level(wT, no[k]);
  const int lp = k > 0 ? sk_hk_level(wT, no[k - 1]) : -1;
  for (int q = lp + 1; q <= l; ++q) ls[q] = k;                  // nodes ascend, so levels do too
  if (k == M - 1)
    for (int q = l + 1; q <= SK_HK_NLEV; ++q) ls[q] = M;
}

// grid (SK_HK_NCH nodes, levels q_lo..q_hi, 2 rules); fixed-order tree reduction (bitwise reproducible)
__global__ void __launch_bounds__(256)
k_hankel_fit(const __grid_constant__ SkHankelPlan H, const double *__restrict__ tab, const double *__restrict__ no1,
             const double *__restrict__ buf1, const double *__restrict__ no2, const double *__restrict__ buf2,
             const long long *__restrict__ lev_start, double *__restrict__ vals /*[2][SK_HK_NLEV][SK_HK_NCH]*/) {
  const int i = blockIdx.x, q = H.q_lo + blockIdx.y, rule = blockIdx.z;
  const double *no = rule ? no2 : no1;
  const double *buf = rule ? buf2 : buf1;
  const long long s0 = lev_start[rule * (SK_HK_NLEV + 1) + q], s1 = lev_start[rule * (SK_HK_NLEV + 1) + q + 1];
  const double rho = 0.5 * sk_hk_level_radius(H.r_hi, q) * (sk_hk_cheb_node(i) + 1.0);
  double acc = 0.0;
  for (long long k = s0 + threadIdx.x; k < s1; k += blockDim.x) acc += sk_hk_fit_term(tab, H.nu, no[k], buf[k], rho);
  __shared__ double sr[256];
  sr[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sr[threadIdx.x] += sr[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) vals[((size_t)rule * SK_HK_NLEV + q) * SK_HK_NCH + i] = sr[0];
}

__global__ void k_hankel_cheb(const __grid_constant__ SkHankelPlan H, const double *__restrict__ vals,
                              double *__restrict__ cheb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = H.q_hi - H.q_lo + 1;
  if (t >= 2 * nq * SK_HK_NCH) return;
  const int m = t % SK_HK_NCH, q = H.q_lo + (t / SK_HK_NCH) % nq, rule = t / (SK_HK_NCH * nq);
  const size_t base = ((size_t)rule * SK_HK_NLEV + q) * SK_HK_NCH;
  cheb[base + m] = sk_hk_cheb_coef(vals + base, m);
}

struct SkHkSrc {
  const double *no[2];
  const double *buf[2];
  double *pos_hi[2];
  double *pos_lo[2];
  sk_cplx *cs[2];
  double *lam[2];
  long long M[2];
};

__global__ void k_hankel_prep(const SkHankelGroup *__restrict__ groups, int gi, double wT, const __grid_constant__ SkHkSrc S) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S.M[0] + S.M[1]) return;
  const int rule = t < S.M[0] ? 0 : 1;
  const long long k = rule ? t - S.M[0] : t;
  sk_hk_source_prep(groups[gi], wT, S.no[rule][k], S.buf[rule][k], &S.pos_hi[rule][k], &S.pos_lo[rule][k], &S.cs[rule][k],
                    &S.lam[rule][k]);
}

// Same deterministic gather as k_spread_modes (sk_kernels.cuh), carrying the K terms of the expansion: the
// strength of term n+1 is the strength of term n times lam_k * ratio[n].  Output layout [nf2][K][2 rules].
template <int W>
__global__ void __launch_bounds__(256)
k_spread_hankel(const __grid_constant__ SkEsPlan P, const SkHankelGroup *__restrict__ groups, int gi,
                const __grid_constant__ SkHankelPlan H, const __grid_constant__ SkHkSrc src, sk_cplx *__restrict__ grid_all) {
  const SkGeom G = groups[gi].G;
  sk_cplx *__restrict__ fft_io = grid_all + groups[gi].grid_off;
  const int r = blockIdx.y;
  const double *__restrict__ ph = src.pos_hi[r];
  const double *__restrict__ pl = src.pos_lo[r];
  const sk_cplx *__restrict__ cs = src.cs[r];
  const double *__restrict__ lam = src.lam[r];
  const long long M = src.M[r];
  const double half = 0.5 * W;
  const long long l_blk = (long long)blockIdx.x * SK_SPREAD_CELLS;
  const long long n_blk = l_blk - G.nf / 2;
  __shared__ long long s_range[2];
  __shared__ int s_l0[SK_SPREAD_CHUNK];
  __shared__ sk_cplx s_cs[SK_SPREAD_CHUNK];
  __shared__ double s_lam[SK_SPREAD_CHUNK];
  __shared__ double s_tap[SK_SPREAD_CHUNK][W + 1];
  if (threadIdx.x < 32) {
    const long long v = sk_warp_lower_bound(ph, M, (double)n_blk - half - 1e-6, false);
    if (threadIdx.x == 0) s_range[0] = v;
  } else if (threadIdx.x < 64) {
    const long long v = sk_warp_lower_bound(ph, M, (double)(n_blk + SK_SPREAD_CELLS - 1) + half + 1e-6, true);
    if (threadIdx.x == 32) s_range[1] = v;
  }
  __syncthreads();
  const long long s0 = s_range[0], s1 = s_range[1];
  const int sub = threadIdx.x & (SK_SPREAD_LANES - 1);
  const int cell = threadIdx.x / SK_SPREAD_LANES;
  const long long l = l_blk + cell;
  const bool live = l < G.nf;
  double ar[SK_HK_K], ai[SK_HK_K];
#pragma unroll
  for (int n = 0; n < SK_HK_K; ++n) ar[n] = ai[n] = 0.0;
  for (long long c0 = s0; c0 < s1; c0 += SK_SPREAD_CHUNK) {
    const int ns = (int)((s1 - c0) < (long long)SK_SPREAD_CHUNK ? (s1 - c0) : (long long)SK_SPREAD_CHUNK);
    if ((int)threadIdx.x < ns) {
      const long long k = c0 + threadIdx.x;
      const double p_hi = ph[k], p_lo = pl[k];
      const double c_first = ceil(p_hi - half);
      const double x0 = (c_first - p_hi) - p_lo;
      double taps[W];
      sk_es_taps<W>(P, 2.0 * (x0 + (half - 0.5)), taps);
#pragma unroll
      for (int i = 0; i < W; ++i) s_tap[threadIdx.x][i] = taps[i];
      s_l0[threadIdx.x] = (int)((long long)c_first - n_blk);
      s_cs[threadIdx.x] = cs[k];
      s_lam[threadIdx.x] = lam[k];
    }
    __syncthreads();
    if (live) {
      int a = 0, b = ns;
      const int want = cell - W + 1;
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (s_l0[mid] < want) a = mid + 1; else b = mid;
      }
      for (int k = a + sub; k < ns; k += SK_SPREAD_LANES) {
        const int i = cell - s_l0[k];
        if (i < 0) break;
        const double wgt = s_tap[k][i];
        const sk_cplx c = s_cs[k];
        const double lk = s_lam[k];
        double cx = wgt * c.x, cy = wgt * c.y;
#pragma unroll
        for (int n = 0; n < SK_HK_K; ++n) {
          ar[n] += cx;
          ai[n] += cy;
          const double f = lk * H.ratio[n];
          cx *= f;
          cy *= f;
        }
      }
    }
    __syncthreads();
  }
  double q = 0.0;
  long long jout = 0;
  if (live) {
    const long long n = l - G.nf / 2;
    q = sk_deconv(P, G.t_cell * fabs((double)n));
    if (n & 1) q = -q;
    jout = n >= 0 ? n : n + G.nf2;
  }
#pragma unroll
  for (int n = 0; n < SK_HK_K; ++n) {
    double vr = ar[n], vi = ai[n];
#pragma unroll
    for (int o = SK_SPREAD_LANES / 2; o > 0; o >>= 1) {
      vr += __shfl_xor_sync(0xffffffffu, vr, o);
      vi += __shfl_xor_sync(0xffffffffu, vi, o);
    }
    if (sub == 0 && live) {
      sk_cplx o;
      o.x = vr * q;
      o.y = vi * q;
      fft_io[(jout * SK_HK_K + n) * 2 + r] = o;
    }
  }
}

template <int W>
__global__ void __launch_bounds__(256)
k_hankel_interp(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkHankelPlan H,
                const SkHankelGroup *__restrict__ groups, const sk_cplx *__restrict__ grid, const double *__restrict__ cheb,
                const double *__restrict__ xs, long long n, double cmul, double xdiv, sk_cplx *__restrict__ stage,
                SkReduceOut *__restrict__ red) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  unsigned int fl = 0;
  if (j < n) {
    const double x = xs[j];
    double f[2];
    sk_hk_point<W>(P, H, groups, grid, cheb, x, f);
    // *c, then / x^(dim/2 - 1)   (src/quadrature.jl:250-254)
    double i1 = sk_mul(f[0], cmul), i2 = sk_mul(f[1], cmul);
    if (xdiv != 0.0) {
      const double den = pow(x, xdiv);
      i1 = i1 / den;
      i2 = i2 / den;
    }
    sk_stage(i1, i2, 1.0, &stage[j], d, fl);
  }
  sk_block_reduce_maxflags(d, fl, red);
}
