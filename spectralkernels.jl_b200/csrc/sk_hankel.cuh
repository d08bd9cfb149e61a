// sk_hankel.cuh -- sm_100a kernels of the O(N) nonuniform Hankel transform (dim >= 2 branch of the K(r)
// path: src/quadrature.jl:137-161, where the reference calls FastHankelTransform.jl's `nufht`).
// The arithmetic is in sk_hankel.h; these are the launch wrappers plus block-level cooperation.
//
//   K9a k_hankel_levels      first source of every dyadic frequency level (both rules)
//   K9b k_hankel_fit         direct sums of each level at the Chebyshev nodes of its local interval
//       k_hankel_cheb        node values -> Chebyshev coefficients of the level
//       k_hankel_local_poly  levels an octave needs -> piecewise expansion of the octave (16 pieces x 16 terms)
//   K9c k_hankel_prep        per group: grid positions, term-0 strengths, term ratios (levels [q_from, q_to))
//       k_spread_hankel      deterministic gather spread of all K terms of a rule in one pass + mode deconvolution
//                            + zero pad (taps evaluated once per source, not once per term); split-K over the source
//                            range for small grids (k_spread_hankel_reduce); accumulating into the running mode
//                            buffer of the shared set of small octaves
//       cuFFT Z2Z            batch 2K interleaved (sk_api.cu)
//   K9d k_hankel_cells       (default) one polynomial per fine-grid cell across the K terms, built per warp; four
//                            Horner chains per target + the octave's local expansion, *c, / x^(dim/2-1), stage
//                            (I2, |I2-I1|), max
//       k_hankel_interp      A/B: one target per thread, per-target taps (the plain restatement of sk_hk_point)
//       k_hankel_interp2     A/B: two targets per thread sharing 256-bit loads; bit-identical to k_hankel_interp
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

// Direct sums of each level at the SK_HK_NCH Chebyshev nodes of its local interval.  grid (SK_HK_NCH nodes, levels
// q_lo..q_hi, 2 rules x SK_HK_FITSPLIT slices of the level's sources): all threads of a block share the node, so
// consecutive sources fall into the same interval of the Bessel table (one L1 wavefront per coefficient load).
// Fixed-order tree reduction per slice, the slices are added in order by k_hankel_cheb: bitwise reproducible.
// (Measured alternatives, profiles/r1_k: one block per source slice with a thread per node is 8x slower -- the
// lanes of a warp then hit 32 different table rows; a warp per 8 nodes with lanes over sources is no faster.)
#define SK_HK_FITSPLIT 8
__global__ void __launch_bounds__(256)
k_hankel_fit(const __grid_constant__ SkHankelPlan H, const double *__restrict__ tab, const double *__restrict__ no1,
             const double *__restrict__ buf1, const double *__restrict__ no2, const double *__restrict__ buf2,
             const long long *__restrict__ lev_start, double *__restrict__ vals /*[FITSPLIT][2][SK_HK_NLEV][SK_HK_NCH]*/) {
  const int i = blockIdx.x, q = H.q_lo + blockIdx.y, rule = blockIdx.z & 1, z = blockIdx.z >> 1;
  const double *no = rule ? no2 : no1;
  const double *buf = rule ? buf2 : buf1;
  const long long l0 = lev_start[rule * (SK_HK_NLEV + 1) + q], l1 = lev_start[rule * (SK_HK_NLEV + 1) + q + 1];
  const long long len = (l1 - l0 + SK_HK_FITSPLIT - 1) / SK_HK_FITSPLIT;
  const long long s0 = l0 + z * len, s1 = (s0 + len) < l1 ? (s0 + len) : l1;
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
  if (threadIdx.x == 0) vals[(((size_t)z * 2 + rule) * SK_HK_NLEV + q) * SK_HK_NCH + i] = sr[0];
}

// node values (slices added in order) -> Chebyshev coefficients; grid (levels, 2 rules), SK_HK_NCH threads
__global__ void __launch_bounds__(SK_HK_NCH)
k_hankel_cheb(const __grid_constant__ SkHankelPlan H, const double *__restrict__ vals, double *__restrict__ cheb) {
  const int q = H.q_lo + blockIdx.x, rule = blockIdx.y, i = threadIdx.x;
  __shared__ double v[SK_HK_NCH];
  double a = 0.0;
  for (int z = 0; z < SK_HK_FITSPLIT; ++z) a += vals[(((size_t)z * 2 + rule) * SK_HK_NLEV + q) * SK_HK_NCH + i];
  v[i] = a;
  __syncthreads();
  cheb[((size_t)q * SK_HK_NCH + i) * 2 + rule] = sk_hk_cheb_coef(v, i);       // layout [NLEV][NCH][2 rules]
}

// per-octave piecewise expansions (sk_hk_local2) from the per-level coefficients: grid (SK_HK_NSUB, q_hi + 1);
// 256 threads = 16 nodes x 16 level lanes: every level's interpolant is evaluated by its own thread, the levels
// are then added in level order (the order of sk_hk_local)
__global__ void __launch_bounds__(256)
k_hankel_local_poly(const __grid_constant__ SkHankelPlan H, const double *__restrict__ cheb, double *__restrict__ loc) {
  const int s = blockIdx.x, tt = blockIdx.y;
  __shared__ double part[SK_HK_NLEV][SK_HK_NLOC][2];
  __shared__ double vals[SK_HK_NLOC * 2];
  const int i = threadIdx.x & (SK_HK_NLOC - 1), ql = threadIdx.x / SK_HK_NLOC;
  const int t_eff = tt < H.q_hi ? tt : 4096;
  const int qe = (t_eff + 1 < H.q_hi) ? t_eff + 1 : H.q_hi;
  const double r = sk_hk_local_node(H, tt, s, i);
  for (int q = H.q_lo + ql; q <= qe; q += 16) sk_hk_local_level(H, cheb, r, q, part[q][i]);
  __syncthreads();
  if (threadIdx.x < SK_HK_NLOC * 2) {
    const int ii = threadIdx.x >> 1, rule = threadIdx.x & 1;
    double a = 0.0;
    for (int q = H.q_lo; q <= qe; ++q) a += part[q][ii][rule];
    vals[2 * ii + rule] = a;
  }
  __syncthreads();
  if (threadIdx.x < SK_HK_NLOC * 2) {
    const int m = threadIdx.x >> 1, rule = threadIdx.x & 1;
    loc[(((size_t)tt * SK_HK_NSUB + s) * SK_HK_NLOC + m) * 2 + rule] = sk_hk_local_coef(vals + rule, m);
  }
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
  sk_hk_source_prep(groups[gi], wT, groups[gi].q_from, groups[gi].q_to, S.no[rule][k], S.buf[rule][k], &S.pos_hi[rule][k],
                    &S.pos_lo[rule][k], &S.cs[rule][k], &S.lam[rule][k]);
}

// Same deterministic gather as k_spread_modes (sk_kernels.cuh), carrying the K terms of the expansion: the
// strength of term n+1 is the strength of term n times lam_k * ratio[n].  Output layout [nf2][K][2 rules].
template <int W>
__global__ void __launch_bounds__(256)
k_spread_hankel(const __grid_constant__ SkEsPlan P, const SkHankelGroup *__restrict__ groups, int gi,
                const __grid_constant__ SkHankelPlan H, const __grid_constant__ SkHkSrc src,
                const long long *__restrict__ lev_start, sk_cplx *__restrict__ fft_io /*FFT input of the group*/, int accumulate,
                sk_cplx *__restrict__ part /*[gridDim.z][nf][K][2] when gridDim.z > 1*/) {
  const SkGeom G = groups[gi].G;
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
  // small grids hold many sources per cell: gridDim.z blocks share a cell block's source range (split-K); their
  // partial sums are added in slice order by k_spread_hankel_reduce, so the result stays reproducible
  long long s0 = s_range[0], s1 = s_range[1];
  {
    // only the sources of the group's levels [q_from, q_to) (everything else has strength 0)
    const long long k_lo = lev_start[r * (SK_HK_NLEV + 1) + groups[gi].q_from];
    const long long k_hi = lev_start[r * (SK_HK_NLEV + 1) + groups[gi].q_to];
    s0 = s0 > k_lo ? s0 : k_lo;
    s1 = s1 < k_hi ? s1 : k_hi;
    if (s1 < s0) s1 = s0;
  }
  if (gridDim.z > 1) {
    const long long len = (s1 - s0 + gridDim.z - 1) / gridDim.z;
    s0 += (long long)blockIdx.z * len;
    s1 = (s0 + len) < s1 ? (s0 + len) : s1;
  }
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
      if (gridDim.z > 1) {
        o.x = vr;
        o.y = vi;
        part[(((size_t)blockIdx.z * G.nf + l) * SK_HK_K + n) * 2 + r] = o;
      } else {
        o.x = vr * q;
        o.y = vi * q;
        sk_cplx *dst = &fft_io[(jout * SK_HK_K + n) * 2 + r];
        if (accumulate) {                   // running modes of a shared set: one writer per element, fixed order
          const sk_cplx prev = *dst;
          o.x += prev.x;
          o.y += prev.y;
        }
        *dst = o;
      }
    }
  }
}

// second half of the split-K spread: add the slices in order, deconvolve the mode, write the FFT input
__global__ void k_spread_hankel_reduce(const __grid_constant__ SkEsPlan P, const SkHankelGroup *__restrict__ groups, int gi,
                                       int nsplit, const sk_cplx *__restrict__ part, sk_cplx *__restrict__ fft_io,
                                       int accumulate) {
  const SkGeom G = groups[gi].G;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.nf * (SK_HK_K * 2)) return;
  const long long l = t / (SK_HK_K * 2);
  const int nr = (int)(t % (SK_HK_K * 2));
  double vr = 0.0, vi = 0.0;
  for (int z = 0; z < nsplit; ++z) {
    const sk_cplx v = part[((size_t)z * G.nf + l) * (SK_HK_K * 2) + nr];
    vr += v.x;
    vi += v.y;
  }
  const long long n = l - G.nf / 2;
  double q = sk_deconv(P, G.t_cell * fabs((double)n));
  if (n & 1) q = -q;
  const long long jout = n >= 0 ? n : n + G.nf2;
  sk_cplx o;
  o.x = vr * q;
  o.y = vi * q;
  if (accumulate) {
    const sk_cplx prev = fft_io[jout * (SK_HK_K * 2) + nr];
    o.x += prev.x;
    o.y += prev.y;
  }
  fft_io[jout * (SK_HK_K * 2) + nr] = o;
}

template <int W>
__global__ void __launch_bounds__(256)
k_hankel_interp(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkHankelPlan H,
                const SkHankelGroup *__restrict__ groups, const sk_cplx *__restrict__ grid, const double *__restrict__ cheb,
                const double *__restrict__ xs, long long n, double cmul, double xdiv, sk_cplx *__restrict__ stage,
                SkReduceOut *__restrict__ red, sk_cplx *__restrict__ raw) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  unsigned int fl = 0;
  if (j < n && raw) {           // raw sums of both rules, layout of k_direct_bessel: raw[2j + rule].x
    double f[2];
    sk_hk_point<W>(P, H, groups, grid, cheb, xs[j], f);
    sk_cplx o;
    o.y = 0.0;
    o.x = f[0]; raw[2 * j] = o;
    o.x = f[1]; raw[2 * j + 1] = o;
  } else if (j < n) {
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

// ---- K9d, production variant ---------------------------------------------------------------------------------
// k_hankel_interp above is bound by the L1/LSU path (ncu: l1tex throughput 95 %, FP64 pipe 25 %): one 16-byte
// load per two FMAs.  Here every thread owns TWO consecutive sorted targets; when they share the group and the
// grid window (the usual case: ~100 targets per cell) each grid value is fetched once with one 256-bit load
// and feeds 8 FMAs.
// The arithmetic per target is the sequence of sk_hk_point, operation for operation: results are bit-identical
// to k_hankel_interp whatever the pairing (tests: test_hankel_interp_variants_agree).
__device__ __forceinline__ void sk_ld256_nc(const void *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

template <int W>
__device__ __forceinline__ void sk_hk_interp_pair(const SkEsPlan &P, const SkHankelPlan &H, const SkHankelGroup &g,
                                                  const sk_cplx *grid, const SkTargetCoord &tA, const SkTargetCoord &tB,
                                                  double rA, double rB, double *outA, double *outB) {
  double tapA[W], tapB[W];
  sk_es_taps<W>(P, tA.s, tapA);
  sk_es_taps<W>(P, tB.s, tapB);
  const double izA = 1.0 / sk_mul(sk_mul(6.283185307179586, g.w_ref), rA);
  const double izB = 1.0 / sk_mul(sk_mul(6.283185307179586, g.w_ref), rB);
  double cA[4] = {0.0, 0.0, 0.0, 0.0}, cB[4] = {0.0, 0.0, 0.0, 0.0};      // (rule0 re, im, rule1 re, im)
  const sk_cplx *gp = grid + (size_t)tA.l0 * (SK_HK_K * 2);
  // each target keeps its own leading block of terms (sk_hk_nterms), exactly as when evaluated alone
  const int nA = sk_hk_nterms(H, sk_mul(sk_mul(6.283185307179586, g.w_ref), rA));
  const int nB = sk_hk_nterms(H, sk_mul(sk_mul(6.283185307179586, g.w_ref), rB));
#pragma unroll 1
  for (int n = (nA > nB ? nA : nB) - 1; n >= 0; --n) {
    double aA[4] = {0.0, 0.0, 0.0, 0.0}, aB[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < W; ++i) {
      double v0, v1, v2, v3;
      sk_ld256_nc(gp + (i * SK_HK_K + n) * 2, v0, v1, v2, v3);
      aA[0] = sk_fma(tapA[i], v0, aA[0]); aA[1] = sk_fma(tapA[i], v1, aA[1]);
      aA[2] = sk_fma(tapA[i], v2, aA[2]); aA[3] = sk_fma(tapA[i], v3, aA[3]);
      aB[0] = sk_fma(tapB[i], v0, aB[0]); aB[1] = sk_fma(tapB[i], v1, aB[1]);
      aB[2] = sk_fma(tapB[i], v2, aB[2]); aB[3] = sk_fma(tapB[i], v3, aB[3]);
    }
    if (n < nA) {
      const double a0 = sk_fma(-cA[1], izA, aA[0]), a1 = sk_fma(cA[0], izA, aA[1]);
      const double a2 = sk_fma(-cA[3], izA, aA[2]), a3 = sk_fma(cA[2], izA, aA[3]);
      cA[0] = a0; cA[1] = a1; cA[2] = a2; cA[3] = a3;
    }
    if (n < nB) {
      const double b0 = sk_fma(-cB[1], izB, aB[0]), b1 = sk_fma(cB[0], izB, aB[1]);
      const double b2 = sk_fma(-cB[3], izB, aB[2]), b3 = sk_fma(cB[2], izB, aB[3]);
      cB[0] = b0; cB[1] = b1; cB[2] = b2; cB[3] = b3;
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const SkTargetCoord &t = u ? tB : tA;
    const double r = u ? rB : rA, iz = u ? izB : izA;
    const double *c = u ? cB : cA;
    const double qf = sk_deconv(P, g.G.t_cell * t.yabs);
    double sn, cs;
    sk_post_phase(g.G, r, &sn, &cs);
    const double er = sk_fma(cs, H.cphi, sn * H.sphi), ei = sk_fma(sn, H.cphi, -cs * H.sphi);
    const double amp = qf * sqrt(0.6366197723675814 * iz);
    double *o = u ? outB : outA;
    o[0] = amp * sk_fma(c[0], er, -c[1] * ei);
    o[1] = amp * sk_fma(c[2], er, -c[3] * ei);
  }
}

#define SK_HK_TPB2 128       // ~160 registers per thread: 3 blocks of 128 threads per SM
template <int W>
__global__ void __launch_bounds__(SK_HK_TPB2, 3)
k_hankel_interp2(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkHankelPlan H,
                 const SkHankelGroup *__restrict__ groups, const sk_cplx *__restrict__ grid, const double *__restrict__ cheb,
                 const double *__restrict__ xs, long long n, double cmul, double xdiv, sk_cplx *__restrict__ stage,
                 SkReduceOut *__restrict__ red, sk_cplx *__restrict__ raw) {
  __shared__ unsigned long long s_max;
  __shared__ unsigned int s_fl, s_cnt;
  if (threadIdx.x == 0) { s_max = 0ull; s_fl = 0u; s_cnt = 0u; }
  __syncthreads();
  const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  double d = 0.0;
  unsigned int fl = 0;
  if (j < n) {
    const bool two = j + 1 < n;
    const double xA = xs[j], xB = two ? xs[j + 1] : xA;
    const int tA = sk_hk_octave(H.r_hi, xA), tB = sk_hk_octave(H.r_hi, xB);
    int gA = sk_hk_group_of_octave(H, tA), gB = sk_hk_group_of_octave(H, tB);
    if (gA >= H.ngroups) gA = -1;
    if (gB >= H.ngroups) gB = -1;
    double fA[2] = {0.0, 0.0}, fB[2] = {0.0, 0.0}, lA[2], lB[2];
    sk_hk_local2(H, cheb, xA, tA, lA);
    sk_hk_local2(H, cheb, xB, tB, lB);
    bool paired = false;
    if (gA >= 0 && gA == gB) {
      const SkTargetCoord cA = sk_target_coord<W>(groups[gA].G, xA), cB = sk_target_coord<W>(groups[gA].G, xB);
      if (cA.l0 == cB.l0) {
        sk_hk_interp_pair<W>(P, H, groups[gA], grid + groups[gA].grid_off, cA, cB, xA, xB, fA, fB);
        paired = true;
      }
    }
    if (!paired) {
      if (gA >= 0) sk_hk_interp_point<W>(P, H, groups[gA], grid + groups[gA].grid_off, xA, fA);
      if (gB >= 0) sk_hk_interp_point<W>(P, H, groups[gB], grid + groups[gB].grid_off, xB, fB);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const double x = u ? xB : xA;
      const double f0 = (u ? fB[0] : fA[0]) + (u ? lB[0] : lA[0]), f1 = (u ? fB[1] : fA[1]) + (u ? lB[1] : lA[1]);
      if (raw) {
        sk_cplx o;
        o.y = 0.0;
        o.x = f0; raw[2 * (j + u)] = o;
        o.x = f1; raw[2 * (j + u) + 1] = o;
        continue;
      }
      double i1 = sk_mul(f0, cmul), i2 = sk_mul(f1, cmul);
      if (xdiv != 0.0) {
        const double den = pow(x, xdiv);
        i1 = i1 / den;
        i2 = i2 / den;
      }
      sk_stage(i1, i2, 1.0, &stage[j + u], d, fl);
    }
  }
  // warps finish at very different times (octaves, slow pairs): no block barrier -- every warp folds its maximum
  // into shared memory and the last one to arrive publishes the block's result
  d = sk_warp_max(d);
  fl = __reduce_or_sync(0xffffffffu, fl);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&s_max, (unsigned long long)__double_as_longlong(d));
    if (fl) atomicOr(&s_fl, fl);
    __threadfence_block();
    if (atomicAdd(&s_cnt, 1u) == (blockDim.x >> 5) - 1) {
      atomicMax(&red->maxbits, atomicMax(&s_max, 0ull));
      const unsigned int f = atomicOr(&s_fl, 0u);
      if (f) atomicOr(&red->flags, f);
    }
  }
}

// ---- K9d, cell polynomials across the K terms (default) ------------------------------------------------------
// See sk_hankel.h ("cell polynomials across the K terms").  One warp owns SK_HK_CT x 32 consecutive sorted targets
// and walks them 32 at a time; the lanes whose targets share a cell build that cell's polynomial together in the
// warp's shared-memory scratch (skipped when the cell is the one already there -- consecutive targets mostly are
// in the same cell) and then evaluate their targets with four Horner chains.  No block-level cooperation: which
// targets share a warp never changes the arithmetic of a target, so results are independent of the tiling and
// of how the targets are sharded over GPUs.  Targets too close to r = 0 for the truncated binomial series
// (SkHkCell::ok == 0: a few cells of the merged group, and the groups of tiny octaves) take sk_hk_interp_point.
#define SK_HK_CT 8
struct SkHkWarpScratch {
  sk_cplx w[SK_HK_K * SK_HK_NJ];          // weights of term n in power j
  double gc[SK_HK_NJ][16][4];             // grid values combined over n, per power j
  double cj[SK_HK_NJ * SK_NC * 4];        // polynomial coefficients per power j
  double coef[SK_NC * 4];                 // the cell polynomial
  double dq[4];                           // deconvolution factor at the 4 Chebyshev nodes of the cell
};

template <int W>
__global__ void __launch_bounds__(256)
k_hankel_cells(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkHankelPlan H,
               const SkHankelGroup *__restrict__ groups, const sk_cplx *__restrict__ grid, const double *__restrict__ loc,
               const double *__restrict__ xs, long long n, double cmul, double xdiv, sk_cplx *__restrict__ stage,
               SkReduceOut *__restrict__ red, sk_cplx *__restrict__ raw) {
  static_assert(W == 16, "the lane <-> (window point, rule) map of the build assumes 16 taps");
  __shared__ double sE[(W / 2) * (SK_NC / 2)], sO[(W / 2) * (SK_NC / 2)];
  __shared__ SkHkWarpScratch sW[8];
  __shared__ sk_cplx sTab[65];
  __shared__ unsigned long long s_max;
  __shared__ unsigned int s_fl, s_cnt;
  for (int t = threadIdx.x; t < (W / 2) * (SK_NC / 2); t += blockDim.x) {
    sE[t] = P.E[t / (SK_NC / 2)][t % (SK_NC / 2)];
    sO[t] = P.O[t / (SK_NC / 2)][t % (SK_NC / 2)];
  }
  if (threadIdx.x == 0) { s_max = 0ull; s_fl = 0u; s_cnt = 0u; }
  if (threadIdx.x >= 128 && threadIdx.x < 128 + 65) sk_sincos2pi_table_fill(sTab, threadIdx.x - 128);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  SkHkWarpScratch &S = sW[wid];
  const long long base = ((long long)blockIdx.x * 8 + wid) * (32 * SK_HK_CT);
  int cur_g = -1;
  long long cur_l0 = -1;
  double d = 0.0;
  unsigned int fl = 0;
  // the distance of the NEXT 32-target step is loaded one step ahead: its latency hides behind this step's work
  double x_next = (base + lane < n) ? xs[base + lane] : 0.0;
#pragma unroll 1
  for (int u = 0; u < SK_HK_CT; ++u) {
    const long long j = base + (long long)u * 32 + lane;
    const bool have = j < n;
    const double x = x_next;
    if (u + 1 < SK_HK_CT) x_next = (j + 32 < n) ? xs[j + 32] : 0.0;
    double f[2] = {0.0, 0.0}, lo[2] = {0.0, 0.0};
    int gi = -1;
    SkTargetCoord tc;
    tc.l0 = -1;
    tc.s = 0.0;
    bool cellpath = false;
    if (have) {
      const int t = sk_hk_octave(H.r_hi, x);
      sk_hk_local2(H, loc, x, t, lo);
      gi = sk_hk_group_of_octave(H, t);
      if (gi >= H.ngroups) gi = -1;
      if (gi >= 0) {
        tc = sk_target_coord<W>(groups[gi].G, x);
        cellpath = sk_hk_cell_ok<W>(groups[gi], tc.l0);
        if (!cellpath) sk_hk_interp_point<W>(P, H, groups[gi], grid + groups[gi].grid_off, x, f);
      }
    }
    unsigned int remaining = __ballot_sync(0xffffffffu, cellpath);
    while (remaining) {
      const int leader = __ffs(remaining) - 1;
      const int gL = __shfl_sync(0xffffffffu, gi, leader);
      const long long cellL = __shfl_sync(0xffffffffu, tc.l0, leader);
      const unsigned int grp = __ballot_sync(0xffffffffu, cellpath && gi == gL && tc.l0 == cellL) & remaining;
      if (gL != cur_g || cellL != cur_l0) {
        // ---- build the polynomial of cell (gL, cellL): all 32 lanes ----
        const SkHankelGroup &g = groups[gL];
        const sk_cplx *gg = grid + g.grid_off;
        const SkHkCell c = sk_hk_cell_setup<W>(H, g, cellL);
        __syncwarp();                                             // the previous cell's evaluations are done
        for (int it = lane; it < c.nt * SK_HK_NJ; it += 32)
          sk_hk_cell_weight(H, c, it / SK_HK_NJ, it % SK_HK_NJ, &S.w[it].x, &S.w[it].y);
        if (lane < 4) {
          const double node = (lane == 0) ? 0.9238795325112867 : (lane == 1) ? 0.3826834323650898
                            : (lane == 2) ? -0.3826834323650898 : -0.9238795325112867;
          S.dq[lane] = sk_deconv(P, g.G.t_cell * fabs(c.ymid - 0.5 * node));
        }
        __syncwarp();
        {
          const int i = lane >> 1, rule = lane & 1;              // one (window point, rule) per lane
          sk_cplx o[SK_HK_NJ];
          sk_hk_cell_combine(S.w, c.nt, gg + ((size_t)(cellL + i) * SK_HK_K) * 2 + rule, o);
#pragma unroll
          for (int jj = 0; jj < SK_HK_NJ; ++jj) { S.gc[jj][i][rule * 2] = o[jj].x; S.gc[jj][i][rule * 2 + 1] = o[jj].y; }
        }
        __syncwarp();
        for (int it = lane; it < SK_HK_NJ * SK_NC * 4; it += 32) {
          const int comp = it & 3, q = (it >> 2) & (SK_NC - 1), jj = it / (SK_NC * 4);
          S.cj[it] = sk_cell_coef<W>(sE, sO, &S.gc[jj][0][comp], 4, q);
        }
        __syncwarp();
        for (int it = lane; it < SK_NC * 4; it += 32) S.coef[it] = sk_hk_cell_shift_add(S.cj, it >> 2, it & 3);
        __syncwarp();
        if (lane < 4) {
          double a[4];
          sk_cheb4_to_monomial(S.dq, a);
          sk_cell_fold(S.coef + lane, 4, a);
        }
        __syncwarp();
        cur_g = gL;
        cur_l0 = cellL;
      }
      if (grp & (1u << lane)) sk_hk_cell_eval(S.coef, sTab, groups[gL].G, x, tc.s, f);
      remaining &= ~grp;
    }
    if (have) {
      const double f0 = f[0] + lo[0], f1 = f[1] + lo[1];
      if (raw) {
        sk_cplx o;
        o.y = 0.0;
        o.x = f0; raw[2 * j] = o;
        o.x = f1; raw[2 * j + 1] = o;
      } else {
        double i1 = sk_mul(f0, cmul), i2 = sk_mul(f1, cmul);
        if (xdiv != 0.0) {
          const double den = pow(x, xdiv);
          i1 = i1 / den;
          i2 = i2 / den;
        }
        sk_stage(i1, i2, 1.0, &stage[j], d, fl);
      }
    }
  }
  d = sk_warp_max(d);
  fl = __reduce_or_sync(0xffffffffu, fl);
  if (lane == 0) {
    atomicMax(&s_max, (unsigned long long)__double_as_longlong(d));
    if (fl) atomicOr(&s_fl, fl);
    __threadfence_block();
    if (atomicAdd(&s_cnt, 1u) == (blockDim.x >> 5) - 1) {
      atomicMax(&red->maxbits, atomicMax(&s_max, 0ull));
      const unsigned int ff = atomicOr(&s_fl, 0u);
      if (ff) atomicOr(&red->flags, ff);
    }
  }
}
