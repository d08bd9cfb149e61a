// sk_kernels.cuh -- sm_100a kernels of the K(r) path.  Each kernel is a launch wrapper around the
// per-element functions of sk_math.h plus the block-level cooperation (reductions, atomics).
//
//   K1  k_gen_sources      updatequadbufs!                      src/quadrature.jl:49-95
//   K2  k_prep_sources     source positions / pre-phase         (inside nufft1d3, src/utils.jl:10)
//       k_spread_modes     ES spread (deterministic gather) + mode deconvolution + zero-pad
//   K3  cuFFT Z2Z          (sk_api.cu)
//   K4  k_interp_session   ES interpolation at the targets, target-side deconvolution, post-phase,
//                          Re/Im select, *c, |I2-I1|, block max     src/quadrature.jl:130-136, :250-258
//       k_interp_cplx      same, complex output (Level-0 sk_nufft1d3)
//   K5  k_accept           I += I2; err += |I2-I1|              src/quadrature.jl:260-262
//       k_commit           ks += I; errs += err                 src/adaptive.jl:163-164
//   K6  k_scan, k_scan_add convergence scan                     src/adaptive.jl:183-199
//   K7  k_direct           direct Fourier summation             src/quadrature.jl:113-128
//   K8  k_make_keys, k_flag_heads, k_scatter_unique, k_gather   unique/sort/scatter, src/adaptive.jl:99-120
#pragma once
#include <cuda_runtime.h>

#include "sk_math.h"

#define SK_FLAG_NAN1 1u
#define SK_FLAG_NAN2 2u
#define SK_FLAG_NAND 4u

struct SkReduceOut {            // device scalars written by the reductions
  unsigned long long maxbits;   // bit pattern of max |I2-I1| (non-negative doubles order like integers)
  unsigned int flags;
  unsigned int _pad;
  long long max_unconv;         // highest non-converged 0-based index, or lo-1
};

__device__ __forceinline__ double sk_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- K1 ---------------------------------------------------------------------------------------------
__global__ void k_gen_sources(const __grid_constant__ SkPanelSpec S, const double *__restrict__ leg_no1,
                              const double *__restrict__ leg_wt1, const double *__restrict__ leg_no2,
                              const double *__restrict__ leg_wt2, const double *__restrict__ jac_no1,
                              const double *__restrict__ jac_wt1, const double *__restrict__ jac_no2,
                              const double *__restrict__ jac_wt2, double *__restrict__ no1, double *__restrict__ buf1,
                              double *__restrict__ no2, double *__restrict__ buf2) {
  const long long M1 = (long long)S.m * S.k, M2 = 2 * M1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < M1) {
    sk_gen_source(S, 0, t, leg_no1, leg_wt1, jac_no1, jac_wt1, &no1[t], &buf1[t]);
  } else if (t < M1 + M2) {
    const long long u = t - M1;
    sk_gen_source(S, 1, u, leg_no2, leg_wt2, jac_no2, jac_wt2, &no2[u], &buf2[u]);
  }
}

// ---- K2 ---------------------------------------------------------------------------------------------
// strengths: real (buf_im == nullptr) or split complex
__global__ void k_prep_sources(const __grid_constant__ SkGeom G, long long M, const double *__restrict__ no,
                               const double *__restrict__ buf_re, const double *__restrict__ buf_im,
                               double *__restrict__ pos_hi, double *__restrict__ pos_lo, sk_cplx *__restrict__ cs) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M) return;
  sk_cplx c;
  sk_source_prep(G, no[k], buf_re[k], buf_im ? buf_im[k] : 0.0, &pos_hi[k], &pos_lo[k], &c.x, &c.y);
  cs[k] = c;
}

// One thread per FFT-input element j of rule blockIdx.y.  Gather: each element sums, in source
// order, the sources whose kernel support covers it -- bitwise reproducible, no atomics.
struct SkSpreadSrc {
  const double *pos_hi[2];
  const double *pos_lo[2];
  const sk_cplx *cs[2];
  long long M[2];
};
__global__ void k_spread_modes(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkGeom G,
                               const __grid_constant__ SkSpreadSrc src, int nrule, sk_cplx *__restrict__ fft_io) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (j >= G.nf2) return;
  sk_cplx o;
  sk_spread_mode(P, G, j, src.pos_hi[r], src.pos_lo[r], src.cs[r], src.M[r], &o.x, &o.y);
  fft_io[j * nrule + r] = o;
}

// ---- K4 ---------------------------------------------------------------------------------------------
// Session interpolation: both rules share the taps.  grid layout [nf2][2] (m-rule, 2m-rule).
template <int W>
__global__ void __launch_bounds__(256)
k_interp_session(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkGeom G, const double *__restrict__ xs,
                 long long n, const sk_cplx *__restrict__ grid, double cmul, int kernel_sin,
                 double *__restrict__ stage_i, double *__restrict__ stage_e, SkReduceOut *__restrict__ red) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  unsigned int fl = 0;
  if (j < n) {
    double fre[2], fim[2];
    sk_interp_point<W, 2>(P, G, xs[j], grid, fre, fim);
    // kernel == :cos -> real part, :sin -> imaginary part (src/quadrature.jl:130-136); then *c (:250-251)
    const double i1 = (kernel_sin ? fim[0] : fre[0]) * cmul;
    const double i2 = (kernel_sin ? fim[1] : fre[1]) * cmul;
    d = fabs(i2 - i1);                                          // :257
    if (i1 != i1) fl |= SK_FLAG_NAN1;
    if (i2 != i2) fl |= SK_FLAG_NAN2;
    if (d != d) { fl |= SK_FLAG_NAND; d = 0.0; }
    stage_i[j] = i2;
    stage_e[j] = d;
  }
  // block max of |I2-I1| (src/quadrature.jl:258) and NaN flags
  __shared__ double smax[8];
  __shared__ unsigned int sfl[8];
  d = sk_warp_max(d);
  fl = __reduce_or_sync(0xffffffffu, fl);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { smax[wid] = d; sfl[wid] = fl; }
  __syncthreads();
  if (wid == 0) {
    d = lane < (blockDim.x >> 5) ? smax[lane] : 0.0;
    fl = lane < (blockDim.x >> 5) ? sfl[lane] : 0u;
    d = sk_warp_max(d);
    fl = __reduce_or_sync(0xffffffffu, fl);
    if (lane == 0) {
      atomicMax(&red->maxbits, (unsigned long long)__double_as_longlong(d));
      if (fl) atomicOr(&red->flags, fl);
    }
  }
}

template <int W>
__global__ void __launch_bounds__(256)
k_interp_cplx(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkGeom G, const double *__restrict__ x,
              long long n, const sk_cplx *__restrict__ grid, sk_cplx *__restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  sk_cplx o;
  sk_interp_point<W, 1>(P, G, x[j], grid, &o.x, &o.y);
  out[j] = o;
}

// ---- K7: direct summation, int[j] = sum_k buf[k] cispi(2 no[k] x[j]) ---------------------------------
// grid (n_targets, 2 rules), one block per (target, rule); deterministic tree reduction.
__global__ void __launch_bounds__(256)
k_direct(const double *__restrict__ no1, const double *__restrict__ buf1, long long M1, const double *__restrict__ no2,
         const double *__restrict__ buf2, long long M2, const double *__restrict__ xs, sk_cplx *__restrict__ sums) {
  const int r = blockIdx.y;
  const double *no = r ? no2 : no1;
  const double *buf = r ? buf2 : buf1;
  const long long M = r ? M2 : M1;
  const double xj = xs[blockIdx.x];
  double ar = 0.0, ai = 0.0;
  for (long long k = threadIdx.x; k < M; k += blockDim.x) {
    double s, c;
    sincospi(sk_mul(sk_mul(2.0, no[k]), xj), &s, &c);       // cispi(2*no[k]*xj), src/quadrature.jl:121
    ar = sk_fma(buf[k], c, ar);
    ai = sk_fma(buf[k], s, ai);
  }
  __shared__ double sr[256], si[256];
  sr[threadIdx.x] = ar;
  si[threadIdx.x] = ai;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sk_cplx o; o.x = sr[0]; o.y = si[0]; sums[(long long)blockIdx.x * 2 + r] = o; }
}

// epilogue of the direct branch: same staging as k_interp_session (single block, n is tiny)
__global__ void k_direct_finish(const sk_cplx *__restrict__ sums, long long n, double cmul, int kernel_sin,
                                double *__restrict__ stage_i, double *__restrict__ stage_e,
                                SkReduceOut *__restrict__ red) {
  for (long long j = threadIdx.x; j < n; j += blockDim.x) {
    const sk_cplx f1 = sums[2 * j], f2 = sums[2 * j + 1];
    const double i1 = (kernel_sin ? f1.y : f1.x) * cmul;
    const double i2 = (kernel_sin ? f2.y : f2.x) * cmul;
    double d = fabs(i2 - i1);
    unsigned int fl = 0;
    if (i1 != i1) fl |= SK_FLAG_NAN1;
    if (i2 != i2) fl |= SK_FLAG_NAN2;
    if (d != d) { fl |= SK_FLAG_NAND; d = 0.0; }
    stage_i[j] = i2;
    stage_e[j] = d;
    atomicMax(&red->maxbits, (unsigned long long)__double_as_longlong(d));
    if (fl) atomicOr(&red->flags, fl);
  }
}

// ---- K5 ---------------------------------------------------------------------------------------------
__global__ void k_accept(double *__restrict__ I, double *__restrict__ err, const double *__restrict__ stage_i,
                         const double *__restrict__ stage_e, long long n, int first) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (first) {            // I = 0 + I2 exactly (I, err start at zero: src/quadrature.jl:174-175)
    I[j] = stage_i[j];
    err[j] = stage_e[j];
  } else {
    I[j] += stage_i[j];   // :261
    err[j] += stage_e[j]; // :262
  }
}

__global__ void k_commit(double *__restrict__ ks, double *__restrict__ errs, const double *__restrict__ I,
                         const double *__restrict__ err, long long n) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  ks[j] += I[j];        // src/adaptive.jl:163
  errs[j] += err[j];    // :164
}

// ---- K6 ---------------------------------------------------------------------------------------------
// The reference walks ix = hi, hi-1, ... while converged (src/adaptive.jl:185-197).  Equivalent:
// new_hi = the largest index whose predicate is false.  lo0 is the 0-based global index of element 0.
__global__ void __launch_bounds__(256)
k_scan(const double *__restrict__ xs, const double *__restrict__ I, long long n, long long lo0, double trunc_a,
       double trunc_num, double xpow, double tau, int criteria, SkReduceOut *__restrict__ red) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long bad = -1;
  if (j < n) {
    const double te = sk_trunc_err(trunc_a, trunc_num, xpow, xs[j], criteria == 0);
    if (!sk_converged(te, I[j], tau, criteria)) bad = lo0 + j;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, bad, o);
    bad = other > bad ? other : bad;
  }
  __shared__ long long sb[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sb[wid] = bad;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) bad = sb[w] > bad ? sb[w] : bad;
    if (bad >= 0) atomicMax(&red->max_unconv, bad);
  }
}

// errs[ix] += 2*trunc_err for the converged tail (src/adaptive.jl:194)
__global__ void k_scan_add(const double *__restrict__ xs, double *__restrict__ errs, long long n, double trunc_a,
                           double trunc_num, double xpow, int criteria) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double te = sk_trunc_err(trunc_a, trunc_num, xpow, xs[j], criteria == 0);
  errs[j] += 2 * te;
}

// ---- K8 ---------------------------------------------------------------------------------------------
// keys: bit patterns of the (non-negative) doubles, which order like unsigned integers; -0.0 -> +0.0.
__global__ void k_make_keys(const double *__restrict__ xs, long long n, unsigned long long *__restrict__ keys,
                            unsigned int *__restrict__ idx, unsigned int *__restrict__ bad) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double x = xs[j];
  if (!(x >= 0.0) || isinf(x)) { atomicOr(bad, 1u); x = 0.0; }
  if (x == 0.0) x = 0.0;
  keys[j] = (unsigned long long)__double_as_longlong(x);
  idx[j] = (unsigned int)j;
}

__global__ void k_flag_heads(const unsigned long long *__restrict__ keys, long long n, unsigned int *__restrict__ head) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  head[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

// uid = inclusive-scan(head) - 1; unique value table and inverse map (original position -> unique id)
__global__ void k_scatter_unique(const unsigned long long *__restrict__ keys, const unsigned int *__restrict__ idx,
                                 const unsigned int *__restrict__ head, const unsigned int *__restrict__ uid_incl,
                                 long long n, double *__restrict__ uxs, unsigned int *__restrict__ inv) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned int u = uid_incl[j] - 1u;
  if (head[j]) uxs[u] = __longlong_as_double((long long)keys[j]);
  inv[idx[j]] = u;
}

__global__ void k_gather(const unsigned int *__restrict__ inv, const double *__restrict__ ks,
                         const double *__restrict__ errs, long long n, double *__restrict__ out_v,
                         double *__restrict__ out_e) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned int u = inv[j];
  out_v[j] = ks[u];
  if (out_e) out_e[j] = errs[u];
}

// number of sorted values <= r (== the largest 1-based index with xs[idx] <= r)
__global__ void k_upper_bound(const double *__restrict__ xs, long long n, double r, long long *__restrict__ out) {
  long long a = 0, b = n;
  while (a < b) {
    const long long mid = (a + b) >> 1;
    if (xs[mid] <= r) a = mid + 1; else b = mid;
  }
  *out = a;
}

__global__ void k_fill(double *__restrict__ a, long long n, double v) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) a[j] = v;
}

// ---- FP64 pipe micro-benchmark (roofline denominator) -------------------------------------------------
// 8 independent dependent-FMA chains per thread; 2 flops per DFMA.
__global__ void __launch_bounds__(256) k_dfma_peak(double *__restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}
