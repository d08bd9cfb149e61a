// sk_kernels.cuh -- sm_100a kernels of the K(r) path.  Each kernel is a launch wrapper around the
// per-element functions of sk_math.h plus the block-level cooperation (shared memory, reductions).
//
//   K1  k_gen_sources       updatequadbufs!                       src/quadrature.jl:49-95
//   K2  k_prep_sources      source positions / pre-phase          (inside nufft1d3, src/utils.jl:10)
//       k_spread_modes      ES spread (deterministic gather, 8 lanes per grid point) + mode
//                           deconvolution + zero-pad
//   K3  cuFFT Z2Z           (sk_api.cu)
//   K4  k_interp_cells      targets -> cell polynomials in shared memory -> Horner; target-side
//                           deconvolution, post-phase, Re/Im select, *c, |I2-I1|, block max
//                                                                 src/quadrature.jl:130-136, :250-258
//       k_interp_session    per-target ES taps (reference-style evaluation; A/B and sparse fallback)
//       k_interp_cplx       complex output (Level-0 sk_nufft1d3)
//   K5  k_accept_add        I += I2; err += |I2-I1|               src/quadrature.jl:260-262
//   K6  k_commit_scan       ks += I; errs += err fused with the convergence predicate
//                                                                 src/adaptive.jl:163-164, :183-199
//       k_scan_add          errs += 2 trunc_err for the converged tail   src/adaptive.jl:194
//   K7  k_direct            direct Fourier summation              src/quadrature.jl:113-128
//   K8  sk_k8.cuh (unique / sort / inverse map without a radix sort), k_gather
//       k_make_keys, k_flag_heads, k_scatter_unique, k_target_summary: the general sort for inputs the bin
//       scheme cannot take                                    src/adaptive.jl:99-120
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "sk_math.h"
#include "sk_k8.cuh"

#define SK_FLAG_NAN1 1u
#define SK_FLAG_NAN2 2u
#define SK_FLAG_NAND 4u
#define SK_FLAG_SKIPPED 8u    // a chained launch whose guard did not hold: nothing was touched

struct SkReduceOut {            // device scalars written by the reductions
  unsigned long long maxbits;   // bit pattern of max |I2-I1| (non-negative doubles order like integers)
  unsigned int flags;
  unsigned int _pad;
  long long max_unconv;         // highest non-converged 0-based index, or lo-1
  unsigned long long rbits;     // bit pattern of the distance at that index (0 if none)
};


// Speculative commit (the common case: the first sub-interval of a panel is accepted and is the whole
// panel).  The interpolation kernel then also does  ks += I2, errs += |I2-I1|  in place (keeping the old
// (ks, errs) pair in `backup` so that a rejected sub-interval can be rolled back bit for bit) and
// evaluates the convergence predicate of src/adaptive.jl:185-197 on panel_ks = I2, so that no separate
// pass over the targets is needed for src/quadrature.jl:261-262, src/adaptive.jl:163-164 and :183-198.
struct SkPeerOut;
struct SkSpec {
  int on;
  int criteria;
  long long lo0;          // global 0-based index of element 0
  double trunc_a, trunc_num, xpow, tau;
  // The truncation half of the predicate, trunc_err(x) < tau, is monotone in the distance x (for trunc_num > 0 the
  // bound min(trunc_a, trunc_num / (2 pi x)) does not increase with x, and IEEE rounding preserves that): the host
  // finds the smallest double xstar with trunc_err(xstar) < tau by bisection on the SAME function (sk_trunc_err), so
  // "x >= xstar" is the same decision for every double x -- one compare per target instead of a division.
  int use_xstar;
  // fresh = 1: (ks, errs) are still zero everywhere in the panel (the first panel of a run covers every positive
  // distance): the old pair is not read (it is (0, 0)), no roll-back copy is written (rolling back = zeroing), and the
  // run needs no memset of the table at all.  Same additions 0 + I2, 0 + |I2-I1|, so the same bits.
  int fresh;
  double xstar;
  sk_cplx *res;           // (ks, errs), pre-offset to element 0
  sk_cplx *backup;        // old (ks, errs), pre-offset
  // Chained launch (sk_subinterval_chain): the NEXT panel's first sub-interval is enqueued behind this panel's before the
  // host has seen this panel's outcome.  It runs only if that outcome is the one the host predicted -- the previous
  // sub-interval accepted (max |I2-I1| below the accept threshold, no NaN: src/quadrature.jl:260) and the scan
  // converged nothing (highest unconverged index = top of the panel, so r_hi and with it the next panel's ends are
  // unchanged: src/adaptive.jl:152) -- and otherwise returns at once with SK_FLAG_SKIPPED, having touched nothing.
  const SkReduceOut *guard;          // nullptr: no guard
  unsigned long long guard_maxbits;  // run only if guard->maxbits < guard_maxbits ...
  long long guard_top;               // ... and guard->max_unconv == guard_top and guard->flags == 0
  // target-sharded run (peer mailboxes): the same guard on the GLOBAL scalars the previous sub-interval's exchange kernel
  // left in device memory -- max |I2-I1| over all ranks below the threshold, no NaN / error / void on any rank, and the
  // global stopping distance still the global r_hi (nothing converged anywhere).  Every rank sees the same words, so
  // every chained launch of a step takes the same decision.
  const SkPeerOut *gguard;           // nullptr: the local form above
  unsigned long long gguard_rbits;   // bit pattern of the global r_hi
  // First panel enqueued behind the SORT (sk_first_panel_early): the number of unique distances and the buffer the
  // unique table ended up in are not known to the host yet -- the kernel takes them from the sort's device-side summary
  // (lo = 0 / 1, the r = 0 row, is known from the first pass).  It skips itself if the sort did not deliver (bad input,
  // bin overflow -> general sort on the host, duplicates dropped -> the host still compacts the table) or if the active
  // set is too small for the NUFFT branch.
  const SkTargetSummary *dyn;        // nullptr: n is the kernel argument
  long long dyn_lo, dyn_min_n;
};

__device__ __forceinline__ bool sk_chain_guard_holds(const SkSpec &spec);

__device__ __forceinline__ double sk_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block max of |I2-I1| (src/quadrature.jl:258) and NaN flags -> device scalars
__device__ __forceinline__ void sk_block_reduce_maxflags(double d, unsigned int fl, SkReduceOut *red) {
  __shared__ double smax[32];
  __shared__ unsigned int sfl[32];
  d = sk_warp_max(d);
  fl = __reduce_or_sync(0xffffffffu, fl);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { smax[wid] = d; sfl[wid] = fl; }
  __syncthreads();
  if (wid == 0) {
    d = lane < nw ? smax[lane] : 0.0;
    fl = lane < nw ? sfl[lane] : 0u;
    d = sk_warp_max(d);
    fl = __reduce_or_sync(0xffffffffu, fl);
    if (lane == 0) {
      atomicMax(&red->maxbits, (unsigned long long)__double_as_longlong(d));
      if (fl) atomicOr(&red->flags, fl);
    }
  }
}

// staging of one target: I2 and |I2-I1| (src/quadrature.jl:250-257) as one 16-byte store
__device__ __forceinline__ void sk_stage(double f1, double f2, double cmul, sk_cplx *dst, double &d, unsigned int &fl) {
  const double i1 = sk_mul(f1, cmul), i2 = sk_mul(f2, cmul);     // explicit roundings: identical in every code path
  double dd = fabs(sk_add(i2, -i1));
  sk_cplx o;
  o.x = i2;
  o.y = dd;
  *dst = o;
  if (dd != dd) {                                                 // NaN in I1 or I2 (or inf - inf): rare path
    fl |= SK_FLAG_NAND;
    if (i1 != i1) fl |= SK_FLAG_NAN1;
    if (i2 != i2) fl |= SK_FLAG_NAN2;
    dd = 0.0;
  }
  d = fmax(d, dd);
}

// per-thread state of the reductions of the interpolation kernels
struct SkAcc {
  double d;                 // max |I2-I1|
  unsigned int fl;          // NaN flags
  long long bad;            // highest non-converged global index seen (speculative scan), -1 if none
  unsigned long long rb;    // bit pattern of its distance
};

// the arithmetic of one target's result, without the stores: (SPEC == false) out = (I2, |I2-I1|) for the staging
// buffer; (SPEC == true) out = new (ks, errs) for `res` (the old pair goes to `backup`), plus the convergence
// predicate of src/adaptive.jl:185-197 on panel_ks = I2
template <bool SPEC>
__device__ __forceinline__ sk_cplx sk_emit_value(const SkSpec &spec, double f1, double f2, double cmul, double x, long long j,
                                                 SkAcc &acc, const sk_cplx old) {
  const double i1 = sk_mul(f1, cmul), i2 = sk_mul(f2, cmul);     // explicit roundings: identical in every code path
  double dd = fabs(sk_add(i2, -i1));
  sk_cplx out;
  if (!SPEC) {
    out.x = i2;
    out.y = dd;
  } else {
    out.x = sk_add(old.x, i2);   // ks += I with I = 0 + I2   (src/quadrature.jl:261, src/adaptive.jl:163)
    out.y = sk_add(old.y, dd);   // errs += err with err = 0 + |I2-I1|
  }
  if (dd != dd) {                                                 // NaN in I1 or I2 (or inf - inf): rare path
    acc.fl |= SK_FLAG_NAND;
    if (i1 != i1) acc.fl |= SK_FLAG_NAN1;
    if (i2 != i2) acc.fl |= SK_FLAG_NAN2;
    dd = 0.0;
  }
  acc.d = fmax(acc.d, dd);
  if (SPEC) {
    bool conv;
    if (spec.use_xstar) {
      conv = (x >= spec.xstar) && ((spec.criteria == 1) || (fabs(i2) < spec.tau));
    } else {
      const double te = sk_trunc_err(spec.trunc_a, spec.trunc_num, spec.xpow, x, spec.criteria == 0);
      conv = sk_converged(te, i2, spec.tau, spec.criteria);
    }
    if (!conv) {
      const long long g = spec.lo0 + j;
      if (g > acc.bad) { acc.bad = g; acc.rb = (unsigned long long)__double_as_longlong(x); }
    }
  }
  return out;
}

// stage (SPEC == false) or commit speculatively (SPEC == true) the target with local index j, distance x
template <bool SPEC>
__device__ __forceinline__ void sk_emit(const SkSpec &spec, double f1, double f2, double cmul, double x, long long j,
                                        sk_cplx *stage, SkAcc &acc, const sk_cplx old) {
  const sk_cplx out = sk_emit_value<SPEC>(spec, f1, f2, cmul, x, j, acc, old);
  if (!SPEC) {
    stage[j] = out;
  } else {
    if (!spec.fresh) spec.backup[j] = old;       // `old` = spec.res[j], loaded early by the caller to hide the latency
    spec.res[j] = out;
  }
}

// asynchronous global -> shared copies (LDGSTS): the data bypasses the register file and the issuing warp does not wait
__device__ __forceinline__ void sk_cp_async8(void *smem_dst, const void *gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void sk_cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void sk_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void sk_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 32-byte global accesses (two (re, im) / (ks, errs) pairs per instruction): a thread that owns adjacent targets moves
// whole sectors
__device__ __forceinline__ void sk_ld256(const void *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void sk_st256(void *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// block reduction of SkAcc -> device scalars
__device__ __forceinline__ void sk_block_reduce_acc(SkAcc a, int spec_on, SkReduceOut *red) {
  __shared__ double smax[32];
  __shared__ unsigned int sfl[32];
  __shared__ long long sbad[32];
  __shared__ unsigned long long srb[32];
  a.d = sk_warp_max(a.d);
  a.fl = __reduce_or_sync(0xffffffffu, a.fl);
  if (spec_on) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long ob = __shfl_xor_sync(0xffffffffu, a.bad, o);
      const unsigned long long orb = __shfl_xor_sync(0xffffffffu, a.rb, o);
      if (ob > a.bad) { a.bad = ob; a.rb = orb; }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { smax[wid] = a.d; sfl[wid] = a.fl; sbad[wid] = a.bad; srb[wid] = a.rb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w) {
      a.d = fmax(a.d, smax[w]);
      a.fl |= sfl[w];
      if (sbad[w] > a.bad) { a.bad = sbad[w]; a.rb = srb[w]; }
    }
    atomicMax(&red->maxbits, (unsigned long long)__double_as_longlong(a.d));
    if (a.fl) atomicOr(&red->flags, a.fl);
    if (spec_on && a.bad >= 0) {             // sorted unique distances: max index <=> max distance
      atomicMax(&red->max_unconv, a.bad);
      atomicMax(&red->rbits, a.rb);
    }
  }
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

// Spread + mode deconvolution.  Every spread-grid cell sums the sources whose kernel support covers it
// in source order -- no atomics, bitwise reproducible.  (The zero-padded part of the FFT input is
// cleared by a memset before this kernel.)  One block = SK_SPREAD_CELLS consecutive cells:
//   0. two warps find the block's source range [s0, s1) with 32-ary searches (sources are sorted);
//   1. in chunks of SK_SPREAD_CHUNK sources: one thread per source evaluates its w tap weights with the
//      shared tap polynomials (Horner, ~8 FMAs per tap -- an exp/sqrt evaluation per (source, cell) pair
//      would cost ~100 instructions) and parks them in shared memory together with its first cell;
//   2. SK_SPREAD_LANES lanes per cell pick up the weights that land on their cell (binary search on the
//      first-cell array, then a short walk) and accumulate weight * strength; the lanes' partial sums are
//      combined in a fixed butterfly order.
// Gauss nodes cluster at the sub-panel ends (hundreds of sources within one kernel width there): the
// chunk loop handles any number of sources per block.
#define SK_SPREAD_LANES 4
#define SK_SPREAD_CELLS (256 / SK_SPREAD_LANES)
#define SK_SPREAD_CHUNK 256
struct SkSpreadSrc {
  const double *pos_hi[2];
  const double *pos_lo[2];
  const sk_cplx *cs[2];
  long long M[2];
};
__device__ __forceinline__ long long sk_warp_lower_bound(const double *__restrict__ ph, long long M, double edge,
                                                         bool strict_greater) {
  // first index with ph[i] >= edge (or > edge when strict_greater): 32-ary search by one full warp
  long long lo = 0, hi = M;
  const int lane = threadIdx.x & 31;
  while (hi > lo) {
    const long long step = (hi - lo + 31) / 32;
    const long long probe = lo + (long long)lane * step;
    bool below = false;
    if (probe < hi) {
      const double v = ph[probe];
      below = strict_greater ? (v <= edge) : (v < edge);
    }
    const int cnt = __popc(__ballot_sync(0xffffffffu, below));            // sorted: the first cnt probes are below
    if (cnt == 0) { hi = lo; break; }
    const long long nlo = lo + (long long)(cnt - 1) * step + 1;
    const long long nhi = lo + (long long)cnt * step;
    lo = nlo;
    hi = nhi < hi ? nhi : hi;
  }
  return lo;
}

template <int W>
__global__ void __launch_bounds__(256)
k_spread_modes(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkGeom G,
               const __grid_constant__ SkSpreadSrc src, int nrule, sk_cplx *__restrict__ fft_io) {
  const int r = blockIdx.y;
  const double *__restrict__ ph = src.pos_hi[r];
  const double *__restrict__ pl = src.pos_lo[r];
  const sk_cplx *__restrict__ cs = src.cs[r];
  const long long M = src.M[r];
  const double half = 0.5 * W;
  const long long l_blk = (long long)blockIdx.x * SK_SPREAD_CELLS;       // first cell of the block
  const long long n_blk = l_blk - G.nf / 2;                              // its signed mode index
  __shared__ long long s_range[2];
  __shared__ int s_l0[SK_SPREAD_CHUNK];                                  // first cell of the source, relative to n_blk
  __shared__ sk_cplx s_cs[SK_SPREAD_CHUNK];
  __shared__ double s_tap[SK_SPREAD_CHUNK][W + 1];                       // +1: spreads the rows over the banks
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
  const int cell = threadIdx.x / SK_SPREAD_LANES;                        // 0 .. SK_SPREAD_CELLS-1
  const long long l = l_blk + cell;
  const bool live = l < G.nf;
  double ar = 0.0, ai = 0.0;
  for (long long c0 = s0; c0 < s1; c0 += SK_SPREAD_CHUNK) {
    const int ns = (int)((s1 - c0) < (long long)SK_SPREAD_CHUNK ? (s1 - c0) : (long long)SK_SPREAD_CHUNK);
    // 1. tap weights of the chunk's sources
    if ((int)threadIdx.x < ns) {
      const long long k = c0 + threadIdx.x;
      const double p_hi = ph[k], p_lo = pl[k];
      const double c_first = ceil(p_hi - half);                          // first cell (mode index) under the kernel
      const double x0 = (c_first - p_hi) - p_lo;                         // in [-W/2, -W/2 + 1)
      double taps[W];
      sk_es_taps<W>(P, 2.0 * (x0 + (half - 0.5)), taps);
#pragma unroll
      for (int i = 0; i < W; ++i) s_tap[threadIdx.x][i] = taps[i];
      s_l0[threadIdx.x] = (int)((long long)c_first - n_blk);
      s_cs[threadIdx.x] = cs[k];
    }
    __syncthreads();
    // 2. cells gather: sources with first cell in [cell - W + 1, cell] (s_l0 is non-decreasing)
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
        ar = sk_fma(wgt, c.x, ar);
        ai = sk_fma(wgt, c.y, ai);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = SK_SPREAD_LANES / 2; o > 0; o >>= 1) {                   // fixed-order butterfly inside the lane group
    ar += __shfl_xor_sync(0xffffffffu, ar, o);
    ai += __shfl_xor_sync(0xffffffffu, ai, o);
  }
  if (sub == 0 && live) {
    const long long n = l - G.nf / 2;                                    // signed mode index
    double q = sk_deconv(P, G.t_cell * fabs((double)n));
    if (n & 1) q = -q;                                                   // shifts the FFT output by nf2/2
    sk_cplx o;
    o.x = ar * q;
    o.y = ai * q;
    const long long j = n >= 0 ? n : n + G.nf2;                          // FFT-input index of mode n
    fft_io[j * nrule + r] = o;
  }
}

// ---- K4 (reference-style per-target taps) -------------------------------------------------------------
// grid layout [nf2][2] (m-rule, 2m-rule); stage[j] = (I2, |I2-I1|)
template <int W, bool SPEC>
__global__ void __launch_bounds__(256)
k_interp_session(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkGeom G, const double *__restrict__ xs,
                 long long n, const sk_cplx *__restrict__ grid, double cmul, int kernel_sin,
                 sk_cplx *__restrict__ stage, const __grid_constant__ SkSpec spec, SkReduceOut *__restrict__ red) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  SkAcc acc = {0.0, 0u, -1, 0ull};
  if (j < n) {
    double fre[2], fim[2];
    const double x = xs[j];
    sk_cplx old;
    old.x = old.y = 0.0;
    if (SPEC && !spec.fresh) old = spec.res[j];
    sk_interp_point<W, 2>(P, G, x, grid, fre, fim);
    // kernel == :cos -> real part, :sin -> imaginary part (src/quadrature.jl:130-136); then *c (:250-251)
    sk_emit<SPEC>(spec, kernel_sin ? fim[0] : fre[0], kernel_sin ? fim[1] : fre[1], cmul, x, j, stage, acc, old);
  }
  sk_block_reduce_acc(acc, SPEC ? 1 : 0, red);
}

// one target from its cell's folded coefficients: Horner, post-phase, Re/Im select (explicit roundings so
// that the block path and the warp path of k_interp_cells give bit-identical values)
// post-phase and Re/Im select of one target from its four Horner sums a = (re1, im1, re2, im2)
__device__ __forceinline__ void sk_cell_finish(const double *a, const sk_cplx *tab, const SkGeom &G, double r, int kernel_sin,
                                               double *f1, double *f2) {
  double sn, cs;
  sk_sincos2pi(tab, sk_frac_prod(G.wc, r, 0.0), &sn, &cs);            // post-phase exp(2 pi i wc r)
  if (kernel_sin) {                                                   // Im((a0 + i a1)(cs + i sn))
    *f1 = sk_fma(a[0], sn, sk_mul(a[1], cs));
    *f2 = sk_fma(a[2], sn, sk_mul(a[3], cs));
  } else {                                                            // Re
    *f1 = sk_fma(a[0], cs, -sk_mul(a[1], sn));
    *f2 = sk_fma(a[2], cs, -sk_mul(a[3], sn));
  }
}
__device__ __forceinline__ void sk_cell_eval(const double *coef, const sk_cplx *tab, const SkGeom &G, double r, double s,
                                             int kernel_sin, double *f1, double *f2) {
  double a[4];
  sk_cell_horner<4>(coef, s, a);
  sk_cell_finish(a, tab, G, r, kernel_sin, f1, f2);
}

// ---- K4 (cell polynomials) -----------------------------------------------------------------------------
// One block = SK_TPB consecutive sorted targets.  Sorted targets touch a contiguous run of fine-grid
// cells; when the run is short enough (ncell <= cmax) the block
//   A. stages the grid window [l_first, l_last + W) of both rules in shared memory (coalesced),
//   B. turns it into SK_NC polynomial coefficients x 4 components (re/im x two rules) per cell,
//   C. evaluates every target by Horner from shared memory (broadcast reads inside a cell).
// Otherwise (sparse targets) it falls back to per-target taps straight from L2.
#define SK_TPT 8
#define SK_TPB (256 * SK_TPT)
// doubles between the coefficient blocks of consecutive cells: 2 more than SK_NC * 4 = 64, so that the blocks of two
// neighbouring cells (a warp's 128 adjacent targets often span two) do not start in the same shared-memory bank
#define SK_CELL_STRIDE (SK_NC * 4 + 2)
template <int W, bool SPEC, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_interp_cells(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkGeom G, const double *__restrict__ xs_arg,
               long long n_arg, const sk_cplx *__restrict__ grid, double cmul, int kernel_sin, int cmax, int tpt,
               sk_cplx *__restrict__ stage, const __grid_constant__ SkSpec spec, SkReduceOut *__restrict__ red) {
  // tpt (4 or 8) targets per thread: 2048-target blocks amortise the per-cell work best; smaller launches use
  // 1024-target blocks so that the grid still fills the 148 SMs several times over
  if (SPEC && !sk_chain_guard_holds(spec)) {           // chained launch whose prediction failed: touch nothing
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&red->flags, SK_FLAG_SKIPPED);
    return;
  }
  const double *xs = xs_arg;
  long long n = n_arg;
  if (SPEC && spec.dyn != nullptr) {                   // enqueued behind the sort: its summary says what there is
    const long long nu = __ldcg(&spec.dyn->n_unique);
    const unsigned int bad = __ldcg(&spec.dyn->bad), over = __ldcg(&spec.dyn->overflow), fixed = __ldcg(&spec.dyn->fixed);
    n = nu - spec.dyn_lo;
    // (fixed: some bin dropped duplicates -- the host still has to compact the table: not this time)
    if (bad || over || fixed || n <= spec.dyn_min_n) {
      if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&red->flags, SK_FLAG_SKIPPED);
      return;
    }
    if ((long long)blockIdx.x * 256 * tpt >= n) return;          // the grid was sized for the number of INPUT distances
  }
  extern __shared__ __align__(16) double smem[];
  double *sE = smem;                                   // [W/2][SK_NC/2]
  double *sO = sE + (W / 2) * (SK_NC / 2);
  double *sWin = sO + (W / 2) * (SK_NC / 2);            // [(cmax + W)][4]
  double *sQ = sWin + (size_t)(cmax + W) * 4;           // [cmax][4]  deconvolution factor at the 4 Chebyshev nodes
  double *sCoef = sQ + (size_t)cmax * 4;                // [cmax][SK_CELL_STRIDE]: (q, comp) at q * 4 + comp
  double *sR = sCoef + (size_t)cmax * SK_CELL_STRIDE;   // [256 * tpt] the block's distances, staged asynchronously
  __shared__ sk_cplx sTab[65];                          // (cos, sin)(2 pi k / 64) for the post-phase
  const int tpb = 256 * tpt;
  const long long j0 = (long long)blockIdx.x * tpb;
  const int cnt = (int)((n - j0) < (long long)tpb ? (n - j0) : (long long)tpb);
  // The block's distances go to shared memory with asynchronous copies (cp.async -> LDGSTS) issued before anything
  // else: they land while phases A and B build the cell polynomials, so phase C starts from shared memory instead of
  // waiting for its first global loads (27 % of the stall samples of the previous version sat on those loads).
  {
    const double *src = xs + j0;
    if ((reinterpret_cast<size_t>(src) & 15) == 0) {
      for (int t = 2 * threadIdx.x; t + 1 < cnt; t += 512) sk_cp_async16(sR + t, src + t);
      if ((cnt & 1) && threadIdx.x == 0) sk_cp_async8(sR + cnt - 1, src + cnt - 1);
    } else {
      for (int t = threadIdx.x; t < cnt; t += 256) sk_cp_async8(sR + t, src + t);
    }
    sk_cp_async_commit();
  }
  const long long l_first = sk_target_coord<W>(G, xs[j0]).l0;
  const long long l_last = sk_target_coord<W>(G, xs[j0 + cnt - 1]).l0;
  const long long ncell_ll = l_last - l_first + 1;
  SkAcc acc = {0.0, 0u, -1, 0ull};
  for (int t = threadIdx.x; t < (W / 2) * (SK_NC / 2); t += blockDim.x) {
    sE[t] = P.E[t / (SK_NC / 2)][t % (SK_NC / 2)];
    sO[t] = P.O[t / (SK_NC / 2)][t % (SK_NC / 2)];
  }
  if (threadIdx.x >= 128 && threadIdx.x < 128 + 65) sk_sincos2pi_table_fill(sTab, threadIdx.x - 128);
  if (ncell_ll >= 1 && ncell_ll <= (long long)cmax) {
    const int ncell = (int)ncell_ll;
    // A: window of (ncell + W - 1) grid points x 2 rules, 32 bytes per point
    const double4 *gsrc = reinterpret_cast<const double4 *>(grid + (size_t)l_first * 2);
    double4 *wdst = reinterpret_cast<double4 *>(sWin);
    for (int t = threadIdx.x; t < ncell + W - 1; t += blockDim.x) wdst[t] = gsrc[t];
    // the target-side deconvolution factor at the 4 Chebyshev nodes of every cell (independent of A)
    for (int t = threadIdx.x; t < ncell * 4; t += blockDim.x) {
      const int k = t & 3, cell = t >> 2;
      const double node = (k == 0) ? 0.9238795325112867 : (k == 1) ? 0.3826834323650898
                        : (k == 2) ? -0.3826834323650898 : -0.9238795325112867;
      const double ymid = (double)(l_first + cell - G.nf2 / 2) + (0.5 * W - 0.5);
      sQ[t] = sk_deconv(P, G.t_cell * fabs(ymid - 0.5 * node));
    }
    __syncthreads();
    // B: coefficient (cell, q, comp) = W/2 FMAs
    for (int t = threadIdx.x; t < ncell * SK_NC * 4; t += blockDim.x) {
      const int comp = t & 3, q = (t >> 2) & (SK_NC - 1), cell = t / (SK_NC * 4);
      sCoef[(size_t)cell * SK_CELL_STRIDE + (t & (SK_NC * 4 - 1))] = sk_cell_coef<W>(sE, sO, sWin + cell * 4 + comp, 4, q);
    }
    __syncthreads();
    // B': fold the deconvolution (a cubic per cell) into the coefficients, one column per thread
    for (int t = threadIdx.x; t < ncell * 4; t += blockDim.x) {
      const int comp = t & 3, cell = t >> 2;
      double a[4];
      sk_cheb4_to_monomial(sQ + cell * 4, a);
      sk_cell_fold(sCoef + (size_t)cell * SK_CELL_STRIDE + comp, 4, a);
    }
    sk_cp_async_wait_all();                                  // this thread's share of sR has landed ...
    __syncthreads();                                         // ... and so has everybody else's
    // C: Horner.  A thread takes FOUR ADJACENT targets at a time: sorted targets next to each other nearly always share
    // their cell (~150 targets per cell at 1e7), so one pair of 16-byte shared-memory loads per coefficient feeds 16
    // FMAs (4 targets x re/im x two rules) instead of 4, and every thread reads and writes 32 / 64 contiguous bytes of
    // xs / (ks, errs).  The few quads that straddle a cell boundary (~2 %) would make half of all warps run a second,
    // divergent pass; their targets are parked in a shared-memory list instead (the window / deconvolution scratch is
    // free by now) and evaluated one per thread after the loop.  Each target sees exactly the operations of
    // sk_cell_eval, in the same order, on every route: values do not depend on how targets are grouped (block path,
    // warp path and the parked targets agree bit for bit).
    unsigned short *sList = reinterpret_cast<unsigned short *>(sWin);
    const int list_cap = (int)((((size_t)(cmax + W) * 4 + (size_t)cmax * 4) * sizeof(double)) / sizeof(unsigned short));
    __shared__ int sListN;
    if (threadIdx.x == 0) sListN = 0;
    __syncthreads();                                         // (also: every thread is done with sWin / sQ)
    const bool aligned32 = ((reinterpret_cast<size_t>(stage + j0) |
                             (SPEC ? (reinterpret_cast<size_t>(spec.res + j0) | reinterpret_cast<size_t>(spec.backup + j0)) : 0)) & 31) == 0;
#pragma unroll 1
    for (int uo = 0; uo < tpt / 4; ++uo) {
      const int t0 = (threadIdx.x + uo * 256) * 4;           // first local target of the quad
      if (t0 >= cnt) continue;
      const int nq = cnt - t0 < 4 ? cnt - t0 : 4;
      double rr[4];
      sk_cplx old[4];
      const bool wide = aligned32 && nq == 4;
      if (wide) {
        {
          const double2 r01 = *reinterpret_cast<const double2 *>(sR + t0), r23 = *reinterpret_cast<const double2 *>(sR + t0 + 2);
          rr[0] = r01.x; rr[1] = r01.y; rr[2] = r23.x; rr[3] = r23.y;
        }
        if (SPEC && !spec.fresh) {
          sk_ld256(spec.res + j0 + t0, old[0].x, old[0].y, old[1].x, old[1].y);
          sk_ld256(spec.res + j0 + t0 + 2, old[2].x, old[2].y, old[3].x, old[3].y);
        } else {
#pragma unroll
          for (int ui = 0; ui < 4; ++ui) old[ui].x = old[ui].y = 0.0;
        }
      } else {
#pragma unroll
        for (int ui = 0; ui < 4; ++ui) {
          rr[ui] = ui < nq ? sR[t0 + ui] : 0.0;
          old[ui].x = old[ui].y = 0.0;
          if (SPEC && !spec.fresh && ui < nq) old[ui] = spec.res[j0 + t0 + ui];
        }
      }
      int cell[4];
      double sv[4];
#pragma unroll
      for (int ui = 0; ui < 4; ++ui) {
        const SkTargetCoord tc = sk_target_coord<W>(G, rr[ui]);
        int cl = (int)(tc.l0 - l_first);
        cell[ui] = cl < 0 ? 0 : (cl >= ncell ? ncell - 1 : cl);
        sv[ui] = tc.s;
      }
      if (nq == 4 && cell[0] == cell[3]) {
        double a[4][4];
        const double *coef = sCoef + (size_t)cell[0] * SK_CELL_STRIDE;
        {
          const double2 c01 = *reinterpret_cast<const double2 *>(coef + (SK_NC - 1) * 4);
          const double2 c23 = *reinterpret_cast<const double2 *>(coef + (SK_NC - 1) * 4 + 2);
#pragma unroll
          for (int ui = 0; ui < 4; ++ui) { a[ui][0] = c01.x; a[ui][1] = c01.y; a[ui][2] = c23.x; a[ui][3] = c23.y; }
        }
#pragma unroll
        for (int q = SK_NC - 2; q >= 0; --q) {
          const double2 c01 = *reinterpret_cast<const double2 *>(coef + q * 4);
          const double2 c23 = *reinterpret_cast<const double2 *>(coef + q * 4 + 2);
#pragma unroll
          for (int ui = 0; ui < 4; ++ui) {
            a[ui][0] = sk_fma(a[ui][0], sv[ui], c01.x);
            a[ui][1] = sk_fma(a[ui][1], sv[ui], c01.y);
            a[ui][2] = sk_fma(a[ui][2], sv[ui], c23.x);
            a[ui][3] = sk_fma(a[ui][3], sv[ui], c23.y);
          }
        }
        sk_cplx out[4];
#pragma unroll
        for (int ui = 0; ui < 4; ++ui) {
          double f1, f2;
          sk_cell_finish(a[ui], sTab, G, rr[ui], kernel_sin, &f1, &f2);
          out[ui] = sk_emit_value<SPEC>(spec, f1, f2, cmul, rr[ui], j0 + t0 + ui, acc, old[ui]);
        }
        if (wide) {
          if (SPEC) {
            if (!spec.fresh) {
              sk_st256(spec.backup + j0 + t0, old[0].x, old[0].y, old[1].x, old[1].y);
              sk_st256(spec.backup + j0 + t0 + 2, old[2].x, old[2].y, old[3].x, old[3].y);
            }
            sk_st256(spec.res + j0 + t0, out[0].x, out[0].y, out[1].x, out[1].y);
            sk_st256(spec.res + j0 + t0 + 2, out[2].x, out[2].y, out[3].x, out[3].y);
          } else {
            sk_st256(stage + j0 + t0, out[0].x, out[0].y, out[1].x, out[1].y);
            sk_st256(stage + j0 + t0 + 2, out[2].x, out[2].y, out[3].x, out[3].y);
          }
        } else {
#pragma unroll
          for (int ui = 0; ui < 4; ++ui) {
            if (!SPEC) stage[j0 + t0 + ui] = out[ui];
            else { if (!spec.fresh) spec.backup[j0 + t0 + ui] = old[ui]; spec.res[j0 + t0 + ui] = out[ui]; }
          }
        }
      } else {
        const int at = atomicAdd(&sListN, nq);
        if (at + nq <= list_cap) {
          for (int ui = 0; ui < nq; ++ui) sList[at + ui] = (unsigned short)(t0 + ui);
        } else {                                             // list full (cannot happen at the densities of this path)
          for (int ui = 0; ui < nq; ++ui) {
            double f1, f2;
            sk_cell_eval(sCoef + (size_t)cell[ui] * SK_CELL_STRIDE, sTab, G, rr[ui], sv[ui], kernel_sin, &f1, &f2);
            sk_emit<SPEC>(spec, f1, f2, cmul, rr[ui], j0 + t0 + ui, stage, acc, old[ui]);
          }
        }
      }
    }
    __syncthreads();
    const int nlist = sListN < list_cap ? sListN : list_cap;
    for (int i = threadIdx.x; i < nlist; i += blockDim.x) {  // the parked targets, one per thread
      const int t = sList[i];
      const double r = sR[t];
      sk_cplx old;
      old.x = old.y = 0.0;
      if (SPEC && !spec.fresh) old = spec.res[j0 + t];
      const SkTargetCoord tc = sk_target_coord<W>(G, r);
      int cl = (int)(tc.l0 - l_first);
      cl = cl < 0 ? 0 : (cl >= ncell ? ncell - 1 : cl);
      double f1, f2;
      sk_cell_eval(sCoef + (size_t)cl * SK_CELL_STRIDE, sTab, G, r, tc.s, kernel_sin, &f1, &f2);
      sk_emit<SPEC>(spec, f1, f2, cmul, r, j0 + t, stage, acc, old);
    }
  } else {
    // Sparse block (more than cmax cells): the same arithmetic, one warp at a time.  The lanes of a warp
    // that share a cell build that cell's coefficients together in a per-warp scratch area (window from
    // L2, 2 coefficient items per lane, fold by 4 lanes) and then evaluate their targets from it -- bit
    // identical to the block path, whatever the tiling (this is what makes results independent of how
    // the targets are sharded over GPUs).
    sk_cp_async_wait_all();                                // (sR is not used on this path)
    __syncthreads();                                       // sE / sO / sTab ready
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *wWin = sCoef + (size_t)wid * (W * 4 + 4 + SK_NC * 4);   // [W][4] window
    double *wQ = wWin + W * 4;                                       // [4] deconvolution nodes
    double *wCoef = wQ + 4;                                          // [SK_NC][4]
#pragma unroll 1
    for (int u = 0; u < tpt; ++u) {
      const int t = threadIdx.x + u * 256;
      const bool have = t < cnt;
      const double r = have ? xs[j0 + t] : 0.0;
      sk_cplx old;
      old.x = old.y = 0.0;
      if (have && SPEC && !spec.fresh) old = spec.res[j0 + t];
      SkTargetCoord tc;
      tc.l0 = -1;
      tc.s = 0.0;
      if (have) tc = sk_target_coord<W>(G, r);
      unsigned int remaining = __ballot_sync(0xffffffffu, have);
      while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const long long cellL = __shfl_sync(0xffffffffu, tc.l0, leader);
        const unsigned int group = __ballot_sync(0xffffffffu, have && tc.l0 == cellL) & remaining;
        // window of W points x 2 rules = W*4 doubles
        const double *gsrc = reinterpret_cast<const double *>(grid + (size_t)cellL * 2);
        for (int i = lane; i < W * 4; i += 32) wWin[i] = gsrc[i];
        if (lane < 4) {
          const double node = (lane == 0) ? 0.9238795325112867 : (lane == 1) ? 0.3826834323650898
                            : (lane == 2) ? -0.3826834323650898 : -0.9238795325112867;
          const double ymid = (double)(cellL - G.nf2 / 2) + (0.5 * W - 0.5);
          wQ[lane] = sk_deconv(P, G.t_cell * fabs(ymid - 0.5 * node));
        }
        __syncwarp();
        for (int i = lane; i < SK_NC * 4; i += 32) {
          const int comp = i & 3, q = i >> 2;
          wCoef[i] = sk_cell_coef<W>(sE, sO, wWin + comp, 4, q);
        }
        __syncwarp();
        if (lane < 4) {
          double a[4];
          sk_cheb4_to_monomial(wQ, a);
          sk_cell_fold(wCoef + lane, 4, a);
        }
        __syncwarp();
        if (group & (1u << lane)) {
          double f1, f2;
          sk_cell_eval(wCoef, sTab, G, r, tc.s, kernel_sin, &f1, &f2);
          sk_emit<SPEC>(spec, f1, f2, cmul, r, j0 + t, stage, acc, old);
        }
        __syncwarp();
        remaining &= ~group;
      }
    }
  }
  sk_block_reduce_acc(acc, SPEC ? 1 : 0, red);
}

// ---- K4 (log-weighted origin sub-interval, src/quadrature.jl:186-228, dim = 1) --------------------------
// Integration by parts: two :cis transforms per rule, A with integrand f + w log w f' (real part used) and
// B with w log w f (imaginary part used);  I_k = (I0 - Re A_k + 2 pi x Im B_k) / (dim - alpha) with the
// boundary term I0 = b^(dim/2+1-alpha) log(b) f(b) J_{-1/2}(2 pi b x),  J_{-1/2}(z) = sqrt(2/(pi z)) cos z.
struct SkLogwArgs {
  double i0_coef;   // b^(dim/2 + 1 - alpha) * log(b) * f(b)
  double denom;     // dim - alpha
  double b;         // right end of the sub-interval
};
template <int W>
__global__ void __launch_bounds__(256)
k_interp_logw(const __grid_constant__ SkEsPlan P, const __grid_constant__ SkGeom G, const double *__restrict__ xs,
              long long n, const sk_cplx *__restrict__ gridA, const sk_cplx *__restrict__ gridB, double cmul,
              const __grid_constant__ SkLogwArgs L, sk_cplx *__restrict__ stage, SkReduceOut *__restrict__ red) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  unsigned int fl = 0;
  if (j < n) {
    const double x = xs[j];
    double are[2], aim[2], bre[2], bim[2];
    sk_interp_point<W, 2>(P, G, x, gridA, are, aim);
    sk_interp_point<W, 2>(P, G, x, gridB, bre, bim);
    double sn, cs;
    sincospi(2.0 * sk_frac_prod(L.b, x, 0.0), &sn, &cs);            // cos(2 pi b x) with the product exact
    const double z = 6.283185307179586 * (L.b * x);
    const double i0 = L.i0_coef * sqrt(2.0 / (3.141592653589793 * z)) * cs;
    const double tx = 6.283185307179586 * x;
    const double f1 = ((i0 - are[0]) + tx * bim[0]) / L.denom;      // src/quadrature.jl:226-227
    const double f2 = ((i0 - are[1]) + tx * bim[1]) / L.denom;
    sk_stage(f1, f2, cmul, &stage[j], d, fl);
  }
  sk_block_reduce_maxflags(d, fl, red);
}

// direct-summation twin of k_interp_logw: sumsA / sumsB hold the :cis sums of both rules per target
__global__ void k_direct_finish_logw(const sk_cplx *__restrict__ sumsA, const sk_cplx *__restrict__ sumsB,
                                     const double *__restrict__ xs, long long n, double cmul,
                                     const __grid_constant__ SkLogwArgs L, sk_cplx *__restrict__ stage,
                                     SkReduceOut *__restrict__ red) {
  double d = 0.0;
  unsigned int fl = 0;
  for (long long j = threadIdx.x; j < n; j += blockDim.x) {
    const double x = xs[j];
    double sn, cs;
    sincospi(2.0 * sk_frac_prod(L.b, x, 0.0), &sn, &cs);
    const double z = 6.283185307179586 * (L.b * x);
    const double i0 = L.i0_coef * sqrt(2.0 / (3.141592653589793 * z)) * cs;
    const double tx = 6.283185307179586 * x;
    const double f1 = ((i0 - sumsA[2 * j].x) + tx * sumsB[2 * j].y) / L.denom;
    const double f2 = ((i0 - sumsA[2 * j + 1].x) + tx * sumsB[2 * j + 1].y) / L.denom;
    sk_stage(f1, f2, cmul, &stage[j], d, fl);
  }
  sk_block_reduce_maxflags(d, fl, red);
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

// epilogue of the direct branches: same staging as the interpolation kernels.  xdiv != 0 divides by
// x^xdiv (dim > 1, src/quadrature.jl:252-254; the reference scales by c first, then divides).
__global__ void __launch_bounds__(256)
k_direct_finish(const sk_cplx *__restrict__ sums, const double *__restrict__ xs, long long n, double cmul, int kernel_sin,
                double xdiv, sk_cplx *__restrict__ stage, SkReduceOut *__restrict__ red) {
  double d = 0.0;
  unsigned int fl = 0;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const sk_cplx f1 = sums[2 * j], f2 = sums[2 * j + 1];
    double i1 = sk_mul(kernel_sin ? f1.y : f1.x, cmul), i2 = sk_mul(kernel_sin ? f2.y : f2.x, cmul);
    if (xdiv != 0.0) {
      const double den = pow(xs[j], xdiv);
      i1 = i1 / den;
      i2 = i2 / den;
    }
    sk_stage(i1, i2, 1.0, &stage[j], d, fl);
  }
  sk_block_reduce_maxflags(d, fl, red);
}

// direct Bessel summation, int[j] = sum_k buf[k] besselj(nu, 2 pi no[k] x[j])  (src/quadrature.jl:145-160).
// This is the reference's own definition of the dim >= 2 transform; the O(N) NUFHT it uses for large
// problems (FastHankelTransform.jl) is not built, so this branch costs O(M N) at any size.
__global__ void __launch_bounds__(256)
k_direct_bessel(int nu, const double *__restrict__ no1, const double *__restrict__ buf1, long long M1,
                const double *__restrict__ no2, const double *__restrict__ buf2, long long M2,
                const double *__restrict__ xs, sk_cplx *__restrict__ sums) {
  const int r = blockIdx.y;
  const double *no = r ? no2 : no1;
  const double *buf = r ? buf2 : buf1;
  const long long M = r ? M2 : M1;
  const double xj = xs[blockIdx.x];
  double acc = 0.0;
  for (long long k = threadIdx.x; k < M; k += blockDim.x) {
    const double z = sk_mul(sk_mul(6.283185307179586, no[k]), xj);       // 2pi*no[k]*xj
    const double jv = nu == 0 ? j0(z) : (nu == 1 ? j1(z) : jn(nu, z));
    acc = sk_fma(buf[k], jv, acc);
  }
  __shared__ double sr[256];
  sr[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sr[threadIdx.x] += sr[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { sk_cplx o; o.x = sr[0]; o.y = 0.0; sums[(long long)blockIdx.x * 2 + r] = o; }
}

// log-weighted origin sub-interval for dim >= 2 (src/quadrature.jl:189, :204-227): sumsA / sumsB hold the Bessel
// sums of orders nu0 = dim/2-1 (integrand f + w log w f') and nu0+1 (integrand w log w f) of both rules;
// I_k = (I0 - A_k + 2 pi x B_k) / (dim - alpha) with I0 = b^(dim/2+1-alpha) log(b) f(b) J_nu0(2 pi b x), then *c and
// / x^(dim/2-1) (:250-254)
__global__ void __launch_bounds__(256)
k_bessel_logw_finish(const sk_cplx *__restrict__ sumsA, const sk_cplx *__restrict__ sumsB, const double *__restrict__ xs,
                     long long n, double cmul, const __grid_constant__ SkLogwArgs L, int nu0, double xdiv,
                     sk_cplx *__restrict__ stage, SkReduceOut *__restrict__ red) {
  double d = 0.0;
  unsigned int fl = 0;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const double x = xs[j];
    const double z = sk_mul(sk_mul(6.283185307179586, L.b), x);
    const double jv = nu0 == 0 ? j0(z) : (nu0 == 1 ? j1(z) : jn(nu0, z));
    const double i0 = L.i0_coef * jv;
    const double tx = 6.283185307179586 * x;
    const double f1 = ((i0 - sumsA[2 * j].x) + tx * sumsB[2 * j].x) / L.denom;
    const double f2 = ((i0 - sumsA[2 * j + 1].x) + tx * sumsB[2 * j + 1].x) / L.denom;
    double i1 = sk_mul(f1, cmul), i2 = sk_mul(f2, cmul);
    if (xdiv != 0.0) {
      const double den = pow(x, xdiv);
      i1 = i1 / den;
      i2 = i2 / den;
    }
    sk_stage(i1, i2, 1.0, &stage[j], d, fl);
  }
  sk_block_reduce_maxflags(d, fl, red);
}

// ---- K5 ---------------------------------------------------------------------------------------------
// pan = (I, err), stage = (I2, |I2-I1|):  I += I2; err += |I2-I1|   (src/quadrature.jl:261-262).
// The first accepted sub-interval of a panel is not added at all: the staging buffer BECOMES the panel
// buffer (pointer swap on the host; 0 + x == x exactly).
__global__ void k_accept_add(sk_cplx *__restrict__ pan, const sk_cplx *__restrict__ stage, long long n) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  sk_cplx p = pan[j];
  const sk_cplx s = stage[j];
  p.x += s.x;
  p.y += s.y;
  pan[j] = p;
}

// ---- scalars exchanged between the ranks of a target-sharded run (NCCL all-reduces on the stream) -------
struct SkGlobalA {                 // after a sub-interval: MAX over ranks
  unsigned long long maxbits;      // bit pattern of max |I2-I1|
  unsigned long long nan1, nan2, nand;
  unsigned long long err;          // a rank failed before it could evaluate the sub-interval: every rank raises
};
struct SkGlobalB {                 // after a convergence scan
  unsigned long long rbits;        // MAX: bit pattern of the stopping distance
  long long n_lb;                  // SUM: targets each rank keeps active if the walk stopped at its own distance
};
__global__ void k_pack_global_a(const SkReduceOut *__restrict__ red, SkGlobalA *__restrict__ g, int idle, int err) {
  g->err = err ? 1ull : 0ull;
  g->maxbits = idle ? 0ull : red->maxbits;
  const unsigned int fl = idle ? 0u : red->flags;
  g->nan1 = (fl & SK_FLAG_NAN1) ? 1ull : 0ull;
  g->nan2 = (fl & SK_FLAG_NAN2) ? 1ull : 0ull;
  g->nand = (fl & SK_FLAG_NAND) ? 1ull : 0ull;
}
__global__ void k_pack_global_b(SkGlobalB *__restrict__ g, unsigned long long rbits, long long n_lb) {
  g->rbits = rbits;
  g->n_lb = n_lb;
}

__global__ void k_pack_global_b_from_red(const SkReduceOut *__restrict__ red, long long lo, SkGlobalB *__restrict__ g) {
  g->rbits = red->max_unconv >= lo ? red->rbits : 0ull;
  const long long n = red->max_unconv - lo + 1;
  g->n_lb = n > 0 ? n : 0;
}

// ---- peer exchange: the scalar collectives of a target-sharded run over NVLink peer memory -------------------------
// One process per GPU.  Every rank owns a 2 KB "mailbox" in its HBM that all its peers have mapped (CUDA IPC,
// sk_comm_peer_export / sk_comm_peer_attach).  A collective point is ONE single-warp kernel enqueued on the compute
// stream right behind the kernel that produced the local scalars:
//   pack the local words  ->  lane r stores them into rank r's mailbox (slot = this rank), fence, then the epoch word
//   ->  lane r waits for rank r's epoch word in its OWN mailbox  ->  warp-level reduction (u64 MAX; the active count is
//   an integer SUM)  ->  the reduced words go straight into the pinned host scalars the caller reads after its (single)
//   stream synchronisation.
// No NCCL launch, no pack kernel, no D2H copy; the transfer is 64 bytes per peer through the NVSwitch.  Two mailbox
// halves alternate with the epoch parity: a rank can be at most one exchange ahead of a peer that is still reading.
// The wait is bounded (timeout_ns of %globaltimer): a lost peer turns into an error word, not a hang.
constexpr int SK_PEER_MAX = 16;
constexpr int SK_PEER_WORDS = 8;       // 7 payload words + the epoch word: one 64-byte line per (half, rank)
enum { SK_PX_A = 0, SK_PX_AB = 1, SK_PX_B_RED = 2, SK_PX_B_IMM = 3, SK_PX_RANGE = 4, SK_PX_RAW = 5, SK_PX_GATHER = 6,
       SK_PX_SUMMARY = 7 };
struct SkPeerArgs {
  unsigned long long *box[SK_PEER_MAX];   // box[r]: rank r's mailbox as mapped on this device (box[rank]: the own one)
  int rank, n;
  unsigned long long epoch;               // 1, 2, 3, ... identical on all ranks (every rank issues the same exchanges)
  unsigned long long timeout_ns;
  int kind, idle, err, raw_op;            // raw_op (SK_PX_RAW, doubles): 0 max, 1 min, 2 sum
  long long lo;
  unsigned long long imm[7];              // SK_PX_B_IMM: rbits, n_lb;  SK_PX_RAW / SK_PX_GATHER: the words themselves
  int nw;
};
struct SkPeerOut {                        // pinned host memory, written by the kernel
  SkGlobalA ga;
  SkGlobalB gb;
  unsigned long long words[SK_PEER_MAX * 7];   // RANGE: 3 reduced words; RAW: nw reduced words; GATHER: [rank][nw]
  unsigned long long status;              // 0 ok, 1 timed out waiting for a peer
  unsigned long long epoch_done;
  unsigned long long void_flag;           // 1: some rank marked this exchange void -- it does not count, whatever its kind
};
#define SK_PX_VOID (~0ull)                  // maxbits of a void sub-interval exchange (fails every guard's "<" test)
#define SK_PX_VOID_BIT (1ull << 63)         // ... and the mark in the epoch word, which every kind of exchange carries

__device__ __forceinline__ bool sk_chain_guard_holds(const SkSpec &spec) {
  if (spec.gguard != nullptr) {
    const unsigned long long mb = __ldcg(&spec.gguard->ga.maxbits);
    const unsigned long long bad = __ldcg(&spec.gguard->ga.nan1) | __ldcg(&spec.gguard->ga.nan2) |
                                   __ldcg(&spec.gguard->ga.nand) | __ldcg(&spec.gguard->ga.err);
    const unsigned long long rb = __ldcg(&spec.gguard->gb.rbits);
    return mb < spec.guard_maxbits && bad == 0ull && rb == spec.gguard_rbits;      // (SK_PX_VOID fails the first test)
  }
  if (spec.guard == nullptr) return true;
  const unsigned long long mb = __ldcg(&spec.guard->maxbits);
  const unsigned int fl = __ldcg(&spec.guard->flags);
  const long long top = __ldcg(&spec.guard->max_unconv);
  return mb < spec.guard_maxbits && fl == 0u && top == spec.guard_top;
}

__device__ __forceinline__ void sk_st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long sk_ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long sk_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long sk_warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
    v = o > v ? o : v;
  }
  return v;
}

// the exchange as seen by ONE rank (a warp); `a` holds that rank's view of the mailboxes
__device__ __forceinline__ void sk_peer_exchange_warp(const SkPeerArgs &a, const SkReduceOut *__restrict__ red,
                                                      const SkK8State *__restrict__ k8, SkPeerOut *__restrict__ out,
                                                      const SkTargetSummary *__restrict__ sum = nullptr,
                                                      SkPeerOut *__restrict__ dout = nullptr) {
  const int lane = threadIdx.x & 31;
  unsigned long long w[7] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
  const bool quiet = a.idle || a.err;
  if (a.kind == SK_PX_A || a.kind == SK_PX_AB) {
    const unsigned int fl = quiet ? 0u : red->flags;
    w[0] = quiet ? 0ull : red->maxbits;
    w[1] = (fl & SK_FLAG_NAN1) ? 1ull : 0ull;
    w[2] = (fl & SK_FLAG_NAN2) ? 1ull : 0ull;
    w[3] = (fl & SK_FLAG_NAND) ? 1ull : 0ull;
    w[4] = a.err ? 1ull : 0ull;
  }
  // void: this rank's sub-interval was enqueued ahead of time under a device-side guard and skipped itself (its reduction
  // slot carries SK_FLAG_SKIPPED).  The exchange still takes place -- the ranks stay paired -- but the all-ones word wins
  // the MAX on every rank: nobody counts this exchange, every rank issues the next one in its place.
  const bool is_void = (a.kind == SK_PX_A || a.kind == SK_PX_AB) && !quiet && (red->flags & SK_FLAG_SKIPPED) != 0u;
  if (is_void) { w[0] = SK_PX_VOID; w[1] = w[2] = w[3] = 0ull; }
  if (((a.kind == SK_PX_AB && !quiet) || a.kind == SK_PX_B_RED) && !is_void) {
    const long long top = red->max_unconv;
    w[5] = top >= a.lo ? red->rbits : 0ull;
    w[6] = (unsigned long long)(top - a.lo + 1 > 0 ? top - a.lo + 1 : 0);
  }
  if (a.kind == SK_PX_B_IMM) { w[5] = a.imm[0]; w[6] = a.imm[1]; }
  if (a.kind == SK_PX_RANGE) { w[0] = k8->kmin_inv; w[1] = k8->kmax; w[2] = k8->bad ? 1ull : 0ull; }
  if (a.kind == SK_PX_SUMMARY) {
    // what every rank contributes to the start of a run (src/adaptive.jl:123, :152): its smallest positive and its largest
    // unique distance and the number of positive unique distances -- straight from the sort's summary; w[2] flags a
    // summary the host still has to redo (general sort, duplicates to compact): then every rank falls back to a gather
    const long long nu = sum->n_unique;
    const long long lo = sum->r0 == 0.0 ? 1 : 0, cnt = nu - lo > 0 ? nu - lo : 0;
    const double rmin = lo ? sum->r1 : sum->r0;
    w[0] = cnt > 0 ? ~(unsigned long long)__double_as_longlong(rmin) : 0ull;
    w[1] = cnt > 0 ? (unsigned long long)__double_as_longlong(sum->r_last) : 0ull;
    w[2] = (sum->bad || sum->overflow || sum->fixed) ? 1ull : 0ull;
    w[6] = (unsigned long long)cnt;
  }
  if (a.kind == SK_PX_RAW || a.kind == SK_PX_GATHER) {
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = a.imm[i];
  }
  const int half = (int)(a.epoch & 1ull);
  bool ok = true, vseen = false;
  if (lane < a.n) {
    unsigned long long *dst = a.box[lane] + (size_t)(half * SK_PEER_MAX + a.rank) * SK_PEER_WORDS;
#pragma unroll
    for (int i = 0; i < 7; ++i) reinterpret_cast<volatile unsigned long long *>(dst)[i] = w[i];
    __threadfence_system();
    sk_st_release_sys(dst + 7, a.epoch | (is_void ? SK_PX_VOID_BIT : 0ull));
    // now the words of rank `lane`
    const unsigned long long *src = a.box[a.rank] + (size_t)(half * SK_PEER_MAX + lane) * SK_PEER_WORDS;
    const unsigned long long t0 = sk_globaltimer();
    unsigned long long word = 0ull;
    while (((word = sk_ld_acquire_sys(src + 7)) & ~SK_PX_VOID_BIT) < a.epoch) {
      if (sk_globaltimer() - t0 > a.timeout_ns) { ok = false; break; }
    }
    vseen = ok && (word & SK_PX_VOID_BIT) != 0ull;
    __threadfence_system();
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = ok ? reinterpret_cast<const volatile unsigned long long *>(src)[i] : 0ull;
  } else {
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = 0ull;       // neutral for MAX over u64 and for the integer SUM
  }
  const unsigned int all_ok = __all_sync(0xffffffffu, ok);
  if (a.kind == SK_PX_GATHER) {
    if (lane < a.n)
      for (int i = 0; i < a.nw; ++i) out->words[lane * a.nw + i] = w[i];
  } else if (a.kind == SK_PX_RAW) {
    for (int i = 0; i < a.nw; ++i) {
      double v = __longlong_as_double((long long)w[i]);
      if (lane >= a.n) v = a.raw_op == 0 ? -CUDART_INF : (a.raw_op == 1 ? CUDART_INF : 0.0);
#pragma unroll
      for (int m = 16; m; m >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v, m);
        v = a.raw_op == 0 ? fmax(v, o) : (a.raw_op == 1 ? fmin(v, o) : v + o);   // commutative: every lane, every rank agrees
      }
      if (lane == 0) out->words[i] = (unsigned long long)__double_as_longlong(v);
    }
  } else {
    unsigned long long r[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) r[i] = sk_warp_max_u64(w[i]);
    long long s = (long long)w[6];
#pragma unroll
    for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) {
      if (a.kind == SK_PX_A || a.kind == SK_PX_AB) {
        out->ga.maxbits = r[0]; out->ga.nan1 = r[1]; out->ga.nan2 = r[2]; out->ga.nand = r[3]; out->ga.err = r[4];
        if (dout) { dout->ga.maxbits = r[0]; dout->ga.nan1 = r[1]; dout->ga.nan2 = r[2]; dout->ga.nand = r[3]; dout->ga.err = r[4]; }
      }
      if (a.kind == SK_PX_AB || a.kind == SK_PX_B_RED || a.kind == SK_PX_B_IMM) {
        out->gb.rbits = r[5]; out->gb.n_lb = s;
        if (dout) { dout->gb.rbits = r[5]; dout->gb.n_lb = s; }
      }
      if (a.kind == SK_PX_RANGE) { out->words[0] = r[0]; out->words[1] = r[1]; out->words[2] = r[2]; }
      if (a.kind == SK_PX_SUMMARY) { out->words[0] = r[0]; out->words[1] = r[1]; out->words[2] = r[2]; out->words[3] = (unsigned long long)s; }
    }
  }
  const unsigned int any_void = __any_sync(0xffffffffu, vseen);
  if (lane == 0) {
    if (!all_ok) out->status = 1ull;
    out->void_flag = any_void ? 1ull : 0ull;
    out->epoch_done = a.epoch;
  }
}

__global__ void __launch_bounds__(32) k_peer_exchange(SkPeerArgs a, const SkReduceOut *__restrict__ red,
                                                      const SkK8State *__restrict__ k8, SkPeerOut *__restrict__ out,
                                                      const SkTargetSummary *__restrict__ sum, SkPeerOut *__restrict__ dout) {
  sk_peer_exchange_warp(a, red, k8, out, sum, dout);
}

// The same protocol with the ranks emulated as the blocks of ONE cooperative launch on one device (tests: separate
// launches that wait for one another must not share a GPU).  boxes: n mailboxes back to back; block b plays rank b with
// its own local scalars red[b] and its own output outs[b].
__global__ void __launch_bounds__(32) k_peer_exchange_emul(SkPeerArgs a, unsigned long long *boxes, const SkReduceOut *__restrict__ red,
                                                           SkPeerOut *__restrict__ outs) {
  SkPeerArgs mine = a;
  mine.rank = blockIdx.x;
  for (int r = 0; r < a.n; ++r) mine.box[r] = boxes + (size_t)r * 2 * SK_PEER_MAX * SK_PEER_WORDS;
  if (a.kind == SK_PX_RAW || a.kind == SK_PX_GATHER)
    for (int i = 0; i < 7; ++i) mine.imm[i] = a.imm[i] + (unsigned long long)blockIdx.x * (a.kind == SK_PX_GATHER ? 1ull : 0ull);
  sk_peer_exchange_warp(mine, red + blockIdx.x, nullptr, outs + blockIdx.x);
}

// Small device <-> host traffic on the critical path goes through one-warp kernels and mapped pinned memory instead of
// cudaMemcpyAsync: a 32-byte copy queued between two dependent kernels costs a trip through the copy engine (~5 us of
// pipeline bubble), a tiny kernel does not.
__global__ void k_red_init(SkReduceOut *__restrict__ red, long long max_unconv_init) {
  red->maxbits = 0ull; red->flags = 0u; red->_pad = 0u; red->max_unconv = max_unconv_init; red->rbits = 0ull;
}
__global__ void k_zero_words(unsigned long long *__restrict__ p, long long nwords) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long long)gridDim.x * blockDim.x) p[i] = 0ull;
}
__global__ void __launch_bounds__(32) k_publish(unsigned long long *__restrict__ host_dst,
                                                const unsigned long long *__restrict__ dev_src, int nwords) {
  for (int i = threadIdx.x; i < nwords; i += 32) host_dst[i] = dev_src[i];
  __threadfence_system();
}

// roll a rejected speculative commit back: res = backup (bit for bit)
__global__ void k_restore(sk_cplx *__restrict__ res, const sk_cplx *__restrict__ backup, long long n) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) res[j] = backup[j];
}

// ---- K6 ---------------------------------------------------------------------------------------------
// res = (ks, errs) += pan = (I, err)  (src/adaptive.jl:163-164), fused with the convergence predicate of
// src/adaptive.jl:185-197 on panel_ks = I.  The reference walks ix = hi, hi-1, ... while converged;
// equivalently new_hi is the largest index whose predicate is false.  lo0 = global 0-based index of
// element 0.  do_commit = 0 runs the scan only.
__global__ void __launch_bounds__(256)
k_commit_scan(const double *__restrict__ xs, const sk_cplx *__restrict__ pan, sk_cplx *__restrict__ res, long long n,
              long long lo0, int do_commit, double trunc_a, double trunc_num, double xpow, double tau, int criteria,
              SkReduceOut *__restrict__ red) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long bad = -1;
  unsigned long long rb = 0;
  if (j < n) {
    const sk_cplx p = pan[j];
    if (do_commit) {
      sk_cplx r = res[j];
      r.x += p.x;
      r.y += p.y;
      res[j] = r;
    }
    const double x = xs[j];
    const double te = sk_trunc_err(trunc_a, trunc_num, xpow, x, criteria == 0);
    if (!sk_converged(te, p.x, tau, criteria)) { bad = lo0 + j; rb = (unsigned long long)__double_as_longlong(x); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long ob = __shfl_xor_sync(0xffffffffu, bad, o);
    const unsigned long long orb = __shfl_xor_sync(0xffffffffu, rb, o);
    if (ob > bad) { bad = ob; rb = orb; }
  }
  __shared__ long long sb[8];
  __shared__ unsigned long long srb[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sb[wid] = bad; srb[wid] = rb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sb[w] > bad) { bad = sb[w]; rb = srb[w]; }
    if (bad >= 0) {                       // sorted unique distances: max index <=> max distance
      atomicMax(&red->max_unconv, bad);
      atomicMax(&red->rbits, rb);
    }
  }
}

// plain commit (no scan follows, e.g. a flush before reading results)
__global__ void k_commit(const sk_cplx *__restrict__ pan, sk_cplx *__restrict__ res, long long n) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  sk_cplx r = res[j];
  const sk_cplx p = pan[j];
  r.x += p.x;
  r.y += p.y;
  res[j] = r;
}

// errs[ix] += 2*trunc_err for the converged tail (src/adaptive.jl:194)
__global__ void k_scan_add(const double *__restrict__ xs, sk_cplx *__restrict__ res, long long n, double trunc_a,
                           double trunc_num, double xpow, int criteria) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double te = sk_trunc_err(trunc_a, trunc_num, xpow, xs[j], criteria == 0);
  res[j].y += 2 * te;
}

// ---- K8 ---------------------------------------------------------------------------------------------
// keys: bit patterns of the (non-negative) doubles, which order like unsigned integers; -0.0 -> +0.0.
// Also reduces OR / AND of all keys: the bits that differ between any two keys are (OR ^ AND), which
// tells the host which 24 bits are worth radix-sorting (SkKeyBits).
struct SkKeyBits {
  unsigned long long bits_or, bits_and;
  unsigned int bad;        // a distance was NaN / negative / infinite
  unsigned int overflow;   // a run was too long for the two-level sort
  unsigned int unsorted;   // some x[j] <= x[j-1]: the input is not already strictly increasing
  unsigned int _pad;
};
__global__ void __launch_bounds__(256)
k_make_keys(const double *__restrict__ xs, long long n, unsigned long long *__restrict__ keys,
            unsigned int *__restrict__ idx, SkKeyBits *__restrict__ kb) {
  // grid-stride: a few thousand blocks, one pair of atomics per block (a single hot address serialises)
  unsigned long long k_or = 0ull, k_and = ~0ull;
  unsigned int bad = 0, unsorted = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    double x = xs[j];
    if (!(x >= 0.0) || isinf(x)) { bad = 1; x = 0.0; }
    if (j > 0 && !(x > xs[j - 1])) unsorted = 1;          // already sorted and unique? (src/adaptive.jl:113)
    if (x == 0.0) x = 0.0;
    const unsigned long long k = (unsigned long long)__double_as_longlong(x);
    keys[j] = k;
    idx[j] = (unsigned int)j;
    k_or |= k;
    k_and &= k;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    k_or |= __shfl_xor_sync(0xffffffffu, k_or, o);
    k_and &= __shfl_xor_sync(0xffffffffu, k_and, o);
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    unsorted |= __shfl_xor_sync(0xffffffffu, unsorted, o);
  }
  __shared__ unsigned long long s_or[8], s_and[8];
  __shared__ unsigned int s_bad[8], s_uns[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_or[wid] = k_or; s_and[wid] = k_and; s_bad[wid] = bad; s_uns[wid] = unsorted; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { k_or |= s_or[w]; k_and &= s_and[w]; bad |= s_bad[w]; unsorted |= s_uns[w]; }
    atomicOr(&kb->bits_or, k_or);
    atomicAnd(&kb->bits_and, k_and);
    if (bad) atomicOr(&kb->bad, 1u);
    if (unsorted) atomicOr(&kb->unsorted, 1u);
  }
}

__global__ void k_flag_heads(const unsigned long long *__restrict__ keys, long long n, unsigned int *__restrict__ head) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  head[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

// uid = inclusive-scan(head); unique value table uxs[uid - 1] and inverse map (original position ->
// unique id).  The guards only matter when the two-level sort overflowed and left garbage behind (the
// caller then redoes the sort): never write out of bounds.
__global__ void k_scatter_unique(const unsigned long long *__restrict__ keys, const unsigned int *__restrict__ idx,
                                 const unsigned int *__restrict__ head, const unsigned int *__restrict__ uid_incl,
                                 long long n, double *__restrict__ uxs, unsigned int *__restrict__ inv) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned int u = uid_incl[j] - 1u;
  if (head[j] && (long long)u < n) uxs[u] = __longlong_as_double((long long)keys[j]);
  if ((long long)idx[j] < n) inv[idx[j]] = u;
}

__global__ void k_target_summary(const double *__restrict__ uxs, const unsigned int *__restrict__ uid_incl, long long n,
                                 const SkKeyBits *__restrict__ kb, SkTargetSummary *__restrict__ out) {
  long long nu = uid_incl[n - 1];
  if (nu < 1 || nu > n) nu = 1;          // only after an overflowed two-level sort
  out->overflow = kb->overflow;
  out->n_unique = nu;
  out->r0 = uxs[0];
  out->r1 = nu > 1 ? uxs[1] : 0.0;
  out->r_last = uxs[nu - 1];
  out->bad = kb->bad;
  out->presorted = 0u;
  out->fixed = 0u;
}

// values and errors back in the ORIGINAL input order (src/adaptive.jl:105-107) from res = (ks, errs): one
// 16-byte random read per target (random reads are ~2.5x cheaper than the random 8-byte writes of a
// scatter from sorted order, measured: profiles/r1_c_launches.txt).
//
// The truncation term of the error estimate, errs[ix] += 2 trunc_err for the targets a panel's scan found converged
// (src/adaptive.jl:194), is applied HERE instead of in a pass of its own: every unique target belongs to at most one
// converged tail [lo, hi) (the panel after which it left the active set), nothing touches it afterwards, and the
// bound depends on the target only through its distance -- which is the input distance x[j] itself (times the
// warping factor of sk_targets_scale), read coalesced.  Same operands, same operations, same rounding as k_scan_add.
#define SK_MAX_TAILS 48
struct SkTailSeg {
  long long lo, hi;                       // 0-based unique ids [lo, hi)
  double trunc_a, trunc_num, xpow;
  int criteria, _pad;
};
struct SkTailList {
  int n, _pad;
  SkTailSeg seg[SK_MAX_TAILS];
};
// A gather can be enqueued ahead of time behind a chained last panel (sk_results_chain_device): it then runs only if that
// panel turned out accepted (max |I2-I1| below the accept threshold, no NaN, not skipped) with EVERY target converged
// (guard->max_unconv == gtop = lo - 1: the adaptive loop ends, src/adaptive.jl:149), and says so in *ran.
struct SkGatherGuard {
  const SkReduceOut *red;            // nullptr: unconditional
  unsigned long long maxbits;
  long long top;
  unsigned int *ran;                 // (mapped pinned host memory) receives `gen` when the gather runs
  unsigned int gen;
  const SkPeerOut *gl;               // sharded run: the accept test is on the GLOBAL scalars of the chained panel's exchange
};
__global__ void __launch_bounds__(256)
k_gather(const unsigned int *__restrict__ inv, const sk_cplx *__restrict__ res, long long n,
         double *__restrict__ out_v, double *__restrict__ out_e, const double *__restrict__ xin, double xscale,
         const __grid_constant__ SkTailList T, const SkGatherGuard gg = SkGatherGuard{nullptr, 0ull, 0, nullptr, 0u, nullptr}) {
  if (gg.red != nullptr) {
    unsigned long long mb = __ldcg(&gg.red->maxbits);
    const unsigned int fl = __ldcg(&gg.red->flags);
    const long long top = __ldcg(&gg.red->max_unconv);
    unsigned long long gbad = 0ull;
    if (gg.gl != nullptr) {          // every rank's max |I2-I1|, NaN / error words (a void exchange fails the "<" test)
      mb = __ldcg(&gg.gl->ga.maxbits);
      gbad = __ldcg(&gg.gl->ga.nan1) | __ldcg(&gg.gl->ga.nan2) | __ldcg(&gg.gl->ga.nand) | __ldcg(&gg.gl->ga.err);
    }
    if (!(mb < gg.maxbits && fl == 0u && top == gg.top && gbad == 0ull)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *gg.ran = gg.gen; __threadfence_system(); }      // (mapped host memory)
  }
  // 4 independent random reads in flight per thread (the kernel is bound by the latency of the 16-byte reads)
  const long long j0 = ((long long)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
  unsigned int u[4];
  sk_cplx r[4];
  double x[4];
  const bool tails = out_e != nullptr && T.n > 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long j = j0 + (long long)i * 256;
    u[i] = j < n ? __ldg(&inv[j]) : 0u;
    x[i] = (tails && j < n) ? xin[j] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = res[u[i]];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long j = j0 + (long long)i * 256;
    if (j < n) {
      out_v[j] = r[i].x;
      if (out_e) {
        double e = r[i].y;
        if (tails) {
          for (int s = 0; s < T.n; ++s)
            if ((long long)u[i] >= T.seg[s].lo && (long long)u[i] < T.seg[s].hi) {
              const double xx = xscale == 1.0 ? x[i] : sk_mul(x[i], xscale);     // the unique table's own rounding
              e += 2 * sk_trunc_err(T.seg[s].trunc_a, T.seg[s].trunc_num, T.seg[s].xpow, xx, T.seg[s].criteria == 0);
              break;
            }
        }
        out_e[j] = e;
      }
    }
  }
}

// lags of point pairs, computed where they are consumed (src/model.jl:53-68: warp_lags = norm(x_j - x_k) for
// every index pair; config 3's 49 995 000 pairwise distances of 1e4 points never cross PCIe).
// pairs == nullptr: all pairs i < j in row-major order of the strict upper triangle (plus nothing for i == j).
__global__ void k_pair_lags(const double *__restrict__ pts, long long npts, int dim, const long long *__restrict__ pairs,
                            long long npairs, double *__restrict__ lags) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npairs) return;
  long long i, j;
  if (pairs) {
    i = pairs[2 * t];
    j = pairs[2 * t + 1];
  } else {
    // t-th pair of the strict upper triangle: row i has npts-1-i entries
    const double nn = (double)npts;
    long long ii = (long long)floor(((2.0 * nn - 1.0) - sqrt((2.0 * nn - 1.0) * (2.0 * nn - 1.0) - 8.0 * (double)t)) * 0.5);
    if (ii < 0) ii = 0;
    while (ii * (2 * npts - ii - 1) / 2 > t) --ii;                       // fix floating-point off-by-one
    while ((ii + 1) * (2 * npts - ii - 2) / 2 <= t) ++ii;
    i = ii;
    j = t - ii * (2 * npts - ii - 1) / 2 + ii + 1;
  }
  double acc = 0.0;
  for (int d = 0; d < dim; ++d) {
    const double df = sk_add(pts[i * dim + d], -pts[j * dim + d]);
    acc = sk_add(acc, sk_mul(df, df));
  }
  lags[t] = dim == 1 ? fabs(sk_add(pts[i], -pts[j])) : sqrt(acc);
}

// lags under a linear warping x -> x / rho (src/model.jl:62-66 with warp(params, x) = x / params[1], the range
// parameter of scripts/fit_vecchia_demo.jl:15): the sorted unique table is the original one times a positive factor
// (monotone: order and the inverse map are unchanged)
__global__ void k_scale_targets(const double *__restrict__ orig, double *__restrict__ uxs, long long n, double f) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) uxs[j] = sk_mul(orig[j], f);
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

__global__ void k_set_zero_lag(sk_cplx *__restrict__ res, double v) {
  res[0].x = v;
  res[0].y = nan("");
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
