// sk_k8.h -- index arithmetic of K8 (unique / sort / inverse map; src/adaptive.jl:99-107, :113-120), written as
// host/device inline functions so that tests/emul can check the monotonicity of the bin maps without a GPU.
//
// The distances are non-negative doubles, whose bit patterns ("keys") order like unsigned integers.  K8 never
// radix-sorts them.  It learns a piecewise-linear estimate of the key distribution from a coarse histogram
// (SK_K8_NC bins over [kmin, kmax]), cuts it into "fine bins" of ~SK_K8_TARGET elements, throws every element into
// its fine bin (one pass, fixed-capacity slots) and finishes each fine bin inside one thread block's shared memory
// (sort, de-duplicate, emit the unique table and the inverse map).  Both maps below are monotone non-decreasing in
// the key, which is all the scheme needs for the bin-major order to be the sorted order.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SK_K8_HD __host__ __device__ __forceinline__
#else
#define SK_K8_HD inline
#endif

#define SK_K8_NC_LOG 13
#define SK_K8_NC (1 << SK_K8_NC_LOG)          // coarse bins
#define SK_K8_TARGET_LOG 10
#define SK_K8_TARGET (1 << SK_K8_TARGET_LOG)  // estimated elements per fine bin
#define SK_K8_CAP 2048                        // slots per fine bin (2x slack over the estimate)
#define SK_K8_NSSB 2048                       // sub-bins of the in-block counting sort
#define SK_K8_SAMPLE_MIN (1 << 20)            // below this many inputs the coarse histogram sees every element
#define SK_K8_FILL_STRIDE 8                   // one fill counter per 32-byte sector (measured: the L2 atomic units
                                              // serialise per sector; eight counters per sector cost 45 % more time)

struct SkK8State {              // device scalars of one sk_targets_set (zero-initialised by a memset)
  unsigned long long kmin_inv;  // ~min key over the positive inputs (atomicMax of ~key; 0 = no positive input)
  unsigned long long kmax;      // max key
  unsigned long long nzero;     // inputs equal to zero (they all map to unique id 0 and are not binned)
  unsigned long long mul;       // sk_k8_mul(kmin, kmax)                              (written by k_k8_plan)
  unsigned long long ndesc;     // number of j with x[j] <= x[j-1]; 0: the input is already strictly increasing
  unsigned int bad;             // a distance was NaN / negative / infinite
  unsigned long long ndup;      // duplicates dropped by k_k8_finish, all bins
  unsigned int overflow;        // 1: a fine bin outgrew its slots
  unsigned int nfine;           // fine bins in use                                   (written by k_k8_plan)
  unsigned int n_slots;         // positive inputs in the bins = sum of the fills       (written by k_k8_scan_bins)
};

// Coarse bins tile the key range [kmin, kmax] exactly: with M = floor(2^64 SK_K8_NC / (kmax - kmin + 1)) the 128-bit
// product (key - kmin) * M has the coarse bin in its high word and the position inside the bin (as a fraction of 2^64)
// in its low word.  A range of at most SK_K8_NC keys gets M = 0: every key is its own coarse bin.
SK_K8_HD unsigned long long sk_k8_mul(unsigned long long kmin, unsigned long long kmax) {
  const unsigned long long range1 = kmax - kmin + 1ull;
  if (range1 <= (unsigned long long)SK_K8_NC) return 0ull;
  return (unsigned long long)((((unsigned __int128)1) << (64 + SK_K8_NC_LOG)) / range1);
}

SK_K8_HD unsigned long long sk_k8_mulhi(unsigned long long a, unsigned long long b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#endif
}

// segment s (32 consecutive inputs) is part of the coarse-histogram sample: a multiplicative hash instead of a fixed
// stride, so that periodic inputs (rows of a distance matrix) do not alias with the sample.  A sample of segments is
// a fair sample of the values only when the input order is unrelated to the values: a (nearly) sorted input -- fewer
// than one descent in 16 elements -- is histogrammed in full (sk_k8_samp).
SK_K8_HD unsigned int sk_k8_samp(unsigned long long n, unsigned long long ndesc) {
  return (n >= (unsigned long long)SK_K8_SAMPLE_MIN && ndesc * 16ull >= n) ? 8u : 1u;
}
SK_K8_HD bool sk_k8_sampled(unsigned long long s, unsigned int samp) {
  if (samp <= 1u) return true;
  const unsigned int h = (unsigned int)(s * 2654435761ull) ^ (unsigned int)((s * 2654435761ull) >> 32);
  return ((h * 2246822519u) >> 16) % samp == 0u;
}

// coarse bin and in-bin fraction (x 2^64) of a key
SK_K8_HD void sk_k8_coarse(unsigned long long key, unsigned long long kmin, unsigned long long mul, unsigned int *c,
                           unsigned long long *frac) {
  const unsigned long long d = key - kmin;
  if (mul == 0ull) { *c = (unsigned int)d; *frac = 0ull; return; }
  *c = (unsigned int)sk_k8_mulhi(d, mul);
  *frac = d * mul;                      // low word of the product
}

// estimated number of elements below `key`: piecewise linear in the key, from the coarse histogram
// (tab[c].x = exclusive cumulative estimate, tab[c].y = estimate inside coarse bin c).  Monotone non-decreasing in key
// because the coarse bin is, tab[c + 1].x = tab[c].x + tab[c].y and the in-bin term is < tab[c].y.
SK_K8_HD unsigned int sk_k8_fine_bin(unsigned long long key, unsigned long long kmin, unsigned long long mul,
                                     const unsigned int *ccum, const unsigned int *ccnt) {
  unsigned int c;
  unsigned long long frac;
  sk_k8_coarse(key, kmin, mul, &c, &frac);
  const unsigned long long est = (unsigned long long)ccum[c] + sk_k8_mulhi(frac, (unsigned long long)ccnt[c]);
  return (unsigned int)(est >> SK_K8_TARGET_LOG);
}

// sub-bin of the in-block counting sort: linear over the bin's own key range [lo, hi]; monotone (integer -> double
// conversion, multiplication by a positive constant and truncation all are)
SK_K8_HD double sk_k8_ssb_scale(unsigned long long lo, unsigned long long hi) {
  return (double)SK_K8_NSSB / ((double)(hi - lo) + 1.0);
}
SK_K8_HD int sk_k8_ssb(unsigned long long key, unsigned long long lo, double scale) {
  const int s = (int)((double)(key - lo) * scale);
  return s < SK_K8_NSSB - 1 ? s : SK_K8_NSSB - 1;
}
