// sk_k8.cuh -- K8: unique / sort / inverse map of the input distances (src/adaptive.jl:99-107, :113-120) as five
// hand-written kernels and no radix sort (index arithmetic and the scheme: sk_k8.h).
//
//   k_k8_stats    one read pass: validity, "already strictly increasing?", key range, number of zeros
//   k_k8_sample   coarse histogram (SK_K8_NC bins over the key range) of a hashed 1-in-`samp` sample of 32-element
//                 segments (of everything below SK_K8_SAMPLE_MIN inputs)
//   k_k8_plan     one block: scan of the coarse histogram -> piecewise-linear estimate of the key distribution
//   k_k8_scatter  every element -> its fine bin (~SK_K8_TARGET elements each, SK_K8_CAP slots): one global atomic for
//                 the slot (one counter per 32-byte sector), one 16-byte (key, index) store; zeros are answered directly
//   k_k8_offsets  one block: exclusive scan of the bins' fill counts = where each bin's unique values start if no
//                 distance occurs twice
//   k_k8_finish   one block per fine bin, in shared memory: counting sort on SK_K8_NSSB sub-bins, exact rank inside
//                 the (tiny) sub-bin groups, first-of-value flags, block scan; emits the sorted unique table and the
//                 inverse map at the bin's offset and records how many duplicates it dropped.  Blocks are independent
//                 of each other (no look-back, no spinning).
//   k_k8_fix_*    only when some bin dropped duplicates (all three return at once otherwise): scan of the dropped
//                 counts, compaction of the unique table into a second buffer, shift of the inverse map
//   k_k8_summary  n_unique, the two smallest and the largest unique distance, flags -> one read-back
//
// HBM traffic per input distance: 8 (stats) + 1 (sample) + 8 + 16 (scatter) + 16 + 8 + 4 (finish) = 61 bytes; the
// radix-sort pipeline this replaces moved ~150.  Nothing here depends on the order in which atomics resolve: the
// unique table is the sorted set and inv[j] is the rank of x[j] in it, whatever slot the element landed in.
#pragma once
#include <cuda_runtime.h>

#include "sk_k8.h"

#define SK_K8_TPB 256
#define SK_K8_EPT (SK_K8_CAP / SK_K8_TPB)

struct SkTargetSummary {        // written by k_k8_summary
  long long n_unique;
  double r0, r1, r_last;        // smallest, second smallest and largest unique distance
  unsigned int bad;
  unsigned int overflow;        // the bin scheme did not apply (clustered / heavily duplicated input): general sort
  unsigned int presorted;       // the input was already strictly increasing: no sort at all
  unsigned int fixed;           // duplicates were dropped: the compacted unique table is in the second buffer
};

__device__ __forceinline__ unsigned long long sk_k8_key(double x, unsigned int *bad) {
  if (!(x >= 0.0) || isinf(x)) { *bad = 1u; x = 0.0; }
  if (x == 0.0) x = 0.0;                                  // -0.0 -> +0.0
  return (unsigned long long)__double_as_longlong(x);
}

// ---- pass 0 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_k8_stats(const double *__restrict__ xs, long long n, SkK8State *__restrict__ st) {
  unsigned long long kmin_inv = 0ull, kmax = 0ull, nz = 0ull, nd = 0ull;
  unsigned int bad = 0;
  // four consecutive distances per thread and step (two 16-byte loads); the distance before the quad comes from the
  // neighbouring lane (lane 0 reads it)
  const long long nquad = (n + 3) / 4;
  const int lane_ = threadIdx.x & 31;
  // two quads per thread and step (the second one block-stride further): four 16-byte loads in flight
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q0 = (long long)blockIdx.x * blockDim.x; q0 < nquad; q0 += 2 * stride) {
    double x[2][4];
    int m[2];
    long long jq[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long q = q0 + h * stride + threadIdx.x, j = 4 * q;
      jq[h] = j;
      m[h] = 0;
      x[h][0] = x[h][1] = x[h][2] = x[h][3] = 0.0;
      if (q < nquad) {
        if (j + 3 < n) {
          const double2 a = *reinterpret_cast<const double2 *>(xs + j), b = *reinterpret_cast<const double2 *>(xs + j + 2);
          x[h][0] = a.x; x[h][1] = a.y; x[h][2] = b.x; x[h][3] = b.y;
          m[h] = 4;
        } else {
          for (; j + m[h] < n; ++m[h]) x[h][m[h]] = xs[j + m[h]];
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double prev = __shfl_up_sync(0xffffffffu, x[h][3], 1);
      if (lane_ == 0 && m[h] > 0 && jq[h] > 0) prev = xs[jq[h] - 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < m[h]) {
          const unsigned long long k = sk_k8_key(x[h][i], &bad);
          if ((jq[h] + i > 0) && !(x[h][i] > prev)) ++nd;     // already sorted and unique? (src/adaptive.jl:113)
          prev = x[h][i];
          if (k == 0ull) ++nz;
          else {
            kmin_inv = kmin_inv > ~k ? kmin_inv : ~k;
            kmax = kmax > k ? kmax : k;
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin_inv, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin_inv = kmin_inv > a ? kmin_inv : a;
    kmax = kmax > b ? kmax : b;
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
    nd += __shfl_xor_sync(0xffffffffu, nd, o);
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  __shared__ unsigned long long s_a[8], s_b[8], s_z[8], s_d[8];
  __shared__ unsigned int s_f[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_a[wid] = kmin_inv; s_b[wid] = kmax; s_z[wid] = nz; s_d[wid] = nd; s_f[wid] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int f = 0;
    nz = nd = 0;
    for (int w = 0; w < 8; ++w) {
      kmin_inv = kmin_inv > s_a[w] ? kmin_inv : s_a[w];
      kmax = kmax > s_b[w] ? kmax : s_b[w];
      nz += s_z[w];
      nd += s_d[w];
      f |= s_f[w];
    }
    if (kmin_inv) atomicMax(&st->kmin_inv, kmin_inv);
    if (kmax) atomicMax(&st->kmax, kmax);
    if (nz) atomicAdd(&st->nzero, nz);
    if (nd) atomicAdd(&st->ndesc, nd);
    if (f) atomicOr(&st->bad, 1u);
  }
}

// ---- pass 1: coarse histogram of a sample --------------------------------------------------------------
// One warp per sampled 32-element segment; shared-memory privatised histogram, flushed with global atomics.
__global__ void __launch_bounds__(512)
k_k8_sample(const double *__restrict__ xs, long long n, const SkK8State *__restrict__ st,
            unsigned int *__restrict__ chist) {
  if (!st->ndesc || !st->kmin_inv) return;
  __shared__ unsigned int s_h[SK_K8_NC];
  for (int t = threadIdx.x; t < SK_K8_NC; t += blockDim.x) s_h[t] = 0u;
  __syncthreads();
  const unsigned long long kmin = ~st->kmin_inv;
  const unsigned long long mul = sk_k8_mul(kmin, st->kmax);
  const unsigned int samp = sk_k8_samp((unsigned long long)n, st->ndesc);
  const long long nseg = (n + 31) / 32;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = warp0; s < nseg; s += nwarp) {
    if (!sk_k8_sampled((unsigned long long)s, samp)) continue;
    const long long j = s * 32 + lane;
    if (j < n) {
      unsigned int bad = 0;
      const unsigned long long k = sk_k8_key(xs[j], &bad);
      if (k) {
        unsigned int cb;
        unsigned long long frac;
        sk_k8_coarse(k, kmin, mul, &cb, &frac);
        atomicAdd(&s_h[cb], 1u);
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < SK_K8_NC; t += blockDim.x) {
    const unsigned int v = s_h[t];
    if (v) atomicAdd(&chist[t], v);
  }
}

// ---- plan: one block of 1024 threads ---------------------------------------------------------------------
// ctab[c] = (estimated number of positive inputs below coarse bin c, estimated number inside it), scaled from the
// sample to the n - nzero positive inputs by integer arithmetic (every estimate is rounded down, so the total never
// exceeds the number of inputs and the host's bound on the number of fine bins holds).
__global__ void __launch_bounds__(1024)
k_k8_plan(SkK8State *__restrict__ st, const unsigned int *__restrict__ chist, long long n, uint2 *__restrict__ ctab) {
  if (!st->ndesc) return;
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_tot;
  const int per = SK_K8_NC / 1024;
  unsigned int h[per];
  unsigned long long loc = 0;
#pragma unroll
  for (int i = 0; i < per; ++i) { h[i] = chist[threadIdx.x * per + i]; loc += h[i]; }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = s_w[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_w[lane] = wi - w;                          // exclusive over the warps
    if (lane == 31) s_tot = wi;
  }
  __syncthreads();
  const unsigned long long sampled = s_tot;
  const unsigned long long npos = (unsigned long long)n - st->nzero;
  unsigned long long run = s_w[wid] + (inc - loc);   // sampled elements below this thread's first coarse bin
  // a bin's estimate is floor(cum_hi * npos / sampled) - floor(cum_lo * npos / sampled): sums telescope exactly
  auto scaled = [&](unsigned long long cum) -> unsigned long long {
    if (sampled == 0ull) return 0ull;
    // cum <= sampled <= 2^31, npos < 2^31: the product fits in 64 bits
    return cum * npos / sampled;
  };
  // up to SK_K8_CAP inputs fit one block of k_k8_finish whatever their distribution: a single fine bin
  const bool single = npos <= (unsigned long long)SK_K8_CAP;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    const unsigned long long lo = scaled(run), hi = scaled(run + h[i]);
    run += h[i];
    uint2 e;
    e.x = single ? 0u : (unsigned int)lo;
    e.y = single ? 0u : (unsigned int)(hi - lo);
    ctab[threadIdx.x * per + i] = e;
  }
  if (threadIdx.x == 0) {
    st->mul = st->kmin_inv ? sk_k8_mul(~st->kmin_inv, st->kmax) : 0ull;
    st->nfine = npos == 0ull ? 0u : (single ? 1u : (unsigned int)((scaled(sampled) >> SK_K8_TARGET_LOG) + 1ull));
  }
}

// ---- pass 2: scatter into the fine bins ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_k8_scatter(const double *__restrict__ xs, long long n, SkK8State *__restrict__ st, const uint2 *__restrict__ ctab,
             unsigned int *__restrict__ fill, ulonglong2 *__restrict__ slots, unsigned int *__restrict__ inv) {
  if (!st->ndesc) return;
  const unsigned long long kmin = ~st->kmin_inv;
  const unsigned long long mul = st->mul;
  unsigned int over = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    unsigned int bad = 0;
    const unsigned long long k = sk_k8_key(xs[j], &bad);
    if (k == 0ull) { inv[j] = 0u; continue; }             // every zero is unique id 0
    unsigned int cb;
    unsigned long long frac;
    sk_k8_coarse(k, kmin, mul, &cb, &frac);
    const uint2 e = __ldg(&ctab[cb]);
    const unsigned int f = (unsigned int)(((unsigned long long)e.x + __umul64hi(frac, (unsigned long long)e.y)) >> SK_K8_TARGET_LOG);
    const unsigned int slot = atomicAdd(&fill[(size_t)f * SK_K8_FILL_STRIDE], 1u);
    if (slot < (unsigned int)SK_K8_CAP) {
      slots[(size_t)f * SK_K8_CAP + slot] = make_ulonglong2(k, (unsigned long long)j);   // (key, index): one 16-byte store
    } else {
      over = 1u;
    }
  }
  if (__any_sync(0xffffffffu, over) && (threadIdx.x & 31) == 0) atomicOr(&st->overflow, 1u);
}

// ---- exclusive scan of min(fill, cap) over the bins in use: one block of 1024 threads ---------------------------------
// src: one value per bin at stride `stride` (the padded fill counters, or the dropped-duplicate counts); dst[b] =
// sum of the values of bins < b; *total = sum over all bins.  which = 0: provisional unique offsets (always);
// which = 1: dropped duplicates (only when there are any).
__global__ void __launch_bounds__(1024)
k_k8_scan_bins(SkK8State *__restrict__ st, const unsigned int *__restrict__ src, int stride, unsigned int cap,
               unsigned int *__restrict__ dst, int which) {
  if (!st->ndesc) return;
  if (which == 1 && st->ndup == 0ull) return;
  __shared__ unsigned int s_w[32];
  __shared__ unsigned int s_run;
  const unsigned int nb = st->nfine;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0u;
  __syncthreads();
  for (unsigned int b0 = 0; b0 < nb; b0 += 1024u) {
    const unsigned int b = b0 + threadIdx.x;
    unsigned int v = 0u;
    if (b < nb) { v = src[(size_t)b * stride]; v = v < cap ? v : cap; }
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    unsigned int wbase = 0u;
    for (int w = 0; w < wid; ++w) wbase += s_w[w];
    const unsigned int base = s_run;
    if (b < nb) dst[b] = base + wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_run = base + wbase + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0 && which == 0) st->n_slots = s_run;
}

// ---- pass 3: finish every fine bin in shared memory ----------------------------------------------------------
__global__ void __launch_bounds__(SK_K8_TPB, 4)
k_k8_finish(SkK8State *__restrict__ st, const unsigned int *__restrict__ fill, const ulonglong2 *__restrict__ slots,
            const unsigned int *__restrict__ uoffp, unsigned int *__restrict__ dup, double *__restrict__ uxs,
            unsigned int *__restrict__ inv) {
  if (!st->ndesc) return;
  __shared__ unsigned long long s_key[SK_K8_CAP];        // placed order, later final (sorted) order
  __shared__ int s_off[SK_K8_NSSB + 1];                  // sub-bin counts, then exclusive offsets (+ total)
  __shared__ unsigned short s_luid[SK_K8_CAP];           // heads at positions <= p
  __shared__ unsigned char s_head[SK_K8_CAP];
  __shared__ unsigned long long s_lo[8], s_hi[8];
  __shared__ int s_w[8];
  const unsigned int bin = blockIdx.x;
  if (bin >= st->nfine) return;
  const unsigned int fl = fill[(size_t)bin * SK_K8_FILL_STRIDE];
  const int cnt = (int)(fl < (unsigned int)SK_K8_CAP ? fl : (unsigned int)SK_K8_CAP);
  const size_t base = (size_t)bin * SK_K8_CAP;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  unsigned long long key[SK_K8_EPT];
  unsigned int oidx[SK_K8_EPT];
  unsigned long long lo = ~0ull, hi = 0ull;
#pragma unroll
  for (int e = 0; e < SK_K8_EPT; ++e) {
    const int t = threadIdx.x + e * SK_K8_TPB;
    key[e] = 0ull;
    oidx[e] = 0u;
    if (t < cnt) {
      const ulonglong2 rec = slots[base + t];
      key[e] = rec.x;
      oidx[e] = (unsigned int)rec.y;
      lo = lo < key[e] ? lo : key[e];
      hi = hi > key[e] ? hi : key[e];
    }
  }
  for (int t = threadIdx.x; t <= SK_K8_NSSB; t += SK_K8_TPB) s_off[t] = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = lo < a ? lo : a;
    hi = hi > b ? hi : b;
  }
  if (lane == 0) { s_lo[wid] = lo; s_hi[wid] = hi; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) { lo = lo < s_lo[w] ? lo : s_lo[w]; hi = hi > s_hi[w] ? hi : s_hi[w]; }
  const double scale = cnt > 0 ? sk_k8_ssb_scale(lo, hi) : 0.0;

  // counting sort on the sub-bins: the returned count is the element's rank inside its sub-bin group
  int ssb[SK_K8_EPT], rk[SK_K8_EPT];
#pragma unroll
  for (int e = 0; e < SK_K8_EPT; ++e) {
    const int t = threadIdx.x + e * SK_K8_TPB;
    ssb[e] = 0; rk[e] = 0;
    if (t < cnt) {
      ssb[e] = sk_k8_ssb(key[e], lo, scale);
      rk[e] = atomicAdd(&s_off[ssb[e]], 1);
    }
  }
  __syncthreads();
  {  // exclusive scan of the SK_K8_NSSB counts; thread t owns sub-bins [t*per, (t+1)*per)
    const int per = SK_K8_NSSB / SK_K8_TPB;
    int c[per], sum = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) { c[i] = s_off[threadIdx.x * per + i]; sum += c[i]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) wbase += (w < wid) ? s_w[w] : 0;
    int run = wbase + inc - sum;
#pragma unroll
    for (int i = 0; i < per; ++i) { s_off[threadIdx.x * per + i] = run; run += c[i]; }
    if (threadIdx.x == SK_K8_TPB - 1) s_off[SK_K8_NSSB] = run;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < SK_K8_EPT; ++e) {
    const int t = threadIdx.x + e * SK_K8_TPB;
    if (t < cnt) s_key[s_off[ssb[e]] + rk[e]] = key[e];
  }
  __syncthreads();
  // exact position inside the group (groups hold ~1 element for smooth inputs; any size is handled)
  int fin[SK_K8_EPT];
  unsigned int headbits = 0;
#pragma unroll
  for (int e = 0; e < SK_K8_EPT; ++e) {
    const int t = threadIdx.x + e * SK_K8_TPB;
    fin[e] = 0;
    if (t < cnt) {
      const int o = s_off[ssb[e]], g_end = s_off[ssb[e] + 1], me = o + rk[e];
      int less = 0, eqb = 0;
      if (g_end - o > 1) {                                    // most groups hold one element
        for (int p = o; p < g_end; ++p) {
          const unsigned long long k2 = s_key[p];
          less += k2 < key[e];
          eqb += (k2 == key[e]) & (p < me);
        }
      }
      fin[e] = o + less + eqb;
      if (eqb == 0) headbits |= 1u << e;
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < SK_K8_EPT; ++e) {
    const int t = threadIdx.x + e * SK_K8_TPB;
    if (t < cnt) { s_key[fin[e]] = key[e]; s_head[fin[e]] = (headbits >> e) & 1u; }
  }
  __syncthreads();
  int nuniq;
  {  // inclusive scan of the head flags over the final positions; thread t owns positions [t*EPT, (t+1)*EPT)
    int c[SK_K8_EPT], sum = 0;
#pragma unroll
    for (int i = 0; i < SK_K8_EPT; ++i) {
      const int p = threadIdx.x * SK_K8_EPT + i;
      c[i] = p < cnt ? (int)s_head[p] : 0;
      sum += c[i];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { wbase += (w < wid) ? s_w[w] : 0; tot += s_w[w]; }
    nuniq = tot;
    int run = wbase + inc - sum;
#pragma unroll
    for (int i = 0; i < SK_K8_EPT; ++i) { run += c[i]; s_luid[threadIdx.x * SK_K8_EPT + i] = (unsigned short)run; }
  }
  // The bin's unique values start where the preceding bins' ELEMENTS end (k_k8_scan_bins): exact when no distance
  // occurs twice.  Dropped duplicates are recorded; k_k8_fix_* then shift the table and the map (rare path).
  const unsigned int ndrop = (unsigned int)(cnt - nuniq);
  if (threadIdx.x == 0) {
    dup[bin] = ndrop;
    if (ndrop) atomicAdd(&st->ndup, (unsigned long long)ndrop);
  }
  const unsigned int uoff = (st->nzero ? 1u : 0u) + uoffp[bin];
  __syncthreads();                                         // s_luid complete
  // (positions strided by the block: consecutive lanes read consecutive shared-memory words and write consecutive
  //  -- or, after a dropped duplicate, nearly consecutive -- table entries)
#pragma unroll
  for (int i = 0; i < SK_K8_EPT; ++i) {
    const int p = threadIdx.x + i * SK_K8_TPB;
    if (p < cnt && s_head[p]) uxs[uoff + s_luid[p] - 1u] = __longlong_as_double((long long)s_key[p]);
  }
#pragma unroll
  for (int e = 0; e < SK_K8_EPT; ++e) {
    const int t = threadIdx.x + e * SK_K8_TPB;
    if (t < cnt) inv[oidx[e]] = uoff + s_luid[fin[e]] - 1u;
  }
}

// (process-per-GPU runs) the local key range as two words for a MAX all-reduce: ~kmin and kmax
__global__ void k_k8_pack_range(const SkK8State *__restrict__ st, unsigned long long *__restrict__ out) {
  out[0] = st->kmin_inv;
  out[1] = st->kmax;
  out[2] = st->bad ? 1ull : 0ull;
}

// already sorted and unique input: the unique table is the input itself and the inverse map is the identity
__global__ void k_k8_identity(const double *__restrict__ xs, long long n, const SkK8State *__restrict__ st,
                              double *__restrict__ uxs, unsigned int *__restrict__ inv) {
  if (st->ndesc) return;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const double x = xs[j];
    uxs[j] = x == 0.0 ? 0.0 : x;
    inv[j] = (unsigned int)j;
  }
}

// ---- duplicates were dropped: compact the unique table and shift the inverse map (all return at once otherwise) -----
// uxs_out[zbase + uoffp[b] - D[b] + i] = uxs[zbase + uoffp[b] + i], i < fill[b] - dup[b]; one block per bin
__global__ void __launch_bounds__(256)
k_k8_fix_uxs(const SkK8State *__restrict__ st, const unsigned int *__restrict__ fill, const unsigned int *__restrict__ uoffp,
             const unsigned int *__restrict__ dup, const unsigned int *__restrict__ dsum, const double *__restrict__ uxs,
             double *__restrict__ uxs_out) {
  if (!st->ndesc || st->ndup == 0ull) return;
  const unsigned int bin = blockIdx.x;
  if (bin >= st->nfine) return;
  const unsigned int zbase = st->nzero ? 1u : 0u;
  unsigned int cnt = fill[(size_t)bin * SK_K8_FILL_STRIDE];
  cnt = cnt < (unsigned int)SK_K8_CAP ? cnt : (unsigned int)SK_K8_CAP;
  const unsigned int nu = cnt - dup[bin], from = zbase + uoffp[bin], to = from - dsum[bin];
  for (unsigned int i = threadIdx.x; i < nu; i += blockDim.x) uxs_out[to + i] = uxs[from + i];
  if (bin == 0 && threadIdx.x == 0 && zbase) uxs_out[0] = 0.0;
}
// inv[j] -= D[bin of inv[j]]: the bin of a provisional unique id by binary search in the provisional offsets
__global__ void __launch_bounds__(256)
k_k8_fix_inv(const SkK8State *__restrict__ st, const unsigned int *__restrict__ uoffp, const unsigned int *__restrict__ dsum,
             long long n, unsigned int *__restrict__ inv) {
  if (!st->ndesc || st->ndup == 0ull) return;
  const unsigned int zbase = st->nzero ? 1u : 0u, nb = st->nfine;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const unsigned int u = inv[j];
    if (u < zbase) continue;                                 // the zero distance
    const unsigned int v = u - zbase;
    unsigned int a = 0u, b = nb;                             // last bin with uoffp[bin] <= v
    while (b - a > 1u) {
      const unsigned int mid = (a + b) >> 1;
      if (__ldg(&uoffp[mid]) <= v) a = mid; else b = mid;
    }
    inv[j] = u - __ldg(&dsum[a]);
  }
}

__global__ void k_k8_summary(const SkK8State *__restrict__ st, double *__restrict__ uxs, double *__restrict__ uxs_fix,
                             long long n, SkTargetSummary *__restrict__ out) {
  const bool presorted = st->ndesc == 0ull;
  const bool fixed = !presorted && st->ndup != 0ull;        // the compacted table lives in uxs_fix
  double *tab = fixed ? uxs_fix : uxs;
  long long nu = presorted ? n : (long long)(st->nzero ? 1 : 0) + (long long)st->n_slots - (long long)st->ndup;
  if (nu < 1 || nu > n) nu = 1;                // only after an overflow / invalid input (the host then discards this)
  if (!presorted && st->nzero) tab[0] = 0.0;
  out->n_unique = nu;
  out->r0 = tab[0];
  out->r1 = nu > 1 ? tab[1] : 0.0;
  out->r_last = tab[nu - 1];
  out->bad = st->bad;
  out->overflow = st->overflow;
  out->presorted = presorted ? 1u : 0u;
  out->fixed = fixed ? 1u : 0u;
}
