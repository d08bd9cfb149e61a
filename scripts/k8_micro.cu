// k8_micro.cu -- throwaway micro-benchmark behind the design of k_k8_scatter: what do 1e7 scattered global atomics and
// stores cost on B200?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o k8_micro k8_micro.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// xs holds uniform doubles in [1, 2): the 52 mantissa bits are uniform, so the bins are equally loaded
__device__ __forceinline__ unsigned int hbin(unsigned long long k, unsigned int nb) { return (unsigned int)(__umul64hi(k << 12, (unsigned long long)nb)); }

// variant: 0 atomics+2 stores, 1 atomics only, 2 stores only (slot from a hash), 3 atomics on counters padded to 32 B,
// 4 atomic + one 16-byte store, 5 atomic (REDG, no return) only
template <int V>
__global__ void __launch_bounds__(256) k_scatter(const double *__restrict__ xs, long long n, unsigned int nb, unsigned int cap,
                                                 unsigned int *__restrict__ fill, unsigned long long *__restrict__ skeys,
                                                 unsigned int *__restrict__ sidx, ulonglong2 *__restrict__ spair) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = (unsigned long long)__double_as_longlong(xs[j]);
    const unsigned int f = hbin(k, nb);
    unsigned int slot;
    if (V == 2) slot = (unsigned int)((j >> 13) & (cap / 2 - 1));
    else if (V == 3) slot = atomicAdd(&fill[f * 8], 1u);
    else if (V == 5) { atomicAdd(&fill[f], 1u); continue; }
    else slot = atomicAdd(&fill[f], 1u);
    if (V == 1 || V == 3) { if (slot == 0xffffffffu) skeys[0] = 1; continue; }
    slot &= cap - 1;
    const size_t at = (size_t)f * cap + slot;
    if (V == 4) { spair[at] = make_ulonglong2(k, (unsigned long long)j); }
    else { skeys[at] = k; sidx[at] = (unsigned int)j; }
  }
}

// block-staged scatter: chunk of CH elements per block iteration, counting sort by bin in shared memory, one global atomic
// per (chunk, bin), contiguous runs out
template <int CH, int TPB>
__global__ void __launch_bounds__(TPB) k_scatter_staged(const double *__restrict__ xs, long long n, unsigned int nb, unsigned int cap,
                                                        unsigned int *__restrict__ fill, unsigned long long *__restrict__ skeys,
                                                        unsigned int *__restrict__ sidx) {
  extern __shared__ unsigned char smraw[];
  unsigned int *cnt = (unsigned int *)smraw;                 // [nb] count, then global base - local offset
  unsigned int *loc = cnt + nb;                              // [nb] exclusive local offsets
  unsigned long long *skey = (unsigned long long *)(loc + nb);               // [CH] (2 nb ints: 8-byte aligned)
  unsigned int *sj = (unsigned int *)(skey + CH);            // [CH]
  unsigned short *sb = (unsigned short *)(sj + CH);          // [CH]
  __shared__ unsigned int s_w[32];
  const int EPT = CH / TPB;
  const long long nchunk = (n + CH - 1) / CH;
  for (long long c = blockIdx.x; c < nchunk; c += gridDim.x) {
    const long long j0 = c * CH;
    for (unsigned int t = threadIdx.x; t < nb; t += TPB) cnt[t] = 0u;
    __syncthreads();
    unsigned long long key[EPT];
    unsigned int b[EPT], r[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const long long j = j0 + threadIdx.x + e * TPB;
      key[e] = j < n ? (unsigned long long)__double_as_longlong(xs[j]) : 0ull;
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const long long j = j0 + threadIdx.x + e * TPB;
      b[e] = hbin(key[e], nb);
      r[e] = j < n ? atomicAdd(&cnt[b[e]], 1u) : 0u;
    }
    __syncthreads();
    // exclusive scan of cnt over nb bins (nb <= 16 * TPB assumed): thread owns a contiguous range
    const unsigned int per = (nb + TPB - 1) / TPB;
    unsigned int sum = 0;
    for (unsigned int i = 0; i < per; ++i) { const unsigned int q = threadIdx.x * per + i; if (q < nb) sum += cnt[q]; }
    unsigned int inc = sum;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    unsigned int wb = 0;
    for (int w = 0; w < wid; ++w) wb += s_w[w];
    unsigned int run = wb + inc - sum;
    for (unsigned int i = 0; i < per; ++i) {
      const unsigned int q = threadIdx.x * per + i;
      if (q < nb) {
        const unsigned int cq = cnt[q];
        loc[q] = run;
        unsigned int gb = 0;
        if (cq) gb = atomicAdd(&fill[q], cq);
        cnt[q] = gb - run;                                   // global slot of staged position p is cnt[bin] + p
        run += cq;
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const long long j = j0 + threadIdx.x + e * TPB;
      if (j < n) { const unsigned int p = loc[b[e]] + r[e]; skey[p] = key[e]; sj[p] = (unsigned int)j; sb[p] = (unsigned short)b[e]; }
    }
    __syncthreads();
    const int m = (int)((n - j0) < CH ? (n - j0) : CH);
    for (int p = threadIdx.x; p < m; p += TPB) {
      const unsigned int bb = sb[p];
      const unsigned int slot = (cnt[bb] + (unsigned int)p) & (cap - 1);
      const size_t at = (size_t)bb * cap + slot;
      skeys[at] = skey[p];
      sidx[at] = sj[p];
    }
    __syncthreads();
  }
}

// random 4-byte stores (the inverse map) and random 16-byte reads (the gather)
__global__ void k_rand_store4(const unsigned int *__restrict__ perm, long long n, unsigned int *__restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[perm[j]] = (unsigned int)j;
}
__global__ void k_rand_read16(const unsigned int *__restrict__ perm, long long n, const double2 *__restrict__ src, double *__restrict__ o1, double *__restrict__ o2) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) { const double2 v = src[perm[j]]; o1[j] = v.x; o2[j] = v.y; }
}
__global__ void k_rand_read8(const unsigned int *__restrict__ perm, long long n, const double *__restrict__ src, double *__restrict__ o1) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) o1[j] = src[perm[j]];
}

int main(int argc, char **argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 10000000LL;
  std::vector<double> h(n);
  std::vector<unsigned int> perm(n);
  unsigned long long s = 88172645463325252ull;
  for (long long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = 1.0 + (double)(s >> 12) / 4503599627370496.0; perm[i] = (unsigned int)i; }
  for (long long i = n - 1; i > 0; --i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; const long long k = s % (i + 1); std::swap(perm[i], perm[k]); }
  double *xs; unsigned int *fill, *sidx, *dperm, *inv; unsigned long long *skeys; ulonglong2 *spair; double2 *res; double *o1, *o2;
  CK(cudaMalloc(&xs, n * 8)); CK(cudaMemcpy(xs, h.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dperm, n * 4)); CK(cudaMemcpy(dperm, perm.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&inv, n * 4)); CK(cudaMalloc(&res, n * 16)); CK(cudaMalloc(&o1, n * 8)); CK(cudaMalloc(&o2, n * 8));
  CK(cudaMemset(res, 0, n * 16));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char *name, auto launch) {
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      launch(true);
      cudaEventRecord(e0); launch(false); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("%-58s %8.1f us\n", name, best * 1e3);
  };
  for (int tl = 10; tl <= 13; ++tl) {
    const unsigned int nb = (unsigned int)(n >> tl) + 2, cap = 2u << tl;
    CK(cudaMalloc(&fill, (size_t)nb * 32)); CK(cudaMalloc(&skeys, (size_t)nb * cap * 8)); CK(cudaMalloc(&sidx, (size_t)nb * cap * 4));
    CK(cudaMalloc(&spair, (size_t)nb * cap * 16));
    const int gs = 148 * 8;
    char nm[128];
#define RUN(V, label) snprintf(nm, sizeof nm, "target 2^%d (%u bins): %s", tl, nb, label); \
    timeit(nm, [&](bool prep) { if (prep) cudaMemset(fill, 0, (size_t)nb * 32); else k_scatter<V><<<gs, 256>>>(xs, n, nb, cap, fill, skeys, sidx, spair); });
    RUN(0, "atomic + 8B + 4B stores (current)")
    RUN(1, "atomic (with return) only")
    RUN(5, "atomic (no return) only")
    RUN(3, "atomic on counters padded to 32 B")
    RUN(2, "8B + 4B stores only")
    RUN(4, "atomic + one 16B store")
    if (nb <= 16 * 512) {
      snprintf(nm, sizeof nm, "target 2^%d: staged, chunk 8192 / 512 thr", tl);
      const size_t sm = (size_t)nb * 8 + 8192 * 14;
      cudaFuncSetAttribute(k_scatter_staged<8192, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      timeit(nm, [&](bool prep) { if (prep) cudaMemset(fill, 0, (size_t)nb * 32); else k_scatter_staged<8192, 512><<<148 * 2, 512, sm>>>(xs, n, nb, cap, fill, skeys, sidx); });
      snprintf(nm, sizeof nm, "target 2^%d: staged, chunk 16384 / 1024 thr", tl);
      const size_t sm2 = (size_t)nb * 8 + 16384 * 14;
      if (sm2 <= 227 * 1024) {
        cudaFuncSetAttribute(k_scatter_staged<16384, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        timeit(nm, [&](bool prep) { if (prep) cudaMemset(fill, 0, (size_t)nb * 32); else k_scatter_staged<16384, 1024><<<148, 1024, sm2>>>(xs, n, nb, cap, fill, skeys, sidx); });
      }
    }
    cudaFree(fill); cudaFree(skeys); cudaFree(sidx); cudaFree(spair);
  }
  timeit("random 4B stores (inverse map), 1e7", [&](bool prep) { if (!prep) k_rand_store4<<<(unsigned)((n + 255) / 256), 256>>>(dperm, n, inv); });
  timeit("random 16B reads + 2 coalesced writes (gather)", [&](bool prep) { if (!prep) k_rand_read16<<<(unsigned)((n + 255) / 256), 256>>>(dperm, n, res, o1, o2); });
  timeit("random 8B reads + 1 coalesced write", [&](bool prep) { if (!prep) k_rand_read8<<<(unsigned)((n + 255) / 256), 256>>>(dperm, n, (const double *)res, o1); });
  for (size_t g : {32, 64, 128}) {
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    char nm[96]; snprintf(nm, sizeof nm, "gather with L2 fetch granularity %zu (got %zu)", g, got);
    timeit(nm, [&](bool prep) { if (!prep) k_rand_read16<<<(unsigned)((n + 255) / 256), 256>>>(dperm, n, res, o1, o2); });
  }
  return 0;
}
