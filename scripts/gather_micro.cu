// micro-benchmark behind the layout of the final gather (results to the input order): random 16-byte reads from a 160 MB
// (ks, errs) table against two passes of random 8-byte reads from 80 MB tables that may stay L2-resident.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/gather_micro.cu -o scripts/gather_micro
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include <cuda_runtime.h>
struct c2 { double x, y; };
__global__ void g_aos(const unsigned *__restrict__ inv, const c2 *__restrict__ res, long long n, double *ov, double *oe) {
  const long long j0 = ((long long)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
  unsigned u[4]; c2 r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { long long j = j0 + i * 256; u[i] = j < n ? inv[j] : 0u; }
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = res[u[i]];
#pragma unroll
  for (int i = 0; i < 4; ++i) { long long j = j0 + i * 256; if (j < n) { ov[j] = r[i].x; oe[j] = r[i].y; } }
}
template <int HINT>
__global__ void g_soa1(const unsigned *__restrict__ inv, const double *__restrict__ tab, long long n, double *out) {
  const long long j0 = ((long long)blockIdx.x * blockDim.x) * 8 + threadIdx.x;
  unsigned u[8]; double r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { long long j = j0 + i * 256; u[i] = j < n ? (HINT ? __ldcs(&inv[j]) : inv[j]) : 0u; }
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __ldg(&tab[u[i]]);
#pragma unroll
  for (int i = 0; i < 8; ++i) { long long j = j0 + i * 256; if (j < n) { if (HINT) __stcs(&out[j], r[i]); else out[j] = r[i]; } }
}
__global__ void g_soa2(const unsigned *__restrict__ inv, const double *__restrict__ ta, const double *__restrict__ tb, long long n,
                       double *ov, double *oe) {
  const long long j0 = ((long long)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
  unsigned u[4]; double a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { long long j = j0 + i * 256; u[i] = j < n ? inv[j] : 0u; }
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = __ldg(&ta[u[i]]); b[i] = __ldg(&tb[u[i]]); }
#pragma unroll
  for (int i = 0; i < 4; ++i) { long long j = j0 + i * 256; if (j < n) { ov[j] = a[i]; oe[j] = b[i]; } }
}
int main() {
  const long long n = 10000000;
  std::vector<unsigned> perm(n);
  std::iota(perm.begin(), perm.end(), 0u);
  std::mt19937_64 rng(1);
  std::shuffle(perm.begin(), perm.end(), rng);
  unsigned *inv; c2 *res; double *ta, *tb, *ov, *oe, *flush;
  cudaMalloc(&inv, 4 * n); cudaMalloc(&res, 16 * n); cudaMalloc(&ta, 8 * n); cudaMalloc(&tb, 8 * n);
  cudaMalloc(&ov, 8 * n); cudaMalloc(&oe, 8 * n); cudaMalloc(&flush, 512ll << 20);
  cudaMemcpy(inv, perm.data(), 4 * n, cudaMemcpyHostToDevice);
  cudaMemset(res, 0, 16 * n); cudaMemset(ta, 0, 8 * n); cudaMemset(tb, 0, 8 * n);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char *name, auto fn) {
    float best = 1e9, tot = 0;
    for (int it = 0; it < 6; ++it) {
      cudaMemsetAsync(flush, it, 512ll << 20);          // flush the L2 between repetitions
      cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) { best = std::min(best, ms); tot += ms; }
    }
    printf("%-58s best %.4f ms  mean %.4f ms\n", name, best, tot / 5);
  };
  const unsigned g4 = (unsigned)((n + 1023) / 1024), g8 = (unsigned)((n + 2047) / 2048);
  timeit("AoS: one pass, 16-byte reads from the 160 MB table", [&] { g_aos<<<g4, 256>>>(inv, res, n, ov, oe); });
  timeit("SoA: one pass, two 8-byte reads (2 x 80 MB tables)", [&] { g_soa2<<<g4, 256>>>(inv, ta, tb, n, ov, oe); });
  timeit("SoA: two passes of 8-byte reads, plain", [&] { g_soa1<0><<<g8, 256>>>(inv, ta, n, ov); g_soa1<0><<<g8, 256>>>(inv, tb, n, oe); });
  timeit("SoA: two passes of 8-byte reads, streams evict-first", [&] { g_soa1<1><<<g8, 256>>>(inv, ta, n, ov); g_soa1<1><<<g8, 256>>>(inv, tb, n, oe); });
  timeit("SoA: ONE pass of 8-byte reads (values only), evict-first", [&] { g_soa1<1><<<g8, 256>>>(inv, ta, n, ov); });
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
