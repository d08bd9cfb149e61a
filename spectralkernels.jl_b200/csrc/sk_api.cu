// sk_api.cu -- the C ABI (include/spectralkernels_b200.h) over the sm_100a kernels.
//
// One sk_ctx owns one CUDA stream, the exp-of-semicircle plan, the quadrature rules, the sorted
// unique targets and every O(N) work array of the adaptive loop.  Only scalars cross the ABI per
// sub-interval (a, b in; max|I2-I1| out) and per outer panel (new highest unconverged index out).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cufft.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/spectralkernels_b200.h"
#include "sk_host_util.h"
#include "sk_kernels.cuh"
#include "sk_hankel.cuh"
#include "sk_rules.cuh"

#define SK_GATHER_SLICES 8

namespace {

// NVTX ranges named after the reference's TimerOutputs stages (src/quadrature.jl:99,108,113,140,145,
// src/adaptive.jl:156,162,182), so that a timeline of the GPU path reads like `SpectralKernels.TIMER`.
// Header-only NVTX3: no cost unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// NCCL is loaded with dlopen when (and only when) a communicator is requested: a single-GPU user needs no
// NCCL at all, and inside a torch process the already loaded libnccl.so.2 is reused.
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd)
        api.handle = h;
    }
  }
  return api.handle ? &api : nullptr;
}

struct HostScalars {          // pinned mirror of the device scalars
  SkReduceOut red;
  SkTargetSummary sum;
  SkK8State k8;
  SkKeyBits kb;
  SkGlobalA ga;
  SkGlobalB gb;
  double hv[32];              // generic host-value collectives
  double r[2];
  SkHankelGroup grp[SK_HK_NGRP];   // transform groups of the current Hankel sub-interval
};

}  // namespace

struct sk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string errmsg;
  double eps = 1e-15;
  SkEsPlan plan;
  std::map<std::tuple<long long, int, int>, cufftHandle> fft_plans;      // (size, batch, 0 compute stream / 1 prefetch stream)

  // rules
  int m = 0, k = 0;
  double p = 0.0;
  bool have_rule = false, have_jac = false, rule_generated = false;
  DevBuf<double> leg_no1, leg_wt1, leg_no2, leg_wt2, jac_no1, jac_wt1, jac_no2, jac_wt2;
  std::vector<double> h_rule[8];
  // generated rules, kept for the life of the context: (n, p) -> (nodes, weights).  Derivative configs flip
  // between p = 0 and p = 1 (src/adaptive.jl:42), and a generation costs ~0.3 s of host time at n = 8192
  // (the generated rules themselves are cached process-wide: rule_cache() below)

  // integrand
  int family = SK_SDF_HOST, deriv = 0, nparam = 0;
  double params[SK_NPARAM_MAX] = {0};

  // sources of the current sub-interval
  DevBuf<double> no1, buf1, no2, buf2, pos_hi1, pos_lo1, pos_hi2, pos_lo2, imz;
  DevBuf<sk_cplx> cs1, cs2, fft, fftB, dsum, dsumB;
  // Source-side prefetch: nodes, strengths, spread and FFT of a sub-interval the driver is about to ask for (the first
  // panel while the distances are still being sorted; the second panel while the first is interpolated) are computed
  // ahead on a second stream into a second buffer set.  A request is served from it only if its panel spec and its
  // transform geometry equal the prefetched ones bit for bit -- otherwise the prefetch is simply dropped.
  struct SrcSet {
    DevBuf<double> no1, buf1, no2, buf2, pos_hi1, pos_lo1, pos_hi2, pos_lo2;
    DevBuf<sk_cplx> cs1, cs2, fft;
  } pf, pf2;                                 // pf: the next request; pf2: the one after it (promoted when pf is taken)
  bool pf_valid = false, need_gen = false, prefetch_on = true;
  SkPanelSpec pf_S, pend_S;
  SkGeom pf_G;
  bool pf2_valid = false;
  SkPanelSpec pf2_S;
  SkGeom pf2_G;
  cudaEvent_t pf2_ev = nullptr;
  cudaStream_t stream_main = nullptr, stream2 = nullptr;
  cudaEvent_t pf_ev = nullptr;
  cudaEvent_t k8_ev = nullptr;
  // sk_targets_set_device: K8 reads the caller's buffer directly while the library's own copy of the distances (the
  // final gather reads them again, possibly in a later call) is made on the copy stream next to the sort
  const double *in_src = nullptr;
  cudaEvent_t in_ev[2] = {nullptr, nullptr};
  cudaEvent_t ev_slice[SK_GATHER_SLICES] = {nullptr};
  bool results_sliced = false;
  int last_logw = 0;
  bool in_group = false, early_pending = false, early_global = false;
  long long n_pf_hits = 0, n_pf_issued = 0;
  DevBuf<double> bufb1, bufb2;               // second integrand of the log-weighted origin sub-interval
  bool have_sources = false;

  // targets
  long long n_in = 0, n_unique = 0;
  bool has_zero = false;
  DevBuf<double> in, uxs, uxs_fix, uxs_orig, out_v, out_e;
  bool have_orig = false;                     // uxs_orig holds the unscaled unique distances (sk_targets_scale)
  double in_scale = 1.0;                      // unique distance = input distance * in_scale (sk_targets_scale)
  SkTailList tails;                           // converged tails whose 2 trunc_err is added by the gather (k_gather)
  double r0_orig = 0, r1_orig = 0, r_last_orig = 0;
  // asynchronous result copies (sk_results_get_async): two slots, a copy stream, events
  DevBuf<double> aout_v[2], aout_e[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_gather[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
  bool copy_pending[2] = {false, false};
  int aslot = 0;
  DevBuf<sk_cplx> res, pan, stage;   // (ks, errs), (I, err), (I2, |I2-I1|) per unique target
  DevBuf<unsigned long long> keys, keys_alt;
  DevBuf<unsigned int> idx, idx_alt, head, uid, inv;
  DevBuf<unsigned char> cub_tmp;
  SkKeyBits *d_kb = nullptr;
  // K8 (sk_k8.cuh): control block (state, coarse histogram, fill counters, offsets, dropped-duplicate counts), the
  // piecewise-linear distribution estimate, and the fine-bin slots
  DevBuf<unsigned char> k8_ctl;
  DevBuf<uint2> k8_ctab;
  DevBuf<ulonglong2> k8_slots;
  bool have_targets = false;

  // panel state (0-based half-open [lo, hi))
  long long lo = 0, hi = 0;
  double r_lo = 0, r_hi = 0;
  bool in_panel = false, staged = false, first_accept = true, commit_pending = false;
  long long pend_lo = 0, pend_hi = 0;        // range of the pending (lazy) commit
  double r0 = 0, r1 = 0, r_last = 0;         // smallest, second smallest, largest unique distance
  double early_lo = 0, early_hi = 0;         // distance range known after the first K8 pass (sk_targets_early_range)
  long long begin_n = -1;                    // between sk_targets_begin* and sk_targets_end
  bool sub_open = false;                     // between sk_subinterval_begin and sk_subinterval_end
  long long scan_hi = -1;                    // 1-based index / distance returned by the last scan
  double scan_r = 0;
  int interp_mode = 0;                       // 0: cell polynomials (default), 1: per-target taps
  // nonuniform Hankel transform (dim >= 2): 0 auto, 1 always the direct Bessel summation, 2 always the O(N) scheme
  int hankel_mode = 0;
  bool hk_tab_ready = false;
  DevBuf<double> hk_tab, hk_vals, hk_cheb, hk_loc, hk_lam1, hk_lam2;
  DevBuf<long long> hk_lev;
  DevBuf<SkHankelGroup> hk_groups;
  DevBuf<sk_cplx> hk_grid, hk_part, hk_modes;
  bool smem_attr_set[SK_WMAX + 1] = {false};
  // target-sharded multi-GPU: scalar NCCL all-reduces on the context's stream
  ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  // ... or as single-warp exchange kernels over peer-mapped mailboxes (sk_comm_peer_*; k_peer_exchange)
  int peer_n = 0;                            // ranks attached (0: no peer exchange)
  unsigned long long *peer_box = nullptr;    // this rank's mailbox (device memory, exported through CUDA IPC)
  unsigned long long *peer_map[SK_PEER_MAX] = {nullptr};   // the peers' mailboxes as mapped here
  unsigned long long peer_epoch = 0;
  SkPeerOut *peer_out[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // pinned: sub-interval / scan, early range, host values,
                                                                           // run summary, chained sub-interval
  SkPeerOut *d_gout[2] = {nullptr, nullptr}; // device mirrors of slots 0 and 4 (the guards of chained launches read them)
  int peer_slot = 0;                         // where the scalars of the sub-interval being finished are: 0, or 4 (chained)
  const SkReduceOut *cur_red = nullptr;      // its local reduction slot (d_red, or d_red2 for an adopted chained launch)
  // chained launches in sharded runs: on for 2 ranks, off beyond (SK_SHARDED_CHAIN=0/1 overrides).  Measured: 2 GPUs
  // 1.234 -> 1.20 ms per step, 8 GPUs 1.258 -> 1.270 ms -- a chained panel cannot start before EVERY rank's previous panel
  // has gone through the exchange, so with more ranks the skew eats what the saved host round trip gives
  bool sharded_chain = false;
  bool peer_summary_sent = false;            // the sort's summary went out behind it (sk_comm_summary)
  double peer_timeout_s = 20.0;
  SkGlobalA *d_ga = nullptr;
  SkGlobalB *d_gb = nullptr;
  double *d_hv = nullptr;
  double g_max_abs = 0;                      // global results of the last collectives
  double g_r_stop = 0;
  long long g_n_lb = 0;
  // speculative commit of the panel's first sub-interval (see sk_subinterval_opts::speculate)
  bool spec_active = false, spec_accepted = false;
  // (ks, errs) = 0 at the start of a run (src/adaptive.jl:122) is not written until somebody needs it: the first panel's
  // speculative commit covers every positive distance and writes the table outright (SkSpec::fresh)
  bool res_zero_pending = false, zero_lag_written = false, spec_fresh = false;
  bool pend_timed = false, pend_spec = false;  // between transform_and_stage_enqueue and _finish
  bool scan_from_spec = false;                 // between converge_scan_enqueue and _finish
  int pend_rc = 0;                             // local result of transform_and_stage_enqueue in a sharded run
  bool pend_ab = false;                        // the sub-interval's collective carried the scan's scalars as well
  long long panel_subs = 0;                  // sub-intervals evaluated in the open panel
  long long n_act_global = 0;                // active targets over all ranks (0: this rank only)
  sk_scan_args spec_args;
  long long spec_new_hi = 0;
  double spec_r = 0;
  SkTargetSummary *d_sum = nullptr;

  SkReduceOut *d_red = nullptr;
  SkReduceOut *red_target = nullptr;         // where the interpolation kernel reduces to (d_red, or the chain's d_red2)
  // chained launch of the next panel's first sub-interval (sk_subinterval_chain, SkSpec::guard)
  struct Chain {
    bool pending = false, adopted = false;
    double a = 0, b = 0, r_lo = 0, r_hi = 0;
    long long lo = 0, hi = 0;
    sk_subinterval_opts o;
    sk_scan_args sa;
  } chain;
  SkGeom chain_G;
  // the FIRST panel's first sub-interval enqueued behind the sort (sk_first_panel_early; SkSpec::dyn)
  struct Early {
    bool pending = false, adopted = false;
    double a = 0, b = 0, r_lo = 0, r_hi = 0;
    long long lo = 0, n_in = 0;
    sk_subinterval_opts o;
    sk_scan_args sa;
    SkGeom G;
  } early1;
  cudaEvent_t ev_sum = nullptr;              // behind the D2H copy of the sort's summary
  // the final gather enqueued behind a chained last panel (sk_results_chain_device; SkGatherGuard)
  struct GatherChain {
    bool pending = false;
    double *vals = nullptr, *errs = nullptr;
    SkTailList tails;
  } gchain;
  unsigned int *d_gran = nullptr, *h_gran = nullptr;
  unsigned int gran_gen = 0;                 // generation written by the chained gather that ran
  SkReduceOut *d_red2 = nullptr, *h_red2 = nullptr;
  cudaEvent_t ev_red = nullptr, ev_red2 = nullptr;
  HostScalars *h_scal = nullptr;  // pinned

  // stats / timing
  sk_stats stats;
  long long launches_total = 0;              // kernels launched since the context was created (sk_stats.launches_total)
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_user[2] = {nullptr, nullptr};
};

namespace {

int fail(sk_ctx *c, int code, const char *fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    c->errmsg = buf;
  }
  return code;
}

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(c, e_ == cudaErrorMemoryAllocation ? SK_ERR_ALLOC : SK_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                        \
  } while (0)

#define LAUNCH_CHECK()                                                                          \
  do {                                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                        \
    if (e_ != cudaSuccess)                                                                      \
      return fail(c, SK_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    c->stats.kernel_launches++, c->launches_total++;                                                                 \
  } while (0)

#define DISPATCH_W(w, CALL)                \
  switch (w) {                             \
    case 4: CALL(4); break;                \
    case 6: CALL(6); break;                \
    case 8: CALL(8); break;                \
    case 10: CALL(10); break;              \
    case 12: CALL(12); break;              \
    case 14: CALL(14); break;              \
    default: CALL(16); break;              \
  }

inline unsigned int nblk(long long n, int b) { return (unsigned int)((n + b - 1) / b); }
int flush_commit(sk_ctx *c);
int ensure_res_zero(sk_ctx *c);
// small scalars to / from the device without the copy engine (k_red_init, k_publish: mapped pinned memory)
int red_init(sk_ctx *c, SkReduceOut *d, long long max_unconv_init);
int publish(sk_ctx *c, void *host_dst, const void *dev_src, size_t bytes);
int chain_discard(sk_ctx *c);
int early_discard(sk_ctx *c);

#define NCK(call)                                                                                     \
  do {                                                                                                \
    ncclResult_t r_ = (call);                                                                         \
    if (r_ != ncclSuccess)                                                                            \
      return fail(c, SK_ERR_CUDA, "NCCL: %s (%s:%d)", nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "error", \
                  __FILE__, __LINE__);                                                                \
  } while (0)

// collective A (after a sub-interval): MAX of max|I2-I1| and of the NaN flags.  Enqueued on the stream
// right behind the kernel that filled d_red; the caller's read-back then also fetches h_scal->ga.
// err != 0: this rank could not evaluate the sub-interval (allocation failure, geometry out of range, ...).  It still
// joins the collective -- its peers are already waiting in it -- with neutral values and the error word set, so that
// every rank sees the failure and raises instead of blocking in the all-reduce for ever.
inline bool sharded(const sk_ctx *c) { return c->comm != nullptr || c->peer_n > 0; }
void peer_release(sk_ctx *c) {
  if (c->peer_n > 0 || c->peer_box) {
    cudaStreamSynchronize(c->stream);
    for (int r = 0; r < SK_PEER_MAX; ++r) {
      if (c->peer_map[r] && c->peer_map[r] != c->peer_box) cudaIpcCloseMemHandle(c->peer_map[r]);
      c->peer_map[r] = nullptr;
    }
    if (c->peer_box) cudaFree(c->peer_box);
    c->peer_box = nullptr;
    c->peer_n = 0;
    c->peer_epoch = 0;
  }
}

// one exchange over the peer mailboxes (see k_peer_exchange): enqueued on the compute stream, the reduced words land in
// pinned host memory (slot 0: sub-interval / scan scalars, 1: the early key range, 2: host values)
int peer_exchange(sk_ctx *c, int slot, int kind, int idle, int err, long long lo, const unsigned long long *imm, int nw,
                  int raw_op, const SkK8State *k8 = nullptr, const SkTargetSummary *sum = nullptr,
                  const SkReduceOut *red = nullptr) {
  SkPeerArgs a;
  std::memset(&a, 0, sizeof(a));
  for (int r = 0; r < c->peer_n; ++r) a.box[r] = c->peer_map[r];
  a.rank = c->comm_rank;
  a.n = c->peer_n;
  a.epoch = ++c->peer_epoch;
  a.timeout_ns = (unsigned long long)(c->peer_timeout_s * 1e9);
  a.kind = kind; a.idle = idle; a.err = err; a.raw_op = raw_op; a.lo = lo; a.nw = nw;
  for (int i = 0; i < 7; ++i) a.imm[i] = (imm && i < nw) ? imm[i] : 0ull;
  SkPeerOut *dout = slot == 0 ? c->d_gout[0] : (slot == 4 ? c->d_gout[1] : nullptr);
  k_peer_exchange<<<1, 32, 0, c->stream>>>(a, red ? red : c->d_red, k8, c->peer_out[slot], sum, dout);
  LAUNCH_CHECK();
  return SK_OK;
}
int peer_check(sk_ctx *c, int slot) {   // after the stream synchronisation
  if (c->peer_n > 0 && c->peer_out[slot]->status) {
    c->peer_out[slot]->status = 0;
    return fail(c, SK_ERR_STATE, "peer exchange timed out: a rank did not reach the collective point within %.0f s", c->peer_timeout_s);
  }
  return SK_OK;
}

int comm_reduce_a(sk_ctx *c, int idle, int err = 0) {
  if (!sharded(c)) return SK_OK;
  if (c->peer_n > 0) return peer_exchange(c, 0, SK_PX_A, idle, err, c->lo, nullptr, 0, 0);
  NcclApi *N = nccl_api();
  k_pack_global_a<<<1, 1, 0, c->stream>>>(c->d_red, c->d_ga, (idle || err) ? 1 : 0, err);
  LAUNCH_CHECK();
  NCK(N->AllReduce(c->d_ga, c->d_ga, 5, ncclUint64, ncclMax, c->comm, c->stream));
  CK(cudaMemcpyAsync(&c->h_scal->ga, c->d_ga, sizeof(SkGlobalA), cudaMemcpyDeviceToHost, c->stream));
  return SK_OK;
}
// collectives A and B of a speculated sub-interval in ONE launch: the interpolation kernel produced max |I2-I1| and
// the scan's (stopping distance, active count) together, so they travel together; the scan that follows then needs
// no collective at all.  Idle ranks contribute neutral values.
int comm_reduce_ab(sk_ctx *c, int idle, int err) {
  if (!sharded(c)) return SK_OK;
  if (c->peer_n > 0) return peer_exchange(c, 0, SK_PX_AB, idle, err, c->lo, nullptr, 0, 0);
  NcclApi *N = nccl_api();
  k_pack_global_a<<<1, 1, 0, c->stream>>>(c->d_red, c->d_ga, (idle || err) ? 1 : 0, err);
  LAUNCH_CHECK();
  if (idle || err) k_pack_global_b<<<1, 1, 0, c->stream>>>(c->d_gb, 0ull, 0);
  else k_pack_global_b_from_red<<<1, 1, 0, c->stream>>>(c->d_red, c->lo, c->d_gb);
  LAUNCH_CHECK();
  NCK(N->GroupStart());
  NCK(N->AllReduce(c->d_ga, c->d_ga, 5, ncclUint64, ncclMax, c->comm, c->stream));
  NCK(N->AllReduce(&c->d_gb->rbits, &c->d_gb->rbits, 1, ncclUint64, ncclMax, c->comm, c->stream));
  NCK(N->AllReduce(&c->d_gb->n_lb, &c->d_gb->n_lb, 1, ncclInt64, ncclSum, c->comm, c->stream));
  NCK(N->GroupEnd());
  CK(cudaMemcpyAsync(&c->h_scal->ga, c->d_ga, sizeof(SkGlobalA), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&c->h_scal->gb, c->d_gb, sizeof(SkGlobalB), cudaMemcpyDeviceToHost, c->stream));
  return SK_OK;
}
// collective B (after a scan): MAX of the stopping distance, SUM of the per-rank lower bounds of the
// number of targets that stay active.  from_red: take the values from d_red (lo = first index of the panel).
int comm_reduce_b(sk_ctx *c, bool from_red, unsigned long long rbits, long long n_lb) {
  if (!sharded(c)) return SK_OK;
  if (c->peer_n > 0) {
    const unsigned long long imm[2] = {rbits, (unsigned long long)n_lb};
    return peer_exchange(c, 0, from_red ? SK_PX_B_RED : SK_PX_B_IMM, 0, 0, c->lo, imm, 2, 0);
  }
  NcclApi *N = nccl_api();
  if (from_red) k_pack_global_b_from_red<<<1, 1, 0, c->stream>>>(c->d_red, c->lo, c->d_gb);
  else k_pack_global_b<<<1, 1, 0, c->stream>>>(c->d_gb, rbits, n_lb);
  LAUNCH_CHECK();
  NCK(N->GroupStart());
  NCK(N->AllReduce(&c->d_gb->rbits, &c->d_gb->rbits, 1, ncclUint64, ncclMax, c->comm, c->stream));
  NCK(N->AllReduce(&c->d_gb->n_lb, &c->d_gb->n_lb, 1, ncclInt64, ncclSum, c->comm, c->stream));
  NCK(N->GroupEnd());
  CK(cudaMemcpyAsync(&c->h_scal->gb, c->d_gb, sizeof(SkGlobalB), cudaMemcpyDeviceToHost, c->stream));
  return SK_OK;
}
const SkGlobalA &global_a(const sk_ctx *c) { return c->peer_n > 0 ? c->peer_out[c->peer_slot]->ga : c->h_scal->ga; }
const SkGlobalB &global_b(const sk_ctx *c) { return c->peer_n > 0 ? c->peer_out[c->peer_slot]->gb : c->h_scal->gb; }
// a void exchange (some rank's chained launch skipped itself, k_peer_exchange) does not count: every rank issues the
// exchange again -- with the same local scalars, or after evaluating the sub-interval it had skipped
inline bool peer_void(const sk_ctx *c) { return c->peer_n > 0 && c->peer_out[c->peer_slot]->void_flag != 0ull; }
void comm_take_a(sk_ctx *c, unsigned int *fl, double *mx) {   // after the stream sync
  if (!sharded(c)) return;
  const SkGlobalA &g = global_a(c);
  std::memcpy(mx, &g.maxbits, sizeof(double));
  *fl = (g.nan1 ? SK_FLAG_NAN1 : 0u) | (g.nan2 ? SK_FLAG_NAN2 : 0u) | (g.nand ? SK_FLAG_NAND : 0u);
}
void comm_take_b(sk_ctx *c) {
  if (!sharded(c)) return;
  const SkGlobalB &g = global_b(c);
  std::memcpy(&c->g_r_stop, &g.rbits, sizeof(double));
  c->g_n_lb = g.n_lb;
}

int width_from_eps(double eps) {
  int w = (int)std::ceil(-std::log10(eps / 10.0));
  if (w & 1) ++w;
  if (w < 4) w = 4;
  if (w > SK_WMAX) w = SK_WMAX;
  return w;
}

int get_fft_plan(sk_ctx *c, long long nf2, int batch, cufftHandle *out) {
  auto key = std::make_tuple(nf2, batch, c->stream == c->stream_main ? 0 : 1);
  auto it = c->fft_plans.find(key);
  if (it != c->fft_plans.end()) {
    *out = it->second;
    return SK_OK;
  }
  cufftHandle h;
  int n[1] = {(int)nf2};
  int embed[1] = {(int)nf2};
  // interleaved batch: element l of transform r lives at [l*batch + r]
  cufftResult r = cufftPlanMany(&h, 1, n, embed, batch, 1, embed, batch, 1, CUFFT_Z2Z, batch);
  if (r != CUFFT_SUCCESS) return fail(c, SK_ERR_CUFFT, "cufftPlanMany(n=%lld, batch=%d) failed: %d", nf2, batch, (int)r);
  r = cufftSetStream(h, c->stream);
  if (r != CUFFT_SUCCESS) return fail(c, SK_ERR_CUFFT, "cufftSetStream failed: %d", (int)r);
  c->fft_plans[key] = h;
  *out = h;
  return SK_OK;
}

template <int W>
int launch_interp_session(sk_ctx *c, const SkGeom &G, const double *xs, long long n, double cmul, int ksin, const SkSpec &spec) {
  if (c->interp_mode == 1) {
    if (spec.on)
      k_interp_session<W, true><<<nblk(n, 256), 256, 0, c->stream>>>(c->plan, G, xs, n, c->fft.p, cmul, ksin, c->stage.p + c->lo, spec, (c->red_target ? c->red_target : c->d_red));
    else
      k_interp_session<W, false><<<nblk(n, 256), 256, 0, c->stream>>>(c->plan, G, xs, n, c->fft.p, cmul, ksin, c->stage.p + c->lo, spec, (c->red_target ? c->red_target : c->d_red));
    return 0;
  }
  // cells the active targets span -> average targets per cell -> how many cells a block may hold
  // targets per thread: big launches take 6144-target blocks (24 per thread: the per-block work -- window, cell
  // polynomials, four barriers -- is amortised over more targets, and 1e7 targets make 5.5 waves of 2 blocks per SM;
  // measured 0.169 -> 0.159 ms per launch against 4096-target blocks, 8192 would leave one block per SM)
  // smaller launches shrink the blocks so that the grid still makes >= 4 waves of the 296 resident blocks
  const int tpt = std::min(24, std::max(4, 4 * (int)(n / (296LL * 4 * 256 * 4))));
  const int tpb = 256 * tpt;
  const double span = (c->r_hi - c->r_lo) * G.kap_hi + 1.0;
  const double per_block = span * (double)tpb / (double)n;
  // cells a block may hold in shared memory (a block that spans more takes the warp path): 1.3 x the average + 8
  const int cmax = std::min(96, std::max(32, 8 * (int)std::ceil((1.3 * per_block + 8.0) / 8.0)));
  const size_t smem = sizeof(double) * (size_t)(2 * (W / 2) * (SK_NC / 2) + (cmax + W) * 4 + cmax * 4 + cmax * SK_CELL_STRIDE + tpb);
  // function attributes are per device: remember per context (one context = one device)
  bool &attr_set = c->smem_attr_set[W];
  if (!attr_set) {
    cudaFuncSetAttribute(k_interp_cells<W, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(k_interp_cells<W, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    attr_set = true;
  }
#define SK_LAUNCH_CELLS(SPECV, MINBV)                                                                                  \
  k_interp_cells<W, SPECV, MINBV><<<nblk(n, tpb), 256, smem, c->stream>>>(c->plan, G, xs, n, c->fft.p, cmul, ksin, cmax, tpt, \
                                                                              c->stage.p + c->lo, spec, (c->red_target ? c->red_target : c->d_red))
  if (spec.on) SK_LAUNCH_CELLS(true, 2); else SK_LAUNCH_CELLS(false, 2);
#undef SK_LAUNCH_CELLS
  return 0;
}
template <int W>
void launch_interp_cplx(sk_ctx *c, const SkGeom &G, const double *x, long long n, sk_cplx *out) {
  k_interp_cplx<W><<<nblk(n, 256), 256, 0, c->stream>>>(c->plan, G, x, n, c->fft.p, out);
}


// Source side of one transform pair: prep + spread/deconvolve/pad + FFT.  nrule = 1 or 2.
// re1/im1 (rule 0) and re2 (rule 1) are the strengths over the nodes c->no1 / c->no2; fft_out receives
// the interleaved grids [nf2][nrule].
int run_source_side(sk_ctx *c, const SkGeom &G, int nrule, long long M1, const double *re1, const double *im1,
                    long long M2, const double *re2, DevBuf<sk_cplx> &fft_out) {
  const size_t need = (size_t)G.nf2 * nrule;
  CK(fft_out.ensure(need));
  CK(c->pos_hi1.ensure(M1));
  CK(c->pos_lo1.ensure(M1));
  CK(c->cs1.ensure(M1));
  k_prep_sources<<<nblk(M1, 256), 256, 0, c->stream>>>(G, M1, c->no1.p, re1, im1, c->pos_hi1.p, c->pos_lo1.p, c->cs1.p);
  LAUNCH_CHECK();
  SkSpreadSrc src;
  std::memset(&src, 0, sizeof(src));
  src.pos_hi[0] = c->pos_hi1.p; src.pos_lo[0] = c->pos_lo1.p; src.cs[0] = c->cs1.p; src.M[0] = M1;
  if (nrule == 2) {
    CK(c->pos_hi2.ensure(M2));
    CK(c->pos_lo2.ensure(M2));
    CK(c->cs2.ensure(M2));
    k_prep_sources<<<nblk(M2, 256), 256, 0, c->stream>>>(G, M2, c->no2.p, re2, nullptr, c->pos_hi2.p, c->pos_lo2.p, c->cs2.p);
    LAUNCH_CHECK();
    src.pos_hi[1] = c->pos_hi2.p; src.pos_lo[1] = c->pos_lo2.p; src.cs[1] = c->cs2.p; src.M[1] = M2;
  }
  CK(cudaMemsetAsync(fft_out.p, 0, sizeof(sk_cplx) * need, c->stream));   // zero-padding of the modes
  dim3 grid(nblk(G.nf, SK_SPREAD_CELLS), nrule);          // SK_SPREAD_LANES lanes per spread-grid cell
#define CALL(WW) k_spread_modes<WW><<<grid, 256, 0, c->stream>>>(c->plan, G, src, nrule, fft_out.p)
  DISPATCH_W(c->plan.w, CALL)
#undef CALL
  LAUNCH_CHECK();
  cufftHandle h;
  int rc = get_fft_plan(c, G.nf2, nrule, &h);
  if (rc != SK_OK) return rc;
  cufftResult fr = cufftExecZ2Z(h, (cufftDoubleComplex *)fft_out.p, (cufftDoubleComplex *)fft_out.p, CUFFT_INVERSE);
  if (fr != CUFFT_SUCCESS) return fail(c, SK_ERR_CUFFT, "cufftExecZ2Z failed: %d", (int)fr);
  c->stats.kernel_launches++, c->launches_total++;
  c->stats.last_nf = G.nf;
  c->stats.last_nf2 = G.nf2;
  return SK_OK;
}

// O(N) nonuniform Hankel transform of the sub-interval's sources at the active targets (sk_hankel.h), staged
// like the other branches.  Returns SK_ERR_UNSUPPORTED (without touching the staging buffers) when the dyadic
// scheme does not apply; the caller then takes the direct Bessel summation.
// nu: Bessel order; sbuf1 / sbuf2: strengths over the nodes c->no1 / c->no2; raw != nullptr: write the two rule sums
// per target (raw[2j + rule].x) instead of staging.
int hankel_stage(sk_ctx *c, double a, double b, int nu, const double *sbuf1, const double *sbuf2, double cmul, double xdiv,
                 long long n_act, long long M1, long long M2, sk_cplx *raw) {
  NvtxRange nvtx("NUFHT call");
  SkHankelPlan H;
  SkHankelGroup *hg = c->h_scal->grp;
  const long long total = sk_hk_make_plan(c->plan, nu, a, b, c->r_lo, c->r_hi, &H, hg);
  if (total < 0 || c->plan.w != 16) return SK_ERR_UNSUPPORTED;
  if (!c->hk_tab_ready) {
    std::vector<double> tab(SK_HK_TAB_SIZE);
    if (sk_plan_bessel_table(SK_HK_NUMAX, SK_HK_TAB_INT, SK_HK_TAB_NC, tab.data()) != 0) return fail(c, SK_ERR_ARG, "Bessel table");
    CK(c->hk_tab.ensure(SK_HK_TAB_SIZE));
    CK(cudaMemcpyAsync(c->hk_tab.p, tab.data(), sizeof(double) * SK_HK_TAB_SIZE, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // (allocated per call below: the number of source slices depends on the rule size)
    CK(c->hk_cheb.ensure(2 * SK_HK_NLEV * SK_HK_NCH));
    CK(c->hk_loc.ensure((size_t)SK_HK_NLEV * SK_HK_NSUB * SK_HK_NLOC * 2));
    CK(c->hk_lev.ensure(2 * (SK_HK_NLEV + 1)));
    CK(c->hk_groups.ensure(SK_HK_NGRP));
    c->hk_tab_ready = true;
  }
  CK(c->hk_grid.ensure((size_t)std::max<long long>(total, 1)));
  CK(c->pos_hi1.ensure(M1)); CK(c->pos_lo1.ensure(M1)); CK(c->cs1.ensure(M1)); CK(c->hk_lam1.ensure(M1));
  CK(c->pos_hi2.ensure(M2)); CK(c->pos_lo2.ensure(M2)); CK(c->cs2.ensure(M2)); CK(c->hk_lam2.ensure(M2));
  if (H.ngroups > 0)
    CK(cudaMemcpyAsync(c->hk_groups.p, hg, sizeof(SkHankelGroup) * H.ngroups, cudaMemcpyHostToDevice, c->stream));
  // local part: level boundaries, node sums, Chebyshev coefficients
  k_hankel_levels<<<nblk(M1 + M2, 256), 256, 0, c->stream>>>(H.wT, c->no1.p, M1, c->no2.p, M2, c->hk_lev.p);
  LAUNCH_CHECK();
  {
    CK(c->hk_vals.ensure((size_t)SK_HK_FITSPLIT * 2 * SK_HK_NLEV * SK_HK_NCH));
    dim3 grid(SK_HK_NCH, H.q_hi - H.q_lo + 1, 2 * SK_HK_FITSPLIT);
    k_hankel_fit<<<grid, 256, 0, c->stream>>>(H, c->hk_tab.p, c->no1.p, sbuf1, c->no2.p, sbuf2, c->hk_lev.p, c->hk_vals.p);
    LAUNCH_CHECK();
    dim3 gc(H.q_hi - H.q_lo + 1, 2);
    k_hankel_cheb<<<gc, SK_HK_NCH, 0, c->stream>>>(H, c->hk_vals.p, c->hk_cheb.p);
    LAUNCH_CHECK();
    // the levels an octave needs, summed once into a piecewise expansion of the octave
    const int t_need = sk_hk_octave(c->r_hi, c->r_lo);
    dim3 gl(SK_HK_NSUB, (t_need < H.q_hi ? t_need : H.q_hi) + 1);
    k_hankel_local_poly<<<gl, 256, 0, c->stream>>>(H, c->hk_cheb.p, c->hk_loc.p);
    LAUNCH_CHECK();
  }
  // asymptotic part: one batched transform (K terms x 2 rules) per group
  SkHkSrc S;
  S.no[0] = c->no1.p; S.no[1] = c->no2.p; S.buf[0] = sbuf1; S.buf[1] = sbuf2;
  S.pos_hi[0] = c->pos_hi1.p; S.pos_hi[1] = c->pos_hi2.p; S.pos_lo[0] = c->pos_lo1.p; S.pos_lo[1] = c->pos_lo2.p;
  S.cs[0] = c->cs1.p; S.cs[1] = c->cs2.p; S.lam[0] = c->hk_lam1.p; S.lam[1] = c->hk_lam2.p;
  S.M[0] = M1; S.M[1] = M2;
  if (total > 0) CK(cudaMemsetAsync(c->hk_grid.p, 0, sizeof(sk_cplx) * (size_t)total, c->stream));   // zero padding
  // deepest octave first: the groups of a shared set add their own levels to a running mode buffer
  for (int gi = H.ngroups - 1; gi >= 0; --gi) {
    const SkGeom &G = hg[gi].G;
    const bool shared = hg[gi].shared != 0;
    sk_cplx *grid_g = c->hk_grid.p + hg[gi].grid_off;
    sk_cplx *dst = grid_g;
    if (shared) {
      CK(c->hk_modes.ensure((size_t)G.nf2 * 2 * SK_HK_K));
      if (hg[gi].shared == 1) CK(cudaMemsetAsync(c->hk_modes.p, 0, sizeof(sk_cplx) * (size_t)G.nf2 * 2 * SK_HK_K, c->stream));
      dst = c->hk_modes.p;
    }
    k_hankel_prep<<<nblk(M1 + M2, 256), 256, 0, c->stream>>>(c->hk_groups.p, gi, H.wT, S);
    LAUNCH_CHECK();
    // enough blocks to fill the GPU: small grids split each cell block's source range (split-K)
    const unsigned int bx = nblk(G.nf, SK_SPREAD_CELLS);
    int nsplit = (int)((148u * 4u + 2u * bx - 1u) / (2u * bx));
    nsplit = nsplit < 1 ? 1 : (nsplit > 64 ? 64 : nsplit);
    if (nsplit > 1) CK(c->hk_part.ensure((size_t)nsplit * G.nf * 2 * SK_HK_K));
    dim3 grid(bx, 2, nsplit);
    k_spread_hankel<16><<<grid, 256, 0, c->stream>>>(c->plan, c->hk_groups.p, gi, H, S, c->hk_lev.p, dst, shared ? 1 : 0,
                                                     c->hk_part.p);
    LAUNCH_CHECK();
    if (nsplit > 1) {
      k_spread_hankel_reduce<<<nblk(G.nf * 2 * SK_HK_K, 256), 256, 0, c->stream>>>(c->plan, c->hk_groups.p, gi, nsplit,
                                                                                  c->hk_part.p, dst, shared ? 1 : 0);
      LAUNCH_CHECK();
    }
    cufftHandle h;
    int rc = get_fft_plan(c, G.nf2, 2 * SK_HK_K, &h);
    if (rc != SK_OK) return rc;
    cufftResult fr = cufftExecZ2Z(h, (cufftDoubleComplex *)dst, (cufftDoubleComplex *)grid_g, CUFFT_INVERSE);   // shared: out of place
    if (fr != CUFFT_SUCCESS) return fail(c, SK_ERR_CUFFT, "cufftExecZ2Z failed: %d", (int)fr);
    c->stats.kernel_launches++, c->launches_total++;
    c->stats.last_nf = G.nf;
    c->stats.last_nf2 = G.nf2;
  }
  if (c->timing) CK(cudaEventRecord(c->ev[1], c->stream));
  if (c->interp_mode == 1)     // A/B: one target per thread, 16-byte loads (the plain restatement of sk_hk_point)
    k_hankel_interp<16><<<nblk(n_act, 256), 256, 0, c->stream>>>(c->plan, H, c->hk_groups.p, c->hk_grid.p, c->hk_loc.p,
                                                                 c->uxs.p + c->lo, n_act, cmul, xdiv,
                                                                 c->stage.p + c->lo, c->d_red, raw);
  else if (c->interp_mode == 2)   // A/B: two targets per thread, 256-bit loads; bit-identical to mode 1
    k_hankel_interp2<16><<<nblk((n_act + 1) / 2, SK_HK_TPB2), SK_HK_TPB2, 0, c->stream>>>(c->plan, H, c->hk_groups.p, c->hk_grid.p,
                                                                            c->hk_loc.p, c->uxs.p + c->lo, n_act, cmul,
                                                                            xdiv, c->stage.p + c->lo, c->d_red, raw);
  else                            // default: cell polynomials across the K terms
    k_hankel_cells<16><<<nblk(n_act, 256 * SK_HK_CT), 256, 0, c->stream>>>(c->plan, H, c->hk_groups.p, c->hk_grid.p, c->hk_loc.p,
                                                                          c->uxs.p + c->lo, n_act, cmul, xdiv,
                                                                          c->stage.p + c->lo, c->d_red, raw);
  LAUNCH_CHECK();
  if (c->timing) CK(cudaEventRecord(c->ev[2], c->stream));
  c->stats.n_hankel++;
  return SK_OK;
}

// ---- source-side prefetch (see sk_ctx::pf) -------------------------------------------------------------------
void swap_src_sets(sk_ctx *c) {
  std::swap(c->no1, c->pf.no1); std::swap(c->buf1, c->pf.buf1); std::swap(c->no2, c->pf.no2); std::swap(c->buf2, c->pf.buf2);
  std::swap(c->pos_hi1, c->pf.pos_hi1); std::swap(c->pos_lo1, c->pf.pos_lo1);
  std::swap(c->pos_hi2, c->pf.pos_hi2); std::swap(c->pos_lo2, c->pf.pos_lo2);
  std::swap(c->cs1, c->pf.cs1); std::swap(c->cs2, c->pf.cs2); std::swap(c->fft, c->pf.fft);
}

// updatequadbufs! (src/quadrature.jl:49-95) for a built-in density: the panel spec of the sub-interval [a, b]
void make_panel_spec(const sk_ctx *c, double a, double b, int logw, SkPanelSpec *out) {
  SkPanelSpec &S = *out;
  std::memset(&S, 0, sizeof(S));
  const bool origin = (a == 0.0 && c->p != 0.0);                    // src/quadrature.jl:185
  S.m = c->m; S.k = c->k;
  S.origin_jacobi = origin ? 1 : 0;
  S.weight_in_f = origin ? 0 : 1;                                   // :230-238 vs :240-247
  S.logw = logw ? 1 : 0;
  S.family = c->family; S.deriv = c->deriv; S.nparam = c->nparam;
  S.p = c->p;
  for (int i = 0; i < c->nparam; ++i) S.params[i] = c->params[i];
  sk_fill_subpanels(a, b, c->k, S.bmad2, S.bpad2);
  S.jac_scale = std::pow(S.bmad2[0], c->p + 1);
}

int launch_gen_sources(sk_ctx *c, const SkPanelSpec &S) {
  NvtxRange nvtx("update quadrature buffers");
  const long long M1 = (long long)c->m * c->k;
  CK(c->no1.ensure(M1)); CK(c->buf1.ensure(M1)); CK(c->no2.ensure(2 * M1)); CK(c->buf2.ensure(2 * M1));
  k_gen_sources<<<nblk(3 * M1, 256), 256, 0, c->stream>>>(S, c->leg_no1.p, c->leg_wt1.p, c->leg_no2.p, c->leg_wt2.p,
                                                          c->jac_no1.p, c->jac_wt1.p, c->jac_no2.p, c->jac_wt2.p,
                                                          c->no1.p, c->buf1.p, c->no2.p, c->buf2.p);
  LAUNCH_CHECK();
  return SK_OK;
}

// nodes, strengths, spread and FFT of (S, G) on the prefetch stream, into the second buffer set
int prefetch_sources(sk_ctx *c, const SkPanelSpec &S, const SkGeom &G) {
  c->pf_valid = false;
  swap_src_sets(c);
  c->stream = c->stream2;
  const long long M1 = (long long)c->m * c->k;
  int rc = launch_gen_sources(c, S);
  if (rc == SK_OK) rc = run_source_side(c, G, 2, M1, c->buf1.p, nullptr, 2 * M1, c->buf2.p, c->fft);
  if (rc == SK_OK && cudaEventRecord(c->pf_ev, c->stream2) != cudaSuccess) rc = SK_ERR_CUDA;
  c->stream = c->stream_main;
  swap_src_sets(c);
  if (rc != SK_OK) return rc;
  c->pf_S = S;
  c->pf_G = G;
  c->pf_valid = true;
  c->n_pf_issued++;
  return SK_OK;
}

// the same into the SECOND prefetch set: the request after the next one (both panels of a typical run are known as soon
// as the distance range is: (0, b1) and (b1, b1 + m k / (2 r_hi)), src/adaptive.jl:152)
int prefetch_sources_second(sk_ctx *c, const SkPanelSpec &S, const SkGeom &G) {
  const bool v1 = c->pf_valid;
  const SkPanelSpec S1 = c->pf_S;
  const SkGeom G1 = c->pf_G;
  std::swap(c->pf, c->pf2);
  std::swap(c->pf_ev, c->pf2_ev);
  const int rc = prefetch_sources(c, S, G);
  c->pf2_valid = c->pf_valid;
  c->pf2_S = c->pf_S;
  c->pf2_G = c->pf_G;
  std::swap(c->pf, c->pf2);
  std::swap(c->pf_ev, c->pf2_ev);
  c->pf_valid = v1;
  c->pf_S = S1;
  c->pf_G = G1;
  return rc;
}
// the first set was just taken (its buffers are the current set now, the old current set sits in pf): the second set moves up
void promote_second_prefetch(sk_ctx *c) {
  c->pf_valid = false;
  if (!c->pf2_valid) return;
  std::swap(c->pf, c->pf2);
  std::swap(c->pf_ev, c->pf2_ev);
  c->pf_S = c->pf2_S;
  c->pf_G = c->pf2_G;
  c->pf_valid = true;
  c->pf2_valid = false;
}

// the device-side form of the scan arguments a speculated sub-interval carries (SkSpec; `fresh` and the chain guard are
// set by the caller)
void fill_spec(sk_ctx *c, const sk_scan_args *sa, SkSpec &spec) {
  spec.on = 1;
  spec.criteria = sa->criteria;
  spec.lo0 = c->lo;
  spec.trunc_a = sa->trunc_a;
  spec.trunc_num = sa->trunc_num;
  spec.xpow = sa->xpow;
  spec.tau = sa->tau;
  // exact threshold distance of the truncation half of the predicate (SkSpec::xstar), dim = 1 (x^1: no pow())
  spec.use_xstar = 0;
  spec.xstar = 0.0;
  if (spec.criteria != 0 && spec.xpow == 1.0 && spec.trunc_num > 0.0 && std::isfinite(spec.trunc_num) &&
      spec.trunc_a == spec.trunc_a && spec.tau == spec.tau) {
    auto pred = [&](double x) { return sk_trunc_err(spec.trunc_a, spec.trunc_num, 1.0, x, 0) < spec.tau; };
    const double dmax = 1.7976931348623157e308, dmin = 4.9406564584124654e-324;
    if (pred(dmin)) { spec.use_xstar = 1; spec.xstar = 0.0; }                 // every positive distance passes
    else if (!pred(dmax)) { spec.use_xstar = 1; spec.xstar = INFINITY; }      // none does
    else {
      unsigned long long lo_b = 1ull, hi_b = 0x7fefffffffffffffull;          // pred(lo) false, pred(hi) true
      while (hi_b - lo_b > 1ull) {
        const unsigned long long mid = lo_b + (hi_b - lo_b) / 2;
        double xm;
        std::memcpy(&xm, &mid, sizeof(double));
        if (pred(xm)) hi_b = mid; else lo_b = mid;
      }
      std::memcpy(&spec.xstar, &hi_b, sizeof(double));
      spec.use_xstar = 1;
    }
  }
  spec.res = c->res.p + c->lo;
  spec.backup = c->stage.p + c->lo;
}

// transform + stage for the sub-interval whose sources are in no1/buf1/no2/buf2, in two halves: everything that is
// enqueued on the context's stream, and the read-back of the reduced scalars after the stream has drained.  One
// context runs the halves back to back (transform_and_stage); a device group (sk_group_*) enqueues on every device
// first and reads back afterwards, so that the devices work concurrently under a single host thread.
int transform_and_stage_enqueue_local(sk_ctx *c, double a, double b, const sk_subinterval_opts *o) {
  const long long n_act = c->hi - c->lo;
  const long long M1 = (long long)c->m * c->k, M2 = 2 * M1;
  const int ksin = o->kernel == SK_KERNEL_SIN;
  // fast = nufft_quad_size_cutoff(length(no2), length(xs)) && length(xs) > 1   (src/quadrature.jl:105, src/utils.jl:39)
  const long long n_cut = c->n_act_global > 0 ? c->n_act_global : n_act;
  const bool bessel = o->kernel == SK_KERNEL_BESSEL;
  const bool fast = !bessel && (M2 * n_cut > (1LL << 18)) && n_cut > 1;
  // speculation is only meaningful for the first sub-interval of the panel (the whole panel)
  const bool spec_on = fast && o->speculate != nullptr && c->panel_subs == 0 && o->speculate->criteria >= 0 &&
                       o->speculate->criteria <= 2;
  SkSpec spec;
  std::memset(&spec, 0, sizeof(spec));
  c->spec_fresh = false;
  if (spec_on) {
    // first panel of a run over every positive distance: the table is still (pending) zero -> written outright
    const bool fresh = c->res_zero_pending && !c->commit_pending && c->lo == (c->has_zero ? 1 : 0) && c->hi == c->n_unique;
    if (fresh) {
      c->res_zero_pending = false;
      if (c->has_zero && !c->zero_lag_written) CK(cudaMemsetAsync(c->res.p, 0, sizeof(sk_cplx), c->stream));
      c->spec_fresh = true;
      spec.fresh = 1;
    } else {
      int rcz = ensure_res_zero(c);      // the table is read: it must hold its zeros now
      if (rcz != SK_OK) return rcz;
    }
    int rc = flush_commit(c);            // res must be current before it is updated in place
    if (rc != SK_OK) return rc;
    fill_spec(c, o->speculate, spec);
  }
  {
    int rci = red_init(c, c->d_red, c->lo - 1);
    if (rci != SK_OK) return rci;
  }
  if (c->timing) CK(cudaEventRecord(c->ev[0], c->stream));
  SkGeom G;
  std::memset(&G, 0, sizeof(G));
  if (fast && sk_make_geom(c->plan, a, b, c->r_lo, c->r_hi, &G) != 0)
    return fail(c, SK_ERR_ARG, "type-3 grid too large for [a,b]=[%g,%g], r in [%g,%g]", a, b, c->r_lo, c->r_hi);
  // built-in density: the sources either wait in the prefetch set (same panel spec, same geometry, bit for bit) or
  // are generated now
  const bool builtin_req = c->need_gen;
  bool served = false;
  if (c->need_gen) {
    c->need_gen = false;
    if (fast && c->pf_valid && std::memcmp(&c->pf_S, &c->pend_S, sizeof(SkPanelSpec)) == 0 &&
        std::memcmp(&c->pf_G, &G, sizeof(SkGeom)) == 0) {
      CK(cudaStreamWaitEvent(c->stream, c->pf_ev, 0));
      swap_src_sets(c);
      served = true;
      c->n_pf_hits++;
      c->stats.last_nf = G.nf;
      c->stats.last_nf2 = G.nf2;
      promote_second_prefetch(c);
    } else {
      int rc = launch_gen_sources(c, c->pend_S);
      if (rc != SK_OK) return rc;
    }
    c->have_sources = true;
  }
  if (!served) c->pf_valid = c->pf2_valid = false;   // a prefetch is good for the very request it was made for
  // dim >= 2: the reference calls nufht whenever its NUFFT cutoff holds (src/quadrature.jl:139-143); here the
  // O(N) scheme is taken when it is cheaper than the direct Bessel summation (more than ~4096 active targets)
  int hk_rc = SK_ERR_UNSUPPORTED;
  bool hk_timed = false;
  if (bessel && c->hankel_mode != 1 && n_cut > 1 && c->r_lo > 0.0 &&
      (c->hankel_mode == 2 || (M2 * n_cut > (1LL << 29)))) {
    hk_rc = hankel_stage(c, a, b, o->nu, c->buf1.p, c->buf2.p, o->cmul, o->xdiv_pow, n_act, M1, M2, nullptr);
    if (hk_rc != SK_OK && hk_rc != SK_ERR_UNSUPPORTED) return hk_rc;
    hk_timed = hk_rc == SK_OK;
  }
  if (fast) {
    NvtxRange nvtx("FINUFFT call");
    if (!served) {
      int rc = run_source_side(c, G, 2, M1, c->buf1.p, nullptr, M2, c->buf2.p, c->fft);
      if (rc != SK_OK) return rc;
    }
    if (c->timing) CK(cudaEventRecord(c->ev[1], c->stream));
#define CALL(WW) launch_interp_session<WW>(c, G, c->uxs.p + c->lo, n_act, o->cmul, ksin, spec)
    DISPATCH_W(c->plan.w, CALL)
#undef CALL
    LAUNCH_CHECK();
    if (c->timing) CK(cudaEventRecord(c->ev[2], c->stream));
    c->stats.n_fast++;
    // the first panel of a run rarely converges anything: the next panel is then (b, b + quadm / (2 r_hi)) over the same
    // targets (src/adaptive.jl:152).  Its source side starts now, next to the interpolation of this one.
    if (c->prefetch_on && builtin_req && spec_on && c->stats.n_panels == 0 && c->r_hi > 0.0) {
      const double a2 = b, b2 = b + (double)((long long)c->m * c->k) / (2 * c->r_hi);
      SkGeom G2;
      if (std::isfinite(b2) && b2 > a2 && sk_make_geom(c->plan, a2, b2, c->r_lo, c->r_hi, &G2) == 0) {
        SkPanelSpec S2;
        make_panel_spec(c, a2, b2, o->logw, &S2);
        const bool waiting = c->pf_valid && std::memcmp(&c->pf_S, &S2, sizeof(SkPanelSpec)) == 0 &&
                             std::memcmp(&c->pf_G, &G2, sizeof(SkGeom)) == 0;      // (prefetched next to the sort already)
        if (!waiting) {
          int rc = prefetch_sources(c, S2, G2);
          if (rc != SK_OK) return rc;
        }
      }
    }
  } else if (bessel && hk_rc == SK_OK) {
    // staged by hankel_stage above
  } else {
    NvtxRange nvtx(bessel ? "direct Bessel summation" : "direct Fourier summation");
    if (n_act > 2000000000LL) return fail(c, SK_ERR_ARG, "too many targets for the direct branch");
    CK(c->dsum.ensure((size_t)n_act * 2));
    dim3 grid((unsigned int)n_act, 2);
    if (bessel)
      k_direct_bessel<<<grid, 256, 0, c->stream>>>(o->nu, c->no1.p, c->buf1.p, M1, c->no2.p, c->buf2.p, M2, c->uxs.p + c->lo, c->dsum.p);
    else
      k_direct<<<grid, 256, 0, c->stream>>>(c->no1.p, c->buf1.p, M1, c->no2.p, c->buf2.p, M2, c->uxs.p + c->lo, c->dsum.p);
    LAUNCH_CHECK();
    k_direct_finish<<<nblk(n_act, 256), 256, 0, c->stream>>>(c->dsum.p, c->uxs.p + c->lo, n_act, o->cmul, ksin,
                                                            bessel ? o->xdiv_pow : 0.0, c->stage.p + c->lo, c->d_red);
    LAUNCH_CHECK();
    c->stats.n_direct++;
  }
  {
    int rcp = publish(c, &c->h_scal->red, c->d_red, sizeof(SkReduceOut));
    if (rcp != SK_OK) return rcp;
  }
  CK(cudaEventRecord(c->ev_red, c->stream));
  c->pend_timed = c->timing && (fast || hk_timed);
  c->pend_spec = spec_on;
  if (spec_on) c->spec_args = *o->speculate;
  return SK_OK;
}

int transform_and_stage_enqueue(sk_ctx *c, double a, double b, const sk_subinterval_opts *o) {
  NvtxRange nvtx("panel integral");
  c->pend_timed = c->pend_spec = false;
  {
    // does this sub-interval speculate (first of its panel, NUFFT branch, scan arguments given)?  Decided from the
    // arguments and the GLOBAL active count only, so that every rank -- also an idle one -- issues the same collectives
    const long long M2 = 2LL * c->m * c->k, n_cut = c->n_act_global > 0 ? c->n_act_global : (c->hi - c->lo);
    c->pend_ab = sharded(c) && o->kernel != SK_KERNEL_BESSEL && (M2 * n_cut > (1LL << 18)) && n_cut > 1 &&
                 o->speculate != nullptr && c->panel_subs == 0 && o->speculate->criteria >= 0 && o->speculate->criteria <= 2;
  }
  const int rc = transform_and_stage_enqueue_local(c, a, b, o);
  c->pend_rc = rc;
  if (!sharded(c)) return rc;
  // sharded run: a rank that failed locally still joins the collective (its peers are waiting in it)
  const std::string msg = c->errmsg;
  c->peer_slot = 0;
  c->cur_red = c->d_red;
  const int rcc = c->pend_ab ? comm_reduce_ab(c, 0, rc != SK_OK) : comm_reduce_a(c, 0, rc != SK_OK);
  if (rc != SK_OK) c->errmsg = msg;
  if (rcc == SK_OK) cudaEventRecord(c->ev_red, c->stream);      // (behind the exchange: _finish may wait on the event only)
  return rcc != SK_OK ? rcc : SK_OK;                     // a local failure is reported by _finish, after the collective
}

// chained launches in a sharded run need the mailboxes (their guards read the global scalars the exchange kernels leave
// in device memory, and a skipped launch turns its exchange void); with NCCL all-reduces they stay off
inline bool chain_allowed(const sk_ctx *c) {
  return !sharded(c) || (c->peer_n > 0 && c->comm == nullptr && c->sharded_chain);
}

// sharded run over mailboxes: while the exchange just read back is void (some rank's chained launch skipped itself),
// issue it again with this rank's local scalars; every rank does the same, so the exchanges stay paired
int peer_resend_while_void(sk_ctx *c) {
  int guard = 0;
  while (peer_void(c)) {
    if (++guard > 8) return fail(c, SK_ERR_STATE, "peer exchange: too many void rounds");
    c->peer_slot = 0;
    const long long lo = c->lo;
    int rc = peer_exchange(c, 0, c->pend_ab ? SK_PX_AB : SK_PX_A, 0, c->pend_rc != SK_OK, lo, nullptr, 0, 0, nullptr, nullptr,
                           c->cur_red);
    if (rc != SK_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
    rc = peer_check(c, 0);
    if (rc != SK_OK) return rc;
  }
  return SK_OK;
}

// flags_out != nullptr: hand the NaN flags to the caller (a device group applies the rule of src/quadrature.jl:165
// to the flags of all devices) instead of raising SK_ERR_NAN here
int transform_and_stage_finish(sk_ctx *c, double *max_abs_diff, unsigned int *flags_out) {
  const long long n_act = c->hi - c->lo;
  if (c->chain.pending && c->chain.adopted) {
    // this sub-interval was enqueued ahead of time (sk_subinterval_chain): its scalars are in the chain's slot
    CK(cudaEventSynchronize(c->ev_red2));
    c->chain.pending = c->chain.adopted = false;
    if (c->h_red2->flags & SK_FLAG_SKIPPED) {
      // the guard did not hold after all (cannot happen when the caller follows src/adaptive.jl:149-200, kept for
      // safety): nothing was touched, evaluate the sub-interval now
      sk_subinterval_opts o = c->chain.o;
      o.speculate = &c->chain.sa;
      make_panel_spec(c, c->chain.a, c->chain.b, o.logw, &c->pend_S);
      c->need_gen = true;
      int rc = sharded(c) ? transform_and_stage_enqueue(c, c->chain.a, c->chain.b, &o)      // (with its exchange)
                          : transform_and_stage_enqueue_local(c, c->chain.a, c->chain.b, &o);
      if (rc != SK_OK) return rc;
      CK(cudaStreamSynchronize(c->stream));
    } else {
      c->h_scal->red = *c->h_red2;
      c->stats.n_chained++;
      if (sharded(c)) { c->peer_slot = 4; c->cur_red = c->d_red2; }
    }
  } else if (c->chain.pending) {
    CK(cudaEventSynchronize(c->ev_red));      // a chained launch is queued behind this sub-interval: do not wait for it
  } else {
    CK(cudaStreamSynchronize(c->stream));
  }
  if (c->early1.pending && c->early1.adopted) {
    c->early1.pending = c->early1.adopted = false;
    if (c->h_scal->red.flags & SK_FLAG_SKIPPED) {
      // the launch behind the sort skipped itself (the sort fell back to the general path, or the active set is too
      // small for the NUFFT branch): evaluate the sub-interval now, the ordinary way
      int rc = chain_discard(c);              // (a launch chained behind it saw the flag and skipped itself too)
      if (rc != SK_OK) return rc;
      sk_subinterval_opts o = c->early1.o;
      o.speculate = &c->early1.sa;
      c->res_zero_pending = true;
      c->spec_fresh = false;
      make_panel_spec(c, c->early1.a, c->early1.b, o.logw, &c->pend_S);
      c->need_gen = true;
      rc = sharded(c) ? transform_and_stage_enqueue(c, c->early1.a, c->early1.b, &o)
                      : transform_and_stage_enqueue_local(c, c->early1.a, c->early1.b, &o);
      if (rc != SK_OK) return rc;
      CK(cudaStreamSynchronize(c->stream));
    } else {
      c->stats.n_chained++;
    }
  }
  if (sharded(c)) {
    if (c->pend_rc != SK_OK) return c->pend_rc;                     // this rank's own failure (message already set)
    int prc = peer_check(c, c->peer_slot);
    if (prc != SK_OK) return prc;
    prc = peer_resend_while_void(c);                                // (some other rank's chained launch skipped itself)
    if (prc != SK_OK) return prc;
    if (global_a(c).err) return fail(c, SK_ERR_STATE, "another rank failed in this sub-interval");
  }
  if (c->pend_timed) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.source_ms += ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
    c->stats.interp_ms += ms;
  }
  unsigned int fl = c->h_scal->red.flags;
  double mx;
  std::memcpy(&mx, &c->h_scal->red.maxbits, sizeof(double));
  comm_take_a(c, &fl, &mx);                 // sharded run: the maximum and the NaN flags over all ranks
  if (sharded(c) && c->pend_ab) comm_take_b(c);
  if (fl & SK_FLAG_NAND) mx = std::nan("");
  *max_abs_diff = mx;
  c->g_max_abs = mx;
  c->staged = true;
  c->panel_subs++;
  c->stats.n_subintervals++;
  c->stats.units += n_act;
  if (c->pend_spec) {
    c->spec_active = true;
    c->spec_accepted = false;
    const long long top = c->h_scal->red.max_unconv;
    c->spec_new_hi = top + 1;
    c->spec_r = 0.0;
    if (top >= c->lo) std::memcpy(&c->spec_r, &c->h_scal->red.rbits, sizeof(double));
    c->stats.n_speculated++;
  }
  if (flags_out) { *flags_out = fl; return SK_OK; }
  // any(isnan, int1) || any(isnan, int2) && throw(...)   (src/quadrature.jl:165)
  if (!(fl & SK_FLAG_NAN1) && (fl & SK_FLAG_NAN2)) return fail(c, SK_ERR_NAN, "NaN detected in panel integral...");
  return SK_OK;
}

int transform_and_stage(sk_ctx *c, double a, double b, const sk_subinterval_opts *o, double *max_abs_diff) {
  int rc = transform_and_stage_enqueue(c, a, b, o);
  if (rc != SK_OK) return rc;
  return transform_and_stage_finish(c, max_abs_diff, nullptr);
}

// a speculative commit that was not accepted is rolled back before anything else touches the panel
int rollback_speculation(sk_ctx *c) {
  if (c->spec_active && !c->spec_accepted) {
    const long long n = c->hi - c->lo;
    if (c->spec_fresh) {                  // the table was zero before the sub-interval: rolling back = zeroing
      CK(cudaMemsetAsync(c->res.p + c->lo, 0, sizeof(sk_cplx) * n, c->stream));
    } else {
      k_restore<<<nblk(n, 256), 256, 0, c->stream>>>(c->res.p + c->lo, c->stage.p + c->lo, n);
      LAUNCH_CHECK();
    }
    c->spec_fresh = false;
    c->spec_active = false;
    c->staged = false;
    c->stats.n_spec_rollbacks++;
  }
  return SK_OK;
}

// a chained launch the host did not pick up: if it ran (its guard held) its speculative commit is rolled back
int chain_discard(sk_ctx *c) {
  c->gchain.pending = false;                 // (a gather chained behind it wrote, at most, into the caller's output arrays)
  if (!c->chain.pending) return SK_OK;
  c->chain.pending = c->chain.adopted = false;
  CK(cudaEventSynchronize(c->ev_red2));
  if (!(c->h_red2->flags & SK_FLAG_SKIPPED)) {
    const long long n = c->chain.hi - c->chain.lo;
    k_restore<<<nblk(n, 256), 256, 0, c->stream>>>(c->res.p + c->chain.lo, c->stage.p + c->chain.lo, n);
    LAUNCH_CHECK();
  }
  return SK_OK;
}

// the first-panel launch queued behind the sort that the host did not pick up: if it ran, the table holds its commit
int early_discard(sk_ctx *c) {
  if (!c->early1.pending) return SK_OK;
  c->early1.pending = c->early1.adopted = false;
  CK(cudaEventSynchronize(c->ev_red));
  if (!(c->h_scal->red.flags & SK_FLAG_SKIPPED) && c->n_unique > c->early1.lo)
    CK(cudaMemsetAsync(c->res.p + c->early1.lo, 0, sizeof(sk_cplx) * (c->n_unique - c->early1.lo), c->stream));
  return SK_OK;
}
bool early_matches(const sk_ctx *c, double a, double b, const sk_subinterval_opts *o) {
  const sk_ctx::Early &h = c->early1;
  return h.pending && !h.adopted && a == h.a && b == h.b && c->lo == h.lo && c->hi == c->n_unique && c->r_lo == h.r_lo &&
         c->r_hi == h.r_hi && c->panel_subs == 0 && !c->timing && c->res_zero_pending && !c->commit_pending &&
         o->cmul == h.o.cmul && o->p == h.o.p && o->kernel == h.o.kernel && o->logw == h.o.logw && o->nu == h.o.nu &&
         o->xdiv_pow == h.o.xdiv_pow && o->speculate != nullptr && std::memcmp(o->speculate, &h.sa, sizeof(sk_scan_args)) == 0;
}

bool chain_matches(const sk_ctx *c, double a, double b, const sk_subinterval_opts *o) {
  const sk_ctx::Chain &h = c->chain;
  return h.pending && !h.adopted && a == h.a && b == h.b && c->lo == h.lo && c->hi == h.hi && c->r_lo == h.r_lo &&
         c->r_hi == h.r_hi && c->panel_subs == 0 && !c->timing && o->cmul == h.o.cmul && o->p == h.o.p &&
         o->kernel == h.o.kernel && o->logw == h.o.logw && o->nu == h.o.nu && o->xdiv_pow == h.o.xdiv_pow &&
         o->speculate != nullptr && std::memcmp(o->speculate, &h.sa, sizeof(sk_scan_args)) == 0;
}

// Rules are generated on the device (k_gauss_rules: Newton in double, then double-double) and kept for the life of
// the context; the host long-double generator (sk_plan_gauss_rule) remains as the cross-check in the tests.  All
// rules a configuration needs that are not cached yet -- up to four: (m, 2m) x (Legendre, Jacobi) -- are generated by
// ONE launch (blockIdx.y = rule): their Newton iterations are long serial recurrences, so they cost the time of the
// largest one, not the sum.
typedef std::map<std::pair<int, double>, std::pair<std::vector<double>, std::vector<double>>> RuleCache;
RuleCache &rule_cache() {          // process-wide: every context of a device group, every Session of a fit shares it
  static RuleCache cache;
  return cache;
}
std::mutex &rule_cache_mutex() {
  static std::mutex mu;
  return mu;
}

int ensure_gauss_rules(sk_ctx *c, const std::vector<std::pair<int, double>> &want) {
  std::lock_guard<std::mutex> lock(rule_cache_mutex());
  std::vector<std::pair<int, double>> todo;
  for (const auto &k : want)
    if (rule_cache().find(k) == rule_cache().end() && std::find(todo.begin(), todo.end(), k) == todo.end()) todo.push_back(k);
  if (todo.empty()) return 0;
  const int nj = (int)todo.size();
  size_t coef_doubles = 0, out_doubles = 0;
  int nmax = 0;
  for (const auto &k : todo) { coef_doubles += 6 * (size_t)k.first; out_doubles += 2 * (size_t)k.first; nmax = std::max(nmax, k.first); }
  std::vector<double> hcoef(coef_doubles), hout(out_doubles);
  std::vector<SkRuleJob> jobs(nj);
  double *dcoef = nullptr, *dout = nullptr;
  SkRuleJob *djobs = nullptr;
  bool ok = cudaMalloc((void **)&dcoef, sizeof(double) * coef_doubles) == cudaSuccess &&
            cudaMalloc((void **)&dout, sizeof(double) * out_doubles) == cudaSuccess &&
            cudaMalloc((void **)&djobs, sizeof(SkRuleJob) * nj) == cudaSuccess;
  int rc = 0;
  if (ok) {
    size_t co = 0, oo = 0;
    for (int j = 0; j < nj && rc == 0; ++j) {
      const int n = todo[j].first;
      if (sk_plan_jacobi_coeffs(n, todo[j].second, &hcoef[co], &hcoef[co + 2 * (size_t)n], &hcoef[co + 4 * (size_t)n]) != 0) rc = -1;
      jobs[j].n = n;
      jobs[j].p = todo[j].second;
      jobs[j].A = (const sk_dd *)(dcoef + co);
      jobs[j].B = (const sk_dd *)(dcoef + co + 2 * (size_t)n);
      jobs[j].C = (const sk_dd *)(dcoef + co + 4 * (size_t)n);
      jobs[j].no = dout + oo;
      jobs[j].wt = dout + oo + n;
      co += 6 * (size_t)n;
      oo += 2 * (size_t)n;
    }
    if (rc == 0) {
      ok = cudaMemcpyAsync(dcoef, hcoef.data(), sizeof(double) * coef_doubles, cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
           cudaMemcpyAsync(djobs, jobs.data(), sizeof(SkRuleJob) * nj, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
      if (ok) {
        dim3 grid((nmax + 63) / 64, nj);
        k_gauss_rules<<<grid, 64, 0, c->stream>>>(djobs, nj);
        c->stats.kernel_launches++, c->launches_total++;
        ok = cudaGetLastError() == cudaSuccess &&
             cudaMemcpyAsync(hout.data(), dout, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
             cudaStreamSynchronize(c->stream) == cudaSuccess;
      }
    }
  }
  if (dcoef) cudaFree(dcoef);
  if (dout) cudaFree(dout);
  if (djobs) cudaFree(djobs);
  if (rc != 0) return rc;
  if (!ok) return -2;
  size_t oo = 0;
  for (int j = 0; j < nj; ++j) {
    const int n = todo[j].first;
    std::vector<double> x(hout.begin() + oo, hout.begin() + oo + n), w(hout.begin() + oo + n, hout.begin() + oo + 2 * (size_t)n);
    oo += 2 * (size_t)n;
    for (int i = 1; i < n; ++i)
      if (!(x[i] > x[i - 1])) return -3;                 // Newton landed on a neighbouring zero
    rule_cache().emplace(todo[j], std::make_pair(std::move(x), std::move(w)));
  }
  return 0;
}

int cached_gauss_rule(sk_ctx *c, int n, double p, std::vector<double> &no, std::vector<double> &wt) {
  if (ensure_gauss_rules(c, {{n, p}}) != 0) return -1;
  std::lock_guard<std::mutex> lock(rule_cache_mutex());
  const auto it = rule_cache().find(std::make_pair(n, p));
  no = it->second.first;
  wt = it->second.second;
  return 0;
}

int upload_rule(sk_ctx *c, DevBuf<double> &dst, const double *src, int n) {
  CK(dst.ensure(n));
  CK(cudaMemcpyAsync(dst.p, src, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  return SK_OK;
}

// General sort for inputs the bin scheme of K8 cannot take (heavily clustered or duplicated distances overflow a fine
// bin): full radix sort of the keys, first-of-value flags, scan, scatter.  Rare path.
int targets_sort_general(sk_ctx *c, long long n_in) {
  CK(c->keys.ensure(n_in));
  CK(c->keys_alt.ensure(n_in));
  CK(c->idx.ensure(n_in));
  CK(c->idx_alt.ensure(n_in));
  CK(c->head.ensure(n_in));
  CK(c->uid.ensure(n_in));
  SkKeyBits kb0;
  kb0.bits_or = 0ull; kb0.bits_and = ~0ull; kb0.bad = 0; kb0.overflow = 0; kb0.unsorted = 0; kb0._pad = 0;
  c->h_scal->kb = kb0;
  CK(cudaMemcpyAsync(c->d_kb, &c->h_scal->kb, sizeof(SkKeyBits), cudaMemcpyHostToDevice, c->stream));
  k_make_keys<<<std::min<unsigned int>(nblk(n_in, 256), 148u * 16u), 256, 0, c->stream>>>(c->in.p, n_in, c->keys.p, c->idx.p, c->d_kb);
  LAUNCH_CHECK();
  cub::DoubleBuffer<unsigned long long> dk(c->keys.p, c->keys_alt.p);
  cub::DoubleBuffer<unsigned int> dv(c->idx.p, c->idx_alt.p);
  size_t tmp_bytes = 0, tmp2 = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n_in, 0, 64, c->stream));
  CK(cub::DeviceScan::InclusiveSum(nullptr, tmp2, c->head.p, c->uid.p, (int)n_in, c->stream));
  CK(c->cub_tmp.ensure(std::max(tmp_bytes, tmp2)));
  tmp_bytes = tmp2 = c->cub_tmp.cap;
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp_bytes, dk, dv, (int)n_in, 0, 63, c->stream));
  c->stats.kernel_launches += 9;
  k_flag_heads<<<nblk(n_in, 256), 256, 0, c->stream>>>(dk.Current(), n_in, c->head.p);
  LAUNCH_CHECK();
  CK(cub::DeviceScan::InclusiveSum(c->cub_tmp.p, tmp2, c->head.p, c->uid.p, (int)n_in, c->stream));
  c->stats.kernel_launches += 2;
  k_scatter_unique<<<nblk(n_in, 256), 256, 0, c->stream>>>(dk.Current(), dv.Current(), c->head.p, c->uid.p, n_in, c->uxs.p, c->inv.p);
  LAUNCH_CHECK();
  k_target_summary<<<1, 1, 0, c->stream>>>(c->uxs.p, c->uid.p, n_in, c->d_kb, c->d_sum);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(&c->h_scal->sum, c->d_sum, sizeof(SkTargetSummary), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return SK_OK;
}

// the key range of the distances being sorted, available as soon as k_k8_stats has run (r_hi = 0: no usable range)
int targets_early_range(sk_ctx *c, double *r_lo, double *r_hi) {
  *r_lo = *r_hi = 0.0;
  if (!c->early_pending) return SK_OK;
  c->early_pending = false;
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(c->k8_ev));
  unsigned long long kmin_inv, kmax, bad;
  if (c->early_global) {
    const unsigned long long *w = c->peer_n > 0 ? c->peer_out[1]->words
                                                : reinterpret_cast<const unsigned long long *>(c->h_scal->hv);
    kmin_inv = w[0]; kmax = w[1]; bad = w[2];
  } else {
    const SkK8State &k8 = c->h_scal->k8;
    kmin_inv = k8.kmin_inv; kmax = k8.kmax; bad = k8.bad;
  }
  if (bad || kmin_inv == 0ull || kmax == 0ull) return SK_OK;
  const unsigned long long kmin = ~kmin_inv;
  std::memcpy(r_lo, &kmin, sizeof(double));
  std::memcpy(r_hi, &kmax, sizeof(double));
  return SK_OK;
}
// source side of the first panel (0, m k / (2 r_hi)) (src/adaptive.jl:152) over the distance range [r_lo, r_hi], on the
// prefetch stream, while the sort is still running
int targets_early_prefetch(sk_ctx *c, double r_lo, double r_hi) {
  CK(cudaSetDevice(c->device));
  const double b1 = 0.0 + (double)((long long)c->m * c->k) / (2 * r_hi);
  SkGeom G1;
  if (!(std::isfinite(b1) && b1 > 0.0) || sk_make_geom(c->plan, 0.0, b1, r_lo, r_hi, &G1) != 0) return SK_OK;
  SkPanelSpec S1;
  make_panel_spec(c, 0.0, b1, c->last_logw, &S1);
  int rc = prefetch_sources(c, S1, G1);
  if (rc != SK_OK) return rc;
  // ... and of the second panel (b1, b1 + m k / (2 r_hi)): the first panel of a run rarely converges anything
  const double b2 = b1 + (double)((long long)c->m * c->k) / (2 * r_hi);
  SkGeom G2;
  if (std::isfinite(b2) && b2 > b1 && sk_make_geom(c->plan, b1, b2, r_lo, r_hi, &G2) == 0) {
    SkPanelSpec S2;
    make_panel_spec(c, b1, b2, c->last_logw, &S2);
    rc = prefetch_sources_second(c, S2, G2);
  }
  return rc;
}

// K8 (sk_k8.cuh): c->in holds the n_in raw distances -> sorted unique table c->uxs, inverse map c->inv.  One host
// synchronisation, at the end (the summary); an already strictly increasing input is detected on the device and
// costs one read pass.  Two halves like transform_and_stage (a device group sorts all chunks concurrently).
int targets_enqueue(sk_ctx *c, long long n_in) {
  NvtxRange nvtx("unique / sort (K8)");
  {
    int rcd = chain_discard(c);
    if (rcd == SK_OK) rcd = early_discard(c);
    if (rcd != SK_OK) return rcd;
  }
  c->have_targets = false;
  c->early_lo = c->early_hi = 0.0;
  const double *src = c->in_src ? c->in_src : c->in.p;      // (the caller's device buffer, or the library's copy)
  if (n_in > 0x7ffffff0LL) return fail(c, SK_ERR_ARG, "n_in too large");
  const size_t nfine_max = (size_t)(n_in >> SK_K8_TARGET_LOG) + 2;
  // control block: state | coarse histogram | fill counters (one per 32-byte sector) -- cleared by one memset;
  // then, not cleared: provisional unique offsets | dropped duplicates per bin | their exclusive sums
  const size_t off_hist = 256, off_fill = off_hist + sizeof(unsigned int) * SK_K8_NC;
  const size_t ctl_bytes = off_fill + sizeof(unsigned int) * nfine_max * SK_K8_FILL_STRIDE;
  const size_t off_uoff = ctl_bytes, off_dup = off_uoff + sizeof(unsigned int) * nfine_max;
  const size_t off_dsum = off_dup + sizeof(unsigned int) * nfine_max, all_bytes = off_dsum + sizeof(unsigned int) * nfine_max;
  static_assert(sizeof(SkK8State) <= 256, "control block layout");
  CK(c->k8_ctl.ensure(all_bytes));
  CK(c->k8_ctab.ensure(SK_K8_NC));
  CK(c->k8_slots.ensure(nfine_max * SK_K8_CAP));
  CK(c->uxs.ensure(n_in));               // sized for the worst case (n_unique <= n_in): no host round trip
  CK(c->uxs_fix.ensure(n_in));           // the compacted table when duplicates were dropped (k_k8_fix_uxs)
  CK(c->inv.ensure(n_in));
  CK(c->res.ensure(n_in));
  CK(c->pan.ensure(n_in));
  CK(c->stage.ensure(n_in));
  SkK8State *st = (SkK8State *)c->k8_ctl.p;
  unsigned int *chist = (unsigned int *)(c->k8_ctl.p + off_hist);
  unsigned int *fill = (unsigned int *)(c->k8_ctl.p + off_fill);
  unsigned int *uoffp = (unsigned int *)(c->k8_ctl.p + off_uoff);
  unsigned int *dup = (unsigned int *)(c->k8_ctl.p + off_dup);
  unsigned int *dsum = (unsigned int *)(c->k8_ctl.p + off_dsum);
  if (c->timing) CK(cudaEventRecord(c->ev[0], c->stream));
  static_assert((256 + sizeof(unsigned int) * SK_K8_NC) % 8 == 0 && (sizeof(unsigned int) * SK_K8_FILL_STRIDE) % 8 == 0, "8-byte words");
  k_zero_words<<<148, 256, 0, c->stream>>>((unsigned long long *)c->k8_ctl.p, (long long)(ctl_bytes / 8));   // (a kernel, not a memset node)
  LAUNCH_CHECK();
  const unsigned int gs = std::min<unsigned int>(nblk(n_in, 256), 148u * 8u);    // grid-stride passes: 2048 threads per SM
  k_k8_stats<<<gs, 256, 0, c->stream>>>(src, n_in, st);
  LAUNCH_CHECK();
  // the first panel is (0, quadm / (2 r_max)) (src/adaptive.jl:152) and r_max is known after this first pass: fetch the
  // key range now, and start the panel's source side on the prefetch stream while the sort runs
  // (a rank of a process-per-GPU run does not know the global distance range yet: no early prefetch there; a device
  //  group issues it for all its devices once every chunk's range is in: sk_group_targets_set)
  const bool early = c->prefetch_on && c->have_rule && c->family != SK_SDF_HOST && c->plan.w > 0 &&
                     (c->in_group || sharded(c) || n_in >= 65536);
  c->early_pending = early;
  c->early_global = false;
  if (early) {
    if (c->peer_n > 0) {
      // (peer mailboxes: the same 3-word MAX as one exchange kernel; the result lands in pinned memory, slot 1)
      int rcx = peer_exchange(c, 1, SK_PX_RANGE, 0, 0, 0, nullptr, 0, 0, st);
      if (rcx != SK_OK) return rcx;
      {
        int rcp = publish(c, &c->h_scal->k8, st, sizeof(SkK8State));                                   // (this rank's zeros)
        if (rcp != SK_OK) return rcp;
      }
      c->early_global = true;
    } else if (c->comm) {
      // process-per-GPU run: the panels are built from the GLOBAL distance range, so the ranks reduce their key
      // ranges here (one 3-word MAX all-reduce behind the first pass; every rank takes this branch or none does:
      // the condition above holds no rank-local quantity)
      NcclApi *N = nccl_api();
      unsigned long long *w = reinterpret_cast<unsigned long long *>(c->d_hv);
      k_k8_pack_range<<<1, 1, 0, c->stream>>>(st, w);
      LAUNCH_CHECK();
      NCK(N->AllReduce(w, w, 3, ncclUint64, ncclMax, c->comm, c->stream));
      CK(cudaMemcpyAsync(c->h_scal->hv, w, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
      c->early_global = true;
    } else {
      int rcp = publish(c, &c->h_scal->k8, st, sizeof(SkK8State));
      if (rcp != SK_OK) return rcp;
    }
    CK(cudaEventRecord(c->k8_ev, c->stream));
  }
  k_k8_sample<<<std::min<unsigned int>(nblk(n_in, 512), 148u), 512, 0, c->stream>>>(src, n_in, st, chist);
  LAUNCH_CHECK();
  k_k8_plan<<<1, 1024, 0, c->stream>>>(st, chist, n_in, c->k8_ctab.p);
  LAUNCH_CHECK();
  k_k8_scatter<<<gs, 256, 0, c->stream>>>(src, n_in, st, c->k8_ctab.p, fill, c->k8_slots.p, c->inv.p);
  LAUNCH_CHECK();
  k_k8_scan_bins<<<1, 1024, 0, c->stream>>>(st, fill, SK_K8_FILL_STRIDE, SK_K8_CAP, uoffp, 0);
  LAUNCH_CHECK();
  k_k8_finish<<<(unsigned int)nfine_max, SK_K8_TPB, 0, c->stream>>>(st, fill, c->k8_slots.p, uoffp, dup, c->uxs.p, c->inv.p);
  LAUNCH_CHECK();
  // (duplicates dropped by a bin: the compaction of the table and the shift of the inverse map are launched by
  //  targets_finish when the summary says so -- three launches less on the common path)
  k_k8_identity<<<gs, 256, 0, c->stream>>>(src, n_in, st, c->uxs.p, c->inv.p);
  LAUNCH_CHECK();
  k_k8_summary<<<1, 1, 0, c->stream>>>(st, c->uxs.p, c->uxs_fix.p, n_in, c->d_sum);
  LAUNCH_CHECK();
  {
    int rcp = publish(c, &c->h_scal->sum, c->d_sum, sizeof(SkTargetSummary));
    if (rcp != SK_OK) return rcp;
  }
  c->peer_summary_sent = false;
  if (c->peer_n > 0 && !c->in_group) {
    // process-per-GPU run over peer mailboxes: what the ranks exchange at the start of a run (global distance range, global
    // active count) goes out right behind the summary -- the host finds it when the sort's synchronisation returns
    // (sk_comm_summary) instead of paying for a synchronous gather
    int rcx = peer_exchange(c, 3, SK_PX_SUMMARY, 0, 0, 0, nullptr, 0, 0, nullptr, c->d_sum);
    if (rcx != SK_OK) return rcx;
    c->peer_summary_sent = true;
  }
  if (c->in_src) CK(cudaStreamWaitEvent(c->stream, c->in_ev[1], 0));   // the library's copy of the distances is complete
  CK(cudaEventRecord(c->ev_sum, c->stream));
  if (early && !c->in_group) {
    double r_lo = 0, r_hi = 0;
    int rc = targets_early_range(c, &r_lo, &r_hi);
    if (rc != SK_OK) return rc;
    c->early_lo = r_lo;
    c->early_hi = r_hi;
    if (r_hi > 0.0) {
      rc = targets_early_prefetch(c, r_lo, r_hi);
      if (rc != SK_OK) return rc;
    }
  }
  return SK_OK;
}

int targets_finish(sk_ctx *c, long long n_in, sk_target_info *info, bool force_general) {
  if (c->early1.pending) CK(cudaEventSynchronize(c->ev_sum));   // the first panel is queued behind the sort: do not wait for it
  else CK(cudaStreamSynchronize(c->stream));
  c->in_src = nullptr;                       // the caller's buffer is not referenced any more
  if (c->h_scal->sum.bad) return fail(c, SK_ERR_INPUT, "distances must be finite and >= 0");
  const bool general = force_general || c->h_scal->sum.overflow != 0;
  if (general) {
    int rc = targets_sort_general(c, n_in);
    if (rc != SK_OK) return rc;
  } else if (c->h_scal->sum.fixed) {
    // some bin dropped duplicates: compact the unique table into the second buffer, shift the inverse map, and take
    // the summary again from the compacted table (same layout of the control block as targets_enqueue)
    const size_t nfine_max = (size_t)(n_in >> SK_K8_TARGET_LOG) + 2;
    const size_t off_fill = 256 + sizeof(unsigned int) * SK_K8_NC;
    const size_t off_uoff = off_fill + sizeof(unsigned int) * nfine_max * SK_K8_FILL_STRIDE;
    const size_t off_dup = off_uoff + sizeof(unsigned int) * nfine_max, off_dsum = off_dup + sizeof(unsigned int) * nfine_max;
    SkK8State *st = (SkK8State *)c->k8_ctl.p;
    unsigned int *fill = (unsigned int *)(c->k8_ctl.p + off_fill), *uoffp = (unsigned int *)(c->k8_ctl.p + off_uoff);
    unsigned int *dup = (unsigned int *)(c->k8_ctl.p + off_dup), *dsum = (unsigned int *)(c->k8_ctl.p + off_dsum);
    const unsigned int gs = std::min<unsigned int>(nblk(n_in, 256), 148u * 8u);
    k_k8_scan_bins<<<1, 1024, 0, c->stream>>>(st, dup, 1, 0xffffffffu, dsum, 1);
    LAUNCH_CHECK();
    k_k8_fix_uxs<<<(unsigned int)nfine_max, 256, 0, c->stream>>>(st, fill, uoffp, dup, dsum, c->uxs.p, c->uxs_fix.p);
    LAUNCH_CHECK();
    k_k8_fix_inv<<<gs, 256, 0, c->stream>>>(st, uoffp, dsum, n_in, c->inv.p);
    LAUNCH_CHECK();
    k_k8_summary<<<1, 1, 0, c->stream>>>(st, c->uxs.p, c->uxs_fix.p, n_in, c->d_sum);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(&c->h_scal->sum, c->d_sum, sizeof(SkTargetSummary), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  const SkTargetSummary sm = c->h_scal->sum;
  if (sm.bad) return fail(c, SK_ERR_INPUT, "distances must be finite and >= 0");
  if (!general && sm.fixed) std::swap(c->uxs, c->uxs_fix);        // duplicates were dropped: the compacted table
  if (c->timing) {
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(cudaEventSynchronize(c->ev[1]));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.sort_ms = ms;
  }
  const long long nu = sm.n_unique;
  c->n_in = n_in;
  c->n_unique = nu;
  c->r0 = sm.r0; c->r1 = sm.r1; c->r_last = sm.r_last;
  c->have_orig = false;
  c->in_scale = 1.0;
  c->tails.n = 0;
  c->has_zero = (sm.r0 == 0.0);
  c->have_targets = true;
  c->in_panel = false;
  c->staged = false;
  c->commit_pending = false;
  c->scan_hi = -1;
  c->stats.sort_two_level = general ? 0 : (sm.presorted ? 2 : 1);
  if (info) {
    info->n_in = n_in;
    info->n_unique = nu;
    info->has_zero = c->has_zero ? 1 : 0;
    info->_pad = 0;
    info->r_min_pos = c->has_zero ? sm.r1 : sm.r0;
    info->r_max = sm.r_last;
  }
  return SK_OK;
}

int targets_from_device_buffer(sk_ctx *c, long long n_in, sk_target_info *info, bool force_general = false) {
  int rc = targets_enqueue(c, n_in);
  if (rc != SK_OK) return rc;
  return targets_finish(c, n_in, info, force_general);
}

int red_init(sk_ctx *c, SkReduceOut *d, long long max_unconv_init) {
  k_red_init<<<1, 1, 0, c->stream>>>(d, max_unconv_init);
  LAUNCH_CHECK();
  return SK_OK;
}
int publish(sk_ctx *c, void *host_dst, const void *dev_src, size_t bytes) {
  k_publish<<<1, 32, 0, c->stream>>>((unsigned long long *)host_dst, (const unsigned long long *)dev_src, (int)(bytes / 8));
  LAUNCH_CHECK();
  return SK_OK;
}
static_assert(sizeof(SkReduceOut) % 8 == 0 && sizeof(SkTargetSummary) % 8 == 0 && sizeof(SkK8State) % 8 == 0, "k_publish moves 8-byte words");

// ks = errs = 0 (src/adaptive.jl:122), written only when a consumer needs it (see res_zero_pending); the r = 0 row keeps
// what sk_zero_lag_set put there
int ensure_res_zero(sk_ctx *c) {
  if (!c->res_zero_pending) return SK_OK;
  c->res_zero_pending = false;
  const long long first = (c->has_zero && c->zero_lag_written) ? 1 : 0;
  if (c->n_unique > first)
    CK(cudaMemsetAsync(c->res.p + first, 0, sizeof(sk_cplx) * (c->n_unique - first), c->stream));
  return SK_OK;
}

// the lazy commit of the last panel, when no scan consumed it
int flush_commit(sk_ctx *c) {
  if (!c->commit_pending) return SK_OK;
  if (c->res_zero_pending) {
    int rcz = ensure_res_zero(c);
    if (rcz != SK_OK) return rcz;
  }
  const long long n = c->pend_hi - c->pend_lo;
  k_commit<<<nblk(n, 256), 256, 0, c->stream>>>(c->pan.p + c->pend_lo, c->res.p + c->pend_lo, n);
  LAUNCH_CHECK();
  c->commit_pending = false;
  return SK_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int sk_abi_version(void) { return SK_ABI_VERSION; }

const char *sk_error_string(int code) {
  switch (code) {
    case SK_OK: return "ok";
    case SK_ERR_CUDA: return "CUDA runtime error";
    case SK_ERR_ARG: return "invalid argument";
    case SK_ERR_STATE: return "invalid call sequence";
    case SK_ERR_NAN: return "NaN detected in panel integral...";
    case SK_ERR_SPLIT: return "sub-interval split too many times (b - a < 1e-16)";
    case SK_ERR_ALLOC: return "out of memory";
    case SK_ERR_CUFFT: return "cuFFT error";
    case SK_ERR_UNSUPPORTED: return "unsupported reference branch";
    case SK_ERR_INPUT: return "invalid distances";
    default: return "unknown error";
  }
}

const char *sk_last_error(const sk_ctx *ctx) { return ctx ? ctx->errmsg.c_str() : "null context"; }

int sk_ctx_create(int device, sk_ctx **out) {
  if (!out) return SK_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return SK_ERR_CUDA;  // no CPU fallback: fail loudly
  if (device < 0 || device >= ndev) return SK_ERR_ARG;
  sk_ctx *c = new sk_ctx();
  c->device = device;
  std::memset(&c->stats, 0, sizeof(c->stats));
  std::memset(&c->tails, 0, sizeof(c->tails));
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void **)&c->d_red, sizeof(SkReduceOut)) != cudaSuccess ||
      cudaMalloc((void **)&c->d_sum, sizeof(SkTargetSummary)) != cudaSuccess ||
      cudaMalloc((void **)&c->d_kb, sizeof(SkKeyBits)) != cudaSuccess ||
      cudaMallocHost((void **)&c->h_scal, sizeof(HostScalars)) != cudaSuccess) {
    delete c;
    return SK_ERR_CUDA;
  }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&c->ev[i]);
  for (int i = 0; i < 2; ++i) cudaEventCreate(&c->ev_user[i]);
  c->stream_main = c->stream;
  // the prefetch / copy stream gets the highest priority: its small source-side kernels (spread, FFT of the NEXT panel)
  // run next to an interpolation kernel whose thousands of queued blocks would otherwise keep them waiting -- and the
  // next panel (chained right behind, sk_subinterval_chain) cannot start before they are done
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->pf_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->pf2_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->k8_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_red, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_red2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_sum, cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc((void **)&c->d_gran, sizeof(unsigned int)) != cudaSuccess ||
      cudaHostAlloc((void **)&c->h_gran, 8, cudaHostAllocDefault) != cudaSuccess ||
      cudaMalloc((void **)&c->d_red2, sizeof(SkReduceOut)) != cudaSuccess ||
      cudaHostAlloc((void **)&c->h_red2, sizeof(SkReduceOut), cudaHostAllocDefault) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->in_ev[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->in_ev[1], cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return SK_ERR_CUDA;
  }
  if (sk_plan_make_es(width_from_eps(c->eps), &c->plan) != 0) {
    delete c;
    return SK_ERR_ARG;
  }
  *c->h_gran = 0u;
  std::memset(c->h_red2, 0, sizeof(SkReduceOut));
  *out = c;
  return SK_OK;
}

int sk_ctx_destroy(sk_ctx *c) {
  if (!c) return SK_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto &kv : c->fft_plans) cufftDestroy(kv.second);
  DevBuf<double> *dbl[] = {&c->leg_no1, &c->leg_wt1, &c->leg_no2, &c->leg_wt2, &c->jac_no1, &c->jac_wt1, &c->jac_no2,
                           &c->jac_wt2, &c->no1, &c->buf1, &c->no2, &c->buf2, &c->pos_hi1, &c->pos_lo1, &c->pos_hi2,
                           &c->pos_lo2, &c->imz, &c->in, &c->uxs, &c->uxs_fix, &c->uxs_orig, &c->out_v, &c->out_e};
  for (auto *b : dbl) b->release();
  c->cs1.release(); c->cs2.release(); c->fft.release(); c->dsum.release();
  c->res.release(); c->pan.release(); c->stage.release();
  c->fftB.release(); c->dsumB.release(); c->bufb1.release(); c->bufb2.release();
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  for (int i = 0; i < 2; ++i) {
    c->aout_v[i].release(); c->aout_e[i].release();
    if (c->ev_gather[i]) cudaEventDestroy(c->ev_gather[i]);
    if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
  }
  c->hk_tab.release(); c->hk_vals.release(); c->hk_cheb.release(); c->hk_loc.release(); c->hk_lam1.release(); c->hk_lam2.release();
  c->hk_lev.release(); c->hk_groups.release(); c->hk_grid.release(); c->hk_part.release(); c->hk_modes.release();
  if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
  peer_release(c);
  for (int i = 0; i < 5; ++i) if (c->peer_out[i]) cudaFreeHost(c->peer_out[i]);
  for (int i = 0; i < 2; ++i) if (c->d_gout[i]) cudaFree(c->d_gout[i]);
  if (c->d_ga) cudaFree(c->d_ga);
  if (c->d_gb) cudaFree(c->d_gb);
  if (c->d_hv) cudaFree(c->d_hv);
  if (c->d_sum) cudaFree(c->d_sum);
  if (c->d_kb) cudaFree(c->d_kb);
  c->keys.release(); c->keys_alt.release(); c->idx.release(); c->idx_alt.release();
  c->head.release(); c->uid.release(); c->inv.release(); c->cub_tmp.release();
  c->k8_ctl.release(); c->k8_ctab.release(); c->k8_slots.release();
  if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
  if (c->pf_ev) cudaEventDestroy(c->pf_ev);
  if (c->pf2_ev) cudaEventDestroy(c->pf2_ev);
  c->pf2.no1.release(); c->pf2.buf1.release(); c->pf2.no2.release(); c->pf2.buf2.release();
  c->pf2.pos_hi1.release(); c->pf2.pos_lo1.release(); c->pf2.pos_hi2.release(); c->pf2.pos_lo2.release();
  c->pf2.cs1.release(); c->pf2.cs2.release(); c->pf2.fft.release();
  if (c->k8_ev) cudaEventDestroy(c->k8_ev);
  for (int i = 0; i < 2; ++i) if (c->in_ev[i]) cudaEventDestroy(c->in_ev[i]);
  for (int i = 0; i < SK_GATHER_SLICES; ++i) if (c->ev_slice[i]) cudaEventDestroy(c->ev_slice[i]);
  c->pf.no1.release(); c->pf.buf1.release(); c->pf.no2.release(); c->pf.buf2.release();
  c->pf.pos_hi1.release(); c->pf.pos_lo1.release(); c->pf.pos_hi2.release(); c->pf.pos_lo2.release();
  c->pf.cs1.release(); c->pf.cs2.release(); c->pf.fft.release();
  if (c->d_red) cudaFree(c->d_red);
  if (c->d_red2) cudaFree(c->d_red2);
  if (c->h_red2) cudaFreeHost(c->h_red2);
  if (c->ev_red) cudaEventDestroy(c->ev_red);
  if (c->ev_red2) cudaEventDestroy(c->ev_red2);
  if (c->ev_sum) cudaEventDestroy(c->ev_sum);
  if (c->d_gran) cudaFree(c->d_gran);
  if (c->h_gran) cudaFreeHost(c->h_gran);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 2; ++i) if (c->ev_user[i]) cudaEventDestroy(c->ev_user[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return SK_OK;
}

int sk_ctx_set_timing(sk_ctx *c, int enabled) {
  if (!c) return SK_ERR_ARG;
  c->timing = enabled != 0;
  c->stats.timing_enabled = enabled != 0;
  return SK_OK;
}

int sk_ctx_set_nufft_eps(sk_ctx *c, double eps) {
  if (!c || !(eps > 0) || !(eps < 1)) return fail(c, SK_ERR_ARG, "eps must be in (0,1)");
  c->eps = eps;
  c->pf_valid = c->pf2_valid = false;
  if (sk_plan_make_es(width_from_eps(eps), &c->plan) != 0) return fail(c, SK_ERR_ARG, "cannot plan for eps=%g", eps);
  return SK_OK;
}

int sk_ctx_synchronize(sk_ctx *c) {
  if (!c) return SK_ERR_ARG;
  CK(cudaStreamSynchronize(c->stream));
  return SK_OK;
}

int sk_ctx_stream(sk_ctx *c, void **stream_out) {
  if (!c || !stream_out) return SK_ERR_ARG;
  *stream_out = (void *)c->stream;
  return SK_OK;
}

int sk_timer_begin(sk_ctx *c) {
  if (!c) return SK_ERR_ARG;
  CK(cudaEventRecord(c->ev_user[0], c->stream));
  return SK_OK;
}

int sk_timer_end(sk_ctx *c, double *ms) {
  if (!c || !ms) return SK_ERR_ARG;
  CK(cudaEventRecord(c->ev_user[1], c->stream));
  CK(cudaEventSynchronize(c->ev_user[1]));
  float f = 0;
  CK(cudaEventElapsedTime(&f, c->ev_user[0], c->ev_user[1]));
  *ms = f;
  return SK_OK;
}

int sk_fp64_peak(sk_ctx *c, double *tflops, double *ms_out) {
  if (!c || !tflops) return SK_ERR_ARG;
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  CK(c->dsum.ensure(16));
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(c->ev_user[0], c->stream));
    k_dfma_peak<<<blocks, 256, 0, c->stream>>>((double *)c->dsum.p, iters, 0.999999, 1e-9);
    CK(cudaEventRecord(c->ev_user[1], c->stream));
    CK(cudaEventSynchronize(c->ev_user[1]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->ev_user[0], c->ev_user[1]));
    if (rep > 0 && f < best) best = f;
  }
  const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return SK_OK;
}

// ---- target-sharded multi-GPU: communicator and collectives ---------------------------------------------
int sk_comm_unique_id(void *out128) {
  if (!out128) return SK_ERR_ARG;
  NcclApi *N = nccl_api();
  if (!N) return SK_ERR_UNSUPPORTED;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (N->GetUniqueId(&id) != ncclSuccess) return SK_ERR_CUDA;
  std::memcpy(out128, &id, sizeof(id));
  return SK_OK;
}

int sk_comm_init(sk_ctx *c, const void *id128, int32_t rank, int32_t nranks) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(c, SK_ERR_ARG, "bad communicator arguments");
  NcclApi *N = nccl_api();
  if (!N) return fail(c, SK_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  CK(cudaSetDevice(c->device));
  if (c->comm) { N->CommDestroy(c->comm); c->comm = nullptr; }
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  NCK(N->CommInitRank(&c->comm, nranks, id, rank));
  c->comm_rank = rank;
  c->comm_size = nranks;
  if (!c->d_ga) CK(cudaMalloc((void **)&c->d_ga, sizeof(SkGlobalA)));
  if (!c->d_gb) CK(cudaMalloc((void **)&c->d_gb, sizeof(SkGlobalB)));
  if (!c->d_hv) CK(cudaMalloc((void **)&c->d_hv, 32 * sizeof(double)));
  return SK_OK;
}

int sk_comm_destroy(sk_ctx *c) {
  if (!c) return SK_ERR_ARG;
  cudaSetDevice(c->device);
  if (c->comm) {
    cudaStreamSynchronize(c->stream);
    nccl_api()->CommDestroy(c->comm);
    c->comm = nullptr;
  }
  peer_release(c);
  c->comm_size = 1;
  c->comm_rank = 0;
  return SK_OK;
}

// ---- peer mailboxes (one process per GPU, NVLink / NVSwitch peer memory; see k_peer_exchange) ------------------
// sk_comm_peer_export: allocate this rank's mailbox and hand out its 64-byte CUDA IPC handle; the host all-gathers the
// handles of all ranks (torch.distributed / MPI / files) and passes them to sk_comm_peer_attach.  The all-gather also
// orders "every mailbox is zeroed" before "any rank writes into a peer's mailbox".
int sk_comm_peer_export(sk_ctx *c, void *handle64) {
  if (!c || !handle64) return fail(c, SK_ERR_ARG, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaSetDevice(c->device));
  peer_release(c);
  const size_t bytes = sizeof(unsigned long long) * 2 * SK_PEER_MAX * SK_PEER_WORDS;
  CK(cudaMalloc((void **)&c->peer_box, bytes));
  CK(cudaMemset(c->peer_box, 0, bytes));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->peer_box));
  std::memcpy(handle64, &h, sizeof(h));
  return SK_OK;
}

int sk_comm_peer_attach(sk_ctx *c, const void *handles, int32_t rank, int32_t nranks) {
  if (!c || !handles || nranks < 1 || nranks > SK_PEER_MAX || rank < 0 || rank >= nranks)
    return fail(c, SK_ERR_ARG, "bad peer arguments (at most %d ranks)", SK_PEER_MAX);
  if (!c->peer_box) return fail(c, SK_ERR_STATE, "sk_comm_peer_export first");
  CK(cudaSetDevice(c->device));
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) { c->peer_map[r] = c->peer_box; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char *)handles + (size_t)r * sizeof(h), sizeof(h));
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_map[r] = (unsigned long long *)p;
  }
  for (int i = 0; i < 5; ++i)
    if (!c->peer_out[i]) {
      CK(cudaHostAlloc((void **)&c->peer_out[i], sizeof(SkPeerOut), cudaHostAllocPortable));
      std::memset(c->peer_out[i], 0, sizeof(SkPeerOut));
    }
  for (int i = 0; i < 2; ++i)
    if (!c->d_gout[i]) {
      CK(cudaMalloc((void **)&c->d_gout[i], sizeof(SkPeerOut)));
      CK(cudaMemset(c->d_gout[i], 0, sizeof(SkPeerOut)));
    }
  c->sharded_chain = nranks <= 2;
  if (const char *sc = std::getenv("SK_SHARDED_CHAIN")) c->sharded_chain = std::atoi(sc) != 0;
  c->peer_slot = 0;
  c->comm_rank = rank;
  c->comm_size = nranks;
  c->peer_n = nranks;
  c->peer_epoch = 0;
  if (const char *t = std::getenv("SK_PEER_TIMEOUT_S")) c->peer_timeout_s = std::max(0.001, std::atof(t));
  return SK_OK;
}

// every rank's k <= 7 doubles: out[r * k + i] = value i of rank r.  One exchange; synchronous.  (Peer mailboxes only;
// with an NCCL communicator use sk_comm_allreduce on a one-hot layout.)
int sk_comm_allgather(sk_ctx *c, const double *vals, int32_t k, double *out) {
  if (!c || !vals || !out || k < 1 || k > 7) return fail(c, SK_ERR_ARG, "bad allgather arguments");
  if (c->peer_n <= 0) return fail(c, SK_ERR_STATE, "sk_comm_allgather needs peer mailboxes (sk_comm_peer_attach)");
  CK(cudaSetDevice(c->device));
  unsigned long long w[7] = {0, 0, 0, 0, 0, 0, 0};
  std::memcpy(w, vals, sizeof(double) * k);
  int rc = SK_OK;
  for (int round = 0;; ++round) {             // (again while void: it met another rank's skipped chained launch)
    if (round > 8) return fail(c, SK_ERR_STATE, "peer exchange: too many void rounds");
    rc = peer_exchange(c, 2, SK_PX_GATHER, 0, 0, 0, w, k, 0);
    if (rc != SK_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
    rc = peer_check(c, 2);
    if (rc != SK_OK) return rc;
    if (!c->peer_out[2]->void_flag) break;
  }
  std::memcpy(out, c->peer_out[2]->words, sizeof(double) * k * c->peer_n);
  return SK_OK;
}

// what the ranks exchange at the start of a run, as delivered behind the last sk_targets_set* / sk_targets_end (peer
// mailboxes only): smallest positive and largest distance over all ranks (0 when no rank has one) and the number of
// positive unique distances summed over the ranks.  *valid = 0: not available (no mailboxes, or some rank's sort took a
// path whose summary the host recomputes): every rank then exchanges them with sk_comm_allgather / sk_comm_allreduce.
int sk_comm_summary(sk_ctx *c, int32_t *valid, double *r_lo, double *r_hi, int64_t *n_active) {
  if (!c || !valid || !r_lo || !r_hi || !n_active) return SK_ERR_ARG;
  *valid = 0; *r_lo = 0.0; *r_hi = 0.0; *n_active = 0;
  if (c->peer_n <= 0 || !c->peer_summary_sent) return SK_OK;
  c->peer_summary_sent = false;
  int rc = peer_check(c, 3);
  if (rc != SK_OK) return rc;
  const unsigned long long *w = c->peer_out[3]->words;
  if (w[2]) return SK_OK;
  const unsigned long long kmin = ~w[0];
  if (w[0]) std::memcpy(r_lo, &kmin, sizeof(double));
  std::memcpy(r_hi, &w[1], sizeof(double));
  *n_active = (int64_t)w[3];
  *valid = 1;
  return SK_OK;
}

// generic collective on up to 32 host doubles (op: 0 max, 1 min, 2 sum); synchronous.  Used once per
// kernel_values call (global distance range / counts) and for the rare exact count.
int sk_comm_allreduce(sk_ctx *c, double *vals, int32_t n, int32_t op) {
  if (!c || !vals || n < 1 || n > 32 || op < 0 || op > 2) return fail(c, SK_ERR_ARG, "bad allreduce arguments");
  if (!sharded(c)) return SK_OK;
  CK(cudaSetDevice(c->device));
  if (c->peer_n > 0) {
    for (int i0 = 0; i0 < n; i0 += 7) {
      const int nw = std::min(7, n - i0);
      unsigned long long w[7] = {0, 0, 0, 0, 0, 0, 0};
      std::memcpy(w, vals + i0, sizeof(double) * nw);
      int rc = SK_OK;
      for (int round = 0;; ++round) {
        if (round > 8) return fail(c, SK_ERR_STATE, "peer exchange: too many void rounds");
        rc = peer_exchange(c, 2, SK_PX_RAW, 0, 0, 0, w, nw, op);
        if (rc != SK_OK) return rc;
        CK(cudaStreamSynchronize(c->stream));
        rc = peer_check(c, 2);
        if (rc != SK_OK) return rc;
        if (!c->peer_out[2]->void_flag) break;
      }
      std::memcpy(vals + i0, c->peer_out[2]->words, sizeof(double) * nw);
    }
    return SK_OK;
  }
  NcclApi *N = nccl_api();
  for (int i = 0; i < n; ++i) c->h_scal->hv[i] = vals[i];
  CK(cudaMemcpyAsync(c->d_hv, c->h_scal->hv, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  NCK(N->AllReduce(c->d_hv, c->d_hv, n, ncclDouble, op == 0 ? ncclMax : (op == 1 ? ncclMin : ncclSum), c->comm, c->stream));
  CK(cudaMemcpyAsync(c->h_scal->hv, c->d_hv, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; ++i) vals[i] = c->h_scal->hv[i];
  return SK_OK;
}

// a rank without active targets takes part in the collective points of the other ranks' sub-intervals
// (which = 0; 2 when the active ranks speculate) and scans (which = 1) with neutral contributions
int sk_comm_idle(sk_ctx *c, int32_t which) {
  if (!c) return SK_ERR_ARG;
  if (!sharded(c)) return SK_OK;
  CK(cudaSetDevice(c->device));
  if (which == 0 || which == 2) {
    c->peer_slot = 0;
    int rc = SK_OK;
    for (int round = 0;; ++round) {           // (again while the exchange is void: an active rank's chained launch skipped)
      if (round > 8) return fail(c, SK_ERR_STATE, "peer exchange: too many void rounds");
      rc = which == 2 ? comm_reduce_ab(c, 1, 0) : comm_reduce_a(c, 1);
      if (rc != SK_OK) return rc;
      CK(cudaStreamSynchronize(c->stream));
      rc = peer_check(c, 0);
      if (rc != SK_OK) return rc;
      if (!peer_void(c)) break;
    }
    unsigned int fl = 0;
    double mx = 0.0;
    comm_take_a(c, &fl, &mx);
    if (which == 2) comm_take_b(c);
    if (fl & SK_FLAG_NAND) mx = std::nan("");
    c->g_max_abs = mx;
    // the idle rank raises what the active ranks raise, so that all ranks leave the adaptive loop together
    if (global_a(c).err) return fail(c, SK_ERR_STATE, "another rank failed in this sub-interval");
    if (!(fl & SK_FLAG_NAN1) && (fl & SK_FLAG_NAN2)) return fail(c, SK_ERR_NAN, "NaN detected in panel integral...");
  } else {
    int rc = comm_reduce_b(c, false, 0ull, 0);
    if (rc != SK_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
    rc = peer_check(c, 0);
    if (rc != SK_OK) return rc;
    comm_take_b(c);
  }
  return SK_OK;
}

// test hook: the exchange protocol with `nranks` ranks emulated as the blocks of one cooperative launch on this
// context's device (ranks as separate launches must not share a GPU).  kind: 0 A, 1 AB, 6 gather of (base + rank).
// maxbits_in[r], rbits_in[r], top_in[r]: rank r's local scalars; out: per rank {maxbits, err, rbits, n_lb, status}.
int sk_comm_peer_selftest(sk_ctx *c, int32_t nranks, int32_t rounds, const uint64_t *maxbits_in, const uint64_t *rbits_in,
                          const int64_t *top_in, int64_t lo, int32_t skip_rank, uint64_t *out5) {
  if (!c || nranks < 1 || nranks > SK_PEER_MAX || rounds < 1 || !maxbits_in || !rbits_in || !top_in || !out5)
    return fail(c, SK_ERR_ARG, "bad selftest arguments");
  CK(cudaSetDevice(c->device));
  unsigned long long *boxes = nullptr;
  SkReduceOut *red = nullptr;
  SkPeerOut *outs = nullptr;
  const size_t bb = sizeof(unsigned long long) * 2 * SK_PEER_MAX * SK_PEER_WORDS * nranks;
  CK(cudaMalloc((void **)&boxes, bb));
  CK(cudaMemset(boxes, 0, bb));
  CK(cudaMalloc((void **)&red, sizeof(SkReduceOut) * nranks));
  CK(cudaMalloc((void **)&outs, sizeof(SkPeerOut) * nranks));
  CK(cudaMemset(outs, 0, sizeof(SkPeerOut) * nranks));
  std::vector<SkReduceOut> hr(nranks);
  int rc = SK_OK;
  for (int it = 1; it <= rounds && rc == SK_OK; ++it) {
    for (int r = 0; r < nranks; ++r) {
      std::memset(&hr[r], 0, sizeof(SkReduceOut));
      hr[r].maxbits = maxbits_in[r] + (unsigned long long)(it - 1);
      hr[r].rbits = rbits_in[r];
      hr[r].max_unconv = top_in[r];
      if (r == skip_rank && it == rounds) hr[r].flags = SK_FLAG_SKIPPED;     // its (chained) launch skipped itself: void
    }
    cudaMemcpyAsync(red, hr.data(), sizeof(SkReduceOut) * nranks, cudaMemcpyHostToDevice, c->stream);
    SkPeerArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = nranks; a.epoch = (unsigned long long)it; a.timeout_ns = 2000000000ull; a.kind = SK_PX_AB; a.lo = lo;
    void *args[] = {&a, &boxes, &red, &outs};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)k_peer_exchange_emul, dim3(nranks), dim3(32), args, 0, c->stream);
    if (e != cudaSuccess) rc = fail(c, SK_ERR_CUDA, "cooperative launch: %s", cudaGetErrorString(e));
  }
  if (rc == SK_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, SK_ERR_CUDA, "selftest kernel failed");
  if (rc == SK_OK) {
    std::vector<SkPeerOut> ho(nranks);
    cudaMemcpy(ho.data(), outs, sizeof(SkPeerOut) * nranks, cudaMemcpyDeviceToHost);
    for (int r = 0; r < nranks; ++r) {
      out5[5 * r + 0] = ho[r].ga.maxbits; out5[5 * r + 1] = ho[r].ga.err; out5[5 * r + 2] = ho[r].gb.rbits;
      out5[5 * r + 3] = (uint64_t)ho[r].gb.n_lb;
      out5[5 * r + 4] = ho[r].status | (ho[r].void_flag << 1) | (ho[r].epoch_done << 8);
    }
  }
  cudaFree(boxes); cudaFree(red); cudaFree(outs);
  return rc;
}

// global results of the last collective points: max |I2-I1| over all ranks (last sub-interval), stopping
// distance over all ranks and the summed lower bound of the active counts (last scan)
int sk_comm_last(sk_ctx *c, double *max_abs_diff, double *r_stop, int64_t *n_active_lb) {
  if (!c) return SK_ERR_ARG;
  if (max_abs_diff) *max_abs_diff = c->g_max_abs;
  if (r_stop) *r_stop = c->g_r_stop;
  if (n_active_lb) *n_active_lb = c->g_n_lb;
  return SK_OK;
}

int sk_host_alloc(size_t bytes, void **out) {
  if (!out) return SK_ERR_ARG;
  return cudaMallocHost(out, bytes) == cudaSuccess ? SK_OK : SK_ERR_ALLOC;
}
int sk_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? SK_OK : SK_ERR_CUDA; }

// ---- Level 0 ------------------------------------------------------------------------------------------
int sk_nufft1d3(sk_ctx *c, int64_t M, const double *w, const double *s, int64_t N, const double *x, double *out,
                double eps) {
  if (!c || M < 0 || N < 0 || (M > 0 && (!w || !s)) || (N > 0 && (!x || !out))) return fail(c, SK_ERR_ARG, "bad pointers/sizes");
  if (N == 0) return SK_OK;
  if (M == 0) { std::memset(out, 0, sizeof(double) * 2 * N); return SK_OK; }
  CK(cudaSetDevice(c->device));
  SkEsPlan saved = c->plan;
  if (eps > 0 && eps != c->eps && sk_plan_make_es(width_from_eps(eps), &c->plan) != 0) return fail(c, SK_ERR_ARG, "bad eps");
  // the gather spread needs ascending sources
  std::vector<long long> perm(M);
  std::iota(perm.begin(), perm.end(), 0LL);
  if (!std::is_sorted(w, w + M)) std::sort(perm.begin(), perm.end(), [&](long long a, long long b) { return w[a] < w[b]; });
  std::vector<double> hw(M), hre(M), him(M);
  for (long long i = 0; i < M; ++i) { hw[i] = w[perm[i]]; hre[i] = s[2 * perm[i]]; him[i] = s[2 * perm[i] + 1]; }
  double xmin = x[0], xmax = x[0];
  for (long long j = 1; j < N; ++j) { xmin = std::min(xmin, x[j]); xmax = std::max(xmax, x[j]); }
  SkGeom G;
  int rc = SK_OK;
  if (sk_make_geom(c->plan, hw.front(), hw.back(), xmin, xmax, &G) != 0) rc = fail(c, SK_ERR_ARG, "type-3 grid too large");
  auto body = [&]() -> int {
    CK(c->no1.ensure(M)); CK(c->buf1.ensure(M)); CK(c->imz.ensure(M));
    CK(c->in.ensure(N)); CK(c->dsum.ensure(N));
    CK(cudaMemcpyAsync(c->no1.p, hw.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->buf1.p, hre.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->imz.p, him.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->in.p, x, sizeof(double) * N, cudaMemcpyHostToDevice, c->stream));
    int r2 = run_source_side(c, G, 1, M, c->buf1.p, c->imz.p, 0, nullptr, c->fft);
    if (r2 != SK_OK) return r2;
#define CALL(WW) launch_interp_cplx<WW>(c, G, c->in.p, N, c->dsum.p)
    DISPATCH_W(c->plan.w, CALL)
#undef CALL
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(out, c->dsum.p, sizeof(double) * 2 * N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SK_OK;
  };
  if (rc == SK_OK) rc = body();
  c->plan = saved;
  c->pf_valid = c->pf2_valid = false;
  c->have_sources = false;
  c->have_targets = false;  // the target buffer was reused
  return rc;
}

// ---- rules ------------------------------------------------------------------------------------------------
int sk_rule_set(sk_ctx *c, int32_t m, int32_t k, double p, const double *leg_no1, const double *leg_wt1,
                const double *leg_no2, const double *leg_wt2, const double *jac_no1, const double *jac_wt1,
                const double *jac_no2, const double *jac_wt2) {
  if (!c || m < 1 || k < 1) return fail(c, SK_ERR_ARG, "quadspec must be positive");
  if (k > SK_KMAX) return fail(c, SK_ERR_UNSUPPORTED, "k = %d sub-panels > %d", k, SK_KMAX);
  if (p != 0.0 && !(p > -1.0)) return fail(c, SK_ERR_ARG, "p needs to be in (-1.0, Inf) to be integrable");  // quadrature.jl:40
  CK(cudaSetDevice(c->device));
  const bool given_leg = leg_no1 && leg_wt1 && leg_no2 && leg_wt2;
  const bool given_jac = jac_no1 && jac_wt1 && jac_no2 && jac_wt2;
  // library-generated rules for the same (m, k, p) are already resident: nothing to do (every kernel_values call
  // of a fitting loop comes through here)
  if (!given_leg && !given_jac && c->have_rule && c->rule_generated && c->m == m && c->k == k && c->p == p) return SK_OK;
  const int sizes[8] = {m, m, 2 * m, 2 * m, m, m, 2 * m, 2 * m};
  for (int i = 0; i < 8; ++i) c->h_rule[i].assign(sizes[i], 0.0);
  {
    std::vector<std::pair<int, double>> want;
    if (!given_leg) { want.push_back({m, 0.0}); want.push_back({2 * m, 0.0}); }
    if (p != 0.0 && !given_jac) { want.push_back({m, p}); want.push_back({2 * m, p}); }
    if (!want.empty() && ensure_gauss_rules(c, want) != 0)
      return fail(c, SK_ERR_ARG, "Gauss rule generation failed for m=%d p=%g", m, p);
  }
  if (given_leg) {
    const double *src[4] = {leg_no1, leg_wt1, leg_no2, leg_wt2};
    for (int i = 0; i < 4; ++i) std::copy(src[i], src[i] + sizes[i], c->h_rule[i].begin());
  } else {
    if (cached_gauss_rule(c, m, 0.0, c->h_rule[0], c->h_rule[1]) != 0 ||
        cached_gauss_rule(c, 2 * m, 0.0, c->h_rule[2], c->h_rule[3]) != 0)
      return fail(c, SK_ERR_ARG, "Gauss-Legendre generation failed for m=%d", m);
  }
  c->have_jac = (p != 0.0);
  if (c->have_jac) {
    if (given_jac) {
      const double *src[4] = {jac_no1, jac_wt1, jac_no2, jac_wt2};
      for (int i = 0; i < 4; ++i) std::copy(src[i], src[i] + sizes[4 + i], c->h_rule[4 + i].begin());
    } else {
      if (cached_gauss_rule(c, m, p, c->h_rule[4], c->h_rule[5]) != 0 ||
          cached_gauss_rule(c, 2 * m, p, c->h_rule[6], c->h_rule[7]) != 0)
        return fail(c, SK_ERR_ARG, "Gauss-Jacobi generation failed for m=%d p=%g", m, p);
    }
  } else {
    for (int i = 0; i < 4; ++i) c->h_rule[4 + i] = c->h_rule[i];  // jacrule = legrule, src/adaptive.jl:49
  }
  DevBuf<double> *dst[8] = {&c->leg_no1, &c->leg_wt1, &c->leg_no2, &c->leg_wt2, &c->jac_no1, &c->jac_wt1, &c->jac_no2, &c->jac_wt2};
  for (int i = 0; i < 8; ++i) {
    int rc = upload_rule(c, *dst[i], c->h_rule[i].data(), sizes[i]);
    if (rc != SK_OK) return rc;
  }
  const long long M1 = (long long)m * k;
  CK(c->no1.ensure(M1)); CK(c->buf1.ensure(M1)); CK(c->no2.ensure(2 * M1)); CK(c->buf2.ensure(2 * M1));
  CK(cudaStreamSynchronize(c->stream));
  c->m = m; c->k = k; c->p = p;
  c->pf_valid = c->pf2_valid = false;
  c->have_rule = true;
  c->rule_generated = !given_leg && !(c->have_jac && given_jac);
  return SK_OK;
}

int sk_rule_get(sk_ctx *c, int32_t which, double *no, double *wt) {
  if (!c || which < 0 || which > 3 || !no || !wt) return fail(c, SK_ERR_ARG, "bad arguments");
  if (!c->have_rule) return fail(c, SK_ERR_STATE, "no rule set");
  const std::vector<double> &a = c->h_rule[2 * which], &b = c->h_rule[2 * which + 1];
  std::copy(a.begin(), a.end(), no);
  std::copy(b.begin(), b.end(), wt);
  return SK_OK;
}

int sk_sdf_builtin(sk_ctx *c, int32_t family, const double *params, int32_t nparams, int32_t deriv_index) {
  if (!c) return SK_ERR_ARG;
  int want = family == SK_SDF_MATERN ? 4 : family == SK_SDF_EXPONENTIAL ? 2 : family == SK_SDF_HOST ? 0 : -1;
  if (want < 0) return fail(c, SK_ERR_ARG, "unknown family %d", family);
  if (nparams != want || (want > 0 && !params)) return fail(c, SK_ERR_ARG, "family %d takes %d parameters", family, want);
  if (deriv_index < 0 || deriv_index > want - (family == SK_SDF_MATERN ? 1 : 0)) return fail(c, SK_ERR_ARG, "bad deriv_index");
  c->family = family;
  c->nparam = nparams;
  c->deriv = deriv_index;
  for (int i = 0; i < nparams; ++i) c->params[i] = params[i];
  return SK_OK;
}

// ---- targets -------------------------------------------------------------------------------------------
int sk_targets_set(sk_ctx *c, const double *xs_host, int64_t n_in, sk_target_info *info) {
  if (!c || !xs_host || n_in < 1) return fail(c, SK_ERR_ARG, "need at least one distance");
  CK(cudaSetDevice(c->device));
  CK(c->in.ensure(n_in));
  CK(cudaMemcpyAsync(c->in.p, xs_host, sizeof(double) * n_in, cudaMemcpyHostToDevice, c->stream));
  return targets_from_device_buffer(c, n_in, info);
}

int sk_targets_set_device(sk_ctx *c, const double *xs_dev, int64_t n_in, sk_target_info *info) {
  if (!c || !xs_dev || n_in < 1) return fail(c, SK_ERR_ARG, "need at least one distance");
  CK(cudaSetDevice(c->device));
  CK(c->in.ensure(n_in));
  if (c->stream2) {
    // K8 reads the caller's buffer; the library's own copy is made on the copy stream, next to the sort (it starts once
    // everything queued so far -- a previous call's gather reads the old copy -- has drained)
    CK(cudaEventRecord(c->in_ev[0], c->stream));
    CK(cudaStreamWaitEvent(c->stream2, c->in_ev[0], 0));
    CK(cudaMemcpyAsync(c->in.p, xs_dev, sizeof(double) * n_in, cudaMemcpyDeviceToDevice, c->stream2));
    CK(cudaEventRecord(c->in_ev[1], c->stream2));
    c->in_src = xs_dev;
  } else {
    CK(cudaMemcpyAsync(c->in.p, xs_dev, sizeof(double) * n_in, cudaMemcpyDeviceToDevice, c->stream));
  }
  const int rc_t = targets_from_device_buffer(c, n_in, info);
  c->in_src = nullptr;
  return rc_t;
}

// sk_targets_set[_device] in two halves: everything is enqueued by _begin (which returns as soon as the first pass over
// the distances has delivered their range), the host does its scalar work for the first panel -- the panel is
// (0, m k / (2 r_max)) (src/adaptive.jl:152), so estimate_tail_decay (:204-220) and the scan arguments can be evaluated
// while the device sorts -- and _end waits for the sort.
int sk_targets_begin(sk_ctx *c, const double *xs_host, int64_t n_in) {
  if (!c || !xs_host || n_in < 1) return fail(c, SK_ERR_ARG, "need at least one distance");
  CK(cudaSetDevice(c->device));
  CK(c->in.ensure(n_in));
  CK(cudaMemcpyAsync(c->in.p, xs_host, sizeof(double) * n_in, cudaMemcpyHostToDevice, c->stream));
  c->begin_n = -1;
  int rc = targets_enqueue(c, n_in);
  if (rc == SK_OK) c->begin_n = n_in;
  return rc;
}
int sk_targets_begin_device(sk_ctx *c, const double *xs_dev, int64_t n_in) {
  if (!c || !xs_dev || n_in < 1) return fail(c, SK_ERR_ARG, "need at least one distance");
  CK(cudaSetDevice(c->device));
  CK(c->in.ensure(n_in));
  c->begin_n = -1;
  if (c->stream2) {
    CK(cudaEventRecord(c->in_ev[0], c->stream));
    CK(cudaStreamWaitEvent(c->stream2, c->in_ev[0], 0));
    CK(cudaMemcpyAsync(c->in.p, xs_dev, sizeof(double) * n_in, cudaMemcpyDeviceToDevice, c->stream2));
    CK(cudaEventRecord(c->in_ev[1], c->stream2));
    c->in_src = xs_dev;
  } else {
    CK(cudaMemcpyAsync(c->in.p, xs_dev, sizeof(double) * n_in, cudaMemcpyDeviceToDevice, c->stream));
  }
  int rc = targets_enqueue(c, n_in);
  if (rc == SK_OK) c->begin_n = n_in;
  else c->in_src = nullptr;
  return rc;
}
// the distance range [r_lo, r_hi] of the distances being sorted (over all ranks when a communicator / mailboxes are
// attached); r_hi = 0 when it is not known before the sort ends (small inputs, host-evaluated densities)
int sk_targets_early_range(sk_ctx *c, double *r_lo, double *r_hi) {
  if (!c || !r_lo || !r_hi) return SK_ERR_ARG;
  *r_lo = c->begin_n >= 0 ? c->early_lo : 0.0;
  *r_hi = c->begin_n >= 0 ? c->early_hi : 0.0;
  return SK_OK;
}
int sk_targets_end(sk_ctx *c, sk_target_info *info) {
  if (!c) return SK_ERR_ARG;
  if (c->begin_n < 0) return fail(c, SK_ERR_STATE, "sk_targets_begin first");
  CK(cudaSetDevice(c->device));
  const long long n_in = c->begin_n;
  c->begin_n = -1;
  const int rc = targets_finish(c, n_in, info, false);
  c->in_src = nullptr;
  return rc;
}

int sk_targets_set_pairs(sk_ctx *c, const double *pts_host, int64_t npts, int32_t dim, const int64_t *pairs_host,
                         int64_t npairs, sk_target_info *info) {
  if (!c || !pts_host || npts < 1 || dim < 1) return fail(c, SK_ERR_ARG, "bad points");
  CK(cudaSetDevice(c->device));
  const long long all_pairs = (long long)npts * (npts - 1) / 2;
  const long long np = pairs_host ? (long long)npairs : all_pairs;
  if (np < 1) return fail(c, SK_ERR_ARG, "no pairs");
  if (pairs_host)
    for (long long t = 0; t < 2 * np; ++t)
      if (pairs_host[t] < 0 || pairs_host[t] >= npts) return fail(c, SK_ERR_ARG, "pair index out of range");
  DevBuf<double> d_pts;
  DevBuf<long long> d_pairs;
  CK(d_pts.ensure((size_t)npts * dim));
  CK(cudaMemcpyAsync(d_pts.p, pts_host, sizeof(double) * npts * dim, cudaMemcpyHostToDevice, c->stream));
  if (pairs_host) {
    CK(d_pairs.ensure((size_t)2 * np));
    CK(cudaMemcpyAsync(d_pairs.p, pairs_host, sizeof(long long) * 2 * np, cudaMemcpyHostToDevice, c->stream));
  }
  CK(c->in.ensure(np));
  k_pair_lags<<<nblk(np, 256), 256, 0, c->stream>>>(d_pts.p, npts, dim, pairs_host ? d_pairs.p : nullptr, np, c->in.p);
  LAUNCH_CHECK();
  int rc = targets_from_device_buffer(c, np, info);
  d_pts.release();
  d_pairs.release();
  return rc;
}

// Linear warping of the lags: every unique distance becomes (original distance) * factor.  The sort, the unique
// table's order and the inverse map stay valid, so a fitting loop whose range parameter enters through the
// warping function x -> x / rho (src/model.jl:62-66, scripts/fit_vecchia_demo.jl:15) re-uses the sorted lags of
// sk_targets_set* for every hyperparameter vector.  The factor always applies to the ORIGINAL distances.
int sk_targets_scale(sk_ctx *c, double factor, sk_target_info *info) {
  if (!c) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "no targets set");
  if (!(factor > 0.0) || std::isinf(factor)) return fail(c, SK_ERR_ARG, "the scale factor must be positive and finite");
  CK(cudaSetDevice(c->device));
  if (!c->have_orig) {
    CK(c->uxs_orig.ensure(c->n_unique));
    CK(cudaMemcpyAsync(c->uxs_orig.p, c->uxs.p, sizeof(double) * c->n_unique, cudaMemcpyDeviceToDevice, c->stream));
    c->r0_orig = c->r0; c->r1_orig = c->r1; c->r_last_orig = c->r_last;
    c->have_orig = true;
  }
  k_scale_targets<<<nblk(c->n_unique, 256), 256, 0, c->stream>>>(c->uxs_orig.p, c->uxs.p, c->n_unique, factor);
  LAUNCH_CHECK();
  c->r0 = c->r0_orig * factor; c->r1 = c->r1_orig * factor; c->r_last = c->r_last_orig * factor;   // same rounding as the kernel
  c->in_scale = factor;
  c->tails.n = 0;
  c->in_panel = false;
  c->staged = false;
  c->commit_pending = false;
  c->scan_hi = -1;
  if (info) {
    info->n_in = c->n_in;
    info->n_unique = c->n_unique;
    info->has_zero = c->has_zero ? 1 : 0;
    info->_pad = 0;
    info->r_min_pos = c->has_zero ? c->r1 : c->r0;
    info->r_max = c->r_last;
  }
  return SK_OK;
}

int sk_target_value(sk_ctx *c, int64_t idx, double *out) {
  if (!c || !out) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "no targets set");
  if (idx < 1 || idx > c->n_unique) return fail(c, SK_ERR_ARG, "index out of range");
  CK(cudaMemcpyAsync(&c->h_scal->r[0], c->uxs.p + (idx - 1), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *out = c->h_scal->r[0];
  return SK_OK;
}

// ---- adaptive loop steps -----------------------------------------------------------------------------------
int sk_run_begin(sk_ctx *c) {
  if (!c) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "sk_targets_set first");
  CK(cudaSetDevice(c->device));
  {
    int rcd = chain_discard(c);
    if (rcd != SK_OK) return rcd;
  }
  const bool t = c->timing;
  const int two = c->stats.sort_two_level;
  const double sort_ms = c->stats.sort_ms;
  std::memset(&c->stats, 0, sizeof(c->stats));
  c->stats.timing_enabled = t;
  c->stats.sort_two_level = two;
  c->stats.sort_ms = sort_ms;
  c->res_zero_pending = true;               // ks = errs = 0, src/adaptive.jl:122: see ensure_res_zero
  c->zero_lag_written = false;
  c->spec_fresh = false;
  c->tails.n = 0;
  c->in_panel = false;
  c->staged = false;
  c->commit_pending = false;
  c->scan_hi = -1;
  return SK_OK;
}

int sk_zero_lag_set(sk_ctx *c, double value) {
  if (!c) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "sk_targets_set first");
  if (!c->has_zero) return SK_OK;
  k_set_zero_lag<<<1, 1, 0, c->stream>>>(c->res.p, value);
  LAUNCH_CHECK();
  c->zero_lag_written = true;               // (a later ensure_res_zero leaves the row alone)
  return SK_OK;
}

int sk_panel_begin(sk_ctx *c, int64_t ix1, int64_t hi, double *r_lo, double *r_hi) {
  if (!c) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "sk_targets_set first");
  if (ix1 < 1 || hi < ix1 || hi > c->n_unique) return fail(c, SK_ERR_ARG, "bad index range [%lld,%lld]", (long long)ix1, (long long)hi);
  CK(cudaSetDevice(c->device));
  int rc = flush_commit(c);
  if (rc != SK_OK) return rc;
  c->lo = ix1 - 1;
  c->hi = hi;
  // the end points are normally known on the host already (target summary / last scan): no D2H
  bool need_lo = true, need_hi = true;
  if (ix1 == 1) { c->r_lo = c->r0; need_lo = false; }
  else if (ix1 == 2) { c->r_lo = c->r1; need_lo = false; }
  if (hi == c->n_unique) { c->r_hi = c->r_last; need_hi = false; }
  else if (hi == c->scan_hi) { c->r_hi = c->scan_r; need_hi = false; }
  else if (hi == 1) { c->r_hi = c->r0; need_hi = false; }
  if (need_lo || need_hi) {
    CK(cudaMemcpyAsync(&c->h_scal->r[0], c->uxs.p + c->lo, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&c->h_scal->r[1], c->uxs.p + (c->hi - 1), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->r_lo = c->h_scal->r[0];
    c->r_hi = c->h_scal->r[1];
  }
  if (r_lo) *r_lo = c->r_lo;
  if (r_hi) *r_hi = c->r_hi;
  c->in_panel = true;
  c->staged = false;
  c->first_accept = true;
  c->spec_active = false;
  c->spec_accepted = false;
  c->panel_subs = 0;
  c->n_act_global = 0;
  return SK_OK;
}

int sk_panel_set_range(sk_ctx *c, double r_lo, double r_hi, int64_t n_active_global) {
  if (!c) return SK_ERR_ARG;
  if (!c->in_panel) return fail(c, SK_ERR_STATE, "no open panel");
  if (!(r_lo <= c->r_lo) || !(r_hi >= c->r_hi)) return fail(c, SK_ERR_ARG, "range must contain the panel's targets");
  c->r_lo = r_lo;
  c->r_hi = r_hi;
  if (n_active_global > 0) {
    if (n_active_global < c->hi - c->lo) return fail(c, SK_ERR_ARG, "global active count below the local one");
    c->n_act_global = n_active_global;
  }
  return SK_OK;
}

static int subinterval_prologue(sk_ctx *c, double a, double b, const sk_subinterval_opts *o) {
  if (!c || !o) return SK_ERR_ARG;
  if (!c->have_rule) return fail(c, SK_ERR_STATE, "sk_rule_set first");
  if (!c->in_panel) return fail(c, SK_ERR_STATE, "sk_panel_begin first");
  if (o->kernel != SK_KERNEL_COS && o->kernel != SK_KERNEL_SIN && o->kernel != SK_KERNEL_BESSEL)
    return fail(c, SK_ERR_UNSUPPORTED, "kernel %d", o->kernel);
  if (o->kernel == SK_KERNEL_BESSEL && (o->nu < 0 || o->nu > 64)) return fail(c, SK_ERR_ARG, "Bessel order %d", o->nu);
  // check_subdivide_failure, src/utils.jl:28-36
  if (!(std::fabs(b - a) > 1e-16))
    return fail(c, SK_ERR_SPLIT,
                "The sub-interval (a, b) = (%.17g, %.17g) has been split too many times (b - a < 1e-16). "
                "Exiting to avoid infinite splitting.", a, b);
  CK(cudaSetDevice(c->device));
  return rollback_speculation(c);
}

static int subinterval_builtin_enqueue(sk_ctx *c, double a, double b, const sk_subinterval_opts *o) {
  int rc = subinterval_prologue(c, a, b, o);
  if (rc != SK_OK) return rc;
  if (c->family == SK_SDF_HOST) return fail(c, SK_ERR_STATE, "no built-in spectral density set: use sk_subinterval_host");
  if (o->p != c->p) return fail(c, SK_ERR_ARG, "opts.p (%g) differs from the rule's p (%g)", o->p, c->p);
  const bool origin = (a == 0.0 && c->p != 0.0);                    // src/quadrature.jl:185
  if (origin && o->logw)
    return fail(c, SK_ERR_UNSUPPORTED, "log-weighted origin sub-interval needs df: use sk_subinterval_logw_host (src/quadrature.jl:186-228)");
  if (c->early1.pending) {
    if (early_matches(c, a, b, o)) {
      // this very sub-interval has been running since the sort ended (sk_first_panel_early)
      c->early1.adopted = true;
      c->res_zero_pending = false;                       // the launch writes the table outright (SkSpec::fresh)
      if (c->has_zero && !c->zero_lag_written) CK(cudaMemsetAsync(c->res.p, 0, sizeof(sk_cplx), c->stream));
      c->spec_fresh = true;
      c->pend_spec = true;
      c->pend_timed = false;
      c->pend_rc = SK_OK;
      c->pend_ab = sharded(c);
      c->peer_slot = 0;
      c->cur_red = c->d_red;
      c->spec_args = c->early1.sa;
      c->need_gen = false;
      c->have_sources = true;
      c->last_logw = o->logw ? 1 : 0;
      c->stats.n_fast++;
      c->stats.last_nf = c->early1.G.nf;
      c->stats.last_nf2 = c->early1.G.nf2;
      return SK_OK;
    }
    rc = early_discard(c);
    if (rc != SK_OK) return rc;
  }
  if (c->chain.pending) {
    if (chain_matches(c, a, b, o)) {
      // this very sub-interval was enqueued ahead of time: nothing to launch, _finish reads the chain's scalars
      c->chain.adopted = true;
      c->pend_spec = true;
      c->pend_timed = false;
      c->pend_rc = SK_OK;
      c->pend_ab = sharded(c);               // (sharded: its exchange carried the scan's scalars)
      c->spec_args = c->chain.sa;
      c->spec_fresh = false;
      c->need_gen = false;
      c->have_sources = true;
      c->last_logw = o->logw ? 1 : 0;
      c->stats.n_fast++;
      c->stats.last_nf = c->chain_G.nf;
      c->stats.last_nf2 = c->chain_G.nf2;
      return SK_OK;
    }
    rc = chain_discard(c);
    if (rc != SK_OK) return rc;
  }
  make_panel_spec(c, a, b, o->logw, &c->pend_S);
  c->need_gen = true;                       // generated -- or taken from the prefetch set -- in transform_and_stage
  c->last_logw = o->logw ? 1 : 0;
  if (c->timing) CK(cudaEventRecord(c->ev[3], c->stream));
  return transform_and_stage_enqueue(c, a, b, o);
}

int sk_subinterval(sk_ctx *c, double a, double b, const sk_subinterval_opts *o, double *max_abs_diff) {
  if (c && !max_abs_diff) return fail(c, SK_ERR_ARG, "null output");
  int rc = subinterval_builtin_enqueue(c, a, b, o);
  if (rc != SK_OK) return rc;
  return transform_and_stage_finish(c, max_abs_diff, nullptr);
}

// Between sk_targets_begin* and sk_targets_end: enqueue the FIRST panel's first sub-interval (0, b1) behind the sort.
// Everything the launch needs is known after the first pass over the distances (their range -> panel ends and transform
// geometry; whether there is an r = 0 row) except the number of unique distances and the buffer the unique table ends up
// in: the kernel reads those from the sort's device-side summary (SkSpec::dyn) and skips itself if the sort did not
// deliver.  The host's sk_run_begin / sk_panel_begin / sk_subinterval[_begin](0, b1, opts) that follow sk_targets_end then
// find the panel already being integrated.  *queued = 0: not applicable, nothing was enqueued.
int sk_first_panel_early(sk_ctx *c, double a, double b, const sk_subinterval_opts *o, int32_t *queued) {
  if (!c || !o || !queued) return SK_ERR_ARG;
  *queued = 0;
  if (c->begin_n < 0 || c->early1.pending || c->chain.pending || !chain_allowed(c) || c->in_group || c->timing ||
      c->interp_mode != 0 || c->family == SK_SDF_HOST || !c->have_rule || o->speculate == nullptr ||
      (o->kernel != SK_KERNEL_COS && o->kernel != SK_KERNEL_SIN) || o->p != c->p || a != 0.0 || !(b > 0.0) ||
      (c->p != 0.0 && o->logw) || o->speculate->criteria < 0 || o->speculate->criteria > 2 ||
      (c->early_global && !(c->peer_n > 0)) || !(c->early_hi > 0.0) || c->plan.w <= 0)
    return SK_OK;
  CK(cudaSetDevice(c->device));
  const SkK8State &k8 = c->h_scal->k8;              // fetched behind the first pass (targets_early_range)
  if (k8.bad) return SK_OK;
  const long long n_in = c->begin_n, lo = k8.nzero ? 1 : 0, M2 = 2LL * c->m * c->k;
  const long long min_n = std::max<long long>(1, (1LL << 18) / M2);         // NUFFT branch: M2 n > 2^18 and n > 1
  if (n_in - lo <= min_n) return SK_OK;
  SkGeom G;
  if (sk_make_geom(c->plan, a, b, c->early_lo, c->early_hi, &G) != 0) return SK_OK;
  SkPanelSpec S;
  make_panel_spec(c, a, b, o->logw, &S);
  if (!c->pf_valid || std::memcmp(&c->pf_S, &S, sizeof(SkPanelSpec)) != 0 || std::memcmp(&c->pf_G, &G, sizeof(SkGeom)) != 0)
    return SK_OK;                                    // its sources are not waiting in the prefetch set
  CK(c->res.ensure(n_in));
  CK(c->stage.ensure(n_in));
  const long long lo_save = c->lo;
  c->lo = lo;                                        // fill_spec offsets the table pointers by c->lo
  SkSpec spec;
  std::memset(&spec, 0, sizeof(spec));
  fill_spec(c, o->speculate, spec);
  spec.fresh = 1;
  spec.dyn = c->d_sum;
  spec.dyn_lo = lo;
  spec.dyn_min_n = min_n;
  {
    int rci = red_init(c, c->d_red, lo - 1);
    if (rci != SK_OK) { c->lo = lo_save; return rci; }
  }
  CK(cudaStreamWaitEvent(c->stream, c->pf_ev, 0));
  swap_src_sets(c);
  c->n_pf_hits++;
  promote_second_prefetch(c);
  const double r_lo_save = c->r_lo, r_hi_save = c->r_hi;
  c->r_lo = c->early_lo; c->r_hi = c->early_hi;      // (the launch helper sizes its blocks from the distance span)
  const int ksin = o->kernel == SK_KERNEL_SIN;
#define CALL(WW) launch_interp_session<WW>(c, G, c->uxs.p + lo, n_in - lo, o->cmul, ksin, spec)
  DISPATCH_W(c->plan.w, CALL)
#undef CALL
  c->lo = lo_save; c->r_lo = r_lo_save; c->r_hi = r_hi_save;
  LAUNCH_CHECK();
  {
    int rcp = publish(c, &c->h_scal->red, c->d_red, sizeof(SkReduceOut));
    if (rcp != SK_OK) return rcp;
  }
  if (sharded(c)) {
    // sharded run: the sub-interval's exchange (max |I2-I1| and the scan's scalars over all ranks) goes out behind it; a
    // launch that skipped itself makes it void and every rank then evaluates the panel the ordinary way
    int rcx = peer_exchange(c, 0, SK_PX_AB, 0, 0, lo, nullptr, 0, 0, nullptr, nullptr, c->d_red);
    if (rcx != SK_OK) return rcx;
  }
  CK(cudaEventRecord(c->ev_red, c->stream));
  sk_ctx::Early &e = c->early1;
  e.pending = true; e.adopted = false;
  e.a = a; e.b = b; e.r_lo = c->early_lo; e.r_hi = c->early_hi; e.lo = lo; e.n_in = n_in;
  e.o = *o; e.sa = *o->speculate; e.o.speculate = nullptr; e.G = G;
  *queued = 1;
  return SK_OK;
}

// Between sk_subinterval_begin and sk_subinterval_end of a panel's speculated first sub-interval: enqueue the NEXT
// panel's first sub-interval (a2, b2) right behind it, guarded on the device (SkSpec::guard): it runs only if this
// sub-interval turns out accepted (max |I2-I1| < accept_below, no NaN) with nothing converged.  The host then finds
// the next panel already integrated when it gets there (its sk_subinterval[_begin] with exactly these arguments picks
// the launch up) -- the scalar work between two panels costs no device idle time.  *chained = 0: not applicable here
// (sharded run, timing on, sources not prefetched, ...): nothing was enqueued, carry on as usual.
int sk_subinterval_chain(sk_ctx *c, double a2, double b2, const sk_subinterval_opts *o, double accept_below, int32_t *chained) {
  if (!c || !o || !chained) return SK_ERR_ARG;
  *chained = 0;
  if (!c->sub_open || !c->pend_spec || c->chain.pending || !chain_allowed(c) || (sharded(c) && !c->pend_ab) || c->in_group ||
      c->timing || c->interp_mode != 0 ||
      c->family == SK_SDF_HOST || o->speculate == nullptr || (o->kernel != SK_KERNEL_COS && o->kernel != SK_KERNEL_SIN) ||
      o->p != c->p || !(a2 > 0.0) || !(b2 > a2) || !(accept_below > 0.0) || c->pend_rc != SK_OK ||
      o->speculate->criteria < 0 || o->speculate->criteria > 2)
    return SK_OK;
  CK(cudaSetDevice(c->device));
  const long long n_act = c->hi - c->lo, M2 = 2LL * c->m * c->k;
  if (!((M2 * n_act > (1LL << 18)) && n_act > 1)) return SK_OK;
  SkGeom G2;
  if (sk_make_geom(c->plan, a2, b2, c->r_lo, c->r_hi, &G2) != 0) return SK_OK;
  SkPanelSpec S2;
  make_panel_spec(c, a2, b2, o->logw, &S2);
  if (!c->pf_valid || std::memcmp(&c->pf_S, &S2, sizeof(SkPanelSpec)) != 0 || std::memcmp(&c->pf_G, &G2, sizeof(SkGeom)) != 0)
    return SK_OK;                                    // the sources of (a2, b2) are not waiting in the prefetch set
  SkSpec spec;
  std::memset(&spec, 0, sizeof(spec));
  fill_spec(c, o->speculate, spec);
  spec.guard = c->d_red;
  std::memcpy(&spec.guard_maxbits, &accept_below, sizeof(double));
  spec.guard_top = c->hi - 1;
  if (sharded(c)) {                                  // the GLOBAL scalars of the sub-interval in flight (its exchange is queued)
    spec.guard = nullptr;
    spec.gguard = c->d_gout[0];
    std::memcpy(&spec.gguard_rbits, &c->r_hi, sizeof(double));
  }
  {
    int rci = red_init(c, c->d_red2, c->lo - 1);
    if (rci != SK_OK) return rci;
  }
  CK(cudaStreamWaitEvent(c->stream, c->pf_ev, 0));
  swap_src_sets(c);                                  // the prefetched sources / grids become the current set
  c->n_pf_hits++;
  promote_second_prefetch(c);
  c->red_target = c->d_red2;
  const int ksin = o->kernel == SK_KERNEL_SIN;
#define CALL(WW) launch_interp_session<WW>(c, G2, c->uxs.p + c->lo, n_act, o->cmul, ksin, spec)
  DISPATCH_W(c->plan.w, CALL)
#undef CALL
  c->red_target = nullptr;
  LAUNCH_CHECK();
  {
    int rcp = publish(c, c->h_red2, c->d_red2, sizeof(SkReduceOut));
    if (rcp != SK_OK) return rcp;
  }
  if (sharded(c)) {
    // its exchange goes out behind it (slot 4): void if the guard did not hold -- on every rank alike, the guard is global
    int rcx = peer_exchange(c, 4, SK_PX_AB, 0, 0, c->lo, nullptr, 0, 0, nullptr, nullptr, c->d_red2);
    if (rcx != SK_OK) return rcx;
  }
  CK(cudaEventRecord(c->ev_red2, c->stream));
  c->chain.pending = true;
  c->chain.adopted = false;
  c->chain.a = a2; c->chain.b = b2; c->chain.r_lo = c->r_lo; c->chain.r_hi = c->r_hi;
  c->chain.lo = c->lo; c->chain.hi = c->hi;
  c->chain.o = *o;
  c->chain.sa = *o->speculate;
  c->chain.o.speculate = nullptr;
  c->chain_G = G2;
  *chained = 1;
  return SK_OK;
}

// Right after a successful sk_subinterval_chain: enqueue the final gather (sk_results_get_device) behind the chained panel,
// guarded on the device -- it runs only if that panel is accepted and converges EVERY target, i.e. if it is the last panel
// of the run (src/adaptive.jl:149); its truncation segment [lo, hi) (src/adaptive.jl:194) is the one sk_converge_apply will
// register.  The host's sk_results_get_device(vals, errs) then finds the results in place.  *queued = 0: nothing enqueued.
int sk_results_chain_device(sk_ctx *c, double *vals_dev, double *errs_dev, double accept_below, int32_t *queued) {
  if (!c || !vals_dev || !queued) return SK_ERR_ARG;
  *queued = 0;
  if (!c->chain.pending || c->chain.adopted || c->gchain.pending || c->timing || c->tails.n != 0 || !(accept_below > 0.0))
    return SK_OK;
  CK(cudaSetDevice(c->device));
  sk_ctx::GatherChain &g = c->gchain;
  std::memset(&g.tails, 0, sizeof(g.tails));
  if (c->chain.sa.criteria != SK_CRIT_PANEL) {
    SkTailSeg &t = g.tails.seg[0];
    t.lo = c->chain.lo; t.hi = c->chain.hi;
    t.trunc_a = c->chain.sa.trunc_a; t.trunc_num = c->chain.sa.trunc_num; t.xpow = c->chain.sa.xpow;
    t.criteria = c->chain.sa.criteria; t._pad = 0;
    g.tails.n = 1;
  }
  SkGatherGuard gg;
  gg.red = c->d_red2;
  std::memcpy(&gg.maxbits, &accept_below, sizeof(double));
  gg.top = c->chain.lo - 1;
  gg.ran = c->h_gran;                                // (mapped pinned memory: no copy back)
  gg.gen = ++c->gran_gen;
  gg.gl = sharded(c) ? c->d_gout[1] : nullptr;       // (the chained panel's exchange is queued in front of this launch)
  k_gather<<<nblk(c->n_in, 1024), 256, 0, c->stream>>>(c->inv.p, c->res.p, c->n_in, vals_dev, errs_dev, c->in.p, c->in_scale,
                                                       g.tails, gg);
  LAUNCH_CHECK();
  g.vals = vals_dev; g.errs = errs_dev;
  g.pending = true;
  *queued = 1;
  return SK_OK;
}

// sk_subinterval in two halves (built-in densities): _begin enqueues the sub-interval, the host does its scalar work for
// the NEXT panel (tail fit, scan arguments: they depend on the panel ends only), _end waits and returns max |I2 - I1|.
int sk_subinterval_begin(sk_ctx *c, double a, double b, const sk_subinterval_opts *o) {
  if (c) c->sub_open = false;
  int rc = subinterval_builtin_enqueue(c, a, b, o);
  if (rc == SK_OK) c->sub_open = true;
  return rc;
}
int sk_subinterval_end(sk_ctx *c, double *max_abs_diff) {
  if (!c || !max_abs_diff) return fail(c, SK_ERR_ARG, "null pointer");
  if (!c->sub_open) return fail(c, SK_ERR_STATE, "sk_subinterval_begin first");
  c->sub_open = false;
  return transform_and_stage_finish(c, max_abs_diff, nullptr);
}

static int subinterval_host_enqueue(sk_ctx *c, double a, double b, const double *no1, const double *buf1, const double *no2,
                                    const double *buf2, const sk_subinterval_opts *o) {
  int rc = subinterval_prologue(c, a, b, o);
  if (rc != SK_OK) return rc;
  rc = chain_discard(c);
  if (rc != SK_OK) return rc;
  rc = early_discard(c);
  if (rc != SK_OK) return rc;
  if (!no1 || !buf1 || !no2 || !buf2) return fail(c, SK_ERR_ARG, "null pointer");
  const long long M1 = (long long)c->m * c->k;
  CK(cudaMemcpyAsync(c->no1.p, no1, sizeof(double) * M1, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->buf1.p, buf1, sizeof(double) * M1, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->no2.p, no2, sizeof(double) * 2 * M1, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->buf2.p, buf2, sizeof(double) * 2 * M1, cudaMemcpyHostToDevice, c->stream));
  c->have_sources = true;
  c->need_gen = false;
  return transform_and_stage_enqueue(c, a, b, o);
}

int sk_subinterval_host(sk_ctx *c, double a, double b, const double *no1, const double *buf1, const double *no2,
                        const double *buf2, const sk_subinterval_opts *o, double *max_abs_diff) {
  if (c && !max_abs_diff) return fail(c, SK_ERR_ARG, "null pointer");
  int rc = subinterval_host_enqueue(c, a, b, no1, buf1, no2, buf2, o);
  if (rc != SK_OK) return rc;
  return transform_and_stage_finish(c, max_abs_diff, nullptr);
}

// Log-weighted origin sub-interval (src/quadrature.jl:186-228, dim = 1): the host evaluates both
// integrands of the integration by parts (it owns f and df) and the boundary-term coefficient.
static int logw_host_enqueue_local(sk_ctx *c, double a, double b, const double *no1, const double *bufa1, const double *bufb1,
                                   const double *no2, const double *bufa2, const double *bufb2, const sk_subinterval_opts *o,
                                   double i0_coef, double denom) {
  int rc = SK_OK;
  const long long M1 = (long long)c->m * c->k, M2 = 2 * M1;
  const long long n_act = c->hi - c->lo;
  CK(c->bufb1.ensure(M1));
  CK(c->bufb2.ensure(M2));
  c->need_gen = false;
  c->pf_valid = c->pf2_valid = false;
  if (no1) {            // host-evaluated integrands (arbitrary closures: Julia owns f and df)
    CK(cudaMemcpyAsync(c->no1.p, no1, sizeof(double) * M1, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->buf1.p, bufa1, sizeof(double) * M1, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->bufb1.p, bufb1, sizeof(double) * M1, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->no2.p, no2, sizeof(double) * M2, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->buf2.p, bufa2, sizeof(double) * M2, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->bufb2.p, bufb2, sizeof(double) * M2, cudaMemcpyHostToDevice, c->stream));
  } else {              // built-in family: both integrands of the integration by parts on the device (df is closed-form)
    SkPanelSpec S;
    make_panel_spec(c, a, b, 1, &S);
    S.variant = 1;      // f + w log w f'   (src/quadrature.jl:192, :210)
    k_gen_sources<<<nblk(3 * M1, 256), 256, 0, c->stream>>>(S, c->leg_no1.p, c->leg_wt1.p, c->leg_no2.p, c->leg_wt2.p, c->jac_no1.p,
                                                            c->jac_wt1.p, c->jac_no2.p, c->jac_wt2.p, c->no1.p, c->buf1.p, c->no2.p, c->buf2.p);
    LAUNCH_CHECK();
    S.variant = 2;      // w log w f        (:198, :216)
    k_gen_sources<<<nblk(3 * M1, 256), 256, 0, c->stream>>>(S, c->leg_no1.p, c->leg_wt1.p, c->leg_no2.p, c->leg_wt2.p, c->jac_no1.p,
                                                            c->jac_wt1.p, c->jac_no2.p, c->jac_wt2.p, c->no1.p, c->bufb1.p, c->no2.p, c->bufb2.p);
    LAUNCH_CHECK();
  }
  c->have_sources = true;
  {
    int rci = red_init(c, c->d_red, 0);
    if (rci != SK_OK) return rci;
  }
  SkLogwArgs L;
  L.i0_coef = i0_coef;
  L.denom = denom;
  L.b = b;
  const long long n_cut = c->n_act_global > 0 ? c->n_act_global : n_act;
  const bool bessel = o->kernel == SK_KERNEL_BESSEL;                  // dim >= 2: orders nu and nu + 1, :204-221
  const bool fast = !bessel && (M2 * n_cut > (1LL << 18)) && n_cut > 1;          // src/quadrature.jl:105
  if (bessel) {
    if (n_act > 2000000000LL) return fail(c, SK_ERR_ARG, "too many targets");
    CK(c->dsum.ensure((size_t)n_act * 2));
    CK(c->dsumB.ensure((size_t)n_act * 2));
    int hk = SK_ERR_UNSUPPORTED;
    if (c->hankel_mode != 1 && n_cut > 1 && c->r_lo > 0.0 && o->nu + 1 <= SK_HK_NUMAX &&
        (c->hankel_mode == 2 || (M2 * n_cut > (1LL << 29)))) {
      hk = hankel_stage(c, a, b, o->nu, c->buf1.p, c->buf2.p, 1.0, 0.0, n_act, M1, M2, c->dsum.p);
      if (hk == SK_OK) {
        CK(cudaStreamSynchronize(c->stream));     // the pinned group table is rewritten by the second plan
        hk = hankel_stage(c, a, b, o->nu + 1, c->bufb1.p, c->bufb2.p, 1.0, 0.0, n_act, M1, M2, c->dsumB.p);
      }
      if (hk != SK_OK && hk != SK_ERR_UNSUPPORTED) return hk;
    }
    if (hk != SK_OK) {
      dim3 grid((unsigned int)n_act, 2);
      k_direct_bessel<<<grid, 256, 0, c->stream>>>(o->nu, c->no1.p, c->buf1.p, M1, c->no2.p, c->buf2.p, M2, c->uxs.p + c->lo, c->dsum.p);
      LAUNCH_CHECK();
      k_direct_bessel<<<grid, 256, 0, c->stream>>>(o->nu + 1, c->no1.p, c->bufb1.p, M1, c->no2.p, c->bufb2.p, M2, c->uxs.p + c->lo, c->dsumB.p);
      LAUNCH_CHECK();
      c->stats.n_direct++;
    }
    k_bessel_logw_finish<<<nblk(n_act, 256), 256, 0, c->stream>>>(c->dsum.p, c->dsumB.p, c->uxs.p + c->lo, n_act, o->cmul, L,
                                                                  o->nu, o->xdiv_pow, c->stage.p + c->lo, c->d_red);
    LAUNCH_CHECK();
  } else if (fast) {
    SkGeom G;
    if (sk_make_geom(c->plan, a, b, c->r_lo, c->r_hi, &G) != 0) return fail(c, SK_ERR_ARG, "type-3 grid too large");
    rc = run_source_side(c, G, 2, M1, c->buf1.p, nullptr, M2, c->buf2.p, c->fft);
    if (rc != SK_OK) return rc;
    rc = run_source_side(c, G, 2, M1, c->bufb1.p, nullptr, M2, c->bufb2.p, c->fftB);
    if (rc != SK_OK) return rc;
#define CALL(WW)                                                                                                     \
  k_interp_logw<WW><<<nblk(n_act, 256), 256, 0, c->stream>>>(c->plan, G, c->uxs.p + c->lo, n_act, c->fft.p, c->fftB.p, \
                                                             o->cmul, L, c->stage.p + c->lo, c->d_red)
    DISPATCH_W(c->plan.w, CALL)
#undef CALL
    LAUNCH_CHECK();
    c->stats.n_fast++;
  } else {
    CK(c->dsum.ensure((size_t)n_act * 2));
    CK(c->dsumB.ensure((size_t)n_act * 2));
    dim3 grid((unsigned int)n_act, 2);
    k_direct<<<grid, 256, 0, c->stream>>>(c->no1.p, c->buf1.p, M1, c->no2.p, c->buf2.p, M2, c->uxs.p + c->lo, c->dsum.p);
    LAUNCH_CHECK();
    k_direct<<<grid, 256, 0, c->stream>>>(c->no1.p, c->bufb1.p, M1, c->no2.p, c->bufb2.p, M2, c->uxs.p + c->lo, c->dsumB.p);
    LAUNCH_CHECK();
    k_direct_finish_logw<<<1, 256, 0, c->stream>>>(c->dsum.p, c->dsumB.p, c->uxs.p + c->lo, n_act, o->cmul, L,
                                                   c->stage.p + c->lo, c->d_red);
    LAUNCH_CHECK();
    c->stats.n_direct++;
  }
  {
    int rcp = publish(c, &c->h_scal->red, c->d_red, sizeof(SkReduceOut));
    if (rcp != SK_OK) return rcp;
  }
  return SK_OK;
}

int sk_subinterval_logw_host(sk_ctx *c, double a, double b, const double *no1, const double *bufa1, const double *bufb1,
                             const double *no2, const double *bufa2, const double *bufb2, const sk_subinterval_opts *o,
                             double i0_coef, double denom, double *max_abs_diff) {
  int rc = subinterval_prologue(c, a, b, o);
  if (rc != SK_OK) return rc;
  rc = chain_discard(c);
  if (rc != SK_OK) return rc;
  rc = early_discard(c);
  if (rc != SK_OK) return rc;
  if (!max_abs_diff) return fail(c, SK_ERR_ARG, "null pointer");
  const bool all_null = !no1 && !bufa1 && !bufb1 && !no2 && !bufa2 && !bufb2;
  if (!all_null && (!no1 || !bufa1 || !bufb1 || !no2 || !bufa2 || !bufb2)) return fail(c, SK_ERR_ARG, "null pointer");
  if (all_null && (c->family == SK_SDF_HOST || c->deriv != 0))
    return fail(c, SK_ERR_STATE, "device-evaluated log-weighted integrands need a built-in density (sk_sdf_builtin, deriv_index 0)");
  if (a != 0.0) return fail(c, SK_ERR_ARG, "the integration-by-parts branch applies to a == 0 only");
  const long long n_act = c->hi - c->lo;
  rc = logw_host_enqueue_local(c, a, b, no1, bufa1, bufb1, no2, bufa2, bufb2, o, i0_coef, denom);
  if (sharded(c)) {       // a rank that failed locally still joins the collective (its peers are waiting in it)
    const std::string msg = c->errmsg;
    c->peer_slot = 0;
    c->cur_red = c->d_red;
    c->pend_ab = false;
    c->pend_rc = rc;
    const int rcc = comm_reduce_a(c, 0, rc != SK_OK);
    if (rcc != SK_OK) return rcc;
    CK(cudaStreamSynchronize(c->stream));
    if (rc != SK_OK) { c->errmsg = msg; return rc; }
    int prc = peer_check(c, 0);
    if (prc != SK_OK) return prc;
    prc = peer_resend_while_void(c);
    if (prc != SK_OK) return prc;
    if (global_a(c).err) return fail(c, SK_ERR_STATE, "another rank failed in this sub-interval");
  } else {
    if (rc != SK_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
  }
  unsigned int fl = c->h_scal->red.flags;
  double mx;
  std::memcpy(&mx, &c->h_scal->red.maxbits, sizeof(double));
  comm_take_a(c, &fl, &mx);
  if (fl & SK_FLAG_NAND) mx = std::nan("");
  *max_abs_diff = mx;
  c->g_max_abs = mx;
  c->staged = true;
  c->panel_subs++;
  c->stats.n_subintervals++;
  c->stats.units += n_act;
  if (!(fl & SK_FLAG_NAN1) && (fl & SK_FLAG_NAN2)) return fail(c, SK_ERR_NAN, "NaN detected in panel integral...");
  return SK_OK;
}

int sk_sources_get(sk_ctx *c, int32_t rule, double *no, double *buf) {
  if (!c || !no || !buf || rule < 0 || rule > 1) return SK_ERR_ARG;
  if (!c->have_sources) return fail(c, SK_ERR_STATE, "no sub-interval evaluated yet");
  const long long M = (long long)c->m * c->k * (rule ? 2 : 1);
  CK(cudaMemcpyAsync(no, rule ? c->no2.p : c->no1.p, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(buf, rule ? c->buf2.p : c->buf1.p, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return SK_OK;
}

int sk_subinterval_accept(sk_ctx *c) {
  if (!c) return SK_ERR_ARG;
  if (!c->staged) return fail(c, SK_ERR_STATE, "no staged sub-interval");
  const long long n = c->hi - c->lo;
  if (c->spec_active && !c->spec_accepted) {
    // the interpolation kernel already applied ks += I2, errs += |I2-I1| (sk_subinterval_opts::speculate)
    c->spec_accepted = true;
    c->first_accept = false;
    c->staged = false;
    c->stats.n_accepted++;
    return SK_OK;
  }
  if (c->first_accept) {
    // I = 0 + I2, err = 0 + |I2-I1| exactly (src/quadrature.jl:174-175, :261-262): the staging buffer
    // becomes the panel buffer, no pass over the data
    std::swap(c->pan, c->stage);
  } else {
    k_accept_add<<<nblk(n, 256), 256, 0, c->stream>>>(c->pan.p + c->lo, c->stage.p + c->lo, n);
    LAUNCH_CHECK();
  }
  c->first_accept = false;
  c->staged = false;
  c->stats.n_accepted++;
  return SK_OK;
}

int sk_panel_commit(sk_ctx *c) {
  NvtxRange nvtx("add panels to integral");
  if (!c) return SK_ERR_ARG;
  if (!c->in_panel) return fail(c, SK_ERR_STATE, "no open panel");
  const long long n = c->hi - c->lo;
  if (c->spec_active && c->spec_accepted) {   // already in (ks, errs)
    c->stats.n_panels++;
    return SK_OK;
  }
  if (c->first_accept)   // nothing was accepted: I = err = 0
    CK(cudaMemsetAsync(c->pan.p + c->lo, 0, sizeof(sk_cplx) * n, c->stream));
  // ks += I; errs += err (src/adaptive.jl:163-164) is deferred and fused with the convergence scan
  // (k_commit_scan); any other consumer flushes it first.
  c->commit_pending = true;
  c->pend_lo = c->lo;
  c->pend_hi = c->hi;
  c->stats.n_panels++;
  return SK_OK;
}

// the scan in two halves (see transform_and_stage_enqueue): enqueue the fused commit + predicate pass (nothing at all
// when the panel was committed speculatively), then read the reduced scalars back
static int converge_scan_enqueue(sk_ctx *c, const sk_scan_args *a) {
  NvtxRange nvtx("check convergence");
  if (!c || !a) return SK_ERR_ARG;
  if (!c->in_panel) return fail(c, SK_ERR_STATE, "no open panel");
  if (a->criteria < 0 || a->criteria > 2) return fail(c, SK_ERR_ARG, "bad criteria");
  CK(cudaSetDevice(c->device));
  c->scan_from_spec = false;
  if (c->spec_active && c->spec_accepted) {
    const sk_scan_args &s0 = c->spec_args;
    const bool same = s0.criteria == a->criteria && std::memcmp(&s0.trunc_a, &a->trunc_a, sizeof(double)) == 0 &&
                      std::memcmp(&s0.trunc_num, &a->trunc_num, sizeof(double)) == 0 && s0.xpow == a->xpow && s0.tau == a->tau;
    if (!same) return fail(c, SK_ERR_STATE, "scan arguments differ from the ones the panel was speculated with");
    c->scan_from_spec = true;
    if (sharded(c) && !c->pend_ab) {   // sharded run: the scan is a collective point for every rank
      unsigned long long rb = 0ull;
      std::memcpy(&rb, &c->spec_r, sizeof(double));
      const long long nlb = c->spec_new_hi - c->lo > 0 ? c->spec_new_hi - c->lo : 0;
      int rcc = comm_reduce_b(c, false, rb, nlb);
      if (rcc != SK_OK) return rcc;
    }
    return SK_OK;
  }
  const long long n = c->hi - c->lo;
  {
    int rci = red_init(c, c->d_red, c->lo - 1);
    if (rci != SK_OK) return rci;
  }
  const int do_commit = (c->commit_pending && c->pend_lo == c->lo && c->pend_hi == c->hi) ? 1 : 0;
  if (c->commit_pending && !do_commit) {
    int rc = flush_commit(c);
    if (rc != SK_OK) return rc;
  }
  {
    int rcz = ensure_res_zero(c);
    if (rcz != SK_OK) return rcz;
  }
  k_commit_scan<<<nblk(n, 256), 256, 0, c->stream>>>(c->uxs.p + c->lo, c->pan.p + c->lo, c->res.p + c->lo, n, c->lo, do_commit,
                                                     a->trunc_a, a->trunc_num, a->xpow, a->tau, a->criteria, c->d_red);
  LAUNCH_CHECK();
  c->commit_pending = false;
  {
    int rcp = publish(c, &c->h_scal->red, c->d_red, sizeof(SkReduceOut));
    if (rcp != SK_OK) return rcp;
  }
  return comm_reduce_b(c, true, 0ull, 0);
}

static int converge_scan_finish(sk_ctx *c, int64_t *new_hi, double *r_at_new_hi) {
  if (c->scan_from_spec) {
    *new_hi = c->spec_new_hi;
    if (r_at_new_hi) *r_at_new_hi = c->spec_r;
    c->scan_hi = c->spec_new_hi;
    c->scan_r = c->spec_r;
    if (sharded(c) && !c->pend_ab) {
      CK(cudaStreamSynchronize(c->stream));
      const int prc = peer_check(c, 0);
      if (prc != SK_OK) return prc;
      comm_take_b(c);
    }
    return SK_OK;
  }
  CK(cudaStreamSynchronize(c->stream));
  comm_take_b(c);
  const long long top = c->h_scal->red.max_unconv;      // 0-based index of the highest unconverged target, or lo-1
  double r = 0.0;
  if (top >= c->lo) std::memcpy(&r, &c->h_scal->red.rbits, sizeof(double));
  *new_hi = top + 1;                                      // 1-based
  if (r_at_new_hi) *r_at_new_hi = r;
  c->scan_hi = top + 1;
  c->scan_r = r;
  return SK_OK;
}

int sk_converge_scan(sk_ctx *c, const sk_scan_args *a, int64_t *new_hi, double *r_at_new_hi) {
  if (!c || !a || !new_hi) return SK_ERR_ARG;
  int rc = converge_scan_enqueue(c, a);
  if (rc != SK_OK) return rc;
  return converge_scan_finish(c, new_hi, r_at_new_hi);
}

int sk_converge_apply(sk_ctx *c, const sk_scan_args *a, int64_t new_hi) {
  if (!c || !a) return SK_ERR_ARG;
  if (!c->in_panel) return fail(c, SK_ERR_STATE, "no open panel");
  if (new_hi < c->lo || new_hi > c->hi) return fail(c, SK_ERR_ARG, "new_hi outside the panel range");
  int rc = flush_commit(c);
  if (rc != SK_OK) return rc;
  const long long nconv = c->hi - new_hi;                 // 0-based indices new_hi .. hi-1
  if (nconv > 0 && a->criteria != SK_CRIT_PANEL) {
    if (c->tails.n < SK_MAX_TAILS) {
      // errs[ix] += 2 trunc_err (src/adaptive.jl:194) rides along with the final gather: no pass of its own
      SkTailSeg &t = c->tails.seg[c->tails.n++];
      t.lo = new_hi; t.hi = c->hi;
      t.trunc_a = a->trunc_a; t.trunc_num = a->trunc_num; t.xpow = a->xpow;
      t.criteria = a->criteria; t._pad = 0;
    } else {
      rc = ensure_res_zero(c);
      if (rc != SK_OK) return rc;
      k_scan_add<<<nblk(nconv, 256), 256, 0, c->stream>>>(c->uxs.p + new_hi, c->res.p + new_hi, nconv, a->trunc_a, a->trunc_num,
                                                          a->xpow, a->criteria);
      LAUNCH_CHECK();
    }
  }
  c->in_panel = false;
  c->spec_active = false;
  return SK_OK;
}

int sk_target_upper_index(sk_ctx *c, double r, int64_t *idx) {
  if (!c || !idx) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "no targets set");
  k_upper_bound<<<1, 1, 0, c->stream>>>(c->uxs.p, c->n_unique, r, &c->d_red->max_unconv);
  LAUNCH_CHECK();
  {
    int rcp = publish(c, &c->h_scal->red, c->d_red, sizeof(SkReduceOut));
    if (rcp != SK_OK) return rcp;
  }
  CK(cudaStreamSynchronize(c->stream));
  *idx = c->h_scal->red.max_unconv;
  return SK_OK;
}

int sk_results_get_device(sk_ctx *c, double *vals_dev, double *errs_dev) {
  if (!c || !vals_dev) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "no targets set");
  if (c->gchain.pending) {
    // a gather was enqueued behind the last panel (sk_results_chain_device): if it ran and is the gather this call would
    // do -- same arrays, same truncation segments, nothing pending -- the results are already in place
    c->gchain.pending = false;
    const bool same = vals_dev == c->gchain.vals && errs_dev == c->gchain.errs && !c->chain.pending && !c->early1.pending &&
                      !c->commit_pending && !c->res_zero_pending && !(c->spec_active && !c->spec_accepted) && !c->timing &&
                      c->tails.n == c->gchain.tails.n &&
                      std::memcmp(c->tails.seg, c->gchain.tails.seg, sizeof(SkTailSeg) * (size_t)c->tails.n) == 0;
    CK(cudaStreamSynchronize(c->stream));
    if (same && *c->h_gran == c->gran_gen) {
      c->stats.n_chained++;
      return SK_OK;
    }
  }
  int rc = rollback_speculation(c);
  if (rc != SK_OK) return rc;
  rc = chain_discard(c);
  if (rc != SK_OK) return rc;
  rc = early_discard(c);
  if (rc != SK_OK) return rc;
  rc = flush_commit(c);
  if (rc != SK_OK) return rc;
  rc = ensure_res_zero(c);
  if (rc != SK_OK) return rc;
  if (c->timing) CK(cudaEventRecord(c->ev[0], c->stream));
  k_gather<<<nblk(c->n_in, 1024), 256, 0, c->stream>>>(c->inv.p, c->res.p, c->n_in, vals_dev, errs_dev, c->in.p, c->in_scale, c->tails);
  LAUNCH_CHECK();
  if (c->timing) CK(cudaEventRecord(c->ev[1], c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (c->timing) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.gather_ms += ms;
  }
  return SK_OK;
}

static int results_enqueue(sk_ctx *c, double *vals, double *errs) {
  NvtxRange nvtx("scatter to input order");
  if (!c || !vals) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "no targets set");
  CK(cudaSetDevice(c->device));
  int rc = rollback_speculation(c);
  if (rc != SK_OK) return rc;
  rc = chain_discard(c);
  if (rc != SK_OK) return rc;
  rc = early_discard(c);
  if (rc != SK_OK) return rc;
  rc = flush_commit(c);
  if (rc != SK_OK) return rc;
  rc = ensure_res_zero(c);
  if (rc != SK_OK) return rc;
  CK(c->out_v.ensure(c->n_in));
  if (errs) CK(c->out_e.ensure(c->n_in));
  if (c->timing) CK(cudaEventRecord(c->ev[0], c->stream));
  // Large outputs are gathered and copied in slices: the copy of slice k (stream2) overlaps the gather of slice k+1
  // (compute stream), so the PCIe link is busy from the first slice on.  (stream2 is otherwise the prefetch stream and
  // idle at this point.)
  const long long n = c->n_in;
  const int nslice = n >= (1LL << 21) ? SK_GATHER_SLICES : 1;
  for (int k = 0; k < nslice; ++k) {
    const long long o = (n * k / nslice) & ~1023LL, e = k + 1 == nslice ? n : ((n * (k + 1) / nslice) & ~1023LL);
    if (e <= o) continue;
    k_gather<<<nblk(e - o, 1024), 256, 0, c->stream>>>(c->inv.p + o, c->res.p, e - o, c->out_v.p + o,
                                                        errs ? c->out_e.p + o : nullptr, c->in.p + o, c->in_scale, c->tails);
    LAUNCH_CHECK();
    if (nslice == 1) {
      if (c->timing) CK(cudaEventRecord(c->ev[1], c->stream));
      CK(cudaMemcpyAsync(vals, c->out_v.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
      if (errs) CK(cudaMemcpyAsync(errs, c->out_e.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    } else {
      if (!c->ev_slice[k]) CK(cudaEventCreateWithFlags(&c->ev_slice[k], cudaEventDisableTiming));
      CK(cudaEventRecord(c->ev_slice[k], c->stream));
      CK(cudaStreamWaitEvent(c->stream2, c->ev_slice[k], 0));
      CK(cudaMemcpyAsync(vals + o, c->out_v.p + o, sizeof(double) * (e - o), cudaMemcpyDeviceToHost, c->stream2));
      if (errs) CK(cudaMemcpyAsync(errs + o, c->out_e.p + o, sizeof(double) * (e - o), cudaMemcpyDeviceToHost, c->stream2));
    }
  }
  if (nslice > 1 && c->timing) CK(cudaEventRecord(c->ev[1], c->stream));
  c->results_sliced = nslice > 1;
  return SK_OK;
}
static int results_finish(sk_ctx *c) {
  CK(cudaStreamSynchronize(c->stream));
  if (c->results_sliced) CK(cudaStreamSynchronize(c->stream2));
  if (c->timing) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.gather_ms += ms;
  }
  return SK_OK;
}

int sk_results_get(sk_ctx *c, double *vals, double *errs) {
  int rc = results_enqueue(c, vals, errs);
  if (rc != SK_OK) return rc;
  return results_finish(c);
}

// Batched evaluations (hyperparameter sweeps over the same targets): gather on the compute stream, copy on a second
// stream while the next run computes.
int sk_results_get_async(sk_ctx *c, double *vals, double *errs) {
  if (!c || !vals) return SK_ERR_ARG;
  if (!c->have_targets) return fail(c, SK_ERR_STATE, "no targets set");
  CK(cudaSetDevice(c->device));
  int rc = rollback_speculation(c);
  if (rc != SK_OK) return rc;
  rc = chain_discard(c);
  if (rc != SK_OK) return rc;
  rc = early_discard(c);
  if (rc != SK_OK) return rc;
  rc = flush_commit(c);
  if (rc != SK_OK) return rc;
  rc = ensure_res_zero(c);
  if (rc != SK_OK) return rc;
  if (!c->copy_stream) {
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&c->ev_gather[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
    }
  }
  const int slot = c->aslot;
  c->aslot ^= 1;
  if (c->copy_pending[slot]) {                 // the copy issued two runs ago from this slot
    CK(cudaEventSynchronize(c->ev_copy[slot]));
    c->copy_pending[slot] = false;
  }
  CK(c->aout_v[slot].ensure(c->n_in));
  if (errs) CK(c->aout_e[slot].ensure(c->n_in));
  k_gather<<<nblk(c->n_in, 1024), 256, 0, c->stream>>>(c->inv.p, c->res.p, c->n_in, c->aout_v[slot].p,
                                                       errs ? c->aout_e[slot].p : nullptr, c->in.p, c->in_scale, c->tails);
  LAUNCH_CHECK();
  CK(cudaEventRecord(c->ev_gather[slot], c->stream));
  CK(cudaStreamWaitEvent(c->copy_stream, c->ev_gather[slot], 0));
  CK(cudaMemcpyAsync(vals, c->aout_v[slot].p, sizeof(double) * c->n_in, cudaMemcpyDeviceToHost, c->copy_stream));
  if (errs) CK(cudaMemcpyAsync(errs, c->aout_e[slot].p, sizeof(double) * c->n_in, cudaMemcpyDeviceToHost, c->copy_stream));
  CK(cudaEventRecord(c->ev_copy[slot], c->copy_stream));
  c->copy_pending[slot] = true;
  return SK_OK;
}

int sk_results_wait(sk_ctx *c) {
  if (!c) return SK_ERR_ARG;
  if (c->copy_stream) CK(cudaStreamSynchronize(c->copy_stream));
  c->copy_pending[0] = c->copy_pending[1] = false;
  return SK_OK;
}

int sk_ctx_set_interp_mode(sk_ctx *c, int mode) {
  if (!c || mode < 0 || mode > 2) return SK_ERR_ARG;
  c->interp_mode = mode;
  return SK_OK;
}

int sk_ctx_set_hankel_mode(sk_ctx *c, int mode) {
  if (!c || mode < 0 || mode > 2) return SK_ERR_ARG;
  c->hankel_mode = mode;
  return SK_OK;
}

int sk_stats_get(sk_ctx *c, sk_stats *out) {
  if (!c || !out) return SK_ERR_ARG;
  c->stats.n_prefetch_issued = c->n_pf_issued;
  c->stats.n_prefetch_hits = c->n_pf_hits;
  c->stats.launches_total = c->launches_total;
  *out = c->stats;
  return SK_OK;
}

// ---- device group: ONE caller, several GPUs -------------------------------------------------------------------
// The reference's API is a single task calling kernel_values(cfg, xs) (src/adaptive.jl:95-108).  A group gives that
// caller N devices behind the same sequence of calls: the distances are cut into N contiguous chunks, every device
// sorts / de-duplicates its own chunk and runs the same adaptive loop on it; each step is enqueued on all devices
// first and read back afterwards, so one host thread keeps N GPUs (and N PCIe links for the upload of the distances
// and the download of the results) busy.  The only exchange is the host-side max of the per-device scalars (max |I2-I1|,
// NaN flags, largest unconverged distance) -- no NCCL, no peer copies.  Every device builds the transform geometry from
// the GLOBAL distance range (sk_panel_set_range), so values and error estimates are bit-identical to a one-device run
// over the same distances.  Indices crossing this API (ix1, hi, new_hi) count over the concatenation of the devices'
// unique tables; equal distances in different chunks count once per chunk.
struct sk_group {
  std::vector<sk_ctx *> ctx;
  std::string errmsg;
  int nuse = 0;                                   // devices holding a chunk (n_in may be smaller than the group)
  long long n_in = 0;
  std::vector<long long> off, cnt, ix1, hi, new_hi;
  std::vector<char> active;
  std::vector<sk_target_info> info;
  bool have_targets = false, has_zero = false, in_panel = false;
  double r_min_pos = 0, r_max = 0, r_hi_g = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
};

namespace {
int gfail(sk_group *g, int dev, int rc) {
  if (g && dev >= 0 && dev < (int)g->ctx.size())
    g->errmsg = "device " + std::to_string(g->ctx[dev]->device) + ": " + g->ctx[dev]->errmsg;
  return rc;
}
int gfailmsg(sk_group *g, int rc, const char *msg) {
  if (g) g->errmsg = msg;
  return rc;
}
long long group_global_hi(const sk_group *g) {
  long long h = g->has_zero ? 1 : 0;
  for (int i = 0; i < g->nuse; ++i) h += std::max<long long>(g->hi[i] - (g->ix1[i] - 1), 0);
  return h;
}
}  // namespace

int sk_group_create(const int32_t *devices, int32_t ndev, sk_group **out) {
  if (!out || !devices || ndev < 1) return SK_ERR_ARG;
  *out = nullptr;
  sk_group *g = new sk_group();
  for (int i = 0; i < ndev; ++i) {            // (a device may be listed more than once: two chunks on one GPU)
    sk_ctx *c = nullptr;
    int rc = sk_ctx_create(devices[i], &c);
    if (rc != SK_OK) { sk_group_destroy(g); return rc; }
    c->in_group = ndev > 1;               // chunks: the first panel's geometry comes from the global range
    g->ctx.push_back(c);
  }
  const size_t n = g->ctx.size();
  g->off.assign(n, 0); g->cnt.assign(n, 0); g->ix1.assign(n, 1); g->hi.assign(n, 0); g->new_hi.assign(n, 0);
  g->active.assign(n, 0);
  g->info.resize(n);
  *out = g;
  return SK_OK;
}

int sk_group_destroy(sk_group *g) {
  if (!g) return SK_OK;
  for (sk_ctx *c : g->ctx) sk_ctx_destroy(c);
  delete g;
  return SK_OK;
}

int sk_group_size(const sk_group *g) { return g ? (int)g->ctx.size() : 0; }
const char *sk_group_last_error(const sk_group *g) { return g ? g->errmsg.c_str() : "null group"; }

int sk_group_ctx(sk_group *g, int32_t i, sk_ctx **out) {
  if (!g || !out || i < 0 || i >= (int)g->ctx.size()) return SK_ERR_ARG;
  *out = g->ctx[i];
  return SK_OK;
}

int sk_group_set_timing(sk_group *g, int enabled) {
  if (!g) return SK_ERR_ARG;
  for (sk_ctx *c : g->ctx) sk_ctx_set_timing(c, enabled);
  return SK_OK;
}

int sk_group_set_nufft_eps(sk_group *g, double eps) {
  if (!g) return SK_ERR_ARG;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    int rc = sk_ctx_set_nufft_eps(g->ctx[i], eps);
    if (rc != SK_OK) return gfail(g, (int)i, rc);
  }
  return SK_OK;
}

int sk_group_synchronize(sk_group *g) {
  if (!g) return SK_ERR_ARG;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    cudaSetDevice(g->ctx[i]->device);
    int rc = sk_ctx_synchronize(g->ctx[i]);
    if (rc != SK_OK) return gfail(g, (int)i, rc);
  }
  return SK_OK;
}

int sk_group_rule_set(sk_group *g, int32_t m, int32_t k, double p, const double *leg_no1, const double *leg_wt1,
                      const double *leg_no2, const double *leg_wt2, const double *jac_no1, const double *jac_wt1,
                      const double *jac_no2, const double *jac_wt2) {
  if (!g) return SK_ERR_ARG;
  // device 0 generates (or receives) the rules; the others receive device 0's, so that all devices integrate with
  // bit-identical nodes and weights
  int rc = sk_rule_set(g->ctx[0], m, k, p, leg_no1, leg_wt1, leg_no2, leg_wt2, jac_no1, jac_wt1, jac_no2, jac_wt2);
  if (rc != SK_OK) return gfail(g, 0, rc);
  sk_ctx *c0 = g->ctx[0];
  for (size_t i = 1; i < g->ctx.size(); ++i) {
    sk_ctx *c = g->ctx[i];
    if (c->have_rule && c->m == m && c->k == k && c->p == p && c->h_rule[0] == c0->h_rule[0] && c->h_rule[4] == c0->h_rule[4] &&
        c->h_rule[1] == c0->h_rule[1] && c->h_rule[5] == c0->h_rule[5])
      continue;
    const bool jac = p != 0.0;
    rc = sk_rule_set(c, m, k, p, c0->h_rule[0].data(), c0->h_rule[1].data(), c0->h_rule[2].data(), c0->h_rule[3].data(),
                     jac ? c0->h_rule[4].data() : nullptr, jac ? c0->h_rule[5].data() : nullptr,
                     jac ? c0->h_rule[6].data() : nullptr, jac ? c0->h_rule[7].data() : nullptr);
    if (rc != SK_OK) return gfail(g, (int)i, rc);
  }
  return SK_OK;
}

int sk_group_rule_get(sk_group *g, int32_t which, double *no, double *wt) {
  if (!g) return SK_ERR_ARG;
  int rc = sk_rule_get(g->ctx[0], which, no, wt);
  return rc == SK_OK ? rc : gfail(g, 0, rc);
}

int sk_group_sdf_builtin(sk_group *g, int32_t family, const double *params, int32_t nparams, int32_t deriv_index) {
  if (!g) return SK_ERR_ARG;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    int rc = sk_sdf_builtin(g->ctx[i], family, params, nparams, deriv_index);
    if (rc != SK_OK) return gfail(g, (int)i, rc);
  }
  return SK_OK;
}

int sk_group_targets_set(sk_group *g, const double *xs_host, int64_t n_in, sk_target_info *info) {
  if (!g || !xs_host || n_in < 1) return gfailmsg(g, SK_ERR_ARG, "need at least one distance");
  g->have_targets = false;
  const int nd = (int)g->ctx.size();
  g->nuse = (int)std::min<long long>(nd, n_in);
  g->n_in = n_in;
  // upload + sort of all chunks enqueued first (N PCIe links when xs_host is pinned memory: sk_host_alloc), read back after
  for (int i = 0; i < g->nuse; ++i) {
    sk_ctx *c = g->ctx[i];
    g->off[i] = (long long)i * n_in / g->nuse;
    g->cnt[i] = (long long)(i + 1) * n_in / g->nuse - g->off[i];
    if (cudaSetDevice(c->device) != cudaSuccess || c->in.ensure(g->cnt[i]) != cudaSuccess ||
        cudaMemcpyAsync(c->in.p, xs_host + g->off[i], sizeof(double) * g->cnt[i], cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
      c->errmsg = "upload of the distances failed";
      return gfail(g, i, SK_ERR_CUDA);
    }
    int rc = targets_enqueue(c, g->cnt[i]);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  {
    // every chunk's key range is in long before its sort ends: the first panel's source side -- built from the GLOBAL
    // range, as sk_panel_set_range will ask for -- starts on every device while the sorts run
    double lo_g = 0.0, hi_g = 0.0;
    long long n_tot = 0;
    for (int i = 0; i < g->nuse; ++i) {
      double lo = 0, hi = 0;
      int rc = targets_early_range(g->ctx[i], &lo, &hi);
      if (rc != SK_OK) return gfail(g, i, rc);
      if (hi > 0.0) {
        if (lo_g == 0.0 || lo < lo_g) lo_g = lo;
        if (hi > hi_g) hi_g = hi;
      }
      n_tot += g->cnt[i];
    }
    if (hi_g > 0.0 && n_tot >= 65536)
      for (int i = 0; i < g->nuse; ++i) {
        int rc = targets_early_prefetch(g->ctx[i], lo_g, hi_g);
        if (rc != SK_OK) return gfail(g, i, rc);
      }
  }
  g->has_zero = false;
  g->r_min_pos = 0.0;
  g->r_max = 0.0;
  long long nu = 0;
  for (int i = 0; i < g->nuse; ++i) {
    sk_ctx *c = g->ctx[i];
    cudaSetDevice(c->device);
    int rc = targets_finish(c, g->cnt[i], &g->info[i], false);
    if (rc != SK_OK) return gfail(g, i, rc);
    const sk_target_info &t = g->info[i];
    g->ix1[i] = t.has_zero ? 2 : 1;
    g->hi[i] = t.n_unique;
    if (t.has_zero) g->has_zero = true;
    if (t.r_min_pos > 0 && (g->r_min_pos == 0.0 || t.r_min_pos < g->r_min_pos)) g->r_min_pos = t.r_min_pos;
    if (t.r_max > g->r_max) g->r_max = t.r_max;
    nu += t.n_unique - (t.has_zero ? 1 : 0);
  }
  g->r_hi_g = g->r_max;
  g->have_targets = true;
  g->in_panel = false;
  if (info) {
    info->n_in = n_in;
    info->n_unique = nu + (g->has_zero ? 1 : 0);
    info->has_zero = g->has_zero ? 1 : 0;
    info->_pad = 0;
    info->r_min_pos = g->r_min_pos;
    info->r_max = g->r_max;
  }
  return SK_OK;
}

int sk_group_run_begin(sk_group *g) {
  if (!g) return SK_ERR_ARG;
  if (!g->have_targets) return gfailmsg(g, SK_ERR_STATE, "sk_group_targets_set first");
  for (int i = 0; i < g->nuse; ++i) {
    int rc = sk_run_begin(g->ctx[i]);
    if (rc != SK_OK) return gfail(g, i, rc);
    g->ix1[i] = g->info[i].has_zero ? 2 : 1;
    g->hi[i] = g->info[i].n_unique;
  }
  g->r_hi_g = g->r_max;
  g->in_panel = false;
  return SK_OK;
}

int sk_group_zero_lag_set(sk_group *g, double value) {
  if (!g) return SK_ERR_ARG;
  for (int i = 0; i < g->nuse; ++i) {
    cudaSetDevice(g->ctx[i]->device);
    int rc = sk_zero_lag_set(g->ctx[i], value);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  return SK_OK;
}

int sk_group_panel_begin(sk_group *g, int64_t ix1, int64_t hi, double *r_lo, double *r_hi) {
  if (!g) return SK_ERR_ARG;
  if (!g->have_targets) return gfailmsg(g, SK_ERR_STATE, "sk_group_targets_set first");
  if (ix1 != (g->has_zero ? 2 : 1) || hi != group_global_hi(g))
    return gfailmsg(g, SK_ERR_ARG, "a group follows the active range of its own scans: pass ix1 / hi as returned");
  long long n_act = 0;
  for (int i = 0; i < g->nuse; ++i) {
    g->active[i] = g->hi[i] >= g->ix1[i];
    if (g->active[i]) n_act += g->hi[i] - g->ix1[i] + 1;
  }
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    sk_ctx *c = g->ctx[i];
    int rc = sk_panel_begin(c, g->ix1[i], g->hi[i], nullptr, nullptr);
    if (rc != SK_OK) return gfail(g, i, rc);
    if (g->ctx.size() > 1) {
      rc = sk_panel_set_range(c, g->r_min_pos, g->r_hi_g, n_act);
      if (rc != SK_OK) return gfail(g, i, rc);
    }
  }
  if (r_lo) *r_lo = g->r_min_pos;
  if (r_hi) *r_hi = g->r_hi_g;
  g->in_panel = true;
  return SK_OK;
}

namespace {
// read the staged scalars of all active devices back and combine them (src/quadrature.jl:258, :165)
int group_subinterval_finish(sk_group *g, double *max_abs_diff) {
  double mx = 0.0;
  unsigned int fl = 0;
  bool nan = false;
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    cudaSetDevice(g->ctx[i]->device);
    double m1 = 0.0;
    unsigned int f1 = 0;
    int rc = transform_and_stage_finish(g->ctx[i], &m1, &f1);
    if (rc != SK_OK) return gfail(g, i, rc);
    fl |= f1;
    if (m1 != m1) nan = true; else mx = std::max(mx, m1);
  }
  *max_abs_diff = nan ? std::nan("") : mx;
  if (!(fl & SK_FLAG_NAN1) && (fl & SK_FLAG_NAN2)) return gfailmsg(g, SK_ERR_NAN, "NaN detected in panel integral...");
  return SK_OK;
}
}  // namespace

int sk_group_subinterval(sk_group *g, double a, double b, const sk_subinterval_opts *o, double *max_abs_diff) {
  if (!g || !o || !max_abs_diff) return gfailmsg(g, SK_ERR_ARG, "null argument");
  if (!g->in_panel) return gfailmsg(g, SK_ERR_STATE, "sk_group_panel_begin first");
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    int rc = subinterval_builtin_enqueue(g->ctx[i], a, b, o);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  return group_subinterval_finish(g, max_abs_diff);
}

int sk_group_subinterval_host(sk_group *g, double a, double b, const double *no1, const double *buf1, const double *no2,
                              const double *buf2, const sk_subinterval_opts *o, double *max_abs_diff) {
  if (!g || !o || !max_abs_diff) return gfailmsg(g, SK_ERR_ARG, "null argument");
  if (!g->in_panel) return gfailmsg(g, SK_ERR_STATE, "sk_group_panel_begin first");
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    int rc = subinterval_host_enqueue(g->ctx[i], a, b, no1, buf1, no2, buf2, o);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  return group_subinterval_finish(g, max_abs_diff);
}

int sk_group_subinterval_accept(sk_group *g) {
  if (!g) return SK_ERR_ARG;
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    cudaSetDevice(g->ctx[i]->device);
    int rc = sk_subinterval_accept(g->ctx[i]);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  return SK_OK;
}

int sk_group_panel_commit(sk_group *g) {
  if (!g) return SK_ERR_ARG;
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    cudaSetDevice(g->ctx[i]->device);
    int rc = sk_panel_commit(g->ctx[i]);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  return SK_OK;
}

int sk_group_converge_scan(sk_group *g, const sk_scan_args *a, int64_t *new_hi, double *r_at_new_hi) {
  if (!g || !a || !new_hi) return SK_ERR_ARG;
  if (!g->in_panel) return gfailmsg(g, SK_ERR_STATE, "no open panel");
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    int rc = converge_scan_enqueue(g->ctx[i], a);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  std::vector<double> r(g->nuse, 0.0);
  double r_g = 0.0;
  for (int i = 0; i < g->nuse; ++i) {
    g->new_hi[i] = g->hi[i];
    if (!g->active[i]) continue;
    cudaSetDevice(g->ctx[i]->device);
    int64_t nh = 0;
    int rc = converge_scan_finish(g->ctx[i], &nh, &r[i]);
    if (rc != SK_OK) return gfail(g, i, rc);
    g->new_hi[i] = nh;
    r_g = std::max(r_g, r[i]);
  }
  // the reference walks down from the largest distance while converged: everything up to the largest unconverged
  // distance of ANY device stays active on every device
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i] || !(r_g > r[i])) continue;
    cudaSetDevice(g->ctx[i]->device);
    int64_t nh = 0;
    int rc = sk_target_upper_index(g->ctx[i], r_g, &nh);
    if (rc != SK_OK) return gfail(g, i, rc);
    g->new_hi[i] = std::min<long long>(nh, g->hi[i]);
  }
  long long h = g->has_zero ? 1 : 0;
  for (int i = 0; i < g->nuse; ++i) h += std::max<long long>(g->new_hi[i] - (g->ix1[i] - 1), 0);
  *new_hi = h;
  if (r_at_new_hi) *r_at_new_hi = r_g;
  g->r_hi_g = r_g;
  return SK_OK;
}

int sk_group_converge_apply(sk_group *g, const sk_scan_args *a, int64_t new_hi) {
  if (!g || !a) return SK_ERR_ARG;
  if (!g->in_panel) return gfailmsg(g, SK_ERR_STATE, "no open panel");
  long long h = g->has_zero ? 1 : 0;
  for (int i = 0; i < g->nuse; ++i) h += std::max<long long>(g->new_hi[i] - (g->ix1[i] - 1), 0);
  if (new_hi != h) return gfailmsg(g, SK_ERR_ARG, "new_hi is not the value sk_group_converge_scan returned");
  for (int i = 0; i < g->nuse; ++i) {
    if (!g->active[i]) continue;
    cudaSetDevice(g->ctx[i]->device);
    int rc = sk_converge_apply(g->ctx[i], a, g->new_hi[i]);
    if (rc != SK_OK) return gfail(g, i, rc);
    g->hi[i] = g->new_hi[i];
  }
  g->in_panel = false;
  return SK_OK;
}

int sk_group_results_get(sk_group *g, double *vals, double *errs) {
  if (!g || !vals) return SK_ERR_ARG;
  if (!g->have_targets) return gfailmsg(g, SK_ERR_STATE, "no targets set");
  for (int i = 0; i < g->nuse; ++i) {
    int rc = results_enqueue(g->ctx[i], vals + g->off[i], errs ? errs + g->off[i] : nullptr);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  for (int i = 0; i < g->nuse; ++i) {
    cudaSetDevice(g->ctx[i]->device);
    int rc = results_finish(g->ctx[i]);
    if (rc != SK_OK) return gfail(g, i, rc);
  }
  return SK_OK;
}

int sk_group_stats_get(sk_group *g, sk_stats *out) {
  if (!g || !out) return SK_ERR_ARG;
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < std::max(g->nuse, 1); ++i) {
    const sk_stats &s = g->ctx[i]->stats;
    // per-step counters are the same on every device that took part: report the largest; work adds up
    out->n_subintervals = std::max(out->n_subintervals, s.n_subintervals);
    out->n_accepted = std::max(out->n_accepted, s.n_accepted);
    out->n_panels = std::max(out->n_panels, s.n_panels);
    out->n_fast = std::max(out->n_fast, s.n_fast);
    out->n_direct = std::max(out->n_direct, s.n_direct);
    out->n_hankel = std::max(out->n_hankel, s.n_hankel);
    out->n_speculated = std::max(out->n_speculated, s.n_speculated);
    out->n_spec_rollbacks = std::max(out->n_spec_rollbacks, s.n_spec_rollbacks);
    out->units += s.units;
    out->kernel_launches += s.kernel_launches;
    out->last_nf = s.last_nf; out->last_nf2 = s.last_nf2;
    out->interp_ms = std::max(out->interp_ms, s.interp_ms);
    out->source_ms = std::max(out->source_ms, s.source_ms);
    out->sort_ms = std::max(out->sort_ms, s.sort_ms);
    out->gather_ms = std::max(out->gather_ms, s.gather_ms);
    out->n_prefetch_issued += g->ctx[i]->n_pf_issued;
    out->n_prefetch_hits += g->ctx[i]->n_pf_hits;
    out->timing_enabled = s.timing_enabled;
    out->sort_two_level = i == 0 ? s.sort_two_level : std::min(out->sort_two_level, s.sort_two_level);
  }
  return SK_OK;
}

// ---- host-side helpers exported for tests ---------------------------------------------------------------
int sk_host_gauss_rule(int32_t n, double p, double *no, double *wt) {
  if (!no || !wt) return SK_ERR_ARG;
  return sk_plan_gauss_rule(n, p, no, wt) == 0 ? SK_OK : SK_ERR_ARG;
}

int sk_host_es_plan(int32_t w, double *beta, int32_t *nc, double *coef, int32_t *nq, double *qc, double *ximax) {
  SkEsPlan P;
  if (sk_plan_make_es(w, &P) != 0) return SK_ERR_ARG;
  if (beta) *beta = P.beta;
  if (nc) *nc = SK_NC;
  if (nq) *nq = P.nq;
  if (ximax) *ximax = P.ximax;
  if (coef)
    for (int i = 0; i < w / 2; ++i)
      for (int q = 0; q < SK_NC / 2; ++q) {
        coef[(i * 2 + 0) * (SK_NC / 2) + q] = P.E[i][q];
        coef[(i * 2 + 1) * (SK_NC / 2) + q] = P.O[i][q];
      }
  if (qc) std::memcpy(qc, P.qc, sizeof(double) * P.nq);
  return SK_OK;
}

}  // extern "C"
