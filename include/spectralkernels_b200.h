/*
 * spectralkernels_b200.h -- C ABI of the B200 (sm_100a) evaluator for the K(r) hot path of
 * pbeckman/SpectralKernels.jl.
 *
 * Plain C: `extern "C"`, pointers and sizes only, no torch / CUDA types in any signature.
 * Every function returns an int status (SK_OK == 0, negative == error), never throws, and
 * never calls back into the host language.  All host buffers are owned by the caller and
 * must stay alive for the duration of the call; all device memory is owned by the sk_ctx.
 * One sk_ctx must be used by one host thread at a time (the reference's config is not
 * re-entrant either: src/adaptive.jl:17-21, src/quadrature.jl:170).
 *
 * There are two levels.  Each entry point cites the reference interface it replaces
 * (paths are into the reference repository).
 *
 *   Level 0  sk_nufft1d3            == finufft1d3(w, s, x)             src/utils.jl:10
 *   Level 1  device-resident session that absorbs every O(N) pass of
 *            kernel_values / _kernel_values / fourier_integrate_interval /
 *            fourier_integrate_panel / updatequadbufs!; the host keeps only the scalar
 *            control flow of src/adaptive.jl:149-200 and src/quadrature.jl:181-272.
 *
 * INTEGRATION.md shows the Julia `ccall` bindings for each entry point.
 */
#ifndef SPECTRALKERNELS_B200_H
#define SPECTRALKERNELS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_ABI_VERSION 2

/* status codes */
#define SK_OK 0
#define SK_ERR_CUDA (-1)        /* a CUDA runtime call failed (sk_last_error has the text)      */
#define SK_ERR_ARG (-2)         /* invalid argument                                            */
#define SK_ERR_STATE (-3)       /* call sequence violated (e.g. sub-interval before targets)   */
#define SK_ERR_NAN (-4)         /* NaN in the 2m-rule panel integral, src/quadrature.jl:165     */
#define SK_ERR_SPLIT (-5)       /* b - a <= 1e-16, src/utils.jl:28-36 (check_subdivide_failure) */
#define SK_ERR_ALLOC (-6)       /* out of device / pinned memory                               */
#define SK_ERR_CUFFT (-7)       /* cuFFT failure                                               */
#define SK_ERR_UNSUPPORTED (-8) /* a reference branch this build does not cover                */
#define SK_ERR_INPUT (-9)       /* targets contain NaN / negative / non-finite values          */

/* integral kernels, src/quadrature.jl:176-180 and :130-136 */
#define SK_KERNEL_COS 0
#define SK_KERNEL_SIN 1
#define SK_KERNEL_BESSEL 2   /* (:J, nu), dim >= 2 (even): sum_k c_k J_nu(2 pi w_k r), src/quadrature.jl:137-161, :179 */

/* convergence criteria, src/adaptive.jl:12, :231-233 */
#define SK_CRIT_PANEL 0
#define SK_CRIT_TAILS 1
#define SK_CRIT_BOTH 2

/* built-in spectral-density families (evaluated on the device) */
#define SK_SDF_HOST 0        /* strengths are supplied by the host (sk_subinterval_host)               */
#define SK_SDF_MATERN 1      /* params (phi, rho, nu, d): phi*(rho^2+w^2)^(-nu-d/2), scripts/matern_pair.jl:17 */
#define SK_SDF_EXPONENTIAL 2 /* params (phi, alpha):      phi*exp(-alpha*|w|), test/derivatives/jacobian.jl:5 */

typedef struct sk_ctx sk_ctx;

typedef struct sk_target_info {
  int64_t n_in;      /* number of input distances                                          */
  int64_t n_unique;  /* number of unique distances (length of the sorted unique set)       */
  int32_t has_zero;  /* 1 if the smallest unique distance is 0 (src/adaptive.jl:133)       */
  int32_t _pad;
  double r_min_pos;  /* smallest strictly positive distance (0 if none)                    */
  double r_max;      /* largest distance                                                   */
} sk_target_info;

typedef struct sk_scan_args {
  double trunc_a;    /* -c/(d+dim)*b^(d+dim), first bound of src/adaptive.jl:225-228 (target independent) */
  double trunc_num;  /* c*b^(d+(dim-1)/2), numerator of the second bound                                 */
  double xpow;       /* (dim+1)/2, exponent of x in the second bound                                     */
  double tau;        /* config.tol*abs(k0)/2, src/adaptive.jl:191                                        */
  int32_t criteria;  /* SK_CRIT_*; the host switches to PANEL after a NaN tail fit (src/adaptive.jl:170-175) */
  int32_t _pad;
} sk_scan_args;

typedef struct sk_subinterval_opts {
  double cmul;       /* config.c, src/adaptive.jl:43-45, applied at src/quadrature.jl:250-251        */
  double p;          /* config.p, src/adaptive.jl:42                                                */
  int32_t kernel;    /* SK_KERNEL_COS / SK_KERNEL_SIN, src/quadrature.jl:177                        */
  int32_t logw;      /* config.logw: multiply the integrand by log(w), src/quadrature.jl:242        */
  int32_t nu;        /* Bessel order for SK_KERNEL_BESSEL: dim/2-1, or dim/2 for the derivative (:179)   */
  int32_t _pad;
  double xdiv_pow;   /* dim/2 - 1: the integrals are divided by x^xdiv_pow (src/quadrature.jl:252-254); 0 for dim = 1 */
  /* Optional (may be NULL).  On the FIRST sub-interval of a panel -- the whole panel [a,b] -- the host may
   * pass the scan arguments it will use for this panel (they depend on (a,b) only: estimate_tail_decay,
   * src/adaptive.jl:204-220).  The interpolation kernel then also applies ks += I2, errs += |I2-I1| and
   * evaluates the convergence predicate, so that accept / commit / scan need no further pass over the
   * targets if the sub-interval is accepted; if it is rejected the update is rolled back bit for bit. */
  const sk_scan_args *speculate;
} sk_subinterval_opts;

typedef struct sk_stats {
  int64_t n_subintervals;   /* sub-intervals evaluated since sk_run_begin                     */
  int64_t n_accepted;
  int64_t n_panels;         /* outer panels committed                                         */
  int64_t units;            /* sum over sub-intervals of N_active  (SURVEY section 8d "unit")  */
  int64_t n_fast;           /* sub-intervals that took the NUFFT branch                        */
  int64_t n_direct;         /* sub-intervals that took the direct-summation branch             */
  int64_t kernel_launches;  /* kernels of this library launched since sk_run_begin             */
  int64_t last_nf;          /* spread-grid size of the last NUFFT                              */
  int64_t last_nf2;         /* FFT size of the last NUFFT                                      */
  int64_t n_speculated;     /* sub-intervals whose commit/scan was fused into the interpolation kernel */
  int64_t n_spec_rollbacks; /* ... of which were rejected and rolled back                      */
  double interp_ms;         /* device time in the interpolation kernel since sk_run_begin      */
  double source_ms;         /* device time in node/strength/spread/FFT since sk_run_begin      */
  int32_t timing_enabled;
  int32_t sort_two_level;   /* last sk_targets_set*: 2 = input was already sorted and unique (no sort), 1 = the bin scheme of K8 (csrc/sk_k8.cuh), 0 = the general sort (clustered / heavily duplicated input) */
  int64_t n_hankel;         /* sub-intervals that took the O(N) nonuniform Hankel transform (dim >= 2) */
  double sort_ms;           /* device time of the last sk_targets_set* (unique / sort / inverse map), timing enabled */
  double gather_ms;         /* device time in the gather to the input order since sk_run_begin  */
  int64_t n_prefetch_issued;/* sub-intervals whose source side (nodes, strengths, spread, FFT) was computed ahead on the
                               second stream, since the context was created ...                 */
  int64_t n_prefetch_hits;  /* ... and how many of them the driver then actually asked for      */
  int64_t n_chained;        /* launches that were enqueued ahead of time (sk_first_panel_early, sk_subinterval_chain,
                               sk_results_chain_device) and picked up, since sk_run_begin      */
  int64_t launches_total;   /* kernels of this library launched since the context was created (the sort of
                               sk_targets_set* runs before sk_run_begin resets the other counters) */
} sk_stats;

/* ---- library --------------------------------------------------------------------------------- */
int sk_abi_version(void);
const char *sk_error_string(int code);
const char *sk_last_error(const sk_ctx *ctx);            /* detail of the last failure on ctx        */

int sk_ctx_create(int device, sk_ctx **out);
int sk_ctx_destroy(sk_ctx *ctx);
int sk_ctx_set_timing(sk_ctx *ctx, int enabled);          /* per-stage cudaEvent timers (NVTX-like)   */
int sk_ctx_set_nufft_eps(sk_ctx *ctx, double eps);        /* default 1e-15, as hard-wired in src/utils.jl:10 */
/* interpolation kernel: 0 (default) = cell polynomials in shared memory (k_interp_cells); 1 = per-target
 * exp-of-semicircle taps, the textbook evaluation (k_interp_session), kept for A/B measurements */
int sk_ctx_set_interp_mode(sk_ctx *ctx, int mode);
/* dim >= 2 (SK_KERNEL_BESSEL), replaces FastHankelTransform.jl's nufht (src/quadrature.jl:139-143):
 * 0 (default) = O(N) nonuniform Hankel transform when more than ~4096 targets are active, the reference's
 * direct Bessel summation (src/quadrature.jl:145-160) below that; 1 = always the direct summation;
 * 2 = always the O(N) transform (orders 0..3) */
int sk_ctx_set_hankel_mode(sk_ctx *ctx, int mode);
int sk_ctx_synchronize(sk_ctx *ctx);
/* the CUDA stream of the context as an opaque handle (cudaStream_t), for event timing by the caller */
int sk_ctx_stream(sk_ctx *ctx, void **stream_out);

/* device-side stopwatch on the context's stream (cudaEvent pair), the equivalent of the reference's
 * TimerOutputs sections (src/SpectralKernels.jl:14): begin records, end records + waits + returns ms */
int sk_timer_begin(sk_ctx *ctx);
int sk_timer_end(sk_ctx *ctx, double *ms);
/* measured FP64 FMA throughput of this GPU (dependent-chain DFMA micro-benchmark, TFLOP/s); the
 * denominator of the FP64 roofline, which MEASURED_PEAKS.json does not contain */
int sk_fp64_peak(sk_ctx *ctx, double *tflops, double *ms);

/* ---- target-sharded multi-GPU (one process per GPU) ---------------------------------------------------
 * Every rank holds its own chunk of the distances; the ranks run ONE adaptive loop in lock step.  With a
 * communicator the scalar reductions of that loop run as NCCL all-reduces on the context's stream, right
 * behind the kernel that produced the local value: sk_subinterval* then return the maximum of max|I2-I1|
 * over all ranks, and sk_converge_scan is a collective point too (sk_comm_last gives the global stopping
 * distance and the summed lower bound of the active counts).  NCCL is dlopen'ed (libnccl.so.2) on first use.
 * sk_comm_unique_id: rank 0 creates the 128-byte id, the host broadcasts it (MPI / torch.distributed / file). */
int sk_comm_unique_id(void *out128);
int sk_comm_init(sk_ctx *ctx, const void *id128, int32_t rank, int32_t nranks);
int sk_comm_destroy(sk_ctx *ctx);
/* all-reduce of up to 32 host doubles (op: 0 max, 1 min, 2 sum); synchronous; no-op without a communicator */
int sk_comm_allreduce(sk_ctx *ctx, double *vals, int32_t n, int32_t op);
/* a rank whose active set is empty joins the others' collective points: which = 0 sub-interval, 1 scan,
 * 2 a sub-interval that carries the scan's scalars too -- the first sub-interval of a panel when opts.speculate is
 * given and the NUFFT branch is taken (2 m k n_active_global > 2^18, kernel cos / sin): its collective then also
 * reduces (stopping distance, active count), and when that sub-interval is accepted the scan is no collective point
 * any more (an idle rank skips its which = 1 call for that panel).  A rank that fails locally still joins the
 * collective with an error word set, so that every rank returns an error instead of blocking. */
int sk_comm_idle(sk_ctx *ctx, int32_t which);
int sk_comm_last(sk_ctx *ctx, double *max_abs_diff, double *r_stop, int64_t *n_active_lb);

/* ---- the same collectives over NVLink / NVSwitch PEER MEMORY instead of NCCL ---------------------------------------
 * Every rank owns a 2 KB mailbox in its HBM which all its peers map through CUDA IPC.  A collective point is then ONE
 * single-warp kernel on the compute stream behind the kernel that produced the local scalars (k_peer_exchange): it packs
 * them, stores 64 bytes into every peer's mailbox, waits (bounded: SK_PEER_TIMEOUT_S, default 20 s) for the peers'
 * words in its own mailbox, reduces them and writes the result into pinned host memory -- no NCCL launch, no pack
 * kernel, no D2H copy; the host still synchronises once per sub-interval, exactly as on one GPU.  With mailboxes
 * attached no NCCL communicator is needed at all (sk_comm_init may be skipped); every sk_subinterval*, sk_converge_scan,
 * sk_comm_idle, sk_comm_allreduce and the early range reduction of sk_targets_set then use the mailboxes.
 *   sk_comm_peer_export  allocate this rank's mailbox, return its 64-byte cudaIpcMemHandle_t
 *   (the host all-gathers the handles: torch.distributed / MPI / a file)
 *   sk_comm_peer_attach  handles = nranks x 64 bytes in rank order; at most 16 ranks, all on one node
 *   sk_comm_allgather    out[r * k + i] = value i (k <= 7) of rank r; synchronous (the once-per-call range / counts)
 * sk_comm_peer_selftest runs the protocol with the ranks emulated as blocks of one cooperative launch on ONE device
 * (out5: per rank maxbits, err, rbits, n_lb, status | void << 1 | epoch << 8; skip_rank >= 0: that rank's last exchange is
 * that of a chained launch which skipped itself, so every rank must see it void) -- a test hook, not part of the path. */
int sk_comm_peer_export(sk_ctx *ctx, void *handle64);
int sk_comm_peer_attach(sk_ctx *ctx, const void *handles, int32_t rank, int32_t nranks);
int sk_comm_allgather(sk_ctx *ctx, const double *vals, int32_t k, double *out);
/* With mailboxes attached, sk_targets_set* / sk_targets_end also deliver what the ranks exchange at the start of a run
 * (src/adaptive.jl:123, :152): the smallest positive and the largest distance over all ranks and the summed number of
 * positive unique distances -- sent behind the sort's summary kernel, so it costs no synchronisation of its own.
 * *valid = 0: not available (some rank's sort took a path whose summary the host recomputes); every rank sees the same
 * answer and then exchanges the three numbers with sk_comm_allgather. */
int sk_comm_summary(sk_ctx *ctx, int32_t *valid, double *r_lo, double *r_hi, int64_t *n_active);
int sk_comm_peer_selftest(sk_ctx *ctx, int32_t nranks, int32_t rounds, const uint64_t *maxbits_in, const uint64_t *rbits_in,
                          const int64_t *top_in, int64_t lo, int32_t skip_rank, uint64_t *out5);

/* pinned host memory for callers that want full PCIe rate (Julia: unsafe_wrap the pointer) */
int sk_host_alloc(size_t bytes, void **out);
int sk_host_free(void *ptr);

/* ---- Level 0: drop-in for finufft1d3(w, s, x), src/utils.jl:10 -------------------------------- */
/* out[j] = sum_k s[k] * exp(+i * 2*pi * x[j] * w[k]);  s and out are interleaved (re, im);        */
/* all pointers are HOST pointers; synchronous.  eps <= 0 selects the context default.             */
int sk_nufft1d3(sk_ctx *ctx, int64_t M, const double *w, const double *s, int64_t N, const double *x,
                double *out, double eps);

/* ---- Level 1: quadrature rules, QuadRule src/quadrature.jl:27-47 ------------------------------- */
/* (m, k) = quadspec; p = config.p.  Node/weight pointers are HOST arrays of length m (…1) and    */
/* 2m (…2) on [-1,1], ascending; pass NULL for all eight to let the library generate them          */
/* (Gauss-Legendre, and Gauss-Jacobi(0,p) when p != 0).                                            */
int sk_rule_set(sk_ctx *ctx, int32_t m, int32_t k, double p,
                const double *leg_no1, const double *leg_wt1, const double *leg_no2, const double *leg_wt2,
                const double *jac_no1, const double *jac_wt1, const double *jac_no2, const double *jac_wt2);
/* which: 0 legendre m, 1 legendre 2m, 2 jacobi m, 3 jacobi 2m; copies to HOST arrays */
int sk_rule_get(sk_ctx *ctx, int32_t which, double *no, double *wt);

/* ---- Level 1: integrand ------------------------------------------------------------------------ */
/* deriv_index 0 = S itself; j >= 1 = dS/dparams[j-1] (the integrands of src/derivatives.jl:63-72) */
int sk_sdf_builtin(sk_ctx *ctx, int32_t family, const double *params, int32_t nparams, int32_t deriv_index);

/* ---- Level 1: targets, replaces unique/sort/Dict of src/adaptive.jl:99-107, :113-120 ----------- */
int sk_targets_set(sk_ctx *ctx, const double *xs_host, int64_t n_in, sk_target_info *info);
int sk_targets_set_device(sk_ctx *ctx, const double *xs_dev, int64_t n_in, sk_target_info *info);
/* sk_targets_set[_device] in two halves.  _begin enqueues the upload and the sort and returns once the first pass over the
 * distances has delivered their range (sk_targets_early_range; r_hi = 0: not known before the sort ends).  The first
 * panel is (0, m k / (2 r_max)) (src/adaptive.jl:152), so the host evaluates estimate_tail_decay (src/adaptive.jl:204-220)
 * and the scan arguments for it while the device sorts; _end waits for the sort.  The device buffer of _begin_device
 * must stay valid until _end returns. */
int sk_targets_begin(sk_ctx *ctx, const double *xs_host, int64_t n_in);
int sk_targets_begin_device(sk_ctx *ctx, const double *xs_dev, int64_t n_in);
int sk_targets_early_range(sk_ctx *ctx, double *r_lo, double *r_hi);
int sk_targets_end(sk_ctx *ctx, sk_target_info *info);
/* Between _begin and _end: enqueue the FIRST panel's first sub-interval (0, b1), b1 = m k / (2 r_max)
 * (src/adaptive.jl:152), behind the sort -- its ends, its transform geometry and whether there is an r = 0 row are known
 * after the first pass; the number of unique distances is read by the kernel from the sort's device-side summary.
 * opts.speculate must carry the scan arguments of (0, b1).  The sk_run_begin / sk_panel_begin / sk_subinterval[_begin]
 * (0, b1, opts) that follow sk_targets_end find the panel already being integrated (sk_stats.n_chained); anything
 * else discards the launch.  *queued = 0: not applicable (sharded run, timing on, host-evaluated density, dim >= 2,
 * small input): nothing was enqueued. */
int sk_first_panel_early(sk_ctx *ctx, double a, double b, const sk_subinterval_opts *opts, int32_t *queued);
/* lags of point pairs computed on the device (src/model.jl:53-68 with NoWarping: lag = norm(pts[i] - pts[j])):
 * pts_host is npts x dim row-major; pairs_host holds npairs 0-based (i, j) index pairs, or NULL for all
 * npts (npts-1) / 2 pairs i < j in row-major order of the strict upper triangle.  The results of
 * sk_results_get are then in pair order.  Replaces the host-side lag list and its upload. */
int sk_targets_set_pairs(sk_ctx *ctx, const double *pts_host, int64_t npts, int32_t dim, const int64_t *pairs_host,
                         int64_t npairs, sk_target_info *info);
/* Linear warping of the lags, warp(params, x) = x / rho (src/model.jl:62-66; the range parameter of
 * scripts/fit_vecchia_demo.jl:15): every unique distance becomes (original distance) * factor, factor > 0.  Order,
 * uniqueness and the inverse map are unchanged, so a fitting loop re-uses the sort of sk_targets_set* for every
 * range value.  The factor always applies to the distances as they were set.  info (may be NULL) receives the scaled
 * r_min_pos / r_max. */
int sk_targets_scale(sk_ctx *ctx, double factor, sk_target_info *info);
/* sorted unique value at 1-based index idx */
int sk_target_value(sk_ctx *ctx, int64_t idx, double *out);

/* ---- Level 1: the adaptive loop's device steps ------------------------------------------------- */
/* ks = errs = 0 (src/adaptive.jl:122) and reset statistics */
int sk_run_begin(sk_ctx *ctx);
/* ks[1] = value, errs[1] = NaN when the first unique distance is 0 (src/adaptive.jl:133-146) */
int sk_zero_lag_set(sk_ctx *ctx, double value);
/* start an outer panel over the 1-based inclusive index range [ix1, hi] (src/adaptive.jl:157-159,
 * src/quadrature.jl:174-175: I = err = 0); returns the smallest / largest active distance */
int sk_panel_begin(sk_ctx *ctx, int64_t ix1, int64_t hi, double *r_lo, double *r_hi);
/* override the distance range the transform geometry is built for (default: the panel's own
 * [r_lo, r_hi]).  A target-sharded multi-GPU run passes the GLOBAL range so that every rank uses the
 * same grids and per-target results do not depend on the sharding.  n_active_global (> 0) is the number of
 * active targets over all ranks: the NUFFT-vs-direct-summation cutoff (src/quadrature.jl:105, src/utils.jl:39)
 * is then decided on the global count, as a single-GPU run over the union would. */
int sk_panel_set_range(sk_ctx *ctx, double r_lo, double r_hi, int64_t n_active_global);
/* one pass of the bisection loop body, src/quadrature.jl:183-258: build both rules on [a,b]
 * (updatequadbufs!, :49-95) from the built-in S, transform (fast or direct, :105-128), select
 * Re/Im, scale by cmul, stage I2 and |I2-I1|, and return max|I2-I1| (NaN if any is NaN). */
int sk_subinterval(sk_ctx *ctx, double a, double b, const sk_subinterval_opts *opts, double *max_abs_diff);
/* sk_subinterval in two halves: _begin enqueues, _end waits and returns max |I2 - I1|.  In between the host does its
 * scalar work for the next panel (its ends are known: src/adaptive.jl:152), which then costs no device idle time. */
int sk_subinterval_begin(sk_ctx *ctx, double a, double b, const sk_subinterval_opts *opts);
int sk_subinterval_end(sk_ctx *ctx, double *max_abs_diff);
/* Between _begin and _end of a panel's speculated first sub-interval: enqueue the NEXT panel's first sub-interval
 * (a2, b2) = (b, b + m k / (2 r_hi)) (src/adaptive.jl:152) behind it, guarded ON THE DEVICE: the kernel runs only if
 * this sub-interval turns out accepted -- max |I2-I1| < accept_below = tol * k0, no NaN (src/quadrature.jl:260) -- and
 * its scan converged nothing (so r_hi, and with it (a2, b2), is what the host will compute); otherwise it returns at
 * once having touched nothing.  opts.speculate must carry the scan arguments of (a2, b2).  When the host reaches that
 * panel, its sk_subinterval[_begin] with exactly these arguments finds the work done (sk_stats.n_chained); any other
 * call discards the launch (and rolls it back if it ran).  *chained = 0: not applicable (sharded run, timing on,
 * sources of (a2, b2) not prefetched, host-evaluated density, dim >= 2): nothing was enqueued. */
int sk_subinterval_chain(sk_ctx *ctx, double a2, double b2, const sk_subinterval_opts *opts, double accept_below,
                         int32_t *chained);
/* Right after a successful sk_subinterval_chain: enqueue the final gather of sk_results_get_device behind the chained panel,
 * guarded on the device: it runs only if that panel is accepted and converges EVERY target, i.e. if it turns out to be the
 * last panel of the run (src/adaptive.jl:149).  sk_results_get_device(vals, errs) with the same arrays then finds the
 * results in place; otherwise it gathers as usual.  *queued = 0: nothing was enqueued. */
int sk_results_chain_device(sk_ctx *ctx, double *vals_dev, double *errs_dev, double accept_below, int32_t *queued);
/* same with host-evaluated nodes and (real) strengths: no1/buf1 length m*k, no2/buf2 length 2*m*k
 * (the buffers of src/adaptive.jl:50-53 after updatequadbufs!) */
int sk_subinterval_host(sk_ctx *ctx, double a, double b, const double *no1, const double *buf1,
                        const double *no2, const double *buf2, const sk_subinterval_opts *opts,
                        double *max_abs_diff);
/* the log-weighted origin sub-interval (config.logw && a == 0 && p != 0, src/quadrature.jl:186-228, dim = 1):
 * integration by parts with two :cis transforms per rule.  bufa* = wt * (f + w log w f')(no*), bufb* =
 * wt * (w log w f)(no*) built by updatequadbufs! with p = config.p (Jacobi first sub-panel);
 * i0_coef = b^(dim/2+1-alpha) * log(b) * f(b), denom = dim - alpha.  Stages
 * I_k = (I0 - Re A_k + 2 pi x Im B_k) / denom * cmul like sk_subinterval.
 * With opts->kernel == SK_KERNEL_BESSEL (dim >= 2, :204-221) the two transforms are Bessel sums of orders
 * opts->nu = dim/2-1 (A) and opts->nu + 1 (B), I0 carries J_nu(2 pi b x) and the result is divided by
 * x^xdiv_pow: I_k = (I0 - A_k + 2 pi x B_k) / denom * cmul / x^(dim/2-1).
 * All six array pointers NULL: a built-in density is set (sk_sdf_builtin, deriv_index 0) and the device evaluates both
 * integrands itself (dS/dw of the shipped families is closed-form): nothing is evaluated on or uploaded from the host. */
int sk_subinterval_logw_host(sk_ctx *ctx, double a, double b, const double *no1, const double *bufa1,
                             const double *bufb1, const double *no2, const double *bufa2, const double *bufb2,
                             const sk_subinterval_opts *opts, double i0_coef, double denom, double *max_abs_diff);
/* copy the nodes / strengths the last sk_subinterval generated to HOST arrays (parity tests) */
int sk_sources_get(sk_ctx *ctx, int32_t rule /*0: m-rule, 1: 2m-rule*/, double *no, double *buf);
/* I += I2; err += |I2-I1| for the staged sub-interval, src/quadrature.jl:260-262 */
int sk_subinterval_accept(sk_ctx *ctx);
/* ks += I; errs += err over the panel range, src/adaptive.jl:163-164 */
int sk_panel_commit(sk_ctx *ctx);
/* convergence scan from hi downwards, src/adaptive.jl:183-199, search part: returns the highest
 * unconverged 1-based index (ix1-1 if every active target converged) and the distance at that index
 * (0 if none).  No side effects. */
int sk_converge_scan(sk_ctx *ctx, const sk_scan_args *args, int64_t *new_hi, double *r_at_new_hi);
/* side-effect part: errs[ix] += 2*trunc_err for new_hi < ix <= hi (src/adaptive.jl:194) and close the
 * panel.  Single GPU: pass the new_hi sk_converge_scan returned.  Target-sharded multi-GPU: pass the
 * local index of the GLOBAL stopping distance (sk_target_upper_index of the max over ranks). */
int sk_converge_apply(sk_ctx *ctx, const sk_scan_args *args, int64_t new_hi);
/* largest 1-based index whose sorted unique distance is <= r (0 if none) */
int sk_target_upper_index(sk_ctx *ctx, double r, int64_t *idx);
/* values and errors in the ORIGINAL input order, duplicates included (src/adaptive.jl:105-107);
 * HOST arrays of length n_in; errs may be NULL */
int sk_results_get(sk_ctx *ctx, double *vals, double *errs);
/* Asynchronous variant for batched evaluations over the same targets (hyperparameter sweeps of a fitting loop, the
 * P_sdf + 2 derivative runs of src/derivatives.jl:86-112): the results are gathered on the compute stream and copied
 * to the (pinned) HOST arrays on a second stream while the next run computes.  Two copies can be in flight; the host
 * arrays must stay alive and untouched until sk_results_wait returns. */
int sk_results_get_async(sk_ctx *ctx, double *vals, double *errs);
int sk_results_wait(sk_ctx *ctx);
/* same into DEVICE arrays (no PCIe traffic) */
int sk_results_get_device(sk_ctx *ctx, double *vals_dev, double *errs_dev);
int sk_stats_get(sk_ctx *ctx, sk_stats *out);

/* ---- device group: ONE caller, several GPUs -------------------------------------------------------------
 * The reference's API is a single task calling kernel_values(cfg, xs) (src/adaptive.jl:95-108); a group gives that
 * caller N devices behind the same sequence of calls as a single context (sk_group_X has the contract of sk_X).
 * The distances are cut into N contiguous chunks of the caller's array, one per device; each device sorts and
 * de-duplicates its chunk (src/adaptive.jl:99, :113-120) and runs the adaptive loop on it in lock step with the
 * others.  Every call enqueues its work on all devices before it reads anything back, so the devices -- and, with
 * host arrays from sk_host_alloc, their PCIe links -- work concurrently under one host thread.  The only exchange is
 * the host-side maximum of the per-device scalars (max |I2-I1| of src/quadrature.jl:258, the NaN flags of :165, the
 * largest unconverged distance of src/adaptive.jl:183-198): no collective library is involved.  All devices build the
 * transform geometry from the global distance range, so values and error estimates equal a one-device run over the
 * same distances bit for bit.  ix1 / hi / new_hi count over the concatenation of the devices' unique tables (equal
 * distances in different chunks count once per chunk); pass them back as returned.  A device may be listed more
 * than once (several chunks on one GPU).                                                                          */
typedef struct sk_group sk_group;
int sk_group_create(const int32_t *devices, int32_t ndev, sk_group **out);
int sk_group_destroy(sk_group *g);
int sk_group_size(const sk_group *g);
int sk_group_ctx(sk_group *g, int32_t i, sk_ctx **out);          /* the i-th device's context (stays owned by g) */
const char *sk_group_last_error(const sk_group *g);
int sk_group_set_timing(sk_group *g, int enabled);
int sk_group_set_nufft_eps(sk_group *g, double eps);
int sk_group_synchronize(sk_group *g);
int sk_group_rule_set(sk_group *g, int32_t m, int32_t k, double p, const double *leg_no1, const double *leg_wt1,
                      const double *leg_no2, const double *leg_wt2, const double *jac_no1, const double *jac_wt1,
                      const double *jac_no2, const double *jac_wt2);
int sk_group_rule_get(sk_group *g, int32_t which, double *no, double *wt);
int sk_group_sdf_builtin(sk_group *g, int32_t family, const double *params, int32_t nparams, int32_t deriv_index);
int sk_group_targets_set(sk_group *g, const double *xs_host, int64_t n_in, sk_target_info *info);
int sk_group_run_begin(sk_group *g);
int sk_group_zero_lag_set(sk_group *g, double value);
int sk_group_panel_begin(sk_group *g, int64_t ix1, int64_t hi, double *r_lo, double *r_hi);
int sk_group_subinterval(sk_group *g, double a, double b, const sk_subinterval_opts *opts, double *max_abs_diff);
int sk_group_subinterval_host(sk_group *g, double a, double b, const double *no1, const double *buf1, const double *no2,
                              const double *buf2, const sk_subinterval_opts *opts, double *max_abs_diff);
int sk_group_subinterval_accept(sk_group *g);
int sk_group_panel_commit(sk_group *g);
int sk_group_converge_scan(sk_group *g, const sk_scan_args *args, int64_t *new_hi, double *r_at_new_hi);
int sk_group_converge_apply(sk_group *g, const sk_scan_args *args, int64_t new_hi);
int sk_group_results_get(sk_group *g, double *vals, double *errs);
int sk_group_stats_get(sk_group *g, sk_stats *out);               /* counters of the step; work and launches summed */

/* ---- host-side plan helpers exported for tests (no GPU needed) --------------------------------- */
/* Gauss-Legendre (p == 0) / Gauss-Jacobi(0,p) on [-1,1], ascending; 0 on success */
int sk_host_gauss_rule(int32_t n, double p, double *no, double *wt);
/* exp-of-semicircle plan: tap polynomial coefficients coef[w/2][2][nc/2] (even, odd parts in s^2,
 * s = 2x), Chebyshev coefficients qc[nq] of (2/w)/phihat(xi) in tau = 2 (xi/ximax)^2 - 1 */
int sk_host_es_plan(int32_t w, double *beta, int32_t *nc, double *coef, int32_t *nq, double *qc, double *ximax);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRALKERNELS_B200_H */
