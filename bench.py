#!/usr/bin/env python
"""
bench.py -- headline benchmark of the K(r) hot path (BASELINE.json):

    K(r) evals/sec at tol = 1e-8 (Matern S, 1e7 r), 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, restated (oracle port)

A "step" is one kernel_values(cfg, rs) call over one batch of synthetic distances:
Matern nu = 1.5, rho = 1, phi such that K(0) = 1; rs ~ U(0,1), seed = rank, unsorted as drawn;
tol = 1e-8, convergence_criteria = :both, quadspec = (2^12, 2^4)  (BASELINE config 2).

value : whole-job K(r) evals/s with the distances already resident in HBM (device pointers in,
        device pointers out), timed with CUDA events on the library's stream, max over ranks.
e2e   : the same metric through the public kernel_values call with HOST (pinned) buffers: the H2D copy
        of the distances and the D2H copy of values and errors are inside the timed region.
N > 1 : every rank evaluates its own batch of n distances (weak scaling); the ranks run ONE adaptive
        loop in lock step through scalar collectives (spectralkernels.jl_b200/sharded.py): single-warp exchange
        kernels over NVLink peer-mapped mailboxes by default, SK_COMM_TRANSPORT=nccl for in-library NCCL
        all-reduces, SK_COMM=torch for torch.distributed (A/B references).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic work per unit = one (active target, sub-interval) pair
FLOPS_PER_UNIT_SURVEY = 764     # F_fused: per-target tap polynomials shared by the m- and 2m-rule grids
# FP64 flops k_interp_cells actually executes per unit at the benchmark's density (ncu op counters:
# DFMA, DADD, DMUL per target; refreshed from profiles/r2_traffic.json when that file is present): cell polynomials
# replace the per-target tap evaluation, so far fewer flops are needed for the same result
FLOPS_PER_UNIT_EXECUTED = 2 * 87.6 + 21.4 + 16.1
# interpolation kernel with the fused (speculative) commit: read r (8) + read (ks,errs) (16) + write (ks,errs)
# (16) + write the roll-back copy (16)
BYTES_PER_UNIT_K4 = 56
# K8 (unique / sort / inverse map, csrc/sk_k8.cuh) per input distance: stats 8 + sample 1 + scatter 8 + 16 + finish
# 16 + 8 + 4; gather to the input order: inv 4 + (ks, errs) 16 + distance 8 + values and errors 16
BYTES_PER_INPUT_K8 = 61
BYTES_PER_INPUT_GATHER = 44
WORKLOAD = ("matern nu=1.5 rho=1 K(0)=1, r~U(0,1) seed=rank unsorted, tol=1e-8, quadspec (4096,16), :both "
            "(BASELINE config 2)")


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu --set full capture
    of this same command (scripts/ncu_summary.py writes profiles/r2_traffic.json); None when no capture is committed."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:
        return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000, help="distances per GPU")
    ap.add_argument("--cpu-sample", type=int, default=10_000_000, help="distances in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--interp-mode", type=int, default=0, help="0 cell polynomials (default), 1 per-target taps (A/B reference)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: n distances per GPU (the headline line); strong: the n distances of the BASELINE workload "
                         "split over the GPUs.  The weak line always carries the strong-scaling leg as well (key 'strong')")
    ap.add_argument("--no-strong-leg", action="store_true")
    return ap.parse_args()


def workload_sdf_params():
    return (1.0 / (np.pi / 2), 1.0, 1.5)            # K(0) = 1 (pattern of scripts/figures/speed_test_plot.jl:27-28)


def make_distances(n: int, rank: int) -> np.ndarray:
    return np.random.default_rng(rank).uniform(0.0, 1.0, n)


class ClockSampler:
    """SM clock and clock-event (throttle) reasons DURING the timed region.  NVML is polled from a thread every 2 ms
    (the timed region of the default run is 10-30 ms: `nvidia-smi -lms` would deliver at most one sample); nvidia-smi
    is only the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, cuda_index: int):
        self.idx, self.rows, self.proc, self.h, self.nv = cuda_index, [], None, None, None
        self.sm, self.mask, self.mx, self.run = [], 0, None, False
        if cuda_index < 0:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                self.h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid))
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                phys = int(vis.split(",")[cuda_index]) if vis and vis.split(",")[cuda_index].isdigit() else cuda_index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.h = self.nv = None

    def _poll(self):
        nv, h = self.nv, self.h
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.idx < 0:
            return
        if self.nv is not None:
            self.run = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.run = False
            self.th.join(timeout=1)
            nv, names = self.nv, []
            for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                              ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                              ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                              ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown)):
                if self.mask & int(bit):
                    names.append(name)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(names), "samples": len(self.sm), "how": "NVML polled every 2 ms from the warm-up through the timed resident and end-to-end steps"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"] if self.idx >= 0 else [],
                    "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "how": "nvidia-smi -lms 100"}


def oracle_cpu_run(n_sample: int, steps: int, warmup: int):
    """The reference's CPU path restated (oracle port): numpy driver + from-scratch CPU type-3 NUFFT
    with all host threads, on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sk_oracle as so
    so.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    phi, rho, nu = workload_sdf_params()
    S = lambda w: phi * (rho ** 2 + w ** 2) ** (-nu - 0.5)
    cfg = so.OracleConfig(S)
    xs = make_distances(10_000_000, 0)[:n_sample]
    for _ in range(warmup):
        so.kernel_values(cfg, xs[: max(1000, n_sample // 50)], k0=1.0, transform="nufft")
    t0 = time.perf_counter()
    for _ in range(steps):
        vals, _ = so.kernel_values(cfg, xs, k0=1.0, transform="nufft")
    dt = time.perf_counter() - t0
    true = (1 + 2 * np.pi * xs) * np.exp(-2 * np.pi * xs)
    return {"evals_per_s": n_sample * steps / dt, "seconds": dt, "threads": so.num_threads(),
            "max_err": float(np.max(np.abs(vals - true)))}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return None
    steps, warmup = max(1, args.steps), max(3, args.warmup)       # the same K and W as the repo's arm
    r = oracle_cpu_run(args.cpu_sample, steps, warmup)
    line = {
        "impl": "reference", "metric": "K(r) evals/sec at tol=1e-8 (Matern S, 1e7 r)", "value": r["evals_per_s"],
        "unit": "evals/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_per_gpu": args.cpu_sample, "n_per_step": args.cpu_sample, "full_n": 10_000_000,
                   "warmup_note": "warm-up steps run on a 1/50 sample (untimed; they only fault the pages in)",
                   "note": "reference (Julia+FINUFFT) cannot run in this image; this is the oracle port of its CPU "
                           "path (CPU restatement, not FINUFFT), all host threads, bounded sample per step"},
        "cpu_baseline": {"value": r["evals_per_s"], "unit": "evals/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{args.cpu_sample} of the 1e7 distances per step"},
        "e2e": {"value": r["evals_per_s"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "max_abs_err_vs_closed_form": r["max_err"],
    }
    return line


def main():
    args = parse()
    # stdout carries exactly ONE line (the JSON); libraries that print banners to stdout (NCCL's version
    # line) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def run(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import spectralkernels_jl_b200 as sk

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    # one process per GPU: stay on the CPUs (NUMA node) of this rank's GPU before any pinned buffer is allocated
    numa_bound = sk.bind_to_gpu_cpus(local_rank) if (world > 1 and not os.environ.get("SK_NO_NUMA_BIND")) else False
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from spectralkernels_jl_b200.sharded import LibComm, TorchComm
        comm = "torch" if os.environ.get("SK_COMM", "lib") == "torch" else "lib"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(3, args.warmup)
    K = max(1, args.steps)
    phi, rho, nu = workload_sdf_params()
    cfg = sk.AdaptiveKernelConfig(sk.Matern(phi, rho, nu), device=local_rank)
    eng = cfg.engine
    if comm == "lib":       # scalar all-reduces inside the library, on its stream (NCCL)
        comm = LibComm.from_torch(eng)
    elif comm == "torch":   # the same reductions through torch.distributed (A/B)
        comm = TorchComm(device=torch.device("cuda", local_rank))
    if args.interp_mode:
        eng.set_interp_mode(args.interp_mode)
    k0 = 1.0
    import gc

    def measure(xs_local: np.ndarray, sample_clocks: bool):
        """W warm-up + K timed resident steps, then K timed end-to-end steps, on this rank's distances."""
        n = xs_local.size
        host_in = sk.PinnedArray(n)
        host_in.array[:] = xs_local
        host_v, host_e = sk.PinnedArray(n), sk.PinnedArray(n)
        d_in = torch.from_numpy(host_in.array).to(f"cuda:{local_rank}")
        d_v = torch.empty(n, dtype=torch.float64, device=d_in.device)
        d_e = torch.empty(n, dtype=torch.float64, device=d_in.device)
        torch.cuda.synchronize()

        def step_resident(trace=None):
            sk.kernel_values(cfg, None, k0=k0, xs_device=(d_in.data_ptr(), n), out_device=(d_v.data_ptr(), d_e.data_ptr()),
                             comm=comm, trace=trace)

        def step_e2e():
            sk.kernel_values(cfg, host_in.array, k0=k0, comm=comm, out_vals=host_v.array, out_errs=host_e.array)

        trace = []
        # clocks and throttle reasons are sampled from the warm-up through the last timed step (resident, per-stage and
        # end-to-end loops: the GPU is under load throughout), so that the short timed region is well covered
        sampler = ClockSampler(local_rank if (rank == 0 and sample_clocks) else -1)
        sampler.start()
        step_resident(trace)
        for _ in range(W - 1):
            step_resident()
        eng.set_timing(True)
        gc.collect()
        gc.disable()                  # no collector pauses inside the timed regions (ranks wait for each other)
        agg = {"units": 0, "interp_ms": 0.0, "source_ms": 0.0, "sort_ms": 0.0, "gather_ms": 0.0, "launches": 0,
               "subintervals": 0, "chained": 0}
        barrier()
        launches0 = eng.stats().get("launches_total", 0)
        t0 = time.perf_counter()
        eng.timer_begin()
        for _ in range(K):
            step_resident()
            st = eng.stats()
            for key_, src in (("units", "units"), ("interp_ms", "interp_ms"), ("source_ms", "source_ms"), ("sort_ms", "sort_ms"),
                              ("gather_ms", "gather_ms"), ("launches", "kernel_launches"), ("subintervals", "n_subintervals")):
                agg[key_] += st[src]
        dev_ms = eng.timer_end()
        agg["launches_total"] = eng.stats().get("launches_total", 0) - launches0      # every kernel of the K steps, sort included
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        eng.set_timing(False)
        # the per-stage timers add event synchronisations to the step: time the same K steps once more without them
        barrier()
        eng.timer_begin()
        for _ in range(K):
            step_resident()
        res_ms = max(eng.timer_end(), 0.0)
        agg["chained"] = eng.stats().get("n_chained", 0)        # launches picked up ahead of time in the last (untimed-stage) step
        barrier()
        # end to end (host buffers)
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        # the same end-to-end steps PIPELINED through the public asynchronous form of the call (async_results=True: the
        # gather runs on the compute stream, the D2H copy of step i on the copy stream while step i + 1 uploads, sorts and
        # integrates; two output slots; results_wait() at the end): what a caller streaming batches of distances gets
        # from the duplex PCIe link.  Reported next to `e2e`, never instead of it.
        pipe_ms = None
        try:
            if world == 1:
                hv2, he2 = sk.PinnedArray(n), sk.PinnedArray(n)
                slots = [(host_v, host_e), (hv2, he2)]
                for i in range(2):
                    sk.kernel_values(cfg, host_in.array, k0=k0, out_vals=slots[i][0].array, out_errs=slots[i][1].array,
                                     async_results=True)
                eng.results_wait()
                t0 = time.perf_counter()
                for i in range(K):
                    sk.kernel_values(cfg, host_in.array, k0=k0, out_vals=slots[i & 1][0].array, out_errs=slots[i & 1][1].array,
                                     async_results=True)
                eng.results_wait()
                pipe_ms = 1e3 * (time.perf_counter() - t0)
                if not (np.array_equal(slots[(K - 1) & 1][0].array, slots[K & 1][0].array)):
                    pipe_ms = None            # (both slots hold the same run's values: anything else disqualifies the number)
                hv2.free(); he2.free()
        except Exception:
            pipe_ms = None
        clocks = sampler.stop()
        gc.enable()
        if world > 1:
            t = torch.tensor([res_ms, e2e_ms, wall_ms, dev_ms], dtype=torch.float64, device=d_in.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res_ms, e2e_ms, wall_ms, dev_ms = t.tolist()
        true = (1 + 2 * np.pi * host_in.array) * np.exp(-2 * np.pi * host_in.array)
        max_err = float(np.max(np.abs(host_v.array - true)))
        same = bool(torch.equal(d_v.cpu(), torch.from_numpy(host_v.array)))
        out = {"n": n, "res_ms": res_ms, "timed_stage_ms": dev_ms, "e2e_ms": e2e_ms, "pipe_ms": pipe_ms, "wall_ms": wall_ms, "agg": agg,
               "clocks": clocks, "trace": trace, "max_err": max_err, "same": same}
        del d_in, d_v, d_e
        host_in.free(); host_v.free(); host_e.free()
        return out

    n = args.n
    strong_only = args.scaling == "strong"
    weak = None if strong_only else measure(make_distances(n, rank), True)
    strong = None
    if (world > 1 or strong_only) and not args.no_strong_leg:
        # the BASELINE workload itself (n distances, seed 0) split over the GPUs: contiguous chunks of the drawn order
        full = make_distances(n, 0)
        lo, hi = rank * n // world, (rank + 1) * n // world
        strong = measure(full[lo:hi].copy(), strong_only)
        del full
    main_ = strong if strong_only else weak

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fp64_tf, fp64_ms = eng.fp64_peak()
        agg, res_ms, e2e_ms = main_["agg"], main_["res_ms"], main_["e2e_ms"]
        nloc = main_["n"]
        total = nloc * world if not strong_only else n
        n_launch = max(1, agg["subintervals"])
        units_per_launch = agg["units"] / n_launch
        k4_ms = agg["interp_ms"] / n_launch
        tr = measured_traffic() or {}
        exe_fpu = float(tr.get("executed_flops_per_unit", FLOPS_PER_UNIT_EXECUTED))
        alg_tf = units_per_launch * FLOPS_PER_UNIT_SURVEY / (k4_ms * 1e-3) / 1e12 if k4_ms > 0 else None
        exe_tf = units_per_launch * exe_fpu / (k4_ms * 1e-3) / 1e12 if k4_ms > 0 else None
        ach_gbs = units_per_launch * BYTES_PER_UNIT_K4 / (k4_ms * 1e-3) / 1e9 if k4_ms > 0 else None
        ms_step = res_ms / K
        # step level: every (target, sub-interval) unit of the step at SURVEY's algorithmic 764 flops against the FP64 peak,
        # over the whole resident step (sort, source side, interpolation, gather, host round trips)
        step_roof_ms = (agg["units"] / K) * FLOPS_PER_UNIT_SURVEY / (fp64_tf * 1e12) * 1e3
        sort_ms, gather_ms = agg["sort_ms"] / K, agg["gather_ms"] / K
        k8 = {"kernel": "K8 unique/sort/inverse map (sk_k8.cuh) + gather", "bound": "hbm", "unit": "GB/s", "peak": hbm_peak,
              "sort_ms": sort_ms, "gather_ms": gather_ms,
              "achieved": nloc * (BYTES_PER_INPUT_K8 + BYTES_PER_INPUT_GATHER) / ((sort_ms + gather_ms) * 1e-3) / 1e9
              if (sort_ms + gather_ms) > 0 else None,
              "bytes_per_input": BYTES_PER_INPUT_K8 + BYTES_PER_INPUT_GATHER}
        k8["frac"] = k8["achieved"] / hbm_peak if k8["achieved"] else None
        line = {
            "metric": "K(r) evals/sec at tol=1e-8 (Matern S, 1e7 r)",
            "value": total * K / (res_ms * 1e-3), "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong_only else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "n_per_gpu": nloc, "k0": "passed (=1.0) in both arms", "nufft_eps": 1e-15,
                       "l2": "inputs + work arrays (>1 GB per step) exceed the 126 MB L2; no explicit flush",
                       "launch_chaining": ("device-guarded: first panel queued behind the sort, second behind the first, gather "
                                           "behind the second (sk_first_panel_early / sk_subinterval_chain / "
                                           "sk_results_chain_device); host scalar work overlapped") if world == 1 else
                                          ("host scalar work overlapped with device work (begin/end halves); chained launches "
                                           + ("(first panel behind the sort, second behind the first: guards on the global scalars, "
                                              "void exchanges)" if (world <= 2 or os.environ.get("SK_SHARDED_CHAIN") == "1")
                                              and os.environ.get("SK_SHARDED_CHAIN") != "0" and getattr(comm, "mode", "") == "peer"
                                              else "off at this rank count")),
                       "parallelism": ("target-sharded, scalar collectives only: "
                                       + {"peer": "single-warp exchange kernels over NVLink peer-mapped mailboxes "
                                                  "(k_peer_exchange) on the compute stream, no NCCL on the data path",
                                          "nccl": "in-library NCCL all-reduces on the compute stream"}.get(
                                              getattr(comm, "mode", ""), "torch.distributed all-reduces")
                                       ) if world > 1 else "single GPU",
                       "cpu_affinity": "each rank pinned to its GPU's local CPUs (NVML)" if numa_bound else "default",
                       "panels": [(t["a"], t["b"], t["hi_before"], t["hi_after"]) for t in main_["trace"] if t["kind"] == "panel"]},
            "e2e": {"value": total * K / (e2e_ms * 1e-3), "unit": "evals/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": 8 * nloc, "d2h_bytes_per_step": 16 * nloc,
                    "note": "kernel_values with pinned host buffers; values and errors both copied back"},
            "e2e_pipelined": ({"value": total * K / (main_["pipe_ms"] * 1e-3), "unit": "evals/s", "ms_per_step": main_["pipe_ms"] / K,
                               "note": "the same host-to-host steps through kernel_values(async_results=True): the D2H copy "
                                       "of step i overlaps step i+1 (duplex PCIe); reported beside e2e, not instead of it"}
                              if main_.get("pipe_ms") else None),
            "gpu_launches": int(agg.get("launches_total") or agg["launches"]),
            "clocks": main_["clocks"],
            "roofline": {"bound": "fp64", "kernel": "k_interp_cells<16>", "achieved": exe_tf, "peak": fp64_tf,
                         "unit": "TFLOP/s", "frac": (exe_tf / fp64_tf) if exe_tf else None,
                         "traffic": tr.get("interp_cells_dram_bytes_per_launch"),
                         "traffic_source": tr.get("source", "no ncu capture committed for this build (profiles/r2_traffic.json)"),
                         "flops_per_unit": exe_fpu,
                         "note": "achieved/frac count the FP64 flops the kernel EXECUTES per unit (ncu op counters): it "
                                 "evaluates cell polynomials, not SURVEY 8(d)'s per-target taps (764 flops/unit), so the "
                                 "algorithmic figure is reported separately (algorithmic_*) and at step level (step)",
                         "algorithmic_flops_per_unit": FLOPS_PER_UNIT_SURVEY, "algorithmic_achieved": alg_tf,
                         "algorithmic_frac": (alg_tf / fp64_tf) if alg_tf else None,
                         "units_per_launch": units_per_launch, "avg_launch_ms": k4_ms,
                         "kernel_share_of_step": (agg["interp_ms"] / K) / (main_["timed_stage_ms"] / K),
                         "peak_source": f"measured live: sk_fp64_peak DFMA micro-benchmark ({fp64_ms:.3f} ms, SM clock "
                                        f"{main_['clocks'].get('sm_mhz')} MHz under load; MEASURED_PEAKS.json has no FP64 figure)",
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": (ach_gbs / hbm_peak) if ach_gbs else None, "bytes_per_unit": BYTES_PER_UNIT_K4,
                                 "peak_source": hbm_src},
                         "step": {"units": agg["units"] / K, "flops_per_unit": FLOPS_PER_UNIT_SURVEY, "roofline_ms": step_roof_ms,
                                  "ms_per_step": ms_step, "frac": step_roof_ms / ms_step,
                                  "note": "whole resident step against the FP64 roofline at SURVEY 8(d)'s algorithmic count"},
                         "sort": k8},
            "source_side_ms_per_step": agg["source_ms"] / K, "interp_ms_per_step": agg["interp_ms"] / K,
            "sort_ms_per_step": sort_ms, "gather_ms_per_step": gather_ms,
            "units_per_step": agg["units"] / K, "wall_ms_per_step": main_["wall_ms"] / K,
            "chained_launches_per_step": int(agg["chained"]),
            "parity": {"max_abs_err_vs_closed_form": main_["max_err"], "resident_equals_e2e_bitwise": main_["same"]},
        }
        if strong is not None and not strong_only:
            line["strong"] = {"workload": f"the {n} distances of the BASELINE workload (seed 0) split over {world} GPUs",
                              "value": n * K / (strong["res_ms"] * 1e-3), "ms_per_step": strong["res_ms"] / K,
                              "e2e_value": n * K / (strong["e2e_ms"] * 1e-3), "e2e_ms_per_step": strong["e2e_ms"] / K,
                              "unit": "evals/s", "n_per_gpu": strong["n"],
                              "max_abs_err_vs_closed_form": strong["max_err"]}
        if world == 1 and not args.no_cpu_baseline:
            r = oracle_cpu_run(args.cpu_sample, 1, 1)
            line["cpu_baseline"] = {"value": r["evals_per_s"], "unit": "evals/s", "cores": r["threads"], "kind": "port",
                                    "sample": f"first {args.cpu_sample} of the 1e7 distances, 1 pass "
                                              f"({r['seconds']:.1f} s); oracle port (CPU restatement, not FINUFFT)"}
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


if __name__ == "__main__":
    main()
