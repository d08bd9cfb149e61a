#!/usr/bin/env python
"""
bench_group.py -- strong scaling of ONE kernel_values call from ONE process over N GPUs (the C ABI's device group,
sk_group_*): the reference's API is a single task calling kernel_values(cfg, xs) (src/adaptive.jl:95-108), so this is
what a Julia caller sees when it hands the library more than one device.

    python bench_group.py [--n 10000000] [--steps 5] [--warmup 3] [--max-gpus 8]

Workload: BASELINE config 2 (bench.py): Matern nu = 1.5, n distances ~ U(0,1) seed 0, tol 1e-8; host (pinned) arrays
in, values + errors out -- the end-to-end metric of bench.py.  One JSON line with a row per device count; the N = 1
row uses the plain single-context session.  Results of every N are compared bit for bit with N = 1.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import spectralkernels_jl_b200 as sk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--max-gpus", type=int, default=8)
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    n = args.n
    xs = sk.PinnedArray(n)
    xs.array[:] = np.random.default_rng(0).uniform(0.0, 1.0, n)
    vals, errs = sk.PinnedArray(n), sk.PinnedArray(n)
    S = sk.Matern(1.0 / (np.pi / 2), 1.0, 1.5)
    rows, ref = [], None
    for nd in [d for d in (1, 2, 4, 8) if d <= min(have, args.max_gpus)]:
        cfg = sk.AdaptiveKernelConfig(S, devices=list(range(nd)))
        for _ in range(max(3, args.warmup)):
            sk.kernel_values(cfg, xs.array, k0=1.0, out_vals=vals.array, out_errs=errs.array)
        cfg.engine.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sk.kernel_values(cfg, xs.array, k0=1.0, out_vals=vals.array, out_errs=errs.array)
        dt = (time.perf_counter() - t0) / args.steps
        if ref is None:
            ref = (vals.array.copy(), errs.array.copy())
        same = bool(np.array_equal(ref[0], vals.array) and np.array_equal(ref[1], errs.array))
        true = (1 + 2 * np.pi * xs.array) * np.exp(-2 * np.pi * xs.array)
        rows.append({"n_gpus": nd, "e2e_ms_per_call": 1e3 * dt, "evals_per_s": n / dt, "bitwise_equal_to_1gpu": same,
                     "max_abs_err_vs_closed_form": float(np.max(np.abs(vals.array - true)))})
        cfg.engine.close()
    print(json.dumps({"metric": "K(r) evals/s, one process, one kernel_values call over N GPUs (device group), end to end",
                      "n": n, "steps": args.steps, "scaling": "strong", "rows": rows,
                      "h2d_bytes_per_call": 8 * n, "d2h_bytes_per_call": 16 * n}), flush=True)


if __name__ == "__main__":
    main()
