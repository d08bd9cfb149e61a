#!/usr/bin/env python
"""
Secondary measurements on the other BASELINE.json configs (one JSON line each; bench.py remains the
headline contract on config 2).  Single GPU.

  config 1  README demo: S = (1+w^2)^-2, 1000 log-spaced r in [1e-6, 1], tol 1e-8   (latency bound)
  config 3  singular Matern alpha = 0.5 as a 1-D kernel on the 49 995 000 pairwise distances of 1e4 random
            2-D points (lags computed on the device from the points, sk_targets_set_pairs)
  config 4  1e6 distances: K, K' (range via warping), dK/dphi, dK/drho, dK/dnu -- 5 adaptive runs over the
            same lags (uploaded and sorted once)
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import spectralkernels_jl_b200 as sk  # noqa: E402


def timed(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def config1():
    rs = 10 ** np.linspace(-6, 0, 1000)
    cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0, 1.0, 1.5))
    k0 = np.pi / 2
    dt, (v, e) = timed(lambda: sk.kernel_values(cfg, rs, k0=k0), warm=3, reps=20)
    err = float(np.max(np.abs(v - (np.pi / 2) * (1 + 2 * np.pi * rs) * np.exp(-2 * np.pi * rs))) / k0)
    return {"config": 1, "workload": "README demo, 1000 log-spaced r", "ms": 1e3 * dt, "evals_per_s": rs.size / dt,
            "max_err_over_k0": err, "stats": cfg.engine.stats()}


def config3():
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 1, (10_000, 2))
    parms = (1.0, 1.0, 1.5)
    cfg = sk.AdaptiveKernelConfig(sk.Matern(*parms), alpha=0.5)
    k0 = sk.compute_k0(cfg)
    n = pts.shape[0] * (pts.shape[0] - 1) // 2
    host_v, host_e = sk.PinnedArray(n), sk.PinnedArray(n)
    tr = []
    dt, _ = timed(lambda: sk.kernel_values(cfg, None, k0=k0, points=pts, out_vals=host_v.array, out_errs=host_e.array,
                                           trace=tr), warm=1, reps=2)
    st = cfg.engine.stats()
    return {"config": 3, "workload": "singular Matern alpha=0.5 (1-D kernel), 49 995 000 pairwise distances of 1e4 2-D points, "
                                     "lags computed on device, values+errors copied back",
            "ms": 1e3 * dt, "evals_per_s": n / dt, "n": n, "units": st["units"], "subintervals": st["n_subintervals"],
            "panels": [(t["a"], t["b"], t["hi_before"], t["hi_after"]) for t in tr if t["kind"] == "panel"][-8:],
            "finite": bool(np.all(np.isfinite(host_v.array)))}


def config4():
    rng = np.random.default_rng(0)
    n = 1_000_000
    pin = sk.PinnedArray(n)
    xs = pin.array
    xs[:] = rng.uniform(0, 1, n)
    S = sk.Matern(1.0 / (np.pi / 2), 1.0, 1.5)
    cfg = sk.AdaptiveKernelConfig(S)
    k0 = 1.0
    bufs = [sk.PinnedArray(n) for _ in range(6)]          # K, errs, K', dK/dphi, dK/drho, dK/dnu (pinned: full PCIe rate)

    def run():
        v, _ = sk.kernel_values(cfg, xs, k0=k0, out_vals=bufs[0].array, out_errs=bufs[1].array)
        dk = sk.kernel_derivative(cfg, xs, k0, reuse_targets=True, out_vals=bufs[2].array)
        d = sk.kernel_sdf_derivatives(cfg, xs, k0, reuse_targets=True, outs=[b.array for b in bufs[3:6]])
        return v, dk, d

    dt, (v, dk, d) = timed(run, warm=2, reps=5)
    true = (1 + 2 * np.pi * xs) * np.exp(-2 * np.pi * xs)
    dtrue = -(2 * np.pi) ** 2 * xs * np.exp(-2 * np.pi * xs)
    return {"config": 4, "workload": "1e6 distances: K, K', dK/dphi, dK/drho, dK/dnu (5 adaptive runs, lags sorted once; "
                                     "pinned host buffers, derivative runs return values only)",
            "ms": 1e3 * dt, "evals_per_s": 5 * xs.size / dt, "max_err_K": float(np.max(np.abs(v - true))),
            "max_err_dK": float(np.max(np.abs(dk - dtrue))),
            "max_err_dphi": float(np.max(np.abs(d[0] - true * (np.pi / 2))))}


def config3_dim2():
    """config 3 (ii): the same 49 995 000 pairwise distances with dim = 2 (J_0 kernel, p = +0.5, c = 2 pi,
    src/adaptive.jl:42-43): the O(N) nonuniform Hankel transform."""
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 1, (10_000, 2))
    parms = (1.0, 1.0, 1.5)
    cfg = sk.AdaptiveKernelConfig(sk.Matern(*parms, d=2), alpha=0.5, dim=2)
    k0 = sk.compute_k0(cfg)
    n = pts.shape[0] * (pts.shape[0] - 1) // 2
    host_v, host_e = sk.PinnedArray(n), sk.PinnedArray(n)
    tr = []
    cfg.engine.set_timing(True)
    dt, _ = timed(lambda: sk.kernel_values(cfg, None, k0=k0, points=pts, out_vals=host_v.array, out_errs=host_e.array,
                                           trace=tr), warm=1, reps=2)
    st = cfg.engine.stats()
    return {"config": "3-dim2", "workload": "singular Matern alpha=0.5, dim=2 (J_0 kernel), 49 995 000 pairwise distances of "
                                            "1e4 2-D points, lags computed on device, values+errors copied back",
            "ms": 1e3 * dt, "evals_per_s": n / dt, "n": n, "units": st["units"], "subintervals": st["n_subintervals"],
            "n_hankel": st["n_hankel"], "interp_ms": st["interp_ms"], "source_ms": st["source_ms"],
            "panels": [(t["a"], t["b"], t["hi_before"], t["hi_after"]) for t in tr if t["kind"] == "panel"][-8:],
            "finite": bool(np.all(np.isfinite(host_v.array)))}


def config2_dim2():
    """1e7 uniform lags, Matern nu = 1.5 in 2-D (J_0 kernel): closed form available."""
    from scipy import special
    rng = np.random.default_rng(0)
    pin = sk.PinnedArray(10_000_000)
    xs = pin.array
    xs[:] = rng.uniform(0, 1, xs.size)
    parms = (1.0, 1.0, 1.5)
    cfg = sk.AdaptiveKernelConfig(sk.Matern(*parms, d=2), dim=2)
    k0 = float(np.pi * parms[0] / (2 ** 0.5 * special.gamma(2.5)) * 2 ** 0.5 * special.gamma(1.5))
    host_v, host_e = sk.PinnedArray(xs.size), sk.PinnedArray(xs.size)
    cfg.engine.set_timing(True)
    dt, _ = timed(lambda: sk.kernel_values(cfg, xs, k0=k0, out_vals=host_v.array, out_errs=host_e.array), warm=1, reps=3)
    st = cfg.engine.stats()
    arg = 2 * np.pi * xs
    true = np.pi * parms[0] / (2 ** 0.5 * special.gamma(2.5)) * special.kv(1.5, arg) * arg ** 1.5
    return {"config": "2-dim2", "workload": "Matern nu=1.5 in 2-D (J_0 kernel), 1e7 uniform lags, end to end",
            "ms": 1e3 * dt, "evals_per_s": xs.size / dt, "units": st["units"], "subintervals": st["n_subintervals"],
            "n_hankel": st["n_hankel"], "interp_ms": st["interp_ms"], "source_ms": st["source_ms"],
            "max_err_over_k0": float(np.max(np.abs(host_v.array - true)) / k0)}


def config2_variants():
    """SURVEY 8(d) secondary sweeps on config 2 (1e7 distances, Matern nu = 1.5, K(0) = 1, end to end from pinned host
    memory): sorted-unique input (no sort: separates the K8 cost), r ~ U(0, 1e3) and log-uniform r in [1e-6, 1]
    (shrinking active sets: more, smaller panels)."""
    rng = np.random.default_rng(0)
    n = 10_000_000
    S = sk.Matern(1.0 / (np.pi / 2), 1.0, 1.5)
    out = []
    for name, gen in (("uniform (0,1), unsorted", lambda: rng.uniform(0, 1, n)),
                      ("uniform (0,1), sorted unique", lambda: np.unique(rng.uniform(0, 1, n))),
                      ("uniform (0,1e3)", lambda: rng.uniform(0, 1e3, n)),
                      ("log-uniform [1e-6,1]", lambda: 10 ** rng.uniform(-6, 0, n))):
        x = gen()
        pin, hv, he = sk.PinnedArray(x.size), sk.PinnedArray(x.size), sk.PinnedArray(x.size)
        pin.array[:] = x
        cfg = sk.AdaptiveKernelConfig(S)
        tr = []
        dt, _ = timed(lambda: sk.kernel_values(cfg, pin.array, k0=1.0, out_vals=hv.array, out_errs=he.array, trace=tr),
                      warm=2, reps=5)
        st = cfg.engine.stats()
        true = (1 + 2 * np.pi * x) * np.exp(-2 * np.pi * x)
        out.append({"input": name, "n": int(x.size), "ms": 1e3 * dt, "evals_per_s": x.size / dt, "units": st["units"],
                    "subintervals": st["n_subintervals"], "sort_path": st["sort_two_level"],
                    "panels": [(t["hi_before"], t["hi_after"]) for t in tr if t["kind"] == "panel"][-6:],
                    "max_err": float(np.max(np.abs(hv.array - true)))})
    return {"config": "2-variants", "workload": "config 2 secondary sweeps, end to end", "runs": out}


if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or [1, 4, 3]
    for c in which:
        print(json.dumps({1: config1, 3: config3, 4: config4, 32: config3_dim2, 22: config2_dim2, 20: config2_variants}[c]()), flush=True)
