import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import spectralkernels_jl_b200 as sk
rng = np.random.default_rng(0)
for rmax in (1.0, 1e3):
    x = rng.uniform(0, rmax, 10_000_000)
    pin, hv, he = sk.PinnedArray(x.size), sk.PinnedArray(x.size), sk.PinnedArray(x.size)
    pin.array[:] = x
    cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0 / (np.pi / 2), 1.0, 1.5))
    eng = cfg.engine
    for rep in range(3):
        t0 = time.perf_counter(); info = eng.targets_set(pin.array); t1 = time.perf_counter()
        sk.kernel_values(cfg, pin.array, k0=1.0, out_vals=hv.array, out_errs=he.array, reuse_targets=True); t2 = time.perf_counter()
        sk.kernel_values(cfg, pin.array, k0=1.0, out_vals=hv.array, out_errs=he.array, reuse_targets=True, want_errors=False); t3 = time.perf_counter()
        print(f"rmax {rmax}: targets_set {1e3*(t1-t0):.2f} ms (sort path {eng.stats()['sort_two_level']}), loop+results {1e3*(t2-t1):.2f} ms, loop+values only {1e3*(t3-t2):.2f}", flush=True)
