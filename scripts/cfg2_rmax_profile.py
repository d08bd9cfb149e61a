"""config 2 with r ~ U(0, 1e3): two kernel_values calls (for ncu launch lists)."""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import spectralkernels_jl_b200 as sk
rng = np.random.default_rng(0)
x = rng.uniform(0, 1e3, 10_000_000)
pin, hv, he = sk.PinnedArray(x.size), sk.PinnedArray(x.size), sk.PinnedArray(x.size)
pin.array[:] = x
cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0 / (np.pi / 2), 1.0, 1.5))
cfg.engine.set_timing(True)
for _ in range(3):
    t0 = time.perf_counter()
    tr = []
    sk.kernel_values(cfg, pin.array, k0=1.0, out_vals=hv.array, out_errs=he.array, trace=tr)
    print("ms", 1e3 * (time.perf_counter() - t0), cfg.engine.stats(), flush=True)
print([(t.get("a"), t.get("b"), t.get("rel_err"), t.get("accepted")) for t in tr if t["kind"] == "subinterval"])
