"""One dim = 2 kernel_values call (1e6 uniform lags) for ncu launch lists."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import spectralkernels_jl_b200 as sk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
xs = rng.uniform(0, 1, n)
cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0, 1.0, 1.5, d=2), dim=2)
for _ in range(2):
    v, e = sk.kernel_values(cfg, xs, k0=2.0943951023931953)
print("ok", cfg.engine.stats())
