"""Small dim = 2 run through the O(N) Hankel path (used under compute-sanitizer)."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import spectralkernels_jl_b200 as sk
rng = np.random.default_rng(0)
xs = np.concatenate([rng.uniform(0, 1, 3000), 10 ** rng.uniform(-6, 0, 500)])
cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0, 1.0, 1.5, d=2), dim=2)
cfg.engine.set_hankel_mode(2)
v, e = sk.kernel_values(cfg, xs, k0=2.0943951023931953)
print("ok", cfg.engine.stats()["n_hankel"], float(v.max()))
