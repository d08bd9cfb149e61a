"""Timing breakdown of one hyperparameter evaluation of bench_vecchia.py (diagnostic)."""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import spectralkernels_jl_b200 as sk
import bench_vecchia as bv
rng = np.random.default_rng(0)
pts = rng.uniform(0, 1, (100_000, 2))
t0 = time.perf_counter(); pairs = bv.knn_pairs(pts); print("pairs s", time.perf_counter() - t0, pairs.shape, flush=True)
eng = sk.Session(0)
eng.set_timing(True)
out = sk.PinnedArray(pairs.shape[0])
for it in range(4):
    cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0, 4.0 + 0.1 * it, 1.5, d=2), dim=2, engine=eng)
    t0 = time.perf_counter(); k0 = sk.compute_k0(cfg); t1 = time.perf_counter()
    tr = []
    sk.kernel_values(cfg, None, k0=k0, points=pts, pairs=pairs, reuse_targets=it > 0, want_errors=False, out_vals=out.array, trace=tr)
    t2 = time.perf_counter()
    st = eng.stats()
    print(f"it {it}: k0 {1e3*(t1-t0):.2f} ms, kernel_values {1e3*(t2-t1):.2f} ms, interp {st['interp_ms']:.2f} source {st['source_ms']:.2f} "
          f"subs {st['n_subintervals']} hankel {st['n_hankel']} direct {st['n_direct']} units {st['units']} launches {st['kernel_launches']}", flush=True)
    print("   ", [(round(t['a']), round(t['b']), t.get('hi_before'), t.get('hi_after')) for t in tr if t['kind'] == 'panel'])
lags = np.linalg.norm(pts[pairs[:, 0]] - pts[pairs[:, 1]], axis=1)
print("lag quantiles", np.quantile(lags, [0, 0.01, 0.5, 0.99, 1.0]), "unique", np.unique(lags).size)
