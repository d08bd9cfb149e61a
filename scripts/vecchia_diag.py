"""Per-vector timing of the Vecchia kernel stage (diagnostic): python scripts/vecchia_diag.py [async]"""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import spectralkernels_jl_b200 as sk
import bench_vecchia as bv
use_async = len(sys.argv) > 1 and sys.argv[1] == "async"
rng = np.random.default_rng(0)
pts = rng.uniform(0, 1, (100_000, 2))
pairs = bv.knn_pairs(pts)
B = 24
hp = np.stack([1.0 + 0.05 * rng.standard_normal(B), 4.0 * np.exp(0.1 * rng.standard_normal(B)), 1.5 + 0.05 * rng.standard_normal(B)], axis=1)
eng = sk.Session(0)
outs = [sk.PinnedArray(pairs.shape[0]), sk.PinnedArray(pairs.shape[0])]
for rep in range(2):
    times = []
    for it in range(B):
        h = hp[it]
        cfg = sk.AdaptiveKernelConfig(sk.Matern(h[0], h[1], h[2], d=2), dim=2, engine=eng)
        t0 = time.perf_counter()
        k0 = sk.compute_k0(cfg)
        t1 = time.perf_counter()
        sk.kernel_values(cfg, None, k0=k0, points=pts, pairs=pairs, reuse_targets=(it > 0 or rep > 0), want_errors=False,
                         out_vals=outs[it & 1].array, async_results=use_async)
        t2 = time.perf_counter()
        times.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), eng.stats()["last_nf2"], eng.stats()["kernel_launches"]))
    eng.results_wait()
    print("rep", rep, "async" if use_async else "sync", " ".join(f"{a:.1f}/{b:.1f}" for a, b, _, _ in times), flush=True)
