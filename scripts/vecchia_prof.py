"""Where does a Vecchia kernel-stage call spend its host time?  cProfile of the set-up calls and of a few vectors
(the hyperparameter vectors of bench_vecchia.py)."""
import cProfile, pstats, sys, time, io
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import spectralkernels_jl_b200 as sk
import bench_vecchia as bv
rng = np.random.default_rng(0)
pts = rng.uniform(0, 1, (100_000, 2))
pairs = bv.knn_pairs(pts)
B = 24
hp = np.stack([1.0 + 0.05 * rng.standard_normal(B), 4.0 * np.exp(0.1 * rng.standard_normal(B)), 1.5 + 0.05 * rng.standard_normal(B)], axis=1)
eng = sk.Session(0)
outs = [sk.PinnedArray(pairs.shape[0]), sk.PinnedArray(pairs.shape[0])]
def run(h, reuse, slot):
    cfg = sk.AdaptiveKernelConfig(sk.Matern(h[0], h[1], h[2], d=2), dim=2, engine=eng)
    k0 = sk.compute_k0(cfg)
    sk.kernel_values(cfg, None, k0=k0, points=pts, pairs=pairs, reuse_targets=reuse, want_errors=False, out_vals=outs[slot].array,
                     async_results=True)
calls = [("first", hp[0], False), ("warm set-up", hp[0], False)] + [(f"vector {i}", hp[i], True) for i in range(8)]
for i, (name, h, reuse) in enumerate(calls):
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable(); run(h, reuse, i & 1); pr.disable()
    dt = time.perf_counter() - t0
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(4)
    st = eng.stats()
    print(f"== {name} {np.round(h, 4).tolist()}: {1e3 * dt:.1f} ms; launches {st['kernel_launches']} nf2 {st['last_nf2']} units {st['units']} "
          f"subs {st['n_subintervals']} hankel {st['n_hankel']} direct {st['n_direct']} prefetch {st['n_prefetch_issued']}/{st['n_prefetch_hits']}")
    print("\n".join(l[:160] for l in s.getvalue().splitlines() if "/" in l or "{" in l), flush=True)
eng.results_wait()
