"""
Writes the distance sets `tests/golden/make_traces.jl` feeds to the unmodified reference, as raw little-endian float64
files under tests/golden/trace_inputs/ (git-ignored: they are regenerated from seeds), so that Julia and numpy see
bit-identical inputs:

    config2_2e3.f64    2000 of the 1e7 config-2 distances (+ the maximum), seed 0
    config2_1e7.f64    all 1e7 (80 MB; only with --full)
    config3_lags.f64   4000 strided pairwise distances of the 1e4 config-3 points (+ the maximum)
    config4_1e6.f64    the 1e6 config-4 distances (8 MB; only with --full; else 4000 of them + the maximum)
    config5_lags.f64   4000 strided lags of the config-5 KNN pair list (+ the maximum; needs scipy)

    python tests/golden/make_trace_inputs.py [--full]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main(full: bool):
    out = os.path.join(HERE, "trace_inputs")
    os.makedirs(out, exist_ok=True)
    xs = np.random.default_rng(0).uniform(0.0, 1.0, 10_000_000)
    sub = xs[:2000].copy()
    sub[-1] = xs.max()
    sub.astype("<f8").tofile(os.path.join(out, "config2_2e3.f64"))
    if full:
        xs.astype("<f8").tofile(os.path.join(out, "config2_1e7.f64"))
    pts = np.random.default_rng(0).uniform(0, 1, (10_000, 2))
    n = pts.shape[0]
    npairs = n * (n - 1) // 2
    t = np.linspace(0, npairs - 1, 4000).astype(np.int64)
    i = np.floor(((2 * n - 1) - np.sqrt((2.0 * n - 1) ** 2 - 8.0 * t)) / 2).astype(np.int64)
    i = np.where(i * (2 * n - i - 1) // 2 > t, i - 1, i)
    i = np.where((i + 1) * (2 * n - i - 2) // 2 <= t, i + 1, i)
    j = t - i * (2 * n - i - 1) // 2 + i + 1
    lag = np.sqrt((pts[i, 0] - pts[j, 0]) ** 2 + (pts[i, 1] - pts[j, 1]) ** 2)
    far = 0.0
    for s in range(0, n, 500):
        d = pts[s:s + 500, None, :] - pts[None, :, :]
        far = max(far, float(np.sqrt((d ** 2).sum(-1)).max()))
    np.append(lag, far).astype("<f8").tofile(os.path.join(out, "config3_lags.f64"))
    x4 = np.random.default_rng(0).uniform(0, 1, 1_000_000)
    if not full:
        x4 = np.append(x4[:: 250], x4.max())
    x4.astype("<f8").tofile(os.path.join(out, "config4_1e6.f64"))
    try:
        sys.path.insert(0, ROOT)
        from bench_vecchia import knn_pairs
        p5 = np.random.default_rng(0).uniform(0, 1, (100_000, 2))
        pairs = knn_pairs(p5)
        d = p5[pairs[:, 0]] - p5[pairs[:, 1]]
        lag5 = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)
        sel = np.linspace(0, lag5.size - 1, 4000).astype(np.int64)
        np.append(lag5[sel], lag5.max()).astype("<f8").tofile(os.path.join(out, "config5_lags.f64"))
    except Exception as e:                      # the package import needs the built library
        print("config 5 skipped:", e)
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main("--full" in sys.argv)
