"""
Device-group parity check on real GPUs (one process; not collected by pytest):

    python tests/multi_gpu_group_check.py

One kernel_values call over all visible GPUs (AdaptiveKernelConfig(..., devices=[0..N-1]) -> sk_group_*) against the
same call on GPU 0 alone: values and error estimates must be equal bit for bit, the sub-interval traces identical.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralkernels_jl_b200 as sk  # noqa: E402


def main():
    n_dev = torch.cuda.device_count()
    ok = True
    for name, S, gen, k0 in (
        ("matern_uniform", sk.Matern(1 / (np.pi / 2), 1.0, 1.5), lambda rng, n: rng.uniform(0, 1, n), 1.0),
        ("slow_decay_logspaced", sk.Matern(1.0, 0.5, 0.55), lambda rng, n: np.sort(10 ** rng.uniform(-4, 0, n))[::-1].copy(), 5.9),
    ):
        xs = gen(np.random.default_rng(7), 1_000_000)
        t1, tg = [], []
        v1, e1 = sk.kernel_values(sk.AdaptiveKernelConfig(S, device=0), xs, k0=k0, trace=t1)
        cfg = sk.AdaptiveKernelConfig(S, devices=list(range(n_dev)))
        vg, eg = sk.kernel_values(cfg, xs, k0=k0, trace=tg)
        key = lambda tr: [(t["a"], t["b"], t["accepted"]) for t in tr if t["kind"] == "subinterval"]
        same = np.array_equal(v1, vg) and np.array_equal(e1, eg, equal_nan=True) and key(t1) == key(tg)
        print(f"[multi_gpu_group_check] {name}: devices={n_dev} bitwise_and_trace_equal={same} "
              f"panels={sum(1 for t in tg if t['kind'] == 'panel')} units={cfg.engine.stats()['units']}", flush=True)
        ok = ok and same
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
