"""K8 timing diagnostics: targets_set on several distance sets, gather, per-kernel sanity vs numpy."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectralkernels_jl_b200 as sk
import torch

eng = sk.Session(0)
rng = np.random.default_rng(0)
cases = {"U(0,1) 1e7": rng.uniform(0, 1, 10_000_000), "logU 1e7": 10 ** rng.uniform(-6, 0, 10_000_000),
         "U(0,1) 1e6": rng.uniform(0, 1, 1_000_000), "U(0,1) 1e3": rng.uniform(0, 1, 1000),
         "dups 1e7": rng.uniform(0, 1, 2_500_000).repeat(4)}
for name, xs in cases.items():
    d = torch.from_numpy(xs).cuda()
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        info = eng.targets_set_device(d.data_ptr(), xs.size)
        t1 = time.perf_counter()
    ux = np.unique(xs)
    ok = info.n_unique == ux.size and eng.target_value(1) == ux[0] and eng.target_value(ux.size) == ux[-1] \
        and eng.target_value(ux.size // 2) == ux[ux.size // 2 - 1]
    print(f"{name}: targets_set {1e3*(t1-t0):.3f} ms, n_unique {info.n_unique}, path {eng.stats()['sort_two_level']}, ok {ok}", flush=True)
