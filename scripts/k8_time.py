"""times sk_targets_set_device on 1e7 U(0,1) distances (CUDA events through the library's stream)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectralkernels_jl_b200 as sk
import torch
eng = sk.Session(0)
xs = np.random.default_rng(0).uniform(0, 1, 10_000_000)
d = torch.from_numpy(xs).cuda()
torch.cuda.synchronize()
best = 1e9
for rep in range(6):
    eng.timer_begin()
    info = eng.targets_set_device(d.data_ptr(), xs.size)
    ms = eng.timer_end()
    best = min(best, ms)
print(f"SK_K8_DBG={os.environ.get('SK_K8_DBG','0')}: targets_set {best:.3f} ms (n_unique {info.n_unique}, path {eng.stats()['sort_two_level']})")
