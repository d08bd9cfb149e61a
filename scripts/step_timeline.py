"""Host-side timeline of one resident kernel_values step (config 2): every library call with its duration, and the host
time between calls (the GPU is idle there: each call ends with a synchronisation).  python scripts/step_timeline.py [n]"""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import torch
import spectralkernels_jl_b200 as sk
from spectralkernels_jl_b200 import adaptive as ad
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "resident"          # resident | e2e
if len(sys.argv) > 3 and sys.argv[3] == "nooverlap":
    ad.OVERLAP_HOST_WORK = False
xs = np.random.default_rng(0).uniform(0, 1, n)
d_in = torch.from_numpy(xs).cuda()
d_v = torch.empty_like(d_in); d_e = torch.empty_like(d_in)
cfg = sk.AdaptiveKernelConfig(sk.Matern(1 / (np.pi / 2), 1.0, 1.5))
eng = cfg.engine
log = []
def wrap(obj, name, label=None):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); t1 = time.perf_counter()
        log.append((label or name, t0, t1)); return r
    setattr(obj, name, g)
for name in dir(eng):
    if not name.startswith("_") and callable(getattr(eng, name)) and name not in ("stats",):
        try: wrap(eng, name)
        except Exception: pass
wrap(ad, "estimate_tail_decay", "HOST estimate_tail_decay"); wrap(ad, "_scan_args", "HOST _scan_args")
hin, hv, he = sk.PinnedArray(n), sk.PinnedArray(n), sk.PinnedArray(n)
hin.array[:] = xs
def step():
    if mode == "e2e":
        sk.kernel_values(cfg, hin.array, k0=1.0, out_vals=hv.array, out_errs=he.array)
    else:
        sk.kernel_values(cfg, None, k0=1.0, xs_device=(d_in.data_ptr(), n), out_device=(d_v.data_ptr(), d_e.data_ptr()))
for _ in range(5): step()
torch.cuda.synchronize()
tot = {}
R = 20
T0 = time.perf_counter()
for _ in range(R):
    log.clear(); t_begin = time.perf_counter(); step(); torch.cuda.synchronize(); t_end = time.perf_counter()
    prev = t_begin
    for name, t0, t1 in log:
        if name.startswith("HOST"):
            tot[name] = tot.get(name, 0) + (t1 - t0); continue      # nested inside the gaps
        tot["gap before " + name] = tot.get("gap before " + name, 0) + (t0 - prev)
        tot[name] = tot.get(name, 0) + (t1 - t0); prev = t1
    tot["tail"] = tot.get("tail", 0) + (t_end - prev)
wall = (time.perf_counter() - T0) / R
print(f"n = {n} {mode} overlap={ad.OVERLAP_HOST_WORK}: wall {1e3 * wall:.3f} ms per step; per-step averages (us):")
for k, v in tot.items():
    print(f"  {1e6 * v / R:9.1f}  {k}")
