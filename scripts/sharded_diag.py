"""per-rank n_chained / timeline of a sharded resident step (torchrun): python -m torch.distributed.run ... scripts/sharded_diag.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import spectralkernels_jl_b200 as sk
from spectralkernels_jl_b200.sharded import LibComm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 10_000_000
xs = np.random.default_rng(rank).uniform(0, 1, n)
d_in = torch.from_numpy(xs).cuda(); d_v = torch.empty_like(d_in); d_e = torch.empty_like(d_in)
cfg = sk.AdaptiveKernelConfig(sk.Matern(1 / (np.pi / 2), 1.0, 1.5), device=local)
eng = cfg.engine
comm = LibComm.from_torch(eng)
def step():
    sk.kernel_values(cfg, None, k0=1.0, xs_device=(d_in.data_ptr(), n), out_device=(d_v.data_ptr(), d_e.data_ptr()), comm=comm)
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
st = eng.stats()
print(f"rank {rank}: {1e3 * dt:.4f} ms/step, n_chained {st['n_chained']}, speculated {st['n_speculated']}, rollbacks {st['n_spec_rollbacks']}, "
      f"prefetch {st['n_prefetch_issued']}/{st['n_prefetch_hits']}, mode {comm.mode}", flush=True)
dist.barrier(); comm.close(); dist.destroy_process_group()
