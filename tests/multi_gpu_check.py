"""
Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank evaluates its own chunk of the distances in one target-sharded kernel_values call (scalar NCCL
all-reduces only); rank 0 then evaluates the union on its GPU alone.  The sharded values and error
estimates must equal the single-GPU ones bit for bit and the panel traces must be identical.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralkernels_jl_b200 as sk  # noqa: E402
from spectralkernels_jl_b200.sharded import LibComm, TorchComm  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    flavour = os.environ.get("SK_COMM", "lib")
    for name, S, gen, k0, kw in (
        ("matern_uniform", sk.Matern(1 / (np.pi / 2), 1.0, 1.5), lambda rng, n: rng.uniform(0, 1, n), 1.0, {}),
        ("slow_decay_logspaced", sk.Matern(1.0, 0.5, 0.55), lambda rng, n: 10 ** rng.uniform(-4, 0, n), 5.9, {}),
        # dim = 2: the O(N) nonuniform Hankel transform (octave groups are built from the global distance range)
        ("matern_2d_hankel", sk.Matern(1.0, 1.0, 1.5, d=2), lambda rng, n: rng.uniform(0, 1, n), 2.0943951023931953,
         {"dim": 2}),
    ):
        n = 400_000
        chunks = [gen(np.random.default_rng(100 + r), n + 1000 * r) for r in range(world)]
        if name == "slow_decay_logspaced":
            chunks[world - 1] = chunks[world - 1] * 1e-2          # the last rank runs out of active targets early
        cfg = sk.AdaptiveKernelConfig(S, device=local, **kw)
        comm = LibComm.from_torch(cfg.engine) if flavour == "lib" else TorchComm(device=torch.device("cuda", local))
        tr = []
        v, e = sk.kernel_values(cfg, chunks[rank], k0=k0, comm=comm, trace=tr)
        if flavour == "lib":
            dist.barrier()
            comm.close()
        # gather everything on rank 0 (padded to the longest chunk)
        m = max(c.size for c in chunks)
        buf = torch.zeros(2, m, dtype=torch.float64, device=f"cuda:{local}")
        buf[0, : v.size] = torch.from_numpy(v)
        buf[1, : e.size] = torch.from_numpy(e)
        out = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        keys = [None] * world
        dist.all_gather_object(keys, [(t["a"], t["b"], t.get("accepted"), t.get("criteria")) for t in tr])
        if rank == 0:
            union = np.concatenate(chunks)
            cfg1 = sk.AdaptiveKernelConfig(S, device=local, **kw)
            tr1 = []
            v1, e1 = sk.kernel_values(cfg1, union, k0=k0, trace=tr1)
            vs = np.concatenate([out[r][0, : chunks[r].size].cpu().numpy() for r in range(world)])
            es = np.concatenate([out[r][1, : chunks[r].size].cpu().numpy() for r in range(world)])
            same_v = np.array_equal(vs, v1)
            same_e = np.array_equal(np.nan_to_num(es, nan=-1), np.nan_to_num(e1, nan=-1))
            key1 = [(t["a"], t["b"], t.get("accepted"), t.get("criteria")) for t in tr1]
            same_t = all(k == key1 for k in keys)
            npan = sum(1 for t in tr1 if t["kind"] == "panel")
            print(f"[multi_gpu_check] {name}: world={world} values_bitwise={same_v} errs_bitwise={same_e} "
                  f"traces_equal={same_t} panels={npan} comm={flavour}/{getattr(comm, "mode", "torch")} host_reductions={comm.n_reductions} "
                  f"max|dv|={np.max(np.abs(vs - v1)):.2e}", flush=True)
            ok = ok and same_v and same_e and same_t
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
