"""
Scalar reductions for a target-sharded multi-GPU run of kernel_values.

Every rank passes its own chunk of the distances (no exchange of targets).  The adaptive loop stays in
lock step through scalar all-reduces: max of the largest unconverged distance (panel choice,
src/adaptive.jl:152), sum of the active counts (NUFFT-vs-direct cutoff, src/quadrature.jl:105), max of
max|I2-I1| (accept test, src/quadrature.jl:258-260) and max of the stopping distance of the convergence
scan (src/adaptive.jl:183-198).  Because every rank builds the
same transform geometry (sk_panel_set_range) the per-target arithmetic does not depend on the sharding:
the values are bit-identical to a single-GPU run over the union of the chunks.

Backend: torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence


class TorchComm:
    """max / min / sum all-reduces and an all-gather of a few doubles per rank.  The device buffers are
    allocated once; every call is one H2D copy, one collective, one D2H copy."""

    def __init__(self, device=None, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._torch, self._dist, self._group = torch, dist, group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if device is None:
            device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        self.device = device
        self.n_reductions = 0

    def _reduce(self, vals: Sequence[float], op) -> List[float]:
        t = self._torch.tensor(list(vals), dtype=self._torch.float64, device=self.device)
        self._dist.all_reduce(t, op=op, group=self._group)
        self.n_reductions += 1
        return t.cpu().tolist()

    def max(self, vals):
        return self._reduce(vals, self._dist.ReduceOp.MAX)

    def min(self, vals):
        return self._reduce(vals, self._dist.ReduceOp.MIN)

    def sum(self, vals):
        return self._reduce(vals, self._dist.ReduceOp.SUM)

    def gather(self, vals):
        t = self._torch.tensor(list(vals), dtype=self._torch.float64, device=self.device)
        out = self._torch.empty(self.world_size * t.numel(), dtype=self._torch.float64, device=self.device)
        self._dist.all_gather_into_tensor(out, t, group=self._group)
        self.n_reductions += 1
        return out.view(self.world_size, t.numel()).cpu().tolist()


class LibComm:
    """The scalar reductions of a target-sharded run done INSIDE libsk_b200, enqueued on the context's stream right
    behind the kernel that produced the local value.  Two transports:

    * `mode="peer"` (default on one node): single-warp exchange kernels over peer-mapped mailboxes in the ranks' HBM
      (NVLink / NVSwitch peer memory through CUDA IPC; sk_comm_peer_export / sk_comm_peer_attach, k_peer_exchange) --
      no NCCL call on the data path at all;
    * `mode="nccl"`: NCCL all-reduces on the context's stream (sk_comm_init) -- the A/B reference.

    Either way the per-sub-interval and per-scan reductions cost no extra host synchronisation; this object only serves
    the once-per-call reductions (global distance range, counts).

    `LibComm.from_torch(engine)` bootstraps through an initialised torch.distributed group (the mailboxes' IPC handles
    are all-gathered; for NCCL rank 0 creates the unique id and broadcasts it)."""
    fused = True

    def __init__(self, engine, rank: int, world_size: int, uid: bytes = None, peer_handles=None):
        self.engine, self.rank, self.world_size = engine, int(rank), int(world_size)
        if peer_handles is not None:
            engine.comm_peer_attach(peer_handles, rank, world_size)
            self.mode = "peer"
        else:
            engine.comm_init(uid, rank, world_size)
            self.mode = "nccl"
        self.n_reductions = 0

    @classmethod
    def from_torch(cls, engine, group=None, mode: str = None):
        import os
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        mode = mode or os.environ.get("SK_COMM_TRANSPORT", "peer")
        if mode == "peer" and world <= 16:
            # every rank must end up on the same transport: a rank that cannot export / map the mailboxes (no peer access,
            # IPC disabled in a container) makes ALL ranks fall back to the NCCL all-reduces
            try:
                mine = engine.comm_peer_export()
            except Exception:
                mine = None
            handles = [None] * world
            dist.all_gather_object(handles, mine, group=group)      # also: every mailbox is zeroed before anyone writes
            comm, ok = None, all(h is not None for h in handles)
            if ok:
                try:
                    comm = cls(engine, rank, world, peer_handles=handles)
                except Exception:
                    ok = False
            flags = [None] * world
            dist.all_gather_object(flags, bool(ok), group=group)
            if all(flags):
                return comm
            engine.comm_destroy()
        box = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(engine, rank, world, uid=box[0])

    def _r(self, vals, op):
        self.n_reductions += 1
        return self.engine.comm_allreduce(vals, op)

    def max(self, vals):
        return self._r(vals, 0)

    def min(self, vals):
        return self._r(vals, 1)

    def sum(self, vals):
        return self._r(vals, 2)

    def gather(self, vals):
        """every rank's values, through ONE sum all-reduce of a one-hot layout (adding zeros is exact)"""
        vals = list(vals)
        k = len(vals)
        if self.mode == "peer" and k <= 7:
            self.n_reductions += 1
            return self.engine.comm_allgather(vals, self.world_size)
        if k * self.world_size > 32:
            raise ValueError("too many values for sk_comm_allreduce")
        buf = [0.0] * (k * self.world_size)
        buf[self.rank * k:(self.rank + 1) * k] = vals
        out = self._r(buf, 2)
        return [out[r * k:(r + 1) * k] for r in range(self.world_size)]

    def close(self):
        self.engine.comm_destroy()
