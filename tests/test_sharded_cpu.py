"""
world_size-2 gloo tests of the host logic of a target-sharded kernel_values run: every rank holds its
own chunk of the distances, the adaptive loop runs in lock step through scalar all-reduces
(spectralkernels.jl_b200/sharded.py), and the result equals the single-process run over the union --
values bit-for-bit, panel traces identical.  Per-target work is done by tests/fake_engine.py (oracle
direct sums) because there is no GPU here; the driver under test is the product's.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, chunks, kw, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import spectralkernels_jl_b200 as sk
    from spectralkernels_jl_b200.sharded import TorchComm
    from fake_engine import FakeEngine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        S = lambda w: np.exp(-np.abs(w)) if kw.get("sdf") == "exp" else (0.25 + w ** 2) ** -1.05
        cfg = sk.AdaptiveKernelConfig(S, engine=FakeEngine(), quadspec=(256, 4), **kw.get("cfg", {}))
        comm = TorchComm()
        trace = []
        vals, errs = sk.kernel_values(cfg, chunks[rank], k0=kw["k0"], comm=comm, trace=trace)
        ret[rank] = (vals, errs, trace, comm.n_reductions)
    finally:
        dist.destroy_process_group()


def _single(xs, kw):
    import spectralkernels_jl_b200 as sk
    from fake_engine import FakeEngine
    S = lambda w: np.exp(-np.abs(w)) if kw.get("sdf") == "exp" else (0.25 + w ** 2) ** -1.05
    cfg = sk.AdaptiveKernelConfig(S, engine=FakeEngine(), quadspec=(256, 4), **kw.get("cfg", {}))
    trace = []
    vals, errs = sk.kernel_values(cfg, xs, k0=kw["k0"], trace=trace)
    return vals, errs, trace


def _key(trace, with_hi=False):
    subs = [(t["a"], t["b"], t["accepted"]) for t in trace if t["kind"] == "subinterval"]
    pans = [(t["a"], t["b"], t["criteria"]) + ((t["hi_before"], t["hi_after"]) if with_hi else ())
            for t in trace if t["kind"] == "panel"]
    return subs, pans


@pytest.mark.parametrize("case", ["slow_decay_logspaced", "exp_with_zero_and_empty_rank_tail"])
def test_two_rank_gloo_equals_single(case):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(5)
    if case == "slow_decay_logspaced":
        xs = 10 ** rng.uniform(-3, 0, 90)
        chunks = [xs[:50], xs[50:]]
        kw = {"k0": 5.9, "sdf": "matern"}
    else:
        # rank 1 only holds small distances: its active set empties panels before rank 0's does
        xs = np.concatenate([rng.uniform(0.5, 3.0, 30), [0.0, 0.0], rng.uniform(1e-3, 2e-2, 25)])
        chunks = [xs[:32], xs[32:]]
        kw = {"k0": 2.0, "sdf": "exp"}
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, chunks, kw, ret), nprocs=2, join=True)
    v1, e1, t1 = _single(xs, kw)
    v2 = np.concatenate([ret[0][0], ret[1][0]])
    e2 = np.concatenate([ret[0][1], ret[1][1]])
    assert np.array_equal(v1, v2)                                  # bit-identical values
    assert np.array_equal(np.isnan(e1), np.isnan(e2)) and np.array_equal(np.nan_to_num(e1), np.nan_to_num(e2))
    assert _key(ret[0][2]) == _key(ret[1][2]) == _key(t1)          # same panels and accept decisions on every rank
    assert ret[0][3] == ret[1][3] > 0                              # same number of scalar reductions
    # the active counts add up to the single-process ones
    p0 = [t for t in ret[0][2] if t["kind"] == "panel"]
    p1 = [t for t in ret[1][2] if t["kind"] == "panel"]
    ps = [t for t in t1 if t["kind"] == "panel"]
    off0 = 1 if np.any(chunks[0] == 0) else 0
    off1 = 1 if np.any(chunks[1] == 0) else 0
    offs = 1 if np.any(xs == 0) else 0
    for a, b, s in zip(p0, p1, ps):
        dup = 0
        assert (a["hi_after"] - off0) + (b["hi_after"] - off1) - dup == s["hi_after"] - offs


def test_driver_with_fake_engine_matches_oracle_driver():
    """The product's host driver (adaptive.py) against the oracle's independent restatement
    (oracle/sk_oracle.py) when both use the same per-target arithmetic: identical values and traces."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sk_oracle as so
    import spectralkernels_jl_b200 as sk
    from fake_engine import FakeEngine
    S = lambda w: (0.25 + w ** 2) ** -1.05
    xs = np.concatenate([[0.0, 0.7, 0.7], 10 ** np.linspace(-3, 0, 40)])
    dS = lambda w: -2.1 * w * (0.25 + w ** 2) ** -2.05
    for kw in ({}, {"derivative": True}, {"alpha": 0.4}, {"convergence_criteria": "tails"},
               {"convergence_criteria": "panel"}, {"tol": 1e-5}, {"alpha": 0.4, "logw": True, "df": dS},
               {"logw": True, "df": dS}):
        cfg = sk.AdaptiveKernelConfig(S, engine=FakeEngine(), quadspec=(256, 4), **kw)
        ocfg = so.OracleConfig(S, quadspec=(256, 4), **kw)
        tg, to = [], []
        vg, eg = sk.kernel_values(cfg, xs, k0=5.9, trace=tg)
        vo, eo = so.kernel_values(ocfg, xs, k0=5.9, trace=to)
        assert np.array_equal(vg, vo), kw
        assert _key(tg, True) == _key(to, True), kw
    # dim = 2: Bessel kernels (src/quadrature.jl:137-161, :176-180, :252-254), direct Bessel sums on both sides
    S2 = lambda w: (0.25 + w ** 2) ** -2.0
    xs2 = np.concatenate([[0.0], 10 ** np.linspace(-2, 0, 25)])
    for kw in ({"dim": 2}, {"dim": 2, "derivative": True}, {"dim": 2, "alpha": 0.5}, {"dim": 4}):
        cfg = sk.AdaptiveKernelConfig(S2, engine=FakeEngine(), quadspec=(256, 4), **kw)
        ocfg = so.OracleConfig(S2, quadspec=(256, 4), **kw)
        tg, to = [], []
        vg, eg = sk.kernel_values(cfg, xs2, k0=3.0, trace=tg)
        vo, eo = so.kernel_values(ocfg, xs2, k0=3.0, trace=to)
        assert np.array_equal(vg, vo), kw
        # error estimates contain 2*trunc_err, whose (c, d) come from two different evaluations of the same
        # rank-one least-squares fit (closed form vs. lstsq): equal to ~1e-13 relative, not bit for bit
        assert np.array_equal(np.isnan(eg), np.isnan(eo)), kw
        assert np.allclose(np.nan_to_num(eg), np.nan_to_num(eo), rtol=1e-10, atol=0), kw
        assert _key(tg, True) == _key(to, True), kw


def test_first_panel_scalars_are_prepared_while_the_targets_are_sorted():
    """sk_targets_begin / _early_range / _end in the host driver (adaptive.py): the tail fit of the first panel is evaluated
    between the two halves (while the device sorts), the panel loop then finds it; values, errors and the trace equal the
    blocking call sequence bit for bit."""
    import spectralkernels_jl_b200 as sk
    from spectralkernels_jl_b200 import adaptive as ad
    from fake_engine import FakeEngine
    S = lambda w: (1.0 + w ** 2) ** -1.3
    xs = np.concatenate([[0.0], 10 ** np.random.default_rng(2).uniform(-3, 0, 40)])
    out = []
    for overlap in (False, True):
        eng = FakeEngine()
        log = []
        real = ad.estimate_tail_decay
        ad.OVERLAP_HOST_WORK = overlap
        ad.estimate_tail_decay = lambda *a, **k: (log.append(len(eng.calls)), real(*a, **k))[1]
        try:
            cfg = sk.AdaptiveKernelConfig(S, engine=eng, quadspec=(256, 4))
            tr = []
            v, e = sk.kernel_values(cfg, xs, k0=1.9, trace=tr)
        finally:
            ad.OVERLAP_HOST_WORK = True
            ad.estimate_tail_decay = real
        key = [(t["kind"], t["a"], t["b"], t.get("accepted"), t.get("hi_after"), t.get("criteria")) for t in tr]
        out.append((v, e, key, [c[0] for c in eng.calls], log))
    (v0, e0, k0_, c0, l0), (v1, e1, k1_, c1, l1) = out
    assert np.array_equal(v0, v1) and np.array_equal(np.nan_to_num(e0, nan=-1), np.nan_to_num(e1, nan=-1)) and k0_ == k1_
    assert "targets_begin" not in c0 and c1[:3] == ["targets_begin", "targets_early_range", "targets_end"]
    # with the halves, the first tail fit ran after targets_early_range and BEFORE targets_end (2 engine calls logged so far)
    assert l1[0] == 2 and len(l1) == len(l0)
