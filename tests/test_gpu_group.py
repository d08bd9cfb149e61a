"""
Device group (-m gpu): one caller, several GPUs behind the call sequence of a single context (sk_group_* in
include/spectralkernels_b200.h; the reference's API is one task calling kernel_values(cfg, xs), src/adaptive.jl:95-108).
A device may be listed more than once, so the chunking, the lock-step loop and the combination of the per-device
scalars are exercised on a one-GPU box too; with more GPUs visible the same checks run across real devices.

Bar: values and error estimates bit-identical to the single-context run over the same distances; identical traces when
no distance is shared between chunks.
"""
import numpy as np
import pytest

import closed_forms as cf
import sk_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sk():
    import spectralkernels_jl_b200 as sk
    return sk


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


def _key(trace):
    subs = [(t["a"], t["b"], t["accepted"]) for t in trace if t["kind"] == "subinterval"]
    pans = [(t["a"], t["b"], t["hi_before"], t["hi_after"], t["criteria"]) for t in trace if t["kind"] == "panel"]
    return subs, pans


def _run(sk, S, xs, engine, **kw):
    k0 = kw.pop("k0", None)
    cfg = sk.AdaptiveKernelConfig(S, engine=engine, **kw)
    tr = []
    k0 = k0 or sk.compute_k0(cfg)
    v, e = sk.kernel_values(cfg, xs, k0=k0, trace=tr)
    return v, e, tr, cfg


@pytest.mark.parametrize("ndev", [1, 2, 3, 8])
def test_group_matches_single_context_bitwise(sk, ndev):
    rng = np.random.default_rng(21)
    xs = rng.uniform(0.0, 1.0, 200_000)                  # distinct with probability one: same unique counts
    S = sk.Matern(1.0 / (np.pi / 2), 1.0, 1.5)
    v1, e1, t1, c1 = _run(sk, S, xs, sk.Session(0))
    vg, eg, tg, cg = _run(sk, S, xs, sk.GroupSession(_devices(ndev)))
    assert np.array_equal(v1, vg) and np.array_equal(e1, eg)
    assert _key(t1) == _key(tg)
    assert np.max(np.abs(vg - (1 + 2 * np.pi * xs) * np.exp(-2 * np.pi * xs))) <= 1e-8
    s1, sg = c1.engine.stats(), cg.engine.stats()
    assert sg["units"] == s1["units"] and sg["n_subintervals"] == s1["n_subintervals"]
    c1.engine.close(); cg.engine.close()


def test_group_shrinking_active_set_zero_lag_and_duplicates(sk):
    """Slow-decay Matern on log-spaced distances (several panels, devices run out of active targets at different
    times), with zero lags and duplicates inside and across chunks."""
    S = sk.Matern(1.0, 0.5, 0.55)
    Sh = lambda w: (0.25 + w ** 2) ** -1.05
    xs = 10 ** np.linspace(-4, 0, 600)
    xs = np.concatenate([xs[::3], [0.0], xs[1::3], xs[:40], [0.0], xs[2::3]])     # chunks see different ranges
    v1, e1, t1, c1 = _run(sk, S, xs, sk.Session(0))
    for ndev in (2, 5):
        vg, eg, tg, cg = _run(sk, S, xs, sk.GroupSession(_devices(ndev)))
        assert np.array_equal(v1, vg)
        assert np.array_equal(e1, eg, equal_nan=True)
        assert [(t["a"], t["b"], t["accepted"]) for t in tg if t["kind"] == "subinterval"] == \
               [(t["a"], t["b"], t["accepted"]) for t in t1 if t["kind"] == "subinterval"]
        cg.engine.close()
    # and the oracle on the same input
    ocfg = so.OracleConfig(Sh)
    k0 = so.compute_k0(ocfg)
    vo, _ = so.kernel_values(ocfg, xs, k0=k0)
    assert np.max(np.abs(v1 - vo)) <= 1e-9 * k0
    c1.engine.close()


def test_group_host_closure_bisection_and_tiny_inputs(sk):
    al = 2000.0
    xs = np.linspace(0.001, 0.05, 400)
    f = lambda w: np.exp(-al * np.abs(w))                 # a closure: strengths evaluated on the host, uploaded to all
    k0 = 2 / al
    v1, e1, t1, c1 = _run(sk, f, xs, sk.Session(0), k0=k0)
    vg, eg, tg, cg = _run(sk, f, xs, sk.GroupSession(_devices(3)), k0=k0)
    assert sum(1 for t in t1 if t["kind"] == "subinterval" and not t["accepted"]) >= 3     # rejected sub-intervals
    assert np.array_equal(v1, vg) and np.array_equal(e1, eg) and _key(t1) == _key(tg)
    assert cg.engine.stats()["n_spec_rollbacks"] >= 1
    # fewer distances than devices; a single distance; the direct branch (<= 2 active targets overall)
    S = sk.Exponential(1.0, 1.0)
    for r in (np.array([0.4, 0.9]), np.array([0.77]), np.array([0.3, 0.0, 1.7])):
        va, _, _, ca = _run(sk, S, r, sk.Session(0), k0=2.0)
        vb, _, _, cb = _run(sk, S, r, sk.GroupSession(_devices(4)), k0=2.0)
        assert np.array_equal(va, vb)
        assert cb.engine.stats()["n_direct"] == ca.engine.stats()["n_direct"]
        ca.engine.close(); cb.engine.close()
    c1.engine.close(); cg.engine.close()


def test_group_config_keyword_and_errors(sk):
    cfg = sk.AdaptiveKernelConfig(sk.Matern(), devices=_devices(2))
    assert isinstance(cfg.engine, sk.GroupSession)
    xs = np.random.default_rng(3).uniform(0, 2, 5000)
    v, _ = sk.kernel_values(cfg, xs, k0=np.pi / 2)
    assert np.max(np.abs(v - cf.readme_cov(xs))) <= 1e-8 * np.pi / 2
    with pytest.raises(sk.SkError):
        sk.kernel_values(cfg, np.array([0.1, -0.2, 0.3]), k0=1.0)
    with pytest.raises(NotImplementedError):
        sk.kernel_values(cfg, None, k0=1.0, points=np.zeros((4, 2)))
