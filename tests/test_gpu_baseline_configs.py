"""
Parity tests at the stated size of BASELINE.json configs 3, 4 and 5 (-m gpu), through the C ABI:

  config 3  singular Matern alpha = 0.5 on the 49 995 000 pairwise distances of 1e4 random 2-D points (seed 0),
            as 1-D kernel and with dim = 2: a strided subsample is re-evaluated by the oracle (direct sums,
            <= 1e-11 K(0) / 2e-11 K(0) with Bessel sums; identical sub-interval trace) and checked against the
            closed form `sing_matern_cov` (scripts/matern_pair.jl:20-33) at 10 tol (test/matern_sdf.jl:62)
  config 4  1e6 distances: K, K'(r), dK/dphi, dK/drho, dK/dnu against mpmath derivatives of `matern_cov`
            (scripts/matern_pair.jl:7-15) on a subsample; upstream threshold 1e-5 (test/derivatives/sdf_params.jl:22-24)
  config 5  the KNN-15 pair list of 1e5 random 2-D points (scripts/fit_vecchia_demo.jl:40-41), dim = 2 Matern,
            every pair against `matern_cov(d = 2)` at 10 tol (test/matern_sdf.jl:27)

and the NaN / Inf semantics of the truncation bound (Julia's `min`, src/adaptive.jl:225-228) on the device.
"""
import math
import os
import sys

import numpy as np
import pytest

import closed_forms as cf
import sk_oracle as so

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sk():
    import spectralkernels_jl_b200 as sk
    return sk


def _subs(trace):
    return [(t["a"], t["b"], t["accepted"]) for t in trace if t["kind"] == "subinterval"]


def _pans(trace):
    return [(t["a"], t["b"], t["criteria"]) for t in trace if t["kind"] == "panel"]


def _pair_of(t, n):
    """(i, j) of the t-th pair of the strict upper triangle in row-major order (the order of sk_targets_set_pairs
    with pairs = NULL)."""
    t = np.asarray(t, dtype=np.int64)
    i = np.floor(((2 * n - 1) - np.sqrt((2.0 * n - 1) ** 2 - 8.0 * t)) / 2).astype(np.int64)
    i = np.where(i * (2 * n - i - 1) // 2 > t, i - 1, i)
    i = np.where((i + 1) * (2 * n - i - 2) // 2 <= t, i + 1, i)
    j = t - i * (2 * n - i - 1) // 2 + i + 1
    return i, j


@pytest.mark.parametrize("dim", [1, 2])
def test_config3_pairwise_singular_matern(sk, dim):
    npts, alpha, tol = 10_000, 0.5, 1e-8
    parms = (1.0, 1.0, 1.5)
    pts = np.random.default_rng(0).uniform(0, 1, (npts, 2))
    npairs = npts * (npts - 1) // 2
    S = sk.Matern(*parms, d=dim)
    Sh = lambda w: cf.matern_sdf(w, parms, d=dim)
    cfg = sk.AdaptiveKernelConfig(S, alpha=alpha, dim=dim, tol=tol)
    ocfg = so.OracleConfig(Sh, alpha=alpha, dim=dim, tol=tol)
    k0 = so.compute_k0(ocfg)
    assert abs(sk.compute_k0(cfg) - k0) <= 1e-9 * k0
    tg = []
    vals, errs = sk.kernel_values(cfg, None, k0=k0, points=pts, trace=tg)
    assert vals.size == npairs and np.all(np.isfinite(vals)) and np.all(np.isfinite(errs))
    st = cfg.engine.stats()
    assert st["n_direct"] == 0 and (st["n_hankel"] > 0) == (dim == 2)
    # strided subsample of the pair list (+ the pair with the largest lag, so that the panel sequence is the same)
    nsub = 2000 if dim == 1 else 160                  # dim = 2: the oracle's Bessel sums cost ~0.1 s per target
    t = np.linspace(0, npairs - 1, nsub).astype(np.int64)
    i, j = _pair_of(t, npts)
    lag = np.sqrt((pts[i, 0] - pts[j, 0]) ** 2 + (pts[i, 1] - pts[j, 1]) ** 2)
    # (the device computes the lags itself; they may differ from numpy's in the last bit -- sqrt of a sum -- which moves
    # K by |K'(r)| ulp(r), far below the bounds used here)
    # the true maximum lag of the full set is what fixes the panels: read it back from the session
    r_max = cfg.engine.target_value(int(cfg.engine._last_targets[0].n_unique))
    lag_o = np.append(lag, r_max)
    to = []
    vo, eo = so.kernel_values(ocfg, lag_o, k0=k0, trace=to)
    lim = (1e-11 if dim == 1 else 2e-11) * k0
    assert np.max(np.abs(vals[t] - vo[:-1])) <= lim
    assert _subs(tg) == _subs(to) and _pans(tg) == _pans(to)                # same sub-intervals, same decisions
    # closed form (r rho <= 2: scripts/matern_pair.jl:22), reference tolerance 10 tol (test/matern_sdf.jl:62)
    sel = np.arange(0, nsub, 4 if dim == 1 else 1)
    true = cf.sing_matern_cov(lag[sel], (*parms, -alpha), d=dim)
    assert np.all(np.abs(vals[t[sel]] - true) / k0 <= 10 * tol)
    # the same subsample evaluated alone on the device
    v_s, _ = sk.kernel_values(sk.AdaptiveKernelConfig(S, alpha=alpha, dim=dim, tol=tol), lag_o, k0=k0)
    assert np.max(np.abs(v_s[:-1] - vals[t])) <= 1e-12 * k0


def _matern_cov_mp(t, phi, alpha, v, d=1):
    import mpmath as mp
    constant = mp.pi ** (mp.mpf(d) / 2) * phi / (2 ** (v - 1) * mp.gamma(v + mp.mpf(d) / 2) * alpha ** (2 * v))
    arg = alpha * 2 * mp.pi * abs(t)
    if arg == 0:
        return constant * 2 ** (v - 1) * mp.gamma(v)
    return constant * mp.besselk(v, arg) * arg ** v


def test_config4_derivatives_1e6(sk):
    import mpmath as mp
    n = 1_000_000
    xs = np.random.default_rng(0).uniform(0, 1, n)
    parms = (1.0 / (math.pi / 2), 1.0, 1.5)                    # K(0) = 1
    tol = 1e-8
    cfg = sk.AdaptiveKernelConfig(sk.Matern(*parms), tol=tol)
    k0 = 1.0
    tr = []
    v, _ = sk.kernel_values(cfg, xs, k0=k0, trace=tr)
    dk = sk.kernel_derivative(cfg, xs, k0, reuse_targets=True)
    dphi, drho, dnu = sk.kernel_sdf_derivatives(cfg, xs, k0, reuse_targets=True)
    # closed forms of K and K' on all 1e6 lags (Matern nu = 3/2)
    true = (1 + 2 * math.pi * xs) * np.exp(-2 * math.pi * xs)
    dtrue = -(2 * math.pi) ** 2 * xs * np.exp(-2 * math.pi * xs)
    assert np.max(np.abs(v - true)) <= 10 * tol * k0                         # test/matern_sdf.jl:27
    assert np.max(np.abs(dk - dtrue)) <= 10 * tol * k0                       # test/matern_sdf.jl:34
    assert np.max(np.abs(dphi - true / parms[0])) <= 10 * tol / parms[0]     # K is linear in phi
    # mpmath derivatives of matern_cov (ForwardDiff in the reference) on a strided subsample
    sel = np.arange(0, n, n // 150)
    ph, rh, nu_ = [mp.mpf(p) for p in parms]
    d_rho, d_nu = [], []
    with mp.workdps(40):
        for t in xs[sel]:
            t = mp.mpf(float(t))
            d_rho.append(float(mp.diff(lambda q: _matern_cov_mp(t, ph, q, nu_), rh)))
            d_nu.append(float(mp.diff(lambda q: _matern_cov_mp(t, ph, rh, q), nu_)))
    assert np.max(np.abs(drho[sel] - np.array(d_rho))) < 1e-5                # upstream threshold (sdf_params.jl:22-24)
    assert np.max(np.abs(dnu[sel] - np.array(d_nu))) < 1e-5
    assert np.max(np.abs(drho[sel] - np.array(d_rho))) <= 100 * tol and np.max(np.abs(dnu[sel] - np.array(d_nu))) <= 100 * tol
    # oracle parity of one derivative integrand on a subsample (own adaptive run, own trace: src/derivatives.jl:68-71)
    sub = np.append(xs[sel], xs.max())
    ex = -parms[2] - 0.5
    f_nu = lambda w: -parms[0] * (parms[1] ** 2 + w ** 2) ** ex * np.log(parms[1] ** 2 + w ** 2)
    ocfg = so.gen_new_sdf_config(so.OracleConfig(lambda w: cf.matern_sdf(w, parms), tol=tol), f_nu)
    vo, _ = so.kernel_values(ocfg, sub, k0=k0, param_derivative=True)
    assert np.max(np.abs(dnu[sel] - vo[:-1])) <= 1e-10 * max(1.0, float(np.max(np.abs(vo))))
    # equal lags get equal derivative values (the scatter to the input order)
    xs2 = xs.copy()
    xs2[1::2] = xs[0:-1:2]
    d2 = sk.kernel_derivative(cfg, xs2, k0)
    assert np.array_equal(d2[1::2], d2[0:-1:2])


def test_config5_vecchia_pair_list(sk):
    sys.path.insert(0, ROOT)
    from bench_vecchia import knn_pairs
    npts, tol = 100_000, 1e-8
    pts = np.random.default_rng(0).uniform(0, 1, (npts, 2))
    pairs = knn_pairs(pts)
    assert 1.3e7 < pairs.shape[0] < 1.4e7
    parms = (1.0, 4.0, 1.5)
    cfg = sk.AdaptiveKernelConfig(sk.Matern(*parms, d=2), dim=2, tol=tol)
    k0 = sk.compute_k0(cfg)
    assert abs(k0 - cf.matern_cov(0.0, parms, d=2)[0]) <= 1e-9 * k0
    tr = []
    v, e = sk.kernel_values(cfg, None, k0=k0, points=pts, pairs=pairs, trace=tr)
    d = pts[pairs[:, 0]] - pts[pairs[:, 1]]
    lag = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)
    assert cfg.engine._last_targets[0].n_unique == np.unique(lag).size      # the device de-duplicates like unique()
    same = pairs[:, 0] == pairs[:, 1]
    assert same.sum() >= npts and np.all(v[same] == k0) and np.all(np.isnan(e[same]))    # src/adaptive.jl:133-146
    true = cf.matern_cov(lag, parms, d=2)
    assert np.max(np.abs(v - true)) / k0 <= 10 * tol                         # test/matern_sdf.jl:27 with dim = 2
    assert cfg.engine.stats()["n_hankel"] > 0
    # oracle (direct Bessel sums) on a small subsample that keeps the panel sequence
    sel = np.linspace(0, lag.size - 1, 60).astype(np.int64)
    sub = np.append(lag[sel][lag[sel] > 0], lag.max())
    ocfg = so.OracleConfig(lambda w: cf.matern_sdf(w, parms, d=2), dim=2, tol=tol)
    to = []
    vo, _ = so.kernel_values(ocfg, sub, k0=k0, trace=to)
    assert np.max(np.abs(v[sel][lag[sel] > 0] - vo[:-1])) <= 2e-11 * k0
    # the first panel is fixed by the largest lag; later ones by the largest lag still unconverged, which a
    # 60-point subsample does not share with 1.36e7 pairs -- so only the first sub-interval is comparable
    assert _subs(tr)[0] == _subs(to)[0] and all(s[2] for s in _subs(tr)) and all(s[2] for s in _subs(to))


def test_truncation_bound_nan_semantics_on_device(sk):
    """src/adaptive.jl:225-228: Julia's min(a, b) propagates NaN, so a NaN truncation bound keeps a target active
    (`trunc_err < tol` is false) where C's fmin would drop the NaN.  Checked through sk_converge_scan (separate scan
    kernel) and through the speculative scan fused into the interpolation kernel."""
    from spectralkernels_jl_b200._capi import ScanArgs, SK_CRIT, SK_KERNEL_COS
    xs = np.linspace(0.05, 1.0, 4000)
    nan, inf = float("nan"), float("inf")
    eng = sk.Session(0)
    eng.rule_set(4096, 16, 0.0)
    S = sk.Matern(1.0, 1.0, 1.5)
    eng.sdf_builtin(S.family, S.params, 0)
    info = eng.targets_set(xs)
    n = int(info.n_unique)
    tau = 10.0                                  # every |panel_k| < tau (K(0) = pi/2): the decision rests on the truncation bound
    cases = [(nan, 1e-3, n), (1e-3, nan, n), (nan, nan, n), (inf, nan, n), (nan, -inf, n),
             (inf, 1e-3, 0), (1e-3, inf, 0), (-inf, 1e-3, 0), (1e-30, 1e-30, 0), (inf, inf, n), (inf, 50.0, None)]
    for crit in ("both", "tails"):
        for ta, tn, expect in cases:
            args = ScanArgs(ta, tn, 1.0, tau, SK_CRIT[crit], 0)
            with np.errstate(all="ignore"):
                te = np.minimum(ta, tn / (2 * np.pi * xs))                  # numpy's minimum propagates NaN like Julia's
            unconv = ~(te < tau)
            want = int(np.max(np.nonzero(unconv)[0]) + 1) if unconv.any() else 0
            if expect is not None:
                assert want == expect
            for speculate in (False, True):
                eng.run_begin()
                eng.panel_begin(1, n)
                eng.subinterval(0.0, 16384.0, 2.0, 0.0, SK_KERNEL_COS, False, speculate=args if speculate else None)
                eng.subinterval_accept()
                eng.panel_commit()
                new_hi, _ = eng.converge_scan(args)
                eng.converge_apply(args, new_hi)
                assert new_hi == want, (crit, ta, tn, speculate, new_hi, want)
                if speculate:
                    assert eng.stats()["n_speculated"] == 1
    eng.close()
