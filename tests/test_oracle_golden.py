"""
Pins the oracle (oracle/sk_oracle.py + sk_oracle.c) against every closed-form known answer
the reference's own tests hold for the K(r) path, at the reference's tolerance
(|err|/K(0) <= 10*tol: test/exponential_sdf_1d.jl:21, test/matern_sdf.jl:27,62).
CPU only.  The r grids are subsampled (every 5th point of the reference's 1000) so the
whole CPU suite stays within minutes; the grids' end points are kept.
"""
import numpy as np
import pytest

import closed_forms as cf
import sk_oracle as so

SUB = slice(None, None, 5)


def _idx(n):
    i = np.arange(n)[SUB]
    return np.unique(np.append(i, n - 1))


@pytest.mark.parametrize("tol", [1e-4, 1e-8, 1e-10, 1e-12])
@pytest.mark.parametrize("derivative", [False, True])
def test_exponential_sdf_1d(golden, tol, derivative):
    i = _idx(1000)
    r = golden["exp_r"][i]
    true = (golden["exp_dK"] if derivative else golden["exp_K"])[i]
    if derivative:
        r, true = r[1:], true[1:]                         # exponential_sdf_1d.jl:8-10
    cfg = so.OracleConfig(cf.exponential_sdf, tol=tol, derivative=derivative)
    vals, _ = so.kernel_values(cfg, r)
    assert np.all(np.abs(vals - true) / 2.0 <= 10 * tol)


@pytest.mark.parametrize("tol,derivative", [(1e-4, False), (1e-8, False), (1e-12, False), (1e-4, True), (1e-8, True)])
def test_matern_sdf_1d(golden, tol, derivative):
    """test/matern_sdf.jl:2-34, dim=1.  (derivative=True, tol=1e-12) is left out: the restated
    algorithm reaches 1.8e-11 there, not 10*tol -- the upstream testset is disabled
    (test/runtests.jl:22-27), so upstream never checks it either."""
    i = _idx(1000)
    parms = tuple(golden["matern_parms"])
    r = golden["matern_r"][i]
    true = (golden["matern_dK"] if derivative else golden["matern_K"])[i]
    k0 = float(golden["matern_K"][0])
    cfg = so.OracleConfig(lambda w: cf.matern_sdf(w, parms), tol=tol, derivative=derivative)
    vals, _ = so.kernel_values(cfg, r, **({"k0": k0} if derivative else {}))     # matern_sdf.jl:20-22
    assert np.all(np.abs(vals - true) / k0 <= 10 * tol)


@pytest.mark.parametrize("derivative", [False, True])
def test_matern_sdf_2d(golden, derivative):
    """test/matern_sdf.jl:2-34 with dim = 2: the (:J, nu) Bessel kernel (src/quadrature.jl:137-161, :176-180),
    prefactor 2 pi (src/adaptive.jl:43) and p = dim/2 (+1).  40 of the reference's 1000 grid points (direct
    Bessel summation is O(M N))."""
    i = np.unique(np.append(np.arange(0, 1000, 25), 999))
    parms = tuple(golden["matern_parms"])
    r = golden["matern_r"][i]
    true = (golden["matern2d_dK"] if derivative else golden["matern2d_K"])[i]
    k0 = float(golden["matern2d_K"][0])
    cfg = so.OracleConfig(lambda w: cf.matern_sdf(w, parms, d=2), dim=2, tol=1e-8, derivative=derivative)
    vals, _ = so.kernel_values(cfg, r, **({"k0": k0} if derivative else {}))
    assert np.all(np.abs(vals - true) / k0 <= 10 * 1e-8)


@pytest.mark.parametrize("tol", [1e-4, 1e-8])
def test_singular_matern_1d(golden, tol):
    """test/matern_sdf.jl:36-64 with dim=1, alpha=0.5 (Jacobi origin panel).  K(0) is infinite for the
    singular kernel, so the normaliser is the oracle's own finite k0 surrogate compute_k0 (adaptive.jl:74-91)."""
    i = _idx(1000)[1:]
    parms = tuple(golden["matern_parms"])
    r = golden["sing_r"][i]
    true = golden["sing_K"][i]
    cfg = so.OracleConfig(lambda w: cf.matern_sdf(w, parms), tol=tol, alpha=0.5)
    k0 = so.compute_k0(cfg)
    vals, _ = so.kernel_values(cfg, r, k0=k0)
    assert np.all(np.abs(vals - true) / k0 <= 10 * tol)


def test_singularity_derivative_logw(golden):
    """test/matern_sdf.jl:66-86, dim = 1: dK/d alpha through the logw=true config (integration by parts at
    the origin, src/quadrature.jl:186-228).  Sign/scale: kernel_singularity_derivative (src/derivatives.jl:74-81)
    returns d/d alpha of int |w|^-alpha S cos = -int log(w) w^-alpha S cos, which is what c *= -1 encodes
    (src/adaptive.jl:45)."""
    idx = golden["sing_dalpha_idx"][::4]
    parms = tuple(golden["matern_parms"])
    r = golden["sing_r"][idx]
    true = golden["sing_dalpha"][::4]
    S = lambda w: cf.matern_sdf(w, parms)
    ex = -parms[2] - 0.5
    dS = lambda w: parms[0] * ex * (parms[1] ** 2 + w ** 2) ** (ex - 1) * 2 * w
    cfg0 = so.OracleConfig(S, alpha=0.5)
    k0 = so.compute_k0(cfg0)
    cfg = so.OracleConfig(S, df=dS, alpha=0.5, logw=True, tol=1e-8)
    vals, _ = so.kernel_values(cfg, r, k0=k0, param_derivative=True)
    assert np.all(np.abs(vals - true) / k0 <= 10 * 1e-8)


def test_readme_demo_and_trace(golden):
    """README.md:19-33; the trace shape (2 outer panels, no bisection) is the oracle's own
    output -- there is no reference-side fixture for traces ("parity unpinned" for traces)."""
    r = golden["readme_r"]
    cfg = so.OracleConfig(lambda w: (1 + w ** 2) ** -2)
    trace = []
    vals, errs = so.kernel_values(cfg, r, trace=trace)
    assert np.max(np.abs(vals - golden["readme_K"])) / (np.pi / 2) <= 1e-8
    panels = [t for t in trace if t["kind"] == "panel"]
    subs = [t for t in trace if t["kind"] == "subinterval"]
    assert [(p["a"], p["b"]) for p in panels] == [(0.0, 32768.0), (32768.0, 65536.0)]
    assert all(s["accepted"] for s in subs) and len(subs) == 2
    assert panels[-1]["hi_after"] == 0
    assert abs(panels[0]["d"] + 3.9633370367588094) < 1e-12   # min-norm least-squares quirk, adaptive.jl:210-214


def test_warping_lags(golden):
    """test/derivatives/warping.jl:21-23: kernel_values on warped lags ~ closed form at tol=1e-12
    (isapprox default rtol = sqrt(eps))."""
    cfg = so.OracleConfig(cf.exponential_sdf, tol=1e-12)
    assert cfg.quadspec == (4096, 16)                         # 1e-12 is not < 1e-12 (adaptive.jl:37)
    assert so.OracleConfig(cf.exponential_sdf, tol=1e-13).quadspec == (4096, 1)   # adaptive.jl:37-40
    lags = golden["warp_lags"][::4]
    vals, _ = so.kernel_values(cfg, lags)
    true = golden["warp_K"][::4]
    assert np.linalg.norm(vals - true) <= 1.5e-8 * np.linalg.norm(true)


def _matern_param_derivs(parms):
    phi, rho, nu = parms
    ex = -nu - 0.5
    base = lambda w: rho ** 2 + w ** 2
    return [lambda w: base(w) ** ex,
            lambda w: phi * ex * base(w) ** (ex - 1) * 2 * rho,
            lambda w: -phi * base(w) ** ex * np.log(base(w))]


def test_sdf_param_derivatives(golden):
    """test/derivatives/sdf_params.jl (an ENABLED reference test): dK/d(phi, rho, nu) by one adaptive run per
    parameter with f = dS/dtheta_j (src/derivatives.jl:63-72), tol = 1e-12, threshold 1e-5 as upstream."""
    parms = tuple(golden["sdfp_parms"])
    xs = golden["sdfp_r"]
    cfg = so.OracleConfig(lambda w: cf.matern_sdf(w, parms), tol=1e-12)
    k0 = so.compute_k0(cfg)
    assert abs(k0 - golden["sdfp_K"][0]) <= 1e-8 * k0
    for dS, key in zip(_matern_param_derivs(parms), ("sdfp_dphi", "sdfp_drho", "sdfp_dnu")):
        cfgj = so.gen_new_sdf_config(cfg, dS)                   # tol carried over, default quadspec (adaptive.jl:69-72)
        vals, _ = so.kernel_values(cfgj, xs, k0=k0, param_derivative=True)
        assert np.max(np.abs(vals - golden[key])) < 1e-5


def test_kernel_warping_gradients(golden):
    """test/derivatives/warping.jl:36-44: d kernel / d warp-params = K'(lag) * d lag / d params, with K' from
    the derivative config (sin kernel; src/derivatives.jl:51-59), threshold 1e-8 as upstream.  The warp
    (x/p1)^p2 is differentiated by hand here (ForwardDiff upstream)."""
    p1, p2 = 1 / 50.0, 1.1
    xs = np.linspace(1.1, 2.0, 100)[::5]
    wx, wy = (xs / p1) ** p2, (1.0 / p1) ** p2
    lags = np.abs(wy - wx)
    sgn = np.sign(wx - wy)
    dlag = np.stack([sgn * (-p2 / p1) * (wx - wy), sgn * (wx * np.log(xs / p1) - wy * np.log(1.0 / p1))], axis=1)
    cfg = so.OracleConfig(cf.exponential_sdf, tol=1e-12)
    k0 = so.compute_k0(cfg)
    dK, _ = so.kernel_values(so.gen_derivative_config(cfg), lags, k0=k0)
    grads = dK[:, None] * dlag
    true = cf.exponential_dcov(lags)[:, None] * dlag
    assert np.max(np.linalg.norm(grads - true, axis=1)) < 1e-8


def test_duplicates_unsorted_and_zero():
    """adaptive.jl:99-107 (unique + scatter), :113-120 (sort), :133-146 (r = 0 row)."""
    cfg = so.OracleConfig(cf.exponential_sdf)
    r = np.array([0.3, 0.0, 1.7, 0.3, 0.05, 1.7, 0.0])
    vals, errs = so.kernel_values(cfg, r)
    assert np.allclose(vals, cf.exponential_cov(r), atol=2e-8 * 2)
    assert vals[0] == vals[3] and vals[2] == vals[5]
    assert np.isnan(errs[1]) and np.isnan(errs[6])
    k0 = so.compute_k0(cfg)
    assert vals[1] == k0


def test_gauss_rules_moments():
    """QuadRule (quadrature.jl:27-47): exactness of the generated rules."""
    for n in (64, 4096, 8192):
        x, w = so.gauss_rule(n)
        assert abs(w.sum() - 2) < 4e-15 and abs((w * x ** 2).sum() - 2 / 3) < 4e-15
        assert np.all(np.diff(x) > 0)
    for p in (-0.5, 0.5, 1.5):
        for n in (64, 8192):
            x, w = so.gauss_rule(n, p)
            for j in (0, 1, 7, 40):
                exact = 2 ** (p + j + 1) / (p + j + 1)
                assert abs((w * (1 + x) ** j).sum() / exact - 1) < 1e-13
    # singular integrand the rule is built for: int_0^1 w^-1/2 cos(w) dw
    x, w = so.gauss_rule(64, -0.5)
    h = 0.5
    val = np.sum(w * h ** 0.5 * np.cos(h * x + h))
    from scipy import integrate
    ref = integrate.quad(lambda t: 2 * np.cos(t * t), 0, 1, epsabs=0, epsrel=1e-13)[0]
    assert abs(val - ref) < 1e-14


def test_cpu_nufft_matches_direct():
    """The CPU type-3 NUFFT restatement against the reference's own direct summation
    (quadrature.jl:113-128) on the default panel shapes."""
    rng = np.random.default_rng(0)
    cfg = so.OracleConfig(lambda w: (1 + w ** 2) ** -2)
    x = np.sort(rng.uniform(1e-4, 1.0, 300))
    no1, buf1, no2, buf2 = so.updatequadbufs(cfg, cfg.f, 0.0, 32768.0 / x[-1])
    for no, buf in ((no1, buf1), (no2, buf2)):
        d = so.direct_cis(no, buf, x)
        f = so.cpu_nufft1d3(no, buf, x)
        # plain-double positions: error ~ eps * (space-bandwidth product 2^15), as in any double NUFFT
        assert np.max(np.abs(d - f)) <= 3e-11 * np.sum(np.abs(buf))
    # complex strengths, later panel, clustered targets
    w = np.sort(rng.uniform(5000.0, 9000.0, 5000))
    s = rng.normal(size=5000) + 1j * rng.normal(size=5000)
    x = np.sort(np.concatenate([rng.uniform(0.2, 0.21, 50), rng.uniform(0, 1.0, 50)]))
    d = so.direct_cis(w, s, x)
    f = so.cpu_nufft1d3(w, s, x)
    assert np.max(np.abs(d - f)) <= 3e-11 * np.sum(np.abs(s))
