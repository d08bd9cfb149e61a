"""
The callers either side of the path (-m gpu): SpectralModel / gen_kernel / SpectralKernel (src/model.jl),
gen_kernel_jacobian (src/derivatives.jl:86-112), the dual-number assembly of ext/SpectralKernelsForwardDiffExt.jl and
build_dense_cov_matrix (src/utils.jl:41-64), replaying the reference's own tests with their thresholds:

    test/derivatives/jacobian.jl:33      max |J_test - J_ref| < 1e-8   (warped 1-D exponential model)
    test/derivatives/warping.jl:21-23    kernel values through a warped model
"""
import numpy as np
import pytest

import closed_forms as cf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sk():
    import spectralkernels_jl_b200 as sk
    return sk


def _jacobian_model(sk):
    # test/derivatives/jacobian.jl:4-19 -- iso_sdf(w, alpha) = exp(-alpha |w|), warp(params, x) = x^params[1],
    # test_params = [1.1, 0.1]: parameter 0 is the density's alpha, parameter 1 the warping exponent
    warp = lambda params, x: x ** params[0]
    xgrid = np.linspace(0.0, 1.0, 20)
    model = sk.SpectralModel(lambda p: sk.Exponential(1.0, p[0]), xgrid, warp=warp, sdf_param_indices=0,
                             warp_param_indices=1, tol=1e-12,
                             dsdfs=lambda p: [sk.Exponential(1.0, p[0]).derivative(2)])       # dS/dalpha
    return model, xgrid


def test_jacobian_jl(sk):
    model, xgrid = _jacobian_model(sk)
    params = np.array([1.1, 0.1])
    al, p = params
    gk = sk.gen_kernel(model, params)
    k0 = gk(xgrid[0], xgrid[0])
    assert abs(k0 - 2 / al) <= 1e-11
    J = sk.gen_kernel_jacobian(model, params, k0)
    pairs = model.kernel_index_pairs
    x, y = xgrid[pairs[:, 0]], xgrid[pairs[:, 1]]
    with np.errstate(divide="ignore", invalid="ignore"):
        wx, wy = x ** p, y ** p
        lx = np.where(x > 0, wx * np.log(np.where(x > 0, x, 1.0)), 0.0)
        ly = np.where(y > 0, wy * np.log(np.where(y > 0, y, 1.0)), 0.0)
    r = np.abs(wx - wy)
    q = al ** 2 + (2 * np.pi * r) ** 2
    J_ref = np.stack([2 / q - 4 * al ** 2 / q ** 2,                                   # d/dalpha 2a/(a^2 + (2 pi r)^2)
                      (-2 * al * 8 * np.pi ** 2 * r / q ** 2) * np.sign(wx - wy) * (lx - ly)], axis=1)
    assert np.max(np.abs(gk.values - 2 * al / q)) <= 1e-10
    assert np.max(np.abs(J - J_ref)) < 1e-8                                           # jacobian.jl:33
    # the lookup of src/model.jl:79-90 in both argument orders, and the full (x, y) product of jacobian.jl:23
    full = np.array([[gk(a, b, params) for a in xgrid] for b in xgrid])
    assert np.allclose(full, full.T) and np.allclose(full, gk.matrix())
    with pytest.raises(KeyError):
        gk(0.123, 0.5)
    # dual numbers (ext/SpectralKernelsForwardDiffExt.jl): identity partials give the Jacobian back
    out, dv = sk.gen_kernel_dual(model, params, np.eye(2))
    assert np.array_equal(out.values, gk.values) and np.max(np.abs(dv - J_ref)) < 1e-8


def test_dense_pairs_warped_kernel_and_dense_matrix(sk):
    pts = np.linspace(1.1, 2.0, 40)
    tp = (1 / 50.0, 1.1)                                                              # test/derivatives/warping.jl:5-8
    warp = lambda params, x: (x / params[0]) ** params[1]
    model = sk.SpectralModel(lambda p: sk.Exponential(1.0, 1.0), pts, warp=warp, sdf_param_indices=(),
                             warp_param_indices=(0, 1), tol=1e-12)
    gk = sk.gen_kernel(model, np.array(tp))
    pr = model.kernel_index_pairs
    assert pr.shape[0] == 40 * 41 // 2 and np.all(pr[:, 0] <= pr[:, 1])
    lag = np.abs(warp(tp, pts[pr[:, 0]]) - warp(tp, pts[pr[:, 1]]))
    assert np.linalg.norm(gk.values - cf.exponential_cov(lag)) <= 1.5e-8 * np.linalg.norm(cf.exponential_cov(lag))
    # an O(n) pair list (what Vecchia's tile_pairs hands to the model, ext/SpectralKernelsVecchiaExt.jl:11-16)
    band = np.array([(i, j) for i in range(40) for j in range(i, min(i + 3, 40))])
    mb = sk.SpectralModel(lambda p: sk.Matern(p[0], p[1], p[2]), pts, kernel_index_pairs=band,
                          sdf_param_indices=(0, 1, 2), tol=1e-10)
    parms = (2.14, 0.97, 0.89)
    kb = sk.gen_kernel(mb, np.array(parms))
    assert np.max(np.abs(kb.values - cf.matern_cov(np.abs(pts[band[:, 0]] - pts[band[:, 1]]), parms))) <= 1e-8 * kb.values.max()
    Jb = sk.gen_kernel_jacobian(mb, np.array(parms), kb(pts[0], pts[0]))
    assert Jb.shape == (band.shape[0], 3) and np.max(np.abs(Jb[:, 0] - kb.values / parms[0])) <= 1e-8   # K linear in phi
    # src/utils.jl:41-64
    x1 = np.sort(np.random.default_rng(5).uniform(0, 3, 60))
    M = sk.build_dense_cov_matrix(sk.AdaptiveKernelConfig(sk.Exponential(1.0, 1.0)), x1)
    assert np.max(np.abs(M - cf.exponential_cov(np.abs(x1[:, None] - x1[None, :])))) <= 1e-7
