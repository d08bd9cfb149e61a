"""
Host mirror of the callers either side of the K(r) path: the model layer that turns points and
index pairs into lags and kernel values back into a lookup table, and the dense covariance helper --

    NoWarping, dense_index_pairs, SpectralModel, gen_kernel_setup, gen_kernel, SpectralKernel   src/model.jl:1-90
    gen_kernel_jacobian                                                                         src/derivatives.jl:86-112
    gen_kernel_dual (what ext/SpectralKernelsForwardDiffExt.jl:7-22 assembles from the Jacobian)
    build_dense_cov_matrix                                                                      src/utils.jl:41-64

with the same names and argument meaning.  What changes underneath: the lags of the index pairs are computed, sorted
and de-duplicated on the device (`sk_targets_set_pairs`), every derivative run of a Jacobian re-uses that sort
(`reuse_targets`), and the values come back as ONE flat array in pair order -- `SpectralKernel.store`, the
`Dict(zip(raw_pairs, values))` of src/model.jl:77, is a view built on first use, because at ~1e7 pairs the dictionary
itself is what dominates the reference's wall time.

Python has no ForwardDiff: the parameter derivatives of the spectral density come from the built-in families' device
generators (or `dsdfs`), and the gradient of the warped lag w.r.t. the warping parameters is either supplied
(`warp_grad`) or taken by central differences of the warping function (cheap: O(pairs) scalar work, no integration).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from .adaptive import AdaptiveKernelConfig, compute_k0, gen_new_sdf_config, kernel_values
from .derivatives import kernel_derivative, kernel_sdf_derivatives, kernel_singularity_derivative
from .sdf import is_builtin


class NoWarping:                                                              # src/model.jl:1-3
    def __call__(self, params, x):
        return x


def dense_index_pairs(pts) -> np.ndarray:
    """All pairs (j, k) with j <= k -- diagonal included -- in the column-major order of
    `vec(collect(Iterators.product(eachindex(pts), eachindex(pts))))` (src/model.jl:15-19); 0-based here."""
    n = len(pts)
    k, j = np.triu_indices(n)                 # row-major over (k, j >= k) ...
    order = np.lexsort((k, j))                # ... re-ordered: second index slowest, first index fastest
    return np.stack([k[order], j[order]], axis=1).astype(np.int64)


class SpectralModel:
    """`SpectralModel(sdf, pts; warp, kernel_index_pairs, sdf_param_indices, warp_param_indices,
    singularity_param_index, verbose, kwargs...)` (src/model.jl:39-47).

    `sdf(sdf_params)` returns the spectral density for those parameters -- a built-in family (`Matern`, `Exponential`:
    evaluated on the device) or any callable S(w) -- the role of `ParametricFunction(cfg.f, params)` (src/model.jl:56).
    Parameter indices are 0-based.  `kwargs` go to `AdaptiveKernelConfig` (tol, alpha, ...); `dim` is the dimension of
    the points, as in the reference."""

    def __init__(self, sdf: Callable, pts, *, warp=None, kernel_index_pairs=None, sdf_param_indices,
                 warp_param_indices=(), singularity_param_index: Optional[int] = None, verbose: bool = False,
                 warp_grad: Optional[Callable] = None, dsdfs: Optional[Callable] = None, df: Optional[Callable] = None,
                 **kwargs):
        self.sdf = sdf
        self.pts = np.asarray(pts, dtype=np.float64)
        if self.pts.ndim == 1:
            self.pts = self.pts.reshape(-1, 1)
        self.warp = warp if warp is not None else NoWarping()
        self.kernel_index_pairs = (dense_index_pairs(self.pts) if kernel_index_pairs is None
                                   else np.ascontiguousarray(kernel_index_pairs, dtype=np.int64).reshape(-1, 2))
        as_tuple = lambda v: (int(v),) if np.isscalar(v) else tuple(int(i) for i in v)
        self.sdf_param_indices = as_tuple(sdf_param_indices)
        self.warp_param_indices = as_tuple(warp_param_indices)
        self.singularity_param_index = singularity_param_index
        self.verbose = verbose
        self.warp_grad, self.dsdfs, self.df = warp_grad, dsdfs, df
        self.cfg_kwargs = dict(kwargs)
        self.cfg_kwargs.setdefault("dim", self.pts.shape[1])                   # src/model.jl:43
        self._engine = None
        self._cache_key = None                                                  # pair list resident on the device?

    # one engine per model: all kernel evaluations of a fit share the device scratch (and the sorted pair lags)
    def _config(self, S, alpha) -> AdaptiveKernelConfig:
        kw = dict(self.cfg_kwargs)
        kw["alpha"] = alpha
        cfg = AdaptiveKernelConfig(S, engine=self._engine, **kw)
        self._engine = cfg.engine
        return cfg


def _warp_points(sm: SpectralModel, warp_params) -> np.ndarray:
    out = np.stack([np.atleast_1d(np.asarray(sm.warp(warp_params, p), dtype=np.float64)) for p in
                    (sm.pts[:, 0] if sm.pts.shape[1] == 1 else sm.pts)])
    return out.reshape(sm.pts.shape[0], -1)


def gen_kernel_setup(sm: SpectralModel, params):
    """src/model.jl:53-68.  Returns (new_cfg, warp_pts, raw_pairs, warp_params): the lags themselves are formed on the
    device from `warp_pts` and `sm.kernel_index_pairs` (norm(warp_pts[j] - warp_pts[k]), :66)."""
    params = np.asarray(params, dtype=np.float64)
    sdf_params = tuple(params[j] for j in sm.sdf_param_indices)
    alpha = 0.0 if sm.singularity_param_index is None else float(params[sm.singularity_param_index])
    new_cfg = sm._config(sm.sdf(sdf_params), alpha)
    warp_params = tuple(params[j] for j in sm.warp_param_indices)
    warp_pts = _warp_points(sm, warp_params)
    return new_cfg, warp_pts, sm.kernel_index_pairs, warp_params


class SpectralKernel:
    """src/model.jl:49-51, :79-90: the covariance lookup `kernel(x, y)` over the point pairs of the model.  Holds the
    flat value array in pair order; `store` (the reference's Dict keyed by raw point pairs) is built on first use."""

    def __init__(self, pts: np.ndarray, pairs: np.ndarray, values: np.ndarray):
        self.pts, self.pairs, self.values = pts, pairs, values
        self._store = None
        self._pos = None

    @staticmethod
    def _key(p):
        return tuple(np.atleast_1d(np.asarray(p, dtype=np.float64)).tolist())

    @property
    def store(self) -> dict:
        if self._store is None:
            keys = [self._key(p) for p in self.pts]
            self._store = {(keys[j], keys[k]): v for (j, k), v in zip(self.pairs.tolist(), self.values.tolist())}
        return self._store

    def __call__(self, x, y, _params=None):                                    # the third argument is swallowed, :90
        kx, ky = self._key(x), self._key(y)
        st = self.store
        if (kx, ky) in st:
            return st[(kx, ky)]
        if (ky, kx) in st:
            return st[(ky, kx)]
        raise KeyError(f"Point pair ({x}, {y}) not in the `SpectralKernel` lookup table.")

    def matrix(self) -> np.ndarray:
        """dense symmetric matrix over the model's points (NaN where a pair is not in the table)"""
        n = self.pts.shape[0]
        M = np.full((n, n), np.nan)
        M[self.pairs[:, 0], self.pairs[:, 1]] = self.values
        M[self.pairs[:, 1], self.pairs[:, 0]] = self.values
        return M


def _values(sm, cfg, warp_pts, pairs, **kw):
    return kernel_values(cfg, None, points=warp_pts, pairs=pairs, verbose=sm.verbose, **kw)


def gen_kernel(sm: SpectralModel, params, *, k0: Optional[float] = None) -> SpectralKernel:
    """src/model.jl:73-77."""
    cfg, warp_pts, pairs, _ = gen_kernel_setup(sm, params)
    vals, _ = _values(sm, cfg, warp_pts, pairs, k0=k0)
    return SpectralKernel(sm.pts, pairs, vals)


def _lag_gradients(sm: SpectralModel, warp_params, pairs) -> np.ndarray:
    """d || warp(theta, x_j) - warp(theta, x_k) || / d theta for every pair: [npairs, len(theta)]
    (warping_gradients, src/derivatives.jl:33-45, without the multipliers)."""
    theta = np.asarray(warp_params, dtype=np.float64)
    if theta.size == 0:
        return np.zeros((pairs.shape[0], 0))
    if sm.warp_grad is not None:
        return np.asarray(sm.warp_grad(theta, sm.pts, pairs), dtype=np.float64).reshape(pairs.shape[0], theta.size)

    def lags(th):
        wp = _warp_points(sm, tuple(th))
        return np.linalg.norm(wp[pairs[:, 0]] - wp[pairs[:, 1]], axis=1)

    out = np.empty((pairs.shape[0], theta.size))
    for m in range(theta.size):                        # central differences, Richardson-extrapolated: O(h^4)
        h = 1e-3 * max(1.0, abs(theta[m]))
        e = np.zeros_like(theta)
        e[m] = h
        d1 = (lags(theta + e) - lags(theta - e)) / (2 * h)
        d2 = (lags(theta + e / 2) - lags(theta - e / 2)) / h
        out[:, m] = (4 * d2 - d1) / 3
    return out


def gen_kernel_jacobian(sm: SpectralModel, params, k0: float) -> np.ndarray:
    """src/derivatives.jl:86-112: d K(pair) / d params, [npairs, nparams], columns in parameter order.  One adaptive
    run per spectral-density parameter, one derivative-config run for all warping parameters (chain rule through the
    warped lag), one log-weighted run for the singularity parameter; the pair lags are sorted once."""
    cfg, warp_pts, pairs, warp_params = gen_kernel_setup(sm, params)
    sdf_params = tuple(np.asarray(params, dtype=np.float64)[j] for j in sm.sdf_param_indices)
    eng = cfg.engine
    eng.targets_set_pairs(warp_pts, pairs)             # the sort every run below re-uses
    dsdfs = sm.dsdfs(sdf_params) if sm.dsdfs is not None else None
    if dsdfs is None and is_builtin(cfg.f):
        # the built-in families list their parameters in constructor order; the model's sdf parameters are a prefix
        dsdfs = [cfg.f.derivative(j) for j in range(1, len(sdf_params) + 1)]
    cols, order = [], []
    for j, d in zip(sm.sdf_param_indices, kernel_sdf_derivatives(cfg, None, k0, dsdfs=dsdfs, reuse_targets=True,
                                                                  points=warp_pts, pairs=pairs)):
        cols.append(d)
        order.append(j)
    if sm.warp_param_indices:
        dK = kernel_derivative(cfg, None, k0, reuse_targets=True, points=warp_pts, pairs=pairs)     # K'(lag)
        g = _lag_gradients(sm, warp_params, pairs)
        for m, j in enumerate(sm.warp_param_indices):
            cols.append(dK * g[:, m])
            order.append(j)
    if sm.singularity_param_index is not None:
        df = sm.df(sdf_params) if sm.df is not None else getattr(cfg.f, "dw", None)
        cols.append(kernel_singularity_derivative(cfg, None, k0, df, reuse_targets=True, points=warp_pts, pairs=pairs))
        order.append(sm.singularity_param_index)
    J = np.empty((pairs.shape[0], len(np.asarray(params))))
    J[:] = 0.0
    for j, c in zip(order, cols):
        J[:, j] = c
    return J


def gen_kernel_dual(sm: SpectralModel, params, partials):
    """What ext/SpectralKernelsForwardDiffExt.jl:7-22 builds for dual-number parameters: the primal kernel and, per
    pair, the partials sum_m J[pair, m] * partials[m, :].  `partials`: [nparams, N].  Returns (SpectralKernel, [npairs, N])."""
    out = gen_kernel(sm, params)
    k0 = out(sm.pts[0], sm.pts[0])                     # :11 -- the (pts[1], pts[1]) entry must be in the pair list
    J = gen_kernel_jacobian(sm, params, k0)
    return out, J @ np.asarray(partials, dtype=np.float64)


def build_dense_cov_matrix(cfg: AdaptiveKernelConfig, pts) -> np.ndarray:
    """src/utils.jl:41-64 (1-D points): the dense covariance matrix from ONE kernel_values call over the zero lag and
    all pairwise distances |pts[i] - pts[j]|, j > i.  The reference sorts the lags itself (sortperm) before the call;
    here the device sorts, and its default pair order -- strict upper triangle, row-major -- is the order of :44-45."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1)
    npt = pts.size
    vals, _ = kernel_values(cfg, None, points=pts.reshape(-1, 1))            # all i < j, row-major
    k00, _ = kernel_values(cfg, np.zeros(1))
    M = np.empty((npt, npt))
    iu = np.triu_indices(npt, k=1)
    M[iu] = vals
    M[(iu[1], iu[0])] = vals
    M[np.diag_indices(npt)] = k00[0]
    return M
