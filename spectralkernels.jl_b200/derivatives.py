"""
Host mirror of the kernel-evaluation parts of src/derivatives.jl: every derivative of K is one more
`kernel_values` run over the SAME lags with a different integrand or kernel, so the lags are uploaded,
sorted and de-duplicated once (`reuse_targets`) and each run only repeats the panel loop.

    kernel_derivative(cfg, lags, k0)          K'(lag)                  src/derivatives.jl:51-59 (the
                                              gen_derivative_config run of kernel_warping_gradients)
    kernel_sdf_derivatives(cfg, lags, k0)     dK/d theta_j             src/derivatives.jl:63-72
    kernel_singularity_derivative(cfg, ...)   dK/d alpha               src/derivatives.jl:74-81

The chain rule through the warping function and the assembly of the Jacobian (src/derivatives.jl:33-45,
:86-112) are host-language autodiff plumbing and stay with the host.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .adaptive import AdaptiveKernelConfig, gen_derivative_config, gen_new_sdf_config, kernel_values
from .sdf import is_builtin


def kernel_derivative(cfg: AdaptiveKernelConfig, lags, k0: float, *, reuse_targets: bool = False, **kw):
    """K'(lag): the sin-kernel run with p + 1 and c * (-2 pi)  (src/adaptive.jl:61-66, src/quadrature.jl:177).
    Only the values are used by the callers (`[1]` in src/derivatives.jl:56), so the error estimates stay on the
    device unless `want_errors=True` is passed; `out_vals` (pinned host memory) avoids the pageable-memory copy."""
    dcfg = gen_derivative_config(cfg)
    kw.setdefault("want_errors", False)
    return kernel_values(dcfg, lags, k0=k0, reuse_targets=reuse_targets, **kw)[0]


def kernel_sdf_derivatives(cfg: AdaptiveKernelConfig, lags, k0: float, *, dsdfs: Optional[List] = None,
                           reuse_targets: bool = False, outs: Optional[List[np.ndarray]] = None, **kw) -> List[np.ndarray]:
    """One adaptive run per spectral-density parameter with f = dS/d theta_j (src/derivatives.jl:63-72).
    For the built-in families the parameter derivatives are device generators (`S.derivative(j)`); for
    other callables pass `dsdfs`, a list of callables."""
    if dsdfs is None:
        if not is_builtin(cfg.f):
            raise TypeError("pass dsdfs=[dS/dtheta_1, ...] for a spectral density that is not a built-in family")
        nparam = {1: 3, 2: 2}[cfg.f.family]
        dsdfs = [cfg.f.derivative(j) for j in range(1, nparam + 1)]
    out = []
    kw.setdefault("want_errors", False)                                      # only `[1]` is used (src/derivatives.jl:70)
    for j, dS in enumerate(dsdfs):
        cfgj = gen_new_sdf_config(cfg, dS)                                   # drops quadspec etc., as the reference
        okw = dict(kw)
        if outs is not None:
            okw["out_vals"] = outs[j]                                        # e.g. pinned host arrays, one per parameter
        out.append(kernel_values(cfgj, lags, k0=k0, param_derivative=True,
                                 reuse_targets=reuse_targets or j > 0, **okw)[0])
    return out


def kernel_singularity_derivative(cfg: AdaptiveKernelConfig, lags, k0: float, df, *, reuse_targets: bool = False, **kw):
    """dK/d alpha through the log-weighted config (src/derivatives.jl:74-81)."""
    acfg = AdaptiveKernelConfig(cfg.f, df=df, derivative=False, alpha=cfg.alpha, dim=cfg.dim, logw=True, tol=cfg.tol,
                                device=cfg.device, nufft_eps=cfg.nufft_eps)
    acfg._engine = cfg._engine
    kw.setdefault("want_errors", False)
    return kernel_values(acfg, lags, k0=k0, param_derivative=True, reuse_targets=reuse_targets, **kw)[0]
