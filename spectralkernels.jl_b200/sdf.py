"""
Built-in spectral-density families.  Each is an ordinary callable S(w) on the host (needed by
compute_k0 and estimate_tail_decay, which stay on the host as in the reference) and also names the
device generator (sk_sdf_builtin) so the 196 608 integrand evaluations per sub-interval of
updatequadbufs! (src/quadrature.jl:49-95) never leave the GPU.  Any other Python callable works too:
it is evaluated on the host and its strengths are uploaded (sk_subinterval_host).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._capi import SK_SDF_EXPONENTIAL, SK_SDF_MATERN


@dataclass(frozen=True)
class Matern:
    """phi * (rho^2 + w^2)^(-nu - d/2)   (matern_sdf, scripts/matern_pair.jl:17).
    deriv = 0: S; 1, 2, 3: dS/dphi, dS/drho, dS/dnu (the integrands of src/derivatives.jl:63-72)."""
    phi: float = 1.0
    rho: float = 1.0
    nu: float = 1.5
    d: int = 1
    deriv: int = 0

    family = SK_SDF_MATERN

    @property
    def params(self):
        return (float(self.phi), float(self.rho), float(self.nu), float(self.d))

    def __call__(self, w):
        w = np.asarray(w, dtype=np.float64)
        base = self.rho ** 2 + w ** 2
        ex = -self.nu - self.d / 2
        if self.deriv == 0:
            return self.phi * base ** ex
        if self.deriv == 1:
            return base ** ex
        if self.deriv == 2:
            return self.phi * ex * base ** (ex - 1.0) * 2.0 * self.rho
        if self.deriv == 3:
            return -self.phi * base ** ex * np.log(base)
        raise ValueError("deriv must be 0..3")

    def derivative(self, j: int) -> "Matern":
        return Matern(self.phi, self.rho, self.nu, self.d, j)

    def dw(self, w):
        """dS/dw, the `df` keyword of AdaptiveKernelConfig (needed by logw=true, src/quadrature.jl:192)."""
        if self.deriv != 0:
            raise ValueError("dw is provided for the density itself only")
        w = np.asarray(w, dtype=np.float64)
        ex = -self.nu - self.d / 2
        return self.phi * ex * (self.rho ** 2 + w ** 2) ** (ex - 1.0) * 2.0 * w


@dataclass(frozen=True)
class Exponential:
    """phi * exp(-alpha |w|)   (test/exponential_sdf_1d.jl:3, test/derivatives/jacobian.jl:5)."""
    phi: float = 1.0
    alpha: float = 1.0
    deriv: int = 0

    family = SK_SDF_EXPONENTIAL

    @property
    def params(self):
        return (float(self.phi), float(self.alpha))

    def __call__(self, w):
        w = np.asarray(w, dtype=np.float64)
        e = np.exp(-self.alpha * np.abs(w))
        if self.deriv == 0:
            return self.phi * e
        if self.deriv == 1:
            return e
        if self.deriv == 2:
            return -self.phi * np.abs(w) * e
        raise ValueError("deriv must be 0..2")

    def derivative(self, j: int) -> "Exponential":
        return Exponential(self.phi, self.alpha, j)

    def dw(self, w):
        if self.deriv != 0:
            raise ValueError("dw is provided for the density itself only")
        w = np.asarray(w, dtype=np.float64)
        return -self.alpha * np.sign(w) * self.phi * np.exp(-self.alpha * np.abs(w))


def is_builtin(f) -> bool:
    return isinstance(f, (Matern, Exponential))
