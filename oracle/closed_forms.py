"""
closed_forms.py -- TEST INFRASTRUCTURE ONLY.

The closed-form known answers the reference's own tests compare `kernel_values`
against.  These are the golden vectors of the path:

  exponential  test/exponential_sdf_1d.jl:3-5     S = exp(-|w|)
  scaled exp   test/derivatives/jacobian.jl:5-6   S = exp(-|w|/alpha)... (2a/(a^2+(2 pi r)^2))
  Matern       scripts/matern_pair.jl:7-17        needs K_nu  (scipy.special.kv)
  singular     scripts/matern_pair.jl:20-33       needs 1F2   (mpmath.hyp1f2)
  README demo  README.md:19-33  S = (1+w^2)^-2 == Matern(phi=1, rho=1, nu=3/2, d=1)
"""
from __future__ import annotations

import math

import numpy as np


# --- exponential: test/exponential_sdf_1d.jl:3-5 ---------------------------------------
def exponential_sdf(w):
    return np.exp(-np.abs(w))


def exponential_cov(r):
    r = np.asarray(r, dtype=float)
    return 2 / (1 + (2 * math.pi * r) ** 2)


def exponential_dcov(r):
    r = np.asarray(r, dtype=float)
    return -(16 * math.pi ** 2 * r) / (1 + (2 * math.pi * r) ** 2) ** 2


# --- Matern: scripts/matern_pair.jl:7-17 -----------------------------------------------
def matern_sdf(w, parms, d=1):
    phi, rho, nu = parms
    return phi * (rho ** 2 + np.asarray(w, dtype=float) ** 2) ** (-nu - d / 2)


def matern_cov(t, parms, d=1):
    from scipy import special
    phi, alpha, v = parms
    t = np.atleast_1d(np.asarray(t, dtype=float))
    constant = math.pi ** (d / 2) * phi / (2 ** (v - 1) * special.gamma(v + d / 2) * alpha ** (2 * v))
    arg = alpha * 2 * math.pi * np.abs(t)
    out = np.empty_like(arg)
    z = arg == 0
    out[z] = constant * 2 ** (v - 1) * special.gamma(v)      # lim x^v K_v(x), x -> 0
    nz = ~z
    out[nz] = constant * special.kv(v, arg[nz]) * arg[nz] ** v
    return out


def matern_dcov(t, parms, d=1):
    """d/dt of matern_cov (the reference uses ForwardDiff of matern_cov, test/matern_sdf.jl:10).
    d/dx [x^v K_v(x)] = -x^v K_{v-1}(x)."""
    from scipy import special
    phi, alpha, v = parms
    t = np.atleast_1d(np.asarray(t, dtype=float))
    constant = math.pi ** (d / 2) * phi / (2 ** (v - 1) * special.gamma(v + d / 2) * alpha ** (2 * v))
    arg = alpha * 2 * math.pi * np.abs(t)
    out = np.zeros_like(arg)
    nz = arg != 0
    out[nz] = -constant * arg[nz] ** v * special.kv(v - 1, arg[nz]) * alpha * 2 * math.pi
    return out


def readme_cov(r):
    """S = (1+w^2)^-2  =>  K(r) = (pi/2)(1 + 2 pi r) exp(-2 pi r)   (README.md:19-33)."""
    r = np.asarray(r, dtype=float)
    return (math.pi / 2) * (1 + 2 * math.pi * r) * np.exp(-2 * math.pi * r)


# --- singular Matern: scripts/matern_pair.jl:20-33 --------------------------------------
def sing_matern_cov_mp(t, params, d=1):
    """scalar mpmath version of sing_matern_cov (params may hold mpf values; used for d/d alpha)."""
    import mpmath as mp
    phi, a, b, p = [mp.mpf(x) for x in params]
    dd = mp.mpf(d)
    tt = mp.mpf(t) + mp.mpf("1e-30")
    z = a ** 2 * mp.pi ** 2 * tt ** 2
    o = mp.pi ** p * (a * tt) ** p * mp.gamma((dd + p) / 2) * \
        mp.hyp1f2((dd + p) / 2, dd / 2, (2 - 2 * b + p) / 2, z) / (mp.gamma(dd / 2) * mp.gamma((2 - 2 * b + p) / 2))
    o -= mp.pi ** (2 * b) * (a * tt) ** (2 * b) * mp.gamma(b + dd / 2) * \
        mp.hyp1f2(b + dd / 2, 1 + b - p / 2, b + dd / 2 - p / 2, z) / \
        (mp.gamma(1 + b - p / 2) * mp.gamma(b + dd / 2 - p / 2))
    o *= phi * a ** (-2 * b) * mp.pi ** (1 + dd / 2 - p) * tt ** (-p) * mp.csc(b * mp.pi - p * mp.pi / 2) / \
        mp.gamma(b + dd / 2)
    return o


def sing_matern_cov(t, params, d=1, dps=40):
    """params = (phi, a, b, p) with p = -alpha (matern_pair.jl:33); evaluated in mpmath."""
    import mpmath as mp
    phi, a, b, p = params
    t = np.atleast_1d(np.asarray(t, dtype=float))
    out = np.empty(t.size)
    with mp.workdps(dps):
        phi, a, b, p, dd = mp.mpf(phi), mp.mpf(a), mp.mpf(b), mp.mpf(p), mp.mpf(d)
        for i, ti in enumerate(t):
            tt = mp.mpf(float(ti)) + mp.mpf("1e-30")           # matern_pair.jl:33
            z = a ** 2 * mp.pi ** 2 * tt ** 2
            o = mp.pi ** p * (a * tt) ** p * mp.gamma((dd + p) / 2) * \
                mp.hyp1f2((dd + p) / 2, dd / 2, (2 - 2 * b + p) / 2, z) / \
                (mp.gamma(dd / 2) * mp.gamma((2 - 2 * b + p) / 2))
            o -= mp.pi ** (2 * b) * (a * tt) ** (2 * b) * mp.gamma(b + dd / 2) * \
                mp.hyp1f2(b + dd / 2, 1 + b - p / 2, b + dd / 2 - p / 2, z) / \
                (mp.gamma(1 + b - p / 2) * mp.gamma(b + dd / 2 - p / 2))
            o *= phi * a ** (-2 * b) * mp.pi ** (1 + dd / 2 - p) * tt ** (-p) * mp.csc(b * mp.pi - p * mp.pi / 2) / \
                mp.gamma(b + dd / 2)
            out[i] = float(o)
    return out
