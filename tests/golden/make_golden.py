"""
Generates tests/golden/closed_forms.npz -- the known-answer vectors the reference's own
tests hold for the K(r) path.  The reference cannot be executed in
this image (no Julia), so the fixtures are the closed forms its tests evaluate in-test:

  exp_*        test/exponential_sdf_1d.jl:3-9      r = range(0, 5.1, length=1000)
  matern_*     test/matern_sdf.jl:4-12 (dim=1)     parms (2.14, 0.97, 0.89), same r grid
  sing_*       test/matern_sdf.jl:38-46 (dim=1)    alpha = 0.5, r = range(0, 1.1, length=1000)
  readme_*     README.md:19-33                     r = 10 .^ range(-6, 0, length=1000)
  sdfp_*       test/derivatives/sdf_params.jl:6-18 parms (2.3, 0.1, 1.75), r = range(0, 3.5, length=30),
               d matern_cov / d(phi, rho, nu) (ForwardDiff in the reference, mpmath.diff here)
  warp_*       test/derivatives/warping.jl:5-22    warped lags and kernel values

Run:  python tests/golden/make_golden.py      (needs scipy + mpmath; a few seconds)
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import closed_forms as cf  # noqa: E402


def matern_cov_mp(t, phi, alpha, v, d=1):
    import mpmath as mp
    constant = mp.pi ** (mp.mpf(d) / 2) * phi / (2 ** (v - 1) * mp.gamma(v + mp.mpf(d) / 2) * alpha ** (2 * v))
    arg = alpha * 2 * mp.pi * abs(t)
    if arg == 0:
        return constant * 2 ** (v - 1) * mp.gamma(v)
    return constant * mp.besselk(v, arg) * arg ** v


def main():
    import mpmath as mp
    out = {}
    r51 = np.linspace(0.0, 5.1, 1000)
    out["exp_r"] = r51
    out["exp_K"] = cf.exponential_cov(r51)
    out["exp_dK"] = cf.exponential_dcov(r51)
    parms = (2.14, 0.97, 0.89)
    out["matern_parms"] = np.array(parms)
    out["matern_r"] = r51
    out["matern_K"] = cf.matern_cov(r51, parms, d=1)
    out["matern_dK"] = cf.matern_dcov(r51, parms, d=1)
    out["matern2d_K"] = cf.matern_cov(r51, parms, d=2)
    out["matern2d_dK"] = cf.matern_dcov(r51, parms, d=2)
    r11 = np.linspace(0.0, 1.1, 1000)
    out["sing_r"] = r11
    out["sing_alpha"] = np.array(0.5)
    out["sing_K"] = cf.sing_matern_cov(r11, (*parms, -0.5), d=1)
    # d/d alpha of the singular kernel (test/matern_sdf.jl:66-86; ForwardDiff there, mpmath.diff here),
    # on every 5th point of the grid (the points the parity tests use)
    import mpmath as mp_
    idx = np.unique(np.append(np.arange(0, 1000, 5), 999))[1:]
    dka = []
    with mp_.workdps(40):
        for t in r11[idx]:
            fun = lambda a_: mp_.mpf(float(cf.sing_matern_cov_mp(float(t), (*parms, -a_), d=1)))
            dka.append(float(mp_.diff(lambda a_: cf.sing_matern_cov_mp(float(t), (*parms, -a_), d=1), mp_.mpf("0.5"))))
    out["sing_dalpha_idx"] = idx
    out["sing_dalpha"] = np.array(dka)
    rr = 10 ** np.linspace(-6, 0, 1000)
    out["readme_r"] = rr
    out["readme_K"] = cf.readme_cov(rr)
    # sdf_params.jl
    p3 = (2.3, 0.1, 1.75)
    r35 = np.linspace(0.0, 3.5, 30)
    out["sdfp_parms"] = np.array(p3)
    out["sdfp_r"] = r35
    d1, d2, d3, k = [], [], [], []
    with mp.workdps(40):
        for t in r35:
            t = mp.mpf(float(t))
            k.append(float(matern_cov_mp(t, mp.mpf(p3[0]), mp.mpf(p3[1]), mp.mpf(p3[2]))))
            d1.append(float(mp.diff(lambda q: matern_cov_mp(t, q, mp.mpf(p3[1]), mp.mpf(p3[2])), mp.mpf(p3[0]))))
            d2.append(float(mp.diff(lambda q: matern_cov_mp(t, mp.mpf(p3[0]), q, mp.mpf(p3[2])), mp.mpf(p3[1]))))
            d3.append(float(mp.diff(lambda q: matern_cov_mp(t, mp.mpf(p3[0]), mp.mpf(p3[1]), q), mp.mpf(p3[2]))))
    out["sdfp_K"] = np.array(k)
    out["sdfp_dphi"] = np.array(d1)
    out["sdfp_drho"] = np.array(d2)
    out["sdfp_dnu"] = np.array(d3)
    # warping.jl: warp(params, x) = (x/params[1])^params[2], params = [1/50, 1.1], pairs (1.0, x)
    tp = (1 / 50.0, 1.1)
    xs = np.linspace(1.1, 2.0, 100)
    lags = np.abs((1.0 / tp[0]) ** tp[1] - (xs / tp[0]) ** tp[1])
    out["warp_lags"] = lags
    out["warp_K"] = cf.exponential_cov(lags)
    np.savez_compressed(os.path.join(HERE, "closed_forms.npz"), **out)
    print("wrote", os.path.join(HERE, "closed_forms.npz"), {k_: np.asarray(v).shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
