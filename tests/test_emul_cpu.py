"""
CPU check of the product's per-element arithmetic (csrc/sk_math.h, sk_host_util.h, sk_plan_host.cpp):
tests/emul/emul.cpp compiles those headers with g++ and runs them in plain loops; numpy supplies the
FFT.  The results are compared with the oracle's direct sums (src/quadrature.jl:113-128).  This is a
test harness, not a product path.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import sk_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = ctypes.POINTER(ctypes.c_double)


class EsPlan(ctypes.Structure):
    _fields_ = [("w", ctypes.c_int32), ("nq", ctypes.c_int32), ("beta", ctypes.c_double), ("ximax", ctypes.c_double),
                ("E", ctypes.c_double * 64), ("O", ctypes.c_double * 64), ("qc", ctypes.c_double * 24)]


class Geom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("wc", "D", "inv_hu", "kap_hi", "kap_lo", "t_cell")] + \
               [("nf", ctypes.c_longlong), ("nf2", ctypes.c_longlong)]


def _ptr(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(ROOT, "tests", "emul", "emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "libsk_emul.so")
    csrc = os.path.join(ROOT, "spectralkernels.jl_b200", "csrc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-fPIC", "-fopenmp", "-std=gnu++17", "-fext-numeric-literals",
                           "-ffp-contract=off", "-I", csrc, "-shared", "-o", out, src,
                           os.path.join(csrc, "sk_plan_host.cpp"), "-lquadmath", "-lm"])
    L = ctypes.CDLL(out)
    L.emul_geom.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4 + [ctypes.c_void_p]
    L.emul_spread.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, dp, dp, dp]
    L.emul_interp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, dp, dp, dp]
    P = EsPlan()
    assert L.emul_es_plan(16, ctypes.byref(P)) == 0
    return L, P


def _nufft(emul, no, c, x):
    L, P = emul
    G = Geom()
    assert L.emul_geom(ctypes.byref(P), no.min(), no.max(), x.min(), x.max(), ctypes.byref(G)) == 0
    cc = np.ascontiguousarray(np.asarray(c, dtype=complex))
    fin = np.zeros(G.nf2, dtype=complex)
    L.emul_spread(ctypes.byref(P), ctypes.byref(G), no.size, _ptr(no), _ptr(cc.view(float)), _ptr(fin.view(float)))
    g = np.ascontiguousarray(np.fft.ifft(fin) * G.nf2)
    out = np.zeros(x.size, dtype=complex)
    L.emul_interp(ctypes.byref(P), ctypes.byref(G), x.size, _ptr(x), _ptr(g.view(float)), _ptr(out.view(float)))
    return out, G


def test_type3_math_default_panels(emul):
    rng = np.random.default_rng(0)
    cfg = so.OracleConfig(lambda w: (1 + w ** 2) ** -2)
    x = np.sort(rng.uniform(1e-4, 1.0, 120))
    for (a, b) in ((0.0, 32768.0 / x[-1]), (32768.0 / x[-1], 65536.0 / x[-1])):
        no1, buf1, no2, buf2 = so.updatequadbufs(cfg, cfg.f, a, b)
        for no, buf in ((no1, buf1), (no2, buf2)):
            f, G = _nufft(emul, no, buf, x)
            assert G.D == 0.0 and G.nf2 == 262144             # 2^18 (power of two preferred)
            assert np.max(np.abs(f - so.direct_cis(no, buf, x))) <= 5e-13 * np.sum(np.abs(buf))


def test_type3_math_general_geometry(emul):
    rng = np.random.default_rng(1)
    w = np.sort(rng.uniform(5000.0, 9000.0, 3000))
    s = rng.normal(size=3000) + 1j * rng.normal(size=3000)
    cases = [np.sort(np.concatenate([rng.uniform(0.2, 0.21, 40), rng.uniform(0, 1.0, 40)])),
             np.sort(rng.uniform(0.7, 0.9, 60)),            # centred targets (D != 0, pre-phase)
             np.sort(rng.uniform(-0.5, 0.9, 60)),           # negative distances
             np.array([0.3, 0.3000001, 0.3000002])]         # degenerate range
    for x in cases:
        f, G = _nufft(emul, w, s, x)
        assert np.max(np.abs(f - so.direct_cis(w, s, x))) <= 5e-13 * np.sum(np.abs(s))
    f, _ = _nufft(emul, np.array([7.0]), np.array([1.0 + 2j]), np.array([0.25, 0.5]))
    assert np.max(np.abs(f - so.direct_cis(np.array([7.0]), np.array([1.0 + 2j]), np.array([0.25, 0.5])))) < 1e-13


def test_cell_polynomial_interpolation(emul):
    """k_interp_cells' block algorithm (cell polynomials in place of per-target taps) against the
    per-target evaluation and against direct sums, for dense, clustered and sparse sorted targets."""
    L, P = emul
    L.emul_interp_cells.restype = ctypes.c_longlong
    L.emul_interp_cells.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, dp, dp, ctypes.c_int,
                                    ctypes.c_int, dp]
    L.emul_interp2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, dp, dp, dp]
    rng = np.random.default_rng(2)
    cfg = so.OracleConfig(lambda w: (1 + w ** 2) ** -2)
    no1, buf1, no2, buf2 = so.updatequadbufs(cfg, cfg.f, 0.0, 32768.0)
    for name, x in (("dense", np.sort(rng.uniform(0.3, 0.3005, 4000))),
                    ("mixed", np.sort(np.concatenate([rng.uniform(0, 1, 1500), rng.uniform(0.6, 0.6002, 2500)]))),
                    ("sparse", np.sort(rng.uniform(1e-4, 1.0, 700)))):
        x[-1] = 1.0
        G = Geom()
        assert L.emul_geom(ctypes.byref(P), 0.0, 32768.0, x.min(), x.max(), ctypes.byref(G)) == 0
        grid2 = np.zeros((G.nf2, 2), dtype=complex)
        for rule, (no, buf) in enumerate(((no1, buf1), (no2, buf2))):
            fin = np.zeros(G.nf2, dtype=complex)
            cc = np.ascontiguousarray(buf.astype(complex))
            L.emul_spread(ctypes.byref(P), ctypes.byref(G), no.size, _ptr(no), _ptr(cc.view(float)), _ptr(fin.view(float)))
            grid2[:, rule] = np.fft.ifft(fin) * G.nf2
        grid2 = np.ascontiguousarray(grid2)
        ref = np.zeros((x.size, 2), dtype=complex)
        L.emul_interp2(ctypes.byref(P), ctypes.byref(G), x.size, _ptr(x), _ptr(grid2.view(float)), _ptr(ref.view(float)))
        out = np.zeros((x.size, 2), dtype=complex)
        npoly = L.emul_interp_cells(ctypes.byref(P), ctypes.byref(G), x.size, _ptr(x), _ptr(grid2.view(float)), 1024, 32,
                                    _ptr(out.view(float)))
        nblocks = -(-x.size // 1024)
        if name == "dense":
            assert npoly >= nblocks - 1      # the last block also holds x = 1.0
        if name == "sparse":
            assert npoly == 0
        scale = np.sum(np.abs(buf2))
        assert np.max(np.abs(out - ref)) <= 4e-15 * scale, name
        sub = slice(None, None, 40)
        assert np.max(np.abs(out[sub, 1] - so.direct_cis(no2, buf2, x[sub]))) <= 5e-13 * scale


def test_device_source_generator_matches_updatequadbufs(emul):
    """sk_gen_source (the K1 body) against the oracle's updatequadbufs (quadrature.jl:49-95)."""
    L, _ = emul
    m, k = 256, 4
    parms = np.array([2.14, 0.97, 0.89, 1.0])
    S = lambda w: parms[0] * (parms[1] ** 2 + w ** 2) ** (-parms[2] - 0.5)
    for alpha, a, b in ((0.0, 0.0, 800.0), (0.5, 0.0, 800.0), (0.5, 800.0, 1700.0)):
        cfg = so.OracleConfig(S, alpha=alpha, quadspec=(m, k))
        p = cfg.p
        origin = (a == 0.0 and p != 0.0)
        if origin:
            ref = so.updatequadbufs(cfg, S, a, b, p=p)
        else:
            g = lambda w: so._pow(w, p) * 1 * S(w)
            ref = so.updatequadbufs(cfg, g, a, b)
        leg, jac = cfg.legrule, cfg.jacrule
        outs = [np.empty(m * k), np.empty(m * k), np.empty(2 * m * k), np.empty(2 * m * k)]
        L.emul_gen_sources.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp,
                                       ctypes.c_int] + [dp] * 12
        L.emul_gen_sources(m, k, a, b, p, int(origin), int(not origin), 0, 1, 0, _ptr(parms), 4,
                           _ptr(leg.no1), _ptr(leg.wt1), _ptr(leg.no2), _ptr(leg.wt2),
                           _ptr(jac.no1), _ptr(jac.wt1), _ptr(jac.no2), _ptr(jac.wt2), *[_ptr(o) for o in outs])
        assert np.array_equal(outs[0], ref[0]) and np.array_equal(outs[2], ref[2])      # nodes: bit-exact
        assert np.max(np.abs(outs[1] / ref[1] - 1)) < 5e-15                             # strengths: pow() ulps
        assert np.max(np.abs(outs[3] / ref[3] - 1)) < 5e-15


def test_device_rule_generator(emul):
    """sk_rules.cuh (double-double Newton, the generator the device runs) against the oracle's long-double
    rules and exact moments."""
    L, _ = emul
    L.emul_gauss_rule_dd.argtypes = [ctypes.c_int, ctypes.c_double, dp, dp]
    for n, p in ((64, 0.0), (4096, 0.0), (8192, -0.5), (4096, 1.5), (8192, 0.5), (1, 0.0), (2, -0.9)):
        no, wt = np.empty(n), np.empty(n)
        assert L.emul_gauss_rule_dd(n, p, _ptr(no), _ptr(wt)) == 0
        xo, wo = so.gauss_rule(n, p)
        assert np.all(np.diff(no) > 0) if n > 1 else True
        assert np.max(np.abs(no - xo)) <= 2.3e-16
        assert np.max(np.abs(wt / wo - 1)) <= 2e-12      # the long-double oracle loses ~1e-12 in the weights next to x = +-1 (1 - x^2 cancels); double-double does not
        for j in (0, 1, 5):
            if j > 2 * n - 1:
                continue                                   # an n-point rule is exact up to degree 2n - 1
            exact = 2 ** (p + j + 1) / (p + j + 1)
            assert abs((wt * (1 + no) ** j).sum() / exact - 1) < 2e-14


def test_lean_sincos(emul):
    L, _ = emul
    L.emul_sincos_err.restype = ctypes.c_double
    L.emul_sincos_err.argtypes = [ctypes.c_int]
    assert L.emul_sincos_err(2_000_003) <= 5e-16


def test_convergence_predicate(emul):
    L, _ = emul
    L.emul_trunc_err.restype = ctypes.c_double
    L.emul_trunc_err.argtypes = [ctypes.c_double] * 4 + [ctypes.c_int]
    L.emul_converged.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int]
    b, c, d, dim = 32768.0, 0.683, -3.96, 1
    ta = -c / (d + dim) * b ** (d + dim)
    tn = c * b ** (d + (dim - 1) / 2)
    for x in (1e-6, 1e-3, 0.5, 1.0):
        assert L.emul_trunc_err(ta, tn, 1.0, x, 0) == so.truncation_error_estimate(b, x, c, d, dim)
    for crit, name in ((0, "panel"), (1, "tails"), (2, "both")):
        for te in (1e-10, 1e-7):
            for pk in (1e-10, -1e-7):
                assert bool(L.emul_converged(te, pk, 5e-9, crit)) == bool(so.check_convergence(te, pk, 5e-9, criteria=name))
