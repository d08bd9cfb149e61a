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
    # Julia's min propagates NaN (src/adaptive.jl:225-228): a NaN bound keeps the target active; +-Inf order normally
    nan, inf = float("nan"), float("inf")
    for ta_, tn_ in ((nan, 1.0), (1.0, nan), (nan, nan), (inf, nan), (nan, -inf)):
        assert np.isnan(L.emul_trunc_err(ta_, tn_, 1.0, 0.5, 0))
        assert not L.emul_converged(L.emul_trunc_err(ta_, tn_, 1.0, 0.5, 0), 0.0, 5e-9, 1)
    assert L.emul_trunc_err(inf, 1.0, 1.0, 0.5, 0) == 1.0 / np.pi
    assert L.emul_trunc_err(-inf, 1.0, 1.0, 0.5, 0) == -inf
    assert L.emul_trunc_err(1.0, inf, 1.0, 0.5, 0) == 1.0
    for crit, name in ((0, "panel"), (1, "tails"), (2, "both")):
        for te in (1e-10, 1e-7):
            for pk in (1e-10, -1e-7):
                assert bool(L.emul_converged(te, pk, 5e-9, crit)) == bool(so.check_convergence(te, pk, 5e-9, criteria=name))


# ---- nonuniform Hankel transform (csrc/sk_hankel.h: the dim >= 2 branch, src/quadrature.jl:137-161) -------
def _hk_api(L):
    sz = [ctypes.c_int() for _ in range(6)]
    L.emul_hk_sizes(*[ctypes.byref(v) for v in sz])
    L.emul_hk_plan.restype = ctypes.c_longlong
    L.emul_hk_plan.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_double] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
    L.emul_hk_fit.argtypes = [ctypes.c_void_p, dp, ctypes.c_int, ctypes.c_longlong, dp, dp, dp]
    L.emul_hk_spread.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_longlong, dp, dp, dp]
    L.emul_hk_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, dp, dp, ctypes.c_longlong, dp, dp]
    L.emul_hk_group_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.emul_bessel_eval.restype = ctypes.c_double
    L.emul_bessel_eval.argtypes = [dp, ctypes.c_int, ctypes.c_double]
    L.emul_hk_level.argtypes = [ctypes.c_double, ctypes.c_double]
    L.emul_hk_octave.argtypes = [ctypes.c_double, ctypes.c_double]
    return [v.value for v in sz]


def _hk_transform(L, P, tab, nu, no1, buf1, no2, buf2, a, b, xs):
    plan_b, grp_b, K, NLEV, NCH, NGRP = _hk_api(L)
    H = ctypes.create_string_buffer(plan_b)
    G = ctypes.create_string_buffer(grp_b * NGRP)
    total = L.emul_hk_plan(ctypes.byref(P), nu, a, b, float(xs.min()), float(xs.max()), H, G)
    assert total >= 0
    info = (ctypes.c_int * 5)()
    L.emul_hk_plan_info(H, info)
    cheb = np.zeros((NLEV, NCH, 2))
    L.emul_hk_fit(H, _ptr(tab), 0, no1.size, _ptr(no1), _ptr(buf1), _ptr(cheb))
    L.emul_hk_fit(H, _ptr(tab), 1, no2.size, _ptr(no2), _ptr(buf2), _ptr(cheb))
    grid = np.zeros(max(total, 1), dtype=complex)
    for g in range(info[4]):
        nf2, off, D = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_double()
        L.emul_hk_group_info(G, g, ctypes.byref(nf2), ctypes.byref(off), ctypes.byref(D))
        fin = np.zeros((nf2.value, K, 2), dtype=complex)
        L.emul_hk_spread(ctypes.byref(P), H, G, g, 0, no1.size, _ptr(no1), _ptr(buf1), _ptr(fin.view(float)))
        L.emul_hk_spread(ctypes.byref(P), H, G, g, 1, no2.size, _ptr(no2), _ptr(buf2), _ptr(fin.view(float)))
        grid[off.value:off.value + nf2.value * K * 2] = (np.fft.ifft(fin, axis=0) * nf2.value).reshape(-1)
    loc = np.zeros((NLEV, 16, 16, 2))
    L.emul_hk_local_poly.argtypes = [ctypes.c_void_p, dp, dp]
    L.emul_hk_local_poly(H, _ptr(cheb), _ptr(loc))
    out = np.zeros((xs.size, 2))
    L.emul_hk_eval(ctypes.byref(P), H, G, _ptr(grid.view(float)), _ptr(loc), xs.size, _ptr(xs), _ptr(out))
    # the cell-polynomial variant of the same transform
    out_c = np.zeros((xs.size, 2))
    ncell = ctypes.c_longlong()
    L.emul_hk_eval_cells.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, dp, dp, ctypes.c_longlong, dp, dp,
                                     ctypes.c_void_p]
    L.emul_hk_eval_cells(ctypes.byref(P), H, G, _ptr(grid.view(float)), _ptr(loc), xs.size, _ptr(xs), _ptr(out_c),
                         ctypes.byref(ncell))
    info = list(info) + [ncell.value, out_c]
    return out, info


def test_bessel_table_and_dyadic_cuts(emul):
    """J_0..J_3 on [0, 64] from the quad-precision-fitted table; level / octave cuts are exact at the boundaries."""
    from scipy import special
    L, _ = emul
    _hk_api(L)
    tab = np.zeros(4 * 32 * 16)
    assert L.emul_bessel_table(_ptr(tab)) == 0
    z = np.concatenate([np.linspace(0, 64, 1501), [1e-9, 1.9999999, 2.0, 2.0000001, 63.999999]])
    for nu, ref in ((0, special.j0), (1, special.j1)):
        got = np.array([L.emul_bessel_eval(_ptr(tab), nu, float(v)) for v in z])
        assert np.max(np.abs(got - ref(z))) <= 2e-15                       # cephes j0/j1: ~1e-16 absolute
    for nu in (2, 3):
        got = np.array([L.emul_bessel_eval(_ptr(tab), nu, float(v)) for v in z])
        assert np.max(np.abs(got - special.jv(nu, z))) <= 5e-13            # scipy jv: ~1e-13
    wT = 5.09
    for q in range(1, 30):
        bq = wT * 2.0 ** (q - 1)
        assert L.emul_hk_level(wT, bq) == q and L.emul_hk_level(wT, np.nextafter(bq, 0)) == q - 1
    assert L.emul_hk_level(wT, 0.0) == 0
    r_hi = 0.8731
    for t in range(0, 30):
        edge = r_hi * 2.0 ** -t
        assert L.emul_hk_octave(r_hi, edge) == t and L.emul_hk_octave(r_hi, np.nextafter(edge, 1)) == max(t - 1, 0)


def test_hankel_plan_invariants(emul):
    """sk_hk_make_plan: every octave that needs an asymptotic part has exactly one group; the levels the groups of a
    shared set spread incrementally partition the levels above the set's first cut; every group's grid covers its
    sources and targets."""
    L, P = emul
    plan_b, grp_b, K, NLEV, NCH, NGRP = _hk_api(L)
    L.emul_hk_group_fields.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(12)
    for case in range(40):
        r_hi = float(10 ** rng.uniform(-3, 3))
        r_lo = r_hi * float(10 ** rng.uniform(-7, 0))
        npan = int(rng.integers(0, 4))
        a = npan * 32768.0 / r_hi * float(rng.uniform(0.5, 1.0)) if npan else 0.0
        b = a + 32768.0 / r_hi * float(rng.uniform(0.3, 3.0))
        H = ctypes.create_string_buffer(plan_b)
        G = ctypes.create_string_buffer(grp_b * NGRP)
        total = L.emul_hk_plan(ctypes.byref(P), int(rng.integers(0, 4)), a, b, r_lo, r_hi, H, G)
        assert total >= 0
        info = (ctypes.c_int * 5)()
        L.emul_hk_plan_info(H, info)
        q_lo, q_hi, t_full, t_last, ng = list(info)
        wT = 32.0 / (2 * np.pi * r_hi)
        assert q_lo == L.emul_hk_level(wT, a) and q_hi == L.emul_hk_level(wT, b)
        t_need = L.emul_hk_octave(r_hi, r_lo)
        n_oct = max(0, min(t_last, t_need) - (t_full + 1) + 1)
        assert ng == n_oct + (1 if t_full >= 0 else 0)
        fields = []
        for gi in range(ng):
            out = (ctypes.c_longlong * 6)()
            geo = (ctypes.c_double * 4)()
            L.emul_hk_group_fields(G, gi, out, geo)
            fields.append((list(out), list(geo)))
        off = 1 if t_full >= 0 else 0
        if t_full >= 0:
            (q_cut, q_from, q_to, shared, nf, nf2), (w_ref, wc, D, inv_hu) = fields[0]
            assert (q_cut, q_from, q_to, shared) == (0, 0, NLEV, 0) and w_ref <= max(a, wT)
        covered = None
        for i, gi in enumerate(range(off, ng)):
            t = t_full + 1 + i
            (q_cut, q_from, q_to, shared, nf, nf2), (w_ref, wc, D, inv_hu) = fields[gi]
            assert q_cut == t + 2 and q_from == q_cut
            assert w_ref <= wT * 2.0 ** (t + 1) * (1 + 1e-15)              # lam = w_ref / w <= 1 for every source of the group
            assert nf2 >= 1.999 * nf and (nf2 & (nf2 - 1) == 0 or (nf2 // 3) & (nf2 // 3 - 1) == 0)
            # the spread grid covers the sources [first cut, b] ...
            w_first = wT * 2.0 ** (t + 1)
            assert abs((w_first - wc) * inv_hu) <= nf / 2 and abs((b - wc) * inv_hu) <= nf / 2
            # ... and the fine grid the targets of the octave
            kappa = nf2 / inv_hu
            for r in (max(r_lo, r_hi * 2.0 ** -(t + 1)), r_hi * 2.0 ** -t):
                assert abs((r - D) * kappa) <= nf2 / 2 - 8
            if shared == 0:
                assert q_to == NLEV
            else:
                assert t >= 3
                nxt = fields[gi + 1][0] if gi + 1 < ng else None
                if shared == 1:
                    assert gi == ng - 1 and q_to == NLEV                       # the deepest octave takes everything above
                else:
                    assert nxt is not None and nxt[3] in (1, 2) and q_to == nxt[1] == q_cut + 1
                    assert fields[gi + 1][1] == fields[gi][1]                   # same w_ref and geometry across the set


@pytest.mark.parametrize("nu,alpha", [(0, 0.0), (1, 0.0), (0, 0.5), (2, 0.0)])
def test_hankel_transform_math(emul, nu, alpha):
    """The O(N) nonuniform Hankel transform (dyadic levels x octaves: Hankel expansion through batched type-3
    NUFFTs + local Chebyshev expansions) against the reference's direct Bessel summation
    (src/quadrature.jl:145-160) on the default quadrature panels: first panel (one transform per octave),
    second panel (all octaves share one transform) and a far panel."""
    L, P = emul
    tab = np.zeros(4 * 32 * 16)
    assert L.emul_bessel_table(_ptr(tab)) == 0
    S = lambda w: (1 + w ** 2) ** -2.5
    cfg = so.OracleConfig(S, dim=2, alpha=alpha, derivative=(nu == 1))      # (the strengths only; nu = 2 is dim = 4's K')
    rng = np.random.default_rng(nu + 1)
    xs = np.sort(np.concatenate([rng.uniform(0, 1, 40), 10 ** rng.uniform(-6, 0, 40), [1.0, 0.5, 0.25]]))
    for (a, b) in ((0.0, 32768.0), (32768.0, 65536.0), (2.0e6, 2.0e6 + 32768.0)):
        if a == 0:
            no1, buf1, no2, buf2 = so.updatequadbufs(cfg, S, a, b, p=cfg.p)
        else:
            no1, buf1, no2, buf2 = so.updatequadbufs(cfg, lambda w: w ** cfg.p * S(w), a, b)
        got, info = _hk_transform(L, P, tab, nu, no1, buf1, no2, buf2, a, b, xs)
        if a == 0:
            assert info[2] == -1 and info[4] >= 10          # no shared transform; one group per octave
        else:
            assert info[2] >= 10 and info[4] <= 2           # the whole panel is asymptotic for most octaves
        ncell, got_c = info[5], info[6]
        assert ncell >= xs.size // 3                         # most targets sit far enough from the origin for the cell path
        for col, (no, buf) in enumerate(((no1, buf1), (no2, buf2))):
            ref = so.direct_bessel(nu, no, buf, xs)
            # the direct sum itself carries ~1e-12 sum|c| (131 072 sequential additions, 2 pi w r rounded)
            assert np.max(np.abs(got[:, col] - ref)) <= 3e-12 * np.sum(np.abs(buf)), (a, col)
            # cell polynomials across the K terms (k_hankel_cells) against the per-target evaluation
            assert np.max(np.abs(got_c[:, col] - got[:, col])) <= 2e-14 * np.sum(np.abs(buf)), (a, col)
        # random strengths: the asymptotic part dominates (a smooth S leaves it at ~1e-5 of the sum)
        b1, b2 = rng.normal(size=no1.size) / no1.size, rng.normal(size=no2.size) / no2.size
        xs3 = np.ascontiguousarray(xs[::3])
        got, info = _hk_transform(L, P, tab, nu, no1, b1, no2, b2, a, b, xs3)
        ref = so.direct_bessel(nu, no2, b2, xs3)
        assert np.max(np.abs(got[:, 1] - ref)) <= 5e-15 * np.sum(np.abs(b2))
        assert np.max(np.abs(info[6][:, 1] - got[:, 1])) <= 1e-15 * np.sum(np.abs(b2))


# ---- K8: unique / sort / inverse map (csrc/sk_k8.h index arithmetic, csrc/sk_k8.cuh steps) ---------------------
def _k8(emul, xs, samp=None):
    L, _ = emul
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n = xs.size
    if samp is None:
        samp = 0                                   # the product's choice (sk_k8_samp)
    uxs = np.empty(n)
    inv = np.empty(n, dtype=np.uint32)
    nu = ctypes.c_longlong()
    diag = (ctypes.c_longlong * 4)()
    L.emul_k8.argtypes = [dp, ctypes.c_longlong, ctypes.c_uint, dp, ctypes.POINTER(ctypes.c_uint32),
                          ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    rc = L.emul_k8(_ptr(xs), n, samp, _ptr(uxs), inv.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(nu), diag)
    return rc, uxs[: nu.value], inv, list(diag)


def _k8_check(emul, xs, samp=None, expect_rc=0):
    rc, uxs, inv, diag = _k8(emul, xs, samp)
    assert rc == expect_rc, (rc, diag)
    if rc == 0:
        ref, rinv = np.unique(np.where(xs == 0, 0.0, xs), return_inverse=True)
        assert np.array_equal(uxs, ref)
        assert np.array_equal(inv, rinv.astype(np.uint32))
    return diag


def test_k8_distributions(emul):
    """The bin scheme sorts and de-duplicates exactly (vs numpy.unique) on the distance sets of the BASELINE configs
    and stays inside its slots: uniform, log-uniform, pairwise distances of random 2-D points, a pair list with a
    heavy zero lag and duplicated lags, tiny inputs, presorted input, constant input."""
    rng = np.random.default_rng(0)
    d = _k8_check(emul, rng.uniform(0, 1, 2_000_000))                       # config 2 (sampled histogram)
    assert d[1] <= 1400 and d[2] <= 12 and d[3] == 0
    d = _k8_check(emul, rng.uniform(0, 1e3, 300_000))
    assert d[1] <= 1400
    d = _k8_check(emul, 10 ** rng.uniform(-6, 0, 1_500_000))                # log-uniform (shrinking active sets)
    assert d[1] <= 1500
    pts = rng.uniform(0, 1, (1500, 2))                                      # config 3: all pairwise distances
    iu = np.triu_indices(1500, 1)
    lag = np.sqrt(((pts[iu[0]] - pts[iu[1]]) ** 2).sum(1))
    d = _k8_check(emul, lag)
    assert d[1] <= 1500
    # config 5: a pair list with ~12% zero lags (the diagonal of every block) and every lag repeated a few times
    base = rng.uniform(0, 0.05, 150_000)
    lagv = np.concatenate([np.zeros(60_000), np.repeat(base, 3), base[:30_000]])
    rng.shuffle(lagv)
    d = _k8_check(emul, lagv)
    assert d[1] <= 1500
    for n in (1, 2, 3, 100, 2048, 2049, 5000):                              # small inputs, around the single-bin limit
        _k8_check(emul, rng.uniform(0, 2, n))
    _k8_check(emul, np.zeros(7))
    _k8_check(emul, np.array([0.0, 0.3, 0.0, 0.3, 1e-300, 5e-324, 1.7e308]))
    _k8_check(emul, np.full(1000, 0.25))                                    # one value, 1000 times: one group of 1000
    d = _k8_check(emul, np.sort(rng.uniform(0, 1, 5000)))
    assert d[3] == 1                                                        # presorted: identity
    srt = np.sort(rng.uniform(0, 1, 3_000_000))
    srt[1000] = srt[999]                                                    # sorted but not unique: the bins must cope
    d = _k8_check(emul, srt)
    assert d[3] == 0 and d[1] <= 1500
    per = np.tile(rng.uniform(0, 1, 256), 8192)                             # periodic input, period = 256 = 8 segments
    _k8_check(emul, per, expect_rc=1)                                       # 256 values x 8192 copies: general sort
    assert _k8(emul, np.array([0.5, -1.0]))[0] == 2 and _k8(emul, np.array([0.5, np.nan]))[0] == 2


def test_k8_clustered_input_overflows_cleanly(emul):
    """Heavily clustered distances outgrow a fine bin: reported (the product then takes the general sort), never wrong."""
    rng = np.random.default_rng(4)
    spread = rng.uniform(0, 2.0, 5000)
    clustered = 0.5 + rng.uniform(0, 1e-9, 3000)
    rc, _, _, diag = _k8(emul, np.concatenate([spread, clustered]))
    assert rc == 1 and diag[1] > 2048
