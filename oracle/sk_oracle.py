"""
sk_oracle.py -- TEST INFRASTRUCTURE ONLY (the parity oracle).

A CPU restatement (numpy + the C helpers in sk_oracle.c) of the K(r) hot path of
pbeckman/SpectralKernels.jl: `AdaptiveKernelConfig` + `kernel_values`.  Every
function cites the reference file:line it follows.  The transform is the
reference's own *direct summation* branch (src/quadrature.jl:113-128), which is
the definition of what `finufft1d3` (src/utils.jl:10) approximates; optionally
the from-scratch CPU type-3 NUFFT in sk_oracle.c (`transform="nufft"`) for sizes
where direct summation is too slow (this is also the CPU baseline in bench.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl
reference` legs may import this module.  The product (spectralkernels.jl_b200)
never does.

Parity pinning (see oracle/README.md): the reference cannot run in this image
(no Julia, no FINUFFT).  The oracle is pinned end-to-end against every
closed-form known answer the reference's own tests hold for this path
(test/exponential_sdf_1d.jl, test/matern_sdf.jl, test/derivatives/warping.jl,
test/derivatives/sdf_params.jl) at the reference's own tolerances -- see
tests/test_oracle_golden.py.  Panel *traces* have no reference-side fixture
anywhere: "parity unpinned" for traces.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsk_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/libsk_oracle.so with oracle/Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "sk_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.sko_direct_cis.argtypes = [ctypes.c_int64, dp, dp, dp, ctypes.c_int64, dp, dp, dp]
        L.sko_direct_cis.restype = None
        L.sko_direct_bessel.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, ctypes.c_int64, dp, dp]
        L.sko_direct_bessel.restype = None
        L.sko_gauss_rule.argtypes = [ctypes.c_int, ctypes.c_double, dp, dp]
        L.sko_gauss_rule.restype = ctypes.c_int
        L.sko_nufft1d3.argtypes = [ctypes.c_int64, dp, dp, ctypes.c_int64, dp, dp, ctypes.c_double]
        L.sko_nufft1d3.restype = ctypes.c_int
        L.sko_num_threads.restype = ctypes.c_int
        L.sko_set_num_threads.argtypes = [ctypes.c_int]
        L.sko_set_num_threads.restype = None
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def num_threads() -> int:
    return int(lib().sko_num_threads())


def set_num_threads(n: int) -> None:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants all host cores."""
    lib().sko_set_num_threads(int(n))


# --------------------------------------------------------------------------- #
# Quadrature rules -- src/quadrature.jl:27-47 (QuadRule)                      #
# --------------------------------------------------------------------------- #
_RULE_CACHE = {}


def gauss_rule(n: int, p: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
    """gausslegendre(n) (p == 0) or gaussjacobi(n, 0.0, p): weight (1+x)^p on [-1,1]."""
    key = (int(n), float(p))
    if key not in _RULE_CACHE:
        x = np.empty(n)
        w = np.empty(n)
        rc = lib().sko_gauss_rule(int(n), float(p), _ptr(x), _ptr(w))
        if rc != 0:
            raise RuntimeError("oracle Gauss rule generation failed (nodes not ascending)")
        _RULE_CACHE[key] = (x, w)
    return _RULE_CACHE[key]


@dataclass
class QuadRule:
    """src/quadrature.jl:27-47"""
    no1: np.ndarray
    wt1: np.ndarray
    no2: np.ndarray
    wt2: np.ndarray

    @staticmethod
    def make(m: int, case: str, p: float = 0.0) -> "QuadRule":
        if case == "legendre" or p == 0.0:           # quadrature.jl:36-38
            (no1, wt1), (no2, wt2) = gauss_rule(m), gauss_rule(2 * m)
        elif case == "jacobi":                        # quadrature.jl:39-42
            if p <= -1.0:
                raise ValueError("p needs to be in (-1.0, Inf) to be integrable")
            (no1, wt1), (no2, wt2) = gauss_rule(m, p), gauss_rule(2 * m, p)
        else:
            raise ValueError("Options are case=:legendre or case=:jacobi")
        return QuadRule(no1, wt1, no2, wt2)


# --------------------------------------------------------------------------- #
# Config -- src/adaptive.jl:2-59                                              #
# --------------------------------------------------------------------------- #
@dataclass
class OracleConfig:
    f: Callable
    df: Optional[Callable] = None
    dim: int = 1
    alpha: float = 0.0
    tol: float = 1e-8
    derivative: bool = False
    logw: bool = False
    convergence_criteria: str = "both"
    tail: Optional[float] = None
    quadspec: Tuple[int, int] = (2 ** 12, 2 ** 4)
    # derived
    c: float = field(init=False)
    p: float = field(init=False)
    legrule: QuadRule = field(init=False)
    jacrule: QuadRule = field(init=False)

    def __post_init__(self):
        if self.convergence_criteria not in ("panel", "tails", "both"):   # adaptive.jl:29-31
            raise ValueError("Argument convergence_criteria must be one of :panel, :tails, :both.")
        if self.alpha >= self.dim:                                        # adaptive.jl:33-35
            raise ValueError("alpha must be less than dim to be integrable.")
        if self.tol < 1e-12 and self.quadspec[0] * self.quadspec[1] > 2 ** 12:   # adaptive.jl:37-40
            self.quadspec = (2 ** 12, 1)
        self.p = -self.alpha + (0 if self.dim == 1 else self.dim / 2) + (1 if self.derivative else 0)  # :42
        c = 2.0 if self.dim == 1 else 2 * math.pi                          # :43
        if self.derivative:
            c *= -2 * math.pi                                              # :44
        if self.logw:
            c *= -1                                                        # :45
        self.c = c
        m, _k = self.quadspec
        self.legrule = QuadRule.make(m, "legendre")                        # :48
        self.jacrule = QuadRule.make(m, "jacobi", self.p) if self.p != 0.0 else self.legrule   # :49

    @property
    def quadsz(self) -> int:                                               # adaptive.jl:93
        return self.quadspec[0] * self.quadspec[1]


def gen_derivative_config(cfg: OracleConfig) -> OracleConfig:             # adaptive.jl:61-66
    return OracleConfig(cfg.f, df=cfg.df, derivative=True, dim=cfg.dim, alpha=cfg.alpha, tol=cfg.tol,
                        logw=cfg.logw, tail=cfg.tail, convergence_criteria=cfg.convergence_criteria,
                        quadspec=cfg.quadspec)


def gen_new_sdf_config(cfg: OracleConfig, new_f, alpha=None) -> OracleConfig:   # adaptive.jl:69-72
    return OracleConfig(new_f, df=cfg.df, dim=cfg.dim, alpha=cfg.alpha if alpha is None else alpha, tol=cfg.tol)


def compute_k0(cfg: OracleConfig) -> float:
    """src/adaptive.jl:74-91.  QuadGK.jl (third party) is replaced by QUADPACK's
    qagi through scipy.integrate.quad with the same atol=0, rtol=min(1e-8, 1e-2 tol)."""
    from scipy import integrate, special
    f = lambda w: float(np.asarray(cfg.f(np.asarray([w], dtype=float)))[0])
    p = cfg.p
    L = 1.0
    while L ** p * abs(f(L)) > abs(f(0.0)) / 2:                            # :78-80
        L *= 2
    if cfg.dim == 1:
        def integrand(w):                                                 # :82
            wl = w * L
            if wl == 0.0 and (p < 0 or cfg.logw):
                return 0.0
            return wl ** p * (math.log(wl) if cfg.logw else 1.0) * f(wl) * L
    else:
        nu = cfg.dim / 2 - 1 + (1 if cfg.derivative else 0)                # :85
        def integrand(w):                                                 # :86
            wl = w * L
            if wl == 0.0 and (p < 0 or cfg.logw):
                return 0.0
            return (math.pi * w) ** nu / special.gamma(nu + 1) * wl ** p * \
                (math.log(wl) if cfg.logw else 1.0) * f(wl) * L
    rtol = min(1e-8, 1e-2 * cfg.tol)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        val = integrate.quad(integrand, 0.0, np.inf, epsabs=0.0, epsrel=max(rtol, 5e-14), limit=400)[0]
    return cfg.c * val                                                     # :88-90


# --------------------------------------------------------------------------- #
# updatequadbufs! -- src/quadrature.jl:49-95                                  #
# --------------------------------------------------------------------------- #
def _subpanel_edges(a: float, b: float, k: int) -> np.ndarray:
    """range(a, b, length=k+1) (quadrature.jl:56).  Julia's StepRangeLen carries the
    step in twice precision, i.e. each element is a + i*(b-a)/k rounded once; long
    double reproduces that for the k used here."""
    al, bl = np.longdouble(a), np.longdouble(b)
    i = np.arange(k + 1, dtype=np.longdouble)
    e = (al + i * ((bl - al) / np.longdouble(k))).astype(np.float64)
    e[0], e[-1] = a, b
    return e


def updatequadbufs(cfg: OracleConfig, f: Callable, a: float, b: float, p: float = 0.0):
    m, k = cfg.quadspec
    leg, jac = cfg.legrule, cfg.jacrule
    no1 = np.empty(m * k); buf1 = np.empty(m * k)
    no2 = np.empty(2 * m * k); buf2 = np.empty(2 * m * k)
    edges = _subpanel_edges(a, b, k)
    first = 0
    if p != 0 and a == 0.0:                                               # :61-78 Jacobi at the origin
        sa, sb = edges[0], edges[1]
        bmad2, bpad2 = (sb - sa) / 2, (sb + sa) / 2
        no1[:m] = bmad2 * jac.no1 + bpad2
        buf1[:m] = jac.wt1 * bmad2 ** (p + 1) * f(no1[:m])
        no2[:2 * m] = bmad2 * jac.no2 + bpad2
        buf2[:2 * m] = jac.wt2 * bmad2 ** (p + 1) * f(no2[:2 * m])
        first = 1
    for i in range(first, k):                                             # :82-92 Legendre elsewhere
        sa, sb = edges[i], edges[i + 1]
        bmad2, bpad2 = (sb - sa) / 2, (sb + sa) / 2
        s1 = slice(i * m, (i + 1) * m)
        no1[s1] = bmad2 * leg.no1 + bpad2
        buf1[s1] = leg.wt1 * bmad2 * _pow(no1[s1], p) * f(no1[s1])
        s2 = slice(i * 2 * m, (i + 1) * 2 * m)
        no2[s2] = bmad2 * leg.no2 + bpad2
        buf2[s2] = leg.wt2 * bmad2 * _pow(no2[s2], p) * f(no2[s2])
    return no1, buf1, no2, buf2


def _pow(x: np.ndarray, p: float) -> np.ndarray:
    return np.ones_like(x) if p == 0 else np.power(x, p)


# --------------------------------------------------------------------------- #
# transforms                                                                  #
# --------------------------------------------------------------------------- #
def direct_cis(no: np.ndarray, buf: np.ndarray, xs: np.ndarray) -> np.ndarray:
    """src/quadrature.jl:113-128: int[j] = sum_k buf[k]*cispi(2*no[k]*x[j])."""
    no = np.ascontiguousarray(no, dtype=np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    bre = np.ascontiguousarray(np.real(buf), dtype=np.float64)
    bim = np.ascontiguousarray(np.imag(buf), dtype=np.float64) if np.iscomplexobj(buf) else None
    ore = np.empty(xs.size); oim = np.empty(xs.size)
    lib().sko_direct_cis(no.size, _ptr(no), _ptr(bre), _ptr(bim) if bim is not None else None,
                         xs.size, _ptr(xs), _ptr(ore), _ptr(oim))
    return ore + 1j * oim


def direct_bessel(nu: int, no: np.ndarray, buf: np.ndarray, xs: np.ndarray) -> np.ndarray:
    """src/quadrature.jl:145-160."""
    no = np.ascontiguousarray(no, dtype=np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    buf = np.ascontiguousarray(buf, dtype=np.float64)
    out = np.empty(xs.size)
    lib().sko_direct_bessel(int(nu), no.size, _ptr(no), _ptr(buf), xs.size, _ptr(xs), _ptr(out))
    return out


def cpu_nufft1d3(w: np.ndarray, s: np.ndarray, x: np.ndarray, eps: float = 1e-15) -> np.ndarray:
    """Contract of finufft1d3(w, s, x) (src/utils.jl:10): f_j = sum_k s_k exp(+i 2 pi x_j w_k)."""
    w = np.ascontiguousarray(w, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    sc = np.ascontiguousarray(np.asarray(s, dtype=np.complex128))
    out = np.empty(x.size, dtype=np.complex128)
    rc = lib().sko_nufft1d3(w.size, _ptr(w), sc.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                            x.size, _ptr(x), out.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                            float(eps))
    if rc != 0:
        raise MemoryError("oracle nufft1d3 failed")
    return out


def nufft_quad_size_cutoff(n_no: int, n_x: int) -> bool:                  # utils.jl:39
    return n_no * n_x > 2 ** 18


def check_subdivide_failure(a: float, b: float):                          # utils.jl:28-36
    if abs(b - a) > 1e-16:
        return
    raise RuntimeError(f"The sub-interval (a, b) = ({a}, {b}) has been split too many times "
                       f"(b - a < 1e-16). Exiting to avoid infinite splitting.")


# --------------------------------------------------------------------------- #
# fourier_integrate_panel -- src/quadrature.jl:97-167                         #
# --------------------------------------------------------------------------- #
def fourier_integrate_panel(cfg, f, a, b, xs, p=0.0, kernel="cis", transform="direct", stats=None):
    check_subdivide_failure(a, b)                                         # :98
    no1, buf1, no2, buf2 = updatequadbufs(cfg, f, a, b, p=p)              # :100-103
    fast = nufft_quad_size_cutoff(no2.size, xs.size) and xs.size > 1      # :105
    if stats is not None:
        stats.append(("fast" if fast else "direct", xs.size))
    if kernel in ("cis", "cos", "sin"):
        if fast and transform == "nufft":                                 # :109-110 (FINUFFT in the reference)
            int1 = cpu_nufft1d3(no1, buf1, xs)
            int2 = cpu_nufft1d3(no2, buf2, xs)
        else:                                                             # :113-128 (the definition)
            int1 = direct_cis(no1, buf1, xs)
            int2 = direct_cis(no2, buf2, xs)
        if kernel == "cos":                                               # :130-132
            int1, int2 = int1.real.copy(), int2.real.copy()
        elif kernel == "sin":                                             # :133-135
            int1, int2 = int1.imag.copy(), int2.imag.copy()
    elif isinstance(kernel, tuple) and len(kernel) == 2 and kernel[0] == "J":     # :137-161
        nu = int(kernel[1])
        if nu != kernel[1]:
            raise ValueError("InexactError: Int64(%r)" % (kernel[1],))    # :138
        int1 = direct_bessel(nu, no1, buf1, xs)
        int2 = direct_bessel(nu, no2, buf2, xs)
    else:
        raise ValueError("integral kernel must be :cis, :sin, :cos, or (:J, nu)")
    # :165  any(isnan,int1) || any(isnan,int2) && throw(...)
    if (not np.isnan(int1).any()) and np.isnan(int2).any():
        raise RuntimeError("NaN detected in panel integral...")
    return int1, int2


# --------------------------------------------------------------------------- #
# fourier_integrate_interval -- src/quadrature.jl:169-275                     #
# --------------------------------------------------------------------------- #
def fourier_integrate_interval(a, b, cfg: OracleConfig, xs, k0, trace=None, transform="direct"):
    from scipy import special
    f, df, dim, alpha = cfg.f, cfg.df, cfg.dim, cfg.alpha
    stack = [(a, b, cfg.tol)]                                             # :173 (LIFO, :2-25)
    I = np.zeros(xs.size)
    err = np.zeros(xs.size)
    if dim == 1:
        kernel = "sin" if cfg.derivative else "cos"                       # :177
    else:
        kernel = ("J", dim / 2) if cfg.derivative else ("J", dim / 2 - 1)  # :179
    while stack:
        _a, _b, _tol = stack.pop()                                        # :183
        if _a == 0.0 and cfg.p != 0.0:                                    # :185
            if cfg.logw:                                                  # :186-228
                I0 = _b ** (dim / 2 + 1 - alpha) * math.log(_b) * _scalar(f, _b) * \
                    special.jv(dim / 2 - 1, 2 * math.pi * _b * xs)
                fa = lambda w: f(w) + w * np.log(w) * df(w)
                fb = lambda w: w * np.log(w) * f(w)
                if dim == 1:
                    I1a, I2a = fourier_integrate_panel(cfg, fa, _a, _b, xs, p=cfg.p, kernel="cis", transform=transform)
                    I1b, I2b = fourier_integrate_panel(cfg, fb, _a, _b, xs, p=cfg.p, kernel="cis", transform=transform)
                    I1a, I2a, I1b, I2b = I1a.real, I2a.real, I1b.imag, I2b.imag
                elif dim == 2:
                    I1a, I2a = fourier_integrate_panel(cfg, fa, _a, _b, xs, p=cfg.p, kernel=("J", int(dim / 2 - 1)))
                    I1b, I2b = fourier_integrate_panel(cfg, fb, _a, _b, xs, p=cfg.p, kernel=("J", int(dim / 2)))
                else:
                    raise NotImplementedError("singularity derivative not implemented in d > 2")
                I1 = (I0 - I1a + 2 * math.pi * xs * I1b) / (dim - alpha)
                I2 = (I0 - I2a + 2 * math.pi * xs * I2b) / (dim - alpha)
            else:                                                         # :230-238
                I1, I2 = fourier_integrate_panel(cfg, f, _a, _b, xs, p=cfg.p, kernel=kernel, transform=transform)
        else:                                                             # :240-247
            pw = cfg.p
            if cfg.logw:
                g = lambda w: _pow(w, pw) * np.log(w) * f(w)
            else:
                g = lambda w: _pow(w, pw) * 1 * f(w)
            I1, I2 = fourier_integrate_panel(cfg, g, _a, _b, xs, kernel=kernel, transform=transform)
        I1 = I1 * cfg.c                                                   # :250-251
        I2 = I2 * cfg.c
        if dim > 1:                                                       # :252-254
            I1 = I1 / xs ** (dim / 2 - 1)
            I2 = I2 / xs ** (dim / 2 - 1)
        _err = np.abs(I2 - I1)                                            # :257
        max_I_error = np.max(_err) if not np.isnan(_err).any() else float("nan")   # :258
        accepted = bool(max_I_error < cfg.tol * k0)                       # :260 (uses cfg.tol, not _tol)
        if trace is not None:
            trace.append({"kind": "subinterval", "a": float(_a), "b": float(_b), "n_act": int(xs.size),
                          "rel_err": float(max_I_error / k0), "accepted": accepted})
        if accepted:
            I += I2                                                       # :261
            err += _err                                                   # :262
        else:                                                             # :268-270
            tl, tr = (9 * _tol / 10, _tol / 10) if _a == 0 else (_tol / 2, _tol / 2)
            mid = (_a + _b) / 2
            stack.append((_a, mid, tl))
            stack.append((mid, _b, tr))
    return I, err


def _scalar(f, w):
    return float(np.asarray(f(np.asarray([w], dtype=float)))[0])


# --------------------------------------------------------------------------- #
# tail estimate / convergence -- src/adaptive.jl:204-233                      #
# --------------------------------------------------------------------------- #
def estimate_tail_decay(cfg: OracleConfig, a, b, d=None):
    nf = 1000                                                             # :208
    start = a + (b - a)                                                   # :210 (equals b up to one rounding)
    ws = np.linspace(start, b, nf) if start != b else np.full(nf, b)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if d is None:
            tmp = np.log(np.abs(cfg.f(ws)))                               # :213
            A = np.stack([np.ones(nf), np.log(ws)], axis=1)
            if not np.all(np.isfinite(tmp)):
                d = float("nan")
            else:
                # Julia's `\` on a rank-deficient tall matrix returns the minimum-norm
                # least-squares solution (pivoted QR + complete orthogonal factorisation);
                # numpy's SVD-based lstsq returns the same solution.            :214
                sol = np.linalg.lstsq(A, tmp, rcond=None)[0]
                d = float(sol[1])
        d = d - cfg.alpha                                                 # :216
        c = float(np.sum(ws ** d * np.abs(cfg.f(ws))) / np.sum(ws ** (2 * d)))   # :218
    return c, d


def truncation_error_estimate(b, x, c, d, dim):                           # adaptive.jl:222-229
    with np.errstate(all="ignore"):
        t1 = -c / (d + dim) * b ** (d + dim)
        t2 = c * b ** (d + (dim - 1) / 2) / (2 * math.pi * x ** ((dim + 1) / 2))
    return np.minimum(t1, t2)


def check_convergence(trunc_err, panel_k, tol, criteria="both"):          # adaptive.jl:231-233
    return (criteria == "panel" or trunc_err < tol) and (criteria == "tails" or abs(panel_k) < tol)


# --------------------------------------------------------------------------- #
# kernel_values / _kernel_values -- src/adaptive.jl:95-202                    #
# --------------------------------------------------------------------------- #
def kernel_values(cfg: OracleConfig, xs, k0=None, param_derivative=False, trace: Optional[List] = None,
                  transform="direct"):
    xs = np.asarray(xs, dtype=np.float64)
    if k0 is None:
        k0 = compute_k0(cfg)                                              # :97
    # unique(xs) keeps first occurrences (:99); values are scattered back through
    # a Dict keyed by x (:105-107).  Order of the unique set is irrelevant for the
    # result because _kernel_values sorts it (:113-120).
    uxs, inv = np.unique(xs, return_inverse=True)                         # sorted unique
    uvals, uerrs = _kernel_values(cfg, uxs, k0, param_derivative=param_derivative, trace=trace,
                                  transform=transform)
    return uvals[inv], uerrs[inv]


def _kernel_values(cfg: OracleConfig, xs, k0, param_derivative=False, trace=None, transform="direct"):
    xs = np.asarray(xs, dtype=np.float64)
    if xs.size > 1 and not np.all(xs[1:] >= xs[:-1]):                     # :113-120
        sp = np.argsort(xs, kind="stable")
        ip = np.empty_like(sp); ip[sp] = np.arange(sp.size)
        skv, serr = _kernel_values(cfg, xs[sp], k0, param_derivative=param_derivative, trace=trace,
                                   transform=transform)
        return skv[ip], serr[ip]
    n = xs.size
    ks = np.zeros(n); errs = np.zeros(n)                                  # :122
    hi = n                                                                # :123 (1-based index)
    quadm = cfg.quadsz
    conv_crit = cfg.convergence_criteria
    a = b = 0.0
    c = d = float("nan")
    ix1 = 1
    if n > 0 and xs[0] == 0:                                              # :133-146
        ix1 = 2
        if cfg.derivative:
            ks[0], errs[0] = 0.0, float("nan")
        elif param_derivative:
            ks[0], errs[0] = compute_k0(cfg), float("nan")
        else:
            ks[0], errs[0] = k0, float("nan")
    ipanel = 0
    while hi > 0 and xs[hi - 1] > 0:                                      # :149
        a, b = b, b + quadm / (2 * xs[hi - 1])                            # :152
        sl = slice(ix1 - 1, hi)
        pk, pe = fourier_integrate_interval(a, b, cfg, xs[sl].copy(), abs(k0), trace=trace,
                                            transform=transform)          # :157-159
        ks[sl] += pk                                                      # :163
        errs[sl] += pe                                                    # :164
        if conv_crit == "panel":                                          # :168
            c, d = float("nan"), float("nan")
        else:
            c, d = estimate_tail_decay(cfg, a, b, d=cfg.tail)
        if (math.isnan(c) or math.isnan(d)) and conv_crit != "panel":     # :170-175
            conv_crit = "panel"
        tau = cfg.tol * abs(k0) / 2                                       # :191
        hi_before = hi
        if conv_crit == "panel":
            te = np.zeros(hi - ix1 + 1)
        else:
            te = truncation_error_estimate(b, xs[sl], c, d, cfg.dim)
        # :185-197 -- walk ix = hi, hi-1, ... while converged.  Vectorised: the walk stops at the
        # largest index whose predicate is false; everything above it gets errs += 2*trunc_err.
        pk_abs = np.abs(pk)
        ok = np.ones(hi - ix1 + 1, dtype=bool)
        if conv_crit != "panel":
            ok &= te < tau                                                # check_convergence, :231-233
        if conv_crit != "tails":
            ok &= pk_abs < tau
        bad = np.nonzero(~ok)[0]
        ix = ix1 - 1 if bad.size == 0 else ix1 + int(bad[-1])
        if ix < hi:
            errs[ix:hi] += 2 * te[ix - ix1 + 1:]
        hi = ix                                                           # :198
        if trace is not None:
            trace.append({"kind": "panel", "index": ipanel, "a": float(a), "b": float(b),
                          "hi_before": int(hi_before), "hi_after": int(hi), "ix1": int(ix1),
                          "c": float(c), "d": float(d), "criteria": conv_crit})
        ipanel += 1
    return ks, errs
