"""
Host side of the K(r) path: the reference's public API for this path with the same names, argument
meaning and error behaviour --

    AdaptiveKernelConfig(f; df, dim, alpha, tol, derivative, logw, convergence_criteria, tail, quadspec)
                                                                         (src/adaptive.jl:2-59)
    kernel_values(cfg, xs; k0, param_derivative, verbose) -> (values, errors)   (src/adaptive.jl:95-108)

-- keeping only the scalar control flow of the adaptive panel driver (src/adaptive.jl:149-200) and of
the bisection loop (src/quadrature.jl:181-272).  Every O(N) pass runs on the GPU behind the C ABI of
include/spectralkernels_b200.h (ctypes here, `ccall` in Julia: INTEGRATION.md).  This module is what
the Julia host does, written in Python because the image has no Julia.  There is no CPU fallback.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import SK_CRIT, SK_KERNEL_BESSEL, SK_KERNEL_COS, SK_KERNEL_SIN, GroupSession, ScanArgs, Session, SkError
from .sdf import is_builtin

_CRITERIA = ("panel", "tails", "both")


def _sym(s) -> str:
    return str(s).lstrip(":")


class AdaptiveKernelConfig:
    """Mirror of `AdaptiveKernelConfig` (src/adaptive.jl:2-59).

    `f` is a callable S(w) accepting numpy arrays, or a built-in family from `sdf` (device-evaluated).
    Extra, optional keywords (reference-compatible defaults): `device` (CUDA ordinal), `devices` (a list of CUDA
    ordinals: the same single-caller `kernel_values(cfg, xs)` then spreads the distances over these GPUs -- one
    process, one host thread, no collective library: the C ABI's "device group"), `nufft_eps` (the reference
    hard-wires 1e-15 in src/utils.jl:10), `engine` (an object implementing the Session interface; used by the
    multi-process tests)."""

    def __init__(self, f: Callable, *, df: Optional[Callable] = None, dim: int = 1, alpha: float = 0.0,
                 tol: float = 1e-8, derivative: bool = False, logw: bool = False,
                 convergence_criteria="both", tail: Optional[float] = None,
                 quadspec: Tuple[int, int] = (2 ** 12, 2 ** 4), device: int = 0, nufft_eps: float = 1e-15,
                 engine=None, devices: Optional[Sequence[int]] = None):
        crit = _sym(convergence_criteria)
        if crit not in _CRITERIA:                                              # adaptive.jl:29-31
            raise ValueError("Argument convergence_criteria must be one of :panel, :tails, :both.")
        if alpha >= dim:                                                       # adaptive.jl:33-35
            raise ValueError("alpha must be less than dim to be integrable.")
        quadspec = (int(quadspec[0]), int(quadspec[1]))
        if tol < 1e-12 and quadspec[0] * quadspec[1] > 2 ** 12:                # adaptive.jl:37-40
            warnings.warn("Tolerances ε < 1e-12 are not recommended. Switching to a smaller quadrature rule for "
                          "higher accuracy (but slower) computations.")
            quadspec = (2 ** 12, 1)
        self.f, self.df = f, df
        self.dim, self.alpha, self.tol = int(dim), float(alpha), float(tol)
        self.derivative, self.logw = bool(derivative), bool(logw)
        self.convergence_criteria, self.tail, self.quadspec = crit, tail, quadspec
        self.p = -self.alpha + (0 if self.dim == 1 else self.dim / 2) + (1 if self.derivative else 0)   # :42
        c = 2.0 if self.dim == 1 else 2 * math.pi                               # :43
        if self.derivative:
            c *= -2 * math.pi                                                   # :44
        if self.logw:
            c *= -1                                                             # :45
        self.c = c
        self.devices = [int(d) for d in devices] if devices is not None else None
        self.device, self.nufft_eps = (self.devices[0] if self.devices else int(device)), float(nufft_eps)
        self._engine = engine
        self._rules = None       # host copies of the canonical rules (for host-evaluated integrands)

    # -- the config owns its device scratch, like cfg.buffers / cfg.splittingheap (adaptive.jl:17-21) --
    @property
    def engine(self):
        if self._engine is None:
            self._engine = GroupSession(self.devices) if (self.devices and len(self.devices) > 1) else Session(self.device)
            if self.nufft_eps != 1e-15:
                self._engine.set_nufft_eps(self.nufft_eps)
        return self._engine

    @property
    def quadsz(self) -> int:                                                    # adaptive.jl:93
        return self.quadspec[0] * self.quadspec[1]

    def _kw(self):
        return dict(df=self.df, dim=self.dim, alpha=self.alpha, tol=self.tol, derivative=self.derivative,
                    logw=self.logw, convergence_criteria=self.convergence_criteria, tail=self.tail,
                    quadspec=self.quadspec, device=self.device, nufft_eps=self.nufft_eps)


def gen_derivative_config(cfg: AdaptiveKernelConfig) -> AdaptiveKernelConfig:   # adaptive.jl:61-66
    kw = cfg._kw()
    kw["derivative"] = True
    out = AdaptiveKernelConfig(cfg.f, **kw)
    out._engine = cfg._engine
    return out


def gen_new_sdf_config(cfg: AdaptiveKernelConfig, new_f, alpha=None) -> AdaptiveKernelConfig:   # adaptive.jl:69-72
    out = AdaptiveKernelConfig(new_f, df=cfg.df, dim=cfg.dim, alpha=cfg.alpha if alpha is None else alpha,
                               tol=cfg.tol, device=cfg.device, nufft_eps=cfg.nufft_eps)
    out._engine = cfg._engine
    return out


def _f_scalar(f, w: float) -> float:
    return float(np.asarray(f(np.asarray([w], dtype=np.float64)))[0])


def compute_k0(cfg: AdaptiveKernelConfig) -> float:
    """src/adaptive.jl:74-91 (host scalar work, stays on the host as in the reference).  QuadGK's
    quadgk(f, 0, Inf; atol=0, rtol=min(1e-8, 1e-2 tol)) becomes QUADPACK qagi with the same tolerances."""
    from scipy import integrate, special
    f, p = cfg.f, cfg.p
    L = 1.0
    while L ** p * abs(_f_scalar(f, L)) > abs(_f_scalar(f, 0.0)) / 2:          # :78-80
        L *= 2
    if cfg.dim == 1:
        pref = lambda w: 1.0
    else:
        nu = cfg.dim / 2 - 1 + (1 if cfg.derivative else 0)                     # :85
        pref = lambda w: (math.pi * w) ** nu / special.gamma(nu + 1)

    def integrand(w):                                                          # :82 / :86
        wl = w * L
        if wl == 0.0 and (p < 0 or cfg.logw):
            return 0.0
        return pref(w) * wl ** p * (math.log(wl) if cfg.logw else 1.0) * _f_scalar(f, wl) * L

    rtol = min(1e-8, 1e-2 * cfg.tol)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        val = integrate.quad(integrand, 0.0, np.inf, epsabs=0.0, epsrel=max(rtol, 5e-14), limit=400)[0]
    return cfg.c * val                                                         # :88-90


def estimate_tail_decay(cfg: AdaptiveKernelConfig, a: float, b: float, d=None):
    """src/adaptive.jl:204-220, including its quirk: `range(a + (b-a), stop=b, length=1000)` is 1000
    copies of b, so the least-squares fit is rank one and Julia's `\\` returns the minimum-norm solution."""
    nf = 1000
    start = a + (b - a)
    if start == b:
        # all 1000 abscissae equal b: the design matrix [1 log b] has identical rows, and the minimum-norm
        # least-squares solution is (1, L) t / (1 + L^2) with L = log b, t = log|f(b)|;  the ratio of sums
        # in :218 collapses to |f(b)| b^d / b^(2d).
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            fb = abs(_f_scalar(cfg.f, b))
            if d is None:
                t = math.log(fb) if fb > 0 else -math.inf
                L = math.log(b)
                d = (L * t / (1.0 + L * L)) if math.isfinite(t) else float("nan")      # :213-214
            d = d - cfg.alpha                                                          # :216
            try:
                c = float((b ** d * fb) / (b ** (2 * d)))                              # :218
            except (OverflowError, ZeroDivisionError):
                c = float("nan")
        return c, d
    ws = np.linspace(start, b, nf)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        fw = np.abs(np.asarray(cfg.f(ws), dtype=np.float64))
        if d is None:
            tmp = np.log(fw)                                                   # :213
            if not np.all(np.isfinite(tmp)):
                d = float("nan")
            else:
                A = np.stack([np.ones(nf), np.log(ws)], axis=1)
                d = float(np.linalg.lstsq(A, tmp, rcond=None)[0][1])           # :214
        d = d - cfg.alpha                                                      # :216
        c = float(np.sum(ws ** d * fw) / np.sum(ws ** (2 * d)))                # :218
    return c, d


def _scan_args(cfg, b, c, d, tau, crit) -> ScanArgs:
    """The target-independent pieces of truncation_error_estimate (src/adaptive.jl:222-229)."""
    dim = cfg.dim
    if crit == "panel":
        ta = tn = 0.0
    else:
        # IEEE arithmetic throughout, as in Julia: d + dim == 0 gives -c/0 = -+Inf (converged by the tail bound),
        # never a Python ZeroDivisionError
        c64, d64, b64 = np.float64(c), np.float64(d), np.float64(b)
        with np.errstate(all="ignore"):
            ta = float(-c64 / (d64 + dim) * b64 ** (d64 + dim))
            tn = float(c64 * b64 ** (d64 + (dim - 1) / 2))
    return ScanArgs(ta, tn, (dim + 1) / 2, float(tau), SK_CRIT[crit], 0)


# The host's scalar work for the next panel (tail fit, scan arguments) runs while the device sorts / integrates
# (sk_targets_begin/_end, sk_subinterval_begin/_end).  False: the plain blocking calls (A/B, diagnostics).
OVERLAP_HOST_WORK = True


def _may_chain(comm) -> bool:
    """Chained launches (sk_first_panel_early, sk_subinterval_chain): single-GPU runs; sharded runs whose collectives go
    over peer mailboxes with 2 ranks (more on request, SK_SHARDED_CHAIN=1: the guards then read the global scalars and a
    skipped launch makes its exchange void -- it pays on 2 GPUs, not on 8, see sk_ctx::sharded_chain)."""
    if comm.world_size == 1:
        return True
    env = os.environ.get("SK_SHARDED_CHAIN")
    on = (env == "1") if env is not None else comm.world_size <= 2          # (same rule as sk_comm_peer_attach)
    return bool(getattr(comm, "fused", False) and getattr(comm, "mode", "") == "peer" and on)


def _panel_scalars(cfg, a: float, b: float, crit: str, tau: float):
    """Everything the convergence scan of panel (a, b) needs that depends on the panel ends only: the tail fit
    (src/adaptive.jl:168-175) and the target-independent pieces of the truncation bound.  Returns
    (c, d, criteria, message, scan arguments)."""
    if crit == "panel":                                                          # :168
        c = d = float("nan")
    else:
        c, d = estimate_tail_decay(cfg, a, b, d=cfg.tail)
    if (math.isnan(c) or math.isnan(d)) and crit != "panel":                     # :170-175
        msg = "\talgebraic tail estimate failed -- using convergence_criteria = :panel"
        crit = "panel"
    elif crit != "panel":
        msg = f"\talgebraic tail estimate S(w) ≈ {c:.2e} * w^({d:.2f})"
    else:
        msg = None
    return c, d, crit, msg, _scan_args(cfg, b, c, d, tau, crit)


class _NoComm:
    """Single-process stand-in for the scalar reductions of a target-sharded run."""
    world_size = 1

    def max(self, vals: Sequence[float]) -> List[float]:
        return list(vals)

    def min(self, vals: Sequence[float]) -> List[float]:
        return list(vals)

    def sum(self, vals: Sequence[float]) -> List[float]:
        return list(vals)

    def gather(self, vals: Sequence[float]) -> List[List[float]]:
        return [list(vals)]


def _host_rules(cfg: AdaptiveKernelConfig, eng):
    if cfg._rules is None:
        leg = eng.rule_get(0) + eng.rule_get(1)
        jac = (eng.rule_get(2) + eng.rule_get(3)) if cfg.p != 0.0 else leg
        cfg._rules = (leg, jac)
    return cfg._rules


def _subpanel_edges(a: float, b: float, k: int) -> np.ndarray:
    """range(a, b, length=k+1), src/quadrature.jl:56 (twice-precision step: one rounding per element)."""
    al, bl = np.longdouble(a), np.longdouble(b)
    e = (al + np.arange(k + 1, dtype=np.longdouble) * ((bl - al) / np.longdouble(k))).astype(np.float64)
    e[0], e[-1] = a, b
    return e


def _host_strengths(cfg: AdaptiveKernelConfig, eng, a: float, b: float, origin: bool, integrand=None):
    """updatequadbufs! (src/quadrature.jl:49-95) on the host for an arbitrary callable S."""
    m, k = cfg.quadspec
    (ln1, lw1, ln2, lw2), (jn1, jw1, jn2, jw2) = _host_rules(cfg, eng)
    p, f = cfg.p, (integrand if integrand is not None else cfg.f)
    if origin:
        g = f
        pw = lambda no: np.ones_like(no) if p == 0 else np.power(no, p)
    else:                                                                        # quadrature.jl:240-247
        lg = (lambda w: np.log(w)) if cfg.logw else (lambda w: 1.0)
        g = lambda w: (np.ones_like(w) if p == 0 else np.power(w, p)) * lg(w) * f(w)
        pw = lambda no: np.ones_like(no)
    edges = _subpanel_edges(a, b, k)
    out = []
    for mm, ln, lw, jn, jw in ((m, ln1, lw1, jn1, jw1), (2 * m, ln2, lw2, jn2, jw2)):
        no = np.empty(mm * k)
        buf = np.empty(mm * k)
        first = 0
        if origin:                                                               # :61-78
            bm, bp = (edges[1] - edges[0]) / 2, (edges[1] + edges[0]) / 2
            no[:mm] = bm * jn + bp
            buf[:mm] = jw * bm ** (p + 1) * g(no[:mm])
            first = 1
        for i in range(first, k):                                                # :82-92
            bm, bp = (edges[i + 1] - edges[i]) / 2, (edges[i + 1] + edges[i]) / 2
            sl = slice(i * mm, (i + 1) * mm)
            no[sl] = bm * ln + bp
            buf[sl] = lw * bm * pw(no[sl]) * g(no[sl])
        out += [no, buf]
    return out


def fourier_integrate_interval(cfg: AdaptiveKernelConfig, eng, a: float, b: float, k0: float, comm, active: bool,
                               verbose: bool = False, trace: Optional[list] = None, speculate=None,
                               n_act_g: Optional[int] = None, spec_state: Optional[dict] = None, while_enqueued=None,
                               chain_out: Optional[Tuple[int, int]] = None):
    """Scalar control flow of src/quadrature.jl:169-275: LIFO bisection, accept test against
    config.tol*k0 (not the split tolerance), 9:1 tolerance split at the origin.  The per-target work of
    every pass happens inside sk_subinterval / sk_subinterval_accept.

    `while_enqueued` (callable): host work that does not depend on this panel's outcome (the next panel's tail fit);
    it runs between sk_subinterval_begin and sk_subinterval_end of the panel's first sub-interval, i.e. while the
    device integrates."""
    nu, xdiv = 0, 0.0
    if cfg.dim == 1:
        kernel = SK_KERNEL_SIN if cfg.derivative else SK_KERNEL_COS              # :177
    else:
        # (:J, dim/2) or (:J, dim/2-1), :179; only even dims are usable (Int64(kernel[2]), :138).  Where the reference
        # calls FastHankelTransform.jl's nufht (:139-143) the library runs its own O(N) nonuniform Hankel transform
        # (orders 0..3, csrc/sk_hankel.h); small active sets take the direct Bessel summation (:145-160).
        order = cfg.dim / 2 if cfg.derivative else cfg.dim / 2 - 1
        if order != int(order):
            raise ValueError(f"InexactError: Int64({order})")                    # :138
        kernel, nu, xdiv = SK_KERNEL_BESSEL, int(order), cfg.dim / 2 - 1         # :252-254
    stack = [(a, b, cfg.tol)]                                                    # :173
    builtin = is_builtin(cfg.f)
    first = True
    spec_state = spec_state if spec_state is not None else {}
    spec_state["ab"] = False            # an idle rank: did the panel's first collective carry the scan's scalars?
    spec_state["first_accepted"] = None
    while stack:
        _a, _b, _tol = stack.pop()                                               # :183
        # the first interval popped is the whole panel: let the device fuse accept / commit / scan into
        # the interpolation kernel (rolled back by the library if the interval is rejected)
        spec = speculate if first else None
        first = False
        origin = (_a == 0.0 and cfg.p != 0.0)                                    # :185
        if abs(_b - _a) <= 1e-16:                                                # utils.jl:28-36
            raise RuntimeError(f"The sub-interval (a, b) = ({_a}, {_b}) has been split too many times "
                               f"(b - a < 1e-16). Exiting to avoid infinite splitting.")
        if active:
            if origin and cfg.logw:                                              # :186-228, integration by parts
                if cfg.dim not in (1, 2):
                    raise NotImplementedError("singularity derivative not implemented in d > 2")   # :222-223
                f, df = cfg.f, cfg.df
                if builtin and getattr(f, "deriv", 0) == 0 and (df is None or getattr(df, "__self__", None) is f):
                    # a shipped family and its own closed-form dS/dw: both integrands are evaluated on the device
                    no1 = ba1 = bb1 = no2 = ba2 = bb2 = None
                else:
                    if df is None:
                        raise TypeError("logw=true needs df (the derivative of the spectral density)")
                    ga = lambda w: f(w) + w * np.log(w) * df(w)                  # :192, :210
                    gb = lambda w: w * np.log(w) * f(w)                          # :198, :216
                    no1, ba1, no2, ba2 = _host_strengths(cfg, eng, _a, _b, True, integrand=ga)
                    _, bb1, _, bb2 = _host_strengths(cfg, eng, _a, _b, True, integrand=gb)
                i0 = _b ** (cfg.dim / 2 + 1 - cfg.alpha) * math.log(_b) * _f_scalar(f, _b)      # :189
                if cfg.dim == 1:
                    mx = eng.subinterval_logw_host(_a, _b, no1, ba1, bb1, no2, ba2, bb2, cfg.c, cfg.p, i0,
                                                   cfg.dim - cfg.alpha)
                else:                                                            # :204-221: (:J, dim/2-1) and (:J, dim/2)
                    mx = eng.subinterval_logw_host(_a, _b, no1, ba1, bb1, no2, ba2, bb2, cfg.c, cfg.p, i0,
                                                   cfg.dim - cfg.alpha, kernel=SK_KERNEL_BESSEL,
                                                   nu=int(cfg.dim / 2 - 1), xdiv_pow=cfg.dim / 2 - 1)
            elif builtin and spec is not None and while_enqueued is not None and OVERLAP_HOST_WORK and \
                    hasattr(eng, "subinterval_begin"):
                eng.subinterval_begin(_a, _b, cfg.c, cfg.p, kernel, cfg.logw, speculate=spec, nu=nu, xdiv_pow=xdiv)
                try:
                    nxt = while_enqueued()
                    if nxt is not None and _may_chain(comm) and hasattr(eng, "subinterval_chain"):
                        # the next panel's first sub-interval goes in behind this one, guarded on the device: it runs
                        # only if this one is accepted (:260) and converges nothing (sk_subinterval_chain)
                        a2, b2, sargs2 = nxt
                        if eng.subinterval_chain(a2, b2, cfg.c, cfg.p, kernel, cfg.logw, sargs2, cfg.tol * k0, nu=nu,
                                                 xdiv_pow=xdiv) and chain_out is not None:
                            # ... and behind it the final gather, which runs if that panel ends the loop (:149)
                            eng.results_chain_device(chain_out[0], chain_out[1], cfg.tol * k0)
                finally:
                    mx = eng.subinterval_end()
            elif builtin:
                mx = eng.subinterval(_a, _b, cfg.c, cfg.p, kernel, cfg.logw, speculate=spec, nu=nu, xdiv_pow=xdiv)
            else:
                no1, buf1, no2, buf2 = _host_strengths(cfg, eng, _a, _b, origin)
                mx = eng.subinterval_host(_a, _b, no1, buf1, no2, buf2, cfg.c, cfg.p, kernel, cfg.logw, speculate=spec,
                                          nu=nu, xdiv_pow=xdiv)
        else:
            mx = 0.0
        if getattr(comm, "fused", False):
            # the library reduced over the ranks on its stream (sk_comm_init); idle ranks join the collective
            if not active:
                # the active ranks' collective carries the scan's scalars too when they speculate (sk_comm_idle)
                ab = spec is not None and kernel != SK_KERNEL_BESSEL and n_act_g is not None and \
                    2 * cfg.quadsz * n_act_g > 2 ** 18 and n_act_g > 1
                eng.comm_idle(2 if ab else 0)
                if ab:
                    spec_state["ab"] = True
                mx = eng.comm_last()[0]
            mx_g = math.inf if math.isnan(mx) else mx
        else:
            # max over all ranks; NaN travels as +inf (both fail the accept test, quadrature.jl:260)
            mx_g = comm.max([math.inf if math.isnan(mx) else mx])[0]
        accepted = mx_g < cfg.tol * k0                                           # :260
        if spec_state["first_accepted"] is None:
            spec_state["first_accepted"] = bool(accepted)
        if verbose:
            word = "converged" if mx_g / k0 <= _tol else "did not converge"
            print(f"\tsubpanel w ∈ [{_a:.2e}, {_b:.2e}] {word} to tolerance {_tol:.2e} with max error {mx_g / k0:.2e}")
        if trace is not None:
            trace.append({"kind": "subinterval", "a": float(_a), "b": float(_b), "rel_err": float(mx_g / k0),
                          "accepted": bool(accepted)})
        if accepted:
            if active:
                eng.subinterval_accept()                                         # :261-262
        else:                                                                    # :268-270
            tl, tr = (9 * _tol / 10, _tol / 10) if _a == 0 else (_tol / 2, _tol / 2)
            mid = (_a + _b) / 2
            if not (_a < mid < _b):
                # the reference would push the same interval again and loop forever once (a+b)/2 rounds
                # to an end point (its guard only fires for b - a <= 1e-16, src/utils.jl:28-36)
                raise RuntimeError(f"The sub-interval (a, b) = ({_a}, {_b}) cannot be split any further. "
                                   f"Exiting to avoid infinite splitting.")
            stack.append((_a, mid, tl))
            stack.append((mid, _b, tr))


def kernel_values(cfg: AdaptiveKernelConfig, xs, *, k0: Optional[float] = None, param_derivative: bool = False,
                  verbose: bool = False, trace: Optional[list] = None, comm=None, want_errors: bool = True,
                  out_vals=None, out_errs=None, xs_device: Optional[Tuple[int, int]] = None,
                  out_device: Optional[Tuple[int, int]] = None, reuse_targets: bool = False,
                  points=None, pairs=None, async_results: bool = False):
    """`kernel_values(config, xs; k0, param_derivative, verbose)` (src/adaptive.jl:95-108): returns
    (values, errors) in the order of `xs`, duplicates included.

    Optional extras: `trace` (list, receives the panel trace), `comm` (scalar reductions of a
    target-sharded multi-GPU run: every rank passes its own chunk of the distances), `xs_device` /
    `out_device` ((pointer, n) / (vals_ptr, errs_ptr): device-resident input and output, no PCIe),
    `reuse_targets` (the engine already holds exactly these distances from the previous call -- the
    P_sdf + 2 derivative runs of src/derivatives.jl:86-112 all use the same lags -- so the upload and the
    sort/unique are skipped), `points` (+ optional `pairs`): evaluate at lag = ||points[i] - points[j]|| for the
    given index pairs (default: all i < j), computed on the device (src/model.jl:53-68 with NoWarping); `xs` is
    ignored and the results come back in pair order; `async_results` (with pinned `out_vals` / `out_errs`): return as
    soon as the copy of the results is queued on the copy stream -- the next kernel_values call on the same engine
    overlaps with it; `cfg.engine.results_wait()` blocks until the arrays are complete."""
    eng = cfg.engine
    comm = comm or _NoComm()
    if k0 is None:
        k0 = compute_k0(cfg)                                                     # :97
    m, k = cfg.quadspec
    eng.rule_set(m, k, cfg.p)
    if is_builtin(cfg.f):
        eng.sdf_builtin(cfg.f.family, cfg.f.params, cfg.f.deriv)
    # unique + sort + inverse map on the device (adaptive.jl:99, :113-120)
    pre = {}         # panel scalars computed ahead of time, keyed by the exact (a, b, criteria) they were computed for
    if reuse_targets and getattr(eng, "_last_targets", None) is not None:
        info, n_in = eng._last_targets
    elif points is not None:
        info = eng.targets_set_pairs(points, pairs)
        n_in = int(info.n_in)
    else:
        if xs_device is None:
            xs = np.ascontiguousarray(xs, dtype=np.float64)
        n_in = int(xs_device[1]) if xs_device is not None else xs.size
        if OVERLAP_HOST_WORK and hasattr(eng, "targets_begin") and n_in > 0:
            # two halves: while the device sorts, the host prepares the first panel (0, m k / (2 r_max)), whose ends
            # are known as soon as the first pass over the distances has delivered their range
            if xs_device is not None:
                eng.targets_begin_device(*xs_device)
            else:
                eng.targets_begin(xs)
            try:
                _, r_early = eng.targets_early_range()
                if r_early > 0 and comm.world_size == 1 or (r_early > 0 and getattr(comm, "fused", False)):
                    b1 = 0.0 + cfg.quadsz / (2 * r_early)
                    ps = _panel_scalars(cfg, 0.0, b1, cfg.convergence_criteria, cfg.tol * abs(k0) / 2)
                    pre[(0.0, b1, cfg.convergence_criteria)] = ps
                    if _may_chain(comm) and cfg.dim == 1 and is_builtin(cfg.f) and ps[2] == cfg.convergence_criteria \
                            and hasattr(eng, "first_panel_early"):
                        # the first panel goes in behind the sort (sk_first_panel_early): its kernel takes the number
                        # of unique distances from the sort's device-side summary
                        eng.first_panel_early(0.0, b1, cfg.c, cfg.p, SK_KERNEL_SIN if cfg.derivative else SK_KERNEL_COS,
                                              cfg.logw, ps[4])
            finally:
                info = eng.targets_end()
        elif xs_device is not None:
            info = eng.targets_set_device(*xs_device)
        else:
            info = eng.targets_set(xs)
    try:
        eng._last_targets = (info, n_in)
    except AttributeError:
        pass
    if verbose:
        print(f"Reducing {info.n_in} to {info.n_unique} unique lags for evaluation...")
    eng.run_begin()
    n = int(info.n_unique)
    ix1 = 1
    if info.has_zero:                                                            # :133-146
        ix1 = 2
        if cfg.derivative:
            eng.zero_lag_set(0.0)
        elif param_derivative:
            eng.zero_lag_set(compute_k0(cfg))
        else:
            eng.zero_lag_set(k0)
    hi = n                                                                       # :123
    quadm = cfg.quadsz
    crit = cfg.convergence_criteria
    a = b = 0.0
    # global distance range over all ranks (scalars only)
    r_hi_local = info.r_max if (n >= ix1) else 0.0
    # one gather of (smallest positive distance, largest distance, active count) per rank
    fused = getattr(comm, "fused", False)
    rmin_local = info.r_min_pos if info.r_min_pos > 0 else math.inf
    # (+-inf would turn into NaN in a one-hot sum: a rank without positive distances sends the largest double)
    summ = eng.comm_summary() if (fused and not reuse_targets and getattr(comm, "mode", "") == "peer"
                                  and hasattr(eng, "comm_summary")) else (False,)
    if summ[0]:
        # the three numbers went round behind the sort's summary kernel (sk_comm_summary): no collective of their own
        r_lo_g = summ[1] if summ[1] > 0 else 1.7976931348623157e308
        r_hi_g, n_act_g = summ[2], int(summ[3])
    else:
        g0 = comm.gather([rmin_local if math.isfinite(rmin_local) else 1.7976931348623157e308, r_hi_local,
                          float(max(n - ix1 + 1, 0))])
        r_lo_g = min(v[0] for v in g0)
        r_hi_g = max(v[1] for v in g0)
        n_act_g = int(sum(v[2] for v in g0))
    m2 = 2 * cfg.quadsz
    ipanel = 0
    tau = cfg.tol * abs(k0) / 2                                                  # :191
    while r_hi_g > 0:                                                            # :149 (hi > 0 && xs[hi] > 0)
        a, b = b, b + quadm / (2 * r_hi_g)                                       # :152
        active = hi >= ix1
        if verbose:
            print(f"\nintegrating panel w ∈ [{a:.2e}, {b:.2e}] (length {b - a:.2e}) to resolve {hi} points "
                  f"x ≤ {r_hi_g:.2e} ")
        if active:
            eng.panel_begin(ix1, hi)
            if comm.world_size > 1:
                eng.panel_set_range(r_lo_g, r_hi_g, n_act_g)
        # The tail fit depends on (a, b) only, so it is evaluated BEFORE the panel is integrated (the
        # reference does it after, src/adaptive.jl:168): the scan arguments can then ride along with the
        # panel's first sub-interval.  It is usually there already: computed while the device was sorting (first
        # panel) or integrating the previous panel.
        key = (a, b, crit)
        pre_hit = key in pre
        c, d, crit, crit_msg, sargs = pre.pop(key) if pre_hit else _panel_scalars(cfg, a, b, crit, tau)
        pre.clear()

        def ahead(a2=b, r=r_hi_g, crit2=crit):
            # the next panel if this one converges nothing (the usual outcome of a run's first panels): same r_hi
            b2 = a2 + quadm / (2 * r)
            if math.isfinite(b2) and b2 > a2:
                ps = _panel_scalars(cfg, a2, b2, crit2, tau)
                pre[(a2, b2, crit2)] = ps
                if ps[2] == crit2:              # (a failed tail fit changes the criteria: the host decides that later)
                    return a2, b2, ps[4]
            return None

        spec_state = {}
        fourier_integrate_interval(cfg, eng, a, b, abs(k0), comm, active, verbose=verbose, trace=trace,
                                   speculate=sargs, n_act_g=n_act_g, spec_state=spec_state,
                                   while_enqueued=ahead if (ipanel == 0 or pre_hit) else None,
                                   # (the gather is chained in single-GPU runs only: behind a sharded panel's exchange
                                   #  kernel it measured 28 us SLOWER than launching it from the host)
                                   chain_out=out_device if (out_device is not None and not async_results
                                                            and comm.world_size == 1) else None)             # :157-159
        if active:
            eng.panel_commit()                                                   # :163-164
        if verbose and crit_msg:
            print(crit_msg)
        hi_before = hi
        if active:
            new_hi, r_stop = eng.converge_scan(sargs)                            # :183-198
        else:
            new_hi, r_stop = hi, 0.0
        # one gather per panel: every rank's stopping distance and the number of targets it keeps active if
        # the walk stopped at ITS OWN stopping distance (a lower bound of what it keeps for the global one)
        if fused:
            # (no collective when the scan's scalars travelled with the panel's accepted first sub-interval)
            if not active and not (spec_state.get("ab") and spec_state.get("first_accepted")):
                eng.comm_idle(1)
            _, r_g, n_lb_g = eng.comm_last()
            g1 = [[r_g, float(n_lb_g)]]
        else:
            g1 = comm.gather([r_stop, float(max(new_hi - ix1 + 1, 0))])
        r_stop_g = max(v[0] for v in g1)
        if comm.world_size > 1 and active and r_stop_g > r_stop:
            new_hi = eng.target_upper_index(r_stop_g)
        if active:
            eng.converge_apply(sargs, new_hi)                                    # :194
        hi = new_hi
        r_hi_g = r_stop_g
        if comm.world_size > 1 and r_hi_g > 0:
            # global active count of the next panel, needed only for the NUFFT-vs-direct cutoff
            # (src/quadrature.jl:105): the lower bounds usually decide it; otherwise one exact sum
            n_lb = int(sum(v[1] for v in g1))
            if m2 * n_lb > 2 ** 18 and n_lb > 1:
                n_act_g = max(n_lb, hi - ix1 + 1)
            else:
                n_act_g = int(comm.sum([float(max(hi - ix1 + 1, 0))])[0])
        if trace is not None:
            trace.append({"kind": "panel", "index": ipanel, "a": float(a), "b": float(b), "hi_before": int(hi_before),
                          "hi_after": int(hi), "ix1": int(ix1), "c": float(c), "d": float(d), "criteria": crit})
        ipanel += 1
    if out_device is not None:
        eng.results_get_device(*out_device)
        return None, None
    if async_results:
        if out_vals is None:
            raise ValueError("async_results needs out_vals (pinned host memory)")
        eng.results_get_async(out_vals, out_errs if want_errors else None)
        return out_vals, (out_errs if want_errors else None)
    return eng.results_get(n_in, want_errors=want_errors, out_vals=out_vals, out_errs=out_errs)   # :105-107
