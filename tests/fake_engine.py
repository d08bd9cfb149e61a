"""
TEST INFRASTRUCTURE: an object with the interface of `_capi.Session` whose per-target work is done by
the oracle's direct sums (numpy + oracle/libsk_oracle.so).  It lets the CPU suite exercise the product's
HOST logic -- the adaptive driver in spectralkernels.jl_b200/adaptive.py and the scalar reductions of a
target-sharded run in sharded.py -- with world_size-2 gloo groups, where no GPU exists.  It is never
used by the product.
"""
import types

import numpy as np

import sk_oracle as so


class FakeEngine:
    def __init__(self):
        self.m = self.k = None
        self.calls = []

    # rules -------------------------------------------------------------------------------------------
    def rule_set(self, m, k, p, leg=None, jac=None):
        self.m, self.k, self.p = m, k, p
        self._leg = so.gauss_rule(m) + so.gauss_rule(2 * m)
        self._jac = (so.gauss_rule(m, p) + so.gauss_rule(2 * m, p)) if p != 0.0 else self._leg

    def rule_get(self, which):
        src = self._leg if which < 2 else self._jac
        w = which % 2
        return src[2 * w], src[2 * w + 1]

    def sdf_builtin(self, family, params, deriv):
        raise AssertionError("the fake engine only takes host-evaluated strengths")

    # targets -----------------------------------------------------------------------------------------
    def targets_set(self, xs):
        xs = np.asarray(xs, dtype=np.float64)
        self.uxs, self.inv = np.unique(xs, return_inverse=True)
        n = self.uxs.size
        pos = self.uxs[self.uxs > 0]
        self.n_in = xs.size
        return types.SimpleNamespace(n_in=xs.size, n_unique=n, has_zero=int(self.uxs[0] == 0),
                                     r_min_pos=float(pos[0]) if pos.size else 0.0, r_max=float(self.uxs[-1]))

    # the two halves of targets_set (sk_targets_begin / _early_range / _end): the host prepares the first panel in between
    def targets_begin(self, xs):
        self._pending = np.asarray(xs, dtype=np.float64)
        self.calls.append(("targets_begin",))

    def targets_early_range(self):
        pos = self._pending[self._pending > 0]
        self.calls.append(("targets_early_range",))
        return (float(pos.min()), float(pos.max())) if pos.size else (0.0, 0.0)

    def targets_end(self):
        self.calls.append(("targets_end",))
        return self.targets_set(self._pending)

    def run_begin(self):
        n = self.uxs.size
        self.ks, self.errs = np.zeros(n), np.zeros(n)

    def zero_lag_set(self, v):
        if self.uxs[0] == 0:
            self.ks[0], self.errs[0] = v, np.nan

    def panel_begin(self, ix1, hi):
        self.lo, self.hi = ix1 - 1, hi
        self.I = np.zeros(hi - ix1 + 1)
        self.err = np.zeros(hi - ix1 + 1)
        self.range = (self.uxs[self.lo], self.uxs[hi - 1])
        return self.range

    def panel_set_range(self, r_lo, r_hi, n_active_global=0):
        assert r_lo <= self.range[0] and r_hi >= self.range[1]
        self.calls.append(("range", r_lo, r_hi))

    def subinterval_host(self, a, b, no1, buf1, no2, buf2, cmul, p, kernel, logw, speculate=None, nu=0, xdiv_pow=0.0):
        x = self.uxs[self.lo:self.hi]
        if kernel == 2:
            i1 = so.direct_bessel(nu, no1, buf1, x) * cmul
            i2 = so.direct_bessel(nu, no2, buf2, x) * cmul
            if xdiv_pow != 0.0:
                i1, i2 = i1 / x ** xdiv_pow, i2 / x ** xdiv_pow
            self._stage = (i2, np.abs(i2 - i1))
            return float(np.max(self._stage[1]))
        f1, f2 = so.direct_cis(no1, buf1, x), so.direct_cis(no2, buf2, x)
        i1 = (f1.imag if kernel == 1 else f1.real) * cmul
        i2 = (f2.imag if kernel == 1 else f2.real) * cmul
        self._stage = (i2, np.abs(i2 - i1))
        return float(np.max(self._stage[1]))

    def subinterval_logw_host(self, a, b, no1, bufa1, bufb1, no2, bufa2, bufb2, cmul, p, i0_coef, denom):
        from scipy import special
        x = self.uxs[self.lo:self.hi]
        i0 = i0_coef * special.jv(-0.5, 2 * np.pi * b * x)
        out = []
        for no, ba, bb in ((no1, bufa1, bufb1), (no2, bufa2, bufb2)):
            fa, fb = so.direct_cis(no, ba, x), so.direct_cis(no, bb, x)
            out.append(((i0 - fa.real) + 2 * np.pi * x * fb.imag) / denom * cmul)
        self._stage = (out[1], np.abs(out[1] - out[0]))
        return float(np.max(self._stage[1]))

    def subinterval(self, *a, **k):
        raise AssertionError("built-in S needs the CUDA engine")

    def subinterval_accept(self):
        self.I += self._stage[0]
        self.err += self._stage[1]

    def panel_commit(self):
        self.ks[self.lo:self.hi] += self.I
        self.errs[self.lo:self.hi] += self.err

    def _trunc(self, a, x):
        if a.criteria == 0:
            return np.zeros_like(x)
        with np.errstate(all="ignore"):
            return np.minimum(a.trunc_a, a.trunc_num / (2 * np.pi * x ** a.xpow))

    def converge_scan(self, a):
        x = self.uxs[self.lo:self.hi]
        te = self._trunc(a, x)
        c1 = (a.criteria == 0) | (te < a.tau)
        c2 = (a.criteria == 1) | (np.abs(self.I) < a.tau)
        bad = np.nonzero(~(c1 & c2))[0]
        if bad.size == 0:
            return self.lo, 0.0
        top = self.lo + bad[-1]
        return int(top + 1), float(self.uxs[top])

    def target_upper_index(self, r):
        return int(np.searchsorted(self.uxs, r, side="right"))

    def converge_apply(self, a, new_hi):
        x = self.uxs[new_hi:self.hi]
        self.errs[new_hi:self.hi] += 2 * self._trunc(a, x)

    def results_get(self, n_in, want_errors=True, out_vals=None, out_errs=None):
        return self.ks[self.inv], (self.errs[self.inv] if want_errors else None)
