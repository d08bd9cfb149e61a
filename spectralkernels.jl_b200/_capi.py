"""
ctypes binding of the C ABI in include/spectralkernels_b200.h (libsk_b200.so).

This is the same boundary the Julia host binds with `ccall`; nothing here
computes anything.  There is no CPU fallback: if the library is missing, cannot be loaded, or no
CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsk_b200.so")

SK_OK = 0
SK_KERNEL_COS, SK_KERNEL_SIN, SK_KERNEL_BESSEL = 0, 1, 2
SK_CRIT = {"panel": 0, "tails": 1, "both": 2}
SK_SDF_HOST, SK_SDF_MATERN, SK_SDF_EXPONENTIAL = 0, 1, 2
SK_ERR_NAN, SK_ERR_SPLIT, SK_ERR_UNSUPPORTED = -4, -5, -8


class SkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[sk_b200 {code}] {msg}")
        self.code = code


class TargetInfo(ctypes.Structure):
    _fields_ = [("n_in", c_int64), ("n_unique", c_int64), ("has_zero", c_int32), ("_pad", c_int32),
                ("r_min_pos", c_double), ("r_max", c_double)]


class ScanArgs(ctypes.Structure):
    _fields_ = [("trunc_a", c_double), ("trunc_num", c_double), ("xpow", c_double), ("tau", c_double),
                ("criteria", c_int32), ("_pad", c_int32)]


class SubintervalOpts(ctypes.Structure):
    _fields_ = [("cmul", c_double), ("p", c_double), ("kernel", c_int32), ("logw", c_int32),
                ("nu", c_int32), ("_pad", c_int32), ("xdiv_pow", c_double), ("speculate", POINTER(ScanArgs))]


class Stats(ctypes.Structure):
    _fields_ = [("n_subintervals", c_int64), ("n_accepted", c_int64), ("n_panels", c_int64), ("units", c_int64),
                ("n_fast", c_int64), ("n_direct", c_int64), ("kernel_launches", c_int64), ("last_nf", c_int64),
                ("last_nf2", c_int64), ("n_speculated", c_int64), ("n_spec_rollbacks", c_int64), ("interp_ms", c_double), ("source_ms", c_double),
                ("timing_enabled", c_int32), ("sort_two_level", c_int32), ("n_hankel", c_int64), ("sort_ms", c_double),
                ("gather_ms", c_double), ("n_prefetch_issued", c_int64), ("n_prefetch_hits", c_int64), ("n_chained", c_int64), ("launches_total", c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


_dp = POINTER(c_double)

# name -> (restype, argtypes); exactly the symbols include/spectralkernels_b200.h declares
SIGNATURES = {
    "sk_abi_version": (c_int, []),
    "sk_error_string": (c_char_p, [c_int]),
    "sk_last_error": (c_char_p, [c_void_p]),
    "sk_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "sk_ctx_destroy": (c_int, [c_void_p]),
    "sk_ctx_set_timing": (c_int, [c_void_p, c_int]),
    "sk_ctx_set_nufft_eps": (c_int, [c_void_p, c_double]),
    "sk_ctx_set_interp_mode": (c_int, [c_void_p, c_int]),
    "sk_ctx_set_hankel_mode": (c_int, [c_void_p, c_int]),
    "sk_ctx_synchronize": (c_int, [c_void_p]),
    "sk_ctx_stream": (c_int, [c_void_p, POINTER(c_void_p)]),
    "sk_timer_begin": (c_int, [c_void_p]),
    "sk_timer_end": (c_int, [c_void_p, _dp]),
    "sk_fp64_peak": (c_int, [c_void_p, _dp, _dp]),
    "sk_comm_unique_id": (c_int, [c_void_p]),
    "sk_comm_init": (c_int, [c_void_p, c_void_p, c_int32, c_int32]),
    "sk_comm_destroy": (c_int, [c_void_p]),
    "sk_comm_allreduce": (c_int, [c_void_p, _dp, c_int32, c_int32]),
    "sk_comm_idle": (c_int, [c_void_p, c_int32]),
    "sk_comm_last": (c_int, [c_void_p, _dp, _dp, POINTER(c_int64)]),
    "sk_targets_begin": (c_int, [c_void_p, _dp, c_int64]),
    "sk_targets_begin_device": (c_int, [c_void_p, c_void_p, c_int64]),
    "sk_targets_early_range": (c_int, [c_void_p, _dp, _dp]),
    "sk_targets_end": (c_int, [c_void_p, c_void_p]),
    "sk_first_panel_early": (c_int, [c_void_p, c_double, c_double, c_void_p, POINTER(c_int32)]),
    "sk_subinterval_begin": (c_int, [c_void_p, c_double, c_double, c_void_p]),
    "sk_subinterval_end": (c_int, [c_void_p, _dp]),
    "sk_subinterval_chain": (c_int, [c_void_p, c_double, c_double, c_void_p, c_double, POINTER(c_int32)]),
    "sk_results_chain_device": (c_int, [c_void_p, c_void_p, c_void_p, c_double, POINTER(c_int32)]),
    "sk_comm_peer_export": (c_int, [c_void_p, c_void_p]),
    "sk_comm_peer_attach": (c_int, [c_void_p, c_void_p, c_int32, c_int32]),
    "sk_comm_allgather": (c_int, [c_void_p, _dp, c_int32, _dp]),
    "sk_comm_summary": (c_int, [c_void_p, POINTER(c_int32), _dp, _dp, POINTER(c_int64)]),
    "sk_comm_peer_selftest": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "sk_host_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "sk_host_free": (c_int, [c_void_p]),
    "sk_nufft1d3": (c_int, [c_void_p, c_int64, _dp, _dp, c_int64, _dp, _dp, c_double]),
    "sk_rule_set": (c_int, [c_void_p, c_int32, c_int32, c_double] + [_dp] * 8),
    "sk_rule_get": (c_int, [c_void_p, c_int32, _dp, _dp]),
    "sk_sdf_builtin": (c_int, [c_void_p, c_int32, _dp, c_int32, c_int32]),
    "sk_targets_set": (c_int, [c_void_p, c_void_p, c_int64, POINTER(TargetInfo)]),
    "sk_targets_set_device": (c_int, [c_void_p, c_void_p, c_int64, POINTER(TargetInfo)]),
    "sk_targets_set_pairs": (c_int, [c_void_p, _dp, c_int64, c_int32, POINTER(c_int64), c_int64, POINTER(TargetInfo)]),
    "sk_targets_scale": (c_int, [c_void_p, c_double, c_void_p]),
    "sk_target_value": (c_int, [c_void_p, c_int64, _dp]),
    "sk_run_begin": (c_int, [c_void_p]),
    "sk_zero_lag_set": (c_int, [c_void_p, c_double]),
    "sk_panel_begin": (c_int, [c_void_p, c_int64, c_int64, _dp, _dp]),
    "sk_panel_set_range": (c_int, [c_void_p, c_double, c_double, c_int64]),
    "sk_subinterval": (c_int, [c_void_p, c_double, c_double, POINTER(SubintervalOpts), _dp]),
    "sk_subinterval_host": (c_int, [c_void_p, c_double, c_double, _dp, _dp, _dp, _dp, POINTER(SubintervalOpts), _dp]),
    "sk_subinterval_logw_host": (c_int, [c_void_p, c_double, c_double, _dp, _dp, _dp, _dp, _dp, _dp,
                                         POINTER(SubintervalOpts), c_double, c_double, _dp]),
    "sk_sources_get": (c_int, [c_void_p, c_int32, _dp, _dp]),
    "sk_subinterval_accept": (c_int, [c_void_p]),
    "sk_panel_commit": (c_int, [c_void_p]),
    "sk_converge_scan": (c_int, [c_void_p, POINTER(ScanArgs), POINTER(c_int64), _dp]),
    "sk_converge_apply": (c_int, [c_void_p, POINTER(ScanArgs), c_int64]),
    "sk_target_upper_index": (c_int, [c_void_p, c_double, POINTER(c_int64)]),
    "sk_results_get": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sk_results_get_async": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sk_results_wait": (c_int, [c_void_p]),
    "sk_results_get_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sk_stats_get": (c_int, [c_void_p, POINTER(Stats)]),
    "sk_group_create": (c_int, [POINTER(c_int32), c_int32, POINTER(c_void_p)]),
    "sk_group_destroy": (c_int, [c_void_p]),
    "sk_group_size": (c_int, [c_void_p]),
    "sk_group_ctx": (c_int, [c_void_p, c_int32, POINTER(c_void_p)]),
    "sk_group_last_error": (c_char_p, [c_void_p]),
    "sk_group_set_timing": (c_int, [c_void_p, c_int]),
    "sk_group_set_nufft_eps": (c_int, [c_void_p, c_double]),
    "sk_group_synchronize": (c_int, [c_void_p]),
    "sk_group_rule_set": (c_int, [c_void_p, c_int32, c_int32, c_double] + [_dp] * 8),
    "sk_group_rule_get": (c_int, [c_void_p, c_int32, _dp, _dp]),
    "sk_group_sdf_builtin": (c_int, [c_void_p, c_int32, _dp, c_int32, c_int32]),
    "sk_group_targets_set": (c_int, [c_void_p, c_void_p, c_int64, POINTER(TargetInfo)]),
    "sk_group_run_begin": (c_int, [c_void_p]),
    "sk_group_zero_lag_set": (c_int, [c_void_p, c_double]),
    "sk_group_panel_begin": (c_int, [c_void_p, c_int64, c_int64, _dp, _dp]),
    "sk_group_subinterval": (c_int, [c_void_p, c_double, c_double, POINTER(SubintervalOpts), _dp]),
    "sk_group_subinterval_host": (c_int, [c_void_p, c_double, c_double, _dp, _dp, _dp, _dp, POINTER(SubintervalOpts), _dp]),
    "sk_group_subinterval_accept": (c_int, [c_void_p]),
    "sk_group_panel_commit": (c_int, [c_void_p]),
    "sk_group_converge_scan": (c_int, [c_void_p, POINTER(ScanArgs), POINTER(c_int64), _dp]),
    "sk_group_converge_apply": (c_int, [c_void_p, POINTER(ScanArgs), c_int64]),
    "sk_group_results_get": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sk_group_stats_get": (c_int, [c_void_p, POINTER(Stats)]),
    "sk_host_gauss_rule": (c_int, [c_int32, c_double, _dp, _dp]),
    "sk_host_es_plan": (c_int, [c_int32, _dp, POINTER(c_int32), _dp, POINTER(c_int32), _dp, _dp]),
}

_lib = None


def load():
    """dlopen libsk_b200.so (raises if it has not been built: run `python spectralkernels.jl_b200/build.py`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SkError(-100, f"{LIB_PATH} is missing: build it with `python spectralkernels.jl_b200/build.py` "
                                f"(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class PinnedArray:
    """A float64 numpy view over cudaMallocHost memory (sk_host_alloc)."""

    def __init__(self, n: int):
        self._ptr = c_void_p()
        rc = load().sk_host_alloc(max(int(n), 1) * 8, byref(self._ptr))
        if rc != SK_OK:
            raise SkError(rc, "sk_host_alloc failed")
        buf = (c_double * max(int(n), 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=int(n))

    def free(self):
        if self._ptr is not None and self._ptr.value:
            load().sk_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Session:
    """Owns one sk_ctx (one GPU, one stream).  Method names follow the C entry points."""

    def __init__(self, device: int = 0, timing: bool = False):
        self._L = load()
        self._h = c_void_p()
        rc = self._L.sk_ctx_create(int(device), byref(self._h))
        if rc != SK_OK:
            raise SkError(rc, "sk_ctx_create failed (is a CUDA device visible? there is no CPU fallback)")
        self.device = int(device)
        if timing:
            self._ck(self._L.sk_ctx_set_timing(self._h, 1))
        self.rule_key = None
        self.sdf_key = None

    # -- plumbing ------------------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != SK_OK:
            detail = self._L.sk_last_error(self._h)
            base = self._L.sk_error_string(rc)
            raise SkError(rc, (detail or base or b"").decode())

    def close(self):
        if self._h is not None and self._h.value:
            self._L.sk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._ck(self._L.sk_ctx_synchronize(self._h))

    def stream_handle(self) -> int:
        s = c_void_p()
        self._ck(self._L.sk_ctx_stream(self._h, byref(s)))
        return s.value or 0

    def timer_begin(self):
        self._ck(self._L.sk_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = c_double()
        self._ck(self._L.sk_timer_end(self._h, byref(ms)))
        return ms.value

    def fp64_peak(self):
        tf, ms = c_double(), c_double()
        self._ck(self._L.sk_fp64_peak(self._h, byref(tf), byref(ms)))
        return tf.value, ms.value

    def set_timing(self, on: bool):
        self._ck(self._L.sk_ctx_set_timing(self._h, 1 if on else 0))

    # -- in-library communicator (NCCL on the context's stream) -------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = load().sk_comm_unique_id(buf)
        if rc != SK_OK:
            raise SkError(rc, "sk_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, nranks: int):
        buf = ctypes.create_string_buffer(bytes(uid), 128)
        self._ck(self._L.sk_comm_init(self._h, buf, int(rank), int(nranks)))
        self.comm_size = int(nranks)

    def comm_destroy(self):
        self._ck(self._L.sk_comm_destroy(self._h))
        self.comm_size = 1

    def comm_allreduce(self, vals, op: int):
        a = _f64(list(vals)).copy()
        self._ck(self._L.sk_comm_allreduce(self._h, _p(a), a.size, int(op)))
        return a.tolist()

    def comm_idle(self, which: int):
        self._ck(self._L.sk_comm_idle(self._h, int(which)))

    # -- the same collectives over peer-mapped mailboxes (NVLink / NVSwitch; k_peer_exchange) -------------
    def comm_peer_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._ck(self._L.sk_comm_peer_export(self._h, buf))
        return buf.raw

    def comm_peer_attach(self, handles, rank: int, nranks: int):
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != 64 * int(nranks):
            raise ValueError("one 64-byte handle per rank")
        buf = ctypes.create_string_buffer(blob, len(blob))
        self._ck(self._L.sk_comm_peer_attach(self._h, buf, int(rank), int(nranks)))
        self.comm_size = int(nranks)

    def comm_allgather(self, vals, nranks: int):
        a = _f64(list(vals)).copy()
        out = np.empty(a.size * int(nranks), dtype=np.float64)
        self._ck(self._L.sk_comm_allgather(self._h, _p(a), a.size, _p(out)))
        return out.reshape(int(nranks), a.size).tolist()

    def comm_summary(self):
        """(valid, r_lo, r_hi, n_active) over all ranks, delivered behind the last targets_set* (peer mailboxes)"""
        ok, lo, hi, n = c_int32(0), c_double(), c_double(), c_int64()
        self._ck(self._L.sk_comm_summary(self._h, byref(ok), byref(lo), byref(hi), byref(n)))
        return bool(ok.value), lo.value, hi.value, n.value

    def comm_peer_selftest(self, maxbits, rbits, top, lo: int, rounds: int = 3, skip_rank: int = -1):
        """the exchange protocol with len(maxbits) ranks emulated on this one device (test hook)"""
        n = len(maxbits)
        mb = np.ascontiguousarray(maxbits, dtype=np.uint64)
        rb = np.ascontiguousarray(rbits, dtype=np.uint64)
        tp = np.ascontiguousarray(top, dtype=np.int64)
        out = np.zeros(5 * n, dtype=np.uint64)
        self._ck(self._L.sk_comm_peer_selftest(self._h, n, int(rounds), mb.ctypes.data, rb.ctypes.data, tp.ctypes.data, int(lo),
                                               int(skip_rank), out.ctypes.data))
        return out.reshape(n, 5)

    def comm_last(self):
        mx, r, n = c_double(), c_double(), c_int64()
        self._ck(self._L.sk_comm_last(self._h, byref(mx), byref(r), byref(n)))
        return mx.value, r.value, n.value

    def set_hankel_mode(self, mode: int):
        """dim >= 2: 0 auto, 1 always the direct Bessel summation, 2 always the O(N) nonuniform Hankel transform."""
        self._ck(self._L.sk_ctx_set_hankel_mode(self._h, int(mode)))

    def set_interp_mode(self, mode: int):
        self._ck(self._L.sk_ctx_set_interp_mode(self._h, int(mode)))

    def set_nufft_eps(self, eps: float):
        self._ck(self._L.sk_ctx_set_nufft_eps(self._h, float(eps)))

    # -- Level 0 ---------------------------------------------------------------------------------------
    def nufft1d3(self, w, s, x, eps: float = 0.0) -> np.ndarray:
        """finufft1d3(w, s, x) of src/utils.jl:10: f_j = sum_k s_k exp(+i 2 pi x_j w_k)."""
        w, x = _f64(w), _f64(x)
        sc = np.ascontiguousarray(np.asarray(s, dtype=np.complex128))
        out = np.empty(x.size, dtype=np.complex128)
        self._ck(self._L.sk_nufft1d3(self._h, w.size, _p(w), sc.view(np.float64).ctypes.data_as(_dp), x.size, _p(x),
                                     out.view(np.float64).ctypes.data_as(_dp), float(eps)))
        return out

    # -- Level 1 ---------------------------------------------------------------------------------------
    def rule_set(self, m: int, k: int, p: float, leg=None, jac=None):
        """leg/jac: optional (no1, wt1, no2, wt2) tuples of host arrays; None = generated by the library."""
        key = (int(m), int(k), float(p), leg is None, jac is None)
        if leg is None and jac is None and key == self.rule_key:
            return
        args = []
        keep = []
        for rule in (leg, jac):
            if rule is None:
                args += [None] * 4
            else:
                arrs = [_f64(a) for a in rule]
                keep += arrs
                args += [_p(a) for a in arrs]
        self._ck(self._L.sk_rule_set(self._h, int(m), int(k), float(p), *args))
        self.rule_key = key
        self.m, self.k, self.p = int(m), int(k), float(p)

    def rule_get(self, which: int):
        n = self.m * (2 if which in (1, 3) else 1)
        no, wt = np.empty(n), np.empty(n)
        self._ck(self._L.sk_rule_get(self._h, int(which), _p(no), _p(wt)))
        return no, wt

    def sdf_builtin(self, family: int, params, deriv_index: int = 0):
        key = (int(family), tuple(float(v) for v in params), int(deriv_index))
        if key == self.sdf_key:                    # unchanged since the last call on this context
            return
        pr = _f64(params)
        self._ck(self._L.sk_sdf_builtin(self._h, int(family), _p(pr) if pr.size else None, pr.size, int(deriv_index)))
        self.sdf_key = key

    def targets_set(self, xs: np.ndarray) -> TargetInfo:
        xs = _f64(xs)
        info = TargetInfo()
        self._ck(self._L.sk_targets_set(self._h, xs.ctypes.data, xs.size, byref(info)))
        self._last_targets = (info, int(info.n_in))          # what kernel_values(reuse_targets=True) picks up
        return info

    def targets_set_device(self, dev_ptr: int, n: int) -> TargetInfo:
        info = TargetInfo()
        self._ck(self._L.sk_targets_set_device(self._h, c_void_p(dev_ptr), int(n), byref(info)))
        self._last_targets = (info, int(info.n_in))
        return info

    def targets_set_pairs(self, pts, pairs=None) -> TargetInfo:
        """pts: (npts, dim) array; pairs: (npairs, 2) int64 0-based index pairs, or None for all i < j."""
        pts = np.ascontiguousarray(np.atleast_2d(np.asarray(pts, dtype=np.float64).T).T if np.ndim(pts) == 1
                                   else np.asarray(pts, dtype=np.float64))
        if pts.ndim == 1:
            pts = pts.reshape(-1, 1)
        info = TargetInfo()
        if pairs is None:
            self._ck(self._L.sk_targets_set_pairs(self._h, _p(pts), pts.shape[0], pts.shape[1], None, 0, byref(info)))
        else:
            pr = np.ascontiguousarray(pairs, dtype=np.int64)
            self._ck(self._L.sk_targets_set_pairs(self._h, _p(pts), pts.shape[0], pts.shape[1],
                                                  pr.ctypes.data_as(POINTER(c_int64)), pr.shape[0], byref(info)))
        self._last_targets = (info, int(info.n_in))
        return info

    # sk_targets_set[_device] in two halves: the host's scalar work for the first panel overlaps the sort
    def targets_begin(self, xs: np.ndarray):
        self._begin_keep = xs                      # (the upload is asynchronous: keep the array alive until targets_end)
        self._ck(self._L.sk_targets_begin(self._h, _p(xs), xs.size))

    def targets_begin_device(self, ptr: int, n: int):
        self._ck(self._L.sk_targets_begin_device(self._h, c_void_p(ptr), int(n)))

    def targets_early_range(self):
        lo, hi = c_double(), c_double()
        self._ck(self._L.sk_targets_early_range(self._h, byref(lo), byref(hi)))
        return lo.value, hi.value

    def first_panel_early(self, a: float, b: float, cmul: float, p: float, kernel: int, logw: bool, speculate) -> bool:
        """between targets_begin* and targets_end: enqueue the first panel's first sub-interval behind the sort"""
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1 if logw else 0, 0, 0, 0.0, ctypes.pointer(speculate))
        done = c_int32(0)
        self._ck(self._L.sk_first_panel_early(self._h, float(a), float(b), byref(o), byref(done)))
        return bool(done.value)

    def targets_end(self) -> TargetInfo:
        info = TargetInfo()
        try:
            self._ck(self._L.sk_targets_end(self._h, byref(info)))
        finally:
            self._begin_keep = None
        return info

    def subinterval_begin(self, a: float, b: float, cmul: float, p: float, kernel: int, logw: bool, speculate=None,
                          nu: int = 0, xdiv_pow: float = 0.0):
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1 if logw else 0, int(nu), 0, float(xdiv_pow),
                            ctypes.pointer(speculate) if speculate is not None else None)
        self._sub_keep = (o, speculate)
        self._ck(self._L.sk_subinterval_begin(self._h, float(a), float(b), byref(o)))

    def subinterval_chain(self, a2: float, b2: float, cmul: float, p: float, kernel: int, logw: bool, speculate,
                          accept_below: float, nu: int = 0, xdiv_pow: float = 0.0) -> bool:
        """enqueue the next panel's first sub-interval behind the one in flight, guarded on the device"""
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1 if logw else 0, int(nu), 0, float(xdiv_pow),
                            ctypes.pointer(speculate))
        done = c_int32(0)
        self._ck(self._L.sk_subinterval_chain(self._h, float(a2), float(b2), byref(o), float(accept_below), byref(done)))
        return bool(done.value)

    def results_chain_device(self, vals_ptr: int, errs_ptr: int, accept_below: float) -> bool:
        """enqueue the final gather behind the chained panel, guarded on the device (it runs if that panel is the last)"""
        done = c_int32(0)
        self._ck(self._L.sk_results_chain_device(self._h, c_void_p(vals_ptr), c_void_p(errs_ptr) if errs_ptr else None,
                                                 float(accept_below), byref(done)))
        return bool(done.value)

    def subinterval_end(self) -> float:
        out = c_double()
        try:
            self._ck(self._L.sk_subinterval_end(self._h, byref(out)))
        finally:
            self._sub_keep = None
        return out.value

    def targets_scale(self, factor: float) -> "TargetInfo":
        """Lags under the linear warping x -> x * factor (range parameter): re-uses the sort of the last targets_set*."""
        info = TargetInfo()
        self._ck(self._L.sk_targets_scale(self._h, float(factor), byref(info)))
        if getattr(self, "_last_targets", None) is not None:
            self._last_targets = (info, self._last_targets[1])
        return info

    def target_value(self, idx: int) -> float:
        out = c_double()
        self._ck(self._L.sk_target_value(self._h, int(idx), byref(out)))
        return out.value

    def run_begin(self):
        self._ck(self._L.sk_run_begin(self._h))

    def zero_lag_set(self, value: float):
        self._ck(self._L.sk_zero_lag_set(self._h, float(value)))

    def panel_begin(self, ix1: int, hi: int):
        lo, hi_ = c_double(), c_double()
        self._ck(self._L.sk_panel_begin(self._h, int(ix1), int(hi), byref(lo), byref(hi_)))
        return lo.value, hi_.value

    def panel_set_range(self, r_lo: float, r_hi: float, n_active_global: int = 0):
        self._ck(self._L.sk_panel_set_range(self._h, float(r_lo), float(r_hi), int(n_active_global)))

    def subinterval(self, a: float, b: float, cmul: float, p: float, kernel: int, logw: bool, speculate=None,
                    nu: int = 0, xdiv_pow: float = 0.0) -> float:
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1 if logw else 0, int(nu), 0, float(xdiv_pow),
                            ctypes.pointer(speculate) if speculate is not None else None)
        out = c_double()
        self._ck(self._L.sk_subinterval(self._h, float(a), float(b), byref(o), byref(out)))
        return out.value

    def subinterval_host(self, a, b, no1, buf1, no2, buf2, cmul, p, kernel, logw, speculate=None,
                         nu: int = 0, xdiv_pow: float = 0.0) -> float:
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1 if logw else 0, int(nu), 0, float(xdiv_pow),
                            ctypes.pointer(speculate) if speculate is not None else None)
        out = c_double()
        no1, buf1, no2, buf2 = _f64(no1), _f64(buf1), _f64(no2), _f64(buf2)
        self._ck(self._L.sk_subinterval_host(self._h, float(a), float(b), _p(no1), _p(buf1), _p(no2), _p(buf2),
                                             byref(o), byref(out)))
        return out.value

    def subinterval_logw_host(self, a, b, no1, bufa1, bufb1, no2, bufa2, bufb2, cmul, p, i0_coef, denom,
                              kernel: int = SK_KERNEL_COS, nu: int = 0, xdiv_pow: float = 0.0) -> float:
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1, int(nu), 0, float(xdiv_pow), None)
        out = c_double()
        if no1 is None:          # built-in density: both integrands are evaluated on the device
            ptrs = [None] * 6
        else:
            arrs = [_f64(x) for x in (no1, bufa1, bufb1, no2, bufa2, bufb2)]
            ptrs = [_p(x) for x in arrs]
        self._ck(self._L.sk_subinterval_logw_host(self._h, float(a), float(b), *ptrs, byref(o),
                                                  float(i0_coef), float(denom), byref(out)))
        return out.value

    def sources_get(self, rule: int):
        n = self.m * self.k * (2 if rule else 1)
        no, buf = np.empty(n), np.empty(n)
        self._ck(self._L.sk_sources_get(self._h, int(rule), _p(no), _p(buf)))
        return no, buf

    def subinterval_accept(self):
        self._ck(self._L.sk_subinterval_accept(self._h))

    def panel_commit(self):
        self._ck(self._L.sk_panel_commit(self._h))

    def converge_scan(self, args: ScanArgs):
        new_hi, r = c_int64(), c_double()
        self._ck(self._L.sk_converge_scan(self._h, byref(args), byref(new_hi), byref(r)))
        return new_hi.value, r.value

    def converge_apply(self, args: ScanArgs, new_hi: int):
        self._ck(self._L.sk_converge_apply(self._h, byref(args), int(new_hi)))

    def target_upper_index(self, r: float) -> int:
        out = c_int64()
        self._ck(self._L.sk_target_upper_index(self._h, float(r), byref(out)))
        return out.value

    def results_get(self, n_in: int, want_errors: bool = True, out_vals=None, out_errs=None):
        vals = out_vals if out_vals is not None else np.empty(n_in)
        errs = (out_errs if out_errs is not None else np.empty(n_in)) if want_errors else None
        self._ck(self._L.sk_results_get(self._h, vals.ctypes.data, errs.ctypes.data if errs is not None else None))
        return vals, errs

    def results_get_async(self, out_vals, out_errs=None):
        """Gather on the compute stream, copy to the (pinned) host arrays on a second stream; `results_wait` blocks."""
        self._ck(self._L.sk_results_get_async(self._h, out_vals.ctypes.data, out_errs.ctypes.data if out_errs is not None else None))

    def results_wait(self):
        self._ck(self._L.sk_results_wait(self._h))

    def results_get_device(self, vals_ptr: int, errs_ptr: int = 0):
        self._ck(self._L.sk_results_get_device(self._h, c_void_p(vals_ptr), c_void_p(errs_ptr) if errs_ptr else None))

    def stats(self) -> dict:
        st = Stats()
        self._ck(self._L.sk_stats_get(self._h, byref(st)))
        return st.as_dict()


class GroupSession:
    """Owns one sk_group: several GPUs behind the call sequence of a single Session (include/spectralkernels_b200.h,
    "device group").  The distances of `targets_set` are cut into contiguous chunks, one per device; every step is
    enqueued on all devices before anything is read back.  Only the entry points the host-array path of
    `kernel_values` uses exist on a group (no device-pointer / pair-list / log-weighted variants)."""

    def __init__(self, devices, timing: bool = False):
        self._L = load()
        self._h = c_void_p()
        devs = (c_int32 * len(devices))(*[int(d) for d in devices])
        rc = self._L.sk_group_create(devs, len(devices), byref(self._h))
        if rc != SK_OK:
            raise SkError(rc, "sk_group_create failed (are these CUDA devices visible? there is no CPU fallback)")
        self.devices = [int(d) for d in devices]
        self.device = self.devices[0]
        if timing:
            self.set_timing(True)
        self.rule_key = None
        self.sdf_key = None

    def _ck(self, rc: int):
        if rc != SK_OK:
            detail = self._L.sk_group_last_error(self._h)
            base = self._L.sk_error_string(rc)
            raise SkError(rc, (detail or base or b"").decode())

    def close(self):
        if self._h is not None and self._h.value:
            self._L.sk_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._ck(self._L.sk_group_synchronize(self._h))

    def set_timing(self, on: bool):
        self._ck(self._L.sk_group_set_timing(self._h, 1 if on else 0))

    def set_nufft_eps(self, eps: float):
        self._ck(self._L.sk_group_set_nufft_eps(self._h, float(eps)))

    def rule_set(self, m: int, k: int, p: float, leg=None, jac=None):
        key = (int(m), int(k), float(p), leg is None, jac is None)
        if leg is None and jac is None and key == self.rule_key:
            return
        args, keep = [], []
        for rule in (leg, jac):
            if rule is None:
                args += [None] * 4
            else:
                arrs = [_f64(a) for a in rule]
                keep += arrs
                args += [_p(a) for a in arrs]
        self._ck(self._L.sk_group_rule_set(self._h, int(m), int(k), float(p), *args))
        self.rule_key = key
        self.m, self.k, self.p = int(m), int(k), float(p)

    def rule_get(self, which: int):
        n = self.m * (2 if which in (1, 3) else 1)
        no, wt = np.empty(n), np.empty(n)
        self._ck(self._L.sk_group_rule_get(self._h, int(which), _p(no), _p(wt)))
        return no, wt

    def sdf_builtin(self, family: int, params, deriv_index: int = 0):
        pr = _f64(params)
        self._ck(self._L.sk_group_sdf_builtin(self._h, int(family), _p(pr) if pr.size else None, pr.size, int(deriv_index)))
        self.sdf_key = (int(family), tuple(pr.tolist()), int(deriv_index))

    def targets_set(self, xs: np.ndarray) -> TargetInfo:
        xs = _f64(xs)
        info = TargetInfo()
        self._ck(self._L.sk_group_targets_set(self._h, xs.ctypes.data, xs.size, byref(info)))
        self._last_targets = (info, int(info.n_in))
        return info

    def _unsupported(self, *a, **k):
        raise NotImplementedError("a device group takes host arrays of distances only (sk_group_targets_set)")

    targets_set_device = targets_set_pairs = targets_scale = subinterval_logw_host = results_get_device = \
        results_get_async = _unsupported

    def run_begin(self):
        self._ck(self._L.sk_group_run_begin(self._h))

    def zero_lag_set(self, value: float):
        self._ck(self._L.sk_group_zero_lag_set(self._h, float(value)))

    def panel_begin(self, ix1: int, hi: int):
        lo, hi_ = c_double(), c_double()
        self._ck(self._L.sk_group_panel_begin(self._h, int(ix1), int(hi), byref(lo), byref(hi_)))
        return lo.value, hi_.value

    def subinterval(self, a: float, b: float, cmul: float, p: float, kernel: int, logw: bool, speculate=None,
                    nu: int = 0, xdiv_pow: float = 0.0) -> float:
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1 if logw else 0, int(nu), 0, float(xdiv_pow),
                            ctypes.pointer(speculate) if speculate is not None else None)
        out = c_double()
        self._ck(self._L.sk_group_subinterval(self._h, float(a), float(b), byref(o), byref(out)))
        return out.value

    def subinterval_host(self, a, b, no1, buf1, no2, buf2, cmul, p, kernel, logw, speculate=None,
                         nu: int = 0, xdiv_pow: float = 0.0) -> float:
        o = SubintervalOpts(float(cmul), float(p), int(kernel), 1 if logw else 0, int(nu), 0, float(xdiv_pow),
                            ctypes.pointer(speculate) if speculate is not None else None)
        out = c_double()
        no1, buf1, no2, buf2 = _f64(no1), _f64(buf1), _f64(no2), _f64(buf2)
        self._ck(self._L.sk_group_subinterval_host(self._h, float(a), float(b), _p(no1), _p(buf1), _p(no2), _p(buf2),
                                                   byref(o), byref(out)))
        return out.value

    def subinterval_accept(self):
        self._ck(self._L.sk_group_subinterval_accept(self._h))

    def panel_commit(self):
        self._ck(self._L.sk_group_panel_commit(self._h))

    def converge_scan(self, args: ScanArgs):
        new_hi, r = c_int64(), c_double()
        self._ck(self._L.sk_group_converge_scan(self._h, byref(args), byref(new_hi), byref(r)))
        return new_hi.value, r.value

    def converge_apply(self, args: ScanArgs, new_hi: int):
        self._ck(self._L.sk_group_converge_apply(self._h, byref(args), int(new_hi)))

    def results_get(self, n_in: int, want_errors: bool = True, out_vals=None, out_errs=None):
        vals = out_vals if out_vals is not None else np.empty(n_in)
        errs = (out_errs if out_errs is not None else np.empty(n_in)) if want_errors else None
        self._ck(self._L.sk_group_results_get(self._h, vals.ctypes.data, errs.ctypes.data if errs is not None else None))
        return vals, errs

    def stats(self) -> dict:
        st = Stats()
        self._ck(self._L.sk_group_stats_get(self._h, byref(st)))
        return st.as_dict()


def bind_to_gpu_cpus(device: int = 0) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device` (its NUMA node / PCIe root), so that
    the pinned host buffers allocated afterwards and the staging copies stay on the GPU's own socket.  With one process
    per GPU on a two-socket box this is what keeps eight concurrent H2D / D2H streams from crossing the inter-socket
    link.  Returns False (and changes nothing) when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


def host_gauss_rule(n: int, p: float = 0.0):
    no, wt = np.empty(n), np.empty(n)
    rc = load().sk_host_gauss_rule(int(n), float(p), _p(no), _p(wt))
    if rc != SK_OK:
        raise SkError(rc, "sk_host_gauss_rule failed")
    return no, wt
