"""
Pins panel traces to the REFERENCE itself: tests/golden/reference_traces.json is written by
tests/golden/make_traces.jl (run with Julia on the unmodified pbeckman/SpectralKernels.jl; the script only replaces
the two printing helpers the package calls under verbose=true, src/utils.jl:12-25, by recorders).  The build image has
no Julia, so the fixture may be absent: the tests then SKIP and DESIGN.md / oracle/README.md keep saying "traces:
parity unpinned".  With the fixture present,

  * (CPU) the oracle's trace for every recorded case must equal the reference's: sub-intervals (a, b, accepted) and
    panels (a, b, hi_before, hi_after) bit for bit, sampled values within 10 tol K(0);
  * (GPU) the product's trace must equal it too.

The distance sets are regenerated from seeds by tests/golden/make_trace_inputs.py (bit-identical in Julia and numpy).
"""
import json
import math
import os

import numpy as np
import pytest

import sk_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "reference_traces.json")
INP = os.path.join(HERE, "golden", "trace_inputs")

needs_fixture = pytest.mark.skipif(not os.path.exists(FIX), reason="reference_traces.json absent: run "
                                   "tests/golden/make_traces.jl with Julia on the reference (traces: parity unpinned)")


def _matern(parms, d=1):
    return lambda w: parms[0] * (parms[1] ** 2 + np.asarray(w, dtype=float) ** 2) ** (-parms[2] - d / 2)


def _cases():
    """name -> (host spectral density, config keywords, kernel_values keywords, input file or array, device family)"""
    p2 = (1.0 / (math.pi / 2), 1.0, 1.5)
    out = {
        "config1_readme": (lambda w: (1 + w ** 2) ** -2.0, {}, {}, 10 ** np.linspace(-6, 0, 1000), ("matern", (1.0, 1.0, 1.5), 1)),
        "config2_2e3.f64": (_matern(p2), {}, {"k0": 1.0}, "config2_2e3.f64", ("matern", p2, 1)),
        "config2_1e7.f64": (_matern(p2), {}, {"k0": 1.0}, "config2_1e7.f64", ("matern", p2, 1)),
        "config3_dim1": (_matern((1.0, 1.0, 1.5), 1), {"alpha": 0.5, "dim": 1}, {}, "config3_lags.f64", ("matern", (1.0, 1.0, 1.5), 1)),
        "config3_dim2": (_matern((1.0, 1.0, 1.5), 2), {"alpha": 0.5, "dim": 2}, {}, "config3_lags.f64", ("matern", (1.0, 1.0, 1.5), 2)),
        "config4_K": (_matern(p2), {}, {"k0": 1.0}, "config4_1e6.f64", ("matern", p2, 1)),
        "config4_dK": (_matern(p2), {"derivative": True}, {"k0": 1.0}, "config4_1e6.f64", ("matern", p2, 1)),
        "config5_dim2": (_matern((1.0, 4.0, 1.5), 2), {"dim": 2}, {}, "config5_lags.f64", ("matern", (1.0, 4.0, 1.5), 2)),
    }
    return out


def _key(trace):
    subs = [(t["a"], t["b"], bool(t["accepted"])) for t in trace if t["kind"] == "subinterval"]
    pans = [(t["a"], t["b"], int(t["hi_before"]), int(t["hi_after"])) for t in trace if t["kind"] == "panel"]
    return subs, pans


def _ref_key(case, unique_offset):
    """The reference counts indices into the sorted unique vector, 1-based -- as the oracle and the product do."""
    return _key(case["trace"])


def _load_inputs(spec):
    if isinstance(spec, str):
        path = os.path.join(INP, spec)
        if not os.path.exists(path):
            pytest.skip(f"{path} absent: run tests/golden/make_trace_inputs.py")
        return np.fromfile(path, dtype="<f8")
    return spec


@needs_fixture
def test_oracle_traces_equal_reference_traces():
    ref = json.load(open(FIX))
    table = _cases()
    checked = 0
    for case in ref["cases"]:
        if case["name"] not in table:
            continue
        S, ckw, kkw, inp, _ = table[case["name"]]
        xs = _load_inputs(inp)
        if xs.size > 20000:
            continue                                   # the oracle's direct sums are for the small cases
        cfg = so.OracleConfig(S, **ckw)
        tr = []
        vals, _ = so.kernel_values(cfg, xs, trace=tr, **kkw)
        assert _key(tr) == _ref_key(case, 0), case["name"]
        k0 = kkw.get("k0") or so.compute_k0(cfg)
        idx = np.asarray(case["sample_index"], dtype=np.int64)
        assert np.max(np.abs(vals[idx] - np.asarray(case["sample_values"]))) <= 10 * cfg.tol * abs(k0), case["name"]
        checked += 1
    assert checked > 0


@needs_fixture
@pytest.mark.gpu
def test_product_traces_equal_reference_traces():
    import spectralkernels_jl_b200 as sk
    ref = json.load(open(FIX))
    table = _cases()
    checked = 0
    for case in ref["cases"]:
        if case["name"] not in table:
            continue
        S, ckw, kkw, inp, (fam, parms, d) = table[case["name"]]
        xs = _load_inputs(inp)
        cfg = sk.AdaptiveKernelConfig(sk.Matern(*parms, d=d), **ckw)
        tr = []
        k0 = kkw.get("k0") or sk.compute_k0(cfg)
        vals, _ = sk.kernel_values(cfg, xs, k0=k0, trace=tr)
        assert _key(tr) == _ref_key(case, 0), case["name"]
        idx = np.asarray(case["sample_index"], dtype=np.int64)
        assert np.max(np.abs(vals[idx] - np.asarray(case["sample_values"]))) <= 10 * cfg.tol * abs(k0), case["name"]
        checked += 1
    assert checked > 0


def test_trace_fixture_tooling_is_consistent():
    """Without Julia the fixture cannot be produced here; what CAN be checked is that the loader's case table, the
    generator script and the input writer name the same cases and files."""
    jl = open(os.path.join(HERE, "golden", "make_traces.jl")).read()
    py = open(os.path.join(HERE, "golden", "make_trace_inputs.py")).read()
    for name, (_, _, _, inp, _) in _cases().items():
        if isinstance(inp, str):
            assert inp in jl and inp in py, inp
        stem = name if not name.endswith(".f64") else name
        assert stem.split("_dim")[0] in jl, name
    for hook in ("print_panel_info", "print_panel_convergence"):
        assert hook in jl
