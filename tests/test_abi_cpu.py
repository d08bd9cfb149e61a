"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/*.h declares,
host-side plan helpers work, and the product refuses to run without a GPU (no CPU fallback)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sk():
    subprocess.check_call([sys.executable, os.path.join(ROOT, "spectralkernels.jl_b200", "build.py")])
    import spectralkernels_jl_b200 as sk
    return sk


def test_header_symbols_exported(sk):
    hdr = open(os.path.join(ROOT, "include", "spectralkernels_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char \*)\s*\*?\s*(sk_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 30
    lib = sk.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(sk._capi.SIGNATURES), declared ^ set(sk._capi.SIGNATURES)
    assert lib.sk_abi_version() == 2


def test_no_cpu_fallback(sk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sk.SkError):
        sk.Session(0)
    cfg = sk.AdaptiveKernelConfig(sk.Matern())
    with pytest.raises(sk.SkError):
        sk.kernel_values(cfg, np.linspace(0, 1, 5), k0=1.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spectralkernels.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "sk_oracle" not in txt and "libsk_oracle" not in txt and "closed_forms" not in txt, fn


def test_host_gauss_rule_matches_oracle(sk):
    import sk_oracle as so
    for n, p in ((64, 0.0), (4096, 0.0), (8192, -0.5), (4096, 1.5)):
        a, b = sk.host_gauss_rule(n, p)
        c, d = so.gauss_rule(n, p)
        assert np.max(np.abs(a - c)) <= 2.3e-16
        assert np.max(np.abs(b / d - 1)) <= 2e-13


def test_config_validation(sk):
    with pytest.raises(ValueError):
        sk.AdaptiveKernelConfig(sk.Matern(), convergence_criteria="nope")
    with pytest.raises(ValueError):
        sk.AdaptiveKernelConfig(sk.Matern(), alpha=1.0)
    with pytest.warns(UserWarning):
        cfg = sk.AdaptiveKernelConfig(sk.Matern(), tol=1e-13)
    assert cfg.quadspec == (4096, 1)
    cfg = sk.AdaptiveKernelConfig(sk.Matern(), derivative=True, alpha=0.25)
    assert cfg.p == 0.75 and cfg.c == 2.0 * (-2 * np.pi)


def test_host_scalars_match_oracle(sk):
    """compute_k0 / estimate_tail_decay are host code on both sides (adaptive.jl:74-91, :204-220)."""
    import sk_oracle as so
    S = sk.Matern(2.14, 0.97, 0.89)
    cfg = sk.AdaptiveKernelConfig(S)
    ocfg = so.OracleConfig(S)
    assert abs(sk.compute_k0(cfg) / so.compute_k0(ocfg) - 1) < 1e-12
    for (a, b) in ((0.0, 32768.0), (32768.0, 65536.0), (65536.0, 5.19e6)):
        assert np.allclose(sk.estimate_tail_decay(cfg, a, b), so.estimate_tail_decay(ocfg, a, b), rtol=1e-13)
