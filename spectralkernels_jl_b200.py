"""
Import shim: the package directory is `spectralkernels.jl_b200/` (a dot is not importable as a
module name), so this module loads it under the name `spectralkernels_jl_b200`:

    import spectralkernels_jl_b200 as sk
    cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0, 1.0, 1.5))
    vals, errs = sk.kernel_values(cfg, rs)
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spectralkernels.jl_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
