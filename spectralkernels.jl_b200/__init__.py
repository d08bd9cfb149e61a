"""
spectralkernels.jl_b200 -- B200 (sm_100a) evaluator for the K(r) hot path of pbeckman/SpectralKernels.jl.

    csrc/                 CUDA kernels + the C ABI (include/spectralkernels_b200.h) -> libsk_b200.so
    _capi.py              ctypes binding of that ABI (what Julia binds with ccall)
    adaptive.py           host mirror of AdaptiveKernelConfig / kernel_values (scalar control flow only)
    sdf.py                built-in spectral-density families with device generators
    derivatives.py        K', dK/dtheta_j, dK/dalpha as further kernel_values runs over the same lags
    sharded.py            scalar reductions for a target-sharded multi-GPU run (torch.distributed)

The directory name contains a dot, so import it through the loader module at the repository root:

    import spectralkernels_jl_b200 as sk
"""
from . import _capi, sdf
from ._capi import GroupSession, PinnedArray, Session, SkError, bind_to_gpu_cpus, host_gauss_rule, load
from .adaptive import (AdaptiveKernelConfig, compute_k0, estimate_tail_decay, gen_derivative_config,
                       gen_new_sdf_config, kernel_values)
from .derivatives import kernel_derivative, kernel_sdf_derivatives, kernel_singularity_derivative
from .model import (NoWarping, SpectralKernel, SpectralModel, build_dense_cov_matrix, dense_index_pairs, gen_kernel,
                    gen_kernel_dual, gen_kernel_jacobian, gen_kernel_setup)
from .sdf import Exponential, Matern

__all__ = ["AdaptiveKernelConfig", "kernel_values", "compute_k0", "estimate_tail_decay", "gen_derivative_config",
           "gen_new_sdf_config", "kernel_derivative", "kernel_sdf_derivatives", "kernel_singularity_derivative", "Matern", "Exponential", "Session", "GroupSession", "SkError", "PinnedArray", "host_gauss_rule", "bind_to_gpu_cpus",
           "load", "sdf", "NoWarping", "SpectralModel", "SpectralKernel", "dense_index_pairs", "gen_kernel_setup", "gen_kernel",
           "gen_kernel_jacobian", "gen_kernel_dual", "build_dense_cov_matrix"]
