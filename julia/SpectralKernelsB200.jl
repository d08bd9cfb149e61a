# SpectralKernelsB200.jl -- the reference-side binding a maintainer of SpectralKernels.jl adds so that the package's
# OWN public API runs on B200s through libsk_b200.so (C ABI: include/spectralkernels_b200.h).
#
# NOT EXECUTED in this repository: the build image has no Julia.  The identical C ABI is exercised through ctypes by
# spectralkernels.jl_b200/_capi.py + adaptive.py, which follows the same control flow call for call.  Keep this file
# thin: only scalars cross the boundary inside the loops.
#
# How the dispatch works (no method of the package is overwritten, nothing is pirated): wrap the spectral density,
#
#     using SpectralKernels, SpectralKernelsB200
#     S   = B200(w -> (1 + w^2)^(-2))                        # any closure: evaluated in Julia, strengths uploaded
#     S   = B200(matern_sdf; builtin=SK_SDF_MATERN)          # f(w, phi, rho, nu): evaluated by the device generator
#     S   = B200(f; devices=[0, 1, 2, 3])                    # one caller, four GPUs (the C ABI's device group)
#     cfg = AdaptiveKernelConfig(S; tol=1e-8)                # unchanged constructor (src/adaptive.jl:24-59)
#     Ks, errs = kernel_values(cfg, rs)                      # unchanged call      (src/adaptive.jl:95-108)
#
# `AdaptiveKernelConfig{S,dS}` carries the type of `f`, so the method below -- defined on configs whose density is a
# `B200`, a `ParametricFunction` of one or a `ParametricDerivative` of one -- is what Julia selects for
# `kernel_values(cfg, xs)`.  Everything above that call keeps working unchanged and now runs on the GPU, because it
# only ever calls `kernel_values`: `gen_kernel` (src/model.jl:73-77, via `ParametricFunction(sm.cfg.f, params)`),
# `kernel_sdf_derivatives` / `kernel_warping_gradients` / `kernel_singularity_derivative` (src/derivatives.jl:51-81),
# the ForwardDiff and ChainRules extensions (ext/SpectralKernelsForwardDiffExt.jl:7-22) and the Vecchia extension
# (ext/SpectralKernelsVecchiaExt.jl:19-27).
#
# This file uses the plain blocking entry points (sk_targets_set, sk_subinterval, sk_results_get).  The ABI also offers
# them in two halves plus device-guarded chained launches (sk_targets_begin/_early_range/_end, sk_first_panel_early,
# sk_subinterval_begin/_chain/_end, sk_results_chain_device: INTEGRATION.md) so that `estimate_tail_decay` for the next
# panel runs while the device works and consecutive panels run back to back; adaptive.py shows the call order.  They only
# move WHEN work is enqueued -- values, error estimates and traces are bit-identical -- so adopting them is optional.
module SpectralKernelsB200

using SpectralKernels
import SpectralKernels: AdaptiveKernelConfig, ParametricFunction, ParametricDerivative, compute_k0,
                        estimate_tail_decay, updatequadbufs!, quadsz, kernel_values, build_dense_cov_matrix

const libsk = get(ENV, "SK_B200_LIB", "libsk_b200.so")

const SK_KERNEL_COS, SK_KERNEL_SIN, SK_KERNEL_BESSEL = Cint(0), Cint(1), Cint(2)
const SK_CRIT = Dict(:panel => Cint(0), :tails => Cint(1), :both => Cint(2))
const SK_SDF_MATERN, SK_SDF_EXPONENTIAL = Cint(1), Cint(2)

struct TargetInfo            # sk_target_info
  n_in::Int64; n_unique::Int64; has_zero::Int32; _pad::Int32; r_min_pos::Float64; r_max::Float64
end
struct ScanArgs              # sk_scan_args
  trunc_a::Float64; trunc_num::Float64; xpow::Float64; tau::Float64; criteria::Int32; _pad::Int32
end
struct SubintervalOpts       # sk_subinterval_opts
  cmul::Float64; p::Float64; kernel::Int32; logw::Int32
  nu::Int32; _pad::Int32     # Bessel order for SK_KERNEL_BESSEL (dim >= 2)
  xdiv_pow::Float64          # dim/2 - 1
  speculate::Ptr{ScanArgs}   # C_NULL, or the panel's scan arguments on the panel's first sub-interval
end

# ---- the wrapper that selects the backend ---------------------------------------------------------------------
struct B200{F} <: Function
  fn::F
  devices::Vector{Int32}                 # one entry: a single context; several: a device group (sk_group_*)
  builtin::Cint                          # 0: evaluate fn in Julia; SK_SDF_MATERN / SK_SDF_EXPONENTIAL: device generator
end
B200(fn; devices=[0], builtin=Cint(0)) = B200(fn, Int32.(collect(devices)), Cint(builtin))
(b::B200)(w, params...) = b.fn(w, params...)

const OnB200 = Union{B200, ParametricFunction{<:B200}, ParametricDerivative{<:B200}}
backend_of(f::B200) = f
backend_of(f::ParametricFunction{<:B200}) = f.fn
backend_of(f::ParametricDerivative{<:B200}) = f.swap_sdf.fn.fn
# device generator for this integrand?  (family, parameters, which parameter derivative); only for a ParametricFunction
# of a built-in family with d = 1 Matern parameters (phi, rho, nu) or exponential parameters (phi, alpha)
builtin_of(f::B200) = nothing
function builtin_of(f::ParametricFunction{<:B200})
  fam = f.fn.builtin
  fam == 0 && return nothing
  prm = collect(Float64, f.params)
  fam == SK_SDF_MATERN && length(prm) == 3 && push!(prm, 1.0)          # (phi, rho, nu, d)
  (fam, prm, Cint(0))
end
function builtin_of(f::ParametricDerivative{F,P,J}) where {F<:B200,P,J}
  b = builtin_of(f.swap_sdf.fn)
  isnothing(b) ? nothing : (b[1], b[2], Cint(J - 1))                   # J - 1: args[1] is the frequency (wrappers.jl:24)
end

# ---- one session (context or group) per task and device list ----------------------------------------------------
mutable struct Session
  h::Ptr{Cvoid}
  group::Bool
  function Session(devices::Vector{Int32})
    r = Ref{Ptr{Cvoid}}(C_NULL)
    grp = length(devices) > 1
    rc = grp ? ccall((:sk_group_create, libsk), Cint, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devices, length(devices), r) :
               ccall((:sk_ctx_create, libsk), Cint, (Cint, Ref{Ptr{Cvoid}}), devices[1], r)
    rc == 0 || error("libsk_b200: cannot open device(s) $devices ($rc).  There is no CPU fallback.")
    s = new(r[], grp)
    finalizer(x -> x.group ? ccall((:sk_group_destroy, libsk), Cint, (Ptr{Cvoid},), x.h) :
                             ccall((:sk_ctx_destroy, libsk), Cint, (Ptr{Cvoid},), x.h), s)
    s
  end
end
# cfg.buffers / cfg.splittingheap are per-config scratch that must not be shared between tasks (src/adaptive.jl:17-21);
# the device scratch follows the same rule: one session per task and device list
session(devices::Vector{Int32}) = get!(() -> Session(devices), task_local_storage(), (:sk_b200, Tuple(devices)))::Session

function ck(s::Session, rc::Cint)
  rc == 0 && return
  msg = s.group ? unsafe_string(ccall((:sk_group_last_error, libsk), Cstring, (Ptr{Cvoid},), s.h)) :
                  unsafe_string(ccall((:sk_last_error, libsk), Cstring, (Ptr{Cvoid},), s.h))
  error("libsk_b200 [$rc]: $msg")     # never throws across the ABI: the C side only returns codes
end
# sk_X for a context, sk_group_X for a group: same arguments
macro sk(s, name, argtypes, args...)
  g = QuoteNode(Symbol("sk_group_", name)); c = QuoteNode(Symbol("sk_", name))
  esc(quote
    ck($s, $s.group ? ccall(($g, libsk), Cint, (Ptr{Cvoid}, $(argtypes.args...)), $s.h, $(args...)) :
                      ccall(($c, libsk), Cint, (Ptr{Cvoid}, $(argtypes.args...)), $s.h, $(args...)))
  end)
end

# --- Level 0: replaces src/utils.jl:10 -----------------------------------------------------------------
function finufft1d3_b200(s::Session, w::Vector{Float64}, c::Vector{ComplexF64}, x::Vector{Float64})
  s.group && error("sk_nufft1d3 takes a single context")
  out = Vector{ComplexF64}(undef, length(x))
  GC.@preserve w c x out ck(s, ccall((:sk_nufft1d3, libsk), Cint,
      (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{ComplexF64}, Int64, Ptr{Float64}, Ptr{ComplexF64}, Float64),
      s.h, length(w), w, c, length(x), x, out, 1e-15))
  out
end

# target-independent pieces of truncation_error_estimate (src/adaptive.jl:222-229), IEEE arithmetic throughout
scan_args(config, b, cc, d, tau, crit) = crit == :panel ?
  ScanArgs(0.0, 0.0, (config.dim + 1)/2, tau, SK_CRIT[:panel], 0) :
  ScanArgs(-cc/(d + config.dim)*b^(d + config.dim), cc*b^(d + (config.dim - 1)/2), (config.dim + 1)/2, tau, SK_CRIT[crit], 0)

# --- Level 1: kernel_values with every O(N) pass on the device ---------------------------------------------------
# `targets`: nothing (upload xs), or a closure that sets the targets itself (pair lists: see build_dense_cov_matrix)
function kernel_values(config::AdaptiveKernelConfig{<:OnB200}, xs::AbstractVector{Float64};
                       k0=compute_k0(config), param_derivative=false, verbose=false, targets=nothing)
  be = backend_of(config.f)
  s  = session(be.devices)
  (m, k) = config.quadspec
  lr, jr = config.legrule, config.jacrule
  # QuadRule (src/quadrature.jl:27-47): pass FastGaussQuadrature's own nodes so both paths share them
  GC.@preserve lr jr @sk(s, rule_set, (Int32, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                                       Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                         m, k, config.p, lr.no1, lr.wt1, lr.no2, lr.wt2, jr.no1, jr.wt1, jr.no2, jr.wt2)
  builtin = builtin_of(config.f)
  if !isnothing(builtin)
    (fam, prm, dj) = builtin
    @sk(s, sdf_builtin, (Int32, Ptr{Float64}, Int32, Int32), fam, prm, length(prm), dj)
  end
  info = Ref(TargetInfo(0, 0, 0, 0, 0.0, 0.0))
  n_out = length(xs)
  if isnothing(targets)
    xv = convert(Vector{Float64}, xs)
    GC.@preserve xv @sk(s, targets_set, (Ptr{Float64}, Int64, Ref{TargetInfo}), xv, length(xv), info)   # adaptive.jl:99,113-120
  else
    n_out = targets(s, info)
  end
  verbose && println("Reducing $(info[].n_in) to $(info[].n_unique) unique lags for evaluation...")
  @sk(s, run_begin, ())
  n   = info[].n_unique
  ix1 = 1
  if info[].has_zero != 0                                                  # adaptive.jl:133-146
    ix1 = 2
    z = config.derivative ? 0.0 : (param_derivative ? compute_k0(config) : k0)
    @sk(s, zero_lag_set, (Float64,), z)
  end
  hi, r_hi   = n, (n >= ix1 ? info[].r_max : 0.0)
  conv_crit  = config.convergence_criteria
  (a, b)     = (0.0, 0.0)
  dim        = config.dim
  if dim == 1
    (kernel, nu, xdiv) = (config.derivative ? SK_KERNEL_SIN : SK_KERNEL_COS, 0, 0.0)   # quadrature.jl:177
  else
    # (:J, dim/2) or (:J, dim/2-1), quadrature.jl:179; Int64(...) throws for odd dim as in the reference (:138).
    # The library replaces FastHankelTransform's nufht (:139-143) by its own O(N) nonuniform Hankel transform
    # (orders 0..3) and takes the direct Bessel summation (:145-160) for small active sets.
    (kernel, nu, xdiv) = (SK_KERNEL_BESSEL, Int64(config.derivative ? dim/2 : dim/2 - 1), dim/2 - 1)
  end
  tau = config.tol*abs(k0)/2
  while r_hi > 0                                                           # adaptive.jl:149
    (a, b) = (b, b + quadsz(config)/(2*r_hi))                              # adaptive.jl:152
    verbose && println("\nintegrating panel w ∈ [$a, $b] (length $(b - a)) to resolve $hi points x ≤ $r_hi")   # utils.jl:12-15
    @sk(s, panel_begin, (Int64, Int64, Ptr{Float64}, Ptr{Float64}), ix1, hi, C_NULL, C_NULL)
    # The tail fit depends on (a, b) only, so it is evaluated BEFORE the panel is integrated (the reference does it
    # after, adaptive.jl:168): the scan arguments then ride along with the panel's first sub-interval and the device
    # fuses accept / commit / scan into the interpolation kernel (sk_subinterval_opts.speculate)
    (cc, d) = (conv_crit == :panel) ? (NaN, NaN) : estimate_tail_decay(config, a, b, d=config.tail)
    if (isnan(cc) || isnan(d)) && conv_crit != :panel; conv_crit = :panel; end     # adaptive.jl:170-175
    sa = Ref(scan_args(config, b, cc, d, tau, conv_crit))
    # ---- fourier_integrate_interval (quadrature.jl:169-275): scalar control flow only ----
    stack = [(a, b, config.tol)]
    first = true
    GC.@preserve sa while !isempty(stack)
      (_a, _b, _tol) = pop!(stack)
      spec  = first ? Base.unsafe_convert(Ptr{ScanArgs}, sa) : Ptr{ScanArgs}(C_NULL)
      first = false
      opts  = Ref(SubintervalOpts(config.c, config.p, kernel, config.logw ? 1 : 0, nu, 0, xdiv, spec))
      mx    = Ref(0.0)
      if _a == 0.0 && config.p != 0.0 && config.logw
        # log-weighted origin sub-interval by parts (quadrature.jl:186-228): Julia owns f and df and evaluates both
        # integrands with updatequadbufs!; the device does the transforms (:cis for dim = 1; Bessel orders dim/2-1 and
        # dim/2 for dim = 2) and I = (I0 - A + 2 pi x B)/(dim - alpha)
        dim <= 2 || error("singularity derivative not implemented in d > 2")                      # :222-223
        s.group && error("the log-weighted origin sub-interval takes a single context")
        (f, df) = (config.f, config.df)
        i0   = _b^(dim/2 + 1 - config.alpha)*log(_b)*f(_b)                                          # :189
        lopt = Ref(SubintervalOpts(config.c, config.p, dim == 1 ? SK_KERNEL_COS : SK_KERNEL_BESSEL, 1,
                                   dim == 1 ? 0 : Int64(dim/2 - 1), 0, dim/2 - 1, C_NULL))
        if !isnothing(builtin) && builtin[3] == 0
          # a shipped family: dS/dw is closed-form, the device evaluates both integrands itself (all pointers NULL)
          ck(s, ccall((:sk_subinterval_logw_host, libsk), Cint,
              (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
               Ptr{Float64}, Ref{SubintervalOpts}, Float64, Float64, Ref{Float64}),
              s.h, _a, _b, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, lopt, i0, dim - config.alpha, mx))
        else
          (no1, ba1, no2, ba2) = updatequadbufs!(config.buffers, config.legrule, config.jacrule,
                                                 w -> f(w) + w*log(w)*df(w), _a, _b; p=config.p)
          (no1, ra1, no2, ra2) = (copy(no1), real.(ba1), copy(no2), real.(ba2))
          (_, bb1, _, bb2)     = updatequadbufs!(config.buffers, config.legrule, config.jacrule,
                                                 w -> w*log(w)*f(w), _a, _b; p=config.p)
          (rb1, rb2) = (real.(bb1), real.(bb2))
          GC.@preserve no1 ra1 rb1 no2 ra2 rb2 ck(s, ccall((:sk_subinterval_logw_host, libsk), Cint,
              (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
               Ptr{Float64}, Ref{SubintervalOpts}, Float64, Float64, Ref{Float64}),
              s.h, _a, _b, no1, ra1, rb1, no2, ra2, rb2, lopt, i0, dim - config.alpha, mx))
        end
      elseif isnothing(builtin)
        origin = (_a == 0.0 && config.p != 0.0)
        f = origin ? config.f : (w -> w^config.p * (config.logw ? log(w) : 1) * config.f(w))
        (no1, buf1, no2, buf2) = updatequadbufs!(config.buffers, config.legrule, config.jacrule, f, _a, _b;
                                                 p=(origin ? config.p : 0))
        rb1, rb2 = real.(buf1), real.(buf2)
        GC.@preserve no1 rb1 no2 rb2 @sk(s, subinterval_host, (Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                                                             Ptr{Float64}, Ref{SubintervalOpts}, Ref{Float64}),
                                         _a, _b, no1, rb1, no2, rb2, opts, mx)
      else
        @sk(s, subinterval, (Float64, Float64, Ref{SubintervalOpts}, Ref{Float64}), _a, _b, opts, mx)
      end
      verbose && SpectralKernels.print_panel_convergence(mx[]/abs(k0), _tol, _a, _b)
      if mx[] < config.tol*abs(k0)                                         # quadrature.jl:260
        @sk(s, subinterval_accept, ())
      else                                                                 # quadrature.jl:268-270
        (tl, tr) = (_a == 0) ? (9_tol/10, _tol/10) : (_tol/2, _tol/2)
        (_a < (_a + _b)/2 < _b) || error("sub-interval ($_a, $_b) cannot be split any further")
        push!(stack, (_a, (_a + _b)/2, tl)); push!(stack, ((_a + _b)/2, _b, tr))
      end
    end
    @sk(s, panel_commit, ())                                               # adaptive.jl:163-164
    newhi, rstop = Ref{Int64}(0), Ref(0.0)
    @sk(s, converge_scan, (Ref{ScanArgs}, Ref{Int64}, Ref{Float64}), sa, newhi, rstop)          # adaptive.jl:183-198
    @sk(s, converge_apply, (Ref{ScanArgs}, Int64), sa, newhi[])                                 # adaptive.jl:194
    hi, r_hi = newhi[], rstop[]
  end
  vals = Vector{Float64}(undef, n_out); errs = similar(vals)
  GC.@preserve vals errs @sk(s, results_get, (Ptr{Float64}, Ptr{Float64}), vals, errs)         # adaptive.jl:105-107
  (vals, errs)
end

# src/utils.jl:41-64 on the device: the pairwise lags |pts[i] - pts[j]|, j > i, are formed, sorted and de-duplicated
# by sk_targets_set_pairs (its default pair order -- strict upper triangle, row-major -- is the order of :44-45), and
# the matrix is filled from the flat result.  The reference's sortperm / invperm disappear.
function build_dense_cov_matrix(cfg::AdaptiveKernelConfig{<:OnB200}, pts::AbstractVector{Float64})
  npt = length(pts)
  pv  = convert(Vector{Float64}, pts)
  set_pairs = (s, info) -> begin
    s.group && error("pair lists take a single context")
    GC.@preserve pv ck(s, ccall((:sk_targets_set_pairs, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Int64}, Int64, Ref{TargetInfo}), s.h, pv, npt, 1, C_NULL, 0, info))
    Int(info[].n_in)
  end
  vals = kernel_values(cfg, Float64[]; targets=set_pairs)[1]
  k00  = kernel_values(cfg, [0.0])[1][1]
  M = Matrix{Float64}(undef, npt, npt)
  j = 0
  for i in 1:npt
    M[i, i] = k00
    M[i, i+1:end] .= vals[(j+1):(j+npt-i)]
    M[i+1:end, i] .= vals[(j+1):(j+npt-i)]
    j += npt - i
  end
  M
end

export B200, Session, finufft1d3_b200, SK_SDF_MATERN, SK_SDF_EXPONENTIAL
end # module
