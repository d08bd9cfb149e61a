# SpectralKernelsB200.jl -- the reference-side binding a maintainer of SpectralKernels.jl would add to
# run `kernel_values` on a B200 through libsk_b200.so (C ABI: include/spectralkernels_b200.h).
#
# NOT EXECUTED in this repository's CI: the build image has no Julia.  The identical C ABI is exercised
# through ctypes by spectralkernels.jl_b200/_capi.py + adaptive.py, which is a line-for-line twin of the
# control flow below.  Keep this file thin: scalars only cross the boundary inside the loops.
#
# Usage:
#   using SpectralKernels, SpectralKernelsB200
#   cfg = AdaptiveKernelConfig(S)                       # unchanged public API (src/adaptive.jl:24-59)
#   Ks, errs = kernel_values_b200(cfg, rs)              # same return as kernel_values (src/adaptive.jl:95-108)
#   # or, for the shipped families, without evaluating S on the host:
#   Ks, errs = kernel_values_b200(cfg, rs; builtin=(SK_SDF_MATERN, [phi, rho, nu, 1.0]))
module SpectralKernelsB200

using SpectralKernels
import SpectralKernels: AdaptiveKernelConfig, compute_k0, estimate_tail_decay, updatequadbufs!, quadsz

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

mutable struct Ctx
  h::Ptr{Cvoid}
  function Ctx(device::Integer=0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:sk_ctx_create, libsk), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
    rc == 0 || error("sk_ctx_create failed ($rc): no CUDA device?  There is no CPU fallback.")
    c = new(r[])
    finalizer(x -> ccall((:sk_ctx_destroy, libsk), Cint, (Ptr{Cvoid},), x.h), c)
    c
  end
end

function ck(c::Ctx, rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:sk_last_error, libsk), Cstring, (Ptr{Cvoid},), c.h))
  error("libsk_b200 [$rc]: $msg")     # never throws across the ABI: the C side only returns codes
end

# --- Level 0: replaces src/utils.jl:10 -----------------------------------------------------------------
function finufft1d3_b200(c::Ctx, w::Vector{Float64}, s::Vector{ComplexF64}, x::Vector{Float64})
  out = Vector{ComplexF64}(undef, length(x))
  GC.@preserve w s x out ck(c, ccall((:sk_nufft1d3, libsk), Cint,
      (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{ComplexF64}, Int64, Ptr{Float64}, Ptr{ComplexF64}, Float64),
      c.h, length(w), w, s, length(x), x, out, 1e-15))
  out
end

# --- Level 1: kernel_values with every O(N) pass on the device ---------------------------------------------
function kernel_values_b200(config::AdaptiveKernelConfig, xs::AbstractVector{Float64};
                            k0=compute_k0(config), param_derivative=false, verbose=false,
                            builtin=nothing, ctx::Ctx=Ctx())
  config.dim == 1 || error("dim > 1 is not built yet in libsk_b200")
  c = ctx
  (m, k) = config.quadspec
  lr, jr = config.legrule, config.jacrule
  # QuadRule (src/quadrature.jl:27-47): pass FastGaussQuadrature's own nodes so both paths share them
  GC.@preserve lr jr ck(c, ccall((:sk_rule_set, libsk), Cint,
      (Ptr{Cvoid}, Int32, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
       Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
      c.h, m, k, config.p, lr.no1, lr.wt1, lr.no2, lr.wt2, jr.no1, jr.wt1, jr.no2, jr.wt2))
  if !isnothing(builtin)
    (fam, prm) = builtin
    ck(c, ccall((:sk_sdf_builtin, libsk), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Int32),
                c.h, fam, prm, length(prm), 0))
  end
  xv = convert(Vector{Float64}, xs)
  info = Ref(TargetInfo(0, 0, 0, 0, 0.0, 0.0))
  GC.@preserve xv ck(c, ccall((:sk_targets_set, libsk), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{TargetInfo}),
                              c.h, xv, length(xv), info))               # unique + sort + inverse map (adaptive.jl:99,113-120)
  ck(c, ccall((:sk_run_begin, libsk), Cint, (Ptr{Cvoid},), c.h))
  n   = info[].n_unique
  ix1 = 1
  if info[].has_zero != 0                                                  # adaptive.jl:133-146
    ix1 = 2
    z = config.derivative ? 0.0 : (param_derivative ? compute_k0(config) : k0)
    ck(c, ccall((:sk_zero_lag_set, libsk), Cint, (Ptr{Cvoid}, Float64), c.h, z))
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
  opts       = Ref(SubintervalOpts(config.c, config.p, kernel, config.logw ? 1 : 0, nu, 0, xdiv, C_NULL))
  # (optional optimisation, see sk_subinterval_opts.speculate: evaluate estimate_tail_decay(config, a, b)
  #  before the panel and pass pointer_from_objref/Ref of the ScanArgs with the panel's first sub-interval)
  tau        = config.tol*abs(k0)/2
  while r_hi > 0                                                           # adaptive.jl:149
    (a, b) = (b, b + quadsz(config)/(2*r_hi))                              # adaptive.jl:152
    rlo, rhi = Ref(0.0), Ref(0.0)
    ck(c, ccall((:sk_panel_begin, libsk), Cint, (Ptr{Cvoid}, Int64, Int64, Ref{Float64}, Ref{Float64}),
                c.h, ix1, hi, rlo, rhi))
    # ---- fourier_integrate_interval (quadrature.jl:169-275): scalar control flow only ----
    stack = [(a, b, config.tol)]
    while !isempty(stack)
      (_a, _b, _tol) = pop!(stack)
      mx = Ref(0.0)
      if _a == 0.0 && config.p != 0.0 && config.logw
        # log-weighted origin sub-interval by parts (quadrature.jl:186-228): Julia owns f and df and evaluates both
        # integrands with updatequadbufs!; the device does the transforms (:cis for dim = 1; Bessel orders dim/2-1 and
        # dim/2 for dim = 2) and I = (I0 - A + 2 pi x B)/(dim - alpha)
        dim <= 2 || error("singularity derivative not implemented in d > 2")                      # :222-223
        (f, df) = (config.f, config.df)
        (no1, ba1, no2, ba2) = updatequadbufs!(config.buffers, config.legrule, config.jacrule,
                                               w -> f(w) + w*log(w)*df(w), _a, _b; p=config.p)
        (no1, ra1, no2, ra2) = (copy(no1), real.(ba1), copy(no2), real.(ba2))
        (_, bb1, _, bb2)     = updatequadbufs!(config.buffers, config.legrule, config.jacrule,
                                               w -> w*log(w)*f(w), _a, _b; p=config.p)
        (rb1, rb2) = (real.(bb1), real.(bb2))
        i0   = _b^(dim/2 + 1 - config.alpha)*log(_b)*f(_b)                                          # :189
        lopt = Ref(SubintervalOpts(config.c, config.p, dim == 1 ? SK_KERNEL_COS : SK_KERNEL_BESSEL, 1,
                                   dim == 1 ? 0 : Int64(dim/2 - 1), 0, dim/2 - 1, C_NULL))
        GC.@preserve no1 ra1 rb1 no2 ra2 rb2 ck(c, ccall((:sk_subinterval_logw_host, libsk), Cint,
            (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ref{SubintervalOpts}, Float64, Float64, Ref{Float64}),
            c.h, _a, _b, no1, ra1, rb1, no2, ra2, rb2, lopt, i0, dim - config.alpha, mx))
      elseif isnothing(builtin)
        origin = (_a == 0.0 && config.p != 0.0)
        f = origin ? config.f : (w -> w^config.p * (config.logw ? log(w) : 1) * config.f(w))
        (no1, buf1, no2, buf2) = updatequadbufs!(config.buffers, config.legrule, config.jacrule, f, _a, _b;
                                                 p=(origin ? config.p : 0))
        rb1, rb2 = real.(buf1), real.(buf2)
        GC.@preserve no1 rb1 no2 rb2 ck(c, ccall((:sk_subinterval_host, libsk), Cint,
            (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ref{SubintervalOpts}, Ref{Float64}), c.h, _a, _b, no1, rb1, no2, rb2, opts, mx))
      else
        ck(c, ccall((:sk_subinterval, libsk), Cint, (Ptr{Cvoid}, Float64, Float64, Ref{SubintervalOpts}, Ref{Float64}),
                    c.h, _a, _b, opts, mx))
      end
      if mx[] < config.tol*abs(k0)                                         # quadrature.jl:260
        ck(c, ccall((:sk_subinterval_accept, libsk), Cint, (Ptr{Cvoid},), c.h))
      else                                                                 # quadrature.jl:268-270
        (tl, tr) = (_a == 0) ? (9_tol/10, _tol/10) : (_tol/2, _tol/2)
        (_a < (_a + _b)/2 < _b) || error("sub-interval ($_a, $_b) cannot be split any further")
        push!(stack, (_a, (_a + _b)/2, tl)); push!(stack, ((_a + _b)/2, _b, tr))
      end
    end
    ck(c, ccall((:sk_panel_commit, libsk), Cint, (Ptr{Cvoid},), c.h))      # adaptive.jl:163-164
    (cc, d) = (conv_crit == :panel) ? (NaN, NaN) : estimate_tail_decay(config, a, b, d=config.tail)
    if (isnan(cc) || isnan(d)) && conv_crit != :panel; conv_crit = :panel; end
    sa = conv_crit == :panel ? ScanArgs(0.0, 0.0, (dim+1)/2, tau, SK_CRIT[:panel], 0) :
         ScanArgs(-cc/(d+dim)*b^(d+dim), cc*b^(d+(dim-1)/2), (dim+1)/2, tau, SK_CRIT[conv_crit], 0)
    newhi, rstop = Ref{Int64}(0), Ref(0.0)
    ck(c, ccall((:sk_converge_scan, libsk), Cint, (Ptr{Cvoid}, Ref{ScanArgs}, Ref{Int64}, Ref{Float64}),
                c.h, Ref(sa), newhi, rstop))                               # adaptive.jl:183-198
    ck(c, ccall((:sk_converge_apply, libsk), Cint, (Ptr{Cvoid}, Ref{ScanArgs}, Int64), c.h, Ref(sa), newhi[]))
    hi, r_hi = newhi[], rstop[]
  end
  vals = Vector{Float64}(undef, length(xv)); errs = similar(vals)
  GC.@preserve vals errs ck(c, ccall((:sk_results_get, libsk), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                                     c.h, vals, errs))                     # adaptive.jl:105-107
  (vals, errs)
end

export Ctx, kernel_values_b200, finufft1d3_b200, SK_SDF_MATERN, SK_SDF_EXPONENTIAL
end # module
