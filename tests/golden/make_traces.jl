# make_traces.jl -- dumps the adaptive panel traces of the UNMODIFIED reference (pbeckman/SpectralKernels.jl) for the
# five BASELINE.json configurations, so that `tests/test_reference_traces.py` can pin the product's and the oracle's
# traces to the reference itself ("panel counts and indexing must be bit-exact").
#
# It cannot run in the build image (no Julia, no FINUFFT); run it once on any machine with Julia >= 1.9:
#
#     julia --project=/path/to/SpectralKernels.jl tests/golden/make_traces.jl tests/golden/reference_traces.json
#
# The project needs SpectralKernels and JSON.  No random numbers are drawn in Julia: every distance set is either
# deterministic or read from the raw float64 files `python tests/golden/make_trace_inputs.py` writes under
# tests/golden/trace_inputs/, so Julia and numpy see bit-identical distances.
#
# How the trace is taken without touching the reference's source: `kernel_values(cfg, xs; verbose=true)` calls two
# printing helpers with the exact Float64 quantities of the adaptive loop,
#     print_panel_info(xs, highest_unconv_ix, a, b)                 src/utils.jl:12, called at src/adaptive.jl:154
#     print_panel_convergence(max_I_error/k0, _tol, _a, _b)         src/utils.jl:17, called at src/quadrature.jl:259
# This script replaces those two helpers by recorders.  Nothing else of the package is altered.
using SpectralKernels, JSON

const TRACE = Vector{Any}()

@eval SpectralKernels function print_panel_info(xs, highest_unconv_ix, a, b)
  push!(Main.TRACE, Dict("kind" => "panel", "a" => a, "b" => b, "hi_before" => highest_unconv_ix,
                         "r_hi" => xs[highest_unconv_ix]))
  nothing
end
@eval SpectralKernels function print_panel_convergence(max_I_error, tol, _a, _b)
  push!(Main.TRACE, Dict("kind" => "subinterval", "a" => _a, "b" => _b, "rel_err" => max_I_error, "split_tol" => tol))
  nothing
end

matern_sdf(w, parms; d=1) = parms[1]*(parms[2]^2 + w^2)^(-parms[3] - d/2)      # scripts/matern_pair.jl:17

readvec(path) = reinterpret(Float64, read(path))                                   # raw little-endian float64

function run_case(name, cfg, xs; kw...)
  empty!(TRACE)
  redirect_stdout(devnull) do
    (vals, errs) = kernel_values(cfg, xs; verbose=true, kw...)
    global LAST = (vals, errs)
  end
  # accepted <=> max_I_error < config.tol*k0 (src/quadrature.jl:260); with rel_err = max_I_error/k0 recorded
  for t in TRACE
    t["kind"] == "subinterval" && (t["accepted"] = t["rel_err"] < cfg.tol)
  end
  # hi_after of a panel is the hi_before of the next one; 0 after the last (src/adaptive.jl:149)
  pans = [t for t in TRACE if t["kind"] == "panel"]
  for (i, t) in enumerate(pans)
    t["hi_after"] = i < length(pans) ? pans[i+1]["hi_before"] : 0
  end
  idx = unique(round.(Int, range(1, length(xs), length=min(length(xs), 64))))
  Dict("name" => name, "n" => length(xs), "trace" => deepcopy(TRACE),
       "sample_index" => idx .- 1, "sample_values" => LAST[1][idx], "sample_errors" => LAST[2][idx])
end

function main(out)
  dir = joinpath(@__DIR__, "trace_inputs")
  cases = Any[]
  # config 1: README demo (README.md:19-33)
  rs = 10 .^ range(-6, 0, length=1000)
  push!(cases, run_case("config1_readme", AdaptiveKernelConfig(w -> (1 + w^2)^(-2); tol=1e-8), collect(rs)))
  # config 2 at reduced and full size: Matern nu = 1.5, rho = 1, K(0) = 1; distances written by make_trace_inputs.py
  S2 = w -> matern_sdf(w, (1/(pi/2), 1.0, 1.5))
  for f in ("config2_2e3.f64", "config2_1e7.f64")
    isfile(joinpath(dir, f)) || continue
    push!(cases, run_case(f, AdaptiveKernelConfig(S2; tol=1e-8), Vector(readvec(joinpath(dir, f))); k0=1.0))
  end
  # config 3: singular Matern alpha = 0.5 on pairwise distances (1-D kernel and dim = 2)
  for (f, dim) in (("config3_lags.f64", 1), ("config3_lags.f64", 2))
    isfile(joinpath(dir, f)) || continue
    S3 = w -> matern_sdf(w, (1.0, 1.0, 1.5); d=dim)
    push!(cases, run_case("config3_dim$(dim)", AdaptiveKernelConfig(S3; tol=1e-8, alpha=0.5, dim=dim),
                          Vector(readvec(joinpath(dir, f)))))
  end
  # config 4: derivative config (K') and one parameter-derivative integrand on the same lags
  f4 = joinpath(dir, "config4_1e6.f64")
  if isfile(f4)
    xs = Vector(readvec(f4))
    cfg = AdaptiveKernelConfig(S2; tol=1e-8)
    push!(cases, run_case("config4_K", cfg, xs; k0=1.0))
    push!(cases, run_case("config4_dK", SpectralKernels.gen_derivative_config(cfg), xs; k0=1.0))
    dnu = w -> -(1/(pi/2))*(1 + w^2)^(-2)*log(1 + w^2)
    push!(cases, run_case("config4_dnu", SpectralKernels.gen_new_sdf_config(cfg, dnu), xs; k0=1.0, param_derivative=true))
  end
  # config 5: Vecchia pair-list lags, dim = 2
  f5 = joinpath(dir, "config5_lags.f64")
  if isfile(f5)
    S5 = w -> matern_sdf(w, (1.0, 4.0, 1.5); d=2)
    push!(cases, run_case("config5_dim2", AdaptiveKernelConfig(S5; tol=1e-8, dim=2), Vector(readvec(f5))))
  end
  open(out, "w") do io
    JSON.print(io, Dict("generator" => "tests/golden/make_traces.jl", "cases" => cases), 1)
  end
end

main(length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "reference_traces.json"))
