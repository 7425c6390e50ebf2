# RegNeuralDEB200.jl -- drop-in Julia binding of libregnde.so for the neural-ODE hot path of
# avik-pal/RegNeuralDE.jl.  NOT EXECUTED in the build image (no Julia toolchain there); it is the
# reference-side stub a maintainer adds, kept 1:1 with the ctypes binding the tests do execute
# (regneuralde/jl_b200/_lib.py, node.py).  See INTEGRATION.md.
#
# It keeps the reference's call surface (src/models/neural_ode.jl:10, :48, :110):
#     node = TrackedNeuralODE(model, tspan, time_dep, regularize, solver; reltol, abstol, ...)
#     res, nfe, sv = node(x, p; func = ..., tspan = ...)
# and replaces only the body: instead of ODEProblem + solve(...; sensealg = SensitivityADPassThrough())
# traced by Tracker, one ccall runs the fused sm_100a stepper, and a Tracker custom gradient calls the
# reverse sweep.
module RegNeuralDEB200

using CUDA, Flux, Tracker
using Tracker: TrackedArray, data, track, @grad

const LIB = get(ENV, "REGNDE_LIB", joinpath(@__DIR__, "..", "regneuralde", "jl_b200", "libregnde.so"))

# ---- mirror of include/regnde.h ------------------------------------------------------------------
struct RndeConfig
    struct_bytes::Int32
    state_dim::Int32; hidden_dim::Int32; batch::Int32
    act_hidden::Int32; act_out::Int32; time_dep::Int32; kblock::Int32
    alg::Int32; reg_kind::Int32; max_steps::Int32; tape_capacity::Int32
    need_backward::Int32; kernel_variant::Int32; dist_mode::Int32
    rank::Int32; nranks::Int32
    t0::Float32; t1::Float32; abstol::Float32; reltol::Float32; dtmin::Float32
    max_saveat::Int32; n_layers::Int32
    global_batch::Int64
    pre_act::Int32; layer_width::NTuple{8,Int32}; layer_act::NTuple{8,Int32}; arith::Int32
    csq_extra::Int32; reserved0::Int32
end

mutable struct RndeStats
    nf::Int32; naccept::Int32; nreject::Int32; n_saved::Int32; retcode::Int32
    t_final::Float32; dt_last::Float32; dt_init::Float32
    RndeStats() = new(0, 0, 0, 0, 0, 0f0, 0f0, 0f0)
end

const ACT_IDENTITY, ACT_TANH = Int32(0), Int32(1)
const ALG_TSIT5, ALG_AUTO_TSIT5 = Int32(0), Int32(1)
# the `func` closures of the reference, enumerated (closures cannot cross a C ABI):
const REG_NONE = Int32(0)
const REG_ERR_DT = Int32(1)          # (u,t,int) -> int.EEst * int.dt               neural_ode.jl:116
const REG_STIFF_DT_ABS = Int32(2)    # (u,t,int) -> abs(int.eigen_est * int.dt)     test/test_node.jl:75
const REG_STIFF_SCALED = Int32(3)    # stability_size * |eigen_est|                 mnist_node.jl:76-79
const REG_ERR_PLUS_STIFF = Int32(4)  # EEst*dt + 0.1*stability_size*eigen_est       mnist_node.jl:88-97

check(rc, h = C_NULL) = rc == 0 || error("regnde: ", unsafe_string(ccall((:rnde_status_string, LIB), Cstring, (Cint,), rc)),
                                         h == C_NULL ? "" : " -- " * unsafe_string(ccall((:rnde_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

# ---- the layer ------------------------------------------------------------------------------------
struct SavedValuesB200{T}
    t::Vector{Float32}
    saveval::T                 # tracked CuVector: agg (mean / maximum / sum) and λ stay in Julia, as in the reference
end

mutable struct TrackedNeuralODE{R,M,P,RE}
    model::M
    p::P
    re::RE
    tspan::Vector{Float32}
    alg::Int32
    reltol::Float32
    abstol::Float32
    time_dep::Bool
    handles::Dict{Tuple{Int,Int32,Bool},Ptr{Cvoid}}
end

_act(l::Dense) = l.σ === identity ? ACT_IDENTITY : (l.σ === tanh || l.σ === CUDA.tanh) ? ACT_TANH : error("unsupported activation")

function TrackedNeuralODE(model, tspan, time_dep, regularize, solver = :Tsit5; reltol = 1.4f-8, abstol = 1.4f-8, kwargs...)
    get(kwargs, :save_everystep, false) && error("save_everystep = true has no call site in the reference; use saveat")
    haskey(kwargs, :saveat) && return TrackedNeuralODEMulti(model, tspan, time_dep, regularize, solver; reltol = reltol, abstol = abstol, kwargs...)
    p, re = Flux.destructure(model)          # W1, b1, W2, b2 -- exactly the layout rnde_forward expects
    alg = solver === :AutoTsit5 ? ALG_AUTO_TSIT5 : ALG_TSIT5
    TrackedNeuralODE{regularize,typeof(model),typeof(p),typeof(re)}(model, p, re, Float32.(tspan), alg, reltol, abstol, time_dep, Dict())
end

function handle!(n::TrackedNeuralODE, D, H, B, reg_kind, need_backward, layers)
    get!(n.handles, (B, reg_kind, need_backward)) do
        cfg = Ref(RndeConfig(sizeof(RndeConfig), D, H, B, _act(layers[1]), _act(layers[2]), 1, 0, n.alg, reg_kind, 0, 256,
                             need_backward, 0, 0, 0, 1, n.tspan[1], n.tspan[2], n.abstol, n.reltol, 0f0, 0, 0, B, 0, ntuple(_ -> Int32(0), 8), ntuple(_ -> Int32(0), 8), 0, 0, 0))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rnde_create, LIB), Cint, (Ref{RndeConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        h[]
    end
end

# raw solve on plain CuArrays: (res, saveval, stats)
function _solve(n::TrackedNeuralODE, h, x::CuMatrix{Float32}, p::CuVector{Float32})
    u = similar(x)
    sv = CUDA.zeros(Float32, 257)
    st = RndeStats()
    GC.@preserve x p u sv begin
        check(ccall((:rnde_forward, LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ref{RndeStats}, Ptr{Cvoid}),
                    h, x, p, u, sv, st, CUDA.stream().handle), h)
    end
    u, sv[1:st.n_saved], st
end

# Tracker custom gradient: what Tracker.gradient(...) differentiates (experiments/mnist_node.jl:229-232)
_solve_tracked(n, h, x, p) = track(_solve_tracked, n, h, x, p)
@grad function _solve_tracked(n, h, x, p)
    u, sv, st = _solve(n, h, data(x), data(p))
    (u, sv, st), function (Δ)
        du, dsv = Δ[1], Δ[2]
        dsvfull = CUDA.zeros(Float32, 257); dsv === nothing || (dsvfull[1:length(dsv)] .= dsv)
        dp = similar(data(p)); dx = similar(data(x))
        GC.@preserve du dsvfull dp dx check(ccall((:rnde_backward, LIB), Cint,
            (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
            h, du === nothing ? CUDA.zeros(Float32, size(u)) : du, dsvfull, dp, dx, CUDA.stream().handle), h)
        (nothing, nothing, dx, dp)
    end
end

# {regularize = true, return_multiple = false}: src/models/neural_ode.jl:110-144
function (n::TrackedNeuralODE{true})(x, p = n.p; func = REG_ERR_DT, tspan = nothing, saveat = nothing)
    saveat === nothing || error("this node was built without saveat: its functor returns the final state only (neural_ode.jl:11)")
    layers = n.model.layers
    D, B = size(x); H = size(layers[1].W, 1)
    h = handle!(n, D, H, B, Int32(func), p isa TrackedArray || x isa TrackedArray, layers)
    tspan === nothing || check(ccall((:rnde_set_tspan, LIB), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), h, tspan[1], tspan[2]), h)
    u, sv, st = _solve_tracked(n, h, x, p)
    return u, Int(data(st).nf), SavedValuesB200(Float32[], sv)
end

# {false, false}: src/models/neural_ode.jl:48-77 -- returns (res, nfe, nothing)
function (n::TrackedNeuralODE{false})(x, p = n.p; func = nothing, tspan = nothing, saveat = nothing)
    layers = n.model.layers
    D, B = size(x); H = size(layers[1].W, 1)
    h = handle!(n, D, H, B, REG_NONE, p isa TrackedArray || x isa TrackedArray, layers)
    tspan === nothing || check(ccall((:rnde_set_tspan, LIB), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), h, tspan[1], tspan[2]), h)
    u, _, st = _solve_tracked(n, h, x, p)
    return u, Int(data(st).nf), nothing
end

# ---- multi-save functors {R,true}: saveat + Tsit5 dense output, chain fields (time_dep = false) -------------------
# src/models/neural_ode.jl:79-108,146-180; call site src/models/time_series.jl:51 with
# gen_dynamics = Chain(x -> tanh.(x), Dense(20,50,tanh), ... x8)   experiments/latent_ode.jl:109-147
mutable struct TrackedNeuralODEMulti{R,M,P,RE}
    model::M; p::P; re::RE; tspan::Vector{Float32}; alg::Int32; reltol::Float32; abstol::Float32; time_dep::Bool
    saveat::Vector{Float32}
    handles::Dict{Tuple{Int,Int32,Bool},Ptr{Cvoid}}
end
function TrackedNeuralODEMulti(model, tspan, time_dep, regularize, solver; reltol, abstol, saveat, kwargs...)
    p, re = Flux.destructure(model)
    alg = solver === :AutoTsit5 ? ALG_AUTO_TSIT5 : ALG_TSIT5
    TrackedNeuralODEMulti{regularize,typeof(model),typeof(p),typeof(re)}(model, p, re, Float32.(tspan), alg, reltol, abstol, time_dep, Float32.(saveat), Dict())
end

# config of a Chain of Dense layers (optionally led by x -> tanh.(x)) evaluated as re(p)(u)
function chain_config(n, D, B, reg_kind, need_backward)
    ls = collect(n.model.layers)
    pre = ls[1] isa Dense ? ACT_IDENTITY : ACT_TANH
    ds = filter(l -> l isa Dense, ls)
    w = ntuple(i -> i <= length(ds) ? Int32(size(ds[i].W, 1)) : Int32(0), 8)
    a = ntuple(i -> i <= length(ds) ? _act(ds[i]) : Int32(0), 8)
    RndeConfig(sizeof(RndeConfig), D, 0, B, 0, 0, 0, 0, n.alg, reg_kind, 0, 256, need_backward, 0, 0, 0, 1,
               n.tspan[1], n.tspan[2], n.abstol, n.reltol, 0f0, 64 * cld(length(n.saveat), 64), length(ds), B, pre, w, a, 0, 0, 0)
end

function (n::TrackedNeuralODEMulti{R})(x, p = n.p; func = REG_ERR_DT, tspan = nothing, saveat = nothing) where {R}
    D, B = size(x)
    reg = R ? Int32(func) : REG_NONE
    tracked = p isa TrackedArray || x isa TrackedArray
    h = get!(n.handles, (B, reg, tracked)) do
        hh = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rnde_create, LIB), Cint, (Ref{RndeConfig}, Ref{Ptr{Cvoid}}), Ref(chain_config(n, D, B, reg, tracked)), hh)); hh[]
    end
    times = saveat === nothing ? n.saveat : Float32.(saveat)          # update_saveat! semantics (neural_ode.jl:35-45)
    check(ccall((:rnde_set_saveat, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Cint), h, times, length(times)), h)
    res, sv, st = _solve_saveat_tracked(h, x, p, length(times))      # res: feat x nsave x batch
    return res, Int(data(st).nf), R ? SavedValuesB200(Float32[], sv) : nothing
end

function _solve_saveat(h, x::CuMatrix{Float32}, p::CuVector{Float32}, nsave)
    res = CUDA.zeros(Float32, size(x, 1), nsave, size(x, 2)); sv = CUDA.zeros(Float32, 257); st = RndeStats()
    GC.@preserve x p res sv check(ccall((:rnde_forward_saveat, LIB), Cint,
        (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ref{RndeStats}, Ptr{Cvoid}),
        h, x, p, CU_NULL, res, sv, st, CUDA.stream().handle), h)
    res, sv[1:st.n_saved], st
end
_solve_saveat_tracked(h, x, p, nsave) = track(_solve_saveat_tracked, h, x, p, nsave)
@grad function _solve_saveat_tracked(h, x, p, nsave)
    res, sv, st = _solve_saveat(h, data(x), data(p), nsave)
    (res, sv, st), function (Δ)
        dres, dsv = Δ[1], Δ[2]
        dsvfull = CUDA.zeros(Float32, 257); dsv === nothing || (dsvfull[1:length(dsv)] .= dsv)
        dp = similar(data(p)); dx = similar(data(x))
        GC.@preserve dres dsvfull dp dx check(ccall((:rnde_backward_saveat, LIB), Cint,
            (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
            h, CU_NULL, dres, dsvfull, dp, dx, CUDA.stream().handle), h)
        (nothing, dx, dp, nothing)
    end
end

# ---- (p::LatentGRU)(x): experiments/latent_ode.jl:39-99 ------------------------------------------------------------
struct RndeGruConfig
    struct_bytes::Int32; in_dim::Int32; hidden_dim::Int32; latent_dim::Int32; batch::Int32; seq_len::Int32; need_backward::Int32; reserved::Int32
end
function gru_forward(g::Ptr{Cvoid}, x::CuArray{Float32,3}, p::CuVector{Float32}, latent_dim)     # x: (2I+1) x T x B as the reference builds it
    out = CUDA.zeros(Float32, 2 * latent_dim, size(x, 3))
    GC.@preserve x p out check(ccall((:rnde_gru_forward, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                                     g, x, p, out, CUDA.stream().handle))
    out
end
gru_tracked(g, x, p, L) = track(gru_tracked, g, x, p, L)
@grad function gru_tracked(g, x, p, L)
    out = gru_forward(g, data(x), data(p), L)
    out, function (Δ)
        dp = similar(data(p))
        GC.@preserve Δ dp check(ccall((:rnde_gru_backward, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}), g, Δ, dp, CUDA.stream().handle))
        (nothing, nothing, dp, nothing)
    end
end

end # module
