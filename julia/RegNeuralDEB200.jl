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

const ARITH_FMA_CHAIN, ARITH_FIXED24, ARITH_SPLITK = Int32(0), Int32(1), Int32(2)
const DETACH_ALL, DETACH_ALL_BUT_FIRST = Int32(0), Int32(1)      # rnde_set_detach: what the tape differentiates (utils.jl:21-23)
const KERNEL_AUTO, KERNEL_CLUSTER4 = Int32(0), Int32(4)

# ---- what the reference passes, recognised (closures and solver objects cannot cross a C ABI) ----------------------
# solver_args... of the constructor: Tsit5() / AutoTsit5(Tsit5()) objects of OrdinaryDiffEq as the reference's scripts pass
# them (test/test_node.jl:9,65; experiments/mnist_node.jl:81,99,116), or the Symbols :Tsit5 / :AutoTsit5.
function _alg(solver)
    solver === :Tsit5 && return ALG_TSIT5
    solver === :AutoTsit5 && return ALG_AUTO_TSIT5
    name = string(nameof(typeof(solver)))
    name == "Tsit5" && return ALG_TSIT5
    # AutoTsit5(Tsit5()) is a CompositeAlgorithm{Tuple{Tsit5,...},AutoSwitch}: only its first algorithm ever steps here
    (name == "CompositeAlgorithm" && string(nameof(typeof(solver.algs[1]))) == "Tsit5") && return ALG_AUTO_TSIT5
    error("regnde: unsupported solver $(typeof(solver)); the CUDA stepper implements Tsit5() and AutoTsit5(Tsit5())")
end

# The `func(u, t, integrator)` closure of SavingCallback (neural_ode.jl:116,152; mnist_node.jl:67,76-79,88-97;
# test_node.jl:75) is probed ONCE with two mock integrators and matched against the four closures the reference uses;
# anything else errors loudly (the saved values are formed on the device from EEst, dt and eigen_est only).
struct MockIntegrator; EEst::Float32; dt::Float32; eigen_est::Float32; end
const _STAB = 1f0 / 3.5068f0          # 1 / alg_stability_size(Tsit5())   (mnist_node.jl:75,87)
_expected(kind, m) = kind == REG_ERR_DT ? m.EEst * m.dt : kind == REG_STIFF_DT_ABS ? abs(m.eigen_est * m.dt) :
                     kind == REG_STIFF_SCALED ? _STAB * abs(m.eigen_est) : m.EEst * m.dt + 0.1f0 * _STAB * m.eigen_est
function _reg_kind(func)
    func isa Integer && return Int32(func)                       # already an RNDE_REG_* value
    probes = (MockIntegrator(2f0, 3f0, 5f0), MockIntegrator(7f0, 0.5f0, -11f0))
    vals = map(m -> Float32(Tracker.data(func(nothing, 0f0, m))), probes)
    for kind in (REG_ERR_DT, REG_STIFF_DT_ABS, REG_STIFF_SCALED, REG_ERR_PLUS_STIFF)
        all(isapprox(v, _expected(kind, m); rtol = 1f-5) for (v, m) in zip(vals, probes)) && return kind
    end
    error("regnde: the SavingCallback closure is none of the four the reference uses (EEst*dt, abs(eigen_est*dt), ",
          "stability_size*|eigen_est|, EEst*dt + 0.1*stability_size*eigen_est); got $(vals) on the probes")
end

# canonical arithmetic: the split-K stepper where it applies (mirrors resolve_arith of regneuralde/jl_b200/node.py)
_arith(D, H) = (D % 8 == 0 && 128 < D ÷ 4 <= 224 && 4 <= H <= 112) ? ARITH_SPLITK : ARITH_FMA_CHAIN

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
    time_dep || error("2-layer fields are time dependent (TDChain, basic.jl:16-28); Chain fields (time_dep = false) are served with saveat")
    p, re = Flux.destructure(model)          # W1, b1, W2, b2 -- exactly the layout rnde_forward expects
    alg = _alg(solver)
    TrackedNeuralODE{regularize,typeof(model),typeof(p),typeof(re)}(model, p, re, Float32.(tspan), alg, reltol, abstol, time_dep, Dict())
end

function handle!(n::TrackedNeuralODE, D, H, B, reg_kind, need_backward, layers)
    get!(n.handles, (B, reg_kind, need_backward)) do
        cfg = Ref(RndeConfig(sizeof(RndeConfig), D, H, B, _act(layers[1]), _act(layers[2]), Int32(n.time_dep), 0, n.alg, reg_kind, 0, 256,
                             need_backward, 0, 0, 0, 1, n.tspan[1], n.tspan[2], n.abstol, n.reltol, 0f0, 0, 0, B, 0, ntuple(_ -> Int32(0), 8), ntuple(_ -> Int32(0), 8), _arith(D, H), 0, 0))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rnde_create, LIB), Cint, (Ref{RndeConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        # _convert_tspan (utils.jl:21-23) makes tspan tracked whenever p is: the first dt stays on the tape (the library's default,
        # spelled out here because it is the reference's behaviour; DETACH_ALL gives the frozen-step adjoint alone)
        check(ccall((:rnde_set_detach, LIB), Cint, (Ptr{Cvoid}, Int32), h[], DETACH_ALL_BUT_FIRST), h[])
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
function (n::TrackedNeuralODE{true})(x, p = n.p; func = (u, t, integrator) -> integrator.EEst * integrator.dt, tspan = nothing, saveat = nothing)
    saveat === nothing || error("this node was built without saveat: its functor returns the final state only (neural_ode.jl:11)")
    layers = n.model.layers
    D, B = size(x); H = size(layers[1].W, 1)
    h = handle!(n, D, H, B, _reg_kind(func), p isa TrackedArray || x isa TrackedArray, layers)
    ts = tspan === nothing ? n.tspan : tspan          # a per-call tspan never outlives the call (neural_ode.jl:53,58)
    check(ccall((:rnde_set_tspan, LIB), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), h, ts[1], ts[2]), h)
    u, sv, st = _solve_tracked(n, h, x, p)
    return u, Int(data(st).nf), SavedValuesB200(Float32[], sv)
end

# {false, false}: src/models/neural_ode.jl:48-77 -- returns (res, nfe, nothing)
function (n::TrackedNeuralODE{false})(x, p = n.p; func = nothing, tspan = nothing, saveat = nothing)
    layers = n.model.layers
    D, B = size(x); H = size(layers[1].W, 1)
    h = handle!(n, D, H, B, REG_NONE, p isa TrackedArray || x isa TrackedArray, layers)
    ts = tspan === nothing ? n.tspan : tspan
    check(ccall((:rnde_set_tspan, LIB), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), h, ts[1], ts[2]), h)
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
    alg = _alg(solver)
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

function (n::TrackedNeuralODEMulti{R})(x, p = n.p; func = (u, t, integrator) -> integrator.EEst * integrator.dt, tspan = nothing, saveat = nothing) where {R}
    D, B = size(x)
    reg = R ? _reg_kind(func) : REG_NONE
    tracked = p isa TrackedArray || x isa TrackedArray
    h = get!(n.handles, (B, reg, tracked)) do
        hh = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rnde_create, LIB), Cint, (Ref{RndeConfig}, Ref{Ptr{Cvoid}}), Ref(chain_config(n, D, B, reg, tracked)), hh)); hh[]
    end
    ts = tspan === nothing ? n.tspan : tspan
    check(ccall((:rnde_set_tspan, LIB), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), h, ts[1], ts[2]), h)
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


# ---- track / untrack (src/RegNeuralDE.jl:24-25) and solution (neural_ode.jl:182-210) --------------------------------
track(m) = Flux.fmap(x -> x isa AbstractArray ? Tracker.param(x) : x, m)
untrack(m) = Flux.fmap(Tracker.data, m)

# solution(n, x, p): the reference returns the full ODESolution; here the fields its callers read (sol.u[end], sol.destats)
struct SolutionB200; u; t::Vector{Float32}; nf::Int; naccept::Int; nreject::Int; retcode::Symbol; end
function solution(n::TrackedNeuralODE, x, p = n.p; tspan = nothing)
    layers = n.model.layers
    D, B = size(x); H = size(layers[1].W, 1)
    h = handle!(n, D, H, B, REG_NONE, false, layers)
    ts = tspan === nothing ? n.tspan : tspan
    check(ccall((:rnde_set_tspan, LIB), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), h, ts[1], ts[2]), h)
    u, _, st = _solve(n, h, data(x), data(p))
    SolutionB200([u], Float32[ts[2]], st.nf, st.naccept, st.nreject, st.retcode == 0 ? :Success : :Failure)
end

# ---- ClassifierNODE (src/models/supervised_classification.jl:2-46) ---------------------------------------------------
struct ClassifierNODE{N,RE1,RE3,T}
    preode::RE1; node::N; postode::RE3
    p1::T; p2::T; p3::T
end
function ClassifierNODE(preode, node, postode)
    p1, re1 = Flux.destructure(preode)
    p3, re3 = Flux.destructure(postode)
    ClassifierNODE(re1, node, re3, p1, node.p, p3)
end
Flux.trainable(m::ClassifierNODE) = (m.p1, m.p2, m.p3)
function (m::ClassifierNODE)(x, p1 = m.p1, p2 = m.p2, p3 = m.p3; node_kwargs...)
    x = m.preode(p1)(x)
    x, nfe, sv = m.node(x, p2; node_kwargs...)
    return m.postode(p3)(x), nfe, sv
end

# The fused training-step path for the experiment's loss (mnist_node.jl:132-152): forward solve, Dense(784,10) head +
# logitcrossentropy (rnde_head_loss_grad), lambda * agg(sv.saveval) and its cotangents on the device (rnde_reg_agg), reverse
# sweep (rnde_backward): one host synchronisation per step.  Returns (loss, (g1, g2, g3), nfe) like
#   gs = Tracker.gradient((p1,p2,p3) -> loss_function(x, y, model, p1, p2, p3; λ), ps...)   (mnist_node.jl:229-232).
function loss_and_gradient(m::ClassifierNODE, x::CuMatrix{Float32}, y::CuMatrix{Float32}; λ = 1f2, func = (u, t, i) -> i.EEst * i.dt,
                           agg = mean, tspan = nothing)
    n = m.node
    layers = n.model.layers
    D, B = size(x); H = size(layers[1].W, 1); C = size(y, 1)
    reg = _reg_kind(func)
    h = handle!(n, D, H, B, reg, true, layers)
    ts = tspan === nothing ? n.tspan : tspan
    check(ccall((:rnde_set_tspan, LIB), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), h, ts[1], ts[2]), h)
    p2, p3 = data(m.p2), data(m.p3)
    u = similar(x); sv = CUDA.zeros(Float32, 257); dsv = CUDA.zeros(Float32, 257)
    loss = CUDA.zeros(Float32, 1); regv = CUDA.zeros(Float32, 1); logits = CUDA.zeros(Float32, C, B)
    du = similar(x); g3 = similar(p3); g2 = similar(p2)
    st = RndeStats(); s = CUDA.stream().handle
    aggk = agg === mean ? Int32(0) : agg === maximum ? Int32(1) : agg === sum ? Int32(2) : error("agg must be mean, maximum or sum")
    GC.@preserve x y p2 p3 u sv dsv loss regv logits du g3 g2 begin
        check(ccall((:rnde_forward, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{RndeStats}, Ptr{Cvoid}),
                    h, x, p2, u, sv, C_NULL, s), h)
        check(ccall((:rnde_head_loss_grad, LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint, Cfloat, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                    h, u, p3, y, C, 1f0, loss, logits, du, g3, s), h)
        check(ccall((:rnde_reg_agg, LIB), Cint, (Ptr{Cvoid}, Cint, Cfloat, Cfloat, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                    h, aggk, λ, 1f0, sv, dsv, regv, s), h)
        check(ccall((:rnde_backward, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                    h, du, dsv, g2, CU_NULL, s), h)
        check(ccall((:rnde_last_stats, LIB), Cint, (Ptr{Cvoid}, Ref{RndeStats}), h, st), h)
    end
    return Array(loss)[1] + Array(regv)[1], (similar(data(m.p1), 0), g2, g3), Int(st.nf)
end

# update_parameters!(ps, gs, opt) with opt = Optimiser(InvDecay(γ), Momentum(η, ρ)) on raw arrays (src/utils.jl:149-156)
function opt_update!(p::CuVector{Float32}, g::CuVector{Float32}, v::CuVector{Float32}, n_update; γ = 1f-5, η = 0.1f0, ρ = 0.9f0)
    isempty(p) && return p                                                   # utils.jl:151
    GC.@preserve p g v check(ccall((:rnde_opt_update, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Cfloat, Cfloat, Cfloat, Ptr{Cvoid}),
                                   C_NULL, p, g, v, length(p), 1f0 / (1f0 + γ * n_update), η, ρ, CUDA.stream().handle))
    p
end

# ---- TrackedNeuralDSDE (src/models/neural_sde.jl:84-146), forward solves with supplied noise -----------------------
struct RndeSdeConfig
    struct_bytes::Int32; state_dim::Int32; hidden_dim::Int32; batch::Int32; alg::Int32; reg_kind::Int32; max_steps::Int32; max_saved::Int32
    t0::Float32; t1::Float32; abstol::Float32; reltol::Float32
end
mutable struct RndeSdeStats
    nfe1::Int32; nfe2::Int32; naccept::Int32; nreject::Int32; n_saved::Int32; draws::Int32; retcode::Int32; reserved::Int32
    t_final::Float32; dt_init::Float32; dt_last::Float32; reserved2::Float32
    RndeSdeStats() = new(0, 0, 0, 0, 0, 0, 0, 0, 0f0, 0f0, 0f0, 0f0)
end
function sde_create(D, H, B, alg, reg_kind, tspan, abstol, reltol)
    s = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rnde_sde_create, LIB), Cint, (Ref{RndeSdeConfig}, Ref{Ptr{Cvoid}}),
                Ref(RndeSdeConfig(sizeof(RndeSdeConfig), D, H, B, alg, reg_kind, 0, 1024, tspan[1], tspan[2], abstol, reltol)), s))
    s[]
end
# x: D x B, p = vcat(p_drift, p_diffusion), normals: D x B x n_draws standard normals (randn! of the caller's RNG)
function sde_forward(s::Ptr{Cvoid}, x::CuMatrix{Float32}, p::CuVector{Float32}, normals::CuArray{Float32,3})
    u = similar(x); sv = CUDA.zeros(Float32, 1024); st = RndeSdeStats()
    GC.@preserve x p normals u sv begin
        rc = ccall((:rnde_sde_forward, LIB), Cint,
                   (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint, CuPtr{Float32}, CuPtr{Float32}, Ref{RndeSdeStats}, Ptr{Cvoid}),
                   s, x, p, normals, size(normals, 3), u, sv, st, CUDA.stream().handle)
        rc == 0 || error("regnde: ", unsafe_string(ccall((:rnde_sde_last_error, LIB), Cstring, (Ptr{Cvoid},), s)))
    end
    u, Int(st.nfe1), Int(st.nfe2), sv[1:st.n_saved]
end

# Tracker.gradient through the SDE solve (mnist_nsde.jl:201-204): sde_enable_tape once per handle, then sde_backward after a forward
sde_enable_tape(s::Ptr{Cvoid}, cap = 1024) = (ccall((:rnde_sde_enable_tape, LIB), Cint, (Ptr{Cvoid}, Int32), s, cap) == 0 || error("regnde: tape"); s)
function sde_backward(s::Ptr{Cvoid}, du::CuMatrix{Float32}, dsv::CuVector{Float32}, np::Int)
    dp = CUDA.zeros(Float32, np); dx = similar(du)
    GC.@preserve du dsv dp dx begin
        rc = ccall((:rnde_sde_backward, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                   s, du, dsv, dp, dx, CUDA.stream().handle)
        rc == 0 || error("regnde: ", unsafe_string(ccall((:rnde_sde_last_error, LIB), Cstring, (Ptr{Cvoid},), s)))
    end
    dx, dp
end

# ---- TrackedFFJORD (src/models/ffjord.jl:1-137, experiments/ffjord_tabular.jl:47-141) ----------------------------------
# The model is the tabular experiment's MLPDynamics of three ConcatSquashLinear layers with its forw_n_back; p = destructure(model)
# (per layer: layer_W, layer_B, bias_W, bias_B, gate_W -- the order the library expects).  The augmented state [z; delta_logp
# (; |f|^2; |e^T J|^2)] is solved on the device; logpz stays in Julia as in the reference.
struct TrackedFFJORD{R,P}
    p::P
    D::Int; H::Int
    tspan::Vector{Float32}; alg::Int32; reltol::Float32; abstol::Float32
    handles::Dict{Tuple{Int,Int32,Bool},Ptr{Cvoid}}
end
TrackedFFJORD(p, D, H, tspan, regularize, solver = :Tsit5; reltol = 1.4f-8, abstol = 1.4f-8) =
    TrackedFFJORD{regularize,typeof(p)}(p, D, H, Float32.(tspan), _alg(solver), reltol, abstol, Dict())

function handle!(n::TrackedFFJORD{R}, B, extra, need_backward) where {R}
    get!(n.handles, (B, Int32(extra), need_backward)) do
        cfg = Ref(RndeConfig(sizeof(RndeConfig), n.D + extra, n.H, B, 0, 0, 1, 0, n.alg, R ? REG_ERR_DT : REG_NONE, 0, 256, need_backward, 0, 0, 0, 1,
                             n.tspan[1], n.tspan[2], n.abstol, n.reltol, 0f0, 0, 0, B, 0, ntuple(_ -> Int32(0), 8), ntuple(_ -> Int32(0), 8), ARITH_FMA_CHAIN, extra, 0))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rnde_create, LIB), Cint, (Ref{RndeConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        h[]
    end
end

# (n::TrackedFFJORD{R})(x, p, e; regularize) -> (logpx, lambda1, lambda2, nfe, sv)       ffjord.jl:68-137
function (n::TrackedFFJORD{R})(x, p = n.p, e = CUDA.randn(Float32, size(x)...); regularize = false) where {R}
    extra = (regularize && !R) ? 3 : 1                                   # the {true} functor ignores the keyword (ffjord.jl:121)
    B = size(x, 2)
    h = handle!(n, B, extra, istracked(p) || istracked(x))
    ed = data(e)
    GC.@preserve ed check(ccall((:rnde_set_noise, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}), h, ed), h)
    u0 = vcat(x, CUDA.zeros(Float32, extra, B))
    stub = (model = nothing, alg = n.alg)                                 # _solve / _solve_tracked only use the handle
    pred, sv, st = (istracked(p) || istracked(x)) ? _solve_tracked(stub, h, u0, p) : _solve(stub, h, data(u0), data(p))
    z, delta_logp = pred[1:n.D, :], pred[n.D+1, :]
    logpz = vec(sum(-(log(2f0 * Float32(pi)) .+ z .* z) ./ 2, dims = 1))
    l1, l2 = extra == 3 ? (pred[n.D+2, :], pred[n.D+3, :]) : (CUDA.zeros(Float32, B), CUDA.zeros(Float32, B))
    logpz .- delta_logp, l1, l2, Int(st.nf), R ? SavedValuesB200(Float32[], sv) : nothing
end

# sample(n, indims, p; nsamples) (ffjord.jl:160-167): the flow integrated backwards over [tspan[2], tspan[1]]
function sample(n::TrackedFFJORD, indims::Int, p = n.p; nsamples::Int = 1)
    h = handle!(n, nsamples, 1, false)
    e0 = CUDA.zeros(Float32, indims, nsamples)
    GC.@preserve e0 check(ccall((:rnde_set_noise, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float32}), h, e0), h)
    check(ccall((:rnde_set_reverse_time, LIB), Cint, (Ptr{Cvoid}, Int32), h, 1), h)
    u, _, _ = try
        _solve((model = nothing, alg = n.alg), h, vcat(CUDA.randn(Float32, indims, nsamples), CUDA.zeros(Float32, 1, nsamples)), data(p))
    finally
        ccall((:rnde_set_reverse_time, LIB), Cint, (Ptr{Cvoid}, Int32), h, 0)
    end
    u[1:indims, :]
end

# update_parameters!(ps, gs, opt) with opt = Optimiser(WeightDecay(wd), ADAM(η, β)) on raw arrays (ffjord_tabular.jl:128); βp: the running
# powers β.^t of this update, kept by the caller as Flux keeps them in the optimiser state
function adam_update!(p::CuVector{Float32}, g::CuVector{Float32}, m::CuVector{Float32}, v::CuVector{Float32}, βp; wd = 1f-5, η = 1f-2, β = (0.9f0, 0.999f0), ϵ = 1f-8)
    isempty(p) && return p
    GC.@preserve p g m v check(ccall((:rnde_adam_update, LIB), Cint,
        (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Ptr{Cvoid}),
        C_NULL, p, g, m, v, length(p), η, β[1], β[2], βp[1], βp[2], ϵ, wd, CUDA.stream().handle))
    p
end

end # module
