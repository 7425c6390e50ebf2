# dump_reference.jl -- run the UNMODIFIED reference path (RegNeuralDE.jl @ its Manifest.toml) on given inputs and
# dump everything the parity tests compare, as .npy (NPZ.jl is a dependency of the reference, Project.toml:25).
# NOT EXECUTED in the build image (no Julia); shipped so that anyone with Julia 1.5 + the reference's Manifest can
# pin the oracle (SURVEY.md 8c vi):   julia --project=/path/to/RegNeuralDE.jl julia/dump_reference.jl in_dir out_dir
#
# in_dir:  x.npy (D x B Float32), p.npy (Flux.destructure order), optional w.npy / ws.npy (loss weights)
# out_dir: u.npy, saveval.npy, nfe.npy, dp.npy, dx.npy
using RegNeuralDE, OrdinaryDiffEq, Flux, CUDA, Tracker, NPZ

in_dir, out_dir = ARGS[1], ARGS[2]
x = npzread(joinpath(in_dir, "x.npy")) |> gpu
p = npzread(joinpath(in_dir, "p.npy")) |> gpu
D, B = size(x)
H = (length(p) - D) ÷ (2D + 2)                      # H(D+1) + H + D(H+1) + D
act_out = get(ENV, "ACT_OUT", "tanh") == "tanh" ? CUDA.tanh : identity
model = TDChain(Dense(D + 1, H, CUDA.tanh), Dense(H + 1, D, act_out)) |> track |> gpu
solver = get(ENV, "SOLVER", "Tsit5") == "AutoTsit5" ? AutoTsit5(Tsit5()) : Tsit5()
node = TrackedNeuralODE(model, [0.0f0, 1.0f0], true, true, solver; save_everystep = false,
                        reltol = 1.4f-8, abstol = 1.4f-8, save_start = false)
func = get(ENV, "FUNC", "err") == "err" ? ((u, t, int) -> int.EEst * int.dt) : ((u, t, int) -> abs(int.eigen_est * int.dt))

w = isfile(joinpath(in_dir, "w.npy")) ? gpu(npzread(joinpath(in_dir, "w.npy"))) : CUDA.ones(Float32, D, B)
pt, xt = Tracker.param(p), Tracker.param(x)
res, nfe, sv = node(xt, pt; func = func)
ws = isfile(joinpath(in_dir, "ws.npy")) ? npzread(joinpath(in_dir, "ws.npy")) : ones(Float32, length(sv.saveval))
loss = sum(w .* res) + sum(ws .* sv.saveval)
Tracker.back!(loss)

mkpath(out_dir)
npzwrite(joinpath(out_dir, "u.npy"), Array(Tracker.data(res)))
npzwrite(joinpath(out_dir, "saveval.npy"), Float32.(Tracker.data.(sv.saveval)))
npzwrite(joinpath(out_dir, "nfe.npy"), [nfe])
npzwrite(joinpath(out_dir, "dp.npy"), Array(Tracker.grad(pt)))
npzwrite(joinpath(out_dir, "dx.npy"), Array(Tracker.grad(xt)))
