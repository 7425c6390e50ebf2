"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_cases.py [case ...]
cases: toy (CTA variant), stream, mnist (cluster-4 forward, tensor-core sweep, tcgen05 weight gradients), ffma4 (FFMA sweep),
cluster8, chain (chain field + saveat), gru, sde.  All backward passes include the first-dt term (a6.cuh), the default."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import regneuralde.jl_b200 as r
from oracle import orc

cases = sys.argv[1:] or ["toy", "stream", "mnist", "ffma4", "cluster8", "chain", "gru", "sde", "ffjord"]
rng = np.random.default_rng(0)


def tdchain(D, H, B, variant, solver, func, act_out=True, cap=16):
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
    node = r.TrackedNeuralODE(model, [0.0, 0.05], True, True, solver, kernel_variant=variant, tape_capacity=cap)
    x = torch.from_numpy(rng.random((D, B), dtype=np.float32)).cuda().requires_grad_(True)
    p = torch.from_numpy(orc.glorot_params(rng, D, H)).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=func)
    (res.sum() + sv.saveval.sum()).backward()
    torch.cuda.synchronize()
    return nfe


for c in cases:
    if c == "toy":
        print(c, tdchain(2, 10, 7, 1, r.Tsit5(), r.ERROR_ESTIMATE, act_out=False))
    elif c == "stream":
        print(c, tdchain(2, 10, 9, 2, r.AutoTsit5(), r.ERROR_PLUS_STIFFNESS, act_out=False))
    elif c == "mnist":
        print(c, tdchain(784, 100, 20, 4, r.AutoTsit5(), r.ERROR_PLUS_STIFFNESS))
    elif c == "ffma4":      # cluster-4 forward with the FFMA sweep (shape outside the tensor-core sweep's range)
        print(c, tdchain(200, 37, 20, 4, r.AutoTsit5(), r.STIFFNESS_ESTIMATE, act_out=False))
    elif c == "cluster8":
        print(c, tdchain(784, 100, 33, 3, r.Tsit5(), r.ERROR_ESTIMATE))
    elif c == "chain":
        W = (50, 20, 50, 20)
        layers, K = [], 20
        for M in W:
            layers.append(r.Dense(K, M, "tanh")); K = M
        node = r.TrackedNeuralODE(r.Chain("tanh", *layers), [0.0, 0.1], False, True, r.Tsit5(), saveat=[0.0, 0.03, 0.1], tape_capacity=16)
        x = torch.randn(20, 6, device="cuda", requires_grad=True); p = node.p.clone().requires_grad_(True)
        res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
        (res.sum() + sv.saveval.sum()).backward(); torch.cuda.synchronize()
        print(c, nfe)
    elif c == "sde":
        D, H, B = 32, 64, 24
        nsde = r.TrackedNeuralDSDE(r.Chain(r.Dense(D, H, "tanh"), r.Dense(H, D, None)), r.Dense(D, D, None), [0.0, 0.2], True, r.SOSRI())
        x = torch.randn(D, B, device="cuda")
        p = nsde.p.clone().requires_grad_(True)
        out = nsde(x, p, func=r.ERROR_ESTIMATE)
        (out[0].sum() + out[3].saveval.sum()).backward()          # forward with the tape, then the reverse sweep (sde_bwd.cuh)
        torch.cuda.synchronize()
        print(c, out[1], out[2])
    elif c == "ffjord":     # forward + reverse sweep through the hand-differentiated ConcatSquash field, then the sampler
        ff = r.TrackedFFJORD(r.CSQDynamics(6, 12), [0.0, 0.1], True, True, r.Tsit5(), tape_capacity=16)
        x = torch.randn(6, 9, device="cuda"); p = ff.p.clone().requires_grad_(True)
        logpx, l1, l2, nfe, sv = ff(x, p, torch.randn(6, 9, device="cuda"))
        (logpx.sum() + sv.saveval.sum()).backward(); torch.cuda.synchronize()
        r.ffjord.sample(ff, 6, ff.p, nsamples=5); torch.cuda.synchronize()
        print(c, nfe)
    elif c == "gru":
        gru = r.LatentGRU(5, 6, 4)
        x = torch.randn(11, 5, 6, device="cuda"); p = gru.p.clone().requires_grad_(True)
        gru(x, p).sum().backward(); torch.cuda.synchronize()
        print(c, "ok")
