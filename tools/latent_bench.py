"""Timing of the Latent-ODE solver half (BASELINE.json configs[3]) on one GPU:
TrackedNeuralODE(gen_dynamics, [0,1], false, REGULARIZE, solver, saveat = 49 irregular times) at batch 512
(experiments/latent_ode.jl:109-147), forward + backward of sum(w .* res) + mean(sv.saveval), CUDA events.
Prints one JSON line (developer evidence for profiles/, not the bench.py headline)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import regneuralde.jl_b200 as r  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = 30
rng = np.random.default_rng(1234)
W = (50, 20, 50, 20, 50, 20, 50, 20)
layers, K = [], 20
for M in W:
    layers.append(r.Dense(K, M, "tanh")); K = M
model = r.Chain("tanh", *layers)
saveat = np.unique(np.concatenate([[0.0], np.sort(rng.random(47)), [1.0]]).astype(np.float32)).tolist()
out = {"workload": "latent_ode_solve", "batch": B, "n_saveat": len(saveat)}
for name, reg, solver, func in [("vanilla", False, r.Tsit5(), None), ("error_est", True, r.Tsit5(), r.ERROR_ESTIMATE),
                                ("error_stiff_est", True, r.AutoTsit5(), r.ERROR_PLUS_STIFFNESS)]:
    node = r.TrackedNeuralODE(model, [0.0, 1.0], False, reg, solver, saveat=saveat, reltol=1.4e-8, abstol=1.4e-8)
    x = torch.from_numpy(rng.standard_normal((20, B)).astype(np.float32)).cuda().requires_grad_(True)
    p = node.p.clone().requires_grad_(True)
    w = torch.randn(20, len(saveat), B, device="cuda")
    tf = tb = 0.0
    for it in range(iters + 3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        res, nfe, sv = node(x, p, func=func)
        e1.record()
        loss = (res * w).sum() + (100.0 * sv.saveval.mean() if reg else 0.0)
        loss.backward()
        e2.record()
        torch.cuda.synchronize()
        if it >= 3:
            tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
        p.grad = None; x.grad = None
    st = node.last_stats
    out[name] = {"nfe": nfe, "naccept": st.naccept, "fwd_ms": tf / iters, "bwd_ms": tb / iters,
                 "samples_per_s": B / ((tf + tb) / iters * 1e-3), "variant": int(node._handles[next(iter(node._handles))].lib.rnde_kernel_variant(node._handles[next(iter(node._handles))].h))}
print(json.dumps(out))
