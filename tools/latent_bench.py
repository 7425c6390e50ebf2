"""Timing of the Latent-ODE row (BASELINE.json configs[3]) on one GPU, PhysioNet-shaped synthetic
batch (37 features, 49 irregular observation times, batch 512; experiments/latent_ode.jl:105-150,226-262):
  (1) the generator ODE solve alone (chain field + saveat), forward + backward;
  (2) the whole training step loss_function(...) -> backward: GRU encoder, rec_to_gen, solve, decoder, likelihood;
  (3) the CPU restatement of (2)'s two hot loops on the host cores (torch GRU oracle + C oracle solve).
Prints one JSON line (developer evidence for profiles/, not the bench.py headline)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import regneuralde.jl_b200 as r  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = 30
rng = np.random.default_rng(1234)
I, T = 37, 49
times = np.unique(np.concatenate([[0.0], np.sort(rng.random(T - 2)), [1.0]]).astype(np.float32))
S = len(times)
data = rng.standard_normal((I, S, B)).astype(np.float32)
mask = (rng.random((I, S, B)) < 0.15).astype(np.float32)
trow = np.broadcast_to(times[None, :, None], (1, S, B)).astype(np.float32).copy()
out = {"workload": "latent_ode physionet-shaped", "batch": B, "features": I, "n_saveat": S}


def timed(fn):
    tot = 0.0
    for it in range(iters + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 3:
            tot += e0.elapsed_time(e1)
    return tot / iters


for name, reg, solver, func in [("vanilla", False, r.Tsit5(), None), ("error_est", True, r.Tsit5(), r.ERROR_ESTIMATE),
                                ("error_stiff_est", True, r.AutoTsit5(), r.ERROR_PLUS_STIFFNESS)]:
    gen = torch.Generator().manual_seed(5)
    model = r.latent_ode_model(I, 40, 50, 20, 50, saveat=times.tolist(), regularize=reg, solver=solver, generator=gen)
    ps = [p.clone().requires_grad_(True) for p in model.trainable()]
    d, m, t = (torch.from_numpy(a).cuda() for a in (data * mask, mask, trow))
    sample = torch.randn(20, B, device="cuda")
    info = {}

    def step():
        total, nfe, parts = r.loss_function(d, m, t, model, *ps, func=func, regularize=reg, lam_r=1.0e3, sample=sample)
        total.backward()
        info["nfe"] = nfe
        for p in ps:
            p.grad = None

    ms = timed(step)
    # solve alone
    z0 = torch.randn(20, B, device="cuda").requires_grad_(True)
    w = torch.randn(20, S, B, device="cuda")

    def solve_only():
        res, nfe, sv = model.node(z0, ps[2], func=func)
        ((res * w).sum() + (1e3 * sv.saveval.mean() if reg else 0.0)).backward()
        z0.grad = None; ps[2].grad = None

    ms_solve = timed(solve_only)
    x = torch.cat([d, m, t], 0)
    ms_gru = timed(lambda: (model.rnn(x, ps[0]).sum().backward(), setattr(ps[0], "grad", None)))
    out[name] = {"nfe": info["nfe"], "naccept": model.node.last_stats.naccept, "step_ms": ms, "solve_fwd_bwd_ms": ms_solve,
                 "gru_fwd_bwd_ms": ms_gru, "samples_per_s": B / (ms * 1e-3)}

# CPU restatement of the two hot loops (error_est), all host cores
from oracle import gru_oracle as G, orc  # noqa: E402
W = (50, 20, 50, 20, 50, 20, 50, 20)
torch.set_num_threads(os.cpu_count())
p1 = torch.tensor(G.glorot_params(rng, I, 40, 50), requires_grad=True)
x_cpu = torch.tensor(np.concatenate([data * mask, mask, trow], 0))
t0 = time.time()
o = G.forward(p1, x_cpu, I, 40, 50); o.sum().backward()
t_gru = time.time() - t0
cfg = orc.OracleConfig(D=20, H=50, B=B, reg_kind=orc.REG_ERR_DT, kblock1=20, widths=W, acts=(1,) * 8, pre_act=1, saveat=times.astype(np.float64))
oc = orc.Oracle(cfg)
pz = orc.glorot_chain_params(rng, 20, W)
z = rng.standard_normal((20, B)).astype(np.float32)
oc.forward(z, pz)
t0 = time.time()
ref = oc.forward(z, pz)
oc.backward(np.zeros((20, B), np.float32), np.ones(len(ref.saveval), np.float32), dusave=np.ones(ref.usave.shape, np.float32))
t_solve = time.time() - t0
out["cpu_port"] = {"cores": os.cpu_count(), "gru_fwd_bwd_ms": t_gru * 1e3, "solve_fwd_bwd_ms": t_solve * 1e3, "nfe": ref.nf,
                   "samples_per_s": B / (t_gru + t_solve)}
print(json.dumps(out))
