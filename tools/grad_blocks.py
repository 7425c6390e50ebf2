"""Developer diagnostic: per parameter block (W1, b1, W2, b2) error of the CUDA gradient vs the CPU Float32 adjoint, toy stiffness case."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import regneuralde.jl_b200 as r
from oracle import orc
D, H, B = 2, 10, 5
rng = np.random.default_rng(7)
p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D))
for variant in (0, 2):
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, True, r.AutoTsit5(), reltol=1.4e-8, abstol=1.4e-8, kernel_variant=variant)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True); x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=r.STIFFNESS_ESTIMATE)
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_ID, alg=1, reg_kind=orc.REG_STIFF_DT_ABS, kblock1=D))
    ref = o.forward(x_np, p_np)
    rng2 = np.random.default_rng(7); rng2.random((D, B)); 
    w = np.random.default_rng(3).standard_normal((D, B)).astype(np.float32); ws = np.random.default_rng(4).standard_normal(len(ref.saveval)).astype(np.float32)
    ((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
    hi, hx, _, _ = o.backward(w, ws, hi=True); c32, cx, _, _ = o.backward(w, ws)
    g = p.grad.cpu().numpy()
    blocks = {"W1": (0, H * (D + 1)), "b1": (H * (D + 1), H * (D + 1) + H), "W2": (H * (D + 1) + H, H * (D + 1) + H + D * (H + 1)), "b2": (H * (D + 1) + H + D * (H + 1), len(g))}
    print(f"variant {variant}: nfe {nfe} naccept {node.last_stats.naccept}; max|g| {np.abs(hi).max():.3e}")
    for k, (a, b) in blocks.items():
        print(f"   {k}: cuda err {np.abs(g[a:b] - hi[a:b]).max():.3e}  cpu32 err {np.abs(c32[a:b] - hi[a:b]).max():.3e}  block max {np.abs(hi[a:b]).max():.3e}")
    print("   time column W1t:", g[H * D:H * (D + 1)][:4], hi[H * D:H * (D + 1)][:4])
