"""Row A6 diagnostics: the first-dt gradient term (detach_dt all_but_first minus all) of the CUDA path, of the CPU Float32
adjoint and of the Float64-cotangent yardstick, per case.  Usage: python tools/a6_check.py"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import regneuralde.jl_b200 as r
from oracle import orc
from test_gpu_parity import A6_CASES, oracle_cfg

for name, D, H, B, act_out, auto, func, variant in A6_CASES:
    rng = np.random.default_rng(11)
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
    fobj = getattr(r, func)
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, True, r.AutoTsit5() if auto else r.Tsit5(), reltol=1.4e-8, abstol=1.4e-8,
                              kernel_variant=variant, detach_dt="first_term_only")
    p = torch.from_numpy(p_np).cuda().requires_grad_(True); x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=fobj)
    o = orc.Oracle(oracle_cfg(D, H, B, act_out, 1 if auto else 0, fobj.kind, arith=node.arith)); ref = o.forward(x_np, p_np)
    w = rng.standard_normal((D, B)).astype(np.float32); ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
    ((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
    torch.cuda.synchronize()
    mp, mx = p.grad.cpu().numpy().astype(np.float64), x.grad.cpu().numpy().astype(np.float64)
    tp, tx, _, _ = o.backward(w, ws, hi=True, first_dt_tracked="term")
    cp, cx, _, _ = o.backward(w, ws, first_dt_tracked="term")
    full, _, _, _ = o.backward(w, ws, hi=True)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    k = float((mp * tp).sum() / (tp.astype(np.float64) ** 2).sum()); kc = float((cp.astype(np.float64) * tp).sum() / (tp.astype(np.float64) ** 2).sum())
    blocks = {"W1": (0, H * D), "W1t": (H * D, H * (D + 1)), "b1": (H * (D + 1), H * (D + 2)), "W2": (H * (D + 2), H * (D + 2) + D * H),
              "W2t": (H * (D + 2) + D * H, H * (D + 2) + D * (H + 1)), "b2": (H * (D + 2) + D * (H + 1), H * (D + 2) + D * (H + 2))}
    per = "  ".join(f"{kk} {np.abs(mp[a:b] - tp[a:b]).max() / np.abs(tp[a:b]).max():.1e}" for kk, (a, b) in blocks.items())
    print(f"{name:58s} term/grad {np.abs(tp).max()/np.abs(full).max():.1e}  cuda: dp {rel(mp,tp):.2e} dx {rel(mx,tx):.2e} scale {k:.5f}   cpu32: dp {rel(cp,tp):.2e} dx {rel(cx,tx):.2e} scale {kc:.5f}", flush=True)
    print("      blocks: " + per, flush=True)
