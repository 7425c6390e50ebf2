"""A/B of the reverse sweep: tensor-core (default) vs FFMA (RNDE_BWD_FFMA=1): gradient error vs the FP64-cotangent oracle and time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import regneuralde.jl_b200 as r
from oracle import orc
D, H, B = 784, 100, 512
rng = np.random.default_rng(7)
p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
w = rng.standard_normal((D, B)).astype(np.float32)
node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.Tsit5(), tape_capacity=128)
x = torch.from_numpy(x_np).cuda().requires_grad_(True); p = torch.from_numpy(p_np).cuda().requires_grad_(True)
res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
ws = rng.standard_normal(len(sv)).astype(np.float32)
o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_TANH, reg_kind=orc.REG_ERR_DT, kblock1=98, arith=node.arith)); ref = o.forward(x_np, p_np)
rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
for name, wsv in (("random cotangents on u only", 0 * ws), ("u + saved values", ws)):
    x.grad = None; p.grad = None
    res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
    ((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(wsv).cuda()).sum()).backward()
    dp_hi, dx_hi, _, _ = o.backward(w, wsv, hi=True); dp_32, dx_32, _, _ = o.backward(w, wsv)
    print(f"{name}: dp err {rel(p.grad.cpu().numpy(), dp_hi):.2e} (cpu fp32 {rel(dp_32, dp_hi):.2e})  dx err {rel(x.grad.cpu().numpy(), dx_hi):.2e} (cpu fp32 {rel(dx_32, dx_hi):.2e})")
hd = next(h for k, h in node._handles.items() if k[2])
du = torch.from_numpy(w).cuda().t().contiguous().view(-1); dsv = torch.zeros(129, device="cuda"); dp = torch.empty_like(p); dx = torch.empty(D * B, device="cuda")
import ctypes as C
torch.cuda.synchronize()
for rep in range(3):
    node(x, p, func=r.ERROR_ESTIMATE)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); hd.check(hd.lib.rnde_backward(hd.h, du.data_ptr(), dsv.data_ptr(), dp.data_ptr(), dx.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "bwd"); e1.record()
    torch.cuda.synchronize()
print("RNDE_BWD_FFMA" in os.environ and "FFMA sweep" or "tensor-core sweep", "backward incl. wgrad ms:", e0.elapsed_time(e1), "nfe", nfe)
