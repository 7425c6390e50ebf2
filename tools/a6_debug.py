"""Developer diagnostic: dL/d(dt_1) of the separate-launch path for one case, per sweep kernel (env RNDE_BWD_FFMA / default)."""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["RNDE_A6_EXTERNAL"] = "1"
import regneuralde.jl_b200 as r
from oracle import orc
from test_gpu_parity import A6_CASES, oracle_cfg
name, D, H, B, act_out, auto, func, variant = A6_CASES[int(sys.argv[1])]
rng = np.random.default_rng(11)
p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
fobj = getattr(r, func)
model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
node = r.TrackedNeuralODE(model, [0.0, 1.0], True, True, r.AutoTsit5() if auto else r.Tsit5(), reltol=1.4e-8, abstol=1.4e-8, kernel_variant=variant)
p = torch.from_numpy(p_np).cuda().requires_grad_(True); x = torch.from_numpy(x_np).cuda().requires_grad_(True)
res, nfe, sv = node(x, p, func=fobj)
o = orc.Oracle(oracle_cfg(D, H, B, act_out, 1 if auto else 0, fobj.kind, arith=node.arith)); ref = o.forward(x_np, p_np)
w = rng.standard_normal((D, B)).astype(np.float32); ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
torch.cuda.synchronize()
hd = next(iter(node._handles.values()))
out = (C.c_float * 2)()
hd.lib.rnde_debug_a6(hd.h, out)
_, _, dtb, tb = o.backward(w, ws, hi=True)
N = len(dtb)
ref_d = dtb[0] + tb[1:].sum() - dtb[-1]
print(name, "BWD_FFMA" if os.environ.get("RNDE_BWD_FFMA") else "TC", "naccept", N, "nrej", ref.nreject, " sums:", out[0], out[1], " oracle dL/d(dt_1):", ref_d,
      " parts: dtbar_1", dtb[0], "sum tbar", tb[1:].sum(), "dtbar_N", dtb[-1])
