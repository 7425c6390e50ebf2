"""Developer diagnostic of the open toy-shape issue: the stiffness-estimate regulariser's parameter gradient
from the generic sweep, with the default library and with a build whose weight-gradient contraction accumulates every product in
Float64 (-DRNDE_WG_DOUBLE).  Usage: python tools/stiff_probe.py [nvcc flags ...]"""
import subprocess, sys
sys.path.insert(0, ".")
from regneuralde.jl_b200 import _lib as L
if len(sys.argv) > 1:
    so = "/tmp/libregnde_probe.so"
    subprocess.run(["nvcc", *L.NVCC_FLAGS, *sys.argv[1:], f"-I{L._INCLUDE}", "-o", so, str(L.sources()[0])], check=True, capture_output=True)
    L.LIB_PATH = type(L.LIB_PATH)(so); L.build = lambda *a, **k: L.LIB_PATH
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import regneuralde.jl_b200 as r
from oracle import orc
from gradbar import cpu32_noise
from test_gpu_parity import make_node, oracle_cfg
for name, D, H, B, act_out, auto, func, variant in [("test_node stiffreg grad", 2, 10, 5, 0, True, "STIFFNESS_ESTIMATE", 0), ("mid combined grad", 20, 50, 100, 1, True, "ERROR_PLUS_STIFFNESS", 0)]:
    rng = np.random.default_rng(7)
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
    fobj = getattr(r, func)
    node = make_node(D, H, act_out, True, r.AutoTsit5() if auto else r.Tsit5(), variant)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True); x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=fobj)
    o = orc.Oracle(oracle_cfg(D, H, B, act_out, 1, fobj.kind, arith=node.arith)); ref = o.forward(x_np, p_np)
    w = rng.standard_normal((D, B)).astype(np.float32); ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
    ((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
    dp_hi, dx_hi, _, _ = o.backward(w, ws, hi=True)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    c_p, c_x = cpu32_noise(o, w, ws, n=5)
    print(name, "e_p %.2e c_p %.2e ratio %.2f | e_x %.2e c_x %.2e" % (rel(p.grad.cpu().numpy(), dp_hi), c_p, rel(p.grad.cpu().numpy(), dp_hi) / c_p, rel(x.grad.cpu().numpy(), dx_hi), c_x))
