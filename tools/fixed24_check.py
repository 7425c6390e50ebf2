"""FIXED24 (exact integer tensor-core) forward stepper against the oracle's arith = 1: bit-identity, NFE, time."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import regneuralde.jl_b200 as r
from regneuralde.jl_b200 import _lib as L
from oracle import orc
bits = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
for (D, H, B, auto, func) in [(784, 100, 16, False, "ERROR_ESTIMATE"), (784, 100, 50, True, "ERROR_PLUS_STIFFNESS"), (784, 100, 512, False, "ERROR_ESTIMATE"), (640, 72, 40, False, None)]:
    rng = np.random.default_rng(1999)
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
    node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, func is not None, r.AutoTsit5() if auto else r.Tsit5(), arith=L.ARITH_FIXED24, tape_capacity=128)
    fobj = getattr(r, func) if func else None
    x = torch.from_numpy(x_np).cuda(); p = torch.from_numpy(p_np).cuda()
    with torch.no_grad():
        res, nfe, sv = node(x, p, func=fobj)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5): node(x, p, func=fobj)
        torch.cuda.synchronize(); dt = (time.time() - t0) / 5
    cfg = orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_TANH, alg=1 if auto else 0, reg_kind=fobj.kind if fobj else 0, kblock1=D // 8, arith=1)
    ref = orc.Oracle(cfg).forward(x_np, p_np)
    ok_u = np.array_equal(bits(res.cpu().numpy()), bits(ref.u))
    ok_sv = True if sv is None else np.array_equal(bits(sv.saveval.cpu().numpy()), bits(ref.saveval))
    print(f"D={D} H={H} B={B}: nfe {nfe} vs {ref.nf}, u bit-equal {ok_u}, saveval bit-equal {ok_sv}, max|du| {np.abs(res.cpu().numpy()-ref.u).max():.3e}, {dt*1e3:.3f} ms per solve (inference)")
