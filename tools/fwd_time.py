"""ms per forward solve (MNIST shape, batch 512, inference and taped) for quick A/B of stepper variants."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pathlib import Path
from regneuralde.jl_b200 import _lib as _L
if len(sys.argv) > 1:
    _L.LIB_PATH = Path(sys.argv[1]).resolve()      # A/B: a variant build of the library
import regneuralde.jl_b200 as r
from oracle import orc
rng = np.random.default_rng(1999)
D, H, B = 784, 100, 512
p = torch.from_numpy(orc.glorot_params(rng, D, H)).cuda(); x = torch.from_numpy(rng.random((D, B), dtype=np.float32)).cuda()
node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.Tsit5(), tape_capacity=128)
for grad in (False, True):
    xx = x.clone().requires_grad_(grad)
    for _ in range(3): res, nfe, sv = node(xx, p, func=r.ERROR_ESTIMATE)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hd = [h for k, h in node._handles.items() if k[2] == grad][0]
    import ctypes as C
    xb = r.colmajor(x); ub = torch.empty(D * B, device="cuda"); svb = torch.zeros(129, device="cuda")
    ts = []
    for _ in range(7):
        torch.cuda.synchronize(); e0.record()
        hd.lib.rnde_forward(hd.h, xb.data_ptr(), p.data_ptr(), ub.data_ptr(), svb.data_ptr(), None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"tape={grad}: {sorted(ts)[3]:.4f} ms per solve, nfe {nfe}")
