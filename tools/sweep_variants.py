"""Developer diagnostic: time of rnde_backward (sweep + weight gradients) at the flagship shape for several builds of the
library (extra nvcc -D flags per variant; "old:<dir>" builds the sources under <dir>).  Usage: python tools/sweep_variants.py"""
import ctypes as C, os, subprocess, sys, importlib
import numpy as np, torch
sys.path.insert(0, ".")
from regneuralde.jl_b200 import _lib as L
VARIANTS = [("current", [])] + [(" ".join(f), f) for f in ([a] for a in sys.argv[1:])]
code = r'''
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
from regneuralde.jl_b200 import _lib as L
L.LIB_PATH = type(L.LIB_PATH)(sys.argv[1])
L.build = lambda *a, **k: L.LIB_PATH
import regneuralde.jl_b200 as r
from oracle import orc
rng = np.random.default_rng(7)
D, H, B = 784, 100, 512
p = torch.from_numpy(orc.glorot_params(rng, D, H)).cuda().requires_grad_(True)
x = torch.from_numpy(rng.random((D, B), dtype=np.float32)).cuda()
for mode in ("all", "all_but_first"):
    node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.Tsit5(), tape_capacity=128, detach_dt=mode)
    ts = []
    for it in range(6):
        res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
        loss = res.sum() + sv.saveval.sum()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss.backward(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"   detach={mode:14s} backward {min(ts[1:]):.3f} ms (nfe {nfe})", flush=True)
'''
open("/tmp/sv_run.py", "w").write(code)
for name, flags in VARIANTS:
    so = f"/tmp/libregnde_{abs(hash(name)) % 10000}.so"
    r_ = subprocess.run(["nvcc", *L.NVCC_FLAGS, *flags, f"-I{L._INCLUDE}", "-o", so, str(L.sources()[0])], capture_output=True, text=True)
    if r_.returncode:
        print(name, "build failed", r_.stderr[-500:]); continue
    print(name, flush=True)
    subprocess.run([sys.executable, "/tmp/sv_run.py", so])
