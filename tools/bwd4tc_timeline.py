"""Developer diagnostic: phase accumulators of bwd4tc_kernel (MNIST shape, batch 512), CTA 0 thread 0, -DRNDE_TIMELINE build in /tmp."""
import ctypes as C, os, subprocess, sys
import numpy as np, torch
sys.path.insert(0, ".")
os.environ["RNDE_DEBUG_TIMELINE"] = "1"
from regneuralde.jl_b200 import _lib as L
tl = "/tmp/libregnde_tl.so"
subprocess.run(["nvcc", *L.NVCC_FLAGS, "-DRNDE_TIMELINE", *sys.argv[1:], f"-I{L._INCLUDE}", "-o", tl, str(L.sources()[0])], check=True, capture_output=True)
L.LIB_PATH = type(L.LIB_PATH)(tl)
import regneuralde.jl_b200 as r
from oracle import orc
rng = np.random.default_rng(7)
D, H, B = 784, 100, 512
p = torch.from_numpy(orc.glorot_params(rng, D, H)).cuda().requires_grad_(True)
x = torch.from_numpy(rng.random((D, B), dtype=np.float32)).cuda()
node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.Tsit5(), tape_capacity=128)
for _ in range(2):
    res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
    (res.sum() + sv.saveval.sum()).backward()
torch.cuda.synchronize()
hd = next(iter(node._handles.values()))
buf = (C.c_longlong * 16)()
names_n = 13
hd.lib.rnde_debug_timeline(hd.h, buf, 16)
nv = 6 * node.last_stats.naccept + 1
names = ["kbar copy", "delta2 + tape + B1 split + sync", "GEMM 1 issue", "GEMM 1 wait", "tcgen05.ld + scatter + wait partials",
         "reduce + delta1 + all-gather wait", "B2 split + sync", "GEMM 2 issue + wait", "tcgen05.ld + transpose + sync", "sync after step entry (skew of the other warps)", "after VJP: zb adjust + distribute", "step entry: StepRec, dsaveval, prefetches", "step entry: cotangent block (tape loads + math)"]
tot = 0
for k, n in enumerate(names):
    print(f"  {n:45s} {buf[k] / nv:8.0f} cycles per VJP"); tot += buf[k] / nv
print("  total", tot, "VJPs", nv)
