"""Developer diagnostic: clock64 phase timeline of fwd4s_kernel<100,784> (default; "fmachain": fwd4_kernel<100,98>) at the MNIST shape, batch 512, CTA 0.
Builds a -DRNDE_TIMELINE copy of the library into /tmp; the product build has the markers compiled out."""
import ctypes as C, os, subprocess, sys
import numpy as np, torch
sys.path.insert(0, ".")
os.environ["RNDE_DEBUG_TIMELINE"] = "1"
from regneuralde.jl_b200 import _lib as L
tl = "/tmp/libregnde_tl.so"
subprocess.run(["nvcc", *L.NVCC_FLAGS, "-DRNDE_TIMELINE", *[a for a in sys.argv[1:] if a.startswith("-D")], f"-I{L._INCLUDE}", "-o", tl, str(L.sources()[0])], check=True, capture_output=True)
L.LIB_PATH = type(L.LIB_PATH)(tl)
import regneuralde.jl_b200 as r
from oracle import orc
rng = np.random.default_rng(1999)
D, H, B = 784, 100, 512
p = torch.from_numpy(orc.glorot_params(rng, D, H)).cuda()
x = torch.from_numpy(rng.random((D, B), dtype=np.float32)).cuda()
ARITH = 1 if "fixed24" in sys.argv else (0 if "fmachain" in sys.argv else 2)      # default: the split-K stepper (fwd4s_kernel)
node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.Tsit5(), tape_capacity=128, arith=ARITH)
need = len(sys.argv) > 1 and sys.argv[1] == "tape"
x.requires_grad_(need)
for _ in range(2):
    res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
torch.cuda.synchronize()
hd = next(iter(node._handles.values()))
buf = (C.c_longlong * 8000)()
hd.lib.rnde_debug_timeline(hd.h, buf, 8000)
a = np.array(list(buf)).reshape(-1, 2)
a = a[: np.nonzero(a[:, 1])[0].max() + 1]
ids, ts = a[:, 0], a[:, 1]
if ARITH == 1:
    xn = {1: "column max of z + sync", 2: "digits of z -> B1 + fence + sync", 3: "GEMM 1 issue + wait", 4: "TMEM -> T -> partial -> scatter", 5: "sync", 6: "wait partials", 7: "reduce + tanh + all-gather", 8: "sync", 9: "wait hidden", 20: "column max + digits of h -> B2 + sync", 21: "GEMM 2 issue + wait", 22: "TMEM -> T -> act -> transpose + sync", 10: "tile read-back + tape"}
names = {1: "stageZ+sync", 2: "layer-1 partial", 3: "sync", 4: "pair+scatter", 5: "sync", 6: "wait partials", 7: "reduce+tanh+gather", 8: "sync",
         9: "wait hidden", 10: "layer-2+tanh", 0: "between evals / combos", 11: "norm: col sumsq + sync", 12: "norm: cluster reduce + sync",
         13: "norm: publish", 14: "norm: grid barrier", 15: "norm: total", 16: "controller"}
dur = {}
for i in range(40, len(ids) - 1):
    dur.setdefault(int(ids[i + 1]), []).append(ts[i + 1] - ts[i])        # time spent reaching marker ids[i+1]
print(f"nfe {nfe} naccept {node.last_stats.naccept} tape={need}")
tot = 0
if ARITH == 1:
    names.update(xn)
if ARITH == 2:
    names.update({4: "layer 1 (8x8 FFMA2) + reduce-scatter + send", 10: "layer 2 (8x8 FFMA2) + reduce-scatter + tanh"})
for k in [1, 2, 3, 4, 5, 6, 7, 8, 9, 20, 21, 22, 10, 0, 11, 12, 13, 14, 15, 16]:
    if k in dur:
        v = np.array(dur[k]); print(f"  -> {names[k]:28s} n={len(v):4d} mean {v.mean():8.0f} median {np.median(v):8.0f}")
steps = np.nonzero(ids == 16)[0]
print("cycles per step (controller to controller):", np.diff(ts[steps]).mean())
