"""Developer diagnostic: phase accumulators of the generic forward stepper on the Latent-ODE chain field.
Needs a library built with -DRNDE_TIMELINE (python tools/latent_timeline.py builds one into /tmp and loads it)."""
import ctypes as C, os, subprocess, sys
import numpy as np, torch
sys.path.insert(0, ".")
os.environ["RNDE_DEBUG_TIMELINE"] = "1"
from regneuralde.jl_b200 import _lib as L
tl = "/tmp/libregnde_tl.so"
subprocess.run(["nvcc", *L.NVCC_FLAGS, "-DRNDE_TIMELINE", f"-I{L._INCLUDE}", "-o", tl, str(L.sources()[0])], check=True)
L.LIB_PATH = type(L.LIB_PATH)(tl)
import regneuralde.jl_b200 as r
rng = np.random.default_rng(1234)
W = (50, 20, 50, 20, 50, 20, 50, 20)
layers, K = [], 20
for M in W:
    layers.append(r.Dense(K, M, "tanh")); K = M
saveat = np.unique(np.concatenate([[0.0], np.sort(rng.random(47)), [1.0]]).astype(np.float32)).tolist()
for need_grad in (False, True):
    node = r.TrackedNeuralODE(r.Chain("tanh", *layers), [0.0, 1.0], False, True, r.Tsit5(), saveat=saveat)
    x = torch.from_numpy(rng.standard_normal((20, 512)).astype(np.float32)).cuda().requires_grad_(need_grad)
    for _ in range(3):
        res, nfe, sv = node(x, node.p, func=r.ERROR_ESTIMATE)
    torch.cuda.synchronize()
    hd = next(iter(node._handles.values()))
    buf = (C.c_longlong * 8)()
    hd.lib.rnde_debug_timeline(hd.h, buf, 8)
    names = ["loopheader", "stage combos", "rhs (6 evals)", "norm", "controller", "saves+apply", "-", "loop top"]
    st = node.last_stats
    print(f"tape={need_grad} nfe {nfe} naccept {st.naccept}; cycles per step:")
    for k in range(8):
        print(f"   {names[k]:14s} {buf[k] / max(st.naccept, 1):10.0f}")
