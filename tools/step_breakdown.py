"""Where a training step's time goes, with CUDA events between the library calls (MNIST shape, batch 512)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import regneuralde.jl_b200 as R
from regneuralde.jl_b200 import _lib as L
import bench
rng = np.random.default_rng(1999)
p2_np, p3_np = bench.init_params(rng)
xs, ys = bench.synth_batches(np.random.default_rng(2000), 4, 512)
node = R.TrackedNeuralODE(R.MLPDynamics(784, 100), [0.0, 1.0], True, True, R.Tsit5(), tape_capacity=128)
clf = R.ClassifierNODE(None, node, R.Dense(784, 10))
clf.p2.copy_(torch.from_numpy(p2_np)); clf.p3.copy_(torch.from_numpy(p3_np)); node.p = clf.p2
opt = R.Optimiser(1.0e-5, 0.1, 0.9)
xd = [torch.from_numpy(x).cuda() for x in xs]; yd = [torch.from_numpy(y).cuda() for y in ys]
hd = node._handle(512, L.REG_ERR_DT, True); lib = hd.lib
ws = clf._workspace(512, xd[0].device, hd.cfg.tape_capacity)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def ev(): return torch.cuda.Event(enable_timing=True)
acc = np.zeros(5); N = 30
for it in range(N + 5):
    x, y = xd[it % 4], yd[it % 4]
    xbuf = R.colmajor(x); ybuf = R.colmajor(y)
    e = [ev() for _ in range(6)]
    e[0].record()
    lib.rnde_forward(hd.h, xbuf.data_ptr(), clf.p2.data_ptr(), ws["u"].data_ptr(), ws["sv"].data_ptr(), None, stream)
    e[1].record()
    lib.rnde_head_loss_grad(hd.h, ws["u"].data_ptr(), clf.p3.data_ptr(), ybuf.data_ptr(), 10, C.c_float(1.0), ws["loss"].data_ptr(), ws["logits"].data_ptr(), ws["du"].data_ptr(), ws["g3"].data_ptr(), stream)
    lib.rnde_reg_agg(hd.h, 0, C.c_float(100.0), C.c_float(1.0), ws["sv"].data_ptr(), ws["dsv"].data_ptr(), ws["reg"].data_ptr(), stream)
    e[2].record()
    lib.rnde_backward(hd.h, ws["du"].data_ptr(), ws["dsv"].data_ptr(), ws["g2"].data_ptr(), None, stream)
    e[3].record()
    R.update_parameters_((clf.p1, clf.p2, clf.p3), (clf.p1, ws["g2"], ws["g3"]), opt)
    e[4].record()
    torch.cuda.synchronize()
    if it >= 5:
        acc += [e[i].elapsed_time(e[i + 1]) for i in range(4)] + [e[0].elapsed_time(e[4])]
names = ["forward solve", "head + regulariser aggregation", "backward (sweep + weight gradients)", "optimiser", "whole step (events)"]
for n, v in zip(names, acc / N): print(f"{n:40s} {v:7.3f} ms")
