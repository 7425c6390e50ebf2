import sys, torch, numpy as np
sys.path.insert(0, ".")
import regneuralde.jl_b200 as R
g = torch.Generator().manual_seed(1)
ff = R.TrackedFFJORD(R.CSQDynamics(43, 100, generator=g), [0.0, 1.0], True, False, R.Tsit5(), tape_capacity=96)
x = torch.randn(43, 1024, generator=g).cuda(); e = torch.randn(43, 1024, generator=g).cuda()
for _ in range(2):
    o = R.ffjord.loss_and_gradient(ff, x, ff.p, e)
torch.cuda.synchronize(); print(o["nfe"])
