"""Two training steps of the Neural-SDE classifier (the bench's secondary row) -- the workload of the ncu captures of row N2."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import regneuralde.jl_b200 as R
D, H, B = 32, 64, 512
gen = torch.Generator().manual_seed(1999)
node = R.TrackedNeuralDSDE(R.Chain(R.Dense(D, H, "tanh"), R.Dense(H, D)), R.Dense(D, D), [0.0, 1.0], True, R.SOSRI(), reltol=1.4e-1, abstol=1.4e-1)
clf = R.ClassifierNSDE(R.Dense(784, D, generator=gen), node, R.Dense(D, 10, generator=gen))
xi = torch.rand(784, B, generator=gen).cuda()
yi = torch.nn.functional.one_hot(torch.randint(0, 10, (B,), generator=gen), 10).T.float().cuda()
z = torch.randn(256, D, B, device="cuda")
for _ in range(2):
    o = clf.loss_and_gradient(xi, yi, lam=1.0e2, trajectories=1, func=R.ERROR_ESTIMATE, noise=z)
torch.cuda.synchronize(); print(float(o["loss"]), int(node.last_stats.naccept), int(node.last_stats.nreject))
