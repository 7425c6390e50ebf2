"""torchrun tool: rnde_allreduce_grads (one-shot push all-reduce over NVLink peer memory) against NCCL on N ranks: values and time."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
import regneuralde.jl_b200 as r
from regneuralde.jl_b200 import _lib as L
node = r.TrackedNeuralODE(r.MLPDynamics(784, 100), [0.0, 1.0], True, True, r.Tsit5(), dist_mode=L.DIST_EXACT, rank=rank, world=world)
node._handle(64, L.REG_ERR_DT, True)
ok = True
for rep in range(4):
    g = torch.from_numpy(np.random.default_rng(10 * rep + rank).standard_normal(158568).astype(np.float32)).cuda()
    ref = g.double(); dist.all_reduce(ref)
    nc = g.clone(); dist.all_reduce(nc)
    node.allreduce_(g); torch.cuda.synchronize()
    err = float((g.double() - ref).abs().max() / ref.abs().max())
    same = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(same, g)
    bitwise = all(torch.equal(same[0], s) for s in same)
    ok = ok and err < 1e-6 and bitwise
    if rank == 0: print(f"rep {rep}: rel err vs float64 sum {err:.2e} (NCCL {float((nc.double()-ref).abs().max()/ref.abs().max()):.2e}), identical on all ranks: {bitwise}")
g = torch.randn(158568, device="cuda"); g3 = torch.randn(7850, device="cuda")
for fn, name in ((lambda: (dist.all_reduce(g), dist.all_reduce(g3)), "NCCL all_reduce x2"), (lambda: node.allreduce_(g, g3), "rnde_allreduce_grads x2")):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    for _ in range(50): fn()
    torch.cuda.synchronize(); dt = (time.time() - t0) / 50
    if rank == 0: print(f"{name}: {dt*1e6:.1f} us per training step's gradients (158568 + 7850 floats), {world} ranks")
if rank == 0: print("OK" if ok else "MISMATCH")
dist.destroy_process_group()
