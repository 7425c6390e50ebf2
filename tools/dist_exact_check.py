"""2..8 GPUs under torchrun: reference-exact data-parallel mode vs the single batched solve (oracle on rank 0).
   torchrun --nproc-per-node N tools/dist_exact_check.py [B_global]"""
import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
import regneuralde.jl_b200 as R
from regneuralde.jl_b200 import _lib as L

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
D, H = 784, 100
Bg = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Bl = Bg // world
rng = np.random.default_rng(1999)
from oracle import orc
p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, Bg), dtype=np.float32)
w_np = rng.standard_normal((D, Bg)).astype(np.float32)
model = R.MLPDynamics(D, H)
node = R.TrackedNeuralODE(model, [0.0, 1.0], True, True, R.Tsit5(), reltol=1.4e-8, abstol=1.4e-8, dist_mode=L.DIST_EXACT, rank=rank, world=world)
p = torch.from_numpy(p_np).cuda().requires_grad_(True)
x = torch.from_numpy(np.ascontiguousarray(x_np[:, rank * Bl:(rank + 1) * Bl])).cuda()
for it in range(3):
    p.grad = None
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    res, nfe, sv = node(x, p, func=R.ERROR_ESTIMATE)
    torch.cuda.synchronize(); t1 = time.time()
    ws = torch.ones_like(sv.saveval)
    # the saved values are shared by all ranks: every rank applies their FULL cotangent to its own columns, the gradients are summed
    loss = (res * torch.from_numpy(np.ascontiguousarray(w_np[:, rank * Bl:(rank + 1) * Bl])).cuda()).sum() + (sv.saveval * ws).sum()
    loss.backward()
    g = p.grad.clone(); dist.all_reduce(g)
    torch.cuda.synchronize(); t2 = time.time()
    if rank == 0: print(f"iter {it}: fwd {1e3*(t1-t0):.2f} ms  bwd+allreduce {1e3*(t2-t1):.2f} ms nfe {nfe} naccept {node.last_stats.naccept}")
# gather results on rank 0 and compare with the single batched solve
us = [torch.empty_like(res.contiguous()) for _ in range(world)]
dist.all_gather(us, res.detach().contiguous())
if rank == 0:
    u = torch.cat(us, dim=1).cpu().numpy()
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=Bg, kblock1=98, reg_kind=orc.REG_ERR_DT, arith=node.arith)); ref = o.forward(x_np, p_np)
    bits = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    print("exact mode vs single batched solve: nfe", nfe, ref.nf, "naccept", node.last_stats.naccept, ref.naccept,
          "| u bit-equal", np.array_equal(bits(u), bits(ref.u)), "| saveval bit-equal", np.array_equal(bits(sv.saveval.detach().cpu().numpy()), bits(ref.saveval)))
    dp, _, _, _ = o.backward(w_np, np.ones(len(ref.saveval), np.float32), hi=True)
    print("gradient (sum over ranks) vs oracle: relerr %.3e" % (np.abs(g.cpu().numpy() - dp).max() / np.abs(dp).max()))
    dp0, _, _, _ = o.backward(w_np, np.ones(len(ref.saveval), np.float32), hi=True, first_dt_tracked=False)
    print("   (against the frozen-step gradient: %.3e; first-dt term / gradient %.1e)" % (np.abs(g.cpu().numpy() - dp0).max() / np.abs(dp0).max(), np.abs(dp - dp0).max() / np.abs(dp).max()))
# the first-dt term alone (rnde_set_detach diagnostic mode): its scalar dL/d(dt_1) is a sum over ALL ranks' columns
node2 = R.TrackedNeuralODE(model, [0.0, 1.0], True, True, R.Tsit5(), reltol=1.4e-8, abstol=1.4e-8, dist_mode=L.DIST_EXACT, rank=rank, world=world,
                           detach_dt="first_term_only")
p2 = torch.from_numpy(p_np).cuda().requires_grad_(True)
res2, _, sv2 = node2(x, p2, func=R.ERROR_ESTIMATE)
((res2 * torch.from_numpy(np.ascontiguousarray(w_np[:, rank * Bl:(rank + 1) * Bl])).cuda()).sum() + sv2.saveval.sum()).backward()
g2 = p2.grad.clone(); dist.all_reduce(g2)
if rank == 0:
    tp, _, _, _ = o.backward(w_np, np.ones(len(ref.saveval), np.float32), hi=True, first_dt_tracked="term")
    a, b = g2.cpu().numpy().astype(np.float64), tp.astype(np.float64)
    sc = float((a * b).sum() / (b * b).sum())
    print("first-dt term alone (sum over ranks) vs oracle: scale %.5f, direction relerr %.2e" % (sc, np.abs(a / sc - b).max() / np.abs(b).max()))
dist.destroy_process_group()
