import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import orc
import regneuralde.jl_b200 as r
D,H,B = 20,50,100
for (auto, func, okind) in ((True, r.ERROR_PLUS_STIFFNESS, 4), (False, r.ERROR_ESTIMATE, 1), (True, r.STIFFNESS_ESTIMATE, 2)):
    rng = np.random.default_rng(7)
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D,B),dtype=np.float32)
    model = r.TDChain(r.Dense(D+1,H,"tanh"), r.Dense(H+1,D,"tanh"))
    node = r.TrackedNeuralODE(model,[0.0,1.0],True,True,r.AutoTsit5() if auto else r.Tsit5(),reltol=1.4e-8,abstol=1.4e-8)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True); x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res,nfe,sv = node(x,p,func=func)
    o = orc.Oracle(orc.OracleConfig(D=D,H=H,B=B,alg=1 if auto else 0,reg_kind=okind)); ref = o.forward(x_np,p_np)
    w = rng.standard_normal((D,B)).astype(np.float32); ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
    for (wu, wsv, name) in ((w, ws, "both"), (w, 0*ws, "u only"), (0*w, ws, "sv only")):
        p.grad = None; x.grad = None
        res,nfe,sv = node(x,p,func=func)
        loss = (res*torch.from_numpy(wu).cuda()).sum() + (sv.saveval*torch.from_numpy(wsv).cuda()).sum()
        loss.backward()
        hi = o.backward(wu,wsv,hi=True)[0]; c32 = o.backward(wu,wsv)[0]; g = p.grad.cpu().numpy()
        offs = [0, H*(D+1), H*(D+1)+H, H*(D+1)+H+D*(H+1), len(g)]
        names = ["W1","b1","W2","b2"]
        parts = []
        for k in range(4):
            a,b_ = offs[k], offs[k+1]
            parts.append("%s cuda %.1e cpu %.1e (max %.1e)" % (names[k], np.abs(g[a:b_]-hi[a:b_]).max(), np.abs(c32[a:b_]-hi[a:b_]).max(), np.abs(hi[a:b_]).max()))
        print(okind, name, " | ".join(parts))
