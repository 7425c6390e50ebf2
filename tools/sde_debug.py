"""Developer diagnostic: the SDE stepper against the oracle on a case with rejections, attempt by attempt."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import regneuralde.jl_b200 as r
from oracle import sde_oracle as S
from test_gpu_nsde import make, node_for
for (D, H, B, tol, auto, scale) in [(32, 64, 7, 0.02, False, 3.0), (12, 20, 9, 0.05, False, 3.0), (32, 64, 40, 0.06, True, 3.0)]:
    p, x, z = make(1999, D, H, B, scale=scale)
    node = node_for(D, H, True, r.AutoSOSRI2() if auto else r.SOSRI(), tol)
    func = r.STIFFNESS_SCALED if auto else r.ERROR_ESTIMATE
    try:
        with torch.no_grad():
            res, nfe1, nfe2, sv = node(torch.from_numpy(x).cuda(), torch.from_numpy(p).cuda(), func=func, noise=torch.from_numpy(z).cuda())
    except Exception as ex:
        print("CUDA solve failed:", ex)
    f, g = S.drift_diffusion(p, np.float32, D=D, H=H)
    ref = S.solve(x, f, g, z, alg=S.ALG_AUTO_SOSRI2 if auto else S.ALG_SOSRI, reg_kind=S.REG_STIFF_SCALED if auto else S.REG_ERR_DT, abstol=tol, reltol=tol)
    st = node.last_stats
    print(f"D={D} B={B} tol={tol}: cuda nacc {st.naccept} nrej {st.nreject} draws {st.draws} retcode {st.retcode} | oracle nacc {ref.naccept} nrej {ref.nreject} draws {ref.draws}")
    got = node.attempts()
    for i, (a, b) in enumerate(zip(got, zip(ref.dts, ref.eests, ref.accepted))):
        bad = abs(a[0] - b[0]) > 1e-5 * b[0] or abs(a[1] - b[1]) > 1e-3 * b[1] or a[2] != b[2]
        if bad or i < 3 or not b[2]:
            print(f"   attempt {i}: cuda dt {a[0]:.7g} EEst {a[1]:.6g} acc {a[2]} | oracle dt {b[0]:.7g} EEst {b[1]:.6g} acc {b[2]} {'<-- differs' if bad else ''}")
        if bad:
            break
    if st.retcode == 0:
        print("   max |u - u_ref| / max|u_ref| =", np.abs(res.cpu().numpy() - ref.u).max() / np.abs(ref.u).max())
