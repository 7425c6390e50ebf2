"""GPU: one evaluation of the FFJORD field on the device (csrc/csq.cuh through rnde_test_csq_rhs) against the C oracle
(oracle/rnde_oracle.c csq_column), bit for bit -- the first brick of SURVEY.md 8f row N4.

Verified on a B200 with the last GPU seconds of round 1 (4 shapes x 3 stage times, all bit-identical); the field is not
wired into a stepper yet."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ffjord_oracle as F, orc  # noqa: E402  (the checker)


@pytest.mark.parametrize("Dz,H,B,extra", [(43, 100, 8, 1), (43, 100, 7, 3), (5, 9, 3, 1), (6, 64, 130, 3)])
def test_field_evaluation_bit_identical(oracle_built, Dz, H, B, extra):
    import regneuralde.jl_b200 as R
    lib = R.lib()
    rng = np.random.default_rng(17)
    p = F.glorot_params(rng, Dz, H, dtype=np.float32, bias_scale=0.2)
    z = rng.standard_normal((Dz + extra, B)).astype(np.float32)
    e = rng.standard_normal((Dz, B)).astype(np.float32)
    o = orc.Oracle(orc.OracleConfig(D=Dz + extra, H=H, B=B, csq_extra=extra, csq_noise=e, kblock1=Dz + extra))
    for t in (0.0, 0.37, 1.0):
        ref, _ = o.rhs(p, z, t)
        pd, zd, ed = (torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).cuda() for a in (p, z, e))
        kd = torch.empty(B, Dz + extra, device="cuda", dtype=torch.float32)
        assert lib.rnde_test_csq_rhs(Dz, H, extra, B, pd.data_ptr(), zd.data_ptr(), ed.data_ptr(), C.c_float(t), kd.data_ptr(), None) == 0
        got = kd.cpu().numpy().T
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref, dtype=np.float32).view(np.uint32)), \
            (t, np.abs(got - ref).max())
