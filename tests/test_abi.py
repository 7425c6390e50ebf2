"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/regnde.h
declares; the product path fails LOUDLY without a CUDA device (no CPU fallback, nothing routed
through oracle/)."""
import ctypes as C
import re
from pathlib import Path

import pathlib
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    txt = (ROOT / "include" / "regnde.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rnde_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    import regneuralde.jl_b200 as R
    path = R.build()
    assert path.exists()
    lib = C.CDLL(str(path))
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/regnde.h but not exported"
    assert lib.rnde_version() == 100
    # the python binding's list is the same surface
    assert set(R._lib.EXPORTS) <= set(names)


def test_config_struct_matches_header():
    from regneuralde.jl_b200 import _lib as L
    # 17 int32 + 5 float + int64 (8-aligned)
    assert C.sizeof(L.Config) == 184
    assert C.sizeof(L.Stats) == 32
    cfg = L.Config()
    cfg.struct_bytes = C.sizeof(L.Config)
    cfg.state_dim, cfg.hidden_dim, cfg.batch, cfg.time_dep = 784, 100, 512, 1
    lib = L.lib()
    assert lib.rnde_num_params(C.byref(cfg)) == 158568          # 100*785+100+784*101+784
    assert lib.rnde_default_kblock(C.byref(cfg)) == 98
    cfg.state_dim, cfg.hidden_dim = 2, 10
    assert lib.rnde_num_params(C.byref(cfg)) == 64              # test/test_node.jl model
    assert lib.rnde_default_kblock(C.byref(cfg)) == 2


def test_no_cpu_fallback():
    """Without a GPU every entry point must refuse; with one this test is skipped."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import regneuralde.jl_b200 as R
    from regneuralde.jl_b200 import _lib as L
    lib = L.lib()
    assert lib.rnde_device_count() == 0
    cfg = L.Config()
    cfg.struct_bytes = C.sizeof(L.Config)
    cfg.state_dim, cfg.hidden_dim, cfg.batch, cfg.time_dep = 2, 10, 1, 1
    cfg.act_hidden, cfg.act_out = 1, 0
    cfg.t0, cfg.t1, cfg.abstol, cfg.reltol = 0.0, 1.0, 1e-6, 1e-6
    h = C.c_void_p()
    assert lib.rnde_create(C.byref(cfg), C.byref(h)) == L.ERR_CUDA
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R.TrackedNeuralODE(R.TDChain(R.Dense(3, 10, "tanh"), R.Dense(11, 2)), [0.0, 1.0], True, True)
    x = torch.zeros(2, 1)
    with pytest.raises(RuntimeError):
        L.require_device()


def test_test_hooks_validate_arguments_before_touching_the_device():
    from regneuralde.jl_b200 import _lib as L
    lib = L.lib()
    assert lib.rnde_test_csq_rhs(43, 100, 2, 8, None, None, None, C.c_float(0.0), None, None) == L.ERR_ARG      # extra must be 1 or 3
    assert lib.rnde_test_csq_rhs(43, 100, 1, 8, None, None, None, C.c_float(0.0), None, None) == L.ERR_ARG      # null buffers
    assert lib.rnde_test_unary_bits(7, 0, 1, None, None) == L.ERR_ARG


def test_product_never_imports_oracle():
    pkg = ROOT / "regneuralde"
    for f in pkg.rglob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f
    # native side: no translation unit includes a file under oracle/, and the shipped library neither links nor names it
    for f in list((pkg / "jl_b200" / "csrc").glob("*")) + list((ROOT / "include").glob("*.h")):
        for line in f.read_text().splitlines():
            if line.lstrip().startswith("#include"):
                assert "oracle" not in line, (f, line)
    from regneuralde.jl_b200 import _lib as L
    so = pathlib.Path(L.build())
    blob = so.read_bytes()
    assert b"rnde_oracle" not in blob and b"librnde_oracle" not in blob, "product library references the oracle"


def test_host_mirror_validation():
    import regneuralde.jl_b200 as R
    with pytest.raises(NotImplementedError):
        R.TDChain(R.Dense(3, 10, "tanh"), R.Dense(11, 10, "tanh"), R.Dense(11, 2))
    with pytest.raises(ValueError):
        R.TDChain(R.Dense(3, 10, "tanh"), R.Dense(10, 2))
    m = R.MLPDynamics(784, 100)
    assert m.destructure().numel() == 158568 and (m.D, m.H) == (784, 100)
    # Flux.destructure order: vec(W1) column-major first
    d = R.Dense(3, 2)
    flat = d.destructure()
    assert flat[1] == d.W[1, 0] and flat[2] == d.W[0, 1]
