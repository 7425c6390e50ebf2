"""CPU: what the shipped library was compiled INTO (cuobjdump -sass of the in-tree libregnde.so).  The claims of DESIGN.md
section 4 about which kernels run on the 5th-generation tensor cores (tcgen05: UTCHMMA / UTCIMMA with LDTM read-back), which
use the bulk-copy (TMA) engine for their tape tiles (UBLKCP) and mbarriers (SYNCS), and that the canonical arithmetic is not
contracted behind our back, are checked against the machine code -- mnemonics per B200_PROFILING.md."""
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    import os
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    from regneuralde.jl_b200 import _lib as L
    out = subprocess.run([CUOBJDUMP, "-sass", str(L.build())], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out or "SM100" in out.upper() or "EF_CUDA_SM100" in out, "library is not built for sm_100a"
    funcs = {}
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    return {k: "\n".join(v) for k, v in funcs.items()}


def body(sass, needle):
    hits = [k for k in sass if needle in k]
    assert len(hits) == 1, (needle, hits)
    return sass[hits[0]]


def count(text, mnemonic):
    return len(re.findall(r"\b%s\b" % re.escape(mnemonic), text))


def test_reverse_sweep_and_weight_gradients_run_on_tcgen05(sass):
    for inst in ("bwd4tc_kernelILb1E", "bwd4tc_kernelILb0E"):                  # with / without the first-dt additions (a6.cuh)
        b = body(sass, inst)
        assert count(b, "UTCHMMA") >= 8 and "LDTM" in b and "UTCBAR" in b      # tcgen05.mma kind::f16, tcgen05.ld, tcgen05.commit
        assert "UBLKCP" in b                                                   # delta tiles leave as bulk stores
    w = body(sass, "wgrad_tc_kernelE")
    assert count(w, "UTCHMMA") >= 3 and "LDTM" in w                            # 3xTF32 split products
    for k, v in sass.items():                                                  # no legacy warp-level MMA anywhere
        assert not re.search(r"\b(HMMA|IMMA|QGMMA|HGMMA)\b", v), k


def test_exact_integer_forward_uses_int8_tensor_cores(sass):
    x = body(sass, "fwd4x_kernel")
    assert count(x, "UTCIMMA") >= 8 and "LDTM" in x and "UTCHMMA" not in x     # tcgen05.mma kind::i8 only


def test_forward_stepper_is_fp32_fma_with_bulk_tape_stores_and_mbarriers(sass):
    f = body(sass, "fwd4_kernelILi100ELi98E")
    assert len(re.findall(r"\bUBLKCP", f)) >= 3                                # sZ, sKt, sH tiles
    assert "SYNCS" in f and "UTCHMMA" not in f and "UTCIMMA" not in f
    ffma = len(re.findall(r"\bFFMA\b", f))
    assert ffma >= 1500, ffma                                                  # unrolled 4x4 register tiles of both layers
    # canonical arithmetic: the state / stage arithmetic is explicit fma and separate add / mul -- with -fmad=false ptxas
    # never sees a contractable mul+add, so FMUL and FADD both remain in the code
    assert re.search(r"\bFADD\b", f) and re.search(r"\bFMUL\b", f)
    assert "MUFU.TANH" not in f and "MUFU.EX2" not in f                        # canon_tanhf, not the SFU approximations


def test_latent_kernels_keep_their_weights_on_chip(sass):
    for name in ("gru_fwd_kernel", "gru_bwd_kernel"):
        g = body(sass, name)
        lds, ldg = len(re.findall(r"\bLDS(\.\w+)*\b", g)), len(re.findall(r"\bLDG(\.\w+)*\b", g))
        assert lds > 2 * ldg, (name, lds, ldg)                                 # weights come from shared memory; global loads are the staging and the tape
