"""GPU: the canonical device math (include/regnde_canon.h compiled by nvcc) must equal the CPU build
of the same header BIT FOR BIT -- this is what makes step counts, NFE and saved values reproducible.
The tanh check is exhaustive over every Float32 in [0, 9.25) (the device uses a branch-free ranged
division there); pow/log10 are sampled densely."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

SRC = r'''
#include "regnde_canon.h"
void t_tanh_bits(unsigned first, long n, float* y) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) y[i] = canon_tanhf(rnde_u2f(first + (unsigned)i));
}
void t_unary_bits(int fn, unsigned first, long n, float* y) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const float x = rnde_u2f(first + (unsigned)i);
        y[i] = fn == 0 ? canon_tanhf(x) : fn == 1 ? canon_sigmoidf(x) : fn == 2 ? canon_softplusf(x) : canon_expnegf(x);
    }
}
void t_powf(const float* x, float e, float* y, float* l, long n) { for (long i = 0; i < n; ++i) { y[i] = canon_powf(x[i], e); l[i] = canon_log10f(x[i]); } }
'''


@pytest.fixture(scope="module")
def cpu(tmp_path_factory):
    d = tmp_path_factory.mktemp("canon_gpu")
    (d / "c.c").write_text(SRC)
    so = d / "c.so"
    subprocess.run(["/usr/bin/gcc", "-O2", "-shared", "-fPIC", "-fopenmp", "-ffp-contract=off", "-mfma", f"-I{ROOT / 'include'}", str(d / "c.c"),
                    "-o", str(so), "-lm"], check=True)
    return C.CDLL(str(so))


def test_tanh_bit_identical_exhaustive(cpu):
    import regneuralde.jl_b200 as R
    lib = R.lib()
    last = int(np.float32(9.25).view(np.uint32))
    chunk = 1 << 26
    y_dev = torch.empty(chunk, device="cuda", dtype=torch.float32)
    y_cpu = np.empty(chunk, dtype=np.float32)
    bad = 0
    for first in range(0, last, chunk):
        n = min(chunk, last - first)
        assert lib.rnde_test_tanh_bits(C.c_uint32(first), n, y_dev.data_ptr(), None) == 0
        g = y_dev[:n].cpu().numpy()
        cpu.t_tanh_bits(C.c_uint(first), C.c_long(n), y_cpu.ctypes.data_as(C.c_void_p))
        bad += int(np.count_nonzero(g.view(np.uint32) != y_cpu[:n].view(np.uint32)))
    assert bad == 0, f"{bad} of {last} inputs differ between GPU and CPU canon_tanhf"
    # negative half follows from the sign transfer; spot-check it and the specials
    x = torch.tensor([-0.0, -1e-40, -0.3, -9.0, -50.0, float("inf"), float("-inf")], device="cuda")
    y = torch.empty_like(x)
    assert lib.rnde_test_tanh(x.data_ptr(), y.data_ptr(), x.numel(), None) == 0
    assert abs(float(y[-2]) - 1.0) < 1e-7 and abs(float(y[-1]) + 1.0) < 1e-7      # clamp at 9.01: saturates within 1 ulp of +-1
    assert torch.signbit(y[0]) and float(y[3]) <= -0.99999


def test_pow_log10_bit_identical(cpu):
    import regneuralde.jl_b200 as R
    lib = R.lib()
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(-40, 10, 1 << 22)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd); ld = torch.empty_like(xd)
    y = np.empty_like(x); l = np.empty_like(x)
    for e in (np.float32(7.0 / 50.0), np.float32(2.0 / 25.0)):
        assert lib.rnde_test_pow(xd.data_ptr(), C.c_float(float(e)), yd.data_ptr(), ld.data_ptr(), x.size, None) == 0
        cpu.t_powf(x.ctypes.data_as(C.c_void_p), C.c_float(float(e)), y.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p), C.c_long(x.size))
        assert np.array_equal(yd.cpu().numpy().view(np.uint32), y.view(np.uint32))
        assert np.array_equal(ld.cpu().numpy().view(np.uint32), l.view(np.uint32))


def test_sigmoid_softplus_bit_identical(cpu):
    """The FFJORD activations (canon_sigmoidf / canon_softplusf): every Float32 with 2^-26 <= |x| < 104,
    both signs -- below 2^-26 the result no longer depends on x beyond the last bit and above 104 it is saturated, so those
    ranges and the specials are sampled."""
    import regneuralde.jl_b200 as R
    lib = R.lib()
    lo, hi = int(np.float32(2.0 ** -26).view(np.uint32)), int(np.float32(104.0).view(np.uint32))
    chunk = 1 << 26
    y_dev = torch.empty(chunk, device="cuda", dtype=torch.float32)
    y_cpu = np.empty(chunk, dtype=np.float32)
    for fn in (1, 2):
        for sign in (0, 0x80000000):
            for first in range(lo, hi, chunk):
                n = min(chunk, hi - first)
                assert lib.rnde_test_unary_bits(fn, C.c_uint32(first | sign), n, y_dev.data_ptr(), None) == 0
                g = y_dev[:n].cpu().numpy()
                cpu.t_unary_bits(fn, C.c_uint(first | sign), C.c_long(n), y_cpu.ctypes.data_as(C.c_void_p))
                bad = int(np.count_nonzero(g.view(np.uint32) != y_cpu[:n].view(np.uint32)))
                assert bad == 0, f"fn {fn}: {bad} inputs differ in the chunk starting at bit pattern {first | sign:#x}"
    # windows over the tiny, saturated and infinite ends (NaNs are left out: their payloads are not part of the contract);
    # canon_expnegf takes a >= 0 only
    inf = 0x7F800000
    for fn in (1, 2, 3):
        for first, n in ((0, 1 << 20), (lo - (1 << 20), 1 << 20), (hi - 512, 1 << 20), (hi + (1 << 22), 1 << 20), (inf - (1 << 20), (1 << 20) + 1)):
            for sign in ((0,) if fn == 3 else (0, 0x80000000)):
                assert lib.rnde_test_unary_bits(fn, C.c_uint32(first | sign), n, y_dev.data_ptr(), None) == 0
                g = y_dev[:n].cpu().numpy()
                cpu.t_unary_bits(fn, C.c_uint(first | sign), C.c_long(n), y_cpu.ctypes.data_as(C.c_void_p))
                assert np.array_equal(g.view(np.uint32), y_cpu[:n].view(np.uint32)), (fn, hex(first | sign))
    # exp(-a) exhaustively over [2^-26, 104) as well (it feeds both activations; its subnormal tail rounds once)
    for first in range(lo, hi, chunk):
        n = min(chunk, hi - first)
        assert lib.rnde_test_unary_bits(3, C.c_uint32(first), n, y_dev.data_ptr(), None) == 0
        g = y_dev[:n].cpu().numpy()
        cpu.t_unary_bits(3, C.c_uint(first), C.c_long(n), y_cpu.ctypes.data_as(C.c_void_p))
        assert np.array_equal(g.view(np.uint32), y_cpu[:n].view(np.uint32)), hex(first)
