import ctypes as C, subprocess, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path('/root/repo'); sys.path.insert(0, str(ROOT))
SRC = r'''
#include "regnde_canon.h"
void t_tanh_bits(unsigned first, long n, float* y) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) y[i] = canon_tanhf(rnde_u2f(first + (unsigned)i));
}
'''
open('/tmp/c.c','w').write(SRC)
subprocess.run(["/usr/bin/gcc","-O2","-shared","-fPIC","-fopenmp","-ffp-contract=off","-mfma",f"-I{ROOT/'include'}","/tmp/c.c","-o","/tmp/c.so","-lm"],check=True)
cpu = C.CDLL('/tmp/c.so')
import regneuralde.jl_b200 as R
lib = R.lib()
last = int(np.float32(9.25).view(np.uint32)); chunk = 1<<26
y_dev = torch.empty(chunk, device='cuda'); y_cpu = np.empty(chunk, np.float32)
tot = 0; ex = []
hist = {}
for first in range(0, last, chunk):
    n = min(chunk, last-first)
    lib.rnde_test_tanh_bits(C.c_uint32(first), n, y_dev.data_ptr(), None)
    g = y_dev[:n].cpu().numpy()
    cpu.t_tanh_bits(C.c_uint(first), C.c_long(n), y_cpu.ctypes.data_as(C.c_void_p))
    d = np.nonzero(g.view(np.uint32) != y_cpu[:n].view(np.uint32))[0]
    tot += len(d)
    for i in d[:3]:
        x = np.uint32(first+i).view(np.float32)
        ex.append((float(x), float(g[i]), float(y_cpu[i]), int(g[i:i+1].view(np.int32)[0]) - int(y_cpu[i:i+1].view(np.int32)[0])))
    if len(d): hist[first] = len(d)
print("total mismatches", tot, "of", last)
print("by chunk (first bits -> count):", {hex(k): v for k, v in hist.items()})
for e in ex[:30]: print("x=%.9g gpu=%.9g cpu=%.9g ulpdiff=%d" % e)
