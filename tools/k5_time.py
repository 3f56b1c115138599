"""Development aid: time the full Ritz lift (lz_ritz_vectors, K5) at 512^3, n = k = 60 - GEMM form vs the
4-columns-per-sweep form (LZ_K5_GEMM=0).  python tools/k5_time.py [side] [n]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lanczos_b200 import _capi, engine  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
k = n
M = side ** 3
ctx = engine.Context.default()
ld = engine.padded_ld(M)
V = torch.rand((n, ld), dtype=torch.float64, device="cuda")
Y = torch.empty((k, ld), dtype=torch.float64, device="cuda")
S = np.asfortranarray(np.linalg.qr(np.random.RandomState(0).randn(n, n))[0])
scale = np.ones(n)
best = 1e9
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _capi.check(ctx.lib.lz_ritz_vectors(ctx.handle, C.c_void_p(V.data_ptr()), ld, n, M, scale.ctypes.data_as(C.c_void_p),
                                        S.ctypes.data_as(C.c_void_p), k, C.c_void_p(Y.data_ptr()), ld))
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
byts = 2.0 * n * 8 * M
flops = 2.0 * n * k * M
print(f"K5 lift {side}^3 n=k={n} gemm={os.environ.get('LZ_K5_GEMM', '1')}: {best*1e3:.2f} ms  "
      f"{byts/best/1e9:.0f} GB/s of 2n8M  {flops/best/1e12:.2f} TFLOP/s fp64", flush=True)
