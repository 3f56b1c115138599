"""Development aid: SELL-32-sigma / CSR SpMV timing on the config-4 graph for a range of sigma.
Kernel variants are selected with LZ_SELL_VARIANT / LZ_SELL_CTAS (read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lanczos_b200 as lz  # noqa: E402
from lanczos_b200 import synth  # noqa: E402

cells = tuple(int(c) for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else (253, 253, 252)))
sigmas = [int(s) for s in (sys.argv[2].split(",") if len(sys.argv) > 2 else "32,64,128,256,512,1024,4096".split(","))]
g = synth.RggGenerator(cells, seed=0)
indptr, indices, data = g.rows(0, g.M)
ctx = lz.Context.default()
x = torch.rand(g.M, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
nnz = indices.numel()
alg = 12.0 * nnz + 16.0 * g.M


def timeit(op, reps=5):
    for _ in range(2):
        op.apply(x, y)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        op.apply(x, y)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


tag = f"var={os.environ.get('LZ_SELL_VARIANT', 'default')} ctas={os.environ.get('LZ_SELL_CTAS', 'default')}"
for fmt, sg in [("sell", s) for s in sigmas] + [("csr", 0)]:
    op = lz.DeviceOperator.from_device_csr(ctx, indptr, indices, data, fmt=fmt, sigma=sg)
    t, st = op.nnz()
    ms = timeit(op)
    print(f"{tag} M={g.M} nnz={nnz} {fmt} sigma={sg}: stored/true={st/max(t,1):.3f} {ms:.3f} ms  {alg/ms/1e6:.0f} GB/s (alg)", flush=True)
    del op
