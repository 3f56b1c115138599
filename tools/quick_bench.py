"""Quick device-side timing of the loop in its main modes (development aid, not the bench)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lanczos_b200 as lz  # noqa: E402


def run(grid, n, **kw):
    op = lz.StencilOperator(grid, 2.0 * len(grid), -1.0)
    M = op.M
    g = torch.Generator(device="cuda").manual_seed(0)
    v0 = torch.rand(M, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    L = lz.Lanczos(op)
    best = None
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.execute_Lanczos(n, v0=v0, **kw)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = L.result.gpu_ms
        best = ms if best is None else min(best, ms)
    per = best / n
    print(f"grid={grid} n={n} {kw}: {per:.4f} ms/step  {1e3/per:.1f} steps/s  "
          f"48N-GB/s={48*M/per/1e6:.0f}  launches={L.result.launches} reorth={L.result.reorth_count} wall={wall*1e3:.1f}ms",
          flush=True)
    del L
    torch.cuda.empty_cache()


if __name__ == "__main__":
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        pass
    big = (512, 512, 512)
    sys.stdout = sys.__stdout__
    import builtins
    _print = builtins.print
    def quiet_print(*a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("+++"):
            return
        _print(*a, **k)
    builtins.print = quiet_print
    run(big, 20, reorth="none", keep_basis=False, ref_compat=False, step_kernel="two_pass")
    run(big, 20, reorth="none", keep_basis=False, ref_compat=False, step_kernel="fused")
    run(big, 40, reorth="selective", cgs_passes=2, step_kernel="fused")
    run(big, 40, reorth="selective", cgs_passes=2, step_kernel="two_pass")
    run(big, 20, reorth="selective", cgs_passes=2)
    run(big, 20, reorth="full")
    run(big, 20, reorth="full", cgs_passes=2)
    run((256, 256, 256), 50, reorth="full")
    run((200, 200), 100, reorth="full")
    run((64, 64, 64), 50, reorth="full")
    # 27-point operator and sparse operators at scale
    def run_op(op, n, label, cls=lz.Lanczos, **kw):
        g = torch.Generator(device="cuda").manual_seed(0)
        v0 = torch.rand(op.shape[0], dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        L = cls(op)
        ex = L.execute_Lanczos if cls is lz.Lanczos else L.execute_LanczosOld
        best = None
        for rep in range(3):
            ex(n, v0=v0, profile=True, **kw)
            ms = L.result.gpu_ms
            best = ms if best is None else min(best, ms)
        print(f"{label} n={n} {kw}: {best/n:.4f} ms/step kernels={ {k: (round(v[0]/max(v[1],1),4), v[1]) for k, v in L.result.kernel_ms.items() if v[1]} }", flush=True)
    run_op(lz.StencilOperator(big, 0.0, 0.0, weights27=lz.reference_T27_weights(-1.0)), 12, "27pt 512^3", reorth="none", keep_basis=False)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import lanczos_oracle as orc
    Hbig = orc.laplacian_csr((256, 256, 128), 6.0, -1.0)
    for fmt in ("sell", "csr"):
        run_op(Hbig, 12, f"7pt-as-{fmt} 8.4M rows", reorth="none", keep_basis=False, fmt=fmt)
    Hd = orc.delaunay_graph_laplacian(1_000_000, seed=0)
    for fmt in ("sell", "csr"):
        run_op(Hd, 30, f"delaunay 1M {fmt}", cls=lz.IrrLanczos, reorth="full", fmt=fmt)
