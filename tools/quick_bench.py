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
