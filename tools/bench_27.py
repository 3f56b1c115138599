"""Development aid: 27-point operator at 512^3, two-pass vs recompute step."""
import builtins, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lanczos_b200 as lz
_p = builtins.print
builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and a[0].startswith("+++")) else _p(*a, **k)
grid = (512, 512, 512)
op = lz.StencilOperator(grid, 0.0, 0.0, weights27=lz.reference_T27_weights(-1.0))
g = torch.Generator(device="cuda").manual_seed(0)
v0 = torch.rand(op.M, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
for kern in ("two_pass", "recompute"):
    L = lz.Lanczos(op)
    for rep in range(2):
        L.execute_Lanczos(12, v0=v0, reorth="none", keep_basis=False, step_kernel=kern, profile=True)
    r = L.result
    print(kern, f"{r.gpu_ms / 12:.4f} ms/step", {k: round(v[0] / max(v[1], 1), 4) for k, v in r.kernel_ms.items() if v[1]}, flush=True)
