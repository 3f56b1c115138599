"""Small driver for ncu captures: a few steps of one mode at 512^3."""
import builtins, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lanczos_b200 as lz
_p = builtins.print
builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and a[0].startswith("+++")) else _p(*a, **k)
mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
grid = (512, 512, 512)
op = lz.StencilOperator(grid, 6.0, -1.0)
g = torch.Generator(device="cuda").manual_seed(0)
v0 = torch.rand(op.M, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
L = lz.Lanczos(op)
L.execute_Lanczos(n, v0=v0, reorth="none", keep_basis=False, step_kernel=mode)
print(mode, L.result.gpu_ms / n, "ms/step")
