"""Small driver for ncu captures of the Gram-Schmidt kernels: n steps of CGS2 at 512^3."""
import builtins, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lanczos_b200 as lz
_p = builtins.print
builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and a[0].startswith("+++")) else _p(*a, **k)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
op = lz.StencilOperator((512, 512, 512), 6.0, -1.0)
g = torch.Generator(device="cuda").manual_seed(0)
v0 = torch.rand(op.M, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
L = lz.Lanczos(op)
L.execute_Lanczos(n, v0=v0, reorth="full", cgs_passes=2, profile=True)
r = L.result
print(f"{r.gpu_ms / n:.3f} ms/step", {k: (round(v[0], 2), v[1]) for k, v in r.kernel_ms.items() if v[1]})
