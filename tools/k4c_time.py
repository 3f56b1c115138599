import sys, builtins, torch
sys.path.insert(0, ".")
import lanczos_b200 as lz
_p = builtins.print
builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and a[0].startswith("+++")) else _p(*a, **k)
op = lz.StencilOperator((512, 512, 512), 6.0, -1.0)
v0 = torch.rand(op.M, dtype=torch.float64, device="cuda") * 2 - 1
L = lz.Lanczos(op)
for rep in range(2):
    L.execute_Lanczos(60, v0=v0, reorth="full", cgs_passes=2, profile=True)
r = L.result
import os
print("cp.async" if os.environ.get("LZ_K4C_TMA") == "0" else "TMA", f"{r.gpu_ms/60:.3f} ms/step", {k: (round(v[0], 2), v[1]) for k, v in r.kernel_ms.items() if v[1]}, flush=True)
