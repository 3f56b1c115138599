import builtins, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lanczos_b200 as lz
op = lz.StencilOperator((512, 512, 512), 0.0, 0.0, weights27=lz.reference_T27_weights(-1.0))
dev = op.device_handle(lz.Context.default())
x = torch.rand(op.M, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for _ in range(3): dev.apply(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): dev.apply(x, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"K1b 27pt 512^3: {ms:.4f} ms  {16*op.M/ms/1e6:.0f} GB/s")
