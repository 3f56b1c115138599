"""Time the windowed SELL apply on the config-4 graph with parts of the kernel switched off (LZ_SELLW_DEBUG)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lanczos_b200 import engine, synth
from lanczos_b200.engine import DeviceCSR

ctx = engine.Context.default()
gen = synth.RggGenerator((253, 253, 252), seed=0)
H = DeviceCSR(*gen.rows(0, gen.M))
op = engine.as_device_operator(H, ctx, fmt="sell")
print("windowed", op.windowed(), "value_free", op.value_free(), "M", gen.M, flush=True)
x = torch.rand(gen.M, dtype=torch.float64, device=ctx.torch_device)
y = torch.empty_like(x)
for mode in sys.argv[1:] or ["0"]:
    var, _, dbg = mode.partition(":")
    os.environ["LZ_SELLW_VARIANT"] = var
    os.environ["LZ_SELLW_DEBUG"] = dbg or "0"
    for _ in range(3):
        op.apply(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.ExternalStream(ctx.stream_handle) if hasattr(ctx, "stream_handle") else None
    e0.record(); 
    for _ in range(20):
        op.apply(x, y)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        op.apply(x, y)
    torch.cuda.synchronize()
    print("mode", mode, "ms/apply %.4f" % ((time.perf_counter() - t0) / 20 * 1e3), flush=True)
