"""Time one SpMV apply on the config-4 graph (50 M vertices) in the windowed SELL form.

    python tools/sellw_time.py [variant ...]     variant 0: two CTAs per SM, one stage each (default)
                                                 variant 1: one CTA per SM, two stages (the fallback)
    LZ_SELL_WINDOW=0 python tools/sellw_time.py  the plain SELL kernel
    LZ_SELLW_BANKS=0 ...                         entries of a row in column order
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lanczos_b200 import engine, synth  # noqa: E402
from lanczos_b200.engine import DeviceCSR  # noqa: E402

ctx = engine.Context.default()
gen = synth.RggGenerator((253, 253, 252), seed=0)
H = DeviceCSR(*gen.rows(0, gen.M))
op = engine.as_device_operator(H, ctx, fmt="sell")
print("windowed", op.windowed(), "value_free", op.value_free(), "M", gen.M, flush=True)
x = torch.rand(gen.M, dtype=torch.float64, device=ctx.torch_device)
y = torch.empty_like(x)
for variant in sys.argv[1:] or ["0"]:
    os.environ["LZ_SELLW_VARIANT"] = variant
    for _ in range(5):
        op.apply(x, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        op.apply(x, y)
    torch.cuda.synchronize()
    print("variant", variant, "ms/apply %.4f" % ((time.perf_counter() - t0) / 20 * 1e3), flush=True)
