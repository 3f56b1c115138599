"""Config 1 (200 x 200, n = 100, full reorth in the reference's form): the persistent kernel against the replayed graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lanczos_b200 as lz

op = lz.StencilOperator((200, 200), 4.0, -1.0, bc="dirichlet")
for persistent in ([True, False] if len(sys.argv) < 2 else [sys.argv[1] == "1"]):
    L = lz.Lanczos(op)
    for _ in range(3):
        L.execute_Lanczos(100, seed=3, persistent=persistent, verbose=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    R = 20
    for _ in range(R):
        L.execute_Lanczos(100, seed=3, persistent=persistent, verbose=False)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / R
    print("persistent", persistent, "kernel", L.result.step_kernel, "graph", getattr(L.result, "graph", None),
          "us/step wall %.2f" % (dt / 100 * 1e6), "loop_ms", getattr(L.result, "loop_ms", None), flush=True)
