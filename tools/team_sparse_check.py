"""torchrun check of the row-sharded sparse path across real GPUs (one process per GPU)."""
import builtins, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
_p = builtins.print
builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and a[0].startswith("+++")) else _p(*a, **k)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import lanczos_b200 as lz
from lanczos_b200.team import TeamLanczos
from oracle import lanczos_oracle as orc
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
H = orc.rgg_graph_laplacian(npts, mean_degree=13.0, seed=4)
t = TeamLanczos(H)
t.execute_LanczosOld(40, seed=11)
ms = t.result.gpu_ms
if rank == 0:
    ref = orc.lanczos(H, 40, seed=11) if npts <= 300000 else None
    if ref is not None:
        ea = np.max(np.abs(np.diag(t.H_eff) - ref["alpha"]) / np.abs(ref["alpha"]))
        eb = np.max(np.abs(np.diag(t.H_eff, 1) - ref["beta"]) / np.abs(ref["beta"]))
        print(f"sparse team world={world} npts={npts}: alpha err {ea:.2e} beta err {eb:.2e}; {ms/40:.4f} ms/step")
        assert ea < 1e-12 and eb < 1e-12
    else:
        print(f"sparse team world={world} npts={npts}: {ms/40:.4f} ms/step")
del t
dist.barrier()
dist.destroy_process_group()
