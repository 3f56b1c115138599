"""torchrun check of the block-wise sharded sparse path on real GPUs (one process per GPU): every
rank generates only its own row block of the config-4 graph on its device; the team solve must
equal the oracle loop on the assembled matrix (small boxes) or a single-GPU solve (large boxes)."""
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
from lanczos_b200 import synth
from lanczos_b200.engine import DeviceCSR
from lanczos_b200.team import TeamLanczos
cells = tuple(int(c) for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else (12, 11, 16)))
n = 30
gen = synth.RggGenerator(cells, seed=3)
t = TeamLanczos(gen.row_block(rank, world))
t.execute_LanczosOld(n, seed=11)
ms = t.result.gpu_ms
if rank == 0:
    if gen.M <= 20000:
        from oracle import lanczos_oracle as orc
        from oracle import rgg_oracle as rgg
        L, _, _ = rgg.rgg_laplacian(cells, synth.MEAN_DEGREE_LAMBDA, 3)
        ref = orc.lanczos(L, n, seed=11)
        ra, rb, how = ref["alpha"], ref["beta"], "oracle"
    else:
        S = lz.IrrLanczos(DeviceCSR(*gen.rows(0, gen.M)))
        S.execute_LanczosOld(n, seed=11)
        ra, rb, how = np.diag(S.H_eff), np.diag(S.H_eff, 1), "single GPU"
    ea = np.max(np.abs(np.diag(t.H_eff) - ra) / np.abs(ra))
    eb = np.max(np.abs(np.diag(t.H_eff, 1) - rb) / np.abs(rb))
    print(f"rgg team world={world} M={gen.M} vs {how}: alpha err {ea:.2e} beta err {eb:.2e}; {ms/n:.4f} ms/step", flush=True)
    assert ea < 1e-12 and eb < 1e-12
del t
dist.barrier()
dist.destroy_process_group()
