"""Development aid: row-sharded two-pass / recompute runs on one GPU, repeated, against the single-shard result.
With LZ_PDL=1 (programmatic dependent launch on) the two-pass runs go wrong from step 2-3 on, nondeterministically;
with the attribute off (the default) every line ends in `first bad []`.  See pdl_enabled() in csrc/capi.cu."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import lanczos_b200 as lz
from lanczos_b200.team import LocalTeamLanczos
import builtins
_p = builtins.print
builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and a[0].startswith("+++")) else _p(*a, **k)
grid=(16,12,20)
op = lz.StencilOperator(grid, 6.25, [-1.0,-0.8,-1.1], bc="periodic")
one = lz.Lanczos(op); one.execute_Lanczos(20, seed=7, step_kernel="two_pass")
a1=np.diag(one.H_eff).copy()
for world in (1,2,4):
  for kern in ("recompute","two_pass"):
    for rep in range(3):
        team = LocalTeamLanczos(op, world)
        team.execute_Lanczos(20, seed=7, step_kernel=kern)
        a=np.diag(team.H_eff)
        bad=np.where(np.abs(a-a1)>1e-9)[0]
        print(world, kern, rep, "first bad", bad[:3], flush=True)
