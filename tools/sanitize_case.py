"""Small cases for compute-sanitizer (tools/sanitize.sh): every kernel family once, sized so that a
run under memcheck / racecheck / synccheck / initcheck finishes in a minute or two.

  smoke      __graft_entry__.smoke(): stencil + SELL + device-generated graph, each checked against the oracle
  team       a 2-shard and a 3-shard LocalTeamLanczos solve (peer ring: flags, slots, halo planes, ghost gather)
  recompute  whole-tile grid (lean KA2 + KB + border kernel), selective re-orthogonalisation that fires, CGS2 with K4c (TMA)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300)))


def case_smoke():
    import __graft_entry__ as g
    g.smoke()


def case_team():
    import lanczos_b200 as lz
    from lanczos_b200.team import LocalTeamLanczos
    from oracle import lanczos_oracle as orc
    grid, n = (16, 12, 10), 16
    H = orc.laplacian_csr(grid, 6.25, -1.0, periodic=True)
    ref = orc.lanczos(H, n, seed=7)
    op = lz.StencilOperator(grid, 6.25, -1.0)
    for world, reorth, passes in ((2, "full", 1), (3, "selective", 2)):
        t = LocalTeamLanczos(op, world)
        t.execute_Lanczos(n, seed=7, reorth=reorth, cgs_passes=passes, select_tol=1e-14 if reorth == "selective" else 0.0)
        e = rel(np.diag(t.H_eff), ref["alpha"])
        assert e < 1e-11, (world, e)
        print(f"team stencil world={world} reorth={reorth}: alpha err {e:.1e}")
    G = orc.delaunay_graph_laplacian(1500, seed=3)
    refg = orc.lanczos(G, 12, seed=5)
    t = LocalTeamLanczos(G, 2)
    t.execute_Lanczos(12, seed=5)
    e = rel(np.diag(t.H_eff), refg["alpha"])
    assert e < 1e-11, e
    print(f"team sparse world=2: alpha err {e:.1e}")


def case_recompute():
    import lanczos_b200 as lz
    from oracle import lanczos_oracle as orc
    grid, n = (64, 16, 6), 20
    H = orc.laplacian_csr(grid, 6.0, -1.0, periodic=True)
    ref = orc.lanczos(H, n, seed=3)
    op = lz.StencilOperator(grid, 6.0, -1.0)
    L = lz.Lanczos(op)
    for kw in (dict(reorth="full", use_cuda=False), dict(reorth="selective", cgs_passes=2, select_tol=1e-13, kb_alpha=True),
               dict(reorth="full", cgs_passes=2), dict(reorth="none", keep_basis=False, kb_alpha=True)):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            L.execute_Lanczos(n, seed=3, verbose=False, **kw)
        e = rel(np.diag(L.H_eff)[:8], ref["alpha"][:8])
        assert e < 1e-9, (kw, e)
        print(f"recompute {kw}: step={L.result.step_kernel} alpha[:8] err {e:.1e} reorths {L.result.reorth_count}")
    L.get_H_eigs()


if __name__ == "__main__":
    {"smoke": case_smoke, "team": case_team, "recompute": case_recompute}[sys.argv[1]]()
    print("case", sys.argv[1], "ok")
