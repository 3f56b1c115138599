"""Row-sharded solves on the GPU.  With one device the P shards are driven by one process
(LocalTeamLanczos: same kernels, same peer-memory exchange, push phase before combine phase);
with >= 2 devices the same test also runs across real GPUs, and through one process per GPU
(TeamLanczos, NCCL only for the handle swap)."""
import os
import socket

import numpy as np
import pytest

from oracle import lanczos_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lz():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lanczos_b200
    return lanczos_b200


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300))


CASES = [
    ((16, 12, 20), "periodic", 4, "full", 1),
    ((16, 12, 10), "dirichlet", 3, "full", 1),        # uneven slabs 4/3/3
    ((16, 12, 9), "periodic", 2, "full", 2),
    ((18, 7, 8), "periodic", 8, "none", 1),           # one plane per shard
    ((40, 36), "periodic", 4, "full", 1),             # 2-D: y-slabs
    ((40, 36), "dirichlet", 3, "selective", 2),
    ((15, 6, 8), "periodic", 2, "full", 1),           # odd nx: scalar path, odd plane
    ((257,), "dirichlet", 4, "full", 1),              # 1-D segments
    ((64, 16, 12), "periodic", 3, "selective", 2),    # whole tiles: alpha inside KB + border kernel, slab-top edge from the ghost plane
    ((64, 8, 8), "dirichlet", 2, "none", 1),
    ((128, 8, 6), "periodic", 6, "none", 1),          # one plane per shard
]


@pytest.mark.parametrize("grid,bc,world,reorth,passes", CASES)
def test_local_team_matches_single_gpu(lz, grid, bc, world, reorth, passes):
    from lanczos_b200.team import LocalTeamLanczos
    dim = len(grid)
    op = lz.StencilOperator(grid, 2.0 * dim + 0.25, [-1.0, -0.8, -1.1][:dim], bc=bc)
    n = 24
    one = lz.Lanczos(op)
    kba = (grid[0] % 64 == 0)        # whole-tile cases: alpha inside KB + border kernel (slab-top edge from the ghost plane)
    one.execute_Lanczos(n, seed=7, reorth=reorth, cgs_passes=passes)
    team = LocalTeamLanczos(op, world)
    team.execute_Lanczos(n, seed=7, reorth=reorth, cgs_passes=passes, kb_alpha=kba)
    assert team.result.alpha_in_update == (kba and reorth != "full")
    a1, b1 = np.diag(one.H_eff), np.diag(one.H_eff, 1)
    a2, b2 = np.diag(team.H_eff), np.diag(team.H_eff, 1)
    tol = 1e-12 if reorth != "none" else 1e-9
    assert rel(a2, a1) < tol
    assert rel(b2, b1) < tol
    # and against the oracle when the run follows the reference (full reorth)
    if reorth == "full" and passes == 1:
        H = orc.laplacian_csr(grid, 2.0 * dim + 0.25, [-1.0, -0.8, -1.1][:dim], periodic=(bc == "periodic"))
        ref = orc.lanczos(H, n, seed=7)
        assert rel(a2, ref["alpha"]) < 1e-12
        assert rel(b2, ref["beta"]) < 1e-12
        V = team.basis_rows_host()
        assert np.max(np.abs(V - ref["V"].T)) < 1e-12


@pytest.mark.parametrize("grid,bc,world", [((16, 12, 20), "periodic", 4), ((16, 12, 10), "dirichlet", 3), ((40, 36), "periodic", 4),
                                           ((15, 6, 8), "periodic", 2), ((64, 16, 9), "periodic", 3), ((64, 32, 4), "dirichlet", 4)])
def test_local_team_step_kernels_agree(lz, grid, bc, world):
    """Sharded runs: the recompute step (default for stencils; halo planes pushed after KB) and the
    two-pass step (halo planes stored by K3 itself) give the same tridiagonal matrix."""
    from lanczos_b200.team import LocalTeamLanczos
    dim = len(grid)
    op = lz.StencilOperator(grid, 2.0 * dim + 0.25, [-1.0, -0.8, -1.1][:dim], bc=bc)
    T = {}
    for kern in ("recompute", "two_pass"):
        team = LocalTeamLanczos(op, world)
        team.execute_Lanczos(20, seed=7, step_kernel=kern)
        assert team.result.step_kernel == kern
        T[kern] = team.H_eff.copy()
    assert rel(np.diag(T["recompute"]), np.diag(T["two_pass"])) < 1e-12
    assert rel(np.diag(T["recompute"], 1), np.diag(T["two_pass"], 1)) < 1e-12


def test_local_team_with_potential_and_clean_start(lz):
    from lanczos_b200.team import LocalTeamLanczos
    H, c, o, pot = orc.deuteron_hamiltonian(12)
    op = lz.StencilOperator((12, 12, 12), c, o, diag=pot)
    one = lz.Lanczos(op)
    one.execute_Lanczos(40, seed=78)
    team = LocalTeamLanczos(op, 3)
    team.execute_Lanczos(40, seed=78)
    assert rel(np.diag(team.H_eff), np.diag(one.H_eff)) < 1e-12
    assert rel(np.diag(team.H_eff, 1), np.diag(one.H_eff, 1)) < 1e-12
    v0 = orc.start_vector(12 ** 3, seed=5)
    one.execute_Lanczos(16, v0=v0, ref_compat=False)
    team.execute_Lanczos(16, v0=v0, ref_compat=False)
    assert rel(np.diag(team.H_eff), np.diag(one.H_eff)) < 1e-12
    V = team.basis_rows_host()
    assert np.max(np.abs(V[0] - v0)) < 1e-15


def test_team_results_identical_on_every_shard_and_reproducible(lz):
    from lanczos_b200.team import LocalTeamLanczos
    op = lz.StencilOperator((32, 16, 24), 6.0, -1.0)
    team = LocalTeamLanczos(op, 4)
    team.execute_Lanczos(20, seed=3)
    T1 = team.H_eff.copy()
    team.execute_Lanczos(20, seed=3)
    assert np.array_equal(T1, team.H_eff)          # rank-ordered sums: bit-reproducible


def test_multi_device_single_process(lz):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from lanczos_b200.team import LocalTeamLanczos
    op = lz.StencilOperator((64, 32, 48), 6.0, -1.0)
    one = lz.Lanczos(op)
    one.execute_Lanczos(30, seed=7)
    ndev = torch.cuda.device_count()
    team = LocalTeamLanczos(op, ndev, devices=list(range(ndev)))
    team.execute_Lanczos(30, seed=7)
    assert rel(np.diag(team.H_eff), np.diag(one.H_eff)) < 1e-12
    assert rel(np.diag(team.H_eff, 1), np.diag(one.H_eff, 1)) < 1e-12


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lanczos_b200 as lz
        from lanczos_b200.team import TeamLanczos
        op = lz.StencilOperator((64, 32, 48), 6.0, -1.0)
        t = TeamLanczos(op)
        t.execute_Lanczos(30, seed=7)
        t.execute_Lanczos(30, seed=7, reorth="selective", cgs_passes=2)
        if rank == 0:
            np.save(out, t.H_eff)
        del t
    finally:
        dist.destroy_process_group()


def test_one_process_per_gpu(lz, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    out = str(tmp_path / "T.npy")
    mp.spawn(_rank_main, args=(world, _free_port(), out), nprocs=world, join=True)
    T = np.load(out)
    one = lz.Lanczos(lz.StencilOperator((64, 32, 48), 6.0, -1.0))
    one.execute_Lanczos(30, seed=7, reorth="selective", cgs_passes=2)
    assert rel(np.diag(T), np.diag(one.H_eff)) < 1e-11


@pytest.mark.parametrize("fmt,world,reorth", [("sell", 3, "full"), ("csr", 2, "full"), ("sell", 4, "selective"), ("sell", 5, "none")])
def test_local_team_sparse_matches_single_gpu(lz, fmt, world, reorth):
    """Row-sharded sparse operator with the ghost-index exchange vs the single-GPU run and the oracle."""
    from lanczos_b200.team import LocalTeamLanczos
    H = orc.delaunay_graph_laplacian(6000, seed=5)
    n = 30
    one = lz.IrrLanczos(H)
    one.execute_LanczosOld(n, seed=3, reorth=reorth, cgs_passes=2 if reorth == "selective" else 1, fmt=fmt)
    team = LocalTeamLanczos(H, world, fmt=fmt)
    team.execute_LanczosOld(n, seed=3, reorth=reorth, cgs_passes=2 if reorth == "selective" else 1)
    tol = 1e-12 if reorth != "none" else 1e-9
    assert rel(np.diag(team.H_eff), np.diag(one.H_eff)) < tol
    assert rel(np.diag(team.H_eff, 1), np.diag(one.H_eff, 1)) < tol
    if reorth == "full":
        ref = orc.lanczos(H, n, seed=3)
        assert rel(np.diag(team.H_eff), ref["alpha"]) < 1e-12
        assert rel(np.diag(team.H_eff, 1), ref["beta"]) < 1e-12


def test_local_team_sparse_rgg(lz):
    from lanczos_b200.team import LocalTeamLanczos
    H = orc.rgg_graph_laplacian(20000, mean_degree=13.0, seed=4)
    ref = orc.lanczos(H, 25, seed=11)
    team = LocalTeamLanczos(H, 4)
    team.execute_LanczosOld(25, seed=11)
    assert rel(np.diag(team.H_eff), ref["alpha"]) < 1e-12
    assert rel(np.diag(team.H_eff, 1), ref["beta"]) < 1e-12


def test_local_team_stencil27(lz):
    from lanczos_b200.team import LocalTeamLanczos
    w = tuple(-v for v in orc.box27_weights(2.0))
    grid = (16, 10, 12)
    H = orc.laplacian27_csr(grid, w, periodic=True)
    ref = orc.lanczos(H, 20, seed=5)
    op = lz.StencilOperator(grid, 0.0, 0.0, weights27=w)
    team = LocalTeamLanczos(op, 3)
    team.execute_Lanczos(20, seed=5)
    assert rel(np.diag(team.H_eff), ref["alpha"]) < 1e-12
    assert rel(np.diag(team.H_eff, 1), ref["beta"]) < 1e-12


def _same_up_to_sign(A, B, tol):
    """columns of A and B agree up to a sign each"""
    for i in range(A.shape[1]):
        d = min(np.linalg.norm(A[:, i] - B[:, i]), np.linalg.norm(A[:, i] + B[:, i]))
        assert d < tol, (i, d)


def test_dropin_classes_row_sharded(lz, tmp_path, capsys):
    """Multi-GPU behind the reference API: Lanczos(H).execute_Lanczos(n, devices=[...]) and everything the
    reference class offers afterwards - H_eff, V, get_H_eigs / H_eigvecs (each shard lifts its rows),
    print_good_eigs (sharded operator apply, sums over the peer ring), ritz_vectors(k), checkpoint -
    against the single-GPU run of the same class (Lanczos.py:132-185)."""
    import torch
    from lanczos_b200 import io as lzio
    d = torch.cuda.current_device()
    ndev = torch.cuda.device_count()
    devices = [i % ndev for i in range(3)] if ndev >= 2 else [d, d, d]
    H, c, o, pot = orc.deuteron_hamiltonian(12)
    op = lz.StencilOperator((12, 12, 12), c, o, diag=pot)
    n = 60
    one = lz.Lanczos(op)
    one.execute_Lanczos(n, seed=78)
    sh = lz.Lanczos(op)
    sh.execute_Lanczos(n, seed=78, devices=devices)
    assert rel(np.diag(sh.H_eff), np.diag(one.H_eff)) < 1e-12 and rel(np.diag(sh.H_eff, 1), np.diag(one.H_eff, 1)) < 1e-12
    assert sh.V.shape == (12 ** 3, n) and np.max(np.abs(sh.V - one.V)) < 1e-11
    one.get_H_eigs()
    sh.get_H_eigs()                                   # includes the reference's normalisation / orthogonality asserts
    np.testing.assert_allclose(sh.H_eigvals, one.H_eigvals, rtol=1e-10, atol=1e-12)
    assert sh.H_eigvecs.shape == (12 ** 3, n)
    _same_up_to_sign(sh.H_eigvecs[:, :5], one.H_eigvecs[:, :5], 1e-8)
    ip1 = one.print_good_eigs(print_nr=5)
    ip2 = sh.print_good_eigs(print_nr=5)
    assert np.max(np.abs(ip1 - ip2)) < 1e-9
    assert "EIGENVALUE AND EIGVENVECTOR COMPARISON" in capsys.readouterr().out
    th1, Y1 = one.ritz_vectors(3)
    th2, Y2 = sh.ritz_vectors(3)
    assert isinstance(Y2, list) and len(Y2) == 3
    Y2h = np.concatenate([y.cpu().numpy() for y in Y2], axis=1)
    _same_up_to_sign(Y2h.T, Y1.cpu().numpy().T, 1e-8)
    # sharded checkpoint: one file per shard, restored onto the same partition and onto one GPU
    path = lzio.save_checkpoint(str(tmp_path / "run.npz"), sh)
    for r in range(3):
        assert os.path.exists(lzio.shard_path(str(tmp_path / "run"), r))
    back = lz.Lanczos(op)
    lzio.restore_checkpoint(back, path, devices=devices)
    back.get_H_eigs()
    np.testing.assert_allclose(back.H_eigvals, one.H_eigvals, rtol=1e-10, atol=1e-12)
    _same_up_to_sign(back.H_eigvecs[:, :3], one.H_eigvecs[:, :3], 1e-8)
    flat = lz.Lanczos(op)
    lzio.restore_checkpoint(flat, path)
    assert np.max(np.abs(flat.V - one.V)) < 1e-11
    # irregular class, sparse row blocks
    G = orc.delaunay_graph_laplacian(5000, seed=2)
    a = lz.IrrLanczos(G)
    a.execute_LanczosOld(30, seed=5)
    b = lz.IrrLanczos(G)
    b.execute_LanczosOld(30, seed=5, devices=devices[:2])
    assert rel(np.diag(b.H_eff), np.diag(a.H_eff)) < 1e-12
    a.get_H_eigs()
    b.get_H_eigs()
    _same_up_to_sign(b.H_eigvecs[:, -3:], a.H_eigvecs[:, -3:], 1e-8)
    assert np.max(np.abs(a.print_good_eigs(print_nr=3) - b.print_good_eigs(print_nr=3))) < 1e-9


def _rank_dropin(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lanczos_b200 as lz
        op = lz.StencilOperator((64, 32, 48), 6.0, -1.0)
        L = lz.Lanczos(op)
        L.execute_Lanczos(30, seed=7, devices="auto", verbose=False)
        L.get_H_eigs()
        ip = L.print_good_eigs(print_nr=2)
        V = L.V                           # assembling the global basis is a collective: every rank asks for it
        if rank == 0:
            np.savez(out, T=L.H_eff, theta=L.H_eigvals, Y=L.H_eigvecs[:, :3], ip=ip, V=V[:, :4])
        del L
    finally:
        dist.destroy_process_group()


def test_dropin_one_process_per_gpu(lz, tmp_path):
    """devices="auto" under a process group: the torchrun form of the same drop-in call."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    out = str(tmp_path / "dropin.npz")
    mp.spawn(_rank_dropin, args=(world, _free_port(), out), nprocs=world, join=True)
    z = np.load(out)
    one = lz.Lanczos(lz.StencilOperator((64, 32, 48), 6.0, -1.0))
    one.execute_Lanczos(30, seed=7)
    one.get_H_eigs()
    assert rel(np.diag(z["T"]), np.diag(one.H_eff)) < 1e-12
    np.testing.assert_allclose(z["theta"], one.H_eigvals, rtol=1e-10, atol=1e-12)
    assert np.max(np.abs(z["V"] - one.V[:, :4])) < 1e-11
    assert np.max(np.abs(z["ip"] - one.print_good_eigs(print_nr=2))) < 1e-9


@pytest.mark.parametrize("world,reorth,passes,sigma", [(3, "selective", 2, 64), (2, "none", 1, 64), (4, "selective", 1, 64),
                                                        (2, "selective", 2, 384), (3, "none", 1, 96), (2, "none", 1, 2048)])
def test_sparse_shards_overlap_interior_apply_with_exchange(lz, world, reorth, passes, sigma):
    """Row shards of a SELL operator: the spans whose rows touch no ghost column are applied while the ghost
    entries and the beta sum travel on a second stream (lz_run_info.overlap); against the same run without
    the overlap, the single-GPU run and the oracle.  Small sigma so that these small shards have interior
    spans at all (and windows that are not a power of two / not a multiple of 8 chunks); select_tol small enough that sweeps fire (the early interior apply is then redone)."""
    from lanczos_b200.team import LocalTeamLanczos
    H = orc.rgg_graph_laplacian(24000, mean_degree=13.0, seed=4)        # cell-ordered: ghosts only near block ends
    n = 30
    kw = dict(reorth=reorth, cgs_passes=passes, select_tol=1e-12 if reorth == "selective" else 0.0)
    one = lz.IrrLanczos(H)
    one.execute_LanczosOld(n, seed=3, sigma=sigma, **kw)
    T = {}
    for ov in (True, False):
        team = LocalTeamLanczos(H, world, fmt="sell", sigma=sigma)
        team.execute_LanczosOld(n, seed=3, overlap=ov, **kw)
        assert team.result.overlap == ov
        if reorth == "selective":
            assert team.result.reorth_count > 0
        T[ov] = team.H_eff.copy()
        team.execute_LanczosOld(n, seed=3, overlap=ov, **kw)
        assert np.array_equal(T[ov], team.H_eff)                        # two streams, still bit-reproducible
    tol = 1e-12 if reorth != "none" else 1e-9
    assert rel(np.diag(T[True]), np.diag(T[False])) < tol and rel(np.diag(T[True], 1), np.diag(T[False], 1)) < tol
    assert rel(np.diag(T[True]), np.diag(one.H_eff)) < tol and rel(np.diag(T[True], 1), np.diag(one.H_eff, 1)) < tol
    if reorth == "selective":
        ref = orc.lanczos(H, n, seed=3)
        assert rel(np.diag(T[True])[:12], ref["alpha"][:12]) < 1e-11


@pytest.mark.parametrize("world,reorth", [(2, "selective"), (3, "none")])
def test_sparse_shards_windowed_interior(lz, monkeypatch, world, reorth):
    """Row shards in the windowed SELL form (csrc/sellw.cu): the interior windows run from the shared-memory stage
    with 16-bit offsets, the windows with ghost columns stay with the plain kernel in pieces."""
    from lanczos_b200.team import LocalTeamLanczos
    H = orc.banded_graph_laplacian(60_011, far=(3000,), seed=9)
    n = 30
    kw = dict(reorth=reorth, cgs_passes=2, select_tol=1e-12 if reorth == "selective" else 0.0)
    monkeypatch.setenv("LZ_SELL_WINDOW", "0")
    plain = LocalTeamLanczos(H, world, fmt="sell", sigma=256)
    plain.execute_LanczosOld(n, seed=3, **kw)
    assert plain.windowed_local() == 0
    monkeypatch.setenv("LZ_SELL_WINDOW", "1")
    monkeypatch.setenv("LZ_SELL_WINDOW_MIN", "1")
    for ov in (True, False):
        team = LocalTeamLanczos(H, world, fmt="sell", sigma=256)
        team.execute_LanczosOld(n, seed=3, overlap=ov, **kw)
        assert team.windowed_local() > 0 and team.result.overlap == ov
        tol = 1e-12 if reorth != "none" else 1e-9
        assert rel(np.diag(team.H_eff), np.diag(plain.H_eff)) < tol
        assert rel(np.diag(team.H_eff, 1), np.diag(plain.H_eff, 1)) < tol
        T = team.H_eff.copy()
        team.execute_LanczosOld(n, seed=3, overlap=ov, **kw)
        assert np.array_equal(T, team.H_eff)
