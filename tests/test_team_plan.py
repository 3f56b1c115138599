"""Row-sharding logic on the CPU.  (1) SlabPlan invariants.  (2) world_size-2/3 gloo runs: every
rank holds only its slab, gets its two ghost planes from the ranks SlabPlan names, and the scalar
sums go through all_reduce - the sharded loop must reproduce the single-process oracle.  This is
the host-side contract that the CUDA team path (csrc/lanczos.cu, csrc/peer.cuh) implements on
NVLink peer memory."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lanczos_b200.team import SlabPlan
from oracle import lanczos_oracle as orc


@pytest.mark.parametrize("grid,world,periodic", [((8, 6, 10), 4, True), ((8, 6, 10), 3, False),
                                                  ((16, 9), 2, True), ((33,), 4, False), ((4, 4, 8), 8, True)])
def test_slab_plan_partitions_rows(grid, world, periodic):
    plan = SlabPlan(grid, world, periodic)
    assert plan.M == int(np.prod(grid))
    rows = [plan.rows(r) for r in range(world)]
    assert rows[0][0] == 0 and rows[-1][1] == plan.M
    for a, b in zip(rows[:-1], rows[1:]):
        assert a[1] == b[0]
    sizes = [plan.local_rows(r) for r in range(world)]
    assert max(sizes) - min(sizes) <= plan.plane            # balanced to within one plane
    nz = plan.grid3[2]
    for r in range(world):
        lo, up = plan.neighbours(r)
        zlo, zup = plan.halo_planes(r)
        z0, z1 = plan.slab(r)
        if periodic:
            assert zlo == (z0 - 1) % nz and zup == z1 % nz
            assert plan.slab(lo)[0] <= zlo < plan.slab(lo)[1]
            assert plan.slab(up)[0] <= zup < plan.slab(up)[1]
        else:
            assert (lo == -1) == (r == 0) and (up == -1) == (r == world - 1)
            assert (zlo is None) == (r == 0) and (zup is None) == (r == world - 1)


def test_slab_plan_rejects_too_many_ranks():
    with pytest.raises(ValueError):
        SlabPlan((4, 4, 2), 3, True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sharded_apply(plan, rank, x_loc, center, off3, periodic):
    """y_loc = (H x)_loc with ghost planes fetched from the neighbours named by the plan."""
    nx, ny, _ = plan.grid3
    nzl = plan.counts[rank]
    X = x_loc.reshape(nzl, ny, nx)
    lo, up = plan.neighbours(rank)
    ghost_lo = np.zeros((ny, nx))
    ghost_hi = np.zeros((ny, nx))
    reqs = []
    # my first plane goes to the lower neighbour (it is the plane above its slab), my last plane up
    if lo >= 0:
        reqs.append(dist.isend(torch.from_numpy(X[0].copy()), dst=lo, tag=1))
    if up >= 0:
        reqs.append(dist.isend(torch.from_numpy(X[-1].copy()), dst=up, tag=2))
    if up >= 0:
        t = torch.zeros(ny, nx, dtype=torch.float64)
        dist.recv(t, src=up, tag=1)
        ghost_hi = t.numpy()
    if lo >= 0:
        t = torch.zeros(ny, nx, dtype=torch.float64)
        dist.recv(t, src=lo, tag=2)
        ghost_lo = t.numpy()
    for q in reqs:
        q.wait()
    Xp = np.concatenate([ghost_lo[None], X, ghost_hi[None]], axis=0)
    Y = center * X + off3[2] * (Xp[:-2] + Xp[2:])
    for ax, o, n_ax in ((2, off3[0], nx), (1, off3[1], ny)):
        if o == 0.0:
            continue
        if periodic:
            Y = Y + o * (np.roll(X, 1, axis=ax) + np.roll(X, -1, axis=ax))
        else:
            P = np.zeros_like(X)
            sl_a = [slice(None)] * 3
            sl_b = [slice(None)] * 3
            sl_a[ax], sl_b[ax] = slice(1, None), slice(None, -1)
            P[tuple(sl_a)] += X[tuple(sl_b)]
            P[tuple(sl_b)] += X[tuple(sl_a)]
            Y = Y + o * P
    return Y.reshape(-1)


def _allsum(v):
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t.item())


def _worker(rank, world, port, grid, periodic, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = SlabPlan(grid, world, periodic)
        center, off = 2.0 * len(grid), [-1.0] * len(grid)
        off3 = plan.off3(off)
        r0, r1 = plan.rows(rank)
        v0 = orc.start_vector(plan.M, seed=99)[r0:r1]
        # the reference loop (Lanczos.py:104-119) on row shards: dots -> all_reduce
        apply = lambda x: _sharded_apply(plan, rank, x, center, off3, periodic)
        V = np.zeros((n, r1 - r0))
        alpha, beta = np.zeros(n), np.zeros(n - 1)
        r = apply(v0)
        a = _allsum(np.dot(r, v0))
        r = r - a * v0
        for j in range(n):
            beta[j - 1] = np.sqrt(_allsum(np.dot(r, r)))
            V[j] = r / beta[j - 1]
            ip = np.array([_allsum(np.dot(V[j], V[i])) for i in range(n)])
            V[j] = 2 * V[j] - (ip[:, None] * V).sum(axis=0)
            r = apply(V[j])
            alpha[j] = _allsum(np.dot(V[j], r))
            r = r - V[j] * alpha[j] - V[j - 1] * beta[j - 1]
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), alpha=alpha, beta=beta)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid,world,periodic", [((6, 5, 8), 2, True), ((6, 5, 7), 3, False), ((10, 9), 2, True)])
def test_sharded_loop_matches_oracle_gloo(tmp_path, grid, world, periodic):
    n = 12
    port = _free_port()
    mp.spawn(_worker, args=(world, port, grid, periodic, n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    H = orc.laplacian_csr(grid, 2.0 * len(grid), -1.0, periodic=periodic)
    ref = orc.lanczos(H, n, seed=99)
    assert np.max(np.abs(got["alpha"] - ref["alpha"]) / np.abs(ref["alpha"])) < 1e-12
    assert np.max(np.abs(got["beta"] - ref["beta"]) / np.abs(ref["beta"])) < 1e-12


# ---- sparse operators: contiguous row blocks + ghost-index exchange -------------------------------

@pytest.mark.parametrize("world", [2, 3, 5])
def test_row_block_plan_reproduces_spmv(world):
    """Single-process check of the plan: local blocks with renumbered columns, fed with the ghost
    entries that the send lists deliver, reproduce H @ x exactly."""
    from lanczos_b200.team import RowBlockPlan
    H = orc.delaunay_graph_laplacian(1500, seed=2)
    plan = RowBlockPlan(H, world)
    x = np.random.RandomState(0).uniform(-1, 1, H.shape[0])
    gather = [np.full(plan.nghost_max, np.nan) for _ in range(world)]
    for r in range(world):                                   # every rank pushes what the others need
        r0, r1 = plan.rows(r)
        send, seg, off = plan.send_lists(r)
        assert seg[0] == 0 and seg[-1] == len(send)
        for q in range(world):
            vals = x[r0:r1][send[seg[q]:seg[q + 1]]]
            gather[q][off[q]:off[q] + len(vals)] = vals
    y = np.zeros_like(x)
    for r in range(world):
        r0, r1 = plan.rows(r)
        indptr, indices, data, ncols = plan.local_csr(r)
        ng = len(plan.ghost_cols[r])
        assert ncols == (r1 - r0) + ng
        assert not np.isnan(gather[r][:ng]).any()            # every ghost entry was delivered
        xloc = np.concatenate([x[r0:r1], gather[r][:ng]])
        import scipy.sparse as sp
        y[r0:r1] = sp.csr_matrix((data, indices, indptr), shape=(r1 - r0, ncols)) @ xloc
    assert np.array_equal(y, H @ x)


def _sparse_worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        from lanczos_b200.team import RowBlockPlan
        H = orc.delaunay_graph_laplacian(1200, seed=4)
        plan = RowBlockPlan(H, world)
        r0, r1 = plan.rows(rank)
        indptr, indices, data, ncols = plan.local_csr(rank)
        A = sp.csr_matrix((data, indices, indptr), shape=(r1 - r0, ncols))
        send, seg, off = plan.send_lists(rank)
        ng = len(plan.ghost_cols[rank])
        segs = plan.owner_segments(rank)

        def apply(x_loc):
            ghosts = np.zeros(ng)
            reqs = []
            for q in range(world):
                if q != rank and seg[q + 1] > seg[q]:
                    reqs.append(dist.isend(torch.from_numpy(x_loc[send[seg[q]:seg[q + 1]]].copy()), dst=q, tag=7))
            for q, a, b in segs:
                if q != rank and b > a:
                    t = torch.zeros(b - a, dtype=torch.float64)
                    dist.recv(t, src=q, tag=7)
                    ghosts[a:b] = t.numpy()
            for rq in reqs:
                rq.wait()
            return A @ np.concatenate([x_loc, ghosts])

        v0 = orc.start_vector(plan.M, seed=99)[r0:r1]
        V = np.zeros((n, r1 - r0))
        alpha, beta = np.zeros(n), np.zeros(n - 1)
        r = apply(v0)
        a = _allsum(np.dot(r, v0))
        r = r - a * v0
        for j in range(n):
            beta[j - 1] = np.sqrt(_allsum(np.dot(r, r)))
            V[j] = r / beta[j - 1]
            ip = np.array([_allsum(np.dot(V[j], V[i])) for i in range(n)])
            V[j] = 2 * V[j] - (ip[:, None] * V).sum(axis=0)
            r = apply(V[j])
            alpha[j] = _allsum(np.dot(V[j], r))
            r = r - V[j] * alpha[j] - V[j - 1] * beta[j - 1]
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), alpha=alpha, beta=beta)
    finally:
        dist.destroy_process_group()


def test_sharded_sparse_loop_matches_oracle_gloo(tmp_path):
    n, world = 10, 2
    mp.spawn(_sparse_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    ref = orc.lanczos(orc.delaunay_graph_laplacian(1200, seed=4), n, seed=99)
    assert np.max(np.abs(got["alpha"] - ref["alpha"]) / np.abs(ref["alpha"])) < 1e-12
    assert np.max(np.abs(got["beta"] - ref["beta"]) / np.abs(ref["beta"])) < 1e-12
