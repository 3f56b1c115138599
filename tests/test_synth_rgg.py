"""BASELINE config 4 input: the hashed random-geometric-graph Laplacian (include/lz_synth.h).

CPU: the NumPy restatement (oracle/rgg_oracle.py) is a graph Laplacian with the advertised density,
liblz_synth.so exports what its header declares, and BlockPlan (row blocks held rank by rank) obeys
the same exchange contract as RowBlockPlan - single process and over gloo with world_size 2.
GPU: the device generator reproduces the oracle's coordinates and sparsity pattern bit for bit, the
device SELL conversion equals the host one, and sharded solves on device-generated blocks agree with
the single-GPU solve and the oracle loop.
"""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lanczos_oracle as orc
from oracle import rgg_oracle as rgg
from lanczos_b200.team import BlockPlan, RowBlock, RowBlockPlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAM = 13.0 / (4.0 * np.pi / 3.0)


# ------------------------------------------------------------------------------------------- CPU
def test_oracle_rgg_is_a_graph_laplacian():
    L, prefix, p = rgg.rgg_laplacian((7, 6, 5), LAM, seed=11)
    M = L.shape[0]
    assert prefix[-1] == M == len(p)
    assert (L != L.T).nnz == 0
    assert np.abs(np.asarray(L.sum(axis=1))).max() == 0.0
    assert L.has_sorted_indices and L.indices.dtype == np.int32
    assert np.all(np.diff(L.indptr) >= 1)                     # every diagonal is stored
    assert np.array_equal(L.diagonal(), np.diff(L.indptr) - 1)
    # all points sit inside their cell
    cell = np.repeat(np.arange(len(prefix) - 1), np.diff(prefix))
    assert np.array_equal(np.floor(p[:, 0]).astype(int), cell % 7)
    assert np.array_equal(np.floor(p[:, 2]).astype(int), cell // 42)


def test_oracle_rgg_density_matches_config4():
    cnt = rgg.cell_counts((48, 48, 48), LAM, seed=0)
    assert abs(cnt.mean() - LAM) < 0.02 and cnt.max() <= 31
    assert abs(cnt.var() - LAM) < 0.1                         # Poisson: variance = mean
    L, _, _ = rgg.rgg_laplacian((10, 10, 10), LAM, seed=0)
    interior = L.diagonal()
    # boundary cells lose neighbours; the bulk sits at ~13 neighbours => ~14 entries per row
    assert 9.0 < interior.mean() < 13.5


def test_synth_library_exports_header_symbols():
    from lanczos_b200 import build as lzbuild, synth
    path = lzbuild.build_synth()
    lib = ctypes.CDLL(path)
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lz_synth.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(lzs_[a-z0-9_]+)\s*\(", text)))
    assert declared == sorted(synth.SIGNATURES)
    assert all(hasattr(lib, s) for s in declared)
    assert ctypes.sizeof(synth.RggParams) == 4 * 4 + 8 + 8 + 32 * 8
    assert synth.poisson_cdf(LAM) == list(rgg.poisson_cdf(LAM))


def _blocks_of(H, starts):
    out = []
    for r in range(len(starts) - 1):
        r0, r1 = starts[r], starts[r + 1]
        lo, hi = H.indptr[r0], H.indptr[r1]
        out.append(RowBlock(H.shape[0], starts, r, (H.indptr[r0:r1 + 1] - lo).astype(np.int32),
                            H.indices[lo:hi].astype(np.int32), H.data[lo:hi].copy()))
    return out


def _exchange_and_apply(plan, x):
    world = plan.world
    gather = [np.full(max(plan.nghost_max, 1), np.nan) for _ in range(world)]
    for r in range(world):
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
        assert ncols == (r1 - r0) + ng and not np.isnan(gather[r][:ng]).any()
        A = sp.csr_matrix((np.asarray(data), np.asarray(indices), np.asarray(indptr)), shape=(r1 - r0, ncols))
        y[r0:r1] = A @ np.concatenate([x[r0:r1], gather[r][:ng]])
    return y


@pytest.mark.parametrize("starts", [[0, 131, 396], [0, 100, 101, 250, 396], [0, 396]])
def test_block_plan_reproduces_spmv(starts):
    L, _, _ = rgg.rgg_laplacian((6, 5, 4), LAM, seed=7)
    assert L.shape[0] == 396
    plan = BlockPlan(_blocks_of(L, starts), len(starts) - 1)
    x = np.random.RandomState(1).uniform(-1, 1, L.shape[0])
    assert np.array_equal(_exchange_and_apply(plan, x), L @ x)


def test_block_plan_equals_row_block_plan():
    H = orc.delaunay_graph_laplacian(1500, seed=2)
    ref = RowBlockPlan(H, 3)
    plan = BlockPlan(_blocks_of(H, ref.starts), 3)
    for r in range(3):
        assert np.array_equal(plan.ghost_cols[r], ref.ghost_cols[r])
        for a, b in zip(plan.local_csr(r)[:3], ref.local_csr(r)[:3]):
            assert np.array_equal(np.asarray(a), b)
        for a, b in zip(plan.send_lists(r), ref.send_lists(r)):
            assert np.array_equal(a, b)


def test_block_plan_rejects_bad_partitions():
    L, _, _ = rgg.rgg_laplacian((4, 4, 4), LAM, seed=1)
    M = L.shape[0]
    with pytest.raises(ValueError):
        BlockPlan(_blocks_of(L, [0, M // 2, M]), 3)
    with pytest.raises(ValueError):
        BlockPlan(_blocks_of(L, [0, M // 2, M])[:1], 2)          # the other rank's ghost list is missing


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _block_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L, _, _ = rgg.rgg_laplacian((6, 5, 4), LAM, seed=7)
        starts = [0, 190, L.shape[0]]
        mine = _blocks_of(L, starts)[rank]                    # this rank only ever sees its own block

        def gather(d):
            parts = [None] * world
            dist.all_gather_object(parts, d)
            out = {}
            for p in parts:
                out.update(p)
            return out

        plan = BlockPlan([mine], world, gather=gather)
        send, seg, off = plan.send_lists(rank)
        ng = len(plan.ghost_cols[rank])
        indptr, indices, data, ncols = plan.local_csr(rank)
        A = sp.csr_matrix((data.numpy(), indices.numpy(), indptr.numpy()), shape=(plan.local_rows(rank), ncols))
        x = np.random.RandomState(5).uniform(-1, 1, L.shape[0])
        r0, r1 = plan.rows(rank)
        other = 1 - rank
        ghosts = torch.zeros(ng, dtype=torch.float64)
        req = dist.isend(torch.from_numpy(x[r0:r1][send[seg[other]:seg[other + 1]]].copy()), dst=other, tag=3)
        dist.recv(ghosts, src=other, tag=3)
        req.wait()
        y = A @ np.concatenate([x[r0:r1], ghosts.numpy()])
        np.save(os.path.join(out_dir, f"y{rank}.npy"), y)
    finally:
        dist.destroy_process_group()


def test_block_plan_over_gloo(tmp_path):
    mp.spawn(_block_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    L, _, _ = rgg.rgg_laplacian((6, 5, 4), LAM, seed=7)
    x = np.random.RandomState(5).uniform(-1, 1, L.shape[0])
    y = np.concatenate([np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")])
    assert np.array_equal(y, L @ x)


# ------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("cells,seed", [((6, 5, 4), 7), ((9, 3, 5), 123456789), ((1, 1, 6), 3), ((12, 11, 10), 0)])
def test_device_generator_matches_oracle_bit_for_bit(cells, seed):
    from lanczos_b200 import synth
    L, prefix, p = rgg.rgg_laplacian(cells, LAM, seed)
    g = synth.RggGenerator(cells, LAM, seed)
    assert g.M == L.shape[0]
    assert np.array_equal(g.prefix.cpu().numpy(), prefix)
    assert np.array_equal(g.positions().cpu().numpy(), p)
    indptr, indices, data = g.rows(0, g.M)
    assert np.array_equal(indptr.cpu().numpy(), L.indptr)
    assert np.array_equal(indices.cpu().numpy(), L.indices)
    assert np.array_equal(data.cpu().numpy(), L.data)
    # a row block in the middle carries global columns
    a, b = g.M // 3, 2 * g.M // 3
    ip, idx, dat = g.rows(a, b)
    lo, hi = L.indptr[a], L.indptr[b]
    assert np.array_equal(ip.cpu().numpy(), L.indptr[a:b + 1] - lo)
    assert np.array_equal(idx.cpu().numpy(), L.indices[lo:hi])
    assert np.array_equal(dat.cpu().numpy(), L.data[lo:hi])


@gpu
@pytest.mark.parametrize("fmt", ["sell", "csr"])
def test_device_csr_operator_equals_host_built_operator(fmt):
    """lz_op_csr_create_dev: same operator (exported CSR, SpMV bits) as the host conversion."""
    import lanczos_b200 as lz
    from lanczos_b200 import engine
    H = orc.delaunay_graph_laplacian(3000, seed=9)
    ctx = lz.Context.default()
    dev = ctx.torch_device
    op_h = lz.DeviceOperator.from_scipy(ctx, H, fmt=fmt, sigma=64)
    op_d = lz.DeviceOperator.from_device_csr(ctx, torch.from_numpy(H.indptr).to(dev), torch.from_numpy(H.indices).to(dev),
                                             torch.from_numpy(H.data).to(dev), fmt=fmt, sigma=64)
    assert op_h.nnz() == op_d.nnz()
    Eh, Ed = op_h.export_csr(), op_d.export_csr()
    assert np.array_equal(Eh.indptr, Ed.indptr) and np.array_equal(Eh.indices, Ed.indices)
    assert np.array_equal(Eh.data, Ed.data)
    x = np.random.RandomState(0).uniform(-1, 1, H.shape[0])
    assert np.array_equal(op_h.apply_host(x), op_d.apply_host(x))
    # ragged: empty rows, one long row, a row count that is not a multiple of 32
    R = sp.random(1003, 1003, density=0.004, random_state=3, format="lil")
    R[17, :] = 1.0
    R = sp.csr_matrix(R + R.T)
    R.sort_indices()
    ip, idx, dat = (torch.from_numpy(a).to(dev) for a in (R.indptr.astype(np.int32), R.indices.astype(np.int32), R.data))
    od = lz.DeviceOperator.from_device_csr(ctx, ip, idx, dat, fmt=fmt, sigma=96)
    oh = lz.DeviceOperator.from_scipy(ctx, R, fmt=fmt, sigma=96)
    xr = np.random.RandomState(1).uniform(-1, 1, 1003)
    assert np.array_equal(od.apply_host(xr), oh.apply_host(xr))
    assert od.nnz() == oh.nnz()
    with pytest.raises(ValueError):
        bad = idx.clone()
        bad[5] = 5000
        lz.DeviceOperator.from_device_csr(ctx, ip, bad, dat, fmt=fmt)


@gpu
def test_device_csr_drop_in_matches_oracle():
    """IrrLanczos on a DeviceCSR (the cupyx-matrix analogue) against the oracle loop."""
    import lanczos_b200 as lz
    from lanczos_b200 import synth
    from lanczos_b200.engine import DeviceCSR
    cells, seed, n = (8, 7, 6), 21, 30
    L, _, _ = rgg.rgg_laplacian(cells, LAM, seed)
    g = synth.RggGenerator(cells, LAM, seed)
    H = DeviceCSR(*g.rows(0, g.M))
    assert (H.get() != L).nnz == 0
    ref = orc.lanczos(L, n, seed=4)
    S = lz.IrrLanczos(H)
    S.execute_LanczosOld(n, seed=4)
    a, b = np.diag(S.H_eff), np.diag(S.H_eff, 1)
    assert np.max(np.abs(a - ref["alpha"]) / np.abs(ref["alpha"])) < 1e-12
    assert np.max(np.abs(b - ref["beta"]) / np.abs(ref["beta"])) < 1e-12


@gpu
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rgg_blocks_match_oracle(world):
    """Row blocks generated on the device, one per shard, ghost lists from BlockPlan: the team solve
    equals the oracle on the assembled matrix and the single-GPU solve."""
    import lanczos_b200 as lz
    from lanczos_b200 import synth
    from lanczos_b200.team import LocalTeamLanczos
    cells, seed, n = (8, 7, 9), 5, 24
    L, _, _ = rgg.rgg_laplacian(cells, LAM, seed)
    ref = orc.lanczos(L, n, seed=99)
    g = synth.RggGenerator(cells, LAM, seed)
    blocks = [g.row_block(r, world) for r in range(world)]
    T = LocalTeamLanczos(blocks, world)
    T.execute_Lanczos(n, seed=99)
    a, b = np.diag(T.H_eff), np.diag(T.H_eff, 1)
    assert np.max(np.abs(a - ref["alpha"]) / np.abs(ref["alpha"])) < 1e-12
    assert np.max(np.abs(b - ref["beta"]) / np.abs(ref["beta"])) < 1e-12
    th = np.linalg.eigvalsh(T.H_eff)
    assert np.max(np.abs(th - ref["theta"])) / np.abs(ref["theta"]).max() < 1e-10


@gpu
def test_rgg_at_scale_properties():
    """2M-vertex graph (oracle cannot hold it): size-independent properties - symmetric pattern via
    x.Ly == y.Lx, L 1 = 0, x.Lx = sum over edges (x_i - x_j)^2 >= 0, SELL == CSR."""
    import lanczos_b200 as lz
    from lanczos_b200 import synth
    from lanczos_b200.engine import DeviceCSR
    g = synth.RggGenerator((88, 88, 84), LAM, seed=1)
    assert abs(g.M / (88 * 88 * 84) - LAM) < 0.01
    indptr, indices, data = g.rows(0, g.M)
    assert abs(indices.numel() / g.M - 14.0) < 0.6
    ctx = lz.Context.default()
    sell = lz.DeviceOperator.from_device_csr(ctx, indptr, indices, data, fmt="sell")
    csr = lz.DeviceOperator.from_device_csr(ctx, indptr, indices, data, fmt="csr")
    gen = torch.Generator(device=ctx.torch_device).manual_seed(0)
    x = torch.rand(g.M, dtype=torch.float64, device=ctx.torch_device, generator=gen) - 0.5
    y = torch.rand(g.M, dtype=torch.float64, device=ctx.torch_device, generator=gen) - 0.5
    Lx, Ly = sell.apply(x), sell.apply(y)
    assert float((Lx - csr.apply(x)).abs().max()) < 1e-12          # different summation order only
    ones = torch.ones_like(x)
    assert float(sell.apply(ones).abs().max()) == 0.0
    assert abs(float(torch.dot(y, Lx) - torch.dot(x, Ly))) < 1e-9 * float(torch.dot(x, Lx))
    assert float(torch.dot(x, Lx)) > 0.0
    t, s = sell.nnz()
    assert t == indices.numel() and s / t < 1.6


@gpu
def test_full_size_config4_properties():
    """BASELINE config 4 at its full size (~50 M vertices, ~7e8 entries, generated and converted on the
    device): L is symmetric and annihilates constants, the SELL kernel agrees with the CSR kernel, and a
    short Lanczos run satisfies the three-term recurrence."""
    import lanczos_b200 as lz
    from lanczos_b200 import synth
    from lanczos_b200.engine import DeviceCSR
    g = synth.RggGenerator((253, 253, 252), seed=0)
    assert abs(g.M - 50.0e6) < 0.2e6
    indptr, indices, data = g.rows(0, g.M)
    assert abs(indices.numel() / g.M - 14.0) < 0.2
    ctx = lz.Context.default()
    dev = ctx.torch_device
    sell = lz.DeviceOperator.from_device_csr(ctx, indptr, indices, data, fmt="sell")
    gen = torch.Generator(device=dev).manual_seed(3)
    x = torch.rand(g.M, dtype=torch.float64, device=dev, generator=gen) - 0.5
    y = torch.rand(g.M, dtype=torch.float64, device=dev, generator=gen) - 0.5
    Lx, Ly = sell.apply(x), sell.apply(y)
    assert float(sell.apply(torch.ones_like(x)).abs().max()) == 0.0
    xLx = float(torch.dot(x, Lx))
    assert xLx > 0.0 and abs(float(torch.dot(y, Lx) - torch.dot(x, Ly))) < 1e-10 * xLx
    csr = lz.DeviceOperator.from_device_csr(ctx, indptr, indices, data, fmt="csr")
    assert float((Lx - csr.apply(x)).abs().max()) < 1e-12
    t, s = sell.nnz()
    assert t == indices.numel() and s / t < 1.05
    del csr, Lx, Ly, y
    S = lz.IrrLanczos(DeviceCSR(indptr, indices, data))
    n = 6
    S.execute_LanczosOld(n, v0=x, reorth="selective", cgs_passes=2)
    res = S.result
    res.normalize_basis()
    V = res.V_dev[:, :g.M]
    T = torch.from_numpy(S.H_eff).to(dev)
    assert ((V @ V.T) - torch.eye(n, dtype=torch.float64, device=dev)).abs().max().item() < 1e-10
    for j in range(1, n - 1):
        r = sell.apply(V[j].contiguous()) - T[j, j] * V[j] - T[j, j + 1] * V[j + 1] - T[j, j - 1] * V[j - 1]
        assert r.norm().item() < 1e-10
