"""Row-sharded (multi-GPU) Lanczos: host-side plan and plumbing.

Vectors and the Krylov basis are split into contiguous row blocks, one per GPU; for structured
grids with the reference's index map i = x + nx*(y + ny*z) (Hamiltonian.py:73-76) contiguous
blocks are z-slabs (y-slabs of a 2-D grid, segments of a 1-D grid).  The exchange itself - halo
planes and the scalar sums - happens inside the CUDA kernels over NVLink peer memory
(csrc/peer.cuh); this module only decides who owns what, allocates the exchange buffers, swaps
their cudaIpc handles through torch.distributed, and calls lz_team_lanczos_run.

    torchrun --nproc-per-node 8 ...:   TeamLanczos(StencilOperator(...))        one process per GPU
    single process, several shards:    LocalTeamLanczos(op, world=4, devices=[0, 0, 0, 0])
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _capi, engine
from ._capi import LanczosBreakdown, RunInfo, RunOpts
from .engine import Context, StencilOperator, padded_ld


# --------------------------------------------------------------------------- the plan (pure host)
@dataclass
class SlabPlan:
    """Who owns which slab.  The sharded axis is the slowest grid axis; it is presented to the
    kernels as z (a 2-D grid (nx, ny) becomes (nx, 1, ny), a 1-D grid (nx,) becomes (1, 1, nx))."""
    grid: tuple
    world: int
    periodic: bool

    def __post_init__(self):
        g = tuple(int(x) for x in self.grid)
        if len(g) == 1:
            self.grid3 = (1, 1, g[0])
        elif len(g) == 2:
            self.grid3 = (g[0], 1, g[1])
        elif len(g) == 3:
            self.grid3 = g
        else:
            raise ValueError("1, 2 or 3 grid dimensions")
        nz = self.grid3[2]
        if self.world < 1 or self.world > nz:
            raise ValueError(f"cannot split {nz} slabs over {self.world} ranks")
        base, extra = divmod(nz, self.world)
        self.counts = [base + (1 if r < extra else 0) for r in range(self.world)]
        self.starts = [sum(self.counts[:r]) for r in range(self.world)]
        self.plane = self.grid3[0] * self.grid3[1]
        self.M = self.plane * nz

    def off3(self, off: Sequence[float]):
        o = tuple(float(x) for x in off)
        if len(o) == 1:
            return (0.0, 0.0, o[0])
        if len(o) == 2:
            return (o[0], 0.0, o[1])
        return o

    def slab(self, rank: int):
        """(z0, z1) planes owned by `rank`."""
        return self.starts[rank], self.starts[rank] + self.counts[rank]

    def rows(self, rank: int):
        """[row0, row1) of the global vector owned by `rank`."""
        z0, z1 = self.slab(rank)
        return z0 * self.plane, z1 * self.plane

    def local_rows(self, rank: int) -> int:
        return self.counts[rank] * self.plane

    def neighbours(self, rank: int):
        """(lower, upper) ranks owning the plane below / above the slab; -1 at a Dirichlet wall.
        With a single rank and periodic boundaries the slab is its own neighbour."""
        lo, up = rank - 1, rank + 1
        if lo < 0:
            lo = self.world - 1 if self.periodic else -1
        if up >= self.world:
            up = 0 if self.periodic else -1
        return lo, up

    def halo_planes(self, rank: int):
        """Global z indices of the two ghost planes of `rank` (None at a wall) - the contract the
        kernels implement, used by the CPU tests."""
        z0, z1 = self.slab(rank)
        nz = self.grid3[2]
        lo = z0 - 1 if z0 > 0 else (nz - 1 if self.periodic else None)
        up = z1 if z1 < nz else (0 if self.periodic else None)
        return lo, up


class RowBlockPlan:
    """Contiguous row blocks of a sparse operator: rank r owns rows [r0, r1); the
    columns of its block are renumbered - owned columns to [0, M_loc), the others to
    M_loc + (position in the rank's sorted list of ghost columns).  The ghost list of a rank is,
    per owner, a contiguous run (owners are contiguous row ranges), so every peer's contribution
    lands as one segment of the rank's gather buffer."""

    def __init__(self, H, world: int, align: int = 32):
        import scipy.sparse as sp
        A = sp.csr_matrix(H)
        if A.shape[0] != A.shape[1]:
            raise ValueError("operator must be square")
        A.sort_indices()
        self.A = A
        self.world = int(world)
        self.M = A.shape[0]
        self.plane = 0
        per = -(-self.M // self.world)
        per = -(-per // align) * align
        self.starts = [min(r * per, self.M) for r in range(self.world + 1)]
        if self.starts[-2] >= self.M:
            raise ValueError(f"cannot split {self.M} rows over {self.world} ranks in blocks of {align}")
        self.ghost_cols = []          # per rank: sorted global columns it needs from others
        for r in range(self.world):
            r0, r1 = self.rows(r)
            cols = A.indices[A.indptr[r0]:A.indptr[r1]]
            self.ghost_cols.append(np.unique(cols[(cols < r0) | (cols >= r1)]).astype(np.int64))
        self.nghost_max = max(len(g) for g in self.ghost_cols)

    def rows(self, rank: int):
        return self.starts[rank], self.starts[rank + 1]

    def local_rows(self, rank: int) -> int:
        return self.starts[rank + 1] - self.starts[rank]

    def owner_segments(self, rank: int):
        """[(owner q, start, end)] : ghost_cols[rank][start:end] are owned by q."""
        g = self.ghost_cols[rank]
        cuts = np.searchsorted(g, self.starts)
        return [(q, int(cuts[q]), int(cuts[q + 1])) for q in range(self.world)]

    def local_csr(self, rank: int):
        """(indptr, indices, data, ncols) of the rank's row block with renumbered columns."""
        r0, r1 = self.rows(rank)
        A = self.A
        lo, hi = A.indptr[r0], A.indptr[r1]
        cols = A.indices[lo:hi].astype(np.int64)
        own = (cols >= r0) & (cols < r1)
        local = np.empty_like(cols)
        local[own] = cols[own] - r0
        local[~own] = (r1 - r0) + np.searchsorted(self.ghost_cols[rank], cols[~own])
        indptr = (A.indptr[r0:r1 + 1] - lo).astype(np.int32)
        return indptr, local.astype(np.int32), np.ascontiguousarray(A.data[lo:hi], dtype=np.float64), \
            (r1 - r0) + len(self.ghost_cols[rank])

    def make_op(self, lib, ctx: Context, rank: int, fmt: int, sigma: int):
        indptr, indices, data, ncols = self.local_csr(rank)
        h = C.c_void_p()
        _capi.check(lib.lz_op_csr_shard_create(
            ctx.handle, self.local_rows(rank), ncols, len(data), indptr.ctypes.data_as(C.c_void_p),
            indices.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), fmt, sigma, C.byref(h)))
        return h

    def send_lists(self, rank: int):
        """What `rank` sends: (send_idx, seg_start[world+1], dst_off[world]).  For destination q the
        entries are q's ghost columns owned by `rank` (as local row indices), and they land at the
        position of that run inside q's ghost list."""
        r0, _ = self.rows(rank)
        idx, seg, off = [], [0], []
        for q in range(self.world):
            if q == rank:
                seg.append(seg[-1])
                off.append(0)
                continue
            _, a, b = self.owner_segments(q)[rank]
            idx.append(self.ghost_cols[q][a:b] - r0)
            seg.append(seg[-1] + (b - a))
            off.append(a)
        send = np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, dtype=np.int32)
        return send, np.asarray(seg, dtype=np.int32), np.asarray(off, dtype=np.int64)


@dataclass(eq=False)
class RowBlock:
    """Rows [starts[rank], starts[rank + 1]) of a square sparse operator with M rows, as CSR with
    GLOBAL column indices.  The arrays are NumPy arrays, CPU tensors or CUDA tensors (int32
    indptr / indices, fp64 data); a rank only ever holds its own block, so operators that do not
    fit one host or one GPU (BASELINE config 4: 5e7 rows, 7e8 entries) can still be sharded."""
    M: int
    starts: Sequence[int]
    rank: int
    indptr: object
    indices: object
    data: object

    @property
    def shape(self):
        return (self.M, self.M)


def _as_tensor(a, dtype):
    import torch
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(dtype).contiguous()


class BlockPlan:
    """RowBlockPlan for operators given block by block: the same numbering contract (owned columns
    first, then the rank's sorted ghost list; every owner's run of a ghost list is contiguous), but
    each rank derives its part from its own block and only the ghost lists travel (`gather`: a
    function that turns this process's {rank: ghost list} into the lists of all ranks -
    torch.distributed.all_gather_object under torchrun, identity when one process holds every block).
    The index work runs in torch on whatever device holds the block (plumbing, not arithmetic)."""

    def __init__(self, blocks: Sequence[RowBlock], world: int, gather=None):
        import torch
        self.world = int(world)
        b0 = blocks[0]
        self.M = int(b0.M)
        self.starts = [int(x) for x in b0.starts]
        self.plane = 0
        if len(self.starts) != self.world + 1 or self.starts[0] != 0 or self.starts[-1] != self.M:
            raise ValueError("RowBlock.starts must be world + 1 offsets from 0 to M")
        if any(self.starts[r + 1] <= self.starts[r] for r in range(self.world)):
            raise ValueError("every rank needs at least one row")
        self.blocks, self._local, mine = {}, {}, {}
        for b in blocks:
            if [int(x) for x in b.starts] != self.starts or int(b.M) != self.M:
                raise ValueError("row blocks disagree about the partition")
            r0, r1 = self.rows(b.rank)
            indptr = _as_tensor(b.indptr, torch.int32)
            indices = _as_tensor(b.indices, torch.int32)
            data = _as_tensor(b.data, torch.float64)
            if indptr.numel() != r1 - r0 + 1:
                raise ValueError(f"block of rank {b.rank}: indptr must have {r1 - r0 + 1} entries")
            if self.world == 1:
                ghosts = torch.zeros(0, dtype=torch.int64, device=indices.device)
                local = indices
            else:
                own = (indices >= r0) & (indices < r1)
                ghosts = torch.unique(indices[~own]).to(torch.int64)          # sorted
                local = torch.where(own, indices - r0,
                                    (r1 - r0) + torch.searchsorted(ghosts, indices.to(torch.int64)).to(torch.int32))
                del own
            self.blocks[b.rank] = b
            self._local[b.rank] = (indptr, local.to(torch.int32).contiguous(), data, (r1 - r0) + int(ghosts.numel()))
            mine[b.rank] = ghosts.cpu().numpy()
        every = mine if gather is None else gather(mine)
        if sorted(every) != list(range(self.world)):
            raise ValueError("ghost lists of some ranks are missing")
        self.ghost_cols = [np.asarray(every[r], dtype=np.int64) for r in range(self.world)]
        self.nghost_max = max(len(g) for g in self.ghost_cols)

    def rows(self, rank: int):
        return self.starts[rank], self.starts[rank + 1]

    def local_rows(self, rank: int) -> int:
        return self.starts[rank + 1] - self.starts[rank]

    owner_segments = RowBlockPlan.owner_segments
    send_lists = RowBlockPlan.send_lists

    def local_csr(self, rank: int):
        """(indptr, indices, data, ncols) with renumbered columns, on the device that holds the block."""
        return self._local[rank]

    def make_op(self, lib, ctx: Context, rank: int, fmt: int, sigma: int):
        indptr, indices, data, ncols = self._local[rank]
        h = C.c_void_p()
        nnz = int(indices.numel())
        if indptr.is_cuda:
            if indptr.device.index != ctx.device:
                raise ValueError(f"the block of rank {rank} lives on another device than its shard")
            import torch
            torch.cuda.current_stream(ctx.device).synchronize()
            _capi.check(lib.lz_op_csr_create_dev(
                ctx.handle, self.local_rows(rank), ncols, nnz, C.c_void_p(indptr.data_ptr()),
                C.c_void_p(indices.data_ptr() if nnz else 0), C.c_void_p(data.data_ptr() if nnz else 0),
                fmt, sigma, C.byref(h)))
        else:
            _capi.check(lib.lz_op_csr_shard_create(
                ctx.handle, self.local_rows(rank), ncols, nnz, C.c_void_p(indptr.data_ptr()),
                C.c_void_p(indices.data_ptr() if nnz else 0), C.c_void_p(data.data_ptr() if nnz else 0),
                fmt, sigma, C.byref(h)))
        return h


# --------------------------------------------------------------------------- device side
class _Shard:
    def __init__(self, ctx: Context, rank: int):
        self.ctx, self.rank = ctx, rank
        self.comm_ptr = None
        self.opened = {}
        self.op_handle = None
        self.keep = ()


class _TeamBase:
    """Shared driver: subclasses provide the shards and the mapped exchange buffers."""

    def __init__(self, H, world: int, fmt: str = "sell", sigma: int = 0):
        import scipy.sparse as sp
        self.H = H
        self.world = int(world)
        self.fmt, self.sigma = fmt, int(sigma)
        if isinstance(H, StencilOperator):
            self.plan = SlabPlan(H.grid, self.world, H.bc == "periodic")
            self.sparse = False
        elif sp.issparse(H):
            self.plan = RowBlockPlan(H, self.world)
            self.sparse = True
        elif isinstance(H, RowBlock) or (isinstance(H, (list, tuple)) and H and all(isinstance(b, RowBlock) for b in H)):
            blocks = [H] if isinstance(H, RowBlock) else list(H)
            self.plan = BlockPlan(blocks, self.world, gather=self._gather_ghost_lists)
            self.sparse = True
        else:
            raise TypeError("row-sharded runs take a StencilOperator, a scipy.sparse matrix or RowBlock(s)")
        self.M = self.plan.M
        self.lib = _capi.load()
        self.team = None
        self.max_steps = 0
        self.Lanczos_has_been_executed = False
        self._results = None

    def _gather_ghost_lists(self, mine: dict) -> dict:
        """{rank: ghost list} of every rank from those of this process (one process: identity)."""
        return mine

    # subclasses: self.shards (list of _Shard), self._map_buffers(nbytes) -> per-shard list of `world` pointers
    def _ensure_team(self, n: int):
        if self.team is not None and n <= self.max_steps:
            return
        self._destroy_team()
        max_steps = max(int(n), 128)
        nbytes = C.c_int64()
        # a single shard wraps periodic boundaries inside its own vector: no ghost planes
        plane = self.plan.plane if self.world > 1 else 0
        nghost = self.plan.nghost_max if self.sparse else 0
        _capi.check(self.lib.lz_comm_bytes(self.world, max_steps, plane, nghost, C.byref(nbytes)))
        tables = self._map_buffers(nbytes.value)
        nl = len(self.shards)
        ranks = (C.c_int * nl)(*[s.rank for s in self.shards])
        ctxs = (C.c_void_p * nl)(*[s.ctx.handle for s in self.shards])
        h = C.c_void_p()
        _capi.check(self.lib.lz_team_create(self.world, nl, ranks, ctxs, self.M, max_steps, plane, nghost, C.byref(h)))
        self.team = h
        self.max_steps = max_steps
        for i, s in enumerate(self.shards):
            ptrs = (C.c_void_p * self.world)(*tables[i])
            lo, up = (-1, -1) if (self.sparse or self.world == 1) else self.plan.neighbours(s.rank)
            _capi.check(self.lib.lz_team_attach(self.team, i, ptrs, lo, up))
            if self.sparse:
                send, seg, off = self.plan.send_lists(s.rank)
                _capi.check(self.lib.lz_team_set_ghosts(
                    self.team, i, len(send), send.ctypes.data_as(C.c_void_p),
                    seg.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p)))
        self._make_ops()

    def _make_ops(self):
        torch = engine._torch()
        H, plan = self.H, self.plan
        if self.sparse:
            f = dict(engine.FORMATS, auto=_capi.LZ_FMT_SELL)[self.fmt]
            for s in self.shards:
                s.op_handle = plan.make_op(self.lib, s.ctx, s.rank, f, self.sigma)
            return
        off3 = plan.off3(H.off)
        for s in self.shards:
            z0, z1 = plan.slab(s.rank)
            shape = (C.c_int64 * 3)(plan.grid3[0], plan.grid3[1], z1 - z0)
            off = (C.c_double * 3)(*off3)
            diag_t, diag_p = None, C.c_void_p(0)
            if H.diag is not None:
                r0, r1 = plan.rows(s.rank)
                if isinstance(H.diag, torch.Tensor):
                    diag_t = H.diag.reshape(-1)[r0:r1].to(device=s.ctx.torch_device, dtype=torch.float64).contiguous()
                else:
                    diag_t = torch.from_numpy(np.ascontiguousarray(np.asarray(H.diag, dtype=np.float64).reshape(-1)[r0:r1])).to(s.ctx.torch_device)
                diag_p = C.c_void_p(diag_t.data_ptr())
            h = C.c_void_p()
            bc = _capi.LZ_BC_PERIODIC if H.bc == "periodic" else _capi.LZ_BC_DIRICHLET
            if H.weights27 is not None:
                w = (C.c_double * 4)(*H.weights27)
                _capi.check(self.lib.lz_op_stencil27_create(s.ctx.handle, shape, bc, w, diag_p, C.byref(h)))
            else:
                _capi.check(self.lib.lz_op_stencil_create(s.ctx.handle, 3, shape, bc, float(H.center), off, diag_p, C.byref(h)))
            s.op_handle = h
            s.keep = (diag_t,)

    def _destroy_team(self):
        if getattr(self, "team", None):
            for s in self.shards:
                if s.op_handle:
                    self.lib.lz_op_destroy(s.op_handle)
                    s.op_handle = None
            self.lib.lz_team_destroy(self.team)
            self.team = None
            self._unmap_buffers()

    def __del__(self):
        try:
            self._destroy_team()
        except Exception:
            pass

    # ---- the loop ---------------------------------------------------------------------------------
    def _local_start_vectors(self, seed, v0):
        """Per local shard: a CUDA tensor with the shard's rows of the start vector.  `v0` may be
        None (the reference's seeded global vector, Lanczos.py:93-97, sliced), a global host array,
        or a list of per-shard CUDA tensors / one CUDA tensor (single local shard)."""
        torch = engine._torch()
        out = []
        if v0 is None or isinstance(v0, np.ndarray) or (isinstance(v0, (list, tuple)) and np.isscalar(v0[0])):
            full = engine.start_vector(self.M, seed, v0)
            for s in self.shards:
                r0, r1 = self.plan.rows(s.rank)
                out.append(torch.from_numpy(np.ascontiguousarray(full[r0:r1])).to(s.ctx.torch_device))
            return out
        np.random.seed(seed)
        vs = list(v0) if isinstance(v0, (list, tuple)) else [v0]
        if len(vs) != len(self.shards):
            raise ValueError("one start-vector shard per local shard")
        for s, v in zip(self.shards, vs):
            t = v.to(device=s.ctx.torch_device, dtype=torch.float64).contiguous().reshape(-1)
            if t.numel() != self.plan.local_rows(s.rank):
                raise ValueError(f"shard of rank {s.rank} must have {self.plan.local_rows(s.rank)} rows")
            out.append(t)
        return out

    def execute_Lanczos(self, n, seed=99, use_cuda=True, v0=None, *, reorth="full", cgs_passes=1,
                        ref_compat=True, keep_basis=True, breakdown_tol=0.0, select_tol=0.0,
                        profile=False, step_kernel="auto", cgs_fused=True, sweep_form=0, kb_alpha=False, overlap=True, **_ignored):
        torch = engine._torch()
        n = int(n)
        if n > self.M:
            raise ValueError("n cannot be larger than M!")
        if ref_compat and n < 2:
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")
        self._ensure_team(n)
        self._results = None
        self.Lanczos_has_been_executed = False
        mode = engine._REORTH[reorth]
        need_basis = keep_basis or mode != _capi.LZ_REORTH_NONE
        starts = self._local_start_vectors(seed, v0)
        nl = len(self.shards)
        Vs, lds = [], []
        for s in self.shards:
            ld = padded_ld(self.plan.local_rows(s.rank))
            lds.append(ld)
            with torch.cuda.device(s.ctx.device):
                Vs.append(torch.empty((n, ld), dtype=torch.float64, device=s.ctx.torch_device) if need_basis else None)
        alpha, beta, scale = np.zeros(n), np.zeros(max(n - 1, 0)), np.ones(n)
        opts = RunOpts(mode, int(cgs_passes), 1 if ref_compat else 0, 1 if profile else 0,
                       engine.STEP_KERNEL[step_kernel], engine.run_flags(cgs_fused, sweep_form, kb_alpha, overlap), float(breakdown_tol), float(select_tol))
        info = RunInfo()
        ops = (C.c_void_p * nl)(*[s.op_handle for s in self.shards])
        v0p = (C.c_void_p * nl)(*[t.data_ptr() for t in starts])
        Vp = (C.c_void_p * nl)(*[(V.data_ptr() if V is not None else 0) for V in Vs])
        ldp = (C.c_int64 * nl)(*lds)
        for s in self.shards:          # inputs were produced on torch's streams
            torch.cuda.synchronize(s.ctx.device)
        self._pre_run_barrier()        # ranks enter the loop together (the device-side waits are bounded)
        status = self.lib.lz_team_lanczos_run(
            self.team, ops, v0p, n, C.byref(opts), alpha.ctypes.data_as(C.c_void_p),
            beta.ctypes.data_as(C.c_void_p), Vp, ldp, scale.ctypes.data_as(C.c_void_p), C.byref(info))
        if status == _capi.LZ_ERR_BREAKDOWN:
            raise LanczosBreakdown(self.lib.lz_last_error().decode(), steps_done=info.steps_done)
        if status in (_capi.LZ_ERR_CUDA, _capi.LZ_ERR_PEER):
            # the ranks may be out of step for good (the library refuses further runs on this team):
            # every rank sees an error - its own or a peer timeout - and builds a fresh team on the next call
            msg = self.lib.lz_last_error().decode()
            try:
                self._destroy_team()
            except Exception:
                pass
            raise RuntimeError(f"lanczos_b200 error {status}: {msg} (the team was discarded; the next run creates a new one)")
        _capi.check(status)
        self.n = n
        self._results = [engine.LanczosResult(s.ctx, n, self.plan.local_rows(s.rank), alpha, beta, V, ld,
                                              scale.copy(), info)
                         for s, V, ld in zip(self.shards, Vs, lds)]
        self._H_eff = self._results[0].tridiagonal()
        self.Lanczos_has_been_executed = True

    execute_LanczosOld = execute_Lanczos

    def _pre_run_barrier(self):
        pass

    @property
    def result(self):
        if not self.Lanczos_has_been_executed:
            raise ValueError("Lanczos Algorithm has not been called.")
        return self._results[0]

    @property
    def results(self):
        if not self.Lanczos_has_been_executed:
            raise ValueError("Lanczos Algorithm has not been called.")
        return self._results

    @property
    def H_eff(self):
        if not self.Lanczos_has_been_executed:
            raise ValueError("Lanczos Algorithm has not been called.")
        return self._H_eff

    def ritz_values(self, k=None):
        theta = np.linalg.eigvalsh(self.H_eff)
        return theta if k is None else theta[:k]

    @property
    def M_local(self):
        return self.plan.local_rows(self.shards[0].rank)

    # ---- what the drop-in classes need after the loop (Lanczos.py:132-185), shard by shard ------------------
    def ritz_vectors_dev(self, S):
        """Per local shard: Y (k, ld) CUDA tensor, Y[c] = this shard's rows of V @ S[:, c] (K5 on every shard;
        the lift needs no exchange: each rank lifts its own rows, Lanczos.py:154-156)."""
        return [r.ritz_vectors_dev(S) for r in self.results]

    def _gather_blocks(self, blocks):
        """Per-local-shard host arrays (k, M_local) -> the global (k, M) array on every process."""
        return np.concatenate(blocks, axis=1)

    def rows_host(self, tensors):
        """(k, M) host array from per-shard CUDA tensors (k, ld >= M_local)."""
        blocks = [t[:, :self.plan.local_rows(s.rank)].cpu().numpy() for s, t in zip(self.shards, tensors)]
        return self._gather_blocks(blocks)

    def basis_rows_host(self):
        """(n, M) host array of the normalised basis, assembled from the shards."""
        return self._gather_blocks([r.basis_rows_host() for r in self.results])

    def apply_dots(self, xs, ys=None):
        """y = H x over the shards (xs / ys: one CUDA tensor of M_local doubles per local shard; ys allocated
        when None).  Returns (x.Hx, Hx.Hx, ys), the sums taken over all ranks in rank order."""
        torch = engine._torch()
        self._ensure_team(max(self.max_steps, 2))
        nl = len(self.shards)
        if ys is None:
            ys = [torch.empty_like(x) for x in xs]
        for s in self.shards:
            torch.cuda.synchronize(s.ctx.device)
        dots = np.zeros(2)
        ops = (C.c_void_p * nl)(*[s.op_handle for s in self.shards])
        xp = (C.c_void_p * nl)(*[x.data_ptr() for x in xs])
        yp = (C.c_void_p * nl)(*[y.data_ptr() for y in ys])
        self._pre_run_barrier()
        _capi.check(self.lib.lz_team_apply_dots(self.team, ops, xp, yp, dots.ctypes.data_as(C.c_void_p)))
        return float(dots[0]), float(dots[1]), ys

    def residual_cosines(self, Ys):
        """cos^2 of the angle between H y and y for every row of the sharded Ritz-vector blocks `Ys`
        (Lanczos.py:171-176) - the operator kernel with its halo / ghost exchange, sums over the peer ring."""
        k = Ys[0].shape[0]
        out = np.zeros(k)
        ys = None
        for i in range(k):
            xs = [Y[i, :self.plan.local_rows(s.rank)] for s, Y in zip(self.shards, Ys)]
            xhx, hxhx, ys = self.apply_dots(xs, ys)
            out[i] = xhx ** 2 / hxhx if hxhx > 0.0 else 0.0
        return out

    def value_free_local(self) -> bool:
        v = C.c_int32()
        _capi.check(self.lib.lz_op_value_free(self.shards[0].op_handle, C.byref(v)))
        return bool(v.value)

    def windowed_local(self) -> int:
        """Largest per-window stage (granules of 32 entries) of the first local shard, 0 = plain SELL kernel."""
        v = C.c_int32()
        _capi.check(self.lib.lz_op_windowed(self.shards[0].op_handle, C.byref(v)))
        return int(v.value)

    def nnz_local(self):
        """(true, stored) entries of the first local shard of a sparse operator."""
        t, s = C.c_int64(), C.c_int64()
        _capi.check(self.lib.lz_op_nnz(self.shards[0].op_handle, C.byref(t), C.byref(s)))
        return t.value, s.value


class LocalTeamLanczos(_TeamBase):
    """All `world` shards driven by this process: on one GPU (tests of the exchange logic - the
    kernels of different shards then run one after the other, push phase before combine phase) or
    on several GPUs (`devices`)."""

    def __init__(self, H, world: int, devices: Optional[List[int]] = None, **kw):
        super().__init__(H, world, **kw)
        torch = engine._torch()
        if devices is None:
            devices = [torch.cuda.current_device()] * self.world
        if len(devices) != self.world:
            raise ValueError("one device per shard")
        # one context (own partials / workspace arena) per shard, even on a shared device
        self.shards = [_Shard(Context(d), r) for r, d in enumerate(devices)]

    def _map_buffers(self, nbytes):
        ptrs = []
        for s in self.shards:
            p = C.c_void_p()
            _capi.check(self.lib.lz_comm_alloc(s.ctx.handle, nbytes, C.byref(p), None))
            s.comm_ptr = p
            ptrs.append(p.value)
        return [list(ptrs) for _ in self.shards]

    def _unmap_buffers(self):
        for s in self.shards:
            if s.comm_ptr:
                self.lib.lz_comm_free(s.ctx.handle, s.comm_ptr)
                s.comm_ptr = None

class TeamLanczos(_TeamBase):
    """One process per GPU (torchrun): this process drives the shard of its rank; exchange
    buffers of the other ranks are mapped through cudaIpc handles swapped over torch.distributed."""

    def __init__(self, H, rank: Optional[int] = None, world: Optional[int] = None, device=None, **kw):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("TeamLanczos needs an initialised torch.distributed process group")
        self.dist = dist
        rank = dist.get_rank() if rank is None else rank
        world = dist.get_world_size() if world is None else world
        super().__init__(H, world, **kw)
        self.rank = rank
        self.shards = [_Shard(Context.default(device), rank)]

    def _map_buffers(self, nbytes):
        s = self.shards[0]
        p = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _capi.check(self.lib.lz_comm_alloc(s.ctx.handle, nbytes, C.byref(p), handle))
        s.comm_ptr = p
        handles = [None] * self.world
        self.dist.all_gather_object(handles, bytes(handle))
        ptrs = []
        for q in range(self.world):
            if q == self.rank:
                ptrs.append(p.value)
                continue
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[q])
            mapped = C.c_void_p()
            _capi.check(self.lib.lz_comm_open(s.ctx.handle, buf, C.byref(mapped)))
            s.opened[q] = mapped
            ptrs.append(mapped.value)
        self.dist.barrier()
        return [ptrs]

    def _pre_run_barrier(self):
        self.dist.barrier()

    def _gather_blocks(self, blocks):
        """The local (k, M_local) block of every rank -> the global (k, M) array on every rank (host side,
        through the process group: this is output plumbing, used when the caller asks for host arrays)."""
        parts = [None] * self.world
        self.dist.all_gather_object(parts, blocks[0])
        return np.concatenate(parts, axis=1)

    def _gather_ghost_lists(self, mine: dict) -> dict:
        parts = [None] * self.dist.get_world_size()
        self.dist.all_gather_object(parts, mine)
        out = {}
        for p in parts:
            out.update(p)
        return out

    def _unmap_buffers(self):
        s = self.shards[0]
        for q, mapped in list(s.opened.items()):
            self.lib.lz_comm_close(s.ctx.handle, mapped)
        s.opened = {}
        try:
            self.dist.barrier()        # nobody frees while a peer still has the buffer mapped
        except Exception:
            pass
        if s.comm_ptr:
            self.lib.lz_comm_free(s.ctx.handle, s.comm_ptr)
            s.comm_ptr = None
