"""Synthetic benchmark inputs generated on the device (liblz_synth.so, include/lz_synth.h).

Not part of the drop-in boundary: bench.py and the tests use it to build the random-geometric-graph
Laplacian of BASELINE config 4 ("irregular 3D random-geometric-graph Laplacian, 50M vertices,
~14 nnz/row"), one row block per GPU, without a host round trip.  The reference has no counterpart
(its irregular operators are built by IrrGrid/IrrLap for a few thousand points).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from .engine import _torch
from .team import RowBlock

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblz_synth.so")

MEAN_DEGREE_LAMBDA = 13.0 / (4.0 * math.pi / 3.0)     # points per unit cell for mean degree 13 at r = 1 cell


class RggParams(C.Structure):
    _fields_ = [("ncx", C.c_int32), ("ncy", C.c_int32), ("ncz", C.c_int32), ("reserved", C.c_int32),
                ("seed", C.c_uint64), ("r2", C.c_double), ("cdf", C.c_double * 32)]


_vp, _i64 = C.c_void_p, C.c_int64
SIGNATURES = {
    "lzs_rgg_cell_counts": [C.POINTER(RggParams), _i64, _i64, _vp, _vp],
    "lzs_rgg_positions": [C.POINTER(RggParams), _vp, _i64, _i64, _vp, _vp],
    "lzs_rgg_row_degrees": [C.POINTER(RggParams), _vp, _i64, _i64, _vp, _vp],
    "lzs_rgg_fill": [C.POINTER(RggParams), _vp, _i64, _i64, _vp, _vp, _vp, _vp],
}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m lanczos_b200.build`")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = C.c_int, args
        _lib = lib
    return _lib


def poisson_cdf(lam: float, n: int = 32):
    """cdf[k] = P(count <= k): fp64 running sums, the table of include/lz_synth.h."""
    p = math.exp(-lam)
    out, run = [], 0.0
    for k in range(n):
        run += p
        out.append(run)
        p = p * lam / (k + 1)
    return out


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"liblz_synth: CUDA error {rc}")


class RggGenerator:
    """Rows of the graph Laplacian of a Poisson random geometric graph in a box of unit cells."""

    def __init__(self, cells, lam: float = MEAN_DEGREE_LAMBDA, seed: int = 0, r2: float = 1.0, device=None):
        torch = _torch()
        self.lib = load()
        self.cells = tuple(int(c) for c in cells)
        self.ncells = self.cells[0] * self.cells[1] * self.cells[2]
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.p = RggParams(self.cells[0], self.cells[1], self.cells[2], 0, int(seed), float(r2),
                           (C.c_double * 32)(*poisson_cdf(lam)))
        with torch.cuda.device(self.device):
            self.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            cnt = torch.empty(self.ncells, dtype=torch.int32, device=self.device)
            _check(self.lib.lzs_rgg_cell_counts(C.byref(self.p), 0, self.ncells, C.c_void_p(cnt.data_ptr()), self.stream))
            self.prefix = torch.zeros(self.ncells + 1, dtype=torch.int64, device=self.device)
            torch.cumsum(cnt, 0, out=self.prefix[1:])
            self.M = int(self.prefix[-1].item())
        if self.M >= 2 ** 31:
            raise ValueError("vertex numbers must fit int32")

    def slab_starts(self, world: int):
        """Row ranges of `world` z-slabs of cells: starts[world + 1]."""
        ncx, ncy, ncz = self.cells
        if world > ncz:
            raise ValueError(f"cannot split {ncz} cell layers over {world} ranks")
        cuts = [(ncz * r // world) * ncx * ncy for r in range(world + 1)]
        torch = _torch()
        idx = torch.tensor(cuts, dtype=torch.int64, device=self.device)
        return [int(v) for v in self.prefix[idx].tolist()]

    def positions(self, row0: int = 0, row1: int | None = None):
        torch = _torch()
        row1 = self.M if row1 is None else row1
        xyz = torch.empty((row1 - row0, 3), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _check(self.lib.lzs_rgg_positions(C.byref(self.p), C.c_void_p(self.prefix.data_ptr()), row0, row1,
                                              C.c_void_p(xyz.data_ptr()), self.stream))
        return xyz

    def rows(self, row0: int, row1: int):
        """(indptr[rows+1] int32, indices int32 GLOBAL columns, data fp64) CUDA tensors of rows [row0, row1)."""
        torch = _torch()
        nrow = row1 - row0
        with torch.cuda.device(self.device):
            pp = C.c_void_p(self.prefix.data_ptr())
            per = torch.empty(max(nrow, 1), dtype=torch.int32, device=self.device)
            _check(self.lib.lzs_rgg_row_degrees(C.byref(self.p), pp, row0, row1, C.c_void_p(per.data_ptr()), self.stream))
            ip64 = torch.zeros(nrow + 1, dtype=torch.int64, device=self.device)
            torch.cumsum(per[:nrow], 0, out=ip64[1:])
            nnz = int(ip64[-1].item())
            if nnz >= 2 ** 31:
                raise ValueError("entries of a row block must fit int32 (scipy CSR layout)")
            indptr = ip64.to(torch.int32)
            del ip64, per
            indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)[:nnz]
            data = torch.empty(max(nnz, 1), dtype=torch.float64, device=self.device)[:nnz]
            _check(self.lib.lzs_rgg_fill(C.byref(self.p), pp, row0, row1, C.c_void_p(indptr.data_ptr()),
                                         C.c_void_p(indices.data_ptr()), C.c_void_p(data.data_ptr()), self.stream))
            torch.cuda.current_stream().synchronize()
        return indptr, indices, data

    def row_block(self, rank: int = 0, world: int = 1) -> RowBlock:
        starts = self.slab_starts(world)
        indptr, indices, data = self.rows(starts[rank], starts[rank + 1])
        return RowBlock(self.M, starts, rank, indptr, indices, data)
