"""NumPy restatement of the synthetic random-geometric-graph Laplacian of BASELINE config 4.
TEST INFRASTRUCTURE ONLY (see oracle/lanczos_oracle.py for the rules).

The reference (jgslunde/Lanczos) has no generator for this input: its irregular operators come
from IrrGrid/IrrLap point clouds (Python/Irregular/IrrHamiltonian.py:35) of a few thousand
points.  The graph is therefore *defined* in include/lz_synth.h (counter-based hash -> Poisson
point process in unit cells -> radius graph -> L = D - A, rows sorted) and generated on the GPU by
lanczos_b200/csrc/synth_rgg.cu; this file states the same definition with NumPy uint64/fp64
arithmetic so that the sparsity pattern can be compared bit for bit (tests/test_synth_rgg.py).
Brute force O(M^2): small boxes only.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

_U64 = np.uint64
_K1, _K2 = _U64(0xbf58476d1ce4e5b9), _U64(0x94d049bb133111eb)
_GOLD, _KP = _U64(0x9e3779b97f4a7c15), _U64(0xd1342543de82ef95)


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> _U64(30))) * _K1
        z = (z ^ (z >> _U64(27))) * _K2
    return z ^ (z >> _U64(31))


def cell_hash(seed, c):
    with np.errstate(over="ignore"):
        return mix64(_U64(seed) + _GOLD * (np.asarray(c, dtype=np.uint64) + _U64(1)))


def unit53(h):
    return (np.asarray(h, dtype=np.uint64) >> _U64(11)).astype(np.float64) * 2.0 ** -53


def poisson_cdf(lam, n=32):
    """cdf[k] = P(count <= k), k < n, fp64 running sums (the table both sides use)."""
    p = math.exp(-lam)
    out, run = [], 0.0
    for k in range(n):
        run += p
        out.append(run)
        p = p * lam / (k + 1)
    return np.asarray(out, dtype=np.float64)


def cell_counts(cells, lam, seed):
    ncx, ncy, ncz = cells
    c = np.arange(ncx * ncy * ncz, dtype=np.int64)
    u = unit53(cell_hash(seed, c))
    cdf = poisson_cdf(lam)
    return (cdf[None, :31] <= u[:, None]).sum(axis=1).astype(np.int32)


def positions(cells, lam, seed):
    """(prefix[ncells+1], xyz[M,3]) - vertices numbered cell by cell."""
    ncx, ncy, ncz = cells
    cnt = cell_counts(cells, lam, seed)
    prefix = np.concatenate([[0], np.cumsum(cnt, dtype=np.int64)])
    cell = np.repeat(np.arange(len(cnt), dtype=np.int64), cnt)
    k = np.arange(prefix[-1], dtype=np.int64) - prefix[cell]
    hc = cell_hash(seed, cell)
    cc = np.stack([cell % ncx, (cell // ncx) % ncy, cell // (ncx * ncy)], axis=1)
    xyz = np.empty((len(cell), 3))
    for a in range(3):
        with np.errstate(over="ignore"):
            h = mix64(hc ^ (_KP * (4 * k + a + 1).astype(np.uint64)))
        xyz[:, a] = cc[:, a].astype(np.float64) + unit53(h)
    return prefix, xyz


def rgg_laplacian(cells, lam, seed, r2=1.0):
    """L = D - A of the radius graph, sorted CSR with int32 indices."""
    prefix, p = positions(cells, lam, seed)
    M = len(p)
    ex = p[:, None, 0] - p[None, :, 0]
    ey = p[:, None, 1] - p[None, :, 1]
    ez = p[:, None, 2] - p[None, :, 2]
    d2 = (ex * ex + ey * ey) + ez * ez
    adj = d2 <= r2
    np.fill_diagonal(adj, False)
    deg = adj.sum(axis=1).astype(np.float64)
    r, c = np.nonzero(adj)
    d = np.arange(M)
    # every diagonal is stored, also the 0 of an isolated vertex (as the generator does)
    L = sp.csr_matrix((np.concatenate([-np.ones(len(r)), deg]), (np.concatenate([r, d]), np.concatenate([c, d]))),
                      shape=(M, M))
    L.sort_indices()
    L.indices = L.indices.astype(np.int32)
    L.indptr = L.indptr.astype(np.int32)
    return L, prefix, p
