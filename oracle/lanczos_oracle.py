"""CPU oracle for the Lanczos tridiagonalization hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy *restatement* of the algorithm that the reference
(jgslunde/Lanczos) runs on its CPU path.  It exists so that the CUDA path in
``lanczos_b200`` can be checked on machines where ``/root/reference`` is absent
(the GPU box).  Nothing in the product (``lanczos_b200/``, ``Python/``) imports it:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may.

Parity pinning: the reference ships no golden vectors or known-answer tests for this
path.  The oracle is therefore pinned against the *live*
reference, executed in the authoring container by ``tests/golden/make_golden.py``;
the resulting alpha/beta/Ritz values and CSR patterns are committed under
``tests/golden/`` and ``tests/test_oracle.py`` holds the oracle to them bit-for-bit
(same NumPy ops in the same order => identical floats).

Reference lines restated by each function are cited as ``Lanczos.py:a-b`` (=
``/root/reference/Python/Regular/Lanczos.py``), ``IrrLanczos.py`` (=
``Python/Irregular/IrrLanczos.py``) and ``Hamiltonian.py`` (=
``Python/Regular/Hamiltonian.py``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "start_vector", "gram_schmidt_row", "tridiagonalize", "assemble_tridiagonal",
    "ritz_pairs", "lanczos", "laplacian_csr", "laplacian27_csr", "box27_weights", "reference_T_csr", "deuteron_potential",
    "deuteron_hamiltonian", "delaunay_graph_laplacian", "rgg_graph_laplacian",
    "csr_matvec_rows", "timed_steps",
]


# --------------------------------------------------------------------------- loop

def start_vector(M, seed=99, v0=None):
    """Start vector of the reference: legacy global MT19937 stream, uniform(-1,1),
    then 2-norm scaling.  Lanczos.py:93-100 / IrrLanczos.py:210-217."""
    np.random.seed(seed)
    x = np.random.uniform(-1, 1, size=(M)) if v0 is None else np.array(v0)
    return x / np.linalg.norm(x)


def gram_schmidt_row(V, j):
    """One classical Gram-Schmidt sweep of row j against *every* row of V (zero rows
    and row j itself included), CPU form of the reference:
        ip = sum(V[j]*V, axis=1);  V[j] = 2 V[j] - sum(ip[:,None]*V, axis=0)
    Lanczos.py:247-249, IrrLanczos.py:463-464 (the Irregular class uses this form on
    both branches, :454-455)."""
    ip = np.sum(V[j] * V, axis=1)
    V[j] = 2 * V[j] - np.sum(ip[:, None] * V, axis=0)


def gram_schmidt_row_blocked(V, j, nrows=None, threads=None):
    """gram_schmidt_row for bases of GB size (full-size config-2 parity test): the same arithmetic,
    bit for bit, without the two (n, M) temporaries of the reference's expressions.
      * ip[r] = np.sum(V[j]*V[r]) row by row - the reduction NumPy runs per row of `np.sum(V[j]*V, axis=1)`
        (pairwise summation along the contiguous axis);
      * the axis-0 sum `np.sum(ip[:,None]*V, axis=0)` adds the rows in order, element by element:
        acc = ip[0]*V[0]; acc += ip[1]*V[1]; ...   (done per column slab, slabs in parallel);
      * rows >= nrows are known to be zero (rows the loop has not reached): they add +0.0 to every sum.
    tests/test_oracle.py holds it to gram_schmidt_row bit-for-bit."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    n, M = V.shape
    nrows = n if nrows is None else nrows
    threads = threads or min(16, os.cpu_count() or 1)
    vj = V[j].copy()
    with ThreadPoolExecutor(threads) as ex:
        ip = np.array(list(ex.map(lambda r: np.sum(vj * V[r]), range(nrows))))
        slab = -(-M // (threads * 4))

        def upd(c0):
            c1 = min(M, c0 + slab)
            acc = ip[0] * V[0, c0:c1]
            for r in range(1, nrows):
                acc += ip[r] * V[r, c0:c1]
            return c0, c1, acc
        out = np.empty(M)
        for c0, c1, acc in ex.map(upd, range(0, M, slab)):
            out[c0:c1] = acc
    V[j] = 2 * vj - out


def gram_schmidt_row_gpu_form(V, j):
    """The sweep as Regular/Lanczos.py:236-238 states it for use_cuda=True (CuPy arrays there;
    the same operations on NumPy arrays here): the self inner product is zeroed,
        ip = sum(V[j]*V, axis=1); ip[j] = 0;  V[j] = V[j] - sum(ip[:,None]*V, axis=0)"""
    ip = np.sum(V[j] * V, axis=1)
    ip[j] = 0
    V[j] = V[j] - np.sum(ip[:, None] * V, axis=0)


def tridiagonalize(H, n, seed=99, v0=None, reorth=True, sweep="cpu", blocked=False):
    """The n-step symmetric Lanczos loop exactly as the reference runs it
    (Lanczos.py:104-119 == IrrLanczos.py:221-238), quirks included:
      * the user's start vector only seeds the pre-step; row 0 of the basis is
        normalize(H v0 - (v0.H v0) v0);
      * beta[j-1] at j=0 lands in beta[-1] and is overwritten by the last step;
      * V[j-1] at j=0 is the (still zero) last row;
      * n == 1 raises IndexError (beta is empty), n > M raises ValueError.
    `sweep`: "cpu" = the 2 V[j] - sum form (Lanczos.py:247-249, IrrLanczos.py both branches),
    "gpu" = the form Regular/Lanczos.py:236-238 states for use_cuda=True (self term dropped).
    Returns (alpha (n,), beta (n-1,), V (n, M) row-major)."""
    M = np.shape(H)[0]
    if n > M:
        raise ValueError("n cannot be larger than M!")       # Lanczos.py:76-77
    q = start_vector(M, seed, v0)
    V = np.zeros((n, M))
    V[0] = q
    alpha = np.zeros(n)
    beta = np.zeros(n - 1)
    r = H * V[0]
    alpha[0] = np.dot(r, V[0])
    r = r - alpha[0] * V[0]
    for j in range(n):
        beta[j - 1] = np.linalg.norm(r)
        V[j] = r / beta[j - 1]
        if reorth and blocked and sweep == "cpu":
            gram_schmidt_row_blocked(V, j, nrows=j + 1)      # same bits, no (n, M) temporaries
        elif reorth:
            (gram_schmidt_row_gpu_form if sweep == "gpu" else gram_schmidt_row)(V, j)
        r = H * V[j]
        alpha[j] = np.dot(V[j], r)
        r = r - V[j] * alpha[j] - V[j - 1] * beta[j - 1]
    return alpha, beta, V


def assemble_tridiagonal(alpha, beta):
    """Dense n x n symmetric tridiagonal H_eff.  Lanczos.py:121-130."""
    n = len(alpha)
    T = np.zeros((n, n))
    idx = np.arange(n)
    T[idx, idx] = alpha
    if n > 1:
        T[idx[:-1], idx[:-1] + 1] = beta
        T[idx[:-1] + 1, idx[:-1]] = beta
    return T


def ritz_pairs(T, V_cols):
    """eigh of the tridiagonal and lift of every eigenvector with the basis
    (V_cols is (M, n), columns = Lanczos vectors).  Lanczos.py:151-156."""
    theta, S = np.linalg.eigh(T)
    n = T.shape[0]
    Y = np.zeros((V_cols.shape[0], n))
    for i in range(n):
        Y[:, i] = np.dot(V_cols, S[:, i])
    return theta, S, Y


def lanczos(H, n, seed=99, v0=None, reorth=True, vectors=False, sweep="cpu", blocked=False):
    """Convenience wrapper: returns dict(alpha, beta, T, V (M,n), theta[, Y])."""
    alpha, beta, V = tridiagonalize(H, n, seed=seed, v0=v0, reorth=reorth, sweep=sweep, blocked=blocked)
    T = assemble_tridiagonal(alpha, beta)
    out = {"alpha": alpha, "beta": beta, "T": T, "V": V.T}
    if vectors:
        out["theta"], out["S"], out["Y"] = ritz_pairs(T, V.T)
    else:
        out["theta"] = np.linalg.eigvalsh(T)
    return out


def timed_steps(H, n, v0, reorth=True):
    """Wall-clock seconds for one n-step run (the CPU baseline leg of bench.py)."""
    import time
    t0 = time.perf_counter()
    tridiagonalize(H, n, v0=v0, reorth=reorth)
    return time.perf_counter() - t0


# ---------------------------------------------------------------- operator fixtures

def _shift1d(n, periodic):
    """(n x n) matrix with ones on the +-1 off-diagonals (wrapped when periodic).
    For n == 2 and periodic the two wrapped neighbours coincide and sum to 2, for
    n == 1 all neighbours are the point itself - exactly what COO->CSR duplicate
    summation does to the reference's emission (Hamiltonian.py:61-68, 92-97)."""
    rows, cols = [], []
    for i in range(n):
        for d in (-1, 1):
            k = i + d
            if periodic:
                k %= n
            elif k < 0 or k >= n:
                continue
            rows.append(i)
            cols.append(k)
    return sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))


def laplacian_csr(shape, center, off, periodic=True, diag=None):
    """Structured-grid operator  H = center*I + sum_axis off[axis]*(S+ + S-) (+ diag)
    with the reference's index map  i = x + y*nx + z*nx*ny  (Hamiltonian.py:73-76;
    2-D: i = x + y*nx, tools2.py:35-38) built with Kronecker products; sorted CSR with
    int32 indices.  ``make_golden.py`` checks it against Hamiltonian.create_sparse_T("7")
    bit-for-bit."""
    shape = tuple(int(s) for s in shape)
    dim = len(shape)
    off = [float(off)] * dim if np.isscalar(off) else [float(o) for o in off]
    M = int(np.prod(shape))
    H = sp.identity(M, format="csr") * float(center)
    for ax in range(dim):
        if off[ax] == 0.0:
            continue
        S = _shift1d(shape[ax], periodic)
        # axis 0 (x) is fastest: kron(I_z, kron(I_y, S_x))
        term = S
        for b in range(ax):            # faster axes go to the right
            term = sp.kron(term, sp.identity(shape[b]), format="csr")
        for b in range(ax + 1, dim):   # slower axes go to the left
            term = sp.kron(sp.identity(shape[b]), term, format="csr")
        H = H + off[ax] * term
    if diag is not None:
        H = H + sp.diags(np.asarray(diag, dtype=np.float64))
    H = sp.csr_matrix(H)
    H.sum_duplicates()
    H.sort_indices()
    H.indices = H.indices.astype(np.int32)
    H.indptr = H.indptr.astype(np.int32)
    return H


def box27_weights(T_factor=1.0):
    """(centre, face, edge, corner) coefficients of the reference's 27-point Laplacian:
    3/13 * {-44/3, 1, 1/2, 1/3} * T_factor (Hamiltonian.get_weights_27point, Hamiltonian.py:116-128)."""
    return tuple(T_factor * 3.0 / 13.0 * w for w in (-44.0 / 3.0, 1.0, 0.5, 1.0 / 3.0))


def laplacian27_csr(shape, weights, periodic=True, diag=None):
    """27-point box operator on a 3-D grid: the coefficient of a neighbour depends on how many of its
    three offsets are non-zero (Hamiltonian.py:24,102-128); index map and periodic wrap as the
    7-point one.  Built from Kronecker products of per-axis neighbour sums; sorted CSR, int32."""
    shape = tuple(int(s) for s in shape)
    assert len(shape) == 3
    A = [[sp.identity(n, format="csr"), _shift1d(n, periodic)] for n in shape]    # per axis: I, S
    M = int(np.prod(shape))
    H = sp.csr_matrix((M, M))
    for ez in (0, 1):
        for ey in (0, 1):
            for ex in (0, 1):
                term = sp.kron(A[2][ez], sp.kron(A[1][ey], A[0][ex], format="csr"), format="csr")
                H = H + float(weights[ex + ey + ez]) * term
    if diag is not None:
        H = H + sp.diags(np.asarray(diag, dtype=np.float64))
    H = sp.csr_matrix(H)
    H.sum_duplicates()
    H.sort_indices()
    H.indices = H.indices.astype(np.int32)
    H.indptr = H.indptr.astype(np.int32)
    return H


def reference_T_csr(N, T_factor=1.0):
    """The reference's 7-point periodic Laplacian T on an N^3 grid, emitted the way
    Hamiltonian.create_sparse_T("7") emits it (Hamiltonian.py:20-21, 61-68, 87-99):
    per row the entries [c, -x, -y, -z, +x, +y, +z] with weights [-6, 1, ...]*T_factor,
    wrapped periodically, COO -> CSR (duplicates summed, indices NOT sorted)."""
    i = np.arange(N ** 3, dtype=np.int64)
    x, y, z = i % N, (i // N) % N, i // (N * N)
    rel = [(0, 0, 0), (-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    w = np.ones(7)
    w[0] = -6
    rows, cols, vals = [], [], []
    for k, (dx, dy, dz) in enumerate(rel):
        xx, yy, zz = (x + dx) % N, (y + dy) % N, (z + dz) % N
        rows.append(i)
        cols.append(xx + yy * N + zz * N * N)
        vals.append(np.full(N ** 3, T_factor * w[k]))
    # interleave so that the COO triplets appear row by row in emission order
    rows = np.stack(rows, 1).ravel()
    cols = np.stack(cols, 1).ravel()
    vals = np.stack(vals, 1).ravel()
    return sp.csr_matrix((vals, (rows, cols)), shape=(N ** 3, N ** 3))


def deuteron_potential(x, y, z):
    """Radial deuteron model potential used by the reference drivers
    (3Ddeuteron.py:51-61 == Irregular/Potentials.py:3-13)."""
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    e_well = 54.531
    e_wells = 65.4823128982115
    e_cores = 40.0 * e_well
    return e_cores * np.exp(-(r / 0.25) ** 4.0) - e_wells * np.exp(-(r / 1.7) ** 4.0)


def deuteron_hamiltonian(N, L=25.0):
    """H = -T + V of 3Ddeuteron.py:63-81 with the 7-point T (grid coordinates are
    linspace(-L/2, L/2, N) while dx = L/N, as in Hamiltonian.py:13-17 - kept as is).
    Returns (H sorted CSR, center coefficient, off coefficient, diag potential)."""
    dx = float(L) / N
    T_factor = 197.327 ** 2 / (2 * 469.4592) * 1 / dx ** 2
    g = np.linspace(-L / 2, L / 2, N)
    # V index = x + y*N + z*N^2 with x = g[i], y = g[j], z = g[k]  (Hamiltonian.py:39-44)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    pot = deuteron_potential(X, Y, Z).ravel()
    H = laplacian_csr((N, N, N), 6.0 * T_factor, -T_factor, periodic=True, diag=pot)
    return H, 6.0 * T_factor, -T_factor, pot


def delaunay_graph_laplacian(npts, seed=0):
    """Unweighted graph Laplacian L = D - A of the Delaunay triangulation of ``npts``
    uniform random points in the unit square (BASELINE config 2).  Sorted
    CSR, int32 indices.  Points are emitted in a cell-ordered sequence so that the row
    numbering has spatial locality."""
    from scipy.spatial import Delaunay
    rng = np.random.RandomState(seed)
    pts = rng.uniform(0.0, 1.0, size=(npts, 2))
    g = max(1, int(np.sqrt(npts / 16.0)))
    cell = np.minimum((pts[:, 1] * g).astype(np.int64), g - 1) * g + \
        np.minimum((pts[:, 0] * g).astype(np.int64), g - 1)
    pts = pts[np.argsort(cell, kind="stable")]
    tri = Delaunay(pts)
    s = tri.simplices
    e = np.concatenate([s[:, [0, 1]], s[:, [1, 2]], s[:, [0, 2]]], axis=0)
    e = np.concatenate([e, e[:, ::-1]], axis=0)
    A = sp.csr_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(npts, npts))
    A.data[:] = 1.0                       # duplicates from shared edges collapse to 1
    deg = np.asarray(A.sum(axis=1)).ravel()
    Lm = sp.csr_matrix(sp.diags(deg) - A)
    Lm.sort_indices()
    Lm.indices = Lm.indices.astype(np.int32)
    Lm.indptr = Lm.indptr.astype(np.int32)
    return Lm


def banded_graph_laplacian(M, far=(9000,), seed=0, keep=0.7):
    """Unweighted graph Laplacian of a locality-ordered graph: vertex i is joined to i +- 1..4, i +- 97..100 and
    i +- (f .. f+2) for f in `far`, every candidate edge kept with probability `keep` (ragged rows)."""
    rng = np.random.RandomState(seed)
    offs = [1, 2, 3, 4, 97, 98, 99, 100] + [f + k for f in far for k in range(3)]
    rows, cols = [], []
    for d in offs:
        i = np.nonzero(rng.random_sample(M - d) < keep)[0]
        rows += [i, i + d]
        cols += [i + d, i]
    r, c = np.concatenate(rows), np.concatenate(cols)
    A = sp.csr_matrix((np.ones(r.size), (r, c)), shape=(M, M))
    return sp.csr_matrix(sp.diags(np.asarray(A.sum(axis=1)).ravel()) - A)


def rgg_graph_laplacian(npts, mean_degree=13.0, seed=0, dim=3):
    """Graph Laplacian of a random geometric graph in the unit cube (BASELINE config 4,
    scaled down for tests): radius chosen for the requested mean degree, vertices in
    cell order.  Sorted CSR, int32 indices."""
    from scipy.spatial import cKDTree
    rng = np.random.RandomState(seed)
    pts = rng.uniform(0.0, 1.0, size=(npts, dim))
    vol_unit_ball = {2: np.pi, 3: 4.0 * np.pi / 3.0}[dim]
    radius = (mean_degree / (npts * vol_unit_ball)) ** (1.0 / dim)
    g = max(1, int(1.0 / (2 * radius)))
    key = np.zeros(npts, dtype=np.int64)
    for ax in range(dim - 1, -1, -1):
        key = key * g + np.minimum((pts[:, ax] * g).astype(np.int64), g - 1)
    pts = pts[np.argsort(key, kind="stable")]
    pairs = cKDTree(pts).query_pairs(radius, output_type="ndarray")
    e = np.concatenate([pairs, pairs[:, ::-1]], axis=0)
    A = sp.csr_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(npts, npts))
    deg = np.asarray(A.sum(axis=1)).ravel()
    Lm = sp.csr_matrix(sp.diags(deg) - A)
    Lm.sort_indices()
    Lm.indices = Lm.indices.astype(np.int32)
    Lm.indptr = Lm.indptr.astype(np.int32)
    return Lm


def csr_matvec_rows(indptr, indices, data, x):
    """Row-by-row CSR product with a sequential sum in stored column order - the
    published algorithm of SciPy sparsetools ``csr_matvec`` (scipy 1.18.1,
    scipy/sparse/sparsetools/csr.h; the dependency behind ``H*v`` at Lanczos.py:108,116).
    Pure-Python loops: small cases only."""
    M = len(indptr) - 1
    y = np.zeros(M)
    for i in range(M):
        s = 0.0
        for k in range(indptr[i], indptr[i + 1]):
            s += data[k] * x[indices[k]]
        y[i] = s
    return y
