"""On-disk formats either side of the hot path, host-side Python like the reference's.

  * the operator cache of `Hamiltonian.create_sparse_T` (Hamiltonian.py:48-53,69):
    `T_matrices/T_N=<N>_Laplace=<points>.npz`, a `scipy.sparse.save_npz` file;
  * the result files of the drivers (3Ddeuteron.py:99-100): `eigvals.npy`, `eigvecs.npy`;
  * the Mathematica-style matrix dump of `MatrixWrite.py:37-62`;
  * a solver checkpoint (alpha, beta, Ritz data and optionally the Krylov basis) - the reference has
    none; it restores a finished run so that `get_H_eigs` / `print_good_eigs` work without re-running.

Nothing here touches the GPU except `restore_checkpoint`, which uploads a saved basis.
"""
from __future__ import annotations

import os
import re

import numpy as np

from .engine import StencilOperator

T_DIR = "T_matrices"


# ---------------------------------------------------------------- operator cache (Hamiltonian.py:48-69)
def t_matrix_path(N: int, points="27", root: str = T_DIR) -> str:
    return os.path.join(root, "T_N=%d_Laplace=%s.npz" % (N, points))


def load_t_matrix(N: int, points="27", root: str = T_DIR):
    """The cached Laplacian of `create_sparse_T` as a scipy CSR matrix, or None when it was never
    written (the reference then builds it with Python loops, Hamiltonian.py:55-69)."""
    import scipy.sparse as sp
    path = t_matrix_path(N, points, root)
    if not os.path.isfile(path):
        return None
    return sp.csr_matrix(sp.load_npz(path))


def save_t_matrix(T, N: int, points="27", root: str = T_DIR) -> str:
    import scipy.sparse as sp
    os.makedirs(root, exist_ok=True)
    path = t_matrix_path(N, points, root)
    sp.save_npz(path, sp.csr_matrix(T))
    return path


def stencil_from_t_matrix(T, N: int, potential=None, sign: float = -1.0) -> StencilOperator:
    """Matrix-free descriptor of H = sign*T + V for a cached reference Laplacian T (H = -T + V,
    3Ddeuteron.py:80): the stencil weights are read off row 0 of T (centre / face / edge / corner by
    the number of non-zero offsets, Hamiltonian.py:102-128) and a sample of rows is checked against
    them.  Raises ValueError when T is not a periodic 7- or 27-point stencil on an N^3 grid."""
    import scipy.sparse as sp
    T = sp.csr_matrix(T)
    M = N ** 3
    if T.shape != (M, M):
        raise ValueError(f"T has shape {T.shape}, expected ({M}, {M})")

    def row_weights(i):
        x, y, z = i % N, (i // N) % N, i // (N * N)
        w = {}
        for k in range(T.indptr[i], T.indptr[i + 1]):
            c = int(T.indices[k])
            off = [(c % N - x), ((c // N) % N - y), (c // (N * N) - z)]
            off = [(o + N // 2) % N - N // 2 if N > 2 else o for o in off]      # periodic wrap -> -1, 0, 1
            if any(abs(o) > 1 for o in off):
                raise ValueError(f"row {i} couples to a point further than one cell away")
            kind = sum(1 for o in off if o != 0)
            w.setdefault(kind, set()).add(float(T.data[k]))
        if any(len(v) != 1 for v in w.values()):
            raise ValueError(f"row {i}: weights are not uniform per neighbour class")
        return {k: v.pop() for k, v in w.items()}

    if N < 3:
        raise ValueError("stencil_from_t_matrix needs N >= 3 (neighbours coincide on smaller periodic grids)")
    w0 = row_weights(0)
    for i in {0, 1, N - 1, N, M // 2, M - 1}:
        if row_weights(i) != w0:
            raise ValueError("T is not a translation-invariant stencil")
    diag = None if potential is None else np.asarray(potential, dtype=np.float64).reshape(-1)
    if set(w0) <= {0, 1}:
        return StencilOperator((N, N, N), sign * w0.get(0, 0.0), sign * w0.get(1, 0.0), diag=diag)
    w = tuple(sign * w0.get(k, 0.0) for k in range(4))
    return StencilOperator((N, N, N), 0.0, 0.0, weights27=w, diag=diag)


# ---------------------------------------------------------------- results (3Ddeuteron.py:99-100)
def save_eigs(eigvals, eigvecs, directory: str = "."):
    """np.save("eigvals.npy", l_L); np.save("eigvecs.npy", v_L)"""
    os.makedirs(directory, exist_ok=True)
    np.save(os.path.join(directory, "eigvals.npy"), np.asarray(eigvals))
    np.save(os.path.join(directory, "eigvecs.npy"), np.asarray(eigvecs))


def load_eigs(directory: str = "."):
    return np.load(os.path.join(directory, "eigvals.npy")), np.load(os.path.join(directory, "eigvecs.npy"))


# ---------------------------------------------------------------- MatrixWrite.py:37-62
def matrix_dat_name(d, N, L, p) -> str:
    return f"matrix_d={d}_N={N}_L={L}_p={p}.dat"


def write_matrix_dat(path: str, H, d: int, L, N: int, p: str) -> str:
    """The Mathematica-readable dump of MatrixWrite.py: header (numd, nrpoints, box, potential), then
    `H = {{rows, cols}, {{row, col, value},\\n ...}};` with values printed as %.17f in COO order."""
    import scipy.sparse as sp
    A = sp.coo_matrix(H)
    # the separator before the third box length is a no-break space (U+00A0) in the reference's source
    # (MatrixWrite.py:41); it is kept so that the files are byte-identical
    head = (f"numd = {d:d};\nnrpoints = {sp.csr_matrix(H).count_nonzero():d};\n"
            f"box = {{{L:g}, {L:g},\u00a0{L:g}}};\npotential = \"{p}\";\nH = {{{{{N**3:d}, {N**3:d}}}, {{")
    body = "".join("{%d, %d, %.17f},\n" % (r, c, v) for r, c, v in zip(A.row, A.col, A.data))
    with open(path, "w", encoding="utf-8") as f:
        f.write(head + body + "}};")
    return path


_ENTRY = re.compile(r"\{(\d+), (\d+), (-?[0-9.]+(?:e[-+]?\d+)?)\},")


def read_matrix_dat(path: str):
    """(meta, scipy COO matrix) from a MatrixWrite dump."""
    import scipy.sparse as sp
    text = open(path, encoding="utf-8").read()
    meta = {"numd": int(re.search(r"numd = (\d+);", text).group(1)),
            "nrpoints": int(re.search(r"nrpoints = (\d+);", text).group(1)),
            "box": tuple(float(x) for x in re.search(r"box = \{([^}]*)\};", text).group(1).split(",")),
            "potential": re.search(r'potential = "([^"]*)";', text).group(1)}
    rows, cols = (int(x) for x in re.search(r"H = \{\{(\d+), (\d+)\}", text).groups())
    ent = _ENTRY.findall(text)
    r = np.array([int(e[0]) for e in ent], dtype=np.int64)
    c = np.array([int(e[1]) for e in ent], dtype=np.int64)
    v = np.array([float(e[2]) for e in ent], dtype=np.float64)
    return meta, sp.coo_matrix((v, (r, c)), shape=(rows, cols))


# ---------------------------------------------------------------- checkpoint (no reference counterpart)
def save_checkpoint(path: str, solver, with_basis: bool = True) -> str:
    """alpha/beta (as H_eff's diagonals), n, M and - optionally - the normalised basis rows of a
    finished run, as one .npz."""
    T = solver.H_eff
    out = {"alpha": np.diag(T).copy(), "beta": np.diag(T, 1).copy(), "n": np.int64(T.shape[0]),
           "M": np.int64(solver.M), "format": np.int64(1)}
    base = path[:-4] if path.endswith(".npz") else path
    team = getattr(solver, "_team", None)
    if team is not None:
        # row-sharded run: the basis is written shard by shard (no process ever holds all of it);
        # the header file carries the partition so that a restore can check it
        out["shard_starts"] = np.asarray([team.plan.rows(r)[0] for r in range(team.world)] + [team.M], dtype=np.int64)
        if with_basis:
            for s, r in zip(team.shards, team.results):
                r0, r1 = team.plan.rows(s.rank)
                np.savez(shard_path(base, s.rank), V_rows=r.basis_rows_host(), rows=np.asarray([r0, r1], dtype=np.int64),
                         format=np.int64(1))
            out["sharded_basis"] = np.int64(1)
        if team.shards[0].rank == 0:
            np.savez(base, **out)
        return base + ".npz"
    if with_basis:
        out["V_rows"] = solver.result.basis_rows_host()          # (n, M), rows = Lanczos vectors
    np.savez(base, **out)
    return base + ".npz"


def shard_path(base: str, rank: int) -> str:
    return "%s.shard%d.npz" % (base, rank)


def load_checkpoint(path: str) -> dict:
    with np.load(path) as z:
        d = {k: z[k] for k in z.files}
    if int(d.get("format", 0)) != 1:
        raise ValueError("not a lanczos_b200 checkpoint")
    return d


def restore_checkpoint(solver, path: str, devices=None, fmt="auto", sigma=0):
    """Put a saved run back into `solver` (a Lanczos / IrrLanczos instance for the same operator):
    H_eff on the host, the basis - when it was saved - back in HBM, so that get_H_eigs, H_eigvecs and
    print_good_eigs work without running the loop again.  A checkpoint of a row-sharded run is
    restored shard by shard onto `devices` (same meaning as in execute_Lanczos; the partition must
    be the one the checkpoint was written with)."""
    import torch
    from . import engine
    from ._capi import RunInfo
    d = load_checkpoint(path)
    n, M = int(d["n"]), int(d["M"])
    if M != solver.M:
        raise ValueError(f"checkpoint is for M = {M}, the operator has M = {solver.M}")
    if "shard_starts" in d and devices is not None:
        team = solver._team_for(devices, fmt, sigma)
        if team is None:
            raise ValueError("restore_checkpoint: `devices` does not describe a row-sharded run")
        starts = [team.plan.rows(r)[0] for r in range(team.world)] + [team.M]
        if list(d["shard_starts"]) != starts:
            raise ValueError("the checkpoint was written with another partition of the rows")
        team._ensure_team(n)
        base = path[:-4] if path.endswith(".npz") else path
        results = []
        for s in team.shards:
            Ml = team.plan.local_rows(s.rank)
            ld = engine.padded_ld(Ml)
            V_dev = None
            if int(d.get("sharded_basis", 0)):
                with np.load(shard_path(base, s.rank)) as z:
                    rows, Vr = z["rows"], z["V_rows"]
                if tuple(rows) != tuple(team.plan.rows(s.rank)):
                    raise ValueError(f"shard file of rank {s.rank} covers other rows")
                with torch.cuda.device(s.ctx.device):
                    V_dev = torch.zeros((n, ld), dtype=torch.float64, device=s.ctx.torch_device)
                    V_dev[:, :Ml] = torch.from_numpy(np.ascontiguousarray(Vr)).to(s.ctx.torch_device)
            info = RunInfo()
            info.steps_done = n
            results.append(engine.LanczosResult(s.ctx, n, Ml, d["alpha"].copy(), d["beta"].copy(), V_dev, ld, np.ones(n), info))
        team.n = n
        team._results = results
        team._H_eff = results[0].tridiagonal()
        team.Lanczos_has_been_executed = True
        solver.n = n
        solver._team = team
        solver._result = results[0]
        solver._H_eff = team._H_eff
        solver._V_host = None
        solver._Y_dev = None
        solver.H_eigs_have_been_found = False
        solver.Lanczos_has_been_executed = True
        return solver
    if "shard_starts" in d and int(d.get("sharded_basis", 0)):
        # sharded checkpoint onto one GPU: stitch the shard files together
        base = path[:-4] if path.endswith(".npz") else path
        world = len(d["shard_starts"]) - 1
        d = dict(d)
        d["V_rows"] = np.concatenate([np.load(shard_path(base, r))["V_rows"] for r in range(world)], axis=1)
    ctx = engine.Context.default()
    ld = engine.padded_ld(M)
    V_dev = None
    if "V_rows" in d:
        V_dev = torch.zeros((n, ld), dtype=torch.float64, device=ctx.torch_device)
        V_dev[:, :M] = torch.from_numpy(np.ascontiguousarray(d["V_rows"])).to(ctx.torch_device)
    info = RunInfo()
    info.steps_done = n
    res = engine.LanczosResult(ctx, n, M, d["alpha"].copy(), d["beta"].copy(), V_dev, ld, np.ones(n), info)
    solver.n = n
    solver._result = res
    solver._H_eff = res.tridiagonal()
    solver._V_host = None
    solver._team = None
    solver._Y_dev = None
    solver.H_eigs_have_been_found = False
    solver.Lanczos_has_been_executed = True
    return solver
