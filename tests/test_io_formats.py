"""On-disk formats either side of the hot path (lanczos_b200/io.py) against
files written by the reference itself (tests/golden/make_golden_io.py)."""
import filecmp
import os

import numpy as np
import pytest
import scipy.sparse as sp

from lanczos_b200 import io as lzio
from oracle import lanczos_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("points", ["7", "27"])
def test_reads_the_reference_t_matrix_cache(points, tmp_path):
    T = lzio.load_t_matrix(4, points, root=GOLD)                  # written by Hamiltonian.create_sparse_T
    assert T is not None and T.shape == (64, 64)
    assert lzio.load_t_matrix(5, points, root=GOLD) is None       # never cached -> the reference would build it
    # the oracle's generator restates the same operator
    R = (orc.reference_T_csr(4, T_factor=1.5) if points == "7"
         else orc.laplacian27_csr((4, 4, 4), orc.box27_weights(1.5), periodic=True))
    assert abs(T - R).max() < 1e-15
    # round trip through our writer: a file the reference can load back (scipy.sparse.load_npz)
    path = lzio.save_t_matrix(T, 4, points, root=str(tmp_path / "T_matrices"))
    assert os.path.basename(path) == "T_N=4_Laplace=%s.npz" % points
    back = sp.load_npz(path)
    assert (back != T).nnz == 0


@pytest.mark.parametrize("points", ["7", "27"])
def test_stencil_descriptor_from_cached_t(points):
    T = lzio.load_t_matrix(4, points, root=GOLD)
    op = lzio.stencil_from_t_matrix(T, 4, sign=1.0)
    assert op.grid == (4, 4, 4) and op.bc == "periodic"
    if points == "7":
        assert op.weights27 is None
        assert op.center == T[0, 0] and op.off == (T[0, 1],) * 3
    else:
        w = op.weights27
        assert w[0] == T[0, 0] and w[1] == T[0, 1] and w[2] == T[0, 5] and w[3] == T[0, 21]
        assert np.allclose(w, orc.box27_weights(1.5), rtol=1e-15)
    with pytest.raises(ValueError):
        lzio.stencil_from_t_matrix(T, 5)
    B = sp.lil_matrix(T)
    B[7, 7] += 1.0
    with pytest.raises(ValueError):
        lzio.stencil_from_t_matrix(B, 4)
        lzio.stencil_from_t_matrix(sp.identity(64) + sp.eye(64, k=9), 4)


def test_matrix_dat_is_byte_identical_to_the_reference_writer(tmp_path):
    A = sp.load_npz(os.path.join(GOLD, "matrix_dat_input.npz"))
    name = lzio.matrix_dat_name(3, 2, 25, "Deuteron")
    assert name == "matrix_d=3_N=2_L=25_p=Deuteron.dat"
    out = lzio.write_matrix_dat(str(tmp_path / name), A, 3, 25, 2, "Deuteron")
    assert filecmp.cmp(out, os.path.join(GOLD, name), shallow=False)
    meta, B = lzio.read_matrix_dat(out)
    assert meta == {"numd": 3, "nrpoints": A.count_nonzero(), "box": (25.0, 25.0, 25.0), "potential": "Deuteron"}
    assert abs(sp.csr_matrix(B) - A).max() < 1e-16


def test_eig_files_round_trip(tmp_path):
    l, v = np.linspace(-2.2, 3.0, 7), np.random.RandomState(0).rand(30, 7)
    lzio.save_eigs(l, v, str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["eigvals.npy", "eigvecs.npy"]      # 3Ddeuteron.py:99-100
    l2, v2 = lzio.load_eigs(str(tmp_path))
    assert np.array_equal(l, l2) and np.array_equal(v, v2)


def test_checkpoint_rejects_foreign_files(tmp_path):
    np.savez(tmp_path / "x.npz", alpha=np.zeros(3))
    with pytest.raises(ValueError):
        lzio.load_checkpoint(str(tmp_path / "x.npz"))


@pytest.mark.gpu
def test_cached_t_matrix_runs_matrix_free_and_checkpoints(tmp_path):
    """H = -T + V from the reference's cache file: the matrix-free operator read off T gives the
    tridiagonal matrix of the stored one; a checkpoint restores the run (Ritz vectors included)."""
    import lanczos_b200 as lz
    T = lzio.load_t_matrix(4, "27", root=GOLD)
    pot = np.linspace(0.0, 2.0, 64)
    H = sp.csr_matrix(-T + sp.diags(pot))
    ref = orc.lanczos(H, 12, seed=3)
    op = lzio.stencil_from_t_matrix(T, 4, potential=pot)
    assert abs(op.tocsr() - H).max() < 1e-14
    L = lz.Lanczos(op)
    L.execute_Lanczos(12, seed=3)
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert np.max(np.abs(a - ref["alpha"]) / np.abs(ref["alpha"])) < 1e-12
    assert np.max(np.abs(b - ref["beta"]) / np.abs(ref["beta"])) < 1e-12
    ck = lzio.save_checkpoint(str(tmp_path / "run.npz"), L)
    vals, vecs = L.H_eigvals.copy(), L.H_eigvecs.copy()
    R = lzio.restore_checkpoint(lz.Lanczos(op), ck)
    assert np.array_equal(R.H_eff, L.H_eff)
    assert np.allclose(R.H_eigvals, vals, rtol=0, atol=1e-14)
    assert np.max(np.abs(np.abs(R.H_eigvecs) - np.abs(vecs))) < 1e-12
    assert np.max(np.abs(R.V - L.V)) == 0.0
    lzio.save_eigs(R.H_eigvals, R.H_eigvecs, str(tmp_path))
    with pytest.raises(ValueError):
        lzio.restore_checkpoint(lz.Lanczos(lz.StencilOperator((5, 5, 5), 6.0, -1.0)), ck)
