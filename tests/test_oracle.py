"""The oracle (oracle/lanczos_oracle.py) is held to the golden vectors produced by the
live reference (tests/golden/make_golden.py).  Same NumPy ops in the same order =>
the comparison is bit-for-bit.  CPU only."""
import hashlib

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lanczos_oracle as orc



def same_floats(got, want, rtol=1e-12):
    """Bit for bit where the host runs the BLAS kernels the golden files were made with (the authoring
    container); on another CPU OpenBLAS may pick different dot/axpy kernels, whose reductions round
    differently - there the restatement must still agree to `rtol` (a warning records the downgrade)."""
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        return False
    if np.array_equal(got, want):
        return True
    ok = np.allclose(got, want, rtol=rtol, atol=rtol * float(np.max(np.abs(want))))
    if ok:
        import warnings
        warnings.warn("oracle agrees with the golden vectors to %.0e but not bit for bit on this host's BLAS" % rtol)
    return ok


def _tri(T):
    return np.diag(T).copy(), np.diag(T, 1).copy()


@pytest.mark.parametrize("N", [2, 3, 5])
def test_pattern_T7_matches_reference(golden, N):
    # kron generator == Hamiltonian.create_sparse_T("7") after sort_indices, bit-exact
    T = orc.laplacian_csr((N, N, N), -6.0 * 1.75, 1.75, periodic=True)
    assert np.array_equal(T.indptr, golden[f"T7_N{N}_indptr"])

    assert np.array_equal(T.indices, golden[f"T7_N{N}_indices"])

    assert same_floats(T.data, golden[f"T7_N{N}_data"])
    # emission-order restatement (COO with duplicates) gives the same matrix
    T2 = orc.reference_T_csr(N, 1.75)
    T2.sum_duplicates()
    T2.sort_indices()
    assert np.array_equal(T2.indptr, golden[f"T7_N{N}_indptr"])

    assert np.array_equal(T2.indices, golden[f"T7_N{N}_indices"])

    assert same_floats(T2.data, golden[f"T7_N{N}_data"])


@pytest.mark.parametrize("N", [3, 5])
def test_deuteron_H_matches_reference(golden, N):
    dx = 25.0 / N
    g = np.linspace(-12.5, 12.5, N)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    pot = orc.deuteron_potential(X, Y, Z).ravel()
    H = orc.laplacian_csr((N, N, N), 6.0 * 1.75, -1.75, periodic=True, diag=pot)
    assert np.array_equal(H.indptr, golden[f"H_N{N}_indptr"])

    assert np.array_equal(H.indices, golden[f"H_N{N}_indices"])

    np.testing.assert_allclose(H.data, golden[f"H_N{N}_data"], rtol=1e-15, atol=0)


def test_c1_small_bit_exact(golden):
    H = orc.laplacian_csr((24, 20), 4.0, -1.0, periodic=True)
    res = orc.lanczos(H, 30, seed=99, vectors=True)
    assert same_floats(res["alpha"], golden["c1s_alpha"])
    assert same_floats(res["beta"], golden["c1s_beta"])
    assert same_floats(res["theta"], golden["c1s_theta"])
    assert same_floats(res["V"][:, :3].T, golden["c1s_V_first3"])


@pytest.mark.parametrize("tag,per", [("c1d", False), ("c1p", True)])
def test_c1_full_bit_exact(golden, tag, per):
    H = orc.laplacian_csr((200, 200), 4.0, -1.0, periodic=per)
    res = orc.lanczos(H, 100, seed=99)
    assert same_floats(res["alpha"], golden[f"{tag}_alpha"])
    assert same_floats(res["beta"], golden[f"{tag}_beta"])
    np.testing.assert_allclose(res["theta"], golden[f"{tag}_theta"], rtol=1e-13, atol=1e-13)


def test_c3_small_user_start_vector(golden):
    H = orc.laplacian_csr((12, 12, 12), 6.0, -1.0, periodic=True)
    v0 = np.random.RandomState(7).uniform(-1, 1, 12 ** 3)
    res = orc.lanczos(H, 40, v0=v0)
    assert same_floats(res["alpha"], golden["c3s_alpha"])
    assert same_floats(res["beta"], golden["c3s_beta"])


def test_deuteron_bit_exact(golden):
    H, _, _, _ = orc.deuteron_hamiltonian(16)
    res = orc.lanczos(H, 120, seed=78)
    assert same_floats(res["alpha"], golden["deut_alpha"])
    assert same_floats(res["beta"], golden["deut_beta"])
    np.testing.assert_allclose(res["theta"], golden["deut_theta"], rtol=1e-12, atol=1e-10)


def test_delaunay_csr_and_csc(golden):
    L = orc.delaunay_graph_laplacian(3000, seed=0)
    sha = hashlib.sha256(L.indptr.tobytes() + L.indices.tobytes()).digest()
    assert np.array_equal(np.frombuffer(sha, dtype=np.uint8), golden["del_indptr_sha"])

    res = orc.lanczos(L, 50, seed=99)
    assert same_floats(res["alpha"], golden["del_alpha"])
    assert same_floats(res["beta"], golden["del_beta"])
    res = orc.lanczos(sp.csc_matrix(L), 50, seed=99)
    assert same_floats(res["alpha"], golden["delcsc_alpha"])
    assert same_floats(res["beta"], golden["delcsc_beta"])


def test_edge_cases(golden):
    H = orc.laplacian_csr((6,), 2.0, -1.0, periodic=False)
    assert same_floats(orc.lanczos(H, 2, seed=3)["T"], golden["n2_T"])
    assert same_floats(orc.lanczos(H, 6, seed=3)["T"], golden["nM_T"])
    with pytest.raises(IndexError):
        orc.tridiagonalize(H, 1)
    with pytest.raises(ValueError):
        orc.tridiagonalize(H, 7)


def test_start_vector_discarded():
    # quirk 1 of SURVEY §0: row 0 of the basis is orthogonal to the user's start vector
    H = orc.laplacian_csr((30, 30), 4.0, -1.0, periodic=False)
    v0 = orc.start_vector(900, seed=99)
    _, _, V = orc.tridiagonalize(H, 5, seed=99)
    assert abs(np.dot(V[0], v0)) < 1e-14


def test_csr_matvec_rows_matches_scipy():
    H = orc.laplacian_csr((5, 4, 3), 6.0, -1.0, periodic=True)
    x = np.random.RandomState(1).uniform(-1, 1, 60)
    y = orc.csr_matvec_rows(H.indptr, H.indices, H.data, x)
    assert same_floats(y, H * x)


def test_rgg_laplacian_shape():
    L = orc.rgg_graph_laplacian(4000, mean_degree=13.0, seed=0)
    assert L.shape == (4000, 4000)
    assert abs(L.sum()) < 1e-9
    assert 9.0 < (L.nnz / 4000.0 - 1.0) < 14.0   # boundary effects lower the mean degree
    assert (abs(L - L.T)).nnz == 0


@pytest.mark.parametrize("N", [2, 3, 5])
def test_pattern_T27_matches_reference(golden, N):
    # the reference's default 27-point Laplacian (Hamiltonian.create_sparse_T("27"))
    T = orc.laplacian27_csr((N, N, N), orc.box27_weights(1.75), periodic=True)
    assert np.array_equal(T.indptr, golden[f"T27_N{N}_indptr"])

    assert np.array_equal(T.indices, golden[f"T27_N{N}_indices"])

    np.testing.assert_allclose(T.data, golden[f"T27_N{N}_data"], rtol=4e-16, atol=1e-15)


def _deuteron27(N):
    dx = 25.0 / N
    Tf = 197.327 ** 2 / (2 * 469.4592) * 1 / dx ** 2
    g = np.linspace(-12.5, 12.5, N)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    pot = orc.deuteron_potential(X, Y, Z).ravel()
    w = tuple(-x for x in orc.box27_weights(Tf))          # H = -T + V
    return orc.laplacian27_csr((N, N, N), w, periodic=True, diag=pot), w, pot


def test_deuteron27_matches_reference(golden):
    H, _, _ = _deuteron27(12)
    res = orc.lanczos(H, 60, seed=78)
    np.testing.assert_allclose(res["alpha"], golden["deut27_alpha"], rtol=1e-12)
    np.testing.assert_allclose(res["beta"], golden["deut27_beta"], rtol=1e-12)


def test_blocked_sweep_is_bit_identical():
    """gram_schmidt_row_blocked (used by the full-size config-2 GPU parity test) against the plain
    restatement of Lanczos.py:247-249, bit for bit, zero tail rows included."""
    rs = np.random.RandomState(1)
    for n, M, j, nrows in [(7, 1000, 3, 7), (9, 100003, 4, 5), (5, 70000, 0, 1), (6, 33, 5, 6), (4, 129, 1, 2)]:
        V = rs.uniform(-1, 1, (n, M)) / np.sqrt(M)
        V[nrows:] = 0.0
        A, B = V.copy(), V.copy()
        orc.gram_schmidt_row(A, j)
        orc.gram_schmidt_row_blocked(B, j, nrows=nrows, threads=3)
        assert np.array_equal(A, B)
    H = orc.delaunay_graph_laplacian(2000, seed=0)
    a, b = orc.lanczos(H, 30, seed=99), orc.lanczos(H, 30, seed=99, blocked=True)
    assert np.array_equal(a["alpha"], b["alpha"]) and np.array_equal(a["beta"], b["beta"]) and np.array_equal(a["V"], b["V"])


def test_gpu_form_sweep_matches_the_reference_expression():
    """Regular/Lanczos.py:236-238 (use_cuda=True branch, CuPy there): the same expressions on NumPy arrays."""
    rs = np.random.RandomState(2)
    V = rs.uniform(-1, 1, (6, 500)) / np.sqrt(500)
    j = 2
    W = V.copy()
    ip = np.sum(W[j] * W, axis=1)
    ip[j] = 0
    W[j] = W[j] - np.sum(ip[:, None] * W, axis=0)
    U = V.copy()
    orc.gram_schmidt_row_gpu_form(U, j)
    assert np.array_equal(U, W)
    # the two forms agree to rounding on a normalised row
    V[j] /= np.linalg.norm(V[j])
    A, B = V.copy(), V.copy()
    orc.gram_schmidt_row(A, j)
    orc.gram_schmidt_row_gpu_form(B, j)
    assert np.max(np.abs(A - B)) < 1e-15


def test_vendored_reference_matches_golden():
    """oracle/_ref (bench.py's reference arm) is the unmodified reference: it reproduces the golden alpha/beta."""
    from oracle import build_ref
    mods = build_ref.load()
    if mods is None:
        pytest.skip("oracle/_ref not built (no /root/reference in reach)")
    import contextlib
    import io
    man = build_ref.manifest()
    assert set(man["files"]) == {"Lanczos.py", "IrrLanczos.py", "Hamiltonian.py"}
    H = orc.laplacian_csr((200, 200), 4.0, -1.0, periodic=False)
    L = mods[0].Lanczos(H)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        L.execute_Lanczos(20, seed=99, use_cuda=False)
    ref = orc.lanczos(H, 20, seed=99)
    assert np.array_equal(np.diag(L.H_eff), ref["alpha"]) and np.array_equal(np.diag(L.H_eff, 1), ref["beta"])


def test_banded_graph_laplacian_generator():
    """Test matrix of the windowed SELL form: symmetric, zero row sums, -1 off the diagonal, columns only in the
    stated bands around the diagonal."""
    L = orc.banded_graph_laplacian(5000, far=(900,), seed=1)
    assert (L != L.T).nnz == 0
    assert np.max(np.abs(L @ np.ones(5000))) == 0.0
    C = L.tocoo()
    d = np.abs(C.row - C.col)
    assert set(np.unique(C.data[d > 0])) == {-1.0}
    assert set(np.unique(d)) <= {0, 1, 2, 3, 4, 97, 98, 99, 100, 900, 901, 902}
