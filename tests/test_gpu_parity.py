"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in classes) against the
oracle and the committed golden vectors of the live reference.

Tolerances (BASELINE.json north_star): sparsity pattern / halo indexing bit-exact; alpha/beta
1e-12 relative (m <= 50, full reorth); converged Ritz values 1e-10 relative.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lanczos_oracle as orc

pytestmark = pytest.mark.gpu

TOL_AB = 1e-12
TOL_RITZ = 1e-10


@pytest.fixture(scope="module")
def lz():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lanczos_b200
    return lanczos_b200


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


GRIDS = [
    ((7,), "periodic"), ((7,), "dirichlet"), ((1,), "periodic"), ((2,), "periodic"),
    ((6, 5), "periodic"), ((6, 5), "dirichlet"), ((5, 4), "periodic"), ((70, 3), "dirichlet"),
    ((4, 3, 5), "periodic"), ((4, 3, 5), "dirichlet"), ((5, 3, 2), "periodic"), ((2, 2, 2), "periodic"),
    ((66, 9, 3), "periodic"), ((130, 17, 4), "dirichlet"), ((3, 1, 4), "periodic"),
]


@pytest.mark.parametrize("grid,bc", GRIDS)
def test_stencil_pattern_bit_exact(lz, grid, bc):
    dim = len(grid)
    center = 2.0 * dim + 1.0      # (+1: keeps the degenerate 1-point grid from cancelling to an empty row)
    A = orc.laplacian_csr(grid, center, -1.0, periodic=(bc == "periodic"))
    op = lz.StencilOperator(grid, center, -1.0, bc=bc)
    E = op.tocsr()
    assert np.array_equal(E.indptr, A.indptr)
    assert np.array_equal(E.indices, A.indices)
    assert np.array_equal(E.data, A.data)
    # probe the kernel itself with unit vectors: column i of H, exactly (entries are small integers)
    M = A.shape[0]
    if M <= 400:
        D = A.toarray()
        for i in range(M):
            e = np.zeros(M)
            e[i] = 1.0
            assert np.array_equal(op.matvec(e), D[:, i]), f"column {i}"


@pytest.mark.parametrize("grid,bc", [((64, 64, 64), "periodic"), ((100, 37, 11), "dirichlet"),
                                      ((33, 20, 7), "periodic"), ((200, 200), "periodic"),
                                      ((200, 200), "dirichlet"), ((1001,), "dirichlet")])
def test_stencil_apply_values(lz, grid, bc):
    dim = len(grid)
    off = [-1.0, -0.75, -1.25][:dim]
    A = orc.laplacian_csr(grid, 2.0 * dim + 0.1, off, periodic=(bc == "periodic"))
    op = lz.StencilOperator(grid, 2.0 * dim + 0.1, off, bc=bc)
    x = np.random.RandomState(5).uniform(-1, 1, A.shape[0])
    y = op.matvec(x)
    ref = A * x
    assert np.max(np.abs(y - ref)) <= 8e-16 * np.max(np.abs(ref)) * 8


def test_stencil_with_potential_matches_reference_H(lz, golden):
    N = 5
    g = np.linspace(-12.5, 12.5, N)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    pot = orc.deuteron_potential(X, Y, Z).ravel()
    op = lz.StencilOperator((N, N, N), 6.0 * 1.75, -1.75, diag=pot)
    E = op.tocsr()
    assert np.array_equal(E.indptr, golden["H_N5_indptr"])
    assert np.array_equal(E.indices, golden["H_N5_indices"])
    np.testing.assert_allclose(E.data, golden["H_N5_data"], rtol=1e-15)


@pytest.mark.parametrize("fmt", ["csr", "sell"])
def test_sparse_apply_and_export(lz, fmt):
    L = orc.delaunay_graph_laplacian(5000, seed=1)
    ctx = lz.Context.default()
    op = lz.DeviceOperator.from_scipy(ctx, L, fmt=fmt)
    E = op.export_csr()
    assert np.array_equal(E.indptr, L.indptr)
    assert np.array_equal(E.indices, L.indices)
    assert np.array_equal(E.data, L.data)
    x = np.random.RandomState(2).uniform(-1, 1, 5000)
    y = op.apply_host(x)
    ref = L * x
    assert np.max(np.abs(y - ref)) <= 1e-14 * np.max(np.abs(ref))
    t, s = op.nnz()
    assert t == L.nnz and (s >= t)
    if fmt == "sell":
        assert s <= 1.15 * t          # sigma-sorting keeps the padding small


def test_sparse_ragged_rows(lz):
    # empty rows, one dense row, M not a multiple of 32
    rs = np.random.RandomState(3)
    M = 77
    A = sp.random(M, M, density=0.05, random_state=rs, format="lil")
    A[5, :] = rs.uniform(-1, 1, M)
    A[9, :] = 0
    A = sp.csr_matrix(A)
    A = sp.csr_matrix(A + A.T)
    A.sort_indices()
    x = rs.uniform(-1, 1, M)
    ctx = lz.Context.default()
    for fmt in ("csr", "sell"):
        op = lz.DeviceOperator.from_scipy(ctx, A, fmt=fmt)
        y = op.apply_host(x)
        np.testing.assert_allclose(y, A * x, rtol=0, atol=1e-14 * np.abs(A * x).max())


def _check_run(L, ref, n_conv=None):
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert rel(a, ref["alpha"]) < TOL_AB
    assert rel(b, ref["beta"]) < TOL_AB


def test_c1_small_vs_golden_and_oracle(lz, golden):
    op = lz.StencilOperator((24, 20), 4.0, -1.0)
    L = lz.Lanczos(op)
    L.execute_Lanczos(30, seed=99)
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert rel(a, golden["c1s_alpha"]) < TOL_AB
    assert rel(b, golden["c1s_beta"]) < TOL_AB
    np.testing.assert_allclose(L.H_eigvals, golden["c1s_theta"], rtol=TOL_RITZ, atol=1e-13)
    # first three Lanczos vectors (rows of the reference's in-loop basis)
    V3 = L.V[:, :3].T
    assert np.max(np.abs(V3 - golden["c1s_V_first3"])) < 1e-13


@pytest.mark.parametrize("tag,bc", [("c1d", "dirichlet"), ("c1p", "periodic")])
def test_c1_full_config(lz, golden, tag, bc):
    """BASELINE config 1: 2-D 5-point Laplacian 200x200, m = 100, lowest 10 Ritz values."""
    op = lz.StencilOperator((200, 200), 4.0, -1.0, bc=bc)
    L = lz.Lanczos(op)
    L.execute_Lanczos(100, seed=99)
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert rel(a[:50], golden[f"{tag}_alpha"][:50]) < TOL_AB
    assert rel(b[:50], golden[f"{tag}_beta"][:50]) < TOL_AB
    assert rel(a, golden[f"{tag}_alpha"]) < 1e-11
    assert rel(b, golden[f"{tag}_beta"]) < 1e-11
    lowest = L.ritz_values(10)
    np.testing.assert_allclose(lowest, golden[f"{tag}_theta"][:10], rtol=TOL_RITZ, atol=1e-12)


def test_c1_same_through_csr_input(lz, golden):
    # the same operator given as the scipy matrix the reference would hold
    H = orc.laplacian_csr((200, 200), 4.0, -1.0, periodic=False)
    for fmt in ("csr", "sell"):
        L = lz.Lanczos(H)
        L.execute_Lanczos(60, seed=99, fmt=fmt)
        a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
        assert rel(a[:50], golden["c1d_alpha"][:50]) < TOL_AB
        assert rel(b[:50], golden["c1d_beta"][:50]) < TOL_AB


def test_c3_small_user_start_vector(lz, golden):
    v0 = np.random.RandomState(7).uniform(-1, 1, 12 ** 3)
    L = lz.Lanczos(lz.StencilOperator((12, 12, 12), 6.0, -1.0))
    L.execute_Lanczos(40, v0=v0)
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert rel(a, golden["c3s_alpha"]) < TOL_AB
    assert rel(b, golden["c3s_beta"]) < TOL_AB
    np.testing.assert_allclose(L.H_eigvals, golden["c3s_theta"], rtol=TOL_RITZ, atol=1e-12)


def test_deuteron_converged_ritz(lz, golden):
    """3Ddeuteron.py recipe at N = 16, n = 120: the low Ritz values converge; they must match the
    reference to 1e-10 and alpha/beta to 1e-12 over the first 50 steps."""
    H, c, o, pot = orc.deuteron_hamiltonian(16)
    op = lz.StencilOperator((16, 16, 16), c, o, diag=pot)
    L = lz.Lanczos(op)
    L.execute_Lanczos(120, seed=78)
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert rel(a[:50], golden["deut_alpha"][:50]) < TOL_AB
    assert rel(b[:50], golden["deut_beta"][:50]) < TOL_AB
    th, ref = L.H_eigvals, golden["deut_theta"]
    # converged = residual estimate small: compare the lowest 8 and highest 8
    scale = np.abs(ref).max()
    assert np.max(np.abs(th[:8] - ref[:8])) < TOL_RITZ * scale
    assert np.max(np.abs(th[-8:] - ref[-8:])) < TOL_RITZ * scale
    # runtime self-checks of the reference (Lanczos.py:157-158) ran inside get_H_eigs
    Y = L.H_eigvecs
    assert Y.shape == (16 ** 3, 120)
    # Ritz vectors against the oracle's lift
    res = orc.lanczos(H, 120, seed=78, vectors=True)
    for i in (0, 1, 2):
        yo = res["Y"][:, i]
        assert min(np.linalg.norm(Y[:, i] - yo), np.linalg.norm(Y[:, i] + yo)) < 1e-8


def test_irregular_delaunay(lz, golden):
    Ld = orc.delaunay_graph_laplacian(3000, seed=0)
    for H in (Ld, sp.csc_matrix(Ld)):
        L = lz.IrrLanczos(H)
        L.execute_LanczosOld(50, seed=99)
        a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
        assert rel(a, golden["del_alpha"]) < TOL_AB
        assert rel(b, golden["del_beta"]) < TOL_AB
    L.get_H_eigs()
    np.testing.assert_allclose(L.H_eigvals, golden["del_theta"], rtol=TOL_RITZ, atol=1e-12)


def test_sweep_form_follows_use_cuda(lz):
    """Regular.execute_Lanczos(use_cuda=True) sweeps in the form of Lanczos.py:236-238 (self term
    dropped), use_cuda=False and IrrLanczos in the 2 V[j] - sum form (:247-249): each against the
    oracle in the same form; the two forms agree to rounding."""
    grid, n = (12, 10, 9), 30
    H = orc.laplacian_csr(grid, 6.5, -1.0, periodic=True)
    op = lz.StencilOperator(grid, 6.5, -1.0)
    gpu_form = orc.lanczos(H, n, seed=4, sweep="gpu")
    cpu_form = orc.lanczos(H, n, seed=4)
    L = lz.Lanczos(op)
    L.execute_Lanczos(n, seed=4)                               # use_cuda=True
    assert rel(np.diag(L.H_eff), gpu_form["alpha"]) < TOL_AB
    assert rel(np.diag(L.H_eff, 1), gpu_form["beta"]) < TOL_AB
    with pytest.warns(RuntimeWarning):
        L.execute_Lanczos(n, seed=4, use_cuda=False)
    assert rel(np.diag(L.H_eff), cpu_form["alpha"]) < TOL_AB
    assert rel(np.diag(L.H_eff, 1), cpu_form["beta"]) < TOL_AB
    Li = lz.IrrLanczos(H)
    Li.execute_LanczosOld(n, seed=4)
    assert rel(np.diag(Li.H_eff), cpu_form["alpha"]) < TOL_AB
    assert rel(gpu_form["alpha"], cpu_form["alpha"]) < 1e-13


@pytest.mark.slow
def test_config2_full_size_vs_oracle(lz):
    """BASELINE config 2 at its stated size: graph Laplacian of a 1 M-vertex 2-D Delaunay mesh, CSR in,
    m = 200, full re-orthogonalisation in the reference's form (IrrLanczos.py:193-260), against the
    oracle on the host (1.6 GB basis).  alpha/beta of the first 50 steps <= 1e-12 relative (the
    north-star bound is stated for m <= 50), every step <= 1e-9, converged Ritz values <= 1e-10."""
    npts, m = 1_000_000, 200
    H = orc.delaunay_graph_laplacian(npts, seed=0)
    assert H.shape[0] == npts and H.indices.dtype == np.int32
    ref = orc.lanczos(H, m, seed=99, blocked=True)            # bit-identical form without (n, M) temporaries
    th, S = np.linalg.eigh(ref["T"])                           # Lanczos.py:151
    L = lz.IrrLanczos(H)
    L.execute_LanczosOld(m, seed=99)
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert rel(a[:50], ref["alpha"][:50]) < TOL_AB
    assert rel(b[:50], ref["beta"][:50]) < TOL_AB
    assert rel(a, ref["alpha"]) < 1e-9
    assert rel(b, ref["beta"]) < 1e-9
    # converged Ritz pairs of the oracle: residual |H y - theta y| <= 1e-8 |H|
    theta = np.linalg.eigvalsh(L.H_eff)
    resid = np.abs(ref["beta"][-1] * S[-1, :])                 # |beta_m s_mi|: the Lanczos residual estimate
    conv = resid < 1e-8 * np.abs(th).max()
    assert conv.sum() >= 1, "no converged Ritz value at m = 200"
    np.testing.assert_allclose(theta[conv], th[conv], rtol=TOL_RITZ, atol=1e-10 * np.abs(th).max())
    # the extreme Ritz value, converged or not, also agrees (it is a function of alpha/beta alone)
    assert abs(theta[-1] - th[-1]) <= 1e-10 * abs(th[-1])
    # and the lifted Ritz vector of the best converged pair matches the oracle's up to sign
    i = int(np.argmin(np.where(conv, resid, np.inf)))
    _, Yg = L.ritz_vectors(m)
    yg = Yg[i].cpu().numpy()
    yo = np.dot(ref["V"], S[:, i])                             # Lanczos.py:155-156, one column
    assert min(np.linalg.norm(yg - yo), np.linalg.norm(yg + yo)) < 1e-7


def test_irregular_rgg_vs_oracle(lz):
    H = orc.rgg_graph_laplacian(20000, mean_degree=13.0, seed=4)
    ref = orc.lanczos(H, 40, seed=11)
    for fmt in ("csr", "sell"):
        L = lz.IrrLanczos(H)
        L.execute_LanczosOld(40, seed=11, fmt=fmt)
        _check_run(L, ref)


def test_edge_cases(lz, golden):
    op = lz.StencilOperator((6,), 2.0, -1.0, bc="dirichlet")
    L = lz.Lanczos(op)
    with pytest.raises(ValueError, match="has not been called"):
        L.H_eff
    with pytest.raises(ValueError, match="n cannot be larger than M"):
        L.execute_Lanczos(7)
    with pytest.raises(IndexError):
        L.execute_Lanczos(1)
    L.execute_Lanczos(2, seed=3)
    np.testing.assert_allclose(L.H_eff, golden["n2_T"], rtol=1e-12, atol=1e-14)
    L.execute_Lanczos(6, seed=3)            # n == M
    np.testing.assert_allclose(L.H_eff[:5, :5], golden["nM_T"][:5, :5], rtol=1e-9, atol=1e-12)
    with pytest.raises(TypeError):
        lz.Lanczos(np.eye(4)).execute_Lanczos(2)


def test_breakdown_is_reported(lz):
    # H = 3*I: every vector is an eigenvector, the pre-step residual is exactly zero
    op = lz.StencilOperator((64,), 3.0, 0.0)
    L = lz.Lanczos(op)
    with pytest.raises(lz.LanczosBreakdown):
        L.execute_Lanczos(5, seed=1, breakdown_tol=1e-10)


def test_static_reorthogonalize(lz):
    rs = np.random.RandomState(0)
    V = rs.uniform(-1, 1, (7, 1000)) / np.sqrt(1000)
    V[5:] = 0.0
    for j in (0, 3, 4):
        Vo = V.copy()
        orc.gram_schmidt_row(Vo, j)
        Vg = V.copy()
        lz.Lanczos.reorthogonalize(Vg, j, use_cuda=False)       # Lanczos.py:247-249
        assert np.max(np.abs(Vg - Vo)) < 1e-15
        Vi = V.copy()
        lz.IrrLanczos.reorthogonalize(Vi, j)                    # IrrLanczos.py:453-455: same form on both branches
        assert np.array_equal(Vi, Vg)
        # Regular with use_cuda=True (the default) drops the self term (Lanczos.py:236-238)
        Vo2 = V.copy()
        orc.gram_schmidt_row_gpu_form(Vo2, j)
        Vg2 = V.copy()
        lz.Lanczos.reorthogonalize(Vg2, j)
        assert np.max(np.abs(Vg2 - Vo2)) < 1e-15
    # rows after j that are not zero take part as well (the reference sums over all rows)
    V2 = rs.uniform(-1, 1, (6, 515)) / np.sqrt(515)
    Vo = V2.copy()
    orc.gram_schmidt_row(Vo, 2)
    Vg = V2.copy()
    lz.Lanczos.reorthogonalize(Vg, 2, use_cuda=False)
    assert np.max(np.abs(Vg - Vo)) < 1e-15


def test_bit_reproducible(lz):
    op = lz.StencilOperator((48, 40, 20), 6.0, -1.0)
    outs = []
    for _ in range(2):
        L = lz.Lanczos(op)
        L.execute_Lanczos(25, seed=5)
        outs.append(L.H_eff.copy())
    assert np.array_equal(outs[0], outs[1])


def test_modes_agree(lz):
    """CGS2 / clean start / no-basis ring mode: same Krylov space, same recurrences."""
    H, c, o, pot = orc.deuteron_hamiltonian(12)
    op = lz.StencilOperator((12, 12, 12), c, o, diag=pot)
    base = lz.Lanczos(op)
    base.execute_Lanczos(60, seed=78)
    T0 = base.H_eff
    L = lz.Lanczos(op)
    L.execute_Lanczos(60, seed=78, cgs_passes=2)
    assert rel(np.diag(L.H_eff)[:50], np.diag(T0)[:50]) < TOL_AB
    assert rel(np.diag(L.H_eff, 1)[:50], np.diag(T0, 1)[:50]) < TOL_AB
    # clean start: q_0 = v0/|v0| -> equals the oracle loop started from that vector w/o pre-step
    v0 = orc.start_vector(12 ** 3, seed=78)
    Lc = lz.Lanczos(op)
    Lc.execute_Lanczos(20, v0=v0, ref_compat=False)
    V = Lc.V
    assert np.max(np.abs(V[:, 0] - v0)) < 1e-15
    G = V.T @ V
    assert np.max(np.abs(G - np.eye(20))) < 1e-12
    R = H @ V - V @ Lc.H_eff
    assert np.max(np.abs(R[:, :-1])) < 1e-10 * np.abs(T0).max()
    # no re-orthogonalisation, ring of three vectors: first steps equal the full run
    Ln = lz.Lanczos(op)
    Ln.execute_Lanczos(12, seed=78, reorth="none", keep_basis=False)
    assert rel(np.diag(Ln.H_eff)[:8], np.diag(T0)[:8]) < 1e-9


def test_selective_reorth(lz):
    """Selective re-orthogonalisation (not in the reference): far fewer sweeps, same answers."""
    H, c, o, pot = orc.deuteron_hamiltonian(16)
    op = lz.StencilOperator((16, 16, 16), c, o, diag=pot)
    full = lz.Lanczos(op)
    full.execute_Lanczos(150, seed=78)
    sel = lz.Lanczos(op)
    sel.execute_Lanczos(150, seed=78, reorth="selective", cgs_passes=2)
    assert 0 < sel.result.reorth_count < 75
    scale = np.abs(full.H_eigvals).max()
    assert np.max(np.abs(sel.ritz_values(8) - full.ritz_values(8))) < TOL_RITZ * scale
    V = sel.V
    G = np.abs(V.T @ V - np.eye(150))
    assert G.max() < 1e-6                      # semi-orthogonality (sqrt(eps) level)
    # a pure Laplacian never loses orthogonality within 60 steps: the monitor must stay quiet
    lap = lz.Lanczos(lz.StencilOperator((40, 40, 40), 6.0, -1.0))
    lap.execute_Lanczos(60, seed=1, reorth="selective", cgs_passes=2)
    assert lap.result.reorth_count <= 2
    V = lap.V
    assert np.abs(V.T @ V - np.eye(60)).max() < 1e-7


def test_large_grid_invariants(lz):
    """Size-independent properties at a size the oracle cannot hold: orthonormal basis and the
    three-term recurrence H V = V T + beta q e^T, checked with device products."""
    import torch
    n, grid = 12, (256, 256, 128)
    op = lz.StencilOperator(grid, 6.0, -1.0)
    M = op.M
    g = torch.Generator(device="cuda").manual_seed(0)
    v0 = torch.rand(M, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    L = lz.Lanczos(op)
    L.execute_Lanczos(n, v0=v0, reorth="selective", cgs_passes=2)
    res = L.result
    res.normalize_basis()
    V = res.V_dev[:, :M]
    G = V @ V.T
    assert (G - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item() < 1e-10
    T = torch.from_numpy(L.H_eff).cuda()
    dev = op.device_handle(res.ctx)
    for j in range(n - 1):
        Hv = dev.apply(V[j].contiguous())
        r = Hv - T[j, j] * V[j] - T[j, j + 1] * V[j + 1]
        if j > 0:
            r = r - T[j, j - 1] * V[j - 1]
        assert r.norm().item() < 1e-11
    # shift invariance of the periodic Laplacian: alpha/beta do not change under a cyclic shift
    v0s = torch.roll(v0.view(grid[2], grid[1], grid[0]), shifts=(3, 5, 7), dims=(0, 1, 2)).reshape(-1).contiguous()
    L2 = lz.Lanczos(op)
    L2.execute_Lanczos(n, v0=v0s, reorth="none", keep_basis=False)
    L3 = lz.Lanczos(op)
    L3.execute_Lanczos(n, v0=v0, reorth="none", keep_basis=False)
    assert rel(np.diag(L2.H_eff), np.diag(L3.H_eff)) < 1e-11


def test_print_good_eigs_matches_the_reference_formula(lz, capsys):
    """The residual diagnostic of print_good_eigs (Lanczos.py:166-185, IrrLanczos.py:331-353) runs on the
    device; its numbers equal the reference's host formula on the oracle's Ritz vectors."""
    H, c, o, pot = orc.deuteron_hamiltonian(10)
    n = 60
    ref = orc.lanczos(H, n, seed=78, vectors=True)
    want = np.array([np.dot((H @ x) / np.linalg.norm(H @ x), x) ** 2 for x in ref["Y"].T])
    for op, cls in ((lz.StencilOperator((10, 10, 10), c, o, diag=pot), lz.Lanczos), (H, lz.Lanczos), (sp_csc(H), lz.IrrLanczos)):
        L = cls(op)
        (L.execute_Lanczos if cls is lz.Lanczos else L.execute_LanczosOld)(n, seed=78)
        got = L.print_good_eigs(tol=0.01, print_nr=5)
        out = capsys.readouterr().out
        assert "EIGENVALUE AND EIGVENVECTOR COMPARISON" in out
        assert got.shape == (n,)
        conv = np.abs(1 - want) < 1e-6                       # converged pairs: cos^2 = 1 to round-off
        assert conv.sum() >= 10
        assert np.max(np.abs(got[conv] - want[conv])) < 1e-10
        assert np.max(np.abs(got - want)) < 1e-6             # unconverged ones: same value up to the Ritz vectors' accuracy


def sp_csc(H):
    import scipy.sparse as sp
    return sp.csc_matrix(H)


def test_full_size_config3_properties(lz):
    """BASELINE config 3 at its full size (512^3 = 134 M unknowns): the recompute step (lean KA2 + KB)
    and the two-pass step (K1 + K3) give the same alpha/beta, the basis is orthonormal and satisfies the
    three-term recurrence, and alpha/beta are invariant under a cyclic shift of the start vector."""
    import torch
    n, grid = 8, (512, 512, 512)
    op = lz.StencilOperator(grid, 6.0, -1.0)
    M = op.M
    g = torch.Generator(device="cuda").manual_seed(5)
    v0 = torch.rand(M, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    T = {}
    for kern in ("two_pass", "recompute"):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, v0=v0, reorth="selective", cgs_passes=2, step_kernel=kern)
        assert L.result.step_kernel == kern
        T[kern] = L.H_eff.copy()
        if kern == "two_pass":
            del L
            torch.cuda.empty_cache()
    assert rel(np.diag(T["recompute"]), np.diag(T["two_pass"])) < 1e-12
    assert rel(np.diag(T["recompute"], 1), np.diag(T["two_pass"], 1)) < 1e-12
    res = L.result
    res.normalize_basis()
    V = res.V_dev[:, :M]
    G = V @ V.T
    assert (G - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item() < 1e-10
    Tm = torch.from_numpy(L.H_eff).cuda()
    dev = op.device_handle(res.ctx)
    for j in range(1, n - 1):
        r = dev.apply(V[j].contiguous()) - Tm[j, j] * V[j] - Tm[j, j + 1] * V[j + 1] - Tm[j, j - 1] * V[j - 1]
        assert r.norm().item() < 1e-11
    del V, G, res, L
    torch.cuda.empty_cache()
    v0s = torch.roll(v0.view(512, 512, 512), shifts=(1, 17, 64), dims=(0, 1, 2)).reshape(-1).contiguous()
    Ls = lz.Lanczos(op)
    Ls.execute_Lanczos(n, v0=v0s, reorth="none", keep_basis=False)
    assert Ls.result.step_kernel == "recompute"
    assert rel(np.diag(Ls.H_eff)[:6], np.diag(T["recompute"])[:6]) < 1e-11


# ---- KF: the single-pass fused step (structured grids, nx % 64 == 0, ny % 8 == 0) ----------------

@pytest.mark.parametrize("grid,n", [((20, 18, 16), 30), ((33, 7, 5), 60), ((64, 48, 6), 112), ((257,), 40)])
def test_cgs2_fused_middle_matches_four_sweeps(lz, grid, n):
    """K4c (update of sweep 1 + dots of sweep 2 from one staged read of the basis) against the four
    separate sweeps: v' is formed with the same operations in the same order, the second sweep's
    coefficients differ only in the order of their partial sums.  n = 112 crosses every tile
    configuration (TC = 256 up to 48 rows, TC = 128 up to ~98) and the unfused fallback beyond."""
    dim = len(grid)
    op = lz.StencilOperator(grid, 2.0 * dim + 0.3, [-1.0, -0.9, -1.1][:dim], bc="dirichlet")
    out = {}
    for fused in (True, False):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, seed=21, reorth="full", cgs_passes=2, cgs_fused=fused, profile=True)
        out[fused] = (np.diag(L.H_eff).copy(), np.diag(L.H_eff, 1).copy(), L.V.copy(), L.result.kernel_ms["gs_fused"][1])
    assert out[True][3] > 0 and out[False][3] == 0
    assert rel(out[True][0], out[False][0]) < 1e-12 and rel(out[True][1], out[False][1]) < 1e-12
    assert np.max(np.abs(out[True][2] - out[False][2])) < 1e-12
    V = out[True][2]
    assert np.max(np.abs(V.T @ V - np.eye(n))) < 5e-14


def test_cgs2_fused_middle_sparse_and_selective(lz):
    G = orc.delaunay_graph_laplacian(5001, seed=6)          # odd row count: the zero-filled tail of the last tile
    T = {}
    for fused in (True, False):
        L = lz.IrrLanczos(G)
        L.execute_LanczosOld(40, seed=2, cgs_passes=2, cgs_fused=fused)
        T[fused] = L.H_eff.copy()
    assert rel(np.diag(T[True]), np.diag(T[False])) < 1e-12 and rel(np.diag(T[True], 1), np.diag(T[False], 1)) < 1e-12
    H, c, o, pot = orc.deuteron_hamiltonian(12)
    op = lz.StencilOperator((12, 12, 12), c, o, diag=pot)
    th = {}
    for fused in (True, False):
        L = lz.Lanczos(op)
        L.execute_Lanczos(80, seed=78, reorth="selective", cgs_passes=2, cgs_fused=fused)
        assert L.result.reorth_count > 0
        th[fused] = np.linalg.eigvalsh(L.H_eff)[:4]
    assert rel(th[True], th[False]) < 1e-10


RECOMPUTE_CASES = [((7,), "periodic"), ((9,), "dirichlet"), ((6, 5), "periodic"), ((70, 3), "dirichlet"),
                   ((4, 3, 5), "periodic"), ((5, 3, 2), "dirichlet"), ((2, 2, 2), "periodic"), ((66, 9, 3), "periodic"),
                   ((130, 17, 4), "dirichlet"), ((64, 16, 12), "periodic"), ((128, 32, 5), "dirichlet"),
                   ((64, 48, 1), "periodic"), ((64, 16, 2), "dirichlet")]      # whole tiles: the lean KA2 kernel


@pytest.mark.parametrize("grid,bc", RECOMPUTE_CASES)
@pytest.mark.parametrize("reorth", ["full", "none"])
def test_recompute_step_matches_two_pass_and_oracle(lz, grid, bc, reorth):
    """KA + KB (H v re-evaluated inside the update, 32 M bytes) against K1 + K3 (48 M bytes): the
    arithmetic of w and of the update is the same operation for operation, so alpha agrees to the
    last bits and beta up to the order of its partial sums; both against the oracle."""
    dim = len(grid)
    off = [-1.0, -0.7, -1.2][:dim]
    op = lz.StencilOperator(grid, 2.0 * dim + 0.5, off, bc=bc)
    M = int(np.prod(grid))
    n = min(16, M)
    if n < 2:
        pytest.skip("n >= 2")
    runs = {}
    for kern in ("two_pass", "recompute", "auto"):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, seed=13, reorth=reorth, keep_basis=True, step_kernel=kern, persistent=False)
        runs[kern] = (np.diag(L.H_eff).copy(), np.diag(L.H_eff, 1).copy(), L.V.copy(), L.result.step_kernel)
    assert runs["two_pass"][3] == "two_pass" and runs["recompute"][3] == "recompute"
    assert runs["auto"][3] == "recompute"                 # the default for matrix-free operators
    a2, b2, V2, _ = runs["two_pass"]
    ar, br, Vr, _ = runs["recompute"]
    k = n if reorth == "full" else min(n, 6)              # without sweeps round-off differences grow
    if M < 64:
        k = min(k, 2)                                     # tiny periodic grids exhaust their Krylov space at once
    assert rel(ar[:k], a2[:k]) < 1e-12 and rel(br[:k], b2[:k]) < 1e-12
    assert np.array_equal(runs["auto"][0], ar) and np.array_equal(runs["auto"][1], br)
    if reorth == "full" and M >= 64:
        H = orc.laplacian_csr(grid, 2.0 * dim + 0.5, off, periodic=(bc == "periodic"))
        ref = orc.lanczos(H, n, seed=13)
        assert rel(ar, ref["alpha"]) < 1e-12 and rel(br, ref["beta"]) < 1e-12
        assert np.max(np.abs(Vr - ref["V"])) < 1e-12


ALPHA_CASES = [((64, 16, 12), "periodic", False), ((128, 32, 5), "dirichlet", False), ((64, 8, 1), "periodic", False),
               ((64, 8, 2), "periodic", True), ((192, 24, 40), "periodic", True), ((64, 64, 33), "dirichlet", True),
               ((256, 64, 70), "periodic", False)]


@pytest.mark.parametrize("grid,bc,pot", ALPHA_CASES)
def test_alpha_accumulated_inside_update(lz, grid, bc, pot):
    """Recompute step on whole 64 x 8 tiles: alpha_{j+1} comes from KB (edges inside a CTA tile) + the
    border kernel (edges across tiles, z-chunks and the periodic wrap) instead of a KA pass; against
    the KA form and against the oracle."""
    M = int(np.prod(grid))
    diag = np.cos(np.arange(M) * 0.37) * 0.3 if pot else None
    off = [-1.0, -0.7, -1.2]
    op = lz.StencilOperator(grid, 6.5, off, bc=bc, diag=diag)
    H = orc.laplacian_csr(grid, 6.5, off, periodic=(bc == "periodic"), diag=diag)
    n = 14
    ref = orc.lanczos(H, n, seed=5, reorth=False)
    res = {}
    for flag in (True, False):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, seed=5, reorth="none", kb_alpha=flag, profile=True)
        assert L.result.step_kernel == "recompute" and L.result.alpha_in_update == flag
        assert (L.result.kernel_ms["border"][1] > 0) == flag
        res[flag] = (np.diag(L.H_eff).copy(), np.diag(L.H_eff, 1).copy())
    k = 8                                                       # without sweeps round-off differences grow
    assert rel(res[True][0][:k], res[False][0][:k]) < 1e-12 and rel(res[True][1][:k], res[False][1][:k]) < 1e-12
    assert rel(res[True][0][:k], ref["alpha"][:k]) < 1e-12 and rel(res[True][1][:k], ref["beta"][:k]) < 1e-12
    # without a basis (three-row ring), and with selective sweeps that fire (alpha re-taken after the sweep)
    R = lz.Lanczos(op)
    R.execute_Lanczos(n, seed=5, reorth="none", keep_basis=False, kb_alpha=True)
    assert R.result.alpha_in_update and rel(np.diag(R.H_eff)[:k], res[True][0][:k]) < 1e-12
    full = orc.lanczos(H, 30, seed=5)
    S = lz.Lanczos(op)
    S.execute_Lanczos(30, seed=5, reorth="selective", cgs_passes=2, select_tol=1e-15, kb_alpha=True)      # fires at once: ~ full CGS2
    assert S.result.alpha_in_update and S.result.reorth_count >= 25
    assert rel(np.diag(S.H_eff), full["alpha"]) < 1e-11 and rel(np.diag(S.H_eff, 1), full["beta"]) < 1e-11


def test_recompute_step_with_potential_27pt_and_ring(lz):
    H, c, o, pot = orc.deuteron_hamiltonian(12)
    n = 40
    ref = orc.lanczos(H, n, seed=78)
    op = lz.StencilOperator((12, 12, 12), c, o, diag=pot)
    for kern in ("recompute", "two_pass"):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, seed=78, step_kernel=kern)
        assert rel(np.diag(L.H_eff), ref["alpha"]) < 1e-12 and rel(np.diag(L.H_eff, 1), ref["beta"]) < 1e-12
    # selective sweeps and the three-row ring
    a = {}
    for kern in ("recompute", "two_pass"):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, seed=78, reorth="selective", cgs_passes=2, step_kernel=kern)
        a[kern] = np.linalg.eigvalsh(L.H_eff)[:3]
        R = lz.Lanczos(op)
        R.execute_Lanczos(10, seed=78, reorth="none", keep_basis=False, step_kernel=kern)
        a[kern + "_ring"] = np.diag(R.H_eff).copy()
    assert rel(a["recompute"], a["two_pass"]) < 1e-9
    assert rel(a["recompute_ring"][:6], a["two_pass_ring"][:6]) < 1e-11
    # the reference's 27-point Laplacian
    w = lz.reference_T27_weights(-1.0)
    grid = (10, 9, 8)
    op27 = lz.StencilOperator(grid, 0.0, 0.0, weights27=w, diag=np.linspace(0.0, 1.0, 720))
    H27 = orc.laplacian27_csr(grid, w, periodic=True, diag=np.linspace(0.0, 1.0, 720))
    ref27 = orc.lanczos(H27, 20, seed=5)
    for kern in ("recompute", "two_pass"):
        L = lz.Lanczos(op27)
        L.execute_Lanczos(20, seed=5, step_kernel=kern)
        assert L.result.step_kernel == kern
        assert rel(np.diag(L.H_eff), ref27["alpha"]) < 1e-12 and rel(np.diag(L.H_eff, 1), ref27["beta"]) < 1e-12
    L = lz.Lanczos(op27)
    L.execute_Lanczos(20, seed=5)
    assert L.result.step_kernel == "two_pass"        # auto: two applies of the 27-point kernel do not pay


def test_recompute_step_whole_tiles_with_potential(lz):
    """The lean KA2 kernel (nx % 64 == 0, ny % 16 == 0) with a potential array, both boundary types."""
    grid = (64, 32, 6)
    pot = np.random.RandomState(8).uniform(0.0, 2.0, int(np.prod(grid)))
    for bc in ("periodic", "dirichlet"):
        op = lz.StencilOperator(grid, 6.2, [-1.0, -0.7, -1.3], bc=bc, diag=pot)
        H = orc.laplacian_csr(grid, 6.2, [-1.0, -0.7, -1.3], periodic=(bc == "periodic"), diag=pot)
        ref = orc.lanczos(H, 16, seed=3)
        for kern in ("recompute", "two_pass"):
            L = lz.Lanczos(op)
            L.execute_Lanczos(16, seed=3, step_kernel=kern)
            assert rel(np.diag(L.H_eff), ref["alpha"]) < 1e-12 and rel(np.diag(L.H_eff, 1), ref["beta"]) < 1e-12


def test_recompute_step_rejected_for_stored_operators(lz):
    H = orc.laplacian_csr((8, 8, 8), 6.0, -1.0)
    L = lz.Lanczos(H)
    with pytest.raises(RuntimeError):
        L.execute_Lanczos(5, step_kernel="recompute")
    L.execute_Lanczos(5, step_kernel="auto")
    assert L.result.step_kernel == "two_pass"


@pytest.mark.parametrize("grid,bc", [((64, 16, 12), "periodic"), ((128, 8, 9), "dirichlet"), ((64, 24, 2), "periodic")])
def test_fused_step_matches_two_pass_and_oracle(lz, grid, bc):
    op = lz.StencilOperator(grid, 6.5, [-1.0, -0.7, -1.2], bc=bc)
    H = orc.laplacian_csr(grid, 6.5, [-1.0, -0.7, -1.2], periodic=(bc == "periodic"))
    n = 20
    ref = orc.lanczos(H, n, seed=13)                  # full reorth; a Laplacian stays orthogonal anyway
    runs = {}
    for kern in ("two_pass", "fused"):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, seed=13, reorth="none", keep_basis=True, step_kernel=kern)
        runs[kern] = (np.diag(L.H_eff).copy(), np.diag(L.H_eff, 1).copy(), L.V.copy(), L.result.kernel_ms)
    a2, b2, V2, _ = runs["two_pass"]
    af, bf, Vf, _ = runs["fused"]
    assert rel(af, a2) < 1e-12 and rel(bf, b2) < 1e-12
    assert rel(af, ref["alpha"]) < 1e-10 and rel(bf, ref["beta"]) < 1e-10
    assert np.max(np.abs(Vf - V2)) < 1e-12
    # ring mode (no basis kept)
    L = lz.Lanczos(op)
    L.execute_Lanczos(n, seed=13, reorth="none", keep_basis=False, step_kernel="fused")
    assert rel(np.diag(L.H_eff), af) < 1e-14


def test_fused_step_with_potential_and_selective_sweeps(lz):
    """Deuteron-like operator on a grid the fused kernel accepts: Ritz values converge, the omega
    monitor fires, and the predicated sweep + recompute path of the fused loop is exercised."""
    grid = (64, 16, 16)
    g = [np.linspace(-12.5, 12.5, m) for m in grid]
    Z, Y, X = np.meshgrid(g[2], g[1], g[0], indexing="ij")
    pot = orc.deuteron_potential(X, Y, Z).ravel()
    dx = 25.0 / 16
    T = 197.327 ** 2 / (2 * 469.4592) / dx ** 2
    H = orc.laplacian_csr(grid, 6.0 * T, -T, periodic=True, diag=pot)
    op = lz.StencilOperator(grid, 6.0 * T, -T, diag=pot)
    n = 140
    ref = orc.lanczos(H, n, seed=78)
    sel = lz.Lanczos(op)
    sel.execute_Lanczos(n, seed=78, reorth="selective", cgs_passes=2, step_kernel="fused")
    assert 0 < sel.result.reorth_count < n // 2
    scale = np.abs(ref["theta"]).max()
    assert np.max(np.abs(sel.ritz_values(6) - ref["theta"][:6])) < TOL_RITZ * scale
    two = lz.Lanczos(op)
    two.execute_Lanczos(n, seed=78, reorth="selective", cgs_passes=2, step_kernel="two_pass")
    assert np.max(np.abs(two.ritz_values(6) - ref["theta"][:6])) < TOL_RITZ * scale
    V = sel.V
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-6
    # first 30 steps (before any sweep) agree with the reference to round-off
    assert rel(np.diag(sel.H_eff)[:30], ref["alpha"][:30]) < 1e-11


def test_fused_step_rejected_where_it_does_not_apply(lz):
    L = lz.Lanczos(lz.StencilOperator((20, 20, 20), 6.0, -1.0))
    with pytest.raises(RuntimeError):
        L.execute_Lanczos(5, reorth="none", step_kernel="fused")
    L.execute_Lanczos(5, reorth="none", step_kernel="auto")     # falls back to the two-pass step


# ---- K1b: the reference's default 27-point Laplacian, matrix-free ------------------------------------

@pytest.mark.parametrize("grid,bc", [((5, 5, 5), "periodic"), ((2, 2, 2), "periodic"), ((6, 5, 4), "dirichlet"),
                                      ((66, 9, 3), "periodic"), ((7, 3, 1), "periodic")])
def test_stencil27_pattern_and_values(lz, golden, grid, bc):
    w = orc.box27_weights(1.75)
    A = orc.laplacian27_csr(grid, w, periodic=(bc == "periodic"))
    op = lz.StencilOperator(grid, 0.0, 0.0, bc=bc, weights27=w)
    E = op.tocsr()
    assert np.array_equal(E.indptr, A.indptr)
    assert np.array_equal(E.indices, A.indices)
    np.testing.assert_allclose(E.data, A.data, rtol=1e-15)
    if grid == (5, 5, 5):
        assert np.array_equal(E.indices, golden["T27_N5_indices"])       # == Hamiltonian.create_sparse_T("27")
        np.testing.assert_allclose(E.data, golden["T27_N5_data"], rtol=1e-15)
    M = A.shape[0]
    D = A.toarray()
    rs = np.random.RandomState(1)
    for i in rs.choice(M, size=min(M, 40), replace=False):          # unit-vector probes of the kernel
        e = np.zeros(M)
        e[i] = 1.0
        np.testing.assert_allclose(op.matvec(e), D[:, i], rtol=1e-15, atol=1e-16)
    x = rs.uniform(-1, 1, M)
    ref = A * x
    assert np.max(np.abs(op.matvec(x) - ref)) <= 1e-14 * np.max(np.abs(ref))


def test_deuteron27_driver_operator(lz, golden):
    """3Ddeuteron.py's actual operator (27-point T, H = -T + V) at N = 12: matrix-free run vs the
    reference's alpha/beta, and the same operator handed over as the scipy matrix."""
    N = 12
    dx = 25.0 / N
    Tf = 197.327 ** 2 / (2 * 469.4592) * 1 / dx ** 2
    g = np.linspace(-12.5, 12.5, N)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    pot = orc.deuteron_potential(X, Y, Z).ravel()
    w = tuple(-v for v in lz.reference_T27_weights(Tf))
    op = lz.StencilOperator((N, N, N), 0.0, 0.0, weights27=w, diag=pot)
    L = lz.Lanczos(op)
    L.execute_Lanczos(60, seed=78)
    a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
    assert rel(a[:50], golden["deut27_alpha"][:50]) < TOL_AB
    assert rel(b[:50], golden["deut27_beta"][:50]) < TOL_AB
    Hs = orc.laplacian27_csr((N, N, N), w, periodic=True, diag=pot)
    L2 = lz.Lanczos(Hs)
    L2.execute_Lanczos(60, seed=78)
    assert rel(np.diag(L2.H_eff)[:50], golden["deut27_alpha"][:50]) < TOL_AB


GUARD = 4096           # doubles on either side of every buffer the library writes


@pytest.mark.parametrize("case", ["stencil_full_cgs2", "stencil_selective_kb_alpha", "sparse_sell", "team", "lift"])
def test_writes_stay_inside_the_callers_buffers(lz, case):
    """compute-sanitizer is closed on the GPU pool (it refuses to start), so out-of-bounds writes are
    hunted with guard bands instead: every buffer the C ABI writes (basis rows incl. the TMA-staged K4c
    tiles and the in-place sweeps, Ritz vectors, apply outputs) sits between bands of a sentinel that
    must survive, and the results must still match the oracle.  Sizes are odd on purpose (ragged last
    tiles / chunks / planes)."""
    import ctypes as C
    import torch
    from lanczos_b200 import _capi, engine
    sentinel = -7.0e300
    dev = torch.device("cuda", torch.cuda.current_device())

    def guarded(rows, ld):
        buf = torch.full((rows * ld + 2 * GUARD,), sentinel, dtype=torch.float64, device=dev)
        return buf, buf[GUARD:GUARD + rows * ld].view(rows, ld)

    def intact(buf, rows, ld):
        torch.cuda.synchronize()
        return bool((buf[:GUARD] == sentinel).all()) and bool((buf[GUARD + rows * ld:] == sentinel).all())

    if case in ("stencil_full_cgs2", "stencil_selective_kb_alpha", "lift"):
        grid = (64, 24, 7) if case != "stencil_full_cgs2" else (66, 9, 5)
        H = orc.laplacian_csr(grid, 6.0, -1.0, periodic=True)
        op = lz.StencilOperator(grid, 6.0, -1.0)
        ctx = engine.Context.default()
        dop = op.device_handle(ctx)
        M, n = op.M, 20
        ld = engine.padded_ld(M)
        buf, V = guarded(n, ld)
        kw = dict(reorth="full", cgs_passes=2) if case != "stencil_selective_kb_alpha" else \
            dict(reorth="selective", cgs_passes=2, select_tol=1e-14, kb_alpha=True)
        v0 = orc.start_vector(M, seed=3)
        res = engine.run_lanczos(dop, v0, n, V_dev=V, **kw)
        assert intact(buf, n, ld)
        ref = orc.lanczos(H, n, seed=3)
        assert rel(res.alpha[:10], ref["alpha"][:10]) < 1e-10
        if case == "lift":
            theta, S = np.linalg.eigh(res.tridiagonal())
            ybuf, Y = guarded(n, ld)
            _capi.check(ctx.lib.lz_ritz_vectors(ctx.handle, C.c_void_p(V.data_ptr()), ld, n, M,
                                                res.row_scale.ctypes.data_as(C.c_void_p),
                                                np.asfortranarray(S).ctypes.data_as(C.c_void_p), n,
                                                C.c_void_p(Y.data_ptr()), ld))
            assert intact(ybuf, n, ld) and intact(buf, n, ld)
            abuf, y = guarded(1, ld)
            _capi.check(ctx.lib.lz_op_apply(dop.handle, C.c_void_p(Y[0].data_ptr()), C.c_void_p(y.data_ptr())))
            assert intact(abuf, 1, ld)
            assert bool((y[0, M:] == sentinel).all())           # the pad of the row is not written either
    elif case == "sparse_sell":
        G = orc.delaunay_graph_laplacian(5003, seed=1)
        ctx = engine.Context.default()
        dop = engine.as_device_operator(G, ctx)
        M, n = G.shape[0], 24
        ld = engine.padded_ld(M)
        buf, V = guarded(n, ld)
        res = engine.run_lanczos(dop, orc.start_vector(M, seed=3), n, V_dev=V, reorth="full", cgs_passes=2)
        assert intact(buf, n, ld)
        assert rel(res.alpha[:10], orc.lanczos(G, n, seed=3)["alpha"][:10]) < 1e-10
    else:
        from lanczos_b200.team import LocalTeamLanczos
        grid = (64, 8, 9)
        op = lz.StencilOperator(grid, 6.0, -1.0)
        t = LocalTeamLanczos(op, 3)
        t.execute_Lanczos(16, seed=3, reorth="selective", cgs_passes=2, select_tol=1e-14, kb_alpha=True)
        ref = orc.lanczos(orc.laplacian_csr(grid, 6.0, -1.0, periodic=True), 16, seed=3)
        assert rel(np.diag(t.H_eff)[:10], ref["alpha"][:10]) < 1e-10


@pytest.mark.parametrize("M,n,k", [(1000, 60, 60), (777, 130, 70), (256 * 3 + 1, 5, 1), (64, 2, 2), (5003, 129, 64), (20001, 33, 17)])
def test_ritz_lift_gemm_against_numpy(lz, M, n, k):
    """K5 (lz_ritz_vectors, the lift loop of get_H_eigs, Lanczos.py:154-156) as a tall-skinny GEMM: row blocks
    of 128 (accumulating passes), column blocks of 64, ragged last tile, odd M - against V.T @ (scale * S)."""
    import ctypes as C
    import torch
    from lanczos_b200 import _capi, engine
    ctx = engine.Context.default()
    rs = np.random.RandomState(M + n)
    ld = engine.padded_ld(M)
    Vh = rs.uniform(-1, 1, (n, M))
    S = np.asfortranarray(rs.uniform(-1, 1, (n, k)))
    scale = rs.uniform(0.5, 2.0, n)
    V = torch.zeros((n, ld), dtype=torch.float64, device=ctx.torch_device)
    V[:, :M] = torch.from_numpy(Vh).to(ctx.torch_device)
    Y = torch.full((k, ld), -3.0, dtype=torch.float64, device=ctx.torch_device)
    torch.cuda.synchronize()
    _capi.check(ctx.lib.lz_ritz_vectors(ctx.handle, C.c_void_p(V.data_ptr()), ld, n, M, scale.ctypes.data_as(C.c_void_p),
                                        S.ctypes.data_as(C.c_void_p), k, C.c_void_p(Y.data_ptr()), ld))
    want = (S * scale[:, None]).T @ Vh
    got = Y[:, :M].cpu().numpy()
    assert np.max(np.abs(got - want)) < 1e-12 * n
    assert bool((Y[:, M:] == -3.0).all())                     # the pad of every row is left alone


@pytest.mark.parametrize("kind", ["stencil_full", "stencil_selective", "sparse_full"])
def test_graph_replay_is_bit_identical(lz, kind):
    """Launch-bound solves are captured into a CUDA graph the second time the same solve is asked for and
    replayed afterwards (lz_run_info.graph): same bits as the plain launches, and a change of any baked-in
    argument (start vector buffer, step count) falls back to plain launches."""
    import torch
    from lanczos_b200 import engine
    ctx = engine.Context.default()
    if kind.startswith("stencil"):
        grid = (200, 200)
        H = orc.laplacian_csr(grid, 4.0, -1.0, periodic=False)
        dop = lz.StencilOperator(grid, 4.0, -1.0, bc="dirichlet").device_handle(ctx)
    else:
        H = orc.delaunay_graph_laplacian(20000, seed=4)
        dop = engine.as_device_operator(H, ctx)
    M, n = H.shape[0], 40
    kw = dict(reorth="selective", cgs_passes=2, select_tol=1e-12) if kind == "stencil_selective" else dict(reorth="full")
    v0 = torch.from_numpy(orc.start_vector(M, seed=9)).to(ctx.torch_device)
    V = torch.empty((n, engine.padded_ld(M)), dtype=torch.float64, device=ctx.torch_device)
    runs = []
    for _ in range(4):
        V.fill_(float("nan"))
        res = engine.run_lanczos(dop, v0, n, V_dev=V, persistent=False, **kw)
        runs.append((res.graph, res.alpha.copy(), res.beta.copy(), V[:, :M].clone(), res.launches))
    assert [r[0] for r in runs] == ["none", "captured", "replayed", "replayed"]
    for g, a, b, Vc, nl in runs[1:]:
        assert np.array_equal(a, runs[0][1]) and np.array_equal(b, runs[0][2]) and nl == runs[0][4]
        assert torch.equal(Vc, runs[0][3])
    ref = orc.lanczos(H, n, seed=9)
    if kind != "stencil_selective":
        assert rel(runs[2][1], ref["alpha"]) < TOL_AB and rel(runs[2][2], ref["beta"]) < TOL_AB
    # another start-vector buffer: not the captured solve
    res = engine.run_lanczos(dop, v0.clone(), n, V_dev=V, persistent=False, **kw)
    assert res.graph == "none" and np.array_equal(res.alpha, runs[0][1])
    res = engine.run_lanczos(dop, v0, n - 1, V_dev=V, persistent=False, **kw)
    assert res.graph == "none" and np.array_equal(res.alpha, runs[0][1][:n - 1])


KBA_CASES = [((64, 16, 12), "periodic", False), ((128, 32, 5), "dirichlet", False), ((64, 16, 2), "periodic", True),
             ((192, 48, 40), "periodic", True), ((64, 64, 33), "dirichlet", True), ((256, 64, 70), "periodic", False)]


@pytest.mark.parametrize("grid,bc,pot", KBA_CASES)
def test_kba_step_matches_ka_kb(lz, grid, bc, pot):
    """One kernel per step (KBA: KB plus the alpha reduction of its output, read back through L2 behind
    per-chunk completion counters) against KA2 + KB: same arithmetic per point, so alpha/beta agree to
    the order of the partial sums; and against the oracle."""
    M = int(np.prod(grid))
    diag = np.cos(np.arange(M) * 0.37) * 0.3 if pot else None
    off = [-1.0, -0.7, -1.2]
    op = lz.StencilOperator(grid, 6.5, off, bc=bc, diag=diag)
    H = orc.laplacian_csr(grid, 6.5, off, periodic=(bc == "periodic"), diag=diag)
    n = 14
    ref = orc.lanczos(H, n, seed=5, reorth=False)
    res = {}
    for flag in (True, False):
        L = lz.Lanczos(op)
        L.execute_Lanczos(n, seed=5, reorth="none", kba=flag, persistent=False)
        assert L.result.step_kernel == "recompute" and L.result.kba == flag
        res[flag] = (np.diag(L.H_eff).copy(), np.diag(L.H_eff, 1).copy(), L.V.copy(), L.result.launches)
        L.execute_Lanczos(n, seed=5, reorth="none", kba=flag, persistent=False)
        assert np.array_equal(np.diag(L.H_eff), res[flag][0])           # static item -> CTA deal: bit-reproducible
    assert res[True][3] < res[False][3]                                  # one launch per step instead of two
    k = 8
    assert rel(res[True][0][:k], res[False][0][:k]) < 1e-12 and rel(res[True][1][:k], res[False][1][:k]) < 1e-12
    assert rel(res[True][0][:k], ref["alpha"][:k]) < 1e-12 and rel(res[True][1][:k], ref["beta"][:k]) < 1e-12
    assert np.max(np.abs(res[True][2][:, :3] - res[False][2][:, :3])) < 1e-13
    R = lz.Lanczos(op)
    R.execute_Lanczos(n, seed=5, reorth="none", keep_basis=False, kba=True)       # three-row ring
    assert R.result.kba and rel(np.diag(R.H_eff)[:k], res[True][0][:k]) < 1e-12
    full = orc.lanczos(H, 30, seed=5)
    S = lz.Lanczos(op)
    S.execute_Lanczos(30, seed=5, reorth="selective", cgs_passes=2, select_tol=1e-15, kba=True)      # sweeps fire: alpha re-taken by KA2
    assert S.result.kba and S.result.reorth_count >= 25
    assert rel(np.diag(S.H_eff), full["alpha"]) < 1e-11 and rel(np.diag(S.H_eff, 1), full["beta"]) < 1e-11


SMALL_CASES = [((200, 200), "dirichlet", False), ((200, 200), "periodic", False), ((1001,), "dirichlet", False),
               ((7,), "periodic", False), ((6, 5), "periodic", True), ((70, 3), "dirichlet", False), ((33, 20, 7), "periodic", True),
               ((64, 64, 64), "periodic", False), ((2, 2, 2), "periodic", False), ((130, 17, 4), "dirichlet", True)]


@pytest.mark.parametrize("grid,bc,pot", SMALL_CASES)
def test_persistent_small_problem_kernel(lz, grid, bc, pot):
    """The whole solve of a small matrix-free problem in one persistent cooperative kernel (small.cu, opt-in
    `persistent=True`; step_kernel "persistent", one launch): against the oracle in both sweep forms, with CGS2, without sweeps, from a clean
    start, and against the kernel-per-phase loop."""
    dim = len(grid)
    M = int(np.prod(grid))
    off = [-1.0, -0.7, -1.2][:dim]
    diag = np.sin(np.arange(M) * 0.21) * 0.4 if pot else None
    op = lz.StencilOperator(grid, 2.0 * dim + 0.5, off, bc=bc, diag=diag)
    H = orc.laplacian_csr(grid, 2.0 * dim + 0.5, off, periodic=(bc == "periodic"), diag=diag)
    n = min(24, M)
    if n < 2:
        pytest.skip("n >= 2")
    tight = M >= 64                                        # tiny periodic grids exhaust their Krylov space at once
    k = n if tight else 2
    L = lz.Lanczos(op)
    for use_cuda, sweep in ((True, "gpu"), (False, "cpu")):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            L.execute_Lanczos(n, seed=13, use_cuda=use_cuda, persistent=True)
        assert L.result.step_kernel == "persistent" and L.result.launches == 1
        ref = orc.lanczos(H, n, seed=13, sweep=sweep)
        a, b = np.diag(L.H_eff), np.diag(L.H_eff, 1)
        assert rel(a[:k], ref["alpha"][:k]) < TOL_AB and rel(b[:max(k - 1, 1)], ref["beta"][:max(k - 1, 1)]) < TOL_AB
        if tight:
            assert np.max(np.abs(L.V - ref["V"])) < 1e-11
    if tight:
        L.get_H_eigs()
        np.testing.assert_allclose(L.H_eigvals, ref["theta"], rtol=TOL_RITZ, atol=1e-12)
    # the kernel-per-phase loop gives the same tridiagonal matrix
    P = lz.Lanczos(op)
    P.execute_Lanczos(n, seed=13, persistent=False)
    assert P.result.step_kernel != "persistent"
    Q = lz.Lanczos(op)
    Q.execute_Lanczos(n, seed=13, persistent=True)
    assert rel(np.diag(Q.H_eff)[:k], np.diag(P.H_eff)[:k]) < TOL_AB
    Q.execute_Lanczos(n, seed=13, persistent=True)
    T1 = Q.H_eff.copy()
    Q.execute_Lanczos(n, seed=13, persistent=True)
    assert np.array_equal(T1, Q.H_eff)                     # bit-reproducible
    if tight:
        # CGS2, no sweeps, clean start
        Q.execute_Lanczos(n, seed=13, cgs_passes=2, persistent=True)
        P.execute_Lanczos(n, seed=13, cgs_passes=2, persistent=False)
        assert Q.result.step_kernel == "persistent" and rel(np.diag(Q.H_eff), np.diag(P.H_eff)) < 1e-11
        Vq = Q.V
        assert np.max(np.abs(Vq.T @ Vq - np.eye(n))) < 1e-13
        Q.execute_Lanczos(8, seed=13, reorth="none", persistent=True)
        r0 = orc.lanczos(H, 8, seed=13, reorth=False)
        assert Q.result.step_kernel == "persistent" and rel(np.diag(Q.H_eff)[:6], r0["alpha"][:6]) < 1e-11
        v0 = orc.start_vector(M, seed=5)
        Q.execute_Lanczos(10, v0=v0, ref_compat=False, persistent=True)
        P.execute_Lanczos(10, v0=v0, ref_compat=False, persistent=False)
        assert Q.result.step_kernel == "persistent" and rel(np.diag(Q.H_eff), np.diag(P.H_eff)) < 1e-11
        assert np.max(np.abs(Q.V[:, 0] - v0 / np.linalg.norm(v0))) < 1e-15


def test_persistent_kernel_reports_breakdown(lz):
    op = lz.StencilOperator((4, 4, 4), 6.0, -1.0)         # 64 unknowns, few distinct eigenvalues
    L = lz.Lanczos(op)
    with pytest.raises(lz.LanczosBreakdown) as e:
        L.execute_Lanczos(40, seed=1, breakdown_tol=1e-10, persistent=True)
    assert 0 < e.value.steps_done < 40


def test_value_free_sell_for_unweighted_graph_laplacians(lz):
    """SELL operators whose off-diagonal entries are all equal (L = D - A of an unweighted graph: BASELINE configs 2
    and 4) are applied from their column indices alone (4 instead of 12 bytes per stored entry); weighted operators,
    and fmt="sell_values", keep the general kernel.  Same y, same exported matrix, same alpha/beta."""
    import torch
    from lanczos_b200 import engine
    ctx = engine.Context.default()
    G = orc.delaunay_graph_laplacian(7003, seed=2)                      # ragged rows, odd row count, padding in every chunk
    x = np.random.RandomState(0).uniform(-1, 1, G.shape[0])
    ops = {f: engine.DeviceOperator.from_scipy(ctx, G, fmt=f) for f in ("sell", "sell_values", "csr")}
    assert ops["sell"].value_free() and not ops["sell_values"].value_free() and not ops["csr"].value_free()
    ref = G * x
    for f, op in ops.items():
        y = op.apply_host(x)
        assert np.max(np.abs(y - ref)) <= 4e-15 * np.max(np.abs(ref)), f
        assert (op.export_csr() != G).nnz == 0, f
    # -2 on every edge (a scaled Laplacian) is uniform too; a weighted one is not; a diagonal-free adjacency matrix is
    assert engine.DeviceOperator.from_scipy(ctx, sp.csr_matrix(2.0 * G), fmt="sell").value_free()
    W = sp.csr_matrix(G.copy())
    W.data = W.data * (1.0 + 0.01 * np.arange(W.nnz))
    wop = engine.DeviceOperator.from_scipy(ctx, W, fmt="sell")
    assert not wop.value_free()
    assert np.max(np.abs(wop.apply_host(x) - W * x)) <= 4e-15 * np.max(np.abs(W * x))
    A = sp.csr_matrix(sp.diags(G.diagonal()) - G)                       # adjacency: no stored diagonal, entries +1
    A.eliminate_zeros()
    aop = engine.DeviceOperator.from_scipy(ctx, A, fmt="sell")
    assert aop.value_free() and np.max(np.abs(aop.apply_host(x) - A * x)) <= 4e-15 * np.max(np.abs(A * x))
    # device-resident CSR arrays (the config-4 path) and the loop
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.torch_device).to(dt)     # noqa: E731
    dop = engine.DeviceOperator.from_device_csr(ctx, t(G.indptr, torch.int32), t(G.indices, torch.int32), t(G.data, torch.float64))
    assert dop.value_free() and np.max(np.abs(dop.apply_host(x) - ref)) <= 4e-15 * np.max(np.abs(ref))
    refl = orc.lanczos(G, 30, seed=5)
    for f in ("sell", "sell_values"):
        L = lz.IrrLanczos(G)
        L.execute_LanczosOld(30, seed=5, fmt=f)
        assert rel(np.diag(L.H_eff), refl["alpha"]) < TOL_AB and rel(np.diag(L.H_eff, 1), refl["beta"]) < TOL_AB


def test_windowed_sell_form(lz, monkeypatch):
    """csrc/sellw.cu: when the columns every sorting window of a value-free operator refers to fit a shared-memory
    stage, x is staged there and the stored column indices are 16-bit offsets.  With the entries in column order
    (LZ_SELLW_BANKS=0) the summation order per row is the plain SELL kernel's and y is bit-identical to it; by
    default a row's entries are stored in a bank-aware order (same sums up to rounding, fixed order).  Ragged rows,
    a last window that is not full, vectors that are only 8-byte aligned, both kernel variants; weighted operators
    and operators whose windows do not fit keep the plain kernel; the loop gives the oracle's alpha/beta."""
    import torch
    from lanczos_b200 import engine
    ctx = engine.Context.default()
    M = 40_037
    G = orc.banded_graph_laplacian(M, seed=3)
    x = np.random.RandomState(0).uniform(-1, 1, M)
    monkeypatch.setenv("LZ_SELL_WINDOW", "0")
    plain = engine.DeviceOperator.from_scipy(ctx, G, fmt="sell")
    assert plain.windowed() == 0 and plain.value_free()
    y_plain = plain.apply_host(x)
    monkeypatch.setenv("LZ_SELL_WINDOW", "1")
    monkeypatch.setenv("LZ_SELL_WINDOW_MIN", "1")
    buf = torch.zeros(M + 1, dtype=torch.float64, device=ctx.torch_device)      # odd rows of a basis with odd M
    buf[1:] = torch.from_numpy(x).to(ctx.torch_device)
    assert buf[1:].data_ptr() % 16 == 8
    for banks in ("1", "0"):
        monkeypatch.setenv("LZ_SELLW_BANKS", banks)
        op = engine.DeviceOperator.from_scipy(ctx, G, fmt="sell")
        assert 0 < op.windowed() <= 448 and op.value_free()
        ys = []
        for variant in ("0", "1"):          # two CTAs per SM with one stage each / one CTA with two stages (the fallback)
            monkeypatch.setenv("LZ_SELLW_VARIANT", variant)
            y = op.apply_host(x)
            assert np.max(np.abs(y - G * x)) <= 4e-15 * np.max(np.abs(G * x)), (banks, variant)
            if banks == "0":
                assert np.array_equal(y, y_plain), variant
            assert np.array_equal(op.apply(buf[1:]).cpu().numpy(), y), (banks, variant)      # no bulk copies
            ys.append(y)
        assert np.array_equal(ys[0], ys[1])
        monkeypatch.setenv("LZ_SELLW_VARIANT", "0")
        assert (op.export_csr() != G).nnz == 0
    monkeypatch.delenv("LZ_SELLW_BANKS")
    # a weighted operator keeps the plain kernel (with stored values it is already close to the HBM bound)
    W = sp.csr_matrix(G.copy())
    W.data = W.data * (1.0 + 0.01 * (np.arange(W.nnz) % 89))
    wop = engine.DeviceOperator.from_scipy(ctx, W, fmt="sell")
    assert wop.windowed() == 0 and not wop.value_free()
    assert np.max(np.abs(wop.apply_host(x) - W * x)) <= 4e-15 * np.max(np.abs(W * x))
    # smaller sorting windows (8 chunks: half of the kernel's warps idle) and one that is no multiple of 8 chunks
    for sigma in (256, 96):
        op = engine.DeviceOperator.from_scipy(ctx, G, fmt="sell", sigma=sigma)
        assert op.windowed() > 0
        assert np.max(np.abs(op.apply_host(x) - G * x)) <= 4e-15 * np.max(np.abs(G * x))
    # too many distinct runs of columns per window: the plain kernel stays
    far = orc.banded_graph_laplacian(M, far=(5000, 9000, 14000, 20000, 27000), seed=3)
    fop = engine.DeviceOperator.from_scipy(ctx, far, fmt="sell")
    assert fop.windowed() == 0
    assert np.max(np.abs(fop.apply_host(x) - far * x)) <= 4e-15 * np.max(np.abs(far * x))
    # the loop (alpha from the windowed kernel's partial sums, bookkeeping tail run by its first 8 warps)
    ref = orc.lanczos(G, 30, seed=5)
    for variant in ("0", "1"):
        monkeypatch.setenv("LZ_SELLW_VARIANT", variant)
        L = lz.IrrLanczos(G)
        L.execute_LanczosOld(30, seed=5)
        assert L._device_op.windowed() > 0
        assert rel(np.diag(L.H_eff), ref["alpha"]) < TOL_AB and rel(np.diag(L.H_eff, 1), ref["beta"]) < TOL_AB
        first = L.H_eff.copy()
        L.execute_LanczosOld(30, seed=5)
        assert np.array_equal(first, L.H_eff)
