"""Drop-in for Python/Regular/Hamiltonian.py (lanczos_b200/hamiltonian.py): the traced potential program
(CPU tests) and the matrix-free T / V / H against the golden CSR matrices of the live reference (GPU tests)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import lanczos_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def deuteron(x, y, z):
    """3Ddeuteron.py:51-61, verbatim arithmetic."""
    r = np.sqrt(x**2 + y**2 + z**2)
    eWell = 54.531
    eWells = 65.4823128982115
    eCores = 40.0*eWell
    rCore = 1.0/4
    rWell = 17.0/10
    fPow = 4.0
    return eCores*np.exp(-(r/rCore)**fPow) - eWells*np.exp(-(r/rWell)**fPow)


def run_program(ops, consts, x, y, z):
    """Host interpreter of the postfix program (what csrc/potential.cu runs per grid point)."""
    from lanczos_b200.hamiltonian import _OP
    name = {v: k for k, v in _OP.items()}
    st = []
    old = np.seterr(all="ignore")              # every branch of the dispatch tables below is evaluated
    for op in ops:
        k = name[op & 0xff]
        if k in "XYZ":
            st.append({"X": x, "Y": y, "Z": z}[k])
        elif k == "CONST":
            st.append(consts[op >> 8])
        elif k in ("ADD", "SUB", "MUL", "DIV", "POW", "MIN", "MAX"):
            b, a = st.pop(), st.pop()
            st.append({"ADD": a + b, "SUB": a - b, "MUL": a * b, "DIV": a / b, "POW": np.power(a, b),
                       "MIN": np.minimum(a, b), "MAX": np.maximum(a, b)}[k])
        else:
            a = st.pop()
            st.append({"NEG": -a, "SQRT": np.sqrt(a), "EXP": np.exp(a), "LOG": np.log(np.abs(a) + 1e-300), "ABS": np.abs(a),
                       "SIN": np.sin(a), "COS": np.cos(a), "TANH": np.tanh(a), "SQUARE": a * a}[k])
    np.seterr(**old)
    assert len(st) == 1
    return st[0]


def test_potential_is_traced_into_a_postfix_program():
    from lanczos_b200.hamiltonian import _Tracer
    ops, consts = _Tracer().trace(deuteron)
    assert len(ops) <= 96 and len(consts) <= 32
    rs = np.random.RandomState(0)
    for _ in range(50):
        x, y, z = rs.uniform(-12.5, 12.5, 3)
        want = deuteron(np.float64(x), np.float64(y), np.float64(z))
        got = run_program(ops, consts, np.float64(x), np.float64(y), np.float64(z))
        assert got == want                      # same operations in the same order: bit-identical on the host
    # other shapes of user code: reflected operands, np.power / np.abs / np.minimum, constants folded by Python
    f = lambda x, y, z: 2.0 - np.abs(x) / (1.0 + np.minimum(y * y, 3.0)) + np.power(z, 3) * 0.5 - (-x)  # noqa: E731
    ops, consts = _Tracer().trace(f)
    assert run_program(ops, consts, 0.3, -1.2, 0.7) == f(0.3, -1.2, 0.7)
    ops, consts = _Tracer().trace(lambda x, y, z: 1.5)                       # constant potential
    assert run_program(ops, consts, 0.0, 0.0, 0.0) == 1.5


def test_branching_potentials_are_not_traced():
    from lanczos_b200.hamiltonian import _Tracer, _Untraceable
    with pytest.raises(_Untraceable):
        _Tracer().trace(lambda x, y, z: 1.0 if x > 0 else 0.0)
    with pytest.raises(_Untraceable):
        _Tracer().trace(lambda x, y, z: np.where(x > 0, 1.0, 0.0))
    with pytest.raises((_Untraceable, TypeError)):
        _Tracer().trace(lambda x, y, z: np.arctan2(x, y))


def test_reference_attributes_and_helpers():
    """Attributes and helper methods of Hamiltonian.py:8-25,73-128 (no GPU needed)."""
    import lanczos_b200 as lz
    H = lz.Hamiltonian(5, 25, deuteron, 1.75)
    assert H.dx == 5.0 and np.array_equal(H.x, np.linspace(-12.5, 12.5, 5))
    w = H.get_weights_27point()
    assert w.shape == (27,) and w[13] == (-44 / 3) * 3.0 / 13 and np.isclose(w.sum(), 0.0, atol=1e-15)
    assert H.unravel_xyz(1, 2, 3) == 1 + 2 * 5 + 3 * 25 and H.ravel_i(86) == (1, 2, 3)
    nb, ww = H.Laplacian_7point(0)
    assert nb == [0, 4, 20, 100, 1, 5, 25] and list(ww) == [-6, 1, 1, 1, 1, 1, 1]
    nb27, _ = H.Laplacian_27point(0)
    assert len(nb27) == 27 and len(set(nb27)) == 27 and nb27[13] == 0


@pytest.fixture(scope="module")
def lz():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lanczos_b200
    return lanczos_b200


@pytest.mark.gpu
@pytest.mark.parametrize("N", [2, 3, 5])
def test_T_V_H_match_the_reference_matrices(lz, golden, N, tmp_path, monkeypatch):
    """create_sparse_T("7"/"27"), create_sparse_V and H = -T + V against the CSR matrices the live reference
    built for the same N, L = 25, T_factor = 1.75 (tests/golden/make_golden.py): pattern bit-exact, values of
    T bit-exact, values of H to 1e-14 (exp / pow run on the device)."""
    monkeypatch.chdir(tmp_path)                       # the T_matrices cache lives in the working directory
    sysm = lz.Hamiltonian(N, 25, deuteron, 1.75)
    for points in ("7", "27"):
        sysm.create_sparse_T(points)
        T = sysm.T_sparse.tocsr()
        assert np.array_equal(T.indptr, golden[f"T{points}_N{N}_indptr"])
        if N >= 3:                                    # N = 2: the reference's COO build leaves duplicates unsorted
            Tg = _csr(golden, f"T{points}_N{N}", N)
            assert np.array_equal(T.indices, Tg.indices) and np.array_equal(T.data, Tg.data)
    sysm.create_sparse_T("7")
    sysm.create_sparse_V()
    assert sysm.potential_evaluated_on == "device"
    H = (-sysm.T_sparse + sysm.V_sparse)
    H.sort_indices()
    assert "matrix-free" in str(H) and H.shape == (N ** 3, N ** 3)
    E = H.tocsr()
    Hg = _csr(golden, f"H_N{N}", N)
    if N >= 3:
        assert np.array_equal(E.indptr, Hg.indptr) and np.array_equal(E.indices, Hg.indices)
        np.testing.assert_allclose(E.data, Hg.data, rtol=1e-14, atol=1e-14 * np.abs(Hg.data).max())
    else:
        assert abs(E - Hg).max() < 1e-12 * abs(Hg).max()


def _csr(golden, key, N):
    import scipy.sparse as sp
    A = sp.csr_matrix((golden[key + "_data"], golden[key + "_indices"], golden[key + "_indptr"]), shape=(N ** 3, N ** 3))
    A.sum_duplicates()
    A.sort_indices()
    return A


@pytest.mark.gpu
def test_potential_on_device_and_host_fallback(lz):
    from lanczos_b200.hamiltonian import evaluate_potential
    N = 24
    g = np.linspace(-12.5, 12.5, N)
    d, how = evaluate_potential(deuteron, g, g, g)
    assert how == "device"
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    want = deuteron(X, Y, Z).ravel()
    got = d.cpu().numpy()
    assert np.max(np.abs(got - want)) <= 4e-16 * np.max(np.abs(want)) + 1e-300
    # a potential that branches cannot be traced: evaluated on the host grid, same layout
    box = lambda x, y, z: np.where(np.abs(x) + 0 * y + 0 * z < 5.0, -1.0, 0.0)     # noqa: E731
    d2, how2 = evaluate_potential(box, g, g[:7], g[:5])
    assert how2 == "host" and d2.numel() == N * 7 * 5
    assert np.array_equal(d2.cpu().numpy().reshape(5, 7, N)[2, 3], np.where(np.abs(g) < 5.0, -1.0, 0.0))
    step = lambda x, y, z: -1.0 if x * x + y * y + z * z < 30.0 else 0.0            # noqa: E731
    d3, how3 = evaluate_potential(step, g[:6], g[:5], g[:4])
    assert how3 == "host-scalar"
    assert d3.cpu().numpy()[1 + 6 * (2 + 5 * 3)] == step(g[1], g[2], g[3])


@pytest.mark.gpu
def test_deuteron_driver_end_to_end(lz, golden):
    """The body of 3Ddeuteron.py:63-97 at N = 12, n = 60 with the drop-in Hamiltonian and Lanczos classes:
    alpha/beta against the live reference's (golden deut27_*, same N, L, T_factor, seed)."""
    n, N, L = 60, 12, 25
    dx = float(L) / N
    hc, rest_energy = 197.327, 469.4592
    T_factor = hc**2/(2*rest_energy) * 1/dx**2
    system = lz.Hamiltonian(N, L, deuteron, T_factor)
    system.create_sparse_T()
    system.create_sparse_V()
    H = (-system.T_sparse + system.V_sparse)
    H.sort_indices()
    TEST = lz.Lanczos(H)
    with pytest.warns(RuntimeWarning):
        TEST.execute_Lanczos(n, use_cuda=False, seed=78)
    a, b = np.diag(TEST.H_eff), np.diag(TEST.H_eff, 1)
    assert np.max(np.abs(a[:50] - golden["deut27_alpha"][:50]) / np.abs(golden["deut27_alpha"][:50])) < 1e-12
    assert np.max(np.abs(b[:50] - golden["deut27_beta"][:50]) / np.abs(golden["deut27_beta"][:50])) < 1e-12
    l_L, v_L = TEST.H_eigvals, TEST.H_eigvecs
    assert v_L.shape == (N ** 3, n)
    np.testing.assert_allclose(l_L[:3], golden["deut27_theta"][:3], rtol=1e-10)
    TEST.print_good_eigs()


@pytest.mark.gpu
def test_reference_driver_script_runs_unchanged(lz, tmp_path):
    """`from Hamiltonian import Hamiltonian; from Lanczos import Lanczos` with Python/Regular on the path - the
    import lines of 3Ddeuteron.py:8,73 - resolve to the drop-ins, and the driver's statements run as written."""
    script = tmp_path / "driver.py"
    script.write_text('''
import sys, numpy as np
sys.path.insert(0, %r)
from Lanczos import Lanczos
def potential(x, y, z):
    r = np.sqrt(x**2 + y**2 + z**2)
    return 2181.24*np.exp(-(r/0.25)**4.0) - 65.4823128982115*np.exp(-(r/1.7)**4.0)
n = 30; N = 16; L = 25
dx = float(L)/N
T_factor = 197.327**2/(2*469.4592) * 1/dx**2
from Hamiltonian import Hamiltonian
system = Hamiltonian(N, L, potential, T_factor)
system.create_sparse_T()
system.create_sparse_V()
T_sparse = system.T_sparse
V_sparse = system.V_sparse
H = (-T_sparse + V_sparse)
H.sort_indices()
print("H MATRIX:")
print(H)
TEST = Lanczos(H)
TEST.execute_Lanczos(n, use_cuda=True, seed=78)
l_L, v_L = TEST.H_eigvals, TEST.H_eigvecs
TEST.print_good_eigs()
np.save("eigvals.npy", l_L)
print("LOWEST", l_L[0])
''' % os.path.join(ROOT, "Python", "Regular"))
    run = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "+++ Lanczos executed successfully." in run.stdout and "matrix-free" in run.stdout
    lowest = float(run.stdout.split("LOWEST")[1].split()[0])
    assert np.isfinite(lowest) and os.path.exists(tmp_path / "eigvals.npy")
    assert np.array_equal(np.sort(np.load(tmp_path / "eigvals.npy")), np.load(tmp_path / "eigvals.npy"))


@pytest.mark.gpu
def test_T_matrix_cache_round_trip(lz, golden, tmp_path, monkeypatch):
    """Hamiltonian.py:48-69: `T_matrices/T_N=<N>_Laplace=<points>.npz` is honoured when it exists (the weights are read
    back from it) and written on request in the reference's own format - a cache written here is what the
    reference's create_sparse_T would load, and vice versa."""
    import scipy.sparse as sp
    monkeypatch.chdir(tmp_path)
    N = 5
    a = lz.Hamiltonian(N, 25, deuteron, 1.75)
    a.create_sparse_T("27", save_cache=True)
    path = tmp_path / "T_matrices" / "T_N=5_Laplace=27.npz"
    assert path.exists()
    T = sp.csr_matrix(sp.load_npz(str(path)))
    T.sort_indices()
    assert np.array_equal(T.indices, golden["T27_N5_indices"]) and np.array_equal(T.data, _csr(golden, "T27_N5", N).data)
    b = lz.Hamiltonian(N, 25, deuteron, 99.0)              # another T_factor: the cached matrix wins, as in the reference
    b.create_sparse_T("27")
    assert b.T_sparse.weights == a.T_sparse.weights
    # a cache written by the reference (here: the golden 7-point matrix) is read back as a 7-point stencil
    sp.save_npz(str(tmp_path / "T_matrices" / "T_N=5_Laplace=7.npz"), _csr(golden, "T7_N5", N))
    c = lz.Hamiltonian(N, 25, deuteron, 1.0)
    c.create_sparse_T("7")
    assert c.T_sparse.weights == (1.75 * -6.0, 1.75, 0.0, 0.0) and c.T_sparse.weights27 is None
