#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference); the GPU box uses the
committed ``*.npz`` files.  The reference imports cupy / cupyx / matplotlib at module
top (Lanczos.py:3-5) - none is installed here - so they are stubbed in sys.modules
before import; only the ``use_cuda=False`` branch is ever executed.

    python tests/golden/make_golden.py

Each fixture stores the inputs that cannot be regenerated bit-for-bit elsewhere (the
operator is rebuilt from its recipe by oracle.lanczos_oracle, the start vector comes
from the legacy NumPy RNG stream) and the reference's outputs.
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("LANCZOS_REF", "/root/reference")
sys.path.insert(0, ROOT)


def _stub_modules():
    cupy = types.ModuleType("cupy")
    cupy.ndarray = type("ndarray", (), {})
    cupyx = types.ModuleType("cupyx")
    cupyx_scipy = types.ModuleType("cupyx.scipy")
    cupyx_sparse = types.ModuleType("cupyx.scipy.sparse")
    cupyx.scipy = cupyx_scipy
    cupyx_scipy.sparse = cupyx_sparse
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    for name, mod in [("cupy", cupy), ("cupyx", cupyx), ("cupyx.scipy", cupyx_scipy),
                      ("cupyx.scipy.sparse", cupyx_sparse), ("matplotlib", mpl),
                      ("matplotlib.pyplot", plt)]:
        sys.modules.setdefault(name, mod)


def load_reference():
    _stub_modules()
    sys.path.insert(0, os.path.join(REF, "Python", "Regular"))
    sys.path.insert(0, os.path.join(REF, "Python", "Irregular"))
    import Lanczos as ref_regular          # noqa: E402
    import IrrLanczos as ref_irregular     # noqa: E402
    import Hamiltonian as ref_hamiltonian  # noqa: E402
    return ref_regular, ref_irregular, ref_hamiltonian


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def run_regular(ref, H, n, seed=99, v0=None):
    with quiet():
        L = ref.Lanczos(H)
        L.execute_Lanczos(n, seed=seed, use_cuda=False, v0=v0)
        L.get_H_eigs()
    return L


def run_irregular(ref, H, n, seed=99, v0=None):
    with quiet():
        L = ref.IrrLanczos(H)
        L.execute_LanczosOld(n, seed=seed, use_cuda=False, v0=v0)
        L.get_H_eigs()
    return L


def tri_parts(T):
    return np.diag(T).copy(), np.diag(T, 1).copy()


def main():
    from oracle import lanczos_oracle as orc
    ref_reg, ref_irr, ref_ham = load_reference()
    out = {}

    # ---- G1: reference T matrix pattern (7-point periodic) at N = 5 and N = 2 --------
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                      # Hamiltonian() creates ./T_matrices
        try:
            for N in (2, 3, 5):
                with quiet():
                    ham = ref_ham.Hamiltonian(N, 25.0, orc.deuteron_potential, 1.75)
                    ham.create_sparse_T("7")
                    ham.create_sparse_V()
                T = ham.T_sparse.copy()
                T.sort_indices()
                out[f"T7_N{N}_indptr"] = T.indptr.astype(np.int32)
                out[f"T7_N{N}_indices"] = T.indices.astype(np.int32)
                out[f"T7_N{N}_data"] = T.data.astype(np.float64)
                with quiet():
                    ham27 = ref_ham.Hamiltonian(N, 25.0, orc.deuteron_potential, 1.75)
                    ham27.create_sparse_T("27")
                T27 = ham27.T_sparse.copy()
                T27.sum_duplicates()
                T27.sort_indices()
                out[f"T27_N{N}_indptr"] = T27.indptr.astype(np.int32)
                out[f"T27_N{N}_indices"] = T27.indices.astype(np.int32)
                out[f"T27_N{N}_data"] = T27.data.astype(np.float64)
                Hd = (-ham.T_sparse + ham.V_sparse)
                Hd.sort_indices()
                out[f"H_N{N}_indptr"] = Hd.indptr.astype(np.int32)
                out[f"H_N{N}_indices"] = Hd.indices.astype(np.int32)
                out[f"H_N{N}_data"] = Hd.data.astype(np.float64)
        finally:
            os.chdir(cwd)

    # ---- G2: config-1 shape, small: 2-D 5-point periodic 24x20, n = 30 --------------
    H = orc.laplacian_csr((24, 20), 4.0, -1.0, periodic=True)
    L = run_regular(ref_reg, H, 30, seed=99)
    a, b = tri_parts(L.H_eff)
    out["c1s_alpha"], out["c1s_beta"], out["c1s_theta"] = a, b, L.H_eigvals.copy()
    out["c1s_V_first3"] = np.ascontiguousarray(L.V[:, :3].T)

    # ---- G3: config-1 full size 200x200 Dirichlet and periodic, n = 100 -------------
    for tag, per in (("c1d", False), ("c1p", True)):
        H = orc.laplacian_csr((200, 200), 4.0, -1.0, periodic=per)
        L = run_regular(ref_reg, H, 100, seed=99)
        a, b = tri_parts(L.H_eff)
        out[f"{tag}_alpha"], out[f"{tag}_beta"], out[f"{tag}_theta"] = a, b, L.H_eigvals.copy()

    # ---- G4: 3-D 7-point periodic 12^3, n = 40, user start vector -------------------
    H = orc.laplacian_csr((12, 12, 12), 6.0, -1.0, periodic=True)
    v0 = np.random.RandomState(7).uniform(-1, 1, 12 ** 3)
    L = run_regular(ref_reg, H, 40, v0=v0)
    a, b = tri_parts(L.H_eff)
    out["c3s_alpha"], out["c3s_beta"], out["c3s_theta"] = a, b, L.H_eigvals.copy()

    # ---- G5: deuteron H (3Ddeuteron.py recipe) N = 16, n = 120: Ritz values converge -
    Hd, _, _, _ = orc.deuteron_hamiltonian(16)
    L = run_regular(ref_reg, Hd, 120, seed=78)
    a, b = tri_parts(L.H_eff)
    out["deut_alpha"], out["deut_beta"], out["deut_theta"] = a, b, L.H_eigvals.copy()

    # ---- G5b: the driver's actual operator: 27-point T (3Ddeuteron.py:76 default), N = 12, n = 60 ----
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            N = 12
            dx = 25.0 / N
            Tf = 197.327 ** 2 / (2 * 469.4592) * 1 / dx ** 2
            with quiet():
                ham = ref_ham.Hamiltonian(N, 25.0, orc.deuteron_potential, Tf)
                ham.create_sparse_T()
                ham.create_sparse_V()
            H27 = (-ham.T_sparse + ham.V_sparse)
            H27.sort_indices()
        finally:
            os.chdir(cwd)
    L = run_regular(ref_reg, H27, 60, seed=78)
    a, b = tri_parts(L.H_eff)
    out["deut27_alpha"], out["deut27_beta"], out["deut27_theta"] = a, b, L.H_eigvals.copy()

    # ---- G6: Irregular: Delaunay graph Laplacian 3000 vertices, CSR and CSC, n = 50 -
    Ld = orc.delaunay_graph_laplacian(3000, seed=0)
    L = run_irregular(ref_irr, Ld, 50, seed=99)
    a, b = tri_parts(L.H_eff)
    out["del_alpha"], out["del_beta"], out["del_theta"] = a, b, L.H_eigvals.copy()
    L = run_irregular(ref_irr, sp.csc_matrix(Ld), 50, seed=99)
    a, b = tri_parts(L.H_eff)
    out["delcsc_alpha"], out["delcsc_beta"] = a, b
    out["del_indptr_sha"] = np.frombuffer(
        __import__("hashlib").sha256(Ld.indptr.tobytes() + Ld.indices.tobytes()).digest(), dtype=np.uint8)

    # ---- G7: edge cases: n = 2 and n = M -------------------------------------------
    H = orc.laplacian_csr((6,), 2.0, -1.0, periodic=False)
    L = run_regular(ref_reg, H, 2, seed=3)
    out["n2_T"] = L.H_eff.copy()
    L = run_regular(ref_reg, H, 6, seed=3)
    out["nM_T"] = L.H_eff.copy()

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), "keys:", len(out))


if __name__ == "__main__":
    main()
