#!/usr/bin/env python
"""Golden files for the on-disk formats (lanczos_b200/io.py), produced by the UNMODIFIED reference.
Runs only in the authoring container (needs /root/reference); the outputs are committed.

  * T_N=4_Laplace=7.npz / T_N=4_Laplace=27.npz : written by Hamiltonian.create_sparse_T itself
    (Hamiltonian.py:48-69) into ./T_matrices of a scratch directory;
  * matrix_d=3_N=2_L=25_p=Deuteron.dat : the writer of MatrixWrite.py cannot be imported here (IrrGrid
    needs matplotlib and breaks under NumPy 2), so its formatting block (the lines between
    "### Wiring to file ###" and the end of the function) is read from the reference tree at run time
    and executed on a stand-in `Ham` object holding a small sparse matrix - the reference's own code
    formats the file, nothing of it is copied into this repository.

    python tests/golden/make_golden_io.py
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, load_reference, quiet  # noqa: E402


def t_matrices():
    _, _, ham = load_reference()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        os.makedirs("T_matrices")
        try:
            for points in ("7", "27"):
                with quiet():
                    h = ham.Hamiltonian(4, 25.0, lambda x, y, z: 0.0, 1.5)
                    h.create_sparse_T(points=points)
                name = "T_N=4_Laplace=%s.npz" % points
                shutil.copy(os.path.join("T_matrices", name), os.path.join(HERE, name))
                print("wrote", name, h.T_sparse.shape, h.T_sparse.nnz)
        finally:
            os.chdir(cwd)


def matrix_dat():
    lines = open(os.path.join(REF, "Python", "Irregular", "MatrixWrite.py")).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "Wiring to file" in l)
    end = next(i for i in range(start, len(lines)) if "outfile.write" in lines[i]) + 1
    # the block sits inside a function body; its multi-line f-string continues at column 0
    block = "\n".join(l[4:] if l.startswith("    ") else l for l in lines[start:end])
    rng = np.random.RandomState(4)
    N = 2
    A = sp.random(N ** 3, N ** 3, density=0.4, random_state=rng, format="csr")
    A = sp.csr_matrix(A + A.T + sp.diags(np.arange(1.0, N ** 3 + 1) / 3.0))
    ham = types.SimpleNamespace(H_sparse=A)
    cwd = os.getcwd()
    os.chdir(HERE)
    try:
        exec(block, {"np": np, "sparse": sp, "Ham": ham, "d": 3, "L": 25, "N": N, "p": "Deuteron"})
    finally:
        os.chdir(cwd)
    sp.save_npz(os.path.join(HERE, "matrix_dat_input.npz"), A)
    print("wrote matrix_d=3_N=2_L=25_p=Deuteron.dat", A.nnz, "entries")


if __name__ == "__main__":
    t_matrices()
    matrix_dat()
