"""Drop-in for Python/Regular/Lanczos.py: class Lanczos (structured-grid Hamiltonians)."""
from __future__ import annotations

from .solver import LanczosBase


class Lanczos(LanczosBase):
    """Lanczos tridiagonalisation of a Hermitian operator, B200 backend.

    Mirrors the reference class (Python/Regular/Lanczos.py:11): construct with the operator,
    call execute_Lanczos(n), read H_eff / V / H_eigvals / H_eigvecs.  `H` is a scipy.sparse
    matrix (what Hamiltonian.py builds) or a matrix-free lanczos_b200.StencilOperator."""

    _GPU_SWEEP_FORM = 1          # Lanczos.py:236-238: with use_cuda=True the sweep drops the self term

    def execute_Lanczos(self, n, seed=99, use_cuda=True, v0=None, **options):
        """Lanczos.py:75-141.  Positional/keyword arguments as in the reference; `options` are the
        keyword-only extras of LanczosBase._execute, which default to the reference's behaviour
        (full re-orthogonalisation in the reference's single-sweep form, the v0-discarding
        pre-step, basis kept)."""
        self._execute(n, seed, use_cuda, v0, **options)

    def get_H_eigs(self):
        """Lanczos.py:145-163 (with the normalisation / orthogonality asserts of :157-158)."""
        self._ritz(check_vectors=True)

    def print_good_eigs(self, tol=0.01, print_nr=20, print_bad=True):
        """Lanczos.py:166-185: cos^2 of the angle between H x and x for every Ritz vector."""
        eigvals = self.H_eigvals
        inner_prod = self._residual_cosines()
        print("__________EIGENVALUE AND EIGVENVECTOR COMPARISON__________")
        print("%12s %12s" % ("Eigval", "Eigvec InnerProd"))
        for i in range(min(print_nr, self.n)):
            if abs(1 - inner_prod[i]) < tol:
                print("%12.4f %20.14f" % (eigvals[i], inner_prod[i]))
            else:
                print("%12.4f %20.14f --- BAD" % (eigvals[i], inner_prod[i]))
        return inner_prod
