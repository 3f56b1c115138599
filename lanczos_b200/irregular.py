"""Drop-in for Python/Irregular/IrrLanczos.py: class IrrLanczos (irregular-mesh operators)."""
from __future__ import annotations

import numpy as np

from .solver import LanczosBase


class IrrLanczos(LanczosBase):
    """Mirrors the reference class (Python/Irregular/IrrLanczos.py:12).  The symmetric loop is
    `execute_LanczosOld` (:193-260) - the one Irr3Ddeuteron.py:39 calls.  The reference's
    `execute_Lanczos` (:77-187) is an experimental two-sided variant that is broken at HEAD
    (SURVEY.md §2.1); here it runs the same symmetric loop."""

    def execute_LanczosOld(self, n, seed=99, use_cuda=True, v0=None, **options):
        """IrrLanczos.py:193-260; `options` as in LanczosBase._execute."""
        self._execute(n, seed, use_cuda, v0, **options)

    def execute_Lanczos(self, n, seed=99, use_cuda=True, v0=None, dtype=np.float64, **kw):
        if np.dtype(dtype) != np.float64:
            raise ValueError("lanczos_b200 computes in float64 only")
        if self.H.shape[0] != self.H.shape[1]:
            raise AssertionError("H must be square")                     # IrrLanczos.py:80
        self.execute_LanczosOld(n, seed=seed, use_cuda=use_cuda, v0=v0, **kw)

    def get_H_eigs(self):
        """IrrLanczos.py:285-305 (no asserts)."""
        self._ritz(check_vectors=False)

    def get_H_eigsOld(self):
        """IrrLanczos.py:264-282 (normalisation assert only)."""
        self._ritz(check_vectors=False)
        self.test_is_normalized(self._H_eigvecs, tol=0.001)

    def print_good_eigs(self, tol=0.01, print_nr=20, print_bad=True, normal_eq=False):
        """IrrLanczos.py:331-353, with the sort index the reference meant (`sort_idxs`)."""
        eigvals = self.H_eigvals
        inner_prod = self._residual_cosines()
        if normal_eq:
            eigvals = np.sqrt(eigvals)
        order = np.argsort(np.abs(eigvals))
        print("__________EIGENVALUE AND EIGVENVECTOR COMPARISON__________")
        print("%12s %12s" % ("Eigval", "Eigvec InnerProd"))
        for i in range(min(print_nr, self.n)):
            k = order[i]
            tag = "" if abs(1 - inner_prod[k]) < tol else " --- BAD"
            print("%12.4f%12.6f%s" % (eigvals[k], inner_prod[k], tag))
        return inner_prod

    print_good_eigsOld = print_good_eigs
