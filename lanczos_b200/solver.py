"""Shared implementation behind the two drop-in classes (regular.Lanczos, irregular.IrrLanczos).

The public surface (constructor, method names, keyword names and defaults, properties, error
types and messages) follows the reference classes - Python/Regular/Lanczos.py:11-337 and
Python/Irregular/IrrLanczos.py:12-554 - while the loop itself runs in liblanczos_b200.so.
Diagnostics that the reference computes on the host after the loop (ARPACK comparison, the
n x n eigenproblem, pretty-printing) stay on the host, as north-star item (4) asks.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import engine
from .engine import Context, LanczosBreakdown, StencilOperator, as_device_operator, operator_rows

_NOT_RUN = "Lanczos Algorithm has not been called."      # Lanczos.py:31


class LanczosBase:
    # The reference has two forms of its Gram-Schmidt sweep (see lz_reorthogonalize): Regular with
    # use_cuda=True drops the self term (Lanczos.py:236-238), everything else uses 2 V[j] - sum
    # (Lanczos.py:247-249, IrrLanczos.py:453-460).  Subclasses say which one `use_cuda=True` means.
    _GPU_SWEEP_FORM = 0

    # ---- construction (Lanczos.py:19-26) --------------------------------------------------
    def __init__(self, H):
        self.H = H
        self.M = operator_rows(H)
        self.Lanczos_has_been_executed = False
        self.H_eigs_have_been_found = False
        self.H_exact_eigs_have_been_found = False
        self._result = None
        self._V_host = None
        self._team = None            # row-sharded run (devices=...): lanczos_b200.team object

    # ---- lazy properties (Lanczos.py:28-66) -----------------------------------------------
    @property
    def H_eff(self):
        if not self.Lanczos_has_been_executed:
            raise ValueError(_NOT_RUN)
        return self._H_eff

    @property
    def V(self):
        """(M, n) array whose columns are the Lanczos vectors (Lanczos.py:139: `V.T`).  The basis
        lives in HBM; it is normalised and copied to the host on first access."""
        if not self.Lanczos_has_been_executed:
            raise ValueError(_NOT_RUN)
        if self._V_host is None:
            self._V_host = (self._team if self._team is not None else self._result).basis_rows_host()
        return self._V_host.T

    @property
    def H_eigvecs(self):
        if not self.H_eigs_have_been_found:
            self.get_H_eigs()
        return self._H_eigvecs

    @property
    def H_eigvals(self):
        if not self.H_eigs_have_been_found:
            self.get_H_eigs()
        return self._H_eigvals

    @property
    def H_eigvals_actual(self):
        if not self.H_exact_eigs_have_been_found:
            self.find_exact_eigs()
            self.H_exact_eigs_have_been_found = True
        return self._H_eigvals_actual

    @property
    def H_eigvecs_actual(self):
        if not self.H_exact_eigs_have_been_found:
            self.find_exact_eigs()
            self.H_exact_eigs_have_been_found = True
        return self._H_eigvecs_actual

    @property
    def result(self):
        """The engine-level result (alpha, beta, device basis, timings)."""
        if not self.Lanczos_has_been_executed:
            raise ValueError(_NOT_RUN)
        return self._result

    # ---- exact eigenpairs for comparison (Lanczos.py:68-71): host ARPACK, a diagnostic -------
    def find_exact_eigs(self, nr_vecs=20):
        import scipy.sparse.linalg
        print("+++ Calculating exact eigs using scipy.sparse.linalg.eigsh.")
        H = self.H.tocsr() if isinstance(self.H, StencilOperator) else self.H
        self._H_eigvals_actual, self._H_eigvecs_actual = scipy.sparse.linalg.eigsh(H, k=nr_vecs, which="SM")
        print("+++ Finished calculating exact eigs.")

    # ---- the loop ---------------------------------------------------------------------------
    def _execute(self, n, seed, use_cuda, v0, *, reorth="full", cgs_passes=1, ref_compat=True,
                 fmt="auto", sigma=0, device=None, keep_basis=True, breakdown_tol=0.0,
                 select_tol=0.0, profile=False, step_kernel="auto", cgs_fused=True, kb_alpha=False, kba=False, persistent=False, verbose=True,
                 devices=None):
        """Keyword-only extras (all default to the reference's behaviour):
        reorth 'full' | 'selective' | 'none'; cgs_passes 1 | 2; ref_compat (the v0-discarding
        pre-step and the (2-|v|^2) sweep form of the reference); fmt 'auto' | 'csr' | 'sell' and
        sigma for sparse operators; device index; keep_basis; breakdown_tol; select_tol;
        profile (per-kernel CUDA-event timing); step_kernel 'auto' | 'two_pass' | 'fused' | 'recompute';
        cgs_fused (CGS2: one read of the basis for the update of sweep 1 and the dots of sweep 2);
        verbose (the reference's '+++' banners);
        devices: row-sharded run over several GPUs behind the same API - a list of device indices
        (this process drives one shard per device) or "auto" (one process per GPU under torchrun: the
        initialised torch.distributed group is the team).  Every result (H_eff, V, H_eigvals,
        H_eigvecs, print_good_eigs) is then assembled from the shards; alpha/beta are bit-identical
        on every rank."""
        if n > self.M:
            raise ValueError("n cannot be larger than M!")                  # Lanczos.py:76-77
        if ref_compat and n < 2:
            # the reference writes beta[j-1] into an empty array (Lanczos.py:107,112)
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")
        if not use_cuda:
            warnings.warn("use_cuda=False: lanczos_b200 has no CPU path, the B200 backend is used",
                          RuntimeWarning, stacklevel=3)
        if verbose:
            print("+++ Executing Lanczos algorithm")
        self.n = n
        team = self._team_for(devices, fmt, sigma)
        if team is not None:
            np.random.seed(seed)
            self._result = None
            self._V_host = None
            self._team = None
            self.Lanczos_has_been_executed = False
            team.execute_Lanczos(n, seed, True, v0, reorth=reorth, cgs_passes=cgs_passes, ref_compat=ref_compat,
                                 keep_basis=keep_basis, breakdown_tol=breakdown_tol, select_tol=select_tol,
                                 profile=profile, step_kernel=step_kernel, cgs_fused=cgs_fused, kb_alpha=kb_alpha,
                                 sweep_form=self._GPU_SWEEP_FORM if use_cuda else 0)
            self._team = team
            self._result = team.result
            self._H_eff = team.H_eff
            self._Y_dev = None
            self.H_eigs_have_been_found = False
            if verbose:
                print("+++ Lanczos executed successfully.")
            self.Lanczos_has_been_executed = True
            return
        self._team = None
        if devices is not None and devices != "auto" and len(list(devices)) == 1:
            device = int(list(devices)[0])
        ctx = Context.default(device)
        # the device copy of the operator (CSR upload / SELL conversion) is made once per instance
        key = (id(ctx), fmt, sigma)
        cache = self.__dict__.setdefault("_op_cache", {})
        if key not in cache:
            cache[key] = as_device_operator(self.H, ctx, fmt=fmt, sigma=sigma)
        op = cache[key]
        torch = engine._torch()
        if isinstance(v0, torch.Tensor):
            np.random.seed(seed)                                             # Lanczos.py:93
            start = v0
        else:
            start = engine.start_vector(self.M, seed, v0)
        self._device_op = op
        # release the previous basis first: at 512^3 it is ~100 GB and the allocator reuses it
        self._result = None
        self._V_host = None
        self.Lanczos_has_been_executed = False
        self._result = engine.run_lanczos(op, start, n, reorth=reorth, cgs_passes=cgs_passes,
                                          ref_compat=ref_compat, keep_basis=keep_basis,
                                          breakdown_tol=breakdown_tol, select_tol=select_tol,
                                          profile=profile, step_kernel=step_kernel, cgs_fused=cgs_fused,
                                          sweep_form=self._GPU_SWEEP_FORM if use_cuda else 0, kb_alpha=kb_alpha, kba=kba,
                                          persistent=persistent)
        self._H_eff = self._result.tridiagonal()
        self._V_host = None
        self._Y_dev = None
        self.H_eigs_have_been_found = False
        if verbose:
            print("+++ Lanczos executed successfully.")
        self.Lanczos_has_been_executed = True

    def _team_for(self, devices, fmt, sigma):
        """The row-sharded driver for `devices` (None: single-GPU run), built once per instance."""
        if devices is None:
            return None
        from . import team as lzteam
        if isinstance(devices, str):
            if devices != "auto":
                raise ValueError("devices: a list of device indices or 'auto'")
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
                return None
            key, make = ("dist", fmt, sigma), (lambda: lzteam.TeamLanczos(self.H, fmt=fmt, sigma=sigma))
        else:
            devs = [int(d) for d in devices]
            if len(devs) <= 1:
                return None
            key = (tuple(devs), fmt, sigma)
            make = (lambda: lzteam.LocalTeamLanczos(self.H, len(devs), devices=devs, fmt=fmt, sigma=sigma))
        cache = self.__dict__.setdefault("_team_cache", {})
        if key not in cache:
            cache[key] = make()
        return cache[key]

    # ---- Ritz pairs (Lanczos.py:145-163) ------------------------------------------------------
    def _ritz(self, check_vectors):
        if not self.Lanczos_has_been_executed:
            raise ValueError(_NOT_RUN)
        print("+++ Converting eigenvectors from H_eff to H basis.")
        theta, S = np.linalg.eigh(self.H_eff)                   # small n x n problem: host LAPACK
        if self._team is not None:
            Y = self._team.ritz_vectors_dev(S)                  # K5 on every shard: each lifts its own rows
            self._Y_dev = Y
            Yh = np.ascontiguousarray(self._team.rows_host(Y).T)
        else:
            Y = self._result.ritz_vectors_dev(S)                # K5 on the device
            self._Y_dev = Y                                     # kept for the residual diagnostics
            Yh = np.ascontiguousarray(Y[:, :self.M].cpu().numpy().T)   # (M, n) like the reference
        if check_vectors:
            self.test_is_normalized(Yh, tol=0.001)
            self.test_is_orthogonal(Yh, tol=0.01)
        self._H_eigvals = theta
        self._H_eigvecs = Yh
        print("+++ Finished Converting.")
        self.H_eigs_have_been_found = True

    def ritz_values(self, k=None):
        """Lowest-k Ritz values without touching the basis (eigvalsh of H_eff)."""
        theta = np.linalg.eigvalsh(self.H_eff)
        return theta if k is None else theta[:k]

    def ritz_vectors(self, k, which="lowest"):
        """(theta[:k], Y) with Y a CUDA tensor (k, M): only the wanted Ritz vectors are lifted,
        (n + k) * 8 * M bytes of traffic instead of the reference's full (M, n) product.  After a
        row-sharded run Y is a list with one (k, M_local) tensor per local shard."""
        theta, S = np.linalg.eigh(self.H_eff)
        sel = np.arange(k) if which == "lowest" else np.arange(len(theta) - k, len(theta))
        if self._team is not None:
            Ys = self._team.ritz_vectors_dev(S[:, sel])
            return theta[sel], [Y[:, :self._team.plan.local_rows(s.rank)] for s, Y in zip(self._team.shards, Ys)]
        Y = self._result.ritz_vectors_dev(S[:, sel])
        return theta[sel], Y[:, :self.M]

    # ---- diagnostics (Lanczos.py:166-222), host-side printing over device products ------------
    def _residual_cosines(self):
        """cos^2 of the angle between H x and x for every Ritz vector (Lanczos.py:171-176:
        `Hx = H*x; Hx /= norm(Hx); dot(Hx, x)**2`), with the products on the device: the operator
        kernel on the Ritz vectors K5 left in HBM, two deterministic dots per vector."""
        import ctypes as C
        from . import _capi
        torch = engine._torch()
        eigvecs = self.H_eigvecs                                  # runs get_H_eigs when needed
        if self._team is not None:
            Ys = getattr(self, "_Y_dev", None)
            if not isinstance(Ys, list):
                Ys = self._team.ritz_vectors_dev(np.linalg.eigh(self.H_eff)[1])
            return self._team.residual_cosines(Ys)
        ctx = Context.default()
        op = getattr(self, "_device_op", None)
        if op is None or op.ctx is not ctx:
            op = as_device_operator(self.H, ctx)
            self._device_op = op
        Y = getattr(self, "_Y_dev", None)
        if Y is None or Y.shape[0] != self.n:
            ld = engine.padded_ld(self.M)
            Y = torch.zeros((self.n, ld), dtype=torch.float64, device=ctx.torch_device)
            Y[:, :self.M] = torch.from_numpy(np.ascontiguousarray(eigvecs.T)).to(ctx.torch_device)
        u = torch.empty(Y.shape[1], dtype=torch.float64, device=ctx.torch_device)
        inner_prod = np.zeros(self.n)
        ux, uu = C.c_double(), C.c_double()
        torch.cuda.current_stream(ctx.device).synchronize()
        for i in range(self.n):
            xi = Y[i]
            _capi.check(ctx.lib.lz_op_apply(op.handle, C.c_void_p(xi.data_ptr()), C.c_void_p(u.data_ptr())))
            _capi.check(ctx.lib.lz_dot(ctx.handle, C.c_void_p(u.data_ptr()), C.c_void_p(xi.data_ptr()), self.M, C.byref(ux)))
            _capi.check(ctx.lib.lz_dot(ctx.handle, C.c_void_p(u.data_ptr()), C.c_void_p(u.data_ptr()), self.M, C.byref(uu)))
            inner_prod[i] = ux.value ** 2 / uu.value if uu.value > 0.0 else 0.0
        return inner_prod

    def compare_eigs(self):
        if not self.Lanczos_has_been_executed:
            raise ValueError(_NOT_RUN)
        print("+++ Comparing to exact eigs.")
        val_a, vec_a = self.H_eigvals_actual, self.H_eigvecs_actual
        val_e, vec_e = self.H_eigvals, self.H_eigvecs
        nr = len(val_a)
        pairs = np.full((nr, 2), np.nan)
        pairs[:, 0] = val_a
        overlap = np.full(nr, np.nan)
        idx_pairs = np.full(nr, np.nan)
        for i in range(self.n):
            o = np.dot(vec_e[:, i], vec_a) ** 2
            k = o.argmax()
            if np.isnan(overlap[k]) or o[k] > overlap[k]:
                pairs[k, 1], overlap[k], idx_pairs[k] = val_e[i], o[k], i
        perc = abs((pairs[:, 0] - pairs[:, 1]) / pairs[:, 1]) * 100
        print("__________EIGENVALUE AND EIGVENVECTOR COMPARISON__________")
        print("%6s %6s %20s %20s %14s %14s" % ("Idx1", "Idx2", "Actual", "Lanczos", "% Diff", "Eigvec Prod"))
        for i in range(nr):
            print("%6d %6.0f %20.10f %20.10f %14.4f %14.4f" % (i, idx_pairs[i], pairs[i, 0], pairs[i, 1], perc[i], overlap[i]))
        return pairs, overlap

    # ---- static helpers kept from the reference (Lanczos.py:233-337) ---------------------------
    @classmethod
    def reorthogonalize(cls, V, j, use_cuda=True):
        """One Gram-Schmidt sweep of row j of V against all rows, in place, in the form the
        reference class uses for this `use_cuda` (Lanczos.py:236-238 / :247-249;
        IrrLanczos.py:453-460).  V: (n, M) host array or CUDA tensor with rows = Lanczos vectors.
        Always computed on the GPU (there is no CPU path)."""
        return engine.reorthogonalize_rows(V, j, form=cls._GPU_SWEEP_FORM if use_cuda else 0)

    @staticmethod
    def get_matched_eigs(v, vL, l, lL):
        n = len(lL)
        overlap = np.zeros(n)
        map_vL2v = np.zeros(n, dtype=int)
        for i in range(n):
            o = np.dot(vL[:, i], v) ** 2
            map_vL2v[i] = o.argmax()
            overlap[i] = o[map_vL2v[i]]
        order = overlap.argsort()[::-1]
        return v[:, map_vL2v[order]], vL[:, order], l[map_vL2v[order]], lL[order]

    @staticmethod
    def test_is_Hermitian(A):
        import scipy.sparse as sp
        B = A.tocsr() if isinstance(A, StencilOperator) else A
        if sp.issparse(B):
            assert abs(B - B.T).max() == 0, "A IS NOT HERMITIAN!"
        else:
            assert (np.asarray(B) == np.asarray(B).T).all(), "A IS NOT HERMITIAN!"

    @staticmethod
    def test_is_normalized(V, tol=0.001, no_assert=False):
        norms = np.linalg.norm(V, axis=0)
        worst = norms[np.argmin(np.abs(norms - 1))]
        if no_assert:
            return worst
        assert np.abs(worst - 1) < tol, "VECTOR HAS NORM %.4f. IS NOT NORMALIZED." % worst

    @staticmethod
    def test_is_orthogonal(V, tol=0.01, no_assert=False):
        G = np.abs(V.T @ V - np.eye(V.shape[1]) * np.linalg.norm(V, axis=0) ** 2)
        at = np.unravel_index(np.argmax(G), G.shape)
        err = np.sqrt(G[at])
        if no_assert:
            return err
        assert err < tol, "VECTORS %d AND %d NOT ORTHOGONAL! INNER PRODUCT %.4f" % (at[0], at[1], err)

    @staticmethod
    def test_is_eigvecs(A, V, tol=0.01, no_assert=False):
        N = np.shape(A)[0]
        errors = np.zeros(V.shape[1])
        for i in range(V.shape[1]):
            t = (A * V[:, i]) / V[:, i] if not isinstance(A, np.ndarray) else np.dot(A, V[:, i]) / V[:, i]
            errors[i] = np.max(t) - np.min(t)
        if no_assert:
            return np.max(errors)
        assert np.max(errors) > tol, "VECTOR NOT EIGENVECTOR."      # (sense as in Lanczos.py:337)


__all__ = ["LanczosBase", "LanczosBreakdown", "StencilOperator"]
