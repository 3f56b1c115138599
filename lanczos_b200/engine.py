"""Host-side plumbing between the Python drop-in classes and the C ABI.

PyTorch is used for device memory, streams and (in ``lanczos_b200.team``) process groups only;
every numerical operation happens inside liblanczos_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import (LZ_BC_DIRICHLET, LZ_BC_PERIODIC, LZ_FMT_CSR, LZ_FMT_SELL, LZ_FMT_SELL_VALUES, LZ_REORTH_FULL,
                    LZ_REORTH_NONE, LZ_REORTH_SELECTIVE, LanczosBreakdown, RunInfo, RunOpts)

_REORTH = {"none": LZ_REORTH_NONE, "full": LZ_REORTH_FULL, "selective": LZ_REORTH_SELECTIVE,
           None: LZ_REORTH_NONE, False: LZ_REORTH_NONE, True: LZ_REORTH_FULL}


# "sell": SELL-32-sigma, value-free when every off-diagonal entry is equal; "sell_values": always with the values
FORMATS = {"csr": LZ_FMT_CSR, "sell": LZ_FMT_SELL, "sell_values": LZ_FMT_SELL_VALUES}
STEP_KERNEL = {"auto": 0, "two_pass": 1, "fused": 2, "recompute": 3, 0: 0, 1: 1, 2: 2, 3: 3}
STEP_KERNEL_NAME = {1: "two_pass", 2: "fused", 3: "recompute", 4: "persistent"}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("lanczos_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch


def run_flags(cgs_fused: bool = True, sweep_form: int = 0, kb_alpha: bool = False, overlap: bool = True,
              kba: bool = False, persistent: bool = False) -> int:
    """lz_run_opts.flags (0 = the library defaults): bit 0 = CGS2 without K4c, bit 1 = the Regular GPU sweep
    form (LZ_SWEEP_GPU), bit 2 = alpha accumulated inside KB + border kernel, bit 3 = sparse row shards without
    the interior/boundary overlap, bit 4 = the single KBA kernel per step, bit 5 = small problems in one persistent
    cooperative kernel."""
    return ((0 if cgs_fused else 1) | (2 if sweep_form == _capi.LZ_SWEEP_GPU else 0) | (4 if kb_alpha else 0) |
            (0 if overlap else 8) | (16 if kba else 0) | (32 if persistent else 0))


def padded_ld(M: int) -> int:
    """Row stride of the basis: rows start on 512-byte boundaries."""
    return (int(M) + 63) // 64 * 64


def reference_T27_weights(T_factor: float = 1.0):
    """(centre, face, edge, corner) coefficients of the reference's 27-point Laplacian T
    (Hamiltonian.get_weights_27point, Hamiltonian.py:116-128)."""
    return tuple(T_factor * 3.0 / 13.0 * w for w in (-44.0 / 3.0, 1.0, 0.5, 1.0 / 3.0))


@dataclass(eq=False)
class StencilOperator:
    """Matrix-free structured-grid operator descriptor
        (H x)_i = (center + diag_i) x_i + sum_axis off[axis] * (x_{i+e_axis} + x_{i-e_axis})
    with the reference's index map i = x + nx*(y + ny*z) (Hamiltonian.py:73-84) and periodic
    (Hamiltonian.py:92-97) or Dirichlet (1Dbox.py:15-22) boundaries.  It stands in for the CSR
    matrix that Hamiltonian.create_sparse_T("7") (+ create_sparse_V) would build; `tocsr()`
    returns exactly that matrix (after sort_indices) for checks on small grids.
    """
    grid: Sequence[int]                    # (nx,), (nx, ny) or (nx, ny, nz); x is the fastest index
    center: float
    off: Sequence[float] | float
    bc: str = "periodic"
    diag: Optional[np.ndarray] = None      # potential on the diagonal, length M (host) or a CUDA tensor
    weights27: Optional[Sequence[float]] = None   # 27-point box stencil (3-D): (centre, face, edge, corner)
                                                  # coefficients; `center`/`off` are then ignored
    _dev: dict = field(default_factory=dict, repr=False, compare=False)

    def __post_init__(self):
        self.grid = tuple(int(s) for s in self.grid)
        if not 1 <= len(self.grid) <= 3:
            raise ValueError("StencilOperator: 1, 2 or 3 dimensions")
        if np.isscalar(self.off):
            self.off = (float(self.off),) * len(self.grid)
        self.off = tuple(float(o) for o in self.off)
        if len(self.off) != len(self.grid):
            raise ValueError("StencilOperator: one off-diagonal coefficient per axis")
        if self.bc not in ("periodic", "dirichlet"):
            raise ValueError("StencilOperator: bc must be 'periodic' or 'dirichlet'")
        if self.weights27 is not None:
            if len(self.grid) != 3 or len(tuple(self.weights27)) != 4:
                raise ValueError("StencilOperator: weights27 needs a 3-D grid and 4 coefficients")
            self.weights27 = tuple(float(w) for w in self.weights27)
        self.M = int(np.prod(self.grid))
        if self.diag is not None and int(np.prod(tuple(self.diag.shape))) != self.M:
            raise ValueError("StencilOperator: diag must have M entries")

    # duck-typed operator protocol of the reference (np.shape(H)[0], H*vec: Lanczos.py:22,108)
    @property
    def shape(self):
        return (self.M, self.M)

    @property
    def ndim(self):
        return 2

    @property
    def dtype(self):
        return np.dtype(np.float64)

    def device_handle(self, ctx: "Context") -> "DeviceOperator":
        key = id(ctx)
        if key not in self._dev:
            self._dev[key] = DeviceOperator.from_stencil(ctx, self)
        return self._dev[key]

    def matvec(self, x):
        """y = H x for a host vector, computed on the GPU (used by the diagnostics)."""
        ctx = Context.default()
        return self.device_handle(ctx).apply_host(np.asarray(x, dtype=np.float64))

    def __mul__(self, x):
        return self.matvec(x)

    def tocsr(self):
        ctx = Context.default()
        return self.device_handle(ctx).export_csr()


class Context:
    """One lz_ctx per (device, stream)."""
    _default = {}

    def __init__(self, device: Optional[int] = None, stream=None):
        torch = _torch()
        self.lib = _capi.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.torch_device = torch.device("cuda", self.device)
        with torch.cuda.device(self.device):
            self.stream = stream if stream is not None else torch.cuda.current_stream()
        h = C.c_void_p()
        _capi.check(self.lib.lz_ctx_create(self.device, C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.handle = h

    @classmethod
    def default(cls, device: Optional[int] = None) -> "Context":
        torch = _torch()
        dev = torch.cuda.current_device() if device is None else int(device)
        key = (dev, torch.cuda.current_stream(dev).cuda_stream)
        if key not in cls._default:
            cls._default[key] = cls(dev)
        return cls._default[key]

    def sync(self):
        _capi.check(self.lib.lz_ctx_sync(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.lz_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class DeviceOperator:
    """An lz_op handle plus the buffers it references."""

    def __init__(self, ctx: Context, handle, M: int, kind: str, keep=()):
        self.ctx, self.handle, self.M, self.kind = ctx, handle, int(M), kind
        self._keep = keep

    @classmethod
    def from_stencil(cls, ctx: Context, st: StencilOperator) -> "DeviceOperator":
        torch = _torch()
        dim = len(st.grid)
        shape = (C.c_int64 * dim)(*st.grid)
        off = (C.c_double * dim)(*st.off)
        diag_t = None
        diag_p = C.c_void_p(0)
        if st.diag is not None:
            if isinstance(st.diag, torch.Tensor):
                diag_t = st.diag.to(device=ctx.torch_device, dtype=torch.float64).contiguous().reshape(-1)
            else:
                diag_t = torch.from_numpy(np.ascontiguousarray(st.diag, dtype=np.float64).reshape(-1)).to(ctx.torch_device)
            diag_p = C.c_void_p(diag_t.data_ptr())
        h = C.c_void_p()
        bc = LZ_BC_PERIODIC if st.bc == "periodic" else LZ_BC_DIRICHLET
        if st.weights27 is not None:
            w = (C.c_double * 4)(*st.weights27)
            _capi.check(ctx.lib.lz_op_stencil27_create(ctx.handle, shape, bc, w, diag_p, C.byref(h)))
        else:
            _capi.check(ctx.lib.lz_op_stencil_create(ctx.handle, dim, shape, bc, float(st.center), off, diag_p, C.byref(h)))
        return cls(ctx, h, st.M, "stencil", keep=(diag_t,))

    @classmethod
    def from_scipy(cls, ctx: Context, H, fmt: str = "auto", sigma: int = 0) -> "DeviceOperator":
        import scipy.sparse as sp
        A = sp.csr_matrix(H) if not sp.isspmatrix_csr(H) else H
        if A.shape[0] != A.shape[1]:
            raise ValueError("operator must be square")
        A = A.astype(np.float64, copy=False)
        if A.nnz >= 2 ** 31:
            raise ValueError("nnz must fit int32 (scipy CSR layout)")
        indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(A.indices, dtype=np.int32)
        data = np.ascontiguousarray(A.data, dtype=np.float64)
        if fmt == "auto":
            fmt = "sell"
        f = FORMATS[fmt]
        h = C.c_void_p()
        _capi.check(ctx.lib.lz_op_csr_create(
            ctx.handle, A.shape[0], A.nnz, indptr.ctypes.data_as(C.c_void_p),
            indices.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), f, int(sigma), C.byref(h)))
        return cls(ctx, h, A.shape[0], fmt)

    @classmethod
    def from_device_csr(cls, ctx: Context, indptr, indices, data, ncols: Optional[int] = None,
                        fmt: str = "auto", sigma: int = 0) -> "DeviceOperator":
        """CSR arrays that already live on the GPU (CUDA tensors: int32 indptr/indices, fp64 data) -
        the counterpart of the cupyx matrix the reference holds in GPU mode (Lanczos.py:88).
        Converted on the device (lz_op_csr_create_dev); `ncols` > M makes a row shard."""
        torch = _torch()
        M = int(indptr.numel()) - 1
        nnz = int(indices.numel())
        for t, dt in ((indptr, torch.int32), (indices, torch.int32), (data, torch.float64)):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dt and t.is_contiguous()):
                raise TypeError("from_device_csr: contiguous CUDA tensors (int32 indptr/indices, float64 data)")
            if t.device.index != ctx.device:
                raise ValueError("from_device_csr: the arrays live on another device than the context")
        if fmt == "auto":
            fmt = "sell"
        f = FORMATS[fmt]
        h = C.c_void_p()
        torch.cuda.current_stream(ctx.device).synchronize()      # the arrays were produced on torch's stream
        _capi.check(ctx.lib.lz_op_csr_create_dev(
            ctx.handle, M, M if ncols is None else int(ncols), nnz, C.c_void_p(indptr.data_ptr()),
            C.c_void_p(indices.data_ptr() if nnz else 0), C.c_void_p(data.data_ptr() if nnz else 0),
            f, int(sigma), C.byref(h)))
        return cls(ctx, h, M, fmt)

    def nnz(self):
        t, s = C.c_int64(), C.c_int64()
        _capi.check(self.ctx.lib.lz_op_nnz(self.handle, C.byref(t), C.byref(s)))
        return t.value, s.value

    def value_free(self) -> bool:
        """True when the operator is applied from its column indices alone (all off-diagonal entries equal)."""
        v = C.c_int32()
        _capi.check(self.ctx.lib.lz_op_value_free(self.handle, C.byref(v)))
        return bool(v.value)

    def windowed(self) -> int:
        """0, or the largest number of 32-entry granules of x one sorting window stages in shared memory when the
        operator runs in the windowed SELL form (16-bit column offsets; csrc/sellw.cu)."""
        v = C.c_int32()
        _capi.check(self.ctx.lib.lz_op_windowed(self.handle, C.byref(v)))
        return int(v.value)

    def apply(self, x, y=None):
        """y = H x for CUDA tensors (enqueued on the context's stream)."""
        torch = _torch()
        if y is None:
            y = torch.empty_like(x)
        _capi.check(self.ctx.lib.lz_op_apply(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr())))
        return y

    def apply_host(self, x: np.ndarray) -> np.ndarray:
        torch = _torch()
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.ctx.torch_device)
        return self.apply(xd).cpu().numpy()

    def export_csr(self):
        import scipy.sparse as sp
        nnz = C.c_int64()
        _capi.check(self.ctx.lib.lz_op_export_csr(self.handle, C.byref(nnz), None, None, None))
        indptr = np.zeros(self.M + 1, dtype=np.int32)
        indices = np.zeros(nnz.value, dtype=np.int32)
        data = np.zeros(nnz.value, dtype=np.float64)
        _capi.check(self.ctx.lib.lz_op_export_csr(
            self.handle, C.byref(nnz), indptr.ctypes.data_as(C.c_void_p),
            indices.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p)))
        return sp.csr_matrix((data, indices, indptr), shape=(self.M, self.M))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.lz_op_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


@dataclass(eq=False)
class DeviceCSR:
    """A square sparse operator whose CSR arrays are CUDA tensors (int32 indptr / indices, fp64 data):
    what `cupyx.scipy.sparse.csr_matrix` is to the reference's GPU mode (Lanczos.py:88)."""
    indptr: object
    indices: object
    data: object

    @property
    def shape(self):
        M = int(self.indptr.numel()) - 1
        return (M, M)

    @property
    def nnz(self):
        return int(self.indices.numel())

    def get(self):
        """Host scipy matrix (the `.get()` of a cupyx matrix, Lanczos.py:137)."""
        import scipy.sparse as sp
        return sp.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()),
                             shape=self.shape)


def as_device_operator(H, ctx: Context, fmt: str = "auto", sigma: int = 0) -> DeviceOperator:
    """Map what the reference accepts as `H` onto a device operator.  scipy.sparse matrices
    (CSR as in Regular, CSC as IrrHamiltonian produces) and StencilOperator descriptors are
    supported; objects that only offer a host-side `H*vec` are not (that would be a CPU path)."""
    import scipy.sparse as sp
    if isinstance(H, DeviceOperator):
        return H
    if isinstance(H, StencilOperator):
        return H.device_handle(ctx)
    if isinstance(H, DeviceCSR):
        return DeviceOperator.from_device_csr(ctx, H.indptr, H.indices, H.data, fmt=fmt, sigma=sigma)
    if sp.issparse(H):
        return DeviceOperator.from_scipy(ctx, H, fmt=fmt, sigma=sigma)
    raise TypeError(
        f"unsupported operator type {type(H).__name__}: pass a scipy.sparse matrix or a "
        "lanczos_b200.StencilOperator (host-callback operators would need a CPU path)")


def operator_rows(H) -> int:
    if isinstance(H, (StencilOperator, DeviceOperator)):
        return H.M
    if isinstance(H, DeviceCSR):
        return H.shape[0]
    return int(np.shape(H)[0])


def start_vector(M: int, seed=99, v0=None) -> np.ndarray:
    """The reference's start vector (Lanczos.py:93-100): the legacy global NumPy stream is
    seeded even when a vector is supplied; the vector is NOT normalised here (the device does)."""
    np.random.seed(seed)
    if v0 is None:
        return np.random.uniform(-1, 1, size=(M))
    x = np.array(v0, dtype=np.float64).reshape(-1)
    if x.shape[0] != M:
        raise ValueError(f"v0 has {x.shape[0]} entries, the operator has {M} rows")
    return x


class LanczosResult:
    """alpha/beta on the host, the Krylov basis on the device."""

    def __init__(self, ctx, n, M, alpha, beta, V_dev, ld, row_scale, info):
        self.ctx, self.n, self.M = ctx, n, M
        self.alpha, self.beta = alpha, beta
        self.V_dev, self.ld, self.row_scale = V_dev, ld, row_scale
        self.steps_done = info.steps_done
        self.reorth_count = info.reorth_count
        self.launches = info.launches
        self.gpu_ms = float(info.gpu_ms)
        self.step_kernel = STEP_KERNEL_NAME.get(info.step_kernel, "two_pass")
        # per-kernel device times (only when run with profile=True)
        self.kernel_ms = {"apply": (float(info.apply_ms), info.apply_launches),
                          "update": (float(info.update_ms), info.update_launches),
                          "dots": (float(info.dots_ms), info.dots_launches),
                          "gs_update": (float(info.gsupd_ms), info.gsupd_launches),
                          "fused": (float(info.fused_ms), info.fused_launches),
                          "gs_fused": (float(info.gsfused_ms), info.gsfused_launches),
                          "border": (float(info.border_ms), info.border_launches)}
        self.alpha_in_update = int(info.alpha_in_update) == 1
        self.kba = int(info.alpha_in_update) == 2
        self.overlap = bool(info.overlap)
        self.graph = {0: "none", 1: "captured", 2: "replayed"}.get(int(info.graph), "none")

    def tridiagonal(self) -> np.ndarray:
        """Dense H_eff like Lanczos.py:121-130."""
        n = self.n
        T = np.zeros((n, n))
        i = np.arange(n)
        T[i, i] = self.alpha
        if n > 1:
            T[i[:-1], i[:-1] + 1] = self.beta
            T[i[:-1] + 1, i[:-1]] = self.beta
        return T

    def normalize_basis(self):
        if self.V_dev is None:
            raise ValueError("the basis was not kept (reorth='none', keep_basis=False)")
        if np.any(self.row_scale != 1.0):
            _capi.check(self.ctx.lib.lz_basis_normalize(
                self.ctx.handle, C.c_void_p(self.V_dev.data_ptr()), self.ld, self.n, self.M,
                self.row_scale.ctypes.data_as(C.c_void_p)))
            self.row_scale[:] = 1.0

    def basis_rows_host(self) -> np.ndarray:
        """(n, M) row-major host copy of the normalised basis (the in-loop layout, Lanczos.py:104)."""
        self.normalize_basis()
        return self.V_dev[:, :self.M].cpu().numpy()

    def ritz_vectors_dev(self, S: np.ndarray):
        """Y (k, ld) CUDA tensor with Y[c] = V @ S[:, c]  (Lanczos.py:154-156)."""
        torch = _torch()
        if self.V_dev is None:
            raise ValueError("the basis was not kept")
        S = np.asfortranarray(S, dtype=np.float64)
        n, k = S.shape
        if n != self.n:
            raise ValueError("S must have n rows")
        Y = torch.empty((k, self.ld), dtype=torch.float64, device=self.ctx.torch_device)
        _capi.check(self.ctx.lib.lz_ritz_vectors(
            self.ctx.handle, C.c_void_p(self.V_dev.data_ptr()), self.ld, n, self.M,
            self.row_scale.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p), k,
            C.c_void_p(Y.data_ptr()), self.ld))
        return Y


def run_lanczos(op: DeviceOperator, v0, n: int, *, reorth="full", cgs_passes=1, ref_compat=True,
                keep_basis=True, breakdown_tol=0.0, select_tol=0.0, V_dev=None,
                profile=False, step_kernel="auto", cgs_fused=True, sweep_form=0, kb_alpha=False, kba=False, persistent=False) -> LanczosResult:
    """Enqueue and run the n-step loop (lz_lanczos_run).  `v0` is a host array (copied through
    pinned memory) or a CUDA tensor of M doubles."""
    torch = _torch()
    ctx = op.ctx
    M = op.M
    n = int(n)
    mode = _REORTH[reorth] if not isinstance(reorth, int) or isinstance(reorth, bool) else reorth
    need_basis = keep_basis or mode != LZ_REORTH_NONE
    with torch.cuda.device(ctx.device):
        if isinstance(v0, torch.Tensor):
            v0_dev = v0.to(device=ctx.torch_device, dtype=torch.float64).contiguous().reshape(-1)
        else:
            host = torch.from_numpy(np.ascontiguousarray(v0, dtype=np.float64).reshape(-1))
            try:
                host = host.pin_memory()
            except RuntimeError:
                pass
            v0_dev = host.to(ctx.torch_device, non_blocking=True)
        if v0_dev.numel() != M:
            raise ValueError(f"v0 has {v0_dev.numel()} entries, the operator has {M} rows")
        ld = padded_ld(M)
        if need_basis and V_dev is None:
            V_dev = torch.empty((n, ld), dtype=torch.float64, device=ctx.torch_device)
        elif not need_basis:
            V_dev = None
        alpha = np.zeros(n)
        beta = np.zeros(max(n - 1, 0))
        scale = np.ones(n)
        opts = RunOpts(mode, int(cgs_passes), 1 if ref_compat else 0, 1 if profile else 0,
                       STEP_KERNEL[step_kernel], run_flags(cgs_fused, sweep_form, kb_alpha, True, kba, persistent), float(breakdown_tol), float(select_tol))
        info = RunInfo()
        status = ctx.lib.lz_lanczos_run(
            ctx.handle, op.handle, C.c_void_p(v0_dev.data_ptr()), n, C.byref(opts),
            alpha.ctypes.data_as(C.c_void_p), beta.ctypes.data_as(C.c_void_p),
            C.c_void_p(V_dev.data_ptr() if V_dev is not None else 0), ld,
            scale.ctypes.data_as(C.c_void_p), C.byref(info))
    if status == _capi.LZ_ERR_BREAKDOWN:
        msg = ctx.lib.lz_last_error().decode()
        raise LanczosBreakdown(msg, steps_done=info.steps_done)
    _capi.check(status)
    return LanczosResult(ctx, n, M, alpha, beta, V_dev, ld, scale, info)


def reorthogonalize_rows(V, j: int, form: int = 0):
    """Lanczos.reorthogonalize(V, j) for a basis held as rows: V is a CUDA tensor (n, ld>=M)
    [in place] or a host (n, M) array [updated in place through the device].  `form`:
    LZ_SWEEP_CPU (2 V[j] - sum, Lanczos.py:247-249) or LZ_SWEEP_GPU (self term dropped, :236-238)."""
    torch = _torch()
    if isinstance(V, torch.Tensor):
        ctx = Context.default(V.device.index)
        n, M = V.shape
        ld = V.stride(0)
        _capi.check(ctx.lib.lz_reorthogonalize(ctx.handle, C.c_void_p(V.data_ptr()), ld, n, M, int(j), int(form)))
        return V
    A = np.asarray(V)
    ctx = Context.default()
    n, M = A.shape
    ld = padded_ld(M)
    Vd = torch.zeros((n, ld), dtype=torch.float64, device=ctx.torch_device)
    Vd[:, :M] = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float64)).to(ctx.torch_device)
    _capi.check(ctx.lib.lz_reorthogonalize(ctx.handle, C.c_void_p(Vd.data_ptr()), ld, n, M, int(j), int(form)))
    V[j] = Vd[j, :M].cpu().numpy()
    return V
