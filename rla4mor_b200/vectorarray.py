"""Minimal device-resident stand-ins for the pyMOR objects the embedding API touches.

The reference's operators take and return pyMOR `VectorArray`s living in
`NumpyVectorSpace`s (rla/embeddings.py:116,121,140,168-171).  pyMOR is not part of
this engine (and is absent from the build image), so the hot path carries its own
small equivalents with the same method names and the same `(len, dim)` row layout
(SURVEY.md Appendix B): a block of m vectors of dimension n is an (m, n) array.  The
data is a CUDA `torch.Tensor`; `to_numpy()` / `from_numpy()` are the host boundary.
"""
import numpy as np

from ._lib import require_cuda


def _torch():
    return require_cuda()


def as_device_block(x, dtype=None):
    """numpy (m, n) / torch tensor / DeviceVectorArray -> CUDA tensor (m, n)."""
    torch = _torch()
    if isinstance(x, DeviceVectorArray):
        t = x.data
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    else:
        assert isinstance(x, torch.Tensor), f"cannot use {type(x)} as a block of vectors"
        assert x.is_cuda, "torch inputs must live on the GPU (there is no CPU path)"
        t = x
    if t.dim() == 1:
        t = t.reshape(1, -1)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


class DeviceVectorSpace:
    """Counterpart of pyMOR's NumpyVectorSpace(dim, id)."""

    def __init__(self, dim, id=None):
        self.dim = int(dim)
        self.id = id

    def __eq__(self, other):
        return isinstance(other, DeviceVectorSpace) and self.dim == other.dim and self.id == other.id

    def __hash__(self):
        return hash((self.dim, self.id))

    def __contains__(self, U):
        return isinstance(U, DeviceVectorArray) and U.space == self

    def __repr__(self):
        return f"DeviceVectorSpace({self.dim}, id={self.id!r})"

    def from_numpy(self, data):
        t = as_device_block(data)
        assert t.shape[1] == self.dim, f"expected vectors of dimension {self.dim}, got {t.shape[1]}"
        return DeviceVectorArray(self, t)

    make_array = from_numpy

    def empty(self, reserve=0):
        torch = _torch()
        return DeviceVectorArray(self, torch.empty((0, self.dim), dtype=torch.float64, device="cuda"))

    def zeros(self, count=1):
        torch = _torch()
        return DeviceVectorArray(self, torch.zeros((count, self.dim), dtype=torch.float64, device="cuda"))


class DeviceVectorArray:
    """Counterpart of pyMOR's NumpyVectorArray: `len(U)` vectors of dimension `U.dim`."""

    def __init__(self, space, data):
        self.space = space
        self.data = data

    @property
    def dim(self):
        return self.space.dim

    def __len__(self):
        return int(self.data.shape[0])

    def __getitem__(self, ind):
        d = self.data[ind]
        if d.dim() == 1:
            d = d.reshape(1, -1)
        return DeviceVectorArray(self.space, d)

    def to_numpy(self, ensure_copy=False):
        return self.data.detach().cpu().numpy()

    def copy(self):
        return DeviceVectorArray(self.space, self.data.clone())

    def append(self, other):
        torch = _torch()
        assert other in self.space
        self.data = torch.cat([self.data.to(other.data.dtype) if len(self) == 0 else self.data, other.data], dim=0)

    def lincomb(self, coefficients):
        """rows of the result = coefficients @ rows  (pyMOR semantics)."""
        torch = _torch()
        c = torch.as_tensor(np.atleast_2d(np.asarray(coefficients)), dtype=self.data.dtype, device=self.data.device)
        assert c.shape[1] == len(self)
        if self.data.dtype == torch.float64 and len(self) > 0:
            from .reductor_ops import gemm_nn                  # (r', r) @ (r, n) on the sketch GEMM
            return DeviceVectorArray(self.space, gemm_nn(c.contiguous(), self.data))
        return DeviceVectorArray(self.space, c @ self.data)

    def norm(self):
        torch = _torch()
        return torch.linalg.norm(self.data, dim=1).cpu().numpy()

    def inner(self, other):
        return (self.data.conj() @ other.data.T).cpu().numpy()

    def scal(self, alpha):
        self.data = self.data * alpha

    def __repr__(self):
        return f"DeviceVectorArray(len={len(self)}, dim={self.dim})"


class IdentityOperator:
    """Counterpart of pyMOR's IdentityOperator(space) (rla/embeddings.py:137-139)."""
    linear = True

    def __init__(self, space):
        self.source = self.range = space

    def apply(self, U, mu=None):
        return U

    def apply_adjoint(self, V, mu=None):
        return V


class MatrixOperator:
    """Dense or CSR matrix operator on the device: counterpart of pyMOR's
    NumpyMatrixOperator for the `sqrt_product` Q and the affine terms A_q.
    `apply(U)` = (M @ U^T)^T, `apply_adjoint(V)` = V @ conj(M)."""
    linear = True

    def __init__(self, matrix, source_id=None, range_id=None):
        torch = _torch()
        import scipy.sparse as sp
        self.sparse = sp.issparse(matrix)
        if self.sparse:
            csr = matrix.tocsr()
            if np.iscomplexobj(csr.data):
                raise TypeError("MatrixOperator: complex sparse matrices are not supported (split real/imaginary parts)")
            csr.sort_indices()
            self.shape = csr.shape
            self.rowptr = torch.from_numpy(csr.indptr.astype(np.int64)).cuda()
            self.col = torch.from_numpy(csr.indices.astype(np.int32)).cuda()
            self.val = torch.from_numpy(csr.data.astype(np.float64)).cuda()
            self._host = csr
            self._t = None
        else:
            self.matrix = as_device_block(matrix)
            self.shape = tuple(self.matrix.shape)
        self.source = DeviceVectorSpace(self.shape[1], source_id)
        self.range = DeviceVectorSpace(self.shape[0], range_id)

    def apply(self, U, mu=None):
        assert U in self.source
        if self.sparse:
            from .reductor_ops import spmm_csr
            return DeviceVectorArray(self.range, spmm_csr(self.rowptr, self.col, self.val, self.shape, U.data))
        from .dense import gauss_apply_explicit, _tma_friendly
        torch = _torch()
        if U.data.dtype == torch.float64 and self.matrix.dtype == torch.float64:
            # (M @ U^T)^T on the tensor-core sketch kernel
            return DeviceVectorArray(self.range, gauss_apply_explicit(_tma_friendly(self.matrix), _tma_friendly(U.data)))
        return DeviceVectorArray(self.range, U.data @ self.matrix.T)

    def apply_adjoint(self, V, mu=None):
        assert V in self.range
        if self.sparse:
            if self._t is None:
                self._t = MatrixOperator(self._host.conj().T.tocsr())
            return DeviceVectorArray(self.source, self._t.apply(DeviceVectorArray(self._t.source, V.data)).data)
        torch = _torch()
        if V.data.dtype == torch.float64 and self.matrix.dtype == torch.float64:
            from .reductor_ops import gemm_nn
            return DeviceVectorArray(self.source, gemm_nn(V.data, self.matrix))
        return DeviceVectorArray(self.source, V.data @ self.matrix.conj())
