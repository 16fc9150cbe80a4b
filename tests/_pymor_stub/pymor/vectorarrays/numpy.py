"""NumpyVectorSpace / NumpyVectorArray: `(len, dim)` row layout (pyMOR 2023.1)."""
import numpy as np

from pymor.vectorarrays.interface import VectorArray, VectorSpace


class NumpyVectorArray(VectorArray):
    def __init__(self, space, data):
        self.space = space
        self._data = data

    # -- basics
    @property
    def dim(self):
        return self.space.dim

    def __len__(self):
        return self._data.shape[0]

    def to_numpy(self, ensure_copy=False):
        return self._data.copy() if ensure_copy else self._data

    def copy(self, deep=False):
        return NumpyVectorArray(self.space, self._data.copy())

    def __getitem__(self, ind):
        if isinstance(ind, (int, np.integer)):
            ind = [ind]
        d = self._data[ind]
        view = NumpyVectorArray(self.space, d)
        view._base, view._ind = self, ind
        return view

    def __delitem__(self, ind):
        self._data = np.delete(self._data, ind, axis=0)

    def _writeback(self):
        # fancy-indexed "views" are copies in NumPy: push in-place changes back to the base
        base = getattr(self, "_base", None)
        if base is not None and not np.shares_memory(self._data, base._data):
            base._data[self._ind] = self._data
            base._writeback()

    def append(self, other, remove_from_other=False):
        assert other in self.space
        dtype = np.promote_types(self._data.dtype, other._data.dtype) if len(self) else other._data.dtype
        self._data = np.concatenate([self._data.astype(dtype, copy=False), other._data.astype(dtype, copy=False)], axis=0)
        if remove_from_other:
            other._data = other._data[:0]

    # -- arithmetic
    def scal(self, alpha):
        alpha = np.asarray(alpha)
        if np.iscomplexobj(alpha) and not np.iscomplexobj(self._data):
            self._data = self._data.astype(complex)
        self._data *= alpha.reshape(-1, 1) if alpha.ndim else alpha
        self._writeback()

    def axpy(self, alpha, x):
        alpha = np.asarray(alpha)
        xd = x._data
        dtype = np.promote_types(np.promote_types(self._data.dtype, xd.dtype), alpha.dtype)
        if dtype != self._data.dtype:
            self._data = self._data.astype(dtype)
        self._data += (alpha.reshape(-1, 1) if alpha.ndim else alpha) * xd
        self._writeback()

    def inner(self, other, product=None):
        if product is not None:
            return product.apply2(self, other)
        return self._data.conj() @ other._data.T

    def pairwise_inner(self, other, product=None):
        if product is not None:
            return product.pairwise_apply2(self, other)
        return np.sum(self._data.conj() * other._data, axis=1)

    def gramian(self, product=None):
        return self.inner(self, product)

    def norm(self, product=None, tol=None, raise_complex=None):
        if product is not None:
            return np.sqrt(product.pairwise_apply2(self, self).real)
        return np.linalg.norm(self._data, axis=1)

    def norm2(self, product=None):
        return self.norm(product) ** 2

    def lincomb(self, coefficients):
        c = np.asarray(coefficients)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        assert c.shape[1] == len(self)
        return NumpyVectorArray(self.space, c @ self._data)

    def conj(self):
        return NumpyVectorArray(self.space, self._data.conj())

    @property
    def real(self):
        return NumpyVectorArray(self.space, self._data.real.copy())

    @property
    def imag(self):
        return NumpyVectorArray(self.space, self._data.imag.copy())

    def __add__(self, other):
        return NumpyVectorArray(self.space, self._data + other._data)

    def __sub__(self, other):
        return NumpyVectorArray(self.space, self._data - other._data)

    def __isub__(self, other):
        self.axpy(-1, other)
        return self

    def __iadd__(self, other):
        self.axpy(1, other)
        return self

    def __mul__(self, alpha):
        return NumpyVectorArray(self.space, self._data * alpha)

    __rmul__ = __mul__

    def __neg__(self):
        return NumpyVectorArray(self.space, -self._data)

    def dofs(self, dof_indices):
        return self._data[:, dof_indices]

    def __repr__(self):
        return f"NumpyVectorArray(len={len(self)}, dim={self.dim}, id={self.space.id!r})"


class _classinstancemethod:
    """Callable on the class (`NumpyVectorSpace.from_numpy(data)`) and on an instance."""

    def __init__(self, cls_fn):
        self.cls_fn = cls_fn
        self.inst_fn = None

    def instancemethod(self, fn):
        self.inst_fn = fn
        return self

    def __get__(self, obj, owner):
        if obj is None:
            return lambda *a, **k: self.cls_fn(owner, *a, **k)
        return lambda *a, **k: self.inst_fn(obj, *a, **k)


class NumpyVectorSpace(VectorSpace):
    def __init__(self, dim, id=None):
        self.dim = int(dim)
        self.id = id

    is_scalar = property(lambda self: self.dim == 1 and self.id is None)

    def __eq__(self, other):
        return type(other) is type(self) and self.dim == other.dim and self.id == other.id

    def __hash__(self):
        return hash((self.dim, self.id))

    def __contains__(self, U):
        return getattr(U, "space", None) == self

    def __repr__(self):
        return f"NumpyVectorSpace({self.dim}, id={self.id!r})"

    @_classinstancemethod
    def from_numpy(cls, data, id=None, ensure_copy=False):
        data = np.asarray(data)
        if data.ndim == 1:
            data = data.reshape(1, -1)
        return NumpyVectorArray(cls(data.shape[1], id), data.copy() if ensure_copy else data)

    @from_numpy.instancemethod
    def from_numpy(self, data, ensure_copy=False):
        data = np.asarray(data)
        if data.ndim == 1:
            data = data.reshape(1, -1)
        assert data.shape[1] == self.dim, f"expected vectors of dimension {self.dim}, got {data.shape[1]}"
        return NumpyVectorArray(self, data.copy() if ensure_copy else data)

    def make_array(self, data):
        return self.from_numpy(data)

    def empty(self, reserve=0):
        return NumpyVectorArray(self, np.empty((0, self.dim)))

    def zeros(self, count=1, reserve=0):
        return NumpyVectorArray(self, np.zeros((count, self.dim)))

    def ones(self, count=1, reserve=0):
        return NumpyVectorArray(self, np.ones((count, self.dim)))

    def full(self, value, count=1, reserve=0):
        return NumpyVectorArray(self, np.full((count, self.dim), value))

    def random(self, count=1, distribution="uniform", random_state=None, seed=None, **kwargs):
        rs = random_state if random_state is not None else np.random.RandomState(seed)
        if distribution == "normal":
            return NumpyVectorArray(self, rs.normal(kwargs.get("loc", 0), kwargs.get("scale", 1), (count, self.dim)))
        return NumpyVectorArray(self, rs.uniform(kwargs.get("low", 0), kwargs.get("high", 1), (count, self.dim)))
