"""Host logic of the sparse triangular solves (rla4mor_b200/factorization.py, csrc/sptrsv.cu):
the C++ plan builder (levels, level-sorted order, groups of the narrow tail, re-packed CSR) is
checked on the CPU by replaying the device schedule step by step in NumPy against SciPy.
No GPU involved."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import splu, spsolve_triangular

from rla4mor_b200.factorization import plan_triangular


def _fem(nx):
    ex = np.ones(nx)
    T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
    return (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx)) + 0.1 * sp.eye(nx * nx)).tocsc()


def _replay(p, B, unit):
    """What trsv_wide_kernel / trsv_groups_kernel compute, step by step, in schedule order:
    X[q] is row order[q]; external column indices are positions, internal ones slots."""
    rowptr, col, val, order, split = p["rowptr"], p["col"], p["val"], p["order"], p["split"]
    X = B[order].copy()
    done = np.zeros(p["n"], dtype=bool)
    for lo, mid, hi, kind in zip(p["step_lo"], p["step_mid"], p["step_hi"], p["step_kind"]):
        new = {}
        if kind == 0:
            for i in range(lo, hi):
                e0, e1 = rowptr[i], rowptr[i + 1]
                assert split[i] == e1 and np.all(done[col[e0:e1]])        # only rows of earlier steps
                new[i] = (X[i] - val[e0:e1] @ X[col[e0:e1]]) / (1.0 if unit else p["diag"][i])
        else:
            assert hi - lo >= 1 and mid - lo <= p["max_multi"]
            for g in range(lo, hi):
                g0, nr = p["grp_start"][g], p["grp_rows"][g]
                assert 1 <= nr <= 32 and (nr > 1) == (g < mid)
                if g > lo:
                    assert g0 == p["grp_start"][g - 1] + p["grp_rows"][g - 1]   # a step is contiguous in the order
                xs = np.zeros((nr,) + X.shape[1:])
                for q in range(nr):
                    i = g0 + q
                    e0, sp, e1 = rowptr[i], split[i], rowptr[i + 1]
                    assert np.all(done[col[e0:sp]])                       # external: earlier steps only
                    assert np.all(col[sp:e1] < q)                         # internal: earlier slots of the group
                    acc = X[i] - val[e0:sp] @ X[col[e0:sp]] - val[sp:e1] @ xs[col[sp:e1]]
                    xs[q] = acc / (1.0 if unit else p["diag"][i])
                    new[i] = xs[q]
        for i, v in new.items():                                          # a step only becomes visible at its end
            X[i] = v
            done[i] = True
    assert done.all()
    out = np.empty_like(X)
    out[order] = X
    return out


@pytest.mark.parametrize("nx", [7, 40])
def test_plan_replay_matches_scipy(nx):
    A = _fem(nx)
    n = A.shape[0]
    lu = splu(A)
    B = np.random.RandomState(0).standard_normal((n, 3))
    for T, lower, unit in ((lu.L, True, True), (lu.U, False, False), (lu.U.T, True, False), (lu.L.T, False, True)):
        p = plan_triangular(T.tocsr(), lower)
        assert sorted(p["order"].tolist()) == list(range(n)) and np.all(p["pos"][p["order"]] == np.arange(n))
        ref = spsolve_triangular(T.tocsr(), B, lower=lower)
        got = _replay(p, B, unit)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-12


def test_plan_small_cases():
    # diagonal matrix: one level; strictly sequential chain: one group after another
    p = plan_triangular(sp.eye(5, format="csr") * 2.0, True)
    assert p["nlevels"] == 1 and p["nnz"] == 0 and np.allclose(p["diag"], 2.0)
    n = 70
    T = (sp.eye(n) + sp.diags([np.ones(n - 1)], [-1])).tocsr()
    p = plan_triangular(T, True)
    assert p["nlevels"] == n and p["nsteps"] == 3 and list(p["step_kind"]) == [1, 1, 1]
    assert list(p["grp_rows"][:p["ngroups"]]) == [32, 32, 6]
    B = np.arange(n, dtype=float).reshape(-1, 1)
    assert np.allclose(_replay(p, B, True), spsolve_triangular(T, B, lower=True))
    with pytest.raises(Exception):
        plan_triangular(sp.csr_matrix(np.array([[1.0, 2.0], [0.0, 1.0]])), True)       # entry above the diagonal
