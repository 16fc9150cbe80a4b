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
    """What the device kernels compute, step by step, in schedule order: X[q] is row order[q];
    external column indices are positions, internal ones slots of the row's group.
    kind 0 / 1: every row of the position range subtracts its external entries (and divides by its
    effective diagonal: 1 inside groups); kind 2: x = Dinv y for every group of the range."""
    rowptr, col, val, order, split = p["rowptr"], p["col"], p["val"], p["order"], p["split"]
    n = p["n"]
    X = B[order].copy()
    done = np.zeros(n, dtype=bool)           # final value available
    summed = np.zeros(n, dtype=bool)         # external entries subtracted
    in_group = np.zeros(n, dtype=bool)
    for g in range(p["ngroups"]):
        g0, nr = p["grp_start"][g], p["grp_rows"][g]
        assert 2 <= nr <= 128 and not in_group[g0:g0 + nr].any()
        in_group[g0:g0 + nr] = True
    assert np.all(p["diag_eff"][in_group] == 1.0) and np.all(p["diag_eff"][~in_group] == p["diag"][~in_group])
    resolved_groups = 0
    for lo, hi, kind in zip(p["step_lo"], p["step_hi"], p["step_kind"]):
        assert hi > lo and kind in (0, 1, 2)
        new = {}
        if kind in (0, 1):
            for i in range(lo, hi):
                e0, sp_, e1 = rowptr[i], split[i], rowptr[i + 1]
                assert np.all(done[col[e0:sp_]]) and not summed[i]         # external: earlier steps only
                assert in_group[i] or sp_ == e1                            # single rows have no internal entries
                new[i] = (X[i] - val[e0:sp_] @ X[col[e0:sp_]]) / (1.0 if unit else p["diag_eff"][i])
                summed[i] = True
            for i, v in new.items():
                X[i] = v
                done[i] = not in_group[i]
        else:
            assert lo == resolved_groups                                    # groups are resolved in order, once
            for g in range(lo, hi):
                g0, nr = p["grp_start"][g], p["grp_rows"][g]
                assert summed[g0:g0 + nr].all() and not done[g0:g0 + nr].any()
                Dinv = p["dinv"][p["dinv_ptr"][g]:p["dinv_ptr"][g + 1]].reshape(nr, nr)
                D = np.zeros((nr, nr))
                for q in range(nr):
                    i = g0 + q
                    sp_, e1 = split[i], rowptr[i + 1]
                    assert np.all(col[sp_:e1] < q)                          # internal: earlier slots of the group
                    D[q, col[sp_:e1]] = val[sp_:e1]
                    D[q, q] = 1.0 if unit else p["diag"][i]
                assert np.allclose(Dinv @ D, np.eye(nr), atol=1e-10) and np.all(np.triu(Dinv, 1) == 0)
                X[g0:g0 + nr] = Dinv @ X[g0:g0 + nr]
                done[g0:g0 + nr] = True
            resolved_groups = hi
    assert done.all() and resolved_groups == p["ngroups"]
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


def test_chain_groups_do_not_depend_on_unrelated_rows():
    """A 384-row dependency chain next to 40 independent 3-row chains: the long chain is cut into three
    groups of 128 whatever happens at the same levels elsewhere (round 1 cut bands of levels short
    as soon as ANY connected component of the band outgrew a group)."""
    n_chain, n_small = 384, 40
    n = n_chain + 3 * n_small
    rows, cols = [], []
    for i in range(1, n_chain):
        rows += [i] * min(i, 5); cols += list(range(max(0, i - 5), i))          # banded chain: each row needs the 5 before it
    for c in range(n_small):
        b = n_chain + 3 * c
        rows += [b + 1, b + 2]; cols += [b, b + 1]
    T = (sp.csr_matrix((np.full(len(rows), 0.3), (rows, cols)), shape=(n, n)) + 2.0 * sp.eye(n)).tocsr()
    p = plan_triangular(T, True)
    sizes = sorted(p["grp_rows"][:p["ngroups"]].tolist(), reverse=True)
    assert sizes[:3] == [128, 128, 128] and sizes[3:] == [3] * n_small
    assert int((p["step_kind"] == 2).sum()) == 3          # three group levels
    B = np.random.RandomState(1).standard_normal((n, 2))
    assert np.allclose(_replay(p, B, False), spsolve_triangular(T, B, lower=True))


def test_plan_small_cases():
    # diagonal matrix: one level; strictly sequential chain: one group after another
    p = plan_triangular(sp.eye(5, format="csr") * 2.0, True)
    assert p["nlevels"] == 1 and p["nnz"] == 0 and np.allclose(p["diag"], 2.0)
    n = 70
    T = (sp.eye(n) + sp.diags([np.ones(n - 1)], [-1])).tocsr()
    p = plan_triangular(T, True)
    # one chain of 70 rows: ONE group; one external-sum step (nothing to subtract) and one resolve step
    assert p["nlevels"] == n and list(p["step_kind"]) == [0, 2]
    assert list(p["grp_rows"][:p["ngroups"]]) == [70]
    B = np.arange(n, dtype=float).reshape(-1, 1)
    assert np.allclose(_replay(p, B, True), spsolve_triangular(T, B, lower=True))
    with pytest.raises(Exception):
        plan_triangular(sp.csr_matrix(np.array([[1.0, 2.0], [0.0, 1.0]])), True)       # entry above the diagonal
