"""Modified Gram-Schmidt with re-iteration (pyMOR 2023.1 `gram_schmidt`): R starts as the
identity, a vector is dropped when its norm is <= atol or falls to <= rtol * its initial norm,
and it is orthogonalised again while the norm shrinks below reiteration_threshold * the old one."""
import numpy as np

from pymor.core.exceptions import AccuracyError
from pymor.core.logger import getLogger


def gram_schmidt(A, product=None, return_R=False, atol=1e-13, rtol=1e-13, offset=0,
                 reiterate=True, reiteration_threshold=9e-1, check=True, check_tol=1e-3, copy=True):
    logger = getLogger("pymor.algorithms.gram_schmidt.gram_schmidt")
    if copy:
        A = A.copy()
    R = np.eye(len(A))
    remove = []
    for i in range(offset, len(A)):
        initial_norm = A[i].norm(product)[0]
        if initial_norm <= atol:
            logger.info(f"Removing vector {i} of norm {initial_norm}")
            remove.append(i)
            continue
        if i == 0:
            A[0].scal(1 / initial_norm)
            R[i, i] = initial_norm
        else:
            norm = initial_norm
            while True:
                for j in range(i):
                    if j in remove:
                        continue
                    p = A[j].inner(A[i], product)[0, 0]
                    A[i].axpy(-p, A[j])
                    common_dtype = np.promote_types(R.dtype, type(p))
                    R = R.astype(common_dtype, copy=False)
                    R[j, i] += p
                old_norm, norm = norm, A[i].norm(product)[0]
                if norm <= rtol * initial_norm:
                    logger.info(f"Removing linearly dependent vector {i}")
                    remove.append(i)
                    break
                if reiterate and norm < reiteration_threshold * old_norm:
                    logger.info(f"Orthonormalizing vector {i} again")
                else:
                    A[i].scal(1 / norm)
                    R[i, i] = norm
                    break
    if remove:
        del A[remove]
        R = np.delete(R, remove, axis=0)
    if check:
        error_matrix = A[offset:len(A)].inner(A, product)
        error_matrix[:len(A) - offset, offset:len(A)] -= np.eye(len(A) - offset)
        if error_matrix.size > 0:
            err = np.max(np.abs(error_matrix))
            if err >= check_tol:
                raise AccuracyError(f"result not orthogonal (max err={err})")
    return (A, R) if return_R else A
