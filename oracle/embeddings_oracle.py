"""CPU restatement of /root/reference/rla/embeddings.py -- TEST INFRASTRUCTURE ONLY.

pyMOR (unpinned, README.md:8 of the reference) is not installed and not under
/root/reference, so the operator classes cannot be executed; what is restated
here is their *arithmetic* on plain NumPy arrays in the reference's layout:
a block of m vectors of dimension n is an (m, n) array, one vector per row
(`VectorArray.to_numpy()`, embeddings.py:169).  pyMOR's
`NumpyMatrixOperator(M).apply(V)` is `(M @ V.to_numpy().T).T` = `V @ M.T`
(published pyMOR behaviour, cited at embeddings.py:177-178,253-254,431-432).
`sqrt_product` is the identity unless a dense `Q` (n_range x n_source) is given.
"""
import numpy as np

from .srht_oracle import fht_oop, srht


# ----------------------------------------------------------------- dimensions
def srht_compute_dim(options, n):
    """SrhtEmbedding.compute_dim, embeddings.py:148-164."""
    range_dim = options.get('range_dim')
    eps = options.get('epsilon')
    delta = options.get('delta')
    d = options.get('oblivious_dim')
    assert range_dim or all([eps, delta, d])                         # :154
    if range_dim is None:
        a = 2 if options.get('dtype') == complex else 1              # :157-159
        range_dim = 2 / (eps ** 2 - eps ** 3 / 3)                    # :160
        range_dim = range_dim * (np.sqrt(a * d) + np.sqrt(8 * np.log(6 * a * n / delta))) ** 2  # :161
        range_dim = range_dim * np.log(3 * a * d / delta)            # :162
        range_dim = int(np.ceil(range_dim))                          # :163
    return range_dim


def gaussian_compute_dim(options):
    """Gaussian / BlockGaussian / Vectorized compute_dim,
    embeddings.py:234-247, 337-350, 409-422."""
    range_dim = options.get('range_dim')
    eps = options.get('epsilon')
    delta = options.get('delta')
    d = options.get('oblivious_dim')
    assert range_dim or all([eps, delta, d])                         # :240
    if range_dim is None:
        a = 2 if options.get('dtype') == complex else 1
        range_dim = 7.87 * (1 / eps ** 2) * (a * 6.9 * d + np.log(1 / delta))  # :245
        range_dim = int(np.ceil(range_dim))
    return range_dim


# ----------------------------------------------------------------------- SRHT
def srht_apply(U, k, seed, Q=None):
    """SrhtEmbedding.apply, embeddings.py:167-172: srht(Q U)."""
    qu = U if Q is None else U @ Q.T
    return srht(qu, k, seed)


def srht_random_rows(n, k, seed, indices):
    """SrhtEmbedding._get_random_rows, embeddings.py:195-209 (note the
    sqrt(n/k) scale, which differs from srht()'s sqrt(2**d/k) when n is not a
    power of two -- SURVEY.md App. A.3)."""
    d = int(np.ceil(np.log2(n)))                                              # :198
    rademacher = np.random.RandomState(seed).choice([-1, 1], (n), replace=True)   # :201
    sampling = np.random.RandomState(seed).choice(2 ** d, k, replace=True)        # :202
    Pt = np.zeros((len(indices), 2 ** d))                                     # :204
    for i, ind in enumerate(indices):
        Pt[i, sampling[ind]] = 1                                              # :205-206
    Pt = fht_oop(Pt)                                                          # :207
    return np.sqrt(n / k) * Pt[:, :n] * rademacher                            # :208


def srht_matrix(n, k, seed, Q=None):
    """get_random_matrix / get_matrix of SrhtEmbedding, embeddings.py:182-192:
    rows 0..k-1, then Q^H applied to each row (`Q.apply_adjoint`)."""
    rmat = srht_random_rows(n, k, seed, np.arange(k))
    return rmat if Q is None else rmat @ Q.conj()


def srht_apply_adjoint(V, n, k, seed, Q=None):
    """SrhtEmbedding.apply_adjoint, embeddings.py:175-178:
    NumpyMatrixOperator(get_matrix().T).apply(V) = V @ get_matrix()."""
    return V @ srht_matrix(n, k, seed, Q)


# ------------------------------------------------------------------- Gaussian
def gaussian_random_matrix(k, n, seed):
    """GaussianEmbedding._compute_random_matrix, embeddings.py:265-270."""
    return np.random.RandomState(seed).normal(size=(k, n), loc=0, scale=1 / np.sqrt(k))


def gaussian_apply(U, theta, Q=None):
    """GaussianEmbedding.apply, embeddings.py:250-254: (Theta @ (Q U)^T)^T."""
    qu = U if Q is None else U @ Q.T
    return (theta @ qu.T).T


def gaussian_matrix(theta, Q=None):
    """GaussianEmbedding._compute_matrix, embeddings.py:258-262."""
    return theta if Q is None else (theta.conj() @ Q.conj()).conj()


# -------------------------------------------------------------- Block Gaussian
def block_sizes(k, max_block_size):
    """embeddings.py:394-400."""
    m = k // max_block_size
    r = k % max_block_size
    sizes = [max_block_size for _ in range(m)]
    if r > 0:
        sizes.append(r)
    return sizes


def block_seeds(seed, n_blocks):
    """embeddings.py:403-407: per-block seeds; `_seed` is bumped until they are
    unique.  Returns (block_seeds, possibly incremented seed)."""
    seeds = np.random.RandomState(seed).randint(0, 2 ** 32 - 1, size=n_blocks)
    while len(np.unique(seeds)) < len(seeds):
        seed += 1
        seeds = np.random.RandomState(seed).randint(0, 2 ** 32 - 1, size=n_blocks)
    return seeds, seed


def block_gaussian_block(k_total, n, b, block_seed):
    """BlockGaussianEmbedding._get_random_block, embeddings.py:452-461: every
    block uses scale 1/sqrt(k_total)."""
    return np.random.RandomState(block_seed).normal(size=(b, n), loc=0, scale=1 / np.sqrt(k_total))


def block_gaussian_random_matrix(k, n, seed, max_block_size):
    """embeddings.py:444-450."""
    sizes = block_sizes(k, max_block_size)
    seeds, _ = block_seeds(seed, len(sizes))
    return np.vstack([block_gaussian_block(k, n, b, s) for b, s in zip(sizes, seeds)])


def block_gaussian_apply(U, k, seed, max_block_size, Q=None):
    """BlockGaussianEmbedding.apply, embeddings.py:425-434: one GEMM per row
    block of Theta, results hstack'ed along the sketch dimension."""
    V = U if Q is None else U @ Q.T
    n = V.shape[1]
    sizes = block_sizes(k, max_block_size)
    seeds, _ = block_seeds(seed, len(sizes))
    lst = []
    for b, s in zip(sizes, seeds):
        gauss = block_gaussian_block(k, n, b, s)
        lst.append((gauss @ V.T).T)
    return np.hstack(lst)


# ------------------------------------------------------------------ Vectorized
def vectorized_apply(U, inner_apply):
    """EmbeddingVectorized.apply, embeddings.py:352-358: x = U.to_numpy().T.flatten()
    (i.e. the (n_vectors, k1) block flattened column-major), then the inner
    embedding is applied to that single vector."""
    x = U.T.flatten()
    return inner_apply(x.reshape(1, -1))
