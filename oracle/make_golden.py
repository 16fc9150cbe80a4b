"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (build container only).

    python -m oracle.make_golden

Imports /root/reference/rla/srht.py by file path (oracle/ref_loader.py) and
records, for seeded inputs, the outputs of the reference's own `srht`,
`fht_oop` and `fht_ip`, plus the sign / row-index draws written exactly as the
reference writes them (srht.py:162-163).  For the embedding classes (pyMOR is
absent, so rla/embeddings.py cannot be imported) the fixtures are produced by
transcribing the cited NumPy expressions line by line here, with the
reference's own fht_oop doing the transform -- see each block's comment.
The fixtures are committed; the GPU box never needs /root/reference.
"""
import os

import numpy as np

from .ref_loader import load_reference_srht

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, m or None for 1-D, n, k, embedding seed, complex?)
SRHT_CASES = [
    ("pow2_small", 3, 64, 10, 0, False),
    ("nonpow2_100", 5, 100, 16, 7, False),
    ("vec1d_37", None, 37, 8, 3, False),
    ("one_tile_4096", 2, 4096, 100, 1, False),
    ("nonpow2_5000", 2, 5000, 64, 5, False),
    ("two_tiles_8197", 4, 8197, 50, 11, False),
    ("k_gt_n", 2, 16, 40, 9, False),
    ("n_is_1", 2, 1, 3, 4, False),
    ("complex_48", 2, 48, 6, 2, True),
    ("mid_65536", 2, 65536, 200, 0, False),
]
FHT_SHAPES = [(8,), (1,), (3, 16), (3, 32), (2, 1024), (2, 2048), (4, 1), (1, 4096), (2, 8192)]


def _input(shape, seed, cplx=False):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal(shape)
    if cplx:
        x = x + 1j * rs.standard_normal(shape)
    return x


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference_srht()

    srht_out = {}
    for name, m, n, k, seed, cplx in SRHT_CASES:
        shape = (n,) if m is None else (m, n)
        x = _input(shape, 1000 + seed, cplx)
        y = ref.srht(x, k, seed=seed)
        d = int(np.ceil(np.log2(n)))
        signs = np.random.RandomState(seed).choice([-1, 1], (n), True)        # srht.py:162
        sampling = np.random.RandomState(seed).choice(range(2 ** d), k, True)  # srht.py:163
        srht_out[name + "__x"] = x if x.size <= 20000 else np.zeros(0)        # big inputs are regenerated from the seed
        srht_out[name + "__y"] = y
        srht_out[name + "__signs"] = signs.astype(np.int8)
        srht_out[name + "__sampling"] = sampling.astype(np.int64)
        srht_out[name + "__meta"] = np.array([-1 if m is None else m, n, k, seed, int(cplx)], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "srht_reference.npz"), **srht_out)

    fht_out = {}
    for i, shape in enumerate(FHT_SHAPES):
        for cplx in (False, True):
            if cplx and i not in (2, 3):
                continue
            a = _input(shape, 2000 + i, cplx)
            tag = f"case{i}_{'c' if cplx else 'r'}"
            fht_out[tag + "__a"] = a
            fht_out[tag + "__oop"] = ref.fht_oop(a)
            b = a.copy()
            ref.fht_ip(b)
            fht_out[tag + "__ip"] = b
    np.savez_compressed(os.path.join(OUT, "fht_reference.npz"), **fht_out)

    # ---- embeddings.py transcriptions (pyMOR absent; NumPy lines copied by citation)
    emb = {}
    # SrhtEmbedding._get_random_rows, embeddings.py:195-209, with the reference's fht_oop
    for tag, n, k, seed, indices in [("rows_pow2", 64, 10, 0, np.arange(10)),
                                     ("rows_nonpow2", 100, 16, 7, np.array([0, 3, 15, 3])),
                                     ("rows_5000", 5000, 12, 5, np.arange(12))]:
        d = int(np.ceil(np.log2(n)))
        rademacher = np.random.RandomState(seed).choice([-1, 1], (n), replace=True)
        sampling = np.random.RandomState(seed).choice(range(2 ** d), k, replace=True)
        Pt = np.zeros((len(indices), 2 ** d))
        for i, ind in enumerate(indices):
            Pt[i, sampling[ind]] = 1
        Pt = ref.fht_oop(Pt)
        emb[tag + "__rows"] = np.sqrt(n / k) * Pt[:, :n] * rademacher
        emb[tag + "__meta"] = np.array([n, k, seed], dtype=np.int64)
        emb[tag + "__indices"] = np.asarray(indices, dtype=np.int64)
    # GaussianEmbedding._compute_random_matrix (:265-270) and apply (:250-254)
    for tag, m, n, k, seed in [("gauss_small", 4, 50, 12, 3), ("gauss_mid", 7, 301, 33, 21)]:
        theta = np.random.RandomState(seed).normal(size=(k, n), loc=0, scale=1 / np.sqrt(k))
        U = _input((m, n), 3000 + seed)
        emb[tag + "__theta"] = theta
        emb[tag + "__U"] = U
        emb[tag + "__Y"] = (theta @ U.T).T
        emb[tag + "__meta"] = np.array([m, n, k, seed], dtype=np.int64)
    # BlockGaussianEmbedding block sizes / seeds / blocks (:393-407, :452-461)
    for tag, m, n, k, seed, mbs in [("block_a", 3, 40, 10, 5, 4), ("block_b", 2, 64, 33, 77, 32)]:
        q, r = k // mbs, k % mbs
        sizes = [mbs for _ in range(q)] + ([r] if r > 0 else [])
        seeds = np.random.RandomState(seed).randint(0, 2 ** 32 - 1, size=len(sizes))
        assert len(np.unique(seeds)) == len(seeds)
        blocks = [np.random.RandomState(s).normal(size=(b, n), loc=0, scale=1 / np.sqrt(k))
                  for b, s in zip(sizes, seeds)]
        U = _input((m, n), 4000 + seed)
        emb[tag + "__sizes"] = np.array(sizes, dtype=np.int64)
        emb[tag + "__seeds"] = np.asarray(seeds, dtype=np.int64)
        emb[tag + "__theta"] = np.vstack(blocks)
        emb[tag + "__U"] = U
        emb[tag + "__Y"] = np.hstack([(g @ U.T).T for g in blocks])
        emb[tag + "__meta"] = np.array([m, n, k, seed, mbs], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "embeddings_transcribed.npz"), **emb)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
