"""CPU oracle for the rla4mor sketching hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference's hot path
(`rla/srht.py`, `rla/embeddings.py`, the sketch arithmetic of
`mor/sketched_reductor.py`).  It exists to *check* the CUDA product in
`rla4mor_b200/`; it is never part of the product.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import or execute anything under
`oracle/`.  Nothing in `rla4mor_b200/` imports it (tests/test_no_oracle_in_product.py
enforces that), and the product raises if its CUDA library is missing.

Parity pin: the reference ships no golden vectors for this path (SURVEY.md
section 8c).  The oracle is pinned instead against *outputs of the reference
itself*, generated in the build container and committed under `tests/golden/`:

  * `oracle/make_golden.py` imports `/root/reference/rla/srht.py` by file path and
    records `srht`, `fht_oop`, `fht_ip` on seeded inputs (`srht_reference.npz`,
    `fht_reference.npz`);
  * `oracle/make_golden_pymor.py` imports `/root/reference/rla/embeddings.py`,
    `mor/sketched_reductor.py` and `utilities/*` UNMODIFIED on top of the tests-only pyMOR
    stand-in `tests/_pymor_stub` (pyMOR is not in the image) and records every embedding
    class (apply, adjoint, explicit matrices, blocks, seeds, dims, caching quirks) and the
    whole SketchedReductor flow (galerkin, minres, empty ROM): `embeddings_reference.npz`,
    `reductor_reference.npz`.

`tests/test_oracle_golden.py` and `tests/test_reference_goldens_cpu.py` check every oracle
function against those fixtures; the GPU tests compare the product with the same fixtures.
What stays a restatement: pyMOR's own side of each call (`NumpyMatrixOperator.apply`,
`gram_schmidt`, `project` / `expand` / `contract`), written in the stub from pyMOR 2023.1's
documented semantics.  `embeddings_transcribed.npz` (round 1: the NumPy lines of
`rla/embeddings.py` re-typed) is kept as a second, independent fixture.
"""
from .srht_oracle import (  # noqa: F401
    rademacher_signs, sampling_indices, srht, fht_oop, fht_ip, srht_closed_form,
)
from .embeddings_oracle import (  # noqa: F401
    srht_compute_dim, gaussian_compute_dim, srht_random_rows, srht_apply,
    srht_apply_adjoint, gaussian_random_matrix, gaussian_apply, block_sizes,
    block_seeds, block_gaussian_block, block_gaussian_apply,
    block_gaussian_random_matrix, vectorized_apply,
)
from .reductor_oracle import (  # noqa: F401
    gram_schmidt, sketch_affine_terms, orthonormalize_sketch, galerkin_system,
    residual_norm,
)
