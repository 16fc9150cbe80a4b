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
itself*: `oracle/make_golden.py` imports `/root/reference/rla/srht.py` by file
path in the build container, runs `srht`, `fht_oop`, `fht_ip` on seeded inputs
and commits the results under `tests/golden/`; `tests/test_oracle_golden.py`
checks every oracle function against those fixtures.  The embedding classes in
`rla/embeddings.py` cannot be imported (pyMOR is not installed and is not under
/root/reference); their restatement is pinned by re-deriving each formula with
plain NumPy inside `make_golden.py` from the cited reference lines -- that part
is "parity unpinned by executable reference" and is flagged as such in
DESIGN.md.
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
