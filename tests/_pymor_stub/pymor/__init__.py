"""TEST INFRASTRUCTURE -- a minimal stand-in for the parts of pyMOR that rla4mor's hot path touches.

pyMOR (unpinned in the reference, README.md:8; the API used is the 2023.1 one) is not in the
build image, so `/root/reference/rla/embeddings.py`, `mor/sketched_reductor.py` and `utilities/*`
cannot be imported as they are.  This package restates, from pyMOR's published behaviour, exactly
the names SURVEY.md Appendix B lists (BasicObject/ImmutableObject with `__auto_init` / `with_` /
`logger`, Operator algebra, NumpyVectorSpace/Array, NumpyMatrixOperator, rule tables with
`insert_rule`, `project` / `expand` / `contract`, `gram_schmidt`, StationaryModel,
ResidualOperator) so that the UNMODIFIED reference modules run on top of it:

  * `oracle/make_golden.py` puts this directory and /root/reference on sys.path, imports the
    reference classes themselves and records their outputs as fixtures under tests/golden/;
  * `tests/test_pymor_adapter.py` runs `rla4mor_b200.pymor_adapter` against it.

It is never imported by the product (`rla4mor_b200/`), and nothing here is tuned for speed.
What is pinned by this route: every line of the reference's own modules.  What stays a
restatement: pyMOR's side of each call (`NumpyMatrixOperator.apply`, `gram_schmidt`, `project`,
`expand`, `contract`), written here from its documented semantics.
"""
__version__ = "0.stub (API of pyMOR 2023.1)"
