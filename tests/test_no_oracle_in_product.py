"""The product never routes through the test oracle or any CPU fallback: walk the AST of every
module under rla4mor_b200/ and reject imports of `oracle` / `tests`, and any string that points
into oracle/ (a ctypes load of oracle/_build, a subprocess of oracle/_ref).  bench.py and
__graft_entry__.py may use the oracle only in their checker / cpu_baseline legs (tested in
test_bench_contract.py)."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rla4mor_b200")


def product_files():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(dirpath, f)


def test_no_oracle_import_or_path_in_product():
    offenders = []
    for path in product_files():
        tree = ast.parse(open(path).read(), filename=path)
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                top = n.split(".")[0]
                if top in ("oracle", "tests", "golden_util"):
                    offenders.append((path, node.lineno, f"import {n}"))
            if isinstance(node, ast.Constant) and isinstance(node.value, str):
                v = node.value
                if "oracle/" in v or "oracle." in v or "_pymor_stub" in v or "/root/reference" in v:
                    # docstrings may mention the oracle; code strings may not
                    parent_is_doc = False
                    for p in ast.walk(tree):
                        body = getattr(p, "body", None)
                        if isinstance(body, list) and body and isinstance(body[0], ast.Expr) and body[0].value is node:
                            parent_is_doc = True
                    if not parent_is_doc:
                        offenders.append((path, node.lineno, v[:60]))
    assert not offenders, offenders


def test_cuda_sources_do_not_reference_the_oracle():
    """No #include of, or path into, oracle/ in the CUDA sources (comments may mention the word)."""
    import re
    for dirpath, _, files in os.walk(os.path.join(PKG, "csrc")):
        for f in files:
            if f.endswith((".cu", ".cuh")):
                for line in open(os.path.join(dirpath, f)):
                    code = line.split("//")[0]
                    assert not re.search(r"oracle", code, re.I), (f, line)


def test_product_fails_loudly_without_the_library(monkeypatch):
    """A missing shared library is an error, not a fallback."""
    import pytest
    from rla4mor_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    import rla4mor_b200._build as b
    monkeypatch.setattr(b, "build_library", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no nvcc")))
    with pytest.raises(_lib.RlaError):
        _lib.lib()
