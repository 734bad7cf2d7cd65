"""Importable alias of ``efficient-rag-with-learned-retrieval-and-uncertainty-quantification_b200/``.

The product package directory is named after the reference repository and contains hyphens,
so it cannot be imported by name.  This shim makes ``import rag_uq_b200`` (and
``rag_uq_b200.ops`` etc.) resolve into that directory; it holds no code of its own.
"""
from pathlib import Path as _Path

_REAL = _Path(__file__).resolve().parent.parent / "efficient-rag-with-learned-retrieval-and-uncertainty-quantification_b200"
__path__ = [str(_REAL)]
exec(compile((_REAL / "__init__.py").read_text(), str(_REAL / "__init__.py"), "exec"))
