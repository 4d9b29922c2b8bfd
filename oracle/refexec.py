"""Execute selected functions of the mounted reference WITHOUT importing it.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Only usable where
``/root/reference`` exists (the build container); the GPU box gets the golden
vectors this produces (``tests/golden/make_golden.py``), never this module's
output at run time.

The reference modules cannot be imported here (faiss, lance, fastcluster,
spectrum_utils are absent -- SURVEY F3), but their pure-numba helpers can be
run by parsing the file, keeping only the wanted ``FunctionDef`` nodes and
``exec``-ing them with ``cache=True`` stripped from the decorators (SURVEY
F4-iii).  No reference source is copied into this repository.
"""
from __future__ import annotations

import ast
import math
import os

REFERENCE_ROOT = os.environ.get("FALCON_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "falcon/cluster/cluster.py"))


class _StripCache(ast.NodeTransformer):
    def visit_Call(self, node):
        self.generic_visit(node)
        node.keywords = [k for k in node.keywords if k.arg != "cache"]
        return node


def load(rel_path: str, names: list[str], extra_globals: dict | None = None) -> dict:
    """Return ``{name: function}`` for ``names`` defined in ``rel_path``."""
    import numba as nb
    import numpy as np
    import scipy.cluster.hierarchy as sch
    from scipy.cluster.hierarchy import fcluster
    from typing import Dict, Iterator, List, Optional, Tuple, Union

    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path) as fh:
        tree = ast.parse(fh.read(), path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    if missing:
        raise KeyError(f"{sorted(missing)} not found in {path}")
    mod = ast.Module(body=[_StripCache().visit(n) for n in keep], type_ignores=[])
    ast.fix_missing_locations(mod)
    g = dict(
        nb=nb, np=np, math=math, sch=sch, fcluster=fcluster,
        Dict=Dict, Iterator=Iterator, List=List, Optional=Optional, Tuple=Tuple, Union=Union,
    )
    if extra_globals:
        g.update(extra_globals)
    exec(compile(mod, path, "exec"), g)
    return {n: g[n] for n in names}


def spectrum_functions() -> dict:
    """``get_dim`` and ``_to_vector`` of falcon/cluster/spectrum.py:172-199, 250-296."""
    return load("falcon/cluster/spectrum.py", ["get_dim", "_to_vector"])


def cluster_functions() -> dict:
    """``_get_cluster_group_idx``, ``_postprocess_cluster``, ``_linkage`` of
    falcon/cluster/cluster.py:334-509."""
    return load(
        "falcon/cluster/cluster.py",
        ["_get_cluster_group_idx", "_postprocess_cluster", "_linkage"],
    )
