"""Import shims that let the UNMODIFIED reference run in the authoring container.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py; never by the product).  None of
the three patches touches hot-path arithmetic:

  1. ``dgl`` is not installed; the reference only uses ``dgl.heterograph(...)`` and
     ``graph.adj_external(etype='bought', scipy_fmt='coo')`` (dataset.py:140-149), which
     must return a scipy COO of ones with rows = user, cols = item.
  2. ``sentence_transformers`` is imported by utils.py:8 but only called when an embedding
     cache file is missing.
  3. ``np.NINF`` (base_model.py:258) was removed in NumPy 2.
"""
import sys
import types

import numpy as np
import scipy.sparse as sp

import os

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    """The unmodified reference tree: /root/reference in the authoring container, else the copy that
    ``__graft_entry__.build()`` leaves in the git-ignored ``baseline/_ref/`` (it travels to the GPU box)."""
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "TextGCN")):
            return cand
    return None


REFERENCE_ROOT = reference_root() or "/root/reference"


class _HeteroGraph:
    def __init__(self, data, device=None):
        self._data = data
        self.ndata = {}

    def adj_external(self, etype=None, scipy_fmt="coo", ctx=None):
        for (src, name, dst), (s, d) in self._data.items():
            if name == etype:
                s = np.asarray(s, dtype=np.int64)
                d = np.asarray(d, dtype=np.int64)
                shape = (int(s.max()) + 1, int(d.max()) + 1)
                return sp.coo_matrix((np.ones(len(s), dtype=np.int64), (s, d)), shape=shape)
        raise KeyError(etype)


def install():
    if "dgl" not in sys.modules:
        dgl = types.ModuleType("dgl")
        dgl.heterograph = lambda data, device=None: _HeteroGraph(data, device)
        sys.modules["dgl"] = dgl
    if "sentence_transformers" not in sys.modules:
        st = types.ModuleType("sentence_transformers")

        class SentenceTransformer:  # pragma: no cover - never called
            def __init__(self, *a, **k):
                raise RuntimeError("sentence_transformers is shimmed; provide embedding cache files")

        st.SentenceTransformer = SentenceTransformer
        sys.modules["sentence_transformers"] = st
    if not hasattr(np, "NINF"):
        np.NINF = -np.inf
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "TextGCN")):
        raise RuntimeError("the reference is neither mounted at /root/reference nor copied to baseline/_ref "
                           "(run `python -c 'import __graft_entry__ as g; g.build()'` where it is mounted)")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
