"""ORACLE support (test infrastructure): run the UNMODIFIED reference Python from
/root/reference in the build container, to validate `oracle/rerank.py` and to
generate the golden vectors under tests/golden/ (see oracle/make_golden.py).

/root/reference does not exist on the GPU box; there the byte-identical copies
that `oracle/build_ref.py` placed under the git-ignored `oracle/_ref/` are loaded
instead (only by bench.py's reference arm / cpu_baseline leg -- the `-m gpu`
tests and `smoke()` use the committed fixtures).  `available()` says whether
either is present.

`import clustercontrast` fails in this image (matplotlib, wandb, faiss are
missing: clustercontrast/__init__.py:3-8 -> trainers.py:11,13,
utils/faiss_rerank.py:14), so the two files on the path are loaded by file path
under stub parent packages, with a stub `faiss` module whose IndexFlatL2 is an
exact search using the canonical key of oracle/rerank.py (faiss itself is an
un-vendored third-party dependency).  Lines faiss_rerank.py:65-123 then execute
verbatim.
"""
import importlib.util
import os
import sys
import types

import numpy as np

_REF_TREE = "/root/reference/cluster-contrast-reid-main"
_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py (git-ignored)
# the reference tree itself in the build container; on the GPU box the byte-identical copies under oracle/_ref
REF_ROOT = _REF_TREE if os.path.isfile(os.path.join(_REF_TREE, "clustercontrast/utils/faiss_rerank.py")) else _REF_COPY


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "clustercontrast/utils/faiss_rerank.py"))


class _IndexFlatL2:
    """Stand-in for faiss.IndexFlatL2 (call sites faiss_utils.py:108-109,
    faiss_rerank.py:60-62): exact k nearest by L2, ascending; ties by index."""

    precomputed = None        # (x, rank): bench.py times the search on its own and hands the result in here

    def __init__(self, d):
        self.d = d
        self._xb = None

    def add(self, xb):
        self._xb = np.ascontiguousarray(xb, dtype=np.float32)

    def search(self, xq, k):
        from oracle.rerank import exact_knn
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        if xq.shape != self._xb.shape or not np.array_equal(xq, self._xb):
            raise NotImplementedError("stub only supports self-search (the reference's use)")
        pre = type(self).precomputed
        if pre is not None and pre[0].shape == xq.shape and pre[1].shape[1] == k and np.array_equal(pre[0], xq):
            return np.zeros(pre[1].shape, np.float32), pre[1].astype(np.int64)   # the reference discards the distances (:62)
        idx, key = exact_knn(self._xb, k, return_keys=True)
        return (2.0 - 2.0 * key).astype(np.float32), idx


class _IndexFlatIP:
    """Stand-in for faiss.IndexFlatIP (call site infomap_cluster.py:70-73): exact k largest inner products,
    descending; ties by index."""

    def __init__(self, d):
        self.d = d
        self._xb = None

    def add(self, xb):
        self._xb = np.ascontiguousarray(xb, dtype=np.float32)

    def search(self, xq, k):
        from oracle.rerank import exact_knn
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        if xq.shape != self._xb.shape or not np.array_equal(xq, self._xb):
            raise NotImplementedError("stub only supports self-search (the reference's use)")
        idx, key = exact_knn(self._xb, k, return_keys=True)
        return key.astype(np.float32), idx


def _stub_faiss():
    m = types.ModuleType("faiss")
    m.get_num_gpus = lambda: 0
    m.IndexFlatL2 = _IndexFlatL2
    m.IndexFlatIP = _IndexFlatIP
    m.METRIC_L2 = 1
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_faiss_rerank():
    """The reference module clustercontrast.utils.faiss_rerank, unmodified."""
    if "rerank" in _cache:
        return _cache["rerank"]
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    sys.modules.setdefault("faiss", _stub_faiss())
    for pkg, sub in (("clustercontrast", "clustercontrast"), ("clustercontrast.utils", "clustercontrast/utils")):
        if pkg not in sys.modules:
            p = types.ModuleType(pkg)
            p.__path__ = [os.path.join(REF_ROOT, sub)]
            sys.modules[pkg] = p
    _load("clustercontrast.utils.faiss_utils", os.path.join(REF_ROOT, "clustercontrast/utils/faiss_utils.py"))
    mod = _load("clustercontrast.utils.faiss_rerank", os.path.join(REF_ROOT, "clustercontrast/utils/faiss_rerank.py"))
    _cache["rerank"] = mod
    return mod


def load_infomap_cluster():
    """The reference module clustercontrast.utils.infomap_cluster, unmodified (the `infomap` package it imports at
    module level is stubbed: only the kNN front end and get_links are exercised)."""
    if "infomap" in _cache:
        return _cache["infomap"]
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    sys.modules.setdefault("faiss", _stub_faiss())
    if not hasattr(sys.modules["faiss"], "IndexFlatIP"):
        sys.modules["faiss"].IndexFlatIP = _IndexFlatIP
    sys.modules.setdefault("infomap", types.ModuleType("infomap"))
    for pkg, sub in (("clustercontrast", "clustercontrast"), ("clustercontrast.utils", "clustercontrast/utils")):
        if pkg not in sys.modules:
            p = types.ModuleType(pkg)
            p.__path__ = [os.path.join(REF_ROOT, sub)]
            sys.modules[pkg] = p
    _load("clustercontrast.utils.infomap_utils", os.path.join(REF_ROOT, "clustercontrast/utils/infomap_utils.py"))
    mod = _load("clustercontrast.utils.infomap_cluster", os.path.join(REF_ROOT, "clustercontrast/utils/infomap_cluster.py"))
    _cache["infomap"] = mod
    return mod


def load_eval_rerank():
    """The reference module clustercontrast/utils/rerank.py (numpy only), unmodified."""
    if "eval_rerank" not in _cache:
        if not available():
            raise RuntimeError("reference tree not present at " + REF_ROOT)
        _cache["eval_rerank"] = _load("_ref_eval_rerank", os.path.join(REF_ROOT, "clustercontrast/utils/rerank.py"))
    return _cache["eval_rerank"]


def load_ranking():
    """The reference module clustercontrast/evaluation_metrics/ranking.py, unmodified (its relative import
    `from ..utils import to_numpy` is served by the real clustercontrast/utils/__init__.py)."""
    if "ranking" in _cache:
        return _cache["ranking"]
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    for pkg, sub in (("clustercontrast", "clustercontrast"), ("clustercontrast.utils", "clustercontrast/utils"),
                     ("clustercontrast.evaluation_metrics", "clustercontrast/evaluation_metrics")):
        if pkg not in sys.modules:
            p = types.ModuleType(pkg)
            p.__path__ = [os.path.join(REF_ROOT, sub)]
            sys.modules[pkg] = p
    if not hasattr(sys.modules["clustercontrast.utils"], "to_numpy"):
        real = _load("_ref_utils_init", os.path.join(REF_ROOT, "clustercontrast/utils/__init__.py"))
        sys.modules["clustercontrast.utils"].to_numpy = real.to_numpy
        sys.modules["clustercontrast.utils"].to_torch = real.to_torch
    mod = _load("clustercontrast.evaluation_metrics.ranking", os.path.join(REF_ROOT, "clustercontrast/evaluation_metrics/ranking.py"))
    _cache["ranking"] = mod
    return mod


def load_cm():
    """The reference module clustercontrast/models/cm.py (torch + numpy only)."""
    if "cm" not in _cache:
        if not available():
            raise RuntimeError("reference tree not present at " + REF_ROOT)
        _cache["cm"] = _load("_ref_cm", os.path.join(REF_ROOT, "clustercontrast/models/cm.py"))
    return _cache["cm"]


def ref_compute_jaccard_distance(x_torch, k1, k2):
    """faiss_rerank.py:30 with search_option=3 (CPU branch :58-62)."""
    import torch
    torch.set_num_threads(1)            # per-row torch.mm thrashes with more
    mod = load_faiss_rerank()
    return mod.compute_jaccard_distance(x_torch, k1=k1, k2=k2, print_flag=False, search_option=3)


def ref_generate_cluster_features(labels, features):
    """The closure at examples/cluster_contrast_train_usl.py:169-182 cannot be imported
    (it is defined inside main_worker); this is the harness twin SURVEY.md section 8c allows,
    followed by F.normalize (:191)."""
    import collections
    import torch
    import torch.nn.functional as F
    centers = collections.defaultdict(list)
    for i, label in enumerate(labels):
        if label == -1:
            continue
        centers[labels[i]].append(features[i])
    centers = [torch.stack(centers[idx], dim=0).mean(0) for idx in sorted(centers.keys())]
    return F.normalize(torch.stack(centers, dim=0), dim=1)
