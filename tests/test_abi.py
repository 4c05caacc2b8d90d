"""CPU: the C-ABI library loads and exports every symbol include/reid_b200.h declares; the Python
surface mirrors the reference's signatures and fails loudly without a GPU (no CPU fallback)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "reid_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(reid_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from reid_gan_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), "missing export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "python binding table and header disagree"
    L = _lib.lib()
    assert L.reid_abi_version() >= 1
    assert L.reid_dbscan_workspace_bytes(1000) > 0          # pure host helpers are callable without a GPU
    assert L.reid_knn_exact_scratch_bytes(1000, 2) == 8000


def test_argument_validation_needs_no_gpu():
    from reid_gan_b200 import _lib
    L = _lib.lib()
    rc = L.reid_reciprocal_masks(None, 10, 5, 5, 0, 10, None, None)
    assert rc == _lib.REID_ERR_INVALID_ARG and "NULL" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "reid_reciprocal_masks")
    buf = (ctypes.c_int32 * 4)()
    rc = L.reid_knn_exact(ctypes.addressof(buf), 10, 4, None, 0, 1, 500, ctypes.addressof(buf), None,
                          ctypes.addressof(buf), 16, None)
    assert rc == _lib.REID_ERR_INVALID_ARG


def test_argument_validation_of_the_sample_first_entries():
    """The entry points of the sample-first symmetric search reject inconsistent optional arguments before any CUDA call."""
    from reid_gan_b200 import _lib
    L = _lib.lib()
    buf = (ctypes.c_int64 * 16)()
    a = (ctypes.addressof(buf) + 15) & ~15                   # some entries check 16-byte alignment first
    # position tables come together
    rc = L.reid_knn_rescore_mapped(a, 8, 64, 0, 8, a, a, a, 1, 16, 0, 4, 0.0, None, 0, a, None, a, a, a, a, a, None, None)
    assert rc == _lib.REID_ERR_INVALID_ARG and "row_pos" in _lib.last_error()
    # ... and cover all N rows
    rc = L.reid_knn_rescore_mapped(a, 8, 64, 0, 4, a, a, a, 1, 16, 0, 4, 0.0, None, 0, a, a, a, a, a, a, a, None, None)
    assert rc == _lib.REID_ERR_INVALID_ARG and "all N rows" in _lib.last_error()
    # column thresholds need the column lists
    rc = L.reid_knn_candidates_tc_abt(a, 512, a, 256, 64, 4, 0, 512, -16, 1, 2, a, a, a, 1, a, None, None, 0, None)
    assert rc == _lib.REID_ERR_INVALID_ARG and "column lists" in _lib.last_error()
    # the emission needs a counter array and a capacity
    rc = L.reid_knn_sample_tau_emit(a, a, a, 2, 8, 16, a, a, a, None, 0, 0, None)
    assert rc == _lib.REID_ERR_INVALID_ARG and "main lists" in _lib.last_error()
    rc = L.reid_features_to_half_gather(a, None, 8, 64, 4, a, a, None)
    assert rc == _lib.REID_ERR_INVALID_ARG
    # the prepass mode may cut its column tiles up to 8 ways, the top-k mode 4
    rc = L.reid_knn_candidates_tc_abt(a, 4096, a, 4096, 64, 4, 0, 256, 16, 8, 2, a, a, a, 0, None, None, None, 0, None)
    assert rc == _lib.REID_ERR_INVALID_ARG and "n_splits" in _lib.last_error()


def test_python_surface_matches_reference_signatures():
    import reid_gan_b200 as rg
    sig = inspect.signature(rg.compute_jaccard_distance)
    names = list(sig.parameters)
    assert names[:6] == ["target_features", "k1", "k2", "print_flag", "search_option", "use_float16"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["k1"], d["k2"], d["print_flag"], d["search_option"], d["use_float16"]) == (20, 6, True, 0, False)
    sig = inspect.signature(rg.ClusterMemory.__init__)
    assert list(sig.parameters)[1:] == ["num_features", "num_samples", "temp", "momentum", "use_hard", "use_conf"]
    assert list(inspect.signature(rg.ClusterMemory.forward).parameters)[1:] == ["inputs", "targets", "gan_inputs",
                                                                                 "conf_weight"]
    assert list(inspect.signature(rg.cm).parameters) == ["inputs", "indexes", "features", "momentum"]
    assert list(inspect.signature(rg.cm_hard).parameters) == ["inputs", "indexes", "features", "momentum"]
    c = rg.DBSCAN(eps=0.6, min_samples=4, metric="precomputed", n_jobs=-1)
    assert c.eps == 0.6 and c.min_samples == 4
    with pytest.raises(ValueError):
        rg.DBSCAN(eps=0.6, metric="euclidean")
    m = rg.ClusterMemory(32, 5)
    assert m.features.shape == (5, 32) and m.gan_features.shape == (5, 32) and m.temp == 0.05 and m.momentum == 0.2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import reid_gan_b200 as rg
    x, _ = rg.synth(64, 32, 4)
    with pytest.raises(RuntimeError):
        rg.compute_jaccard_distance(x, k1=5, k2=2, print_flag=False)
    with pytest.raises(RuntimeError):
        rg.DBSCAN(eps=0.5, min_samples=2).fit_predict(np.zeros((4, 4), np.float32))
    with pytest.raises(RuntimeError):
        rg.ClusterMemory(32, 4)(x[:8], torch.zeros(8, dtype=torch.long))
    with pytest.raises(RuntimeError):
        rg.generate_cluster_features(np.zeros(64, np.int64), x)


def test_synth_is_deterministic_and_unit_norm():
    import reid_gan_b200 as rg
    a, ia = rg.synth(100, 64, 7, 0.8, 3)
    b, ib = rg.synth(100, 64, 7, 0.8, 3)
    assert torch.equal(a, b) and torch.equal(ia, ib)
    assert torch.allclose(a.norm(dim=1), torch.ones(100), atol=1e-6)
