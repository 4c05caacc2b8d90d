"""Developer timing of the exact re-score variants (needs the -DREID_DEV library: REID_DEV=1 bash reid-gan_b200/csrc/build.sh).
Usage: python scripts/dev_rescore_variants.py [N]"""
import os
import sys
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libreid_b200.so", "libreid_b200_dev.so")
from reid_gan_b200 import faiss_rerank as fr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32621
x, _ = rg.synth(N, 2048, max(1, N // 31), 0.8, 0)
x = x.cuda()
ref = None
for variant, waves in ((1, 2), (2, 1000), (2, 8), (2, 4), (2, 2), (2, 1), (3, 2), (3, 1), (2, 1000), (2, 2)):
    os.environ["REID_RESCORE_VARIANT"] = str(variant)
    os.environ["REID_MMA_WAVES"] = str(waves)
    for _ in range(2):
        idx, key, info = fr.knn_search(x, 30, "tc")
    torch.cuda.synchronize()
    _lib.profiler.start()
    for _ in range(5):
        idx, key, info = fr.knn_search(x, 30, "tc")
    torch.cuda.synchronize()
    _lib.profiler.stop()
    ms = _lib.profiler.summary()["reid_knn_rescore"]
    if ref is None:
        ref = (idx.clone(), key.clone())
    same = torch.equal(idx, ref[0]) and torch.equal(key, ref[1])
    print("variant %d waves %d: reid_knn_rescore %.4f ms  identical_to_variant_1=%s uncertified=%d" % (variant, waves, ms[1] / ms[0], same, info["uncertified_rows"]), flush=True)
