"""Developer A/B of the Jaccard eps-graph stage: partner-count guess on / off, three data shapes."""
import sys
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import _lib, pipeline, faiss_rerank as fr

for name, gen, kw in (("c2", "synth", dict(N=32621, D=2048, n_ids=1041, noise=0.8, seed=0)),
                      ("market", "synth", dict(N=12936, D=2048, n_ids=751, noise=0.8, seed=0)),
                      ("hard", "synth_hard", dict(N=20480, D=2048, seed=0))):
    x, _ = getattr(rg, gen)(**kw)
    x = x.cuda()
    ref = None
    for guess in (False, True):
        fr.PARTNER_GUESS = guess
        fr._guess_hint.clear()
        for _ in range(3):
            out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
        torch.cuda.synchronize()
        _lib.profiler.start()
        for _ in range(5):
            out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
        torch.cuda.synchronize()
        _lib.profiler.stop()
        p = _lib.profiler.summary()
        lab = out["labels"].cpu()
        if ref is None:
            ref = lab
        print("%-7s guess=%-5s (adaptive hint: %s) eps_graph %.3f ms  pass(sum of calls) %.3f ms  same_labels=%s" % (
            name, guess, fr._guess_hint.get(x.shape[0], True), p["reid_jaccard_eps_graph"][1] / 5, sum(v[1] for v in p.values()) / 5, torch.equal(lab, ref)), flush=True)
