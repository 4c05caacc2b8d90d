"""GPU (-m gpu): ClusterMemory / cm / cm_hard against the golden vectors of the reference's cm.py
(loss and centroids within 1e-4 relative, BASELINE.json north_star)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CM = sorted(glob.glob(os.path.join(GOLD, "cm_*.npz")))
RTOL, ATOL = 1e-4, 2e-6


@pytest.mark.parametrize("hard", [False, True])
@pytest.mark.parametrize("path", CM, ids=[os.path.basename(p)[:-4] for p in CM])
def test_cluster_memory_module(path, hard):
    import reid_gan_b200 as rg
    g = np.load(path)
    tag = "hard" if hard else "cm"
    C, D = g["features"].shape
    mem = rg.ClusterMemory(D, C, temp=float(g["temp"]), momentum=float(g["momentum"]), use_hard=hard).cuda()
    mem.features = torch.from_numpy(g["features"]).cuda()
    x = torch.from_numpy(g["inputs"]).cuda().requires_grad_(True)
    t = torch.from_numpy(g["targets"]).cuda()
    loss = mem(x, t)
    assert loss.shape == (x.shape[0],)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g["loss_" + tag], rtol=RTOL, atol=ATOL)
    before = mem.features.clone()
    loss.backward(torch.from_numpy(g["grad_loss"]).cuda())
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["grad_" + tag], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(mem.features.cpu().numpy(), g["features_after_" + tag], rtol=RTOL, atol=ATOL)
    touched = (mem.features != before).any(dim=1).cpu().numpy()
    assert set(np.nonzero(touched)[0]) <= set(g["targets"].tolist())
    np.testing.assert_allclose(mem.features.norm(dim=1).cpu().numpy()[touched], 1.0, atol=1e-6)


@pytest.mark.parametrize("hard", [False, True])
@pytest.mark.parametrize("path", CM, ids=[os.path.basename(p)[:-4] for p in CM])
def test_module_level_functions(path, hard):
    """cm()/cm_hard() composed with torch ops exactly as ClusterMemory.forward does (cm.py:125-135)."""
    import torch.nn.functional as F
    import reid_gan_b200 as rg
    g = np.load(path)
    tag = "hard" if hard else "cm"
    f = torch.from_numpy(g["features"]).cuda()
    x = torch.from_numpy(g["inputs"]).cuda().requires_grad_(True)
    t = torch.from_numpy(g["targets"]).cuda()
    out = (rg.cm_hard if hard else rg.cm)(F.normalize(x, dim=1), t, f, float(g["momentum"]))
    out = out / float(g["temp"])
    loss = F.cross_entropy(out, t, reduction="none")
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g["loss_" + tag], rtol=RTOL, atol=ATOL)
    loss.backward(torch.from_numpy(g["grad_loss"]).cuda())
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["grad_" + tag], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(f.cpu().numpy(), g["features_after_" + tag], rtol=RTOL, atol=ATOL)


def test_full_size_against_oracle():
    """BASELINE configs[2]: B=256 (16 x 16), ~700 clusters x 2048-d, temp 0.05, momentum 0.2."""
    import reid_gan_b200 as rg
    from oracle import memory as omem
    x, ids = rg.synth(22000, 2048, 700, 0.8, 0)
    cen = torch.nn.functional.normalize(torch.stack([x[ids == k].mean(0) for k in range(700)]), dim=1)
    inp, tgt = rg.synth_cm_batch(x, ids, ids, 16, 16, seed=1)
    for hard in (False, True):
        mem = rg.ClusterMemory(2048, 700, temp=0.05, momentum=0.2, use_hard=hard).cuda()
        mem.features = cen.clone().cuda()
        xi = inp.clone().cuda().requires_grad_(True)
        loss = mem(xi, tgt.cuda())
        loss.mean().backward()
        l_ref, z, xhat, nrm = omem.cm_forward(inp.numpy(), tgt.numpy(), cen.numpy(), 0.05)
        np.testing.assert_allclose(loss.detach().cpu().numpy(), l_ref, rtol=RTOL, atol=ATOL)
        g_ref = omem.cm_backward(np.full(256, 1 / 256, np.float32), z, tgt.numpy(), cen.numpy(), xhat, nrm, 0.05)
        np.testing.assert_allclose(xi.grad.cpu().numpy(), g_ref, rtol=1e-3, atol=1e-7)
        f_ref = omem.cm_hard_update(cen.numpy(), xhat, tgt.numpy(), 0.2)[0] if hard else \
            omem.cm_update(cen.numpy(), xhat, tgt.numpy(), 0.2)
        np.testing.assert_allclose(mem.features.cpu().numpy(), f_ref, rtol=RTOL, atol=ATOL)
