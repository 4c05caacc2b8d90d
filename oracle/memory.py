"""ORACLE (test infrastructure, never on the product path).

CPU restatement (numpy, fp32 with fp64 only where noted) of the cluster-memory loss:
clustercontrast/models/cm.py:9-76 (CM, CM_Hard, cm, cm_hard) and :110-137
(ClusterMemory.forward), plus the autograd chain the reference leaves to torch
(outputs /= temp, F.cross_entropy(reduction='none'), F.normalize).

Pinned in tests/test_oracle.py against the reference cm.py run on CPU through
oracle/ref_shim.load_cm() and in tests/golden/cm_*.npz.
"""
import numpy as np


def normalize_rows(x, eps=1e-12):
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    return x / np.maximum(n, np.float32(eps))[:, None], n


def cm_forward(inputs, targets, features, temp):
    """cm.py:125,16,134-135 -> (loss (B,), logits z (B,C) after /temp, xhat, norms)."""
    xhat, nrm = normalize_rows(inputs)
    z = (xhat.astype(np.float64) @ np.asarray(features, np.float64).T).astype(np.float32)
    z = (z / np.float32(temp)).astype(np.float32)
    m = z.max(axis=1, keepdims=True)
    lse = (m[:, 0] + np.log(np.exp((z - m).astype(np.float64)).sum(axis=1))).astype(np.float32)
    loss = lse - z[np.arange(z.shape[0]), targets]
    return loss.astype(np.float32), z, xhat, nrm


def cm_backward(grad_loss, z, targets, features, xhat, nrm, temp):
    """d loss / d inputs through CE, /temp, mm (cm.py:26,56: PRE-update centroids) and normalize."""
    B, C = z.shape
    m = z.max(axis=1, keepdims=True)
    p = np.exp((z - m).astype(np.float64))
    p /= p.sum(axis=1, keepdims=True)
    p[np.arange(B), targets] -= 1.0
    gz = p * np.asarray(grad_loss, np.float64)[:, None] / float(temp)
    gxh = gz @ np.asarray(features, np.float64)
    xh = xhat.astype(np.float64)
    gx = (gxh - xh * (xh * gxh).sum(axis=1, keepdims=True)) / np.maximum(nrm.astype(np.float64), 1e-12)[:, None]
    return gx.astype(np.float32)


def cm_update(features, xhat, targets, momentum):
    """cm.py:29-31: sequential per-sample momentum update in batch order, in place."""
    f = np.array(features, dtype=np.float32, copy=True)
    mom = np.float32(momentum)
    for x, y in zip(xhat, targets):
        v = mom * f[y] + (np.float32(1.0) - mom) * x
        f[y] = v / np.sqrt(np.sum(v * v, dtype=np.float32))
    return f


def cm_hard_update(features, xhat, targets, momentum):
    """cm.py:58-70: per distinct label, hardest positive = first argmin of x.f[label]
    (pre-update centroid), one momentum update with it."""
    f = np.array(features, dtype=np.float32, copy=True)
    mom = np.float32(momentum)
    groups = {}
    for b, y in enumerate(np.asarray(targets).tolist()):
        groups.setdefault(y, []).append(b)
    chosen = {}
    for y, members in groups.items():
        d = np.array([np.dot(xhat[b].astype(np.float64), f[y].astype(np.float64)) for b in members]).astype(np.float32)
        b = members[int(np.argmin(d))]
        chosen[y] = b
        v = f[y] * mom + (np.float32(1.0) - mom) * xhat[b]
        f[y] = v / np.sqrt(np.sum(v * v, dtype=np.float32))
    return f, chosen
