"""Synthetic ResNet-50-shaped re-ID features.

The reference extracts (N, 2048) L2-normalised fp32 features with its backbone
(clustercontrast/models/resnet.py:90-94, evaluators.py:30-68); there are no
datasets or checkpoints in this environment, so every test and benchmark uses
the seeded generator below: `n_ids` unit-norm identity centres plus isotropic
noise, re-normalised.  It is host-side only (CPU torch generator) so the same
seed gives the same bytes on every box.
"""
import torch
import torch.nn.functional as F


def synth(N, D=2048, n_ids=None, noise=0.8, seed=0):
    """Return (x, ids): x (N, D) fp32 unit-norm CPU tensor, ids (N,) int64."""
    if n_ids is None:
        n_ids = max(1, N // 31)
    g = torch.Generator().manual_seed(int(seed))
    centres = F.normalize(torch.randn(n_ids, D, generator=g), dim=1)
    ids = torch.randint(0, n_ids, (N,), generator=g)
    x = centres[ids] + noise * torch.randn(N, D, generator=g) / (D ** 0.5)
    x = F.normalize(x, dim=1).contiguous()
    return x, ids


def synth_cm_batch(x, ids, centroid_labels, num_ids=16, num_instances=16, seed=0):
    """A ClusterMemory batch shaped like the reference's identity sampler
    (utils/data/sampler.py:69-107): `num_ids` labels x `num_instances` members,
    features un-normalised by a random positive scale.
    Returns (inputs (B, D) fp32, targets (B,) int64)."""
    g = torch.Generator().manual_seed(int(seed) + 7919)
    labels = torch.unique(centroid_labels[centroid_labels >= 0])
    pick = labels[torch.randperm(labels.numel(), generator=g)[:num_ids]]
    rows, tg = [], []
    for lab in pick.tolist():
        members = torch.nonzero(centroid_labels == lab).flatten()
        sel = members[torch.randint(0, members.numel(), (num_instances,), generator=g)]
        rows.append(sel)
        tg.append(torch.full((num_instances,), lab, dtype=torch.int64))
    rows = torch.cat(rows)
    tg = torch.cat(tg)
    perm = torch.randperm(rows.numel(), generator=g)
    rows, tg = rows[perm], tg[perm]
    scale = 1.0 + 0.1 * torch.randn(rows.numel(), 1, generator=g)
    inputs = (x[rows] * scale.abs().clamp_min(0.1)).contiguous()
    return inputs, tg


def synth_device(N, D=2048, n_ids=None, noise=0.8, seed=0, device="cuda"):
    """Same construction as `synth`, generated on the device (seeded CUDA generator) -- for the scale sweep
    (N = 100k / 250k), where the host generator would take longer than the passes being measured.  The bytes
    differ from `synth`'s (different RNG), but are identical on every rank of one box."""
    if n_ids is None:
        n_ids = max(1, N // 31)
    g = torch.Generator(device=device).manual_seed(int(seed))
    centres = F.normalize(torch.randn(n_ids, D, generator=g, device=device), dim=1)
    ids = torch.randint(0, n_ids, (N,), generator=g, device=device)
    x = torch.empty((N, D), dtype=torch.float32, device=device)
    step = 1 << 16
    for a in range(0, N, step):                      # chunked: bounds the temporaries at 250k x 2048
        b = min(N, a + step)
        x[a:b] = F.normalize(centres[ids[a:b]] + noise * torch.randn(b - a, D, generator=g, device=device) / (D ** 0.5), dim=1)
    return x, ids


def synth_hard(N, D=2048, seed=0, iso_frac=0.04, hub_frac=0.005, dup_frac=0.003):
    """A harder set than `synth` for the parity gates: heavy-tailed identity sizes (2 .. ~600 rows), a noise level
    per identity in [0.8, 1.25], `iso_frac` isolated rows (random directions: DBSCAN noise points), `hub_frac` hub
    rows (normalised sums of 3-6 identity centres plus a little noise: they sit in many neighbour lists and give
    long inverted-index columns) and `dup_frac` exact duplicates of other rows (ties broken by index; self not at
    rank 0).  Rows are shuffled, so identities are not contiguous.  Returns (x (N, D) fp32 unit-norm, ids (N,)):
    ids >= 0 identity, -1 isolated, -2 hub; a duplicate carries the id of its source."""
    g = torch.Generator().manual_seed(int(seed) + 104729)
    n_iso, n_hub, n_dup = int(N * iso_frac), int(N * hub_frac), int(N * dup_frac)
    n_body = N - n_iso - n_hub - n_dup
    sizes = []
    while sum(sizes) < n_body:
        u = float(torch.rand(1, generator=g))
        sizes.append(int(min(600, 2 + 5.0 * (max(u, 1e-4) ** -0.75))))
    sizes[-1] -= sum(sizes) - n_body
    if sizes[-1] <= 0:
        sizes.pop()
        sizes[-1] += n_body - sum(sizes)
    n_ids = len(sizes)
    centres = F.normalize(torch.randn(n_ids, D, generator=g), dim=1)
    level = 0.8 + 0.45 * torch.rand(n_ids, generator=g)
    ids_body = torch.repeat_interleave(torch.arange(n_ids), torch.tensor(sizes))
    body = centres[ids_body] + level[ids_body, None] * torch.randn(n_body, D, generator=g) / (D ** 0.5)
    iso = torch.randn(n_iso, D, generator=g)
    hubs = torch.zeros(n_hub, D)
    for h in range(n_hub):
        m = int(torch.randint(3, 7, (1,), generator=g))
        pick = torch.randint(0, n_ids, (m,), generator=g)
        hubs[h] = centres[pick].sum(0) / m ** 0.5 + 0.3 * torch.randn(D, generator=g) / (D ** 0.5)
    x = F.normalize(torch.cat([body, iso, hubs]), dim=1)
    ids = torch.cat([ids_body, torch.full((n_iso,), -1), torch.full((n_hub,), -2)])
    src = torch.randint(0, x.shape[0], (n_dup,), generator=g)
    x = torch.cat([x, x[src]])
    ids = torch.cat([ids, ids[src]])
    perm = torch.randperm(N, generator=g)
    return x[perm].contiguous(), ids[perm].contiguous()
