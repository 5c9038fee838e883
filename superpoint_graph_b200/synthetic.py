"""Synthetic S3DIS-shaped superpoint-graph batches; host-side numpy only.

Shapes and statistics follow what the reference's loader produces (learning/spg.py:130-236,
learning/ecc/GraphConvInfo.py:48-58): a symmetric kNN superpoint graph with ~10 in-edges per node
sorted by target, 13 standardised edge features, superpoints resampled to exactly `npts` points
with replacement (spg.py:209-214), xyz centred and divided by the bounding-box diameter,
`clouds_flag = -1` for superpoints with fewer than `minpts` raw points, labels with 5 % ignored.
"""
import numpy as np
import torch


def _knn_edges(rng, n, k):
    """Directed symmetric kNN edge list [E,2] (source, target) on random 3-D centroids."""
    pts = rng.uniform(0.0, 10.0, size=(n, 3))
    try:
        from scipy.spatial import cKDTree
        _, nbr = cKDTree(pts).query(pts, k=min(k + 1, n))
        nbr = nbr[:, 1:]
    except Exception:  # tiny fallback, O(n^2)
        d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d, np.inf)
        nbr = np.argsort(d, axis=1)[:, :k]
    src = np.repeat(np.arange(n), nbr.shape[1])
    dst = nbr.reshape(-1)
    pairs = np.concatenate([np.stack([src, dst], 1), np.stack([dst, src], 1)], 0)
    pairs = np.unique(pairs, axis=0)
    return pairs[pairs[:, 0] != pairs[:, 1]]


def make_batch(n_nodes=2048, k=8, nfeat=14, npts=128, n_edge_feats=13, n_classes=13, minpts=40,
               seed=1, isolated_frac=0.01):
    """Returns a dict of CPU torch tensors:
    clouds [Nv,F,L] f32, clouds_global [Nv] f32, clouds_flag [N] int64 (0 | -1),
    idxn [E] int64, degs [N] int64, edgefeats [E,Fe] f32, labels [N] int64."""
    rng = np.random.default_rng(seed)
    E = _knn_edges(rng, n_nodes, k)
    # a few nodes without in-edges (zero-degree rows must come out as zeros)
    iso = rng.choice(n_nodes, size=max(1, int(isolated_frac * n_nodes)), replace=False)
    E = E[~np.isin(E[:, 1], iso)]
    order = E[:, 1].argsort()  # sort by target, numpy default kind (GraphConvInfo.py:50)
    idxn = E[order, 0].astype(np.int64)
    degs = np.bincount(E[:, 1], minlength=n_nodes).astype(np.int64)
    edgefeats = rng.standard_normal((E.shape[0], n_edge_feats)).astype(np.float32)

    raw_counts = np.clip(rng.lognormal(np.log(200.0), 1.2, size=n_nodes), 1, 10000).astype(np.int64)
    flag = np.where(raw_counts < minpts, -1, 0).astype(np.int64)
    nv = int((flag == 0).sum())
    clouds = np.empty((nv, nfeat, npts), dtype=np.float32)
    xyz = rng.standard_normal((nv, npts, 3)).astype(np.float32)
    xyz -= xyz.mean(1, keepdims=True)
    diam = (xyz.max(1) - xyz.min(1)).max(1)
    xyz /= (diam[:, None, None] + 1e-10)
    clouds[:, :3, :] = xyz.transpose(0, 2, 1)
    if nfeat > 3:
        rest = rng.uniform(-0.5, 0.5, size=(nv, nfeat - 3, npts)).astype(np.float32)
        if nfeat >= 14:
            rest[:, -3:, :] = rng.uniform(0.0, 1.0, size=(nv, 3, npts))  # XYZ room-relative
        clouds[:, 3:, :] = rest
    clouds_global = rng.uniform(0.1, 3.0, size=nv).astype(np.float32)
    labels = rng.integers(0, n_classes, size=n_nodes).astype(np.int64)
    labels[rng.random(n_nodes) < 0.05] = -100
    return {
        "clouds": torch.from_numpy(clouds),
        "clouds_global": torch.from_numpy(clouds_global),
        "clouds_flag": torch.from_numpy(flag),
        "idxn": torch.from_numpy(idxn),
        "degs": torch.from_numpy(degs),
        "edgefeats": torch.from_numpy(edgefeats),
        "labels": torch.from_numpy(labels),
    }


def batch_counts(batch):
    """(superpoints N, embedded superpoints Nv, points Nv*L, directed edges E)."""
    nv, _, L = batch["clouds"].shape
    return int(batch["degs"].numel()), int(nv), int(nv * L), int(batch["idxn"].numel())
