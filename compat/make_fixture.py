"""Writes a small synthetic S3DIS-layout dataset (the wire format between partition/ and learning/):

    <root>/superpoint_graphs/Area_k/<room>.h5   sp_labels, sp_centroids, sp_length, sp_volume, sp_surface,
                                                 sp_point_count, source, target, se_delta_mean, se_delta_std
                                                 (partition/provider.py:558-600, read by learning/spg.py:66-103)
    <root>/parsed/Area_k/<room>.h5               one dataset per superpoint: [n_pts, 14] = xyz rgb e lpsv XYZ
                                                 (learning/s3dis_dataset.py:151-158, read by spg.py:198-205)

With the h5py stand-in (compat/h5py.py) the files are .npz archives under the .h5 names; with a real h5py
they are real HDF5 files — the writer below only uses `h5py.File(..., 'w').create_dataset`.
"""
import os
import sys

import numpy as np


def write_room(h5py, root, area, room, n_sp, rng, n_classes=13):
    from scipy.spatial import cKDTree
    cent = rng.uniform(0, 10, size=(n_sp, 3)).astype(np.float32)
    _, nbr = cKDTree(cent).query(cent, k=min(6, n_sp))
    src = np.repeat(np.arange(n_sp), nbr.shape[1] - 1)
    dst = nbr[:, 1:].reshape(-1)
    pairs = np.unique(np.concatenate([np.stack([src, dst], 1), np.stack([dst, src], 1)], 0), axis=0)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    E = pairs.shape[0]
    counts = np.clip(rng.lognormal(np.log(120.0), 1.0, size=n_sp), 5, 2000).astype(np.int64)
    labels = np.zeros((n_sp, n_classes + 1), dtype=np.uint32)
    cls = rng.integers(0, n_classes, size=n_sp)
    labels[np.arange(n_sp), 1 + cls] = counts
    unl = rng.random(n_sp) < 0.05
    labels[unl] = 0
    labels[unl, 0] = counts[unl]
    gdir = os.path.join(root, "superpoint_graphs", "Area_%d" % area)
    pdir = os.path.join(root, "parsed", "Area_%d" % area)
    os.makedirs(gdir, exist_ok=True)
    os.makedirs(pdir, exist_ok=True)
    with h5py.File(os.path.join(gdir, room + ".h5"), "w") as f:
        f.create_dataset("sp_labels", data=labels)
        f.create_dataset("sp_centroids", data=cent)
        f.create_dataset("sp_length", data=rng.uniform(0.1, 2, size=(n_sp, 1)).astype(np.float32))
        f.create_dataset("sp_surface", data=rng.uniform(0.1, 2, size=(n_sp, 1)).astype(np.float32))
        f.create_dataset("sp_volume", data=rng.uniform(0.1, 2, size=(n_sp, 1)).astype(np.float32))
        f.create_dataset("sp_point_count", data=counts[:, None].astype(np.uint64))
        f.create_dataset("source", data=pairs[:, :1].astype(np.uint32))
        f.create_dataset("target", data=pairs[:, 1:].astype(np.uint32))
        f.create_dataset("se_delta_mean", data=(cent[pairs[:, 0]] - cent[pairs[:, 1]]).astype(np.float32))
        f.create_dataset("se_delta_std", data=rng.uniform(0, 1, size=(E, 3)).astype(np.float32))
    with h5py.File(os.path.join(pdir, room + ".h5"), "w") as f:
        for i in range(n_sp):
            n = int(counts[i])
            P = np.empty((n, 14), dtype=np.float32)
            P[:, :3] = cent[i] + rng.standard_normal((n, 3)) * 0.3
            P[:, 3:11] = rng.uniform(-0.5, 0.5, size=(n, 8))
            P[:, 11:14] = rng.uniform(0, 1, size=(n, 3))
            f.create_dataset("%d" % i, data=P)
    return n_sp, E


def make(root, rooms_per_area=2, n_sp=120, seed=0):
    import h5py
    rng = np.random.default_rng(seed)
    total = [0, 0]
    for area in range(1, 7):
        for r in range(rooms_per_area):
            n, e = write_room(h5py, root, area, "office_%d" % (r + 1), n_sp + 7 * r + area, rng)
            total[0] += n
            total[1] += e
    return tuple(total)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    print(make(sys.argv[1]))
