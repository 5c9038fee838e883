"""CPU restatement of the steps either side of the path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and the golden generator,
never by the product.

* `load_superpoint`, `augment_cloud`, `stack_clouds` follow learning/spg.py:198-236, :238-260 and
  :146-166 of the reference.  Pinned by tests/golden/loader_*.npz, produced by the reference's own
  `load_superpoint` (test-mode RNG, h5py replaced by an in-memory stand-in) — see
  tests/golden/make_golden.py.  The reference's `augment_cloud` needs `transforms3d`, which is not
  in this image: its three matrices are restated from the library's documented definitions
  (`zfdir2mat(s)` = s*I, `zfdir2mat(-1, axis)` = reflection across the plane normal to `axis`,
  `axangle2mat([0,0,1], a)` = rotation about z), i.e. augmentation parity is unpinned.
* `ConfusionMatrix` follows learning/metrics.py:8-79 (pinned by golden: the reference class runs
  here, it only needs numpy).
"""
import math
import random

import numpy as np


ATTRIB_COLUMNS = (("xyz", (0, 1, 2)), ("rgb", (3, 4, 5)), ("e", (6,)), ("lpsv", (7, 8, 9, 10)),
                  ("XYZ", (11, 12, 13)), ("d", (14,)))


def attrib_columns(pc_attribs):
    """Source columns selected by `--pc_attribs` (ref: learning/spg.py:221-229; substring tests in
    this fixed order).  Empty string = every column of the parsed array."""
    if pc_attribs == "":
        return None
    if "d" in pc_attribs:
        # the reference appends the 1-D slice P[:,14] and np.concatenate then raises (spg.py:228,230)
        raise ValueError("all the input arrays must have same number of dimensions")
    cols = []
    for key, cc in ATTRIB_COLUMNS:
        if key in pc_attribs:
            cols.extend(cc)
    return cols


def sample_indices(n, npts, rs):
    """Row of every output point (ref: learning/spg.py:209-214): more than npts points ->
    `rs.choice(n, npts)`; fewer -> the points themselves followed by `rs.choice(n, npts-n)`."""
    if n > npts:
        return rs.choice(n, npts)
    if n < npts:
        return np.concatenate([np.arange(n), rs.choice(n, npts - n)])
    return np.arange(n)


def test_rng(sp_id, test_seed_offset):
    """Evaluation draws are reproducible per superpoint (ref: learning/spg.py:207)."""
    return np.random.RandomState(seed=sp_id + test_seed_offset)


def augment_matrix(pc_augm_scale, pc_augm_rot, pc_augm_mirror_prob, rnd=random):
    """The 3x3 of augment_cloud (ref: learning/spg.py:240-253), drawing from `rnd` in the same order."""
    M = np.eye(3)
    if pc_augm_scale > 1:
        s = rnd.uniform(1 / pc_augm_scale, pc_augm_scale)
        M = np.dot(np.eye(3) * s, M)
    if pc_augm_rot == 1:
        angle = rnd.uniform(0, 2 * math.pi)
        c, s_ = math.cos(angle), math.sin(angle)
        R = np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]])
        M = np.dot(R, M)
    if pc_augm_mirror_prob > 0:
        if rnd.random() < pc_augm_mirror_prob / 2:
            M = np.dot(np.diag([-1.0, 1.0, 1.0]), M)
        if rnd.random() < pc_augm_mirror_prob / 2:
            M = np.dot(np.diag([1.0, -1.0, 1.0]), M)
    return M


def jitter_noise(shape, rs=np.random):
    """ref: learning/spg.py:256-258 (sigma 0.01, clip 0.05)."""
    sigma, clip = 0.01, 0.05
    return np.clip(sigma * rs.randn(*shape), -1 * clip, clip).astype(np.float32)


def load_superpoint(P, ii, pc_attribs="xyzrgbelpsvXYZ", pc_xyznormalize=1, M=None, noise=None):
    """P [n, C] float32 parsed points of one superpoint, ii the sampled rows (sample_indices).
    Returns (cloud [npts, F] float32, diameter float32[1]) as the reference's load_superpoint does
    for a superpoint with at least ptn_minpts points."""
    P = P.astype(np.float32)[ii, ...]
    if pc_xyznormalize:
        diameter = np.max(np.max(P[:, :3], axis=0) - np.min(P[:, :3], axis=0))
        P[:, :3] = (P[:, :3] - np.mean(P[:, :3], axis=0, keepdims=True)) / (diameter + 1e-10)
    else:
        diameter = 0.0
        P[:, :3] = (P[:, :3] - np.mean(P[:, :3], axis=0, keepdims=True))
    cols = attrib_columns(pc_attribs)
    if cols is not None:
        P = P[:, cols]
    if M is not None:
        P[:, :3] = np.dot(P[:, :3], M.T)
    if noise is not None:
        P = P + noise
    return P, np.array([diameter], dtype=np.float32)


def stack_clouds(clouds):
    """[npts, F] per superpoint -> [Nv, F, npts] (ref: learning/spg.py:152,161 `cloud.T`, np.stack)."""
    return np.stack([c.T for c in clouds])


class ConfusionMatrix(object):
    """ref: learning/metrics.py:8-79, vectorised where the reference loops."""

    def __init__(self, number_of_labels=2):
        self.number_of_labels = number_of_labels
        self.confusion_matrix = np.zeros((number_of_labels, number_of_labels))

    def count_predicted_batch(self, ground_truth_vec, predicted):
        for i in range(ground_truth_vec.shape[0]):
            self.confusion_matrix[:, predicted[i]] += ground_truth_vec[i, :]

    def get_intersection_union_per_class(self):
        cm = self.confusion_matrix
        diag = np.diag(cm)
        div = cm.sum(1) + cm.sum(0) - diag
        div = np.where(diag == 0, 1, div)
        return [float(d) / v for d, v in zip(diag, div)]

    def get_overall_accuracy(self):
        tot = self.confusion_matrix.sum()
        return float(np.trace(self.confusion_matrix)) / (tot if tot != 0 else 1)

    def get_average_intersection_union(self):
        values = self.get_intersection_union_per_class()
        class_seen = ((self.confusion_matrix.sum(1) + self.confusion_matrix.sum(0)) != 0).sum()
        return sum(values) / class_seen

    def get_mean_class_accuracy(self):
        re = 0
        for i in range(self.number_of_labels):
            re = re + self.confusion_matrix[i][i] / max(1, np.sum(self.confusion_matrix[i, :]))
        return re / self.number_of_labels


def eval_bookkeeping(outputs, label_mode, label_vec, n_classes):
    """One evaluation batch (ref: learning/main.py:257-262): returns (predictions of every node,
    confusion matrix [C,C], n_valid, n_correct)."""
    pred = np.argmax(outputs, 1)
    idx = label_mode != -100
    cm = ConfusionMatrix(n_classes)
    if idx.sum() > 0:
        cm.count_predicted_batch(label_vec[idx, ...], pred[idx])
    return pred, cm.confusion_matrix, int(idx.sum()), int((pred[idx] == label_mode[idx]).sum())


# ------------------------------------------------------------------------ label up-sampling
def reduced_labels2full(labels_red, components, n_ver):
    """partition/provider.py:630-635."""
    labels_full = np.zeros((n_ver,), dtype='uint8')
    for i_com in range(0, len(components)):
        labels_full[components[i_com]] = labels_red[i_com]
    return labels_full


def interpolate_labels(xyz_up, xyz, labels):
    """partition/provider.py:676-682: 1-NN label transfer.  The reference calls scikit-learn
    (NearestNeighbors(n_neighbors=1, algorithm='kd_tree'), unpinned in the reference; 1.x here), whose
    published algorithm is an exact nearest-neighbour search on float64 copies of the coordinates; it is
    restated as a blocked brute-force argmin of the float64 squared distance (same result whenever the
    nearest neighbour is unique)."""
    if len(labels.shape) > 1 and labels.shape[1] > 1:
        labels = np.argmax(labels, axis=1)
    ref = np.asarray(xyz, dtype=np.float64)
    out = np.empty(xyz_up.shape[0], dtype=np.int64)
    for i in range(0, xyz_up.shape[0], 2048):
        q = np.asarray(xyz_up[i:i + 2048], dtype=np.float64)
        d = ((q[:, None, :] - ref[None, :, :]) ** 2).sum(-1)
        out[i:i + 2048] = d.argmin(1)
    return labels[out].flatten(), out
