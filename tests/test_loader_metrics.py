"""Steps either side of the path: the per-superpoint batch
loader and the evaluation bookkeeping.  Golden files come from the reference's own
`load_superpoint` / `ConfusionMatrix` (tests/golden/make_golden.py loader)."""
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import loader_ref as lr  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("s3dis", "sema", "nonorm")


def _gold(name):
    return np.load(os.path.join(GOLD, name))


def _case(g, tag):
    norm, npts, minpts, off = [int(v) for v in g[tag + "_cfg"]]
    return SimpleNamespace(pc_xyznormalize=norm, ptn_npts=npts, ptn_minpts=minpts,
                           pc_attribs=str(g[tag + "_attribs"]), pc_augm_scale=0, pc_augm_rot=1,
                           pc_augm_mirror_prob=0, pc_augm_jitter=1), off


# ------------------------------------------------------------------ CPU: oracle and host logic
@pytest.mark.parametrize("tag", CASES)
def test_oracle_load_superpoint_matches_reference_bitwise(tag):
    g = _gold("loader_clouds.npz")
    args, off = _case(g, tag)
    flags, clouds, diams = [], [], []
    for sid, n in enumerate(g["counts"]):
        if n < args.ptn_minpts:
            flags.append(-1)
            continue
        flags.append(0)
        ii = lr.sample_indices(int(n), args.ptn_npts, lr.test_rng(sid, off))
        c, d = lr.load_superpoint(g["P%d" % sid], ii, args.pc_attribs, args.pc_xyznormalize)
        clouds.append(c)
        diams.append(d)
    assert np.array_equal(np.array(flags), g[tag + "_flag"])
    assert np.array_equal(lr.stack_clouds(clouds), g[tag + "_clouds"])
    assert np.array_equal(np.concatenate(diams), g[tag + "_global"])


def test_oracle_confusion_matrix_matches_reference():
    m = _gold("metrics_confusion.npz")
    cm = lr.ConfusionMatrix(13)
    for b in range(3):
        pred, c, _, _ = lr.eval_bookkeeping(m["o%d" % b], m["t%d" % b], m["tvec%d" % b], 13)
        assert np.array_equal(pred, m["pred%d" % b])
        cm.confusion_matrix += c
    assert np.array_equal(cm.confusion_matrix, m["cm"])
    assert cm.get_overall_accuracy() == float(m["oa"])
    assert np.allclose(cm.get_average_intersection_union(), float(m["miou"]), rtol=1e-15)
    assert np.allclose(cm.get_intersection_union_per_class(), m["iou"], rtol=1e-15)
    assert np.allclose(cm.get_mean_class_accuracy(), float(m["macc"]), rtol=1e-15)


def test_host_mirror_matches_oracle_draw_for_draw():
    from superpoint_graph_b200 import spg_loader, spg_metrics
    for n in (1, 40, 127, 128, 129, 1000):
        a = spg_loader.sample_indices(n, 128, np.random.RandomState(5))
        b = lr.sample_indices(n, 128, np.random.RandomState(5))
        assert np.array_equal(a, b) and a.shape == (128,) and a.max() < n
    for scale, rot, mirror in ((0, 1, 0), (1.5, 1, 0.9), (2.0, 0, 1.0), (0, 0, 0)):
        args = SimpleNamespace(pc_augm_scale=scale, pc_augm_rot=rot, pc_augm_mirror_prob=mirror)
        M1 = spg_loader.augment_matrix(args, random.Random(3))
        M2 = lr.augment_matrix(scale, rot, mirror, random.Random(3))
        assert np.array_equal(M1, M2)
        if scale == 0:
            assert np.allclose(M1 @ M1.T, np.eye(3), atol=1e-12)  # rotations / reflections only
    assert spg_loader.attrib_columns("xyzrgbelpsvXYZ", 15) == list(range(14))
    assert spg_loader.attrib_columns("xyzelpsv", 15) == lr.attrib_columns("xyzelpsv")
    assert spg_loader.attrib_columns("", 15) == list(range(15))
    with pytest.raises(ValueError):  # the reference cannot concatenate its 1-D 'd' column either
        spg_loader.attrib_columns("xyzd", 15)
    # host interface of the metrics mirror against the reference's matrix
    m = _gold("metrics_confusion.npz")
    cm = spg_metrics.ConfusionMatrix(13)
    for b in range(3):
        idx = m["t%d" % b] != -100
        cm.count_predicted_batch(m["tvec%d" % b][idx], m["pred%d" % b][idx])
    assert np.array_equal(cm.confusion_matrix, m["cm"])
    assert cm.get_overall_accuracy() == float(m["oa"])
    assert np.allclose(cm.get_intersection_union_per_class(), m["iou"], rtol=1e-15)
    assert np.allclose(cm.get_average_intersection_union(), float(m["miou"]), rtol=1e-15)
    assert np.allclose(cm.get_mean_class_accuracy(), float(m["macc"]), rtol=1e-15)
    hard = spg_metrics.ConfusionMatrix(4)
    hard.count_predicted_batch_hard(np.array([0, 1, 1, 3]), np.array([0, 2, 2, 3]))
    hard.count_predicted(2, 2, 5)
    assert hard.get_count(1, 2) == 2 and hard.get_count(2, 2) == 5 and hard.count_gt(1) == 2


# ------------------------------------------------------------------ GPU: kernels
def _store(g, dev):
    from superpoint_graph_b200.spg_loader import SuperpointStore
    st = SuperpointStore()
    st.add("mem", {sid: g["P%d" % sid] for sid in range(len(g["counts"]))})
    return st.finalize(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_cloud_build_matches_reference_bitwise(tag):
    from superpoint_graph_b200.spg_loader import load_superpoints
    dev = torch.device("cuda:0")
    g = _gold("loader_clouds.npz")
    args, off = _case(g, tag)
    flags, clouds, diam = load_superpoints(_store(g, dev), "mem", range(len(g["counts"])), args,
                                           train=False, test_seed_offset=off)
    assert np.array_equal(flags, g[tag + "_flag"])
    assert np.array_equal(clouds.cpu().numpy(), g[tag + "_clouds"])
    assert np.array_equal(diam.cpu().numpy(), g[tag + "_global"])


@pytest.mark.gpu
def test_cloud_build_training_augmentation_vs_oracle():
    """Training mode: same draws as the reference's loader (numpy global state for the samples and
    the jitter, `random` for the 3x3); the rotation is evaluated in double on both sides."""
    from superpoint_graph_b200.spg_loader import load_superpoints
    dev = torch.device("cuda:0")
    g = _gold("loader_clouds.npz")
    args, _ = _case(g, "s3dis")
    args.pc_augm_scale, args.pc_augm_mirror_prob = 1.3, 0.8
    ids = [sid for sid, n in enumerate(g["counts"]) if n >= args.ptn_minpts]
    np.random.seed(4)
    random.seed(4)
    _, clouds, diam = load_superpoints(_store(g, dev), "mem", ids, args, train=True)
    np.random.seed(4)
    random.seed(4)
    want, wdiam = [], []
    for sid in ids:
        n = int(g["counts"][sid])
        ii = lr.sample_indices(n, args.ptn_npts, np.random.random.__self__)
        M = lr.augment_matrix(args.pc_augm_scale, args.pc_augm_rot, args.pc_augm_mirror_prob)
        noise = lr.jitter_noise((args.ptn_npts, 14))
        c, d = lr.load_superpoint(g["P%d" % sid], ii, args.pc_attribs, 1, M=M, noise=noise)
        want.append(c)
        wdiam.append(d)
    want = lr.stack_clouds(want)
    got = clouds.cpu().numpy()
    assert np.array_equal(diam.cpu().numpy(), np.concatenate(wdiam))
    assert np.array_equal(got[:, 3:], want[:, 3:])            # untouched by the 3x3: exact
    np.testing.assert_allclose(got[:, :3], want[:, :3], rtol=0, atol=2e-7)  # BLAS vs FMA order in double


@pytest.mark.gpu
def test_cloud_build_device_rng_properties():
    """device_rng=True: nothing per point is drawn on the host.  The original points come first
    (spg.py:212-214), every sampled row is a row of the same superpoint, the jitter is clipped
    N(0, 0.01) noise."""
    from superpoint_graph_b200.spg_loader import load_superpoints
    dev = torch.device("cuda:0")
    g = _gold("loader_clouds.npz")
    args, _ = _case(g, "s3dis")
    args.pc_augm_rot = 0
    st = _store(g, dev)
    ids = [sid for sid, n in enumerate(g["counts"]) if n >= args.ptn_minpts]
    args.pc_augm_jitter = 0
    _, clean, _ = load_superpoints(st, "mem", ids, args, train=True, device_rng=True, seed=9)
    _, again, _ = load_superpoints(st, "mem", ids, args, train=True, device_rng=True, seed=9)
    _, other, _ = load_superpoints(st, "mem", ids, args, train=True, device_rng=True, seed=10)
    assert torch.equal(clean, again) and not torch.equal(clean, other)
    clean = clean.cpu().numpy()
    for k, sid in enumerate(ids):
        P = g["P%d" % sid]
        n = P.shape[0]
        rgb = clean[k, 3:6].T  # columns 3..5 are copied unchanged
        if n <= args.ptn_npts:
            assert np.array_equal(rgb[:n], P[:, 3:6])
        rows = {tuple(r) for r in P[:, 3:6].tolist()}
        assert all(tuple(r) in rows for r in rgb.tolist())
        if n > 4 * args.ptn_npts:
            assert len({tuple(r) for r in rgb.tolist()}) > args.ptn_npts // 2  # not a constant draw
    args.pc_augm_jitter = 1
    _, noisy, _ = load_superpoints(st, "mem", ids, args, train=True, device_rng=True, seed=9)
    d = (noisy.cpu().numpy() - clean).ravel()
    assert np.abs(d).max() <= 0.05 + 1e-6
    assert abs(d.std() - 0.01) < 1e-3 and abs(d.mean()) < 5e-4


@pytest.mark.gpu
def test_confusion_count_matches_reference_exactly():
    from superpoint_graph_b200.spg_metrics import ConfusionMatrix, MultiSampleMean
    dev = torch.device("cuda:0")
    m = _gold("metrics_confusion.npz")
    cm = ConfusionMatrix(13)
    n_valid = n_ok = 0
    for b in range(3):
        o, t, tv = m["o%d" % b], m["t%d" % b], m["tvec%d" % b]
        pred = cm.count_predicted_batch_device(torch.from_numpy(o).to(dev), torch.from_numpy(t).to(dev),
                                               torch.from_numpy(tv).to(dev), want_predictions=True)
        assert np.array_equal(pred.cpu().numpy(), m["pred%d" % b])  # ties -> first maximum
        n_valid += int((t != -100).sum())
        n_ok += int((m["pred%d" % b] == t).sum())
    assert np.array_equal(cm.confusion_matrix, m["cm"])
    assert cm.get_overall_accuracy() == float(m["oa"])
    assert np.allclose(cm.get_average_intersection_union(), float(m["miou"]), rtol=1e-15)
    assert cm.accuracy() == 100.0 * n_ok / n_valid
    # wide class counts (more classes than lanes) and an all-unlabelled batch
    rng = np.random.default_rng(3)
    C, n = 70, 257
    o = rng.standard_normal((n, C)).astype(np.float32)
    tv = rng.integers(0, 50, size=(n, C))
    t = rng.integers(0, C, size=n)
    t[::3] = -100
    big = ConfusionMatrix(C)
    big.count_predicted_batch_device(torch.from_numpy(o).to(dev), torch.from_numpy(t).to(dev),
                                     torch.from_numpy(tv).to(dev))
    _, want, _, _ = lr.eval_bookkeeping(o, t, tv, C)
    assert np.array_equal(big.confusion_matrix, want)
    none = ConfusionMatrix(C)
    none.count_predicted_batch_device(torch.from_numpy(o).to(dev),
                                      torch.full((n,), -100, dtype=torch.int64, device=dev),
                                      torch.from_numpy(tv).to(dev))
    assert none.confusion_matrix.sum() == 0 and none.accuracy() == 0
    # eval_final's multi-sample average
    ms = MultiSampleMean()
    for s in m["ms_samples"]:
        ms.add(torch.from_numpy(s).to(dev))
    assert np.array_equal(ms.value().cpu().numpy(), m["ms_mean"])


@pytest.mark.gpu
def test_cloud_build_unpadded_rows_take_the_scalar_path():
    """The C-ABI accepts any row pitch; 15-float rows (the parsed files as they are) cannot be
    fetched with 128-bit loads and go through the scalar path — same bits."""
    from superpoint_graph_b200 import ops
    dev = torch.device("cuda:0")
    g = _gold("loader_clouds.npz")
    args, off = _case(g, "s3dis")
    ids = [sid for sid, n in enumerate(g["counts"]) if n >= args.ptn_minpts]
    pts = np.concatenate([g["P%d" % sid] for sid in ids], 0)
    counts = np.array([g["counts"][sid] for sid in ids])
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    idx = np.stack([lr.sample_indices(int(g["counts"][sid]), 128, lr.test_rng(sid, off)) for sid in ids])
    cols = lr.attrib_columns(args.pc_attribs)
    clouds = torch.empty((len(ids), len(cols), 128), device=dev)
    diam = torch.empty(len(ids), device=dev)
    ops.cloud_build(torch.from_numpy(pts).to(dev), torch.from_numpy(starts.astype(np.int64)).to(dev),
                    torch.from_numpy(counts.astype(np.int32)).to(dev),
                    torch.from_numpy(idx.astype(np.int32)).to(dev),
                    torch.tensor(cols, dtype=torch.int32, device=dev), 128, True, None, None, 0.0, 0.05, 0,
                    clouds, diam)
    assert np.array_equal(clouds.cpu().numpy(), g["s3dis_clouds"])
    assert np.array_equal(diam.cpu().numpy(), g["s3dis_global"])


# ------------------------------------------------------------------ label up-sampling (8(f) rank 4)
def test_oracle_upsampling_matches_reference_golden(golden_dir):
    """oracle/loader_ref restatements against outputs of the reference's own function text
    (partition/provider.py:630-635,676-682, generated by tests/golden/make_golden.py)."""
    from oracle import loader_ref
    g = np.load(os.path.join(golden_dir, "upsampling.npz"))
    lab, _ = loader_ref.interpolate_labels(g["xyz_up"], g["xyz"], g["logits"])
    assert np.array_equal(lab, g["lab_up"])
    lab, _ = loader_ref.interpolate_labels(g["xyz_up"], g["xyz"], g["hard"])
    assert np.array_equal(lab, g["lab_up_hard"])
    comps = [np.nonzero(g["comp_of"] == c)[0] for c in range(40)]
    assert np.array_equal(loader_ref.reduced_labels2full(g["labels_red"], comps, 700), g["full"])


@pytest.mark.gpu
def test_gpu_upsampling_bit_exact(golden_dir):
    """spg_nn1_interpolate / spg_labels_to_points through the Python mirrors: labels identical to the
    reference's (golden), neighbour indices identical to the oracle's, batched queries included; a
    larger random case against the oracle."""
    from oracle import loader_ref
    from superpoint_graph_b200 import spg_metrics
    g = np.load(os.path.join(golden_dir, "upsampling.npz"))
    dev = torch.device("cuda:0")
    lab, idx = spg_metrics.interpolate_labels(torch.from_numpy(g["xyz_up"]).to(dev), g["xyz"], g["logits"], return_index=True)
    assert np.array_equal(lab.cpu().numpy(), g["lab_up"])
    _, ref_idx = loader_ref.interpolate_labels(g["xyz_up"], g["xyz"], g["hard"])
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref_idx)
    lab_b = spg_metrics.interpolate_labels(g["xyz_up"], g["xyz"], g["hard"], ver_batch=777)
    assert np.array_equal(lab_b.cpu().numpy(), g["lab_up_hard"])
    comps = [np.nonzero(g["comp_of"] == c)[0] for c in range(40)]
    full = spg_metrics.reduced_labels2full(g["labels_red"], comps, 700)
    assert full.dtype == torch.uint8 and np.array_equal(full.cpu().numpy(), g["full"])
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-50, 50, size=(3001, 3)).astype(np.float32)  # not a multiple of the 1024-point tile
    up = rng.uniform(-55, 55, size=(20000, 3)).astype(np.float32)
    labs = rng.integers(0, 8, size=3001)
    want, want_idx = loader_ref.interpolate_labels(up, xyz, labs)
    got, got_idx = spg_metrics.interpolate_labels(up, xyz, labs, return_index=True)
    assert np.array_equal(got_idx.cpu().numpy().astype(np.int64), want_idx)
    assert np.array_equal(got.cpu().numpy(), want)
