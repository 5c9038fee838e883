"""The BASELINE.json configurations as named workloads (model flags, synthetic batch shape, mode).

One table shared by bench.py, the parity tests and the reference arm, so that "the benchmarked
configuration" and "the tested configuration" cannot drift apart.  Flags follow the reference's
documented command lines (S3DIS.md:27-30, Semantic3D.md:20-22, vKITTI3D.md:46-50) and
learning/main.py's defaults; the batches are synthetic (superpoint_graph_b200.synthetic).
"""
from .synthetic import make_batch
from .trainer import make_args

WORKLOADS = {
    # configs[1]: S3DIS Area-5 fold training, gru_10_1_1_1_0, fp32 (main.py:49,97-99: batch 2,
    # hardcutoff 512 -> 2 x 512 superpoints)
    "s3dis_train": dict(
        title="configs[1]: S3DIS-shaped training step, gru_10_1_1_1_0,f_13, fp32, 2 scenes x %(half)d superpoints per GPU",
        args=dict(), batch=dict(nfeat=14, n_classes=13, minpts=40), nodes=1024, train=True, dtype="f32"),
    # configs[0]: one whole room, eval mode, PointNet + 1 x ECC (main.py:229-264, batch_size 1)
    "room_fwd": dict(
        title="configs[0]: S3DIS-shaped single-room forward (PointNet + 1 x ECC, eval mode), %(nodes)d superpoints",
        args=dict(model_config="gru_1_1_1_1_0,f_13"), batch=dict(nfeat=14, n_classes=13, minpts=40),
        nodes=1536, train=False, dtype="f32"),
    # configs[2]: Semantic3D reduced-8 inference, gru_10,f_8, xyzrgbelpsv (F=11), whole-scene graphs
    "sema3d_eval": dict(
        title="configs[2]: Semantic3D-shaped inference, gru_10,f_8, F=11, eval mode, %(nodes)d superpoints",
        args=dict(model_config="gru_10,f_8", node_feats=11, ptn_nfeat_stn=11, classes=8),
        batch=dict(nfeat=11, n_classes=8, minpts=40), nodes=20000, train=False, dtype="f32"),
    # configs[3]: vKITTI3D SPG widths (vKITTI3D.md:46-50), batch 4 x hardcutoff 256, minpts 15, xyzXYZrgb (F=9).
    # Training runs in fp32; the inference forward exists in bf16 arithmetic (PointNet trunk: kind::f16).
    "vkitti_train": dict(
        title="configs[3] widths: vKITTI3D-shaped training step, gru_10_1_1_1_0,f_13, F=9, fp32, 4 scenes x %(quarter)d superpoints per GPU",
        args=dict(node_feats=9, ptn_nfeat_stn=9, ptn_widths=[[64, 64, 128], [64, 32, 32]],
                  ptn_widths_stn=[[32, 64], [32, 16]]),
        batch=dict(nfeat=9, n_classes=13, minpts=15), nodes=1024, train=True, dtype="f32"),
    "vkitti_eval": dict(
        title="configs[3]: vKITTI3D-shaped inference (PointNet embeddings + ECC), gru_10_1_1_1_0,f_13, F=9, bf16 trunk, %(nodes)d superpoints per GPU",
        args=dict(node_feats=9, ptn_nfeat_stn=9, ptn_widths=[[64, 64, 128], [64, 32, 32]],
                  ptn_widths_stn=[[32, 64], [32, 16]]),
        batch=dict(nfeat=9, n_classes=13, minpts=15), nodes=8192, train=False, dtype="bf16"),
    # configs[4]: synthetic sweep, S3DIS architecture, vector and matrix filters
    "sweep_vv": dict(
        title="configs[4]: synthetic sweep training step, gru_10_1_1_1_0,f_13, fp32, %(nodes)d superpoints per GPU",
        args=dict(), batch=dict(nfeat=14, n_classes=13, minpts=40), nodes=10000, train=True, dtype="f32"),
    "sweep_mat": dict(
        title="configs[4]: synthetic sweep training step, gru_10_0,f_13 (matrix filters), fp32, %(nodes)d superpoints per GPU",
        args=dict(model_config="gru_10_0,f_13"), batch=dict(nfeat=14, n_classes=13, minpts=40),
        nodes=10000, train=True, dtype="f32"),
}


def get(name, nodes=None):
    """-> dict(name, title, margs, nodes, train, dtype, batch_kwargs)."""
    w = WORKLOADS[name]
    n = int(nodes if nodes is not None else w["nodes"])
    return dict(name=name, margs=make_args(**w["args"]), nodes=n, train=w["train"], dtype=w["dtype"],
                batch_kwargs=dict(w["batch"]),
                title=w["title"] % dict(nodes=n, half=n // 2, quarter=n // 4))


def batch(w, seed):
    return make_batch(n_nodes=w["nodes"], seed=seed, **w["batch_kwargs"])


def oracle_cfg(margs):
    """(pcfg, mcfg) dictionaries that drive oracle/nets_ref for the same flags (tests, bench)."""
    conf = margs.model_config.split(",")[0].split("_")
    nrep = int(conf[1])
    vv = bool(int(conf[2])) if len(conf) > 2 else True
    layernorm = bool(int(conf[3])) if len(conf) > 3 else True
    ingate = bool(int(conf[4])) if len(conf) > 4 else True
    cat_all = bool(int(conf[5])) if len(conf) > 5 else True
    H = margs.ptn_widths[1][-1]
    pcfg = dict(n_conv=len(margs.ptn_widths[0]), n_fc=len(margs.ptn_widths[1]),
                n_conv_stn=len(margs.ptn_widths_stn[0]), n_fc_stn=len(margs.ptn_widths_stn[1]),
                nfeat_stn=margs.ptn_nfeat_stn)
    mcfg = dict(fnet_widths=[margs.edge_feats] + list(margs.fnet_widths) + [H if vv else H * H],
                bnidx=margs.fnet_bnidx, nrepeats=nrep, layernorm=layernorm, ingate=ingate, cat_all=cat_all)
    return pcfg, mcfg
