"""ORACLE (test infrastructure, never imported by the product path).

Functional CPU restatement (torch.nn.functional on explicit state dicts) of the dense part of the
reference's learning path: STNkD / PointNet, the filter-generating network, GRUCellEx, the
recurrent ECC module and GraphNetwork, plus one full training step.  State-dict keys are the
reference's, so the same dict drives the reference modules, this oracle and the
CUDA modules.  Pinned against the reference by tests/golden/make_golden.py.
"""
import torch
import torch.nn.functional as F

from . import ecc_ref


# ------------------------------------------------------------------------------ building blocks
def _bn(x, sd, key, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm1d as instantiated at learning/pointnet.py:31,43,87,103, graphnet.py:29."""
    return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'], sd[key + '.weight'],
                        sd[key + '.bias'], training, momentum, eps)


def conv_stack(x, sd, prefix, n_layers, training):
    """[Conv1d(k=1), BatchNorm1d, ReLU] * n (learning/pointnet.py:27-37, 83-96); x is [B,C,L]."""
    for i in range(n_layers):
        x = F.conv1d(x, sd['%s%d.weight' % (prefix, 3 * i)], sd['%s%d.bias' % (prefix, 3 * i)])
        x = F.relu(_bn(x, sd, '%s%d' % (prefix, 3 * i + 1), training))
    return x


def fc_stack(x, sd, prefix, n_layers, training, last_ac=True):
    """[Linear, BatchNorm1d, ReLU] * n with the activation of the last layer optional
    (learning/pointnet.py:39-49, 98-110; prelast_do = 0)."""
    for i in range(n_layers):
        x = F.linear(x, sd['%s%d.weight' % (prefix, 3 * i)], sd['%s%d.bias' % (prefix, 3 * i)])
        if i < n_layers - 1 or last_ac:
            x = F.relu(_bn(x, sd, '%s%d' % (prefix, 3 * i + 1), training))
    return x


def stn_forward(x, sd, prefix, n_conv, n_fc, training, K=2):
    """STNkD.forward, learning/pointnet.py:55-61: convs, max over points, fcs, proj, + identity."""
    x = conv_stack(x, sd, prefix + 'convs.', n_conv, training)
    x = F.max_pool1d(x, x.size(2)).squeeze(2)
    x = fc_stack(x, sd, prefix + 'fcs.', n_fc, training, last_ac=True)
    x = F.linear(x, sd[prefix + 'proj.weight'], sd[prefix + 'proj.bias'])
    return x.view(-1, K, K) + torch.eye(K, dtype=x.dtype).unsqueeze(0)


def pointnet_forward(x, x_global, sd, cfg, training, prefix=''):
    """PointNet.forward, learning/pointnet.py:120-133.
    cfg: dict(n_conv, n_fc, n_conv_stn, n_fc_stn, nfeat_stn)."""
    if cfg['nfeat_stn'] > 0:
        T = stn_forward(x[:, :cfg['nfeat_stn'], :], sd, prefix + 'stn.', cfg['n_conv_stn'],
                        cfg['n_fc_stn'], training)
        xy = torch.bmm(x[:, :2, :].transpose(1, 2), T).transpose(1, 2)      # :123
        x = torch.cat([xy, x[:, 2:, :]], 1)                                 # :124
    x = conv_stack(x, sd, prefix + 'convs.', cfg['n_conv'], training)
    x = F.max_pool1d(x, x.size(2)).squeeze(2)                               # :127
    if x_global is not None:
        x = torch.cat([x, x_global.view(x.shape[0], -1)], 1)                # :128-132
    return fc_stack(x, sd, prefix + 'fcs.', cfg['n_fc'], training, last_ac=False)


def pointnet_forward_ragged(points, offsets, x_global, sd, cfg, training, prefix=''):
    """PointNet.forward (learning/pointnet.py:120-133) WITHOUT the loader's resample-to-ptn_npts
    (spg.py:209-214): `points` [P,F] of all superpoints back to back, `offsets` [B+1].  The 1x1 convolutions
    and their BatchNorm see all points at once ([1,F,P]: BatchNorm1d reduces over batch and length alike),
    every max-pool runs over one superpoint's own points.  No reference counterpart exists for unequal
    segment lengths (parity unpinned there); with equal lengths it IS pointnet_forward (tested)."""
    B = len(offsets) - 1
    x = points.t().unsqueeze(0)                                             # [1,F,P]

    def segmax(y):                                                          # [1,C,P] -> [B,C]
        return torch.stack([y[0, :, int(offsets[b]):int(offsets[b + 1])].max(1)[0] for b in range(B)])

    if cfg['nfeat_stn'] > 0:
        h = conv_stack(x[:, :cfg['nfeat_stn'], :], sd, prefix + 'stn.convs.', cfg['n_conv_stn'], training)
        h = fc_stack(segmax(h), sd, prefix + 'stn.fcs.', cfg['n_fc_stn'], training, last_ac=True)
        T = F.linear(h, sd[prefix + 'stn.proj.weight'], sd[prefix + 'stn.proj.bias']).view(-1, 2, 2)
        T = T + torch.eye(2, dtype=T.dtype).unsqueeze(0)
        seg = torch.repeat_interleave(torch.arange(B), torch.as_tensor(offsets[1:]) - torch.as_tensor(offsets[:-1]))
        xy = torch.bmm(points[:, None, :2], T[seg]).squeeze(1)              # row vector times T (:123)
        x = torch.cat([xy, points[:, 2:]], 1).t().unsqueeze(0)
    h = segmax(conv_stack(x, sd, prefix + 'convs.', cfg['n_conv'], training))
    if x_global is not None:
        h = torch.cat([h, x_global.view(B, -1)], 1)
    return fc_stack(h, sd, prefix + 'fcs.', cfg['n_fc'], training, last_ac=False)


def cloud_embed(clouds, clouds_global, clouds_flag, sd, cfg, training, prefix=''):
    """CloudEmbedder.run_full, learning/pointnet.py:147-158: PointNet on the valid clouds,
    scattered into zero descriptors."""
    idx_valid = torch.nonzero(clouds_flag.eq(0)).reshape(-1)
    out = pointnet_forward(clouds, clouds_global, sd, cfg, training, prefix)
    desc = out.new_zeros((clouds_flag.numel(), out.shape[1]))
    return desc.index_copy(0, idx_valid, out)


def fnet_forward(ef, sd, prefix, widths, bnidx, training):
    """create_fnet, learning/graphnet.py:17-34: Linear(+BN at bnidx)+ReLU ... Linear."""
    idx = 0
    x = ef
    n_hidden = len(widths) - 2
    for k in range(n_hidden):
        x = F.linear(x, sd['%s%d.weight' % (prefix, idx)], sd['%s%d.bias' % (prefix, idx)])
        idx += 1
        if bnidx == k:
            x = _bn(x, sd, '%s%d' % (prefix, idx), training)
            idx += 1
        x = F.relu(x)
        idx += 1
    x = F.linear(x, sd['%s%d.weight' % (prefix, idx)], sd.get('%s%d.bias' % (prefix, idx)))
    if bnidx == len(widths) - 1:
        x = _bn(x, sd, '%s%d' % (prefix, idx + 1), training)
    return x


def gru_cell_ex(x, h, sd, prefix, layernorm=True, ingate=True):
    """GRUCellEx.forward, learning/modules.py:224-251."""
    if ingate:
        x = torch.sigmoid(F.linear(h, sd[prefix + 'ig.weight'], sd[prefix + 'ig.bias'])) * x  # :226
    gi = F.linear(x, sd[prefix + 'weight_ih'])                                                # :239
    gh = F.linear(h, sd[prefix + 'weight_hh'])                                                # :240
    if layernorm:                                                                             # :218-222
        gi = F.instance_norm(gi.unsqueeze(1), eps=1e-5).squeeze(1)
        gh = F.instance_norm(gh.unsqueeze(1), eps=1e-5).squeeze(1)
    i_r, i_i, i_n = gi.chunk(3, 1)
    h_r, h_i, h_n = gh.chunk(3, 1)
    b_ir, b_ii, b_in = sd[prefix + 'bias_ih'].chunk(3)
    b_hr, b_hi, b_hn = sd[prefix + 'bias_hh'].chunk(3)
    r = torch.sigmoid(i_r + b_ir + h_r + b_hr)                                                # :247
    z = torch.sigmoid(i_i + b_ii + h_i + b_hi)                                                # :248
    n = torch.tanh(i_n + b_in + r * (h_n + b_hn))                                             # :249
    return n + z * (h - n)                                                                    # :250


def rnn_ecc_forward(hx, edgefeats, idxn, degs, sd, prefix, mcfg, training, ecc_mode='vec'):
    """RNNGraphConvModule.forward, learning/modules.py:152-183.
    mcfg: dict(fnet_widths (incl. in/out), bnidx, nrepeats, layernorm, ingate, cat_all)."""
    w = fnet_forward(edgefeats, sd, prefix + '_fnet.', mcfg['fnet_widths'], mcfg['bnidx'], training)
    nc = hx.size(1)
    if w.size(1) != nc:
        w = w.view(-1, nc, nc)                                                                # :164
    hxs = [hx]
    for _ in range(mcfg['nrepeats']):                                                         # :171
        if ecc_mode == 'loop':
            inp = ecc_ref.GraphConvLoop.apply(hx, w, idxn, degs)
        else:
            inp = ecc_ref.graph_conv_forward(hx, w, idxn, None, degs)                         # :175
        hx = gru_cell_ex(inp, hx, sd, prefix + '_cell.', mcfg['layernorm'], mcfg['ingate'])   # :180
        hxs.append(hx)
    return torch.cat(hxs, 1) if mcfg['cat_all'] else hx                                      # :183


def graphnet_forward(emb, edgefeats, idxn, degs, sd, mcfg, training, prefix='', ecc_mode='vec'):
    """GraphNetwork.forward for '<gru_...>,f_<classes>' configs (learning/graphnet.py:95-98)."""
    x = rnn_ecc_forward(emb, edgefeats, idxn, degs, sd, prefix + '0.', mcfg, training, ecc_mode)
    return F.linear(x, sd[prefix + '1.weight'], sd[prefix + '1.bias'])


# ----------------------------------------------------------------------------------- full step
def spg_forward(batch, sd_ptn, sd_ecc, pcfg, mcfg, training, ecc_mode='vec'):
    """learning/main.py:202-203: embeddings = CloudEmbedder.run(...); outputs = model.ecc(...)."""
    emb = cloud_embed(batch['clouds'], batch['clouds_global'], batch['clouds_flag'], sd_ptn, pcfg,
                      training)
    return graphnet_forward(emb, batch['edgefeats'], batch['idxn'], batch['degs'], sd_ecc, mcfg,
                            training, ecc_mode=ecc_mode)


def is_param(key):
    return not (key.endswith('running_mean') or key.endswith('running_var') or
                key.endswith('num_batches_tracked'))


class RefTrainer(object):
    """One reference training step on CPU: forward, weighted CE, backward, element-wise gradient
    clamp, Adam (learning/main.py:199-213, 433-437)."""

    def __init__(self, sd_ptn, sd_ecc, pcfg, mcfg, lr=1e-2, grad_clip=1.0, class_weights=None,
                 ecc_mode='vec'):
        self.sd_ptn = {k: v.clone() for k, v in sd_ptn.items()}
        self.sd_ecc = {k: v.clone() for k, v in sd_ecc.items()}
        self.pcfg, self.mcfg, self.grad_clip = pcfg, mcfg, grad_clip
        self.class_weights, self.ecc_mode = class_weights, ecc_mode
        self.params = []
        # main.py:421-425 registers model.ecc before model.ptn
        for sd in (self.sd_ecc, self.sd_ptn):
            for k, v in sd.items():
                if is_param(k):
                    v.requires_grad_(True)
                    self.params.append(v)
        self.opt = torch.optim.Adam(self.params, lr=lr)

    def step(self, batch):
        self.opt.zero_grad()
        out = spg_forward(batch, self.sd_ptn, self.sd_ecc, self.pcfg, self.mcfg, True, self.ecc_mode)
        loss = F.cross_entropy(out, batch['labels'], weight=self.class_weights)
        loss.backward()
        if self.grad_clip > 0:
            for p in self.params:
                p.grad.clamp_(-self.grad_clip, self.grad_clip)
        self.opt.step()
        return float(loss.detach()), out.detach()
