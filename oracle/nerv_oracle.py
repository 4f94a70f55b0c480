"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A plain restatement, on the CPU, of the reference's arithmetic for the frame-fitting hot path
(SURVEY.md section 8a, rows A1-A14).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module, and only as the checker / the timed CPU
baseline — the product path (package dir + liborepnerv.so) never touches it.

Each function cites the reference file:line (under /root/reference) it restates.  It is written with
plain torch tensor ops (fp32 by default, any dtype accepted so tests can run it in fp64), no autograd
tricks beyond `torch.autograd.grad` through the restated forward, and no import of the reference.
The functions are device-agnostic: `tests/` also runs them on CUDA tensors with TF32 switched off
(`torch.backends.cudnn.allow_tf32 = False`) as the full-size fp32 GPU oracle of SURVEY.md 8c-iv — still
library (cuDNN/cuBLAS fp32) arithmetic of the restated algorithm, never the product kernels.

Pinning: `tests/golden/make_golden.py` imports the UNMODIFIED reference `model.py` / `utils.py` in the
build container and stores small input/output vectors in `tests/golden/*.pt`; `tests/test_oracle_golden.py`
checks this oracle against them.  The SSIM / MS-SSIM part restates the third-party
`pytorch_msssim==0.2.1` (reference requirements.txt:3), which is absent from /root/reference and from the
container: **parity unpinned** for that sub-function (it is cross-checked against an independent
float64 direct-window implementation, and against the separately written stand-in of the package that the golden
generator uses — tests/golden/_shim/pytorch_msssim.py — in the tests instead).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- A1 positional encoding
def pos_encoding(pos, lbase, levels):
    """reference utils.py:121-129.  pos: [B] -> [B, 2*levels] = stack([sin, cos] per level, dim 1)."""
    pe = []
    for i in range(levels):
        v = pos * lbase ** i * math.pi
        pe += [torch.sin(v), torch.cos(v)]
    return torch.stack(pe, 1)


# ----------------------------------------------------------------------------- A2 stem
def silu(x):
    return x * torch.sigmoid(x)


def stem_forward(embed, W1, b1, W2, b2, fc_dim, fc_h, fc_w, act='swish'):
    """reference model.py:174-188 (MLP: Linear, act, Linear, act) and :612-613 (view to NCHW)."""
    h = activation(F.linear(embed, W1, b1), act)
    o = activation(F.linear(h, W2, b2), act)
    return o.view(o.size(0), fc_dim, fc_h, fc_w)


# ----------------------------------------------------------------------------- A3 ERB fold
def erb_fold(w3x3, b3x3, w1x3, b1x3, w3x1, b3x1, w1, w2, w3):
    """reference model.py:450-516.
    1x3 fills the middle ROW (F.pad(w,(0,0,1,1)), :495), 3x1 the middle COLUMN (F.pad(w,(1,1,0,0))).
    Sequential branch (:510-515): T = W2 (*) W1 over the 2*Cin axis, then W3 applied per tap."""
    k13 = torch.zeros_like(w3x3)
    k13[:, :, 1:2, :] = w1x3
    k31 = torch.zeros_like(w3x3)
    k31[:, :, :, 1:2] = w3x1
    T = torch.einsum('omhw,mi->oihw', w2, w1[:, :, 0, 0])
    kseq = torch.einsum('po,oihw->pihw', w3[:, :, 0, 0], T)
    K = w3x3 + (k13 + k31) + kseq
    b = b3x3 + (b1x3 + b3x1)
    return K, b


def erb_fold_backward(dK, db, w1, w2, w3):
    """Analytic gradient of `erb_fold` (what autograd derives for model.py:450-516)."""
    W1, W3 = w1[:, :, 0, 0], w3[:, :, 0, 0]
    T = torch.einsum('omhw,mi->oihw', w2, W1)
    g = {
        'w3x3': dK.clone(), 'w1x3': dK[:, :, 1:2, :].clone(), 'w3x1': dK[:, :, :, 1:2].clone(),
        'b3x3': db.clone(), 'b1x3': db.clone(), 'b3x1': db.clone(),
    }
    g['w3'] = torch.einsum('pihw,oihw->po', dK, T)[:, :, None, None]
    dT = torch.einsum('po,pihw->oihw', W3, dK)
    g['w2'] = torch.einsum('oihw,mi->omhw', dT, W1)
    g['w1'] = torch.einsum('omhw,oihw->mi', w2, dT)[:, :, None, None]
    return g


# ----------------------------------------------------------------------------- 8f-4: the other linear branch sets
def seqconv3x3_forward(x, k0, b0, scale, bias, mask):
    """reference model.py:277-289 (SeqConv3x3.forward): 1x1 conv with bias, border padded WITH that bias, depthwise
    3x3 with kernel scale*mask and bias."""
    y0 = F.conv2d(x, k0, b0)
    y0 = F.pad(y0, (1, 1, 1, 1), 'constant', 0)
    b0p = b0.view(1, -1, 1, 1)
    frame = torch.ones_like(y0)
    frame[:, :, 1:-1, 1:-1] = 0
    y0 = y0 * (1 - frame) + b0p * frame
    return F.conv2d(y0, scale * mask, bias, groups=k0.shape[0])


def branch_set_forward(x, sd, p, branch_type):
    """reference model.py:541-565: the EXPLICIT multi-branch forward of ACB / RepVGG / DBB / ECB (pre-PixelShuffle)."""
    c = lambda n, pad: F.conv2d(x, sd[p + n + '.weight'], sd.get(p + n + '.bias'), padding=pad)     # noqa: E731
    if branch_type == 'ACB':
        return c('rbr_3x3_branch', 1) + c('rbr_3x1_branch', (1, 0)) + c('rbr_1x3_branch', (0, 1))
    if branch_type == 'RepVGG':
        return c('rbr_3x3_branch', 1) + c('rbr_1x1_branch', 0)
    seq = F.conv2d(F.conv2d(x, sd[p + 'rbr_1x1_3x3_branch_1x1.weight']), sd[p + 'rbr_1x1_3x3_branch_3x3.weight'],
                   padding=1)
    if branch_type == 'DBB':
        avg = F.avg_pool2d(F.conv2d(x, sd[p + 'rbr_1x1_avg_branch_1x1.weight']), 3, 1, 1)
        return c('rbr_3x3_branch', 1) + c('rbr_1x1_branch', 0) + seq + avg
    if branch_type == 'ECB':
        out = c('rbr_3x3_branch', 1) + seq
        for e in EDGE_BRANCHES:
            q = p + e + '.'
            out = out + seqconv3x3_forward(x, sd[q + 'k0'], sd[q + 'b0'], sd[q + 'scale'], sd[q + 'bias'], sd[q + 'mask'])
        return out
    raise KeyError(branch_type)


EDGE_BRANCHES = ('rbr_conv1x1_sbx_branch', 'rbr_conv1x1_sby_branch', 'rbr_conv1x1_lpl_branch')


def branch_set_fold(sd, p):
    """The single 3x3 kernel + bias equal to `branch_set_forward` (every branch is linear; the reference itself has no
    fold for these sets — for SeqConv3x3 it is the closed form of its rep_params, model.py:291-300).  The branch set is
    recognised from the keys present under prefix `p`."""
    K, b = sd[p + 'rbr_3x3_branch.weight'], sd[p + 'rbr_3x3_branch.bias']
    if p + 'rbr_3x1_branch.weight' in sd:                                   # ACB
        K = K + F.pad(sd[p + 'rbr_1x3_branch.weight'], (0, 0, 1, 1)) + F.pad(sd[p + 'rbr_3x1_branch.weight'], (1, 1, 0, 0))
        b = b + sd[p + 'rbr_1x3_branch.bias'] + sd[p + 'rbr_3x1_branch.bias']
    if p + 'rbr_1x1_branch.weight' in sd:                                   # RepVGG, DBB
        K = K + F.pad(sd[p + 'rbr_1x1_branch.weight'], (1, 1, 1, 1))
        b = b + sd[p + 'rbr_1x1_branch.bias']
    if p + 'rbr_1x1_3x3_branch_1x1.weight' in sd:                           # DBB, ECB
        K = K + torch.einsum('omhw,mi->oihw', sd[p + 'rbr_1x1_3x3_branch_3x3.weight'],
                             sd[p + 'rbr_1x1_3x3_branch_1x1.weight'][:, :, 0, 0])
    if p + 'rbr_1x1_avg_branch_1x1.weight' in sd:                           # DBB
        K = K + (sd[p + 'rbr_1x1_avg_branch_1x1.weight'] / 9).expand(-1, -1, 3, 3)
    for e in EDGE_BRANCHES:                                                 # ECB
        q = p + e + '.'
        if q + 'k0' in sd:
            dw = sd[q + 'scale'] * sd[q + 'mask']                           # [Cout,1,3,3]
            K = K + dw * sd[q + 'k0']
            b = b + sd[q + 'b0'] * dw.sum((1, 2, 3)) + sd[q + 'bias']
    return K, b


# ----------------------------------------------------------------------------- A4-A6 block / head / generator
def activation(x, act='swish'):
    """reference model.py:86-117 (ActivationLayer)."""
    if act == 'swish':
        return silu(x)
    return {'relu': F.relu, 'leaky': lambda v: F.leaky_relu(v, 0.01), 'leaky01': lambda v: F.leaky_relu(v, 0.1),
            'relu6': F.relu6, 'gelu': F.gelu, 'sin': torch.sin, 'softplus': F.softplus,
            'hardswish': F.hardswish}[act](x)


def block_forward(x, K, b, stride, act='swish'):
    """reference model.py:539 (F.conv2d 3x3 pad 1) and :567 (PixelShuffle -> Identity -> activation)."""
    return activation(F.pixel_shuffle(F.conv2d(x, K, b, stride=1, padding=1), stride), act)


def head_forward(x, Wh, bh, sigmoid=False):
    """reference model.py:620-623."""
    o = F.conv2d(x, Wh, bh)
    return torch.sigmoid(o) if sigmoid else (torch.tanh(o) + 1) * 0.5


def block_kernel(sd, prefix):
    """(K, b) of block `prefix` ('layers.i.') from a state dict in train (ERB / vanilla) or deploy layout
    (key names: SURVEY.md section 5 / reference model.py:316-343)."""
    if prefix + 'rbr_reparam.weight' in sd:
        return sd[prefix + 'rbr_reparam.weight'], sd[prefix + 'rbr_reparam.bias']
    if prefix + 'branch.weight' in sd:
        return sd[prefix + 'branch.weight'], sd[prefix + 'branch.bias']
    p = prefix
    if p + 'rbr_1x1_3x3_1x1_branch_1x1_1.weight' not in sd:
        return branch_set_fold(sd, p)                                       # ACB / RepVGG / DBB / ECB
    return erb_fold(sd[p + 'rbr_3x3_branch.weight'], sd[p + 'rbr_3x3_branch.bias'],
                    sd[p + 'rbr_1x3_branch.weight'], sd[p + 'rbr_1x3_branch.bias'],
                    sd[p + 'rbr_3x1_branch.weight'], sd[p + 'rbr_3x1_branch.bias'],
                    sd[p + 'rbr_1x1_3x3_1x1_branch_1x1_1.weight'], sd[p + 'rbr_1x1_3x3_1x1_branch_3x3.weight'],
                    sd[p + 'rbr_1x1_3x3_1x1_branch_1x1_2.weight'])


def generator_forward(sd, embed, cfg, return_features=False):
    """reference model.py:611-625 for the single-resolution configuration.
    cfg: dict(fc_h, fc_w, fc_dim, strides, sigmoid)."""
    act = cfg.get('act', 'swish')
    x = stem_forward(embed, sd['stem.0.weight'], sd['stem.0.bias'], sd['stem.2.weight'], sd['stem.2.bias'],
                     cfg['fc_dim'], cfg['fc_h'], cfg['fc_w'], act)
    feats = [x]
    for i, s in enumerate(cfg['strides']):
        p = f'layers.{i}.'
        if cfg.get('explicit_branches') and (p + 'rbr_3x3_branch.weight') in sd and \
                (p + 'rbr_1x1_3x3_1x1_branch_1x1_1.weight') not in sd:
            # the way the reference itself runs ACB / RepVGG / DBB / ECB (model.py:541-565): one convolution per branch
            x = activation(F.pixel_shuffle(branch_set_forward(x, sd, p, cfg['explicit_branches']), s), act)
        else:
            K, b = block_kernel(sd, p)
            x = block_forward(x, K, b, s, act)
        feats.append(x)
    last = len(cfg['strides']) - 1
    img = head_forward(x, sd[f'head_layers.{last}.weight'], sd[f'head_layers.{last}.bias'], cfg.get('sigmoid', False))
    return (img, feats) if return_features else img


def generator_forward_multi(sd, embed, cfg):
    """reference model.py:611-625 with sin_res=False: the image of EVERY stage that owns a head (`head_layers.i.*`
    present in the state dict), in stage order."""
    _, feats = generator_forward(sd, embed, cfg, return_features=True)
    return [head_forward(feats[i + 1], sd[f'head_layers.{i}.weight'], sd[f'head_layers.{i}.bias'], cfg.get('sigmoid', False))
            for i in range(len(cfg['strides'])) if f'head_layers.{i}.weight' in sd]


def multires_loss(sd, embed, data, cfg, lw, loss_type='Fusion6'):
    """reference main_train.py:238-244: per-stage targets by adaptive average pooling, per-stage losses, weight `lw` on
    all but the last.  Returns (loss_sum, images, targets)."""
    imgs = generator_forward_multi(sd, embed, cfg)
    targets = [F.adaptive_avg_pool2d(data, x.shape[-2:]) for x in imgs]
    losses = [loss_fn(o, t, loss_type) for o, t in zip(imgs, targets)]
    losses = [l * (lw if i < len(losses) - 1 else 1) for i, l in enumerate(losses)]
    return sum(losses), imgs, targets


# ----------------------------------------------------------------------------- A7/A10 SSIM family
def _gauss_1d(size=11, sigma=1.5, dtype=torch.float32):
    """pytorch_msssim 0.2.1 `_fspecial_gauss_1d`."""
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _gauss_filter(x, win):
    """pytorch_msssim 0.2.1 `gaussian_filter`: separable VALID filtering, H pass then W pass, per channel."""
    C = x.shape[1]
    w = win.to(device=x.device, dtype=x.dtype).view(1, 1, -1)
    out = F.conv2d(x, w.view(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)
    return F.conv2d(out, w.view(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)


def ssim_maps(X, Y, data_range=1.0, K=(0.01, 0.03)):
    """pytorch_msssim 0.2.1 `_ssim` up to the per-pixel maps. Returns (ssim_map, cs_map)."""
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    win = _gauss_1d(dtype=X.dtype)
    mu1, mu2 = _gauss_filter(X, win), _gauss_filter(Y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _gauss_filter(X * X, win) - mu1_sq
    s2 = _gauss_filter(Y * Y, win) - mu2_sq
    s12 = _gauss_filter(X * Y, win) - mu1_mu2
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    return ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs, cs


def ssim(X, Y):
    """pytorch_msssim.ssim(X, Y, data_range=1, size_average=True) as called at reference utils.py:160."""
    m, _ = ssim_maps(X, Y)
    return torch.flatten(m, 2).mean(-1).mean()


def ms_ssim(X, Y):
    """pytorch_msssim.ms_ssim(X, Y, data_range=1, size_average=True) as called at reference utils.py:205."""
    weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=X.dtype, device=X.device)
    mcs = []
    for i in range(5):
        m, cs = ssim_maps(X, Y)
        ssim_pc, cs_pc = torch.flatten(m, 2).mean(-1), torch.flatten(cs, 2).mean(-1)
        if i < 4:
            mcs.append(torch.relu(cs_pc))
            padding = [s % 2 for s in X.shape[2:]]
            X = F.avg_pool2d(X, kernel_size=2, padding=padding)
            Y = F.avg_pool2d(Y, kernel_size=2, padding=padding)
    stack = torch.stack(mcs + [torch.relu(ssim_pc)], dim=0)
    return torch.prod(stack ** weights.view(-1, 1, 1), dim=0).mean()


def ssim_direct_f64(X, Y):
    """Independent float64 SSIM: explicit 11x11 window sums per output pixel (no separable filtering, no
    conv).  Used to cross-check `ssim`, whose third-party original is not available (parity unpinned)."""
    X, Y = X.double(), Y.double()
    g = _gauss_1d(dtype=torch.float64)
    w2 = torch.outer(g, g)
    B, C, H, W = X.shape
    Hv, Wv = H - 10, W - 10
    px = X.unfold(2, 11, 1).unfold(3, 11, 1)        # [B,C,Hv,Wv,11,11]
    py = Y.unfold(2, 11, 1).unfold(3, 11, 1)
    m1, m2 = (px * w2).sum((-1, -2)), (py * w2).sum((-1, -2))
    s1 = (px * px * w2).sum((-1, -2)) - m1 * m1
    s2 = (py * py * w2).sum((-1, -2)) - m2 * m2
    s12 = (px * py * w2).sum((-1, -2)) - m1 * m2
    C1, C2 = 1e-4, 9e-4
    m = ((2 * m1 * m2 + C1) / (m1 * m1 + m2 * m2 + C1)) * ((2 * s12 + C2) / (s1 + s2 + C2))
    assert m.shape[-2:] == (Hv, Wv)
    return m.mean()


# loss_type -> weights of (mean |p-t|, mean (p-t)^2, 1 - SSIM): reference utils.py:142-166
LOSS_WEIGHTS = {'L2': (0.0, 1.0, 0.0), 'L1': (1.0, 0.0, 0.0), 'SSIM': (0.0, 0.0, 1.0),
                'Fusion1': (0.0, 0.3, 0.7), 'Fusion2': (0.3, 0.0, 0.7), 'Fusion3': (0.0, 0.5, 0.5),
                'Fusion4': (0.5, 0.0, 0.5), 'Fusion5': (0.0, 0.7, 0.3), 'Fusion6': (0.7, 0.0, 0.3),
                'Fusion7': (0.3, 0.7, 0.0), 'Fusion8': (0.5, 0.5, 0.0), 'Fusion9': (0.9, 0.0, 0.1)}


def loss_fn(pred, target, loss_type='Fusion6'):
    """reference utils.py:139-166 for the L1 / L2 / SSIM combinations."""
    w_l1, w_mse, w_ssim = LOSS_WEIGHTS[loss_type]
    target = target.detach()
    loss = 0
    if w_l1:
        loss = loss + w_l1 * torch.mean(torch.abs(pred - target))
    if w_mse:
        loss = loss + w_mse * F.mse_loss(pred, target)
    if w_ssim:
        loss = loss + w_ssim * (1 - ssim(pred, target))
    return loss


def psnr(output, target):
    """reference utils.py:194-195."""
    return -10 * torch.log10(F.mse_loss(output, target, reduction='mean'))


# ----------------------------------------------------------------------------- A9 LR schedule / Adam
def lr_at(epoch, it, data_size, lr, warmup, epochs, lr_type='cosine'):
    """reference utils.py:240-259 (warmup already int(ratio*epochs), main_train.py:111)."""
    e = epoch + float(it) / data_size
    if lr_type == 'cosine':
        mult = 0.5 * (math.cos(math.pi * (e - warmup) / (epochs - warmup)) + 1.0)
    elif lr_type == 'const':
        mult = 1
    else:
        raise NotImplementedError
    if e < warmup:
        mult = 0.1 + 0.9 * e / warmup
    return lr * mult


def adam_step(p, g, m, v, t, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay / amsgrad) as used at reference main_train.py:196, :250."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    step_size = lr / (1 - beta1 ** t)
    denom = v.sqrt() / math.sqrt(1 - beta2 ** t) + eps
    return p - step_size * m / denom, m, v


# ----------------------------------------------------------------------------- A12/A13 prune / quantise
def prune_threshold(tensors, amount):
    """k-th smallest |w| (k = round(amount*N)) over all tensors: the cut torch.nn.utils.prune's
    global_unstructured(L1Unstructured) applies (reference main_eval.py:587)."""
    flat = torch.cat([t.reshape(-1).abs() for t in tensors])
    k = int(round(amount * flat.numel()))
    if k == 0:
        return None, 0
    return torch.kthvalue(flat, k).values.item(), k


def quantize_per_tensor(t, bit=8, axis=-1):
    """reference utils.py:11-67 restated without the Python row loop."""
    if axis == -1:
        nz = t[t != 0]
        t_min, t_max = nz.min(), nz.max()
        scale = (t_max - t_min) / 2 ** bit
    else:
        tm = t.transpose(0, axis).reshape(t.size(axis), -1) if axis else t.reshape(t.size(0), -1)
        big = torch.finfo(t.dtype).max
        mn = torch.where(tm != 0, tm, torch.full_like(tm, big)).min(1).values
        mx = torch.where(tm != 0, tm, torch.full_like(tm, -big)).max(1).values
        empty = (tm != 0).sum(1) == 0
        mn = torch.where(empty, torch.zeros_like(mn), mn)
        mx = torch.where(empty, torch.zeros_like(mx), mx)
        shape = [1] * t.dim()
        shape[axis] = -1
        t_min = mn.view(shape)
        scale = ((mx - mn) / 2 ** bit).view(shape)
    q = ((t - t_min) / (scale + 1e-19)).round()
    return q, t_min + scale * q


# ----------------------------------------------------------------------------- one training step (A8)
def train_step(sd, opt_state, embed, target, cfg, lr, t, loss_type='Fusion6', beta1=0.5):
    """One iteration of reference main_train.py:238-250 on a state dict (dict name -> tensor):
    forward, Fusion loss, backward (autograd through the restated forward), Adam.  Returns
    (new_sd, new_opt_state, loss, img, grads)."""
    # SeqConv3x3.mask is an nn.Parameter with requires_grad=False (model.py:222): a constant of the model
    params = {k: (v.detach().clone() if k.endswith('.mask') else v.detach().clone().requires_grad_(True))
              for k, v in sd.items()}
    img = generator_forward(params, embed, cfg)
    loss = loss_fn(img, target, loss_type)
    names = [k for k in params if params[k].requires_grad]
    const = {k: params[k] for k in params if not params[k].requires_grad}
    grads = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
    new_sd, new_state, gdict = {}, {}, {}
    for n, g in zip(names, grads):
        g = torch.zeros_like(params[n]) if g is None else g
        m, v = opt_state.get(n, (torch.zeros_like(g), torch.zeros_like(g)))
        p, m, v = adam_step(params[n].detach(), g, m, v, t, lr, beta1=beta1)
        new_sd[n], new_state[n], gdict[n] = p, (m, v), g
    new_sd.update(const)
    new_sd = {k: new_sd[k] for k in sd}                   # keep the state-dict order
    return new_sd, new_state, loss.detach(), img.detach(), gdict


# ----------------------------------------------------------------------------- 8f-3 prune-then-finetune
def finetune_steps(sd, masks, frozen, embed, target, cfg, lrs, loss_type='Fusion6', beta1=0.5):
    """reference main_eval.py:446-507 on a state dict: `masks` (name -> 0/1 tensor) are the weight_mask of
    torch.nn.utils.prune — the forward reads weight_orig * mask, so the gradient of weight_orig is the masked gradient;
    `frozen` names never move (the ERB quirk, SURVEY.md 2.1 row 19: an ERB block reads the `.weight` computed at prune
    time, so its pruned branch kernels stay at weight_orig * mask of that moment whatever Adam does to weight_orig).
    Fresh Adam state (:426, :489-490).  `sd` holds weight_orig; returns (sd of weight_orig, effective weights, losses)."""
    sd = {k: v.clone() for k, v in sd.items()}
    stale = {n: sd[n] * masks[n] for n in frozen}
    state, losses = {}, []
    for t, lr in enumerate(lrs, 1):
        params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
        # frozen: VALUE of the stale tensor, GRADIENT still reaching weight_orig through the retained graph of
        # weight_orig * mask (main_eval.py:476-480 retain_graph=True) — Adam keeps moving the unused weight_orig
        eff = {k: ((stale[k] + (params[k] * masks[k] - (params[k] * masks[k]).detach())) if k in frozen else
                   (params[k] * masks[k] if k in masks else params[k])) for k in params}
        loss = loss_fn(generator_forward(eff, embed, cfg), target, loss_type)
        names = list(params)
        grads = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
        for n, g in zip(names, grads):
            if g is None:
                continue                    # the reference's Adam skips parameters without a gradient
            m, v = state.get(n, (torch.zeros_like(g), torch.zeros_like(g)))
            sd[n], m, v = adam_step(sd[n], g, m, v, t, lr, beta1=beta1)
            state[n] = (m, v)
        losses.append(float(loss.detach()))
    eff = {k: (stale[k] if k in frozen else (sd[k] * masks[k] if k in masks else sd[k])) for k in sd}
    return sd, eff, losses


# ----------------------------------------------------------------------------- reference-shaped random state
def random_state(cfg, w, seed=1, deploy=False, branch_type='ERB'):
    """A random state dict with the reference's parameter names / shapes (model.py:571-609, :316-343) drawn from the
    nn.Linear / nn.Conv2d default distribution U(-1/sqrt(fan_in), 1/sqrt(fan_in)) — for the timed CPU / library legs
    of bench.py, which must not import the package.  cfg: dict(fc_h, fc_w, fc_dim, strides); w: dict(stem_dim_num,
    expansion, reduction, lower_width).  `deploy` gives the single-branch layout (rbr_reparam)."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    stem_dim = int(str(w['stem_dim_num']).split('_')[0])
    fd, fh, fw = cfg['fc_dim'], cfg['fc_h'], cfg['fc_w']
    sd = {'stem.0.weight': u((stem_dim, 80), 80), 'stem.0.bias': u((stem_dim,), 80),
          'stem.2.weight': u((fd * fh * fw, stem_dim), stem_dim), 'stem.2.bias': u((fd * fh * fw,), stem_dim)}
    c = fd
    for i, s in enumerate(cfg['strides']):
        cnew = int(c * w['expansion']) if i == 0 else max(c // w['reduction'], w['lower_width'])
        co, p = cnew * s * s, f'layers.{i}.'
        if deploy:
            sd[p + 'rbr_reparam.weight'], sd[p + 'rbr_reparam.bias'] = u((co, c, 3, 3), 9 * c), u((co,), 9 * c)
        elif branch_type == 'NeRV_vanilla':
            sd[p + 'branch.weight'], sd[p + 'branch.bias'] = u((co, c, 3, 3), 9 * c), u((co,), 9 * c)
        elif branch_type in ('ACB', 'RepVGG', 'DBB', 'ECB'):                       # model.py:345-393
            sd[p + 'rbr_3x3_branch.weight'], sd[p + 'rbr_3x3_branch.bias'] = u((co, c, 3, 3), 9 * c), u((co,), 9 * c)
            if branch_type == 'ACB':
                sd[p + 'rbr_3x1_branch.weight'], sd[p + 'rbr_3x1_branch.bias'] = u((co, c, 3, 1), 3 * c), u((co,), 3 * c)
                sd[p + 'rbr_1x3_branch.weight'], sd[p + 'rbr_1x3_branch.bias'] = u((co, c, 1, 3), 3 * c), u((co,), 3 * c)
            if branch_type in ('RepVGG', 'DBB'):
                sd[p + 'rbr_1x1_branch.weight'], sd[p + 'rbr_1x1_branch.bias'] = u((co, c, 1, 1), c), u((co,), c)
            if branch_type in ('DBB', 'ECB'):
                sd[p + 'rbr_1x1_3x3_branch_1x1.weight'] = u((2 * c, c, 1, 1), c)
                sd[p + 'rbr_1x1_3x3_branch_3x3.weight'] = u((co, 2 * c, 3, 3), 18 * c)
            if branch_type == 'DBB':
                sd[p + 'rbr_1x1_avg_branch_1x1.weight'] = u((co, c, 1, 1), c)
            if branch_type == 'ECB':
                masks = {'sbx': [[1, 0, -1], [2, 0, -2], [1, 0, -1]], 'sby': [[1, 2, 1], [0, 0, 0], [-1, -2, -1]],
                         'lpl': [[0, 1, 0], [1, -4, 1], [0, 1, 0]]}
                for e, mk in masks.items():
                    q = p + f'rbr_conv1x1_{e}_branch.'
                    sd[q + 'k0'], sd[q + 'b0'] = u((co, c, 1, 1), c), u((co,), c)
                    sd[q + 'scale'] = torch.randn((co, 1, 1, 1), generator=g) * 1e-3
                    sd[q + 'bias'] = torch.randn((co,), generator=g) * 1e-3
                    sd[q + 'mask'] = torch.tensor(mk, dtype=torch.float32).view(1, 1, 3, 3).repeat(co, 1, 1, 1)
        else:
            sd[p + 'rbr_3x3_branch.weight'], sd[p + 'rbr_3x3_branch.bias'] = u((co, c, 3, 3), 9 * c), u((co,), 9 * c)
            sd[p + 'rbr_3x1_branch.weight'], sd[p + 'rbr_3x1_branch.bias'] = u((co, c, 3, 1), 3 * c), u((co,), 3 * c)
            sd[p + 'rbr_1x3_branch.weight'], sd[p + 'rbr_1x3_branch.bias'] = u((co, c, 1, 3), 3 * c), u((co,), 3 * c)
            sd[p + 'rbr_1x1_3x3_1x1_branch_1x1_1.weight'] = u((2 * c, c, 1, 1), c)
            sd[p + 'rbr_1x1_3x3_1x1_branch_3x3.weight'] = u((co, 2 * c, 3, 3), 18 * c)
            sd[p + 'rbr_1x1_3x3_1x1_branch_1x1_2.weight'] = u((co, co, 1, 1), co)
        c = cnew
    last = len(cfg['strides']) - 1
    sd[f'head_layers.{last}.weight'], sd[f'head_layers.{last}.bias'] = u((3, c, 1, 1), c), u((3,), c)
    return sd
