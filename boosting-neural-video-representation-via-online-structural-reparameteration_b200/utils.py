"""Host-side mirror of the reference `utils.py` for the frame-fitting hot path.

Same names, argument meaning and error behaviour as reference utils.py (PositionalEncoding :110-129,
loss_fn :139-189, psnr_fn :191-199, msssim_fn :201-211, adjust_lr :240-259, quantize_per_tensor :11-67,
RoundTensor :213-238); the arithmetic runs in liborepnerv.so.  Loss types built from L1 / MSE / SSIM terms (L2, L1, SSIM, Fusion1..9,
reference utils.py:142-166) run on the device; the MS-SSIM-loss and FFT variants (Fusion10..15, :167-188) raise
NotImplementedError.
"""
import math
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr

# loss_type -> (weight of mean|p-t|, weight of mean (p-t)^2, weight of (1-SSIM)); reference utils.py:142-166
LOSS_TERMS = {
    'L2': (0.0, 1.0, 0.0), 'L1': (1.0, 0.0, 0.0), 'SSIM': (0.0, 0.0, 1.0),
    'Fusion1': (0.0, 0.3, 0.7), 'Fusion2': (0.3, 0.0, 0.7), 'Fusion3': (0.0, 0.5, 0.5), 'Fusion4': (0.5, 0.0, 0.5),
    'Fusion5': (0.0, 0.7, 0.3), 'Fusion6': (0.7, 0.0, 0.3), 'Fusion7': (0.3, 0.7, 0.0), 'Fusion8': (0.5, 0.5, 0.0),
    'Fusion9': (0.9, 0.0, 0.1),
}
_L1_SSIM_LOSSES = LOSS_TERMS          # former name (round 1)
# (pred ptr, pred version, target ptr, target version) -> out5 of the last loss / stats evaluation: the reference loop
# calls loss_fn, psnr_fn and msssim_fn on the SAME (output, target) pair every step (main_train.py:242, :253-254); the
# second and third call reuse the first one's MSE / scale-0 SSIM statistics instead of filtering the frame again.
_last_stats = {}
_workspaces = {}


def _workspace(kind, nbytes, device):
    key = (kind, str(device))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _cuda_f32(t):
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    return t.detach().to(torch.float32).contiguous()


class PositionalEncoding(nn.Module):
    """Reference utils.py:110-129.  forward(pos[B]) -> [B, 2*levels] on the GPU (so the caller's
    `.cuda(non_blocking=True)`, main_train.py:235, is a no-op)."""

    def __init__(self, pe_embed):
        super().__init__()
        self.pe_embed = pe_embed.lower()
        if self.pe_embed == 'none':
            self.embed_length = 1
        else:
            self.lbase, self.levels = [float(x) for x in pe_embed.split('_')]
            self.levels = int(self.levels)
            self.embed_length = 2 * self.levels
            # lbase**i in Python double, rounded to fp32 when it meets the fp32 tensor — as in the reference
            self._freqs_host = torch.tensor([self.lbase ** i for i in range(self.levels)], dtype=torch.float64
                                            ).to(torch.float32)
            self._freqs_dev = {}

    def freqs(self, device):
        key = str(device)
        if key not in self._freqs_dev:
            self._freqs_dev[key] = self._freqs_host.to(device)
        return self._freqs_dev[key]

    def forward(self, pos):
        if self.pe_embed == 'none':
            return pos[:, None]
        pos = _cuda_f32(pos).reshape(-1)
        out = torch.empty(pos.numel(), self.embed_length, dtype=torch.float32, device=pos.device)
        check(_lib.lib().onr_pos_encoding(ptr(pos), pos.numel(), ptr(self.freqs(pos.device)), self.levels,
                                          ptr(out), _lib.stream()), "onr_pos_encoding")
        return out


class _StatsKey:
    """Identity (not address: the caching allocator recycles addresses) + version of a (pred, target) pair."""

    def __init__(self, pred, target):
        self.p, self.t = weakref.ref(pred), weakref.ref(target)
        self.v = (pred._version, target._version)

    def matches(self, pred, target):
        return self.p() is pred and self.t() is target and self.v == (pred._version, target._version)


def _cached_stats(pred, target):
    c = _last_stats.get(str(pred.device))
    return c if (c is not None and c[0].matches(pred, target)) else None


class _FusionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, w_l1, w_mse, w_ssim):
        lib = _lib.lib()
        B, C, H, W = pred.shape
        if C != 3:
            raise NotImplementedError("loss kernels expect 3-channel images")
        p, t = pred.detach().contiguous(), _cuda_f32(target)
        out5 = torch.empty(5, dtype=torch.float32, device=p.device)
        need_grad = pred.requires_grad
        grad = torch.empty_like(p) if need_grad else None
        work = _workspace("loss", lib.onr_loss_workspace_bytes(B, H, W), p.device)
        check(lib.onr_fusion_loss(ptr(p), ptr(t), B, H, W, w_l1, w_mse, w_ssim, 1.0, ptr(out5), ptr(grad),
                                  ptr(work), _lib.stream()), "onr_fusion_loss")
        ctx.grad = grad
        ctx.mark_non_differentiable(out5)
        return out5[0], out5

    @staticmethod
    def backward(ctx, gloss, _gout5):
        # dloss/dpred was produced by the forward pass with unit upstream gradient; apply autograd's upstream scalar
        # (a device tensor: no host sync) in place
        g = ctx.grad
        gl = gloss.detach().to(torch.float32).contiguous()
        check(_lib.lib().onr_scale_by_device_scalar(ptr(g), g.numel(), ptr(gl), _lib.stream()),
              "onr_scale_by_device_scalar")
        return g, None, None, None, None


def loss_fn(pred, target, args):
    """Reference utils.py:139-189 (`target.detach()` included)."""
    lt = args.loss_type
    if lt not in LOSS_TERMS:
        raise NotImplementedError(
            f"loss_type {lt!r}: only the L1 / L2 / SSIM combinations {sorted(LOSS_TERMS)} run on the B200 hot path")
    w_l1, w_mse, w_ssim = LOSS_TERMS[lt]
    _last_stats.pop(str(pred.device), None)
    loss, out5 = _FusionLoss.apply(pred, target.detach(), w_l1, w_mse, w_ssim)
    if pred.is_cuda and target.is_cuda and pred.dtype == target.dtype == torch.float32 and pred.is_contiguous() \
            and target.is_contiguous():
        # the kernels read exactly these two tensors: later psnr_fn / msssim_fn calls on them reuse the statistics
        _last_stats[str(pred.device)] = (_StatsKey(pred, target), out5, _workspaces[("loss", str(pred.device))])
    return loss


def adaptive_avg_pool2d(data, size):
    """F.adaptive_avg_pool2d(data, size) on the device (reference main_train.py:239: the frame pooled to the
    resolution of each head).  The identity when the size already matches (the single-resolution case)."""
    Ho, Wo = int(size[0]), int(size[1])
    if tuple(data.shape[-2:]) == (Ho, Wo):
        return data
    x = _cuda_f32(data)
    B, C, H, W = x.shape
    out = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=x.device)
    check(_lib.lib().onr_adaptive_avg_pool(ptr(x), B * C, H, W, Ho, Wo, ptr(out), _lib.stream()),
          "onr_adaptive_avg_pool")
    return out


def frame_stats(pred, target):
    """out5 = [Fusion6 loss, L1, SSIM, MSE, PSNR] of a [B,3,H,W] pair, no gradient."""
    lib = _lib.lib()
    cached = _cached_stats(pred, target)
    if cached is not None:
        return cached[1]                       # L1 / SSIM / MSE / PSNR do not depend on the loss weights
    p, t = _cuda_f32(pred), _cuda_f32(target)
    B, _, H, W = p.shape
    out5 = torch.empty(5, dtype=torch.float32, device=p.device)
    work = _workspace("loss", lib.onr_loss_workspace_bytes(B, H, W), p.device)
    check(lib.onr_fusion6_fwd_bwd(ptr(p), ptr(t), B, H, W, 0.7, 0.3, 1.0, ptr(out5), None, ptr(work),
                                  _lib.stream()), "onr_fusion6_fwd_bwd")
    _last_stats[str(p.device)] = (_StatsKey(pred, target), out5, work)
    return out5


def psnr_fn(output_list, target_list):
    """Reference utils.py:191-199: -10 log10(MSE over the whole batch), shape (batch, num_stage)."""
    psnr_list = []
    for output, target in zip(output_list, target_list):
        psnr = frame_stats(output, target)[4]
        psnr_list.append(psnr.view(1, 1).expand(output.size(0), -1))
    return torch.cat(psnr_list, dim=1)


def msssim_fn(output_list, target_list):
    """Reference utils.py:201-211: ms_ssim(size_average=True) when H >= 160 else 0."""
    lib = _lib.lib()
    vals = []
    for output, target in zip(output_list, target_list):
        if output.size(-2) >= 160:
            p, t = _cuda_f32(output), _cuda_f32(target)
            B, _, H, W = p.shape
            out1 = torch.empty(1, dtype=torch.float32, device=p.device)
            work = _workspace("msssim", lib.onr_msssim_workspace_bytes(B, H, W), p.device)
            cached = _cached_stats(output, target)
            loss_work = cached[2] if cached is not None else None
            check(lib.onr_msssim(ptr(p), ptr(t), B, H, W, ptr(out1), ptr(work), ptr(loss_work), _lib.stream()),
                  "onr_msssim")
            vals.append(out1.view(1))
        else:
            vals.append(torch.zeros(1, device=output.device))
    msssim = torch.cat(vals, dim=0)
    return msssim.view(1, -1).expand(output_list[-1].size(0), -1)


def lr_multiplier(cur_epoch, cur_iter, data_size, args):
    """The schedule of reference utils.py:240-254 (warm-up 0.1 -> 1, then cosine / step / const)."""
    cur_epoch = cur_epoch + (float(cur_iter) / data_size)
    if args.lr_type == 'cosine':
        lr_mult = 0.5 * (math.cos(math.pi * (cur_epoch - args.warmup) / (args.epochs - args.warmup)) + 1.0)
    elif args.lr_type == 'step':
        lr_mult = 0.1 ** (sum(cur_epoch >= np.array(args.lr_steps)))
    elif args.lr_type in ('const', 'plateau'):
        lr_mult = 1
    else:
        raise NotImplementedError
    if cur_epoch < args.warmup:
        lr_mult = 0.1 + 0.9 * cur_epoch / args.warmup
    return lr_mult


def adjust_lr(optimizer, cur_epoch, cur_iter, data_size, args):
    """Reference utils.py:240-259."""
    lr = args.lr * lr_multiplier(cur_epoch, cur_iter, data_size, args)
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr
    return lr


def quantize_per_tensor(t, bit=8, axis=-1):
    """Reference utils.py:11-67.  Returns (quant_t, new_t) on t's device, same shape and dtype float32."""
    lib = _lib.lib()
    src_device = t.device
    x = _cuda_f32(t)
    if axis == -1 or x.dim() < 2:
        rows, view = 1, x.reshape(1, -1)
        restore = lambda y: y.reshape(x.shape)
    elif axis == 0:
        rows, view = x.size(0), x.reshape(x.size(0), -1)
        restore = lambda y: y.reshape(x.shape)
    elif axis == 1:
        xt = x.transpose(0, 1).contiguous()
        rows, view = xt.size(0), xt.reshape(xt.size(0), -1)
        restore = lambda y: y.reshape(xt.shape).transpose(0, 1).contiguous()
    else:
        raise NotImplementedError(f"quant axis {axis}")
    q = torch.empty_like(view)
    new = torch.empty_like(view)
    out_q, out_new = [], []
    # the kernel grid carries one row per blockIdx.y (<= 65535)
    for r0 in range(0, rows, 65535):
        r1 = min(rows, r0 + 65535)
        scratch = torch.empty((r1 - r0) * 2, dtype=torch.int32, device=x.device)
        check(lib.onr_quant_rows(ptr(view[r0:r1]), r1 - r0, view.size(1), bit, ptr(q[r0:r1]), ptr(new[r0:r1]),
                                 ptr(scratch), _lib.stream()), "onr_quant_rows")
    return restore(q).to(src_device), restore(new).to(src_device)


def global_magnitude_threshold(tensors, amount):
    """k-th smallest |w| over `tensors` with k = round(amount * N) — the threshold
    torch.nn.utils.prune.global_unstructured(L1Unstructured) uses (reference main_eval.py:587) — by an
    exact 4-pass radix select on the device.  Returns (threshold float or None when k == 0, k)."""
    lib = _lib.lib()
    ts = [_cuda_f32(t).reshape(-1) for t in tensors]
    n = sum(t.numel() for t in ts)
    k = int(round(amount * n)) if isinstance(amount, float) else int(amount)
    if k <= 0:
        return None, 0
    hist = torch.zeros(256, dtype=torch.int64, device=ts[0].device)
    prefix, mask, remaining = 0, 0, k
    for shift in (24, 16, 8, 0):
        hist.zero_()
        for t in ts:
            check(lib.onr_abs_radix_hist(ptr(t), t.numel(), prefix, mask, shift, ptr(hist), _lib.stream()),
                  "onr_abs_radix_hist")
        h = hist.cpu().tolist()
        acc = 0
        for b in range(256):
            if acc + h[b] >= remaining:
                prefix |= b << shift
                mask |= 0xFF << shift
                remaining -= acc
                break
            acc += h[b]
    thr = torch.tensor([prefix], dtype=torch.int32).view(torch.float32).item()
    return thr, k


def RoundTensor(x, num=2, group_str=False):
    """Reference utils.py:213-238."""
    if group_str:
        return '/'.join(','.join(str(round(e, num)) for e in x[i].tolist()) for i in range(x.size(0)))
    return ','.join(str(round(e, num)) for e in x.flatten().tolist())
