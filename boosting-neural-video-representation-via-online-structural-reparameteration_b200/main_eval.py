"""Eval / decode driver with the reference's CLI (main_eval.py, no-finetune path :551-827) on the B200 path.

    python -m orepnerv.main_eval <same model flags as training> --branch_type ERB --outf bunny --suffix erb \
        --prune_ratio 0.2 --quant_bit 8 --eval_only [--dump_images]

Flow (reference line numbers):
  1. build the deploy-state Generator for ERB (:163-170), load `model_latest_deploy.pth` (ERB) or
     `model_latest.pth` (vanilla) after stripping thop keys, strict=False (:554-567, :599-611);
  2. global magnitude pruning of stem Linear weights + block conv weights (:572-587, :616-641) — the threshold is
     found by the device radix select, the masks are registered through torch.nn.utils.prune so the state dict has the
     reference's weight_orig / weight_mask entries;
  3. `--quant_bit`: quantise EVERY state-dict tensor (:660-669), Huffman-length statistics and bpp (:673-729), load back
     (:703).  Reference behaviour kept on purpose: weight_mask is quantised too, which turns the mask
     into all ones, and "8 bit" has 257 levels;
  4. decode loop (:738-827): `fwd_num` timed forwards per frame, 5 + 50 forwards on the first frame for FPS, PSNR /
     MS-SSIM accumulation, optional PNG dump.
`--finetune` (prune-then-finetune, main_eval.py:213-545) is outside the hot path and raises NotImplementedError.
"""
import os
import time

import torch
import torch.nn.utils.prune as prune

from .cli_common import (FrameCache, build_model, build_parser, finish_args, huffman_avg_bits, prepare_outdir,
                         strip_profiler_keys)
from .model import NeRVBlock
from .utils import RoundTensor, frame_stats, global_magnitude_threshold, msssim_fn, quantize_per_tensor


def prunable_modules(model):
    mods = [m for m in (model.stem[0], model.stem[2])]
    for layer in model.layers:
        if isinstance(layer, NeRVBlock):
            mods.append(layer.single_conv())
    return mods


def global_prune(model, amount):
    """prune.global_unstructured(L1Unstructured, amount) with the k-th value found on the device.  Exactly
    k = round(amount * N) entries are masked, as `torch.topk` does in the reference (main_eval.py:587): everything
    strictly below the k-th magnitude, plus as many of the entries EQUAL to it (ties: exact zeros, repeated quantised
    values) as are needed to reach k, taken in flat-index order."""
    mods = prunable_modules(model)
    ws = [m.weight.detach() for m in mods]
    thr, k = global_magnitude_threshold(ws, float(amount))
    if thr is None:
        masks = [torch.ones_like(w) for w in ws]
    else:
        flat = torch.cat([w.abs().reshape(-1) for w in ws])
        below = flat < thr
        ties = flat == thr
        need = k - int(below.sum())
        pruned = below | (ties & (torch.cumsum(ties, 0) <= need))
        keep = (~pruned).to(ws[0].dtype)
        masks, off = [], 0
        for w in ws:
            masks.append(keep[off:off + w.numel()].view_as(w))
            off += w.numel()
    for m, mask in zip(mods, masks):
        prune.custom_from_mask(m, 'weight', mask)
    total = sum(m.weight_mask.numel() for m in mods)
    zeros = sum(int((m.weight_mask == 0).sum()) for m in mods)
    return zeros, total


def prune_and_quantise(model, args, n_frames, frame_hw, log_bpp=None):
    """Steps 2-3 of the reference flow on a loaded deploy / vanilla model: global magnitude pruning
    (main_eval.py:572-587) then quantisation of every state-dict tensor + Huffman / bpp statistics (:652-729)
    and `load_state_dict` (:703).  Returns the info text the reference writes to its `only_prune*` file."""
    info = ''
    if args.prune_ratio < 1:
        zeros, total = global_prune(model, args.prune_ratio)
        msg = f'global prune: target {args.prune_ratio}, actual {zeros / total:.3f} ({zeros}/{total} mask zeros)'
        print(msg)
        info += msg + '\n'
    if args.quant_bit != -1:
        with torch.no_grad():
            cur = model.state_dict()
            symbols = []
            for k, v in cur.items():
                large = v.dim() in {2, 4} and 'bias' not in k
                q, new_v = quantize_per_tensor(v, args.quant_bit, args.quant_axis if large else -1)
                symbols.append(q[v != 0].flatten())
                cur[k] = new_v.to(v.device).type_as(v)
            avg_bits, total_bits, n_sym = huffman_avg_bits(torch.cat(symbols))
            model.load_state_dict(cur)
            H, W = frame_hw
            bpp = total_bits / (n_frames * H * W)
            msg = (f'quantised {len(symbols)} tensors to {args.quant_bit} bit; Huffman {avg_bits:.4f} bit/symbol over '
                   f'{n_sym} symbols, efficiency {avg_bits / args.quant_bit:.4f}; total {int(total_bits)} bits, '
                   f'{n_frames} frames {H}x{W}, BPP={bpp:.6f}')
            print(msg)
            info += msg + '\n'
            if log_bpp:
                with open(log_bpp, 'a') as f:
                    f.write(msg + '\n')
    return info


def decode_clip(model, pe, cache, args, log_path=None, local_rank=0, fwd_num=10, quiet=False):
    """The reference decode loop (main_eval.py:738-827): per frame `fwd_num` timed forwards (host clock around
    `torch.cuda.synchronize()`, as the reference times it), on the first frame 5 + 50 extra forwards for the
    "first frame FPS", PSNR / MS-SSIM accumulation, optional PNG dump.
    Returns dict(psnr, msssim, fps, fps_first_frame, frames)."""
    psnrs, msssims, times = [], [], []
    model.eval()
    eval_str, fps0 = '', None
    with torch.no_grad():
        for i in range(len(cache)):
            embed = pe(cache.t[i:i + 1])
            target = cache.frames[i:i + 1].float().div(255)
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(fwd_num):
                out = model(embed)
            torch.cuda.synchronize()
            times.append(time.time() - t0)
            if i == 0:
                for _ in range(5):
                    model(embed)
                torch.cuda.synchronize()
                t0 = time.time()
                for _ in range(50):
                    model(embed)
                torch.cuda.synchronize()
                fps0 = 50 / (time.time() - t0)
                eval_str = f'[first frame] FPS: {fps0:.2f}\n'
                if not quiet:
                    print(eval_str.strip())
            if getattr(args, 'dump_images', False):
                from torchvision.utils import save_image
                vis = f'{args.outf}/visualize'
                os.makedirs(vis, exist_ok=True)
                save_image(out[-1][0], f'{vis}/pred_{i}.png')
            psnrs.append(frame_stats(out[0], target)[4].view(1))
            msssims.append(msssim_fn(out, [target]).view(1))
            if i % args.print_freq == 0 or i == len(cache) - 1:
                fps = fwd_num * (i + 1) / sum(times)
                print_str = 'Rank:{}, Step [{}/{}], PSNR: {}, MSSSIM: {} FPS: {}'.format(
                    local_rank, i + 1, len(cache), RoundTensor(torch.cat(psnrs).mean().view(1), 2, False),
                    RoundTensor(torch.cat(msssims).mean().view(1), 4, False), round(fps, 2))
                if not quiet:
                    print(print_str)
                if log_path:
                    with open(log_path, 'a') as f:
                        f.write(print_str + '\n' + eval_str + '\n')
    return dict(psnr=torch.cat(psnrs).mean().item(), msssim=torch.cat(msssims).mean().item(),
                fps=fwd_num * len(cache) / sum(times), fps_first_frame=fps0, frames=len(cache))


def main(argv=None):
    args = finish_args(build_parser(eval_mode=True).parse_args(argv))
    if args.finetune:
        raise NotImplementedError("prune-then-finetune (reference main_eval.py:213-545) is outside the B200 hot path")
    local_rank = 0
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    prepare_outdir(args, 0)
    torch.manual_seed(args.manualSeed)
    info = ''

    erb = args.branch_type == 'ERB'
    pe, model = build_model(args, device, deploy=True if erb else args.deploy)
    ckpt_name = 'model_latest_deploy.pth' if erb else 'model_latest.pth'
    path = os.path.join(args.outf, ckpt_name)
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    try:
        checkpoint = torch.load(path, map_location='cpu', weights_only=True)
    except Exception:
        checkpoint = torch.load(path, map_location='cpu', weights_only=False)
    state = checkpoint['state_dict'] if isinstance(checkpoint, dict) and 'state_dict' in checkpoint else checkpoint
    model.load_state_dict(strip_profiler_keys(state), strict=False)
    info += f'loaded {path}\n'

    cache = FrameCache(args.dataset, device, vid_list=args.vid, frame_gap=args.test_gap)
    info += prune_and_quantise(model, args, len(cache), tuple(cache.frames.shape[-2:]),
                               log_bpp='{}/bpp_rank{}.txt'.format(args.outf, local_rank)
                               if args.quant_bit != -1 else None)

    only_name = 'only_prune{:.2f}_quant{}.txt'.format(args.prune_ratio, args.quant_bit if args.quant_bit > 0 else 'full')
    log_path = '{}/{}'.format(args.outf, only_name)
    with open(log_path, 'w', encoding='utf-8') as f:
        f.write(info)
    res = decode_clip(model, pe, cache, args, log_path=log_path, local_rank=local_rank,
                      fwd_num=getattr(args, 'fwd_num', 10))
    return res['psnr'], res['msssim']


if __name__ == '__main__':
    main()
