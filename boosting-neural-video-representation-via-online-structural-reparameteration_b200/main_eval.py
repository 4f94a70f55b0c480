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
`--finetune` with `--prune_ratio < 1` (prune-then-finetune, main_eval.py:213-545) replaces steps 1-2:
  1'. build the TRAIN-state Generator, load `model_latest.pth` (:213-236);
  2'. global magnitude pruning over the train-state tensors — stem Linear weights + `branch` (vanilla) or the six ERB
      branch convolutions of every block (:239-368) — again with the device radix select;
  3'. fresh Adam (:426, the checkpoint's optimizer state is skipped :429-437), `--finetune_epochs` epochs of the training
      step with epoch numbering continued at the checkpoint's epoch (:446-507) on `trainer.FrameFitter`: the pruning
      re-parameterisation `weight = weight_orig * weight_mask` is carried by keeping the weights masked and multiplying
      the flat gradient buffer by the masks before every Adam step (one launch);
  4'. ERB: `switch_to_deploy` on every block (:530-541); then steps 3-4 above on the fine-tuned model.
Reference quirk kept by default: an ERB block never CALLS its branch convolutions, so prune's
forward-pre-hook never refreshes their `.weight` — the fold keeps reading the tensor computed at prune time while Adam
moves the unused `weight_orig`.  Fine-tuning therefore trains only the branch biases, the stem (whose Linear modules
are called) and the head; the pruned branch kernels stay frozen.  `ONR_FINETUNE_BRANCHES=1` fixes that instead: the
branch kernels train under their masks (mask-aware fold).
"""
import os
import time

import torch
import torch.nn.utils.prune as prune

from .cli_common import (FrameCache, build_model, build_parser, finish_args, huffman_avg_bits, prepare_outdir,
                         strip_profiler_keys)
from .model import NeRVBlock, _ERB_BRANCHES
from .utils import RoundTensor, frame_stats, global_magnitude_threshold, msssim_fn, quantize_per_tensor


def prunable_modules(model):
    mods = [m for m in (model.stem[0], model.stem[2])]
    for layer in model.layers:
        if isinstance(layer, NeRVBlock):
            mods.append(layer.single_conv())
    return mods


def global_masks(ws, amount):
    """Masks of prune.global_unstructured(L1Unstructured, amount) over the tensors `ws`, with the k-th value found on
    the device.  Exactly k = round(amount * N) entries are masked, as `torch.topk` does in the reference
    (main_eval.py:587): everything strictly below the k-th magnitude, plus as many of the entries EQUAL to it (ties:
    exact zeros, repeated quantised values) as are needed to reach k, taken in flat-index order."""
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
    return masks


def global_prune(model, amount):
    """Deploy / vanilla model: global masks registered through torch.nn.utils.prune (main_eval.py:572-587, :616-641)."""
    mods = prunable_modules(model)
    masks = global_masks([m.weight.detach() for m in mods], amount)
    for m, mask in zip(mods, masks):
        prune.custom_from_mask(m, 'weight', mask)
    total = sum(m.weight_mask.numel() for m in mods)
    zeros = sum(int((m.weight_mask == 0).sum()) for m in mods)
    return zeros, total


def train_state_prunable(model):
    """(parameter-name prefix, module) of every tensor the finetune path prunes, in the reference's order
    (main_eval.py:239-352): stem Linear layers, then per block `branch` (vanilla; `rbr_reparam` if already deployed)
    or the six ERB branch convolutions."""
    out = [('stem.0', model.stem[0]), ('stem.2', model.stem[2])]
    for l, layer in enumerate(model.layers):
        if hasattr(layer, 'branch'):
            out.append((f'layers.{l}.branch', layer.branch))
        elif hasattr(layer, 'rbr_3x3_branch'):
            # reference order :320-350: 3x3, 3x1, 1x3, 1x1_1, 3x3 (sequence), 1x1_2 == the creation order
            out += [(f'layers.{l}.{n}', getattr(layer, n)) for n in _ERB_BRANCHES if hasattr(layer, n)]
        elif hasattr(layer, 'rbr_reparam'):
            out.append((f'layers.{l}.rbr_reparam', layer.rbr_reparam))
    return out


def prune_finetune(model, pe, cache, args, start_epoch, log_path=None, local_rank=0):
    """Steps 2'-4' of the module docstring on a loaded train-state model.  Returns the info text; on return the model
    is in the state the reference reaches at main_eval.py:545 — ERB blocks deployed, the still-pruned modules (stem;
    vanilla `branch`) carrying prune's weight_orig / weight_mask with the ORIGINAL values under the masked entries
    (they never receive a gradient), which is what the quantisation step then reads."""
    from datetime import datetime
    from . import sharding
    from .main_train import fit_epoch
    from .optim import FusedAdam
    from .trainer import FrameFitter
    if args.branch_type not in ('NeRV_vanilla', 'ERB'):
        raise NotImplementedError(f'prune-then-finetune handles NeRV_vanilla and ERB (reference main_eval.py:238, '
                                  f':297), not {args.branch_type}')
    info = ''
    targets = train_state_prunable(model)
    for name, _ in targets:
        info += f'prune list += {name}.weight\n'
    ws = [m.weight.detach() for _, m in targets]
    masks = global_masks(ws, args.prune_ratio)
    total = sum(m.numel() for m in masks)
    zeros = sum(int((m == 0).sum()) for m in masks)
    msg = (f'global prune (train state): target {args.prune_ratio}, actual {zeros / max(total, 1):.3f} '
           f'({zeros}/{total} mask zeros)')
    print(msg)
    info += msg + '\n'
    originals = [w.clone() for w in ws]
    train_branches = os.environ.get('ONR_FINETUNE_BRANCHES', '0') == '1'
    grad_masks = {}
    with torch.no_grad():
        for (name, m), mask in zip(targets, masks):
            m.weight.mul_(mask)
            frozen = ('.rbr_' in name and 'rbr_reparam' not in name) and not train_branches
            grad_masks[name + '.weight'] = torch.zeros_like(mask) if frozen else mask
    model._weights_epoch = getattr(model, '_weights_epoch', 0) + 1
    if args.branch_type == 'ERB':
        info += ('ERB branch kernels train under their masks (ONR_FINETUNE_BRANCHES=1)\n' if train_branches else
                 'ERB branch kernels stay frozen at their pruned values (reference behaviour, main_eval.py:476-480)\n')

    optimizer = FusedAdam(model.parameters(), betas=(args.beta, 0.999))
    data_size = len(cache)
    spe = sharding.steps_per_epoch(data_size, args.batchSize)
    total_epochs = start_epoch + args.finetune_epochs
    fitter = FrameFitter(model, pe, args, optimizer=optimizer, steps_per_epoch=spe, data_size=data_size,
                         grad_masks=grad_masks, epoch_offset=start_epoch, epoch_mod=total_epochs)
    best_psnr, best_msssim = 0.0, 0.0
    start = datetime.now()
    model.train()
    for epoch in range(start_epoch, total_epochs):
        t0 = datetime.now()
        st = fit_epoch(fitter, cache, args, epoch, total_epochs, spe, 1, 0, local_rank, log_path)
        best_psnr, best_msssim = max(best_psnr, float(st[4])), max(best_msssim, float(st[5]))
        now = datetime.now()
        line = '\t{}p: current: {:.2f}\t best: {:.2f}\t msssim_best: {:.4f}\t'.format(
            fitter.H, float(st[4]), best_psnr, best_msssim)
        line += 'Time/epoch: \tCurrent:{:.2f} \tAverage:{:.2f}'.format(
            (now - t0).total_seconds(), (now - start).total_seconds() / (epoch + 1 - start_epoch))
        print(line, flush=True)
        if log_path:
            with open(log_path, 'a') as f:
                f.write(line + '\n')
    fitter.release_graph()
    for p_ in model.parameters():
        p_.grad = None
    del fitter
    model._executors = {}
    if args.branch_type == 'ERB':
        n = 0
        for layer in model.layers:
            if isinstance(layer, NeRVBlock):
                layer.switch_to_deploy()
                n += 1
        line = f'finetune done: {n} NeRVBlocks switched to deploy state'
        print(line)
        info += line + '\n'
    # hand the still-pruned modules back in torch.nn.utils.prune's layout: weight_orig (fine-tuned where the mask is 1,
    # the original value where it is 0) + weight_mask
    with torch.no_grad():
        for (name, m), mask, w0 in zip(targets, masks, originals):
            if '.rbr_' in name and not hasattr(model.layers[int(name.split('.')[1])], name.split('.')[2]):
                continue                                   # an ERB branch that switch_to_deploy has folded away
            m.weight.add_((1 - mask) * w0)
            prune.custom_from_mask(m, 'weight', mask)
    model._weights_epoch = getattr(model, '_weights_epoch', 0) + 1
    return info


def prune_and_quantise(model, args, n_frames, frame_hw, log_bpp=None, prune_now=True):
    """Steps 2-3 of the reference flow on a loaded deploy / vanilla model: global magnitude pruning
    (main_eval.py:572-587) then quantisation of every state-dict tensor + Huffman / bpp statistics (:652-729)
    and `load_state_dict` (:703).  Returns the info text the reference writes to its `only_prune*` file."""
    info = ''
    if args.prune_ratio < 1 and prune_now:
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
            psnrs.append(frame_stats(out[-1], target)[4].view(1))   # the full-resolution stage
            msssims.append(msssim_fn(out[-1:], [target]).view(1))
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
    local_rank = 0
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    prepare_outdir(args, 0)
    torch.manual_seed(args.manualSeed)
    info = ''

    erb = args.branch_type == 'ERB'
    finetune = bool(args.finetune) and args.prune_ratio < 1          # reference main_eval.py:213-214
    # reference :163-180: ERB without finetune evaluates the deploy-state model, everything else the train state
    pe, model = build_model(args, device, deploy=False if (finetune or args.finetune) else (True if erb else args.deploy))
    ckpt_name = 'model_latest_deploy.pth' if (erb and not args.finetune) else 'model_latest.pth'
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
    if finetune:
        ft_name = 'finetune_e{}_pr{:.2f}_q{}.txt'.format(args.finetune_epochs, args.prune_ratio,
                                                        args.quant_bit if args.quant_bit != -1 else 'none')
        ft_log = '{}/{}'.format(args.outf, ft_name)
        train_cache = (cache if args.frame_gap == args.test_gap else
                       FrameCache(args.dataset, device, vid_list=args.vid, frame_gap=args.frame_gap))
        start_epoch = int(checkpoint['epoch']) if isinstance(checkpoint, dict) and 'epoch' in checkpoint else 0
        ft_info = prune_finetune(model, pe, train_cache, args, start_epoch, log_path=ft_log, local_rank=local_rank)
        with open(ft_log, 'a') as f:
            f.write(ft_info)
        info += ft_info
    info += prune_and_quantise(model, args, len(cache), tuple(cache.frames.shape[-2:]),
                               log_bpp='{}/bpp_rank{}.txt'.format(args.outf, local_rank)
                               if args.quant_bit != -1 else None, prune_now=not finetune)

    only_name = 'only_prune{:.2f}_quant{}.txt'.format(args.prune_ratio, args.quant_bit if args.quant_bit > 0 else 'full')
    log_path = '{}/{}'.format(args.outf, only_name)
    with open(log_path, 'w', encoding='utf-8') as f:
        f.write(info)
    res = decode_clip(model, pe, cache, args, log_path=log_path, local_rank=local_rank,
                      fwd_num=getattr(args, 'fwd_num', 10))
    return res['psnr'], res['msssim']


if __name__ == '__main__':
    main()
