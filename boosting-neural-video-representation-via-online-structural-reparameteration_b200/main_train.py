"""Training driver with the reference's CLI (main_train.py) on top of the B200 hot path.

    python -m orepnerv.main_train -e 300 --lr 0.0005 -b 1 --embed 1.25_40 --stem_dim_num 512_1 \
        --fc_hw_dim 9_16_26 --expansion 1 --reduction 2 --lower_width 96 --strides 5 2 2 2 2 --single_res \
        --loss Fusion6 --warmup 0.2 --lr_type cosine --norm none --act swish --branch_type ERB \
        --dataset synthetic:132x720x1280 --outf bunny --suffix erb
    torchrun --nproc-per-node 8 -m orepnerv.main_train ... -d          (frame-sharded data parallel)

Kept from the reference: flags and their prefix matching (`--loss` -> `--loss_type`), output directory naming,
log line formats, per-epoch checkpoint files and their dict layout (main_train.py:292-358): `model_latest.pth`,
`model_latest_deploy.pth` (folded), `*_train_best*.pth`, `model_val_best.pth`.
Replaced: the per-step body (main_train.py:229-254) is `FrameFitter.step` — sm_100a kernels on a frame cache
in HBM, replayed as one CUDA graph; with -d the gradients are all-reduced over NCCL.
"""
import os
import random
import time
from copy import deepcopy
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist

from . import sharding
from .cli_common import FrameCache, build_model, build_parser, finish_args, prepare_outdir
from .model import NeRVBlock
from .optim import FusedAdam
from .trainer import FrameFitter
from .utils import RoundTensor, frame_stats, msssim_fn


def main(argv=None):
    args = finish_args(build_parser().parse_args(argv))
    world = int(os.environ.get("WORLD_SIZE", "1")) if args.distributed else 1
    rank = int(os.environ.get("RANK", "0")) if args.distributed else 0
    local_rank = int(os.environ.get("LOCAL_RANK", "0")) if args.distributed else 0
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prepare_outdir(args, rank)
    train(local_rank, rank, world, args)
    if world > 1:
        dist.destroy_process_group()


@torch.no_grad()
def evaluate(model, cache, pe, args, local_rank=0, log_path=None):
    """Reference main_train.py:377-438: decode every `test_gap`-th frame with the current (train-state) model,
    `--eval_fps` repeating each forward 10 times for the FPS figure (:397), per-stage PSNR / MS-SSIM against the frame
    pooled to each stage's resolution (:421-427), the progress line of :428-436.  The MACs line of :407-417 (thop) is
    replaced by the analytic count of the convolutions.  Returns (psnr[num_stage], msssim[num_stage],
    decode frames/s)."""
    from .utils import adaptive_avg_pool2d, psnr_fn
    psnr_list, msssim_list, time_list = [], [], []
    model.eval()
    frames = list(range(0, len(cache), args.test_gap))
    fwd_num = 10 if getattr(args, 'eval_fps', False) else 1
    val_psnr = val_msssim = None
    for n, i in enumerate(frames):
        if n > 10 and args.debug:
            break
        embed_input = pe(cache.t[i:i + 1])
        data = cache.frames[i:i + 1].float().div(255)
        for _ in range(fwd_num):
            torch.cuda.synchronize()
            t0 = time.time()
            output_list = model(embed_input)
            torch.cuda.synchronize()
            time_list.append(time.time() - t0)
        if n == 0:
            print(f"MACs: {decoder_macs(model) / 10 ** 9:.2f}G")
        target_list = [adaptive_avg_pool2d(data, x.shape[-2:]) for x in output_list]
        psnr_list.append(psnr_fn(output_list, target_list))
        msssim_list.append(msssim_fn(output_list, target_list))
        val_psnr = torch.cat(psnr_list, dim=0).mean(0)
        val_msssim = torch.cat(msssim_list, dim=0).float().mean(0)
        if n % args.print_freq == 0 or n == len(frames) - 1:
            fps = fwd_num * (n + 1) / sum(time_list)
            print_str = 'Rank:{}, Step [{}/{}], PSNR: {}, MSSSIM: {} FPS: {}'.format(
                local_rank, n + 1, len(frames), RoundTensor(val_psnr, 2, False), RoundTensor(val_msssim, 4, False),
                round(fps, 2))
            print(print_str)
            if log_path:
                with open(log_path, 'a') as f:
                    f.write(print_str + '\n')
    model.train()
    return val_psnr.cpu(), val_msssim.cpu(), len(time_list) / max(sum(time_list), 1e-9)


def decoder_macs(model):
    """Multiply-accumulates of one decoded frame (stem + block convolutions + heads): what `thop.profile` reports at
    reference main_train.py:407-417, counted from the module shapes."""
    macs = sum(m.in_features * m.out_features for m in model.stem if hasattr(m, 'in_features'))
    h, w = model.fc_h, model.fc_w
    for blk, head in zip(model.layers, model.head_layers):
        macs += h * w * blk.out_channels * blk.ngf * 9
        h, w = h * blk.stride, w * blk.stride
        if head is not None:
            macs += h * w * head.in_channels * 3
    return macs


def fit_epoch(fitter, cache, args, epoch, total_epochs, spe, world, rank, local_rank, log_path=None):
    """One epoch of the reference step loop (main_train.py:229-267; main_eval.py:450-507 runs the same loop for
    prune-then-finetune): frame order, `FrameFitter.step` per iteration, the reference's progress line every
    `print_freq` steps.  Returns the epoch means [loss, L1, SSIM, MSE, PSNR, MS-SSIM] as a CPU tensor."""
    device, B, data_size = fitter.dev, args.batchSize, len(cache)
    order = sharding.shard_indices(data_size, world, rank, epoch, seed=args.manualSeed) if B == 1 else None
    stats = []
    n_steps = spe if not args.debug else min(spe, 11)
    for i in range(n_steps):
        if B == 1:
            idx = torch.tensor([order[i]], device=device)
        else:
            perm = sharding.epoch_permutation(data_size, epoch, args.manualSeed)
            sl = [perm[(i * world * B + rank * B + k) % data_size] for k in range(B)]
            idx = torch.tensor(sl, device=device)
        out = fitter.step(cache.frames[idx], cache.t[idx])
        stats.append(out[:6].clone())
        if i % args.print_freq == 0 or i == n_steps - 1:
            st = torch.stack(stats).mean(0).cpu()
            lr = fitter.opt.param_groups[0]['lr']
            print_str = '[{}] Rank:{}, Epoch[{}/{}], Step [{}/{}], lr:{:.2e} PSNR: {}, MSSSIM: {}'.format(
                datetime.now().strftime("%Y/%m/%d %H:%M:%S"), local_rank, epoch + 1, total_epochs, i + 1, n_steps,
                lr, RoundTensor(st[4:5], 2, False), RoundTensor(st[5:6], 4, False))
            print(print_str, flush=True)
            if rank == 0 and log_path:
                with open(log_path, 'a') as f:
                    f.write(print_str + '\n')
    return torch.stack(stats).mean(0).cpu()


def fit_epoch_modules(model, pe, optimizer, cache, args, epoch, total_epochs, rank, local_rank, log_path=None):
    """One epoch of the reference loop written against the drop-in MODULES (main_train.py:229-267 verbatim:
    `model(embed)`, per-stage targets by adaptive average pooling, `loss_fn` per stage weighted by --lw,
    `loss_sum.backward()`, `optimizer.step()`, `psnr_fn`, `msssim_fn`) — the path multi-resolution heads
    (sin_res=False) train through; the single-resolution configurations use `FrameFitter` (fit_epoch) instead.
    Single process.  Returns the epoch means (psnr[num_stage], msssim[num_stage]) as CPU tensors."""
    from .utils import adaptive_avg_pool2d, adjust_lr, loss_fn, psnr_fn
    device = next(model.parameters()).device
    B, data_size = args.batchSize, len(cache)
    perm = sharding.epoch_permutation(data_size, epoch, args.manualSeed)
    n_steps = data_size // B if not args.debug else min(data_size // B, 11)
    psnr_list, msssim_list = [], []
    for i in range(n_steps):
        idx = torch.tensor(perm[i * B:(i + 1) * B], device=device)
        data = cache.frames[idx].float().div(255)
        embed_input = pe(cache.t[idx])
        output_list = model(embed_input)
        target_list = [adaptive_avg_pool2d(data, x.shape[-2:]) for x in output_list]
        loss_list = [loss_fn(output, target, args) for output, target in zip(output_list, target_list)]
        loss_list = [loss_list[k] * (args.lw if k < len(loss_list) - 1 else 1) for k in range(len(loss_list))]
        loss_sum = sum(loss_list)
        lr = adjust_lr(optimizer, epoch % total_epochs, i, data_size, args)
        optimizer.zero_grad()
        loss_sum.backward()
        optimizer.step()
        with torch.no_grad():
            psnr_list.append(psnr_fn(output_list, target_list))
            msssim_list.append(msssim_fn(output_list, target_list))
        if i % args.print_freq == 0 or i == n_steps - 1:
            train_psnr = torch.cat(psnr_list, dim=0).mean(0).cpu()
            train_msssim = torch.cat(msssim_list, dim=0).float().mean(0).cpu()
            print_str = '[{}] Rank:{}, Epoch[{}/{}], Step [{}/{}], lr:{:.2e} PSNR: {}, MSSSIM: {}'.format(
                datetime.now().strftime("%Y/%m/%d %H:%M:%S"), local_rank, epoch + 1, total_epochs, i + 1, n_steps,
                lr, RoundTensor(train_psnr, 2, False), RoundTensor(train_msssim, 4, False))
            print(print_str, flush=True)
            if rank == 0 and log_path:
                with open(log_path, 'a') as f:
                    f.write(print_str + '\n')
    return torch.cat(psnr_list, dim=0).mean(0).cpu(), torch.cat(msssim_list, dim=0).float().mean(0).cpu()


def train(local_rank, rank, world, args):
    torch.manual_seed(args.manualSeed)
    np.random.seed(args.manualSeed)
    random.seed(args.manualSeed)
    device = torch.device('cuda', local_rank)
    train_best_psnr, train_best_msssim, val_best_psnr, val_best_msssim = [torch.tensor(0.0) for _ in range(4)]
    is_train_best = False

    pe, model = build_model(args, device)
    total_params = sum(p.numel() for p in model.parameters()) / 1e6
    log_path = '{}/rank{}.txt'.format(args.outf, rank)
    if rank == 0:
        print(f'{args}\n {model}\n Model Params: {total_params}M')
        with open(log_path, 'a') as f:
            f.write(str(model) + '\n' + f'Params: {total_params}M\n')
    try:
        from torch.utils.tensorboard import SummaryWriter
        writer = SummaryWriter(os.path.join(args.outf, f'param_{total_params}M', 'tensorboard')) if rank == 0 else None
    except Exception:
        writer = None
    print("Use GPU: {} for training".format(local_rank))

    optimizer = FusedAdam(model.parameters(), betas=(args.beta, 0.999))
    cache = FrameCache(args.dataset, device, vid_list=args.vid, frame_gap=args.frame_gap)
    data_size = len(cache)                                           # reference: len(train_dataset)
    spe = sharding.steps_per_epoch(data_size, world * args.batchSize)
    multi = not args.single_res
    if multi and world > 1:
        raise SystemExit('orepnerv: multi-resolution heads train through the module API in a single process (drop -d)')
    fitter = None if multi else FrameFitter(model, pe, args, optimizer=optimizer, world_size=world,
                                            steps_per_epoch=spe, data_size=data_size)
    H_out, W_out = (int(x) for x in cache.frames.shape[-2:])
    start = datetime.now()
    for epoch in range(args.epochs):
        epoch_start = datetime.now()
        if multi:
            train_psnr, train_msssim = fit_epoch_modules(model, pe, optimizer, cache, args, epoch, args.epochs, rank,
                                                         local_rank, log_path)
        else:
            st = fit_epoch(fitter, cache, args, epoch, args.epochs, spe, world, rank, local_rank, log_path)
            train_psnr, train_msssim = st[4:5], st[5:6]
        if rank == 0:
            h, w = H_out, W_out
            is_train_best = bool(train_psnr[-1] > train_best_psnr)
            train_best_psnr = train_psnr[-1] if is_train_best else train_best_psnr
            train_best_msssim = train_msssim[-1] if train_msssim[-1] > train_best_msssim else train_best_msssim
            if writer is not None:
                writer.add_scalar(f'Train/PSNR_{h}X{w}_gap{args.frame_gap}', train_psnr[-1].item(), epoch + 1)
                writer.add_scalar(f'Train/MSSSIM_{h}X{w}_gap{args.frame_gap}', train_msssim[-1].item(), epoch + 1)
                writer.add_scalar('Train/lr', optimizer.param_groups[0]['lr'], epoch + 1)
            now = datetime.now()
            print_str = '\t{}p: current: {:.2f}\t best: {:.2f}\t msssim_best: {:.4f}\t'.format(
                h, train_psnr[-1].item(), float(train_best_psnr), float(train_best_msssim))
            print_str += "Time/epoch: \tCurrent:{:.2f} \tAverage:{:.2f}".format(
                (now - epoch_start).total_seconds(), (now - start).total_seconds() / (epoch + 1))
            print(print_str, flush=True)
            with open(log_path, 'a') as f:
                f.write(print_str + '\n')

        save_checkpoint = {
            'epoch': epoch + 1, 'state_dict': model.state_dict(), 'train_best_psnr': train_best_psnr,
            'train_best_msssim': train_best_msssim, 'val_best_psnr': val_best_psnr,
            'val_best_msssim': val_best_msssim, 'optimizer': optimizer.state_dict()}

        if (epoch + 1) % args.eval_freq == 0 or epoch > args.epochs - 10:
            val_psnr, val_msssim, fps = evaluate(model, cache, pe, args, local_rank, log_path if rank == 0 else None)
            if rank == 0:
                is_val_best = bool(val_psnr[-1] > val_best_psnr)
                val_best_psnr = val_psnr[-1] if is_val_best else val_best_psnr
                val_best_msssim = val_msssim[-1] if val_msssim[-1] > val_best_msssim else val_best_msssim
                print_str = f'Eval best_PSNR at epoch{epoch + 1}:'
                print_str += '\t{}p: current: {:.2f}\tbest: {:.2f} \tbest_msssim: {:.4f}\t decode fps: {:.1f}'.format(
                    H_out, val_psnr[-1].item(), float(val_best_psnr), float(val_best_msssim), fps)
                print(print_str)
                with open(log_path, 'a') as f:
                    f.write(print_str + '\n')
                if is_val_best:
                    torch.save(save_checkpoint, '{}/model_val_best.pth'.format(args.outf))

        if rank == 0:
            torch.save(save_checkpoint, '{}/model_latest.pth'.format(args.outf))
            if is_train_best:
                torch.save(save_checkpoint, '{}/model_train_best.pth'.format(args.outf))
            if args.branch_type == 'ERB':
                copy_model = deepcopy(model)
                for layer in copy_model.layers:
                    if isinstance(layer, NeRVBlock):
                        layer.switch_to_deploy()
                deploy_checkpoint = dict(save_checkpoint, state_dict=copy_model.state_dict())
                torch.save(deploy_checkpoint, '{}/model_latest_deploy.pth'.format(args.outf))
                if is_train_best:
                    torch.save(deploy_checkpoint, '{}/model_train_best_deploy.pth'.format(args.outf))
                if epoch == args.epochs - 1:
                    n_dep = sum(p.numel() for p in copy_model.parameters()) / 1e6
                    with open(log_path, 'a') as f:
                        f.write(f'Deploy Rep-Model Params: {n_dep:.3f}M\n')
    if rank == 0:
        print("Training complete in: " + str(datetime.now() - start))


if __name__ == '__main__':
    main()
