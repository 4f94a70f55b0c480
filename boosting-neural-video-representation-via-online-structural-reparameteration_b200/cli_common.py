"""Pieces shared by main_train.py and main_eval.py: the reference's argparse surface (main_train.py:39-109,
main_eval.py:31-104), output-directory naming (main_train.py:111-147), the frame source and the model
factory.  Host-side Python only; nothing here computes on tensors except loading frames.
"""
import argparse
import os
import re
import shutil

import torch

from .data import synthetic_clip
from .model import Generator
from .utils import PositionalEncoding


def build_parser(eval_mode=False):
    p = argparse.ArgumentParser(fromfile_prefix_chars="@", epilog=SUPPORTED)
    # dataset parameters
    p.add_argument('--vid', default=[None], type=int, nargs='+', help='video id list for training')
    p.add_argument('--scale', type=int, default=1)
    p.add_argument('--frame_gap', type=int, default=1, help='frame selection gap')
    p.add_argument('--augment', type=int, default=0)
    p.add_argument('--dataset', type=str, default='UVG',
                   help="directory name under ../data (PNG/JPG frames, as in the reference) or "
                        "'synthetic:<frames>x<H>x<W>' for the built-in synthetic clip")
    p.add_argument('--test_gap', default=1, type=int, help='evaluation gap')
    # architecture
    p.add_argument('--embed', type=str, default='1.25_80')
    p.add_argument('--stem_dim_num', type=str, default='1024_1')
    p.add_argument('--fc_hw_dim', type=str, default='9_16_128')
    p.add_argument('--expansion', type=float, default=8)
    p.add_argument('--reduction', type=int, default=2)
    p.add_argument('--strides', type=int, nargs='+', default=[5, 3, 2, 2, 2])
    p.add_argument('--num_blocks', type=int, default=1)
    p.add_argument('--norm', default='none', type=str, choices=['none', 'bn', 'in'])
    p.add_argument('--act', type=str, default='gelu',
                   choices=['relu', 'leaky', 'leaky01', 'relu6', 'gelu', 'swish', 'softplus', 'hardswish'])
    p.add_argument('--lower_width', type=int, default=32)
    p.add_argument("--single_res", action='store_true')
    p.add_argument("--conv_type", default='conv', type=str, choices=['conv', 'deconv', 'bilinear'])
    p.add_argument("--branch_type", default='NeRV_vanilla', type=str,
                   choices=['NeRV_vanilla', 'ERB', 'ACB', 'RepVGG', 'DBB', 'ECB'])
    # training
    p.add_argument('-j', '--workers', type=int, default=4)
    p.add_argument('-b', '--batchSize', type=int, default=1)
    p.add_argument('--not_resume_epoch', action='store_true')
    p.add_argument('-e', '--epochs', type=int, default=150)
    if eval_mode:
        p.add_argument('--cycles', type=int, default=1)
        p.add_argument('--finetune', action='store_true', default=False)
        p.add_argument('--finetune_epochs', type=int, default=100)
    p.add_argument('--warmup', type=float, default=0.2)
    p.add_argument('--lr', type=float, default=0.001)
    p.add_argument('--lr_type', type=str, default='cosine')
    p.add_argument('--lr_steps', default=[], type=float, nargs="+")
    p.add_argument('--beta', type=float, default=0.5)
    p.add_argument('--loss_type', type=str, default='L2')
    p.add_argument('--lw', type=float, default=1.0)
    p.add_argument('--sigmoid', action='store_true')
    # evaluation
    p.add_argument('--deploy', action='store_true', default=False)
    p.add_argument('--eval_only', action='store_true', default=False)
    p.add_argument('--eval_freq', type=int, default=50)
    p.add_argument('--quant_bit', type=int, default=-1)
    p.add_argument('--quant_axis', type=int, default=0)
    p.add_argument('--dump_images', action='store_true', default=False)
    p.add_argument('--eval_fps', action='store_true', default=False)
    p.add_argument('--prune_steps', type=float, nargs='+', default=[0., ])
    p.add_argument('--prune_ratio', type=float, default=1.0)
    # distributed / misc
    p.add_argument('--manualSeed', type=int, default=1)
    p.add_argument('--init_method', default='tcp://127.0.0.1:9888', type=str)
    p.add_argument('-d', '--distributed', action='store_true', default=False,
                   help='frame-sharded data parallel; launch with torchrun, one process per GPU')
    p.add_argument('--debug', action='store_true')
    p.add_argument('-p', '--print_freq', default=50, type=int)
    p.add_argument('--weight', default='None', type=str)
    p.add_argument('--overwrite', action='store_true')
    p.add_argument('--outf', default='unify')
    p.add_argument('--suffix', default='')
    return p


SUPPORTED = ("supported on the B200 hot path: --branch_type NeRV_vanilla|ERB (the north-star path) and ACB|RepVGG|DBB|ECB "
             "(folded online into one convolution like ERB), every --act (swish is fused into the convolution, the "
             "others take one extra elementwise pass), --norm none, --single_res (multi-resolution heads with --lw run through "
             "the module API), "
             "--num_blocks 1, --stem_dim_num <dim>_1, --conv_type conv, --loss_type L2|L1|SSIM|Fusion1..Fusion9, "
             "--lr_type cosine|const|step, --finetune with --prune_ratio < 1 (NeRV_vanilla|ERB); README recipe: --embed 1.25_40 "
             "--stem_dim_num 512_1 --fc_hw_dim 9_16_26 --expansion 1 --reduction 2 --lower_width 96 --strides 5 2 2 2 2 "
             "--single_res --act swish --loss Fusion6 --branch_type ERB")


def validate_args(args):
    """One clear error for every flag combination outside the hot path (of the reference's own defaults the
    multi-resolution heads are), instead of a NotImplementedError or a plan failure deep inside the first step."""
    from .model import ACT_CODES, SUPPORTED_BRANCHES
    from .utils import LOSS_TERMS
    bad = []
    if args.branch_type not in SUPPORTED_BRANCHES:
        bad.append(f'--branch_type {args.branch_type}')
    if args.act not in ACT_CODES:
        bad.append(f'--act {args.act}')
    if getattr(args, 'finetune', False) and args.prune_ratio < 1 and args.branch_type not in ('NeRV_vanilla', 'ERB'):
        bad.append(f'--finetune with --branch_type {args.branch_type} (the reference handles NeRV_vanilla and ERB, '
                   'main_eval.py:238, :297)')
    if args.norm != 'none':
        bad.append(f'--norm {args.norm}')
    if not args.single_res:
        # multi-resolution heads (model.py:598-608): streaming head kernels up to 128 channels, plain wide-head kernels
        # up to 1024 on the earlier stages
        try:
            width, wide = int(str(args.fc_hw_dim).split('_')[2]), []
            for i, s_ in enumerate(args.strides):
                width = int(width * args.expansion) if i == 0 else max(width // (1 if s_ == 1 else args.reduction),
                                                                         args.lower_width)
                wide.append(width)
            if max(wide) > 1024 or wide[-1] > 128:
                bad.append(f'multi-resolution heads on stages of {max(wide)} channels (at most 1024 per early stage, '
                           '128 on the last)')
        except (IndexError, ValueError):
            pass
    if args.num_blocks != 1:
        bad.append(f'--num_blocks {args.num_blocks}')
    if args.conv_type != 'conv':
        bad.append(f'--conv_type {args.conv_type}')
    if args.loss_type not in LOSS_TERMS:
        bad.append(f'--loss_type {args.loss_type}')
    try:
        if int(str(args.stem_dim_num).split('_')[1]) != 1:
            bad.append(f'--stem_dim_num {args.stem_dim_num}')
        int(str(args.fc_hw_dim).split('_')[2])
    except (IndexError, ValueError):
        bad.append(f'--stem_dim_num {args.stem_dim_num} / --fc_hw_dim {args.fc_hw_dim}')
    if bad:
        raise SystemExit('orepnerv: unsupported configuration: ' + '; '.join(bad) + '\n' + SUPPORTED)


def finish_args(args):
    """Derived fields exactly like reference main_train.py:111-138."""
    validate_args(args)
    args.warmup = int(args.warmup * args.epochs)
    if args.debug:
        args.eval_freq = 1
        args.outf = 'result/debug'
    else:
        args.outf = os.path.join('result', args.outf)
    args.exp_id = (f'{args.dataset}/embed{args.embed}_{args.stem_dim_num}_fc_{args.fc_hw_dim}__exp{args.expansion}'
                   f'_reduce{args.reduction}_low{args.lower_width}_blk{args.num_blocks}_gap{args.frame_gap}'
                   f'_e{args.epochs}_warm{args.warmup}_b{args.batchSize}_{args.conv_type}_lr{args.lr}_{args.lr_type}'
                   f'_{args.loss_type}_act{args.act}_{args.suffix}')
    args.outf = os.path.join(args.outf, f'{args.suffix}')
    return args


def prepare_outdir(args, rank=0):
    if rank == 0:
        if args.overwrite and os.path.isdir(args.outf) and not args.eval_only:
            print('Will overwrite the existing output dir!')
            shutil.rmtree(args.outf)
        os.makedirs(args.outf, exist_ok=True)


def build_model(args, device, deploy=None):
    pe = PositionalEncoding(args.embed)
    args.embed_length = pe.embed_length
    model = Generator(embed_length=args.embed_length, stem_dim_num=args.stem_dim_num, fc_hw_dim=args.fc_hw_dim,
                      expansion=args.expansion, num_blocks=args.num_blocks, norm=args.norm, act=args.act, bias=True,
                      reduction=args.reduction, conv_type=args.conv_type, stride_list=args.strides,
                      sin_res=args.single_res, lower_width=args.lower_width, sigmoid=args.sigmoid,
                      deploy=args.deploy if deploy is None else deploy, branch_type=args.branch_type)
    return pe, model.to(device)


class FrameCache:
    """The whole clip as uint8 [N,3,H,W] resident in HBM plus the normalised indices i/N — the on-GPU
    replacement of reference CustomDataSet (model.py:11-70): same sorted file listing, `vid_list` / `frame_gap`
    sub-sampling, portrait frames transposed, index i/N.  365 MB for Bunny, 3.7 GB for a 600-frame 1080p clip.

    Sample k of the reference is (image of listing position k*frame_gap, frame_idx[k*frame_gap]) with
    `frame_idx = [i/N_listing]`, filtered to `vid_list` when one is given, and `len = len(frame_idx) // frame_gap`
    (model.py:37-44, :57-70).  Note what that means for `--vid`: the INDICES are those of the selected frames but the
    IMAGES are still taken from the head of the listing (`frame_path` is never filtered, model.py:58-59) — kept as is,
    the cache reproduces the reference's samples, not its intent."""

    def __init__(self, dataset, device, vid_list=(None,), frame_gap=1):
        m = re.fullmatch(r'synthetic:(\d+)x(\d+)x(\d+)', dataset)
        if m:
            n_listing, h, w = (int(x) for x in m.groups())
            load = None
        else:
            main_dir = f'../data/{dataset.lower()}'
            names = sorted(os.listdir(main_dir))                            # reference model.py:27-28
            n_listing = len(names)
            load = lambda pos: self._load_image(os.path.join(main_dir, names[pos]))    # noqa: E731
        frame_idx = [float(i) / n_listing for i in range(n_listing)]       # reference model.py:37
        if vid_list is not None and None not in list(vid_list):
            frame_idx = [frame_idx[i] for i in vid_list]                    # reference model.py:40-41
        valid = [k * frame_gap for k in range(len(frame_idx) // frame_gap)]    # reference model.py:50, :57
        if load is None:
            frames = synthetic_clip(n_listing, h, w, device=device)[valid]
        else:
            frames = torch.stack([load(v) for v in valid]).to(device)
        self.frames = frames.contiguous()
        self.t = torch.tensor([frame_idx[v] for v in valid], dtype=torch.float32, device=device)

    @staticmethod
    def _load_image(path):
        import numpy as np
        from PIL import Image
        img = np.asarray(Image.open(path).convert('RGB'))                   # reference model.py:60
        t = torch.from_numpy(img.copy()).permute(2, 0, 1)                   # uint8 [3,H,W]; k/255 on the device == ToTensor
        if t.size(1) > t.size(2):                                           # reference model.py:66-67
            t = t.permute(0, 2, 1)
        return t

    def __len__(self):
        return self.frames.size(0)


def strip_profiler_keys(state_dict):
    """thop leaves total_ops / total_params buffers in reference checkpoints (main_eval.py:229-234)."""
    return {k: v for k, v in state_dict.items() if 'total_ops' not in k and 'total_params' not in k}


def huffman_avg_bits(symbols):
    """Average Huffman code length (bits/symbol) of an integer tensor — the statistic the reference derives
    with dahuffman.HuffmanCodec.from_data (main_eval.py:673-692).  Returns (avg_bits, total_bits, n_symbols)."""
    import heapq
    vals, counts = torch.unique(symbols.reshape(-1), return_counts=True)
    counts = counts.tolist()
    if len(counts) == 1:
        return 1.0, float(counts[0]), 1
    # dahuffman adds an end-of-file symbol with count 1 to the table
    heap = [(c, i, None) for i, c in enumerate(counts)] + [(1, len(counts), None)]
    heapq.heapify(heap)
    depth = [0] * (len(counts) + 1)
    groups = {i: [i] for i in range(len(counts) + 1)}
    nxt = len(counts) + 1
    while len(heap) > 1:
        c1, i1, _ = heapq.heappop(heap)
        c2, i2, _ = heapq.heappop(heap)
        members = groups.pop(i1) + groups.pop(i2)
        for mbr in members:
            depth[mbr] += 1
        groups[nxt] = members
        heapq.heappush(heap, (c1 + c2, nxt, None))
        nxt += 1
    total_bits = float(sum(c * d for c, d in zip(counts, depth)))
    return total_bits / sum(counts), total_bits, len(counts)
