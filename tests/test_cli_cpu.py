"""CPU tests of the CLI layer: the reference's argparse surface (prefix matching included), derived fields,
and the Huffman code-length statistic that replaces dahuffman."""
import heapq

import torch

from orepnerv.cli_common import build_parser, finish_args, huffman_avg_bits, strip_profiler_keys

README_FLAGS = ("-e 300 --lr 0.0005 -b 1 --embed 1.25_40 --stem_dim_num 512_1 --fc_hw_dim 9_16_26 --expansion 1 "
                "--reduction 2 --lower_width 96 --strides 5 2 2 2 2 --num_blocks 1 --single_res --loss Fusion6 "
                "--warmup 0.2 --lr_type cosine --norm none --act swish --branch_type ERB --outf bunny --suffix erb "
                "--dataset bunny").split()


def test_readme_command_line_parses_like_the_reference():
    args = finish_args(build_parser().parse_args(README_FLAGS))
    assert args.loss_type == 'Fusion6'                 # `--loss` resolves by prefix (reference README.md:49)
    assert args.warmup == 60 and args.epochs == 300    # int(0.2 * 300), reference main_train.py:111
    assert args.outf == 'result/bunny/erb'
    assert args.strides == [5, 2, 2, 2, 2] and args.single_res and args.branch_type == 'ERB'
    ev = finish_args(build_parser(eval_mode=True).parse_args(README_FLAGS + ['--prune_ratio', '0.2', '--quant_bit', '8']))
    assert ev.prune_ratio == 0.2 and ev.quant_bit == 8 and ev.finetune is False


def test_strip_profiler_keys():
    sd = {'stem.0.weight': 1, 'total_ops': 2, 'layers.0.total_params': 3}
    assert list(strip_profiler_keys(sd)) == ['stem.0.weight']


def _huffman_lengths(counts):
    heap = [(c, i, [i]) for i, c in enumerate(counts)]
    heapq.heapify(heap)
    depth = [0] * len(counts)
    nxt = len(counts)
    while len(heap) > 1:
        c1, _, m1 = heapq.heappop(heap)
        c2, _, m2 = heapq.heappop(heap)
        for m in m1 + m2:
            depth[m] += 1
        heapq.heappush(heap, (c1 + c2, nxt, m1 + m2))
        nxt += 1
    return depth


def test_huffman_avg_bits():
    g = torch.Generator().manual_seed(0)
    sym = torch.randint(0, 40, (5000,), generator=g).float()
    sym[:2000] = 7.0
    avg, total, n = huffman_avg_bits(sym)
    vals, counts = torch.unique(sym, return_counts=True)
    depth = _huffman_lengths(counts.tolist() + [1])          # + dahuffman's EOF symbol
    expect = sum(c * d for c, d in zip(counts.tolist(), depth))
    assert n == len(vals) and abs(avg - expect / 5000) < 1e-12
    p = counts.double() / counts.sum()
    assert avg >= float(-(p * p.log2()).sum()) - 1e-9


def _reference_cli():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "cli_flags.json")) as f:
        return json.load(f)


def test_every_reference_flag_exists_with_the_same_default():
    """tests/golden/cli_flags.json is the argparse surface of the unmodified reference (main_train.py / main_eval.py,
    extracted by tests/golden/make_cli_golden.py).  The mirrored parsers must know every flag (long and short
    spelling) and give it the reference default."""
    ref = _reference_cli()
    for script, eval_mode in (("main_train.py", False), ("main_eval.py", True)):
        parser = build_parser(eval_mode=eval_mode)
        known = {s: a for a in parser._actions for s in a.option_strings}
        for spec in ref[script]:
            for flag in spec['flags']:
                assert flag in known, (script, flag)
            act = known[spec['flags'][-1]]
            if 'default' in spec and not isinstance(spec['default'], str):
                assert act.default == spec['default'], (script, spec['flags'], act.default, spec['default'])
            elif spec.get('action') == 'store_true':
                assert act.default is False and act.nargs == 0, (script, spec['flags'])
            if 'nargs' in spec:
                assert act.nargs == spec['nargs'], (script, spec['flags'])


def _write_clip(root, n=7, h=6, w=9, portrait_at=2):
    import numpy as np
    from PIL import Image
    import os
    d = os.path.join(root, "data", "clipx")
    os.makedirs(d)
    rng = np.random.RandomState(0)
    for i in range(n):
        shape = (w, h, 3) if i == portrait_at else (h, w, 3)            # one portrait frame (transposed on load)
        Image.fromarray(rng.randint(0, 256, shape, dtype=np.uint8)).save(os.path.join(d, f"f{i:03d}.png"))
    os.makedirs(os.path.join(root, "run"))
    return os.path.join(root, "run")


def test_frame_cache_reproduces_the_reference_dataset(tmp_path, monkeypatch):
    """FrameCache (directory branch: PIL decode once, uint8, portrait transpose, i/N index, --vid / frame_gap) against
    the UNMODIFIED reference CustomDataSet (model.py:11-70) on the same directory of PNG frames — including its
    `--vid` behaviour (indices of the selection, images from the head of the listing)."""
    import os
    import sys
    import numpy as np
    import pytest
    from orepnerv.cli_common import FrameCache
    run = _write_clip(str(tmp_path))
    monkeypatch.chdir(run)                                                  # the reference reads ../data/<dataset>
    cases = [dict(vid_list=[None], frame_gap=1), dict(vid_list=[None], frame_gap=3), dict(vid_list=[1, 4, 6, 2], frame_gap=1),
             dict(vid_list=[5, 0, 3], frame_gap=2)]
    caches = [FrameCache("ClipX", torch.device("cpu"), **kw) for kw in cases]
    c0 = caches[0]
    assert c0.frames.dtype == torch.uint8 and tuple(c0.frames.shape) == (7, 3, 6, 9)
    assert torch.equal(c0.t, torch.arange(7, dtype=torch.float32) / 7)
    assert len(caches[1]) == 2 and len(caches[2]) == 4 and len(caches[3]) == 1
    if not os.path.isdir("/root/reference"):
        pytest.skip("the reference tree is not on this machine: checked the layout only")
    sys.path.insert(0, "/root/reference")
    monkeypatch.setattr(np, "asfarray", lambda a: np.asarray(a, dtype=float), raising=False)   # removed in NumPy 2
    import model as ref_model                                               # reference, unmodified
    from torchvision import transforms
    for kw, cache in zip(cases, caches):
        ds = ref_model.CustomDataSet("../data/clipx", transforms.ToTensor(), **kw)
        assert len(ds) == len(cache), kw
        for k in range(len(ds)):
            img, idx = ds[k]
            assert torch.equal(cache.frames[k].float().div(255), img), (kw, k)     # bit-identical to ToTensor
            assert cache.t[k].item() == idx.item(), (kw, k)


def test_cli_accepts_the_widened_rows_and_rejects_the_rest():
    import pytest
    base = [a for a in README_FLAGS]
    for extra in (['--branch_type', 'DBB'], ['--branch_type', 'ECB', '--act', 'gelu'], ['--act', 'leaky01'],
                  ['--fc_hw_dim', '9_16_128', '--expansion', '8']):
        args = finish_args(build_parser().parse_args(base + extra))
        assert args.branch_type in ('ERB', 'DBB', 'ECB')
    ev = finish_args(build_parser(eval_mode=True).parse_args(base + ['--prune_ratio', '0.4', '--finetune',
                                                                     '--finetune_epochs', '5']))
    assert ev.finetune and ev.finetune_epochs == 5
    for extra in (['--norm', 'bn'], ['--num_blocks', '2'], ['--conv_type', 'deconv'], ['--loss_type', 'Fusion10']):
        with pytest.raises(SystemExit):
            finish_args(build_parser().parse_args(base + extra))
    no_single = [a for a in base if a != '--single_res']
    assert not finish_args(build_parser().parse_args(no_single + ['--lw', '0.5'])).single_res   # <= 128 channels per stage
    finish_args(build_parser().parse_args(no_single + ['--fc_hw_dim', '9_16_128', '--expansion', '8']))   # wide-head kernels
    with pytest.raises(SystemExit):                                         # a head on a 2048-channel stage
        finish_args(build_parser().parse_args(no_single + ['--fc_hw_dim', '9_16_128', '--expansion', '16']))
    with pytest.raises(SystemExit):                                         # the reference finetunes NeRV_vanilla | ERB only
        finish_args(build_parser(eval_mode=True).parse_args(base + ['--branch_type', 'DBB', '--prune_ratio', '0.4',
                                                                    '--finetune']))


def test_evaluate_host_logic(monkeypatch, tmp_path, capsys):
    """main_train.evaluate (reference main_train.py:377-438): frame selection by test_gap, --eval_fps repeats, per-stage
    metrics, the reference's progress line and the analytic MACs — executed on the CPU with the device metric calls
    replaced by the oracle (the arithmetic itself is covered by the GPU parity tests)."""
    import argparse
    import orepnerv.utils as U
    from oracle import nerv_oracle as O
    from orepnerv import main_train
    from orepnerv.model import Generator
    import torch.nn.functional as F
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(U, "adaptive_avg_pool2d", lambda d, size: F.adaptive_avg_pool2d(d, tuple(size)))
    monkeypatch.setattr(U, "psnr_fn", lambda outs, tgts: torch.cat(
        [O.psnr(o, t).view(1, 1).expand(o.size(0), -1) for o, t in zip(outs, tgts)], dim=1))
    monkeypatch.setattr(main_train, "msssim_fn", lambda outs, tgts: torch.zeros(1, len(outs)).expand(outs[-1].size(0), -1))
    gen = Generator(embed_length=8, stem_dim_num='16_1', fc_hw_dim='3_4_4', expansion=1, num_blocks=1, norm='none',
                    act='swish', bias=True, reduction=2, conv_type='conv', stride_list=[2, 2], sin_res=False,
                    lower_width=4, sigmoid=False, deploy=False, branch_type='NeRV_vanilla')
    calls = []

    def fake_forward(embed):
        calls.append(float(embed.sum()))
        g = torch.Generator().manual_seed(len(calls))
        return [torch.rand(1, 3, 6, 8, generator=g), torch.rand(1, 3, 12, 16, generator=g)]
    gen.forward = fake_forward

    class Clip:
        frames = torch.randint(0, 256, (5, 3, 12, 16), generator=torch.Generator().manual_seed(0)).to(torch.uint8)
        t = torch.arange(5, dtype=torch.float32) / 5

        def __len__(self):
            return 5
    args = argparse.Namespace(test_gap=2, eval_fps=True, debug=False, print_freq=1)
    log = tmp_path / "rank0.txt"
    psnr, msssim, fps = main_train.evaluate(gen, Clip(), lambda t: t.view(1, 1).repeat(1, 8), args, 0, str(log))
    assert len(calls) == 3 * 10                                   # frames 0, 2, 4; --eval_fps: 10 forwards each
    assert psnr.shape == (2,) and msssim.shape == (2,) and fps > 0 and gen.training
    lines = log.read_text().strip().split('\n')
    assert len(lines) == 3 and lines[-1].startswith('Rank:0, Step [3/3], PSNR: ') and ' FPS: ' in lines[-1]
    # stem 8*16 + 16*48, blocks 3*4*16*4*9 + 6*8*16*4*9, heads 6*8*4*3 + 12*16*4*3
    assert main_train.decoder_macs(gen) == 8 * 16 + 16 * 48 + 12 * 16 * 4 * 9 + 48 * 16 * 4 * 9 + 48 * 12 + 192 * 12
    assert 'MACs: ' in capsys.readouterr().out
