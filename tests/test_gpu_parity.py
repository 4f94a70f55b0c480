"""GPU parity tests (B200): the CUDA path, called through the C ABI / the reference-shaped Python API,
against (a) the golden vectors produced by the unmodified reference and (b) the CPU oracle on the same
seeded inputs.

Tolerances (stated per north_star):
  * ERB fold and its backward: fp32 kernels, rel-L2 <= 1e-5 vs the fp32 oracle;
  * positional encoding: <= 2e-6 abs (device sinf/cosf vs CPU, arguments up to ~2e4 rad);
  * anything that passes through the bf16 tensor-core convolutions (activations, images, losses, gradients):
    rel-L2 <= 2e-2 on tensors, |d loss| <= 2e-3; integer results (quantisation codes, prune masks): bit exact.
"""
import argparse
import math
import os
import subprocess

import pytest
import torch

import fullsize_util as U      # tests/ is on sys.path
from oracle import nerv_oracle as O

pytestmark = pytest.mark.gpu
# movement of each trained tensor vs the reference / oracle (U.movement_ok).  Measured on B200 over all these tests
# (profiles/r03_movement_ratios.jsonl): median 0.2 - 1.5 %, worst 9.2 % (tiny_erb, a 1x3 branch of 48 elements whose
# gradient signs sit in the bf16 noise: Adam turns a sign flip into a full lr-sized step).  A wrong Adam bias correction
# is off by 15x at step 1.  The wgrad accumulates with atomics (run-to-run noise in the last bits decides those signs), so the
# bound keeps 2.7x headroom over the worst measured ratio.
MOVE_TOL = 0.25
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from orepnerv import _lib
    _lib.lib()          # raises unless the device is sm_100
    return torch.device("cuda:0")


def build(cfg, branch_type, dev, deploy=False, seed=1):
    from orepnerv.model import Generator
    from orepnerv.utils import PositionalEncoding
    torch.manual_seed(seed)
    pe = PositionalEncoding(cfg['embed'])
    gen = Generator(embed_length=pe.embed_length, stem_dim_num=cfg['stem_dim_num'], fc_hw_dim=cfg['fc_hw_dim'],
                    expansion=cfg['expansion'], num_blocks=1, norm='none', act='swish', bias=True,
                    reduction=cfg['reduction'], conv_type='conv', stride_list=cfg['strides'], sin_res=True,
                    lower_width=cfg['lower_width'], sigmoid=False, deploy=deploy, branch_type=branch_type)
    return pe, gen.to(dev)


def ocfg(cfg):
    fh, fw, fd = [int(x) for x in cfg['fc_hw_dim'].split('_')]
    return dict(fc_h=fh, fc_w=fw, fc_dim=fd, strides=cfg['strides'], sigmoid=False)


# ------------------------------------------------------------------------------------------- kernels
def test_selftest_binary_tcgen05_vs_simt(dev):
    """tcgen05 fprop / dgrad / wgrad against the SIMT kernels on identical bf16 inputs (device cross-check)."""
    exe = os.path.join(ROOT, "boosting-neural-video-representation-via-online-structural-reparameteration_b200",
                       "onr_selftest")
    for op in ("fprop", "dgrad", "wgrad"):
        for shape in ("tiny", "l0", "l1", "l2s", "b2", "u3", "wide"):
            r = subprocess.run([exe, op, shape, "0"], capture_output=True, text=True, timeout=120)
            assert r.returncode == 0, r.stdout + r.stderr


def test_positional_encoding(dev, golden):
    from orepnerv.utils import PositionalEncoding
    m = golden('misc.pt')
    pe = PositionalEncoding('1.25_40')
    out = pe(m['pe_pos'])
    assert out.is_cuda and out.shape == m['pe_embed'].shape
    assert (out.cpu() - m['pe_embed']).abs().max().item() <= 2e-6


@pytest.mark.parametrize("cin,cout", [(4, 16), (12, 32), (26, 650), (96, 384), (112, 2800), (26, 864)])
def test_erb_fold_forward_backward(dev, cin, cout):
    from orepnerv.model import NeRVBlock
    torch.manual_seed(3)
    s = int(round(math.sqrt(cout // max(cin, 1)))) if cout % cin == 0 else 1
    blk = NeRVBlock(ngf=cin, new_ngf=cout, stride=1, bias=True, norm='none', act='swish', deploy=False,
                    conv_type='conv', branch_type='ERB').to(dev)
    K, b = blk.get_equivalent_kernel_bias()
    names = ['rbr_3x3_branch.weight', 'rbr_3x3_branch.bias', 'rbr_1x3_branch.weight', 'rbr_1x3_branch.bias',
             'rbr_3x1_branch.weight', 'rbr_3x1_branch.bias', 'rbr_1x1_3x3_1x1_branch_1x1_1.weight',
             'rbr_1x1_3x3_1x1_branch_3x3.weight', 'rbr_1x1_3x3_1x1_branch_1x1_2.weight']
    sd = {k: v.detach().cpu() for k, v in blk.state_dict().items()}
    ws = [sd[n] for n in names]
    # the reference arithmetic in float64 is the ground truth of the fp32 gate (an fp32 CPU einsum over K = 9*Cout =
    # 25 200 terms carries ~1e-5 of rounding error of its own)
    ws = [w.double() for w in ws]
    K_ref, b_ref = O.erb_fold(*ws)
    assert rel_l2(K, K_ref) <= 1e-5 and rel_l2(b, b_ref) <= 1e-5
    g = torch.Generator().manual_seed(4)
    dK, db = torch.randn(K.shape, generator=g), torch.randn(b.shape, generator=g)
    (K * dK.to(dev)).sum().add((b * db.to(dev)).sum()).backward()
    ref = O.erb_fold_backward(dK.double(), db.double(), ws[6], ws[7], ws[8])
    order = ['w3x3', 'b3x3', 'w1x3', 'b1x3', 'w3x1', 'b3x1', 'w1', 'w2', 'w3']
    params = dict(blk.named_parameters())
    for n, o in zip(names, order):
        assert rel_l2(params[n].grad, ref[o]) <= 1e-5, n
    # bit-reproducible: split-K partials are combined in a fixed order (replicas of a data-parallel run must not drift)
    first = {n: params[n].grad.clone() for n in names}
    for p_ in params.values():
        p_.grad = None
    K2, b2 = blk.get_equivalent_kernel_bias()
    assert torch.equal(K2, K) and torch.equal(b2, b)
    (K2 * dK.to(dev)).sum().add((b2 * db.to(dev)).sum()).backward()
    for n in names:
        assert torch.equal(params[n].grad, first[n]), n


@pytest.mark.parametrize("cin,cnew,s,h,w", [(4, 4, 2, 6, 8), (26, 26, 5, 9, 16), (26, 96, 2, 13, 21), (96, 96, 3, 8, 16)])
def test_block_forward_backward(dev, cin, cnew, s, h, w):
    """NeRVBlock.forward on NCHW fp32 tensors vs F.conv2d + pixel_shuffle + SiLU of the oracle."""
    from orepnerv.model import NeRVBlock
    torch.manual_seed(5)
    blk = NeRVBlock(ngf=cin, new_ngf=cnew, stride=s, bias=True, norm='none', act='swish', deploy=False,
                    conv_type='conv', branch_type='NeRV_vanilla').to(dev)
    x = torch.randn(2, cin, h, w)
    xg = x.to(dev).requires_grad_(True)
    y = blk(xg)
    Kc, bc = blk.branch.weight.detach().cpu().requires_grad_(True), blk.branch.bias.detach().cpu().requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    y_ref = O.block_forward(xc, Kc, bc, s)
    assert y.shape == y_ref.shape
    assert rel_l2(y, y_ref) <= 1e-2
    gy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(6))
    y.backward(gy.to(dev))
    y_ref.backward(gy)
    assert rel_l2(xg.grad, xc.grad) <= 2e-2
    assert rel_l2(blk.branch.weight.grad, Kc.grad) <= 2e-2
    assert rel_l2(blk.branch.bias.grad, bc.grad) <= 2e-2


# ------------------------------------------------------------------------------------------- whole decoder
@pytest.mark.parametrize("name,bt", [("tiny_erb.pt", "ERB"), ("tiny_vanilla.pt", "NeRV_vanilla"), ("small_erb.pt", "ERB")])
def test_generator_against_reference_golden(dev, golden, name, bt):
    from orepnerv.utils import loss_fn, psnr_fn
    g = golden(name)
    pe, gen = build(g['cfg'], bt, dev)
    embed = pe(g['pos'])
    img = gen(embed)[0]
    assert rel_l2(img, g['img']) <= 1e-2
    args = argparse.Namespace(loss_type='Fusion6')
    loss = loss_fn(img, g['target'].to(dev), args)
    assert abs(loss.item() - g['loss'].item()) <= 2e-3
    loss.backward()
    for k, p in gen.named_parameters():
        ref = g['grads'][k]
        assert p.grad is not None, k
        # tiny tensors (biases of 3-16 elements) carry bf16 noise: compare against the gradient scale
        err = (p.grad.cpu() - ref).norm().item()
        assert err <= 3e-2 * ref.norm().item() + 1e-6, (k, err, ref.norm().item())
    psnr = psnr_fn([img.detach()], [g['target'].to(dev)])
    assert abs(psnr[0, 0].item() - g['psnr'][0, 0].item()) <= 0.05


def test_deploy_equals_train_state(dev, golden):
    """switch_to_deploy + state-dict round trip: deploy decode == train-state decode (reference gate:
    identical; here both go through the same bf16 operand packing, so they are bit identical too)."""
    import copy
    g = golden("small_erb.pt")
    pe, gen = build(g['cfg'], "ERB", dev)
    embed = pe(g['pos'])
    with torch.no_grad():
        img_train = gen(embed)[0]
    dep = copy.deepcopy(gen)
    for blk in dep.layers:
        blk.switch_to_deploy()
    assert list(dep.state_dict().keys()) == list(g['deploy_state'].keys())
    for i, (K_ref, b_ref) in enumerate(g['folded']):
        assert rel_l2(dep.layers[i].rbr_reparam.weight, K_ref) <= 1e-5
        assert rel_l2(dep.layers[i].rbr_reparam.bias, b_ref) <= 1e-5
    _, dep2 = build(g['cfg'], "ERB", dev, deploy=True)
    dep2.load_state_dict(dep.state_dict())
    with torch.no_grad():
        img_dep = dep2(embed)[0]
    assert torch.equal(img_dep, img_train)
    assert rel_l2(img_dep, g['deploy_img']) <= 1e-2


@pytest.mark.parametrize("name,bt", [("tiny_erb.pt", "ERB"), ("tiny_vanilla.pt", "NeRV_vanilla")])
def test_training_steps_reference_loop(dev, golden, name, bt):
    """The reference's own loop (main_train.py:238-250) written against our modules + FusedAdam."""
    from orepnerv.optim import FusedAdam
    from orepnerv.utils import loss_fn, adjust_lr
    g = golden(name)
    pe, gen = build(g['cfg'], bt, dev)
    embed = pe(g['pos'])
    target = g['target'].to(dev)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5)
    opt = FusedAdam(gen.parameters(), betas=(0.5, 0.999))
    for i, (loss_ref, lr_ref) in enumerate(zip(g['train_losses'], g['train_lrs'])):
        out = gen(embed)[0]
        loss = loss_fn(out, target, args)
        lr = adjust_lr(opt, 0, i, 4, args)
        assert lr == lr_ref
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(loss.item() - loss_ref) <= 3e-3
    sd = gen.state_dict()
    # Adam normalises the update to ~lr per element, so compare the MOVEMENT of each tensor
    for k, v in g['trained_state'].items():
        moved_ref = v - g['init_state'][k]
        moved = sd[k].cpu() - g['init_state'][k]
        assert U.movement_ok(f"reference_loop[{name}]", k, moved, moved_ref, MOVE_TOL), k
    osd = opt.state_dict()
    assert set(osd['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}


def test_frame_fitter_matches_oracle_steps(dev, golden):
    """Fast path (FrameFitter, CUDA graph) vs the oracle's train_step on the same frames / schedule."""
    from orepnerv.trainer import FrameFitter
    g = golden("small_erb.pt")
    cfg = g['cfg']
    pe, gen = build(cfg, "ERB", dev)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5, beta=0.5,
                              batchSize=2)
    fit = FrameFitter(gen, pe, args, data_size=4, steps_per_epoch=2, use_graph=True, with_msssim=False)
    frames_u8 = (g['target'] * 255).round().to(torch.uint8)
    target = frames_u8.float().div(255)
    sd, state = {k: v.clone() for k, v in g['init_state'].items()}, {}
    embed = O.pos_encoding(g['pos'], 1.25, 40)
    for t in range(4):
        out = fit.step(frames_u8.to(dev), g['pos'].to(dev)).clone()
        epoch, it = t // 2, t % 2
        lr = O.lr_at(epoch, it, 4, 5e-4, 1, 5)
        sd, state, loss, img, _ = O.train_step(sd, state, embed, target, ocfg(cfg), lr, t + 1)
        assert abs(out[0].item() - loss.item()) <= 3e-3, (t, out[0].item(), loss.item())
        assert abs(out[4].item() - O.psnr(img, target).item()) <= 0.05
    for k, v in gen.state_dict().items():
        moved_ref = sd[k] - g['init_state'][k]
        moved = v.cpu() - g['init_state'][k]
        assert U.movement_ok("frame_fitter_steps", k, moved, moved_ref, MOVE_TOL), k


# ------------------------------------------------------------------------------------------- loss / metrics
@pytest.mark.parametrize("shape", [(1, 3, 40, 56), (2, 3, 77, 130)])
def test_fusion6_forward_backward(dev, shape):
    from orepnerv.utils import loss_fn
    g = torch.Generator().manual_seed(8)
    a = torch.rand(shape, generator=g)
    b = (a + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    for lt in ("Fusion6", "L1", "SSIM", "L2", "Fusion1", "Fusion7"):
        p = a.clone().requires_grad_(True)
        ref = O.loss_fn(p, b, lt)
        ref.backward()
        pg = a.to(dev).requires_grad_(True)
        loss = loss_fn(pg, b.to(dev), argparse.Namespace(loss_type=lt))
        loss.backward()
        assert abs(loss.item() - ref.item()) <= 2e-6
        assert rel_l2(pg.grad, p.grad) <= 2e-5


def test_metrics_against_golden(dev, golden):
    from orepnerv.utils import psnr_fn, msssim_fn
    m = golden('misc.pt')
    a, b = m['metric_in']
    assert abs(psnr_fn([a.to(dev)], [b.to(dev)]).item() - m['metric_psnr'].item()) <= 1e-3
    assert abs(msssim_fn([a.to(dev)], [b.to(dev)]).item() - m['metric_msssim'].item()) <= 2e-5
    # odd sizes exercise avg_pool2d padding = dim % 2 (U1080 path: 135 is odd)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(1, 3, 270, 201, generator=g)
    y = (x + 0.05 * torch.randn(x.shape, generator=g)).clamp(0, 1)
    assert abs(msssim_fn([x.to(dev)], [y.to(dev)]).item() - O.ms_ssim(x, y).item()) <= 2e-5
    small = torch.rand(1, 3, 64, 64)
    assert msssim_fn([small.to(dev)], [small.to(dev)]).item() == 0.0      # H < 160 -> 0 (utils.py:204-207)
    # the training step hands MS-SSIM the loss workspace so that scale 0 (= the SSIM of the loss) is not filtered
    # twice: same value as the standalone evaluation
    from orepnerv import _lib
    from orepnerv._lib import check, ptr
    lib = _lib.lib()
    xd, yd = x.to(dev).contiguous(), y.to(dev).contiguous()
    B, _, H, W = xd.shape
    lw = torch.zeros(lib.onr_loss_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    mw = torch.zeros(lib.onr_msssim_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    out, ms = torch.zeros(8, device=dev), torch.zeros(2, device=dev)
    check(lib.onr_fusion6_fwd_bwd(ptr(xd), ptr(yd), B, H, W, 0.7, 0.3, 1.0, ptr(out), None, ptr(lw), _lib.stream()),
          "fusion6")
    check(lib.onr_msssim(ptr(xd), ptr(yd), B, H, W, ptr(ms[0:1]), ptr(mw), ptr(lw), _lib.stream()), "msssim")
    check(lib.onr_msssim(ptr(xd), ptr(yd), B, H, W, ptr(ms[1:2]), ptr(mw), None, _lib.stream()), "msssim")
    assert abs(ms[0].item() - ms[1].item()) <= 1e-6
    assert abs(ms[0].item() - O.ms_ssim(x, y).item()) <= 2e-5


def test_adam_matches_oracle(dev):
    from orepnerv.optim import FusedAdam
    torch.manual_seed(10)
    shapes = [(3,), (17, 5), (4, 3, 3, 3), (1025,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    ref = [(p.detach().cpu().clone(), torch.zeros(s), torch.zeros(s)) for p, s in zip(ps, shapes)]
    opt = FusedAdam(ps, lr=1e-3, betas=(0.5, 0.999))
    for t in range(1, 6):
        gs = [torch.randn(s) for s in shapes]
        for p, gr in zip(ps, gs):
            p.grad = gr.to(dev)
        opt.param_groups[0]['lr'] = 1e-3 * t
        opt.step()
        ref = [O.adam_step(p, gr, m, v, t, 1e-3 * t) for (p, m, v), gr in zip(ref, gs)]
    for p, (pr, _, _) in zip(ps, ref):
        torch.testing.assert_close(p.detach().cpu(), pr, rtol=2e-5, atol=2e-7)


# ------------------------------------------------------------------------------------------- eval-side ops
def test_quantize_bit_exact(dev, golden):
    from orepnerv.utils import quantize_per_tensor
    m = golden('misc.pt')
    qi, qo = m['quant_in'], m['quant_out']
    for key, (t, bit, axis) in {'t4_axis0': (qi['t4'], 8, 0), 't4_axis1': (qi['t4'], 8, 1),
                                't2_axis0': (qi['t2'], 8, 0), 't2_axis-1': (qi['t2'], 6, -1),
                                't1_axis-1': (qi['t1'], 8, -1)}.items():
        q, new = quantize_per_tensor(t.to(dev), bit, axis)
        assert torch.equal(q.cpu(), qo[key][0]), key
        assert torch.equal(new.cpu(), qo[key][1]), key
    # 257 levels for "8 bit" 
    t = torch.linspace(-1, 1, 4096).view(1, -1)
    q, _ = quantize_per_tensor(t.to(dev), 8, 0)
    assert q.min().item() == 0 and q.max().item() == 256


def test_global_prune_threshold(dev):
    from orepnerv.utils import global_magnitude_threshold
    torch.manual_seed(11)
    ts = [torch.randn(40, 80), torch.randn(300, 26, 3, 3), torch.randn(7)]
    for amount in (0.2, 0.5, 0.013):
        thr_ref, k_ref = O.prune_threshold(ts, amount)
        thr, k = global_magnitude_threshold([t.to(dev) for t in ts], amount)
        assert k == k_ref and thr == thr_ref


def test_eval_pipeline_prune_quant(dev, golden):
    """Reference main_eval.py:551-712 flow on a deploy model: global prune (torch.nn.utils.prune drives the
    masks, as in the reference), 8-bit quantisation of the state dict, load back, decode."""
    import torch.nn.utils.prune as prune
    from orepnerv.utils import quantize_per_tensor
    g = golden("small_erb.pt")
    pe, dep = build(g['cfg'], "ERB", dev, deploy=True)
    dep.load_state_dict({k: v for k, v in g['deploy_state'].items()})
    mods = [dep.stem[0], dep.stem[2]] + [blk.rbr_reparam for blk in dep.layers]
    prune.global_unstructured([(m, 'weight') for m in mods], pruning_method=prune.L1Unstructured, amount=0.2)
    embed = pe(g['pos'])
    with torch.no_grad():
        img_pruned = dep(embed)[0]
    # oracle on the pruned weights
    sd_ref = {k: v.clone() for k, v in g['deploy_state'].items()}
    thr, _ = O.prune_threshold([sd_ref[k] for k in ('stem.0.weight', 'stem.2.weight', 'layers.0.rbr_reparam.weight',
                                                    'layers.1.rbr_reparam.weight')], 0.2)
    for k in ('stem.0.weight', 'stem.2.weight', 'layers.0.rbr_reparam.weight', 'layers.1.rbr_reparam.weight'):
        sd_ref[k] = sd_ref[k] * (sd_ref[k].abs() > thr)
    img_ref = O.generator_forward(sd_ref, O.pos_encoding(g['pos'], 1.25, 40), ocfg(g['cfg']))
    assert rel_l2(img_pruned, img_ref) <= 1e-2
    # quantise every state-dict tensor like main_eval.py:660-669, 703 and decode again
    cur = dep.state_dict()
    for k, v in cur.items():
        large = v.dim() in {2, 4} and 'bias' not in k
        _, new_v = quantize_per_tensor(v, 8, 0 if large else -1)
        q_ref, new_ref = O.quantize_per_tensor(v.cpu(), 8, 0 if large else -1)
        assert torch.equal(new_v.cpu(), new_ref), k
        cur[k] = new_v.to(v.device).type_as(v)
    dep.load_state_dict(cur)
    with torch.no_grad():
        img_q = dep(embed)[0]
    assert torch.isfinite(img_q).all() and rel_l2(img_q, img_pruned) < 0.2
    # what the reference decodes after load_state_dict (main_eval.py:703): every pruned module recomputes
    # weight = weight_orig * weight_mask from the QUANTISED entries on its next forward — stem Linears included
    sd_q = {}
    for k, v in cur.items():
        if k.endswith('weight_orig'):
            sd_q[k[:-len('_orig')]] = (v * cur[k[:-len('orig')] + 'mask']).cpu()
        elif not k.endswith('weight_mask'):
            sd_q[k] = v.cpu()
    img_ref_q = O.generator_forward(sd_q, O.pos_encoding(g['pos'], 1.25, 40), ocfg(g['cfg']))
    assert rel_l2(img_q, img_ref_q) <= 1e-2
    # A13 quirk: quantising weight_mask (min = max = 1 per row) turns every mask entry into 1
    # (rows that were pruned completely keep their zeros)
    assert all((v == 1).float().mean().item() > 0.99 for k, v in cur.items() if k.endswith('weight_mask'))


# ------------------------------------------------------------------------------------------- full size
def test_s720_properties(dev):
    """BASELINE configs[1] geometry (fc 9_16_26, strides 5 2 2 2 2 -> 720x1280): size-independent properties.
    (1) deploy decode == train-state decode bit for bit; (2) the fold is linear in the 3x3 branch;
    (3) one FrameFitter step lowers the loss on the frame it was fitted to; (4) block-4 tcgen05 kernels
    agree with the SIMT kernels at full size (selftest l3)."""
    import copy
    from orepnerv.trainer import FrameFitter
    from orepnerv.data import synthetic_clip
    cfg = dict(embed='1.25_40', stem_dim_num='512_1', fc_hw_dim='9_16_26', expansion=1, reduction=2,
               lower_width=96, strides=[5, 2, 2, 2, 2])
    pe, gen = build(cfg, "ERB", dev)
    assert sum(p.numel() for p in gen.parameters()) == 7576025
    pos = torch.tensor([3 / 132], device=dev)
    with torch.no_grad():
        img = gen(pe(pos))[0]
    assert img.shape == (1, 3, 720, 1280) and torch.isfinite(img).all()
    dep = copy.deepcopy(gen)
    for blk in dep.layers:
        blk.switch_to_deploy()
    assert sum(p.numel() for p in dep.parameters()) == 3201905
    with torch.no_grad():
        img_dep = dep(pe(pos))[0]
    assert torch.equal(img, img_dep)
    blk = gen.layers[2]
    with torch.no_grad():
        K0, _ = blk.get_equivalent_kernel_bias()
        blk.rbr_3x3_branch.weight.mul_(2.0)
        K1, _ = blk.get_equivalent_kernel_bias()
        blk.rbr_3x3_branch.weight.mul_(0.5)
        assert rel_l2(K1 - K0, blk.rbr_3x3_branch.weight) <= 1e-5
    frames = synthetic_clip(2, 720, 1280, device=dev)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=0, epochs=300, beta=0.5,
                              batchSize=1)
    fit = FrameFitter(gen, pe, args, data_size=2, steps_per_epoch=2, use_graph=True)
    losses = []
    for _ in range(12):
        losses.append(fit.step(frames[0:1], pos).clone())
    losses = torch.stack(losses).cpu()
    assert torch.isfinite(losses).all()
    assert losses[-1, 0] < losses[0, 0] and losses[-1, 4] > losses[0, 4]       # loss down, PSNR up
    assert 0.0 < losses[-1, 5] <= 1.0                                           # MS-SSIM in range
    exe = os.path.join(ROOT, "boosting-neural-video-representation-via-online-structural-reparameteration_b200",
                       "onr_selftest")
    for op in ("fprop", "dgrad", "wgrad"):
        r = subprocess.run([exe, op, "l3", "0"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr


def test_cli_train_then_eval(dev, tmp_path, monkeypatch):
    """README workflow end to end at toy size: main_train (ERB, synthetic clip) writes the reference's checkpoint
    files; main_eval loads the deploy checkpoint, prunes 20 %, quantises to 8 bit and decodes."""
    from orepnerv import main_train, main_eval
    monkeypatch.chdir(tmp_path)
    flags = ("-e 3 --lr 0.002 -b 1 --embed 1.25_40 --stem_dim_num 64_1 --fc_hw_dim 3_4_12 --expansion 1 --reduction 2 "
             "--lower_width 8 --strides 3 2 --single_res --loss Fusion6 --warmup 0.2 --lr_type cosine --norm none "
             "--act swish --branch_type ERB --outf toy --suffix erb --dataset synthetic:6x18x24 --eval_freq 1 -p 100"
             ).split()
    main_train.main(flags + ['--overwrite'])
    out = tmp_path / 'result' / 'toy' / 'erb'
    for f in ('model_latest.pth', 'model_latest_deploy.pth', 'model_train_best.pth', 'model_val_best.pth', 'rank0.txt'):
        assert (out / f).exists(), f
    ck = torch.load(out / 'model_latest.pth', weights_only=True)
    assert set(ck) == {'epoch', 'state_dict', 'train_best_psnr', 'train_best_msssim', 'val_best_psnr',
                       'val_best_msssim', 'optimizer'}
    assert 'layers.0.rbr_1x1_3x3_1x1_branch_3x3.weight' in ck['state_dict']
    dk = torch.load(out / 'model_latest_deploy.pth', weights_only=True)
    assert 'layers.0.rbr_reparam.weight' in dk['state_dict'] and 'layers.0.rbr_3x3_branch.weight' not in dk['state_dict']
    psnr_full, _ = main_eval.main(flags + ['--eval_only'])
    psnr_pq, _ = main_eval.main(flags + ['--eval_only', '--prune_ratio', '0.2', '--quant_bit', '8'])
    assert psnr_full > 5.0 and psnr_pq > 5.0 and abs(psnr_full - psnr_pq) < 3.0
    assert (out / 'only_prune0.20_quant8.txt').exists() and (out / 'bpp_rank0.txt').exists()


@pytest.mark.parametrize("cfg,bt", [
    (dict(embed='1.25_40', stem_dim_num='64_1', fc_hw_dim='2_3_112', expansion=1, reduction=2, lower_width=96,
          strides=[5, 2]), "ERB"),                      # NeRV-L width (BASELINE configs[2]): 112 -> 2800 -> 112 -> 384
    (dict(embed='1.25_40', stem_dim_num='64_1', fc_hw_dim='2_3_26', expansion=1, reduction=2, lower_width=96,
          strides=[5, 3, 2]), "ERB"),                   # UVG stride list 5 3 2 (BASELINE configs[3]): N = 864 block
    (dict(embed='1.25_40', stem_dim_num='64_1', fc_hw_dim='2_3_26', expansion=1, reduction=2, lower_width=96,
          strides=[5, 3, 2]), "NeRV_vanilla"),
])
def test_generator_other_geometries_vs_oracle(dev, cfg, bt):
    """Channel widths / strides of the other BASELINE configs at a small spatial size: image, loss and every
    parameter gradient against the CPU oracle (fp32), same seed and target."""
    from orepnerv.utils import loss_fn
    pe, gen = build(cfg, bt, dev)
    sd = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    pos = torch.tensor([0.125, 0.625])
    img = gen(pe(pos))[0]
    H, W = img.shape[-2:]
    g = torch.Generator().manual_seed(21)
    target = torch.randint(0, 256, (2, 3, H, W), generator=g).float().div(255)
    loss = loss_fn(img, target.to(dev), argparse.Namespace(loss_type='Fusion6'))
    loss.backward()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    img_ref = O.generator_forward(params, O.pos_encoding(pos, 1.25, 40), ocfg(cfg))
    loss_ref = O.loss_fn(img_ref, target)
    grads_ref = torch.autograd.grad(loss_ref, list(params.values()))
    assert rel_l2(img, img_ref) <= 1e-2
    assert abs(loss.item() - loss_ref.item()) <= 2e-3
    for (k, _), gr in zip(params.items(), grads_ref):
        pg = dict(gen.named_parameters())[k].grad
        err = (pg.cpu() - gr).norm().item()
        assert err <= 4e-2 * gr.norm().item() + 1e-6, (k, err, gr.norm().item())


def test_host_pipeline_matches_plain_steps(dev, golden):
    """FrameFitter.step_host (prefetched H2D / deferred D2H on a copy stream) must produce exactly the metrics of
    plain FrameFitter.step on the same frames, one call later."""
    from orepnerv.trainer import FrameFitter
    g = golden("small_erb.pt")
    frames_u8 = (g['target'] * 255).round().to(torch.uint8)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5, beta=0.5,
                              batchSize=2)
    outs = []
    for mode in ("plain", "host"):
        pe, gen = build(g['cfg'], "ERB", dev)
        fit = FrameFitter(gen, pe, args, data_size=4, steps_per_epoch=2, use_graph=True, with_msssim=False)
        res = []
        if mode == "plain":
            for t in range(5):
                res.append(fit.step(frames_u8.to(dev), g['pos'].to(dev))[:5].cpu().clone())
        else:
            fp, tp = frames_u8.pin_memory(), g['pos'].clone().pin_memory()
            fit.host_pipeline_begin(fp, tp)
            for t in range(5):
                prev = fit.step_host(fp if t < 4 else None, tp if t < 4 else None)
                if prev is not None:
                    res.append(prev[:5].clone())
            res.append(fit.host_pipeline_end()[:5].clone())
        outs.append(torch.stack(res))
    assert outs[0].shape == outs[1].shape == (5, 5)
    # training kernels accumulate in a run-dependent order (atomics, concurrent issuers): compare within bf16 noise
    torch.testing.assert_close(outs[0], outs[1], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("B,H,W,C", [(1, 20, 36, 96), (1, 33, 44, 26), (2, 20, 36, 96), (1, 16, 200, 112)])
def test_head_kernels_vs_torch(dev, B, H, W, C):
    """RGB head (model.py:601, :620-623) forward / backward kernels against plain PyTorch fp32 on the same bf16
    activations.  B = 1 takes the bulk-copy streaming kernels (720 and 1452 pixels: ragged last 128-pixel tile),
    B = 2 the register-fed ones; the fused backward must agree with its split halves."""
    from orepnerv import _lib
    from orepnerv._lib import check, ptr
    lib = _lib.lib()
    g = torch.Generator().manual_seed(B * 1000 + C)
    Cp = (C + 31) // 32 * 32
    y = torch.zeros(B, H, W, Cp)
    y[..., :C] = torch.randn(B, H, W, C, generator=g)
    dsilu = torch.zeros(B, H, W, Cp)
    dsilu[..., :C] = torch.rand(B, H, W, C, generator=g) * 1.2 - 0.1
    y_d, ds_d = y.to(dev).bfloat16().contiguous(), dsilu.to(dev).bfloat16().contiguous()
    Wh = (torch.randn(3, C, generator=g) * 0.2).to(dev)
    bh = (torch.randn(3, generator=g) * 0.1).to(dev)
    gimg = torch.randn(B, 3, H, W, generator=g).to(dev)
    img = torch.zeros(B, 3, H, W, device=dev)
    st = _lib.stream()
    check(lib.onr_head_fwd(ptr(y_d), B, H, W, C, Cp, ptr(Wh), ptr(bh), 0, ptr(img), st), "head_fwd")
    yf = y_d.float()[..., :C].requires_grad_(True)
    Wr, br = Wh.clone().requires_grad_(True), bh.clone().requires_grad_(True)
    ref = (torch.tanh(torch.einsum('bhwc,kc->bkhw', yf, Wr) + br.view(1, 3, 1, 1)) + 1) * 0.5
    assert (img - ref).abs().max().item() <= 2e-5
    ref.backward(gimg)
    dz_ref = yf.grad * ds_d.float()[..., :C]

    def run(fused):
        gW, gb = torch.zeros(3, C, device=dev), torch.zeros(3, device=dev)
        dz = torch.zeros(B, H, W, Cp, device=dev, dtype=torch.bfloat16)
        if fused:
            check(lib.onr_head_bwd(ptr(gimg), ptr(img), ptr(y_d), ptr(ds_d), B, H, W, C, Cp, ptr(Wh), 0, ptr(gW),
                                   ptr(gb), ptr(dz), st), "head_bwd")
        else:
            check(lib.onr_head_bwd_dz(ptr(gimg), ptr(img), ptr(ds_d), B, H, W, C, Cp, ptr(Wh), 0, ptr(dz), st), "dz")
            check(lib.onr_head_bwd_gw(ptr(gimg), ptr(img), ptr(y_d), B, H, W, C, Cp, 0, ptr(gW), ptr(gb), st), "gw")
        return gW, gb, dz.float()

    for fused in (True, False):
        gW, gb, dz = run(fused)
        assert rel_l2(gW, Wr.grad) <= 1e-4, fused
        assert rel_l2(gb, br.grad) <= 1e-4, fused
        assert rel_l2(dz[..., :C], dz_ref) <= 6e-3, fused          # dz is stored as bf16
        assert dz[..., C:].abs().max().item() == 0.0 if Cp > C else True
    if B == 1 and (H * W) % 4 == 0:
        # z mode (last block of a training model stores only its pre-activation): SiLU / SiLU' evaluated in the kernels
        z_d = y_d                                            # reuse the same bf16 values as pre-activations
        zf = z_d.float()[..., :C].requires_grad_(True)
        Wz, bz = Wh.clone().requires_grad_(True), bh.clone().requires_grad_(True)
        act = torch.nn.functional.silu(zf)
        ref_z = (torch.tanh(torch.einsum('bhwc,kc->bkhw', act, Wz) + bz.view(1, 3, 1, 1)) + 1) * 0.5
        img_z = torch.zeros(B, 3, H, W, device=dev)
        check(lib.onr_head_fwd_z(ptr(z_d), B, H, W, C, Cp, ptr(Wh), ptr(bh), 0, ptr(img_z), st), "head_fwd_z")
        assert (img_z - ref_z).abs().max().item() <= 2e-3       # tanh.approx SiLU
        ref_z.backward(gimg)
        gW, gb = torch.zeros(3, C, device=dev), torch.zeros(3, device=dev)
        dz = torch.zeros(B, H, W, Cp, device=dev, dtype=torch.bfloat16)
        check(lib.onr_head_bwd_z(ptr(gimg), ptr(img_z), ptr(z_d), B, H, W, C, Cp, ptr(Wh), 0, ptr(gW), ptr(gb), ptr(dz),
                                 st), "head_bwd_z")
        assert rel_l2(gW, Wz.grad) <= 5e-3 and rel_l2(gb, bz.grad) <= 5e-3
        assert rel_l2(dz.float()[..., :C], zf.grad) <= 1e-2


def test_fold_ahead_matches_default(dev, golden, monkeypatch):
    """ONR_FOLD_AHEAD=1 (per-block Adam + next-step fold inside the backward) is a pure re-scheduling: same metrics
    and parameters as the default order after a few steps."""
    from orepnerv.trainer import FrameFitter
    g = golden("small_erb.pt")
    frames_u8 = (g['target'] * 255).round().to(torch.uint8)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5, beta=0.5,
                              batchSize=2)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ONR_FOLD_AHEAD", mode)
        pe, gen = build(g['cfg'], "ERB", dev)
        fit = FrameFitter(gen, pe, args, data_size=4, steps_per_epoch=2, use_graph=True, with_msssim=False)
        assert fit.fold_ahead == (mode == "1")
        outs = [fit.step(frames_u8.to(dev), g['pos'].to(dev))[:5].cpu().clone() for _ in range(4)]
        res[mode] = (torch.stack(outs), {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()})
    torch.testing.assert_close(res["0"][0], res["1"][0], rtol=2e-3, atol=2e-4)
    for k in res["0"][1]:
        assert rel_l2(res["1"][1][k], res["0"][1][k]) <= 2e-3, k


def test_decode_sees_parameters_updated_by_the_fitter(dev, golden):
    """The fitter updates parameters through raw pointers inside a CUDA graph; the decode-side packed-operand cache
    must notice (reference main_train.py evaluates the same model object between training epochs)."""
    import copy
    from orepnerv.trainer import FrameFitter
    g = golden("small_erb.pt")
    frames_u8 = (g['target'] * 255).round().to(torch.uint8)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-3, lr_type='cosine', warmup=0, epochs=5, beta=0.5,
                              batchSize=2)
    pe, gen = build(g['cfg'], "ERB", dev)
    embed = pe(g['pos'])
    with torch.no_grad():
        before = gen(embed)[0].clone()
    fit = FrameFitter(gen, pe, args, data_size=4, steps_per_epoch=2, use_graph=True, with_msssim=False)
    for _ in range(3):
        fit.step(frames_u8.to(dev), g['pos'].to(dev))
    with torch.no_grad():
        after = gen(embed)[0].clone()
        fresh = copy.deepcopy(gen)(embed)[0]          # new executor, packs from the current parameters
    assert not torch.equal(after, before)
    assert torch.equal(after, fresh)


@pytest.mark.parametrize("name", ["small_erb.pt", "tiny_vanilla.pt"])
def test_decode_fused_head_matches_unfused(dev, golden, monkeypatch, name):
    """Decode with the RGB head fused into the last block's epilogue (ONR_CONV_FPROP_HEAD, ONR_DECODE_FUSED=1) against the
    unfused pair of kernels and against the oracle: same image up to the bf16 rounding of the last activation that the
    fused path no longer performs."""
    g = golden(name)
    bt = "ERB" if "erb" in name else "NeRV_vanilla"
    imgs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("ONR_DECODE_FUSED", fused)
        pe, gen = build(g['cfg'], bt, dev)
        with torch.no_grad():
            imgs[fused] = gen(pe(g['pos']))[0].clone()
        ex = gen.executor(g['pos'].numel(), False)
        assert ex._decode_fused == (fused == "1")
        with torch.no_grad():
            again = gen(pe(g['pos']))[0]
        assert torch.equal(again, imgs[fused])                   # decode is deterministic
    assert rel_l2(imgs["1"], imgs["0"]) <= 3e-3
    assert rel_l2(imgs["1"], g['img']) <= 1e-2


def test_a13_prune_quant_against_reference_golden(dev, golden):
    """SURVEY 8a-A13 known-answer test.  The golden file was produced by the UNMODIFIED reference
    (tests/golden/make_golden.py:a13_prune_quant): deploy model -> prune.global_unstructured 0.2 -> quantize_per_tensor
    over every state-dict entry (weight_orig AND weight_mask) -> load_state_dict -> decode.  Our eval driver stages
    (main_eval.global_prune, prune_and_quantise) must give the same mask count, bit-identical quantised tensors — the
    masks come back as all ones, i.e. the reference's quantisation undoes its pruning, a quirk kept on purpose — and
    the same decoded image."""
    from orepnerv import main_eval
    g = golden("a13_prune_quant.pt")
    pe, dep = build(g['cfg'], "ERB", dev, deploy=True)
    dep.load_state_dict(g['deploy_state'])
    args = argparse.Namespace(prune_ratio=0.2, quant_bit=8, quant_axis=0)
    zeros, total = main_eval.global_prune(dep, 0.2)
    assert (zeros, total) == (g['mask_zeros'], g['mask_total'])
    for k, v in dep.state_dict().items():
        assert torch.equal(v.cpu(), g['pruned_state'][k]), k            # same masks, same weight_orig
    args.prune_ratio = 1.0                                               # already pruned: quantise only
    main_eval.prune_and_quantise(dep, args, 2, (18, 24))
    sd = dep.state_dict()
    assert set(sd) == set(g['quant_state'])
    for k, v in sd.items():
        assert torch.equal(v.cpu(), g['quant_state'][k]), k
        if k.endswith('weight_mask'):
            assert bool((v == 1).all()), k                               # the quirk: pruning is undone
    with torch.no_grad():
        img = dep(pe(g['pos']))[0]
    assert rel_l2(img, g['img']) <= 1e-2


def test_global_prune_hits_exactly_k_under_ties(dev, golden):
    """Coarsely quantised weights have many equal magnitudes: exactly round(amount * N) entries must be masked, as
    torch.topk does in the reference (main_eval.py:587), not every entry tied with the threshold."""
    from orepnerv import main_eval
    g = golden("small_erb.pt")
    pe, dep = build(g['cfg'], "ERB", dev, deploy=True)
    dep.load_state_dict(g['deploy_state'])
    with torch.no_grad():
        for m in main_eval.prunable_modules(dep):
            m.weight.copy_((m.weight * 20).round() / 20)                 # values k/20: heavy ties, many exact zeros
    zeros, total = main_eval.global_prune(dep, 0.3)
    assert zeros == int(round(0.3 * total))
