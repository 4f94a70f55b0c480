"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/model.py, utils.py) on
the CPU of the build container.  The GPU box has no /root/reference: tests only read the committed .pt
files.  Run:  python tests/golden/make_golden.py
"""
import argparse
import os
import sys
from copy import deepcopy

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))     # pytorch_msssim stand-in (see its docstring)
sys.path.insert(0, "/root/reference")
import model as ref_model      # noqa: E402  (reference, unmodified)
import utils as ref_utils      # noqa: E402  (reference, unmodified)

TINY = dict(embed='1.25_4', stem_dim_num='16_1', fc_hw_dim='3_4_4', expansion=1, reduction=2, lower_width=4,
            strides=[2, 2])
SMALL = dict(embed='1.25_40', stem_dim_num='64_1', fc_hw_dim='3_4_12', expansion=1, reduction=2, lower_width=8,
             strides=[3, 2])


def build(cfg, branch_type, deploy=False, seed=1):
    torch.manual_seed(seed)
    pe = ref_utils.PositionalEncoding(cfg['embed'])
    gen = ref_model.Generator(embed_length=pe.embed_length, stem_dim_num=cfg['stem_dim_num'],
                              fc_hw_dim=cfg['fc_hw_dim'], expansion=cfg['expansion'], num_blocks=1, norm='none',
                              act='swish', bias=True, reduction=cfg['reduction'], conv_type='conv',
                              stride_list=cfg['strides'], sin_res=True, lower_width=cfg['lower_width'],
                              sigmoid=False, deploy=deploy, branch_type=branch_type)
    return pe, gen


def frames(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, 3, h, w), generator=g).float().div(255)


def case(cfg, branch_type, name, n_steps=3):
    pe, gen = build(cfg, branch_type)
    out = {'cfg': cfg, 'branch_type': branch_type}
    out['init_state'] = {k: v.clone() for k, v in gen.state_dict().items()}
    pos = torch.tensor([0.25, 0.7])
    embed = pe(pos)
    out['pos'], out['embed'] = pos, embed.clone()
    img = gen(embed)[0]
    out['img'] = img.detach().clone()
    H, W = img.shape[-2:]
    target = frames(2, H, W, 7)
    out['target'] = target
    if branch_type == 'ERB':
        out['folded'] = [tuple(t.detach().clone() for t in blk.get_equivalent_kernel_bias()) for blk in gen.layers]
        dep = deepcopy(gen)
        for blk in dep.layers:
            blk.switch_to_deploy()
        out['deploy_state'] = {k: v.clone() for k, v in dep.state_dict().items()}
        out['deploy_img'] = dep(embed)[0].detach().clone()
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5)
    loss = ref_utils.loss_fn(img, target, args)
    out['loss'] = loss.detach().clone()
    gen.zero_grad()
    loss.backward()
    out['grads'] = {k: p.grad.detach().clone() for k, p in gen.named_parameters() if p.grad is not None}
    out['psnr'] = ref_utils.psnr_fn([img.detach()], [target]).clone()
    # a few optimisation steps exactly like reference main_train.py:238-250
    opt = torch.optim.Adam(gen.parameters(), betas=(0.5, 0.999))
    gen.zero_grad()
    losses, lrs = [], []
    for i in range(n_steps):
        o = gen(embed)[0]
        l = ref_utils.loss_fn(o, target, args)
        lrs.append(ref_utils.adjust_lr(opt, 0, i, 4, args))
        opt.zero_grad()
        l.backward()
        opt.step()
        losses.append(l.item())
    out['train_losses'], out['train_lrs'] = losses, lrs
    out['trained_state'] = {k: v.clone() for k, v in gen.state_dict().items()}
    torch.save(out, os.path.join(HERE, name))
    print(name, 'img', tuple(img.shape), 'loss', float(loss.detach()), 'params', sum(p.numel() for p in gen.parameters()))


def misc():
    out = {}
    pe = ref_utils.PositionalEncoding('1.25_40')
    pos = torch.tensor([0.0, 1 / 132, 0.5, 131 / 132, 599 / 600], dtype=torch.float32)
    out['pe_pos'], out['pe_embed'] = pos, pe(pos)
    g = torch.Generator().manual_seed(3)
    t4 = torch.randn(6, 5, 3, 3, generator=g)
    t4[t4.abs() < 0.3] = 0
    t4[2] = 0                       # an all-zero row
    t2 = torch.randn(7, 9, generator=g)
    t1 = torch.randn(11, generator=g)
    out['quant_in'] = {'t4': t4, 't2': t2, 't1': t1}
    out['quant_out'] = {
        't4_axis0': ref_utils.quantize_per_tensor(t4, 8, 0), 't4_axis1': ref_utils.quantize_per_tensor(t4, 8, 1),
        't2_axis0': ref_utils.quantize_per_tensor(t2, 8, 0), 't2_axis-1': ref_utils.quantize_per_tensor(t2, 6, -1),
        't1_axis-1': ref_utils.quantize_per_tensor(t1, 8, -1),
    }
    args = argparse.Namespace(lr=5e-4, lr_type='cosine', warmup=60, epochs=300)

    class _Opt:
        param_groups = [{'lr': 0.0}]
    sched = []
    for epoch, it in [(0, 0), (0, 66), (30, 5), (59, 131), (60, 0), (150, 17), (299, 131)]:
        sched.append((epoch, it, ref_utils.adjust_lr(_Opt(), epoch, it, 132, args)))
    out['lr_sched'] = sched
    # SSIM / MS-SSIM / PSNR metric values through the reference's own call sites (pytorch_msssim is the shim)
    a = frames(1, 176, 192, 11)
    b = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    out['metric_in'] = (a, b)
    out['metric_psnr'] = ref_utils.psnr_fn([a], [b])
    out['metric_msssim'] = ref_utils.msssim_fn([a], [b])
    torch.save(out, os.path.join(HERE, 'misc.pt'))
    print('misc.pt written')


def a13_prune_quant():
    """SURVEY 8a-A13 known-answer test with the reference's own calls: deploy-state model ->
    prune.global_unstructured(stem Linears + rbr_reparam convs, L1Unstructured, 0.2) (main_eval.py:572-587) ->
    quantize_per_tensor over EVERY state-dict entry, weight_orig and weight_mask included (main_eval.py:660-669) ->
    load_state_dict (:703) -> decode.  Stores the deploy state, the quantised state dict and the decoded image."""
    import torch.nn.utils.prune as prune
    g = torch.load(os.path.join(HERE, 'small_erb.pt'), weights_only=False)
    pe, dep = build(SMALL, 'ERB', deploy=True)
    dep.load_state_dict(g['deploy_state'])
    mods = [dep.stem[0], dep.stem[2]] + [blk.rbr_reparam for blk in dep.layers]
    prune.global_unstructured([(m, 'weight') for m in mods], pruning_method=prune.L1Unstructured, amount=0.2)
    zeros = sum(int((m.weight_mask == 0).sum()) for m in mods)
    total = sum(m.weight_mask.numel() for m in mods)
    with torch.no_grad():
        cur = dep.state_dict()
        pruned_state = {k: v.clone() for k, v in cur.items()}
        for k, v in cur.items():
            large = v.dim() in {2, 4} and 'bias' not in k
            _, new_v = ref_utils.quantize_per_tensor(v, 8, 0 if large else -1)
            cur[k] = new_v.detach().type_as(v)
        dep.load_state_dict(cur)
        img = dep(pe(g['pos']))[0].detach().clone()
    out = {'cfg': SMALL, 'pos': g['pos'], 'deploy_state': g['deploy_state'], 'mask_zeros': zeros, 'mask_total': total,
           'pruned_state': pruned_state, 'quant_state': {k: v.clone() for k, v in cur.items()}, 'img': img}
    torch.save(out, os.path.join(HERE, 'a13_prune_quant.pt'))
    ones = {k: float((v == 1).float().mean()) for k, v in cur.items() if k.endswith('weight_mask')}
    print('a13_prune_quant.pt: mask zeros', zeros, '/', total, '; fraction of ones in the quantised masks', ones)


def finetune_case(branch_type, src, name, amount=0.3, start_epoch=3, finetune_epochs=2, iters=2):
    """Prune-then-finetune (reference main_eval.py:213-545) with the reference's own classes and torch.nn.utils.prune:
    a line-for-line mirror of the loop (:239-368 prune list + global_unstructured, :426 fresh Adam, :446-507 the steps
    with `adjust_lr(epoch % total_epochs)`, `backward(retain_graph=ERB)` and `optimizer.state.clear()` before the first
    step, :530-541 switch_to_deploy) minus `.cuda()` and the data loader.  The ERB run exhibits the frozen-branch quirk
    (SURVEY.md 2.1 row 19): the golden records that the pruned branch kernels do not move."""
    import torch.nn.utils.prune as prune
    g = torch.load(os.path.join(HERE, src), weights_only=False)
    cfg = g['cfg']
    pe, gen = build(cfg, branch_type)
    gen.load_state_dict(g['trained_state'])
    start_state = {k: v.clone() for k, v in gen.state_dict().items()}
    param_list, names = [], []
    if branch_type == 'NeRV_vanilla':
        for k, v in gen.named_parameters():
            if 'weight' in k:
                if 'stem' in k:
                    param_list.append(gen.stem[int(k.split('.')[1])]); names.append(k)
                elif 'layers' in k[:6]:
                    param_list.append(gen.layers[int(k.split('.')[1])].branch); names.append(k)
    else:
        for k, v in gen.named_parameters():
            if 'weight' in k and 'stem' in k:
                param_list.append(gen.stem[int(k.split('.')[1])]); names.append(k)
        for li, layer in enumerate(gen.layers):
            for bn in ('rbr_3x3_branch', 'rbr_3x1_branch', 'rbr_1x3_branch', 'rbr_1x1_3x3_1x1_branch_1x1_1',
                       'rbr_1x1_3x3_1x1_branch_3x3', 'rbr_1x1_3x3_1x1_branch_1x1_2'):
                if hasattr(layer, bn):
                    param_list.append(getattr(layer, bn)); names.append(f'layers.{li}.{bn}.weight')
    prune.global_unstructured([(m, 'weight') for m in param_list], pruning_method=prune.L1Unstructured, amount=amount)
    masks = {n: m.weight_mask.detach().clone() for n, m in zip(names, param_list)}
    opt = torch.optim.Adam(gen.parameters(), betas=(0.5, 0.999), foreach=False)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5)
    embed, target = g['embed'], g['target']
    total_epochs = start_epoch + finetune_epochs
    losses, lrs = [], []
    for epoch in range(start_epoch, total_epochs):
        gen.train()
        for i in range(iters):
            out = gen(embed)[0]
            loss = ref_utils.loss_fn(out, target, args)
            lrs.append(ref_utils.adjust_lr(opt, epoch % total_epochs, i, 4, args))
            opt.zero_grad()
            loss.backward(retain_graph=(branch_type == 'ERB'))
            if epoch == start_epoch and i == 0:
                opt.state.clear()
            opt.step()
            losses.append(loss.item())
    pre_deploy = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    eff = {n: m.weight.detach().clone() for n, m in zip(names, param_list)}      # what the forward reads
    if branch_type == 'ERB':
        for layer in gen.layers:
            layer.switch_to_deploy()
    final = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    with torch.no_grad():
        img = gen(embed)[0].detach().clone()
    out = dict(cfg=cfg, branch_type=branch_type, start_state=start_state, masks=masks, mask_names=names, amount=amount,
               start_epoch=start_epoch, finetune_epochs=finetune_epochs, iters=iters, data_size=4, losses=losses, lrs=lrs,
               pre_deploy_state=pre_deploy, effective_weights=eff, final_state=final, embed=embed, pos=g['pos'],
               target=target, img=img)
    torch.save(out, os.path.join(HERE, name))
    moved = {n: float((eff[n] - start_state[n] * masks[n]).abs().max()) for n in names}
    print(name, 'losses', losses, 'lrs', lrs, 'max |effective weight - pruned start|', moved)


MULTI = dict(embed='1.25_40', stem_dim_num='64_1', fc_hw_dim='4_6_12', expansion=1, reduction=2, lower_width=8,
             strides=[3, 2])


def multires_case(name, lw=0.7, n_steps=3):
    """Multi-resolution heads (sin_res=False, reference model.py:598-608, :615-623) through the reference's own
    training iteration (main_train.py:238-250): one RGB head per stage, the frame pooled to every head's resolution,
    per-stage Fusion6 losses summed with weight --lw on all but the last."""
    import torch.nn.functional as F
    cfg = MULTI
    torch.manual_seed(1)
    pe = ref_utils.PositionalEncoding(cfg['embed'])
    gen = ref_model.Generator(embed_length=pe.embed_length, stem_dim_num=cfg['stem_dim_num'], fc_hw_dim=cfg['fc_hw_dim'],
                              expansion=cfg['expansion'], num_blocks=1, norm='none', act='swish', bias=True,
                              reduction=cfg['reduction'], conv_type='conv', stride_list=cfg['strides'], sin_res=False,
                              lower_width=cfg['lower_width'], sigmoid=False, deploy=False, branch_type='ERB')
    out = {'cfg': cfg, 'lw': lw, 'init_state': {k: v.clone() for k, v in gen.state_dict().items()}}
    pos = torch.tensor([0.25, 0.7])
    embed = pe(pos)
    out['pos'], out['embed'] = pos, embed.clone()
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5, lw=lw)

    def step_loss(data):
        output_list = gen(embed)
        target_list = [F.adaptive_avg_pool2d(data, x.shape[-2:]) for x in output_list]
        loss_list = [ref_utils.loss_fn(o, t, args) for o, t in zip(output_list, target_list)]
        loss_list = [loss_list[i] * (args.lw if i < len(loss_list) - 1 else 1) for i in range(len(loss_list))]
        return output_list, target_list, loss_list, sum(loss_list)

    imgs = gen(embed)
    H, W = imgs[-1].shape[-2:]
    data = frames(2, H, W, 7)
    out['target'] = data
    output_list, target_list, loss_list, loss_sum = step_loss(data)
    out['imgs'] = [o.detach().clone() for o in output_list]
    out['targets'] = [t.clone() for t in target_list]
    out['losses'] = [l.detach().clone() for l in loss_list]
    out['loss_sum'] = loss_sum.detach().clone()
    gen.zero_grad()
    loss_sum.backward()
    out['grads'] = {k: p.grad.detach().clone() for k, p in gen.named_parameters()}
    out['psnr'] = ref_utils.psnr_fn([o.detach() for o in output_list], target_list).clone()
    opt = torch.optim.Adam(gen.parameters(), betas=(0.5, 0.999))
    losses = []
    for i in range(n_steps):
        _, _, _, l = step_loss(data)
        ref_utils.adjust_lr(opt, 0, i, 4, args)
        opt.zero_grad()
        l.backward()
        opt.step()
        losses.append(l.item())
    out['train_losses'] = losses
    out['trained_state'] = {k: v.clone() for k, v in gen.state_dict().items()}
    torch.save(out, os.path.join(HERE, name))
    print(name, [tuple(o.shape) for o in output_list], 'loss', float(loss_sum.detach()), losses)


if __name__ == '__main__':
    what = sys.argv[1:] or ['base']
    if 'multires' in what:
        multires_case('small_erb_multires.pt')
    if 'base' in what:
        case(TINY, 'ERB', 'tiny_erb.pt')
        case(TINY, 'NeRV_vanilla', 'tiny_vanilla.pt')
        case(SMALL, 'ERB', 'small_erb.pt')
        misc()
        a13_prune_quant()
    if 'finetune' in what:
        finetune_case('ERB', 'small_erb.pt', 'finetune_erb.pt')
        finetune_case('NeRV_vanilla', 'tiny_vanilla.pt', 'finetune_vanilla.pt')
    if 'branches' in what:
        for bt in ('ACB', 'RepVGG', 'DBB', 'ECB'):
            case(SMALL, bt, f'small_{bt.lower()}.pt')
