"""CPU tests: the oracle (oracle/nerv_oracle.py) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  These pin the oracle; the GPU tests then compare the CUDA path
with the oracle.  Tolerances: fp32 round-off only (the oracle runs the same math in a different op order)."""
import math

import pytest
import torch

from oracle import nerv_oracle as O


def cfg_of(g):
    c = g['cfg']
    fh, fw, fd = [int(x) for x in c['fc_hw_dim'].split('_')]
    return dict(fc_h=fh, fc_w=fw, fc_dim=fd, strides=c['strides'], sigmoid=False)


def pe_params(g):
    lbase, levels = g['cfg']['embed'].split('_')
    return float(lbase), int(levels)


@pytest.mark.parametrize("name", ["tiny_erb.pt", "tiny_vanilla.pt", "small_erb.pt"])
def test_forward_loss_grads(golden, name):
    g = golden(name)
    lbase, levels = pe_params(g)
    embed = O.pos_encoding(g['pos'], lbase, levels)
    assert torch.equal(embed, g['embed'])                      # same ops, same order: bit exact
    sd = g['init_state']
    img = O.generator_forward(sd, embed, cfg_of(g))
    torch.testing.assert_close(img, g['img'], rtol=1e-5, atol=1e-6)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = O.loss_fn(O.generator_forward(params, embed, cfg_of(g)), g['target'])
    torch.testing.assert_close(loss.detach(), g['loss'], rtol=1e-5, atol=1e-6)
    grads = torch.autograd.grad(loss, list(params.values()))
    for (k, _), gr in zip(params.items(), grads):
        torch.testing.assert_close(gr, g['grads'][k], rtol=2e-4, atol=1e-7, msg=lambda m: f"{k}: {m}")
    torch.testing.assert_close(O.psnr(img, g['target']).view(1, 1).expand(2, 1), g['psnr'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["tiny_erb.pt", "small_erb.pt"])
def test_fold_and_deploy(golden, name):
    g = golden(name)
    sd = g['init_state']
    for i, (K_ref, b_ref) in enumerate(g['folded']):
        K, b = O.block_kernel(sd, f'layers.{i}.')
        assert (K - K_ref).norm() / K_ref.norm() < 1e-6        # north-star gate is 1e-5
        assert (b - b_ref).norm() / b_ref.norm() < 1e-6
        torch.testing.assert_close(g['deploy_state'][f'layers.{i}.rbr_reparam.weight'], K_ref)
    # deploy-state forward == train-state forward
    lbase, levels = pe_params(g)
    img_d = O.generator_forward(g['deploy_state'], O.pos_encoding(g['pos'], lbase, levels), cfg_of(g))
    torch.testing.assert_close(img_d, g['deploy_img'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(g['deploy_img'], g['img'], rtol=1e-5, atol=1e-6)


def test_fold_backward_analytic():
    torch.manual_seed(0)
    cin, cout = 5, 12
    ws = [torch.randn(cout, cin, 3, 3), torch.randn(cout), torch.randn(cout, cin, 1, 3), torch.randn(cout),
          torch.randn(cout, cin, 3, 1), torch.randn(cout), torch.randn(2 * cin, cin, 1, 1),
          torch.randn(cout, 2 * cin, 3, 3), torch.randn(cout, cout, 1, 1)]
    ws = [w.double().requires_grad_(True) for w in ws]
    K, b = O.erb_fold(*ws)
    dK, db = torch.randn_like(K), torch.randn_like(b)
    auto = torch.autograd.grad([K, b], ws, [dK, db])
    ana = O.erb_fold_backward(dK, db, ws[6].detach(), ws[7].detach(), ws[8].detach())
    order = ['w3x3', 'b3x3', 'w1x3', 'b1x3', 'w3x1', 'b3x1', 'w1', 'w2', 'w3']
    for name, a in zip(order, auto):
        torch.testing.assert_close(ana[name], a, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", ["tiny_erb.pt", "tiny_vanilla.pt"])
def test_training_steps(golden, name):
    g = golden(name)
    lbase, levels = pe_params(g)
    embed = O.pos_encoding(g['pos'], lbase, levels)
    sd, state = {k: v.clone() for k, v in g['init_state'].items()}, {}
    for i, (loss_ref, lr_ref) in enumerate(zip(g['train_losses'], g['train_lrs'])):
        lr = O.lr_at(0, i, 4, 5e-4, 1, 5)
        assert abs(lr - lr_ref) < 1e-12
        sd, state, loss, _, _ = O.train_step(sd, state, embed, g['target'], cfg_of(g), lr, i + 1)
        assert abs(loss.item() - loss_ref) < 2e-5
    for k, v in g['trained_state'].items():
        torch.testing.assert_close(sd[k], v, rtol=1e-3, atol=2e-5, msg=lambda m: f"{k}: {m}")


def test_misc(golden):
    m = golden('misc.pt')
    assert torch.equal(O.pos_encoding(m['pe_pos'], 1.25, 40), m['pe_embed'])
    qi, qo = m['quant_in'], m['quant_out']
    for key, (t, bit, axis) in {'t4_axis0': (qi['t4'], 8, 0), 't4_axis1': (qi['t4'], 8, 1),
                                't2_axis0': (qi['t2'], 8, 0), 't2_axis-1': (qi['t2'], 6, -1),
                                't1_axis-1': (qi['t1'], 8, -1)}.items():
        q, new = O.quantize_per_tensor(t, bit, axis)
        assert torch.equal(q, qo[key][0]), key
        assert torch.equal(new, qo[key][1]), key
    for epoch, it, lr_ref in m['lr_sched']:
        assert abs(O.lr_at(epoch, it, 132, 5e-4, 60, 300) - lr_ref) < 1e-15
    a, b = m['metric_in']
    torch.testing.assert_close(O.psnr(a, b).view(1, 1), m['metric_psnr'])
    torch.testing.assert_close(O.ms_ssim(a, b).view(1, 1), m['metric_msssim'])


def test_ssim_against_independent_f64():
    """pytorch_msssim is not vendored by the reference (parity unpinned): cross-check the separable fp32
    restatement against a direct 11x11-window float64 evaluation."""
    g = torch.Generator().manual_seed(5)
    a = torch.rand(2, 3, 40, 56, generator=g)
    b = (a + 0.1 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    assert abs(O.ssim(a, b).item() - O.ssim_direct_f64(a, b).item()) < 2e-6
    assert abs(O.ssim(a, a).item() - 1.0) < 1e-6


def test_prune_threshold_matches_torch_prune():
    import torch.nn as nn
    import torch.nn.utils.prune as prune
    torch.manual_seed(2)
    mods = [nn.Linear(7, 9), nn.Conv2d(3, 5, 3)]
    thr, k = O.prune_threshold([m.weight.detach() for m in mods], 0.3)
    prune.global_unstructured([(m, 'weight') for m in mods], pruning_method=prune.L1Unstructured, amount=0.3)
    for m in mods:
        assert torch.equal(m.weight_mask, (m.weight_orig.abs() > thr).float())
    assert sum(int((m.weight_mask == 0).sum()) for m in mods) == k


def test_ssim_shim_is_a_second_independent_restatement():
    """The golden vectors that pass through pytorch_msssim (loss, MS-SSIM metric) are produced with
    tests/golden/_shim/pytorch_msssim.py, written after the structure of the published 0.2.1 package and WITHOUT the
    oracle: two separate restatements must agree (the real package is absent: parity with it stays unpinned)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "_shim", "pytorch_msssim.py")
    with open(path) as f:
        src = f.read()
    assert "import" in src and "from oracle" not in src and "import oracle" not in src and "nerv_oracle as" not in src
    spec = importlib.util.spec_from_file_location("pytorch_msssim_shim", path)
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    g = torch.Generator().manual_seed(6)
    for shape in [(2, 3, 40, 56), (1, 3, 177, 201)]:
        a = torch.rand(shape, generator=g)
        b = (a + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
        assert abs(shim.ssim(a, b, data_range=1, size_average=True).item() - O.ssim(a, b).item()) < 1e-6
        ar = a.clone().requires_grad_(True)
        ao = a.clone().requires_grad_(True)
        (1 - shim.ssim(ar, b, data_range=1, size_average=True)).backward()
        (1 - O.ssim(ao, b)).backward()
        torch.testing.assert_close(ar.grad, ao.grad, rtol=1e-4, atol=1e-9)
    a = torch.rand(1, 3, 177, 201, generator=g)             # odd sizes: avg_pool2d padding = dim % 2
    b = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    assert abs(shim.ms_ssim(a, b, data_range=1, size_average=True).item() - O.ms_ssim(a, b).item()) < 1e-6
