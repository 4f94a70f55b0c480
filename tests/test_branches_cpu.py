"""CPU tests for SURVEY.md 8f-3 / 8f-4 (prune-then-finetune, the other branch sets, the activation table):
  * the oracle's restatement of ACB / RepVGG / DBB / ECB against golden vectors of the UNMODIFIED reference
    (tests/golden/make_golden.py branches): forward, loss, every parameter gradient, three optimisation steps;
  * the element arithmetic of csrc/fold_branches.cuh — the functions the CUDA kernels run per thread — compiled with
    g++ (tests/native/fold_branches_host.cpp) against the oracle's fold and autograd, so the index math is verified
    before any GPU time is spent;
  * csrc/act.cuh (value / derivative of every activation) compiled the same way against torch;
  * the oracle's prune-then-finetune against the golden produced by the reference's own loop + torch.nn.utils.prune
    (including the frozen-branch quirk of the ERB path).
"""
import ctypes as C
import os
import subprocess

import pytest
import torch
import torch.nn.functional as F

from oracle import nerv_oracle as O

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "boosting-neural-video-representation-via-online-structural-reparameteration_b200")
BRANCH_GOLDENS = [("small_acb.pt", "ACB"), ("small_repvgg.pt", "RepVGG"), ("small_dbb.pt", "DBB"), ("small_ecb.pt", "ECB")]


def cfg_of(g, act='swish'):
    c = g['cfg']
    fh, fw, fd = [int(x) for x in c['fc_hw_dim'].split('_')]
    return dict(fc_h=fh, fc_w=fw, fc_dim=fd, strides=c['strides'], sigmoid=False, act=act)


# ------------------------------------------------------------------------------------------- oracle vs reference golden
@pytest.mark.parametrize("name,bt", BRANCH_GOLDENS)
def test_oracle_branch_sets_match_reference(golden, name, bt):
    g = golden(name)
    sd, embed = g['init_state'], g['embed']
    img = O.generator_forward(sd, embed, cfg_of(g))
    torch.testing.assert_close(img, g['img'], rtol=1e-5, atol=1e-6)        # folded conv == explicit multi-branch forward
    params = {k: (v.clone() if k.endswith('.mask') else v.clone().requires_grad_(True)) for k, v in sd.items()}
    loss = O.loss_fn(O.generator_forward(params, embed, cfg_of(g)), g['target'])
    torch.testing.assert_close(loss.detach(), g['loss'], rtol=1e-5, atol=1e-6)
    names = [k for k in params if params[k].requires_grad]
    assert set(names) == set(g['grads'])
    for k, gr in zip(names, torch.autograd.grad(loss, [params[k] for k in names])):
        torch.testing.assert_close(gr, g['grads'][k], rtol=5e-4, atol=2e-7, msg=lambda m: f"{k}: {m}")
    # explicit restatement of model.py:541-565 == the fold, block by block, in float64
    sd64 = {k: v.double() for k, v in sd.items()}
    x = torch.randn(2, cfg_of(g)['fc_dim'], 5, 7, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    K, b = O.block_kernel(sd64, 'layers.0.')
    assert (F.conv2d(x, K, b, padding=1) - O.branch_set_forward(x, sd64, 'layers.0.', bt)).abs().max() < 1e-12
    # three optimisation steps (main_train.py:238-250 as run by the golden generator)
    s, state = {k: v.clone() for k, v in sd.items()}, {}
    for i, loss_ref in enumerate(g['train_losses']):
        s, state, l, _, _ = O.train_step(s, state, embed, g['target'], cfg_of(g), O.lr_at(0, i, 4, 5e-4, 1, 5), i + 1)
        assert abs(l.item() - loss_ref) < 2e-5
    assert list(s) == list(g['trained_state'])
    for k, v in g['trained_state'].items():
        torch.testing.assert_close(s[k], v, rtol=1e-3, atol=2e-5, msg=lambda m: f"{k}: {m}")


# ------------------------------------------------------------------------------------------- fold arithmetic of the kernels
@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libfoldhost.so")
    src = os.path.join(ROOT, "tests", "native", "fold_branches_host.cpp")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


@pytest.mark.parametrize("bt", ["ACB", "RepVGG", "DBB", "ECB"])
@pytest.mark.parametrize("cin,cout", [(3, 8), (12, 108)])
def test_branch_fold_kernel_arithmetic(hostlib, bt, cin, cout):
    from orepnerv import branches
    from orepnerv.model import NeRVBlock
    torch.manual_seed(11)
    blk = NeRVBlock(ngf=cin, new_ngf=cout, stride=1, bias=True, norm='none', act='swish', deploy=False,
                    conv_type='conv', branch_type=bt)
    with torch.no_grad():                                   # scale / bias of SeqConv3x3 are ~1e-3 at init: make them count
        for n, p in blk.named_parameters():
            if n.endswith('.scale') or n.endswith('.b0'):
                p.copy_(torch.randn_like(p))
    assert blk.fold_kind() == "set"
    slots = branches.branch_slots(blk)
    s = branches.make_branch_set(cin, cout, {slot: t.detach() for slot, _, t in slots})
    K, b = torch.empty(cout, cin, 3, 3), torch.empty(cout)
    hostlib.host_branch_fold_fwd(C.byref(s), C.c_void_p(K.data_ptr()), C.c_void_p(b.data_ptr()))
    sd64 = {'L.' + n: t.detach().double().requires_grad_(not n.endswith('.mask')) for _, n, t in slots}
    K_ref, b_ref = O.branch_set_fold(sd64, 'L.')
    assert ((K - K_ref).norm() / K_ref.norm()).item() < 1e-6 and ((b - b_ref).norm() / b_ref.norm()).item() < 1e-6
    # backward: dK, dbias -> every branch gradient, against autograd through the oracle's fold
    gen = torch.Generator().manual_seed(12)
    dK, db = torch.randn(K.shape, generator=gen), torch.randn(b.shape, generator=gen)
    names = [n for _, n, _ in slots if not n.endswith('.mask')]
    ref = torch.autograd.grad([K_ref, b_ref], [sd64['L.' + n] for n in names], [dK.double(), db.double()])
    grads = {n: torch.full_like(t, float('nan')) for _, n, t in slots if not n.endswith('.mask')}
    gset = branches.make_branch_set(cin, cout, {slot: grads.get(n) for slot, n, _ in slots})
    hostlib.host_branch_fold_bwd(C.byref(s), C.c_void_p(dK.data_ptr()), C.c_void_p(db.data_ptr()), C.byref(gset))
    for n, r in zip(names, ref):
        assert torch.isfinite(grads[n]).all(), n            # every element written
        assert ((grads[n] - r).norm() / (r.norm() + 1e-30)).item() < 2e-6, n


def test_activation_table_matches_torch(tmp_path):
    src = tmp_path / "act_host.cpp"
    src.write_text('#include "%s/csrc/act.cuh"\n'
                   'extern "C" void host_act(const float* z, int n, int act, float* y, float* d) {\n'
                   '    for (int i = 0; i < n; ++i) onr::act_value_grad(z[i], act, y + i, d + i);\n}\n' % PKG)
    out = str(tmp_path / "libacthost.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", out, str(src)], check=True)
    lib = C.CDLL(out)
    from orepnerv.model import ACT_CODES
    z = torch.cat([torch.linspace(-30, 30, 4001), torch.tensor([-3.0, 3.0, 0.0, 6.0, 20.0, 20.5])]).float()
    for name, code in ACT_CODES.items():
        y, d = torch.empty_like(z), torch.empty_like(z)
        lib.host_act(C.c_void_p(z.data_ptr()), z.numel(), code, C.c_void_p(y.data_ptr()), C.c_void_p(d.data_ptr()))
        zr = z.clone().double().requires_grad_(True)
        yr = O.activation(zr, name)
        dr, = torch.autograd.grad(yr.sum(), zr)
        torch.testing.assert_close(y.double(), yr.detach(), rtol=2e-6, atol=2e-6, msg=lambda m: f"{name}: {m}")
        torch.testing.assert_close(d.double(), dr, rtol=2e-6, atol=2e-6, msg=lambda m: f"{name}' : {m}")


# ------------------------------------------------------------------------------------------- prune-then-finetune
@pytest.mark.parametrize("name", ["finetune_erb.pt", "finetune_vanilla.pt"])
def test_oracle_finetune_matches_reference(golden, name):
    g = golden(name)
    erb = g['branch_type'] == 'ERB'
    masks = g['masks']
    frozen = {n for n in masks if '.rbr_' in n} if erb else set()
    sd, eff, losses = O.finetune_steps(g['start_state'], masks, frozen, g['embed'], g['target'], cfg_of(g), g['lrs'])
    for a, b in zip(losses, g['losses']):
        assert abs(a - b) < 2e-5
    for n in masks:
        if n in frozen:                                                                  # the quirk: frozen
            assert torch.equal(eff[n], g['start_state'][n] * masks[n])
            assert torch.equal(eff[n], g['effective_weights'][n])
        assert torch.equal(eff[n] * (1 - masks[n]), torch.zeros_like(eff[n]))            # pruned entries stay pruned
    pre = g['pre_deploy_state']
    for k, v in sd.items():
        ref = pre[k] if k in pre else pre[k + '_orig']
        # tolerance relative to the MOVEMENT of the tensor (Adam steps are lr-sized whatever the gradient scale)
        move = (ref - g['start_state'][k]).norm()
        assert (v - ref).norm() <= 1e-3 * move + 1e-9, k
    # the schedule of the loop: epoch numbering continues at the checkpoint's epoch, modulus start + finetune epochs
    it = iter(g['lrs'])
    for epoch in range(g['start_epoch'], g['start_epoch'] + g['finetune_epochs']):
        for i in range(g['iters']):
            assert abs(O.lr_at(epoch % (g['start_epoch'] + g['finetune_epochs']), i, g['data_size'], 5e-4, 1, 5)
                       - next(it)) < 1e-15
    if erb:     # after switch_to_deploy the folded kernels equal the fold of the frozen pruned branches
        fin = g['final_state']
        full = dict(sd)
        full.update(eff)
        for i in range(len(cfg_of(g)['strides'])):
            K, b = O.block_kernel(full, f'layers.{i}.')
            torch.testing.assert_close(K, fin[f'layers.{i}.rbr_reparam.weight'], rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(b, fin[f'layers.{i}.rbr_reparam.bias'], rtol=1e-4, atol=1e-6)


def test_train_state_prune_list_order():
    """main_eval.py:239-352: stem Linears first, then per block the branch convolutions in creation order."""
    from orepnerv.main_eval import train_state_prunable
    from orepnerv.model import Generator
    kw = dict(embed_length=8, stem_dim_num='16_1', fc_hw_dim='3_4_4', expansion=1, num_blocks=1, norm='none',
              act='swish', bias=True, reduction=2, conv_type='conv', stride_list=[2, 2], sin_res=True,
              lower_width=4, sigmoid=False, deploy=False)
    erb = [n for n, _ in train_state_prunable(Generator(branch_type='ERB', **kw))]
    assert erb[:2] == ['stem.0', 'stem.2'] and erb[2:8] == [
        'layers.0.rbr_3x3_branch', 'layers.0.rbr_3x1_branch', 'layers.0.rbr_1x3_branch',
        'layers.0.rbr_1x1_3x3_1x1_branch_1x1_1', 'layers.0.rbr_1x1_3x3_1x1_branch_3x3',
        'layers.0.rbr_1x1_3x3_1x1_branch_1x1_2'] and len(erb) == 14
    van = [n for n, _ in train_state_prunable(Generator(branch_type='NeRV_vanilla', **kw))]
    assert van == ['stem.0', 'stem.2', 'layers.0.branch', 'layers.1.branch']


# ------------------------------------------------------------------------------------------- multi-resolution heads
def test_oracle_multires_matches_reference(golden):
    g = golden("small_erb_multires.pt")
    cfg = cfg_of(g)
    params = {k: v.clone().requires_grad_(True) for k, v in g['init_state'].items()}
    loss, imgs, targets = O.multires_loss(params, g['embed'], g['target'], cfg, g['lw'])
    assert len(imgs) == len(g['imgs']) == 2
    for a, b in zip(imgs, g['imgs']):
        torch.testing.assert_close(a.detach(), b, rtol=1e-5, atol=1e-6)
    for a, b in zip(targets, g['targets']):
        torch.testing.assert_close(a, b, rtol=0, atol=1e-7)
    torch.testing.assert_close(loss.detach(), g['loss_sum'], rtol=1e-5, atol=1e-6)
    for (k, p), gr in zip(params.items(), torch.autograd.grad(loss, list(params.values()))):
        torch.testing.assert_close(gr, g['grads'][k], rtol=5e-4, atol=2e-7, msg=lambda m: f"{k}: {m}")


def test_generator_multires_layout(golden):
    """sin_res=False: a head on every stage, same state-dict keys and initial values as the reference."""
    from orepnerv.model import Generator
    g = golden("small_erb_multires.pt")
    c = g['cfg']
    torch.manual_seed(1)
    gen = Generator(embed_length=80, stem_dim_num=c['stem_dim_num'], fc_hw_dim=c['fc_hw_dim'], expansion=c['expansion'],
                    num_blocks=1, norm='none', act='swish', bias=True, reduction=c['reduction'], conv_type='conv',
                    stride_list=c['strides'], sin_res=False, lower_width=c['lower_width'], sigmoid=False, deploy=False,
                    branch_type='ERB')
    sd = gen.state_dict()
    assert list(sd) == list(g['init_state'])
    assert all(torch.equal(sd[k], v) for k, v in g['init_state'].items())


@pytest.mark.parametrize("bt", ["DBB", "ECB"])
def test_branch_fold_backward_is_linear_in_dk(hostlib, bt):
    """The data-parallel exchange all-reduces the folded-kernel gradient dK|dbias and runs the fold backward on the sum on
    every rank (SURVEY.md 8e option 2).  That is only right because the fold backward is linear in (dK, dbias) for fixed
    branch parameters: bwd(dK1 + dK2) == bwd(dK1) + bwd(dK2) — checked on the kernels' own arithmetic."""
    from orepnerv import branches
    from orepnerv.model import NeRVBlock
    torch.manual_seed(21)
    cin, cout = 6, 24
    blk = NeRVBlock(ngf=cin, new_ngf=cout, stride=1, bias=True, norm='none', act='swish', deploy=False,
                    conv_type='conv', branch_type=bt)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith('.scale') or n.endswith('.b0'):
                p.copy_(torch.randn_like(p))
    slots = branches.branch_slots(blk)
    s = branches.make_branch_set(cin, cout, {slot: t.detach() for slot, _, t in slots})
    gen = torch.Generator().manual_seed(22)

    def bwd(dK, db):
        grads = {n: torch.zeros_like(t) for _, n, t in slots if not n.endswith('.mask')}
        gset = branches.make_branch_set(cin, cout, {slot: grads.get(n) for slot, n, _ in slots})
        hostlib.host_branch_fold_bwd(C.byref(s), C.c_void_p(dK.data_ptr()), C.c_void_p(db.data_ptr()), C.byref(gset))
        return grads

    dK1, dK2 = torch.randn(cout, cin, 3, 3, generator=gen), torch.randn(cout, cin, 3, 3, generator=gen)
    db1, db2 = torch.randn(cout, generator=gen), torch.randn(cout, generator=gen)
    g1, g2, g12 = bwd(dK1, db1), bwd(dK2, db2), bwd(dK1 + dK2, db1 + db2)
    for n in g12:
        torch.testing.assert_close(g12[n], g1[n] + g2[n], rtol=1e-4, atol=1e-5, msg=lambda m: f"{n}: {m}")
