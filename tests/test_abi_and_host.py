"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/orepnerv.h declares, the
host-side mirror of the reference API builds models with bit-identical initial parameters / state-dict
layout, and unsupported configurations fail loudly."""
import copy
import math
import os

import pytest
import torch

from orepnerv import _lib
from orepnerv.model import Generator, NeRVBlock
from orepnerv.utils import PositionalEncoding, adjust_lr, lr_multiplier
from orepnerv import sharding


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(declared) == set(_lib._SIGNATURES), set(declared) ^ set(_lib._SIGNATURES)
    assert lib.onr_abi_version() == 1


def test_tile_n_rule():
    from orepnerv.engine import conv_tile_n
    for n in (32, 96, 128, 384, 800, 864, 3200):
        bn, nt = conv_tile_n(n)
        assert bn % 32 == 0 and bn <= 256 and bn * nt >= n and bn * (nt - 1) < n
    assert conv_tile_n(384) == (128, 3)         # block 1-4 forward: 2 sub-tiles share each weight tile
    assert conv_tile_n(96) == (96, 1)           # dgrad of those blocks
    assert conv_tile_n(32) == (32, 1)
    lib = _lib.load()
    assert lib.onr_conv_tile_n(33, None, None) < 0        # not a multiple of 32 -> error code, message set
    assert b"multiple of 32" in lib.onr_last_error()


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.lib()
    pe = PositionalEncoding('1.25_4')
    with pytest.raises(RuntimeError):
        pe(torch.tensor([0.5]))


def build(g, branch_type, deploy=False):
    c = g['cfg']
    torch.manual_seed(1)
    pe = PositionalEncoding(c['embed'])
    gen = Generator(embed_length=pe.embed_length, stem_dim_num=c['stem_dim_num'], fc_hw_dim=c['fc_hw_dim'],
                    expansion=c['expansion'], num_blocks=1, norm='none', act='swish', bias=True,
                    reduction=c['reduction'], conv_type='conv', stride_list=c['strides'], sin_res=True,
                    lower_width=c['lower_width'], sigmoid=False, deploy=deploy, branch_type=branch_type)
    return pe, gen


@pytest.mark.parametrize("name,bt", [("tiny_erb.pt", "ERB"), ("tiny_vanilla.pt", "NeRV_vanilla"),
                                     ("small_erb.pt", "ERB"), ("small_acb.pt", "ACB"), ("small_repvgg.pt", "RepVGG"),
                                     ("small_dbb.pt", "DBB"), ("small_ecb.pt", "ECB")])
def test_init_is_bit_identical_to_reference(golden, name, bt):
    g = golden(name)
    _, gen = build(g, bt)
    sd = gen.state_dict()
    assert list(sd.keys()) == list(g['init_state'].keys())
    for k, v in g['init_state'].items():
        assert sd[k].dtype == torch.float32 and sd[k].shape == v.shape
        assert torch.equal(sd[k], v), k


def test_deploy_layout_and_checkpoint_roundtrip(golden, tmp_path):
    g = golden("tiny_erb.pt")
    _, dep = build(g, "ERB", deploy=True)
    assert list(dep.state_dict().keys()) == list(g['deploy_state'].keys())
    dep.load_state_dict(g['deploy_state'])
    path = os.path.join(tmp_path, "model_latest_deploy.pth")
    torch.save({'epoch': 1, 'state_dict': dep.state_dict()}, path)
    _, dep2 = build(g, "ERB", deploy=True)
    dep2.load_state_dict(torch.load(path, weights_only=True)['state_dict'])
    for k, v in g['deploy_state'].items():
        assert torch.equal(dep2.state_dict()[k], v)


def test_deepcopy_drops_executors(golden):
    g = golden("tiny_erb.pt")
    _, gen = build(g, "ERB")
    gen._executors = {"x": object()}
    cp = copy.deepcopy(gen)
    assert cp._executors == {} and list(cp.state_dict()) == list(gen.state_dict())


@pytest.mark.parametrize("kw", [dict(branch_type='OREPA'), dict(act='mish'), dict(norm='bn'), dict(num_blocks=2),
                                dict(stem_dim_num='16_2')])
def test_out_of_scope_configs_raise(kw):
    base = dict(embed_length=8, stem_dim_num='16_1', fc_hw_dim='3_4_4', expansion=1, num_blocks=1, norm='none',
                act='swish', bias=True, reduction=2, conv_type='conv', stride_list=[2, 2], sin_res=True,
                lower_width=4, sigmoid=False, deploy=False, branch_type='ERB')
    base.update(kw)
    with pytest.raises(NotImplementedError):
        Generator(**base)


def test_lr_schedule_matches_golden(golden):
    import argparse
    m = golden('misc.pt')
    args = argparse.Namespace(lr=5e-4, lr_type='cosine', warmup=60, epochs=300)

    class Opt:
        param_groups = [{'lr': 0.0}]
    for epoch, it, lr_ref in m['lr_sched']:
        assert adjust_lr(Opt(), epoch, it, 132, args) == lr_ref


def test_shard_indices_follow_drop_last_batches():
    """K ranks x batch 1 == the reference DataLoader with -b K and drop_last=True (main_train.py:207-209): floor(N/K)
    steps per epoch, disjoint frames, the N % K tail of the epoch's permutation skipped."""
    for n, k in [(132, 8), (132, 1), (600, 8), (7, 4)]:
        per_rank = [sharding.shard_indices(n, k, r, epoch=3) for r in range(k)]
        steps = sharding.steps_per_epoch(n, k)
        assert steps == n // k
        assert all(len(p) == steps for p in per_rank)
        seen = [i for p in per_rank for i in p]
        assert len(set(seen)) == len(seen) == steps * k and set(seen) <= set(range(n))
        perm = sharding.epoch_permutation(n, 3)
        assert set(seen) == set(perm[:steps * k])
    assert sharding.shard_indices(10, 2, 0, 0, shuffle=False) == [0, 2, 4, 6, 8]
    assert sharding.steps_per_epoch(132, 8) == 16
    # fewer frames than ranks: every rank still gets one frame
    assert [len(sharding.shard_indices(2, 4, r, 0)) for r in range(4)] == [1, 1, 1, 1]


def test_flat_gradient_buffer_is_aligned_and_complete(golden):
    """alloc_grads: one flat fp32 buffer (the all-reduce payload), one view per parameter, every view starting on a
    256-byte boundary (the fused Adam kernel takes its 128-bit path only on 16-byte aligned tensors), no overlap."""
    g = golden("tiny_erb.pt")
    _, gen = build(g, "ERB")
    grads = gen.alloc_grads()
    flat = grads.pop("__flat__")
    offsets = grads.pop("__offsets__")
    named = dict(gen.named_parameters())
    assert all(offsets[n] == ((grads[n].data_ptr() - flat.data_ptr()) // 4, named[n].numel()) for n in named)
    assert set(grads) == set(named)
    spans = []
    for n, p in named.items():
        v = grads[n]
        assert v.shape == p.shape and v.dtype == torch.float32
        off = (v.data_ptr() - flat.data_ptr())
        assert off % 256 == 0, n
        spans.append((off // 4, off // 4 + p.numel()))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert spans[-1][1] <= flat.numel() < spans[-1][1] + 64
    flat.fill_(1.0)
    assert all(float(grads[n].sum()) == named[n].numel() for n in named)
