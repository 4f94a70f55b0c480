"""Driver entry points.

build(): compiles every CUDA source of the package for sm_100a into the in-tree shared library
         (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo) plus the device self-test binary, loads the
         library and checks it exports every symbol include/orepnerv.h declares.  Needs no GPU.
         (The oracle is pure Python: there is no C restatement or oracle/_ref to compile — the reference is
         a Python/PyTorch program, see DESIGN.md.)
smoke(): one small invocation of the hot path on cuda:0 — decoder forward, Fusion6 loss, full backward and a
         fused Adam step of a small ERB model — checked against the CPU oracle.
"""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "orepnerv_build", os.path.join(ROOT, "boosting-neural-video-representation-via-online-structural-"
                                             "reparameteration_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = mod.build_all(force=False, verbose=bool(os.environ.get("ONR_BUILD_VERBOSE")))
    from orepnerv import _lib
    lib = _lib.load()
    missing = [s for s in _lib.declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"{path} does not export: {missing}")
    import orepnerv.model, orepnerv.utils, orepnerv.trainer, orepnerv.optim  # noqa: F401,E401
    print(f"built {path}; {len(_lib.declared_symbols())} C-ABI symbols exported")


def smoke():
    import argparse
    import torch
    from oracle import nerv_oracle as O
    from orepnerv.model import Generator
    from orepnerv.optim import FusedAdam
    from orepnerv.utils import PositionalEncoding, loss_fn, adjust_lr

    dev = torch.device("cuda:0")
    cfg = dict(embed='1.25_40', stem_dim_num='64_1', fc_hw_dim='3_4_12', strides=[3, 2])
    torch.manual_seed(1)
    pe = PositionalEncoding(cfg['embed'])
    gen = Generator(embed_length=pe.embed_length, stem_dim_num=cfg['stem_dim_num'], fc_hw_dim=cfg['fc_hw_dim'],
                    expansion=1, num_blocks=1, norm='none', act='swish', bias=True, reduction=2, conv_type='conv',
                    stride_list=cfg['strides'], sin_res=True, lower_width=8, sigmoid=False, deploy=False,
                    branch_type='ERB')
    sd0 = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    gen = gen.to(dev)
    pos = torch.tensor([0.25, 0.7])
    g = torch.Generator().manual_seed(7)
    target = torch.randint(0, 256, (2, 3, 18, 24), generator=g).float().div(255)
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5)
    opt = FusedAdam(gen.parameters(), betas=(0.5, 0.999))
    img = gen(pe(pos))[0]
    loss = loss_fn(img, target.to(dev), args)
    adjust_lr(opt, 0, 0, 4, args)
    opt.zero_grad()
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    ocfg = dict(fc_h=3, fc_w=4, fc_dim=12, strides=cfg['strides'], sigmoid=False)
    embed = O.pos_encoding(pos, 1.25, 40)
    lr = O.lr_at(0, 0, 4, 5e-4, 1, 5)
    _, _, loss_ref, img_ref, grads_ref = O.train_step(sd0, {}, embed, target, ocfg, lr, 1)
    rel = ((img.detach().cpu() - img_ref).norm() / img_ref.norm()).item()
    dl = abs(loss.item() - loss_ref.item())
    print(f"smoke: image rel-L2 vs oracle {rel:.3e}, |loss diff| {dl:.3e}, loss {loss.item():.5f}")
    assert rel < 1e-2 and dl < 2e-3, "CUDA path disagrees with the CPU oracle"
    for k, p in gen.named_parameters():
        moved = (p.detach().cpu() - sd0[k]).abs().max().item()
        assert moved > 0, f"{k} was not updated"


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "build"
    {"build": build, "smoke": smoke}[cmd]()
