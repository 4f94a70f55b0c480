"""Helpers of the full-size GPU parity tests: the fp32 GPU oracle and the convergence run.

The oracle side is `oracle/nerv_oracle.py` evaluated on CUDA tensors with TF32 switched off (cuDNN / cuBLAS fp32
arithmetic of the restated reference algorithm).  Test infrastructure only: nothing in the package imports this.
"""
import argparse
import contextlib
import json
import math
import os
import time

import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import nerv_oracle as O  # noqa: E402

# BASELINE.json configs[1..3] geometries (README flags: --embed 1.25_40 --stem_dim_num 512_1 --reduction 2
# --expansion 1 --lower_width 96)
CONFIGS = {
    "S720": dict(embed='1.25_40', stem_dim_num='512_1', fc_hw_dim='9_16_26', expansion=1, reduction=2,
                 lower_width=96, strides=[5, 2, 2, 2, 2], H=720, W=1280),
    "L720": dict(embed='1.25_40', stem_dim_num='512_1', fc_hw_dim='9_16_112', expansion=1, reduction=2,
                 lower_width=96, strides=[5, 2, 2, 2, 2], H=720, W=1280),
    "U1080": dict(embed='1.25_40', stem_dim_num='512_1', fc_hw_dim='9_16_26', expansion=1, reduction=2,
                  lower_width=96, strides=[5, 3, 2, 2, 2], H=1080, W=1920),
}


@contextlib.contextmanager
def fp32_oracle_math():
    """TF32 off everywhere: the oracle's convolutions / matmuls run in true fp32 on the GPU."""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def rel_l2(a, b):
    a, b = a.detach().double(), b.detach().double().to(a.device)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


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


def oracle_embed(pos, dev):
    """reference main_train.py:234-235: the positional encoding is evaluated on the CPU in fp32, then uploaded."""
    return O.pos_encoding(pos.detach().cpu().float(), 1.25, 40).to(dev)


def record(name, payload):
    """Measured parity numbers are appended to gpurun_out/parity_fullsize.jsonl (copied to profiles/ by hand)."""
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_fullsize.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **payload}) + "\n")
    except OSError:
        pass


def movement_ok(tag, key, moved, moved_ref, tol):
    """Adam normalises every update to ~lr per element, so trained tensors are compared by the MOVEMENT of each tensor:
    ||moved - moved_ref|| <= tol * ||moved_ref||.  The measured ratio is appended to gpurun_out/movement.jsonl so the
    tolerances in the tests can be (and were) set from measurements."""
    err, ref = (moved - moved_ref).norm().item(), moved_ref.norm().item()
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "movement.jsonl"), "a") as f:
            f.write(json.dumps({"test": tag, "tensor": key, "ratio": err / (ref + 1e-30), "ref_norm": ref}) + "\n")
    except OSError:
        pass
    return err <= tol * ref + 1e-7


def one_step_parity(name, dev, branch_type="ERB"):
    """Forward, Fusion6 loss and every parameter gradient of ONE full-size frame: CUDA path (reference-shaped API)
    vs the fp32 GPU oracle on the same parameters / frame.  Returns the measured errors."""
    from orepnerv.data import synthetic_clip
    from orepnerv.utils import loss_fn
    cfg = CONFIGS[name]
    pe, gen = build(cfg, branch_type, dev)
    pos = torch.tensor([5 / 132])
    frame = synthetic_clip(6, cfg['H'], cfg['W'], device=dev)[5:6]
    target = frame.float().div(255)
    img = gen(pe(pos))[0]
    loss = loss_fn(img, target, argparse.Namespace(loss_type='Fusion6'))
    loss.backward()
    torch.cuda.synchronize()
    res = {"config": name, "branch_type": branch_type}
    with fp32_oracle_math():
        params = {k: v.detach().clone().requires_grad_(True) for k, v in gen.state_dict().items()}
        img_ref = O.generator_forward(params, oracle_embed(pos, dev), ocfg(cfg))
        loss_ref = O.loss_fn(img_ref, target)
        grads_ref = torch.autograd.grad(loss_ref, list(params.values()))
    res["img_rel_l2"] = rel_l2(img, img_ref)
    res["loss"], res["loss_ref"] = loss.item(), loss_ref.item()
    res["psnr"], res["psnr_ref"] = O.psnr(img.detach(), target).item(), O.psnr(img_ref.detach(), target).item()
    gerr = {}
    named = dict(gen.named_parameters())
    for (k, _), gr in zip(params.items(), grads_ref):
        gerr[k] = (named[k].grad - gr).norm().item() / (gr.norm().item() + 1e-30)
    res["grad_rel_l2_max"] = max(gerr.values())
    res["grad_rel_l2_worst"] = max(gerr, key=gerr.get)
    res["grad_rel_l2"] = gerr
    # the folded kernels themselves (fp32 gate of north_star: 1e-5)
    if branch_type == "ERB":
        ferr = []
        with torch.no_grad(), fp32_oracle_math():
            for i, blk in enumerate(gen.layers):
                K, b = blk.get_equivalent_kernel_bias()
                K_ref, b_ref = O.block_kernel({k: v.detach().double() for k, v in params.items()}, f'layers.{i}.')
                ferr.append(max(rel_l2(K, K_ref), rel_l2(b, b_ref)))
        res["fold_rel_l2_max"] = max(ferr)
    return res


def frame_order(n_frames, epoch, seed=1234):
    g = torch.Generator().manual_seed(seed + epoch)
    return torch.randperm(n_frames, generator=g).tolist()


def convergence_run(dev, name="S720", n_frames=16, epochs=30, lr=5e-4, log=None, repeats=1):
    """The README recipe (main_train.py:222-267: Adam beta 0.5, cosine schedule with 20 % warm-up, Fusion6, batch 1)
    on a reduced clip, run twice on the same frames / seed / frame order: (a) the CUDA path (FrameFitter), (b) the
    fp32 GPU oracle.  Returns per-epoch mean training PSNR of both plus the final full-clip eval PSNR of both."""
    from orepnerv.data import synthetic_clip
    from orepnerv.trainer import FrameFitter
    cfg = CONFIGS[name]
    clip = synthetic_clip(n_frames, cfg['H'], cfg['W'], device=dev)
    warm = int(0.2 * epochs)
    args = argparse.Namespace(loss_type='Fusion6', lr=lr, lr_type='cosine', warmup=warm, epochs=epochs, beta=0.5,
                              batchSize=1)
    t_all = torch.arange(n_frames, dtype=torch.float32) / n_frames
    out = {"config": name, "n_frames": n_frames, "epochs": epochs, "lr": lr}

    # ---- (a) ours (repeats > 1: the run is repeated from the same initial state; wgrad accumulates with atomics, so two
    #      runs differ in the last bits per step and the chaotic fit amplifies that — the spread is the noise floor)
    init = None
    ours_evals = []
    for rep in range(repeats):
        pe, gen = build(cfg, "ERB", dev)
        if init is None:
            init = {k: v.detach().clone() for k, v in gen.state_dict().items()}
        fit = FrameFitter(gen, pe, args, data_size=n_frames, steps_per_epoch=n_frames, use_graph=True,
                          with_msssim=False)
        t_dev = t_all.to(dev)
        t0 = time.time()
        curve = []
        for ep in range(epochs):
            acc = []
            for i in frame_order(n_frames, ep):
                acc.append(fit.step(clip[i:i + 1], t_dev[i:i + 1])[4:5].clone())
            curve.append(torch.cat(acc).mean().item())
        torch.cuda.synchronize()
        with torch.no_grad():
            ps = []
            for i in range(n_frames):
                img = gen(pe(t_dev[i:i + 1]))[0]
                ps.append(O.psnr(img, clip[i:i + 1].float().div(255)).item())
        ours_evals.append(sum(ps) / len(ps))
        if rep == 0:
            out["ours_s"] = time.time() - t0
            out["ours_train_psnr"] = curve
            out["ours_eval_psnr"] = ours_evals[0]
        del fit, gen
        torch.cuda.empty_cache()
    out["ours_eval_psnr_repeats"] = ours_evals

    # ---- (b) fp32 GPU oracle: same initial state, same frames, same order, same schedule
    oc = ocfg(cfg)
    oracle_evals = []
    for rep in range(repeats):
        sd, state = {k: v.clone() for k, v in init.items()}, {}
        t0 = time.time()
        curve = []
        step = 0
        with fp32_oracle_math():
            for ep in range(epochs):
                acc = []
                for it, i in enumerate(frame_order(n_frames, ep)):
                    step += 1
                    lr_t = O.lr_at(ep, it, n_frames, lr, warm, epochs)
                    target = clip[i:i + 1].float().div(255)
                    sd, state, loss, img, _ = O.train_step(sd, state, oracle_embed(t_all[i:i + 1], dev), target, oc,
                                                           lr_t, step)
                    acc.append(O.psnr(img, target).reshape(1))
                curve.append(torch.cat(acc).mean().item())
                if log and rep == 0:
                    log(f"oracle epoch {ep}: train PSNR {curve[-1]:.3f} (ours {out['ours_train_psnr'][ep]:.3f})")
            torch.cuda.synchronize()
            with torch.no_grad():
                ps = []
                for i in range(n_frames):
                    img = O.generator_forward(sd, oracle_embed(t_all[i:i + 1], dev), oc)
                    ps.append(O.psnr(img, clip[i:i + 1].float().div(255)).item())
            oracle_evals.append(sum(ps) / len(ps))
        if rep == 0:
            out["oracle_s"] = time.time() - t0
            out["oracle_train_psnr"] = curve
            out["oracle_eval_psnr"] = oracle_evals[0]
    out["oracle_eval_psnr_repeats"] = oracle_evals
    out["delta_eval_psnr"] = out["ours_eval_psnr"] - out["oracle_eval_psnr"]
    out["delta_train_psnr_last"] = out["ours_train_psnr"][-1] - out["oracle_train_psnr"][-1]
    return out


if __name__ == "__main__":
    # python tests/fullsize_util.py [config] [frames] [epochs] [repeats]  -> gpurun_out/convergence_<config>_*.json
    import sys
    name = sys.argv[1] if len(sys.argv) > 1 else "S720"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    e = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    r = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    res = convergence_run(torch.device("cuda:0"), name, n, e, log=lambda s: print(s, flush=True), repeats=r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"convergence_{name}_{n}f_{e}e.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if not isinstance(v, list)}))
