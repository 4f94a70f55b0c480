"""torchrun worker of tests/test_gpu_dp.py: the SHIPPED data-parallel path (FrameFitter, world 2, one CUDA graph with
the bucketed NCCL all-reduces captured inside) against the oracle's batch-2 step.

Each rank fits its own frame of the 2-frame golden clip (batch 1 per rank); K ranks x batch 1 must equal the reference
run with `-b K` (every loss term is a batch mean).  Checks: (1) per-step loss/PSNR of the two ranks
average to the oracle's batch-2 values, (2) the parameters of the ranks stay BIT-identical to each other, (3) the
parameter movement matches the oracle's, (4) both exchange schemes (bucketed dK / flat) give the same parameters up to
summation order.  Prints one JSON line on rank 0.
"""
import argparse
import gc
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from oracle import nerv_oracle as O
    from orepnerv.trainer import FrameFitter
    from fullsize_util import build, ocfg
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    g = torch.load(os.path.join(ROOT, "tests", "golden", "small_erb.pt"), map_location="cpu", weights_only=False)
    cfg = g['cfg']
    frames_u8 = (g['target'] * 255).round().to(torch.uint8)        # [2,3,H,W]
    target = frames_u8.float().div(255)
    steps = 4
    args = argparse.Namespace(loss_type='Fusion6', lr=5e-4, lr_type='cosine', warmup=1, epochs=5, beta=0.5, batchSize=1)
    results = {}
    for scheme in ("bucket", "flat"):
        os.environ["ONR_DP_EXCHANGE"] = scheme
        pe, gen = build(cfg, "ERB", dev)
        fit = FrameFitter(gen, pe, args, world_size=world, data_size=4, steps_per_epoch=2, use_graph=True,
                          with_msssim=False)
        outs = []
        for t in range(steps):
            outs.append(fit.step(frames_u8[rank:rank + 1].to(dev), g['pos'][rank:rank + 1].to(dev))[:5].clone())
        outs = torch.stack(outs)                                   # [steps, 5] loss, l1, ssim, mse, psnr
        gathered = [torch.zeros_like(outs) for _ in range(world)]
        dist.all_gather(gathered, outs)
        flat = torch.cat([p.detach().reshape(-1) for p in gen.parameters()])
        flats = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(flats, flat)
        results[scheme] = (torch.stack(gathered).cpu(), [f.cpu() for f in flats],
                           {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()})
        fit.release_graph()
        del fit, gen
        gc.collect()
    if rank == 0:
        sd, state = {k: v.clone() for k, v in g['init_state'].items()}, {}
        embed = O.pos_encoding(g['pos'], 1.25, 40)
        ref = []
        for t in range(steps):
            lr = O.lr_at(t // 2, t % 2, 4, 5e-4, 1, 5)
            sd, state, loss, img, _ = O.train_step(sd, state, embed, target, ocfg(cfg), lr, t + 1)
            ref.append((loss.item(), torch.mean((img - target) ** 2).item()))
        outs, flats, state_b = results["bucket"]
        rep = {"world": world, "ranks_bit_identical": all(torch.equal(flats[0], f) for f in flats[1:])}
        # batch-mean loss == mean of the per-rank losses; batch MSE == mean of the per-rank MSEs
        rep["max_loss_err"] = max(abs(outs[:, t, 0].mean().item() - ref[t][0]) for t in range(steps))
        rep["max_mse_rel_err"] = max(abs(outs[:, t, 3].mean().item() - ref[t][1]) / ref[t][1] for t in range(steps))
        worst = 0.0
        for k, v in state_b.items():
            moved_ref = sd[k] - g['init_state'][k]
            moved = v - g['init_state'][k]
            worst = max(worst, ((moved - moved_ref).norm() / (moved_ref.norm() + 1e-12)).item())
        rep["max_movement_rel_err"] = worst
        state_f = results["flat"][2]
        rep["bucket_vs_flat_rel"] = max(((state_b[k] - state_f[k]).norm() / (state_f[k].norm() + 1e-12)).item()
                                        for k in state_b)
        rep["flat_ranks_bit_identical"] = all(torch.equal(results["flat"][1][0], f) for f in results["flat"][1][1:])
        print("DP_RESULT " + json.dumps(rep), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
