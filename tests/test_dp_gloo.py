"""CPU, world_size 2, gloo: the data-parallel contract of SURVEY.md 8e — K ranks x batch 1 with averaged
gradients equals one batch-K step — checked with the oracle as the per-rank compute and the package's own
frame sharding."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import nerv_oracle as O
    from orepnerv import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_erb.pt"), weights_only=False)
    c = g['cfg']
    fh, fw, fd = [int(x) for x in c['fc_hw_dim'].split('_')]
    cfg = dict(fc_h=fh, fc_w=fw, fc_dim=fd, strides=c['strides'], sigmoid=False)
    n_frames = 2
    idx = sharding.shard_indices(n_frames, world, rank, epoch=0, shuffle=False)
    assert idx == [rank]
    pos, target = g['pos'][idx], g['target'][idx]
    params = {k: v.clone().requires_grad_(True) for k, v in g['init_state'].items()}
    loss = O.loss_fn(O.generator_forward(params, O.pos_encoding(pos, 1.25, 4), cfg), target)
    grads = torch.autograd.grad(loss, list(params.values()))
    flat = torch.cat([x.reshape(-1) for x in grads])
    dist.all_reduce(flat)
    flat /= world
    if rank == 0:
        q.put(flat)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_average_equals_batch_two(golden):
    g = golden("tiny_erb.pt")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = torch.cat([g['grads'][k].reshape(-1) for k in g['init_state'].keys()])
    # Fusion6 is a batch mean of L1 and of per-(frame, channel) SSIM means -> averaging is exact up to fp32
    torch.testing.assert_close(flat, ref, rtol=2e-4, atol=1e-7)
