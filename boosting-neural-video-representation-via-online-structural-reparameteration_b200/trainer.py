"""Fast path of the frame-fitting step: one call = the body of the reference hot loop.

Reference main_train.py:229-254, per iteration:
    embed = PE(norm_idx); data.cuda(); out = model(embed); loss = Fusion6(out, data); adjust_lr();
    zero_grad(); loss.backward(); optimizer.step(); psnr_fn(); msssim_fn()
`FrameFitter.step` runs exactly that sequence as liborepnerv.so kernels on persistent buffers with no
autograd graph, no per-step allocation and no host synchronisation, and (optionally) replays it as ONE
CUDA graph.  With world_size > 1 every rank fits its own frame (frame-sharded data parallel) and the
flat fp32 gradient buffer is all-reduced over NCCL before the fused Adam step, which matches the
reference run with `-b world_size` because every loss term is a batch mean.
"""
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr
from .optim import FusedAdam
from .utils import LOSS_TERMS, lr_multiplier


class _Exchange:
    """What the executor's backward calls for the data-parallel gradient exchange: `exchange(t)` all-reduces a bucket;
    `all_gather_slots` / `stem_slots` serve the factored stem exchange (engine.NetExecutor.backward)."""

    def __init__(self, world):
        self.world = world          # (no reference to the fitter: a cycle would keep its CUDA graph alive past `del`)
        self._slots = None
        if os.environ.get("ONR_DP_STEM", "factors") != "factors":
            self.all_gather_slots = None        # ONR_DP_STEM=allreduce: all-reduce the stem matrices (A/B timing)
        # The stem gather closes the backward on the main stream; on the default communicator it queues behind the last
        # blocks' bucket all-reduces (one NCCL stream per communicator, FIFO).  Its own communicator
        # (ONR_DP_GATHER_GROUP=1) lets it run beside them — measured on 8 B200s: 6113 vs 6140 frames/s with the shared
        # one, so the shared communicator stays the default.
        self.gather_group = (dist.new_group() if (world > 1 and os.environ.get("ONR_DP_GATHER_GROUP", "0") == "1")
                             else None)

    def __call__(self, t):
        dist.all_reduce(t)

    def stem_slots(self, ex):
        if self._slots is None:
            gen = ex.gen
            n = int(ex.lib.onr_stem_factor_floats(gen.fc_dim, gen.fc_h, gen.fc_w, ex.hid, ex.E))
            self._slots = torch.zeros(self.world, n, dtype=torch.float32, device=ex.dev)
        return self._slots, dist.get_rank(), self.world

    def all_gather_slots(self, slots, rank):
        dist.all_gather_into_tensor(slots, slots[rank], group=self.gather_group)


class FrameFitter:
    def __init__(self, model, pe, args, optimizer=None, world_size=1, steps_per_epoch=None, data_size=None,
                 use_graph=True, with_msssim=True, grad_masks=None, epoch_offset=0, epoch_mod=None):
        """grad_masks (prune-then-finetune, reference main_eval.py:213-545): dict parameter name -> fp32 mask of the
        parameter's shape.  Every step the gradients are multiplied by the masks before Adam — the gradient of prune's
        `weight = weight_orig * weight_mask` — so a masked entry (zero gradient from the first step on a fresh
        optimizer) never moves; an all-zero mask freezes a tensor.
        epoch_offset / epoch_mod: the LR schedule is evaluated at epoch (epoch_offset + step // steps_per_epoch) %
        epoch_mod (default: args.epochs) while the Adam step count starts at 1 (main_eval.py:446-466)."""
        self.lib = _lib.lib()
        self.model, self.pe, self.args = model, pe, args
        self.world = world_size
        self.B = args.batchSize
        if args.loss_type not in LOSS_TERMS:
            raise NotImplementedError(f"loss_type {args.loss_type!r} is outside the B200 hot path")
        self.w_l1, self.w_mse, self.w_ssim = LOSS_TERMS[args.loss_type]
        self.ex = model.executor(self.B, True)
        if self.ex.multi:
            raise NotImplementedError("FrameFitter fits the single-resolution head; multi-resolution heads "
                                      "(sin_res=False) train through the module API (main_train does that)")
        dev = self.ex.dev
        self.dev = dev
        H, W = self.ex.H, self.ex.W
        self.H, self.W = H, W
        # persistent gradients: one flat fp32 buffer, parameters' .grad are views into it
        self.grads = model.alloc_grads()
        self.flat_grad = self.grads["__flat__"]
        for n, p in model.named_parameters():
            if p.requires_grad:                       # SeqConv3x3.mask (ECB) is a constant Parameter
                p.grad = self.grads[n]
        self.opt = optimizer if optimizer is not None else FusedAdam(model.parameters(), lr=args.lr,
                                                                      betas=(args.beta, 0.999))
        if not isinstance(self.opt, FusedAdam):
            raise TypeError("FrameFitter drives orepnerv.optim.FusedAdam")
        self.opt.fused_zero_grad = True
        self.opt.grad_scale = 1.0 / world_size
        self.data_size = data_size if data_size is not None else 1
        self.steps_per_epoch = steps_per_epoch if steps_per_epoch is not None else self.data_size
        self.device_sched = args.lr_type in ('cosine', 'const')
        self.with_msssim = with_msssim and H >= 160 and min(H, W) > 160
        # static I/O buffers (CUDA-graph friendly)
        self.frame_u8 = torch.zeros(self.B, 3, H, W, dtype=torch.uint8, device=dev)
        self.t_norm = torch.zeros(self.B, dtype=torch.float32, device=dev)
        self.target = torch.zeros(self.B, 3, H, W, dtype=torch.float32, device=dev)
        self.gimg = torch.zeros(self.B, 3, H, W, dtype=torch.float32, device=dev)
        self.out = torch.zeros(8, dtype=torch.float32, device=dev)   # loss, l1, ssim, mse, psnr, msssim, lr, -
        self.loss_work = torch.empty(self.lib.onr_loss_workspace_bytes(self.B, H, W), dtype=torch.uint8, device=dev)
        self.ms_work = (torch.empty(self.lib.onr_msssim_workspace_bytes(self.B, H, W), dtype=torch.uint8, device=dev)
                        if self.with_msssim else None)
        self.freqs = pe.freqs(dev)
        # Fold-ahead (ONR_FOLD_AHEAD=1; single GPU, device-side LR schedule): a block's parameters are updated on its
        # side stream as soon as its gradients exist and the block is re-folded / re-packed for the NEXT step right
        # there, beside the tail of the backward, instead of at the head of the next step where the first
        # convolutions wait for it.  Same results; measured SLOWER on B200 (759 vs 777 frames/s: the step is bound by
        # total work, and six Adam launches replace one), hence off by default.
        self.fold_ahead = (world_size == 1 and args.lr_type in ('cosine', 'const')
                           and os.environ.get("ONR_FOLD_AHEAD", "0") == "1")
        # Early Adam (ONR_ADAM_EARLY, default blocks 2,3,4): the update of a late block runs on that block's side stream
        # as soon as its gradients are final (after its all-reduce + fold backward), beside the rest of the backward,
        # so the Adam launch at the end of the step — on the critical path — only covers the early blocks, stem and head.
        # Needs the device-side schedule (the tick moves to the head of the step).
        early = os.environ.get("ONR_ADAM_EARLY", "2,3,4")
        self.adam_early = ([int(x) for x in early.split(",") if x.strip() != ""]
                           if (args.lr_type in ('cosine', 'const') and not self.fold_ahead) else [])
        named = list(model.named_parameters())
        self._block_params = [[p for n, p in named if n.startswith(f"layers.{l}.")] for l in range(self.ex.L)]
        in_blocks = {id(p) for ps in self._block_params for p in ps}
        self._rest_params = [p for _, p in named if id(p) not in in_blocks]
        self.adam_early = [l for l in self.adam_early if 0 <= l < self.ex.L]
        self.epoch_offset = int(epoch_offset)
        self.epoch_mod = int(epoch_mod) if epoch_mod is not None else int(args.epochs)
        self.grad_mask = None
        if grad_masks:
            # one flat mask beside the flat gradient buffer (ones outside the masked tensors): ONE launch per step
            self.grad_mask = torch.ones_like(self.flat_grad)
            offs = self.grads["__offsets__"]
            for n, m in grad_masks.items():
                o, k = offs[n]
                if m.numel() != k:
                    raise ValueError(f"grad mask of {n} has {m.numel()} elements, the parameter {k}")
                self.grad_mask[o:o + k] = m.reshape(-1).to(self.dev, torch.float32)
            # the masks apply to the complete gradients, right before the one Adam launch at the end of the step
            self.fold_ahead, self.adam_early = False, []
        early_ids = {id(p) for l in self.adam_early for p in self._block_params[l]}
        self._late_params = [p for _, p in named if id(p) not in early_ids]
        self._weights_valid = False
        # Gradient exchange (world > 1).  "bucket" (default): per-block all-reduce of the folded-kernel gradients
        # dK|dbias (12.8 MB at S720 instead of the 30 MB of branch gradients) issued on the block's side stream as soon
        # as its wgrad is done, head / stem gradients as their own buckets; everything, NCCL calls included, is
        # captured in ONE CUDA graph.  "flat" (ONR_DP_EXCHANGE=flat): round 1's single all-reduce of the whole flat
        # buffer after the backward, kept for A/B timing.
        self.exchange = os.environ.get("ONR_DP_EXCHANGE", "bucket")
        if self.exchange not in ("bucket", "flat"):
            raise ValueError("ONR_DP_EXCHANGE must be bucket or flat")
        self._exchange = _Exchange(world_size) if world_size > 1 else None
        self.graph = None
        self.use_graph = use_graph and self.device_sched
        self.host_step = 0
        self.launches_per_step = None

    # ------------------------------------------------------------------------------------------
    def _body(self):
        self._body_pre()
        if self.world > 1 and self.exchange == "flat":
            dist.all_reduce(self.flat_grad)          # one exposed all-reduce of every branch gradient (round-1 scheme)
        self._body_post()

    def _reduce(self, t):
        """Gradient exchange of one bucket: NCCL all-reduce (sum) over NVLink on the current (side) stream; the
        average is applied as grad_scale = 1/world inside the fused Adam kernel."""
        dist.all_reduce(t)


    def _body_pre(self):
        """frame conversion, forward, loss + its gradient, backward -> local gradients in self.flat_grad"""
        lib, st = self.lib, _lib.stream()
        B, H, W = self.B, self.H, self.W
        if self.fold_ahead or self.adam_early:
            self._tick()        # the per-block Adam updates inside the backward need this step's lr / step count
        # the uint8 -> fp32 conversion of the target frame is only needed by the loss: it runs on a side stream beside
        # the stem and the first convolutions instead of in front of them
        main = torch.cuda.current_stream()
        if getattr(self, "_aux_stream", None) is None:
            self._aux_stream = torch.cuda.Stream(device=self.dev)
        fork0 = torch.cuda.Event()
        fork0.record(main)
        self._aux_stream.wait_event(fork0)
        with torch.cuda.stream(self._aux_stream):
            check(lib.onr_frame_u8_to_f32(ptr(self.frame_u8), self.frame_u8.numel(), ptr(self.target), _lib.stream()),
                  "u8_to_f32")
            target_ready = torch.cuda.Event()
            target_ready.record(self._aux_stream)
        img = self.ex.forward(t_norm=self.t_norm, freqs=self.freqs, refresh=not self.fold_ahead)
        main.wait_event(target_ready)
        check(lib.onr_fusion_loss(ptr(img), ptr(self.target), B, H, W, self.w_l1, self.w_mse, self.w_ssim, 1.0,
                                  ptr(self.out), ptr(self.gimg), ptr(self.loss_work), st), "onr_fusion_loss")
        ms_done = None
        if self.with_msssim:
            # the MS-SSIM metric (reference main_train.py:254) runs on a side stream, concurrently with the whole
            # backward; its first scale is the SSIM the loss has just evaluated, so it starts after the loss
            main = torch.cuda.current_stream()
            if getattr(self, "_ms_stream", None) is None:
                self._ms_stream = torch.cuda.Stream(device=self.dev)
            fork = torch.cuda.Event()
            fork.record(main)
            self._ms_stream.wait_event(fork)
            with torch.cuda.stream(self._ms_stream):
                check(lib.onr_msssim(ptr(img), ptr(self.target), B, H, W, ptr(self.out[5:6]), ptr(self.ms_work),
                                     ptr(self.loss_work), _lib.stream()), "onr_msssim")
                ms_done = torch.cuda.Event()
                ms_done.record(self._ms_stream)
        hook = self._update_block if self.fold_ahead else (self._adam_block if self.adam_early else None)
        self.ex.backward(self.gimg, self.grads, block_hook=hook,
                         reduce=self._exchange if (self.world > 1 and self.exchange == "bucket") else None)
        if ms_done is not None:
            torch.cuda.current_stream().wait_event(ms_done)

    def _update_block(self, l):
        """backward hook (runs on block l's side stream): Adam on the block's tensors, then fold + pack for the next step"""
        self.opt.step_params(self._block_params[l])
        self.ex.fold_pack_block(l)

    def _adam_block(self, l):
        """backward hook (block l's side stream, its gradients final): early Adam for the late blocks"""
        if l in self.adam_early:
            self.opt.step_params(self._block_params[l])

    def _tick(self):
        lr_dev, step_dev = self.opt.device_scalars(self.dev)
        a = self.args
        check(self.lib.onr_sched_tick_ex(ptr(step_dev), ptr(lr_dev), float(a.lr), self.steps_per_epoch, self.data_size,
                                         int(a.warmup), int(a.epochs), 0 if a.lr_type == 'cosine' else 1,
                                         self.epoch_offset, self.epoch_mod, _lib.stream()), "onr_sched_tick_ex")

    def _body_post(self):
        """LR schedule tick + fused Adam (+ gradient averaging and zeroing); with fold-ahead only the tensors outside
        the blocks (stem, head) are left to update here"""
        if self.fold_ahead:
            self.opt.step_params(self._rest_params)
            self.opt._step_count_host += 1
            return
        if self.adam_early:
            self.opt.step_params(self._late_params)
            self.opt._step_count_host += 1
            return
        if self.device_sched:
            self._tick()
        if self.grad_mask is not None:
            check(self.lib.onr_mul_inplace_f32(ptr(self.flat_grad), ptr(self.grad_mask), self.flat_grad.numel(),
                                               _lib.stream()), "onr_mul_inplace_f32")
        self.opt.step(device_schedule=self.device_sched)

    def release_graph(self):
        """Drops the captured step graph.  With world > 1 it holds captured NCCL kernels: release it (and synchronise)
        before `dist.destroy_process_group()`, which otherwise blocks on the communicator the live graph references."""
        torch.cuda.synchronize()
        self.graph = None

    def refresh_weights(self):
        """(Re)folds and packs every block from the current parameters.  Needed before the first step and after
        anything outside `step` changed the parameters (checkpoint load, roll-back) when fold-ahead is on."""
        main = torch.cuda.current_stream()
        for ev in self.ex.refresh_weights():
            main.wait_event(ev)
        self._weights_valid = True

    def _host_lr(self):
        t = self.host_step
        epoch = (self.epoch_offset + t // self.steps_per_epoch) % self.epoch_mod
        it = t % self.steps_per_epoch
        lr = self.args.lr * lr_multiplier(epoch, it, self.data_size, self.args)
        for g in self.opt.param_groups:
            g['lr'] = lr
        return lr

    def load_inputs(self, frame_u8, t_norm):
        """Copies one step's inputs (device or pinned-host tensors) into the static buffers."""
        self.frame_u8.copy_(frame_u8.reshape(self.frame_u8.shape), non_blocking=True)
        self.t_norm.copy_(t_norm.reshape(self.t_norm.shape), non_blocking=True)

    def step(self, frame_u8=None, t_norm=None):
        """One optimisation step on the given frame(s).  Returns the device tensor
        [loss, L1, SSIM, MSE, PSNR, MS-SSIM, ...] of this step (no host sync)."""
        if frame_u8 is not None:
            self.load_inputs(frame_u8, t_norm)
        lr = self._host_lr()
        if self.fold_ahead and not self._weights_valid:
            self.refresh_weights()
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
            self.opt._step_count_host += 1
        else:
            self._body()
        self.host_step += 1
        # parameters changed through raw pointers: invalidate operand caches keyed on tensor versions
        self.model._weights_epoch = getattr(self.model, "_weights_epoch", 0) + 1
        return self.out

    def _capture(self):
        # warm-up outside the graph on a side stream (allocations, lazy inits), then capture.
        # The warm-up runs REAL steps; their effect on parameters/optimizer state is rolled back.
        params = [p.detach().clone() for p in self.model.parameters()]
        for gi, group in enumerate(self.opt.param_groups):
            self.opt._table(gi, group)                       # materialise exp_avg / exp_avg_sq
        moments = [(st['exp_avg'].clone(), st['exp_avg_sq'].clone()) for st in self.opt.state.values()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, q in zip(self.model.parameters(), params):
                p.copy_(q)
            for st, (m0, v0) in zip(self.opt.state.values(), moments):
                st['exp_avg'].copy_(m0)
                st['exp_avg_sq'].copy_(v0)
            self.flat_grad.zero_()
            _, step_dev = self.opt.device_scalars(self.dev)
            step_dev.fill_(self.host_step)
        self.opt._step_count_host = self.host_step
        if self.fold_ahead:
            self.refresh_weights()       # the packed operands must match the rolled-back parameters
        g = torch.cuda.CUDAGraph()
        if self.world == 1:
            with torch.cuda.graph(g):
                self._body()
        else:
            # the NCCL all-reduces are captured with the kernels (NCCL >= 2.9.6 supports stream capture; every rank
            # captures and replays the same sequence).  thread_local mode: the process group's watchdog thread may
            # touch the CUDA API while this thread captures.
            dist.barrier()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._body()
        # the capture pass itself did not execute; fix the host-side counter it advanced
        self.opt._step_count_host = self.host_step
        self.graph = g

    # ------------------------------------------------------------------------------------------ host pipeline
    # End-to-end driving from HOST buffers (what a data loader hands over): every step still uploads its own frame
    # (uint8, pinned) + index and downloads its own metrics, but on a copy stream, one step ahead / behind, so the
    # PCIe traffic overlaps the kernels of the neighbouring steps.
    def host_pipeline_begin(self, frame_u8_host, t_host):
        dev = self.dev
        self._cs = torch.cuda.Stream(device=dev)
        self._stage_f = [torch.empty_like(self.frame_u8) for _ in range(2)]
        self._stage_t = [torch.empty_like(self.t_norm) for _ in range(2)]
        self._ev_h2d = [torch.cuda.Event() for _ in range(2)]
        self._ev_used = [torch.cuda.Event() for _ in range(2)]
        self._out_ring = torch.zeros(4, 8, dtype=torch.float32, device=dev)
        self._host_ring = torch.zeros(4, 8, dtype=torch.float32).pin_memory()
        self._ev_out = [torch.cuda.Event() for _ in range(4)]
        self._ev_d2h = [torch.cuda.Event() for _ in range(4)]
        self._hp_i = 0
        self._hp_pending = None
        self._prefetch(0, frame_u8_host, t_host, first=True)

    def _prefetch(self, slot, frame_u8_host, t_host, first=False):
        cs = self._cs
        if not first:
            cs.wait_event(self._ev_used[slot])          # the step that read this staging slot has consumed it
        with torch.cuda.stream(cs):
            self._stage_f[slot].copy_(frame_u8_host.reshape(self.frame_u8.shape), non_blocking=True)
            self._stage_t[slot].copy_(t_host.reshape(self.t_norm.shape), non_blocking=True)
            self._ev_h2d[slot].record(cs)

    def step_host(self, next_frame_u8_host=None, next_t_host=None):
        """Runs one step on the inputs prefetched earlier, prefetches the next step's inputs, and returns the
        metrics of the PREVIOUS step as a host tensor (None on the first call)."""
        i = self._hp_i
        slot, r = i % 2, i % 4
        main = torch.cuda.current_stream()
        main.wait_event(self._ev_h2d[slot])
        self.frame_u8.copy_(self._stage_f[slot], non_blocking=True)
        self.t_norm.copy_(self._stage_t[slot], non_blocking=True)
        self._ev_used[slot].record(main)
        if next_frame_u8_host is not None:
            self._prefetch((i + 1) % 2, next_frame_u8_host, next_t_host, first=(i == 0))
        self.step()
        self._out_ring[r].copy_(self.out, non_blocking=True)
        self._ev_out[r].record(main)
        self._cs.wait_event(self._ev_out[r])
        with torch.cuda.stream(self._cs):
            self._host_ring[r].copy_(self._out_ring[r], non_blocking=True)
            self._ev_d2h[r].record(self._cs)
        prev = None
        if self._hp_pending is not None:
            self._ev_d2h[self._hp_pending].synchronize()
            prev = self._host_ring[self._hp_pending].clone()
        self._hp_pending = r
        self._hp_i = i + 1
        return prev

    def host_pipeline_end(self):
        """Metrics of the last step (blocks until its download has landed)."""
        if self._hp_pending is None:
            return None
        self._ev_d2h[self._hp_pending].synchronize()
        out = self._host_ring[self._hp_pending].clone()
        self._hp_pending = None
        return out

    def sync_step_counter(self):
        """Align the device-side step counter with the host (call after loading a checkpoint)."""
        _, step_dev = self.opt.device_scalars(self.dev)
        step_dev.fill_(self.host_step)
        self.opt._step_count_host = self.host_step
        self._weights_valid = False          # parameters may have been replaced: re-fold before the next step
