"""Fused multi-tensor Adam behind the torch.optim.Optimizer interface.

Drop-in for `optim.Adam(model.parameters(), betas=(args.beta, 0.999))` (reference main_train.py:196,
:248-250): same hyper-parameters, same per-parameter state keys (`step`, `exp_avg`, `exp_avg_sq`), hence
the same `optimizer.state_dict()` layout in checkpoints (main_train.py:300).  `step()` is ONE launch of
onr_adam_multi over all parameter tensors; gradients can be averaged (grad_scale) and re-zeroed in the
same pass.
"""
import torch

from . import _lib
from ._lib import check, ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam mirrors the reference configuration: no weight decay / amsgrad")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self._step_count_host = 0
        self._tables = {}
        self._ring = None
        self._ring_pos = 0
        self.grad_scale = 1.0
        self.fused_zero_grad = False

    # ---- state in torch.optim.Adam's layout ------------------------------------------------------
    def _ensure_state(self, p):
        st = self.state[p]
        if 'exp_avg' not in st:
            st['step'] = torch.tensor(float(self._step_count_host))
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def state_dict(self):
        for st in self.state.values():
            if 'step' in st:
                st['step'] = torch.tensor(float(self._step_count_host))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [float(st['step']) for st in self.state.values() if 'step' in st]
        self._step_count_host = int(max(steps)) if steps else 0
        self._tables = {}

    def zero_grad(self, set_to_none=True):
        """torch.optim.Optimizer.zero_grad, except that gradients living in a Generator's persistent flat buffer
        (bound to `p.grad` by `loss.backward()`, see model._GeneratorFunction) stay bound and are cleared with ONE
        memset instead of being dropped and re-allocated every step (reference main_train.py:248)."""
        owners = []
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                owner = getattr(p, "_onr_grad_owner", None)
                if owner is not None and owner._pgrads is not None and \
                        p.grad.data_ptr() == owner._pgrads["__flat__"].data_ptr() + 4 * owner._pgrads["__offsets__"].get(
                            getattr(p, "_onr_name", ""), (-1, 0))[0]:
                    if all(owner is not o for o in owners):
                        owners.append(owner)
                elif set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()
        for owner in owners:
            owner.zero_persistent_grads()

    # ---- table of raw pointers -------------------------------------------------------------------
    def _table(self, gi, group, subset=None):
        params = [p for p in (group['params'] if subset is None else subset) if p.grad is not None]
        if not params:
            return None
        if subset is not None:
            gi = (gi, tuple(id(p) for p in subset))
        rows = []
        for p in params:
            if p.grad.dtype != torch.float32 or p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("FusedAdam needs fp32 CUDA parameters and gradients")
            st = self._ensure_state(p)
            rows.append((p.data_ptr(), p.grad.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                         p.numel()))
        key = tuple(rows)
        cached = self._tables.get(gi)
        if cached is None or cached[0] != key:
            dev = params[0].device
            lib = _lib.lib()
            per_block, per_call = int(lib.onr_adam_block_elems()), int(lib.onr_adam_max_tensors())
            calls = []
            for i in range(0, len(rows), per_call):
                part = rows[i:i + per_call]
                table = torch.tensor(part, dtype=torch.int64).to(dev)
                calls.append((table, sum(-(-r[4] // per_block) for r in part), len(part)))
            cached = (key, calls, dev)
            self._tables[gi] = cached
        return cached

    def device_scalars(self, dev):
        """(lr_dev fp32[1], step_dev int32[1]) read by the kernel."""
        if getattr(self, "_lr_dev", None) is None or self._lr_dev.device != dev:
            self._lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._hyp_dev = torch.zeros(2, dtype=torch.float32, device=dev)   # kernel scratch (bias corrections)
        return self._lr_dev, self._step_dev

    def _upload(self, lr, step, dev):
        lr_dev, step_dev = self.device_scalars(dev)
        if self._ring is None:
            # pinned staging ring so an async copy never reads a slot the host has already rewritten
            self._ring = (torch.zeros(512, dtype=torch.float32).pin_memory(),
                          torch.zeros(512, dtype=torch.int32).pin_memory())
        i = self._ring_pos
        self._ring_pos = (i + 1) % 512
        if self._ring_pos == 0:
            torch.cuda.current_stream().synchronize()
        self._ring[0][i] = lr
        self._ring[1][i] = step
        lr_dev.copy_(self._ring[0][i:i + 1], non_blocking=True)
        step_dev.copy_(self._ring[1][i:i + 1], non_blocking=True)

    @torch.no_grad()
    def step_params(self, params):
        """Adam update of a subset of the (single) parameter group, reading the learning rate and step count the
        caller has already advanced on the device (onr_sched_tick).  Lets a trainer update a block's parameters as
        soon as its gradients exist; the host step count is the caller's to advance once per optimisation step."""
        if len(self.param_groups) != 1:
            raise RuntimeError("step_params needs a single parameter group")
        lib, group = _lib.lib(), self.param_groups[0]
        cached = self._table(0, group, subset=list(params))
        if cached is None:
            return
        _, calls, dev = cached
        lr_dev, step_dev = self.device_scalars(dev)
        b1, b2 = group['betas']
        for table, total_blocks, n in calls:
            check(lib.onr_adam_multi(ptr(table), n, total_blocks, ptr(lr_dev), ptr(step_dev),
                                     ptr(self._hyp_dev), b1, b2,
                                     group['eps'], float(self.grad_scale), 1 if self.fused_zero_grad else 0,
                                     _lib.stream()), "onr_adam_multi")

    @torch.no_grad()
    def step(self, closure=None, device_schedule=False):
        """One Adam update.  With device_schedule=True the caller has already advanced the device-side
        step counter / learning rate (onr_sched_tick), e.g. inside a captured CUDA graph."""
        loss = closure() if closure is not None else None
        lib = _lib.lib()
        self._step_count_host += 1
        for gi, group in enumerate(self.param_groups):
            cached = self._table(gi, group)
            if cached is None:
                continue
            _, calls, dev = cached
            if not device_schedule:
                self._upload(float(group['lr']), self._step_count_host, dev)
            lr_dev, step_dev = self.device_scalars(dev)
            b1, b2 = group['betas']
            for table, total_blocks, n in calls:
                check(lib.onr_adam_multi(ptr(table), n, total_blocks, ptr(lr_dev), ptr(step_dev),
                                         ptr(self._hyp_dev), b1, b2,
                                         group['eps'], float(self.grad_scale), 1 if self.fused_zero_grad else 0,
                                         _lib.stream()), "onr_adam_multi")
            if not torch.cuda.is_current_stream_capturing():
                # the kernel writes through raw pointers: tell autograd / the decode-side operand cache
                for p in group['params']:
                    torch.autograd.graph.increment_version(p)
        return loss
