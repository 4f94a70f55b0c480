"""Host-side mirror of the reference `model.py` for the frame-fitting hot path.

Same public surface (reference model.py:303-625): `Generator(**kargs)`, `NeRVBlock(**kargs)` with
`get_equivalent_kernel_bias()` / `switch_to_deploy()`, identical sub-module names, construction order
(so `torch.manual_seed(s)` yields bit-identical initial parameters) and `state_dict()` keys/shapes/dtypes
(so reference checkpoints load here and ours load there).  What differs is below the surface: no
`F.conv2d`, `nn.PixelShuffle`, `nn.SiLU`, `nn.Linear` kernel is ever dispatched — the modules are
parameter containers and every arithmetic step runs in liborepnerv.so (sm_100a) via `engine.NetExecutor`.

Scope: branch_type NeRV_vanilla | ERB (the north-star path) and, folded online the same way,
ACB | RepVGG | DBB | ECB (8f-4, branches.py); every activation of the reference's ActivationLayer; norm 'none',
num_blocks 1, single-resolution head.  Anything else raises NotImplementedError — there is no fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib, branches
from ._lib import check, ptr
from .engine import NetExecutor, pad32, conv_tile_n, _conv_plan, _wgrad_plan

SUPPORTED_BRANCHES = ("NeRV_vanilla", "ERB") + branches.FOLDED_BRANCH_TYPES
# reference model.py:86-117 (ActivationLayer) -> activation code of the kernels (csrc/act.cuh)
ACT_CODES = {"swish": 0, "relu": 1, "leaky": 2, "leaky01": 3, "relu6": 4, "gelu": 5, "softplus": 6, "hardswish": 7,
             "sin": 8}
_ERB_BRANCHES = ("rbr_3x3_branch", "rbr_3x1_branch", "rbr_1x3_branch", "rbr_1x1_3x3_1x1_branch_1x1_1",
                 "rbr_1x1_3x3_1x1_branch_3x3", "rbr_1x1_3x3_1x1_branch_1x1_2")


def _require(cond, what):
    if not cond:
        raise NotImplementedError(
            f"{what} is outside the B200 hot path (supported: branch_type NeRV_vanilla|ERB|ACB|RepVGG|DBB|ECB, act "
            f"{'|'.join(ACT_CODES)}, norm none, num_blocks 1, single_res); no fallback path exists")


def _fire_prune_hooks(conv):
    """torch.nn.utils.prune recomputes `.weight = weight_orig * weight_mask` in a forward-pre-hook of the
    conv module (reference main_eval.py:572-587 relies on that).  We never call the module, so run the
    hooks by hand before reading `.weight`."""
    for hook in conv._forward_pre_hooks.values():
        hook(conv, None)


class _FoldFunction(torch.autograd.Function):
    """ERB online fold (reference model.py:450-516) with its analytic backward, on the tensor-core
    fold plan of the block (csrc/fold_tc.cu, 3xTF32 split precision: fp32-accurate)."""

    @staticmethod
    def forward(ctx, blk, w3x3, b3x3, w1x3, b1x3, w3x1, b3x1, w1, w2, w3):
        lib = _lib.lib()
        st = _lib.stream()
        cout, cin = w3x3.shape[0], w3x3.shape[1]
        plan = blk.fold_plan()
        Kt = torch.empty(cout, 9, cin, dtype=torch.float32, device=w3x3.device)
        bias = torch.empty(cout, dtype=torch.float32, device=w3x3.device)
        plan.fwd(blk, Kt, bias, st)
        K = torch.empty(cout, cin, 3, 3, dtype=torch.float32, device=w3x3.device)
        check(lib.onr_tapmajor_permute(ptr(Kt), ptr(K), cin, cout, 1, st), "onr_tapmajor_permute")
        ctx.plan = plan
        ctx.shapes = [t.shape for t in (w3x3, b3x3, w1x3, b1x3, w3x1, b3x1, w1, w2, w3)]
        return K, bias

    @staticmethod
    def backward(ctx, dK, dbias):
        lib = _lib.lib()
        st = _lib.stream()
        plan = ctx.plan
        cout, cin = plan.cout, plan.cin
        dK = dK.contiguous()
        dbias = dbias.contiguous() if dbias is not None else torch.zeros(cout, device=dK.device)
        dKt = torch.empty(cout, 9, cin, dtype=torch.float32, device=dK.device)
        check(lib.onr_tapmajor_permute(ptr(dK), ptr(dKt), cin, cout, 0, st), "onr_tapmajor_permute")
        grads = [torch.zeros(s, dtype=torch.float32, device=dK.device) for s in ctx.shapes]
        plan.bwd(dKt, dbias, grads, st)
        return (None,) + tuple(grads)


class _BranchSetFoldFunction(torch.autograd.Function):
    """Online fold of an ACB / RepVGG / DBB / ECB branch set into one 3x3 kernel + bias (csrc/fold_branches.cu; the
    reference runs model.py:541-565 un-folded) with its analytic backward.  `tensors` follow branches.branch_slots."""

    @staticmethod
    def forward(ctx, blk, *tensors):
        lib, st = _lib.lib(), _lib.stream()
        dev = tensors[0].device
        K = torch.empty(blk.out_channels, blk.ngf, 3, 3, dtype=torch.float32, device=dev)
        bias = torch.empty(blk.out_channels, dtype=torch.float32, device=dev)
        branches.fold_fwd(lib, blk, K, bias, st)
        ctx.blk = blk
        return K, bias

    @staticmethod
    def backward(ctx, dK, dbias):
        lib, st = _lib.lib(), _lib.stream()
        blk = ctx.blk
        dK = dK.contiguous()
        dbias = dbias.contiguous() if dbias is not None else torch.zeros(blk.out_channels, device=dK.device)
        slots = branches.branch_slots(blk)
        grads = {name: (torch.zeros_like(t) if t.requires_grad else None) for _, name, t in slots}
        branches.fold_bwd(lib, blk, dK, dbias, grads.get, st)
        return (None,) + tuple(grads[name] for _, name, _ in slots)


class _BlockFunction(torch.autograd.Function):
    """conv3x3 + PixelShuffle + SiLU of ONE block on NCHW fp32 tensors (module-boundary path used when a
    NeRVBlock is called on its own; the Generator uses the fused executor instead)."""

    @staticmethod
    def forward(ctx, x, K, bias, stride, cnew, act=0):
        lib = _lib.lib()
        st = _lib.stream()
        B, cin, H, W = x.shape
        s = stride
        cpi, cpo = pad32(cin), pad32(cnew)
        nk = s * s * cpo
        bn, nt = conv_tile_n(nk)
        npad = bn * nt
        bn2, nt2 = conv_tile_n(cpi)
        cpi_rows = bn2 * nt2
        dev = x.device
        bf16 = torch.bfloat16
        train = torch.is_grad_enabled() and (x.requires_grad or K.requires_grad or bias.requires_grad)
        xh = torch.empty(B, H, W, cpi, dtype=bf16, device=dev)
        check(lib.onr_nchw_to_nhwc_bf16(ptr(x.detach().contiguous()), B, cin, H, W, cpi, ptr(xh), st), "to_nhwc")
        wf = torch.empty(9, npad, cpi, dtype=bf16, device=dev)
        wd = torch.empty(9, cpi_rows, nk, dtype=bf16, device=dev)
        bias_p = torch.empty(npad, dtype=torch.float32, device=dev)
        check(lib.onr_pack_weights(ptr(K.detach().contiguous()), ptr(bias.detach().contiguous()), cin, cnew, s,
                                   npad, cpi_rows, ptr(wf), ptr(wd), ptr(bias_p), st), "onr_pack_weights")
        y = torch.empty(B, H * s, W * s, cpo, dtype=bf16, device=dev)
        d = torch.empty_like(y)
        plan = _conv_plan(lib, kind=_lib.CONV_FPROP_TRAIN if act == 0 else _lib.CONV_FPROP_Z, B=B, H=H, W=W,
                          a=ptr(xh), a_cp=cpi, a_s=1, w=ptr(wf), n_rows=npad, n_total=nk, out=ptr(y), out_cp=cpo,
                          out_s=s, out_d=ptr(d) if act == 0 else None, bias_p=ptr(bias_p), dmul=None)
        check(lib.onr_conv_plan_run(plan.handle, st), "onr_conv_plan_run")
        if act != 0:
            check(lib.onr_act_map(ptr(y), ptr(d), B * H * s * W * s, cnew, cpo, act, st), "onr_act_map")
        out = torch.empty(B, cnew, H * s, W * s, dtype=torch.float32, device=dev)
        check(lib.onr_nhwc_bf16_to_nchw(ptr(y), B, cnew, H * s, W * s, cpo, ptr(out), st), "to_nchw")
        ctx.geom = (B, cin, H, W, s, cnew, cpi, cpo, nk, cpi_rows)
        ctx.save_for_backward(xh, d, wd)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.lib()
        st = _lib.stream()
        B, cin, H, W, s, cnew, cpi, cpo, nk, cpi_rows = ctx.geom
        xh, d, wd = ctx.saved_tensors
        dev, bf16 = gout.device, torch.bfloat16
        gy = torch.empty(B, H * s, W * s, cpo, dtype=bf16, device=dev)
        check(lib.onr_nchw_to_nhwc_bf16(ptr(gout.contiguous()), B, cnew, H * s, W * s, cpo, ptr(gy), st), "to_nhwc")
        dz = gy * d                                   # boundary path only: dZ = dY * SiLU'(z)
        dKp = torch.zeros(nk, 9, cpi, dtype=torch.float32, device=dev)
        dbp = torch.zeros(nk, dtype=torch.float32, device=dev)
        wplan = _wgrad_plan(lib, B=B, H=H, W=W, x=ptr(xh), x_cp=cpi, dz=ptr(dz), dz_cp=cpo, s=s,
                            dKp=ptr(dKp), dbias_p=ptr(dbp))
        check(lib.onr_wgrad_plan_run(wplan.handle, st), "onr_wgrad_plan_run")
        dK = torch.empty(cnew * s * s, cin, 3, 3, dtype=torch.float32, device=dev)
        db = torch.empty(cnew * s * s, dtype=torch.float32, device=dev)
        check(lib.onr_unpack_wgrad(ptr(dKp), ptr(dbp), cin, cnew, s, ptr(dK), ptr(db), st), "onr_unpack_wgrad")
        ones = torch.ones(B, H, W, cpi, dtype=bf16, device=dev)
        dxh = torch.empty(B, H, W, cpi, dtype=bf16, device=dev)
        dplan = _conv_plan(lib, kind=_lib.CONV_DGRAD, B=B, H=H, W=W, a=ptr(dz), a_cp=cpo, a_s=s, w=ptr(wd),
                           n_rows=cpi_rows, n_total=cpi, out=ptr(dxh), out_cp=cpi, out_s=1, out_d=None,
                           bias_p=None, dmul=ptr(ones))
        check(lib.onr_conv_plan_run(dplan.handle, st), "onr_conv_plan_run(dgrad)")
        dx = torch.empty(B, cin, H, W, dtype=torch.float32, device=dev)
        check(lib.onr_nhwc_bf16_to_nchw(ptr(dxh), B, cin, H, W, cpi, ptr(dx), st), "to_nchw")
        torch.cuda.current_stream().synchronize()     # plans (TMA descriptors) must outlive the launches
        return dx, dK, db, None, None, None


def activation_module(act_type):
    """reference model.py:86-117 (ActivationLayer) — kept for module traversal / print(model); never called."""
    table = {'relu': lambda: nn.ReLU(True), 'leaky': lambda: nn.LeakyReLU(inplace=True),
             'leaky01': lambda: nn.LeakyReLU(negative_slope=0.1, inplace=True), 'relu6': lambda: nn.ReLU6(inplace=True),
             'gelu': nn.GELU, 'swish': lambda: nn.SiLU(inplace=True), 'softplus': nn.Softplus,
             'hardswish': lambda: nn.Hardswish(inplace=True), 'sin': lambda: torch.sin}
    if act_type not in table:
        raise KeyError(f"Unknown activation function {act_type}.")
    return table[act_type]()


class NeRVBlock(nn.Module):
    """Reference model.py:303-567.  conv3x3(ngf -> new_ngf*stride^2) -> PixelShuffle(stride) -> SiLU."""

    def __init__(self, **kargs):
        super().__init__()
        self.ngf, self.new_ngf, self.stride = kargs['ngf'], kargs['new_ngf'], kargs['stride']
        self.deploy = kargs['deploy']
        self.branch_type = kargs['branch_type']
        _require(self.branch_type in SUPPORTED_BRANCHES, f"branch_type {self.branch_type!r}")
        _require(kargs.get('norm', 'none') == 'none', f"norm {kargs.get('norm')!r}")
        self.act_name = kargs.get('act', 'swish')
        _require(self.act_name in ACT_CODES, f"act {self.act_name!r}")
        # parameter-free modules kept so that print(model) / module traversal look like the reference
        self.up_scale = nn.PixelShuffle(self.stride)
        self.norm = nn.Identity()
        self.act = activation_module(self.act_name)
        self.out_channels = self.new_ngf * self.stride * self.stride
        ci, co = self.ngf, self.out_channels
        if self.deploy:
            self.rbr_reparam = nn.Conv2d(ci, co, (3, 3), 1, 1, bias=True)
        elif self.branch_type == "NeRV_vanilla":
            self.branch = nn.Conv2d(ci, co, (3, 3), 1, 1, bias=kargs.get('bias', True))
            _require(self.branch.bias is not None, "bias=False")
        elif self.branch_type in branches.FOLDED_BRANCH_TYPES:
            branches.create_branches(self, self.branch_type, ci, co)
        else:  # ERB: creation order matches reference model.py:324-343 (same RNG consumption)
            self.rbr_3x3_branch = nn.Conv2d(ci, co, (3, 3), 1, 1)
            self.rbr_3x1_branch = nn.Conv2d(ci, co, (3, 1), 1, (1, 0))
            self.rbr_1x3_branch = nn.Conv2d(ci, co, (1, 3), 1, (0, 1))
            self.rbr_1x1_3x3_1x1_branch_1x1_1 = nn.Conv2d(ci, 2 * ci, (1, 1), 1, 0, bias=False)
            self.rbr_1x1_3x3_1x1_branch_3x3 = nn.Conv2d(2 * ci, co, (3, 3), 1, 1, bias=False)
            self.rbr_1x1_3x3_1x1_branch_1x1_2 = nn.Conv2d(co, co, (1, 1), 1, 0, bias=False)

    # ---- structure queries used by the executor -------------------------------------------------
    def is_erb_train(self):
        return (not self.deploy) and self.branch_type == "ERB" and hasattr(self, "rbr_3x3_branch")

    def fold_kind(self):
        """'erb' (tensor-core fold, fold_tc.cu), 'set' (ACB / RepVGG / DBB / ECB, fold_branches.cu) or None (a single
        convolution: vanilla or deploy state)."""
        if self.is_erb_train():
            return "erb"
        if (not self.deploy) and self.branch_type in branches.FOLDED_BRANCH_TYPES and hasattr(self, "rbr_3x3_branch"):
            return "set"
        return None

    @property
    def act_code(self):
        return ACT_CODES[self.act_name]

    def fold_plan(self):
        """Tensor-core fold plan of this block (created on first use; holds the fold's workspace on the device)."""
        dev = self.rbr_3x3_branch.weight.device
        plan = self.__dict__.get("_fold_plan_obj")
        if plan is None or plan.work.device != dev:
            from .engine import FoldPlan
            plan = FoldPlan(_lib.lib(), self.ngf, self.out_channels, True, dev)
            self.__dict__["_fold_plan_obj"] = plan
        return plan

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k != "_fold_plan_obj":                    # plans hold raw device pointers: never copied
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def single_conv_name(self):
        return "rbr_reparam" if (self.deploy or not hasattr(self, "branch")) else "branch"

    def single_conv(self):
        conv = getattr(self, self.single_conv_name())
        _fire_prune_hooks(conv)
        return conv

    # ---- reference API -----------------------------------------------------------------------
    def get_equivalent_kernel_bias(self):
        """Reference model.py:450-478; differentiable w.r.t. the nine branch tensors."""
        kind = self.fold_kind()
        if kind == "set":
            # the reference has no fold for these sets (AttributeError); here they fold like ERB
            return _BranchSetFoldFunction.apply(self, *[t for _, _, t in branches.branch_slots(self)])
        if kind is None:
            raise AttributeError("get_equivalent_kernel_bias needs the training branches")
        b = self
        return _FoldFunction.apply(
            self, b.rbr_3x3_branch.weight, b.rbr_3x3_branch.bias, b.rbr_1x3_branch.weight, b.rbr_1x3_branch.bias,
            b.rbr_3x1_branch.weight, b.rbr_3x1_branch.bias, b.rbr_1x1_3x3_1x1_branch_1x1_1.weight,
            b.rbr_1x1_3x3_1x1_branch_3x3.weight, b.rbr_1x1_3x3_1x1_branch_1x1_2.weight)

    def switch_to_deploy(self):
        """Reference model.py:395-448: fold once into `rbr_reparam`, drop the branches."""
        if getattr(self, 'deploy', False) or not hasattr(self, 'rbr_3x3_branch'):
            if hasattr(self, 'rbr_reparam'):
                self.deploy = True
            return
        with torch.no_grad():
            kernel, bias = self.get_equivalent_kernel_bias()
        if not hasattr(self, 'rbr_reparam'):
            self.rbr_reparam = nn.Conv2d(self.ngf, self.out_channels, (3, 3), 1, 1, bias=True)
        self.rbr_reparam.weight.data = kernel
        self.rbr_reparam.bias.data = bias
        for name in _ERB_BRANCHES + ("branch",) + branches.SET_BRANCH_MODULES:
            if hasattr(self, name):
                self.__delattr__(name)
        self.__dict__.pop("_fold_plan_obj", None)
        self.deploy = True

    def forward(self, x):
        """Reference model.py:518-567 on NCHW fp32 tensors (module-boundary path)."""
        if self.fold_kind() is not None:
            K, b = self.get_equivalent_kernel_bias()
        else:
            conv = self.single_conv()
            K, b = conv.weight, conv.bias
        return _BlockFunction.apply(x, K, b, self.stride, self.new_ngf, self.act_code)


class _FnModule(nn.Module):
    """nn.Sequential slot for a bare function (the reference puts torch.sin itself into the stem's Sequential, which
    torch >= 1.9 rejects; the index layout stem.0 / stem.2 is what matters for the state dict)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


class _GeneratorFunction(torch.autograd.Function):
    """Whole-decoder autograd node: forward = executor.forward, backward = executor.backward.

    The image is written straight into a fresh tensor (no copy).  The backward writes the parameter gradients into
    the Generator's PERSISTENT flat gradient buffer and binds `p.grad` to views of it, so the reference loop
    (`optimizer.zero_grad(); loss.backward(); optimizer.step()`, main_train.py:248-250) costs one memset and no
    per-step allocation; only a second backward without zero_grad in between (gradient accumulation) takes the
    generic path that hands temporaries to autograd."""

    @staticmethod
    def forward(ctx, gen, ex, embed, *params):
        img = torch.empty(ex.B, 3, ex.H, ex.W, dtype=torch.float32, device=ex.dev)
        early = ex.head_stages[:-1]                   # multi-resolution heads (sin_res=False): one image per stage
        outs = {l: torch.empty(ex.B, 3, ex.geoms[l].ho, ex.geoms[l].wo, dtype=torch.float32, device=ex.dev)
                for l in early}
        ex.forward(embed=embed, out=img, outs=outs)
        ctx.gen, ctx.ex, ctx.early = gen, ex, early
        return tuple(outs[l] for l in early) + (img,)

    @staticmethod
    def backward(ctx, *gimgs_all):
        gen, ex = ctx.gen, ctx.ex
        gimg = gimgs_all[-1]
        gimgs = {l: g.contiguous() for l, g in zip(ctx.early, gimgs_all[:-1])} or None
        named = list(gen.named_parameters())
        pg = gen.persistent_grads()
        owned = [p.grad is pg[n] for n, p in named]
        # (SeqConv3x3.mask is a Parameter that never requires a gradient: it takes no part in the binding)
        fast = all((not p.requires_grad) or p.grad is None or o for (n, p), o in zip(named, owned))
        if fast and (gen._grads_clean or not any(owned)):
            if not gen._grads_clean:
                pg["__flat__"].zero_()
            ex.backward(gimg.contiguous(), pg, gimgs=gimgs)
            gen._grads_clean = False
            for n, p in named:
                if p.requires_grad:
                    p.grad = pg[n]
            return (None, None, None) + (None,) * len(named)
        grads = gen.alloc_grads()
        ex.backward(gimg.contiguous(), grads, gimgs=gimgs)
        out = [grads[n] if p.requires_grad else None for n, p in named]
        return (None, None, None) + tuple(out)


class Generator(nn.Module):
    """Reference model.py:571-625."""

    def __init__(self, **kargs):
        super().__init__()
        stem_dim, stem_num = [int(x) for x in kargs['stem_dim_num'].split('_')]
        self.fc_h, self.fc_w, self.fc_dim = [int(x) for x in kargs['fc_hw_dim'].split('_')]
        _require(stem_num == 1, f"stem_dim_num with {stem_num} hidden layers")
        self.act_name = kargs.get('act', 'swish')
        _require(self.act_name in ACT_CODES, f"act {self.act_name!r}")
        _require(kargs.get('num_blocks', 1) == 1, "num_blocks > 1")
        self.sin_res = bool(kargs.get('sin_res', True))
        _require(kargs.get('bias', True), "bias=False")
        # reference MLP(): Linear, act, Linear, act with one shared activation module (model.py:184-188)
        act_fn = activation_module(self.act_name)
        if not isinstance(act_fn, nn.Module):          # 'sin' is the bare torch.sin in the reference (model.py:104)
            act_fn = _FnModule(act_fn)
        self.stem = nn.Sequential(nn.Linear(kargs['embed_length'], stem_dim), act_fn,
                                  nn.Linear(stem_dim, self.fc_h * self.fc_w * self.fc_dim), act_fn)
        self.layers, self.head_layers = [nn.ModuleList() for _ in range(2)]
        ngf = self.fc_dim
        strides = list(kargs['stride_list'])
        for i, stride in enumerate(strides):
            if i == 0:
                new_ngf = int(ngf * kargs['expansion'])
            else:
                new_ngf = max(ngf // (1 if stride == 1 else kargs['reduction']), kargs['lower_width'])
            self.layers.append(NeRVBlock(ngf=ngf, new_ngf=new_ngf, stride=stride, bias=kargs['bias'],
                                         norm=kargs['norm'], act=kargs['act'], deploy=kargs['deploy'],
                                         conv_type=kargs.get('conv_type', 'conv'),
                                         branch_type=kargs['branch_type']))
            ngf = new_ngf
            # reference model.py:598-608: one RGB head on the last stage (sin_res) or on every stage
            self.head_layers.append(nn.Conv2d(ngf, 3, 1, 1, bias=kargs['bias'])
                                    if (i == len(strides) - 1 or not self.sin_res) else None)
        self.sigmoid = kargs['sigmoid']
        self._executors = {}
        self._pgrads = None
        self._grads_clean = False

    # ---- helpers for the executor ---------------------------------------------------------------
    @property
    def act_code(self):
        return ACT_CODES[self.act_name]

    def head_name(self):
        return f"head_layers.{len(self.layers) - 1}"

    def head_conv(self):
        return self.head_layers[len(self.layers) - 1]

    def executor(self, batch, train):
        key = (batch, bool(train), tuple(b.fold_kind() for b in self.layers),
               str(next(self.parameters()).device))
        ex = self._executors.get(key)
        if ex is None:
            ex = NetExecutor(self, batch, train)
            self._executors = {key: ex}      # one live executor: buffers are large (hundreds of MB)
        return ex

    def alloc_grads(self):
        """Zeroed fp32 gradient tensors, one per parameter, carved from a single flat buffer (each slice starts
        on a 256-byte boundary so the optimizer and the kernels can use 128-bit accesses)."""
        named = list(self.named_parameters())
        align = 64
        total = sum(-(-p.numel() // align) * align for _, p in named)
        flat = torch.zeros(total, dtype=torch.float32, device=named[0][1].device)
        grads, offsets, off = {}, {}, 0
        for n, p in named:
            grads[n] = flat[off:off + p.numel()].view_as(p)
            offsets[n] = (off, p.numel())
            off += -(-p.numel() // align) * align
        grads["__flat__"] = flat
        grads["__offsets__"] = offsets
        return grads

    def persistent_grads(self):
        """The flat gradient buffer `loss.backward()` writes into (allocated once per device; see
        _GeneratorFunction).  `_grads_clean` tracks whether it is known to be all zeros."""
        dev = next(self.parameters()).device
        if self._pgrads is None or self._pgrads["__flat__"].device != dev:
            self._pgrads = self.alloc_grads()
            self._grads_clean = True
            for n, p in self.named_parameters():
                p._onr_grad_owner, p._onr_name = self, n
        return self._pgrads

    def zero_persistent_grads(self):
        if self._pgrads is not None and not self._grads_clean:
            self._pgrads["__flat__"].zero_()
            self._grads_clean = True

    def __deepcopy__(self, memo):
        # executors hold raw device pointers / TMA descriptors of THIS model's buffers: never copy them
        # (reference main_train.py:332 deep-copies the model every epoch for the deploy checkpoint)
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_executors":
                new.__dict__[k] = {}
            elif k == "_pgrads":
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._grads_clean = False
        return new

    def forward(self, input):
        """input: embedding [B, 2*levels] (reference model.py:611-625). Returns the list of images, one per head
        stage ([img[B,3,H,W]] with sin_res)."""
        B = input.size(0)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        ex = self.executor(B, needs_grad)
        if needs_grad:
            params = [p for _, p in self.named_parameters()]
            return list(_GeneratorFunction.apply(self, ex, input.detach(), *params))
        with torch.no_grad():
            return ex.decode(input)
