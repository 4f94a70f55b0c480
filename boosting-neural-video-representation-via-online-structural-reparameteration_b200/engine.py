"""Device-side executor of the frame-fitting hot path.

One `NetExecutor` owns, for a `Generator` at a fixed batch size, every persistent HBM buffer (NHWC bf16
activations, SiLU' maps, dZ maps, folded kernels and their packed bf16 operand copies, gradient
staging) and the TMA-descriptor plans of the tcgen05 convolution kernels, and sequences the kernels of
liborepnerv.so for

    forward  : PE + stem -> per block [ERB fold -> pack -> conv3x3+PixelShuffle+SiLU] -> RGB head
    backward : head -> per block [wgrad, dgrad (x SiLU'), unpack, fold backward] -> stem

which is reference main_train.py:238 (`model(embed_input)`) and :249 (`loss_sum.backward()`) for
model.py:611-625 / :518-567.  PyTorch only provides the allocations and the stream.
"""
import ctypes as C
import math

import os

import torch
import torch.distributed

from . import _lib, branches
from ._lib import ConvDesc, WgradDesc, check, ptr


def pad32(c):
    return (c + 31) // 32 * 32


def conv_tile_n(n_total):
    bn, nt = C.c_int(0), C.c_int(0)
    check(_lib.load().onr_conv_tile_n(n_total, C.byref(bn), C.byref(nt)), "onr_conv_tile_n")
    return bn.value, nt.value


class BlockGeom:
    """Shapes of one NeRVBlock (reference model.py:303-343) in the padded NHWC / implicit-GEMM world."""

    def __init__(self, cin, cnew, s, h, w):
        self.cin, self.cnew, self.s, self.h, self.w = cin, cnew, s, h, w
        self.cout = cnew * s * s                 # reference conv output channels (PixelShuffle order)
        self.cpi, self.cpo = pad32(cin), pad32(cnew)
        self.nk = s * s * self.cpo               # GEMM N of fprop / K of dgrad, n' = (i*s+j)*Cpo + c
        bn, nt = conv_tile_n(self.nk)
        self.npad = bn * nt                      # rows of the packed fprop weights
        bn2, nt2 = conv_tile_n(self.cpi)
        self.cpi_rows = bn2 * nt2                # rows of the packed dgrad weights
        self.ho, self.wo = h * s, w * s


class _Plan:
    """RAII wrapper around onr_conv_plan / onr_wgrad_plan handles."""

    def __init__(self, handle, destroy):
        self.handle, self._destroy = handle, destroy

    def __del__(self):
        try:
            if self.handle:
                self._destroy(self.handle)
        except Exception:
            pass


def _conv_plan(lib, **kw):
    d = ConvDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    h = C.c_void_p()
    check(lib.onr_conv_plan_create(C.byref(h), C.byref(d)), "onr_conv_plan_create")
    return _Plan(h, lib.onr_conv_plan_destroy)


class FoldPlan:
    """Tensor-core ERB fold of one (Cin, Cout) block (csrc/fold_tc.cu): owns the workspace and the plan handle."""

    def __init__(self, lib, cin, cout, train, device):
        self.lib, self.cin, self.cout, self.train = lib, cin, cout, train
        nbytes = lib.onr_fold_workspace_bytes(cin, cout, 1 if train else 0)
        # filled with 0xFF (= NaN as fp32) once: a kernel that ever reads an operand element nobody wrote shows up as NaN
        # instead of as a result that depends on what the allocator handed out
        self.work = torch.full((nbytes + 1024,), 0xFF, dtype=torch.uint8, device=device)
        base = (self.work.data_ptr() + 1023) // 1024 * 1024
        h = C.c_void_p()
        check(lib.onr_fold_plan_create(C.byref(h), cin, cout, C.c_void_p(base), 1 if train else 0),
              "onr_fold_plan_create")
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                self.lib.onr_fold_plan_destroy(self.handle)
        except Exception:
            pass

    def fwd(self, blk, Kt, bias, st):
        b = blk
        check(self.lib.onr_fold_plan_fwd(
            self.handle, ptr(b.rbr_3x3_branch.weight), ptr(b.rbr_3x3_branch.bias),
            ptr(b.rbr_1x3_branch.weight), ptr(b.rbr_1x3_branch.bias),
            ptr(b.rbr_3x1_branch.weight), ptr(b.rbr_3x1_branch.bias),
            ptr(b.rbr_1x1_3x3_1x1_branch_1x1_1.weight), ptr(b.rbr_1x1_3x3_1x1_branch_3x3.weight),
            ptr(b.rbr_1x1_3x3_1x1_branch_1x1_2.weight), ptr(Kt), ptr(bias), st), "onr_fold_plan_fwd")

    def bwd(self, dKt, dbias, g, st):
        """g: the nine branch gradient tensors in the order (3x3 w, b, 1x3 w, b, 3x1 w, b, w1, w2, w3)."""
        check(self.lib.onr_fold_plan_bwd(self.handle, ptr(dKt), ptr(dbias), *[ptr(t) for t in g], st),
              "onr_fold_plan_bwd")


def _wgrad_plan(lib, **kw):
    d = WgradDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    h = C.c_void_p()
    check(lib.onr_wgrad_plan_create(C.byref(h), C.byref(d)), "onr_wgrad_plan_create")
    return _Plan(h, lib.onr_wgrad_plan_destroy)


class NetExecutor:
    def __init__(self, gen, batch, train):
        self.lib = _lib.lib()
        self.gen, self.B, self.train = gen, batch, train
        dev = next(gen.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("orepnerv Generator must live on a CUDA (sm_100a) device; no CPU path exists")
        self.dev = dev
        B = batch
        bf16, f32 = torch.bfloat16, torch.float32

        def zeros(*shape, dtype=f32):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        # ---- geometry -------------------------------------------------------------------------
        self.geoms = []
        h, w, c = gen.fc_h, gen.fc_w, gen.fc_dim
        for blk in gen.layers:
            g = BlockGeom(blk.ngf, blk.new_ngf, blk.stride, h, w)
            assert g.cin == c
            self.geoms.append(g)
            h, w, c = g.ho, g.wo, g.cnew
        self.H, self.W, self.C_last = h, w, c
        L = len(self.geoms)
        self.L = L
        # stages that carry an RGB head: the last one always; every stage with sin_res=False (reference
        # model.py:598-608).  The head kernels keep the 3 x C weights in registers / shared memory: C <= 128 per stage.
        self.head_stages = [l for l in range(L) if gen.head_layers[l] is not None]
        assert self.head_stages and self.head_stages[-1] == L - 1
        # (wider early stages — sin_res=False with the reference's default widths — take the plain wide-head kernels,
        # at most 1024 channels)
        for l in self.head_stages:
            limit = 128 if l == L - 1 else 1024
            if pad32(self.geoms[l].cnew) > limit:
                raise NotImplementedError(
                    f"RGB head on a stage of {self.geoms[l].cnew} channels: the head kernels take at most {limit} here")
        self.multi = len(self.head_stages) > 1

        # ---- stem ------------------------------------------------------------------------------
        lin1, lin2 = gen.stem[0], gen.stem[2]
        self.E, self.hid = lin1.in_features, lin1.out_features
        self.embed = zeros(B, self.E)
        self.pre1, self.h1, self.dh1 = zeros(B, self.hid), zeros(B, self.hid), zeros(B, self.hid)

        # Training, single image: the LAST block stores only its bf16 pre-activation z (ONR_CONV_FPROP_Z) and the head
        # kernels evaluate SiLU(z) / SiLU'(z) themselves — its output feeds nothing else, so the SiLU' map (177 MB at
        # 720p) is never written and the block's epilogue, which set the pace of that kernel, shrinks to bias + convert.
        self._head_fused = os.environ.get("ONR_HEAD_FUSED", "1") != "0"
        # activation code of the run (csrc/act.cuh).  swish (0) is fused into the convolution epilogue; any other
        # activation runs the blocks in pre-activation mode (ONR_CONV_FPROP_Z) followed by onr_act_map, and the
        # swish-only shortcuts (z-only last block, fused decode head) are off
        self.act = int(getattr(gen, "act_code", 0))
        swish = self.act == 0
        self._last_z = (swish and train and B == 1 and (self.H * self.W) % 4 == 0 and self._head_fused
                        and os.environ.get("ONR_LAST_Z", "1") != "0" and os.environ.get("ONR_HEAD_STREAM", "1") != "0")
        # Decode with the RGB head fused into the last block's epilogue (ONR_CONV_FPROP_HEAD; ONR_DECODE_FUSED=1): the last
        # activation (177 MB at 720p) is never written, only the image is.  Measured on B200 it is NOT faster — the fused
        # kernel needs one 96-wide N tile per sub-pixel and takes 0.221 ms against 0.177 + 0.044 ms for the 128-wide
        # convolution plus the streaming head kernel (2853 vs 2890 frames/s) — so it is an option that saves memory,
        # not the default.
        self._decode_fused = (swish and (not train) and pad32(self.C_last) <= 256
                              and os.environ.get("ONR_DECODE_FUSED", "0") == "1")
        # ---- activations: x[l] is the input of block l (x[L] = last block output, or its pre-activation) ----
        self.x, self.d, self.dz = [], [], []
        for l in range(L + 1):
            if l == 0:
                hh, ww, cp = gen.fc_h, gen.fc_w, self.geoms[0].cpi
            else:
                g = self.geoms[l - 1]
                hh, ww, cp = g.ho, g.wo, g.cpo
            self.x.append(None if (l == L and self._decode_fused) else zeros(B, hh, ww, cp, dtype=bf16))
            # x[0]/d[0] come from the stem kernel which always writes both
            self.d.append(zeros(B, hh, ww, cp, dtype=bf16) if ((train and not (l == L and self._last_z)) or l == 0)
                          else None)
            self.dz.append(zeros(B, hh, ww, cp, dtype=bf16) if train else None)
        self._img_static = zeros(B, 3, self.H, self.W)      # CUDA-graph paths always decode into this one
        self.img = self._img_static
        # multi-resolution heads: image (and, in training, the head's dz contribution) of every earlier head stage
        self._img_stage = {l: zeros(B, 3, self.geoms[l].ho, self.geoms[l].wo) for l in self.head_stages[:-1]}
        self.imgs = dict(self._img_stage)
        self._dz_head = ({l: zeros(B, self.geoms[l].ho, self.geoms[l].wo, self.geoms[l].cpo, dtype=bf16)
                          for l in self.head_stages[:-1]} if train else {})

        # ---- per block weights / gradient staging ------------------------------------------------
        self.K, self.bias, self.T, self.wf, self.wd, self.bias_p = [], [], [], [], [], []
        # ERB fold on the tensor cores (3xTF32, csrc/fold_tc.cu) unless ONR_FOLD_SIMT=1 asks for round 1's fp32 SIMT
        # contractions (kept for A/B timing and as a device-side cross-check)
        self._fold_tc = os.environ.get("ONR_FOLD_SIMT", "0") != "1"
        self.fold = []
        self.dKp, self.dbias_p, self.dK, self.dbias, self.dT, self.dKb = [], [], [], [], [], []
        # the wgrad kernels accumulate (red.global.add) into dKp / dbias_p: all of them live in one pool that a
        # single memset clears per step
        pool_off, total = [], 0
        for g in self.geoms:
            a = total
            b = a + -(-(g.nk * 9 * g.cpi) // 64) * 64
            total = b + -(-g.nk // 64) * 64
            pool_off.append((a, b))
        self._wgrad_pool = zeros(total) if train else None
        for (g, blk), (off_k, off_b) in zip(zip(self.geoms, gen.layers), pool_off):
            kind = blk.fold_kind()
            erb = kind == "erb"
            # folded kernel: tap-major [Cout][9][Cin] on the tensor-core path, OIHW on the SIMT path (ONR_FOLD_SIMT=1)
            # and for the ACB / RepVGG / DBB / ECB branch sets (fold_branches.cu)
            self.K.append(zeros(g.cout, g.cin, 3, 3) if kind else None)
            self.bias.append(zeros(g.cout) if kind else None)
            self.T.append(zeros(g.cout, g.cin, 3, 3) if (erb and not self._fold_tc) else None)
            self.fold.append(FoldPlan(self.lib, g.cin, g.cout, train, dev) if (erb and self._fold_tc) else None)
            self.wf.append(zeros(9, g.npad, g.cpi, dtype=bf16))
            self.wd.append(zeros(9, g.cpi_rows, g.nk, dtype=bf16) if train else None)
            self.bias_p.append(zeros(g.npad))
            if train:
                self.dKp.append(self._wgrad_pool[off_k:off_k + g.nk * 9 * g.cpi].view(g.nk, 9, g.cpi))
                self.dbias_p.append(self._wgrad_pool[off_b:off_b + g.nk])
                # dK | dbias of a block live in ONE buffer: it is the block's gradient-exchange bucket (SURVEY.md 8e
                # option 2: all-reduce the folded-kernel gradient, then run the linear fold backward on every rank)
                nK = g.cout * g.cin * 9
                bucket = zeros(nK + g.cout)
                self.dKb.append(bucket)
                self.dK.append(bucket[:nK].view(g.cout, g.cin, 3, 3))
                self.dbias.append(bucket[nK:])
                self.dT.append(zeros(g.cout, g.cin, 3, 3) if (erb and not self._fold_tc) else None)
            else:
                for lst in (self.dKp, self.dbias_p, self.dK, self.dbias, self.dT, self.dKb):
                    lst.append(None)

        # ---- plans -----------------------------------------------------------------------------
        lib = self.lib
        self.fprop, self.dgrad, self.wgrad = [], [], []
        for l, g in enumerate(self.geoms):
            last_z = self._last_z and l == L - 1
            kind = _lib.CONV_FPROP_Z if (last_z or not swish) else (_lib.CONV_FPROP_TRAIN if train
                                                                     else _lib.CONV_FPROP_INFER)
            extra = {}
            if self._decode_fused and l == L - 1:
                head = gen.head_conv()
                kind = _lib.CONV_FPROP_HEAD
                extra = dict(head_w=ptr(head.weight), head_b=ptr(head.bias), head_c=self.C_last,
                             use_sigmoid=1 if gen.sigmoid else 0, img=ptr(self.img))
            self.fprop.append(_conv_plan(
                lib, kind=kind, B=B, H=g.h, W=g.w, a=ptr(self.x[l]), a_cp=g.cpi, a_s=1,
                w=ptr(self.wf[l]), n_rows=g.npad, n_total=g.nk,
                out=ptr(self.x[l + 1]), out_cp=g.cpo, out_s=g.s,
                out_d=ptr(self.d[l + 1]) if (train and swish and not last_z) else None, bias_p=ptr(self.bias_p[l]),
                dmul=None,
                **extra))
            if train:
                self.dgrad.append(_conv_plan(
                    lib, kind=_lib.CONV_DGRAD, B=B, H=g.h, W=g.w,
                    a=ptr(self.dz[l + 1]), a_cp=g.cpo, a_s=g.s,
                    w=ptr(self.wd[l]), n_rows=g.cpi_rows, n_total=g.cpi,
                    out=ptr(self.dz[l]), out_cp=g.cpi, out_s=1, out_d=None, bias_p=None,
                    dmul=ptr(self.d[l])))
                self.wgrad.append(_wgrad_plan(
                    lib, B=B, H=g.h, W=g.w, x=ptr(self.x[l]), x_cp=g.cpi,
                    dz=ptr(self.dz[l + 1]), dz_cp=g.cpo, s=g.s,
                    dKp=ptr(self.dKp[l]), dbias_p=ptr(self.dbias_p[l])))
        self._wgrad_on_side = os.environ.get("ONR_WGRAD_SIDE", "1") != "0"
        self._fold_chain = int(os.environ.get("ONR_FOLD_CHAIN", "0"))

    # ------------------------------------------------------------------------------------- helpers
    def _block_kernel(self, l):
        """(K, bias, tap_major) of block l, folding the ERB branches on the device when needed."""
        blk = self.gen.layers[l]
        g = self.geoms[l]
        st = _lib.stream()
        if blk.fold_kind() == "set":
            branches.fold_fwd(self.lib, blk, self.K[l], self.bias[l], st)
            return self.K[l], self.bias[l], False
        if blk.is_erb_train():
            if self._fold_tc:
                self.fold[l].fwd(blk, self.K[l], self.bias[l], st)
                return self.K[l], self.bias[l], True
            b = blk
            check(self.lib.onr_erb_fold_fwd(
                ptr(b.rbr_3x3_branch.weight), ptr(b.rbr_3x3_branch.bias),
                ptr(b.rbr_1x3_branch.weight), ptr(b.rbr_1x3_branch.bias),
                ptr(b.rbr_3x1_branch.weight), ptr(b.rbr_3x1_branch.bias),
                ptr(b.rbr_1x1_3x3_1x1_branch_1x1_1.weight), ptr(b.rbr_1x1_3x3_1x1_branch_3x3.weight),
                ptr(b.rbr_1x1_3x3_1x1_branch_1x1_2.weight),
                g.cin, g.cout, ptr(self.K[l]), ptr(self.bias[l]), ptr(self.T[l]), st), "onr_erb_fold_fwd")
            return self.K[l], self.bias[l], False
        conv = blk.single_conv()
        return conv.weight.detach(), conv.bias.detach(), False

    def _side_streams(self):
        if getattr(self, "_side", None) is None:
            self._side = [torch.cuda.Stream(device=self.dev) for _ in range(self.L)]
        return self._side

    def _weights_key(self):
        """Identity + version of every tensor the packed operands depend on (decode-time cache key)."""
        key = []
        for blk in self.gen.layers:
            if blk.fold_kind() == "set":
                ts = [t for _, _, t in branches.branch_slots(blk)]
            elif blk.is_erb_train():
                ts = [getattr(blk, n).weight for n in ("rbr_3x3_branch", "rbr_1x3_branch", "rbr_3x1_branch",
                      "rbr_1x1_3x3_1x1_branch_1x1_1", "rbr_1x1_3x3_1x1_branch_3x3", "rbr_1x1_3x3_1x1_branch_1x1_2")]
                ts += [blk.rbr_3x3_branch.bias, blk.rbr_1x3_branch.bias, blk.rbr_3x1_branch.bias]
            else:
                conv = getattr(blk, blk.single_conv_name())
                ts = ([conv.weight_orig, conv.weight_mask] if hasattr(conv, "weight_orig") else [conv.weight])
                ts.append(conv.bias)
            key += [(t.data_ptr(), t._version) for t in ts]
        for lin in (self.gen.stem[0], self.gen.stem[2]):
            # pruned stem layers (main_eval.py:572-587) carry weight_orig / weight_mask; `.weight` is derived
            ts = [lin.weight_orig, lin.weight_mask] if hasattr(lin, "weight_orig") else [lin.weight]
            key += [(t.data_ptr(), t._version) for t in ts]
        # FrameFitter updates parameters through raw pointers inside a CUDA graph (no per-tensor version bump): it
        # advances this counter instead
        key.append(getattr(self.gen, "_weights_epoch", 0))
        return tuple(key)

    def refresh_weights(self):
        """Fold (ERB) and pack every block kernel into the bf16 operand layouts.

        Each block's fold runs on its own side stream (they are independent, small fp32 GEMMs that leave
        most SMs idle) so the five folds overlap each other and the stem / earlier blocks' convolutions.
        Returns one event per block; the caller makes the main stream wait on event l before conv l."""
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        events = []
        chain = self._fold_chain if self.train else 0
        for l, g in enumerate(self.geoms):
            side = self._side_streams()[l]
            side.wait_event(fork)
            if l > 0 and chain == 1:
                side.wait_event(events[l - 1])      # one fold after the other, block 0 first
            elif l > 0 and chain == 2:
                side.wait_event(events[0])          # block 0 (which gates the first convolution) alone, then the rest
            with torch.cuda.stream(side):
                self.fold_pack_block(l)
                ev = torch.cuda.Event()
                ev.record(side)
            events.append(ev)
        return events

    def fold_pack_block(self, l):
        """Fold (ERB) + pack block l's kernel into the bf16 operand layouts on the current stream."""
        g, st = self.geoms[l], _lib.stream()
        K, b, tap_major = self._block_kernel(l)
        pack = self.lib.onr_pack_weights_t if tap_major else self.lib.onr_pack_weights
        check(pack(ptr(K), ptr(b), g.cin, g.cnew, g.s, g.npad, g.cpi_rows,
                   ptr(self.wf[l]), ptr(self.wd[l]) if self.train else None, ptr(self.bias_p[l]), st),
              "onr_pack_weights")

    # ------------------------------------------------------------------------------------- forward
    def forward(self, embed=None, t_norm=None, freqs=None, refresh=True, out=None, outs=None):
        """Runs the decoder; returns the image [B,3,H,W] fp32: `out` when given (it then also becomes the image the
        next `backward` differentiates through), else the executor's own static buffer.
        Either `embed` [B,2L] (reference entry) or `t_norm` [B] + `freqs` [L] (fused PE) is given.
        With multi-resolution heads the images of the earlier stages land in `outs[l]` when given, else in static
        buffers; `self.imgs` holds them either way."""
        self.img = out if out is not None else self._img_static
        self.imgs = {l: (outs[l] if outs is not None else self._img_stage[l]) for l in self._img_stage}
        gen, st = self.gen, _lib.stream()
        lin1, lin2 = gen.stem[0], gen.stem[2]
        g0 = self.geoms[0]
        if refresh and not self.train:
            # decode: re-fold / re-pack only when a parameter changed since the last call
            key = self._weights_key()
            refresh = key != getattr(self, "_packed_key", None)
            self._packed_key = key
        if refresh or self.train:
            # torch.nn.utils.prune recomputes `.weight = weight_orig * weight_mask` in a forward-pre-hook; the stem
            # Linear modules are never called here, so run their hooks by hand (the reference re-evaluates them on
            # every forward, main_eval.py:572-587 + model.py:612)
            for lin in (lin1, lin2):
                for hook in lin._forward_pre_hooks.values():
                    hook(lin, None)
        events = self.refresh_weights() if refresh else None     # forks side streams first
        if t_norm is None:
            if embed is None:
                raise ValueError("forward needs embed or t_norm")
            if embed.data_ptr() != self.embed.data_ptr():
                self.embed.copy_(embed.reshape(self.B, self.E))
        elif freqs is None:
            raise ValueError("t_norm needs the frequency table of the PositionalEncoding")
        check(self.lib.onr_pe_stem_fwd_act(
            ptr(t_norm), self.B, ptr(freqs), self.E // 2,
            ptr(lin1.weight), ptr(lin1.bias), self.hid, ptr(lin2.weight), ptr(lin2.bias),
            gen.fc_dim, gen.fc_h, gen.fc_w, g0.cpi,
            ptr(self.embed), ptr(self.pre1), ptr(self.h1), ptr(self.x[0]), ptr(self.d[0]), self.act, st),
            "onr_pe_stem_fwd")
        main = torch.cuda.current_stream()
        head = gen.head_conv()
        if self._decode_fused:
            check(self.lib.onr_conv_plan_set_head(self.fprop[-1].handle, ptr(self.img), ptr(head.weight), ptr(head.bias)),
                  "onr_conv_plan_set_head")
        for l in range(self.L):
            if events is not None:
                main.wait_event(events[l])
            check(self.lib.onr_conv_plan_run(self.fprop[l].handle, st), "onr_conv_plan_run(fprop)")
            if self.act != 0:
                g = self.geoms[l]
                check(self.lib.onr_act_map(ptr(self.x[l + 1]), ptr(self.d[l + 1]) if self.train else None,
                                           self.B * g.ho * g.wo, g.cnew, g.cpo, self.act, st), "onr_act_map")
            if l in self.imgs:                                   # multi-resolution head of this stage
                g, hl = self.geoms[l], gen.head_layers[l]
                check(self.lib.onr_head_fwd(ptr(self.x[l + 1]), self.B, g.ho, g.wo, g.cnew, g.cpo, ptr(hl.weight),
                                            ptr(hl.bias), 1 if gen.sigmoid else 0, ptr(self.imgs[l]), st),
                      "onr_head_fwd(stage)")
        if self._decode_fused:
            return self.img
        head_fwd = self.lib.onr_head_fwd_z if self._last_z else self.lib.onr_head_fwd
        check(head_fwd(
            ptr(self.x[self.L]), self.B, self.H, self.W, self.C_last, self.geoms[-1].cpo,
            ptr(head.weight), ptr(head.bias), 1 if gen.sigmoid else 0, ptr(self.img), st), "onr_head_fwd")
        return self.img

    # ------------------------------------------------------------------------------------- decode
    def decode(self, embed):
        """Decode (reference main_eval.py:753-762 `model(embed_input)` under no_grad) as ONE CUDA-graph replay: stem,
        the block convolutions and the head are captured once per set of packed weights — a 720p decode is eight
        kernels of 5-160 us, so the launch gaps of eager issue are a visible share of the frame.  Weights are re-folded /
        re-packed eagerly (outside the graph) whenever a parameter changed, which also drops the captured graph.
        Returns fresh image tensors, like the reference: [image of every head stage ..., final image].
        ONR_DECODE_GRAPH=0 keeps the eager launches."""
        assert not self.train

        def snapshot():
            return [self._img_stage[l].clone() for l in self.head_stages[:-1]] + [self._img_static.clone()]

        # (multi-process runs keep the eager launches: a process group's watchdog thread touches the CUDA API, which a
        # capture in the default global error mode does not tolerate)
        multi_proc = torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1
        if os.environ.get("ONR_DECODE_GRAPH", "1") == "0" or multi_proc or torch.cuda.is_current_stream_capturing():
            img = torch.empty(self.B, 3, self.H, self.W, dtype=torch.float32, device=self.dev)
            outs = {l: torch.empty_like(self._img_stage[l]) for l in self.head_stages[:-1]}
            self.forward(embed=embed, out=img, outs=outs)
            return [outs[l] for l in self.head_stages[:-1]] + [img]
        stale = self._weights_key() != getattr(self, "_packed_key", None)
        if stale or getattr(self, "_decode_graph", None) is None:
            self._decode_graph = None
            self.forward(embed=embed)                     # eager: refreshes the operands; the warm-up of the capture
            out = snapshot()
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.forward(embed=self.embed, refresh=False)
            self._decode_graph = g
            return out
        self.embed.copy_(embed.reshape(self.B, self.E))
        self._decode_graph.replay()
        return snapshot()

    # ------------------------------------------------------------------------------------ backward
    def backward(self, gimg, grads, block_hook=None, reduce=None, gimgs=None):
        """Backward of `forward`. `grads` maps parameter name -> fp32 gradient tensor; every tensor must be
        zero on entry (the kernels accumulate into the ERB branch / stem / head gradients and overwrite the
        single-branch conv gradients).

        reduce(t), if given (data-parallel fitting), sums the fp32 tensor `t` over the ranks in place on the current
        stream.  It is called once per exchange bucket as soon as that bucket is complete — head gradients after the
        head backward, each block's folded-kernel gradient `dK|dbias` right after its wgrad on the block's side stream
        (the linear fold backward then runs on the summed dK on every rank), stem gradients at the end — so the
        exchange overlaps the remaining dgrad chain.

        gimgs, with multi-resolution heads: {stage l: dL/d(image of stage l)} for the earlier head stages; each head's
        backward adds its contribution to dz[l+1] before block l's wgrad / dgrad consume it.

        block_hook(l), if given, is called on block l's side stream once that block's parameter gradients are
        complete AND its dgrad has been issued (nothing in this backward reads the block's weights, packed operands
        or T afterwards): a trainer can update the block and re-fold it for the next step right there."""
        assert self.train
        gen, st, lib = self.gen, _lib.stream(), self.lib
        head = gen.head_conv()
        hname = gen.head_name()
        gL = self.geoms[-1]
        main = torch.cuda.current_stream()
        joins = []
        fork = torch.cuda.Event()
        fork.record(main)
        hside = self._side_streams()[0]
        hside.wait_event(fork)
        with torch.cuda.stream(hside):
            self._wgrad_pool.zero_()
            pool_clear = torch.cuda.Event()
            pool_clear.record(hside)
        if self._last_z:
            # the same single pass, reading only the pre-activation z of the last block
            check(lib.onr_head_bwd_z(
                ptr(gimg), ptr(self.img), ptr(self.x[self.L]), self.B, self.H, self.W, self.C_last, gL.cpo,
                ptr(head.weight), 1 if gen.sigmoid else 0, ptr(grads[hname + ".weight"]), ptr(grads[hname + ".bias"]),
                ptr(self.dz[self.L]), st), "onr_head_bwd_z")
            if reduce is not None:
                head_done = torch.cuda.Event()
                head_done.record(main)
                with torch.cuda.stream(hside):
                    hside.wait_event(head_done)
                    reduce(self._flat_span(grads, [hname + ".weight", hname + ".bias"]))
                    ev = torch.cuda.Event()
                    ev.record(hside)
                    joins.append(ev)
        elif self._head_fused:
            # one pass over the pixels: dz = (Wh^T g_pre) * SiLU' and the head weight/bias gradient reduction
            check(lib.onr_head_bwd(
                ptr(gimg), ptr(self.img), ptr(self.x[self.L]), ptr(self.d[self.L]), self.B, self.H, self.W,
                self.C_last, gL.cpo, ptr(head.weight), 1 if gen.sigmoid else 0, ptr(grads[hname + ".weight"]),
                ptr(grads[hname + ".bias"]), ptr(self.dz[self.L]), st), "onr_head_bwd")
            if reduce is not None:
                head_done = torch.cuda.Event()
                head_done.record(main)
                with torch.cuda.stream(hside):
                    hside.wait_event(head_done)
                    reduce(self._flat_span(grads, [hname + ".weight", hname + ".bias"]))
                    ev = torch.cuda.Event()
                    ev.record(hside)
                    joins.append(ev)
        else:
            # split: the reduction on a side stream, only dz on the critical path
            with torch.cuda.stream(hside):
                check(lib.onr_head_bwd_gw(
                    ptr(gimg), ptr(self.img), ptr(self.x[self.L]), self.B, self.H, self.W, self.C_last, gL.cpo,
                    1 if gen.sigmoid else 0, ptr(grads[hname + ".weight"]), ptr(grads[hname + ".bias"]),
                    _lib.stream()), "onr_head_bwd_gw")
                if reduce is not None:
                    reduce(self._flat_span(grads, [hname + ".weight", hname + ".bias"]))
                ev = torch.cuda.Event()
                ev.record(hside)
                joins.append(ev)
            check(lib.onr_head_bwd_dz(
                ptr(gimg), ptr(self.img), ptr(self.d[self.L]), self.B, self.H, self.W, self.C_last, gL.cpo,
                ptr(head.weight), 1 if gen.sigmoid else 0, ptr(self.dz[self.L]), st), "onr_head_bwd_dz")
        if not self._wgrad_on_side:
            main.wait_event(pool_clear)
        for l in reversed(range(self.L)):
            g, blk = self.geoms[l], gen.layers[l]
            if gimgs is not None and l in self._dz_head and gimgs.get(l) is not None:
                # head of an earlier stage (reference model.py:619-623 under autograd): its gradient w.r.t. the block
                # output joins the one the next block's dgrad has just written into dz[l+1]
                hl, nm = gen.head_layers[l], f"head_layers.{l}"
                check(lib.onr_head_bwd(
                    ptr(gimgs[l]), ptr(self.imgs[l]), ptr(self.x[l + 1]), ptr(self.d[l + 1]), self.B, g.ho, g.wo, g.cnew,
                    g.cpo, ptr(hl.weight), 1 if gen.sigmoid else 0, ptr(grads[nm + ".weight"]), ptr(grads[nm + ".bias"]),
                    ptr(self._dz_head[l]), st), "onr_head_bwd(stage)")
                check(lib.onr_add_bf16(ptr(self.dz[l + 1]), ptr(self._dz_head[l]), self.dz[l + 1].numel(), st),
                      "onr_add_bf16")
                if reduce is not None:
                    reduce(self._flat_span(grads, [nm + ".weight", nm + ".bias"]))
            # dz[l+1] is ready: wgrad -> un-pack -> fold backward of block l run on a side stream beside the
            # dgrad chain (only the dgrads and the stem backward are on the critical path)
            if not self._wgrad_on_side:
                check(lib.onr_wgrad_plan_run(self.wgrad[l].handle, st), "onr_wgrad_plan_run")
            done = torch.cuda.Event()
            done.record(main)
            side = self._side_streams()[l]
            side.wait_event(done)
            with torch.cuda.stream(side):
                sst = _lib.stream()
                if self._wgrad_on_side:
                    side.wait_event(pool_clear)
                    check(lib.onr_wgrad_plan_run(self.wgrad[l].handle, sst), "onr_wgrad_plan_run")
                folded = blk.fold_kind() is not None
                if folded:
                    dK, db = self.dK[l], self.dbias[l]
                else:   # single-branch block: dK is the parameter gradient itself (grads are zero on entry)
                    name = f"layers.{l}." + blk.single_conv_name()
                    dK, db = grads[name + ".weight"], grads[name + ".bias"]
                unpack = lib.onr_unpack_wgrad_t if (blk.is_erb_train() and self._fold_tc) else lib.onr_unpack_wgrad
                check(unpack(ptr(self.dKp[l]), ptr(self.dbias_p[l]), g.cin, g.cnew, g.s, ptr(dK), ptr(db), sst),
                      "onr_unpack_wgrad")
                if reduce is not None:
                    reduce(self.dKb[l] if folded else
                           self._flat_span(grads, [name + ".weight", name + ".bias"]))
                if folded:
                    self.scatter_block_grads(l, grads)
                ev = torch.cuda.Event()
                ev.record(side)
                joins.append(ev)
            check(lib.onr_conv_plan_run(self.dgrad[l].handle, st), "onr_conv_plan_run(dgrad)")
            if block_hook is not None:
                issued = torch.cuda.Event()
                issued.record(main)
                side.wait_event(issued)
                with torch.cuda.stream(side):
                    block_hook(l)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    joins.append(ev)
        lin1 = gen.stem[0]
        lin2 = gen.stem[2]
        g0 = self.geoms[0]
        gather = getattr(reduce, "all_gather_slots", None) if reduce is not None else None
        if gather is not None and self.B == 1:
            # data-parallel, one frame per rank: exchange the rank-1 FACTORS of the stem gradients (an all-gather of
            # ~21 KB per rank) instead of all-reducing the 7.7 .. 33 MB matrices at the very end of the backward
            slots, rank, world = reduce.stem_slots(self)
            check(lib.onr_stem_bwd_factors_act(
                ptr(self.dz[0]), ptr(self.embed), self.E, ptr(self.pre1), ptr(self.h1), self.hid, ptr(lin2.weight),
                gen.fc_dim, gen.fc_h, gen.fc_w, g0.cpi, ptr(slots[rank]), ptr(self.dh1), self.act, st),
                "onr_stem_bwd_factors")
            gather(slots, rank)
            check(lib.onr_stem_grads_from_factors(
                ptr(slots), world, self.E, self.hid, gen.fc_dim, gen.fc_h, gen.fc_w,
                ptr(grads["stem.0.weight"]), ptr(grads["stem.0.bias"]), ptr(grads["stem.2.weight"]),
                ptr(grads["stem.2.bias"]), st), "onr_stem_grads_from_factors")
        else:
            check(lib.onr_stem_bwd_act(
                ptr(self.dz[0]), self.B, ptr(self.embed), self.E, ptr(self.pre1), ptr(self.h1), self.hid,
                ptr(lin2.weight), gen.fc_dim, gen.fc_h, gen.fc_w, g0.cpi,
                ptr(grads["stem.0.weight"]), ptr(grads["stem.0.bias"]),
                ptr(grads["stem.2.weight"]), ptr(grads["stem.2.bias"]), ptr(self.dh1), self.act, st), "onr_stem_bwd")
            if reduce is not None:
                reduce(self._flat_span(grads, ["stem.0.weight", "stem.0.bias", "stem.2.weight", "stem.2.bias"]))
        for ev in joins:
            main.wait_event(ev)

    @staticmethod
    def _flat_span(grads, names):
        """The slice of the flat gradient buffer that covers the (consecutive) parameters `names`."""
        flat, offs = grads["__flat__"], grads["__offsets__"]
        lo = min(offs[n][0] for n in names)
        hi = max(offs[n][0] + offs[n][1] for n in names)
        return flat[lo:hi]

    def scatter_block_grads(self, l, grads):
        """dK/dbias of a multi-branch block l -> gradients of its branch tensors (fold backward)."""
        g, blk, st = self.geoms[l], self.gen.layers[l], _lib.stream()
        pfx = f"layers.{l}."
        if blk.fold_kind() == "set":
            branches.fold_bwd(self.lib, blk, self.dK[l], self.dbias[l], lambda n: grads.get(pfx + n), st)
        elif self._fold_tc:
            names = ("rbr_3x3_branch.weight", "rbr_3x3_branch.bias", "rbr_1x3_branch.weight", "rbr_1x3_branch.bias",
                     "rbr_3x1_branch.weight", "rbr_3x1_branch.bias", "rbr_1x1_3x3_1x1_branch_1x1_1.weight",
                     "rbr_1x1_3x3_1x1_branch_3x3.weight", "rbr_1x1_3x3_1x1_branch_1x1_2.weight")
            self.fold[l].bwd(self.dK[l], self.dbias[l], [grads[pfx + n] for n in names], st)
        else:
            b = blk
            check(self.lib.onr_erb_fold_bwd(
                ptr(self.dK[l]), ptr(self.dbias[l]),
                ptr(b.rbr_1x1_3x3_1x1_branch_1x1_1.weight), ptr(b.rbr_1x1_3x3_1x1_branch_3x3.weight),
                ptr(b.rbr_1x1_3x3_1x1_branch_1x1_2.weight), ptr(self.T[l]), g.cin, g.cout,
                ptr(grads[pfx + "rbr_3x3_branch.weight"]), ptr(grads[pfx + "rbr_3x3_branch.bias"]),
                ptr(grads[pfx + "rbr_1x3_branch.weight"]), ptr(grads[pfx + "rbr_1x3_branch.bias"]),
                ptr(grads[pfx + "rbr_3x1_branch.weight"]), ptr(grads[pfx + "rbr_3x1_branch.bias"]),
                ptr(grads[pfx + "rbr_1x1_3x3_1x1_branch_1x1_1.weight"]),
                ptr(grads[pfx + "rbr_1x1_3x3_1x1_branch_3x3.weight"]),
                ptr(grads[pfx + "rbr_1x1_3x3_1x1_branch_1x1_2.weight"]),
                ptr(self.dT[l]), st), "onr_erb_fold_bwd")
