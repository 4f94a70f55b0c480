"""The reference's other linear branch sets — ACB, RepVGG, DBB, ECB (model.py:345-393, SeqConv3x3 :191-300) — as
parameter containers plus the slot table that maps their tensors onto `onr_branch_set` (include/orepnerv.h).

The reference trains these blocks with an explicit multi-branch forward (model.py:541-565: 2-8 cuDNN convolutions, an
average pool and the additions per block) and cannot deploy them (`switch_to_deploy` hard-codes the ERB names).  Here
they take the same route as ERB: the branches are folded into ONE 3x3 kernel on the device every step
(csrc/fold_branches.cu), the block runs the single tcgen05 convolution, and the fold's backward scatters dK / dbias to
the branch parameters — same function, same gradients, one convolution instead of up to eight, and
`switch_to_deploy()` works for all of them.
"""
import ctypes as C

import torch
import torch.nn as nn

from ._lib import BranchSet, check

FOLDED_BRANCH_TYPES = ("ACB", "RepVGG", "DBB", "ECB")
EDGE_BRANCHES = ("rbr_conv1x1_sbx_branch", "rbr_conv1x1_sby_branch", "rbr_conv1x1_lpl_branch")
# every sub-module name a training block of these types may own (deleted by switch_to_deploy, model.py:434-443)
SET_BRANCH_MODULES = ("rbr_3x3_branch", "rbr_3x1_branch", "rbr_1x3_branch", "rbr_1x1_branch", "rbr_1x1_3x3_branch_1x1",
                      "rbr_1x1_3x3_branch_3x3", "rbr_1x1_avg_branch_1x1", "rbr_1x1_avg_branch_avg") + EDGE_BRANCHES


class SeqConv3x3(nn.Module):
    """reference model.py:191-300: 1x1 convolution (k0, b0) followed by a fixed depthwise 3x3 edge filter
    (`scale * mask`, + bias).  Parameter container only — creation order and RNG consumption as in the reference
    (Conv2d init, then randn scale, randn bias), `mask` registered as a Parameter with requires_grad=False."""

    _MASKS = {
        'conv1x1-sobelx': {(0, 0): 1.0, (1, 0): 2.0, (2, 0): 1.0, (0, 2): -1.0, (1, 2): -2.0, (2, 2): -1.0},
        'conv1x1-sobely': {(0, 0): 1.0, (0, 1): 2.0, (0, 2): 1.0, (2, 0): -1.0, (2, 1): -2.0, (2, 2): -1.0},
        'conv1x1-laplacian': {(0, 1): 1.0, (1, 0): 1.0, (1, 2): 1.0, (2, 1): 1.0, (1, 1): -4.0},
    }

    def __init__(self, seq_type, inp_planes, out_planes):
        super().__init__()
        if seq_type not in self._MASKS:
            raise ValueError('the type of seqconv is not supported!')
        self.type, self.inp_planes, self.out_planes = seq_type, inp_planes, out_planes
        conv0 = nn.Conv2d(inp_planes, out_planes, kernel_size=1, padding=0)
        self.k0 = conv0.weight
        self.b0 = conv0.bias
        self.scale = nn.Parameter(torch.randn(size=(out_planes, 1, 1, 1)) * 1e-3)
        self.bias = nn.Parameter(torch.reshape(torch.randn(out_planes) * 1e-3, (out_planes,)))
        mask = torch.zeros((out_planes, 1, 3, 3), dtype=torch.float32)
        for (h, w), v in self._MASKS[seq_type].items():
            mask[:, 0, h, w] = v
        self.mask = nn.Parameter(data=mask, requires_grad=False)


def create_branches(blk, branch_type, ci, co):
    """The sub-modules of a training block, in the reference's creation order (model.py:345-393)."""
    blk.rbr_3x3_branch = nn.Conv2d(ci, co, (3, 3), 1, 1)
    if branch_type == "ACB":
        blk.rbr_3x1_branch = nn.Conv2d(ci, co, (3, 1), 1, (1, 0))
        blk.rbr_1x3_branch = nn.Conv2d(ci, co, (1, 3), 1, (0, 1))
    elif branch_type == "RepVGG":
        blk.rbr_1x1_branch = nn.Conv2d(ci, co, (1, 1), 1, 0)
    elif branch_type == "DBB":
        blk.rbr_1x1_branch = nn.Conv2d(ci, co, (1, 1), 1, 0)
        blk.rbr_1x1_3x3_branch_1x1 = nn.Conv2d(ci, 2 * ci, (1, 1), 1, 0, bias=False)
        blk.rbr_1x1_3x3_branch_3x3 = nn.Conv2d(2 * ci, co, (3, 3), 1, 1, bias=False)
        blk.rbr_1x1_avg_branch_1x1 = nn.Conv2d(ci, co, (1, 1), 1, 0, bias=False)
        blk.rbr_1x1_avg_branch_avg = nn.AvgPool2d(kernel_size=3, stride=1, padding=1)
    elif branch_type == "ECB":
        blk.rbr_1x1_3x3_branch_1x1 = nn.Conv2d(ci, 2 * ci, (1, 1), 1, 0, bias=False)
        blk.rbr_1x1_3x3_branch_3x3 = nn.Conv2d(2 * ci, co, (3, 3), 1, 1, bias=False)
        blk.rbr_conv1x1_sbx_branch = SeqConv3x3('conv1x1-sobelx', ci, co)
        blk.rbr_conv1x1_sby_branch = SeqConv3x3('conv1x1-sobely', ci, co)
        blk.rbr_conv1x1_lpl_branch = SeqConv3x3('conv1x1-laplacian', ci, co)
    else:
        raise KeyError(branch_type)


def branch_slots(blk):
    """[(slot, parameter name relative to the block, tensor)] for the branches the block owns.  slot is a field of
    onr_branch_set, or (field, e) for the per-SeqConv3x3 arrays."""
    out = []

    def conv(mod_name, wslot, bslot=None):
        m = getattr(blk, mod_name, None)
        if m is None:
            return
        out.append((wslot, mod_name + ".weight", m.weight))
        if bslot is not None and m.bias is not None:
            out.append((bslot, mod_name + ".bias", m.bias))

    conv("rbr_3x3_branch", "w3x3", "b3x3")
    conv("rbr_1x3_branch", "w1x3", "b1x3")
    conv("rbr_3x1_branch", "w3x1", "b3x1")
    conv("rbr_1x1_branch", "w1x1", "b1x1")
    conv("rbr_1x1_3x3_branch_1x1", "seq_w1")
    conv("rbr_1x1_3x3_branch_3x3", "seq_w2")
    conv("rbr_1x1_avg_branch_1x1", "avg_w")
    for e, name in enumerate(EDGE_BRANCHES):
        m = getattr(blk, name, None)
        if m is not None:
            for field, attr in (("edge_k0", "k0"), ("edge_b0", "b0"), ("edge_scale", "scale"), ("edge_bias", "bias"),
                                ("edge_mask", "mask")):
                out.append(((field, e), f"{name}.{attr}", getattr(m, attr)))
    return out


def make_branch_set(cin, cout, slot_tensors):
    """onr_branch_set from {slot: tensor or None} (fp32, contiguous; data pointers are taken as they are NOW)."""
    s = BranchSet()
    s.cin, s.cout = cin, cout
    for slot, t in slot_tensors.items():
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("branch tensors must be contiguous fp32")
        if isinstance(slot, tuple):
            getattr(s, slot[0])[slot[1]] = t.data_ptr()
        else:
            setattr(s, slot, t.data_ptr())
    return s


def fold_fwd(lib, blk, K, bias, st):
    """K[Cout,Cin,3,3], bias[Cout] <- the single-convolution equivalent of the block's branch set."""
    s = make_branch_set(blk.ngf, blk.out_channels, {slot: t for slot, _, t in branch_slots(blk)})
    check(lib.onr_branch_fold_fwd(C.byref(s), K.data_ptr(), bias.data_ptr(), st), "onr_branch_fold_fwd")


def fold_bwd(lib, blk, dK, dbias, grad_of, st):
    """Scatters dK / dbias to the branch gradients; grad_of(name) -> the gradient tensor of the block-relative
    parameter `name` (or None to skip it).  SeqConv3x3.mask is a constant and is skipped."""
    slots = branch_slots(blk)
    s = make_branch_set(blk.ngf, blk.out_channels, {slot: t for slot, _, t in slots})
    g = make_branch_set(blk.ngf, blk.out_channels,
                        {slot: (None if name.endswith(".mask") else grad_of(name)) for slot, name, _ in slots})
    check(lib.onr_branch_fold_bwd(C.byref(s), dK.data_ptr(), dbias.data_ptr(), C.byref(g), st), "onr_branch_fold_bwd")
