"""ctypes binding of liborepnerv.so (the C ABI declared in include/orepnerv.h).

There is exactly one compute path: the sm_100a kernels in that library.  If the library is missing,
cannot be loaded, or the current device is not a B200-class GPU, the callers raise — nothing here (or
anywhere in the package) falls back to PyTorch ops or to the CPU oracle.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liborepnerv.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "orepnerv.h")

_lib = None
_device_checked = False

vp = C.c_void_p
i32 = C.c_int
f32 = C.c_float
sz = C.c_size_t
u32 = C.c_uint32


class ConvDesc(C.Structure):
    """onr_conv_desc"""
    _fields_ = [
        ("kind", i32), ("B", i32), ("H", i32), ("W", i32),
        ("a", vp), ("a_cp", i32), ("a_s", i32),
        ("w", vp), ("n_rows", i32),
        ("n_total", i32),
        ("out", vp), ("out_cp", i32), ("out_s", i32),
        ("out_d", vp),
        ("bias_p", vp),
        ("dmul", vp),
        ("head_w", vp), ("head_b", vp), ("head_c", i32), ("use_sigmoid", i32), ("img", vp),
    ]


class WgradDesc(C.Structure):
    """onr_wgrad_desc"""
    _fields_ = [
        ("B", i32), ("H", i32), ("W", i32),
        ("x", vp), ("x_cp", i32),
        ("dz", vp), ("dz_cp", i32), ("s", i32),
        ("dKp", vp),
        ("dbias_p", vp),
    ]


class BranchSet(C.Structure):
    """onr_branch_set"""
    _fields_ = [("cin", i32), ("cout", i32)] + \
        [(n, vp) for n in ("w3x3", "b3x3", "w1x3", "b1x3", "w3x1", "b3x1", "w1x1", "b1x1", "seq_w1", "seq_w2", "avg_w")] + \
        [(n, vp * 3) for n in ("edge_k0", "edge_b0", "edge_scale", "edge_bias", "edge_mask")]


CONV_FPROP_TRAIN, CONV_FPROP_INFER, CONV_DGRAD, CONV_FPROP_Z, CONV_FPROP_HEAD = 0, 1, 2, 3, 4

_SIGNATURES = {
    "onr_abi_version": (i32, []),
    "onr_last_error": (C.c_char_p, []),
    "onr_check_device": (i32, []),
    "onr_launch_count": (C.c_ulonglong, []),
    "onr_pe_stem_fwd": (i32, [vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "onr_pe_stem_fwd_act": (i32, [vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]),
    "onr_stem_bwd_act": (i32, [vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]),
    "onr_stem_bwd_factors_act": (i32, [vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp]),
    "onr_act_map": (i32, [vp, vp, sz, i32, i32, i32, vp]),
    "onr_add_bf16": (i32, [vp, vp, sz, vp]),
    "onr_adaptive_avg_pool": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "onr_pos_encoding": (i32, [vp, i32, vp, i32, vp, vp]),
    "onr_frame_u8_to_f32": (i32, [vp, sz, vp, vp]),
    "onr_stem_bwd": (i32, [vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "onr_stem_factor_floats": (sz, [i32, i32, i32, i32, i32]),
    "onr_stem_bwd_factors": (i32, [vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp]),
    "onr_stem_grads_from_factors": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "onr_erb_fold_fwd": (i32, [vp] * 9 + [i32, i32, vp, vp, vp, vp]),
    "onr_erb_fold_bwd": (i32, [vp] * 6 + [i32, i32] + [vp] * 10 + [vp]),
    "onr_fold_workspace_bytes": (sz, [i32, i32, i32]),
    "onr_fold_plan_create": (i32, [C.POINTER(vp), i32, i32, vp, i32]),
    "onr_fold_plan_destroy": (None, [vp]),
    "onr_fold_plan_fwd": (i32, [vp] + [vp] * 9 + [vp, vp, vp]),
    "onr_fold_plan_bwd": (i32, [vp, vp, vp] + [vp] * 9 + [vp]),
    "onr_branch_fold_fwd": (i32, [C.POINTER(BranchSet), vp, vp, vp]),
    "onr_branch_fold_bwd": (i32, [C.POINTER(BranchSet), vp, vp, C.POINTER(BranchSet), vp]),
    "onr_tapmajor_permute": (i32, [vp, vp, i32, i32, i32, vp]),
    "onr_pack_weights_t": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "onr_unpack_wgrad_t": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "onr_pack_weights": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "onr_unpack_wgrad": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "onr_conv_plan_create": (i32, [C.POINTER(vp), C.POINTER(ConvDesc)]),
    "onr_conv_plan_run": (i32, [vp, vp]),
    "onr_conv_plan_set_head": (i32, [vp, vp, vp, vp]),
    "onr_conv_plan_destroy": (None, [vp]),
    "onr_conv_plan_info": (i32, [vp] + [C.POINTER(i32)] * 5),
    "onr_conv_plan_set_prof": (i32, [vp, vp, C.POINTER(i32)]),
    "onr_conv_tile_n": (i32, [i32, C.POINTER(i32), C.POINTER(i32)]),
    "onr_wgrad_plan_create": (i32, [C.POINTER(vp), C.POINTER(WgradDesc)]),
    "onr_wgrad_plan_run": (i32, [vp, vp]),
    "onr_wgrad_plan_destroy": (None, [vp]),
    "onr_nchw_to_nhwc_bf16": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "onr_nhwc_bf16_to_nchw": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "onr_head_fwd": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "onr_head_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp]),
    "onr_head_fwd_z": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "onr_head_bwd_z": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp]),
    "onr_head_bwd_dz": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp]),
    "onr_head_bwd_gw": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "onr_loss_workspace_bytes": (sz, [i32, i32, i32]),
    "onr_fusion6_fwd_bwd": (i32, [vp, vp, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp]),
    "onr_fusion_loss": (i32, [vp, vp, i32, i32, i32, f32, f32, f32, f32, vp, vp, vp, vp]),
    "onr_scale_by_device_scalar": (i32, [vp, sz, vp, vp]),
    "onr_msssim_workspace_bytes": (sz, [i32, i32, i32]),
    "onr_msssim": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "onr_adam_block_elems": (sz, []),
    "onr_adam_max_tensors": (i32, []),
    "onr_adam_multi": (i32, [vp, i32, sz, vp, vp, vp, f32, f32, f32, f32, i32, vp]),
    "onr_sched_tick": (i32, [vp, vp, C.c_double, i32, i32, i32, i32, i32, vp]),
    "onr_sched_tick_ex": (i32, [vp, vp, C.c_double, i32, i32, i32, i32, i32, i32, i32, vp]),
    "onr_mul_inplace_f32": (i32, [vp, vp, sz, vp]),
    "onr_abs_radix_hist": (i32, [vp, sz, u32, u32, i32, vp, vp]),
    "onr_apply_magnitude_mask": (i32, [vp, sz, f32, vp, vp, vp]),
    "onr_quant_rows": (i32, [vp, i32, sz, i32, vp, vp, vp, vp]),
}


def declared_symbols():
    """Every function name include/orepnerv.h declares (used by the CPU test that checks exports)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(onr_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load the shared library (no GPU needed) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.onr_abi_version() != 1:
        raise RuntimeError("liborepnerv.so ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().onr_last_error().decode(errors="replace")


_SYNC_DEBUG = bool(os.environ.get("ONR_SYNC_DEBUG"))


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"liborepnerv {what} failed (rc={rc}): {last_error()}")
    if _SYNC_DEBUG:
        # debugging aid: serialise after every library call and leave a trail (the last line printed before a hang
        # names the kernel that never finished)
        import sys
        import torch
        print(f"[onr] {what} ...", file=sys.stderr, end="", flush=True)
        torch.cuda.synchronize()
        print(" ok", file=sys.stderr, flush=True)


def lib():
    """Library handle for compute calls: also verifies once that the current device is sm_100."""
    global _device_checked
    l = load()
    if not _device_checked:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("orepnerv needs a CUDA device (B200, sm_100a); there is no CPU path")
        check(l.onr_check_device(), "onr_check_device")
        _device_checked = True
    return l


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
