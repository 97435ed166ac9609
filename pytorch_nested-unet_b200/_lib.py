"""ctypes binding of libnunet_b200.so (the C ABI declared in include/nunet_b200.h).

The product path has NO fallback: if the library is missing or a call fails, an exception is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnunet_b200.so")

MAX_SRC = 5
ABI_VERSION = 2   # NUNET_ABI_VERSION of include/nunet_b200.h these structures mirror
SRC_INPUT, SRC_SAME, SRC_POOL, SRC_UP, SRC_GRAD, SRC_ACT, SRC_POOL_ACT, SRC_UP_ACT = 0, 1, 2, 3, 4, 5, 6, 7
ACT_KIND = {SRC_SAME: SRC_ACT, SRC_POOL: SRC_POOL_ACT, SRC_UP: SRC_UP_ACT}
PACK_FWD, PACK_DGRAD, PACK_FWD_C32, PACK_DGRAD_C32 = 0, 1, 2, 3


class Src(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data2", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("k2", C.c_void_p), ("C", C.c_int32), ("kind", C.c_int32)]


class BnExchange(C.Structure):  # nunet_bn_exchange_t
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("data", C.c_void_p * 8), ("flags", C.c_void_p * 8), ("seq", C.c_void_p)]


class BnTail(C.Structure):      # nunet_bn_tail_t
    _fields_ = [("counter", C.c_void_p), ("count", C.c_double), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("conv_bias", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("momentum", C.c_float), ("eps", C.c_float),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("save_mean", C.c_void_p), ("save_invstd", C.c_void_p), ("sums_out", C.c_void_p),
                ("xch", BnExchange)]


class BnBwdTail(C.Structure):   # nunet_bn_bwd_tail_t
    _fields_ = [("counter", C.c_void_p), ("count", C.c_double), ("gamma", C.c_void_p), ("save_mean", C.c_void_p), ("save_invstd", C.c_void_p),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("k1", C.c_void_p), ("k2", C.c_void_p), ("k3", C.c_void_p),
                ("grad_scale", C.c_float), ("sums_out", C.c_void_p), ("xch", BnExchange)]


class ConvDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("num_src", C.c_int32), ("src", Src * MAX_SRC),
                ("Cout", C.c_int32), ("wpack", C.c_void_p), ("out", C.c_void_p), ("stat_part", C.c_void_p), ("wpack_layout", C.c_int32),
                ("bn", BnTail), ("num_out_seg", C.c_int32), ("out_seg", C.c_void_p * MAX_SRC), ("out_seg_c", C.c_int32 * MAX_SRC),
                ("mask_y", C.c_void_p), ("mask_scale", C.c_void_p), ("mask_shift", C.c_void_p), ("mask_mean", C.c_void_p),
                ("mask_invstd", C.c_void_p), ("bn_bwd", BnBwdTail)]


class ConvPlan(C.Structure):
    _fields_ = [("T", C.c_int32), ("NT", C.c_int32), ("CK", C.c_int32), ("S", C.c_int32), ("P", C.c_int32),
                ("nchunks", C.c_int32), ("n_tiles", C.c_int32), ("grid_m", C.c_int32), ("cin_pad", C.c_int32),
                ("smem_bytes", C.c_int32), ("tmem_cols", C.c_int32), ("ntg", C.c_int32), ("tma_a_stage_bytes", C.c_int32),
                ("tma_smem_bytes", C.c_int32), ("tma_epi_warps", C.c_int32), ("strips", C.c_int32), ("wpack_elems", C.c_int64)]


class GradCons(C.Structure):
    _fields_ = [("dx", C.c_void_p), ("pitch", C.c_int32), ("coff", C.c_int32), ("kind", C.c_int32)]


class RouterDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("y", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("num_cons", C.c_int32), ("cons", GradCons * 6), ("dlogits", C.c_void_p), ("wh", C.c_void_p),
                ("K", C.c_int32), ("gz", C.c_void_p), ("part", C.c_void_p), ("bn", BnBwdTail),
                ("head_part", C.c_void_p), ("dwh", C.c_void_p), ("dbh", C.c_void_p)]


class WgradPlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("KP", "S", "Pg", "cpc", "nchunks_total", "splits", "m_tiles", "n_tiles", "cin_pad",
                                         "smem_bytes", "triple", "tma", "tma_PW", "tma_r", "tma_S", "tma_G", "tma_Nt", "tma_splits",
                                         "tma_cpc", "tma_chunks", "tma_smem_bytes")] + [("part_elems", C.c_int64)]


class WgradDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("num_src", C.c_int32), ("src", Src * MAX_SRC),
                ("Cout", C.c_int32), ("gy", C.c_void_p), ("part", C.c_void_p), ("dw", C.c_void_p), ("cin_real", C.c_int32)]


class PackJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p)] + [(n, C.c_int32) for n in ("Cout", "Cin", "mode", "NT", "CK", "nchunks", "n_tiles", "pad_")]


class ActivateDesc(C.Structure):   # nunet_activate_desc_t
    _fields_ = [("y", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("scale", C.c_void_p),
                ("shift", C.c_void_p), ("a", C.c_void_p), ("a_pool", C.c_void_p), ("a_up", C.c_void_p), ("wh", C.c_void_p), ("bh", C.c_void_p),
                ("K", C.c_int32), ("pad_", C.c_int32), ("logits", C.c_void_p)]


class NunetError(RuntimeError):
    pass


_lib = None

_SIGS = {
    "nunet_version": ([], C.c_int),
    "nunet_conv3x3_plan": ([C.c_int32] * 5 + [C.POINTER(ConvPlan)], C.c_int),
    "nunet_conv3x3_fwd": ([C.POINTER(ConvDesc), C.c_void_p], C.c_int),
    "nunet_conv3x3_c32_eligible": ([C.c_int32] * 5, C.c_int),
    "nunet_pack_weights": ([C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(ConvPlan), C.c_void_p, C.c_void_p], C.c_int),
    "nunet_pack_weights_batched": ([C.c_void_p, C.c_int32, C.c_void_p], C.c_int),
    "nunet_bn_workspace_bytes": ([C.c_int32], C.c_int),
    "nunet_bn_finalize_train": ([C.c_void_p, C.c_int32, C.c_int32, C.c_double] + [C.c_void_p] * 5 + [C.c_float, C.c_float]
                                + [C.c_void_p] * 8, C.c_int),
    "nunet_bn_finalize_from_sums": ([C.c_void_p, C.c_int32, C.c_double] + [C.c_void_p] * 5 + [C.c_float, C.c_float]
                                    + [C.c_void_p] * 5, C.c_int),
    "nunet_head_fwd": ([C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p], C.c_int),
    "nunet_loss_blocks_per_sample": ([C.c_int32, C.c_int32], C.c_int),
    "nunet_loss_iou_fwd": ([C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 6, C.c_int),
    "nunet_iou_counts": ([C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_loss_bwd": ([C.c_void_p] * 4 + [C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_loss_iou_fwd_multi": ([C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 6, C.c_int),
    "nunet_loss_bwd_multi": ([C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_head_bwd_params_rows": ([], C.c_int),
    "nunet_head_bwd_params": ([C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_void_p] * 3 + [C.c_int32, C.c_void_p], C.c_int),
    "nunet_grad_router_rows": ([C.c_int32] * 4, C.c_int),
    "nunet_grad_router": ([C.POINTER(RouterDesc), C.c_void_p], C.c_int),
    "nunet_up_transpose": ([C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p, C.c_void_p], C.c_int),
    "nunet_bn_backward_reduce": ([C.c_void_p, C.c_int32, C.c_int32, C.c_double] + [C.c_void_p] * 12, C.c_int),
    "nunet_bn_backward_from_sums": ([C.c_void_p, C.c_int32, C.c_double] + [C.c_void_p] * 8 + [C.c_float, C.c_void_p], C.c_int),
    "nunet_conv3x3_wgrad_plan": ([C.c_int32] * 5 + [C.POINTER(WgradPlan)], C.c_int),
    "nunet_conv3x3_wgrad": ([C.POINTER(WgradDesc), C.c_void_p], C.c_int),
    "nunet_lovasz_scratch_elems": ([C.c_int32, C.c_int32], C.c_longlong),
    "nunet_lovasz_hinge_fwd": ([C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5, C.c_int),
    "nunet_lovasz_hinge_bwd": ([C.c_void_p, C.c_void_p, C.c_float, C.c_longlong, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_sigmoid_u8": ([C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_input_u8": ([C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p], C.c_int),
    "nunet_optim_chunk_elems": ([], C.c_int),
    "nunet_optim_step": ([C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_bn_activate": ([C.POINTER(ActivateDesc), C.c_void_p], C.c_int),
    "nunet_bn_apply": ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int32, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_bn_bwd_apply": ([C.c_void_p] * 5 + [C.c_longlong, C.c_int32, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_upsample2x": ([C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p], C.c_int),
    "nunet_maxpool2x2": ([C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p], C.c_int),
    "nunet_input_to_nhwc32": ([C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p], C.c_int),
    "nunet_check_nchw_to_nhwc": ([C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p], C.c_int),
    "nunet_check_conv3x3": ([C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_check_bn": ([C.c_void_p, C.c_longlong, C.c_int32] + [C.c_void_p] * 4 + [C.c_float, C.c_float, C.c_int32] + [C.c_void_p] * 5, C.c_int),
    "nunet_check_relu_bn_bwd": ([C.c_void_p, C.c_void_p, C.c_longlong, C.c_int32] + [C.c_void_p] * 9, C.c_int),
    "nunet_check_conv3x3_dgrad": ([C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_check_conv3x3_wgrad": ([C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_check_route_bwd": ([C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_check_head_bwd": ([C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_int32] + [C.c_void_p] * 4, C.c_int),
    "nunet_check_bn_relu": ([C.c_void_p, C.c_longlong, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p], C.c_int),
    "nunet_check_route": ([C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p], C.c_int),
    "nunet_check_head": ([C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p], C.c_int),
    "nunet_bn_fold_eval": ([C.c_int32] + [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 3, C.c_int),
}


def lib():
    """Loads the shared library once; raises if it is not built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NunetError("libnunet_b200.so is not built: run `python pytorch_nested-unet_b200/build.py` "
                             "(or __graft_entry__.build()). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = res
        if l.nunet_version() != ABI_VERSION:
            raise NunetError("libnunet_b200.so has ABI version %d, this binding expects %d: rebuild the library" % (l.nunet_version(), ABI_VERSION))
        _lib = l
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc, what):
    if rc != 0:
        if rc >= 1000:
            raise NunetError("%s: CUDA error %d" % (what, rc - 1000))
        raise NunetError("%s: error %d (%s)" % (what, rc, {1: "bad argument", 2: "unsupported shape", 3: "shared memory"}.get(rc, "?")))
