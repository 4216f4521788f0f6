"""ctypes binding of the C ABI declared in include/cervix_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``csrc/libcervix_b200.so``.  There is no fallback: if the library is missing the import of any
compute path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcervix_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
EUNSUPPORTED = -3


class ConvDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "h", "w", "cin", "cout", "kh", "kw", "stride", "pad", "dil",
                                          "ho", "wo", "dtype")]


class GemmProblem(C.Structure):
    """cvx_gemm_problem of include/cervix_b200.h."""
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("bias", C.c_void_p), ("c", C.c_void_p), ("rowsum", C.c_void_p),
                ("lda_m", C.c_int64), ("lda_k", C.c_int64), ("ldb_n", C.c_int64), ("ldb_k", C.c_int64), ("ldc", C.c_int64),
                ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("reserved", C.c_int32)]


class AugSample(C.Structure):
    """cvx_aug_sample of include/cervix_b200.h (the host packs these as numpy records: utils/dataloader.py AUG_SAMPLE)."""
    _fields_ = [("src_off", C.c_int64), ("lab_off", C.c_int64), ("tmp_off", C.c_int64), ("xtab", C.c_int32), ("ytab", C.c_int32),
                ("xnn", C.c_int32), ("ynn", C.c_int32), ("rot", C.c_int32), ("lut", C.c_int32), ("ih", C.c_int32),
                ("iw", C.c_int32), ("nh", C.c_int32), ("nw", C.c_int32), ("xtaps", C.c_int32), ("ytaps", C.c_int32),
                ("dx", C.c_int32), ("dy", C.c_int32), ("flip", C.c_int32), ("blur", C.c_int32), ("rotate", C.c_int32),
                ("reserved", C.c_int32)]


MAX_GEMM_PROBLEMS = 16
MAX_PARAM_SETS = 8


class ParamSets(C.Structure):
    """cvx_param_sets of include/cervix_b200.h."""
    _fields_ = [("w", C.c_void_p * MAX_PARAM_SETS), ("b", C.c_void_p * MAX_PARAM_SETS), ("dw", C.c_void_p * MAX_PARAM_SETS),
                ("db", C.c_void_p * MAX_PARAM_SETS), ("row_start", C.c_int32 * (MAX_PARAM_SETS + 1)), ("sets", C.c_int32)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_D = C.POINTER(ConvDesc)

# name -> argtypes (restype is always int unless listed in _RESTYPES)
PROTOTYPES = {
    "cvx_abi_version": [],
    "cvx_last_error": [],
    "cvx_launch_count": [],
    "cvx_device_is_sm100": [],
    "cvx_nchw_to_nhwc": [_P, _P, _I, _I, _I, _I, _I, _P],
    "cvx_nhwc_to_nchw": [_P, _P, _I, _I, _I, _I, _I, _P],
    "cvx_pack_weight": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cvx_unpack_wgrad": [_P, _P, _I, _I, _I, _I, _P],
    "cvx_pack_dw_weight": [_P, _P, _I, _P],
    "cvx_unpack_dw_wgrad": [_P, _P, _I, _P],
    "cvx_copy_channels": [_P, _I, _I, _P, _I, _I, _L, _I, _I, _P],
    "cvx_conv_fwd": [_D, _P, _P, _P, _P, _P],
    "cvx_conv_dgrad": [_D, _P, _P, _P, _P],
    "cvx_conv_wgrad": [_D, _P, _P, _P, _P],
    "cvx_bias_grad": [_P, _P, _P, _L, _I, _I, _P],
    "cvx_conv_fwd_tc": [_D, _P, _P, _P, _P, _P],
    "cvx_conv_dgrad_tc": [_D, _P, _P, _P, _P],
    "cvx_conv_wgrad_tc": [_D, _P, _P, _P, _P],
    "cvx_conv_tc_set_pairs": [C.c_int, C.c_int],
    "cvx_im2col_narrow": [_D, _P, _P, _I, _P],
    "cvx_subsample": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cvx_subsample_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cvx_dwconv_fwd": [_D, _P, _P, _P, _I, _P],
    "cvx_dwconv_bwd_data": [_D, _P, _P, _P, _P, _I, _P],
    "cvx_dwconv_bwd_weight": [_D, _P, _P, _P, _P, _I, _P],
    "cvx_bn_forward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _F, _F, _P],
    "cvx_bn_backward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    "cvx_relu_fwd": [_P, _P, _L, _I, _P],
    "cvx_relu_bwd": [_P, _P, _P, _L, _I, _P],
    "cvx_add": [_P, _P, _P, _L, _I, _P],
    "cvx_spatial_reduce": [_P, _P, _I, _I, _I, _F, _I, _P],
    "cvx_spatial_broadcast": [_P, _P, _I, _I, _I, _F, _I, _P],
    "cvx_upsample_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cvx_upsample_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cvx_upsample_into": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "cvx_upsample_from_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "cvx_upsample_to_nchw_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cvx_upsample_to_nchw_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cvx_maxpool3x3s2_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "cvx_maxpool3x3s2_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cvx_dropout_fwd": [_P, _P, _P, _L, _F, C.c_uint64, _P, _I, _P],
    "cvx_dropout_bwd": [_P, _P, _P, _L, _F, _I, _P],
    "cvx_dropout_bwd_seeded": [_P, _P, _L, _F, C.c_uint64, _P, _I, _P],
    "cvx_seg_loss_stats": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P],
    "cvx_seg_loss_finalize": [_P, _P, _I, _F, _F, _P],
    "cvx_seg_loss_grad": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _P],
    "cvx_conv_fwd_tc_ex": [_D, _P, _P, _P, _P, _P, _P, _P, _P],
    "cvx_conv_fwd_tc_act": [_D, _P, _P, _P, _P, _P, _I, _P, _P],
    "cvx_conv_dgrad_tc_ex": [_D, _P, _P, _P, _P, _P, _P, _P],
    "cvx_dwf_fwd": [_D, _P, _P, _P, _P, _I, _P, _P, _P],
    "cvx_dwf_bwd": [_D, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "cvx_bn_stats": [_P, _P, _L, _I, _I, _P],
    "cvx_bn_affine": [_P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _P],
    "cvx_pw_fold": [_P, _P, _P, _P, _P, _P, _I, _I, _P],
    "cvx_affine_act": [_P, _P, _P, _P, _P, _L, _I, _I, _P],
    "cvx_bn_bwd_sums": [_P, _P, _P, _P, _L, _I, _I, _P],
    "cvx_bn_bwd_coef": [_P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "cvx_bn_bwd_affine": [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "cvx_pw_bwd_coef": [_P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "cvx_seg_layernorm_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P],
    "cvx_seg_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P],
    "cvx_gelu_fwd": [_P, _P, _L, _P],
    "cvx_gelu_bwd": [_P, _P, _P, _L, _P],
    "cvx_graph_gather": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "cvx_gate_pool_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "cvx_gate_pool_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "cvx_attn_small_fwd": [_P, _P, _P, _I, _I, _I, _I, _F, _F, C.c_uint64, _P, _P],
    "cvx_attn_small_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, C.c_uint64, _P, _P],
    "cvx_l2norm_fwd": [_P, _P, _P, _I, _I, _P],
    "cvx_l2norm_bwd": [_P, _P, _P, _P, _I, _I, _P],
    "cvx_rows_gather": [_P, _P, _P, _P, _I, _I, _P],
    "cvx_rows_scatter_add": [_P, _P, _P, _P, _I, _I, _I, _P],
    "cvx_softmax_ce": [_P, _P, _P, _P, _I, _I, _F, _P],
    "cvx_masked_mse": [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _P],
    "cvx_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P],
    "cvx_adam_step_dev": [_P, _P, _P, _P, _L, _P, _P, _P],
    "cvx_sgd_step": [_P, _P, _P, _L, _F, _F, _F, _I, _I, _F, _P],
    "cvx_sgd_step_dev": [_P, _P, _P, _L, _P, _I, _P],
    "cvx_gemm_grouped": [C.POINTER(GemmProblem), _I, _P],
    "cvx_segtab_layernorm_fwd": [_P, C.POINTER(ParamSets), _P, _P, _P, _P, _P, _I, _I, _F, _I, _P],
    "cvx_segtab_layernorm_bwd": [_P, _P, C.POINTER(ParamSets), _P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P],
    "cvx_segtab_gate_pool_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "cvx_segtab_gate_pool_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "cvx_segtab_bcast_add": [_P, _P, _P, _P, _P, _I, _I, _P],
    "cvx_segtab_bcast_add_bwd": [_P, _P, _P, _P, _P, _I, _I, _P],
    "cvx_multi_gather_chunk": [],
    "cvx_set_ws_prezeroed": [_I],
    "cvx_debug_tc_trace": [_I, _P],
    "cvx_aug_resize_rows": [_P, _I, _P, _P, _P, _L, _P],
    "cvx_aug_compose": [_P, _I, _P, _P, _P, _P, _P, _I, _I, _P],
    "cvx_aug_blur5": [_P, _I, _P, _P, _I, _I, _P],
    "cvx_aug_rotate_jitter": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "cvx_multi_gather": [_P, _P, _P, _P, _P, _I, _P, _P],
    "cvx_seg_postprocess": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "cvx_confusion_matrix": [_P, _P, _L, _I, _P, _P],
    "cvx_split_patches": [_P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _P],
    "cvx_split_patches_u8": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, C.POINTER(C.c_float),
                             C.POINTER(C.c_float), _I, _P],
    "cvx_finish_batch_u8": [_P, _P, _L, _P, _P, _L, _I, _I, _P],
}
_RESTYPES = {"cvx_last_error": C.c_char_p, "cvx_launch_count": C.c_int64}

_lib = None


class CervixError(RuntimeError):
    pass


def load():
    """Load libcervix_b200.so and bind every symbol of include/cervix_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CervixError(
            "cervix_b200: %s not found - build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.cvx_abi_version() != 1:
        raise CervixError("cervix_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = _lib.cvx_last_error().decode("utf-8", "replace") if _lib is not None else ""
        raise CervixError("%s failed (rc=%d): %s" % (what, rc, msg))
