"""ctypes binding of the C-ABI extension (``libb200dq.so``, declared in ``include/b200dq.h``).

The library is the product: there is NO fallback.  Importing this module on a machine where the
shared object has not been built raises, and every entry point raises on a non-zero status.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200dq.so")


class ExtensionMissing(RuntimeError):
    pass


class TapGemmDesc(C.Structure):
    _fields_ = [
        ("a_ptr", C.c_void_p), ("a_dims", C.c_longlong * 5), ("a_strides", C.c_longlong * 5),
        ("b_ptr", C.c_void_p), ("b_rows", C.c_longlong), ("b_k", C.c_longlong),
        ("b_batch", C.c_longlong), ("b_batch_stride", C.c_longlong),
        ("num_taps", C.c_int), ("kchunks", C.c_int),
        ("tap_c", C.c_int * 16), ("tap_w", C.c_int * 16), ("tap_p", C.c_int * 16),
        ("tap_h", C.c_int * 16), ("tap_bk", C.c_int * 16),
        ("TW", C.c_int), ("TH", C.c_int), ("TN", C.c_int),
        ("Wout", C.c_int), ("Hout", C.c_int), ("NB", C.c_int), ("Cout", C.c_int),
        ("out", C.c_void_p), ("oN", C.c_longlong), ("oH", C.c_longlong), ("oW", C.c_longlong),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("rN", C.c_longlong), ("rH", C.c_longlong), ("rW", C.c_longlong),
        ("alpha", C.c_float), ("out_f32", C.c_int), ("block_n", C.c_int), ("m_tiles_per_cta", C.c_int),
        ("relu", C.c_int),
    ]


class MmDesc(C.Structure):
    _fields_ = [
        ("a_ptr", C.c_void_p), ("a_dims", C.c_longlong * 5), ("a_strides", C.c_longlong * 5),
        ("b_ptr", C.c_void_p), ("b_dims", C.c_longlong * 5), ("b_strides", C.c_longlong * 5),
        ("a_mn", C.c_int), ("b_mn", C.c_int), ("ntaps", C.c_int), ("taps_per_cta", C.c_int),
        ("tap_c", C.c_int * 12), ("tap_w", C.c_int * 12), ("tap_p", C.c_int * 12), ("tap_h", C.c_int * 12),
        ("KW", C.c_int), ("KH", C.c_int), ("KN", C.c_int),
        ("ktiles_w", C.c_int), ("ktiles_h", C.c_int), ("kblocks", C.c_int),
        ("splits", C.c_int), ("batches", C.c_int),
        ("M", C.c_int), ("N", C.c_int),
        ("out", C.c_void_p), ("oZ", C.c_longlong), ("oT", C.c_longlong), ("oM", C.c_longlong),
        ("alpha", C.c_float), ("out_f32", C.c_int), ("block_n", C.c_int), ("b_strip", C.c_int),
        ("colsum", C.c_void_p),
    ]


_lib = None

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes (restype is always int status).  Mirrors include/b200dq.h.
class PconvTapsDesc(C.Structure):
    _fields_ = [
        ("a_ptr", C.c_void_p), ("a_dims", C.c_longlong * 5), ("a_strides", C.c_longlong * 5),
        ("b_ptr", C.c_void_p), ("b_k", C.c_longlong), ("kchunks", C.c_int), ("nr", C.c_int), ("ns", C.c_int),
        ("row_c", C.c_int * 9), ("row_p", C.c_int * 9), ("row_dh", C.c_int * 9), ("col_dw", C.c_int * 3),
        ("wcol", C.c_int * 27), ("NB", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("out", C.c_void_p), ("oN", C.c_longlong), ("oH", C.c_longlong), ("oW", C.c_longlong), ("bias", C.c_void_p)]


class Conv2dGeom(C.Structure):
    _fields_ = [("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
                ("ksize", C.c_int), ("stride", C.c_int)]


SIGNATURES = {
    "b2dq_version": [],
    # operator-level entries (csrc/oplevel.cu); the geometry struct is passed by reference
    "b2dq_conv2d_out_hw": [_vp, _vp],
    "b2dq_conv2d_fwd_workspace_bytes": [_vp],
    "b2dq_conv2d_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _ll, _vp],
    "b2dq_conv2d_dgrad": [_vp, _vp, _vp, _vp, _vp],
    "b2dq_conv2d_wgrad_workspace_bytes": [_vp, _i],
    "b2dq_conv2d_wgrad": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp],
    "b2dq_groupnorm_workspace_bytes": [_i, _i, _i, _i, _i],
    "b2dq_groupnorm_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _f, _i, _vp],
    "b2dq_groupnorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _vp],
    "b2dq_groupnorm_swish_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp],
    "b2dq_groupnorm_swish_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp],
    "b2dq_attention_workspace_bytes": [_i, _i, _i, _i],
    "b2dq_attention_fwd": [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _f, _vp],
    "b2dq_attention_bwd": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _f, _vp],
    "b2dq_vq_prepare_codebook": [_vp, _vp, _vp, _i, _i, _vp],
    "b2dq_vq_search_plan": [_i, _i, _i, _i, C.POINTER(C.c_int)],
    "b2dq_vq_search_workspace_bytes": [_i, _i],      # returns a byte count, not a status
    "b2dq_vq_search_gather": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _i, _i, _i, _i, _vp, _i, _vp],
    "b2dq_vq_ema_finalize": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _vp],
    "b2dq_vq_bwd": [_vp, _vp, _vp, _vp, _vp, _f, _vp, _ll, _i, _vp],
    "b2dq_tapgemm": [C.POINTER(TapGemmDesc), _vp],
    "b2dq_mmgemm": [C.POINTER(MmDesc), _vp],
    "b2dq_pconv3x3": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "b2dq_gn_finalize_tiles": [_vp, _vp, _i, _i, _i, _f, _vp],
    "b2dq_colsum_reduce": [_vp, _vp, _i, _i, _vp],
    "b2dq_wgrad_reduce": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b2dq_wgrad_reduce_bias": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "b2dq_gn_chunks": [_i, _i],
    "b2dq_gn_stats": [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "b2dq_gn_apply": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b2dq_gn_bwd_stats": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b2dq_gn_bwd_apply": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b2dq_gn_bwd_param": [_vp, _vp, _i, _i, _vp],
    "b2dq_gn_bwd_fused_workspace_bytes": [_i, _i, _i, _i],   # returns a byte count, not a status
    "b2dq_gn_bwd_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _vp],
    "b2dq_gn_bwd_fused_plan": [_i, _i, _i, C.POINTER(C.c_int)],
    "b2dq_gn_fwd_fused_workspace_bytes": [_i, _i, _i, _i],   # returns a byte count, not a status
    "b2dq_gn_fwd_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _f, _i, _vp],
    "b2dq_nchw_f32_to_nhwc_bf16": [_vp, _vp, _i, _i, _i, _vp],
    "b2dq_nhwc_bf16_to_nchw_f32": [_vp, _vp, _i, _i, _i, _vp],
    "b2dq_nhwc_f32_to_nchw_f32": [_vp, _vp, _i, _i, _i, _vp],
    "b2dq_nchw_f32_to_nhwc_f32": [_vp, _vp, _i, _i, _i, _vp],
    "b2dq_upsample2x": [_vp, _vp, _i, _i, _i, _i, _vp],
    "b2dq_upsample2x_bwd": [_vp, _vp, _i, _i, _i, _i, _vp],
    "b2dq_softmax_rows": [_vp, _vp, _ll, _i, _i, _vp],
    "b2dq_softmax_bwd_rows": [_vp, _vp, _vp, _ll, _i, _f, _vp],
    "b2dq_add_bf16": [_vp, _vp, _vp, _ll, _vp],
    "b2dq_im2col3x3_small": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "b2dq_pack_weights": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "b2dq_pack_weights_multi": [_vp, _i, _ll, _i, _vp],
    "b2dq_pconv_taps": [_vp, _i, _vp],
    "b2dq_upconv_pack": [_vp, _vp, _vp, _i, _i, _vp],
    "b2dq_upconv_wgrad_reduce": [_vp, _vp, _i, _i, _i, _vp],
    "b2dq_lpips_head_chunks": [_i, _i],              # returns a count, not a status
    "b2dq_lpips_head_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _f, _vp],
    "b2dq_lpips_head_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _f, _vp],
    "b2dq_maxpool2x2": [_vp, _vp, _i, _i, _i, _i, _vp],
    "b2dq_maxpool2x2_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "b2dq_relu_bwd": [_vp, _vp, _vp, _ll, _vp],
    "b2dq_lrelu_bwd": [_vp, _vp, _vp, _ll, _f, _vp],
    "b2dq_im2col_window": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "b2dq_bias_grad_blocks": [_ll],
    "b2dq_bias_grad": [_vp, _vp, _vp, _ll, _i, _vp],
    "b2dq_cast_f32_to_bf16": [_vp, _vp, _ll, _vp],
    "b2dq_patch_entropy": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "b2dq_permuter_forward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                              C.POINTER(C.c_longlong), _vp],
    "b2dq_permuter_backward": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _ll, _vp],
}


def lib():
    """Load the extension (once).  Raises ExtensionMissing if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU/PyTorch fallback for the DQ-VAE hot path.")
        l = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = l
    return _lib


def check(status, what):
    if status != 0:
        raise RuntimeError(f"{what} failed with status {status}")
