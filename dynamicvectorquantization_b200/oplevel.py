"""ctypes wrappers of the operator-level C entry points (include/b200dq.h, csrc/oplevel.cu).

One C call per operator: the geometry goes in as a struct, the library picks tiles, tap tables, split-K factors and
the kernel.  `kernels.py` makes the same choices in Python for the autograd Functions of `ops.py`; these wrappers are
the route a non-Python host would take, and `tests/test_gpu_oplevel.py` holds both routes to bit-identical results.
"""
import ctypes as C

import torch

from . import _cabi
from .kernels import BF16, _ptr, _stream, check


def _geom(n, h, w, cin, cout, ksize, stride):
    return _cabi.Conv2dGeom(n, h, w, cin, cout, ksize, stride)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def conv2d_fwd(x, wpack, bias, ksize, stride, cout, residual=None, act=0, want_stats=False):
    """x NHWC bf16 -> y NHWC bf16 (and the GroupNorm(32) statistics of y when want_stats and the shape emits them)."""
    nb, h, w, cin = x.shape
    g = _geom(nb, h, w, cin, cout, ksize, stride)
    lib = _cabi.lib()
    hw = (C.c_int * 2)()
    check(lib.b2dq_conv2d_out_hw(C.byref(g), hw), "conv2d_out_hw")
    y = torch.empty(nb, hw[0], hw[1], cout, dtype=BF16, device=x.device)
    ws_bytes = lib.b2dq_conv2d_fwd_workspace_bytes(C.byref(g))
    stats = None
    if want_stats and ws_bytes > 0:
        stats = torch.empty(nb, 32, 2, dtype=torch.float32, device=x.device)
    ws = _ws(ws_bytes, x.device)
    check(lib.b2dq_conv2d_fwd(C.byref(g), _ptr(x), _ptr(wpack), _ptr(bias), _ptr(residual), _ptr(y), int(act),
                              _ptr(stats), _ptr(ws), ws_bytes, _stream()), "conv2d_fwd")
    return (y, stats) if want_stats else y


def conv2d_dgrad(dy, wpack_dgrad, ksize, stride, cin, in_hw):
    nb, ho, wo, cout = dy.shape
    h, w = in_hw
    g = _geom(nb, h, w, cin, cout, ksize, stride)
    dx = torch.empty(nb, h, w, cin, dtype=BF16, device=dy.device)
    check(_cabi.lib().b2dq_conv2d_dgrad(C.byref(g), _ptr(dy), _ptr(wpack_dgrad), _ptr(dx), _stream()), "conv2d_dgrad")
    return dx


def conv2d_wgrad(x, dy, ksize, stride, want_bias=False):
    nb, h, w, cin = x.shape
    cout = dy.shape[-1]
    g = _geom(nb, h, w, cin, cout, ksize, stride)
    lib = _cabi.lib()
    ws_bytes = lib.b2dq_conv2d_wgrad_workspace_bytes(C.byref(g), int(want_bias))
    assert ws_bytes >= 0
    ws = _ws(ws_bytes, x.device)
    dw = torch.empty(cout, cin, ksize, ksize, dtype=torch.float32, device=x.device)
    db = torch.empty(cout, dtype=torch.float32, device=x.device) if want_bias else None
    check(lib.b2dq_conv2d_wgrad(C.byref(g), _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), _ptr(ws), ws_bytes, _stream()),
          "conv2d_wgrad")
    return (dw, db) if want_bias else dw


def groupnorm_fwd(x, gamma, beta, act, groups=32, eps=1e-6):
    nb, h, w, c = x.shape
    lib = _cabi.lib()
    ws_bytes = lib.b2dq_groupnorm_workspace_bytes(nb, h * w, c, groups, 0)
    assert ws_bytes >= 0
    ws = _ws(ws_bytes, x.device)
    y = torch.empty_like(x)
    stats = torch.empty(nb, groups, 2, dtype=torch.float32, device=x.device)
    check(lib.b2dq_groupnorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(stats), _ptr(ws), ws_bytes, nb, h * w,
                                 c, groups, float(eps), int(act), _stream()), "groupnorm_fwd")
    return y, stats


def groupnorm_bwd(dy, x, stats, gamma, beta, act, groups=32, add=None):
    nb, h, w, c = x.shape
    lib = _cabi.lib()
    ws_bytes = lib.b2dq_groupnorm_workspace_bytes(nb, h * w, c, groups, 1)
    assert ws_bytes >= 0
    ws = _ws(ws_bytes, x.device)
    dx = torch.empty_like(x)
    dgb = torch.empty(2, c, dtype=torch.float32, device=x.device)
    check(lib.b2dq_groupnorm_bwd(_ptr(dy), _ptr(x), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(dx), _ptr(dgb), _ptr(add),
                                 _ptr(ws), ws_bytes, nb, h * w, c, groups, int(act), _stream()), "groupnorm_bwd")
    return dx, dgb[0], dgb[1]


def attention_fwd(qkv, scale=None):
    """qkv [N,T,3C] bf16 (q | k | v) -> (out [N,T,C], probs [N,T,T])."""
    nb, t, c3 = qkv.shape
    c = c3 // 3
    scale = float(int(c) ** -0.5) if scale is None else float(scale)
    lib = _cabi.lib()
    ws_bytes = lib.b2dq_attention_workspace_bytes(nb, t, c, 0)
    assert ws_bytes >= 0
    ws = _ws(ws_bytes, qkv.device)
    out = torch.empty(nb, t, c, dtype=BF16, device=qkv.device)
    probs = torch.empty(nb, t, t, dtype=BF16, device=qkv.device)
    check(lib.b2dq_attention_fwd(_ptr(qkv), _ptr(out), _ptr(probs), _ptr(ws), ws_bytes, nb, t, c, scale, _stream()),
          "attention_fwd")
    return out, probs


def attention_bwd(qkv, probs, dout, scale=None):
    nb, t, c3 = qkv.shape
    c = c3 // 3
    scale = float(int(c) ** -0.5) if scale is None else float(scale)
    lib = _cabi.lib()
    ws_bytes = lib.b2dq_attention_workspace_bytes(nb, t, c, 1)
    assert ws_bytes >= 0
    ws = _ws(ws_bytes, qkv.device)
    dqkv = torch.empty_like(qkv)
    check(lib.b2dq_attention_bwd(_ptr(qkv), _ptr(probs), _ptr(dout), _ptr(dqkv), _ptr(ws), ws_bytes, nb, t, c, scale,
                                 _stream()), "attention_bwd")
    return dqkv
