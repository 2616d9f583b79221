"""Operator-level C entry points (include/b200dq.h "operator-level entry points", csrc/oplevel.cu) against the
descriptor-level route the autograd Functions take (kernels.py): same kernels, same descriptors, so the results are
required to be bit-identical, and both are anchored to the fp32 torch operator the reference calls
(nn.Conv2d / GroupNorm / AttnBlock, modules/diffusionmodules/model.py:29-35,88-115,170-188)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def _mods():
    from dynamicvectorquantization_b200 import kernels as kn, oplevel as op
    return kn, op


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).pow(2).sum() / b.pow(2).sum().clamp_min(1e-30)).sqrt().item()


CONV_CASES = [
    # n, h, w, cin, cout, ksize, stride
    (4, 128, 128, 128, 128, 3, 1),      # persistent strip kernel (+ GroupNorm statistics from its epilogue)
    (2, 32, 32, 64, 256, 3, 1),
    (2, 32, 32, 256, 256, 1, 1),
    (3, 16, 16, 512, 512, 3, 1),
    (2, 64, 64, 128, 128, 3, 2),        # Downsample: pad (0,1,0,1), stride 2
    (1, 24, 40, 64, 192, 3, 1),         # ragged tiles
    (3, 256, 256, 128, 128, 3, 2),      # Downsample at full width: data gradient on the strip kernel (b2dq_pconv_taps)
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_forward_and_gradients_match_the_descriptor_route_and_torch(case):
    kn, op = _mods()
    n, h, w, cin, cout, k, stride = case
    torch.manual_seed(cin + cout + k + stride)
    dev = "cuda"
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, k, k, device=dev) * (cin * k * k) ** -0.5
    b = torch.randn(cout, device=dev) * 0.1
    xh = _nhwc(x).to(BF16)
    wf, wd = kn.pack_weights(wt)
    ho, wo = h // stride, w // stride
    res = torch.randn(n, ho, wo, cout, device=dev).to(BF16)

    y_ref = kn.conv_fwd(xh, wf, b, k, stride, cout, residual=res)
    ref_stats = kn.last_conv_stats
    y, stats = op.conv2d_fwd(xh, wf, b, k, stride, cout, residual=res, want_stats=True)
    assert torch.equal(y, y_ref)
    assert (stats is None) == (ref_stats is None)
    if stats is not None:
        assert torch.equal(stats, ref_stats)
    # fp32 operator of the reference on the same bf16 operands
    xr, wr = xh.float().permute(0, 3, 1, 2), wt.to(BF16).float()
    if stride == 1:
        t = F.conv2d(xr, wr, b, padding=k // 2)
    else:
        t = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wr, b, stride=2)
    t = _nhwc(t) + res.float()
    assert _rel(y, t) < 6e-3

    dy = torch.randn(n, ho, wo, cout, device=dev).to(BF16)
    dx_ref = kn.conv_dgrad(dy, wd, k, stride, cin, (h, w))
    dx = op.conv2d_dgrad(dy, wd, k, stride, cin, (h, w))
    assert torch.equal(dx, dx_ref)

    dw_ref, db_ref = kn.conv_wgrad(xh, dy, k, stride, want_bias=True)
    dw, db = op.conv2d_wgrad(xh, dy, k, stride, want_bias=True)
    assert torch.equal(dw, dw_ref) and torch.equal(db, db_ref)
    assert torch.equal(op.conv2d_wgrad(xh, dy, k, stride), dw_ref)
    # against autograd of the fp32 operator
    xr = xr.clone().requires_grad_(True)
    wr = wr.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    if stride == 1:
        t = F.conv2d(xr, wr, br, padding=k // 2)
    else:
        t = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wr, br, stride=2)
    t.backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(dx, _nhwc(xr.grad)) < 6e-3
    assert _rel(dw, wr.grad) < 2e-3
    assert _rel(db, br.grad) < 1e-4


def test_conv2d_rejects_bad_geometry_and_small_workspace():
    import ctypes as C
    from dynamicvectorquantization_b200 import _cabi
    lib = _cabi.lib()
    g = _cabi.Conv2dGeom(2, 31, 32, 64, 64, 3, 2)          # odd H under stride 2
    out = (C.c_int * 2)()
    assert lib.b2dq_conv2d_out_hw(C.byref(g), out) == -1
    g = _cabi.Conv2dGeom(2, 32, 32, 64, 64, 5, 1)          # 5x5 is not a model layer
    assert lib.b2dq_conv2d_fwd_workspace_bytes(C.byref(g)) == -1
    g = _cabi.Conv2dGeom(2, 32, 32, 64, 64, 3, 1)
    x = torch.zeros(2, 32, 32, 64, device="cuda", dtype=BF16)
    dy = torch.zeros(2, 32, 32, 64, device="cuda", dtype=BF16)
    dw = torch.zeros(64, 64, 3, 3, device="cuda")
    ws = torch.zeros(16, device="cuda", dtype=torch.uint8)
    assert lib.b2dq_conv2d_wgrad(C.byref(g), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, ws.data_ptr(), 16,
                                 None) == -2
    st = torch.zeros(2, 32, 2, device="cuda")
    assert lib.b2dq_conv2d_fwd(C.byref(g), x.data_ptr(), x.data_ptr(), None, None, dy.data_ptr(), 0, st.data_ptr(),
                               None, 0, None) == -3   # this shape's kernel does not emit statistics


@pytest.mark.parametrize("shape,act", [((4, 64, 64, 128), 1), ((2, 32, 32, 256), 0), ((3, 16, 16, 512), 1),
                                       ((2, 12, 12, 64), 1), ((2, 10, 10, 32), 1)])   # the last one: 1 channel per group, statistics + apply pair
def test_groupnorm_forward_backward_match_the_descriptor_route_and_torch(shape, act):
    kn, op = _mods()
    torch.manual_seed(sum(shape))
    n, h, w, c = shape
    dev = "cuda"
    x = (torch.randn(n, h, w, c, device=dev) * 1.5 + 0.3).to(BF16)
    gamma = 1 + 0.2 * torch.randn(c, device=dev)
    beta = 0.1 * torch.randn(c, device=dev)
    dy = torch.randn(n, h, w, c, device=dev).to(BF16)
    add = torch.randn(n, h, w, c, device=dev).to(BF16)

    y_ref, st_ref = kn.gn_forward(x, gamma, beta, act)
    y, st = op.groupnorm_fwd(x, gamma, beta, act)
    assert torch.equal(y, y_ref) and torch.equal(st, st_ref)
    dx_ref, dg_ref, db_ref = kn.gn_bwd(dy, x, st_ref, gamma, beta, act, add=add)
    dx, dg, db = op.groupnorm_bwd(dy, x, st, gamma, beta, act, add=add)
    assert torch.equal(dx, dx_ref) and torch.equal(dg, dg_ref) and torch.equal(db, db_ref)

    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    t = F.group_norm(xr, 32, gr, br, eps=1e-6)
    if act:
        t = t * torch.sigmoid(t)
    t.backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(y, _nhwc(t.detach())) < 4e-3
    assert _rel(dx.float() - add.float(), _nhwc(xr.grad)) < 1.5e-2
    assert _rel(dg, gr.grad) < 3e-3 and _rel(db, br.grad) < 3e-3


@pytest.mark.parametrize("n,t,c", [(2, 256, 256), (2, 1024, 256), (3, 64, 512)])
def test_attention_core_matches_the_descriptor_route_and_torch(n, t, c):
    kn, op = _mods()
    from dynamicvectorquantization_b200 import ops
    torch.manual_seed(t + c)
    dev = "cuda"
    qkv = torch.randn(n, t, 3 * c, device=dev).to(BF16)
    do = torch.randn(n, t, c, device=dev).to(BF16)
    scale = float(c) ** -0.5

    # the route AttnQKVFn takes
    q, k, v = qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]
    s = torch.empty(n, t, t, dtype=torch.float32, device=dev)
    ops._mm(q, False, k, False, t, t, c, s, alpha=scale, out_f32=True)
    p_ref = kn.softmax_rows(s, t)
    o_ref = torch.empty(n, t, c, dtype=BF16, device=dev)
    ops._mm(p_ref, False, v, True, t, c, t, o_ref)
    dqkv_ref = torch.empty_like(qkv)
    ops._mm(p_ref, True, do, True, t, c, t, dqkv_ref[..., 2 * c:])
    dp = torch.empty(n, t, t, dtype=BF16, device=dev)
    ops._mm(do, False, v, False, t, t, c, dp)
    ds = kn.softmax_bwd_rows(p_ref, dp, t, scale)
    ops._mm(ds, False, k, True, t, c, t, dqkv_ref[..., :c])
    ops._mm(ds, True, q, True, t, c, t, dqkv_ref[..., c:2 * c])

    o, p = op.attention_fwd(qkv)
    assert torch.equal(p, p_ref) and torch.equal(o, o_ref)
    dqkv = op.attention_bwd(qkv, p, do)
    assert torch.equal(dqkv, dqkv_ref)

    qr = qkv.float().clone().requires_grad_(True)
    qq, kk, vv = qr[..., :c], qr[..., c:2 * c], qr[..., 2 * c:]
    tt = torch.softmax(qq @ kk.transpose(1, 2) * scale, dim=-1) @ vv
    tt.backward(do.float())
    assert _rel(o, tt.detach()) < 8e-3
    assert _rel(dqkv, qr.grad) < 2e-2
