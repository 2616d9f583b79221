"""torch.autograd.Function wrappers around the C-ABI kernels.

Internal activation format: NHWC bf16 ([N,H,W,C] contiguous CUDA tensors).  Parameters stay the
reference's fp32 OIHW ``nn.Parameter``s (master copy, optimizer-visible); their bf16 GEMM packings
are derived caches keyed on the parameter's version counter.  All Functions are re-entrant (saved
tensors are never modified), because the reference loss calls ``torch.autograd.grad(...,
retain_graph=True)`` on ``decoder.conv_out.weight`` before ``backward``
(modules/losses/vqperceptual_multidisc.py:102-113).
"""
import weakref

import torch

from . import kernels as kn

BF16 = torch.bfloat16

FOLD_UPSAMPLE = True   # Upsample(with_conv): four parity-class 2x2 convolutions instead of upsample + conv3x3

# GroupNorm statistics produced by the epilogue of the conv that made a tensor: (weakref(tensor), stats).
# Consumed by the very next gn_swish on THAT tensor object; anything else recomputes them.
_pending_stats = None

def _packed(weight, kind):
    """bf16 GEMM packing of an OIHW fp32 parameter, cached ON the parameter object (so the cache
    dies with it) and validated against the storage pointer and the in-place version counter."""
    cache = weight.__dict__.setdefault("_b2_packs", {})
    ent = cache.get(kind)
    stamp = (weight.data_ptr(), weight._version, weight.device)
    if ent is not None and ent[0] == stamp:
        return ent[1]
    w = weight.detach()
    co, ci, r, s = w.shape
    if kind in ("fwd", "dgrad") and w.is_cuda and w.dtype == torch.float32:
        # one launch makes both packings when the data gradient will be wanted too (training)
        other = "dgrad" if kind == "fwd" else "fwd"
        both = weight.requires_grad or kind == "dgrad"   # (grad mode is off inside Function.forward)
        oent = cache.get(other)
        need_other = both and not (oent is not None and oent[0] == stamp)
        fwd, dgr = kn.pack_weights(w, want_fwd=(kind == "fwd" or need_other), want_dgrad=(kind == "dgrad" or need_other))
        if need_other:
            cache[other] = (stamp, dgr if other == "dgrad" else fwd)
        p = fwd if kind == "fwd" else dgr
    elif kind in ("up_fwd", "up_dgrad"):            # nearest x2 + conv3x3 folded to four 2x2 parity convolutions
        fwd, dgr = kn.upconv_pack(w)
        other = "up_dgrad" if kind == "up_fwd" else "up_fwd"
        cache[other] = (stamp, dgr if other == "up_dgrad" else fwd)
        p = fwd if kind == "up_fwd" else dgr
    elif kind == "fwd":
        p = kn.pack_weight_fwd(w)
    elif kind == "dgrad":
        p = kn.pack_weight_dgrad(w)
    elif kind == "fwd_pad16":                       # Cout < 16 (conv_out): pad rows to 16
        p = torch.zeros(16, r * s * ci, dtype=BF16, device=w.device)
        p[:co] = kn.pack_weight_fwd(w)
    elif kind == "col_fwd":                         # Cin*9 <= 64 (conv_in): [Cout, 64], col = t*Cin + c
        p = torch.zeros(co, 64, dtype=BF16, device=w.device)
        p[:, :r * s * ci] = w.permute(0, 2, 3, 1).reshape(co, -1).to(BF16)
    elif kind == "col_dgrad":                       # Cout*9 <= 64 (conv_out): [Cin, 64], col = t*Cout + co
        p = torch.zeros(ci, 64, dtype=BF16, device=w.device)
        p[:, :r * s * co] = w.permute(1, 2, 3, 0).reshape(ci, -1).to(BF16)
    elif kind == "col_dgrad_x2":                    # col_dgrad with every tap column doubled (hi/lo halves of dy)
        p = torch.zeros(ci, 64, dtype=BF16, device=w.device)
        p[:, :2 * r * s * co] = w.permute(1, 2, 3, 0).reshape(ci, r * s, 1, co).expand(ci, r * s, 2, co).reshape(ci, -1).to(BF16)
    else:
        raise ValueError(kind)
    cache[kind] = (stamp, p)
    return p


def prepack(module):
    """Refresh every stale "fwd" / "dgrad" packing of `module`'s convolution weights in ONE launch
    (b2dq_pack_weights_multi).  After an optimizer step all of them are stale and the lazy path of `_packed` would
    repack them one launch per layer; a forward that starts with `prepack(self)` finds them fresh instead.  Only
    packings the lazy path has created before are refreshed (the cache is the registry), into the same buffers, so the
    first step and modules called on their own behave exactly as without this call."""
    todo = []
    for p in module.parameters():
        cache = p.__dict__.get("_b2_packs")
        if not cache or not p.is_cuda or p.dtype != torch.float32 or p.dim() != 4:
            continue
        stamp = (p.data_ptr(), p._version, p.device)
        f, d = cache.get("fwd"), cache.get("dgrad")
        sf = f is not None and f[0] != stamp and f[0][2] == p.device
        sd = d is not None and d[0] != stamp and d[0][2] == p.device
        if sf or sd:
            todo.append((p, cache, stamp, f[1] if sf else None, d[1] if sd else None))
    if len(todo) < 2:
        return 0
    key = tuple((p.data_ptr(), 0 if f is None else f.data_ptr(), 0 if d is None else d.data_ptr())
                for p, _, _, f, d in todo)
    tabs = module.__dict__.setdefault("_b2_pack_tables", {})
    ent = tabs.get(key)
    if ent is None:
        if torch.cuda.is_current_stream_capturing():
            return 0                                  # the table upload cannot be captured: lazy path this time
        rows, start, max_rs = [], 0, 1
        for p, _, _, f, d in todo:
            co, ci, r, s_ = p.shape
            rs, tci = r * s_, (ci + 31) // 32
            if rs > 16:
                return 0
            max_rs = max(max_rs, rs)
            rows.append([key[len(rows)][0], key[len(rows)][1], key[len(rows)][2], co | (ci << 32), rs | (tci << 32), start])
            start += ((co + 31) // 32) * tci
        ent = (torch.tensor(rows, dtype=torch.int64).to(todo[0][0].device), start, max_rs)
        if len(tabs) >= 4:                            # e.g. autoencoder / discriminator weights in alternation
            tabs.clear()
        tabs[key] = ent
    table, total, max_rs = ent
    kn.pack_weights_multi(table, len(todo), total, max_rs)
    for p, cache, stamp, f, d in todo:
        if f is not None:
            cache["fwd"] = (stamp, f)
        if d is not None:
            cache["dgrad"] = (stamp, d)
    return len(todo)


def invalidate_caches(module):
    """Drop every derived bf16 copy (GEMM weight packings, VQ search codebook + norms) held for `module`'s
    parameters.  The caches are validated by (data_ptr, tensor._version), and a write THROUGH `.data`
    (`p.data.copy_(...)`, `dist.broadcast(p.data, 0)`, `nn.init.*_(p.data)`, EMA weight swaps such as LitEma) does
    not bump the parameter's version counter: call this after any such write that happens after the first forward.
    Ordinary optimizer steps, `load_state_dict` and `p.copy_()` under no_grad do bump it and need nothing."""
    for p in module.parameters():
        p.__dict__.pop("_b2_packs", None)
    for m in module.modules():
        m.__dict__.pop("_b2_pack_tables", None)
        if hasattr(m, "_cb_key"):
            m._cb_key = None


def _f32(t):
    return None if t is None else t.detach().float().contiguous()


class Conv2dFn(torch.autograd.Function):
    """NHWC bf16 convolution (3x3 s1 p1 | 1x1 | 3x3 s2 with the Downsample padding) + bias
    (+ residual, which may be broadcast over the batch)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, ksize, stride):
        cout, cin = weight.shape[0], weight.shape[1]
        y = kn.conv_fwd(x, _packed(weight, "fwd"), _f32(bias), ksize, stride, cout,
                        residual=None if residual is None else _bcast_res(residual, x.shape[0]))
        ctx.save_for_backward(x, weight)
        ctx.meta = (ksize, stride, cin, cout, bias is not None,
                    None if residual is None else tuple(residual.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        ksize, stride, cin, cout, has_bias, res_shape = ctx.meta
        dy = dy.contiguous()
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = kn.conv_dgrad(dy, _packed(weight, "dgrad"), ksize, stride, cin, x.shape[1:3])
        want_db = has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            if want_db:
                dw, db = kn.conv_wgrad(x, dy, ksize, stride, want_bias=True)
            else:
                dw = kn.conv_wgrad(x, dy, ksize, stride)
        elif want_db:
            db = kn.bias_grad(dy)
        if res_shape is not None and ctx.needs_input_grad[3]:
            dres = dy if res_shape[0] == dy.shape[0] else dy.float().sum(0, keepdim=True).to(BF16)
        return dx, dw, db, dres, None, None


def _bcast_res(res, nb):
    if res.shape[0] == nb:
        return res.contiguous()
    return res.expand(nb, *res.shape[1:]).contiguous()


class ConvInFn(torch.autograd.Function):
    """3x3 s1 p1 convolution of a few-channel image (Cin*9 <= 64, e.g. RGB -> 128): the 3x3 window
    is gathered to 64 columns and contracted as one GEMM tap (EncoderDual.py:41)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        col = kn.im2col3x3_small(x)
        dims, strs = kn.nhwc_view(col)
        y = torch.empty(nb, h, w, cout, dtype=BF16, device=x.device)
        wp = _packed(weight, "col_fwd")
        kn.tapgemm(col, dims, strs, wp, cout, 64, [(0, 0, 0, 0, 0)], 1, y, 0,
                   (h * w * cout, w * cout, cout), w, h, nb, cout, bias=_f32(bias))
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        dw = db = None
        if ctx.needs_input_grad[1]:
            col = kn.im2col3x3_small(x)
            dwc = _col_wgrad(dy, col, cout)                        # [cout, 64]
            dw = dwc[:, :9 * cin].reshape(cout, 3, 3, cin).permute(0, 3, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = kn.bias_grad(dy)
        return None, dw, db                                        # no gradient to the image


def _col_wgrad(a_nhwc, col, m_channels):
    """out[m, j] = sum_pixels a[pixel, m] * col[pixel, j]  (fp32 [m_channels, 64])."""
    nb, h, w, _ = a_nhwc.shape
    adims, astrs = kn.nhwc_view(a_nhwc)
    bdims, bstrs = kn.nhwc_view(col)
    kw, kh, kq = kn.tile_shape(w, h, nb, pixels=64)
    ktw, kth = (w + kw - 1) // kw, (h + kh - 1) // kh
    kblocks = ktw * kth * ((nb + kq - 1) // kq)
    splits = max(1, min(148, kblocks // 4))
    partial = torch.empty(splits, 1, m_channels, 64, dtype=torch.float32, device=col.device)
    kn.mmgemm(a_nhwc, adims, astrs, True, col, bdims, bstrs, True, m_channels, 64, kblocks, partial,
              (m_channels * 64, m_channels * 64, 64), kbox=(kw, kh, kq), ktiles=(ktw, kth), splits=splits,
              out_f32=True, block_n=128)
    return partial.sum(0)[0]


class ConvOutFn(torch.autograd.Function):
    """3x3 s1 p1 convolution to a few channels (128 -> RGB, DecoderPositional.py:91), fp32 NHWC
    output.  Backward gathers the 3x3 window of dy once and reuses it for dX and dW."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        dims, strs = kn.nhwc_view(x)
        taps = [(0, s - 1, 0, r - 1, (r * 3 + s) * cin) for r, s in kn.TAPS_3x3]
        y = torch.empty(nb, h, w, cout, dtype=torch.float32, device=x.device)
        wp = _packed(weight, "fwd_pad16")
        kn.tapgemm(x, dims, strs, wp, 16, wp.shape[1], taps, cin // 64, y, 0,
                   (h * w * cout, w * cout, cout), w, h, nb, cout, bias=_f32(bias), out_f32=True, block_n=16)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        dyb = dy.to(BF16).contiguous()
        colf = kn.im2col3x3_small(dyb, flip=True)                  # [N,H,W,64], col = t*cout + co
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(nb, h, w, cin, dtype=BF16, device=x.device)
            dims, strs = kn.nhwc_view(colf)
            kn.tapgemm(colf, dims, strs, _packed(weight, "col_dgrad"), cin, 64, [(0, 0, 0, 0, 0)], 1, dx, 0,
                       (h * w * cin, w * cin, cin), w, h, nb, cin)
        if ctx.needs_input_grad[1]:
            dwc = _col_wgrad(x, colf, cin)                          # [cin, 64]: [ci, t*cout + co]
            dw = dwc[:, :9 * cout].reshape(cin, 3, 3, cout).permute(3, 0, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum((0, 1, 2))
        return dx, dw, db


class GroupNormSwishFn(torch.autograd.Function):
    """GroupNorm(32, eps=1e-6, affine) optionally followed by swish (model.py:29-35)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, swish, stats=None):
        g, b = _f32(gamma), _f32(beta)
        if stats is not None:
            y = kn.gn_apply(x, stats, g, b, swish)
        else:
            y, stats = kn.gn_forward(x, g, b, swish)
        ctx.save_for_backward(x, stats, g, b)
        ctx.swish = swish
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, g, b = ctx.saved_tensors
        dx, dg, db = kn.gn_bwd(dy.contiguous(), x, stats, g, b, ctx.swish)
        return dx, dg, db, None, None


class ResnetBlockFn(torch.autograd.Function):
    """x + conv2(swish(GN(conv1(swish(GN(x)))))) with an optional 1x1 / 3x3 shortcut convolution
    (model.py:117-137) as ONE autograd node, so the backward can order and fuse its kernels:
    the gradient arriving through the residual branch is added inside the last GroupNorm-backward
    pass (no separate accumulation kernel), GroupNorm statistics come from the producing conv's
    epilogue when available."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, c1w, c1b, n2w, n2b, c2w, c2b, scw, scb, in_stats):
        g1, b1, g2, b2 = _f32(n1w), _f32(n1b), _f32(n2w), _f32(n2b)
        cout = c1w.shape[0]
        if in_stats is not None:
            st1 = in_stats
            a1 = kn.gn_apply(x, st1, g1, b1, True)
        else:
            a1, st1 = kn.gn_forward(x, g1, b1, True)
        h1 = kn.conv_fwd(a1, _packed(c1w, "fwd"), _f32(c1b), 3, 1, cout)
        st2 = kn.last_conv_stats
        if st2 is not None:
            a2 = kn.gn_apply(h1, st2, g2, b2, True)
        else:
            a2, st2 = kn.gn_forward(h1, g2, b2, True)
        sc = x if scw is None else kn.conv_fwd(x, _packed(scw, "fwd"), _f32(scb), scw.shape[-1], 1, cout)
        out = kn.conv_fwd(a2, _packed(c2w, "fwd"), _f32(c2b), 3, 1, cout, residual=sc)
        ctx.save_for_backward(x, a1, h1, a2, st1, st2, g1, b1, g2, b2, c1w, c2w, scw)
        ctx.has_sc = scw is not None
        ctx.bias_flags = (c1b is not None, c2b is not None, scb is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, a1, h1, a2, st1, st2, g1, b1, g2, b2, c1w, c2w, scw = ctx.saved_tensors
        dout = dout.contiguous()
        cin, cout = c1w.shape[1], c1w.shape[0]
        hw = x.shape[1:3]
        hb1, hb2, hbs = ctx.bias_flags
        d_a2 = kn.conv_dgrad(dout, _packed(c2w, "dgrad"), 3, 1, cout, hw)
        dw2, db2 = kn.conv_wgrad(a2, dout, 3, 1, want_bias=True)
        d_h1, dg2, dbt2 = kn.gn_bwd(d_a2, h1, st2, g2, b2, True)
        d_a1 = kn.conv_dgrad(d_h1, _packed(c1w, "dgrad"), 3, 1, cin, hw)
        dw1, db1 = kn.conv_wgrad(a1, d_h1, 3, 1, want_bias=True)
        dws = dbs = None
        if ctx.has_sc:
            k = scw.shape[-1]
            res_grad = kn.conv_dgrad(dout, _packed(scw, "dgrad"), k, 1, cin, hw)
            dws, dbs = kn.conv_wgrad(x, dout, k, 1, want_bias=True)
        else:
            res_grad = dout
        dx, dg1, dbt1 = kn.gn_bwd(d_a1, x, st1, g1, b1, True, add=res_grad)
        return (dx, dg1, dbt1, dw1, db1 if hb1 else None, dg2, dbt2, dw2, db2 if hb2 else None,
                dws, dbs if hbs else None, None)


def resnet_block(x, blk):
    """blk: nn.Module with norm1/conv1/norm2/conv2 (+ nin_shortcut / conv_shortcut)."""
    global _pending_stats
    in_stats = None
    if _pending_stats is not None and _pending_stats[0]() is x:
        in_stats = _pending_stats[1]
    _pending_stats = None
    sc = None
    if blk.in_channels != blk.out_channels:
        sc = blk.conv_shortcut if blk.use_conv_shortcut else blk.nin_shortcut
    y = ResnetBlockFn.apply(x, blk.norm1.weight, blk.norm1.bias, blk.conv1.weight, blk.conv1.bias,
                            blk.norm2.weight, blk.norm2.bias, blk.conv2.weight, blk.conv2.bias,
                            None if sc is None else sc.weight, None if sc is None else sc.bias, in_stats)
    stats = kn.last_conv_stats
    kn.last_conv_stats = None
    _pending_stats = (weakref.ref(y), stats) if stats is not None else None
    return y


class AttentionFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(C)) v over the T = h*w positions of each image, single head of width C
    (model.py:176-188).  q, k, v: [B, T, C] bf16 (any row stride that is a multiple of 8)."""

    @staticmethod
    def forward(ctx, q, k, v):
        bsz, t, c = q.shape
        scale = float(int(c) ** -0.5)
        s = torch.empty(bsz, t, t, dtype=torch.float32, device=q.device)
        _mm(q, False, k, False, t, t, c, s, alpha=scale, out_f32=True)
        p = kn.softmax_rows(s, t)
        o = torch.empty(bsz, t, c, dtype=BF16, device=q.device)
        _mm(p, False, v, True, t, c, t, o)
        ctx.save_for_backward(q, k, v, p)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, p = ctx.saved_tensors
        do = do.contiguous()
        bsz, t, c = q.shape
        dv = torch.empty(bsz, t, c, dtype=BF16, device=q.device)
        _mm(p, True, do, True, t, c, t, dv)                       # P^T dO
        dp = torch.empty(bsz, t, t, dtype=BF16, device=q.device)
        _mm(do, False, v, False, t, t, c, dp)                     # dO V^T
        ds = kn.softmax_bwd_rows(p, dp, t, ctx.scale)
        dq = torch.empty(bsz, t, c, dtype=BF16, device=q.device)
        _mm(ds, False, k, True, t, c, t, dq)                      # dS K
        dk = torch.empty(bsz, t, c, dtype=BF16, device=q.device)
        _mm(ds, True, q, True, t, c, t, dk)                       # dS^T Q
        return dq, dk, dv


def _mm(a, a_mn, b, b_mn, M, N, K, out, alpha=1.0, out_f32=False):
    """Batched out[z] = alpha * A B^T with A:[M,K] (K-major) or stored [K,M] (a_mn), same for B.
    out [bsz, M, N] may be a column slice of a wider buffer (unit stride along N)."""
    bsz = a.shape[0]
    assert a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1

    def view(t, mn, rows):
        sb, sr = t.stride(0), t.stride(1)
        if mn:   # stored [K][rows]
            return (rows, K, 1, 1, bsz), (1, sr, sb, sb, sb)
        return (K, rows, 1, 1, bsz), (1, sr, sb, sb, sb)

    ad, as_ = view(a, a_mn, M)
    bd, bs = view(b, b_mn, N)
    kb = (K + 63) // 64                       # a ragged last block reads zeros (TMA out-of-bounds fill)
    kn.mmgemm(a, ad, as_, a_mn, b, bd, bs, b_mn, M, N, kb, out, (out.stride(0), 0, out.stride(1)),
              kbox=(64, 1, 1), ktiles=(kb, 1), batches=bsz, alpha=alpha, out_f32=out_f32)


def _packed_cat(weights, kind):
    """bf16 packing of several 1x1 OIHW weights stacked along the output channels: "fwd" -> [sum Cout, Cin],
    "dgrad" -> [Cin, sum Cout].  Cached on the first parameter, validated against every member's version."""
    head = weights[0]
    cache = head.__dict__.setdefault("_b2_packs", {})
    stamp = tuple((w.data_ptr(), w._version, w.device) for w in weights)
    ent = cache.get("cat_" + kind)
    if ent is not None and ent[0] == stamp:
        return ent[1]
    if head.is_cuda and head.dtype == torch.float32 and all(w.shape[2:] == (1, 1) for w in weights):
        # one concatenation + ONE packing launch give both layouts of the stacked weight: [sum Cout, Cin] and its
        # transpose [Cin, sum Cout] (= the per-weight data-gradient packings side by side)
        fwd, dgr = kn.pack_weights(torch.cat([w.detach() for w in weights], dim=0))
        cache["cat_fwd"], cache["cat_dgrad"] = (stamp, fwd), (stamp, dgr)
        return fwd if kind == "fwd" else dgr
    parts = [kn.pack_weight_fwd(w) if kind == "fwd" else kn.pack_weight_dgrad(w) for w in weights]
    p = torch.cat(parts, dim=0 if kind == "fwd" else 1).contiguous()
    cache["cat_" + kind] = (stamp, p)
    return p


class AttnQKVFn(torch.autograd.Function):
    """q, k, v = three 1x1 convolutions of the same normalised input, then softmax(q k^T / sqrt(C)) v
    (model.py:170-188), as ONE autograd node: one [C, 3C] GEMM produces q | k | v side by side, the attention
    GEMMs read them as column slices, the backward writes dq | dk | dv into one buffer that feeds ONE data-gradient
    GEMM and ONE weight-gradient GEMM (no three-way gradient accumulation, a third of the launches)."""

    @staticmethod
    def forward(ctx, hn, qw, qb, kw, kb, vw, vb):
        bsz, h, w, c = hn.shape
        t = h * w
        bias = None if qb is None else torch.cat([_f32(qb), _f32(kb), _f32(vb)])
        qkv = kn.conv_fwd(hn, _packed_cat((qw, kw, vw), "fwd"), bias, 1, 1, 3 * c).view(bsz, t, 3 * c)
        q, k, v = qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]
        scale = float(int(c) ** -0.5)
        s = torch.empty(bsz, t, t, dtype=torch.float32, device=hn.device)
        _mm(q, False, k, False, t, t, c, s, alpha=scale, out_f32=True)
        p = kn.softmax_rows(s, t)
        o = torch.empty(bsz, t, c, dtype=BF16, device=hn.device)
        _mm(p, False, v, True, t, c, t, o)
        ctx.save_for_backward(hn, qkv, p, qw, kw, vw)
        ctx.scale, ctx.has_bias = scale, qb is not None
        return o

    @staticmethod
    def backward(ctx, do):
        hn, qkv, p, qw, kw, vw = ctx.saved_tensors
        do = do.contiguous()
        bsz, h, w, c = hn.shape
        t = h * w
        q, k, v = qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]
        dqkv = torch.empty_like(qkv)
        _mm(p, True, do, True, t, c, t, dqkv[..., 2 * c:])               # dV = P^T dO
        dp = torch.empty(bsz, t, t, dtype=BF16, device=hn.device)
        _mm(do, False, v, False, t, t, c, dp)                           # dP = dO V^T
        ds = kn.softmax_bwd_rows(p, dp, t, ctx.scale)
        _mm(ds, False, k, True, t, c, t, dqkv[..., :c])                 # dQ = dS K
        _mm(ds, True, q, True, t, c, t, dqkv[..., c:2 * c])             # dK = dS^T Q
        g = dqkv.view(bsz, h, w, 3 * c)
        d_hn = None
        if ctx.needs_input_grad[0]:
            d_hn = kn.conv_dgrad(g, _packed_cat((qw, kw, vw), "dgrad"), 1, 1, c, (h, w))
        dws = [None] * 3
        dbs = [None] * 3
        if any(ctx.needs_input_grad[i] for i in (1, 3, 5)):
            if ctx.has_bias:
                dw, db = kn.conv_wgrad(hn, g, 1, 1, want_bias=True)
                dbs = list(db.split(c))
            else:
                dw = kn.conv_wgrad(hn, g, 1, 1)
            dws = list(dw.split(c, dim=0))
        return d_hn, dws[0], dbs[0], dws[1], dbs[1], dws[2], dbs[2]


class Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return kn.upsample2x(x)

    @staticmethod
    def backward(ctx, g):
        return kn.upsample2x_bwd(g.contiguous())


class UpsampleConvFn(torch.autograd.Function):
    """conv3x3(nearest_upsample_x2(x)) + bias (Upsample.forward, model.py:49-53) without materialising the
    upsampled tensor: four 2x2 parity-class convolutions of the low-resolution input (kernels.upconv_*)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = kn.upconv_fwd(x, _packed(weight, "up_fwd"), _f32(bias), weight.shape[0])
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = kn.upconv_dgrad(dy, _packed(weight, "up_dgrad"), weight.shape[1])
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            if want_db:
                dw, db = kn.upconv_wgrad(x, dy, want_bias=True)
            else:
                dw = kn.upconv_wgrad(x, dy)
        elif want_db:
            db = kn.bias_grad(dy)
        return dx, dw, db


class ToNHWCFn(torch.autograd.Function):
    """NCHW fp32 -> NHWC bf16 (module boundary)."""

    @staticmethod
    def forward(ctx, x):
        return kn.nchw_f32_to_nhwc_bf16(x.float().contiguous())

    @staticmethod
    def backward(ctx, g):
        return kn.nhwc_bf16_to_nchw_f32(g.contiguous())


class ToNCHWFn(torch.autograd.Function):
    """NHWC (bf16 or fp32) -> NCHW fp32 (module boundary)."""

    @staticmethod
    def forward(ctx, x):
        ctx.was_bf16 = x.dtype == BF16
        if ctx.was_bf16:
            return kn.nhwc_bf16_to_nchw_f32(x.contiguous())
        return kn.nhwc_f32_to_nchw_f32(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        g = g.float().contiguous()
        if ctx.was_bf16:
            return kn.nchw_f32_to_nhwc_bf16(g)
        return kn.nchw_f32_to_nhwc_f32(g)


class ToNHWC32Fn(torch.autograd.Function):
    """NCHW fp32 -> NHWC fp32 (kept in fp32: the reference-facing VQ entry point)."""

    @staticmethod
    def forward(ctx, x):
        return kn.nchw_f32_to_nhwc_f32(x.float().contiguous())

    @staticmethod
    def backward(ctx, g):
        return kn.nhwc_f32_to_nchw_f32(g.float().contiguous())


ToNCHWInvFn = ToNHWC32Fn


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return kn.add_bf16(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, g


# convenience wrappers -------------------------------------------------------------------------
def conv2d(x, conv, residual=None, stride=None):
    """x NHWC bf16; conv: an nn.Conv2d used as a parameter container."""
    global _pending_stats
    k = conv.kernel_size[0]
    st = conv.stride[0] if stride is None else stride
    y = Conv2dFn.apply(x, conv.weight, conv.bias, residual, k, st)
    stats = kn.last_conv_stats                      # set by the forward that just ran (or None)
    kn.last_conv_stats = None
    _pending_stats = (weakref.ref(y), stats) if stats is not None else None
    return y


def gn_swish(x, norm, swish=True):
    global _pending_stats
    stats = None
    if _pending_stats is not None and _pending_stats[0]() is x and norm.num_groups == 32:
        stats = _pending_stats[1]
    _pending_stats = None
    return GroupNormSwishFn.apply(x, norm.weight, norm.bias, swish, stats)


def to_nhwc(x):
    return ToNHWCFn.apply(x)


def to_nchw(x):
    return ToNCHWFn.apply(x)
