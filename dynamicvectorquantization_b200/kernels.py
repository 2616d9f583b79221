"""Thin, autograd-free launchers for the C-ABI kernels, operating on torch CUDA tensors.

Activations are NHWC bf16 ([N,H,W,C] contiguous).  Every function runs on the current CUDA
stream and raises if the extension is missing or a launch fails - there is no fallback path.
"""
import ctypes as C

import torch

from . import _cabi
from ._cabi import MmDesc, TapGemmDesc
from ._cabi import check as _check

BF16 = torch.bfloat16

# kernels launched per C-ABI call (for bench.py's gpu_launches); everything else launches one
_MULTI = {"vq_ema_finalize": 2, "gn_stats": 2, "gn_bwd_stats": 2, "gn_bwd_apply": 2, "bias_grad": 2}   # others: 1
_launches = 0


def check(status, what):
    global _launches
    _check(status, what)
    _launches += _MULTI.get(what, 1)


def launch_count():
    """Number of kernels of libb200dq.so launched by this process so far."""
    return _launches


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _pow2ceil(v):
    p = 1
    while p < v:
        p *= 2
    return p


def tile_shape(wout, hout, nb, pixels=128):
    tw = min(pixels, _pow2ceil(wout))
    th = min(pixels // tw, _pow2ceil(hout))
    tn = pixels // (tw * th)
    return tw, th, tn


# ------------------------------------------------------------------------------------------ VQ
class Codebook:
    """bf16 copy + squared norms of the fp32 codebook (search operands)."""

    def __init__(self, K, Cdim, device):
        self.K, self.C = K, Cdim
        kpad = (K + 255) // 256 * 256
        self.cb = torch.empty(K, Cdim, dtype=BF16, device=device)
        self.sqnorm = torch.empty(kpad, dtype=torch.float32, device=device)

    def refresh(self, weight_f32):
        """weight_f32: [K(+1), C] fp32 contiguous; rows [:K] are used."""
        check(_cabi.lib().b2dq_vq_prepare_codebook(_ptr(weight_f32), _ptr(self.cb), _ptr(self.sqnorm),
                                                   self.K, self.C, _stream()), "vq_prepare_codebook")


def vq_search_gather(x_bf16, codebook, weight_f32, x_f32=None, row_mask=None, want_xq_bf16=True,
                     want_xq_f32=False, counts=None, sums=None, loss_acc=None, max_ctas=0, split=True):
    """x_bf16 [N,C].  Returns (codes int64 [N], xq_bf16 | None, xq_f32 | None).
    split=True lets small-N calls spread the codebook over several CTAs per row tile (same results)."""
    N, Cd = x_bf16.shape
    dev = x_bf16.device
    ws, ws_bytes = None, 0
    if split and N > 0:
        ws_bytes = _cabi.lib().b2dq_vq_search_workspace_bytes(N, codebook.K)
        if ws_bytes > 0:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    codes = torch.empty(N, dtype=torch.int64, device=dev)
    xq_b = torch.empty(N, Cd, dtype=BF16, device=dev) if want_xq_bf16 else None
    xq_f = torch.empty(N, Cd, dtype=torch.float32, device=dev) if want_xq_f32 else None
    check(_cabi.lib().b2dq_vq_search_gather(
        _ptr(x_bf16), _ptr(x_f32), _ptr(codebook.cb), _ptr(codebook.sqnorm), _ptr(weight_f32),
        _ptr(row_mask), _ptr(codes), _ptr(xq_b), _ptr(xq_f), _ptr(loss_acc), _ptr(counts), _ptr(sums),
        N, Cd, codebook.K, max_ctas, _ptr(ws), ws_bytes, _stream()), "vq_search_gather")
    return codes, xq_b, xq_f


def vq_ema_finalize(counts, sums, restart_rows, cluster_size_ema, embed_ema, weight_f32, decay, eps,
                    restart):
    K, Cd = embed_ema.shape
    dead = torch.empty(K, dtype=torch.uint8, device=embed_ema.device)
    nscr = torch.empty(1, dtype=torch.float32, device=embed_ema.device)
    check(_cabi.lib().b2dq_vq_ema_finalize(_ptr(counts), _ptr(sums), _ptr(restart_rows),
                                           _ptr(cluster_size_ema), _ptr(embed_ema), _ptr(weight_f32),
                                           _ptr(dead), _ptr(nscr), K, Cd, float(decay), float(eps),
                                           int(bool(restart)), _stream()), "vq_ema_finalize")
    return dead


def vq_bwd(g_xq, x, xq, row_mask, g_loss, coef):
    g_x = torch.empty_like(x)
    n_rows, Cd = x.shape
    check(_cabi.lib().b2dq_vq_bwd(_ptr(g_xq), _ptr(x), _ptr(xq), _ptr(row_mask), _ptr(g_loss),
                                  float(coef), _ptr(g_x), n_rows, Cd, _stream()), "vq_bwd")
    return g_x


# ------------------------------------------------------------------------------------------ conv
def _fill(arr, vals):
    for i, v in enumerate(vals):
        arr[i] = v


def tapgemm(a, a_dims, a_strides, b, b_rows, b_k, taps, kchunks, out, out_off, ostr, wout, hout, nb,
            cout, bias=None, residual=None, res_off=0, rstr=(0, 0, 0), alpha=1.0, out_f32=False,
            block_n=0, b_batch=1, b_batch_stride=0, m_tiles_per_cta=0, relu=False):
    """taps: list of (dc, dw, dp, dh, bk)."""
    d = TapGemmDesc()
    d.a_ptr = a.data_ptr()
    _fill(d.a_dims, a_dims)
    _fill(d.a_strides, a_strides)
    d.b_ptr = b.data_ptr()
    d.b_rows, d.b_k, d.b_batch, d.b_batch_stride = b_rows, b_k, b_batch, b_batch_stride
    d.num_taps, d.kchunks = len(taps), kchunks
    for i, (tc, tw, tp, th, bk) in enumerate(taps):
        d.tap_c[i], d.tap_w[i], d.tap_p[i], d.tap_h[i], d.tap_bk[i] = tc, tw, tp, th, bk
    d.TW, d.TH, d.TN = tile_shape(wout, hout, nb)
    d.Wout, d.Hout, d.NB, d.Cout = wout, hout, nb, cout
    esz = 4 if out_f32 else 2
    d.out = out.data_ptr() + out_off * esz
    d.oN, d.oH, d.oW = ostr
    d.bias = _ptr(bias)
    d.residual = None if residual is None else residual.data_ptr() + res_off * 2
    d.rN, d.rH, d.rW = rstr
    if block_n == 0:
        m_tiles = -(-wout // d.TW) * -(-hout // d.TH) * -(-nb // d.TN)
        block_n = _pick_block_n(m_tiles, cout)
    d.alpha, d.out_f32, d.block_n, d.m_tiles_per_cta = alpha, int(out_f32), block_n, (m_tiles_per_cta or FORCE_MT)
    d.relu = int(relu)                       # 0 none, 1 ReLU, 2 LeakyReLU(0.2)
    check(_cabi.lib().b2dq_tapgemm(C.byref(d), _stream()), "tapgemm")


def nhwc_view(x):
    """5-D TMA view (c, w, p, h, n) of a contiguous NHWC tensor: dims, strides (elements)."""
    nb, h, w, c = x.shape
    return (c, w, 1, h, nb), (1, c, w * c, w * c, h * w * c)


def parity_view(x):
    """Stride-2 view [N, H/2, 2, W/2, 2C] as (c2, w2, p, h2, n)."""
    nb, h, w, c = x.shape
    return (2 * c, w // 2, 2, h // 2, nb), (1, 2 * c, w * c, 2 * w * c, h * w * c)


TAPS_3x3 = [(r, s) for r in range(3) for s in range(3)]


def pack_weight_fwd(w):
    """OIHW fp32 -> [Cout, R*S*Cin] bf16 (tap-major, Cin contiguous)."""
    co, ci, r, s = w.shape
    return w.detach().permute(0, 2, 3, 1).reshape(co, r * s * ci).to(BF16).contiguous()


def pack_weight_dgrad(w):
    """OIHW fp32 -> [Cin, R*S*Cout] bf16 (for the data gradient: contraction over Cout)."""
    co, ci, r, s = w.shape
    return w.detach().permute(1, 2, 3, 0).reshape(ci, r * s * co).to(BF16).contiguous()


def pack_weights(w, want_fwd=True, want_dgrad=True):
    """Both bf16 packings of an OIHW fp32 weight in ONE launch -> (fwd | None, dgrad | None)."""
    co, ci, r, s = w.shape
    w = w.detach().contiguous()
    fwd = torch.empty(co, r * s * ci, dtype=BF16, device=w.device) if want_fwd else None
    dgr = torch.empty(ci, r * s * co, dtype=BF16, device=w.device) if want_dgrad else None
    check(_cabi.lib().b2dq_pack_weights(_ptr(w), _ptr(fwd), _ptr(dgr), co, ci, r, s, _stream()), "pack_weights")
    return fwd, dgr


def pack_weights_multi(table, n_items, total_tiles, max_rs):
    """table: device int64 [n_items, 6] records (see b2dq_pack_weights_multi); repacks all of them in one launch."""
    check(_cabi.lib().b2dq_pack_weights_multi(_ptr(table), int(n_items), int(total_tiles), int(max_rs), _stream()),
          "pack_weights_multi")


def conv_fwd(x, wpack, bias, ksize, stride, cout, residual=None, out_f32=False, relu=False):
    """x NHWC bf16; wpack from pack_weight_fwd; stride-2 uses pad (0,1,0,1) like Downsample.
    relu: clamp the output at zero in the epilogue (tap-GEMM path only)."""
    global last_conv_stats
    last_conv_stats = None
    nb, h, w, cin = x.shape
    assert cin % 64 == 0, "Cin must be a multiple of 64 (edge layers use the im2col path)"
    if not out_f32 and not relu and _pconv_ok(ksize, stride, w, cin, cout, nb, h):
        return pconv3x3(x, wpack, bias, residual, dgrad=False, want_stats=FUSE_GN_STATS)
    kch = cin // 64
    if stride == 1:
        dims, strs = nhwc_view(x)
        if ksize == 3:
            taps = [(0, s - 1, 0, r - 1, (r * 3 + s) * cin) for r, s in TAPS_3x3]
        else:
            taps = [(0, 0, 0, 0, 0)]
        ho, wo = h, w
    else:
        assert ksize == 3 and h % 2 == 0 and w % 2 == 0
        dims, strs = parity_view(x)
        taps = [((s % 2) * cin, s // 2, r % 2, r // 2, (r * 3 + s) * cin) for r, s in TAPS_3x3]
        ho, wo = h // 2, w // 2
    out = torch.empty(nb, ho, wo, cout, dtype=torch.float32 if out_f32 else BF16, device=x.device)
    ostr = (ho * wo * cout, wo * cout, cout)
    tapgemm(x, dims, strs, wpack, wpack.shape[0], wpack.shape[1], taps, kch, out, 0, ostr, wo, ho, nb,
            cout, bias=bias, residual=residual, rstr=ostr, out_f32=out_f32, relu=relu)
    return out


def conv_dgrad(dy, wdpack, ksize, stride, cin, in_hw):
    """dy NHWC bf16 [N,Ho,Wo,Cout] -> dx NHWC bf16 [N,H,W,Cin]."""
    nb, ho, wo, cout = dy.shape
    assert cout % 64 == 0
    if _pconv_ok(ksize, stride, wo, cout, cin, nb, ho):
        return pconv3x3(dy, wdpack, None, None, dgrad=True)
    kch = cout // 64
    h, w = in_hw
    dx = torch.empty(nb, h, w, cin, dtype=BF16, device=dy.device)
    dims, strs = nhwc_view(dy)
    if stride == 1:
        if ksize == 3:
            taps = [(0, 1 - s, 0, 1 - r, (r * 3 + s) * cout) for r, s in TAPS_3x3]
        else:
            taps = [(0, 0, 0, 0, 0)]
        tapgemm(dy, dims, strs, wdpack, wdpack.shape[0], wdpack.shape[1], taps, kch, dx, 0,
                (h * w * cin, w * cin, cin), w, h, nb, cin)
    else:
        # y[oh,ow] = sum W[r,s] x[2oh+r, 2ow+s]  =>  dx[2i+ph, 2j+pw] gathers the taps with
        # r = ph (mod 2): (r, dh) in {(0,0),(2,-1)} for ph=0 and {(1,0)} for ph=1 (same for columns).
        rsel = {0: [(0, 0), (2, -1)], 1: [(1, 0)]}
        strips = _pconv_taps_ok(wo, cout, cin, nb, ho)
        for ph in (0, 1):
            for pw in (0, 1):
                if strips:                   # persistent strip kernel: one strip serves the class's column taps
                    pconv_taps(dy, dims, strs, wdpack, kch, [(0, 0, dh) for _, dh in rsel[ph]],
                               [dw for _, dw in rsel[pw]],
                               [[(r * 3 + s) * cout for s, _ in rsel[pw]] for r, _ in rsel[ph]], nb, ho, wo, dx,
                               (ph * w + pw) * cin, (h * w * cin, 2 * w * cin, 2 * cin))
                    continue
                taps = [(0, dw, 0, dh, (r * 3 + s) * cout) for r, dh in rsel[ph] for s, dw in rsel[pw]]
                tapgemm(dy, dims, strs, wdpack, wdpack.shape[0], wdpack.shape[1], taps, kch, dx,
                        (ph * w + pw) * cin, (h * w * cin, 2 * w * cin, 2 * cin), wo, ho, nb, cin)
    return dx


def mmgemm(a, a_dims, a_strides, a_mn, b, b_dims, b_strides, b_mn, M, N, kblocks, out, ostr,
           taps=((0, 0, 0, 0),), kbox=(64, 1, 1), ktiles=(0, 0), splits=1, batches=1, alpha=1.0,
           out_f32=False, block_n=0, out_off=0, taps_per_cta=0, b_strip=False, colsum=None):
    d = MmDesc()
    d.a_ptr, d.b_ptr = a.data_ptr(), b.data_ptr()
    _fill(d.a_dims, a_dims); _fill(d.a_strides, a_strides)
    _fill(d.b_dims, b_dims); _fill(d.b_strides, b_strides)
    d.a_mn, d.b_mn, d.ntaps, d.taps_per_cta = int(a_mn), int(b_mn), len(taps), taps_per_cta
    for i, (tc, tw, tp, th) in enumerate(taps):
        d.tap_c[i], d.tap_w[i], d.tap_p[i], d.tap_h[i] = tc, tw, tp, th
    d.KW, d.KH, d.KN = kbox
    d.ktiles_w, d.ktiles_h, d.kblocks = ktiles[0], ktiles[1], kblocks
    d.splits, d.batches, d.M, d.N = splits, batches, M, N
    esz = 4 if out_f32 else 2
    d.out = out.data_ptr() + out_off * esz
    d.oZ, d.oT, d.oM = ostr
    d.alpha, d.out_f32, d.block_n, d.b_strip = alpha, int(out_f32), block_n, int(b_strip)
    d.colsum = _ptr(colsum)
    check(_cabi.lib().b2dq_mmgemm(C.byref(d), _stream()), "mmgemm")


NUM_SMS = 148
FORCE_MT = 0          # tests / tuning: force m_tiles_per_cta of the tap GEMM (0 = library heuristic)
FUSE_BIAS_GRAD = True    # 3x3 convs: bias gradient from the weight-gradient GEMM (dY x ones on the tensor core)
USE_WGRAD_STRIP = True   # 3x3 s1 weight gradient: one 66-pixel activation strip per k-block for a filter row
USE_PCONV = True      # persistent strip kernel for 3x3 s1 layers with 128 output channels and W % 128 == 0


def _pconv_ok(ksize, stride, w, cin, cout, nb, h):
    return (USE_PCONV and ksize == 3 and stride == 1 and cout == 128 and w % 128 == 0 and cin % 64 == 0
            and nb * h * (w // 128) >= 2 * NUM_SMS)


FUSE_GN_STATS = True     # pconv forward also emits the GroupNorm statistics of its output
last_conv_stats = None   # (mean, rstd) [N,32,2] of the most recent conv_fwd output, or None


def pconv3x3(x, wpack, bias, residual, dgrad, want_stats=False):
    """Persistent strip kernel (b2dq_pconv3x3): 3x3 stride-1 convolution (or its data gradient) to 128 channels."""
    global last_conv_stats
    nb, h, w, cin = x.shape
    lib = _cabi.lib()
    out = torch.empty(nb, h, w, 128, dtype=BF16, device=x.device)
    part = torch.empty(nb * h * (w // 128), 64, dtype=torch.float32, device=x.device) if want_stats else None
    check(lib.b2dq_pconv3x3(_ptr(x), _ptr(wpack), _ptr(out), _ptr(bias), _ptr(residual), _ptr(part), nb, h, w, cin,
                            int(dgrad), 0, _stream()), "pconv3x3")
    if want_stats:
        stats = torch.empty(nb, 32, 2, dtype=torch.float32, device=x.device)
        check(lib.b2dq_gn_finalize_tiles(_ptr(part), _ptr(stats), nb, h, w, 1e-6, _stream()), "gn_finalize_tiles")
        last_conv_stats = stats
    return out


import os as _os0
USE_PCONV_TAPS = _os0.environ.get("B2DQ_PCONV_TAPS", "1") != "0"  # parity classes (folded up-convolution, stride-2 data gradient) on the persistent strip kernel


def _pconv_taps_ok(w_tiles, cin, cout, nb, h):
    return (USE_PCONV and USE_PCONV_TAPS and cout == 128 and w_tiles % 128 == 0 and cin % 64 == 0
            and nb * h * (w_tiles // 128) >= 2 * NUM_SMS)


def pconv_taps(a, a_dims, a_strides, b, kchunks, rows, cols, wcol, nb, h, w, out, out_off, ostr, bias=None):
    """b2dq_pconv_taps: rows = [(c, p, dh)] row taps, cols = [dw] column taps, wcol[r][s] = weight column base."""
    d = _cabi.PconvTapsDesc()
    d.a_ptr = a.data_ptr()
    _fill(d.a_dims, a_dims)
    _fill(d.a_strides, a_strides)
    d.b_ptr, d.b_k, d.kchunks = b.data_ptr(), b.shape[1], kchunks
    d.nr, d.ns = len(rows), len(cols)
    for i, (rc, rp, rdh) in enumerate(rows):
        d.row_c[i], d.row_p[i], d.row_dh[i] = rc, rp, rdh
        for j in range(len(cols)):
            d.wcol[i * 3 + j] = wcol[i][j]
    _fill(d.col_dw, cols)
    d.NB, d.H, d.W = nb, h, w
    d.out = out.data_ptr() + out_off * 2
    d.oN, d.oH, d.oW = ostr
    d.bias = _ptr(bias)
    check(_cabi.lib().b2dq_pconv_taps(C.byref(d), 0, _stream()), "pconv_taps")


import os as _os
# k-blocks (of 64 pixels) a CTA must keep for the split-K weight gradient to be cut into TWO waves of CTAs: every
# split writes (and the reduction re-reads) a full fp32 copy of the weight gradient, so short CTAs pay more for their
# epilogue and partial traffic than the second wave gains in balance
# (measured on B200, tools/gpu_r2k.sh: ONE wave wins on every layer shape of the model - 0.511 -> 0.486 ms at
# 128->128 @256^2, 0.085 -> 0.064 ms at 256->256 @32^2, step 58.8 -> 57.3 ms - so the default never cuts two)
WGRAD_TWO_WAVE_MIN_KB = int(_os.environ.get("B2DQ_WGRAD_TWO_WAVE_MIN_KB", str(1 << 30)))


def _wgrad_splits(kblocks, ctas_per_split):
    """Split-K factor so that one launch (ctas_per_split output tiles x splits CTAs, 1 CTA/SM)
    fills the 148 SMs in whole waves, with at least 4 k-blocks per CTA."""
    target = max(1, NUM_SMS // max(1, ctas_per_split))
    if kblocks >= WGRAD_TWO_WAVE_MIN_KB * target * 2:   # plenty of work: two full waves balance better than one
        target *= 2
    return max(1, min(target, kblocks // 4 if kblocks >= 4 else 1))


def _pick_block_n(m_tiles, cout):
    """128 x BN output tiles: prefer BN=256 (A tile read once) unless that leaves the grid under
    two waves of 148 SMs, where BN=128 (2 CTAs/SM, epilogue overlap) balances better."""
    if cout % 256 == 0 and m_tiles * (cout // 256) >= 4 * NUM_SMS:
        return 256
    if cout <= 16:
        return 16
    if cout <= 64:
        return 64
    return 128


def conv_wgrad(x, dy, ksize, stride, want_bias=False):
    """dW (fp32, OIHW) for y = conv(x): x NHWC bf16 [N,H,W,Cin], dy NHWC bf16 [N,Ho,Wo,Cout].
    want_bias: also return db = sum_pixels dy (3x3 convs get it from the same GEMM as dY x ones)."""
    nb, h, w, cin = x.shape
    _, ho, wo, cout = dy.shape
    if stride == 1:
        bdims, bstrs = nhwc_view(x)
        if ksize == 3:
            taps = [(0, s - 1, 0, r - 1) for r, s in TAPS_3x3]
        else:
            taps = [(0, 0, 0, 0)]
    else:
        bdims, bstrs = parity_view(x)
        taps = [((s % 2) * cin, s // 2, r % 2, r // 2) for r, s in TAPS_3x3]
    adims, astrs = nhwc_view(dy)
    kw, kh, kn = tile_shape(wo, ho, nb, pixels=64)
    ktw, kth = (wo + kw - 1) // kw, (ho + kh - 1) // kh
    kblocks = ktw * kth * ((nb + kn - 1) // kn)
    ntaps = len(taps)
    mt, nt = (cout + 127) // 128, (cin + 127) // 128
    ngroups = (ntaps + 2) // 3                 # <= 3 taps (TMEM accumulators) per CTA, folded into the grid
    splits = _wgrad_splits(kblocks, mt * nt * ngroups)
    partial = torch.empty(splits, ntaps, cout, cin, dtype=torch.float32, device=x.device)
    fuse_bias = want_bias and ntaps >= 3 and FUSE_BIAS_GRAD
    colsum = torch.empty(splits, cout, dtype=torch.float32, device=x.device) if fuse_bias else None
    mmgemm(dy, adims, astrs, True, x, bdims, bstrs, True, cout, cin, kblocks, partial,
           (ntaps * cout * cin, cout * cin, cin), taps=taps, kbox=(kw, kh, kn), ktiles=(ktw, kth),
           splits=splits, out_f32=True, block_n=128, taps_per_cta=min(3, ntaps),
           b_strip=USE_WGRAD_STRIP and ksize == 3 and stride == 1 and (kw, kh, kn) == (64, 1, 1),
           colsum=colsum)
    dw = torch.empty(cout, cin, ksize, ksize, dtype=torch.float32, device=x.device)
    if fuse_bias:                              # weight and bias gradient finished by one reduction launch
        db = torch.empty(cout, dtype=torch.float32, device=x.device)
        check(_cabi.lib().b2dq_wgrad_reduce_bias(_ptr(partial), _ptr(dw), splits, ntaps, cout, cin, _ptr(colsum), _ptr(db),
                                                 _stream()), "wgrad_reduce_bias")
        return dw, db
    check(_cabi.lib().b2dq_wgrad_reduce(_ptr(partial), _ptr(dw), splits, ntaps, cout, cin, 0, _stream()),
          "wgrad_reduce")
    if not want_bias:
        return dw
    return dw, bias_grad(dy)


# ------------------------------------------------------------------------------------------ upsample + conv
# nearest-neighbour x2 followed by a 3x3 convolution (model.py:49-53) = four 2x2 convolutions of the LOW-resolution
# input, one per output parity class (ph, pw): output row 2i+ph reads low-res rows {i-1: r=0 | i: r=1,2} for ph = 0
# and {i: r=0,1 | i+1: r=2} for ph = 1 (same for columns), so the nine taps collapse onto four whose weights are
# sums of the original ones.  16 tap-GEMM steps per low-res pixel instead of 36, and the upsampled tensor is never
# written.  _UP_M[ph, a, r] = 1 when filter row r lands on low-res offset a of parity ph.
_UP_M = ((( 1, 0, 0), (0, 1, 1)), ((1, 1, 0), (0, 0, 1)))
_UP_OFF = ((-1, 0), (0, 1))                     # low-res offset of slot a for parity ph


_up_m_cache = {}


def _up_m(ref):
    """_UP_M as a tensor on ref's device (cached: no host-to-device copy inside a CUDA-graph capture)."""
    key = (ref.device, ref.dtype)
    if key not in _up_m_cache:
        _up_m_cache[key] = torch.tensor(_UP_M, dtype=ref.dtype, device=ref.device)
    return _up_m_cache[key]


def upconv_fold(w):
    """OIHW fp32 [Cout,Cin,3,3] -> folded weights [2(ph),2(pw),2(a),2(b),Cout,Cin] (fp32)."""
    m = _up_m(w)
    return torch.einsum("par,qbs,oirs->pqaboi", m, m, w)


def upconv_unfold_grad(dwf):
    """Gradient w.r.t. the folded weights [2,2,2,2,Cout,Cin] -> gradient w.r.t. the 3x3 weights (OIHW)."""
    m = _up_m(dwf)
    return torch.einsum("par,qbs,pqaboi->oirs", m, m, dwf).contiguous()


def upconv_pack(w):
    """-> (fwd [Cout, 16*Cin], dgrad [Cin, 16*Cout]) bf16; column block ((ph*2+pw)*2+a)*2+b."""
    co, ci = w.shape[0], w.shape[1]
    if w.is_cuda and w.dtype == torch.float32:          # fold + both packings in one launch
        w = w.detach().contiguous()
        fwd = torch.empty(co, 16 * ci, dtype=BF16, device=w.device)
        dgr = torch.empty(ci, 16 * co, dtype=BF16, device=w.device)
        check(_cabi.lib().b2dq_upconv_pack(_ptr(w), _ptr(fwd), _ptr(dgr), co, ci, _stream()), "upconv_pack")
        return fwd, dgr
    wf = upconv_fold(w.detach().float())
    fwd = wf.permute(4, 0, 1, 2, 3, 5).reshape(co, 16 * ci).to(BF16).contiguous()
    dgr = wf.permute(5, 0, 1, 2, 3, 4).reshape(ci, 16 * co).to(BF16).contiguous()
    return fwd, dgr


def upconv_fwd(x, wpack_fwd, bias, cout):
    """x NHWC bf16 [N,H,W,Cin] -> conv3x3(upsample2x(x)) [N,2H,2W,Cout] as four parity-class tap GEMMs."""
    nb, h, w, cin = x.shape
    assert cin % 64 == 0
    out = torch.empty(nb, 2 * h, 2 * w, cout, dtype=BF16, device=x.device)
    dims, strs = nhwc_view(x)
    strips = _pconv_taps_ok(w, cin, cout, nb, h)
    for ph in (0, 1):
        for pw in (0, 1):
            cls = ph * 2 + pw
            if strips:                       # persistent strip kernel: a 129-pixel strip serves both column slots
                pconv_taps(x, dims, strs, wpack_fwd, cin // 64, [(0, 0, _UP_OFF[ph][a]) for a in (0, 1)],
                           [_UP_OFF[pw][b] for b in (0, 1)],
                           [[((cls * 2 + a) * 2 + b) * cin for b in (0, 1)] for a in (0, 1)], nb, h, w, out,
                           (ph * 2 * w + pw) * cout, (4 * h * w * cout, 4 * w * cout, 2 * cout), bias=bias)
                continue
            taps = [(0, _UP_OFF[pw][b], 0, _UP_OFF[ph][a], ((cls * 2 + a) * 2 + b) * cin)
                    for a in (0, 1) for b in (0, 1)]
            tapgemm(x, dims, strs, wpack_fwd, cout, 16 * cin, taps, cin // 64, out, (ph * 2 * w + pw) * cout,
                    (4 * h * w * cout, 4 * w * cout, 2 * cout), w, h, nb, cout, bias=bias)
    return out


def upconv_dgrad(dy, wpack_dgrad, cin):
    """dy NHWC bf16 [N,2H,2W,Cout] -> gradient w.r.t. the low-resolution input [N,H,W,Cin]: one 16-tap GEMM
    over the parity view of dy."""
    nb, h2, w2, cout = dy.shape
    h, w = h2 // 2, w2 // 2
    assert cout % 64 == 0
    dx = torch.empty(nb, h, w, cin, dtype=BF16, device=dy.device)
    dims, strs = parity_view(dy)
    taps = []
    for ph in (0, 1):
        for pw in (0, 1):
            cls = ph * 2 + pw
            for a in (0, 1):
                for b in (0, 1):
                    # y[2i+ph, 2j+pw] reads x[i+dh, j+dw]  =>  dx[i, j] gathers dy[2(i-dh)+ph, 2(j-dw)+pw]
                    taps.append((pw * cout, -_UP_OFF[pw][b], ph, -_UP_OFF[ph][a], ((cls * 2 + a) * 2 + b) * cout))
    tapgemm(dy, dims, strs, wpack_dgrad, cin, 16 * cout, taps, cout // 64, dx, 0, (h * w * cin, w * cin, cin), w, h,
            nb, cin)
    return dx


UPCONV_WGRAD_TAPS_PER_CTA = int(_os.environ.get("B2DQ_UPCONV_WGRAD_TPC", "1"))   # slots of a class sharing one dY tile per CTA


def upconv_wgrad(x, dy, want_bias=False):
    """dW (fp32 OIHW [Cout,Cin,3,3]) [+ db] of conv3x3(upsample2x(x)): per parity class a 4-tap weight-gradient
    GEMM of the strided class view of dy against the low-resolution x, unfolded onto the nine filter taps."""
    nb, h, w, cin = x.shape
    cout = dy.shape[-1]
    bdims, bstrs = nhwc_view(x)
    kw, kh, kn = tile_shape(w, h, nb, pixels=64)
    ktw, kth = (w + kw - 1) // kw, (h + kh - 1) // kh
    kblocks = ktw * kth * ((nb + kn - 1) // kn)
    mt, nt = (cout + 127) // 128, (cin + 127) // 128
    splits = _wgrad_splits(kblocks, mt * nt * -(-4 // UPCONV_WGRAD_TAPS_PER_CTA))
    dy6 = dy.view(nb, h, 2, w, 2, cout)
    partial = torch.empty(4, splits, 4, cout, cin, dtype=torch.float32, device=x.device)
    for ph in (0, 1):
        for pw in (0, 1):
            dyc = dy6[:, :, ph, :, pw, :]                     # [N,H,W,Cout] view, pixel stride 2*Cout
            adims = (cout, w, 1, h, nb)
            astrs = (1, 2 * cout, 2 * w * cout, 4 * w * cout, 4 * h * w * cout)
            taps = [(0, _UP_OFF[pw][b], 0, _UP_OFF[ph][a]) for a in (0, 1) for b in (0, 1)]
            mmgemm(dyc, adims, astrs, True, x, bdims, bstrs, True, cout, cin, kblocks, partial[ph * 2 + pw],
                   (4 * cout * cin, cout * cin, cin), taps=taps, kbox=(kw, kh, kn), ktiles=(ktw, kth), splits=splits,
                   out_f32=True, block_n=128, taps_per_cta=UPCONV_WGRAD_TAPS_PER_CTA)
    # splits summed and the 16 class/slot gradients scattered back onto the nine filter taps in one launch
    dw = torch.empty(cout, cin, 3, 3, dtype=torch.float32, device=x.device)
    check(_cabi.lib().b2dq_upconv_wgrad_reduce(_ptr(partial), _ptr(dw), splits, cout, cin, _stream()),
          "upconv_wgrad_reduce")
    return (dw, bias_grad(dy)) if want_bias else dw


def bias_grad(dy):
    cch = dy.shape[-1]
    rows = dy.numel() // cch
    lib = _cabi.lib()
    out = torch.empty(cch, dtype=torch.float32, device=dy.device)
    part = torch.empty(lib.b2dq_bias_grad_blocks(rows) * cch, dtype=torch.float32, device=dy.device)
    check(lib.b2dq_bias_grad(_ptr(dy), _ptr(out), _ptr(part), rows, cch, _stream()), "bias_grad")
    return out


# ------------------------------------------------------------------------------------------ GN
# GroupNorm is two passes over the tensor (reduce, then apply).  Optionally the passes run in image
# groups small enough for the second pass to hit L2 (statistics are per image, so grouping is exact).
# Measured on B200 (tools/kernel_bench.py gn): at batch 32 / 256x256x128 the per-group launches cost far
# more than the saved HBM reads (bwd 4.5 ms grouped vs 0.65 ms whole), so grouping is OFF by default.
GN_L2_BUDGET_BYTES = 1 << 60


def _gn_groups(nb, per_image_bytes, tensors):
    g = max(1, GN_L2_BUDGET_BYTES // max(1, per_image_bytes * tensors))
    return nb if g >= nb else g


def gn_stats(x, groups=32, eps=1e-6):
    nb, h, w, c = x.shape
    lib = _cabi.lib()
    stats = torch.empty(nb, groups, 2, dtype=torch.float32, device=x.device)
    ws = torch.empty(nb * lib.b2dq_gn_chunks(nb, h * w) * groups * 2, dtype=torch.float32, device=x.device)
    check(lib.b2dq_gn_stats(_ptr(x), _ptr(stats), _ptr(ws), nb, h * w, c, groups, eps, _stream()), "gn_stats")
    return stats


def gn_apply(x, stats, gamma, beta, swish, groups=32):
    nb, h, w, c = x.shape
    y = torch.empty_like(x)
    check(_cabi.lib().b2dq_gn_apply(_ptr(x), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(y), nb, h * w, c,
                                    groups, int(swish), _stream()), "gn_apply")
    return y


def gn_forward(x, gamma, beta, swish, groups=32, eps=1e-6):
    """stats + apply.  Returns (y, stats).  One persistent kernel (statistics, team barrier per image, apply from L2)
    when the shape allows it, else the statistics kernels followed by the apply kernel."""
    nb, h, w, c = x.shape
    if USE_GN_FUSED and int(swish) in (0, 1):
        lib = _cabi.lib()
        ws_bytes = lib.b2dq_gn_fwd_fused_workspace_bytes(nb, h * w, c, groups)
        if ws_bytes > 0:
            y = torch.empty_like(x)
            stats = torch.empty(nb, groups, 2, dtype=torch.float32, device=x.device)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            check(lib.b2dq_gn_fwd_fused(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(stats), _ptr(ws), ws_bytes, nb,
                                        h * w, c, groups, float(eps), int(swish), _stream()), "gn_fwd_fused")
            return y, stats
    g = _gn_groups(nb, h * w * c * 2, 1)
    if g >= nb:
        stats = gn_stats(x, groups, eps)
        return gn_apply(x, stats, gamma, beta, swish, groups), stats
    lib = _cabi.lib()
    y = torch.empty_like(x)
    stats = torch.empty(nb, groups, 2, dtype=torch.float32, device=x.device)
    ws = torch.empty(g * lib.b2dq_gn_chunks(g, h * w) * groups * 2, dtype=torch.float32, device=x.device)
    for n0 in range(0, nb, g):
        n = min(g, nb - n0)
        xs, ys, ss = x[n0:n0 + n], y[n0:n0 + n], stats[n0:n0 + n]
        check(lib.b2dq_gn_stats(_ptr(xs), _ptr(ss), _ptr(ws), n, h * w, c, groups, eps, _stream()), "gn_stats")
        check(lib.b2dq_gn_apply(_ptr(xs), _ptr(ss), _ptr(gamma), _ptr(beta), _ptr(ys), n, h * w, c, groups,
                                int(swish), _stream()), "gn_apply")
    return y, stats


USE_GN_FUSED = True      # GroupNorm backward as one persistent kernel (2 reads + 1 write of HBM) instead of 5 passes


def gn_bwd_kernel_name():
    """Which kernels gn_bwd launches (for bench.py's roofline_gn label)."""
    return "gn_bwd_fused_kernel" if USE_GN_FUSED else "gn_bwd_partial + gn_bwd_reduce + gn_bwd_apply"


def gn_bwd_fused_plan(nb, hw, c):
    """(teams = images in flight, CTAs per team, rows per CTA, grid) of the fused backward for an [nb,hw,c] tensor."""
    out = (C.c_int * 4)()
    _cabi.lib().b2dq_gn_bwd_fused_plan(nb, hw, c, out)
    return tuple(out)


def gn_bwd(dy, x, stats, gamma, beta, swish, groups=32, add=None, ws_nc=None):
    """Returns (dx bf16, dgamma f32, dbeta f32); add (bf16, like x) is summed into dx (residual gradient);
    ws_nc [N,C,2]: the reduction pass was already done by the producer of dy (pconv dgrad epilogue)."""
    nb, h, w, c = x.shape
    l = _cabi.lib()
    if ws_nc is None and USE_GN_FUSED and int(swish) in (0, 1):
        ws_bytes = l.b2dq_gn_bwd_fused_workspace_bytes(nb, h * w, c, groups)
        if ws_bytes > 0:
            dx = torch.empty_like(x)
            dgb = torch.empty(2, c, dtype=torch.float32, device=x.device)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            check(l.b2dq_gn_bwd_fused(_ptr(dy), _ptr(x), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(dx), _ptr(dgb),
                                      _ptr(add), _ptr(ws), ws_bytes, nb, h * w, c, groups, int(swish), _stream()),
                  "gn_bwd_fused")
            return dx, dgb[0], dgb[1]
    if ws_nc is not None:
        dx = torch.empty_like(x)
        dgb = torch.empty(2, c, dtype=torch.float32, device=x.device)
        check(l.b2dq_gn_bwd_apply(_ptr(dy), _ptr(x), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(ws_nc), _ptr(dx),
                                  _ptr(dgb), _ptr(add), nb, h * w, c, groups, int(swish), _stream()), "gn_bwd_apply")
        return dx, dgb[0], dgb[1]
    g = _gn_groups(nb, h * w * c * 2, 2)
    ws = torch.empty(nb, c, 2, dtype=torch.float32, device=x.device)
    part = torch.empty(g * l.b2dq_gn_chunks(g, h * w) * c * 2, dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x)
    dgb = torch.empty(2, c, dtype=torch.float32, device=x.device)
    for n0 in range(0, nb, g):
        n = min(g, nb - n0)
        dys, xs, ss, wss, dxs = dy[n0:n0 + n], x[n0:n0 + n], stats[n0:n0 + n], ws[n0:n0 + n], dx[n0:n0 + n]
        check(l.b2dq_gn_bwd_stats(_ptr(dys), _ptr(xs), _ptr(ss), _ptr(gamma), _ptr(beta), _ptr(part), _ptr(wss),
                                  n, h * w, c, groups, int(swish), _stream()), "gn_bwd_stats")
        adds = None if add is None else add[n0:n0 + n]
        check(l.b2dq_gn_bwd_apply(_ptr(dys), _ptr(xs), _ptr(ss), _ptr(gamma), _ptr(beta), _ptr(wss), _ptr(dxs),
                                  None, _ptr(adds), n, h * w, c, groups, int(swish), _stream()), "gn_bwd_apply_nodgb")
    check(l.b2dq_gn_bwd_param(_ptr(ws), _ptr(dgb), nb, c, _stream()), "gn_bwd_param")
    return dx, dgb[0], dgb[1]


# ------------------------------------------------------------------------------------------ misc
def nchw_f32_to_nhwc_bf16(x):
    nb, c, h, w = x.shape
    out = torch.empty(nb, h, w, c, dtype=BF16, device=x.device)
    check(_cabi.lib().b2dq_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(out), nb, c, h * w, _stream()), "to_nhwc")
    return out


def nhwc_bf16_to_nchw_f32(x):
    nb, h, w, c = x.shape
    out = torch.empty(nb, c, h, w, dtype=torch.float32, device=x.device)
    check(_cabi.lib().b2dq_nhwc_bf16_to_nchw_f32(_ptr(x), _ptr(out), nb, c, h * w, _stream()), "to_nchw")
    return out


def nhwc_f32_to_nchw_f32(x):
    nb, h, w, c = x.shape
    out = torch.empty(nb, c, h, w, dtype=torch.float32, device=x.device)
    check(_cabi.lib().b2dq_nhwc_f32_to_nchw_f32(_ptr(x), _ptr(out), nb, c, h * w, _stream()), "to_nchw32")
    return out


def nchw_f32_to_nhwc_f32(x):
    nb, c, h, w = x.shape
    out = torch.empty(nb, h, w, c, dtype=torch.float32, device=x.device)
    check(_cabi.lib().b2dq_nchw_f32_to_nhwc_f32(_ptr(x), _ptr(out), nb, c, h * w, _stream()), "to_nhwc32")
    return out


def upsample2x(x):
    nb, h, w, c = x.shape
    out = torch.empty(nb, 2 * h, 2 * w, c, dtype=BF16, device=x.device)
    check(_cabi.lib().b2dq_upsample2x(_ptr(x), _ptr(out), nb, h, w, c, _stream()), "upsample2x")
    return out


def upsample2x_bwd(g):
    nb, h2, w2, c = g.shape
    out = torch.empty(nb, h2 // 2, w2 // 2, c, dtype=BF16, device=g.device)
    check(_cabi.lib().b2dq_upsample2x_bwd(_ptr(g), _ptr(out), nb, h2 // 2, w2 // 2, c, _stream()),
          "upsample2x_bwd")
    return out


def softmax_rows(s, T):
    p = torch.empty(s.shape, dtype=BF16, device=s.device)
    check(_cabi.lib().b2dq_softmax_rows(_ptr(s), _ptr(p), s.numel() // T, T,
                                        int(s.dtype == torch.float32), _stream()), "softmax_rows")
    return p


def softmax_bwd_rows(p, dp, T, scale):
    ds = torch.empty_like(p)
    check(_cabi.lib().b2dq_softmax_bwd_rows(_ptr(p), _ptr(dp), _ptr(ds), p.numel() // T, T, float(scale),
                                            _stream()), "softmax_bwd_rows")
    return ds


def add_bf16(a, b):
    o = torch.empty_like(a)
    check(_cabi.lib().b2dq_add_bf16(_ptr(a), _ptr(b), _ptr(o), a.numel(), _stream()), "add_bf16")
    return o


def patch_entropy(x_nchw, bins, patch, sigma):
    """x [B,3,H,W] fp32 NCHW -> [B, H/patch, W/patch] fp32 (dqvae_dual_entropy.py:25-63)."""
    assert x_nchw.dtype == torch.float32 and x_nchw.dim() == 4 and x_nchw.shape[1] == 3
    x_nchw = x_nchw.contiguous()
    b, _, h, w = x_nchw.shape
    out = torch.empty(b, h // patch, w // patch, dtype=torch.float32, device=x_nchw.device)
    check(_cabi.lib().b2dq_patch_entropy(_ptr(x_nchw), _ptr(bins), _ptr(out), b, h, w, patch, bins.numel(),
                                         float(sigma), _stream()), "patch_entropy")
    return out


def permuter_forward(indices, grain, coarse_hw, fine_hw, coarse_len, fine_len, region_first, codes6):
    """permuter.py:50-109.  indices [B,F,F], grain [B,Hc,Hc] int64 -> six [B,L] int64 tensors."""
    import ctypes
    b = indices.shape[0]
    dev = indices.device
    outs = [torch.empty(b, n, dtype=torch.int64, device=dev) for n in (coarse_len,) * 3 + (fine_len,) * 3]
    arr = (ctypes.c_longlong * 6)(*[int(c) for c in codes6])
    check(_cabi.lib().b2dq_permuter_forward(_ptr(indices), _ptr(grain), *[_ptr(o) for o in outs], b, coarse_hw,
                                            fine_hw, coarse_len, fine_len, int(bool(region_first)), arr,
                                            _stream()), "permuter_forward")
    return outs


def permuter_backward(coarse_content, fine_content, coarse_position, fine_position, coarse_hw, fine_hw,
                      coarse_position_eos, fine_position_eos):
    """permuter.py:111-132 -> target [B,F,F] int64."""
    b = coarse_content.shape[0]
    out = torch.empty(b, fine_hw, fine_hw, dtype=torch.int64, device=coarse_content.device)
    check(_cabi.lib().b2dq_permuter_backward(_ptr(coarse_content), _ptr(fine_content), _ptr(coarse_position),
                                             _ptr(fine_position), _ptr(out), b, coarse_hw, fine_hw,
                                             coarse_content.shape[1], fine_content.shape[1],
                                             int(coarse_position_eos), int(fine_position_eos), _stream()),
          "permuter_backward")
    return out


def im2col3x3_small(x, flip=False):
    nb, h, w, cs = x.shape
    out = torch.empty(nb, h, w, 64, dtype=BF16, device=x.device)
    check(_cabi.lib().b2dq_im2col3x3_small(_ptr(x), _ptr(out), nb, h, w, cs, int(flip), _stream()),
          "im2col3x3_small")
    return out


def im2col_window(x, k, stride, sgn, off, out_hw):
    """x [N,Hs,Ws,Cs] bf16 (k*k*Cs <= 64) -> [N,Ho,Wo,64] bf16: column (r*k+s)*Cs + c holds
    x[n, oh*stride + sgn*r + off, ow*stride + sgn*s + off, c] (zero outside x / in the unused columns)."""
    nb, hs, ws, cs = x.shape
    ho, wo = out_hw
    out = torch.empty(nb, ho, wo, 64, dtype=BF16, device=x.device)
    check(_cabi.lib().b2dq_im2col_window(_ptr(x), _ptr(out), nb, hs, ws, ho, wo, cs, k, stride, sgn, off, _stream()),
          "im2col_window")
    return out


def lrelu_bwd(dy, y, slope=0.2):
    """dy * (y > 0 ? 1 : slope) for a LeakyReLU whose OUTPUT is y (bf16, same shape)."""
    dx = torch.empty_like(dy)
    check(_cabi.lib().b2dq_lrelu_bwd(_ptr(dy), _ptr(y), _ptr(dx), dy.numel(), float(slope), _stream()), "lrelu_bwd")
    return dx


# ------------------------------------------------------------------------------------------ 4x4 convolutions (PatchGAN)
# modules/discriminator/model.py:37-66: Conv2d(k=4, stride 2 | 1, padding 1).  y[oh,ow] = sum_{r,s} W[r,s] x[st*oh-1+r, st*ow-1+s].
def _taps4(stride, cin):
    """(channel offset, dw, row parity, dh, weight column) of the 16 filter taps in the view conv4x4_fwd reads."""
    taps = []
    for r in range(4):
        for s in range(4):
            if stride == 2:                       # parity view [N, H/2, 2, W/2, 2C]: row 2oh-1+r = 2(oh+dh)+p
                dh, p = divmod(r - 1, 2)
                dw, q = divmod(s - 1, 2)
                taps.append((q * cin, dw, p, dh, (r * 4 + s) * cin))
            else:
                taps.append((0, s - 1, 0, r - 1, (r * 4 + s) * cin))
    return taps


def conv4x4_out_hw(h, w, stride):
    return (h // 2, w // 2) if stride == 2 else (h - 1, w - 1)


def conv4x4_fwd(x, wpack, bias, stride, cout, out_f32=False, block_n=0, act=0):
    """x NHWC bf16 [N,H,W,Cin] (Cin % 64 == 0); wpack [Cout(+pad), 16*Cin] from pack_weight_fwd."""
    nb, h, w, cin = x.shape
    assert cin % 64 == 0 and (stride == 1 or (h % 2 == 0 and w % 2 == 0))
    ho, wo = conv4x4_out_hw(h, w, stride)
    dims, strs = parity_view(x) if stride == 2 else nhwc_view(x)
    out = torch.empty(nb, ho, wo, cout, dtype=torch.float32 if out_f32 else BF16, device=x.device)
    ostr = (ho * wo * cout, wo * cout, cout)
    tapgemm(x, dims, strs, wpack, wpack.shape[0], wpack.shape[1], _taps4(stride, cin), cin // 64, out, 0, ostr, wo, ho,
            nb, cout, bias=bias, out_f32=out_f32, block_n=block_n, relu=act)
    return out


def conv4x4_dgrad(dy, wdpack, stride, cin, in_hw):
    """dy NHWC bf16 [N,Ho,Wo,Cout] -> dx NHWC bf16 [N,H,W,Cin]; wdpack [Cin, 16*Cout] from pack_weight_dgrad."""
    nb, ho, wo, cout = dy.shape
    assert cout % 64 == 0
    h, w = in_hw
    dx = torch.empty(nb, h, w, cin, dtype=BF16, device=dy.device)
    dims, strs = nhwc_view(dy)
    kch = cout // 64
    if stride == 1:                               # dx[ih,iw] = sum W[r,s]^T dy[ih+1-r, iw+1-s]
        taps = [(0, 1 - s, 0, 1 - r, (r * 4 + s) * cout) for r in range(4) for s in range(4)]
        tapgemm(dy, dims, strs, wdpack, wdpack.shape[0], wdpack.shape[1], taps, kch, dx, 0, (h * w * cin, w * cin, cin),
                w, h, nb, cin)
        return dx
    # stride 2: input row 2i+ph receives filter rows r with r = ph+1 (mod 2), from output row i + (ph+1-r)/2
    rsel = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}
    for ph in (0, 1):
        for pw in (0, 1):
            taps = [(0, dw, 0, dh, (r * 4 + s) * cout) for r, dh in rsel[ph] for s, dw in rsel[pw]]
            tapgemm(dy, dims, strs, wdpack, wdpack.shape[0], wdpack.shape[1], taps, kch, dx, (ph * w + pw) * cin,
                    (h * w * cin, 2 * w * cin, 2 * cin), wo, ho, nb, cin)
    return dx


def conv4x4_wgrad(x, dy, stride):
    """dW fp32 OIHW [Cout,Cin,4,4] of y = conv4x4(x): the 16 taps go through the weight-gradient GEMM in two launches
    of 8 (its tap table holds 12)."""
    nb, h, w, cin = x.shape
    _, ho, wo, cout = dy.shape
    bdims, bstrs = parity_view(x) if stride == 2 else nhwc_view(x)
    taps = [t[:4] for t in _taps4(stride, cin)]
    adims, astrs = nhwc_view(dy)
    kw, kh, kn = tile_shape(wo, ho, nb, pixels=64)
    ktw, kth = (wo + kw - 1) // kw, (ho + kh - 1) // kh
    kblocks = ktw * kth * ((nb + kn - 1) // kn)
    mt, nt = (cout + 127) // 128, (cin + 127) // 128
    splits = _wgrad_splits(kblocks, mt * nt * 3)
    partial = torch.empty(splits, 16, cout, cin, dtype=torch.float32, device=x.device)
    for half in (0, 1):
        mmgemm(dy, adims, astrs, True, x, bdims, bstrs, True, cout, cin, kblocks, partial,
               (16 * cout * cin, cout * cin, cin), taps=taps[8 * half:8 * half + 8], kbox=(kw, kh, kn), ktiles=(ktw, kth),
               splits=splits, out_f32=True, block_n=128, taps_per_cta=3, out_off=8 * half * cout * cin)
    dw = torch.empty(cout, cin, 4, 4, dtype=torch.float32, device=x.device)
    check(_cabi.lib().b2dq_wgrad_reduce(_ptr(partial), _ptr(dw), splits, 16, cout, cin, 0, _stream()), "wgrad_reduce")
    return dw


# ------------------------------------------------------------------------------------------ BatchNorm (PatchGAN)
# nn.BatchNorm2d over [N,H,W] per channel = the GroupNorm kernels on the tensor seen as ONE image with one group per
# channel (N = 1, G = C); activation code 2 = LeakyReLU(0.2).
def bn_stats(x, eps=1e-5):
    """x NHWC bf16 -> stats [1, C, 2] = (batch mean, 1/sqrt(biased batch variance + eps)) per channel."""
    nb, h, w, c = x.shape
    return gn_stats(x.view(1, nb * h, w, c), groups=c, eps=eps)


def bn_apply(x, stats, gamma, beta, act):
    nb, h, w, c = x.shape
    return gn_apply(x.view(1, nb * h, w, c), stats, gamma, beta, act, groups=c).view(nb, h, w, c)


def bn_bwd(dy, x, stats, gamma, beta, act, batch_stats=True):
    """-> (dx, dgamma, dbeta).  batch_stats=False (eval mode: the statistics are constants, not functions of x):
    dx = dz * gamma * rstd without the two mean-subtraction terms."""
    nb, h, w, c = x.shape
    dyv, xv = dy.view(1, nb * h, w, c), x.view(1, nb * h, w, c)
    if batch_stats:
        dx, dg, db = gn_bwd(dyv, xv, stats, gamma, beta, act, groups=c)
        return dx.view(nb, h, w, c), dg, db
    l = _cabi.lib()
    hw = nb * h * w
    ws = torch.empty(1, c, 2, dtype=torch.float32, device=x.device)
    part = torch.empty(l.b2dq_gn_chunks(1, hw) * c * 2, dtype=torch.float32, device=x.device)
    check(l.b2dq_gn_bwd_stats(_ptr(dyv), _ptr(xv), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(part), _ptr(ws), 1, hw, c,
                              c, int(act), _stream()), "gn_bwd_stats")
    dgb = torch.empty(2, c, dtype=torch.float32, device=x.device)
    check(l.b2dq_gn_bwd_param(_ptr(ws), _ptr(dgb), 1, c, _stream()), "gn_bwd_param")
    dx = torch.empty_like(x)
    zero = torch.zeros_like(ws)
    check(l.b2dq_gn_bwd_apply(_ptr(dyv), _ptr(xv), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(zero), _ptr(dx), None, None,
                              1, hw, c, c, int(act), _stream()), "gn_bwd_apply_nodgb")
    return dx, dgb[0], dgb[1]


# ------------------------------------------------------------------------------------------ VGG pieces
def maxpool2x2(x):
    """2x2 / stride 2 max pooling of NHWC bf16."""
    nb, h, w, c = x.shape
    y = torch.empty(nb, h // 2, w // 2, c, dtype=BF16, device=x.device)
    check(_cabi.lib().b2dq_maxpool2x2(_ptr(x), _ptr(y), nb, h, w, c, _stream()), "maxpool2x2")
    return y


def maxpool2x2_bwd(dy, x):
    """Gradient of maxpool2x2 w.r.t. its input x (first maximum of a window gets the gradient)."""
    nb, h, w, c = x.shape
    dx = torch.empty_like(x)
    check(_cabi.lib().b2dq_maxpool2x2_bwd(_ptr(dy), _ptr(x), _ptr(dx), nb, h, w, c, _stream()), "maxpool2x2_bwd")
    return dx


def relu_bwd(dy, y):
    """dy * (y > 0) for a ReLU whose OUTPUT is y (bf16, same shape)."""
    dx = torch.empty_like(dy)
    check(_cabi.lib().b2dq_relu_bwd(_ptr(dy), _ptr(y), _ptr(dx), dy.numel(), _stream()), "relu_bwd")
    return dx


def lpips_head_fwd(f0, f1, w, seed=None, p_drop=0.0):
    """f0, f1 NHWC bf16 [N,H,W,C]; w fp32 [C]; seed: int64 CUDA tensor [1] (dropout) or None.
    Returns the per-image spatial mean [N] (fp32)."""
    nb, h, wd, c = f0.shape
    lib = _cabi.lib()
    chunks = lib.b2dq_lpips_head_chunks(nb, h * wd)
    part = torch.empty(nb, chunks, dtype=torch.float32, device=f0.device)
    check(lib.b2dq_lpips_head_fwd(_ptr(f0), _ptr(f1), _ptr(w), _ptr(part), nb, h * wd, c, _ptr(seed), float(p_drop),
                                  _stream()), "lpips_head_fwd")
    return part.sum(1) / float(h * wd)


def lpips_head_bwd(f0, f1, w, g, want0, want1, seed=None, p_drop=0.0):
    """g fp32 [N] = gradient w.r.t. the spatial means.  Returns (df0 | None, df1 | None) in bf16."""
    nb, h, wd, c = f0.shape
    d0 = torch.empty_like(f0) if want0 else None
    d1 = torch.empty_like(f1) if want1 else None
    check(_cabi.lib().b2dq_lpips_head_bwd(_ptr(f0), _ptr(f1), _ptr(w), _ptr(g), _ptr(d0), _ptr(d1), nb, h * wd, c,
                                          _ptr(seed), float(p_drop), _stream()), "lpips_head_bwd")
    return d0, d1
