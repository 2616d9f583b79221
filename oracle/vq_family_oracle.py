"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the sibling quantizers that share the DQ-VAE search path
(SURVEY.md 8f row 2): ``quantize2.VectorQuantize2``, ``quantize2_list.VectorQuantize2``,
``quantize_rqvae.RQBottleneck`` and ``quantize_vqgan.VectorQuantizer2`` of
``/root/reference/modules/vector_quantization/``.  Numpy restatements; every function cites the lines
it follows.  Only ``tests/`` may import this module - it is the checker, never the product path.

Pinning: no reference test covers these classes; the oracle is pinned against outputs of the
reference's own classes run in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/vq_family.npz``; checked by ``tests/test_oracle_golden.py``).
"""
import numpy as np

from . import vq_oracle as vo


def _mean(a):
    return np.float32(np.mean(a, dtype=np.float64))


def _nearest(x, weight, search_bf16=False):
    """Nearest code of every row (quantize2.py:30-55).  search_bf16: evaluate the distances on the
    bf16-rounded operands the CUDA search multiplies (the gather / loss keep the fp32 rows)."""
    if search_bf16:
        return vo.find_nearest_embedding(vo.bf16_round(x), np.concatenate([vo.bf16_round(weight[:-1]), weight[-1:]]))
    return vo.find_nearest_embedding(x, weight)


# ------------------------------------------------------------------ quantize2.VectorQuantize2
def vq2_forward(x, weight, beta=0.25, legacy=True, search_bf16=False):
    """quantize2.py:157-189 on flattened rows x [N,C] (eval).  Returns (x_q, loss, idx, d loss / d x).

    legacy (:175): beta*mean((sg(xq)-x)^2) + mean((xq-sg(x))^2); the codebook is EMA-frozen so only the
    first term has a gradient.  Otherwise (:178): mean((x-sg(xq))^2)."""
    idx = _nearest(x, weight, search_bf16)
    e = weight[idx]
    m = _mean((e - x) ** 2)
    if legacy:
        loss, gscale = np.float32(beta) * m + m, beta
    else:
        loss, gscale = m, 1.0
    gx = (2.0 * gscale / x.size) * (x - e)
    return x + (e - x), loss, idx, gx.astype(np.float32)


# ------------------------------------------------------------------ quantize2_list.VectorQuantize2
def vq2_list_forward(x_list, weight, beta=0.25, decay=0.99, train=False, cs=None, em=None, restart_rows=None,
                     eps=1e-5, search_bf16=False):
    """quantize2_list.py:148-166.  Ragged list of [n_i, C] rows; every item is searched against the
    codebook as it stands when the item is reached: in training the EMA update + re-normalisation
    (:117-127) runs after EACH item.  restart_rows[i] replays the rows item i would draw.
    Returns (xq_list, loss, idx_list, (weight, cs, em))."""
    w = weight.copy()
    xq_list, idx_list = [], []
    loss = np.float32(0.0)
    for i, x in enumerate(x_list):
        idx = _nearest(x, w, search_bf16)
        e = w[idx]
        if train:
            cs, em = vo.update_buffers(x, idx, cs, em, decay,
                                       restart_rows=None if restart_rows is None else restart_rows[i])
            w[:-1] = vo.update_embedding(cs, em, eps)
        m = _mean((e - x) ** 2)
        loss = loss + (np.float32(beta) * m + m)                         # :155
        xq_list.append(x + (e - x))                                      # :158
        idx_list.append(idx)
    return xq_list, np.float32(loss / len(x_list)), idx_list, (w, cs, em)


# ------------------------------------------------------------------ quantize_rqvae.RQBottleneck
def rq_to_code_shape(x, latent_shape, code_shape):
    """quantize_rqvae.py:216-225: [B,H,W,D] -> [B,h,w,rH*rW*D]."""
    B, H, W, D = x.shape
    rH, rW = latent_shape[0] // code_shape[0], latent_shape[1] // code_shape[1]
    x = x.reshape(B, H // rH, rH, W // rW, rW, D).transpose(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H // rH, W // rW, -1)


def rq_to_latent_shape(x, latent_shape, code_shape):
    """quantize_rqvae.py:227-237."""
    B, h, w, _ = x.shape
    D = latent_shape[2]
    rH, rW = latent_shape[0] // code_shape[0], latent_shape[1] // code_shape[1]
    x = x.reshape(B, h, w, rH, rW, D).transpose(0, 1, 3, 2, 4, 5)
    return x.reshape(B, h * rH, w * rW, D)


def rq_forward(x, weights, latent_shape, code_shape, train=False, states=None, decay=0.99, restart_rows=None,
               eps=1e-5, search_bf16=False):
    """quantize_rqvae.py:239-296.  x [B,H,W,D]; weights = one [K+1,C] array per depth (the SAME array
    object repeated when the codebook is shared).  Residual loop (:259-268): code d is the nearest row
    to the running residual; in training each depth's codebook is EMA-updated right after its search.
    Loss (:283-296) = mean over depth of mean((x - sg(agg_d))^2).  restart_rows[d]: the [K,C] rows depth d
    draws, or a callable residual_rows -> [K,C].
    Returns (quants [B,H,W,D], loss, codes [B,h,w,d], d loss / d x in code shape, states)."""
    xr = rq_to_code_shape(x, latent_shape, code_shape).astype(np.float32)
    shp = xr.shape
    rows = xr.reshape(-1, shp[-1])
    residual = rows.copy()
    agg = np.zeros_like(rows)
    depth = code_shape[-1]
    codes, losses = [], []
    gx = np.zeros_like(rows)
    for d in range(depth):
        w = weights[d]
        idx = _nearest(residual, w, search_bf16)
        e = w[idx]
        if train:
            cs, em = states[d]
            rr = None if restart_rows is None else restart_rows[d]
            if callable(rr):                                            # rows drawn from THIS depth's residual
                rr = rr(residual)
            cs, em = vo.update_buffers(residual, idx, cs, em, decay, restart_rows=rr)
            w[:-1] = vo.update_embedding(cs, em, eps)                   # in place: shared codebooks see it
            for j in range(depth):
                if weights[j] is w:
                    states[j] = (cs, em)
        residual = residual - e
        agg = agg + e
        losses.append(_mean((rows - agg) ** 2))
        gx += (2.0 / (rows.size * depth)) * (rows - agg)
        codes.append(idx.reshape(shp[:-1] + (1,)))
    quants = rq_to_latent_shape(agg.reshape(shp), latent_shape, code_shape)
    quants = x + (quants - x)                                           # :279
    return quants, np.float32(np.mean(losses)), np.concatenate(codes, -1), gx.reshape(shp), states


def rq_embed_code(codes, weights, latent_shape, code_shape):
    """quantize_rqvae.py:298-312: sum over depth of the code embeddings, back in latent shape."""
    emb = sum(weights[d][codes[..., d]] for d in range(codes.shape[-1]))
    return rq_to_latent_shape(emb, latent_shape, code_shape)


# ------------------------------------------------------------------ quantize_vqgan.VectorQuantizer2
def vqgan_forward(z, emb, beta=0.25, legacy=True, search_bf16=False):
    """quantize_vqgan.py:271-312 on rows z [N,C] with a LEARNABLE codebook emb [K,C] (no padding row).

    d = |z|^2 + |e|^2 - 2 z e^T (:280-282), argmin, straight-through.  legacy (:295): mean((sg(zq)-z)^2) +
    beta*mean((zq-sg(z))^2); else (:292) beta on the first term.
    Returns (z_q, loss, idx, d loss / d z, d loss / d emb)."""
    z = z.astype(np.float32)
    zs_, es_ = (vo.bf16_round(z), vo.bf16_round(emb)) if search_bf16 else (z, emb)
    d = (zs_ * zs_).sum(1, keepdims=True, dtype=np.float32) + (es_ * es_).sum(1, dtype=np.float32)[None] \
        - np.float32(2.0) * (zs_ @ es_.T)
    idx = d.argmin(1).astype(np.int64)
    e = emb[idx]
    m = _mean((e - z) ** 2)
    zs, es = (1.0, beta) if legacy else (beta, 1.0)
    loss = np.float32(zs) * m + np.float32(es) * m
    gz = (2.0 * zs / z.size) * (z - e)
    ge = np.zeros_like(emb, dtype=np.float64)
    np.add.at(ge, idx, (2.0 * es / z.size) * (e - z).astype(np.float64))
    return z + (e - z), loss, idx, gz.astype(np.float32), ge.astype(np.float32)
