"""Sibling quantizers of the reference that share the DQ-VAE search path, all
on the fused sm_100a search kernel (csrc/vq.cu) - no [N,K] distance matrix, no one-hot:

* ``VectorQuantize2``      - ``modules/vector_quantization/quantize2.py:135-209`` (no grain mask,
                              ``commit_loss_legacy`` switch),
* ``VectorQuantize2List``  - ``quantize2_list.py:134-187`` (ragged list of [n_i, C] sequences, the EMA
                              codebook moves between items),
* ``RQBottleneck``         - ``quantize_rqvae.py:147-400`` (residual quantization, depth loop over
                              shared or separate EMA codebooks),
* ``VectorQuantizer2``     - ``quantize_vqgan.py:213-341`` (learnable codebook: the gradient w.r.t. the
                              embedding is assembled from the per-code counts / sums the search kernel
                              accumulates, count_k * e_k - sum_k, instead of a scatter-add).

Same constructor arguments, buffers / parameters and return conventions as the reference classes.
There is no CPU path: every forward needs CUDA tensors and the built extension.
"""
from typing import Iterable

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .. import ops
from .quantize import BF16, VQEmbedding, _SearchOperands, _VQFn


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("the B200 quantizers need CUDA tensors; there is no CPU fallback")


def _search_rows(emb, rows_f32, value_scale, grad_scale):
    """rows [N,C] fp32 (autograd-tracked) through one EMA codebook: (xq rows fp32, loss, codes), with the
    EMA / restart / re-normalisation applied afterwards in training (quantize2.py:117-127)."""
    train = emb.training and emb.ema
    if train:
        emb._ensure_acc()
    rows_f32 = rows_f32.float().contiguous()
    xq, loss, codes = _VQFn.apply(rows_f32.detach().to(BF16), rows_f32, None, emb, value_scale, grad_scale, train)
    if train:
        det = rows_f32.detach()
        emb._ema_step(lambda idx: det if idx is None else det[idx], det.shape[0])
    return xq, loss, codes


def _soft_codes(codebook, x, temp, stochastic):
    distances = codebook.compute_distances(x)
    soft_code = F.softmax(-distances / temp, dim=-1)
    if stochastic:
        flat = soft_code.reshape(-1, soft_code.shape[-1])
        code = torch.multinomial(flat, 1).reshape(*soft_code.shape[:-1])
    else:
        code = distances.argmin(dim=-1)
    return soft_code, code


class VectorQuantize2(nn.Module):
    """quantize2.py:135-209."""

    def __init__(self, codebook_size, codebook_dim=None, accept_image_fmap=True, commitment_beta=0.25,
                 decay=0.99, restart_unused_codes=True, channel_last=False, commit_loss_legacy=True):
        super().__init__()
        self.accept_image_fmap = accept_image_fmap
        self.beta = commitment_beta
        self.channel_last = channel_last
        self.restart_unused_codes = restart_unused_codes
        self.commit_loss_legacy = commit_loss_legacy
        self.codebook = VQEmbedding(codebook_size, codebook_dim, decay=decay,
                                    restart_unused_codes=restart_unused_codes)
        self.codebook.weight.data.uniform_(-1.0 / codebook_size, 1.0 / codebook_size)

    def forward(self, x, *ignorewargs, **ignorekwargs):
        _require_cuda(x)
        need_transpose = not self.channel_last and not self.accept_image_fmap
        if self.accept_image_fmap:
            b, c, height, width = x.shape
            rows = ops.ToNCHWInvFn.apply(x)                      # [B,H,W,C] fp32, differentiable
        else:
            rows = x.transpose(1, 2) if need_transpose else x
        shape = rows.shape
        # legacy (:175): beta*mean((sg(xq)-x)^2) + mean((xq-sg(x))^2), the codebook is EMA-frozen;
        # otherwise (:178): mean((x-sg(xq))^2)
        scales = (1.0 + self.beta, self.beta) if self.commit_loss_legacy else (1.0, 1.0)
        xq, loss, codes = _search_rows(self.codebook, rows.reshape(-1, shape[-1]), *scales)
        x_q = xq.reshape(shape)
        if self.accept_image_fmap:
            x_q = ops.to_nchw(x_q)
            codes = codes.reshape(b, height, width)
        else:
            codes = codes.reshape(shape[:-1])
            if need_transpose:
                x_q = x_q.transpose(1, 2).contiguous()
        return x_q, loss, (None, None, codes)

    @torch.no_grad()
    def get_soft_codes(self, x, temp=1.0, stochastic=False):
        return _soft_codes(self.codebook, x, temp, stochastic)

    def get_codebook_entry(self, indices, *kwargs):
        return self.codebook.embed(indices)


class VectorQuantize2List(nn.Module):
    """quantize2_list.py:134-187 (class name there: VectorQuantize2)."""

    def __init__(self, codebook_size, codebook_dim=None, commitment_beta=0.25, decay=0.99,
                 restart_unused_codes=True):
        super().__init__()
        self.beta = commitment_beta
        self.restart_unused_codes = restart_unused_codes
        self.codebook = VQEmbedding(codebook_size, codebook_dim, decay=decay,
                                    restart_unused_codes=restart_unused_codes)
        self.codebook.weight.data.uniform_(-1.0 / codebook_size, 1.0 / codebook_size)

    def forward(self, x_list, *ignorewargs, **ignorekwargs):
        batch_size = len(x_list)
        x_q_list, x_code_list = [], []
        loss = 0.
        for x in x_list:
            _require_cuda(x)
            shape = x.shape
            # one search per item, in order: in training the codebook an item sees already contains the
            # EMA updates of the items before it (:117-127 runs inside every self.codebook(...) call)
            xq, loss_i, codes = _search_rows(self.codebook, x.reshape(-1, shape[-1]), 1.0 + self.beta, self.beta)
            loss = loss + loss_i
            x_q_list.append(xq.reshape(shape))
            x_code_list.append(codes.reshape(shape[:-1]))
        loss = loss / batch_size
        return x_q_list, loss, (None, None, x_code_list)

    @torch.no_grad()
    def get_soft_codes(self, x, temp=1.0, stochastic=False):
        return _soft_codes(self.codebook, x, temp, stochastic)

    def get_codebook_entry(self, indices, *kwargs):
        return self.codebook.embed(indices)


class _RQFn(torch.autograd.Function):
    """Residual quantization of rows [N,C] over `depth` codebooks (quantize_rqvae.py:239-296).

    forward: (aggregated quants [N,C], loss = mean_d mean((x-agg_d)^2), codes [N,depth]).
    backward: straight-through for the quants; d loss / d x = 2/(N*C*depth) * sum_d (x - agg_d), and
    x - agg_d is exactly the residual left after depth d, so the loop keeps their running sum."""

    @staticmethod
    def forward(ctx, rows, bottleneck):
        n, c = rows.shape
        depth = len(bottleneck.codebooks)
        residual = rows.detach().clone()
        agg = torch.zeros_like(residual)
        res_sum = torch.zeros_like(residual)
        loss = torch.zeros((), dtype=torch.float32, device=rows.device)
        codes = []
        for d in range(depth):
            emb = bottleneck.codebooks[d]
            train = emb.training and emb.ema
            if train:
                emb._ensure_acc()
            with torch.no_grad():
                xq, part, code = _VQFn.apply(residual.to(BF16), residual, None, emb, 1.0, 0.0, train)
            if train:
                det = residual
                emb._ema_step(lambda idx, det=det: det if idx is None else det[idx], n)
            residual = residual - xq
            agg = agg + xq
            res_sum += residual
            loss = loss + part                       # part = mean((residual_in - e)^2) = mean((x - agg_d)^2)
            codes.append(code.unsqueeze(-1))
        ctx.save_for_backward(res_sum)
        ctx.coef = 2.0 / float(n * c * depth)
        codes = torch.cat(codes, dim=-1)
        ctx.mark_non_differentiable(codes)
        return agg, loss / depth, codes

    @staticmethod
    def backward(ctx, g_q, g_loss, _g_codes):
        (res_sum,) = ctx.saved_tensors
        g = torch.zeros_like(res_sum) if g_q is None else g_q
        if g_loss is not None:
            g = g + (ctx.coef * g_loss) * res_sum
        return g, None


class RQBottleneck(nn.Module):
    """quantize_rqvae.py:147-400: residual quantization with `code_shape[-1]` EMA codebooks."""

    def __init__(self, latent_shape, code_shape, n_embed, decay=0.99, shared_codebook=False,
                 restart_unused_codes=True, commitment_loss="cumsum"):
        super().__init__()
        if not len(code_shape) == len(latent_shape) == 3:
            raise ValueError("incompatible code shape or latent shape")
        if any([y % x != 0 for x, y in zip(code_shape[:2], latent_shape[:2])]):
            raise ValueError("incompatible code shape or latent shape")
        embed_dim = int(np.prod(latent_shape[:2]) // np.prod(code_shape[:2]) * latent_shape[2])
        self.latent_shape = torch.Size(latent_shape)
        self.code_shape = torch.Size(code_shape)
        self.shape_divisor = torch.Size([latent_shape[i] // code_shape[i] for i in range(len(latent_shape))])
        self.shared_codebook = shared_codebook
        if self.shared_codebook:
            if isinstance(n_embed, Iterable) or isinstance(decay, Iterable):
                raise ValueError("Shared codebooks are incompatible with list types of momentums or sizes: "
                                 "Change it into int")
        self.restart_unused_codes = restart_unused_codes
        depth = self.code_shape[-1]
        self.n_embed = n_embed if isinstance(n_embed, Iterable) else [n_embed for _ in range(depth)]
        self.decay = decay if isinstance(decay, Iterable) else [decay for _ in range(depth)]
        assert len(self.n_embed) == depth
        assert len(self.decay) == depth
        if self.shared_codebook:
            codebook0 = VQEmbedding(self.n_embed[0], embed_dim, decay=self.decay[0],
                                    restart_unused_codes=restart_unused_codes)
            self.codebooks = nn.ModuleList([codebook0 for _ in range(depth)])
        else:
            self.codebooks = nn.ModuleList([
                VQEmbedding(self.n_embed[idx], embed_dim, decay=self.decay[idx],
                            restart_unused_codes=restart_unused_codes) for idx in range(depth)])
        self.commitment_loss = commitment_loss

    def to_code_shape(self, x):
        (B, H, W, D) = x.shape
        (rH, rW, _) = self.shape_divisor
        x = x.reshape(B, H // rH, rH, W // rW, rW, D).permute(0, 1, 3, 2, 4, 5)
        return x.reshape(B, H // rH, W // rW, -1)

    def to_latent_shape(self, x):
        (B, h, w, _) = x.shape
        (_, _, D) = self.latent_shape
        (rH, rW, _) = self.shape_divisor
        x = x.reshape(B, h, w, rH, rW, D).permute(0, 1, 3, 2, 4, 5)
        return x.reshape(B, h * rH, w * rW, D)

    def _run(self, x_reshaped):
        shape = x_reshaped.shape
        rows = x_reshaped.reshape(-1, shape[-1]).float().contiguous()
        agg, loss, codes = _RQFn.apply(rows, self)
        return agg.reshape(shape), loss, codes.reshape(*shape[:-1], -1)

    @torch.no_grad()
    def quantize(self, x):
        """x [B,h,w,embed_dim] -> (list of the aggregated quants after each depth, codes [B,h,w,d])."""
        _require_cuda(x)
        residual = x.detach().clone().float()
        agg = torch.zeros_like(residual)
        quant_list, code_list = [], []
        for emb in self.codebooks:
            quant, code = emb(residual)
            residual.sub_(quant)
            agg.add_(quant)
            quant_list.append(agg.clone())
            code_list.append(code.unsqueeze(-1))
        return quant_list, torch.cat(code_list, dim=-1)

    def forward(self, x):
        _require_cuda(x)
        x_reshaped = self.to_code_shape(x)
        agg, commitment_loss, codes = self._run(x_reshaped)
        # straight-through (:279): value x + (q - x), gradient of the quants flows to x unchanged
        quants_trunc = x + (self.to_latent_shape(agg) - x).detach()
        return quants_trunc, commitment_loss, codes

    def compute_commitment_loss(self, x, quant_list):
        loss_list = [(x - quant.detach()).pow(2.0).mean() for quant in quant_list]
        return torch.mean(torch.stack(loss_list))

    def _embeds(self, code):
        code_slices = torch.chunk(code, chunks=code.shape[-1], dim=-1)
        if self.shared_codebook:
            return [self.codebooks[0].embed(s) for s in code_slices]
        return [self.codebooks[i].embed(s) for i, s in enumerate(code_slices)]

    @torch.no_grad()
    def embed_code(self, code):
        assert code.shape[1:] == self.code_shape
        embeds = torch.cat(self._embeds(code), dim=-2).sum(-2)
        return self.to_latent_shape(embeds)

    @torch.no_grad()
    def embed_code_with_depth(self, code, to_latent_shape=False):
        assert code.shape[-1] == self.code_shape[-1]
        embeds = self._embeds(code)
        if to_latent_shape:
            embeds = [self.to_latent_shape(embed.squeeze(-2)).unsqueeze(-2) for embed in embeds]
        return torch.cat(embeds, dim=-2), None

    @torch.no_grad()
    def embed_partial_code(self, code, code_idx, decode_type="select"):
        assert code.shape[1:] == self.code_shape
        assert code_idx < code.shape[-1]
        B, h, w, _ = code.shape
        embeds = self._embeds(code)
        if decode_type == "select":
            embeds = embeds[code_idx].view(B, h, w, -1)
        elif decode_type == "add":
            embeds = torch.cat(embeds[:code_idx + 1], dim=-2).sum(-2)
        else:
            raise NotImplementedError(f"{decode_type} is not implemented in partial decoding")
        return self.to_latent_shape(embeds)

    @torch.no_grad()
    def get_soft_codes(self, x, temp=1.0, stochastic=False):
        x = self.to_code_shape(x)
        residual = x.detach().clone()
        soft_code_list, code_list = [], []
        for codebook in self.codebooks:
            soft_code, code = _soft_codes(codebook, residual, temp, stochastic)
            residual -= codebook.embed(code)
            code_list.append(code.unsqueeze(-1))
            soft_code_list.append(soft_code.unsqueeze(-2))
        return torch.cat(soft_code_list, dim=-2), torch.cat(code_list, dim=-1)


class _LearnableSearch(_SearchOperands):
    """Search operands of a plain nn.Embedding (no padding row, trained by the optimizer)."""

    def __init__(self, embedding):
        self._embedding = embedding
        self.n_embed = embedding.num_embeddings

    @property
    def weight(self):
        return self._embedding.weight


class _VQLearnFn(torch.autograd.Function):
    """rows [N,C] fp32 + learnable codebook [K,C] -> (z_q rows, loss, codes) for quantize_vqgan.py:271-312.

    loss = (zs + es) * mean((e-z)^2); d loss/d z = zs * 2 (z-e)/(N*C); d loss/d e_k = es * 2/(N*C) *
    (count_k e_k - sum_k) with count_k / sum_k the per-code statistics the search kernel accumulated."""

    @staticmethod
    def forward(ctx, rows, weight, state, zs, es):
        n, c = rows.shape
        state._ensure_acc()
        state._acc.zero_()
        sums, counts = state._acc_views()
        loss_acc = torch.zeros(1, dtype=torch.float32, device=rows.device)
        from .. import kernels as kn
        w = weight.detach()
        codes, _, xq = kn.vq_search_gather(rows.to(BF16), state._codebook(), w, x_f32=rows, want_xq_bf16=False,
                                           want_xq_f32=True, counts=counts, sums=sums, loss_acc=loss_acc)
        ctx.save_for_backward(rows, xq, w, sums.clone(), counts.clone())
        ctx.zs, ctx.es, ctx.inv = zs, es, 2.0 / float(n * c)
        ctx.mark_non_differentiable(codes)
        return xq, loss_acc[0] * ((zs + es) / float(n * c)), codes

    @staticmethod
    def backward(ctx, g_zq, g_loss, _g_codes):
        rows, xq, w, sums, counts = ctx.saved_tensors
        g_rows = torch.zeros_like(rows) if g_zq is None else g_zq       # z + sg(z_q - z): identity to z
        g_w = None
        if g_loss is not None:
            g_rows = g_rows + (ctx.zs * ctx.inv * g_loss) * (rows - xq)
            g_w = (ctx.es * ctx.inv * g_loss) * (counts.unsqueeze(1) * w - sums)
        return g_rows, g_w, None, None, None


class VectorQuantizer2(nn.Module):
    """quantize_vqgan.py:213-341 (learnable codebook, optional index remapping)."""

    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.legacy = legacy
        self.embedding = nn.Embedding(self.n_e, self.e_dim)
        self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)
        self.remap = remap
        if self.remap is not None:
            self.register_buffer("used", torch.tensor(np.load(self.remap)))
            self.re_embed = self.used.shape[0]
            self.unknown_index = unknown_index                  # "random" or "extra" or integer
            if self.unknown_index == "extra":
                self.unknown_index = self.re_embed
                self.re_embed = self.re_embed + 1
            print(f"Remapping {self.n_e} indices to {self.re_embed} indices. "
                  f"Using {self.unknown_index} for unknown indices.")
        else:
            self.re_embed = n_e
        self.sane_index_shape = sane_index_shape
        self._search = _LearnableSearch(self.embedding)

    def remap_to_used(self, inds):
        ishape = inds.shape
        assert len(ishape) > 1
        inds = inds.reshape(ishape[0], -1)
        used = self.used.to(inds)
        match = (inds[:, :, None] == used[None, None, ...]).long()
        new = match.argmax(-1)
        unknown = match.sum(2) < 1
        if self.unknown_index == "random":
            new[unknown] = torch.randint(0, self.re_embed, size=new[unknown].shape).to(device=new.device)
        else:
            new[unknown] = self.unknown_index
        return new.reshape(ishape)

    def unmap_to_all(self, inds):
        ishape = inds.shape
        assert len(ishape) > 1
        inds = inds.reshape(ishape[0], -1)
        used = self.used.to(inds)
        if self.re_embed > self.used.shape[0]:                  # extra token
            inds[inds >= self.used.shape[0]] = 0
        back = torch.gather(used[None, :][inds.shape[0] * [0], :], 1, inds)
        return back.reshape(ishape)

    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        assert temp is None or temp == 1.0, "Only for interface compatible with Gumbel"
        assert rescale_logits is False, "Only for interface compatible with Gumbel"
        assert return_logits is False, "Only for interface compatible with Gumbel"
        _require_cuda(z)
        b, c, h, w = z.shape
        rows = ops.ToNCHWInvFn.apply(z).reshape(-1, self.e_dim).float().contiguous()
        # legacy (:295): mean((sg(zq)-z)^2) + beta*mean((zq-sg(z))^2); fixed (:292): beta on the first term
        zs, es = (1.0, self.beta) if self.legacy else (self.beta, 1.0)
        zq, loss, codes = _VQLearnFn.apply(rows, self.embedding.weight, self._search, zs, es)
        z_q = ops.to_nchw(zq.reshape(b, h, w, c))
        min_encoding_indices = codes
        if self.remap is not None:
            min_encoding_indices = self.remap_to_used(min_encoding_indices.reshape(b, -1)).reshape(-1, 1)
        if self.sane_index_shape:
            min_encoding_indices = min_encoding_indices.reshape(b, h, w)
        return z_q, loss, (None, None, min_encoding_indices)

    def get_codebook_entry(self, indices, shape=None):
        if self.remap is not None:
            indices = self.unmap_to_all(indices.reshape(shape[0], -1)).reshape(-1)
        z_q = self.embedding(indices)
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
        return z_q

    @torch.no_grad()
    def embed_code_with_depth(self, code, to_latent_shape=False):
        code_slices = torch.chunk(code, chunks=code.shape[-1], dim=-1)
        embeds = [self.embedding(code_slice) for code_slice in code_slices]
        if to_latent_shape:
            embeds = [self.to_latent_shape(embed.squeeze(-2)).unsqueeze(-2) for embed in embeds]
        return torch.cat(embeds, dim=-2), None
