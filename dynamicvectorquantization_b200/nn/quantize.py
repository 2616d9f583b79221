"""EMA vector quantizer of DQ-VAE on the fused sm_100a search kernel.

Mirror of ``modules/vector_quantization/quantize2_mask.py`` (reference): ``VQEmbedding`` (:10-132)
and ``VectorQuantize2`` (:135-209) with the same constructor arguments, parameters / buffers
(``weight [K+1,C]`` frozen, ``cluster_size_ema [K]``, ``embed_ema [K,C]``), forward contract
``forward(x, codebook_mask=None, ...) -> (x_q, loss, (None, None, codes))`` and helpers.

What differs is how it is computed: one kernel does distance + argmin + gather + masked loss +
per-code count/sum accumulation on bf16 operands with fp32 accumulation (no [N,K] matrix, no
one-hot), a second tiny pair of kernels applies the EMA / restart / re-normalisation.
"""
import numpy as np
import torch
import torch.distributed as dist
from torch import nn
from torch.nn import functional as F

from .. import kernels as kn
from .. import ops

BF16 = torch.bfloat16


class _VQFn(torch.autograd.Function):
    """rows [N,C] -> (x_q rows, value_scale * sum_rows m*|e-x|^2 / (N*C), codes); the gradient of the
    loss w.r.t. the rows is grad_scale * 2 m (x-e) / (N*C).  (value, grad) = (1+beta, beta) is the
    EMA-frozen commitment loss of quantize2_mask.py:172-179.

    fp32 mode (x_f32 given): x_q = exact fp32 codebook rows, loss/EMA from the fp32 rows.
    bf16 mode: everything from the bf16 rows (fused-model path)."""

    @staticmethod
    def forward(ctx, x_bf16, x_f32, row_mask, emb, value_scale, grad_scale, accumulate):
        n, c = x_bf16.shape
        loss_acc = torch.zeros(1, dtype=torch.float32, device=x_bf16.device)
        counts = sums = None
        if accumulate:
            emb._acc.zero_()
            sums, counts = emb._acc_views()
        w = emb.weight.detach()
        codes, xq_b, xq_f = kn.vq_search_gather(
            x_bf16, emb._codebook(), w, x_f32=x_f32, row_mask=row_mask,
            want_xq_bf16=x_f32 is None, want_xq_f32=x_f32 is not None,
            counts=counts, sums=sums, loss_acc=loss_acc)
        loss = loss_acc[0] * (value_scale / float(n * c))
        xq = xq_f if x_f32 is not None else xq_b
        ctx.save_for_backward(x_bf16 if x_f32 is None else x_f32, xq, row_mask)
        ctx.coef = 2.0 * grad_scale / float(n * c)
        ctx.mark_non_differentiable(codes)
        return xq, loss, codes

    @staticmethod
    def backward(ctx, g_xq, g_loss, _g_codes):
        x, xq, row_mask = ctx.saved_tensors
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=x.device)
        if g_xq is None:
            g_xq = torch.zeros_like(xq)
        if x.dtype == BF16:
            g = kn.vq_bwd(g_xq.contiguous(), x, xq, row_mask, g_loss.reshape(1).float().contiguous(), ctx.coef)
            return g, None, None, None, None, None, None
        m = 1.0 if row_mask is None else row_mask.unsqueeze(1)
        g = g_xq + (ctx.coef * g_loss) * m * (x - xq)
        return None, g, None, None, None, None, None


def reduce_ema_stats(acc):
    """Data-parallel exchange of the per-code statistics: ONE all-reduce of the packed
    [K*C sums | K counts] buffer instead of the reference's two (quantize2_mask.py:86-88)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def share_restart_rows(rows):
    """Every rank restarts dead codes from rank 0's candidate rows (quantize2_mask.py:99-100)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(rows, 0)
    return rows


class _SearchOperands:
    """What the search kernel needs besides the rows, for any owner of a ``weight`` ([K(+1), C] fp32)
    and ``n_embed``: the derived bf16 codebook + squared norms (refreshed when the weight changes) and
    the packed [K*C sums | K counts] accumulator the kernel adds per-code statistics into."""
    _cb = None
    _cb_key = None
    _acc = None

    # ---- derived search operands (bf16 codebook + squared norms), refreshed when weight changes
    def _codebook(self):
        w = self.weight
        key = (w.data_ptr(), w._version, self._dirty_tick if hasattr(self, "_dirty_tick") else 0)
        if self._cb is None or self._cb_key != key or self._cb.cb.device != w.device:
            if self._cb is None or self._cb.cb.device != w.device:
                self._cb = kn.Codebook(self.n_embed, w.shape[1], w.device)
            self._cb.refresh(w.detach().contiguous())
            self._cb_key = key
        return self._cb

    def _mark_dirty(self):
        self._dirty_tick = getattr(self, "_dirty_tick", 0) + 1

    def _acc_views(self):
        k, c = self.n_embed, self.weight.shape[1]
        return self._acc[:k * c].view(k, c), self._acc[k * c:]

    def _ensure_acc(self):
        k, c = self.n_embed, self.weight.shape[1]
        if self._acc is None or self._acc.device != self.weight.device:
            self._acc = torch.zeros(k * c + k, dtype=torch.float32, device=self.weight.device)


class VQEmbedding(_SearchOperands, nn.Embedding):
    """VQ embedding module with EMA update (quantize2_mask.py:10-132)."""

    def __init__(self, n_embed, embed_dim, ema=True, decay=0.99, restart_unused_codes=True, eps=1e-5):
        super().__init__(n_embed + 1, embed_dim, padding_idx=n_embed)
        self.ema = ema
        self.decay = decay
        self.eps = eps
        self.restart_unused_codes = restart_unused_codes
        self.n_embed = n_embed
        if not ema:
            # the reference trains a non-EMA codebook through the (x_q - sg(x))^2 term (:177-179); this class only
            # implements the EMA-frozen codebook every stage-1 config uses (learnable codebooks: quantize_vqgan)
            raise NotImplementedError("VQEmbedding(ema=False) is not implemented on the B200 path; use "
                                      "modules.vector_quantization.quantize_vqgan.VectorQuantizer2 for a learnable codebook")
        if embed_dim % 64 != 0 or embed_dim > 256:
            raise ValueError(f"VQEmbedding (B200): embed_dim must be a multiple of 64 and <= 256 (the search kernel "
                             f"stages the rows in 64-channel chunks), got {embed_dim}")
        if self.ema:
            _ = [p.requires_grad_(False) for p in self.parameters()]
            # padding index is not updated by EMA; embed_ema starts from the N(0,1) init (:27)
            self.register_buffer("cluster_size_ema", torch.zeros(n_embed))
            self.register_buffer("embed_ema", self.weight[:-1, :].detach().clone())
        self._cb = None
        self._cb_key = None
        self._acc = None
        # defer_ema: forward only accumulates the per-code statistics; the caller applies the EMA /
        # restart / re-normalisation later with apply_deferred_ema().  Result-identical (the gather
        # uses the pre-update codebook either way, :119-126) and lets a CUDA-graph-captured step keep
        # the NCCL exchange outside the graph.
        self.defer_ema = False
        self._deferred = None

    @torch.no_grad()
    def compute_distances(self, inputs):
        """||x||^2 + ||e||^2 - 2 x e^T with bf16 operands / fp32 accumulation (:29-48)."""
        c = self.weight.shape[1]
        shape = inputs.shape
        x = inputs.reshape(-1, c).to(BF16).contiguous()
        cb = self._codebook()
        n, k = x.shape[0], self.n_embed
        dot = torch.empty(n, k, dtype=torch.float32, device=x.device)
        kn.mmgemm(x, (c, n, 1, 1, 1), (1, c, n * c, n * c, n * c), False,
                  cb.cb, (c, k, 1, 1, 1), (1, c, k * c, k * c, k * c), False,
                  n, k, c // 64, dot, (0, 0, k), kbox=(64, 1, 1), ktiles=(c // 64, 1), alpha=-2.0,
                  out_f32=True, block_n=128)
        d = dot + x.float().pow(2).sum(1, keepdim=True) + cb.sqnorm[:k].unsqueeze(0)
        return d.reshape(*shape[:-1], k)

    @torch.no_grad()
    def find_nearest_embedding(self, inputs):
        c = self.weight.shape[1]
        x = inputs.reshape(-1, c).to(BF16).contiguous()
        codes, _, _ = kn.vq_search_gather(x, self._codebook(), self.weight.detach(), want_xq_bf16=False)
        return codes.reshape(inputs.shape[:-1])

    @torch.no_grad()
    def _tile_with_noise(self, x, target_n):
        b, embed_dim = x.shape
        n_repeats = (target_n + b - 1) // b
        std = x.new_ones(embed_dim) * 0.01 / np.sqrt(embed_dim)
        x = x.repeat(n_repeats, 1)
        return x + torch.rand_like(x) * std

    @torch.no_grad()
    def _ema_step(self, rows_f32_fn, n_vectors):
        """EMA + restart + re-normalisation after the search kernel has filled self._acc
        (:86-105 and :107-115).  rows_f32_fn(idx) returns fp32 input rows for the restart."""
        if self.defer_ema:
            self._deferred = (rows_f32_fn(None), n_vectors)     # the rows tensor itself (no closure: picklable)
            return
        self._ema_step_now(rows_f32_fn, n_vectors)

    @torch.no_grad()
    def apply_deferred_ema(self, keep=False):
        """Apply the EMA / restart / re-normalisation of the last deferred training forward, once: a second call
        without a new forward raises.  keep=True leaves it pending (a CUDA-graph replay refills the SAME rows
        buffer and statistics in place, so the owner of the graph re-applies it after every replay)."""
        if self._deferred is None:
            raise RuntimeError("apply_deferred_ema(): no deferred codebook update is pending")
        rows, n = self._deferred
        if not keep:
            self._deferred = None
        self._ema_step_now(lambda idx: rows if idx is None else rows[idx], n)

    @torch.no_grad()
    def _ema_step_now(self, rows_f32_fn, n_vectors):
        reduce_ema_stats(self._acc)
        sums, counts = self._acc_views()
        restart_rows = None
        if self.restart_unused_codes:
            k = self.n_embed
            if n_vectors < k:
                vectors = self._tile_with_noise(rows_f32_fn(None), k)
                restart_rows = vectors[torch.randperm(vectors.shape[0], device=vectors.device)][:k]
            else:
                perm = torch.randperm(n_vectors, device=self.weight.device)   # same RNG draw as :97
                restart_rows = rows_f32_fn(perm[:k])
            restart_rows = share_restart_rows(restart_rows.float().contiguous())
        kn.vq_ema_finalize(counts, sums, restart_rows, self.cluster_size_ema, self.embed_ema,
                           self.weight.data, self.decay, self.eps, self.restart_unused_codes)
        self._mark_dirty()

    def forward(self, inputs):
        """inputs [..., C] -> (embeds [..., C] fp32, idxs [...]) like the reference (:117-128)."""
        c = self.weight.shape[1]
        flat32 = inputs.reshape(-1, c).float().contiguous()
        train = self.training and self.ema
        if train:
            self._ensure_acc()
        xq, _, codes = _VQFn.apply(flat32.to(BF16), flat32, None, self, 1.0, 0.0, train)
        if train:
            self._ema_step(lambda idx: flat32 if idx is None else flat32[idx], flat32.shape[0])
        return xq.detach().reshape(inputs.shape), codes.reshape(inputs.shape[:-1])

    def embed(self, idxs):
        return super().forward(idxs)


class VectorQuantize2(nn.Module):
    """quantize2_mask.py:135-209."""

    def __init__(self, codebook_size, codebook_dim=None, accept_image_fmap=True, commitment_beta=0.25,
                 decay=0.99, restart_unused_codes=True, channel_last=False):
        super().__init__()
        self.accept_image_fmap = accept_image_fmap
        self.beta = commitment_beta
        self.channel_last = channel_last
        self.restart_unused_codes = restart_unused_codes
        self.codebook = VQEmbedding(codebook_size, codebook_dim, decay=decay,
                                    restart_unused_codes=restart_unused_codes)
        self.codebook.weight.data.uniform_(-1.0 / codebook_size, 1.0 / codebook_size)

    # ---- fused-model path: NHWC bf16 rows in, NHWC bf16 rows out
    def forward_rows(self, rows_bf16, row_mask):
        """rows_bf16 [N,C] (autograd-tracked), row_mask [N] fp32 or None -> (xq rows bf16, loss, codes)."""
        emb = self.codebook
        train = self.training and emb.ema
        if train:
            emb._ensure_acc()
        xq, loss, codes = _VQFn.apply(rows_bf16, None, row_mask, emb, 1.0 + self.beta, self.beta, train)
        if train:
            det = rows_bf16.detach()
            emb._ema_step(lambda idx: det if idx is None else det[idx], det.shape[0])
        return xq, loss, codes

    # ---- reference-facing path
    def forward(self, x, codebook_mask=None, *ignorewargs, **ignorekwargs):
        if not x.is_cuda:
            raise RuntimeError("VectorQuantize2 (B200) needs CUDA tensors; there is no CPU fallback")
        need_transpose = not self.channel_last and not self.accept_image_fmap
        if self.accept_image_fmap:
            b, c, height, width = x.shape
            rows = ops.ToNCHWInvFn.apply(x)                   # [B,H,W,C] fp32, differentiable
        else:
            if need_transpose:
                x = x.transpose(1, 2)
            rows = x
        shape = rows.shape
        c = shape[-1]
        flat32 = rows.reshape(-1, c).float().contiguous()
        mask_rows = None
        if codebook_mask is not None:
            mask_rows = codebook_mask.reshape(-1).float().contiguous()
            assert mask_rows.numel() == flat32.shape[0], "codebook_mask must have one value per position"
        emb = self.codebook
        train = self.training and emb.ema
        if train:
            emb._ensure_acc()
        xq, loss, codes = _VQFn.apply(flat32.detach().to(BF16), flat32, mask_rows, emb, 1.0 + self.beta, self.beta,
                                      train)
        if train:
            det = flat32.detach()
            emb._ema_step(lambda idx: det if idx is None else det[idx], det.shape[0])
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
        distances = self.codebook.compute_distances(x)
        soft_code = F.softmax(-distances / temp, dim=-1)
        if stochastic:
            flat = soft_code.reshape(-1, soft_code.shape[-1])
            code = torch.multinomial(flat, 1).reshape(*soft_code.shape[:-1])
        else:
            code = distances.argmin(dim=-1)
        return soft_code, code

    def get_codebook_entry(self, indices, *kwargs):
        return self.codebook.embed(indices)
