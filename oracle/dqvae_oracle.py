"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the convolutional part of the DQ-VAE stage-1 path.

A functional, plain-PyTorch fp32 restatement of the reference modules, driven by a flat
``state_dict`` with the reference's key names.  Each function cites the reference lines it follows
(paths relative to /root/reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.

Pinning: the reference has no tests; ``tests/golden/make_golden.py`` runs the reference's own
classes in the build container on weights produced by :func:`make_weights` and stores their
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against
those files.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- configs (configs/stage1/*.yml)
DUAL_CFG = dict(
    ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16, 32), in_channels=3,
    resolution=256, z_channels=256,                        # dqvae-dual-r-05_imagenet.yml:6-16
    dec_ch=128, dec_ch_mult=(1, 1, 2, 2), dec_attn_resolutions=(32,), out_ch=3, latent_size=32,
    codebook_size=1024, codebook_dim=256, beta=0.25, decay=0.99,
    router="feature",
)
# reduced-width variant used for fast CPU tests (same topology, 1/4 width, 64x64 images)
TINY_CFG = dict(
    ch=32, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(4, 8), in_channels=3,
    resolution=64, z_channels=64,
    dec_ch=32, dec_ch_mult=(1, 1, 2, 2), dec_attn_resolutions=(8,), out_ch=3, latent_size=8,
    codebook_size=128, codebook_dim=64, beta=0.25, decay=0.99,
    router="feature",
)


# 64-channel-granular variant (every conv input is a multiple of 64 channels, as the CUDA kernels
# require) used by the GPU parity tests: same topology, half width, 64x64 images
SMALL_CFG = dict(
    ch=64, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(4, 8), in_channels=3,
    resolution=64, z_channels=64,
    dec_ch=64, dec_ch_mult=(1, 1, 2, 2), dec_attn_resolutions=(8,), out_ch=3, latent_size=8,
    codebook_size=128, codebook_dim=64, beta=0.25, decay=0.99,
    router="feature",
)


TRIPLE_CFG = dict(
    ch=128, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolutions=(8, 16, 32), in_channels=3,
    resolution=256, z_channels=256,                        # dqvae-triple-r-03-03_imagenet.yml:6-16
    dec_ch=128, dec_ch_mult=(1, 1, 2, 2), dec_attn_resolutions=(32,), out_ch=3, latent_size=32,
    codebook_size=1024, codebook_dim=256, beta=0.25, decay=0.99,
    router="feature", grains=3,
)
TINY_TRIPLE_CFG = dict(
    ch=32, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolutions=(2, 4, 8), in_channels=3,
    resolution=64, z_channels=64,
    dec_ch=32, dec_ch_mult=(1, 1, 2, 2), dec_attn_resolutions=(8,), out_ch=3, latent_size=8,
    codebook_size=128, codebook_dim=64, beta=0.25, decay=0.99,
    router="feature", grains=3,
)
SMALL_TRIPLE_CFG = dict(
    ch=64, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolutions=(4, 8, 16), in_channels=3,
    resolution=128, z_channels=64,
    dec_ch=64, dec_ch_mult=(1, 1, 2, 2), dec_attn_resolutions=(16,), out_ch=3, latent_size=16,
    codebook_size=128, codebook_dim=64, beta=0.25, decay=0.99,
    router="feature", grains=3,
)
# entropy-routed dual model (dqvae-entropy-dual-r05_imagenet.yml): no router parameters
ENTROPY_CFG = dict(DUAL_CFG, router="entropy")
SMALL_ENTROPY_CFG = dict(SMALL_CFG, router="entropy")
TINY_ENTROPY_CFG = dict(TINY_CFG, router="entropy")


# --------------------------------------------------------------------------- parameter inventory
def _conv(shapes, p, cin, cout, k):
    shapes[p + ".weight"] = (cout, cin, k, k)
    shapes[p + ".bias"] = (cout,)


def _norm(shapes, p, c):
    shapes[p + ".weight"] = (c,)
    shapes[p + ".bias"] = (c,)


def _resblock(shapes, p, cin, cout):
    _norm(shapes, p + ".norm1", cin)
    _conv(shapes, p + ".conv1", cin, cout, 3)
    _norm(shapes, p + ".norm2", cout)
    _conv(shapes, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(shapes, p + ".nin_shortcut", cin, cout, 1)


def _attn(shapes, p, c):
    _norm(shapes, p + ".norm", c)
    for n in ("q", "k", "v", "proj_out"):
        _conv(shapes, f"{p}.{n}", c, c, 1)


def encoder_shapes(cfg, prefix="encoder"):
    """Parameter names/shapes of DualGrainEncoder (modules/dynamic_modules/EncoderDual.py:16-86)."""
    s = {}
    ch, mult = cfg["ch"], cfg["ch_mult"]
    _conv(s, f"{prefix}.conv_in", cfg["in_channels"], ch, 3)
    res = cfg["resolution"]
    in_mult = (1,) + tuple(mult)
    block_in = ch
    for lvl in range(len(mult)):
        block_in, block_out = ch * in_mult[lvl], ch * mult[lvl]
        for b in range(cfg["num_res_blocks"]):
            _resblock(s, f"{prefix}.down.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
            if res in cfg["attn_resolutions"]:
                _attn(s, f"{prefix}.down.{lvl}.attn.{b}", block_in)
        if lvl != len(mult) - 1:
            _conv(s, f"{prefix}.down.{lvl}.downsample.conv", block_in, block_in, 3)
            res //= 2
    if cfg.get("grains", 2) == 3:                         # EncoderTriple.py:63-93
        c_med = block_in // (mult[-1] // mult[-2])
        heads = (("coarse", block_in), ("median", c_med), ("fine", c_med // (mult[-2] // mult[-3])))
    else:
        heads = (("coarse", block_in), ("fine", block_in // (mult[-1] // mult[-2])))
    for grain, c in heads:
        _resblock(s, f"{prefix}.mid_{grain}.block_1", c, c)
        _attn(s, f"{prefix}.mid_{grain}.attn_1", c)
        _resblock(s, f"{prefix}.mid_{grain}.block_2", c, c)
        _norm(s, f"{prefix}.norm_out_{grain}", c)
        _conv(s, f"{prefix}.conv_out_{grain}", c, cfg["z_channels"], 3)
    if cfg["router"] == "feature":                       # RouterDual.py:7-32 / RouterTriple.py:6-44
        z, ng = cfg["z_channels"], len(heads)
        s[f"{prefix}.router.gate.0.weight"] = (ng * z, ng * z)
        s[f"{prefix}.router.gate.0.bias"] = (ng * z,)
        s[f"{prefix}.router.gate.2.weight"] = (ng, ng * z)
        s[f"{prefix}.router.gate.2.bias"] = (ng,)
        for grain, _ in heads:
            _norm(s, f"{prefix}.router.feature_norm_{grain}", z)
    return s


def decoder_shapes(cfg, prefix="decoder"):
    """Parameter names/shapes of DecoderPositional.Decoder (:42-107), position_type fourier+learned."""
    s = {}
    ch, mult = cfg["dec_ch"], cfg["dec_ch_mult"]
    nres = len(mult)
    block_in = ch * mult[-1]
    res = cfg["resolution"] // 2 ** (nres - 1)
    zc = cfg["z_channels"]
    _conv(s, f"{prefix}.conv_in", zc, block_in, 3)
    _resblock(s, f"{prefix}.mid.block_1", block_in, block_in)
    _attn(s, f"{prefix}.mid.attn_1", block_in)
    _resblock(s, f"{prefix}.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = ch * mult[lvl]
        for b in range(cfg["num_res_blocks"] + 1):
            _resblock(s, f"{prefix}.up.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
            if res in cfg["dec_attn_resolutions"]:
                _attn(s, f"{prefix}.up.{lvl}.attn.{b}", block_in)
        if lvl != 0:
            _conv(s, f"{prefix}.up.{lvl}.upsample.conv", block_in, block_in, 3)
            res *= 2
    _norm(s, f"{prefix}.norm_out", block_in)
    _conv(s, f"{prefix}.conv_out", block_in, cfg["out_ch"], 3)
    _conv(s, f"{prefix}.position_bias_fourier.lff.ffm.conv", 2, zc, 1)
    s[f"{prefix}.position_bias_learned.row_embed.weight"] = (cfg["latent_size"], zc)
    s[f"{prefix}.position_bias_learned.col_embed.weight"] = (cfg["latent_size"], zc)
    return s


def model_shapes(cfg):
    """All tensors of DualGrainVQModel minus the loss (models/stage1_dynamic/dqvae_dual_feat.py:26-35)."""
    s = {}
    s.update(encoder_shapes(cfg))
    s.update(decoder_shapes(cfg))
    K, C = cfg["codebook_size"], cfg["codebook_dim"]
    s["quantize.codebook.weight"] = (K + 1, C)
    s["quantize.codebook.cluster_size_ema"] = (K,)
    s["quantize.codebook.embed_ema"] = (K, C)
    _conv(s, "quant_conv", cfg["z_channels"], C, 1)
    _conv(s, "post_quant_conv", C, cfg["z_channels"], 1)
    return s


def make_weights(shapes, seed=0):
    """Deterministic weights that do not depend on module construction order: every tensor gets its
    own generator seeded by (seed, crc32(name)).  Convs/linears ~ U(+-1/sqrt(fan_in)), norm scales
    ~ 1 + 0.1 N(0,1), biases 0.05 N(0,1), codebook rows ~ N(0,1) (data scale)."""
    import zlib
    sd = {}
    for name in sorted(shapes):
        shape = shapes[name]
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if name.endswith("cluster_size_ema"):
            t = torch.ones(shape)
        elif "codebook" in name:
            t = torch.randn(shape, generator=g)
        elif "embed.weight" in name:
            t = torch.randn(shape, generator=g) * 0.5
        elif len(shape) == 1:
            is_scale = name.endswith(".weight")
            t = torch.randn(shape, generator=g) * (0.1 if is_scale else 0.05) + (1.0 if is_scale else 0.0)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[name] = t.float()
    return sd


# --------------------------------------------------------------------------- building blocks
def swish(x):
    """modules/diffusionmodules/model.py:29-31."""
    return x * torch.sigmoid(x)


def group_norm(sd, p, x):
    """model.py:34-35: GroupNorm(32, C, eps=1e-6, affine)."""
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def conv2d(sd, p, x, stride=1, padding=None):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd[p + ".bias"], stride=stride,
                    padding=w.shape[-1] // 2 if padding is None else padding)


def resnet_block(sd, p, x):
    """model.py:117-137 with temb=None, dropout p=0."""
    h = conv2d(sd, p + ".conv1", swish(group_norm(sd, p + ".norm1", x)))
    h = conv2d(sd, p + ".conv2", swish(group_norm(sd, p + ".norm2", h)))
    if p + ".nin_shortcut.weight" in sd:
        x = conv2d(sd, p + ".nin_shortcut", x)
    return x + h


def attn_block(sd, p, x):
    """model.py:168-192: single-head attention over the h*w positions, scale C^-1/2."""
    h = group_norm(sd, p + ".norm", x)
    q, k, v = (conv2d(sd, f"{p}.{n}", h) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).transpose(1, 2)           # [b, T, c]
    k = k.reshape(b, c, hh * ww)                           # [b, c, T]
    w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    o = torch.bmm(v.reshape(b, c, hh * ww), w.transpose(1, 2)).reshape(b, c, hh, ww)
    return x + conv2d(sd, p + ".proj_out", o)


def downsample(sd, p, x):
    """model.py:68-72: zero-pad right/bottom by 1, 3x3 stride 2."""
    return conv2d(sd, p + ".conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)


def upsample(sd, p, x):
    """model.py:49-53: nearest x2 then 3x3 conv."""
    return conv2d(sd, p + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))


# --------------------------------------------------------------------------- encoder / router / decoder
def feature_router(sd, p, h_fine, h_coarse):
    """RouterDual.py:35-43."""
    hf = F.group_norm(h_fine, 32, sd[p + ".feature_norm_fine.weight"], sd[p + ".feature_norm_fine.bias"], eps=1e-6)
    hc = F.group_norm(h_coarse, 32, sd[p + ".feature_norm_coarse.weight"], sd[p + ".feature_norm_coarse.bias"], eps=1e-6)
    z = torch.cat([hc, F.avg_pool2d(hf, 2, 2)], dim=1).permute(0, 2, 3, 1)
    z = F.silu(F.linear(z, sd[p + ".gate.0.weight"], sd[p + ".gate.0.bias"]))
    return F.linear(z, sd[p + ".gate.2.weight"], sd[p + ".gate.2.bias"])


def entropy_router(entropy, threshold):
    """RouterDual.py:53-57."""
    fine = (entropy > threshold).long().unsqueeze(-1)
    coarse = (entropy <= threshold).long().unsqueeze(-1)
    return torch.cat([coarse, fine], dim=-1)


def patch_entropy(x, patch=16, lo=-1.0):
    """models/stage1_dynamic/dqvae_dual_entropy.py:25-63 (32 bins on [-1,1], sigma 0.01, eps 1e-40).  lo=0.0 gives
    the variant of scripts/tools/calculate_entropy_thresholds.py:65-79, whose bins span [0,1]."""
    b = x.shape[0]
    gray = 0.2989 * x[:, 0:1] + 0.5870 * x[:, 1:2] + 0.1140 * x[:, 2:]
    u = F.unfold(gray, kernel_size=patch, stride=patch).transpose(1, 2)      # [b, P, patch*patch]
    nper = u.shape[1]
    u = u.reshape(b * nper, -1)
    bins = torch.linspace(lo, 1, 32, device=x.device)
    k = torch.exp(-0.5 * ((u.unsqueeze(2) - bins.view(1, 1, -1)) / torch.tensor(0.01)).pow(2))
    pdf = k.mean(dim=1)
    pdf = pdf / (pdf.sum(dim=1, keepdim=True) + 1e-40) + 1e-40
    ent = -(pdf * torch.log(pdf)).sum(dim=1)
    hw = x.shape[-1] // patch
    return ent.reshape(b, hw, hw)


def entropy_thresholds(batches, patch=16):
    """scripts/tools/calculate_entropy_thresholds.py:95-117: patch entropies (bins on [0,1]) of every batch, sorted;
    threshold "i" (i = 1..99) = sorted[(size * i) // 100]."""
    ent = np.sort(np.concatenate([patch_entropy(b, patch, lo=0.0).reshape(-1).numpy() for b in batches]))
    size = ent.shape[0]
    return {str(i + 1): float(ent[int((size * (i + 1)) // 100)]) for i in range(99)}


def gumbel_softmax_hard(logits, noise, forced_index=None):
    """torch.nn.functional.gumbel_softmax(logits, tau=1, hard=True, dim=-1) with the Gumbel noise given
    (noise = -log(Exp(1)) drawn by the caller; EncoderDual.py:132-133, EncoderTriple.py:157-158): the value is
    the one-hot of argmax(softmax(logits + noise)), the gradient is that of the soft sample.
    forced_index replays the arg-max of another run (teacher forcing across a near-tie)."""
    y_soft = torch.softmax(logits + noise, dim=-1)
    index = y_soft.argmax(dim=-1, keepdim=True) if forced_index is None else forced_index.unsqueeze(-1)
    y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
    return y_hard - y_soft.detach() + y_soft


def dual_encoder(sd, cfg, x, x_entropy=None, forced_gate=None, entropy_threshold=None, p="encoder",
                 gumbel_noise=None, forced_index=None):
    """EncoderDual.py:89-156.  Default = eval mode (no gumbel noise; `forced_gate` [B,h,w,2] overrides the
    router output).  `gumbel_noise` [B,h,w,2] switches to the TRAINING-mode routing of a feature router
    (update_router and self.training, :132-133 and :142-145): hard gumbel-softmax sample with that noise, and
    h_dual multiplied by gate.max(dim=1) - value 1, but it carries the gradient that trains the router."""
    nlev = len(cfg["ch_mult"])
    res = cfg["resolution"]
    h = conv2d(sd, p + ".conv_in", x)
    h_fine = None
    for lvl in range(nlev):
        for b in range(cfg["num_res_blocks"]):
            h = resnet_block(sd, f"{p}.down.{lvl}.block.{b}", h)
            if res in cfg["attn_resolutions"]:
                h = attn_block(sd, f"{p}.down.{lvl}.attn.{b}", h)
        if lvl == nlev - 2:
            h_fine = h
        if lvl != nlev - 1:
            h = downsample(sd, f"{p}.down.{lvl}.downsample", h)
            res //= 2
    heads = {}
    for grain, t in (("coarse", h), ("fine", h_fine)):
        t = resnet_block(sd, f"{p}.mid_{grain}.block_1", t)
        t = attn_block(sd, f"{p}.mid_{grain}.attn_1", t)
        t = resnet_block(sd, f"{p}.mid_{grain}.block_2", t)
        t = swish(group_norm(sd, f"{p}.norm_out_{grain}", t))
        heads[grain] = conv2d(sd, f"{p}.conv_out_{grain}", t)
    h_coarse, h_fine = heads["coarse"], heads["fine"]
    if forced_gate is not None:
        gate = forced_gate
    elif cfg["router"] == "feature":
        gate = feature_router(sd, p + ".router", h_fine, h_coarse)
    else:
        gate = entropy_router(x_entropy, entropy_threshold)
    if gumbel_noise is not None:
        gate = gumbel_softmax_hard(gate, gumbel_noise, forced_index)
    gate = gate.permute(0, 3, 1, 2)
    indices = gate.argmax(dim=1)
    up = h_coarse.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
    idx_rep = indices.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2).unsqueeze(1)
    h_dual = torch.where(idx_rep == 0, up, h_fine)
    if gumbel_noise is not None:
        gate_grad = gate.max(dim=1, keepdim=True)[0]
        h_dual = h_dual * gate_grad.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
    mask = torch.where(idx_rep == 0, torch.full_like(idx_rep, 0.25, dtype=torch.float32),
                       torch.ones_like(idx_rep, dtype=torch.float32))
    return dict(h_dual=h_dual, indices=indices, codebook_mask=mask, gate=gate,
                h_fine=h_fine, h_coarse=h_coarse)


def triple_router(sd, p, h_fine, h_median, h_coarse):
    """RouterTriple.py:46-55."""
    def gn(t, n):
        return F.group_norm(t, 32, sd[f"{p}.feature_norm_{n}.weight"], sd[f"{p}.feature_norm_{n}.bias"], eps=1e-6)
    z = torch.cat([gn(h_coarse, "coarse"), F.avg_pool2d(gn(h_median, "median"), 2, 2),
                   F.avg_pool2d(gn(h_fine, "fine"), 4, 4)], dim=1).permute(0, 2, 3, 1)
    z = F.silu(F.linear(z, sd[p + ".gate.0.weight"], sd[p + ".gate.0.bias"]))
    return F.linear(z, sd[p + ".gate.2.weight"], sd[p + ".gate.2.bias"])


def triple_encoder(sd, cfg, x, forced_gate=None, p="encoder"):
    """EncoderTriple.py:95-183 in eval mode (0 coarse / 1 median / 2 fine; masks 1/16, 1/4, 1)."""
    nlev = len(cfg["ch_mult"])
    res = cfg["resolution"]
    h = conv2d(sd, p + ".conv_in", x)
    taps = {}
    for lvl in range(nlev):
        for b in range(cfg["num_res_blocks"]):
            h = resnet_block(sd, f"{p}.down.{lvl}.block.{b}", h)
            if res in cfg["attn_resolutions"]:
                h = attn_block(sd, f"{p}.down.{lvl}.attn.{b}", h)
        taps[lvl] = h
        if lvl != nlev - 1:
            h = downsample(sd, f"{p}.down.{lvl}.downsample", h)
            res //= 2
    heads = {}
    for grain, t in (("coarse", taps[nlev - 1]), ("median", taps[nlev - 2]), ("fine", taps[nlev - 3])):
        t = resnet_block(sd, f"{p}.mid_{grain}.block_1", t)
        t = attn_block(sd, f"{p}.mid_{grain}.attn_1", t)
        t = resnet_block(sd, f"{p}.mid_{grain}.block_2", t)
        heads[grain] = conv2d(sd, f"{p}.conv_out_{grain}", swish(group_norm(sd, f"{p}.norm_out_{grain}", t)))
    gate = forced_gate if forced_gate is not None else triple_router(
        sd, p + ".router", heads["fine"], heads["median"], heads["coarse"])
    gate = gate.permute(0, 3, 1, 2)
    indices = gate.argmax(dim=1)
    up_c = heads["coarse"].repeat_interleave(4, dim=-1).repeat_interleave(4, dim=-2)
    up_m = heads["median"].repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
    idx = indices.repeat_interleave(4, dim=-1).repeat_interleave(4, dim=-2).unsqueeze(1)
    h_triple = torch.where(idx == 0, up_c, torch.where(idx == 1, up_m, heads["fine"]))
    one = torch.ones((), dtype=torch.float32)
    mask = torch.where(idx == 0, 0.0625 * one, torch.where(idx == 1, 0.25 * one, one))
    return dict(h_dual=h_triple, indices=indices, codebook_mask=mask, gate=gate)


def budget_loss_triple(gate, target_fine=0.3, target_median=0.3, gamma=1.0, min_grain=8, median_grain=16,
                       max_grain=32):
    """modules/dynamic_modules/budget.py:43-59."""
    n = gate.size(0)
    med = (gate[:, 0] + 4.0 * gate[:, 1] + gate[:, 2]).sum() / n - min_grain ** 2
    r_med = med / (median_grain ** 2 - min_grain ** 2)
    fine = (gate[:, 0] + 16.0 * gate[:, 2] + gate[:, 1]).sum() / n - min_grain ** 2
    r_fine = fine / (max_grain ** 2 - min_grain ** 2)
    return gamma * F.mse_loss(r_fine, torch.full_like(r_fine, target_fine)) + \
        F.mse_loss(r_med, torch.full_like(r_med, target_median))


def position_bias(sd, cfg, p="decoder"):
    """fourier_embedding.py:5-55 (linspace coords, sin(conv1x1)) + DecoderPositional.py:27-39."""
    n = cfg["latent_size"]
    lin = torch.linspace(-1, 1, n)
    coord = torch.stack([lin.view(1, n).expand(n, n), lin.view(n, 1).expand(n, n)], 0).unsqueeze(0)
    four = torch.sin(conv2d(sd, p + ".position_bias_fourier.lff.ffm.conv", coord))
    row = sd[p + ".position_bias_learned.row_embed.weight"]      # indexed by h
    col = sd[p + ".position_bias_learned.col_embed.weight"]      # indexed by w
    learned = (col.unsqueeze(0) + row.unsqueeze(1)).permute(2, 0, 1).unsqueeze(0)
    return four + learned


def decoder(sd, cfg, z, p="decoder"):
    """DecoderPositional.py:109-145 (position_type fourier+learned; grain_indices is unused there)."""
    nres = len(cfg["dec_ch_mult"])
    res = cfg["resolution"] // 2 ** (nres - 1)
    h = conv2d(sd, p + ".conv_in", z + position_bias(sd, cfg, p))
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for lvl in reversed(range(nres)):
        for b in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(sd, f"{p}.up.{lvl}.block.{b}", h)
            if res in cfg["dec_attn_resolutions"]:
                h = attn_block(sd, f"{p}.up.{lvl}.attn.{b}", h)
        if lvl != 0:
            h = upsample(sd, f"{p}.up.{lvl}.upsample", h)
            res *= 2
    return conv2d(sd, p + ".conv_out", swish(group_norm(sd, p + ".norm_out", h)))


# --------------------------------------------------------------------------- VQ (torch version, differentiable)
def vq_forward(sd, cfg, h, mask, search_bf16=False, p="quantize.codebook", forced_codes=None):
    """quantize2_mask.py:157-191 in eval mode.  search_bf16 evaluates the nearest-code search on
    bf16-rounded operands (what the CUDA kernel multiplies); everything else stays fp32."""
    b, c, hh, ww = h.shape
    x = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    w = sd[p + ".weight"]
    xs, cb = (x.detach(), w[:-1])
    if search_bf16:
        xs, cb = xs.bfloat16().float(), cb.bfloat16().float()
    flat = xs.reshape(-1, c)
    d = (flat.pow(2).sum(1, keepdim=True) + cb.t().pow(2).sum(0, keepdim=True)) - 2.0 * flat @ cb.t()
    codes = d.argmin(-1).reshape(b, hh * ww)
    if forced_codes is not None:          # teacher forcing: replay the codes of another run
        codes = forced_codes.reshape(b, hh * ww)
    xq = w[codes]
    m = mask.permute(0, 2, 3, 1).reshape(b, hh * ww, 1)
    loss = cfg["beta"] * torch.mean((xq.detach() - x) ** 2 * m) + torch.mean((xq - x.detach()) ** 2 * m)
    xq = x + (xq - x).detach()
    return xq.reshape(b, hh, ww, c).permute(0, 3, 1, 2), loss, codes.reshape(b, hh, ww)


def budget_loss_dual(gate, target_ratio=0.5, gamma=10.0, min_grain=16, max_grain=32):
    """modules/dynamic_modules/budget.py:15-28 with calculate_all=True (returns 2*gamma*MSE(1-r,1-t))."""
    beta = (1.0 * gate[:, 0] + 4.0 * gate[:, 1]).sum() / gate.size(0) - min_grain ** 2
    ratio = beta / (max_grain ** 2 - min_grain ** 2)
    last = gamma * F.mse_loss(1 - ratio, 1 - torch.full_like(ratio, target_ratio))
    return last + last


def model_forward(sd, cfg, x, search_bf16=False, forced_gate=None, x_entropy=None, entropy_threshold=None,
                  forced_codes=None, gumbel_noise=None, forced_index=None):
    """models/stage1_dynamic/dqvae_dual_feat.py:59-78: encode -> quant_conv -> VQ -> post_quant_conv
    -> decode.  Returns dict(xrec, qloss, codes, indices, gate, h_dual)."""
    if cfg.get("grains", 2) == 3:
        enc = triple_encoder(sd, cfg, x, forced_gate=forced_gate)
    else:
        if cfg["router"] == "entropy" and x_entropy is None and forced_gate is None:
            x_entropy = patch_entropy(x, patch=cfg["resolution"] // (cfg["latent_size"] // 2))
        enc = dual_encoder(sd, cfg, x, x_entropy=x_entropy, forced_gate=forced_gate,
                           entropy_threshold=entropy_threshold, gumbel_noise=gumbel_noise, forced_index=forced_index)
    h = conv2d(sd, "quant_conv", enc["h_dual"])
    quant, qloss, codes = vq_forward(sd, cfg, h, enc["codebook_mask"], search_bf16=search_bf16,
                                     forced_codes=forced_codes)
    xrec = decoder(sd, cfg, conv2d(sd, "post_quant_conv", quant))
    return dict(xrec=xrec, qloss=qloss, codes=codes, indices=enc["indices"], gate=enc["gate"],
                h_dual=enc["h_dual"], h_pre_vq=h)
