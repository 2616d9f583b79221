"""Mint golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference

Imports the reference's own classes from /root/reference (with a 3-line pytorch_lightning shim,
SURVEY.md 8c), loads deterministic weights from oracle.dqvae_oracle.make_weights into them, runs
them on seeded inputs (CPU, fp32) and stores the results as small .npz fixtures next to this
file.  /root/reference does not exist on the GPU box: tests only read the .npz files.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
os.chdir(REF)  # the reference resolves relative paths (thresholds json) from its root

pl = types.ModuleType("pytorch_lightning")
pl.LightningModule = nn.Module
sys.modules["pytorch_lightning"] = pl

from oracle import dqvae_oracle as orc  # noqa: E402

from modules.vector_quantization.quantize2_mask import VectorQuantize2  # noqa: E402
from utils.utils import instantiate_from_config  # noqa: E402


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


# ----------------------------------------------------------------------------------- VQ goldens
def vq_goldens():
    torch.manual_seed(1234)
    K, C, B, H = 64, 64, 2, 8
    vq = VectorQuantize2(codebook_size=K, codebook_dim=C)
    g = torch.Generator().manual_seed(5)
    w = torch.randn(K + 1, C, generator=g)
    with torch.no_grad():
        vq.codebook.weight.copy_(w)
        vq.codebook.embed_ema.copy_(w[:-1])
        vq.codebook.cluster_size_ema.fill_(1.0)
    x = torch.randn(B, C, H, H, generator=g)
    mask = torch.where(torch.rand(B, 1, H, H, generator=g) > 0.5, 1.0, 0.25)
    # eval forward + backward
    vq.eval()
    xin = x.clone().requires_grad_(True)
    xq, loss, (_, _, codes) = vq(xin, codebook_mask=mask)
    gq = torch.randn(xq.shape, generator=g)
    (xq * gq).sum().backward(retain_graph=True)
    gx_ste = xin.grad.clone()
    xin.grad = None
    loss.backward()
    # three train steps with the restart rows replayed: make randperm deterministic by seeding
    vq.train()
    states = []
    xs = []
    for step in range(3):
        xt = torch.randn(B, C, H, H, generator=g) * (1 + step)
        torch.manual_seed(100 + step)
        perm = torch.randperm(B * H * H)  # what _update_buffers will draw after manual_seed
        torch.manual_seed(100 + step)
        xq_t, loss_t, (_, _, codes_t) = vq(xt, codebook_mask=mask)
        flat = xt.permute(0, 2, 3, 1).reshape(-1, C)
        states.append(dict(codes=codes_t.clone(), xq=xq_t.detach().clone(), loss=loss_t.detach().clone(),
                           cs=vq.codebook.cluster_size_ema.clone(), em=vq.codebook.embed_ema.clone(),
                           w=vq.codebook.weight.detach().clone(), restart=flat[perm][:K].clone()))
        xs.append(xt)
    save("vq_small.npz", weight=w, x=x, mask=mask, xq=xq, loss=loss, codes=codes, gq=gq, gx_ste=gx_ste,
         gx_loss=xin.grad,
         **{f"t{i}_{k}": v for i, s in enumerate(states) for k, v in s.items()},
         **{f"t{i}_x": v for i, v in enumerate(xs)})


# ----------------------------------------------------------------------------------- model goldens
def build_reference_modules(cfg, yaml_name="dqvae-dual-r-05_imagenet.yml"):
    conf = yaml.safe_load(open(os.path.join(REF, "configs/stage1", yaml_name)))["model"]["params"]
    enc_p, dec_p = dict(conf["encoderconfig"]["params"]), dict(conf["decoderconfig"]["params"])
    enc_p.update(ch=cfg["ch"], resolution=cfg["resolution"], z_channels=cfg["z_channels"],
                 attn_resolutions=list(cfg["attn_resolutions"]))
    enc_p["router_config"] = dict(enc_p["router_config"])
    enc_p["router_config"]["params"] = dict(enc_p["router_config"]["params"], num_channels=cfg["z_channels"])
    dec_p.update(ch=cfg["dec_ch"], in_ch=cfg["z_channels"], resolution=cfg["resolution"],
                 attn_resolutions=list(cfg["dec_attn_resolutions"]), latent_size=cfg["latent_size"])
    enc = instantiate_from_config(dict(target=conf["encoderconfig"]["target"], params=enc_p))
    dec = instantiate_from_config(dict(target=conf["decoderconfig"]["target"], params=dec_p))
    vq_p = dict(conf["vqconfig"]["params"], codebook_size=cfg["codebook_size"], codebook_dim=cfg["codebook_dim"])
    vq = instantiate_from_config(dict(target=conf["vqconfig"]["target"], params=vq_p))
    m = nn.Module()
    m.encoder, m.decoder, m.quantize = enc, dec, vq
    m.quant_conv = nn.Conv2d(cfg["z_channels"], cfg["codebook_dim"], 1)
    m.post_quant_conv = nn.Conv2d(cfg["codebook_dim"], cfg["z_channels"], 1)
    return m


def model_goldens(cfg, tag, batch, seed):
    m = build_reference_modules(cfg)
    shapes = orc.model_shapes(cfg)
    ref_sd = m.state_dict()
    assert set(ref_sd) == set(shapes), (sorted(set(ref_sd) ^ set(shapes)))
    for k in shapes:
        assert tuple(ref_sd[k].shape) == tuple(shapes[k]), (k, ref_sd[k].shape, shapes[k])
    sd = orc.make_weights(shapes, seed=seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(seed + 77)
    x = torch.rand(batch, 3, cfg["resolution"], cfg["resolution"], generator=g) * 2 - 1
    h_dict = m.encoder(x, None)
    h = m.quant_conv(h_dict["h_dual"])
    quant, qloss, (_, _, codes) = m.quantize(x=h, temp=0.0, codebook_mask=h_dict["codebook_mask"])
    xrec = m.decoder(m.post_quant_conv(quant), h_dict["indices"])
    loss = (xrec - x).abs().mean() + qloss
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    pick = ["encoder.conv_in.weight", "encoder.down.0.block.0.conv1.weight", "encoder.down.0.block.0.norm1.weight",
            "encoder.down.3.attn.0.q.weight", "encoder.down.2.block.0.nin_shortcut.weight",
            "encoder.down.1.downsample.conv.weight", "encoder.conv_out_fine.bias",
            "quant_conv.weight", "post_quant_conv.weight", "decoder.conv_in.weight",
            "decoder.up.3.attn.1.proj_out.weight", "decoder.up.1.upsample.conv.weight",
            "decoder.up.0.block.2.conv2.weight", "decoder.norm_out.weight", "decoder.conv_out.weight",
            "decoder.conv_out.bias", "decoder.position_bias_learned.row_embed.weight",
            "decoder.position_bias_fourier.lff.ffm.conv.weight"]
    gsel = {"grad__" + k.replace(".", "__"): grads[k] for k in pick}
    gnorm = {k: float(v.double().pow(2).sum().sqrt()) for k, v in grads.items()}
    names = sorted(gnorm)
    save(f"model_{tag}.npz", x=x, xrec=xrec, qloss=qloss, codes=codes.to(torch.int16), indices=h_dict["indices"].to(torch.int8),
         gate=h_dict["gate"], h_dual=h_dict["h_dual"].to(torch.float16) if tag != "tiny" else h_dict["h_dual"],
         loss=loss, grad_norm_names=np.array(names), grad_norms=np.array([gnorm[n] for n in names]), **gsel)


def train_routing_goldens():
    """Training-mode routing of the reference's DualGrainEncoder (EncoderDual.py:130-149): hard gumbel-softmax
    sample of the router logits, h_dual multiplied by gate.max(dim=1) (the path that trains the router), dual
    budget loss on the sampled gate.  The Gumbel noise F.gumbel_softmax draws is the first RNG use of the forward
    (dropout p = 0 draws nothing), so it is replayed by drawing it here after the same seed."""
    from modules.dynamic_modules.budget import BudgetConstraint_RatioMSE_DualGrain
    cfg = orc.TINY_CFG
    m = build_reference_modules(cfg)
    m.load_state_dict(orc.make_weights(orc.model_shapes(cfg), seed=9), strict=True)
    m.train()
    g = torch.Generator().manual_seed(91)
    x = torch.rand(2, 3, cfg["resolution"], cfg["resolution"], generator=g) * 2 - 1
    lat = cfg["latent_size"]
    probe = torch.randn(2, cfg["z_channels"], lat, lat, generator=g)           # cotangent of h_dual
    torch.manual_seed(777)
    noise = -torch.empty(2, lat // 2, lat // 2, 2).exponential_().log()
    torch.manual_seed(777)
    hd = m.encoder(x, None)
    budget = BudgetConstraint_RatioMSE_DualGrain(target_ratio=0.5, gamma=10.0, min_grain_size=lat // 2,
                                                 max_grain_size=lat, calculate_all=True)(hd["gate"])
    ((hd["h_dual"] * probe).sum() + budget).backward()
    grads = {n: p.grad for n, p in m.encoder.named_parameters() if p.grad is not None}
    pick = ["router.gate.0.weight", "router.gate.2.weight", "router.gate.2.bias", "router.feature_norm_fine.weight",
            "conv_out_fine.weight", "conv_out_coarse.bias", "down.0.block.0.conv1.weight", "conv_in.weight"]
    save("model_tiny_train_routing.npz", x=x, probe=probe, noise=noise, gate=hd["gate"], indices=hd["indices"],
         h_dual=hd["h_dual"], mask=hd["codebook_mask"], budget=budget,
         grad_norm_names=np.array(sorted(grads)),
         grad_norms=np.array([float(grads[n].double().pow(2).sum().sqrt()) for n in sorted(grads)]),
         **{"grad__" + k.replace(".", "__"): grads[k] for k in pick})


def triple_entropy_goldens():
    """Tiny triple-grain model (EncoderTriple + RouterTriple + triple budget loss) and the entropy
    branch (Entropy module + fixed-threshold router) of the reference, eval mode."""
    import json
    import tempfile
    from modules.dynamic_modules.budget import (BudgetConstraint_NormedSeperateRatioMSE_TripleGrain,
                                                BudgetConstraint_RatioMSE_DualGrain)
    conf = yaml.safe_load(open(os.path.join(REF, "configs/stage1/dqvae-triple-r-03-03_imagenet.yml")))["model"]["params"]
    cfg = orc.TINY_TRIPLE_CFG
    enc_p = dict(conf["encoderconfig"]["params"], ch=cfg["ch"], resolution=cfg["resolution"],
                 z_channels=cfg["z_channels"], attn_resolutions=list(cfg["attn_resolutions"]))
    enc_p["router_config"] = dict(enc_p["router_config"], params=dict(enc_p["router_config"]["params"],
                                                                      num_channels=cfg["z_channels"]))
    dec_p = dict(conf["decoderconfig"]["params"], ch=cfg["dec_ch"], in_ch=cfg["z_channels"],
                 resolution=cfg["resolution"], attn_resolutions=list(cfg["dec_attn_resolutions"]),
                 latent_size=cfg["latent_size"])
    m = nn.Module()
    m.encoder = instantiate_from_config(dict(target=conf["encoderconfig"]["target"], params=enc_p))
    m.decoder = instantiate_from_config(dict(target=conf["decoderconfig"]["target"], params=dec_p))
    m.quantize = instantiate_from_config(dict(target=conf["vqconfig"]["target"], params=dict(
        conf["vqconfig"]["params"], codebook_size=cfg["codebook_size"], codebook_dim=cfg["codebook_dim"])))
    m.quant_conv = nn.Conv2d(cfg["z_channels"], cfg["codebook_dim"], 1)
    m.post_quant_conv = nn.Conv2d(cfg["codebook_dim"], cfg["z_channels"], 1)
    shapes = orc.model_shapes(cfg)
    ref_sd = m.state_dict()
    assert set(ref_sd) == set(shapes), sorted(set(ref_sd) ^ set(shapes))
    assert all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)
    m.load_state_dict(orc.make_weights(shapes, seed=5), strict=True)
    m.eval()
    g = torch.Generator().manual_seed(55)
    x = torch.rand(2, 3, cfg["resolution"], cfg["resolution"], generator=g) * 2 - 1
    hd = m.encoder(x, None)
    h = m.quant_conv(hd["h_triple"])
    quant, qloss, (_, _, codes) = m.quantize(x=h, temp=0.0, codebook_mask=hd["codebook_mask"])
    xrec = m.decoder(m.post_quant_conv(quant), None)
    budget = BudgetConstraint_NormedSeperateRatioMSE_TripleGrain(
        target_fine_ratio=0.3, target_median_ratio=0.3, gamma=1.0, min_grain_size=2, median_grain_size=4,
        max_grain_size=8)(hd["gate"])
    save("model_tiny_triple.npz", x=x, xrec=xrec, qloss=qloss, codes=codes.to(torch.int16),
         indices=hd["indices"].to(torch.int8), gate=hd["gate"], mask=hd["codebook_mask"], budget=budget)

    # entropy branch: Entropy module + fixed-threshold router + dual budget loss
    from models.stage1_dynamic.dqvae_dual_entropy import Entropy
    from modules.dynamic_modules.RouterDual import DualGrainFixedEntropyRouter
    xe = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    xe[:, :, :32] = xe[:, :, :32].mean(dim=(2, 3), keepdim=True) + 0.02 * xe[:, :, :32]   # flat top half
    ent = Entropy(16, 64, 64)(xe)
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump({"50": 1.5}, f)
    router = DualGrainFixedEntropyRouter(json_path=f.name, fine_grain_ratito=0.5)
    gate_e = router(entropy=ent)
    bud = BudgetConstraint_RatioMSE_DualGrain(target_ratio=0.5, gamma=10.0, min_grain_size=4, max_grain_size=8,
                                              calculate_all=True)(gate_e.permute(0, 3, 1, 2).float())
    save("entropy_small.npz", x=xe, entropy=ent, gate=gate_e, threshold=1.5, budget=bud)


def triple_backward_goldens():
    """Gradients of the tiny triple-grain model of the reference (EncoderTriple + RouterTriple, eval-mode routing) under
    |xrec - x|.mean() + qloss + triple budget loss: norms of every parameter gradient and a few full tensors
    (SURVEY 8 rows a3 + a18).  Same construction and seeds as triple_entropy_goldens."""
    from modules.dynamic_modules.budget import BudgetConstraint_NormedSeperateRatioMSE_TripleGrain
    conf = yaml.safe_load(open(os.path.join(REF, "configs/stage1/dqvae-triple-r-03-03_imagenet.yml")))["model"]["params"]
    cfg = orc.TINY_TRIPLE_CFG
    enc_p = dict(conf["encoderconfig"]["params"], ch=cfg["ch"], resolution=cfg["resolution"],
                 z_channels=cfg["z_channels"], attn_resolutions=list(cfg["attn_resolutions"]))
    enc_p["router_config"] = dict(enc_p["router_config"], params=dict(enc_p["router_config"]["params"],
                                                                      num_channels=cfg["z_channels"]))
    dec_p = dict(conf["decoderconfig"]["params"], ch=cfg["dec_ch"], in_ch=cfg["z_channels"],
                 resolution=cfg["resolution"], attn_resolutions=list(cfg["dec_attn_resolutions"]),
                 latent_size=cfg["latent_size"])
    m = nn.Module()
    m.encoder = instantiate_from_config(dict(target=conf["encoderconfig"]["target"], params=enc_p))
    m.decoder = instantiate_from_config(dict(target=conf["decoderconfig"]["target"], params=dec_p))
    m.quantize = instantiate_from_config(dict(target=conf["vqconfig"]["target"], params=dict(
        conf["vqconfig"]["params"], codebook_size=cfg["codebook_size"], codebook_dim=cfg["codebook_dim"])))
    m.quant_conv = nn.Conv2d(cfg["z_channels"], cfg["codebook_dim"], 1)
    m.post_quant_conv = nn.Conv2d(cfg["codebook_dim"], cfg["z_channels"], 1)
    m.load_state_dict(orc.make_weights(orc.model_shapes(cfg), seed=5), strict=True)
    m.eval()
    g = torch.Generator().manual_seed(55)
    x = torch.rand(2, 3, cfg["resolution"], cfg["resolution"], generator=g) * 2 - 1
    hd = m.encoder(x, None)
    h = m.quant_conv(hd["h_triple"])
    quant, qloss, (_, _, codes) = m.quantize(x=h, temp=0.0, codebook_mask=hd["codebook_mask"])
    xrec = m.decoder(m.post_quant_conv(quant), None)
    budget = BudgetConstraint_NormedSeperateRatioMSE_TripleGrain(
        target_fine_ratio=0.3, target_median_ratio=0.3, gamma=1.0, min_grain_size=2, median_grain_size=4,
        max_grain_size=8)(hd["gate"])
    loss = (xrec - x).abs().mean() + qloss + budget
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    pick = [k for k in ("encoder.conv_in.weight", "encoder.conv_out_fine.weight", "encoder.conv_out_median.weight",
                        "encoder.conv_out_coarse.bias", "encoder.norm_out_median.weight", "quant_conv.weight",
                        "post_quant_conv.bias", "decoder.conv_in.weight", "decoder.norm_out.weight",
                        "decoder.conv_out.weight") if k in grads]
    names = sorted(grads)
    save("model_tiny_triple_grads.npz", x=x, loss=loss, indices=hd["indices"].to(torch.int8), codes=codes.to(torch.int16),
         grad_norm_names=np.array(names),
         grad_norms=np.array([float(grads[n].double().pow(2).sum().sqrt()) for n in names]),
         **{"grad__" + k.replace(".", "__"): grads[k] for k in pick})


# ----------------------------------------------------------------------------------- sibling quantizers
def vq_family_goldens():
    """quantize2 / quantize2_list / quantize_rqvae / quantize_vqgan of the reference.
    Training passes store the restart rows each update drew (same replay trick as vq_goldens)."""
    from modules.vector_quantization.quantize2 import VectorQuantize2 as VQ2
    from modules.vector_quantization.quantize2_list import VectorQuantize2 as VQ2List
    from modules.vector_quantization.quantize_rqvae import RQBottleneck
    from modules.vector_quantization.quantize_vqgan import VectorQuantizer2
    out = {}
    g = torch.Generator().manual_seed(77)

    def put(prefix, **kw):
        for k, v in kw.items():
            out[f"{prefix}_{k}"] = v.detach().clone() if torch.is_tensor(v) else v

    # ---- quantize2.VectorQuantize2, both loss flavours
    K, C = 64, 64
    w = torch.randn(K + 1, C, generator=g)
    x = torch.randn(2, C, 8, 8, generator=g)
    gq = torch.randn(2, C, 8, 8, generator=g)
    put("q2", weight=w, x=x, gq=gq)
    for legacy in (True, False):
        vq = VQ2(codebook_size=K, codebook_dim=C, commit_loss_legacy=legacy).eval()
        with torch.no_grad():
            vq.codebook.weight.copy_(w)
        xin = x.clone().requires_grad_(True)
        xq, loss, (_, _, codes) = vq(xin)
        (xq * gq).sum().backward(retain_graph=True)
        g_ste = xin.grad.clone()
        xin.grad = None
        loss.backward()
        put(f"q2_legacy{int(legacy)}", xq=xq, loss=loss, codes=codes, gx_ste=g_ste, gx_loss=xin.grad)

    # ---- quantize2_list.VectorQuantize2: ragged list, eval + one training pass
    K = 16
    wl = torch.randn(K + 1, C, generator=g)
    xs = [torch.randn(n, C, generator=g) for n in (5, 77, 130)]
    vq = VQ2List(codebook_size=K, codebook_dim=C).eval()
    with torch.no_grad():
        vq.codebook.weight.copy_(wl)
        vq.codebook.embed_ema.copy_(wl[:-1])
        vq.codebook.cluster_size_ema.fill_(1.0)
    xin = [t.clone().requires_grad_(True) for t in xs]
    xq_l, loss, (_, _, code_l) = vq(xin)
    loss.backward()
    put("ql", weight=wl, loss=loss, n_items=len(xs))
    for i in range(len(xs)):
        put(f"ql_{i}", x=xs[i], xq=xq_l[i], codes=code_l[i], gx=xin[i].grad)
    vq.train()
    # replay the RNG calls of _update_buffers item by item (quantize2_list.py:92-98)
    torch.manual_seed(321)
    restart = []
    for t in xs:
        v = t
        if v.shape[0] < K:
            v = vq.codebook._tile_with_noise(v, K)
        restart.append(v[torch.randperm(v.shape[0])][:K].clone())
    torch.manual_seed(321)
    xq_t, loss_t, (_, _, code_t) = vq([t.clone() for t in xs])
    put("ql_train", loss=loss_t, w=vq.codebook.weight, cs=vq.codebook.cluster_size_ema, em=vq.codebook.embed_ema)
    for i in range(len(xs)):
        put(f"ql_train_{i}", xq=xq_t[i], codes=code_t[i], restart=restart[i])

    # ---- quantize_rqvae.RQBottleneck: separate codebooks (depth 3) and a shared one (2x2 patches, depth 2)
    for tag, latent, code, shared, K in (("rq", (8, 8, 64), (8, 8, 3), False, 32),
                                          ("rqs", (8, 8, 64), (4, 4, 2), True, 32)):
        rq = RQBottleneck(latent_shape=latent, code_shape=code, n_embed=K, shared_codebook=shared).eval()
        ws = []
        with torch.no_grad():
            for d, cb in enumerate(rq.codebooks):
                if shared and d > 0:
                    ws.append(ws[0])
                    continue
                wd = torch.randn(K + 1, cb.weight.shape[1], generator=g) * (0.6 ** d)
                cb.weight.copy_(wd)
                cb.embed_ema.copy_(wd[:-1])
                cb.cluster_size_ema.fill_(1.0)
                ws.append(wd)
        xr = torch.randn(2, *latent, generator=g)
        gqr = torch.randn(2, *latent, generator=g)
        xin = xr.clone().requires_grad_(True)
        q, loss, codes = rq(xin)
        (q * gqr).sum().backward(retain_graph=True)
        g_ste = xin.grad.clone()
        xin.grad = None
        loss.backward()
        put(tag, x=xr, gq=gqr, quants=q, loss=loss, codes=codes, gx_ste=g_ste, gx_loss=xin.grad,
            embed=rq.embed_code(codes), depth=code[2])
        for d in range(code[2]):
            put(f"{tag}_w{d}", v=ws[d])
        rq.train()
        xt = torch.randn(2, *latent, generator=g) * 1.5
        # replay: depth d draws randperm over the residual rows entering depth d (quantize_rqvae.py:259-268)
        torch.manual_seed(654)
        n_rows = 2 * code[0] * code[1]
        perms = [torch.randperm(n_rows) for _ in range(code[2])]
        hooks, seen = [], []
        for cb in (list(rq.codebooks) if not shared else [rq.codebooks[0]]):
            hooks.append(cb.register_forward_pre_hook(lambda m, a: seen.append(a[0].detach().reshape(-1, a[0].shape[-1]).clone())))
        torch.manual_seed(654)
        q_t, loss_t, codes_t = rq(xt)
        for h in hooks:
            h.remove()
        put(f"{tag}_train", x=xt, quants=q_t, loss=loss_t, codes=codes_t)
        for d in range(code[2]):
            cb = rq.codebooks[d]
            put(f"{tag}_train_d{d}", restart=seen[d][perms[d]][:K], w=cb.weight, cs=cb.cluster_size_ema,
                em=cb.embed_ema)

    # ---- quantize_vqgan.VectorQuantizer2 (learnable codebook)
    n_e = 64
    we = torch.randn(n_e, C, generator=g) * 0.7
    z = torch.randn(2, C, 8, 8, generator=g)
    gz = torch.randn(2, C, 8, 8, generator=g)
    put("vg", weight=we, z=z, gq=gz)
    for legacy in (True, False):
        q = VectorQuantizer2(n_e, C, beta=0.25, legacy=legacy, sane_index_shape=not legacy)
        with torch.no_grad():
            q.embedding.weight.copy_(we)
        zin = z.clone().requires_grad_(True)
        zq, loss, (_, _, idx) = q(zin)
        (zq * gz).sum().backward(retain_graph=True)
        g_ste = zin.grad.clone()
        assert q.embedding.weight.grad is None or float(q.embedding.weight.grad.abs().max()) == 0.0
        zin.grad = None
        loss.backward()
        put(f"vg_legacy{int(legacy)}", zq=zq, loss=loss, idx=idx, gz_ste=g_ste, gz_loss=zin.grad,
            gw=q.embedding.weight.grad, entry=q.get_codebook_entry(idx.reshape(-1), (2, 8, 8, C)))
    save("vq_family.npz", **out)


# ----------------------------------------------------------------------------------- stage-2 permuter
def permuter_goldens():
    """modules/dynamic_modules/permuter.py of the reference: both fine orderings, forward + forward_back,
    on code maps whose coarse cells carry one code (what the dual-grain encoder emits), plus the edge cases
    all-coarse / all-fine samples, a sequence with duplicate positions and one without eos."""
    from modules.dynamic_modules.permuter import DualGrainSeperatePermuter
    out = {}
    g = torch.Generator().manual_seed(11)
    for tag, hw1, fhw, b in (("p8", 4, 8, 5), ("p32", 16, 32, 3), ("p8b", 4, 8, 4), ("p32b", 16, 32, 2)):
        grain = torch.randint(0, 2, (b, hw1, hw1), generator=g)
        if not tag.endswith("b"):                                          # "b": padded length < full length
            grain[0] = 0                                                   # all coarse: empty fine sequence
            grain[1] = 1                                                   # all fine: empty coarse sequence
        rep = grain.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
        fine_codes = torch.randint(0, 1024, (b, fhw, fhw), generator=g)
        cell_codes = torch.randint(0, 1024, (b, hw1, hw1), generator=g).repeat_interleave(2, -1).repeat_interleave(2, -2)
        indices = fine_codes * rep + cell_codes * (1 - rep)
        out[f"{tag}_indices"], out[f"{tag}_grain"] = indices, grain
        for order in ("region-first", "row-first"):
            p = DualGrainSeperatePermuter(coarse_hw=hw1, fine_hw=fhw, coarse_position_pad_code=hw1 * hw1,
                                          coarse_position_eos_code=hw1 * hw1 + 1, fine_position_order=order)
            o = p(indices, grain)
            back = p.forward_back(o["coarse_content"], o["fine_content"], o["coarse_position"], o["fine_position"])
            assert torch.equal(back, indices)                              # the reference's own round-trip check
            k = order[:3]
            for name, v in o.items():
                out[f"{tag}_{k}_{name}"] = v
            out[f"{tag}_{k}_back"] = back
    # forward_back on hand-made sequences: duplicates (last wins), garbage after the eos, a missing coarse eos
    p = DualGrainSeperatePermuter(coarse_hw=4, fine_hw=8, coarse_position_pad_code=16, coarse_position_eos_code=17)
    cc = torch.tensor([[5, 6, 7, 1025, 9, 1024], [1, 2, 3, 4, 5, 6]])
    cp = torch.tensor([[0, 3, 0, 17, 2, 16], [0, 1, 2, 3, 4, 5]])           # sample 1 never reaches the eos
    fc = torch.tensor([[40, 41, 42, 1025, 50], [60, 61, 1025, 1024, 1024]])
    fp = torch.tensor([[9, 9, 63, 1025, 0], [0, 10, 1025, 1024, 1024]])
    out["pb_cc"], out["pb_cp"], out["pb_fc"], out["pb_fp"] = cc, cp, fc, fp
    out["pb_back"] = p.forward_back(cc, fc, cp, fp)
    save("permuter.npz", **out)


def loss_goldens():
    """Stage-1 training loss of the reference (LPIPS + PatchGAN + adaptive weight) on seeded
    weights: torchvision's VGG16 is built WITHOUT the pretrained file (offline) and every tensor of the loss
    module is overwritten from oracle.loss_oracle.make_loss_weights.  xrec is a seeded toy "decoder tail"
    conv3x3(feat, w_last) so that the adaptive weight has a last layer to differentiate."""
    import torchvision
    import torch.nn.functional as F
    import modules.losses.lpips as ref_lpips
    from oracle import loss_oracle as lo
    tv_vgg16 = torchvision.models.vgg16

    class _Models:                       # what `from torchvision import models` gives lpips.py, minus the download
        @staticmethod
        def vgg16(pretrained=True):
            return tv_vgg16(weights=None)
    ref_lpips.models = _Models
    from modules.losses.vqperceptual_multidisc import VQLPIPSWithDiscriminator
    disc_cfg = {"target": "modules.discriminator.model.NLayerDiscriminator",
                "params": {"input_nc": 3, "ndf": 64, "n_layers": 3, "use_actnorm": False}}
    budget_cfg = {"target": "modules.dynamic_modules.budget.BudgetConstraint_RatioMSE_DualGrain",
                  "params": {"target_ratio": 0.5, "gamma": 10.0, "min_grain_size": 16, "max_grain_size": 32,
                             "calculate_all": True}}
    loss = VQLPIPSWithDiscriminator(disc_start=0, disc_config=disc_cfg, disc_init=True, codebook_weight=1.0,
                                    pixelloss_weight=1.0, disc_factor=1.0, disc_weight=1.0, perceptual_weight=1.0,
                                    disc_conditional=False, disc_adaptive_loss=True, disc_loss="hinge",
                                    disc_weight_max=0.75, budget_loss_config=budget_cfg)
    sd = lo.make_loss_weights(seed=21)
    ref_sd = loss.state_dict()
    mine = {k[len("loss."):]: v for k, v in sd.items()}
    missing = [k for k in ref_sd if k not in mine and "num_batches" not in k and "scaling_layer" not in k]
    extra = [k for k in mine if k not in ref_sd]
    assert not missing and not extra, (missing, extra)
    for k, v in mine.items():
        assert ref_sd[k].shape == v.shape, k
    loss.load_state_dict(mine, strict=False)
    out = {}
    for mode in ("eval", "train"):
        # "train" = what Lightning's model.train() leaves: BatchNorm uses batch statistics.  The Dropout in
        # front of every LPIPS lin head would also turn on (LPIPS().eval() at :74 does not survive
        # model.train()); it is switched off here so that the fixture is deterministic - the product keeps
        # the reference behaviour and the GPU test checks the dropout statistics separately.
        loss.train(mode == "train")
        loss.perceptual_loss.eval()
        loss.load_state_dict(mine, strict=False)                     # reset BatchNorm running statistics
        x, feat, w_last, qloss, gate = lo.toy_inputs()
        w_last.requires_grad_(True)
        feat.requires_grad_(True)
        xrec = F.conv2d(feat, w_last, padding=1)
        l0, log0 = loss(qloss, x, xrec, 0, 0, last_layer=w_last, split="train", gate=gate)
        g_w, g_feat = torch.autograd.grad(l0, [w_last, feat])
        bn_after0 = {k: v.clone() for k, v in loss.state_dict().items() if "running" in k}
        l1, log1 = loss(qloss, x, xrec.detach(), 1, 0, last_layer=w_last, split="train")
        dparams = dict(loss.discriminator.named_parameters())
        g_d = torch.autograd.grad(l1, list(dparams.values()))
        bn_after1 = {k: v.clone() for k, v in loss.state_dict().items() if "running" in k}
        p = mode + "_"
        out.update({p + "loss0": l0, p + "g_w_last": g_w, p + "g_feat": g_feat, p + "loss1": l1, p + "xrec": xrec})
        out.update({p + "log0_" + k.replace("train_", ""): v for k, v in log0.items()})
        out.update({p + "log1_" + k.replace("train_", ""): v for k, v in log1.items()})
        out.update({p + "gd_norm_" + k: g.norm() for k, g in zip(dparams, g_d)})
        out[p + "gd_main.0.weight"] = g_d[list(dparams).index("main.0.weight")]
        out[p + "gd_main.11.weight"] = g_d[list(dparams).index("main.11.weight")]
        out.update({p + "bn0_" + k: v for k, v in bn_after0.items()})
        out.update({p + "bn1_" + k: v for k, v in bn_after1.items()})
        # LPIPS alone, per image
        out[p + "lpips"] = loss.perceptual_loss(x, xrec.detach())
    save("loss_small.npz", **out)


def threshold_inputs(seed=3, n_batches=3, b=4, res=64, patch=16):
    """Seeded image batches in [0,1] mixing noise and flat patches (regenerated by the tests)."""
    g = torch.Generator().manual_seed(seed)
    k = res // patch
    out = []
    for _ in range(n_batches):
        x = torch.rand(b, 3, res, res, generator=g)
        flat = torch.rand(b, 3, k, k, generator=g).repeat_interleave(patch, 2).repeat_interleave(patch, 3)
        pick = (torch.rand(b, 1, k, k, generator=g) > 0.5).float().repeat_interleave(patch, 2).repeat_interleave(patch, 3)
        out.append(pick * x + (1 - pick) * flat)
    return out


def threshold_goldens():
    """scripts/tools/calculate_entropy_thresholds.py of the reference: its Entropy class (bins on [0,1]) and its
    percentile procedure (:95-117), run on seeded batches (the data-set loaders are not importable offline)."""
    src = open(os.path.join(REF, "scripts/tools/calculate_entropy_thresholds.py")).read().split("if __name__")[0]
    src = src.replace("from data.imagenet_lmdb import Imagenet_LMDB", "").replace("from data.ffhq_lmdb import FFHQ_LMDB", "")
    ns = {}
    exec(compile(src, "calculate_entropy_thresholds", "exec"), ns)
    model = ns["Entropy"](16, 64, 64)
    with torch.no_grad():
        for i, image in enumerate(threshold_inputs()):
            if i == 0:
                entropy_numpy = model(image).view(-1).cpu().numpy()
            else:
                entropy_numpy = np.concatenate((entropy_numpy, model(image).view(-1).cpu().numpy()))
    entropy_numpy = np.sort(entropy_numpy)
    size = entropy_numpy.shape[0]
    th = np.array([entropy_numpy[int((size * (i + 1)) // 100)] for i in range(99)], dtype=np.float32)
    save("entropy_thresholds.npz", thresholds=th, entropies=entropy_numpy)


if __name__ == "__main__":
    what = sys.argv[1:] or ["vq", "family", "permuter", "tiny", "dual", "variants", "loss", "thresholds", "routing",
                            "triple_grads"]
    if "triple_grads" in what:
        triple_backward_goldens()
    if "routing" in what:
        train_routing_goldens()
    if "thresholds" in what:
        threshold_goldens()
    if "loss" in what:
        loss_goldens()
    if "variants" in what:
        triple_entropy_goldens()
    if "vq" in what:
        vq_goldens()
    if "family" in what:
        vq_family_goldens()
    if "permuter" in what:
        permuter_goldens()
    if "tiny" in what:
        model_goldens(orc.TINY_CFG, "tiny", batch=2, seed=3)
    if "dual" in what:
        model_goldens(orc.DUAL_CFG, "dual", batch=1, seed=7)
