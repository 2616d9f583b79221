"""Module-level parity: the overlay classes are driven through their
reference-facing interface (NCHW fp32 in / out, reference constructor arguments) and compared with
the oracle functions on the same state_dict."""
import numpy as np
import pytest
import torch

from parity_util import audit_codes, bf16_operands

BF = torch.bfloat16


def rel_rms(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(((a - b).pow(2).mean() / b.pow(2).mean().clamp_min(1e-30)).sqrt())


def _overlay():
    from dynamicvectorquantization_b200 import configs
    configs.activate_overlay()


def _randomise(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if p.dim() == 1:
                p.copy_((1.0 if n.endswith("weight") else 0.0) + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * (p[0].numel()) ** -0.5)
    return {k: v.detach().clone() for k, v in mod.state_dict().items()}


def _prefixed(sd, prefix):
    return {f"{prefix}.{k}": v for k, v in sd.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,res", [(128, 128, 32), (128, 256, 16), (64, 64, 128)])
def test_resnet_block_public_interface(cin, cout, res):
    _overlay()
    from modules.diffusionmodules.model import ResnetBlock
    from oracle import dqvae_oracle as orc
    blk = ResnetBlock(in_channels=cin, out_channels=cout, temb_channels=0, dropout=0.0)
    sd = _randomise(blk, 1)
    x = torch.randn(2, cin, res, res, generator=torch.Generator().manual_seed(2))
    xr = x.clone().requires_grad_(True)
    ref = orc.resnet_block(_prefixed(sd, "b"), "b", xr)
    gy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    ref.backward(gy)
    blk = blk.cuda()
    xd = x.cuda().requires_grad_(True)
    out = blk(xd, None)
    out.backward(gy.cuda())
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel_rms(out.detach(), ref.detach()) < 1e-2
    assert rel_rms(xd.grad, xr.grad) < 3e-2
    with pytest.raises(NotImplementedError):
        blk(xd, torch.zeros(2, 8, device="cuda"))           # timestep embeddings are not part of the path


@pytest.mark.gpu
@pytest.mark.parametrize("c,res", [(256, 32), (512, 16), (512, 8)])
def test_attn_block_public_interface(c, res):
    _overlay()
    from modules.diffusionmodules.model import AttnBlock
    from oracle import dqvae_oracle as orc
    blk = AttnBlock(c)
    sd = _randomise(blk, 4)
    x = torch.randn(2, c, res, res, generator=torch.Generator().manual_seed(5))
    xr = x.clone().requires_grad_(True)
    ref = orc.attn_block(_prefixed(sd, "a"), "a", xr)
    gy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(6))
    ref.backward(gy)
    blk = blk.cuda()
    xd = x.cuda().requires_grad_(True)
    out = blk(xd)
    out.backward(gy.cuda())
    assert rel_rms(out.detach(), ref.detach()) < 1e-2
    assert rel_rms(xd.grad, xr.grad) < 3e-2
    wq = blk.q.weight.grad
    assert wq is not None and torch.isfinite(wq).all()


@pytest.mark.gpu
def test_resampling_blocks_public_interface():
    _overlay()
    from modules.diffusionmodules.model import Downsample, Upsample
    from oracle import dqvae_oracle as orc
    for cls, fn, c, res in ((Downsample, orc.downsample, 128, 64), (Upsample, orc.upsample, 256, 16)):
        m = cls(c, True)
        sd = _randomise(m, 7)
        x = torch.randn(2, c, res, res, generator=torch.Generator().manual_seed(8))
        xr = x.clone().requires_grad_(True)
        sdr = {k: v.clone().requires_grad_(True) for k, v in _prefixed(sd, "m").items()}
        ref = fn(sdr, "m", xr)
        gy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(9))
        ref.backward(gy)
        m = m.cuda()
        xd = x.cuda().requires_grad_(True)
        out = m(xd)
        out.backward(gy.cuda())
        assert out.shape == ref.shape
        assert rel_rms(out.detach(), ref.detach()) < 1e-2, cls.__name__
        assert rel_rms(xd.grad, xr.grad) < 2e-2, cls.__name__
        assert rel_rms(m.conv.weight.grad, sdr["m.conv.weight"].grad) < 2e-2, cls.__name__
        assert rel_rms(m.conv.bias.grad, sdr["m.conv.bias"].grad) < 2e-2, cls.__name__


@pytest.mark.gpu
def test_decoder_public_interface_and_last_layer():
    """Decoder.forward(h, grain_indices) on NCHW fp32 + the re-entrant autograd contract of the reference
    loss: torch.autograd.grad(..., decoder.conv_out.weight, retain_graph=True) twice, then backward."""
    _overlay()
    from modules.dynamic_modules.DecoderPositional import Decoder
    from oracle import dqvae_oracle as orc
    cfg = orc.SMALL_CFG
    dec = Decoder(ch=64, in_ch=64, out_ch=3, ch_mult=[1, 1, 2, 2], num_res_blocks=2, resolution=64,
                  attn_resolutions=[8], latent_size=8, window_size=2, position_type="fourier+learned")
    sd = {k[len("decoder."):]: v for k, v in orc.make_weights(orc.decoder_shapes(cfg), seed=3).items()}
    dec.load_state_dict(sd, strict=True)
    z = torch.randn(2, 64, 8, 8, generator=torch.Generator().manual_seed(1))
    ref = orc.decoder(_prefixed(sd, "decoder"), cfg, z)
    dec = dec.cuda()
    zd = z.cuda().requires_grad_(True)
    out = dec(zd, None)
    assert rel_rms(out.detach(), ref) < 2e-2
    last = dec.conv_out.weight
    g1 = torch.autograd.grad(out.abs().mean(), last, retain_graph=True)[0]
    g2 = torch.autograd.grad(out.pow(2).mean(), last, retain_graph=True)[0]
    out.mean().backward()
    assert torch.isfinite(g1).all() and torch.isfinite(g2).all() and last.grad is not None and zd.grad is not None


@pytest.mark.gpu
def test_vq_training_trajectory_matches_oracle():
    """Three training forwards of VectorQuantize2 (EMA counts/sums, dead-code restart with the rows the
    module really drew, re-normalised weights) vs the numpy restatement of quantize2_mask.py:66-115."""
    _overlay()
    from modules.vector_quantization.quantize2_mask import VectorQuantize2
    from oracle import vq_oracle as vo
    K, C, B, H = 96, 64, 2, 12
    vq = VectorQuantize2(codebook_size=K, codebook_dim=C).cuda().train()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.randn(K + 1, C, generator=g))
        vq.codebook.embed_ema.copy_(vq.codebook.weight[:-1])
        vq.codebook.cluster_size_ema.fill_(1.0)
    w = vq.codebook.weight.detach().cpu().numpy().copy()
    cs = np.ones(K, np.float32)
    em = w[:-1].copy()
    for step in range(3):
        x = torch.randn(B, C, H, H, generator=g) * (1.0 + 0.5 * step)
        mask = torch.where(torch.rand(B, 1, H, H, generator=g) > 0.5, 1.0, 0.25)
        torch.manual_seed(50 + step)
        perm = torch.randperm(B * H * H, device="cuda")            # what _ema_step will draw
        torch.manual_seed(50 + step)
        xq, loss, (_, _, codes) = vq(x.cuda(), codebook_mask=mask.cuda())
        flat = x.permute(0, 2, 3, 1).reshape(-1, C).numpy()
        got = codes.reshape(-1).cpu().numpy()
        audit_codes(*bf16_operands(flat, w), got, f"trajectory step {step}")
        assert np.allclose(xq.detach().permute(0, 2, 3, 1).reshape(-1, C).cpu().numpy(), w[got], atol=2e-6)
        restart = flat[perm.cpu().numpy()][:K]
        cs, em = vo.update_buffers(flat, got, cs, em, 0.99, restart_rows=restart)
        w[:-1] = vo.update_embedding(cs, em)
        assert np.allclose(vq.codebook.cluster_size_ema.cpu().numpy(), cs, rtol=1e-5, atol=1e-6), step
        assert np.allclose(vq.codebook.embed_ema.cpu().numpy(), em, rtol=1e-4, atol=1e-5), step
        assert np.allclose(vq.codebook.weight.detach().cpu().numpy(), w, rtol=1e-4, atol=1e-5), step


@pytest.mark.gpu
def test_vq_sequence_input_and_helpers():
    """accept_image_fmap=False / channel_last=True path, get_codebook_entry, compute_distances."""
    _overlay()
    from modules.vector_quantization.quantize2_mask import VectorQuantize2
    from oracle import vq_oracle as vo
    K, C = 80, 64
    vq = VectorQuantize2(codebook_size=K, codebook_dim=C, accept_image_fmap=False, channel_last=True).cuda().eval()
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.randn(K + 1, C, generator=torch.Generator().manual_seed(1)))
    x = torch.randn(3, 50, C, generator=torch.Generator().manual_seed(2))
    xq, loss, (_, _, codes) = vq(x.cuda())
    w = vq.codebook.weight.detach().cpu().numpy()
    assert codes.shape == (3, 50)
    audit_codes(*bf16_operands(x.reshape(-1, C).numpy(), w), codes.reshape(-1).cpu().numpy(), "sequence input")
    ent = vq.get_codebook_entry(codes)
    assert torch.equal(ent, vq.codebook.weight[codes])
    d = vq.codebook.compute_distances(x.cuda())
    dref = vo.compute_distances(vo.bf16_round(x.reshape(-1, C).numpy()), np.concatenate([vo.bf16_round(w[:-1]), w[-1:]]))
    assert np.allclose(d.reshape(-1, K).cpu().numpy(), dref, rtol=1e-3, atol=1e-2)
    soft, code = vq.get_soft_codes(x.cuda(), temp=1.0)
    assert soft.shape == (3, 50, K) and torch.allclose(soft.sum(-1), torch.ones(3, 50, device="cuda"), atol=1e-4)


# --------------------------------------------------------------------------- CPU: fp32 PyTorch parts
def test_routers_budget_entropy_match_oracle_on_cpu():
    _overlay()
    from modules.dynamic_modules.RouterDual import DualGrainFeatureRouter
    from modules.dynamic_modules.RouterTriple import TripleGrainFeatureRouter
    from modules.dynamic_modules.budget import (BudgetConstraint_NormedSeperateRatioMSE_TripleGrain,
                                                BudgetConstraint_RatioMSE_DualGrain)
    from oracle import dqvae_oracle as orc
    g = torch.Generator().manual_seed(0)
    r2 = DualGrainFeatureRouter(num_channels=64, normalization_type="group-32", gate_type="2layer-fc-SiLu")
    sd = _randomise(r2, 1)
    hf, hc = torch.randn(2, 64, 8, 8, generator=g), torch.randn(2, 64, 4, 4, generator=g)
    assert torch.allclose(r2(h_fine=hf, h_coarse=hc), orc.feature_router(_prefixed(sd, "r"), "r", hf, hc), atol=1e-5)
    r3 = TripleGrainFeatureRouter(num_channels=64, normalization_type="group-32", gate_type="2layer-fc-SiLu")
    sd3 = _randomise(r3, 2)
    hm, hc2 = torch.randn(2, 64, 4, 4, generator=g), torch.randn(2, 64, 2, 2, generator=g)
    assert torch.allclose(r3(h_fine=hf, h_median=hm, h_coarse=hc2),
                          orc.triple_router(_prefixed(sd3, "r"), "r", hf, hm, hc2), atol=1e-5)
    gate2 = torch.softmax(torch.randn(3, 2, 16, 16, generator=g), 1)
    b2 = BudgetConstraint_RatioMSE_DualGrain(target_ratio=0.5, gamma=10.0, min_grain_size=16, max_grain_size=32,
                                             calculate_all=True)
    assert torch.allclose(b2(gate2), orc.budget_loss_dual(gate2))
    gate3 = torch.softmax(torch.randn(3, 3, 8, 8, generator=g), 1)
    b3 = BudgetConstraint_NormedSeperateRatioMSE_TripleGrain(target_fine_ratio=0.3, target_median_ratio=0.3, gamma=1.0,
                                                             min_grain_size=8, median_grain_size=16, max_grain_size=32)
    assert torch.allclose(b3(gate3), orc.budget_loss_triple(gate3))


@pytest.mark.gpu
@pytest.mark.parametrize("patch,size,batch", [(16, 256, 3), (16, 64, 2), (8, 64, 2), (4, 32, 1)])
def test_patch_entropy_kernel_matches_oracle(patch, size, batch):
    """csrc/entropy.cu vs the fp32 restatement of dqvae_dual_entropy.py:25-63: noise patches (entropy ~3),
    constant patches (entropy ~0, every other bin at the 1e-40 floor) and smooth ramps."""
    _overlay()
    from models.stage1_dynamic.dqvae_dual_entropy import Entropy
    from oracle import dqvae_oracle as orc
    g = torch.Generator().manual_seed(patch * 1000 + size)
    x = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    n = size // patch
    kind = torch.randint(0, 3, (batch, 1, n, n), generator=g).repeat_interleave(patch, 2).repeat_interleave(patch, 3)
    const = (torch.rand(batch, 3, n, n, generator=g) * 2 - 1).repeat_interleave(patch, 2).repeat_interleave(patch, 3)
    ramp = torch.linspace(-1, 1, size).view(1, 1, 1, size).expand(batch, 3, size, size)
    x = torch.where(kind == 0, x, torch.where(kind == 1, const, ramp)).contiguous()
    ref = orc.patch_entropy(x, patch)
    mod = Entropy(patch, size, size)
    got = mod(x.cuda())
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert torch.allclose(got.cpu(), ref, rtol=2e-5, atol=2e-6), float((got.cpu() - ref).abs().max())
    assert float(ref.min()) < 0.1 < 2.5 < float(ref.max()) or patch == 4      # both regimes are present
    with pytest.raises(RuntimeError):
        mod(x)                                                                 # no CPU fallback


@pytest.mark.gpu
def test_entropy_threshold_tool_matches_the_reference_procedure(tmp_path):
    """nn/thresholds.py (fused entropy kernel, bins on [0,1], device-side sort) against the oracle restatement of
    scripts/tools/calculate_entropy_thresholds.py and the fixture minted from the reference tool itself, on images
    that mix flat and noise patches."""
    import json
    import os
    import numpy as np
    from dynamicvectorquantization_b200.nn.thresholds import EntropyThresholds
    from oracle import dqvae_oracle as orc
    from test_oracle_golden import _threshold_inputs
    batches = _threshold_inputs()
    acc = EntropyThresholds(patch_size=16, image_size=64)
    for b in batches:
        acc.update(b.cuda())
    got = acc.thresholds()
    ref = orc.entropy_thresholds(batches, patch=16)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "entropy_thresholds.npz"))["thresholds"]
    assert list(got) == [str(i) for i in range(1, 100)]
    for i, k in enumerate(ref):
        assert abs(got[k] - ref[k]) <= 2e-4 * abs(ref[k]) + 2e-5, (k, got[k], ref[k])
        assert abs(got[k] - float(gold[i])) <= 2e-4 * abs(float(gold[i])) + 2e-5, (k, got[k], float(gold[i]))
    acc.save(str(tmp_path / "t.json"))
    assert json.load(open(tmp_path / "t.json"))["50"] == got["50"]
