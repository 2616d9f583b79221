"""GPU parity of the stage-1 training loss: LPIPS on the tensor-core convolution kernels, the
PatchGAN discriminator and the adaptive-weight loss, against oracle/loss_oracle.py (fp32, CPU) and the fixtures
minted from the reference's own classes (tests/golden/loss_small.npz).

Tolerances: the VGG16 stack runs in bf16 (13 convolution layers, fp32 accumulation), so LPIPS values are compared
at 2e-2 relative and its gradients at 6e-2 relative RMS.  The PatchGAN runs on the same bf16 tensor-core kernels
(5 convolution layers, fp32 accumulation, fp32 BatchNorm statistics): logits / discriminator-loss terms at 1e-2,
its gradients at 3e-2 relative RMS (written next to each assert)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_rms(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(((a - b).pow(2).mean() / b.pow(2).mean().clamp_min(1e-30)).sqrt())


def _loss_module(monkeypatch, seed=21, budget=True):
    monkeypatch.setenv("B200DQ_ALLOW_RANDOM_VGG", "1")
    from dynamicvectorquantization_b200 import configs
    configs.activate_overlay()
    from modules.losses.vqperceptual_multidisc import VQLPIPSWithDiscriminator
    from oracle import loss_oracle as lo
    disc_cfg = {"target": "modules.discriminator.model.NLayerDiscriminator",
                "params": {"input_nc": 3, "ndf": 64, "n_layers": 3, "use_actnorm": False}}
    loss = VQLPIPSWithDiscriminator(disc_start=0, disc_config=disc_cfg, disc_init=True, disc_weight_max=0.75,
                                    budget_loss_config=configs._BUDGET_DUAL if budget else None)
    sd = lo.make_loss_weights(seed)
    loss.load_state_dict({k[len("loss."):]: v for k, v in sd.items()}, strict=False)
    return loss.cuda(), sd


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


@pytest.mark.parametrize("res,nb", [(64, 2), (256, 2)])
def test_patchgan_layers_match_fp32_ops(res, nb):
    """Every PatchGAN layer on its own, on IDENTICAL bf16 inputs and bf16-rounded weights, against torch's fp32 ops on
    the GPU (F.conv2d, nn.BatchNorm2d in training mode, F.leaky_relu): the 3-channel stem (im2col + GEMM + LeakyReLU
    epilogue; image / weight / bias gradients), the 4x4 stride-2 and stride-1 convolutions (16-tap GEMMs; data and
    weight gradients), BatchNorm + LeakyReLU (values, running statistics, dx / dgamma / dbeta) and the one-channel
    head (two-term bf16 cotangent).  Isolated layers see no activation-sign flips, so the bounds are the bf16
    rounding of the outputs: 4e-3 for bf16 tensors, 1e-3 for fp32 results."""
    from dynamicvectorquantization_b200.nn import discriminator as D
    g = torch.Generator(device="cuda").manual_seed(res)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x = (torch.rand(nb, res, res, 3, device="cuda", generator=g) * 2 - 1).to(BF)
    w0, b0 = (rnd(64, 3, 4, 4) * 0.2).requires_grad_(True), (rnd(64) * 0.1).requires_grad_(True)
    xg = x.clone().requires_grad_(True)
    y = D._StemFn.apply(xg, w0, b0)
    xr = _nchw(x).requires_grad_(True)
    w0r, b0r = w0.detach().to(BF).float().requires_grad_(True), b0.detach().clone().requires_grad_(True)
    ref = F.leaky_relu(F.conv2d(xr, w0r, b0r, stride=2, padding=1), 0.2)
    assert rel_rms(_nchw(y.detach()), ref.detach()) < 4e-3
    cot = rnd(*ref.shape)
    gy = torch.autograd.grad(y, [xg, w0, b0], cot.permute(0, 2, 3, 1).to(BF).contiguous())
    gr = torch.autograd.grad(ref, [xr, w0r, b0r], cot.to(BF).float())
    assert rel_rms(_nchw(gy[0]), gr[0]) < 4e-3 and rel_rms(gy[1], gr[1]) < 1e-3 and rel_rms(gy[2], gr[2]) < 1e-3
    h = y.detach()
    for cin, cout, stride in ((64, 128, 2), (128, 256, 2), (256, 512, 1)):
        w = (rnd(cout, cin, 4, 4) * (cin * 16) ** -0.5).requires_grad_(True)
        hg = h.clone().requires_grad_(True)
        z = D._Conv4x4Fn.apply(hg, w, None, stride)
        hr = _nchw(h).requires_grad_(True)
        wr = w.detach().to(BF).float().requires_grad_(True)
        zr = F.conv2d(hr, wr, None, stride=stride, padding=1)
        assert rel_rms(_nchw(z.detach()), zr.detach()) < 4e-3, (cin, cout, stride)
        cot = rnd(*zr.shape)
        gz = torch.autograd.grad(z, [hg, w], cot.permute(0, 2, 3, 1).to(BF).contiguous())
        gzr = torch.autograd.grad(zr, [hr, wr], cot.to(BF).float())
        assert rel_rms(_nchw(gz[0]), gzr[0]) < 4e-3 and rel_rms(gz[1], gzr[1]) < 1e-3, (cin, cout, stride)
        bn = torch.nn.BatchNorm2d(cout).cuda().train()
        with torch.no_grad():
            bn.weight.normal_(1, 0.2); bn.bias.normal_(0, 0.2)
        bn2 = torch.nn.BatchNorm2d(cout).cuda().train()
        bn2.load_state_dict(bn.state_dict())
        zd = z.detach().clone().requires_grad_(True)
        a = D._batchnorm_lrelu(zd, bn)
        zrr = _nchw(z.detach()).requires_grad_(True)
        ar = F.leaky_relu(bn2(zrr), 0.2)
        assert rel_rms(_nchw(a.detach()), ar.detach()) < 4e-3
        assert rel_rms(bn.running_mean, bn2.running_mean) < 1e-5 and rel_rms(bn.running_var, bn2.running_var) < 1e-5
        assert int(bn.num_batches_tracked) == 1
        cot = rnd(*ar.shape)
        ga = torch.autograd.grad(a, [zd, bn.weight, bn.bias], cot.permute(0, 2, 3, 1).to(BF).contiguous())
        gar = torch.autograd.grad(ar, [zrr, bn2.weight, bn2.bias], cot.to(BF).float())
        assert rel_rms(_nchw(ga[0]), gar[0]) < 4e-3 and rel_rms(ga[1], gar[1]) < 1e-4 and rel_rms(ga[2], gar[2]) < 1e-4
        # evaluation mode: running statistics are constants of the backward
        bn.eval(); bn2.eval()
        ze = z.detach().clone().requires_grad_(True)
        ae = D._batchnorm_lrelu(ze, bn)
        zer = _nchw(z.detach()).requires_grad_(True)
        aer = F.leaky_relu(bn2(zer), 0.2)
        assert rel_rms(_nchw(ae.detach()), aer.detach()) < 4e-3
        ge = torch.autograd.grad(ae, [ze, bn.weight, bn.bias], cot.permute(0, 2, 3, 1).to(BF).contiguous())
        ger = torch.autograd.grad(aer, [zer, bn2.weight, bn2.bias], cot.to(BF).float())
        assert rel_rms(_nchw(ge[0]), ger[0]) < 4e-3 and rel_rms(ge[1], ger[1]) < 1e-4 and rel_rms(ge[2], ger[2]) < 1e-4
        h = a.detach()
    wh, bh = (rnd(1, 512, 4, 4) * (512 * 16) ** -0.5).requires_grad_(True), rnd(1).requires_grad_(True)
    hg = h.clone().requires_grad_(True)
    o = D._HeadFn.apply(hg, wh, bh)
    hr = _nchw(h).requires_grad_(True)
    whr, bhr = wh.detach().to(BF).float().requires_grad_(True), bh.detach().clone().requires_grad_(True)
    orf = F.conv2d(hr, whr, bhr, stride=1, padding=1)
    assert rel_rms(_nchw(o.detach()), orf.detach()) < 1e-4
    for kind in ("random", "mean"):
        cot = rnd(*orf.shape) if kind == "random" else torch.full_like(orf, -1.0 / orf.numel())
        go = torch.autograd.grad(o, [hg, wh, bh], cot.permute(0, 2, 3, 1).contiguous(), retain_graph=True)
        gor = torch.autograd.grad(orf, [hr, whr, bhr], cot, retain_graph=True)
        assert rel_rms(_nchw(go[0]), gor[0]) < 4e-3 and rel_rms(go[1], gor[1]) < 1e-3 and rel_rms(go[2], gor[2]) < 1e-5
        assert abs(float(go[1].norm() / gor[1].norm()) - 1.0) < 1e-3          # no 2^-9 bias from a bf16 cotangent


@pytest.mark.parametrize("batch,res,train", [(2, 64, True), (2, 64, False), (3, 128, True), (4, 256, True)])
def test_patchgan_on_the_kernels_matches_oracle(batch, res, train):
    """NLayerDiscriminator (modules/discriminator/model.py:17-67) through its reference-facing interface (NCHW fp32
    in, NCHW fp32 logits out) on the hand-written kernels against the fp32 oracle: logits, updated running
    statistics, and the gradients w.r.t. every parameter and w.r.t. the input image (the generator-loss path), for a
    random cotangent and for the constant cotangent of a mean (hinge) loss.

    Gradient bounds: the network is piecewise linear (four LeakyReLU layers).  The bf16 activations differ from the
    fp32 oracle's by ~2e-3 relative, which flips the sign of the few pre-activations that lie that close to zero; a
    flipped unit changes its local derivative between 1 and 0.2.  With a flip probability of ~2.5e-3 per unit and
    layer that is ~5 % RMS per LeakyReLU layer, ~10-14 % end to end (measured), as noise that does not turn the
    gradient (cosine >= 0.99) nor scale it (norms within 5 %); the layers themselves are exact to bf16 rounding
    (test_patchgan_layers_match_fp32_ops)."""
    from dynamicvectorquantization_b200 import configs
    configs.activate_overlay()
    from modules.discriminator.model import NLayerDiscriminator
    from oracle import loss_oracle as lo
    sd = {k: v for k, v in lo.make_loss_weights(seed=31).items() if k.startswith("loss.discriminator.")}
    disc = NLayerDiscriminator(input_nc=3, ndf=64, n_layers=3, use_actnorm=False)
    disc.load_state_dict({k[len("loss.discriminator."):]: v for k, v in sd.items()}, strict=False)
    disc = disc.cuda().train(train)
    g = torch.Generator().manual_seed(batch * res)
    x = torch.rand(batch, 3, res, res, generator=g) * 2 - 1
    xd = x.cuda().requires_grad_(True)
    logits = disc(xd)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = dict(sd); full.update(params)
    xr = x.clone().requires_grad_(True)
    new_stats = {}
    ref = lo.discriminator(full, xr, train=train, new_stats=new_stats)
    assert logits.shape == ref.shape
    e = rel_rms(logits.detach(), ref.detach())
    assert e < 1e-2, f"logits rel-RMS {e}"
    if train:
        for k, v in new_stats.items():
            got = disc.state_dict()[k[len("loss.discriminator."):]].cpu()
            if k.endswith("running_var"):
                assert torch.allclose(got, v, rtol=1e-2), k
            else:       # means of zero-centred activations: judged against the batch standard deviation of the channel
                std = ((new_stats[k.replace("running_mean", "running_var")] - 0.9) / 0.1).clamp_min(0).sqrt()
                assert float(((got - v).abs() / (0.1 * std + 1e-12)).max()) < 1e-2, k
        assert int(disc.main[3].num_batches_tracked) == 1
    names = [k for k, _ in disc.named_parameters()]
    for kind in ("random", "mean"):
        cot = torch.randn(ref.shape, generator=g) if kind == "random" else torch.full(ref.shape, -1.0 / ref.numel())
        got = torch.autograd.grad(logits, [xd] + list(disc.parameters()), cot.cuda(), retain_graph=True)
        want = torch.autograd.grad(ref, [xr] + [params["loss.discriminator." + n] for n in names], cot, retain_graph=True)
        errs = {n: rel_rms(a, b) for n, a, b in zip(["input"] + names, got, want)}
        coss = {n: _cos(a, b) for n, a, b in zip(["input"] + names, got, want)}
        worst, wcos = max(errs.items(), key=lambda kv: kv[1]), min(coss.items(), key=lambda kv: kv[1])
        print(f"patchgan {batch}x{res} train={train} cotangent={kind}: logits {e:.2e}, worst gradient rel-RMS {worst}, "
              f"lowest cosine {wcos}")
        assert worst[1] < 0.16 and wcos[1] > 0.99, (errs, coss)      # measured: 0.10-0.14, cosines 0.990-0.995
        for n, a, b in zip(["input"] + names, got, want):
            assert abs(float(a.norm()) / float(b.norm()) - 1.0) < 5e-2, (kind, n, float(a.norm()), float(b.norm()))


def test_maxpool_and_relu_gradients_match_torch():
    from dynamicvectorquantization_b200 import kernels as kn
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 12, 20, 64, generator=g).clamp_min(0).to(BF)       # ReLU output: many exact ties at 0
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    y = kn.maxpool2x2(x.cuda())
    assert torch.equal(y.cpu().float(), yr.detach().permute(0, 2, 3, 1))
    dy = torch.randn(3, 6, 10, 64, generator=g).to(BF)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    dx = kn.maxpool2x2_bwd(dy.cuda(), x.cuda())
    assert torch.equal(dx.cpu().float(), xr.grad.permute(0, 2, 3, 1))
    d2 = torch.randn(3, 12, 20, 64, generator=g).to(BF)
    got = kn.relu_bwd(d2.cuda(), x.cuda()).cpu()
    assert torch.equal(got, torch.where(x > 0, d2, torch.zeros_like(d2)))


@pytest.mark.parametrize("res,batch", [(64, 2), (128, 3)])
def test_lpips_matches_oracle(monkeypatch, res, batch):
    from oracle import loss_oracle as lo
    loss, sd = _loss_module(monkeypatch)
    lp = loss.perceptual_loss.eval()
    g = torch.Generator().manual_seed(res)
    x = torch.rand(batch, 3, res, res, generator=g) * 2 - 1
    y = (x + 0.3 * torch.randn(batch, 3, res, res, generator=g)).clamp(-1, 1)
    yr = y.clone().requires_grad_(True)
    ref = lo.lpips(sd, x, yr)
    ref.sum().backward()
    yd = y.cuda().requires_grad_(True)
    got = lp(x.cuda(), yd)
    got.sum().backward()
    assert got.shape == (batch, 1, 1, 1)
    assert torch.allclose(got.detach().cpu(), ref.detach(), rtol=2e-2, atol=1e-4), (got.flatten(), ref.flatten())
    assert rel_rms(yd.grad, yr.grad) < 6e-2
    # gradient w.r.t. the first argument too (both sides are differentiable in the reference)
    xd = x.cuda().requires_grad_(True)
    lp(xd, y.cuda()).sum().backward()
    xr = x.clone().requires_grad_(True)
    lo.lpips(sd, xr, y).sum().backward()
    assert rel_rms(xd.grad, xr.grad) < 6e-2
    # features through the reference-facing NCHW interface
    feats = lp.net(lp.scaling_layer(x.cuda()))
    shift = torch.tensor(lo.SHIFT)[None, :, None, None]
    scale = torch.tensor(lo.SCALE)[None, :, None, None]
    ref_feats = lo.vgg_features(sd, (x - shift) / scale)
    for a, b in zip(feats, ref_feats):
        assert a.shape == b.shape and rel_rms(a, b) < 2e-2


def test_lpips_training_mode_dropout_is_live_and_unbiased(monkeypatch):
    loss, _ = _loss_module(monkeypatch)
    lp = loss.perceptual_loss
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    y = (x + 0.3 * torch.randn(2, 3, 64, 64, generator=g).cuda()).clamp(-1, 1)
    with torch.no_grad():
        ref = lp.eval()(x, y)
        lp.train()                                    # what Lightning's model.train() does to the frozen metric
        torch.manual_seed(0)
        draws = torch.stack([lp(x, y) for _ in range(24)])
    assert float(draws.std(0).max()) > 0, "dropout in front of the lin heads should be active in training mode"
    assert torch.allclose(draws.mean(0), ref, rtol=0.1), (draws.mean(0).flatten(), ref.flatten())


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_full_loss_matches_oracle_and_reference_golden(monkeypatch, mode):
    from oracle import loss_oracle as lo
    gold = np.load(os.path.join(GOLD, "loss_small.npz"))
    loss, sd = _loss_module(monkeypatch)
    loss.train(mode == "train")
    loss.perceptual_loss.eval()                       # deterministic comparison (see the dropout test)
    x, feat, w_last, qloss, gate = lo.toy_inputs()
    xd, gd, qd = x.cuda(), gate.cuda(), qloss.cuda()
    wl = w_last.cuda().requires_grad_(True)
    ft = feat.cuda().requires_grad_(True)
    xrec = F.conv2d(ft, wl, padding=1)
    l0, log0 = loss(qd, xd, xrec, 0, 0, last_layer=wl, split="train", gate=gd)
    gw, gf = torch.autograd.grad(l0, [wl, ft])
    # oracle on the CPU, same inputs
    w2 = w_last.clone().requires_grad_(True)
    f2 = feat.clone().requires_grad_(True)
    xrec2 = F.conv2d(f2, w2, padding=1)
    r0, rlog = lo.loss_forward(sd, qloss, x, xrec2, 0, 0, last_layer=w2, gate=gate, disc_weight_max=0.75,
                               train=(mode == "train"), budget=lo.budget_loss_dual)
    rgw, rgf = torch.autograd.grad(r0, [w2, f2])
    p = mode + "_"
    for got, ref, key, tol in ((l0, r0, "loss0", 2e-2), (log0["train_p_loss"], rlog["p_loss"], "log0_p_loss", 2e-2),
                               (log0["train_g_loss"], rlog["g_loss"], "log0_g_loss", 1e-2),
                               (log0["train_d_weight"], rlog["d_weight"], "log0_d_weight", 5e-2),
                               (log0["train_budget_loss"], rlog["budget_loss"], "log0_budget_loss", 1e-5)):
        assert abs(float(got) - float(ref)) <= tol * abs(float(ref)) + 1e-6, (key, float(got), float(ref))
        assert abs(float(got) - float(gold[p + key])) <= tol * abs(float(gold[p + key])) + 1e-6, (key, "golden")
    # the GAN half of these gradients passes through the piecewise-linear PatchGAN in bf16 (see the bound discussion in
    # test_patchgan_on_the_kernels_matches_oracle): direction and size are pinned, the RMS bound is 0.12
    for got_g, ref_g, key in ((gw, rgw, "g_w_last"), (gf, rgf, "g_feat")):
        for want in (ref_g, torch.from_numpy(gold[p + key])):
            assert rel_rms(got_g, want) < 0.12 and _cos(got_g, want) > 0.993, (key, rel_rms(got_g, want), _cos(got_g, want))
            assert abs(float(got_g.norm()) / float(want.norm()) - 1.0) < 3e-2, key
    # discriminator pass
    l1, log1 = loss(qd, xd, xrec.detach(), 1, 0, last_layer=wl, split="train")
    assert abs(float(l1) - float(gold[p + "loss1"])) <= 1e-2 * abs(float(gold[p + "loss1"])) + 1e-6
    dparams = dict(loss.discriminator.named_parameters())
    gds = torch.autograd.grad(l1, list(dparams.values()))
    for k, gr in zip(dparams, gds):
        ref = float(gold[p + "gd_norm_" + k])
        assert abs(float(gr.norm()) - ref) <= 2e-2 * ref + 1e-7, (k, float(gr.norm()), ref)
    if mode == "train":
        for k, v in loss.state_dict().items():
            if "running" in k:
                # bf16 convolutions feed these statistics; means of zero-centred activations are judged against the
                # channel's batch standard deviation (recovered from the golden running variance)
                ref_v = torch.from_numpy(gold[p + "bn1_" + k])
                if k.endswith("running_var"):
                    assert torch.allclose(v.cpu(), ref_v, rtol=2e-2), k
                else:
                    rv = torch.from_numpy(gold[p + "bn1_" + k.replace("running_mean", "running_var")])
                    assert float(((v.cpu() - ref_v).abs() / (rv.sqrt() + 1e-12)).max()) < 2e-3, k


def test_training_step_under_the_real_loss(monkeypatch):
    """Both optimizer passes of training_step (dqvae_dual_feat.py:88-119) with the reference's loss config on the
    reduced-width dual-grain model: finite losses, gradients where the reference has them."""
    monkeypatch.setenv("B200DQ_ALLOW_RANDOM_VGG", "1")
    from dynamicvectorquantization_b200 import configs
    cfg = configs.scaled_dual_config()
    cfg["params"]["lossconfig"] = configs.real_loss_config(cfg["params"]["lossconfig"]["params"]["budget_loss_config"])
    torch.manual_seed(0)
    model = configs.build_model(cfg).cuda().train()
    model.learning_rate, model.min_learning_rate = 1e-4, 1e-6          # what train.py:243-267 injects
    model.steps_per_epoch, model.training_steps, model.max_epoch = 10, 100, 10
    (opt_ae, opt_disc), scheds = model.configure_optimizers()
    assert len(scheds) == 2
    x = torch.rand(4, 3, 64, 64, device="cuda") * 2 - 1
    batch = {"image": x.permute(0, 2, 3, 1).contiguous()}
    for it in range(2):
        opt_ae.zero_grad(set_to_none=True)
        l0 = model.training_step(batch, it, 0)
        l0.backward()
        assert torch.isfinite(l0)
        assert model.decoder.conv_out.weight.grad is not None and model.encoder.conv_in.weight.grad is not None
        assert all(p.grad is None for p in model.loss.perceptual_loss.parameters())
        opt_ae.step()
        opt_disc.zero_grad(set_to_none=True)
        l1 = model.training_step(batch, it, 1)
        l1.backward()
        assert torch.isfinite(l1)
        assert all(p.grad is not None for p in model.loss.discriminator.parameters())
        opt_disc.step()


@pytest.mark.parametrize("c,hw", [(64, (9, 11)), (128, (16, 16)), (256, (7, 5)), (512, (4, 4)), (64, (64, 64))])
def test_lpips_head_kernel_matches_fp32_reference(c, hw):
    """csrc/lpips.cu against the reference's own formulation (normalize_tensor, squared difference, 1x1 lin, spatial
    mean; lpips.py:44-55,116-122) in fp32 on the same bf16-rounded features: values 1e-5, gradients (bf16 outputs) 1e-2."""
    from dynamicvectorquantization_b200 import kernels as kn
    g = torch.Generator().manual_seed(c + hw[0])
    n, (h, w_) = 3, hw
    f0 = torch.randn(n, h, w_, c, generator=g).clamp_min(0).to(BF)          # ReLU features (incl. all-zero pixels)
    f1 = torch.randn(n, h, w_, c, generator=g).clamp_min(0).to(BF)
    f0[0, 0, 0] = 0
    f1[1, 1, 1] = 0
    lw = torch.rand(c, generator=g) * 2 / c
    a = f0.float().requires_grad_(True)
    b = f1.float().requires_grad_(True)
    na = a / (a.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    nb = b / (b.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    ref = (((na - nb) ** 2) * lw).sum(-1).mean((1, 2))
    up = torch.randn(n, generator=g)
    (ref * up).sum().backward()
    got = kn.lpips_head_fwd(f0.cuda(), f1.cuda(), lw.cuda())
    assert torch.allclose(got.cpu(), ref.detach(), rtol=1e-5, atol=1e-7)
    d0, d1 = kn.lpips_head_bwd(f0.cuda(), f1.cuda(), lw.cuda(), up.cuda(), True, True)
    live0 = f0.float().pow(2).sum(-1) > 0                                     # |f| = 0: the reference divides by eps
    live1 = f1.float().pow(2).sum(-1) > 0
    assert rel_rms(d0.cpu()[live0], a.grad[live0]) < 1e-2
    assert rel_rms(d1.cpu()[live1], b.grad[live1]) < 1e-2
    only1 = kn.lpips_head_bwd(f0.cuda(), f1.cuda(), lw.cuda(), up.cuda(), False, True)
    assert only1[0] is None and torch.equal(only1[1], d1)
    # dropout: same mask forward and backward (finite-difference-free check: gradient is zero where the value's
    # mask is zero), keep rate ~ 1 - p, rescaled by 1 / (1 - p)
    seed = torch.tensor([1234567], dtype=torch.int64, device="cuda")
    ones = torch.ones(c, device="cuda")
    fa = torch.ones(n, h, w_, c).to(BF).cuda()
    fb = torch.zeros(n, h, w_, c).to(BF).cuda()
    v = kn.lpips_head_fwd(fa, fb, ones, seed, 0.5)                           # = mean_hw sum_c mask_c * 2 / C
    assert torch.allclose(v.cpu(), torch.ones(n), atol=6.0 / (c * h * w_) ** 0.5)
    seed2 = seed + 1
    r1 = kn.lpips_head_fwd(f0.cuda(), f1.cuda(), lw.cuda(), seed, 0.5)
    r2 = kn.lpips_head_fwd(f0.cuda(), f1.cuda(), lw.cuda(), seed2, 0.5)
    assert not torch.equal(r1, r2), "a different seed must draw a different mask"
    assert torch.equal(r1, kn.lpips_head_fwd(f0.cuda(), f1.cuda(), lw.cuda(), seed, 0.5)), "same seed, same mask"
    # backward uses the forward's mask: d val / d f1 summed against f1 ... (val is homogeneous of degree 0 in f1, so
    # instead check linearity in the upstream gradient and agreement of the masked value with a masked reference)
    dd0, dd1 = kn.lpips_head_bwd(f0.cuda(), f1.cuda(), lw.cuda(), up.cuda(), True, True, seed, 0.5)
    ee0, ee1 = kn.lpips_head_bwd(f0.cuda(), f1.cuda(), lw.cuda(), (2 * up).cuda(), True, True, seed, 0.5)
    assert rel_rms(ee1.float()[live1.cuda()], 2 * dd1.float()[live1.cuda()]) < 1e-2
