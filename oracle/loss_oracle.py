"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the stage-1 training loss of DQ-VAE.

A functional fp32 PyTorch restatement, on a flat state_dict, of what ``training_step`` evaluates
after the autoencoder forward:

* ``modules/losses/lpips.py``  (LPIPS: ScalingLayer :59-66, torchvision VGG16 feature slices
  :78-113, channel normalisation :116-118, 1x1 ``lin`` heads + spatial average :44-55,121-122);
* ``modules/discriminator/model.py:17-67`` (PatchGAN ``NLayerDiscriminator``: 4x4 convolutions,
  BatchNorm2d in training mode, LeakyReLU(0.2));
* ``modules/losses/vqperceptual_multidisc.py:23-26,102-194`` (hinge losses, adaptive discriminator
  weight from the two gradients w.r.t. the decoder's last layer, ``disc_weight_max`` clamp, codebook
  and budget terms; the discriminator pass on detached images).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

Pinning: the reference ships no test for this path.  ``tests/golden/make_golden.py::loss_goldens``
runs the reference's own classes in the build container (torchvision's VGG16 with seeded random
weights - the pretrained file cannot be downloaded offline - and seeded non-negative ``lin`` heads)
and stores inputs/outputs in ``tests/golden/loss_small.npz``; ``tests/test_oracle_golden.py`` pins
this file to them.
"""
import torch
import torch.nn.functional as F

# torchvision.models.vgg16().features: index of every convolution, grouped as the reference slices
# them (lpips.py:88-97); a 2x2 max-pool precedes every slice but the first.
VGG_SLICES = [[0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28]]
VGG_CHANNELS = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256),
                (256, 512), (512, 512), (512, 512), (512, 512), (512, 512), (512, 512)]
LPIPS_CHNS = [64, 128, 256, 512, 512]                                     # lpips.py:15
SHIFT = [-.030, -.088, -.188]                                             # lpips.py:62
SCALE = [.458, .448, .450]                                                # lpips.py:63


def lpips_shapes(prefix="loss.perceptual_loss"):
    shapes, ci = {}, 0
    for s, convs in enumerate(VGG_SLICES):
        for idx in convs:
            cin, cout = VGG_CHANNELS[ci]
            ci += 1
            shapes[f"{prefix}.net.slice{s + 1}.{idx}.weight"] = (cout, cin, 3, 3)
            shapes[f"{prefix}.net.slice{s + 1}.{idx}.bias"] = (cout,)
    for k, c in enumerate(LPIPS_CHNS):
        shapes[f"{prefix}.lin{k}.model.1.weight"] = (1, c, 1, 1)           # model = [Dropout, Conv2d] (lpips.py:72-75)
    return shapes


def disc_shapes(input_nc=3, ndf=64, n_layers=3, prefix="loss.discriminator"):
    """State-dict layout of NLayerDiscriminator with BatchNorm (discriminator/model.py:37-62)."""
    shapes = {f"{prefix}.main.0.weight": (ndf, input_nc, 4, 4), f"{prefix}.main.0.bias": (ndf,)}
    idx, mult = 2, 1
    for n in range(1, n_layers + 1):
        prev, mult = mult, min(2 ** n, 8)
        shapes[f"{prefix}.main.{idx}.weight"] = (ndf * mult, ndf * prev, 4, 4)     # bias=False before BatchNorm (:30-33)
        for nm in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"{prefix}.main.{idx + 1}.{nm}"] = (ndf * mult,)
        idx += 3
    shapes[f"{prefix}.main.{idx}.weight"] = (1, ndf * mult, 4, 4)
    shapes[f"{prefix}.main.{idx}.bias"] = (1,)
    return shapes


def make_loss_weights(seed=0, ndf=64, n_layers=3):
    """Deterministic weights for the loss modules, independent of module construction order:
    He-scaled VGG convolutions (activations keep O(1) scale through 13 layers), non-negative lin heads
    (like the trained LPIPS heads), discriminator as weights_init leaves it (N(0,0.02) convolutions,
    BatchNorm weight N(1,0.02), bias 0; discriminator/model.py:8-14)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in lpips_shapes().items():
        if ".net." in k:
            if k.endswith("weight"):
                fan_in = shp[1] * 9
                sd[k] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
            else:
                sd[k] = torch.randn(shp, generator=g) * 0.05
        else:
            sd[k] = torch.rand(shp, generator=g) * (2.0 / shp[1])
    for k, shp in disc_shapes(ndf=ndf, n_layers=n_layers).items():
        if k.endswith("running_mean"):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shp)
        elif len(shp) == 4:
            sd[k] = torch.randn(shp, generator=g) * 0.02
        elif k.endswith("main.0.bias") or (k.endswith("bias") and shp == (1,)):
            sd[k] = torch.randn(shp, generator=g) * 0.02
        elif k.endswith("weight"):
            sd[k] = 1.0 + torch.randn(shp, generator=g) * 0.02
        else:
            sd[k] = torch.zeros(shp)
    return sd


# ------------------------------------------------------------------------------------- LPIPS
def vgg_features(sd, x, prefix="loss.perceptual_loss"):
    """lpips.py:99-113: relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 of torchvision's VGG16."""
    outs, h = [], x
    for s, convs in enumerate(VGG_SLICES):
        if s > 0:
            h = F.max_pool2d(h, 2, 2)
        for idx in convs:
            h = F.relu(F.conv2d(h, sd[f"{prefix}.net.slice{s + 1}.{idx}.weight"],
                                sd[f"{prefix}.net.slice{s + 1}.{idx}.bias"], padding=1))
        outs.append(h)
    return outs


def normalize_tensor(x, eps=1e-10):
    """lpips.py:116-118."""
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def lpips(sd, inp, target, prefix="loss.perceptual_loss"):
    """lpips.py:44-56 (eval mode: Dropout is the identity).  Returns [B,1,1,1]."""
    shift = torch.tensor(SHIFT, dtype=inp.dtype, device=inp.device)[None, :, None, None]
    scale = torch.tensor(SCALE, dtype=inp.dtype, device=inp.device)[None, :, None, None]
    f0 = vgg_features(sd, (inp - shift) / scale, prefix)
    f1 = vgg_features(sd, (target - shift) / scale, prefix)
    val = 0
    for k in range(5):
        d = (normalize_tensor(f0[k]) - normalize_tensor(f1[k])) ** 2
        val = val + F.conv2d(d, sd[f"{prefix}.lin{k}.model.1.weight"]).mean([2, 3], keepdim=True)
    return val


# ------------------------------------------------------------------------------ discriminator
def discriminator(sd, x, n_layers=3, train=True, prefix="loss.discriminator", new_stats=None):
    """discriminator/model.py:37-67.  BatchNorm2d in training mode normalises with the batch's biased
    variance; when ``new_stats`` is a dict, the updated running statistics (momentum 0.1, unbiased
    variance) are written into it - the oracle itself never mutates ``sd``."""
    h = F.leaky_relu(F.conv2d(x, sd[f"{prefix}.main.0.weight"], sd[f"{prefix}.main.0.bias"], stride=2, padding=1), 0.2)
    idx = 2
    for n in range(1, n_layers + 1):
        stride = 2 if n < n_layers else 1
        h = F.conv2d(h, sd[f"{prefix}.main.{idx}.weight"], None, stride=stride, padding=1)
        bn = f"{prefix}.main.{idx + 1}"
        if train:
            mean = h.mean([0, 2, 3])
            var = h.var([0, 2, 3], unbiased=False)
            if new_stats is not None:
                cnt = h.numel() / h.shape[1]
                new_stats[bn + ".running_mean"] = 0.9 * sd[bn + ".running_mean"] + 0.1 * mean.detach()
                new_stats[bn + ".running_var"] = 0.9 * sd[bn + ".running_var"] + 0.1 * var.detach() * cnt / (cnt - 1)
        else:
            mean, var = sd[bn + ".running_mean"], sd[bn + ".running_var"]
        h = (h - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + 1e-5)
        h = h * sd[bn + ".weight"][None, :, None, None] + sd[bn + ".bias"][None, :, None, None]
        h = F.leaky_relu(h, 0.2)
        idx += 3
    return F.conv2d(h, sd[f"{prefix}.main.{idx}.weight"], sd[f"{prefix}.main.{idx}.bias"], stride=1, padding=1)


def hinge_d_loss(logits_real, logits_fake):
    """vqperceptual_multidisc.py:23-27."""
    return 0.5 * (torch.mean(F.relu(1. - logits_real)) + torch.mean(F.relu(1. + logits_fake)))


def budget_loss_dual(gate, target_ratio=0.5, gamma=10.0, min_grain=16, max_grain=32):
    """budget.py:15-28 with calculate_all=True (returns the 'last' term twice, :26)."""
    b = gate.shape[0]
    beta = 1.0 * gate[:, 0].sum() + 4.0 * gate[:, 1].sum()
    ratio = (beta / b - min_grain ** 2) / (max_grain ** 2 - min_grain ** 2)
    last = gamma * (1 - ratio - (1 - target_ratio)) ** 2
    return last + last


# --------------------------------------------------------------------------------- the loss
def loss_forward(sd, codebook_loss, inputs, reconstructions, optimizer_idx, global_step, last_layer=None,
                 gate=None, disc_start=0, codebook_weight=1.0, disc_factor=1.0, disc_weight=1.0,
                 perceptual_weight=1.0, disc_weight_max=None, n_layers=3, train=True, budget=None, new_stats=None):
    """vqperceptual_multidisc.py:115-194 (hinge GAN loss, adaptive weight).  Returns (loss, log dict).

    optimizer_idx 0: nll = mean(|x - xrec| + w_p * lpips) (:116-124); g = -mean(D(xrec)) (:135);
      d_weight = clamp(|d nll / d last_layer| / (|d g / d last_layer| + 1e-4), 0, 1e4) * disc_weight,
      clamped to disc_weight_max (:102-113,137-144); loss = nll + d_weight * disc_factor * g +
      codebook_weight * mean(codebook_loss) [+ budget(gate)] (:148-153).
    optimizer_idx 1: disc_factor * hinge(D(x.detach()), D(xrec.detach())) (:178-187).
    disc_factor is zeroed while global_step < disc_start (:16-19)."""
    rec = torch.abs(inputs - reconstructions)
    if perceptual_weight > 0:
        p_loss = lpips(sd, inputs, reconstructions)
        rec = rec + perceptual_weight * p_loss
    else:
        p_loss = torch.zeros(1)
    nll = rec.mean()
    factor = disc_factor if global_step >= disc_start else 0.0
    if optimizer_idx == 0:
        logits_fake = discriminator(sd, reconstructions, n_layers, train, new_stats=new_stats)
        g_loss = -logits_fake.mean()
        if last_layer is not None and last_layer.requires_grad:
            nll_g = torch.autograd.grad(nll, last_layer, retain_graph=True)[0]
            g_g = torch.autograd.grad(g_loss, last_layer, retain_graph=True)[0]
            d_weight = (nll_g.norm() / (g_g.norm() + 1e-4)).clamp(0.0, 1e4).detach() * disc_weight
        else:
            d_weight = torch.tensor(0.0)
        if disc_weight_max is not None:
            d_weight = d_weight.clamp(max=disc_weight_max)
        loss = nll + d_weight * factor * g_loss + codebook_weight * codebook_loss.mean()
        log = {"nll_loss": nll.detach(), "rec_loss": rec.detach().mean(), "p_loss": p_loss.detach().mean(),
               "d_weight": d_weight.detach(), "g_loss": g_loss.detach(), "quant_loss": codebook_loss.detach().mean()}
        if gate is not None and budget is not None:
            bl = budget(gate)
            loss = loss + bl
            log["budget_loss"] = bl.detach()
        log["total_loss"] = loss.detach()
        return loss, log
    logits_real = discriminator(sd, inputs.detach(), n_layers, train, new_stats=new_stats)
    # the fake pass sees the statistics the real pass left behind only through running_* (unused in
    # training-mode normalisation), so the two passes are independent here
    logits_fake = discriminator(sd, reconstructions.detach(), n_layers, train)
    d_loss = factor * hinge_d_loss(logits_real, logits_fake)
    return d_loss, {"disc_loss": d_loss.detach(), "logits_real": logits_real.detach().mean(),
                    "logits_fake": logits_fake.detach().mean()}


def toy_inputs(seed=11, b=2, res=64, cfeat=8):
    """Seeded stand-in for "decoder output as a function of its last layer": xrec = conv3x3(feat, w_last).
    Shared by tests/golden/make_golden.py::loss_goldens and the tests, so the fixture stores outputs only."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 3, res, res, generator=g) * 2 - 1
    feat = torch.randn(b, cfeat, res, res, generator=g) * 0.5
    w_last = torch.randn(3, cfeat, 3, 3, generator=g) * 0.1
    qloss = torch.rand((), generator=g) * 0.1
    fine = (torch.rand(b, 1, 4, 4, generator=g) > 0.4).float()
    gate = torch.cat([1 - fine, fine], 1)
    return x, feat, w_last, qloss, gate
