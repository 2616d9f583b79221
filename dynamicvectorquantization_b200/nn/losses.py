"""Stage-1 training loss: L1 + LPIPS + adaptive-weight PatchGAN + codebook + budget terms.

Public surface of ``modules/losses/vqperceptual_multidisc.py`` (reference): the GAN loss helpers (:17-47) and
``VQLPIPSWithDiscriminator`` (:50-194) - same constructor arguments, same attribute / sub-module names
(``perceptual_loss``, ``discriminator``, ``budget_loss`` ...), same call contract

    loss, log = module(codebook_loss, inputs, reconstructions, optimizer_idx, global_step,
                       last_layer=None, cond=None, split="train", gate=None)

and the same keys in the returned log dict.  The body is organised around the two passes Lightning drives
(``optimizer_idx`` 0 = autoencoder, 1 = discriminator) instead of one long branch:

* ``_autoencoder_pass``: nll = mean(|x - xrec| + w_p * LPIPS(x, xrec)) (:116-124), g = GAN loss of D(xrec)
  (:127-135), adaptive weight |d nll / d last_layer| / (|d g / d last_layer| + 1e-4) clamped to [0, 1e4] and to
  ``disc_weight_max`` (:102-113,137-146), total = nll + d_weight * disc_factor * g + codebook_weight *
  mean(codebook_loss) [+ budget(gate)] (:148-153);
* ``_discriminator_pass``: disc_factor * GAN loss of D(x.detach()), D(xrec.detach()) (:178-187).  The reference
  evaluates L1 + LPIPS before branching and discards them in this pass; they are not computed here.

The perceptual term runs on the tensor-core kernels (``nn/lpips.py``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .discriminator import weights_init
from .lpips import LPIPS

try:
    from utils.utils import instantiate_from_config
except Exception:
    from ..config import instantiate_from_config


class DummyLoss(nn.Module):
    """Placeholder loss (vqperceptual_multidisc.py:13-15)."""


def adopt_weight(weight, global_step, threshold=0, value=0.):
    """`value` until `threshold` steps have passed, `weight` afterwards (:17-20)."""
    return value if global_step < threshold else weight


def log(t, eps=1e-10):
    return (t + eps).log()


# ---- GAN objectives (:25-47); *_d_* take (logits_real, logits_fake), *_g_* / *_gen_* take logits_fake
def hinge_d_loss(logits_real, logits_fake):
    real_term = F.relu(1. - logits_real).mean()
    fake_term = F.relu(1. + logits_fake).mean()
    return 0.5 * (real_term + fake_term)


def hinge_g_loss(logits_fake):
    return -logits_fake.mean()


def vanilla_d_loss(logits_real, logits_fake):
    return 0.5 * (F.softplus(-logits_real).mean() + F.softplus(logits_fake).mean())


def bce_discr_loss(logits_real, logits_fake):
    return (-log(1 - logits_fake.sigmoid()) - log(logits_real.sigmoid())).mean()


def bce_gen_loss(logits_fake):
    return -log(logits_fake.sigmoid()).mean()


_GAN_OBJECTIVES = {                      # name -> (discriminator objective, generator objective)   (:78-88)
    "hinge": (hinge_d_loss, hinge_g_loss),
    "vanilla": (vanilla_d_loss, hinge_g_loss),
    "bce": (bce_discr_loss, bce_gen_loss),
}


class VQLPIPSWithDiscriminator(nn.Module):
    def __init__(self, disc_start, disc_config, disc_init, codebook_weight=1.0, pixelloss_weight=1.0,
                 disc_factor=1.0, disc_weight=1.0, perceptual_weight=1.0, disc_conditional=False,
                 disc_adaptive_loss=True, disc_loss="hinge", disc_weight_max=None, budget_loss_config=None):
        super().__init__()
        if disc_loss not in _GAN_OBJECTIVES:
            raise AssertionError(f"Unknown GAN loss '{disc_loss}'.")
        self.disc_loss, self.gen_loss = _GAN_OBJECTIVES[disc_loss]
        # reconstruction / perceptual term
        self.pixel_weight = pixelloss_weight
        self.perceptual_weight = perceptual_weight
        self.perceptual_loss = LPIPS().eval()
        self.codebook_weight = codebook_weight
        # adversarial term
        self.discriminator = instantiate_from_config(disc_config)
        if disc_init:
            self.discriminator = self.discriminator.apply(weights_init)
        self.discriminator_iter_start = disc_start
        self.disc_factor = disc_factor
        self.discriminator_weight = disc_weight
        self.disc_weight_max = disc_weight_max
        self.disc_adaptive_loss = disc_adaptive_loss
        self.disc_conditional = disc_conditional
        # grain budget term
        self.budget_loss_config = budget_loss_config
        if budget_loss_config is not None:
            self.budget_loss = instantiate_from_config(budget_loss_config)
        print(f"VQLPIPSWithDiscriminator running with {disc_loss} loss.")

    # ------------------------------------------------------------------ pieces
    def calculate_adaptive_weight(self, nll_loss, g_loss, last_layer=None):
        """Balance of the two gradients that reach the decoder's last layer (:102-113)."""
        layer = last_layer if last_layer is not None else self.last_layer[0]
        nll_grads, g_grads = (torch.autograd.grad(term, layer, retain_graph=True)[0] for term in (nll_loss, g_loss))
        ratio = nll_grads.norm() / (g_grads.norm() + 1e-4)
        return ratio.clamp(0.0, 1e4).detach() * self.discriminator_weight

    def _logits(self, images, cond):
        if cond is None:
            assert not self.disc_conditional
            return self.discriminator(images)
        assert self.disc_conditional
        return self.discriminator(torch.cat((images, cond), dim=1))

    def _disc_factor(self, global_step):
        return adopt_weight(self.disc_factor, global_step, threshold=self.discriminator_iter_start)

    # ------------------------------------------------------------------ optimizer_idx == 0
    def _autoencoder_pass(self, codebook_loss, inputs, reconstructions, global_step, last_layer, cond, split, gate):
        inputs, reconstructions = inputs.contiguous(), reconstructions.contiguous()
        rec_loss = (inputs - reconstructions).abs()
        if self.perceptual_weight > 0:
            p_loss = self.perceptual_loss(inputs, reconstructions)
            rec_loss = rec_loss + self.perceptual_weight * p_loss
        else:
            p_loss = torch.tensor([0.0])
        nll_loss = rec_loss.mean()
        g_loss = self.gen_loss(self._logits(reconstructions, cond))

        if self.disc_adaptive_loss:
            try:
                d_weight = self.calculate_adaptive_weight(nll_loss, g_loss, last_layer=last_layer)
            except RuntimeError:                      # no graph to differentiate (evaluation)
                assert not self.training
                d_weight = torch.tensor(0.0)
            if self.disc_weight_max is not None:
                d_weight.clamp_max_(self.disc_weight_max)
        else:
            d_weight = torch.tensor(self.disc_weight_max)

        disc_factor = self._disc_factor(global_step)
        quant_loss = codebook_loss.mean()
        loss = nll_loss + d_weight * disc_factor * g_loss + self.codebook_weight * quant_loss
        terms = dict(quant_loss=quant_loss.detach(), nll_loss=nll_loss.detach().mean(),
                     rec_loss=rec_loss.detach().mean(), p_loss=p_loss.detach().mean(), d_weight=d_weight.detach(),
                     disc_factor=torch.tensor(disc_factor), g_loss=g_loss.detach().mean())
        if gate is not None and self.budget_loss_config is not None:
            budget_loss = self.budget_loss(gate=gate)
            loss = loss + budget_loss
            terms["budget_loss"] = budget_loss.detach().mean()
        terms["total_loss"] = loss.clone().detach().mean()
        return loss, {f"{split}_{name}": value for name, value in terms.items()}

    # ------------------------------------------------------------------ optimizer_idx == 1
    def _discriminator_pass(self, inputs, reconstructions, global_step, cond, split):
        logits_real = self._logits(inputs.contiguous().detach(), cond)
        logits_fake = self._logits(reconstructions.contiguous().detach(), cond)
        d_loss = self._disc_factor(global_step) * self.disc_loss(logits_real, logits_fake)
        terms = dict(disc_loss=d_loss.clone().detach().mean(), logits_real=logits_real.detach().mean(),
                     logits_fake=logits_fake.detach().mean())
        return d_loss, {f"{split}_{name}": value for name, value in terms.items()}

    def forward(self, codebook_loss, inputs, reconstructions, optimizer_idx, global_step, last_layer=None,
                cond=None, split="train", gate=None):
        if optimizer_idx == 0:
            return self._autoencoder_pass(codebook_loss, inputs, reconstructions, global_step, last_layer, cond,
                                          split, gate)
        if optimizer_idx == 1:
            return self._discriminator_pass(inputs, reconstructions, global_step, cond, split)
