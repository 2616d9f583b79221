"""Stage-1 DQ-VAE models (dual-grain feature / entropy routed, triple-grain) on the sm_100a path.

Mirrors ``models/stage1_dynamic/dqvae_dual_feat.py`` (:8-192), ``dqvae_dual_entropy.py`` (:13-261)
and ``dqvae_triple_feat.py`` of the reference: constructor arguments, sub-module names
(``encoder, decoder, loss, quantize, quant_conv, post_quant_conv``), ``encode / decode / forward``
return structures, ``training_step(batch, batch_idx, optimizer_idx)``, ``validation_step``,
``configure_optimizers``, ``get_last_layer``, ``log_images``, ``get_code_emb_with_depth``,
``init_from_ckpt``.  ``forward`` chains encoder head -> quant_conv -> VQ -> post_quant_conv ->
decoder in NHWC bf16 without leaving the C-ABI kernels.
"""
import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import kernels as kn
from .. import ops
from .quantize import VectorQuantize2

try:
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    _HAVE_PL = True
except Exception:
    _HAVE_PL = False

    class _Base(nn.Module):
        """Minimal stand-in for LightningModule when pytorch_lightning is not installed."""
        current_epoch = 0
        global_step = 0

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

try:
    from utils.utils import instantiate_from_config
except Exception:
    from ..config import instantiate_from_config


# ---- LR lambdas (models/stage1/utils.py:6-24)
def _fn_linear_warmup(warmup_steps, step):
    return float(step) / float(max(1, warmup_steps)) if step < warmup_steps else 1.0


def _fn_linear_warmup_cosine(warmup_steps, max_steps, multipler_min, step):
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    m = 0.5 * (math.cos((step - warmup_steps) / (max_steps - warmup_steps) * math.pi) + 1)
    return max(m, multipler_min)


def Scheduler_LinearWarmup(warmup_steps):
    return partial(_fn_linear_warmup, warmup_steps)


def Scheduler_LinearWarmup_CosineDecay(warmup_steps, max_steps, multipler_min):
    return partial(_fn_linear_warmup_cosine, warmup_steps, max_steps, multipler_min)


class Entropy(nn.Sequential):
    """Per-patch grey-level entropy (dqvae_dual_entropy.py:13-63): soft histogram with 32 bins on
    [-1,1], sigma 0.01, eps 1e-40.  forward() runs the fused CUDA kernel (csrc/entropy.cu: one read of
    the image, no [B*patches, pixels, bins] intermediate); entropy() keeps the reference's helper
    signature for callers that bring their own flattened patches."""

    def __init__(self, patch_size, image_width, image_height):
        super().__init__()
        self.width = image_width
        self.height = image_height
        self.psize = patch_size
        self.patch_num = int(self.width * self.height / self.psize ** 2)
        self.hw = int(self.width // self.psize)
        self.unfold = torch.nn.Unfold(kernel_size=(self.psize, self.psize), stride=self.psize)
        self._bins = None

    def entropy(self, values, bins, sigma, batch):
        epsilon = 1e-40
        residuals = values.unsqueeze(2) - bins.unsqueeze(0).unsqueeze(0)
        kernel_values = torch.exp(-0.5 * (residuals / sigma).pow(2))
        pdf = torch.mean(kernel_values, dim=1)
        normalization = torch.sum(pdf, dim=1).unsqueeze(1) + epsilon
        pdf = pdf / normalization + epsilon
        ent = -torch.sum(pdf * torch.log(pdf), dim=1)
        return ent.reshape(batch, self.hw, self.hw)

    @torch.no_grad()
    def forward(self, inputs):
        if not inputs.is_cuda:
            raise RuntimeError("Entropy (B200) needs CUDA tensors; there is no CPU fallback")
        assert inputs.shape[-2] == self.height and inputs.shape[-1] == self.width
        if self._bins is None or self._bins.device != inputs.device:
            self._bins = torch.linspace(-1, 1, 32).to(device=inputs.device)
        return kn.patch_entropy(inputs.float(), self._bins, self.psize, 0.01)


def _disabled_train(self, mode=True):
    return self


class _GrainVQModelBase(_Base):
    _h_key = "h_dual"
    _uses_entropy = False

    def _init_common(self, encoderconfig, decoderconfig, lossconfig, vqconfig, quant_before_dim,
                     quant_after_dim, quant_sample_temperature, ckpt_path, ignore_keys, image_key,
                     monitor, warmup_epochs, loss_with_epoch, scheduler_type):
        self.image_key = image_key
        self.encoder = instantiate_from_config(encoderconfig)
        self.decoder = instantiate_from_config(decoderconfig)
        self.loss = instantiate_from_config(lossconfig)
        self.quantize = instantiate_from_config(vqconfig)
        self.quant_conv = torch.nn.Conv2d(quant_before_dim, quant_after_dim, 1)
        self.post_quant_conv = torch.nn.Conv2d(quant_after_dim, quant_before_dim, 1)
        self.quant_sample_temperature = quant_sample_temperature
        for name in ("encoder", "decoder", "quantize"):
            mod = type(getattr(self, name)).__module__
            if not mod.startswith("dynamicvectorquantization_b200"):
                import warnings
                warnings.warn(f"{name} resolved to {mod}, not to the B200 overlay: the reference tree is ahead of the "
                              f"overlay on sys.path (start the job with `python -m dynamicvectorquantization_b200.launch "
                              f"<script> ...`) or the config names a class this package does not provide")

    def _finish_init(self, ckpt_path, ignore_keys, monitor, warmup_epochs, loss_with_epoch, scheduler_type):
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)
        if monitor is not None:
            self.monitor = monitor
        self.warmup_epochs = warmup_epochs
        self.loss_with_epoch = loss_with_epoch
        self.scheduler_type = scheduler_type

    def init_from_ckpt(self, path, ignore_keys=list()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            for ik in ignore_keys:
                if k.startswith(ik):
                    print("Deleting key {} from state_dict.".format(k))
                    del sd[k]
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    # ------------------------------------------------------------------ hot path
    def _quantize_nhwc(self, h_nchw, codebook_mask):
        """h (NCHW fp32, small) -> quant_conv -> VQ.  Returns (quant NHWC bf16, loss, info)."""
        if isinstance(self.quantize, VectorQuantize2) and self.quantize.accept_image_fmap:
            hq = ops.conv2d(ops.to_nhwc(h_nchw), self.quant_conv)
            b, hh, ww, c = hq.shape
            mask_rows = None if codebook_mask is None else codebook_mask.reshape(-1).float().contiguous()
            xq, loss, codes = self.quantize.forward_rows(hq.view(-1, c), mask_rows)
            return xq.view(b, hh, ww, c), loss, (None, None, codes.view(b, hh, ww))
        # any other quantizer class: go through its own NCHW fp32 interface
        hq = ops.to_nchw(ops.conv2d(ops.to_nhwc(h_nchw), self.quant_conv))
        quant, loss, info = self.quantize(x=hq, temp=self.quant_sample_temperature, codebook_mask=codebook_mask)
        return ops.to_nhwc(quant), loss, info

    def _decode_nhwc(self, quant_nhwc):
        dec = self.decoder
        if hasattr(dec, "position_bias_nchw"):
            b, hh, ww, c = quant_nhwc.shape
            pos = dec.position_bias_nchw(hh, ww, quant_nhwc.device).permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
            h = ops.Conv2dFn.apply(quant_nhwc, self.post_quant_conv.weight, self.post_quant_conv.bias, pos, 1, 1)
            return ops.to_nchw(dec.forward_nhwc(h))
        q = ops.to_nchw(ops.conv2d(quant_nhwc, self.post_quant_conv))
        return dec(q, None)

    def _encode_impl(self, x):
        x_entropy = self.entropy_calculation(x) if self._uses_entropy else None
        h_dict = self.encoder(x, x_entropy)
        quant, emb_loss, info = self._quantize_nhwc(h_dict[self._h_key], h_dict["codebook_mask"])
        return quant, emb_loss, info, h_dict["indices"], h_dict["gate"], x_entropy

    def encode(self, x):
        quant, emb_loss, info, grain_indices, gate, x_entropy = self._encode_impl(x)
        out = (ops.to_nchw(quant), emb_loss, info, grain_indices, gate)
        return out + (x_entropy,) if self._uses_entropy else out

    def decode(self, quant, grain_indices=None):
        return self._decode_nhwc(ops.to_nhwc(quant))

    def forward(self, input):
        ops.prepack(self)                      # all weight packings an optimizer step made stale, in one launch
        quant, diff, _, grain_indices, gate, x_entropy = self._encode_impl(input)
        dec = self._decode_nhwc(quant)
        out = (dec, diff, grain_indices, gate)
        return out + (x_entropy,) if self._uses_entropy else out

    # ------------------------------------------------------------------ training-loop surface
    def get_input(self, batch, k):
        x = batch[k]
        if len(x.shape) == 3:
            x = x[..., None]
        if x.size(1) != 3:
            x = x.permute(0, 3, 1, 2).to(memory_format=torch.contiguous_format).float()
        return x

    def _step_arg(self):
        return self.current_epoch if self.loss_with_epoch else self.global_step

    def training_step(self, batch, batch_idx, optimizer_idx):
        x = self.get_input(batch, self.image_key)
        xrec, qloss, indices, gate = self(x)[:4]
        ratio = indices.sum() / (indices.size(0) * indices.size(1) * indices.size(2))
        if optimizer_idx == 0:
            aeloss, log_dict_ae = self.loss(qloss, x, xrec, optimizer_idx, self._step_arg(),
                                            last_layer=self.get_last_layer(), split="train", gate=gate)
            self.log("train_aeloss", aeloss, prog_bar=False, logger=True, on_step=True, on_epoch=True)
            self.log("train_fine_ratio", ratio, prog_bar=True, logger=True, on_step=True, on_epoch=True)
            rec_loss = log_dict_ae["train_rec_loss"]
            self.log("train_rec_loss", rec_loss, prog_bar=True, logger=True, on_step=True, on_epoch=True, sync_dist=True)
            del log_dict_ae["train_rec_loss"]
            self.log_dict(log_dict_ae, prog_bar=False, logger=True, on_step=True, on_epoch=True)
            return aeloss
        if optimizer_idx == 1:
            discloss, log_dict_disc = self.loss(qloss, x, xrec, optimizer_idx, self._step_arg(),
                                                last_layer=self.get_last_layer(), split="train")
            self.log("train_discloss", discloss, prog_bar=False, logger=True, on_step=True, on_epoch=True)
            self.log_dict(log_dict_disc, prog_bar=False, logger=True, on_step=True, on_epoch=True)
            return discloss

    def validation_step(self, batch, batch_idx):
        x = self.get_input(batch, self.image_key)
        xrec, qloss, indices, gate = self(x)[:4]
        ratio = indices.sum() / (indices.size(0) * indices.size(1) * indices.size(2))
        self.log("val_fine_ratio", ratio, prog_bar=True, logger=True, on_step=True, on_epoch=True)
        aeloss, log_dict_ae = self.loss(qloss, x, xrec, 0, self._step_arg(), last_layer=self.get_last_layer(),
                                        split="val", gate=gate)
        discloss, log_dict_disc = self.loss(qloss, x, xrec, 1, self._step_arg(),
                                            last_layer=self.get_last_layer(), split="val")
        rec_loss = log_dict_ae["val_rec_loss"]
        self.log("val_rec_loss", rec_loss, prog_bar=True, logger=True, on_step=True, on_epoch=True, sync_dist=True)
        del log_dict_ae["val_rec_loss"]
        self.log("val_aeloss", aeloss, prog_bar=False, logger=True, on_step=True, on_epoch=True, sync_dist=True)
        self.log_dict(log_dict_ae)
        self.log_dict(log_dict_disc)
        return self.log_dict

    def configure_optimizers(self):
        lr = self.learning_rate
        opt_ae = torch.optim.Adam(list(self.encoder.parameters()) + list(self.decoder.parameters()) +
                                  list(self.quantize.parameters()) + list(self.quant_conv.parameters()) +
                                  list(self.post_quant_conv.parameters()), lr=lr, betas=(0.5, 0.9))
        opt_disc = torch.optim.Adam(self.loss.discriminator.parameters(), lr=lr, betas=(0.5, 0.9))
        warmup_steps = self.steps_per_epoch * self.warmup_epochs
        if self.scheduler_type == "linear-warmup":
            mk = lambda opt: torch.optim.lr_scheduler.LambdaLR(opt, Scheduler_LinearWarmup(warmup_steps))
        elif self.scheduler_type == "linear-warmup_cosine-decay":
            mmin = self.min_learning_rate / self.learning_rate
            mk = lambda opt: torch.optim.lr_scheduler.LambdaLR(
                opt, Scheduler_LinearWarmup_CosineDecay(warmup_steps=warmup_steps, max_steps=self.training_steps,
                                                        multipler_min=mmin))
        else:
            raise NotImplementedError()
        scheds = [{"scheduler": mk(o), "interval": "step", "frequency": 1} for o in (opt_ae, opt_disc)]
        return [opt_ae, opt_disc], scheds

    def get_last_layer(self):
        try:
            return self.decoder.conv_out.weight
        except AttributeError:
            return self.decoder.last_layer

    def log_images(self, batch, **kwargs):
        from modules.dynamic_modules.utils import draw_dual_grain_256res_color  # reference helper
        log = dict()
        x = self.get_input(batch, self.image_key).to(self.device)
        out = self(x)
        log["inputs"] = x
        log["reconstructions"] = out[0]
        log["grain_color"] = draw_dual_grain_256res_color(images=x.clone(), indices=out[2], scaler=0.7)
        return log

    def get_code_emb_with_depth(self, code):
        """dqvae_dual_feat.py:191-192 / dqvae_triple_feat.py:217-218 delegate to the quantizer's
        embed_code_with_depth (a (tensor, None)-style return for the quantizers that define it); the entropy model
        (dqvae_dual_entropy.py:258-262) uses get_codebook_entry.  quantize2_mask.VectorQuantize2 - the quantizer of
        every stage-1 config - has no embed_code_with_depth (the reference raises AttributeError there): fall back
        to get_codebook_entry instead of failing."""
        if not self._uses_entropy and hasattr(self.quantize, "embed_code_with_depth"):
            return self.quantize.embed_code_with_depth(code)
        return self.quantize.get_codebook_entry(code)


class DualGrainVQModel(_GrainVQModelBase):
    """models/stage1_dynamic/dqvae_dual_feat.py:8-46."""

    def __init__(self, encoderconfig, decoderconfig, lossconfig, vqconfig, quant_before_dim, quant_after_dim,
                 quant_sample_temperature=0., ckpt_path=None, ignore_keys=[], image_key="image", monitor=None,
                 warmup_epochs=0, loss_with_epoch=True, scheduler_type="linear-warmup_cosine-decay"):
        super().__init__()
        self._init_common(encoderconfig, decoderconfig, lossconfig, vqconfig, quant_before_dim, quant_after_dim,
                          quant_sample_temperature, ckpt_path, ignore_keys, image_key, monitor, warmup_epochs,
                          loss_with_epoch, scheduler_type)
        self._finish_init(ckpt_path, ignore_keys, monitor, warmup_epochs, loss_with_epoch, scheduler_type)


class DualGrainEntropyVQModel(_GrainVQModelBase):
    """models/stage1_dynamic/dqvae_dual_entropy.py:66-113 (class name there: DualGrainVQModel)."""
    _uses_entropy = True

    def __init__(self, encoderconfig, decoderconfig, lossconfig, vqconfig, quant_before_dim, quant_after_dim,
                 quant_sample_temperature=0., ckpt_path=None, ignore_keys=[], image_key="image", monitor=None,
                 warmup_epochs=0, loss_with_epoch=True, scheduler_type="linear-warmup_cosine-decay",
                 entropy_patch_size=16, image_size=256):
        super().__init__()
        self._init_common(encoderconfig, decoderconfig, lossconfig, vqconfig, quant_before_dim, quant_after_dim,
                          quant_sample_temperature, ckpt_path, ignore_keys, image_key, monitor, warmup_epochs,
                          loss_with_epoch, scheduler_type)
        self.entropy_patch_size = entropy_patch_size
        self.image_size = image_size
        self.entropy_calculation = Entropy(entropy_patch_size, image_size, image_size).eval()
        self.entropy_calculation.train = _disabled_train.__get__(self.entropy_calculation)
        self._finish_init(ckpt_path, ignore_keys, monitor, warmup_epochs, loss_with_epoch, scheduler_type)


class TripleGrainVQModel(_GrainVQModelBase):
    """models/stage1_dynamic/dqvae_triple_feat.py."""
    _h_key = "h_triple"

    def __init__(self, encoderconfig, decoderconfig, lossconfig, vqconfig, quant_before_dim, quant_after_dim,
                 quant_sample_temperature=0., ckpt_path=None, ignore_keys=[], image_key="image", monitor=None,
                 warmup_epochs=0, loss_with_epoch=True, scheduler_type="linear-warmup_cosine-decay"):
        super().__init__()
        self._init_common(encoderconfig, decoderconfig, lossconfig, vqconfig, quant_before_dim, quant_after_dim,
                          quant_sample_temperature, ckpt_path, ignore_keys, image_key, monitor, warmup_epochs,
                          loss_with_epoch, scheduler_type)
        self._finish_init(ckpt_path, ignore_keys, monitor, warmup_epochs, loss_with_epoch, scheduler_type)

    def decode_code(self, code_b):
        return self.decode(self.quantize.get_codebook_entry(code_b).permute(0, 3, 1, 2))


class SurrogateAELoss(nn.Module):
    """Reconstruction surrogate used by tests and bench.py in place of the reference's
    VQLPIPSWithDiscriminator (whose LPIPS needs downloaded VGG16 weights):
    L1(x, xrec) + codebook_weight * qloss + budget(gate).  Same call contract as
    modules/losses/vqperceptual_multidisc.py:115-194 (returns (loss, log dict))."""

    def __init__(self, codebook_weight=1.0, budget_loss_config=None):
        super().__init__()
        self.codebook_weight = codebook_weight
        self.budget_loss = None if budget_loss_config is None else instantiate_from_config(budget_loss_config)
        self.discriminator = nn.Conv2d(3, 1, 1)      # placeholder so configure_optimizers has parameters

    def forward(self, codebook_loss, inputs, reconstructions, optimizer_idx, global_step, last_layer=None,
                split="train", gate=None):
        rec = (inputs.contiguous() - reconstructions.contiguous()).abs().mean()
        if optimizer_idx == 0:
            loss = rec + self.codebook_weight * codebook_loss.mean()
            if self.budget_loss is not None and gate is not None:
                loss = loss + self.budget_loss(gate=gate)
            return loss, {f"{split}_total_loss": loss.detach(), f"{split}_quant_loss": codebook_loss.detach().mean(),
                          f"{split}_rec_loss": rec.detach()}
        d = self.discriminator(reconstructions.detach()).mean() * 0.0
        return d, {f"{split}_disc_loss": d.detach()}
