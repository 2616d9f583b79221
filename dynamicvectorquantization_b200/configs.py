"""Python restatement of the model sections of the reference's ``configs/stage1/*.yml`` (used where
the YAML files are not available - GPU box, bench.py) with the loss swapped for the surrogate.
``tests/test_dropin_conformance.py`` checks these against the YAML files in the build container.
Targets are the REFERENCE's dotted paths: they resolve to this package through the overlay
(``dynamicvectorquantization_b200/overlay`` first on ``sys.path``)."""
import copy
import os
import sys

OVERLAY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "overlay")


def activate_overlay():
    """Put the overlay tree first on sys.path (idempotent)."""
    if OVERLAY in sys.path:
        sys.path.remove(OVERLAY)
    sys.path.insert(0, OVERLAY)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(1, root)


_VQ = dict(target="modules.vector_quantization.quantize2_mask.VectorQuantize2",
           params=dict(codebook_size=1024, codebook_dim=256, channel_last=False, accept_image_fmap=True,
                       commitment_beta=0.25, decay=0.99, restart_unused_codes=True))
_DEC = dict(target="modules.dynamic_modules.DecoderPositional.Decoder",
            params=dict(ch=128, in_ch=256, out_ch=3, ch_mult=[1, 1, 2, 2], num_res_blocks=2, resolution=256,
                        attn_resolutions=[32], latent_size=32, window_size=2, position_type="fourier+learned"))
_TAIL = dict(quant_before_dim=256, quant_after_dim=256, quant_sample_temperature=0.0, image_key="image",
             monitor="val_rec_loss", warmup_epochs=0.1, scheduler_type="linear-warmup_cosine-decay")
_BUDGET_DUAL = dict(target="modules.dynamic_modules.budget.BudgetConstraint_RatioMSE_DualGrain",
                    params=dict(target_ratio=0.5, gamma=10.0, min_grain_size=16, max_grain_size=32, calculate_all=True))
_BUDGET_TRIPLE = dict(target="modules.dynamic_modules.budget.BudgetConstraint_NormedSeperateRatioMSE_TripleGrain",
                      params=dict(target_fine_ratio=0.3, target_median_ratio=0.3, gamma=1.0, min_grain_size=8,
                                  median_grain_size=16, max_grain_size=32))


def _surrogate(budget):
    p = dict(codebook_weight=1.0)
    if budget is not None:
        p["budget_loss_config"] = budget
    return dict(target="dynamicvectorquantization_b200.nn.model.SurrogateAELoss", params=p)


def real_loss_config(budget):
    """lossconfig of configs/stage1/dqvae-dual-r-05_imagenet.yml:36-63 (LPIPS + PatchGAN + adaptive weight)."""
    p = dict(disc_start=0,
             disc_config=dict(target="modules.discriminator.model.NLayerDiscriminator",
                              params=dict(input_nc=3, ndf=64, n_layers=3, use_actnorm=False)),
             disc_init=True, codebook_weight=1.0, pixelloss_weight=1.0, disc_factor=1.0, disc_weight=1.0,
             perceptual_weight=1.0, disc_conditional=False, disc_loss="hinge", disc_weight_max=0.75)
    if budget is not None:
        p["budget_loss_config"] = budget
    return dict(target="modules.losses.vqperceptual_multidisc.VQLPIPSWithDiscriminator", params=p)


def _enc_dual(router):
    return dict(target="modules.dynamic_modules.EncoderDual.DualGrainEncoder",
                params=dict(ch=128, ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[16, 32],
                            dropout=0.0, resamp_with_conv=True, in_channels=3, resolution=256, z_channels=256,
                            router_config=router))


STAGE1 = {
    "dqvae-dual-r-05": dict(
        target="models.stage1_dynamic.dqvae_dual_feat.DualGrainVQModel",
        params=dict(encoderconfig=_enc_dual(dict(
            target="modules.dynamic_modules.RouterDual.DualGrainFeatureRouter",
            params=dict(num_channels=256, normalization_type="group-32", gate_type="2layer-fc-SiLu"))),
            decoderconfig=_DEC, lossconfig=_surrogate(_BUDGET_DUAL), vqconfig=_VQ, **_TAIL)),
    "dqvae-entropy-dual-r05": dict(
        target="models.stage1_dynamic.dqvae_dual_entropy.DualGrainVQModel",
        params=dict(encoderconfig=_enc_dual(dict(
            target="modules.dynamic_modules.RouterDual.DualGrainFixedEntropyRouter",
            params=dict(json_path="scripts/tools/thresholds/entropy_thresholds_imagenet_train_patch-16.json",
                        fine_grain_ratito=0.5))),
            decoderconfig=_DEC, lossconfig=_surrogate(None), vqconfig=_VQ, **_TAIL)),
    "dqvae-triple-r-03-03": dict(
        target="models.stage1_dynamic.dqvae_triple_feat.TripleGrainVQModel",
        params=dict(encoderconfig=dict(
            target="modules.dynamic_modules.EncoderTriple.TripleGrainEncoder",
            params=dict(ch=128, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2, attn_resolutions=[8, 16, 32],
                        dropout=0.0, resamp_with_conv=True, in_channels=3, resolution=256, z_channels=256,
                        router_config=dict(target="modules.dynamic_modules.RouterTriple.TripleGrainFeatureRouter",
                                           params=dict(num_channels=256, normalization_type="group-32",
                                                       gate_type="2layer-fc-SiLu")))),
            decoderconfig=_DEC, lossconfig=_surrogate(_BUDGET_TRIPLE), vqconfig=_VQ, **_TAIL)),
}
STAGE1["dqvae-entropy-dual-r05"]["params"]["encoderconfig"]["params"]["update_router"] = False


def stage1_config(name):
    return copy.deepcopy(STAGE1[name])


def scaled_dual_config(ch=64, resolution=64, z_channels=64, codebook_size=128, attn=(8, 16)):
    """Same topology as dqvae-dual-r-05 at reduced width / resolution (parity tests)."""
    cfg = stage1_config("dqvae-dual-r-05")
    p = cfg["params"]
    lat = resolution // 8
    p["encoderconfig"]["params"].update(ch=ch, resolution=resolution, z_channels=z_channels,
                                        attn_resolutions=[lat // 2, lat])
    p["encoderconfig"]["params"]["router_config"]["params"]["num_channels"] = z_channels
    p["decoderconfig"]["params"].update(ch=ch, in_ch=z_channels, resolution=resolution,
                                        attn_resolutions=[lat], latent_size=lat)
    p["vqconfig"]["params"].update(codebook_size=codebook_size, codebook_dim=z_channels)
    p["lossconfig"]["params"]["budget_loss_config"]["params"].update(min_grain_size=lat // 2, max_grain_size=lat)
    p.update(quant_before_dim=z_channels, quant_after_dim=z_channels)
    return cfg


def scaled_triple_config(ch=64, resolution=128, z_channels=64, codebook_size=128):
    """dqvae-triple-r-03-03 topology at reduced width / resolution (parity tests)."""
    cfg = stage1_config("dqvae-triple-r-03-03")
    p = cfg["params"]
    lat = resolution // 8
    p["encoderconfig"]["params"].update(ch=ch, resolution=resolution, z_channels=z_channels,
                                        attn_resolutions=[lat // 4, lat // 2, lat])
    p["encoderconfig"]["params"]["router_config"]["params"]["num_channels"] = z_channels
    p["decoderconfig"]["params"].update(ch=ch, in_ch=z_channels, resolution=resolution,
                                        attn_resolutions=[lat], latent_size=lat)
    p["vqconfig"]["params"].update(codebook_size=codebook_size, codebook_dim=z_channels)
    p["lossconfig"]["params"]["budget_loss_config"]["params"].update(
        min_grain_size=lat // 4, median_grain_size=lat // 2, max_grain_size=lat)
    p.update(quant_before_dim=z_channels, quant_after_dim=z_channels)
    return cfg


def scaled_entropy_config(json_path, ch=64, resolution=64, z_channels=64, codebook_size=128):
    """dqvae-entropy-dual-r05 topology at reduced width; `json_path` holds the entropy thresholds."""
    cfg = stage1_config("dqvae-entropy-dual-r05")
    small = scaled_dual_config(ch=ch, resolution=resolution, z_channels=z_channels, codebook_size=codebook_size)
    p, sp = cfg["params"], small["params"]
    router = p["encoderconfig"]["params"]["router_config"]
    router["params"]["json_path"] = json_path
    p["encoderconfig"]["params"] = dict(sp["encoderconfig"]["params"], router_config=router, update_router=False)
    p["decoderconfig"], p["vqconfig"] = sp["decoderconfig"], sp["vqconfig"]
    p.update(quant_before_dim=z_channels, quant_after_dim=z_channels,
             entropy_patch_size=resolution // (resolution // 8 // 2), image_size=resolution)
    return cfg


def build_model(cfg):
    activate_overlay()
    from dynamicvectorquantization_b200.config import instantiate_from_config
    return instantiate_from_config(cfg)
