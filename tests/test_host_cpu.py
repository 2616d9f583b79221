"""CPU-only tests: host-side logic, C-ABI surface, drop-in conformance with the reference tree,
and the data-parallel exchange of the VQ statistics (gloo, world_size 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


# ------------------------------------------------------------------------------- C ABI surface
def test_cabi_exports_every_declared_symbol():
    from dynamicvectorquantization_b200 import _cabi, build
    build.build()
    lib = _cabi.lib()
    header = open(os.path.join(ROOT, "include", "b200dq.h")).read()
    declared = set(re.findall(r"\bint\s+(b2dq_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b2dq_version() == 0


def test_struct_layouts_match_header_field_order():
    from dynamicvectorquantization_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "b200dq.h")).read()
    for cname, cls in (("b2dq_tapgemm_desc", _cabi.TapGemmDesc), ("b2dq_mm_desc", _cabi.MmDesc)):
        body = header[header.index("typedef struct " + cname):]
        body = body[body.index("{") + 1:body.index("} " + cname)]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            stmt = re.sub(r"^(const\s+)?(void\*|float\*|long long|int|float)\s*", "", stmt)
            for piece in stmt.split(","):
                names.append(re.sub(r"[\*\s]|\[.*?\]", "", piece))
        assert names == [f[0] for f in cls._fields_], (cname, names)


def test_extension_missing_is_loud(monkeypatch, tmp_path):
    from dynamicvectorquantization_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.ExtensionMissing):
        _cabi.lib()


def test_cpu_tensors_are_rejected_not_silently_computed():
    from dynamicvectorquantization_b200 import configs
    configs.activate_overlay()
    from modules.vector_quantization.quantize2_mask import VectorQuantize2
    vq = VectorQuantize2(codebook_size=64, codebook_dim=64)
    with pytest.raises(RuntimeError):
        vq(torch.zeros(1, 64, 4, 4))


# ------------------------------------------------------------------------------- host geometry
def test_tile_shapes_and_tap_tables():
    from dynamicvectorquantization_b200 import kernels as kn
    for w, h, nb in [(256, 256, 32), (64, 64, 2), (32, 32, 1), (16, 16, 32), (8, 8, 3), (4, 4, 5)]:
        tw, th, tn = kn.tile_shape(w, h, nb)
        assert tw * th * tn == 128 and tw <= max(w, 1) * 2
        kw, kh, kq = kn.tile_shape(w, h, nb, pixels=64)
        assert kw * kh * kq == 64
    # stride-2 taps address x[2*oh + r, 2*ow + s] through the [N, H/2, 2, W/2, 2C] view
    cin = 64
    for r, s in kn.TAPS_3x3:
        dc, dw, dp, dh = (s % 2) * cin, s // 2, r % 2, r // 2
        for oh, ow in [(0, 0), (3, 5)]:
            assert 2 * (oh + dh) + dp == 2 * oh + r and 2 * (ow + dw) + dc // cin == 2 * ow + s


def test_weight_packings_are_consistent():
    from dynamicvectorquantization_b200 import kernels as kn
    w = torch.randn(8, 4, 3, 3)
    f, d = kn.pack_weight_fwd(w).float(), kn.pack_weight_dgrad(w).float()
    wb = w.bfloat16().float()
    for r, s in kn.TAPS_3x3:
        t = r * 3 + s
        assert torch.equal(f[:, t * 4:(t + 1) * 4], wb[:, :, r, s])
        assert torch.equal(d[:, t * 8:(t + 1) * 8], wb[:, :, r, s].t())


def test_weight_pack_cache_follows_the_parameter():
    """The bf16 packings are cached on the parameter and refreshed on in-place updates; a new
    parameter that happens to reuse a freed storage address must not see the old packing."""
    from dynamicvectorquantization_b200 import ops
    w = torch.nn.Parameter(torch.randn(8, 4, 3, 3))
    p1 = ops._packed(w, "fwd")
    assert ops._packed(w, "fwd") is p1
    with torch.no_grad():
        w.mul_(2.0)
    p2 = ops._packed(w, "fwd")
    assert p2 is not p1 and torch.equal(p2.float(), (p1.float() * 2).bfloat16().float())
    ptr = w.data_ptr()
    del w, p1, p2
    seen = False
    for _ in range(8):
        w2 = torch.nn.Parameter(torch.randn(8, 4, 3, 3))
        seen |= w2.data_ptr() == ptr
        assert torch.equal(ops._packed(w2, "fwd").float(),
                           w2.detach().permute(0, 2, 3, 1).reshape(8, -1).bfloat16().float())
        del w2


# ------------------------------------------------------------------------------- drop-in conformance
@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("name,yml", [("dqvae-dual-r-05", "dqvae-dual-r-05_imagenet.yml"),
                                      ("dqvae-entropy-dual-r05", "dqvae-entropy-dual-r05_imagenet.yml"),
                                      ("dqvae-triple-r-03-03", "dqvae-triple-r-03-03_imagenet.yml")])
def test_configs_match_reference_yaml(name, yml):
    import yaml
    from dynamicvectorquantization_b200 import configs
    ref = yaml.safe_load(open(os.path.join(REF, "configs/stage1", yml)))["model"]
    mine = configs.stage1_config(name)
    assert mine["target"] == ref["target"]
    for key in ("encoderconfig", "decoderconfig", "vqconfig"):
        assert mine["params"][key] == ref["params"][key], key
    for key in ("quant_before_dim", "quant_after_dim", "quant_sample_temperature", "image_key", "monitor",
                "warmup_epochs", "scheduler_type"):
        assert mine["params"][key] == ref["params"][key], key
    # the real loss (bench.py --loss real, tests/test_gpu_loss.py) is the YAML's lossconfig, budget term included
    budget = ref["params"]["lossconfig"]["params"].get("budget_loss_config")
    assert configs.real_loss_config(budget) == ref["params"]["lossconfig"], name
    if name == "dqvae-dual-r-05":
        assert configs._BUDGET_DUAL == budget
    if name == "dqvae-triple-r-03-03":
        assert configs._BUDGET_TRIPLE == budget


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_overlay_loads_reference_yaml_and_state_dict_keys_match():
    """The reference's own YAML (unchanged apart from the loss, which needs downloaded VGG weights)
    instantiates through the overlay, and the resulting state_dict has the reference's keys/shapes."""
    import subprocess
    code = r'''
import sys, types, yaml, torch, torch.nn as nn
sys.path.insert(0, %r); sys.path.insert(0, %r + "/dynamicvectorquantization_b200/overlay"); sys.path.append(%r)
import os; os.chdir(%r)
from utils.utils import instantiate_from_config          # the REFERENCE's plugin loader
conf = yaml.safe_load(open("configs/stage1/dqvae-dual-r-05_imagenet.yml"))["model"]
conf["params"]["lossconfig"] = {"target": "dynamicvectorquantization_b200.nn.model.SurrogateAELoss"}
m = instantiate_from_config(conf)
assert type(m).__module__.startswith("dynamicvectorquantization_b200"), type(m)
mine = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("loss.")}
# now the reference classes themselves (overlay removed from the path)
sys.path = [p for p in sys.path if "overlay" not in p]
for k in [k for k in sys.modules if k.split(".")[0] in ("modules", "models")]:
    del sys.modules[k]
pl = types.ModuleType("pytorch_lightning"); pl.LightningModule = nn.Module; sys.modules["pytorch_lightning"] = pl
from modules.dynamic_modules.EncoderDual import DualGrainEncoder
from modules.dynamic_modules.DecoderPositional import Decoder
from modules.vector_quantization.quantize2_mask import VectorQuantize2
assert DualGrainEncoder.__module__ == "modules.dynamic_modules.EncoderDual"
p = conf["params"]
ref = {}
for pref, mod in (("encoder.", DualGrainEncoder(**p["encoderconfig"]["params"])), ("decoder.", Decoder(**p["decoderconfig"]["params"])),
                  ("quantize.", VectorQuantize2(**p["vqconfig"]["params"]))):
    ref.update({pref + k: tuple(v.shape) for k, v in mod.state_dict().items()})
for k in ("quant_conv", "post_quant_conv"):
    ref[k + ".weight"] = (256, 256, 1, 1); ref[k + ".bias"] = (256,)
assert mine == ref, sorted(set(mine.items()) ^ set(ref.items()))[:10]
# seeded default init is identical (same construction order and init calls)
torch.manual_seed(5); a = DualGrainEncoder(**p["encoderconfig"]["params"]).state_dict()
sys.path.insert(0, %r + "/dynamicvectorquantization_b200/overlay")
for k in [k for k in sys.modules if k.split(".")[0] in ("modules", "models")]:
    del sys.modules[k]
from modules.dynamic_modules.EncoderDual import DualGrainEncoder as Mine
torch.manual_seed(5); b = Mine(**p["encoderconfig"]["params"]).state_dict()
assert all(torch.equal(a[k], b[k]) for k in a), "seeded init differs"
print("CONFORMANCE_OK")
''' % (ROOT, ROOT, REF, REF, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "CONFORMANCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_loss_overlay_conforms_to_the_reference_loss_module():
    """modules.losses.vqperceptual_multidisc.VQLPIPSWithDiscriminator through the overlay, built from the
    reference's own YAML lossconfig: same state_dict keys / shapes as the reference class (torchvision VGG16
    without the pretrained file), same seeded discriminator initialisation, same loss helpers on CPU tensors."""
    import subprocess
    code = r'''
import os, sys, yaml, torch
os.environ["B200DQ_ALLOW_RANDOM_VGG"] = "1"
sys.path.insert(0, %r); sys.path.insert(0, %r + "/dynamicvectorquantization_b200/overlay"); sys.path.append(%r)
os.chdir(%r)
from utils.utils import instantiate_from_config
conf = yaml.safe_load(open("configs/stage1/dqvae-dual-r-05_imagenet.yml"))["model"]["params"]["lossconfig"]
torch.manual_seed(3)
mine = instantiate_from_config(conf)
assert type(mine).__module__.startswith("dynamicvectorquantization_b200"), type(mine)
assert type(mine.perceptual_loss).__module__.startswith("dynamicvectorquantization_b200")
assert type(mine.discriminator).__module__.startswith("dynamicvectorquantization_b200")
assert not mine.perceptual_loss.training and all(not p.requires_grad for p in mine.perceptual_loss.parameters())
mk = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
md = {k: v.clone() for k, v in mine.discriminator.state_dict().items()}
import dynamicvectorquantization_b200.nn.losses as L
x, y = torch.randn(4, 1, 6, 6), torch.randn(4, 1, 6, 6)
mine_vals = [float(L.hinge_d_loss(x, y)), float(L.vanilla_d_loss(x, y)), float(L.bce_discr_loss(x, y)),
             float(L.hinge_g_loss(y)), float(L.bce_gen_loss(y)), L.adopt_weight(1.0, 3, threshold=5), L.adopt_weight(1.0, 7, threshold=5)]
# the reference classes themselves
sys.path = [p for p in sys.path if "overlay" not in p]
for k in [k for k in sys.modules if k.split(".")[0] in ("modules", "models")]:
    del sys.modules[k]
import torchvision
import modules.losses.lpips as ref_lpips
tv = torchvision.models.vgg16
class _M:
    @staticmethod
    def vgg16(pretrained=True):
        return tv(weights=None)
ref_lpips.models = _M
torch.manual_seed(3)
ref = instantiate_from_config(conf)
assert type(ref).__module__ == "modules.losses.vqperceptual_multidisc"
rk = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
assert mk == rk, sorted(set(mk.items()) ^ set(rk.items()))[:10]
# lin heads come from the reference tree's modules/lpips/vgg.pth in both
for k in range(5):
    assert torch.equal(getattr(mine.perceptual_loss, "lin%%d" %% k).model[1].weight, getattr(ref.perceptual_loss, "lin%%d" %% k).model[1].weight)
import modules.losses.vqperceptual_multidisc as R
ref_vals = [float(R.hinge_d_loss(x, y)), float(R.vanilla_d_loss(x, y)), float(R.bce_discr_loss(x, y)),
            float(R.hinge_g_loss(y)), float(R.bce_gen_loss(y)), R.adopt_weight(1.0, 3, threshold=5), R.adopt_weight(1.0, 7, threshold=5)]
assert mine_vals == ref_vals, (mine_vals, ref_vals)
# discriminator: identical module structure and init distribution (weights_init); same forward on CPU given same weights
ref.discriminator.load_state_dict(md)
from dynamicvectorquantization_b200.nn.discriminator import NLayerDiscriminator
d2 = NLayerDiscriminator(input_nc=3, ndf=64, n_layers=3); d2.load_state_dict(md)
img = torch.randn(2, 3, 64, 64)
# the overlay's forward runs on the CUDA kernels only: on CPU tensors it must fail loudly, never fall back; its
# parameter containers are the reference's own nn.Sequential (same modules at the same indices, same state)
assert [type(a).__name__ for a in d2.main] == [type(a).__name__ for a in ref.discriminator.main]
assert [tuple(v.shape) for v in d2.state_dict().values()] == [tuple(v.shape) for v in ref.discriminator.state_dict().values()]
assert torch.equal(ref.discriminator.eval()(img), d2.main.eval()(img))
try:
    d2.eval()(img)
    raise AssertionError("NLayerDiscriminator ran on CPU tensors")
except RuntimeError as e:
    assert "no CPU fallback" in str(e)
print("LOSS_CONFORMANCE_OK")
''' % (ROOT, ROOT, REF, REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert "LOSS_CONFORMANCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_lpips_refuses_to_run_without_weights_unless_told(monkeypatch, tmp_path):
    """No silent random perceptual metric: without the pretrained files the constructor raises."""
    import subprocess
    code = "import os, sys; sys.path.insert(0, %r); os.chdir(%r); os.environ.pop('B200DQ_ALLOW_RANDOM_VGG', None)\n" \
           "os.environ['TORCH_HOME'] = %r\n" \
           "from dynamicvectorquantization_b200.nn.lpips import LPIPS\n" \
           "try:\n    LPIPS()\n    print('BUILT')\nexcept RuntimeError as e:\n    print('RAISED', str(e)[:60])\n" % (ROOT, str(tmp_path), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "RAISED" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_checkpoint_loads_through_init_from_ckpt(tmp_path):
    """Checkpoint I/O conformance: a Lightning-style checkpoint {"state_dict": ...} written from
    the REFERENCE's modules (encoder, decoder, quantizer, 1x1 convs, loss incl. LPIPS + discriminator) loads into
    the overlay model through the reference's own ckpt_path / ignore_keys arguments (dqvae_dual_feat.py:36-57),
    tensor for tensor, and round-trips back into the reference modules."""
    import subprocess
    code = r'''
import os, sys, types, yaml, torch, torch.nn as nn
os.environ["B200DQ_ALLOW_RANDOM_VGG"] = "1"
ROOT, REF, TMP = %r, %r, %r
sys.path.insert(0, ROOT); sys.path.append(REF); os.chdir(REF)
pl = types.ModuleType("pytorch_lightning"); pl.LightningModule = nn.Module; sys.modules["pytorch_lightning"] = pl
import torchvision
import modules.losses.lpips as ref_lpips
tv = torchvision.models.vgg16
class _M:
    @staticmethod
    def vgg16(pretrained=True):
        return tv(weights=None)
ref_lpips.models = _M
from utils.utils import instantiate_from_config
conf = yaml.safe_load(open("configs/stage1/dqvae-dual-r-05_imagenet.yml"))["model"]
p = conf["params"]
torch.manual_seed(11)
parts = {"encoder": instantiate_from_config(p["encoderconfig"]), "decoder": instantiate_from_config(p["decoderconfig"]),
         "quantize": instantiate_from_config(p["vqconfig"]), "loss": instantiate_from_config(p["lossconfig"]),
         "quant_conv": nn.Conv2d(256, 256, 1), "post_quant_conv": nn.Conv2d(256, 256, 1)}
assert type(parts["encoder"]).__module__ == "modules.dynamic_modules.EncoderDual"          # the reference's classes
sd = {}
for name, m in parts.items():
    with torch.no_grad():
        for t in m.state_dict().values():
            if t.dtype.is_floating_point:
                t.add_(torch.randn_like(t) * 0.01)                                           # a "trained" state
    sd.update({name + "." + k: v.clone() for k, v in m.state_dict().items()})
ckpt = os.path.join(TMP, "last.ckpt")
torch.save({"state_dict": sd, "epoch": 3, "global_step": 1234}, ckpt)
# ---- now the overlay
for k in [k for k in sys.modules if k.split(".")[0] in ("modules", "models")]:
    del sys.modules[k]
sys.path.insert(0, ROOT + "/dynamicvectorquantization_b200/overlay")
conf2 = yaml.safe_load(open("configs/stage1/dqvae-dual-r-05_imagenet.yml"))["model"]
conf2["params"]["ckpt_path"] = ckpt
model = instantiate_from_config(conf2)
assert type(model).__module__.startswith("dynamicvectorquantization_b200")
mine = model.state_dict()
assert set(mine) == set(sd), sorted(set(mine) ^ set(sd))[:10]
bad = [k for k in sd if not torch.equal(mine[k], sd[k])]
assert not bad, bad[:10]
# ignore_keys drops whole sub-trees, as in the reference
conf3 = yaml.safe_load(open("configs/stage1/dqvae-dual-r-05_imagenet.yml"))["model"]
conf3["params"].update(ckpt_path=ckpt, ignore_keys=["loss.discriminator", "quantize"])
torch.manual_seed(5)
m3 = instantiate_from_config(conf3).state_dict()
assert torch.equal(m3["encoder.conv_in.weight"], sd["encoder.conv_in.weight"])
assert not torch.equal(m3["loss.discriminator.main.0.weight"], sd["loss.discriminator.main.0.weight"])
assert not torch.equal(m3["quantize.codebook.weight"], sd["quantize.codebook.weight"])
# and back: the overlay's state_dict loads into the reference modules strictly
for name, m in parts.items():
    m.load_state_dict({k[len(name) + 1:]: v for k, v in mine.items() if k.startswith(name + ".")}, strict=True)
print("CKPT_CONFORMANCE_OK")
''' % (ROOT, REF, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert "CKPT_CONFORMANCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_loss_module_host_logic_matches_reference_golden(monkeypatch, mode):
    """The loss module's own logic (adaptive weight, clamps, hinge terms, budget, BatchNorm bookkeeping, log keys)
    on CPU: the two parts that need the CUDA kernels are substituted - the perceptual term by the pinned fp32 oracle,
    the discriminator's forward by its own nn.Sequential of parameter containers run through torch ops - everything
    else runs as shipped and must reproduce the fixture minted from the reference's class."""
    monkeypatch.setenv("B200DQ_ALLOW_RANDOM_VGG", "1")
    import torch.nn.functional as F
    from dynamicvectorquantization_b200 import configs
    from dynamicvectorquantization_b200.nn.losses import VQLPIPSWithDiscriminator
    from oracle import loss_oracle as lo
    configs.activate_overlay()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "loss_small.npz"))
    disc_cfg = {"target": "modules.discriminator.model.NLayerDiscriminator",
                "params": {"input_nc": 3, "ndf": 64, "n_layers": 3, "use_actnorm": False}}
    loss = VQLPIPSWithDiscriminator(disc_start=0, disc_config=disc_cfg, disc_init=True, disc_weight_max=0.75,
                                    budget_loss_config=configs._BUDGET_DUAL)
    sd = lo.make_loss_weights(seed=21)
    loss.load_state_dict({k[len("loss."):]: v for k, v in sd.items()}, strict=False)
    loss.train(mode == "train")
    monkeypatch.setattr(loss.perceptual_loss, "forward", lambda a, b: lo.lpips(sd, a, b))
    monkeypatch.setattr(loss.discriminator, "forward", lambda inp: loss.discriminator.main(inp))
    x, feat, w_last, qloss, gate = lo.toy_inputs()
    w_last.requires_grad_(True)
    feat.requires_grad_(True)
    xrec = F.conv2d(feat, w_last, padding=1)
    l0, log0 = loss(qloss, x, xrec, 0, 0, last_layer=w_last, split="train", gate=gate)
    gw, gf = torch.autograd.grad(l0, [w_last, feat])
    p = mode + "_"

    def close(a, key, rtol=2e-4, atol=1e-7):
        b = gold[p + key]
        a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
        assert np.allclose(a, b, rtol=rtol, atol=atol), (key, float(np.abs(a - b).max()))
    close(l0, "loss0")
    assert sorted(log0) == sorted("train_" + k[len(p + "log0_"):] for k in gold.files if k.startswith(p + "log0_"))
    for k in gold.files:
        if k.startswith(p + "log0_"):
            close(log0["train_" + k[len(p + "log0_"):]], k[len(p):], rtol=1e-3)
    close(gw, "g_w_last", rtol=2e-3, atol=1e-6 * float(np.abs(gold[p + "g_w_last"]).max()) + 1e-9)
    close(gf, "g_feat", rtol=2e-3, atol=2e-3 * float(np.abs(gold[p + "g_feat"]).max()))
    l1, log1 = loss(qloss, x, xrec.detach(), 1, 0, last_layer=w_last, split="train")
    close(l1, "loss1")
    assert sorted(log1) == ["train_disc_loss", "train_logits_fake", "train_logits_real"]
    close(log1["train_logits_real"], "log1_logits_real", atol=1e-6)
    dparams = dict(loss.discriminator.named_parameters())
    for k, gr in zip(dparams, torch.autograd.grad(l1, list(dparams.values()))):
        close(gr.norm(), "gd_norm_" + k, rtol=2e-3, atol=1e-7)
    if mode == "train":
        for k, v in loss.state_dict().items():
            if "running" in k:
                close(v, "bn1_" + k, rtol=1e-4, atol=1e-6)
    # evaluation without a graph: the adaptive weight falls back to 0 like the reference (:139-142)
    loss.eval()
    with torch.no_grad():
        lv, logv = loss(qloss, x, xrec.detach(), 0, 0, last_layer=w_last, split="val", gate=gate)
    assert float(logv["val_d_weight"]) == 0.0 and torch.isfinite(lv)
    # discriminator warm-up: before disc_start the adversarial factor is 0
    loss.discriminator_iter_start = 10
    _, logw = loss(qloss, x, F.conv2d(feat, w_last, padding=1), 0, 3, last_layer=w_last, split="train", gate=gate)
    assert float(logw["train_disc_factor"]) == 0.0


# ------------------------------------------------------------------------------- folded upsample + conv
def test_upsample_conv_fold_is_exact_in_fp32():
    """The algebra behind kernels.upconv_*: conv3x3(nearest_upsample_x2(x)) equals, per output parity class, a 2x2
    convolution of the low-resolution input with the folded weights and offsets _UP_OFF (fp32 identity), the
    un-fold maps gradients back onto the 3x3 taps, and the bf16 packings address the blocks the tap tables name."""
    import torch.nn.functional as F
    from dynamicvectorquantization_b200 import kernels as kn
    g = torch.Generator().manual_seed(0)
    n, ci, co, h, w = 2, 5, 4, 6, 7
    x = torch.randn(n, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g, requires_grad=True)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), wt, padding=1)
    wf = kn.upconv_fold(wt)
    wf.retain_grad()
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for ph in (0, 1):
        for pw in (0, 1):
            acc = 0
            for a in (0, 1):
                for b in (0, 1):
                    dh, dw = kn._UP_OFF[ph][a], kn._UP_OFF[pw][b]
                    acc = acc + torch.einsum("nihw,oi->nohw", xp[:, :, 1 + dh:1 + dh + h, 1 + dw:1 + dw + w], wf[ph, pw, a, b])
            out[:, :, ph::2, pw::2] = acc
    assert torch.allclose(out, ref, atol=1e-5)
    up = torch.randn(ref.shape, generator=g)
    (gref,) = torch.autograd.grad((ref * up).sum(), wt, retain_graph=True)
    (out * up).sum().backward()
    assert torch.allclose(kn.upconv_unfold_grad(wf.grad), gref, atol=1e-4)
    assert torch.allclose(wt.grad, gref, atol=1e-4)
    # packings: column block ((ph*2+pw)*2+a)*2+b of the forward pack is wf[ph,pw,a,b] (Cout x Cin), of the data-gradient
    # pack its transpose
    fwd, dgr = kn.upconv_pack(wt)
    assert fwd.shape == (co, 16 * ci) and dgr.shape == (ci, 16 * co) and fwd.dtype == torch.bfloat16
    for ph in (0, 1):
        for pw in (0, 1):
            for a in (0, 1):
                for b in (0, 1):
                    blk = ((ph * 2 + pw) * 2 + a) * 2 + b
                    assert torch.equal(fwd[:, blk * ci:(blk + 1) * ci], wf[ph, pw, a, b].detach().to(torch.bfloat16))
                    assert torch.equal(dgr[:, blk * co:(blk + 1) * co], wf[ph, pw, a, b].detach().t().to(torch.bfloat16))


# ------------------------------------------------------------------------------- VQ work schedule
def _vq_items(n_tiles_codebook, plan, cta):
    """Python mirror of VqSched::next (csrc/vq.cu): the (row tile, j0, j1, nsplit, tail slot) items of one CTA."""
    grid, rounds, tail_tiles, q, _ = plan
    nn = n_tiles_codebook
    items = [(r * grid + cta, 0, nn, 1, 0) for r in range(rounds)]
    U = tail_tiles * nn
    u = min(cta * q, U)
    uend = min(U, u + q)
    while u < uend:
        tl, j0 = divmod(u, nn)
        j1 = min(nn, j0 + (uend - u))
        nsplit = ((tl + 1) * nn - 1) // q - (tl * nn) // q + 1
        items.append((rounds * grid + tl, j0, j1, nsplit, tl))
        u += j1 - j0
    return items


@pytest.mark.parametrize("G", [148, 132, 37, 7, 1])
def test_vq_stream_k_schedule_covers_every_tile_exactly_once(G):
    """The schedule b2dq_vq_search_plan hands the kernel: every (row tile, codebook tile) unit is searched exactly
    once, a CTA never sees a row tile twice, nsplit equals the number of CTAs that really share the tile (the
    arrival counter's target), tail slots stay inside the workspace, and a split is only planned where it pays."""
    import ctypes
    from dynamicvectorquantization_b200 import _cabi, build
    build.build()
    lib = _cabi.lib()
    out = (ctypes.c_int * 5)()
    cases = [(65536, 1024), (65536, 8192), (65536, 16384), (32768, 2048), (32768, 2304), (40000, 8192), (2048, 16384),
             (256, 16384), (1, 1800), (300, 4096), (1100, 5000), (19000, 4096), (128 * G, 4096), (128 * G + 1, 4096),
             (128 * (2 * G - 1), 2048), (5000, 256), (77, 300)]
    for N, K in cases:
        for allow in (0, 1):
            assert lib.b2dq_vq_search_plan(N, K, G, allow, out) == 0
            plan = tuple(out)
            grid, rounds, tail_tiles, q, split = plan
            tiles, nn = -(-N // 128), -(-K // 256)
            assert 1 <= grid <= G and rounds * grid + tail_tiles == tiles, (N, K, G, plan)
            if not allow or nn < 8:
                assert not split and q == nn
            seen = {}
            sharers = {}
            for cta in range(grid):
                mine = _vq_items(nn, plan, cta)
                assert len({it[0] for it in mine}) == len(mine), "a CTA got the same row tile twice"
                for tile, j0, j1, nsplit, tl in mine:
                    assert 0 <= j0 < j1 <= nn and tile < tiles
                    for j in range(j0, j1):
                        assert (tile, j) not in seen, (N, K, G, tile, j)
                        seen[(tile, j)] = cta
                    sharers.setdefault(tile, []).append((nsplit, tl))
                    if nsplit > 1:
                        assert split and 0 <= tl < tail_tiles
                        assert lib.b2dq_vq_search_workspace_bytes(N, K) >= tail_tiles * 128 * 8 + tail_tiles * 4
                    else:
                        assert (j0, j1) == (0, nn), "an unshared tile must be searched whole by its CTA"
            assert len(seen) == tiles * nn, (N, K, G, plan, len(seen))
            for tile, lst in sharers.items():
                assert all(ns == len(lst) for ns, _ in lst), (N, K, G, tile, lst)
            if split:                                  # the cut must shorten the last wave
                longest = max(sum(j1 - j0 for _, j0, j1, _, _ in _vq_items(nn, plan, c)) for c in range(grid))
                assert longest < (rounds + 1) * nn


# ------------------------------------------------------------------------------- data parallel (gloo)
def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from dynamicvectorquantization_b200.nn import quantize as Q
    from oracle import vq_oracle as vo
    K, C, N = 16, 8, 40
    g = np.random.RandomState(0)
    w = np.concatenate([g.randn(K, C).astype(np.float32), np.zeros((1, C), np.float32)])
    x_all = g.randn(world * N, C).astype(np.float32)
    x = x_all[rank * N:(rank + 1) * N]
    idx = vo.find_nearest_embedding(x, w)
    acc = torch.zeros(K * C + K)
    acc[:K * C].view(K, C).index_add_(0, torch.from_numpy(idx), torch.from_numpy(x))
    acc[K * C:] += torch.bincount(torch.from_numpy(idx), minlength=K).float()
    Q.reduce_ema_stats(acc)
    rows = torch.from_numpy(x[:K].copy())
    Q.share_restart_rows(rows)
    # single-process statistics over the concatenated batch must equal the reduced ones
    idx_all = vo.find_nearest_embedding(x_all, w)
    sums = np.zeros((K, C), np.float32); np.add.at(sums, idx_all, x_all)
    ok = np.allclose(acc[:K * C].view(K, C).numpy(), sums, atol=1e-5) and \
        np.array_equal(acc[K * C:].numpy(), np.bincount(idx_all, minlength=K).astype(np.float32)) and \
        np.array_equal(rows.numpy(), x_all[:K])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_vq_statistics_exchange_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in procs]
    assert res == [(0, True), (1, True)]


def test_bench_reference_arm_json_contract():
    """--impl reference prints one JSON line with the contract keys (rank != 0 prints nothing)."""
    import json
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_launcher_puts_the_overlay_ahead_of_the_script_directory(tmp_path):
    """`PYTHONPATH=<overlay>:... python train.py` from the reference root does NOT activate the
    overlay (sys.path[0] = the script's directory wins for namespace packages).  A script placed in a stand-in
    reference tree (its own modules/dynamic_modules/EncoderDual.py) is run (a) the documented-in-round-1 way and (b)
    through `python -m dynamicvectorquantization_b200.launch`: only (b) resolves the overlay class."""
    import subprocess
    fake = tmp_path / "ref"
    (fake / "modules" / "dynamic_modules").mkdir(parents=True)
    (fake / "modules" / "dynamic_modules" / "EncoderDual.py").write_text("class DualGrainEncoder:\n    pass\n")
    (fake / "modules" / "dynamic_modules" / "only_in_reference.py").write_text("MARK = 'reference'\n")
    (fake / "probe.py").write_text(
        "import sys, os\nsys.path.append(os.getcwd())\n"                       # what train.py:5 does
        "import modules.dynamic_modules.EncoderDual as m\n"
        "import modules.dynamic_modules.only_in_reference as o\n"
        "print('RESOLVED', m.__file__, o.MARK, sys.argv[1:])\n")
    from dynamicvectorquantization_b200 import configs
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([configs.OVERLAY, ROOT, str(fake)]))
    plain = subprocess.run([sys.executable, "probe.py", "--x"], cwd=fake, env=env, capture_output=True, text=True)
    assert plain.returncode == 0, plain.stderr
    assert str(fake) in plain.stdout.split("RESOLVED")[1], "expected the PYTHONPATH-only launch to pick the reference file"
    launched = subprocess.run([sys.executable, "-m", "dynamicvectorquantization_b200.launch", "probe.py", "--x"], cwd=fake,
                              env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True)
    assert launched.returncode == 0, launched.stderr
    out = launched.stdout.split("RESOLVED")[1]
    assert os.path.join("dynamicvectorquantization_b200", "overlay") in out, out
    assert "reference" in out and "['--x']" in out            # fall-through to the reference tree and argv still work


def test_conv4x4_tap_tables_match_torch_conv():
    """PatchGAN 4x4 convolutions (modules/discriminator/model.py:37-66) on the tap GEMM: the tap tables of
    kernels.conv4x4_fwd / conv4x4_dgrad / the stem's image gradient are checked on the CPU by emulating the GEMM's
    addressing (shifted boxes of the NHWC tensor or of its stride-2 parity view, zero fill outside, strided parity
    classes of the output) in PyTorch against F.conv2d and its autograd."""
    import torch.nn.functional as F
    from dynamicvectorquantization_b200 import kernels as kn

    def box(view, n0, h0, p, w0, c0, cin, hh, ww):
        """view [N, H2, P, W2, C2] -> [N, hh, ww, cin] box starting at (h0, w0) with zero fill outside."""
        N, H2, P, W2, C2 = view.shape
        out = torch.zeros(N, hh, ww, cin, dtype=view.dtype)
        for i in range(hh):
            for j in range(ww):
                y, x = h0 + i, w0 + j
                if 0 <= y < H2 and 0 <= x < W2:
                    out[:, i, j] = view[:, y, p, x, c0:c0 + cin]
        return out

    g = torch.Generator().manual_seed(0)
    for stride, h, w, cin, cout in ((2, 8, 12, 4, 5), (1, 6, 7, 3, 4)):
        x = torch.randn(2, h, w, cin, generator=g, dtype=torch.float64)
        wt = torch.randn(cout, cin, 4, 4, generator=g, dtype=torch.float64)
        ref = F.conv2d(x.permute(0, 3, 1, 2), wt, stride=stride, padding=1).permute(0, 2, 3, 1)
        ho, wo = kn.conv4x4_out_hw(h, w, stride)
        assert ref.shape[1:3] == (ho, wo)
        wf = wt.permute(0, 2, 3, 1).reshape(cout, 16 * cin)                       # pack_weight_fwd layout
        view = x.view(2, h // 2, 2, w // 2, 2 * cin) if stride == 2 else x.view(2, h, 1, w, cin)
        y = torch.zeros(2, ho, wo, cout, dtype=torch.float64)
        for (dc, dw, dp, dh, bk) in kn._taps4(stride, cin):
            y += box(view, 0, dh, dp, dw, dc, cin, ho, wo) @ wf[:, bk:bk + cin].t()
        assert torch.allclose(y, ref, atol=1e-10), f"forward taps, stride {stride}"
        # data gradient: taps of kernels.conv4x4_dgrad over dy, weights in pack_weight_dgrad layout
        dy = torch.randn(2, ho, wo, cout, generator=g, dtype=torch.float64)
        xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
        F.conv2d(xr, wt, stride=stride, padding=1).backward(dy.permute(0, 3, 1, 2))
        ref_dx = xr.grad.permute(0, 2, 3, 1)
        wd = wt.permute(1, 2, 3, 0).reshape(cin, 16 * cout)
        dyv = dy.view(2, ho, 1, wo, cout)
        dx = torch.zeros(2, h, w, cin, dtype=torch.float64)
        if stride == 1:
            for r in range(4):
                for s_ in range(4):
                    dx += box(dyv, 0, 1 - r, 0, 1 - s_, 0, cout, h, w) @ wd[:, (r * 4 + s_) * cout:(r * 4 + s_ + 1) * cout].t()
        else:
            rsel = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}                     # as in kernels.conv4x4_dgrad
            for ph in (0, 1):
                for pw in (0, 1):
                    acc = torch.zeros(2, ho, wo, cin, dtype=torch.float64)
                    for r, dh in rsel[ph]:
                        for s_, dw in rsel[pw]:
                            acc += box(dyv, 0, dh, 0, dw, 0, cout, ho, wo) @ wd[:, (r * 4 + s_) * cout:(r * 4 + s_ + 1) * cout].t()
                    dx[:, ph::2, pw::2] = acc
        assert torch.allclose(dx, ref_dx, atol=1e-10), f"data-gradient taps, stride {stride}"


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from dynamicvectorquantization_b200.ddp import BucketedGradExchange
    torch.manual_seed(0)                                    # same parameters on every rank
    net = torch.nn.Sequential(torch.nn.Linear(12, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                              torch.nn.Linear(40, 3))
    unused = torch.nn.Parameter(torch.zeros(5))             # a parameter that never receives a gradient
    params = list(net.parameters()) + [unused]
    ex = BucketedGradExchange(params, bucket_mb=0.0005)     # ~130 floats per bucket: several buckets
    ok = len(ex.buckets) >= 3 and sum(b[2] for b in ex.buckets) == len(params)
    ok = ok and ex.buckets[0][1] == ex.flat.numel() and ex.buckets[-1][0] == 0         # reverse order, whole buffer
    for step in range(2):
        x = torch.randn(6, 12, generator=torch.Generator().manual_seed(10 * step + rank))
        ex.begin_step()
        net(x).pow(2).mean().backward()
        ex.finish()
        got = [p.grad.clone() for p in params]
        # reference: local gradients of every rank, averaged by hand
        ref_net = torch.nn.Sequential(*[type(m)(m.in_features, m.out_features) if isinstance(m, torch.nn.Linear) else type(m)()
                                        for m in net])
        ref_net.load_state_dict(net.state_dict())
        ref_net(x).pow(2).mean().backward()
        for g_, p in zip(got, ref_net.parameters()):
            mine = [torch.zeros_like(p.grad) for _ in range(world)]
            dist.all_gather(mine, p.grad)
            ok = ok and torch.allclose(g_, sum(mine) / world, rtol=1e-6, atol=1e-8)
        ok = ok and float(got[-1].abs().max()) == 0.0 and all(p.grad.data_ptr() >= ex.flat.data_ptr() for p in params)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_bucketed_gradient_exchange_world_size_2():
    """dynamicvectorquantization_b200/ddp.py on gloo: gradients are views into one flat buffer, buckets are cut in
    reverse parameter order and cover the buffer exactly, every bucket is averaged over the ranks as soon as its
    last gradient is accumulated (or in finish() for parameters without a gradient) - equal to averaging the ranks'
    local gradients by hand, over two steps."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in procs]
    assert res == [(0, True), (1, True)]


def test_oplevel_geometry_and_workspace_sizes_match_the_python_heuristics():
    """The operator-level C entries (csrc/oplevel.cu) derive tiles and split-K factors themselves; their host-only
    helpers must agree with kernels.py, or the two routes would stop being bit-identical."""
    import ctypes as C
    from dynamicvectorquantization_b200 import _cabi, kernels as kn
    lib = _cabi.lib()
    for (n, h, w, cin, cout, k, stride) in [(32, 256, 256, 128, 128, 3, 1), (32, 32, 32, 256, 256, 3, 1),
                                            (2, 32, 32, 256, 768, 1, 1), (32, 128, 128, 128, 128, 3, 2),
                                            (1, 24, 40, 64, 192, 3, 1)]:
        g = _cabi.Conv2dGeom(n, h, w, cin, cout, k, stride)
        out = (C.c_int * 2)()
        assert lib.b2dq_conv2d_out_hw(C.byref(g), out) == 0
        ho, wo = h // stride, w // stride
        assert (out[0], out[1]) == (ho, wo)
        pconv = kn._pconv_ok(k, stride, w, cin, cout, n, h)
        assert (lib.b2dq_conv2d_fwd_workspace_bytes(C.byref(g)) > 0) == pconv
        kw, kh, kn_ = kn.tile_shape(wo, ho, n, pixels=64)
        kblocks = -(-wo // kw) * -(-ho // kh) * -(-n // kn_)
        ntaps = k * k
        splits = kn._wgrad_splits(kblocks, -(-cout // 128) * -(-cin // 128) * ((ntaps + 2) // 3))
        want = (splits * ntaps * cout * cin * 4 + 255) // 256 * 256
        assert lib.b2dq_conv2d_wgrad_workspace_bytes(C.byref(g), 0) == want
    assert lib.b2dq_groupnorm_workspace_bytes(32, 65536, 128, 32, 1) >= lib.b2dq_gn_bwd_fused_workspace_bytes(32, 65536, 128, 32)
    assert lib.b2dq_groupnorm_workspace_bytes(2, 100, 96, 33, 0) == -1
    assert lib.b2dq_attention_workspace_bytes(2, 1024, 256, 0) == 2 * 1024 * 1024 * 4
