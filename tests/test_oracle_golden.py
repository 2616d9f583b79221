"""CPU tests: the oracle restatements reproduce what the reference's own classes computed
(golden .npz minted by tests/golden/make_golden.py from /root/reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import dqvae_oracle as orc
from oracle import vq_oracle as vo

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return {k: v for k, v in np.load(os.path.join(G, name), allow_pickle=False).items()}


def _flat(x_nchw):
    b, c, h, w = x_nchw.shape
    return np.ascontiguousarray(x_nchw.transpose(0, 2, 3, 1).reshape(-1, c))


def test_vq_oracle_eval_matches_reference():
    g = _load("vq_small.npz")
    x, w, mask = _flat(g["x"]), g["weight"], _flat(g["mask"])[:, 0]
    xq, loss, idx = vo.vq_forward(x, w, mask=mask, beta=0.25)
    assert np.array_equal(idx, g["codes"].reshape(-1))
    assert np.array_equal(xq, _flat(g["xq"]))           # x + (e - x) evaluated in fp32, like :182
    assert np.allclose(xq, w[idx], rtol=0, atol=1e-6)   # ... which is the gathered row up to 1 ulp
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    # gradients: straight-through is identity; commitment term = 2*beta*m*(x-e)/numel
    assert np.array_equal(g["gx_ste"], g["gq"])
    ref = 2 * 0.25 * mask[:, None] * (x - w[idx]) / x.size
    assert np.allclose(_flat(g["gx_loss"]), ref, rtol=1e-5, atol=1e-9)


def test_vq_oracle_training_trajectory_matches_reference():
    g = _load("vq_small.npz")
    w = g["weight"].copy()
    cs = np.ones(w.shape[0] - 1, np.float32)
    em = w[:-1].copy()
    for t in range(3):
        x = _flat(g[f"t{t}_x"])
        idx = vo.find_nearest_embedding(x, w)
        assert np.array_equal(idx, g[f"t{t}_codes"].reshape(-1))
        assert np.allclose(w[idx], _flat(g[f"t{t}_xq"]), rtol=0, atol=1e-5)  # gather uses the pre-update codebook
        cs, em = vo.update_buffers(x, idx, cs, em, 0.99, restart_rows=g[f"t{t}_restart"])
        w[:-1] = vo.update_embedding(cs, em)
        assert np.allclose(cs, g[f"t{t}_cs"], rtol=1e-6, atol=1e-7)
        assert np.allclose(em, g[f"t{t}_em"], rtol=1e-5, atol=1e-6)
        assert np.allclose(w, g[f"t{t}_w"], rtol=1e-5, atol=1e-6)


def test_bf16_round_matches_torch():
    a = np.random.RandomState(0).randn(4096).astype(np.float32) * 3
    assert np.array_equal(vo.bf16_round(a), torch.from_numpy(a).bfloat16().float().numpy())


@pytest.mark.parametrize("tag,cfg,seed", [("tiny", orc.TINY_CFG, 3), ("dual", orc.DUAL_CFG, 7)])
def test_model_oracle_matches_reference(tag, cfg, seed):
    path = os.path.join(G, f"model_{tag}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not minted")
    g = _load(f"model_{tag}.npz")
    sd = orc.make_weights(orc.model_shapes(cfg), seed=seed)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "cluster_size" not in k and "embed_ema" not in k}
    full = dict(sd); full.update(params)
    x = torch.from_numpy(g["x"])
    out = orc.model_forward(full, cfg, x)
    assert np.array_equal(out["codes"].numpy(), g["codes"].astype(np.int64))
    assert np.array_equal(out["indices"].numpy(), g["indices"].astype(np.int64))
    assert torch.allclose(out["xrec"], torch.from_numpy(g["xrec"]), rtol=1e-4, atol=1e-5)
    assert abs(float(out["qloss"]) - float(g["qloss"])) < 1e-5 * abs(float(g["qloss"])) + 1e-8
    loss = (out["xrec"] - x).abs().mean() + out["qloss"]
    loss.backward()
    for key in g:
        if key.startswith("grad__"):
            name = key[len("grad__"):].replace("__", ".")
            assert torch.allclose(params[name].grad, torch.from_numpy(g[key]), rtol=2e-3, atol=1e-6), name
    names = list(g["grad_norm_names"])
    for n, ref in zip(names, g["grad_norms"]):
        if n in params and params[n].grad is not None:
            got = float(params[n].grad.double().pow(2).sum().sqrt())
            assert abs(got - ref) <= 2e-3 * ref + 1e-7, (n, got, ref)


def test_train_mode_routing_oracle_matches_reference():
    """EncoderDual.py:130-149 in training mode (hard gumbel-softmax sample with the reference's noise replayed,
    gate_grad multiply, budget loss): values and the gradients that train the router."""
    g = _load("model_tiny_train_routing.npz")
    cfg = orc.TINY_CFG
    sd = orc.make_weights(orc.model_shapes(cfg), seed=9)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("encoder.")}
    full = dict(sd); full.update(params)
    enc = orc.dual_encoder(full, cfg, torch.from_numpy(g["x"]), gumbel_noise=torch.from_numpy(g["noise"]))
    assert np.array_equal(enc["indices"].numpy(), g["indices"])
    assert torch.allclose(enc["gate"], torch.from_numpy(g["gate"]), atol=1e-6)
    assert torch.allclose(enc["h_dual"], torch.from_numpy(g["h_dual"]), rtol=1e-4, atol=1e-5)
    assert np.array_equal(enc["codebook_mask"].numpy(), g["mask"])
    lat = cfg["latent_size"]
    budget = orc.budget_loss_dual(enc["gate"], min_grain=lat // 2, max_grain=lat)
    assert abs(float(budget) - float(g["budget"])) < 1e-5 * abs(float(g["budget"])) + 1e-8
    ((enc["h_dual"] * torch.from_numpy(g["probe"])).sum() + budget).backward()
    assert 0 < int(enc["indices"].sum()) < enc["indices"].numel()              # both grains sampled
    for key in g:
        if key.startswith("grad__"):
            name = "encoder." + key[len("grad__"):].replace("__", ".")
            assert torch.allclose(params[name].grad, torch.from_numpy(g[key]), rtol=2e-3, atol=1e-6), name
    for n, ref in zip(g["grad_norm_names"], g["grad_norms"]):
        got = float(params["encoder." + str(n)].grad.double().pow(2).sum().sqrt())
        assert abs(got - ref) <= 2e-3 * ref + 1e-7, (n, got, ref)
    assert float(params["encoder.router.gate.0.weight"].grad.abs().max()) > 0  # the router does get a gradient


def test_triple_oracle_matches_reference():
    g = _load("model_tiny_triple.npz")
    cfg = orc.TINY_TRIPLE_CFG
    sd = orc.make_weights(orc.model_shapes(cfg), seed=5)
    out = orc.model_forward(sd, cfg, torch.from_numpy(g["x"]))
    assert np.array_equal(out["indices"].numpy(), g["indices"].astype(np.int64))
    assert np.array_equal(out["codes"].numpy(), g["codes"].astype(np.int64))
    assert torch.allclose(out["xrec"], torch.from_numpy(g["xrec"]), rtol=1e-4, atol=1e-5)
    assert abs(float(out["qloss"]) - float(g["qloss"])) < 1e-5 * abs(float(g["qloss"]))
    b = orc.budget_loss_triple(out["gate"], min_grain=2, median_grain=4, max_grain=8)
    assert abs(float(b) - float(g["budget"])) < 1e-5 * abs(float(g["budget"])) + 1e-8


def test_triple_oracle_backward_matches_reference():
    """Rows a3 + a18 for the triple-grain model: the oracle's autograd of |xrec - x|.mean() + qloss + triple budget loss
    against the gradients of the reference's own EncoderTriple / Decoder / VectorQuantize2 (eval-mode routing)."""
    g = _load("model_tiny_triple_grads.npz")
    cfg = orc.TINY_TRIPLE_CFG
    sd = orc.make_weights(orc.model_shapes(cfg), seed=5)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "cluster_size" not in k and "embed_ema" not in k}
    full = dict(sd); full.update(params)
    x = torch.from_numpy(g["x"])
    out = orc.model_forward(full, cfg, x)
    assert np.array_equal(out["indices"].numpy(), g["indices"].astype(np.int64))
    assert np.array_equal(out["codes"].numpy(), g["codes"].astype(np.int64))
    loss = (out["xrec"] - x).abs().mean() + out["qloss"] + orc.budget_loss_triple(out["gate"], min_grain=2, median_grain=4,
                                                                                  max_grain=8)
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    checked = 0
    for key in g:
        if key.startswith("grad__"):
            name = key[len("grad__"):].replace("__", ".")
            assert torch.allclose(params[name].grad, torch.from_numpy(g[key]), rtol=2e-3, atol=1e-6), name
            checked += 1
    assert checked >= 8
    seen = 0
    for n, ref in zip(g["grad_norm_names"], g["grad_norms"]):
        n = str(n)
        if n in params and params[n].grad is not None:
            got = float(params[n].grad.double().pow(2).sum().sqrt())
            assert abs(got - ref) <= 2e-3 * ref + 1e-7, (n, got, ref)
            seen += 1
    assert seen >= 400                                          # every trainable tensor of the model


def test_entropy_oracle_matches_reference():
    g = _load("entropy_small.npz")
    x = torch.from_numpy(g["x"])
    ent = orc.patch_entropy(x, patch=16)
    assert torch.allclose(ent, torch.from_numpy(g["entropy"]), rtol=1e-5, atol=1e-6)
    gate = orc.entropy_router(ent, float(g["threshold"]))
    assert np.array_equal(gate.numpy(), g["gate"])
    assert 0 < int(gate[..., 1].sum()) < gate[..., 1].numel()          # both grains present
    b = orc.budget_loss_dual(gate.permute(0, 3, 1, 2).float(), min_grain=4, max_grain=8)
    assert abs(float(b) - float(g["budget"])) < 1e-6 * abs(float(g["budget"])) + 1e-9


# --------------------------------------------------------------------------- sibling quantizers (8f row 2)
from oracle import vq_family_oracle as vf  # noqa: E402


def _unflat(rows, like_nchw):
    b, c, h, w = like_nchw.shape
    return rows.reshape(b, h, w, c).transpose(0, 3, 1, 2)


@pytest.mark.parametrize("legacy", [True, False])
def test_family_quantize2_oracle_matches_reference(legacy):
    g = _load("vq_family.npz")
    p = f"q2_legacy{int(legacy)}"
    x = _flat(g["q2_x"])
    xq, loss, idx, gx = vf.vq2_forward(x, g["q2_weight"], beta=0.25, legacy=legacy)
    assert np.array_equal(idx, g[p + "_codes"].reshape(-1))
    assert np.array_equal(_unflat(xq, g["q2_x"]), g[p + "_xq"])
    assert abs(float(loss) - float(g[p + "_loss"])) <= 1e-6 * abs(float(g[p + "_loss"]))
    assert np.array_equal(g[p + "_gx_ste"], g["q2_gq"])
    assert np.allclose(_unflat(gx, g["q2_x"]), g[p + "_gx_loss"], rtol=1e-5, atol=1e-10)


def test_family_quantize2_list_oracle_matches_reference():
    g = _load("vq_family.npz")
    n = int(g["ql_n_items"])
    xs = [g[f"ql_{i}_x"] for i in range(n)]
    xq, loss, idx, _ = vf.vq2_list_forward(xs, g["ql_weight"], beta=0.25)
    for i in range(n):
        assert np.array_equal(idx[i], g[f"ql_{i}_codes"])
        assert np.array_equal(xq[i], g[f"ql_{i}_xq"])
        e = g["ql_weight"][idx[i]]
        assert np.allclose(g[f"ql_{i}_gx"], 2 * 0.25 * (xs[i] - e) / xs[i].size / n, rtol=1e-5, atol=1e-10)
    assert abs(float(loss) - float(g["ql_loss"])) <= 1e-6 * abs(float(g["ql_loss"]))
    # training: the codebook moves between items
    k = g["ql_weight"].shape[0] - 1
    xq, loss, idx, (w, cs, em) = vf.vq2_list_forward(
        xs, g["ql_weight"], beta=0.25, train=True, cs=np.ones(k, np.float32), em=g["ql_weight"][:-1].copy(),
        restart_rows=[g[f"ql_train_{i}_restart"] for i in range(n)])
    for i in range(n):
        assert np.array_equal(idx[i], g[f"ql_train_{i}_codes"])
        assert np.allclose(xq[i], g[f"ql_train_{i}_xq"], rtol=0, atol=1e-6)
    assert np.allclose(w, g["ql_train_w"], rtol=1e-5, atol=1e-6)
    assert np.allclose(cs, g["ql_train_cs"], rtol=1e-6, atol=1e-7)
    assert np.allclose(em, g["ql_train_em"], rtol=1e-5, atol=1e-6)
    assert abs(float(loss) - float(g["ql_train_loss"])) <= 1e-5 * abs(float(g["ql_train_loss"]))


@pytest.mark.parametrize("tag,latent,code,shared", [("rq", (8, 8, 64), (8, 8, 3), False),
                                                     ("rqs", (8, 8, 64), (4, 4, 2), True)])
def test_family_rq_oracle_matches_reference(tag, latent, code, shared):
    g = _load("vq_family.npz")
    depth = code[2]
    ws = [g[f"{tag}_w{d}_v"].copy() for d in range(depth)]
    if shared:
        ws = [ws[0]] * depth
    q, loss, codes, gx, _ = vf.rq_forward(g[f"{tag}_x"], ws, latent, code)
    assert np.array_equal(codes, g[f"{tag}_codes"])
    assert np.allclose(q, g[f"{tag}_quants"], rtol=0, atol=2e-6)
    assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 1e-6 * abs(float(g[f"{tag}_loss"]))
    assert np.array_equal(g[f"{tag}_gx_ste"], g[f"{tag}_gq"])
    assert np.allclose(vf.rq_to_latent_shape(gx, latent, code), g[f"{tag}_gx_loss"], rtol=1e-4, atol=1e-9)
    assert np.allclose(vf.rq_embed_code(codes, ws, latent, code), g[f"{tag}_embed"], rtol=0, atol=2e-6)
    # one training pass, every depth updating its codebook right after its own search
    k = ws[0].shape[0] - 1
    states = [(np.ones(k, np.float32), ws[d][:-1].copy()) for d in range(depth)]
    if shared:
        states = [states[0]] * depth
    q, loss, codes, _, states = vf.rq_forward(
        g[f"{tag}_train_x"], ws, latent, code, train=True, states=list(states),
        restart_rows=[g[f"{tag}_train_d{d}_restart"] for d in range(depth)])
    assert np.array_equal(codes, g[f"{tag}_train_codes"])
    assert np.allclose(q, g[f"{tag}_train_quants"], rtol=0, atol=2e-6)
    assert abs(float(loss) - float(g[f"{tag}_train_loss"])) <= 1e-5 * abs(float(g[f"{tag}_train_loss"]))
    for d in range(depth):
        assert np.allclose(ws[d], g[f"{tag}_train_d{d}_w"], rtol=1e-5, atol=1e-6), d
        assert np.allclose(states[d][0], g[f"{tag}_train_d{d}_cs"], rtol=1e-6, atol=1e-7), d
        assert np.allclose(states[d][1], g[f"{tag}_train_d{d}_em"], rtol=1e-5, atol=1e-6), d


@pytest.mark.parametrize("legacy", [True, False])
def test_family_vqgan_oracle_matches_reference(legacy):
    g = _load("vq_family.npz")
    p = f"vg_legacy{int(legacy)}"
    z = _flat(g["vg_z"])
    zq, loss, idx, gz, gw = vf.vqgan_forward(z, g["vg_weight"], beta=0.25, legacy=legacy)
    assert np.array_equal(idx, g[p + "_idx"].reshape(-1))
    assert np.array_equal(_unflat(zq, g["vg_z"]), g[p + "_zq"])
    assert abs(float(loss) - float(g[p + "_loss"])) <= 1e-6 * abs(float(g[p + "_loss"]))
    assert np.array_equal(g[p + "_gz_ste"], g["vg_gq"])
    assert np.allclose(_unflat(gz, g["vg_z"]), g[p + "_gz_loss"], rtol=1e-5, atol=1e-10)
    assert np.allclose(gw, g[p + "_gw"], rtol=1e-4, atol=1e-8)
    assert np.allclose(_unflat(g["vg_weight"][idx], g["vg_z"]), g[p + "_entry"], rtol=0, atol=0)


# --------------------------------------------------------------------------- stage-2 permuter (8f row 3)
from oracle import permuter_oracle as po  # noqa: E402

PERMUTER_CASES = [("p8", 4, 8), ("p32", 16, 32), ("p8b", 4, 8), ("p32b", 16, 32)]


@pytest.mark.parametrize("tag,hw1,fhw", PERMUTER_CASES)
@pytest.mark.parametrize("order", ["region-first", "row-first"])
def test_permuter_oracle_matches_reference(tag, hw1, fhw, order):
    g = _load("permuter.npz")
    codes = dict(coarse_position_pad_code=hw1 * hw1, coarse_position_eos_code=hw1 * hw1 + 1)
    o = po.forward(g[f"{tag}_indices"], g[f"{tag}_grain"], hw1, fhw, order, **codes)
    k = order[:3]
    for name, v in o.items():
        assert np.array_equal(v, g[f"{tag}_{k}_{name}"]), name
    back = po.forward_back(o["coarse_content"], o["fine_content"], o["coarse_position"], o["fine_position"],
                           hw1, fhw, **codes)
    assert np.array_equal(back, g[f"{tag}_{k}_back"]) and np.array_equal(back, g[f"{tag}_indices"])


def test_permuter_oracle_backward_edge_cases_match_reference():
    g = _load("permuter.npz")
    back = po.forward_back(g["pb_cc"], g["pb_fc"], g["pb_cp"], g["pb_fp"], 4, 8,
                           coarse_position_pad_code=16, coarse_position_eos_code=17)
    assert np.array_equal(back, g["pb_back"])
    assert back[0, 1, 1] == 41 and back[0, 0, 0] == 7 and back[1].sum() == 60 + 61   # last wins; no eos -> no spread


# ------------------------------------------------------------------------------------ training loss
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_loss_oracle_matches_reference(mode):
    """oracle.loss_oracle (LPIPS + PatchGAN + adaptive weight) vs the reference's VQLPIPSWithDiscriminator run in
    the build container on the same seeded weights (tests/golden/make_golden.py::loss_goldens)."""
    import torch.nn.functional as F
    from oracle import loss_oracle as lo
    g = np.load(os.path.join(G, "loss_small.npz"))
    sd = lo.make_loss_weights(seed=21)
    x, feat, w_last, qloss, gate = lo.toy_inputs()
    w_last.requires_grad_(True)
    feat.requires_grad_(True)
    xrec = F.conv2d(feat, w_last, padding=1)
    assert np.allclose(xrec.detach().numpy(), g[mode + "_xrec"], atol=1e-6)
    train = mode == "train"
    stats0 = {}
    l0, log0 = lo.loss_forward(sd, qloss, x, xrec, 0, 0, last_layer=w_last, gate=gate, disc_weight_max=0.75,
                               train=train, budget=lo.budget_loss_dual, new_stats=stats0)
    gw, gf = torch.autograd.grad(l0, [w_last, feat])

    def close(a, key, rtol=2e-4, atol=1e-7):
        b = g[mode + "_" + key]
        a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
        assert np.allclose(a, b, rtol=rtol, atol=atol), (key, float(np.abs(a - b).max()), float(np.abs(b).max()))
    close(l0, "loss0")
    close(log0["p_loss"], "log0_p_loss")
    close(log0["d_weight"], "log0_d_weight", rtol=1e-3)
    close(log0["g_loss"], "log0_g_loss")
    close(log0["budget_loss"], "log0_budget_loss")
    close(gw, "g_w_last", rtol=2e-3, atol=1e-6 * float(np.abs(g[mode + "_g_w_last"]).max()) + 1e-9)
    close(gf, "g_feat", rtol=2e-3, atol=2e-3 * float(np.abs(g[mode + "_g_feat"]).max()))
    close(lo.lpips(sd, x, xrec.detach()), "lpips")
    if train:
        for k, v in stats0.items():
            close(v, "bn0_" + k[len("loss."):], rtol=1e-4, atol=1e-6)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "discriminator" in k and "running" not in k}
    sd1 = dict(sd); sd1.update(params)
    l1, log1 = lo.loss_forward(sd1, qloss, x, xrec.detach(), 1, 0, train=train)
    close(l1, "loss1")
    close(log1["logits_real"], "log1_logits_real", atol=1e-6)
    close(log1["logits_fake"], "log1_logits_fake", atol=1e-6)
    names = list(params)
    gd = torch.autograd.grad(l1, [params[k] for k in names])
    for k, gr in zip(names, gd):
        short = k[len("loss.discriminator."):]
        close(gr.norm(), "gd_norm_" + short, rtol=2e-3, atol=1e-7)
    close(gd[names.index("loss.discriminator.main.0.weight")], "gd_main.0.weight", rtol=2e-3,
          atol=2e-3 * float(np.abs(g[mode + "_gd_main.0.weight"]).max()))


def _threshold_inputs(seed=3, n_batches=3, b=4, res=64, patch=16):
    g = torch.Generator().manual_seed(seed)
    k = res // patch
    out = []
    for _ in range(n_batches):
        x = torch.rand(b, 3, res, res, generator=g)
        flat = torch.rand(b, 3, k, k, generator=g).repeat_interleave(patch, 2).repeat_interleave(patch, 3)
        pick = (torch.rand(b, 1, k, k, generator=g) > 0.5).float().repeat_interleave(patch, 2).repeat_interleave(patch, 3)
        out.append(pick * x + (1 - pick) * flat)
    return out


def test_entropy_threshold_oracle_matches_reference_tool():
    """orc.entropy_thresholds vs the reference's scripts/tools/calculate_entropy_thresholds.py procedure."""
    g = np.load(os.path.join(G, "entropy_thresholds.npz"))
    th = orc.entropy_thresholds(_threshold_inputs(), patch=16)
    got = np.array([th[str(i)] for i in range(1, 100)], dtype=np.float32)
    assert np.allclose(got, g["thresholds"], rtol=1e-6, atol=1e-7)
