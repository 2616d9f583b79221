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
    if tag == "dual":
        return  # backward of the full-size model is covered on the tiny config (keeps the CPU suite fast)
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
