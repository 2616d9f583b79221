"""GPU parity of the stage-2 tokenisation path: csrc/permuter.cu through the
reference-facing DualGrainSeperatePermuter vs the numpy oracle and the goldens minted from the
reference class - bit-exact (int64 index work) - and the encode -> permute -> un-permute -> decode
round trip through the dual-grain model."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _g():
    return {k: v for k, v in np.load(os.path.join(G, "permuter.npz"), allow_pickle=False).items()}


def _permuter(hw1, fhw, order):
    from dynamicvectorquantization_b200 import configs
    configs.activate_overlay()
    from modules.dynamic_modules.permuter import DualGrainSeperatePermuter
    return DualGrainSeperatePermuter(coarse_hw=hw1, fine_hw=fhw, coarse_position_pad_code=hw1 * hw1,
                                     coarse_position_eos_code=hw1 * hw1 + 1, fine_position_order=order)


@pytest.mark.parametrize("tag,hw1,fhw", [("p8", 4, 8), ("p32", 16, 32), ("p8b", 4, 8), ("p32b", 16, 32)])
@pytest.mark.parametrize("order", ["region-first", "row-first"])
def test_permuter_matches_reference_golden(tag, hw1, fhw, order):
    g = _g()
    p = _permuter(hw1, fhw, order)
    idx = torch.from_numpy(g[f"{tag}_indices"]).cuda()
    grain = torch.from_numpy(g[f"{tag}_grain"]).cuda()
    o = p(idx, grain)
    k = order[:3]
    for name in ("coarse_content", "fine_content", "coarse_position", "fine_position", "coarse_segment",
                 "fine_segment"):
        assert o[name].dtype == torch.int64 and o[name].is_contiguous()
        assert np.array_equal(o[name].cpu().numpy(), g[f"{tag}_{k}_{name}"]), name
    back = p.forward_back(o["coarse_content"], o["fine_content"], o["coarse_position"], o["fine_position"])
    assert torch.equal(back, idx)
    with pytest.raises(RuntimeError):
        p(idx.cpu(), grain.cpu())


def test_permuter_backward_edge_cases_match_reference():
    g = _g()
    p = _permuter(4, 8, "region-first")
    back = p.forward_back(*[torch.from_numpy(g[k]).cuda() for k in ("pb_cc", "pb_fc", "pb_cp", "pb_fp")])
    assert np.array_equal(back.cpu().numpy(), g["pb_back"])      # duplicates: last wins; no eos: no spread


@pytest.mark.parametrize("order", ["region-first", "row-first"])
@pytest.mark.parametrize("hw1,hw2,batch", [(16, 2, 64), (8, 4, 7), (32, 2, 3), (1, 2, 2)])
def test_permuter_matches_oracle_random(order, hw1, hw2, batch):
    """Sizes beyond the goldens (up to a 64x64 fine map = four 1024-thread chunks, 4x4 sub-cells, a single
    cell), random grains incl. values outside {0,1} (neither sequence takes them), against the oracle."""
    from oracle import permuter_oracle as po
    fhw = hw1 * hw2
    g = torch.Generator().manual_seed(hw1 * 100 + hw2)
    grain = torch.randint(0, 3 if hw1 == 8 else 2, (batch, hw1, hw1), generator=g)
    idx = torch.randint(0, 1024, (batch, fhw, fhw), generator=g)
    p = _permuter(hw1, fhw, order)
    codes = dict(coarse_position_pad_code=hw1 * hw1, coarse_position_eos_code=hw1 * hw1 + 1,
                 fine_position_pad_code=max(1024, fhw * fhw), fine_position_eos_code=max(1024, fhw * fhw) + 1)
    p.fine_position_pad_code, p.fine_position_eos_code = codes["fine_position_pad_code"], codes["fine_position_eos_code"]
    o = p(idx.cuda(), grain.cuda())
    ref = po.forward(idx.numpy(), grain.numpy(), hw1, fhw, order, **codes)
    for name, v in ref.items():
        assert np.array_equal(o[name].cpu().numpy(), v), name
    back = p.forward_back(o["coarse_content"], o["fine_content"], o["coarse_position"], o["fine_position"])
    ref_back = po.forward_back(ref["coarse_content"], ref["fine_content"], ref["coarse_position"],
                               ref["fine_position"], hw1, fhw, **codes)
    assert np.array_equal(back.cpu().numpy(), ref_back)


def test_tokenise_roundtrip_through_the_dual_model():
    """dqtransformer_uncond_entropy.py:167-178: encode -> permuter -> forward_back -> code embeddings ->
    decode reproduces the direct reconstruction (eval mode, small dual config)."""
    from dynamicvectorquantization_b200 import configs
    from oracle import dqvae_oracle as orc
    ocfg = orc.SMALL_CFG
    model = configs.build_model(configs.scaled_dual_config())
    sd = orc.make_weights(orc.model_shapes(ocfg), seed=11)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    x = (torch.rand(3, 3, 64, 64, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    with torch.no_grad():
        quant, _, info, grain_indices, gate = model.encode(x)
        codes = info[2]                                                     # [B, 8, 8]
        p = _permuter(4, 8, "region-first")
        z = p(codes, grain_indices)
        back = p.forward_back(z["coarse_content"], z["fine_content"], z["coarse_position"], z["fine_position"])
        assert torch.equal(back, codes)          # coarse cells carry one code, so the map is reproduced exactly
        n_coarse = int((grain_indices == 0).sum(dim=(1, 2)).max())
        assert z["coarse_content"].shape[1] == n_coarse + 1
        emb = model.get_code_emb_with_depth(back)                          # [B, 8, 8, C]
        rec = model.decode(emb.permute(0, 3, 1, 2).contiguous())
        direct = model.decode(quant)
    assert torch.allclose(rec, direct, atol=2e-2), float((rec - direct).abs().max())
