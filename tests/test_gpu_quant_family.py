"""GPU parity of the sibling quantizers through their reference-facing interface:
the CUDA modules vs the numpy oracle searched on the same bf16-rounded operands (codes bit-exact), and
vs the goldens minted from the reference's own classes (tests/golden/vq_family.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


def _g():
    return {k: v for k, v in np.load(os.path.join(G, "vq_family.npz"), allow_pickle=False).items()}


def _overlay():
    from dynamicvectorquantization_b200 import configs
    configs.activate_overlay()


def _flat(x_nchw):
    b, c, h, w = x_nchw.shape
    return np.ascontiguousarray(x_nchw.transpose(0, 2, 3, 1).reshape(-1, c))


def _unflat(rows, like_nchw):
    b, c, h, w = like_nchw.shape
    return rows.reshape(b, h, w, c).transpose(0, 3, 1, 2)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _load_codebook(cb, w):
    with torch.no_grad():
        cb.weight.copy_(_t(w))
        cb.embed_ema.copy_(_t(w[:-1]))
        cb.cluster_size_ema.fill_(1.0)


@pytest.mark.parametrize("legacy", [True, False])
def test_quantize2_matches_oracle_and_golden(legacy):
    _overlay()
    from modules.vector_quantization.quantize2 import VectorQuantize2
    from oracle import vq_family_oracle as vf
    g = _g()
    p = f"q2_legacy{int(legacy)}"
    vq = VectorQuantize2(codebook_size=64, codebook_dim=64, commit_loss_legacy=legacy).cuda().eval()
    _load_codebook(vq.codebook, g["q2_weight"])
    x = _t(g["q2_x"]).requires_grad_(True)
    xq, loss, (_, _, codes) = vq(x)
    rows = _flat(g["q2_x"])
    rxq, rloss, ridx, rgx = vf.vq2_forward(rows, g["q2_weight"], 0.25, legacy, search_bf16=True)
    assert np.array_equal(codes.reshape(-1).cpu().numpy(), ridx)
    assert codes.shape == (2, 8, 8) and codes.dtype == torch.int64
    assert np.allclose(xq.detach().cpu().numpy(), _unflat(g["q2_weight"][ridx], g["q2_x"]), atol=1e-6)
    assert abs(float(loss) - float(rloss)) < 1e-5 * float(rloss)
    gq = _t(g["q2_gq"])
    (xq * gq).sum().backward(retain_graph=True)
    assert torch.equal(x.grad, gq)
    x.grad = None
    loss.backward()
    assert np.allclose(x.grad.cpu().numpy(), _unflat(rgx, g["q2_x"]), rtol=1e-4, atol=1e-9)
    if np.array_equal(ridx, g[p + "_codes"].reshape(-1)):           # no bf16 near-tie on this input
        assert abs(float(loss) - float(g[p + "_loss"])) < 1e-5 * float(g[p + "_loss"])
        assert np.allclose(x.grad.cpu().numpy(), g[p + "_gx_loss"], rtol=1e-4, atol=1e-9)
    # sequence input [B, C, N] (accept_image_fmap=False, channel_last=False): transposed in and out
    vs = VectorQuantize2(codebook_size=64, codebook_dim=64, accept_image_fmap=False,
                         commit_loss_legacy=legacy).cuda().eval()
    _load_codebook(vs.codebook, g["q2_weight"])
    xs = _t(g["q2_x"]).reshape(2, 64, 64)
    xq_s, loss_s, (_, _, codes_s) = vs(xs)
    assert xq_s.shape == (2, 64, 64) and torch.equal(codes_s.reshape(-1), codes.reshape(-1))
    assert torch.allclose(xq_s, xq.detach().reshape(2, 64, 64)) and abs(float(loss_s) - float(loss)) < 1e-6
    with pytest.raises(RuntimeError):
        vq(torch.from_numpy(g["q2_x"]))                             # no CPU fallback


def test_quantize2_list_ragged_eval_and_training():
    _overlay()
    from modules.vector_quantization.quantize2_list import VectorQuantize2
    from oracle import vq_family_oracle as vf
    g = _g()
    n = int(g["ql_n_items"])
    K = g["ql_weight"].shape[0] - 1
    xs = [g[f"ql_{i}_x"] for i in range(n)]
    vq = VectorQuantize2(codebook_size=K, codebook_dim=64).cuda().eval()
    _load_codebook(vq.codebook, g["ql_weight"])
    xin = [_t(a).requires_grad_(True) for a in xs]
    xq_l, loss, (_, _, code_l) = vq(xin)
    rxq, rloss, ridx, _ = vf.vq2_list_forward(xs, g["ql_weight"], 0.25, search_bf16=True)
    loss.backward()
    same_as_ref = True
    for i in range(n):
        assert code_l[i].shape == (xs[i].shape[0],)
        assert np.array_equal(code_l[i].cpu().numpy(), ridx[i]), i
        assert np.allclose(xq_l[i].detach().cpu().numpy(), g["ql_weight"][ridx[i]], atol=1e-6)
        e = g["ql_weight"][ridx[i]]
        assert np.allclose(xin[i].grad.cpu().numpy(), 2 * 0.25 * (xs[i] - e) / xs[i].size / n, rtol=1e-4, atol=1e-9)
        same_as_ref &= np.array_equal(ridx[i], g[f"ql_{i}_codes"])
    assert abs(float(loss) - float(rloss)) < 1e-5 * float(rloss)
    if same_as_ref:
        assert abs(float(loss) - float(g["ql_loss"])) < 1e-5 * float(g["ql_loss"])
    # training: replay the CUDA RNG draws of the per-item updates (tile-with-noise when n_i < K, then randperm)
    vq.train()
    dev_xs = [_t(a) for a in xs]
    torch.manual_seed(99)
    restart = []
    for t in dev_xs:
        v = t
        if v.shape[0] < K:
            v = vq.codebook._tile_with_noise(v, K)
        restart.append(v[torch.randperm(v.shape[0], device="cuda")][:K].cpu().numpy())
    torch.manual_seed(99)
    xq_t, loss_t, (_, _, code_t) = vq(dev_xs)
    rxq, rloss, ridx, (w, cs, em) = vf.vq2_list_forward(
        xs, g["ql_weight"], 0.25, train=True, cs=np.ones(K, np.float32), em=g["ql_weight"][:-1].copy(),
        restart_rows=restart, search_bf16=True)
    for i in range(n):
        assert np.array_equal(code_t[i].cpu().numpy(), ridx[i]), i
    assert np.allclose(vq.codebook.cluster_size_ema.cpu().numpy(), cs, rtol=1e-5, atol=1e-6)
    assert np.allclose(vq.codebook.embed_ema.cpu().numpy(), em, rtol=1e-4, atol=1e-5)
    assert np.allclose(vq.codebook.weight.detach().cpu().numpy(), w, rtol=1e-4, atol=1e-5)
    assert abs(float(loss_t) - float(rloss)) < 1e-4 * float(rloss)


@pytest.mark.parametrize("tag,latent,code,shared", [("rq", (8, 8, 64), (8, 8, 3), False),
                                                     ("rqs", (8, 8, 64), (4, 4, 2), True)])
def test_rq_bottleneck_matches_oracle_and_golden(tag, latent, code, shared):
    _overlay()
    from modules.vector_quantization.quantize_rqvae import RQBottleneck
    from oracle import vq_family_oracle as vf
    g = _g()
    depth, K = code[2], 32
    rq = RQBottleneck(latent_shape=latent, code_shape=code, n_embed=K, shared_codebook=shared).cuda().eval()
    ws = [g[f"{tag}_w{d}_v"].copy() for d in range(depth)]
    if shared:
        ws = [ws[0]] * depth
    for d in range(1 if shared else depth):
        _load_codebook(rq.codebooks[d], ws[d])
    x = _t(g[f"{tag}_x"]).requires_grad_(True)
    q, loss, codes = rq(x)
    rq_, rloss, rcodes, rgx, _ = vf.rq_forward(g[f"{tag}_x"], ws, latent, code, search_bf16=True)
    assert np.array_equal(codes.cpu().numpy(), rcodes) and codes.dtype == torch.int64
    assert np.allclose(q.detach().cpu().numpy(), rq_, atol=3e-6)
    assert abs(float(loss) - float(rloss)) < 1e-5 * float(rloss)
    gq = _t(g[f"{tag}_gq"])
    (q * gq).sum().backward(retain_graph=True)
    assert torch.equal(x.grad, gq)
    x.grad = None
    loss.backward()
    assert np.allclose(x.grad.cpu().numpy(), vf.rq_to_latent_shape(rgx, latent, code), rtol=1e-4, atol=1e-9)
    if np.array_equal(rcodes, g[f"{tag}_codes"]):
        assert np.allclose(q.detach().cpu().numpy(), g[f"{tag}_quants"], atol=3e-6)
        assert abs(float(loss) - float(g[f"{tag}_loss"])) < 1e-5 * float(g[f"{tag}_loss"])
        assert np.allclose(x.grad.cpu().numpy(), g[f"{tag}_gx_loss"], rtol=1e-4, atol=1e-9)
    assert np.allclose(rq.embed_code(codes).cpu().numpy(), vf.rq_embed_code(rcodes, ws, latent, code), atol=3e-6)
    ql, codes2 = rq.quantize(rq.to_code_shape(x.detach()))
    assert torch.equal(codes2, codes) and len(ql) == depth
    emb_d, _ = rq.embed_code_with_depth(codes)
    assert emb_d.shape == codes.shape + (ws[0].shape[1],)
    part = rq.embed_partial_code(codes, depth - 1, decode_type="add")
    assert torch.allclose(part, rq.embed_code(codes), atol=1e-6)
    # one training pass: depth d draws randperm over its residual rows right after its own search
    rq.train()
    xt = g[f"{tag}_train_x"]
    n_rows = 2 * code[0] * code[1]
    torch.manual_seed(7)
    perms = [torch.randperm(n_rows, device="cuda").cpu().numpy() for _ in range(depth)]
    torch.manual_seed(7)
    q_t, loss_t, codes_t = rq(_t(xt))
    states = [(np.ones(K, np.float32), ws[d][:-1].copy()) for d in range(depth)]
    if shared:
        states = [states[0]] * depth
    rr = [(lambda res, p=perms[d]: res[p][:K]) for d in range(depth)]
    rq_t, rloss_t, rcodes_t, _, states = vf.rq_forward(xt, ws, latent, code, train=True, states=list(states),
                                                       restart_rows=rr, search_bf16=True)
    assert np.array_equal(codes_t.cpu().numpy(), rcodes_t)
    assert abs(float(loss_t) - float(rloss_t)) < 1e-4 * float(rloss_t)
    for d in range(depth):
        cb = rq.codebooks[d]
        assert np.allclose(cb.cluster_size_ema.cpu().numpy(), states[d][0], rtol=1e-5, atol=1e-6), d
        assert np.allclose(cb.embed_ema.cpu().numpy(), states[d][1], rtol=1e-4, atol=1e-5), d
        assert np.allclose(cb.weight.detach().cpu().numpy(), ws[d], rtol=1e-4, atol=1e-5), d


@pytest.mark.parametrize("legacy", [True, False])
def test_vqgan_quantizer_learnable_codebook(legacy, tmp_path):
    _overlay()
    from modules.vector_quantization.quantize_vqgan import VectorQuantizer2
    from oracle import vq_family_oracle as vf
    g = _g()
    p = f"vg_legacy{int(legacy)}"
    q = VectorQuantizer2(64, 64, beta=0.25, legacy=legacy, sane_index_shape=not legacy).cuda()
    with torch.no_grad():
        q.embedding.weight.copy_(_t(g["vg_weight"]))
    z = _t(g["vg_z"]).requires_grad_(True)
    zq, loss, (_, _, idx) = q(z)
    rows = _flat(g["vg_z"])
    rzq, rloss, ridx, rgz, rgw = vf.vqgan_forward(rows, g["vg_weight"], 0.25, legacy, search_bf16=True)
    assert idx.shape == ((2, 8, 8) if not legacy else (128,))
    assert np.array_equal(idx.reshape(-1).cpu().numpy(), ridx)
    assert np.allclose(zq.detach().cpu().numpy(), _unflat(g["vg_weight"][ridx], g["vg_z"]), atol=1e-6)
    assert abs(float(loss) - float(rloss)) < 1e-5 * float(rloss)
    gq = _t(g["vg_gq"])
    (zq * gq).sum().backward(retain_graph=True)
    assert torch.equal(z.grad, gq)
    assert q.embedding.weight.grad is None or float(q.embedding.weight.grad.abs().max()) == 0.0
    z.grad = None
    q.embedding.weight.grad = None
    loss.backward()
    assert np.allclose(z.grad.cpu().numpy(), _unflat(rgz, g["vg_z"]), rtol=1e-4, atol=1e-9)
    assert np.allclose(q.embedding.weight.grad.cpu().numpy(), rgw, rtol=1e-3, atol=1e-8)
    if np.array_equal(ridx, g[p + "_idx"].reshape(-1)):
        assert abs(float(loss) - float(g[p + "_loss"])) < 1e-5 * float(g[p + "_loss"])
        assert np.allclose(q.embedding.weight.grad.cpu().numpy(), g[p + "_gw"], rtol=1e-3, atol=1e-8)
        assert np.allclose(z.grad.cpu().numpy(), g[p + "_gz_loss"], rtol=1e-4, atol=1e-9)
    ent = q.get_codebook_entry(idx.reshape(-1), (2, 8, 8, 64))
    assert np.allclose(ent.detach().cpu().numpy(), g[p + "_entry"] if np.array_equal(ridx, g[p + "_idx"].reshape(-1))
                       else _unflat(g["vg_weight"][ridx], g["vg_z"]), atol=0)
    # the optimizer moves the codebook and the next search sees the new rows
    opt = torch.optim.SGD(q.parameters(), lr=50.0)
    opt.step()
    w1 = q.embedding.weight.detach().cpu().numpy()
    _, _, (_, _, idx1) = q(z.detach())
    _, _, ridx1, _, _ = vf.vqgan_forward(rows, w1, 0.25, legacy, search_bf16=True)
    assert np.array_equal(idx1.reshape(-1).cpu().numpy(), ridx1)
    # index remapping (:247-269): only the listed codes are "used", the rest map to the extra token
    used = np.arange(0, 64, 2)
    np.save(tmp_path / "used.npy", used)
    qr = VectorQuantizer2(64, 64, beta=0.25, remap=str(tmp_path / "used.npy"), unknown_index="extra").cuda()
    with torch.no_grad():
        qr.embedding.weight.copy_(_t(g["vg_weight"]))
    _, _, (_, _, ridx_m) = qr(z.detach())
    exp = np.where(ridx % 2 == 0, ridx // 2, len(used))
    assert np.array_equal(ridx_m.reshape(-1).cpu().numpy(), exp)
