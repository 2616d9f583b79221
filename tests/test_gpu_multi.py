"""Multi-GPU (NCCL, one process per GPU) checks of the data-parallel path; skipped with < 2 GPUs.

Invariants: after a training step on different shards, every rank holds bit-identical
codebook state (weight / cluster_size_ema / embed_ema) and identical parameters, and the EMA
statistics equal those of a single process that saw the concatenated batch."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from dynamicvectorquantization_b200 import configs
    from oracle import dqvae_oracle as orc
    torch.manual_seed(7)
    model = configs.build_model(configs.scaled_dual_config())
    sd = orc.make_weights(orc.model_shapes(orc.SMALL_CFG), seed=2)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().train()
    for p in model.loss.parameters():
        p.requires_grad_(False)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
    params = [p for n, p in model.named_parameters() if not n.startswith("loss.") and p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, betas=(0.5, 0.9))
    g = torch.Generator().manual_seed(100 + rank)
    x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        xrec, qloss, indices, gate = ddp(x)
        loss, _ = model.loss(qloss, x, xrec, 0, 0, last_layer=None, split="train", gate=gate)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    cb = model.quantize.codebook
    flat = torch.cat([cb.weight.flatten(), cb.cluster_size_ema, cb.embed_ema.flatten(),
                      model.decoder.conv_out.weight.flatten(), model.encoder.conv_in.weight.flatten()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered[1:])
    finite = bool(torch.isfinite(flat).all()) and bool(torch.isfinite(loss))
    ema_ok, ema_msg = _concatenated_batch_ema_check(rank, world)
    q.put((rank, same, finite, float(cb.cluster_size_ema.sum()), ema_ok, ema_msg))
    dist.destroy_process_group()


def _concatenated_batch_ema_check(rank, world):
    """quantize2_mask.py:86-105 under data parallelism: each rank quantizes ITS shard in training mode; afterwards
    every rank's EMA state must equal what the numpy oracle computes in a single process from the concatenated
    batch (all ranks' rows and codes) with rank 0's restart rows (the reference broadcasts them from rank 0)."""
    import numpy as np
    import torch.distributed as dist
    from modules.vector_quantization.quantize2_mask import VectorQuantize2
    from oracle import vq_oracle as vo
    K, C, B, H = 96, 64, 2, 12
    vq = VectorQuantize2(codebook_size=K, codebook_dim=C).cuda().train()
    g = torch.Generator().manual_seed(3)
    w0 = torch.randn(K + 1, C, generator=g)
    with torch.no_grad():
        vq.codebook.weight.copy_(w0)
        vq.codebook.embed_ema.copy_(w0[:-1])
        vq.codebook.cluster_size_ema.fill_(1.0)
    x = torch.randn(B, C, H, H, generator=torch.Generator().manual_seed(500 + rank)).cuda()
    torch.manual_seed(50)
    perm = torch.randperm(B * H * H, device="cuda")                 # what the module draws for the restart rows
    torch.manual_seed(50)
    _, _, (_, _, codes) = vq(x)
    rows = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    all_rows = [torch.empty_like(rows) for _ in range(world)]
    all_codes = [torch.empty_like(codes.reshape(-1)) for _ in range(world)]
    dist.all_gather(all_rows, rows)
    dist.all_gather(all_codes, codes.reshape(-1).contiguous())
    restart = all_rows[0][perm][:K].cpu().numpy()                    # rank 0's candidates, same perm on every rank
    cat_rows = torch.cat(all_rows).cpu().numpy()
    cat_codes = torch.cat(all_codes).cpu().numpy()
    w = w0.numpy().copy()
    cs, em = vo.update_buffers(cat_rows, cat_codes, np.ones(K, np.float32), w[:-1].copy(), 0.99, restart_rows=restart)
    w[:-1] = vo.update_embedding(cs, em)
    cb = vq.codebook
    ok = (np.allclose(cb.cluster_size_ema.cpu().numpy(), cs, rtol=1e-5, atol=1e-6)
          and np.allclose(cb.embed_ema.cpu().numpy(), em, rtol=1e-4, atol=1e-5)
          and np.allclose(cb.weight.detach().cpu().numpy(), w, rtol=1e-4, atol=1e-5))
    restarted = int((cs == 1.0).sum())
    return ok, f"rank {rank}: {restarted} of {K} codes restarted, max |cs diff| " \
               f"{float(np.abs(cb.cluster_size_ema.cpu().numpy() - cs).max()):.2e}"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_ddp_keeps_codebook_and_params_identical():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=600) for _ in range(world))
    [p.join(60) for p in procs]
    for rank, same, finite, total, ema_ok, ema_msg in res:
        assert same, f"rank {rank}: codebook/parameters diverged across ranks"
        assert finite
        print(ema_msg)
        assert ema_ok, f"EMA state differs from a single process on the concatenated batch ({ema_msg})"
    assert abs(res[0][3] - res[1][3]) == 0
