"""One training step of the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (launch list or full capture of selected kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dynamicvectorquantization_b200 import configs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REAL = len(sys.argv) > 2 and sys.argv[2] == "real"     # the reference's full loss (LPIPS + PatchGAN), both passes
if REAL:
    os.environ.setdefault("B200DQ_ALLOW_RANDOM_VGG", "1")
torch.manual_seed(2021)
cfg = configs.stage1_config("dqvae-dual-r-05")
if REAL:
    cfg["params"]["lossconfig"] = configs.real_loss_config(configs._BUDGET_DUAL)
model = configs.build_model(cfg).cuda().train()
if not REAL:
    for p in model.loss.parameters():
        p.requires_grad_(False)
else:
    opt_d = torch.optim.Adam(model.loss.discriminator.parameters(), lr=1e-4, betas=(0.5, 0.9), fused=True)
params = [p for n, p in model.named_parameters() if not n.startswith("loss.") and p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-4, betas=(0.5, 0.9), fused=True)   # as bench.py
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1


def step():
    opt.zero_grad(set_to_none=True)
    xrec, qloss, indices, gate = model(x)
    loss, _ = model.loss(qloss, x, xrec, 0, 0, last_layer=model.get_last_layer() if REAL else None, split="train",
                         gate=gate)
    loss.backward()
    opt.step()
    if REAL:
        opt_d.zero_grad(set_to_none=True)
        xrec, qloss, indices, gate = model(x)
        l1, _ = model.loss(qloss, x, xrec, 1, 0, last_layer=model.get_last_layer(), split="train")
        l1.backward()
        opt_d.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, batch", B)
