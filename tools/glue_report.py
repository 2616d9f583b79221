"""Which Python lines of the package launch non-b2 (ATen / cuBLAS) kernels in one training step, and how much GPU
time those kernels take: torch.profiler with stacks, kernels attributed to the innermost package frame."""
import os
import sys
import collections

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity

from dynamicvectorquantization_b200 import configs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(2021)
cfg = configs.stage1_config("dqvae-dual-r-05")
model = configs.build_model(cfg).cuda().train()
for p in model.loss.parameters():
    p.requires_grad_(False)
params = [p for n, p in model.named_parameters() if not n.startswith("loss.") and p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-4, betas=(0.5, 0.9), fused=True)
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1


def step():
    opt.zero_grad(set_to_none=True)
    xrec, qloss, indices, gate = model(x)
    loss, _ = model.loss(qloss, x, xrec, 0, 0, last_layer=None, split="train", gate=gate)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()

by_site = collections.defaultdict(lambda: [0, 0.0, set()])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU:
        continue
    ks = [k for k in ev.kernels if "b2::" not in k.name]
    if not ks:
        continue
    # leaf aten ops only (parents repeat the kernels of their children)
    if any(c.kernels for c in ev.cpu_children):
        continue
    site = "?"
    for fr in ev.stack:
        if "dynamicvectorquantization_b200" in fr or "profile_step" in fr or "glue_report" in fr:
            site = fr.strip()
            break
    if site == "?":
        site = str([list(sh) for sh in (ev.input_shapes or []) if sh])[:100]
    a = by_site[(site, ev.name)]
    a[0] += len(ks)
    a[1] += sum(k.duration for k in ks)
    a[2].update(k.name[:50] for k in ks)
tot = sum(a[1] for a in by_site.values())
print("non-b2 kernels: %d launches, %.1f us" % (sum(a[0] for a in by_site.values()), tot))
for (site, op), a in sorted(by_site.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%8.1f us  n=%3d  %-28s %s" % (a[1], a[0], op, site[-110:]))
