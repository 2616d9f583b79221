"""Per-tap correctness of the persistent strip conv under both descriptor modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from dynamicvectorquantization_b200 import kernels as kn
BF = torch.bfloat16
nb, h, w, c = 2, 8, 256, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(nb, h, w, c, generator=g).to(BF)
for mode in ("0", "1"):
    os.environ["B2DQ_PCONV_MODE"] = mode
    for r in range(3):
        for s in range(3):
            wt = torch.zeros(c, c, 3, 3)
            wt[:, :, r, s] = torch.randn(c, c, generator=g) * c ** -0.5
            wt = wt.to(BF).float()
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, padding=1).permute(0, 2, 3, 1)
            y = kn.pconv3x3(x.cuda(), kn.pack_weight_fwd(wt.cuda()), None, None, False).float().cpu()
            err = ((y - ref).pow(2).sum() / ref.pow(2).sum()).sqrt().item()
            # error split by output column position within the 8-pixel swizzle period
            per = [((y[:, :, i::8] - ref[:, :, i::8]).pow(2).sum() / ref[:, :, i::8].pow(2).sum()).sqrt().item() for i in range(8)]
            print(f"mode {mode} tap r={r} s={s}: rel {err:.3f}  by (w%8): " + " ".join(f"{e:.2f}" for e in per))
