"""Launch the three headline kernels a few times (for `ncu --set full -k regex:...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicvectorquantization_b200 import kernels as kn
dev, BF = "cuda", torch.bfloat16
nb, hw, c = 32, 256, 128
x = torch.randn(nb, hw, hw, c, device=dev).to(BF)
w = torch.randn(c, c, 3, 3, device=dev) * (c * 9) ** -0.5
wp = kn.pack_weight_fwd(w)
bias = torch.zeros(c, device=dev)
dy = torch.randn(nb, hw, hw, c, device=dev).to(BF)
N, C, K = 65536, 256, 1024
xv = torch.randn(N, C, device=dev)
wv = torch.cat([xv[torch.randperm(N, device=dev)[:K]] + 0.1 * torch.randn(K, C, device=dev), torch.zeros(1, C, device=dev)])
cb = kn.Codebook(K, C, dev); cb.refresh(wv)
xb = xv.to(BF)
which = sys.argv[1:] or ["conv", "wgrad", "vq", "gn"]
gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)
res = torch.randn(nb, hw, hw, c, device=dev).to(BF)
for _ in range(3):
    if "conv" in which:
        kn.conv_fwd(x, wp, bias, 3, 1, c)                       # pconv3x3_kernel
        kn.USE_PCONV = False; kn.conv_fwd(x, wp, bias, 3, 1, c); kn.USE_PCONV = True   # tapgemm_kernel
    if "wgrad" in which:
        kn.conv_wgrad(x, dy, 3, 1)
    if "vq" in which:
        kn.vq_search_gather(xb, cb, wv)
    if "gn" in which:
        y, st = kn.gn_forward(x, gam, bet, True)                # gn_fwd_fused_kernel
        kn.gn_bwd(dy, x, st, gam, bet, True, add=res)           # gn_bwd_fused_kernel
torch.cuda.synchronize()
print("done")
