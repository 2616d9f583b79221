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
which = sys.argv[1:] or ["conv", "wgrad", "vq", "gn", "conv256", "upconv"]
gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)
res = torch.randn(nb, hw, hw, c, device=dev).to(BF)
# 3x3 256 -> 256 at 64x64 (tapgemm_kernel<256,3,2>) and the folded up-convolution 128 -> 128, 128^2 -> 256^2 (four
# parity-class launches of tapgemm_kernel<128,2,2>)
x256 = torch.randn(nb, 64, 64, 256, device=dev).to(BF)
w256 = kn.pack_weight_fwd(torch.randn(256, 256, 3, 3, device=dev) * (256 * 9) ** -0.5)
b256 = torch.zeros(256, device=dev)
xlo = torch.randn(nb, 128, 128, c, device=dev).to(BF)
wup, _ = kn.upconv_pack(w)
for _ in range(3):
    if "conv" in which:
        kn.conv_fwd(x, wp, bias, 3, 1, c)                       # pconv3x3_kernel
        kn.USE_PCONV = False; kn.conv_fwd(x, wp, bias, 3, 1, c); kn.USE_PCONV = True   # tapgemm_kernel
    if "wgrad" in which:
        kn.conv_wgrad(x, dy, 3, 1)
    if "vq" in which:
        kn.vq_search_gather(xb, cb, wv)
    if "conv256" in which:
        kn.conv_fwd(x256, w256, b256, 3, 1, 256)
    if "upconv" in which:
        kn.upconv_fwd(xlo, wup, bias, c)
    if "gn" in which:
        y, st = kn.gn_forward(x, gam, bet, True)                # gn_fwd_fused_kernel
        kn.gn_bwd(dy, x, st, gam, bet, True)                    # gn_bwd_fused_kernel (as bench.py's roofline_gn: no residual add)
torch.cuda.synchronize()
print("done")
