"""Layer-by-layer check of the PatchGAN kernels against torch ops on the GPU (fp32), to localise a mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dynamicvectorquantization_b200 import kernels as kn, ops
from dynamicvectorquantization_b200.nn import discriminator as D

BF = torch.bfloat16
torch.manual_seed(0)


def rr(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(((a - b).pow(2).mean() / b.pow(2).mean().clamp_min(1e-30)).sqrt())


def nchw(t):
    return t.float().permute(0, 3, 1, 2)


for res, nb in ((64, 2), (256, 2)):
    x = (torch.rand(nb, res, res, 3, device="cuda") * 2 - 1).to(BF)
    w0 = (torch.randn(64, 3, 4, 4, device="cuda") * 0.2).requires_grad_(True)
    b0 = (torch.randn(64, device="cuda") * 0.1).requires_grad_(True)
    xg = x.clone().requires_grad_(True)
    y = D._StemFn.apply(xg, w0, b0)
    xr = nchw(x).requires_grad_(True)
    w0r, b0r = w0.detach().to(BF).float().requires_grad_(True), b0.detach().clone().requires_grad_(True)
    ref = F.leaky_relu(F.conv2d(xr, w0r, b0r, stride=2, padding=1), 0.2)
    print(res, "stem fwd", rr(nchw(y), ref))
    cot = torch.randn_like(ref)
    gy = torch.autograd.grad(y, [xg, w0, b0], cot.permute(0, 2, 3, 1).to(BF).contiguous())
    gr = torch.autograd.grad(ref, [xr, w0r, b0r], cot.to(BF).float())
    print(res, "stem dx", rr(nchw(gy[0]), gr[0]), "dw", rr(gy[1], gr[1]), "db", rr(gy[2], gr[2]))
    h = y.detach()
    for cin, cout, stride in ((64, 128, 2), (128, 256, 2), (256, 512, 1)):
        w = (torch.randn(cout, cin, 4, 4, device="cuda") * (cin * 16) ** -0.5).requires_grad_(True)
        hg = h.clone().requires_grad_(True)
        z = D._Conv4x4Fn.apply(hg, w, None, stride)
        hr = nchw(h).requires_grad_(True)
        wr = w.detach().to(BF).float().requires_grad_(True)
        zr = F.conv2d(hr, wr, None, stride=stride, padding=1)
        print(res, f"conv {cin}->{cout} s{stride} fwd", rr(nchw(z), zr), tuple(z.shape))
        cot = torch.randn_like(zr)
        gz = torch.autograd.grad(z, [hg, w], cot.permute(0, 2, 3, 1).to(BF).contiguous())
        gzr = torch.autograd.grad(zr, [hr, wr], cot.to(BF).float())
        print(res, "   dx", rr(nchw(gz[0]), gzr[0]), "dw", rr(gz[1], gzr[1]))
        bn = torch.nn.BatchNorm2d(cout).cuda().train()
        with torch.no_grad():
            bn.weight.normal_(1, 0.2); bn.bias.normal_(0, 0.2)
        bn2 = torch.nn.BatchNorm2d(cout).cuda().train()
        bn2.load_state_dict(bn.state_dict())
        zd = z.detach().clone().requires_grad_(True)
        a = D._batchnorm_lrelu(zd, bn)
        zrr = nchw(z.detach()).requires_grad_(True)
        ar = F.leaky_relu(bn2(zrr), 0.2)
        print(res, "   bn+lrelu fwd", rr(nchw(a), ar), "running", rr(bn.running_mean, bn2.running_mean), rr(bn.running_var, bn2.running_var))
        cot = torch.randn_like(ar)
        ga = torch.autograd.grad(a, [zd, bn.weight, bn.bias], cot.permute(0, 2, 3, 1).to(BF).contiguous())
        gar = torch.autograd.grad(ar, [zrr, bn2.weight, bn2.bias], cot.to(BF).float())
        print(res, "   bn bwd dx", rr(nchw(ga[0]), gar[0]), "dg", rr(ga[1], gar[1]), "db", rr(ga[2], gar[2]))
        h = a.detach()
    wh = (torch.randn(1, 512, 4, 4, device="cuda") * (512 * 16) ** -0.5).requires_grad_(True)
    bh = torch.randn(1, device="cuda").requires_grad_(True)
    hg = h.clone().requires_grad_(True)
    o = D._HeadFn.apply(hg, wh, bh)
    hr = nchw(h).requires_grad_(True)
    whr, bhr = wh.detach().to(BF).float().requires_grad_(True), bh.detach().clone().requires_grad_(True)
    orf = F.conv2d(hr, whr, bhr, stride=1, padding=1)
    print(res, "head fwd", rr(nchw(o), orf), tuple(o.shape))
    for kind in ("random", "mean"):
        cot = torch.randn_like(orf) if kind == "random" else torch.full_like(orf, -1.0 / orf.numel())
        go = torch.autograd.grad(o, [hg, wh, bh], cot.permute(0, 2, 3, 1).contiguous(), retain_graph=True)
        gor = torch.autograd.grad(orf, [hr, whr, bhr], cot, retain_graph=True)
        print(res, f"head bwd ({kind}) dx", rr(nchw(go[0]), gor[0]), "dw", rr(go[1], gor[1]), "db", rr(go[2], gor[2]),
              "norm ratio dw", float(go[1].norm() / gor[1].norm()))
