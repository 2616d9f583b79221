import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dynamicvectorquantization_b200 import configs, kernels as kn
from oracle import dqvae_oracle as orc

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).sum() / b.pow(2).sum().clamp_min(1e-30))

model = configs.build_model(configs.stage1_config("dqvae-dual-r-05"))
sd = orc.make_weights(orc.model_shapes(orc.DUAL_CFG), seed=17)
model.load_state_dict(sd, strict=False)
model = model.cuda().eval()
x = (torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(3)) * 2 - 1).cuda()
with torch.no_grad():
    outs = []
    for i in range(4):
        hd = model.encoder(x, None)
        quant, loss, info, gi, gate = model.encode(x)
        xrec = model.decode(quant)
        outs.append((hd["h_dual"].clone(), info[2].clone(), xrec.clone(), quant.clone()))
    for i in range(1, 4):
        print(f"run {i} vs 0: h_dual equal {torch.equal(outs[i][0], outs[0][0])} maxdiff {float((outs[i][0]-outs[0][0]).abs().max()):.3e} "
              f"codes differ {int((outs[i][1] != outs[0][1]).sum())}  xrec rel {rel(outs[i][2], outs[0][2]):.2e}")
    # which kernel family is non-deterministic?  repeat single ops
    h = torch.randn(1, 256, 256, 128, device="cuda").to(torch.bfloat16)
    w = torch.randn(128, 128, 3, 3, device="cuda") * 0.03
    wp = kn.pack_weight_fwd(w); b = torch.zeros(128, device="cuda")
    for name, fn in (("pconv", lambda: kn.conv_fwd(h, wp, b, 3, 1, 128)),
                     ("gn_stats", lambda: kn.gn_stats(h)),
                     ("gn_apply", lambda: kn.gn_apply(h, kn.gn_stats(h), torch.ones(128, device="cuda"), b, True))):
        r0 = fn(); same = all(torch.equal(fn(), r0) for _ in range(5))
        print(name, "deterministic:", same)
    kn.USE_PCONV = False
    r0 = kn.conv_fwd(h, wp, b, 3, 1, 128); print("tapgemm deterministic:", all(torch.equal(kn.conv_fwd(h, wp, b, 3, 1, 128), r0) for _ in range(5)))
    kn.USE_PCONV = True
    r1 = kn.conv_fwd(h, wp, b, 3, 1, 128); print("pconv == tapgemm:", torch.equal(r0, r1), float((r0.float()-r1.float()).abs().max()))
