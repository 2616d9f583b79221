import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dynamicvectorquantization_b200 import configs
from oracle import dqvae_oracle as orc

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).sum() / b.pow(2).sum().clamp_min(1e-30))

for name, cfgname, ocfg in (("dual", "dqvae-dual-r-05", orc.DUAL_CFG), ("triple", "dqvae-triple-r-03-03", orc.TRIPLE_CFG)):
    for seed in (7, 17):
        model = configs.build_model(configs.stage1_config(cfgname))
        sd = orc.make_weights(orc.model_shapes(ocfg), seed=seed)
        model.load_state_dict(sd, strict=False)
        model = model.cuda().eval()
        x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(3)) * 2 - 1
        with torch.no_grad():
            xrec, qloss, indices, gate = model(x.cuda())
            info = model.encode(x.cuda())[2]
            out = orc.model_forward(sd, ocfg, x, forced_gate=gate.cpu().permute(0, 2, 3, 1), forced_codes=info[2].cpu())
            # decoder-only comparison: feed the oracle's post_quant input through the product decoder
            e_dec = rel(xrec, out["xrec"])
            # encoder-only: latent before VQ
            hd = model.encoder(x.cuda(), None)
            key = "h_triple" if name == "triple" else "h_dual"
            e_enc = rel(hd[key], out["h_dual"])
        print(f"{name} seed {seed}: xrec rel-MSE {e_dec:.2e}  |xrec| rms {float(out['xrec'].pow(2).mean().sqrt()):.3f} "
              f"h rel-MSE {e_enc:.2e} fine-frac {float((indices>0).float().mean()):.2f} qloss {float(qloss):.3f}/{float(out['qloss']):.3f}")
        del model
        torch.cuda.empty_cache()
