import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dynamicvectorquantization_b200 import configs
from oracle import dqvae_oracle as orc

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).sum() / b.pow(2).sum().clamp_min(1e-30))

f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False); json.dump({"50": 1.5}, f); f.close()
for name, cfgfn, ocfg in (("entropy", lambda: configs.scaled_entropy_config(f.name), orc.SMALL_ENTROPY_CFG),
                          ("dual", lambda: configs.scaled_dual_config(), orc.SMALL_CFG)):
    model = configs.build_model(cfgfn())
    sd = orc.make_weights(orc.model_shapes(ocfg), seed=31)
    print(name, model.load_state_dict(sd, strict=False))
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    for flat in (False, True):
        xx = x.clone()
        if flat:
            xx[:, :, :32] = xx[:, :, :32].mean(dim=(2, 3), keepdim=True) + 0.02 * xx[:, :, :32]
        with torch.no_grad():
            ent = model.entropy_calculation(xx.cuda()) if name == "entropy" else None
            hd = model.encoder(xx.cuda(), ent)
            key = "h_dual"
            if name == "entropy":
                oenc = orc.dual_encoder(sd, ocfg, xx, x_entropy=orc.patch_entropy(xx, 16), entropy_threshold=1.5)
            else:
                oenc = orc.dual_encoder(sd, ocfg, xx, forced_gate=hd["gate"].cpu().permute(0, 2, 3, 1))
            print(name, "flat" if flat else "noise", "h_dual rel", rel(hd[key], oenc["h_dual"]),
                  "idx equal", bool((hd["indices"].cpu() == oenc["indices"]).all()),
                  "h_fine rel", None)
            out = model(xx.cuda())
            info = model.encode(xx.cuda())[2]
            oo = orc.model_forward(sd, ocfg, xx, entropy_threshold=1.5,
                                   forced_gate=None if name == "entropy" else hd["gate"].cpu().permute(0, 2, 3, 1),
                                   forced_codes=info[2].cpu())
            print("   xrec rel", rel(out[0], oo["xrec"]), "qloss", float(out[1]), float(oo["qloss"]))
