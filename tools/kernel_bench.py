"""Per-kernel timing on one GPU (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicvectorquantization_b200 import kernels as kn

dev = "cuda"
BF = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def timeit_graph(fn, launches=20, reps=5):
    """Kernel-only time of a launch-bound call: `launches` calls captured in one CUDA graph (no host work
    between them), best of `reps` replays, L2 flushed before each replay."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(launches):
            fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / launches)
    ts.sort()
    return ts[len(ts) // 2]


def vq(N=65536, C=256):
    for K in (256, 1024, 8192, 16384):
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(N, C, device=dev, generator=g)
        w = torch.cat([x[torch.randperm(N, device=dev)[:K]] + 0.1 * torch.randn(K, C, device=dev), torch.zeros(1, C, device=dev)])
        cb = kn.Codebook(K, C, dev); cb.refresh(w)
        xb = x.to(BF)
        ms = timeit(lambda: kn.vq_search_gather(xb, cb, w))
        msg = timeit_graph(lambda: kn.vq_search_gather(xb, cb, w))
        fl = 2.0 * N * K * C
        by = 2 * N * C + 2 * K * C + 8 * N + 2 * N * C
        print(json.dumps(dict(k="vq_search_gather", N=N, K=K, ms_eager=round(ms, 4), ms=round(msg, 4), tflops=round(fl / msg / 1e9, 1),
                              alg_GBs=round(by / msg / 1e6, 1))))
        counts = torch.zeros(K, device=dev); sums = torch.zeros(K, C, device=dev); loss = torch.zeros(1, device=dev)
        ms = timeit_graph(lambda: kn.vq_search_gather(xb, cb, w, counts=counts, sums=sums, loss_acc=loss))
        print(json.dumps(dict(k="vq_search_gather+ema", N=N, K=K, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))))


def conv(nb, h, w, cin, cout, k, stride):
    x = torch.randn(nb, h, w, cin, device=dev).to(BF)
    wt = torch.randn(cout, cin, k, k, device=dev) * (cin * k * k) ** -0.5
    b = torch.zeros(cout, device=dev)
    wp, wd = kn.pack_weight_fwd(wt), kn.pack_weight_dgrad(wt)
    ho, wo = h // stride, w // stride
    dy = torch.randn(nb, ho, wo, cout, device=dev).to(BF)
    fl = 2.0 * nb * ho * wo * cin * cout * k * k
    r = dict(k=f"conv{k}x{k}s{stride}", nb=nb, hw=h, cin=cin, cout=cout)
    def fwd_mt(mt):
        def f():
            kn.FORCE_MT = mt
            kn.conv_fwd(x, wp, b, k, stride, cout)
            kn.FORCE_MT = 0
        return f
    def fwd_nopconv():
        kn.USE_PCONV = False
        kn.conv_fwd(x, wp, b, k, stride, cout)
        kn.USE_PCONV = True
    for name, fn in (("fwd", lambda: kn.conv_fwd(x, wp, b, k, stride, cout)),
                     ("fwd_tap", fwd_nopconv),
                     ("dgrad", lambda: kn.conv_dgrad(dy, wd, k, stride, cin, (h, w))),
                     ("wgrad", lambda: kn.conv_wgrad(x, dy, k, stride))):
        ms = timeit(fn, iters=5, warm=2)
        r[name + "_ms"] = round(ms, 3); r[name + "_tflops"] = round(fl / ms / 1e9, 1)
    print(json.dumps(r))


def upconv(nb, h, w, cin, cout):
    """Folded nearest-x2 + 3x3 convolution (four 2x2 parity classes) forward / data gradient, low-res input h x w."""
    x = torch.randn(nb, h, w, cin, device=dev).to(BF)
    wt = torch.randn(cout, cin, 3, 3, device=dev) * (cin * 9) ** -0.5
    b = torch.zeros(cout, device=dev)
    wf, wd = kn.upconv_pack(wt)
    dy = torch.randn(nb, 2 * h, 2 * w, cout, device=dev).to(BF)
    fl = 2.0 * nb * (2 * h) * (2 * w) * cin * cout * 4
    r = dict(k="upconv", nb=nb, hw=h, cin=cin, cout=cout)
    for name, fn in (("fwd", lambda: kn.upconv_fwd(x, wf, b, cout)), ("dgrad", lambda: kn.upconv_dgrad(dy, wd, cin))):
        ms = timeit(fn, iters=5, warm=2)
        r[name + "_ms"] = round(ms, 3); r[name + "_tflops"] = round(fl / ms / 1e9, 1)
    print(json.dumps(r))


def gn(nb, h, w, c):
    x = torch.randn(nb, h, w, c, device=dev).to(BF)
    g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    dy = torch.randn(nb, h, w, c, device=dev).to(BF)
    by = x.numel() * 2
    ms1 = timeit(lambda: kn.gn_stats(x)); st = kn.gn_stats(x)
    ms2 = timeit(lambda: kn.gn_apply(x, st, g, b, True))
    ms3 = timeit(lambda: kn.gn_bwd(dy, x, st, g, b, True))
    old = kn.GN_L2_BUDGET_BYTES
    res = {}
    for mb in (70, 140, 280, 1 << 20):
        kn.GN_L2_BUDGET_BYTES = mb << 20
        res[f"bwd_{mb}MB"] = round(timeit(lambda: kn.gn_bwd(dy, x, st, g, b, True)), 3)
        res[f"fwd_{mb}MB"] = round(timeit(lambda: kn.gn_forward(x, g, b, True)), 3)
    kn.GN_L2_BUDGET_BYTES = old
    print(json.dumps(dict(k="gn_image_groups", nb=nb, hw=h, c=c, **res)))
    print(json.dumps(dict(k="gn", nb=nb, hw=h, c=c, stats_ms=round(ms1, 3), stats_GBs=round(by / ms1 / 1e6), apply_ms=round(ms2, 3),
                          apply_GBs=round(2 * by / ms2 / 1e6), bwd_ms=round(ms3, 3), bwd_GBs=round(5 * by / ms3 / 1e6))))


def vqk():
    """VQ search+gather at K=1024 (the codebook size of every stage-1 config): grid sweep and stream-K tail
    (B2DQ_VQ_SPLIT_MIN_TILES=4 lets the 4-codebook-tile rows be shared)."""
    C, K = 256, 1024
    for N in (65536, 32768):
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(N, C, device=dev, generator=g)
        w = torch.cat([x[torch.randperm(N, device=dev)[:K]] + 0.1 * torch.randn(K, C, device=dev), torch.zeros(1, C, device=dev)])
        cb = kn.Codebook(K, C, dev); cb.refresh(w)
        xb = x.to(BF)
        ref = kn.vq_search_gather(xb, cb, w, split=False)[0]
        for mc in (0, 128, 74):
            for split in (False, True):
                ms = timeit_graph(lambda: kn.vq_search_gather(xb, cb, w, max_ctas=mc, split=split))
                same = bool(torch.equal(kn.vq_search_gather(xb, cb, w, max_ctas=mc, split=split)[0], ref))
                print(json.dumps(dict(k="vq K=1024", N=N, max_ctas=mc, split=split, ms=round(ms, 4), same=same,
                                      tflops=round(2.0 * N * K * C / ms / 1e9, 1),
                                      env=os.environ.get("B2DQ_VQ_SPLIT_MIN_TILES", "8"))), flush=True)


def pconv():
    """Persistent strip convolution 128->128 at 256x256, batch 32: plain / +residual / +residual+stats / data gradient."""
    nb, h, w, c = 32, 256, 256, 128
    x = torch.randn(nb, h, w, c, device=dev).to(BF)
    res = torch.randn(nb, h, w, c, device=dev).to(BF)
    wt = torch.randn(c, c, 3, 3, device=dev) * (c * 9) ** -0.5
    b = torch.zeros(c, device=dev)
    wp, wd = kn.pack_weight_fwd(wt), kn.pack_weight_dgrad(wt)
    fl = 2.0 * nb * h * w * c * c * 9
    for name, fn in (("fwd", lambda: kn.pconv3x3(x, wp, b, None, False)),
                     ("fwd+res", lambda: kn.pconv3x3(x, wp, b, res, False)),
                     ("fwd+res+stats", lambda: kn.pconv3x3(x, wp, b, res, False, want_stats=True)),
                     ("fwd+stats", lambda: kn.pconv3x3(x, wp, b, None, False, want_stats=True)),
                     ("dgrad", lambda: kn.pconv3x3(x, wd, None, None, True))):
        for _ in range(3):
            fn()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s_.record()
        for _ in range(10):
            fn()
        e_.record(); torch.cuda.synchronize()
        ms = s_.elapsed_time(e_) / 10
        print(json.dumps(dict(k="pconv3x3 " + name, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))), flush=True)


def gnf():
    """GroupNorm backward: fused persistent kernel vs the separate-kernel path on every shape of the dual-config step
    (back-to-back launches, as inside a step).  B2DQ_GN_L2_BUDGET_MB selects the images in flight (teams)."""
    shapes = [(32, 256, 256, 128), (32, 128, 128, 128), (32, 128, 128, 256), (32, 64, 64, 128), (32, 64, 64, 256),
              (32, 32, 32, 256), (32, 16, 16, 256), (32, 16, 16, 512)]
    for nb, h, w, c in shapes:
        x = (torch.randn(nb, h, w, c, device=dev) * 1.5 + 0.3).to(BF)
        g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        dy = torch.randn(nb, h, w, c, device=dev).to(BF)
        add = torch.randn(nb, h, w, c, device=dev).to(BF)
        st = kn.gn_stats(x)
        by = 3 * x.numel() * 2
        r = dict(k="gn_bwd", nb=nb, hw=h, c=c, plan=kn.gn_bwd_fused_plan(nb, h * w, c),
                 budget_mb=os.environ.get("B2DQ_GN_L2_BUDGET_MB", "default"))
        for name, fused, a in (("fused", True, None), ("fused_add", True, add), ("split", False, None), ("split_add", False, add)):
            kn.USE_GN_FUSED = fused
            fn = lambda: kn.gn_bwd(dy, x, st, g, b, True, add=a)
            for _ in range(3):
                fn()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); s_.record()
            for _ in range(10):
                fn()
            e_.record(); torch.cuda.synchronize()
            ms = s_.elapsed_time(e_) / 10
            byt = by + (x.numel() * 2 if a is not None else 0)
            r[name + "_ms"] = round(ms, 4); r[name + "_GBs"] = round(byt / ms / 1e6)
        for name, fused in (("fwd_fused", True), ("fwd_split", False)):
            kn.USE_GN_FUSED = fused
            fn = lambda: kn.gn_forward(x, g, b, True)
            for _ in range(3):
                fn()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); s_.record()
            for _ in range(10):
                fn()
            e_.record(); torch.cuda.synchronize()
            ms = s_.elapsed_time(e_) / 10
            r[name + "_ms"] = round(ms, 4); r[name + "_GBs"] = round(2 * x.numel() * 2 / ms / 1e6)
        kn.USE_GN_FUSED = True
        dxa = kn.gn_bwd(dy, x, st, g, b, True)
        kn.USE_GN_FUSED = False
        dxb = kn.gn_bwd(dy, x, st, g, b, True)
        kn.USE_GN_FUSED = True
        r["max_abs_diff_dx"] = float((dxa[0].float() - dxb[0].float()).abs().max())
        r["dg_finite"] = bool(torch.isfinite(dxa[1]).all())
        print(json.dumps(r), flush=True)


def vq_small():
    """Small-N calls (residual quantizer depth step, stage-2 sampling): codebook split on / off."""
    C = 256
    for N, K in ((2048, 16384), (2048, 1024), (256, 16384), (32768, 1024)):
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(N, C, device=dev, generator=g)
        w = torch.cat([torch.randn(K, C, device=dev, generator=g), torch.zeros(1, C, device=dev)])
        cb = kn.Codebook(K, C, dev); cb.refresh(w)
        xb = x.to(BF)
        fl = 2.0 * N * K * C
        for split in (False, True):
            ms = timeit_graph(lambda: kn.vq_search_gather(xb, cb, w, split=split))
            print(json.dumps(dict(k="vq_search_gather", N=N, K=K, split=split, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))))


if __name__ == "__main__":
    what = sys.argv[1:] or ["vq", "conv", "gn"]
    if "vq" in what:
        vq()
        vq_small()
    if "conv" in what:
        for c in [(32, 256, 256, 128, 128, 3, 1), (32, 128, 128, 128, 128, 3, 1), (32, 64, 64, 256, 256, 3, 1), (32, 32, 32, 256, 256, 3, 1),
                  (32, 16, 16, 512, 512, 3, 1), (32, 32, 32, 256, 256, 1, 1), (32, 256, 256, 128, 128, 3, 2), (32, 128, 128, 256, 256, 3, 1)]:
            conv(*c)
    if "tap" in what:
        for c in [(32, 64, 64, 256, 256, 3, 1), (32, 32, 32, 256, 256, 3, 1), (32, 16, 16, 512, 512, 3, 1),
                  (32, 32, 32, 256, 768, 1, 1), (32, 32, 32, 256, 256, 1, 1), (32, 256, 256, 128, 128, 3, 2),
                  (32, 128, 128, 128, 256, 3, 1), (32, 256, 256, 64, 128, 1, 1)]:
            conv(*c)
        for c in [(32, 128, 128, 128, 128), (32, 64, 64, 256, 256), (32, 32, 32, 256, 256), (32, 16, 16, 512, 512)]:
            upconv(*c)
    if "vqk" in what:
        vqk()
    if "pconv" in what:
        pconv()
    if "gnf" in what:
        gnf()
    if "gn" in what:
        gn(32, 256, 256, 128); gn(32, 64, 64, 256); gn(32, 16, 16, 512)
