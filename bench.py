#!/usr/bin/env python
"""Headline benchmark: images/sec of the DQ-VAE stage-1 forward+backward (dqvae-dual-r-05,
256x256, bf16 compute, 32 images per GPU) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle port) on the host cores

A step = encoder -> quant_conv -> VQ (search + EMA update) -> post_quant_conv -> decoder forward,
surrogate loss |xrec - x|.mean() + qloss + budget(gate) (the reference loss needs downloaded VGG16
weights: SURVEY.md 8c), backward, gradient all-reduce (DDP, N > 1) and the Adam step of the
autoencoder optimizer (dqvae_dual_feat.py:144-149).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (256x256 DQ-VAE fwd+bwd)"
WORKLOADS = {
    "dqvae-dual-r-05": "dqvae-dual-r-05 (F=16/F=8, K=1024) 256x256 bf16, batch 32 per GPU",
    "dqvae-entropy-dual-r05": "dqvae-entropy-dual-r05 (fixed entropy router) 256x256 bf16, batch 32 per GPU",
    "dqvae-triple-r-03-03": "dqvae-triple-r-03-03 (F=32/16/8, K=1024) 256x256 bf16, batch 32 per GPU",
}
WORKLOAD = WORKLOADS["dqvae-dual-r-05"]
FLOP_PER_IMAGE_FWD = 392.8e9 + 0.27e9          # BASELINE.md section 2 (2*MAC, attention included)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="--impl reference: wall-clock bound of the run")
    ap.add_argument("--cpu-sample-batch", type=int, default=2, help="--impl reference: images per CPU step")
    ap.add_argument("--config", default="dqvae-dual-r-05", choices=sorted(WORKLOADS),
                    help="headline = dqvae-dual-r-05; the others are parity-test configs that can be timed too")
    ap.add_argument("--no-real-loss", action="store_true",
                    help="N=1: do not append the `real_loss_step` object (a short `--loss real` run in a child process)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: exchange the gradient buckets only after the backward (A/B of the overlap)")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph (N=1)")
    ap.add_argument("--loss", default="surrogate", choices=["surrogate", "real"],
                    help="real: time the reference's full training_step (both optimizer passes, LPIPS + PatchGAN + "
                         "adaptive weight; SURVEY 8f row 1) on one GPU instead of the headline fwd+bwd step")
    ap.add_argument("--aux", action="store_true",
                    help="instead of the headline step, time the widened rows (SURVEY 8f: patch entropy, stage-2 "
                         "permuter, residual quantizer) on one GPU, one JSON line each with roofline + cpu_baseline")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference / CPU arm
def usable_cores():
    """CPU cores this process may really use: affinity mask and cgroup quota, not the host total."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def pick_threads(torch):
    """All usable host threads is not the fastest setting for batch-1 convolutions on a many-core
    host (oneDNN oversubscribes): time a proxy layer (3x3 128->128 at 256x256, fwd+bwd) at a few
    thread counts and keep the best."""
    import torch.nn.functional as F
    cores = usable_cores()
    cands = sorted({c for c in (cores, cores // 2, cores // 4, 64, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    x = torch.randn(1, 128, 256, 256)
    w = torch.randn(128, 128, 3, 3, requires_grad=True)
    best = (None, 1e30)
    for t in cands:
        torch.set_num_threads(t)
        F.conv2d(x, w, padding=1).sum().backward()          # warm the pool at this size
        t0 = time.perf_counter()
        for _ in range(2):
            F.conv2d(x, w, padding=1).sum().backward()
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (t, dt)
    torch.set_num_threads(best[0])
    return best[0], cores


REF_YAML = {"dqvae-dual-r-05": "dqvae-dual-r-05_imagenet.yml",
            "dqvae-entropy-dual-r05": "dqvae-entropy-dual-r05_imagenet.yml",
            "dqvae-triple-r-03-03": "dqvae-triple-r-03-03_imagenet.yml"}


def find_reference_tree():
    """The reference's own source tree, if one travelled with the repo: baseline/_ref (git-ignored copy made by
    __graft_entry__.build() in the build container, or dropped there by the driver), else /root/reference."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "modules", "dynamic_modules", "EncoderDual.py")):
            return cand
    return None


def reference_step_fn(cfg_name, batch, ref_root):
    """One fwd+bwd of the REFERENCE's own modules (encoder, quantizer, 1x1 convs, decoder built from the
    reference's YAML by its own instantiate_from_config, fp32, torch CPU) under the surrogate loss of the B200 arm.
    Only pytorch_lightning is shimmed (LightningModule = nn.Module; it is not installed here)."""
    import types
    import torch
    import torch.nn as nn
    import yaml
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = nn.Module
        sys.modules["pytorch_lightning"] = pl
    sys.path.insert(0, ref_root)                  # before the repo root: `modules`, `models`, `utils` = reference's
    cwd = os.getcwd()
    os.chdir(ref_root)                            # the reference resolves relative paths (threshold JSON) from its root
    try:
        from utils.utils import instantiate_from_config
        conf = yaml.safe_load(open(os.path.join(ref_root, "configs", "stage1", REF_YAML[cfg_name])))["model"]["params"]
        torch.manual_seed(2021)
        m = nn.Module()
        m.encoder = instantiate_from_config(conf["encoderconfig"])
        m.decoder = instantiate_from_config(conf["decoderconfig"])
        m.quantize = instantiate_from_config(conf["vqconfig"])
        m.quant_conv = nn.Conv2d(conf["quant_before_dim"], conf["quant_after_dim"], 1)
        m.post_quant_conv = nn.Conv2d(conf["quant_after_dim"], conf["quant_before_dim"], 1)
        budget = instantiate_from_config(conf["lossconfig"]["params"]["budget_loss_config"]) \
            if "budget_loss_config" in conf["lossconfig"]["params"] else None
        entropy = None
        if cfg_name == "dqvae-entropy-dual-r05":
            from models.stage1_dynamic.dqvae_dual_entropy import Entropy
            entropy = Entropy(conf.get("entropy_patch_size", 16), conf.get("image_size", 256), conf.get("image_size", 256))
    finally:
        os.chdir(cwd)
    for mod in (m.encoder, m.decoder, m.quantize):
        assert type(mod).__module__.split(".")[0] == "modules" and ref_root in sys.modules[type(mod).__module__].__file__
    m.train()
    hkey = "h_triple" if "triple" in cfg_name else "h_dual"
    g = torch.Generator().manual_seed(2021)
    x = torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1
    params = [p for p in m.parameters() if p.requires_grad]

    def step():
        for p in params:
            p.grad = None
        hd = m.encoder(x, entropy(x) if entropy is not None else None)
        h = m.quant_conv(hd[hkey])
        quant, qloss, _ = m.quantize(x=h, temp=0.0, codebook_mask=hd["codebook_mask"])
        xrec = m.decoder(m.post_quant_conv(quant), hd["indices"])
        loss = (xrec - x).abs().mean() + qloss
        if budget is not None:
            loss = loss + budget(hd["gate"])
        loss.backward()
        return float(loss.detach())
    return step


def oracle_step_fn(cfg_name, batch, seed=0):
    """One fwd+bwd of the fp32 oracle port (reference algorithm) on the host cores."""
    import torch
    from oracle import dqvae_oracle as orc
    ocfg = orc.DUAL_CFG
    sd = orc.make_weights(orc.model_shapes(ocfg), seed=seed)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "ema" not in k and "codebook" not in k}
    full = dict(sd); full.update(params)
    g = torch.Generator().manual_seed(2021)
    x = torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1

    def step():
        for p in params.values():
            p.grad = None
        out = orc.model_forward(full, ocfg, x)
        loss = (out["xrec"] - x).abs().mean() + out["qloss"] + orc.budget_loss_dual(out["gate"].float())
        loss.backward()
        return float(loss.detach())
    return step


def cpu_step_fn(cfg_name, batch):
    """-> (step, kind, what): the reference's own modules when its tree is available, else the oracle port."""
    import contextlib
    ref_root = find_reference_tree()
    if ref_root is not None:
        try:
            with contextlib.redirect_stdout(sys.stderr):      # the reference's constructors print banners
                fn = reference_step_fn(cfg_name, batch, ref_root)
            return (fn, "reference",
                    f"the reference's own modules imported from {os.path.relpath(ref_root, ROOT) if ref_root.startswith(ROOT) else ref_root} "
                    f"(built from configs/stage1/{REF_YAML[cfg_name]}; training mode: gumbel routing + EMA codebook update)")
        except Exception as e:
            sys.stderr.write(f"[bench] reference modules unusable ({type(e).__name__}: {e}); timing the oracle port\n")
    return (oracle_step_fn(cfg_name, batch), "port",
            "fp32 oracle port of the reference modules (no reference tree next to the repo)")


def e2e_loop(step, x_host, dev, steps, sink):
    """The end-to-end leg: every step copies its input from pinned host memory and reads its result back.  The copy of
    step i+1 is issued on a copy stream while step i computes (two device buffers, the usual prefetching input
    pipeline of a training loop: ``pin_memory`` + ``non_blocking``); only the first copy is exposed.  `sink(out)`
    issues the device->host read of the step's result."""
    import torch
    cur_stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty(x_host.shape, dtype=x_host.dtype, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(consumed[b])
            bufs[b].copy_(x_host, non_blocking=True)
            ready[b].record(copy_stream)

    copy_stream.wait_stream(cur_stream)
    prefetch(0)
    for i in range(steps):
        if i + 1 < steps:
            prefetch(i + 1)
        b = i & 1
        cur_stream.wait_event(ready[b])
        out = step(bufs[b])
        consumed[b].record(cur_stream)
        sink(out)
    cur_stream.wait_stream(copy_stream)


def workload_config(args, world):
    """The `config` object - identical in both arms (the reference arm times a bounded per-step sample of it)."""
    return {"workload": WORKLOADS[args.config], "global_batch": world * args.batch, "parallelism": f"dp{world}",
            "step": "fwd + bwd (surrogate L1 + qloss + budget loss) + DDP all-reduce + Adam",
            "l2": "per-step working set (~40 GB of activations) >> 126 MB L2, no flush needed"}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads, cores = pick_threads(torch)
    sample_b = max(1, args.cpu_sample_batch)
    step, kind, what = cpu_step_fn(args.config, sample_b)
    t0 = time.perf_counter(); step(); first = time.perf_counter() - t0
    budget_s = args.cpu_budget_s
    warm = max(0, min(args.warmup - 1, int(budget_s * 0.2 / max(first, 1e-3))))
    for _ in range(warm):
        step()
    steps = max(1, min(args.steps, int(budget_s * 0.8 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    ips = sample_b * steps / dt
    sample = (f"{steps} x fwd+bwd of {sample_b} images per step (of the {world * args.batch}-image batch of the config: "
              f"a CPU step of the full batch needs ~25 fp32 activations of 2 GiB each per 32 images and minutes per "
              f"step; the per-image cost of the convolutions does not depend on the batch) - {what}; {threads} "
              f"threads = fastest of the thread counts tried on the {cores} usable cores")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm + 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from dynamicvectorquantization_b200 import configs, kernels as kn
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")    # NCCL work is captured in a CUDA graph
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(2021)                      # the reference's default seed (train.py:41)
    global WORKLOAD
    WORKLOAD = WORKLOADS[args.config]
    cfg = configs.stage1_config(args.config)
    if args.config == "dqvae-entropy-dual-r05":
        # the router reads its threshold from a JSON of the reference tree (scripts/tools/thresholds/...);
        # off the reference tree, write the one value it uses (key "50" = 1.6778)
        jp = cfg["params"]["encoderconfig"]["params"]["router_config"]["params"]["json_path"]
        if not os.path.exists(jp):
            import tempfile
            f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
            json.dump({"50": 1.6778}, f); f.close()
            cfg["params"]["encoderconfig"]["params"]["router_config"]["params"]["json_path"] = f.name
    model = configs.build_model(cfg).to(dev)
    model.train()
    for p in model.loss.parameters():
        p.requires_grad_(False)
    model.learning_rate = 4.5e-6 * world * args.batch
    ae_params = [p for n, p in model.named_parameters() if not n.startswith("loss.") and p.requires_grad]
    use_graph = not args.no_graph
    graph_ddp = use_graph and world > 1          # N > 1: graph-captured fwd+bwd, one flat gradient all-reduce
    # the reference's optimizer (dqvae_dual_feat.py:144-149), in torch's single-pass fused CUDA form
    try:
        opt = torch.optim.Adam(ae_params, lr=model.learning_rate, betas=(0.5, 0.9), capturable=use_graph, fused=True)
    except (TypeError, RuntimeError, ValueError):
        opt = torch.optim.Adam(ae_params, lr=model.learning_rate, betas=(0.5, 0.9), capturable=use_graph)
    net = model
    ex = None                                    # bucketed gradient exchange overlapped with the backward (N > 1, graph)
    ex_in_graph = False
    if graph_ddp:
        from dynamicvectorquantization_b200 import ops as b2ops
        from dynamicvectorquantization_b200.ddp import BucketedGradExchange
        # every rank starts from rank 0's parameters / buffers (what DDP's constructor does)
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
        b2ops.invalidate_caches(model)          # writes through .data do not bump the version the caches key on
        ex = BucketedGradExchange(ae_params, bucket_mb=float(os.environ.get("B2DQ_BUCKET_MB", "25")),
                                  overlap=not args.no_overlap)
    elif world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    B = args.batch
    g = torch.Generator().manual_seed(2021 + rank)
    x_host = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).pin_memory()
    x_dev = x_host.to(dev)
    loss_host = torch.zeros(1).pin_memory()

    def fwd_bwd(x):
        if ex is not None:
            ex.begin_step()                      # gradients dropped: the backward assigns, the exchange gathers them
        else:
            opt.zero_grad(set_to_none=True)
        xrec, qloss, indices, gate = net(x)[:4]
        loss, _ = model.loss(qloss, x, xrec, 0, 0, last_layer=None, split="train", gate=gate)
        loss.backward()                          # buckets of the flat buffer are all-reduced as they complete
        if ex is not None and ex_in_graph:
            ex.finish()
        return loss

    exchange_note = None
    if graph_ddp:
        exchange_note = ("the whole step (fwd, bwd, exchanges, Adam) is ONE CUDA graph: the packed VQ-statistics all-reduce "
                         "+ restart-row broadcast (1 MiB) sit inside the forward, the gradient is exchanged in %d buckets "
                         "of <= %s MB (reverse parameter order, NCCL AVG over NVLink) on a side stream as soon as a "
                         "bucket's last gradient is written, %s" % (
                             len(ex.buckets), os.environ.get("B2DQ_BUCKET_MB", "25"),
                             "overlapped with the rest of the backward" if ex.overlap else "NOT overlapped (--no-overlap)"))
    elif world > 1:
        exchange_note = "torch DistributedDataParallel (bucketed all-reduce overlapped with the backward), eager launches"

    def finish(loss):
        if ex is not None:
            if not ex_in_graph:
                ex.finish()
            if model.quantize.codebook.defer_ema:
                model.quantize.codebook.apply_deferred_ema(keep=True)   # packed all-reduce + rank-0 restart rows
        opt.step()
        return loss

    def step(x):
        return finish(fwd_bwd(x))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):         # also runs the 3 EMA warm-up passes of SURVEY 8d
        step(x_dev)
    barrier()
    # Single-GPU: capture the whole step (forward, backward, EMA update, Adam) in ONE CUDA graph so
    # the ~1900 launches of a step are replayed without host involvement.
    graph = None
    launches_per_step = None
    if use_graph:
        try:
            static_x = x_dev.clone()
            l0 = kn.launch_count()
            graph = torch.cuda.CUDAGraph()
            whole_step_graph = True
            if graph_ddp:
                try:                              # NCCL collectives (VQ statistics, gradient buckets) inside the capture
                    ex_in_graph = True
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        static_loss = step(static_x)
                except Exception as e:
                    sys.stderr.write(f"[bench] whole-step capture with the NCCL exchanges failed ({type(e).__name__}: {e}); "
                                     f"capturing fwd+bwd only, exchanging after the graph\n")
                    ex_in_graph = False
                    whole_step_graph = False
                    exchange_note += " [whole-step capture failed: gradient / statistics exchanged eagerly after the graph]"
                    model.quantize.codebook.defer_ema = True
                    torch.cuda.synchronize()
                    for _ in range(2):
                        step(static_x)
                    torch.cuda.synchronize()
                    l0 = kn.launch_count()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        static_loss = fwd_bwd(static_x)
            else:
                with torch.cuda.graph(graph):
                    static_loss = step(static_x)
            launches_per_step = kn.launch_count() - l0
            if not whole_step_graph:              # EMA finalize kernels run eagerly after each replay
                l1 = kn.launch_count()
                finish(static_loss)
                launches_per_step += kn.launch_count() - l1

            def step(x):                          # noqa: F811  (graph replay with the eager signature)
                if x is not static_x:
                    static_x.copy_(x, non_blocking=True)
                graph.replay()
                return static_loss if whole_step_graph else finish(static_loss)
            for _ in range(2):
                step(static_x)
            x_dev = static_x
        except Exception as e:                    # capture is an optimisation, not a requirement
            sys.stderr.write(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly\n")
            graph = None
            torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = kn.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step(x_dev)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = kn.launch_count() - launches0
    if graph is not None:
        launches = launches_per_step * args.steps   # replayed from the graph, counted at capture
    # end-to-end: host buffers in, loss out, every step
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(step, x_host, dev, args.steps, lambda l: loss_host.copy_(l.detach().reshape(1), non_blocking=True))
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    def leave():
        """N > 1: the captured graph holds NCCL work; tearing the communicator down under it can block for minutes.
        Everything is measured and printed by now: synchronise, meet the other ranks once more, and leave the
        process without the destructors."""
        if world > 1:
            sys.stdout.flush(); sys.stderr.flush()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        leave()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ips = world * B * args.steps / (ms / 1e3)
    ips_e2e = world * B * args.steps / (ms_e2e / 1e3)
    roof, roof_vq, roof_gn = kernel_rooflines(torch, kn, dev, peaks, with_cpu=(world == 1 and not args.no_cpu_baseline))
    line = {"metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world), "cuda_graph": graph is not None,
            "model_tflops_per_step": 3 * FLOP_PER_IMAGE_FWD * B / 1e12,
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4 * world,
                    "d2h_bytes_per_step": 4 * world,
                    "input_pipeline": "pinned host batch -> device on a copy stream, double-buffered: the copy of step "
                                      "i+1 runs under step i (all copies inside the timed region)"},
            "model_flops_utilisation": {"achieved_tflops": 3 * FLOP_PER_IMAGE_FWD * ips / world / 1e12,
                                        "peak_tflops": peaks.get("bf16_tflops_sustained"),
                                        "note": "algorithmic conv+attention FLOPs (BASELINE.md) x3 for fwd+bwd, per GPU"},
            "roofline": roof, "roofline_vq": roof_vq, "roofline_gn": roof_gn}
    if world > 1:
        line["exchange"] = exchange_note
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    if world == 1 and not args.no_real_loss:
        line["real_loss_step"] = real_loss_line(args)
    print(json.dumps(line), flush=True)
    leave()


def kernel_rooflines(torch, kn, dev, peaks, with_cpu=False):
    """Live CUDA-event timing (current stream) of the dominant kernel of the step - the 3x3
    128->128 convolution at 256x256, batch 32 (77 + 135 GF/img of the 393 GF/img forward are this
    shape) - and of the VQ search kernel at the microbench shape."""
    BF = torch.bfloat16
    nb, hw, c = 32, 256, 128
    x = torch.randn(nb, hw, hw, c, device=dev).to(BF)
    w = torch.randn(c, c, 3, 3, device=dev) * (c * 9) ** -0.5
    wp = kn.pack_weight_fwd(w)
    bias = torch.zeros(c, device=dev)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    ms = timed(lambda: kn.conv_fwd(x, wp, bias, 3, 1, c), 20)
    flops = 2.0 * nb * hw * hw * c * c * 9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:
        pass
    from dynamicvectorquantization_b200.build import source_sha

    def traffic(key):
        """DRAM bytes per launch from the committed ncu capture - only if it was taken from THESE kernel sources."""
        if prof.get(key + "_csrc_sha16") == source_sha(key):
            return prof.get(key + "_dram_bytes_per_launch")
        return None

    burst, sustained = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    roof = {"kernel": "pconv3x3_kernel (conv3x3 128->128 @256x256, batch 32, forward)", "bound": "tensor",
            "achieved": flops / ms / 1e9, "peak": burst, "unit": "TFLOP/s",
            "frac": (flops / ms / 1e9 / burst) if burst else None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone, 3 warm-up + 20 launches)"
            if burst else "unavailable",
            "ms_per_launch": ms, "traffic": traffic("pconv"),
            "traffic_source": "profiles/ncu_summary.json (ncu --set full of the same kernel sources; null if the sources changed)",
            "algorithmic_bytes": 2 * 2 * nb * hw * hw * c + 2 * 9 * c * c,
            "sustained_peak": sustained,
            "frac_of_sustained_peak": (flops / ms / 1e9 / sustained) if sustained else None}
    peak_t = burst
    del x, w, wp
    # GroupNorm(+swish) backward on the largest activation of the step ([32,256,256,128] bf16: 15 of the 69
    # GroupNorm layers, most of the family's bytes).  Algorithmic traffic = read dy, read x, write dx.
    peak_h = peaks.get("hbm_gbs")
    xg = torch.randn(nb, hw, hw, c, device=dev).to(BF)
    dyg = torch.randn(nb, hw, hw, c, device=dev).to(BF)
    gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    _, st = kn.gn_forward(xg, gam, bet, True)
    ms_gn = timed(lambda: kn.gn_bwd(dyg, xg, st, gam, bet, True), 20)
    by_gn = 3 * xg.numel() * 2
    roof_gn = {"kernel": "GroupNorm(32)+swish backward on [32,256,256,128] bf16 (" + kn.gn_bwd_kernel_name() + ")",
               "bound": "hbm", "achieved": by_gn / ms_gn / 1e6, "peak": peak_h, "unit": "GB/s",
               "frac": (by_gn / ms_gn / 1e6 / peak_h) if peak_h else None, "ms_per_launch": ms_gn,
               "algorithmic_bytes": by_gn, "traffic": traffic("gn"),
               "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peak_h else "unavailable",
               "timing": "3 warm-up + 20 back-to-back launches, CUDA events; 1.6 GB per launch >> 126 MB L2"}
    del xg, dyg
    # VQ search (+gather) at N=65536, C=256: algorithmic bytes = 2NC + 2KC + 8N + 2NC.  The kernel (tens of
    # microseconds) is shorter than the host side of one launch, so 20 launches are captured in a CUDA
    # graph and the replay is timed (CUDA events, L2 flushed before each replay).
    N, C = 65536, 256
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def graph_ms(fn, launches=20, reps=7):
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
        return sorted(ts)[len(ts) // 2]

    sweep = []
    for K in (256, 1024, 8192, 16384):
        gen = torch.Generator(device=dev).manual_seed(0)
        xv = torch.randn(N, C, device=dev, generator=gen)
        wv = torch.cat([xv[torch.randperm(N, device=dev, generator=gen)[:K]] +
                        0.1 * torch.randn(K, C, device=dev, generator=gen), torch.zeros(1, C, device=dev)])
        cb = kn.Codebook(K, C, dev); cb.refresh(wv)
        xb = xv.to(BF)
        if K == 1024:
            wv_k1024 = wv
        msv = graph_ms(lambda: kn.vq_search_gather(xb, cb, wv))
        by = 2 * N * C + 2 * K * C + 8 * N + 2 * N * C
        parity = None
        if with_cpu:
            # the codes of the very buffers just timed, audited on a row sample against the fp64 search of the
            # oracle on the same bf16 operands (SURVEY 8d tie policy); a real mismatch fails the bench
            from oracle import vq_oracle as vo
            codes = kn.vq_search_gather(xb, cb, wv)[0]
            rows = torch.arange(0, N, N // 4096, device=dev)[:4096]
            wr = torch.cat([wv[:-1].to(BF).float(), wv[-1:]]).cpu().numpy()
            parity = vo.audit_codes(xb[rows].float().cpu().numpy(), wr, codes[rows].cpu().numpy())
            parity["rows_checked"] = int(rows.numel())
            assert parity["real"] == 0, f"VQ codes of the timed buffers differ from the oracle at K={K}: {parity}"
        sweep.append({"K": K, "ms_per_launch": msv, "GBps": by / msv / 1e6, "codes_vs_fp64_oracle": parity,
                      "hbm_frac": (by / msv / 1e6 / peak_h) if peak_h else None,
                      "tensor_tflops": 2.0 * N * K * C / msv / 1e9,
                      "tensor_frac": (2.0 * N * K * C / msv / 1e9 / peak_t) if peak_t else None,
                      "tensor_frac_of_sustained": (2.0 * N * K * C / msv / 1e9 / sustained) if sustained else None})
    k1 = sweep[1]
    # the reference algorithm of the same op on the host cores (numpy oracle port, quantize2_mask.py:29-55):
    # distances + argmin of a bounded row sample against the K=1024 codebook, scaled to rows/s
    vq_cpu = None
    try:
        if not with_cpu:
            raise RuntimeError("skipped (N > 1 or --no-cpu-baseline)")
        import numpy as np
        from oracle import vq_oracle as vo
        rows = 8192
        xs = xb[:rows].float().cpu().numpy()
        ws = vo.bf16_round(wv_k1024.cpu().numpy())
        vo.find_nearest_embedding(xs[:256], ws)
        t0 = time.perf_counter()
        reps = 0
        while reps < 1 or (time.perf_counter() - t0 < 3.0 and reps < 10):
            vo.find_nearest_embedding(xs, ws); reps += 1
        dt = (time.perf_counter() - t0) / reps
        vq_cpu = {"value": rows / dt, "unit": "rows/s", "cores": usable_cores(), "kind": "port",
                  "sample": f"numpy oracle (addmm + argmin, fp32) on {rows} of the 65536 rows, K=1024",
                  "gpu_rows_per_s": N / (k1["ms_per_launch"] / 1e3)}
    except Exception as e:                                   # a reported baseline, never a reason to fail the bench
        vq_cpu = {"unavailable": f"{type(e).__name__}: {e}"}
    roof_vq = {"kernel": "vq_search_kernel (N=65536, C=256, K=1024, search+gather)", "bound": "tensor",
               "note": "dense [N,C]x[C,K] contraction above the ridge for K >= 1024: tensor-bound (SURVEY 8d); the HBM "
                       "fraction is reported as the metric asks; K=256 is the memory-leaning point of the sweep",
               "achieved": k1["GBps"], "peak": peak_h, "unit": "GB/s", "frac": k1["hbm_frac"],
               "tensor_tflops": k1["tensor_tflops"], "tensor_frac": k1["tensor_frac"],
               "tensor_peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)",
               "ms_per_launch": k1["ms_per_launch"], "timing": "20 launches per CUDA-graph replay, median of 7, L2 flushed",
               "traffic": traffic("vq"), "sweep": sweep, "cpu_baseline": vq_cpu,
               "codes_vs_fp64_oracle": k1.get("codes_vs_fp64_oracle"),
               "burst_peak_tflops": burst, "tensor_frac_of_burst": (k1["tensor_tflops"] / burst) if burst else None}
    # the training shape of every stage-1 config: 32 images -> N = 32768 rows, K = 1024, search + gather + statistics
    try:
        Nt, K = 32768, 1024
        xt = xb[:Nt].contiguous()
        cbt = kn.Codebook(K, C, dev); cbt.refresh(wv_k1024)
        counts, sums = torch.zeros(K, device=dev), torch.zeros(K, C, device=dev)
        lacc = torch.zeros(1, device=dev)
        mst = graph_ms(lambda: kn.vq_search_gather(xt, cbt, wv_k1024, counts=counts, sums=sums, loss_acc=lacc))
        roof_vq["training_shape"] = {"N": Nt, "K": K, "ms_per_launch": mst, "tensor_tflops": 2.0 * Nt * K * C / mst / 1e9,
                                     "tensor_frac_of_burst": (2.0 * Nt * K * C / mst / 1e9 / burst) if burst else None,
                                     "what": "search + gather + loss + in-kernel EMA counts/sums"}
    except Exception as e:
        roof_vq["training_shape"] = {"unavailable": f"{type(e).__name__}: {e}"}
    return roof, roof_vq, roof_gn


def cpu_baseline(args):
    """The reported CPU baseline of the B200 arm = a short run of the reference arm in its OWN process (this one
    has the overlay on sys.path, so `modules.*` would resolve to the B200 classes here)."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config, "--steps", "6",
           "--warmup", "1", "--cpu-budget-s", "25", "--cpu-sample-batch", "1", "--batch", str(args.batch)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
        line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        return line["cpu_baseline"]
    except Exception as e:                        # a reported baseline, never a reason to fail the bench
        return {"unavailable": f"{type(e).__name__}: {e}"}


def real_loss_line(args):
    """`--loss real` (the reference's full training_step: LPIPS + PatchGAN + adaptive weight, both optimizer passes)
    timed in a child process after the headline run; carried in the headline line as `real_loss_step`."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    cmd = [sys.executable, os.path.abspath(__file__), "--loss", "real", "--steps", str(min(args.steps, 10)), "--warmup", "3",
           "--batch", str(args.batch)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        return {k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "gpu_launches", "e2e", "config")}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_aux(args):
    """each kernel timed alone with CUDA events (L2 flushed between launches), against the
    measured HBM copy bandwidth, with the CPU oracle of the same op timed on a bounded sample beside it."""
    import numpy as np
    import torch
    from dynamicvectorquantization_b200 import configs, kernels as kn
    from oracle import dqvae_oracle as orc
    from oracle import permuter_oracle as po
    from oracle import vq_family_oracle as vf
    configs.activate_overlay()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_h = peaks.get("hbm_gbs") or 6577.7
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    threads, cores = pick_threads(torch)

    def gpu_ms(fn, iters=20):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]

    def cpu_s(fn, budget=5.0):
        fn()
        t0 = time.perf_counter(); n = 0
        while n < 1 or (time.perf_counter() - t0 < budget and n < 20):
            fn(); n += 1
        return (time.perf_counter() - t0) / n

    def emit(name, unit_name, units, ms, bytes_, cpu_units, cpu_sec, sample, launches=1):
        print(json.dumps({
            "metric": f"{unit_name}/sec ({name})", "value": units / ms * 1e3, "unit": f"{unit_name}/s", "n_gpus": 1,
            "ms_per_call": ms, "higher_is_better": True, "data": "synthetic", "gpu_launches": launches,
            "config": {"workload": name, "l2": "256 MB flush between launches"},
            "roofline": {"bound": "hbm", "achieved": bytes_ / ms / 1e6, "peak": peak_h, "unit": "GB/s",
                         "frac": bytes_ / ms / 1e6 / peak_h, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback"},
            "cpu_baseline": {"value": cpu_units / cpu_sec, "unit": f"{unit_name}/s", "cores": threads, "kind": "port",
                             "sample": sample}}), flush=True)

    # ---- patch entropy (a10): B=32 images 256x256, 16x16 patches; bytes = image once + 4 B per patch
    from models.stage1_dynamic.dqvae_dual_entropy import Entropy
    b = 32
    x = torch.rand(b, 3, 256, 256, device=dev) * 2 - 1
    ent = Entropy(16, 256, 256)
    ms = gpu_ms(lambda: ent(x))
    xc = x[:4].cpu()
    sec = cpu_s(lambda: orc.patch_entropy(xc, 16))
    emit("patch entropy 256x256, 16x16 patches, batch 32", "images", b, ms, b * 3 * 256 * 256 * 4 + b * 256 * 4,
         4, sec, f"oracle (fp32 torch CPU, {threads} threads) on 4 images")
    # ---- stage-2 permuter (8f row 3): B=256 code maps 32x32 + grain maps 16x16, ~50 % fine
    b = 256
    grain = torch.randint(0, 2, (b, 16, 16), device=dev)
    idx = torch.randint(0, 1024, (b, 32, 32), device=dev)
    from modules.dynamic_modules.permuter import DualGrainSeperatePermuter
    perm = DualGrainSeperatePermuter()
    out = perm(idx, grain)
    lc, lf = out["coarse_content"].shape[1], out["fine_content"].shape[1]
    ms = gpu_ms(lambda: perm(idx, grain))
    gi, gg = idx[:8].cpu().numpy(), grain[:8].cpu().numpy()
    sec = cpu_s(lambda: po.forward(gi, gg))
    emit("dual-grain permuter forward, 256 code maps 32x32 (incl. the 2-int length read-back)", "maps", b, ms,
         b * (1024 + 256) * 8 + b * 3 * (lc + lf) * 8, 8, sec, "numpy oracle on 8 maps, 1 thread", launches=1)
    args4 = (out["coarse_content"], out["fine_content"], out["coarse_position"], out["fine_position"])
    ms = gpu_ms(lambda: perm.forward_back(*args4))
    a4 = [t[:8].cpu().numpy() for t in args4]
    sec = cpu_s(lambda: po.forward_back(*a4))
    emit("dual-grain permuter forward_back, 256 sequences -> 32x32 code maps", "maps", b, ms,
         b * 2 * (lc + lf) * 8 + b * 1024 * 8, 8, sec, "numpy oracle (element loop like the reference) on 8 maps, 1 thread")
    # ---- residual quantizer (8f row 2): B=32 latents 8x8x256, depth 4, K=16384 (RQ-VAE shape), eval forward
    from modules.vector_quantization.quantize_rqvae import RQBottleneck
    rq = RQBottleneck(latent_shape=(8, 8, 256), code_shape=(8, 8, 4), n_embed=16384, shared_codebook=True).to(dev).eval()
    with torch.no_grad():
        rq.codebooks[0].weight.normal_()
    z = torch.randn(32, 8, 8, 256, device=dev)
    with torch.no_grad():
        ms_eager = gpu_ms(lambda: rq(z))
        # the depth loop is ~60 short launches: eager time is the host's launch rate, so the call is also
        # timed as a CUDA-graph replay (what a stage-2 sampler that captures its step would see)
        ms = ms_eager
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                rq(z)
            ms = gpu_ms(g.replay)
        except Exception as e:
            sys.stderr.write(f"[bench --aux] graph capture of the RQ forward failed ({type(e).__name__}: {e})\n")
    n, c, k, d = 32 * 64, 256, 16384, 4
    w = rq.codebooks[0].weight.detach().cpu().numpy()
    zc = z[:4].cpu().numpy()
    sec = cpu_s(lambda: vf.rq_forward(zc, [w] * d, (8, 8, 256), (8, 8, 4)))
    flops = 2.0 * n * k * c * d
    line_bytes = d * (2 * n * c + 2 * k * c + 8 * n + 4 * n * c)
    print(json.dumps({
        "metric": "latents/sec (RQ bottleneck 8x8x256, depth 4, K=16384, eval forward)", "value": 32 / ms * 1e3,
        "unit": "latents/s", "n_gpus": 1, "ms_per_call": ms, "ms_per_call_eager": ms_eager,
        "higher_is_better": True, "data": "synthetic", "gpu_launches": 2 * d,
        "roofline": {"bound": "tensor", "achieved": flops / ms / 1e9, "peak": peaks.get("bf16_tflops_sustained"),
                     "unit": "TFLOP/s", "frac": (flops / ms / 1e9 / peaks["bf16_tflops_sustained"])
                     if peaks.get("bf16_tflops_sustained") else None, "traffic": None,
                     "note": f"N={n} rows per depth = 16 row tiles: the codebook is split over 9 CTAs per tile "
                             f"(144 of 148 SMs busy); timed as a CUDA-graph replay; algorithmic bytes {line_bytes}"},
        "cpu_baseline": {"value": 4 / sec, "unit": "latents/s", "cores": threads, "kind": "port",
                         "sample": f"numpy oracle on 4 latents ({threads} BLAS threads)"}}), flush=True)


def run_real_loss(args):
    """training_step of dqvae_dual_feat.py:88-119 as Lightning drives it with two optimizers -
    pass 0: AE forward, L1 + LPIPS + adaptive-weight GAN + codebook + budget loss, backward, Adam(AE);
    pass 1: AE forward again, hinge loss of the PatchGAN on real / reconstructed images, backward, Adam(D).
    VGG16 / lin heads are randomly initialised (no pretrained files offline): identical arithmetic."""
    os.environ.setdefault("B200DQ_ALLOW_RANDOM_VGG", "1")
    import torch
    from dynamicvectorquantization_b200 import configs, kernels as kn
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(2021)
    cfg = configs.stage1_config("dqvae-dual-r-05")
    cfg["params"]["lossconfig"] = configs.real_loss_config(configs._BUDGET_DUAL)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the loss module prints a banner like the reference's
        model = configs.build_model(cfg).to(dev).train()
    B = args.batch
    model.learning_rate = 4.5e-6 * B
    ae = [p for n, p in model.named_parameters() if not n.startswith("loss.") and p.requires_grad]
    dp = list(model.loss.discriminator.parameters())
    mk = lambda ps: torch.optim.Adam(ps, lr=model.learning_rate, betas=(0.5, 0.9), capturable=True, fused=True)
    opt_ae, opt_d = mk(ae), mk(dp)
    g = torch.Generator().manual_seed(2021)
    x_host = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).pin_memory()
    x_dev = x_host.to(dev)
    loss_host = torch.zeros(2).pin_memory()

    def step(x):
        opt_ae.zero_grad(set_to_none=True)
        xrec, qloss, indices, gate = model(x)[:4]
        l0, _ = model.loss(qloss, x, xrec, 0, 0, last_layer=model.get_last_layer(), split="train", gate=gate)
        l0.backward()
        opt_ae.step()
        opt_d.zero_grad(set_to_none=True)
        xrec, qloss, indices, gate = model(x)[:4]
        l1, _ = model.loss(qloss, x, xrec, 1, 0, last_layer=model.get_last_layer(), split="train")
        l1.backward()
        opt_d.step()
        return torch.stack([l0.detach(), l1.detach()])

    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    torch.cuda.synchronize()
    graph, launches_per_step = None, None
    if not args.no_graph:
        try:
            static_x = x_dev.clone()
            l0c = kn.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = step(static_x)
            launches_per_step = kn.launch_count() - l0c
            eager = step

            def step(x):                              # noqa: F811
                if x is not static_x:
                    static_x.copy_(x, non_blocking=True)
                graph.replay()
                return static_out
            step(static_x)
            x_dev = static_x
        except Exception as e:
            sys.stderr.write(f"[bench --loss real] CUDA graph capture failed ({type(e).__name__}: {e}); eager\n")
            graph = None
            torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start()
    c0 = kn.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step(x_dev)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = launches_per_step * args.steps if graph is not None else kn.launch_count() - c0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(step, x_host, dev, args.steps, lambda out: loss_host.copy_(out, non_blocking=True))
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop()
    print(json.dumps({
        "metric": "images/sec (256x256 DQ-VAE training_step, LPIPS + PatchGAN loss, both optimizer passes)",
        "value": B * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOADS["dqvae-dual-r-05"] + ", lossconfig of the reference YAML", "global_batch": B,
                   "step": "2 AE forwards + AE backward + LPIPS fwd/bwd + 3 discriminator forwards + backwards + 2 Adam",
                   "cuda_graph": graph is not None, "weights": "random init (VGG16 / lin heads / AE / D)",
                   "discriminator": "PatchGAN on the hand-written kernels (4x4 tap GEMMs, BatchNorm + LeakyReLU kernels)"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": B * args.steps / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                "d2h_bytes_per_step": 8, "input_pipeline": "double-buffered prefetch on a copy stream"},
        "last_losses": [float(v) for v in loss_host]}), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.loss == "real":
        run_real_loss(a)
    elif a.aux:
        run_aux(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
