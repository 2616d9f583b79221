"""Data-parallel gradient exchange of the stage-1 step, overlapped with the backward pass.

The reference trains under Lightning's ``DistributedDataParallel`` (``train.py:230``: ``accelerator="ddp"``): the
autoencoder gradients are averaged over the ranks in buckets while the backward is still running.  The overlay
modules work under ``torch.nn.parallel.DistributedDataParallel`` as they are (``tests/test_gpu_multi.py``); this
module is the same exchange in a form that can be captured in ONE CUDA graph together with the forward and the
backward (DDP's reducer cannot): after the exchange every gradient is a view into a flat fp32 buffer, the buffer is cut into
buckets in REVERSE parameter order (the decoder's gradients are ready first), a post-accumulate hook counts a
bucket's parameters down and, when the last one has its gradient, copies the bucket's gradients into its slice with
one multi-tensor copy, re-points the parameters' ``.grad`` at the slices and issues ``all_reduce(AVG)`` of the slice
on a side stream.  ``finish()`` joins the side stream.  Result: the NCCL kernels of all but the last bucket run under
the remaining backward GEMMs instead of after them.

Averages, not sums (``ReduceOp.AVG``), like DDP.  The summation order inside a bucket is NCCL's; it is the same on
every rank, so parameters stay bit-identical across ranks.
"""
import torch
import torch.distributed as dist


class BucketedGradExchange:
    def __init__(self, params, bucket_mb=25.0, group=None, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.overlap = overlap
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        self._span = {}
        self._view = {}
        for p in self.params:
            self._view[id(p)] = self.flat[off:off + p.numel()].view_as(p)
            self._span[id(p)] = (off, off + p.numel())
            off += p.numel()
        # buckets over the flat buffer, walking the parameters from the last one backwards
        cap = max(1, int(bucket_mb * (1 << 20) / 4))
        self.buckets = []                              # [lo, hi, n_params]
        self._members = []                             # parameters of each bucket
        hi = total
        members = []
        self._bucket_of = {}
        for p in reversed(self.params):
            lo = self._span[id(p)][0]
            self._bucket_of[id(p)] = len(self.buckets)
            members.append(p)
            if hi - lo >= cap:
                self.buckets.append([lo, hi, len(members)])
                self._members.append(members)
                hi, members = lo, []
        if members:
            self.buckets.append([0, hi, len(members)])
            self._members.append(members)
        self._pending = [b[2] for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        for p in self.params:
            p.grad = None
            p.register_post_accumulate_grad_hook(self._make_hook(p))

    # ------------------------------------------------------------------ per-step protocol
    def begin_step(self):
        """Drop the gradients (the backward then ASSIGNS each new gradient instead of adding it into an existing one:
        no per-parameter accumulation kernel) and re-arm the buckets."""
        for p in self.params:
            p.grad = None
        self._pending = [b[2] for b in self.buckets]
        self._launched = [False] * len(self.buckets)

    def finish(self):
        """Exchange whatever has not been sent yet (parameters that received no gradient this step), then make the
        current stream wait for every bucket.  Afterwards every ``p.grad`` is a view into the averaged flat buffer."""
        for i in range(len(self.buckets)):
            if not self._launched[i]:
                self._launch(i)
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    # ------------------------------------------------------------------ internals
    def _make_hook(self, p):
        def hook(param):
            i = self._bucket_of[id(p)]
            self._pending[i] -= 1
            if self._pending[i] == 0 and self.overlap and not self._launched[i]:
                self._launch(i)
        return hook

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _launch(self, i):
        """Move the bucket's gradients into its slice of the flat buffer with ONE multi-tensor copy, make the slices
        the parameters' gradients, and all-reduce the slice."""
        self._launched[i] = True
        lo, hi, _ = self.buckets[i]
        srcs, dsts = [], []
        for p in self._members[i]:
            v = self._view[id(p)]
            g = p.grad
            if g is None:
                v.zero_()                                            # no gradient this step: contributes zeros
            elif g.data_ptr() != v.data_ptr():
                srcs.append(g)
                dsts.append(v)
            p.grad = v
        if srcs:
            torch._foreach_copy_(dsts, srcs)
        if self._world() == 1:
            return
        chunk = self.flat[lo:hi]
        if self._side is not None:
            self._side.wait_stream(torch.cuda.current_stream())     # the bucket's gradients are complete on this stream
            with torch.cuda.stream(self._side):
                dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group)
        else:                                                        # CPU tensors (gloo tests): no AVG on gloo
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            chunk.div_(self._world())
