"""Scene-sharded data parallelism.

The reference has no multi-GPU mode.  Here every rank runs the whole network on its own scenes and builds its own
rulebooks; the only exchange is the all-reduce (mean) of the parameter gradients.  BatchNorm statistics stay per rank
(no SyncBN), which is what running the reference independently on each shard would do.

Gradients live as views into ONE flat fp32 buffer (UNet-m64: 43.4 M parameters = 174 MB), cut into buckets of
`bucket_mb`.  `BucketedGradAllReduce` launches the all-reduce of a bucket from autograd's post-accumulate hooks, as soon
as the last gradient of that bucket has been written, so the collective runs on NCCL's stream over NVLink/NVSwitch
UNDER the rest of the backward pass; `finish()` (before the optimizer step) only waits for the tail.  The buckets follow
the order in which gradients became ready in the first backward pass (deepest-consumed layers first), recorded once
and then frozen, like torch DDP's bucket rebuild.  The reduction is ReduceOp.AVG on NCCL (no separate scaling pass);
gloo (CPU tests) has no AVG, so there it is SUM followed by one multiply.  There is no compute kernel that feeds a
collective tile by tile on this path (the all-reduce consumes whole finished weight gradients), hence no fused
compute+collective kernel."""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """One flat bucket, one all-reduce after the backward pass (kept for small models and as the simplest correct form)."""

    def __init__(self, params, world_size=None, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = world_size if world_size is not None else dist.get_world_size(group)
        self.flat = None
        if self.params:
            self._layout(self.params)

    def _layout(self, order):
        """(re)build the flat buffer with the parameters in `order`; existing gradient values are carried over"""
        dev, dt = order[0].device, order[0].dtype
        flat = torch.zeros(sum(p.numel() for p in order), device=dev, dtype=dt)
        off = 0
        self.offsets = {}
        for p in order:
            n = p.numel()
            view = flat[off:off + n].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view                      # autograd accumulates into this view in place
            self.offsets[id(p)] = (off, off + n)
            off += n
        self.flat = flat

    def check_views(self):
        """True when every .grad still aliases the flat buffer (zero_grad(set_to_none=True) would break it)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def _avg(self, t, async_op=False):
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=False)
        t.mul_(1.0 / self.world)
        return w

    def all_reduce(self):
        if self.flat is None or self.world == 1:
            return
        if not self.check_views():
            raise RuntimeError("parameter .grad no longer aliases the flat bucket; use zero_grad(set_to_none=False)")
        self._avg(self.flat)

    finish = all_reduce


class BucketedGradAllReduce(FlatGradAllReduce):
    """Bucketed all-reduce overlapped with the backward pass.  Use: construct once, run backward, call finish() before
    optimizer.step(), keep zero_grad(set_to_none=False)."""

    def __init__(self, params, world_size=None, group=None, bucket_mb=25.0):
        super().__init__(params, world_size, group)
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.order = []              # parameters in the order their gradients became ready (first pass)
        self.frozen = False
        self.buckets = []            # [(start, end, n_params)]
        self.bucket_of = {}
        self.pending = []
        self.work = []
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _freeze(self):
        seen = {id(p) for p in self.order}
        order = self.order + [p for p in self.params if id(p) not in seen]     # never-ready parameters go last
        self._layout(order)
        self.buckets, self.bucket_of = [], {}
        start, count = 0, 0
        for p in order:
            s, e = self.offsets[id(p)]
            self.bucket_of[id(p)] = len(self.buckets)
            count += 1
            if (e - start) * self.flat.element_size() >= self.bucket_bytes:
                self.buckets.append((start, e, count))
                start, count = e, 0
        if count:
            self.buckets.append((start, self.flat.numel(), count))
        self.pending = [b[2] for b in self.buckets]
        self.frozen = True

    def _on_grad(self, p):
        if self.world == 1:
            return
        if not self.frozen:
            self.order.append(p)
            return
        b = self.bucket_of[id(p)]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            s, e, _ = self.buckets[b]
            # NCCL's stream waits for the work enqueued so far on the current (autograd) stream, then reduces this bucket
            # while the backward pass keeps going
            self.work.append((b, self._avg(self.flat[s:e], async_op=True)))

    def finish(self):
        """Wait for the buckets in flight and reduce whatever has not been launched (first pass: everything)."""
        if self.flat is None or self.world == 1:
            return
        if not self.frozen:
            self._avg(self.flat)              # first pass: one flat all-reduce, then fix the bucket order
            self._freeze()
            return
        if not self.check_views():
            raise RuntimeError("parameter .grad no longer aliases the flat bucket; use zero_grad(set_to_none=False)")
        launched = {b for b, _ in self.work}
        for b, (s, e, _) in enumerate(self.buckets):
            if b not in launched:             # parameters that received no gradient this step
                self.work.append((b, self._avg(self.flat[s:e], async_op=True)))
        for _, w in self.work:
            if w is not None and hasattr(w, "wait"):
                w.wait()
        self.work = []
        self.pending = [b[2] for b in self.buckets]

    all_reduce = finish
