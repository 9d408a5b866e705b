"""Scene-sharded data parallelism.

The reference has no multi-GPU mode.  Here every rank runs the whole network on its own scenes and builds its
own rulebooks; the only exchange is one all-reduce (sum, then 1/world) of the parameter gradients, which are
laid out as views into ONE flat fp32 buffer (UNet-m64: 43.4 M parameters = 174 MB) so that a single NCCL call
over NVLink/NVSwitch moves them.  BatchNorm statistics stay per rank (no SyncBN), which is what running the
reference independently on each shard would do.  Works with the gloo backend for CPU tests."""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    def __init__(self, params, world_size=None, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = world_size if world_size is not None else dist.get_world_size(group)
        if not self.params:
            self.flat = None
            return
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            n = p.numel()
            view = self.flat[off:off + n].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view                      # autograd accumulates into this view in place
            off += n

    def check_views(self):
        """True when every .grad still aliases the flat buffer (zero_grad(set_to_none=True) would break it)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def all_reduce(self):
        if self.flat is None or self.world == 1:
            return
        if not self.check_views():
            raise RuntimeError("parameter .grad no longer aliases the flat bucket; use zero_grad(set_to_none=False)")
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
