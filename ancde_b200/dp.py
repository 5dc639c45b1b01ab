"""Batch-sharded data parallelism for ANCDE training.

The dual-NCDE forward and backward need no communication: with fixed-grid rk4 every sample's two
recurrences are independent, so rank r simply takes its slice of the batch.  The only exchange is
ONE sum all-reduce of a flat FP32 gradient buffer per step (~0.5 MB for the Sepsis model), issued
through torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """All parameters' gradients as views of one contiguous FP32 buffer."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        # the parameter VALUES live in one buffer too (p.data become views): whole-model elementwise work -- the weight
        # regulariser of the training step -- is a handful of kernels instead of a handful per parameter
        self.flat_param = torch.empty(self.numel, dtype=torch.float32, device=dev)
        self.offset = {}
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError('FlatGradBucket expects float32 parameters')
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            p.grad = self.flat[off:off + n].view_as(p)
            self.offset[id(p)] = (off, n)
            off += n

    def segments(self, params):
        """(start, end, seg) of a run of parameters that sit next to each other in the flat buffers: seg[i] is the index
        (within `params`) of the parameter that element start + i belongs to."""
        spans = sorted(self.offset[id(p)] for p in params)
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            if o0 + n0 != o1:
                raise ValueError('the parameters are not contiguous in the flat buffer')
        start, end = spans[0][0], spans[-1][0] + spans[-1][1]
        seg = torch.cat([torch.full((n,), i, dtype=torch.int64) for i, (_, n) in enumerate(spans)])
        return start, end, seg.to(self.flat.device)

    def zero(self):
        """Zeroes every gradient in one fill (the views stay attached to the parameters)."""
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """Sum all-reduce of the flat buffer, then 1/world (mean-reduced losses)."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)


def shard_slice(n, rank, world):
    """Contiguous equal shards; the reference uses drop_last=True (datasets/common.py:17-18)."""
    per = n // world
    return slice(rank * per, (rank + 1) * per)


def shard(x, rank, world):
    """This rank's slice of dim 0 of a batch tensor."""
    return x[shard_slice(x.shape[0], rank, world)]
