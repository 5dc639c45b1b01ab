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
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError('FlatGradBucket expects float32 parameters')
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

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
