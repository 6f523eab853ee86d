"""Data-parallel gradient exchange: one flat fp32 arena per optimisation phase, all-reduced in place.

Replaces DistributedDataParallel's bucketed reducer (reference: Trainer.py:44): every parameter's
`.grad` is a view into one contiguous buffer, so backward writes gradients straight into the arena,
a single NCCL all-reduce (NVLS in-switch reduction on NVSwitch systems) averages it across ranks, and
the fused Adam step consumes the same views.  No parameter broadcast per step, no unused-parameter
search: parameters outside the phase do not take part.
"""
import torch
import torch.distributed as dist


class GradArena:
    def __init__(self, params, world_size=1, chunk_mb=64):
        self.params = [p for p in params if p.requires_grad]
        # de-duplicate (motion_enc is shared by the CmpFlow and MotMask networks)
        seen, uniq = set(), []
        for p in self.params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        self.world_size = world_size
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.numel = total
        self.chunk = max(1, int(chunk_mb * (1 << 20) // 4))

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        """Average the arena over the ranks (no-op for a single process)."""
        if self.world_size <= 1 or not dist.is_initialized():
            return
        # one collective for the whole arena: NVSwitch bandwidth does not depend on peer count, so
        # bucket size only has to amortise launch latency (185-209 MB here).
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / self.world_size)

    def check_views(self):
        """True if every parameter's .grad still aliases the arena (autograd accumulates in place)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)
