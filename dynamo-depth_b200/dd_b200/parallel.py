"""Data-parallel plumbing: replica synchronisation at construction + one flat fp32 gradient arena per phase.

Replaces DistributedDataParallel (reference: Trainer.py:44):

* `broadcast_module_state` -- what DDP's constructor does: every parameter and buffer of rank 0 is copied to all
  ranks (one flat broadcast per dtype), so replicas start from identical weights whatever their local seeds were.
* `GradArena` -- every parameter's `.grad` is a view into one contiguous buffer laid out in the model's fixed module
  order (networks.model.MODULE_ORDER), so backward writes gradients straight into the arena and the fused Adam step
  consumes the same views.  The arena is cut into one chunk per sub-module; a chunk is all-reduced (NCCL `AVG`: the
  1/world factor is applied inside the collective, no extra pass over the arena) as soon as autograd has accumulated
  the last gradient of the chunk, i.e. while the rest of backward is still running -- the reducer of DDP that fires
  inside `loss.backward()` (Trainer.py:148), without buckets, copies or an unused-parameter search.
  Chunks are always issued in the same static order on every rank (reverse module order = the order backward
  completes them), and the arena layout is hashed and compared across ranks before the first collective.
"""
import hashlib

import torch
import torch.distributed as dist


def _dist_on(world_size):
    return world_size > 1 and dist.is_available() and dist.is_initialized()


def broadcast_module_state(module, src=0, group=None):
    """Copy rank `src`'s parameters and buffers to every rank (DDP construction semantics, Trainer.py:44).
    One flat broadcast per dtype; returns the number of elements broadcast."""
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    total = 0
    for dtype in sorted(by_dtype, key=str):
        group_t = by_dtype[dtype]
        flat = torch.cat([t.reshape(-1) for t in group_t])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in group_t:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        total += off
    return total


def layout_digest(named_sizes):
    """63-bit digest of [(name, numel), ...] -- equal on two ranks iff their arenas are laid out identically."""
    h = hashlib.sha1()
    for name, n in named_sizes:
        h.update(f"{name}:{int(n)};".encode())
    return int.from_bytes(h.digest()[:8], "big") >> 1


def assert_same_across_ranks(value, what, device, group=None):
    t = torch.tensor([value, -value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    if int(t[0]) != value or int(-t[1]) != value:
        raise RuntimeError(f"{what} differs between ranks (this rank {value}, max {int(t[0])}, min {int(-t[1])})")


class GradArena:
    def __init__(self, params, world_size=1, names=None, chunk_ids=None, overlap=True, group=None):
        """params: parameters in a rank-independent order; names: matching "<module>.<param>" strings (layout check);
        chunk_ids: one hashable per parameter, equal for the parameters of one chunk (consecutive runs = chunks)."""
        params = list(params)
        names = list(names) if names is not None else [f"p{i}" for i in range(len(params))]
        chunk_ids = list(chunk_ids) if chunk_ids is not None else [0] * len(params)
        assert len(names) == len(params) == len(chunk_ids)
        seen, keep = set(), []
        for i, p in enumerate(params):      # de-duplicate (motion_enc is shared by the CmpFlow and MotMask networks)
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                keep.append(i)
        self.params = [params[i] for i in keep]
        self.names = [names[i] for i in keep]
        self.world_size = world_size
        self.group = group
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.numel = total
        self.chunks = []                    # [start, end, [param indices]] in arena order
        off, last = 0, object()
        for j, p in enumerate(self.params):
            n = p.numel()
            # same strides as the parameter (channels_last ResNet weights): the fused optimiser requires it, and a dense
            # parameter of n elements maps one-to-one onto its n-element slice of the arena
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            p.grad = self.flat[off:off + n].as_strided(p.size(), p.stride()) if dense else self.flat[off:off + n].view_as(p)
            cid = chunk_ids[keep[j]]
            if not self.chunks or cid != last:
                self.chunks.append([off, off + n, [j]])
                last = cid
            else:
                self.chunks[-1][1] = off + n
                self.chunks[-1][2].append(j)
            off += n
        self.order = list(range(len(self.chunks)))[::-1]      # issue order: backward finishes the last module first
        self.digest = layout_digest([(n, p.numel()) for n, p in zip(self.names, self.params)])
        self.overlap = bool(overlap) and _dist_on(world_size)
        self._hooks, self._works = [], []
        self._expected = None               # per chunk: set of parameter indices that received a gradient in step 0
        self._fired = [set() for _ in self.chunks]
        self._next = 0                      # position in self.order of the next chunk to issue
        self._param_chunk = {}
        self.last_collectives = 0
        if _dist_on(world_size):
            assert_same_across_ranks(self.digest, "gradient-arena layout (parameter names / sizes / order)", dev, group)
            avg = getattr(dist.ReduceOp, "AVG", None)
            self._avg = avg if (avg is not None and dist.get_backend(group) == "nccl") else None
        if self.overlap:
            for ci, (_, _, idxs) in enumerate(self.chunks):
                for j in idxs:
                    self._param_chunk[j] = ci
                    self._hooks.append(self.params[j].register_post_accumulate_grad_hook(self._make_hook(j)))

    # ------------------------------------------------------------------ collectives
    def _reduce(self, lo, hi):
        view = self.flat[lo:hi]
        if self._avg is not None:
            return dist.all_reduce(view, op=self._avg, group=self.group, async_op=True), None
        return dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True), view

    def _issue_ready(self, force=False):
        while self._next < len(self.order):
            ci = self.order[self._next]
            if not force and (self._expected is None or not self._expected[ci] <= self._fired[ci]):
                break
            lo, hi, _ = self.chunks[ci]
            self._works.append(self._reduce(lo, hi))
            self._next += 1

    def _make_hook(self, j):
        def hook(_param):
            ci = self._param_chunk[j]
            if self._expected is not None and self.order.index(ci) < self._next and j not in self._expected[ci]:
                raise RuntimeError(f"gradient of {self.names[j]} arrived after its arena chunk was all-reduced "
                                   "(the set of parameters receiving gradients changed inside a phase)")
            self._fired[ci].add(j)
            if self._expected is not None:
                self._issue_ready()
        return hook

    def all_reduce(self, group=None):
        """Average the arena over the ranks: issues whatever backward has not already issued from its hooks (all of it
        in the first step of a phase, which only records which parameters receive gradients), then orders the current
        stream after the collectives.  No-op for a single process."""
        if not _dist_on(self.world_size):
            return
        if group is not None:
            self.group = group
        self._issue_ready(force=True)
        for work, scale_view in self._works:
            work.wait()
            if scale_view is not None:       # backends without AVG (gloo on the CPU test tier)
                scale_view.mul_(1.0 / self.world_size)
        self.last_collectives = len(self._works)
        if self._expected is None and self.overlap:
            self._expected = [set(f) for f in self._fired]
        self._works, self._next = [], 0
        self._fired = [set() for _ in self.chunks]

    # ------------------------------------------------------------------ misc
    def zero(self):
        self.flat.zero_()

    def release(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def check_views(self):
        """True if every parameter's .grad still aliases the arena (autograd accumulates in place)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base and p.grad.stride() == p.stride()
                   for p in self.params)
