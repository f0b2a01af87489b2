"""Batch sharding of the layer across the GPUs of one box (one process per GPU, ``torch.distributed``).

The ray-shooting map is sample-wise, so the data-parallel layout needs no collective in forward or backward:
every rank holds a replica of the (KB-sized) plan and a contiguous slice of the batch.  The
only exchange is optional: ``all_gather_outputs`` for a downstream loss that couples samples across the
batch (NCCL all-gather forward; the backward is the matching reduce-scatter of the incoming gradient: this rank's
rows, summed over the ranks' copies of the loss).  The reference has no distributed code (single device).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world_size):
    """[lo, hi) of this rank's contiguous slice; the first ``batch % world_size`` ranks get one more sample."""
    base, extra = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, rank=None, world_size=None):
    """This rank's slice of a replicated batch tensor (a view, no copy)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


class _AllGatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_local, sizes, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        ctx.rank, ctx.sizes, ctx.group = rank, sizes, group
        y_local = y_local.contiguous()
        if len(set(sizes)) == 1:
            out = y_local.new_empty((sum(sizes),) + tuple(y_local.shape[1:]))
            dist.all_gather_into_tensor(out, y_local, group=group)
            return out
        # ragged shards: pad every shard to the largest one (equal-size gather works on every backend)
        top = max(sizes)
        padded = y_local.new_zeros((top,) + tuple(y_local.shape[1:]))
        padded[:y_local.shape[0]] = y_local
        parts = [torch.empty_like(padded) for _ in sizes]
        dist.all_gather(parts, padded, group=group)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)

    @staticmethod
    def backward(ctx, g):
        # every rank holds the gradient w.r.t. the full gathered tensor of ITS copy of the loss; the gradient of
        # this rank's rows is the sum over ranks of those copies' slices: a reduce-scatter (1/world of the bytes an
        # all-reduce of the whole [B, ...] gradient would move).  The incoming gradient is never modified in place:
        # autograd may share it with other consumers (retain_grad, hooks).
        sizes, rank, group = ctx.sizes, ctx.rank, ctx.group
        world = len(sizes)
        lo = sum(sizes[:rank])
        if dist.get_backend(group) == "nccl":
            if len(set(sizes)) == 1:
                out = g.new_empty((sizes[rank],) + tuple(g.shape[1:]))
                dist.reduce_scatter_tensor(out, g.contiguous(), op=dist.ReduceOp.SUM, group=group)
                return out, None, None
            top = max(sizes)
            padded = g.new_zeros((world * top,) + tuple(g.shape[1:]))
            off = 0
            for r, s in enumerate(sizes):
                padded[r * top:r * top + s] = g[off:off + s]
                off += s
            out = g.new_empty((top,) + tuple(g.shape[1:]))
            dist.reduce_scatter_tensor(out, padded, op=dist.ReduceOp.SUM, group=group)
            return out[:sizes[rank]], None, None
        # backends without reduce-scatter (gloo, the CPU tests): all-reduce a private copy, keep this rank's rows
        total = g.clone(memory_format=torch.contiguous_format)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return total[lo:lo + sizes[rank]], None, None


def all_gather_outputs(y_local, batch=None, group=None):
    """[B_local, ...] on every rank -> [B, ...] on every rank, differentiable.

    ``batch`` (the global batch size) is only needed when it does not divide evenly.  If the downstream loss
    is computed redundantly on every rank and then averaged by DDP-style gradient averaging, scale it by
    ``1 / world_size`` as usual.
    """
    world = dist.get_world_size(group)
    if batch is None:
        sizes = [int(y_local.shape[0])] * world
    else:
        sizes = [shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world)]
    return _AllGatherRows.apply(y_local, sizes, group)


# --------------------------------------------------------------------------- the exchange step over peer memory
class PeerGather:
    """All-gather of the layer's output over NVLink / NVSwitch peer memory instead of a library collective.

    Every rank owns a [world * rows, k] float32 buffer in *symmetric memory* (``torch.distributed._symmetric_memory``:
    the plumbing that exchanges the handles; one allocation per rank, mapped into every peer, plus -- on NVSwitch boxes
    -- one multicast mapping that a single store reaches all ranks through).  Two ways to fill it:

    * ``all_gather(y_local)``: ONE kernel of this library (``rayen_gather_push_f32``) reads this rank's rows and stores
      them into every rank's buffer: one ``multimem.st`` per 16 bytes through the multicast mapping (the switch
      replicates), or plain peer stores when there is no multicast mapping.
    * ``fused_output()``: the address to hand to the forward kernels as their ``y`` -- this rank's rows *inside the
      multicast mapping*.  The kernels' ordinary ``y`` stores are then the all-gather: no separate pass over ``y`` at
      all (the epilogue of the forward kernels writes every rank's copy).  Only for plans whose forward kernels never
      read ``y`` back (n <= 32 without a big LMI) and only with a multicast mapping.

    Both are followed by ``barrier()`` (the symmetric-memory handle's device-side barrier on the current stream) before
    anyone reads the gathered rows.  Two buffers alternate, so that a rank may start step t+1 while a peer still reads
    the rows of step t.  Backward of the exchange: NCCL reduce-scatter of the incoming gradient (``_AllGatherRows``).
    """

    def __init__(self, rows_local, k, group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _cabi
        self._cabi = _cabi
        self.group = dist.group.WORLD if group is None else group
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.rows, self.k = int(rows_local), int(k)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = device
        self.buffers, self.handles = [], []
        for _ in range(2):
            buf = symm_mem.empty((self.world * self.rows, self.k), dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(buf, self.group)
            self.buffers.append(buf)
            self.handles.append(hdl)
        self.turn = 0
        import ctypes
        self._ptr_arrays = []
        for hdl in self.handles:
            arr = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            self._ptr_arrays.append(arr)
        self.multicast = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.handles]

    @property
    def has_multicast(self):
        return all(m != 0 for m in self.multicast)

    def _next(self):
        self.turn ^= 1
        return self.turn

    def barrier(self, which=None):
        self.handles[self.turn if which is None else which].barrier(channel=0)

    def all_gather(self, y_local, use_multicast=True):
        """y_local [rows, k] (float32, contiguous, this device) -> the gathered [world * rows, k] buffer (valid until the
        call after next).  One kernel + one barrier on the current stream."""
        assert y_local.shape == (self.rows, self.k) and y_local.dtype == torch.float32 and y_local.is_contiguous()
        t = self._next()
        import ctypes
        mc = self.multicast[t] if (use_multicast and self.multicast[t]) else 0
        rc = self._cabi.lib().rayen_gather_push_f32(
            y_local.data_ptr(), self.rows, self.k, ctypes.cast(self._ptr_arrays[t], ctypes.POINTER(ctypes.c_void_p)),
            self.world, mc or None, self.rank * self.rows, torch.cuda.current_stream(self.device).cuda_stream)
        self._cabi.check(rc, "rayen_gather_push_f32")
        self.barrier(t)
        return self.buffers[t]

    def fused_output(self):
        """(address for the forward kernels' y, buffer index): this rank's rows inside the multicast mapping."""
        if not self.has_multicast:
            raise RuntimeError("PeerGather.fused_output needs an NVSwitch multicast mapping (symmetric memory multicast_ptr)")
        t = self._next()
        return self.multicast[t] + 4 * self.rank * self.rows * self.k, t


class _GatheredRayShoot(torch.autograd.Function):
    """q [rows, n] on every rank -> y [world * rows, k] on every rank: the layer's forward kernels store y straight
    into every rank's gathered buffer through the multicast mapping (PeerGather.fused_output), then one barrier.
    Backward: reduce-scatter of the incoming gradient (this rank's rows, summed over the ranks), then the layer's
    closed-form backward."""

    @staticmethod
    def forward(ctx, q, module, gather):
        from . import _cabi
        from .constraint_module import _raw_stream
        v = q.detach()
        if v.dtype != torch.float32:
            v = v.float()
        if v.stride(1) != 1 or v.stride(0) < v.shape[1]:
            v = v.contiguous()
        B, cols = v.shape
        assert B == gather.rows and module.k == gather.k
        f = module._packed.fields
        if f.get("wide") or f.get("lmi_big"):
            raise RuntimeError("the fused all-gather epilogue needs forward kernels that never read y back "
                               "(n <= 32, LMI size <= 32); use PeerGather.all_gather for this set")
        st = module._launch_state(v.device)
        if not getattr(st, "_coalesced", False):
            st.plan.set_coalesced_output(True)   # whole rows per store instruction: full-width NVLink writes
            st._coalesced = True
        aux = torch.empty((2 * B + st.ws_words(B),), dtype=torch.float32, device=v.device)
        base = aux.data_ptr()
        y_ptr, t = gather.fused_output()
        want_grad = 1 if ctx.needs_input_grad[0] else 0
        with torch.cuda.device(v.device):
            rc = st.forward(st.handle, v.data_ptr(), v.stride(0), y_ptr, base, base + 4 * B, B, module._mode, want_grad,
                            base + 8 * B, _raw_stream(st.index))
        if rc != 0:
            _cabi.check(rc, "rayen_forward_f32")
        gather.barrier(t)
        ctx.module, ctx.in_dtype, ctx.aux, ctx.have_dkappa, ctx.state, ctx.gather = module, q.dtype, aux, want_grad, st, gather
        ctx.save_for_backward(v)
        module._last_aux = (aux, B)
        return gather.buffers[t]

    @staticmethod
    def backward(ctx, g_full):
        from . import _cabi
        from .constraint_module import _raw_stream
        (v,) = ctx.saved_tensors
        gather, st, module, aux = ctx.gather, ctx.state, ctx.module, ctx.aux
        B, cols = v.shape
        gy = g_full.new_empty((B, gather.k))
        dist.reduce_scatter_tensor(gy, g_full.contiguous().float(), op=dist.ReduceOp.SUM, group=gather.group)
        gv = torch.empty((B, cols), dtype=torch.float32, device=v.device)
        base = aux.data_ptr()
        with torch.cuda.device(v.device):
            rc = st.backward(st.handle, v.data_ptr(), v.stride(0), gy.data_ptr(), base, base + 4 * B, gv.data_ptr(), cols, B,
                             module._mode, ctx.have_dkappa, base + 8 * B, _raw_stream(st.index))
        if rc != 0:
            _cabi.check(rc, "rayen_backward_f32")
        return (gv if ctx.in_dtype == torch.float32 else gv.to(ctx.in_dtype)), None, None


def forward_gathered(module, q, gather):
    """``ConstraintModule`` forward on this rank's rows ``q`` [rows, n] with the all-gather fused into the forward
    kernels' epilogue: returns the gathered output [world * rows, k] (a view of the symmetric buffer, valid until the call
    after next), differentiable with respect to ``q``."""
    return _GatheredRayShoot.apply(q, module, gather)
