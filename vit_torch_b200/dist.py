"""Batch-sharded data parallelism for the fused ViT path: one process per GPU, bucketed NCCL gradient all-reduce
overlapped with backward (the reference itself is single-GPU, utils_network.py:406-452).

Buckets are the flat fp32 gradient buffers that functional.BlockFn.backward fills (one per transformer block, in
reverse layer order as backward runs) plus one trailing bucket for the remaining parameters (patch_embed, cls/pos,
final norm, head). Each block bucket is all-reduced (average) on a dedicated communication stream as soon as the block's
backward has been enqueued, so the 12 collectives overlap the rest of backward; `finish()` joins the streams before the
optimizer step. No DistributedDataParallel wrapper: unused parameters (the DINO fine-tune `head`)
simply never produce a bucket.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import functional as Fn


_partition_limit = 0     # SMs left to the persistent kernels while collectives overlap them (0 = no partition)


def configure_sm_partition(nccl_ctas: int = 4) -> None:
    """Call BEFORE dist.init_process_group. Caps the NCCL kernels at `nccl_ctas` thread blocks (NCCL_MAX_CTAS, unless
    the user already set it) and sizes the persistent GEMM grids for the remaining SMs. The GEMMs are statically
    scheduled, one CTA per SM: if an overlapping all-reduce kernel holds an SM when a GEMM launches, that GEMM's late CTA
    still owns 1/148 of the tiles and the kernel takes twice as long; leaving the SMs free costs nccl_ctas/148 of GEMM
    throughput instead. Measured on 2 B200 (ViT-B/16 bs128): NCCL default 21.09 ms/step, 16 blocks 20.87, 8 blocks 20.60,
    4 blocks 20.16; 4 blocks still move the 343 MB of fp32 gradients well inside one backward pass."""
    import os

    from . import ops
    if nccl_ctas <= 0:
        return
    global _partition_limit
    os.environ.setdefault("NCCL_MAX_CTAS", str(nccl_ctas))
    n = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    # The limit is applied only while all-reduces are in flight (GradAllReducer: from the first bucket of a backward
    # pass until finish()): forward, loss and optimiser run on all SMs. ViT-B/16 bs128 has 591 output tiles in its
    # N = 768 GEMMs -- 3.99 waves on 148 SMs but 4.10 (= 5 rounds) on 144 -- so a permanent limit costs the forward
    # proj / fc2 GEMMs a whole extra round.
    _partition_limit = max(n - int(os.environ["NCCL_MAX_CTAS"]), n // 2)


def partition_limit(nccl_ctas: int) -> int:
    """SMs left to the persistent kernels while a collective capped at nccl_ctas thread blocks shares the GPU."""
    n = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return max(n - int(nccl_ctas), n // 2)


def default_mode(world: int) -> str:
    """Data-parallel mode `bench.py --dp auto` (and a caller without a preference) should use for `world` GPUs of one
    NVSwitch node. Measured on B200, ViT-B/16 bs128 per GPU (profiles/r02_summary.md section 4): the split-graph mode wins
    at 2 GPUs (19.20 vs 19.36 ms/step) and loses at 8 (20.67 vs 20.10: its partitioned second graph costs more than the
    hidden part of the all-reduce saves), so it is the default for 2 GPUs only."""
    if world <= 1:
        return "none"
    return "split" if world == 2 else "deferred"


class GradAllReducer:
    def __init__(self, model: torch.nn.Module, process_group=None, overlap: bool = True, compress: bool = False,
                 split: bool = False, split_ctas: int = 4, register_arena: bool = False):
        self.model = model
        # register_arena: the Trainer's gradient arena is allocated by NCCL's own allocator (ncclMemAlloc) and registered
        # with the communicator, so that NVLS (in-switch) all-reduces read and write it zero-copy
        self.register_arena = bool(register_arena)
        self.arena_registered = False
        self._pools = []
        self.compress = bool(compress) and not overlap     # deferred mode only: bf16 gradients on the wire
        self._wire = None
        # split mode (graph Trainer): backward is captured as two graphs; the gradients of the first half (the last
        # blocks) are all-reduced on a side stream by a communicator capped at `split_ctas` thread blocks WHILE the second
        # graph runs on the other SMs; only the second half's all-reduce is exposed.
        self.split = bool(split) and not overlap
        self.split_ctas = int(split_ctas)
        self.pg_side = None
        self._side_done = None
        if self.split and dist.is_initialized() and dist.get_world_size(process_group) > 1 \
                and next(model.parameters()).is_cuda:
            self.world = dist.get_world_size(process_group)
            self.cuda = True
            self.side_group()           # collective call: every rank constructs its reducer
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        self.cuda = next(model.parameters()).is_cuda
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        self._pending = []   # (flat buffer, work handle)
        self._bucketed_ptrs = set()
        self.collectives = 0
        self._limited = False
        if self.world > 1:
            Fn.grad_bucket_hooks.append(self._on_bucket)

    def alloc_arena(self, numel: int, device):
        """Zeroed fp32 buffer from the communicator's registered memory pool, or None (plain allocation) when
        registration was not asked for or this torch / NCCL build cannot do it."""
        if not self.register_arena or not (dist.is_initialized() and self.world > 1 and self.cuda):
            return None
        try:
            pg = self.pg if self.pg is not None else dist.group.WORLD
            backend = pg._get_backend(torch.device(device))
            pool = torch.cuda.MemPool(backend.mem_allocator)
            with torch.cuda.use_mem_pool(pool):
                buf = torch.zeros((int(numel),), dtype=torch.float32, device=device)
            backend.register_mem_pool(pool)
            self._pools.append(pool)          # the pool owns the memory: keep it alive as long as the reducer
            self.arena_registered = True
            return buf
        except Exception as exc:              # older NCCL / torch without ncclMemAlloc: fall back to a plain buffer
            import sys
            print(f"[vit_torch_b200] gradient arena not registered with NCCL ({type(exc).__name__}: {exc})",
                  file=sys.stderr)
            return None

    def close(self):
        if self._on_bucket in Fn.grad_bucket_hooks:
            Fn.grad_bucket_hooks.remove(self._on_bucket)

    # called from BlockFn.backward on the compute stream once the block's flat gradient buffer is fully enqueued
    def _on_bucket(self, buf: torch.Tensor, params, views):
        if not self.overlap:
            self._pending.append((buf, None, params, views))
            return
        if self.cuda:
            if _partition_limit and not self._limited:      # collectives in flight from here to finish()
                from . import ops
                ops.set_sm_limit(_partition_limit)
                self._limited = True
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                work = dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
            buf.record_stream(self.comm_stream)
        else:  # gloo (CPU tests): no AVG, no streams
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self.collectives += 1
        self._pending.append((buf, work, params, views))

    def discard_pending(self):
        """Forget bucket notifications (used after a CUDA-graph capture of backward: replays fire no hooks)."""
        self._pending.clear()

    def side_group(self):
        """Communicator for the overlapped half: NCCL thread blocks capped (they share the GPU with the second graph)."""
        if self.pg_side is None and self.world > 1 and self.cuda:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = self.split_ctas
            opts.config.min_ctas = min(self.split_ctas, 1) or 1
            self.pg_side = dist.new_group(ranks=list(range(self.world)), pg_options=opts)
        return self.pg_side

    def reduce_first_half(self, arena, end: int) -> None:
        """Enqueue the all-reduce of arena[0:end] on the communication stream, ordered after everything enqueued so far
        on the current stream. Returns immediately; reduce_second_half() joins it."""
        if self.world <= 1 or end <= 0:
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            dist.all_reduce(arena.buf[:end], op=dist.ReduceOp.AVG, group=self.side_group())
            self._side_done = torch.cuda.Event()
            self._side_done.record(self.comm_stream)
        self.collectives += 1

    def reduce_second_half(self, arena, start: int) -> None:
        """All-reduce arena[start:used] on the current stream (all SMs), then join the first half."""
        if self.world <= 1:
            return
        if arena.off > start:
            dist.all_reduce(arena.buf[start:arena.off], op=dist.ReduceOp.AVG, group=self.pg)
            self.collectives += 1
        self._reduce_rest(arena.buf.untyped_storage().data_ptr())
        if self._side_done is not None:
            torch.cuda.current_stream().wait_event(self._side_done)
            self._side_done = None

    def _reduce_arena(self, arena):
        """ONE collective over the used part of the gradient arena (average). compress: the arena is cast to bf16 by one
        kernel, reduced in bf16 (half the bytes on NVLink) and cast back into the fp32 arena the optimiser reads."""
        buf = arena.used()
        if self.compress and self.cuda:
            from . import ops
            if self._wire is None or self._wire.numel() < buf.numel():
                self._wire = torch.empty((arena.buf.numel(),), dtype=torch.bfloat16, device=buf.device)
            wire = self._wire[: buf.numel()]
            ops.cast_bf16(buf, out=wire)
            dist.all_reduce(wire, op=dist.ReduceOp.AVG, group=self.pg)
            ops.cast_f32_from_bf16(wire, buf)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.AVG if self.cuda else dist.ReduceOp.SUM, group=self.pg)
            if not self.cuda:
                buf.div_(self.world)
        self.collectives += 1

    def finish_static(self, arena):
        """Deferred reduction without bucket notifications (the Trainer replays a captured forward + backward): one
        collective over the used part of the gradient arena, one over the gradients that live outside it."""
        if self.world <= 1:
            return
        op = dist.ReduceOp.AVG if self.cuda else dist.ReduceOp.SUM
        base = None
        if arena is not None and arena.off > 0:
            self._reduce_arena(arena)
            base = arena.buf.untyped_storage().data_ptr()
        self._reduce_rest(base)

    def _reduce_rest(self, arena_base) -> None:
        """All-reduce (average) the gradients that do not live in the arena (e.g. pos_embed when its gradient comes back
        through torch's interpolate for an input that is not the training resolution)."""
        op = dist.ReduceOp.AVG if self.cuda else dist.ReduceOp.SUM
        rest = [p.grad for p in self.model.parameters()
                if p.grad is not None and (arena_base is None or p.grad.untyped_storage().data_ptr() != arena_base)]
        if rest:
            flat = torch.cat([g.reshape(-1) for g in rest])
            dist.all_reduce(flat, op=op, group=self.pg)
            if not self.cuda:
                flat.div_(self.world)
            self.collectives += 1
            off = 0
            for g in rest:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n

    def finish(self):
        """Join the overlapped collectives, reduce whatever was not covered by a block bucket, and make sure every
        param.grad holds the averaged gradient. Call after loss.backward(), before optimizer.step()."""
        if self.world <= 1:
            return
        if self._limited:       # what is launched from here on is ordered after the joined collectives: all SMs again
            from . import ops
            ops.set_sm_limit(0)
            self._limited = False
        covered = set()
        # deferred mode with a gradient arena (train.Trainer): the blocks' buffers are adjacent slices of one buffer,
        # so ONE collective at full NVLink bandwidth covers them all
        coalesced = False
        arena = Fn.grad_arena
        if (arena is not None and self._pending and all(w is None for _, w, _, _ in self._pending)
                and all(b.untyped_storage().data_ptr() == arena.buf.untyped_storage().data_ptr()
                        for b, _, _, _ in self._pending)):
            self._reduce_arena(arena)
            coalesced = True
        for buf, work, params, views in self._pending:
            if coalesced:
                pass
            elif work is None:
                dist.all_reduce(buf, op=dist.ReduceOp.AVG if self.cuda else dist.ReduceOp.SUM, group=self.pg)
                self.collectives += 1
            else:
                work.wait()
            if not self.cuda and not coalesced:
                buf.div_(self.world)
            # .grad normally IS the bucket view (autograd steals it when .grad was None); if autograd cloned or
            # accumulated instead, overwrite it with the reduced bucket contents.
            for p, v in zip(params, views):
                if p is None or v is None:
                    continue
                covered.add(id(p))
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    p.grad.copy_(v)
        self._pending.clear()
        abase = arena.buf.untyped_storage().data_ptr() if (coalesced and arena is not None) else None
        rest = [p.grad for p in self.model.parameters() if p.grad is not None and id(p) not in covered
                and (abase is None or p.grad.untyped_storage().data_ptr() != abase)]   # arena slices are reduced already
        if rest:
            flat = torch.cat([g.reshape(-1) for g in rest])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG if self.cuda else dist.ReduceOp.SUM, group=self.pg)
            if not self.cuda:
                flat.div_(self.world)
            self.collectives += 1
            off = 0
            for g in rest:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
