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


class GradAllReducer:
    def __init__(self, model: torch.nn.Module, process_group=None, overlap: bool = True, compress: bool = False):
        self.model = model
        self.compress = bool(compress) and not overlap     # deferred mode only: bf16 gradients on the wire
        self._wire = None
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
        rest = [p.grad for p in self.model.parameters()
                if p.grad is not None and (base is None or p.grad.untyped_storage().data_ptr() != base)]
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
