"""The reference's optimisation step (utils_network.py:406-452) around the fused model: forward, CrossEntropyLoss,
zero_grad / backward, (DP: bucketed all-reduce), SGD with momentum 0.9 (utils_network.py:119-126)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def reset_parameters_like_zoo(m: nn.Module) -> None:
    """VisionModelZoo.reset_parameters (models/vision_all.py:322-329): recursive reset_parameters(); applied to DINO
    models built with pretrained=False (models/vision_all.py:157-158)."""
    for c in m.children():
        reset_parameters_like_zoo(c)
    if hasattr(m, "reset_parameters"):
        m.reset_parameters()


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD(params, lr, momentum) -- the reference's 'sgd' entry (utils_network.py:119-126) -- as ONE
    multi-tensor kernel per step that also refreshes the bf16 GEMM-operand copies of the weights
    (vitk_sgd_momentum_multi). Same constructor keywords, `param_groups[i]['lr']` stays the knob LambdaLR turns
    (utils_network.py:218-225), state_dict() keeps torch's `momentum_buffer` naming. dampening / nesterov /
    weight_decay other than the reference's zeros raise."""

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0, weight_decay=0, nesterov=False):
        if dampening != 0 or weight_decay != 0 or nesterov:
            raise NotImplementedError("FusedSGD implements the reference configuration: plain momentum SGD")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=0, weight_decay=0, nesterov=False))
        self._plan = {}       # group index -> cached launch plan
        self._pinned = []     # rotating pinned staging buffers for the pointer tables
        self._slot = 0

    def _build_plan(self, ps, dev):
        import numpy as np

        from . import ops
        chunk = ops._lib.load().vitk_sgd_chunk_elems()
        rows, cmap, ws = [], [], []
        for t, p in enumerate(ps):
            st = self.state[p]
            if "momentum_buffer" not in st:  # torch: buf = g on the first step == momentum * 0 + g
                st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            w = self._bf16_copy(p)
            ws.append(w)
            rows.append([p.data_ptr(), 0, st["momentum_buffer"].data_ptr(), w.data_ptr() if w is not None else 0,
                         p.numel()])
            cmap += [(t, c) for c in range((p.numel() + chunk - 1) // chunk)]
        table = np.asarray(rows, dtype=np.int64).reshape(-1, 5)
        cmap_dev = torch.tensor(cmap, dtype=torch.int32).view(-1, 2).to(dev)
        return dict(ids=[id(p) for p in ps], table=table, cmap=cmap_dev, nchunks=len(cmap), ws=ws,
                    dev_table=torch.empty((len(ps), 5), dtype=torch.int64, device=dev))

    @staticmethod
    def _bf16_copy(p):
        """The live bf16 GEMM-operand copy of `p` held by functional.bf16_weight's cache, or None."""
        from . import functional
        ent = functional._wcache.get(id(p))
        if ent is not None and ent[0]() is p and ent[1] == p._version and ent[2] == p.data_ptr():
            return ent[3]
        return None

    @torch.no_grad()
    def step(self, closure=None):
        from . import ops
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise ops._lib.VitkError("FusedSGD needs contiguous fp32 CUDA parameters (no CPU fallback)")
            plan = self._plan.get(gi)
            if (plan is None or plan["ids"] != [id(p) for p in ps]
                    or any(self._bf16_copy(p) is not w for p, w in zip(ps, plan["ws"]))):
                plan = self._plan[gi] = self._build_plan(ps, ps[0].device)
            tab = plan["table"]
            for t, p in enumerate(ps):
                g = p.grad
                if not (g.is_contiguous() and g.dtype == torch.float32):
                    g = p.grad = g.contiguous().float()
                tab[t, 0] = p.data_ptr()
                tab[t, 1] = g.data_ptr()
            if len(self._pinned) < 8:
                self._pinned.append(torch.empty((len(ps), 5), dtype=torch.int64).pin_memory())
            pin = self._pinned[self._slot % len(self._pinned)]
            self._slot += 1
            if pin.shape != (len(ps), 5):
                pin = self._pinned[(self._slot - 1) % len(self._pinned)] = torch.empty((len(ps), 5),
                                                                                        dtype=torch.int64).pin_memory()
            pin.numpy()[...] = tab
            plan["dev_table"].copy_(pin, non_blocking=True)
            ops.sgd_momentum_multi(plan["dev_table"], plan["cmap"], plan["nchunks"], group["lr"], group["momentum"],
                                   1.0, False)
        return loss


class Trainer:
    """The reference's optimisation step (utils_network.py:406-452). With graph=True the whole step -- forward, loss,
    backward, fused optimiser: ~300 kernel launches -- is captured once into a CUDA graph on static input buffers and
    replayed (single GPU only: capturing the overlapped NCCL all-reduces measured no gain on 2 GPUs -- 20.11 vs 20.16
    ms/step -- and hung at process-group teardown; the arithmetic and launch sequence are those of the eager step)."""

    def __init__(self, model: nn.Module, lr: float = 1e-3, momentum: float = 0.9, reducer=None, fused_opt: bool = True,
                 graph: bool = False):
        self.model = model
        self.reducer = reducer
        params = [p for p in model.parameters() if p.requires_grad]
        self.opt = (FusedSGD if fused_opt else torch.optim.SGD)(params, lr=lr, momentum=momentum)
        self.extra_launches_per_step = 0
        self.use_graph = bool(graph) and reducer is None
        self._graph = None
        self._sx = self._sy = self._sloss = None
        self._eager_steps = 0
        self.launches_per_step = None   # libvitk launches inside one captured step
        self.last_correct = None        # 0-d device tensor: correct argmax predictions of the last step's batch

    def _step_eager(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        out = self.model(x)
        if out.dim() == 2 and out.dtype == torch.float32 and out.stride(1) == 1 and y.dtype == torch.int64:
            # fused loss + gradient + argmax-accuracy count (utils_network.py:85-95, 429-433), no host sync
            from . import functional
            loss, self.last_correct = functional.cross_entropy(out, y)
        else:   # e.g. soft / non-int64 targets: torch's loss (still on the GPU)
            loss = F.cross_entropy(out, y)
            self.last_correct = None
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return loss.detach()

    def _capture(self, x: torch.Tensor, y: torch.Tensor) -> None:
        from . import ops
        # (no warm-up step here: it would be an extra optimisation step; step() has already run two eager ones)
        self._sx, self._sy = x.clone(), y.clone()
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        with torch.cuda.graph(self._graph):
            self._sloss = self._step_eager(self._sx, self._sy)
        self.launches_per_step = ops.launch_count - n0

    def static_inputs(self):
        """(images, labels) device buffers the captured step reads; fill them (e.g. `copy_` from pinned host memory)
        and call step_static(). None before the graph exists."""
        return (self._sx, self._sy) if self._graph is not None else None

    def step_static(self) -> torch.Tensor:
        self._graph.replay()
        return self._sloss

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if not self.use_graph:
            return self._step_eager(x, y)
        if self._graph is None:
            if self._eager_steps < 2:        # first steps eager: bf16 weight cache, optimiser state, autotuned plans
                self._eager_steps += 1
                return self._step_eager(x, y)
            self._capture(x, y)
        if x.shape != self._sx.shape or y.shape != self._sy.shape or x.dtype != self._sx.dtype:
            return self._step_eager(x, y)    # a different batch shape (e.g. the last batch of an epoch)
        if x.data_ptr() != self._sx.data_ptr():
            self._sx.copy_(x, non_blocking=True)
        if y.data_ptr() != self._sy.data_ptr():
            self._sy.copy_(y, non_blocking=True)
        return self.step_static()
