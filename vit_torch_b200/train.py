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


class _FusedMultiTensor(torch.optim.Optimizer):
    """Shared plumbing of the fused multi-tensor optimisers: ONE kernel per step over a device table of per-tensor
    pointers (parameter, gradient, state buffers, live bf16 GEMM-operand copy), hyper-parameters in DEVICE memory so
    that a step captured in a CUDA graph keeps following `param_groups[i]['lr']` (the knob the reference's LambdaLR
    turns, utils_network.py:218-225): sync_hyper() rewrites them with a stream-ordered copy whenever they changed."""

    _state_names: tuple = ()
    _hp_sync_len = 0          # leading entries of the device hyper-parameter vector owned by the host
    _hp_len = 0

    def _init_fused(self):
        self._plan = {}       # group index -> cached launch plan
        self._pinned = []     # rotating pinned staging buffers for the pointer tables
        self._slot = 0
        self._hp_dev = {}     # group index -> device fp32 vector
        self._hp_host = {}    # group index -> last values written
        self._cap = {}        # group index -> (pinned table, device table) owned by a CUDA-graph capture

    def _hyper(self, group) -> list:
        raise NotImplementedError

    def _launch(self, plan, hp_dev) -> None:
        raise NotImplementedError

    def _hp(self, gi, dev):
        if gi not in self._hp_dev:
            self._hp_dev[gi] = torch.zeros((self._hp_len,), dtype=torch.float32, device=dev)
            self._hp_host[gi] = None
        return self._hp_dev[gi]

    def sync_hyper(self) -> None:
        """Make the device copies of the hyper-parameters match param_groups (no-op when nothing changed). Called by
        step(); the graph Trainer calls it before every replay."""
        for gi, group in enumerate(self.param_groups):
            if gi not in self._hp_dev:
                continue
            cur = [float(v) for v in self._hyper(group)]
            if cur != self._hp_host[gi]:
                self._hp_dev[gi][: self._hp_sync_len].copy_(torch.tensor(cur, dtype=torch.float32), non_blocking=False)
                self._hp_host[gi] = cur

    def _build_plan(self, ps, dev):
        import numpy as np

        from . import ops
        chunk = ops._lib.load().vitk_sgd_chunk_elems()
        ncol = 4 + len(self._state_names)
        rows, cmap, ws = [], [], []
        for t, p in enumerate(ps):
            st = self.state[p]
            for nm in self._state_names:
                if nm not in st:   # zeros reproduce torch's first step (buf = g; exp_avg = (1 - beta1) g; ...)
                    st[nm] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            w = self._bf16_copy(p)
            ws.append(w)
            rows.append([p.data_ptr(), 0] + [st[nm].data_ptr() for nm in self._state_names] +
                        [w.data_ptr() if w is not None else 0, p.numel()])
            cmap += [(t, c) for c in range((p.numel() + chunk - 1) // chunk)]
        table = np.asarray(rows, dtype=np.int64).reshape(-1, ncol)
        cmap_dev = torch.tensor(cmap, dtype=torch.int32).view(-1, 2).to(dev)
        return dict(ids=[id(p) for p in ps], table=table, cmap=cmap_dev, nchunks=len(cmap), ws=ws,
                    state_ptrs=self._state_ptrs(ps),
                    dev_table=torch.empty((len(ps), ncol), dtype=torch.int64, device=dev))

    def _state_ptrs(self, ps):
        return [tuple(self.state[p][nm].data_ptr() if nm in self.state[p] else 0 for nm in self._state_names)
                for p in ps]

    def prepare_capture(self) -> None:
        """Call right before a CUDA-graph capture of step(): the captured step gets its OWN pinned staging buffer and
        device pointer table, which no eager step() ever rewrites. (The graph holds the H2D copy of the table; with the
        rotating buffers an eager step for another batch shape would, eight steps later, overwrite the buffer that copy
        reads and every replay would then apply stale gradient pointers.) Allocation happens here because pinned /
        device allocations are not allowed while a stream is capturing."""
        self._cap = {}
        for gi, group in enumerate(self.param_groups):
            n = len(group["params"])
            if n == 0:
                continue
            ncol = 4 + len(self._state_names)
            dev = group["params"][0].device
            self._cap[gi] = (torch.empty((n, ncol), dtype=torch.int64).pin_memory(),
                             torch.empty((n, ncol), dtype=torch.int64, device=dev))

    def load_state_dict(self, state_dict):
        """torch replaces self.state with new tensors: the cached launch plans hold raw pointers into the old ones."""
        super().load_state_dict(state_dict)
        self._plan = {}
        self._after_load()

    def __setstate__(self, state):
        super().__setstate__(state)
        self._plan = {}

    def _after_load(self) -> None:
        pass

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
                    raise ops._lib.VitkError(f"{type(self).__name__} needs contiguous fp32 CUDA parameters "
                                             "(no CPU fallback)")
            plan = self._plan.get(gi)
            if (plan is None or plan["ids"] != [id(p) for p in ps]
                    or any(self._bf16_copy(p) is not w for p, w in zip(ps, plan["ws"]))
                    or plan["state_ptrs"] != self._state_ptrs(ps)):
                plan = self._plan[gi] = self._build_plan(ps, ps[0].device)
            tab = plan["table"]
            for t, p in enumerate(ps):
                g = p.grad
                if not (g.is_contiguous() and g.dtype == torch.float32):
                    g = p.grad = g.contiguous().float()
                tab[t, 0] = p.data_ptr()
                tab[t, 1] = g.data_ptr()
            capturing = torch.cuda.is_current_stream_capturing()
            if capturing:
                cap = self._cap.get(gi)
                if cap is None or cap[0].shape[0] < tab.shape[0]:
                    raise RuntimeError(f"{type(self).__name__}.step() under CUDA-graph capture needs prepare_capture() "
                                       "first (capture-private pointer tables)")
                pin, dev_table = cap[0][: tab.shape[0]], cap[1][: tab.shape[0]]
            else:
                if len(self._pinned) < 8:
                    self._pinned.append(torch.empty(tab.shape, dtype=torch.int64).pin_memory())
                pin = self._pinned[self._slot % len(self._pinned)]
                self._slot += 1
                if pin.shape != tab.shape:
                    pin = self._pinned[(self._slot - 1) % len(self._pinned)] = torch.empty(
                        tab.shape, dtype=torch.int64).pin_memory()
                dev_table = plan["dev_table"]
            pin.numpy()[...] = tab
            dev_table.copy_(pin, non_blocking=True)
            hp = self._hp(gi, ps[0].device)
            if not capturing:
                self.sync_hyper()     # (a capture records the launch only; the values are synced around it)
            self._launch(dict(plan, dev_table=dev_table) if capturing else plan, hp)
        return loss


class FusedSGD(_FusedMultiTensor):
    """torch.optim.SGD(params, lr, momentum) -- the reference's 'sgd' entry (utils_network.py:119-126) -- as ONE
    multi-tensor kernel per step that also refreshes the bf16 GEMM-operand copies of the weights
    (vitk_sgd_momentum_multi_hp). Same constructor keywords, `param_groups[i]['lr']` stays the knob LambdaLR turns
    (utils_network.py:218-225), state_dict() keeps torch's `momentum_buffer` naming. dampening / nesterov /
    weight_decay other than the reference's zeros raise."""

    _state_names = ("momentum_buffer",)
    _hp_sync_len = 3
    _hp_len = 3

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0, weight_decay=0, nesterov=False):
        if dampening != 0 or weight_decay != 0 or nesterov:
            raise NotImplementedError("FusedSGD implements the reference configuration: plain momentum SGD")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=0, weight_decay=0, nesterov=False))
        self._init_fused()

    def _hyper(self, group):
        return [group["lr"], group["momentum"], 1.0]

    def _launch(self, plan, hp_dev):
        from . import ops
        ops.sgd_momentum_multi_hp(plan["dev_table"], plan["cmap"], plan["nchunks"], hp_dev)


class FusedAdam(_FusedMultiTensor):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) -- the reference's 'adam' entry
    (utils_network.py:119-126) -- as one multi-tensor kernel (+ a one-thread tick kernel that advances the step counter
    and the bias corrections on the device) that also refreshes the bf16 weight copies (vitk_adam_multi). State keeps
    torch's names (`step`, `exp_avg`, `exp_avg_sq`; `step` is one device scalar shared by the group). amsgrad /
    maximize raise."""

    _state_names = ("exp_avg", "exp_avg_sq")
    _hp_sync_len = 8
    _hp_len = 11
    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError(f"{type(self).__name__}: amsgrad / maximize are not implemented")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._init_fused()

    def _hyper(self, group):
        b1, b2 = group["betas"]
        return [group["lr"], b1, b2, group["eps"], group["weight_decay"], 1.0 if self._decoupled else 0.0,
                1.0 - b1, 1.0 - b2]

    def _build_plan(self, ps, dev):
        plan = super()._build_plan(ps, dev)
        gi = next(i for i, g in enumerate(self.param_groups) if any(p is ps[0] for p in g["params"]))
        step = self._hp(gi, dev)[8]
        loaded = [self.state[p].get("step") for p in ps]
        loaded = [float(t) for t in loaded if t is not None and t.data_ptr() != step.data_ptr()]
        if loaded:      # resumed from a state_dict: continue the bias corrections from the loaded step count
            step.fill_(max(loaded))
        for p in ps:
            self.state[p]["step"] = step      # 0-d view of the group's device step counter
        return plan

    def _launch(self, plan, hp_dev):
        from . import ops
        ops.adam_multi(plan["dev_table"], plan["cmap"], plan["nchunks"], hp_dev)


class FusedAdamW(FusedAdam):
    """torch.optim.AdamW (decoupled weight decay, default 1e-2): the reference's 'adamw' entry."""

    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False,
                 maximize=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                         maximize=maximize)


class Trainer:
    """The reference's optimisation step (utils_network.py:406-452). With graph=True the whole step -- forward, loss,
    backward, fused optimiser: ~300 kernel launches -- is captured once into a CUDA graph on static input buffers and
    replayed (single GPU only: capturing the overlapped NCCL all-reduces measured no gain on 2 GPUs -- 20.11 vs 20.16
    ms/step -- and hung at process-group teardown; the arithmetic and launch sequence are those of the eager step)."""

    def __init__(self, model: nn.Module, lr: float = 1e-3, momentum: float = 0.9, reducer=None, fused_opt: bool = True,
                 graph: bool = False, optimizer: str = "sgd", strict_graph: bool = False):
        self.model = model
        self.strict_graph = strict_graph    # True: a failed capture raises (bench.py) instead of falling back to eager
        self.reducer = reducer
        params = [p for p in model.parameters() if p.requires_grad]
        # `optimizer`: the reference's --opt names (utils_network.py:119-126); sgd / adam / adamw have fused kernels
        if optimizer == "sgd":
            self.opt = (FusedSGD if fused_opt else torch.optim.SGD)(params, lr=lr, momentum=momentum)
        elif optimizer == "adam":
            self.opt = (FusedAdam if fused_opt else torch.optim.Adam)(params, lr=lr)
        elif optimizer == "adamw":
            self.opt = (FusedAdamW if fused_opt else torch.optim.AdamW)(params, lr=lr)
        else:
            raise NotImplementedError(f"optimizer {optimizer!r}: pass a torch.optim optimiser through the reference loop")
        self.extra_launches_per_step = 0
        # gradient arena: the per-block flat gradient buffers of a step are adjacent slices of one buffer (one memset
        # per step; one all-reduce per step in deferred data-parallel mode)
        self.arena = None
        if params and params[0].is_cuda:
            from . import functional
            # (+ slack: the head's weight gradients are padded to 8 rows, d_pos doubles as the source of d_cls / d_bias)
            self.arena = functional.GradArena(sum((p.numel() + 3) // 4 * 4 + 4 for p in params) + (1 << 16)
                                              + 2 * max((p.numel() for p in params if p.dim() == 3), default=0),
                                              params[0].device, alloc=getattr(reducer, "alloc_arena", None))
        # Data parallel: with the deferred all-reduce (dist.GradAllReducer(overlap=False)) forward + loss + backward are
        # captured and the collective and the optimiser kernel run eagerly after the replay; the overlapped mode issues
        # NCCL calls from inside backward and stays eager.
        self.use_graph = bool(graph) and (reducer is None or not getattr(reducer, "overlap", True))
        self.opt_in_graph = reducer is None
        # split mode (dist.GradAllReducer(split=True)): backward captured as two graphs around block `split_at`
        self.split = self.use_graph and reducer is not None and getattr(reducer, "split", False)
        self.split_at = None
        self._graph2 = None
        self._split_off = 0
        self._graph = None
        self._sx = self._sy = self._sloss = None
        self._eager_steps = 0
        self.launches_per_step = None   # libvitk launches inside one captured step
        self.last_correct = None        # 0-d device tensor: correct argmax predictions of the last step's batch

    def _step_eager(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        loss = self._fwd_bwd(x, y)
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return loss

    def _fwd_bwd(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self.arena is not None:
            from . import functional
            functional.grad_arena = self.arena
            self.arena.begin()              # (the previous step's gradients were consumed by its optimiser step)
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
        return loss.detach()

    def _blocks(self):
        m = self.model
        for cand in (m, getattr(m, "backbone", None)):
            blocks = getattr(cand, "blocks", None) if cand is not None else None
            if blocks is not None and len(blocks) >= 4 and all(p.requires_grad for p in blocks[0].parameters()):
                return blocks
        return None

    def _capture_split(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Two graphs: (1) forward + loss + backward down to the input of block `split_at`, (2) the rest of backward.
        The gradient arena fills in backward order, so graph 1 owns arena[0:_split_off) -- reduced on the side stream
        while graph 2 replays on grids sized for the SMs NCCL leaves free."""
        from . import dist as vdist
        from . import functional, ops
        blocks = self._blocks()
        k = self.split_at if self.split_at is not None else max(1, len(blocks) // 3)
        self.split_at = k
        stash = {}

        def cut(_mod, args):     # forward pre-hook of block k: cut the autograd graph at its input
            xk = args[0]
            xd = xk.detach().requires_grad_(True)
            stash["xk"], stash["xd"] = xk, xd
            # keep the gradient OBJECT block k's backward returns: it carries the bf16 copy / column sums that the next
            # block's backward reuses (functional.BlockFn side channel); .grad would be a bare alias of it
            xd.register_hook(lambda g: stash.__setitem__("g", g))
            return (xd,) + tuple(args[1:])

        self._sx, self._sy = x.clone(), y.clone()
        torch.cuda.synchronize()
        self._graph, self._graph2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        handle = blocks[k].register_forward_pre_hook(cut)
        try:
            with torch.cuda.graph(self._graph):
                self._sloss = self._fwd_bwd(self._sx, self._sy)       # backward stops at the detached input of block k
            self._split_off = self.arena.off
            limit = vdist.partition_limit(self.reducer.split_ctas)
            ops.set_sm_limit(limit)                                   # graph 2 shares the GPU with the side all-reduce
            try:
                with torch.cuda.graph(self._graph2, pool=self._graph.pool()):
                    stash["xk"].backward(stash.get("g", stash["xd"].grad))
            finally:
                ops.set_sm_limit(0)
        finally:
            handle.remove()
        self.launches_per_step = ops.launch_count - n0 + 1
        self.reducer.discard_pending()
        self._static_grads = [(p, p.grad) for p in self.model.parameters() if p.grad is not None]

    def _capture(self, x: torch.Tensor, y: torch.Tensor) -> None:
        from . import ops
        if self.split and self._blocks() is not None and self.arena is not None:
            return self._capture_split(x, y)
        # (no warm-up step here: it would be an extra optimisation step; step() has already run two eager ones)
        self._sx, self._sy = x.clone(), y.clone()
        if self.opt_in_graph and hasattr(self.opt, "prepare_capture"):
            self.opt.prepare_capture()
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        with torch.cuda.graph(self._graph):
            if self.opt_in_graph:
                self._sloss = self._step_eager(self._sx, self._sy)
            else:
                self._sloss = self._fwd_bwd(self._sx, self._sy)
        self.launches_per_step = ops.launch_count - n0
        if not self.opt_in_graph:
            self.reducer.discard_pending()      # bucket hooks fired during capture; replays reduce the arena directly
            self.launches_per_step += 1         # the optimiser kernel launched after every replay
            # the gradient tensors the captured backward writes: re-attached before every eager tail, because an eager
            # step in between (another batch shape) replaces .grad
            self._static_grads = [(p, p.grad) for p in self.model.parameters() if p.grad is not None]

    def static_inputs(self):
        """(images, labels) device buffers the captured step reads; fill them (e.g. `copy_` from pinned host memory)
        and call step_static(). None before the graph exists."""
        return (self._sx, self._sy) if self._graph is not None else None

    def _finish_static(self) -> None:
        """Data-parallel tail of a captured step: all-reduce the gradients (arena + the rest), then the optimiser."""
        for p, g in self._static_grads:
            p.grad = g
        self.reducer.finish_static(self.arena)
        self.opt.step()

    def step_static(self) -> torch.Tensor:
        if self.opt_in_graph and hasattr(self.opt, "sync_hyper"):
            self.opt.sync_hyper()       # lr schedules act on the captured step through device-resident hyper-parameters
        self._graph.replay()
        if self._graph2 is not None:
            # first half of the gradients: all-reduce on the side stream while the second backward graph runs
            self.reducer.reduce_first_half(self.arena, self._split_off)
            self._graph2.replay()
            for p, g in self._static_grads:
                p.grad = g
            self.reducer.reduce_second_half(self.arena, self._split_off)
            self.opt.step()
            return self._sloss
        if not self.opt_in_graph:
            self._finish_static()
        return self._sloss

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if not self.use_graph:
            return self._step_eager(x, y)
        if self._graph is None:
            if self._eager_steps < 2:        # first steps eager: bf16 weight cache, optimiser state, autotuned plans
                self._eager_steps += 1
                return self._step_eager(x, y)
            try:
                self._capture(x, y)
            except Exception as exc:    # keep training: fall back to eager launches (same kernels, same collectives)
                if self.strict_graph:
                    raise
                import sys
                print(f"[vit_torch_b200] CUDA-graph capture of the step failed ({type(exc).__name__}: {exc}); "
                      "continuing with eager launches", file=sys.stderr)
                self.use_graph = False
                self._graph = None
                torch.cuda.synchronize()
                return self._step_eager(x, y)
        if x.shape != self._sx.shape or y.shape != self._sy.shape or x.dtype != self._sx.dtype:
            return self._step_eager(x, y)    # a different batch shape (e.g. the last batch of an epoch)
        if x.data_ptr() != self._sx.data_ptr():
            self._sx.copy_(x, non_blocking=True)
        if y.data_ptr() != self._sy.data_ptr():
            self._sy.copy_(y, non_blocking=True)
        return self.step_static()
