"""torch.autograd.Functions that sequence the sm_100a kernels for the ViT encoder hot path.

One Function per transformer Block (forward and hand-written backward), one for PatchEmbed + token assembly and one
for the final LayerNorm on the class token. Numerics policy: fp32 residual stream and fp32 master
weights / gradients; bf16 GEMM + attention operands and saved activations; fp32 LayerNorm / softmax / GELU math.

Reference semantics: timm/DINO Block, LayerScale_Block (models/cait.py:147-150),
PatchEmbed + prepare_tokens (models/cait.py:229-234, models/deit.py:35-43), final norm (models/cait.py:242-246).
"""
from __future__ import annotations

import weakref

import torch

from . import ops

# ----------------------------------------------------------------------------------------------------------------
# bf16 weight cache: master weights stay fp32 nn.Parameters; a bf16 copy is refreshed when the parameter changes
# ----------------------------------------------------------------------------------------------------------------
_wcache: dict[int, tuple] = {}


def bf16_weight(p: torch.Tensor) -> torch.Tensor:
    """bf16 copy of a 2-D (or conv 4-D, viewed 2-D) fp32 weight, cached on (identity, version, storage)."""
    key = id(p)
    ent = _wcache.get(key)
    ver, dp = p._version, p.data_ptr()
    if ent is not None and ent[0]() is p and ent[1] == ver and ent[2] == dp:
        return ent[3]
    src = p.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    w = ops.cast_bf16(src.view(src.shape[0], -1))
    _wcache[key] = (weakref.ref(p), ver, dp, w)
    if len(_wcache) > 4096:
        for k in [k for k, v in _wcache.items() if v[0]() is None]:
            del _wcache[k]
    return w


class GradArena:
    """One persistent fp32 buffer that the per-block flat gradient buffers of a step are carved from (opt-in: the
    Trainer owns one). begin() zeroes it with a single memset -- instead of one torch fill per block -- and rewinds it;
    because the blocks' buffers are then adjacent, a data-parallel step can all-reduce them with ONE collective
    (dist.GradAllReducer, deferred mode). Gradients handed out are views: they are valid until the next begin()."""

    def __init__(self, numel: int, device, alloc=None):
        # alloc(numel, device) -> zeroed fp32 tensor or None: lets a data-parallel reducer place the arena in memory
        # registered with its communicator (dist.GradAllReducer.alloc_arena)
        buf = alloc(max(int(numel), 4), device) if alloc is not None else None
        self.buf = buf if buf is not None else torch.zeros((max(int(numel), 4),), dtype=torch.float32, device=device)
        self.off = 0
        self.high = 0            # high-water mark of the last step: only that part needs zeroing / reducing

    def begin(self) -> None:
        self.high = max(self.high, self.off)
        if self.high:
            self.buf[: self.high].zero_()
        self.off = 0

    def take(self, n: int, device):
        if self.buf.device != device or self.off + n > self.buf.numel():
            return None
        out = self.buf[self.off:self.off + n]
        self.off += n
        return out

    def used(self) -> torch.Tensor:
        return self.buf[: self.off]


grad_arena: GradArena | None = None      # set by train.Trainer; None = every block allocates its own zero buffer


def grad_zeros(shape, device):
    """A zeroed fp32 gradient buffer: a slice of the step's gradient arena when there is one (zeroed by its single
    memset, reduced by its single all-reduce), else a fresh tensor."""
    n = 1
    for v in shape:
        n *= int(v)
    buf = grad_arena.take((max(n, 4) + 3) // 4 * 4, device) if grad_arena is not None else None
    if buf is None:
        return torch.zeros(tuple(shape), dtype=torch.float32, device=device)
    return buf[:n].view(tuple(shape))


def _flat_grads(params, needs, device):
    """One contiguous fp32 zero buffer holding the gradients of all params that need one; returns views."""
    sizes = [p.numel() if (p is not None and need) else 0 for p, need in zip(params, needs)]
    total = sum(((s + 3) // 4) * 4 for s in sizes)  # keep every view 16-byte aligned
    buf = grad_arena.take(max(total, 4), device) if grad_arena is not None else None
    if buf is None:
        buf = torch.zeros((max(total, 4),), dtype=torch.float32, device=device)
    views, off = [], 0
    for p, s in zip(params, sizes):
        if s == 0:
            views.append(None)
        else:
            views.append(buf[off:off + s].view(p.shape))
            off += ((s + 3) // 4) * 4
    return buf, views


def _th_s_f32() -> bool:
    """Talking-heads logit planes in fp32 instead of bf16 (VITK_TH_S_F32=1; twice the plane traffic)."""
    import os
    return os.environ.get("VITK_TH_S_F32", "0") == "1"


def _dy_bf16(dy2d, M, D, gamma=None, rowscale=None, rows_per_sample=0):
    """bf16 copy of an fp32 residual-stream gradient (scaled by LayerScale gamma / DropPath), reusing the copy the
    producing kernel already wrote when there is one."""
    side = getattr(dy2d, "_vitk_bf16", None)
    if side is not None and gamma is None and rowscale is None and side.shape == (M, D):
        return side
    return ops.scale_cast(dy2d, M, D, colscale=gamma, rowscale=rowscale, rows_per_sample=rows_per_sample)


class BlockFn(torch.autograd.Function):
    """x -> x + g1 * DropPath(Attn(LN1 x)) -> + g2 * DropPath(Mlp(LN2 .)); g1/g2 = None for plain ViT blocks.
    `rowscale` = None or a pair (rs_attn, rs_mlp) of per-sample DropPath factors mask / keep_prob: the reference calls
    drop_path once per branch (models/cait.py:148-149, timm Block), i.e. two independent draws."""

    @staticmethod
    def forward(ctx, x, num_heads, eps, scale, rowscale, n1w, n1b, qkvw, qkvb, projw, projb, n2w, n2b, fc1w, fc1b,
                fc2w, fc2b, g1, g2, thl_w=None, thl_b=None, thw_w=None, thw_b=None):
        B, N, D = x.shape
        M = B * N
        H = num_heads
        d = D // H
        x0 = x.contiguous().view(M, D)
        dev = x.device
        need_bwd = any(ctx.needs_input_grad)
        wqkv, wproj, wfc1, wfc2 = bf16_weight(qkvw), bf16_weight(projw), bf16_weight(fc1w), bf16_weight(fc2w)
        hidden = fc1w.shape[0]
        rs1, rs2 = rowscale if rowscale is not None else (None, None)

        h1, mean1, rstd1 = ops.layernorm_fwd(x0, n1w, n1b, eps)
        qkv = torch.empty((M, 3 * D), dtype=torch.bfloat16, device=dev)
        ops.gemm(h1, wqkv, epilogue=ops.EPI_STORE_BF16, bias=qkvb, out=qkv)
        th = thl_w is not None
        S = Pm = rmax = rsum = lse2 = None
        if not th:
            o, lse2 = ops.attn_fwd(qkv, B, N, H, d, scale)
        else:
            # talking-heads attention (models/cait.py:111-128): raw logits by a batched tcgen05 GEMM reading q/k in
            # place, one fused mixing/softmax/mixing pass, P'.V by a second batched GEMM
            Np = (N + 7) // 8 * 8
            o = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
            if ops.th_gemm_ok(N, d, Np):
                # short sequences (CaiT at 224 px): whole-image products, one thread block per (image, row block)
                # walking the heads; the logit planes are bf16 as under torch.autocast (VITK_TH_S_F32=1: fp32)
                S = ops.th_scores(qkv, 0, qkv, D, B, N, H, d, Np, out_f32=_th_s_f32())
                Pm, rmax, rsum = ops.th_mix_fwd(S, thl_w, thl_b, thw_w, thw_b, scale, B, H, N, Np)
                ops.th_apply(Pm, qkv, 2 * D, o, 0, B, N, H, d, Np)
            else:
                S = torch.empty((B, H, N, Np), dtype=torch.float32, device=dev)
                ops.gemm_batched(qkv, 3 * D, d, N * 3 * D, False, qkv, 3 * D, d, N * 3 * D, False, N, N, d, H, B, S, Np,
                                 N * Np, H * N * Np, a_off=0, b_off=D, out_f32=True)
                Pm, rmax, rsum = ops.th_mix_fwd(S, thl_w, thl_b, thw_w, thw_b, scale, B, H, N, Np)
                ops.gemm_batched(Pm, Np, N * Np, H * N * Np, False, qkv, 3 * D, d, N * 3 * D, True, N, d, N, H, B, o, D,
                                 d, N * D, b_off=2 * D)
            if not need_bwd:
                S = Pm = None
        x1 = torch.empty((M, D), dtype=torch.float32, device=dev)
        f1 = torch.empty((M, D), dtype=torch.bfloat16, device=dev) if (g1 is not None and need_bwd) else None
        ops.gemm(o, wproj, epilogue=ops.EPI_RESID_F32, bias=projb, gamma=g1, resid=x0, out=x1, out2=f1,
                 rowscale=rs1, rows_per_sample=N)
        h2, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b, eps)
        a = torch.empty((M, hidden), dtype=torch.bfloat16, device=dev) if need_bwd else None
        g = torch.empty((M, hidden), dtype=torch.bfloat16, device=dev)
        ops.gemm(h2, wfc1, epilogue=ops.EPI_BIAS_GELU, bias=fc1b, out=a, out2=g)
        x2 = torch.empty((M, D), dtype=torch.float32, device=dev)
        f2 = torch.empty((M, D), dtype=torch.bfloat16, device=dev) if (g2 is not None and need_bwd) else None
        ops.gemm(g, wfc2, epilogue=ops.EPI_RESID_F32, bias=fc2b, gamma=g2, resid=x1, out=x2, out2=f2,
                 rowscale=rs2, rows_per_sample=N)

        if need_bwd:
            ctx.save_for_backward(x0, mean1, rstd1, h1, qkv, o, lse2, x1, mean2, rstd2, h2, a, g, f1, f2, n1w, qkvw,
                                  projw, n2w, fc1w, fc2w, g1, g2, rs1, rs2, wqkv, wproj, wfc1, wfc2, S, Pm, rmax, rsum,
                                  thl_w, thl_b, thw_w, thw_b)
            ctx.dims = (B, N, D, H, d, scale)
            ctx.has = (qkvb is not None, projb is not None, fc1b is not None, fc2b is not None)
            ctx.param_refs = [n1w, n1b, qkvw, qkvb, projw, projb, n2w, n2b, fc1w, fc1b, fc2w, fc2b, g1, g2, thl_w,
                              thl_b, thw_w, thw_b]
        return x2.view(B, N, D)

    @staticmethod
    def backward(ctx, dout):
        (x0, mean1, rstd1, h1, qkv, o, lse2, x1, mean2, rstd2, h2, a, g, f1, f2, n1w, qkvw, projw, n2w, fc1w, fc2w,
         g1, g2, rs1, rs2, wqkv, wproj, wfc1, wfc2, S, Pm, rmax, rsum, thl_w, thl_b, thw_w, thw_b) = ctx.saved_tensors
        B, N, D, H, d, scale = ctx.dims
        M = B * N
        dev = dout.device
        hidden = fc1w.shape[0]
        needs = ctx.needs_input_grad
        # parameter order of forward(): index 5.. = n1w n1b qkvw qkvb projw projb n2w n2b fc1w fc1b fc2w fc2b g1 g2
        shapes_like = [n1w, n1w, qkvw, qkvw[:, 0] if ctx.has[0] else None, projw, projw[:, 0] if ctx.has[1] else None,
                       n2w, n2w, fc1w, fc1w[:, 0] if ctx.has[2] else None, fc2w, fc2w[:, 0] if ctx.has[3] else None,
                       g1, g2, thl_w, thl_b, thw_w, thw_b]
        buf, gv = _flat_grads(shapes_like, needs[5:23], dev)
        (dn1w, dn1b, dqkvw, dqkvb, dprojw, dprojb, dn2w, dn2b, dfc1w, dfc1b, dfc2w, dfc2b, dg1, dg2, dthl_w, dthl_b,
         dthw_w, dthw_b) = gv

        dy = dout if dout.is_contiguous() else dout.contiguous()
        dx2 = dy.view(M, D)
        if getattr(dout, "_vitk_bf16", None) is not None:
            dx2._vitk_bf16 = dout._vitk_bf16
            dx2._vitk_colsum = getattr(dout, "_vitk_colsum", None)

        # ---- Mlp branch: x2 = x1 + g2 * rowscale * (fc2(gelu(fc1(h2))))
        fc2b_done = False
        if g2 is not None and f2 is not None:
            # LayerScale: bf16(dY * gamma * rowscale), d_gamma = sum dY * rowscale * f and the fc2 bias gradient in ONE pass
            dx2b = ops.layerscale_bwd(dx2, f2, g2, rs2, N, dgamma=dg2, dbias=dfc2b)
            fc2b_done = True
        else:
            dx2b = _dy_bf16(dx2, M, D, g2, rs2, N)
        da = torch.empty((M, hidden), dtype=torch.bfloat16, device=dev)
        # fc2 dgrad with GELU' fused; the same epilogue reduces da over rows = fc1 bias gradient
        ops.gemm(dx2b, wfc2, b_mn=True, epilogue=ops.EPI_DGELU, aux=a, out=da, colsum=dfc1b)
        if dfc2w is not None:
            ops.gemm(dx2b, g, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dfc2w)
        if dfc2b is not None and not fc2b_done:
            side = getattr(dx2, "_vitk_colsum", None) if dx2b is getattr(dx2, "_vitk_bf16", None) else None
            if side is not None:
                dfc2b.add_(side)        # column sums of dx2b were produced by the kernel that wrote dx2b
            else:
                ops.colsum_accum(dx2b, dfc2b)
        dh2 = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
        ops.gemm(da, wfc1, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=dh2)
        if dfc1w is not None:
            ops.gemm(da, h2, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dfc1w)
        del da
        # ---- LN2 backward + residual: dx1 = dx2 + LN2'(dh2); bf16 copy (x g1) feeds the proj dgrad/wgrad
        plain1 = g1 is None and rs1 is None
        dx1, dx1b = ops.layernorm_bwd(dh2, x1, n2w, mean2, rstd2, dres=dx2, dweight=dn2w, dbias=dn2b, want_bf16=plain1,
                                      dxsum=dprojb if plain1 else None)
        projb_done = plain1
        if g1 is not None and f1 is not None:
            dx1b = ops.layerscale_bwd(dx1, f1, g1, rs1, N, dgamma=dg1, dbias=dprojb)
            projb_done = True
        elif not plain1:
            dx1b = ops.scale_cast(dx1, M, D, colscale=g1, rowscale=rs1, rows_per_sample=N)
        # ---- attention branch
        do = dh2  # reuse the buffer: dO [M, D] bf16
        ops.gemm(dx1b, wproj, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=do)
        if dprojw is not None:
            ops.gemm(dx1b, o, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dprojw)
        if dprojb is not None and not projb_done:
            ops.colsum_accum(dx1b, dprojb)
        attn_dbias_done = False
        if thl_w is None:
            dqkv = ops.attn_bwd(qkv, o, do, lse2, B, N, H, d, scale, dbias=dqkvb)  # qkv bias gradient fused in
            attn_dbias_done = dqkvb is not None
        else:
            Np = S.shape[-1]
            zeros = lambda t: torch.zeros_like(t, dtype=torch.float32)
            dthl_w = dthl_w if dthl_w is not None else zeros(thl_w)
            dthl_b = dthl_b if dthl_b is not None else zeros(thl_b)
            dthw_w = dthw_w if dthw_w is not None else zeros(thw_w)
            dthw_b = dthw_b if dthw_b is not None else zeros(thw_b)
            dqkv = torch.empty((M, 3 * D), dtype=torch.bfloat16, device=dev)
            if ops.th_gemm_ok(N, d, Np) and (S.dtype == torch.bfloat16 or ops.th_mix_bf16_dp(Np)):
                # (the three products also reduce their outputs over rows: the qkv bias gradient, no separate pass)
                cs = (lambda k: dqkvb[k * D:(k + 1) * D]) if dqkvb is not None else (lambda k: None)
                dPm = ops.th_scores(do, 0, qkv, 2 * D, B, N, H, d, Np)                  # dP'[i,j] = dO_i . v_j (bf16)
                ops.th_apply_t(Pm, do, 0, dqkv, 2 * D, B, N, H, d, Np, colsum=cs(2))    # dV = P'^T dO
                dS = ops.th_mix_bwd(S, dPm, rmax, rsum, thl_w, thl_b, thw_w, thw_b, scale, dthl_w, dthl_b, dthw_w,
                                    dthw_b, B, H, N, Np)
                del dPm
                ops.th_apply(dS, qkv, D, dqkv, 0, B, N, H, d, Np, colsum=cs(0))         # dQ = dS K
                ops.th_apply_t(dS, qkv, 0, dqkv, D, B, N, H, d, Np, colsum=cs(1))       # dK = dS^T Q
                del dS
                attn_dbias_done = dqkvb is not None
            else:
                dp16 = ops.th_mix_bf16_dp(Np)       # version-2 mixing kernels read dP' in bf16
                dPm = torch.empty((B, H, N, Np), dtype=torch.bfloat16 if dp16 else torch.float32, device=dev)
                ops.gemm_batched(do, D, d, N * D, False, qkv, 3 * D, d, N * 3 * D, False, N, N, d, H, B, dPm, Np, N * Np,
                                 H * N * Np, b_off=2 * D, out_f32=not dp16)     # dP'[i,j] = dO_i . v_j
                ops.gemm_batched(Pm, Np, N * Np, H * N * Np, True, do, D, d, N * D, True, N, d, N, H, B, dqkv, 3 * D, d,
                                 N * 3 * D, out_off=2 * D)                              # dV = P'^T dO
                dS = ops.th_mix_bwd(S, dPm, rmax, rsum, thl_w, thl_b, thw_w, thw_b, scale, dthl_w, dthl_b, dthw_w,
                                    dthw_b, B, H, N, Np)
                del dPm
                ops.gemm_batched(dS, Np, N * Np, H * N * Np, False, qkv, 3 * D, d, N * 3 * D, True, N, d, N, H, B, dqkv,
                                 3 * D, d, N * 3 * D, b_off=D, out_off=0)               # dQ = dS K
                ops.gemm_batched(dS, Np, N * Np, H * N * Np, True, qkv, 3 * D, d, N * 3 * D, True, N, d, N, H, B, dqkv,
                                 3 * D, d, N * 3 * D, b_off=0, out_off=D)               # dK = dS^T Q
                del dS
            if not needs[19]: dthl_w = None
            if not needs[20]: dthl_b = None
            if not needs[21]: dthw_w = None
            if not needs[22]: dthw_b = None
        dh1 = do
        ops.gemm(dqkv, wqkv, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=dh1)
        if dqkvw is not None:
            ops.gemm(dqkv, h1, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dqkvw)
        if dqkvb is not None and not attn_dbias_done:
            ops.colsum_accum(dqkv, dqkvb)
        # ---- LN1 backward + residual; the bf16 copy is handed to the previous block through a side channel
        side_sum = grad_zeros((D,), dev)       # (arena slice: zeroed by the step's single memset)
        dx0, dx0b = ops.layernorm_bwd(dh1, x0, n1w, mean1, rstd1, dres=dx1, dweight=dn1w, dbias=dn1b, want_bf16=True,
                                      dxsum=side_sum)
        dx = dx0.view(B, N, D)
        dx._vitk_bf16 = dx0b
        dx._vitk_colsum = side_sum
        if grad_bucket_hooks:
            # hand the hooks their own alias views so the returned ones stay uniquely referenced (autograd then adopts
            # them as .grad without cloning)
            alias = [None if v is None else v.view(v.shape) for v in gv]
            for hook in grad_bucket_hooks:
                hook(buf, ctx.param_refs, alias)
        return (dx, None, None, None, None, dn1w, dn1b, dqkvw, dqkvb, dprojw, dprojb, dn2w, dn2b, dfc1w, dfc1b,
                dfc2w, dfc2b, dg1, dg2, dthl_w, dthl_b, dthw_w, dthw_b)


# callbacks(flat_grad_buffer, params, grad_views) fired when one block's parameter gradients are complete (used by dist.py)
grad_bucket_hooks: list = []


def patch_embed_tma_ok(dtype, C, H, W, P) -> bool:
    """Geometry the im2col-free PatchEmbed kernels take (patch_embed.cu): a patch row of P pixels must be 16 / 32 / 64 /
    128 bytes (fp32: P = 4..32, bf16: P = 8..64) and one row of patches must fit a 64-row pipeline stage."""
    import os
    if os.environ.get("VITK_PATCH_EMBED", "tma") != "tma":
        return False
    esz = 4 if dtype == torch.float32 else 2
    rb = P * esz
    return (rb in (16, 32, 64, 128) and W // P <= 64 and H // P <= 256 and (C * P) % (128 // rb) == 0
            and (W * esz) % 16 == 0 and (H * W * esz) % 16 == 0)


class TokensFn(torch.autograd.Function):
    """PatchEmbed (Conv2d(C,D,P,P) as an im2col-free patch GEMM: TMA gathers the patches straight from NCHW) + prefix
    tokens + positional embedding -> [B, N, D] fp32. img: fp32 (tf32 tensor-core path on the fp32 pixels and master
    weight), bf16, or uint8 with `norm` = (mean[C], std[C]) device vectors (ToTensor + Normalize run on the device,
    utils_datasets.py:573-580). Geometries the TMA path does not take fall back to an explicit patch matrix."""

    @staticmethod
    def forward(ctx, img, conv_w, conv_b, pos, prefix, patch, norm=None):
        B, C, Hh, Ww = img.shape
        D = conv_w.shape[0]
        P = patch
        n = (Hh // P) * (Ww // P)
        T = 0 if prefix is None else prefix.shape[-2]
        N = n + T
        assert pos.shape[-2] == N and pos.shape[-1] == D, f"pos_embed {tuple(pos.shape)} vs tokens {N}x{D}"
        if img.dtype == torch.uint8:
            if norm is None:
                raise ValueError("uint8 images need the (mean, std) of the dataset's Normalize: "
                                 "PatchEmbed.set_input_normalization(mean, std)")
            x = ops.normalize_u8(img.contiguous(), norm[0], norm[1])
        elif img.dtype in (torch.float32, torch.bfloat16):
            x = img.contiguous()
        else:
            x = img.float().contiguous()
        pos2d = pos.detach().reshape(N, D).contiguous()
        out = torch.empty((B, N, D), dtype=torch.float32, device=img.device)
        tma = patch_embed_tma_ok(x.dtype, C, Hh, Ww, P)
        patches = None
        if tma:
            w2d = conv_w.detach().reshape(D, -1) if x.dtype == torch.float32 else bf16_weight(conv_w)
            ops.patch_embed_fwd(x, w2d.contiguous(), conv_b, pos2d, out, P, T)
        else:
            patches = ops.patchify(x if x.dtype == torch.float32 else x.float(), P)
            ops.gemm(patches, bf16_weight(conv_w), epilogue=ops.EPI_TOKENS_F32, bias=conv_b, resid=pos2d,
                     out=out.view(B * N, D), tok=(n, N, T))
        if T:
            ops.prefix_tokens(prefix.detach().reshape(T, D).contiguous(), pos2d, out, B, T, N, D)
        ctx.save_for_backward(x if tma else patches, conv_w, pos, prefix)
        ctx.dims = (B, n, N, T, D, conv_b is not None, P, tma)
        return out

    @staticmethod
    def backward(ctx, dout):
        src, conv_w, pos, prefix = ctx.saved_tensors
        B, n, N, T, D, has_bias, P, tma = ctx.dims
        needs = ctx.needs_input_grad
        dy = dout if dout.is_contiguous() else dout.contiguous()
        dev = dy.device
        dw = db = dpos = dprefix = None
        want_db = needs[2] and has_bias
        if needs[3] or (needs[4] and T) or (want_db and tma):
            dp = grad_zeros((N * D,), dev)                                      # sum over the batch: d_pos
            ops.colsum_f32_accum(dy, N * D, B, N * D, dp)
            if needs[3]:
                dpos = dp.view(pos.shape)
            if needs[4] and T:
                dprefix = grad_zeros((T * D,), dev).copy_(dp[:T * D]).view(prefix.shape)
            if want_db and tma:                                                # bias gradient = d_pos summed over patches
                db = grad_zeros((D,), dev)
                ops.colsum_f32_accum(dp[T * D:], D, n, D, db)
        if tma:
            if needs[1]:
                dw = grad_zeros(conv_w.shape, dev)
                # the weight gradient runs in bf16 (tcgen05 has no MN-major tf32 form for 64-byte rows): an fp32
                # image is cast once (NCHW -> NCHW, not a patch matrix); the bf16 token gradient usually exists already
                img16 = src if src.dtype == torch.bfloat16 else ops.cast_bf16(src)
                side = getattr(dout, "_vitk_bf16", None)
                dyt = side if side is not None and side.numel() == dy.numel() else ops.scale_cast(dy, B * N, D)
                ops.patch_embed_wgrad(img16, dyt.view(B, N, D), dw.view(D, -1), P, T)
        elif needs[1] or want_db:
            dyb = ops.scale_cast(dy, B * n, D, rows_per_group=n, group_stride=N * D, offset_elems=T * D)
            if needs[1]:
                dw = grad_zeros(conv_w.shape, dev)
                ops.gemm(dyb, src, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dw.view(D, -1))
            if want_db:
                db = grad_zeros((D,), dev)
                ops.colsum_accum(dyb, db)
        return None, dw, db, dpos, dprefix, None, None


class TokenNormFn(torch.autograd.Function):
    """LayerNorm of one token (index `tok`) of every image: x [B, N, D] fp32 -> [B, D] fp32 (norm(x)[:, tok])."""

    @staticmethod
    def forward(ctx, x, w, b, eps, tok):
        B, N, D = x.shape
        xc = x if x.is_contiguous() else x.contiguous()
        xv = xc.view(-1)[tok * D:]
        y, mean, rstd = ops.layernorm_fwd_rows(xv, N * D, B, D, w, b, eps, out_f32=True)
        ctx.save_for_backward(xc, w, mean, rstd)
        ctx.dims = (B, N, D, tok)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w, mean, rstd = ctx.saved_tensors
        B, N, D, tok = ctx.dims
        dy = dy.contiguous().float()
        dx = torch.zeros((B, N, D), dtype=torch.float32, device=dy.device)
        dw = grad_zeros((D,), dy.device)
        db = grad_zeros((D,), dy.device)
        ops.layernorm_bwd_rows(dy, xc.view(-1)[tok * D:], N * D, B, D, w, mean, rstd, dx=dx.view(-1)[tok * D:],
                               dx_stride=N * D, dweight=dw, dbias=db)
        return dx, dw, db, None, None


class ClassAttnBlockFn(torch.autograd.Function):
    """CaiT LayerScale_Block_CA (models/cait.py:75-84):
         u = cat(cls, x); cls += g1 * ClassAttn(LN1 u); cls += g2 * Mlp(LN2 cls)
    x [B, n, C] fp32 (patch tokens, unchanged), cls [B, 1, C] fp32 -> new cls [B, 1, C] fp32.
    No concatenation is materialised: LN1 runs on the patch rows and on the class rows separately and the class-
    attention kernel takes the class row and the patch rows of K / V through two pointers. K and V come from one
    tcgen05 GEMM against the stacked [Wk; Wv] weight (one dgrad, one wgrad)."""

    @staticmethod
    def forward(ctx, x, cls, num_heads, eps, scale, n1w, n1b, qw, qb, kw, kb, vw, vb, projw, projb, n2w, n2b, fc1w,
                fc1b, fc2w, fc2b, g1, g2):
        B, n, C = x.shape
        H = num_heads
        d = C // H
        dev = x.device
        xr = x.contiguous().view(B * n, C)
        c0 = cls.contiguous().view(B, C)
        need_bwd = any(ctx.needs_input_grad)
        wq, wp, w1, w2 = bf16_weight(qw), bf16_weight(projw), bf16_weight(fc1w), bf16_weight(fc2w)
        wkv = torch.empty((2 * C, C), dtype=torch.bfloat16, device=dev)
        ops.cast_bf16(kw.detach(), out=wkv[:C])
        ops.cast_bf16(vw.detach(), out=wkv[C:])
        bkv = None
        if kb is not None:
            bkv = torch.cat((kb.detach(), vb.detach()))
        hidden = fc1w.shape[0]

        hx, mx, rx = ops.layernorm_fwd(xr, n1w, n1b, eps)
        hc, mc, rc = ops.layernorm_fwd(c0, n1w, n1b, eps)
        q = torch.empty((B, C), dtype=torch.bfloat16, device=dev)
        ops.gemm(hc, wq, epilogue=ops.EPI_STORE_BF16, bias=qb, out=q)
        kvx = torch.empty((B * n, 2 * C), dtype=torch.bfloat16, device=dev)
        ops.gemm(hx, wkv, epilogue=ops.EPI_STORE_BF16, bias=bkv, out=kvx)
        kvc = torch.empty((B, 2 * C), dtype=torch.bfloat16, device=dev)
        ops.gemm(hc, wkv, epilogue=ops.EPI_STORE_BF16, bias=bkv, out=kvc)
        kc, vc = kvc[:, :C], kvc[:, C:]   # strided views (pitch 2C), read in place
        a, p = ops.class_attn_fwd(q, kc, kvx[:, :C], vc, kvx[:, C:], 2 * C, 2 * C, scale, B, H, n, d)
        c1 = torch.empty((B, C), dtype=torch.float32, device=dev)
        f1 = torch.empty((B, C), dtype=torch.bfloat16, device=dev) if need_bwd else None
        ops.gemm(a, wp, epilogue=ops.EPI_RESID_F32, bias=projb, gamma=g1, resid=c0, out=c1, out2=f1)
        h2, m2, r2 = ops.layernorm_fwd(c1, n2w, n2b, eps)
        pre = torch.empty((B, hidden), dtype=torch.bfloat16, device=dev) if need_bwd else None
        act = torch.empty((B, hidden), dtype=torch.bfloat16, device=dev)
        ops.gemm(h2, w1, epilogue=ops.EPI_BIAS_GELU, bias=fc1b, out=pre, out2=act)
        c2 = torch.empty((B, C), dtype=torch.float32, device=dev)
        f2 = torch.empty((B, C), dtype=torch.bfloat16, device=dev) if need_bwd else None
        ops.gemm(act, w2, epilogue=ops.EPI_RESID_F32, bias=fc2b, gamma=g2, resid=c1, out=c2, out2=f2)
        if need_bwd:
            ctx.save_for_backward(xr, c0, hx, mx, rx, hc, mc, rc, q, kvx, kvc, a, p, c1, f1, h2, m2, r2, pre, act, f2,
                                  n1w, qw, kw, vw, projw, n2w, fc1w, fc2w, g1, g2, wq, wkv, wp, w1, w2)
            ctx.dims = (B, n, C, H, d, scale)
            ctx.has = (qb is not None, kb is not None, projb is not None, fc1b is not None, fc2b is not None)
            ctx.param_refs = [n1w, n1b, qw, qb, kw, kb, vw, vb, projw, projb, n2w, n2b, fc1w, fc1b, fc2w, fc2b, g1, g2]
        return c2.view(B, 1, C)

    @staticmethod
    def backward(ctx, dout):
        (xr, c0, hx, mx, rx, hc, mc, rc, q, kvx, kvc, a, p, c1, f1, h2, m2, r2, pre, act, f2, n1w, qw, kw, vw, projw,
         n2w, fc1w, fc2w, g1, g2, wq, wkv, wp, w1, w2) = ctx.saved_tensors
        B, n, C, H, d, scale = ctx.dims
        dev = dout.device
        hidden = fc1w.shape[0]
        needs = ctx.needs_input_grad
        has_qb, has_kb, has_pb, has_b1, has_b2 = ctx.has
        vec = n1w
        like = [n1w, n1w, qw, vec if has_qb else None, kw, vec if has_kb else None, vw, vec if has_kb else None, projw,
                vec if has_pb else None, n2w, n2w, fc1w, fc1w[:, 0] if has_b1 else None, fc2w, vec if has_b2 else None,
                g1, g2]
        buf, gv = _flat_grads(like, needs[5:23], dev)
        (dn1w, dn1b, dqw, dqb, dkw, dkb, dvw, dvb, dprojw, dprojb, dn2w, dn2b, dfc1w, dfc1b, dfc2w, dfc2b, dg1, dg2) = gv
        dc2 = dout.contiguous().view(B, C).float()
        # ---- Mlp on the class token
        if dg2 is not None:
            ops.colsum_prod_accum(dc2, f2, dg2)
        dc2b = ops.scale_cast(dc2, B, C, colscale=g2)
        dact = torch.empty((B, hidden), dtype=torch.bfloat16, device=dev)
        ops.gemm(dc2b, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=pre, out=dact)
        if dfc2w is not None:
            ops.gemm(dc2b, act, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dfc2w)
        if dfc2b is not None:
            ops.colsum_accum(dc2b, dfc2b)
        dh2 = torch.empty((B, C), dtype=torch.bfloat16, device=dev)
        ops.gemm(dact, w1, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=dh2)
        if dfc1w is not None:
            ops.gemm(dact, h2, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dfc1w)
        if dfc1b is not None:
            ops.colsum_accum(dact, dfc1b)
        dc1, _ = ops.layernorm_bwd(dh2, c1, n2w, m2, r2, dres=dc2, dweight=dn2w, dbias=dn2b)
        # ---- class attention
        if dg1 is not None:
            ops.colsum_prod_accum(dc1, f1, dg1)
        dc1b = ops.scale_cast(dc1, B, C, colscale=g1)
        da = torch.empty((B, C), dtype=torch.bfloat16, device=dev)
        ops.gemm(dc1b, wp, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=da)
        if dprojw is not None:
            ops.gemm(dc1b, a, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dprojw)
        if dprojb is not None:
            ops.colsum_accum(dc1b, dprojb)
        dq = torch.empty((B, C), dtype=torch.bfloat16, device=dev)
        dkvx = torch.empty((B * n, 2 * C), dtype=torch.bfloat16, device=dev)
        dkvc = torch.empty((B, 2 * C), dtype=torch.bfloat16, device=dev)
        ops.class_attn_bwd(q, kvc[:, :C], kvx[:, :C], kvc[:, C:], kvx[:, C:], 2 * C, 2 * C, p, da, scale, dq,
                           dkvc[:, :C], dkvx[:, :C], dkvc[:, C:], dkvx[:, C:], 2 * C, 2 * C, B, H, n, d)
        # ---- q / k / v Linears: dgrad to the LN1 outputs, wgrad into the stacked [dWk; dWv] view of the flat buffer
        dhx = torch.empty((B * n, C), dtype=torch.bfloat16, device=dev)
        ops.gemm(dkvx, wkv, b_mn=True, epilogue=ops.EPI_STORE_BF16, out=dhx)
        dhc32 = torch.empty((B, C), dtype=torch.float32, device=dev)
        ops.gemm(dkvc, wkv, b_mn=True, epilogue=ops.EPI_STORE_F32, out=dhc32)
        dhc32b = torch.empty((B, C), dtype=torch.float32, device=dev)
        ops.gemm(dq, wq, b_mn=True, epilogue=ops.EPI_RESID_F32, resid=dhc32, out=dhc32b)
        for dw, dy_sl, in ((dkw, slice(0, C)), (dvw, slice(C, 2 * C))):
            if dw is not None:
                ops.gemm(dkvx[:, dy_sl], hx, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dw)
                ops.gemm(dkvc[:, dy_sl], hc, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dw)
        for db, dy_sl in ((dkb, slice(0, C)), (dvb, slice(C, 2 * C))):
            if db is not None:
                ops.colsum_accum(dkvx[:, dy_sl], db)
                ops.colsum_accum(dkvc[:, dy_sl], db)
        if dqw is not None:
            ops.gemm(dq, hc, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, out=dqw)
        if dqb is not None:
            ops.colsum_accum(dq, dqb)
        # ---- LN1 backward: patch rows (gradient of x) and class rows (+ residual path of cls)
        dx, _ = ops.layernorm_bwd(dhx, xr, n1w, mx, rx, dweight=dn1w, dbias=dn1b)
        dcls = torch.empty((B, C), dtype=torch.float32, device=dev)
        ops.layernorm_bwd_rows(dhc32b, c0, C, B, C, n1w, mc, rc, dres=dc1, dx=dcls, dx_stride=C, dweight=dn1w,
                               dbias=dn1b)
        if grad_bucket_hooks:
            alias = [None if v is None else v.view(v.shape) for v in gv]
            for hook in grad_bucket_hooks:
                hook(buf, ctx.param_refs, alias)
        return (dx.view(B, n, C), dcls.view(B, 1, C), None, None, None, dn1w, dn1b, dqw, dqb, dkw, dkb, dvw, dvb, dprojw,
                dprojb, dn2w, dn2b, dfc1w, dfc1b, dfc2w, dfc2b, dg1, dg2)


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class HeadFn(torch.autograd.Function):
    """The fc classifier head of models/vision_all.py:299-320: Linear(+bias)+GELU ... Linear(bias=False) on [B, F]
    features, as a chain of tcgen05 GEMMs (bias+GELU fused in the epilogue; dgrad with GELU' fused; split-K wgrad).
    args: x [B, F] fp32, n_layers, then (weight, bias-or-None) per layer; GELU follows every layer but the last."""

    @staticmethod
    def forward(ctx, x, n_layers, *wb):
        B, F0 = x.shape
        dev = x.device
        ws, bs = wb[0::2], wb[1::2]
        need_bwd = any(ctx.needs_input_grad)
        h = ops.scale_cast(x.contiguous(), B, F0) if F0 % 8 == 0 else None
        if h is None:
            raise NotImplementedError("head input width must be a multiple of 8")
        saved_in, saved_pre = [h], []
        out = None
        for i in range(n_layers):
            w = bf16_weight(ws[i])
            n_out, k_in = ws[i].shape
            if k_in % 8:
                raise NotImplementedError("head layer widths (except the last) must be multiples of 8")
            last = i == n_layers - 1
            if last:
                out = torch.zeros((B, _pad8(n_out)), dtype=torch.float32, device=dev)
                ops.gemm(h, w, N=n_out, epilogue=ops.EPI_STORE_F32, bias=bs[i], out=out)
            else:
                if n_out % 8:
                    raise NotImplementedError("head layer widths (except the last) must be multiples of 8")
                pre = torch.empty((B, n_out), dtype=torch.bfloat16, device=dev) if need_bwd else None
                act = torch.empty((B, n_out), dtype=torch.bfloat16, device=dev)
                ops.gemm(h, w, epilogue=ops.EPI_BIAS_GELU, bias=bs[i], out=pre, out2=act)
                saved_pre.append(pre)
                saved_in.append(act)
                h = act
        n_last = ws[-1].shape[0]
        if need_bwd:
            ctx.save_for_backward(*saved_in, *saved_pre, *ws)
            ctx.meta = (B, n_layers, [b is not None for b in bs], n_last)
        return out[:, :n_last]

    @staticmethod
    def backward(ctx, dout):
        B, L, has_b, n_last = ctx.meta
        t = ctx.saved_tensors
        ins, pres, ws = t[:L], t[L:2 * L - 1], t[2 * L - 1:]
        dev = dout.device
        needs = ctx.needs_input_grad
        grads = [None] * (2 * L)
        # dY of the last layer as a bf16 matrix with a 16-byte aligned pitch
        dy = torch.zeros((B, _pad8(n_last)), dtype=torch.float32, device=dev)
        dy[:, :n_last] = dout
        dyb = ops.scale_cast(dy, B, _pad8(n_last))
        n_cur = n_last
        for i in range(L - 1, -1, -1):
            w = bf16_weight(ws[i])
            n_out, k_in = ws[i].shape
            dyv = dyb[:, :n_out] if dyb.shape[1] != n_out else dyb
            if needs[2 + 2 * i]:
                dw = grad_zeros((_pad8(n_out), k_in), dev)
                ops.gemm(dyv, ins[i], a_mn=True, b_mn=True, M=n_out, epilogue=ops.EPI_ATOMIC_F32, out=dw)
                grads[2 * i] = dw[:n_out]
            if has_b[i] and needs[3 + 2 * i]:
                db = grad_zeros((_pad8(n_out),), dev)
                ops.colsum_accum(dyb, db)
                grads[2 * i + 1] = db[:n_out]
            if i > 0:
                nxt = torch.empty((B, k_in), dtype=torch.bfloat16, device=dev)
                ops.gemm(dyv, w, b_mn=True, K=n_out, epilogue=ops.EPI_DGELU, aux=pres[i - 1], out=nxt)
                dyb = nxt
            elif needs[0]:
                dx = torch.empty((B, k_in), dtype=torch.float32, device=dev)
                ops.gemm(dyv, w, b_mn=True, K=n_out, epilogue=ops.EPI_STORE_F32, out=dx)
                return (dx, None, *grads)
        return (None, None, *grads)


class CrossEntropyFn(torch.autograd.Function):
    """nn.CrossEntropyLoss() (mean) of utils_network.py:429-433 as one kernel that also produces the gradient and the
    argmax-accuracy count of utils_network.py:85-95. Returns (loss, n_correct) as 0-d device tensors (no host sync);
    n_correct carries no gradient."""

    @staticmethod
    def forward(ctx, logits, labels):
        out2, dl = ops.cross_entropy(logits, labels, want_grad=ctx.needs_input_grad[0])
        if dl is not None:
            ctx.save_for_backward(dl)
        ncorrect = out2[1]
        ctx.mark_non_differentiable(ncorrect)
        return out2[0], ncorrect

    @staticmethod
    def backward(ctx, dloss, _dcorrect):
        (dl,) = ctx.saved_tensors
        return dl * dloss, None


def cross_entropy(logits, labels):
    """(mean cross-entropy, number of correct argmax predictions): fused drop-in for the reference's loss + accuracy."""
    return CrossEntropyFn.apply(logits, labels)
