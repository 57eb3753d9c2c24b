/*
 * vitk.h -- C ABI of libvitk.so: hand-written sm_100a (B200) kernels for the ViT encoder hot path of
 * khuongnd6/ViT_torch (DINO ViT-S/B 16/8, DeiT, CaiT backbones).
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (torch allocates them); the library never allocates
 *     device memory and keeps no global state except a host-side TMA tensor-map cache guarded by a mutex.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and returns 0 (VITK_OK)
 *     or a negative error code. Nothing throws.
 *   - bf16 matrices are row-major with an explicit leading dimension in ELEMENTS (multiple of 8); fp32 vectors are
 *     dense. "rows" is always B*N tokens flattened.
 *   - The reference has no FFI of its own (pure PyTorch): each entry cites the reference Python it replaces
 *     (paths under the reference tree) and INTEGRATION.md shows the ctypes binding a maintainer would add.
 */
#ifndef VITK_H_
#define VITK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define VITK_OK 0
#define VITK_ERR_ARG (-1)         /* bad shape / alignment / null pointer */
#define VITK_ERR_UNSUPPORTED (-2) /* combination not compiled in */
#define VITK_ERR_CUDA (-3)        /* launch failed (cudaGetLastError) */
#define VITK_ERR_TMAP (-4)        /* cuTensorMapEncodeTiled failed */

/* GEMM epilogues (vitk_gemm_bf16 `epilogue`) */
#define VITK_EPI_STORE_BF16 0 /* out_bf16 = acc (+bias)                                              */
#define VITK_EPI_BIAS_GELU 1  /* pre = acc+bias; out_bf16 = gelu'(pre) [opt]; out2_bf16 = gelu(pre) (fc1+act) */
#define VITK_EPI_RESID_F32 2  /* v = acc+bias; [out2_bf16 = v]; out_f32 = resid + gamma*v (proj / fc2)  */
#define VITK_EPI_DGELU 3      /* out_bf16 = acc * aux_bf16, aux = gelu'(pre) from BIAS_GELU (fc2 dgrad)   */
#define VITK_EPI_ATOMIC_F32 4 /* out_f32 += acc, split-K                                  (wgrad)       */
#define VITK_EPI_STORE_F32 5  /* out_f32 = acc (+bias)                                                  */
#define VITK_EPI_TOKENS_F32 6 /* PatchEmbed token assembly: row (b,p) -> out_f32[b*tok_N+tok_T+p] = acc+bias+pos[tok_T+p] */

/* Library/ABI version and build info. */
int vitk_abi_version(void);

/* Size the persistent kernels' grids for at most n SMs (0 = all). Data-parallel runs leave a few SMs to the NCCL
 * kernels that overlap backward (vit_torch_b200/dist.py). Process-wide. */
int vitk_set_sm_limit(int n);

/*
 * D[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM), fused epilogue.
 *   a_mn_major = 0: A stored [M, lda] (K contiguous);  1: A stored [K, lda] (M contiguous)
 *   b_mn_major = 0: B stored [N, ldb] (K contiguous);  1: B stored [K, ldb] (N contiguous)
 * Replaces nn.Linear forward (models/cait.py:99,102,113,126; timm/DINO Attention.qkv/proj, Mlp.fc1/fc2 --
 * in-repo witness models/swin.py:24-30) and the autograd-generated dgrad (dX = dY W: A=dY, B=W mn-major) and
 * wgrad (dW = dY^T X: A=dY mn-major, B=X mn-major, VITK_EPI_ATOMIC_F32 accumulating into the fp32 .grad).
 * N % 8 == 0, lda/ldb % 8 == 0, 16-byte aligned pointers. splits: 0 = auto (only used by VITK_EPI_ATOMIC_F32).
 */
int vitk_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major, int M,
                   int N, int K, int epilogue, const float* bias, const float* gamma, const float* resid,
                   long long ldr, void* out, long long ldo, void* out2, long long ldo2, const void* aux,
                   long long ldaux, int splits, void* stream);

/*
 * Same GEMM with the extra epilogue operands:
 *   rowscale[row / rows_per_sample] multiplies the branch in VITK_EPI_RESID_F32 (DropPath: mask/keep_prob per sample,
 *   models/cait.py:67,140; identity when null);
 *   tok_n / tok_N / tok_T drive VITK_EPI_TOKENS_F32 (PatchEmbed + cls/pos assembly: `resid` = pos_embed fp32 [tok_N, ldr],
 *   models/cait.py:229-233, models/deit.py:35-42, DINO prepare_tokens);
 *   colsum (optional fp32 [N], +=): column sums of the values stored to `out` by VITK_EPI_STORE_BF16 / VITK_EPI_DGELU
 *   -- the bias gradient of the Linear whose dY this GEMM produces, without another pass over dY.
 */
int vitk_gemm_bf16_ex(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major, int M,
                      int N, int K, int epilogue, const float* bias, const float* gamma, const float* resid,
                      long long ldr, void* out, long long ldo, void* out2, long long ldo2, const void* aux,
                      long long ldaux, int splits, const float* rowscale, int rows_per_sample, int tok_n, int tok_N,
                      int tok_T, float* colsum, void* stream);

/*
 * Batched GEMM: nbatch_h * nbatch_b independent problems D_b[M,N] = A_b[M,K] * B_b[N,K]^T addressed by element strides
 * per (inner, outer) batch index -- attention-style batches over (head, image) reading q/k/v in place from the qkv
 * Linear output. Epilogues VITK_EPI_STORE_BF16 / VITK_EPI_STORE_F32 only. Used by the CaiT talking-heads attention
 * (models/cait.py:116-125): S = q k^T, O = P' v, and their backward products. All strides % 8 == 0.
 */
int vitk_gemm_bf16_batched(const void* A, long long lda, long long sa_h, long long sa_b, int a_mn_major, const void* B,
                           long long ldb, long long sb_h, long long sb_b, int b_mn_major, int M, int N, int K,
                           int nbatch_h, int nbatch_b, int epilogue, void* out, long long ldo, long long so_h,
                           long long so_b, void* stream);

/*
 * LayerNorm forward over the last dim (nn.LayerNorm(D, eps=1e-6): models/cait.py:64,68,203,259; models/deit.py:98).
 *   x fp32 [rows, D] -> y bf16 [rows, D]; saves mean/rstd fp32 [rows]. D % 4 == 0, D <= 1024.
 */
int vitk_layernorm_fwd(const float* x, const float* weight, const float* bias, void* y_bf16, float* mean, float* rstd,
                       long long rows, int D, float eps, void* stream);

/*
 * LayerNorm backward fused with the residual-gradient add:
 *   dx_f32 = dres_f32 (optional) + LN'(dy_bf16);  dx_bf16 (optional) = bf16(dx_f32 * colscale[D] (optional))
 *   dweight/dbias fp32 [D] are ACCUMULATED (+=) with atomics.
 */
int vitk_layernorm_bwd(const void* dy_bf16, const float* x, const float* weight, const float* mean, const float* rstd,
                       const float* dres, float* dx, void* dx_bf16, const float* colscale, float* dweight,
                       float* dbias, long long rows, int D, void* stream);

/* Strided / fp32-output variants: x rows x_stride elements apart (final norm on the cls rows only: DINO
 * `norm(x)[:, 0]`, models/cait.py:244-246); y_bf16 and/or y_f32 [rows, D] dense. dy fp32 or bf16 [rows, D] dense;
 * dres / dx rows dx_stride apart. dxsum (optional, fp32 [D], +=) receives the column sums of the dx_bf16 copy, i.e.
 * the bias gradient of the Linear that consumes it (saves a separate pass over dx). */
int vitk_layernorm_fwd_ex(const float* x, long long x_stride, const float* weight, const float* bias, void* y_bf16,
                          float* y_f32, float* mean, float* rstd, long long rows, int D, float eps, void* stream);
int vitk_layernorm_bwd_ex(const void* dy, int dy_is_f32, const float* x, long long x_stride, const float* weight,
                          const float* mean, const float* rstd, const float* dres, float* dx, long long dx_stride,
                          void* dx_bf16, const float* colscale, float* dweight, float* dbias, float* dxsum,
                          long long rows, int D, void* stream);

/* Fused mean softmax cross-entropy + its gradient + argmax accuracy of logits fp32 [rows, C] (row pitch ld), labels
 * int64 [rows]: out2[0] = mean loss, out2[1] = number of rows whose first-maximum column equals the label (as a float),
 * dlogits (optional, row pitch ldd) = (softmax - onehot) / rows_that_count. Rows labelled -100 (nn.CrossEntropyLoss's
 * default ignore_index) are left out of the mean and get a zero gradient; any other label outside [0, C) makes the loss
 * NaN (torch raises a device assert there). Deterministic, no host synchronisation. Replaces
 * nn.CrossEntropyLoss + autograd (utils_network.py:429-433, main.py:244) and classification_count_correct
 * (utils_network.py:85-95). */
int vitk_cross_entropy(const float* logits, long long ld, const long long* labels, int rows, int C, float* out2,
                       float* dlogits, long long ldd, void* stream);

/* out[c] += sum_r x[r*ldx + c], fp32 (d_pos = sum_b dX[b,:,:], d_cls; SURVEY App. A.3 token assembly). cols % 4 == 0. */
int vitk_colsum_f32(const float* x, long long ldx, long long rows, long long cols, float* out, void* stream);

/* out[c] += sum_r a_f32[r,c] * b_bf16[r,c]  (LayerScale d_gamma = sum_rows dY o f; models/cait.py:144-149). */
int vitk_colsum_prod(const float* a, long long lda, const void* b_bf16, long long ldb, long long rows, int N, float* out,
                     void* stream);
/* Same with the DropPath factor of the branch (models/cait.py:67,140,148-149; timm Block): the forward value is
 * x + gamma * rowscale[row / rows_per_sample] * f, so d_gamma = sum_rows dY * rowscale * f. rowscale may be null. */
int vitk_colsum_prod_ex(const float* a, long long lda, const void* b_bf16, long long ldb, long long rows, int N,
                        float* out, const float* rowscale, long long rows_per_sample, void* stream);

/* LayerScale branch backward in one pass (models/cait.py:144-149, y = x + gamma * DropPath(f)):
 *   out_bf16[r,c] = bf16(dy[r,c] * gamma[c] * rowscale[r / rows_per_sample])   (gradient of f: operand of the dgrad / wgrad)
 *   dgamma[c] += sum_r dy * rowscale * f ;  dbias[c] += sum_r dy * gamma * rowscale   (bias gradient of f's Linear)
 * rowscale / dgamma / dbias may be null. out_bf16 is dense [rows, N]. N % 8 == 0. */
int vitk_layerscale_bwd(const float* dy, long long lddy, const void* f_bf16, long long ldf, const float* gamma,
                        const float* rowscale, long long rows_per_sample, long long rows, int N, void* out_bf16,
                        float* dgamma, float* dbias, void* stream);

/* out_bf16[r,:] = bf16(x[(r / rows_per_group) * group_stride + (r % rows_per_group) * D + :] * colscale * rowscale[r / rows_per_sample])
 * -- the bf16 copy of a residual-stream gradient that the dgrad/wgrad GEMMs read (LayerScale / DropPath folded in),
 * also used to compact dX[:, T:, :] into patch rows for the PatchEmbed wgrad. D % 8 == 0. */
int vitk_scale_cast(const float* x, long long rows_per_group, long long group_stride, long long rows, int D,
                    const float* colscale, const float* rowscale, long long rows_per_sample, void* out_bf16,
                    void* stream);

/* x fp32 [B,C,H,W] -> bf16 [B*(H/P)*(W/P), C*P*P] patch rows, k = (c,i,j): the A operand of the PatchEmbed GEMM
 * (Conv2d(C,D,P,P) + flatten(2).transpose(1,2); in-repo analogue models/swin.py:434,445). P % 4 == 0. */
int vitk_patchify(const float* x, void* out_bf16, int B, int C, int H, int W, int P, void* stream);

/*
 * PatchEmbed as an im2col-free patch GEMM: Conv2d(C, D, kernel=P, stride=P)(x).flatten(2).transpose(1,2) + bias + pos
 * written straight into the token buffer (DINO / timm PatchEmbed, call site models/vision_all.py:156,161-167; in-repo
 * witness models/swin.py:434-445; token assembly models/cait.py:229-234, models/deit.py:35-43). The A operand is
 * gathered by TMA from the NCHW image (rank-5 tensor map {px, pc, py, pr, c*b}); no [B*n, C*P*P] patch matrix exists.
 *   img: fp32 [B,C,H,W] (tcgen05 kind::tf32 on the fp32 pixels and the fp32 weight) or bf16 [B,C,H,W] (kind::f16 with
 *   the bf16 weight copy); weight [D, C*P*P] of the same type; bias fp32 [D] or null; pos fp32 [tok_N, ldpos];
 *   out fp32 [B, tok_N, D]: out[b, tok_T + p, :] = patch_p . W^T + bias + pos[tok_T + p]. P * sizeof(elem) in {16,32,64,128}.
 */
int vitk_patch_embed_fwd(const void* img, int img_is_bf16, const void* weight, const float* bias, const float* pos,
                         long long ldpos, float* out, int B, int C, int H, int W, int P, int D, int tok_N, int tok_T,
                         void* stream);
/* Its weight gradient (autograd of the conv): dW fp32 [D, C*P*P] += sum_{b,p} dY[b, tok_T+p, :]^T
 * patch_p, with dY read in place from a bf16 token gradient [B, dy_tok_N, D] and the patches gathered from the bf16
 * image again (img_is_bf16 must be 1: MN-major tf32 operands need 128-byte rows, so fp32 images are cast once by the
 * caller). Both operands are MN-major; images are split over thread blocks (red.global.add). */
int vitk_patch_embed_wgrad(const void* img, int img_is_bf16, const void* dy, int dy_tok_N, int dy_tok_T, float* dW,
                           int B, int C, int H, int W, int P, int D, void* stream);
/* Device input pipeline: uint8 [B,C,H,W] -> bf16 (x/255 - mean[c]) / std[c], i.e. ToTensor + Normalize
 * (utils_datasets.py:573-580) after the H2D copy of the raw bytes. H*W % 16 == 0. */
int vitk_normalize_u8(const void* x_u8, void* y_bf16, const float* mean, const float* stdv, int B, int C, int H, int W,
                      void* stream);

/* out[b, t, :] = tok[t,:] + pos[t,:] for the T prefix tokens (cls, dist) of every image. */
int vitk_prefix_tokens(const float* tok, const float* pos, float* out, int B, int T, long long tokens_per_image, int D,
                       void* stream);

/* Column sums of a bf16 matrix accumulated (+=) into fp32 out[N] (bias gradients: db = sum_rows dY). */
int vitk_colsum_bf16(const void* x_bf16, long long ldx, long long rows, int N, float* out, void* stream);

/* Multi-tensor SGD with momentum, torch.optim.SGD semantics with dampening 0 / no nesterov / no weight decay (the
 * reference's 'sgd' optimiser, utils_network.py:119-126, stepped at utils_network.py:439-442), fused with the refresh
 * of the bf16 GEMM-operand copy of each weight:  buf = momentum*buf + grad_scale*g ; p -= lr*buf ; w_bf16 = bf16(p).
 *   table:     device array of { float* p; const float* g; float* buf; bf16* w_or_null; long long n; } (40 bytes each)
 *   chunk_map: device array of int2 { tensor index, chunk index }, one thread block per vitk_sgd_chunk_elems() elements
 *   first_step: momentum buffers are uninitialised (buf = grad_scale*g), as torch does on the first step. */
int vitk_sgd_chunk_elems(void);
int vitk_sgd_momentum_multi(const void* table, const void* chunk_map, int num_chunks, float lr, float momentum,
                            float grad_scale, int first_step, void* stream);

/* Same update with the hyper-parameters read from DEVICE memory: hyper = { lr, momentum, grad_scale } (fp32). A step
 * captured in a CUDA graph then follows the reference's LambdaLR schedule (utils_network.py:218-225) without
 * re-capture: the host rewrites `hyper` (stream-ordered copy) between replays. */
int vitk_sgd_momentum_multi_hp(const void* table, const void* chunk_map, int num_chunks, const float* hyper,
                               void* stream);

/* Multi-tensor Adam / AdamW (torch.optim.Adam / AdamW without amsgrad / maximize: the reference's 'adam' and 'adamw'
 * optimisers, utils_network.py:119-126), fused with the bf16 weight refresh. table: device array of
 * { float* p; const float* g; float* exp_avg; float* exp_avg_sq; bf16* w_or_null; int64 n } per tensor; chunk_map as
 * for SGD. hyper (device, fp32[11], read AND written) = { lr, beta1, beta2, eps, weight_decay, decoupled (1 = AdamW,
 * 0 = Adam: L2 term added to the gradient), 1-beta1, 1-beta2, step, step_size, inv_sqrt_bias2 }: each call first advances
 * `step` by one and derives the last two (bias corrections) on the device, so the step counter is graph-capturable.
 * Two launches. */
int vitk_adam_multi(const void* table, const void* chunk_map, int num_chunks, float* hyper, void* stream);

/* bf16 -> fp32 of n elements (gradients that travelled through a bf16 all-reduce back into the fp32 arena). */
int vitk_cast_bf16_f32(const void* x_bf16, float* y, long long n, void* stream);

/* fp32 -> bf16 cast of n elements (weights, activations). n % 8 == 0 not required. */
int vitk_cast_f32_bf16(const float* x, void* y_bf16, long long n, void* stream);

/*
 * Fused attention forward: out = softmax(q k^T * scale) v per (batch, head); never materialises [B,H,N,N].
 *   qkv bf16 [B, N, 3, H, d] (the raw qkv Linear output), out bf16 [B, N, H, d] (heads merged, what proj consumes),
 *   lse2 fp32 [B, H, N] = log2-domain logsumexp of the scaled scores (saved for backward). d in {64, 48}.
 * Replaces the core of timm/DINO Attention.forward (in-repo witness models/swin.py:119-144).
 */
int vitk_attn_fwd(const void* qkv_bf16, void* out_bf16, float* lse2, int B, int N, int H, int d, float scale,
                  void* stream);

/*
 * Fused attention backward (recompute): given qkv, the forward output O, dO and lse2, writes dqkv bf16
 * [B, N, 3, H, d] in place (dQ | dK | dV, the layout the qkv dgrad/wgrad GEMMs consume). delta fp32 [B,H,N] is a
 * caller-provided scratch (rowsum(dO o O)). Deterministic: no atomics. Autograd of the eager attention in the
 * reference.
 */
int vitk_attn_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2, float* delta,
                  void* dqkv_bf16, int B, int N, int H, int d, float scale, void* stream);

/* vitk_attn_bwd that also accumulates (+=) the column sums of dqkv into dqkv_bias_grad fp32 [3*H*d] -- the gradient of
 * the qkv Linear bias (db = sum_rows dY) -- from the registers that hold the dQ / dK / dV tiles. */
int vitk_attn_bwd_ex(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2, float* delta,
                     void* dqkv_bf16, float* dqkv_bias_grad, int B, int N, int H, int d, float scale, void* stream);

/*
 * Whole-head attention backward for short sequences (d = 64, N <= 256: every ViT-S/B 16 / DeiT configuration at 224 px):
 * one thread block per (image, head) stages Q, K, V, dO of the head once, computes S and dP ONCE per tile pair, keeps
 * dQ (4 query tiles), dK and dV in TMEM and forms delta in its prologue -- no atomics, no workspace, no second kernel.
 * Same inputs / outputs / determinism as vitk_attn_bwd_ex. Opt-in: vitk_attn_bwd_head_supported(N, d) returns 1 only
 * with VITK_ATTN_BWD_HEAD=1 in the environment (measured slower than the two-kernel backward at 197 tokens: one block
 * per SM leaves one tile chain in flight instead of two; kept parity-tested).
 */
int vitk_attn_bwd_head_supported(int N, int d);
int vitk_attn_bwd_head(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2,
                       void* dqkv_bf16, float* dqkv_bias_grad, int B, int N, int H, int d, float scale, void* stream);

/*
 * Single-kernel attention backward (d = 64): S / dP and the elementwise pass are computed once per (key block, query
 * tile); dV, dK accumulate in TMEM, dQ tiles are summed across key blocks with fp32 atomics into the caller-provided
 * workspace dq_f32_ws [B*N, H*d] (zeroed inside) and then written as bf16 into dqkv. Launches: delta pre-pass, memset,
 * fused kernel, conversion. Same inputs / outputs as vitk_attn_bwd; dQ is not bit-reproducible run to run (atomics).
 */
int vitk_attn_bwd_fused(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2, float* delta,
                        float* dq_f32_ws, void* dqkv_bf16, int B, int N, int H, int d, float scale, void* stream);

/*
 * CaiT talking-heads mixing (models/cait.py:116-125), forward: S fp32 [B,H,N,Np] (raw q.k logits, row pitch Np >= N,
 * Np % 8 == 0) -> Pm bf16 [B,H,N,Np] = Ww softmax_j(scale*Wl S + bl) + bw (pad columns zero); rowmax/rowsum fp32
 * [B,H,N] are saved for backward. wl/ww fp32 [H,H] (= proj_l.weight / proj_w.weight), bl/bw fp32 [H]. H in {2,4,6,8,16}.
 */
int vitk_th_mix_fwd(const float* S, const float* wl, const float* bl, const float* ww, const float* bw, float scale,
                    void* Pm_bf16, float* rowmax, float* rowsum, int B, int H, int N, int Np, void* stream);
/* Backward of the above: dPm fp32 [B,H,N,Np] -> dS bf16 [B,H,N,Np] (gradient w.r.t. the RAW logits, i.e. scale folded
 * in); dwl/dbl/dww/dbw are accumulated (+=). */
int vitk_th_mix_bwd(const float* S, const float* dPm, const float* rowmax, const float* rowsum, const float* wl,
                    const float* bl, const float* ww, const float* bw, float scale, void* dS_bf16, float* dwl, float* dbl,
                    float* dww, float* dbw, int B, int H, int N, int Np, void* stream);
/* Version 2 of the mixing kernels (one warp per query row, every head mix and the weight-gradient products on the
 * tensor cores through mma.sync tf32, 3xTF32 for the logits) serves rows of up to 208 keys: vitk_th_mix_fwd picks it
 * by itself; its backward takes dP' in bf16 (half the traffic of the fp32 form). vitk_th_mix_supports_bf16_dp(Np)
 * returns 1 when that backward is available for rows of pitch Np. */
int vitk_th_mix_supports_bf16_dp(int Np);
int vitk_th_mix_bwd_bf16(const float* S, const void* dPm_bf16, const float* rowmax, const float* rowsum, const float* wl,
                         const float* bl, const float* ww, const float* bw, float scale, void* dS_bf16, float* dwl,
                         float* dbl, float* dww, float* dbw, int B, int H, int N, int Np, void* stream);

/* bf16 logit planes S bf16 [B,H,N,Np] -- what the reference's `q @ k.transpose(-2, -1)` produces under
 * torch.autocast(bfloat16) (models/cait.py:116) -- with the version-2 mixing kernels; otherwise as vitk_th_mix_fwd /
 * vitk_th_mix_bwd_bf16 (bf16 values are exact in tf32, so the 3xTF32 logit mix drops to two products). */
int vitk_th_mix_fwd_s16(const void* S_bf16, const float* wl, const float* bl, const float* ww, const float* bw, float scale,
                        void* Pm_bf16, float* rowmax, float* rowsum, int B, int H, int N, int Np, void* stream);
int vitk_th_mix_bwd_s16(const void* S_bf16, const void* dPm_bf16, const float* rowmax, const float* rowsum, const float* wl,
                        const float* bl, const float* ww, const float* bw, float scale, void* dS_bf16, float* dwl,
                        float* dbl, float* dww, float* dbw, int B, int H, int N, int Np, void* stream);

/*
 * The per-(image, head) products around the talking-heads mixes (models/cait.py:116 `q @ k^T`, :125 `attn @ v`, and
 * their autograd) for short sequences: N <= 208 tokens, d in {48, 64}, Np % 8 == 0 (vitk_th_gemm_supported). Operands
 * are token-major bf16 matrices [B*N, ld] read in place (head h = columns col0 + h*d .. + d of a matrix that has
 * `cols` valid columns) and planes [B,H,N,Np]:
 *   vitk_th_scores : out[b,h,i,j] = sum_e a[b*N+i, a_col0+h*d+e] * b[b*N+j, b_col0+h*d+e]   (out bf16, or fp32 if out_f32;
 *                    pad columns [N, Np) are written as zeros)
 *   vitk_th_apply  : out[b*N+i, o_col0+h*d+e] = sum_j p[b,h,i,j] * x[b*N+j, x_col0+h*d+e]    (transpose = 0)
 *                    out[b*N+j, o_col0+h*d+e] = sum_i p[b,h,i,j] * x[b*N+i, x_col0+h*d+e]    (transpose = 1)
 *                    p bf16 planes whose pad columns are zero (as vitk_th_mix_* write them); out bf16 [B*N, ldo];
 *                    colsum: optional fp32 [H*d], += column sums of the H*d output columns (the slice of the qkv bias
 *                    gradient that belongs to this product: autograd of `self.qkv` with qkv_bias, models/cait.py:99).
 */
int vitk_th_gemm_supported(int N, int d, int Np);
int vitk_th_scores(const void* a_bf16, long long lda, int a_cols, int a_col0, const void* b_bf16, long long ldb, int b_cols,
                   int b_col0, void* out, int out_f32, int B, int N, int H, int d, int Np, void* stream);
int vitk_th_apply(const void* p_bf16, const void* x_bf16, long long ldx, int x_cols, int x_col0, void* out_bf16,
                  long long ldo, int o_col0, int transpose, float* colsum, int B, int N, int H, int d, int Np, void* stream);

/*
 * CaiT class attention (models/cait.py:38-55): one query row per (image, head). q bf16 [B,C] (unscaled), keys/values:
 * row 0 = class token kc/vc bf16 [B, ldc], rows 1..n = patch tokens kx/vx bf16 [B*n, ldkv]. out bf16 [B,C];
 * p fp32 [B,H,n+1] (softmax probabilities, saved for backward). d = C/H in {48, 64}.
 */
int vitk_class_attn_fwd(const void* q, const void* kc, const void* kx, const void* vc, const void* vx, long long ldkv,
                        long long ldc, float scale, void* out, float* p, int B, int H, int n, int d, void* stream);
int vitk_class_attn_bwd(const void* q, const void* kc, const void* kx, const void* vc, const void* vx, long long ldkv,
                        long long ldc, const float* p, const void* dout, float scale, void* dq, void* dkc, void* dkx,
                        void* dvc, void* dvx, long long lddkv, long long lddc, int B, int H, int n, int d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITK_H_ */
