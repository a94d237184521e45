/*
 * ffm_b200.h — C ABI of libffm_b200.so: the sm_100a (B200) kernels behind the FairLoRA training
 * hot path of Harvard-AI-and-Robotics-Lab/FairFedMed.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry point below therefore
 * cites the reference *Python call site* it replaces (paths relative to the reference repo root).
 * A maintainer binds these with ctypes (see INTEGRATION.md and fairfedmed_b200/_cabi.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing
 *     and never throws: it returns 0 on success or a negative errno-style code
 *     (-22 invalid argument, -5 CUDA error, -38 not supported); ffm_last_error() gives the text;
 *   - matrices are dense row-major; "bf16" means __nv_bfloat16 storage, "f32" IEEE binary32;
 *   - callers own all workspaces (sizes from the *_workspace_bytes helpers).
 */
#ifndef FFM_B200_H_
#define FFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef FFM_STREAM_T
#define FFM_STREAM_T
#ifdef __CUDACC__
#include <cuda_runtime.h>
typedef cudaStream_t ffm_stream_t;
#else
typedef void* ffm_stream_t;
#endif
#endif

/* ---------------------------------------------------------------- library ------------------ */
const char* ffm_last_error(void);
int ffm_version(void);

/*
 * Instrumentation used by bench.py (no reference counterpart).
 *   ffm_launch_count : number of kernels this library has launched since the last reset (host-side counter).
 *   ffm_profile_*    : when enabled, every launch of the fused SVLoRA GEMM kernel is bracketed by CUDA events on
 *                      its own stream; ffm_profile_read synchronises them and returns per-launch milliseconds and
 *                      (T, K, N) triples (host arrays), then clears the record list.  N is negated for launches of the
 *                      adapter-free build (ffm_frozen_linear): their FLOPs are 2*T*K*|N| without the low-rank terms.
 */
long long ffm_launch_count(int reset);
int ffm_profile_enable(int enable);
int ffm_profile_read(float* ms_host, int* tkn_host, int max_records);

/* ------------------------------------------------- FairLoRA / SVLoRA linear ------------------ */
/* Maximum adapter rank the fused GEMM was built for (32: the RN50 recipe, scripts/fairfedlora_fairfedmed_rn50.sh:37). */
int ffm_svlora_max_rank(void);
/* Padded rank used for rank r: 16 for r <= 16 (ViT recipes, r = 12), 32 for r <= 32; 0 if r is unsupported.  It is the
 * row length of the h / z side outputs of ffm_svlora_fwd. */
int ffm_svlora_padded_rank(int r);

size_t ffm_svlora_fwd_workspace_bytes(int T, int K, int N, int n_samples);
size_t ffm_svlora_bwd_workspace_bytes(int T, int K, int N, int n_samples);

/*
 * ffm_svlora_prepare — the per-call operand preparation of ffm_svlora_fwd on its own: fills `workspace`
 * (ffm_svlora_fwd_workspace_bytes(T, K, N, n_samples); T is ignored) with the bf16 adapter tiles of both directions
 * and the scaled singular values.  It depends on the parameters and the attribute rows only, not on activations, so a
 * host can issue it for every layer ahead of time (another stream) and then call ffm_svlora_fwd with
 * lora_a = lora_b = NULL (s_eff ignored) and that workspace: the forward then skips its own preparation launch.
 */
int ffm_svlora_prepare(const float* lora_a, const float* lora_b, const float* s_eff, void* workspace,
                       size_t workspace_bytes, int K, int N, int r, int n_samples, float scaling, ffm_stream_t stream);

/*
 * Forward of FairLoRALinear / SVLoRALinear / LoRALinear
 *   replaces trainers/GLP_OT_SVLoRA.py:450-482 (FairLoRALinear.forward), :308-312, :241-242.
 *
 *   h[t,:]  = x[t,:] · A                                      (f32 [T,rp], rp = ffm_svlora_padded_rank(r); columns >= r are 0)
 *   u[t,:]  = x[t,:] · W^T + bias + scaling · (h[t,:] ⊙ s_eff[sample(t),:]) · B
 *   y       = act ? QuickGELU(u) : u          (clip/model.py:313-315 fused when act = 1)
 *   y_dact  = QuickGELU'(u) (only when act = 1 and y_dact != NULL): the one thing the backward pass needs from the
 *             activation, so u itself is never written.
 *
 *   sample(t) = ((t / row_div) mod b_prime) / num_slices.  row_div = 1: rows are sequence-first [L, B', C] like the
 *               reference's activations (clip/model.py:438-440); row_div = L: rows are batch-first [B', L, C]
 *               (used by fairfedmed_b200.clip_model so attention needs no transposes).  OCT volumes fold
 *               num_slices slice-images per sample into B' (trainers/GLP_OT_SVLoRA.py:473-475).
 *
 *   z[t,:]  = bf16(scaling · h[t,:] ⊙ s_eff[sample(t),:])   (bf16 [T,rp], optional): the operand d_lora_b needs,
 *             produced by the epilogue anyway (it feeds the rank-16 tensor-core update), so the backward pass does not
 *             recompute it.
 *
 *   x [T,K] bf16, W [N,K] bf16 (frozen nn.Linear weight), bias [N] f32 or NULL,
 *   lora_a [K,r] f32, lora_b [r,N] f32, s_eff [n_samples,r] f32 (from ffm_seff), y / y_dact [T,N] bf16.
 *   workspace: ffm_svlora_fwd_workspace_bytes(); afterwards it holds the bf16 adapter tiles of BOTH directions.  Keep it
 *   alive and pass it to ffm_svlora_bwd as fwd_workspace to save that call its own preparation launch.
 *   Requirements: K % 8 == 0, N % 8 == 0, 1 <= r <= ffm_svlora_max_rank(), 16-byte aligned pointers.
 */
int ffm_svlora_fwd(const void* x, const void* w, const float* bias, const float* lora_a, const float* lora_b,
                   const float* s_eff, void* y, void* y_dact, float* h_out, void* z_out, void* workspace,
                   size_t workspace_bytes, int T, int K, int N, int r, int n_samples, int b_prime, int num_slices,
                   int row_div, float scaling, int act, ffm_stream_t stream);

/*
 * Backward of the same module (autograd of trainers/GLP_OT_SVLoRA.py:450-482; W and bias frozen :375-376).
 *
 *   dzu = dy · B^T                         dz = scaling·dzu       dh = dz ⊙ s_eff[sample]
 *   dx  = dy · W + dh · A^T               (bf16 [T,K]); if gelu_dact != NULL (bf16 [T,K], the y_dact output of the
 *         *preceding* layer's forward) dx is multiplied element-wise by it, i.e. back through that QuickGELU
 *   d_lora_a [K,r] = x^T · dh             d_lora_b [r,N] = z^T · dy       (z from ffm_svlora_fwd)
 *   d_s_eff [n_samples,r] = scaling · sum_{t in sample} dzu ⊙ h
 *
 *   Three launches: the fused tcgen05 GEMM (dx, dzu, dh), one kernel for both adapter contractions (x and dy are
 *   each read from HBM once), one kernel that folds their partials and does the per-sample segmented reduction.
 *   w_t is W transposed, [K,N] bf16 (the frozen weight is transposed once at module construction).
 *   h (f32 [T,rp]) and z (bf16 [T,rp]) are the side outputs of ffm_svlora_fwd; fwd_workspace is that call's workspace
 *   (adapter tiles already prepared) or NULL (then lora_a / lora_b / s_eff are converted again here).
 */
int ffm_svlora_bwd(const void* dy, const void* x, const void* w_t, const float* lora_a, const float* lora_b,
                   const float* s_eff, const float* h, const void* z, const void* fwd_workspace,
                   const void* gelu_dact, void* dx, float* d_lora_a, float* d_lora_b, float* d_s_eff, void* workspace,
                   size_t workspace_bytes, int T, int K, int N, int r, int n_samples, int b_prime, int num_slices,
                   int row_div, float scaling, ffm_stream_t stream);

/*
 * The same backward in two separately launchable phases (bit mask): FFM_BWD_DX = the fused dX GEMM (writes dx and the
 * side outputs dzu / dh into `workspace`), FFM_BWD_PARAMS = the adapter-gradient contractions and the per-sample ds_eff
 * (read x, dy, h, z and the side outputs of phase DX from the SAME workspace).  Only dx feeds the layers below, so a
 * host may issue phase PARAMS on another stream once phase DX has completed (event) and let it overlap the rest of
 * the backward pass; keep the workspace alive until it has run.  ffm_svlora_bwd == both phases, one stream.
 */
#define FFM_BWD_DX 1
#define FFM_BWD_PARAMS 2
int ffm_svlora_bwd_phase(const void* dy, const void* x, const void* w_t, const float* lora_a, const float* lora_b,
                         const float* s_eff, const float* h, const void* z, const void* fwd_workspace,
                         const void* gelu_dact, void* dx, float* d_lora_a, float* d_lora_b, float* d_s_eff,
                         void* workspace, size_t workspace_bytes, int T, int K, int N, int r, int n_samples,
                         int b_prime, int num_slices, int row_div, float scaling, int phases, ffm_stream_t stream);

/* ------------------------------------------------- residual add + LayerNorm (frozen blocks) -- */
/*
 * The glue between the adapted MLP and the frozen attention of ResidualAttentionBlock (clip/model.py:354-357):
 *   x = x + branch;  h = LayerNorm(x)        LayerNorm with fp32 statistics like clip/model.py:304-310
 * fused into one pass (row-sized tensors moved: 2 reads + 2 writes instead of 3 reads + 2 writes + 1 read + 1 write).
 *   x, res, sum_out, ln_out: bf16 [rows, C];  gamma, beta: f32 [C] (frozen);  mean_out, rstd_out: f32 [rows] (saved
 *   for the backward).  res == NULL: plain LayerNorm of x (sum_out unused).  C in {256, 512, 768, 1024}.
 * Backward (LayerNorm weights are frozen in the FairLoRA recipe, trainers/GLP_OT_SVLoRA.py:822-829):
 *   dx = LayerNorm'(d_ln; s, mean, rstd, gamma) + d_res        (d_res == NULL: no residual gradient to merge)
 *   the same dx is the gradient of both x and res.
 */
int ffm_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* sum_out,
                          void* ln_out, float* mean_out, float* rstd_out, int rows, int C, float eps,
                          ffm_stream_t stream);
int ffm_add_layernorm_bwd(const void* d_ln, const void* d_res, const void* s, const float* gamma, const float* mean,
                          const float* rstd, void* dx, int rows, int C, ffm_stream_t stream);

/* ------------------------------------------------- frozen projections (scope row f1) ---------- */
/*
 * y[T, N] = x[T, K] · w[N, K]^T + bias — the in_proj / out_proj of nn.MultiheadAttention inside
 * ResidualAttentionBlock (clip/model.py:350-352) on the same tcgen05 / TMA pipeline as ffm_svlora_fwd, built without the
 * adapter side product (UMMA N = 192, no fix-up).  The backward of a frozen projection is the same call on the
 * transposed weight: dx[T, K] = dy[T, N] · w_t[K, N]^T.  bf16 operands, fp32 bias (or NULL), bf16 result.
 */
int ffm_frozen_linear(const void* x, const void* w, const float* bias, void* y, int T, int K, int N, ffm_stream_t stream);

/* ------------------------------------------------- frozen attention core (scope row f1) ------- */
/*
 * softmax(Q K^T / sqrt(head_dim)) V per (sample, head) — the core of nn.MultiheadAttention as ResidualAttentionBlock
 * calls it (clip/model.py:350-352, need_weights=False; causal = the text tower's upper-triangular -inf mask,
 * clip/model.py:520-526).  q, k, v are read in place from the packed in_proj output and dq, dk, dv are written packed
 * the same way, so there is no head split / transpose / concat around the kernels:
 *   qkv, d_qkv  bf16 [B, L, 3, H, head_dim] (batch_first = 1) or [L, B, 3, H, head_dim] (0, the reference's layout)
 *   out, d_out  bf16 [B, L, H*head_dim]     or [L, B, H*head_dim]
 *   lse         f32  [B*H, L]  base-2 log-sum-exp of the scaled scores, forward -> backward
 * head_dim must be 64 and L <= ffm_attention_max_len() (208): the CLIP towers (197 image tokens, 77 text tokens).
 * Deterministic (no atomics).  No dropout (the reference trains with attention dropout 0).
 */
int ffm_attention_max_len(void);
int ffm_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int head_dim, int causal,
                      int batch_first, ffm_stream_t stream);
int ffm_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* d_qkv, int B, int L,
                      int H, int head_dim, int causal, int batch_first, ffm_stream_t stream);

/* ------------------------------------------------- ViT input side (scope row f3) -------------- */
/*
 * ffm_patchify_normalize — CustomCLIP.forward's `image / 255`, `(image - mean) / std`
 * (trainers/GLP_OT_SVLoRA.py:679-693), the cast to half precision and the im2col of the stride-`patch` convolution
 * ModifiedVisionTransformer.conv1 (clip/model.py:431-433) in one pass:
 *   patches[b*G + gy*gw + gx, c*patch*patch + py*patch + px] =
 *       bf16((image[b, c, gy*patch+py, gx*patch+px] / 255 - mean[c]) / std[c])         (IEEE fp32 divisions)
 *   image f32 [Bp, C, H, W] (0..255; div255 = 0 skips the first division: OCT inputs are already min-max scaled),
 *   patches bf16 [Bp*G, C*patch*patch], mean / std f32 [C].  patch % 8 == 0, H % patch == W % patch == 0.
 * The patch embedding is then the GEMM patches · conv1.weight.flatten(1)^T.
 *
 * ffm_vit_embed_ln — clip/model.py:434-440 and the first block's ln_1 (:354):
 *   x0[b, 0, :]   = LN_pre(class_embedding + positional_embedding[0])
 *   x0[b, l>0, :] = LN_pre(patch_emb[b*G + l-1, :] + positional_embedding[l]),   h0 = LN_1(x0)
 *   patch_emb bf16 [Bp*G, C]; tables / LayerNorm parameters f32; x0, h0 bf16 [Bp, G+1, C]; mean1 / rstd1 f32
 *   [Bp*(G+1)] or NULL (statistics of LN_1, the format ffm_add_layernorm_bwd takes).  C in {256, 512, 768, 1024}.
 * Forward only: conv1, the embeddings and ln_pre are frozen and the image needs no gradient.
 */
int ffm_patchify_normalize(const float* image, void* patches, const float* mean, const float* stdv, int Bp, int C,
                           int H, int W, int patch, int div255, ffm_stream_t stream);
int ffm_vit_embed_ln(const void* patch_emb, const float* class_embedding, const float* positional_embedding,
                     const float* ln_pre_gamma, const float* ln_pre_beta, const float* ln1_gamma, const float* ln1_beta,
                     void* x0, void* h0, float* mean1, float* rstd1, int Bp, int G, int C, float eps_pre, float eps_1,
                     ffm_stream_t stream);

/*
 * OCT input side — trainers/GLP_OT_SVLoRA.py:686-693 after the trainable slice projection y = proj_per_3d_slice(slices / 255):
 *   z = (y - amin(y)) / (amax(y) - amin(y) + 1e-5) per slice-image over (C, H, W), then (z - mean[c]) / std[c], cast and
 *   im2col for the stride-`patch` convolution.
 * ffm_oct_minmax_patchify: y f32 [Bp, C, H, W] -> lo / hi f32 [Bp] and patches bf16 [Bp*G, C*patch*patch]
 *   (two launches: one reduction pass, one conversion pass).
 * ffm_oct_input_bwd: d_patches bf16 (gradient of `patches`) -> d_y f32 [Bp, C, H, W], including the min / max terms
 *   (torch.amin / amax split the gradient evenly among the pixels attaining the extremum); deterministic; `ws` of at
 *   least ffm_oct_input_bwd_ws_bytes(Bp) bytes.
 */
int ffm_oct_minmax_patchify(const float* y, float* lo, float* hi, void* patches, const float* mean, const float* stdv,
                            int Bp, int C, int H, int W, int patch, ffm_stream_t stream);
size_t ffm_oct_input_bwd_ws_bytes(int Bp);
int ffm_oct_input_bwd(const void* d_patches, const float* y, const float* lo, const float* hi, const float* stdv,
                      float* d_y, void* ws, size_t ws_bytes, int Bp, int C, int H, int W, int patch, ffm_stream_t stream);

/*
 * OCT slice projection — proj_per_3d_slice = Conv2d(dim_per_3d_slice -> 3, kernel 5, padding 2) on image / 255
 * (trainers/GLP_OT_SVLoRA.py:587-595, :684); in_scale (1 / 255) is folded into the weights.
 *   x f32 [Bp, Cin, H, W] (the raw slices), w f32 [Cout, Cin, 5, 5], bias f32 [Cout], y / dy f32 [Bp, Cout, H, W];
 *   Cin <= 32, Cout <= 4, W a multiple of 4.
 * ffm_oct_slice_conv_wgrad: dw [Cout, Cin, 5, 5] and dbias [Cout] of the same expression (the input is data and needs no
 * gradient); deterministic; `ws` of at least ffm_oct_slice_conv_wgrad_ws_bytes(Cin) bytes.
 */
size_t ffm_oct_slice_conv_wgrad_ws_bytes(int Cin);
int ffm_oct_slice_conv_fwd(const float* x, const float* w, const float* bias, float* y, int Bp, int Cin, int Cout, int H,
                           int W, float in_scale, ffm_stream_t stream);
int ffm_oct_slice_conv_wgrad(const float* x, const float* dy, float* dw, float* dbias, void* ws, size_t ws_bytes, int Bp,
                             int Cin, int Cout, int H, int W, float in_scale, ffm_stream_t stream);

/*
 * Merged weight of a plain LoRA projection — LoRALinear.weight(x, attr), trainers/GLP_OT_SVLoRA.py:235-239, consumed by the
 * RN50 attention pool (clip/model.py:88-97):  out[o, i] = W[o, i] + scaling * sum_j A[i, j] * B[j, o]
 *   W, out f32 [out_f, in_f]; A = lora_A.weight f32 [in_f, r]; B = lora_B.weight f32 [r, out_f]; r <= 32.
 * ffm_lora_merged_weight_bwd: dWm f32 [out_f, in_f] -> dA [in_f, r], dB [r, out_f] (deterministic; `ws` of at least
 * ffm_lora_merged_weight_ws_bytes(out_f, in_f) bytes).
 */
size_t ffm_lora_merged_weight_ws_bytes(int out_f, int in_f);
int ffm_lora_merged_weight(const float* W, const float* A, const float* B, float* out, int out_f, int in_f, int r,
                           float scaling, ffm_stream_t stream);
int ffm_lora_merged_weight_bwd(const float* dWm, const float* A, const float* B, float* dA, float* dB, float* ws,
                               size_t ws_bytes, int out_f, int in_f, int r, float scaling, ffm_stream_t stream);

/*
 * nn.AvgPool2d(k) of the ResNet trunk (clip/model.py:30, :42, :108) on channels-last fp32 activations:
 *   x, dx [B, H, W, C] (NHWC memory), y, dy [B, H/k, W/k, C], f32 (elem_bf16 = 0) or bf16 (elem_bf16 = 1, fp32
 *   accumulation); H, W multiples of k, C a multiple of 4 (fp32) / 8 (bf16).
 */
int ffm_avgpool_nhwc_fwd(const void* x, void* y, int B, int H, int W, int C, int k, int elem_bf16, ffm_stream_t stream);
int ffm_avgpool_nhwc_bwd(const void* dy, void* dx, int B, int H, int W, int C, int k, int elem_bf16, ffm_stream_t stream);
/* out f32 [n] = in bf16 [n] (contiguous, n a multiple of 8): the adapters' bf16 output back into the fp32 ResNet trunk
 * (the `.to(x.dtype)` at the end of FairLoRALinear.forward, trainers/GLP_OT_SVLoRA.py:479-482). */
int ffm_widen_bf16(const void* in, float* out, int64_t n, ffm_stream_t stream);
/* y = relu(a + b) — the closing `relu(out + identity)` of a bottleneck (clip/model.py:56-58) — and its backward
 * g = dy * [y > 0] (the gradient of both a and b); f32, same memory layout for all operands, n a multiple of 4. */
int ffm_add_relu(const float* a, const float* b, float* y, int64_t n, ffm_stream_t stream);
int ffm_relu_mask(const float* dy, const float* y, float* g, int64_t n, ffm_stream_t stream);

/*
 * Training-mode nn.BatchNorm2d (+ the ReLU that follows it) of the ResNet trunk (clip/model.py:18-58) on channels-last fp32
 * activations viewed as x f32 [R = B*H*W, C]; gamma / beta f32 [C] (trainable in the RN50 recipe).
 * ffm_bn_relu_fwd: y = [relu](gamma * (x - mean) * rstd + beta) with the batch statistics (biased variance), which are also
 *   returned (mean_out, rstd_out f32 [C]) for the backward; running_mean / running_var (may both be NULL) are updated as torch
 *   does (momentum, unbiased variance).  ffm_bn_relu_bwd: dx, dgamma, dbeta (the ReLU mask is recomputed from x).
 * `ws` of at least ffm_bn_ws_bytes(C) bytes; C a multiple of 4; deterministic.
 */
size_t ffm_bn_ws_bytes(int C);
int ffm_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float* y,
                    float* mean_out, float* rstd_out, void* ws, size_t ws_bytes, int64_t R, int C, float momentum, float eps,
                    int relu, ffm_stream_t stream);
int ffm_bn_relu_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* mean,
                    const float* rstd, float* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int64_t R, int C,
                    int relu, ffm_stream_t stream);

/*
 * Group mixing of singular values — trainers/GLP_OT_SVLoRA.py:453-467.
 *   attr != NULL: pi[b,g] = lambda (0.7 in the reference) if attr[b]==g else (1-lambda)/(G-1)
 *   attr == NULL: n_samples must be 1 and pi = 1/G
 *   s_eff[b,:] = sum_g pi[b,g] S[g,:] (+ S_global when not NULL)
 * attr is int64 [n_samples] ON THE DEVICE (the host mirror copies the reference's CPU tensor once per step).
 */
int ffm_seff(const long long* attr, const float* S, const float* S_global, float* s_eff, int n_samples, int G, int r,
             float lambda, ffm_stream_t stream);

/* Transpose of ffm_seff: dS[g,:] = sum_b pi[b,g] ds_eff[b,:]; dS_global = sum_b ds_eff[b,:] (optional). */
int ffm_ds(const long long* attr, const float* ds_eff, float* dS, float* dS_global, int n_samples, int G, int r,
           float lambda, ffm_stream_t stream);

/* ------------------------------------------------------ GLP_OT head -------------------------- */
#define FFM_OT_NONE 0
#define FFM_OT_SINKHORN 1
#define FFM_OT_COT 2

size_t ffm_ot_head_workspace_bytes(int M, int Bp, int D, int n_prompts, int n_cls);

/*
 * Forward of the GLP_OT head — trainers/GLP_OT_SVLoRA.py:713-757 with Sinkhorn :615-634 and
 * entropic_COT_fast :636-675.
 *
 *   img  [M+1, Bp, D] bf16 or f32 (row 0 = pooled token, skipped like :696-697), txt [n_prompts, n_cls, D] f32
 *   sim[b*n_cls+c, m, n] = <img_n[m+1,b,:], txt_n[n,c,:]>      (both L2-normalised, F.normalize eps 1e-12)
 *   K = exp(-(1-sim)/eps);  T = Sinkhorn(K, 1/M, 1/n_prompts) | COT(...) | 1
 *   sim_op[b*n_cls+c] = sum(T*sim) (mean for OT=None);  logits[b,c] = exp(logit_scale) * mean_slices(sim_op)
 *
 *   Outputs: logits [Bp/num_slices, n_cls] f32; T_out [Bp*n_cls, M, n_prompts] f32 (may be NULL when mode = NONE);
 *   status_out int32[2] = {iterations run, 1 if T contains NaN (reference returns None :738-743)}.
 *   img_is_bf16 selects the storage type of img.  img_batch_first = 1: img (and d_img of the backward) is stored
 *   batch-first [Bp, M+1, D] (the layout fairfedmed_b200's ViT tower keeps internally) instead of the reference's
 *   sequence-first [M+1, Bp, D] — same values, no transposition pass.  8 <= D <= 1024, D % 8 == 0.
 */
int ffm_ot_head_fwd(const void* img, int img_is_bf16, int img_batch_first, const float* txt, const float* logit_scale,
                    float* logits, float* T_out, float* sim_out, float* inv_norm_out, int32_t* status_out,
                    void* workspace, size_t workspace_bytes, int M, int Bp, int D, int n_prompts, int n_cls,
                    int num_slices, int mode, float eps, float thresh, int max_iter, float top_percent,
                    ffm_stream_t stream);

/*
 * Backward of the head. The transport plan is a constant (computed under no_grad, :734), so gradients flow
 * through sim only:  d_sim = T * d_sim_op (or 1/(M*n_prompts) for OT=None), then through both normalisations.
 *   d_img [M+1, Bp, D] (same storage type as img; row 0 receives zeros), d_txt [n_prompts, n_cls, D] f32,
 *   d_logit_scale f32[1].
 */
int ffm_ot_head_bwd(const void* img, int img_is_bf16, int img_batch_first, const float* txt, const float* logit_scale,
                    const float* d_logits, const float* T_plan, const float* sim, const float* inv_norm, void* d_img,
                    float* d_txt, float* d_logit_scale, void* workspace, size_t workspace_bytes, int M, int Bp, int D,
                    int n_prompts, int n_cls, int num_slices, int mode, ffm_stream_t stream);

/*
 * Stand-alone Sinkhorn / COT on a pre-built kernel matrix K [P, M, N] f32 (u = 1/M, v = v_mass/N rows):
 * the persistent kernel used by the head, exposed for parity tests and the bandwidth stress shape.
 * status_out as above.
 */
size_t ffm_sinkhorn_workspace_bytes(int P, int M, int N);
int ffm_sinkhorn(const float* Kmat, float* T_out, int32_t* status_out, void* workspace, size_t workspace_bytes, int P,
                 int M, int N, int mode, float v_mass, float thresh, int max_iter, ffm_stream_t stream);

/* ------------------------------------------------- federated aggregation --------------------- */
/*
 * Server aggregation — utils/fed_utils.py:42-100 (average_weights_EMA) and :6-40 (average_weights).
 * Clients are sharded one per rank; each rank scales its flat f32 parameter buffer locally, the host mirror
 * all-reduces (NCCL sum) the result, and the epilogue applies shared-half-S and the EMA with the previous global.
 *
 *   seg_kind[i] / seg_off[i] / seg_len[i] describe contiguous segments of the flat buffer:
 *     kind 0: ordinary tensor, weight = w_scalar (= n_k / sum n_k, 0 for unselected clients)
 *     kind 1: lora_S tensor [G, r] with shape[0]==G: row g weighted by w_group[g] (= n_{k,g} / sum_k n_{k,g})
 */
int ffm_fedavg_scale(const float* flat_in, float* flat_out, const int32_t* seg_kind, const int64_t* seg_off,
                     const int64_t* seg_len, int n_seg, int64_t n_elem, float w_scalar, const float* w_group, int G,
                     int r, ffm_stream_t stream);

/*
 * out = (1-beta_decay) * shared_half(avg) + beta_decay * prev_global; shared_half replaces the first r/2
 * columns of every kind-1 segment by their mean over the G rows (utils/fed_utils.py:90-98) when enabled.
 */
int ffm_fedavg_epilogue(const float* avg, const float* prev_global, float* out, const int32_t* seg_kind,
                        const int64_t* seg_off, const int64_t* seg_len, int n_seg, int64_t n_elem, float beta_decay,
                        int shared_half_s, int G, int r, ffm_stream_t stream);

/* --------------------------------------------------- fairness metrics ------------------------ */
/*
 * Group-wise rank statistics — evaluation/metrics.py:340-356 (compute_auc), :513-547 (equity_scaled_AUC),
 * :486-511 (equity_scaled_accuracy), :248-292 (DPD / EOD / AOD inputs).
 *
 *   prob [N, 2] f32 (softmax), label int32 [N] in {0,1}, attrs int32 [n_attr, N] (-1 = unknown).
 *   For every attribute a and group id g in [-1, max_groups) plus the "overall" slot, computes per probability
 *   column c in {0,1} the tie-aware Mann–Whitney counts of the one-vs-rest problem (positives = label==c):
 *       gt[c]  = #{(i,j): label_i==c, label_j!=c, prob[i,c] >  prob[j,c]}
 *       eq[c]  = #{(i,j): label_i==c, label_j!=c, prob[i,c] == prob[j,c]}
 *   (AUC_c = (gt + eq/2)/(P*Nn), exactly sklearn's roc_auc_score on stored f32 values) and the confusion
 *   counts of pred = argmax(prob): tp, fp, tn, fn.
 *
 *   counts_out: uint64 [n_slots, 8] = {gt0, eq0, gt1, eq1, tp, fp, tn, fn}, n_slots = 1 + n_attr*(max_groups+1);
 *   slot 0 = overall, slot 1 + a*(max_groups+1) + (g+1) = (attribute a, group g).
 */
size_t ffm_group_auc_workspace_bytes(int N, int n_attr, int max_groups);
int ffm_group_auc(const float* prob, const int32_t* label, const int32_t* attrs, uint64_t* counts_out,
                  void* workspace, size_t workspace_bytes, int N, int n_attr, int max_groups, ffm_stream_t stream);

/* ------------------------------------------------------ training utils ----------------------- */
/*
 * Fused SGD over a flat f32 parameter buffer (torch.optim.SGD semantics, momentum/weight decay/no nesterov),
 * applied `n_steps` times with the SAME gradient: the reference registers one optimizer under two model names,
 * so optimizer.step() runs twice per iteration (trainers/GLP_OT_SVLoRA.py:864-871,
 * Dassl/dassl/engine/trainer.py:333-342).  first_step != 0 initialises the momentum buffer with the gradient.
 */
int ffm_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                 float weight_decay, int n_steps, int first_step, ffm_stream_t stream);

/*
 * ffm_sgd_step with the learning rate read from DEVICE memory (`lr_dev`, one float) and a zero-initialised momentum
 * buffer standing in for torch's lazily created one (momentum * 0 + d == d, so no first-step flag): the form a captured
 * CUDA graph of the training step uses — the StepLR schedule (Dassl/dassl/optim/lr_scheduler.py) then only rewrites
 * that float.
 */
int ffm_sgd_step_dev_lr(float* param, const float* grad, float* momentum_buf, int64_t n, const float* lr_dev,
                        float momentum, float weight_decay, int n_steps, ffm_stream_t stream);

#ifdef __cplusplus
}
#endif

#endif /* FFM_B200_H_ */
