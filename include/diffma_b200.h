/*
 * diffma_b200 C-ABI -- the drop-in boundary of the B200-native DiffMa hot path.
 *
 * What it replaces.  The reference (wongzbb/DiffMa-Diffusion-Mamba) reaches its native code through
 * the Python import surface of two third-party wheels:
 *
 *   block/mamba.py:11     from mamba_ssm.ops.selective_scan_interface import selective_scan_fn, mamba_inner_fn
 *   block/mamba.py:13     from causal_conv1d import causal_conv1d_fn, causal_conv1d_update
 *   block/mamba2.py:17    from mamba_ssm.ops.triton.layernorm_gated import RMSNorm as RMSNormGated
 *   block/mamba2.py:20-21 from mamba_ssm.ops.triton.ssd_combined import mamba_chunk_scan_combined,
 *                                                                        mamba_split_conv1d_scan_combined
 *
 * Beneath those wrappers upstream binds two pybind11 modules, `selective_scan_cuda.{fwd,bwd}` and
 * `causal_conv1d_cuda.{causal_conv1d_fwd,causal_conv1d_bwd}`, plus Triton kernels for Mamba-2.  The
 * entry points below are what a binding for this path uses instead: plain pointers and sizes, no torch
 * types, no exceptions.  INTEGRATION.md shows the ctypes stub on the reference side.
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative dm_status otherwise; never throw, never exit;
 *   - never allocate, never synchronise, never copy host<->device: the caller owns every buffer
 *     (scratch included, passed as named fields) and the work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so every call is CUDA-graph capturable;
 *   - re-entrant; the only global state is one-time cudaFuncSetAttribute per kernel;
 *   - activations are "tokens-major": element (b, token, channel) of a tensor lives at
 *     base + b*batch_stride + token*token_stride + channel  (strides in ELEMENTS, channel stride 1).
 *     Callers holding the reference's (B, C, L) layout transpose first (the Python shim does);
 *   - activation pointers must be 16-byte aligned and their strides multiples of 16 bytes.
 */
#ifndef DIFFMA_B200_H
#define DIFFMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_ABI_VERSION 6

typedef enum {
    DM_OK = 0,
    DM_ERR_INVALID_ARG = -1,   /* null pointer, bad size, misaligned pointer/stride                      */
    DM_ERR_UNSUPPORTED = -2,   /* shape/dtype combination this build has no kernel for                   */
    DM_ERR_CUDA = -4           /* a CUDA runtime call failed; see dm_last_cuda_error()                   */
} dm_status;

typedef enum { DM_F32 = 0, DM_BF16 = 1 } dm_dtype;

#define DM_MAX_GROUPS 4

/* Row placement of per-direction outputs. */
#define DM_OUT_SCAN_ORDER 0    /* row j of direction k  = j-th scanned token  (what mamba_inner_fn returns) */
#define DM_OUT_TOKEN_ORDER 1   /* row order[k][j]: already un-permuted, so merging directions is a plain sum */

/* ------------------------------------------------------------------------------------------------------
 * Mamba-1 forward: causal conv1d + SiLU -> x_proj -> dt_proj -> softplus -> selective scan -> D skip ->
 * SiLU(z) gate.  Replaces, for one or several directions in one call, the part of upstream
 * `MambaInnerFn.forward` between the in-projection and the out-projection (reference call sites
 * block/mamba.py:346-393): causal_conv1d_cuda.causal_conv1d_fwd + two cuBLAS GEMMs +
 * selective_scan_cuda.fwd, and the CrossScan / CrossMerge gathers around them (block/mamba.py:32-82).
 *
 * One "group" = one mixer (its own weights and activations).  A Spiral block has two mixers on two
 * inputs (block/mamba_block.py:107-108); passing both as groups runs them in one launch.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    /* activations in */
    const void* xz;            /* (B, L_src, 2*d_inner) act dtype: x = channels [0,d_inner), z = [d_inner,2*d_inner) */
    int64_t xz_batch_stride, xz_token_stride;
    /* activations out */
    void* out;                 /* y * silu(z), act dtype; element (b,k,row,c) at
                                  b*out_batch_stride + k*out_dir_stride + row*out_token_stride + c        */
    int64_t out_batch_stride, out_dir_stride, out_token_stride;
    /* intermediates, written by the forward, kept by the caller for the backward (or reused as scratch) */
    void* u;                   /* (B, n_dir, seqlen, d_inner) act dtype, contiguous: silu(conv1d(x)) in scan order */
    float* x_dbl;              /* (B, n_dir, seqlen, dt_rank + 2*d_state) fp32, contiguous: [dt_low | B | C]        */
    /* weights */
    const float* conv_weight;  /* (d_inner, d_conv) fp32   -- conv1d.weight viewed (d, w)                 */
    const float* conv_bias;    /* (d_inner) fp32 or NULL                                                  */
    const void* x_proj_weight; /* (dt_rank + 2*d_state, d_inner) act dtype, row-major                     */
    const void* dt_proj_weight;/* (d_inner, dt_rank) act dtype, row-major                                 */
    const float* dt_bias;      /* (d_inner) fp32 or NULL  -- `delta_bias`                                 */
    const float* A;            /* (d_inner, d_state) fp32 -- already -exp(A_log)                          */
    const float* D;            /* (d_inner) fp32 or NULL                                                  */
    /* training only, or NULL: the forward also stores the recurrence state BEFORE every c-th scanned token,
     * c = dm_mamba1_bwd_chunk_tokens(): (B, n_dir, ceil(seqlen/c), d_inner, d_state) fp32, contiguous.  Handed to
     * dm_mamba1_scan_bwd as `state_workspace` with `states_valid = 1` it saves the backward its forward sweep. */
    float* chunk_states;
    /* optional, or NULL: (B, n_dir, seqlen, d_inner) fp16, contiguous, scan order: delta = softplus(dt_proj(dt_low) + bias).
     * When given (bf16 activations), the scan kernel reads it instead of evaluating dt_proj + softplus itself. */
    void* delta;
} dm_mamba1_group;

typedef struct {
    int32_t batch;             /* B                                                                       */
    int32_t n_dir;             /* scanned sequences per batch element (spiral 3, zig 1, vim 2, vmamba/eff 4) */
    int32_t seqlen;            /* tokens per scanned sequence (L, or L/4 for the EfficientVMamba split)    */
    int32_t d_inner, d_state, dt_rank, d_conv;
    int32_t act_dtype;         /* dm_dtype of xz / out / u / x_proj / dt_proj                             */
    int32_t out_order;         /* DM_OUT_SCAN_ORDER or DM_OUT_TOKEN_ORDER                                 */
    int32_t n_groups;          /* 1..DM_MAX_GROUPS                                                        */
    const int32_t* order;      /* device (n_dir, seqlen) gather table: scanned token j of direction k is
                                  source token order[k*seqlen+j]; NULL = every direction is the identity.
                                  A first entry order[k*seqlen] == -1 marks direction k as the identity.   */
    dm_mamba1_group group[DM_MAX_GROUPS];
    /* Optional scratch of the scan kernel's dynamic schedule (persistent warps consuming a ready queue of ~40-token
     * segments; the recurrence state crosses segments through this buffer).  NULL = static schedule, one warp per
     * (sequence, 64 channels).  At least dm_mamba1_sched_workspace_bytes(...) bytes, 16-byte aligned, ZEROED ONCE by
     * the caller when allocated: every launch leaves it re-armed.  One workspace must not be used by launches that
     * can run concurrently (different streams). */
    void* sched_workspace;
    int64_t sched_workspace_bytes;
    /* Non-zero: the z half of `xz` already holds silu(z) (the in-projection's epilogue applied it once per source
     * token, dm_gemm_bf16_tn_ex `silu_from`), so the scan multiplies by it instead of evaluating SiLU once per
     * direction.  Inference only: the backward needs the raw z and rejects this flag. */
    int32_t z_is_gated;
    int32_t reserved_;
} dm_mamba1_args;

int dm_mamba1_scan_fwd(const dm_mamba1_args* args, void* stream);

/* Size of `sched_workspace` for a launch of n_groups x batch x n_dir sequences of d_inner channels (0 = none). */
int64_t dm_mamba1_sched_workspace_bytes(int32_t batch, int32_t n_dir, int32_t d_inner, int32_t n_groups);

/* The two kernels of dm_mamba1_scan_fwd separately, for profiling and the backward's recomputation:
 * phase 1 = gather + conv1d + SiLU + x_proj (writes u, x_dbl); phase 2 = dt_proj + scan + gate (reads them). */
int dm_mamba1_scan_phase(const dm_mamba1_args* args, int phase, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Mamba-1 backward of dm_mamba1_scan_fwd (upstream `MambaInnerFn.backward` minus the out-projection:
 * selective_scan_cuda.bwd + causal_conv1d_cuda.causal_conv1d_bwd; reached from reference train.py:259).
 * `args` is the forward's struct (xz, u, x_dbl as the forward wrote them, weights, order table, and the out_*
 * strides, which here describe `dout`).  All gradient buffers are fp32, contiguous, in SCAN order; buffers marked
 * "accumulated" must be zeroed by the caller.  The GEMM-shaped gradients (through x_proj / dt_proj) are the host's
 * job between the two phases:
 *   phase 1  reverse scan: needs dout; writes d_xz_scan[..., d_inner:] (dz), du (scan part), ddelta (d delta_raw),
 *            accumulates d_x_dbl[..., dt_rank:] (dB, dC), dA, dD, d_dt_bias; uses state_workspace.
 *   host     d_x_dbl[..., :dt_rank] = ddelta . W_dt ; dW_dt = ddelta^T . dt_low ; du += d_x_dbl . W_x ; dW_x = d_x_dbl^T . u
 *   phase 2  conv backward: reads du (total), writes d_xz_scan[..., :d_inner] (dx), accumulates d_conv_weight / bias.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* dout;          /* gradient of `out` (act dtype), addressed with args->group[g].out_*_stride            */
    float* d_xz_scan;          /* (B, n_dir, seqlen, 2*d_inner): [dx | dz] per scanned token                           */
    float* du;                 /* (B, n_dir, seqlen, d_inner)                                                         */
    float* ddelta;             /* (B, n_dir, seqlen, d_inner)                                                         */
    float* d_x_dbl;            /* (B, n_dir, seqlen, dt_rank + 2*d_state), accumulated: [d dt_low | dB | dC]           */
    float* dA;                 /* (d_inner, d_state), accumulated                                                     */
    float* dD;                 /* (d_inner), accumulated, or NULL                                                     */
    float* d_dt_bias;          /* (d_inner), accumulated, or NULL                                                     */
    float* state_workspace;    /* (B, n_dir, ceil(seqlen/c), d_inner, d_state) scratch, c = dm_mamba1_bwd_chunk_tokens() */
    float* d_conv_weight;      /* (d_inner, d_conv), accumulated                                                      */
    float* d_conv_bias;        /* (d_inner), accumulated, or NULL                                                     */
    int64_t states_valid;      /* 1: state_workspace holds the forward's `chunk_states`; 0: scratch, the kernel fills it */
} dm_mamba1_bwd_group;

int dm_mamba1_scan_bwd(const dm_mamba1_args* args, const dm_mamba1_bwd_group* grads /* [n_groups] */, int phase,
                       void* stream);
int dm_mamba1_bwd_chunk_tokens(void);     /* tokens between two saved states of `state_workspace` (4 in this build) */

/* ------------------------------------------------------------------------------------------------------
 * Mamba-2 forward: split [z | x | B | C | dt] -> causal conv1d + SiLU over [x|B|C] -> softplus(dt + bias)
 * -> SSD state recurrence S_t = exp(dt A_h) S_{t-1} + dt x_t (x) B_t,  y_t = S_t C_t + D_h x_t
 * -> gate v = y * silu(z), and the per-token sum of squares of v that the gated RMSNorm needs.
 * Replaces the part of upstream `mamba_split_conv1d_scan_combined` before the RMSNorm scale and the
 * out-projection (reference call sites block/mamba2.py:392-696): causal_conv1d_fwd + the five Triton SSD
 * kernels + the first pass of `_layer_norm_fwd_1pass_kernel`.  ngroups = 1, norm_before_gate = False.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* zxbcdt;        /* (B, L_src, 2*d_inner + 2*d_state + nheads) act dtype                    */
    int64_t in_batch_stride, in_token_stride;
    void* out;                 /* v = y * silu(z), act dtype, addressed like dm_mamba1_group.out           */
    int64_t out_batch_stride, out_dir_stride, out_token_stride;
    float* sumsq;              /* fp32, ZEROED BY THE CALLER, or NULL to skip: sum_c v[b,k,row,c]^2 (fp32 v, before
                                  rounding), one value per OUTPUT row, at b*sumsq_batch_stride +
                                  k*sumsq_dir_stride + row (row as in `out`: scanned index or source token) */
    int64_t sumsq_batch_stride, sumsq_dir_stride;
    const float* conv_weight;  /* (d_inner + 2*d_state, d_conv) fp32                                      */
    const float* conv_bias;    /* (d_inner + 2*d_state) fp32 or NULL                                      */
    const float* dt_bias;      /* (nheads) fp32 or NULL                                                   */
    const float* A;            /* (nheads) fp32 -- already -exp(A_log)                                    */
    const float* D;            /* (nheads) fp32 or NULL                                                   */
} dm_mamba2_group;

typedef struct {
    int32_t batch, n_dir, seqlen;
    int32_t d_inner, d_state, nheads, d_conv;     /* headdim = d_inner / nheads                           */
    int32_t act_dtype, out_order, n_groups;
    int32_t gate;              /* 1: out = y * silu(z); 0: out = y (caller gates)                          */
    const int32_t* order;      /* as in dm_mamba1_args                                                    */
    dm_mamba2_group group[DM_MAX_GROUPS];
} dm_mamba2_args;

int dm_mamba2_ssd_fwd(const dm_mamba2_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Mamba-2 backward of dm_mamba2_ssd_fwd (upstream MambaSplitConv1dScanCombinedFn.backward minus the RMSNorm scale and the
 * out-projection; reference block/mamba2.py:392 differentiated from train.py:258-259 with --use-mamba2).  The SSD
 * recurrence is the S6 recurrence with A[d, n] = A_head(d), delta[d] = dt_head(d): the reverse scan and the conv backward
 * of the x channels are dm_mamba1_scan_bwd on SSD operands; this entry point prepares those operands and finishes the
 * B / C channels.  `args` is the forward's struct; all buffers contiguous, scan order.
 *   phase 0  conv1d + SiLU of x | B | C in scan order -> u, and x_dbl rows in dm_mamba1's format [dt hi | dt lo | B | C]
 *            (raw dt of head h in dt_low slot h; the S6 view's dt_proj is the one-hot head map (d_inner, 32));
 *   then     dm_mamba1_scan_bwd phase 1 with xz = zxbcdt - d_inner (z at offset d_inner), per-channel A / D / dt_bias,
 *            and phase 2 with xz = zxbcdt + d_inner (x at offset 0), conv_weight = rows [0, d_inner) of the Mamba-2 conv;
 *   phase 2  conv backward of B | C from d_x_dbl[..., 32:64] -> d_bc, accumulates d_conv_weight / bias rows >= d_inner.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    void* u;                   /* (B, n_dir, seqlen, d_inner) act dtype                      [phase 0 writes]          */
    float* x_dbl;              /* (B, n_dir, seqlen, 64) fp32                                [phase 0 writes]          */
    const float* d_x_dbl;      /* (B, n_dir, seqlen, 64) fp32 from the reverse scan          [phase 2 reads]           */
    float* d_bc;               /* (B, n_dir, seqlen, 2*d_state) fp32: gradient of raw B | C  [phase 2 writes]          */
    float* d_conv_weight;      /* (d_inner + 2*d_state, d_conv) fp32, accumulated            [phase 2: rows >= d_inner] */
    float* d_conv_bias;        /* (d_inner + 2*d_state) fp32, accumulated, or NULL                                      */
} dm_mamba2_bwd_group;
int dm_mamba2_ssd_bwd(const dm_mamba2_args* args, const dm_mamba2_bwd_group* grads /* [n_groups] */, int phase,
                      void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Row-wise glue of Spiral_MambaBlock.forward (reference block/mamba_block.py:100-115), one warp per token row.
 * x / skip / out are the fp32 residual stream (rows = batch*seqlen, d_model = 512 in this build);
 * `mod` is the adaLN output (batch, 3*d_model) = [shift | scale | gate] with row stride mod_batch_stride.
 *
 * dm_spiral_pre      out2[0] = modulate(LayerNorm(x + skip)), out2[1] = out2[0] * w[row]   (act dtype, (2, rows, d))
 *                    -- lines 101-105, plus the long-skip add of model.py:290-292 when skip != NULL; w may be NULL.
 * dm_spiral_post_ln  out = LayerNorm(cat(ab[0], ab[1])) (rows, 2*d), act dtype            -- attention_network[0]
 * dm_spiral_post_mix alpha = sigmoid(w3 . silu(hidden) + b3); out = (x + skip) + gate * (alpha*a + (1-alpha)*b)
 *                    -- attention_network[2..4] and lines 112-114; `hidden` = attention_network[1] output (rows, d).
 * ---------------------------------------------------------------------------------------------------- */
int dm_spiral_pre(const float* x, const float* skip, const float* ln_weight, const float* ln_bias, const float* mod,
                  int64_t mod_batch_stride, const float* w, void* out2, int32_t batch, int32_t seqlen,
                  int32_t d_model, float eps, int32_t act_dtype, void* stream);
int dm_spiral_post_ln(const void* ab, const float* ln_weight, const float* ln_bias, void* out, int32_t rows,
                      int32_t d_model, float eps, int32_t act_dtype, void* stream);
int dm_spiral_post_mix(const float* x, const float* skip, const void* ab, const void* hidden, const float* w3,
                       const float* b3, const float* mod, int64_t mod_batch_stride, float* out, int32_t batch,
                       int32_t seqlen, int32_t d_model, int32_t act_dtype, void* stream);
/* dm_spiral_post_mix of block i and dm_spiral_pre of block i+1 (or of the final layer: ln_weight = 1, ln_bias = 0,
 * eps 1e-6, w = NULL) in one pass over the row: x_out as dm_spiral_post_mix's `out`, out2 as dm_spiral_pre's `out2`
 * computed from x_out (+ skip_next).  Saves one launch and one read of the residual stream per block boundary. */
int dm_spiral_post_mix_pre(const float* x, const float* skip, const void* ab, const void* hidden, const float* w3,
                           const float* b3, const float* mod, int64_t mod_batch_stride, float* x_out,
                           const float* skip_next, const float* ln_weight, const float* ln_bias, const float* mod_next,
                           int64_t mod_next_batch_stride, const float* w, void* out2, int32_t batch, int32_t seqlen,
                           int32_t d_model, float eps, int32_t act_dtype, void* stream);

/* The same two operations with the attention network's LayerNorm FOLDED around its Linear (reference
 * block/mamba_block.py:110: attention_network = LayerNorm(2D) -> Linear(2D, D) -> SiLU -> Linear(D, 1) -> Sigmoid):
 *     Linear(LN(cat(a, b)))[n] = rstd * (a . W'_a[n] + b . W'_b[n] - mean * colsum[n]) + cvec[n],
 *     W' = W * gamma (per input column), colsum[n] = sum_k W'[n][k], cvec[n] = sum_k beta[k] W[n][k] + bias[n],
 * so the caller runs the GEMM on the raw out-projection results -- g2[0] = a W'_a^T, g2[1] = b W'_b^T, (2, rows, D), fp32 or
 * the activation dtype -- and this kernel, which reads a and b anyway, supplies the row's mean and rstd: dm_spiral_post_ln and
 * its (rows, 2D) round trip disappear.  out2 == NULL: plain post_mix (writes x_out only); otherwise post_mix + pre of the
 * next block as dm_spiral_post_mix_pre.  Inference only (no adjoint is provided for this form). */
typedef struct dm_spiral_fold_args {
    const float* x; const float* skip;        /* (B, L, D) fp32; skip may be NULL                                      */
    const void* ab;                           /* (2, B*L, D) activation dtype                                          */
    const void* g2; int32_t g2_dtype;         /* (2, B*L, D); DM_F32 or the activation dtype                           */
    int32_t act_dtype;
    const float* colsum; const float* cvec;   /* (D) fp32 each                                                         */
    const float* w3; const float* b3;         /* attention_network[3]: (D), (1) fp32                                   */
    const float* mod; int64_t mod_batch_stride;
    float* x_out;
    const float* skip_next; const float* ln_weight; const float* ln_bias;        /* pre part, as dm_spiral_post_mix_pre */
    const float* mod_next; int64_t mod_next_batch_stride; const float* w;
    void* out2;                               /* NULL: no pre part                                                     */
    int32_t batch, seqlen, d_model;
    float eps, ln2_eps;                       /* eps of the next block's LayerNorm / of the attention LayerNorm        */
} dm_spiral_fold_args;
int dm_spiral_post_mix_fold(const dm_spiral_fold_args* args, void* stream);

/* Head of DiffMa.forward in one launch (reference model.py:264-281, visionEmbedding / PatchEmbed, TimestepEmbedder):
 *   h (B, L, D) fp32        = PatchEmbed(x) + pos_embed: x (B, C, S, S) fp32, patch_weight (C*p*p, D) fp32 (the conv weight
 *                             unfolded and transposed), pos_bias (L, D) fp32 = pos_embed + conv bias
 *   silu_c (B, 2D) act dtype = silu(cat(t_table[t] + y, t_table[t] + y2_mean)): what every adaLN Linear consumes; t_table
 *                             (table_rows, D) fp32 = TimestepEmbedder evaluated at the integer steps, t int64 (a step outside
 *                             the table poisons its row with NaN), y / y2_mean (B, D) fp32 (y2: see y2_tokens).
 * D = 512. */
int dm_step_head(const float* x, const float* patch_weight, const float* pos_bias, float* h, int32_t batch, int32_t channels,
                 int32_t image_size, int32_t patch, const int64_t* t, const float* t_table, int32_t table_rows, const float* y,
                 const float* y2_mean, int32_t y2_tokens, void* silu_c, int32_t d_model, int32_t act_dtype, void* stream);
/* y2_tokens <= 1: y2_mean is the pooled (B, D) tensor; y2_tokens = T > 1: y2_mean points at the un-pooled (B, T, D) tensor and
 * the kernel takes the token mean itself (reference model.py:276 `torch.mean(y2, dim=1)`). */

/* Tail of DiffMa.forward (reference model.py:295-301): FinalLayer.linear on the modulated rows hn (B*L, D) and unpatchify in one
 * launch: out (B, out_channels, S, S), S = grid_side * patch, out[b][c][gy p + py][gx p + px] = (hn . W^T + bias)[(py p + px) *
 * out_channels + c] of token (gy, gx).  bf16 only, D = 512, patch * patch * out_channels <= 128 (else DM_ERR_UNSUPPORTED:
 * the caller runs the GEMM and the permute). */
int dm_final_linear_unpatchify(const void* hn, const void* weight, const void* bias, void* out, int32_t batch,
                               int32_t grid_side, int32_t patch, int32_t out_channels, int32_t d_model, int32_t act_dtype,
                               void* stream);

/* Adjoints of the three row kernels above for the training step (autograd of block/mamba_block.py:100-115, reached from
 * train.py:259).  Per-batch-element and per-parameter gradients are ACCUMULATED (atomics) into buffers the caller zeroed:
 *   dm_spiral_pre_bwd       d_out2 (2, rows, d) act dtype -> dx (rows, d) fp32 [= d skip]; d_mod[:, 0:d] += d shift,
 *                           d_mod[:, d:2d] += d scale; d_ln_weight / d_ln_bias (d) += .  w == NULL: mask of ones.
 *   dm_spiral_post_mix_bwd  d_x_out (rows, d) fp32 [= d x = d skip] -> d_ab (2, rows, d) and d_hidden (rows, d) act dtype
 *                           (overwritten); d_mod[:, 2d:3d] += d gate; d_w3 (d) +=, d_b3 (1) += .
 *   dm_spiral_post_ln_bwd   d_out (rows, 2d) act dtype -> d_ab += LayerNorm backward; d_ln_weight / d_ln_bias (2d) += . */
int dm_spiral_pre_bwd(const float* x, const float* skip, const float* ln_weight, const float* ln_bias, const float* mod,
                      int64_t mod_batch_stride, const float* w, const void* d_out2, float* dx, float* d_mod,
                      int64_t d_mod_batch_stride, float* d_ln_weight, float* d_ln_bias, int32_t batch, int32_t seqlen,
                      int32_t d_model, float eps, int32_t act_dtype, void* stream);
int dm_spiral_post_mix_bwd(const float* d_x_out, const void* ab, const void* hidden, const float* w3, const float* b3,
                           const float* mod, int64_t mod_batch_stride, void* d_ab, void* d_hidden, float* d_mod,
                           int64_t d_mod_batch_stride, float* d_w3, float* d_b3, int32_t batch, int32_t seqlen,
                           int32_t d_model, int32_t act_dtype, void* stream);
int dm_spiral_post_ln_bwd(const void* ab, const float* ln_weight, const void* d_out, void* d_ab, float* d_ln_weight,
                          float* d_ln_bias, int32_t batch, int32_t seqlen, int32_t d_model, float eps, int32_t act_dtype,
                          void* stream);

/* Adjoint of the CrossScan gather (reference block/mamba.py:48-57): dst[r][l][:] = sum_k src[r][index[l*n_dir + k]][:].
 * src (n_groups, rows_per_group, channels) fp32 scan-order rows, index (src_len * n_dir) int32 row numbers inside a group,
 * dst (n_groups, src_len, channels) act dtype.  channels % 8 == 0. */
int dm_merge_directions(const float* src, const int32_t* index, void* dst, int32_t n_groups, int32_t src_len,
                        int32_t n_dir, int32_t rows_per_group, int32_t channels, int32_t act_dtype, void* stream);
/* The same with the output row assembled from up to 4 column segments living in different fp32 tensors (segment s: base
 * pointer, columns (multiple of 8), source row stride in elements): [dz | dx | dB dC | d dt] of the Mamba-2 backward
 * without a concatenated (B, n_dir, seqlen, 2096) intermediate. */
typedef struct { const float* src; int32_t channels; int32_t row_stride; } dm_merge_segment;
int dm_merge_directions_multi(const dm_merge_segment* segments, int32_t n_segments, const int32_t* index, void* dst,
                              int32_t n_groups, int32_t src_len, int32_t n_dir, int32_t rows_per_group, int32_t act_dtype,
                              void* stream);

/* ------------------------------------------------------------------------------------------------------
 * One reverse-diffusion update as one elementwise kernel (reference diffusion/gaussian_diffusion.py p_mean_variance
 * :254-332 with LEARNED_RANGE variance + epsilon prediction, and p_sample :376-417):
 *   model_out (N, 2C, H, W) fp32 = [eps | v], x (N, C, H, W), noise like x, t (N) int64 step indices,
 *   table (9, n_steps) fp32 rows: sqrt_acp, sqrt_1m_acp, sqrt_recip_acp, sqrt_recipm1_acp, posterior_variance,
 *   posterior_log_variance_clipped, posterior_mean_coef1, posterior_mean_coef2, log_betas.
 *   sample = mean + [t != 0] * exp(0.5*logvar) * noise ; pred_xstart optional (may be NULL).  chw = C*H*W.
 * ---------------------------------------------------------------------------------------------------- */
int dm_p_sample_update(const float* model_out, const float* x, const float* noise, const float* table, const int64_t* t,
                       float* sample, float* pred_xstart, int32_t batch, int32_t chw, int32_t n_steps,
                       int32_t clip_denoised, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Batched bf16 GEMM on tcgen05 / TMEM / TMA:  C[g] = rowscale[g] (.) (A[g] . B[g]^T), fp32 accumulation.
 *   A (groups, M, K), B (groups, N, K), C (groups, M, N), all bf16, K (resp. N) contiguous, strides in elements and
 *   multiples of 8; row_scale (groups, M) fp32 or NULL.  Replaces the cuBLAS calls behind the reference's
 *   in-projection (block/mamba.py:333-337, block/mamba2.py:382) and out-projection (inside mamba_inner_fn /
 *   mamba_split_conv1d_scan_combined); the row scale carries the soft mask  (x*w).W = w (.) (x.W).
 * ---------------------------------------------------------------------------------------------------- */
int dm_gemm_bf16_tn(const void* A, int64_t a_group_stride, int64_t a_row_stride, const void* B, int64_t b_group_stride,
                    int64_t b_row_stride, void* C, int64_t c_group_stride, int64_t c_row_stride, const float* row_scale,
                    int32_t groups, int32_t M, int32_t N, int32_t K, void* stream);

/* The same kernel with everything it fuses that a library GEMM cannot:
 *   C[g] = epilogue( (A_0[g] + ... + A_{n_sum-1}[g]) . B[g]^T ),  epilogue(v) = silu_{col >= silu_from}( row_scale * v + bias )
 *   - n_sum (1 or 3) slices of every A row, a_sum_stride elements apart, are summed in shared memory in front of the
 *     MMA: the CrossMerge direction sum (reference block/mamba.py:60-82) as the out-projection's A producer, so the
 *     contraction is K = d_inner instead of n_dir * d_inner and the merged tensor never exists in HBM;
 *   - bias (groups, N) fp32 or NULL: attention_network[1]'s bias (reference block/mamba_block.py:52-53);
 *   - silu_from (multiple of 32; >= N disables): SiLU on columns [silu_from, N): the in-projection emits silu(z) for
 *     the gate once per source token (see dm_mamba1_args.z_is_gated). */
typedef struct {
    const void* A;             /* (groups, M, n_sum, K) bf16 as strides: group, row, summed slice; K contiguous      */
    int64_t a_group_stride, a_row_stride, a_sum_stride;
    int32_t n_sum, reserved_;
    const void* B;             /* (groups, N, K) bf16, K contiguous                                                   */
    int64_t b_group_stride, b_row_stride;
    void* C;                   /* (groups, M, N) bf16, N contiguous                                                   */
    int64_t c_group_stride, c_row_stride;
    const float* row_scale;    /* (groups, M) fp32 or NULL                                                            */
    const float* bias;         /* (groups, N) fp32 or NULL                                                            */
    int32_t silu_from;
    int32_t groups, M, N, K;
} dm_gemm_args;
int dm_gemm_bf16_tn_ex(const dm_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Training-step tail as one pass over flat fp32 buffers: torch.optim.AdamW's update (reference train.py:201,262) and
 * the EMA update `ema = decay*ema + (1-decay)*param` (reference train.py:34-43,264) for `n` contiguous parameters.
 *   param *= 1 - lr*weight_decay ; m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; g = grad * grad_scale
 *   param -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) ; ema (may be NULL) as above, from the UPDATED param.
 * `step` points at a DEVICE float holding t (the number of steps including this one), so the call can be captured in
 * a CUDA graph and replayed while the caller increments the counter on the device.  Any sub-range of the flat
 * buffers may be passed (per gradient bucket, as soon as its all-reduce has landed).
 * ---------------------------------------------------------------------------------------------------- */
int dm_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, const float* step,
                      int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay, double ema_decay,
                      double grad_scale, void* stream);   /* hyper-parameters in double: 1 - beta etc. are formed before rounding */
/* Same update with the arguments in a struct and one more output: `shadow_bf16` (may be NULL, 8-byte aligned) receives
 * the UPDATED parameters rounded to bf16 -- the compute-dtype weights of the next autocast step, so the training loop
 * needs no per-tensor fp32 -> bf16 cast (reference: torch.autocast re-casts every weight in every forward, train.py:252). */
typedef struct dm_adamw_args {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
    float* ema;                    /* may be NULL */
    const float* step;             /* device scalar, as above */
    void* shadow_bf16;             /* may be NULL */
    int64_t n;
    double lr, beta1, beta2, eps, weight_decay, ema_decay, grad_scale;
} dm_adamw_args;
int dm_adamw_ema_step_ex(const dm_adamw_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------ */
int dm_version(void);                     /* DM_ABI_VERSION of the loaded library                        */
const char* dm_status_string(int status);
int dm_last_cuda_error(void);             /* cudaError_t of the last DM_ERR_CUDA on this thread           */
const char* dm_build_info(void);          /* "sm_100a nvcc 12.9 ..."                                      */

#ifdef __cplusplus
}
#endif
#endif /* DIFFMA_B200_H */
