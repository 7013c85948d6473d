/*
 * radialog_b200 — C-ABI of the B200-native RaDialog image->report hot path.
 *
 * The reference (ChantalMP/RaDialog) is pure Python/PyTorch and has no FFI; its boundary for this path is a set of
 * Python call surfaces.  This header is the C boundary a maintainer binds from Python
 * (ctypes stub in INTEGRATION.md); each entry point cites the reference code it replaces (paths relative to the
 * reference root).  Conventions:
 *   - every function returns 0 on success, <0 on error; rd_last_error() gives the message (thread-local);
 *   - all pointers named *_dev are device pointers owned by the caller (e.g. torch tensors' data_ptr());
 *   - `stream` is a cudaStream_t passed as void*; no function synchronises unless it says so;
 *   - `dtype`: RD_F16 (reference dtype, test.py:289) or RD_BF16; accumulation is always fp32;
 *   - handles are not thread-safe, distinct handles are independent.
 */
#ifndef RADIALOG_B200_H
#define RADIALOG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RD_F16 0
#define RD_BF16 1

#define RD_OK 0
#define RD_ERR_INVALID -1
#define RD_ERR_CUDA -2
#define RD_ERR_UNSUPPORTED -3

const char* rd_last_error(void);
int rd_version(void);
/* 1 if device `dev` is sm_100 (B200); the library refuses to run anywhere else. */
int rd_device_ok(int dev);

/* ------------------------------------------------------------------------------------------------------------
 * Op level
 * ---------------------------------------------------------------------------------------------------------- */

/* Epilogue of rd_linear.  Order of operations (each "T(.)" is one rounding to the storage dtype):
 *   v = acc (+ bias[n])                                           fp32
 *   residual, res_mode 2 : v += residual[m,n]   (Q-Former / conv residuals, evaluated in fp32 by the reference)
 *   act == RD_ACT_NONE   : y = T(v)
 *   act == RD_ACT_RELU   : y = T(max(v,0))            torchvision Bottleneck / biovil_t/modules.py:45
 *   act == RD_ACT_GELU   : y = T(gelu_erf(v))         Qformer.py:358-361
 *   act == RD_ACT_SWIGLU : W has 2N rows (gate rows [0,N), up rows [N,2N)):
 *                          y = T( T(silu(T(acc_gate))) * T(acc_up) )      modeling_llama_imgemb.py:158-159
 *   lora_r > 0           : y = T( y + T(lora_scale * T(sum_r lora_t[m,r]*lora_B[n,r])) )   peft unmerged LoRA
 *   residual, res_mode 1 : y = T( residual[m,n] + y )       fp16 residual stream, modeling_llama_imgemb.py:302,308 */
#define RD_ACT_NONE 0
#define RD_ACT_RELU 1
#define RD_ACT_GELU 2
#define RD_ACT_SWIGLU 3

typedef struct rd_epilogue {
  const float* bias_dev;     /* [N] fp32 or NULL */
  const void* residual_dev;  /* [M, ld_res] storage dtype or NULL */
  int64_t ld_res;
  int res_mode;              /* 1 or 2, see above */
  int act;
  const void* lora_t_dev;    /* [M, lora_r] storage dtype: A.x already computed (rd_rmsnorm) */
  const void* lora_b_dev;    /* [N, lora_r] storage dtype */
  int lora_r;
  float lora_scale;
} rd_epilogue;

/* out[M,N] = epilogue(x[M,K] . W[N,K]^T); row-major, leading dimensions in elements (multiples of 8).
 * Replaces every nn.Linear / 1x1 conv on the path (q/k/v/o/gate/up/down/lm_head: modeling_llama_imgemb.py:153-159,
 * 178-181,680; Q-Former denses: Qformer.py:128-130,282,352,367; img_proj_layer: test.py:295).
 * `algo`: 0 auto (= tcgen05 tensor-core tiles at every M), 1 streaming CUDA-core GEMV (M<=4), 2 tcgen05,
 * 3 SIMT validation kernel.  `ws_dev` is split-K workspace (may be NULL when algo picks no split;
 * size from rd_linear_workspace_bytes).                                                                        */
int rd_linear(const void* x_dev, int64_t ldx, const void* w_dev, int64_t ldw, void* out_dev, int64_t ldo,
              int M, int N, int K, const rd_epilogue* epi, int dtype, int algo, void* ws_dev, int64_t ws_bytes,
              void* stream);
/* Split-K workspace: must be zero-initialised once by its owner (the kernels re-arm their counters themselves). */
int64_t rd_linear_workspace_bytes(int M, int N, int K);
/* Test hook: force the split-K factor of the tcgen05 path (0 = heuristic). */
int rd_linear_force_splits(int splits);
/* Split-K reduction: 0 (default) thread-block cluster + distributed shared memory when splits <= 8, 1 always through
 * the global fp32 workspace.  Both reduce in fixed split order (deterministic). */
int rd_linear_splitk_mode(int mode);
/* Decode tiles (M <= 32 tokens): 1 parks the weight k-blocks in tensor memory - the epilogue warps copy every weight
 * tile that lands in shared memory into a ring of TMEM slots (tcgen05.st) and tcgen05.mma reads its A operand from TMEM - so a
 * CTA buffers ~1.8x more weight bytes ahead of the activations it depends on; 0 (default: measured no faster) = both operands
 * from shared memory.  Same
 * products, same accumulation order: bit-identical results. */
int rd_linear_tmem_staging(int on);
/* Wide token counts (M > 128: prefill, convolutions, Q-Former): 1 (default) runs them on the persistent kernel of linear_wide.cu -
 * one CTA (pair) per SM walks the output tiles, two TMEM accumulator buffers so the epilogue of tile i (TMA residual load / TMA
 * store) overlaps the MMAs of tile i+1; 0 keeps every shape on the one-tile-per-CTA kernel.  Same products and k order either
 * way.  rd_linear_wide_min_tiles: GEMMs with fewer output tiles stay on the one-tile-per-CTA kernel (default 1 = none).
 * rd_linear_wide_force_nt: token-tile width of that kernel (multiple of 16, <= 256; 0 = by problem size). */
int rd_linear_wide_persistent(int on);
int rd_linear_wide_min_tiles(int n);
int rd_linear_wide_force_nt(int nt);
int rd_linear_wide_force_stages(int stages);   /* 2..6 pipeline stages (the rest of the 224 KB holds epilogue chunk buffers); 0 = by K */
/* 1 (default): the persistent kernel runs as CTA pairs - clusters of two CTAs issue one 256-row tcgen05.mma.cta_group::2 per
 * k-step, each CTA staging its own 128 weight rows and HALF of the shared token tile (a third less shared-memory and L2 traffic
 * per flop, 6 instead of 4 pipeline stages); 0: every CTA on its own.  Bit-identical results. */
int rd_linear_wide_pair(int on);
/* Convolution as an implicit GEMM (biovil_t/resnet.py:25-47 Bottleneck conv2 3x3 and the stride-2 1x1 downsample; eval BatchNorm
 * folded into w / bias): out[(b,oh,ow), n] = epilogue(sum_{kh,kw,c} x[b, oh*stride-pad+kh, ow*stride-pad+kw, c] . w[n, (kh*ks+kw)*C+c]).
 * x is NHWC [B,H,W,C], w is [Cout, ks*ks*C] - the column order of rd_im2col_nhwc, so the result is bit-identical to
 * rd_im2col_nhwc + rd_linear; the gather happens inside the GEMM's TMA producer (im2col-mode tensor map), no matrix in HBM.
 * Returns 1 = launched, 0 = shape not eligible (C % 64 != 0, M <= 128, unaligned out: use the explicit path), < 0 = error.
 * rd_conv_set_implicit(0) makes it always return 0 (test hook). */
int rd_conv_nhwc_implicit(const void* x_dev, const void* w_dev, void* out_dev, int64_t ldo, int B, int H, int W, int C, int Cout,
                          int ks, int stride, int pad, const rd_epilogue* epi, int dtype, void* stream);
int rd_conv_set_implicit(int on);
/* Launch every kernel with the programmatic-dependent-launch attribute (prologue of kernel N+1 — barrier init, TMEM
 * allocation, the first weight tiles — overlaps the tail of kernel N).  On by default; 0 switches it off. */
int rd_set_pdl(int on);

/* LlamaRMSNorm.forward (modeling_llama_imgemb.py:85-93): fp32 mean of squares, x*rsqrt in fp32, round, then
 * multiply by the weight in the storage dtype.  If lora_a_dev != NULL also writes lora_t[M, lora_rows] =
 * T(out[m,:] . lora_a[r,:]) (the lora_A Linear of peft, fed to rd_linear's epilogue).                           */
int rd_rmsnorm(const void* x_dev, const void* w_dev, void* out_dev, int M, int H, float eps,
               const void* lora_a_dev, int lora_rows, void* lora_t_dev, int dtype, void* stream);

/* nn.LayerNorm over the last dim (fp32 statistics): ln_vision blip2.py:199-205 (eps 1e-5), BERT post-LN
 * Qformer.py:285-289,371-375 (eps 1e-12).  gamma/beta fp32.                                                     */
int rd_layernorm(const void* x_dev, const float* gamma_dev, const float* beta_dev, void* out_dev, int M, int H,
                 float eps, int dtype, void* stream);

/* apply_rotary_pos_emb (modeling_llama_imgemb.py:135-142) on q,k of the qkv buffer [M, ldq] + KV-cache append (replaces
 * the torch.cat at :209-212).  Token m = b*q_len+i goes to cache slot ctx_len[0]+i.  q is rotated in place.
 * cos/sin: [max_pos, hd] storage dtype (tables of :97-109 cast as in :123-124).  Caches: [B, nh, cmax, hd].
 * peft LoRA (unmerged, finetune.py:167-173): if lora_b_dev != NULL the row also holds t = T(lora_A . x) in columns
 * [3H, 3H+2r) (q's r values, then v's; produced by the QKV GEMM whose weight carries the lora_A rows) and
 * q,v <- T( T(Wx) + T(lora_scale * T(lora_B . t)) ) before the rotation / append; lora_b_dev is [2H, r] (q rows, v rows). */
int rd_rope_kv_store(void* qkv_dev, int64_t ldq, const int32_t* pos_dev, const int32_t* ctx_len_dev, const void* cos_dev,
                     const void* sin_dev, void* kcache_dev, void* vcache_dev, int B, int q_len, int nh, int hd,
                     int cmax, const void* lora_b_dev, int lora_r, float lora_scale, int dtype, void* stream);

/* LlamaAttention core (modeling_llama_imgemb.py:216-234) for q_len new tokens per row against the cache:
 * fp16 scores, /sqrt(hd), + causal/padding mask, clamp to finfo.min, fp32 softmax rounded to the storage dtype,
 * .V.  keymask[B,cmax] is the HF attention_mask (1 = attend).  out[M, nh*hd].                                   */
int rd_attention(const void* qkv_dev, int64_t ldq, const void* kcache_dev, const void* vcache_dev,
                 const uint8_t* keymask_dev, const int32_t* ctx_len_dev, void* out_dev, int B, int q_len, int nh,
                 int hd, int cmax, int dtype, void* stream);

/* q_len >= 4 launches whose keys fit one UMMA tile (ctx + q_len <= 256) run both contractions on the tcgen05 tensor cores
 * (csrc/attention_prefill_tc.cu: S = Q.K^T and O = P.V as UMMA tiles, the rounding points above applied in the TMEM -> register
 * pass between them); longer contexts use the SIMT kernels.  Test hook: 0 forces the SIMT kernels, 1 (default) restores. */
int rd_attention_set_tensor_core(int on);

/* Single-token decode: rd_rope_kv_store + rd_attention (q_len == 1) fused in one launch; pos_dev[B] are the position
 * ids of the new tokens, which are appended at cache slot ctx_len[0].  The KV sweep uses bulk asynchronous copies
 * (cp.async.bulk + mbarrier) through a shared-memory ring.  ctx_lower_bound: host-known lower bound of ctx_len[0]
 * (0 if unknown) — cache rows below it are requested before the programmatic-launch wait.  LoRA arguments as above. */
int rd_attention_decode(const void* qkv_dev, int64_t ldq, const int32_t* pos_dev, const void* cos_dev, const void* sin_dev,
                        void* kcache_dev, void* vcache_dev, const uint8_t* keymask_dev, const int32_t* ctx_len_dev,
                        void* out_dev, int B, int nh, int hd, int cmax, int ctx_lower_bound, const void* lora_b_dev,
                        int lora_r, float lora_scale, int dtype, void* stream);

/* rd_attention_decode with q/k/v (+ the LoRA t columns) handed over as the QKV GEMM's fp32 split-K partials
 * qkv_part_dev[n_part][part_stride] (row b at b*ldq): summed in split order and rounded once, T(Wx), as they are read. */
int rd_attention_decode_partials(const float* qkv_part_dev, int n_part, long long part_stride, int64_t ldq, const int32_t* pos_dev,
                                 const void* cos_dev, const void* sin_dev, void* kcache_dev, void* vcache_dev,
                                 const uint8_t* keymask_dev, const int32_t* ctx_len_dev, void* out_dev, int B, int nh, int hd, int cmax,
                                 int ctx_lower_bound, const void* lora_b_dev, int lora_r, float lora_scale, int dtype, void* stream);

/* LlamaModel.forward splice (modeling_llama_imgemb.py:571-594, split_at_img :498-520): out[b,t,:] = img[b,t-p,:]
 * for t in [p,p+32) where p = first index of 32000 in row b (0 if none), else embed[ids[b,t]].  img may be NULL
 * (plain embedding).                                                                                            */
int rd_embed_splice(const int64_t* ids_dev, const void* embed_dev, const void* img_dev, void* out_dev, int B, int T,
                    int H, int vocab, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * LLM engine: LlamaForCausalLM.forward + transformers 4.28.1 greedy_search
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct rd_llm rd_llm;

typedef struct rd_llm_config {
  int vocab, hidden, inter, layers, heads, max_pos;
  float rms_eps;
  int dtype;
  int lora_r;          /* 0 = no adapter */
  float lora_scale;    /* alpha / r (finetune.py:167-168 -> 2.0) */
  int qformer_hidden;  /* 768 */
  int max_batch, max_ctx;
  int pad_id, eos_id, img_id;
} rd_llm_config;

/* weight slots for rd_llm_set_weight; layer = -1 for the global ones */
#define RD_W_EMBED 0       /* [vocab, H]                      model.embed_tokens.weight            */
#define RD_W_FINAL_NORM 1  /* [H]                             model.norm.weight                    */
#define RD_W_LM_HEAD 2     /* [vocab, H]                      lm_head.weight                       */
#define RD_W_IMG_PROJ_W 3  /* [H, 768]                        model.img_proj_layer.weight          */
#define RD_W_IMG_PROJ_B 4  /* [H] fp32                        model.img_proj_layer.bias            */
#define RD_W_ROPE_COS 5    /* [max_pos, hd]                                                        */
#define RD_W_ROPE_SIN 6
#define RD_W_QKV 10        /* [3H (+2r), H]  q_proj|k_proj|v_proj rows, then q lora_A and v lora_A rows when an adapter is loaded */
#define RD_W_O 11          /* [H, H]                                                               */
#define RD_W_GATE_UP 12    /* [2I, H]  gate_proj rows then up_proj rows                            */
#define RD_W_DOWN 13       /* [H, I]                                                               */
#define RD_W_LN1 14        /* [H] input_layernorm                                                  */
#define RD_W_LN2 15        /* [H] post_attention_layernorm                                         */
#define RD_W_LORA_A 16     /* unused by the engine (lora_A rides in RD_W_QKV); kept for ABI stability  */
#define RD_W_LORA_B 17     /* [2H, r]  q lora_B rows then v lora_B rows                            */

int rd_llm_create(const rd_llm_config* cfg, rd_llm** out);
void rd_llm_destroy(rd_llm* h);
int rd_llm_set_weight(rd_llm* h, int layer, int slot, const void* ptr_dev);
/* GEMM path: 0 auto, 1 GEMV, 2 tcgen05, 3 SIMT (validation) */
int rd_llm_set_algo(rd_llm* h, int algo);
/* Single-token steps with B <= 32 (default ON): the QKV GEMM (q_proj|k_proj|v_proj (+lora_A), modeling_llama_imgemb.py:178-181)
 * leaves its fp32 split-K partials in an L2-resident slab and the attention kernel sums them in split order and applies the
 * single rounding T(Wx) while reading q/k/v - the GEMM has no cross-CTA reduction tail.  0 = the GEMM reduces them itself
 * (thread-block cluster + distributed shared memory).  Same rounding contract and the same summation order either way.  */
int rd_llm_set_qkv_partials(rd_llm* h, int on);
/* Single-token steps with B <= 32 (default ON): o_proj and down_proj (modeling_llama_imgemb.py:245,159) leave their fp32 split-K
 * partials in an L2-resident slab; the norm launch that follows each of them sums the partials in split order, adds the residual
 * (x = T(x + T(Wx)), :302,308) and applies the next LlamaRMSNorm (:305; input_layernorm of the next layer :287; model.norm :658).
 * The GEMMs lose their cross-CTA reduction tail.  Same rounding points as the cluster reduction + rd_rmsnorm (only the fp32
 * order of the norm's sum of squares differs: a token row is split over a 4-CTA cluster).  0 switches it off.                */
int rd_llm_set_od_partials(rd_llm* h, int on);
/* Decode step: bytes of W_qkv / W_o / W_gate|up that the (latency-bound, HBM-idle) norm and attention kernels pull into
 * the 126 MB L2 with cp.async.bulk.prefetch ahead of the GEMM that streams them; 0,0,0 turns it off. */
int rd_llm_set_l2_prefetch(rd_llm* h, long long qkv_bytes, long long o_bytes, long long gate_up_bytes);

/* Start generation: clears the KV cache, infers attention_mask = (ids != pad) and position_ids = cumsum-1
 * (prepare_inputs_for_generation, modeling_llama_imgemb.py:795-836), runs the prefill forward over ids[B,T] with
 * the image rows spliced in (img_embeds_dev: [B,32,768] storage dtype, or NULL = no image), and selects the first
 * token (greedy).  If all_logits_dev != NULL the full [B,T,vocab] logits are written there (parity dumps; the
 * reference computes them at :768) — otherwise only the last position goes through lm_head.
 * suppress_eos != 0 masks EOS in the argmax (fixed-work throughput runs).                                       */
int rd_llm_prefill(rd_llm* h, const int64_t* ids_dev, const void* img_embeds_dev, int B, int T,
                   void* all_logits_dev, int suppress_eos, void* stream);
/* Multi-turn: append ids[B,T] to the existing context (KV prefix reuse; identical tokens to a full re-prefill of
 * the concatenated conversation, demo.py:282-297) and select the next token.                                    */
int rd_llm_extend(rd_llm* h, const int64_t* ids_dev, int B, int T, int suppress_eos, void* stream);
/* Roll the context back to its first new_ctx cached tokens before rd_llm_extend (npos_host[b] = attended tokens among
 * them, HOST pointer, consumed before the call returns).  Stream-ordered; does not synchronise. */
int rd_llm_truncate(rd_llm* h, int new_ctx, const int32_t* npos_host, void* stream);
/* A CUDA-graph replay of a captured rd_llm_decode_step advances the device-side counters only; this keeps the host
 * mirror used for bounds checks in step (n may be -1 to undo the bump of the capture call itself). */
int rd_llm_note_replayed_steps(rd_llm* h, int n);
/* One greedy decode step for all rows (graph-capturable: no host-dependent state). */
int rd_llm_decode_step(rd_llm* h, void* stream);
/* Device-side generation state: tokens chosen so far [B, n_generated] (row-major, ld = max_ctx), per-row
 * finished flags (1 = row has emitted EOS), last-step logits [B, vocab].                                                                */
int rd_llm_state(rd_llm* h, const int64_t** gen_tokens_dev, const int32_t** finished_dev,
                 const void** last_logits_dev, const void** hidden_dev, int* n_generated_host);
/* Beam search (generate(num_beams=k), test.py:467,629): LlamaForCausalLM._reorder_cache (modeling_llama_imgemb.py:838-843) on the
 * flat KV cache - cache row r of every layer becomes cache row beam_idx_dev[r] (int32[B]).  Call outside stream capture and
 * discard captured decode graphs afterwards (the cache buffers are swapped).  Pair it with rd_llm_force_tokens to feed the
 * tokens the beam scorer selected.                                                                                          */
int rd_llm_reorder_cache(rd_llm* h, const int32_t* beam_idx_dev, void* stream);
/* Teacher forcing for parity tests: the next rd_llm_decode_step consumes toks_dev[B] (int64) instead of the engine's own
 * greedy choice, which stays in the generation record.  Mirrors feeding the reference's `input_ids[:, -1:]`
 * (prepare_inputs_for_generation, modeling_llama_imgemb.py:799-801) with an externally chosen token.               */
int rd_llm_force_tokens(rd_llm* h, const int64_t* toks_dev, void* stream);
/* Device word: 0 while any row is unfinished, else the number of generated tokens at which the last row emitted EOS
 * (transformers 4.28.1 greedy_search: `unfinished_sequences.max() == 0`).  Copy it asynchronously to poll.          */
int rd_llm_done_flag(rd_llm* h, const uint32_t** flag_dev);
/* Per-kernel-class CUDA-event timing of the eager path (bench.py roofline): enable, run steps, read back.
 * classes: 0 rmsnorm 1 qkv 2 rope 3 attn 4 o 5 gate_up 6 down 7 lm_head 8 argmax 9 embed                         */
int rd_llm_profile(rd_llm* h, int enable);
int rd_llm_profile_read(rd_llm* h, float* ms_per_class, int* launches_per_class, int n_classes);
int64_t rd_llm_launch_count(rd_llm* h);

/* ------------------------------------------------------------------------------------------------------------
 * Vision engine: Blip2Qformer.forward_image (blip2_qformer.py:467-484) = BioViL-T ResNet-50 trunk
 * (biovil_t/resnet.py:25-47) + backbone_to_vit + projector (biovil_t/encoder.py:124-130, modules.py:43-47) +
 * ln_vision + Q-Former query branch (Qformer.py:804-965)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct rd_vision rd_vision;

typedef struct rd_vision_config {
  int image_size;
  int layers[4];
  int width;
  int backbone_to_vit, joint;
  int num_query, q_hidden, q_heads, q_layers, q_inter, cross_freq;
  float ln_vision_eps, q_ln_eps;
  int dtype;
  int max_batch;
  /* two-image (temporal) branch, VisionTransformerPooler (biovil_t/transformer.py:28-75): 0 blocks = not configured */
  int pooler_blocks, pooler_heads, pooler_hidden;
  float pooler_ln_eps;
} rd_vision_config;

int rd_vision_create(const rd_vision_config* cfg, rd_vision** out);
void rd_vision_destroy(rd_vision* h);
/* Weights by name (the reference state_dict key with BatchNorm already folded by the host packer, see
 * radialog_b200/vision.py): conv weights are [Cout, kh*kw*Cin] storage dtype (NHWC im2col order), biases and
 * LayerNorm parameters fp32.                                                                                    */
int rd_vision_set_weight(rd_vision* h, const char* name, const void* ptr_dev);
/* images: [B,3,S,S] fp32 NCHW in [0,1] (ReportDataset.py:96-106).  q_out: [B,32,768] fp32, image_embeds
 * [B,196,1408] fp32 (may be NULL).                                                                              */
int rd_vision_forward(rd_vision* h, const float* images_dev, int B, float* q_out_dev, float* image_embeds_dev,
                      void* stream);
/* Two-image (temporal) mode - MultiImageEncoder.forward with a previous image (biovil_t/encoder.py:117-123): trunk +
 * backbone_to_vit on both images, VisionTransformerPooler (biovil_t/transformer.py:28-118) over the 2 x 196 tokens, the current
 * image's pooled tokens as the second half of the projector input; then as rd_vision_forward.  Extra weights by name:
 * vp{i}.ln1/.ln2 (.g/.b), vp{i}.qkv.w [3C,C], vp{i}.proj (.w/.b), vp{i}.fc1, vp{i}.fc2, vp.post (.g/.b), vp.pos_type [2P,C]
 * (sine position + type embedding, storage dtype), proj1f (.w [J,2C] / .b).  Allocates on first use.                       */
int rd_vision_forward_temporal(rd_vision* h, const float* images_dev, const float* prev_images_dev, int B, float* q_out_dev,
                               float* image_embeds_dev, void* stream);
int64_t rd_vision_launch_count(rd_vision* h);

/* ------------------------------------------------------------------------------------------------------------
 * Image preprocessing on the device
 * ---------------------------------------------------------------------------------------------------------- */
/* Bit-exact replacement of the reference's per-image CPU pipeline: remap_to_uint8 (demo.py:173-203) -> PIL "L" image
 * (demo.py:218) -> Resize(resize) -> CenterCrop(crop) -> ToTensor -> ExpandChannels
 * (model/lavis/data/ReportDataset.py:80-106, create_chest_xray_transform_for_inference(512, 448)).
 * img_dev: [H,W] row-major grey image on the device, dtype 0 = uint8, 1 = uint16, 2 = float32; remap = 0 skips
 * remap_to_uint8 (uint8 input only: an already remapped PIL image).  out_dev: float32 [3, crop, crop].          */
typedef struct rd_preproc rd_preproc;
int rd_preproc_create(int max_h, int max_w, int resize, int crop, rd_preproc** out);
void rd_preproc_destroy(rd_preproc* p);
int rd_preproc_run(rd_preproc* p, const void* img_dev, int dtype, int H, int W, int remap, float* out_dev, void* stream);
int64_t rd_preproc_launch_count(rd_preproc* p);

#ifdef __cplusplus
}
#endif
#endif
