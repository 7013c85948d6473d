// Vision engine: Blip2Qformer.forward_image (blip2_qformer.py:467-484) as a native runtime.
//   BioViL-T ResNet-50 trunk (biovil_t/resnet.py:25-47, torchvision Bottleneck v1.5) on NHWC activations: every conv is
//   an rd_linear (tcgen05) GEMM with BatchNorm folded into weight+bias by the host packer and ReLU / residual in the
//   epilogue; 3x3 and strided convs go through an im2col gather.
//   backbone_to_vit (encoder.py:126) -> projector MLP (modules.py:43-47; the constant missing_previous_emb half of
//   its input, encoder.py:128-130, is folded into the first bias) -> NCHW-flat token reinterpretation + ln_vision
//   (blip2_qformer.py:469, blip2.py:199-205) -> Q-Former query branch (Qformer.py:804-965): the query-embedding
//   LayerNorm is input independent and constant-folded, the six cross-attention K/V projections run as one GEMM.
#include <map>
#include <string>
#include <vector>
#include "common.cuh"

extern "C" int rd_stem_im2col(const float*, void*, int, int, int, int, void*);
extern "C" int rd_im2col_nhwc(const void*, void*, int, int, int, int, int, int, int, int, void*);
extern "C" int rd_conv_nhwc_implicit(const void*, const void*, void*, int64_t, int, int, int, int, int, int, int, int, const rd_epilogue*, int, void*);
extern "C" int rd_maxpool3x3s2(const void*, void*, int, int, int, int, int, void*);
extern "C" int rd_ln_vision_tokens(const void*, const float*, const float*, void*, float*, int, int, int, float, int, void*);
extern "C" int rd_small_attention(const void*, int64_t, const void*, const void*, int64_t, void*, int64_t, int, int, int, int, int, int, void*);
extern "C" int rd_broadcast_rows(const void*, void*, int64_t, int, int, void*);
extern "C" int rd_cast_f32(const void*, float*, int64_t, int, void*);
extern "C" int rd_add_rows_bcast(const void*, const void*, void*, int64_t, int64_t, int, void*);

static constexpr int STEM_KP = 152;   // 7*7*3 = 147 padded to a multiple of 8

struct rd_vision {
  rd_vision_config c;
  std::map<std::string, const void*> w;
  char *actA = nullptr, *actB = nullptr, *t1 = nullptr, *t2 = nullptr, *idt = nullptr, *col = nullptr;
  char *emb = nullptr, *kv = nullptr, *hq = nullptr, *qkv = nullptr, *ctx = nullptr, *tmp = nullptr, *ffn = nullptr, *ws = nullptr;
  // two-image (temporal) branch, allocated on first use: fused [B*P, 2C], pooler token streams [B*2P, C] x3, qkv [B*2P, 3C]
  char *fused = nullptr, *px = nullptr, *py = nullptr, *pt = nullptr, *pqkv = nullptr;
  int64_t ws_bytes = 0;
  int64_t launches = 0;
};

static int valloc(char** p, int64_t bytes) {
  RD_CHECK_CUDA(cudaMalloc((void**)p, (size_t)(bytes > 0 ? bytes : 16)));
  RD_CHECK_CUDA(cudaMemset(*p, 0, (size_t)(bytes > 0 ? bytes : 16)));
  return RD_OK;
}

extern "C" int rd_vision_create(const rd_vision_config* cfg, rd_vision** out) {
  RD_REQUIRE(cfg && out, "rd_vision_create: null argument");
  RD_REQUIRE(cfg->image_size % 32 == 0 && cfg->width % 8 == 0, "rd_vision_create: image_size must be a multiple of 32 and width of 8");
  RD_REQUIRE(cfg->q_hidden % cfg->q_heads == 0, "rd_vision_create: q_hidden must divide by q_heads");
  int dev = 0;
  RD_CHECK_CUDA(cudaGetDevice(&dev));
  if (!rd_device_ok(dev)) return RD_ERR_UNSUPPORTED;
  rd_vision* h = new rd_vision();
  h->c = *cfg;
  const int64_t B = cfg->max_batch, S = cfg->image_size, W0 = cfg->width, e = 2;
  // walk the trunk to size the buffers
  int64_t act = B * (S / 2) * (S / 2) * W0;                // stem output
  int64_t colmax = B * (S / 2) * (S / 2) * STEM_KP;
  int64_t t1max = 0, t2max = 0, idtmax = 0;
  int64_t hw = S / 4, inpl = W0;
  act = std::max(act, B * hw * hw * W0);
  for (int li = 0; li < 4; ++li) {
    const int64_t planes = W0 << li;
    for (int b = 0; b < cfg->layers[li]; ++b) {
      const int stride = (b == 0 && li > 0) ? 2 : 1;
      const int64_t ohw = hw / stride;
      t1max = std::max(t1max, B * hw * hw * planes);
      colmax = std::max(colmax, B * ohw * ohw * 9 * planes);
      t2max = std::max(t2max, B * ohw * ohw * planes);
      if (b == 0) { idtmax = std::max(idtmax, B * ohw * ohw * planes * 4); if (stride == 2) colmax = std::max(colmax, B * ohw * ohw * inpl); }
      act = std::max(act, B * ohw * ohw * planes * 4);
      hw = ohw; inpl = planes * 4;
    }
  }
  const int64_t P = hw * hw, J = cfg->joint, Hq = cfg->q_hidden, Q = cfg->num_query;
  int ncross = 0;
  for (int i = 0; i < cfg->q_layers; ++i) ncross += (i % cfg->cross_freq == 0);
  act = std::max(act, B * P * J);
  int r = RD_OK;
  auto A = [&](char** p, int64_t elems) { if (r == RD_OK) r = valloc(p, elems * e); };
  A(&h->actA, act); A(&h->actB, act); A(&h->t1, t1max); A(&h->t2, t2max); A(&h->idt, idtmax); A(&h->col, colmax);
  A(&h->emb, B * P * J); A(&h->kv, B * P * ncross * 2 * Hq); A(&h->hq, B * Q * Hq); A(&h->qkv, B * Q * 3 * Hq); A(&h->ctx, B * Q * Hq);
  A(&h->tmp, B * Q * Hq); A(&h->ffn, B * Q * cfg->q_inter);
  h->ws_bytes = rd_linear_workspace_bytes(64, std::max<int64_t>(3 * Hq, cfg->q_inter), std::max<int64_t>(cfg->q_inter, J));
  if (r == RD_OK) r = valloc(&h->ws, h->ws_bytes);
  if (r != RD_OK) { rd_vision_destroy(h); return r; }
  *out = h;
  return RD_OK;
}

extern "C" void rd_vision_destroy(rd_vision* h) {
  if (!h) return;
  void* ptrs[] = {h->actA, h->actB, h->t1, h->t2, h->idt, h->col, h->emb, h->kv, h->hq, h->qkv, h->ctx, h->tmp, h->ffn, h->ws,
                  h->fused, h->px, h->py, h->pt, h->pqkv};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
}

extern "C" int rd_vision_set_weight(rd_vision* h, const char* name, const void* ptr) {
  RD_REQUIRE(h && name && ptr, "rd_vision_set_weight: null argument");
  RD_REQUIRE(((uintptr_t)ptr & 15) == 0, "rd_vision_set_weight: %s must be 16-byte aligned", name);
  h->w[name] = ptr;
  return RD_OK;
}

extern "C" int64_t rd_vision_launch_count(rd_vision* h) { return h ? h->launches : 0; }

namespace {
struct Ctx {
  rd_vision* h; cudaStream_t st; int dt; int err = RD_OK;
  const void* W(const std::string& n) {
    auto it = h->w.find(n);
    if (it == h->w.end()) { if (err == RD_OK) { rd_set_error("rd_vision: weight '%s' not set", n.c_str()); err = RD_ERR_INVALID; } return nullptr; }
    return it->second;
  }
  // out[M,N] = act(x . W^T + b (+ res))
  void gemm(const void* x, int64_t ldx, const std::string& wname, bool bias, void* out, int64_t ldo, int M, int N, int K, int act,
            const void* res = nullptr, int64_t ld_res = 0) {
    if (err != RD_OK) return;
    rd_epilogue e{};
    const void* w = W(wname + ".w");
    if (bias) e.bias_dev = (const float*)W(wname + ".b");
    if (err != RD_OK) return;
    e.act = act;
    if (res) { e.residual_dev = res; e.ld_res = ld_res; e.res_mode = 2; }
    h->launches++;
    err = rd_linear(x, ldx, w, K, out, ldo, M, N, K, &e, dt, 2, h->ws, h->ws_bytes, st);
  }
  // out[(b,oh,ow), Cout] = act(conv_ks x ks(x NHWC) + b (+ res)): implicit GEMM (im2col-mode TMA inside the GEMM's producer) when the
  // shape allows, else an explicit im2col matrix in h->col followed by the GEMM.  Same products and k order either way.
  void conv(const void* x, int B, int Hi, int Wi, int C, int ks, int stride, int pad, const std::string& wname, void* out, int64_t ldo, int Cout,
            int act, const void* res = nullptr, int64_t ld_res = 0) {
    if (err != RD_OK) return;
    rd_epilogue e{};
    const void* w = W(wname + ".w");
    e.bias_dev = (const float*)W(wname + ".b");
    if (err != RD_OK) return;
    e.act = act;
    if (res) { e.residual_dev = res; e.ld_res = ld_res; e.res_mode = 2; }
    const int r = rd_conv_nhwc_implicit(x, w, out, ldo, B, Hi, Wi, C, Cout, ks, stride, pad, &e, dt, st);
    if (r == 1) { h->launches++; return; }
    if (r < 0) { err = r; return; }
    const int OH = (Hi + 2 * pad - ks) / stride + 1, OW = (Wi + 2 * pad - ks) / stride + 1;
    chk(rd_im2col_nhwc(x, h->col, B, Hi, Wi, C, ks, stride, pad, dt, st));
    gemm(h->col, (int64_t)ks * ks * C, wname, true, out, ldo, B * OH * OW, Cout, ks * ks * C, act, res, ld_res);
  }
  void ln(const void* x, const std::string& name, void* out, int M, int H, float eps) {
    if (err != RD_OK) return;
    const float* g = (const float*)W(name + ".g"); const float* b = (const float*)W(name + ".b");
    if (err != RD_OK) return;
    h->launches++;
    err = rd_layernorm(x, g, b, out, M, H, eps, dt, st);
  }
  void chk(int r) { if (err == RD_OK) { err = r; h->launches++; } }
};
}  // namespace

// ResNet-50 trunk + backbone_to_vit of B images: patch tokens [B*P, C] (NHWC) written to patch_out with row pitch ld_patch.
// Returns through cur/nxt the two ping-pong activation buffers (cur = free to overwrite afterwards).
static void trunk_b2v(rd_vision* h, Ctx& X, const float* images, int B, void* patch_out, int64_t ld_patch, char** cur_out, char** nxt_out) {
  const rd_vision_config& c = h->c;
  const int S = c.image_size, W0 = c.width, dt = c.dtype;
  void* st = X.st;
  // ---- stem: conv1 7x7/2 + BN + ReLU, max-pool 3x3/2 --------------------------------------------------------------
  int hw = S / 2;
  X.chk(rd_stem_im2col(images, h->col, B, S, STEM_KP, dt, st));
  X.gemm(h->col, STEM_KP, "conv1", true, h->actA, W0, B * hw * hw, W0, STEM_KP, RD_ACT_RELU);
  X.chk(rd_maxpool3x3s2(h->actA, h->actB, B, hw, hw, W0, dt, st));
  hw /= 2;
  char* cur = h->actB; char* nxt = h->actA;
  int inpl = W0;
  // ---- layer1..4: Bottleneck = 1x1 -> 3x3(stride) -> 1x1, + identity/downsample, ReLU ---------------------------------
  for (int li = 0; li < 4; ++li) {
    const int planes = W0 << li;
    for (int b = 0; b < c.layers[li]; ++b) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const int stride = (b == 0 && li > 0) ? 2 : 1;
      const int ohw = hw / stride;
      const int Min = B * hw * hw, Mout = B * ohw * ohw;
      X.gemm(cur, inpl, p + ".conv1", true, h->t1, planes, Min, planes, inpl, RD_ACT_RELU);
      X.conv(h->t1, B, hw, hw, planes, 3, stride, 1, p + ".conv2", h->t2, planes, planes, RD_ACT_RELU);
      const void* idt = cur;
      if (b == 0) {
        if (stride == 2) X.conv(cur, B, hw, hw, inpl, 1, 2, 0, p + ".downsample", h->idt, planes * 4, planes * 4, RD_ACT_NONE);
        else X.gemm(cur, inpl, p + ".downsample", true, h->idt, planes * 4, Mout, planes * 4, inpl, RD_ACT_NONE);
        idt = h->idt;
      }
      X.gemm(h->t2, planes, p + ".conv3", true, nxt, planes * 4, Mout, planes * 4, planes, RD_ACT_RELU, idt, planes * 4);
      std::swap(cur, nxt);
      hw = ohw; inpl = planes * 4;
    }
  }
  X.gemm(cur, inpl, "b2v", false, patch_out, ld_patch, B * hw * hw, c.backbone_to_vit, inpl, RD_ACT_NONE);
  *cur_out = nxt; *nxt_out = cur;      // both are free now (the trunk output has been consumed by backbone_to_vit)
}

// projector's second conv -> NCHW-flat token reinterpretation + ln_vision -> Q-Former; `proj1_out` holds ReLU(BN(conv1)) [B*P, J]
static int tail_from_proj1(rd_vision* h, Ctx& X, char* proj1_out, char* other, int B, float* q_out, float* image_embeds) {
  const rd_vision_config& c = h->c;
  const int dt = c.dtype;
  void* st = X.st;
  const int g = c.image_size / 32, P = g * g, J = c.joint, MP = B * P;
  char* cur = other;
  X.gemm(proj1_out, J, "proj2", true, cur, J, MP, J, J, RD_ACT_NONE);
  if (X.err == RD_OK) {
    const float* gg = (const float*)X.W("ln_vision.g"); const float* bb = (const float*)X.W("ln_vision.b");
    if (X.err == RD_OK) X.chk(rd_ln_vision_tokens(cur, gg, bb, h->emb, image_embeds, B, P, J, c.ln_vision_eps, dt, st));
  }
  // ---- Q-Former ---------------------------------------------------------------------------------------------------------
  const int Hq = c.q_hidden, Q = c.num_query, MQ = B * Q, hd = Hq / c.q_heads;
  int ncross = 0;
  for (int i = 0; i < c.q_layers; ++i) ncross += (i % c.cross_freq == 0);
  const int ldkv = ncross * 2 * Hq;
  X.gemm(h->emb, J, "q.cross_kv", true, h->kv, ldkv, MP, ldkv, J, RD_ACT_NONE);
  if (X.err == RD_OK) { const void* h0 = X.W("q.h0"); if (X.err == RD_OK) X.chk(rd_broadcast_rows(h0, h->hq, (int64_t)Q * Hq, B, dt, st)); }
  int ci = 0;
  for (int i = 0; i < c.q_layers; ++i) {
    const std::string p = "q" + std::to_string(i);
    X.gemm(h->hq, Hq, p + ".self_qkv", true, h->qkv, 3 * Hq, MQ, 3 * Hq, Hq, RD_ACT_NONE);
    if (X.err == RD_OK) X.chk(rd_small_attention(h->qkv, 3 * Hq, h->qkv + (int64_t)Hq * 2, h->qkv + (int64_t)2 * Hq * 2, 3 * Hq, h->ctx, Hq, B,
                                                 c.q_heads, hd, Q, Q, dt, st));
    X.gemm(h->ctx, Hq, p + ".self_out", true, h->tmp, Hq, MQ, Hq, Hq, RD_ACT_NONE, h->hq, Hq);
    X.ln(h->tmp, p + ".self_ln", h->hq, MQ, Hq, c.q_ln_eps);
    if (i % c.cross_freq == 0) {
      X.gemm(h->hq, Hq, p + ".cross_q", true, h->qkv, Hq, MQ, Hq, Hq, RD_ACT_NONE);
      const char* kbase = h->kv + (int64_t)ci * 2 * Hq * 2;
      if (X.err == RD_OK) X.chk(rd_small_attention(h->qkv, Hq, kbase, kbase + (int64_t)Hq * 2, ldkv, h->ctx, Hq, B, c.q_heads, hd, Q, P, dt, st));
      X.gemm(h->ctx, Hq, p + ".cross_out", true, h->tmp, Hq, MQ, Hq, Hq, RD_ACT_NONE, h->hq, Hq);
      X.ln(h->tmp, p + ".cross_ln", h->hq, MQ, Hq, c.q_ln_eps);
      ++ci;
    }
    X.gemm(h->hq, Hq, p + ".ffn1", true, h->ffn, c.q_inter, MQ, c.q_inter, Hq, RD_ACT_GELU);
    X.gemm(h->ffn, c.q_inter, p + ".ffn2", true, h->tmp, Hq, MQ, Hq, c.q_inter, RD_ACT_NONE, h->hq, Hq);
    X.ln(h->tmp, p + ".ffn_ln", h->hq, MQ, Hq, c.q_ln_eps);
  }
  if (X.err == RD_OK) X.chk(rd_cast_f32(h->hq, q_out, (int64_t)MQ * Hq, dt, st));
  return X.err;
}

extern "C" int rd_vision_forward(rd_vision* h, const float* images, int B, float* q_out, float* image_embeds, void* stream) {
  RD_REQUIRE(h && images && q_out, "rd_vision_forward: null argument");
  RD_REQUIRE(B > 0 && B <= h->c.max_batch, "rd_vision_forward: B=%d out of range (max_batch %d)", B, h->c.max_batch);
  const rd_vision_config& c = h->c;
  Ctx X{h, (cudaStream_t)stream, c.dtype};
  const int g = c.image_size / 32, MP = B * g * g, J = c.joint;
  char *cur, *nxt;
  trunk_b2v(h, X, images, B, h->t1, c.backbone_to_vit, &cur, &nxt);
  // projector conv1: the constant missing_previous_emb half of its input (encoder.py:128-130) is folded into the bias
  X.gemm(h->t1, c.backbone_to_vit, "proj1", true, nxt, J, MP, J, c.backbone_to_vit, RD_ACT_RELU);
  return tail_from_proj1(h, X, nxt, cur, B, q_out, image_embeds);
}

// Two-image (temporal) mode: MultiImageEncoder.forward with a previous image (biovil_t/encoder.py:117-123) - both images go
// through the trunk + backbone_to_vit, VisionTransformerPooler (biovil_t/transformer.py:28-118) runs over the 2 x P tokens
// (pre-LN blocks; the sine position + type embedding is added to the normalised input of every block's attention), and the
// current image's pooled tokens replace missing_previous_emb as the second half of the projector's input.
extern "C" int rd_vision_forward_temporal(rd_vision* h, const float* images, const float* prev_images, int B, float* q_out,
                                          float* image_embeds, void* stream) {
  RD_REQUIRE(h && images && prev_images && q_out, "rd_vision_forward_temporal: null argument");
  RD_REQUIRE(B > 0 && B <= h->c.max_batch, "rd_vision_forward_temporal: B=%d out of range (max_batch %d)", B, h->c.max_batch);
  const rd_vision_config& c = h->c;
  RD_REQUIRE(c.pooler_blocks > 0 && c.pooler_heads > 0 && c.backbone_to_vit % c.pooler_heads == 0, "rd_vision_forward_temporal: pooler not configured");
  const int C = c.backbone_to_vit, hdp = C / c.pooler_heads;
  RD_REQUIRE(hdp == 32 || hdp == 64, "rd_vision_forward_temporal: pooler head_dim must be 32 or 64 (got %d)", hdp);
  cudaStream_t st = (cudaStream_t)stream;
  const int g = c.image_size / 32, P = g * g, J = c.joint, MP = B * P, M2 = B * 2 * P, e = 2;
  if (h->fused == nullptr) {       // first use: allocate (call outside stream capture)
    const int64_t Bm = c.max_batch;
    RD_CHECK(valloc(&h->fused, Bm * P * 2 * C * e)); RD_CHECK(valloc(&h->px, Bm * 2 * P * C * e)); RD_CHECK(valloc(&h->py, Bm * 2 * P * C * e));
    RD_CHECK(valloc(&h->pt, Bm * 2 * P * std::max(C, c.pooler_hidden) * e)); RD_CHECK(valloc(&h->pqkv, Bm * 2 * P * 3 * C * e));
  }
  Ctx X{h, st, c.dtype};
  char *cur, *nxt;
  // current image -> first half of the fused projector input AND first P tokens of every image's pooler stream
  trunk_b2v(h, X, images, B, h->fused, 2 * C, &cur, &nxt);
  if (X.err != RD_OK) return X.err;
  for (int b = 0; b < B; ++b)      // fused[b, p, 0:C] -> px[b, p, :]
    RD_CHECK_CUDA(cudaMemcpy2DAsync(h->px + (int64_t)b * 2 * P * C * e, (size_t)C * e, h->fused + (int64_t)b * P * 2 * C * e, (size_t)2 * C * e,
                                    (size_t)C * e, P, cudaMemcpyDeviceToDevice, st));
  // previous image -> tokens P..2P-1 (through a dense scratch, then one strided copy)
  trunk_b2v(h, X, prev_images, B, h->py, C, &cur, &nxt);
  if (X.err != RD_OK) return X.err;
  RD_CHECK_CUDA(cudaMemcpy2DAsync(h->px + (int64_t)P * C * e, (size_t)2 * P * C * e, h->py, (size_t)P * C * e, (size_t)P * C * e, B,
                                  cudaMemcpyDeviceToDevice, st));
  // ---- VisionTransformerPooler blocks -------------------------------------------------------------------------------------
  char* x = h->px; char* y = h->py;
  const void* emb = X.W("vp.pos_type");
  if (X.err != RD_OK) return X.err;
  for (int i = 0; i < c.pooler_blocks; ++i) {
    const std::string p = "vp" + std::to_string(i);
    X.ln(x, p + ".ln1", h->pt, M2, C, c.pooler_ln_eps);
    X.chk(rd_add_rows_bcast(h->pt, emb, h->pt, (int64_t)M2 * C, (int64_t)2 * P * C, c.dtype, st));
    X.gemm(h->pt, C, p + ".qkv", false, h->pqkv, 3 * C, M2, 3 * C, C, RD_ACT_NONE);
    if (X.err == RD_OK) X.chk(rd_small_attention(h->pqkv, 3 * C, h->pqkv + (int64_t)C * e, h->pqkv + (int64_t)2 * C * e, 3 * C, h->pt, C, B,
                                                 c.pooler_heads, hdp, 2 * P, 2 * P, c.dtype, st));
    X.gemm(h->pt, C, p + ".proj", true, y, C, M2, C, C, RD_ACT_NONE, x, C);                  // x + proj(attn)
    X.ln(y, p + ".ln2", h->pt, M2, C, c.pooler_ln_eps);
    X.gemm(h->pt, C, p + ".fc1", true, h->pqkv, c.pooler_hidden, M2, c.pooler_hidden, C, RD_ACT_GELU);
    X.gemm(h->pqkv, c.pooler_hidden, p + ".fc2", true, x, C, M2, C, c.pooler_hidden, RD_ACT_NONE, y, C);   // y + mlp(norm2(y))
  }
  X.ln(x, "vp.post", y, M2, C, c.pooler_ln_eps);
  if (X.err != RD_OK) return X.err;
  for (int b = 0; b < B; ++b)      // current image's tokens -> second half of the fused projector input
    RD_CHECK_CUDA(cudaMemcpy2DAsync(h->fused + (int64_t)b * P * 2 * C * e + (int64_t)C * e, (size_t)2 * C * e, y + (int64_t)b * 2 * P * C * e,
                                    (size_t)C * e, (size_t)C * e, P, cudaMemcpyDeviceToDevice, st));
  X.gemm(h->fused, 2 * C, "proj1f", true, nxt, J, MP, J, 2 * C, RD_ACT_RELU);
  return tail_from_proj1(h, X, nxt, cur, B, q_out, image_embeds);
}
