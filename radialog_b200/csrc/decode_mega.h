// Host interface of the persistent decode-layer kernel (decode_mega.cu), used by engine_llm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct rd_mega;

struct MegaLayerDesc {
  const void *qkv, *o, *gate_up, *down;   // [3H+2r, H], [H, H], [2I, H], [H, I]   (row-major, K contiguous)
  const void *ln1, *ln2;                  // [H]
  const void* lora_b;                     // [2H, r] or nullptr
  void *kc, *vc;                          // this layer's slab of the flat KV cache [B_max, nh, cmax, 128]
};

struct MegaCreate {
  int H, I, nh, layers, lora_r, dtype, max_batch, cmax;
  float lora_scale, eps;
};

struct MegaStep {
  void *x, *qkv, *att, *mid;              // activations of the engine: [B,H], [B,3H+2r], [B,H], [B,I]
  const uint8_t* keymask;                 // [B_max, cmax]
  const int32_t* ctx_len;                 // [1] cached tokens (the new token goes to slot ctx_len[0])
  const int32_t* pos;                     // [B] RoPE position of the new token
  const void *cos, *sin;                  // [max_pos, 128]
  int B, layer_begin, layer_end;
};

// nullptr-safe: returns a reason string (static storage) when the shape cannot run on the persistent kernel, else nullptr
const char* rd_mega_unsupported_reason(const MegaCreate* c);
int rd_mega_create(const MegaCreate* c, const MegaLayerDesc* layers, rd_mega** out);
void rd_mega_destroy(rd_mega* m);
int rd_mega_launch(rd_mega* m, const MegaStep* s, cudaStream_t st);
