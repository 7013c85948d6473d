// Non-GEMM kernels of the image path: BioViL-T ResNet-50 trunk glue (im2col, max-pool), the projector->token
// reinterpretation + ln_vision, LayerNorm, and the Q-Former's small attention.  All convolutions themselves run
// as rd_linear (tcgen05) GEMMs on NHWC activations.
#include "common.cuh"

bool rd_pdl_enabled();

// ------------------------------------------------------------------------------------------------
// stem: 7x7 / stride 2 / pad 3 im2col straight from the fp32 NCHW input (biovil_t/resnet.py:34)
// col[(b,oh,ow), (kh*7+kw)*3 + c], zero-padded to KP columns
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void stem_im2col_kernel(const float* __restrict__ img, T* __restrict__ col, int B, int S, int OH, int KP) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunks = KP / 8;
  const int64_t total = (int64_t)B * OH * OH * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % chunks);
    const int64_t row = idx / chunks;
    const int ow = (int)(row % OH), oh = (int)((row / OH) % OH), b = (int)(row / ((int64_t)OH * OH));
    Vec8<T> o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kidx = ch * 8 + e;
      float v = 0.f;
      if (kidx < 147) {
        const int c = kidx % 3, kw = (kidx / 3) % 7, kh = kidx / 21;
        const int ih = oh * 2 - 3 + kh, iw = ow * 2 - 3 + kw;
        if (ih >= 0 && ih < S && iw >= 0 && iw < S) v = img[(((int64_t)b * 3 + c) * S + ih) * S + iw];
      }
      o.v[e] = Tr<T>::r(v);
    }
    *reinterpret_cast<uint4*>(col + row * KP + ch * 8) = *reinterpret_cast<uint4*>(&o);
  }
}

extern "C" int rd_stem_im2col(const float* img, void* col, int B, int S, int KP, int dtype, void* stream) {
  RD_REQUIRE(KP % 8 == 0 && KP >= 147 && S % 2 == 0, "rd_stem_im2col: bad shape");
  const int OH = S / 2;
  const int64_t total = (int64_t)B * OH * OH * (KP / 8);
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(stem_im2col_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, rd_pdl_enabled(), img, (T*)col, B, S, OH, KP));
    return RD_OK;
  });
}

// ------------------------------------------------------------------------------------------------
// generic NHWC im2col (k x k, stride, pad), 8 channels (16 B) per thread.  k=1,stride=2 is the row gather of
// the Bottleneck downsample conv.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void im2col_nhwc_kernel(const T* __restrict__ in, T* __restrict__ col, int B, int H, int W, int C, int k, int stride,
                                   int pad, int OH, int OW) {
  pdl_launch_dependents();
  pdl_wait();
  const int c8 = C / 8;
  const int64_t total = (int64_t)B * OH * OW * k * k * c8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(idx % c8);
    int64_t r = idx / c8;
    const int kw = (int)(r % k); r /= k;
    const int kh = (int)(r % k); r /= k;
    const int64_t row = r;
    const int ow = (int)(row % OW), oh = (int)((row / OW) % OH), b = (int)(row / ((int64_t)OH * OW));
    const int ih = oh * stride - pad + kh, iw = ow * stride - pad + kw;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = *reinterpret_cast<const uint4*>(in + (((int64_t)b * H + ih) * W + iw) * C + cc * 8);
    *reinterpret_cast<uint4*>(col + (row * k * k + kh * k + kw) * C + cc * 8) = v;
  }
}

extern "C" int rd_im2col_nhwc(const void* in, void* col, int B, int H, int W, int C, int k, int stride, int pad, int dtype,
                              void* stream) {
  RD_REQUIRE(C % 8 == 0, "rd_im2col_nhwc: C must be a multiple of 8 (got %d)", C);
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  const int64_t total = (int64_t)B * OH * OW * k * k * (C / 8);
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(im2col_nhwc_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, rd_pdl_enabled(), (const T*)in, (T*)col,
                            B, H, W, C, k, stride, pad, OH, OW));
    return RD_OK;
  });
}

// max-pool 3x3 / stride 2 / pad 1, NHWC (biovil_t/resnet.py:37)
template <class T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C, int OH, int OW) {
  pdl_launch_dependents();
  pdl_wait();
  const int c8 = C / 8;
  const int64_t total = (int64_t)B * OH * OW * c8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(idx % c8);
    const int64_t row = idx / c8;
    const int ow = (int)(row % OW), oh = (int)((row / OW) % OH), b = (int)(row / ((int64_t)OH * OW));
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int ih = oh * 2 - 1 + kh, iw = ow * 2 - 1 + kw;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
          Vec8<T> v = ld16(in + (((int64_t)b * H + ih) * W + iw) * C + cc * 8);
#pragma unroll
          for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], Tr<T>::f(v.v[e]));
        }
      }
    Vec8<T> o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = Tr<T>::r(m[e]);
    *reinterpret_cast<uint4*>(out + row * C + cc * 8) = *reinterpret_cast<uint4*>(&o);
  }
}

extern "C" int rd_maxpool3x3s2(const void* in, void* out, int B, int H, int W, int C, int dtype, void* stream) {
  RD_REQUIRE(C % 8 == 0, "rd_maxpool3x3s2: C must be a multiple of 8");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)B * OH * OW * (C / 8);
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(maxpool3x3s2_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, rd_pdl_enabled(), (const T*)in, (T*)out, B, H,
                            W, C, OH, OW));
    return RD_OK;
  });
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (fp32 statistics, two-pass variance like torch).  Optional gather implements the reference's
// projected_patch_embeddings.reshape(B,-1,1408) (blip2_qformer.py:469): a raw reinterpretation of the NCHW buffer
// [J, P] as [P, J] tokens — token t, element e is NCHW-flat index f = t*J + e, i.e. channel f / P of pixel f % P,
// read here from the NHWC GEMM output [B, P, J].
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
layernorm_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ out,
                 float* __restrict__ out_f32, int H, float eps, int gather_P) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float srow[];
  __shared__ float sred[8];
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (gather_P > 0) {
    const int b = m / gather_P, t = m % gather_P;
    const T* base = x + (int64_t)b * gather_P * H;
    for (int e = tid; e < H; e += 256) {
      const int64_t f = (int64_t)t * H + e;
      const int c = (int)(f / gather_P), pix = (int)(f % gather_P);
      srow[e] = Tr<T>::f(base[(int64_t)pix * H + c]);
    }
  } else {
    for (int e = tid; e < H; e += 256) srow[e] = Tr<T>::f(x[(int64_t)m * H + e]);
  }
  __syncthreads();
  float s = 0.f;
  for (int e = tid; e < H; e += 256) s += srow[e];
  s = warp_sum(s);
  if (lane == 0) sred[warp] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mean += sred[i];
  mean /= (float)H;
  __syncthreads();
  float v = 0.f;
  for (int e = tid; e < H; e += 256) { float d = srow[e] - mean; v = fmaf(d, d, v); }
  v = warp_sum(v);
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) var += sred[i];
  const float rstd = 1.0f / sqrtf(var / (float)H + eps);
  for (int e = tid; e < H; e += 256) {
    const float y = (srow[e] - mean) * rstd * gamma[e] + beta[e];
    out[(int64_t)m * H + e] = Tr<T>::r(y);
    if (out_f32) out_f32[(int64_t)m * H + e] = y;
  }
}

static int launch_ln(const void* x, const float* g, const float* b, void* out, float* out_f32, int M, int H, float eps, int gather_P,
                     int dtype, void* stream) {
  RD_REQUIRE(M > 0 && H > 0 && H * 4 <= 48 * 1024, "layernorm: bad shape M=%d H=%d", M, H);
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(layernorm_kernel<T>, dim3(M), dim3(256), (size_t)H * 4, (cudaStream_t)stream, rd_pdl_enabled(), (const T*)x, g, b,
                            (T*)out, out_f32, H, eps, gather_P));
    return RD_OK;
  });
}

extern "C" int rd_layernorm(const void* x, const float* g, const float* b, void* out, int M, int H, float eps, int dtype, void* stream) {
  return launch_ln(x, g, b, out, nullptr, M, H, eps, 0, dtype, stream);
}
extern "C" int rd_ln_vision_tokens(const void* proj_nhwc, const float* g, const float* b, void* out, float* out_f32, int B, int P, int J,
                                   float eps, int dtype, void* stream) {
  return launch_ln(proj_nhwc, g, b, out, out_f32, B * P, J, eps, P, dtype, stream);
}

// ------------------------------------------------------------------------------------------------
// Q-Former attention (Qformer.py:198-268): 32 queries x {32 | 196} keys per (image, head); all masks are zero on
// this path, scale 1/sqrt(hd) applied after QK^T (:244), fp32 softmax.  One CTA per (head, image); K/V of the head
// live in shared memory (row stride padded by one word against bank conflicts); one warp per query row.
// ------------------------------------------------------------------------------------------------
template <class T, int HD>
__global__ void __launch_bounds__(256)
small_attention_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, const T* __restrict__ v, int64_t ldkv,
                       T* __restrict__ out, int64_t ldo, int q_len, int kv_len) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int LDS = HD + 2;
  extern __shared__ uint8_t sm_raw[];
  T* sk = reinterpret_cast<T*>(sm_raw);
  T* sv = sk + (size_t)kv_len * LDS;
  float* sp = reinterpret_cast<float*>(sv + (size_t)kv_len * LDS);     // [8 warps][kv_len]
  float* sq = sp + 8 * kv_len;                                          // [8 warps][HD]
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* kb = k + (int64_t)b * kv_len * ldkv + h * HD;
  const T* vb = v + (int64_t)b * kv_len * ldkv + h * HD;
  if ((ldkv & 7) == 0 && ((reinterpret_cast<uintptr_t>(kb) | reinterpret_cast<uintptr_t>(vb)) & 15) == 0) {
    // 16-byte global loads, four K and four V requests in flight per thread before the first one is consumed (the rows are
    // padded by one word in shared memory, so they are stored as 4-byte words); the 4-byte loop below paid one global
    // round trip per element pair and was most of the kernel's time for the 196-key cross attention
    constexpr int C8 = HD / 8;
    const int n16 = kv_len * C8;
    for (int base = tid; base < n16; base += 256 * 4) {
      uint4 kk[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 256;
        if (i < n16) {
          const int j = i / C8, c = i % C8;
          kk[u] = *reinterpret_cast<const uint4*>(kb + (int64_t)j * ldkv + c * 8);
          vv[u] = *reinterpret_cast<const uint4*>(vb + (int64_t)j * ldkv + c * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 256;
        if (i < n16) {
          const int j = i / C8, c = i % C8;
          uint32_t* dk = reinterpret_cast<uint32_t*>(sk + j * LDS + c * 8);
          uint32_t* dv = reinterpret_cast<uint32_t*>(sv + j * LDS + c * 8);
          dk[0] = kk[u].x; dk[1] = kk[u].y; dk[2] = kk[u].z; dk[3] = kk[u].w;
          dv[0] = vv[u].x; dv[1] = vv[u].y; dv[2] = vv[u].z; dv[3] = vv[u].w;
        }
      }
    }
  } else {
    for (int i = tid; i < kv_len * (HD / 2); i += 256) {
      const int j = i / (HD / 2), d2 = i % (HD / 2);
      *reinterpret_cast<uint32_t*>(sk + j * LDS + d2 * 2) = *reinterpret_cast<const uint32_t*>(kb + (int64_t)j * ldkv + d2 * 2);
      *reinterpret_cast<uint32_t*>(sv + j * LDS + d2 * 2) = *reinterpret_cast<const uint32_t*>(vb + (int64_t)j * ldkv + d2 * 2);
    }
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)HD);
  float* myp = sp + warp * kv_len;
  float* myq = sq + warp * HD;
  for (int i = warp; i < q_len; i += 8) {
    const T* qr = q + ((int64_t)b * q_len + i) * ldq + h * HD;
    for (int d = lane; d < HD; d += 32) myq[d] = Tr<T>::f(qr[d]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < kv_len; j += 32) {
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) s = fmaf(myq[d], Tr<T>::f(sk[j * LDS + d]), s);
      s *= scale;
      myp[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < kv_len; j += 32) { float e = expf(myp[j] - mx); myp[j] = e; sum += e; }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < HD; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < kv_len; ++j) acc = fmaf(myp[j], Tr<T>::f(sv[j * LDS + d]), acc);
      out[((int64_t)b * q_len + i) * ldo + h * HD + d] = Tr<T>::r(acc * inv);
    }
    __syncwarp();
  }
}

extern "C" int rd_small_attention(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out, int64_t ldo, int B,
                                  int heads, int hd, int q_len, int kv_len, int dtype, void* stream) {
  RD_REQUIRE(hd == 64 || hd == 32, "rd_small_attention: head_dim must be 32 or 64 (got %d)", hd);
  const size_t smem = (size_t)2 * kv_len * (hd + 2) * 2 + (size_t)8 * kv_len * 4 + (size_t)8 * hd * 4;
  RD_REQUIRE(smem <= 200 * 1024, "rd_small_attention: kv_len %d too large", kv_len);
  RD_DISPATCH_DTYPE(dtype, T, {
    if (hd == 64) {
      RD_SMEM_ATTR_ONCE(200 * 1024, small_attention_kernel<T, 64>);
      RD_CHECK_CUDA(rd_launch(small_attention_kernel<T, 64>, dim3(heads, B), dim3(256), smem, (cudaStream_t)stream, rd_pdl_enabled(), (const T*)q,
                              ldq, (const T*)k, (const T*)v, ldkv, (T*)out, ldo, q_len, kv_len));
    } else {
      RD_SMEM_ATTR_ONCE(200 * 1024, small_attention_kernel<T, 32>);
      RD_CHECK_CUDA(rd_launch(small_attention_kernel<T, 32>, dim3(heads, B), dim3(256), smem, (cudaStream_t)stream, rd_pdl_enabled(), (const T*)q,
                              ldq, (const T*)k, (const T*)v, ldkv, (T*)out, ldo, q_len, kv_len));
    }
    return RD_OK;
  });
}

// rows [R, H] broadcast to [B, R, H] (the constant-folded query embedding LayerNorm, Qformer.py:106) and T -> fp32 export
template <class T>
__global__ void broadcast_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t n_src, int64_t n_total) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i % n_src];
}
template <class T>
__global__ void cast_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = Tr<T>::f(src[i]);
}
extern "C" int rd_broadcast_rows(const void* src, void* dst, int64_t n_src, int B, int dtype, void* stream) {
  const int64_t n = n_src * B;
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(broadcast_rows_kernel<T>, dim3((unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256)), dim3(256), 0,
                            (cudaStream_t)stream, rd_pdl_enabled(), (const T*)src, (T*)dst, n_src, n));
    return RD_OK;
  });
}
extern "C" int rd_cast_f32(const void* src, float* dst, int64_t n, int dtype, void* stream) {
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(cast_f32_kernel<T>, dim3((unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256)), dim3(256), 0,
                            (cudaStream_t)stream, rd_pdl_enabled(), (const T*)src, dst, n));
    return RD_OK;
  });
}

// out[i] = T(a[i] + emb[i % period]): the position + type embedding added to the normalised tokens of every pooler block
// (Block.with_pos_and_type_embed, biovil_t/transformer.py:209-215); 8 elements per thread
template <class T>
__global__ void add_rows_bcast_kernel(const T* __restrict__ a, const T* __restrict__ emb, T* __restrict__ out, int64_t n8, int64_t period8) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const Vec8<T> x = ld16(a + i * 8), e = ld16(emb + (i % period8) * 8);
    Vec8<T> o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = Tr<T>::r(Tr<T>::f(x.v[k]) + Tr<T>::f(e.v[k]));
    *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(&o);
  }
}
extern "C" int rd_add_rows_bcast(const void* a, const void* emb, void* out, int64_t n, int64_t period, int dtype, void* stream) {
  RD_REQUIRE(n % 8 == 0 && period % 8 == 0 && period > 0, "rd_add_rows_bcast: sizes must be multiples of 8");
  const int64_t n8 = n / 8;
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(add_rows_bcast_kernel<T>, dim3((unsigned)((n8 + 255) / 256 > 1184 ? 1184 : (n8 + 255) / 256)), dim3(256), 0,
                            (cudaStream_t)stream, rd_pdl_enabled(), (const T*)a, (const T*)emb, (T*)out, n8, period / 8));
    return RD_OK;
  });
}
