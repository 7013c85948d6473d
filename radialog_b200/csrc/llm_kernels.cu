// Non-GEMM kernels of the Llama decoder path.  Every rounding point follows the reference
// (model/lavis/models/blip2_models/modeling_llama_imgemb.py).
#include "common.cuh"

bool rd_pdl_enabled();

// ------------------------------------------------------------------------------------------------
// RMSNorm (+ LoRA-A side product)                                      modeling_llama_imgemb.py:85-93
// one CTA per token row; the normalised row stays in shared memory for the lora_A dot products
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ out, int H, float eps,
               const T* __restrict__ lora_a, int lora_rows, T* __restrict__ lora_t, const void* pf_ptr, long long pf_bytes) {
  pdl_launch_dependents();
  l2_prefetch_slice(pf_ptr, pf_bytes, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x);     // weights: no dependency on x
  extern __shared__ float srow[];      // H floats
  __shared__ float sred[8];
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* xr = x + (int64_t)m * H;
  if (H <= 8 * 256 * 4 && lora_a == nullptr) {
    // rows of up to 8192 elements (Vicuna: 4096) stay in registers, and the norm weights - which do not depend on the previous
    // kernel - are requested BEFORE the programmatic-launch wait: one global round trip after the wait instead of two
    // (same per-thread element order and reduction tree as the general path below: bit-identical results)
    Vec8<T> wv[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int k = tid * 8 + u * 2048; if (k < H) wv[u] = ld16(w + k); }
    pdl_wait();
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int k = tid * 8 + u * 2048; if (k < H) xv[u] = ld16(xr + k); }
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (tid * 8 + u * 2048 < H) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float f = Tr<T>::f(xv[u].v[e]); ss = fmaf(f, f, ss); }
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) sred[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += sred[i];
    const float rs = 1.0f / sqrtf(tot / (float)H + eps);       // torch.rsqrt(variance + eps), fp32
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = tid * 8 + u * 2048;
      if (k < H) {
        Vec8<T> o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float y = Tr<T>::rr(Tr<T>::f(xv[u].v[e]) * rs);          // .to(weight.dtype)
          o.v[e] = Tr<T>::r(Tr<T>::f(wv[u].v[e]) * y);                   // weight * hidden_states
        }
        *reinterpret_cast<uint4*>(out + (int64_t)m * H + k) = *reinterpret_cast<uint4*>(&o);
      }
    }
    return;
  }
  pdl_wait();
  float ss = 0.f;
  for (int k = tid * 8; k < H; k += 256 * 8) {
    Vec8<T> v = ld16(xr + k);
#pragma unroll
    for (int e = 0; e < 8; ++e) { float f = Tr<T>::f(v.v[e]); srow[k + e] = f; ss = fmaf(f, f, ss); }
  }
  ss = warp_sum(ss);
  if (lane == 0) sred[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += sred[i];
  const float rs = 1.0f / sqrtf(tot / (float)H + eps);       // torch.rsqrt(variance + eps), fp32
  for (int k = tid * 8; k < H; k += 256 * 8) {
    Vec8<T> wv = ld16(w + k), o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = Tr<T>::rr(srow[k + e] * rs);                   // .to(weight.dtype)
      float r = Tr<T>::rr(Tr<T>::f(wv.v[e]) * y);              // weight * hidden_states
      o.v[e] = Tr<T>::r(r);
      srow[k + e] = r;
    }
    *reinterpret_cast<uint4*>(out + (int64_t)m * H + k) = *reinterpret_cast<uint4*>(&o);
  }
  if (lora_a == nullptr) return;
  __syncthreads();
  // lora_A: Linear(H -> rows), fp32 accumulate.  The A rows are L2 resident; all chunks of a row are requested
  // before the first is consumed (the loop used to run at one L2 round trip per 256 elements).
  for (int r = warp; r < lora_rows; r += 8) {
    const T* ar = lora_a + (int64_t)r * H;
    float acc = 0.f;
    for (int k0 = lane * 8; k0 < H; k0 += 32 * 8 * 8) {
      Vec8<T> av[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * 256;
        if (k < H) av[u] = ld16(ar + k);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * 256;
        if (k < H) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = fmaf(Tr<T>::f(av[u].v[e]), srow[k + e], acc);
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) lora_t[(int64_t)m * lora_rows + r] = Tr<T>::r(acc);
  }
}

// Split-K partials -> residual stream -> RMSNorm in one launch (decode, o_proj / down_proj).  The GEMM left its fp32 split-K
// partials in part[splits][slab_rows][H]; this kernel finishes the projection the way the GEMM epilogue would -
// x[m,:] = T(x[m,:] + T(sum over splits, split order)) - and applies the LlamaRMSNorm that follows in LlamaDecoderLayer.forward
// (modeling_llama_imgemb.py:85-93,302-305; model.norm after the last layer) to the new row: xn = T(w * T(x * rstd)).
// A token row is split over a thread-block cluster of 4 CTAs (a row's partials are 8 x 16 KB of L2 reads: one SM's ~100 GB/s
// of L2 ingest would make it the slowest kernel of the layer); the four quarter sums of squares meet over distributed shared
// memory in rank order.  Same rounding points as rd_rmsnorm; only the fp32 order of the sum of squares differs.
constexpr int RNP_CL = 4, RNP_THREADS = 128, RNP_MAXG = 4;

template <class T>
__global__ void __launch_bounds__(RNP_THREADS)
rmsnorm_partials_kernel(const float* __restrict__ part, int splits, int64_t slab_stride, T* __restrict__ x, const T* __restrict__ w,
                        T* __restrict__ out, int H, float eps) {
  pdl_launch_dependents();
  __shared__ float swarp[RNP_THREADS / 32];
  __shared__ float s_parts[RNP_CL];     // the four quarter sums of this row, each pushed here by the CTA that owns it
  __shared__ __align__(8) uint64_t xbar;  // completes when all four quarter sums (16 bytes) have landed
  const int c = blockIdx.x, m = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Hq = H / RNP_CL, groups = Hq / 8;
  const int col0 = c * Hq;
  Vec8<T> wv[RNP_MAXG], xv[RNP_MAXG];
#pragma unroll
  for (int u = 0; u < RNP_MAXG; ++u) { const int g = tid + u * RNP_THREADS; if (g < groups) wv[u] = ld16(w + col0 + g * 8); }      // no dependency on the GEMM
  // the exchange barrier is armed, and the whole cluster knows it, BEFORE the PDL wait: the only cluster-wide barrier of the kernel
  // costs nothing on the dependent path, which is left with four asynchronous remote stores and a local mbarrier wait
  if (tid == 0) {
    const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&xbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(4 * RNP_CL) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  pdl_wait();
  T* xr = x + (int64_t)m * H + col0;
  const float* pr = part + (int64_t)m * H + col0;
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < RNP_MAXG; ++u) {
    const int g = tid + u * RNP_THREADS;
    if (g < groups) {
      xv[u] = ld16(xr + g * 8);
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int s0 = 0; s0 < splits; s0 += 8) {            // 16 independent 16-byte L2 loads in flight per thread
        float4 v[8][2];
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s0 + s < splits) {
            const float4* p4 = reinterpret_cast<const float4*>(pr + (int64_t)(s0 + s) * slab_stride + g * 8);
            v[s][0] = __ldcg(p4); v[s][1] = __ldcg(p4 + 1);
          }
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s0 + s < splits) {
            acc[0] += v[s][0].x; acc[1] += v[s][0].y; acc[2] += v[s][0].z; acc[3] += v[s][0].w;
            acc[4] += v[s][1].x; acc[5] += v[s][1].y; acc[6] += v[s][1].z; acc[7] += v[s][1].w;
          }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xv[u].v[e] = Tr<T>::r(Tr<T>::f(xv[u].v[e]) + Tr<T>::rr(acc[e]));       // residual + T(Wx), rounded: the new residual stream
        const float f = Tr<T>::f(xv[u].v[e]);
        ss = fmaf(f, f, ss);
      }
      *reinterpret_cast<uint4*>(xr + g * 8) = *reinterpret_cast<const uint4*>(&xv[u]);
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) swarp[warp] = ss;
  __syncthreads();
  if (tid < RNP_CL) {
    // every CTA PUSHES its quarter sum into all four CTAs' shared memory (st.async: the store completes the receiver's mbarrier);
    // everybody then waits on its own barrier and reads locally - no cluster barrier on the dependent path, none before exit
    const float mine = ((swarp[0] + swarp[1]) + swarp[2]) + swarp[3];
    uint32_t la = (uint32_t)__cvta_generic_to_shared(&s_parts[c]), lb = (uint32_t)__cvta_generic_to_shared(&xbar), ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"((uint32_t)tid));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"((uint32_t)tid));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(mine)), "r"(rb) : "memory");
  }
  {
    const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&xbar);
    uint32_t done, spins = 0;
    long long t0 = 0;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(ba) : "memory");
      if (!done && (++spins & 0x3FFu) == 0) {          // bounded: a protocol bug must trap, not hang the GPU
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ll) { printf("rmsnorm_partials: exchange barrier timed out (block %d,%d)\n", blockIdx.x, blockIdx.y); __trap(); }
      }
    } while (!done);
  }
  const float tot = ((s_parts[0] + s_parts[1]) + s_parts[2]) + s_parts[3];             // rank order: the same total in all four CTAs
  const float rs = 1.0f / sqrtf(tot / (float)H + eps);       // torch.rsqrt(variance + eps), fp32
#pragma unroll
  for (int u = 0; u < RNP_MAXG; ++u) {
    const int g = tid + u * RNP_THREADS;
    if (g < groups) {
      Vec8<T> o;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = Tr<T>::rr(Tr<T>::f(xv[u].v[e]) * rs);          // .to(weight.dtype)
        o.v[e] = Tr<T>::r(Tr<T>::f(wv[u].v[e]) * y);                   // weight * hidden_states
      }
      *reinterpret_cast<uint4*>(out + (int64_t)m * H + col0 + g * 8) = *reinterpret_cast<uint4*>(&o);
    }
  }
}

int rd_rmsnorm_partials(const float* part, int splits, int64_t slab_stride, void* x, const void* w, void* out, int M, int H, float eps,
                        int dtype, void* stream) {
  RD_REQUIRE(M > 0 && H > 0 && H % (8 * RNP_CL) == 0 && H <= 8 * RNP_CL * RNP_THREADS * RNP_MAXG && splits >= 1,
             "rd_rmsnorm_partials: bad shape M=%d H=%d splits=%d", M, H, splits);
  RD_DISPATCH_DTYPE(dtype, T, {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(RNP_CL, M); cfg.blockDim = dim3(RNP_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = RNP_CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (rd_pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    RD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rmsnorm_partials_kernel<T>, part, splits, slab_stride, (T*)x, (const T*)w, (T*)out, H, eps));
    return RD_OK;
  });
}

static int rmsnorm_impl(const void* x, const void* w, void* out, int M, int H, float eps, const void* lora_a,
                        int lora_rows, void* lora_t, const void* pf_ptr, long long pf_bytes, int dtype, void* stream) {
  RD_REQUIRE(M > 0 && H > 0 && H % 8 == 0, "rd_rmsnorm: bad shape M=%d H=%d", M, H);
  RD_REQUIRE(H * 4 <= 200 * 1024, "rd_rmsnorm: H=%d too large", H);
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_SMEM_ATTR_ONCE(200 * 1024, rmsnorm_kernel<T>);
    RD_CHECK_CUDA(rd_launch(rmsnorm_kernel<T>, dim3(M), dim3(256), (size_t)H * 4, (cudaStream_t)stream, rd_pdl_enabled(),
                            (const T*)x, (const T*)w, (T*)out, H, eps, (const T*)lora_a, lora_rows, (T*)lora_t, pf_ptr, pf_bytes));
    return RD_OK;
  });
}

extern "C" int rd_rmsnorm(const void* x, const void* w, void* out, int M, int H, float eps, const void* lora_a,
                          int lora_rows, void* lora_t, int dtype, void* stream) {
  return rmsnorm_impl(x, w, out, M, H, eps, lora_a, lora_rows, lora_t, nullptr, 0, dtype, stream);
}
// rd_rmsnorm that also pulls [pf_ptr, pf_ptr+pf_bytes) (the next GEMM's weights) into L2 while it runs
extern "C" int rd_rmsnorm_prefetch(const void* x, const void* w, void* out, int M, int H, float eps, const void* pf_ptr,
                                   long long pf_bytes, int dtype, void* stream) {
  return rmsnorm_impl(x, w, out, M, H, eps, nullptr, 0, nullptr, pf_ptr, pf_bytes, dtype, stream);
}

// ------------------------------------------------------------------------------------------------
// RoPE + KV-cache append                                 modeling_llama_imgemb.py:128-142, 209-212
// q_embed = (q*cos) + (rotate_half(q)*sin): three separately rounded ops; cache keeps post-RoPE K.
// ------------------------------------------------------------------------------------------------
// peft LoRA on q_proj / v_proj, applied here instead of in the QKV GEMM epilogue: the GEMM also produced
// t = T(lora_A . xn) as 2r extra output columns (q's r values, then v's), and  y = T( T(Wx) + T(scale * T(lora_B . t)) ).
template <class T>
__device__ __forceinline__ float lora_add(float y, const T* __restrict__ brow, const float* t, int r, float scale) {
  float s = 0.f;
  if (r == 8) {                                   // the adapter rank of the reference (finetune.py:167): one 128-bit row
    const Vec8<T> bv = ld16(brow);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(Tr<T>::f(bv.v[i]), t[i], s);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < r) s = fmaf(Tr<T>::f(brow[i]), t[i], s);
  }
  return Tr<T>::rr(y + Tr<T>::rr(scale * Tr<T>::rr(s)));
}

template <class T>
__global__ void __launch_bounds__(256)
rope_kv_store_kernel(T* __restrict__ qkv, int64_t ldq, const int32_t* __restrict__ pos, const int32_t* __restrict__ ctx_len,
                     const T* __restrict__ cos_t, const T* __restrict__ sin_t, T* __restrict__ kc, T* __restrict__ vc,
                     int q_len, int nh, int hd, int cmax, const T* __restrict__ lora_b, int lora_r, float lora_scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, b = m / q_len, i = m % q_len;
  const int H = nh * hd, half = hd / 2;
  const int slot = ctx_len[0] + i;
  const int p = pos[m];
  T* row = qkv + (int64_t)m * ldq;
  float tq[16], tv[16];                    // lora_A outputs of this token (only read when lora_r > 0)
#pragma unroll
  for (int i = 0; i < 16; ++i) { tq[i] = 0.f; tv[i] = 0.f; }
  if (lora_r == 8) {
    const Vec8<T> a = ld16(row + 3 * H), bb = ld16(row + 3 * H + 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) { tq[i] = Tr<T>::f(a.v[i]); tv[i] = Tr<T>::f(bb.v[i]); }
  } else if (lora_r > 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < lora_r) { tq[i] = Tr<T>::f(row[3 * H + i]); tv[i] = Tr<T>::f(row[3 * H + lora_r + i]); }
  }
  const T* cr = cos_t + (int64_t)p * hd;
  const T* sr = sin_t + (int64_t)p * hd;
  for (int idx = threadIdx.x; idx < nh * half; idx += blockDim.x) {
    const int h = idx / half, d = idx % half;
    const float c_lo = Tr<T>::f(cr[d]), c_hi = Tr<T>::f(cr[d + half]);
    const float s_lo = Tr<T>::f(sr[d]), s_hi = Tr<T>::f(sr[d + half]);
    const int64_t cache_off = (((int64_t)b * nh + h) * cmax + slot) * hd;
    const int n_lo = h * hd + d, n_hi = n_lo + half;
    {
      T* q = row + h * hd;
      float lo = Tr<T>::f(q[d]), hi = Tr<T>::f(q[d + half]);
      if (lora_r > 0) {
        lo = lora_add<T>(lo, lora_b + (int64_t)n_lo * lora_r, tq, lora_r, lora_scale);
        hi = lora_add<T>(hi, lora_b + (int64_t)n_hi * lora_r, tq, lora_r, lora_scale);
      }
      q[d] = Tr<T>::r(Tr<T>::rr(lo * c_lo) + Tr<T>::rr(-hi * s_lo));
      q[d + half] = Tr<T>::r(Tr<T>::rr(hi * c_hi) + Tr<T>::rr(lo * s_hi));
    }
    {
      const T* k = row + H + h * hd;
      float lo = Tr<T>::f(k[d]), hi = Tr<T>::f(k[d + half]);
      kc[cache_off + d] = Tr<T>::r(Tr<T>::rr(lo * c_lo) + Tr<T>::rr(-hi * s_lo));
      kc[cache_off + d + half] = Tr<T>::r(Tr<T>::rr(hi * c_hi) + Tr<T>::rr(lo * s_hi));
    }
    {
      const T* v = row + 2 * H + h * hd;
      float lo = Tr<T>::f(v[d]), hi = Tr<T>::f(v[d + half]);
      if (lora_r > 0) {
        lo = lora_add<T>(lo, lora_b + (int64_t)(H + n_lo) * lora_r, tv, lora_r, lora_scale);
        hi = lora_add<T>(hi, lora_b + (int64_t)(H + n_hi) * lora_r, tv, lora_r, lora_scale);
      }
      vc[cache_off + d] = Tr<T>::r(lo);
      vc[cache_off + d + half] = Tr<T>::r(hi);
    }
  }
}

// Vectorised variant for head_dim 128 and LoRA rank 0 / 8: one thread per (head, 8-dim chunk of the low half) holds the
// eight (d, d + 64) pairs of q, k and v as 128-bit values; every load of a thread is independent of the others, so the
// kernel costs one L2 round trip instead of a chain of 2-byte loads per element (91 us -> per launch at B x T = 2048).
// Same formulas and rounding points as rope_kv_store_kernel.
template <class T, int R>
__global__ void __launch_bounds__(256)
rope_kv_store_vec_kernel(T* __restrict__ qkv, int64_t ldq, const int32_t* __restrict__ pos, const int32_t* __restrict__ ctx_len,
                         const T* __restrict__ cos_t, const T* __restrict__ sin_t, T* __restrict__ kc, T* __restrict__ vc,
                         int q_len, int nh, int cmax, const T* __restrict__ lora_b, float lora_scale, int M, int tokens_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128, HALF = 64;
  const int H = nh * HD;
  const int ctx0 = ctx_len[0];
  const int m_begin = blockIdx.x * tokens_per_cta, m_end = min(M, m_begin + tokens_per_cta);
  for (int it = threadIdx.x; it < nh * 8; it += blockDim.x) {
    const int h = it >> 3, d0 = (it & 7) * 8;
    // the 32 lora_B rows of this thread's (head, dim chunk) are loaded ONCE and reused for every token of the CTA's group
    // (one CTA per token re-read all 128 KB of lora_B from L2 per token: 268 MB per layer at B x T = 2048)
    Vec8<T> bq_lo[R == 8 ? 8 : 1], bq_hi[R == 8 ? 8 : 1], bv_lo[R == 8 ? 8 : 1], bv_hi[R == 8 ? 8 : 1];
    if (R == 8) {
      const T* bq = lora_b + (int64_t)(h * HD + d0) * 8;
      const T* bv = lora_b + (int64_t)(H + h * HD + d0) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        bq_lo[e] = ld16(bq + e * 8); bq_hi[e] = ld16(bq + (HALF + e) * 8);
        bv_lo[e] = ld16(bv + e * 8); bv_hi[e] = ld16(bv + (HALF + e) * 8);
      }
    }
    auto lora = [&](float y, const Vec8<T>& brow, const float* t) {
      float sdot = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) sdot = fmaf(Tr<T>::f(brow.v[r]), t[r], sdot);
      return Tr<T>::rr(y + Tr<T>::rr(lora_scale * Tr<T>::rr(sdot)));
    };
    for (int m = m_begin; m < m_end; ++m) {
      const int b = m / q_len, i = m % q_len;
      const int slot = ctx0 + i;
      const int p = pos[m];
      T* row = qkv + (int64_t)m * ldq;
      float tq[8], tv[8];
      if (R == 8) {
        const Vec8<T> a = ld16(row + 3 * H), bb = ld16(row + 3 * H + 8);
#pragma unroll
        for (int r = 0; r < 8; ++r) { tq[r] = Tr<T>::f(a.v[r]); tv[r] = Tr<T>::f(bb.v[r]); }
      }
      const Vec8<T> c_lo = ld16(cos_t + (int64_t)p * HD + d0), c_hi = ld16(cos_t + (int64_t)p * HD + d0 + HALF);
      const Vec8<T> s_lo = ld16(sin_t + (int64_t)p * HD + d0), s_hi = ld16(sin_t + (int64_t)p * HD + d0 + HALF);
      T* qp = row + h * HD + d0;
      const T* kp = row + H + h * HD + d0;
      const T* vp = row + 2 * H + h * HD + d0;
      const Vec8<T> q_lo = ld16(qp), q_hi = ld16(qp + HALF), k_lo = ld16(kp), k_hi = ld16(kp + HALF), v_lo = ld16(vp), v_hi = ld16(vp + HALF);
      Vec8<T> oq_lo, oq_hi, ok_lo, ok_hi, ov_lo, ov_hi;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float cl = Tr<T>::f(c_lo.v[e]), ch = Tr<T>::f(c_hi.v[e]), sl = Tr<T>::f(s_lo.v[e]), sh = Tr<T>::f(s_hi.v[e]);
        float lo = Tr<T>::f(q_lo.v[e]), hi = Tr<T>::f(q_hi.v[e]);
        if (R == 8) { lo = lora(lo, bq_lo[e], tq); hi = lora(hi, bq_hi[e], tq); }
        oq_lo.v[e] = Tr<T>::r(Tr<T>::rr(lo * cl) + Tr<T>::rr(-hi * sl));       // q*cos + rotate_half(q)*sin
        oq_hi.v[e] = Tr<T>::r(Tr<T>::rr(hi * ch) + Tr<T>::rr(lo * sh));
        lo = Tr<T>::f(k_lo.v[e]); hi = Tr<T>::f(k_hi.v[e]);
        ok_lo.v[e] = Tr<T>::r(Tr<T>::rr(lo * cl) + Tr<T>::rr(-hi * sl));
        ok_hi.v[e] = Tr<T>::r(Tr<T>::rr(hi * ch) + Tr<T>::rr(lo * sh));
        lo = Tr<T>::f(v_lo.v[e]); hi = Tr<T>::f(v_hi.v[e]);
        if (R == 8) { lo = lora(lo, bv_lo[e], tv); hi = lora(hi, bv_hi[e], tv); }
        ov_lo.v[e] = Tr<T>::r(lo);
        ov_hi.v[e] = Tr<T>::r(hi);
      }
      const int64_t cache_off = (((int64_t)b * nh + h) * cmax + slot) * HD + d0;
      *reinterpret_cast<uint4*>(qp) = *reinterpret_cast<const uint4*>(&oq_lo);
      *reinterpret_cast<uint4*>(qp + HALF) = *reinterpret_cast<const uint4*>(&oq_hi);
      *reinterpret_cast<uint4*>(kc + cache_off) = *reinterpret_cast<const uint4*>(&ok_lo);
      *reinterpret_cast<uint4*>(kc + cache_off + HALF) = *reinterpret_cast<const uint4*>(&ok_hi);
      *reinterpret_cast<uint4*>(vc + cache_off) = *reinterpret_cast<const uint4*>(&ov_lo);
      *reinterpret_cast<uint4*>(vc + cache_off + HALF) = *reinterpret_cast<const uint4*>(&ov_hi);
    }
  }
}

extern "C" int rd_rope_kv_store(void* qkv, int64_t ldq, const int32_t* pos, const int32_t* ctx_len, const void* cos_t,
                                const void* sin_t, void* kc, void* vc, int B, int q_len, int nh, int hd, int cmax,
                                const void* lora_b, int lora_r, float lora_scale, int dtype, void* stream) {
  RD_REQUIRE(B > 0 && q_len > 0 && hd % 2 == 0, "rd_rope_kv_store: bad shape");
  RD_REQUIRE(lora_b == nullptr || (lora_r > 0 && lora_r <= 16), "rd_rope_kv_store: lora_r must be in [1,16] (got %d)", lora_r);
  RD_REQUIRE(ldq >= 3 * (int64_t)nh * hd + 2 * (lora_b ? lora_r : 0), "rd_rope_kv_store: ldq %lld too small", (long long)ldq);
  RD_DISPATCH_DTYPE(dtype, T, {
    const int lr = lora_b ? lora_r : 0;
    if (hd == 128 && (lr == 0 || lr == 8) && ldq % 8 == 0) {
      // token groups: ~2 CTAs per SM when there are enough tokens, never more than 8 tokens (a serial chain) per CTA
      const int M = B * q_len;
      int tpc = (M + 295) / 296;
      tpc = tpc < 1 ? 1 : (tpc > 8 ? 8 : tpc);
      const int grid = (M + tpc - 1) / tpc;
      if (lr == 8) {
        RD_CHECK_CUDA(rd_launch(rope_kv_store_vec_kernel<T, 8>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, rd_pdl_enabled(),
                                (T*)qkv, ldq, pos, ctx_len, (const T*)cos_t, (const T*)sin_t, (T*)kc, (T*)vc, q_len, nh, cmax,
                                (const T*)lora_b, lora_scale, M, tpc));
      } else {
        RD_CHECK_CUDA(rd_launch(rope_kv_store_vec_kernel<T, 0>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, rd_pdl_enabled(),
                                (T*)qkv, ldq, pos, ctx_len, (const T*)cos_t, (const T*)sin_t, (T*)kc, (T*)vc, q_len, nh, cmax,
                                (const T*)lora_b, lora_scale, M, tpc));
      }
      return RD_OK;
    }
    RD_CHECK_CUDA(rd_launch(rope_kv_store_kernel<T>, dim3(B * q_len), dim3(256), 0, (cudaStream_t)stream, rd_pdl_enabled(),
                            (T*)qkv, ldq, pos, ctx_len, (const T*)cos_t, (const T*)sin_t, (T*)kc, (T*)vc, q_len, nh, hd, cmax,
                            (const T*)lora_b, lora_b ? lora_r : 0, lora_scale));
    return RD_OK;
  });
}

// ------------------------------------------------------------------------------------------------
// Attention against the flat KV cache                               modeling_llama_imgemb.py:216-234
// One CTA per (query row, head, batch row); 8 key-groups x 16 lanes, each lane owns 8 of the 128 dims
// (one 128-bit load per key per lane).  Scores are materialised in shared memory so the softmax is the
// reference's two-pass fp32 softmax with probabilities rounded to the storage dtype before P.V.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_THREADS = 128;
constexpr int ATT_GROUPS = ATT_THREADS / 16;
constexpr int ATT_U = 3;

// FUSED (decode, q_len == 1): the CTA first applies RoPE to its head's q and k, appends k,v to the cache at slot ctx
// and keeps the three 128-vectors in shared memory, so the single-token step needs no separate rope/append launch.
template <class T, bool FUSED>
__global__ void __launch_bounds__(ATT_THREADS, 7)
attention_kernel(const T* __restrict__ qkv, int64_t ldq, T* __restrict__ kc, T* __restrict__ vc,
                 const uint8_t* __restrict__ keymask, const int32_t* __restrict__ ctx_len_p, T* __restrict__ out,
                 int q_len, int nh, int cmax, const int32_t* __restrict__ pos, const T* __restrict__ cos_t,
                 const T* __restrict__ sin_t) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128;
  __shared__ float s_q[FUSED ? HD : 1], s_k[FUSED ? HD : 1], s_v[FUSED ? HD : 1];
  extern __shared__ float sc[];                 // scores / probabilities [c_tot]
  __shared__ float sred[ATT_THREADS / 32];
  __shared__ float spart[ATT_GROUPS][HD];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 4, l16 = tid & 15;
  const int ctx = ctx_len_p[0];
  const int c_tot = ctx + q_len;
  const int jcausal = ctx + i;                  // last key this query may see
  const uint8_t* km = keymask + (int64_t)b * cmax;
  const float lowest = Tr<T>::lowest();

  int any = 1;                                  // decode: the token just selected always attends to itself
  if (!FUSED) {
    any = 0;
    for (int j = tid; j <= jcausal; j += ATT_THREADS) any |= km[j];
    any = __syncthreads_or(any);
  }
  // rows with at least one visible key: masked keys past the causal limit contribute exp(min - max) == 0 exactly,
  // so they are skipped.  All-masked (left-pad) rows are evaluated literally over every key like the reference.
  const int jend = any ? (jcausal + 1) : c_tot;

  const int64_t m = (int64_t)b * q_len + i;
  const T* kbase = kc + ((int64_t)b * nh + h) * cmax * HD;
  const T* vbase = vc + ((int64_t)b * nh + h) * cmax * HD;
  float q[8];
  if (FUSED) {
    constexpr int half = HD / 2;
    const int H = nh * HD;
    const T* row = qkv + m * ldq;
    const int64_t slot_off = (((int64_t)b * nh + h) * cmax + ctx) * HD;
    if (tid < half) {
      const int d = tid, p = pos[b];
      const float c_lo = Tr<T>::f(cos_t[(int64_t)p * HD + d]), c_hi = Tr<T>::f(cos_t[(int64_t)p * HD + d + half]);
      const float s_lo = Tr<T>::f(sin_t[(int64_t)p * HD + d]), s_hi = Tr<T>::f(sin_t[(int64_t)p * HD + d + half]);
      float lo = Tr<T>::f(row[h * HD + d]), hi = Tr<T>::f(row[h * HD + d + half]);
      s_q[d] = Tr<T>::rr(Tr<T>::rr(lo * c_lo) + Tr<T>::rr(-hi * s_lo));            // q*cos + rotate_half(q)*sin
      s_q[d + half] = Tr<T>::rr(Tr<T>::rr(hi * c_hi) + Tr<T>::rr(lo * s_hi));
      lo = Tr<T>::f(row[H + h * HD + d]); hi = Tr<T>::f(row[H + h * HD + d + half]);
      const float k_lo = Tr<T>::rr(Tr<T>::rr(lo * c_lo) + Tr<T>::rr(-hi * s_lo));
      const float k_hi = Tr<T>::rr(Tr<T>::rr(hi * c_hi) + Tr<T>::rr(lo * s_hi));
      s_k[d] = k_lo; s_k[d + half] = k_hi;
      kc[slot_off + d] = Tr<T>::r(k_lo); kc[slot_off + d + half] = Tr<T>::r(k_hi);
    } else {
      const int d = tid - half;
      const T v_lo = row[2 * H + h * HD + d], v_hi = row[2 * H + h * HD + d + half];
      s_v[d] = Tr<T>::f(v_lo); s_v[d + half] = Tr<T>::f(v_hi);
      vc[slot_off + d] = v_lo; vc[slot_off + d + half] = v_hi;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = s_q[l16 * 8 + e];
  } else {
    Vec8<T> qv = ld16(qkv + m * ldq + h * HD + l16 * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = Tr<T>::f(qv.v[e]);
  }
  // key/value row j: from the cache, except the row this CTA has just appended (not visible through the .nc path)
  auto load_kv = [&](const T* base, const float* fresh, int j) {
    Vec8<T> r;
    if (FUSED && j == ctx) {
#pragma unroll
      for (int e = 0; e < 8; ++e) r.v[e] = Tr<T>::r(fresh[l16 * 8 + e]);
    } else {
      r = ld_stream16(base + (int64_t)j * HD + l16 * 8);
    }
    return r;
  };
  const float sqrt_d = 11.313708498984761f;  // math.sqrt(128)
  const unsigned hmask = 0xFFFFu << (lane & 16);
  // The KV sweep is latency bound (a few keys per 16-lane group): software-pipelined, ATT_U keys per group and
  // iteration with the next batch's 128-bit loads already in flight while the current batch is reduced.
  constexpr int STEP = ATT_GROUPS * ATT_U;
  auto load_batch = [&](const T* base, const float* fresh, int jb, Vec8<T>(&dst)[ATT_U]) {
#pragma unroll
    for (int u = 0; u < ATT_U; ++u) {
      const int j = jb + u * ATT_GROUPS;
      if (j < jend) dst[u] = load_kv(base, fresh, j);
    }
  };
  {
    Vec8<T> cur[ATT_U];
    if (g < jend) load_batch(kbase, s_k, g, cur);
    for (int jb = g; jb < jend; jb += STEP) {
      Vec8<T> nxt[ATT_U];
      if (jb + STEP < jend) load_batch(kbase, s_k, jb + STEP, nxt);
      float d[ATT_U];
#pragma unroll
      for (int u = 0; u < ATT_U; ++u) {
        d[u] = 0.f;
        if (jb + u * ATT_GROUPS < jend) {
#pragma unroll
          for (int e = 0; e < 8; ++e) d[u] = fmaf(q[e], Tr<T>::f(cur[u].v[e]), d[u]);
        }
      }
      // the two 16-lane key groups of a warp may run different trip counts: shuffle within the half-warp only
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < ATT_U; ++u) d[u] += __shfl_xor_sync(hmask, d[u], o);
      }
      if (l16 == 0) {
#pragma unroll
        for (int u = 0; u < ATT_U; ++u) {
          const int j = jb + u * ATT_GROUPS;
          if (j < jend) {
            float s = Tr<T>::rr(d[u]);                             // matmul output in the storage dtype
            s = Tr<T>::rr(s / sqrt_d);                             // / math.sqrt(head_dim)
            float madd = km[j] ? 0.f : lowest;                     // _expand_mask
            if (q_len > 1 && j > jcausal) madd = Tr<T>::rr(madd + lowest);   // + _make_causal_mask (may be -inf)
            s = Tr<T>::rr(s + madd);
            s = fmaxf(s, lowest);                                  // torch.max(attn_weights, finfo.min)
            sc[j] = s;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < ATT_U; ++u) cur[u] = nxt[u];
    }
  }
  // first batch of V rows is requested before the softmax so its latency hides behind the reductions
  Vec8<T> vcur[ATT_U];
  if (g < jend) load_batch(vbase, s_v, g, vcur);
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < jend; j += ATT_THREADS) mx = fmaxf(mx, sc[j]);
  mx = warp_max(mx);
  if (lane == 0) sred[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(sred[0], sred[1]), fmaxf(sred[2], sred[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < jend; j += ATT_THREADS) { float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) sred[warp] = sum;
  __syncthreads();
  sum = (sred[0] + sred[1]) + (sred[2] + sred[3]);
  for (int j = tid; j < jend; j += ATT_THREADS) sc[j] = Tr<T>::rr(sc[j] / sum);   // softmax(fp32).to(dtype)
  __syncthreads();

  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int jb = g; jb < jend; jb += STEP) {
    Vec8<T> vnxt[ATT_U];
    if (jb + STEP < jend) load_batch(vbase, s_v, jb + STEP, vnxt);
#pragma unroll
    for (int u = 0; u < ATT_U; ++u) {
      const int j = jb + u * ATT_GROUPS;
      if (j < jend) {
        const float pj = sc[j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, Tr<T>::f(vcur[u].v[e]), acc[e]);
      }
    }
#pragma unroll
    for (int u = 0; u < ATT_U; ++u) vcur[u] = vnxt[u];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) spart[g][l16 * 8 + e] = acc[e];
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int gg = 0; gg < ATT_GROUPS; ++gg) o += spart[gg][tid];
    out[m * (int64_t)(nh * HD) + h * HD + tid] = Tr<T>::r(o);
  }
}

// ------------------------------------------------------------------------------------------------
// Prefill / extend attention (q_len > 1) with the K and V rows of one (sequence, head) staged ONCE in shared memory:
// one CTA per (head, sequence), every warp takes query rows warp, warp + 8, ...  The kernel above re-reads up to c keys
// and values from L2 for every query row (64x redundant at T = 64: 306 us per layer at B = 32); here they are read once.
// Same rounding points (modeling_llama_imgemb.py:216-234): scores T(q.k) -> T(/sqrt(d)) -> T(+ mask) -> max(finfo.min),
// fp32 softmax rounded to the storage dtype before P.V; rows with no visible key are evaluated over every key.
// ------------------------------------------------------------------------------------------------
constexpr int ATP_THREADS = 256, ATP_WARPS = ATP_THREADS / 32;

template <class T>
__global__ void __launch_bounds__(ATP_THREADS, 2)
attention_prefill_kernel(const T* __restrict__ qkv, int64_t ldq, const T* __restrict__ kc, const T* __restrict__ vc,
                         const uint8_t* __restrict__ keymask, const int32_t* __restrict__ ctx_len_p, T* __restrict__ out,
                         int q_len, int nh, int cmax) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128;
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g2 = lane >> 4, l16 = lane & 15;
  const int ctx = ctx_len_p[0];
  const int c_tot = ctx + q_len;
  T* ks = reinterpret_cast<T*>(smem_att);                               // [c_tot][128]
  T* vs = ks + (size_t)cmax * HD;                                       // [c_tot][128]
  float* sc = reinterpret_cast<float*>(vs + (size_t)cmax * HD) + (size_t)warp * cmax;    // [warps][cmax]
  __shared__ int s_first;
  const uint8_t* km = keymask + (int64_t)b * cmax;
  const T* kbase = kc + ((int64_t)b * nh + h) * cmax * HD;
  const T* vbase = vc + ((int64_t)b * nh + h) * cmax * HD;
  if (tid == 0) s_first = 0x7fffffff;
  __syncthreads();
  for (int idx = tid; idx < c_tot * 16; idx += ATP_THREADS) {           // 16-byte chunks, coalesced
    reinterpret_cast<uint4*>(ks)[idx] = *reinterpret_cast<const uint4*>(kbase + (size_t)idx * 8);
    reinterpret_cast<uint4*>(vs)[idx] = *reinterpret_cast<const uint4*>(vbase + (size_t)idx * 8);
  }
  int first = 0x7fffffff;                                               // first key that is not padding
  for (int j = tid; j < c_tot; j += ATP_THREADS)
    if (km[j]) { first = j; break; }
  if (first != 0x7fffffff) atomicMin(&s_first, first);
  __syncthreads();
  first = s_first;
  const float lowest = Tr<T>::lowest();
  const float sqrt_d = 11.313708498984761f;  // math.sqrt(128)
  const unsigned hmask = 0xFFFFu << (lane & 16);

  for (int i = warp; i < q_len; i += ATP_WARPS) {
    const int jcausal = ctx + i;
    const bool any = first <= jcausal;
    const int jend = any ? (jcausal + 1) : c_tot;
    const int64_t m = (int64_t)b * q_len + i;
    float q[8];
    {
      const Vec8<T> qv = ld16(qkv + m * ldq + h * HD + l16 * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) q[e] = Tr<T>::f(qv.v[e]);
    }
    // ---- scores: one key per half-warp and iteration, four keys in flight per half-warp ----
    float mx = -INFINITY;
    for (int jb = 0; jb < jend; jb += 8) {
      float d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jb + 2 * u + g2;
        d[u] = 0.f;
        if (j < jend) {
          const Vec8<T> kk = *reinterpret_cast<const Vec8<T>*>(ks + (size_t)j * HD + l16 * 8);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[u] = fmaf(q[e], Tr<T>::f(kk.v[e]), d[u]);
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) d[u] += __shfl_xor_sync(hmask, d[u], o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jb + 2 * u + g2;
        if (j < jend) {
          float s = Tr<T>::rr(d[u]);                             // matmul output in the storage dtype
          s = Tr<T>::rr(s / sqrt_d);                             // / math.sqrt(head_dim)
          float madd = km[j] ? 0.f : lowest;                     // _expand_mask
          if (j > jcausal) madd = Tr<T>::rr(madd + lowest);      // + _make_causal_mask (may be -inf)
          s = Tr<T>::rr(s + madd);
          s = fmaxf(s, lowest);                                  // torch.max(attn_weights, finfo.min)
          if (l16 == 0) sc[j] = s;
          mx = fmaxf(mx, s);
        }
      }
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < jend; j += 32) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
    sum = warp_sum(sum);
    __syncwarp();
    for (int j = lane; j < jend; j += 32) sc[j] = Tr<T>::rr(sc[j] / sum);   // softmax(fp32).to(dtype)
    __syncwarp();
    // ---- P.V ----
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = g2; j < jend; j += 2) {
      const Vec8<T> vv = *reinterpret_cast<const Vec8<T>*>(vs + (size_t)j * HD + l16 * 8);
      const float pj = sc[j];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, Tr<T>::f(vv.v[e]), acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_down_sync(0xffffffffu, acc[e], 16);
    if (lane < 16) {
      Vec8<T> o;
#pragma unroll
      for (int e = 0; e < 8; ++e) o.v[e] = Tr<T>::r(acc[e]);
      *reinterpret_cast<uint4*>(out + m * (int64_t)(nh * HD) + h * HD + l16 * 8) = *reinterpret_cast<const uint4*>(&o);
    }
    __syncwarp();
  }
}

int rd_attention_prefill_tc(const void* qkv, int64_t ldq, const void* kc, const void* vc, const uint8_t* keymask, const int32_t* ctx_len,
                            int ctx_upper_bound, void* out, int B, int q_len, int nh, int cmax, int dtype, cudaStream_t st);

// ctx_upper_bound: host-known upper bound of ctx_len[0] (-1 = unknown: the cache capacity is assumed); it sizes the shared
// memory / TMEM of the tensor-core kernel, which serves every q_len >= 4 launch whose keys fit one UMMA tile (<= 256).
int rd_attention_bounded(const void* qkv, int64_t ldq, const void* kc, const void* vc, const uint8_t* keymask,
                         const int32_t* ctx_len, int ctx_upper_bound, void* out, int B, int q_len, int nh, int hd, int cmax, int dtype,
                         void* stream) {
  RD_REQUIRE(hd == 128, "rd_attention: head_dim must be 128 (Vicuna-7B); got %d", hd);
  RD_REQUIRE(B > 0 && q_len > 0 && cmax > 0 && cmax * 4 <= 160 * 1024, "rd_attention: bad shape");
  if (q_len >= 4) {
    const int r = rd_attention_prefill_tc(qkv, ldq, kc, vc, keymask, ctx_len, ctx_upper_bound, out, B, q_len, nh, cmax, dtype, (cudaStream_t)stream);
    if (r != 0) return r < 0 ? r : RD_OK;
  }
  // several query rows per (sequence, head) and a cache that fits shared memory: K / V staged once per CTA
  const size_t atp_smem = (size_t)cmax * 128 * 2 * 2 + (size_t)ATP_WARPS * cmax * 4;
  if (q_len >= 4 && atp_smem <= 200 * 1024) {
    RD_DISPATCH_DTYPE(dtype, T, {
      RD_SMEM_ATTR_ONCE(200 * 1024, attention_prefill_kernel<T>);
      RD_CHECK_CUDA(rd_launch(attention_prefill_kernel<T>, dim3(nh, B), dim3(ATP_THREADS), atp_smem, (cudaStream_t)stream, rd_pdl_enabled(),
                              (const T*)qkv, ldq, (const T*)kc, (const T*)vc, keymask, ctx_len, (T*)out, q_len, nh, cmax));
      return RD_OK;
    });
  }
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_SMEM_ATTR_ONCE(160 * 1024, attention_kernel<T, false>);
    RD_CHECK_CUDA(rd_launch(attention_kernel<T, false>, dim3(q_len, nh, B), dim3(ATT_THREADS), (size_t)cmax * 4, (cudaStream_t)stream,
                            rd_pdl_enabled(), (const T*)qkv, ldq, (T*)kc, (T*)vc, keymask, ctx_len, (T*)out, q_len, nh, cmax,
                            (const int32_t*)nullptr, (const T*)nullptr, (const T*)nullptr));
    return RD_OK;
  });
}

extern "C" int rd_attention(const void* qkv, int64_t ldq, const void* kc, const void* vc, const uint8_t* keymask,
                            const int32_t* ctx_len, void* out, int B, int q_len, int nh, int hd, int cmax, int dtype,
                            void* stream) {
  return rd_attention_bounded(qkv, ldq, kc, vc, keymask, ctx_len, -1, out, B, q_len, nh, hd, cmax, dtype, stream);
}

// ------------------------------------------------------------------------------------------------
// Embedding gather with <IMG> splice                     modeling_llama_imgemb.py:498-520, 581-588
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(128)
embed_splice_kernel(const int64_t* __restrict__ ids, const T* __restrict__ embed, const T* __restrict__ img,
                    T* __restrict__ out, int Tlen, int H, int vocab, int img_id, int n_img) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, b = m / Tlen, t = m % Tlen;
  __shared__ int s_first;
  const T* src;
  if (img != nullptr) {
    if (threadIdx.x == 0) s_first = 0x7fffffff;
    __syncthreads();
    int first = 0x7fffffff;
    for (int j = threadIdx.x; j < Tlen; j += blockDim.x)
      if (ids[(int64_t)b * Tlen + j] == img_id) first = min(first, j);
    if (first != 0x7fffffff) atomicMin(&s_first, first);
    __syncthreads();
    const int p = (s_first == 0x7fffffff) ? 0 : s_first;   // rows without <IMG> default to position 0 (:507-510)
    if (t >= p && t < p + n_img) src = img + ((int64_t)b * n_img + (t - p)) * H;
    else {
      int64_t id = ids[m];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
      src = embed + id * H;
    }
  } else {
    int64_t id = ids[m];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    src = embed + id * H;
  }
  for (int k = threadIdx.x * 8; k < H; k += blockDim.x * 8)
    *reinterpret_cast<uint4*>(out + (int64_t)m * H + k) = *reinterpret_cast<const uint4*>(src + k);
}

extern "C" int rd_embed_splice(const int64_t* ids, const void* embed, const void* img, void* out, int B, int T_, int H,
                               int vocab, int dtype, void* stream) {
  RD_REQUIRE(B > 0 && T_ > 0 && H % 8 == 0, "rd_embed_splice: bad shape");
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(embed_splice_kernel<T>, dim3(B * T_), dim3(128), 0, (cudaStream_t)stream, rd_pdl_enabled(),
                            ids, (const T*)embed, (const T*)img, (T*)out, T_, H, vocab, 32000, 32));
    return RD_OK;
  });
}

// Single-token step: embedding rows of the current tokens + the cos / sin rows of every sequence's position gathered into
// rope_rows[B][2][hd] - the attention kernels of all 32 layers then read them with one load instead of pos[b] -> table[pos]
// (two dependent L2 round trips at the head of every layer's attention prologue).
template <class T>
__global__ void __launch_bounds__(128)
embed_decode_kernel(const int64_t* __restrict__ ids, const T* __restrict__ embed, T* __restrict__ out, int H, int vocab,
                    const int32_t* __restrict__ pos, const T* __restrict__ cos_t, const T* __restrict__ sin_t, T* __restrict__ rope_rows, int hd) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  int64_t id = ids[b];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const T* src = embed + id * H;
  const int p = pos[b];
  for (int k = threadIdx.x * 8; k < H; k += blockDim.x * 8)
    *reinterpret_cast<uint4*>(out + (int64_t)b * H + k) = *reinterpret_cast<const uint4*>(src + k);
  for (int d = threadIdx.x; d < hd; d += blockDim.x) {
    rope_rows[(int64_t)b * 2 * hd + d] = cos_t[(int64_t)p * hd + d];
    rope_rows[(int64_t)b * 2 * hd + hd + d] = sin_t[(int64_t)p * hd + d];
  }
}

int rd_embed_decode(const int64_t* ids, const void* embed, void* out, int B, int H, int vocab, const int32_t* pos, const void* cos_t,
                    const void* sin_t, void* rope_rows, int hd, int dtype, void* stream) {
  RD_REQUIRE(B > 0 && H % 8 == 0 && rope_rows != nullptr, "rd_embed_decode: bad arguments");
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(embed_decode_kernel<T>, dim3(B), dim3(128), 0, (cudaStream_t)stream, rd_pdl_enabled(), ids, (const T*)embed, (T*)out,
                            H, vocab, pos, (const T*)cos_t, (const T*)sin_t, (T*)rope_rows, hd));
    return RD_OK;
  });
}

// ------------------------------------------------------------------------------------------------
// Generation bookkeeping
// ------------------------------------------------------------------------------------------------
// attention_mask = ids != pad (HF 4.28.1 generate infers it; test.py:304 relies on that); position_ids =
// cumsum(mask)-1 with pads forced to 1 (prepare_inputs_for_generation, modeling_llama_imgemb.py:804-808).
// One thread per row (T is at most a few hundred and this runs once per prefill).
__global__ void llm_prep_kernel(const int64_t* __restrict__ ids, uint8_t* __restrict__ keymask, int32_t* __restrict__ pos,
                                int32_t* __restrict__ npos, const int32_t* __restrict__ ctx_len, int B, int Tlen, int cmax,
                                int pad_id) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int c0 = ctx_len[0];
  int n = npos[b];
  for (int t = 0; t < Tlen; ++t) {
    const int mk = ids[(int64_t)b * Tlen + t] != pad_id;
    keymask[(int64_t)b * cmax + c0 + t] = (uint8_t)mk;
    n += mk;
    pos[(int64_t)b * Tlen + t] = mk ? (n - 1) : 1;
  }
  npos[b] = n;
}

extern "C" int rd_llm_prep(const int64_t* ids, uint8_t* keymask, int32_t* pos, int32_t* npos, const int32_t* ctx_len,
                           int B, int T_, int cmax, int pad_id, void* stream) {
  RD_CHECK_CUDA(rd_launch(llm_prep_kernel, dim3((B + 63) / 64), dim3(64), 0, (cudaStream_t)stream, rd_pdl_enabled(), ids,
                          keymask, pos, npos, ctx_len, B, T_, cmax, pad_id));
  return RD_OK;
}

// Greedy selection of HF 4.28.1 greedy_search on the last-position logits: argmax in the storage dtype (first
// index on ties), finished rows emit pad, unfinished &= (tok != eos), attention mask grows by a column of ones,
// next position = number of attended tokens so far.  The last CTA to finish advances the shared counters, so the
// whole decode step has no host-visible state (CUDA-graph replayable).
template <class T>
__global__ void __launch_bounds__(1024)
argmax_step_kernel(const T* __restrict__ logits, int64_t ld, int V, int64_t* __restrict__ cur_tok,
                   int64_t* __restrict__ gen, int64_t ldgen, int32_t* __restrict__ finished, uint8_t* __restrict__ keymask,
                   int cmax, int32_t* __restrict__ pos_cur, int32_t* __restrict__ npos, int32_t* __restrict__ ctx_len,
                   int32_t* __restrict__ n_gen, uint32_t* __restrict__ done_ctr, int q_len, int pad_id, int eos_id,
                   int suppress_eos) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* row = logits + (int64_t)b * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = tid; j < V; j += 1024) {
    float v = Tr<T>::f(row[j]);
    if (suppress_eos && j == eos_id) v = Tr<T>::lowest();
    if (v > best || (v == best && j < bi)) { best = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  __shared__ float sv[32];
  __shared__ int si[32];
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = sv[lane]; bi = si[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      const int ctx = ctx_len[0], step = n_gen[0];
      const int fin = finished[b];
      const int64_t tok = fin ? (int64_t)pad_id : (int64_t)bi;
      cur_tok[b] = tok;
      gen[(int64_t)b * ldgen + step] = tok;
      finished[b] = fin || (tok == eos_id);
      if (ctx + q_len < cmax) keymask[(int64_t)b * cmax + ctx + q_len] = 1;
      pos_cur[b] = npos[b];
      npos[b] += 1;
      __threadfence();
      const uint32_t prev = atomicAdd(done_ctr, 1u);
      if (prev == gridDim.x - 1) {
        *done_ctr = 0;
        ctx_len[0] = ctx + q_len;
        n_gen[0] = step + 1;
        // HF greedy_search stops once unfinished_sequences.max() == 0: publish that for the host's (lagging, non-blocking) poll
        __threadfence();
        int all_fin = 1;
        for (int r = 0; r < (int)gridDim.x; ++r) all_fin &= (__ldcg(finished + r) != 0);
        done_ctr[1] = all_fin ? (uint32_t)(step + 1) : 0u;     // 0 = still running, else number of tokens generated
      }
    }
  }
}

extern "C" int rd_argmax_step(const void* logits, int64_t ld, int V, int64_t* cur_tok, int64_t* gen, int64_t ldgen,
                              int32_t* finished, uint8_t* keymask, int cmax, int32_t* pos_cur, int32_t* npos,
                              int32_t* ctx_len, int32_t* n_gen, uint32_t* done_ctr, int B, int q_len, int pad_id,
                              int eos_id, int suppress_eos, int dtype, void* stream) {
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(argmax_step_kernel<T>, dim3(B), dim3(1024), 0, (cudaStream_t)stream, rd_pdl_enabled(),
                            (const T*)logits, ld, V, cur_tok, gen, ldgen, finished, keymask, cmax, pos_cur, npos, ctx_len,
                            n_gen, done_ctr, q_len, pad_id, eos_id, suppress_eos));
    return RD_OK;
  });
}


// ------------------------------------------------------------------------------------------------
// Beam search: LlamaForCausalLM._reorder_cache (modeling_llama_imgemb.py:838-843) on the flat KV cache
// new_cache[row r] = old_cache[beam_idx[r]] for every layer; only the ctx_len[0] cached tokens are moved.
// grid (heads, rows, layers); src and dst are distinct buffers (the engine swaps them afterwards).
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
kv_reorder_kernel(const T* __restrict__ src_k, const T* __restrict__ src_v, T* __restrict__ dst_k, T* __restrict__ dst_v,
                  const int32_t* __restrict__ beam_idx, const int32_t* __restrict__ ctx_len, int nh, int cmax, int hd,
                  int64_t layer_elems) {
  const int h = blockIdx.x, r = blockIdx.y, l = blockIdx.z;
  const int src = beam_idx[r];
  const int64_t n16 = (int64_t)ctx_len[0] * hd * (int64_t)sizeof(T) / 16;
  const int64_t so = (int64_t)l * layer_elems + ((int64_t)src * nh + h) * cmax * hd;
  const int64_t dof = (int64_t)l * layer_elems + ((int64_t)r * nh + h) * cmax * hd;
  const uint4* sk = reinterpret_cast<const uint4*>(src_k + so);
  const uint4* sv = reinterpret_cast<const uint4*>(src_v + so);
  uint4* dk = reinterpret_cast<uint4*>(dst_k + dof);
  uint4* dv = reinterpret_cast<uint4*>(dst_v + dof);
  for (int64_t i = threadIdx.x; i < n16; i += blockDim.x) { dk[i] = sk[i]; dv[i] = sv[i]; }
}

int rd_kv_reorder(const void* src_k, const void* src_v, void* dst_k, void* dst_v, const int32_t* beam_idx, const int32_t* ctx_len,
                  int rows, int nh, int cmax, int hd, int layers, int64_t layer_elems, int dtype, void* stream) {
  RD_REQUIRE(rows > 0 && layers > 0 && (hd * 2) % 16 == 0, "rd_kv_reorder: bad shape");
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_CHECK_CUDA(rd_launch(kv_reorder_kernel<T>, dim3(nh, rows, layers), dim3(256), 0, (cudaStream_t)stream, false, (const T*)src_k,
                            (const T*)src_v, (T*)dst_k, (T*)dst_v, beam_idx, ctx_len, nh, cmax, hd, layer_elems));
    return RD_OK;
  });
}
