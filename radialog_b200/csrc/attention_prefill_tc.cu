// Prefill / extend attention (q_len > 1) on the tcgen05 tensor cores.                 modeling_llama_imgemb.py:216-234
//
// One CTA per (head, sequence, 128-query tile), 128 threads.  Both contractions of LlamaAttention.forward run as UMMA tiles
// with fp32 accumulators in TMEM; the reference's rounding points are applied in the TMEM -> register pass between them:
//
//   S[q, key]  = Q[q, :] . K[key, :]          UMMA 128 x keys x 128   (A = Q tile, B = K rows, both K-major, 128B swizzle)
//   s = T(S) -> T(s / sqrt(128)) -> T(s + mask) -> max(s, finfo.min)     one thread owns one query row = one TMEM lane:
//   p = T(softmax_fp32(s))                                               the row softmax needs no shuffles
//   O[q, :]    = P[q, :] . V[:, :]            UMMA 128 x 128 x keys   (A = P written by the softmax threads, B = V^T)
//
// V sits in the cache as [key, 128] (head dim contiguous), i.e. MN-major for the second contraction; the loader transposes
// it on the way into shared memory so that both operands of both UMMAs use the one K-major / SWIZZLE_128B layout the GEMM
// kernels use (linear_tc.cu): rows of 128 B, 8-row groups 1024 B apart, 16-byte chunk index XOR (row & 7).
// Replaces attention_prefill_kernel (SIMT, 216 us per layer at B = 32, T = 64) whenever the keys fit one UMMA (<= 256);
// same masks, same "row with no visible key is evaluated over every key" rule, same rounding contract.
#include "common.cuh"
#include "tc_ptx.cuh"

bool rd_pdl_enabled();

namespace {

using namespace tcptx;

constexpr int HD = 128;
constexpr int QT = 128;             // query rows per CTA (UMMA M)
constexpr int KB = 64;              // elements per 128-byte swizzle row
constexpr int TILE_BYTES = 128 * 128;   // one [128 rows x 64 elements] K-major block

// byte offset of the 16-byte chunk `chunk` (8 elements) of row `r` inside a K-major SWIZZLE_128B block
__device__ __forceinline__ uint32_t sw128(int r, int chunk) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

template <class T>
__global__ void __launch_bounds__(128, 3)
attention_prefill_tc_kernel(const T* __restrict__ qkv, int64_t ldq, const T* __restrict__ kc, const T* __restrict__ vc,
                            const uint8_t* __restrict__ keymask, const int32_t* __restrict__ ctx_len_p, T* __restrict__ out,
                            int q_len, int nh, int cmax, int tmem_cols) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ int s_first;
  __shared__ float s_pad[256];                      // additive padding mask per key: 0 or finfo.min (_expand_mask)

  const int h = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * QT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ctx = ctx_len_p[0];
  const int c_tot = ctx + q_len;
  const int keys_pad = (c_tot + 15) & ~15;          // UMMA N / K granularity
  const int nkb = (keys_pad + KB - 1) / KB;         // 64-key blocks of the second contraction
  const int q_valid = min(QT, q_len - q0);

  uint8_t* sQ = smem;                               // [2 hd blocks][128 rows x 128 B]
  uint8_t* sKP = sQ + 2 * TILE_BYTES;               // K: [2 hd blocks][keys_pad rows x 128 B]; later P: [nkb][128 rows x 128 B]
  uint8_t* sV = sKP + nkb * TILE_BYTES;             // V^T: [nkb key blocks][128 hd rows x 128 B]

  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_first = 0x7fffffff;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- operands of S = Q . K^T: 16-byte chunks, coalesced global reads, swizzled shared-memory writes ----------------------
  const T* kbase = kc + ((int64_t)b * nh + h) * cmax * HD;
  const T* vbase = vc + ((int64_t)b * nh + h) * cmax * HD;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int idx = tid; idx < QT * 16; idx += 128) {                     // Q tile: 128 rows x 16 chunks
    const int r = idx >> 4, c = idx & 15;
    uint4 v = zero4;
    if (r < q_valid) v = *reinterpret_cast<const uint4*>(qkv + ((int64_t)b * q_len + q0 + r) * ldq + h * HD + c * 8);
    *reinterpret_cast<uint4*>(sQ + (c >> 3) * TILE_BYTES + sw128(r, c & 7)) = v;
  }
  const int kblk_bytes = keys_pad * 128;
  for (int idx = tid; idx < keys_pad * 16; idx += 128) {               // K rows
    const int r = idx >> 4, c = idx & 15;
    uint4 v = zero4;
    if (r < c_tot) v = *reinterpret_cast<const uint4*>(kbase + (size_t)r * HD + c * 8);
    *reinterpret_cast<uint4*>(sKP + (c >> 3) * kblk_bytes + sw128(r, c & 7)) = v;
  }
  {                                                                    // first key that is not padding
    const uint8_t* km = keymask + (int64_t)b * cmax;
    int first = 0x7fffffff;
    for (int j = tid; j < c_tot; j += 128)
      if (km[j]) { first = j; break; }
    for (int j = tid; j < keys_pad; j += 128) s_pad[j] = (j < c_tot && km[j]) ? 0.f : Tr<T>::lowest();
    __syncthreads();                                                   // s_first initialised
    if (first != 0x7fffffff) atomicMin(&s_first, first);
  }
  fence_proxy_async_smem();                                            // generic-proxy writes -> visible to the UMMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  const uint32_t idesc_s = make_idesc(Tr<T>::umma_fmt, QT, keys_pad);
  if (warp == 0) {
    if (elect_one()) {
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = make_smem_desc(smem_u32(sQ + kb * TILE_BYTES));
        const uint64_t db = make_smem_desc(smem_u32(sKP + kb * kblk_bytes));
#pragma unroll
        for (int k = 0; k < KB / 16; ++k)
          tc_mma_f16(tmem_base, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc_s, (kb > 0 || k > 0) ? 1u : 0u);
      }
      tc_commit(&mma_bar);
    }
    __syncwarp();
  }

  // ---- while the first UMMA runs: V^T into shared memory.  A thread takes 8 consecutive keys of one head dim (one 16-byte
  // ---- chunk of the K-major block); a warp covers 32 consecutive dims, so every global load instruction reads 64 contiguous bytes
  {
    const int d = tid;                                                 // HD == blockDim.x == 128
    for (int jb = 0; jb < keys_pad; jb += 32) {                        // 4 chunks = 32 independent 2-byte loads in flight
      Vec8<T> v8[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = jb + u * 8 + e;
          v8[u].v[e] = (j < c_tot) ? vbase[(size_t)j * HD + d] : Tr<T>::r(0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j0 = jb + u * 8;
        if (j0 < keys_pad)
          *reinterpret_cast<uint4*>(sV + (j0 >> 6) * TILE_BYTES + sw128(d, (j0 & 63) >> 3)) = *reinterpret_cast<const uint4*>(&v8[u]);
      }
    }
  }

  // ---- scores -> probabilities, one query row per thread ------------------------------------------------------------------------
  mbar_wait(&mma_bar, 0, 1);
  tc_fence_after();
  const int first = s_first;
  const int i = q0 + tid;                                              // query index of this thread's row
  const int jcausal = ctx + i;
  const bool row_ok = tid < q_valid;
  const bool any = first <= jcausal;
  const int jend = row_ok ? (any ? (jcausal + 1) : c_tot) : 0;
  const float lowest = Tr<T>::lowest();
  const float sqrt_d = 11.313708498984761f;                            // math.sqrt(128)
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  auto score = [&](float acc, int j) {
    float s = Tr<T>::rr(acc);                                          // matmul output in the storage dtype
    s = Tr<T>::rr(s / sqrt_d);                                         // / math.sqrt(head_dim)
    float madd = s_pad[j];                                             // _expand_mask
    if (j > jcausal) madd = Tr<T>::rr(madd + lowest);                  // + _make_causal_mask (may be -inf)
    s = Tr<T>::rr(s + madd);
    return fmaxf(s, lowest);                                           // torch.max(attn_weights, finfo.min)
  };
  // Pass 1 reads the accumulator row from TMEM, applies the rounding chain once and parks the (storage-dtype exact) scores in
  // this thread's own row of the P block - the K operand's shared memory, free now that the first UMMA has completed; passes
  // 2 and 3 work from there.  A thread only ever touches its own row, so no barrier is needed between the passes.
  __syncthreads();            // every thread is past the first UMMA's completion: the K block may be overwritten
  float mx = -INFINITY;
  for (int c = 0; c < keys_pad; c += 16) {                             // pass 1: scores + row maximum
    uint32_t r[16];
    tc_ld16(taddr + c, r);
    tc_wait_ld();
    Vec8<T> s16[2];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float sv = lowest;
      if (c + e < jend) { sv = score(__uint_as_float(r[e]), c + e); mx = fmaxf(mx, sv); }
      s16[e >> 3].v[e & 7] = Tr<T>::r(sv);
    }
    uint8_t* blk = sKP + (c >> 6) * TILE_BYTES;
    *reinterpret_cast<uint4*>(blk + sw128(tid, (c & 63) >> 3)) = *reinterpret_cast<const uint4*>(&s16[0]);
    *reinterpret_cast<uint4*>(blk + sw128(tid, ((c & 63) >> 3) + 1)) = *reinterpret_cast<const uint4*>(&s16[1]);
  }
  float sum = 0.f;
  for (int c = 0; c < keys_pad; c += 8) {                              // pass 2: denominator
    const Vec8<T> s8 = *reinterpret_cast<const Vec8<T>*>(sKP + (c >> 6) * TILE_BYTES + sw128(tid, (c & 63) >> 3));
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (c + e < jend) sum += expf(Tr<T>::f(s8.v[e]) - mx);
  }
  for (int c = 0; c < keys_pad; c += 8) {                              // pass 3: p = T(exp / sum) -> A operand of the second UMMA
    uint4* slot = reinterpret_cast<uint4*>(sKP + (c >> 6) * TILE_BYTES + sw128(tid, (c & 63) >> 3));
    const Vec8<T> s8 = *reinterpret_cast<const Vec8<T>*>(slot);
    Vec8<T> p8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float pv = 0.f;
      if (c + e < jend) pv = expf(Tr<T>::f(s8.v[e]) - mx) / sum;
      p8.v[e] = Tr<T>::r(pv);                                          // softmax(fp32).to(dtype)
    }
    *slot = *reinterpret_cast<const uint4*>(&p8);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- O = P . V (the accumulator reuses the TMEM columns of S: every thread has read its scores) --------------------------
  const uint32_t idesc_o = make_idesc(Tr<T>::umma_fmt, QT, HD);
  if (warp == 0) {
    if (elect_one()) {
      for (int ks = 0; ks < keys_pad / 16; ++ks) {
        const int kb = ks >> 2, k = ks & 3;
        const uint64_t da = make_smem_desc(smem_u32(sKP + kb * TILE_BYTES)) + (uint64_t)((k * 32) >> 4);
        const uint64_t db = make_smem_desc(smem_u32(sV + kb * TILE_BYTES)) + (uint64_t)((k * 32) >> 4);
        tc_mma_f16(tmem_base, da, db, idesc_o, ks > 0 ? 1u : 0u);
      }
      tc_commit(&mma_bar);
    }
    __syncwarp();
  }
  mbar_wait(&mma_bar, 1, 2);
  tc_fence_after();
  T* orow = out + ((int64_t)b * q_len + i) * (int64_t)(nh * HD) + h * HD;
  for (int c = 0; c < HD; c += 16) {
    uint32_t r[16];
    tc_ld16(taddr + c, r);
    tc_wait_ld();
    if (row_ok) {
      Vec8<T> o16[2];
#pragma unroll
      for (int e = 0; e < 16; ++e) o16[e >> 3].v[e & 7] = Tr<T>::r(__uint_as_float(r[e]));
      *reinterpret_cast<uint4*>(orow + c) = *reinterpret_cast<const uint4*>(&o16[0]);
      *reinterpret_cast<uint4*>(orow + c + 8) = *reinterpret_cast<const uint4*>(&o16[1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

}  // namespace

static int g_attn_tc = 1;     // test hook: 0 = always the SIMT kernels
extern "C" int rd_attention_set_tensor_core(int on) { g_attn_tc = on; return RD_OK; }

// Returns 1 if the launch was taken (keys fit one UMMA tile), 0 if the caller should use the SIMT kernels, <0 on error.
int rd_attention_prefill_tc(const void* qkv, int64_t ldq, const void* kc, const void* vc, const uint8_t* keymask, const int32_t* ctx_len,
                            int ctx_upper_bound, void* out, int B, int q_len, int nh, int cmax, int dtype, cudaStream_t st) {
  if (!g_attn_tc) return 0;
  // keys of the launch: ctx_len[0] + q_len (device value); the host bound sizes shared memory / TMEM and gates the path
  const int c_max = ctx_upper_bound < 0 ? cmax : (ctx_upper_bound + q_len < cmax ? ctx_upper_bound + q_len : cmax);
  if (c_max > 256 || ldq % 8 != 0) return 0;
  const int keys_pad = (c_max + 15) & ~15, nkb = (keys_pad + 63) / 64;
  const size_t smem = (size_t)(2 + 2 * nkb) * TILE_BYTES + 1024;
  const int tmem_cols = keys_pad <= 128 ? 128 : 256;
  const dim3 grid(nh, B, (q_len + QT - 1) / QT);
  RD_DISPATCH_DTYPE(dtype, T, {
    RD_SMEM_ATTR_ONCE(200 * 1024, attention_prefill_tc_kernel<T>);
    RD_CHECK_CUDA(rd_launch(attention_prefill_tc_kernel<T>, grid, dim3(128), smem, st, rd_pdl_enabled(), (const T*)qkv, ldq, (const T*)kc,
                            (const T*)vc, keymask, ctx_len, (T*)out, q_len, nh, cmax, tmem_cols));
    return 1;
  });
}
