// Single-token decode attention over the flat KV cache (q_len == 1), fused with RoPE and the KV append.
// Replaces LlamaAttention.forward's cache torch.cat + two bmm + softmax (modeling_llama_imgemb.py:205-234) for the
// decode step; rounding points as in llm_kernels.cu / SURVEY.md Appendix B.
//
// The K rows (and the V rows) of one (sequence, head) are one contiguous [ctx, 128] block of the cache, so the sweep is
// done with bulk asynchronous copies (cp.async.bulk, the 1-D TMA path) through a ring of 32-key shared-memory buffers
// with mbarrier completion instead of per-lane global loads: one elected thread keeps RING copies in flight, the rest of
// the CTA computes from shared memory.  The first RING K chunks only cover cache rows written by EARLIER decode steps,
// so they are requested before griddepcontrol.wait and stream in while the QKV GEMM of this layer is still finishing.
// Safety of that early read.  With programmatic dependent launch every kernel releases its successor at its own start, so the
// pre-wait sections of a whole chain of kernels can run at once - bounded only by SM residency.  The newest cache row was written
// by the previous decode step's attention kernel of the same layer: one step back.  With GEMMs that fill the machine (every
// Vicuna-7B GEMM: 190-260 CTAs x 100 KB of shared memory) at most two or three kernels are co-resident, the writer is > 200
// kernels back and the early read is safe.  With a toy model a whole step fits on the SMs at once and the early read DID return
// stale rows (tests/test_gpu_llm.py eager-vs-graph, found when the pre-wait section grew).  So (1) the caller passes a non-zero
// `prefetch_keys` only for models whose GEMMs fill the machine (engine_llm.cu), and (2) the chunk holding the newest row is never
// part of the early prefetch.  `prefetch_keys` is a host-known lower bound of the context; chunks past it are not requested early.
#include "common.cuh"

bool rd_pdl_enabled();

namespace {

constexpr int HD = 128;
constexpr int CH = 32;                     // keys per chunk: 32 x 256 B = 8 KB
constexpr int CHUNK_BYTES = CH * HD * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done, spins = 0;
  long long t0 = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && (++spins & 0x3FFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) { printf("attention_decode: mbarrier wait timed out (block %d,%d)\n", blockIdx.x, blockIdx.y); __trap(); }
    }
  } while (!done);
}
// K/V rows are read once per step and the next read is 13 GB of traffic later: an L2 evict-first policy keeps the sweep (60-100 MB
// per layer) from pushing out the small data every layer re-reads (split-K slabs, activations, masks, lora_B rows)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull) : "memory");
}

template <class T, int THREADS, int RING>
__global__ void __launch_bounds__(THREADS)
attention_decode_kernel(const T* __restrict__ qkv, int64_t ldq, T* __restrict__ kc, T* __restrict__ vc,
                        const uint8_t* __restrict__ keymask, const int32_t* __restrict__ ctx_len_p, T* __restrict__ out, int nh,
                        int cmax, const int32_t* __restrict__ pos, const T* __restrict__ cos_t, const T* __restrict__ sin_t,
                        int prefetch_keys, const T* __restrict__ lora_b, int lora_r, float lora_scale,
                        const void* pf0, long long pf0_bytes, const void* pf1, long long pf1_bytes,
                        const float* __restrict__ qkv_part, int n_part, long long part_stride, const T* __restrict__ rope_rows) {
  constexpr int GROUPS = THREADS / 16;
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(128) uint8_t smem[];
  T* ring = reinterpret_cast<T*>(smem);                                   // [RING][CH][HD]
  float* sc = reinterpret_cast<float*>(smem + RING * CHUNK_BYTES);        // [cmax + 1]
  __shared__ __align__(8) uint64_t full_bar[RING];
  __shared__ float s_q[HD], s_k[HD], s_v[HD];
  __shared__ float sred[WARPS];
  float (*spart)[HD] = reinterpret_cast<float (*)[HD]>(smem);            // reuses the ring once the sweep is over
  static_assert(GROUPS * HD * 4 <= RING * CHUNK_BYTES, "spart must fit in the ring");

  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 4, l16 = tid & 15;
  const T* kbase = kc + ((int64_t)b * nh + h) * cmax * HD;
  const T* vbase = vc + ((int64_t)b * nh + h) * cmax * HD;

#ifdef RD_ATT_PROF
  long long pt[8]; int pn = 0;
#define APROF() { if (tid == 0 && pn < 8) pt[pn++] = clock64(); }
#else
#define APROF()
#endif
  APROF()
  pdl_launch_dependents();
  {                     // weights of the next GEMMs -> L2 while this latency-bound kernel leaves HBM idle
    const int cta = blockIdx.y * gridDim.x + blockIdx.x, n_ctas = gridDim.x * gridDim.y;
    l2_prefetch_slice(pf0, pf0_bytes, cta, n_ctas, tid, THREADS);
    l2_prefetch_slice(pf1, pf1_bytes, cta, n_ctas, tid, THREADS);
  }
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) mbar_init(&full_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // chunk sequence: K chunks 0..nK-1 then V chunks 0..nV-1; sequence entry s lives in ring slot s % RING
  int n_pre = prefetch_keys > 0 ? (prefetch_keys - 1) / CH : 0;     // whole chunks strictly below the newest cached row
  n_pre = n_pre > RING ? RING : n_pre;
  if (tid == 0) {
    for (int s = 0; s < n_pre; ++s) {            // cache rows of earlier steps: independent of the previous kernel
      mbar_expect_tx(&full_bar[s], CHUNK_BYTES);
      bulk_g2s(ring + (size_t)s * CH * HD, kbase + (size_t)s * CH * HD, CHUNK_BYTES, &full_bar[s]);
    }
  }
  APROF()
  pdl_wait();
  APROF()
  // lora_B rows of the two q (threads 0..63) or two v (threads 64..127) outputs this thread finishes: requested first, so they are in
  // flight together with the rotary-row / projection loads of the prologue instead of after them.  They are an HBM miss (weights last
  // touched a step ago) and the longest item of the prologue; requesting them - or even an L2 prefetch hint for them - BEFORE the PDL
  // wait made eager (non-graph) decode steps of a toy model return different tokens (tools/debug_graph_eager.py; an unrelated
  // constant load or a delay in the same place does not), which is not understood - so nothing touches lora_B before the wait.
  constexpr int half_hd = HD / 2;
  Vec8<T> lb_lo, lb_hi;
  if (lora_r == 8 && tid < 2 * half_hd) {
    const int n_row = tid < half_hd ? h * HD + tid : nh * HD + h * HD + (tid - half_hd);
    lb_lo = ld16_keep(lora_b + (int64_t)n_row * 8);
    lb_hi = ld16_keep(lora_b + (int64_t)(n_row + half_hd) * 8);
  }

  const int ctx = ctx_len_p[0];                  // cached keys; the new token goes to slot ctx
  const int n_kv = (ctx + CH - 1) / CH;          // chunks that hold real keys
  const int nK = n_kv > n_pre ? n_kv : n_pre;    // K entries of the sequence (prefetched-but-unused ones are skipped)
  const int total = nK + n_kv;
  int issued = n_pre;
  auto issue = [&](int s) {                      // tid 0 only
    const bool is_k = s < nK;
    const int c = is_k ? s : s - nK;
    int keys = ctx - c * CH;
    keys = keys > CH ? CH : keys;
    const int slot = s % RING;
    const uint32_t bytes = (uint32_t)keys * HD * 2;
    mbar_expect_tx(&full_bar[slot], bytes);
    bulk_g2s(ring + (size_t)slot * CH * HD, (is_k ? kbase : vbase) + (size_t)c * CH * HD, bytes, &full_bar[slot]);
  };
  if (tid == 0) {
    while (issued < total && issued < RING) { issue(issued); ++issued; }
  }

  // ---- RoPE of this head's q and k, KV append (modeling_llama_imgemb.py:135-142, 209-212) ----------------------------
  {
    constexpr int half = HD / 2;
    const int H = nh * HD;
    const T* row = qkv + (int64_t)b * ldq;
    const int64_t slot_off = (((int64_t)b * nh + h) * cmax + ctx) * HD;
    // q/k/v (and the LoRA t columns) of this row: either the QKV GEMM's rounded output, or - split-K GEMM without a reduction
    // pass - its fp32 partial slabs [n_part][rows][ldq], summed here in split order and rounded once: T(Wx) either way
    const float* prow = qkv_part != nullptr ? qkv_part + (int64_t)b * ldq : nullptr;
    auto ldv = [&](int col) -> float {               // one value (generic path)
      if (prow == nullptr) return Tr<T>::f(row[col]);
      float a = 0.f;
      for (int s = 0; s < n_part; ++s) a += __ldcg(prow + (int64_t)s * part_stride + col);
      return Tr<T>::rr(a);
    };
    // every value this thread needs (2 or 4 projection outputs + the 8 LoRA t columns), requested together: with partial slabs
    // these are L2 round trips, and one dependent chain per value would cost more than the reduction pass it replaces
    auto gather = [&](const int* cols, int n, float* o) {
      if (prow != nullptr && n_part <= 2) {
        float a[12], c[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) if (i < n) a[i] = __ldcg(prow + cols[i]);
#pragma unroll
        for (int i = 0; i < 12; ++i) c[i] = (i < n && n_part == 2) ? __ldcg(prow + part_stride + cols[i]) : 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) if (i < n) o[i] = Tr<T>::rr(a[i] + c[i]);         // (0 + p0) + p1, split order
      } else if (prow == nullptr && n > 4 - 0 && (n == 12 || n == 10)) {
        // rounded GEMM output: the 8 LoRA t columns are one aligned 16-byte load
        const int np = n - 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < np) o[i] = Tr<T>::f(row[cols[i]]);
        const Vec8<T> tv = ld16(row + cols[np]);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[np + i] = Tr<T>::f(tv.v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) if (i < n) o[i] = ldv(cols[i]);
      }
    };
    // peft LoRA (unmerged): the GEMM wrote t = T(lora_A . xn) after the 3H projection columns (q's r values, then v's)
    auto lora8 = [&](float y, const Vec8<T>& bv, const float* t) {
      float sdot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sdot = fmaf(Tr<T>::f(bv.v[i]), t[i], sdot);
      return Tr<T>::rr(y + Tr<T>::rr(lora_scale * Tr<T>::rr(sdot)));
    };
    auto lora_any = [&](float y, int n_row, int tcol) {
      const T* brow = lora_b + (int64_t)n_row * lora_r;
      float sdot = 0.f;
      for (int i = 0; i < lora_r; ++i) sdot = fmaf(Tr<T>::f(brow[i]), ldv(tcol + i), sdot);
      return Tr<T>::rr(y + Tr<T>::rr(lora_scale * Tr<T>::rr(sdot)));
    };
    const bool l8 = lora_r == 8;
    if (tid < half) {
      const int d = tid;
      int cols[12] = {h * HD + d, h * HD + d + half, H + h * HD + d, H + h * HD + d + half, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < 8; ++i) cols[4 + i] = 3 * H + i;
      // cos / sin of this sequence's position: from the per-step gathered rows when the engine provides them (one load, in flight
      // with the projection values below), else through pos[b] (two dependent loads)
      const T* cs = rope_rows != nullptr ? rope_rows + (int64_t)b * 2 * HD : cos_t + (int64_t)pos[b] * HD;
      const T* sn = rope_rows != nullptr ? cs + HD : sin_t + (int64_t)pos[b] * HD;
      const float c_lo = Tr<T>::f(cs[d]), c_hi = Tr<T>::f(cs[d + half]);
      const float s_lo = Tr<T>::f(sn[d]), s_hi = Tr<T>::f(sn[d + half]);
      float v[12];
      gather(cols, l8 ? 12 : 4, v);
      float lo = v[0], hi = v[1];
      const float klo = v[2], khi = v[3];
      if (l8) { lo = lora8(lo, lb_lo, v + 4); hi = lora8(hi, lb_hi, v + 4); }
      else if (lora_r > 0) { lo = lora_any(lo, h * HD + d, 3 * H); hi = lora_any(hi, h * HD + d + half, 3 * H); }
      s_q[d] = Tr<T>::rr(Tr<T>::rr(lo * c_lo) + Tr<T>::rr(-hi * s_lo));
      s_q[d + half] = Tr<T>::rr(Tr<T>::rr(hi * c_hi) + Tr<T>::rr(lo * s_hi));
      const float k_lo = Tr<T>::rr(Tr<T>::rr(klo * c_lo) + Tr<T>::rr(-khi * s_lo));
      const float k_hi = Tr<T>::rr(Tr<T>::rr(khi * c_hi) + Tr<T>::rr(klo * s_hi));
      s_k[d] = k_lo; s_k[d + half] = k_hi;
      kc[slot_off + d] = Tr<T>::r(k_lo); kc[slot_off + d + half] = Tr<T>::r(k_hi);
    } else if (tid < 2 * half) {
      const int d = tid - half;
      int cols[12] = {2 * H + h * HD + d, 2 * H + h * HD + d + half, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < 8; ++i) cols[2 + i] = 3 * H + 8 + i;
      float v[12];
      gather(cols, l8 ? 10 : 2, v);
      float v_lo = v[0], v_hi = v[1];
      if (l8) { v_lo = lora8(v_lo, lb_lo, v + 2); v_hi = lora8(v_hi, lb_hi, v + 2); }
      else if (lora_r > 0) {
        v_lo = lora_any(v_lo, H + h * HD + d, 3 * H + lora_r);
        v_hi = lora_any(v_hi, H + h * HD + d + half, 3 * H + lora_r);
      }
      s_v[d] = v_lo; s_v[d + half] = v_hi;
      vc[slot_off + d] = Tr<T>::r(v_lo); vc[slot_off + d + half] = Tr<T>::r(v_hi);
    }
  }
  __syncthreads();
  APROF()
  float q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) q[e] = s_q[l16 * 8 + e];

  const uint8_t* km = keymask + (int64_t)b * cmax;
  const float lowest = Tr<T>::lowest();
  const float sqrt_d = 11.313708498984761f;      // math.sqrt(128)
  const unsigned hmask = 0xFFFFu << (lane & 16);
  auto score_of = [&](float dot, int j) {          // modeling_llama_imgemb.py:216-230 (decode: padding mask only)
    float s = Tr<T>::rr(dot);
    s = Tr<T>::rr(s / sqrt_d);
    s = Tr<T>::rr(s + (km[j] ? 0.f : lowest));
    return fmaxf(s, lowest);
  };
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  for (int s = 0; s < total; ++s) {
    const int slot = s % RING;
    mbar_wait(&full_bar[slot], (uint32_t)(s / RING) & 1u);
    const T* buf = ring + (size_t)slot * CH * HD;
    if (s < nK) {
      if (s < n_kv) {                              // ---- scores of this K chunk ----
        int keys = ctx - s * CH;
        keys = keys > CH ? CH : keys;
        // the CH / GROUPS keys of a 16-lane group are independent: all dot products first, then the shuffle trees
        // interleaved (one dependent chain per key would run at instruction latency); uniform trip count keeps the
        // half-warp shuffles converged
        constexpr int KPG = CH / GROUPS;
        float dd[KPG];
#pragma unroll
        for (int u = 0; u < KPG; ++u) {
          const int jl = g + u * GROUPS;
          float d = 0.f;
          if (jl < keys) {
            const Vec8<T> kk = *reinterpret_cast<const Vec8<T>*>(buf + jl * HD + l16 * 8);
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(q[e], Tr<T>::f(kk.v[e]), d);
          }
          dd[u] = d;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
          for (int u = 0; u < KPG; ++u) dd[u] += __shfl_xor_sync(hmask, dd[u], o);
        }
#pragma unroll
        for (int u = 0; u < KPG; ++u) {
          const int jl = g + u * GROUPS;
          if (l16 == 0 && jl < keys) sc[s * CH + jl] = score_of(dd[u], s * CH + jl);
        }
      }
    } else {
      const int c = s - nK;
      if (c == 0) {                                // ---- first V chunk: finish the scores, softmax (fp32, rounded) ----
        APROF()
        if (g == 0) {
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) d = fmaf(q[e], s_k[l16 * 8 + e], d);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(hmask, d, o);
          if (l16 == 0) sc[ctx] = score_of(d, ctx);
        }
        __syncthreads();
        float mx = -INFINITY;
        for (int j = tid; j <= ctx; j += THREADS) mx = fmaxf(mx, sc[j]);
        mx = warp_max(mx);
        if (lane == 0) sred[warp] = mx;
        __syncthreads();
        mx = sred[0];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) mx = fmaxf(mx, sred[w]);
        __syncthreads();
        float sum = 0.f;
        for (int j = tid; j <= ctx; j += THREADS) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
        sum = warp_sum(sum);
        if (lane == 0) sred[warp] = sum;
        __syncthreads();
        sum = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) sum += sred[w];
        for (int j = tid; j <= ctx; j += THREADS) sc[j] = Tr<T>::rr(sc[j] / sum);     // softmax(fp32).to(dtype)
        __syncthreads();
        APROF()
      }
      int keys = ctx - c * CH;
      keys = keys > CH ? CH : keys;
#pragma unroll
      for (int u = 0; u < CH / GROUPS; ++u) {      // ---- P.V of this V chunk ----
        const int jl = g + u * GROUPS;
        if (jl < keys) {
          const Vec8<T> vv = *reinterpret_cast<const Vec8<T>*>(buf + jl * HD + l16 * 8);
          const float pj = sc[c * CH + jl];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, Tr<T>::f(vv.v[e]), acc[e]);
        }
      }
    }
    __syncthreads();                               // everyone is done with this ring slot
    if (tid == 0 && issued < total) { issue(issued); ++issued; }
  }
  if (n_kv == 0) {                                 // empty cache: the softmax over the single new key
    if (g == 0) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(q[e], s_k[l16 * 8 + e], d);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(hmask, d, o);
      if (l16 == 0) sc[0] = Tr<T>::rr(1.0f);       // exp(s - s) / 1
    }
    __syncthreads();
  }
  if (g == 0) {                                    // the token just appended
    const float pj = sc[ctx];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, s_v[l16 * 8 + e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) spart[g][l16 * 8 + e] = acc[e];
  __syncthreads();
  APROF()
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int gg = 0; gg < GROUPS; ++gg) o += spart[gg][tid];
    out[(int64_t)b * (nh * HD) + h * HD + tid] = Tr<T>::r(o);
  }
#ifdef RD_ATT_PROF
  APROF()
  if (tid == 0 && ((blockIdx.x == 0 && blockIdx.y == 0) || (blockIdx.x == 17 && blockIdx.y == 20)) && ctx_len_p[0] % 40 == 0)
    printf("att prof cta(%d,%d) ctx %d: pre-wait %lld wait %lld prologue %lld K-loop %lld softmax %lld V-loop %lld out %lld (cycles)\n", blockIdx.x, blockIdx.y,
           ctx_len_p[0], pt[1] - pt[0], pt[2] - pt[1], pt[3] - pt[2], pt[4] - pt[3], pt[5] - pt[4], pt[6] - pt[5], pt[7] - pt[6]);
#endif
}

}  // namespace

static const void* g_rope_rows = nullptr;     // [B][2][128] cos | sin rows of the step (consumed and cleared by the next launch)
extern "C" int rd_attention_decode_set_rope_rows(const void* rows) { g_rope_rows = rows; return RD_OK; }
static const void* g_pf0 = nullptr; static long long g_pf0_bytes = 0;
static const void* g_pf1 = nullptr; static long long g_pf1_bytes = 0;
// weights the next rd_attention_decode launch should pull into L2 (consumed by that launch)
extern "C" int rd_attention_decode_set_l2_prefetch(const void* p0, long long b0, const void* p1, long long b1) {
  g_pf0 = p0; g_pf0_bytes = b0; g_pf1 = p1; g_pf1_bytes = b1; return RD_OK;
}
static int g_attn_prefetch = 1;      // test hook: 0 disables the pre-wait prefetch
extern "C" int rd_attention_decode_set_prefetch(int on) { g_attn_prefetch = on; return RD_OK; }

// Single-token decode: RoPE + KV append + attention in one launch (rd_rope_kv_store + rd_attention with q_len == 1).
// ctx_lower_bound: a host-known lower bound of ctx_len[0] (0 if unknown); only used to size the early prefetch.
static int attention_decode_impl(const void* qkv, int64_t ldq, const int32_t* pos, const void* cos_t, const void* sin_t,
                                 void* kc, void* vc, const uint8_t* keymask, const int32_t* ctx_len, void* out, int B, int nh,
                                 int hd, int cmax, int ctx_lower_bound, const void* lora_b, int lora_r, float lora_scale,
                                 const float* qkv_part, int n_part, long long part_stride, int dtype, void* stream) {
  RD_REQUIRE(hd == 128, "rd_attention_decode: head_dim must be 128 (Vicuna-7B); got %d", hd);
  RD_REQUIRE(B > 0 && nh > 0 && cmax > 0, "rd_attention_decode: bad shape");
  RD_REQUIRE(qkv_part == nullptr || (n_part >= 1 && n_part <= 16), "rd_attention_decode: bad partial count %d", n_part);
  const bool wide = (int64_t)B * nh <= 296;          // few (sequence, head) pairs: more threads per pair
  const int ring = wide ? 4 : 3;                     // 128-thread CTAs: 7 resident per SM so that B=32 x 32 heads is one wave
  const size_t smem = (size_t)ring * CHUNK_BYTES + ((size_t)(cmax + 1) * 4 + 127) / 128 * 128;
  RD_REQUIRE(smem <= 200 * 1024, "rd_attention_decode: cmax %d too large", cmax);
  const int pref = g_attn_prefetch ? (ctx_lower_bound < 0 ? 0 : ctx_lower_bound) : 0;
  RD_DISPATCH_DTYPE(dtype, T, {
    if (wide) {
      RD_SMEM_ATTR_ONCE(200 * 1024, attention_decode_kernel<T, 512, 4>);
      RD_CHECK_CUDA(rd_launch(attention_decode_kernel<T, 512, 4>, dim3(nh, B), dim3(512), smem, (cudaStream_t)stream, rd_pdl_enabled(),
                              (const T*)qkv, ldq, (T*)kc, (T*)vc, keymask, ctx_len, (T*)out, nh, cmax, pos, (const T*)cos_t, (const T*)sin_t, pref,
                              (const T*)lora_b, lora_b ? lora_r : 0, lora_scale, g_pf0, g_pf0_bytes, g_pf1, g_pf1_bytes, qkv_part, n_part, part_stride,
                              (const T*)g_rope_rows));
    } else {
      RD_SMEM_ATTR_ONCE(200 * 1024, attention_decode_kernel<T, 128, 3>);
      RD_CHECK_CUDA(rd_launch(attention_decode_kernel<T, 128, 3>, dim3(nh, B), dim3(128), smem, (cudaStream_t)stream, rd_pdl_enabled(),
                              (const T*)qkv, ldq, (T*)kc, (T*)vc, keymask, ctx_len, (T*)out, nh, cmax, pos, (const T*)cos_t, (const T*)sin_t, pref,
                              (const T*)lora_b, lora_b ? lora_r : 0, lora_scale, g_pf0, g_pf0_bytes, g_pf1, g_pf1_bytes, qkv_part, n_part, part_stride,
                              (const T*)g_rope_rows));
    }
    g_pf0 = nullptr; g_pf1 = nullptr; g_pf0_bytes = 0; g_pf1_bytes = 0; g_rope_rows = nullptr;
    return RD_OK;
  });
}

extern "C" int rd_attention_decode(const void* qkv, int64_t ldq, const int32_t* pos, const void* cos_t, const void* sin_t,
                                   void* kc, void* vc, const uint8_t* keymask, const int32_t* ctx_len, void* out, int B, int nh,
                                   int hd, int cmax, int ctx_lower_bound, const void* lora_b, int lora_r, float lora_scale,
                                   int dtype, void* stream) {
  RD_REQUIRE(qkv != nullptr, "rd_attention_decode: null qkv");
  return attention_decode_impl(qkv, ldq, pos, cos_t, sin_t, kc, vc, keymask, ctx_len, out, B, nh, hd, cmax, ctx_lower_bound, lora_b,
                               lora_r, lora_scale, nullptr, 0, 0, dtype, stream);
}

// Same, with q/k/v (+ LoRA t columns) handed over as the QKV GEMM's fp32 split-K partials [n_part][part_stride] (row b at
// b * ldq): the sum over the splits (fixed order) and the single rounding T(Wx) happen here instead of in a reduction pass.
extern "C" int rd_attention_decode_partials(const float* qkv_part, int n_part, long long part_stride, int64_t ldq, const int32_t* pos,
                                            const void* cos_t, const void* sin_t, void* kc, void* vc, const uint8_t* keymask,
                                            const int32_t* ctx_len, void* out, int B, int nh, int hd, int cmax, int ctx_lower_bound,
                                            const void* lora_b, int lora_r, float lora_scale, int dtype, void* stream) {
  RD_REQUIRE(qkv_part != nullptr, "rd_attention_decode_partials: null partials");
  return attention_decode_impl(nullptr, ldq, pos, cos_t, sin_t, kc, vc, keymask, ctx_len, out, B, nh, hd, cmax, ctx_lower_bound, lora_b,
                               lora_r, lora_scale, qkv_part, n_part, part_stride, dtype, stream);
}
