// Persistent decode-layer kernel: ALL transformer layers of one single-token decode step in ONE launch (sm_100a).
//
// Replaces, for q_len == 1 and B <= 32, the per-layer launch sequence of engine_llm.cu
//   rmsnorm -> QKV GEMM -> RoPE+append+attention -> O GEMM(+res) -> rmsnorm -> gate|up GEMM(+SwiGLU) -> down GEMM(+res)
// (LlamaDecoderLayer.forward, modeling_llama_imgemb.py:266-318; rounding points of SURVEY.md Appendix B unchanged).
//
// Why: a decode step is HBM-bound (13.2 GB of weights per step), and with one kernel per op every launch boundary
// drains the weight stream (prologue + split-K tail ~10 us per GEMM, profiles/ncu_s2_linear.md).  Here one CTA per SM
// stays resident for the whole step and a dedicated producer thread streams the weight tiles of ALL phases and layers
// through one TMA ring; it never waits for anything except a free ring slot, so the HBM pipe stays busy across phase
// boundaries while the consumers synchronise.
//
//   warp 0     : weight producer   (cp.async.bulk.tensor 2D, 128B swizzle, evict-first) - runs ahead across phases/layers
//   warp 1     : tcgen05.mma issuer (UMMA 128 x 32, fp32 accumulators in TMEM, 4 accumulator buffers)
//   warp 2     : activation producer (TMA of the [32 x 64] token tiles, gated on the grid barrier of the previous phase)
//   warps 4-11 : workers: RMSNorm applied to the token tiles in shared memory, GEMM epilogues (tcgen05.ld, stream-K
//                fix-up, residual add, SwiGLU, sum-of-squares partials for the next RMSNorm), RoPE + KV append + attention.
//
// Work split: every GEMM phase is cut "stream-K" style: the (weight tile, k-block) units of the phase are dealt to the
// CTAs in equal contiguous runs (host-built table), so all SMs pull the same number of weight bytes whatever the tile
// count (96 / 32 / 86 / 32 tiles vs 148 SMs).  A tile whose k range is shared by several CTAs is reduced through an
// fp32 workspace in L2 by the LAST CTA to finish it, in fixed split order (deterministic).  Phases are separated by a
// grid barrier (one atomic counter); 5 per layer.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "decode_mega.h"
#include "tc_ptx.cuh"

bool rd_pdl_enabled();

namespace {
using namespace tcptx;

constexpr int TILE_N = 128;            // weight rows per tile (UMMA M)
constexpr int BK = 64;                 // k-block: 64 x 2 B = one 128-byte swizzle row
constexpr int NT = 32;                 // token columns (UMMA N)
constexpr int UK = 16;
constexpr int W_SLOT = TILE_N * BK * 2;   // 16 KB
constexpr int X_SLOT = NT * BK * 2;       // 4 KB
constexpr int SW = 8;                  // weight ring slots (128 KB)
constexpr int SX = 20;                 // token-tile ring slots (80 KB); doubles as the attention scratch
// Accumulation chains: tcgen05.mma into ONE accumulator is a dependent chain, and at N = 32 an MMA is so short that the
// chain runs at pipeline latency (~125 ns per MMA measured, 0.5 us per k-block).  Each segment therefore accumulates its
// four k-steps per k-block into FOUR independent 32-column accumulators (gate|up: two each), summed in the epilogue.
constexpr int NCHAIN = 4;
constexpr int NACC = 2, ACC_COLS = NCHAIN * NT, TMEM_COLS = NACC * ACC_COLS;
constexpr int WORK_WARPS = 8, WORKERS = WORK_WARPS * 32, THREADS = 128 + WORKERS;
constexpr int MAX_G = 160, MAX_SEG = 2, MAX_SPLIT = 8;     // MAX_SEG <= NACC: a phase never waits for its own epilogues
constexpr int PART_STRIDE = 2 * NT * TILE_N;      // floats per (tile, split) slab of the stream-K workspace
constexpr int ATT_WARP_BYTES = SX * X_SLOT / WORK_WARPS;     // 10 KB: [scores / partial acc 2 KB][chunk 4 KB][chunk 4 KB]
constexpr int ATT_CH = 16;             // keys per bulk-copy chunk (16 x 256 B = 4 KB)
constexpr int ATT_MAX_CTX = 1023;      // scores of one (sequence, head) live in 2 KB of shared memory as 2-byte values
static_assert(ATT_WARP_BYTES == 2048 + 2 * ATT_CH * 256, "attention scratch layout");

struct Seg { int tile, kb0, kb1, split, nsplits, pad0, pad1, pad2; };
constexpr int SMEM_RING = SW * W_SLOT + SX * X_SLOT;
constexpr int CONS_R = 32;               // ring of "k-block consumed" barriers (one tcgen05.commit per k-block group frees its W and X slots)
constexpr int N_BARS = SW + CONS_R + 2 * SX + 2 * NACC + 2 * WORK_WARPS;
static_assert(CONS_R > SX && CONS_R > SW, "consumed-barrier ring must be longer than both operand rings");
struct CtaSched { int nseg[4]; Seg seg[4][MAX_SEG]; };
constexpr int LNW_KB = 64;                // k-blocks of RMSNorm weights staged per phase (8 KB)
constexpr int SMEM_MISC = LNW_KB * BK * 2 + N_BARS * 8 + 16 + 32 * 4 + 8 * 32 * 4 + WORK_WARPS * 2 * 4 + (int)sizeof(CtaSched) + 64 + SW * 4;
constexpr int SMEM_BYTES = SMEM_RING + 1024 + ((SMEM_MISC + 127) / 128) * 128;

enum { G_QKV = 0, G_O = 1, G_GU = 2, G_DN = 3 };

struct Sched {
  int nseg[4][MAX_G];
  Seg seg[4][MAX_G][MAX_SEG];
};

struct LayerDev {
  const void *ln1, *ln2, *lora_b;
  void *kc, *vc;
};

constexpr int MAX_LAYERS = 48;
// Tensor maps of every layer's weights travel as a __grid_constant__ kernel parameter (constant bank): TMA issue from a
// descriptor in plain global memory measured ~0.8 us per cp.async.bulk.tensor (one L2/DRAM round trip per instruction).
struct WeightMaps { CUtensorMap m[MAX_LAYERS * 4]; };

struct MegaParams {
  const LayerDev* lay;          // [layers]
  const Sched* sched;
  void *x, *qkv, *att, *mid;
  float* ws;                    // stream-K partials [tile][MAX_SPLIT][PART_STRIDE]
  float* ssq;                   // [H/128][32] sum-of-squares partials of the residual stream
  uint32_t* sync;               // [0] grid-barrier counter, [1] exit counter, [32..] per-tile arrival counters
  const uint8_t* keymask;
  const int32_t *ctx_len, *pos;
  const void *cos_t, *sin_t;
  int B, H, I, nh, n_qkv, ldq, cmax, lora_r, l0, l1, G, att_P, sw, dbg, cb;
  float lora_scale, eps;
  unsigned long long* trace;    // development aid: [cta][layer][phase 0..4][8] globaltimer stamps (nullptr = off)
};

__device__ __forceinline__ void trace_stamp(const MegaParams& p, int l, int ph, int ev) {
  if (p.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[(((size_t)blockIdx.x * (p.l1 - p.l0) + (l - p.l0)) * 5 + ph) * 16 + ev] = t;
  }
}
__device__ __forceinline__ int phase_of_g(int g) { return g == 0 ? 0 : g + 1; }

// One lane polls the mbarrier, the rest of the warp waits at the warp barrier: hundreds of threads spinning on
// mbarrier.try_wait saturate the SM's synchronisation unit and slow every other barrier operation of the CTA
// (measured: the MMA issue loop ran at ~0.7 us per k-block with 256 workers polling one accumulator barrier).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int tag, int lane) {
  if (lane == 0) mbar_wait(bar, parity, tag);
  __syncwarp();
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); }
__device__ __forceinline__ void team_bar(int team, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(threads) : "memory");
}

__device__ __forceinline__ void grid_wait(const uint32_t* ctr, uint32_t target, int tag) {
  if (target == 0) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (ld_acquire_gpu(ctr) < target) {
    if ((++spins & 0xFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 6000000000ll) {
        printf("decode_mega: grid barrier timed out (tag %d, block %d, target %u, have %u)\n", tag, blockIdx.x, target, ld_acquire_gpu(ctr));
        __trap();
      }
    }
  }
}

__device__ __forceinline__ int phase_epoch(int l_rel, int g) { return 5 * l_rel + (g == G_QKV ? 0 : g == G_O ? 2 : g == G_GU ? 3 : 4); }

template <class T> __device__ __forceinline__ T ldcg_t(const T* p) {
  const unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
  return *reinterpret_cast<const T*>(&u);
}
template <class T> __device__ __forceinline__ Vec8<T> ldcg16(const T* p) {
  Vec8<T> r;
  *reinterpret_cast<uint4*>(&r) = __ldcg(reinterpret_cast<const uint4*>(p));
  return r;
}

// ------------------------------------------------------------------------------------------------------------
// attention of one decode step for this CTA's (sequence, head) items: RoPE + KV append + softmax(QK^T).V with the
// reference's rounding points (modeling_llama_imgemb.py:205-234; same arithmetic as attention_decode.cu).  A team of
// P warps shares one item: K/V chunks round-robin over the team, scores in the team leader's scratch.
// ------------------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void attention_items(const MegaParams& p, const LayerDev& L, int cta, int ww, int lane, uint8_t* scratch,
                                                uint64_t* att_bar, float* s_red, uint32_t& seq) {
  constexpr int HD = 128;
  const int P = p.att_P, team = ww / P, pw = ww % P, teams = WORK_WARPS / P;
  const int g2 = lane >> 4, l16 = lane & 15;
  const unsigned hmask = 0xFFFFu << (lane & 16);
  T* sc = reinterpret_cast<T*>(scratch + (size_t)(team * P) * ATT_WARP_BYTES);
  float* pacc = reinterpret_cast<float*>(scratch + (size_t)ww * ATT_WARP_BYTES);
  uint8_t* cbuf = scratch + (size_t)ww * ATT_WARP_BYTES + 2048;
  uint64_t* bar = att_bar + ww * 2;
  const int nh = p.nh, H = p.H, cmax = p.cmax, lora_r = p.lora_r;
  const T* kc = reinterpret_cast<const T*>(L.kc);
  const T* vc = reinterpret_cast<const T*>(L.vc);
  const T* lora_b = reinterpret_cast<const T*>(L.lora_b);
  const T* cos_t = reinterpret_cast<const T*>(p.cos_t);
  const T* sin_t = reinterpret_cast<const T*>(p.sin_t);
  const int ctx = p.ctx_len[0];
  const int nK = (ctx + ATT_CH - 1) / ATT_CH;
  const int n_my = nK > pw ? (nK - pw + P - 1) / P : 0;
  const float lowest = Tr<T>::lowest();
  const float sqrt_d = 11.313708498984761f;
  const int nitems = p.B * nh;

  long long tA = 0, tB = 0, tC = 0, tD = 0, tE = 0;
  for (int item = team * p.G + cta; item < nitems; item += teams * p.G) {
    const long long c_start = clock64();
    const int b = item / nh, h = item % nh;
    const T* kbase = kc + ((int64_t)b * nh + h) * cmax * HD;
    const T* vbase = vc + ((int64_t)b * nh + h) * cmax * HD;
    auto issue = [&](int s) {                    // lane 0: chunk s of this warp's K-then-V sequence
      const bool is_k = s < n_my;
      const int c = pw + (is_k ? s : s - n_my) * P;
      int keys = ctx - c * ATT_CH;
      keys = keys > ATT_CH ? ATT_CH : keys;
      const uint32_t q = seq + (uint32_t)s;
      uint64_t* bb = bar + (q & 1u);
      mbar_expect_tx(bb, (uint32_t)keys * HD * 2);
      bulk_g2s(cbuf + (q & 1u) * (ATT_CH * HD * 2), (is_k ? kbase : vbase) + (size_t)c * ATT_CH * HD, (uint32_t)keys * HD * 2, bb);
    };
    if (lane == 0) {
      if (2 * n_my > 0) issue(0);
      if (2 * n_my > 1) issue(1);
    }
    // ---- RoPE of q, of k, the LoRA'd v, KV append -------------------------------------------------------------------
    // The two half-warps split the work: lanes 0-15 produce q (LoRA + RoPE), lanes 16-31 produce k (RoPE) and v (LoRA);
    // every lane owns dims [D, D+8), its RoPE partner dims live in lane ^ 8.  All loads of a lane are independent of
    // each other (one L2 round trip), results are exchanged with shuffles.
    const T* row = reinterpret_cast<const T*>(p.qkv) + (int64_t)b * p.ldq;
    const int D = l16 * 8;
    const bool lo_half = D < 64;
    const int ppos = p.pos[b];
    const Vec8<T> cv = ld16(cos_t + (int64_t)ppos * HD + D), sv = ld16(sin_t + (int64_t)ppos * HD + D);
    const Vec8<T> A = ldcg16(row + (g2 ? H : 0) + h * HD + D);                    // q (lanes 0-15) or k (lanes 16-31)
    const Vec8<T> Bv = ldcg16(row + 2 * H + h * HD + D);                           // v
    float xl[8];                                                                     // LoRA'd q (lanes 0-15) / v (lanes 16-31)
#pragma unroll
    for (int e = 0; e < 8; ++e) xl[e] = Tr<T>::f(g2 ? Bv.v[e] : A.v[e]);
    if (lora_r > 0) {
      const T* tsrc = row + 3 * H + (g2 ? lora_r : 0);
      const T* brows = lora_b + ((int64_t)(g2 ? H : 0) + h * HD + D) * lora_r;
      float t[16];
      if (lora_r == 8) {                         // the adapter rank of the reference (finetune.py:167): 128-bit rows
        const Vec8<T> tt = ldcg16(tsrc);
        Vec8<T> br[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) br[e] = ld16(brows + e * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = Tr<T>::f(tt.v[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float sdot = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) sdot = fmaf(Tr<T>::f(br[e].v[i]), t[i], sdot);
          xl[e] = Tr<T>::rr(xl[e] + Tr<T>::rr(p.lora_scale * Tr<T>::rr(sdot)));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = i < lora_r ? Tr<T>::f(ldcg_t(tsrc + i)) : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float sdot = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < lora_r) sdot = fmaf(Tr<T>::f(brows[e * lora_r + i]), t[i], sdot);
          xl[e] = Tr<T>::rr(xl[e] + Tr<T>::rr(p.lora_scale * Tr<T>::rr(sdot)));
        }
      }
    }
    float q[8], kn[8], vn[8];
    {
      float rot[8];                              // RoPE: q (lanes 0-15) from the LoRA'd values, k (lanes 16-31) from the raw ones
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float own = g2 ? Tr<T>::f(A.v[e]) : xl[e];
        const float oth = __shfl_xor_sync(0xffffffffu, own, 8);
        const float c = Tr<T>::f(cv.v[e]), sn = Tr<T>::f(sv.v[e]);
        rot[e] = Tr<T>::rr(Tr<T>::rr(own * c) + Tr<T>::rr((lo_half ? -oth : oth) * sn));
      }
      if (pw == 0 && g2 == 1) {                  // append the new token's k (post-RoPE) and v to the cache
        Vec8<T> ko, vo;
#pragma unroll
        for (int e = 0; e < 8; ++e) { ko.v[e] = Tr<T>::r(rot[e]); vo.v[e] = Tr<T>::r(xl[e]); }
        const int64_t slot_off = (((int64_t)b * nh + h) * cmax + ctx) * HD + D;
        *reinterpret_cast<uint4*>(reinterpret_cast<T*>(L.kc) + slot_off) = *reinterpret_cast<const uint4*>(&ko);
        *reinterpret_cast<uint4*>(reinterpret_cast<T*>(L.vc) + slot_off) = *reinterpret_cast<const uint4*>(&vo);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        q[e] = __shfl_sync(0xffffffffu, rot[e], l16);
        kn[e] = __shfl_sync(0xffffffffu, rot[e], 16 + l16);
        vn[e] = __shfl_sync(0xffffffffu, xl[e], 16 + l16);
      }
    }
    const uint8_t* km = p.keymask + (int64_t)b * cmax;
    auto score_of = [&](float dot, unsigned keep) {      // modeling_llama_imgemb.py:216-230 (decode: padding mask only)
      float s = Tr<T>::rr(dot);
      s = Tr<T>::rr(s / sqrt_d);
      s = Tr<T>::rr(s + (keep ? 0.f : lowest));
      return fmaxf(s, lowest);
    };
    // padding-mask bytes: lane l16 holds the byte of key l16 of the current chunk, fetched one chunk ahead (a per-key
    // global load inside the key loop costs an L2 round trip per key: L1 is invalidated by every gpu-scope fence)
    auto km_load = [&](int s) -> unsigned {
      const int j = (pw + s * P) * ATT_CH + l16;
      return (s < n_my && j < ctx) ? (unsigned)km[j] : 0u;
    };
    const unsigned km_new = km[ctx];
    unsigned km_cur = km_load(0);
    const long long c_pro = clock64();
    // ---- scores of this warp's K chunks ------------------------------------------------------------------------------
    float lmax = -INFINITY;
    for (int s = 0; s < n_my; ++s) {
      const unsigned km_next = km_load(s + 1);
      const uint32_t qi = seq + (uint32_t)s;
      mbar_wait_warp(bar + (qi & 1u), (qi >> 1) & 1u, 40, lane);
      const T* buf = reinterpret_cast<const T*>(cbuf + (qi & 1u) * (ATT_CH * HD * 2));
      const int c = pw + s * P;
      int keys = ctx - c * ATT_CH;
      keys = keys > ATT_CH ? ATT_CH : keys;
      // the 8 key pairs of the chunk are independent: all dot products first, then the shuffle trees interleaved, then the
      // rounding chain of the scores - one dependent chain per key pair would run at instruction latency
      float dd[ATT_CH / 2];
#pragma unroll
      for (int i = 0; i < ATT_CH / 2; ++i) {
        const int jl = g2 + 2 * i;
        float d = 0.f;
        if (jl < keys) {
          const Vec8<T> kk = *reinterpret_cast<const Vec8<T>*>(buf + jl * HD + D);
#pragma unroll
          for (int e = 0; e < 8; ++e) d = fmaf(q[e], Tr<T>::f(kk.v[e]), d);
        }
        dd[i] = d;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < ATT_CH / 2; ++i) dd[i] += __shfl_xor_sync(hmask, dd[i], o);
      }
#pragma unroll
      for (int i = 0; i < ATT_CH / 2; ++i) {
        const int jl = g2 + 2 * i;
        const unsigned keep = __shfl_sync(hmask, km_cur, (lane & 16) + jl);
        if (jl < keys) {
          const float svv = score_of(dd[i], keep);
          if (l16 == 0) sc[c * ATT_CH + jl] = Tr<T>::r(svv);
          lmax = fmaxf(lmax, svv);
        }
      }
      km_cur = km_next;
      __syncwarp();
      if (lane == 0 && s + 2 < 2 * n_my) issue(s + 2);
    }
    if (pw == 0) {                               // the token just appended
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(q[e], kn[e], d);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(hmask, d, o);
      const float svv = score_of(d, km_new);
      if (lane == 0) sc[ctx] = Tr<T>::r(svv);
      lmax = fmaxf(lmax, svv);
    }
    const long long c_sc = clock64();
    lmax = warp_max(lmax);
    float mx = lmax;
    if (P > 1) {
      if (lane == 0) s_red[ww * 2] = lmax;
      team_bar(team, P * 32);
      mx = s_red[(team * P) * 2];
      for (int i = 1; i < P; ++i) mx = fmaxf(mx, s_red[(team * P + i) * 2]);
    } else {
      __syncwarp();
    }
    // ---- softmax denominator (fp32), modeling_llama_imgemb.py:233 ----------------------------------------------------
    float lsum = 0.f;
    for (int j = pw * 32 + lane; j <= ctx; j += P * 32) lsum += expf(Tr<T>::f(sc[j]) - mx);
    lsum = warp_sum(lsum);
    float sum = lsum;
    if (P > 1) {
      if (lane == 0) s_red[ww * 2 + 1] = lsum;
      team_bar(team, P * 32);
      sum = 0.f;
      for (int i = 0; i < P; ++i) sum += s_red[(team * P + i) * 2 + 1];
    }
    const long long c_sm = clock64();
    // ---- P.V over this warp's V chunks -------------------------------------------------------------------------------
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = n_my; s < 2 * n_my; ++s) {
      const uint32_t qi = seq + (uint32_t)s;
      mbar_wait_warp(bar + (qi & 1u), (qi >> 1) & 1u, 41, lane);
      const T* buf = reinterpret_cast<const T*>(cbuf + (qi & 1u) * (ATT_CH * HD * 2));
      const int c = pw + (s - n_my) * P;
      int keys = ctx - c * ATT_CH;
      keys = keys > ATT_CH ? ATT_CH : keys;
      float pj[ATT_CH / 2];
#pragma unroll
      for (int i = 0; i < ATT_CH / 2; ++i) {       // softmax(fp32).to(dtype), independent per key
        const int jl = g2 + 2 * i;
        pj[i] = jl < keys ? Tr<T>::rr(expf(Tr<T>::f(sc[c * ATT_CH + jl]) - mx) / sum) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < ATT_CH / 2; ++i) {
        const int jl = g2 + 2 * i;
        if (jl < keys) {
          const Vec8<T> vv = *reinterpret_cast<const Vec8<T>*>(buf + jl * HD + D);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj[i], Tr<T>::f(vv.v[e]), acc[e]);
        }
      }
      __syncwarp();
      if (lane == 0 && s + 2 < 2 * n_my) issue(s + 2);
    }
    if (pw == 0 && g2 == 0) {
      const float pj = Tr<T>::rr(expf(Tr<T>::f(sc[ctx]) - mx) / sum);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vn[e], acc[e]);
    }
    const long long c_pv = clock64();
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_down_sync(0xffffffffu, acc[e], 16);
    if (P > 1) {
      if (pw != 0 && lane < 16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) pacc[D + e] = acc[e];
      }
      team_bar(team, P * 32);
      if (pw == 0 && lane < 16) {
        for (int i = 1; i < P; ++i) {
          const float* pa = reinterpret_cast<const float*>(scratch + (size_t)(team * P + i) * ATT_WARP_BYTES);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += pa[D + e];
        }
      }
    }
    if (pw == 0 && lane < 16) {
      Vec8<T> o;
#pragma unroll
      for (int e = 0; e < 8; ++e) o.v[e] = Tr<T>::r(acc[e]);
      *reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.att) + (int64_t)b * H + h * HD + D) = *reinterpret_cast<const uint4*>(&o);
    }
    seq += (uint32_t)(2 * n_my);
    if (P > 1) team_bar(team, P * 32);           // partial accumulators / scores are free for the next item
    else __syncwarp();
    const long long c_end = clock64();
    tA += c_pro - c_start; tB += c_sc - c_pro; tC += c_sm - c_sc; tD += c_pv - c_sm; tE += c_end - c_pv;
  }
  if (p.trace != nullptr && lane == 0) {
    unsigned long long* o = p.trace + (size_t)p.G * (p.l1 - p.l0) * 80 + 320 + (size_t)p.G * 12 + ((size_t)cta * WORK_WARPS + ww) * 8;
    o[0] += (unsigned long long)tA; o[1] += (unsigned long long)tB; o[2] += (unsigned long long)tC; o[3] += (unsigned long long)tD;
    o[4] += (unsigned long long)tE;
  }
}

// ------------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(THREADS, 1)
decode_mega_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_att,
                   const __grid_constant__ CUtensorMap map_mid, const __grid_constant__ WeightMaps wmaps, const MegaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wring = smem;
  uint8_t* xring = smem + SW * W_SLOT;
  uint8_t* s_lnw = smem + SMEM_RING;                     // [LNW_KB][64] RMSNorm weights of this CTA's k-blocks (norm phases)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_RING + LNW_KB * BK * 2);
  uint64_t* w_full = bars;
  uint64_t* cons = w_full + SW;
  uint64_t* x_full = cons + CONS_R;
  uint64_t* xn_full = x_full + SX;
  uint64_t* acc_full = xn_full + SX;
  uint64_t* acc_empty = acc_full + NACC;
  uint64_t* att_bar = acc_empty + NACC;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(att_bar + 2 * WORK_WARPS);
  uint32_t* s_flag = tmem_ptr_smem + 1;
  float* s_rstd = reinterpret_cast<float*>(tmem_ptr_smem + 4);
  float* s_ssq = s_rstd + 32;            // [8][32]
  float* s_red = s_ssq + 8 * 32;         // [WORK_WARPS][2]
  CtaSched* sched = reinterpret_cast<CtaSched*>(s_red + WORK_WARPS * 2 + 2);      // this CTA's slice of the work table
  uint32_t* s_wkb = reinterpret_cast<uint32_t*>(sched + 1);                       // [SW] k-block that consumes the tile in each W slot

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int KB_H = p.H / BK, KB_I = p.I / BK;

  pdl_launch_dependents();
  if (threadIdx.x >= 128 && threadIdx.x < 128 + 4 * MAX_SEG) {
    const int g = (threadIdx.x - 128) / MAX_SEG, i = (threadIdx.x - 128) % MAX_SEG;
    sched->seg[g][i] = p.sched->seg[g][cta][i];
    if (i == 0) sched->nseg[g] = p.sched->nseg[g][cta];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < SW; ++i) mbar_init(&w_full[i], 1);
    for (int i = 0; i < CONS_R; ++i) mbar_init(&cons[i], 1);
    for (int i = 0; i < SX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&xn_full[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], WORKERS); }
    for (int i = 0; i < 2 * WORK_WARPS; ++i) mbar_init(&att_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // The three control warps run their loops warp-uniformly (all 32 lanes, uniform registers) and elect one lane only for
  // the asynchronous instruction itself: a lane-0-only loop makes ptxas wrap every UTCHMMA / UTMALDG operand in a
  // waterfall (ELECT / R2UR.BROADCAST / BRA) and the serial instruction stream of the MMA loop, not HBM, set the pace
  // (measured ~2000 cycles per k-block against a budget of 700).
  if (warp == 0) {
    // ===================== weight producer: never waits on anything but a free ring slot =====================
    // (Tried and dropped, both measured with tools/trace_mega.py: a second cursor pulling tiles beyond the ring into L2 with
    // cp.async.bulk.prefetch.tensor - either continuously or only while the ring is stalled in a phase boundary.  Neither
    // shortened the streams and both lengthened every phase (extra traffic competing with the latency-critical fix-up /
    // barrier round trips), so the ring is the only prefetch.)
    uint32_t slot = 0, ki = 0, filled = 0;
    for (int l = p.l0; l < p.l1; ++l) {
      for (int g = 0; g < 4; ++g) {
        const CUtensorMap* map = &wmaps.m[l * 4 + g];
        const int ns = sched->nseg[g];
        const int reps = g == G_GU ? 2 : 1;
        bool first = true;
        for (int s = 0; s < ns; ++s) {
          const Seg sg = sched->seg[g][s];
          const int row = sg.tile * TILE_N;
          for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
            // the one (gate|up: two) slots of this k-block; their previous tiles must have been read by their MMAs
            uint32_t sl[2] = {0, 0};
            for (int r = 0; r < reps; ++r) {
              sl[r] = slot;
              if (filled >= SW) {
                const uint32_t old = s_wkb[slot];
                mbar_wait(&cons[old & (CONS_R - 1)], (old / CONS_R) & 1u, 1);
              } else {
                ++filled;
              }
              if (++slot == SW) slot = 0;
            }
            __syncwarp();
            if (elect_one()) {
              if (first) trace_stamp(p, l, phase_of_g(g), 7);
              for (int r = 0; r < reps; ++r) {
                s_wkb[sl[r]] = ki;
                if (p.dbg & 2) {
                  mbar_arrive(&w_full[sl[r]]);      // diagnostic: no weight traffic
                } else {
                  mbar_expect_tx(&w_full[sl[r]], W_SLOT);
                  tma_load_2d(wring + sl[r] * W_SLOT, map, &w_full[sl[r]], kb * BK, r == 0 ? row : p.I + row, HINT_EVICT_FIRST);
                }
              }
            }
            __syncwarp();
            first = false;
            ++ki;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(Tr<T>::umma_fmt, TILE_N, NT);
    const uint64_t dw0 = make_smem_desc(smem_u32(wring)), dx0 = make_smem_desc(smem_u32(xring));
    uint32_t ws = 0, wph = 0, xs = 0, xph = 0, sc = 0, xn_par = 0, ci = 0;
    for (int l = p.l0; l < p.l1; ++l) {
      for (int g = 0; g < 4; ++g) {
        const bool norm = (g == G_QKV || g == G_GU), gu = (g == G_GU);
        const int ns = sched->nseg[g];
        for (int s = 0; s < ns; ++s) {
          const Seg sg = sched->seg[g][s];
          const uint32_t a = sc % NACC;
          mbar_wait(&acc_empty[a], ((sc / NACC) & 1u) ^ 1u, 2);
          tc_fence_after();
          const uint32_t d0 = tmem_base + a * ACC_COLS;
          for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
            const uint32_t ws0 = ws;
            mbar_wait(&w_full[ws0], wph, 3);
            if (++ws == SW) { ws = 0; wph ^= 1u; }
            uint32_t ws1 = ws0;
            if (gu) {
              ws1 = ws;
              mbar_wait(&w_full[ws1], wph, 4);
              if (++ws == SW) { ws = 0; wph ^= 1u; }
            }
            if (norm) {
              mbar_wait(&xn_full[xs], (xn_par >> xs) & 1u, 5);
              xn_par ^= 1u << xs;
            } else {
              mbar_wait(&x_full[xs], xph, 6);
            }
            tc_fence_after();
            const uint64_t da = dw0 + (uint64_t)(ws0 * (W_SLOT >> 4)), du = dw0 + (uint64_t)(ws1 * (W_SLOT >> 4));
            const uint64_t db = dx0 + (uint64_t)(xs * (X_SLOT >> 4));
            const uint32_t accf = kb > sg.kb0 ? 1u : 0u;
            __syncwarp();
            if (elect_one()) {
              if (s == 0 && kb == sg.kb0) trace_stamp(p, l, phase_of_g(g), 4);
              if (!(p.dbg & 4)) {
#pragma unroll
                for (int k = 0; k < BK / UK; ++k) {
                  const uint64_t koff = (uint64_t)((k * UK * 2) >> 4);
                  if (gu) {          // chains: gate k even/odd -> columns 0 / 32, up -> 64 / 96
                    const uint32_t af = (accf || k >= 2) ? 1u : 0u;
                    tc_mma_f16(d0 + (k & 1) * NT, da + koff, db + koff, idesc, af);
                    tc_mma_f16(d0 + (2 + (k & 1)) * NT, du + koff, db + koff, idesc, af);
                  } else {           // chain k -> columns 32 k
                    tc_mma_f16(d0 + k * NT, da + koff, db + koff, idesc, accf);
                  }
                }
              }
              tc_commit(&cons[ci]);               // frees this k-block's W and X slots once the MMAs have read them
              if (kb == sg.kb1 - 1) {
                tc_commit(&acc_full[a]);
                if (s == ns - 1) trace_stamp(p, l, phase_of_g(g), 5);
              }
            }
            __syncwarp();
            ci = (ci + 1) & (CONS_R - 1);
            if (++xs == SX) { xs = 0; xph ^= 1u; }
          }
          ++sc;
        }
      }
    }
  } else if (warp == 2) {
    // ===================== activation producer =====================
    pdl_wait();
    uint32_t xi = 0, slot = 0;
    for (int l = p.l0; l < p.l1; ++l) {
      for (int g = 0; g < 4; ++g) {
        const int ns = sched->nseg[g];
        if (ns == 0) continue;
        if (lane == 0) grid_wait(p.sync, (uint32_t)phase_epoch(l - p.l0, g) * (uint32_t)p.G, 10 + g);
        __syncwarp();
        fence_proxy_async_all();
        if (lane == 0) trace_stamp(p, l, phase_of_g(g), 6);
        const CUtensorMap* map = (g == G_QKV || g == G_GU) ? &map_x : (g == G_O ? &map_att : &map_mid);
        for (int s = 0; s < ns; ++s) {
          const Seg sg = sched->seg[g][s];
          for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
            if (xi >= SX) {
              const uint32_t old = xi - SX;
              mbar_wait(&cons[old & (CONS_R - 1)], (old / CONS_R) & 1u, 7);
            }
            __syncwarp();
            if (elect_one()) {
              if (p.dbg & 1) {
                mbar_arrive(&x_full[slot]);       // diagnostic: no activation traffic
              } else {
                mbar_expect_tx(&x_full[slot], X_SLOT);
                tma_load_2d(xring + slot * X_SLOT, map, &x_full[slot], kb * BK, 0, HINT_EVICT_LAST);
              }
            }
            __syncwarp();
            ++xi;
            if (++slot == SX) slot = 0;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== workers =====================
    pdl_wait();
    const int ww = warp - 4, wtid = threadIdx.x - 128;
    const int quad = warp & 3;
    const int col0 = ww >= 4 ? 16 : 0;
    const int n_local = quad * 32 + lane;
    const int B = p.B, H = p.H;
    T* xg = reinterpret_cast<T*>(p.x);
    uint32_t xi = 0, sc = 0, att_seq = 0, epoch = 0;
    uint32_t* tile_ctr = p.sync + 32;

    for (int l = p.l0; l < p.l1; ++l) {
      const LayerDev L = p.lay[l];
      for (int ph = 0; ph < 5; ++ph) {
        if (ph == 0 || ph == 3) {
          // RMSNorm weights of this CTA's k-blocks -> shared memory (constants: fetched while the grid barrier is pending)
          const int g = ph == 0 ? G_QKV : G_GU;
          const T* lnw = reinterpret_cast<const T*>(g == G_QKV ? L.ln1 : L.ln2);
          const int ns = sched->nseg[g];
          const int len0 = ns > 0 ? sched->seg[g][0].kb1 - sched->seg[g][0].kb0 : 0;
          int n_x = 0;
          for (int s = 0; s < ns; ++s) n_x += sched->seg[g][s].kb1 - sched->seg[g][s].kb0;
          n_x = n_x > LNW_KB ? LNW_KB : n_x;
          for (int idx = wtid; idx < n_x * 8; idx += WORKERS) {
            const int i = idx >> 3, c = idx & 7;
            const int kb = i < len0 ? sched->seg[g][0].kb0 + i : sched->seg[g][1].kb0 + (i - len0);
            *reinterpret_cast<uint4*>(s_lnw + (size_t)idx * 16) = *reinterpret_cast<const uint4*>(lnw + kb * BK + c * 8);
          }
        }
        if (ph == 1) {
          // cached K/V rows of this warp's (sequence, head) items -> L2 while the QKV phase drains: the rows were written by
          // earlier steps, and with them L2-resident the chunk copies of the attention sweep are L2 hits, not HBM round trips
          const int P = p.att_P, team = ww / P, teams = WORK_WARPS / P;
          if (ww % P == 0) {
            const int ctx = p.ctx_len[0];
            const int per = (ctx + 31) / 32, r0 = lane * per;
            int nr = ctx - r0;
            nr = nr > per ? per : nr;
            for (int item = team * p.G + cta; item < B * p.nh; item += teams * p.G) {
              if (nr > 0) {
                const size_t off = ((size_t)item * p.cmax + r0) * 128 * sizeof(T);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(L.kc) + off), "r"(nr * 256) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(L.vc) + off), "r"(nr * 256) : "memory");
              }
            }
          }
        }
        // wait for the previous phase of the whole grid
        if (wtid == 0) grid_wait(p.sync, epoch * (uint32_t)p.G, 20 + ph);
        worker_bar();
        if (wtid == 0) trace_stamp(p, l, ph, 0);
        if (ph == 1) {
          attention_items<T>(p, L, cta, ww, lane, xring, att_bar, s_red, att_seq);
        } else {
          const int g = ph == 0 ? G_QKV : ph - 1;
          const bool norm = (g == G_QKV || g == G_GU), gu = (g == G_GU);
          const int ns = sched->nseg[g];
          int n_x = 0;
          for (int s = 0; s < ns; ++s) n_x += sched->seg[g][s].kb1 - sched->seg[g][s].kb0;
          if (norm && ns > 0) {
            // ---- RMSNorm statistics: rstd[j] = rsqrt(mean(x_j^2) + eps), fp32 (modeling_llama_imgemb.py:85-93) ----
            if (g == G_QKV && l == p.l0) {
              for (int j = ww; j < B; j += WORK_WARPS) {
                float ss = 0.f;
                for (int k = lane * 8; k < H; k += 256) {
                  const Vec8<T> v = ldcg16(xg + (int64_t)j * H + k);
#pragma unroll
                  for (int e = 0; e < 8; ++e) { const float f = Tr<T>::f(v.v[e]); ss = fmaf(f, f, ss); }
                }
                ss = warp_sum(ss);
                if (lane == 0) s_rstd[j] = 1.0f / sqrtf(ss / (float)H + p.eps);
              }
            } else {
              // fixed-order sum of the per-tile partials the o_proj / down_proj epilogues left in L2
              const int j = wtid & 31, part = wtid >> 5, nt = H / TILE_N;
              float v = 0.f;
#pragma unroll 4
              for (int t = part; t < nt; t += 8) v += __ldcg(p.ssq + t * 32 + j);
              s_ssq[part * 32 + j] = v;
              worker_bar();
              if (wtid < B) {
                float tot = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) tot += s_ssq[q * 32 + wtid];
                s_rstd[wtid] = 1.0f / sqrtf(tot / (float)H + p.eps);
              }
            }
            worker_bar();
            // ---- normalise the token tiles in place: xn = T(w * T(x * rstd)) ----
            const T* lnw = reinterpret_cast<const T*>(g == G_QKV ? L.ln1 : L.ln2);
            // Each warp owns every 8th tile and transforms it alone (wait / 8 rounds of 32 chunks / fence / arrive), so the
            // eight latency chains run in parallel instead of all 256 threads serialising on every tile.
            int it = 0;
            for (int s = 0; s < ns; ++s) {
              const Seg sg = sched->seg[g][s];
              for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++it) {
                const uint32_t slot = xi % SX, par = (xi / SX) & 1u;
                ++xi;
                if ((it & (WORK_WARPS - 1)) != ww) continue;
                mbar_wait_warp(&x_full[slot], par, 30, lane);
                const int pc = lane & 7;
#pragma unroll
                for (int i = 0; i < NT / 4; ++i) {
                  const int r = (lane >> 3) + 4 * i;
                  if (r >= B) break;
                  const int c = pc ^ (r & 7);
                  const float rs = s_rstd[r];
                  uint4* cp = reinterpret_cast<uint4*>(xring + slot * X_SLOT + r * 128 + pc * 16);
                  uint4 raw = *cp;
                  const Vec8<T> xv = *reinterpret_cast<const Vec8<T>*>(&raw);
                  const Vec8<T> wv = it < LNW_KB ? *reinterpret_cast<const Vec8<T>*>(s_lnw + (size_t)(it * 8 + c) * 16) : ld16(lnw + kb * BK + c * 8);
                  Vec8<T> o;
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float y = Tr<T>::rr(Tr<T>::f(xv.v[e]) * rs);
                    o.v[e] = Tr<T>::r(Tr<T>::f(wv.v[e]) * y);
                  }
                  *cp = *reinterpret_cast<const uint4*>(&o);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&xn_full[slot]);
              }
            }
          } else {
            xi += (uint32_t)n_x;
          }
          if (wtid == 0) trace_stamp(p, l, ph, 1);
          // ---- epilogues of this CTA's segments ----
          for (int s = 0; s < ns; ++s) {
            const Seg sg = sched->seg[g][s];
            const uint32_t a = sc % NACC;
            if (wtid == 0) mbar_wait(&acc_full[a], (sc / NACC) & 1u, 31);
            worker_bar();
            if (wtid == 0 && s == 0) trace_stamp(p, l, ph, 2);
            ++sc;
            tc_fence_after();
            const uint32_t taddr = tmem_base + a * ACC_COLS + col0 + ((uint32_t)(quad * 32) << 16);
            float acc[16], accu[16];
            {
              uint32_t r0[16], r1[16];
              tc_ld16(taddr, r0);
              tc_ld16(taddr + NT, r1);
              tc_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
              tc_ld16(taddr + 2 * NT, r0);
              tc_ld16(taddr + 3 * NT, r1);
              tc_wait_ld();
              if (gu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) accu[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { acc[j] = (acc[j] + __uint_as_float(r0[j])) + __uint_as_float(r1[j]); accu[j] = 0.f; }
              }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[a]);
            // residual operands of the tile this CTA will finish: requested before the fix-up waits, consumed after them
            float res[16];
            if ((g == G_O || g == G_DN) && sg.split == 0) {
              const int n = sg.tile * TILE_N + n_local;
#pragma unroll
              for (int j = 0; j < 16; ++j) res[j] = (col0 + j < B && n < H) ? Tr<T>::f(ldcg_t(xg + (int64_t)(col0 + j) * H + n)) : 0.f;
            }
            // Stream-K fix-up with a FIXED finaliser: the CTA that holds the head of the tile's k range (split 0; it is that
            // CTA's last segment, so it finishes when the phase ends) sums the other contributors' fp32 partials from L2 onto
            // its own accumulator, in split order (deterministic).  Contributors only publish and move on.
            bool finalise = true;
            if (sg.nsplits > 1 && sg.split != 0) {
              float* part = p.ws + ((int64_t)sg.tile * MAX_SPLIT + sg.split) * PART_STRIDE;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (col0 + j < B) {
                  __stcg(part + (col0 + j) * TILE_N + n_local, acc[j]);
                  if (gu) __stcg(part + (NT + col0 + j) * TILE_N + n_local, accu[j]);
                }
              }
              __threadfence();
              worker_bar();
              if (wtid == 0) atomicAdd(tile_ctr + sg.tile, 1u);
              finalise = false;
            } else if (sg.nsplits > 1) {
              if (wtid == 0) {
                grid_wait(tile_ctr + sg.tile, (uint32_t)sg.nsplits - 1, 50);
                tile_ctr[sg.tile] = 0;                    // re-arm for the next phase that uses this tile index
              }
              worker_bar();
              if (wtid == 0 && s == ns - 1) trace_stamp(p, l, ph, 8);
              __threadfence();
              const float* ps0 = p.ws + (int64_t)sg.tile * MAX_SPLIT * PART_STRIDE + n_local;
              if (gu) {
#pragma unroll 1
                for (int sp = 1; sp < sg.nsplits; sp += 2) {     // two contributors (64 loads) in flight per round
                  const bool two = sp + 1 < sg.nsplits;
                  const float* pa = ps0 + (int64_t)sp * PART_STRIDE;
                  const float* pb = pa + PART_STRIDE;
                  float v0[16], v1[16], u0[16], u1[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const bool ok = col0 + j < B;
                    v0[j] = ok ? __ldcg(pa + (col0 + j) * TILE_N) : 0.f;
                    v1[j] = (ok && two) ? __ldcg(pb + (col0 + j) * TILE_N) : 0.f;
                    u0[j] = ok ? __ldcg(pa + (NT + col0 + j) * TILE_N) : 0.f;
                    u1[j] = (ok && two) ? __ldcg(pb + (NT + col0 + j) * TILE_N) : 0.f;
                  }
#pragma unroll
                  for (int j = 0; j < 16; ++j) { acc[j] = (acc[j] + v0[j]) + v1[j]; accu[j] = (accu[j] + u0[j]) + u1[j]; }
                }
              } else {
#pragma unroll 1
                for (int sp = 1; sp < sg.nsplits; sp += 4) {     // four contributors (64 loads) in flight per round
                  float v[4][16];
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float* pq = ps0 + (int64_t)(sp + q) * PART_STRIDE;
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[q][j] = (sp + q < sg.nsplits && col0 + j < B) ? __ldcg(pq + (col0 + j) * TILE_N) : 0.f;
                  }
#pragma unroll
                  for (int j = 0; j < 16; ++j) acc[j] = (((acc[j] + v[0][j]) + v[1][j]) + v[2][j]) + v[3][j];
                }
              }
            }
            if (wtid == 0 && s == ns - 1) trace_stamp(p, l, ph, 9);
            if (finalise) {
              const int n = sg.tile * TILE_N + n_local;
              if (g == G_QKV) {
                if (n < p.n_qkv) {
                  T* o = reinterpret_cast<T*>(p.qkv) + n;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (col0 + j < B) o[(int64_t)(col0 + j) * p.ldq] = Tr<T>::r(acc[j]);
                }
              } else if (g == G_GU) {
                if (n < p.I) {
                  T* o = reinterpret_cast<T*>(p.mid) + n;
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    if (col0 + j < B) {
                      const float gg = Tr<T>::rr(acc[j]), uu = Tr<T>::rr(accu[j]);
                      o[(int64_t)(col0 + j) * p.I] = Tr<T>::r(Tr<T>::rr(silu_f(gg)) * uu);     // T(T(silu(T(g))) * T(u))
                    }
                  }
                }
              } else {
                // o_proj / down_proj: residual add in the storage dtype + sum of squares of the new residual stream
                float yy[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  yy[j] = 0.f;
                  if (col0 + j < B && n < H) {
                    const T y = Tr<T>::r(res[j] + Tr<T>::rr(acc[j]));
                    xg[(int64_t)(col0 + j) * H + n] = y;
                    const float f = Tr<T>::f(y);
                    yy[j] = f * f;
                  }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float t = warp_sum(yy[j]);
                  if (lane == 0) s_ssq[quad * 32 + col0 + j] = t;
                }
                worker_bar();
                if (wtid < B) __stcg(p.ssq + sg.tile * 32 + wtid, s_ssq[wtid] + s_ssq[32 + wtid] + s_ssq[64 + wtid] + s_ssq[96 + wtid]);
                worker_bar();
              }
            }
          }
        }
        // arrive at the grid barrier that closes this phase
        worker_bar();
        if (wtid == 0) trace_stamp(p, l, ph, 3);
        if (wtid == 0) {
          __threadfence();
          atomicAdd(p.sync, 1u);
          trace_stamp(p, l, ph, 10);
        }
        ++epoch;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  if (threadIdx.x == 0) {
    // the last CTA to leave re-arms the counters for the next launch (nobody polls them any more)
    const uint32_t prev = atomicAdd(p.sync + 1, 1u);
    if (prev == (uint32_t)p.G - 1) {
      p.sync[0] = 0;
      p.sync[1] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
int make_map(CUtensorMap* map, const void* ptr, int64_t ld, int rows, int K, int box_rows, int dtype) {
  PFN_encodeTiled enc = get_encode();
  RD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == RD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                   gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RD_REQUIRE(r == CUDA_SUCCESS, "decode_mega: cuTensorMapEncodeTiled failed (%d) ptr=%p ld=%lld rows=%d K=%d", (int)r, ptr, (long long)ld, rows, K);
  return RD_OK;
}

// Deal the (tile, k-block) units of one GEMM phase to `G` CTAs in contiguous, equal runs.  Returns false if the split
// would need more than MAX_SEG segments per CTA or MAX_SPLIT contributors per tile.
bool build_phase(Sched* sc, int g, int tiles, int kb, int G, int G_all) {
  for (int c = 0; c < G_all; ++c) sc->nseg[g][c] = 0;
  const long long U = (long long)tiles * kb;
  std::vector<int> per_tile(tiles, 0);
  for (int c = 0; c < G; ++c) {
    long long u0 = U * c / G, u1 = U * (c + 1) / G;
    while (u0 < u1) {
      const int t = (int)(u0 / kb), k0 = (int)(u0 % kb);
      const int k1 = (int)((u1 - u0) < (kb - k0) ? k0 + (u1 - u0) : kb);
      int& n = sc->nseg[g][c];
      if (n >= MAX_SEG) return false;
      Seg& s = sc->seg[g][c][n++];
      s.tile = t; s.kb0 = k0; s.kb1 = k1; s.split = per_tile[t]++; s.nsplits = 0;
      u0 += k1 - k0;
    }
  }
  for (int t = 0; t < tiles; ++t) if (per_tile[t] > MAX_SPLIT) return false;
  for (int c = 0; c < G; ++c)
    for (int i = 0; i < sc->nseg[g][c]; ++i) sc->seg[g][c][i].nsplits = per_tile[sc->seg[g][c][i].tile];
  return true;
}

}  // namespace

static unsigned long long* g_mega_trace = nullptr;
// development aid (tools/trace_mega.py): device buffer of [ctas][layers][5][8] u64 globaltimer stamps, nullptr = off
extern "C" int rd_mega_set_trace(void* buf) { g_mega_trace = (unsigned long long*)buf; return RD_OK; }
extern "C" int rd_mega_ctas(rd_mega* m);

struct rd_mega {
  MegaCreate c;
  int G = 0, n_qkv = 0;
  WeightMaps* wmaps_host = nullptr;
  LayerDev* lay = nullptr;
  Sched* sched = nullptr;
  float *ws = nullptr, *ssq = nullptr;
  uint32_t* sync = nullptr;
};

const char* rd_mega_unsupported_reason(const MegaCreate* c) {
  if (!c) return "null config";
  if (c->H % 128 != 0) return "hidden size is not a multiple of 128";
  if (c->I % 64 != 0) return "intermediate size is not a multiple of 64";
  if (c->H / c->nh != 128) return "head_dim != 128";
  if (c->lora_r > 16) return "lora_r > 16";
  if (c->cmax > ATT_MAX_CTX + 1) return "max_ctx > 1024";
  if (c->H / 128 > 480) return "hidden size too large";
  if (c->layers > MAX_LAYERS) return "too many layers for the tensor-map parameter block";
  return nullptr;
}

int rd_mega_create(const MegaCreate* c, const MegaLayerDesc* layers, rd_mega** out) {
  RD_REQUIRE(c && layers && out, "rd_mega_create: null argument");
  const char* why = rd_mega_unsupported_reason(c);
  RD_REQUIRE(why == nullptr, "rd_mega_create: %s", why);
  int dev = 0, sms = 0;
  RD_CHECK_CUDA(cudaGetDevice(&dev));
  RD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  rd_mega* m = new rd_mega();
  m->c = *c;
  m->G = sms < MAX_G ? sms : MAX_G;
  if (const char* e = getenv("RD_MEGA_CTAS")) { const int g = atoi(e); if (g > 0 && g < m->G) m->G = g; }
  m->n_qkv = 3 * c->H + 2 * c->lora_r;
  const int H = c->H, I = c->I;
  // schedule
  std::vector<Sched> hs(1);
  Sched* sc = &hs[0];
  memset(sc, 0, sizeof(Sched));
  const int tiles[4] = {(m->n_qkv + TILE_N - 1) / TILE_N, H / TILE_N, (I + TILE_N - 1) / TILE_N, H / TILE_N};
  const int kbs[4] = {H / BK, H / BK, H / BK, I / BK};
  int max_tiles = 0;
  for (int g = 0; g < 4; ++g) {
    max_tiles = tiles[g] > max_tiles ? tiles[g] : max_tiles;
    const long long U = (long long)tiles[g] * kbs[g];
    int Gp0 = (int)(U / 8 > 0 ? U / 8 : 1);         // at least ~8 k-blocks per active CTA
    Gp0 = Gp0 > m->G ? m->G : Gp0;
    bool ok = false;
    for (int Gp = Gp0; Gp <= m->G && !ok; ++Gp) ok = build_phase(sc, g, tiles[g], kbs[g], Gp, m->G);
    for (int Gp = Gp0 - 1; Gp >= 1 && !ok; --Gp) ok = build_phase(sc, g, tiles[g], kbs[g], Gp, m->G);
    if (!ok) {
      delete m;
      rd_set_error("rd_mega_create: cannot schedule GEMM phase %d (%d tiles x %d k-blocks)", g, tiles[g], kbs[g]);
      return RD_ERR_UNSUPPORTED;
    }
  }
  // tensor maps of the weights
  m->wmaps_host = new WeightMaps();
  memset(m->wmaps_host, 0, sizeof(WeightMaps));
  CUtensorMap* hm = m->wmaps_host->m;
  std::vector<LayerDev> hl(c->layers);
  for (int l = 0; l < c->layers; ++l) {
    const MegaLayerDesc& d = layers[l];
    int r = make_map(&hm[l * 4 + G_QKV], d.qkv, H, m->n_qkv, H, TILE_N, c->dtype);
    if (r == RD_OK) r = make_map(&hm[l * 4 + G_O], d.o, H, H, H, TILE_N, c->dtype);
    if (r == RD_OK) r = make_map(&hm[l * 4 + G_GU], d.gate_up, H, 2 * I, H, TILE_N, c->dtype);
    if (r == RD_OK) r = make_map(&hm[l * 4 + G_DN], d.down, I, H, I, TILE_N, c->dtype);
    if (r != RD_OK) { rd_mega_destroy(m); return r; }
    hl[l].ln1 = d.ln1; hl[l].ln2 = d.ln2; hl[l].lora_b = d.lora_b; hl[l].kc = d.kc; hl[l].vc = d.vc;
  }
  auto fail = [&](cudaError_t e) { rd_set_error("rd_mega_create: CUDA error %s", cudaGetErrorString(e)); rd_mega_destroy(m); return RD_ERR_CUDA; };
  cudaError_t e;
  if ((e = cudaMalloc((void**)&m->lay, hl.size() * sizeof(LayerDev))) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpy(m->lay, hl.data(), hl.size() * sizeof(LayerDev), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&m->sched, sizeof(Sched))) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpy(m->sched, sc, sizeof(Sched), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
  const size_t ws_bytes = (size_t)max_tiles * MAX_SPLIT * PART_STRIDE * 4;
  if ((e = cudaMalloc((void**)&m->ws, ws_bytes)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&m->ssq, (size_t)(H / TILE_N) * 32 * 4)) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(m->ssq, 0, (size_t)(H / TILE_N) * 32 * 4)) != cudaSuccess) return fail(e);
  const size_t sync_bytes = (size_t)(32 + max_tiles + 32) * 4;
  if ((e = cudaMalloc((void**)&m->sync, sync_bytes)) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(m->sync, 0, sync_bytes)) != cudaSuccess) return fail(e);
  *out = m;
  return RD_OK;
}

extern "C" int rd_mega_ctas(rd_mega* m) { return m ? m->G : 0; }

void rd_mega_destroy(rd_mega* m) {
  if (!m) return;
  void* ptrs[] = {m->lay, m->sched, m->ws, m->ssq, m->sync};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete m->wmaps_host;
  delete m;
}

template <class T>
static int launch_mega(rd_mega* m, const MegaStep* s, cudaStream_t st) {
  const MegaCreate& c = m->c;
  RD_SMEM_ATTR_ONCE(SMEM_BYTES, decode_mega_kernel<T>);
  CUtensorMap map_x, map_att, map_mid;
  RD_CHECK(make_map(&map_x, s->x, c.H, s->B, c.H, NT, c.dtype));
  RD_CHECK(make_map(&map_att, s->att, c.H, s->B, c.H, NT, c.dtype));
  RD_CHECK(make_map(&map_mid, s->mid, c.I, s->B, c.I, NT, c.dtype));
  MegaParams p{};
  p.lay = m->lay; p.sched = m->sched;
  p.x = s->x; p.qkv = s->qkv; p.att = s->att; p.mid = s->mid;
  p.ws = m->ws; p.ssq = m->ssq; p.sync = m->sync;
  p.keymask = s->keymask; p.ctx_len = s->ctx_len; p.pos = s->pos; p.cos_t = s->cos; p.sin_t = s->sin;
  p.B = s->B; p.H = c.H; p.I = c.I; p.nh = c.nh; p.n_qkv = m->n_qkv; p.ldq = m->n_qkv; p.cmax = c.cmax; p.lora_r = c.lora_r;
  p.l0 = s->layer_begin; p.l1 = s->layer_end; p.G = m->G;
  p.lora_scale = c.lora_scale; p.eps = c.eps;
  p.trace = g_mega_trace;
  p.sw = SW;
  p.cb = 1;
  if (const char* e = getenv("RD_MEGA_CB")) { const int v = atoi(e); if (v >= 1 && v <= 2) p.cb = v; }
  if (const char* e = getenv("RD_MEGA_DBG")) p.dbg = atoi(e);
  if (const char* e = getenv("RD_MEGA_SW")) { const int v = atoi(e); if (v >= 2 && v <= SW) p.sw = v; }
  // warps per (sequence, head): as many as keeps every item in one round
  const int items = s->B * c.nh;
  int P = 8;
  while (P > 1 && (WORK_WARPS / P) * m->G < items) P >>= 1;
  p.att_P = P;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m->G); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = rd_pdl_enabled() ? 1 : 0;
  RD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_mega_kernel<T>, map_x, map_att, map_mid, *m->wmaps_host, p));
  return RD_OK;
}

int rd_mega_launch(rd_mega* m, const MegaStep* s, cudaStream_t st) {
  RD_REQUIRE(m && s, "rd_mega_launch: null argument");
  RD_REQUIRE(s->B > 0 && s->B <= NT && s->B <= m->c.max_batch, "rd_mega_launch: B=%d out of range (1..%d)", s->B, NT);
  RD_REQUIRE(s->layer_begin >= 0 && s->layer_begin < s->layer_end && s->layer_end <= m->c.layers, "rd_mega_launch: bad layer range");
  RD_DISPATCH_DTYPE(m->c.dtype, T, { return launch_mega<T>(m, s, st); });
}
