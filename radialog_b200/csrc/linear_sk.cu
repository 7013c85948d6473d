// Stream-K decode GEMM (sm_100a):  out[M,N] = epilogue( x[M,K] . W[N,K]^T ),  M <= 32 tokens.
//
// The decode GEMMs of a Vicuna-7B step are HBM-bound and have few weight tiles (97 / 32 / 86 / 32 / 251 tiles of 128 rows
// against 148 SMs).  linear_tc.cu covers them with tile x split-K grids (192-258 CTAs), which leaves some SMs with two
// CTAs' worth of bytes and others with one.  Here exactly one CTA per SM takes an EQUAL contiguous run of the
// (weight tile, k-block) units of the GEMM (host-built table), so every SM pulls the same number of weight bytes; a tile whose
// k range is shared by several CTAs is finished by the CTA that owns the head of the range, which adds the other
// contributors' fp32 partials from L2 onto its own accumulator in split order (deterministic).  Optionally the RMSNorm in
// front of the GEMM is applied to the token tiles in shared memory between the TMA and the MMA (no norm kernel, no xn
// buffer), with the row statistics taken from sum-of-squares partials that the producing GEMM's epilogue writes.
// Machinery (rings, consumed-barrier ring, warp-uniform control loops, fixed finaliser) as in decode_mega.cu, where each
// piece was measured; this is the one-GEMM-per-launch form that keeps programmatic dependent launch between kernels:
// <= 110 KB of shared memory per CTA, so the next kernel's CTAs are co-resident and prefetch their weights under this
// kernel's tail.
//
//   warp 0   : TMA producer (weights: evict-first, issued before griddepcontrol.wait; token tiles: evict-last)
//   warp 1   : tcgen05.mma issuer (UMMA 128 x 32, fp32 accumulators in TMEM, two accumulator buffers)
//   warps 2-5: RMSNorm of the token tiles (optional), epilogue (tcgen05.ld, stream-K fix-up, SwiGLU / residual, sum of squares)
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <vector>
#include "common.cuh"
#include "linear_sk.h"
#include "tc_ptx.cuh"

bool rd_pdl_enabled();

namespace {
using namespace tcptx;

constexpr int TILE_N = 128, BK = 64, NT = 32, UK = 16;
constexpr int W_SLOT = TILE_N * BK * 2;     // 16 KB
constexpr int X_SLOT = NT * BK * 2;         // 4 KB
constexpr int SW = 5;                       // weight ring (80 KB)
constexpr int SX = 4;                       // token-tile ring (16 KB)
constexpr int CONS_R = 16;                  // ring of "k-block consumed" barriers
constexpr int NACC = 2, ACC_COLS = 64, TMEM_COLS = NACC * ACC_COLS;
constexpr int EPI_WARPS = 4, EPI_THREADS = 128, THREADS = 64 + EPI_THREADS;
constexpr int MAX_G = 320, MAX_SEG = 4, MAX_SPLIT = 8;
constexpr int PART_STRIDE = 2 * NT * TILE_N;
constexpr int LNW_KB = 48;                  // k-blocks of RMSNorm weights staged per CTA (6 KB)
static_assert(CONS_R > SW && CONS_R > SX, "consumed-barrier ring must be longer than both operand rings");

struct Seg { int tile, kb0, kb1, split, nsplits, pad0, pad1, pad2; };
struct Sched {
  int nseg[MAX_G];
  Seg seg[MAX_G][MAX_SEG];
};
struct CtaSched { int nseg; int pad[3]; Seg seg[MAX_SEG]; };

constexpr int SMEM_RING = SW * W_SLOT + SX * X_SLOT;
constexpr int N_BARS = SW + CONS_R + 2 * SX + 2 * NACC;
constexpr int SMEM_MISC = LNW_KB * BK * 2 + N_BARS * 8 + 16 + 32 * 4 + 4 * 32 * 4 + (int)sizeof(CtaSched) + SW * 4 + 64;
constexpr int SMEM_BYTES = SMEM_RING + 1024 + ((SMEM_MISC + 127) / 128) * 128;

struct SkParams {
  const Sched* sched;
  void* out;
  const void* residual;
  float* ws;                 // partials [tile][MAX_SPLIT][PART_STRIDE]
  uint32_t* tile_ctr;        // per-tile arrival counters (zero between launches)
  const float* ssq_in;       // fused RMSNorm (nullptr = off)
  const void* ln_w;
  float* ssq_out;
  int64_t ldo, ld_res;
  int M, N, K, mode, ssq_tiles;
  float eps;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

__device__ __forceinline__ void spin_until(const uint32_t* ctr, uint32_t target) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (ld_acquire_gpu(ctr) < target) {
    if ((++spins & 0xFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 6000000000ll) {
        printf("linear_sk_kernel: stream-K fix-up timed out (block %d, target %u, have %u)\n", blockIdx.x, target, ld_acquire_gpu(ctr));
        __trap();
      }
    }
  }
}

template <class T> __device__ __forceinline__ T ldcg_t(const T* p) {
  const unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
  return *reinterpret_cast<const T*>(&u);
}

template <class T>
__global__ void __launch_bounds__(THREADS, 2)
linear_sk_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const SkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wring = smem;
  uint8_t* xring = smem + SW * W_SLOT;
  uint8_t* s_lnw = smem + SMEM_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_RING + LNW_KB * BK * 2);
  uint64_t* w_full = bars;
  uint64_t* cons = w_full + SW;
  uint64_t* x_full = cons + CONS_R;
  uint64_t* xn_full = x_full + SX;
  uint64_t* acc_full = xn_full + SX;
  uint64_t* acc_empty = acc_full + NACC;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + NACC);
  float* s_rstd = reinterpret_cast<float*>(tmem_ptr_smem + 4);
  float* s_ssq = s_rstd + 32;            // [4][32]
  CtaSched* sched = reinterpret_cast<CtaSched*>(s_ssq + 4 * 32);
  uint32_t* s_wkb = reinterpret_cast<uint32_t*>(sched + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const bool gu = p.mode == RD_SK_SWIGLU;
  const bool norm = p.ssq_in != nullptr;

  pdl_launch_dependents();
  if (threadIdx.x >= 64 && threadIdx.x < 64 + MAX_SEG) {
    const int i = threadIdx.x - 64;
    sched->seg[i] = p.sched->seg[cta][i];
    if (i == 0) sched->nseg = p.sched->nseg[cta];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < SW; ++i) mbar_init(&w_full[i], 1);
    for (int i = 0; i < CONS_R; ++i) mbar_init(&cons[i], 1);
    for (int i = 0; i < SX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&xn_full[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_THREADS); }
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
  const int ns = sched->nseg;
  const int reps = gu ? 2 : 1;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // pass 0 (before griddepcontrol.wait): the first SW weight tiles - they never depend on the previous kernel;
    // pass 1: everything else, token tiles included.
    uint32_t pre_w = 0;
    for (int pass = 0; pass < 2; ++pass) {
      uint32_t wslot = 0, wi = 0, ki = 0, xslot = 0;
      if (pass == 1) pdl_wait();
      for (int s = 0; s < ns; ++s) {
        const Seg sg = sched->seg[s];
        const int row = sg.tile * TILE_N;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          if (pass == 0 && wi + (uint32_t)reps > (uint32_t)SW) { s = ns; break; }
          uint32_t sl[2] = {0, 0};
          bool need[2] = {false, false};
          for (int r = 0; r < reps; ++r) {
            sl[r] = wslot;
            need[r] = pass == 0 || wi >= pre_w;
            if (need[r] && wi >= (uint32_t)SW) {        // the slot's previous tile must have been read by its MMAs
              const uint32_t old = s_wkb[wslot];
              mbar_wait(&cons[old & (CONS_R - 1)], (old / CONS_R) & 1u, 1);
            }
            ++wi;
            if (++wslot == SW) wslot = 0;
          }
          if (pass == 1 && ki >= (uint32_t)SX) {
            const uint32_t old = ki - SX;
            mbar_wait(&cons[old & (CONS_R - 1)], (old / CONS_R) & 1u, 2);
          }
          __syncwarp();
          if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
              if (!need[r]) continue;
              s_wkb[sl[r]] = ki;
              mbar_expect_tx(&w_full[sl[r]], W_SLOT);
              tma_load_2d(wring + sl[r] * W_SLOT, &map_w, &w_full[sl[r]], kb * BK, r == 0 ? row : p.N + row, HINT_EVICT_FIRST);
            }
            if (pass == 1) {
              mbar_expect_tx(&x_full[xslot], X_SLOT);
              tma_load_2d(xring + xslot * X_SLOT, &map_x, &x_full[xslot], kb * BK, 0, HINT_EVICT_LAST);
            }
          }
          __syncwarp();
          ++ki;
          if (++xslot == SX) xslot = 0;
        }
      }
      if (pass == 0) pre_w = wi;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(Tr<T>::umma_fmt, TILE_N, NT);
    const uint64_t dw0 = make_smem_desc(smem_u32(wring)), dx0 = make_smem_desc(smem_u32(xring));
    uint32_t ws = 0, wph = 0, xs = 0, xph = 0, ci = 0;
    for (int s = 0; s < ns; ++s) {
      const Seg sg = sched->seg[s];
      const uint32_t a = (uint32_t)s % NACC;
      mbar_wait(&acc_empty[a], (((uint32_t)s / NACC) & 1u) ^ 1u, 3);
      tc_fence_after();
      const uint32_t d0 = tmem_base + a * ACC_COLS;
      for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
        const uint32_t ws0 = ws;
        mbar_wait(&w_full[ws0], wph, 4);
        if (++ws == SW) { ws = 0; wph ^= 1u; }
        uint32_t ws1 = ws0;
        if (gu) {
          ws1 = ws;
          mbar_wait(&w_full[ws1], wph, 5);
          if (++ws == SW) { ws = 0; wph ^= 1u; }
        }
        mbar_wait(norm ? &xn_full[xs] : &x_full[xs], xph, 6);
        tc_fence_after();
        const uint64_t da = dw0 + (uint64_t)(ws0 * (W_SLOT >> 4)), du = dw0 + (uint64_t)(ws1 * (W_SLOT >> 4));
        const uint64_t db = dx0 + (uint64_t)(xs * (X_SLOT >> 4));
        const uint32_t accf = kb > sg.kb0 ? 1u : 0u;
        __syncwarp();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t koff = (uint64_t)((k * UK * 2) >> 4);
            const uint32_t af = (accf || k > 0) ? 1u : 0u;
            tc_mma_f16(d0, da + koff, db + koff, idesc, af);
            if (gu) tc_mma_f16(d0 + NT, du + koff, db + koff, idesc, af);
          }
          tc_commit(&cons[ci]);                   // frees this k-block's W and X slots once the MMAs have read them
          if (kb == sg.kb1 - 1) tc_commit(&acc_full[a]);
        }
        __syncwarp();
        ci = (ci + 1) & (CONS_R - 1);
        if (++xs == SX) { xs = 0; xph ^= 1u; }
      }
    }
  } else {
    // ===================== workers (warps 2-5) =====================
    const int ww = warp - 2, wtid = threadIdx.x - 64;
    const int quad = warp & 3;
    const int n_local = quad * 32 + lane;
    const int B = p.M;
    if (norm) {
      // RMSNorm weights of this CTA's k-blocks -> shared memory (constants: before griddepcontrol.wait)
      const T* lnw = reinterpret_cast<const T*>(p.ln_w);
      int n_x = 0;
      for (int s = 0; s < ns; ++s) n_x += sched->seg[s].kb1 - sched->seg[s].kb0;
      n_x = n_x > LNW_KB ? LNW_KB : n_x;
      for (int idx = wtid; idx < n_x * 8; idx += EPI_THREADS) {
        int i = idx >> 3, kb = 0;
        for (int s = 0; s < ns; ++s) {
          const int len = sched->seg[s].kb1 - sched->seg[s].kb0;
          if (i < len) { kb = sched->seg[s].kb0 + i; break; }
          i -= len;
        }
        *reinterpret_cast<uint4*>(s_lnw + (size_t)idx * 16) = *reinterpret_cast<const uint4*>(lnw + kb * BK + (idx & 7) * 8);
      }
    }
    pdl_wait();
    if (norm && ns > 0) {
      // rstd[j] = rsqrt(mean(x_j^2) + eps) from the producer's per-tile partials, fixed order (modeling_llama_imgemb.py:85-93)
      const int j = wtid & 31, part = wtid >> 5;
      float v = 0.f;
#pragma unroll 4
      for (int t = part; t < p.ssq_tiles; t += 4) v += __ldcg(p.ssq_in + t * 32 + j);
      s_ssq[part * 32 + j] = v;
      epi_bar();
      if (wtid < B) {
        const float tot = ((s_ssq[wtid] + s_ssq[32 + wtid]) + s_ssq[64 + wtid]) + s_ssq[96 + wtid];
        s_rstd[wtid] = 1.0f / sqrtf(tot / (float)p.K + p.eps);
      }
      epi_bar();
      // normalise the token tiles in place, xn = T(w * T(x * rstd)); each warp owns every 4th tile
      const T* lnw = reinterpret_cast<const T*>(p.ln_w);
      uint32_t xs = 0, xph = 0;
      int it = 0;
      for (int s = 0; s < ns; ++s) {
        const Seg sg = sched->seg[s];
        for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++it) {
          const uint32_t slot = xs, par = xph;
          if (++xs == SX) { xs = 0; xph ^= 1u; }
          if ((it & (EPI_WARPS - 1)) != ww) continue;
          if (lane == 0) mbar_wait(&x_full[slot], par, 7);
          __syncwarp();
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
    }
    // ---- epilogues of this CTA's segments ----
    const T* resg = reinterpret_cast<const T*>(p.residual);
    T* outg = reinterpret_cast<T*>(p.out);
    for (int s = 0; s < ns; ++s) {
      const Seg sg = sched->seg[s];
      const uint32_t a = (uint32_t)s % NACC;
      const int n = sg.tile * TILE_N + n_local;
      const bool contributor = sg.nsplits > 1 && sg.split != 0;
      const bool finaliser = sg.nsplits > 1 && sg.split == 0;
      if (wtid == 0) {
        mbar_wait(&acc_full[a], ((uint32_t)s / NACC) & 1u, 8);
        if (finaliser) {
          spin_until(p.tile_ctr + sg.tile, (uint32_t)sg.nsplits - 1);
          p.tile_ctr[sg.tile] = 0;                  // re-arm for the next launch
        }
      }
      epi_bar();
      tc_fence_after();
      if (finaliser) __threadfence();
      const uint32_t tbase = tmem_base + a * ACC_COLS + ((uint32_t)(quad * 32) << 16);
      float* part = p.ws + ((int64_t)sg.tile * MAX_SPLIT + sg.split) * PART_STRIDE + n_local;
      const float* ps0 = p.ws + (int64_t)sg.tile * MAX_SPLIT * PART_STRIDE + n_local;
#pragma unroll 1
      for (int col0 = 0; col0 < NT; col0 += 16) {       // two halves of 16 token columns (register budget)
        float acc[16], accu[16];
        {
          uint32_t r0[16], r1[16];
          tc_ld16(tbase + col0, r0);
          if (gu) tc_ld16(tbase + NT + col0, r1);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) { acc[j] = __uint_as_float(r0[j]); accu[j] = gu ? __uint_as_float(r1[j]) : 0.f; }
        }
        if (col0 == 16) {                                 // both halves are in registers / consumed: free the accumulator
          tc_fence_before();
          mbar_arrive(&acc_empty[a]);
        }
        if (col0 >= B) continue;
        if (contributor) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (col0 + j < B) {
              __stcg(part + (col0 + j) * TILE_N, acc[j]);
              if (gu) __stcg(part + (NT + col0 + j) * TILE_N, accu[j]);
            }
          }
          continue;
        }
        float res[16];
        if (p.mode == RD_SK_RES1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) res[j] = (col0 + j < B && n < p.N) ? Tr<T>::f(ldcg_t(resg + (int64_t)(col0 + j) * p.ld_res + n)) : 0.f;
        }
        if (finaliser) {
          if (gu) {
#pragma unroll 1
            for (int sp = 1; sp < sg.nsplits; ++sp) {
              const float* pa = ps0 + (int64_t)sp * PART_STRIDE;
              float v0[16], u0[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const bool ok = col0 + j < B;
                v0[j] = ok ? __ldcg(pa + (col0 + j) * TILE_N) : 0.f;
                u0[j] = ok ? __ldcg(pa + (NT + col0 + j) * TILE_N) : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) { acc[j] += v0[j]; accu[j] += u0[j]; }
            }
          } else {
#pragma unroll 1
            for (int sp = 1; sp < sg.nsplits; sp += 4) {   // four contributors (64 loads) in flight per round
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
        if (n < p.N) {
          if (p.mode == RD_SK_PLAIN) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + j < B) outg[(int64_t)(col0 + j) * p.ldo + n] = Tr<T>::r(acc[j]);
          } else if (p.mode == RD_SK_SWIGLU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (col0 + j < B) {
                const float gg = Tr<T>::rr(acc[j]), uu = Tr<T>::rr(accu[j]);
                outg[(int64_t)(col0 + j) * p.ldo + n] = Tr<T>::r(Tr<T>::rr(silu_f(gg)) * uu);     // T(T(silu(T(g))) * T(u))
              }
            }
          }
        }
        if (p.mode == RD_SK_RES1) {
          float yy[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            yy[j] = 0.f;
            if (col0 + j < B && n < p.N) {
              const T y = Tr<T>::r(res[j] + Tr<T>::rr(acc[j]));        // residual add in the storage dtype
              outg[(int64_t)(col0 + j) * p.ldo + n] = y;
              const float f = Tr<T>::f(y);
              yy[j] = f * f;
            }
          }
          if (p.ssq_out != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float t = warp_sum(yy[j]);
              if (lane == 0) s_ssq[quad * 32 + col0 + j] = t;
            }
          }
        }
      }
      if (contributor) {
        __threadfence();
        epi_bar();
        if (wtid == 0) atomicAdd(p.tile_ctr + sg.tile, 1u);
      } else if (p.mode == RD_SK_RES1 && p.ssq_out != nullptr) {
        epi_bar();
        if (wtid < B) __stcg(p.ssq_out + sg.tile * 32 + wtid, ((s_ssq[wtid] + s_ssq[32 + wtid]) + s_ssq[64 + wtid]) + s_ssq[96 + wtid]);
        epi_bar();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

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
  RD_REQUIRE(r == CUDA_SUCCESS, "linear_sk: cuTensorMapEncodeTiled failed (%d) ptr=%p ld=%lld rows=%d K=%d", (int)r, ptr, (long long)ld, rows, K);
  return RD_OK;
}

// equal contiguous runs of (tile, k-block) units over G CTAs; false if a CTA would need more than MAX_SEG segments or a tile
// more than MAX_SPLIT contributors
bool build_sched(Sched* sc, int tiles, int kb, int G, int G_all) {
  for (int c = 0; c < G_all; ++c) sc->nseg[c] = 0;
  const long long U = (long long)tiles * kb;
  std::vector<int> per_tile(tiles, 0);
  for (int c = 0; c < G; ++c) {
    long long u0 = U * c / G, u1 = U * (c + 1) / G;
    while (u0 < u1) {
      const int t = (int)(u0 / kb), k0 = (int)(u0 % kb);
      const int k1 = (int)((u1 - u0) < (kb - k0) ? k0 + (u1 - u0) : kb);
      int& n = sc->nseg[c];
      if (n >= MAX_SEG) return false;
      Seg& s = sc->seg[c][n++];
      s.tile = t; s.kb0 = k0; s.kb1 = k1; s.split = per_tile[t]++; s.nsplits = 0;
      u0 += k1 - k0;
    }
  }
  for (int t = 0; t < tiles; ++t) if (per_tile[t] > MAX_SPLIT) return false;
  for (int c = 0; c < G; ++c)
    for (int i = 0; i < sc->nseg[c]; ++i) sc->seg[c][i].nsplits = per_tile[sc->seg[c][i].tile];
  return true;
}

struct Plan { Sched* dev = nullptr; int tiles = 0, kbs = 0, max_seg = 0; };

}  // namespace

struct rd_sk {
  int G = 0;
  std::map<long long, Plan> plans;
  float* ws = nullptr;
  uint32_t* ctr = nullptr;
  int ws_tiles = 0;
};

int rd_sk_create(rd_sk** out) {
  RD_REQUIRE(out, "rd_sk_create: null argument");
  int dev = 0, sms = 0;
  RD_CHECK_CUDA(cudaGetDevice(&dev));
  RD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  rd_sk* c = new rd_sk();
  c->G = sms < MAX_G ? sms : MAX_G;
  if (const char* e = getenv("RD_SK_CTAS_PER_SM")) { const int v = atoi(e); if (v == 2 && 2 * sms <= MAX_G) c->G = 2 * sms; }
  *out = c;
  return RD_OK;
}

void rd_sk_destroy(rd_sk* c) {
  if (!c) return;
  for (auto& kv : c->plans) if (kv.second.dev) cudaFree(kv.second.dev);
  if (c->ws) cudaFree(c->ws);
  if (c->ctr) cudaFree(c->ctr);
  delete c;
}

static long long plan_key(int N, int K, int mode) { return ((long long)N << 34) | ((long long)K << 4) | (long long)mode; }

int rd_sk_plan(rd_sk* c, int N, int K, int mode) {
  RD_REQUIRE(c, "rd_sk_plan: null context");
  RD_REQUIRE(K % BK == 0, "rd_sk_plan: K=%d must be a multiple of %d", K, BK);
  const long long key = plan_key(N, K, mode);
  if (c->plans.count(key)) return RD_OK;
  const int tiles = (N + TILE_N - 1) / TILE_N, kbs = K / BK;
  std::vector<Sched> hs(1);
  memset(&hs[0], 0, sizeof(Sched));
  const long long U = (long long)tiles * kbs;
  int G0 = (int)(U / 4 > 0 ? U / 4 : 1);          // at least ~4 k-blocks per active CTA
  G0 = G0 > c->G ? c->G : G0;
  bool ok = false;
  for (int g = G0; g >= 1 && !ok; --g) ok = build_sched(&hs[0], tiles, kbs, g, c->G);
  RD_REQUIRE(ok, "rd_sk_plan: cannot schedule N=%d K=%d", N, K);
  Plan pl;
  pl.tiles = tiles; pl.kbs = kbs;
  for (int i = 0; i < c->G; ++i) pl.max_seg = hs[0].nseg[i] > pl.max_seg ? hs[0].nseg[i] : pl.max_seg;
  RD_CHECK_CUDA(cudaMalloc((void**)&pl.dev, sizeof(Sched)));
  RD_CHECK_CUDA(cudaMemcpy(pl.dev, &hs[0], sizeof(Sched), cudaMemcpyHostToDevice));
  if (tiles > c->ws_tiles) {
    if (c->ws) cudaFree(c->ws);
    if (c->ctr) cudaFree(c->ctr);
    c->ws = nullptr; c->ctr = nullptr;
    RD_CHECK_CUDA(cudaMalloc((void**)&c->ws, (size_t)tiles * MAX_SPLIT * PART_STRIDE * 4));
    RD_CHECK_CUDA(cudaMalloc((void**)&c->ctr, (size_t)(tiles + 32) * 4));
    RD_CHECK_CUDA(cudaMemset(c->ctr, 0, (size_t)(tiles + 32) * 4));
    c->ws_tiles = tiles;
  }
  c->plans[key] = pl;
  return RD_OK;
}

template <class T>
static int launch_sk(rd_sk* c, const Plan& pl, const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N,
                     int K, int mode, const void* residual, int64_t ld_res, const SkNorm* norm, float* ssq_out, int dtype, cudaStream_t st) {
  RD_SMEM_ATTR_ONCE(SMEM_BYTES, linear_sk_kernel<T>);
  CUtensorMap map_w, map_x;
  RD_CHECK(make_map(&map_w, w, ldw, mode == RD_SK_SWIGLU ? 2 * N : N, K, TILE_N, dtype));
  RD_CHECK(make_map(&map_x, x, ldx, M, K, NT, dtype));
  SkParams p{};
  p.sched = pl.dev; p.out = out; p.residual = residual; p.ws = c->ws; p.tile_ctr = c->ctr;
  p.ssq_in = norm ? norm->ssq : nullptr; p.ln_w = norm ? norm->ln_w : nullptr; p.ssq_tiles = norm ? norm->ssq_tiles : 0;
  p.eps = norm ? norm->eps : 0.f;
  p.ssq_out = ssq_out; p.ldo = ldo; p.ld_res = ld_res; p.M = M; p.N = N; p.K = K; p.mode = mode;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c->G); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = rd_pdl_enabled() ? 1 : 0;
  RD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, linear_sk_kernel<T>, map_w, map_x, p));
  return RD_OK;
}

int rd_sk_max_segments(rd_sk* c, int N, int K, int mode) {
  if (!c || rd_sk_plan(c, N, K, mode) != RD_OK) return 1 << 30;
  return c->plans[plan_key(N, K, mode)].max_seg;
}

int rd_sk_linear(rd_sk* c, const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                 int mode, const void* residual, int64_t ld_res, const SkNorm* norm, float* ssq_out, int dtype, cudaStream_t st) {
  RD_REQUIRE(c && x && w && out, "rd_sk_linear: null argument");
  RD_REQUIRE(M > 0 && M <= NT, "rd_sk_linear: M=%d out of range (1..%d)", M, NT);
  RD_REQUIRE(mode != RD_SK_RES1 || residual != nullptr, "rd_sk_linear: RES1 needs a residual");
  RD_CHECK(rd_sk_plan(c, N, K, mode));
  const Plan& pl = c->plans[plan_key(N, K, mode)];
  RD_REQUIRE(norm == nullptr || pl.max_seg <= NACC, "rd_sk_linear: fused RMSNorm needs <= %d segments per CTA (shape N=%d K=%d has %d)",
             NACC, N, K, pl.max_seg);
  RD_DISPATCH_DTYPE(dtype, T, {
    return launch_sk<T>(c, pl, x, ldx, w, ldw, out, ldo, M, N, K, mode, residual, ld_res, norm, ssq_out, dtype, st);
  });
}
