// Persistent tcgen05 GEMM for wide token counts (prefill, the ResNet convolutions, the image-token side of the Q-Former):
//
//   out[M,N] = epilogue( x[M,K] . W[N,K]^T ),   M > 128, fp16/bf16, fp32 accumulate            (same contract as linear_tc.cu)
//
// linear_tc.cu runs one (128 weight rows x NT tokens) tile per CTA: TMEM alloc, pipeline fill and - above all - the epilogue
// (128 x 256 outputs per tile) are serial with the tile's MMAs, and with one 200 KB CTA per SM nothing else covers them.  On
// the prefill shapes that is 25-50 % of a tile's time (o_proj: 28 us of MMA in a 50 us tile); on the K = 64 layer-1
// convolutions the tile is ALL epilogue.  This kernel keeps the same swap-AB tile (weights = UMMA A, tokens = UMMA B, 128 x NT)
// but is persistent - one CTA per SM walks a static list of tiles - and splits the work so that all three stages overlap:
//
//   warp 0    TMA producer: 2..4-stage ring of (16 KB weight k-block + NT x 128 B token k-block), runs ahead across tiles
//   warp 1    tcgen05.mma issuer; the 512 TMEM columns hold TWO accumulator buffers, tile i+1 accumulates while
//   warps 2-9 drain tile i (two warps per TMEM lane quadrant, each half of a chunk's token columns):
//             tcgen05.ld -> registers -> epilogue arithmetic -> shared memory -> TMA store, in chunks of CR <= 32
//             tokens.  The residual tile arrives by TMA too (3..15 chunks ahead, same shared-memory buffer the result is
//             written back into), so the epilogue issues no per-thread global loads or stores at all.
//
// SwiGLU (gate|up) tiles hold 64 gate rows + 64 up rows of W in ONE 128-row A tile (two 64-row TMA boxes), so the MMA shape
// and the single 256-column accumulator are the same as for every other GEMM and double buffering still fits in TMEM; the two
// lane halves exchange T(g) / T(u) through 8 KB of shared memory and all eight epilogue warps share the silu work.
//
// The token-tile width NT is a launch parameter (any multiple of 16 up to 256): the host picks the NT that wastes the
// fewest MMA cycles to round quantisation (tiles / SMs), e.g. 240 instead of 256 for the 2048-token prefill o_proj/down
// (288 tiles = 1.95 rounds instead of 256 = 1.73 -> 2).
//
// Arithmetic and rounding points are those of linear_tc.cu's epilogues (T(Wx) then the residual add, fp32 bias/act, the
// SwiGLU triple rounding); the k-blocks are accumulated in the same order, so the two kernels agree bit for bit.
#include <cuda.h>
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "tc_ptx.cuh"

bool rd_pdl_enabled();
int rd_tc_make_map(CUtensorMap* map, const void* ptr, int64_t ld, int rows, int cols, int box_rows, int box_cols, int dtype, int swizzle128);

static int g_wide_persistent = 1;    // test hook: 0 = wide shapes stay on linear_tc_kernel
extern "C" int rd_linear_wide_persistent(int on) { g_wide_persistent = on; return RD_OK; }
static int g_wide_force_nt = 0;      // test hook: token-tile width (0 = heuristic)
extern "C" int rd_linear_wide_force_nt(int nt) { g_wide_force_nt = nt; return RD_OK; }
static int g_wide_force_stages = 0;  // test hook: pipeline stages 2..4 (0 = by K)
extern "C" int rd_linear_wide_force_stages(int s) { g_wide_force_stages = (s >= 2 && s <= 6) ? s : 0; return RD_OK; }
static int g_wide_pair = 1;          // 1: CTA pairs (cta_group::2, 256-row UMMA, each CTA stages half of the token tile)
extern "C" int rd_linear_wide_pair(int on) { g_wide_pair = on; return RD_OK; }
static int g_wide_min_tiles = 1;     // test hook: GEMMs with fewer output tiles stay on the one-tile-per-CTA kernel
extern "C" int rd_linear_wide_min_tiles(int n) { g_wide_min_tiles = n; return RD_OK; }

namespace {
using namespace tcptx;

constexpr int WROWS = 128;                       // weight rows per tile (UMMA M)
constexpr int BK = 64;                           // 64 x 2 B = one 128-byte swizzle row
constexpr int UK = 16;
constexpr int W_BYTES = WROWS * BK * 2;          // 16 KB
constexpr int X_BYTES_MAX = 256 * BK * 2;        // 32 KB
constexpr int STAGE_BYTES = W_BYTES + X_BYTES_MAX;
constexpr int PAIR_X_BYTES_MAX = 128 * BK * 2;   // CTA pair: each CTA stages half of the token tile
constexpr int PAIR_STAGE_BYTES = W_BYTES + PAIR_X_BYTES_MAX;   // 32 KB
constexpr int MAX_STAGES = 4;
constexpr int MAX_STAGES_PAIR = 6;
constexpr int MAX_NBUF = 16;
constexpr int BUF_BYTES = 32 * 128 * 2;          // epilogue chunk buffer: CR <= 32 token rows of 128 features (residual in, result out)
static_assert(MAX_STAGES * STAGE_BYTES == MAX_STAGES_PAIR * PAIR_STAGE_BYTES, "both variants split the same 192 KB ring");
constexpr int RING_EPI_BYTES = MAX_STAGES * STAGE_BYTES + 4 * BUF_BYTES;   // 224 KB split between the pipeline and the chunk buffers:
                                                 // 4 stages + 4 buffers (deep K), 3 + 10, or 2 + 16 (K <= 128: the tile is all epilogue)
constexpr int BAR_BYTES = 512;
constexpr int SMEM_BYTES = RING_EPI_BYTES + BAR_BYTES + 1024;
constexpr int EPI_WARPS = 8;                     // two warps per TMEM lane quadrant: each takes half of a chunk's token columns
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr uint64_t HINT_NORMAL = 0x1000000000000000ull;

enum { W_PLAIN = 0, W_RES1 = 1, W_AFFINE = 2 };

struct WideParams {
  int M, N, K;
  int NT, CR;                 // token-tile width, epilogue chunk rows (CR divides NT, CR in {16, 32})
  int m_tiles, n_tiles, tiles;
  int m_fast;                 // consecutive tiles (= CTAs running together) share the weight tile
  int mode, act, has_res;
  int stages, nbuf;           // pipeline stages, 8 KB epilogue chunk buffers
  // implicit-GEMM convolution (conv_ks > 0): the token operand is gathered from the NHWC activation [B, H, W, C] by im2col-mode TMA
  // loads - token m = output pixel (b, oh, ow), k-block kb = 64 channels of filter tap (kb * 64) / C - instead of read from a
  // materialised [M, ks*ks*C] matrix
  int conv_ks, conv_C, conv_OW, conv_OHW, conv_stride, conv_pad;
  const float* bias;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(EPI_THREADS) : "memory"); }

// ---- CTA pair (cta_group::2) helpers: the two CTAs of a cluster run one 256-row UMMA; the even CTA (rank 0) issues it ----
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {      // shared::cluster address of CTA `rank`'s copy
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory whose transaction bytes are counted on the mbarrier at cluster address `bar_cluster`
// (the leader CTA's "stage full" barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// im2col-mode TMA load (NHWC tensor map from cuTensorMapEncodeIm2col): `pixelsPerColumn` output pixels starting at the window whose
// top-left input position is (w, h) of image n, 64 channels from c, filter tap (off_w, off_h); out-of-image taps are zero-filled
__device__ __forceinline__ void tma_load_im2col(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c, int w, int h, int n,
                                                     uint16_t off_w, uint16_t off_h, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h),
        "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// MMA completion -> the mbarrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

template <class T, bool SWIGLU, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
linear_wide_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                   const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res, const WideParams p) {
  constexpr int FEATS = SWIGLU ? 64 : 128;                   // output features per tile
  constexpr int CBUF = SWIGLU ? BUF_BYTES / 2 : BUF_BYTES;   // bytes between chunk buffers
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int STAGES = p.stages;
  constexpr int SB = PAIR ? PAIR_STAGE_BYTES : STAGE_BYTES;    // bytes per pipeline stage
  constexpr int MAXS = PAIR ? MAX_STAGES_PAIR : MAX_STAGES;
  // CTA pair: both CTAs walk the same list of (256 weight rows x NT tokens) tiles; CTA `rank` owns weight rows [128 rank, +128)
  // (its A operand and its accumulator lanes) and stages token rows [NT/2 rank, +NT/2) of the shared B operand
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;           // index / count of the tile walkers
  const int n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // SwiGLU result rows are 64 features wide: twice as many (half-size) buffers, minus the T(g) / T(u) exchange (2 x 4 KB)
  const int NBUF = SWIGLU ? (2 * p.nbuf - 2 < MAX_NBUF ? 2 * p.nbuf - 2 : MAX_NBUF) : p.nbuf;
  uint8_t* epi_s = smem + STAGES * SB;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_EPI_BYTES);
  uint64_t* empty_bar = full_bar + MAXS;
  uint64_t* tfull = empty_bar + MAXS;            // [2] accumulator buffer complete
  uint64_t* tempty = tfull + 2;                  // [2] accumulator buffer drained (every epilogue thread of the tile arrives; pair: on the leader's)
  uint64_t* res_full = tempty + 2;               // [NBUF] residual chunk landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_full + MAX_NBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_total = (p.K + BK - 1) / BK;
  const int NT = p.NT, CR = p.CR;
  const int XROWS = PAIR ? NT / 2 : NT;          // token rows this CTA stages per k-block
  const uint32_t stage_tx = (uint32_t)((PAIR ? 2 : 1) * (W_BYTES + XROWS * BK * 2));   // pair: both CTAs' bytes land on the leader's barrier

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_out)) : "memory");
    if (p.has_res) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_res)) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], PAIR ? 2 * EPI_THREADS : EPI_THREADS); }
    for (int b = 0; b < MAX_NBUF; ++b) mbar_init(&res_full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();      // the peer's TMA loads and barrier arrivals target this CTA's (now initialised) mbarriers
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto tile_coords = [&](int t, int& m0, int& n0) {
    int mt, nt;
    if (p.m_fast) { mt = t % p.m_tiles; nt = t / p.m_tiles; } else { nt = t % p.n_tiles; mt = t / p.n_tiles; }
    m0 = mt * NT;
    n0 = (PAIR ? nt * 2 + (int)rank : nt) * FEATS;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    auto load = [&](void* dst, const CUtensorMap* map, int s, int c0, int c1) {
      if (PAIR) tma_load_2d_pair(dst, map, mapa_u32(smem_u32(&full_bar[s]), 0u), c0, c1, HINT_NORMAL);
      else tma_load_2d(dst, map, &full_bar[s], c0, c1, HINT_NORMAL);
    };
    auto load_w = [&](int s, int kb, int n0) {
      uint8_t* sp = smem + s * SB;
      if (SWIGLU) {
        load(sp, &map_w, s, kb * BK, n0);                        // 64 gate rows
        load(sp + W_BYTES / 2, &map_w, s, kb * BK, p.N + n0);    // 64 up rows
      } else {
        load(sp, &map_w, s, kb * BK, n0);
      }
    };
    const bool expects = !PAIR || rank == 0;     // pair: the leader's barrier carries the expected byte count of both CTAs
    // token operand of k-block kb for the tile whose first token is m0
    int cv_w = 0, cv_h = 0, cv_n = 0;            // conv: window origin of this CTA's first output pixel of the current tile
    auto set_tile_x = [&](int m0) {
      if (p.conv_ks > 0) {
        int ms = m0 + (int)rank * XROWS;
        if (ms >= p.M) ms = 0;                    // pair: the peer's half lies wholly past the last pixel - any valid window (results clipped)
        cv_n = ms / p.conv_OHW;
        const int r = ms - cv_n * p.conv_OHW, oh = r / p.conv_OW, ow = r - oh * p.conv_OW;
        cv_w = ow * p.conv_stride - p.conv_pad;
        cv_h = oh * p.conv_stride - p.conv_pad;
      }
    };
    auto load_x = [&](int s, int kb, int m0) {
      uint8_t* dst = smem + s * SB + W_BYTES;
      if (p.conv_ks > 0) {
        const int k0 = kb * BK, tap = k0 / p.conv_C, c0 = k0 - tap * p.conv_C, kh = tap / p.conv_ks, kw = tap - kh * p.conv_ks;
        if (PAIR) tma_load_im2col_pair(dst, &map_x, mapa_u32(smem_u32(&full_bar[s]), 0u), c0, cv_w, cv_h, cv_n, (uint16_t)kw, (uint16_t)kh, HINT_NORMAL);
        else tma_load_im2col(dst, &map_x, &full_bar[s], c0, cv_w, cv_h, cv_n, (uint16_t)kw, (uint16_t)kh);
      } else {
        load(dst, &map_x, s, kb * BK, m0 + (int)rank * XROWS);
      }
    };
    // weights never depend on the previous kernel: the first stages' weight k-blocks are requested before the PDL wait
    int pre = 0;
    if (unit < p.tiles) {
      int m0, n0;
      tile_coords(unit, m0, n0);
      pre = kb_total < STAGES ? kb_total : STAGES;
      if (elect_one()) {
        for (int i = 0; i < pre; ++i) {
          if (expects) mbar_expect_tx(&full_bar[i], stage_tx);
          load_w(i, i, n0);
        }
      }
      __syncwarp();
    }
    pdl_wait();
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int t = unit; t < p.tiles; t += n_units) {
      int m0, n0;
      tile_coords(t, m0, n0);
      set_tile_x(m0);
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1u, 1);        // first pass: the barrier's "previous phase" counts as complete
        __syncwarp();
        if (elect_one()) {
          if (it >= pre) {
            if (expects) mbar_expect_tx(&full_bar[s], stage_tx);
            load_w(s, kb, n0);
          }
          load_x(s, kb, m0);
        }
        __syncwarp();
        if (it < STAGES) ++it;
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (pair: the leader CTA issues the 256-row UMMA for both) =====================
    const uint32_t idesc = make_idesc(Tr<T>::umma_fmt, PAIR ? 2 * WROWS : WROWS, NT);
    const uint64_t d0 = make_smem_desc(smem_u32(smem));
    int s = 0, tl = 0;
    uint32_t ph = 0;
    for (int t = unit; t < p.tiles; t += n_units, ++tl) {
      const int buf = tl & 1;
      mbar_wait(&tempty[buf], (uint32_t)((tl >> 1) & 1) ^ 1u, 2);      // epilogue(s) have drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full_bar[s], ph, 3);
        tc_fence_after();
        __syncwarp();
        if (elect_one()) {
          const uint64_t da = d0 + (uint64_t)((uint32_t)s * (SB >> 4));
          const uint64_t db = da + (uint64_t)(W_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t koff = (uint64_t)((k * UK * 2) >> 4);
            if (PAIR) tc_mma_f16_pair(d_tmem, da + koff, db + koff, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else tc_mma_f16(d_tmem, da + koff, db + koff, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) {
            tc_commit_pair(&empty_bar[s]);
            if (kb == kb_total - 1) tc_commit_pair(&tfull[buf]);
          } else {
            tc_commit(&empty_bar[s]);
            if (kb == kb_total - 1) tc_commit(&tfull[buf]);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= 2) {
    // ===================== epilogue (warps 2..9: TMEM lane quadrant = warp & 3; `sub` = which half of a chunk's columns) ==========
    const int e = threadIdx.x - 64;
    const int quad = warp & 3;
    const int sub = (warp - 2) >> 2;
    const bool works = (CR == 32) || sub == 0;     // 16-token chunks are not split: the second warp of a quadrant only keeps the barriers
    const int n_local = quad * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int mode = p.mode, act = p.act;
    const bool has_res = p.has_res != 0;
    auto n_chunks_of = [&](int t) {
      int m0, n0;
      tile_coords(t, m0, n0);
      const int rows = p.M - m0 < NT ? p.M - m0 : NT;
      return (rows + CR - 1) / CR;
    };
    pdl_wait();                                   // residual reads and every store come after the previous kernel
    // residual prefetch iterator (thread e == 32): runs D chunks ahead of the chunk being finished
    // (the tile's coordinates and chunk count are kept between calls: the per-chunk path is an mbarrier arm + one TMA issue, no
    // integer divisions - this thread's warp sits on every chunk's critical path)
    int pf_t = unit, pf_c = 0, pf_b = 0, pf_m0 = 0, pf_n0 = 0, pf_nch = 0;
    if (pf_t < p.tiles) { tile_coords(pf_t, pf_m0, pf_n0); pf_nch = n_chunks_of(pf_t); }
    auto issue_res = [&]() {
      if (pf_t >= p.tiles) return;
      mbar_expect_tx(&res_full[pf_b], (uint32_t)(CR * FEATS * 2));
      tma_load_2d(epi_s + pf_b * CBUF, &map_res, &res_full[pf_b], pf_n0, pf_m0 + pf_c * CR, HINT_NORMAL);
      if (++pf_b == NBUF) pf_b = 0;
      if (++pf_c == pf_nch) {
        pf_c = 0;
        pf_t += n_units;
        if (pf_t < p.tiles) { tile_coords(pf_t, pf_m0, pf_n0); pf_nch = n_chunks_of(pf_t); }
      }
    };
    // stores allowed to be still reading their buffer after a new one is committed; the residual of chunk g + D goes into the
    // buffer chunk g - E - 1 was stored from.  The store side (thread 0) and the residual loads (thread 32) are separate
    // threads: each chunk's serial path is one TMA issue.
    const int E = NBUF >= 8 ? 4 : 1, D = NBUF - E - 1;
    if (has_res && e == 32) {
      for (int i = 0; i < D; ++i) issue_res();
    }
    int g = 0, tl = 0, cb = 0;                    // cb / cph: chunk buffer index and its mbarrier phase (g % NBUF, (g / NBUF) & 1)
    uint32_t cph = 0;
#ifdef RD_WIDE_PROF
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define PROF(i) { const long long _n = clock64(); pf[i] += _n - pt; pt = _n; }
#else
#define PROF(i)
#endif
    for (int t = unit; t < p.tiles; t += n_units, ++tl) {
      int m0, n0;
      tile_coords(t, m0, n0);
      const int buf = tl & 1;
      const int n = n0 + n_local;
      float bias_n = 0.f;
      if (!SWIGLU && mode == W_AFFINE && p.bias != nullptr && n < p.N) bias_n = p.bias[n];
      const int nch = n_chunks_of(t);
      PROF(0)
      mbar_wait(&tfull[buf], (uint32_t)((tl >> 1) & 1), 4);
      tc_fence_after();
      PROF(1)
      for (int c = 0; c < nch; ++c, ++g) {
        const int b = cb;
        T* bufp = reinterpret_cast<T*>(epi_s + b * CBUF);
        epi_bar(1);                               // buffer b and the exchange area are free (thread 0 has waited for the store)
        PROF(2)
        if (has_res) {
          if (e == 32) issue_res();               // chunk g + D
          mbar_wait(&res_full[b], cph, 5);
        }
        PROF(3)
        uint32_t r0[16];
        const int jb0 = (CR == 32) ? sub * 16 : 0;   // first token row of the chunk this thread finishes (16 rows)
        if (works) tc_ld16(taddr + (uint32_t)(buf * 256 + c * CR + jb0), r0);
        tc_wait_ld();
        PROF(4)
        if (c == nch - 1) {                       // accumulators of this tile are in registers: hand the buffer back
          tc_fence_before();
          if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0u));     // the leader's MMA warp waits for both CTAs
          else mbar_arrive(&tempty[buf]);
        }
        if (SWIGLU) {
          // lanes 0..63 hold the gate rows, lanes 64..127 the up rows of the same 64 features: both halves park T(.) of their
          // accumulators (the reference rounds g and u to the storage type before anything else), then ALL epilogue warps share the
          // silu / multiply work: thread -> (feature e & 63, quarter of the chunk's token rows e >> 6)
          T* exg = reinterpret_cast<T*>(epi_s + NBUF * CBUF);
          T* exu = exg + 32 * 64;
          T* mine = (quad >= 2 ? exu : exg) + (n_local & 63) + jb0 * 64;
          if (works) {
#pragma unroll
            for (int j = 0; j < 16; ++j) mine[j * 64] = Tr<T>::r(__uint_as_float(r0[j]));
          }
          epi_bar(3);
          const int f = e & 63, quarter = CR >> 2, j0 = (e >> 6) * quarter;     // 256 threads: 64 features x 4 groups of token rows
          if (CR == 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float gt = Tr<T>::f(exg[(j0 + j) * 64 + f]), u = Tr<T>::f(exu[(j0 + j) * 64 + f]);
              bufp[(j0 + j) * 64 + f] = Tr<T>::r(Tr<T>::rr(silu_f(gt)) * u);      // T(T(silu(T(g))) * T(u))
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float gt = Tr<T>::f(exg[(j0 + j) * 64 + f]), u = Tr<T>::f(exu[(j0 + j) * 64 + f]);
              bufp[(j0 + j) * 64 + f] = Tr<T>::r(Tr<T>::rr(silu_f(gt)) * u);
            }
          }
        } else {
          // the residual values of the 16 rows are read together before the first one is consumed (one shared-memory latency
          // per 16 elements instead of per element), and the activation switch sits outside the element loop
          auto finish = [&](const uint32_t (&r)[16], int jb) {
            T* q = bufp + jb * 128 + n_local;
            float res[16];
            if (has_res) {
#pragma unroll
              for (int j = 0; j < 16; ++j) res[j] = Tr<T>::f(q[j * 128]);
            }
            if (mode == W_RES1) {
#pragma unroll
              for (int j = 0; j < 16; ++j) q[j * 128] = Tr<T>::r(res[j] + Tr<T>::rr(__uint_as_float(r[j])));   // residual add after rounding Wx
            } else if (mode == W_AFFINE) {
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + bias_n;
              if (has_res) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += res[j];
              }
              if (act == RD_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
              } else if (act == RD_ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) q[j * 128] = Tr<T>::r(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) q[j * 128] = Tr<T>::r(__uint_as_float(r[j]));
            }
          };
          if (works) finish(r0, jb0);
        }
        fence_proxy_async_smem();
        PROF(5)
        epi_bar(2);
        PROF(6)
        if (e == 0) {
          tma_store_2d(&map_out, bufp, n0, m0 + c * CR);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (E == 4) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");     // stores <= g - E have read their buffers
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        PROF(7)
        if (++cb == NBUF) { cb = 0; cph ^= 1u; }
      }
    }
    if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef RD_WIDE_PROF
    if (blockIdx.x == 1 && (e == 0 || e == 37))
      printf("wide prof e=%d tiles=%d chunks=%d: tile-setup %lld tfull-wait %lld bar1 %lld res-wait %lld tmem-ld %lld compute %lld bar2 %lld store+issue %lld (cycles)\n",
             e, tl, g, pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[6], pf[7]);
#endif
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();      // nobody leaves while the peer may still read this CTA's operands or signal its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int sm_count() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

// token-tile width.  Measured on B200 (tools/bench_wide.py): with a 4..6-stage ring a tile's time barely depends on its width
// between 160 and 256 tokens (the per-k-block TMA round trip, not the MMA, paces the narrower tiles), so big problems take 256.
// Mid-size problems (Q-Former: 1024 query tokens x 768 features = 24 tiles of 256 tokens for 148 SMs) take the widest of
// 256 / 128 / 64 that still yields ~100 tiles (qf dense 36.8 -> 14.3 us, fc2 42 -> 24 us against the one-tile-per-CTA kernel).
int choose_nt(int M, int w_tiles) {
  const int m16 = (M + 15) / 16 * 16;
  for (int nt : {256, 128}) {
    if ((long long)((M + nt - 1) / nt) * w_tiles >= 120) return nt < m16 ? nt : m16;
  }
  return 64 < m16 ? 64 : m16;
}

// implicit-GEMM convolution: geometry of the NHWC activation the token operand is gathered from (nullptr = plain GEMM)
struct ConvGeom { int B, H, W, C, ks, stride, pad, OH, OW; };

typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                     const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC tensor map in im2col mode (rank 4: C, W, H, N): the bounding box of window origins is [-pad, dim + pad - ks] per spatial
// dimension, windows advance by `stride`, a load brings `pixels` consecutive output pixels (w fastest, then h, then n) x 64
// channels into a 128-byte-swizzled [pixels][64] tile - the same shared-memory image as a box of a materialised im2col matrix.
struct Im2colKey { const void* ptr; ConvGeom g; int pixels, dtype; };
struct Im2colSlot { Im2colKey k; CUtensorMap m; };

int make_im2col_map(CUtensorMap* map, const void* x, const ConvGeom& g, int pixels, int dtype) {
  static thread_local std::vector<Im2colSlot>* cache = nullptr;
  if (cache == nullptr) cache = new std::vector<Im2colSlot>();
  for (const Im2colSlot& sl : *cache) {
    const Im2colKey& k = sl.k;
    if (k.ptr == x && k.pixels == pixels && k.dtype == dtype && k.g.B == g.B && k.g.H == g.H && k.g.W == g.W && k.g.C == g.C && k.g.ks == g.ks &&
        k.g.stride == g.stride && k.g.pad == g.pad) {
      *map = sl.m;
      return RD_OK;
    }
  }
  static PFN_encodeIm2col enc = nullptr;
  if (enc == nullptr) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = (PFN_encodeIm2col)fp;
  }
  RD_REQUIRE(enc != nullptr, "cuTensorMapEncodeIm2col entry point not available");
  cuuint64_t gdim[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t gstr[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.W * g.C * 2, (cuuint64_t)g.H * g.W * g.C * 2};
  int lower[2] = {-g.pad, -g.pad};
  int upper[2] = {g.pad - (g.ks - 1), g.pad - (g.ks - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  Im2colSlot sl;
  CUresult r = enc(&sl.m, dtype == RD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr,
                   lower, upper, (cuuint32_t)BK, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d) B=%d H=%d W=%d C=%d ks=%d stride=%d pixels=%d", (int)r, g.B, g.H, g.W, g.C, g.ks,
             g.stride, pixels);
  // drivers up to CUDA 13.1 set a descriptor bit that is wrong for tensors under 128 KB (same fix-up as CUTLASS's
  // copy_traits_sm90_im2col.hpp applies)
  {
    int drv = 0;
    if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (int64_t)g.B * g.H * g.W * g.C * 2 < 131072)
      reinterpret_cast<uint64_t*>(&sl.m)[1] &= ~(1llu << 21);
  }
  sl.k = Im2colKey{x, g, pixels, dtype};
  if (cache->size() < 256) cache->push_back(sl);
  *map = sl.m;
  return RD_OK;
}

template <class T, bool SWIGLU, bool PAIR>
int launch_wide(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                const EpiParams& epi, int mode, int dtype, cudaStream_t st, int nt, const ConvGeom* conv = nullptr) {
  constexpr int FEATS = SWIGLU ? 64 : 128;
  const int sms = sm_count();
  WideParams p{};
  p.M = M; p.N = N; p.K = K;
  constexpr int TILE_FEATS = PAIR ? 2 * FEATS : FEATS;       // a CTA pair covers two weight tiles
  constexpr int MAXS = PAIR ? MAX_STAGES_PAIR : MAX_STAGES;
  constexpr int SB = PAIR ? PAIR_STAGE_BYTES : STAGE_BYTES;
  const int units = PAIR ? sms / 2 : sms;
  p.n_tiles = (N + TILE_FEATS - 1) / TILE_FEATS;
  p.NT = nt;
  p.CR = (p.NT % 32 == 0) ? 32 : 16;
  p.m_tiles = (M + p.NT - 1) / p.NT;
  const long long tiles = (long long)p.m_tiles * p.n_tiles;
  // (a convolution is taken at any size: even one tile per CTA saves the im2col round trip through HBM)
  if ((conv == nullptr && tiles * (PAIR ? 2 : 1) < g_wide_min_tiles) || tiles > 0x3fffffff) return 0;
  p.tiles = (int)tiles;
  p.m_fast = ((int64_t)(SWIGLU ? 2 : 1) * N > (int64_t)M) ? 1 : 0;
  p.mode = mode; p.act = epi.act; p.has_res = (!SWIGLU && epi.residual != nullptr) ? 1 : 0; p.bias = epi.bias;
  {
    // 224 KB of shared memory: deep-K tiles want pipeline stages, K <= 128 tiles (one or two k-blocks, all epilogue) want the
    // residual prefetch to run far ahead instead
    const int kb = (K + BK - 1) / BK;
    p.stages = g_wide_force_stages > 0 ? g_wide_force_stages : (kb <= 2 ? 2 : kb <= 4 ? 3 : MAXS);
    if (p.stages > MAXS) p.stages = MAXS;
    p.nbuf = 4 + (MAXS - p.stages) * (SB / BUF_BYTES);
    if (p.nbuf > MAX_NBUF) p.nbuf = MAX_NBUF;
  }
  CUtensorMap map_w, map_x, map_out, map_res;
  RD_CHECK(rd_tc_make_map(&map_w, w, ldw, SWIGLU ? 2 * N : N, K, FEATS, BK, dtype, 1));
  if (conv != nullptr) {
    p.conv_ks = conv->ks; p.conv_C = conv->C; p.conv_OW = conv->OW; p.conv_OHW = conv->OH * conv->OW; p.conv_stride = conv->stride; p.conv_pad = conv->pad;
    RD_CHECK(make_im2col_map(&map_x, x, *conv, PAIR ? p.NT / 2 : p.NT, dtype));
  } else {
    RD_CHECK(rd_tc_make_map(&map_x, x, ldx, M, K, PAIR ? p.NT / 2 : p.NT, BK, dtype, 1));
  }
  RD_CHECK(rd_tc_make_map(&map_out, out, ldo, M, N, p.CR, FEATS, dtype, 0));
  if (p.has_res) RD_CHECK(rd_tc_make_map(&map_res, epi.residual, epi.ld_res, M, N, p.CR, FEATS, dtype, 0));
  else map_res = map_out;
  RD_SMEM_ATTR_ONCE(SMEM_BYTES, linear_wide_kernel<T, SWIGLU, PAIR>);
  cudaLaunchConfig_t cfg{};
  const unsigned walkers = (unsigned)(p.tiles < units ? p.tiles : units);
  cfg.gridDim = dim3(PAIR ? 2 * walkers : walkers); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (rd_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  RD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, linear_wide_kernel<T, SWIGLU, PAIR>, map_w, map_x, map_out, map_res, p));
  return 1;
}

static int wide_dispatch(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                         const EpiParams& epi, int dtype, cudaStream_t st, const ConvGeom* conv) {
  if (!g_wide_persistent || M <= 128) return 0;
  const bool sw = epi.act == RD_ACT_SWIGLU;
  const bool simple = epi.bias == nullptr && (epi.act == RD_ACT_NONE || sw);
  int mode;
  if (epi.lora_r != 0) return 0;
  if (simple && epi.residual == nullptr) mode = W_PLAIN;
  else if (!sw && simple && epi.residual != nullptr && epi.res_mode == 1) mode = W_RES1;
  else if (!sw && (epi.residual == nullptr || epi.res_mode == 2)) mode = W_AFFINE;
  else return 0;
  // TMA store / residual load: 16-byte aligned base and row pitch
  if (((uintptr_t)out & 15) != 0 || ldo % 8 != 0) return 0;
  if (epi.residual != nullptr && (((uintptr_t)epi.residual & 15) != 0 || epi.ld_res % 8 != 0)) return 0;
  // CTA pairs need a token tile that splits into two swizzle-aligned halves
  const int w_tiles = (N + (sw ? 64 : 128) - 1) / (sw ? 64 : 128);
  const int nt = g_wide_force_nt > 0 ? g_wide_force_nt : choose_nt(M, w_tiles);
  // ... and pay off when the MMAs dominate: deep K, and enough weight tiles that the second CTA of a pair is not idle
  const bool pair = g_wide_pair && nt % 16 == 0 && K >= 8 * BK && w_tiles >= 2 && (w_tiles % 2 == 0 || w_tiles >= 8);
  RD_DISPATCH_DTYPE(dtype, T, {
    if (pair) {
      if (sw) return launch_wide<T, true, true>(x, ldx, w, ldw, out, ldo, M, N, K, epi, mode, dtype, st, nt, conv);
      return launch_wide<T, false, true>(x, ldx, w, ldw, out, ldo, M, N, K, epi, mode, dtype, st, nt, conv);
    }
    if (sw) return launch_wide<T, true, false>(x, ldx, w, ldw, out, ldo, M, N, K, epi, mode, dtype, st, nt, conv);
    return launch_wide<T, false, false>(x, ldx, w, ldw, out, ldo, M, N, K, epi, mode, dtype, st, nt, conv);
  });
}

}  // namespace

// 1: launched, 0: shape / epilogue not handled here (caller falls through to linear_tc_kernel), < 0: error
int rd_linear_wide_try(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                       const EpiParams& epi, int dtype, cudaStream_t st) {
  return wide_dispatch(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, st, nullptr);
}

static int g_conv_implicit = 1;     // test hook: 0 = convolutions always go through the explicit im2col matrix
extern "C" int rd_conv_set_implicit(int on) { g_conv_implicit = on; return RD_OK; }

// Convolution as an implicit GEMM: out[(b,oh,ow), n] = epilogue( sum_{kh,kw,c} x[b, oh*stride-pad+kh, ow*stride-pad+kw, c] . w[n, (kh*ks+kw)*C + c] ),
// NHWC activations, weights [Cout, ks*ks*C] (the layout rd_im2col_nhwc + rd_linear use; eval-BatchNorm folded into w / bias).
// Returns 1 if launched, 0 if this shape has to take the explicit path (C not a multiple of 64, tiny M, unaligned output), < 0 on error.
extern "C" int rd_conv_nhwc_implicit(const void* x_dev, const void* w_dev, void* out_dev, int64_t ldo, int B, int H, int W, int C, int Cout,
                                     int ks, int stride, int pad, const rd_epilogue* epi, int dtype, void* stream) {
  if (!g_conv_implicit) return 0;
  if (C % BK != 0 || ks < 1 || ks > 7 || stride < 1 || pad < 0 || pad > 8 || ((uintptr_t)x_dev & 15) != 0) return 0;
  ConvGeom g{B, H, W, C, ks, stride, pad, (H + 2 * pad - ks) / stride + 1, (W + 2 * pad - ks) / stride + 1};
  if (g.OH < 1 || g.OW < 1) return 0;
  const int64_t M = (int64_t)B * g.OH * g.OW;
  if (M > 0x7fffffff) return 0;
  EpiParams e{};
  if (epi != nullptr) {
    e.bias = epi->bias_dev; e.residual = epi->residual_dev; e.ld_res = epi->ld_res; e.res_mode = epi->res_mode; e.act = epi->act;
    e.lora_t = epi->lora_t_dev; e.lora_b = epi->lora_b_dev; e.lora_r = epi->lora_r; e.lora_scale = epi->lora_scale;
  }
  const int K = ks * ks * C;
  return wide_dispatch(x_dev, 0, w_dev, K, out_dev, ldo, (int)M, Cout, K, e, dtype, (cudaStream_t)stream, &g);
}
