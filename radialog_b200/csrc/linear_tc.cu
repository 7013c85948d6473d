// tcgen05 tensor-core path of rd_linear (sm_100a).
//
//   out[M,N] = epilogue( x[M,K] . W[N,K]^T )        both operands K-major (row-major), fp16/bf16, fp32 accumulate
//
// "Swap-AB" tiling: the WEIGHT tile is the UMMA A operand (128 rows of W fill the 128 TMEM lanes) and the
// ACTIVATION tile is the UMMA B operand (NT token rows -> NT accumulator columns, NT in {16..256}).  That keeps
// the tensor core fed at any token count: a 32-row decode batch streams every weight byte exactly once through
// TMA -> swizzled smem -> tcgen05.mma with no padding of the token dimension to 128, which is what makes the
// decode step HBM-bound rather than tile-quantisation-bound; the same kernel at NT=256 is the prefill / Q-Former /
// conv GEMM.
//
// One CTA = one 128-row weight tile x one NT-token tile x one K split.
//   warp 0   : TMA producer  (cp.async.bulk.tensor 2D, 128B swizzle, mbarrier complete_tx, L2 cache hints)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer, tcgen05.commit -> mbarriers
//   warps 2-5: epilogue: tcgen05.ld accumulators -> registers -> fused epilogue (bias/act/SwiGLU/LoRA/residual)
// Split-K (decode: few weight tiles, many SMs) writes fp32 partials to a workspace; the last CTA of a tile reduces
// them in fixed split order, so results are deterministic run to run.
// PDL: weight tiles never depend on the previous kernel, so the producer issues the first pipeline stages of W
// before griddepcontrol.wait and only then loads the activations.
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

// Tuning knobs of the decode tiles (compile-time; defaults = the measured best)
#ifndef RD_DECODE_SMEM_KB
#define RD_DECODE_SMEM_KB 0            // 0: 100 KB (108 KB gate|up) of pipeline smem per CTA -> two CTAs per SM
#endif
#ifndef RD_DECODE_MIN_CTAS
#define RD_DECODE_MIN_CTAS 2           // __launch_bounds__ minimum CTAs per SM of the decode tiles (register cap)
#endif
#ifndef RD_TS_TMEM_COLS
#define RD_TS_TMEM_COLS 256
#endif
#ifndef RD_TS_W_KB
#define RD_TS_W_KB 64
#endif
#ifndef RD_TS_XST
#define RD_TS_XST 8
#endif
#ifndef RD_TS_DEFAULT
#define RD_TS_DEFAULT 0
#endif


bool rd_pdl_enabled();
unsigned long long* rd_linear_trace_buffer();
static int g_force_generic_epilogue = 0;   // test hook
extern "C" int rd_linear_force_generic_epilogue(int on) { g_force_generic_epilogue = on; return RD_OK; }
static int g_wide_epi = 1;        // test hook: 0 = direct (per-thread strided) epilogue also for wide token tiles
extern "C" int rd_linear_wide_epilogue(int on) { g_wide_epi = on; return RD_OK; }
static int g_ts_mode = RD_TS_DEFAULT;   // 1: decode tiles (M <= 32) park weight k-blocks in TMEM (A operand from tensor memory)
extern "C" int rd_linear_tmem_staging(int on) { g_ts_mode = on; return RD_OK; }
static int g_splitk_mode = 0;     // 0: cluster/DSMEM reduction when possible, 1: always the global workspace
extern "C" int rd_linear_splitk_mode(int mode) { g_splitk_mode = mode; return RD_OK; }

namespace {

// The fused epilogue is called from several fully unrolled accumulator loops; inlining it there makes a ~24k-instruction
// kernel whose straight-line epilogue runs at instruction-fetch latency (measured: ~250 ns per output element).  One
// out-of-line copy keeps the epilogue resident in the instruction cache.
template <class T>
__device__ __noinline__ T epilogue_call(const EpiParams& p, float acc, float acc_up, int m, int n, const float* lora_b_row) {
  return epilogue_elem<T>(p, acc, acc_up, m, n, lora_b_row);
}

constexpr int BLOCK_N = 128;   // weight rows per tile  (UMMA M)
constexpr int BLOCK_K = 64;    // 64 x 2 B = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;
constexpr uint64_t HINT_EVICT_NORMAL = 0x1000000000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must surface as a launch failure with a message, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  long long t0 = 0;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && (++spins & 0x3FFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) {      // ~2 s at 2 GHz
        printf("linear_tc_kernel: mbarrier wait timed out (tag %d, block %d,%d,%d, thread %d, parity %u)\n", tag, blockIdx.x,
               blockIdx.y, blockIdx.z, threadIdx.x, parity);
        __trap();
      }
    }
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// read a float from the same smem offset of CTA `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  uint32_t raddr;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr));
  return v;
}
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int slot) {
  if (trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    trace[(size_t)cta * 16 + slot] = t;
  }
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one lane of a fully converged warp (elect.sync).  The producer and MMA warps run their loops warp-uniformly and elect a
// lane only for the asynchronous instruction: with a lane-0-only loop ptxas wraps every UTMALDG / UTCHMMA operand in an
// ELECT / R2UR.BROADCAST / BRA waterfall and the serial instruction stream (~0.7 us per k-block, measured with per-role
// clock64 accounting in the persistent decode kernel) - not HBM - paces a small-N tile.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): K-major operand, 128-byte swizzle, rows of 128 B,
// 8-row groups 1024 B apart.  [0,14) addr>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4,
// [46,48) version=1 (sm_100), [61,64) layout=2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=f32, A/B = fmt (0 f16, 1 bf16), both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int umma_m, int umma_n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}


// ------------------------------------------------------------------------------------------------
// Epilogue specialisation.  The generic fused epilogue (epilogue_elem) re-reads its kernel parameters through the
// uniform datapath for every element, which measured ~250 ns per output element on the decode shapes; the hot
// decode/prefill cases therefore get branch-free paths whose operands are hoisted into registers once per thread.
// ------------------------------------------------------------------------------------------------
enum { EPI_GENERIC = 0, EPI_PLAIN = 1, EPI_RES1 = 2, EPI_LORA16 = 3, EPI_AFFINE = 4 };

template <class T> struct EpiCtx {
  T* outp;              // out + n
  int64_t ldo;
  const T* resp;        // residual + n            (EPI_RES1)
  int64_t ld_res;
  const T* lora_t;      // [M,16]                  (EPI_LORA16)
  float lora_scale;
  float b[16];          // lora_B[n, 0..16)        (EPI_LORA16)
  float bias_n;         // bias[n] (0 if none)      (EPI_AFFINE)
  int act;              // RD_ACT_* held in a real register (EPI_AFFINE)
  const T* res2p;       // residual + n for res_mode 2, or nullptr (EPI_AFFINE)
  const T* res_s;       // staged residual tile [NT][128] (nullptr = read global)
  const T* lt_s;        // staged lora_t tile [NT][16]    (nullptr = read global)
  int n_local;
};

template <class T, bool SWIGLU, int MODE>
__device__ __forceinline__ float finish_store(const EpiParams& ep, const EpiCtx<T>& cx, float acc, float accu, int m, int n, int j,
                                              bool has_res = false, float res_pre = 0.f) {
  T y;
  if (MODE == EPI_PLAIN) {
    if (SWIGLU) {
      const float g = Tr<T>::rr(acc), u = Tr<T>::rr(accu);
      y = Tr<T>::r(Tr<T>::rr(silu_f(g)) * u);                           // T(T(silu(T(g))) * T(u))
    } else {
      y = Tr<T>::r(acc);
    }
  } else if (MODE == EPI_RES1) {
    const float res = has_res ? res_pre : (cx.res_s ? Tr<T>::f(cx.res_s[j * BLOCK_N + cx.n_local]) : Tr<T>::f(cx.resp[(int64_t)m * cx.ld_res]));
    y = Tr<T>::r(res + Tr<T>::rr(acc));                                  // fp16 residual add after rounding Wx
  } else if (MODE == EPI_LORA16) {
    const T* trow = cx.lt_s ? cx.lt_s + j * 16 : cx.lora_t + (int64_t)m * 16;
    const Vec8<T> t0 = ld16(trow), t1 = ld16(trow + 8);
    float sdot = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) sdot += Tr<T>::f(t0.v[r]) * cx.b[r];
#pragma unroll
    for (int r = 0; r < 8; ++r) sdot += Tr<T>::f(t1.v[r]) * cx.b[8 + r];
    y = Tr<T>::r(Tr<T>::rr(acc) + Tr<T>::rr(cx.lora_scale * Tr<T>::rr(sdot)));
  } else if (MODE == EPI_AFFINE) {
    // bias / ReLU / GELU / fp32 residual (convs, Q-Former, img_proj): operands live in registers, no parameter re-reads
    float v = acc + cx.bias_n;
    if (has_res) v += res_pre; else if (cx.res2p != nullptr) v += Tr<T>::f(cx.res2p[(int64_t)m * cx.ld_res]);
    if (cx.act == RD_ACT_RELU) v = fmaxf(v, 0.0f);
    else if (cx.act == RD_ACT_GELU) v = gelu_erf(v);
    y = Tr<T>::r(v);
  } else {
    y = epilogue_elem<T>(ep, acc, accu, m, n, nullptr);
  }
  cx.outp[(int64_t)m * cx.ldo] = y;
  return Tr<T>::f(y);
}

#define EPI_DISPATCH(MODEVAR, ...)                                             \
  switch (MODEVAR) {                                                           \
    case EPI_PLAIN: { constexpr int MODE = EPI_PLAIN; __VA_ARGS__ } break;     \
    case EPI_RES1: { constexpr int MODE = EPI_RES1; __VA_ARGS__ } break;       \
    case EPI_LORA16: { constexpr int MODE = EPI_LORA16; __VA_ARGS__ } break;   \
    case EPI_AFFINE: { constexpr int MODE = EPI_AFFINE; __VA_ARGS__ } break;   \
    default: { constexpr int MODE = EPI_GENERIC; __VA_ARGS__ } break;          \
  }

constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int NT, bool SWIGLU> struct TcCfg {
  static constexpr int ACCS = SWIGLU ? 2 : 1;
  static constexpr int A_BYTES = BLOCK_N * BLOCK_K * 2;          // 16 KB per weight tile
  static constexpr int B_BYTES = NT * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = ACCS * A_BYTES + B_BYTES;
  // small-token (decode) tiles: keep two CTAs resident per SM so one CTA's prologue/epilogue hides behind the
  // other's weight stream; wide tiles take the whole SM.
  // decode tiles (NT <= 32, plain accumulator) stage the epilogue's residual / lora_t operands in smem while the
  // weight stream runs, so the tail after the last MMA does not wait on global loads
  static constexpr bool STAGE_EPI = (NT <= 32) && !SWIGLU;
  static constexpr int EXTRA_BYTES = STAGE_EPI ? (NT * BLOCK_N * 2 + NT * 16 * 2) : 0;     // residual tile + lora_t tile
  static constexpr int SMEM_BUDGET = (NT <= 32 && RD_DECODE_SMEM_KB > 0) ? RD_DECODE_SMEM_KB * 1024
                                     : (NT <= 64) ? (STAGE_EPI ? 100 * 1024 : 108 * 1024) : 200 * 1024;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = tmem_cols_for(ACCS * NT);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EXTRA_BYTES;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};

// TMEM-staged decode tiles (TS): the weight tile is the UMMA A operand, and tcgen05.mma can read A from tensor memory.  The
// epilogue warps - idle while the weights stream - copy every weight k-block that lands in shared memory into a ring of TMEM
// slots (thread = weight row, tcgen05.st 32x32b.x32) and free the shared-memory stage at once; the MMAs read A from TMEM and
// only the small token tile from shared memory.  A CTA then buffers its 64 KB W ring PLUS up to 112 KB of weights in TMEM, all
// of it fetched BEFORE the activations exist (weights never depend on the previous kernel): the dead time of a decode-step
// boundary (reduction tail of kernel N + launch + activation load of kernel N+1) is covered by ~1.8x more prefetched bytes.
template <int NT, bool SWIGLU> struct TsCfg {
  static constexpr int ACCS = SWIGLU ? 2 : 1;
  static constexpr int W_STAGE = ACCS * BLOCK_N * BLOCK_K * 2;            // 16 KB (32 KB gate|up) per weight k-block
  static constexpr int X_STAGE = NT * BLOCK_K * 2;
  static constexpr int WST = RD_TS_W_KB * 1024 / W_STAGE;                 // shared-memory W ring: 64 KB
  static constexpr int XST = RD_TS_XST;                                   // token-tile ring
  static constexpr int TMEM_COLS = RD_TS_TMEM_COLS;                       // two CTAs per SM share the 512 columns
  static constexpr int ACC_COLS = tmem_cols_for(ACCS * NT);
  static constexpr int SLOT_COLS = 32 * ACCS;                             // 64 K-elements of a 16-bit tile = 32 packed columns
  static constexpr int TSLOTS = (TMEM_COLS - ACC_COLS) / SLOT_COLS;       // 7 (3 for gate|up) weight k-blocks parked in TMEM
  static constexpr int RING_BYTES = WST * W_STAGE + XST * X_STAGE;
  static constexpr int BAR_BYTES = 512;
  static constexpr bool STAGE_EPI = (NT <= 32) && !SWIGLU;
  static constexpr int EXTRA_BYTES = STAGE_EPI ? (NT * BLOCK_N * 2 + NT * 16 * 2) : 0;
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 + BAR_BYTES + EXTRA_BYTES;
  static constexpr bool OK = TSLOTS >= 2 && TSLOTS <= 8 && WST >= 2 && WST <= 8 && XST >= 2 && XST <= 8;
};

struct TcParams {
  int M, N, K;
  int64_t ldo;
  int splits;
  int epi_mode;          // EPI_*: which specialised epilogue applies
  int cluster;           // 1: the splits of a tile form a thread-block cluster (1,1,splits), reduction over DSMEM
  float* ws_part;        // [splits][tiles][ACCS][NT][128] fp32
  uint32_t* ws_ctr;      // [tiles]
  uint64_t hint_w, hint_x;
  unsigned long long* trace;   // development aid: [cta][8] globaltimer stamps (nullptr = off)
  // "partials out" split-K (decode QKV GEMM, m_tiles == 1, plain epilogue): every CTA stores its fp32 partial tile to the slab
  // part_out[split][NT][N] and is done - no cluster, no DSMEM, no reduction pass.  The consumer (attention_decode_kernel) sums
  // the splits in fixed order and applies the single rounding T(Wx) when it reads q/k/v.
  float* part_out;
  int m_fast;                  // grid is (m_tiles, n_tiles, splits) instead of (n_tiles, m_tiles, splits)
  int stages;                  // smem pipeline stages of this launch (2 .. TcCfg::STAGES)
  int wide_epi;                // NT >= 64: transposed epilogue through shared memory (coalesced residual loads / stores)
  EpiParams epi;
};

__device__ __forceinline__ void tc_mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class T, int NT, bool SWIGLU, bool TS = false>
__global__ void __launch_bounds__(TC_THREADS, (NT <= 32) ? RD_DECODE_MIN_CTAS : (NT <= 64) ? 2 : 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, T* __restrict__ out,
                 const TcParams p) {
  using Cfg = TcCfg<NT, SWIGLU>;
  using Ts = TsCfg<(NT <= 32 ? NT : 32), SWIGLU>;
  // pipeline depth is a launch parameter (<= Cfg::STAGES): a GEMM with one or two k-blocks per CTA (the layer-1 convolutions:
  // K = 64) asks for one or two stages' worth of shared memory, so several of its latency-bound CTAs fit on an SM
  const int STAGES = TS ? 1 : p.stages;
  const int RING_BYTES = TS ? Ts::RING_BYTES : STAGES * Cfg::STAGE_BYTES;
  constexpr int BAR_BYTES = TS ? Ts::BAR_BYTES : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint32_t* flag_smem = tmem_ptr_smem + 1;
  // TS barriers (same area, own layout): W ring full/empty, token ring full/empty, TMEM slot ready/empty, accumulators, misc
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + RING_BYTES);
  uint64_t* w_empty = w_full + 8;
  uint64_t* x_full = w_empty + 8;
  uint64_t* x_empty = x_full + 8;
  uint64_t* a_ready = x_empty + 8;
  uint64_t* a_empty = a_ready + 8;
  if (TS) {
    tmem_full_bar = a_empty + 8;
    tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    flag_smem = tmem_ptr_smem + 1;
  }
  volatile uint32_t* dep_flag = flag_smem + 1;                       // TS: set once the producing kernel's output may be read

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // raster order: blockIdx.x runs fastest in the hardware's CTA dispatch.  m_fast puts the token tiles there, so that CTAs
  // running together share one WEIGHT tile and walk the (small, L2-resident) activations - the weights then cross HBM once
  // instead of once per token tile (prefill: 100-180 MB of weights cycled 8 times through a 126 MB L2 miss almost always)
  const int n_tile = p.m_fast ? blockIdx.y : blockIdx.x, m_tile = p.m_fast ? blockIdx.x : blockIdx.y, split = blockIdx.z;
  const int n_tiles_g = p.m_fast ? gridDim.y : gridDim.x;
  const int n0 = n_tile * BLOCK_N, m0 = m_tile * NT;
  const int kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = (int)(((int64_t)kb_total * split) / p.splits);
  const int kb_end = (int)(((int64_t)kb_total * (split + 1)) / p.splits);
  const int nkb = kb_end - kb_begin;

  pdl_launch_dependents();
  if (threadIdx.x == 0) trace_stamp(p.trace, 0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    if (TS) {
      for (int s = 0; s < 8; ++s) {
        mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 4); mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1);
        mbar_init(&a_ready[s], 4); mbar_init(&a_empty[s], 1);
      }
      *dep_flag = 0u;
    } else {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TS ? Ts::TMEM_COLS : Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) trace_stamp(p.trace, 1);
  const int quad = warp & 3;                      // TMEM lane quadrant an epilogue warp may access
  const int n_local = quad * 32 + lane;
  const int n = n0 + n_local;
  const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
  const int m_valid = min(NT, p.M - m0);
  const int tiles = gridDim.x * gridDim.y;
  const int tile_id = m_tile * n_tiles_g + n_tile;
  float* red = reinterpret_cast<float*>(smem);    // cluster split-K: partial tile parked in the pipeline smem
  EpiCtx<T> cx;                                   // epilogue operands hoisted into registers (epilogue warps only)

  if (TS && warp == 0) {
    // ===================== TS: weight producer - never waits for the previous kernel =====================
    for (int i = 0; i < nkb; ++i) {
      const int s = i % Ts::WST;
      if (i >= Ts::WST) mbar_wait(&w_empty[s], (uint32_t)((i / Ts::WST) & 1) ^ 1u, 1);
      __syncwarp();
      if (elect_one()) {
        uint8_t* sp = smem + s * Ts::W_STAGE;
        mbar_expect_tx(&w_full[s], Ts::W_STAGE);
        tma_load_2d(sp, &map_w, &w_full[s], (kb_begin + i) * BLOCK_K, n0, p.hint_w);
        if (SWIGLU) tma_load_2d(sp + Cfg::A_BYTES, &map_w, &w_full[s], (kb_begin + i) * BLOCK_K, p.N + n0, p.hint_w);
      }
      __syncwarp();
    }
  } else if (TS && warp == 1) {
    // ===================== TS: token-tile loader + MMA issuer (A from TMEM, B from shared memory) =====================
    constexpr uint32_t idesc = make_idesc(Tr<T>::umma_fmt, BLOCK_N, NT);
    uint8_t* xring = smem + Ts::WST * Ts::W_STAGE;
    auto load_x = [&](int j) {
      const int sx = j % Ts::XST;
      mbar_expect_tx(&x_full[sx], Ts::X_STAGE);
      tma_load_2d(xring + sx * Ts::X_STAGE, &map_x, &x_full[sx], (kb_begin + j) * BLOCK_K, m0, p.hint_x);
    };
    pdl_wait();                                  // activations were written by the previous kernel
    if (lane == 0) *dep_flag = 1u;
    if (elect_one()) {
      trace_stamp(p.trace, 2);
      for (int j = 0; j < nkb && j < Ts::XST; ++j) load_x(j);
    }
    __syncwarp();
    const uint64_t dx0 = make_smem_desc(smem_u32(xring));
    for (int i = 0; i < nkb; ++i) {
      const int t = i % Ts::TSLOTS, sx = i % Ts::XST;
      mbar_wait(&a_ready[t], (uint32_t)((i / Ts::TSLOTS) & 1), 2);
      mbar_wait(&x_full[sx], (uint32_t)((i / Ts::XST) & 1), 2);
      tc_fence_after();
      __syncwarp();
      if (elect_one()) {
        if (i == 0) trace_stamp(p.trace, 3);
        const uint64_t db = dx0 + (uint64_t)((uint32_t)sx * (Ts::X_STAGE >> 4));
        const uint32_t ta = tmem_base + (uint32_t)(Ts::ACC_COLS + t * Ts::SLOT_COLS);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
          const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
          tc_mma_ts_f16(tmem_base, ta + (uint32_t)(k * 8), db + koff, idesc, acc);
          if (SWIGLU) tc_mma_ts_f16(tmem_base + NT, ta + 32u + (uint32_t)(k * 8), db + koff, idesc, acc);
        }
        tc_commit(&a_empty[t]);                   // the TMEM slot and the token stage are free once these MMAs have read them
        tc_commit(&x_empty[sx]);
        if (i == nkb - 1) {
          tc_commit(tmem_full_bar);
          trace_stamp(p.trace, 4);
        }
      }
      __syncwarp();
      // refill the token stage released one iteration ago (its MMAs have had a k-block's time to complete)
      if (i >= 1 && i - 1 + Ts::XST < nkb) {
        const int j = i - 1;
        mbar_wait(&x_empty[j % Ts::XST], (uint32_t)((j / Ts::XST) & 1), 2);
        __syncwarp();
        if (elect_one()) load_x(j + Ts::XST);
        __syncwarp();
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    auto stage_ptr = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
    auto load_w = [&](int s, int kb) {
      uint8_t* sp = stage_ptr(s);
      tma_load_2d(sp, &map_w, &full_bar[s], kb * BLOCK_K, n0, p.hint_w);
      if (SWIGLU) tma_load_2d(sp + Cfg::A_BYTES, &map_w, &full_bar[s], kb * BLOCK_K, p.N + n0, p.hint_w);
    };
    auto load_x = [&](int s, int kb) {
      tma_load_2d(stage_ptr(s) + Cfg::ACCS * Cfg::A_BYTES, &map_x, &full_bar[s], kb * BLOCK_K, m0, p.hint_x);
    };
    const int pre = nkb < STAGES ? nkb : STAGES;
    if (elect_one()) {
      for (int i = 0; i < pre; ++i) {            // weights first: independent of the previous kernel
        mbar_expect_tx(&full_bar[i], Cfg::STAGE_BYTES);
        load_w(i, kb_begin + i);
      }
    }
    __syncwarp();
    pdl_wait();                                  // activations were written by the previous kernel
    if (elect_one()) {
      trace_stamp(p.trace, 2);
      for (int i = 0; i < pre; ++i) load_x(i, kb_begin + i);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int i = pre; i < nkb; ++i) {
      mbar_wait(&empty_bar[s], ph, 1);
      __syncwarp();
      if (elect_one()) {
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        load_w(s, kb_begin + i);
        load_x(s, kb_begin + i);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc(Tr<T>::umma_fmt, BLOCK_N, NT);
    const uint64_t d0 = make_smem_desc(smem_u32(smem));
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[s], ph, 2);
      tc_fence_after();
      const uint64_t da = d0 + (uint64_t)((uint32_t)s * (Cfg::STAGE_BYTES >> 4));
      const uint64_t du = da + (uint64_t)(Cfg::A_BYTES >> 4);
      const uint64_t db = da + (uint64_t)((Cfg::ACCS * Cfg::A_BYTES) >> 4);
      __syncwarp();
      if (elect_one()) {
        if (i == 0) trace_stamp(p.trace, 3);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
          const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);     // advance inside the 128-byte swizzle row
          tc_mma_f16(tmem_base, da + koff, db + koff, idesc, acc);
          if (SWIGLU) tc_mma_f16(tmem_base + NT, du + koff, db + koff, idesc, acc);
        }
        tc_commit(&empty_bar[s]);                 // frees the smem stage once these MMAs have read it
        if (i == nkb - 1) {
          tc_commit(tmem_full_bar);               // accumulators complete
          trace_stamp(p.trace, 4);
        }
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    bool res_staged = false;
    auto stage_residual = [&]() {
      // EPI_RES1: the residual values this thread will add, parked in shared memory (thread-private slots: no barrier needed)
      if (!(Cfg::STAGE_EPI && p.epi_mode == EPI_RES1 && n < p.N && (p.splits == 1 || p.cluster))) return;
      T* res_s = reinterpret_cast<T*>(smem + RING_BYTES + BAR_BYTES);
      const T* resp = reinterpret_cast<const T*>(p.epi.residual) + n;
      const int j0 = p.splits == 1 ? 0 : split, jstep = p.splits == 1 ? 1 : p.splits;
      for (int jb = j0; jb < m_valid; jb += 8 * jstep) {
        T tmp[8];
_Pragma("unroll")
        for (int t = 0; t < 8; ++t) { const int j = jb + t * jstep; if (j < m_valid) tmp[t] = resp[(int64_t)(m0 + j) * p.epi.ld_res]; }
_Pragma("unroll")
        for (int t = 0; t < 8; ++t) { const int j = jb + t * jstep; if (j < m_valid) res_s[j * BLOCK_N + n_local] = tmp[t]; }
      }
    };
    if (TS) {
      // ---- stagers: weight k-blocks shared memory -> TMEM (thread = weight row), W stage released immediately ----
      const int wrow = quad * 32 + lane;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Ts::WST, t = i % Ts::TSLOTS;
        mbar_wait(&w_full[s], (uint32_t)((i / Ts::WST) & 1), 4);
        if (i >= Ts::TSLOTS) mbar_wait(&a_empty[t], (uint32_t)((i / Ts::TSLOTS) & 1) ^ 1u, 4);
        tc_fence_after();
_Pragma("unroll")
        for (int a = 0; a < Cfg::ACCS; ++a) {
          const uint8_t* wt = smem + s * Ts::W_STAGE + a * Cfg::A_BYTES + (wrow >> 3) * 1024 + (wrow & 7) * 128;
          uint32_t r[32];
_Pragma("unroll")
          for (int c = 0; c < 8; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(wt + ((c ^ (wrow & 7)) << 4));
            r[c * 4 + 0] = v.x; r[c * 4 + 1] = v.y; r[c * 4 + 2] = v.z; r[c * 4 + 3] = v.w;
          }
          tc_st32(taddr + (uint32_t)(Ts::ACC_COLS + t * Ts::SLOT_COLS + a * 32), r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&a_ready[t]); mbar_arrive(&w_empty[s]); }
        if (!res_staged && *dep_flag != 0u) {      // the previous kernel is done: fetch the residual while the stream still runs
          pdl_wait();
          stage_residual();
          res_staged = true;
        }
        __syncwarp();
      }
    }
    pdl_wait();
    cx.outp = out + n; cx.ldo = p.ldo;
    cx.resp = reinterpret_cast<const T*>(p.epi.residual) + n; cx.ld_res = p.epi.ld_res;
    cx.lora_t = reinterpret_cast<const T*>(p.epi.lora_t); cx.lora_scale = p.epi.lora_scale;
    cx.bias_n = 0.f; cx.act = RD_ACT_NONE; cx.res2p = nullptr;
    if (p.epi_mode == EPI_AFFINE) {
      if (p.epi.bias != nullptr && n < p.N) cx.bias_n = p.epi.bias[n];
      // park act / residual pointer in ordinary registers (opaque moves) so the per-element code does not go back to the
      // constant bank through the uniform datapath
      int act_r = p.epi.act;
      asm volatile("mov.b32 %0, %0;" : "+r"(act_r));
      cx.act = act_r;
      unsigned long long rp = (p.epi.residual != nullptr && p.epi.res_mode == 2) ? (unsigned long long)(cx.resp) : 0ull;
      asm volatile("mov.b64 %0, %0;" : "+l"(rp));
      cx.res2p = reinterpret_cast<const T*>(rp);
    }
    if (p.epi_mode == EPI_LORA16 && n < p.N) {
      const T* brow = reinterpret_cast<const T*>(p.epi.lora_b) + (int64_t)n * 16;
      const Vec8<T> b0 = ld16(brow), b1 = ld16(brow + 8);
_Pragma("unroll")
      for (int r = 0; r < 8; ++r) { cx.b[r] = Tr<T>::f(b0.v[r]); cx.b[8 + r] = Tr<T>::f(b1.v[r]); }
    }
    cx.res_s = nullptr; cx.lt_s = nullptr; cx.n_local = n_local;
    if (Cfg::STAGE_EPI) {
      // while the weight stream runs: pull the residual values this thread will add and the lora_t rows of the tile
      T* res_s = reinterpret_cast<T*>(smem + RING_BYTES + BAR_BYTES);
      T* lt_s = res_s + NT * BLOCK_N;
      if (p.epi_mode == EPI_RES1 && n < p.N && (p.splits == 1 || p.cluster)) {
        if (!res_staged) stage_residual();
        cx.res_s = res_s;
      }
      if (p.epi_mode == EPI_LORA16) {
        const int e = threadIdx.x - 64;          // 128 epilogue threads, one 16-byte chunk each per round
        for (int i = e; i < m_valid * 2; i += 128)
          *reinterpret_cast<uint4*>(lt_s + i * 8) = *reinterpret_cast<const uint4*>(cx.lora_t + (int64_t)m0 * 16 + i * 8);
        cx.lt_s = lt_s;
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
    }
    mbar_wait(tmem_full_bar, 0, 3);
    tc_fence_after();
    if (threadIdx.x == 64) trace_stamp(p.trace, 5);
    if (NT >= 64 && p.wide_epi) {
      // ---- wide token tiles (prefill, convolutions): transposed epilogue.  A thread owns one weight row n, so the direct
      // ---- epilogue issues one 2-byte residual load and one 2-byte store per (token, thread) at a stride of ldo - with 256
      // ---- tokens per tile that, not the MMAs, set the pace (prefill o_proj tile: 17 us of MMA, ~90 us of epilogue).
      // ---- Phase A parks the fp32 values [token][n] in the (now idle) pipeline smem, phase B walks it token-major: each
      // ---- thread finishes 4 consecutive n of a token with one 8-byte residual load and one 8-byte store (a warp covers the
      // ---- tile's 128 columns of a token row: 256 contiguous bytes).  Same arithmetic and rounding points as the direct path.
      float* stg = reinterpret_cast<float*>(smem);
      const int mode = p.epi_mode;
      for (int c = 0; c < m_valid; c += 16) {
        uint32_t r[16], ru[16];
        tc_ld16(taddr + c, r);
        if (SWIGLU) tc_ld16(taddr + NT + c, ru);
        tc_wait_ld();
_Pragma("unroll")
        for (int j = 0; j < 16; ++j) {
          if (c + j < m_valid) {
            float v = __uint_as_float(r[j]);
            if (SWIGLU) {
              const float g = Tr<T>::rr(v), u = Tr<T>::rr(__uint_as_float(ru[j]));
              v = Tr<T>::rr(Tr<T>::rr(silu_f(g)) * u);                    // T(T(silu(T(g))) * T(u))
            } else if (mode == EPI_AFFINE) {
              v += cx.bias_n;
            }
            stg[(c + j) * BLOCK_N + n_local] = v;
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = threadIdx.x - 64;
      const int n4 = (e & 31) * 4, nn = n0 + n4;
      if (nn < p.N) {
        const bool has_res = (mode == EPI_RES1) || (mode == EPI_AFFINE && p.epi.residual != nullptr && p.epi.res_mode == 2);
        const T* resb = reinterpret_cast<const T*>(p.epi.residual) + nn;
        T* outb = out + nn;
        const int act = p.epi.act;
        constexpr int UNR = 4;
        for (int mb = e >> 5; mb < m_valid; mb += 4 * UNR) {
          float4 v[UNR];
          uint2 rv[UNR];
_Pragma("unroll")
          for (int u = 0; u < UNR; ++u) {
            const int m = mb + 4 * u;
            if (m < m_valid) {
              v[u] = *reinterpret_cast<const float4*>(stg + m * BLOCK_N + n4);
              if (has_res) rv[u] = *reinterpret_cast<const uint2*>(resb + (int64_t)(m0 + m) * p.epi.ld_res);
            }
          }
_Pragma("unroll")
          for (int u = 0; u < UNR; ++u) {
            const int m = mb + 4 * u;
            if (m < m_valid) {
              const float a[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
              const T* rt = reinterpret_cast<const T*>(&rv[u]);
              T y[4];
_Pragma("unroll")
              for (int q = 0; q < 4; ++q) {
                if (mode == EPI_RES1) {
                  y[q] = Tr<T>::r(Tr<T>::f(rt[q]) + Tr<T>::rr(a[q]));       // fp16 residual add after rounding Wx
                } else if (mode == EPI_AFFINE) {
                  float w = a[q];
                  if (has_res) w += Tr<T>::f(rt[q]);
                  if (act == RD_ACT_RELU) w = fmaxf(w, 0.0f);
                  else if (act == RD_ACT_GELU) w = gelu_erf(w);
                  y[q] = Tr<T>::r(w);
                } else {
                  y[q] = Tr<T>::r(a[q]);
                }
              }
              *reinterpret_cast<uint2*>(outb + (int64_t)(m0 + m) * p.ldo) = *reinterpret_cast<const uint2*>(y);
            }
          }
        }
      }
    } else if (NT <= 32 && !SWIGLU && p.part_out != nullptr) {
      // ---- partials out: publish the fp32 partial tile (a warp stores 128 contiguous bytes per token) and leave ----
      float* part = p.part_out + (int64_t)split * NT * p.N;
      for (int c = 0; c < m_valid; c += 16) {
        uint32_t r[16];
        tc_ld16(taddr + c, r);
        tc_wait_ld();
        if (n < p.N) {
_Pragma("unroll")
          for (int j = 0; j < 16; ++j)
            if (c + j < m_valid) __stcg(part + (int64_t)(c + j) * p.N + n, __uint_as_float(r[j]));
        }
      }
    } else if (p.splits == 1) {
      EPI_DISPATCH(p.epi_mode,
        for (int c = 0; c < m_valid; c += 16) {
          uint32_t r[16], ru[16];
          tc_ld16(taddr + c, r);
          if (SWIGLU) tc_ld16(taddr + NT + c, ru);
          tc_wait_ld();
          if (n < p.N) {
            // residual operands of the 16 columns are requested together, before the first one is consumed
            const T* rsrc = (MODE == EPI_RES1 && cx.res_s == nullptr) ? cx.resp : ((MODE == EPI_AFFINE) ? cx.res2p : nullptr);
            float resv[16];
            if ((MODE == EPI_RES1 || MODE == EPI_AFFINE) && rsrc != nullptr) {
_Pragma("unroll")
              for (int j = 0; j < 16; ++j) resv[j] = (c + j < m_valid) ? Tr<T>::f(rsrc[(int64_t)(m0 + c + j) * cx.ld_res]) : 0.f;
            }
_Pragma("unroll")
            for (int j = 0; j < 16; ++j) {
              if (c + j < m_valid)
                finish_store<T, SWIGLU, MODE>(p.epi, cx, __uint_as_float(r[j]), SWIGLU ? __uint_as_float(ru[j]) : 0.f, m0 + c + j, n, c + j,
                                                      (MODE == EPI_RES1 || MODE == EPI_AFFINE) && rsrc != nullptr, resv[j]);
            }
          }
        }
      )
    } else if (p.cluster) {
      // split-K inside a thread-block cluster: park the fp32 partial tile in this CTA's (now idle) pipeline smem;
      // after the cluster barrier every CTA reduces its own slice of the token columns over DSMEM.
      for (int c = 0; c < m_valid; c += 16) {
_Pragma("unroll")
        for (int a = 0; a < Cfg::ACCS; ++a) {
          uint32_t r[16];
          tc_ld16(taddr + a * NT + c, r);
          tc_wait_ld();
_Pragma("unroll")
          for (int j = 0; j < 16; ++j)
            if (c + j < m_valid) red[(a * NT + c + j) * BLOCK_N + n_local] = __uint_as_float(r[j]);
        }
      }
    } else {
      // split-K through a global workspace: publish the fp32 partial tile, the last CTA of the tile reduces all
      // splits in fixed order
      float* part = p.ws_part + ((int64_t)split * tiles + tile_id) * (Cfg::ACCS * NT * BLOCK_N);
      for (int c = 0; c < m_valid; c += 16) {
_Pragma("unroll")
        for (int a = 0; a < Cfg::ACCS; ++a) {
          uint32_t r[16];
          tc_ld16(taddr + a * NT + c, r);
          tc_wait_ld();
_Pragma("unroll")
          for (int j = 0; j < 16; ++j)
            if (c + j < m_valid) __stcg(part + ((int64_t)a * NT + c + j) * BLOCK_N + n_local, __uint_as_float(r[j]));
        }
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const uint32_t prev = atomicAdd(p.ws_ctr + tile_id, 1u);
        *flag_smem = (prev == (uint32_t)p.splits - 1) ? 1u : 0u;
        if (prev == (uint32_t)p.splits - 1) p.ws_ctr[tile_id] = 0;     // re-arm for the next launch
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*flag_smem) {
        __threadfence();
        if (n < p.N) {
          EPI_DISPATCH(p.epi_mode,
            for (int c = 0; c < m_valid; c += 8) {
              float acc[8], accu[8];
_Pragma("unroll")
              for (int j = 0; j < 8; ++j) { acc[j] = 0.f; accu[j] = 0.f; }
              for (int s = 0; s < p.splits; ++s) {       // 8 (16) independent loads in flight per split
                const float* ps = p.ws_part + ((int64_t)s * tiles + tile_id) * (Cfg::ACCS * NT * BLOCK_N);
_Pragma("unroll")
                for (int j = 0; j < 8; ++j) {
                  if (c + j < m_valid) {
                    acc[j] += __ldcg(ps + (int64_t)(c + j) * BLOCK_N + n_local);
                    if (SWIGLU) accu[j] += __ldcg(ps + ((int64_t)NT + c + j) * BLOCK_N + n_local);
                  }
                }
              }
_Pragma("unroll")
              for (int j = 0; j < 8; ++j)
                if (c + j < m_valid) finish_store<T, SWIGLU, MODE>(p.epi, cx, acc[j], accu[j], m0 + c + j, n, c + j);
            }
          )
        }
      }
    }
    tc_fence_before();
    if (threadIdx.x == 64) trace_stamp(p.trace, 11);
  }
  if (p.cluster) {
    // every thread of every CTA in the cluster: partial tiles are complete and visible cluster-wide
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 64) trace_stamp(p.trace, 6);
    if (warp >= 2 && n < p.N) {
      const uint32_t red_addr = smem_u32(red);
      // this CTA's slice of the token columns: j = split, split + splits, ...; RT columns per round so that their partials
      // are in flight together over DSMEM; summed in fixed split order.  Register budget: RT x SMAX partials (x2 with SwiGLU) -
      // with at most 4 splits (gate|up: 3) twice as many columns fit in a round, i.e. one DSMEM round trip less in the tail.
      auto reduce = [&](auto smax_c, auto rt_c) {
        constexpr int SMAX = decltype(smax_c)::value, RT = decltype(rt_c)::value;
        EPI_DISPATCH(p.epi_mode,
          for (int jb = split; jb < m_valid; jb += RT * p.splits) {
            float v[RT][SMAX], vu[SWIGLU ? RT : 1][SMAX];
_Pragma("unroll")
            for (int t = 0; t < RT; ++t) {
              const int j = jb + t * p.splits;
_Pragma("unroll")
              for (int s = 0; s < SMAX; ++s) {
                v[t][s] = 0.f;
                if (SWIGLU) vu[t][s] = 0.f;
                if (j < m_valid && s < p.splits) {
                  v[t][s] = ld_dsmem_f32(red_addr + (uint32_t)((j * BLOCK_N + n_local) * 4), (uint32_t)s);
                  if (SWIGLU) vu[t][s] = ld_dsmem_f32(red_addr + (uint32_t)(((NT + j) * BLOCK_N + n_local) * 4), (uint32_t)s);
                }
              }
            }
_Pragma("unroll")
            for (int t = 0; t < RT; ++t) {
              const int j = jb + t * p.splits;
              if (j < m_valid) {
                float acc = 0.f, accu = 0.f;
_Pragma("unroll")
                for (int s = 0; s < SMAX; ++s) { acc += v[t][s]; if (SWIGLU) accu += vu[t][s]; }
                finish_store<T, SWIGLU, MODE>(p.epi, cx, acc, accu, m0 + j, n, j);
              }
            }
          }
        )
      };
      if (p.splits <= 4) reduce(std::integral_constant<int, 4>{}, std::integral_constant<int, 8>{});
      else reduce(std::integral_constant<int, 8>{}, std::integral_constant<int, SWIGLU ? 4 : 8>{});
    }
    // nobody leaves (and frees its smem) while a peer may still be reading it
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 64) trace_stamp(p.trace, 7);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TS ? Ts::TMEM_COLS : Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
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

// cuTensorMapEncodeTiled costs a few microseconds of host time; a step re-uses the same few hundred (pointer, shape) pairs
// (weights, the engine's activation buffers), so the encoded descriptors are kept in a small open-addressing cache.
// Two kinds of 2-D map over a row-major [rows, cols] 16-bit matrix: swizzle128 = 1 is the K-major UMMA operand (box of
// 64 columns = one 128-byte swizzle row), swizzle128 = 0 a dense box (epilogue tiles: TMA store / residual load).
struct MapKey { const void* ptr; int64_t ld; int rows, cols, box_rows, box_cols, dtype, swz; };
struct MapSlot { MapKey k; CUtensorMap m; bool used; };
constexpr int MAP_CACHE = 4096;

int encode_map(CUtensorMap* map, const MapKey& k) {
  PFN_encodeTiled enc = get_encode();
  RD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)k.cols, (cuuint64_t)k.rows};
  cuuint64_t gstr[1] = {(cuuint64_t)k.ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)k.box_cols, (cuuint32_t)k.box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, k.dtype == RD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(k.ptr),
                   gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, k.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) ptr=%p ld=%lld rows=%d cols=%d box=%dx%d", (int)r, k.ptr, (long long)k.ld,
             k.rows, k.cols, k.box_rows, k.box_cols);
  return RD_OK;
}

int make_map_ex(CUtensorMap* map, const MapKey& key) {
  static thread_local MapSlot* cache = nullptr;
  if (cache == nullptr) cache = static_cast<MapSlot*>(calloc(MAP_CACHE, sizeof(MapSlot)));
  if (cache == nullptr) return encode_map(map, key);
  uint64_t hsh = (uint64_t)(uintptr_t)key.ptr * 0x9E3779B97F4A7C15ull ^ ((uint64_t)key.ld << 32) ^ ((uint64_t)key.rows << 17) ^ ((uint64_t)key.cols << 3) ^
                 (uint64_t)key.box_rows ^ ((uint64_t)key.box_cols << 9) ^ ((uint64_t)key.dtype << 60) ^ ((uint64_t)key.swz << 59);
  hsh ^= hsh >> 29;
  auto same = [&](const MapKey& a) {
    return a.ptr == key.ptr && a.ld == key.ld && a.rows == key.rows && a.cols == key.cols && a.box_rows == key.box_rows &&
           a.box_cols == key.box_cols && a.dtype == key.dtype && a.swz == key.swz;
  };
  for (int probe = 0; probe < 8; ++probe) {
    MapSlot& sl = cache[(hsh + probe) & (MAP_CACHE - 1)];
    if (sl.used && same(sl.k)) {
      *map = sl.m;
      return RD_OK;
    }
    if (!sl.used) {
      RD_CHECK(encode_map(&sl.m, key));
      sl.k = key;
      sl.used = true;
      *map = sl.m;
      return RD_OK;
    }
  }
  MapSlot& sl = cache[hsh & (MAP_CACHE - 1)];                         // neighbourhood full: overwrite the home slot
  RD_CHECK(encode_map(&sl.m, key));
  sl.k = key;
  sl.used = true;
  *map = sl.m;
  return RD_OK;
}

int make_map(CUtensorMap* map, const void* ptr, int64_t ld, int rows, int K, int box_rows, int dtype) {
  return make_map_ex(map, MapKey{ptr, ld, rows, K, box_rows, BLOCK_K, dtype, 1});
}

// token-tile width.  Many tokens but only one or two k-blocks (the 1x1 convolutions of ResNet layer1/2, K = 64 / 128): such a CTA
// is all prologue + epilogue, so it gets the 128-token tile whose parked epilogue tile (64 KB) lets two CTAs share an SM.
// Mid-size problems (Q-Former: 1024 query tokens x 768 features = 24 tiles of 256 tokens on 148 SMs) take the widest tile that
// still yields ~100 CTAs.
int pick_nt(int M, int N, int K) {
  if (M > 128 && K <= 2 * BLOCK_K) return 128;
  if (M <= 128) return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128;
  const long long n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  for (int nt : {256, 128}) {
    if (n_tiles * ((M + nt - 1) / nt) >= 96) return nt;
  }
  return 64;
}

// Max CTAs of the decode-tile kernel that are co-resident when launched as clusters of (1,1,cs).  Measured on B200
// (ncu launch__cluster_max_active with two ~110 KB CTAs per SM: 71 clusters of 4 = 284 CTAs; timing sweeps agree:
// 258 CTAs in clusters of 3 run as one wave, 288 do not).  cudaOccupancyMaxActiveClusters under-reports here (it
// answers as if one CTA fitted per SM), so the measured figure is used; RD_DEBUG_OCC=1 prints both.
template <class T, int NT, bool SWIGLU>
int resident_capacity(int cs) {
  using Cfg = TcCfg<NT, SWIGLU>;
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cs < 1 || cs > 8) return 1;
  if (cache[cs]) return cache[cs];
  const int per_sm = (227 * 1024) / (Cfg::SMEM_BYTES + 1024);
  const int model = per_sm >= 2 ? (284 / cs) * cs : (148 / cs) * cs;      // (3+ per SM by smem still counts as 2: registers / TMEM)
  if (getenv("RD_DEBUG_OCC")) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148, 1, cs); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)cs;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t oe = cudaOccupancyMaxActiveClusters(&n, linear_tc_kernel<T, NT, SWIGLU>, &cfg);
    cudaGetLastError();
    fprintf(stderr, "radialog_b200: NT=%d swiglu=%d cluster %d: occupancy API %d clusters (%s), model %d CTAs\n", NT, (int)SWIGLU, cs, n,
            cudaGetErrorString(oe), model);
  }
  cache[cs] = model;
  return model;
}

// Split-K factor for few-tile (decode) GEMMs.  The kernel is HBM-bound, so what matters is that (a) every CTA of the
// grid is co-resident in ONE wave (equal shares of the stream finish together; a cluster size whose clusters do not
// pack into the GPCs silently costs a second wave), (b) there are comfortably more CTAs than SMs pulling on HBM, and
// (c) the split count is as small as that allows (less reduction work).
template <class T, int NT, bool SWIGLU>
int choose_splits(int tiles, int kb) {
  if (NT > 64) return 1;                          // wide token tiles (prefill / conv): tensor-bound, many tiles
  const int want = 240;                           // ~1.6 CTAs per SM: measured sweet spot of the split sweep (profiles/)
  const int smax = kb / 4 > 0 ? (kb / 4 > 8 ? 8 : kb / 4) : 1;   // >= 4 k-blocks per CTA, portable cluster size <= 8
  int best = 1, best_ctas = tiles;
  for (int s = 1; s <= smax; ++s) {
    const int ctas = tiles * s;
    if (s > 1 && ctas > resident_capacity<T, NT, SWIGLU>(g_splitk_mode == 0 ? s : 1)) continue;
    if (ctas >= want) return s;
    if (ctas > best_ctas) { best = s; best_ctas = ctas; }
  }
  return best;
}

// set by rd_linear_tc_fused for the launch it wraps (single-threaded use per handle, like the rest of the library)
static const TcFuse* g_fuse = nullptr;

template <class T, int NT, bool SWIGLU>
int launch_tc(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
              const EpiParams& epi, int dtype, void* ws, int64_t ws_bytes, int splits, cudaStream_t st) {
  using Cfg = TcCfg<NT, SWIGLU>;
  RD_SMEM_ATTR_ONCE(Cfg::SMEM_BYTES, linear_tc_kernel<T, NT, SWIGLU>);
  CUtensorMap map_w, map_x;
  RD_CHECK(make_map(&map_w, w, ldw, SWIGLU ? 2 * N : N, K, BLOCK_N, dtype));
  RD_CHECK(make_map(&map_x, x, ldx, M, K, NT, dtype));
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N, m_tiles = (M + NT - 1) / NT;
  const int kb_total = (K + BLOCK_K - 1) / BLOCK_K;
  if (splits <= 0) splits = choose_splits<T, NT, SWIGLU>(n_tiles * m_tiles, kb_total);
  if (splits > kb_total) splits = kb_total;
  if (splits > 8 && (ws == nullptr || g_splitk_mode == 0)) splits = 8;
  if (splits > 16) splits = 16;
  if (splits > 1 && ws == nullptr && g_splitk_mode != 0) splits = 1;
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.ldo = ldo; p.splits = splits; p.epi = epi;
  p.trace = rd_linear_trace_buffer();
  {
    const bool simple = epi.bias == nullptr && (epi.act == RD_ACT_NONE || (SWIGLU && epi.act == RD_ACT_SWIGLU));
    p.epi_mode = EPI_GENERIC;
    if (simple && epi.residual == nullptr && epi.lora_r == 0) p.epi_mode = EPI_PLAIN;
    else if (!SWIGLU && simple && epi.residual != nullptr && epi.res_mode == 1 && epi.lora_r == 0) p.epi_mode = EPI_RES1;
    else if (!SWIGLU && simple && epi.residual == nullptr && epi.lora_r == 16) p.epi_mode = EPI_LORA16;
    else if (!SWIGLU && epi.lora_r == 0 && epi.act != RD_ACT_SWIGLU && (epi.residual == nullptr || epi.res_mode == 2)) p.epi_mode = EPI_AFFINE;
    if (g_force_generic_epilogue) p.epi_mode = EPI_GENERIC;
  }
  // decode: every weight byte is read once (evict-first), the small activation tile is shared by all CTAs (evict-last)
  p.hint_w = m_tiles == 1 ? HINT_EVICT_FIRST : HINT_EVICT_NORMAL;
  p.hint_x = m_tiles == 1 ? HINT_EVICT_LAST : HINT_EVICT_NORMAL;
  // split-K reduction: thread-block cluster + DSMEM when the splits fit a portable cluster (<= 8) and the partial tile fits the
  // pipeline smem; otherwise fp32 partials through the global workspace.
  const bool use_cluster = splits > 1 && splits <= 8 && g_splitk_mode == 0 &&
                           Cfg::ACCS * NT * BLOCK_N * 4 <= Cfg::STAGES * Cfg::STAGE_BYTES;
  p.cluster = use_cluster ? 1 : 0;
  {
    const bool res_ok = epi.residual == nullptr || (epi.ld_res % 4 == 0 && ((uintptr_t)epi.residual & 7) == 0);
    p.wide_epi = (NT >= 64 && splits == 1 && g_wide_epi && (p.epi_mode == EPI_PLAIN || p.epi_mode == EPI_RES1 || p.epi_mode == EPI_AFFINE) &&
                  N % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)out & 7) == 0 && res_ok &&
                  (int64_t)NT * BLOCK_N * 4 <= (int64_t)Cfg::STAGES * Cfg::STAGE_BYTES) ? 1 : 0;
  }
  if (g_fuse != nullptr) {
    const TcFuse& f = *g_fuse;
    RD_REQUIRE(NT <= 32 && m_tiles == 1, "rd_linear_tc_fused: needs M <= 32");
    if (f.part_out != nullptr) {
      const bool plain = !SWIGLU && p.epi_mode == EPI_PLAIN;
      const int64_t need = (int64_t)splits * NT * N * 4;
      if (!plain || f.part_bytes < need) {
        rd_set_error("rd_linear_tc_fused: partials-out needs a plain epilogue and a slab of %lld bytes (got %lld)", (long long)need, (long long)f.part_bytes);
        return RD_ERR_INVALID;
      }
      p.part_out = f.part_out;
      p.cluster = 0;
      if (f.splits_out != nullptr) { f.splits_out[0] = splits; f.splits_out[1] = NT; }
    }
  }
  const bool cluster_launch = use_cluster && p.part_out == nullptr;
  if (splits > 1 && !use_cluster && p.part_out == nullptr) {
    const int64_t part_bytes = (int64_t)splits * n_tiles * m_tiles * Cfg::ACCS * NT * BLOCK_N * 4;
    const int64_t need = part_bytes + ((int64_t)n_tiles * m_tiles * 4 + 255) / 256 * 256;
    RD_REQUIRE(ws != nullptr && ws_bytes >= need, "rd_linear: split-K workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
    p.ws_ctr = reinterpret_cast<uint32_t*>(ws);                                   // counters first (zero-initialised by the owner)
    p.ws_part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ((int64_t)n_tiles * m_tiles * 4 + 255) / 256 * 256);
  }
  // pipeline depth: no deeper than the k-blocks a CTA owns, no shallower than what the epilogue needs to park its tile in
  {
    const int kb_cta = (kb_total + splits - 1) / splits;
    int st_n = kb_cta < 2 ? 2 : kb_cta;
    int64_t park = 0;
    if (use_cluster) park = (int64_t)Cfg::ACCS * NT * BLOCK_N * 4;
    if (p.wide_epi) park = (int64_t)NT * BLOCK_N * 4;
    const int st_park = (int)((park + Cfg::STAGE_BYTES - 1) / Cfg::STAGE_BYTES);
    if (st_n < st_park) st_n = st_park;
    p.stages = st_n > Cfg::STAGES ? Cfg::STAGES : st_n;
  }
  const int smem_bytes = Cfg::SMEM_BYTES - (Cfg::STAGES - p.stages) * Cfg::STAGE_BYTES;
  // the operand that is larger in HBM should be the one consecutive CTAs share (see the kernel's raster-order note)
  p.m_fast = (m_tiles > 1 && splits == 1 && !p.cluster && (int64_t)(SWIGLU ? 2 : 1) * N > (int64_t)M) ? 1 : 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = p.m_fast ? dim3(m_tiles, n_tiles, splits) : dim3(n_tiles, m_tiles, splits); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (rd_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_launch) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = (unsigned)splits;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  if constexpr (NT <= 32) {
    // TMEM-staged variant: decode tiles with a plain / residual / SwiGLU epilogue (see TsCfg)
    const bool ts = TsCfg<NT, SWIGLU>::OK && g_ts_mode && m_tiles == 1 && (p.epi_mode == EPI_PLAIN || p.epi_mode == EPI_RES1) &&
                    kb_total / splits >= 2;
    if (ts) {
      using Ts = TsCfg<NT, SWIGLU>;
      RD_SMEM_ATTR_ONCE(Ts::SMEM_BYTES, linear_tc_kernel<T, NT, SWIGLU, true>);
      cfg.dynamicSmemBytes = Ts::SMEM_BYTES;
      RD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, linear_tc_kernel<T, NT, SWIGLU, true>, map_w, map_x, (T*)out, p));
      return RD_OK;
    }
  }
  RD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, linear_tc_kernel<T, NT, SWIGLU>, map_w, map_x, (T*)out, p));
  return RD_OK;
}

template <class T, bool SWIGLU>
int dispatch_nt(int nt, const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                const EpiParams& epi, int dtype, void* ws, int64_t ws_bytes, int splits, cudaStream_t st) {
  switch (nt) {
    case 16: return launch_tc<T, 16, SWIGLU>(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, splits, st);
    case 32: return launch_tc<T, 32, SWIGLU>(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, splits, st);
    case 64: return launch_tc<T, 64, SWIGLU>(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, splits, st);
    case 128: return launch_tc<T, 128, SWIGLU>(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, splits, st);
    default: return launch_tc<T, 256, SWIGLU>(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, splits, st);
  }
}

}  // namespace

// development aid: per-CTA globaltimer stamps; with a non-zero stride every launch gets its own slab of the buffer
static unsigned long long* g_trace = nullptr;
static long long g_trace_stride = 0, g_trace_launch = 0;
extern "C" int rd_linear_set_trace(void* buf) { g_trace = (unsigned long long*)buf; g_trace_stride = 0; g_trace_launch = 0; return RD_OK; }
extern "C" int rd_linear_set_trace_strided(void* buf, long long stride_u64) {
  g_trace = (unsigned long long*)buf; g_trace_stride = stride_u64; g_trace_launch = 0; return RD_OK;
}
extern "C" long long rd_linear_trace_launches() { return g_trace_launch; }
unsigned long long* rd_linear_trace_buffer() { return g_trace ? g_trace + (g_trace_launch++) * g_trace_stride : nullptr; }
static int g_force_splits = 0;   // test hook: 0 = heuristic
extern "C" int rd_linear_force_splits(int s) { g_force_splits = s; return RD_OK; }

int64_t rd_linear_tc_workspace_bytes(int M, int N, int K) {
  const int nt = pick_nt(M, N, K);
  const int64_t n_tiles = (N + BLOCK_N - 1) / BLOCK_N, m_tiles = (M + nt - 1) / nt;
  const int64_t splits = 16;      // upper bound of pick_splits / the test hook
  return (n_tiles * m_tiles * 4 + 255) / 256 * 256 + splits * n_tiles * m_tiles * 2 * nt * BLOCK_N * 4 + 256;
}

int rd_linear_tc_fused(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                       const EpiParams& epi, int dtype, void* ws, int64_t ws_bytes, const TcFuse* fuse, cudaStream_t st) {
  g_fuse = fuse;
  const int r = rd_linear_tc(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, st);
  g_fuse = nullptr;
  return r;
}

int rd_tc_make_map(CUtensorMap* map, const void* ptr, int64_t ld, int rows, int cols, int box_rows, int box_cols, int dtype, int swizzle128) {
  return make_map_ex(map, MapKey{ptr, ld, rows, cols, box_rows, box_cols, dtype, swizzle128 ? 1 : 0});
}

int rd_linear_wide_try(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                       const EpiParams& epi, int dtype, cudaStream_t st);

int rd_linear_tc(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                 const EpiParams& epi, int dtype, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (g_fuse == nullptr && g_force_splits <= 0 && !g_force_generic_epilogue) {
    // wide token counts with enough tiles to overlap: the persistent kernel (linear_wide.cu)
    const int r = rd_linear_wide_try(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, st);
    if (r != 0) return r < 0 ? r : RD_OK;
  }
  const int nt = pick_nt(M, N, K);
  const int splits = g_force_splits > 0 ? g_force_splits : 0;     // 0: chosen per kernel variant from its occupancy
  const bool sw = epi.act == RD_ACT_SWIGLU;
  RD_DISPATCH_DTYPE(dtype, T, {
    if (sw) return dispatch_nt<T, true>(nt, x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, splits, st);
    return dispatch_nt<T, false>(nt, x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, splits, st);
  });
}
