// Shared device/host helpers for libradialog_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/radialog_b200.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
void rd_set_error(const char* fmt, ...);

#define RD_CHECK_CUDA(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      rd_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RD_ERR_CUDA;                                                                         \
    }                                                                                             \
  } while (0)

#define RD_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      rd_set_error(__VA_ARGS__);   \
      return RD_ERR_INVALID;       \
    }                              \
  } while (0)

#define RD_CHECK(expr)             \
  do {                             \
    int _r = (expr);               \
    if (_r != RD_OK) return _r;    \
  } while (0)

#define RD_LAUNCH_CHECK() RD_CHECK_CUDA(cudaGetLastError())

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: set it once per (kernel instantiation, device), so
// that one process can drive engines on several GPUs.  Usage: RD_SMEM_ATTR_ONCE(bytes, kernel<T, ...>);
#define RD_SMEM_ATTR_ONCE(bytes, ...)                                                                                   \
  do {                                                                                                                  \
    static bool _rd_attr_done[64] = {};                                                                                 \
    int _rd_dev = 0;                                                                                                    \
    RD_CHECK_CUDA(cudaGetDevice(&_rd_dev));                                                                             \
    if (_rd_dev < 0 || _rd_dev >= 64 || !_rd_attr_done[_rd_dev]) {                                                      \
      RD_CHECK_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
      if (_rd_dev >= 0 && _rd_dev < 64) _rd_attr_done[_rd_dev] = true;                                                  \
    }                                                                                                                   \
  } while (0)

// ------------------------------------------------------------------------------------------------
// storage-type traits: every "T(.)" rounding point of the reference goes through Tr<T>::r
// ------------------------------------------------------------------------------------------------
template <class T> struct Tr;
template <> struct Tr<__half> {
  static __device__ __forceinline__ float f(__half x) { return __half2float(x); }
  static __device__ __forceinline__ __half r(float x) { return __float2half_rn(x); }
  static __device__ __forceinline__ float rr(float x) { return __half2float(__float2half_rn(x)); }
  static constexpr float lowest() { return -65504.0f; }   // torch.finfo(torch.float16).min
  static constexpr int umma_fmt = 0;                       // UMMA F16F32Format::F16
};
template <> struct Tr<__nv_bfloat16> {
  static __device__ __forceinline__ float f(__nv_bfloat16 x) { return __bfloat162float(x); }
  static __device__ __forceinline__ __nv_bfloat16 r(float x) { return __float2bfloat16_rn(x); }
  static __device__ __forceinline__ float rr(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
  static constexpr float lowest() { return -3.3895313892515355e38f; }  // torch.finfo(torch.bfloat16).min
  static constexpr int umma_fmt = 1;                                    // UMMA F16F32Format::BF16
};

#define RD_DISPATCH_DTYPE(dtype, T, ...)                          \
  do {                                                            \
    if ((dtype) == RD_F16) {                                      \
      using T = __half;                                           \
      __VA_ARGS__                                                 \
    } else if ((dtype) == RD_BF16) {                              \
      using T = __nv_bfloat16;                                    \
      __VA_ARGS__                                                 \
    } else {                                                      \
      rd_set_error("unsupported dtype %d", (int)(dtype));         \
      return RD_ERR_UNSUPPORTED;                                  \
    }                                                             \
  } while (0)

// 8 storage elements = one 128-bit load
template <class T> struct alignas(16) Vec8 { T v[8]; };

template <class T>
__device__ __forceinline__ Vec8<T> ld_stream16(const T* p) {   // streaming read-once data (weights, KV)
  Vec8<T> r;
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  *reinterpret_cast<uint4*>(&r) = u;
  return r;
}
// small weights that every decode step re-reads (lora_B rows, 4 MB over all layers): loaded with an L2 evict-last policy so that
// the 13 GB of evict-first weight traffic streaming through the 126 MB L2 in between does not push them out
template <class T>
__device__ __forceinline__ Vec8<T> ld16_keep(const T* p) {
  Vec8<T> r;
  uint4 u;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p), "l"(0x14F0000000000000ull));
  *reinterpret_cast<uint4*>(&r) = u;
  return r;
}
template <class T>
__device__ __forceinline__ Vec8<T> ld16(const T* p) {          // cached (activations)
  Vec8<T> r;
  *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(p);
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Programmatic dependent launch: let the next kernel's prologue overlap this kernel's tail, and wait for the
// previous kernel's memory before touching its outputs.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// L2 prefetch of a slice of an upcoming weight matrix (cp.async.bulk.prefetch.L2): the small latency-bound kernels between
// two GEMMs leave HBM idle; they use that time to pull the next GEMM's weights into the 126 MB L2.  CTA `cta` of `n_ctas`
// requests its share of [ptr, ptr+bytes), spread over its threads in 16 KB requests; nothing waits on them.
__device__ __forceinline__ void l2_prefetch_slice(const void* ptr, long long bytes, int cta, int n_ctas, int tid, int n_threads) {
  if (ptr == nullptr || bytes <= 0) return;
  const long long per = ((bytes + n_ctas - 1) / n_ctas + 127) / 128 * 128;     // this CTA's share
  const long long off = (long long)cta * per;
  if (off >= bytes) return;
  const long long len = (bytes - off < per ? bytes - off : per) / 16 * 16;
  const char* base = reinterpret_cast<const char*>(ptr) + off;
  constexpr long long CHUNK = 16384;                                             // one request per thread and round
  for (long long o = (long long)tid * CHUNK; o < len; o += (long long)n_threads * CHUNK) {
    const unsigned sz = (unsigned)(len - o < CHUNK ? len - o : CHUNK);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + o), "r"(sz) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// shared epilogue (rd_linear): see include/radialog_b200.h for the rounding contract
// ------------------------------------------------------------------------------------------------
struct EpiParams {
  const float* bias;
  const void* residual;
  int64_t ld_res;
  int res_mode;
  int act;
  const void* lora_t;
  const void* lora_b;
  int lora_r;
  float lora_scale;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// acc: fp32 accumulator of out[m,n]; acc_up: accumulator of the paired up_proj row (SWIGLU only)
// lora_b_row: optional register copy of lora_B[n, 0..lora_r) (lora_r <= 16) for callers whose thread owns one output
// column n across many rows m (tcgen05 epilogue); nullptr = read it from global memory per element.
constexpr int RD_MAX_LORA_REG = 16;
template <class T>
__device__ __forceinline__ void load_lora_b_row(const EpiParams& p, int n, float (&breg)[RD_MAX_LORA_REG]) {
  const T* b = reinterpret_cast<const T*>(p.lora_b) + (int64_t)n * p.lora_r;
#pragma unroll
  for (int r = 0; r < RD_MAX_LORA_REG; ++r) breg[r] = r < p.lora_r ? Tr<T>::f(b[r]) : 0.f;
}

template <class T>
__device__ __forceinline__ T epilogue_elem(const EpiParams& p, float acc, float acc_up, int m, int n,
                                           const float* lora_b_row = nullptr) {
  float v = acc;
  if (p.bias) v += p.bias[n];
  float y;
  if (p.act == RD_ACT_SWIGLU) {
    float g = Tr<T>::rr(acc), u = Tr<T>::rr(acc_up);
    y = Tr<T>::rr(Tr<T>::rr(silu_f(g)) * u);
  } else {
    const bool res2 = p.residual && p.res_mode == 2;
    if (res2) v += Tr<T>::f(reinterpret_cast<const T*>(p.residual)[(int64_t)m * p.ld_res + n]);
    if (p.act == RD_ACT_RELU) v = fmaxf(v, 0.0f);
    else if (p.act == RD_ACT_GELU) v = gelu_erf(v);
    if (res2) return Tr<T>::r(v);
    y = Tr<T>::rr(v);
  }
  if (p.lora_r > 0) {
    const T* t = reinterpret_cast<const T*>(p.lora_t) + (int64_t)m * p.lora_r;
    float s = 0.f;
    if (lora_b_row != nullptr && p.lora_r == RD_MAX_LORA_REG) {
      // lora_t row: two 128-bit loads, identical across the warp (same m) -> one L1 broadcast each
      Vec8<T> t0 = ld16(t), t1 = ld16(t + 8);
#pragma unroll
      for (int r = 0; r < 8; ++r) s += Tr<T>::f(t0.v[r]) * lora_b_row[r];
#pragma unroll
      for (int r = 0; r < 8; ++r) s += Tr<T>::f(t1.v[r]) * lora_b_row[8 + r];
    } else {
      const T* b = reinterpret_cast<const T*>(p.lora_b) + (int64_t)n * p.lora_r;
      for (int r = 0; r < p.lora_r; ++r) s += Tr<T>::f(t[r]) * Tr<T>::f(b[r]);
    }
    y = Tr<T>::rr(y + Tr<T>::rr(p.lora_scale * Tr<T>::rr(s)));
  }
  if (p.residual) y = Tr<T>::rr(Tr<T>::f(reinterpret_cast<const T*>(p.residual)[(int64_t)m * p.ld_res + n]) + y);
  return Tr<T>::r(y);
}

static inline EpiParams make_epi(const rd_epilogue* e) {
  EpiParams p{};
  if (e) {
    p.bias = e->bias_dev; p.residual = e->residual_dev; p.ld_res = e->ld_res; p.res_mode = e->res_mode ? e->res_mode : 1;
    p.act = e->act; p.lora_t = e->lora_t_dev; p.lora_b = e->lora_b_dev; p.lora_r = e->lora_r; p.lora_scale = e->lora_scale;
  } else {
    p.res_mode = 1;
  }
  return p;
}

// kernel-launch helper with optional programmatic-dependent-launch attribute
template <class... KArgs, class... Args>
static inline cudaError_t rd_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// internal entry points shared between translation units
int rd_linear_tc(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                 const EpiParams& epi, int dtype, void* ws, int64_t ws_bytes, cudaStream_t st);
int64_t rd_linear_tc_workspace_bytes(int M, int N, int K);
// decode (M <= 32) extra of the tcgen05 GEMM
struct TcFuse {
  // partials-out split-K (see TcParams::part_out in linear_tc.cu): the GEMM leaves its fp32 split-K partials in
  // part_out[splits][NT][N] for the consumer to sum; splits_out[0..1] receives the split count and the slab's row count NT.
  float* part_out = nullptr;
  int64_t part_bytes = 0;
  int* splits_out = nullptr;
};
int rd_linear_tc_fused(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                       const EpiParams& epi, int dtype, void* ws, int64_t ws_bytes, const TcFuse* fuse, cudaStream_t st);
