// rd_linear: out[M,N] = epilogue(x[M,K] . W[N,K]^T)
//   * M <= 4   : streaming GEMV on CUDA cores, 128-bit no-allocate weight loads, fp32 accumulate.  This is the
//                single-token decode path: HBM-bound, 2*N*K bytes per call.
//   * otherwise: tcgen05 tensor-core tiles (linear_tc.cu).
//   * algo 3   : plain SIMT tiled kernel, kept as the on-device cross-check for the two above.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// GEMV: one warp per pair of weight rows; lanes stride K in 16-byte chunks; KU chunks in flight per row
// ------------------------------------------------------------------------------------------------
template <class T> struct Pair2;
template <> struct Pair2<__half> {
  static __device__ __forceinline__ float2 cvt(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
};
template <> struct Pair2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 cvt(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
};

template <class T>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = Pair2<T>::cvt(u.x), b = Pair2<T>::cvt(u.y), c = Pair2<T>::cvt(u.z), d = Pair2<T>::cvt(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}

constexpr int GEMV_WARPS = 8;
constexpr int GEMV_KU = 4;

template <class T, int MB>
__global__ void __launch_bounds__(GEMV_WARPS * 32)
gemv_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ w, int64_t ldw, T* __restrict__ out,
            int64_t ldo, int M, int N, int K, EpiParams epi) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int task = blockIdx.x * GEMV_WARPS + (threadIdx.x >> 5);
  const bool swiglu = epi.act == RD_ACT_SWIGLU;
  int n0, n1;
  if (swiglu) { n0 = task; n1 = task + N; if (task >= N) return; }
  else { n0 = 2 * task; n1 = n0 + 1; if (n0 >= N) return; }
  const bool has1 = swiglu || (n1 < N);
  const T* w0 = w + (int64_t)n0 * ldw;
  const T* w1 = w + (int64_t)(has1 ? n1 : n0) * ldw;

  float acc0[MB], acc1[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }

  // the weight stream does not depend on the previous kernel: issue the first loads before waiting on it
  const int kstep = 256 * GEMV_KU;
  uint4 wa[GEMV_KU], wb[GEMV_KU];
  int kbase = lane * 8;
#pragma unroll
  for (int u = 0; u < GEMV_KU; ++u) {
    int kk = kbase + u * 256;
    if (kk < K) { wa[u] = ldg_stream(w0 + kk); wb[u] = ldg_stream(w1 + kk); }
    else { wa[u] = make_uint4(0, 0, 0, 0); wb[u] = make_uint4(0, 0, 0, 0); }
  }
  pdl_wait();
  for (; kbase < K; kbase += kstep) {
    uint4 na[GEMV_KU], nb[GEMV_KU];
    const int knext = kbase + kstep;
#pragma unroll
    for (int u = 0; u < GEMV_KU; ++u) {           // software pipeline: next block's weights in flight
      int kk = knext + u * 256;
      if (kk < K) { na[u] = ldg_stream(w0 + kk); nb[u] = ldg_stream(w1 + kk); }
      else { na[u] = make_uint4(0, 0, 0, 0); nb[u] = make_uint4(0, 0, 0, 0); }
    }
#pragma unroll
    for (int u = 0; u < GEMV_KU; ++u) {
      int kk = kbase + u * 256;
      if (kk < K) {
        float fa[8], fb[8];
        unpack8<T>(wa[u], fa);
        unpack8<T>(wb[u], fb);
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          if (m < M) {
            float fx[8];
            unpack8<T>(*reinterpret_cast<const uint4*>(x + (int64_t)m * ldx + kk), fx);
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc0[m] = fmaf(fa[e], fx[e], acc0[m]); acc1[m] = fmaf(fb[e], fx[e], acc1[m]); }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GEMV_KU; ++u) { wa[u] = na[u]; wb[u] = nb[u]; }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) { acc0[m] = warp_sum(acc0[m]); acc1[m] = warp_sum(acc1[m]); }
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      if (m < M) {
        if (swiglu) {
          out[(int64_t)m * ldo + n0] = epilogue_elem<T>(epi, acc0[m], acc1[m], m, n0);
        } else {
          out[(int64_t)m * ldo + n0] = epilogue_elem<T>(epi, acc0[m], 0.f, m, n0);
          if (has1) out[(int64_t)m * ldo + n1] = epilogue_elem<T>(epi, acc1[m], 0.f, m, n1);
        }
      }
    }
  }
}

template <class T>
static int launch_gemv(const T* x, int64_t ldx, const T* w, int64_t ldw, T* out, int64_t ldo, int M, int N, int K,
                       const EpiParams& epi, cudaStream_t st, bool pdl) {
  int tasks = epi.act == RD_ACT_SWIGLU ? N : (N + 1) / 2;
  dim3 grid((tasks + GEMV_WARPS - 1) / GEMV_WARPS), block(GEMV_WARPS * 32);
  cudaError_t e;
  if (M == 1) e = rd_launch(gemv_kernel<T, 1>, grid, block, 0, st, pdl, x, ldx, w, ldw, out, ldo, M, N, K, epi);
  else if (M == 2) e = rd_launch(gemv_kernel<T, 2>, grid, block, 0, st, pdl, x, ldx, w, ldw, out, ldo, M, N, K, epi);
  else e = rd_launch(gemv_kernel<T, 4>, grid, block, 0, st, pdl, x, ldx, w, ldw, out, ldo, M, N, K, epi);
  RD_CHECK_CUDA(e);
  return RD_OK;
}

// ------------------------------------------------------------------------------------------------
// SIMT tiled kernel (validation / odd shapes): 64x64 tile, BK 32, 4x4 per thread
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
simt_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ w, int64_t ldw, T* __restrict__ out,
            int64_t ldo, int M, int N, int K, EpiParams epi) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int BM = 64, BN = 64, BK = 32;
  __shared__ float sx[BK][BM + 1], sw[BK][BN + 1], su[BK][BN + 1];
  const bool swiglu = epi.act == RD_ACT_SWIGLU;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {}, accu[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      int r = i / BK, c = i % BK;
      int m = m0 + r, n = n0 + r, k = k0 + c;
      sx[c][r] = (m < M && k < K) ? Tr<T>::f(x[(int64_t)m * ldx + k]) : 0.f;
      sw[c][r] = (n < N && k < K) ? Tr<T>::f(w[(int64_t)n * ldw + k]) : 0.f;
      if (swiglu) su[c][r] = (n < N && k < K) ? Tr<T>::f(w[(int64_t)(n + N) * ldw + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4], bu[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sx[k][ty * 4 + i]; b[i] = sw[k][tx * 4 + i]; bu[i] = swiglu ? su[k][tx * 4 + i] : 0.f; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = fmaf(a[i], b[j], acc[i][j]); accu[i][j] = fmaf(a[i], bu[j], accu[i][j]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) out[(int64_t)m * ldo + n] = epilogue_elem<T>(epi, acc[i][j], accu[i][j], m, n);
    }
}

// ------------------------------------------------------------------------------------------------
// dispatcher
// ------------------------------------------------------------------------------------------------
// programmatic dependent launch between the kernels of a step: on by default (3.77 vs 4.02 ms per B=32 decode step)
static bool g_pdl = true;
extern "C" int rd_set_pdl(int on) { g_pdl = on != 0; return RD_OK; }
bool rd_pdl_enabled() { return g_pdl; }

extern "C" int64_t rd_linear_workspace_bytes(int M, int N, int K) { return rd_linear_tc_workspace_bytes(M, N, K); }

extern "C" int rd_linear(const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N,
                         int K, const rd_epilogue* e, int dtype, int algo, void* ws, int64_t ws_bytes, void* stream) {
  RD_REQUIRE(M > 0 && N > 0 && K > 0, "rd_linear: bad shape M=%d N=%d K=%d", M, N, K);
  RD_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "rd_linear: K, ldx, ldw must be multiples of 8 (K=%d ldx=%lld ldw=%lld)",
             K, (long long)ldx, (long long)ldw);
  RD_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "rd_linear: x and W must be 16-byte aligned");
  EpiParams epi = make_epi(e);
  RD_REQUIRE(epi.lora_r == 0 || (epi.lora_t && epi.lora_b), "rd_linear: lora_r set without lora_t/lora_b");
  cudaStream_t st = (cudaStream_t)stream;
  // auto: the tcgen05 path at every M.  Measured on B200 it also wins at M = 1..4 (TMA streams + programmatic-launch weight
  // prefetch: 2.85 vs 3.31 ms per Vicuna-7B decode step); the CUDA-core GEMV stays available as algo 1.
  if (algo == 0) algo = 2;
  if (algo == 1) {
    RD_REQUIRE(M <= 4, "rd_linear: GEMV path needs M<=4 (got %d)", M);
    RD_DISPATCH_DTYPE(dtype, T, { return launch_gemv<T>((const T*)x, ldx, (const T*)w, ldw, (T*)out, ldo, M, N, K, epi, st, g_pdl); });
  } else if (algo == 2) {
    return rd_linear_tc(x, ldx, w, ldw, out, ldo, M, N, K, epi, dtype, ws, ws_bytes, st);
  } else if (algo == 3) {
    dim3 grid((N + 63) / 64, (M + 63) / 64), block(256);
    RD_DISPATCH_DTYPE(dtype, T, {
      RD_CHECK_CUDA(rd_launch(simt_kernel<T>, grid, block, 0, st, g_pdl, (const T*)x, ldx, (const T*)w, ldw, (T*)out, ldo, M, N, K, epi));
      return RD_OK;
    });
  }
  rd_set_error("rd_linear: unknown algo %d", algo);
  return RD_ERR_INVALID;
}
