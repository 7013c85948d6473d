// Probe (development aid, run on the GPU box): does tcgen05.mma read its A operand from TMEM the way the staged decode GEMM
// needs?  One CTA: W[128 x 64] and X[32 x 64] (bf16) go to shared memory in the K-major SWIZZLE_128B layout; D0 = W.X^T with A
// from shared memory (the production path) and D1 with A copied by the threads into TMEM (thread t = row t: the row's 128 bytes
// as 32 packed columns via tcgen05.st 32x32b.x32) and read by tcgen05.mma [d], [a_tmem], b_desc.  Both are checked against a host
// reference.   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/ts_probe tools/probes/ts_probe.cu && /tmp/ts_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../../radialog_b200/csrc/tc_ptx.cuh"

using namespace tcptx;

__device__ __forceinline__ uint32_t sw128(int r, int chunk) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4)); }

__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* W, const __nv_bfloat16* X, float* D0, float* D1) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;              // 128 rows x 128 B
  uint8_t* sX = smem + 16384;      // 32 rows x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 128 * 8; i += 128) { int r = i >> 3, c = i & 7; *(uint4*)(sW + sw128(r, c)) = *(const uint4*)(W + r * 64 + c * 8); }
  for (int i = tid; i < 32 * 8; i += 128) { int r = i >> 3, c = i & 7; *(uint4*)(sX + sw128(r, c)) = *(const uint4*)(X + r * 64 + c * 8); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  const uint32_t idesc = make_idesc(1, 128, 32);
  // ---- D0: A from shared memory ----
  if (warp == 0 && elect_one()) {
    const uint64_t da = make_smem_desc(smem_u32(sW)), db = make_smem_desc(smem_u32(sX));
    for (int k = 0; k < 4; ++k) tc_mma_f16(tb + 0, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, k > 0);
    tc_commit(&bar);
  }
  // ---- stage A into TMEM columns [64, 96): thread t owns row t ----
  {
    uint32_t r[32];
    for (int c = 0; c < 8; ++c) {
      uint4 v = *(const uint4*)(sW + sw128(tid, c));
      r[c * 4 + 0] = v.x; r[c * 4 + 1] = v.y; r[c * 4 + 2] = v.z; r[c * 4 + 3] = v.w;
    }
    const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + 64;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  mbar_wait(&bar, 0, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- D1: A from TMEM ----
  if (warp == 0 && elect_one()) {
    const uint64_t db = make_smem_desc(smem_u32(sX));
    for (int k = 0; k < 4; ++k) tc_mma_ts(tb + 32, tb + 64 + k * 8, db + (uint64_t)((k * 32) >> 4), idesc, k > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 1, 2);
  tc_fence_after();
  const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 32; c += 16) {
    uint32_t a[16], b[16];
    tc_ld16(tl + c, a);
    tc_ld16(tl + 32 + c, b);
    tc_wait_ld();
    for (int j = 0; j < 16; ++j) { D0[tid * 32 + c + j] = __uint_as_float(a[j]); D1[tid * 32 + c + j] = __uint_as_float(b[j]); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128) : "memory"); }
}

int main() {
  const int N = 128, M = 32, K = 64;
  __nv_bfloat16 *hW = (__nv_bfloat16*)malloc(N * K * 2), *hX = (__nv_bfloat16*)malloc(M * K * 2);
  float* ref = (float*)malloc(N * M * 4);
  srand(1);
  for (int i = 0; i < N * K; ++i) hW[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
  for (int i = 0; i < M * K; ++i) hX[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += __bfloat162float(hW[n * K + k]) * __bfloat162float(hX[m * K + k]);
      ref[n * M + m] = s;
    }
  __nv_bfloat16 *dW, *dX; float *d0, *d1;
  cudaMalloc(&dW, N * K * 2); cudaMalloc(&dX, M * K * 2); cudaMalloc(&d0, N * M * 4); cudaMalloc(&d1, N * M * 4);
  cudaMemcpy(dW, hW, N * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dX, hX, M * K * 2, cudaMemcpyHostToDevice);
  cudaMemset(d0, 0, N * M * 4); cudaMemset(d1, 0, N * M * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 16384 + 4096 + 1024>>>(dW, dX, d0, d1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  float* h0 = (float*)malloc(N * M * 4); float* h1 = (float*)malloc(N * M * 4);
  cudaMemcpy(h0, d0, N * M * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h1, d1, N * M * 4, cudaMemcpyDeviceToHost);
  double e0 = 0, e1 = 0, mx = 0;
  for (int i = 0; i < N * M; ++i) { e0 = fmax(e0, fabs(h0[i] - ref[i])); e1 = fmax(e1, fabs(h1[i] - ref[i])); mx = fmax(mx, fabs(ref[i])); }
  printf("max|ref| %.4f   A-from-smem max err %.6f   A-from-TMEM max err %.6f\n", mx, e0, e1);
  printf("sample: ref %.4f %.4f %.4f | smem %.4f %.4f %.4f | tmem %.4f %.4f %.4f\n", ref[0], ref[1], ref[33], h0[0], h0[1], h0[33], h1[0], h1[1], h1[33]);
  return (e0 < 1e-3 && e1 < 1e-3) ? 0 : 1;
}
