// Host interface of the stream-K decode GEMM (linear_sk.cu), used by engine_llm.cu for single-token steps (M <= 32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct rd_sk;

enum { RD_SK_PLAIN = 0, RD_SK_RES1 = 1, RD_SK_SWIGLU = 2 };

// RMSNorm fused on the INPUT of the GEMM: x is the raw residual stream, rstd comes from the per-tile sum-of-squares partials
// that the producing GEMM's epilogue left at `ssq` ([ssq_tiles][32] fp32), xn = T(w * T(x * rstd)) is formed in shared memory.
struct SkNorm {
  const float* ssq;
  int ssq_tiles;
  const void* ln_w;
  float eps;
};

int rd_sk_create(rd_sk** out);
void rd_sk_destroy(rd_sk* c);
// builds (and caches) the work table of a shape; allocates, so call it outside stream capture before the first launch
int rd_sk_plan(rd_sk* c, int N, int K, int mode);
// segments per CTA of the shape's work table (fused RMSNorm needs <= 2); huge if the shape cannot be planned
int rd_sk_max_segments(rd_sk* c, int N, int K, int mode);
// out[M,N] = epilogue(x[M,K] . W[N,K]^T), M <= 32.  mode RD_SK_SWIGLU: W has 2N rows (gate rows, then up rows).
// RD_SK_RES1: out = T(residual + T(acc)); with ssq_out != nullptr also writes sum_n out[m,n]^2 per 128-row tile to ssq_out[tile][m].
int rd_sk_linear(rd_sk* c, const void* x, int64_t ldx, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                 int mode, const void* residual, int64_t ld_res, const SkNorm* norm, float* ssq_out, int dtype, cudaStream_t st);
