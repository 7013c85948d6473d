// GPU-side chest-X-ray image preprocessing: replaces, bit-exactly, the reference's per-image CPU pipeline
//   remap_to_uint8 (demo.py:173-203) -> PIL "L" image (demo.py:218) -> Resize(512) -> CenterCrop(448) -> ToTensor ->
//   ExpandChannels (ReportDataset.py:80-106)
// so that the image goes from a raw grey array on the device to the float32 [3,448,448] tensor forward_image expects
// without a host round trip.  The resize is Pillow's 8-bit bilinear resample (Resample.c: double-precision coefficients,
// 22-bit fixed point, horizontal pass then vertical pass with an 8-bit intermediate); the coefficients are computed on the
// host with the same double arithmetic, the passes run as integer kernels, so every output byte equals Pillow's.
// HBM-bound byte work: one thread per output pixel, coalesced along x; only the rows / columns that survive the centre
// crop are computed.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

struct Axis {
  int in_size = 0, out_size = 0, ksize = 0;
  std::vector<int> bounds;     // [out][2] = (first input index, tap count)
  std::vector<int> kk;         // [out][ksize]
};

// Resample.c precompute_coeffs (bilinear, support 1.0, box = whole axis) + normalize_coeffs_8bpc
void precompute(Axis& a, int in_size, int out_size) {
  a.in_size = in_size; a.out_size = out_size;
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  a.ksize = (int)ceil(support) * 2 + 1;
  a.bounds.assign((size_t)out_size * 2, 0);
  a.kk.assign((size_t)out_size * a.ksize, 0);
  std::vector<double> k(a.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < a.ksize; ++x) k[x] = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0.0) v = -v;
      const double w = v < 1.0 ? 1.0 - v : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < a.ksize; ++x)
      a.kk[(size_t)xx * a.ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
    a.bounds[(size_t)xx * 2] = xmin;
    a.bounds[(size_t)xx * 2 + 1] = xmax;
  }
}

template <class TIn> __device__ __forceinline__ double as_double(TIn v) { return (double)v; }

// per-block min / max in double (exact for u8 / u16 / f32 inputs)
template <class TIn>
__global__ void __launch_bounds__(256)
minmax_partial_kernel(const TIn* __restrict__ img, long long n, double* __restrict__ part) {
  double mn = INFINITY, mx = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = as_double(img[i]);
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
  __shared__ double smn[256], smx[256];
  smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      smn[threadIdx.x] = smn[threadIdx.x + s] < smn[threadIdx.x] ? smn[threadIdx.x + s] : smn[threadIdx.x];
      smx[threadIdx.x] = smx[threadIdx.x + s] > smx[threadIdx.x] ? smx[threadIdx.x + s] : smx[threadIdx.x];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = smn[0]; part[2 * blockIdx.x + 1] = smx[0]; }
}

__global__ void minmax_final_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  double mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    mn = part[2 * i] < mn ? part[2 * i] : mn;
    mx = part[2 * i + 1] > mx ? part[2 * i + 1] : mx;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
  }
  if (threadIdx.x == 0) { out[0] = mn; out[1] = mx - mn; }     // array -= array.min(); then array.max()
}

// remap_to_uint8 (demo.py:184,200-203) in the same float64 operation order, fused with the horizontal resample pass:
// one thread per (needed source row, needed output column)
template <class TIn, bool REMAP>
__global__ void __launch_bounds__(256)
resample_h_kernel(const TIn* __restrict__ img, int W, const double* __restrict__ mm, const int* __restrict__ bounds,
                  const int* __restrict__ kk, int ksize, int row0, int nrows, int col0, int ncols, uint8_t* __restrict__ tmp) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (ox >= ncols || r >= nrows) return;
  const int xx = col0 + ox, y = row0 + r;
  const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
  const double mn = REMAP ? mm[0] : 0.0, mx = REMAP ? mm[1] : 1.0;
  const TIn* src = img + (long long)y * W + xmin;
  int ss = 1 << (PRECISION_BITS - 1);
  for (int x = 0; x < n; ++x) {
    int px;
    if (REMAP) {
      double a = as_double(src[x]);
      a -= mn; a /= mx; a *= 255.0;
      px = (int)(unsigned char)a;
    } else {
      px = (int)src[x];
    }
    ss += px * kk[(long long)xx * ksize + x];
  }
  ss >>= PRECISION_BITS;
  tmp[(long long)r * ncols + ox] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

// identity "horizontal pass" (width unchanged): just the remap + column crop
template <class TIn, bool REMAP>
__global__ void __launch_bounds__(256)
remap_copy_kernel(const TIn* __restrict__ img, int W, const double* __restrict__ mm, int row0, int nrows, int col0, int ncols,
                  uint8_t* __restrict__ tmp) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (ox >= ncols || r >= nrows) return;
  const TIn v = img[(long long)(row0 + r) * W + col0 + ox];
  int px;
  if (REMAP) {
    double a = as_double(v);
    a -= mm[0]; a /= mm[1]; a *= 255.0;
    px = (int)(unsigned char)a;
  } else {
    px = (int)v;
  }
  tmp[(long long)r * ncols + ox] = (uint8_t)px;
}

// vertical pass over the 8-bit intermediate + centre crop + ToTensor (/255 in float32) + ExpandChannels (3 copies)
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, int ncols, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                  int row0, int top, int crop, int identity, float* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (ox >= crop) return;
  int v;
  if (identity) {
    v = tmp[(long long)(top + oy - row0) * ncols + ox];
  } else {
    const int yy = top + oy;
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    int ss = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < n; ++y) ss += (int)tmp[(long long)(ymin + y - row0) * ncols + ox] * kk[(long long)yy * ksize + y];
    ss >>= PRECISION_BITS;
    v = ss < 0 ? 0 : (ss > 255 ? 255 : ss);
  }
  const float f = (float)v / 255.0f;
  const long long plane = (long long)crop * crop, o = (long long)oy * crop + ox;
  out[o] = f; out[plane + o] = f; out[2 * plane + o] = f;
}

}  // namespace

struct rd_preproc {
  int max_h = 0, max_w = 0, resize = 0, crop = 0;
  int H = -1, W = -1;                       // shape the cached coefficient tables belong to
  int new_h = 0, new_w = 0, top = 0, left = 0, row0 = 0, nrows = 0;
  Axis ax, ay;
  int *d_bx = nullptr, *d_kx = nullptr, *d_by = nullptr, *d_ky = nullptr;
  size_t cap_kx = 0, cap_ky = 0, cap_bx = 0, cap_by = 0;
  uint8_t* d_tmp = nullptr;
  double *d_part = nullptr, *d_mm = nullptr;
  int64_t launches = 0;
};

extern "C" int rd_preproc_create(int max_h, int max_w, int resize, int crop, rd_preproc** out) {
  RD_REQUIRE(out && max_h > 0 && max_w > 0 && resize >= crop && crop > 0, "rd_preproc_create: bad arguments");
  int dev = 0;
  RD_CHECK_CUDA(cudaGetDevice(&dev));
  if (!rd_device_ok(dev)) return RD_ERR_UNSUPPORTED;
  rd_preproc* p = new rd_preproc();
  p->max_h = max_h; p->max_w = max_w; p->resize = resize; p->crop = crop;
  cudaError_t e = cudaMalloc((void**)&p->d_tmp, (size_t)max_h * (size_t)(crop > max_w ? crop : max_w));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_part, 2 * 1024 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_mm, 2 * sizeof(double));
  if (e != cudaSuccess) {
    rd_set_error("rd_preproc_create: CUDA error %s", cudaGetErrorString(e));
    if (p->d_tmp) cudaFree(p->d_tmp);
    if (p->d_part) cudaFree(p->d_part);
    delete p;
    return RD_ERR_CUDA;
  }
  *out = p;
  return RD_OK;
}

extern "C" void rd_preproc_destroy(rd_preproc* p) {
  if (!p) return;
  void* ptrs[] = {p->d_bx, p->d_kx, p->d_by, p->d_ky, p->d_tmp, p->d_part, p->d_mm};
  for (void* q : ptrs) if (q) cudaFree(q);
  delete p;
}

extern "C" int64_t rd_preproc_launch_count(rd_preproc* p) { return p ? p->launches : 0; }

static int upload(int** dptr, size_t* cap, const std::vector<int>& v, cudaStream_t st) {
  if (v.size() > *cap) {
    if (*dptr) RD_CHECK_CUDA(cudaFree(*dptr));
    *dptr = nullptr;
    RD_CHECK_CUDA(cudaMalloc((void**)dptr, v.size() * sizeof(int)));
    *cap = v.size();
  }
  RD_CHECK_CUDA(cudaMemcpyAsync(*dptr, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  return RD_OK;
}

// img_dev: [H,W] row-major grey image; dtype 0 = uint8, 1 = uint16, 2 = float32.  remap != 0 applies remap_to_uint8 first
// (demo.py load_image); remap == 0 takes a uint8 image as it is (an already remapped PIL "L" image).
// out_dev: float32 [3, crop, crop].  Not capturable in a CUDA graph when the image shape changes (coefficient upload).
extern "C" int rd_preproc_run(rd_preproc* p, const void* img_dev, int dtype, int H, int W, int remap, float* out_dev, void* stream) {
  RD_REQUIRE(p && img_dev && out_dev, "rd_preproc_run: null argument");
  RD_REQUIRE(H > 0 && W > 0 && H <= p->max_h && W <= p->max_w, "rd_preproc_run: image %dx%d outside the handle's limits %dx%d", H, W, p->max_h, p->max_w);
  RD_REQUIRE(dtype >= 0 && dtype <= 2, "rd_preproc_run: dtype must be 0 (uint8), 1 (uint16) or 2 (float32)");
  RD_REQUIRE(remap || dtype == 0, "rd_preproc_run: remap=0 needs a uint8 image");
  cudaStream_t st = (cudaStream_t)stream;
  if (H != p->H || W != p->W) {
    RD_CHECK_CUDA(cudaStreamSynchronize(st));          // the host tables may still be in flight to the device
    // torchvision Resize(int): smaller edge -> resize, the other int(resize * long / short)
    const int shrt = W <= H ? W : H, lng = W <= H ? H : W;
    const int new_long = (int)((double)p->resize * lng / shrt);
    p->new_w = W <= H ? p->resize : new_long;
    p->new_h = W <= H ? new_long : p->resize;
    precompute(p->ax, W, p->new_w);
    precompute(p->ay, H, p->new_h);
    // torchvision center_crop: int(round((size - crop) / 2.0)) with Python's round-half-to-even = nearbyint
    p->top = (int)nearbyint((p->new_h - p->crop) / 2.0);
    p->left = (int)nearbyint((p->new_w - p->crop) / 2.0);
    // source rows the cropped output rows read in the vertical pass
    if (p->new_h == H) { p->row0 = p->top; p->nrows = p->crop; }
    else {
      int lo = H, hi = 0;
      for (int yy = p->top; yy < p->top + p->crop; ++yy) {
        const int a = p->ay.bounds[2 * yy], b = a + p->ay.bounds[2 * yy + 1];
        lo = a < lo ? a : lo; hi = b > hi ? b : hi;
      }
      p->row0 = lo; p->nrows = hi - lo;
    }
    RD_CHECK(upload(&p->d_bx, &p->cap_bx, p->ax.bounds, st));
    RD_CHECK(upload(&p->d_kx, &p->cap_kx, p->ax.kk, st));
    RD_CHECK(upload(&p->d_by, &p->cap_by, p->ay.bounds, st));
    RD_CHECK(upload(&p->d_ky, &p->cap_ky, p->ay.kk, st));
    p->H = H; p->W = W;
  }
  const long long n = (long long)H * W;
  const int nblk = (int)((n + 256 * 16 - 1) / (256 * 16) < 1024 ? (n + 256 * 16 - 1) / (256 * 16) : 1024);
  const dim3 gh((p->crop + 255) / 256, p->nrows), gv((p->crop + 255) / 256, p->crop);
  const bool hid = p->new_w == W;
#define RD_PRE_DISPATCH(TIN)                                                                                                       \
  do {                                                                                                                             \
    if (remap) {                                                                                                                   \
      minmax_partial_kernel<TIN><<<nblk, 256, 0, st>>>((const TIN*)img_dev, n, p->d_part);                                         \
      minmax_final_kernel<<<1, 32, 0, st>>>(p->d_part, nblk, p->d_mm);                                                             \
      p->launches += 2;                                                                                                            \
      if (hid) remap_copy_kernel<TIN, true><<<gh, 256, 0, st>>>((const TIN*)img_dev, W, p->d_mm, p->row0, p->nrows, p->left, p->crop, p->d_tmp); \
      else resample_h_kernel<TIN, true><<<gh, 256, 0, st>>>((const TIN*)img_dev, W, p->d_mm, p->d_bx, p->d_kx, p->ax.ksize, p->row0, p->nrows, p->left, p->crop, p->d_tmp); \
    } else {                                                                                                                       \
      if (hid) remap_copy_kernel<TIN, false><<<gh, 256, 0, st>>>((const TIN*)img_dev, W, p->d_mm, p->row0, p->nrows, p->left, p->crop, p->d_tmp); \
      else resample_h_kernel<TIN, false><<<gh, 256, 0, st>>>((const TIN*)img_dev, W, p->d_mm, p->d_bx, p->d_kx, p->ax.ksize, p->row0, p->nrows, p->left, p->crop, p->d_tmp); \
    }                                                                                                                              \
  } while (0)
  if (dtype == 0) RD_PRE_DISPATCH(uint8_t);
  else if (dtype == 1) RD_PRE_DISPATCH(uint16_t);
  else RD_PRE_DISPATCH(float);
#undef RD_PRE_DISPATCH
  resample_v_kernel<<<gv, 256, 0, st>>>(p->d_tmp, p->crop, p->d_by, p->d_ky, p->ay.ksize, p->row0, p->top, p->crop, p->new_h == H ? 1 : 0, out_dev);
  p->launches += 2;
  RD_LAUNCH_CHECK();
  return RD_OK;
}
