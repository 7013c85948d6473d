// Error reporting, version and device gate of libradialog_b200.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void rd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* rd_last_error(void) { return g_err; }
extern "C" int rd_version(void) { return 100; }

extern "C" int rd_device_ok(int dev) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    rd_set_error("rd_device_ok: no CUDA device %d", dev);
    return 0;
  }
  if (p.major != 10) {
    rd_set_error("rd_device_ok: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, p.major, p.minor);
    return 0;
  }
  return 1;
}
