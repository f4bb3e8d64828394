#include "common.cuh"

#include <stdarg.h>

#include <mutex>

#include "../../include/cmu_b200.h"

namespace cmu {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
                (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
                (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
                rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

unsigned long long g_launch_count = 0;
static int g_knobs[16] = {0};
int debug_knob(int key) { return (key >= 0 && key < 16) ? g_knobs[key] : 0; }

}  // namespace cmu

extern "C" {

const char* cmu_last_error(void) { return cmu::last_error().c_str(); }

int cmu_version(void) { return 100; }

long long cmu_launch_count(void) { return (long long)cmu::g_launch_count; }

int cmu_debug_set(int key, int value) {
  if (key < 0 || key >= 16) return cmu::fail("cmu_debug_set: bad key %d", key);
  cmu::g_knobs[key] = value;
  return 0;
}

int cmu_device_check(void) {
  int dev = 0;
  CMU_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CMU_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return cmu::fail("device %d is sm_%d%d; this library is sm_100a only", dev, prop.major, prop.minor);
  return 0;
}

}  // extern "C"
