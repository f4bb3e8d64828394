// Shared host-side helpers of the C-ABI library: error reporting, launch checks, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace cmu {

std::string& last_error();
int fail(const char* fmt, ...);

#define CMU_CHECK_CUDA(expr)                                                                    \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return ::cmu::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define CMU_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) return ::cmu::fail(__VA_ARGS__); \
  } while (0)

// every kernel launch of the library goes through this macro: error check + launch counter (cmu_launch_count)
extern unsigned long long g_launch_count;
#define CMU_LAUNCH_CHECK()                 \
  do {                                     \
    ++::cmu::g_launch_count;               \
    CMU_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

// Encodes a bf16 tiled tensor map with SWIZZLE_128B (inner box = 64 elements = 128 bytes).
// dims/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);

int num_sms();
int debug_knob(int key);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace cmu
