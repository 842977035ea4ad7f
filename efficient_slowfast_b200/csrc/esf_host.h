// Host-side helpers shared by the .cu translation units of libesf_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/esf.h"

namespace esf {

int set_error(int code, const char* fmt, ...);  // stores a thread-local message, returns `code`
void count_launch(int n = 1);

#define ESF_CHECK_ARG(cond, ...)                                \
  do {                                                          \
    if (!(cond)) return ::esf::set_error(ESF_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define ESF_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return ::esf::set_error(ESF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                              __LINE__);                                                                 \
  } while (0)

// checks the launch itself (configuration errors and sticky errors such as a device-side trap)
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(ESF_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  count_launch();
  return ESF_OK;
}

inline bool view_ok(const esf_view* v) {
  return v && v->ptr && v->B > 0 && v->T > 0 && v->H > 0 && v->W > 0 && v->C > 0 && v->sW >= v->C;
}

inline bool is16(int dtype) { return dtype == ESF_BF16 || dtype == ESF_F16; }

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int num_sms();  // SM count of the CURRENT device (cached per device ordinal); 0 when there is no device

// Kernel attributes (max dynamic shared memory) are per DEVICE: a process-wide "already set" flag would skip the second
// GPU of a process.  Returns the slot of the current device in a per-call-site table (or nullptr: always set).
constexpr int kMaxDevices = 64;
inline unsigned char* device_slot(unsigned char (&table)[kMaxDevices]) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  return &table[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();  // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time libcuda)

}  // namespace esf

// One planned kernel launch behind the opaque esf_op handle of the C ABI.
struct esf_op {
  virtual int launch(cudaStream_t stream) = 0;
  virtual ~esf_op() {}
};
